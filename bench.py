#!/usr/bin/env python
"""bench.py — VideoMAE ViT-S pretraining step throughput (clips/s), BASELINE.json's metric.

  python bench.py --gpus N --steps K --warmup W                 our sm_100a path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W  the CPU oracle port of the same step on the host cores

A "step" is one full training step of the north-star configuration (BASELINE config 2): ViT-S/16 VideoMAE,
16x224x224 clips, tubelet 2x16x16, tube mask 0.9, per-GPU batch 64, bf16 GEMM/attention operands with fp32 accumulation:
forward, fused masked-MSE loss, backward, (N>1: bucketed gradient allreduce overlapped with backward), AdamW update.
Synthetic clips and random-init weights (there is no dataset/checkpoint access). One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "videomae_vit_s_pretrain_clips_per_sec"
UNIT = "clips/s"
PER_GPU_BATCH = 64
WORKLOAD = "ViT-S/16 VideoMAE pretrain step, 16x224x224, tubelet 2x16x16, mask 0.9 (BASELINE configs[1]; configs[2] for N>1)"
STEP_GFLOP_PER_CLIP = 139.96  # 3 x 46.65 GFLOP forward


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.file = index, None, None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                          str(self.index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        rows = [r.strip().split(", ") for r in self.file.read().strip().splitlines() if r.strip()]
        self.file.close()
        os.unlink(self.file.name)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (its NUMA node), so that the pinned staging buffers are
    first-touched there and the H2D copies of 8 ranks do not all cross one socket. Returns a short description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        maskw = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(maskw) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed) or sorted(allowed)
        os.sched_setaffinity(0, cpus)
        return {"cpus": f"{cpus[0]}-{cpus[-1]}", "n": len(cpus), "all": sorted(allowed)}
    except Exception as exc:   # containers without NVML / cpusets: leave the affinity alone
        return {"cpus": None, "error": type(exc).__name__}


def make_inputs(batch, cfg, seed):
    from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask, visible_indices

    T, S = cfg["dataset"]["clip_len"], cfg["dataset"]["image_size"]
    pt, ps, _ = cfg["model"]["tubelet"]
    g = torch.Generator().manual_seed(seed)
    clip = torch.randn(batch, 3, T, S, S, generator=g)
    torch.manual_seed(42 + seed)  # rank-seeded masks
    mask = get_tube_mask(batch, T // pt, (S // ps) ** 2, cfg["ssl"]["mask_ratio"])
    return clip, mask, visible_indices(mask)


# ======================================================================================================================
# reference arm: the CPU oracle port of the same step on the host cores
# ======================================================================================================================
def cpu_reference_step_factory(cfg, batch, threads):
    from oracle import restatement as R
    from ssl_vit_video_analytics_b200.models import VideoMAEViT

    torch.set_num_threads(threads)
    R.USE_SDPA = True   # the reference's own attention call (tiny_vit.py:103): the fastest CPU path PyTorch has
    torch.manual_seed(0)
    model = VideoMAEViT(cfg)  # parameter container only (constructed on CPU; its forward is never called here)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=5e-4, weight_decay=0.05)
    clip, mask, _ = make_inputs(batch, cfg, seed=0)
    m = cfg["model"]

    def step():
        opt.zero_grad(set_to_none=True)
        pred = R.videomae_vit_forward(sd, clip, mask, m["depth"], m["num_heads"], m["decoder_depth"], m["decoder_num_heads"], tuple(m["tubelet"]))
        loss = R.masked_mse(pred, clip, mask, m["tubelet"][0], m["tubelet"][1], True)
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def time_cpu_reference_model_a(steps, warmup, threads):
    """BASELINE.md §3 row "CPU-ref": the reference's step body (src/train_ssl_mae.py:71-89) on TinyVideoMAE at C1 (B=2, T=8, 112x112,
    tube mask 0.9), fp32, through the oracle restatement (pinned against the live reference by tests/golden), + AdamW(5e-4, 0.05)."""
    from oracle import restatement as R
    from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask
    from ssl_vit_video_analytics_b200.models import TinyVideoMAE, tiny_vit_21m_variant

    torch.set_num_threads(threads)
    R.USE_SDPA = True
    cfg = {"dataset": {"clip_len": 8, "image_size": 112}, "model": {"decoder_embed_dim": 384, "decoder_depth": 4, "decoder_num_heads": 6, "decoder_dropout": 0.0}}
    torch.manual_seed(0)
    model = TinyVideoMAE(tiny_vit_21m_variant(img_size=112, use_checkpoint=False, drop_path_rate=0.0), cfg)   # parameter container only
    sd0 = R.perturb_state_dict(model.state_dict(), seed=1234)
    sd = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and "running" not in k else v) for k, v in sd0.items()}
    params = [v for k, v in sd.items() if torch.is_tensor(v) and v.requires_grad and not k.startswith("encoder.stages.3.")]
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=0.05)
    g = torch.Generator().manual_seed(0)
    clip = torch.randn(2, 3, 8, 112, 112, generator=g)

    def step():
        mask = get_tube_mask(2, 8, 196, 0.9)
        opt.zero_grad(set_to_none=True)
        pred = R.tiny_video_mae_forward(sd, clip, mask, 4, 6)
        loss = R.masked_mse(pred, clip, mask, 1, 8, True)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return 2 * steps / dt, dt / steps * 1e3


def time_model_a(dev, steps, warmup, peaks):
    """Secondary line: the DROP-IN model (reference TinyVideoMAE on the TinyViT-21M encoder, the path with a live reference) at the
    reference's own configs/ssl_mae.yaml shapes (B=40, T=16, 112x112, patch 8, mask 0.9), perturbed init,
    forward + fused loss + backward + FusedAdamW on our kernels, eager launches. 740.3 GFLOP per clip and step."""
    from ssl_vit_video_analytics_b200 import ops
    from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask
    from ssl_vit_video_analytics_b200.models import TinyVideoMAE, tiny_vit_21m_variant
    from ssl_vit_video_analytics_b200.optim import FusedAdamW
    from ssl_vit_video_analytics_b200.parallel import GradientBucketer
    from ssl_vit_video_analytics_b200.train_ssl_mae import mae_step_loss

    B = 40
    tcfg = {"dataset": {"clip_len": 16, "image_size": 112},
            "model": {"decoder_embed_dim": 384, "decoder_depth": 4, "decoder_num_heads": 6, "decoder_dropout": 0.0}}
    torch.manual_seed(0)
    model = TinyVideoMAE(tiny_vit_21m_variant(img_size=112, use_checkpoint=False, drop_path_rate=0.0), tcfg)
    gen = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".bn.weight"):
                p.copy_(torch.rand(p.shape, generator=gen) + 0.5)
            elif name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.02)
    model = model.to(dev).train()
    ddp = GradientBucketer(model, bucket_mb=32.0, exclude=("encoder.stages.3.",), broadcast=False)
    ddp.world = 1
    opt = FusedAdamW([p for n, p in model.named_parameters() if not n.startswith("encoder.stages.3.")], lr=5e-4, weight_decay=0.05)
    g = torch.Generator().manual_seed(0)
    clip = torch.randn(B, 3, 16, 112, 112, generator=g).to(dev)
    torch.manual_seed(42)
    mask = get_tube_mask(B, 16, 196, 0.9).to(dev)

    def step():
        ddp.zero_grad()
        loss = mae_step_loss(model, clip, mask, True)[0]
        loss.backward()
        ddp.finish_gradient_sync()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    val = B / (ms / 1e3)
    out = {"metric": "tinyvideomae_pretrain_clips_per_sec", "value": val, "unit": UNIT, "ms_per_step": ms, "steps": steps, "per_gpu_batch": B,
           "workload": "drop-in TinyVideoMAE (TinyViT-21M encoder), 16x112x112, patch 8, mask 0.9 (configs/ssl_mae.yaml), use_checkpoint=False, drop-path/dropout 0, perturbed init, eager launches",
           "step_gflop_per_clip": 740.3, "frac_of_tensor_roofline": (B * 740.3 / 1e3) / (ms / 1e3) / peaks["bf16_tflops_sustained"],
           "final_loss": float(loss.detach()), "gpu_launches_per_step": (ops.launch_count() - n0) / steps}
    del model, ddp, opt
    torch.cuda.empty_cache()
    return out


def time_gpu_eager(cfg, batch, steps, warmup, dev):
    """BASELINE.md §3 row "GPU-eager-ref": the SAME model-B step written in plain PyTorch (the oracle restatement: nn.functional
    conv3d / linear / layer_norm / scaled_dot_product_attention / gelu, i.e. cuDNN + cuBLAS + flash-SDPA), bf16 autocast, same batch,
    torch's fused AdamW, on the same B200. None of our kernels run here. This is the library baseline our kernels have to beat."""
    from oracle import restatement as R
    from ssl_vit_video_analytics_b200.models import VideoMAEViT

    R.USE_SDPA = True
    torch.manual_seed(0)
    model = VideoMAEViT(cfg)   # parameter container only
    sd = {k: v.detach().to(dev).requires_grad_(True) for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=5e-4, weight_decay=0.05, fused=True)
    clip, mask, _ = make_inputs(batch, cfg, seed=0)
    clip, mask = clip.to(dev), mask.to(dev)
    m = cfg["model"]

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = R.videomae_vit_forward(sd, clip, mask, m["depth"], m["num_heads"], m["decoder_depth"], m["decoder_num_heads"], tuple(m["tubelet"]))
        loss = R.masked_mse(pred, clip, mask, m["tubelet"][0], m["tubelet"][1], True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "final_loss": float(loss.detach()),
           "what": f"oracle/restatement.py videomae_vit_forward on cuda, bf16 autocast, F.scaled_dot_product_attention + cuBLAS + cuDNN, "
                   f"torch.optim.AdamW(fused=True), batch {batch}, {steps} steps after {warmup} warm-up; embeds all tubelets then gathers (eager has no fused gather)"}
    del sd, opt
    torch.cuda.empty_cache()
    return out


def time_cpu_reference(cfg, batch, steps, warmup, threads):
    step = cpu_reference_step_factory(cfg, batch, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 2
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    val, ms = time_cpu_reference(cfg, batch, steps, warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_step_batch": batch, "device": "cpu"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"oracle/restatement.py videomae_vit_forward + masked MSE + backward + AdamW, fp32, batch {batch} clips/step, "
                                   f"{steps} steps after {warmup} warm-up (capped from the requested {args.steps}/{args.warmup})"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ======================================================================================================================
# our arm
# ======================================================================================================================
def run_ours(args, cfg):
    import torch.distributed as dist

    from ssl_vit_video_analytics_b200 import ops
    from ssl_vit_video_analytics_b200.graph_step import GraphedTrainStep, use_side_stream
    from ssl_vit_video_analytics_b200.models import VideoMAEViT
    from ssl_vit_video_analytics_b200.optim import FusedAdamW
    from ssl_vit_video_analytics_b200.parallel import GradientBucketer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a B200: the sm_100a kernels have no fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)   # before any pinned allocation: first touch then lands on the GPU's own NUMA node
    if world > 1:
        # --nccl-channels N caps NCCL_MAX_NCHANNELS (A/B knob). Measured at 8 GPUs: default 17.80 ms/step, 4 channels 18.70, 2 channels
        # 20.9 — the collective of the LAST bucket (it completes with the end of backward) is on the critical path, so a slower
        # allreduce costs more than the SMs a faster one borrows from the persistent kernels. Default: NCCL's own choice.
        if args.nccl_channels > 0:
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(args.nccl_channels))
        dist.init_process_group("nccl", device_id=dev)
    use_graph = args.graph != "off"
    if use_graph:
        # everything runs on a non-default stream from the start: autograd binds each gradient accumulator to the stream it first
        # saw, and a capture may not synchronise with the legacy default stream
        use_side_stream(dev)
    torch.manual_seed(0)
    if args.workload == "tinyvit":
        # the drop-in TinyVideoMAE at the reference's own configs/ssl_mae.yaml shapes (reported next to the headline)
        from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask
        from ssl_vit_video_analytics_b200.models import TinyVideoMAE, tiny_vit_21m_variant
        from ssl_vit_video_analytics_b200.train_ssl_mae import mae_step_loss

        B = 40 if args.batch is None else args.batch
        tcfg = {"dataset": {"clip_len": 16, "image_size": 112},
                "model": {"decoder_embed_dim": 384, "decoder_depth": 4, "decoder_num_heads": 6, "decoder_dropout": 0.0}}
        model = TinyVideoMAE(tiny_vit_21m_variant(img_size=112, use_checkpoint=False, drop_path_rate=0.0), tcfg)
        # "perturbed init" (SURVEY.md §0 fact 6): at the reference's literal init every stride-2 MBConv ends in a BatchNorm
        # with gamma = 0, the encoder output is identically zero and the arithmetic is mostly on zeros
        gen = torch.Generator().manual_seed(1234)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith(".bn.weight"):
                    p.copy_(torch.rand(p.shape, generator=gen) + 0.5)
                elif name.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.02)
        model = model.to(dev).train()
        ddp = GradientBucketer(model, bucket_mb=32.0, exclude=("encoder.stages.3.",))
        g = torch.Generator().manual_seed(rank)
        clip_h = torch.randn(B, 3, 16, 112, 112, generator=g)
        torch.manual_seed(42 + rank)
        mask_d = get_tube_mask(B, 16, 196, 0.9).to(dev)

        def loss_fn(clip):
            return mae_step_loss(model, clip, mask_d, True)[0]
    else:
        B = PER_GPU_BATCH if args.batch is None else args.batch
        model = VideoMAEViT(cfg).to(dev).train()
        ddp = GradientBucketer(model, bucket_mb=args.bucket_mb, exclude=())
        clip_h, mask_h, vis_h = make_inputs(B, cfg, seed=rank)
        mask_d, vis_d = mask_h.to(dev), vis_h.to(dev)
        # every input buffer owns its (static) mask tensors: the e2e loop draws a fresh tube mask per step on the host, like
        # train_ssl_mae.py:71, and uploads it next to the clip
        masks_of = {}

        def loss_fn(clip):
            md, vd = masks_of.get(clip.data_ptr(), (mask_d, vis_d))
            return model.training_loss(clip, md, vd)
    if args.optimizer == "fused":
        opt = FusedAdamW([p for p in model.parameters()], lr=5e-4, weight_decay=0.05)   # row a18: one launch, device-side hyper-parameters
    else:   # A/B: PyTorch's multi-tensor AdamW
        opt = torch.optim.AdamW([p for p in model.parameters()], lr=5e-4, weight_decay=0.05, fused=True, capturable=use_graph)
    clip_pinned = [clip_h.pin_memory(), clip_h.clone().pin_memory()]
    clip_d = clip_h.to(dev)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def eager_step(clip):
        ddp.zero_grad()
        loss = loss_fn(clip)
        loss.backward()
        ddp.finish_gradient_sync()
        opt.step()
        return loss

    # The whole step (zero-grad, forward, loss, backward, gradient sync, AdamW) is captured once per input buffer into a CUDA
    # graph and replayed: ~400 launches per step cost the host 12 ms of the 18 ms a step takes on the GPU, which leaves no
    # slack for the copy/readback bookkeeping of the e2e loop. The graphs share one memory pool (they never run concurrently
    # and hand nothing to each other but parameters, gradients and optimizer state, which live outside the pool).
    # N > 1: the NCCL collectives stay outside the graphs (capturing the hook-driven bucket allreduce deadlocked when tried):
    # graph 1 = zero-grad + forward + loss + backward into the gradient buckets, then the bucket allreduces eagerly on the same
    # stream, then graph 2 = AdamW. The allreduce (119 MB over NVLink) is no longer hidden behind backward; what it costs is
    # less than what the host-side launch work cost.
    def fwd_bwd(clip):
        ddp.zero_grad()
        loss = loss_fn(clip)
        loss.backward()
        ddp.finish_gradient_sync()
        return loss

    # --allreduce captured (default for N > 1): the whole step is ONE graph also under torchrun — the bucket allreduces are
    # launched by the gradient hooks on the communication stream, which forks from the capture stream at "bucket complete" and
    # joins before AdamW, so NCCL runs underneath the rest of backward inside the replayed graph.
    # --allreduce between: graph 1 (forward+backward), eager allreduces, graph 2 (AdamW): the round-1 arrangement, exposed collective.
    ar_captured = world > 1 and args.allreduce == "captured"
    graphed = GraphedTrainStep(eager_step if (world == 1 or ar_captured) else fwd_bwd)
    opt_graph = torch.cuda.CUDAGraph() if (use_graph and world > 1 and not ar_captured) else None

    def step(clip):
        if not use_graph:
            return eager_step(clip)
        loss = graphed.replay(clip)
        if opt_graph is not None:
            ddp.allreduce_now()
            opt_graph.replay()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (eager: also initialises everything a capture must not do for the first time), then the captures
    for _ in range(args.warmup):
        loss = eager_step(clip_d)
    barrier()
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(clip_d), torch.empty_like(clip_d)]
    refresh_masks = args.workload == "vit_s" and os.environ.get("VMAE_BENCH_STATIC_MASK", "0") != "1"
    if refresh_masks:
        from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask, visible_indices

        mask_pinned = [mask_h.clone().pin_memory() for _ in range(2)]
        vis_pinned = [vis_h.clone().pin_memory() for _ in range(2)]
        for b in bufs:
            masks_of[b.data_ptr()] = (mask_d.clone(), vis_d.clone())
        Tp_, Lp_ = mask_h.shape[1], mask_h.shape[2]
    if use_graph:
        bufs[0].copy_(clip_d)
        bufs[1].copy_(clip_d)
        ddp.defer_comm = world > 1 and not ar_captured
        try:
            for t in (clip_d, bufs[0], bufs[1]):
                graphed.capture(t)
            if opt_graph is not None:
                with torch.cuda.graph(opt_graph, pool=graphed.pool()):
                    opt.step()
        except Exception as exc:   # a failed capture leaves no graph behind: measure the eager step instead, and say so
            print(f"bench.py: CUDA-graph capture failed ({type(exc).__name__}: {exc}); timing eager launches", file=sys.stderr)
            use_graph = False
            ddp.defer_comm = False
            torch.cuda.synchronize()
        for _ in range(args.warmup):
            loss = step(clip_d)
        barrier()

    # ---- value: inputs resident in HBM (the 616 MB clip + ~20 GB of activations exceed the 126 MB L2 every step)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(clip_d)
    e1.record()
    barrier()
    launches = ops.launch_count()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(loss.detach())

    # ---- e2e: pinned-host clip -> device every step, loss read back every step; the copy of step i+1 is issued on a
    #      side stream while step i computes (the loader-prefetch pattern of train_ssl_mae.py:67 non_blocking copies)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    mask_bytes = 0

    def issue_copy(i):
        nonlocal mask_bytes
        if refresh_masks:   # a fresh tube mask per step, drawn on the host as the reference does (train_ssl_mae.py:71)
            ready[i % 2].synchronize()        # the pinned mask staging of this slot is free once its previous H2D copy (two steps ago) is done
            mh = get_tube_mask(B, Tp_, Lp_, cfg["ssl"]["mask_ratio"])
            mask_pinned[i % 2].copy_(mh)
            vis_pinned[i % 2].copy_(visible_indices(mh))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            bufs[i % 2].copy_(clip_pinned[i % 2], non_blocking=True)
            if refresh_masks:
                md, vd = masks_of[bufs[i % 2].data_ptr()]
                md.copy_(mask_pinned[i % 2], non_blocking=True)
                vd.copy_(vis_pinned[i % 2], non_blocking=True)
                mask_bytes = mask_pinned[0].numel() * mask_pinned[0].element_size() + vis_pinned[0].numel() * vis_pinned[0].element_size()
            ready[i % 2].record(copy_stream)

    loss_hosts = [torch.zeros((), dtype=torch.float32).pin_memory(), torch.zeros((), dtype=torch.float32).pin_memory()]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        """Every step: H2D copy of its clip batch (issued one step ahead on the copy stream) and a D2H read of its loss. The loss
        of step i is copied to pinned memory asynchronously and read by the host while step i+1 is already enqueued, so the host
        never drains the GPU queue (a blocking .item() per step idles the GPU while the next step's first kernels are launched)."""
        cur = torch.cuda.current_stream(dev)
        for b in range(2):
            consumed[b].record(cur)
        issue_copy(0)
        last = 0.0
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)
            cur.wait_event(ready[i % 2])
            loss = step(bufs[i % 2])
            consumed[i % 2].record(cur)
            loss_hosts[i % 2].copy_(loss.detach(), non_blocking=True)
            loss_ready[i % 2].record(cur)
            if i > 0:
                loss_ready[(i - 1) % 2].synchronize()
                last = float(loss_hosts[(i - 1) % 2])
        loss_ready[(n - 1) % 2].synchronize()
        last = float(loss_hosts[(n - 1) % 2])
        return last

    # the H2D copy of one clip batch on its own (no compute): the floor the e2e step cannot go below
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        h0.record(copy_stream)
        for i in range(4):
            bufs[i % 2].copy_(clip_pinned[i % 2], non_blocking=True)
        h1.record(copy_stream)
    torch.cuda.synchronize()
    h2d_ms_alone = h0.elapsed_time(h1) / 4.0

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    _ = time.perf_counter() - t0

    # ---- the same loop fed with decoded uint8 frames (`vmae_frames_to_clip` normalises on the GPU): a quarter of
    #      the H2D bytes. Reported next to the headline e2e, not instead of it: the reference's step takes the float clip.
    e2e_u8 = None
    if args.workload == "vit_s":
        T_, S_ = cfg["dataset"]["clip_len"], cfg["dataset"]["image_size"]
        gen8 = torch.Generator().manual_seed(7 + rank)
        frames_pinned = [torch.randint(0, 256, (B, T_, S_, S_, 3), dtype=torch.uint8, generator=gen8).pin_memory() for _ in range(2)]
        frames_d = [torch.empty((B, T_, S_, S_, 3), dtype=torch.uint8, device=dev) for _ in range(2)]

        def issue_copy_u8(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                frames_d[i % 2].copy_(frames_pinned[i % 2], non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_u8_loop(n):
            cur = torch.cuda.current_stream(dev)
            for b in range(2):
                consumed[b].record(cur)
            issue_copy_u8(0)
            last = 0.0
            for i in range(n):
                if i + 1 < n:
                    issue_copy_u8(i + 1)
                cur.wait_event(ready[i % 2])
                ops.frames_to_clip(frames_d[i % 2], out=bufs[i % 2])
                consumed[i % 2].record(cur)          # the uint8 staging buffer is free again once it has been converted
                loss = step(bufs[i % 2])
                loss_hosts[i % 2].copy_(loss.detach(), non_blocking=True)
                loss_ready[i % 2].record(cur)
                if i > 0:
                    loss_ready[(i - 1) % 2].synchronize()
                    last = float(loss_hosts[(i - 1) % 2])
            loss_ready[(n - 1) % 2].synchronize()
            return float(loss_hosts[(n - 1) % 2])

        e2e_u8_loop(2)
        barrier()
        e0.record()
        e2e_u8_loop(args.steps)
        e1.record()
        barrier()
        u8_ms = max_over_ranks(e0.elapsed_time(e1))
        e2e_u8 = {"value": world * B * args.steps / (u8_ms / 1e3), "unit": UNIT, "ms_per_step": u8_ms / args.steps,
                  "h2d_bytes_per_step": frames_pinned[0].numel(), "d2h_bytes_per_step": 4,
                  "note": "same loop fed with pinned uint8 frames [B,T,H,W,3]; ToTensor + Normalize + channel order on the GPU (vmae_frames_to_clip)"}

    # ---- roofline leg: the same step with CUDA events around every launch of ours (separate region: events perturb)
    # Every profiled step is queued behind a ~30 ms spin kernel: the host needs ~12 ms to enqueue the ~400 eager launches of a
    # step, so without the head start the GPU catches up with the host and the event pairs also measure the host's launch gaps.
    ddp.defer_comm = False
    ops.profile_start()
    for _ in range(2):
        torch.cuda._sleep(int(30e-3 * 1.9e9))
        eager_step(clip_d)
    prof = ops.profile_stop()
    for v in prof.values():
        v["ms"] /= 2.0
        v["launches"] /= 2.0
        v["work"] /= 2.0

    def leave():
        """Multi-rank exit. With NCCL kernels captured in CUDA graphs, destroying the communicator while the graphs are alive hangs
        (observed: rank 1 stuck in destroy_process_group until torchrun's timeout), so: everybody syncs, the graphs are dropped,
        and the process leaves without the NCCL teardown (the OS reclaims it)."""
        if world == 1:
            return
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)

    if rank != 0:
        leave()
        return
    peaks = load_peaks()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    if args.workload == "tinyvit":
        workload, step_gflop = "drop-in TinyVideoMAE (TinyViT-21M encoder), 16x112x112, patch 8, mask 0.9 (configs/ssl_mae.yaml)", 740.3
    else:
        workload, step_gflop = WORKLOAD, STEP_GFLOP_PER_CLIP
    tensor_classes = ["gemm_fwd", "gemm_dgrad", "gemm_wgrad", "attn_fwd", "attn_bwd"]
    dom = max((k for k in tensor_classes if k in prof), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    achieved = d["work"] / (d["ms"] / 1e3) / 1e12
    traffic = None
    try:   # DRAM bytes per launch of that kernel class from the committed `ncu --set full` capture (same shapes)
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(dom) if args.workload == "vit_s" else None
    except OSError:
        pass
    roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic, "peak_source": peaks["source"] + ", sustained bf16",
                "launches_per_step": d["launches"], "ms_per_step": d["ms"],
                "step_frac_of_tensor_roofline": (B * step_gflop / 1e3) / (ms_step / 1e3) / peaks["bf16_tflops_sustained"],
                "by_kernel_class": {k: {"ms_per_step": round(v["ms"], 3), "launches_per_step": v["launches"],
                                        ("tflops" if k in tensor_classes else "gbs"): round(v["work"] / (v["ms"] / 1e3) / (1e12 if k in tensor_classes else 1e9), 1)}
                                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": world * B, "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "l2_policy": "inputs larger than L2 (clip batch of hundreds of MB, >10 GB activations per step)", "optimizer": ("optim.FusedAdamW (vmae_adamw_multi)" if args.optimizer == "fused" else "torch.optim.AdamW(fused)"),
                   "cuda_graph": use_graph, "grad_allreduce": ("none" if world == 1 else (("bucketed NCCL allreduce on a forked stream inside the captured step, overlapped with backward" if ar_captured else "after backward, between two graphs") if use_graph else "bucketed, hook-driven, overlapped with backward (eager launches)")), "final_loss": final_loss},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": clip_h.numel() * 4 + mask_bytes, "d2h_bytes_per_step": 4, "numa_binding": {k: v for k, v in numa.items() if k != "all"},
                "mask": ("fresh tube mask drawn on the host and uploaded every step" if refresh_masks else "static"),
                "ms_per_step": e2e_ms / args.steps, "h2d_ms_alone": h2d_ms_alone, "uint8_frames_input": e2e_u8, "note": "pinned fp32 clip batch copied H2D every step (double-buffered on a copy stream); the loss of every step is copied D2H to pinned memory and read by the host one step later, inside the timed region"},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if world == 1 and not args.no_model_a and args.workload == "vit_s":
        try:
            torch.cuda.empty_cache()
            line["model_a"] = time_model_a(dev, 5, 3, peaks)
        except Exception as ex:
            line["model_a"] = {"value": None, "unit": UNIT, "workload": f"failed: {ex!r}"}
    if world == 1 and not args.no_gpu_eager and args.workload == "vit_s":
        try:
            torch.cuda.empty_cache()
            ge = time_gpu_eager(cfg, B, max(3, min(args.steps, 10)), 3, dev)
            ge["ours_over_eager"] = value / ge["value"]
            line["gpu_eager_baseline"] = ge
        except Exception as ex:
            line["gpu_eager_baseline"] = {"value": None, "unit": UNIT, "what": f"failed: {ex!r}"}
    if world == 1 and not args.no_cpu_baseline and args.workload == "vit_s":
        if numa.get("all"):
            os.sched_setaffinity(0, numa["all"])   # the CPU legs use every host core again
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        try:
            va, msa = time_cpu_reference_model_a(3, 1, threads)
            line["cpu_baseline_model_a"] = {"value": va, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": msa,
                                            "sample": "oracle restatement of the reference step (train_ssl_mae.py:71-89) on TinyVideoMAE at C1: B=2, T=8, 112x112, mask 0.9, fp32, AdamW; 3 steps after 1 warm-up"}
        except Exception as ex:
            line["cpu_baseline_model_a"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"failed: {ex!r}"}
        try:
            v, ms = time_cpu_reference(cfg, 2, 3, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "oracle/restatement.py (fp32 PyTorch CPU port of the same step incl. AdamW), batch 2 clips/step, 3 steps after 1 warm-up"}
        except Exception as ex:  # the baseline must never take the measured line down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"failed: {ex!r}"}
    print(json.dumps(line), flush=True)
    leave()


# ======================================================================================================================
# BASELINE config 5: FedAvg simulation, 5 clients mapped across the GPUs of one box (src/federated/fed_loop.py:65-147)
# ======================================================================================================================
def run_fedavg(args, cfg):
    """One round = every client runs `--local-steps` MAE pretraining steps on its own GPU from the current global model, then the
    models are aggregated with the weighted allreduce `federated.fed_loop.fedavg_allreduce` (pre-scale by w_k / sum w, NCCL
    allreduce(sum); clients = ranks 0..4, weights = synthetic dataset sizes as in run_federated.py:288; further ranks are idle and
    contribute weight 0). The reference does this sequentially on one device and aggregates on the CPU (fed_loop.py:109-119)."""
    import torch.distributed as dist

    from ssl_vit_video_analytics_b200 import ops
    from ssl_vit_video_analytics_b200.federated.fed_loop import fedavg_allreduce
    from ssl_vit_video_analytics_b200.graph_step import GraphedTrainStep, use_side_stream
    from ssl_vit_video_analytics_b200.models import VideoMAEViT
    from ssl_vit_video_analytics_b200.optim import FusedAdamW
    from ssl_vit_video_analytics_b200.parallel import GradientBucketer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    use_side_stream(dev)
    sizes = [1907, 1893, 1921, 1880, 1936]                       # client dataset sizes
    n_clients = min(len(sizes), world)
    is_client = rank < n_clients
    weight = float(sizes[rank]) if is_client else 0.0
    B = PER_GPU_BATCH if args.batch is None else args.batch
    torch.manual_seed(0)
    model = VideoMAEViT(cfg).to(dev).train()
    ddp = GradientBucketer(model, bucket_mb=32.0, exclude=())    # broadcasts rank 0's weights = the initial global model
    ddp.world = 1                                                # clients train independently: no gradient exchange inside a round
    opt = FusedAdamW(list(model.parameters()), lr=5e-4, weight_decay=0.05)
    clip_h, mask_h, vis_h = make_inputs(B, cfg, seed=rank)
    clip_d, mask_d, vis_d = clip_h.to(dev), mask_h.to(dev), vis_h.to(dev)

    def local_step(clip):
        ddp.zero_grad()
        loss = model.training_loss(clip, mask_d, vis_d)
        loss.backward()
        ddp.finish_gradient_sync()
        opt.step()
        return loss

    for _ in range(3):
        local_step(clip_d)
    graphed = GraphedTrainStep(local_step)
    graphed.capture(clip_d)

    def one_round():
        if is_client:
            for _ in range(args.local_steps):
                graphed.replay(clip_d)
        fedavg_allreduce(model, weight)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 2)):
        one_round()
    barrier()
    ops.reset_launch_count()
    e0, e1, ea = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rounds = max(1, args.steps // max(1, args.local_steps))
    agg_ms = 0.0
    e0.record()
    for _ in range(rounds):
        if is_client:
            for _ in range(args.local_steps):
                graphed.replay(clip_d)
        ea.record()
        fedavg_allreduce(model, weight)
        e1.record()
        e1.synchronize()
        agg_ms += ea.elapsed_time(e1)
    barrier()
    total_ms = e0.elapsed_time(e1)
    if not is_client:
        agg_ms = 0.0   # an idle rank enters the allreduce at once and waits there for the clients' local steps: not aggregation time
    t = torch.tensor([total_ms, agg_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, agg_ms = float(t[0]), float(t[1])
    check = torch.cat([p.detach().flatten()[:64].double() for p in list(model.parameters())[:4]]).sum().reshape(1)
    lo, hi = check.clone(), check.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        clips = n_clients * B * args.local_steps * rounds
        n_floats = sum(v.numel() for v in model.state_dict().values() if torch.is_floating_point(v))
        line = {"metric": "fedavg_round_clips_per_sec", "value": clips / (total_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": rounds * args.local_steps,
                "warmup": args.warmup, "ms_per_step": total_ms / (rounds * args.local_steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "FedAvg simulation (BASELINE configs[4]): 5 clients on the first 5 GPUs, local ViT-S VideoMAE pretraining steps + weighted-allreduce aggregation per round",
                           "clients": n_clients, "client_weights": sizes[:n_clients], "local_steps": args.local_steps, "rounds": rounds, "per_gpu_batch": B,
                           "idle_ranks": max(0, world - n_clients)},
                "fedavg": {"round_ms": total_ms / rounds, "aggregate_ms": agg_ms / rounds, "state_floats": n_floats,
                           "aggregate_gbs_per_gpu": 2 * 4 * n_floats / (agg_ms / rounds / 1e3) / 1e9 if agg_ms > 0 else None,
                           "replicas_identical_after_round": bool(float(lo) == float(hi))},
                "gpu_launches": ops.launch_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 64 = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-model-a", action="store_true", help="skip the secondary line of the drop-in TinyVideoMAE (configs/ssl_mae.yaml shapes)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the plain-PyTorch (cuBLAS/cuDNN/SDPA) comparator leg")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"], help="fused = optim.FusedAdamW (default); torch = torch.optim.AdamW(fused=True), for A/B runs")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the training step from CUDA graphs (default on; under torchrun the NCCL allreduce runs between two graphs)")
    ap.add_argument("--allreduce", default="captured", choices=["captured", "between"],
                    help="N > 1 with graphs: captured = NCCL bucket allreduces inside the one captured step, overlapped with backward (default); "
                         "between = eager allreduces between a forward+backward graph and an AdamW graph")
    ap.add_argument("--nccl-channels", type=int, default=0, help="cap NCCL_MAX_NCHANNELS (0 = NCCL's default; A/B knob)")
    ap.add_argument("--bucket-mb", type=float, default=32.0, help="gradient bucket size (the last bucket's allreduce is exposed: smaller = shorter tail, more launches)")
    ap.add_argument("--workload", default="vit_s", choices=["vit_s", "tinyvit", "fedavg"],
                    help="vit_s = BASELINE headline (default); tinyvit = the drop-in TinyVideoMAE at configs/ssl_mae.yaml shapes; "
                         "fedavg = BASELINE config 5 (5 clients across the GPUs, local steps + weighted-allreduce aggregation)")
    ap.add_argument("--local-steps", type=int, default=4, help="fedavg: local training steps per client and round")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    from ssl_vit_video_analytics_b200.models.videomae_vit import DEFAULT_CONFIG

    if args.impl == "reference":
        run_reference(args, DEFAULT_CONFIG)
    elif args.workload == "fedavg":
        run_fedavg(args, DEFAULT_CONFIG)
    else:
        run_ours(args, DEFAULT_CONFIG)


if __name__ == "__main__":
    main()
