"""Developer check (GPU): vmae_attn_fwd / vmae_attn_bwd vs an fp32 PyTorch reference. modes: fwd | bwd | perf"""
import ctypes as C
import math
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import _lib as L  # noqa: E402

lib = L.load()
FUSED = "--two-kernel" not in sys.argv   # single-pass backward (head_dim 64) unless --two-kernel is given
dev = torch.device("cuda:0")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ref_attn(qkv, scale):
    # qkv [B, L, 3, H, hd] -> fp32 reference output [B, L, H*hd], lse [B, H, L]
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3)]
    s = (q @ k.transpose(-1, -2)) * scale
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ v
    return o.permute(0, 2, 1, 3).reshape(qkv.shape[0], qkv.shape[1], -1), lse


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)).item()


CASES = [(1, 64, 1, 64), (1, 128, 1, 64), (1, 100, 2, 64), (2, 160, 6, 64), (2, 300, 3, 64), (2, 1568, 6, 64),
         (1, 128, 1, 32), (3, 196, 12, 32), (2, 784, 6, 32)]


def mode_fwd():
    ok = True
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, Lq, H, hd) in CASES:
        qkv = torch.randn(B, Lq, 3, H, hd, device=dev, generator=g).bfloat16()
        scale = 1.0 / math.sqrt(hd)
        out = torch.full((B, Lq, H * hd), float("nan"), device=dev, dtype=torch.bfloat16)
        lse = torch.full((B, H, Lq), float("nan"), device=dev, dtype=torch.float32)
        L.check(lib.vmae_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, Lq, H, hd, scale, 0.0, 0, stream()), "attn_fwd")
        torch.cuda.synchronize()
        ro, rl = ref_attn(qkv, scale)
        e1, e2 = rel(out, ro), (lse - rl).abs().max().item()
        good = e1 < 2e-2 and e2 < 2e-2
        ok &= good
        print(f"{'OK ' if good else 'BAD'} fwd B={B} L={Lq} H={H} hd={hd}: out rel={e1:.3e} lse abs={e2:.3e}", flush=True)
    return ok


def mode_bwd():
    ok = True
    g = torch.Generator(device="cuda").manual_seed(1)
    for (B, Lq, H, hd) in CASES:
        qkv = torch.randn(B, Lq, 3, H, hd, device=dev, generator=g).bfloat16()
        dout = torch.randn(B, Lq, H * hd, device=dev, generator=g).bfloat16()
        scale = 1.0 / math.sqrt(hd)
        out = torch.empty(B, Lq, H * hd, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, Lq, device=dev, dtype=torch.float32)
        L.check(lib.vmae_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, Lq, H, hd, scale, 0.0, 0, stream()), "attn_fwd")
        dqkv = torch.full_like(qkv, float("nan"))
        delta = torch.empty(B, H, Lq, device=dev, dtype=torch.float32)
        dqacc = torch.empty(B, Lq, H * hd, device=dev, dtype=torch.float32) if (FUSED and hd == 64) else None
        L.check(lib.vmae_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dqacc), ptr(dqkv), B, Lq, H, hd, scale, 0.0, 0, stream()), "attn_bwd")
        torch.cuda.synchronize()
        qf = qkv.float().requires_grad_(True)
        ro, _ = ref_attn(qf, scale)
        ro.backward(dout.float())
        errs = [rel(dqkv[:, :, i], qf.grad[:, :, i]) for i in range(3)]
        good = all(e < 3e-2 for e in errs)
        ok &= good
        print(f"{'OK ' if good else 'BAD'} bwd B={B} L={Lq} H={H} hd={hd}: dq rel={errs[0]:.3e} dk rel={errs[1]:.3e} dv rel={errs[2]:.3e}", flush=True)
    return ok


def mode_perf():
    g = torch.Generator(device="cuda").manual_seed(2)
    for (B, Lq, H, hd) in [(64, 1568, 6, 64), (64, 160, 6, 64), (640, 196, 12, 32)]:
        qkv = torch.randn(B, Lq, 3, H, hd, device=dev, generator=g).bfloat16()
        dout = torch.randn(B, Lq, H * hd, device=dev, generator=g).bfloat16()
        scale = 1.0 / math.sqrt(hd)
        out = torch.empty(B, Lq, H * hd, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, Lq, device=dev, dtype=torch.float32)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(B, H, Lq, device=dev, dtype=torch.float32)
        dqacc = torch.empty(B, Lq, H * hd, device=dev, dtype=torch.float32) if (FUSED and hd == 64) else None
        flops_f = 4.0 * B * H * Lq * Lq * hd

        def timeit(fn, iters=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        tf = timeit(lambda: L.check(lib.vmae_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, Lq, H, hd, scale, 0.0, 0, stream()), "f"))
        tb = timeit(lambda: L.check(lib.vmae_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dqacc), ptr(dqkv), B, Lq, H, hd, scale, 0.0, 0, stream()), "b"))
        q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
        qr, kr, vr = [t.detach().clone().requires_grad_(True) for t in (q, k, v)]
        ts = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        o = torch.nn.functional.scaled_dot_product_attention(qr, kr, vr)
        do = torch.randn_like(o)
        tsb = timeit(lambda: torch.autograd.grad(o, (qr, kr, vr), do, retain_graph=True))
        print(f"perf B={B} L={Lq} H={H} hd={hd}: fwd {tf*1e3:.0f} us ({flops_f/tf/1e9:.0f} TF/s) bwd {tb*1e3:.0f} us ({2.5*flops_f/tb/1e9:.0f} TF/s alg) | "
              f"torch SDPA fwd {ts*1e3:.0f} us ({flops_f/ts/1e9:.0f} TF/s) bwd {tsb*1e3:.0f} us ({2.5*flops_f/tsb/1e9:.0f} TF/s)", flush=True)
    return True


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    t0 = time.time()
    ok = {"fwd": mode_fwd, "bwd": mode_bwd, "perf": mode_perf}[mode]()
    print(f"mode {mode}: {'ALL OK' if ok else 'FAILURES'} ({time.time()-t0:.1f}s)")
    sys.exit(0 if ok else 1)
