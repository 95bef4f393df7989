"""Data-parallel training over one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The reference is single-device (`train_ssl_mae.py:132`); this adds what SURVEY.md §8(e) specifies: the batch is sharded
across ranks, weights are replicated, and gradients are averaged with a bucketed allreduce that is launched on a side
stream as soon as the last gradient of a bucket has been accumulated, so the collective overlaps the rest of backward.
Buckets are flat fp32 buffers; every `param.grad` is a view into its bucket, so the weight-gradient kernels and autograd
accumulate straight into the communication buffer (no copy-in / copy-out). BatchNorm statistics stay rank-local, as in
the (SyncBN-free) reference.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


class _Bucket:
    def __init__(self, params: List[nn.Parameter], device, dtype=torch.float32):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat = torch.zeros(self.numel, device=device, dtype=dtype)
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.pending = len(params)
        self.fired = {}
        self.work = None
        self.event: Optional[torch.cuda.Event] = None


class GradientBucketer(nn.Module):
    """Wraps a module: `forward` is the wrapped forward; after `loss.backward()` call `finish_gradient_sync()`.

    bucket_mb sizes buckets for launch latency, not link count (NVSwitch gives every peer full bandwidth).
    Parameters that receive no gradient in a step (e.g. `encoder.stages.3.*` of TinyVideoMAE) must be excluded with
    `exclude` or they would stall their bucket; by default parameters whose name matches `exclude` are left out.
    """

    def __init__(self, module: nn.Module, process_group=None, bucket_mb: float = 32.0, exclude=("encoder.stages.3.",),
                 broadcast: bool = True):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad and not any(n.startswith(e) for e in exclude)]
        self.device = named[0][1].device
        self.on_cuda = self.device.type == "cuda"
        if broadcast and self.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        # reverse registration order ~ the order in which backward finishes the gradients
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets: List[_Bucket] = []
        cur: List[nn.Parameter] = []
        size = 0
        for _, p in reversed(named):
            cur.append(p)
            size += p.numel()
            if size >= cap:
                self.buckets.append(_Bucket(cur, self.device))
                cur, size = [], 0
        if cur:
            self.buckets.append(_Bucket(cur, self.device))
        self._bucket_of = {}
        for b in self.buckets:
            for p in b.params:
                self._bucket_of[p] = b
                p.register_post_accumulate_grad_hook(self._on_grad_ready)
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.on_cuda else None
        self.launched = 0
        # defer_comm: backward only fills the buckets (nothing is sent from the hooks); the caller runs `allreduce_now()`
        # afterwards. This is the mode for a forward+backward that is replayed from a CUDA graph (graph_step.py): the
        # collectives stay outside the graph.
        self.defer_comm = False
        if self.on_cuda:
            # fused blocks accumulate straight into the bucket views of THESE parameters and report completion per parameter
            from . import ops

            ops.enable_direct_grad([p for b in self.buckets for p in b.params], self._on_direct_grad)

    # ------------------------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = False):  # keep the grad views alive
        for b in self.buckets:
            b.flat.zero_()
            b.pending = len(b.params)
            b.fired = {}
            b.work = None

    def _on_grad_ready(self, p: nn.Parameter):
        b = self._bucket_of[p]
        if b.fired.get(id(p), 0):
            return  # already signalled this step (a fused block reported it directly and autograd's hook ran as well)
        b.fired[id(p)] = 1
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _on_direct_grad(self, p):
        if p in self._bucket_of:
            self._on_grad_ready(p)

    def _launch(self, b: _Bucket):
        self.launched += 1
        if self.world == 1 or self.defer_comm:
            return
        if self.on_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group)
                b.event = torch.cuda.Event()
                b.event.record(self.comm_stream)
        else:
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish_gradient_sync(self):
        """Make the averaged gradients visible to the compute stream (call once after backward)."""
        for b in self.buckets:
            if b.pending != 0:
                names = {id(p): n for n, p in self.module.named_parameters()}
                missing = [names.get(id(p), "?") for p in b.params if b.fired.get(id(p), 0) == 0]
                raise RuntimeError(f"a gradient bucket was not completed by backward: a parameter received no gradient "
                                   f"(missing {missing[:8]}); exclude parameters that never get a gradient via "
                                   f"GradientBucketer(exclude=...)")
            if self.world == 1 or self.defer_comm:
                continue
            if self.on_cuda:
                torch.cuda.current_stream(self.device).wait_event(b.event)
            else:
                b.work.wait()
                b.flat.div_(self.world)

    def allreduce_now(self):
        """Average every bucket across ranks on the current stream (the `defer_comm` counterpart of the hook-driven path)."""
        if self.world == 1:
            return
        for b in self.buckets:
            if self.on_cuda:
                dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group)
                b.flat.div_(self.world)
