"""CUDA-graph replay of a whole training step (zero-grad, forward, loss, backward, optimizer).

One VideoMAE ViT-S step is ~400 kernel launches; enqueueing them costs the host ~12 ms of the ~18 ms the step runs on a
B200, so any host-side bookkeeping (input prefetch, loss read-back, logging) eats into the slack. Captured once per input
buffer and replayed, the step costs the host one launch.

Rules the caller has to follow (they are the usual ones of `torch.cuda.graph`):
  * run everything — model construction, `GradientBucketer`, the eager warm-up steps — on a NON-default stream
    (`use_side_stream()` below): autograd binds every gradient accumulator to the stream it first saw, and a capture may not
    synchronise with the legacy default stream;
  * the optimizer must keep everything it reads on the device (`optim.FusedAdamW` does; `torch.optim.AdamW` needs
    `capturable=True`), and its state must exist before the capture (run at least one eager step with the same optimizer object
    first): state created inside a capture is re-initialised by every replay;
  * gradient addresses must not change between the eager warm-up and the capture: `FusedAdamW` reads gradients through a
    device-side pointer table built by the warm-up step, so zero gradients in place (`zero_grad(set_to_none=False)`, or the
    fixed grad views of `parallel.GradientBucketer`); `FusedAdamW.step()` raises under capture if a gradient moved;
  * inputs are static buffers: copy each batch into one of the buffers that were captured, then `replay(buffer)`;
  * decoder dropout: the kernels take their seeds as host-side arguments (drawn from torch's CPU generator per call), so a
    captured step would replay the SAME keep masks every iteration — `ops.new_dropout_seeds` refuses to run under capture:
    capture only with `decoder_dropout = 0` (as the parity and benchmark configurations do) or run eagerly;
  * collectives stay outside the graph: with `parallel.GradientBucketer` set `defer_comm = True`, capture forward+backward,
    call `allreduce_now()` eagerly after each replay and run (or separately capture) the optimizer step after it.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch

from . import ops


def use_side_stream(device) -> torch.cuda.Stream:
    """Make a fresh non-default stream current on `device` (call before building the model) and return it."""
    s = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(s)
    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)   # the capture stream differs by design
    return s


class GraphedTrainStep:
    """`step_fn(input_buffer) -> loss` captured per input buffer; `replay(input_buffer)` returns the (static) loss tensor.

    All captures share one memory pool: they never run concurrently and hand nothing to each other but parameters,
    gradients and optimizer state, which live outside the pool.
    """

    def __init__(self, step_fn: Callable[[torch.Tensor], torch.Tensor]):
        self.step_fn = step_fn
        self._graphs: Dict[int, Tuple[torch.cuda.CUDAGraph, torch.Tensor, int]] = {}

    def capture(self, buf: torch.Tensor) -> None:
        if not buf.is_cuda:
            raise RuntimeError("GraphedTrainStep.capture needs a CUDA input buffer")
        if torch.cuda.current_stream(buf.device) == torch.cuda.default_stream(buf.device):
            raise RuntimeError("GraphedTrainStep: run on a non-default stream (graph_step.use_side_stream) from model construction on")
        g = torch.cuda.CUDAGraph()
        pool = next(iter(self._graphs.values()))[0].pool() if self._graphs else None
        before = ops.launch_count()
        with torch.cuda.graph(g, pool=pool):
            out = self.step_fn(buf)
        self._graphs[buf.data_ptr()] = (g, out, ops.launch_count() - before)

    def replay(self, buf: torch.Tensor) -> torch.Tensor:
        try:
            g, out, n = self._graphs[buf.data_ptr()]
        except KeyError:
            raise RuntimeError("GraphedTrainStep.replay: this buffer was not captured (call capture(buffer) first)") from None
        g.replay()
        ops.add_launch_count(n)   # a replay launches the kernels that were counted while it was captured
        return out

    def pool(self):
        """Memory pool of the captures (pass it to further `torch.cuda.graph(..., pool=...)` captures of the same step)."""
        return next(iter(self._graphs.values()))[0].pool() if self._graphs else None

    def launches_per_step(self, buf: torch.Tensor) -> int:
        return self._graphs[buf.data_ptr()][2]
