"""Drop-in for `src/federated/fed_loop.py::fedavg_aggregate` (:14-62) and its multi-GPU mapping.

`fedavg_aggregate(global_model, client_states, client_weights)` keeps the reference contract: floating tensors become
sum_i (w_i / sum w) * theta_i accumulated in client order, `num_batches_tracked` takes the max over clients, other
integer buffers are copied from client 0, keys missing from any client keep the global value; the new state (CPU
tensors) is loaded into `global_model` (strict) and returned. When the client tensors live on a CUDA device the
weighted sum runs there through `vmae_scale_f32` (separate IEEE mul and add, so the result is bit-identical to the
reference's CPU loop).

`fedavg_allreduce` is the same aggregation with client k resident on rank k: pre-scale by w_k / sum w, allreduce(sum)
over NCCL; allreduce(max) for `num_batches_tracked`; broadcast from rank 0 for other integer buffers.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .. import ops


def _is_float_tensor(t) -> bool:
    return torch.is_tensor(t) and torch.is_floating_point(t)


def fedavg_aggregate(global_model, client_states: List[Dict[str, torch.Tensor]], client_weights: List[float]):
    if len(client_states) == 0:
        raise RuntimeError("[ERROR] No client states provided for aggregation.")
    if len(client_states) != len(client_weights):
        raise RuntimeError("[ERROR] client_states and client_weights length mismatch.")
    total_w = float(sum(client_weights))
    if total_w <= 0:
        raise RuntimeError("[ERROR] total client weight must be > 0.")

    global_state = global_model.state_dict()
    new_state = {}
    for key, g_t in global_state.items():
        if any(key not in cs for cs in client_states):
            new_state[key] = g_t.detach().cpu()
            continue
        if _is_float_tensor(g_t):
            tensors = [cs[key].detach() for cs in client_states]
            on_gpu = all(t.is_cuda for t in tensors) and g_t.dtype == torch.float32
            if on_gpu:
                acc = torch.zeros_like(tensors[0], dtype=torch.float32)
                for t, w in zip(tensors, client_weights):
                    ops.scale_f32(t.to(torch.float32), float(w) / total_w, out=acc, accumulate=True)
                new_state[key] = acc.cpu()
            else:
                acc = torch.zeros_like(g_t.detach().cpu())
                for t, w in zip(tensors, client_weights):
                    acc += t.cpu().to(acc.dtype) * (float(w) / total_w)
                new_state[key] = acc
        elif "num_batches_tracked" in key:
            vals = torch.stack([cs[key].detach().cpu().to(torch.long) for cs in client_states], dim=0)
            new_state[key] = vals.max(dim=0).values
        else:
            new_state[key] = client_states[0][key].detach().cpu()
    global_model.load_state_dict(new_state, strict=True)
    return new_state


@torch.no_grad()
def fedavg_allreduce(model, client_weight: float, group=None):
    """In-place FedAvg over the ranks of `group` (one client per rank; weight 0 for idle ranks)."""
    import torch.distributed as dist

    state = model.state_dict()
    some = next(iter(state.values()))
    w = torch.tensor([float(client_weight)], dtype=torch.float64, device=some.device)
    dist.all_reduce(w, op=dist.ReduceOp.SUM, group=group)
    total_w = float(w.item())
    if total_w <= 0:
        raise RuntimeError("[ERROR] total client weight must be > 0.")
    alpha = float(client_weight) / total_w
    floats = [t for t in state.values() if _is_float_tensor(t)]
    if floats:
        flat = torch.cat([t.reshape(-1).to(torch.float32) for t in floats])
        if flat.is_cuda:
            ops.scale_f32(flat, alpha, out=flat)
        else:
            flat.mul_(alpha)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        views, off = [], 0
        for t in floats:
            n = t.numel()
            views.append(flat[off:off + n].view_as(t))
            off += n
        if all(t.dtype == torch.float32 for t in floats):
            torch._foreach_copy_(floats, views)        # one multi-tensor launch instead of ~200 small copies
        else:
            for t, v in zip(floats, views):
                t.copy_(v.to(t.dtype))
    for key, t in state.items():
        if _is_float_tensor(t):
            continue
        if "num_batches_tracked" in key:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        else:
            src = dist.get_global_rank(group, 0) if group is not None else 0
            dist.broadcast(t, src=src, group=group)
    return state
