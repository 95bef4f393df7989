"""AdamW on the GPU in one launch (`torch.optim.AdamW(model.parameters(), lr, weight_decay)` at
`src/train_ssl_mae.py:163`, `src/mae/engine.py` via `build_optimizer`).

Same constructor arguments, `param_groups`, `state` and update rule as `torch.optim.AdamW` (decoupled weight decay, bias
correction — with ONE step count per parameter group where torch keeps one per tensor; the two differ only for a parameter
that starts receiving gradients later than its group, e.g. a backbone unfrozen by `two_stage` finetune: give such parameters
their own group); the difference is where things live: every hyper-parameter the update reads
(lr, betas, eps, weight decay, step) sits in a small device array per parameter group, and all tensors of a group are updated by
ONE kernel launch through a device-side chunk table (`vmae_adamw_multi`). Nothing is read from the host at run time, so a step
captured in a CUDA graph (graph_step.py) stays correct while a scheduler changes `param_groups[i]["lr"]` between replays — call
`sync_hyperparameters()` (done automatically by an eager `step()`) after changing them.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError(f"FusedAdamW: invalid hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans = {}   # group index -> plan (chunk table on the device, its host/pinned images, hp device array, ...)

    # ------------------------------------------------------------------------------------------------------------
    def _plan(self, gi, group):
        """Chunk table of a group. The parameter / state columns are built once; the gradient column is refreshed (one small
        async copy from pinned memory, no synchronisation) whenever the gradients moved, e.g. after `zero_grad(set_to_none=True)`."""
        ps = [p for p in group["params"] if p.grad is not None]
        if not ps:
            return None
        pkey = tuple((p.data_ptr(), p.numel()) for p in ps)
        gkey = tuple(p.grad.data_ptr() for p in ps)
        plan = self._plans.get(gi)
        if plan is not None and plan["pkey"] == pkey and plan["gkey"] == gkey:
            return plan
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.is_contiguous()):
                raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients (there is no fallback path)")
        dev = ps[0].device
        if plan is None or plan["pkey"] != pkey:
            chunk = L.load().vmae_adamw_chunk()
            idx, off, cnt, cols = [], [], [], [[], [], []]
            for k, p in enumerate(ps):
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=dev)   # kept for state_dict compatibility
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                n = p.numel()
                for o in range(0, n, chunk):
                    idx.append(k)
                    off.append(4 * o)
                    cnt.append(min(chunk, n - o))
                    cols[0].append(p.data_ptr() + 4 * o)
                    cols[1].append(st["exp_avg"].data_ptr() + 4 * o)
                    cols[2].append(st["exp_avg_sq"].data_ptr() + 4 * o)
            # the step count survives a rebuild: from the previous plan (device-to-device), else from a loaded state_dict
            hp = torch.zeros(6, dtype=torch.float32, device=dev)
            if plan is not None:
                hp[5:6].copy_(plan["hp"][5:6])
            else:
                hp[5:6].copy_(self.state[ps[0]]["step"].reshape(1).to(torch.float32))
            host = torch.zeros((len(idx), 5), dtype=torch.int64)
            host[:, 0] = torch.tensor(cols[0], dtype=torch.int64)
            host[:, 2] = torch.tensor(cols[1], dtype=torch.int64)
            host[:, 3] = torch.tensor(cols[2], dtype=torch.int64)
            host[:, 4] = torch.tensor(cnt, dtype=torch.int64)
            plan = {"pkey": pkey, "gkey": None, "idx": torch.tensor(idx, dtype=torch.int64), "off": torch.tensor(off, dtype=torch.int64),
                    "host": host, "staging": [host.clone().pin_memory() for _ in range(2)], "staged": [None, None], "turn": 0,
                    "table": torch.empty((len(idx), 5), dtype=torch.int64, device=dev), "chunks": len(idx), "hp": hp, "hyper": None, "params": ps}
            self._plans[gi] = plan
        # gradient column -> pinned staging buffer (two of them, each guarded by the event of its last copy) -> device table
        t = plan["turn"]
        plan["turn"] = 1 - t
        if plan["staged"][t] is not None:
            plan["staged"][t].synchronize()
        stage = plan["staging"][t]
        stage.copy_(plan["host"])
        stage[:, 1] = torch.tensor(gkey, dtype=torch.int64)[plan["idx"]] + plan["off"]
        plan["table"].copy_(stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        plan["staged"][t] = ev
        plan["gkey"] = gkey
        return plan

    def sync_hyperparameters(self):
        """Push lr / betas / eps / weight decay of every group to its device array (one tiny async copy, only when they changed)."""
        for gi, group in enumerate(self.param_groups):
            plan = self._plans.get(gi)
            if plan is None:
                continue
            host = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]))
            if plan["hyper"] != host:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("FusedAdamW: hyper-parameters changed inside a CUDA-graph capture; call sync_hyperparameters() before it")
                plan["hp"][:5].copy_(torch.tensor(host, dtype=torch.float32), non_blocking=False)
                plan["hyper"] = host

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        for gi, group in enumerate(self.param_groups):
            plan = self._plans.get(gi) if capturing else self._plan(gi, group)
            if plan is None:
                if capturing:
                    raise RuntimeError("FusedAdamW: run one eager step() before capturing (the chunk table and the state are built there)")
                continue
            if capturing:
                # The chunk table was filled by the eager warm-up step and cannot be rebuilt inside a capture (it is an H2D
                # copy). A captured step must therefore keep every gradient at the address the warm-up saw: use
                # `zero_grad(set_to_none=False)` or `parallel.GradientBucketer` (fixed grad views). With the torch default
                # `set_to_none=True`, backward would allocate fresh .grad tensors in the graph pool and every replay would
                # apply AdamW to the stale warm-up gradients.
                ps = plan["params"]
                if any(p.grad is None for p in ps) or tuple(p.grad.data_ptr() for p in ps) != plan["gkey"]:
                    raise RuntimeError(
                        "FusedAdamW.step() inside a CUDA-graph capture: a gradient is missing or has moved since the eager warm-up "
                        "step. Keep gradient addresses fixed in a captured step (zero_grad(set_to_none=False) or GradientBucketer).")
            self.sync_hyperparameters()
            L.check(lib.vmae_adamw_multi(ops._p(plan["table"]), plan["chunks"], ops._p(plan["hp"]), ops._stream()), "vmae_adamw_multi")
            ops._count(2)
        return loss

    def state_dict(self):
        # the per-parameter "step" entries mirror the group counter (torch.optim.AdamW layout)
        for plan in self._plans.values():
            for p in plan["params"]:
                self.state[p]["step"].copy_(plan["hp"][5])
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans = {}   # rebuilt (with the loaded step count) by the next step()
