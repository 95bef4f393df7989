"""CPU: stack-2 training-loop semantics (mae/engine.py:14-170, train_mae.py:45-69) on a stub MAE model —
mask-ratio schedule, gradient accumulation, clipping, non-finite skip, one scheduler step per optimizer step."""
import math

import torch

from oracle import restatement as R
from ssl_vit_video_analytics_b200.mae import engine as E


class Log:
    def __init__(self):
        self.lines = []

    def write(self, msg):
        self.lines.append(msg)


class StubMAE(torch.nn.Module):
    """forward(clip, token_mask) -> (pred, target) on the masked rows, like TinyViTMAE (tinyvit_mae.py:82-115)."""

    def __init__(self, poison_step=-1):
        super().__init__()
        self.lin = torch.nn.Linear(8, 8)
        self.calls = 0
        self.poison_step = poison_step
        self.seen_masks = []

    def forward(self, clip, token_mask):
        B, C, T = clip.shape[:3]
        tok = clip.mean(dim=(3, 4)).permute(0, 2, 1)                       # [B, T, C]
        tok = tok.repeat_interleave(9, dim=1)[..., :1].expand(-1, -1, 8)      # [B, T*9, 8]
        self.seen_masks.append(token_mask.clone())
        pred = self.lin(tok)[token_mask]
        target = tok[token_mask].detach() * 2.0
        if self.calls == self.poison_step:
            pred = pred * float("nan")
        self.calls += 1
        return pred, target


def _loader(n):
    g = torch.Generator().manual_seed(0)
    return [torch.randn(2, 3, 4, 8, 8, generator=g) for _ in range(n)]


def test_scheduler_is_linear_warmup_then_cosine():
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    sch = E.build_scheduler({"training": {"epochs": 20}, "scheduler": {"warmup_epochs": 4, "min_lr_ratio": 0.1}}, opt)
    lrs = []
    for _ in range(20):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    assert lrs[:4] == [0.25, 0.5, 0.75, 1.0]
    assert abs(lrs[4] - 1.0) < 1e-12 and abs(lrs[12] - (0.1 + 0.9 * 0.5 * (1 + math.cos(math.pi * 8 / 16)))) < 1e-12
    assert all(a >= b for a, b in zip(lrs[4:], lrs[5:]))


def test_epoch_semantics_accumulation_clip_skip_and_schedule():
    cfg = {"training": {"grad_accum": 2, "grad_clip": 0.5, "amp": False, "log_interval": 1},
           "mae": {"mask_mode": "tube", "mask_ratio": 0.5, "loss_type": "l2", "normalize_target": True,
                   "mask_ratio_schedule": [{"start": 1, "end": 3, "value": 0.25}, {"start": 3, "end": 100, "value": 0.75}]},
           "model": {"stage4_pool": 3}, "optimizer": {"lr": 1e-2}}
    torch.manual_seed(0)
    model = StubMAE(poison_step=2)
    opt = E.build_optimizer(cfg, model)
    steps = {"opt": 0, "sch": 0}
    orig_step = opt.step
    opt.step = lambda *a, **k: (steps.__setitem__("opt", steps["opt"] + 1), orig_step(*a, **k))[1]

    class Sch:
        def step(self):
            steps["sch"] += 1

    log = Log()
    w0 = model.lin.weight.detach().clone()
    out = E.train_one_epoch(1, model, _loader(6), opt, None, Sch(), torch.device("cpu"), cfg, log, loss_fn=R.mae_recon_loss)
    # 6 batches, accumulation 2 -> optimizer steps after batches 1, 3, 5; batch 2 is non-finite: skipped (gradients cleared), so the step
    # after batch 3 carries only batch 3's gradient; the scheduler steps with the optimizer
    assert steps == {"opt": 3, "sch": 3}
    assert any("not finite" in ln for ln in log.lines) and out["mask_ratio"] == 0.25
    assert math.isfinite(out["loss"]) and set(out) == {"loss", "l1", "l2", "pred_std", "target_std", "mask_ratio", "time_sec"}
    assert not torch.equal(w0, model.lin.weight)
    m = model.seen_masks[0]
    assert m.shape == (2, 4 * 9) and int(m[0].sum()) == 4 * max(1, round(9 * 0.25))     # tube mask: same patches in every frame
    out3 = E.train_one_epoch(3, StubMAE(), _loader(2), opt, None, None, torch.device("cpu"), cfg, Log(), loss_fn=R.mae_recon_loss)
    assert out3["mask_ratio"] == 0.75


def test_gradient_clipping_bounds_the_update():
    cfg = {"training": {"grad_accum": 1, "grad_clip": 1e-3, "amp": False}, "mae": {"mask_ratio": 0.5}, "model": {"stage4_pool": 3}}
    model = StubMAE()
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    w0 = torch.cat([p.detach().flatten().clone() for p in model.parameters()])
    E.train_one_epoch(1, model, _loader(1), opt, None, None, torch.device("cpu"), cfg, Log(), loss_fn=R.mae_recon_loss)
    w1 = torch.cat([p.detach().flatten() for p in model.parameters()])
    assert float((w1 - w0).norm()) <= 1e-3 * 1.0001           # SGD with lr 1: the step is the clipped gradient


def test_fused_adamw_host_side_contract():
    """optim.FusedAdamW: torch.optim.AdamW's constructor contract (defaults, param groups, validation); no CPU update path."""
    import pytest
    import torch

    from ssl_vit_video_analytics_b200.optim import FusedAdamW

    p = torch.nn.Parameter(torch.zeros(4))
    opt = FusedAdamW([{"params": [p], "weight_decay": 0.0}], lr=5e-4, weight_decay=0.05)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(4))], lr=5e-4, weight_decay=0.05)
    assert opt.defaults["betas"] == ref.defaults["betas"] and opt.defaults["eps"] == ref.defaults["eps"]
    assert opt.param_groups[0]["weight_decay"] == 0.0 and opt.param_groups[0]["lr"] == 5e-4
    for bad in (dict(lr=-1.0), dict(betas=(1.0, 0.9)), dict(eps=-1e-8), dict(weight_decay=-0.1)):
        with pytest.raises(ValueError):
            FusedAdamW([p], **bad)
    opt.step()                      # no gradient yet: nothing to do, nothing to launch
    p.grad = torch.zeros(4)
    with pytest.raises(RuntimeError, match="fp32 CUDA"):
        opt.step()                  # a CPU parameter fails loudly (no fallback)


def test_graphed_train_step_host_side_errors():
    """graph_step.GraphedTrainStep refuses what cannot work, without a GPU: a CPU buffer, a buffer that was never captured."""
    import pytest
    import torch

    from ssl_vit_video_analytics_b200.graph_step import GraphedTrainStep

    gs = GraphedTrainStep(lambda clip: clip.sum())
    with pytest.raises(RuntimeError, match="CUDA input buffer"):
        gs.capture(torch.zeros(2))
    with pytest.raises(RuntimeError, match="not captured"):
        gs.replay(torch.zeros(2))
    assert gs.pool() is None
