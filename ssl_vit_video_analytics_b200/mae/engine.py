"""Drop-in for `src/mae/engine.py` (:14-170) and the optimizer / scheduler builders of `src/mae/train_mae.py` (:45-69): the
stack-2 training-loop semantics on top of this package's modules —

  * per-epoch mask ratio from `mae.mask_ratio_schedule` (get_mask_ratio), token mask per step (make_token_mask);
  * autocast, GradScaler scale / unscale / step / update, gradient accumulation over `training.grad_accum` steps;
  * gradient clipping at `training.grad_clip` (default 1.0) on the unscaled gradients, optimizer.zero_grad(set_to_none=True);
  * a non-finite loss skips the step (and clears the gradients);
  * the scheduler (cosine with linear warm-up, LambdaLR) is stepped once per optimizer step, as the reference does;
  * the returned meter dictionary and the log line.

The reference's loop has three call-site bugs that make it fail on its own modules; they are fixed here only:
the model is called as `model(clip, token_mask)` (TinyViTMAE.forward takes no `stage4_pool=` keyword, tinyvit_mae.py:82),
and the statistics are read under the keys `reconstruction_error_stats` actually returns (`mae_l1` / `mae_l2`, losses.py:96-108).
"""
from __future__ import annotations

import math
import time
from typing import Any, Dict

import torch
import torch.nn as nn

from .losses import mae_recon_loss, reconstruction_error_stats
from .masking import count_masked, count_visible, get_mask_ratio, make_token_mask


def build_optimizer(cfg: Dict[str, Any], model: nn.Module) -> torch.optim.Optimizer:
    """train_mae.py:45-50 (AdamW, betas default (0.9, 0.95), weight decay on every parameter); `optim.FusedAdamW` on CUDA parameters."""
    opt_cfg = cfg.get("optimizer", {})
    lr = float(opt_cfg.get("lr", opt_cfg.get("base_lr", 3e-4)))
    wd = float(opt_cfg.get("weight_decay", 0.05))
    betas = tuple(opt_cfg.get("betas", [0.9, 0.95]))
    params = list(model.parameters())
    if len(params) > 0 and all(p.is_cuda for p in params):
        from ..optim import FusedAdamW   # one launch per group, device-side hyper-parameters (graph-capturable)

        return FusedAdamW(params, lr=lr, weight_decay=wd, betas=betas)
    return torch.optim.AdamW(params, lr=lr, weight_decay=wd, betas=betas)   # host-side logic tests run on CPU modules


def build_scheduler(cfg: Dict[str, Any], optimizer: torch.optim.Optimizer):
    """train_mae.py:53-69: linear warm-up over `scheduler.warmup_epochs`, cosine to `min_lr_ratio` at `training.epochs`."""
    sch_cfg = cfg.get("scheduler", {})
    epochs = int(cfg.get("training", {}).get("epochs", 200))
    warmup = int(sch_cfg.get("warmup_epochs", 10))
    min_lr_ratio = float(sch_cfg.get("min_lr_ratio", 0.05))

    def lr_lambda(ep: int):
        if ep < warmup:
            return float(ep + 1) / float(max(1, warmup))
        t = (ep - warmup) / float(max(1, epochs - warmup))
        return min_lr_ratio + (1.0 - min_lr_ratio) * 0.5 * (1.0 + math.cos(math.pi * t))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lr_lambda)


def train_one_epoch(epoch: int, model: nn.Module, loader, optimizer: torch.optim.Optimizer, scaler, scheduler, device: torch.device,
                    cfg: Dict[str, Any], logger, loss_fn=None) -> Dict[str, float]:
    """`loss_fn(pred, target, loss_type=..., normalize_target=...)` defaults to `losses.mae_recon_loss` (the sm_100a row-loss
    kernels, CUDA only); the host-logic tests of this loop pass the CPU oracle's loss instead."""
    loss_fn = loss_fn or mae_recon_loss
    model.train()
    training, mae_cfg, model_cfg = cfg.get("training", {}), cfg.get("mae", {}), cfg.get("model", {})
    log_interval = int(training.get("log_interval", 20))
    grad_clip = float(training.get("grad_clip", 1.0))
    grad_accum = int(training.get("grad_accum", 1))
    amp = bool(training.get("amp", True))
    stage4_pool = int(model_cfg.get("stage4_pool", 3))
    tokens_per_frame = stage4_pool * stage4_pool
    mask_mode = str(mae_cfg.get("mask_mode", "tube"))
    mask_ratio = get_mask_ratio(epoch, mae_cfg.get("mask_ratio_schedule", []), float(mae_cfg.get("mask_ratio", 0.8)))
    loss_type = str(mae_cfg.get("loss_type", "l2"))
    normalize_target = bool(mae_cfg.get("normalize_target", True))
    device = torch.device(device)
    use_scaler = scaler is not None and amp

    n_steps, sums = 0, {"loss": 0.0, "l1": 0.0, "l2": 0.0, "pred_std": 0.0, "target_std": 0.0}
    t0 = last = time.time()
    data_t = iter_t = 0.0
    optimizer.zero_grad(set_to_none=True)

    for step, clip in enumerate(loader):
        now = time.time()
        data_t += now - last
        clip = clip.to(device, non_blocking=True)
        if clip.dim() == 4:
            clip = clip.unsqueeze(0)
        if clip.dim() != 5:
            raise RuntimeError(f"Expected clip dim=5, got {clip.shape}")
        B, T = clip.size(0), clip.size(2)
        token_mask = make_token_mask(B=B, T=T, tokens_per_frame=tokens_per_frame, mask_ratio=mask_ratio, mode=mask_mode, device=device)

        with torch.amp.autocast(device_type=device.type, enabled=amp and device.type == "cuda"):
            pred, target = model(clip, token_mask)
            loss = loss_fn(pred, target, loss_type=loss_type, normalize_target=normalize_target)

        if not torch.isfinite(loss):
            logger.write(f"[WARN] epoch={epoch} step={step} loss is not finite: {float(loss)} -> skip step")
            optimizer.zero_grad(set_to_none=True)
            last = time.time()
            continue

        loss_scaled = loss / float(grad_accum)
        if use_scaler:
            scaler.scale(loss_scaled).backward()
        else:
            loss_scaled.backward()

        if (step + 1) % grad_accum == 0:
            if grad_clip > 0:
                if use_scaler:
                    scaler.unscale_(optimizer)
                torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
            if use_scaler:
                scaler.step(optimizer)
                scaler.update()
            else:
                optimizer.step()
            optimizer.zero_grad(set_to_none=True)
            if scheduler is not None:
                scheduler.step()

        stats = reconstruction_error_stats(pred.detach().float(), target.detach().float())
        loss_val = float(loss.detach().item())
        sums["loss"] += loss_val
        sums["l1"] += float(stats["mae_l1"])
        sums["l2"] += float(stats["mae_l2"])
        sums["pred_std"] += float(stats["pred_std"])
        sums["target_std"] += float(stats["target_std"])
        n_steps += 1
        iter_t += time.time() - now
        last = time.time()

        if step % log_interval == 0:
            lr = optimizer.param_groups[0]["lr"]
            total = len(loader) if hasattr(loader, "__len__") else -1
            msg = (f"ep={epoch} step={step}/{total} lr={lr:.6g} loss={loss_val:.4f} l1={stats['mae_l1']:.4f} l2={stats['mae_l2']:.4f} "
                   f"std_pred={stats['pred_std']:.3f} std_tgt={stats['target_std']:.3f} mask={mask_ratio:.2f} "
                   f"tokens(vis={count_visible(token_mask)},mask={count_masked(token_mask)}) t_data={data_t:.2f}s t_iter={iter_t:.2f}s")
            if stats["pred_std"] < 0.05:
                msg += " [WARN:pred_std_low]"
            logger.write(msg)
            data_t = iter_t = 0.0

    t1 = time.time()
    if n_steps == 0:
        return {"loss": 1e9, "l1": 1e9, "l2": 1e9, "pred_std": 0.0, "target_std": 0.0, "mask_ratio": mask_ratio, "time_sec": t1 - t0}
    out = {k: v / n_steps for k, v in sums.items()}
    out.update({"mask_ratio": mask_ratio, "time_sec": t1 - t0})
    return out
