"""Drop-in for `src/mae/losses.py` (:19-144): latent-space MAE losses on gathered rows [N_mask, D].
Per-row normalisation uses the BIASED variance and eps=1e-6 inside the sqrt (:36-38). CUDA tensors go to the fused row-loss
kernels (`vmae_row_loss_fwd/bwd`: one pass for statistics + loss, one for the gradient; the target is detached, as at the
reference's call site tinyvit_mae.py:112). There is no CPU path: CPU tensors raise. The pixel-space loss of the training step
is `ops.masked_mse`."""
from __future__ import annotations

import torch


def _rows(pred: torch.Tensor, target: torch.Tensor):
    """The losses run on the sm_100a row-loss kernels only: CUDA tensors of equal shape, viewed as [rows, last dim]. The
    target is a constant of the loss (the reference's call site passes `tok[mask].detach()`, tinyvit_mae.py:113-114)."""
    if not (pred.is_cuda and target.is_cuda):
        raise RuntimeError("ssl_vit_video_analytics_b200.mae.losses run on CUDA (sm_100a) tensors only; there is no CPU path")
    if pred.shape != target.shape:
        raise ValueError(f"pred {tuple(pred.shape)} and target {tuple(target.shape)} must have the same shape")
    if target.requires_grad:
        raise RuntimeError("the fused row losses treat the target as a constant: pass target.detach()")
    d = pred.shape[-1]
    return pred.reshape(-1, d), target.reshape(-1, d)


def mae_l2_loss(pred: torch.Tensor, target: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    from .. import ops

    p, t = _rows(pred, target)
    return ops.row_loss(p, t, 0, normalize, 1.0e-6)


def mae_l1_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    from .. import ops

    p, t = _rows(pred, target)
    return ops.row_loss(p, t, 1, False, 0.0)


def mae_cosine_loss(pred: torch.Tensor, target: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    from .. import ops

    p, t = _rows(pred, target)
    return ops.row_loss(p, t, 2, False, eps)


def build_mae_loss(pred: torch.Tensor, target: torch.Tensor, loss_type: str = "l2", normalize: bool = True) -> torch.Tensor:
    if loss_type == "l2":
        return mae_l2_loss(pred, target, normalize=normalize)
    if loss_type == "l1":
        return mae_l1_loss(pred, target)
    if loss_type == "cosine":
        return mae_cosine_loss(pred, target)
    raise ValueError(f"Unknown MAE loss type: {loss_type}")


@torch.no_grad()
def reconstruction_error_stats(pred: torch.Tensor, target: torch.Tensor) -> dict:
    err = pred - target
    return {
        "mae_l1": err.abs().mean().item(),
        "mae_l2": (err ** 2).mean().item(),
        "pred_std": pred.std(dim=0).mean().item(),
        "target_std": target.std(dim=0).mean().item(),
    }


def mae_recon_loss(pred, target, *args, **kwargs):
    """Backward-compatible entry (:115-144): (pred, target[, normalize] | [, loss_type, normalize]) plus the keyword
    spellings loss_type= / normalize_target= / normalize=."""
    loss_type = kwargs.pop("loss_type", "l2")
    if "normalize_target" in kwargs:
        normalize = bool(kwargs.pop("normalize_target"))
    else:
        normalize = bool(kwargs.pop("normalize", True))
    if len(args) == 1:
        normalize = bool(args[0])
    elif len(args) >= 2:
        loss_type, normalize = str(args[0]), bool(args[1])
    return build_mae_loss(pred, target, loss_type=loss_type, normalize=normalize)
