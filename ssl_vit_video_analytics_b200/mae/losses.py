"""Drop-in for `src/mae/losses.py` (:19-144): latent-space MAE losses on gathered rows [N_mask, D].
Per-row normalisation uses the BIASED variance and eps=1e-6 inside the sqrt (:36-38). CUDA tensors go to the fused row-loss
kernels (`vmae_row_loss_fwd/bwd`: one pass for statistics + loss, one for the gradient; the target is detached, as at the
reference's call site tinyvit_mae.py:112). CPU tensors keep the reference's tensor-op formulation (the reference's losses are
device-agnostic and its CPU tests call them); the pixel-space loss of the training step is `ops.masked_mse`."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _normalise_rows(t: torch.Tensor) -> torch.Tensor:
    mu = t.mean(dim=-1, keepdim=True)
    var = t.var(dim=-1, keepdim=True, unbiased=False)
    return (t - mu) / (var + 1.0e-6).sqrt()


def _on_gpu(pred, target) -> bool:
    return pred.is_cuda and target.is_cuda and pred.dim() == 2 and pred.shape == target.shape and not target.requires_grad


def mae_l2_loss(pred: torch.Tensor, target: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    if _on_gpu(pred, target):
        from .. import ops

        return ops.row_loss(pred, target, 0, normalize, 1.0e-6)
    if normalize:
        target = _normalise_rows(target)
    return ((pred - target) ** 2).mean()


def mae_l1_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    if _on_gpu(pred, target):
        from .. import ops

        return ops.row_loss(pred, target, 1, False, 0.0)
    return (pred - target).abs().mean()


def mae_cosine_loss(pred: torch.Tensor, target: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    if _on_gpu(pred, target):
        from .. import ops

        return ops.row_loss(pred, target, 2, False, eps)
    p = F.normalize(pred, dim=-1, eps=eps)
    t = F.normalize(target, dim=-1, eps=eps)
    return 1.0 - (p * t).sum(dim=-1).mean()


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
