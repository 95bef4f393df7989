"""Supervised finetune path on the MAE encoder. The reference's finetune driver follows one pattern
(train_finetune.py:19-40, heads.py:6-21): encode every frame, global-average-pool the feature map, average over time,
`nn.Linear(embed_dim, num_classes)`, cross-entropy; the privacy harness perturbs the pooled video embedding with
`add_gaussian_noise` (feature_noise.py:4-7 at run_privacy.py:299). Here that pattern sits on the TinyViT encoder that the MAE
stage pre-trains:

  * all B*T frames go through `TinyViT.forward` (four stages, tiny_vit.py:178-186) in ONE batched pass on the sm_100a kernels
    (the reference loops over frames in Python);
  * GAP + temporal mean = one mean over the T*h*w rows of a clip (the squeeze-excite pooling kernel, fp32 accumulation);
  * optional Gaussian feature noise (Philox kernel) on the pooled embedding; the 576 -> num_classes head runs on the tcgen05 GEMM
    and the loss is the fused softmax cross-entropy (`ops.softmax_cross_entropy`).

`resolve_mode` / `apply_finetune_mode` reproduce the four experiment modes (train_finetune.py:198-210,289-313) and
`checkpoint.load_pretrained_ssl` loads the SSL encoder into `.backbone`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..privacy.feature_noise import add_gaussian_noise
from .tiny_vit import tiny_vit_21m_variant

FINETUNE_MODES = ("ft_random", "linear_probe", "ft_ssl", "two_stage")


def linear_head(z: torch.Tensor, head: nn.Linear) -> torch.Tensor:
    """`head(z)` on the tcgen05 GEMM (bf16 operands, fp32 accumulation). The kernel tiles N in multiples of 8, so a 101-class head
    is evaluated with its weight zero-padded to 104 rows and the padding columns are dropped again (views, no copy of z)."""
    n = head.out_features
    pad = (-n) % 8
    w, b = head.weight, head.bias
    if pad:
        w = torch.nn.functional.pad(w, (0, 0, 0, pad))
        b = torch.nn.functional.pad(b, (0, pad)) if b is not None else None
    y = ops.linear(z, w, b)
    return y[:, :n] if pad else y


class TinyViTVideoClassifier(nn.Module):
    def __init__(self, num_classes: int, encoder: nn.Module | None = None, img_size: int = 112, embed_dim: int = 576,
                 noise_sigma: float = 0.0):
        super().__init__()
        self.backbone = encoder if encoder is not None else tiny_vit_21m_variant(img_size=img_size, use_checkpoint=False)
        self.classifier = nn.Linear(embed_dim, num_classes)
        self.noise_sigma = float(noise_sigma)

    def video_embedding(self, clip: torch.Tensor) -> torch.Tensor:
        """clip [B, C, T, H, W] -> fp32 [B, embed_dim]: mean over frames of the spatially pooled stage-4 features."""
        if not clip.is_cuda:
            raise RuntimeError("TinyViTVideoClassifier runs on CUDA (sm_100a) tensors only; there is no CPU path")
        B, C, T, H, W = clip.shape
        frames = clip.permute(0, 2, 1, 3, 4).flatten(0, 1)
        feat = self.backbone(frames)                                       # [B*T, D, h, w]
        D, h, w = feat.shape[1], feat.shape[2], feat.shape[3]
        rows = ops.as_nhwc_bf16(feat).view(B, T * h * w, D)                # channels-last rows of one clip are contiguous
        return ops.SEPoolFn.apply(rows)                                    # mean over (t, y, x), fp32 accumulation

    def forward(self, clip: torch.Tensor) -> torch.Tensor:
        z = self.video_embedding(clip)
        if self.noise_sigma > 0.0:
            z = add_gaussian_noise(z, self.noise_sigma)
        return linear_head(z, self.classifier)

    @staticmethod
    def loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """nn.CrossEntropyLoss() of train_finetune.py:86-107 as one fused softmax-CE kernel pair."""
        return ops.softmax_cross_entropy(logits, labels)


def resolve_mode(ft_cfg: dict, cli_mode: str | None = None) -> str:
    """train_finetune.py:198-210"""
    mode = cli_mode or ft_cfg.get("experiment", {}).get("mode", "ft_ssl")
    if mode not in FINETUNE_MODES:
        raise ValueError(f"[ERROR] Unknown mode={mode}, must be one of {sorted(FINETUNE_MODES)}")
    return mode


def set_requires_grad(module: nn.Module, flag: bool) -> None:
    for p in module.parameters():
        p.requires_grad = flag


def apply_finetune_mode(model: nn.Module, mode: str, epoch: int = 0, stage1_epochs: int = 0) -> bool:
    """Freeze / unfreeze the backbone as the four modes prescribe (train_finetune.py:289-313 and the epoch loop's unfreeze at the
    stage boundary). Returns True when the backbone is trainable in this epoch."""
    if mode not in FINETUNE_MODES:
        raise ValueError(f"unknown mode {mode}")
    if mode == "two_stage" and stage1_epochs <= 0:
        raise ValueError("[ERROR] two_stage requires training.stage1_epochs > 0")
    frozen = mode == "linear_probe" or (mode == "two_stage" and epoch < stage1_epochs)
    set_requires_grad(model.backbone, not frozen)
    return not frozen
