"""Drop-in for `src/models/tinyvit_mae.py` (:12-115): `MAEDecoder` and the latent-space `TinyViTMAE`.

`MAEDecoder` (proj -> pre-LN transformer blocks -> pred) runs on the fused sm_100a blocks. `TinyViTMAE`'s mask-token
substitution (`torch.where`, :107-108) and row-major boolean gathers (:113-114) are index kernels with a static row count
(`vmae_mask_token_blend_*`, `vmae_mask_row_indices`, `vmae_gather_rows`): bit-exact gather order, no dynamic-shape sync. Its encoder in
the reference is timm's `tiny_vit_21m_224` (un-vendored third-party code, `tinyvit_backbone.py:29`); timm is not part
of this image, so the encoder here is the in-repo TinyViT followed by the same adaptive average pooling to s x s
tokens (`tinyvit_backbone.py:74-81`) — same architecture family, NOT key-compatible with timm checkpoints
(parity unpinned for that boundary).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from .mae_vit_adapter import DecoderLayer, DecoderStack
from .tiny_vit import tiny_vit_21m_variant


class MAEDecoder(nn.Module):
    def __init__(self, embed_dim: int, decoder_dim: int, depth: int, num_heads: int, mlp_ratio: float, dropout: float = 0.1):
        super().__init__()
        self.proj = nn.Linear(embed_dim, decoder_dim)
        self.blocks = DecoderStack(lambda: DecoderLayer(decoder_dim, num_heads, int(decoder_dim * mlp_ratio), dropout), depth)
        self.pred = nn.Linear(decoder_dim, embed_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = ops.LinearFn.apply(x, self.proj.weight, self.proj.bias, L.ACT_NONE)
        h = self.blocks(ops.cast_f32_autograd(h))
        return ops.linear(ops.cast_bf16_autograd(h), self.pred.weight, self.pred.bias)


class TinyViTBackbone(nn.Module):
    """[B,3,H,W] -> [B, pool*pool, C] tokens: full 4-stage TinyViT, adaptive average pool (on the channels-last map, our kernel),
    flatten. fp32 tokens, as `F.adaptive_avg_pool2d` of an fp32 map would give."""

    embed_dim = 576   # stage-4 width of the 21M variant (the reference infers it with a dummy forward, tinyvit_mae.py:64-68)

    def __init__(self, backbone: str = "tiny_vit_21m_224", pretrained: str = ""):
        super().__init__()
        self.net = tiny_vit_21m_variant(use_checkpoint=False)
        self._load_info = {"missing": [], "unexpected": []}
        if pretrained:
            sd = torch.load(pretrained, map_location="cpu")
            if isinstance(sd, dict) and "model" in sd:
                sd = sd["model"]
            missing, unexpected = self.net.load_state_dict(sd, strict=False)
            self._load_info = {"missing": list(missing), "unexpected": list(unexpected)}

    @property
    def load_info(self):
        return self._load_info

    def forward(self, x: torch.Tensor, pool_size: int = 3) -> torch.Tensor:
        with torch.autocast("cuda", dtype=torch.bfloat16):            # keep the map bf16 channels-last up to the pooling kernel
            f = self.net(x)
        return ops.AdaptiveAvgPoolFn.apply(ops.as_nhwc_bf16(f), int(pool_size))


class TinyViTMAE(nn.Module):
    def __init__(self, backbone_name: str, pretrained: str, stage4_pool: int, decoder_dim: int, decoder_depth: int,
                 decoder_num_heads: int, mlp_ratio: float, drop_rate: float = 0.0, attn_drop_rate: float = 0.0):
        super().__init__()
        self.stage4_pool = int(stage4_pool)
        self.tokens_per_frame = self.stage4_pool * self.stage4_pool
        self.encoder = TinyViTBackbone(backbone=backbone_name, pretrained=pretrained)
        self.embed_dim = int(self.encoder.embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        nn.init.trunc_normal_(self.mask_token, std=0.02)
        self.decoder = MAEDecoder(self.embed_dim, int(decoder_dim), int(decoder_depth), int(decoder_num_heads), float(mlp_ratio),
                                  dropout=float(drop_rate))
        self.count_flag = None   # device int32: set by a sync-free forward (n_mask given) when a clip's mask count differed

    def forward(self, clip: torch.Tensor, token_mask: torch.Tensor, n_mask: int | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """clip [B,C,T,H,W], token_mask bool [B, N] -> (pred [N_mask, D], target [N_mask, D]) in the row-major order of the
        reference's boolean gathers `rec[token_mask]`, `tok[token_mask].detach()` (tinyvit_mae.py:113-114).
        The gathers run as index kernels with a static row count: both modes of `make_token_mask` mask the same number of tokens
        in every clip. `n_mask` (masked tokens per clip) makes the call sync-free (CUDA-graph capturable); without it the count is
        read back once — the reference's boolean indexing synchronises there too — and checked to be uniform."""
        B, Cc, T, H, W = clip.shape
        P = self.tokens_per_frame
        N = T * P
        assert token_mask.shape == (B, N), f"token_mask must be [B,{N}] got {tuple(token_mask.shape)}"
        if not clip.is_cuda:
            raise RuntimeError("TinyViTMAE runs on CUDA (sm_100a) tensors only; there is no CPU path")
        m8 = token_mask.to(torch.uint8).contiguous()
        flag = None
        if n_mask is None:
            counts = m8.sum(dim=1, dtype=torch.int32)
            n0, uniform = torch.stack([counts[0], (counts == counts[0]).all().to(torch.int32)]).tolist()
            if not uniform:
                raise ValueError("TinyViTMAE: every clip must mask the same number of tokens (as mae.masking.make_token_mask does)")
            n_mask = int(n0)
        else:
            flag = torch.zeros((), device=clip.device, dtype=torch.int32)
            self.count_flag = flag
        if n_mask == 0:
            empty = torch.empty((0, self.embed_dim), device=clip.device)
            return empty, empty.clone()
        frames = clip.permute(0, 2, 1, 3, 4).reshape(B * T, Cc, H, W).contiguous()
        with ops.weight_cache_scope(self):
            tok = self.encoder(frames, pool_size=self.stage4_pool).reshape(B, N, self.embed_dim)
            tok_in = ops.MaskTokenBlendFn.apply(tok, m8, self.mask_token)
            rec = self.decoder(tok_in)
        idx = ops.mask_row_indices(m8, n_mask, 1, flag)
        return ops.GatherRowsFn.apply(rec, idx), ops.GatherRowsFn.apply(tok.detach(), idx)
