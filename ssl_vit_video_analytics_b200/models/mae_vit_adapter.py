"""Drop-in for `src/models/mae_vit_adapter.py` (TinyVideoMAE): same constructor, forward signature and 289-key
state_dict. The decoder runs as fused sm_100a transformer blocks; the positional-embedding add and mask-token
substitution (:97-104) are one kernel; `enc_to_dec`, `decoder_norm`, `decoder_pred` are the tcgen05 GEMM / LayerNorm.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """timm.layers.trunc_normal_ as used by the reference (:3,59-61); same sampler as torch.nn.init.trunc_normal_."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class _SelfAttentionParams(nn.Module):
    """Parameter container with nn.MultiheadAttention's key layout: in_proj_weight [3D,D] (q|k|v row blocks),
    in_proj_bias, out_proj.{weight,bias}."""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.0)
        nn.init.constant_(self.out_proj.bias, 0.0)


class DecoderLayer(nn.Module):
    """nn.TransformerEncoderLayer(d_model, nhead, dim_feedforward, batch_first=True, norm_first=True, activation="gelu")
    with identical parameter names (:40-47)."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.1):
        super().__init__()
        # construction order matters for RNG-identical initialisation with the reference
        self.self_attn = _SelfAttentionParams(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.nhead = nhead

    def forward(self, x, seeds=None):
        """x: fp32 [B, N, D] -> fp32 [B, N, D]. In training mode with dropout p > 0 the four dropout sites of the reference
        layer are active (attention probabilities in the flash kernel, dropout1, dropout after GELU, dropout2); the masks come
        from a stateless hash seeded from torch's CPU generator, so runs are reproducible under torch.manual_seed but the
        realisations differ from PyTorch's Philox stream (statistical, not bit-wise, equivalence)."""
        B, N, D = x.shape
        a = self.self_attn
        p = float(self.dropout.p) if self.training else 0.0
        y = ops.transformer_block(x.reshape(B * N, D), self.norm1, a.in_proj_weight, a.in_proj_bias, a.out_proj, self.norm2,
                                  self.linear1, self.linear2, self.nhead, B, N, p_drop=p, seeds=seeds)
        return y.view(B, N, D)


class DecoderStack(nn.Module):
    """nn.TransformerEncoder(layer, num_layers): children live under `.layers.{i}` (:48)."""

    def __init__(self, make_layer, num_layers):
        super().__init__()
        first = make_layer()
        # nn.TransformerEncoder deep-copies one initialised layer: every layer starts from the same values
        import copy

        self.layers = nn.ModuleList([first] + [copy.deepcopy(first) for _ in range(num_layers - 1)])
        self.num_layers = num_layers

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class TinyVideoMAE(nn.Module):
    def __init__(self, encoder, config):
        super().__init__()
        self.encoder = encoder
        self.clip_len = config["dataset"]["clip_len"]
        self.img_size = config["dataset"]["image_size"]
        self.patch_size = 8
        self.num_patches = (self.img_size // self.patch_size) ** 2
        self.encoder_dim = 384  # TinyViT stage-3 width (:17-18)
        self.decoder_embed_dim = config["model"]["decoder_embed_dim"]
        self.decoder_num_heads = config["model"]["decoder_num_heads"]
        self.decoder_depth = config["model"]["decoder_depth"]
        # not a reference key: the reference hard-wires nn.TransformerEncoderLayer's default 0.1 (:40-47)
        dec_dropout = float(config["model"].get("decoder_dropout", 0.1))

        D = self.decoder_embed_dim
        self.enc_to_dec = nn.Linear(self.encoder_dim, D, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, D))
        self.temporal_pos_embed = nn.Parameter(torch.zeros(1, self.clip_len, 1, D))
        self.spatial_pos_embed = nn.Parameter(torch.zeros(1, 1, self.num_patches, D))
        self.decoder_blocks = DecoderStack(lambda: DecoderLayer(D, self.decoder_num_heads, int(D * 4), dec_dropout),
                                           self.decoder_depth)
        self.decoder_norm = nn.LayerNorm(D)
        self.decoder_pred = nn.Linear(D, self.patch_size ** 2 * 3, bias=True)
        self.initialize_weights()

    def initialize_weights(self):
        trunc_normal_(self.temporal_pos_embed, std=0.02)
        trunc_normal_(self.spatial_pos_embed, std=0.02)
        trunc_normal_(self.mask_token, std=0.02)
        self.apply(self._init_weights)  # note: also re-initialises every nn.Linear of the encoder (:64-70)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x, mask):
        """x: [B, 3, T, H, W] float; mask: [B, T, L] bool (True = masked) -> pred [B, T*L, 192]."""
        if not x.is_cuda:
            raise RuntimeError("TinyVideoMAE runs on CUDA (sm_100a) tensors only; there is no CPU path")
        with ops.weight_cache_scope(self):   # one launch: bf16 operand copies of every weight matrix, valid for this forward
            return self._forward(x, mask)

    def _forward(self, x, mask):
        B, Cc, T, H, W = x.shape
        Lp = self.num_patches
        frames = x.permute(0, 2, 1, 3, 4).flatten(0, 1)
        latent = self.encoder.forward_stage3(frames)                       # [B*T, 384, S/8, S/8]
        tok = latent.permute(0, 2, 3, 1).reshape(B * T * Lp, self.encoder_dim)
        z = ops.LinearFn.apply(tok, self.enc_to_dec.weight, self.enc_to_dec.bias, L.ACT_NONE)   # bf16 [B*T*L, D]
        m8 = mask.reshape(B, T * Lp).to(torch.uint8).contiguous()
        x_dec = ops.DecoderInputFn.apply(z, None, m8, self.temporal_pos_embed, self.spatial_pos_embed, self.mask_token,
                                         B, T, Lp)                          # fp32 [B, T*L, D]
        x_dec = self.decoder_blocks(x_dec)
        y = ops.LayerNormFn.apply(x_dec, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_norm.eps, ops.BF16)
        return ops.linear(y, self.decoder_pred.weight, self.decoder_pred.bias)
