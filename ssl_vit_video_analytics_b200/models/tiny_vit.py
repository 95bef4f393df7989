"""Drop-in for the reference encoder `src/models/tiny_vit.py` (constructors, forward signatures and state_dict keys
unchanged; file:line citations refer to that file).

Every layer runs on the hand-written sm_100a kernels through `ops`; feature maps are channels-last bf16 internally.
Transformer parts (Attention :86-106, Mlp :74-84, TinyViTBlock :108-130): tcgen05 GEMMs with fused epilogues, tcgen05
flash attention, warp-shuffle LayerNorm. Convolutional stages (Conv2d_BN :12-18, SELayer :20-34, MBConv :36-56,
PatchEmbed :62-72): 1x1 convolutions and the squeeze-excite Linears on the same GEMM over rows = pixels, the 3x3 stem
convolutions as im2col + GEMM, depthwise 3x3 on the TMA-staged depthwise kernel, BatchNorm (train-mode batch statistics
with running-stat update; eval-mode running statistics) with the following GELU and the MBConv residual fused into its
apply pass. CPU tensors and layer configurations without a kernel raise — there is no fallback path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.utils.checkpoint as checkpoint

from .. import _lib as L
from .. import ops


class DropPath(nn.Module):
    """Per-sample stochastic depth (the reference imports timm.layers.DropPath, tiny_vit.py:5). `keep_override` (a list of
    0/1 tensors [B], used cyclically call by call) replaces the Bernoulli draw: parity tests feed both sides the same masks."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep
        self.keep_override = None
        self._calls = 0

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        if self.keep_override is not None:
            mask = self.keep_override[self._calls % len(self.keep_override)].to(device=x.device, dtype=x.dtype).reshape(shape).clone()
            self._calls += 1
        else:
            mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask

    def extra_repr(self):
        return f"drop_prob={self.drop_prob:0.3f}"


def _drop_path_active(m: nn.Module) -> bool:
    return isinstance(m, DropPath) and m.drop_prob > 0.0 and m.training


def _nchw_view(rows_nhwc):
    """[N, H, W, C] contiguous -> NCHW view of the same (channels-last) memory."""
    return rows_nhwc.permute(0, 3, 1, 2)


class Conv2d_BN(nn.Sequential):
    """conv (no bias) followed by BatchNorm2d; children are named 'c' and 'bn' (tiny_vit.py:12-18).

    On CUDA the feature map lives channels-last in bf16 and every piece runs on the sm_100a kernels:
      * 1x1 convolution            -> tcgen05 GEMM over rows = pixels,
      * dense 3x3 (pad 1, stem)    -> im2col kernel + the same GEMM (the 3-channel image is gathered from fp32 NCHW),
      * depthwise 3x3 (pad 1)      -> depthwise kernel,
      * train-mode BatchNorm       -> statistics + apply kernels, with the following GELU and the MBConv residual fused.
    `forward(x, act=False, residual=None)`: act/residual are the fusion hooks used by MBConv / PatchEmbed."""

    def __init__(self, a, b, ks=1, stride=1, pad=0, dilation=1, groups=1, bn_weight_init=1):
        super().__init__()
        self.add_module("c", nn.Conv2d(a, b, ks, stride, pad, dilation, groups, bias=False))
        self.add_module("bn", nn.BatchNorm2d(b))
        nn.init.constant_(self.bn.weight, bn_weight_init)
        nn.init.constant_(self.bn.bias, 0)

    def _kind(self):
        c = self.c
        if c.dilation != (1, 1) or c.out_channels % 8 != 0:
            return None
        if c.kernel_size == (1, 1) and c.stride == (1, 1) and c.padding == (0, 0) and c.groups == 1 and c.in_channels % 8 == 0:
            return "pointwise"
        if c.kernel_size == (3, 3) and c.padding == (1, 1) and c.stride in ((1, 1), (2, 2)):
            if c.groups == c.in_channels == c.out_channels:
                return "depthwise"
            if c.groups == 1:
                return "dense3x3" if c.in_channels % 8 == 0 else "image3x3"
        return None

    def _conv_rows(self, x, kind):
        """-> (rows bf16 [M, Cout], (N, Ho, Wo))"""
        c = self.c
        st = c.stride[0]
        if kind == "image3x3":
            kp = (c.in_channels * 9 + 7) // 8 * 8
            col, ho, wo = ops.im2col3x3_image(x, st, kp)
            w2d = torch.nn.functional.pad(c.weight.reshape(c.out_channels, -1), (0, kp - c.in_channels * 9))
            return ops.LinearFn.apply(col, w2d, None, L.ACT_NONE), (x.shape[0], ho, wo)
        xr = ops.as_nhwc_bf16(x)
        n, h, w, cin = xr.shape
        if kind == "pointwise":
            return ops.LinearFn.apply(xr.view(n * h * w, cin), c.weight, None, L.ACT_NONE), (n, h, w)
        if kind == "depthwise":
            y = ops.DepthwiseConv3x3Fn.apply(xr, c.weight, st)
            return y.view(-1, cin), (n, y.shape[1], y.shape[2])
        col = ops.Im2col3x3Fn.apply(xr, st)                                   # dense3x3, (kh, kw, c) column order
        w2d = c.weight.permute(0, 2, 3, 1).reshape(c.out_channels, -1)
        ho, wo = (h - 1) // st + 1, (w - 1) // st + 1
        return ops.LinearFn.apply(col, w2d, None, L.ACT_NONE), (n, ho, wo)

    def forward(self, x, act=False, residual=None):
        if not x.is_cuda:
            raise RuntimeError("Conv2d_BN runs on CUDA (sm_100a) tensors only; there is no CPU path")
        kind = self._kind()
        bn = self.bn
        if kind is None:
            raise RuntimeError(f"Conv2d_BN: no sm_100a kernel for {self.c} (supported: 1x1 s1, 3x3 pad 1 stride 1|2 dense or "
                               f"depthwise, channels a multiple of 8, dilation 1)")
        if bn.momentum is None or not bn.track_running_stats:
            raise RuntimeError("Conv2d_BN: BatchNorm2d without running statistics / with cumulative momentum is not supported")
        rows, (n, ho, wo) = self._conv_rows(x, kind)
        res_rows = ops.as_nhwc_bf16(residual).view(rows.shape) if residual is not None else None
        a = L.ACT_GELU if act else L.ACT_NONE
        if bn.training:
            y = ops.BatchNormActFn.apply(rows, bn.weight, bn.bias, bn.running_mean, bn.running_var, res_rows, bn.eps, bn.momentum, a)
            with torch.no_grad():
                bn.num_batches_tracked += 1
        else:   # eval(): running statistics, same apply kernel (vmae_bn_apply)
            y = ops.BatchNormEvalFn.apply(rows, bn.weight, bn.bias, bn.running_mean, bn.running_var, res_rows, bn.eps, a)
        return _nchw_view(y.view(n, ho, wo, -1))


class SELayer(nn.Module):
    """Squeeze-excite: GAP -> Linear(c, c/r, no bias) -> ReLU -> Linear(c/r, c, no bias) -> sigmoid gate (tiny_vit.py:20-34)."""

    def __init__(self, channel, reduction=4):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False),
            nn.Sigmoid(),
        )

    def forward(self, x):
        n, c = x.shape[0], x.shape[1]
        if not x.is_cuda or c % 32 != 0:
            raise RuntimeError("SELayer runs on CUDA (sm_100a) feature maps with a multiple of 32 channels; there is no fallback path")
        # pool / gating kernels on the channels-last rows; the two bias-free Linears on the tcgen05 GEMM; ReLU and
        # sigmoid are elementwise glue on the tiny [n, c/4] and [n, c] tensors
        xr = ops.as_nhwc_bf16(x)
        h_, w_ = xr.shape[1], xr.shape[2]
        rows = xr.view(n, h_ * w_, c)
        pooled = ops.SEPoolFn.apply(rows)
        hid = torch.relu(ops.cast_f32_autograd(ops.LinearFn.apply(pooled, self.fc[0].weight, None, L.ACT_NONE)))
        gate = torch.sigmoid(ops.cast_f32_autograd(ops.LinearFn.apply(hid, self.fc[2].weight, None, L.ACT_NONE)))
        return _nchw_view(ops.SEScaleFn.apply(rows, gate).view(n, h_, w_, c))


class MBConv(nn.Module):
    """expand 1x1 -> depthwise 3x3 -> SE -> project 1x1 (BN gamma initialised to 0); residual only when the shape is kept
    (tiny_vit.py:36-56). GELUs are fused into the preceding BatchNorm kernels, the residual into the last one."""

    def __init__(self, in_chans, out_chans, expand_ratio=4, stride=1, drop_path=0.0):
        super().__init__()
        mid = int(in_chans * expand_ratio)
        self.use_res_connect = stride == 1 and in_chans == out_chans
        seq = []
        if expand_ratio != 1:
            seq += [Conv2d_BN(in_chans, mid, ks=1), nn.GELU()]
        seq += [
            Conv2d_BN(mid, mid, ks=3, stride=stride, pad=1, groups=mid),
            nn.GELU(),
            SELayer(mid),
            Conv2d_BN(mid, out_chans, ks=1, bn_weight_init=0),
        ]
        self.conv = nn.Sequential(*seq)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        mods = list(self.conv)
        y = x
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, Conv2d_BN):
                fuse_act = i + 1 < len(mods) and isinstance(mods[i + 1], nn.GELU)
                last = i == len(mods) - 1
                fuse_res = last and self.use_res_connect and not _drop_path_active(self.drop_path)
                y = m(y, act=fuse_act, residual=x if fuse_res else None)
                if fuse_res:
                    return y
                i += 2 if fuse_act else 1
            else:
                y = m(y)
                i += 1
        return x + self.drop_path(y) if self.use_res_connect else y


class PatchEmbed(nn.Module):
    """Stem: conv3x3 s2 -> BN -> GELU -> conv3x3 s1 -> BN (tiny_vit.py:62-72)."""

    def __init__(self, in_chans, embed_dim):
        super().__init__()
        self.patch_embed = nn.Sequential(
            Conv2d_BN(in_chans, embed_dim // 2, ks=3, stride=2, pad=1),
            nn.GELU(),
            Conv2d_BN(embed_dim // 2, embed_dim, ks=3, stride=1, pad=1),
        )

    def forward(self, x):
        y = self.patch_embed[0](x, act=True)      # GELU (index 1) fused into the BatchNorm kernel
        return self.patch_embed[2](y)


class Mlp(nn.Module):
    """fc2(GELU(fc1(x))) — bias+GELU fused in the fc1 GEMM epilogue (tiny_vit.py:74-84)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        if not isinstance(self.act, nn.GELU) or getattr(self.act, "approximate", "none") != "none":
            raise RuntimeError("Mlp: only the exact-erf nn.GELU activation has an sm_100a epilogue")
        h = ops.LinearFn.apply(x, self.fc1.weight, self.fc1.bias, L.ACT_GELU)
        h = self.drop(h)
        y = ops.linear(h, self.fc2.weight, self.fc2.bias)
        return self.drop(y)


class Attention(nn.Module):
    """qkv Linear -> global softmax attention (default 1/sqrt(key_dim) scale; `window_size` and `self.scale` are unused
    in the reference as well) -> proj (tiny_vit.py:86-106). qkv rows are ordered [q|k|v][head][key_dim]."""

    def __init__(self, dim, key_dim, num_heads=8, window_size=7):
        super().__init__()
        self.num_heads = num_heads
        self.scale = key_dim ** -0.5
        self.key_dim = key_dim
        self.qkv = nn.Linear(dim, key_dim * num_heads * 3)
        self.proj = nn.Linear(key_dim * num_heads, dim)

    def forward(self, x):
        qkv = ops.LinearFn.apply(x, self.qkv.weight, self.qkv.bias, L.ACT_NONE)  # [B, L, 3*H*hd] bf16
        o = ops.attention(qkv, self.num_heads)
        return ops.linear(o, self.proj.weight, self.proj.bias)


class TinyViTBlock(nn.Module):
    """Pre-LN transformer block on an NCHW feature map (tiny_vit.py:108-130). Tokens are taken as the channels-last
    view of the map, so consecutive blocks exchange data without layout copies."""

    def __init__(self, dim, num_heads, window_size=7, mlp_ratio=4.0, drop_path=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, key_dim=dim // num_heads, num_heads=num_heads, window_size=window_size)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))

    def forward_tokens(self, t, B, Lq):
        """t: fp32 [B*Lq, C] residual stream -> fp32 [B*Lq, C]."""
        if _drop_path_active(self.drop_path):
            t3 = t.view(B, Lq, -1)
            a = self.attn(ops.layer_norm(t3, self.norm1.weight, self.norm1.bias, self.norm1.eps, ops.BF16))
            t3 = t3 + self.drop_path(ops.cast_f32_autograd(a))
            m = self.mlp(ops.layer_norm(t3, self.norm2.weight, self.norm2.bias, self.norm2.eps, ops.BF16))
            t3 = t3 + self.drop_path(ops.cast_f32_autograd(m))
            return t3.reshape(B * Lq, -1)
        return ops.transformer_block(t, self.norm1, self.attn.qkv.weight, self.attn.qkv.bias, self.attn.proj, self.norm2,
                                     self.mlp.fc1, self.mlp.fc2, self.attn.num_heads, B, Lq)

    def forward(self, x):
        B, Cc, H, W = x.shape
        tok = x.permute(0, 2, 3, 1).reshape(B * H * W, Cc)  # free when x is channels_last
        t = ops.cast_f32_autograd(tok.contiguous())
        t = self.forward_tokens(t, B, H * W)
        if x.dtype == ops.BF16:
            t = ops.cast_bf16_autograd(t)
        return t.view(B, H, W, Cc).permute(0, 3, 1, 2)


class TinyViT(nn.Module):
    """Four stages: conv stage, then three (down-sample MBConv + transformer blocks) stages (tiny_vit.py:136-186)."""

    def __init__(self, img_size=112, in_chans=3, embed_dims=[96, 192, 384, 576], depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_sizes=[7, 7, 14, 7], drop_path_rate=0.1, use_checkpoint=True):
        super().__init__()
        self.use_checkpoint = use_checkpoint
        self.patch_embed = PatchEmbed(in_chans, embed_dims[0])
        self.stages = nn.ModuleList()
        rates = [r.item() for r in torch.linspace(0, drop_path_rate, sum(depths))]
        k = 0
        for s in range(4):
            blocks = []
            if s > 0:
                blocks.append(MBConv(embed_dims[s - 1], embed_dims[s], stride=2))
            for _ in range(depths[s]):
                if s == 0:
                    blocks.append(MBConv(embed_dims[s], embed_dims[s], drop_path=rates[k]))
                else:
                    blocks.append(TinyViTBlock(dim=embed_dims[s], num_heads=num_heads[s], window_size=window_sizes[s],
                                               drop_path=rates[k]))
                k += 1
            self.stages.append(nn.Sequential(*blocks))

    def _run_stage(self, stage, x):
        if self.use_checkpoint and self.training:
            return checkpoint.checkpoint(stage, x, use_reentrant=False)
        return stage(x)

    @staticmethod
    def _boundary(x):
        """Internally the CUDA path keeps bf16 channels-last maps; outside autocast callers get fp32 like the reference."""
        if x.is_cuda and x.dtype == ops.BF16 and not torch.is_autocast_enabled("cuda"):
            return ops.cast_f32_autograd(x.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
        return x

    def forward_stage3(self, x):
        """MAE pre-training interface: stem + stages 0..2 -> [BT, 384, S/8, S/8] (tiny_vit.py:166-176)."""
        with ops.weight_cache_scope(self):
            x = self.patch_embed(x)
            for s in range(3):
                x = self._run_stage(self.stages[s], x)
            return self._boundary(x)

    def forward(self, x):
        """Classification / finetune interface: all four stages (tiny_vit.py:178-186)."""
        with ops.weight_cache_scope(self):
            x = self.patch_embed(x)
            for stage in self.stages:
                x = self._run_stage(stage, x)
            return self._boundary(x)


def tiny_vit_21m_variant(img_size=112, use_checkpoint=True, **kwargs):
    return TinyViT(img_size=img_size, embed_dims=[96, 192, 384, 576], depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 18],
                   use_checkpoint=use_checkpoint, **kwargs)
