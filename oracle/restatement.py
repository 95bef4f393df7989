"""ORACLE — test infrastructure, not product code.

CPU (PyTorch fp32, functional) restatement of the reference's video-MAE hot path, written from the reference's
algorithm; every function cites the reference file:line it follows (paths relative to the reference root). Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module —
as the checker or the CPU baseline, never as the thing shipped. The product path (ssl_vit_video_analytics_b200) never
imports it.

Pinning: the reference holds no tests or golden vectors, so this restatement is pinned against outputs
of the live reference generated in the authoring container by `oracle/make_golden.py` and committed under
`tests/golden/` (checked by tests/test_oracle_golden.py). The ViT-S tubelet model (`videomae_vit_*`) has NO reference
implementation: it is a composition of reference blocks defined here, and its parity is "vs in-repo restatement".
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------------------------------
# masks, patchify, loss
# ----------------------------------------------------------------------------------------------------------------------
def get_tube_mask(batch_size: int, num_frames: int, num_patches: int, mask_ratio: float) -> Tensor:
    """datasets/mae_loader.py:80-90 — per sample rand(L), mask the int(ratio*L) largest, repeat over frames."""
    k = int(mask_ratio * num_patches)
    rows = []
    for _ in range(batch_size):
        noise = torch.rand(num_patches)
        row = torch.zeros(num_patches)
        row[torch.argsort(noise, descending=True)[:k]] = 1
        rows.append(row)
    return torch.stack(rows).unsqueeze(1).repeat(1, num_frames, 1).bool()


def make_token_mask(B: int, T: int, P: int, ratio: float, mode: str, device="cpu") -> Tensor:
    """mae/masking.py:21-57."""
    N = T * P
    mask = torch.zeros((B, N), dtype=torch.bool, device=device)
    if mode == "random":
        k = max(1, int(round(N * ratio)))
        for b in range(B):
            mask[b, torch.randperm(N, device=device)[:k]] = True
        return mask
    if mode == "tube":
        k = max(1, int(round(P * ratio)))
        for b in range(B):
            s = torch.randperm(P, device=device)[:k]
            for t in range(T):
                mask[b, t * P + s] = True
        return mask
    raise ValueError(f"Unknown mask_mode: {mode}")


def patchify(imgs: Tensor, pt: int = 1, ps: int = 8) -> Tensor:
    """train_ssl_mae.py:26-31 for pt=1 (einsum 'bcthpwq->bthwpqc'); generalised to (pt,ps,ps) tubelets with element
    order (dt, py, px, c) for the tubelet model."""
    B, C, T, H, W = imgs.shape
    x = imgs.reshape(B, C, T // pt, pt, H // ps, ps, W // ps, ps)
    x = x.permute(0, 2, 4, 6, 3, 5, 7, 1)
    return x.reshape(B, (T // pt) * (H // ps) * (W // ps), pt * ps * ps * C)


def normalise_target(target: Tensor, unbiased: bool = True) -> Tensor:
    """train_ssl_mae.py:74-77 (torch.var default = unbiased); mae/losses.py:36-38 uses unbiased=False."""
    mean = target.mean(dim=-1, keepdim=True)
    var = target.var(dim=-1, keepdim=True, unbiased=unbiased)
    return (target - mean) / (var + 1.0e-6) ** 0.5


def masked_mse(pred: Tensor, clip: Tensor, mask: Tensor, pt: int = 1, ps: int = 8, norm_pix: bool = True) -> Tensor:
    """train_ssl_mae.py:72-84."""
    target = patchify(clip.float(), pt, ps)
    if norm_pix:
        target = normalise_target(target)
    loss_map = ((pred.float() - target) ** 2).mean(dim=-1)
    m = mask.reshape(mask.shape[0], -1)
    return (loss_map * m).sum() / (m.sum() + 1e-6)


def mae_l2_loss(pred: Tensor, target: Tensor, normalize: bool = True) -> Tensor:
    """mae/losses.py:19-41."""
    if normalize:
        target = normalise_target(target, unbiased=False)
    return ((pred - target) ** 2).mean()


def mae_l1_loss(pred: Tensor, target: Tensor) -> Tensor:
    """mae/losses.py:44-49."""
    return (pred - target).abs().mean()


def mae_cosine_loss(pred: Tensor, target: Tensor, eps: float = 1e-6) -> Tensor:
    """mae/losses.py:52-63."""
    p = F.normalize(pred, dim=-1, eps=eps)
    t = F.normalize(target, dim=-1, eps=eps)
    return 1.0 - (p * t).sum(dim=-1).mean()


def mae_recon_loss(pred: Tensor, target: Tensor, loss_type: str = "l2", normalize_target: bool = True) -> Tensor:
    """mae/losses.py:69-90 (keyword form used by mae/engine.py:87)."""
    if loss_type == "l2":
        return mae_l2_loss(pred, target, normalize_target)
    if loss_type == "l1":
        return mae_l1_loss(pred, target)
    if loss_type == "cosine":
        return mae_cosine_loss(pred, target)
    raise ValueError(f"Unknown MAE loss type: {loss_type}")


# ----------------------------------------------------------------------------------------------------------------------
# TinyViT encoder (models/tiny_vit.py)
# ----------------------------------------------------------------------------------------------------------------------
def _conv_bn(sd: SD, p: str, x: Tensor, stride: int, pad: int, groups: int, bn_train: bool) -> Tensor:
    """tiny_vit.py:12-18 — conv (no bias) then BatchNorm2d (eps 1e-5); bn_train uses batch statistics."""
    y = F.conv2d(x, sd[p + ".c.weight"], None, stride, pad, 1, groups)
    if bn_train:
        return F.batch_norm(y, None, None, sd[p + ".bn.weight"], sd[p + ".bn.bias"], True, 0.0, 1e-5)
    return F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"], sd[p + ".bn.bias"],
                        False, 0.0, 1e-5)


def _drop_path(y: Tensor, drop) -> Tensor:
    """timm DropPath (tiny_vit.py:5) with the per-sample keep mask supplied: y * keep[b] / keep_prob. drop = (keep_prob, keep[B])."""
    if drop is None:
        return y
    keep_prob, keep = drop
    return y * (keep.to(y.dtype).reshape((y.shape[0],) + (1,) * (y.ndim - 1)) / keep_prob)


def _mbconv(sd: SD, p: str, x: Tensor, stride: int, bn_train: bool, drop=None) -> Tensor:
    """tiny_vit.py:36-56 (+ SELayer :20-34); expand_ratio 4; residual iff stride 1 and in == out."""
    cin = x.shape[1]
    h = F.gelu(_conv_bn(sd, p + ".conv.0", x, 1, 0, 1, bn_train))
    mid = h.shape[1]
    h = F.gelu(_conv_bn(sd, p + ".conv.2", h, stride, 1, mid, bn_train))
    g = h.mean(dim=(2, 3))
    g = torch.sigmoid(F.linear(F.relu(F.linear(g, sd[p + ".conv.4.fc.0.weight"])), sd[p + ".conv.4.fc.2.weight"]))
    h = h * g[:, :, None, None]
    h = _conv_bn(sd, p + ".conv.5", h, 1, 0, 1, bn_train)
    return x + _drop_path(h, drop) if (stride == 1 and cin == h.shape[1]) else h


USE_SDPA = False   # True: F.scaled_dot_product_attention, the call the reference itself makes (tiny_vit.py:103); same arithmetic


def attention_core(qkv: Tensor, num_heads: int) -> Tensor:
    """tiny_vit.py:97-104 / the MHA core of nn.TransformerEncoderLayer: qkv [B, L, 3*H*hd] (q|k|v, head, dim)."""
    B, L, W3 = qkv.shape
    hd = W3 // (3 * num_heads)
    q, k, v = qkv.reshape(B, L, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    if USE_SDPA:
        return F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, num_heads * hd)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, num_heads * hd)


def transformer_block(x: Tensor, ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b,
                      num_heads: int, drop=None) -> Tensor:
    """Pre-LN block: tiny_vit.py:122-126 and nn.TransformerEncoderLayer(norm_first=True, gelu) with dropout inactive.
    drop = (keep_prob, keep_attn[B], keep_mlp[B]): the block's DropPath (:124-125), two draws of the same module, masks supplied."""
    D = x.shape[-1]
    d1 = None if drop is None else (drop[0], drop[1])
    d2 = None if drop is None else (drop[0], drop[2])
    h = F.layer_norm(x, (D,), ln1_w, ln1_b, 1e-5)
    x = x + _drop_path(F.linear(attention_core(F.linear(h, qkv_w, qkv_b), num_heads), proj_w, proj_b), d1)
    h = F.layer_norm(x, (D,), ln2_w, ln2_b, 1e-5)
    return x + _drop_path(F.linear(F.gelu(F.linear(h, fc1_w, fc1_b)), fc2_w, fc2_b), d2)


def _tinyvit_block(sd: SD, p: str, x: Tensor, num_heads: int, drop=None) -> Tensor:
    """tiny_vit.py:108-130 on an NCHW map."""
    B, C, H, W = x.shape
    t = x.flatten(2).transpose(1, 2)
    t = transformer_block(t, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], sd[p + ".attn.qkv.weight"], sd[p + ".attn.qkv.bias"],
                          sd[p + ".attn.proj.weight"], sd[p + ".attn.proj.bias"], sd[p + ".norm2.weight"], sd[p + ".norm2.bias"],
                          sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"], sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"],
                          num_heads, drop)
    return t.transpose(1, 2).reshape(B, C, H, W)


TINYVIT_DEPTHS = (2, 2, 6, 2)
TINYVIT_HEADS = (3, 6, 12, 18)


def tinyvit_forward(sd: SD, x: Tensor, prefix: str = "", num_stages: int = 3, bn_train: bool = True, drop_masks=None) -> Tensor:
    """tiny_vit.py:166-186 (forward_stage3 = 3 stages, forward = 4) for the 21M variant (:188-191).
    drop_masks: {"stages.S.J": (keep_prob, keep[B])} for MBConv blocks, (keep_prob, keep_attn[B], keep_mlp[B]) for TinyViTBlocks —
    the stochastic-depth masks of :146,158-162 supplied from outside (None = drop-path inactive)."""
    dm = drop_masks or {}
    p = prefix + "patch_embed.patch_embed"
    x = F.gelu(_conv_bn(sd, p + ".0", x, 2, 1, 1, bn_train))
    x = _conv_bn(sd, p + ".2", x, 1, 1, 1, bn_train)
    for s in range(num_stages):
        j = 0
        if s > 0:
            x = _mbconv(sd, f"{prefix}stages.{s}.{j}", x, 2, bn_train)
            j += 1
        for _ in range(TINYVIT_DEPTHS[s]):
            if s == 0:
                x = _mbconv(sd, f"{prefix}stages.{s}.{j}", x, 1, bn_train, dm.get(f"stages.{s}.{j}"))
            else:
                x = _tinyvit_block(sd, f"{prefix}stages.{s}.{j}", x, TINYVIT_HEADS[s], dm.get(f"stages.{s}.{j}"))
            j += 1
    return x


# ----------------------------------------------------------------------------------------------------------------------
# decoder side (models/mae_vit_adapter.py)
# ----------------------------------------------------------------------------------------------------------------------
def _decoder(sd: SD, x: Tensor, depth: int, heads: int) -> Tensor:
    """mae_vit_adapter.py:40-53,110-115."""
    for i in range(depth):
        p = f"decoder_blocks.layers.{i}"
        x = transformer_block(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], sd[p + ".self_attn.in_proj_weight"],
                              sd[p + ".self_attn.in_proj_bias"], sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"],
                              sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], sd[p + ".linear1.weight"], sd[p + ".linear1.bias"],
                              sd[p + ".linear2.weight"], sd[p + ".linear2.bias"], heads)
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd["decoder_norm.weight"], sd["decoder_norm.bias"], 1e-5)
    return F.linear(x, sd["decoder_pred.weight"], sd["decoder_pred.bias"])


def tiny_video_mae_forward(sd: SD, clip: Tensor, mask: Tensor, decoder_depth: int, decoder_heads: int, bn_train: bool = True,
                           drop_masks=None) -> Tensor:
    """mae_vit_adapter.py:75-117 with decoder dropout inactive; encoder drop-path only through supplied masks."""
    B, C, T, H, W = clip.shape
    frames = clip.permute(0, 2, 1, 3, 4).flatten(0, 1)
    latent = tinyvit_forward(sd, frames, "encoder.", 3, bn_train, drop_masks).flatten(2).transpose(1, 2)
    L = latent.shape[1]
    x = F.linear(latent, sd["enc_to_dec.weight"], sd["enc_to_dec.bias"]).view(B, T, L, -1)
    x = x + (sd["temporal_pos_embed"] + sd["spatial_pos_embed"])
    m = mask.unsqueeze(-1).type_as(x)
    x = (x * (1 - m) + sd["mask_token"] * m).flatten(1, 2)
    return _decoder(sd, x, decoder_depth, decoder_heads)


# ----------------------------------------------------------------------------------------------------------------------
# stack 2: latent-space MAE (models/tinyvit_mae.py); the timm encoder is out of scope, tokens are an input here
# ----------------------------------------------------------------------------------------------------------------------
def mae_decoder_forward(sd: SD, x: Tensor, depth: int, heads: int, prefix: str = "decoder.") -> Tensor:
    """tinyvit_mae.py:12-32 — proj -> nn.TransformerEncoder(norm_first, gelu) x depth -> pred (no final norm), dropout inactive."""
    x = F.linear(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])
    for i in range(depth):
        p = f"{prefix}blocks.layers.{i}"
        x = transformer_block(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], sd[p + ".self_attn.in_proj_weight"],
                              sd[p + ".self_attn.in_proj_bias"], sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"],
                              sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], sd[p + ".linear1.weight"], sd[p + ".linear1.bias"],
                              sd[p + ".linear2.weight"], sd[p + ".linear2.bias"], heads)
    return F.linear(x, sd[prefix + "pred.weight"], sd[prefix + "pred.bias"])


def tinyvit_mae_forward(sd: SD, tok: Tensor, token_mask: Tensor, depth: int, heads: int):
    """tinyvit_mae.py:104-114 from the encoder tokens tok [B, N, D] on: mask-token substitution (torch.where), decoder,
    row-major boolean gathers -> (pred [N_mask, D], target [N_mask, D] detached)."""
    B, N, D = tok.shape
    tok_in = torch.where(token_mask.unsqueeze(-1), sd["mask_token"].expand(B, N, D), tok)
    rec = mae_decoder_forward(sd, tok_in, depth, heads)
    return rec[token_mask], tok[token_mask].detach()


# ----------------------------------------------------------------------------------------------------------------------
# north-star ViT-S tubelet model (no reference implementation; composition of the blocks above)
# ----------------------------------------------------------------------------------------------------------------------
def videomae_vit_forward(sd: SD, clip: Tensor, mask: Tensor, depth: int, heads: int, decoder_depth: int, decoder_heads: int,
                         tubelet=(2, 16, 16)) -> Tensor:
    pt, ps, _ = tubelet
    B = clip.shape[0]
    tok = F.conv3d(clip, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=tubelet)  # [B, E, T', h, w]
    E, Tp = tok.shape[1], tok.shape[2]
    Lp = tok.shape[3] * tok.shape[4]
    tok = tok.flatten(3).permute(0, 2, 3, 1)                                                         # [B, T', L, E]
    tok = tok + (sd["enc_temporal_pos_embed"] + sd["enc_spatial_pos_embed"])
    vis = ~mask[:, 0]
    n_vis = int(vis[0].sum())
    idx = torch.stack([torch.nonzero(vis[b]).flatten() for b in range(B)])                           # ascending
    x = torch.gather(tok, 2, idx[:, None, :, None].expand(B, Tp, n_vis, E)).reshape(B, Tp * n_vis, E)
    for i in range(depth):
        p = f"blocks.{i}"
        x = transformer_block(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], sd[p + ".attn.qkv.weight"], sd[p + ".attn.qkv.bias"],
                              sd[p + ".attn.proj.weight"], sd[p + ".attn.proj.bias"], sd[p + ".norm2.weight"], sd[p + ".norm2.bias"],
                              sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"], sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"], heads)
    x = F.layer_norm(x, (E,), sd["norm.weight"], sd["norm.bias"], 1e-5)
    z = F.linear(x, sd["enc_to_dec.weight"], sd["enc_to_dec.bias"]).view(B, Tp, n_vis, -1)
    D = z.shape[-1]
    full = torch.zeros(B, Tp, Lp, D, dtype=z.dtype, device=z.device)
    full = full.scatter(2, idx[:, None, :, None].expand(B, Tp, n_vis, D), z)
    full = full + (sd["temporal_pos_embed"] + sd["spatial_pos_embed"])
    m = mask.unsqueeze(-1).type_as(full)
    x_dec = (full * (1 - m) + sd["mask_token"] * m).flatten(1, 2)
    return _decoder(sd, x_dec, decoder_depth, decoder_heads)


def videomae_vit_classifier_forward(sd: SD, clip: Tensor, depth: int, heads: int, tubelet=(2, 16, 16), sigma: float = 0.0, eps: Tensor = None):
    """BASELINE config 4 (composition, no reference implementation): the ViT-S tubelet encoder above on ALL tokens -> LayerNorm ->
    mean over tokens (the GAP + temporal mean of train_finetune.py:19-40) -> z + sigma * eps (feature_noise.py:4-7, eps supplied)
    -> nn.Linear head (heads.py:6-21). Keys: backbone.{patch_embed,enc_*_pos_embed,blocks,norm}.*, classifier.*. -> (logits, z)."""
    B = clip.shape[0]
    p = "backbone."
    tok = F.conv3d(clip, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=tubelet)
    E = tok.shape[1]
    tok = tok.flatten(3).permute(0, 2, 3, 1) + (sd[p + "enc_temporal_pos_embed"] + sd[p + "enc_spatial_pos_embed"])
    x = tok.reshape(B, -1, E)
    for i in range(depth):
        q = f"{p}blocks.{i}"
        x = transformer_block(x, sd[q + ".norm1.weight"], sd[q + ".norm1.bias"], sd[q + ".attn.qkv.weight"], sd[q + ".attn.qkv.bias"],
                              sd[q + ".attn.proj.weight"], sd[q + ".attn.proj.bias"], sd[q + ".norm2.weight"], sd[q + ".norm2.bias"],
                              sd[q + ".mlp.fc1.weight"], sd[q + ".mlp.fc1.bias"], sd[q + ".mlp.fc2.weight"], sd[q + ".mlp.fc2.bias"], heads)
    x = F.layer_norm(x, (E,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    z = x.mean(dim=1)
    if sigma > 0:
        z = add_gaussian_noise(z, sigma, eps)
    return F.linear(z, sd["classifier.weight"], sd["classifier.bias"]), z


# ----------------------------------------------------------------------------------------------------------------------
# feature noise, FedAvg
# ----------------------------------------------------------------------------------------------------------------------
def add_gaussian_noise(z: Tensor, sigma: float, eps: Tensor) -> Tensor:
    """privacy/feature_noise.py:4-7 with the normal draw supplied (exact-parity form)."""
    return z if sigma <= 0 else z + eps * sigma


def fedavg(global_state: SD, client_states, client_weights) -> SD:
    """federated/fed_loop.py:14-62 on plain state dicts."""
    total = float(sum(client_weights))
    out = {}
    for k, g in global_state.items():
        if any(k not in cs for cs in client_states):
            out[k] = g.detach().clone()
        elif torch.is_floating_point(g):
            acc = torch.zeros_like(g)
            for cs, w in zip(client_states, client_weights):
                acc += cs[k].to(acc.dtype) * (float(w) / total)
            out[k] = acc
        elif "num_batches_tracked" in k:
            out[k] = torch.stack([cs[k].to(torch.long) for cs in client_states]).max(dim=0).values
        else:
            out[k] = client_states[0][k].detach().clone()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# deterministic "perturbed init" (at the literal init the encoder output is identically zero)
# ----------------------------------------------------------------------------------------------------------------------
def perturb_state_dict(sd: SD, seed: int = 1234) -> SD:
    """BN gamma ~ U(0.5, 1.5); every bias (and BN beta) ~ N(0, 0.02^2); LayerNorm gamma ~ U(0.8, 1.2). Weights unchanged."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        v = v.clone()
        if not torch.is_floating_point(v) or "running_" in k:
            out[k] = v
            continue
        if k.endswith(".bn.weight"):
            v = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("bias"):
            v = torch.randn(v.shape, generator=g) * 0.02
        elif ("norm" in k) and k.endswith(".weight") and v.ndim == 1:
            v = torch.rand(v.shape, generator=g) * 0.4 + 0.8
        out[k] = v
    return out
