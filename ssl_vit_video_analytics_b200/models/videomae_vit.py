"""North-star configuration (BASELINE.json): ViT-S/16 VideoMAE with tubelet embedding and encoder-side masking.

The reference has no such module (its MAE encoder is TinyViT and nothing is gathered — SURVEY.md §0 fact 2), so this
model is assembled from the reference's own blocks and keeps its `forward(x, mask) -> pred` contract:
  * tubelet embed = Conv3d(3, D, (2,16,16), stride (2,16,16)) evaluated on the VISIBLE tubes only
    (gather kernel -> tcgen05 GEMM with bias + positional rows in the epilogue);
  * encoder block = LayerNorm -> Attention -> Mlp exactly as tiny_vit.py:74-130 in token form (x12) + final LayerNorm;
  * enc_to_dec / factorised positional embeddings / mask-token substitution / decoder / decoder_norm / decoder_pred
    exactly as mae_vit_adapter.py:24-53,93-115, visible tokens scattered back to their positions;
  * loss = train_ssl_mae.py:72-84 generalised to (2,16,16) tubelets (`training_loss`).
Its fp32 oracle is `oracle/restatement.py::videomae_vit_forward` (a restatement, not the reference).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .mae_vit_adapter import DecoderLayer, DecoderStack, trunc_normal_
from .tiny_vit import Attention, Mlp


class TubeletEmbed(nn.Module):
    def __init__(self, in_chans, embed_dim, tubelet):
        super().__init__()
        self.tubelet = tuple(tubelet)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=self.tubelet, stride=self.tubelet)


def tubelet_embed(x, vis_idx, proj, tpos, spos, pt, ps, implicit=None):
    """Visible-tube Conv3d embedding + positional rows -> fp32 [B*T'*n_vis, E].
    ps == 16 (the VideoMAE geometry): ONE implicit-GEMM kernel gathering the clip by 5-D TMA boxes (`ops.TubeletEmbedFn`).
    Other patch sizes: gather kernel -> tcgen05 GEMM with the positional rows as fp32 residual (two kernels, same arithmetic)."""
    E = proj.weight.shape[0]
    if implicit is None:
        implicit = ops.tubelet_embed_supported(ps, E) and IMPLICIT_TUBELET_EMBED
    if implicit:
        return ops.TubeletEmbedFn.apply(x, vis_idx, proj.weight, proj.bias, tpos, spos, pt, ps)
    patches, pos_rows = ops.TubeletGatherFn.apply(x, vis_idx, tpos, spos, pt, ps)
    return ops.LinearResidualF32Fn.apply(patches, proj.weight, proj.bias, pos_rows)


IMPLICIT_TUBELET_EMBED = True   # A/B switch for measurements (bench.py --embed gather)


class EncoderBlock(nn.Module):
    """TinyViTBlock in token form (no NCHW round trip): norm1, attn{qkv,proj}, norm2, mlp{fc1,fc2}."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, key_dim=dim // num_heads, num_heads=num_heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))

    def forward_tokens(self, t, B, Lq):
        return ops.transformer_block(t, self.norm1, self.attn.qkv.weight, self.attn.qkv.bias, self.attn.proj, self.norm2,
                                     self.mlp.fc1, self.mlp.fc2, self.attn.num_heads, B, Lq)


DEFAULT_CONFIG = {
    "dataset": {"clip_len": 16, "image_size": 224},
    "model": {
        "embed_dim": 384, "depth": 12, "num_heads": 6, "tubelet": [2, 16, 16],
        "decoder_embed_dim": 384, "decoder_depth": 4, "decoder_num_heads": 6, "decoder_dropout": 0.0,
    },
    "ssl": {"mask_ratio": 0.9, "norm_pix_loss": True},
}


class VideoMAEViT(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        cfg = config or DEFAULT_CONFIG
        m = cfg["model"]
        self.clip_len = cfg["dataset"]["clip_len"]
        self.img_size = cfg["dataset"]["image_size"]
        self.tubelet = tuple(m.get("tubelet", (2, 16, 16)))
        pt, ps, ps2 = self.tubelet
        assert ps == ps2, "square spatial patches only"
        self.frames_out = self.clip_len // pt
        self.num_patches = (self.img_size // ps) ** 2
        self.embed_dim = m.get("embed_dim", 384)
        self.num_heads = m.get("num_heads", 6)
        self.depth = m.get("depth", 12)
        self.decoder_embed_dim = m["decoder_embed_dim"]
        self.decoder_num_heads = m["decoder_num_heads"]
        self.decoder_depth = m["decoder_depth"]
        self.norm_pix_loss = bool(cfg.get("ssl", {}).get("norm_pix_loss", True))
        self.feature_noise_sigma = float(cfg.get("privacy", {}).get("feature_noise_sigma", 0.0))   # configs/privacy.yaml:32 sweeps 0.05 / 0.1 / 0.2
        self.noise_eps_out = None   # tests: an fp32 [B * visible tokens, D] tensor that receives the normals drawn by the epilogue
        E, D = self.embed_dim, self.decoder_embed_dim

        self.patch_embed = TubeletEmbed(3, E, self.tubelet)
        self.enc_temporal_pos_embed = nn.Parameter(torch.zeros(1, self.frames_out, 1, E))
        self.enc_spatial_pos_embed = nn.Parameter(torch.zeros(1, 1, self.num_patches, E))
        self.blocks = nn.ModuleList([EncoderBlock(E, self.num_heads) for _ in range(self.depth)])
        self.norm = nn.LayerNorm(E)

        self.enc_to_dec = nn.Linear(E, D, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, D))
        self.temporal_pos_embed = nn.Parameter(torch.zeros(1, self.frames_out, 1, D))
        self.spatial_pos_embed = nn.Parameter(torch.zeros(1, 1, self.num_patches, D))
        self.decoder_blocks = DecoderStack(
            lambda: DecoderLayer(D, self.decoder_num_heads, int(D * 4), float(m.get("decoder_dropout", 0.0))), self.decoder_depth)
        self.decoder_norm = nn.LayerNorm(D)
        self.decoder_pred = nn.Linear(D, pt * ps * ps * 3, bias=True)
        self.initialize_weights()

    def initialize_weights(self):
        for p in (self.enc_temporal_pos_embed, self.enc_spatial_pos_embed, self.temporal_pos_embed, self.spatial_pos_embed,
                  self.mask_token):
            trunc_normal_(p, std=0.02)
        w = self.patch_embed.proj.weight
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.constant_(self.patch_embed.proj.bias, 0)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ------------------------------------------------------------------------------------------------------------
    def visible_indices(self, mask, n_vis=None):
        """mask [B, T', L] bool (tube mask: identical over T') -> int32 [B, n_vis], ascending visible patch ids.
        n_vis defaults to the count in clip 0 (one host sync); pass it to stay sync-free (CUDA-graph capture)."""
        frame = mask[:, 0].to(torch.uint8).contiguous()
        if n_vis is None:
            # one host sync anyway: also check that this IS a tube mask with the same visible count in every clip (a
            # `make_token_mask(mode="random")` mask would give the encoder and the decoder different token sets)
            counts = (frame == 0).sum(dim=1)
            ok = bool(((mask == mask[:, :1]).all() & (counts == counts[0]).all()).item())
            if not ok:
                raise ValueError("VideoMAEViT needs a tube mask (identical over T') with the same number of visible patches in "
                                 "every clip, e.g. datasets.mae_loader.get_tube_mask")
            n_vis = int(counts[0].item())
        return ops.mask_to_indices(frame, n_vis)

    def forward(self, x, mask, vis_idx=None):
        """x: fp32 [B, 3, T, H, W]; mask: bool [B, T', L] (True = masked) -> pred [B, T'*L, pt*ps*ps*3]."""
        B = x.shape[0]
        Tp, Lp, E, D = self.frames_out, self.num_patches, self.embed_dim, self.decoder_embed_dim
        pt, ps, _ = self.tubelet
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        if vis_idx is None:
            vis_idx = self.visible_indices(mask)
        with ops.weight_cache_scope(self):   # one launch: bf16 operand copies of every weight matrix, valid for this forward
            return self._forward_tokens(x, mask, vis_idx, B, Tp, Lp, pt, ps)

    def _forward_tokens(self, x, mask, vis_idx, B, Tp, Lp, pt, ps):
        n_vis = vis_idx.shape[1]
        Lv = Tp * n_vis
        t = tubelet_embed(x, vis_idx, self.patch_embed.proj, self.enc_temporal_pos_embed, self.enc_spatial_pos_embed, pt, ps)
        for blk in self.blocks:
            t = blk.forward_tokens(t, B, Lv)
        z = ops.LayerNormFn.apply(t, self.norm.weight, self.norm.bias, self.norm.eps, ops.BF16)
        # encoder output: enc_to_dec, with the privacy feature perturbation z' = z + N(0, sigma^2 I) (feature_noise.py:4-7) drawn in
        # the GEMM epilogue when `privacy.feature_noise_sigma` > 0 (no extra pass over z)
        z = ops.linear_with_feature_noise(z, self.enc_to_dec.weight, self.enc_to_dec.bias, self.feature_noise_sigma, eps_out=self.noise_eps_out)
        m8 = mask.reshape(B, Tp * Lp).to(torch.uint8).contiguous()
        x_dec = ops.DecoderInputFn.apply(z, vis_idx, m8, self.temporal_pos_embed, self.spatial_pos_embed, self.mask_token,
                                         B, Tp, Lp)
        x_dec = self.decoder_blocks(x_dec)
        y = ops.LayerNormFn.apply(x_dec, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_norm.eps, ops.BF16)
        return ops.linear(y, self.decoder_pred.weight, self.decoder_pred.bias)

    def training_loss(self, x, mask, vis_idx=None):
        """pred + fused normalised-pixel masked MSE (one scalar); the path bench.py times."""
        pt, ps, _ = self.tubelet
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = self.forward(x, mask, vis_idx)
        return ops.masked_mse(pred, x if x.dtype == torch.float32 else x.float(), mask, pt, ps, 1 if self.norm_pix_loss else 0)


# ======================================================================================================================
# BASELINE config 4: supervised finetune of the ViT-S encoder (UCF101-shaped, 101 classes) with feature noise
# ======================================================================================================================
ENCODER_KEYS = ("patch_embed.", "enc_temporal_pos_embed", "enc_spatial_pos_embed", "blocks.", "norm.")


class VideoMAEViTClassifier(nn.Module):
    """The reference's finetune pattern (train_finetune.py:19-40,84-124; heads.py:6-21: encode -> global average pool -> temporal
    mean -> nn.Linear -> CrossEntropyLoss) on the ViT-S/16 tubelet encoder that `VideoMAEViT` pre-trains, all tokens visible:

        clip -> tubelet embed + positional rows (all 8 x 196 tubes) -> 12 encoder blocks -> LayerNorm
             -> mean over the 1568 tokens (+ feature noise z' = z + N(0, sigma^2 I), feature_noise.py:4-7, in the same kernel)
             -> 384 -> num_classes head on the tcgen05 GEMM -> fused softmax cross-entropy (`loss`).

    Encoder parameter names equal VideoMAEViT's (`patch_embed.*`, `enc_*_pos_embed`, `blocks.*`, `norm.*`) under the prefix
    `backbone.`, so `load_pretrained_mae(mae.state_dict())` is the MAE -> finetune hand-off (train_finetune.py:43-63)."""

    def __init__(self, num_classes: int = 101, config=None, noise_sigma: float = 0.0):
        super().__init__()
        cfg = config or DEFAULT_CONFIG
        m = cfg["model"]
        self.clip_len = cfg["dataset"]["clip_len"]
        self.img_size = cfg["dataset"]["image_size"]
        self.tubelet = tuple(m.get("tubelet", (2, 16, 16)))
        pt, ps, _ = self.tubelet
        self.frames_out = self.clip_len // pt
        self.num_patches = (self.img_size // ps) ** 2
        E = self.embed_dim = m.get("embed_dim", 384)
        self.backbone = nn.Module()
        bb = self.backbone
        bb.patch_embed = TubeletEmbed(3, E, self.tubelet)
        bb.enc_temporal_pos_embed = nn.Parameter(torch.zeros(1, self.frames_out, 1, E))
        bb.enc_spatial_pos_embed = nn.Parameter(torch.zeros(1, 1, self.num_patches, E))
        bb.blocks = nn.ModuleList([EncoderBlock(E, m.get("num_heads", 6)) for _ in range(m.get("depth", 12))])
        bb.norm = nn.LayerNorm(E)
        self.classifier = nn.Linear(E, num_classes)
        self.noise_sigma = float(noise_sigma)
        self.noise_eps_out = None
        trunc_normal_(bb.enc_temporal_pos_embed, std=0.02)
        trunc_normal_(bb.enc_spatial_pos_embed, std=0.02)
        w = bb.patch_embed.proj.weight
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.constant_(bb.patch_embed.proj.bias, 0)
        self.apply(VideoMAEViT._init_weights)
        self._all_idx = None

    def load_pretrained_mae(self, mae_state: dict, strict: bool = True):
        """Copy the encoder tensors of a `VideoMAEViT` state_dict into `.backbone` (decoder-side keys are ignored)."""
        sd = {k: v for k, v in mae_state.items() if k.startswith(ENCODER_KEYS)}
        return self.backbone.load_state_dict(sd, strict=strict)

    def video_embedding(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [B, 3, T, H, W] -> fp32 [B, E]: mean over all tokens of the normalised encoder output (+ feature noise)."""
        if not x.is_cuda:
            raise RuntimeError("VideoMAEViTClassifier runs on CUDA (sm_100a) tensors only; there is no CPU path")
        B = x.shape[0]
        bb = self.backbone
        Tp, Lp, E = self.frames_out, self.num_patches, self.embed_dim
        pt, ps, _ = self.tubelet
        x = x.float().contiguous()
        if self._all_idx is None or self._all_idx.shape[0] != B or self._all_idx.device != x.device:
            self._all_idx = torch.arange(Lp, device=x.device, dtype=torch.int32).repeat(B, 1).contiguous()   # every tube is visible
        with ops.weight_cache_scope(self):
            t = tubelet_embed(x, self._all_idx, bb.patch_embed.proj, bb.enc_temporal_pos_embed, bb.enc_spatial_pos_embed, pt, ps)
            for blk in bb.blocks:
                t = blk.forward_tokens(t, B, Tp * Lp)
            z = ops.LayerNormFn.apply(t, bb.norm.weight, bb.norm.bias, bb.norm.eps, ops.BF16)
            sigma = self.noise_sigma if self.noise_sigma > 0 else 0.0
            return ops.meanpool_with_feature_noise(z.view(B, Tp * Lp, E), sigma, eps_out=self.noise_eps_out)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .video_classifier import linear_head

        with torch.autocast("cuda", dtype=torch.bfloat16):
            z = self.video_embedding(x)
            return linear_head(z, self.classifier)

    @staticmethod
    def loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return ops.softmax_cross_entropy(logits, labels)
