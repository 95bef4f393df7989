"""ORACLE tooling — generate tests/golden/* by running the LIVE reference (authoring container only).

    python -m oracle.make_golden        # needs /root/reference (or VMAE_REFERENCE_ROOT)

Writes small fixtures (a few hundred KB in total):
  tests/golden/init_checksums.json   per-tensor (sum, abs-sum) of TinyVideoMAE's state_dict built under seed 0
  tests/golden/stack1_tiny.pt        reference pred / loss / gradients for a tiny clip, perturbed init
  tests/golden/host_fns.pt           get_tube_mask, make_token_mask, patchify, stack-2 losses, fedavg known answers
"""
from __future__ import annotations

import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_shim, restatement as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

FULL_CFG = {"dataset": {"clip_len": 16, "image_size": 112}, "model": {"decoder_embed_dim": 384, "decoder_depth": 4, "decoder_num_heads": 6}}
TINY_CFG = {"dataset": {"clip_len": 2, "image_size": 32}, "model": {"decoder_embed_dim": 128, "decoder_depth": 2, "decoder_num_heads": 2}}


def build_reference(ref, cfg, seed=0):
    torch.manual_seed(seed)
    enc = ref.tiny_vit.tiny_vit_21m_variant(img_size=cfg["dataset"]["image_size"], use_checkpoint=False, drop_path_rate=0.0)
    model = ref.mae_vit_adapter.TinyVideoMAE(enc, cfg)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return model


def checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def main():
    ref = reference_shim.load()
    os.makedirs(GOLD, exist_ok=True)

    # ---- 1. init checksums (full config, seed 0)
    model = build_reference(ref, FULL_CFG)
    with open(os.path.join(GOLD, "init_checksums.json"), "w") as f:
        json.dump({"config": FULL_CFG, "seed": 0, "num_keys": len(model.state_dict()), "checksums": checksums(model.state_dict())}, f)

    # ---- 2. tiny fwd/bwd, perturbed init, train-mode BN, dropout / drop-path off
    model = build_reference(ref, TINY_CFG)
    sd = R.perturb_state_dict(model.state_dict(), seed=1234)
    model.load_state_dict(sd)
    model.train()
    g = torch.Generator().manual_seed(7)
    clip = torch.randn(2, 3, 2, 32, 32, generator=g)
    torch.manual_seed(42)
    mask = ref.mae_loader.get_tube_mask(2, 2, 16, 0.75)
    pred = model(clip, mask)
    target = reference_shim.patchify_reference(clip, 8)
    mean, var = target.mean(dim=-1, keepdim=True), target.var(dim=-1, keepdim=True)
    target = (target - mean) / (var + 1.0e-6) ** 0.5
    loss_map = ((pred - target) ** 2).mean(dim=-1)
    mflat = mask.flatten(1, 2)
    loss = (loss_map * mflat).sum() / (mflat.sum() + 1e-6)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    keep = ["decoder_pred.bias", "decoder_pred.weight", "mask_token", "temporal_pos_embed", "spatial_pos_embed", "enc_to_dec.bias",
            "decoder_norm.weight", "decoder_blocks.layers.0.self_attn.in_proj_bias", "decoder_blocks.layers.1.linear1.bias",
            "encoder.stages.2.6.mlp.fc2.bias", "encoder.stages.1.1.norm1.weight", "encoder.patch_embed.patch_embed.0.bn.weight"]
    torch.save({
        "config": TINY_CFG, "seed": 0, "perturb_seed": 1234, "clip_seed": 7, "mask_seed": 42, "mask_ratio": 0.75,
        "clip": clip, "mask": mask, "pred": pred.detach(), "loss": loss.detach(),
        "grad_norms": {k: float(v.norm()) for k, v in grads.items()},
        "grads": {k: grads[k] for k in keep},
        "state_checksums": checksums(sd),
    }, os.path.join(GOLD, "stack1_tiny.pt"))

    # ---- 3. host functions
    out = {}
    torch.manual_seed(42)
    out["tube_mask_b4_t3_l196_r0.9_seed42"] = ref.mae_loader.get_tube_mask(4, 3, 196, 0.9)
    torch.manual_seed(5)
    out["tube_mask_b2_t8_l49_r0.5_seed5"] = ref.mae_loader.get_tube_mask(2, 8, 49, 0.5)
    torch.manual_seed(11)
    out["token_mask_tube_b3_t4_p9_r0.6_seed11"] = ref.masking.make_token_mask(3, 4, 9, 0.6, "tube", torch.device("cpu"))
    torch.manual_seed(12)
    out["token_mask_random_b3_t4_p9_r0.6_seed12"] = ref.masking.make_token_mask(3, 4, 9, 0.6, "random", torch.device("cpu"))
    sched = [{"start": 1, "end": 5, "value": 0.5}, {"start": 5, "end": 9, "value": 0.75}]
    out["mask_ratio_schedule"] = [ref.masking.get_mask_ratio(e, sched, 0.9) for e in range(0, 12)]
    imgs = torch.arange(2 * 3 * 2 * 16 * 16, dtype=torch.float32).reshape(2, 3, 2, 16, 16)
    out["patchify_arange_2x3x2x16x16_p8"] = reference_shim.patchify_reference(imgs, 8)
    g = torch.Generator().manual_seed(3)
    p, t = torch.randn(37, 48, generator=g), torch.randn(37, 48, generator=g) * 3 + 1
    out["loss_inputs"] = (p, t)
    out["mae_l2_norm"] = ref.losses.mae_l2_loss(p, t, True)
    out["mae_l2_raw"] = ref.losses.mae_l2_loss(p, t, False)
    out["mae_l1"] = ref.losses.mae_l1_loss(p, t)
    out["mae_cos"] = ref.losses.mae_cosine_loss(p, t)
    out["mae_recon_positional"] = ref.losses.mae_recon_loss(p, t, "l1", False)
    out["recon_stats"] = ref.losses.reconstruction_error_stats(p, t)
    # fedavg on a small BN model
    def small():
        return nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4), nn.Flatten(), nn.LazyLinear(5)) if False else nn.Sequential(
            nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4))
    torch.manual_seed(21)
    glob = small()
    clients = []
    for i in range(3):
        m = small()
        m[1].num_batches_tracked.fill_(i + 2)
        clients.append({k: v.detach().clone() for k, v in m.state_dict().items()})
    weights = [1907.0, 1893.0, 1921.0]
    out["fedavg_clients"] = clients
    out["fedavg_weights"] = weights
    out["fedavg_global_before"] = {k: v.detach().clone() for k, v in glob.state_dict().items()}
    out["fedavg_result"] = ref.fed_loop.fedavg_aggregate(glob, clients, weights)
    z = torch.randn(4, 8, generator=g)
    out["noise_sigma0_is_identity"] = ref.feature_noise.add_gaussian_noise(z, 0.0) is z
    torch.save(out, os.path.join(GOLD, "host_fns.pt"))
    sizes = {f: os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD)}
    print("golden written:", sizes)


if __name__ == "__main__":
    main()
