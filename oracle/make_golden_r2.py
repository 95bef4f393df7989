"""ORACLE tooling — second set of golden vectors from the LIVE reference (authoring container only).

    python -m oracle.make_golden_r2        # needs /root/reference (or VMAE_REFERENCE_ROOT)

  tests/golden/stack1_defaults.pt   TinyVideoMAE with the reference's literal encoder defaults — `use_checkpoint=True`
                                    (train_ssl_mae.py:144) and `drop_path_rate=0.1` (tiny_vit.py:137) — one train-mode step on a
                                    tiny clip: the stochastic-depth keep masks the reference drew (recorded from its DropPath
                                    modules), pred, loss, selected gradients, all gradient norms, and the BatchNorm counters /
                                    running statistics after the step (checkpointed stages update them twice).
                                    Decoder dropout is set to 0 (its realisation is RNG-stream specific).
  tests/golden/stack2_mae.pt        `TinyViTMAE.forward` (tinyvit_mae.py:82-115) and `MAEDecoder` (:12-32) executed from the
                                    reference source with the timm-backed `TinyViTBackbone` replaced by a stub that returns
                                    recorded tokens (timm is not installed and its arithmetic is out of scope):
                                    tokens, token mask, gathered pred / target rows, loss (mae_l2_loss) and gradients.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_shim, restatement as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TINY_CFG = {"dataset": {"clip_len": 2, "image_size": 32}, "model": {"decoder_embed_dim": 128, "decoder_depth": 2, "decoder_num_heads": 2}}


def checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def stack1_defaults(ref):
    torch.manual_seed(0)
    enc = ref.tiny_vit.tiny_vit_21m_variant(img_size=32)            # literal defaults: use_checkpoint=True, drop_path_rate=0.1
    assert enc.use_checkpoint is True
    model = ref.mae_vit_adapter.TinyVideoMAE(enc, TINY_CFG)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    sd = R.perturb_state_dict(model.state_dict(), seed=4321)
    model.load_state_dict(sd)
    model.train()

    # record the keep masks in the FIRST forward only (the non-reentrant checkpoint recompute replays the same RNG state)
    drop_cls = sys.modules["timm.layers"].DropPath
    recorded, recording = {}, [True]
    names = {id(m): n for n, m in model.encoder.named_modules() if isinstance(m, drop_cls)}

    def hooked(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if recording[0]:
            recorded.setdefault(names[id(self)], []).append(m.flatten().clone())
        if keep > 0.0 and self.scale_by_keep:
            m = m / keep
        return x * m

    orig = drop_cls.forward
    drop_cls.forward = hooked
    try:
        g = torch.Generator().manual_seed(17)
        clip = torch.randn(3, 3, 2, 32, 32, generator=g)
        torch.manual_seed(43)
        mask = ref.mae_loader.get_tube_mask(3, 2, 16, 0.75)
        torch.manual_seed(99)                                       # the DropPath draws
        pred = model(clip, mask)
        recording[0] = False
        target = reference_shim.patchify_reference(clip, 8)
        mean, var = target.mean(dim=-1, keepdim=True), target.var(dim=-1, keepdim=True)
        target = (target - mean) / (var + 1.0e-6) ** 0.5
        mflat = mask.flatten(1, 2)
        loss = ((((pred - target) ** 2).mean(dim=-1)) * mflat).sum() / (mflat.sum() + 1e-6)
        loss.backward()
    finally:
        drop_cls.forward = orig
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    keep = ["decoder_pred.bias", "mask_token", "enc_to_dec.bias", "encoder.stages.2.6.mlp.fc2.bias", "encoder.stages.2.1.attn.qkv.bias",
            "encoder.stages.1.1.attn.proj.weight", "encoder.stages.1.1.norm1.weight", "encoder.stages.0.1.conv.0.bn.weight", "encoder.patch_embed.patch_embed.0.bn.weight"]
    after = model.state_dict()
    drop_masks = {}
    for name, lst in recorded.items():                               # "stages.S.J.drop_path" -> "stages.S.J"
        key = name[: -len(".drop_path")]
        p = float(dict(model.encoder.named_modules())[name].drop_prob)
        drop_masks[key] = (1.0 - p, *lst)
    torch.save({
        "config": TINY_CFG, "seed": 0, "perturb_seed": 4321, "clip": clip, "mask": mask, "drop_masks": drop_masks,
        "pred": pred.detach(), "loss": loss.detach(), "grads": {k: grads[k] for k in keep},
        "grad_norms": {k: float(v.norm()) for k, v in grads.items()},
        "num_batches_tracked": {k: int(v) for k, v in after.items() if k.endswith("num_batches_tracked")},
        "running_checksums": checksums({k: v for k, v in after.items() if "running_" in k}),
    }, os.path.join(GOLD, "stack1_defaults.pt"))
    return drop_masks


def stack2_mae(ref_root):
    """Run the reference's tinyvit_mae.py with `models.tinyvit_backbone` replaced by a token-replaying stub."""
    import importlib.util

    D, B, T, P = 64, 2, 4, 9
    g = torch.Generator().manual_seed(23)
    tokens = torch.randn(B * T, P, D, generator=g) * 0.7 + 0.1

    class TinyViTBackbone(nn.Module):                                # stub: replays recorded tokens (timm is out of scope)
        def __init__(self, backbone="", pretrained=""):
            super().__init__()
            self.scale = nn.Parameter(torch.ones(()))                # so that the "encoder" has a parameter and tok needs grad

        def forward(self, x, pool_size=3):
            if x.shape[0] == 1:                                      # the constructor's dummy forward (:64-68)
                return torch.zeros(1, pool_size * pool_size, D)
            return tokens * self.scale

    pkg = types.ModuleType("_ref_models")
    pkg.__path__ = [os.path.join(ref_root, "src", "models")]
    sys.modules["_ref_models"] = pkg
    stub = types.ModuleType("_ref_models.tinyvit_backbone")
    stub.TinyViTBackbone = TinyViTBackbone
    sys.modules["_ref_models.tinyvit_backbone"] = stub
    spec = importlib.util.spec_from_file_location("_ref_models.tinyvit_mae", os.path.join(ref_root, "src", "models", "tinyvit_mae.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_models.tinyvit_mae"] = mod
    spec.loader.exec_module(mod)
    losses = reference_shim._load_by_path("_ref_losses2", "src/mae/losses.py")
    masking = reference_shim._load_by_path("_ref_masking2", "src/mae/masking.py")

    torch.manual_seed(5)
    model = mod.TinyViTMAE(backbone_name="stub", pretrained="", stage4_pool=3, decoder_dim=64, decoder_depth=2, decoder_num_heads=1, mlp_ratio=2.0)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    sd = R.perturb_state_dict(model.state_dict(), seed=77)
    model.load_state_dict(sd)
    model.train()
    out = {"tokens": tokens, "B": B, "T": T, "P": P, "D": D, "decoder": {"decoder_dim": 64, "depth": 2, "num_heads": 1, "mlp_ratio": 2.0}, "state": {k: v.clone() for k, v in sd.items()}, "cases": {}}
    clip = torch.zeros(B, 3, T, 8, 8)
    for mode, seed in (("tube", 31), ("random", 32)):
        torch.manual_seed(seed)
        tm = masking.make_token_mask(B, T, P, 0.6, mode, torch.device("cpu"))
        model.zero_grad(set_to_none=True)
        pred, target = model(clip, tm)
        loss = losses.mae_l2_loss(pred, target, True)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        out["cases"][mode] = {"token_mask": tm, "pred": pred.detach().clone(), "target": target.detach().clone(), "loss": loss.detach().clone(),
                              "grad_norms": {k: float(v.norm()) for k, v in grads.items()},
                              "grads": grads if mode == "tube" else {k: grads[k] for k in ("mask_token", "encoder.scale", "decoder.pred.bias")}}
    torch.save(out, os.path.join(GOLD, "stack2_mae.pt"))


def main():
    ref = reference_shim.load()
    os.makedirs(GOLD, exist_ok=True)
    dm = stack1_defaults(ref)
    print("stack1_defaults: drop-path masks for", sorted(dm))
    stack2_mae(reference_shim.REF_ROOT)
    print({f: os.path.getsize(os.path.join(GOLD, f)) for f in sorted(os.listdir(GOLD))})


if __name__ == "__main__":
    main()
