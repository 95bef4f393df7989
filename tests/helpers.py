"""Shared test helpers: model construction identical to oracle/make_golden.py, gradient comparison utilities."""
import torch

FULL_CFG = {"dataset": {"clip_len": 16, "image_size": 112},
            "model": {"decoder_embed_dim": 384, "decoder_depth": 4, "decoder_num_heads": 6, "decoder_dropout": 0.0}}
TINY_CFG = {"dataset": {"clip_len": 2, "image_size": 32},
            "model": {"decoder_embed_dim": 128, "decoder_depth": 2, "decoder_num_heads": 2, "decoder_dropout": 0.0}}
CFG1 = {"dataset": {"clip_len": 8, "image_size": 112},
        "model": {"decoder_embed_dim": 384, "decoder_depth": 4, "decoder_num_heads": 6, "decoder_dropout": 0.0}}


def build_ours(cfg, seed=0):
    from ssl_vit_video_analytics_b200.models import TinyVideoMAE, tiny_vit_21m_variant

    torch.manual_seed(seed)
    enc = tiny_vit_21m_variant(img_size=cfg["dataset"]["image_size"], use_checkpoint=False, drop_path_rate=0.0)
    return TinyVideoMAE(enc, cfg)


def checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def group_of(key: str) -> str:
    if key.startswith("encoder.patch_embed"):
        return "stem"
    if key.startswith("encoder.stages."):
        return "stage" + key.split(".")[2]
    if key.startswith("blocks."):
        return "encoder"
    if key.startswith("decoder_blocks"):
        return "decoder"
    return "head"


def grad_report(ours: dict, ref: dict):
    """global cosine, per-group cosine over the keys present in both."""
    keys = [k for k in ref if k in ours]
    glob = cosine(torch.cat([ours[k].flatten().double().cpu() for k in keys]), torch.cat([ref[k].flatten().double() for k in keys]))
    groups = {}
    for k in keys:
        groups.setdefault(group_of(k), []).append(k)
    per = {g: cosine(torch.cat([ours[k].flatten().double().cpu() for k in ks]), torch.cat([ref[k].flatten().double() for k in ks]))
           for g, ks in groups.items()}
    return glob, per
