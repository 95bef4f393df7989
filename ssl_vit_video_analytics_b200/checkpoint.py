"""Checkpoint hand-off between this package and the reference: the file formats and key-prefix conventions
of the reference, so that checkpoints written by either side load in the other.

  * stack 1 writes the BARE encoder state_dict every 10 epochs as `encoder_ep{N}.pth` (train_ssl_mae.py:190-194 through
    utils.save_checkpoint, utils.py:73-77);
  * finetuning reads a checkpoint that is either a bare state_dict or a wrapper {"model": ...} and maps `encoder.*` keys
    onto the backbone (train_finetune.py:43-62);
  * the federated code accepts wrappers keyed "model" | "state_dict" | "net" | "weights" (federated/utils_fed.py:31-44,50-71);
  * stack 2 saves {"epoch", "model", "optimizer", "scaler", "scheduler", "best", "cfg_hash"} (mae/utils.py:56-58,73-92).
Pure host code (torch.save / torch.load); tensors are moved to CPU on save so files are device-independent.
"""
from __future__ import annotations

import hashlib
import json
import os
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import torch

WRAPPER_KEYS = ("model", "state_dict", "net", "weights")


def _cpu_state(sd: Dict[str, Any]) -> Dict[str, Any]:
    return {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in sd.items()}


def save_checkpoint(state: Dict[str, Any], filename) -> None:
    """utils.save_checkpoint (utils.py:73-77): create the directory, torch.save the object as is."""
    filename = Path(filename)
    filename.parent.mkdir(parents=True, exist_ok=True)
    torch.save(state, filename)


def encoder_checkpoint_path(save_dir, epoch: int) -> Path:
    return Path(save_dir) / f"encoder_ep{epoch}.pth"


def save_encoder_checkpoint(model, save_dir, epoch: int) -> Path:
    """What stack 1 writes: the bare `model.encoder.state_dict()` (train_ssl_mae.py:190-194)."""
    path = encoder_checkpoint_path(save_dir, epoch)
    save_checkpoint(_cpu_state(model.encoder.state_dict()), path)
    return path


def extract_state_dict(ckpt_obj) -> Optional[Dict[str, Any]]:
    """federated/utils_fed.py:31-44: unwrap {"model" | "state_dict" | "net" | "weights": {...}} or accept a bare state_dict."""
    if ckpt_obj is None:
        return None
    if isinstance(ckpt_obj, dict):
        for k in WRAPPER_KEYS:
            if k in ckpt_obj and isinstance(ckpt_obj[k], dict):
                return ckpt_obj[k]
        if any(torch.is_tensor(v) for v in ckpt_obj.values()):
            return ckpt_obj
    return None


def strip_prefix(state: Dict[str, Any], prefix: str = "encoder.") -> Dict[str, Any]:
    """Keep only the keys under `prefix` and drop it (train_finetune.py:51-55: `encoder.*` -> backbone keys)."""
    return {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}


def add_prefix(state: Dict[str, Any], prefix: str = "encoder.") -> Dict[str, Any]:
    return {prefix + k: v for k, v in state.items()}


def load_encoder_checkpoint(encoder: torch.nn.Module, ckpt, strict: bool = True) -> Tuple[list, list]:
    """Load an encoder from any of the reference's formats: a path or an object; bare encoder keys, `encoder.`-prefixed keys
    of a whole TinyVideoMAE state_dict, or one of the wrappers above. Returns (missing, unexpected) like load_state_dict."""
    if isinstance(ckpt, (str, os.PathLike)):
        ckpt = torch.load(ckpt, map_location="cpu")
    state = extract_state_dict(ckpt)
    if state is None:
        raise ValueError("unrecognised checkpoint format: expected a state_dict or a wrapper keyed " + " | ".join(WRAPPER_KEYS))
    if any(k.startswith("encoder.") for k in state):
        state = strip_prefix(state, "encoder.")
    res = encoder.load_state_dict(state, strict=strict)
    return list(res.missing_keys), list(res.unexpected_keys)


def load_pretrained_ssl(model, ckpt_path) -> bool:
    """train_finetune.py:43-62: load `encoder.*` weights of an SSL checkpoint into `model.backbone` (non-strict); False when the
    path is missing. Bare `encoder_ep{N}.pth` files (no prefix) are accepted as well — the reference silently loads nothing
    from them, which is its hand-off bug, fixed on our side only."""
    if ckpt_path is None or not os.path.isfile(ckpt_path):
        return False
    ckpt = torch.load(ckpt_path, map_location="cpu")
    state = ckpt.get("model", ckpt) if isinstance(ckpt, dict) else ckpt
    backbone_state = strip_prefix(state, "encoder.") if any(k.startswith("encoder.") for k in state) else dict(state)
    model.backbone.load_state_dict(backbone_state, strict=False)
    return True


def config_hash(cfg: Dict[str, Any]) -> str:
    """mae/utils.py:56-58"""
    s = json.dumps(cfg, sort_keys=True, ensure_ascii=False)
    return hashlib.sha256(s.encode("utf-8")).hexdigest()[:12]


def save_mae_checkpoint(path, model, optimizer, scaler, scheduler, epoch: int, best: Dict[str, Any], cfg: Dict[str, Any]) -> None:
    """Stack-2 training checkpoint (mae/utils.py:73-92), same dictionary layout."""
    ckpt = {
        "epoch": epoch,
        "model": _cpu_state(model.state_dict()),
        "optimizer": optimizer.state_dict(),
        "scaler": (scaler.state_dict() if scaler is not None else None),
        "scheduler": (scheduler.state_dict() if scheduler is not None else None),
        "best": best,
        "cfg_hash": config_hash(cfg),
    }
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    torch.save(ckpt, path)


def load_mae_checkpoint(path, model, optimizer=None, scaler=None, scheduler=None, map_location="cpu") -> Dict[str, Any]:
    """Resume from a stack-2 checkpoint; returns the dictionary (epoch, best, cfg_hash ...)."""
    ckpt = torch.load(path, map_location=map_location)
    model.load_state_dict(ckpt["model"])
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    if scaler is not None and ckpt.get("scaler") is not None:
        scaler.load_state_dict(ckpt["scaler"])
    if scheduler is not None and ckpt.get("scheduler") is not None:
        scheduler.load_state_dict(ckpt["scheduler"])
    return ckpt
