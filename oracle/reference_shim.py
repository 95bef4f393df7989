"""ORACLE tooling — imports the LIVE reference (read-only checkout, default /root/reference) so that golden vectors can
be generated in the authoring container. Never used at run time on the GPU box (the checkout does not exist there).

Two shims are needed: (1) `timm` is not installed and the reference only needs
`timm.layers.{DropPath, trunc_normal_}`; (2) the reference's `datasets/` namespace package is shadowed by the
HuggingFace `datasets` distribution, so `datasets/mae_loader.py` is loaded by path.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("VMAE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models"))


def _install_timm_shim():
    if "timm" in sys.modules:
        return

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                m.div_(keep)
            return x * m

    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")
    layers.DropPath = DropPath
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.layers = layers
    sys.modules["timm"] = timm
    sys.modules["timm.layers"] = layers


def _load_by_path(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """-> namespace with the reference modules of the hot path."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    _install_timm_shim()
    ns = types.SimpleNamespace()
    ns.tiny_vit = _load_by_path("_ref_tiny_vit", "src/models/tiny_vit.py")
    ns.mae_vit_adapter = _load_by_path("_ref_mae_vit_adapter", "src/models/mae_vit_adapter.py")
    ns.mae_loader = _load_by_path("_ref_mae_loader", "src/datasets/mae_loader.py")
    ns.masking = _load_by_path("_ref_masking", "src/mae/masking.py")
    ns.losses = _load_by_path("_ref_losses", "src/mae/losses.py")
    ns.feature_noise = _load_by_path("_ref_feature_noise", "src/privacy/feature_noise.py")
    # fed_loop imports `federated.comm_cost` at module import: provide it by path first
    fed_pkg = types.ModuleType("federated")
    fed_pkg.__path__ = [os.path.join(REF_ROOT, "src", "federated")]
    sys.modules.setdefault("federated", fed_pkg)
    _load_by_path("federated.comm_cost", "src/federated/comm_cost.py")
    ns.fed_loop = _load_by_path("_ref_fed_loop", "src/federated/fed_loop.py")
    return ns


def patchify_reference(imgs, p=8):
    """train_ssl_mae.py:26-31 executed verbatim from the reference source (the module itself drags in tensorboard,
    torchvision and the shadowed `datasets` import, so only this function is extracted)."""
    src = open(os.path.join(REF_ROOT, "src", "train_ssl_mae.py"), encoding="utf-8").read()
    start = src.index("def patchify(")
    end = src.index("def setup_logger(")
    scope = {"torch": torch}
    exec(src[start:end], scope)  # noqa: S102 - executes the reference's own function, unmodified
    return scope["patchify"](imgs, p)
