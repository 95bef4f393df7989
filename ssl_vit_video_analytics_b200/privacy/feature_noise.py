"""Drop-in for `src/privacy/feature_noise.py::add_gaussian_noise` (:4-7): z' = z + N(0, sigma^2 I); returns `z` itself
when sigma <= 0. The normals come from the library's Philox4x32-10 stream (statistical, not bit, parity with torch)."""
from __future__ import annotations

import torch

from .. import ops

_offset = 0


def add_gaussian_noise(z, sigma, seed=None, eps_out=None):
    """`eps_out` (optional, fp32, z's shape, CUDA) receives the normals that were drawn, so that `out == z + sigma * eps_out`
    can be checked exactly (the parity contract of SURVEY.md §8 a19)."""
    global _offset
    if sigma <= 0:
        return z
    if seed is None:
        seed = torch.initial_seed()
    out = ops.GaussianNoiseFn.apply(z, float(sigma), int(seed) & 0xFFFFFFFFFFFFFFFF, _offset, eps_out)
    _offset += (z.numel() + 3) // 4
    return out


def apply_feature_mask(z, mask_ratio):
    if mask_ratio <= 0:
        return z
    keep = torch.bernoulli(torch.full_like(z, 1.0 - mask_ratio))
    return z * keep
