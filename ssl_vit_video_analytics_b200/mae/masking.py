"""Drop-in for `src/mae/masking.py` (:9-65): mask-ratio schedule, random / tube token masks, counters.
RNG draws (`torch.randperm(..., device=device)`, one per sample) follow the reference order, so masks are bit-exact
for the same generator state on the same device."""
from __future__ import annotations

from typing import Any, Dict, List

import torch


def get_mask_ratio(epoch: int, schedule: List[Dict[str, Any]], default: float) -> float:
    for seg in schedule or []:
        lo, hi = int(seg.get("start", 1)), int(seg.get("end", 10 ** 9))
        if lo <= epoch < hi:
            return float(seg.get("value", default))
    return float(default)


def make_token_mask(B: int, T: int, tokens_per_frame: int, mask_ratio: float, mode: str, device: torch.device) -> torch.Tensor:
    """-> bool [B, T*tokens_per_frame], True = masked."""
    P = tokens_per_frame
    N = T * P
    if mode == "random":
        k = max(1, int(round(N * mask_ratio)))
        mask = torch.zeros((B, N), dtype=torch.bool, device=device)
        for b in range(B):
            mask[b, torch.randperm(N, device=device)[:k]] = True
        return mask
    if mode == "tube":
        k = max(1, int(round(P * mask_ratio)))
        frame = torch.zeros((B, P), dtype=torch.bool, device=device)
        for b in range(B):
            frame[b, torch.randperm(P, device=device)[:k]] = True
        return frame[:, None, :].expand(B, T, P).reshape(B, N).contiguous()
    raise ValueError(f"Unknown mask_mode: {mode}")


def count_masked(mask: torch.Tensor) -> int:
    return int(mask.sum().item())


def count_visible(mask: torch.Tensor) -> int:
    return int((~mask).sum().item())
