"""Drop-in for `src/datasets/mae_loader.py::get_tube_mask` (:80-90). Host-side, bit-exact with the reference for the
same torch CPU RNG state: one `torch.rand(L)` draw per sample, the `int(ratio*L)` largest draws are masked, and the
spatial mask is repeated over all frames. (`LazyVideoMAEDataset` — JPEG folder I/O — is out of scope, SURVEY.md §2 #4.)
"""
from __future__ import annotations

import torch


def get_tube_mask(batch_size, num_frames, num_patches, mask_ratio):
    num_mask = int(mask_ratio * num_patches)
    frame_masks = torch.zeros(batch_size, num_patches, dtype=torch.bool)
    for b in range(batch_size):
        # per-sample draw: keeps the generator stream identical to the reference loop
        order = torch.argsort(torch.rand(num_patches), descending=True)
        frame_masks[b, order[:num_mask]] = True
    return frame_masks[:, None, :].expand(batch_size, num_frames, num_patches).contiguous()


def visible_indices(mask):
    """Ascending ids of the visible patches per clip for a tube mask [B, T, L] -> int32 [B, n_vis] (host tensor).
    This is the gather index list of the north-star model; the reference itself never gathers."""
    frame = mask[:, 0]
    n_vis = int((~frame[0]).sum())
    rows = []
    for b in range(frame.shape[0]):
        ids = torch.nonzero(~frame[b]).flatten()
        if ids.numel() != n_vis:
            raise ValueError("visible_indices: clips keep different numbers of patches")
        rows.append(ids)
    return torch.stack(rows).to(torch.int32)


def sample_frame_indices(total_frames: int, clip_len: int, stride: int, rng=None):
    """Temporal window of `LazyVideoMAEDataset.__getitem__` (:44-59): `clip_len` frame indices. Videos shorter than
    clip_len*stride frames are sampled uniformly (linspace, truncated to int); otherwise a random start (numpy's global RNG, as in
    the reference, or `rng`) and a fixed stride. Always exactly clip_len indices. Host-side; the decoded uint8 frames then go to the
    GPU as [B, T, H, W, 3] and `ops.frames_to_clip` does ToTensor + Normalize + the channel reversal + the [C, T, H, W] permute."""
    import numpy as np

    window = clip_len * stride
    if total_frames < window:
        idx = np.linspace(0, total_frames - 1, clip_len).astype(int)
    else:
        start = (rng or np.random).randint(0, total_frames - window + 1)
        idx = np.arange(start, start + window, stride)
    return idx[:clip_len]
