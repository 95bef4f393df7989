"""Drop-in for `src/datasets/mae_loader.py::get_tube_mask` (:80-90). Host-side, bit-exact with the reference for the
same torch CPU RNG state: one `torch.rand(L)` draw per sample, the `int(ratio*L)` largest draws are masked, and the
spatial mask is repeated over all frames.

`LazyVideoMAEDataset` (:7-78) is mirrored as a host-side drop-in: same split-file format, constructor, temporal
window, error handling and output as the reference when a per-frame `transform` is given, plus `uint8_frames=True`, which returns
the decoded frames untouched as uint8 [T, H, W, 3] so that the batch crosses PCIe at a quarter of the bytes and
`ops.frames_to_clip` does PILToTensor + ConvertImageDtype + Normalize + channel reversal + permute in one kernel on the GPU.
JPEG decoding itself stays PIL on the host, as in the reference.
"""
from __future__ import annotations

import os

import torch
from torch.utils.data import Dataset


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


class LazyVideoMAEDataset(Dataset):
    """Folder-of-JPEG-frames clips listed in a split file (first whitespace-separated field of every line = frame directory),
    `mae_loader.py:7-78`. `__getitem__` returns what the reference returns — float [C, T, H, W] with the channel order reversed after
    a tensor-producing `transform`, an all-zero clip for a missing or empty directory, a black frame for an unreadable file — or, with
    `uint8_frames=True`, the raw frames as uint8 [T, H, W, 3] (RGB as decoded) for GPU-side normalisation by `ops.frames_to_clip`."""

    def __init__(self, split_file, clip_len=32, stride=2, image_size=112, transform=None, uint8_frames=False):
        if not os.path.exists(split_file):
            raise FileNotFoundError(f"Index file not found: {split_file}")
        if uint8_frames and transform is not None:
            raise ValueError("LazyVideoMAEDataset: uint8_frames=True hands the frames to the GPU untransformed; pass transform=None")
        self.clip_len, self.stride, self.image_size = clip_len, stride, image_size
        self.transform, self.uint8_frames = transform, uint8_frames
        with open(split_file, "r") as f:
            fields = (line.split() for line in f)
            self.samples = [fl[0] for fl in fields if fl]

    def __len__(self):
        return len(self.samples)

    def _load_frame(self, frame_path):
        from PIL import Image

        try:
            return Image.open(frame_path).convert("RGB")
        except Exception:   # unreadable / truncated file: a black frame, as the reference does
            return Image.new("RGB", (self.image_size, self.image_size), (0, 0, 0))

    def _empty(self):
        if self.uint8_frames:
            return torch.zeros(self.clip_len, self.image_size, self.image_size, 3, dtype=torch.uint8)
        return torch.zeros(3, self.clip_len, self.image_size, self.image_size)

    def __getitem__(self, index):
        import numpy as np

        video_dir = self.samples[index]
        if not os.path.exists(video_dir):
            return self._empty()
        names = sorted(n for n in os.listdir(video_dir) if n.endswith(".jpg"))
        if not names:
            return self._empty()
        frames = []
        for i in sample_frame_indices(len(names), self.clip_len, self.stride):
            img = self._load_frame(os.path.join(video_dir, names[i]))
            if self.uint8_frames:
                frames.append(torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()))   # [H, W, 3]
                continue
            if self.transform:
                img = self.transform(img)
            if isinstance(img, torch.Tensor):
                img = img[[2, 1, 0], :, :]   # the reference reverses the channel order of every tensor frame (:71-73)
            frames.append(img)
        if self.uint8_frames:
            return torch.stack(frames)                        # [T, H, W, 3]
        return torch.stack(frames).permute(1, 0, 2, 3)        # [T, C, H, W] -> [C, T, H, W]
