"""CPU: `datasets.mae_loader.LazyVideoMAEDataset` — same clips as the live reference class on the same folders and RNG
state (when the reference checkout is present), the reference's error handling, and the uint8 hand-off mode that feeds
`ops.frames_to_clip`."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ssl_vit_video_analytics_b200.datasets.mae_loader import LazyVideoMAEDataset  # noqa: E402

MEAN = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
STD = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)


def _to_float_normalised(img):
    """PILToTensor + ConvertImageDtype(float) + Normalize of train_ssl_mae.py:137-141, spelled out (no torchvision needed)."""
    t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
    return (t - MEAN) / STD


def _make_videos(tmp_path, lengths, size=16):
    from PIL import Image

    rng = np.random.RandomState(0)
    dirs = []
    for v, n in enumerate(lengths):
        d = tmp_path / f"video_{v:02d}"
        d.mkdir()
        for i in range(n):
            Image.fromarray(rng.randint(0, 256, size=(size, size, 3), dtype=np.uint8)).save(d / f"frame_{i:04d}.jpg", quality=95)
        (d / "notes.txt").write_text("not a frame")
        dirs.append(str(d))
    split = tmp_path / "train.txt"
    split.write_text("".join(f"{d} {v}\n" for v, d in enumerate(dirs)) + "\n")   # label column and a blank line, as real split files have
    return str(split), dirs


def test_split_file_and_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        LazyVideoMAEDataset(str(tmp_path / "missing.txt"))
    split, dirs = _make_videos(tmp_path, [5, 0])
    ds = LazyVideoMAEDataset(split, clip_len=4, stride=2, image_size=16, transform=_to_float_normalised)
    assert len(ds) == 2 and ds.samples == dirs
    with pytest.raises(ValueError):
        LazyVideoMAEDataset(split, transform=_to_float_normalised, uint8_frames=True)
    # empty directory and a directory that vanished: all-zero clip of the configured geometry
    assert torch.equal(ds[1], torch.zeros(3, 4, 16, 16))
    ds.samples.append(str(tmp_path / "gone"))
    assert torch.equal(ds[2], torch.zeros(3, 4, 16, 16))
    # unreadable frame -> black image -> the normalised value of 0 in every channel
    bad = os.path.join(dirs[0], "frame_0000.jpg")
    with open(bad, "wb") as f:
        f.write(b"not a jpeg")
    clip = ds[0]   # 5 frames < 4 * 2: linspace picks frames 0, 1, 2, 4
    assert clip.shape == (3, 4, 16, 16)
    black = ((0.0 - MEAN) / STD).view(3)[[2, 1, 0]]
    assert torch.allclose(clip[:, 0], black.view(3, 1, 1).expand(3, 16, 16))


def test_temporal_window_and_channel_order(tmp_path):
    split, dirs = _make_videos(tmp_path, [7, 40])
    ds = LazyVideoMAEDataset(split, clip_len=8, stride=2, image_size=16, transform=_to_float_normalised)
    from PIL import Image

    names = sorted(n for n in os.listdir(dirs[0]) if n.endswith(".jpg"))
    short = ds[0]   # 7 frames < 16: np.linspace(0, 6, 8).astype(int) = 0 0 1 2 3 4 5 6
    want = [0, 0, 1, 2, 3, 4, 5, 6]
    for t, i in enumerate(want):
        ref = _to_float_normalised(Image.open(os.path.join(dirs[0], names[i])).convert("RGB"))[[2, 1, 0]]
        assert torch.equal(short[:, t], ref)
    np.random.seed(3)
    start = np.random.randint(0, 40 - 16 + 1)
    np.random.seed(3)
    long = ds[1]
    names = sorted(n for n in os.listdir(dirs[1]) if n.endswith(".jpg"))
    for t in range(8):
        ref = _to_float_normalised(Image.open(os.path.join(dirs[1], names[start + 2 * t])).convert("RGB"))[[2, 1, 0]]
        assert torch.equal(long[:, t], ref)


def test_uint8_mode_is_the_same_clip_before_normalisation(tmp_path):
    split, _ = _make_videos(tmp_path, [12])
    f32 = LazyVideoMAEDataset(split, clip_len=4, stride=2, image_size=16, transform=_to_float_normalised)
    u8 = LazyVideoMAEDataset(split, clip_len=4, stride=2, image_size=16, uint8_frames=True)
    np.random.seed(11)
    a = f32[0]
    np.random.seed(11)
    frames = u8[0]
    assert frames.dtype == torch.uint8 and frames.shape == (4, 16, 16, 3)
    # what ops.frames_to_clip computes on the GPU (tests/test_kernels_gpu.py::test_frames_to_clip_matches_loader_transform), on the host
    x = frames.permute(3, 0, 1, 2).float() / 255.0                       # [C, T, H, W], RGB
    x = (x - MEAN.view(3, 1, 1, 1)) / STD.view(3, 1, 1, 1)
    assert torch.equal(x[[2, 1, 0]], a)
    assert torch.equal(LazyVideoMAEDataset(split, clip_len=4, image_size=16, uint8_frames=True)._empty(), torch.zeros(4, 16, 16, 3, dtype=torch.uint8))


def test_same_clips_as_the_live_reference_class(tmp_path):
    from oracle import reference_shim

    if not reference_shim.available():
        pytest.skip("reference checkout not present (GPU box)")
    ref_cls = reference_shim.load().mae_loader.LazyVideoMAEDataset
    split, _ = _make_videos(tmp_path, [6, 30, 0, 17])
    kw = dict(clip_len=8, stride=2, image_size=16, transform=_to_float_normalised)
    ours, ref = LazyVideoMAEDataset(split, **kw), ref_cls(split, **kw)
    assert len(ours) == len(ref) and ours.samples == ref.samples
    for i in range(len(ref)):
        np.random.seed(100 + i)
        want = ref[i]
        np.random.seed(100 + i)
        got = ours[i]
        assert got.dtype == want.dtype and got.shape == want.shape and torch.equal(got, want), i
    loader = torch.utils.data.DataLoader(ours, batch_size=2, shuffle=False, num_workers=0)
    assert next(iter(loader)).shape == (2, 3, 8, 16, 16)
