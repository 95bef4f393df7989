"""Developer check (GPU): tubelet embedding, implicit-GEMM kernel vs gather kernel + GEMM, forward and backward, bench shape."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import ops  # noqa: E402
from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask, visible_indices  # noqa: E402

dev = torch.device("cuda:0")
B, T, S, pt, ps, E = 64, 16, 224, 2, 16, 384
clip = torch.randn(B, 3, T, S, S, device=dev)
mask = get_tube_mask(B, T // pt, (S // ps) ** 2, 0.9)
vis = visible_indices(mask).to(dev)
w = (torch.randn(E, 3, pt, ps, ps, device=dev) * 0.03).requires_grad_(True)
b = torch.zeros(E, device=dev, requires_grad=True)
tpos = torch.zeros(1, T // pt, 1, E, device=dev, requires_grad=True)
spos = torch.zeros(1, 1, (S // ps) ** 2, E, device=dev, requires_grad=True)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def implicit():
    return ops.TubeletEmbedFn.apply(clip, vis, w, b, tpos, spos, pt, ps)


def two_kernel():
    patches, pos_rows = ops.TubeletGatherFn.apply(clip, vis, tpos, spos, pt, ps)
    return ops.LinearResidualF32Fn.apply(patches, w, b, pos_rows)


for name, fn in (("implicit", implicit), ("gather+gemm", two_kernel)):
    with torch.no_grad():
        t_f = timeit(fn)
    out = fn()
    dy = torch.randn_like(out)

    def fb():
        o = fn()
        o.backward(dy)

    t_fb = timeit(fb)
    print(f"{name}: fwd {t_f:.1f} us, fwd+bwd {t_fb:.1f} us")
