"""Developer check (GPU): train-mode BatchNorm forward / backward throughput at the TinyViT stage-0 shape."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


p = lambda t: t.data_ptr() if t is not None else None
st = torch.cuda.current_stream().cuda_stream
for (M, C) in [(640 * 56 * 56, 384), (640 * 56 * 56, 96), (640 * 14 * 14, 768)]:
    x = torch.randn(M, C, device=dev).bfloat16()
    dy = torch.randn(M, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    g, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    mean, rstd = torch.empty(C, device=dev), torch.empty(C, device=dev)
    ws = torch.empty(2 * C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    nb = M * C * 2
    for act in (0, 1):
        t = timeit(lambda: L.check(lib.vmae_bn_train_fwd(p(x), p(g), p(b), None, p(y), p(mean), p(rstd), None, None, p(ws), M, C, 1e-5, 0.1, act, st), "f"))
        print(f"bn fwd  M={M} C={C} act={act}: {t:8.1f} us {3 * nb / t / 1e3:6.0f} GB/s")
        t = timeit(lambda: L.check(lib.vmae_bn_train_bwd(p(x), p(dy), p(g), p(b), p(mean), p(rstd), p(dx), p(dg), p(db), p(ws), M, C, act, st), "b"))
        print(f"bn bwd  M={M} C={C} act={act}: {t:8.1f} us {5 * nb / t / 1e3:6.0f} GB/s")
    del x, dy, y, dx
