"""Developer check (GPU): depthwise 3x3 throughput at the TinyViT shapes (batch 40 clips x 16 frames = 640 images)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda:0")


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


for (N, H, W, C, s) in [(640, 56, 56, 384, 1), (640, 56, 56, 384, 2), (640, 28, 28, 768, 2)]:
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = torch.randn(C, 1, 3, 3, device=dev) * 0.3
    lib = L.load()
    p = lambda t_: t_.data_ptr() if t_ is not None else None
    st = torch.cuda.current_stream().cuda_stream
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    y = torch.empty(N, Ho, Wo, C, device=dev, dtype=torch.bfloat16)
    w9 = w.reshape(C, 9).contiguous()
    dy = torch.randn(N, Ho, Wo, C, device=dev).bfloat16()
    dx = torch.empty_like(x)
    dw = torch.zeros(C, 9, device=dev)
    bi, bo = x.numel() * 2, y.numel() * 2
    t = timeit(lambda: L.check(lib.vmae_dwconv3x3_fwd(p(x), p(w9), p(y), N, H, W, C, s, st), "f"))
    print(f"[{N},{H},{W},{C}] s{s} fwd   {t:8.1f} us {(bi + bo) / t / 1e3:6.0f} GB/s")
    t = timeit(lambda: L.check(lib.vmae_dwconv3x3_bwd(p(x), p(dy), p(w9), p(dx), None, N, H, W, C, s, st), "d"))
    print(f"[{N},{H},{W},{C}] s{s} dgrad {t:8.1f} us {(bi + bo) / t / 1e3:6.0f} GB/s")
    t = timeit(lambda: L.check(lib.vmae_dwconv3x3_bwd(p(x), p(dy), p(w9), None, p(dw), N, H, W, C, s, st), "w"))
    print(f"[{N},{H},{W},{C}] s{s} wgrad {t:8.1f} us {(bi + bo) / t / 1e3:6.0f} GB/s")
    del x, y, dy, dx
