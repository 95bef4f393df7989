"""Developer check (GPU): LayerNorm fwd/bwd throughput at the decoder shape (rows = 64*1568, cols = 384)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
rows, cols = 64 * 1568, 384
x = torch.randn(rows, cols, device=dev)
dy = torch.randn(rows, cols, device=dev).bfloat16()
res = torch.randn(rows, cols, device=dev)
gm = torch.randn(cols, device=dev)
bt = torch.randn(cols, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


y, mean, rstd = ops.layernorm_fwd(x, gm, bt)
t = timeit(lambda: ops.layernorm_fwd(x, gm, bt))
b = rows * cols * (4 + 2)
print(f"ln fwd  f32->bf16: {t:.1f} us  {b / t / 1e3:.0f} GB/s")
t = timeit(lambda: ops.layernorm_bwd(dy, x, gm, mean, rstd, dx_resid=res, want_bf16=True))
b = rows * cols * (2 + 4 + 4 + 4 + 2)
print(f"ln bwd  (dy bf16, x f32, resid f32 -> dx f32 + bf16): {t:.1f} us  {b / t / 1e3:.0f} GB/s")
t = timeit(lambda: ops.layernorm_bwd(dy, x, gm, mean, rstd, dx_resid=None, want_bf16=False))
b = rows * cols * (2 + 4 + 4)
print(f"ln bwd  (dy bf16, x f32 -> dx f32): {t:.1f} us  {b / t / 1e3:.0f} GB/s")
rows2 = 64 * 160
x2 = torch.randn(rows2, cols, device=dev)
dy2 = torch.randn(rows2, cols, device=dev).bfloat16()
y2, m2, r2 = ops.layernorm_fwd(x2, gm, bt)
print(f"encoder size ({rows2} rows): fwd {timeit(lambda: ops.layernorm_fwd(x2, gm, bt)):.1f} us, "
      f"bwd {timeit(lambda: ops.layernorm_bwd(dy2, x2, gm, m2, r2, dx_resid=x2, want_bf16=True)):.1f} us (includes two torch.zeros fills)")
