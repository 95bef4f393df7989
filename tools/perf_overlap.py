"""Developer probe (GPU): how much of a split-K wgrad GEMM hides under a bandwidth-bound LayerNorm backward when the two run on
different streams? Prints sequential and concurrent times for the decoder shapes (M = 64 x 1568 rows, D = 384)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M, D, Hd = 100352, 384, 1536
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, device=dev, generator=g)
dy = torch.randn(M, D, device=dev, generator=g).bfloat16()
gamma = torch.ones(D, device=dev)
beta = torch.zeros(D, device=dev)
_, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
dres = torch.randn(M, D, device=dev, generator=g)
dh = torch.randn(M, Hd, device=dev, generator=g).bfloat16()
ln2 = torch.randn(M, D, device=dev, generator=g).bfloat16()
dw = torch.zeros(Hd, D, device=dev)
db = torch.zeros(Hd, device=dev)
dgam = torch.zeros(D, device=dev)
dbet = torch.zeros(D, device=dev)
side = torch.cuda.Stream(priority=0)
main = torch.cuda.current_stream()


def ln():
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx_resid=dres, dgamma=dgam, dbeta=dbet, want_bf16=True)


def wg():
    ops.linear_wgrad_bias(dh, ln2, dw, db)


def dg():
    ops.linear_dgrad(dh, w1)


w1 = (torch.randn(Hd, D, device=dev, generator=g) * 0.05).bfloat16()


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


def both(a, b):
    def f():
        side.wait_stream(main)
        with torch.cuda.stream(side):
            b()
        a()
        main.wait_stream(side)
    return f


t_ln, t_wg, t_dg = timeit(ln), timeit(wg), timeit(dg)
print(f"alone: layernorm_bwd {t_ln:.0f} us | wgrad fc1 {t_wg:.0f} us | dgrad fc1 {t_dg:.0f} us")
print(f"sequential ln+wgrad {timeit(lambda: (ln(), wg())):.0f} us | concurrent {timeit(both(ln, wg)):.0f} us")
print(f"sequential dgrad+wgrad {timeit(lambda: (dg(), wg())):.0f} us | concurrent {timeit(both(dg, wg)):.0f} us")
print(f"sequential dgrad+ln+wgrad {timeit(lambda: (dg(), ln(), wg())):.0f} us | concurrent (wgrad aside) {timeit(both(lambda: (dg(), ln()), wg)):.0f} us")
