"""Tiny instance of every kernel family behind the C ABI, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py`
(one tool per gpurun call). Shapes are the smallest that still exercise tails (ragged rows / partial tiles) so that the tools finish."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import _lib as L, ops  # noqa: E402
from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask, visible_indices  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 200, 192, 128
x = torch.randn(M, K, device=dev, generator=g).bfloat16()
w = (torch.randn(N, K, device=dev, generator=g) * 0.05).bfloat16()
b = torch.randn(N, device=dev, generator=g)
res = torch.randn(M, N, device=dev, generator=g)
pre16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
pre8 = torch.empty(M, N, device=dev, dtype=torch.uint8)
ops.gemm(x, w, M, N, K, bias=b)
ops.gemm(x, w, M, N, K, bias=b, act=L.ACT_GELU, aux_out=pre16)
ops.gemm(x, w, M, N, K, bias=b, act=L.ACT_GELU, aux_out=pre8)
ops.gemm(x, w, M, N, K, bias=b, residual=res, out_dtype=torch.float32)
ops.gemm(x, w, M, N, K, aux_in=pre16)
ops.gemm(x, w, M, N, K, aux_in=pre8)
dy = torch.randn(M, N, device=dev, generator=g).bfloat16()
ops.linear_dgrad(dy, w)
ops.linear_wgrad_bias(dy, x)
ops.LinearNoiseFn.apply(x, w.float(), b, 0.1, 1, 0, torch.empty(M, N, device=dev))
print("gemm ok", flush=True)
xf = torch.randn(M, N, device=dev, generator=g)
gm, bt = torch.ones(N, device=dev), torch.zeros(N, device=dev)
y, mean, rstd = ops.layernorm_fwd(xf, gm, bt)
ops.layernorm_bwd(dy, xf, gm, mean, rstd, dx_resid=xf, want_bf16=True)
print("layernorm ok", flush=True)
for (B, Lq, H, hd) in [(1, 160, 2, 64), (1, 100, 1, 32)]:
    qkv = torch.randn(B, Lq, 3 * H * hd, device=dev, generator=g).bfloat16()
    do = torch.randn(B * Lq, H * hd, device=dev, generator=g).bfloat16()
    out, lse = ops.attn_fwd(qkv, B, Lq, H, hd)
    ops.attn_bwd(qkv, out, do, lse, B, Lq, H, hd, fused=True)     # single-pass v2 kernel (head_dim 64)
    ops.attn_bwd(qkv, out, do, lse, B, Lq, H, hd, fused=False)    # two-kernel form
print("attention ok", flush=True)
clip = torch.randn(2, 3, 4, 64, 64, device=dev, generator=g)
mask = get_tube_mask(2, 2, 16, 0.75)
vis = visible_indices(mask).to(dev)
we = (torch.randn(384, 3, 2, 16, 16, device=dev, generator=g) * 0.03).requires_grad_(True)
be = torch.zeros(384, device=dev, requires_grad=True)
tp, sp = torch.zeros(1, 2, 1, 384, device=dev, requires_grad=True), torch.zeros(1, 1, 16, 384, device=dev, requires_grad=True)
emb = ops.TubeletEmbedFn.apply(clip, vis, we, be, tp, sp, 2, 16)
emb.backward(torch.ones_like(emb))
pred = torch.randn(2, 32, 1536, device=dev, generator=g).bfloat16().requires_grad_(True)
ops.masked_mse(pred, clip, mask.to(dev), 2, 16, 1).backward()
print("tubelet embed + loss ok", flush=True)
tok = torch.randn(2, 18, 64, device=dev, generator=g).requires_grad_(True)
m8 = (torch.rand(2, 18, device=dev, generator=g) < 0.5).to(torch.uint8)
m8[:, :9] = 1
m8[:, 9:] = 0
mt = torch.zeros(1, 1, 64, device=dev, requires_grad=True)
o = ops.MaskTokenBlendFn.apply(tok, m8, mt)
idx = ops.mask_row_indices(m8, 9, 1)
ops.GatherRowsFn.apply(o, idx).sum().backward()
ops.softmax_cross_entropy(torch.randn(5, 101, device=dev, generator=g).requires_grad_(True), torch.randint(0, 101, (5,), device=dev, generator=g)).backward()
ops.MeanPoolNoiseFn.apply(torch.randn(2, 33, 128, device=dev, generator=g).bfloat16(), 0.1, 1, 0, None)
ops.GaussianNoiseFn.apply(torch.randn(1000, device=dev, generator=g), 0.1, 1, 0, None)
ops.AdaptiveAvgPoolFn.apply(torch.randn(2, 7, 7, 64, device=dev, generator=g).bfloat16(), 3)
print("stack-2 / head kernels ok", flush=True)
from ssl_vit_video_analytics_b200.optim import FusedAdamW  # noqa: E402
p = torch.nn.Parameter(torch.randn(5000, device=dev, generator=g))
p.grad = torch.randn_like(p)
FusedAdamW([p], lr=1e-3).step()
torch.cuda.synchronize()
print("sanitize smoke: all kernels ran")
