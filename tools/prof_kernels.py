"""Launch a handful of representative kernels once each (for ncu). usage: prof_kernels.py [gemm|attn|all]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device="cuda").manual_seed(0)
M = 100352
if what in ("gemm", "all"):
    x = torch.randn(M, 384, device=dev, generator=g).bfloat16()
    w1 = (torch.randn(1536, 384, device=dev, generator=g) * 0.05).bfloat16()
    wp = (torch.randn(384, 384, device=dev, generator=g) * 0.05).bfloat16()
    b1 = torch.randn(1536, device=dev, generator=g)
    bp = torch.randn(384, device=dev, generator=g)
    res = torch.randn(M, 384, device=dev, generator=g)
    pre = torch.empty(M, 1536, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(x, w1, M, 1536, 384)                                              # plain, N=1536
        ops.gemm(x, w1, M, 1536, 384, bias=b1, act=L.ACT_GELU, aux_out=pre)        # fc1
        ops.gemm(x, wp, M, 384, 384, bias=bp, residual=res, out_dtype=torch.float32)  # proj + residual
        ops.linear_dgrad(x, w1.t().contiguous(), aux_in=pre)                       # fc2 dgrad * gelu'
        ops.linear_wgrad(pre, x)                                                   # wgrad 1536x384
if what in ("attn", "all"):
    B, Lq, H, hd = 64, 1568, 6, 64
    qkv = torch.randn(B, Lq, 3 * H * hd, device=dev, generator=g).bfloat16()
    dout = torch.randn(B * Lq, H * hd, device=dev, generator=g).bfloat16()
    for _ in range(2):
        out, lse = ops.attn_fwd(qkv, B, Lq, H, hd)
        ops.attn_bwd(qkv, out, dout, lse, B, Lq, H, hd)
torch.cuda.synchronize()
print("done")
