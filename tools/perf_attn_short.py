"""Developer probe (GPU): the encoder-shape attention of the benchmark (B=64, H=6, L=160 visible tokens, head_dim 64): forward, both
backward forms and torch SDPA (cuDNN) side by side. usage: python tools/perf_attn_short.py [L]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import ops  # noqa: E402

B, H, hd = 64, 6, 64
L = int(sys.argv[1]) if len(sys.argv) > 1 else 160
qkv = torch.randn(B, L, 3, H, hd, device="cuda").bfloat16()
flat = qkv.view(B * L, -1)
o, lse = ops.attn_fwd(flat, B, L, H, hd)
do = torch.randn_like(o)


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print(f"L={L}: fwd {t(lambda: ops.attn_fwd(flat, B, L, H, hd)):.1f} us | bwd single-pass {t(lambda: ops.attn_bwd(flat, o, do, lse, B, L, H, hd, fused=True)):.1f} us"
      f" | bwd two-kernel {t(lambda: ops.attn_bwd(flat, o, do, lse, B, L, H, hd, fused=False)):.1f} us")
q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
qr, kr, vr = [x.detach().clone().requires_grad_(True) for x in (q, k, v)]
oo = torch.nn.functional.scaled_dot_product_attention(qr, kr, vr)
dd = torch.randn_like(oo)
print(f"torch SDPA: fwd {t(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)):.1f} us | bwd {t(lambda: torch.autograd.grad(oo, (qr, kr, vr), dd, retain_graph=True)):.1f} us")
