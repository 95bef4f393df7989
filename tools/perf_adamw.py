"""AdamW step time on the ViT-S VideoMAE parameter set: optim.FusedAdamW vs torch.optim.AdamW(fused=True). python tools/perf_adamw.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200.models import VideoMAEViT  # noqa: E402
from ssl_vit_video_analytics_b200.models.videomae_vit import DEFAULT_CONFIG  # noqa: E402
from ssl_vit_video_analytics_b200.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda:0")
model = VideoMAEViT(DEFAULT_CONFIG).to(dev)
ps = list(model.parameters())
for p in ps:
    p.grad = torch.randn_like(p) * 0.01
n = sum(p.numel() for p in ps)
for name, opt in (("torch fused", torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.05, fused=True)), ("FusedAdamW", FusedAdamW(ps, lr=1e-4, weight_decay=0.05))):
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:12s}: {ms * 1e3:7.1f} us/step  ({28.0 * n / ms / 1e6:.0f} GB/s for {n / 1e6:.1f} M parameters, {len(ps)} tensors)")
