"""How long does the host need to ENQUEUE one training step (no synchronisation) vs. how long the GPU needs to run it?
python tools/cpu_issue_time.py   (one B200)"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from ssl_vit_video_analytics_b200.models import VideoMAEViT  # noqa: E402
from ssl_vit_video_analytics_b200.parallel import GradientBucketer  # noqa: E402

dev = torch.device("cuda:0")
from ssl_vit_video_analytics_b200.models.videomae_vit import DEFAULT_CONFIG as cfg  # noqa: E402
B = bench.PER_GPU_BATCH
torch.manual_seed(0)
model = VideoMAEViT(cfg).to(dev).train()
ddp = GradientBucketer(model, bucket_mb=32.0, exclude=())
clip_h, mask_h, vis_h = bench.make_inputs(B, cfg, seed=0)
clip, mask_d, vis_d = clip_h.to(dev), mask_h.to(dev), vis_h.to(dev)
opt = torch.optim.AdamW(list(model.parameters()), lr=5e-4, weight_decay=0.05, fused=True)


def step():
    ddp.zero_grad()
    loss = model.training_loss(clip, mask_d, vis_d)
    loss.backward()
    ddp.finish_gradient_sync()
    opt.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
for n in (1, 2):
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{n} step(s): host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, until the GPU is done {1e3 * (t2 - t0) / n:.2f} ms/step")
