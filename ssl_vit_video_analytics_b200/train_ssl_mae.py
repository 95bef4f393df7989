"""Training-step driver mirroring `src/train_ssl_mae.py` (`patchify` :26-31, `train_one_epoch` :52-123).

The step is the reference's: tube mask (host RNG) -> normalised-pixel target -> bf16 forward -> masked MSE -> backward
-> optimizer step. The target construction and the loss are ONE fused kernel (`ops.masked_mse`) that reads the clip
once instead of materialising `patchify(clip)` and three normalisation passes.
"""
from __future__ import annotations

import time

import torch

from . import ops
from .datasets.mae_loader import get_tube_mask


def patchify(imgs, p=8):
    """[B,C,T,H,W] -> [B, T*(H/p)*(W/p), p*p*C], element order (py, px, c) — layout contract of the prediction head."""
    B, Cc, T, H, W = imgs.shape
    hp, wp = H // p, W // p
    x = imgs.reshape(B, Cc, T, hp, p, wp, p).permute(0, 2, 3, 5, 4, 6, 1)
    return x.reshape(B, T * hp * wp, p * p * Cc)


def unpatchify(tokens, T, H, W, p=8, C=3):
    """Inverse of `patchify` (visualize_mae.py:19-31)."""
    B = tokens.shape[0]
    hp, wp = H // p, W // p
    x = tokens.reshape(B, T, hp, wp, p, p, C).permute(0, 6, 1, 2, 4, 3, 5)
    return x.reshape(B, C, T, H, W)


def mae_step_loss(model, clip, mask, norm_pix_loss=True, patch=8):
    """pred = model(clip, mask) under bf16 autocast, then the fused masked MSE (train_ssl_mae.py:72-84)."""
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        pred = model(clip, mask)
    return ops.masked_mse(pred, clip, mask, 1, patch, 1 if norm_pix_loss else 0), pred


def train_one_epoch(model, loader, optimizer, scaler, epoch, device, config, writer=None, logger=None):
    model.train()
    ssl_cfg, train_cfg = config["ssl"], config["training"]
    num_steps = len(loader)
    batch_size = train_cfg["batch_size"]
    t_epoch = t_log = time.time()
    total_loss = 0.0
    for step, clip in enumerate(loader):
        clip = clip.to(device, non_blocking=True)
        if clip.dtype == torch.uint8:   # LazyVideoMAEDataset(uint8_frames=True): [B, T, H, W, 3] frames, normalised on the GPU
            clip = ops.frames_to_clip(clip)
        B, Cc, T, H, W = clip.shape
        Lp = (H // 8) * (W // 8)
        mask = get_tube_mask(B, T, Lp, ssl_cfg["mask_ratio"]).to(device)
        loss, pred = mae_step_loss(model, clip, mask, ssl_cfg["norm_pix_loss"])
        optimizer.zero_grad()
        if scaler is not None:
            scaler.scale(loss).backward()
            scaler.step(optimizer)
            scaler.update()
        else:
            loss.backward()
            optimizer.step()
        total_loss += loss.item()
        if step % train_cfg["log_interval"] == 0 and step > 0:
            now = time.time()
            throughput = train_cfg["log_interval"] * batch_size / (now - t_log)
            t_log = now
            pred_std = pred[mask.flatten(1, 2)].float().std().item()
            if logger is not None:
                logger.info(f"Epoch [{epoch}] [{step:4d}/{num_steps}] Loss: {loss.item():.4f} | Std: {pred_std:.3f} | "
                            f"Speed: {throughput:.1f} samples/s")
            if writer is not None:
                gstep = (epoch - 1) * num_steps + step
                writer.add_scalar("Train/Loss", loss.item(), gstep)
                writer.add_scalar("Train/Throughput", throughput, gstep)
    return total_loss / max(1, num_steps), time.time() - t_epoch
