"""Developer check (GPU): LayerNorm, colsum, casts, masked MSE, gather/scatter helpers, noise vs PyTorch fp32."""
import ctypes as C
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
OK = True


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rep(name, got, ref, tol):
    global OK
    torch.cuda.synchronize()
    err = (got.float() - ref.float()).abs().max().item()
    sc = ref.float().abs().max().item() + 1e-9
    good = err / sc < tol and not torch.isnan(got.float()).any().item()
    OK &= good
    print(f"{'OK ' if good else 'BAD'} {name}: rel={err/sc:.3e}", flush=True)


g = torch.Generator(device="cuda").manual_seed(0)


def rn(*s):
    return torch.randn(*s, device=dev, generator=g)


# ---- LayerNorm fwd/bwd
for rows, cols in [(1000, 384), (333, 192), (64, 576), (50, 512), (7, 1024)]:
    x = rn(rows, cols) * 2 + 0.5
    gm, bt = rn(cols), rn(cols)
    y = torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    L.check(lib.vmae_layernorm_fwd(ptr(x), 1, ptr(gm), ptr(bt), ptr(y), 0, ptr(mean), ptr(rstd), rows, cols, 1e-5, st()), "ln")
    xr = x.clone().requires_grad_(True)
    gr, br = gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (cols,), gr, br, 1e-5)
    rep(f"ln fwd {rows}x{cols}", y, yr, 1e-2)
    dy = rn(rows, cols).bfloat16()
    res = rn(rows, cols)
    yr.backward(dy.float())
    dx = torch.empty(rows, cols, device=dev)
    dxb = torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)
    dg, db = torch.zeros(cols, device=dev), torch.zeros(cols, device=dev)
    L.check(lib.vmae_layernorm_bwd(ptr(dy), 0, ptr(x), 1, ptr(gm), ptr(mean), ptr(rstd), ptr(res), ptr(dx), ptr(dxb), ptr(dg), ptr(db), rows, cols, st()), "lnb")
    rep(f"ln bwd dx {rows}x{cols}", dx, xr.grad + res, 1e-4)
    rep(f"ln bwd dx_bf16 {rows}x{cols}", dxb, xr.grad + res, 1e-2)
    rep(f"ln bwd dgamma {rows}x{cols}", dg, gr.grad, 1e-4)
    rep(f"ln bwd dbeta {rows}x{cols}", db, br.grad, 1e-4)

# ---- colsum / cast / scale
for M, N in [(1000, 384), (5000, 1536), (77, 192)]:
    dy = rn(M, N).bfloat16()
    db = torch.ones(N, device=dev)
    L.check(lib.vmae_colsum_accum(ptr(dy), 0, N, ptr(db), M, N, st()), "colsum")
    rep(f"colsum {M}x{N}", db, dy.float().sum(0) + 1, 1e-4)
x = rn(100003)
xb = torch.empty(100003, device=dev, dtype=torch.bfloat16)
L.check(lib.vmae_cast_f32_to_bf16(ptr(x), ptr(xb), x.numel(), st()), "cast")
rep("cast f32->bf16", xb, x.bfloat16(), 1e-6)
xf = torch.empty(100003, device=dev)
L.check(lib.vmae_cast_bf16_to_f32(ptr(xb), ptr(xf), x.numel(), st()), "cast")
rep("cast bf16->f32", xf, xb.float(), 1e-6)
y = torch.ones(100003, device=dev)
L.check(lib.vmae_scale_f32(ptr(x), ptr(y), 0.25, 1, x.numel(), st()), "scale")
rep("scale accumulate", y, 1 + 0.25 * x, 1e-6)


# ---- masked MSE (both tubelet geometries)
def patchify(imgs, pt, ps):
    B, Cc, T, H, W = imgs.shape
    x = imgs.reshape(B, Cc, T // pt, pt, H // ps, ps, W // ps, ps)
    x = x.permute(0, 2, 4, 6, 3, 5, 7, 1)  # b t' h w dt py px c
    return x.reshape(B, (T // pt) * (H // ps) * (W // ps), pt * ps * ps * Cc)


for (B, T, S, pt, ps, norm) in [(2, 8, 112, 1, 8, 1), (2, 4, 64, 2, 16, 1), (1, 2, 32, 1, 8, 2), (1, 2, 32, 1, 8, 0)]:
    clip = rn(B, 3, T, S, S)
    tgt = patchify(clip, pt, ps)
    N, D = tgt.shape[1], tgt.shape[2]
    if norm:
        mean = tgt.mean(-1, keepdim=True)
        var = tgt.var(-1, keepdim=True, unbiased=(norm == 1))
        tgt = (tgt - mean) / (var + 1e-6).sqrt()
    pred = rn(B, N, D).bfloat16()
    mask = (torch.rand(B, N, device=dev, generator=g) < 0.9)
    m8 = mask.to(torch.uint8)
    acc = torch.zeros(2, device=dev)
    diff = torch.full((B, N, D), float("nan"), device=dev, dtype=torch.bfloat16)
    L.check(lib.vmae_masked_mse_fwd(ptr(clip), ptr(pred), 0, ptr(m8), ptr(acc), ptr(diff), B, T, S, S, pt, ps, norm, st()), "mse")
    loss = torch.empty(1, device=dev)
    scale = torch.empty(1, device=dev)
    L.check(lib.vmae_masked_mse_finalize(ptr(acc), ptr(loss), ptr(scale), D, st()), "fin")
    pr = pred.float().requires_grad_(True)
    lm = ((pr - tgt) ** 2).mean(-1)
    lref = (lm * mask).sum() / (mask.sum() + 1e-6)
    lref.backward()
    rep(f"mse loss B={B} T={T} S={S} pt={pt} ps={ps} norm={norm}", loss, lref.detach().reshape(1), 1e-4)
    gout = torch.full((1,), 3.0, device=dev)
    dpred = torch.empty(B, N, D, device=dev, dtype=torch.bfloat16)
    L.check(lib.vmae_masked_mse_bwd(ptr(diff), ptr(scale), ptr(gout), ptr(dpred), dpred.numel(), st()), "mseb")
    rep("mse dpred", dpred, 3.0 * pr.grad, 2e-2)

# ---- mask -> indices, gather tubelets, decoder input fwd/bwd, pos rows bwd
B, T, S, pt, ps, D = 3, 4, 64, 2, 16, 384
Tp, Lp = T // pt, (S // ps) ** 2
n_vis = 3
mask_frame = torch.ones(B, Lp, device=dev, dtype=torch.bool)
for b in range(B):
    perm = torch.randperm(Lp, device=dev, generator=g)[:n_vis]
    mask_frame[b, perm] = False
vis = torch.empty(B, n_vis, device=dev, dtype=torch.int32)
L.check(lib.vmae_mask_to_indices(ptr(mask_frame.to(torch.uint8)), ptr(vis), B, Lp, n_vis, st()), "m2i")
ref_vis = torch.stack([torch.nonzero(~mask_frame[b]).flatten() for b in range(B)]).int()
rep("mask_to_indices", vis, ref_vis, 1e-9)
clip = rn(B, 3, T, S, S)
tpos, spos, mtok = rn(Tp, D), rn(Lp, D), rn(D)
K = 3 * pt * ps * ps
patches = torch.empty(B * Tp * n_vis, K, device=dev, dtype=torch.bfloat16)
pos_rows = torch.empty(B * Tp * n_vis, D, device=dev)
L.check(lib.vmae_gather_tubelets(ptr(clip), ptr(vis), ptr(patches), ptr(tpos), ptr(spos), ptr(pos_rows), B, T, S, S, pt, ps, n_vis, D, st()), "gather")
# reference: conv3d-order unfold
u = clip.reshape(B, 3, Tp, pt, S // ps, ps, S // ps, ps).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, Tp, Lp, K)
ref_p = torch.stack([u[b][:, ref_vis[b].long()] for b in range(B)]).reshape(-1, K)
rep("gather_tubelets", patches, ref_p, 1e-2)
ref_pos = torch.stack([tpos[:, None, :] + spos[ref_vis[b].long()][None] for b in range(B)]).reshape(-1, D)
rep("pos_rows", pos_rows, ref_pos, 1e-6)
mask_full = mask_frame[:, None, :].expand(B, Tp, Lp).contiguous()
z = rn(B * Tp * n_vis, D).bfloat16()
x_dec = torch.full((B, Tp, Lp, D), float("nan"), device=dev)
L.check(lib.vmae_decoder_input_fwd(ptr(z), ptr(vis), ptr(mask_full.to(torch.uint8)), ptr(tpos), ptr(spos), ptr(mtok), ptr(x_dec), B, Tp, Lp, n_vis, D, st()), "dif")
zr = z.float().requires_grad_(True)
tr, sr, mr = tpos.clone().requires_grad_(True), spos.clone().requires_grad_(True), mtok.clone().requires_grad_(True)
full = torch.zeros(B, Tp, Lp, D, device=dev)
zz = zr.reshape(B, Tp, n_vis, D)
full = full.index_put((torch.arange(B, device=dev)[:, None, None], torch.arange(Tp, device=dev)[None, :, None], ref_vis.long()[:, None, :]), zz)
xr = full + (tr[None, :, None, :] + sr[None, None, :, :])
mm = mask_full[..., None].float()
xr = xr * (1 - mm) + mr * mm
rep("decoder_input fwd (gather form)", x_dec, xr, 1e-6)
dx = rn(B, Tp, Lp, D)
xr.backward(dx)
dz = torch.full_like(z, float("nan"))
dt_, ds_, dm_ = torch.zeros(Tp, D, device=dev), torch.zeros(Lp, D, device=dev), torch.zeros(D, device=dev)
L.check(lib.vmae_decoder_input_bwd(ptr(dx), ptr(vis), ptr(mask_full.to(torch.uint8)), ptr(dz), ptr(dt_), ptr(ds_), ptr(dm_), B, Tp, Lp, n_vis, D, st()), "dib")
rep("decoder_input bwd dz", dz, zr.grad, 1e-2)
rep("decoder_input bwd dtpos", dt_, tr.grad, 1e-4)
rep("decoder_input bwd dspos", ds_, sr.grad, 1e-4)
rep("decoder_input bwd dmask_token", dm_, mr.grad, 1e-4)
# blend form (vis_idx NULL)
zf = rn(B * Tp * Lp, D).bfloat16()
L.check(lib.vmae_decoder_input_fwd(ptr(zf), None, ptr(mask_full.to(torch.uint8)), ptr(tpos), ptr(spos), ptr(mtok), ptr(x_dec), B, Tp, Lp, 0, D, st()), "dif2")
xr2 = (zf.float().reshape(B, Tp, Lp, D) + (tpos[None, :, None] + spos[None, None])) * (1 - mm) + mtok * mm
rep("decoder_input fwd (blend form)", x_dec, xr2, 1e-6)
dzf = torch.full_like(zf, float("nan"))
L.check(lib.vmae_decoder_input_bwd(ptr(dx), None, ptr(mask_full.to(torch.uint8)), ptr(dzf), None, None, None, B, Tp, Lp, 0, D, st()), "dib2")
rep("decoder_input bwd dz (blend form)", dzf, (dx * (1 - mm)).reshape(-1, D), 1e-2)
dxr = rn(B * Tp * n_vis, D)
dt2, ds2 = torch.zeros(Tp, D, device=dev), torch.zeros(Lp, D, device=dev)
L.check(lib.vmae_pos_rows_bwd(ptr(dxr), ptr(vis), ptr(dt2), ptr(ds2), B, Tp, n_vis, D, st()), "prb")
rep("pos_rows_bwd dtpos", dt2, dxr.reshape(B, Tp, n_vis, D).sum((0, 2)), 1e-4)
ds_ref = torch.zeros(Lp, D, device=dev).index_add_(0, ref_vis.long()[:, None, :].expand(B, Tp, n_vis).reshape(-1), dxr)
rep("pos_rows_bwd dspos", ds2, ds_ref, 1e-4)

# ---- gaussian noise
n = 1 << 20
zz = rn(n)
out = torch.empty(n, device=dev)
eps = torch.empty(n, device=dev)
L.check(lib.vmae_add_gaussian_noise(ptr(zz), ptr(out), 1, ptr(eps), n, 0.1, 1234, 0, st()), "noise")
rep("noise z + sigma*eps", out, zz + 0.1 * eps, 1e-6)
torch.cuda.synchronize()
m, s = eps.mean().item(), eps.std().item()
good = abs(m) < 5e-3 and abs(s - 1) < 5e-3
OK &= good
print(f"{'OK ' if good else 'BAD'} noise moments mean={m:.4f} std={s:.4f} kurt={(eps**4).mean().item():.3f}")
print("ALL OK" if OK else "FAILURES")
sys.exit(0 if OK else 1)
