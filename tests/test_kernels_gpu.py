"""GPU: every kernel behind the C ABI against a plain PyTorch fp32 reference of the same op (bf16 tolerance for the
tensor-core kernels, fp32 tolerance for the reductions), including ragged / tail shapes."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 1.5e-2   # relative to the largest reference magnitude
F32_TOL = 2e-5


@pytest.fixture(scope="module")
def env():
    from ssl_vit_video_analytics_b200 import _lib, ops

    dev = torch.device("cuda:0")
    return ops, _lib, dev, torch.Generator(device="cuda").manual_seed(0)


def relerr(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert torch.isfinite(got).all()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-12))


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 384, 384), (1000, 1152, 384), (512, 1536, 384), (384, 384, 1536), (300, 192, 384),
                                   (128, 96, 200), (77, 576, 72)])
def test_gemm_forward_epilogues(env, M, N, K):
    ops, L, dev, g = env
    x = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device=dev, generator=g)
    res = torch.randn(M, N, device=dev, generator=g)
    ref = x.float() @ w.float().t()
    assert relerr(ops.gemm(x, w, M, N, K), ref) < BF16_TOL
    assert relerr(ops.gemm(x, w, M, N, K, bias=b), ref + b) < BF16_TOL
    dact = torch.empty(M, N, device=dev, dtype=torch.bfloat16)      # the GELU epilogue saves gelu'(pre-activation)
    y = ops.gemm(x, w, M, N, K, bias=b, act=L.ACT_GELU, aux_out=dact)
    pre = (ref + b).requires_grad_(True)
    F.gelu(pre).sum().backward()
    assert relerr(y, F.gelu(ref + b)) < BF16_TOL and relerr(dact, pre.grad) < BF16_TOL
    y32 = ops.gemm(x, w, M, N, K, bias=b, residual=res, out_dtype=torch.float32)
    assert y32.dtype == torch.float32 and relerr(y32, ref + b + res) < 1e-4
    h = torch.randn(M, N, device=dev, generator=g).bfloat16()
    assert relerr(ops.gemm(x, w, M, N, K, aux_in=h), ref * h.float()) < BF16_TOL      # dgrad epilogue: multiply by the saved factor


@pytest.mark.parametrize("M,Np,Kp", [(128, 64, 64), (1000, 1152, 384), (512, 1536, 384), (384, 384, 1536), (300, 192, 384)])
def test_gemm_dgrad_mn_major_b(env, M, Np, Kp):
    ops, L, dev, g = env
    dy = torch.randn(M, Np, device=dev, generator=g).bfloat16()
    w = (torch.randn(Np, Kp, device=dev, generator=g) * 0.05).bfloat16()
    assert relerr(ops.linear_dgrad(dy, w), dy.float() @ w.float()) < BF16_TOL


@pytest.mark.parametrize("T,Np,Kp", [(64, 128, 64), (4096, 1152, 384), (10240, 1536, 384), (3000, 384, 1536), (1000, 192, 384), (50, 64, 72)])
def test_gemm_wgrad_split_k_accumulates(env, T, Np, Kp):
    ops, L, dev, g = env
    dy = (torch.randn(T, Np, device=dev, generator=g) * 0.1).bfloat16()
    x = torch.randn(T, Kp, device=dev, generator=g).bfloat16()
    dw = torch.ones(Np, Kp, device=dev)
    ops.linear_wgrad(dy, x, dw)                       # accumulates into an existing gradient buffer
    assert relerr(dw, dy.float().t() @ x.float() + 1.0) < 1e-4


@pytest.mark.parametrize("T,Np,Kp", [(64, 128, 64), (4096, 1152, 384), (10240, 1536, 384), (3000, 384, 1536), (1000, 192, 384), (50, 64, 72)])
def test_wgrad_with_fused_bias_gradient(env, T, Np, Kp):
    ops, L, dev, g = env
    dy = (torch.randn(T, Np, device=dev, generator=g) * 0.1).bfloat16()
    x = torch.randn(T, Kp, device=dev, generator=g).bfloat16()
    dw = torch.ones(Np, Kp, device=dev)
    db = torch.full((Np,), 2.0, device=dev)
    ops.linear_wgrad_bias(dy, x, dw, db)               # both accumulate into existing buffers
    assert relerr(dw, dy.float().t() @ x.float() + 1.0) < 1e-4
    assert relerr(db, dy.float().sum(0) + 2.0) < 1e-4


def test_gemm_rejects_bad_arguments(env):
    ops, L, dev, g = env
    x = torch.zeros(128, 64, device=dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.gemm(x, x, 128, 60, 64)
    with pytest.raises(RuntimeError, match="accumulate"):
        ops.gemm(x, x, 128, 128, 64, k_splits=2)


@pytest.mark.parametrize("rows,cols", [(1000, 384), (333, 192), (64, 576), (50, 512), (7, 1024), (1, 96)])
def test_layernorm_fwd_bwd(env, rows, cols):
    ops, L, dev, g = env
    x = torch.randn(rows, cols, device=dev, generator=g) * 2 + 0.5
    gm, bt = torch.randn(cols, device=dev, generator=g), torch.randn(cols, device=dev, generator=g)
    y, mean, rstd = ops.layernorm_fwd(x, gm, bt)
    xr, gr, br = x.clone().requires_grad_(True), gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (cols,), gr, br, 1e-5)
    assert relerr(y, yr) < BF16_TOL
    y32, _, _ = ops.layernorm_fwd(x, gm, bt, out_dtype=torch.float32)
    assert relerr(y32, yr) < F32_TOL
    dy = torch.randn(rows, cols, device=dev, generator=g).bfloat16()
    res = torch.randn(rows, cols, device=dev, generator=g)
    yr.backward(dy.float())
    dx, dxb, dg, db = ops.layernorm_bwd(dy, x, gm, mean, rstd, dx_resid=res, want_bf16=True)
    assert relerr(dx, xr.grad + res) < F32_TOL and relerr(dxb, xr.grad + res) < BF16_TOL
    assert relerr(dg, gr.grad) < 1e-4 and relerr(db, br.grad) < 1e-4


def _ref_attn(qkv, H):
    B, Lq, W3 = qkv.shape
    hd = W3 // (3 * H)
    q, k, v = qkv.reshape(B, Lq, 3, H, hd).permute(2, 0, 3, 1, 4).float()
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, Lq, H * hd), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,Lq,H,hd", [(1, 16, 1, 64), (1, 64, 1, 64), (1, 100, 2, 64), (2, 160, 6, 64), (2, 300, 3, 64), (2, 1568, 6, 64),
                                       (9, 520, 7, 64), (1, 128, 1, 32), (3, 196, 12, 32), (2, 784, 6, 32), (1, 49, 18, 32)])
def test_flash_attention_fwd_bwd(env, B, Lq, H, hd):
    ops, L, dev, g = env
    qkv = torch.randn(B, Lq, 3 * H * hd, device=dev, generator=g).bfloat16()
    dout = torch.randn(B * Lq, H * hd, device=dev, generator=g).bfloat16()
    out, lse = ops.attn_fwd(qkv, B, Lq, H, hd)
    qf = qkv.float().requires_grad_(True)
    ro, rl = _ref_attn(qf, H)
    assert relerr(out.view(B, Lq, -1), ro) < BF16_TOL
    assert float((lse - rl).abs().max()) < 1e-2
    ro.backward(dout.float().view(B, Lq, -1))
    # both backward implementations: the two-kernel form (every head_dim) and the single-pass kernel (head_dim 64)
    for fused in ([False, True] if hd == 64 else [False]):
        dqkv = ops.attn_bwd(qkv, out, dout, lse, B, Lq, H, hd, fused=fused)
        for i, name in enumerate("qkv"):
            sl = slice(i * H * hd, (i + 1) * H * hd)
            assert relerr(dqkv[..., sl], qf.grad[..., sl]) < 2.5e-2, (name, fused)


def test_attention_rejects_unsupported_head_dim(env):
    ops, L, dev, g = env
    qkv = torch.zeros(1, 16, 3 * 48, device=dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="head_dim"):
        ops.attn_fwd(qkv, 1, 16, 1, 48)


@pytest.mark.parametrize("B,T,S,pt,ps,norm", [(2, 8, 112, 1, 8, 1), (2, 4, 64, 2, 16, 1), (1, 2, 32, 1, 8, 2), (1, 2, 32, 1, 8, 0),
                                                  (3, 6, 80, 2, 16, 0), (2, 2, 48, 2, 16, 2)])
def test_masked_mse_matches_reference_formula(env, B, T, S, pt, ps, norm):
    from oracle import restatement as R

    ops, L, dev, g = env
    clip = torch.randn(B, 3, T, S, S, device=dev, generator=g)
    N, D = (T // pt) * (S // ps) ** 2, pt * ps * ps * 3
    pred = torch.randn(B, N, D, device=dev, generator=g).bfloat16().requires_grad_(True)
    mask = torch.rand(B, N, device=dev, generator=g) < 0.9
    loss = ops.masked_mse(pred, clip, mask, pt, ps, norm)
    (loss * 3.0).backward()
    pr = pred.detach().float().cpu().requires_grad_(True)
    tgt = R.patchify(clip.cpu(), pt, ps)
    if norm:
        tgt = R.normalise_target(tgt, unbiased=(norm == 1))
    lm = ((pr - tgt) ** 2).mean(-1)
    lref = (lm * mask.cpu()).sum() / (mask.cpu().sum() + 1e-6)
    (lref * 3.0).backward()
    assert float(loss) == pytest.approx(float(lref), rel=1e-4)
    assert relerr(pred.grad, pr.grad) < BF16_TOL
    assert float(pred.grad[~mask].abs().max()) == 0.0      # visible tokens get exactly zero gradient


def test_masked_mse_all_visible_is_zero_not_nan(env):
    ops, L, dev, g = env
    clip = torch.randn(1, 3, 2, 16, 16, device=dev, generator=g)
    pred = torch.randn(1, 8, 192, device=dev, generator=g).bfloat16()
    loss = ops.masked_mse(pred, clip, torch.zeros(1, 8, dtype=torch.bool, device=dev), 1, 8, 1)
    assert float(loss) == 0.0


def test_mask_indices_gather_scatter(env):
    ops, L, dev, g = env
    B, T, S, pt, ps, D, n_vis = 3, 4, 64, 2, 16, 384, 3
    Tp, Lp = T // pt, (S // ps) ** 2
    frame = torch.ones(B, Lp, device=dev, dtype=torch.bool)
    for b in range(B):
        frame[b, torch.randperm(Lp, device=dev, generator=g)[:n_vis]] = False
    vis = ops.mask_to_indices(frame.to(torch.uint8), n_vis)
    ref_vis = torch.stack([torch.nonzero(~frame[b]).flatten() for b in range(B)])
    assert torch.equal(vis.long(), ref_vis)                 # gather indices: bit-exact
    clip = torch.randn(B, 3, T, S, S, device=dev, generator=g)
    tpos = torch.randn(1, Tp, 1, D, device=dev, generator=g, requires_grad=True)
    spos = torch.randn(1, 1, Lp, D, device=dev, generator=g, requires_grad=True)
    patches, pos_rows = ops.TubeletGatherFn.apply(clip, vis, tpos, spos, pt, ps)
    K = 3 * pt * ps * ps
    u = clip.reshape(B, 3, Tp, pt, S // ps, ps, S // ps, ps).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, Tp, Lp, K)
    ref_p = torch.stack([u[b][:, ref_vis[b]] for b in range(B)]).reshape(-1, K)
    assert relerr(patches, ref_p) < BF16_TOL
    ref_pos = torch.stack([tpos[0, :, 0][:, None] + spos[0, 0][ref_vis[b]][None] for b in range(B)]).reshape(-1, D)
    assert relerr(pos_rows, ref_pos) < 1e-6
    gr = torch.randn_like(pos_rows)
    pos_rows.backward(gr)
    gt, gs = torch.autograd.grad(ref_pos, (tpos, spos), gr)
    assert relerr(tpos.grad, gt) < 1e-4 and relerr(spos.grad, gs) < 1e-4
    # decoder input: scatter + pos + mask token, forward and backward
    mask_full = frame[:, None, :].expand(B, Tp, Lp).reshape(B, Tp * Lp).contiguous()
    z = torch.randn(B * Tp * n_vis, D, device=dev, generator=g).bfloat16().requires_grad_(True)
    tp2 = torch.randn(1, Tp, 1, D, device=dev, generator=g, requires_grad=True)
    sp2 = torch.randn(1, 1, Lp, D, device=dev, generator=g, requires_grad=True)
    mt = torch.randn(1, 1, D, device=dev, generator=g, requires_grad=True)
    x_dec = ops.DecoderInputFn.apply(z, vis, mask_full.to(torch.uint8), tp2, sp2, mt, B, Tp, Lp)
    full = torch.zeros(B, Tp, Lp, D, device=dev)
    full = full.index_put((torch.arange(B, device=dev)[:, None, None], torch.arange(Tp, device=dev)[None, :, None], ref_vis[:, None, :]),
                          z.float().reshape(B, Tp, n_vis, D))
    m = mask_full.view(B, Tp, Lp, 1).float()
    xr = ((full + (tp2 + sp2)) * (1 - m) + mt * m).reshape(B, Tp * Lp, D)
    assert relerr(x_dec, xr) < 1e-6
    gx = torch.randn_like(x_dec)
    refs = torch.autograd.grad(xr, (z, tp2, sp2, mt), gx)
    x_dec.backward(gx)
    assert relerr(z.grad, refs[0]) < BF16_TOL
    for a, r in zip((tp2.grad, sp2.grad, mt.grad), refs[1:]):
        assert relerr(a, r) < 1e-4


def test_small_helpers(env):
    ops, L, dev, g = env
    dy = torch.randn(5000, 1536, device=dev, generator=g).bfloat16()
    assert relerr(ops.colsum(dy), dy.float().sum(0)) < 1e-4
    x = torch.randn(100003, device=dev, generator=g)
    assert torch.equal(ops.cast_bf16(x), x.bfloat16()) and torch.equal(ops.cast_f32(x.bfloat16()), x.bfloat16().float())
    y = torch.ones_like(x)
    ops.scale_f32(x, 0.25, out=y, accumulate=True)
    assert torch.equal(y, 1 + x * 0.25)                       # separate IEEE mul + add: bit-exact
    dact = torch.randn(4096, device=dev, generator=g).bfloat16()
    d = torch.randn(4096, device=dev, generator=g).bfloat16()
    assert relerr(ops.gelu_bwd(d, dact), d.float() * dact.float()) < BF16_TOL


def test_gaussian_noise_statistics_and_determinism(env):
    from ssl_vit_video_analytics_b200.privacy.feature_noise import add_gaussian_noise

    ops, L, dev, g = env
    z = torch.randn(1 << 20, device=dev, generator=g)
    a = ops.GaussianNoiseFn.apply(z, 0.1, 1234, 0)
    b = ops.GaussianNoiseFn.apply(z, 0.1, 1234, 0)
    c = ops.GaussianNoiseFn.apply(z, 0.1, 1235, 0)
    assert torch.equal(a, b) and not torch.equal(a, c)
    eps = (a - z) / 0.1
    assert abs(float(eps.mean())) < 5e-3 and abs(float(eps.std()) - 1) < 5e-3 and abs(float((eps ** 4).mean()) - 3) < 0.05
    zz = z[:1000].clone().requires_grad_(True)
    out = add_gaussian_noise(zz, 0.2)
    out.sum().backward()
    assert torch.equal(zz.grad, torch.ones_like(zz)) and out is not zz
    assert add_gaussian_noise(zz, 0.0) is zz


def test_fedavg_aggregate_on_device_is_bit_exact_with_cpu(env):
    import torch.nn as nn

    from oracle import restatement as R
    from ssl_vit_video_analytics_b200.federated.fed_loop import fedavg_aggregate

    ops, L, dev, g = env
    torch.manual_seed(0)
    glob = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Flatten(), nn.Linear(8, 4)).to(dev)
    clients = []
    for i in range(5):
        m = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Flatten(), nn.Linear(8, 4)).to(dev)
        m[1].num_batches_tracked.fill_(i)
        clients.append({k: v.detach().clone() for k, v in m.state_dict().items()})
    w = [1907.0, 1893.0, 1921.0, 1880.0, 1936.0]
    ref = R.fedavg({k: v.cpu() for k, v in glob.state_dict().items()}, [{k: v.cpu() for k, v in c.items()} for c in clients], w)
    new = fedavg_aggregate(glob, clients, w)
    for k in ref:
        assert torch.equal(new[k], ref[k]), k


# ---- K7: TinyViT convolution kit ------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,Cc,act,res", [(3000, 96, True, False), (5000, 384, True, True), (777, 48, False, False), (64, 768, False, True)])
def test_batchnorm_train_fwd_bwd(env, M, Cc, act, res):
    ops, L, dev, g = env
    x = (torch.randn(M, Cc, device=dev, generator=g) * 1.7 + 0.3).bfloat16()
    gm = torch.rand(Cc, device=dev, generator=g) + 0.5
    bt = torch.randn(Cc, device=dev, generator=g) * 0.1
    rm, rv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    r = torch.randn(M, Cc, device=dev, generator=g).bfloat16() if res else None
    xg = x.clone().requires_grad_(True)
    gg, bg = gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    y = ops.BatchNormActFn.apply(xg, gg, bg, rm, rv, r, 1e-5, 0.1, L.ACT_GELU if act else L.ACT_NONE)
    # reference: nn.functional.batch_norm in train mode on the NCHW view
    xr = x.float().clone().requires_grad_(True)
    gr, br = gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    yr = F.batch_norm(xr.t().reshape(1, Cc, M, 1), rm2, rv2, gr, br, True, 0.1, 1e-5).reshape(Cc, M).t()
    if act:
        yr = F.gelu(yr)
    if res:
        yr = yr + r.float()
    assert relerr(y, yr) < BF16_TOL
    assert relerr(rm, rm2) < 1e-3 and relerr(rv, rv2) < 1e-3
    dy = torch.randn(M, Cc, device=dev, generator=g).bfloat16()
    y.backward(dy)
    yr.backward(dy.float())
    assert relerr(xg.grad, xr.grad) < 2e-2
    assert relerr(gg.grad, gr.grad) < 2e-2 and relerr(bg.grad, br.grad) < 2e-2


@pytest.mark.parametrize("N,H,W,Cc,stride", [(2, 14, 14, 384, 1), (3, 28, 28, 768, 2), (1, 7, 9, 96, 2), (2, 8, 8, 48, 1),
                                             (2, 56, 56, 128, 1), (2, 56, 56, 64, 2), (1, 60, 61, 64, 1), (1, 57, 117, 64, 2), (3, 5, 3, 64, 1),
                                             (2, 1, 9, 64, 1), (1, 3, 300, 64, 1)])
def test_depthwise_conv3x3_fwd_bwd(env, N, H, W, Cc, stride):
    ops, L, dev, g = env
    x = torch.randn(N, H, W, Cc, device=dev, generator=g).bfloat16().requires_grad_(True)
    w = (torch.randn(Cc, 1, 3, 3, device=dev, generator=g) * 0.3).requires_grad_(True)
    y = ops.DepthwiseConv3x3Fn.apply(x, w, stride)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, stride, 1, 1, Cc)
    assert y.shape == yr.permute(0, 2, 3, 1).shape
    assert relerr(y, yr.permute(0, 2, 3, 1)) < BF16_TOL
    dy = torch.randn_like(yr)
    y.backward(dy.permute(0, 2, 3, 1).bfloat16())
    yr.backward(dy.bfloat16().float())
    assert relerr(x.grad, xr.grad.permute(0, 2, 3, 1)) < BF16_TOL
    assert relerr(w.grad, wr.grad) < 2e-3


def test_squeeze_excite_pool_and_scale(env):
    ops, L, dev, g = env
    N, HW, Cc = 3, 49, 384
    x = torch.randn(N, HW, Cc, device=dev, generator=g).bfloat16().requires_grad_(True)
    gate = torch.rand(N, Cc, device=dev, generator=g).requires_grad_(True)
    pooled = ops.SEPoolFn.apply(x)
    y = ops.SEScaleFn.apply(x, gate)
    xr, gr = x.detach().float().requires_grad_(True), gate.detach().clone().requires_grad_(True)
    pr, yr = xr.mean(1), xr * gr[:, None, :]
    assert relerr(pooled, pr) < 1e-5 and relerr(y, yr) < BF16_TOL
    dp, dy = torch.randn_like(pr), torch.randn_like(yr).bfloat16()
    (pooled * dp).sum().backward(retain_graph=True)
    y.backward(dy)
    (pr * dp).sum().backward()
    yr.backward(dy.float())
    assert relerr(x.grad, xr.grad) < 2e-2 and relerr(gate.grad, gr.grad) < 1e-3


@pytest.mark.parametrize("stride", [1, 2])
def test_im2col_col2im_and_image_im2col(env, stride):
    ops, L, dev, g = env
    N, H, W, Cc = 2, 9, 10, 48
    x = torch.randn(N, H, W, Cc, device=dev, generator=g).bfloat16().requires_grad_(True)
    col = ops.Im2col3x3Fn.apply(x, stride)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    unf = F.unfold(xr, 3, padding=1, stride=stride)                    # [N, C*9, P] in (c, kh, kw) order
    P = unf.shape[-1]
    ref = unf.view(N, Cc, 9, P).permute(0, 3, 2, 1).reshape(N * P, 9 * Cc)   # -> (kh, kw, c)
    assert torch.equal(col.float(), ref.bfloat16().float())
    dcol = torch.randn_like(ref).bfloat16()
    col.backward(dcol)
    ref.backward(dcol.float())
    assert relerr(x.grad, xr.grad.permute(0, 2, 3, 1)) < BF16_TOL
    img = torch.randn(N, 3, H, W, device=dev, generator=g)
    c2, ho, wo = ops.im2col3x3_image(img, stride, 32)
    u2 = F.unfold(img, 3, padding=1, stride=stride).permute(0, 2, 1).reshape(-1, 27)
    assert torch.equal(c2[:, :27].float(), u2.bfloat16().float()) and float(c2[:, 27:].abs().max()) == 0.0


# ---- dropout (nn.TransformerEncoderLayer(dropout=0.1) of the reference decoder, mae_vit_adapter.py:40-47) -------------------
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_element_dropout_mask_statistics_and_apply(env, p):
    ops, L, dev, g = env
    n = 1 << 20
    seed = 1234567
    keep = ops.dropout_mask(n, p, seed, dev).bool()
    assert abs(float(keep.float().mean()) - (1 - p)) < 3e-3                       # 1M Bernoulli draws: sigma ~ 5e-4
    assert torch.equal(keep, ops.dropout_mask(n, p, seed, dev).bool())             # same seed -> same mask
    assert float((keep ^ ops.dropout_mask(n, p, seed + 1, dev).bool()).float().mean()) > 0.1
    pair = keep.view(-1, 2).float()                                                # the two halves of one hash word are independent
    assert abs(float((pair[:, 0] * pair[:, 1]).mean()) - (1 - p) ** 2) < 3e-3
    p_eff = round(p * 65536) / 65536
    src = torch.randn(n, device=dev, generator=g)
    add = torch.randn(n, device=dev, generator=g)
    ref = add + keep.float() * src / (1 - p_eff)
    assert relerr(ops.dropout_apply(src, p, seed, add=add, out_dtype=torch.float32), ref) < 1e-6
    assert relerr(ops.dropout_apply(src.bfloat16(), p, seed, out_dtype=torch.bfloat16), keep.float() * src.bfloat16().float() / (1 - p_eff)) < 5e-3
    assert torch.equal(ops.dropout_apply(src, 0.0, seed, out_dtype=torch.float32), src)


@pytest.mark.parametrize("B,Lq,H,hd", [(1, 64, 1, 64), (2, 160, 3, 64), (1, 300, 2, 64), (2, 196, 4, 32)])
def test_flash_attention_dropout_fwd_bwd(env, B, Lq, H, hd):
    """The kernels regenerate the keep mask from (seed, b, h, q, k); with that mask materialised, a plain PyTorch attention with
    explicit dropout must give the same output and the same q/k/v gradients."""
    ops, L, dev, g = env
    p, seed = 0.1, 987654321
    p_eff = round(p * 65536) / 65536
    qkv = torch.randn(B, Lq, 3, H, hd, device=dev, generator=g).bfloat16()
    dout = torch.randn(B, Lq, H * hd, device=dev, generator=g).bfloat16()
    keep = ops.attn_dropout_mask(B, Lq, H, p, seed, dev).float()                   # [B, H, L, L]
    assert abs(float(keep.mean()) - (1 - p_eff)) < 0.01
    out, lse = ops.attn_fwd(qkv.view(B * Lq, -1), B, Lq, H, hd, p, seed)
    q32 = qkv.float().requires_grad_(True)
    q, k, v = [q32[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    prob = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    ref = ((prob * keep / (1 - p_eff)) @ v).permute(0, 2, 1, 3).reshape(B, Lq, H * hd)
    ref.backward(dout.float())
    assert relerr(out.view(B, Lq, -1), ref) < BF16_TOL
    for fused in ([False, True] if hd == 64 else [False]):   # two-kernel form and single-pass kernel regenerate the same mask
        dqkv = ops.attn_bwd(qkv.view(B * Lq, -1), out, dout.view(B * Lq, -1), lse, B, Lq, H, hd, p, seed, fused=fused).view(B, Lq, 3, H, hd)
        for i, name in enumerate("qkv"):
            assert relerr(dqkv[:, :, i], q32.grad[:, :, i]) < 2 * BF16_TOL, (name, fused)
    out0, _ = ops.attn_fwd(qkv.view(B * Lq, -1), B, Lq, H, hd)                      # p = 0 path untouched
    assert relerr(out0.view(B, Lq, -1), (prob @ v).permute(0, 2, 1, 3).reshape(B, Lq, H * hd)) < BF16_TOL


def test_transformer_block_with_dropout_matches_explicit_masks(env):
    """One decoder layer with p = 0.1 at all four dropout sites vs fp32 PyTorch using the very masks the kernels generate."""
    ops, L, dev, g = env
    B, Lq, D, H, p = 2, 160, 384, 6, 0.1
    hd = D // H
    p_eff = round(p * 65536) / 65536
    torch.manual_seed(0)
    ln1, ln2 = torch.nn.LayerNorm(D).to(dev), torch.nn.LayerNorm(D).to(dev)
    qkv, proj = torch.nn.Linear(D, 3 * D).to(dev), torch.nn.Linear(D, D).to(dev)
    fc1, fc2 = torch.nn.Linear(D, 4 * D).to(dev), torch.nn.Linear(4 * D, D).to(dev)
    mods = [ln1, qkv, proj, ln2, fc1, fc2]
    x = torch.randn(B * Lq, D, device=dev, generator=g)
    seeds = [11, 22, 33, 44]
    xo = x.clone().requires_grad_(True)
    y = ops.transformer_block(xo, ln1, qkv.weight, qkv.bias, proj, ln2, fc1, fc2, H, B, Lq, p_drop=p, seeds=seeds)
    dy = torch.randn(B * Lq, D, device=dev, generator=g)
    y.backward(dy)
    got = {n_: p_.grad.clone() for m in mods for n_, p_ in [(f"{id(m)}.{k}", v) for k, v in m.named_parameters()]}
    gx = xo.grad.clone()
    for m in mods:
        m.zero_grad()
    # reference with explicit masks
    s_attn, s_d1, s_act, s_d2 = seeds
    k_attn = ops.attn_dropout_mask(B, Lq, H, p, s_attn, dev).float()
    k_d1 = ops.dropout_mask(B * Lq * D, p, s_d1, dev).float().view(B * Lq, D)
    k_act = ops.dropout_mask(B * Lq * 4 * D, p, s_act, dev).float().view(B * Lq, 4 * D)
    k_d2 = ops.dropout_mask(B * Lq * D, p, s_d2, dev).float().view(B * Lq, D)
    xr = x.clone().requires_grad_(True)
    qkv_r = qkv(ln1(xr)).view(B, Lq, 3, H, hd)
    q, k, v = [qkv_r[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    prob = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1) * k_attn / (1 - p_eff)
    o = (prob @ v).permute(0, 2, 1, 3).reshape(B * Lq, D)
    x1 = xr + proj(o) * k_d1 / (1 - p_eff)
    x2 = x1 + fc2(F.gelu(fc1(ln2(x1))) * k_act / (1 - p_eff)) * k_d2 / (1 - p_eff)
    x2.backward(dy)
    assert relerr(y, x2) < BF16_TOL
    assert relerr(gx, xr.grad) < 2 * BF16_TOL
    for m in mods:
        for kname, prm in m.named_parameters():
            ref_g, our_g = prm.grad, got[f"{id(m)}.{kname}"]
            cos = float((ref_g.flatten().double() @ our_g.flatten().double()) / (ref_g.norm().double() * our_g.norm().double() + 1e-300))
            assert cos > 0.999, (type(m).__name__, kname, cos)


@pytest.mark.parametrize("rows,cols,dt", [(1000, 576, torch.float32), (333, 192, torch.bfloat16), (7, 100, torch.float32)])
def test_stack2_row_losses_fwd_bwd(env, rows, cols, dt):
    """mae/losses.py:19-63 on CUDA tensors = the fused row-loss kernels; reference = the same formulas with tensor ops in fp32."""
    from ssl_vit_video_analytics_b200.mae import losses as ML

    ops, L, dev, g = env
    pred = torch.randn(rows, cols, device=dev, generator=g).to(dt)
    target = (torch.randn(rows, cols, device=dev, generator=g) * 2 + 0.5).to(dt)

    def ref_l2(p, t, normalize):
        if normalize:
            t = (t - t.mean(-1, keepdim=True)) / (t.var(-1, keepdim=True, unbiased=False) + 1e-6).sqrt()
        return ((p - t) ** 2).mean()

    cases = [("l2n", lambda p, t: ML.mae_l2_loss(p, t, True), lambda p, t: ref_l2(p, t, True)),
             ("l2", lambda p, t: ML.mae_l2_loss(p, t, False), lambda p, t: ref_l2(p, t, False)),
             ("l1", ML.mae_l1_loss, lambda p, t: (p - t).abs().mean()),
             ("cos", ML.mae_cosine_loss, lambda p, t: 1.0 - (F.normalize(p, dim=-1, eps=1e-6) * F.normalize(t, dim=-1, eps=1e-6)).sum(-1).mean())]
    tol = 1e-5 if dt == torch.float32 else 2e-2
    for name, ours, ref in cases:
        po = pred.clone().requires_grad_(True)
        lo = ours(po, target)
        (lo * 1.7).backward()
        pr = pred.float().clone().requires_grad_(True)
        lr = ref(pr, target.float())
        (lr * 1.7).backward()
        assert abs(float(lo) - float(lr)) <= 1e-5 + 1e-4 * abs(float(lr)), (name, float(lo), float(lr))
        assert relerr(po.grad, pr.grad) < tol, name
    assert ML.mae_recon_loss(pred, target, "l1", False).is_cuda


@pytest.mark.parametrize("flip", [True, False])
def test_frames_to_clip_matches_loader_transform(env, flip):
    """uint8 frames -> ToTensor + Normalize + channel reversal + [T,C,H,W]->[C,T,H,W] (mae_loader.py:63-77), done on the GPU."""
    ops, L, dev, g = env
    B, T, H, W = 2, 3, 17, 21
    frames = torch.randint(0, 256, (B, T, H, W, 3), device=dev, dtype=torch.uint8, generator=g)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    x = frames.float().permute(0, 1, 4, 2, 3) / 255.0                                  # [B, T, C, H, W] = ToTensor per frame
    x = (x - torch.tensor(mean, device=dev).view(1, 1, 3, 1, 1)) / torch.tensor(std, device=dev).view(1, 1, 3, 1, 1)
    if flip:
        x = x[:, :, [2, 1, 0]]
    ref = x.permute(0, 2, 1, 3, 4).contiguous()
    got = ops.frames_to_clip(frames, mean, std, flip_channels=flip)
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert float((got - ref).abs().max()) < 2e-6


def test_attn_bwd_c_abi_accepts_workspace_for_any_head_dim(env):
    """vmae_attn_bwd with a dQ accumulator: head_dim 64 takes the single-pass kernel, head_dim 32 silently uses the two-kernel form
    (the workspace is simply not needed) — same results either way."""
    ops, L, dev, g = env
    lib = L.load()
    p = lambda t: t.data_ptr() if t is not None else None
    st = torch.cuda.current_stream().cuda_stream
    for (B, Lq, H, hd) in [(2, 256, 3, 32), (2, 256, 3, 64)]:
        qkv = torch.randn(B, Lq, 3 * H * hd, device=dev, generator=g).bfloat16()
        dout = torch.randn(B * Lq, H * hd, device=dev, generator=g).bfloat16()
        out, lse = ops.attn_fwd(qkv, B, Lq, H, hd)
        ref = ops.attn_bwd(qkv, out, dout, lse, B, Lq, H, hd, fused=False)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(B, H, Lq, device=dev)
        acc = torch.empty(B * Lq, H * hd, device=dev)
        L.check(lib.vmae_attn_bwd(p(qkv), p(out), p(dout), p(lse), p(delta), p(acc), p(dqkv), B, Lq, H, hd, 1.0 / math.sqrt(hd), 0.0, 0, st), "bwd")
        assert relerr(dqkv, ref) < 1e-2


def test_fused_adamw_matches_torch_adamw(env):
    """optim.FusedAdamW (one launch per group, device-side hyper-parameters) vs torch.optim.AdamW: same trajectory over several
    steps incl. a learning-rate change, odd sizes, parameters that are misaligned views, two groups, and a state_dict round trip."""
    from ssl_vit_video_analytics_b200.optim import FusedAdamW

    ops, L, dev, g = env
    shapes = [(384, 384), (1536,), (3, 5, 7), (1,), (40000,), (33, 17)]
    base = torch.randn(sum(int(torch.tensor(s).prod()) for s in shapes) + 3, device=dev, generator=g)

    def make():
        ps, off = [], 1   # even entries: own (aligned) allocation; odd entries: views at an odd float offset (scalar path)
        for i, s in enumerate(shapes):
            n = int(torch.tensor(s).prod())
            t = base[off:off + n].clone().view(s) if i % 2 == 0 else base.clone()[off:off + n].view(s)
            off += n
            ps.append(torch.nn.Parameter(t))
        return ps

    pa, pb = make(), make()
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    oa = torch.optim.AdamW([{"params": pa[:3]}, {"params": pa[3:], "weight_decay": 0.0}], **kw)
    ob = FusedAdamW([{"params": pb[:3]}, {"params": pb[3:], "weight_decay": 0.0}], **kw)
    for it in range(6):
        if it == 3:
            for o in (oa, ob):
                o.param_groups[0]["lr"] = 1e-3
        if it == 4:   # checkpoint round trip in the middle of training
            sd = ob.state_dict()
            ob = FusedAdamW([{"params": pb[:3]}, {"params": pb[3:], "weight_decay": 0.0}], **kw)
            ob.load_state_dict(sd)
            assert ob.param_groups[0]["lr"] == 1e-3
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=dev, generator=g) * (0.1 + it)
            a.grad = gr.clone()
            b.grad = gr.clone()
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        assert relerr(b.detach(), a.detach()) < 2e-6, a.shape
    sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
    for k in sa:
        assert float(sb[k]["step"]) == float(sa[k]["step"]) == 6.0
        assert relerr(sb[k]["exp_avg"], sa[k]["exp_avg"]) < 2e-6
        assert relerr(sb[k]["exp_avg_sq"], sa[k]["exp_avg_sq"]) < 2e-6
    with pytest.raises(RuntimeError, match="fp32 CUDA"):
        p = torch.nn.Parameter(torch.zeros(4))
        p.grad = torch.zeros(4)
        FusedAdamW([p]).step()


@pytest.mark.gpu
def test_gemm_cta_pairs_with_multicast_b_tile(env):
    """VMAE_GEMM_PAIRS=1 launches every GEMM as clusters of two CTAs that share the B tile by TMA multicast (csrc/gemm.cu, CL = 2).
    The switch is read once per process, so the three operand layouts and all epilogues run in a child process."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    child_env = dict(os.environ, VMAE_GEMM_PAIRS="1")
    r = subprocess.run([sys.executable, str(root / "tools" / "dev_check_gemm.py"), "all"], env=child_env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mode all: ALL OK" in r.stdout and "BAD" not in r.stdout
