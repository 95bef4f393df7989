"""Tensor-level wrappers over the C ABI (include/vmae.h) and the autograd Functions built from them.

Everything here launches hand-written sm_100a kernels from libvmae_sm100a.so on the current CUDA stream. PyTorch is
used for device memory (caching allocator), streams and autograd bookkeeping only. There is no fallback: a CPU tensor
or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os as _os
import weakref
from typing import Optional

import torch

from . import _lib as L

BF16, F32 = torch.bfloat16, torch.float32

# number of kernels of ours launched since reset_launch_count() (bench.py reports it as gpu_launches)
_launches = 0


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def launch_count() -> int:
    return _launches


def add_launch_count(n: int) -> None:
    """A CUDA-graph replay launches the kernels that were counted while it was captured."""
    _count(n)


def _count(n: int = 1) -> None:
    global _launches
    _launches += n


# optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline leg)
_PROF = None


def profile_start() -> None:
    global _PROF
    _PROF = {}


def profile_stop() -> dict:
    """-> {kernel class: {"ms": total device ms, "launches": n, "work": total flops (or bytes)}}"""
    global _PROF
    rec, _PROF = _PROF, None
    torch.cuda.synchronize()
    out = {}
    for name, evs in (rec or {}).items():
        out[name] = {"ms": sum(e0.elapsed_time(e1) for e0, e1, _ in evs), "launches": len(evs), "work": float(sum(w for _, _, w in evs))}
    return out


def _run(name: str, work: float, call):
    if _PROF is None:
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = call()
    e1.record()
    _PROF.setdefault(name, []).append((e0, e1, work))
    return rc


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == BF16:
        return L.DTYPE_BF16
    if t.dtype == F32:
        return L.DTYPE_F32
    raise TypeError(f"unsupported dtype {t.dtype} (bf16 or fp32 only)")


def _need_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ssl_vit_video_analytics_b200 ops run on CUDA (sm_100a) tensors only; there is no CPU path")


# ======================================================================================================================
# raw kernels
# ======================================================================================================================
def gemm(a, b, M, N, K, *, a_major=L.MAJOR_K, b_major=L.MAJOR_K, bias=None, residual=None, aux_in=None, aux_out=None,
         out=None, out_dtype=BF16, accumulate=False, act=L.ACT_NONE, k_splits=1):
    """out[M,N] = epi(A[M,K] @ B[N,K]^T); see vmae_gemm_bf16 in include/vmae.h for operand-major conventions."""
    _need_cuda(a, b)
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(-1) == 1 and b.stride(-1) == 1
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    lib = L.load()
    # the gelu' side tensor may be one byte per element (vmae.h: VMAE_ACT_GELU_Q8 / VMAE_ACT_AUX_Q8): selected by its dtype
    if aux_out is not None and aux_out.dtype == torch.uint8:
        assert act == L.ACT_GELU
        act = L.ACT_GELU_Q8
    if aux_in is not None and aux_in.dtype == torch.uint8:
        act = L.ACT_AUX_Q8
    kind = "gemm_wgrad" if a_major == L.MAJOR_MN else ("gemm_dgrad" if b_major == L.MAJOR_MN else "gemm_fwd")
    rc = _run(kind, 2.0 * M * N * K, lambda: lib.vmae_gemm_bf16(
        _p(a), a.stride(0), a_major, _p(b), b.stride(0), b_major, M, N, K, _p(bias), _p(residual),
        residual.stride(0) if residual is not None else 0, _p(aux_in), aux_in.stride(0) if aux_in is not None else 0, _p(aux_out),
        aux_out.stride(0) if aux_out is not None else 0, _p(out), out.stride(0), _dt(out), 1 if accumulate else 0, act, k_splits,
        _stream()))
    L.check(rc, "vmae_gemm_bf16")
    _count()
    return out


def _wgrad_splits(tokens: int, n_out: int, k_out: int, sms: int = 148) -> int:
    """Split-K factor of dW[n_out, k_out] += dy^T x over `tokens`. Mirrors the kernel's tile choice (pick_block_n in gemm.cu: BLOCK_N
    in {192, 128}, cost rounds * (BLOCK_N + 48)) and minimises rounds * (k-blocks per split * (BLOCK_N + 48) + epilogue): the number of
    units should land just under a multiple of the SM count with the larger tile (measured on the decoder shapes: qkv 138 -> 100 us)."""
    m_tiles = (n_out + 127) // 128
    kb = (tokens + 63) // 64
    best, best_cost = 1, None
    for ks in range(1, min(kb, 64) + 1):
        cand = []
        for bn in (192, 128):
            units = m_tiles * ((k_out + bn - 1) // bn) * ks
            cand.append((((units + sms - 1) // sms) * (bn + 48), bn, units))
        _, bn, units = min(cand)           # what the kernel will pick for this ks
        rounds = (units + sms - 1) // sms
        cost = rounds * (((kb + ks - 1) // ks) * (bn + 48) + 2500)
        if best_cost is None or cost < best_cost:
            best, best_cost = ks, cost
    return best


def linear_fwd(x, w_bf16, bias=None, *, act=L.ACT_NONE, residual=None, aux_out=None, out_dtype=BF16):
    M, K = x.shape
    N = w_bf16.shape[0]
    return gemm(x, w_bf16, M, N, K, bias=bias, residual=residual, aux_out=aux_out, act=act,
                out_dtype=F32 if residual is not None else out_dtype)


def linear_dgrad(dy, w_bf16, *, aux_in=None):
    """dx[M,K] = dy[M,N] @ W[N,K] (optionally * gelu'(aux_in))"""
    M, N = dy.shape
    K = w_bf16.shape[1]
    return gemm(dy, w_bf16, M, K, N, b_major=L.MAJOR_MN, aux_in=aux_in)


def linear_wgrad(dy, x, dw=None):
    """dW[N,K] (+)= dy[M,N]^T @ x[M,K] in fp32 (split-K with red.add)"""
    M, N = dy.shape
    K = x.shape[1]
    if dw is None:
        dw = torch.zeros((N, K), device=dy.device, dtype=F32)
    gemm(dy, x, N, K, M, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=dw, accumulate=True, k_splits=_wgrad_splits(M, N, K))
    return dw


def linear_wgrad_bias(dy, x, dw=None, db=None):
    """dW[N,K] += dy^T x and db[N] += colsum(dy) in ONE launch (the bias gradient rides on an extra N=16 UMMA)."""
    M, N = dy.shape
    K = x.shape[1]
    if dw is None:
        dw = torch.zeros((N, K), device=dy.device, dtype=F32)
    if db is None:
        db = torch.zeros((N,), device=dy.device, dtype=F32)
    rc = _run("gemm_wgrad", 2.0 * M * N * K, lambda: L.load().vmae_linear_wgrad(
        _p(dy), dy.stride(0), _p(x), x.stride(0), M, N, K, _p(dw), dw.stride(0), _p(db), _wgrad_splits(M, N, K), _stream()))
    L.check(rc, "vmae_linear_wgrad")
    _count()
    return dw, db


def colsum(dy, db=None):
    M, N = dy.shape
    if db is None:
        db = torch.zeros((N,), device=dy.device, dtype=F32)
    L.check(_run("colsum", float(M) * N * dy.element_size(), lambda: L.load().vmae_colsum_accum(
        _p(dy), _dt(dy), dy.stride(0), _p(db), M, N, _stream())), "vmae_colsum_accum")
    _count()
    return db


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    if x.dtype == BF16:
        return x if x.is_contiguous() else x.contiguous()
    x = x.contiguous()
    assert x.dtype == F32
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    L.check(_run("cast", 6.0 * x.numel(), lambda: L.load().vmae_cast_f32_to_bf16(_p(x), _p(y), x.numel(), _stream())), "vmae_cast_f32_to_bf16")
    _count()
    return y


# id(parameter) -> (weakref to the parameter, bf16 operand copy made by WeightCache.refresh() for the current step)
_WEIGHT_BF16 = {}


def cast_weight(w: torch.Tensor) -> torch.Tensor:
    """bf16 GEMM operand of a weight: the WeightCache copy while the owning model's forward scope is open, else a cast."""
    e = _WEIGHT_BF16.get(id(w))
    if e is not None and e[0]() is w and e[1].device == w.device:   # identity check: ids are recycled after a model dies
        return e[1]
    return cast_bf16(w)


class WeightCache:
    """bf16 copies of all matrix-shaped fp32 parameters of a module, refreshed with ONE kernel launch per step
    (instead of one cast launch per weight). The model calls refresh() at the start of forward."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.ndim >= 2 and p.dtype == F32 and p.is_cuda and p.numel() >= 1024]
        self.key = tuple((p.data_ptr(), p.numel()) for p in self.params)
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8  # keep every copy 16-byte aligned
        self.flat = torch.empty(total, device=dev, dtype=BF16)
        self.copies = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
        table = []
        for p, c in zip(self.params, self.copies):
            table += [p.data_ptr(), c.data_ptr(), p.numel()]
        self.table = torch.tensor(table, dtype=torch.int64).to(dev)

    def stale(self) -> bool:
        return self.key != tuple((p.data_ptr(), p.numel()) for p in self.params)

    def refresh(self) -> None:
        L.check(_run("cast", 6.0 * self.flat.numel(), lambda: L.load().vmae_multi_cast_f32_to_bf16(
            _p(self.table), len(self.params), _stream())), "vmae_multi_cast_f32_to_bf16")
        _count()
        for p, c in zip(self.params, self.copies):
            _WEIGHT_BF16[id(p)] = (weakref.ref(p), c)

    def release(self) -> None:
        for p in self.params:
            _WEIGHT_BF16.pop(id(p), None)


def refresh_weight_cache(module: torch.nn.Module):
    """(Re)build the module's cache if parameters moved, then one launch that refreshes every bf16 copy. Returns the cache."""
    cache = module.__dict__.get("_vmae_weight_cache")
    if cache is None or cache.stale():
        if not any(p.is_cuda for p in module.parameters()):
            return None
        cache = WeightCache(module)
        module.__dict__["_vmae_weight_cache"] = cache
    cache.refresh()
    return cache


_SCOPE_DEPTH = 0


class weight_cache_scope:
    """`with weight_cache_scope(model):` around a model's forward. The OUTERMOST scope refreshes the bf16 copies of all of the
    model's weight matrices with one launch and publishes them to `cast_weight`; nested scopes (an encoder called by the MAE
    wrapper) reuse them. On exit the copies are withdrawn, so a weight is never served from a copy made before the last
    optimizer step: outside a scope (a bare sub-module call, an activation-checkpoint recompute) `cast_weight` casts afresh.
    Backward is unaffected: the autograd Functions save the bf16 operands they used."""

    def __init__(self, module: torch.nn.Module):
        self.module = module
        self.cache = None
        self.owner = False

    def __enter__(self):
        global _SCOPE_DEPTH
        self.owner = _SCOPE_DEPTH == 0
        _SCOPE_DEPTH += 1
        if self.owner:
            self.cache = refresh_weight_cache(self.module)
        return self

    def __exit__(self, *exc):
        global _SCOPE_DEPTH
        _SCOPE_DEPTH -= 1
        if self.owner and self.cache is not None:
            self.cache.release()
        return False


def cast_f32(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    if x.dtype == F32:
        return x if x.is_contiguous() else x.contiguous()
    x = x.contiguous()
    assert x.dtype == BF16
    y = torch.empty(x.shape, device=x.device, dtype=F32)
    L.check(L.load().vmae_cast_bf16_to_f32(_p(x), _p(y), x.numel(), _stream()), "vmae_cast_bf16_to_f32")
    _count()
    return y


def layernorm_fwd(x, gamma, beta, eps=1e-5, out_dtype=BF16, save_stats=True):
    rows, cols = x.shape
    y = torch.empty((rows, cols), device=x.device, dtype=out_dtype)
    mean = torch.empty((rows,), device=x.device, dtype=F32) if save_stats else None
    rstd = torch.empty((rows,), device=x.device, dtype=F32) if save_stats else None
    L.check(_run("layernorm_fwd", float(rows) * cols * (x.element_size() + y.element_size()), lambda: L.load().vmae_layernorm_fwd(
        _p(x), _dt(x), _p(gamma), _p(beta), _p(y), _dt(y), _p(mean), _p(rstd), rows, cols, eps, _stream())), "vmae_layernorm_fwd")
    _count()
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, *, dx_resid=None, dgamma=None, dbeta=None, want_bf16=False, want_f32=True):
    rows, cols = x.shape
    dx = torch.empty((rows, cols), device=x.device, dtype=F32) if want_f32 else None
    dx_b = torch.empty((rows, cols), device=x.device, dtype=BF16) if want_bf16 else None
    if dgamma is None:
        dgamma = torch.zeros((cols,), device=x.device, dtype=F32)
    if dbeta is None:
        dbeta = torch.zeros((cols,), device=x.device, dtype=F32)
    nbytes = float(rows) * cols * (dy.element_size() + x.element_size() + (4 if dx_resid is not None else 0) + (4 if want_f32 else 0)
                                   + (2 if want_bf16 else 0))
    L.check(_run("layernorm_bwd", nbytes, lambda: L.load().vmae_layernorm_bwd(
        _p(dy), _dt(dy), _p(x), _dt(x), _p(gamma), _p(mean), _p(rstd), _p(dx_resid), _p(dx), _p(dx_b), _p(dgamma), _p(dbeta), rows, cols,
        _stream())), "vmae_layernorm_bwd")
    _count()
    return dx, dx_b, dgamma, dbeta


def attn_fwd(qkv, B, Lq, H, hd, p_drop=0.0, seed=0):
    out = torch.empty((B * Lq, H * hd), device=qkv.device, dtype=BF16)
    lse = torch.empty((B, H, Lq), device=qkv.device, dtype=F32)
    L.check(_run("attn_fwd", 4.0 * B * H * Lq * Lq * hd, lambda: L.load().vmae_attn_fwd(
        _p(qkv), _p(out), _p(lse), B, Lq, H, hd, 1.0 / math.sqrt(hd), float(p_drop), int(seed), _stream())), "vmae_attn_fwd")
    _count()
    return out, lse


GELU_PRIME_Q8 = _os.environ.get("VMAE_GELU_Q8", "1") != "0"   # fused blocks save gelu' as uint8 (A/B switch)
ATTN_BWD_FUSED = True   # single-pass backward (head_dim 64); False forces the two-kernel form
ATTN_BWD_FUSED_MIN_L = int(_os.environ.get("VMAE_ATTN_FUSED_MIN_L", "128"))


def attn_bwd(qkv, out, dout, lse, B, Lq, H, hd, p_drop=0.0, seed=0, fused=None):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, H, Lq), device=qkv.device, dtype=F32)
    # measured: the single-pass kernel wins from a few hundred tokens up (L=1568: 1059 vs 1322 us); short sequences (L=160) are
    # setup-dominated and stay on the two-kernel form
    use_fused = ((ATTN_BWD_FUSED and Lq >= ATTN_BWD_FUSED_MIN_L) if fused is None else fused) and hd == 64
    dq_acc = torch.empty((B * Lq, H * hd), device=qkv.device, dtype=F32) if use_fused else None
    L.check(_run("attn_bwd", 10.0 * B * H * Lq * Lq * hd, lambda: L.load().vmae_attn_bwd(
        _p(qkv), _p(out), _p(dout), _p(lse), _p(delta), _p(dq_acc), _p(dqkv), B, Lq, H, hd, 1.0 / math.sqrt(hd), float(p_drop), int(seed),
        _stream())), "vmae_attn_bwd")
    _count(4 if use_fused else 3)
    return dqkv


def attn_dropout_mask(B, Lq, H, p_drop, seed, device):
    """uint8 [B, H, L, L] keep mask of the in-kernel attention-probability dropout (test helper)."""
    mask = torch.empty((B, H, Lq, Lq), device=device, dtype=torch.uint8)
    L.check(L.load().vmae_attn_dropout_mask(_p(mask), B, Lq, H, float(p_drop), int(seed), _stream()), "vmae_attn_dropout_mask")
    return mask


def dropout_apply(src, p_drop, seed, *, add=None, out_dtype=None, out=None):
    """out = (add or 0) + keep(seed, i) * src / (1 - p): nn.Dropout with a stateless, reproducible mask (same seed -> same mask,
    which is how backward re-applies it). src bf16/fp32; add fp32 residual (optional); out bf16/fp32, may be src itself."""
    _need_cuda(src)
    assert src.is_contiguous() and src.numel() % 4 == 0
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=out_dtype or src.dtype)
    if add is not None:
        assert add.dtype == F32 and add.is_contiguous() and add.numel() == src.numel()
    n = src.numel()
    nbytes = float(n) * (src.element_size() + out.element_size() + (4 if add is not None else 0))
    L.check(_run("dropout", nbytes, lambda: L.load().vmae_dropout_apply(_p(src), _dt(src), _p(add), _p(out), _dt(out), n, float(p_drop), int(seed),
                                                                       _stream())), "vmae_dropout_apply")
    _count()
    return out


def dropout_mask(n, p_drop, seed, device):
    """uint8 [n] keep mask dropout_apply uses for (p_drop, seed) (test helper)."""
    mask = torch.empty((n,), device=device, dtype=torch.uint8)
    L.check(L.load().vmae_dropout_mask(_p(mask), n, float(p_drop), int(seed), _stream()), "vmae_dropout_mask")
    return mask


def new_dropout_seeds(k):
    """k fresh 63-bit seeds from torch's CPU generator (deterministic under torch.manual_seed, no device sync)."""
    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        raise RuntimeError("dropout inside a CUDA-graph capture: the seeds are host-side kernel arguments, so every replay would reuse "
                           "the same keep masks; capture with dropout 0 or run the step eagerly")
    return [int(v) for v in torch.randint(0, 2 ** 62, (k,), dtype=torch.int64).tolist()]


# ======================================================================================================================
# autograd Functions
# ======================================================================================================================
def _amp_dtype_out(like_fp32: bool = False):
    """dtype a Linear-like op should return: bf16 under CUDA autocast (as nn.Linear would), else fp32."""
    if like_fp32:
        return F32
    return BF16 if torch.is_autocast_enabled("cuda") else F32


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b): nn.Linear (+ optional GELU) on the tcgen05 GEMM. x: [..., K] bf16/fp32; y bf16.
    `weight` may carry trailing singleton/kernel dims (e.g. a 1x1 Conv2d weight [N, K, 1, 1]); it is used as [N, -1]."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        _need_cuda(x, weight)
        shp = x.shape
        x2 = cast_bf16(x.reshape(-1, shp[-1]))
        n_out = weight.shape[0]
        wb = cast_weight(weight).reshape(n_out, -1)
        bf = bias.float().contiguous() if bias is not None else None
        pre = torch.empty((x2.shape[0], n_out), device=x.device, dtype=BF16) if act == L.ACT_GELU else None
        y = linear_fwd(x2, wb, bf, act=act, aux_out=pre)
        ctx.save_for_backward(x2, wb, pre)
        ctx.has_bias = bias is not None
        ctx.in_shape = shp
        ctx.in_dtype = x.dtype
        ctx.w_shape = weight.shape
        ctx.params = (weight, bias)   # the Parameter objects: direct gradient accumulation (enable_direct_grad) needs their .grad
        return y.reshape(*shp[:-1], n_out)

    @staticmethod
    def backward(ctx, dy):
        x2, wb, pre = ctx.saved_tensors
        dy2 = cast_bf16(dy.reshape(-1, dy.shape[-1]))
        if pre is not None:
            dy2 = gelu_bwd(dy2, pre)
        dx = linear_dgrad(dy2, wb) if ctx.needs_input_grad[0] else None
        dw = db = None
        need_w = ctx.needs_input_grad[1]
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        weight, bias = ctx.params
        # accumulate straight into .grad where the owner of the buffers opted in: no zero-filled temporary, no "grad += new" pass
        t_w = _direct_grad(weight) if need_w else None
        t_b = _direct_grad(bias) if need_b else None
        if need_w and need_b and (t_w is None) == (t_b is None):
            dw, db = linear_wgrad_bias(dy2, x2, None if t_w is None else t_w.view(t_w.shape[0], -1), t_b)
        else:
            if need_w:
                dw = linear_wgrad(dy2, x2, None if t_w is None else t_w.view(t_w.shape[0], -1))
            if need_b:
                db = colsum(dy2, t_b)
        if t_b is not None:
            db = None
            _grad_done(bias)
        if t_w is not None:
            dw = None
            _grad_done(weight)
        if dw is not None:
            dw = dw.reshape(ctx.w_shape)
        if dx is not None:
            dx = dx.reshape(ctx.in_shape)
            if ctx.in_dtype == F32:
                dx = cast_f32(dx)
        return dx, dw, db, None


class LinearNoiseFn(torch.autograd.Function):
    """y = x W^T + b + sigma * eps with eps ~ N(0, I) drawn in the GEMM epilogue (privacy/feature_noise.py:4-7 fused into the
    producing Linear; Philox keyed on (seed, offset, element)). eps_out (optional fp32 [rows, N]) receives eps. The noise is an
    additive constant: gradients are those of the Linear."""

    @staticmethod
    def forward(ctx, x, weight, bias, sigma, seed, offset, eps_out):
        _need_cuda(x, weight)
        shp = x.shape
        x2 = cast_bf16(x.reshape(-1, shp[-1]))
        n_out = weight.shape[0]
        wb = cast_weight(weight).reshape(n_out, -1)
        bf = bias.float().contiguous() if bias is not None else None
        M, K = x2.shape
        y = torch.empty((M, n_out), device=x.device, dtype=BF16)
        if eps_out is not None:
            assert eps_out.dtype == F32 and eps_out.is_contiguous() and eps_out.numel() == M * n_out and eps_out.is_cuda
        L.check(_run("gemm_fwd", 2.0 * M * n_out * K, lambda: L.load().vmae_gemm_bf16_noise(
            _p(x2), x2.stride(0), _p(wb), wb.stride(0), M, n_out, K, _p(bf), _p(y), y.stride(0), float(sigma), int(seed) & 0xFFFFFFFFFFFFFFFF,
            int(offset), _p(eps_out), _stream())), "vmae_gemm_bf16_noise")
        _count()
        ctx.save_for_backward(x2, wb)
        ctx.has_bias = bias is not None
        ctx.in_shape = shp
        ctx.in_dtype = x.dtype
        ctx.w_shape = weight.shape
        return y.reshape(*shp[:-1], n_out)

    @staticmethod
    def backward(ctx, dy):
        x2, wb = ctx.saved_tensors
        dy2 = cast_bf16(dy.reshape(-1, dy.shape[-1]))
        dx = linear_dgrad(dy2, wb) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] and ctx.has_bias and ctx.needs_input_grad[2]:
            dw, db = linear_wgrad_bias(dy2, x2)
        elif ctx.needs_input_grad[1]:
            dw = linear_wgrad(dy2, x2)
        elif ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        if dw is not None:
            dw = dw.reshape(ctx.w_shape)
        if dx is not None:
            dx = dx.reshape(ctx.in_shape)
            if ctx.in_dtype == F32:
                dx = cast_f32(dx)
        return dx, dw, db, None, None, None, None


_noise_offset = [0]


def linear_with_feature_noise(x, weight, bias, sigma, seed=None, eps_out=None):
    """nn.Linear followed by add_gaussian_noise(., sigma), as one kernel. sigma <= 0: the plain Linear (the reference returns z
    itself). Successive calls advance a process-wide Philox offset so that they draw fresh normals under one seed."""
    if sigma <= 0:
        return LinearFn.apply(x, weight, bias, L.ACT_NONE)
    if seed is None:
        seed = torch.initial_seed()
    rows = x.numel() // x.shape[-1]
    off = _noise_offset[0]
    _noise_offset[0] += (rows * weight.shape[0] + 3) // 4
    return LinearNoiseFn.apply(x, weight, bias, float(sigma), int(seed), off, eps_out)


def gelu_bwd(dy_bf16, pre_bf16):
    """dy * dact, where dact = gelu'(pre-activation) as saved by the GELU epilogue; used by the unfused LinearFn only."""
    out = torch.empty_like(dy_bf16)
    L.check(L.load().vmae_gelu_bwd(_p(dy_bf16), _p(pre_bf16), _p(out), dy_bf16.numel(), _stream()), "vmae_gelu_bwd")
    _count()
    return out


def linear(x, weight, bias=None, act=L.ACT_NONE):
    y = LinearFn.apply(x, weight, bias, act)
    out_dt = _amp_dtype_out()
    return y if out_dt == BF16 else cast_f32_autograd(y)


class _CastF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return cast_f32(x)

    @staticmethod
    def backward(ctx, g):
        return cast_bf16(g)


class _CastBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.src = x.dtype
        return cast_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return cast_f32(g) if ctx.src == F32 else g


def cast_f32_autograd(x):
    return x if x.dtype == F32 else _CastF32.apply(x)


def cast_bf16_autograd(x):
    return x if x.dtype == BF16 else _CastBF16.apply(x)


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dim; fp32/bf16 in, bf16 or fp32 out."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        _need_cuda(x, weight)
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.dtype not in (BF16, F32):
            raise TypeError(f"LayerNormFn: dtype {x2.dtype}")
        x2 = x2.contiguous()
        g, b = weight.float().contiguous(), bias.float().contiguous()
        y, mean, rstd = layernorm_fwd(x2, g, b, eps, out_dtype)
        ctx.save_for_backward(x2, g, mean, rstd)
        ctx.in_shape = shp
        ctx.in_dtype = x.dtype
        ctx.params = (weight, bias)
        return y.reshape(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, g, mean, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        want_b = ctx.in_dtype == BF16
        weight, bias = ctx.params
        t_g = _direct_grad(weight) if ctx.needs_input_grad[1] else None
        t_b = _direct_grad(bias) if ctx.needs_input_grad[2] else None
        if t_g is None or t_b is None:   # the kernel writes both or neither in place
            t_g = t_b = None
        dx, dxb, dg, db = layernorm_bwd(dy2, x2, g, mean, rstd, dgamma=t_g, dbeta=t_b, want_bf16=want_b, want_f32=not want_b)
        if t_g is not None:
            dg = db = None
            _grad_done(bias)
            _grad_done(weight)
        dxo = dxb if want_b else dx
        return dxo.reshape(ctx.in_shape), dg, db, None, None


def layer_norm(x, weight, bias, eps=1e-5, out_dtype=None):
    if out_dtype is None:
        out_dtype = F32 if (x.dtype == F32 and not torch.is_autocast_enabled("cuda")) else (x.dtype if x.dtype == BF16 else F32)
    return LayerNormFn.apply(x, weight, bias, eps, out_dtype)


class AttentionFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(hd)) v on the packed projection qkv [B, L, 3*H*hd] -> [B, L, H*hd] (bf16)."""

    @staticmethod
    def forward(ctx, qkv, num_heads):
        _need_cuda(qkv)
        B, Lq, W3 = qkv.shape
        hd = W3 // (3 * num_heads)
        q = cast_bf16(qkv)
        out, lse = attn_fwd(q, B, Lq, num_heads, hd)
        ctx.save_for_backward(q, out, lse)
        ctx.dims = (B, Lq, num_heads, hd)
        ctx.in_dtype = qkv.dtype
        return out.reshape(B, Lq, num_heads * hd)

    @staticmethod
    def backward(ctx, dout):
        q, out, lse = ctx.saved_tensors
        B, Lq, H, hd = ctx.dims
        d = cast_bf16(dout.reshape(B * Lq, H * hd))
        dqkv = attn_bwd(q, out, d, lse, B, Lq, H, hd)
        if ctx.in_dtype == F32:
            dqkv = cast_f32(dqkv)
        return dqkv, None


def attention(qkv, num_heads):
    return AttentionFn.apply(qkv, num_heads)


# ((data pointer, numel), fp32 stream gradient, its bf16 copy) handed from one block's backward to the next one's. Matching is
# by storage, not by tensor identity: between two decoder layers the gradient passes through view/reshape backward nodes, which
# wrap the same memory in a new tensor. The entry holds a strong reference, so the memory cannot be recycled while it lives and
# autograd cannot accumulate into it in place (its use count is > 1).
_GRAD_SHADOW = [None]


def _shadow_key(t):
    return (t.data_ptr(), t.numel())

# Direct gradient accumulation (opt-in). For a parameter marked by `enable_direct_grad(params, hook)` that already owns a contiguous
# fp32 .grad (e.g. a view into a GradientBucketer bucket) the fused block, LinearFn and LayerNormFn accumulate into it in place (TMA reduce-add / red.add)
# and returns None to autograd: no zero-filled temporary, no extra "grad += new" pass. Because autograd then never sees that
# gradient, tensor hooks do not fire and `torch.autograd.grad(...)` would get None — hence opt-in, by whoever owns the .grad
# buffers and calls `.backward()`. `hook(param)` (per parameter, not one global) replaces the post-accumulate hook autograd
# would have fired. GRAD_READY_HOOK is an additional process-wide observer (tests).
GRAD_READY_HOOK = None


def enable_direct_grad(params, hook=None) -> None:
    for p in params:
        p._vmae_direct_grad = True
        p._vmae_grad_hook = hook


def disable_direct_grad(params) -> None:
    for p in params:
        p._vmae_direct_grad = False
        p._vmae_grad_hook = None


def _direct_grad(p):
    if not getattr(p, "_vmae_direct_grad", False):
        return None
    g = getattr(p, "grad", None)
    if g is not None and g.dtype == F32 and g.is_contiguous() and g.shape == p.shape and g.is_cuda:
        return g
    return None


def _grad_done(p):
    h = getattr(p, "_vmae_grad_hook", None)
    if h is not None:
        h(p)
    if GRAD_READY_HOOK is not None:
        GRAD_READY_HOOK(p)


_SHADOW_CB = [False]


def _clear_shadow():
    _GRAD_SHADOW[0] = None
    _SHADOW_CB[0] = False


def _publish_shadow(dx, dx_b):
    """Hand (dx, bf16 copy) to the next block's backward; dropped when this backward pass ends (no reference outlives it)."""
    _GRAD_SHADOW[0] = (_shadow_key(dx), dx, dx_b)
    if not _SHADOW_CB[0]:
        _SHADOW_CB[0] = True
        torch.autograd.Variable._execution_engine.queue_callback(_clear_shadow)


class TransformerBlockFn(torch.autograd.Function):
    """One pre-LN transformer block, forward and backward as a fixed schedule of fused kernels:

        x1 = x  + proj(attn(qkv(LN1(x))))          x2 = x1 + fc2(GELU(fc1(LN2(x1))))

    It is the arithmetic of both TinyViTBlock (tiny_vit.py:108-130) and nn.TransformerEncoderLayer(norm_first=True,
    activation="gelu") (mae_vit_adapter.py:40-48) with dropout / drop-path inactive. Residual stream fp32 (as the
    reference's decoder under autocast), GEMM/attention operands bf16, fp32 accumulation.
    Fusions: bias / bias+GELU(+gelu' saved for backward) / bias+residual in GEMM epilogues, gelu' multiply in the fc2-dgrad epilogue,
    residual-gradient add + bf16 copy in the LayerNorm backward.
    """

    @staticmethod
    def forward(ctx, x, ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b, num_heads, B, Lq,
                p_drop=0.0, seeds=None):
        _need_cuda(x)
        D = x.shape[-1]
        assert x.dtype == F32 and x.is_contiguous() and x.shape[0] == B * Lq
        hd = D // num_heads
        params = (ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b)
        f = lambda t: t.float().contiguous()
        ln1_w, ln1_b, ln2_w, ln2_b = f(ln1_w), f(ln1_b), f(ln2_w), f(ln2_b)
        qkv_b, proj_b, fc1_b, fc2_b = f(qkv_b), f(proj_b), f(fc1_b), f(fc2_b)
        wq, wp, w1, w2 = cast_weight(qkv_w), cast_weight(proj_w), cast_weight(fc1_w), cast_weight(fc2_w)
        drop = p_drop > 0.0
        if drop:   # nn.TransformerEncoderLayer's four dropout sites: attention probabilities, dropout1, dropout (after GELU), dropout2
            s_attn, s_d1, s_act, s_d2 = seeds
        ln1, mean1, rstd1 = layernorm_fwd(x, ln1_w, ln1_b)
        qkv = linear_fwd(ln1, wq, qkv_b)
        o, lse = attn_fwd(qkv, B, Lq, num_heads, hd, p_drop, s_attn if drop else 0)
        if drop:
            x1 = dropout_apply(linear_fwd(o, wp, proj_b), p_drop, s_d1, add=x, out_dtype=F32)
        else:
            x1 = linear_fwd(o, wp, proj_b, residual=x)
        ln2, mean2, rstd2 = layernorm_fwd(x1, ln2_w, ln2_b)
        # gelu'(fc1 output), saved for backward: one byte per element (quantisation step 0.0049) unless the dropout path needs to
        # rescale it in bf16 — the fc1 GEMM is bound by HBM write bandwidth and this is a quarter of what it writes
        h_pre = torch.empty((x.shape[0], w1.shape[0]), device=x.device, dtype=torch.uint8 if (GELU_PRIME_Q8 and not drop and w1.shape[0] % 16 == 0) else BF16)
        act = linear_fwd(ln2, w1, fc1_b, act=L.ACT_GELU, aux_out=h_pre)
        if drop:   # same mask on the activation and on the saved gelu' (so backward needs no extra pass)
            dropout_apply(act, p_drop, s_act, out=act)
            dropout_apply(h_pre, p_drop, s_act, out=h_pre)
            x2 = dropout_apply(linear_fwd(act, w2, fc2_b), p_drop, s_d2, add=x1, out_dtype=F32)
        else:
            x2 = linear_fwd(act, w2, fc2_b, residual=x1)
        ctx.save_for_backward(x, ln1_w, mean1, rstd1, ln1, qkv, o, lse, x1, ln2_w, mean2, rstd2, ln2, h_pre, act, wq, wp, w1, w2)
        ctx.dims = (B, Lq, num_heads, hd)
        ctx.params = params
        ctx.drop = (p_drop, tuple(seeds)) if drop else None
        return x2

    @staticmethod
    def backward(ctx, dx2):
        (x, ln1_w, mean1, rstd1, ln1, qkv, o, lse, x1, ln2_w, mean2, rstd2, ln2, h_pre, act, wq, wp, w1, w2) = ctx.saved_tensors
        B, Lq, H, hd = ctx.dims
        dx2 = dx2.contiguous()
        p_drop, (s_attn, s_d1, s_act, s_d2) = ctx.drop if ctx.drop is not None else (0.0, (0, 0, 0, 0))
        # the LayerNorm backward of the block that produced dx2 already wrote a bf16 copy of it (same kernel, no extra read)
        sh = _GRAD_SHADOW[0]
        if p_drop > 0.0:
            dx2_b = dropout_apply(dx2, p_drop, s_d2, out_dtype=BF16)      # gradient entering fc2 = dropout2 mask * dx2 / (1-p)
        else:
            hit = sh is not None and dx2.dtype == F32 and sh[0] == _shadow_key(dx2)
            dx2_b = sh[2].view(dx2.shape) if hit else cast_bf16(dx2)
        _GRAD_SHADOW[0] = None
        P = ctx.params  # ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b
        tgt = [(_direct_grad(p) if ctx.needs_input_grad[i + 1] else None) for i, p in enumerate(P)]
        (t_g1, t_b1, t_wq, t_bq, t_wp, t_bp, t_g2, t_b2, t_w1, t_bb1, t_w2, t_bb2) = tgt
        # ---- MLP branch
        dh = linear_dgrad(dx2_b, w2, aux_in=h_pre)          # d(fc1 out) = (dx2 W2) * gelu'(fc1 out); h_pre holds gelu'
        dw2, db2 = linear_wgrad_bias(dx2_b, act, t_w2, t_bb2)
        dln2 = linear_dgrad(dh, w1)
        dw1, db1 = linear_wgrad_bias(dh, ln2, t_w1, t_bb1)
        dx1, dx1_b, dg2, dbeta2 = layernorm_bwd(dln2, x1, ln2_w, mean2, rstd2, dx_resid=dx2, dgamma=t_g2, dbeta=t_b2, want_bf16=True)
        # ---- attention branch
        if p_drop > 0.0:
            dx1_b = dropout_apply(dx1, p_drop, s_d1, out_dtype=BF16)      # gradient entering out_proj = dropout1 mask * dx1 / (1-p)
        do = linear_dgrad(dx1_b, wp)
        dwp, dbp = linear_wgrad_bias(dx1_b, o, t_wp, t_bp)
        dqkv = attn_bwd(qkv, o, do, lse, B, Lq, H, hd, p_drop, s_attn)
        dln1 = linear_dgrad(dqkv, wq)
        dwq, dbq = linear_wgrad_bias(dqkv, ln1, t_wq, t_bq)
        dx, dx_b, dg1, dbeta1 = layernorm_bwd(dln1, x, ln1_w, mean1, rstd1, dx_resid=dx1, dgamma=t_g1, dbeta=t_b1, want_bf16=True)
        _publish_shadow(dx, dx_b)
        grads = [dg1, dbeta1, dwq, dbq, dwp, dbp, dg2, dbeta2, dw1, db1, dw2, db2]
        out = []
        # reverse order ~ the order in which the gradients became final (fc2 first), so buckets close as early as possible
        for i in reversed(range(12)):
            if tgt[i] is not None:
                _grad_done(P[i])
        for i in range(12):
            out.append(None if tgt[i] is not None else grads[i])
        return (dx, *out, None, None, None, None, None)


def transformer_block(x, ln1, qkv_w, qkv_b, proj, ln2, fc1, fc2, num_heads, B, Lq, p_drop=0.0, seeds=None):
    """p_drop > 0: dropout at the four sites of nn.TransformerEncoderLayer (training mode); seeds = 4 integers (new_dropout_seeds)."""
    if p_drop > 0.0 and seeds is None:
        seeds = new_dropout_seeds(4)
    return TransformerBlockFn.apply(x, ln1.weight, ln1.bias, qkv_w, qkv_b, proj.weight, proj.bias, ln2.weight, ln2.bias,
                                    fc1.weight, fc1.bias, fc2.weight, fc2.bias, num_heads, B, Lq, float(p_drop), seeds)


class LinearResidualF32Fn(torch.autograd.Function):
    """fp32 out = x W^T + b + residual (fp32). Used for the tubelet embedding (+pos rows) in the ViT encoder."""

    @staticmethod
    def forward(ctx, x_bf16, weight, bias, residual):
        wb = cast_weight(weight).reshape(weight.shape[0], -1)
        y = linear_fwd(x_bf16, wb, bias.float().contiguous() if bias is not None else None, residual=residual.contiguous())
        ctx.save_for_backward(x_bf16, wb)
        ctx.w_shape = weight.shape
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wb = ctx.saved_tensors
        dy = dy.contiguous()
        dyb = cast_bf16(dy)
        dx = linear_dgrad(dyb, wb) if ctx.needs_input_grad[0] else None
        dw = linear_wgrad(dyb, x).reshape(ctx.w_shape) if ctx.needs_input_grad[1] else None
        db = colsum(dyb) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, dy if ctx.needs_input_grad[3] else None


# ---- MAE-specific -----------------------------------------------------------------------------------------------
def mask_to_indices(mask_frame_u8, n_vis):
    B, Lp = mask_frame_u8.shape
    vis = torch.empty((B, n_vis), device=mask_frame_u8.device, dtype=torch.int32)
    L.check(L.load().vmae_mask_to_indices(_p(mask_frame_u8), _p(vis), B, Lp, n_vis, _stream()), "vmae_mask_to_indices")
    _count()
    return vis


class TubeletGatherFn(torch.autograd.Function):
    """visible-tubelet im2col + encoder positional rows. Non-differentiable w.r.t. the clip; the positional tables get
    their gradient from the rows' gradient (scatter-add)."""

    @staticmethod
    def forward(ctx, clip, vis_idx, tpos, spos, pt, ps):
        B, Cc, T, H, W = clip.shape
        n_vis = vis_idx.shape[1]
        D = tpos.shape[-1]
        Tp = T // pt
        rows = B * Tp * n_vis
        patches = torch.empty((rows, Cc * pt * ps * ps), device=clip.device, dtype=BF16)
        pos_rows = torch.empty((rows, D), device=clip.device, dtype=F32)
        tp2, sp2 = tpos.reshape(Tp, D).float().contiguous(), spos.reshape(-1, D).float().contiguous()
        L.check(L.load().vmae_gather_tubelets(_p(clip), _p(vis_idx), _p(patches), _p(tp2), _p(sp2), _p(pos_rows), B, T, H, W,
                                              pt, ps, n_vis, D, _stream()), "vmae_gather_tubelets")
        _count()
        ctx.save_for_backward(vis_idx)
        ctx.dims = (B, Tp, n_vis, D, sp2.shape[0])
        ctx.shapes = (tpos.shape, spos.shape)
        ctx.mark_non_differentiable(patches)
        return patches, pos_rows

    @staticmethod
    def backward(ctx, _dpatches, dpos_rows):
        (vis_idx,) = ctx.saved_tensors
        B, Tp, n_vis, D, Lp = ctx.dims
        dt = torch.zeros((Tp, D), device=dpos_rows.device, dtype=F32)
        ds = torch.zeros((Lp, D), device=dpos_rows.device, dtype=F32)
        L.check(L.load().vmae_pos_rows_bwd(_p(dpos_rows.contiguous()), _p(vis_idx), _p(dt), _p(ds), B, Tp, n_vis, D, _stream()),
                "vmae_pos_rows_bwd")
        _count()
        return None, None, dt.reshape(ctx.shapes[0]), ds.reshape(ctx.shapes[1]), None, None


class TubeletEmbedFn(torch.autograd.Function):
    """Tubelet embedding of the visible tubes as ONE implicit-GEMM kernel (K6): Conv3d(3, E, (pt,16,16), stride = kernel) evaluated
    where the tube mask is 0 + bias + factorised positional rows, output fp32 [B*T'*n_vis, E] in gathered order. The clip is read
    in place through 5-D TMA boxes (no patch matrix); backward re-gathers the same boxes for the weight gradient and reduces the
    bias / positional gradients from dy. The clip gets no gradient (it is data)."""

    @staticmethod
    def forward(ctx, clip, vis_idx, weight, bias, tpos, spos, pt, ps):
        _need_cuda(clip, vis_idx, weight)
        B, Cc, T, H, W = clip.shape
        assert Cc == 3 and clip.dtype == F32 and clip.is_contiguous() and vis_idx.dtype == torch.int32 and vis_idx.is_contiguous()
        E = weight.shape[0]
        n_vis = vis_idx.shape[1]
        Tp = T // pt
        wb = cast_weight(weight).reshape(E, -1)
        rows = B * Tp * n_vis
        out = torch.empty((rows, E), device=clip.device, dtype=F32)
        tp2, sp2 = tpos.reshape(Tp, E).float().contiguous(), spos.reshape(-1, E).float().contiguous()
        bf = bias.float().contiguous() if bias is not None else None
        L.check(_run("gemm_fwd", 2.0 * rows * E * wb.shape[1], lambda: L.load().vmae_tubelet_embed_fwd(
            _p(clip), _p(vis_idx), _p(wb), _p(bf), _p(tp2), _p(sp2), _p(out), B, T, H, W, pt, ps, n_vis, E, _stream())), "vmae_tubelet_embed_fwd")
        _count()
        ctx.save_for_backward(clip, vis_idx)
        ctx.geom = (B, T, H, W, pt, ps, n_vis, E, Tp, sp2.shape[0])
        ctx.shapes = (weight.shape, tpos.shape, spos.shape)
        ctx.has_bias = bias is not None
        ctx.params = (weight, bias, tpos, spos)
        return out

    @staticmethod
    def backward(ctx, dy):
        clip, vis_idx = ctx.saved_tensors
        B, T, H, W, pt, ps, n_vis, E, Tp, Lp = ctx.geom
        weight, bias, tpos, spos = ctx.params
        dy = dy.contiguous()
        dyb = cast_bf16(dy)
        dw = db = dt = ds = None
        # every kernel below only ADDS into its gradient buffer, so an opted-in .grad (enable_direct_grad) can take the place of the
        # zero-filled temporary
        if ctx.needs_input_grad[2]:
            t_w = _direct_grad(weight)
            dw = t_w.view(E, -1) if t_w is not None else torch.zeros((E, 3 * pt * ps * ps), device=dy.device, dtype=F32)
            L.check(_run("gemm_wgrad", 2.0 * dy.shape[0] * E * dw.shape[1], lambda: L.load().vmae_tubelet_embed_wgrad(
                _p(clip), _p(vis_idx), _p(dyb), _p(dw), B, T, H, W, pt, ps, n_vis, E, _stream())), "vmae_tubelet_embed_wgrad")
            _count()
            if t_w is not None:
                dw = None
                _grad_done(weight)
            else:
                dw = dw.reshape(ctx.shapes[0])
        if ctx.has_bias and ctx.needs_input_grad[3]:
            t_b = _direct_grad(bias)
            db = colsum(dyb, t_b)
            if t_b is not None:
                db = None
                _grad_done(bias)
        if ctx.needs_input_grad[4] or ctx.needs_input_grad[5]:
            t_t, t_s = _direct_grad(tpos), _direct_grad(spos)
            direct = t_t is not None and t_s is not None and ctx.needs_input_grad[4] and ctx.needs_input_grad[5]
            dt = t_t.view(Tp, E) if direct else torch.zeros((Tp, E), device=dy.device, dtype=F32)
            ds = t_s.view(Lp, E) if direct else torch.zeros((Lp, E), device=dy.device, dtype=F32)
            L.check(L.load().vmae_pos_rows_bwd(_p(dy), _p(vis_idx), _p(dt), _p(ds), B, Tp, n_vis, E, _stream()), "vmae_pos_rows_bwd")
            _count()
            if direct:
                dt = ds = None
                _grad_done(spos)
                _grad_done(tpos)
            else:
                dt, ds = dt.reshape(ctx.shapes[1]), ds.reshape(ctx.shapes[2])
        return None, None, dw, db, dt, ds, None, None


def tubelet_embed_supported(ps, E):
    return ps == 16 and E % 128 == 0 and 128 <= E <= 384


class DecoderInputFn(torch.autograd.Function):
    """x_dec = visible ? z + (tpos + spos) : mask_token  (mae_vit_adapter.py:97-107). z bf16 rows; fp32 out [B,Tp*L,D]."""

    @staticmethod
    def forward(ctx, z, vis_idx, mask_u8, tpos, spos, mask_token, B, Tp, Lp):
        D = z.shape[-1]
        n_vis = vis_idx.shape[1] if vis_idx is not None else 0
        x = torch.empty((B, Tp * Lp, D), device=z.device, dtype=F32)
        z = cast_bf16(z)
        tp2, sp2, mt = tpos.reshape(Tp, D).float().contiguous(), spos.reshape(Lp, D).float().contiguous(), mask_token.reshape(D).float().contiguous()
        L.check(L.load().vmae_decoder_input_fwd(_p(z), _p(vis_idx), _p(mask_u8), _p(tp2), _p(sp2), _p(mt), _p(x), B, Tp, Lp, n_vis,
                                                D, _stream()), "vmae_decoder_input_fwd")
        _count()
        ctx.save_for_backward(vis_idx, mask_u8)
        ctx.dims = (B, Tp, Lp, n_vis, D, z.shape)
        ctx.shapes = (tpos.shape, spos.shape, mask_token.shape)
        ctx.params = (tpos, spos, mask_token)
        return x

    @staticmethod
    def backward(ctx, dx):
        vis_idx, mask_u8 = ctx.saved_tensors
        B, Tp, Lp, n_vis, D, zshape = ctx.dims
        tpos, spos, mask_token = ctx.params
        dx = dx.contiguous()
        dz = torch.empty(zshape, device=dx.device, dtype=BF16)
        # the kernel only ADDS into the three parameter gradients: opted-in .grad buffers replace the zero-filled temporaries
        tg = [_direct_grad(p) for p in (tpos, spos, mask_token)]
        direct = all(t is not None for t in tg) and all(ctx.needs_input_grad[3:6])
        dt = tg[0].view(Tp, D) if direct else torch.zeros((Tp, D), device=dx.device, dtype=F32)
        ds = tg[1].view(Lp, D) if direct else torch.zeros((Lp, D), device=dx.device, dtype=F32)
        dm = tg[2].view(D) if direct else torch.zeros((D,), device=dx.device, dtype=F32)
        L.check(L.load().vmae_decoder_input_bwd(_p(dx), _p(vis_idx), _p(mask_u8), _p(dz), _p(dt), _p(ds), _p(dm), B, Tp, Lp, n_vis,
                                                D, _stream()), "vmae_decoder_input_bwd")
        _count()
        if direct:
            for p in (mask_token, spos, tpos):
                _grad_done(p)
            return dz, None, None, None, None, None, None, None, None
        return dz, None, None, dt.reshape(ctx.shapes[0]), ds.reshape(ctx.shapes[1]), dm.reshape(ctx.shapes[2]), None, None, None


class MaskedMSEFn(torch.autograd.Function):
    """Fused patchify + per-patch normalise + masked MSE (train_ssl_mae.py:26-31,72-84). Returns the scalar loss."""

    @staticmethod
    def forward(ctx, pred, clip, mask_u8, pt, ps, norm_pix):
        _need_cuda(pred, clip, mask_u8)
        B, Cc, T, H, W = clip.shape
        pred_c = pred.contiguous()
        clip = clip.contiguous()
        assert clip.dtype == F32 and Cc == 3
        D = pt * ps * ps * 3
        acc = torch.zeros((2,), device=pred.device, dtype=F32)
        diff = torch.empty(pred_c.shape, device=pred.device, dtype=BF16)
        lib = L.load()
        L.check(_run("masked_mse_fwd", float(clip.numel()) * 4 + float(pred_c.numel()) * (pred_c.element_size() + 2), lambda: lib.vmae_masked_mse_fwd(
            _p(clip), _p(pred_c), _dt(pred_c), _p(mask_u8), _p(acc), _p(diff), B, T, H, W, pt, ps, norm_pix, _stream())), "vmae_masked_mse_fwd")
        loss = torch.empty((1,), device=pred.device, dtype=F32)
        scale = torch.empty((1,), device=pred.device, dtype=F32)
        L.check(lib.vmae_masked_mse_finalize(_p(acc), _p(loss), _p(scale), D, _stream()), "vmae_masked_mse_finalize")
        _count(2)
        ctx.save_for_backward(diff, scale)
        ctx.pred_dtype = pred.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        diff, scale = ctx.saved_tensors
        g = gout.reshape(1).float().contiguous()
        dpred = torch.empty_like(diff)
        L.check(L.load().vmae_masked_mse_bwd(_p(diff), _p(scale), _p(g), _p(dpred), diff.numel(), _stream()), "vmae_masked_mse_bwd")
        _count()
        if ctx.pred_dtype == F32:
            dpred = cast_f32(dpred)
        return dpred, None, None, None, None, None


def masked_mse(pred, clip, mask, pt, ps, norm_pix=1):
    m = mask.reshape(mask.shape[0], -1)
    m8 = m.to(torch.uint8) if m.dtype != torch.uint8 else m
    return MaskedMSEFn.apply(pred, clip, m8.contiguous(), pt, ps, norm_pix)


def frames_to_clip(frames_u8, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), flip_channels=True, out=None):
    """Decoded uint8 frames [B, T, H, W, 3] on the GPU -> normalised fp32 clip [B, 3, T, H, W] (ToTensor + Normalize + the loader's
    channel reversal and permute, mae_loader.py:63-77): the host uploads a quarter of the bytes of a float clip."""
    import ctypes

    _need_cuda(frames_u8)
    assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 5 and frames_u8.shape[-1] == 3 and frames_u8.is_contiguous()
    B, T, H, W, _ = frames_u8.shape
    if out is None:
        out = torch.empty((B, 3, T, H, W), device=frames_u8.device, dtype=F32)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    L.check(L.load().vmae_frames_to_clip(_p(frames_u8), _p(out), B, T, H, W, ctypes.cast(m3, ctypes.c_void_p), ctypes.cast(s3, ctypes.c_void_p),
                                         1 if flip_channels else 0, _stream()), "vmae_frames_to_clip")
    _count()
    return out


class RowLossFn(torch.autograd.Function):
    """Stack-2 latent losses on gathered rows (mae/losses.py:19-63): mode 0 = L2 (optionally with per-row biased-variance target
    normalisation), 1 = L1, 2 = cosine. pred [rows, cols] bf16/fp32 gets the gradient; target is detached as in the reference."""

    @staticmethod
    def forward(ctx, pred, target, mode, normalize, eps):
        _need_cuda(pred, target)
        assert pred.shape == target.shape and pred.dim() == 2
        pred = pred.contiguous()
        target = target.detach().contiguous()
        if pred.dtype not in (BF16, F32):
            pred = pred.float()
        if target.dtype not in (BF16, F32):
            target = target.float()
        rows, cols = pred.shape
        lib = L.load()
        partials = torch.empty((lib.vmae_row_loss_grid(rows),), device=pred.device, dtype=F32)
        loss = torch.empty((), device=pred.device, dtype=F32)
        L.check(lib.vmae_row_loss_fwd(_p(pred), _dt(pred), _p(target), _dt(target), rows, cols, int(mode), int(bool(normalize)), float(eps),
                                      _p(partials), _p(loss), _stream()), "vmae_row_loss_fwd")
        _count(2)
        ctx.save_for_backward(pred, target)
        ctx.cfg = (int(mode), int(bool(normalize)), float(eps))
        return loss

    @staticmethod
    def backward(ctx, gout):
        pred, target = ctx.saved_tensors
        mode, normalize, eps = ctx.cfg
        rows, cols = pred.shape
        dpred = torch.empty_like(pred)
        g = gout.detach().float().contiguous()
        L.check(L.load().vmae_row_loss_bwd(_p(pred), _dt(pred), _p(target), _dt(target), rows, cols, mode, normalize, eps, _p(g), _p(dpred),
                                           _stream()), "vmae_row_loss_bwd")
        _count()
        return dpred, None, None, None, None


def row_loss(pred, target, mode, normalize=True, eps=1e-6):
    return RowLossFn.apply(pred, target, mode, normalize, eps)


# ---- stack 2 (tinyvit_mae.py:104-114): mask-token substitution and the boolean row gathers, sync-free ---------------------
def mask_row_indices(mask_u8, n_sel, want=1, flag=None):
    """mask uint8 [B, N] -> int32 [B * n_sel] global row numbers (b*N + n) of the entries with (mask != 0) == want, ascending:
    the row order of `x[mask]`. `flag` (int32 scalar on the device, zeroed by the caller) records a per-clip count mismatch."""
    B, N = mask_u8.shape
    idx = torch.empty((B * n_sel,), device=mask_u8.device, dtype=torch.int32)
    L.check(L.load().vmae_mask_row_indices(_p(mask_u8), _p(idx), _p(flag), B, N, int(n_sel), int(want), _stream()), "vmae_mask_row_indices")
    _count()
    return idx


class MaskTokenBlendFn(torch.autograd.Function):
    """torch.where(mask[..., None], mask_token, tok) (tinyvit_mae.py:107-108). tok [B, N, D] bf16/fp32; mask uint8 [B, N]."""

    @staticmethod
    def forward(ctx, tok, mask_u8, mask_token):
        _need_cuda(tok, mask_u8, mask_token)
        tok = tok.contiguous()
        if tok.dtype not in (BF16, F32):
            tok = tok.float()
        D = tok.shape[-1]
        rows = tok.numel() // D
        out = torch.empty_like(tok)
        mt = mask_token.reshape(D).float().contiguous()
        L.check(L.load().vmae_mask_token_blend_fwd(_p(tok), _dt(tok), _p(mask_u8), _p(mt), _p(out), rows, D, _stream()), "vmae_mask_token_blend_fwd")
        _count()
        ctx.save_for_backward(mask_u8)
        ctx.mt_shape = mask_token.shape
        ctx.mt_dtype = mask_token.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask_u8,) = ctx.saved_tensors
        dout = dout.contiguous()
        D = dout.shape[-1]
        rows = dout.numel() // D
        dtok = torch.empty_like(dout) if ctx.needs_input_grad[0] else None
        dmt = torch.zeros((D,), device=dout.device, dtype=F32) if ctx.needs_input_grad[2] else None
        L.check(L.load().vmae_mask_token_blend_bwd(_p(dout), _dt(dout), _p(mask_u8), _p(dtok), _p(dmt), rows, D, _stream()), "vmae_mask_token_blend_bwd")
        _count()
        return dtok, None, (dmt.reshape(ctx.mt_shape).to(ctx.mt_dtype) if dmt is not None else None)


class GatherRowsFn(torch.autograd.Function):
    """out[i, :] = src2d[idx[i], :]; backward scatters into zeros (the rows are unique)."""

    @staticmethod
    def forward(ctx, src, idx):
        _need_cuda(src, idx)
        D = src.shape[-1]
        src2 = src.reshape(-1, D).contiguous()
        if src2.dtype not in (BF16, F32):
            src2 = src2.float()
        out = torch.empty((idx.numel(), D), device=src.device, dtype=src2.dtype)
        L.check(L.load().vmae_gather_rows(_p(src2), _dt(src2), _p(idx), _p(out), idx.numel(), D, _stream()), "vmae_gather_rows")
        _count()
        ctx.save_for_backward(idx)
        ctx.src_shape = src.shape
        ctx.n_rows = src2.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dout = dout.contiguous()
        D = dout.shape[-1]
        dsrc = torch.empty((ctx.n_rows, D), device=dout.device, dtype=dout.dtype)
        L.check(L.load().vmae_scatter_rows(_p(dout), _dt(dout), _p(idx), _p(dsrc), idx.numel(), ctx.n_rows, D, _stream()), "vmae_scatter_rows")
        _count(2)
        return dsrc.reshape(ctx.src_shape), None


class AdaptiveAvgPoolFn(torch.autograd.Function):
    """F.adaptive_avg_pool2d(f, (s, s)).flatten(2).transpose(1, 2) (tinyvit_backbone.py:74-81): bf16 NHWC [N,H,W,C] -> fp32 [N, s*s, C]."""

    @staticmethod
    def forward(ctx, x_nhwc, s):
        _need_cuda(x_nhwc)
        N, H, W, Cc = x_nhwc.shape
        assert x_nhwc.dtype == BF16 and x_nhwc.is_contiguous()
        out = torch.empty((N, s * s, Cc), device=x_nhwc.device, dtype=F32)
        L.check(L.load().vmae_adaptive_avgpool_fwd(_p(x_nhwc), _p(out), N, H, W, Cc, int(s), _stream()), "vmae_adaptive_avgpool_fwd")
        _count()
        ctx.geom = (N, H, W, Cc, int(s))
        return out

    @staticmethod
    def backward(ctx, dout):
        N, H, W, Cc, s = ctx.geom
        dx = torch.empty((N, H, W, Cc), device=dout.device, dtype=BF16)
        L.check(L.load().vmae_adaptive_avgpool_bwd(_p(dout.float().contiguous()), _p(dx), N, H, W, Cc, s, _stream()), "vmae_adaptive_avgpool_bwd")
        _count()
        return dx, None


class MeanPoolNoiseFn(torch.autograd.Function):
    """z[b] = mean over tokens of x[b] (+ sigma * eps in the same kernel): bf16 [B, n, D] -> fp32 [B, D]. Gradient: dz / n broadcast."""

    @staticmethod
    def forward(ctx, x, sigma, seed, offset, eps_out):
        _need_cuda(x)
        B, n, D = x.shape
        xb = cast_bf16(x)
        out = torch.empty((B, D), device=x.device, dtype=F32)
        if eps_out is not None:
            assert eps_out.dtype == F32 and eps_out.is_contiguous() and eps_out.numel() == B * D and eps_out.is_cuda
        L.check(L.load().vmae_meanpool_noise(_p(xb), _p(out), _p(eps_out), B, n, D, float(max(sigma, 0.0)), int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset),
                                             _stream()), "vmae_meanpool_noise")
        _count()
        ctx.shape = (B, n, D)
        ctx.in_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, dz):
        B, n, D = ctx.shape
        dx = torch.empty((B, n, D), device=dz.device, dtype=BF16)
        L.check(L.load().vmae_se_scale(None, None, _p(dz.float().contiguous()), _p(dx), B, n, D, 1.0 / n, _stream()), "vmae_se_scale")
        _count()
        return (cast_f32(dx) if ctx.in_dtype == F32 else dx), None, None, None, None


def meanpool_with_feature_noise(x, sigma=0.0, seed=None, eps_out=None):
    if seed is None:
        seed = torch.initial_seed()
    off = 0
    if sigma > 0:
        off = _noise_offset[0]
        _noise_offset[0] += (x.shape[0] * x.shape[2] + 3) // 4
    return MeanPoolNoiseFn.apply(x, float(sigma), int(seed), off, eps_out)


class SoftmaxCrossEntropyFn(torch.autograd.Function):
    """nn.CrossEntropyLoss (mean) on logits [rows, C] bf16/fp32 and int64 labels, fused: one pass for lse + loss, one for dlogits."""

    @staticmethod
    def forward(ctx, logits, labels):
        _need_cuda(logits, labels)
        assert logits.dim() == 2 and labels.shape == (logits.shape[0],)
        lg = logits.contiguous()
        if lg.dtype not in (BF16, F32):
            lg = lg.float()
        lb = labels.to(torch.int64).contiguous()
        rows, Cn = lg.shape
        lse = torch.empty((rows,), device=lg.device, dtype=F32)
        row_loss = torch.empty((rows,), device=lg.device, dtype=F32)
        loss = torch.empty((), device=lg.device, dtype=F32)
        L.check(L.load().vmae_softmax_ce_fwd(_p(lg), _dt(lg), _p(lb), _p(lse), _p(row_loss), _p(loss), rows, Cn, _stream()), "vmae_softmax_ce_fwd")
        _count(2)
        ctx.save_for_backward(lg, lb, lse)
        return loss

    @staticmethod
    def backward(ctx, gout):
        lg, lb, lse = ctx.saved_tensors
        rows, Cn = lg.shape
        dl = torch.empty_like(lg)
        g = gout.detach().reshape(1).float().contiguous()
        L.check(L.load().vmae_softmax_ce_bwd(_p(lg), _dt(lg), _p(lb), _p(lse), _p(g), _p(dl), rows, Cn, _stream()), "vmae_softmax_ce_bwd")
        _count()
        return dl, None


def softmax_cross_entropy(logits, labels):
    return SoftmaxCrossEntropyFn.apply(logits, labels)


class GaussianNoiseFn(torch.autograd.Function):
    """z' = z + sigma * eps, eps ~ N(0, I) from Philox (privacy/feature_noise.py:4-7); gradient is the identity."""

    @staticmethod
    def forward(ctx, z, sigma, seed, offset, eps_out=None):
        """eps_out (optional fp32 tensor of z's shape) receives the unit normals that were drawn: out == z + sigma * eps_out."""
        _need_cuda(z)
        zc = z.contiguous()
        out = torch.empty_like(zc)
        if eps_out is not None:
            assert eps_out.dtype == F32 and eps_out.is_contiguous() and eps_out.numel() == zc.numel() and eps_out.is_cuda
        L.check(L.load().vmae_add_gaussian_noise(_p(zc), _p(out), _dt(zc), _p(eps_out), zc.numel(), float(sigma), int(seed), int(offset),
                                                 _stream()), "vmae_add_gaussian_noise")
        _count()
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None, None


def scale_f32(x, alpha, out=None, accumulate=False):
    _need_cuda(x)
    xc = x.contiguous()
    if out is None:
        out = torch.empty_like(xc)
    L.check(L.load().vmae_scale_f32(_p(xc), _p(out), float(alpha), 1 if accumulate else 0, xc.numel(), _stream()), "vmae_scale_f32")
    _count()
    return out


# ======================================================================================================================
# K7 — TinyViT convolution kit (channels-last bf16 feature maps; tiny_vit.py:12-72)
# ======================================================================================================================
def as_nhwc_bf16(x):
    """NCHW tensor (any memory format / dtype) -> contiguous bf16 [N, H, W, C]; free for a channels-last bf16 input."""
    t = x.permute(0, 2, 3, 1)
    if t.dtype != BF16:
        t = cast_bf16_autograd(t.contiguous())
    return t if t.is_contiguous() else t.contiguous()


class BatchNormActFn(torch.autograd.Function):
    """Train-mode BatchNorm2d over rows [M, C] (+ optional exact GELU, + optional residual), bf16 in / bf16 out, fp32
    statistics; updates running_mean / running_var in place like nn.BatchNorm2d (momentum, unbiased variance)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, eps, momentum, act):
        _need_cuda(x, gamma)
        M, Cc = x.shape
        g, b = gamma.float().contiguous(), beta.float().contiguous()
        y = torch.empty_like(x)
        mean = torch.empty((Cc,), device=x.device, dtype=F32)
        rstd = torch.empty((Cc,), device=x.device, dtype=F32)
        ws = torch.empty((2 * Cc,), device=x.device, dtype=F32)
        nbytes = float(M) * Cc * 2 * (3 + (1 if residual is not None else 0))
        L.check(_run("batchnorm_fwd", nbytes, lambda: L.load().vmae_bn_train_fwd(
            _p(x), _p(g), _p(b), _p(residual), _p(y), _p(mean), _p(rstd), _p(running_mean), _p(running_var), _p(ws), M, Cc, float(eps),
            float(momentum), int(act), _stream())), "vmae_bn_train_fwd")
        _count(3)
        ctx.save_for_backward(x, g, b, mean, rstd)
        ctx.act = int(act)
        ctx.has_res = residual is not None
        ctx.mark_non_differentiable(mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, b, mean, rstd = ctx.saved_tensors
        M, Cc = x.shape
        dy = cast_bf16(dy.contiguous())
        dx = torch.empty_like(x)
        dg = torch.zeros((Cc,), device=x.device, dtype=F32)
        db = torch.zeros((Cc,), device=x.device, dtype=F32)
        ws = torch.empty((2 * Cc,), device=x.device, dtype=F32)
        L.check(_run("batchnorm_bwd", float(M) * Cc * 2 * 5, lambda: L.load().vmae_bn_train_bwd(
            _p(x), _p(dy), _p(g), _p(b), _p(mean), _p(rstd), _p(dx), _p(dg), _p(db), _p(ws), M, Cc, ctx.act, _stream())), "vmae_bn_train_bwd")
        _count(2)
        return dx, dg, db, None, None, (dy if ctx.has_res else None), None, None, None


class BatchNormEvalFn(torch.autograd.Function):
    """Eval-mode BatchNorm2d over rows [M, C] with the running statistics (+ optional exact GELU, + optional residual):
    `vmae_bn_apply`. Inference only — the backward of the eval form is not part of the training path and raises."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, eps, act):
        _need_cuda(x, gamma)
        M, Cc = x.shape
        g, b = gamma.float().contiguous(), beta.float().contiguous()
        mean = running_mean.float().contiguous()
        rstd = torch.rsqrt(running_var.float() + eps).contiguous()       # [C] glue
        y = torch.empty_like(x)
        nbytes = float(M) * Cc * 2 * (2 + (1 if residual is not None else 0))
        L.check(_run("batchnorm_fwd", nbytes, lambda: L.load().vmae_bn_apply(
            _p(x), _p(g), _p(b), _p(mean), _p(rstd), _p(residual), _p(y), M, Cc, int(act), _stream())), "vmae_bn_apply")
        _count()
        return y

    @staticmethod
    def backward(ctx, dy):
        raise RuntimeError("eval-mode BatchNorm (running statistics) has no backward on the sm_100a path: call model.train() to "
                           "train (the reference trains with batch statistics, tiny_vit.py:12-18)")


class DepthwiseConv3x3Fn(torch.autograd.Function):
    """Depthwise 3x3, pad 1, stride 1|2 on NHWC bf16; weight = nn.Conv2d weight [C, 1, 3, 3] (fp32)."""

    @staticmethod
    def forward(ctx, x, weight, stride):
        _need_cuda(x, weight)
        N, H, W, Cc = x.shape
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        w = weight.float().contiguous()
        y = torch.empty((N, Ho, Wo, Cc), device=x.device, dtype=BF16)
        L.check(_run("dwconv_fwd", float(x.numel() + y.numel()) * 2, lambda: L.load().vmae_dwconv3x3_fwd(
            _p(x), _p(w), _p(y), N, H, W, Cc, stride, _stream())), "vmae_dwconv3x3_fwd")
        _count()
        ctx.save_for_backward(x, w)
        ctx.stride = stride
        ctx.w_shape = weight.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        N, H, W, Cc = x.shape
        dy = cast_bf16(dy.contiguous())
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros((Cc, 9), device=x.device, dtype=F32) if ctx.needs_input_grad[1] else None
        L.check(_run("dwconv_bwd", float(x.numel() * 2 + dy.numel() * 2) * 2, lambda: L.load().vmae_dwconv3x3_bwd(
            _p(x), _p(dy), _p(w), _p(dx), _p(dw), N, H, W, Cc, ctx.stride, _stream())), "vmae_dwconv3x3_bwd")
        _count(2)
        return dx, (dw.reshape(ctx.w_shape) if dw is not None else None), None


class SEPoolFn(torch.autograd.Function):
    """Global average pool per image: bf16 [N, HW, C] -> fp32 [N, C]."""

    @staticmethod
    def forward(ctx, x):
        N, HW, Cc = x.shape
        out = torch.empty((N, Cc), device=x.device, dtype=F32)
        L.check(L.load().vmae_se_pool(_p(x), None, _p(out), N, HW, Cc, 1.0 / HW, _stream()), "vmae_se_pool")
        _count()
        ctx.shape = (N, HW, Cc)
        return out

    @staticmethod
    def backward(ctx, dpooled):
        N, HW, Cc = ctx.shape
        dx = torch.empty((N, HW, Cc), device=dpooled.device, dtype=BF16)
        # dx[n,p,c] = dpooled[n,c] / HW  (broadcast form of the gating kernel)
        L.check(L.load().vmae_se_scale(None, None, _p(dpooled.float().contiguous()), _p(dx), N, HW, Cc, 1.0 / HW, _stream()), "vmae_se_scale")
        _count()
        return dx


class SEScaleFn(torch.autograd.Function):
    """y[n,p,c] = x[n,p,c] * gate[n,c]; x bf16 [N,HW,C], gate fp32 [N,C]."""

    @staticmethod
    def forward(ctx, x, gate):
        N, HW, Cc = x.shape
        gate = gate.float().contiguous()
        y = torch.empty_like(x)
        L.check(_run("se_scale", float(x.numel()) * 4, lambda: L.load().vmae_se_scale(_p(x), _p(gate), None, _p(y), N, HW, Cc, 0.0, _stream())),
                "vmae_se_scale")
        _count()
        ctx.save_for_backward(x, gate)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gate = ctx.saved_tensors
        N, HW, Cc = x.shape
        dy = cast_bf16(dy.contiguous())
        dx = torch.empty_like(x)
        dgate = torch.empty((N, Cc), device=x.device, dtype=F32)
        lib = L.load()
        L.check(lib.vmae_se_scale(_p(dy), _p(gate), None, _p(dx), N, HW, Cc, 0.0, _stream()), "vmae_se_scale")
        L.check(lib.vmae_se_pool(_p(x), _p(dy), _p(dgate), N, HW, Cc, 1.0, _stream()), "vmae_se_pool")
        _count(2)
        return dx, dgate


class Im2col3x3Fn(torch.autograd.Function):
    """NHWC bf16 -> [N*Ho*Wo, 9*C] window matrix in (kh, kw, c) order (pad 1); backward is the col2im gather."""

    @staticmethod
    def forward(ctx, x, stride):
        N, H, W, Cc = x.shape
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        col = torch.empty((N * Ho * Wo, 9 * Cc), device=x.device, dtype=BF16)
        L.check(L.load().vmae_im2col3x3(_p(x), _p(col), N, H, W, Cc, stride, _stream()), "vmae_im2col3x3")
        _count()
        ctx.geom = (N, H, W, Cc, stride)
        return col

    @staticmethod
    def backward(ctx, dcol):
        N, H, W, Cc, stride = ctx.geom
        dcol = cast_bf16(dcol.contiguous())
        dx = torch.empty((N, H, W, Cc), device=dcol.device, dtype=BF16)
        L.check(L.load().vmae_col2im3x3(_p(dcol), _p(dx), N, H, W, Cc, stride, _stream()), "vmae_col2im3x3")
        _count()
        return dx, None


def im2col3x3_image(x_nchw_f32, stride, kp):
    """fp32 NCHW image -> bf16 [N*Ho*Wo, kp] window matrix in (c, kh, kw) order, zero padded to kp columns (no gradient:
    the input is data)."""
    x = x_nchw_f32.detach().float().contiguous()
    N, Cin, H, W = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    col = torch.empty((N * Ho * Wo, kp), device=x.device, dtype=BF16)
    L.check(L.load().vmae_im2col3x3_nchw_f32(_p(x), _p(col), N, Cin, H, W, stride, kp, _stream()), "vmae_im2col3x3_nchw_f32")
    _count()
    return col, Ho, Wo
