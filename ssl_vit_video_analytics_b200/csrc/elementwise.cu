// K3 / K8 / K9 and small HBM-bound helpers: LayerNorm fwd/bwd, dtype casts, bias-gradient column sums,
// Gaussian feature noise (Philox), fp32 scale/axpy. One warp owns one row; 128-bit loads; warp-shuffle reductions.
#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"
#include "../../include/vmae.h"

namespace vmae {

constexpr int EW_THREADS = 256;
constexpr int EW_WARPS = EW_THREADS / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
// "raw" 4-element vectors: what a prefetch holds in registers before it is converted (bf16 stays packed: half the registers)
template <typename T> struct Raw4;
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__nv_bfloat16> { using type = uint2; };
__device__ __forceinline__ float4 ldraw4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint2 ldraw4(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ float4 cvt4(float4 v) { return v; }
__device__ __forceinline__ float4 cvt4(uint2 u) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta, biased variance, eps inside the sqrt (nn.LayerNorm).
// ------------------------------------------------------------------------------------------------------------------
template <typename XT, typename YT, int NV>
__global__ void __launch_bounds__(EW_THREADS) layernorm_fwd_kernel(const XT* __restrict__ x, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, YT* __restrict__ y,
                                                                   float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                   long long rows, int cols, float eps) {
  pdl_wait();   // programmatic dependent launch: this grid may have started while its producer was still draining
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long warp_stride = (long long)gridDim.x * EW_WARPS;
  float4 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < cols) {
      g[i] = __ldg(reinterpret_cast<const float4*>(gamma + c));
      b[i] = __ldg(reinterpret_cast<const float4*>(beta + c));
    }
  }
  const float inv_n = 1.0f / (float)cols;
  // software pipeline: the loads of this warp's NEXT row are issued before the current row is reduced, so a warp keeps two
  // rows of HBM traffic in flight instead of stalling on one (the kernel is latency-, not bandwidth-limited otherwise)
  typename Raw4<XT>::type nxt[NV];
  if (warp_global < rows) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) nxt[i] = ldraw4(x + warp_global * cols + c);
    }
  }
  for (long long r = warp_global; r < rows; r += warp_stride) {
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        v[i] = cvt4(nxt[i]);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    if (r + warp_stride < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < cols) nxt[i] = ldraw4(x + (r + warp_stride) * cols + c);
      }
    }
    const float mean = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
    YT* yr = y + r * cols;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        float4 o;
        o.x = (v[i].x - mean) * rstd * g[i].x + b[i].x;
        o.y = (v[i].y - mean) * rstd * g[i].y + b[i].y;
        o.z = (v[i].z - mean) * rstd * g[i].z + b[i].z;
        o.w = (v[i].w - mean) * rstd * g[i].w + b[i].w;
        store4<YT>(yr + c, o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm backward. dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; plus the residual-stream
// gradient if given. dgamma/dbeta: per-lane register partials -> smem cross-warp reduce -> one red.add per column/CTA.
// ------------------------------------------------------------------------------------------------------------------
template <typename DT, typename XT, int NV>
__global__ void __launch_bounds__(EW_THREADS, (NV <= 4) ? 2 : 1) layernorm_bwd_kernel(const DT* __restrict__ dy, const XT* __restrict__ x,
                                                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ dx_resid,
                                                                   float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_out_bf16,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   long long rows, int cols) {
  extern __shared__ float red_smem[];  // [EW_WARPS][2][cols]
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * EW_WARPS + warp;
  const long long warp_stride = (long long)gridDim.x * EW_WARPS;
  float4 g[NV], acc_g[NV], acc_b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    acc_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols) g[i] = __ldg(reinterpret_cast<const float4*>(gamma + c));
  }
  const float inv_n = 1.0f / (float)cols;
  // software pipeline (see the forward): x, dy, the residual gradient and the row statistics of the NEXT row are requested
  // before the current row is reduced
  typename Raw4<XT>::type nx[NV];
  typename Raw4<DT>::type nd[NV];
  float4 nr[NV];
  float nmu = 0.f, nrs = 0.f;
  auto prefetch = [&](long long r) {
    nmu = __ldg(mean + r);
    nrs = __ldg(rstd + r);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        nx[i] = ldraw4(x + r * cols + c);
        nd[i] = ldraw4(dy + r * cols + c);
        if (dx_resid) nr[i] = __ldg(reinterpret_cast<const float4*>(dx_resid + r * cols + c));
      }
    }
  };
  if (warp_global < rows) prefetch(warp_global);
  for (long long r = warp_global; r < rows; r += warp_stride) {
    const float mu = nmu, rs = nrs;
    float4 xh[NV], d[NV], rv[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        const float4 xv = cvt4(nx[i]);
        d[i] = cvt4(nd[i]);
        if (dx_resid) rv[i] = nr[i];
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        const float g0 = d[i].x * g[i].x, g1 = d[i].y * g[i].y, g2 = d[i].z * g[i].z, g3 = d[i].w * g[i].w;
        s1 += (g0 + g1) + (g2 + g3);
        s2 += (g0 * xh[i].x + g1 * xh[i].y) + (g2 * xh[i].z + g3 * xh[i].w);
        acc_g[i].x += d[i].x * xh[i].x; acc_g[i].y += d[i].y * xh[i].y;
        acc_g[i].z += d[i].z * xh[i].z; acc_g[i].w += d[i].w * xh[i].w;
        acc_b[i].x += d[i].x; acc_b[i].y += d[i].y; acc_b[i].z += d[i].z; acc_b[i].w += d[i].w;
      }
    }
    if (r + warp_stride < rows) prefetch(r + warp_stride);
    const float c1 = warp_sum(s1) * inv_n;
    const float c2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < cols) {
        float4 o;
        o.x = rs * (d[i].x * g[i].x - c1 - xh[i].x * c2);
        o.y = rs * (d[i].y * g[i].y - c1 - xh[i].y * c2);
        o.z = rs * (d[i].z * g[i].z - c1 - xh[i].z * c2);
        o.w = rs * (d[i].w * g[i].w - c1 - xh[i].w * c2);
        if (dx_resid) { o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w; }
        if (dx_out) *reinterpret_cast<float4*>(dx_out + r * cols + c) = o;
        if (dx_out_bf16) store4<__nv_bfloat16>(dx_out_bf16 + r * cols + c, o);
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr) return;
  float* sg = red_smem + (size_t)warp * 2 * cols;
  float* sb = sg + cols;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < cols) {
      *reinterpret_cast<float4*>(sg + c) = acc_g[i];
      *reinterpret_cast<float4*>(sb + c) = acc_b[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += EW_THREADS) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < EW_WARPS; ++w) {
      tg += red_smem[(size_t)w * 2 * cols + c];
      tb += red_smem[(size_t)w * 2 * cols + cols + c];
    }
    if (dgamma) atomicAdd(dgamma + c, tg);
    if (dbeta) atomicAdd(dbeta + c, tb);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Element dropout with a stateless keep mask: one 32-bit hash per element pair, 16 bits per element.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ew_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// keep bits of elements 2*pair (low half) and 2*pair+1 (high half)
__device__ __forceinline__ uint32_t ew_drop_word(unsigned long long seed, long long pair) {
  const uint32_t hi = ew_mix32(static_cast<uint32_t>(static_cast<unsigned long long>(pair) >> 32) ^ static_cast<uint32_t>(seed >> 32));
  return ew_mix32(static_cast<uint32_t>(pair) * 0x9E3779B1u + static_cast<uint32_t>(seed) + hi);
}
template <typename ST, typename DT>
__global__ void __launch_bounds__(EW_THREADS) dropout_apply_kernel(const ST* __restrict__ src, const float* __restrict__ add, DT* __restrict__ dst,
                                                                    long long n, uint32_t thr, float rp, unsigned long long seed) {
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = load4<ST>(src + 4 * i);
    const uint32_t w0 = ew_drop_word(seed, 2 * i), w1 = ew_drop_word(seed, 2 * i + 1);
    v.x = ((w0 & 0xffffu) < thr) ? 0.f : v.x * rp;
    v.y = ((w0 >> 16) < thr) ? 0.f : v.y * rp;
    v.z = ((w1 & 0xffffu) < thr) ? 0.f : v.z * rp;
    v.w = ((w1 >> 16) < thr) ? 0.f : v.w * rp;
    if (add) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(add) + i);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    store4<DT>(dst + 4 * i, v);
  }
}
__global__ void __launch_bounds__(EW_THREADS) dropout_mask_kernel(uint8_t* __restrict__ mask, long long n, uint32_t thr, unsigned long long seed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t w = ew_drop_word(seed, i >> 1);
    mask[i] = (((i & 1) ? (w >> 16) : (w & 0xffffu)) >= thr);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// casts / scale
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long n8 = n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    reinterpret_cast<uint4*>(y)[i] = u;
  }
  for (long long i = n8 * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}
// one launch for a whole parameter set: entry t of the table = {src fp32, dst bf16, element count}
struct CastEntry {
  const float* src;
  __nv_bfloat16* dst;
  long long n;
};
__global__ void __launch_bounds__(EW_THREADS) multi_cast_f32_bf16_kernel(const CastEntry* __restrict__ table) {
  const CastEntry e = table[blockIdx.y];
  const long long n8 = e.n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(e.src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(e.src) + 2 * i + 1);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    reinterpret_cast<uint4*>(e.dst)[i] = u;
  }
  for (long long i = n8 * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += stride) e.dst[i] = __float2bfloat16_rn(e.src[i]);
}

// ----------------------------------------------------------------------------------------------------------------------
// AdamW over many tensors in one launch (train_ssl_mae.py:163: torch.optim.AdamW on every parameter, decoupled weight decay).
// table: one entry per chunk of at most ADAMW_CHUNK elements of one tensor; hp (device): lr, beta1, beta2, eps,
// weight_decay, step (float, already incremented for this update by adamw_tick_kernel). Everything the update depends on
// lives on the device, so a captured step replays correctly while the host changes lr between replays.
// ----------------------------------------------------------------------------------------------------------------------
struct AdamWEntry {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};
constexpr int ADAMW_CHUNK = 16384;
__global__ void adamw_tick_kernel(float* hp) { hp[5] += 1.0f; }
__global__ void __launch_bounds__(EW_THREADS) adamw_multi_kernel(const AdamWEntry* __restrict__ table, const float* __restrict__ hp) {
  const AdamWEntry e = table[blockIdx.x];
  const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], step = hp[5];
  const float bc1 = 1.0f - powf(b1, step), bc2 = 1.0f - powf(b2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.0f - lr * wd;
  auto upd = [&](float& p, float g, float& m, float& v) {
    p *= decay;
    m = b1 * m + (1.0f - b1) * g;            // torch: exp_avg.lerp_(grad, 1 - beta1)
    v = b2 * v + (1.0f - b2) * g * g;        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p -= step_size * (m / denom);
  };
  const bool vec = (((uintptr_t)e.p | (uintptr_t)e.g | (uintptr_t)e.m | (uintptr_t)e.v) & 15) == 0;
  const long long n4 = vec ? e.n / 4 : 0;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 p = reinterpret_cast<float4*>(e.p)[i], m = reinterpret_cast<float4*>(e.m)[i], v = reinterpret_cast<float4*>(e.v)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(e.g) + i);
    upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
    reinterpret_cast<float4*>(e.p)[i] = p;
    reinterpret_cast<float4*>(e.m)[i] = m;
    reinterpret_cast<float4*>(e.v)[i] = v;
  }
  for (long long i = n4 * 4 + threadIdx.x; i < e.n; i += blockDim.x) {
    float p = e.p[i], m = e.m[i], v = e.v[i];
    upd(p, e.g[i], m, v);
    e.p[i] = p; e.m[i] = m; e.v[i] = v;
  }
}

__global__ void __launch_bounds__(EW_THREADS) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    reinterpret_cast<float4*>(y)[i] = load4<__nv_bfloat16>(x + 4 * i);
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __bfloat162float(x[i]);
}
__global__ void __launch_bounds__(EW_THREADS) scale_f32_kernel(const float* __restrict__ x, float* __restrict__ y, float alpha, int accumulate, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // separate IEEE multiply and add (no FMA contraction): bit-identical to the reference's CPU `acc += t * w`
    const float v = __fmul_rn(alpha, x[i]);
    y[i] = accumulate ? __fadd_rn(y[i], v) : v;
  }
}

__global__ void __launch_bounds__(EW_THREADS) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pre,
                                                              __nv_bfloat16* __restrict__ out, long long n) {
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 d = load4<__nv_bfloat16>(dy + 4 * i);
    const float4 h = load4<__nv_bfloat16>(pre + 4 * i);
    const float4 o = make_float4(d.x * h.x, d.y * h.y, d.z * h.z, d.w * h.w);
    store4<__nv_bfloat16>(out + 4 * i, o);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// bias gradient: db[n] += sum_m dy[m,n]. grid = (col blocks of 256, row slabs); lane owns 8 consecutive columns.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) colsum_kernel(const T* __restrict__ dy, long long ld, float* __restrict__ db, long long M, int N,
                                                            long long rows_per_block) {
  __shared__ float red[EW_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (col < N) {
    for (long long r = r0 + warp; r < r1; r += EW_WARPS) {
      const float4 a = load4<T>(dy + r * ld + col);
      const float4 b = load4<T>(dy + r * ld + col + 4);
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;  // 256 threads <-> 256 columns
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < EW_WARPS; ++w) t += red[w][c];
  if (blockIdx.x * 256 + c < N) atomicAdd(db + blockIdx.x * 256 + c, t);
}

// ------------------------------------------------------------------------------------------------------------------
// Gaussian feature noise: Philox4x32-10 + Box-Muller, 4 normals per counter.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) gaussian_noise_kernel(const T* __restrict__ z, T* __restrict__ out, float* __restrict__ eps_out,
                                                                    long long n, float sigma, unsigned long long seed, unsigned long long offset) {
  const long long n4 = (n + 3) / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 e = philox_normal4(seed, offset + (unsigned long long)i);
    const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long k = 4 * i + j;
      if (k < n) {
        float v;
        if constexpr (sizeof(T) == 4) v = (float)z[k]; else v = __bfloat162float(z[k]);
        v = __fmaf_rn(sigma, ev[j], v);   // one rounding: round(z + sigma * eps)
        if constexpr (sizeof(T) == 4) out[k] = v; else out[k] = __float2bfloat16_rn(v);
        if (eps_out) eps_out[k] = ev[j];
      }
    }
  }
}

static inline int ew_grid(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace vmae

using namespace vmae;

template <typename XT, typename YT>
static int ln_fwd_dispatch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, long long rows, int cols,
                           float eps, cudaStream_t s) {
  const int nv = (cols + 127) / 128;
  const int grid = ew_grid(rows, EW_WARPS);
#define LN_CASE(NV)                                                                                                             \
  case NV:                                                                                                                      \
    VMAE_CUDA(launch_pdl(PDL_LN, layernorm_fwd_kernel<XT, YT, NV>, dim3(grid), dim3(EW_THREADS), 0, s, (const XT*)x, gamma, beta, (YT*)y, mean, rstd, rows, cols, eps)); \
    break;
  switch (nv) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
    default: return fail(VMAE_ERR_UNSUPPORTED, "layernorm: cols=%d > 1024 unsupported", cols);
  }
#undef LN_CASE
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype, float* mean,
                                  float* rstd, long long rows, int cols, float eps, void* stream) {
  VMAE_CHECK_ARG(x && gamma && beta && y, "vmae_layernorm_fwd: null pointer");
  VMAE_CHECK_ARG(rows >= 0 && cols > 0 && cols % 4 == 0, "vmae_layernorm_fwd: cols=%d must be a positive multiple of 4", cols);
  if (rows == 0) return VMAE_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == VMAE_DTYPE_F32 && y_dtype == VMAE_DTYPE_BF16) return ln_fwd_dispatch<float, __nv_bfloat16>(x, gamma, beta, y, mean, rstd, rows, cols, eps, s);
  if (x_dtype == VMAE_DTYPE_F32 && y_dtype == VMAE_DTYPE_F32) return ln_fwd_dispatch<float, float>(x, gamma, beta, y, mean, rstd, rows, cols, eps, s);
  if (x_dtype == VMAE_DTYPE_BF16 && y_dtype == VMAE_DTYPE_BF16) return ln_fwd_dispatch<__nv_bfloat16, __nv_bfloat16>(x, gamma, beta, y, mean, rstd, rows, cols, eps, s);
  if (x_dtype == VMAE_DTYPE_BF16 && y_dtype == VMAE_DTYPE_F32) return ln_fwd_dispatch<__nv_bfloat16, float>(x, gamma, beta, y, mean, rstd, rows, cols, eps, s);
  return fail(VMAE_ERR_BAD_ARG, "vmae_layernorm_fwd: bad dtypes %d/%d", x_dtype, y_dtype);
}

template <typename DT, typename XT>
static int ln_bwd_dispatch(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const float* dx_resid,
                           float* dx_out, void* dx_out_bf16, float* dgamma, float* dbeta, long long rows, int cols, cudaStream_t s) {
  const int nv = (cols + 127) / 128;
  long long g = (rows + EW_WARPS - 1) / EW_WARPS;
  // one resident wave (2 CTAs per SM at 128 registers): every CTA ends with 2 * cols atomics on the same dgamma / dbeta
  // addresses, so extra waves only add contention (the 10240-row encoder launches were atomics-bound at 8 waves: 20 -> 9 us)
  const long long cap = (long long)num_sms() * 2;
  const int grid = (int)(g > cap ? cap : (g < 1 ? 1 : g));
  const size_t smem = (size_t)EW_WARPS * 2 * cols * sizeof(float);
#define LN_CASE(NV)                                                                                                      \
  case NV: {                                                                                                             \
    auto k = layernorm_bwd_kernel<DT, XT, NV>;                                                                           \
    if (smem > 40 * 1024) VMAE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    VMAE_CUDA(launch_pdl(PDL_LN, k, dim3(grid), dim3(EW_THREADS), smem, s, (const DT*)dy, (const XT*)x, gamma, mean, rstd, dx_resid, dx_out,  \
                         (__nv_bfloat16*)dx_out_bf16, dgamma, dbeta, rows, cols));                                        \
  } break;
  switch (nv) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
    default: return fail(VMAE_ERR_UNSUPPORTED, "layernorm: cols=%d > 1024 unsupported", cols);
  }
#undef LN_CASE
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* gamma, const float* mean,
                                  const float* rstd, const float* dx_resid, float* dx_out, void* dx_out_bf16, float* dgamma, float* dbeta,
                                  long long rows, int cols, void* stream) {
  VMAE_CHECK_ARG(dy && x && gamma && mean && rstd, "vmae_layernorm_bwd: null pointer");
  VMAE_CHECK_ARG(dx_out || dx_out_bf16, "vmae_layernorm_bwd: no output");
  VMAE_CHECK_ARG(rows >= 0 && cols > 0 && cols % 4 == 0, "vmae_layernorm_bwd: cols=%d must be a positive multiple of 4", cols);
  if (rows == 0) return VMAE_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dy_dtype == VMAE_DTYPE_BF16 && x_dtype == VMAE_DTYPE_F32) return ln_bwd_dispatch<__nv_bfloat16, float>(dy, x, gamma, mean, rstd, dx_resid, dx_out, dx_out_bf16, dgamma, dbeta, rows, cols, s);
  if (dy_dtype == VMAE_DTYPE_F32 && x_dtype == VMAE_DTYPE_F32) return ln_bwd_dispatch<float, float>(dy, x, gamma, mean, rstd, dx_resid, dx_out, dx_out_bf16, dgamma, dbeta, rows, cols, s);
  if (dy_dtype == VMAE_DTYPE_BF16 && x_dtype == VMAE_DTYPE_BF16) return ln_bwd_dispatch<__nv_bfloat16, __nv_bfloat16>(dy, x, gamma, mean, rstd, dx_resid, dx_out, dx_out_bf16, dgamma, dbeta, rows, cols, s);
  if (dy_dtype == VMAE_DTYPE_F32 && x_dtype == VMAE_DTYPE_BF16) return ln_bwd_dispatch<float, __nv_bfloat16>(dy, x, gamma, mean, rstd, dx_resid, dx_out, dx_out_bf16, dgamma, dbeta, rows, cols, s);
  return fail(VMAE_ERR_BAD_ARG, "vmae_layernorm_bwd: bad dtypes %d/%d", dy_dtype, x_dtype);
}

template <typename ST>
static int dropout_dispatch(const void* src, const float* add, void* dst, int dst_dtype, long long n, uint32_t thr, float rp,
                            unsigned long long seed, cudaStream_t s) {
  const int grid = ew_grid(n / 4, EW_THREADS);
  if (dst_dtype == VMAE_DTYPE_BF16)
    dropout_apply_kernel<ST, __nv_bfloat16><<<grid, EW_THREADS, 0, s>>>((const ST*)src, add, (__nv_bfloat16*)dst, n, thr, rp, seed);
  else
    dropout_apply_kernel<ST, float><<<grid, EW_THREADS, 0, s>>>((const ST*)src, add, (float*)dst, n, thr, rp, seed);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_dropout_apply(const void* src, int src_dtype, const float* add, void* dst, int dst_dtype, long long n, float p_drop,
                                  unsigned long long seed, void* stream) {
  VMAE_CHECK_ARG(n >= 0 && n % 4 == 0 && (n == 0 || (src && dst)), "vmae_dropout_apply: bad args (n must be a multiple of 4)");
  VMAE_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "vmae_dropout_apply: p_drop=%f outside [0, 1)", (double)p_drop);
  VMAE_CHECK_ARG((src_dtype == VMAE_DTYPE_BF16 || src_dtype == VMAE_DTYPE_F32) && (dst_dtype == VMAE_DTYPE_BF16 || dst_dtype == VMAE_DTYPE_F32),
                 "vmae_dropout_apply: bad dtypes %d/%d", src_dtype, dst_dtype);
  if (n == 0) return VMAE_OK;
  const uint32_t thr = static_cast<uint32_t>(p_drop * 65536.0f + 0.5f);
  const float rp = 65536.0f / (65536.0f - static_cast<float>(thr));
  if (src_dtype == VMAE_DTYPE_BF16) return dropout_dispatch<__nv_bfloat16>(src, add, dst, dst_dtype, n, thr, rp, seed, (cudaStream_t)stream);
  return dropout_dispatch<float>(src, add, dst, dst_dtype, n, thr, rp, seed, (cudaStream_t)stream);
}
extern "C" int vmae_dropout_mask(void* mask, long long n, float p_drop, unsigned long long seed, void* stream) {
  VMAE_CHECK_ARG(n >= 0 && (n == 0 || mask), "vmae_dropout_mask: bad args");
  VMAE_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "vmae_dropout_mask: p_drop=%f outside [0, 1)", (double)p_drop);
  if (n == 0) return VMAE_OK;
  dropout_mask_kernel<<<ew_grid(n, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>((uint8_t*)mask, n, static_cast<uint32_t>(p_drop * 65536.0f + 0.5f), seed);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
  VMAE_CHECK_ARG(n >= 0 && (n == 0 || (x && y)), "vmae_cast_f32_to_bf16: bad args");
  if (n == 0) return VMAE_OK;
  VMAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "vmae_cast_f32_to_bf16: pointers must be 16-byte aligned");
  cast_f32_bf16_kernel<<<ew_grid(n / 8 + 1, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, n);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_multi_cast_f32_to_bf16(const void* table, int count, void* stream) {
  VMAE_CHECK_ARG(count >= 0 && (count == 0 || table), "vmae_multi_cast_f32_to_bf16: bad args");
  if (count == 0) return VMAE_OK;
  VMAE_CHECK_ARG(count <= 65535, "vmae_multi_cast_f32_to_bf16: at most 65535 tensors per call");
  multi_cast_f32_bf16_kernel<<<dim3(16, count), EW_THREADS, 0, (cudaStream_t)stream>>>((const CastEntry*)table);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_adamw_multi(const void* table, int chunks, float* hp, void* stream) {
  VMAE_CHECK_ARG(chunks >= 0 && hp && (chunks == 0 || table), "vmae_adamw_multi: bad args");
  adamw_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hp);
  VMAE_LAUNCH_CHECK();
  if (chunks == 0) return VMAE_OK;
  adamw_multi_kernel<<<chunks, EW_THREADS, 0, (cudaStream_t)stream>>>((const AdamWEntry*)table, hp);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_adamw_chunk(void) { return ADAMW_CHUNK; }
extern "C" int vmae_cast_bf16_to_f32(const void* x, float* y, long long n, void* stream) {
  VMAE_CHECK_ARG(n >= 0 && (n == 0 || (x && y)), "vmae_cast_bf16_to_f32: bad args");
  if (n == 0) return VMAE_OK;
  VMAE_CHECK_ARG(((uintptr_t)x & 7) == 0 && ((uintptr_t)y & 15) == 0, "vmae_cast_bf16_to_f32: pointers misaligned");
  cast_bf16_f32_kernel<<<ew_grid(n / 4 + 1, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, n);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_scale_f32(const float* x, float* y, float alpha, int accumulate, long long n, void* stream) {
  VMAE_CHECK_ARG(n >= 0 && (n == 0 || (x && y)), "vmae_scale_f32: bad args");
  if (n == 0) return VMAE_OK;
  scale_f32_kernel<<<ew_grid(n, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>(x, y, alpha, accumulate, n);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_gelu_bwd(const void* dy, const void* pre, void* out, long long n, void* stream) {
  VMAE_CHECK_ARG(n >= 0 && n % 4 == 0 && (n == 0 || (dy && pre && out)), "vmae_gelu_bwd: bad args (n must be a multiple of 4)");
  if (n == 0) return VMAE_OK;
  gelu_bwd_kernel<<<ew_grid(n / 4, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)pre,
                                                                                       (__nv_bfloat16*)out, n);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_colsum_accum(const void* dy, int dy_dtype, long long ld, float* db, long long M, int N, void* stream) {
  VMAE_CHECK_ARG(dy && db && M >= 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "vmae_colsum_accum: bad args (N=%d ld=%lld must be multiples of 8)", N, ld);
  if (M == 0) return VMAE_OK;
  const int col_blocks = (N + 255) / 256;
  long long slabs = ((long long)num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (slabs > (M + 63) / 64) slabs = (M + 63) / 64;
  if (slabs < 1) slabs = 1;
  const long long rows_per_block = (M + slabs - 1) / slabs;
  dim3 grid(col_blocks, (unsigned)((M + rows_per_block - 1) / rows_per_block));
  if (dy_dtype == VMAE_DTYPE_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, ld, db, M, N, rows_per_block);
  else if (dy_dtype == VMAE_DTYPE_F32)
    colsum_kernel<float><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>((const float*)dy, ld, db, M, N, rows_per_block);
  else
    return fail(VMAE_ERR_BAD_ARG, "vmae_colsum_accum: bad dtype %d", dy_dtype);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_add_gaussian_noise(const void* z, void* out, int dtype, float* eps_out, long long n, float sigma, unsigned long long seed,
                                       unsigned long long offset, void* stream) {
  VMAE_CHECK_ARG(n >= 0 && (n == 0 || (z && out)), "vmae_add_gaussian_noise: bad args");
  if (n == 0) return VMAE_OK;
  const int grid = ew_grid((n + 3) / 4, EW_THREADS);
  if (dtype == VMAE_DTYPE_F32)
    gaussian_noise_kernel<float><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>((const float*)z, (float*)out, eps_out, n, sigma, seed, offset);
  else if (dtype == VMAE_DTYPE_BF16)
    gaussian_noise_kernel<__nv_bfloat16><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, (__nv_bfloat16*)out, eps_out, n, sigma, seed, offset);
  else
    return fail(VMAE_ERR_BAD_ARG, "vmae_add_gaussian_noise: bad dtype %d", dtype);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
