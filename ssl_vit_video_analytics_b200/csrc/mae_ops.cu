// K4 / K5 / K6 — the MAE-specific HBM-bound kernels: fused patchify + per-patch normalise + masked MSE,
// tube-mask index compaction, visible-tubelet gather (im2col of the kept tubes only), decoder-input assembly
// (pos-embed add + mask-token substitution) and their backward passes. One warp per token row throughout.
#include "common.cuh"
#include "../../include/vmae.h"

namespace vmae {

constexpr int MO_THREADS = 256;
constexpr int MO_WARPS = MO_THREADS / 32;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int mo_grid(long long rows) {
  long long g = (rows + MO_WARPS - 1) / MO_WARPS;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------------------------
// K4 forward. One warp per token. The token's pixels are staged in shared memory in prediction order (dt,py,px,c)
// so that both the strided pixel reads (ps-float segments) and the pred/diff accesses are coalesced.
// ------------------------------------------------------------------------------------------------------------------
constexpr int MO_PF = 6;   // prediction chunks (256 elements each) requested ahead of the pixel gather: a 1536-wide token row
// CPS / CPT: compile-time patch size / tubelet depth (0 = use the run-time arguments). The index arithmetic of the gather is
// four divisions per load; with run-time divisors they were most of the kernel's instructions.
template <typename PT, int CPS, int CPT>
__global__ void __launch_bounds__(MO_THREADS) masked_mse_fwd_kernel(const float* __restrict__ clip, const PT* __restrict__ pred,
                                                                    const unsigned char* __restrict__ mask, float* __restrict__ acc,
                                                                    __nv_bfloat16* __restrict__ diff, int B, int T, int H, int W, int pt_arg,
                                                                    int ps_arg, int norm_pix) {
  const int ps = CPS ? CPS : ps_arg, pt = CPT ? CPT : pt_arg;
  extern __shared__ float tok_smem[];  // [MO_WARPS][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Tp = T / pt, hp = H / ps, wp = W / ps;
  const int N = Tp * hp * wp;
  const int D = pt * ps * ps * 3;
  float* buf = tok_smem + (size_t)warp * D;
  const long long total = (long long)B * N;
  const long long wstride = (long long)gridDim.x * MO_WARPS;
  float loss_acc = 0.f, cnt_acc = 0.f;
  for (long long tok = (long long)blockIdx.x * MO_WARPS + warp; tok < total; tok += wstride) {
    const bool masked = mask[tok] != 0;
    if (!masked) {
      if (diff) {
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int e = lane * 8; e < D; e += 256) *reinterpret_cast<uint4*>(diff + tok * D + e) = z;
      }
      continue;
    }
    const int b = (int)(tok / N);
    int n = (int)(tok % N);
    const int w0 = n % wp; n /= wp;
    const int h0 = n % hp;
    const int t0 = n / hp;
    // the token's prediction row does not depend on the pixel gather: request its first MO_PF chunks now, so that both
    // streams are in flight together (a warp works through its tokens one after the other)
    uint4 pf[MO_PF];
    if constexpr (sizeof(PT) == 2) {
#pragma unroll
      for (int k = 0; k < MO_PF; ++k) {
        const int e = lane * 8 + k * 256;
        if (e < D) pf[k] = __ldg(reinterpret_cast<const uint4*>(pred + tok * D + e));
      }
    }
    // gather pixels: source order (c, dt, py, px), px fastest; one 16-byte load covers 4 consecutive px (ps % 4 == 0)
    float s = 0.f;
    for (int e = lane * 4; e < D; e += 128) {
      const int px = e % ps;
      int r = e / ps;
      const int py = r % ps; r /= ps;
      const int dt = r % pt;
      const int c = r / pt;
      const float4 v = __ldg(reinterpret_cast<const float4*>(
          clip + ((((long long)b * 3 + c) * T + (t0 * pt + dt)) * H + (h0 * ps + py)) * W + (w0 * ps + px)));
      float* dst = buf + ((dt * ps + py) * ps + px) * 3 + c;
      dst[0] = v.x; dst[3] = v.y; dst[6] = v.z; dst[9] = v.w;
      s += (v.x + v.y) + (v.z + v.w);
    }
    __syncwarp();
    float mean = 0.f, inv_std = 1.f;
    if (norm_pix) {
      mean = warp_sum_f(s) / (float)D;
      float q = 0.f;
      for (int e = lane; e < D; e += 32) {
        const float a = buf[e] - mean;
        q += a * a;
      }
      const float var = warp_sum_f(q) / (float)(norm_pix == 1 ? D - 1 : D);
      inv_std = rsqrtf(var + 1e-6f);
    }
    float sq = 0.f;
    auto chunk = [&](const int e, const uint4 u_pre, const bool have_pre) {
      float pv[8];
      if constexpr (sizeof(PT) == 2) {
        const uint4 u = have_pre ? u_pre : __ldg(reinterpret_cast<const uint4*>(pred + tok * D + e));
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(p2[t]);
          pv[2 * t] = f.x; pv[2 * t + 1] = f.y;
        }
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(pred + tok * D + e));
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(pred + tok * D + e + 4));
        pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; pv[4] = c4.x; pv[5] = c4.y; pv[6] = c4.z; pv[7] = c4.w;
      }
      float d[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        d[t] = pv[t] - (buf[e + t] - mean) * inv_std;
        sq += d[t] * d[t];
      }
      if (diff) {
        uint4 u;
        __nv_bfloat162 a0 = __floats2bfloat162_rn(d[0], d[1]), a1 = __floats2bfloat162_rn(d[2], d[3]);
        __nv_bfloat162 a2 = __floats2bfloat162_rn(d[4], d[5]), a3 = __floats2bfloat162_rn(d[6], d[7]);
        u.x = *reinterpret_cast<uint32_t*>(&a0); u.y = *reinterpret_cast<uint32_t*>(&a1);
        u.z = *reinterpret_cast<uint32_t*>(&a2); u.w = *reinterpret_cast<uint32_t*>(&a3);
        *reinterpret_cast<uint4*>(diff + tok * D + e) = u;
      }
    };
    if constexpr (sizeof(PT) == 2) {
#pragma unroll
      for (int k = 0; k < MO_PF; ++k) {
        const int e = lane * 8 + k * 256;
        if (e < D) chunk(e, pf[k], true);
      }
      for (int e = lane * 8 + MO_PF * 256; e < D; e += 256) chunk(e, make_uint4(0, 0, 0, 0), false);
    } else {
      for (int e = lane * 8; e < D; e += 256) chunk(e, make_uint4(0, 0, 0, 0), false);
    }
    __syncwarp();
    loss_acc += warp_sum_f(sq) / (float)D;
    cnt_acc += 1.f;
  }
  __shared__ float red[2][MO_WARPS];
  if (lane == 0) { red[0][warp] = loss_acc; red[1][warp] = cnt_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < MO_WARPS; ++w) { a += red[0][w]; c += red[1][w]; }
    if (c > 0.f) { atomicAdd(acc, a); atomicAdd(acc + 1, c); }
  }
}

// The VideoMAE tubelet (2 x 16 x 16 pixels x 3 channels, D = 1536, bf16 predictions) without shared memory: a lane loads the
// same four pixels of all three channels, which are 12 CONSECUTIVE elements of the prediction-ordered row (dt, py, px, c), so
// the channel interleave happens in registers. All 48 pixel values and the 48 predictions of a lane are requested before
// anything is consumed (a warp walks its tokens one after the other: memory-level parallelism has to come from inside a token).
__global__ void __launch_bounds__(MO_THREADS) masked_mse_fwd_p16t2_kernel(const float* __restrict__ clip, const __nv_bfloat16* __restrict__ pred,
                                                                          const unsigned char* __restrict__ mask, float* __restrict__ acc,
                                                                          __nv_bfloat16* __restrict__ diff, int B, int T, int H, int W,
                                                                          int norm_pix) {
  constexpr int D = 1536, NIT = 4;   // 128 groups of 12 elements, 32 lanes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Tp = T / 2, hp = H / 16, wp = W / 16;
  const int N = Tp * hp * wp;
  const long long plane = (long long)H * W, chan = (long long)T * plane;
  const long long total = (long long)B * N;
  const long long wstride = (long long)gridDim.x * MO_WARPS;
  float loss_acc = 0.f, cnt_acc = 0.f;
  for (long long tok = (long long)blockIdx.x * MO_WARPS + warp; tok < total; tok += wstride) {
    if (mask[tok] == 0) {
      if (diff) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int e = lane * 8; e < D; e += 256) *reinterpret_cast<uint4*>(diff + tok * D + e) = z;
      }
      continue;
    }
    const int b = (int)(tok / N);
    int n = (int)(tok % N);
    const int w0 = n % wp; n /= wp;
    const int h0 = n % hp;
    const int t0 = n / hp;
    const float* base = clip + (long long)b * 3 * chan + (long long)(t0 * 2) * plane + (long long)(h0 * 16) * W + w0 * 16;
    const uint2* prow = reinterpret_cast<const uint2*>(pred + tok * D);
    float x[NIT][12];
    uint2 pu[NIT][3];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int g = it * 32 + lane;   // group: (dt, py, px / 4)
      const long long off = (long long)(g >> 6) * plane + (long long)((g >> 2) & 15) * W + (g & 3) * 4;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + c * chan + off));
        x[it][c] = v.x; x[it][3 + c] = v.y; x[it][6 + c] = v.z; x[it][9 + c] = v.w;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) pu[it][k] = __ldg(prow + g * 3 + k);
    }
    float mean = 0.f, inv_std = 1.f;
    if (norm_pix) {
      float s = 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it)
#pragma unroll
        for (int i = 0; i < 12; ++i) s += x[it][i];
      mean = warp_sum_f(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it)
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const float a = x[it][i] - mean;
          q += a * a;
        }
      const float var = warp_sum_f(q) / (float)(norm_pix == 1 ? D - 1 : D);
      inv_std = rsqrtf(var + 1e-6f);
    }
    float sq = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int g = it * 32 + lane;
      uint2 out[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pu[it][k]);
        const float2 f0 = __bfloat1622float2(p2[0]), f1 = __bfloat1622float2(p2[1]);
        const float d0 = f0.x - (x[it][4 * k] - mean) * inv_std, d1 = f0.y - (x[it][4 * k + 1] - mean) * inv_std;
        const float d2 = f1.x - (x[it][4 * k + 2] - mean) * inv_std, d3 = f1.y - (x[it][4 * k + 3] - mean) * inv_std;
        sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        __nv_bfloat162 a0 = __floats2bfloat162_rn(d0, d1), a1 = __floats2bfloat162_rn(d2, d3);
        out[k].x = *reinterpret_cast<uint32_t*>(&a0);
        out[k].y = *reinterpret_cast<uint32_t*>(&a1);
      }
      if (diff) {
        uint2* drow = reinterpret_cast<uint2*>(diff + tok * D);
#pragma unroll
        for (int k = 0; k < 3; ++k) drow[g * 3 + k] = out[k];
      }
    }
    loss_acc += warp_sum_f(sq) / (float)D;
    cnt_acc += 1.f;
  }
  __shared__ float red[2][MO_WARPS];
  if (lane == 0) { red[0][warp] = loss_acc; red[1][warp] = cnt_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < MO_WARPS; ++w) { a += red[0][w]; c += red[1][w]; }
    if (c > 0.f) { atomicAdd(acc, a); atomicAdd(acc + 1, c); }
  }
}

__global__ void masked_mse_finalize_kernel(const float* acc, float* loss, float* scale, int D) {
  const float denom = acc[1] + 1e-6f;
  loss[0] = acc[0] / denom;
  if (scale) scale[0] = 2.0f / ((float)D * denom);
}

__global__ void __launch_bounds__(MO_THREADS) masked_mse_bwd_kernel(const __nv_bfloat16* __restrict__ diff, const float* __restrict__ scale,
                                                                    const float* __restrict__ gout, __nv_bfloat16* __restrict__ dpred, long long n) {
  const float f = scale[0] * (gout ? gout[0] : 1.0f);
  const long long n8 = n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(diff) + i);
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint4 o;
    uint32_t* po = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 v = __bfloat1622float2(p2[t]);
      __nv_bfloat162 r = __floats2bfloat162_rn(v.x * f, v.y * f);
      po[t] = *reinterpret_cast<uint32_t*>(&r);
    }
    reinterpret_cast<uint4*>(dpred)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// tube-mask compaction: ascending indices of the visible (mask == 0) patches of one frame, per clip.
// ------------------------------------------------------------------------------------------------------------------
__global__ void mask_to_indices_kernel(const unsigned char* __restrict__ mask_frame, int* __restrict__ vis_idx, int L, int n_vis) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int base = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    const bool vis = (l < L) && (mask_frame[(long long)b * L + l] == 0);
    const unsigned bal = __ballot_sync(0xffffffffu, vis);
    const int pos = base + __popc(bal & ((1u << lane) - 1u));
    if (vis && pos < n_vis) vis_idx[(long long)b * n_vis + pos] = l;
    base += __popc(bal);
  }
  // fewer visible patches than n_vis: pad with -1 so that the mismatch is detectable
  for (int i = base + lane; i < n_vis; i += 32) vis_idx[(long long)b * n_vis + i] = -1;
}

// ------------------------------------------------------------------------------------------------------------------
// K6 stage 1: gather the visible tubelets into a dense bf16 [rows, 3*pt*ps*ps] matrix (Conv3d weight order).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MO_THREADS) gather_tubelets_kernel(const float* __restrict__ clip, const int* __restrict__ vis_idx,
                                                                     __nv_bfloat16* __restrict__ patches, const float* __restrict__ tpos,
                                                                     const float* __restrict__ spos, float* __restrict__ pos_rows, int B, int T,
                                                                     int H, int W, int pt, int ps, int n_vis, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Tp = T / pt, wp = W / ps;
  const int K = 3 * pt * ps * ps;
  const long long rows = (long long)B * Tp * n_vis;
  const long long wstride = (long long)gridDim.x * MO_WARPS;
  for (long long row = (long long)blockIdx.x * MO_WARPS + warp; row < rows; row += wstride) {
    const int i = (int)(row % n_vis);
    const int tp = (int)((row / n_vis) % Tp);
    const int b = (int)(row / ((long long)n_vis * Tp));
    const int l = vis_idx[(long long)b * n_vis + i];
    const int h0 = l / wp, w0 = l % wp;
    for (int k = lane * 4; k < K; k += 128) {
      const int px = k % ps;
      int r = k / ps;
      const int py = r % ps; r /= ps;
      const int dt = r % pt;
      const int c = r / pt;
      const float4 v = __ldg(reinterpret_cast<const float4*>(
          clip + ((((long long)b * 3 + c) * T + (tp * pt + dt)) * H + (h0 * ps + py)) * W + (w0 * ps + px)));
      __nv_bfloat162 a0 = __floats2bfloat162_rn(v.x, v.y), a1 = __floats2bfloat162_rn(v.z, v.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a0);
      u.y = *reinterpret_cast<uint32_t*>(&a1);
      *reinterpret_cast<uint2*>(patches + row * K + k) = u;
    }
    if (pos_rows) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tpos + (long long)tp * D + d));
        const float4 s = __ldg(reinterpret_cast<const float4*>(spos + (long long)l * D + d));
        *reinterpret_cast<float4*>(pos_rows + row * D + d) = make_float4(a.x + s.x, a.y + s.y, a.z + s.z, a.w + s.w);
      }
    }
  }
}

// position of patch l in the ascending visible list of clip b, or -1 (warp-cooperative search)
__device__ __forceinline__ int find_visible_slot(const int* __restrict__ vis, int n_vis, int l, int lane) {
  int found = -1;
  for (int i0 = 0; i0 < n_vis; i0 += 32) {
    const int i = i0 + lane;
    const bool hit = (i < n_vis) && (vis[i] == l);
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (bal) { found = i0 + __ffs(bal) - 1; break; }
  }
  return found;
}

// ------------------------------------------------------------------------------------------------------------------
// K5 forward: decoder input = visible ? z + tpos[t] + spos[l] : mask_token   (mae_vit_adapter.py:97-104)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MO_THREADS) decoder_input_fwd_kernel(const __nv_bfloat16* __restrict__ z, const int* __restrict__ vis_idx,
                                                                       const unsigned char* __restrict__ mask, const float* __restrict__ tpos,
                                                                       const float* __restrict__ spos, const float* __restrict__ mask_token,
                                                                       float* __restrict__ x_dec, int B, int Tp, int L, int n_vis, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)B * Tp * L;
  const long long wstride = (long long)gridDim.x * MO_WARPS;
  for (long long row = (long long)blockIdx.x * MO_WARPS + warp; row < rows; row += wstride) {
    const int l = (int)(row % L);
    const int tp = (int)((row / L) % Tp);
    const int b = (int)(row / ((long long)L * Tp));
    float* dst = x_dec + row * D;
    if (mask[row]) {
      for (int d = lane * 4; d < D; d += 128)
        *reinterpret_cast<float4*>(dst + d) = __ldg(reinterpret_cast<const float4*>(mask_token + d));
      continue;
    }
    long long zrow = row;
    if (vis_idx) {
      const int slot = find_visible_slot(vis_idx + (long long)b * n_vis, n_vis, l, lane);
      if (slot < 0) {  // mask and index list disagree: poison the row so that it cannot go unnoticed
        for (int d = lane * 4; d < D; d += 128) *reinterpret_cast<float4*>(dst + d) = make_float4(NAN, NAN, NAN, NAN);
        continue;
      }
      zrow = ((long long)b * Tp + tp) * n_vis + slot;
    }
    for (int d = lane * 4; d < D; d += 128) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(z + zrow * D + d));
      const float2 z0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 z1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      const float4 a = __ldg(reinterpret_cast<const float4*>(tpos + (long long)tp * D + d));
      const float4 s = __ldg(reinterpret_cast<const float4*>(spos + (long long)l * D + d));
      // reference order: (z + (tpos + spos))
      *reinterpret_cast<float4*>(dst + d) = make_float4(z0.x + (a.x + s.x), z0.y + (a.y + s.y), z1.x + (a.z + s.z), z1.y + (a.w + s.w));
    }
  }
}

// K5 backward. One CTA per (clip, temporal position): its L rows share one temporal-position row, so the dtpos contribution is
// summed in registers and shared memory and leaves the CTA as ONE red.add per column (row-wise atomics on the Tp rows of dtpos
// serialised ~1 250 same-address updates per element: 95 us for a 24-us read). Masked rows reduce into dmask_token the same
// way; visible rows produce dz (bf16) and scatter-add into dspos (L distinct rows: no contention worth removing).
__global__ void __launch_bounds__(MO_THREADS) decoder_input_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ vis_idx,
                                                                       const unsigned char* __restrict__ mask, __nv_bfloat16* __restrict__ dz,
                                                                       float* __restrict__ dtpos, float* __restrict__ dspos,
                                                                       float* __restrict__ dmask_token, int B, int Tp, int L, int n_vis, int D) {
  extern __shared__ float mt_smem[];  // [MO_WARPS][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.x, b = seg / Tp, tp = seg % Tp;
  constexpr int MAXV = 8;  // D <= 1024
  float4 acc[MAXV], acc_tp[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) acc[i] = acc_tp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = warp; l < L; l += MO_WARPS) {
    const long long row = (long long)seg * L + l;
    const float* src = dx + row * D;
    float4 v[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int d = (i * 32 + lane) * 4;
      v[i] = (d < D) ? __ldg(reinterpret_cast<const float4*>(src + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (mask[row]) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) { acc[i].x += v[i].x; acc[i].y += v[i].y; acc[i].z += v[i].z; acc[i].w += v[i].w; }
      if (!vis_idx && dz) {  // blend form: dz covers every token; masked rows get zero gradient
        for (int d = lane * 4; d < D; d += 128) *reinterpret_cast<uint2*>(dz + row * D + d) = make_uint2(0u, 0u);
      }
      continue;
    }
    long long zrow = row;
    if (vis_idx) {
      const int slot = find_visible_slot(vis_idx + (long long)b * n_vis, n_vis, l, lane);
      if (slot < 0) continue;
      zrow = ((long long)b * Tp + tp) * n_vis + slot;
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int d = (i * 32 + lane) * 4;
      if (d >= D) continue;
      if (dz) {
        __nv_bfloat162 a0 = __floats2bfloat162_rn(v[i].x, v[i].y), a1 = __floats2bfloat162_rn(v[i].z, v[i].w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a0);
        u.y = *reinterpret_cast<uint32_t*>(&a1);
        *reinterpret_cast<uint2*>(dz + zrow * D + d) = u;
      }
      acc_tp[i].x += v[i].x; acc_tp[i].y += v[i].y; acc_tp[i].z += v[i].z; acc_tp[i].w += v[i].w;
      if (dspos) {
        float* p = dspos + (long long)l * D + d;
        atomicAdd(p, v[i].x); atomicAdd(p + 1, v[i].y); atomicAdd(p + 2, v[i].z); atomicAdd(p + 3, v[i].w);
      }
    }
  }
  float* mine = mt_smem + (size_t)warp * D;
  auto block_sum_to = [&](const float4 (&a)[MAXV], float* dst) {   // dst[d] += sum over the CTA's warps of a
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int d = (i * 32 + lane) * 4;
      if (d < D) *reinterpret_cast<float4*>(mine + d) = a[i];
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += MO_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < MO_WARPS; ++w) t += mt_smem[(size_t)w * D + d];
      atomicAdd(dst + d, t);
    }
    __syncthreads();
  };
  if (dmask_token) block_sum_to(acc, dmask_token);
  if (dtpos) block_sum_to(acc_tp, dtpos + (long long)tp * D);
}

// Encoder-side positional rows (tubelet gather backward): one CTA per (clip, temporal position), as above.
__global__ void __launch_bounds__(MO_THREADS) pos_rows_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ vis_idx,
                                                                  float* __restrict__ dtpos, float* __restrict__ dspos, int B, int Tp, int n_vis,
                                                                  int D) {
  extern __shared__ float mt_smem[];  // [MO_WARPS][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.x, b = seg / Tp, tp = seg % Tp;
  constexpr int MAXV = 8;  // D <= 1024
  float4 acc_tp[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) acc_tp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = warp; i0 < n_vis; i0 += MO_WARPS) {
    const long long row = (long long)seg * n_vis + i0;
    const int l = vis_idx[(long long)b * n_vis + i0];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int d = (i * 32 + lane) * 4;
      if (d >= D) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(dx + row * D + d));
      acc_tp[i].x += v.x; acc_tp[i].y += v.y; acc_tp[i].z += v.z; acc_tp[i].w += v.w;
      float* q = dspos + (long long)l * D + d;
      atomicAdd(q, v.x); atomicAdd(q + 1, v.y); atomicAdd(q + 2, v.z); atomicAdd(q + 3, v.w);
    }
  }
  float* mine = mt_smem + (size_t)warp * D;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int d = (i * 32 + lane) * 4;
    if (d < D) *reinterpret_cast<float4*>(mine + d) = acc_tp[i];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += MO_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < MO_WARPS; ++w) t += mt_smem[(size_t)w * D + d];
    atomicAdd(dtpos + (long long)tp * D + d, t);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Stack-2 latent-space losses on gathered rows [rows, cols] (mae/losses.py:19-63): mode 0 = L2 (target optionally
// normalised per row with the BIASED variance, eps inside the sqrt, :36-38), 1 = L1, 2 = cosine (F.normalize eps). One warp
// per row; per-block partial sums are written out and added in a fixed order by the finalize kernel (deterministic).
// The target is a detached tensor in the reference (tinyvit_mae.py:112), so only pred receives a gradient.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RL_THREADS = 256;
template <typename T>
__device__ __forceinline__ float rl_ld(const T* p, long long i);
template <>
__device__ __forceinline__ float rl_ld<float>(const float* p, long long i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float rl_ld<__nv_bfloat16>(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float rl_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// per-row quantities shared by forward and backward
template <typename PT, typename TT>
__device__ __forceinline__ void rl_row_stats(const PT* pred, const TT* target, long long base, int cols, int lane, int mode, int normalize, float eps,
                                             float& a, float& b, float& c) {
  // mode 0: a = target mean, b = 1/sqrt(var + eps) (or 0, 1 without normalisation); mode 2: a = |p|^2, b = |t|^2, c = p.t
  a = 0.f; b = 1.f; c = 0.f;
  if (mode == 0 && normalize) {
    float s = 0.f, q = 0.f;
    for (int i = lane; i < cols; i += 32) { const float t = rl_ld<TT>(target, base + i); s += t; q = fmaf(t, t, q); }
    s = rl_warp_sum(s); q = rl_warp_sum(q);
    a = s / cols;
    b = rsqrtf(fmaxf(q / cols - a * a, 0.f) + eps);
  } else if (mode == 2) {
    float pp = 0.f, tt = 0.f, pt = 0.f;
    for (int i = lane; i < cols; i += 32) {
      const float x = rl_ld<PT>(pred, base + i), t = rl_ld<TT>(target, base + i);
      pp = fmaf(x, x, pp); tt = fmaf(t, t, tt); pt = fmaf(x, t, pt);
    }
    a = rl_warp_sum(pp); b = rl_warp_sum(tt); c = rl_warp_sum(pt);
  }
}
template <typename PT, typename TT>
__global__ void __launch_bounds__(RL_THREADS) row_loss_fwd_kernel(const PT* __restrict__ pred, const TT* __restrict__ target, long long rows, int cols,
                                                                  int mode, int normalize, float eps, float* __restrict__ partials) {
  __shared__ float red[RL_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  for (long long r = (long long)blockIdx.x * (RL_THREADS / 32) + warp; r < rows; r += (long long)gridDim.x * (RL_THREADS / 32)) {
    const long long base = r * cols;
    float a, b, c;
    rl_row_stats<PT, TT>(pred, target, base, cols, lane, mode, normalize, eps, a, b, c);
    if (mode == 2) {
      if (lane == 0) acc += c / (fmaxf(sqrtf(a), eps) * fmaxf(sqrtf(b), eps));   // cosine of the row
    } else {
      float s = 0.f;
      for (int i = lane; i < cols; i += 32) {
        const float d = rl_ld<PT>(pred, base + i) - (rl_ld<TT>(target, base + i) - a) * b;
        s += (mode == 0) ? d * d : fabsf(d);
      }
      acc += s;   // every lane keeps its share; reduced below
    }
  }
  acc = rl_warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < RL_THREADS / 32; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}
__global__ void row_loss_finalize_kernel(const float* __restrict__ partials, int n, long long rows, int cols, int mode, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < n; ++i) t += partials[i];
    *loss = (mode == 2) ? 1.0f - t / (float)rows : t / ((float)rows * (float)cols);
  }
}
template <typename PT, typename TT, typename GT>
__global__ void __launch_bounds__(RL_THREADS) row_loss_bwd_kernel(const PT* __restrict__ pred, const TT* __restrict__ target, long long rows, int cols,
                                                                  int mode, int normalize, float eps, const float* __restrict__ gout,
                                                                  GT* __restrict__ dpred) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float g = __ldg(gout);
  for (long long r = (long long)blockIdx.x * (RL_THREADS / 32) + warp; r < rows; r += (long long)gridDim.x * (RL_THREADS / 32)) {
    const long long base = r * cols;
    float a, b, c;
    rl_row_stats<PT, TT>(pred, target, base, cols, lane, mode, normalize, eps, a, b, c);
    if (mode == 2) {
      const float np = fmaxf(sqrtf(a), eps), nt = fmaxf(sqrtf(b), eps);
      const float cosv = c / (np * nt);
      const bool clamped = sqrtf(a) <= eps;   // F.normalize divides by the constant eps there: d(p/eps) has no projection term
      for (int i = lane; i < cols; i += 32) {
        const float x = rl_ld<PT>(pred, base + i), t = rl_ld<TT>(target, base + i);
        const float d = clamped ? (t / nt) / np : (t / nt - cosv * x / np) / np;
        const float v = -g / (float)rows * d;
        if (sizeof(GT) == 2) reinterpret_cast<__nv_bfloat16*>(dpred)[base + i] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(dpred)[base + i] = v;
      }
    } else {
      const float k = g / ((float)rows * (float)cols);
      for (int i = lane; i < cols; i += 32) {
        const float d = rl_ld<PT>(pred, base + i) - (rl_ld<TT>(target, base + i) - a) * b;
        const float v = (mode == 0) ? 2.0f * k * d : k * ((d > 0.f) - (d < 0.f));
        if (sizeof(GT) == 2) reinterpret_cast<__nv_bfloat16*>(dpred)[base + i] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(dpred)[base + i] = v;
      }
    }
  }
}
static inline int rl_grid(long long rows) {
  long long g = (rows + RL_THREADS / 32 - 1) / (RL_THREADS / 32);
  const long long cap = (long long)num_sms() * 4;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}


// ------------------------------------------------------------------------------------------------------------------
// Input pipeline tail: decoded uint8 frames [B, T, H, W, 3] -> normalised fp32 clip [B, 3, T, H, W], i.e. what
// transforms.ToTensor + Normalize(mean, std) produce per frame followed by the loader's channel reversal and the
// [T, C, H, W] -> [C, T, H, W] permute (mae_loader.py:63-77). One thread per pixel: three byte loads, three coalesced stores.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frames_to_clip_kernel(const uint8_t* __restrict__ frames, float* __restrict__ clip, long long npix, long long thw,
                                                             float s0, float s1, float s2, float b0, float b1, float b2, int flip) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < npix; i += gridDim.x * 256LL) {
    const uint8_t* px = frames + 3 * i;
    const float v0 = fmaf(static_cast<float>(px[0]), s0, b0), v1 = fmaf(static_cast<float>(px[1]), s1, b1),
                v2 = fmaf(static_cast<float>(px[2]), s2, b2);
    const long long bidx = i / thw, r = i % thw;
    float* o = clip + bidx * 3 * thw + r;
    o[0] = flip ? v2 : v0;
    o[thw] = v1;
    o[2 * thw] = flip ? v0 : v2;
  }
}

}  // namespace vmae

using namespace vmae;

extern "C" int vmae_masked_mse_fwd(const float* clip, const void* pred, int pred_dtype, const unsigned char* mask, float* acc, void* diff,
                                   int B, int T, int H, int W, int pt, int ps, int norm_pix, void* stream) {
  VMAE_CHECK_ARG(clip && pred && mask && acc, "vmae_masked_mse_fwd: null pointer");
  VMAE_CHECK_ARG(B > 0 && pt > 0 && ps > 0 && T % pt == 0 && H % ps == 0 && W % ps == 0, "vmae_masked_mse_fwd: clip %dx%dx%d not divisible by tubelet %dx%dx%d", T, H, W, pt, ps, ps);
  VMAE_CHECK_ARG(ps % 4 == 0 && W % 4 == 0 && ((uintptr_t)clip & 15) == 0, "vmae_masked_mse_fwd: patch size and width must be multiples of 4 and the clip 16-byte aligned");
  VMAE_CHECK_ARG(norm_pix >= 0 && norm_pix <= 2, "vmae_masked_mse_fwd: norm_pix must be 0, 1 or 2");
  const int D = pt * ps * ps * 3;
  VMAE_CHECK_ARG(D % 8 == 0, "vmae_masked_mse_fwd: token dim %d must be a multiple of 8", D);
  const size_t smem = (size_t)MO_WARPS * D * sizeof(float);
  VMAE_CHECK_ARG(smem <= 200 * 1024, "vmae_masked_mse_fwd: token dim %d too large", D);
  const long long total = (long long)B * (T / pt) * (H / ps) * (W / ps);
  const int grid = mo_grid(total);
  cudaStream_t s = (cudaStream_t)stream;
  if (pred_dtype == VMAE_DTYPE_BF16 && ps == 16 && pt == 2 && W % 4 == 0) {   // the VideoMAE tubelet: register-only kernel
    masked_mse_fwd_p16t2_kernel<<<grid, MO_THREADS, 0, s>>>(clip, (const __nv_bfloat16*)pred, mask, acc, (__nv_bfloat16*)diff, B, T, H, W, norm_pix);
  } else if (pred_dtype == VMAE_DTYPE_BF16) {
    auto k = masked_mse_fwd_kernel<__nv_bfloat16, 0, 0>;
    if (smem > 40 * 1024) VMAE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, MO_THREADS, smem, s>>>(clip, (const __nv_bfloat16*)pred, mask, acc, (__nv_bfloat16*)diff, B, T, H, W, pt, ps, norm_pix);
  } else if (pred_dtype == VMAE_DTYPE_F32) {
    auto k = masked_mse_fwd_kernel<float, 0, 0>;
    if (smem > 40 * 1024) VMAE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, MO_THREADS, smem, s>>>(clip, (const float*)pred, mask, acc, (__nv_bfloat16*)diff, B, T, H, W, pt, ps, norm_pix);
  } else {
    return fail(VMAE_ERR_BAD_ARG, "vmae_masked_mse_fwd: bad pred dtype %d", pred_dtype);
  }
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_masked_mse_finalize(const float* acc, float* loss, float* scale, int D, void* stream) {
  VMAE_CHECK_ARG(acc && loss && D > 0, "vmae_masked_mse_finalize: bad args");
  masked_mse_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(acc, loss, scale, D);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_masked_mse_bwd(const void* diff, const float* scale, const float* gout, void* dpred, long long n, void* stream) {
  VMAE_CHECK_ARG(diff && scale && dpred && n >= 0 && n % 8 == 0, "vmae_masked_mse_bwd: bad args (n must be a multiple of 8)");
  if (n == 0) return VMAE_OK;
  long long g = (n / 8 + MO_THREADS - 1) / MO_THREADS;
  const long long cap = (long long)num_sms() * 8;
  masked_mse_bwd_kernel<<<(int)(g > cap ? cap : g), MO_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)diff, scale, gout, (__nv_bfloat16*)dpred, n);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_mask_to_indices(const unsigned char* mask_frame, int* vis_idx, int B, int L, int n_vis, void* stream) {
  VMAE_CHECK_ARG(mask_frame && vis_idx && B > 0 && L > 0 && n_vis > 0 && n_vis <= L, "vmae_mask_to_indices: bad args");
  mask_to_indices_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(mask_frame, vis_idx, L, n_vis);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_gather_tubelets(const float* clip, const int* vis_idx, void* patches, const float* temporal_pos, const float* spatial_pos,
                                    float* pos_rows, int B, int T, int H, int W, int pt, int ps, int n_vis, int D, void* stream) {
  VMAE_CHECK_ARG(clip && vis_idx && patches, "vmae_gather_tubelets: null pointer");
  VMAE_CHECK_ARG(B > 0 && pt > 0 && ps > 0 && T % pt == 0 && H % ps == 0 && W % ps == 0 && ps % 4 == 0 && W % 4 == 0, "vmae_gather_tubelets: bad geometry");
  VMAE_CHECK_ARG(!pos_rows || (temporal_pos && spatial_pos && D % 4 == 0), "vmae_gather_tubelets: pos_rows needs both tables and D %% 4 == 0");
  const long long rows = (long long)B * (T / pt) * n_vis;
  gather_tubelets_kernel<<<mo_grid(rows), MO_THREADS, 0, (cudaStream_t)stream>>>(clip, vis_idx, (__nv_bfloat16*)patches, temporal_pos, spatial_pos,
                                                                                  pos_rows, B, T, H, W, pt, ps, n_vis, D);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_decoder_input_fwd(const void* z, const int* vis_idx, const unsigned char* mask, const float* temporal_pos,
                                      const float* spatial_pos, const float* mask_token, float* x_dec, int B, int Tp, int L, int n_vis, int D,
                                      void* stream) {
  VMAE_CHECK_ARG(z && mask && temporal_pos && spatial_pos && mask_token && x_dec, "vmae_decoder_input_fwd: null pointer");
  VMAE_CHECK_ARG(B > 0 && Tp > 0 && L > 0 && D > 0 && D % 4 == 0, "vmae_decoder_input_fwd: bad shape");
  const long long rows = (long long)B * Tp * L;
  decoder_input_fwd_kernel<<<mo_grid(rows), MO_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, vis_idx, mask, temporal_pos, spatial_pos,
                                                                                    mask_token, x_dec, B, Tp, L, n_vis, D);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_decoder_input_bwd(const float* dx_dec, const int* vis_idx, const unsigned char* mask, void* dz, float* dtpos, float* dspos,
                                      float* dmask_token, int B, int Tp, int L, int n_vis, int D, void* stream) {
  VMAE_CHECK_ARG(dx_dec && mask, "vmae_decoder_input_bwd: null pointer");
  VMAE_CHECK_ARG(B > 0 && Tp > 0 && L > 0 && D > 0 && D % 4 == 0 && D <= 1024, "vmae_decoder_input_bwd: bad shape (D <= 1024, D %% 4 == 0)");
  const size_t smem = (size_t)MO_WARPS * D * sizeof(float);
  decoder_input_bwd_kernel<<<B * Tp, MO_THREADS, smem, (cudaStream_t)stream>>>(dx_dec, vis_idx, mask, (__nv_bfloat16*)dz, dtpos, dspos,
                                                                                 dmask_token, B, Tp, L, n_vis, D);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_pos_rows_bwd(const float* dx, const int* vis_idx, float* dtpos, float* dspos, int B, int Tp, int n_vis, int D, void* stream) {
  VMAE_CHECK_ARG(dx && vis_idx && dtpos && dspos && B > 0 && Tp > 0 && n_vis > 0 && D % 4 == 0 && D <= 1024,
                 "vmae_pos_rows_bwd: bad args (D <= 1024, D %% 4 == 0)");
  const size_t smem = (size_t)MO_WARPS * D * sizeof(float);
  pos_rows_bwd_kernel<<<B * Tp, MO_THREADS, smem, (cudaStream_t)stream>>>(dx, vis_idx, dtpos, dspos, B, Tp, n_vis, D);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_row_loss_grid(long long rows) { return rl_grid(rows); }

extern "C" int vmae_row_loss_fwd(const void* pred, int pred_dtype, const void* target, int target_dtype, long long rows, int cols, int mode,
                                 int normalize, float eps, float* partials, float* loss, void* stream) {
  VMAE_CHECK_ARG(pred && target && partials && loss, "vmae_row_loss_fwd: null pointer");
  VMAE_CHECK_ARG(rows > 0 && cols > 0 && mode >= 0 && mode <= 2, "vmae_row_loss_fwd: bad rows=%lld cols=%d mode=%d", rows, cols, mode);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = rl_grid(rows);
#define RL_FWD(PT, TT) row_loss_fwd_kernel<PT, TT><<<grid, RL_THREADS, 0, s>>>((const PT*)pred, (const TT*)target, rows, cols, mode, normalize, eps, partials)
  if (pred_dtype == VMAE_DTYPE_BF16 && target_dtype == VMAE_DTYPE_BF16) RL_FWD(__nv_bfloat16, __nv_bfloat16);
  else if (pred_dtype == VMAE_DTYPE_BF16 && target_dtype == VMAE_DTYPE_F32) RL_FWD(__nv_bfloat16, float);
  else if (pred_dtype == VMAE_DTYPE_F32 && target_dtype == VMAE_DTYPE_BF16) RL_FWD(float, __nv_bfloat16);
  else if (pred_dtype == VMAE_DTYPE_F32 && target_dtype == VMAE_DTYPE_F32) RL_FWD(float, float);
  else return fail(VMAE_ERR_BAD_ARG, "vmae_row_loss_fwd: bad dtypes %d/%d", pred_dtype, target_dtype);
#undef RL_FWD
  VMAE_LAUNCH_CHECK();
  row_loss_finalize_kernel<<<1, 32, 0, s>>>(partials, grid, rows, cols, mode, loss);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_row_loss_bwd(const void* pred, int pred_dtype, const void* target, int target_dtype, long long rows, int cols, int mode,
                                 int normalize, float eps, const float* gout, void* dpred, void* stream) {
  VMAE_CHECK_ARG(pred && target && gout && dpred, "vmae_row_loss_bwd: null pointer");
  VMAE_CHECK_ARG(rows > 0 && cols > 0 && mode >= 0 && mode <= 2, "vmae_row_loss_bwd: bad rows=%lld cols=%d mode=%d", rows, cols, mode);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = rl_grid(rows);
  // dpred has the dtype of pred
#define RL_BWD(PT, TT) row_loss_bwd_kernel<PT, TT, PT><<<grid, RL_THREADS, 0, s>>>((const PT*)pred, (const TT*)target, rows, cols, mode, normalize, eps, gout, (PT*)dpred)
  if (pred_dtype == VMAE_DTYPE_BF16 && target_dtype == VMAE_DTYPE_BF16) RL_BWD(__nv_bfloat16, __nv_bfloat16);
  else if (pred_dtype == VMAE_DTYPE_BF16 && target_dtype == VMAE_DTYPE_F32) RL_BWD(__nv_bfloat16, float);
  else if (pred_dtype == VMAE_DTYPE_F32 && target_dtype == VMAE_DTYPE_BF16) RL_BWD(float, __nv_bfloat16);
  else if (pred_dtype == VMAE_DTYPE_F32 && target_dtype == VMAE_DTYPE_F32) RL_BWD(float, float);
  else return fail(VMAE_ERR_BAD_ARG, "vmae_row_loss_bwd: bad dtypes %d/%d", pred_dtype, target_dtype);
#undef RL_BWD
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_frames_to_clip(const void* frames_u8, float* clip, int B, int T, int H, int W, const float* mean3, const float* std3,
                                   int flip_channels, void* stream) {
  VMAE_CHECK_ARG(frames_u8 && clip && mean3 && std3 && B > 0 && T > 0 && H > 0 && W > 0, "vmae_frames_to_clip: bad args");
  const long long thw = (long long)T * H * W, npix = (long long)B * thw;
  float s[3], b[3];
  for (int c = 0; c < 3; ++c) {   // (u/255 - mean) / std = u * (1/(255 std)) - mean/std; mean3/std3 are HOST pointers
    VMAE_CHECK_ARG(std3[c] != 0.f, "vmae_frames_to_clip: std[%d] == 0", c);
    s[c] = 1.0f / (255.0f * std3[c]);
    b[c] = -mean3[c] / std3[c];
  }
  long long g = (npix + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  frames_to_clip_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)frames_u8, clip, npix, thw, s[0], s[1], s[2], b[0], b[1], b[2],
                                                                     flip_channels);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
