// K7 — depthwise 3x3 convolution (pad 1, stride 1 or 2) on NHWC bf16, TMA-tiled. Replaces the depthwise nn.Conv2d inside
// MBConv (tiny_vit.py:36-56) forward, input gradient (stride 1: the same kernel with the filter flipped) and weight gradient.
//
// One CTA tile = TH x TW output pixels x 64 channels. The input patch with its one-pixel halo arrives as ONE 4-D TMA box
// (C, W, H, N) — out-of-image halo pixels are zero-filled by the TMA unit, so there is no boundary code — into a
// two-stage shared-memory ring: the box of tile t+1 is in flight while tile t is computed, and every input byte crosses
// HBM once (the 9 taps are read from shared memory). A thread owns 8 channels (one 16-byte vector) of a strip of
// consecutive output columns and keeps a sliding 3x3 window of packed vectors in registers.
// HBM-bound: algorithmic bytes = input + output (+ dy for the weight gradient), 2 bytes per element.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/vmae.h"

namespace vmae {

struct DwTiles {
  int N, H, W, C, Ho, Wo;        // input H x W, output Ho x Wo
  int TH, TW;                    // output tile
  int in_rows, in_cols;          // input box (with halo)
  int tiles_h, tiles_w, cblocks; // tile grid; cblocks = C / 64
  int LC, SW;                    // 32 pixel lanes = (32 / LC) rows x LC column strips of SW columns
  long long total;
};

__device__ __forceinline__ void dw_unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = __bfloat1622float2(p[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 dw_pack8(const float (&f)[8]) {
  uint4 u;
  uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    __nv_bfloat162 v = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
    w[t] = *reinterpret_cast<uint32_t*>(&v);
  }
  return u;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

constexpr int DW_THREADS = 256;
constexpr int DW_CB = 64;   // channels per tile

__device__ __forceinline__ void dw_decode(const DwTiles& t, long long idx, int& cb, int& n, int& th, int& tw) {
  const long long per_cb = (long long)t.N * t.tiles_h * t.tiles_w;
  cb = static_cast<int>(idx / per_cb);
  long long r = idx % per_cb;
  n = static_cast<int>(r / (t.tiles_h * t.tiles_w));
  r %= (t.tiles_h * t.tiles_w);
  th = static_cast<int>(r / t.tiles_w);
  tw = static_cast<int>(r % t.tiles_w);
}

// MODE 0: out[n, ho, wo, c] = sum_{kh,kw} in[n, ho*S + kh - 1, wo*S + kw - 1, c] * w[c, kh, kw]   (flip: w[c, 2-kh, 2-kw])
// MODE 1: dw[c, kh, kw] += sum_{n,ho,wo} dy[n, ho, wo, c] * in[n, ho*S + kh - 1, wo*S + kw - 1, c]
template <int S, int MODE>
__global__ void __launch_bounds__(DW_THREADS, (MODE == 0) ? 2 : 1)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_dy, const float* __restrict__ w,
                  __nv_bfloat16* __restrict__ out, float* __restrict__ dw, const DwTiles t, int flip) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t in_bytes = static_cast<uint32_t>(t.in_rows) * t.in_cols * DW_CB * 2;
  const uint32_t dy_bytes = (MODE == 1) ? static_cast<uint32_t>(t.TH) * t.TW * DW_CB * 2 : 0u;
  const uint32_t stage_bytes = in_bytes + dy_bytes;
  __shared__ __align__(8) unsigned long long bars[2];
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full = [&](int s) { return bar0 + 8 * s; };
  const int tid = threadIdx.x;
  const int cv = tid & 7, lane = tid >> 3;
  const int lr = lane / t.LC, lcg = lane % t.LC;

  if (tid == 0) {
    tma_prefetch_desc(&map_in);
    if (MODE == 1) tma_prefetch_desc(&map_dy);
    mbar_init(full(0), 1);
    mbar_init(full(1), 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](long long idx, int s) {
    int cb, n, th, tw;
    dw_decode(t, idx, cb, n, th, tw);
    mbar_expect_tx(full(s), stage_bytes);
    tma_load_4d(sbase + s * stage_bytes, &map_in, full(s), cb * DW_CB, tw * t.TW * S - 1, th * t.TH * S - 1, n);
    if (MODE == 1) tma_load_4d(sbase + s * stage_bytes + in_bytes, &map_dy, full(s), cb * DW_CB, tw * t.TW, th * t.TH, n);
  };
  if (tid == 0 && static_cast<long long>(blockIdx.x) < t.total) issue(blockIdx.x, 0);

  float wt[9][8];      // MODE 0: filter taps; MODE 1: gradient accumulators
  int cur_cb = -1;
  uint32_t k = 0;
  for (long long idx = blockIdx.x; idx < t.total; idx += gridDim.x, ++k) {
    const int s = k & 1;
    if (tid == 0 && idx + gridDim.x < t.total) issue(idx + gridDim.x, s ^ 1);   // stage s^1 was released by the barrier below
    int cb, n, th, tw;
    dw_decode(t, idx, cb, n, th, tw);
    const int c0 = cb * DW_CB + cv * 8;
    if (cb != cur_cb) {
      if (MODE == 1 && cur_cb >= 0) {   // flush the accumulators of the previous channel block
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v = wt[tap][j];
            v += __shfl_xor_sync(0xffffffffu, v, 8);    // lanes of a warp that share cv (tid & 7): pixel lanes 4w .. 4w+3
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if ((tid & 31) < 8) atomicAdd(dw + (cur_cb * DW_CB + cv * 8 + j) * 9 + tap, v);
          }
      }
      cur_cb = cb;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 8; ++j) wt[tap][j] = (MODE == 0) ? __ldg(w + (c0 + j) * 9 + (flip ? 8 - tap : tap)) : 0.f;
    }
    mbar_wait(full(s), (k >> 1) & 1);
    const uint32_t tile = sbase + s * stage_bytes;
    const int ho = th * t.TH + lr;
    if (lr < t.TH && ho < t.Ho) {
      const int c_begin = lcg * t.SW;
      const int c_end = min(min(c_begin + t.SW, t.TW), t.Wo - tw * t.TW);
      // address of input pixel (r, cc) of the box: ((r * in_cols + cc) * 64 + cv * 8) * 2 bytes
      const uint32_t row0 = tile + (static_cast<uint32_t>(lr * S) * t.in_cols * DW_CB + cv * 8) * 2;
      const uint32_t row_stride = static_cast<uint32_t>(t.in_cols) * DW_CB * 2;
      uint4 win[3][3];
      if (c_begin < c_end) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < (S == 1 ? 2 : 1); ++kw)
            win[kh][kw + (S == 1 ? 1 : 2)] = lds_u4(row0 + kh * row_stride + static_cast<uint32_t>(c_begin * S + kw) * DW_CB * 2);
      }
      for (int c = c_begin; c < c_end; ++c) {
        // slide: stride 1 keeps two columns, stride 2 keeps one
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if (S == 1) {
            win[kh][0] = win[kh][1];
            win[kh][1] = win[kh][2];
            win[kh][2] = lds_u4(row0 + kh * row_stride + static_cast<uint32_t>(c + 2) * DW_CB * 2);
          } else {
            win[kh][0] = win[kh][2];
            win[kh][1] = lds_u4(row0 + kh * row_stride + static_cast<uint32_t>(2 * c + 1) * DW_CB * 2);
            win[kh][2] = lds_u4(row0 + kh * row_stride + static_cast<uint32_t>(2 * c + 2) * DW_CB * 2);
          }
        }
        if (MODE == 0) {
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              float xv[8];
              dw_unpack8(win[kh][kw], xv);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv[j], wt[kh * 3 + kw][j], acc[j]);
            }
          const long long pix = (static_cast<long long>(n) * t.Ho + ho) * t.Wo + tw * t.TW + c;
          *reinterpret_cast<uint4*>(out + pix * t.C + c0) = dw_pack8(acc);
        } else {
          float dv[8];
          dw_unpack8(lds_u4(tile + in_bytes + (static_cast<uint32_t>(lr * t.TW + c) * DW_CB + cv * 8) * 2), dv);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              float xv[8];
              dw_unpack8(win[kh][kw], xv);
#pragma unroll
              for (int j = 0; j < 8; ++j) wt[kh * 3 + kw][j] = fmaf(dv[j], xv[j], wt[kh * 3 + kw][j]);
            }
        }
      }
    }
    __syncthreads();   // everybody is done with stage s: the producer may overwrite it two tiles from now
  }
  if (MODE == 1 && cur_cb >= 0) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = wt[tap][j];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((tid & 31) < 8) atomicAdd(dw + (cur_cb * DW_CB + cv * 8 + j) * 9 + tap, v);
      }
  }
}

// Input gradient of the stride-2 convolution: dx[hi, wi] = sum over the taps (kh, kw) with hi + 1 - kh and wi + 1 - kw even of
// dy[(hi + 1 - kh) / 2, (wi + 1 - kw) / 2] * w[kh, kw] — one, two or four taps depending on the parities of hi and wi.
// Tile = TH x TW dx pixels (both even, tile origin even); the dy box is (TH/2 + 1) x (TW/2 + 1), zero-filled past Ho/Wo.
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_tma_dgrad_s2_kernel(const __grid_constant__ CUtensorMap map_dy, const float* __restrict__ w, __nv_bfloat16* __restrict__ dx,
                           const DwTiles t) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t stage_bytes = static_cast<uint32_t>(t.in_rows) * t.in_cols * DW_CB * 2;
  __shared__ __align__(8) unsigned long long bars[2];
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full = [&](int s) { return bar0 + 8 * s; };
  const int tid = threadIdx.x;
  const int cv = tid & 7, lane = tid >> 3;
  const int lr = lane / t.LC, lcg = lane % t.LC;
  if (tid == 0) {
    tma_prefetch_desc(&map_dy);
    mbar_init(full(0), 1);
    mbar_init(full(1), 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](long long idx, int s) {
    int cb, n, th, tw;
    dw_decode(t, idx, cb, n, th, tw);
    mbar_expect_tx(full(s), stage_bytes);
    tma_load_4d(sbase + s * stage_bytes, &map_dy, full(s), cb * DW_CB, tw * t.TW / 2, th * t.TH / 2, n);
  };
  if (tid == 0 && static_cast<long long>(blockIdx.x) < t.total) issue(blockIdx.x, 0);
  float wt[9][8];
  int cur_cb = -1;
  uint32_t k = 0;
  for (long long idx = blockIdx.x; idx < t.total; idx += gridDim.x, ++k) {
    const int s = k & 1;
    if (tid == 0 && idx + gridDim.x < t.total) issue(idx + gridDim.x, s ^ 1);
    int cb, n, th, tw;
    dw_decode(t, idx, cb, n, th, tw);
    const int c0 = cb * DW_CB + cv * 8;
    if (cb != cur_cb) {
      cur_cb = cb;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 8; ++j) wt[tap][j] = __ldg(w + (c0 + j) * 9 + tap);
    }
    mbar_wait(full(s), (k >> 1) & 1);
    const uint32_t tile = sbase + s * stage_bytes;
    const int hi = th * t.TH + lr;          // here H, W are the dx (= forward input) extents
    if (lr < t.TH && hi < t.H) {
      const int q = lr >> 1;                // dy row of the box for kh = 1 (even hi) / kh = 2 (odd hi)
      const bool odd_row = lr & 1;
      const int c_begin = lcg * t.SW;       // SW is even
      const int c_end = min(min(c_begin + t.SW, t.TW), t.W - tw * t.TW);
      const uint32_t row_stride = static_cast<uint32_t>(t.in_cols) * DW_CB * 2;
      const uint32_t r0 = tile + static_cast<uint32_t>(q) * row_stride + cv * 16;
      for (int c = c_begin; c < c_end; c += 2) {
        const int m = c >> 1;
        float a0[8], a1[8], b0[8], b1[8];   // dy[q, m], dy[q, m+1], dy[q+1, m], dy[q+1, m+1]
        dw_unpack8(lds_u4(r0 + static_cast<uint32_t>(m) * DW_CB * 2), a0);
        dw_unpack8(lds_u4(r0 + static_cast<uint32_t>(m + 1) * DW_CB * 2), a1);
        float e[8], o[8];                   // dx at the even column c and the odd column c + 1
        if (!odd_row) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            e[j] = a0[j] * wt[4][j];
            o[j] = fmaf(a1[j], wt[3][j], a0[j] * wt[5][j]);
          }
        } else {
          dw_unpack8(lds_u4(r0 + row_stride + static_cast<uint32_t>(m) * DW_CB * 2), b0);
          dw_unpack8(lds_u4(r0 + row_stride + static_cast<uint32_t>(m + 1) * DW_CB * 2), b1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            e[j] = fmaf(b0[j], wt[1][j], a0[j] * wt[7][j]);
            o[j] = fmaf(b1[j], wt[0][j], fmaf(b0[j], wt[2][j], fmaf(a1[j], wt[6][j], a0[j] * wt[8][j])));
          }
        }
        const long long pix = (static_cast<long long>(n) * t.H + hi) * t.W + tw * t.TW + c;
        *reinterpret_cast<uint4*>(dx + pix * t.C + c0) = dw_pack8(e);
        if (c + 1 < c_end) *reinterpret_cast<uint4*>(dx + (pix + 1) * t.C + c0) = dw_pack8(o);
      }
    }
    __syncthreads();
  }
}

static int dw_make_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int box_w, int box_h) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[4] = {(uint32_t)DW_CB, (uint32_t)box_w, (uint32_t)box_h, 1};
  return make_tmap_bf16(m, base, 4, dims, str, box, 0);
}

static bool dw_plan(DwTiles& t, int N, int H, int W, int C, int stride) {
  if (C % DW_CB != 0) return false;
  t.N = N; t.H = H; t.W = W; t.C = C;
  t.Ho = (H + 2 - 3) / stride + 1;
  t.Wo = (W + 2 - 3) / stride + 1;
  t.TW = t.Wo < 56 ? t.Wo : 56;
  t.TH = (stride == 1) ? 4 : 2;
  if (t.TH > t.Ho) t.TH = (t.Ho >= 2) ? 2 : 1;
  t.in_cols = (t.TW - 1) * stride + 3;
  t.in_rows = (t.TH - 1) * stride + 3;
  if (t.in_cols > 256 || t.in_rows > 256) return false;
  t.tiles_h = (t.Ho + t.TH - 1) / t.TH;
  t.tiles_w = (t.Wo + t.TW - 1) / t.TW;
  t.cblocks = C / DW_CB;
  t.LC = 32 / t.TH;
  t.SW = (t.TW + t.LC - 1) / t.LC;
  t.total = (long long)t.cblocks * N * t.tiles_h * t.tiles_w;
  return true;
}

template <int S, int MODE>
static int dw_launch(const void* in, const void* dy, const float* w, void* out, float* dwg, const DwTiles& t, int in_H, int in_W, int flip,
                     cudaStream_t stream) {
  CUtensorMap m_in, m_dy;
  int rc = dw_make_map(&m_in, in, t.N, in_H, in_W, t.C, t.in_cols, t.in_rows);
  if (rc) return rc;
  m_dy = m_in;
  if (MODE == 1) {
    rc = dw_make_map(&m_dy, dy, t.N, t.Ho, t.Wo, t.C, t.TW, t.TH);
    if (rc) return rc;
  }
  const size_t stage = (size_t)t.in_rows * t.in_cols * DW_CB * 2 + (MODE == 1 ? (size_t)t.TH * t.TW * DW_CB * 2 : 0);
  const size_t smem = 2 * stage;
  auto kern = dwconv_tma_kernel<S, MODE>;
  VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = (long long)num_sms() * (MODE == 0 ? 2 : 1);   // the weight-gradient kernel keeps 72 accumulators: one CTA per SM
  if (grid > t.total) grid = t.total;
  kern<<<(unsigned)grid, DW_THREADS, smem, stream>>>(m_in, m_dy, w, (__nv_bfloat16*)out, dwg, t, flip);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

// Returns 1 when the TMA path does not cover the geometry (the caller falls back to the generic kernels), 0 on success, <0 on error.
int dwconv_tma_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride, cudaStream_t stream) {
  DwTiles t;
  if (!dw_plan(t, N, H, W, C, stride)) return 1;
  return stride == 1 ? dw_launch<1, 0>(x, nullptr, w, y, nullptr, t, H, W, 0, stream) : dw_launch<2, 0>(x, nullptr, w, y, nullptr, t, H, W, 0, stream);
}
int dwconv_tma_dgrad_s1(const void* dy, const float* w, void* dx, int N, int H, int W, int C, cudaStream_t stream) {
  DwTiles t;
  if (!dw_plan(t, N, H, W, C, 1)) return 1;
  return dw_launch<1, 0>(dy, nullptr, w, dx, nullptr, t, H, W, 1, stream);
}
int dwconv_tma_dgrad_s2(const void* dy, const float* w, void* dx, int N, int H, int W, int C, cudaStream_t stream) {
  if (C % DW_CB != 0) return 1;
  DwTiles t;
  t.N = N; t.H = H; t.W = W; t.C = C;
  t.Ho = (H - 1) / 2 + 1;
  t.Wo = (W - 1) / 2 + 1;
  t.TW = W < 56 ? ((W + 1) & ~1) : 56;
  t.TH = H >= 3 ? 4 : 2;
  t.in_cols = t.TW / 2 + 1;
  t.in_rows = t.TH / 2 + 1;
  t.tiles_h = (H + t.TH - 1) / t.TH;
  t.tiles_w = (W + t.TW - 1) / t.TW;
  t.cblocks = C / DW_CB;
  t.LC = 32 / t.TH;
  t.SW = (((t.TW + t.LC - 1) / t.LC) + 1) & ~1;
  t.total = (long long)t.cblocks * N * t.tiles_h * t.tiles_w;
  CUtensorMap m_dy;
  int rc = dw_make_map(&m_dy, dy, N, t.Ho, t.Wo, C, t.in_cols, t.in_rows);
  if (rc) return rc;
  const size_t smem = 2 * (size_t)t.in_rows * t.in_cols * DW_CB * 2;
  VMAE_CUDA(cudaFuncSetAttribute(dwconv_tma_dgrad_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = (long long)num_sms() * 2;
  if (grid > t.total) grid = t.total;
  dwconv_tma_dgrad_s2_kernel<<<(unsigned)grid, DW_THREADS, smem, stream>>>(m_dy, w, (__nv_bfloat16*)dx, t);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
int dwconv_tma_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int stride, cudaStream_t stream) {
  DwTiles t;
  if (!dw_plan(t, N, H, W, C, stride)) return 1;
  return stride == 1 ? dw_launch<1, 1>(x, dy, nullptr, nullptr, dw, t, H, W, 0, stream) : dw_launch<2, 1>(x, dy, nullptr, nullptr, dw, t, H, W, 0, stream);
}

}  // namespace vmae
