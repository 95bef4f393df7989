// Row-level mask / gather kernels of the latent-space MAE (reference: src/models/tinyvit_mae.py:104-114) and the fused
// softmax cross-entropy of the finetune head (reference: src/train_finetune.py:86-107, nn.CrossEntropyLoss).
// All HBM-bound: one warp per row, 128-bit accesses where the row pitch allows, shuffle reductions, no host sync.
#include "common.cuh"
#include "philox.cuh"

namespace vmae {

constexpr int S2_THREADS = 256;
constexpr int S2_WARPS = S2_THREADS / 32;

static inline int s2_grid(long long rows) {
  long long g = (rows + S2_WARPS - 1) / S2_WARPS;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename T>
__device__ __forceinline__ float ld_f(const T* p) {
  if constexpr (sizeof(T) == 4) return (float)*p; else return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void st_f(T* p, float v) {
  if constexpr (sizeof(T) == 4) *p = v; else *p = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------------------------
// Per-clip compaction of a [B, N] uint8 mask: ascending GLOBAL row indices b*N + n of the entries equal to `want`,
// exactly n_sel per clip (row-major order over (b, n) = the order of torch's boolean gather rec[mask]). A clip with
// fewer hits pads with -1 and raises bit 0 of *flag; a clip with more raises bit 1 (the caller checks it lazily).
// ------------------------------------------------------------------------------------------------------------------
__global__ void mask_row_indices_kernel(const unsigned char* __restrict__ mask, int* __restrict__ idx, int* __restrict__ flag, int N, int n_sel, int want) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int base = 0;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int n = n0 + lane;
    const bool hit = (n < N) && ((mask[(long long)b * N + n] != 0) == (want != 0));
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int pos = base + __popc(bal & ((1u << lane) - 1u));
    if (hit && pos < n_sel) idx[(long long)b * n_sel + pos] = b * N + n;
    base += __popc(bal);
  }
  for (int i = base + lane; i < n_sel; i += 32) idx[(long long)b * n_sel + i] = -1;
  if (lane == 0 && flag && base != n_sel) atomicOr(flag, base < n_sel ? 1 : 2);
}

// out[r, :] = mask[r] ? mask_token : tok[r, :]        (torch.where(mask[..., None], mask_token, tok), tinyvit_mae.py:107-108)
template <typename T>
__global__ void __launch_bounds__(S2_THREADS) mask_token_blend_fwd_kernel(const T* __restrict__ tok, const unsigned char* __restrict__ mask,
                                                                          const float* __restrict__ mask_token, T* __restrict__ out, long long rows, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wstride = (long long)gridDim.x * S2_WARPS;
  for (long long r = (long long)blockIdx.x * S2_WARPS + warp; r < rows; r += wstride) {
    const bool m = mask[r] != 0;
    for (int c = lane; c < D; c += 32) st_f(out + r * D + c, m ? mask_token[c] : ld_f(tok + r * D + c));
  }
}

// dtok[r, :] = mask[r] ? 0 : dout[r, :];   dmask_token[:] += sum over masked rows of dout[r, :]
template <typename T>
__global__ void __launch_bounds__(S2_THREADS) mask_token_blend_bwd_kernel(const T* __restrict__ dout, const unsigned char* __restrict__ mask,
                                                                          T* __restrict__ dtok, float* __restrict__ dmask_token, long long rows, int D) {
  extern __shared__ float s_acc[];   // [S2_WARPS][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < S2_WARPS * D; c += S2_THREADS) s_acc[c] = 0.f;
  __syncthreads();
  float* acc = s_acc + warp * D;
  const long long wstride = (long long)gridDim.x * S2_WARPS;
  for (long long r = (long long)blockIdx.x * S2_WARPS + warp; r < rows; r += wstride) {
    const bool m = mask[r] != 0;
    for (int c = lane; c < D; c += 32) {
      const float g = ld_f(dout + r * D + c);
      if (dtok) st_f(dtok + r * D + c, m ? 0.f : g);
      if (m) acc[c] += g;
    }
  }
  __syncthreads();
  if (dmask_token)
    for (int c = threadIdx.x; c < D; c += S2_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < S2_WARPS; ++w) t += s_acc[w * D + c];
      if (t != 0.f) atomicAdd(dmask_token + c, t);
    }
}

// out[i, :] = src[idx[i], :]   (idx < 0: zero row)
template <typename T>
__global__ void __launch_bounds__(S2_THREADS) gather_rows_kernel(const T* __restrict__ src, const int* __restrict__ idx, T* __restrict__ out,
                                                                 long long n_out, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wstride = (long long)gridDim.x * S2_WARPS;
  for (long long i = (long long)blockIdx.x * S2_WARPS + warp; i < n_out; i += wstride) {
    const long long r = idx[i];
    for (int c = lane; c < D; c += 32) st_f(out + i * D + c, r >= 0 ? ld_f(src + r * D + c) : 0.f);
  }
}

// dst[idx[i], :] = src[i, :]  (dst zero-filled by the caller side of the ABI; idx unique, idx < 0 skipped)
template <typename T>
__global__ void __launch_bounds__(S2_THREADS) scatter_rows_kernel(const T* __restrict__ src, const int* __restrict__ idx, T* __restrict__ dst,
                                                                  long long n_src, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wstride = (long long)gridDim.x * S2_WARPS;
  for (long long i = (long long)blockIdx.x * S2_WARPS + warp; i < n_src; i += wstride) {
    const long long r = idx[i];
    if (r < 0) continue;
    for (int c = lane; c < D; c += 32) dst[r * D + c] = src[i * D + c];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// softmax cross-entropy, mean over rows: one warp per row. fwd: lse[r], row_loss[r] = lse - logit[label];
// finalize: loss = mean(row_loss) in a fixed order (deterministic). bwd: dlogits = (softmax - onehot) * gout / rows.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(S2_THREADS) softmax_ce_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ labels,
                                                                    float* __restrict__ lse, float* __restrict__ row_loss, long long rows, int C) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wstride = (long long)gridDim.x * S2_WARPS;
  for (long long r = (long long)blockIdx.x * S2_WARPS + warp; r < rows; r += wstride) {
    const T* row = logits + r * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, ld_f(row + c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(ld_f(row + c) - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float l = mx + __logf(s);
      const long long y = labels[r];
      lse[r] = l;
      row_loss[r] = (y >= 0 && y < C) ? l - ld_f(row + y) : 0.f;
    }
  }
}

__global__ void softmax_ce_finalize_kernel(const float* __restrict__ row_loss, long long rows, float* __restrict__ loss) {
  __shared__ float red[32];
  float t = 0.f;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) t += row_loss[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    *loss = s / (float)rows;
  }
}

template <typename T>
__global__ void __launch_bounds__(S2_THREADS) softmax_ce_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ labels,
                                                                    const float* __restrict__ lse, const float* __restrict__ gout,
                                                                    T* __restrict__ dlogits, long long rows, int C) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wstride = (long long)gridDim.x * S2_WARPS;
  const float g = gout[0] / (float)rows;
  for (long long r = (long long)blockIdx.x * S2_WARPS + warp; r < rows; r += wstride) {
    const float l = lse[r];
    const long long y = labels[r];
    for (int c = lane; c < C; c += 32) {
      const float p = __expf(ld_f(logits + r * C + c) - l);
      st_f(dlogits + r * C + c, (p - (c == y ? 1.f : 0.f)) * g);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// F.adaptive_avg_pool2d(f, (s, s)) on a channels-last bf16 map (tinyvit_backbone.py:74-81): bin i covers
// [floor(i*H/s), ceil((i+1)*H/s)). out fp32 [N, s*s, C]. One thread per (n, bin, channel-pair).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bin_range(int i, int in, int s, int& lo, int& hi) {
  lo = (i * in) / s;
  hi = ((i + 1) * in + s - 1) / s;
}

__global__ void __launch_bounds__(256) adaptive_avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C,
                                                                   int s) {
  const long long total = (long long)N * s * s * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int bx = (int)(r % s); r /= s;
    const int by = (int)(r % s);
    const int n = (int)(r / s);
    int y0, y1, x0, x1;
    bin_range(by, H, s, y0, y1);
    bin_range(bx, W, s, x0, x1);
    float acc = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int xx = x0; xx < x1; ++xx) acc += __bfloat162float(x[(((long long)n * H + y) * W + xx) * C + c]);
    out[i] = acc / (float)((y1 - y0) * (x1 - x0));
  }
}

__global__ void __launch_bounds__(256) adaptive_avgpool_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dx, int N, int H, int W,
                                                                   int C, int s) {
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int xx = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    float acc = 0.f;
    for (int by = 0; by < s; ++by) {
      int y0, y1;
      bin_range(by, H, s, y0, y1);
      if (y < y0 || y >= y1) continue;
      for (int bx = 0; bx < s; ++bx) {
        int x0, x1;
        bin_range(bx, W, s, x0, x1);
        if (xx < x0 || xx >= x1) continue;
        acc += dout[(((long long)n * s + by) * s + bx) * C + c] / (float)((y1 - y0) * (x1 - x0));
      }
    }
    dx[i] = __float2bfloat16_rn(acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Video embedding of the finetune / privacy path: z[b, :] = mean over the n token rows of clip b (GAP + temporal mean of
// train_finetune.py:19-40 in token form), with the feature perturbation z' = z + sigma * N(0, 1) (feature_noise.py:4-7,
// applied to exactly this pooled embedding at run_privacy.py:299) added in the same kernel. x bf16 [B, n, D] -> out fp32 [B, D].
// Block = (clip, 64-channel slab): 8 lanes x 16 B across the slab, 32 row groups down the tokens, smem tree over the groups.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) meanpool_noise_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, float* __restrict__ eps_out,
                                                             int n, int D, float sigma, unsigned long long seed, unsigned long long offset) {
  __shared__ float red[32][65];
  const int slabs = D / 64;
  const int b = blockIdx.x / slabs, slab = blockIdx.x % slabs;
  const int cl = threadIdx.x & 7, rg = threadIdx.x >> 3;
  const __nv_bfloat16* base = x + (static_cast<long long>(b) * n) * D + slab * 64 + cl * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = rg; t < n; t += 32) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(t) * D));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h[k]);
      acc[2 * k] += f.x;
      acc[2 * k + 1] += f.y;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rg][cl * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) t += red[g][threadIdx.x];
    float z = t / static_cast<float>(n);
    const long long i = static_cast<long long>(b) * D + slab * 64 + threadIdx.x;
    if (sigma > 0.f) {
      const float4 e4 = philox_normal4(seed, offset + static_cast<unsigned long long>(i >> 2));
      const int lane4 = static_cast<int>(i & 3);
      const float e = lane4 == 0 ? e4.x : (lane4 == 1 ? e4.y : (lane4 == 2 ? e4.z : e4.w));
      z = __fmaf_rn(sigma, e, z);
      if (eps_out) eps_out[i] = e;
    }
    out[i] = z;
  }
}

}  // namespace vmae

using namespace vmae;

extern "C" int vmae_meanpool_noise(const void* x, float* out, float* eps_out, int B, int n, int D, float sigma, unsigned long long seed,
                                   unsigned long long offset, void* stream) {
  VMAE_CHECK_ARG(x && out && B > 0 && n > 0 && D > 0 && D % 64 == 0 && sigma >= 0.f, "vmae_meanpool_noise: bad args (D %% 64 == 0, sigma >= 0)");
  meanpool_noise_kernel<<<B * (D / 64), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, eps_out, n, D, sigma, seed, offset);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_adaptive_avgpool_fwd(const void* x, float* out, int N, int H, int W, int C, int s, void* stream) {
  VMAE_CHECK_ARG(x && out && N > 0 && H > 0 && W > 0 && C > 0 && s > 0 && s <= H && s <= W, "vmae_adaptive_avgpool_fwd: bad args");
  const long long total = (long long)N * s * s * C;
  long long grid = (total + 255) / 256;
  if (grid > (long long)num_sms() * 16) grid = (long long)num_sms() * 16;
  adaptive_avgpool_fwd_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, N, H, W, C, s);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_adaptive_avgpool_bwd(const float* dout, void* dx, int N, int H, int W, int C, int s, void* stream) {
  VMAE_CHECK_ARG(dout && dx && N > 0 && H > 0 && W > 0 && C > 0 && s > 0 && s <= H && s <= W, "vmae_adaptive_avgpool_bwd: bad args");
  const long long total = (long long)N * H * W * C;
  long long grid = (total + 255) / 256;
  if (grid > (long long)num_sms() * 16) grid = (long long)num_sms() * 16;
  adaptive_avgpool_bwd_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(dout, (__nv_bfloat16*)dx, N, H, W, C, s);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_mask_row_indices(const unsigned char* mask, int* idx, int* flag, int B, int N, int n_sel, int want, void* stream) {
  VMAE_CHECK_ARG(mask && idx && B > 0 && N > 0 && n_sel > 0 && n_sel <= N, "vmae_mask_row_indices: bad args");
  VMAE_CHECK_ARG((long long)B * N < (1ll << 31), "vmae_mask_row_indices: B*N must fit an int32 row index");
  mask_row_indices_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(mask, idx, flag, N, n_sel, want);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

#define S2_DISPATCH(dtype, CALL_BF16, CALL_F32)                                    \
  if ((dtype) == VMAE_DTYPE_BF16) { CALL_BF16; }                                    \
  else if ((dtype) == VMAE_DTYPE_F32) { CALL_F32; }                                 \
  else return fail(VMAE_ERR_BAD_ARG, "bad dtype %d", (int)(dtype));

extern "C" int vmae_mask_token_blend_fwd(const void* tok, int dtype, const unsigned char* mask, const float* mask_token, void* out, long long rows,
                                         int D, void* stream) {
  VMAE_CHECK_ARG(tok && mask && mask_token && out && rows > 0 && D > 0, "vmae_mask_token_blend_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  S2_DISPATCH(dtype,
              (mask_token_blend_fwd_kernel<__nv_bfloat16><<<s2_grid(rows), S2_THREADS, 0, s>>>((const __nv_bfloat16*)tok, mask, mask_token, (__nv_bfloat16*)out, rows, D)),
              (mask_token_blend_fwd_kernel<float><<<s2_grid(rows), S2_THREADS, 0, s>>>((const float*)tok, mask, mask_token, (float*)out, rows, D)));
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_mask_token_blend_bwd(const void* dout, int dtype, const unsigned char* mask, void* dtok, float* dmask_token, long long rows, int D,
                                         void* stream) {
  VMAE_CHECK_ARG(dout && mask && rows > 0 && D > 0 && D <= 4096, "vmae_mask_token_blend_bwd: bad args (D <= 4096)");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)S2_WARPS * D * sizeof(float);
  int grid = s2_grid(rows);
  if (grid > num_sms()) grid = num_sms();   // one atomic per column per CTA
  if (smem > 48 * 1024) {
    VMAE_CUDA(cudaFuncSetAttribute(mask_token_blend_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VMAE_CUDA(cudaFuncSetAttribute(mask_token_blend_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  S2_DISPATCH(dtype,
              (mask_token_blend_bwd_kernel<__nv_bfloat16><<<grid, S2_THREADS, smem, s>>>((const __nv_bfloat16*)dout, mask, (__nv_bfloat16*)dtok, dmask_token, rows, D)),
              (mask_token_blend_bwd_kernel<float><<<grid, S2_THREADS, smem, s>>>((const float*)dout, mask, (float*)dtok, dmask_token, rows, D)));
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_gather_rows(const void* src, int dtype, const int* idx, void* out, long long n_out, int D, void* stream) {
  VMAE_CHECK_ARG(src && idx && out && n_out > 0 && D > 0, "vmae_gather_rows: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  S2_DISPATCH(dtype,
              (gather_rows_kernel<__nv_bfloat16><<<s2_grid(n_out), S2_THREADS, 0, s>>>((const __nv_bfloat16*)src, idx, (__nv_bfloat16*)out, n_out, D)),
              (gather_rows_kernel<float><<<s2_grid(n_out), S2_THREADS, 0, s>>>((const float*)src, idx, (float*)out, n_out, D)));
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_scatter_rows(const void* src, int dtype, const int* idx, void* dst, long long n_src, long long n_dst_rows, int D, void* stream) {
  VMAE_CHECK_ARG(src && idx && dst && n_src > 0 && n_dst_rows > 0 && D > 0, "vmae_scatter_rows: bad args");
  VMAE_CHECK_ARG(dtype == VMAE_DTYPE_BF16 || dtype == VMAE_DTYPE_F32, "vmae_scatter_rows: bad dtype %d", dtype);
  cudaStream_t s = (cudaStream_t)stream;
  VMAE_CUDA(cudaMemsetAsync(dst, 0, (size_t)n_dst_rows * D * (dtype == VMAE_DTYPE_F32 ? 4 : 2), s));
  S2_DISPATCH(dtype,
              (scatter_rows_kernel<__nv_bfloat16><<<s2_grid(n_src), S2_THREADS, 0, s>>>((const __nv_bfloat16*)src, idx, (__nv_bfloat16*)dst, n_src, D)),
              (scatter_rows_kernel<float><<<s2_grid(n_src), S2_THREADS, 0, s>>>((const float*)src, idx, (float*)dst, n_src, D)));
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_softmax_ce_fwd(const void* logits, int dtype, const long long* labels, float* lse, float* row_loss, float* loss, long long rows,
                                   int C, void* stream) {
  VMAE_CHECK_ARG(logits && labels && lse && row_loss && loss && rows > 0 && C > 0, "vmae_softmax_ce_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  S2_DISPATCH(dtype,
              (softmax_ce_fwd_kernel<__nv_bfloat16><<<s2_grid(rows), S2_THREADS, 0, s>>>((const __nv_bfloat16*)logits, labels, lse, row_loss, rows, C)),
              (softmax_ce_fwd_kernel<float><<<s2_grid(rows), S2_THREADS, 0, s>>>((const float*)logits, labels, lse, row_loss, rows, C)));
  VMAE_LAUNCH_CHECK();
  softmax_ce_finalize_kernel<<<1, 256, 0, s>>>(row_loss, rows, loss);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_softmax_ce_bwd(const void* logits, int dtype, const long long* labels, const float* lse, const float* gout, void* dlogits,
                                   long long rows, int C, void* stream) {
  VMAE_CHECK_ARG(logits && labels && lse && gout && dlogits && rows > 0 && C > 0, "vmae_softmax_ce_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  S2_DISPATCH(dtype,
              (softmax_ce_bwd_kernel<__nv_bfloat16><<<s2_grid(rows), S2_THREADS, 0, s>>>((const __nv_bfloat16*)logits, labels, lse, gout, (__nv_bfloat16*)dlogits, rows, C)),
              (softmax_ce_bwd_kernel<float><<<s2_grid(rows), S2_THREADS, 0, s>>>((const float*)logits, labels, lse, gout, (float*)dlogits, rows, C)));
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
