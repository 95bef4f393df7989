// K7 — TinyViT convolution kit on channels-last (NHWC) bf16 feature maps, rows = pixels:
//   train-mode BatchNorm (batch statistics over N*H*W, eps 1e-5, momentum 0.1) with optional fused exact GELU and residual,
//   depthwise 3x3 convolution (stride 1/2, pad 1) forward / dgrad / wgrad,
//   squeeze-excite global average pool and channel gating, and the 3-channel stem convolution.
// Replaces cuDNN/ATen ops behind Conv2d_BN, MBConv, SELayer, PatchEmbed (tiny_vit.py:12-72). All kernels are HBM-bound:
// a thread owns 8 consecutive channels (one 16-byte bf16 vector) so every access is a full, coalesced row segment.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/vmae.h"

namespace vmae {

constexpr int CV_THREADS = 256;
constexpr int CV_WARPS = CV_THREADS / 32;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = __bfloat1622float2(p[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    __nv_bfloat162 v = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
    w[t] = *reinterpret_cast<uint32_t*>(&v);
  }
  return u;
}
// same polynomial Phi as the GEMM epilogue (exact-erf GELU to 1.6e-5)
__device__ __forceinline__ float cv_normal_cdf(float x) {
  const float xc = fminf(fmaxf(x, -4.2f), 4.2f);
  const float u = xc * xc;
  float p = 5.3164563029728384e-11f;
  p = fmaf(p, u, -5.110036394029294e-09f);
  p = fmaf(p, u, 2.1774354763692827e-07f);
  p = fmaf(p, u, -5.478650564327836e-06f);
  p = fmaf(p, u, 9.182147914543748e-05f);
  p = fmaf(p, u, -0.0010986902052536607f);
  p = fmaf(p, u, 0.00978302676230669f);
  p = fmaf(p, u, -0.0662931576371193f);
  p = fmaf(p, u, 0.3988766670227051f);
  return fmaf(xc, p, 0.5f);
}
__device__ __forceinline__ float cv_gelu(float x) { return x * cv_normal_cdf(x); }
__device__ __forceinline__ float cv_gelu_grad(float x) {
  return fmaf(x, 0.39894228040143268f * __expf(-0.5f * x * x), cv_normal_cdf(x));
}

// gelu'(z) for a pair of lanes on packed FFMA2 (half the issue slots of the scalar form; same polynomial, pdf by one ex2)
__device__ __forceinline__ void cv_gelu_grad2(float z0, float z1, float& d0, float& d1) {
  const f32x2 xc2 = pk2(fminf(fmaxf(z0, -4.2f), 4.2f), fminf(fmaxf(z1, -4.2f), 4.2f));
  const f32x2 u2 = mul2(xc2, xc2);
  f32x2 q2 = fma2(pk2(5.3164563029728384e-11f), u2, pk2(-5.110036394029294e-09f));
  q2 = fma2(q2, u2, pk2(2.1774354763692827e-07f));
  q2 = fma2(q2, u2, pk2(-5.478650564327836e-06f));
  q2 = fma2(q2, u2, pk2(9.182147914543748e-05f));
  q2 = fma2(q2, u2, pk2(-0.0010986902052536607f));
  q2 = fma2(q2, u2, pk2(0.00978302676230669f));
  q2 = fma2(q2, u2, pk2(-0.0662931576371193f));
  q2 = fma2(q2, u2, pk2(0.3988766670227051f));
  const f32x2 cdf2 = fma2(xc2, q2, pk2(0.5f));
  float e0, e1;
  up2(fma2(u2, pk2(-0.72134752044448170f), pk2(-1.3257480647361592f)), e0, e1);   // log2(pdf)
  float p0, p1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e1));
  up2(fma2(xc2, pk2(p0, p1), cdf2), d0, d1);
}

// ------------------------------------------------------------------------------------------------------------------
// BatchNorm kernels. Thread block = (cvec channel-vectors) x (rows): threadIdx.x owns 8 consecutive channels for the whole
// launch (per-channel parameters live in registers), threadIdx.y walks rows with a grid stride. A row of C channels is
// cvec consecutive 16-byte vectors, so every warp access is contiguous for any C (48 ... 768).
// MODE 0: acc[0][c] += sum x,  acc[1][c] += sum x^2                                   (statistics)
// MODE 1: dz = dy * (act ? gelu'(z) : 1), z = xhat*gamma+beta; acc[0][c] += sum dz, acc[1][c] += sum dz*xhat (backward)
// ------------------------------------------------------------------------------------------------------------------
constexpr int BN_MAX_ROWS_PER_BLOCK = 64;
constexpr int BN_U = 2;   // rows per pipeline step

template <int MODE>
__global__ void __launch_bounds__(CV_THREADS) bn_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                                               float* __restrict__ acc, long long M, int C, int cv0) {
  extern __shared__ float bn_red[];   // [2][blockDim.y][blockDim.x * 8]
  const int cv = cv0 + threadIdx.x;   // channel vector owned by this thread
  const int c0 = cv * 8;
  const int nx = blockDim.x, ny = blockDim.y;
  float a0[8], a1[8], mu[8], rs[8], g[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; mu[j] = 0.f; rs[j] = 1.f; g[j] = 1.f; b[j] = 0.f; }
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mean[c0 + j]; rs[j] = rstd[c0 + j]; g[j] = gamma[c0 + j]; b[j] = beta[c0 + j]; }
  }
  // software pipeline: BN_U rows per step, and the next BN_U rows are requested before the current ones are consumed
  // (2 * BN_U rows of HBM traffic in flight per thread; the GELU-gradient arithmetic overlaps the loads)
  const long long rstep = (long long)gridDim.x * ny;
  long long r = (long long)blockIdx.x * ny + threadIdx.y;
  uint4 nxt_x[BN_U], nxt_d[BN_U];
#pragma unroll
  for (int u = 0; u < BN_U; ++u) {
    nxt_x[u] = make_uint4(0, 0, 0, 0);
    nxt_d[u] = make_uint4(0, 0, 0, 0);
    if (r + u * rstep < M) {
      nxt_x[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * rstep) * C + c0));
      if (MODE == 1) nxt_d[u] = __ldg(reinterpret_cast<const uint4*>(dy + (r + u * rstep) * C + c0));
    }
  }
  for (; r < M; r += BN_U * rstep) {
    uint4 cur_x[BN_U], cur_d[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) { cur_x[u] = nxt_x[u]; cur_d[u] = nxt_d[u]; }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long rn = r + (BN_U + u) * rstep;
      if (rn < M) {
        nxt_x[u] = __ldg(reinterpret_cast<const uint4*>(x + rn * C + c0));
        if (MODE == 1) nxt_d[u] = __ldg(reinterpret_cast<const uint4*>(dy + rn * C + c0));
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      if (r + u * rstep >= M) break;
      float xv[8];
      unpack8(cur_x[u], xv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += xv[j]; a1[j] = fmaf(xv[j], xv[j], a1[j]); }
      } else {
        float dv[8];
        unpack8(cur_d[u], dv);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float xh0 = (xv[j] - mu[j]) * rs[j], xh1 = (xv[j + 1] - mu[j + 1]) * rs[j + 1];
          float dz0 = dv[j], dz1 = dv[j + 1];
          if (act) {
            float d0, d1;
            cv_gelu_grad2(fmaf(xh0, g[j], b[j]), fmaf(xh1, g[j + 1], b[j + 1]), d0, d1);
            dz0 *= d0;
            dz1 *= d1;
          }
          a0[j] += dz0;
          a0[j + 1] += dz1;
          a1[j] = fmaf(dz0, xh0, a1[j]);
          a1[j + 1] = fmaf(dz1, xh1, a1[j + 1]);
        }
      }
    }
  }
  float* r0 = bn_red + (size_t)threadIdx.y * nx * 8 + threadIdx.x * 8;
  float* r1 = r0 + (size_t)ny * nx * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) { r0[j] = a0[j]; r1[j] = a1[j]; }
  __syncthreads();
  const int tid = threadIdx.y * nx + threadIdx.x;
  for (int c = tid; c < nx * 8; c += nx * ny) {
    float t0 = 0.f, t1 = 0.f;
    for (int q = 0; q < ny; ++q) { t0 += bn_red[(size_t)q * nx * 8 + c]; t1 += bn_red[(size_t)(ny + q) * nx * 8 + c]; }
    atomicAdd(acc + cv0 * 8 + c, t0);
    atomicAdd(acc + C + cv0 * 8 + c, t1);
  }
}

// mean / rstd from the sums, running-statistics update (momentum, unbiased variance) — one small block
__global__ void bn_finalize_kernel(const float* __restrict__ acc, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, long long M, int C, float eps,
                                   float momentum) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const float m = acc[c] / (float)M;
    const float var = fmaxf(acc[C + c] / (float)M - m * m, 0.f);
    mean[c] = m;
    rstd[c] = rsqrtf(var + eps);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
      const float unbiased = (M > 1) ? var * ((float)M / (float)(M - 1)) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
}

// y = act((x - mean) * rstd * gamma + beta) (+ residual)
__global__ void __launch_bounds__(CV_THREADS) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                              const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const __nv_bfloat16* __restrict__ residual,
                                                              __nv_bfloat16* __restrict__ y, long long M, int C, int act, int cv0) {
  const int c0 = (cv0 + threadIdx.x) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = rstd[c0 + j] * gamma[c0 + j];
    sh[j] = fmaf(-mean[c0 + j], sc[j], beta[c0 + j]);
  }
  const long long rstep = (long long)gridDim.x * blockDim.y;
  long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  uint4 nxt_x = make_uint4(0, 0, 0, 0), nxt_r = make_uint4(0, 0, 0, 0);
  if (r < M) {
    nxt_x = __ldg(reinterpret_cast<const uint4*>(x + r * C + c0));
    if (residual) nxt_r = __ldg(reinterpret_cast<const uint4*>(residual + r * C + c0));
  }
  for (; r < M; r += rstep) {
    float xv[8], o[8];
    unpack8(nxt_x, xv);
    const uint4 cur_r = nxt_r;
    if (r + rstep < M) {   // next row in flight while this one is normalised / activated
      nxt_x = __ldg(reinterpret_cast<const uint4*>(x + (r + rstep) * C + c0));
      if (residual) nxt_r = __ldg(reinterpret_cast<const uint4*>(residual + (r + rstep) * C + c0));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(xv[j], sc[j], sh[j]);
      o[j] = act ? cv_gelu(z) : z;
    }
    if (residual) {
      float rv[8];
      unpack8(cur_r, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += rv[j];
    }
    *reinterpret_cast<uint4*>(y + r * C + c0) = pack8(o);
  }
}

// dx = gamma * rstd * (dz - sum(dz)/M - xhat * sum(dz*xhat)/M); dgamma += sum(dz*xhat), dbeta += sum(dz) (done by block 0)
__global__ void __launch_bounds__(CV_THREADS) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ acc, __nv_bfloat16* __restrict__ dx,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C,
                                                                  int act, int cv0) {
  const int c0 = (cv0 + threadIdx.x) * 8;
  const float inv_m = 1.0f / (float)M;
  float mu[8], rs[8], g[8], b[8], s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = mean[c0 + j]; rs[j] = rstd[c0 + j]; g[j] = gamma[c0 + j]; b[j] = beta[c0 + j];
    s0[j] = acc[c0 + j] * inv_m; s1[j] = acc[C + c0 + j] * inv_m;
  }
  if (blockIdx.x == 0 && threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (dbeta) atomicAdd(dbeta + c0 + j, acc[c0 + j]);
      if (dgamma) atomicAdd(dgamma + c0 + j, acc[C + c0 + j]);
    }
  }
  const long long rstep = (long long)gridDim.x * blockDim.y;
  long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  uint4 nxt_x[BN_U], nxt_d[BN_U];
#pragma unroll
  for (int u = 0; u < BN_U; ++u) {
    nxt_x[u] = make_uint4(0, 0, 0, 0);
    nxt_d[u] = make_uint4(0, 0, 0, 0);
    if (r + u * rstep < M) {
      nxt_x[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * rstep) * C + c0));
      nxt_d[u] = __ldg(reinterpret_cast<const uint4*>(dy + (r + u * rstep) * C + c0));
    }
  }
  for (; r < M; r += BN_U * rstep) {
    uint4 cur_x[BN_U], cur_d[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) { cur_x[u] = nxt_x[u]; cur_d[u] = nxt_d[u]; }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long rn = r + (BN_U + u) * rstep;
      if (rn < M) {
        nxt_x[u] = __ldg(reinterpret_cast<const uint4*>(x + rn * C + c0));
        nxt_d[u] = __ldg(reinterpret_cast<const uint4*>(dy + rn * C + c0));
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long ru = r + u * rstep;
      if (ru >= M) break;
      float xv[8], dv[8], o[8];
      unpack8(cur_x[u], xv);
      unpack8(cur_d[u], dv);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float xh0 = (xv[j] - mu[j]) * rs[j], xh1 = (xv[j + 1] - mu[j + 1]) * rs[j + 1];
        float dz0 = dv[j], dz1 = dv[j + 1];
        if (act) {
          float d0, d1;
          cv_gelu_grad2(fmaf(xh0, g[j], b[j]), fmaf(xh1, g[j + 1], b[j + 1]), d0, d1);
          dz0 *= d0;
          dz1 *= d1;
        }
        o[j] = g[j] * rs[j] * (dz0 - s0[j] - xh0 * s1[j]);
        o[j + 1] = g[j + 1] * rs[j + 1] * (dz1 - s0[j + 1] - xh1 * s1[j + 1]);
      }
      *reinterpret_cast<uint4*>(dx + ru * C + c0) = pack8(o);
    }
  }
}

// launch geometry shared by the BatchNorm kernels: the channel vectors are split into chunks of <= 128 (blockDim.x),
// blockDim.y rows per block so that a block has ~256 threads; one launch per chunk (1 for C <= 1024)
struct BnLaunch {
  dim3 block;
  int grid;
  size_t smem;
};
static inline BnLaunch bn_launch(long long M, int nx) {
  BnLaunch l;
  int ny = CV_THREADS / nx;
  if (ny < 1) ny = 1;
  if (ny > BN_MAX_ROWS_PER_BLOCK) ny = BN_MAX_ROWS_PER_BLOCK;
  l.block = dim3(nx, ny);
  long long g = (M + ny - 1) / ny;
  const long long cap = (long long)num_sms() * 8;
  l.grid = (int)(g > cap ? cap : (g < 1 ? 1 : g));
  l.smem = (size_t)2 * ny * nx * 8 * sizeof(float);
  return l;
}

static inline int ew_blocks(long long items) {
  long long g = (items + CV_THREADS - 1) / CV_THREADS;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace vmae

using namespace vmae;

extern "C" int vmae_bn_train_fwd(const void* x, const float* gamma, const float* beta, const void* residual, void* y, float* mean, float* rstd,
                                 float* running_mean, float* running_var, float* workspace, long long M, int C, float eps, float momentum,
                                 int act, void* stream) {
  VMAE_CHECK_ARG(x && gamma && beta && y && mean && rstd && workspace, "vmae_bn_train_fwd: null pointer");
  VMAE_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0, "vmae_bn_train_fwd: C=%d must be a positive multiple of 8", C);
  cudaStream_t s = (cudaStream_t)stream;
  VMAE_CUDA(cudaMemsetAsync(workspace, 0, 2 * (size_t)C * sizeof(float), s));
  const int cvec = C / 8;
  for (int cv0 = 0; cv0 < cvec; cv0 += 128) {
    const int nx = (cvec - cv0 < 128) ? cvec - cv0 : 128;
    const BnLaunch l = bn_launch(M, nx);
    bn_reduce_kernel<0><<<l.grid, l.block, l.smem, s>>>((const __nv_bfloat16*)x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, workspace, M, C, cv0);
    VMAE_LAUNCH_CHECK();
  }
  bn_finalize_kernel<<<1, 256, 0, s>>>(workspace, mean, rstd, running_mean, running_var, M, C, eps, momentum);
  VMAE_LAUNCH_CHECK();
  for (int cv0 = 0; cv0 < cvec; cv0 += 128) {
    const int nx = (cvec - cv0 < 128) ? cvec - cv0 : 128;
    const BnLaunch l = bn_launch(M, nx);
    bn_apply_kernel<<<l.grid, l.block, 0, s>>>((const __nv_bfloat16*)x, mean, rstd, gamma, beta, (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, M, C,
                                               act, cv0);
    VMAE_LAUNCH_CHECK();
  }
  return VMAE_OK;
}

extern "C" int vmae_bn_apply(const void* x, const float* gamma, const float* beta, const float* mean, const float* rstd, const void* residual,
                             void* y, long long M, int C, int act, void* stream) {
  VMAE_CHECK_ARG(x && gamma && beta && mean && rstd && y && M > 0 && C > 0 && C % 8 == 0, "vmae_bn_apply: bad args");
  const int cvec = C / 8;
  for (int cv0 = 0; cv0 < cvec; cv0 += 128) {
    const int nx = (cvec - cv0 < 128) ? cvec - cv0 : 128;
    const BnLaunch l = bn_launch(M, nx);
    bn_apply_kernel<<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, mean, rstd, gamma, beta, (const __nv_bfloat16*)residual,
                                                                  (__nv_bfloat16*)y, M, C, act, cv0);
    VMAE_LAUNCH_CHECK();
  }
  return VMAE_OK;
}

extern "C" int vmae_bn_train_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const float* mean, const float* rstd,
                                 void* dx, float* dgamma, float* dbeta, float* workspace, long long M, int C, int act, void* stream) {
  VMAE_CHECK_ARG(x && dy && gamma && beta && mean && rstd && dx && workspace, "vmae_bn_train_bwd: null pointer");
  VMAE_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0, "vmae_bn_train_bwd: C=%d must be a positive multiple of 8", C);
  cudaStream_t s = (cudaStream_t)stream;
  VMAE_CUDA(cudaMemsetAsync(workspace, 0, 2 * (size_t)C * sizeof(float), s));
  const int cvec = C / 8;
  for (int cv0 = 0; cv0 < cvec; cv0 += 128) {
    const int nx = (cvec - cv0 < 128) ? cvec - cv0 : 128;
    const BnLaunch l = bn_launch(M, nx);
    bn_reduce_kernel<1><<<l.grid, l.block, l.smem, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, mean, rstd, gamma, beta, act, workspace, M, C,
                                                        cv0);
    VMAE_LAUNCH_CHECK();
  }
  for (int cv0 = 0; cv0 < cvec; cv0 += 128) {
    const int nx = (cvec - cv0 < 128) ? cvec - cv0 : 128;
    const BnLaunch l = bn_launch(M, nx);
    bn_bwd_apply_kernel<<<l.grid, l.block, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, mean, rstd, gamma, beta, workspace,
                                                   (__nv_bfloat16*)dx, dgamma, dbeta, M, C, act, cv0);
    VMAE_LAUNCH_CHECK();
  }
  return VMAE_OK;
}

// ==================================================================================================================
// depthwise 3x3 (pad 1, stride 1 or 2) on NHWC bf16. block = 16 channel-vectors x 16 pixels; a thread keeps its 8x9
// filter taps in registers and walks pixels with a grid stride. weight: fp32 [C, 9] (= nn.Conv2d weight [C,1,3,3]).
// ==================================================================================================================
namespace vmae {

// TMA-tiled kernels (dwconv_tma.cu); return 1 when the geometry is not covered (C % 64 != 0), in which case the generic
// kernels below run instead
int dwconv_tma_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride, cudaStream_t stream);
int dwconv_tma_dgrad_s1(const void* dy, const float* w, void* dx, int N, int H, int W, int C, cudaStream_t stream);
int dwconv_tma_dgrad_s2(const void* dy, const float* w, void* dx, int N, int H, int W, int C, cudaStream_t stream);
int dwconv_tma_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int stride, cudaStream_t stream);

struct DwGeom {
  int N, H, W, C, Ho, Wo, stride;
};

__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                         __nv_bfloat16* __restrict__ y, DwGeom g) {
  const int cv = blockIdx.x * 16 + threadIdx.x;   // channel vector
  if (cv * 8 >= g.C) return;
  const int c0 = cv * 8;
  float wt[9][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t][j] = __ldg(w + (c0 + j) * 9 + t);
  const int npix = g.N * g.Ho * g.Wo;   // < 2^31 (checked on the host): 32-bit index arithmetic, 64-bit only for byte offsets
  for (int p = blockIdx.y * 16 + threadIdx.y; p < npix; p += gridDim.y * 16) {
    const int wo = p % g.Wo;
    const int t_ = p / g.Wo;
    const int ho = t_ % g.Ho;
    const int n = t_ / g.Ho;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * g.stride + kh - 1;
      if (hi < 0 || hi >= g.H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * g.stride + kw - 1;
        if (wi < 0 || wi >= g.W) continue;
        float xv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * g.H + hi) * g.W + wi) * g.C + c0)), xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv[j], wt[kh * 3 + kw][j], acc[j]);
      }
    }
    *reinterpret_cast<uint4*>(y + (long long)p * g.C + c0) = pack8(acc);
  }
}

__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                                                           __nv_bfloat16* __restrict__ dx, DwGeom g) {
  const int cv = blockIdx.x * 16 + threadIdx.x;
  if (cv * 8 >= g.C) return;
  const int c0 = cv * 8;
  float wt[9][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t][j] = __ldg(w + (c0 + j) * 9 + t);
  const int npix = g.N * g.H * g.W;
  for (int p = blockIdx.y * 16 + threadIdx.y; p < npix; p += gridDim.y * 16) {
    const int wi = p % g.W;
    const int t_ = p / g.W;
    const int hi = t_ % g.H;
    const int n = t_ / g.H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = hi + 1 - kh;               // = ho * stride
      if (th < 0 || th % g.stride != 0) continue;
      const int ho = th / g.stride;
      if (ho >= g.Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = wi + 1 - kw;
        if (tw < 0 || tw % g.stride != 0) continue;
        const int wo = tw / g.stride;
        if (wo >= g.Wo) continue;
        float dv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (((long long)n * g.Ho + ho) * g.Wo + wo) * g.C + c0)), dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(dv[j], wt[kh * 3 + kw][j], acc[j]);
      }
    }
    *reinterpret_cast<uint4*>(dx + (long long)p * g.C + c0) = pack8(acc);
  }
}

// dw[c, tap] += sum over output pixels of dy * x(tap); 72 register accumulators per thread, smem reduce over the 16 pixel
// lanes of the block, one atomicAdd per (channel, tap) per block.
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                           float* __restrict__ dw, DwGeom g) {
  __shared__ float red[16][16][9];   // [pixel lane][channel vector][tap] for one of the 8 channels at a time
  const int cv = blockIdx.x * 16 + threadIdx.x;
  const bool live = cv * 8 < g.C;
  const int c0 = cv * 8;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int npix = g.N * g.Ho * g.Wo;
  if (live) {
    for (int p = blockIdx.y * 16 + threadIdx.y; p < npix; p += gridDim.y * 16) {
      const int wo = p % g.Wo;
      const int t_ = p / g.Wo;
      const int ho = t_ % g.Ho;
      const int n = t_ / g.Ho;
      float dv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (long long)p * g.C + c0)), dv);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho * g.stride + kh - 1;
        if (hi < 0 || hi >= g.H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int wi = wo * g.stride + kw - 1;
          if (wi < 0 || wi >= g.W) continue;
          float xv[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * g.H + hi) * g.W + wi) * g.C + c0)), xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[kh * 3 + kw][j] = fmaf(dv[j], xv[j], acc[kh * 3 + kw][j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int t = 0; t < 9; ++t) red[threadIdx.y][threadIdx.x][t] = acc[t][j];
    __syncthreads();
    if (threadIdx.y == 0 && live) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) s += red[q][threadIdx.x][t];
        atomicAdd(dw + (c0 + j) * 9 + t, s);
      }
    }
    __syncthreads();
  }
}

// ==================================================================================================================
// squeeze-excite: per-image global average pool over HW rows, and channel gating. MODE 1 of the pool computes the gate
// gradient sum_p dy * x instead.
// ==================================================================================================================
template <int MODE>
__global__ void __launch_bounds__(CV_THREADS) se_pool_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                             float* __restrict__ out, int HW, int C, float scale) {
  __shared__ float red[CV_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const long long base = (long long)blockIdx.y * HW;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (c0 < C) {
    for (int p = warp; p < HW; p += CV_WARPS) {
      float xv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (base + p) * C + c0)), xv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += xv[j];
      } else {
        float dv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (base + p) * C + c0)), dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaf(xv[j], dv[j], a[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = a[j];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < CV_WARPS; ++w) t += red[w][c];
    out[(long long)blockIdx.y * C + blockIdx.x * 256 + c] = t * scale;
  }
}

// y[n,p,c] = a[n,p,c] * gate[n,c] (+ add[n,c])        (forward gating; backward dx = dy * gate + dpooled / HW)
__global__ void __launch_bounds__(CV_THREADS) se_scale_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ gate,
                                                              const float* __restrict__ add, __nv_bfloat16* __restrict__ y, long long M, int HW,
                                                              int C, float add_scale) {
  const int cvec = C / 8;
  const long long total = M * cvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c0 = (int)(i % cvec) * 8;
    const long long n = (i / cvec) / HW;
    float v[8], o[8];
    if (a) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(a) + i), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[j] * __ldg(gate + n * C + c0 + j);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    if (add) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(__ldg(add + n * C + c0 + j), add_scale, o[j]);
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(o);
  }
}

// ==================================================================================================================
// im2col / col2im for the 3x3 stem convolutions (the GEMM itself is the tcgen05 kernel)
// ==================================================================================================================
// NHWC bf16 -> cols [N*Ho*Wo, 9*C], tap-major (kh, kw, c); C % 8 == 0
__global__ void __launch_bounds__(CV_THREADS) im2col3x3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, DwGeom g) {
  const int cvec = g.C / 8;
  const long long total = (long long)g.N * g.Ho * g.Wo * 9 * cvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int cv = (int)(i % cvec);
    const int tap = (int)((i / cvec) % 9);
    const long long p = i / ((long long)cvec * 9);
    const int wo = (int)(p % g.Wo);
    const int ho = (int)((p / g.Wo) % g.Ho);
    const int n = (int)(p / ((long long)g.Wo * g.Ho));
    const int hi = ho * g.stride + tap / 3 - 1, wi = wo * g.stride + tap % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * g.H + hi) * g.W + wi) * g.C + cv * 8));
    reinterpret_cast<uint4*>(col)[i] = v;
  }
}
// dx[n,hi,wi,c] = sum over taps of dcol[(n,ho,wo), tap, c] with (ho,wo) the output pixels that read (hi,wi) through `tap`
__global__ void __launch_bounds__(CV_THREADS) col2im3x3_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, DwGeom g) {
  const int cvec = g.C / 8;
  const long long total = (long long)g.N * g.H * g.W * cvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int cv = (int)(i % cvec);
    const long long p = i / cvec;
    const int wi = (int)(p % g.W);
    const int hi = (int)((p / g.W) % g.H);
    const int n = (int)(p / ((long long)g.W * g.H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int th = hi + 1 - tap / 3, tw = wi + 1 - tap % 3;
      if (th < 0 || tw < 0 || th % g.stride != 0 || tw % g.stride != 0) continue;
      const int ho = th / g.stride, wo = tw / g.stride;
      if (ho >= g.Ho || wo >= g.Wo) continue;
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dcol + ((((long long)n * g.Ho + ho) * g.Wo + wo) * 9 + tap) * g.C + cv * 8)), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    reinterpret_cast<uint4*>(dx)[i] = pack8(acc);
  }
}
// fp32 NCHW image (Cin <= 3... any small Cin) -> cols [N*Ho*Wo, KP] bf16, order (ci, kh, kw) as nn.Conv2d's weight.view(Cout, Cin*9),
// zero-padded to KP columns (KP % 8 == 0 so that the GEMM operand rows are 16-byte aligned)
__global__ void __launch_bounds__(CV_THREADS) im2col3x3_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int N, int Cin,
                                                                    int H, int W, int Ho, int Wo, int stride, int KP) {
  const long long total = (long long)N * Ho * Wo * KP;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const int k = (int)(i % KP);
    const long long p = i / KP;
    float v = 0.f;
    if (k < Cin * 9) {
      const int ci = k / 9, tap = k % 9;
      const int wo = (int)(p % Wo);
      const int ho = (int)((p / Wo) % Ho);
      const int n = (int)(p / ((long long)Wo * Ho));
      const int hi = ho * stride + tap / 3 - 1, wi = wo * stride + tap % 3 - 1;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(x + (((long long)n * Cin + ci) * H + hi) * W + wi);
    }
    col[i] = __float2bfloat16_rn(v);
  }
}

static inline dim3 dw_grid(int C, long long npix) {
  const int cx = (C / 8 + 15) / 16;
  long long gy = (npix + 15) / 16;
  const long long cap = ((long long)num_sms() * 16 + cx - 1) / cx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(cx, (unsigned)gy);
}

}  // namespace vmae

using namespace vmae;

static int dw_geom(DwGeom& g, int N, int H, int W, int C, int stride, const char* who) {
  if (!(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2) && (long long)N * H * W < (1ll << 31)))
    return fail(VMAE_ERR_BAD_ARG, "%s: bad geometry N=%d H=%d W=%d C=%d stride=%d (C %% 8 == 0, stride 1|2)", who, N, H, W, C, stride);
  g.N = N; g.H = H; g.W = W; g.C = C; g.stride = stride;
  g.Ho = (H + 2 - 3) / stride + 1;
  g.Wo = (W + 2 - 3) / stride + 1;
  return VMAE_OK;
}

extern "C" int vmae_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride, void* stream) {
  VMAE_CHECK_ARG(x && w && y, "vmae_dwconv3x3_fwd: null pointer");
  DwGeom g;
  int rc = dw_geom(g, N, H, W, C, stride, "vmae_dwconv3x3_fwd");
  if (rc) return rc;
  rc = dwconv_tma_fwd(x, w, y, N, H, W, C, stride, (cudaStream_t)stream);
  if (rc <= 0) return rc;
  dwconv_fwd_kernel<<<dw_grid(C, (long long)N * g.Ho * g.Wo), dim3(16, 16), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, g);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_dwconv3x3_bwd(const void* x, const void* dy, const float* w, void* dx, float* dw, int N, int H, int W, int C, int stride,
                                  void* stream) {
  VMAE_CHECK_ARG(x && dy && w, "vmae_dwconv3x3_bwd: null pointer");
  DwGeom g;
  int rc = dw_geom(g, N, H, W, C, stride, "vmae_dwconv3x3_bwd");
  if (rc) return rc;
  if (dx) {
    rc = (stride == 1) ? dwconv_tma_dgrad_s1(dy, w, dx, N, H, W, C, (cudaStream_t)stream)
                       : dwconv_tma_dgrad_s2(dy, w, dx, N, H, W, C, (cudaStream_t)stream);
    if (rc < 0) return rc;
    if (rc == 1) {
      dwconv_dgrad_kernel<<<dw_grid(C, (long long)N * H * W), dim3(16, 16), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, g);
      VMAE_LAUNCH_CHECK();
    }
  }
  if (dw) {
    rc = dwconv_tma_wgrad(x, dy, dw, N, H, W, C, stride, (cudaStream_t)stream);
    if (rc < 0) return rc;
    if (rc == 1) {
      dim3 grid = dw_grid(C, (long long)N * g.Ho * g.Wo);
      if (grid.y > (unsigned)(num_sms() * 4)) grid.y = num_sms() * 4;
      dwconv_wgrad_kernel<<<grid, dim3(16, 16), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, g);
      VMAE_LAUNCH_CHECK();
    }
  }
  return VMAE_OK;
}

extern "C" int vmae_se_pool(const void* x, const void* dy, float* out, int N, int HW, int C, float scale, void* stream) {
  VMAE_CHECK_ARG(x && out && N > 0 && HW > 0 && C > 0 && C % 8 == 0, "vmae_se_pool: bad args");
  dim3 grid((C + 255) / 256, N);
  if (dy) se_pool_kernel<1><<<grid, CV_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, out, HW, C, scale);
  else se_pool_kernel<0><<<grid, CV_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, nullptr, out, HW, C, scale);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_se_scale(const void* a, const float* gate, const float* add, void* y, int N, int HW, int C, float add_scale, void* stream) {
  VMAE_CHECK_ARG((a ? gate != nullptr : add != nullptr) && y && N > 0 && HW > 0 && C > 0 && C % 8 == 0, "vmae_se_scale: bad args");
  const long long M = (long long)N * HW;
  se_scale_kernel<<<ew_blocks(M * (C / 8)), CV_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, gate, add, (__nv_bfloat16*)y, M, HW, C, add_scale);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_im2col3x3(const void* x, void* col, int N, int H, int W, int C, int stride, void* stream) {
  VMAE_CHECK_ARG(x && col, "vmae_im2col3x3: null pointer");
  DwGeom g;
  int rc = dw_geom(g, N, H, W, C, stride, "vmae_im2col3x3");
  if (rc) return rc;
  im2col3x3_kernel<<<ew_blocks((long long)N * g.Ho * g.Wo * 9 * (C / 8)), CV_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, g);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_col2im3x3(const void* dcol, void* dx, int N, int H, int W, int C, int stride, void* stream) {
  VMAE_CHECK_ARG(dcol && dx, "vmae_col2im3x3: null pointer");
  DwGeom g;
  int rc = dw_geom(g, N, H, W, C, stride, "vmae_col2im3x3");
  if (rc) return rc;
  col2im3x3_kernel<<<ew_blocks((long long)N * H * W * (C / 8)), CV_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, g);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
extern "C" int vmae_im2col3x3_nchw_f32(const float* x, void* col, int N, int Cin, int H, int W, int stride, int KP, void* stream) {
  VMAE_CHECK_ARG(x && col && N > 0 && Cin > 0 && H > 0 && W > 0 && (stride == 1 || stride == 2) && KP % 8 == 0 && KP >= Cin * 9,
                 "vmae_im2col3x3_nchw_f32: bad args (KP %% 8 == 0, KP >= 9*Cin)");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  im2col3x3_nchw_kernel<<<ew_blocks((long long)N * Ho * Wo * KP), CV_THREADS, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)col, N, Cin, H, W, Ho, Wo, stride, KP);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
