// Micro-benchmarks of the per-SM resources the attention kernels lean on (run on the GPU box):
//   tcgen05.ld / tcgen05.st throughput, MUFU.EX2, bf16x2 pack, FMNMX chains — as cycles per warp-instruction for
//   1..8 resident row-warps per SM. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_sm.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../ssl_vit_video_analytics_b200/csrc/ptx.cuh"

using namespace vmae;

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 2^x for a pair, on the FMA pipe: round-to-nearest split + degree-3 minimax polynomial (rel. err 7.5e-5), exponent by integer add
__device__ __forceinline__ unsigned long long exp2_poly2(unsigned long long x2) {
  float lo, hi; up2(x2, lo, hi);
  x2 = pk2(fmaxf(lo, -125.f), fmaxf(hi, -125.f));
  const unsigned long long MAGIC = pk2(12582912.f, 12582912.f), NMAGIC = pk2(-12582912.f, -12582912.f), M1 = pk2(-1.f, -1.f);
  const unsigned long long C0 = pk2(0.9999280572f, 0.9999280572f), C1 = pk2(0.6932609677f, 0.6932609677f),
                           C2 = pk2(0.2426111251f, 0.2426111251f), C3 = pk2(0.0551716648f, 0.0551716648f);
  const unsigned long long t2 = add2(x2, MAGIC);
  const unsigned long long fl2 = add2(t2, NMAGIC);
  const unsigned long long fr2 = fma2(fl2, M1, x2);
  unsigned long long p2 = fma2(C3, fr2, C2);
  p2 = fma2(p2, fr2, C1);
  p2 = fma2(p2, fr2, C0);
  float tl, th, pl, ph; up2(t2, tl, th); up2(p2, pl, ph);
  const float rl = __uint_as_float(__float_as_uint(pl) + (__float_as_uint(tl) << 23));
  const float rh = __uint_as_float(__float_as_uint(ph) + (__float_as_uint(th) << 23));
  return pk2(rl, rh);
}

// mode 0: ld32 x4 + one wait; 1: ld32 + wait each; 2: st32 x4 + wait; 3: 128 ex2 (4 independent chains); 4: 64 packs; 5: ld32x4 + 128 ex2
// 6: 128 fmax chain (1 chain); 7: fmax 4 chains; 8: 128 ex2 in 16 chains; 9: softmax-like 128 x (ffma, ex2, fadd) + 64 packs
template <int MODE>
__global__ void __launch_bounds__(128, 2) bench_kernel(long long* cycles, float* sink, int iters, int active_warps) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  float acc = threadIdx.x * 1e-3f;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  for (int c = 0; c < 8; ++c) tmem_st32(tmem + lane_off + c * 32, r);
  tmem_wait_st();
  __syncthreads();
  long long t0 = clock64();
  if (warp < active_warps) {
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0 || MODE == 5) {
        uint32_t a[32], b[32], c[32], d[32];
        tmem_ld32(tmem + lane_off + 0, a);
        tmem_ld32(tmem + lane_off + 32, b);
        tmem_ld32(tmem + lane_off + 64, c);
        tmem_ld32(tmem + lane_off + 96, d);
        tmem_wait_ld();
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i] ^ b[i] ^ c[i] ^ d[i]);
        } else {
          float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s0 += ex2a(__uint_as_float(a[i]) * 1e-30f - acc);
            s1 += ex2a(__uint_as_float(b[i]) * 1e-30f - acc);
            s2 += ex2a(__uint_as_float(c[i]) * 1e-30f - acc);
            s3 += ex2a(__uint_as_float(d[i]) * 1e-30f - acc);
          }
          acc = (s0 + s1 + s2 + s3) * 1e-3f;
        }
      } else if (MODE == 1) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t a[32];
          tmem_ld32(tmem + lane_off + c * 32, a);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i]);
        }
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] += it;
        tmem_st32(tmem + lane_off + 0, r);
        tmem_st32(tmem + lane_off + 32, r);
        tmem_st32(tmem + lane_off + 64, r);
        tmem_st32(tmem + lane_off + 96, r);
        tmem_wait_st();
      } else if (MODE == 3) {
        float s0 = acc, s1 = acc + 1, s2 = acc + 2, s3 = acc + 3;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s0 = ex2a(s0 * 0.5f); s1 = ex2a(s1 * 0.5f); s2 = ex2a(s2 * 0.5f); s3 = ex2a(s3 * 0.5f);
        }
        acc = s0 + s1 + s2 + s3;
      } else if (MODE == 8) {   // 16 independent ex2 chains x 8 = 128 ex2
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = acc + i;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = ex2a(v[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i];
        acc *= 1e-3f;
      } else if (MODE == 9) {   // softmax-like: 128 x (ffma, ex2, fadd) + 64 packs, independent
        float v[32]; float s0 = 0.f, s1 = 0.f; uint32_t u = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = ex2a(fmaf(__uint_as_float(r[i]), 1e-9f, acc + k));
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            s0 += v[i]; s1 += v[i + 1];
            __nv_bfloat162 t = __floats2bfloat162_rn(v[i], v[i + 1]);
            u ^= *reinterpret_cast<uint32_t*>(&t);
          }
        }
        acc = (s0 + s1) * 1e-3f + __uint_as_float(u & 0x0fffffffu);
      } else if (MODE == 10 || MODE == 11 || MODE == 12) {   // the forward's whole row phase: 4 loads, max, 128 exps (+sum, pack), 4 stores
        uint32_t sc[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tmem + lane_off + c * 32, sc[c]);
        tmem_wait_ld();
        float mx4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mx4[c] = __uint_as_float(sc[c][0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx4[c] = fmaxf(mx4[c], __uint_as_float(sc[c][i]));
        }
        const float msl = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * 1e-9f + acc;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fmaf(__uint_as_float(sc[c][i]), 1e-9f, -msl), p1 = fmaf(__uint_as_float(sc[c][i + 1]), 1e-9f, -msl);
            if (MODE != 11) { p0 = ex2a(p0); p1 = ex2a(p1); } else { p0 = p0 * 0.001f + 0.5f; p1 = p1 * 0.001f + 0.5f; }
            l4[c] += p0 + p1;
            if (MODE == 12) {
              sc[c][i >> 1] = __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632);   // truncating bf16x2 pack on the ALU pipe
            } else {
              __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
              sc[c][i >> 1] = *reinterpret_cast<uint32_t*>(&t);
            }
          }
        }
        acc = ((l4[0] + l4[1]) + (l4[2] + l4[3])) * 1e-4f;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st16(tmem + lane_off + 128 + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&sc[c][0]));
        tmem_wait_st();
      } else if (MODE >= 20) {   // row phase with packed f32x2 arithmetic; (MODE - 20) of every 16 pairs use the polynomial exp2
        constexpr int NPOLY = MODE - 20;
        uint32_t sc[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tmem + lane_off + c * 32, sc[c]);
        tmem_wait_ld();
        float mx4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mx4[c] = __uint_as_float(sc[c][0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx4[c] = fmaxf(mx4[c], __uint_as_float(sc[c][i]));
        }
        const float msl = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * 1e-9f + acc;
        const unsigned long long sl2_2 = pk2(1e-9f, 1e-9f), nmsl2 = pk2(-msl, -msl);
        unsigned long long l2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            unsigned long long x2 = fma2(pk2(__uint_as_float(sc[c][i]), __uint_as_float(sc[c][i + 1])), sl2_2, nmsl2);
            unsigned long long p2;
            if (((i >> 1) % 16) * NPOLY % 16 < NPOLY) {   // spread NPOLY polynomial pairs evenly over each 16
              p2 = exp2_poly2(x2);
            } else {
              float a, b; up2(x2, a, b);
              p2 = pk2(ex2a(a), ex2a(b));
            }
            l2[c] = add2(l2[c], p2);
            float p0, p1; up2(p2, p0, p1);
            __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
            sc[c][i >> 1] = *reinterpret_cast<uint32_t*>(&t);
          }
        }
        float a0, a1, b0, b1; up2(add2(l2[0], l2[1]), a0, a1); up2(add2(l2[2], l2[3]), b0, b1);
        acc = ((a0 + a1) + (b0 + b1)) * 1e-4f;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st16(tmem + lane_off + 128 + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&sc[c][0]));
        tmem_wait_st();
      } else if (MODE == 4) {
        uint32_t u = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          __nv_bfloat162 t = __floats2bfloat162_rn(acc + i, acc - i);
          u ^= *reinterpret_cast<uint32_t*>(&t);
        }
        acc += __uint_as_float(u & 0x3fffffffu);
      } else if (MODE == 6) {
        float m = acc;
#pragma unroll
        for (int i = 0; i < 128; ++i) m = fmaxf(m, __uint_as_float(r[i & 31]) + (i >> 5));
        acc = m;
      } else if (MODE == 7) {
        float m0 = acc, m1 = acc, m2 = acc, m3 = acc;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          m0 = fmaxf(m0, __uint_as_float(r[i])); m1 = fmaxf(m1, __uint_as_float(r[i]) + 1.f);
          m2 = fmaxf(m2, __uint_as_float(r[i]) + 2.f); m3 = fmaxf(m3, __uint_as_float(r[i]) + 3.f);
        }
        acc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0 && warp < active_warps) cycles[blockIdx.x * 4 + warp] = t1 - t0;
  sink[blockIdx.x * 128 + threadIdx.x] = acc + __uint_as_float(r[3]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int active_warps, double unit_per_iter) {
  int sms = 148, iters = 2000;
  int grid = sms * ctas_per_sm;
  long long* d_cyc; float* d_sink;
  cudaMalloc(&d_cyc, grid * 4 * sizeof(long long));
  cudaMalloc(&d_sink, grid * 128 * sizeof(float));
  cudaMemset(d_cyc, 0, grid * 4 * sizeof(long long));
  bench_kernel<MODE><<<grid, 128>>>(d_cyc, d_sink, iters, active_warps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  static long long h[148 * 8 * 4];
  cudaMemcpy(h, d_cyc, grid * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0; long long mx = 0; int n = 0;
  for (int i = 0; i < grid * 4; ++i) if (h[i] > 0) { sum += h[i]; if (h[i] > mx) mx = h[i]; ++n; }
  printf("%-28s ctas/SM=%d warps/CTA=%d : %8.1f cyc/iter avg (max %8.1f)  -> %.2f cyc per unit/warp\n", name, ctas_per_sm, active_warps,
         sum / n / iters, (double)mx / iters, sum / n / iters / unit_per_iter);
  cudaFree(d_cyc); cudaFree(d_sink);
}

int main(int argc, char** argv) {
  if (argc > 1) {   // MUFU / softmax-like scaling with the number of resident warps per SM sub-partition
    for (int c = 1; c <= 2; ++c) {
      run<10>("fwd row phase", c, 4, 1); run<11>("fwd row phase, no MUFU", c, 4, 1); run<12>("fwd row phase, PRMT pack", c, 4, 1);
      run<20>("row phase f32x2 poly 0/16", c, 4, 1); run<23>("row phase f32x2 poly 3/16", c, 4, 1); run<24>("row phase f32x2 poly 4/16", c, 4, 1);
      run<25>("row phase f32x2 poly 5/16", c, 4, 1); run<26>("row phase f32x2 poly 6/16", c, 4, 1); run<28>("row phase f32x2 poly 8/16", c, 4, 1);
    }
    return 0;
  }

  for (int c = 1; c <= 2; ++c)
    for (int w = 1; w <= 4; w *= 2) {
      run<0>("ld32 x4, one wait", c, w, 4);
      run<1>("ld32+wait x4", c, w, 4);
      run<2>("st32 x4", c, w, 4);
      run<3>("ex2 x128 (4 chains)", c, w, 128);
      run<4>("pack bf16x2 x64", c, w, 64);
      run<5>("ld32x4 + 128 ex2", c, w, 1);
      run<8>("ex2 x128 (16 chains)", c, w, 128);
      run<9>("softmax-like x128", c, w, 128);
      run<6>("fmax chain x128", c, w, 128);
      run<7>("fmax 4 chains x128", c, w, 128);
    }
  return 0;
}
