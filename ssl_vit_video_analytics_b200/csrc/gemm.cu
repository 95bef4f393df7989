// K1 — bf16 GEMM on tcgen05/TMEM with TMA operand staging and fused epilogues (sm_100a only).
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )            (fp32 accumulation in TMEM)
//
// replaces nn.Linear forward / dgrad / wgrad of the reference's transformer blocks
// (tiny_vit.py:74-106 Mlp/Attention, mae_vit_adapter.py:24,40-53 enc_to_dec/decoder/pred head).
//
// Operand "majorness" (which dimension is contiguous in memory) is a template parameter so that all three
// training GEMMs run from the tensors as they already sit in HBM, with no transposes:
//   forward  y  = x  W^T : A = x  [M,K]  K-major,   B = W  [N,K]        K-major
//   dgrad    dx = dy W   : A = dy [M,N'] K-major,   B = W  [N'(=K),K'(=N)] -> stored [Kc,N] = MN-major
//   wgrad    dW = dy^T x : A = dy stored [Kc,M] = MN-major, B = x stored [Kc,N] = MN-major (split-K, fp32 atomics)
//
// Structure: persistent CTAs (grid = #SMs), 10 warps: warp0 = TMA producer, warp1 = UMMA issuer (+TMEM alloc),
// warps 2..9 = epilogue (two warps per TMEM lane quadrant, each owning half of the tile's columns).
// Pipelines: smem ring full/empty (TMA <-> UMMA), TMEM accumulator double buffer full/empty (UMMA <-> epilogue).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/vmae.h"

namespace vmae {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = (2 + GEMM_EPI_WARPS) * 32;
constexpr int GEMM_SMEM_BUDGET = 193 * 1024;
constexpr int GEMM_STAGING_BYTES = GEMM_EPI_WARPS * 4096;  // one 32x32 fp32 tile per epilogue warp

struct GemmParams {
  int M, N, K;
  int k_splits;
  const float* bias;            // [N] fp32 or null
  const float* residual;        // [M, ld_res] fp32 or null  (added after activation)
  long long ld_res;
  const __nv_bfloat16* aux_in;  // [M, ld_aux_in] bf16 or null: multiply by gelu'(aux_in)
  long long ld_aux_in;
  __nv_bfloat16* aux_out;       // [M, ld_aux_out] bf16 or null: pre-activation copy (bias added)
  long long ld_aux_out;
  void* out;                    // bf16 or fp32 [M, ld_out]
  long long ld_out;
  int out_fp32;
  int atomic_add;               // fp32 out only: red.add instead of store (split-K / gradient accumulation)
  int act;                      // 0 none, 1 GELU(erf)
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (GEMM_SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (GEMM_SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_STAGING_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

// Exact-erf GELU evaluated without MUFU: Phi(x) = 0.5 + x * P(x^2) on |x| <= 4.2 (degree-8 minimax fit, max abs error
// 1.6e-5 in fp32 Horner form, 1.3e-5 beyond the clamp) — two orders of magnitude below bf16 resolution of the outputs.
__device__ __forceinline__ float normal_cdf(float x) {
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) { return x * normal_cdf(x); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float pdf = 0.39894228040143268f * ex2_approx(-0.72134752044448170f * x * x);  // exp(-x^2/2)
  return fmaf(x, pdf, normal_cdf(x));
}
__device__ __forceinline__ void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_f4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  return u;
}

template <int BLOCK_N, int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = staging_base + GEMM_STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int kb_per_split = (num_kb + p.k_splits - 1) / p.k_splits;
  const int total_units = m_tiles * n_tiles * p.k_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto decode = [&](int unit, int& m0, int& n0, int& kb0, int& kb1) {
    const int nt = unit % n_tiles;
    const int rest = unit / n_tiles;
    const int mt = rest % m_tiles;
    const int ks = rest / m_tiles;
    m0 = mt * GEMM_BLOCK_M;
    n0 = nt * BLOCK_N;
    kb0 = ks * kb_per_split;
    kb1 = min(kb0 + kb_per_split, num_kb);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int m0, n0, kb0, kb1;
        decode(unit, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const int k0 = kb * GEMM_BLOCK_K;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < GEMM_BLOCK_M / 64; ++i)
              tma_load_2d(sa + i * (GEMM_BLOCK_K * 128), &tma_a, full_bar(stage), m0 + i * 64, k0);
          } else {
            tma_load_2d(sa, &tma_a, full_bar(stage), k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i)
              tma_load_2d(sb + i * (GEMM_BLOCK_K * 128), &tma_b, full_bar(stage), n0 + i * 64, k0);
          } else {
            tma_load_2d(sb, &tma_b, full_bar(stage), k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BLOCK_M, BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int m0, n0, kb0, kb1;
        decode(unit, m0, n0, kb0, kb1);
        if (kb0 >= kb1) continue;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            // K-major: advance 16 elements (32 B) inside the 128 B swizzle row; SBO = 8 rows * 128 B.
            // MN-major: advance 16 k-rows of 128 B; LBO = next 64-wide MN chunk (one TMA box), SBO = 8 k-rows.
            const uint64_t adesc = A_MN ? umma_smem_desc_sw128(sa + k * (GEMM_UMMA_K * 128), GEMM_BLOCK_K * 128, 1024)
                                        : umma_smem_desc_sw128(sa + k * (GEMM_UMMA_K * 2), 0, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(sb + k * (GEMM_UMMA_K * 128), GEMM_BLOCK_K * 128, 1024)
                                        : umma_smem_desc_sw128(sb + k * (GEMM_UMMA_K * 2), 0, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // smem slot reusable once these UMMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // Per 32-column chunk: (1) TMEM -> registers (thread = row) -> warp-private swizzled fp32 staging tile in smem;
    // (2) re-read with a coalesced mapping (8 lanes x float4 = one 128 B row segment, 4 rows per instruction) and run
    // the whole epilogue (bias, GELU, gelu', residual, dtype conversion, store / red.add) on that mapping, so every
    // global access touches full 64/128-byte row segments.
    const int ew = warp - 2;
    const int quad = warp & 3;              // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;               // which half of the tile's columns
    constexpr int COLS_PER_WARP = BLOCK_N / 2;
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    const uint32_t stage_tile = staging_base + ew * 4096;
    const int c4 = lane & 7;                // float4 index inside the 32-column chunk
    const int rsub = lane >> 3;             // row inside a group of 4
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int m0, n0, kb0, kb1;
      decode(unit, m0, n0, kb0, kb1);
      if (kb0 >= kb1) continue;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = half * COLS_PER_WARP + c * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + col0, r);
        tmem_wait_ld();
        if (c == CHUNKS - 1) {
          // every TMEM read of this accumulator stage has completed -> hand it back to the UMMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_f4(stage_tile + lane * 128 + ((j ^ (lane & 7)) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int n = n0 + col0 + c4 * 4;
        const bool col_ok = n < p.N;    // N % 8 == 0 and n % 4 == 0: a float4 is either fully inside or fully outside
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rl = it * 4 + rsub;
          float4 v = ld_shared_f4(stage_tile + rl * 128 + ((c4 ^ (rl & 7)) << 4));
          const int row = m0 + quad * 32 + rl;
          if (!col_ok || row >= p.M) continue;
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          const long long ro = static_cast<long long>(row);
          if (p.act == 1) {
            if (p.aux_out) *reinterpret_cast<uint2*>(p.aux_out + ro * p.ld_aux_out + n) = pack_bf16x4(v);
            v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
          }
          if (p.aux_in) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(p.aux_in + ro * p.ld_aux_in + n));
            const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            v.x *= gelu_erf_grad(h0.x); v.y *= gelu_erf_grad(h0.y); v.z *= gelu_erf_grad(h1.x); v.w *= gelu_erf_grad(h1.y);
          }
          if (p.residual) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.residual + ro * p.ld_res + n));
            v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
          }
          if (p.out_fp32) {
            float* dst = reinterpret_cast<float*>(p.out) + ro * p.ld_out + n;
            if (p.atomic_add) red_add_f4(dst, v);
            else *reinterpret_cast<float4*>(dst) = v;
          } else {
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ro * p.ld_out + n) = pack_bf16x4(v);
          }
        }
        __syncwarp();  // staging tile is reused by the next chunk
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BLOCK_N, int A_MN, int B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN>;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

static int pick_block_n(int N, int b_mn) {
  (void)b_mn;
  const int cands[4] = {256, 192, 128, 64};
  for (int c : cands)
    if (N % c == 0) return c;
  // no exact fit: minimise padded columns, prefer the larger tile on ties
  int best = 128, best_waste = 1 << 30;
  for (int c : cands) {
    const int waste = ((N + c - 1) / c) * c - N;
    if (waste < best_waste) { best = c; best_waste = waste; }
  }
  return best;
}

}  // namespace vmae

using namespace vmae;

extern "C" int vmae_gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major,
                              int M, int N, int K, const float* bias, const float* residual, long long ld_res,
                              const void* aux_in, long long ld_aux_in, void* aux_out, long long ld_aux_out, void* out,
                              long long ld_out, int out_dtype, int accumulate, int act, int k_splits, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VMAE_CHECK_ARG(A && B && out, "vmae_gemm_bf16: null operand");
  VMAE_CHECK_ARG(M > 0 && N > 0 && K > 0, "vmae_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  VMAE_CHECK_ARG(N % 8 == 0, "vmae_gemm_bf16: N=%d must be a multiple of 8", N);
  VMAE_CHECK_ARG(a_major == VMAE_MAJOR_K || a_major == VMAE_MAJOR_MN, "vmae_gemm_bf16: bad a_major %d", a_major);
  VMAE_CHECK_ARG(b_major == VMAE_MAJOR_K || b_major == VMAE_MAJOR_MN, "vmae_gemm_bf16: bad b_major %d", b_major);
  VMAE_CHECK_ARG(!(a_major == VMAE_MAJOR_MN && b_major == VMAE_MAJOR_K), "vmae_gemm_bf16: (A MN-major, B K-major) is not instantiated");
  VMAE_CHECK_ARG(out_dtype == VMAE_DTYPE_BF16 || out_dtype == VMAE_DTYPE_F32, "vmae_gemm_bf16: bad out_dtype %d", out_dtype);
  VMAE_CHECK_ARG(!accumulate || out_dtype == VMAE_DTYPE_F32, "vmae_gemm_bf16: accumulate needs fp32 output");
  VMAE_CHECK_ARG(act == VMAE_ACT_NONE || act == VMAE_ACT_GELU, "vmae_gemm_bf16: bad act %d", act);
  if (k_splits < 1) k_splits = 1;
  VMAE_CHECK_ARG(k_splits == 1 || accumulate, "vmae_gemm_bf16: split-K needs accumulate=1");
  VMAE_CHECK_ARG(k_splits == 1 || (!bias && !residual && !aux_in && !aux_out && act == VMAE_ACT_NONE),
                 "vmae_gemm_bf16: split-K supports only the plain accumulate epilogue");
  VMAE_CHECK_ARG(ld_out % (out_dtype == VMAE_DTYPE_F32 ? 4 : 8) == 0, "vmae_gemm_bf16: ld_out misaligned");

  const int block_n = pick_block_n(N, b_major);
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  if (k_splits > num_kb) k_splits = num_kb;
  {  // no empty splits
    const int per = (num_kb + k_splits - 1) / k_splits;
    k_splits = (num_kb + per - 1) / per;
  }

  CUtensorMap ta, tb;
  int rc;
  if (a_major == VMAE_MAJOR_K) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, (uint32_t)GEMM_BLOCK_M};
    rc = make_tmap_bf16(&ta, A, 2, dims, str, box, 128);
  } else {
    const uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
    const uint64_t str[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, 64};
    rc = make_tmap_bf16(&ta, A, 2, dims, str, box, 128);
  }
  if (rc) return rc;
  if (b_major == VMAE_MAJOR_K) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {64, (uint32_t)block_n};
    rc = make_tmap_bf16(&tb, B, 2, dims, str, box, 128);
  } else {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    const uint64_t str[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {64, 64};
    rc = make_tmap_bf16(&tb, B, 2, dims, str, box, 128);
  }
  if (rc) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.k_splits = k_splits;
  p.bias = bias; p.residual = residual; p.ld_res = ld_res;
  p.aux_in = reinterpret_cast<const __nv_bfloat16*>(aux_in); p.ld_aux_in = ld_aux_in;
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(aux_out); p.ld_aux_out = ld_aux_out;
  p.out = out; p.ld_out = ld_out; p.out_fp32 = (out_dtype == VMAE_DTYPE_F32); p.atomic_add = accumulate; p.act = act;

  const int m_tiles = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_tiles = (N + block_n - 1) / block_n;
  const long long units = (long long)m_tiles * n_tiles * k_splits;
  const int grid = (int)(units < num_sms() ? units : num_sms());

#define VMAE_GEMM_DISPATCH(BN)                                                                    \
  if (a_major == VMAE_MAJOR_K && b_major == VMAE_MAJOR_K) return launch_gemm<BN, 0, 0>(ta, tb, p, grid, stream); \
  if (a_major == VMAE_MAJOR_K && b_major == VMAE_MAJOR_MN) return launch_gemm<BN, 0, 1>(ta, tb, p, grid, stream); \
  return launch_gemm<BN, 1, 1>(ta, tb, p, grid, stream);
  switch (block_n) {
    case 256: VMAE_GEMM_DISPATCH(256)
    case 192: VMAE_GEMM_DISPATCH(192)
    case 128: VMAE_GEMM_DISPATCH(128)
    default: VMAE_GEMM_DISPATCH(64)
  }
#undef VMAE_GEMM_DISPATCH
}
