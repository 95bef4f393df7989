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
constexpr int GEMM_SMEM_BUDGET = 200 * 1024;

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
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int BLOCK_N, int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
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
    const int ew = warp - 2;
    const int quad = warp & 3;              // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;               // which half of the tile's columns
    constexpr int COLS_PER_WARP = BLOCK_N / 2;
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int m0, n0, kb0, kb1;
      decode(unit, m0, n0, kb0, kb1);
      if (kb0 >= kb1) continue;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = half * COLS_PER_WARP + c * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + col0, r);
        tmem_wait_ld();
        const int n = n0 + col0;
        if (n >= p.N) continue;  // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n + j < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (!row_ok) continue;
        if (p.act == 1) {
          if (p.aux_out) {
            __nv_bfloat16* dst = p.aux_out + static_cast<long long>(row) * p.ld_aux_out + n;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n + j < p.N) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(dst + j) = pk;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (p.aux_in) {
          const __nv_bfloat16* src = p.aux_in + static_cast<long long>(row) * p.ld_aux_in + n;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (n + j < p.N) {
              const uint4 pk = __ldg(reinterpret_cast<const uint4*>(src + j));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __bfloat1622float2(h2[t]);
                v[j + 2 * t] *= gelu_erf_grad(f.x);
                v[j + 2 * t + 1] *= gelu_erf_grad(f.y);
              }
            }
          }
        }
        if (p.residual) {
          const float* src = p.residual + static_cast<long long>(row) * p.ld_res + n;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n + j < p.N) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(src + j));
              v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
            }
          }
        }
        if (p.out_fp32) {
          float* dst = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ld_out + n;
          if (p.atomic_add) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) atomicAdd(dst + j, v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (n + j < p.N) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ld_out + n;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (n + j < p.N) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(dst + j) = pk;
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (wait::ld above) -> hand it back to the UMMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
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
