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
//   wgrad    dW = dy^T x : A = dy stored [Kc,M] = MN-major, B = x stored [Kc,N] = MN-major (split-K, TMA reduce-add)
//
// Structure: persistent CTAs (grid = #SMs), 10 warps: warp0 = TMA producer, warp1 = UMMA issuer (+TMEM alloc),
// warps 2..9 = epilogue (two warps per TMEM lane quadrant, each owning half of the tile's columns).
// Pipelines: smem ring full/empty (TMA <-> UMMA), TMEM accumulator double buffer full/empty (UMMA <-> epilogue).
//
// Epilogue (per warp, per 32-column chunk): tcgen05.ld (thread = row) -> epilogue math in registers -> 16-byte stores
// into a warp-private, hardware-swizzled staging box -> ONE TMA store (or TMA reduce-add) of the 32x32 box. Tensors the
// epilogue reads (fp32 residual, bf16 pre-activation for gelu') arrive the same way: a TMA load of the next chunk's box
// is in flight while the current chunk is processed. The threads therefore issue no global loads/stores at all; all
// global traffic is full-line bulk copies and edge tiles are clipped by the TMA unit.
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "philox.cuh"
#include "../../include/vmae.h"

namespace vmae {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 8;   // default: two epilogue warps per TMEM lane quadrant
// EW = 12 (three per quadrant, 64 columns each of a 192-wide tile) is instantiated for the epilogues with an input stream or real
// arithmetic. Measured at M = 100352 (us, 8 -> 12 warps): gelu' multiply (dgrad of fc2) 161 -> 156, 131 -> 117 (N = 1152), 52 -> 45 (N = 384):
// used by default; fp32 residual 72 -> 68 (K = 384) but 135 -> 149 (K = 1536: the larger staging leaves 3 instead of 4 operand stages):
// not used; GELU with the saved gelu' 177 -> 184: no gain, because that kernel is not epilogue-bound but HBM-WRITE-bound — it writes
// 616 MB (activation + gelu'), and a write-only stream on this part tops out at 3.9 TB/s (tools/bw_probe.py: memset 3.9, copy 6.5,
// read-only 5.9 TB/s), i.e. 158 us.
constexpr int gemm_threads(int ew) { return (2 + ew) * 32; }
constexpr int GEMM_SMEM_LIMIT = 227 * 1024;

enum : int {
  EPI_BF16 = 0,        // out bf16 = acc (+bias)
  EPI_BF16_GELU = 1,   // out bf16 = gelu(acc + bias), aux_out bf16 = gelu'(acc + bias)   (saved for backward)
  EPI_BF16_DGELU = 2,  // out bf16 = acc * aux_in                       (aux_in = the saved gelu')
  EPI_F32 = 3,         // out f32  = acc (+bias)
  EPI_F32_RESID = 4,   // out f32  = acc (+bias) + residual
  EPI_F32_ACCUM = 5,   // out f32 += acc            (TMA reduce-add; split-K)
  EPI_BF16_GELU_Q8 = 7,   // EPI_BF16_GELU with gelu' stored as uint8 (see VMAE_ACT_GELU_Q8 in vmae.h)
  EPI_BF16_DGELU_Q8 = 8,  // EPI_BF16_DGELU reading that uint8 tensor
  EPI_BF16_NOISE = 6,  // out bf16 = acc (+bias) + sigma * N(0,1)  (privacy/feature_noise.py:4-7 fused into the producing GEMM; Philox
                       //            keyed on (seed, offset + flattened element index / 4): the stream of the stand-alone kernel)
};

template <int EPI, int EW = GEMM_EPI_WARPS>
struct EpiTraits {
  static constexpr bool OUT_F32 = (EPI == EPI_F32 || EPI == EPI_F32_RESID || EPI == EPI_F32_ACCUM);
  static constexpr bool HAS_IN = (EPI == EPI_BF16_DGELU || EPI == EPI_F32_RESID || EPI == EPI_BF16_DGELU_Q8);
  static constexpr bool DUAL = (EPI == EPI_BF16_GELU || EPI == EPI_BF16_GELU_Q8);
  static constexpr bool Q8 = (EPI == EPI_BF16_GELU_Q8 || EPI == EPI_BF16_DGELU_Q8);
  static constexpr int IN_BYTES = (EPI == EPI_BF16_DGELU_Q8) ? 32 * 32 : 32 * 32 * (OUT_F32 ? 4 : 2);   // bytes of the epilogue-input box
  static constexpr int BOX_BYTES = 32 * 32 * (OUT_F32 ? 4 : 2);     // one 32x32 staging box
  static constexpr int BUF_BYTES = BOX_BYTES * (DUAL ? 2 : 1);     // per buffer (dual: activation + pre-activation)
  // double buffered, except split-K accumulation: its epilogue runs once per long K range, and the 16 KB it gives back (with four
  // epilogue warps) are a fifth operand stage — wgrad went from ~870 to ~1000 TF/s on the decoder shapes with the deeper ring
  static constexpr int NBUF = (EPI == EPI_F32_ACCUM) ? 1 : 2;
  static constexpr int WARP_BYTES = NBUF * BUF_BYTES;
  static constexpr int STAGING_BYTES = EW * WARP_BYTES;
};

struct GemmParams {
  int M, N, K;
  int k_splits;
  const float* bias;  // [N] fp32 or null
  float* bias_grad;   // EPI_F32_ACCUM only: [M] fp32, += sum over the contraction dimension of A (ones-column UMMA)
  // EPI_BF16_NOISE only
  float sigma;
  unsigned long long seed, offset;
  float* eps_out;     // optional fp32 [M, N] (dense): the unit normals that were drawn (exact-parity tests)
};

template <int BLOCK_N, int EPI, int EW = GEMM_EPI_WARPS>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = EpiTraits<EPI, EW>::STAGING_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int ONES_BYTES = (EPI == EPI_F32_ACCUM) ? 2048 : 0;   // 16 x 64 bf16 tile of 1.0 (row-sum UMMA operand)
  static constexpr int AVAIL = GEMM_SMEM_LIMIT - STAGING_BYTES - ONES_BYTES - BAR_BYTES - 1024;
#ifndef VMAE_GEMM_MAX_STAGES
#define VMAE_GEMM_MAX_STAGES 8
#endif
  static constexpr int STAGES = (AVAIL / STAGE_BYTES) > VMAE_GEMM_MAX_STAGES ? VMAE_GEMM_MAX_STAGES : (AVAIL / STAGE_BYTES);
  static constexpr int TMEM_NEED = 2 * BLOCK_N + ((EPI == EPI_F32_ACCUM) ? 32 : 0);
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + ONES_BYTES + BAR_BYTES + 1024;  // +1024: alignment slack
  static_assert(TMEM_NEED <= 512, "accumulating GEMMs use BLOCK_N <= 192 (room for the row-sum columns)");
  static_assert(STAGES >= 2, "not enough shared memory for a double-buffered operand ring");
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
__device__ __forceinline__ float ex2_approx_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) { return x * normal_cdf(x); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float pdf = 0.39894228040143268f * ex2_approx_f(-0.72134752044448170f * x * x);  // exp(-x^2/2)
  return fmaf(x, pdf, normal_cdf(x));
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }

// Staging-box addressing (must match the TMA swizzle of the box's tensor map):
//  fp32 box: 32 rows x 128 B, SWIZZLE_128B: 16-byte slot j of row r lives at r*128 + ((j ^ (r & 7)) << 4)
//  bf16 box: 32 rows x  64 B, SWIZZLE_64B : 16-byte slot j of row r lives at r*64  + ((j ^ ((r >> 1) & 3)) << 4)
__device__ __forceinline__ uint32_t box_f32_slot(uint32_t box, int r, int j) { return box + r * 128 + ((j ^ (r & 7)) << 4); }
__device__ __forceinline__ uint32_t box_bf16_slot(uint32_t box, int r, int j) { return box + r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

// CL = 2: the kernel runs as CTA PAIRS (thread-block clusters of two). The two CTAs of a pair work on vertically adjacent tiles
// (rows m0 and m0 + 128, same columns, same K range), so they need the SAME B tile: each loads one half of it and TMA multicast
// delivers that half into both CTAs' shared memory. Everything else (A tile, UMMAs, TMEM accumulators, epilogue) stays per CTA.
// Why: with 128 x 192 tiles one K block needs 40 KB of operands per 384 UMMA cycles = 107 B/clk/SM, 15.8 KB/clk over 148 SMs,
// about twice what the L2 delivers to the SMs — which is why every tile shape sat at 50-60 % tensor-pipe activity no matter how deep
// the operand ring was. Sharing B takes a pair's L2 reads from 80 to 56 KB per K block (-30 %).
// The only cross-CTA state is the operand ring: a stage is refilled (by BOTH producers, each writing into both CTAs) only when both
// CTAs' UMMAs have drained it, so empty barriers count two arrivals, one of them the peer's multicast commit.
template <int BLOCK_N, int A_MN, int B_MN, int EPI, int EW = GEMM_EPI_WARPS, int CL = 1>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, EPI, EW>;
  using ET = EpiTraits<EPI, EW>;
  static_assert(EW % 4 == 0 && BLOCK_N % (EW / 4 * 32) == 0, "every TMEM lane quadrant gets EW/4 warps with whole 32-column chunks");
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t ones_base = staging_base + Cfg::STAGING_BYTES;
  const uint32_t bar_base = ones_base + Cfg::ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (16 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (18 + a); };
  auto in_bar = [&](int w, int b) { return bar_base + 8u * (20 + 2 * w + b); };  // 2 * EW barriers: epilogue-input boxes
  static_assert(20 + 2 * EW <= 60, "barrier slots");
  const uint32_t tmem_ptr_addr = bar_base + 8u * 60;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  static_assert(CL == 1 || CL == 2, "single CTAs or CTA pairs");
  const int cta_rank = CL == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_unit = static_cast<int>(blockIdx.x) / CL;     // persistent loop: a pair walks the (pair-)units together
  const int unit_stride = static_cast<int>(gridDim.x) / CL;
  const int m_tiles = ((p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M + CL - 1) / CL;   // CL == 2: pairs of row tiles (the last may hang over M)
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int kb_per_split = (num_kb + p.k_splits - 1) / p.k_splits;  // host guarantees: no split is empty
  const int total_units = m_tiles * n_tiles * p.k_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_out);
    if (ET::HAS_IN || ET::DUAL) tma_prefetch_desc(&tma_aux);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EW);
    }
    for (int w = 0; w < EW; ++w) {
      mbar_init(in_bar(w, 0), 1);
      mbar_init(in_bar(w, 1), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (EPI == EPI_F32_ACCUM && warp == 2) {
    // all-ones operand: layout-agnostic (every element is 1.0), so any valid descriptor over these 2 KB works
#pragma unroll
    for (int i = 0; i < 4; ++i) sts_u4(ones_base + (lane * 4 + i) * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();   // the peer's barriers exist before any multicast load or commit can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel; from here on this grid reads
  // and writes global memory, so it waits for that kernel to complete, and lets the next kernel's CTAs start their prologue
  // as soon as CTAs of this grid retire
  pdl_wait();
  pdl_launch_dependents();

  auto decode = [&](int unit, int& m0, int& n0, int& kb0, int& kb1) {
    const int nt = unit % n_tiles;
    const int rest = unit / n_tiles;
    const int mt = rest % m_tiles;
    const int ks = rest / m_tiles;
    m0 = (mt * CL + cta_rank) * GEMM_BLOCK_M;
    n0 = nt * BLOCK_N;
    kb0 = ks * kb_per_split;
    kb1 = min(kb0 + kb_per_split, num_kb);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
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
          if (CL == 2) {
            // this CTA fetches its half of the pair's B tile; multicast writes it into both CTAs and completes the bytes on both
            // CTAs' full barriers (the peer's half arrives the same way: STAGE_BYTES land here in total)
            if (B_MN) {
              constexpr int NB = BLOCK_N / 64, NB0 = (NB + 1) / 2;   // 64-column boxes: rank 0 takes the first NB0
              const int i0 = cta_rank == 0 ? 0 : NB0, i1 = cta_rank == 0 ? NB0 : NB;
              for (int i = i0; i < i1; ++i)
                tma_load_2d_multicast(sb + i * (GEMM_BLOCK_K * 128), &tma_b, full_bar(stage), n0 + i * 64, k0, 0x3);
            } else {
              tma_load_2d_multicast(sb + cta_rank * (BLOCK_N / 2) * 128, &tma_b, full_bar(stage), k0, n0 + cta_rank * (BLOCK_N / 2), 0x3);
            }
          } else if (B_MN) {
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
      // Descriptors are built ONCE; per UMMA only the 14-bit start-address field (bytes >> 4) advances. Building a
      // descriptor from scratch costs ~15 dependent ALU ops per operand, which made this single issuing thread — not the
      // tensor pipe — the limiter (4 UMMAs of 128 cycles each per K block need < 512 cycles of issue work).
      // K-major: advance 16 elements (32 B) inside the 128 B swizzle row; SBO = 8 rows * 128 B.
      // MN-major: advance 16 k-rows of 128 B; LBO = next 64-wide MN chunk (one TMA box), SBO = 8 k-rows.
      const uint64_t dA0 = A_MN ? umma_smem_desc_sw128(smem_base, GEMM_BLOCK_K * 128, 1024) : umma_smem_desc_sw128(smem_base, 0, 1024);
      const uint64_t dB0 = B_MN ? umma_smem_desc_sw128(smem_base + Cfg::A_BYTES, GEMM_BLOCK_K * 128, 1024)
                                : umma_smem_desc_sw128(smem_base + Cfg::A_BYTES, 0, 1024);
      const uint64_t dOnes = umma_smem_desc_sw128(ones_base, 0, 1024);
      constexpr uint32_t A_STEP = (A_MN ? GEMM_UMMA_K * 128 : GEMM_UMMA_K * 2) >> 4;
      constexpr uint32_t B_STEP = (B_MN ? GEMM_UMMA_K * 128 : GEMM_UMMA_K * 2) >> 4;
      constexpr uint32_t STAGE_STEP = Cfg::STAGE_BYTES >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
        int m0, n0, kb0, kb1;
        decode(unit, m0, n0, kb0, kb1);
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        const bool row_sums = (EPI == EPI_F32_ACCUM) && p.bias_grad != nullptr && n0 == 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t dA = dA0 + static_cast<uint64_t>(stage * STAGE_STEP);
          const uint64_t dB = dB0 + static_cast<uint64_t>(stage * STAGE_STEP);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
            umma_bf16(d_tmem, dA + k * A_STEP, dB + k * B_STEP, idesc, acc);
            if (row_sums) {
              // row sums of A over the contraction dimension = bias gradient: D2[128 x 16] += A * ones
              constexpr uint32_t idesc_ones = umma_idesc_bf16(GEMM_BLOCK_M, 16, A_MN, 0);
              umma_bf16(tmem_base + 2 * BLOCK_N + as * 16, dA + k * A_STEP, dOnes + k * ((GEMM_UMMA_K * 2) >> 4), idesc_ones, acc);
            }
          }
          // smem slot reusable once these UMMAs retire (pairs: tell the peer's producer too — it writes half of B into this CTA)
          if (CL == 2) umma_commit_multicast(empty_bar(stage), 0x3);
          else umma_commit(empty_bar(stage));
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
    const int half = ew >> 2;               // which slice of the tile's columns (EW / 4 slices)
    constexpr int COLS_PER_WARP = BLOCK_N / (EW / 4);
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    const uint32_t my_staging = staging_base + ew * ET::WARP_BYTES;
    const int r = lane;                     // row inside the 32-row box
    int as = 0;
    uint32_t aphase = 0;
    uint32_t item = 0;                      // running (unit, chunk) counter: selects the staging buffer and barrier phase

    // coordinates of chunk c of a unit for this warp
    auto chunk_coords = [&](int unit, int c, int& gm, int& gn) {
      int m0, n0, kb0, kb1;
      decode(unit, m0, n0, kb0, kb1);
      gm = m0 + quad * 32;
      gn = n0 + half * COLS_PER_WARP + c * 32;
    };
    if (ET::HAS_IN && lane == 0 && first_unit < total_units) {
      int gm, gn;
      chunk_coords(first_unit, 0, gm, gn);
      mbar_expect_tx(in_bar(ew, 0), ET::IN_BYTES);
      tma_load_2d(my_staging, &tma_aux, in_bar(ew, 0), gn, gm);
    }
    for (int unit = first_unit; unit < total_units; unit += unit_stride) {
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (EPI == EPI_F32_ACCUM && p.bias_grad != nullptr && half == 0) {
        int m0, n0, kb0, kb1;
        decode(unit, m0, n0, kb0, kb1);
        if (n0 == 0) {
          uint32_t rsum[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + 2 * BLOCK_N + as * 16, rsum);
          tmem_wait_ld();
          const int row = m0 + quad * 32 + lane;
          if (row < p.M) atomicAdd(p.bias_grad + row, __uint_as_float(rsum[0]));
        }
      }
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c, ++item) {
        const int b = ET::NBUF == 2 ? (item & 1) : 0;
        const uint32_t buf = my_staging + b * ET::BUF_BYTES;
        int gm, gn;
        chunk_coords(unit, c, gm, gn);
        uint32_t acc[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + half * COLS_PER_WARP + c * 32, acc);
        tmem_wait_ld();
        if (c == CHUNKS - 1) {
          // every TMEM read of this accumulator stage has completed -> hand it back to the UMMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        if (EPI != EPI_F32_ACCUM && EPI != EPI_BF16_DGELU && EPI != EPI_BF16_DGELU_Q8 && p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (gn + j < p.N) {  // N % 8 == 0: a float4 is either fully inside or fully outside
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (EPI == EPI_BF16_NOISE) {
          // z' = z + sigma * eps: this thread owns 32 consecutive columns of row gm + r, i.e. 8 Philox counters of the flattened
          // [M, N] output (N % 8 == 0 and gn % 32 == 0, so the 4-element groups never straddle a row)
          const long long row = gm + r;
          const long long base = row * p.N + gn;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 e = philox_normal4(p.seed, p.offset + static_cast<unsigned long long>((base + 4 * j) >> 2));
            v[4 * j] = __fmaf_rn(p.sigma, e.x, v[4 * j]);
            v[4 * j + 1] = __fmaf_rn(p.sigma, e.y, v[4 * j + 1]);
            v[4 * j + 2] = __fmaf_rn(p.sigma, e.z, v[4 * j + 2]);
            v[4 * j + 3] = __fmaf_rn(p.sigma, e.w, v[4 * j + 3]);
            if (p.eps_out != nullptr && row < p.M && gn + 4 * j < p.N) *reinterpret_cast<float4*>(p.eps_out + base + 4 * j) = e;
          }
        }
        if (ET::HAS_IN) mbar_wait(in_bar(ew, b), (item >> 1) & 1u);
        if (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_BF16_NOISE || EPI == EPI_BF16_GELU_Q8) {
          if (EPI == EPI_BF16_GELU || EPI == EPI_BF16_GELU_Q8) {
            // one evaluation of Phi(x) serves both gelu(x) = x Phi(x) and gelu'(x) = Phi(x) + x phi(x); the derivative is
            // what backward needs, so it is stored instead of the pre-activation (same bytes, and the dgrad epilogue
            // becomes a plain multiply)
            // Packed-pair arithmetic (FFMA2/FMUL2): the epilogue is issue-bound, not latency-bound, so halving the FMA
            // count is what brings it under the 128x256x384 main loop. Same polynomial as normal_cdf(); pdf via one ex2
            // of (log2(1/sqrt(2 pi)) - u log2(e)/2) on the clamped argument (the tail error, <3e-4, is below bf16).
            float d[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const f32x2 x2 = pk2(v[j], v[j + 1]);
              const f32x2 xc2 = pk2(fminf(fmaxf(v[j], -4.2f), 4.2f), fminf(fmaxf(v[j + 1], -4.2f), 4.2f));
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
              up2(fma2(u2, pk2(-0.72134752044448170f), pk2(-1.3257480647361592f)), e0, e1);
              const f32x2 pdf2 = pk2(ex2_approx_f(e0), ex2_approx_f(e1));
              up2(fma2(xc2, pdf2, cdf2), d[j], d[j + 1]);
              up2(mul2(x2, cdf2), v[j], v[j + 1]);
            }
            if (EPI == EPI_BF16_GELU_Q8) {
              // gelu' in [-0.129, 1.129] -> one byte: q = rn((g + 0.13) * 255 / 1.26); the box is 32 rows x 32 bytes, unswizzled
              // Rounding by magic addend (adding 1.5 * 2^23 rounds to the integer grid and leaves it in the low mantissa byte; packed FMA + add per pair) and three
              // byte permutes per word instead of clamp + F2I + shift + or per element: gelu' is bounded by [-0.1290, 1.1290]
              // analytically (and the polynomial's argument is clamped), so q stays inside [0, 255] without a clamp.
              uint32_t q[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float t0, t1, t2, t3;
                up2(add2(fma2(pk2(d[4 * j], d[4 * j + 1]), pk2(202.38095f), pk2(26.309524f)), pk2(12582912.f)), t0, t1);
                up2(add2(fma2(pk2(d[4 * j + 2], d[4 * j + 3]), pk2(202.38095f), pk2(26.309524f)), pk2(12582912.f)), t2, t3);
                const uint32_t lo = __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x0040);
                const uint32_t hi = __byte_perm(__float_as_uint(t2), __float_as_uint(t3), 0x0040);
                q[j] = __byte_perm(lo, hi, 0x5410);
              }
              sts_u4(buf + ET::BOX_BYTES + r * 32, q[0], q[1], q[2], q[3]);
              sts_u4(buf + ET::BOX_BYTES + r * 32 + 16, q[4], q[5], q[6], q[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts_u4(box_bf16_slot(buf + ET::BOX_BYTES, r, j), pack2(d[8 * j], d[8 * j + 1]), pack2(d[8 * j + 2], d[8 * j + 3]),
                       pack2(d[8 * j + 4], d[8 * j + 5]), pack2(d[8 * j + 6], d[8 * j + 7]));
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts_u4(box_bf16_slot(buf, r, j), pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                   pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
        } else if (EPI == EPI_BF16_DGELU_Q8) {
          // the input box holds 32 rows x 32 bytes of quantised gelu' (unswizzled); the bf16 result goes to the same buffer's
          // 64-byte-row swizzled layout, which starts at the same address: read this row's 32 bytes first, then write
          const uint4 qa = lds_u4(buf + r * 32), qb = lds_u4(buf + r * 32 + 16);
          __syncwarp();   // every lane has read its bytes before any lane's (larger) output row overwrites the region
          const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
          // byte -> float without I2F: the byte becomes the low mantissa byte of 2^23 (one permute), minus 2^23 is exact
          float h[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const f32x2 a2 = pk2(__uint_as_float(__byte_perm(qw[j], 0x4B000000u, 0x7540)), __uint_as_float(__byte_perm(qw[j], 0x4B000000u, 0x7541)));
            const f32x2 b2 = pk2(__uint_as_float(__byte_perm(qw[j], 0x4B000000u, 0x7542)), __uint_as_float(__byte_perm(qw[j], 0x4B000000u, 0x7543)));
            up2(fma2(add2(a2, pk2(-8388608.f)), pk2(0.0049411765f), pk2(-0.13f)), h[4 * j], h[4 * j + 1]);
            up2(fma2(add2(b2, pk2(-8388608.f)), pk2(0.0049411765f), pk2(-0.13f)), h[4 * j + 2], h[4 * j + 3]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts_u4(box_bf16_slot(buf, r, j), pack2(v[8 * j] * h[8 * j], v[8 * j + 1] * h[8 * j + 1]), pack2(v[8 * j + 2] * h[8 * j + 2], v[8 * j + 3] * h[8 * j + 3]),
                   pack2(v[8 * j + 4] * h[8 * j + 4], v[8 * j + 5] * h[8 * j + 5]), pack2(v[8 * j + 6] * h[8 * j + 6], v[8 * j + 7] * h[8 * j + 7]));
        } else if (EPI == EPI_BF16_DGELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t a = box_bf16_slot(buf, r, j);
            const uint4 h = lds_u4(a);
            const float2 h0 = unpack2(h.x), h1 = unpack2(h.y), h2 = unpack2(h.z), h3 = unpack2(h.w);
            sts_u4(a, pack2(v[8 * j] * h0.x, v[8 * j + 1] * h0.y), pack2(v[8 * j + 2] * h1.x, v[8 * j + 3] * h1.y),
                   pack2(v[8 * j + 4] * h2.x, v[8 * j + 5] * h2.y), pack2(v[8 * j + 6] * h3.x, v[8 * j + 7] * h3.y));
          }
        } else if (EPI == EPI_F32_RESID) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = box_f32_slot(buf, r, j);
            const float4 q = lds_f4(a);
            sts_f4(a, v[4 * j] + q.x, v[4 * j + 1] + q.y, v[4 * j + 2] + q.z, v[4 * j + 3] + q.w);
          }
        } else {  // EPI_F32, EPI_F32_ACCUM
#pragma unroll
          for (int j = 0; j < 8; ++j) sts_f4(box_f32_slot(buf, r, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (EPI == EPI_F32_ACCUM) tma_reduce_add_2d(&tma_out, buf, gn, gm);
          else tma_store_2d(&tma_out, buf, gn, gm);
          if (EPI == EPI_BF16_GELU || EPI == EPI_BF16_GELU_Q8) tma_store_2d(&tma_aux, buf + ET::BOX_BYTES, gn, gm);
          tma_store_commit();
          tma_store_wait_read<ET::NBUF - 1>();  // the buffer the next item writes has been read out -> free for reuse
          if (ET::HAS_IN) {
            // prefetch the input box of the next item into the other buffer
            int nunit = unit, nc = c + 1;
            if (nc == CHUNKS) { nunit = unit + unit_stride; nc = 0; }
            if (nunit < total_units) {
              int ngm, ngn;
              chunk_coords(nunit, nc, ngm, ngn);
              mbar_expect_tx(in_bar(ew, b ^ 1), ET::IN_BYTES);
              tma_load_2d(my_staging + (b ^ 1) * ET::BUF_BYTES, &tma_aux, in_bar(ew, b ^ 1), ngn, ngm);
            }
          }
        }
        __syncwarp();
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (lane == 0) tma_store_wait<0>();  // all bulk stores of this warp have completed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();   // no CTA retires while its peer may still arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

struct GemmMaps {
  CUtensorMap a, b, out, aux;
  CUtensorMap b_half;   // CTA pairs, K-major B: the same tensor with a box of BLOCK_N / 2 rows
  bool pairs;           // launch as CTA pairs (clusters of two) sharing the B tile by TMA multicast
};

// How many CTA pairs of this kernel can be resident at once (one CTA per SM, both SMs of a pair in one GPC: a GPC with an odd
// number of free SMs leaves one out). Queried once per instantiation.
template <typename K>
static int max_resident_pairs(K kern, int threads, int smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * num_sms());
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = num_sms() / 2 - 4;   // conservative: never more pairs than could be co-resident
  }
  return n;
}

template <int BLOCK_N, int A_MN, int B_MN, int EPI, int EW = GEMM_EPI_WARPS>
static int launch_gemm(const GemmMaps& m, const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, EPI, EW>;
  if (m.pairs) {
    auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, EPI, EW, 2>;
    static thread_local int max_pairs = 0;
    if (max_pairs == 0) {
      VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      max_pairs = max_resident_pairs(kern, gemm_threads(EW), Cfg::SMEM_BYTES);
    }
    const int m_pairs = ((p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M + 1) / 2;
    const long long units = (long long)m_pairs * ((p.N + BLOCK_N - 1) / BLOCK_N) * p.k_splits;
    const int pairs = (int)(units < max_pairs ? units : max_pairs);
    VMAE_CUDA(launch_pdl_cluster(PDL_GEMM, 2, kern, dim3(2 * pairs), dim3(gemm_threads(EW)), Cfg::SMEM_BYTES, stream, m.a, B_MN ? m.b : m.b_half, m.out, m.aux, p));
    return VMAE_OK;
  }
  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, EPI, EW>;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  VMAE_CUDA(launch_pdl(PDL_GEMM, kern, dim3(grid), dim3(gemm_threads(EW)), Cfg::SMEM_BYTES, stream, m.a, m.b, m.out, m.aux, p));
  return VMAE_OK;
}

template <int BLOCK_N, int EPI, int EW = GEMM_EPI_WARPS>
static int dispatch_major(int a_major, int b_major, const GemmMaps& m, const GemmParams& p, int grid, cudaStream_t stream) {
  if (a_major == VMAE_MAJOR_K && b_major == VMAE_MAJOR_K) return launch_gemm<BLOCK_N, 0, 0, EPI, EW>(m, p, grid, stream);
  if (a_major == VMAE_MAJOR_K && b_major == VMAE_MAJOR_MN) return launch_gemm<BLOCK_N, 0, 1, EPI, EW>(m, p, grid, stream);
  return launch_gemm<BLOCK_N, 1, 1, EPI, EW>(m, p, grid, stream);
}

// Epilogues that run with 12 epilogue warps on 192-wide tiles (bit mask over EPI ids; VMAE_GEMM_EW12 overrides for A/B runs)
static int ew12_mask() {
  static const int mask = [] { const char* e = getenv("VMAE_GEMM_EW12"); return e ? atoi(e) : ((1 << EPI_BF16_DGELU) | (1 << EPI_BF16_DGELU_Q8)); }();
  return mask;
}
static bool use_ew12(int epi, int a_major) { return a_major == VMAE_MAJOR_K && ((ew12_mask() >> epi) & 1); }

template <int BLOCK_N>
static int dispatch_epi(int epi, int a_major, int b_major, const GemmMaps& m, const GemmParams& p, int grid, cudaStream_t stream) {
  if constexpr (BLOCK_N == 192) {
    if (use_ew12(epi, a_major)) {
      const bool bmn = b_major == VMAE_MAJOR_MN;
      if (epi == EPI_BF16_GELU && !bmn) return launch_gemm<192, 0, 0, EPI_BF16_GELU, 12>(m, p, grid, stream);
      if (epi == EPI_BF16_DGELU) return bmn ? launch_gemm<192, 0, 1, EPI_BF16_DGELU, 12>(m, p, grid, stream) : launch_gemm<192, 0, 0, EPI_BF16_DGELU, 12>(m, p, grid, stream);
      if (epi == EPI_BF16_DGELU_Q8) return bmn ? launch_gemm<192, 0, 1, EPI_BF16_DGELU_Q8, 12>(m, p, grid, stream) : launch_gemm<192, 0, 0, EPI_BF16_DGELU_Q8, 12>(m, p, grid, stream);
      if (epi == EPI_F32_RESID && !bmn) return launch_gemm<192, 0, 0, EPI_F32_RESID, 12>(m, p, grid, stream);
    }
  }
  switch (epi) {
    case EPI_BF16: return dispatch_major<BLOCK_N, EPI_BF16>(a_major, b_major, m, p, grid, stream);
    case EPI_BF16_GELU: return dispatch_major<BLOCK_N, EPI_BF16_GELU>(a_major, b_major, m, p, grid, stream);
    case EPI_BF16_DGELU: return dispatch_major<BLOCK_N, EPI_BF16_DGELU>(a_major, b_major, m, p, grid, stream);
    case EPI_F32: return dispatch_major<BLOCK_N, EPI_F32>(a_major, b_major, m, p, grid, stream);
    case EPI_F32_RESID: return dispatch_major<BLOCK_N, EPI_F32_RESID>(a_major, b_major, m, p, grid, stream);
    case EPI_BF16_NOISE: return launch_gemm<BLOCK_N, 0, 0, EPI_BF16_NOISE>(m, p, grid, stream);   // forward layout only
    case EPI_BF16_GELU_Q8: return launch_gemm<BLOCK_N, 0, 0, EPI_BF16_GELU_Q8>(m, p, grid, stream);   // forward layout only
    case EPI_BF16_DGELU_Q8:
      if (a_major != VMAE_MAJOR_K) return fail(VMAE_ERR_UNSUPPORTED, "quantised gelu' input needs a K-major A operand");
      return b_major == VMAE_MAJOR_MN ? launch_gemm<BLOCK_N, 0, 1, EPI_BF16_DGELU_Q8>(m, p, grid, stream) : launch_gemm<BLOCK_N, 0, 0, EPI_BF16_DGELU_Q8>(m, p, grid, stream);
    default:
      if constexpr (BLOCK_N <= 192) {
        // split-K accumulation runs its epilogue once per long K range: four epilogue warps with single-buffered staging (16 KB
        // instead of 64) leave room for five operand stages instead of three
        static const bool ew4 = [] { const char* e = getenv("VMAE_GEMM_ACCUM_EW4"); return e ? atoi(e) != 0 : true; }();
        if (ew4) return dispatch_major<BLOCK_N, EPI_F32_ACCUM, 4>(a_major, b_major, m, p, grid, stream);
        return dispatch_major<BLOCK_N, EPI_F32_ACCUM>(a_major, b_major, m, p, grid, stream);
      }
      else return fail(VMAE_ERR_UNSUPPORTED, "accumulating GEMM with BLOCK_N=256 is not instantiated");
  }
}

// Tile width: minimise (rounds over the SMs) x (time of one tile). A tile's time is ~ BLOCK_N plus a fixed part
// (pipeline fill, epilogue tail); wide tiles re-read A less, narrow tiles quantise better when M is small.
static int pick_block_n(int M, int N, int k_splits, bool accumulate, bool ew12) {
  if (ew12 && N >= 192) return 192;   // the 12-warp epilogue exists for 192-wide tiles
  const int cands[3] = {256, 192, 128};
  const long long m_tiles = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int sms = num_sms();
  int best = 128;
  long long best_cost = -1;
  for (int c : cands) {
    if (accumulate && c > 192) continue;  // accumulating GEMMs keep 32 TMEM columns for the row sums
    const long long units = m_tiles * ((N + c - 1) / c) * k_splits;
    const long long rounds = (units + sms - 1) / sms;
    const long long cost = rounds * (c + 48);
    if (best_cost < 0 || cost < best_cost) { best = c; best_cost = cost; }
  }
  return best;
}

static int make_box_map(CUtensorMap* m, const void* base, int dtype, int M, int N, long long ld) {
  const int es = dtype == VMAE_DTYPE_F32 ? 4 : (dtype == VMAE_DTYPE_U8 ? 1 : 2);
  const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
  const uint64_t str[1] = {(uint64_t)ld * es};
  const uint32_t box[2] = {32, 32};
  return make_tmap(m, base, dtype, 2, dims, str, box, dtype == VMAE_DTYPE_U8 ? 0 : 32 * es);   // the 32-byte uint8 rows are not swizzled
}

}  // namespace vmae

using namespace vmae;

static int gemm_impl(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major,
                     int M, int N, int K, const float* bias, const float* residual, long long ld_res,
                     const void* aux_in, long long ld_aux_in, void* aux_out, long long ld_aux_out, void* out,
                     long long ld_out, int out_dtype, int accumulate, int act, int k_splits, float* bias_grad, void* stream_,
                     float noise_sigma = 0.f, unsigned long long noise_seed = 0, unsigned long long noise_offset = 0, float* eps_out = nullptr,
                     bool noise = false) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VMAE_CHECK_ARG(A && B && out, "vmae_gemm_bf16: null operand");
  VMAE_CHECK_ARG(M > 0 && N > 0 && K > 0, "vmae_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  VMAE_CHECK_ARG(N % 8 == 0, "vmae_gemm_bf16: N=%d must be a multiple of 8", N);
  VMAE_CHECK_ARG(a_major == VMAE_MAJOR_K || a_major == VMAE_MAJOR_MN, "vmae_gemm_bf16: bad a_major %d", a_major);
  VMAE_CHECK_ARG(b_major == VMAE_MAJOR_K || b_major == VMAE_MAJOR_MN, "vmae_gemm_bf16: bad b_major %d", b_major);
  VMAE_CHECK_ARG(!(a_major == VMAE_MAJOR_MN && b_major == VMAE_MAJOR_K), "vmae_gemm_bf16: (A MN-major, B K-major) is not instantiated");
  VMAE_CHECK_ARG(out_dtype == VMAE_DTYPE_BF16 || out_dtype == VMAE_DTYPE_F32, "vmae_gemm_bf16: bad out_dtype %d", out_dtype);
  VMAE_CHECK_ARG(!accumulate || out_dtype == VMAE_DTYPE_F32, "vmae_gemm_bf16: accumulate needs fp32 output");
  VMAE_CHECK_ARG(act >= VMAE_ACT_NONE && act <= VMAE_ACT_AUX_Q8, "vmae_gemm_bf16: bad act %d", act);
  if (k_splits < 1) k_splits = 1;
  VMAE_CHECK_ARG(k_splits == 1 || accumulate, "vmae_gemm_bf16: split-K needs accumulate=1");
  VMAE_CHECK_ARG(act != VMAE_ACT_AUX_Q8 || aux_in, "vmae_gemm_bf16: act=AUX_Q8 needs aux_in");
  VMAE_CHECK_ARG(!accumulate || (!bias && !residual && !aux_in && !aux_out && act == VMAE_ACT_NONE),
                 "vmae_gemm_bf16: accumulate supports only the plain epilogue");
  VMAE_CHECK_ARG(ld_out % (out_dtype == VMAE_DTYPE_F32 ? 4 : 8) == 0, "vmae_gemm_bf16: ld_out misaligned");

  // epilogue selection (the fused combinations the training step uses; anything else is refused, never emulated)
  int epi;
  if (noise) {
    VMAE_CHECK_ARG(out_dtype == VMAE_DTYPE_BF16 && !accumulate && act == VMAE_ACT_NONE && !residual && !aux_in && !aux_out &&
                       a_major == VMAE_MAJOR_K && b_major == VMAE_MAJOR_K,
                   "vmae_gemm_bf16_noise: forward layout, bf16 output, bias only");
    epi = EPI_BF16_NOISE;
  } else if (accumulate) {
    epi = EPI_F32_ACCUM;
  } else if (act == VMAE_ACT_GELU || act == VMAE_ACT_GELU_Q8) {
    VMAE_CHECK_ARG(out_dtype == VMAE_DTYPE_BF16 && aux_out && !residual && !aux_in,
                   "vmae_gemm_bf16: act=GELU needs bf16 output with aux_out (pre-activation) and no residual/aux_in");
    VMAE_CHECK_ARG(act == VMAE_ACT_GELU || (a_major == VMAE_MAJOR_K && b_major == VMAE_MAJOR_K && ld_aux_out % 16 == 0),
                   "vmae_gemm_bf16: GELU_Q8 needs the forward layout and ld_aux_out %% 16 == 0");
    epi = act == VMAE_ACT_GELU ? EPI_BF16_GELU : EPI_BF16_GELU_Q8;
  } else if (aux_in) {
    VMAE_CHECK_ARG(out_dtype == VMAE_DTYPE_BF16 && !residual && !aux_out && !bias, "vmae_gemm_bf16: aux_in (gelu') needs a plain bf16 output");
    VMAE_CHECK_ARG(act == VMAE_ACT_NONE || (act == VMAE_ACT_AUX_Q8 && ld_aux_in % 16 == 0), "vmae_gemm_bf16: aux_in takes act NONE (bf16) or AUX_Q8 (uint8, ld %% 16 == 0)");
    epi = act == VMAE_ACT_AUX_Q8 ? EPI_BF16_DGELU_Q8 : EPI_BF16_DGELU;
  } else if (residual) {
    VMAE_CHECK_ARG(out_dtype == VMAE_DTYPE_F32 && !aux_out, "vmae_gemm_bf16: residual needs fp32 output");
    epi = EPI_F32_RESID;
  } else {
    VMAE_CHECK_ARG(!aux_out, "vmae_gemm_bf16: aux_out without act=GELU");
    epi = out_dtype == VMAE_DTYPE_F32 ? EPI_F32 : EPI_BF16;
  }

  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  if (k_splits > num_kb) k_splits = num_kb;
  {  // no empty splits
    const int per = (num_kb + k_splits - 1) / k_splits;
    k_splits = (num_kb + per - 1) / per;
  }

  const bool ew12 = use_ew12(epi, a_major) && !(epi == EPI_BF16_GELU && b_major == VMAE_MAJOR_MN) && !(epi == EPI_F32_RESID && b_major == VMAE_MAJOR_MN);
  const int block_n = pick_block_n(M, N, k_splits, accumulate != 0, ew12);
  GemmMaps maps;
  int rc;
  if (a_major == VMAE_MAJOR_K) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, (uint32_t)GEMM_BLOCK_M};
    rc = make_tmap_bf16(&maps.a, A, 2, dims, str, box, 128);
  } else {
    const uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
    const uint64_t str[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, 64};
    rc = make_tmap_bf16(&maps.a, A, 2, dims, str, box, 128);
  }
  if (rc) return rc;
  if (b_major == VMAE_MAJOR_K) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {64, (uint32_t)block_n};
    rc = make_tmap_bf16(&maps.b, B, 2, dims, str, box, 128);
  } else {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    const uint64_t str[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {64, 64};
    rc = make_tmap_bf16(&maps.b, B, 2, dims, str, box, 128);
  }
  if (rc) return rc;
  // CTA pairs (see the kernel's CL parameter): worth it when there are at least two row tiles per column tile to pair up
  static const int pair_mode = [] { const char* e = getenv("VMAE_GEMM_PAIRS"); return e ? atoi(e) : 0; }();
  maps.pairs = pair_mode != 0 && (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M >= 2;
  maps.b_half = maps.b;
  if (maps.pairs && b_major == VMAE_MAJOR_K) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {64, (uint32_t)block_n / 2};
    rc = make_tmap_bf16(&maps.b_half, B, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  rc = make_box_map(&maps.out, out, out_dtype, M, N, ld_out);
  if (rc) return rc;
  maps.aux = maps.out;
  if (epi == EPI_BF16_GELU) rc = make_box_map(&maps.aux, aux_out, VMAE_DTYPE_BF16, M, N, ld_aux_out);
  else if (epi == EPI_BF16_GELU_Q8) rc = make_box_map(&maps.aux, aux_out, VMAE_DTYPE_U8, M, N, ld_aux_out);
  else if (epi == EPI_BF16_DGELU) rc = make_box_map(&maps.aux, aux_in, VMAE_DTYPE_BF16, M, N, ld_aux_in);
  else if (epi == EPI_BF16_DGELU_Q8) rc = make_box_map(&maps.aux, aux_in, VMAE_DTYPE_U8, M, N, ld_aux_in);
  else if (epi == EPI_F32_RESID) rc = make_box_map(&maps.aux, residual, VMAE_DTYPE_F32, M, N, ld_res);
  if (rc) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.k_splits = k_splits;
  p.bias = bias;
  p.bias_grad = bias_grad;
  p.sigma = noise_sigma;
  p.seed = noise_seed;
  p.offset = noise_offset;
  p.eps_out = eps_out;

  const int m_tiles = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_tiles = (N + block_n - 1) / block_n;
  const long long units = (long long)m_tiles * n_tiles * k_splits;
  const int grid = (int)(units < num_sms() ? units : num_sms());
  switch (block_n) {
    case 256: return dispatch_epi<256>(epi, a_major, b_major, maps, p, grid, stream);
    case 192: return dispatch_epi<192>(epi, a_major, b_major, maps, p, grid, stream);
    default: return dispatch_epi<128>(epi, a_major, b_major, maps, p, grid, stream);
  }
}

extern "C" int vmae_gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major,
                              int M, int N, int K, const float* bias, const float* residual, long long ld_res,
                              const void* aux_in, long long ld_aux_in, void* aux_out, long long ld_aux_out, void* out,
                              long long ld_out, int out_dtype, int accumulate, int act, int k_splits, void* stream) {
  return gemm_impl(A, lda, a_major, B, ldb, b_major, M, N, K, bias, residual, ld_res, aux_in, ld_aux_in, aux_out, ld_aux_out, out, ld_out,
                   out_dtype, accumulate, act, k_splits, nullptr, stream);
}

extern "C" int vmae_linear_wgrad(const void* dy, long long ld_dy, const void* x, long long ld_x, long long tokens, int n_out, int k_in,
                                 float* dw, long long ld_dw, float* dbias, int k_splits, void* stream) {
  VMAE_CHECK_ARG(tokens > 0 && tokens < (1ll << 31), "vmae_linear_wgrad: bad token count %lld", tokens);
  return gemm_impl(dy, ld_dy, VMAE_MAJOR_MN, x, ld_x, VMAE_MAJOR_MN, n_out, k_in, (int)tokens, nullptr, nullptr, 0, nullptr, 0, nullptr, 0,
                   dw, ld_dw, VMAE_DTYPE_F32, 1, VMAE_ACT_NONE, k_splits, dbias, stream);
}

extern "C" int vmae_gemm_bf16_noise(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const float* bias, void* out,
                                    long long ld_out, float sigma, unsigned long long seed, unsigned long long offset, float* eps_out, void* stream) {
  VMAE_CHECK_ARG(sigma >= 0.f, "vmae_gemm_bf16_noise: sigma=%f must be >= 0", (double)sigma);
  return gemm_impl(A, lda, VMAE_MAJOR_K, B, ldb, VMAE_MAJOR_K, M, N, K, bias, nullptr, 0, nullptr, 0, nullptr, 0, out, ld_out, VMAE_DTYPE_BF16, 0,
                   VMAE_ACT_NONE, 1, nullptr, stream, sigma, seed, offset, eps_out, true);
}
