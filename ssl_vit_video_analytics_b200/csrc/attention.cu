// K2 — flash attention forward / backward on tcgen05 + TMEM (sm_100a), non-causal, no mask, optional in-kernel dropout of
// the attention probabilities. Replaces F.scaled_dot_product_attention (tiny_vit.py:103) and the attention core of
// nn.TransformerEncoderLayer (mae_vit_adapter.py:40-48). Input is the packed projection qkv[B, L, 3, H, hd]; q/k/v and the
// per-head slices are addressed through TMA tensor-map coordinates, so no permute/contiguous copies are ever made.
//
// All kernels here share one skeleton: a CTA owns 128 rows (TMEM lanes) and has 256 threads = a control warpgroup (warp 0
// issues the UMMAs, warp 1 the TMA loads; shrunk to 40 registers) and a row warpgroup (one thread per row, grown to 216
// registers with setmaxnreg). "Score-like" tiles are produced by UMMA into TMEM; a row thread pulls its whole row into
// registers with back-to-back tcgen05.ld (one wait) and hands the columns straight back, so the next tile is computed
// underneath the element-wise phase. Two CTAs are co-resident per SM.
//
//   forward : S = Q K^T -> online softmax -> P (bf16, written back to TMEM as the A operand) -> O += P V_j in TMEM (lazy rescale)
//   dQ      : S = Q K^T, dP = dO V^T -> dS = P*(dP - delta) (smem operand tile) -> dQ += dS K          (TMEM accumulate)
//   dK/dV   : S^T = K Q^T, dP^T = V dO^T -> P^T, dS^T (smem operand tiles) -> dV += P^T dO, dK += dS^T Q
// The dQ + dK/dV pair serves head_dim 32, short sequences and callers without an fp32 dQ workspace; head_dim 64 otherwise
// runs the single-pass backward in attention_bwd_fused.cu.
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/vmae.h"

namespace vmae {

#include "att_trace.cuh"

constexpr int ATT_FWD_THREADS = 256;  // all kernels: warpgroup 0 = control (warp 0 UMMA issuer, warp 1 TMA producer), warpgroup 1 = one thread per row
// Two CTAs per SM leave 128 registers per thread at launch; the control warpgroup shrinks to ATT_CTRL_REGS and the row
// warpgroup grows to ATT_ROW_REGS (setmaxnreg), enough to keep a whole 128-column score row in registers.
constexpr int ATT_CTRL_REGS = 40;
constexpr int ATT_ROW_REGS = 216;
static_assert(ATT_CTRL_REGS + ATT_ROW_REGS <= 256, "per-CTA register pool: 256 threads x 128");
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {  // single MUFU.EX2; exp2f() costs ~10 instructions (denormal handling)
#ifdef VMAE_ATT_NOEXP   // developer experiment: how much of the kernel time is the MUFU pipe? (wrong numerics)
  return x * 0.001f + 0.5f;
#endif
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#include "att_drop.cuh"
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// 2^x for a pair of lanes on the FMA pipe (the MUFU pipe delivers only 16 ex2 per clock per SM and is the forward kernel's
// first bound): round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax polynomial for 2^f (relative error 7.5e-5,
// far below the bf16 rounding of P), exponent added as an integer. Arguments are clamped at -125 (result ~ 0).
__device__ __forceinline__ f32x2 exp2_poly2(f32x2 x2) {
  float lo, hi;
  up2(x2, lo, hi);
  x2 = pk2(fmaxf(lo, -125.f), fmaxf(hi, -125.f));
  const f32x2 t2 = add2(x2, pk2(12582912.f));            // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const f32x2 fr2 = fma2(add2(t2, pk2(-12582912.f)), pk2(-1.f), x2);
  f32x2 p2 = fma2(pk2(0.0551716648f), fr2, pk2(0.2426111251f));
  p2 = fma2(p2, fr2, pk2(0.6932609677f));
  p2 = fma2(p2, fr2, pk2(0.9999280572f));
  float tl, th, pl, ph;
  up2(t2, tl, th);
  up2(p2, pl, ph);
  return pk2(__uint_as_float(__float_as_uint(pl) + (__float_as_uint(tl) << 23)), __uint_as_float(__float_as_uint(ph) + (__float_as_uint(th) << 23)));
}
#ifndef VMAE_FWD_POLY_PAIRS
#define VMAE_FWD_POLY_PAIRS 4
#endif
constexpr int FWD_POLY_PAIRS = VMAE_FWD_POLY_PAIRS;   // of every 16 column pairs, this many take the polynomial instead of MUFU (measured optimum 5-6)
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Write 32 consecutive columns [col0, col0+32) of row r (as bf16) into a 128-row K-major SWIZZLE_128B operand tile:
// 64-column chunks of 128 rows x 128 B; inside a row the 16-byte piece index is XORed with (r & 7).
__device__ __forceinline__ void store_tile_row32(uint32_t tile, int r, int col0, const float (&v)[32]) {
  const uint32_t row_base = tile + (col0 >> 6) * (128 * 128) + r * 128;
  const int piece0 = (col0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t addr = row_base + (((piece0 + i) ^ (r & 7)) << 4);
    st_shared_v4(addr, pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                 pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  }
}
// Write 32 consecutive K-elements of this thread's row (TMEM lane) as a bf16 UMMA A-operand held in TMEM: 16 packed
// 32-bit columns starting at column (element index / 2) of the operand region.
__device__ __forceinline__ void store_tmem_operand32(uint32_t taddr, const float (&v)[32]) {
  uint32_t u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) u[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  tmem_st16(taddr, u);
}
__device__ __forceinline__ void check_smem_alignment(const void* p) {
  if (threadIdx.x == 0 && (smem_u32(p) & 1023u) != 0) {
    printf("vmae attention: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
}

// ======================================================================================================================
// forward
// ======================================================================================================================
template <int HD, int BKV>
struct FwdCfg {
  static constexpr int ROWB = HD * 2;  // bytes per operand row == swizzle width (64 or 128)
  static constexpr int Q_BYTES = 128 * ROWB;
  static constexpr int KV_BYTES = BKV * ROWB;
  static constexpr int NS = 3;                            // K/V ring depth
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;           // K[NS]
  static constexpr int OFF_V = OFF_K + NS * KV_BYTES;     // V[NS]
  static constexpr int OFF_BAR = OFF_V + NS * KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
  // TMEM columns: S fp32 [0,BKV) | O fp32 [BKV, BKV+HD) | P bf16 A-operand (BKV/2 packed columns)
  static constexpr int TMEM_NEED = BKV + HD + BKV / 2;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : 256;
  static_assert(TMEM_NEED <= 256, "two CTAs per SM share the 512 TMEM columns");
};

template <int HD, int BKV, bool DROP>
__global__ void __launch_bounds__(ATT_FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int L, int H, float scale, const AttnDrop drop) {
  using Cfg = FwdCfg<HD, BKV>;
  static_assert(BKV == 128, "the row threads hold one 128-column score row in registers");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  check_smem_alignment(smem_raw);
  const uint32_t sbase = smem_u32(smem_raw);
  constexpr int NS = Cfg::NS;
  const uint32_t sQ = sbase + Cfg::OFF_Q;
  auto sK = [&](int s) { return sbase + Cfg::OFF_K + s * Cfg::KV_BYTES; };
  auto sV = [&](int s) { return sbase + Cfg::OFF_V + s * Cfg::KV_BYTES; };
  const uint32_t bar = sbase + Cfg::OFF_BAR;
  const uint32_t q_full = bar, s_full = bar + 8, p_ready = bar + 16, o_full = bar + 24, s_free = bar + 104;
  auto kv_full = [&](int s) { return bar + 32 + 8 * s; };
  auto kv_empty = [&](int s) { return bar + 64 + 8 * s; };
  const uint32_t tmem_ptr_addr = bar + 96;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + Cfg::OFF_BAR + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (L + BKV - 1) / BKV;
  ATT_TRACE_INIT();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(s_free, 4);
    mbar_init(o_full, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tS = tmem, tO = tmem + BKV, tP = tmem + BKV + HD;   // P: bf16 A-operand, BKV/2 packed columns
  pdl_wait();               // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp < 4) {
    setmaxnreg_dec<ATT_CTRL_REGS>();
    if (warp == 1 && elect_one()) {   // TMA producer
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      tma_load_3d(sQ, &map_q, q_full, h * HD, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % NS;
        mbar_wait(kv_empty(s), ((j / NS) & 1) ^ 1u);
        mbar_expect_tx(kv_full(s), 2 * Cfg::KV_BYTES);
        tma_load_3d(sK(s), &map_kv, kv_full(s), (H + h) * HD, j * BKV, b);
        tma_load_3d(sV(s), &map_kv, kv_full(s), (2 * H + h) * HD, j * BKV, b);
      }
    } else if (warp == 0 && elect_one()) {   // UMMA issuer
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
      // descriptor bases are built once; inside the loops only the 14-bit start-address field advances (bytes >> 4).
      // (Building a descriptor costs ~15 dependent ALU ops; with 32-cycle UMMAs that issue cost would dominate.)
      const uint64_t dQ = umma_smem_desc(sQ, 0, 8 * Cfg::ROWB, Cfg::ROWB);
      uint64_t dK[NS], dV[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        dK[s] = umma_smem_desc(sK(s), 0, 8 * Cfg::ROWB, Cfg::ROWB);
        dV[s] = umma_smem_desc(sV(s), 0, 8 * Cfg::ROWB, Cfg::ROWB);
      }
      auto issue_s = [&](int j) {
        const int s = j % NS;
        mbar_wait(kv_full(s), (j / NS) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tS, dQ + 2 * k, dK[s] + 2 * k, idesc_s, k > 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % NS;
        ATT_STAMP(1, j, 0);
        if (j + 1 < nkv) {   // the rows hold S_j in registers: compute S_{j+1} underneath their softmax
          mbar_wait(s_free, j & 1);
          tc_fence_after();
          ATT_STAMP(1, j, 1);
          issue_s(j + 1);
        }
        ATT_STAMP(1, j, 2);
        mbar_wait(p_ready, j & 1);   // P_j (bf16) is in TMEM
        tc_fence_after();
        ATT_STAMP(1, j, 3);
        // O += P_j V_j with P read straight from TMEM (TS form): P never touches shared memory
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          umma_bf16_ts(tO, tP + kk * 8, dV[s] + (kk * 16 * Cfg::ROWB) / 16, idesc_o, (j > 0 || kk > 0));
        umma_commit(o_full);
        umma_commit(kv_empty(s));
        ATT_STAMP(1, j, 4);
      }
    }
  } else {
    setmaxnreg_inc<ATT_ROW_REGS>();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = scale * LOG2E;
    const uint32_t drop_base = (static_cast<uint32_t>((b * H + h) * L + q0 + row)) * DROP_CQ + drop.seed_lo;
    // Online softmax with a lazily updated reference maximum: O accumulates in TMEM across all KV blocks and is only
    // rescaled (TMEM -> registers -> TMEM) when some row of the warp sees its maximum grow by more than 2^8; otherwise the
    // stale reference is kept (P <= 256 stays exact enough in bf16 and l / O carry the same factor, which cancels).
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      ATT_STAMP(0, j, 0);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      ATT_STAMP(0, j, 1);
      // The whole score row goes to registers with four back-to-back TMEM loads and ONE wait (a load + wait pair costs
      // ~160 cycles, four pipelined loads ~280), and S is read once instead of twice. As soon as the loads have landed
      // the S columns are free again: the MMA warp computes S_{j+1} underneath this block's softmax.
      uint32_t sc[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + lane_off + c * 32, sc[c]);
      tmem_wait_ld();
      tc_fence_before();
      if (lane == 0) mbar_arrive(s_free);
      ATT_STAMP(0, j, 2);
      const int valid = min(BKV, L - j * BKV);
      if (valid < BKV) {   // ragged last block: columns past the sequence end never win the max and give p == 0
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) sc[c][i] = 0xff800000u;   // -inf
      }
      // row maximum over the 128 scores with 3-input FMNMX3 (sm_100): 64 instead of 127 instructions on the half-rate ALU pipe,
      // four independent chains
      // eight independent chains: with four, the 16-deep dependent FMNMX3 chains (~6 cycles each) made this phase latency-bound
      float mx8[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v = fmaxf(__uint_as_float(sc[c][16 * hh]), __uint_as_float(sc[c][16 * hh + 1]));
#pragma unroll
          for (int i = 2; i < 16; i += 2) v = fmax3(v, __uint_as_float(sc[c][16 * hh + i]), __uint_as_float(sc[c][16 * hh + i + 1]));
          mx8[2 * c + hh] = v;
        }
      }
      const float mx = fmax3(fmax3(mx8[0], mx8[1], mx8[2]), fmax3(mx8[3], mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]));
      if (j == 0) {
        m = mx;
      } else {
        const bool need = (mx - m) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(o_full, (j - 1) & 1);   // P_{j-1} V_{j-1} has retired: O is stable
          tc_fence_after();
          const float m_new = need ? mx : m;
          const float alpha = ex2_approx((m - m_new) * sl2);  // exactly 1 for rows that keep their reference
          m = m_new;
          l *= alpha;
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tO + lane_off + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(tO + lane_off + c * 32, r);
          }
          tmem_wait_st();
        }
      }
      const float msl = m * sl2;
      ATT_STAMP(0, j, 3);
      ATT_STAMP(0, j, 4);
      f32x2 l4[4] = {0ull, 0ull, 0ull, 0ull};
      const f32x2 sl2_2 = pk2(sl2), nmsl_2 = pk2(-msl);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // packed in place (slot i/2 has already been consumed when pair i is written): no second register array
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const f32x2 x2 = fma2(pk2(__uint_as_float(sc[c][i]), __uint_as_float(sc[c][i + 1])), sl2_2, nmsl_2);
          f32x2 p2;
          if ((((i >> 1) * FWD_POLY_PAIRS) & 15) < FWD_POLY_PAIRS) {   // FWD_POLY_PAIRS of 16 pairs, evenly spread
            p2 = exp2_poly2(x2);
          } else {
            float a, b2;
            up2(x2, a, b2);
            p2 = pk2(ex2_approx(a), ex2_approx(b2));
          }
          l4[c] = add2(l4[c], p2);   // the normaliser sums the un-dropped probabilities
          float p0, p1;
          up2(p2, p0, p1);
          if (DROP) {
            const uint32_t w = drop_mix32(drop_base + static_cast<uint32_t>((j * BKV + c * 32 + i) >> 1) * DROP_CK) ^ drop.seed_hi;
            if ((w & 0xffffu) < drop.thr) p0 = 0.f;
            if ((w >> 16) < drop.thr) p1 = 0.f;
          }
          sc[c][i >> 1] = pack_bf16x2(p0, p1);
        }
      }
      {
        float a, b2;
        up2(add2(add2(l4[0], l4[1]), add2(l4[2], l4[3])), a, b2);
        l += a + b2;
      }
      // P_j stays in registers until P_{j-1} V_{j-1} has retired (its P operand and its K/V stage are then free). Waiting
      // here instead of before the exponentials takes the MMA round trip (observe p_ready, issue, execute behind the
      // other CTA's MMAs, commit) off the critical path: by now it finished ~1000 cycles ago.
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(tP + lane_off + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&sc[c][0]));
      ATT_STAMP(0, j, 5);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);   // one arrival per warp: mbarrier operations serialise
      ATT_STAMP(0, j, 6);
    }
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = (DROP ? drop.rp : 1.0f) / l;
    const int q_row = q0 + row;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tO + lane_off + c * 32, r);
      tmem_wait_ld();
      if (q_row < L) {
        __nv_bfloat16* dst = out + (static_cast<long long>(b) * L + q_row) * (H * HD) + h * HD + c * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
          pk.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
          pk.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
          pk.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + i) = pk;
        }
      }
    }
    if (q_row < L) lse[(static_cast<long long>(b) * H + h) * L + q_row] = m * scale + __logf(l);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, Cfg::TMEM_COLS); }
}

// ======================================================================================================================
// Forward variants that were built, validated against the fp32 reference and measured on (B, H, L, hd) = (64, 6, 1568, 64), then
// removed because none beat the kernel above (392 us; cuDNN SDPA 358-364 us on the same box). Numbers, traces and the pipe
// micro-benchmark behind them: profiles/r2_attn_fwd_experiments.txt, tools/microbench_sfu.cu, DESIGN.md section 5.
//   * one CTA per SM, two 128-row query tiles in explicit ping-pong over a shared K/V ring, one UMMA-issuing warp per tile, the
//     exponential phases alternated by a token (two mbarriers): 456 us (464 us without the token);
//   * this kernel with the block's 128 key columns split over two row warpgroups (four row warps per sub-partition, shared row
//     maximum through shared memory): 419 us;
//   * both combined (640 threads, 104 registers per row thread): 488 us (500 us without the token);
//   * one CTA per SM, eight row warps of 64 columns with two score register sets per thread (block j+1's TMEM load and row maximum
//     issued under block j's exponentials, independent per-half maxima / normalisers / O accumulators merged once at the end):
//     612 us — the exponentials did overlap the next block's maximum (1 200 cycles for both), but ~750 cycles per block of barrier
//     round trips (54 cycles each even when already satisfied, tools/microbench_mbar.cu), TMEM load issue, P store and hand-off
//     stayed serial, with no second CTA on the SM to cover them;
//   * the same pipelining inside the two-CTA layout: 64-column K/V blocks (a thread holds two blocks of scores in 216 registers), S
//     issued two blocks ahead, P double-buffered in TMEM so that the row warps wait for one barrier per block: 445 us (head_dim 32,
//     L = 196: 333 vs 301 us). Twice as many blocks means twice the per-block hand-offs, and S = Q K^T with N = 64 runs at the 48-cycle
//     shared-memory-read floor instead of 32 (tools/microbench_umma.cu). A lesson kept from it: an mbarrier's single parity bit cannot
//     name a phase two completions back, so a consumer that skips waits may only wait for phase n after it KNOWS phase n-1 completed.
// What they showed: a sub-partition retires one MUFU.EX2 warp-instruction per 8 cycles and overlaps it fully with FFMA2 / FADD2 /
// F2FP / FMNMX, but only with >= 2 warps inside the exponential loop; a lone warp runs the loop in 1 250-1 500 cycles per 128
// columns against 768 of MUFU time, so strict alternation trades the MUFU collisions for the lone-warp inefficiency and loses the
// second CTA's cover of prologue and epilogue. XU pipe 49 % busy, tensor pipe 32 %, issue slots 47 % (ncu) for the kernel above.
// ======================================================================================================================

// ======================================================================================================================
// backward: delta = rowsum(dO * O), then one templated kernel in two modes.
// ======================================================================================================================
// HD/8 consecutive lanes share one (row, head): every lane reads one 16-byte piece of O and of dO, so a warp's loads are
// fully coalesced (one thread per row-head would use 16 bytes of every 128-byte line per load instruction).
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                         float* __restrict__ delta, int B, int L, int H, float4* __restrict__ zero_f4) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int LPR = HD / 8;   // lanes per (row, head): 8 or 4
  const long long total = static_cast<long long>(B) * L * H * LPR;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long total_pad = (total + 31) & ~31LL;   // whole warps stay in the loop: the shuffles below need every lane
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total_pad; t += stride) {
    float acc = 0.f;
    if (t < total) {
      // the fp32 dQ accumulator of the single-pass backward ([B, L, H*HD] floats = two float4 per thread here) is cleared in the same
      // pass: its 154 MB of writes ride under this kernel's 154 MB of reads instead of a separate memset (35 + 40 -> ~47 us at batch 64)
      if (zero_f4 != nullptr) {
        zero_f4[2 * t] = make_float4(0.f, 0.f, 0.f, 0.f);
        zero_f4[2 * t + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(o) + t), d = __ldg(reinterpret_cast<const uint4*>(dout) + t);
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fa = __bfloat1622float2(a2[k]), fd = __bfloat1622float2(d2[k]);
        acc += fa.x * fd.x + fa.y * fd.y;
      }
    }
#pragma unroll
    for (int off = LPR / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (t < total && (t % LPR) == 0) {
      const long long idx = t / LPR;   // (b, l, h) flattened, h fastest
      const int hh = static_cast<int>(idx % H);
      const long long bl = idx / H;
      const int ll = static_cast<int>(bl % L);
      const int bb = static_cast<int>(bl / L);
      delta[(static_cast<long long>(bb) * H + hh) * L + ll] = acc;
    }
  }
}

constexpr int BWD_DQ = 0;
constexpr int BWD_DKDV = 1;

template <int HD, int BS, int MODE>
struct BwdCfg {
  static constexpr int ROWB = HD * 2;
  static constexpr int R_BYTES = 128 * ROWB;   // resident tile
  static constexpr int Y_BYTES = BS * ROWB;    // streamed tile
  static constexpr int E_BYTES = 128 * BS * 2; // bf16 operand tile written by the row threads
  static constexpr int NUM_E = (MODE == BWD_DKDV) ? 2 : 1;
  static constexpr int NS = 3;                 // streamed-tile ring depth (2 stages measured 7% slower)
  static constexpr int OFF_R0 = 0;
  static constexpr int OFF_R1 = OFF_R0 + R_BYTES;
  static constexpr int OFF_Y = OFF_R1 + R_BYTES;           // [NS stages][2 tiles]
  static constexpr int OFF_E = OFF_Y + NS * 2 * Y_BYTES;   // [NUM_E]
  static constexpr int OFF_STAT = OFF_E + NUM_E * E_BYTES; // [2][BS] floats: lse*log2(e) and delta of the streamed query block (dK/dV)
  static constexpr int OFF_BAR = OFF_STAT + (MODE == BWD_DKDV ? 2 * BS * 4 : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
  static_assert(SMEM_BYTES <= 115712, "two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2");
  static constexpr int TMEM_NEED = 2 * BS + NUM_E * HD;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
};

// map_res_* have 128-row boxes (resident tiles), map_str_* have BS-row boxes (streamed tiles).
// 256 threads: warpgroup 0 = control (warp 0 UMMA issuer, warp 1 TMA producer), warpgroup 1 = one thread per row. As in the
// forward, the row threads pull both score-like tiles of an iteration into registers with back-to-back TMEM loads and
// hand the TMEM columns straight back, so T(it+1) is computed underneath the element-wise phase of iteration it.
template <int HD, int BS, int MODE, bool DROP>
__global__ void __launch_bounds__(ATT_FWD_THREADS, 2)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_res_qkv, const __grid_constant__ CUtensorMap map_str_qkv,
                const __grid_constant__ CUtensorMap map_res_do, const __grid_constant__ CUtensorMap map_str_do,
                const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int L, int H,
                float scale, const AttnDrop drop) {
  using Cfg = BwdCfg<HD, BS, MODE>;
  static_assert(BS == 64, "the row threads hold two 64-column tiles in registers");
  constexpr int NS = Cfg::NS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  check_smem_alignment(smem_raw);
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sR0 = sbase + Cfg::OFF_R0, sR1 = sbase + Cfg::OFF_R1;
  auto sY = [&](int s, int t) { return sbase + Cfg::OFF_Y + (s * 2 + t) * Cfg::Y_BYTES; };
  auto sE = [&](int t) { return sbase + Cfg::OFF_E + t * Cfg::E_BYTES; };
  float* stat = reinterpret_cast<float*>(smem_raw + Cfg::OFF_STAT);
  const uint32_t bar = sbase + Cfg::OFF_BAR;
  const uint32_t res_full = bar, t_full = bar + 8, t_free = bar + 16, e_ready = bar + 24, e_free = bar + 32, acc_full = bar + 40;
  auto st_full = [&](int s) { return bar + 48 + 8 * s; };
  auto st_empty = [&](int s) { return bar + 72 + 8 * s; };
  const uint32_t tmem_ptr_addr = bar + 96;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + Cfg::OFF_BAR + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int niter = (L + BS - 1) / BS;
#ifdef VMAE_ATT_TRACE
  const bool att_traced = blockIdx.x == g_att_trace_cta[0] && blockIdx.y == g_att_trace_cta[1] && blockIdx.z == g_att_trace_cta[2] &&
                          (g_att_trace_cta[3] < 0 || g_att_trace_cta[3] == MODE);
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_res_qkv);
    tma_prefetch_desc(&map_str_qkv);
    tma_prefetch_desc(&map_res_do);
    tma_prefetch_desc(&map_str_do);
    mbar_init(res_full, 1);
    mbar_init(t_full, 1);
    mbar_init(t_free, 4);
    mbar_init(e_ready, 4);
    mbar_init(e_free, 1);
    mbar_init(acc_full, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(st_full(s), 1); mbar_init(st_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tT0 = tmem, tT1 = tmem + BS, tA0 = tmem + 2 * BS, tA1 = tmem + 2 * BS + HD;
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) {
    setmaxnreg_dec<ATT_CTRL_REGS>();
    if (warp == 1 && elect_one()) {   // TMA producer
      mbar_expect_tx(res_full, 2 * Cfg::R_BYTES);
      if (MODE == BWD_DQ) {
        tma_load_3d(sR0, &map_res_qkv, res_full, h * HD, r0, b);          // Q
        tma_load_3d(sR1, &map_res_do, res_full, h * HD, r0, b);           // dO
      } else {
        tma_load_3d(sR0, &map_res_qkv, res_full, (H + h) * HD, r0, b);    // K
        tma_load_3d(sR1, &map_res_qkv, res_full, (2 * H + h) * HD, r0, b);// V
      }
      for (int it = 0; it < niter; ++it) {
        const int s = it % NS;
        mbar_wait(st_empty(s), ((it / NS) & 1) ^ 1u);
        mbar_expect_tx(st_full(s), 2 * Cfg::Y_BYTES);
        if (MODE == BWD_DQ) {
          tma_load_3d(sY(s, 0), &map_str_qkv, st_full(s), (H + h) * HD, it * BS, b);      // K_j
          tma_load_3d(sY(s, 1), &map_str_qkv, st_full(s), (2 * H + h) * HD, it * BS, b);  // V_j
        } else {
          tma_load_3d(sY(s, 0), &map_str_qkv, st_full(s), h * HD, it * BS, b);            // Q_i
          tma_load_3d(sY(s, 1), &map_str_do, st_full(s), h * HD, it * BS, b);             // dO_i
        }
      }
    } else if (warp == 0 && elect_one()) {   // UMMA issuer
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, BS, 0, 0);
      constexpr uint32_t idesc_a = umma_idesc_bf16(128, HD, 0, 1);
      // descriptor bases are built once; inside the loops only the 14-bit start-address field advances (bytes >> 4)
      const uint64_t dR0 = umma_smem_desc(sR0, 0, 8 * Cfg::ROWB, Cfg::ROWB), dR1 = umma_smem_desc(sR1, 0, 8 * Cfg::ROWB, Cfg::ROWB);
      uint64_t dY0[NS], dY1[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        dY0[s] = umma_smem_desc(sY(s, 0), 0, 8 * Cfg::ROWB, Cfg::ROWB);
        dY1[s] = umma_smem_desc(sY(s, 1), 0, 8 * Cfg::ROWB, Cfg::ROWB);
      }
      const uint64_t dE0 = umma_smem_desc(sE(0), 0, 1024, 128), dE1 = umma_smem_desc(sE(1), 0, 1024, 128);
      auto issue_t = [&](int it) {
        const int s = it % NS;
        mbar_wait(st_full(s), (it / NS) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tT0, dR0 + 2 * k, dY0[s] + 2 * k, idesc_t, k > 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tT1, dR1 + 2 * k, dY1[s] + 2 * k, idesc_t, k > 0);
        umma_commit(t_full);
      };
      mbar_wait(res_full, 0);
      issue_t(0);
      for (int it = 0; it < niter; ++it) {
        const int s = it % NS;
        ATT_STAMP(1, it, 0);
        if (it + 1 < niter) {   // the rows hold T(it) in registers: compute T(it+1) underneath their element-wise phase
          mbar_wait(t_free, it & 1);
          tc_fence_after();
          ATT_STAMP(1, it, 1);
          issue_t(it + 1);
        }
        ATT_STAMP(1, it, 2);
        mbar_wait(e_ready, it & 1);   // E(it) is in shared memory
        tc_fence_after();
        ATT_STAMP(1, it, 3);
#pragma unroll
        for (int kk = 0; kk < BS / 16; ++kk)  // ACC0 += E0 * Y0 (Y0 read as an MN-major B operand: rows = contraction index)
          umma_bf16(tA0, dE0 + 2 * kk, dY0[s] + (kk * 16 * Cfg::ROWB) / 16, idesc_a, (it > 0 || kk > 0));
        if (MODE == BWD_DKDV) {
#pragma unroll
          for (int kk = 0; kk < BS / 16; ++kk)
            umma_bf16(tA1, dE1 + 2 * kk, dY1[s] + (kk * 16 * Cfg::ROWB) / 16, idesc_a, (it > 0 || kk > 0));
        }
        umma_commit(st_empty(s));
        umma_commit(e_free);
        ATT_STAMP(1, it, 4);
      }
      umma_commit(acc_full);
    }
  } else {
    setmaxnreg_inc<ATT_ROW_REGS>();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = scale * LOG2E;
    const int g_row = r0 + row;
    const long long stat_base = (static_cast<long long>(b) * H + h) * L;
    // dropout hash base: DQ threads own a query row (keys vary), dK/dV threads own a key (query rows vary)
    const uint32_t drop_base = (MODE == BWD_DQ) ? static_cast<uint32_t>((b * H + h) * L + g_row) * DROP_CQ + drop.seed_lo
                                                : static_cast<uint32_t>(g_row >> 1) * DROP_CK + drop.seed_lo;
    float lse_row = 0.f, delta_row = 0.f;
    if (MODE == BWD_DQ && g_row < L) {
      lse_row = lse[stat_base + g_row] * LOG2E;
      delta_row = delta[stat_base + g_row];
    }
    // dK/dV: the statistics of the streamed query block (one lse and one delta per column) are the same for every thread.
    // They live in one 512-byte shared buffer: each row thread fetches one value of the NEXT block at the top of an
    // iteration (global-load latency hidden behind the iteration) and deposits it once everybody has finished reading
    // the current block. Columns past the sequence end get lse = +inf, i.e. p == 0 without a predicate.
    const int tid = threadIdx.x - 128;
    const float stat_mul = (tid < BS) ? LOG2E : 1.0f;   // applied when the value is deposited: the load must not be consumed early
    auto fetch_stat = [&](int it) {
      const int g = it * BS + (tid & (BS - 1));
      float v = (tid < BS) ? INFINITY : 0.f;
      if (g < L) v = __ldg(((tid < BS) ? lse : delta) + stat_base + g);
      return v;
    };
    if (MODE == BWD_DKDV) stat[tid] = fetch_stat(0) * stat_mul;
    for (int it = 0; it < niter; ++it) {
      const int valid = min(BS, L - it * BS);
      float stat_next = 0.f;
      if (MODE == BWD_DKDV && it + 1 < niter) stat_next = fetch_stat(it + 1);
      ATT_STAMP(0, it, 0);
      mbar_wait(t_full, it & 1);
      tc_fence_after();
      ATT_STAMP(0, it, 1);
      uint32_t ts[2][32], tp[2][32];   // T0 (scores) and T1 (dP-like) of this row: 2 x 64 fp32
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(tT0 + lane_off + c * 32, ts[c]);
        tmem_ld32(tT1 + lane_off + c * 32, tp[c]);
      }
      tmem_wait_ld();
      tc_fence_before();
      if (lane == 0) mbar_arrive(t_free);
      ATT_STAMP(0, it, 2);
      // p -> ts (in place), ds -> tp (in place)
      if (MODE == BWD_DQ) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float p = ex2_approx(fmaf(__uint_as_float(ts[c][i]), sl2, -lse_row));
            if (valid < BS && c * 32 + i >= valid) p = 0.f;
            float dp = __uint_as_float(tp[c][i]);
            if (DROP) {   // d(prob) = keep ? d(dropped prob) / (1 - p) : 0
              const int col = it * BS + c * 32 + i;
              const uint32_t w = drop_mix32(drop_base + static_cast<uint32_t>(col >> 1) * DROP_CK) ^ drop.seed_hi;
              dp = (((col & 1) ? (w >> 16) : (w & 0xffffu)) < drop.thr) ? 0.f : dp * drop.rp;
            }
            tp[c][i] = __float_as_uint(p * (dp - delta_row));
          }
      } else {
        named_bar_sync(1, 128);   // the statistics deposited at the end of the previous iteration are visible
        const float* st_lse = stat;
        const float* st_delta = stat + BS;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 l4 = *reinterpret_cast<const float4*>(st_lse + c * 32 + 4 * i4);
            const float4 d4 = *reinterpret_cast<const float4*>(st_delta + c * 32 + 4 * i4);
            const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int i = 4 * i4 + t;
              const float p = ex2_approx(fmaf(__uint_as_float(ts[c][i]), sl2, -lv[t]));
              float dp = __uint_as_float(tp[c][i]);
              float pd = p;   // the dropped probability (scaled by 1/(1-p) when dV is written out)
              if (DROP) {
                const uint32_t rid = static_cast<uint32_t>((b * H + h) * L + it * BS + c * 32 + i);   // query row of this column
                const uint32_t w = drop_mix32(rid * DROP_CQ + drop_base) ^ drop.seed_hi;
                const bool dropped = (((g_row & 1) ? (w >> 16) : (w & 0xffffu)) < drop.thr);
                dp = dropped ? 0.f : dp * drop.rp;
                pd = dropped ? 0.f : p;
              }
              ts[c][i] = __float_as_uint(pd);
              tp[c][i] = __float_as_uint(p * (dp - dv[t]));
            }
          }
        named_bar_sync(1, 128);   // everybody has read this block's statistics
        stat[tid] = stat_next * stat_mul;
      }
      ATT_STAMP(0, it, 3);
      if (it > 0) mbar_wait(e_free, (it - 1) & 1);   // the MMAs that read E(it-1) have retired
      ATT_STAMP(0, it, 4);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        store_tile_row32(sE(0), row, c * 32, reinterpret_cast<const float(&)[32]>(tp[c]));
        if (MODE == BWD_DKDV) store_tile_row32(sE(1), row, c * 32, reinterpret_cast<const float(&)[32]>(ts[c]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(e_ready);   // one arrival per warp: mbarrier operations serialise
      ATT_STAMP(0, it, 5);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // write the accumulators: DQ -> q slot; DKDV -> k slot (ACC0) and v slot (ACC1)
    const long long row_off = (static_cast<long long>(b) * L + g_row) * (3 * H * HD);
#pragma unroll
    for (int a = 0; a < Cfg::NUM_E; ++a) {
      const int part = (MODE == BWD_DQ) ? 0 : (a == 0 ? 1 : 2);
      const uint32_t tA = (a == 0) ? tA0 : tA1;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tA + lane_off + c * 32, r);
        tmem_wait_ld();
        if (g_row < L) {
          __nv_bfloat16* dst = dqkv + row_off + (part * H + h) * HD + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 pk;
            // dS was accumulated without the softmax scale; ACC1 of the dK/dV mode is dV = P^T dO (no scale)
            const float f = (MODE == BWD_DKDV && a == 1) ? (DROP ? drop.rp : 1.0f) : scale;
            pk.x = pack_bf16x2(__uint_as_float(r[i]) * f, __uint_as_float(r[i + 1]) * f);
            pk.y = pack_bf16x2(__uint_as_float(r[i + 2]) * f, __uint_as_float(r[i + 3]) * f);
            pk.z = pack_bf16x2(__uint_as_float(r[i + 4]) * f, __uint_as_float(r[i + 5]) * f);
            pk.w = pack_bf16x2(__uint_as_float(r[i + 6]) * f, __uint_as_float(r[i + 7]) * f);
            *reinterpret_cast<uint4*>(dst + i) = pk;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, Cfg::TMEM_COLS); }
}

static int make_map_rows(CUtensorMap* m, const void* base, int B, int L, int width, int hd, int box_rows) {
  const uint64_t dims[3] = {(uint64_t)width, (uint64_t)L, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)width * 2, (uint64_t)L * width * 2};
  const uint32_t box[3] = {(uint32_t)hd, (uint32_t)box_rows, 1};
  return make_tmap_bf16(m, base, 3, dims, str, box, hd * 2);
}

static AttnDrop make_drop(float p_drop, unsigned long long seed) { return make_attn_drop(p_drop, seed); }

template <int HD, int BKV>
static int launch_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, float p_drop, unsigned long long seed,
                      cudaStream_t stream) {
  using Cfg = FwdCfg<HD, BKV>;
  CUtensorMap mq, mkv;
  int rc = make_map_rows(&mq, qkv, B, L, 3 * H * HD, HD, 128);
  if (rc) return rc;
  rc = make_map_rows(&mkv, qkv, B, L, 3 * H * HD, HD, BKV);
  if (rc) return rc;
  const AttnDrop drop = make_drop(p_drop, seed);
  dim3 grid((L + 127) / 128, H, B);
  if (drop.thr > 0) {
    auto kern = attn_fwd_kernel<HD, BKV, true>;
    VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    VMAE_CUDA(launch_pdl(PDL_ATTN, kern, grid, dim3(ATT_FWD_THREADS), Cfg::SMEM_BYTES, stream, mq, mkv, (__nv_bfloat16*)out, lse, L, H, scale, drop));
  } else {
    auto kern = attn_fwd_kernel<HD, BKV, false>;
    VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    VMAE_CUDA(launch_pdl(PDL_ATTN, kern, grid, dim3(ATT_FWD_THREADS), Cfg::SMEM_BYTES, stream, mq, mkv, (__nv_bfloat16*)out, lse, L, H, scale, drop));
  }
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

template <int HD, int BS, int MODE>
static int launch_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int B, int L, int H,
                      float scale, float p_drop, unsigned long long seed, cudaStream_t stream) {
  using Cfg = BwdCfg<HD, BS, MODE>;
  CUtensorMap m_res_qkv, m_str_qkv, m_res_do, m_str_do;
  int rc = make_map_rows(&m_res_qkv, qkv, B, L, 3 * H * HD, HD, 128);
  if (rc) return rc;
  rc = make_map_rows(&m_str_qkv, qkv, B, L, 3 * H * HD, HD, BS);
  if (rc) return rc;
  rc = make_map_rows(&m_res_do, dout, B, L, H * HD, HD, 128);
  if (rc) return rc;
  rc = make_map_rows(&m_str_do, dout, B, L, H * HD, HD, BS);
  if (rc) return rc;
  const AttnDrop drop = make_drop(p_drop, seed);
  dim3 grid((L + 127) / 128, H, B);
  if (drop.thr > 0) {
    auto kern = attn_bwd_kernel<HD, BS, MODE, true>;
    VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    VMAE_CUDA(launch_pdl(PDL_ATTN, kern, grid, dim3(ATT_FWD_THREADS), Cfg::SMEM_BYTES, stream, m_res_qkv, m_str_qkv, m_res_do, m_str_do, lse, delta,
                         (__nv_bfloat16*)dqkv, L, H, scale, drop));
  } else {
    auto kern = attn_bwd_kernel<HD, BS, MODE, false>;
    VMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    VMAE_CUDA(launch_pdl(PDL_ATTN, kern, grid, dim3(ATT_FWD_THREADS), Cfg::SMEM_BYTES, stream, m_res_qkv, m_str_qkv, m_res_do, m_str_do, lse, delta,
                         (__nv_bfloat16*)dqkv, L, H, scale, drop));
  }
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

// keep mask of the attention-probability dropout, materialised for tests: mask[b, h, q, k] = 1 (kept) / 0 (dropped)
__global__ void __launch_bounds__(256) attn_dropout_mask_kernel(uint8_t* __restrict__ mask, long long rows, int L, const AttnDrop drop) {
  const long long total = rows * ((L + 1) / 2);
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += gridDim.x * 256LL) {
    const long long rid = idx / ((L + 1) / 2);
    const int kp = static_cast<int>(idx % ((L + 1) / 2));
    const uint32_t w = drop_word(drop, static_cast<uint32_t>(rid), static_cast<uint32_t>(kp));
    mask[rid * L + 2 * kp] = (w & 0xffffu) >= drop.thr;
    if (2 * kp + 1 < L) mask[rid * L + 2 * kp + 1] = (w >> 16) >= drop.thr;
  }
}

}  // namespace vmae

using namespace vmae;

#ifdef VMAE_ATT_TRACE
extern "C" int vmae_att_trace_select(int bx, int by, int bz, int mode) { return att_trace_select_tu(bx, by, bz, mode); }
extern "C" int vmae_att_trace_read(long long* host) { return att_trace_read_tu(host); }
#endif

extern "C" int vmae_attn_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int hd, float scale, float p_drop,
                             unsigned long long seed, void* stream) {
  VMAE_CHECK_ARG(qkv && out && lse, "vmae_attn_fwd: null pointer");
  VMAE_CHECK_ARG(B > 0 && L > 0 && H > 0, "vmae_attn_fwd: bad shape B=%d L=%d H=%d", B, L, H);
  VMAE_CHECK_ARG(B <= 65535 && H <= 65535, "vmae_attn_fwd: B and H must be <= 65535");
  VMAE_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "vmae_attn_fwd: p_drop=%f outside [0, 1)", (double)p_drop);
  VMAE_CHECK_ARG((long long)B * H * L < (1LL << 32), "vmae_attn_fwd: B*H*L must fit 32 bits");
  cudaStream_t s = (cudaStream_t)stream;
  if (hd == 64) return launch_fwd<64, 128>(qkv, out, lse, B, L, H, scale, p_drop, seed, s);
  if (hd == 32) return launch_fwd<32, 128>(qkv, out, lse, B, L, H, scale, p_drop, seed, s);
  return fail(VMAE_ERR_UNSUPPORTED, "vmae_attn_fwd: head_dim %d unsupported (32 or 64)", hd);
}

extern "C" int vmae_attn_dropout_mask(void* mask, int B, int L, int H, float p_drop, unsigned long long seed, void* stream) {
  VMAE_CHECK_ARG(mask && B > 0 && L > 0 && H > 0, "vmae_attn_dropout_mask: bad args");
  VMAE_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "vmae_attn_dropout_mask: p_drop=%f outside [0, 1)", (double)p_drop);
  VMAE_CHECK_ARG((long long)B * H * L < (1LL << 32), "vmae_attn_dropout_mask: B*H*L must fit 32 bits");
  attn_dropout_mask_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>((uint8_t*)mask, (long long)B * H * L, L, make_drop(p_drop, seed));
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

namespace vmae {
int attn_bwd_fused_hd64(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, float* dq_accum, int B, int L, int H,
                        float scale, float p_drop, unsigned long long seed, cudaStream_t stream, bool accum_zeroed);   // attention_bwd_fused.cu
}

extern "C" int vmae_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, float* dq_accum, void* dqkv,
                             int B, int L, int H, int hd, float scale, float p_drop, unsigned long long seed, void* stream) {
  VMAE_CHECK_ARG(qkv && out && dout && lse && delta && dqkv, "vmae_attn_bwd: null pointer");
  VMAE_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "vmae_attn_bwd: p_drop=%f outside [0, 1)", (double)p_drop);
  VMAE_CHECK_ARG((long long)B * H * L < (1LL << 32), "vmae_attn_bwd: B*H*L must fit 32 bits");
  VMAE_CHECK_ARG(B > 0 && L > 0 && H > 0, "vmae_attn_bwd: bad shape B=%d L=%d H=%d", B, L, H);
  VMAE_CHECK_ARG(B <= 65535 && H <= 65535, "vmae_attn_bwd: B and H must be <= 65535");
  cudaStream_t s = (cudaStream_t)stream;
  // delta: hd/8 lanes per (row, head), two 16-byte loads per thread; enough CTAs that one pass (or a few) covers everything
  const long long total = (long long)B * L * H * (hd / 8);
  long long grid_ll = (total + 255) / 256;
  if (grid_ll > (long long)num_sms() * 128) grid_ll = (long long)num_sms() * 128;
  const int grid = (int)grid_ll;
  int rc;
  if (hd == 64) {
    VMAE_CUDA(launch_pdl(PDL_ATTN, attn_delta_kernel<64>, dim3(grid), dim3(256), 0, s, (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, B, L, H,
                         reinterpret_cast<float4*>(dq_accum)));
    // single-pass kernel (5 GEMM units, one exponential per score, dQ reduced through L2) when the caller provides the fp32
    // dQ accumulator; the two-kernel form otherwise
    if (dq_accum != nullptr) return attn_bwd_fused_hd64(qkv, dout, lse, delta, dqkv, dq_accum, B, L, H, scale, p_drop, seed, s, /*accum_zeroed=*/true);
    rc = launch_bwd<64, 64, BWD_DQ>(qkv, dout, lse, delta, dqkv, B, L, H, scale, p_drop, seed, s);
    if (rc) return rc;
    return launch_bwd<64, 64, BWD_DKDV>(qkv, dout, lse, delta, dqkv, B, L, H, scale, p_drop, seed, s);
  }
  if (hd == 32) {
    VMAE_CUDA(launch_pdl(PDL_ATTN, attn_delta_kernel<32>, dim3(grid), dim3(256), 0, s, (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, B, L, H,
                         static_cast<float4*>(nullptr)));
    rc = launch_bwd<32, 64, BWD_DQ>(qkv, dout, lse, delta, dqkv, B, L, H, scale, p_drop, seed, s);
    if (rc) return rc;
    return launch_bwd<32, 64, BWD_DKDV>(qkv, dout, lse, delta, dqkv, B, L, H, scale, p_drop, seed, s);
  }
  return fail(VMAE_ERR_UNSUPPORTED, "vmae_attn_bwd: head_dim %d unsupported (32 or 64)", hd);
}
