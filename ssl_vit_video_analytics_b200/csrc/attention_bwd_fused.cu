// K2 — single-pass flash-attention backward for head_dim 64 (the decoder / ViT-S shape), tcgen05 + TMEM, sm_100a.
//
// The two-kernel backward in attention.cu recomputes S and dP twice (7 GEMM units and 2 exponentials per score). This kernel
// does the textbook 5 units and 1 exponential: a CTA owns 128 key/value rows (dK, dV accumulate in TMEM over the whole
// query sweep) and, per 128-query block, also produces that block's partial dQ = dS K, which is added to an fp32
// accumulator in global memory with TMA reduce-add (L2 atomics, one 32 x 32 box per row warp). Persistent, one CTA per SM
// (it needs 448 TMEM columns); 384 threads = a control warpgroup (TMA producer warp, UMMA issuer warp; shrunk to 64 registers)
// and eight row warps (216 registers): a row thread owns one QUERY row of the block (its lse and delta are two scalars in
// registers) and half of the 128 key columns, so every SM sub-partition has two row warps whose phases overlap.
//
//   per query block i:   T0 = Q_i K^T   (S, 128 q x 128 kv)         T1 = dO_i V^T   (dP)
//                        P = exp2(T0 * scale*log2e - lse_q),         dS = P * (T1 - delta_q)      -> two [q x kv] bf16 tiles in smem
//                        dV += P^T dO_i     dK += dS^T Q_i           (A operand = the P / dS tile read MN-major: no transpose pass)
//                        dQ_i = dS K                                 (A operand = the dS tile read K-major)
// The row warps never meet at a CTA-wide barrier: they synchronise only through the mbarriers they share with the UMMA warp.
// (A first version kept key rows per thread and P^T in TMEM; it needed per-column statistics staged through shared memory —
// 32 LDS.128 per thread and iteration and visible short-scoreboard stalls — and measured 1.3 % slower.)
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/vmae.h"

namespace vmae {

#include "att_trace.cuh"
#include "att_drop.cuh"

namespace fb {
constexpr int HD = 64;
constexpr int BQ = 128;                    // query rows per block (columns of the score tiles)
// G = number of row warpgroups (4 warps each): a row thread owns one key row and 128 / G query columns of every score tile.
template <int G>
struct Cfg {
  static_assert(G == 2 || G == 4, "two or four row warpgroups");
  static constexpr int THREADS = 128 + G * 128;     // warpgroup 0: TMA warp, UMMA warp (+2 idle); then G row warpgroups
  static constexpr int COLS = 128 / G;              // query columns per row thread
  static constexpr int NCH = COLS / 32;             // 32-column register chunks
  static constexpr int DQC = 64 / G;                // dQ-tile columns read out per row thread
  // setmaxnreg budgets: G = 2: 128 x 64 + 256 x 216 <= 384 x 168; G = 4: 128 x 56 + 512 x 104 <= 640 x 96
  static constexpr int CTRL_REGS = (G == 2) ? 64 : 56;
  static constexpr int ROW_REGS = (G == 2) ? 216 : 104;
};
constexpr int NS = 3;                      // Q/dO ring depth
constexpr int TILE = 128 * HD * 2;         // one [128 x 64] bf16 operand tile: 16 KB
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + TILE;
constexpr int OFF_Y = OFF_V + TILE;        // [NS][2] tiles: Q_i, dO_i
constexpr int OFF_E = OFF_Y + NS * 2 * TILE;       // dS^T: [128 key rows x 128 queries] bf16 = two 64-column SW128 chunks (32 KB)
constexpr int OFF_P = OFF_E + 2 * TILE;            // P: same layout as the dS tile (32 KB)
constexpr int OFF_DQ = OFF_P + 2 * TILE;           // dQ staging: two [128 x 32] fp32 SW128 boxes (32 KB)
constexpr int OFF_BAR = OFF_DQ + 2 * TILE;
constexpr int SMEM_BYTES = OFF_BAR + 128;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");
// TMEM columns
constexpr uint32_t T0 = 0, T1 = 128, DK = 256, DV = 320, DQ = 384;
constexpr float LOG2E_F = 1.4426950408889634f;
}  // namespace fb

__device__ __forceinline__ float fb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for two lanes on the FMA pipe (see attention.cu): degree-3 minimax polynomial, relative error 7.5e-5
__device__ __forceinline__ void fb_exp2_poly2(float x0, float x1, float& y0, float& y1) {
  const f32x2 x2 = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const f32x2 t2 = add2(x2, pk2(12582912.f));
  const f32x2 fr2 = fma2(add2(t2, pk2(-12582912.f)), pk2(-1.f), x2);
  f32x2 p2 = fma2(pk2(0.0551716648f), fr2, pk2(0.2426111251f));
  p2 = fma2(p2, fr2, pk2(0.6932609677f));
  p2 = fma2(p2, fr2, pk2(0.9999280572f));
  float tl, th, pl, ph;
  up2(t2, tl, th);
  up2(p2, pl, ph);
  y0 = __uint_as_float(__float_as_uint(pl) + (__float_as_uint(tl) << 23));
  y1 = __uint_as_float(__float_as_uint(ph) + (__float_as_uint(th) << 23));
}
#ifndef VMAE_FB_GROUPS
#define VMAE_FB_GROUPS 2   // row warpgroups of the single-pass backward (2 or 4; 4 measured slower: 1377 vs 1008 us, register-starved at 104/thread)
#endif
#ifndef VMAE_FB2_POLY
#define VMAE_FB2_POLY 0   // v2: of every 8 column pairs, this many take the polynomial exp2 (the v2 element-wise phase is MUFU-bound: 1024 cycles per block)
#endif
#ifndef VMAE_FB_POLY
#define VMAE_FB_POLY 0   // of every 8 column quads, this many take the polynomial exp2 (measured: no gain here, the phase is not MUFU-bound)
#endif
__device__ __forceinline__ uint32_t fb_pack(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void fb_sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fb_sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void fb_tma_reduce_add_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// map_qkv: 128-row boxes over qkv (K/V resident per work item, Q streamed); map_do: 128-row boxes over dO;
// map_dq: fp32 [B, L, H*64], box DQC x 32.
// Persistent: grid = #SMs; a CTA walks work items w = (b, h, key tile) with stride gridDim.x (consecutive items share (b, h),
// so co-running CTAs stream the same Q_i / dO_i blocks out of L2). TMEM is allocated once, the mbarrier phases run on across
// items, and the next item's K/V load overlaps the current item's dK/dV read-out.
// DQRED: the dQ partials leave the SM as vector reductions straight from registers (REDG.ADD.F32x4) instead of being staged in
// shared memory for a TMA reduce-add: the kernel is bound by shared-memory bandwidth (UMMA operand reads + P/dS stores + TMA fills
// ~ 336 KB per 128 x 128 block at 128 B/clk), and the staging round trip is 64 KB of it.
template <int G, bool DROP, bool DQRED>
__global__ void __launch_bounds__(fb::Cfg<G>::THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                      const __grid_constant__ CUtensorMap map_dq, const float* __restrict__ lse, const float* __restrict__ delta,
                      __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dq_accum, int L, int H, int B, float scale, const AttnDrop drop) {
  using namespace fb;
  using C = Cfg<G>;
  constexpr int COLS = C::COLS, NCH = C::NCH, DQC = C::DQC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (threadIdx.x == 0 && (sbase & 1023u) != 0) { printf("vmae attention: dynamic shared memory base is not 1024-byte aligned\n"); __trap(); }
  const uint32_t sK = sbase + OFF_K, sV = sbase + OFF_V, sE = sbase + OFF_E, sP = sbase + OFF_P, sDQ = sbase + OFF_DQ;
  auto sY = [&](int s, int t) { return sbase + OFF_Y + (s * 2 + t) * TILE; };
  const uint32_t bar = sbase + OFF_BAR;
  const uint32_t res_full = bar, t_full = bar + 8, t_free = bar + 16, p_ready = bar + 24, acc_done = bar + 32, kv_free = bar + 104;
  auto st_full = [&](int s) { return bar + 40 + 8 * s; };
  auto st_empty = [&](int s) { return bar + 64 + 8 * s; };
  const uint32_t tmem_ptr_addr = bar + 96;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + OFF_BAR + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int niter = (L + BQ - 1) / BQ;
  const int ntiles = (L + 127) / 128;
  const int total = B * H * ntiles;
  auto decode = [&](int w, int& b, int& h, int& r0) {
    r0 = (w % ntiles) * 128;
    h = (w / ntiles) % H;
    b = w / (ntiles * H);
  };
#ifdef VMAE_ATT_TRACE
  const bool att_cta = static_cast<int>(blockIdx.x) == g_att_trace_cta[0];
  bool att_traced = false;   // set per work item: the stamps of item number g_att_trace_cta[1] of that CTA are kept
#define FB_TRACE_ITEM(n) att_traced = att_cta && static_cast<int>(n) == g_att_trace_cta[1]
#else
#define FB_TRACE_ITEM(n) do { } while (0)
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dq);
    mbar_init(res_full, 1);
    mbar_init(t_full, 1);
    mbar_init(t_free, 4 * G);
    mbar_init(p_ready, 4 * G);
    mbar_init(acc_done, 1);
    mbar_init(kv_free, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(st_full(s), 1); mbar_init(st_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_addr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  pdl_wait();               // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp < 4) {
    setmaxnreg_dec<C::CTRL_REGS>();
  if (warp == 0) {
    if (elect_one()) {   // TMA producer
      uint32_t gi = 0, item_n = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++item_n) {
        int b, h, r0;
        decode(w, b, h, r0);
        if (item_n > 0) mbar_wait(kv_free, (item_n - 1) & 1);   // the previous item's last MMAs have finished reading K / V
        mbar_expect_tx(res_full, 2 * TILE);
        tma_load_3d(sK, &map_qkv, res_full, (H + h) * HD, r0, b);
        tma_load_3d(sV, &map_qkv, res_full, (2 * H + h) * HD, r0, b);
        for (int it = 0; it < niter; ++it, ++gi) {
          const int s = gi % NS;
          mbar_wait(st_empty(s), ((gi / NS) & 1) ^ 1u);
          mbar_expect_tx(st_full(s), 2 * TILE);
          tma_load_3d(sY(s, 0), &map_qkv, st_full(s), h * HD, it * BQ, b);   // Q_i
          tma_load_3d(sY(s, 1), &map_do, st_full(s), h * HD, it * BQ, b);    // dO_i
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // UMMA issuer
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, BQ, 0, 0);    // T = A(K-major) x B(K-major), N = 128
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, HD, 0, 1);    // dQ = dS (K-major A) x K (MN-major B), N = 64
      constexpr uint32_t idesc_a = umma_idesc_bf16(128, HD, 1, 1);    // dV / dK += P^T | dS^T (MN-major A) x dO | Q (MN-major B)
      // every operand tile sits at a fixed offset from the 1024-aligned smem base, so one descriptor (two registers) serves all
      // of them: only the 14-bit start-address field (bytes >> 4) differs
      const uint64_t d0 = umma_smem_desc(sbase, 0, 1024, 128);
      const uint64_t dK_k = d0 + (OFF_K >> 4);                        // K as K-major A (and, advanced by rows, as MN-major B)
      const uint64_t dV_k = d0 + (OFF_V >> 4);
      auto dQ_ = [&](int s) { return d0 + ((OFF_Y + (s * 2) * TILE) >> 4); };
      auto dO_ = [&](int s) { return d0 + ((OFF_Y + (s * 2 + 1) * TILE) >> 4); };
      const uint64_t dE_k = d0 + (OFF_E >> 4);                                                   // dS [q x kv], K-major: 64-key chunks 16 KB apart
      const uint64_t dmn = umma_smem_desc(sbase, 128 * 128, 1024, 128);                          // MN-major: LBO = stride between the two 64-wide M chunks
      const uint64_t dE_mn = dmn + (OFF_E >> 4), dP_mn = dmn + (OFF_P >> 4);                     // dS^T and P^T read out of the [q x kv] tiles
      auto issue_t = [&](uint32_t g) {
        const int s = g % NS;
        mbar_wait(st_full(s), (g / NS) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem + T0, dQ_(s) + 2 * k, dK_k + 2 * k, idesc_t, k > 0);   // S = Q_i K^T
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem + T1, dO_(s) + 2 * k, dV_k + 2 * k, idesc_t, k > 0);   // dP = dO_i V^T
        umma_commit(t_full);
      };
      uint32_t gi = 0, item_n = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++item_n) {
        FB_TRACE_ITEM(item_n);
        mbar_wait(res_full, item_n & 1);
        if (gi > 0) { mbar_wait(t_free, (gi - 1) & 1); }   // the rows have pulled the previous item's last score tiles out of TMEM
        tc_fence_after();
        issue_t(gi);
        for (int it = 0; it < niter; ++it, ++gi) {
          const int s = gi % NS;
          ATT_STAMP(1, it, 0);
          if (it + 1 < niter) {   // the rows hold T(it) in registers: T(it+1) is computed underneath their element-wise phase
            mbar_wait(t_free, gi & 1);
            tc_fence_after();
            ATT_STAMP(1, it, 1);
            issue_t(gi + 1);
          }
          ATT_STAMP(1, it, 2);
          mbar_wait(p_ready, gi & 1);   // P^T(it) is in TMEM, dS^T(it) in shared memory
          tc_fence_after();
          ATT_STAMP(1, it, 3);
#pragma unroll
          for (int kk = 0; kk < BQ / 16; ++kk)   // dV += P^T dO_i  (contraction over the 128 queries; P^T = the P tile read MN-major)
            umma_bf16(tmem + DV, dP_mn + kk * 128, dO_(s) + kk * 128, idesc_a, (it > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < BQ / 16; ++kk)   // dK += dS^T Q_i
            umma_bf16(tmem + DK, dE_mn + kk * 128, dQ_(s) + kk * 128, idesc_a, (it > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < 128 / 16; ++kk)  // dQ_i = dS K      (contraction over the 128 keys of this CTA; dS read K-major)
            umma_bf16(tmem + DQ, dE_k + (kk >> 2) * 1024 + 2 * (kk & 3), dK_k + kk * 128, idesc_q, kk > 0);
          umma_commit(st_empty(s));
          umma_commit(acc_done);
          if (it + 1 == niter) umma_commit(kv_free);
          ATT_STAMP(1, it, 4);
        }
      }
    }
  }
  } else {
    setmaxnreg_inc<C::ROW_REGS>();
    // 4 * G row warps = G groups of four: a thread owns key row `row` (= TMEM lane) and the COLS query columns of its group of
    // every score tile, so each SM sub-partition has G row warps whose phases overlap. Group g also owns DQC columns of each
    // dQ tile and, in the item tail, COLS of the 128 dK | dV accumulator columns.
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = scale * LOG2E_F;
    // Row statistics: a thread owns query row `row` of every block, so lse and delta of that row are two scalars. They are
    // loaded one block ahead with a clamped index, and nothing touches the loaded values until the next iteration starts
    // (a select or multiply right behind the load would stall the warp for the full DRAM latency).
    auto stat_ptr = [&](const float* src, int b, int h, int it) { return src + (static_cast<long long>(b) * H + h) * L + min(it * BQ + row, L - 1); };
    // this warp's 32 rows x DQC columns of a finished dQ tile. stage: TMEM -> registers -> staging box; issue (after a proxy
    // fence + __syncwarp by the caller): one bulk reduction into the global accumulator
    const uint32_t my_box = sDQ + (warp - 4) * (32 * DQC * 4);
    auto stage_dq = [&]() {
      uint32_t r[DQC];
      if (DQC == 32) tmem_ld32(tmem + DQ + lane_off + grp * DQC, reinterpret_cast<uint32_t(&)[32]>(r));
      else tmem_ld16(tmem + DQ + lane_off + grp * DQC, reinterpret_cast<uint32_t(&)[16]>(r));
      if (lane == 0) tma_store_wait_read<0>();   // this warp's previous reduction has read its staging box
      tmem_wait_ld();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < DQC / 4; ++j) {
        // 128-byte rows use the 128B TMA swizzle; 64-byte rows (G = 4) are stored linearly
        const uint32_t addr = (DQC == 32) ? my_box + lane * 128 + ((j ^ (lane & 7)) << 4) : my_box + lane * 64 + (j << 4);
        fb_sts_f4(addr, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
    };
    auto issue_dq = [&](int b, int h, int qb) {
      if (lane == 0) {
        fb_tma_reduce_add_3d(&map_dq, my_box, h * HD + grp * DQC, qb * BQ + quad * 32, b);
        tma_store_commit();
      }
    };
    // DQRED form: load this thread's DQC columns of the finished dQ tile (async), later add them to the accumulator row
    uint32_t rq[DQC];
    auto load_dq = [&]() {
      if (DQC == 32) tmem_ld32(tmem + DQ + lane_off + grp * DQC, reinterpret_cast<uint32_t(&)[32]>(rq));
      else tmem_ld16(tmem + DQ + lane_off + grp * DQC, reinterpret_cast<uint32_t(&)[16]>(rq));
    };
    auto red_dq = [&](int b, int h, int qb) {
      const int q = qb * BQ + row;
      if (q < L) {
        float* dst = dq_accum + (static_cast<long long>(b) * L + q) * (H * HD) + h * HD + grp * DQC;
#pragma unroll
        for (int j = 0; j < DQC / 4; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(__uint_as_float(rq[4 * j])),
                       "f"(__uint_as_float(rq[4 * j + 1])), "f"(__uint_as_float(rq[4 * j + 2])), "f"(__uint_as_float(rq[4 * j + 3]))
                       : "memory");
      }
    };

    uint32_t gi = 0;
    float nxt_lse = 0.f, nxt_delta = 0.f;
    if (static_cast<int>(blockIdx.x) < total) {
      int b, h, r0;
      decode(blockIdx.x, b, h, r0);
      nxt_lse = __ldg(stat_ptr(lse, b, h, 0));
      nxt_delta = __ldg(stat_ptr(delta, b, h, 0));
    }
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      int b, h, r0;
      decode(w, b, h, r0);
      const int g_row = r0 + row;
      FB_TRACE_ITEM((w - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x));
      const bool more_items = w + static_cast<int>(gridDim.x) < total;
      int nb = b, nh = h, nr0 = r0;
      if (more_items) decode(w + gridDim.x, nb, nh, nr0);
      for (int it = 0; it < niter; ++it, ++gi) {
        // this block's row statistics (loaded an iteration ago), then the request for the next block's
        const uint32_t drop_base = static_cast<uint32_t>((b * H + h) * L + it * BQ + row) * DROP_CQ + drop.seed_lo;
        const bool in_range = it * BQ + row < L;
        const float lse_row = in_range ? nxt_lse * LOG2E_F : INFINITY;   // +inf past the end of the sequence: p == 0
        const float delta_row = in_range ? nxt_delta : 0.f;
        const bool last = it + 1 == niter;
        if (!last || more_items) {
          nxt_lse = __ldg(stat_ptr(lse, last ? nb : b, last ? nh : h, last ? 0 : it + 1));
          nxt_delta = __ldg(stat_ptr(delta, last ? nb : b, last ? nh : h, last ? 0 : it + 1));
        }
        ATT_STAMP_T(0, it, 0, 128);
        ATT_STAMP_T(0, it, 1, 128);
        mbar_wait(t_full, gi & 1);
        tc_fence_after();
        ATT_STAMP_T(0, it, 2, 128);
        uint32_t ts[NCH][32], tp[NCH][32];   // this group's columns of T0 (scores) and T1 (dP) of this key row
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tmem_ld32(tmem + T0 + lane_off + grp * COLS + c * 32, ts[c]);
          tmem_ld32(tmem + T1 + lane_off + grp * COLS + c * 32, tp[c]);
        }
        tmem_wait_ld();
        tc_fence_before();
        if (lane == 0) mbar_arrive(t_free);   // the tiles are in registers: T(it+1) is computed underneath the element-wise phase
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fb_ex2(fmaf(__uint_as_float(ts[c][i]), sl2, -lse_row));
            const float p1 = fb_ex2(fmaf(__uint_as_float(ts[c][i + 1]), sl2, -lse_row));
            float dp0 = __uint_as_float(tp[c][i]), dp1 = __uint_as_float(tp[c][i + 1]);
            float pd0 = p0, pd1 = p1;   // the dropped probabilities (scaled by 1/(1-p) when dV is written out)
            if (DROP) {   // same keep mask as the forward kernel: this thread owns query row `row`, the key pair varies
              const int col = r0 + grp * COLS + c * 32 + i;   // global key index (even)
              const uint32_t wd = drop_mix32(drop_base + static_cast<uint32_t>(col >> 1) * DROP_CK) ^ drop.seed_hi;
              const bool d0 = (wd & 0xffffu) < drop.thr, d1 = (wd >> 16) < drop.thr;
              dp0 = d0 ? 0.f : dp0 * drop.rp; pd0 = d0 ? 0.f : pd0;
              dp1 = d1 ? 0.f : dp1 * drop.rp; pd1 = d1 ? 0.f : pd1;
            }
            ts[c][i] = __float_as_uint(pd0);
            ts[c][i + 1] = __float_as_uint(pd1);
            tp[c][i] = __float_as_uint(p0 * (dp0 - delta_row));
            tp[c][i + 1] = __float_as_uint(p1 * (dp1 - delta_row));
          }
        ATT_STAMP_T(0, it, 3, 128);
        if (it > 0) {
          // the MMAs of the previous block have retired: its dQ tile is complete, and P^T / dS^T may be overwritten
          mbar_wait(acc_done, (gi - 1) & 1);
          tc_fence_after();
          ATT_STAMP_T(0, it, 4, 128);
          if (DQRED) load_dq();   // in flight underneath the tile stores below
          else stage_dq();        // its bulk reduction is issued behind the proxy fence that also publishes dS^T below
          ATT_STAMP_T(0, it, 5, 128);
        }
        // P and dS columns of this query row -> the two [q x kv] shared-memory tiles (64-key SW128 chunks)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col = grp * COLS + c * 32;   // first of these 32 key columns inside the tile
          const uint32_t off = (col >> 6) * TILE + row * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t sw = ((((col & 63) >> 3) + i) ^ (row & 7)) * 16;
            fb_sts_u4(sP + off + sw, fb_pack(__uint_as_float(ts[c][8 * i]), __uint_as_float(ts[c][8 * i + 1])),
                      fb_pack(__uint_as_float(ts[c][8 * i + 2]), __uint_as_float(ts[c][8 * i + 3])),
                      fb_pack(__uint_as_float(ts[c][8 * i + 4]), __uint_as_float(ts[c][8 * i + 5])),
                      fb_pack(__uint_as_float(ts[c][8 * i + 6]), __uint_as_float(ts[c][8 * i + 7])));
            fb_sts_u4(sE + off + sw, fb_pack(__uint_as_float(tp[c][8 * i]), __uint_as_float(tp[c][8 * i + 1])),
                      fb_pack(__uint_as_float(tp[c][8 * i + 2]), __uint_as_float(tp[c][8 * i + 3])),
                      fb_pack(__uint_as_float(tp[c][8 * i + 4]), __uint_as_float(tp[c][8 * i + 5])),
                      fb_pack(__uint_as_float(tp[c][8 * i + 6]), __uint_as_float(tp[c][8 * i + 7])));
          }
        }
        ATT_STAMP_T(0, it, 6, 128);
        if (DQRED && it > 0) tmem_wait_ld();   // the dQ tile is in registers before p_ready lets the next dQ MMA overwrite it
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        if (it > 0) { if (DQRED) red_dq(b, h, it - 1); else issue_dq(b, h, it - 1); }
        ATT_STAMP_T(0, it, 7, 128);
      }
      // item tail: last dQ tile, then the dK / dV accumulators of this key tile (the producer is already loading the next K / V)
      ATT_STAMP_T(0, niter, 0, 128);
      mbar_wait(acc_done, (gi - 1) & 1);
      tc_fence_after();
      ATT_STAMP_T(0, niter, 1, 128);
      if (DQRED) {
        load_dq();
        tmem_wait_ld();
        red_dq(b, h, niter - 1);
      } else {
        stage_dq();
        fence_proxy_async_smem();
        __syncwarp();
        issue_dq(b, h, niter - 1);
      }
      ATT_STAMP_T(0, niter, 2, 128);
      const long long row_off = (static_cast<long long>(b) * L + g_row) * (3 * H * HD);
      // this group's COLS of the 128 accumulator columns [dK (64) | dV (64)]: all loads in flight, one wait
      constexpr int ACC_A = 0;
      const int acol = grp * COLS;                     // 0..127 over dK | dV
      const int part = 1 + (acol >> 6);                // 1 = k slot, 2 = v slot of dqkv
      const uint32_t tA = tmem + ((acol >> 6) == ACC_A ? DK : DV) + (acol & 63);
      uint32_t ra[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld32(tA + lane_off + c * 32, ra[c]);
      tmem_wait_ld();
      if (g_row < L) {
        const float f = (part == 1) ? scale : (DROP ? drop.rp : 1.0f);   // dS was accumulated without the softmax scale, P^T without 1/(1-p)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          __nv_bfloat16* dst = dqkv + row_off + (part * H + h) * HD + (acol & 63) + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 pk;
            pk.x = fb_pack(__uint_as_float(ra[c][i]) * f, __uint_as_float(ra[c][i + 1]) * f);
            pk.y = fb_pack(__uint_as_float(ra[c][i + 2]) * f, __uint_as_float(ra[c][i + 3]) * f);
            pk.z = fb_pack(__uint_as_float(ra[c][i + 4]) * f, __uint_as_float(ra[c][i + 5]) * f);
            pk.w = fb_pack(__uint_as_float(ra[c][i + 6]) * f, __uint_as_float(ra[c][i + 7]) * f);
            *reinterpret_cast<uint4*>(dst + i) = pk;
          }
        }
      }
      ATT_STAMP_T(0, niter, 3, 128);
    }
    if (lane == 0) tma_store_wait<0>();   // this warp's bulk reductions have completed before the CTA retires
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ======================================================================================================================
// v2 — the same arithmetic with the per-block chain broken up (round 2).
//
// The clock64 timeline of v1 (tools/trace_attn, B=64 L=1568) shows one block taking ~3500 cycles, of which the tensor pipe works
// ~1660: MMA(it) [dV, dK, dQ: 24 N=64 UMMAs at the 48-cycle shared-memory-read floor, measured in tools/microbench_umma.cu]
// -> rows wait acc_done -> dQ tile TMEM->smem (300) -> P/dS tile stores (770, the st.shared limit of ~85 B/clk) -> proxy fence
// (360) -> p_ready -> MMA(it+1) ...: the row tail and the MMAs alternate instead of overlapping, because the single P/dS
// tiles and the single dQ accumulator may only be rewritten after the previous block's MMAs have retired. Here
//   * the dS tile and the dQ accumulator (TMEM) are double-buffered, the P tile is released by its own barrier right after
//     the dV UMMAs (p_free), so the row warps store P/dS of block it while MMA(it-1) is still running;
//   * the stores are interleaved with the element-wise math chunk by chunk (the LSU drains them underneath the next chunk);
//   * the dQ read-out of block it-1 happens AFTER p_ready(it) has been signalled, i.e. underneath MMA(it);
//   * the streamed Q/dO ring is released after dK (dQ = dS K does not read it), and has 2 stages (the second dS tile took
//     the third stage's 32 KB);
//   * the element-wise phase uses packed f32x2 arithmetic (the phase is issue-bound: 2 row warps per sub-partition).
// ======================================================================================================================
namespace fb2 {
using fb::HD;
using fb::BQ;
using fb::TILE;
using fb::LOG2E_F;
constexpr int NS = 3;
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + TILE;
constexpr int OFF_Y = OFF_V + TILE;                // [NS][2] tiles: Q_i, dO_i
constexpr int OFF_E = OFF_Y + NS * 2 * TILE;       // dS: 2 buffers x [128 q x 128 kv] bf16 (two 64-key SW128 chunks each, 32 KB)
constexpr int OFF_P = OFF_E + 2 * 2 * TILE;        // P: one buffer, same layout
// The dQ staging boxes (eight [32 x 32] fp32 SW128 boxes = 32 KB) live in the dS buffer that is idle: while block it is produced,
// dS[(it-1)&1] has been consumed by MMA(it-1) and is rewritten only by block it+1. Warp w's 4 KB box is exactly the sub-tile
// (64-key chunk w/4, rows 32 (w%4) ..) that the same warp writes as dS, so only its own bulk-reduction read has to be awaited.
constexpr int OFF_BAR = OFF_P + 2 * TILE;
constexpr int SMEM_BYTES = OFF_BAR + 128;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");
constexpr uint32_t T0 = 0, T1 = 128, DK = 256, DV = 320, DQ0 = 384;   // DQ buffer b at DQ0 + 64 b
constexpr int THREADS = 384;
constexpr int CTRL_REGS = 64, ROW_REGS = 216;
constexpr int FB2_POLY = VMAE_FB2_POLY;
}  // namespace fb2

// DQB: the dQ partials are reduced in bf16 straight into the q slot of dqkv (TMA reduce-add on a bf16 tensor map; dS carries the softmax
// scale) instead of an fp32 accumulator that is zeroed before and converted after the kernel (memset 154 MB + convert 231 MB per launch
// at batch 64). 13 bf16 additions per element: relative error ~0.5 %, the size of the bf16 rounding of the other gradients.
template <bool DROP, bool DQB>
__global__ void __launch_bounds__(fb2::THREADS, 1)
attn_bwd_fused2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                       const __grid_constant__ CUtensorMap map_dq, const float* __restrict__ lse, const float* __restrict__ delta,
                       __nv_bfloat16* __restrict__ dqkv, int L, int H, int B, float scale, const AttnDrop drop) {
  using namespace fb2;
  constexpr int COLS = 64, NCH = 2, DQC = 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (threadIdx.x == 0 && (sbase & 1023u) != 0) { printf("vmae attention: dynamic shared memory base is not 1024-byte aligned\n"); __trap(); }
  const uint32_t sK = sbase + OFF_K, sV = sbase + OFF_V, sP = sbase + OFF_P;
  auto sE = [&](uint32_t b) { return sbase + OFF_E + b * (2 * TILE); };
  auto sY = [&](int s, int t) { return sbase + OFF_Y + (s * 2 + t) * TILE; };
  const uint32_t bar = sbase + OFF_BAR;
  const uint32_t res_full = bar, t_full = bar + 8, t_free = bar + 16, p_ready = bar + 24, acc_done = bar + 32, kv_free = bar + 40, p_free = bar + 48;
  auto st_full = [&](int s) { return bar + 56 + 8 * s; };
  auto st_empty = [&](int s) { return bar + 80 + 8 * s; };
  static_assert(NS <= 3, "barrier slots: st_full at +56.., st_empty at +80.., TMEM pointer at +112");
  const uint32_t tmem_ptr_addr = bar + 112;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + OFF_BAR + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int niter = (L + BQ - 1) / BQ;
  const int ntiles = (L + 127) / 128;
  const int total = B * H * ntiles;
  auto decode = [&](int w, int& b, int& h, int& r0) {
    r0 = (w % ntiles) * 128;
    h = (w / ntiles) % H;
    b = w / (ntiles * H);
  };
#ifdef VMAE_ATT_TRACE
  const bool att_cta = static_cast<int>(blockIdx.x) == g_att_trace_cta[0];
  bool att_traced = false;
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dq);
    mbar_init(res_full, 1);
    mbar_init(t_full, 1);
    mbar_init(t_free, 8);
    mbar_init(p_ready, 8);
    mbar_init(acc_done, 1);
    mbar_init(kv_free, 1);
    mbar_init(p_free, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(st_full(s), 1); mbar_init(st_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_addr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) {
    setmaxnreg_dec<CTRL_REGS>();
    if (warp == 0) {
      if (elect_one()) {   // TMA producer
        uint32_t gi = 0, item_n = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x, ++item_n) {
          int b, h, r0;
          decode(w, b, h, r0);
          if (item_n > 0) mbar_wait(kv_free, (item_n - 1) & 1);
          mbar_expect_tx(res_full, 2 * TILE);
          tma_load_3d(sK, &map_qkv, res_full, (H + h) * HD, r0, b);
          tma_load_3d(sV, &map_qkv, res_full, (2 * H + h) * HD, r0, b);
          for (int it = 0; it < niter; ++it, ++gi) {
            const int s = gi % NS;
            mbar_wait(st_empty(s), ((gi / NS) & 1) ^ 1u);
            mbar_expect_tx(st_full(s), 2 * TILE);
            tma_load_3d(sY(s, 0), &map_qkv, st_full(s), h * HD, it * BQ, b);   // Q_i
            tma_load_3d(sY(s, 1), &map_do, st_full(s), h * HD, it * BQ, b);    // dO_i
          }
        }
      }
    } else if (warp == 2) {
      // Second UMMA issuer: the score tiles T0 = Q_i K^T, T1 = dO_i V^T of the NEXT block. They are independent of the accumulating
      // UMMAs of warp 1 (different TMEM columns), and issuing them from their own thread matters: warp 1 blocks for ~1400 cycles
      // while it feeds the 24 dV/dK/dQ UMMAs of block it-1 into the pipe, and T(it+1) would be queued only after that — the row
      // warps then waited ~400 cycles per block for it (timeline). From here it is queued the moment the rows release T(it).
      if (elect_one()) {
        constexpr uint32_t idesc_t = umma_idesc_bf16(128, BQ, 0, 0);
        const uint64_t d0 = umma_smem_desc(sbase, 0, 1024, 128);
        const uint64_t dK_k = d0 + (OFF_K >> 4);
        const uint64_t dV_k = d0 + (OFF_V >> 4);
        uint32_t g = 0, item_n = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x, ++item_n) {
          mbar_wait(res_full, item_n & 1);
          for (int it = 0; it < niter; ++it, ++g) {
            if (g > 0) mbar_wait(t_free, (g - 1) & 1);   // the rows hold T(g-1) in registers
            const int s = g % NS;
            mbar_wait(st_full(s), (g / NS) & 1);
            tc_fence_after();
            const uint64_t dq = d0 + ((OFF_Y + (s * 2) * TILE) >> 4), dd = d0 + ((OFF_Y + (s * 2 + 1) * TILE) >> 4);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem + T0, dq + 2 * k, dK_k + 2 * k, idesc_t, k > 0);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem + T1, dd + 2 * k, dV_k + 2 * k, idesc_t, k > 0);
            umma_commit(t_full);
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {   // UMMA issuer of the accumulating products
        constexpr uint32_t idesc_q = umma_idesc_bf16(128, HD, 0, 1);
        constexpr uint32_t idesc_a = umma_idesc_bf16(128, HD, 1, 1);
        const uint64_t d0 = umma_smem_desc(sbase, 0, 1024, 128);
        const uint64_t dK_k = d0 + (OFF_K >> 4);
        auto dQ_ = [&](int s) { return d0 + ((OFF_Y + (s * 2) * TILE) >> 4); };
        auto dO_ = [&](int s) { return d0 + ((OFF_Y + (s * 2 + 1) * TILE) >> 4); };
        const uint64_t dmn = umma_smem_desc(sbase, 128 * 128, 1024, 128);
        const uint64_t dP_mn = dmn + (OFF_P >> 4);
        uint32_t gi = 0, item_n = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x, ++item_n) {
#ifdef VMAE_ATT_TRACE
          att_traced = att_cta && static_cast<int>(item_n) == g_att_trace_cta[1];
#endif
          mbar_wait(res_full, item_n & 1);
          for (int it = 0; it < niter; ++it, ++gi) {
            const int s = gi % NS;
            const uint32_t eb = gi & 1u;
            const uint64_t dE_k = d0 + ((OFF_E + eb * 2 * TILE) >> 4), dE_mn = dmn + ((OFF_E + eb * 2 * TILE) >> 4);
            ATT_STAMP(1, it, 0);
            mbar_wait(p_ready, gi & 1);
            tc_fence_after();
            ATT_STAMP(1, it, 3);
#pragma unroll
            for (int kk = 0; kk < BQ / 16; ++kk)   // dV += P^T dO_i
              umma_bf16(tmem + DV, dP_mn + kk * 128, dO_(s) + kk * 128, idesc_a, (it > 0 || kk > 0));
            umma_commit(p_free);                   // the P tile may be rewritten
#pragma unroll
            for (int kk = 0; kk < BQ / 16; ++kk)   // dK += dS^T Q_i
              umma_bf16(tmem + DK, dE_mn + kk * 128, dQ_(s) + kk * 128, idesc_a, (it > 0 || kk > 0));
            umma_commit(st_empty(s));              // Q_i / dO_i are no longer read: the producer refills the stage
#pragma unroll
            for (int kk = 0; kk < 128 / 16; ++kk)  // dQ_i = dS K   -> dQ accumulator buffer gi & 1
              umma_bf16(tmem + DQ0 + eb * 64, dE_k + (kk >> 2) * 1024 + 2 * (kk & 3), dK_k + kk * 128, idesc_q, kk > 0);
            umma_commit(acc_done);
            if (it + 1 == niter) umma_commit(kv_free);
            ATT_STAMP(1, it, 4);
          }
        }
      }
    }
  } else {
    setmaxnreg_inc<ROW_REGS>();
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = scale * LOG2E_F;
    const f32x2 sl2_2 = pk2(sl2), dsc2 = pk2(DQB ? scale : 1.0f);
    auto stat_ptr = [&](const float* src, int b, int h, int it) { return src + (static_cast<long long>(b) * H + h) * L + min(it * BQ + row, L - 1); };
    // dQ tile `buf` (TMEM) -> this warp's staging box inside the idle dS buffer `buf` (the one MMA(buf's block) has just released)
    auto box_of = [&](uint32_t buf) { return sE(buf) + (warp - 4) * (32 * DQC * 4); };
    auto stage_dq = [&](uint32_t buf) {
      uint32_t r[32];
      tmem_ld32(tmem + DQ0 + buf * 64 + lane_off + grp * DQC, r);
      tmem_wait_ld();
      const uint32_t box = box_of(buf);
      if (DQB) {   // bf16 box: 32 rows x 64 B, SWIZZLE_64B (16-byte slot j of row r at r*64 + ((j ^ ((r >> 1) & 3)) << 4))
#pragma unroll
        for (int j = 0; j < 4; ++j)
          fb_sts_u4(box + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), fb_pack(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1])),
                    fb_pack(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])), fb_pack(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])),
                    fb_pack(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])));
      } else {
#pragma unroll
        for (int j = 0; j < DQC / 4; ++j)
          fb_sts_f4(box + lane * 128 + ((j ^ (lane & 7)) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                    __uint_as_float(r[4 * j + 3]));
      }
    };
    auto issue_dq = [&](uint32_t buf, int b, int h, int qb) {
      if (lane == 0) {
        fb_tma_reduce_add_3d(&map_dq, box_of(buf), h * HD + grp * DQC, qb * BQ + quad * 32, b);
        tma_store_commit();
      }
    };

    uint32_t gi = 0;
    float nxt_lse = 0.f, nxt_delta = 0.f;
    uint32_t ts[NCH][32], tp[NCH][32];   // this thread's 64 columns of T0 (scores) and T1 (dP) of the current block
    auto load_t = [&]() {                // issue the TMEM loads of the block whose t_full phase is g (asynchronous)
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld32(tmem + T0 + lane_off + grp * COLS + c * 32, ts[c]);
        tmem_ld32(tmem + T1 + lane_off + grp * COLS + c * 32, tp[c]);
      }
    };
    if (static_cast<int>(blockIdx.x) < total) {
      int b, h, r0;
      decode(blockIdx.x, b, h, r0);
      nxt_lse = __ldg(stat_ptr(lse, b, h, 0));
      nxt_delta = __ldg(stat_ptr(delta, b, h, 0));
      mbar_wait(t_full, 0);
      tc_fence_after();
      load_t();
    }
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      int b, h, r0;
      decode(w, b, h, r0);
      const int g_row = r0 + row;
#ifdef VMAE_ATT_TRACE
      att_traced = att_cta && (w - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) == g_att_trace_cta[1];
#endif
      const bool more_items = w + static_cast<int>(gridDim.x) < total;
      int nb = b, nh = h, nr0 = r0;
      if (more_items) decode(w + gridDim.x, nb, nh, nr0);
      for (int it = 0; it < niter; ++it, ++gi) {
        const uint32_t drop_base = static_cast<uint32_t>((b * H + h) * L + it * BQ + row) * DROP_CQ + drop.seed_lo;
        const bool in_range = it * BQ + row < L;
        const float lse_row = in_range ? nxt_lse * LOG2E_F : INFINITY;
        const float delta_row = in_range ? nxt_delta : 0.f;
        const bool last = it + 1 == niter;
        if (!last || more_items) {
          nxt_lse = __ldg(stat_ptr(lse, last ? nb : b, last ? nh : h, last ? 0 : it + 1));
          nxt_delta = __ldg(stat_ptr(delta, last ? nb : b, last ? nh : h, last ? 0 : it + 1));
        }
        const f32x2 nlse2 = pk2(-lse_row), ndelta2 = pk2(-delta_row);
        ATT_STAMP_T(0, it, 0, 128);
        // the score tiles of this block were requested at the end of the previous iteration (or in the prologue)
        tmem_wait_ld();
        tc_fence_before();
        if (lane == 0) mbar_arrive(t_free);
        ATT_STAMP_T(0, it, 1, 128);
        const uint32_t sEb = sE(gi & 1u);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t pp[16], dd[16];   // packed bf16 pairs of this chunk: P and dS
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1;
            up2(fma2(pk2(__uint_as_float(ts[c][i]), __uint_as_float(ts[c][i + 1])), sl2_2, nlse2), x0, x1);
            float p0, p1;
            if ((((i >> 1) * FB2_POLY) & 7) < FB2_POLY) fb_exp2_poly2(x0, x1, p0, p1);   // FB2_POLY of every 8 pairs on the FMA pipe
            else { p0 = fb_ex2(x0); p1 = fb_ex2(x1); }
            f32x2 dp2 = pk2(__uint_as_float(tp[c][i]), __uint_as_float(tp[c][i + 1]));
            float pd0 = p0, pd1 = p1;
            if (DROP) {
              const int col = r0 + grp * COLS + c * 32 + i;
              const uint32_t wd = drop_mix32(drop_base + static_cast<uint32_t>(col >> 1) * DROP_CK) ^ drop.seed_hi;
              const bool d0 = (wd & 0xffffu) < drop.thr, d1 = (wd >> 16) < drop.thr;
              float a0, a1;
              up2(dp2, a0, a1);
              a0 = d0 ? 0.f : a0 * drop.rp; pd0 = d0 ? 0.f : pd0;
              a1 = d1 ? 0.f : a1 * drop.rp; pd1 = d1 ? 0.f : pd1;
              dp2 = pk2(a0, a1);
            }
            float s0, s1;
            if (DQB) up2(mul2(mul2(pk2(p0, p1), add2(dp2, ndelta2)), dsc2), s0, s1);   // dS carries the softmax scale
            else up2(mul2(pk2(p0, p1), add2(dp2, ndelta2)), s0, s1);
            pp[i >> 1] = fb_pack(pd0, pd1);
            dd[i >> 1] = fb_pack(s0, s1);
          }
          const int col = grp * COLS + c * 32;
          const uint32_t off = (col >> 6) * TILE + row * 128;
          if (c == 0) {
            ATT_STAMP_T(0, it, 2, 128);
            // the P tile is free once the dV UMMAs of the previous block have retired (a chunk of math after its p_ready);
            // this block's dS buffer was last read by MMA(gi-2), whose completion the previous iteration waited for
            if (gi > 0) mbar_wait(p_free, (gi - 1) & 1);
            // this warp's dS sub-tile of buffer gi & 1 doubled as its dQ staging box one block ago: that bulk reduction (issued
            // ~a chunk of math ago) must have read the box before the first store lands on it
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            ATT_STAMP_T(0, it, 3, 128);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t sw = ((((col & 63) >> 3) + i) ^ (row & 7)) * 16;
            fb_sts_u4(sP + off + sw, pp[4 * i], pp[4 * i + 1], pp[4 * i + 2], pp[4 * i + 3]);
            fb_sts_u4(sEb + off + sw, dd[4 * i], dd[4 * i + 1], dd[4 * i + 2], dd[4 * i + 3]);
          }
        }
        ATT_STAMP_T(0, it, 4, 128);
        if (it > 0) {
          // dQ tile of the previous block (MMA(it-1) retired while this block's tiles were being produced): TMEM -> staging box;
          // published by the same proxy fence as the P / dS tiles, reduced into the accumulator underneath MMA(it)
          mbar_wait(acc_done, (gi - 1) & 1);
          tc_fence_after();
          ATT_STAMP_T(0, it, 5, 128);
          stage_dq((gi - 1) & 1u);
        }
        ATT_STAMP_T(0, it, 6, 128);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        if (it > 0) issue_dq((gi - 1) & 1u, b, h, it - 1);
        // request the next block's score tiles (T(gi+1) was issued when this block's tiles were released and ran behind MMA(gi-1)):
        // the loads land underneath the loop top. The last block of an item leaves this to the item tail: the next item's first T
        // needs the next K / V, which is loaded only after this block's MMAs.
        if (!last) {
          mbar_wait(t_full, (gi + 1) & 1);
          tc_fence_after();
          load_t();
        }
        ATT_STAMP_T(0, it, 7, 128);
      }
      // item tail
      ATT_STAMP_T(0, niter, 0, 128);
      mbar_wait(acc_done, (gi - 1) & 1);
      tc_fence_after();
      stage_dq((gi - 1) & 1u);
      fence_proxy_async_smem();
      __syncwarp();
      issue_dq((gi - 1) & 1u, b, h, niter - 1);
      const long long row_off = (static_cast<long long>(b) * L + g_row) * (3 * H * HD);
      const int acol = grp * COLS;
      const int part = 1 + (acol >> 6);
      const uint32_t tA = tmem + ((acol >> 6) == 0 ? DK : DV) + (acol & 63);
      uint32_t ra[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld32(tA + lane_off + c * 32, ra[c]);
      tmem_wait_ld();
      if (g_row < L) {
        const float f = (part == 1) ? (DQB ? 1.0f : scale) : (DROP ? drop.rp : 1.0f);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          __nv_bfloat16* dst = dqkv + row_off + (part * H + h) * HD + (acol & 63) + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 pk;
            pk.x = fb_pack(__uint_as_float(ra[c][i]) * f, __uint_as_float(ra[c][i + 1]) * f);
            pk.y = fb_pack(__uint_as_float(ra[c][i + 2]) * f, __uint_as_float(ra[c][i + 3]) * f);
            pk.z = fb_pack(__uint_as_float(ra[c][i + 4]) * f, __uint_as_float(ra[c][i + 5]) * f);
            pk.w = fb_pack(__uint_as_float(ra[c][i + 6]) * f, __uint_as_float(ra[c][i + 7]) * f);
            *reinterpret_cast<uint4*>(dst + i) = pk;
          }
        }
      }
      ATT_STAMP_T(0, niter, 1, 128);
      if (more_items) {   // the next item's first score tiles (their UMMAs were issued once its K / V had landed)
        mbar_wait(t_full, gi & 1);
        tc_fence_after();
        load_t();
      }
    }
    if (lane == 0) tma_store_wait<0>();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// dq (bf16, q slot of dqkv[B, L, 3, H, 64]) = scale * dq_accum (fp32 [B, L, H*64])
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, long long rows, int H,
                                                              float scale) {
  pdl_wait();
  pdl_launch_dependents();
  const int vec_per_row = H * fb::HD / 4;
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += gridDim.x * 256LL) {
    const long long r = idx / vec_per_row;
    const int c = static_cast<int>(idx % vec_per_row) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(acc + r * (H * fb::HD) + c));
    uint2 u;
    u.x = fb_pack(v.x * scale, v.y * scale);
    u.y = fb_pack(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(dqkv + r * (3 * H * fb::HD) + c) = u;
  }
}

#ifdef VMAE_ATT_TRACE
int att_trace_select_fused(int bx, int by, int bz) { return att_trace_select_tu(bx, by, bz, -1); }
int att_trace_read_fused(long long* host) { return att_trace_read_tu(host); }
#endif

// Returns 0 on success, <0 on error. dq_accum: fp32 [B, L, H*64] workspace (zeroed here).
int attn_bwd_fused_hd64(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, float* dq_accum, int B, int L, int H,
                        float scale, float p_drop, unsigned long long seed, cudaStream_t stream, bool accum_zeroed) {
  using namespace fb;
  CUtensorMap m_qkv, m_do, m_dq;
  {
    const uint64_t dims[3] = {(uint64_t)3 * H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)3 * H * HD * 2, (uint64_t)L * 3 * H * HD * 2};
    const uint32_t box[3] = {(uint32_t)HD, 128, 1};
    int rc = make_tmap_bf16(&m_qkv, qkv, 3, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)H * HD * 2, (uint64_t)L * H * HD * 2};
    const uint32_t box[3] = {(uint32_t)HD, 128, 1};
    int rc = make_tmap_bf16(&m_do, dout, 3, dims, str, box, 128);
    if (rc) return rc;
  }
  constexpr int G = VMAE_FB_GROUPS;
  {
    const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)H * HD * 4, (uint64_t)L * H * HD * 4};
    const uint32_t box[3] = {(uint32_t)Cfg<G>::DQC, 32, 1};   // one reduction per row warp: 32 query rows x DQC head-dim columns
    int rc = make_tmap(&m_dq, dq_accum, VMAE_DTYPE_F32, 3, dims, str, box, Cfg<G>::DQC * 4 == 128 ? 128 : 0);
    if (rc) return rc;
  }
  const long long items = (long long)B * H * ((L + 127) / 128);
  const int grid = (int)(items < num_sms() ? items : num_sms());
  const AttnDrop drop = make_attn_drop(p_drop, seed);
  static const bool use_v2 = [] { const char* e = getenv("VMAE_FB_V2"); return e ? atoi(e) != 0 : true; }();
  // opt-in: measured 867 -> 834 us at B=64 L=1568, but the dQ error (max-norm, vs fp32 reference) grows from 4.2e-3 to 1.2e-2 —
  // 13 bf16 additions per element. The default keeps the fp32 accumulator: parity before 4 %.
  static const bool dq_bf16 = [] { const char* e = getenv("VMAE_FB_DQ_BF16"); return e ? atoi(e) != 0 : false; }();
  if (!(use_v2 && G == 2 && dq_bf16) && !accum_zeroed) VMAE_CUDA(cudaMemsetAsync(dq_accum, 0, (size_t)B * L * H * HD * sizeof(float), stream));
  if (use_v2 && G == 2 && dq_bf16) {
    // bf16 reduction straight into the q slot of dqkv [B, L, 3, H, 64]: zero the slot (strided), no accumulator memset, no convert kernel
    CUtensorMap m_dqb;
    {
      const uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
      const uint64_t str[2] = {(uint64_t)3 * H * HD * 2, (uint64_t)L * 3 * H * HD * 2};
      const uint32_t box[3] = {32, 32, 1};
      int rc = make_tmap_bf16(&m_dqb, dqkv, 3, dims, str, box, 64);
      if (rc) return rc;
    }
    VMAE_CUDA(cudaMemset2DAsync(dqkv, (size_t)3 * H * HD * 2, 0, (size_t)H * HD * 2, (size_t)B * L, stream));
    if (drop.thr > 0) {
      VMAE_CUDA(cudaFuncSetAttribute(attn_bwd_fused2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb2::SMEM_BYTES));
      VMAE_CUDA(launch_pdl(PDL_ATTN_FUSED, attn_bwd_fused2_kernel<true, true>, dim3(grid), dim3(fb2::THREADS), fb2::SMEM_BYTES, stream, m_qkv, m_do, m_dqb, lse,
                           delta, (__nv_bfloat16*)dqkv, L, H, B, scale, drop));
    } else {
      VMAE_CUDA(cudaFuncSetAttribute(attn_bwd_fused2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb2::SMEM_BYTES));
      VMAE_CUDA(launch_pdl(PDL_ATTN_FUSED, attn_bwd_fused2_kernel<false, true>, dim3(grid), dim3(fb2::THREADS), fb2::SMEM_BYTES, stream, m_qkv, m_do, m_dqb, lse,
                           delta, (__nv_bfloat16*)dqkv, L, H, B, scale, drop));
    }
    VMAE_LAUNCH_CHECK();
    return VMAE_OK;
  }
  if (use_v2 && G == 2) {
    if (drop.thr > 0) {
      VMAE_CUDA(cudaFuncSetAttribute(attn_bwd_fused2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb2::SMEM_BYTES));
      VMAE_CUDA(launch_pdl(PDL_ATTN_FUSED, attn_bwd_fused2_kernel<true, false>, dim3(grid), dim3(fb2::THREADS), fb2::SMEM_BYTES, stream, m_qkv, m_do, m_dq, lse,
                           delta, (__nv_bfloat16*)dqkv, L, H, B, scale, drop));
    } else {
      VMAE_CUDA(cudaFuncSetAttribute(attn_bwd_fused2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb2::SMEM_BYTES));
      VMAE_CUDA(launch_pdl(PDL_ATTN_FUSED, attn_bwd_fused2_kernel<false, false>, dim3(grid), dim3(fb2::THREADS), fb2::SMEM_BYTES, stream, m_qkv, m_do, m_dq, lse,
                           delta, (__nv_bfloat16*)dqkv, L, H, B, scale, drop));
    }
  } else {
  static const bool dqred = [] { const char* e = getenv("VMAE_FB_DQRED"); return e ? atoi(e) != 0 : false; }();   // measured: 1385 vs 1001 us — 16-byte lane reductions to 32 different rows cost twice the L2 sector operations of the TMA box
#define FB_LAUNCH(DROPV, REDV)                                                                                                                   \
  do {                                                                                                                                           \
    VMAE_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<G, DROPV, REDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));            \
    VMAE_CUDA(launch_pdl(PDL_ATTN_FUSED, attn_bwd_fused_kernel<G, DROPV, REDV>, dim3(grid), dim3(Cfg<G>::THREADS), SMEM_BYTES, stream, m_qkv,  \
                         m_do, m_dq, lse, delta, (__nv_bfloat16*)dqkv, dq_accum, L, H, B, scale, drop));                                        \
  } while (0)
  if (drop.thr > 0) { if (dqred) FB_LAUNCH(true, true); else FB_LAUNCH(true, false); }
  else { if (dqred) FB_LAUNCH(false, true); else FB_LAUNCH(false, false); }
#undef FB_LAUNCH
  }
  VMAE_LAUNCH_CHECK();
  const long long rows = (long long)B * L;
  long long g = (rows * (H * HD / 4) + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  VMAE_CUDA(launch_pdl(PDL_ATTN_FUSED, attn_dq_convert_kernel, dim3((unsigned)g), dim3(256), 0, stream, (const float*)dq_accum, (__nv_bfloat16*)dqkv, rows, H, scale));
  return VMAE_OK;
}

}  // namespace vmae

#ifdef VMAE_ATT_TRACE
extern "C" int vmae_att_trace_select_fused(int bx, int by, int bz) { return vmae::att_trace_select_fused(bx, by, bz); }
extern "C" int vmae_att_trace_read_fused(long long* host) { return vmae::att_trace_read_fused(host); }
#endif
