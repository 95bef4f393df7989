// K6 — tubelet patch embedding of the VISIBLE tubes as ONE implicit-GEMM kernel (north_star bullet 1).
//
//   out[m, :] = W · patch(m) + bias + pos_t[tp(m)] + pos_s[l(m)]          m = (clip b, temporal slot tp, i-th visible patch)
//
// i.e. nn.Conv3d(3, E, (pt,16,16), stride (pt,16,16)) evaluated only where the tube mask is 0, the factorised positional
// embedding added and the result written in gathered (visible-token) order — what `gather_tubelets_kernel` + GEMM + residual rows
// did in three passes with a 31 MB patch matrix in between. No patch matrix exists here:
//   * the A operand is gathered by TMA: one 5-D box (16 px x 4 py x 1 frame x 1 channel x 1 clip = 64 fp32) per tubelet row and
//     64-wide K block, addressed straight in the clip [B,3,T,H,W] with the (b, tp, l) of that row (l from vis_idx) — the unit
//     zero-fills nothing and copies only visible pixels;
//   * eight converter warps turn the fp32 staging rows into the bf16, 128B-swizzled K-major UMMA tile (the clip stays fp32 in
//     HBM, as the loader delivers it, and is read once: 10 % of it);
//   * tcgen05.mma accumulates all E <= 384 output columns of the 128 rows in TMEM (two UMMAs per K step: N = 256 + N = E-256),
//     so a row is gathered exactly once;
//   * the epilogue adds bias and both positional rows and writes fp32.
// The weight gradient kernel below gathers the same boxes (MN-major B operand) against dY^T; the bias / positional gradients
// are column and scatter sums of dY (vmae_colsum_accum, vmae_pos_rows_bwd).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/vmae.h"

namespace vmae {
namespace te {
constexpr int BM = 128;                  // tubelet rows per CTA
constexpr int BK = 64;                   // K block = one gather box per row: 4 pixel rows x 16 pixels of one (channel, frame)
constexpr int NMAX = 384;
constexpr int STAGES = 2;
constexpr int G_BYTES = BM * BK * 4;     // fp32 staging of the gather boxes: 32 KB
constexpr int A_BYTES = BM * BK * 2;     // bf16 A tile: 16 KB
constexpr int B_BYTES = NMAX * BK * 2;   // bf16 weight tile: 48 KB
constexpr int STAGE_BYTES = G_BYTES + A_BYTES + B_BYTES;
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack
constexpr int THREADS = 320;             // warp 0: TMA producer (all lanes issue gather boxes), warp 1: UMMA, warps 2-9: convert + epilogue
constexpr int PS = 16;

struct Geom {
  int B, T, H, W, pt, n_vis, E, rows, Tp, wp, K;
};

// row m -> clip coordinates of its tube: b, first frame, top-left pixel
__device__ __forceinline__ void row_coords(const Geom& g, const int* __restrict__ vis_idx, int m, int& b, int& tp, int& l) {
  const int i = m % g.n_vis;
  tp = (m / g.n_vis) % g.Tp;
  b = m / (g.n_vis * g.Tp);
  l = __ldg(vis_idx + static_cast<long long>(b) * g.n_vis + i);
}

// fp32 staging rows (64 floats) -> bf16 K-major SW128 tile rows (128 B): two rows per warp instruction, 16 lanes x float4 each
__device__ __forceinline__ void convert_rows(uint32_t staging, uint32_t tile, int row0, int nrows, int lane) {
  const int l16 = lane & 15, sub = lane >> 4;
  for (int r = row0 + sub; r < row0 + nrows; r += 2) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(staging + r * 256 + l16 * 16) : "memory");
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    const uint32_t dst = tile + r * 128 + ((((l16 >> 1) ^ (r & 7))) << 4) + (l16 & 1) * 8;
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(*reinterpret_cast<uint32_t*>(&lo)), "r"(*reinterpret_cast<uint32_t*>(&hi)) : "memory");
  }
}
}  // namespace te

__global__ void __launch_bounds__(te::THREADS, 1)
tubelet_embed_fwd_kernel(const __grid_constant__ CUtensorMap map_clip, const __grid_constant__ CUtensorMap map_w, const int* __restrict__ vis_idx,
                         const float* __restrict__ bias, const float* __restrict__ tpos, const float* __restrict__ spos, float* __restrict__ out,
                         const te::Geom g, const int rows_per_cta) {
  using namespace te;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sG = [&](int s) { return sbase + s * STAGE_BYTES; };
  auto sA = [&](int s) { return sbase + s * STAGE_BYTES + G_BYTES; };
  auto sB = [&](int s) { return sbase + s * STAGE_BYTES + G_BYTES + A_BYTES; };
  const uint32_t bar = sbase + OFF_BAR;
  auto full_g = [&](int s) { return bar + 8 * s; };
  auto full_a = [&](int s) { return bar + 16 + 8 * s; };
  auto full_b = [&](int s) { return bar + 32 + 8 * s; };
  auto empty = [&](int s) { return bar + 48 + 8 * s; };
  const uint32_t acc_full = bar + 64, tmem_ptr_addr = bar + 72;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A CTA owns rows_per_cta <= 128 tubelet rows (the UMMA tile is always 128 rows; the unused rows cost tensor time only, and this
  // kernel is bound by the TMA gather: ~90 cycles per 256-byte box). The host sizes rows_per_cta so that ONE wave covers all SMs.
  const int m0 = blockIdx.x * rows_per_cta;
  const int m_end = min(m0 + rows_per_cta, g.rows);
  const int nkb = g.K / BK;
  const int n1 = g.E < 256 ? g.E : 256, n2 = g.E - n1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_clip);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_g(s), 1);
      mbar_init(full_a(s), 8);
      mbar_init(full_b(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_addr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;

  if (warp == 0) {
    // ---- producer: weight tile by one elected lane, gather boxes by all lanes (rows lane, lane+32, lane+64, lane+96)
    int cb[4], ct[4], cy[4], cx[4];
    bool ok[4];
    const int valid = m_end - m0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + lane + 32 * r;
      ok[r] = m < m_end;
      int b = 0, tp = 0, l = 0;
      if (ok[r]) row_coords(g, vis_idx, m, b, tp, l);
      cb[r] = b; ct[r] = tp * g.pt; cy[r] = (l / g.wp) * PS; cx[r] = (l % g.wp) * PS;
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      if (lane == 0) {
        mbar_wait(empty(s), ((kb / STAGES) & 1) ^ 1u);
        mbar_expect_tx(full_b(s), static_cast<uint32_t>(g.E) * 128u);
        for (int j = 0; j < g.E / 64; ++j) tma_load_2d(sB(s) + j * (64 * 128), &map_w, full_b(s), kb * BK, j * 64);   // [64 out x 64 k] boxes
        mbar_expect_tx(full_g(s), static_cast<uint32_t>(valid) * 256u);
      }
      __syncwarp();
      // K index order of the Conv3d weight: (c, dt, py, px): K block kb = channel c, frame dt, pixel rows py0 .. py0+3
      const int per_c = g.pt * (PS / 4);
      const int c = kb / per_c, dt = (kb % per_c) / (PS / 4), py0 = (kb % (PS / 4)) * 4;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (ok[r]) tma_load_5d(sG(s) + (lane + 32 * r) * 256, &map_clip, full_g(s), cx[r], cy[r] + py0, ct[r] + dt, c, cb[r]);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc1 = umma_idesc_bf16(BM, n1, 0, 0);
      const uint32_t idesc2 = umma_idesc_bf16(BM, n2 > 0 ? n2 : 16, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full_b(s), ph);
        mbar_wait(full_a(s), ph);
        tc_fence_after();
        const uint64_t dA = umma_smem_desc_sw128(sA(s), 0, 1024);
        const uint64_t dB1 = umma_smem_desc_sw128(sB(s), 0, 1024);
        const uint64_t dB2 = umma_smem_desc_sw128(sB(s) + 256 * 128, 0, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
          umma_bf16(tmem, dA + 2 * k, dB1 + 2 * k, idesc1, acc);
          if (n2 > 0) umma_bf16(tmem + 256, dA + 2 * k, dB2 + 2 * k, idesc2, acc);
        }
        umma_commit(empty(s));
      }
      umma_commit(acc_full);
    }
  } else {
    // ---- converter warps: 16 rows each per K block
    const int cw = warp - 2;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(full_g(s), (kb / STAGES) & 1);
      convert_rows(sG(s), sA(s), cw * 16, 16, lane);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(s));
    }
    // ---- epilogue: TMEM lane quadrant = warp & 3, column half = (warp - 2) >> 2
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int quad = warp & 3, half = cw >> 2;
    const int m = m0 + quad * 32 + lane;
    const bool live = m < m_end;
    int b = 0, tp = 0, l = 0;
    if (live) row_coords(g, vis_idx, m, b, tp, l);
    const int ncol = g.E / 2;
    for (int c0 = half * ncol; c0 < (half + 1) * ncol; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(quad * 32) << 16) + c0, acc);
      tmem_wait_ld();
      if (live) {
        float* dst = out + static_cast<long long>(m) * g.E + c0;
        const float* bp = bias + c0;
        const float* tq = tpos + static_cast<long long>(tp) * g.E + c0;
        const float* sq = spos + static_cast<long long>(l) * g.E + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bp + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(tq + j));
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(sq + j));
          float4 o;
          o.x = __uint_as_float(acc[j]) + b4.x + (t4.x + s4.x);
          o.y = __uint_as_float(acc[j + 1]) + b4.y + (t4.y + s4.y);
          o.z = __uint_as_float(acc[j + 2]) + b4.z + (t4.z + s4.z);
          o.w = __uint_as_float(acc[j + 3]) + b4.w + (t4.w + s4.w);
          *reinterpret_cast<float4*>(dst + j) = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ----------------------------------------------------------------------------------------------------------------------
// weight gradient: dW[n, k] += sum_m dY[m, n] * patch(m)[k]       (fp32, red.add; split over the tubelet rows)
// CTA = (128 output channels, 128 K columns, one slice of the rows); per K step (64 tubelet rows): A = dY^T tile (MN-major, TMA),
// B = the gathered patch columns (MN-major: [64 rows][64 k] x 2 chunks, same boxes + converter as the forward kernel).
// ----------------------------------------------------------------------------------------------------------------------
namespace tw {
constexpr int BN_OUT = 128;              // output channels (UMMA M)
constexpr int BKC = 128;                 // K columns of dW (UMMA N): two gather boxes per row
constexpr int BR = 64;                   // tubelet rows per step (UMMA K block)
constexpr int STAGES = 2;
constexpr int G_BYTES = BR * BKC * 4;    // 32 KB fp32 staging
constexpr int A_BYTES = BN_OUT * BR * 2; // 16 KB  dY^T
constexpr int B_BYTES = BKC * BR * 2;    // 16 KB  patches
constexpr int STAGE_BYTES = G_BYTES + A_BYTES + B_BYTES;
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 320;
}  // namespace tw

__global__ void __launch_bounds__(tw::THREADS, 1)
tubelet_embed_wgrad_kernel(const __grid_constant__ CUtensorMap map_clip, const __grid_constant__ CUtensorMap map_dy, const int* __restrict__ vis_idx,
                           float* __restrict__ dw, const te::Geom g, int steps_per_split) {
  using namespace tw;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sG = [&](int s) { return sbase + s * STAGE_BYTES; };
  auto sA = [&](int s) { return sbase + s * STAGE_BYTES + G_BYTES; };
  auto sB = [&](int s) { return sbase + s * STAGE_BYTES + G_BYTES + A_BYTES; };
  const uint32_t bar = sbase + OFF_BAR;
  auto full_g = [&](int s) { return bar + 8 * s; };
  auto full_a = [&](int s) { return bar + 16 + 8 * s; };   // dY^T tile landed (TMA)
  auto full_b = [&](int s) { return bar + 32 + 8 * s; };   // patch tile converted (8 warps)
  auto empty = [&](int s) { return bar + 48 + 8 * s; };
  const uint32_t acc_full = bar + 64, tmem_ptr_addr = bar + 72;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN_OUT, kc0 = blockIdx.y * BKC;
  const int total_steps = (g.rows + BR - 1) / BR;
  const int st0 = blockIdx.z * steps_per_split;
  int st1 = st0 + steps_per_split;
  st1 = st1 > total_steps ? total_steps : st1;
  const int nst = st1 - st0;   // host guarantees >= 1

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_clip);
    tma_prefetch_desc(&map_dy);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_g(s), 1);
      mbar_init(full_a(s), 1);
      mbar_init(full_b(s), 8);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_addr, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;

  if (warp == 0) {
    const int per_c = g.pt * (te::PS / 4);
    for (int i = 0; i < nst; ++i) {
      const int s = i % STAGES;
      const int row0 = (st0 + i) * BR;
      int valid = g.rows - row0;
      valid = valid > BR ? BR : valid;
      if (lane == 0) {
        mbar_wait(empty(s), ((i / STAGES) & 1) ^ 1u);
        // dY^T: two 64-channel chunks of [64 rows x 64 channels]; rows past the end are zero-filled by the TMA unit
        mbar_expect_tx(full_a(s), A_BYTES);
        tma_load_2d(sA(s), &map_dy, full_a(s), n0, row0);
        tma_load_2d(sA(s) + BR * 128, &map_dy, full_a(s), n0 + 64, row0);
        mbar_expect_tx(full_g(s), static_cast<uint32_t>(valid) * 2u * 256u);
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int rr = lane + 32 * r;
        const int m = row0 + rr;
        if (m < g.rows) {
          int b, tp, l;
          te::row_coords(g, vis_idx, m, b, tp, l);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int kb = (kc0 >> 6) + j;   // 64-wide K block index of the Conv3d weight
            const int c = kb / per_c, dt = (kb % per_c) / (te::PS / 4), py0 = (kb % (te::PS / 4)) * 4;
            tma_load_5d(sG(s) + (j * BR + rr) * 256, &map_clip, full_g(s), (l % g.wp) * te::PS, (l / g.wp) * te::PS + py0, tp * g.pt + dt, c, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN_OUT, BKC, 1, 1);
      for (int i = 0; i < nst; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(full_a(s), ph);
        mbar_wait(full_b(s), ph);
        tc_fence_after();
        // MN-major operands: LBO = distance between the 64-wide MN chunks (64 rows x 128 B), SBO = 8 rows; a UMMA K step = 16 rows
        const uint64_t dA = umma_smem_desc_sw128(sA(s), BR * 128, 1024);
        const uint64_t dB = umma_smem_desc_sw128(sB(s), BR * 128, 1024);
#pragma unroll
        for (int k = 0; k < BR / 16; ++k) umma_bf16(tmem, dA + k * ((16 * 128) >> 4), dB + k * ((16 * 128) >> 4), idesc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty(s));
      }
      umma_commit(acc_full);
    }
  } else {
    const int cw = warp - 2;
    for (int i = 0; i < nst; ++i) {
      const int s = i % STAGES;
      const int row0 = (st0 + i) * BR;
      int valid = g.rows - row0;
      valid = valid > BR ? BR : valid;
      mbar_wait(full_g(s), (i / STAGES) & 1);
      // 2 chunks x 64 rows = 128 staging rows; warp cw converts 16 of them; rows past the end of the tensor become zeros
      const int base = cw * 16;                  // staging row index: chunk j = base / 64, row = base % 64
      const int j = base >> 6, r0 = base & 63;
      if (r0 + 16 <= valid) {
        te::convert_rows(sG(s) + j * BR * 256, sB(s) + j * BR * 128, r0, 16, lane);
      } else {
        const int l16 = lane & 15, sub = lane >> 4;
        for (int r = r0 + sub; r < r0 + 16; r += 2) {
          if (r < valid) {
            te::convert_rows(sG(s) + j * BR * 256, sB(s) + j * BR * 128, r, 1, l16);   // one row, lanes 0-15 pattern
          } else {
            const uint32_t dst = sB(s) + j * BR * 128 + r * 128 + ((((l16 >> 1) ^ (r & 7))) << 4) + (l16 & 1) * 8;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(0u), "r"(0u) : "memory");
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_b(s));
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int quad = warp & 3, half = cw >> 2;
    const int n = n0 + quad * 32 + lane;
    for (int c0 = half * (BKC / 2); c0 < (half + 1) * (BKC / 2); c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(quad * 32) << 16) + c0, acc);
      tmem_wait_ld();
      if (n < g.E) {
        float* dst = dw + static_cast<long long>(n) * g.K + kc0 + c0;
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q), "f"(__uint_as_float(acc[q])), "f"(__uint_as_float(acc[q + 1])),
                       "f"(__uint_as_float(acc[q + 2])), "f"(__uint_as_float(acc[q + 3]))
                       : "memory");
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

static int make_clip_map(CUtensorMap* m, const float* clip, const te::Geom& g) {
  const uint64_t dims[5] = {(uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T, 3, (uint64_t)g.B};
  const uint64_t str[4] = {(uint64_t)g.W * 4, (uint64_t)g.H * g.W * 4, (uint64_t)g.T * g.H * g.W * 4, (uint64_t)3 * g.T * g.H * g.W * 4};
  const uint32_t box[5] = {16, 4, 1, 1, 1};
  return make_tmap(m, clip, VMAE_DTYPE_F32, 5, dims, str, box, 0);
}

static int check_geom(te::Geom& g, int B, int T, int H, int W, int pt, int ps, int n_vis, int E, const char* who) {
  if (!(B > 0 && pt > 0 && T % pt == 0 && n_vis > 0)) return fail(VMAE_ERR_BAD_ARG, "%s: bad geometry", who);
  if (ps != 16 || H % 16 != 0 || W % 16 != 0) return fail(VMAE_ERR_UNSUPPORTED, "%s: spatial patch must be 16 (got %d) and H, W multiples of 16", who, ps);
  if (E % 64 != 0 || E < 64 || E > te::NMAX) return fail(VMAE_ERR_UNSUPPORTED, "%s: embed dim %d (multiple of 64, <= %d)", who, E, te::NMAX);
  g.B = B; g.T = T; g.H = H; g.W = W; g.pt = pt; g.n_vis = n_vis; g.E = E;
  g.Tp = T / pt; g.wp = W / 16; g.K = 3 * pt * 256;
  const long long rows = (long long)B * g.Tp * n_vis;
  if (rows >= (1ll << 31)) return fail(VMAE_ERR_BAD_ARG, "%s: too many rows", who);
  g.rows = (int)rows;
  return VMAE_OK;
}

}  // namespace vmae

using namespace vmae;

extern "C" int vmae_tubelet_embed_fwd(const float* clip, const int* vis_idx, const void* weight_bf16, const float* bias, const float* temporal_pos,
                                      const float* spatial_pos, float* out, int B, int T, int H, int W, int pt, int ps, int n_vis, int E, void* stream) {
  VMAE_CHECK_ARG(clip && vis_idx && weight_bf16 && temporal_pos && spatial_pos && out, "vmae_tubelet_embed_fwd: null pointer");
  te::Geom g;
  int rc = check_geom(g, B, T, H, W, pt, ps, n_vis, E, "vmae_tubelet_embed_fwd");
  if (rc) return rc;
  CUtensorMap m_clip, m_w;
  rc = make_clip_map(&m_clip, clip, g);
  if (rc) return rc;
  {
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)E};
    const uint64_t str[1] = {(uint64_t)g.K * 2};
    const uint32_t box[2] = {64, 64};
    rc = make_tmap_bf16(&m_w, weight_bf16, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  static thread_local bool attr_set = false;
  if (!attr_set) {
    VMAE_CUDA(cudaFuncSetAttribute(tubelet_embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, te::SMEM_BYTES));
    attr_set = true;
  }
  int rpc = (g.rows + num_sms() - 1) / num_sms();   // one wave: every SM gathers an equal share of the rows
  rpc = (rpc + 7) / 8 * 8;
  if (rpc > te::BM) rpc = te::BM;
  const int grid = (g.rows + rpc - 1) / rpc;
  tubelet_embed_fwd_kernel<<<grid, te::THREADS, te::SMEM_BYTES, (cudaStream_t)stream>>>(m_clip, m_w, vis_idx, bias, temporal_pos, spatial_pos, out, g, rpc);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}

extern "C" int vmae_tubelet_embed_wgrad(const float* clip, const int* vis_idx, const void* dy_bf16, float* dw, int B, int T, int H, int W, int pt, int ps,
                                        int n_vis, int E, void* stream) {
  VMAE_CHECK_ARG(clip && vis_idx && dy_bf16 && dw, "vmae_tubelet_embed_wgrad: null pointer");
  te::Geom g;
  int rc = check_geom(g, B, T, H, W, pt, ps, n_vis, E, "vmae_tubelet_embed_wgrad");
  if (rc) return rc;
  VMAE_CHECK_ARG(E % 128 == 0 && g.K % 128 == 0, "vmae_tubelet_embed_wgrad: E and K must be multiples of 128");
  CUtensorMap m_clip, m_dy;
  rc = make_clip_map(&m_clip, clip, g);
  if (rc) return rc;
  {
    const uint64_t dims[2] = {(uint64_t)E, (uint64_t)g.rows};
    const uint64_t str[1] = {(uint64_t)E * 2};
    const uint32_t box[2] = {64, 64};
    rc = make_tmap_bf16(&m_dy, dy_bf16, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  static thread_local bool attr_set = false;
  if (!attr_set) {
    VMAE_CUDA(cudaFuncSetAttribute(tubelet_embed_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tw::SMEM_BYTES));
    attr_set = true;
  }
  const int total_steps = (g.rows + tw::BR - 1) / tw::BR;
  const int tiles = (E / tw::BN_OUT) * (g.K / tw::BKC);
  int splits = num_sms() / tiles;                               // one wave: at most one CTA per SM
  if (splits > total_steps) splits = total_steps;
  if (splits < 1) splits = 1;
  const int per = (total_steps + splits - 1) / splits;
  splits = (total_steps + per - 1) / per;                       // no empty split
  dim3 grid(E / tw::BN_OUT, g.K / tw::BKC, splits);
  tubelet_embed_wgrad_kernel<<<grid, tw::THREADS, tw::SMEM_BYTES, (cudaStream_t)stream>>>(m_clip, m_dy, vis_idx, dw, g, per);
  VMAE_LAUNCH_CHECK();
  return VMAE_OK;
}
