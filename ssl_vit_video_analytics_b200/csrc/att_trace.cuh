// Developer timeline support for the attention kernels (tools/trace_attn.cu): clock64 stamps of one CTA, compiled in only
// with -DVMAE_ATT_TRACE. One private copy per translation unit (no relocatable device code needed); include inside namespace vmae.
// Layout: [role 0 = one row warp's lane 0 | 1 = UMMA thread][iteration < 32][slot < 8].
#pragma once
#ifdef VMAE_ATT_TRACE
static __device__ long long g_att_trace[2 * 32 * 8];
static __device__ int g_att_trace_cta[4];   // block x, y, z, backward mode filter (-1 = any)
#define ATT_TRACE_INIT()                                                                                               \
  const bool att_traced = blockIdx.x == g_att_trace_cta[0] && blockIdx.y == g_att_trace_cta[1] && blockIdx.z == g_att_trace_cta[2]
#define ATT_STAMP_T(role, it, slot, row_thread)                                                                        \
  do {                                                                                                                 \
    if (att_traced && (it) < 32 && ((role) == 1 || threadIdx.x == (row_thread))) g_att_trace[((role) * 32 + (it)) * 8 + (slot)] = clock64(); \
  } while (0)
#define ATT_STAMP(role, it, slot) ATT_STAMP_T(role, it, slot, 128)
static inline int att_trace_select_tu(int bx, int by, int bz, int mode) {
  int h[4] = {bx, by, bz, mode};
  cudaMemcpyToSymbol(g_att_trace_cta, h, sizeof(h));
  static long long zero[2 * 32 * 8] = {0};
  return (int)cudaMemcpyToSymbol(g_att_trace, zero, sizeof(zero));
}
static inline int att_trace_read_tu(long long* host) { return (int)cudaMemcpyFromSymbol(host, g_att_trace, sizeof(long long) * 2 * 32 * 8); }
#else
#define ATT_TRACE_INIT() do { } while (0)
#define ATT_STAMP_T(role, it, slot, row_thread) do { } while (0)
#define ATT_STAMP(role, it, slot) do { } while (0)
#endif
