// Host-side plumbing shared by all translation units of libvmae_sm100a.so:
// error reporting (vmae_last_error), TMA tensor-map construction, device properties.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include "../../include/vmae.h"

namespace vmae {

// error codes returned by every extern "C" entry point: VMAE_ERR_* from include/vmae.h
#define VMAE_OK 0

std::string& last_error_ref();  // thread-local, defined in api.cu

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

#define VMAE_CHECK_ARG(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) return ::vmae::fail(VMAE_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define VMAE_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return ::vmae::fail(VMAE_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
                          cudaGetErrorString(e__));                                                  \
  } while (0)

#define VMAE_LAUNCH_CHECK()                                                                          \
  do {                                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                            \
    if (e__ != cudaSuccess)                                                                          \
      return ::vmae::fail(VMAE_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,          \
                          cudaGetErrorString(e__));                                                  \
  } while (0)

int num_sms();  // cached multiprocessor count of the current device

// Launch with programmatic dependent launch enabled (the kernel must call pdl_wait() before its first global-memory access).
// Kernel groups: 1 = GEMM, 2 = LayerNorm, 4 = attention forward / two-kernel backward / delta, 8 = single-pass attention backward.
// VMAE_PDL=<mask> in the environment selects the groups that get the attribute (default: PDL_DEFAULT_MASK, the measured optimum);
// the others are launched with plain stream serialisation.
constexpr int PDL_GEMM = 1, PDL_LN = 2, PDL_ATTN = 4, PDL_ATTN_FUSED = 8;
constexpr int PDL_DEFAULT_MASK = PDL_GEMM | PDL_ATTN_FUSED;   // measured (ms/step): none 19.0, GEMM 18.85, +fused attention bwd 18.70, +LayerNorm 19.1 (worse: multi-wave grids gain nothing from an early start)
bool pdl_enabled(int group);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int group, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(group) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same, launched as thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(int group, int cluster_x, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(group) ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster_x;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Build a tiled TMA descriptor over a bf16 tensor. dims/strides innermost-first; strides in BYTES for dims 1..rank-1.
// swizzle_bytes: 0, 64 or 128; the inner box extent in bytes must not exceed it.
int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes);
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

}  // namespace vmae
