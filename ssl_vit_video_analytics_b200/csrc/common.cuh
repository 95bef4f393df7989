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

// Build a tiled TMA descriptor over a bf16 tensor. dims/strides innermost-first; strides in BYTES for dims 1..rank-1.
// swizzle_bytes: 0, 64 or 128; the inner box extent in bytes must not exceed it.
int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes);
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

}  // namespace vmae
