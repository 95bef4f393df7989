// libvmae_sm100a.so — library-wide host plumbing (see include/vmae.h for the C ABI).
#include <cstdlib>
#include "common.cuh"
#include "../../include/vmae.h"

namespace vmae {

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  // resolved through the runtime so the library carries no link-time dependency on libcuda.so
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes) {
  return make_tmap(out, base, VMAE_DTYPE_BF16, rank, dims, strides_bytes, box, swizzle_bytes);
}

int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(VMAE_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return fail(VMAE_ERR_BAD_ARG, "TMA base pointer %p not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) return fail(VMAE_ERR_BAD_ARG, "TMA stride %llu (dim %d) not a multiple of 16 bytes", (unsigned long long)gstr[i], i + 1);
  CUresult r = fn(out, dtype == VMAE_DTYPE_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (dtype == VMAE_DTYPE_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VMAE_ERR_DRIVER, "cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
  return VMAE_OK;
}

bool pdl_enabled(int group) {
  static const int mask = []() {
    const char* e = getenv("VMAE_PDL");
    return e ? atoi(e) : PDL_DEFAULT_MASK;
  }();
  return (mask & group) != 0;
}

}  // namespace vmae

extern "C" {

const char* vmae_last_error(void) { return vmae::last_error_ref().c_str(); }

int vmae_abi_version(void) { return VMAE_ABI_VERSION; }

int vmae_device_sm_count(void) { return vmae::num_sms(); }

}  // extern "C"
