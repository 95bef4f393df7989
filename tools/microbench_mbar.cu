// Micro-benchmark (run on the GPU box): what does a SATISFIED mbarrier wait cost a warp? The attention kernels' row warps execute 2-5
// of them per 128 x 128 block. Measures, per call and per warp (clock64 around 1000 back-to-back calls on a barrier whose phase has
// already completed): mbarrier.try_wait.parity, mbarrier.test_wait.parity, a dependent ld.shared for scale, and
// tcgen05.fence::after_thread_sync.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_mbar tools/microbench_mbar.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void bench(long long* out, int iters) {
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t chain[32];
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) chain[threadIdx.x] = smem_u32(&chain[(threadIdx.x + 1) & 31]);
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");   // phase 0 completes
  __syncthreads();
  uint32_t acc = 0, addr = smem_u32(&chain[threadIdx.x & 31]);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      uint32_t done;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0u) : "memory");
      acc += done;
      if (!done) break;
    } else if (MODE == 1) {
      uint32_t done;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0u) : "memory");
      acc += done;
      if (!done) break;
    } else if (MODE == 2) {
      asm volatile("ld.shared.u32 %0, [%0];" : "+r"(addr)::"memory");
    } else {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      acc += i;
    }
  }
  const long long t1 = clock64();
  if (acc == 0xdeadbeef || addr == 1) out[1] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const char* names[] = {"mbarrier.try_wait.parity (satisfied)", "mbarrier.test_wait.parity (satisfied)", "dependent ld.shared", "tcgen05.fence::after_thread_sync"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int threads : {32, 128, 256}) {
      const int iters = 1000;
      if (mode == 0) bench<0><<<1, threads>>>(d, iters);
      if (mode == 1) bench<1><<<1, threads>>>(d, iters);
      if (mode == 2) bench<2><<<1, threads>>>(d, iters);
      if (mode == 3) bench<3><<<1, threads>>>(d, iters);
      cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%-42s threads=%3d: %7.1f cycles per call\n", names[mode], threads, (double)h / iters);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
