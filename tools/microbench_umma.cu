// Micro-benchmark (run on the GPU box): what paces tcgen05.mma on one SM?
//   cycles per UMMA (M = 128, K = 16, bf16) for N in {64, 128, 192, 256}, A from shared memory (SS) or from TMEM (TS),
//   alone and with W other warps streaming st.shared.v4 / ld.shared.v4 into a disjoint region of shared memory.
// Hypothesis under test: one SM's shared memory moves ~128 B/clk for ALL clients together (UMMA operand reads, TMA fills,
// st.shared), so an SS UMMA with N <= 128 (>= 128 B of operands per clock at full rate) cannot run at the tensor peak.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_umma tools/microbench_umma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ssl_vit_video_analytics_b200/csrc/ptx.cuh"

using namespace vmae;

constexpr int OPER_BYTES = 64 * 1024;     // operand region: A tile 128 x 64 (16 KB) + B tile up to 256 x 64 (32 KB)
constexpr int SCRATCH_BYTES = 64 * 1024;  // region the traffic warps hammer

// traffic: 0 = none, 1 = st.shared.v4, 2 = ld.shared.v4
template <int N, bool TS>
__global__ void __launch_bounds__(288, 1) umma_bench(long long* out, int n_mma, int traffic, int traffic_warps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar_storage;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = smem_u32(&bar_storage);
  for (int i = threadIdx.x; i < (OPER_BYTES + SCRATCH_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); stop_flag = 0; }
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  long long t0 = 0, t1 = 0;
  long long traffic_ops = 0;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      const uint64_t dA = umma_smem_desc(sbase, 0, 1024, 128);
      const uint64_t dB = umma_smem_desc(sbase + 16 * 1024, 0, 1024, 128);
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // the 4 K = 16 slices of a 64-wide SW128 K panel, like a real main loop
          if (TS) umma_bf16_ts(tmem, tmem + 256 + 8 * k, dB + 2 * k, idesc, 1);
          else umma_bf16(tmem, dA + 2 * k, dB + 2 * k, idesc, 1);
        }
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      t1 = clock64();
      stop_flag = 1;
      out[blockIdx.x * 4 + 0] = t1 - t0;
    }
  } else if (warp - 1 < traffic_warps && traffic != 0) {
    const uint32_t my = sbase + OPER_BYTES + ((warp - 1) * 32 + lane) * 16;   // conflict-free: consecutive lanes, consecutive 16 B
    uint32_t a = lane, b = warp, c = 3, d = 4;
    while (!stop_flag) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const uint32_t addr = my + (r & 7) * (8 * 32 * 16);
        if (traffic == 1) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
        else asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
      }
      traffic_ops += 16;
    }
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(out + blockIdx.x * 4 + 1), static_cast<unsigned long long>(traffic_ops));
    if (a == 0xdeadbeef) out[3] = a + b + c + d;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS>
void run(const char* name, int traffic, int tw) {
  long long* d;
  cudaMalloc(&d, 4 * sizeof(long long));
  cudaMemset(d, 0, 4 * sizeof(long long));
  const int n_mma = 4096;
  cudaFuncSetAttribute(umma_bench<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, OPER_BYTES + SCRATCH_BYTES);
  umma_bench<N, TS><<<1, 288, OPER_BYTES + SCRATCH_BYTES>>>(d, n_mma, traffic, tw);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double cyc = (double)h[0] / n_mma;
  const double ideal = N / 2.0;   // 128 x N x 16 MACs at 4096 MAC/clk
  const double oper = (TS ? 0 : 128 * 16 * 2) + N * 16 * 2;
  const double traffic_bytes = (double)h[1] * 32 * 16;   // per warp-instruction: 32 lanes x 16 B
  printf("%-4s N=%3d traffic=%s x%d: %7.1f cyc/UMMA (ideal %5.1f, %4.0f%% of peak)  operand B/clk %6.1f  traffic B/clk %6.1f  total %6.1f  [%s]\n", name, N,
         traffic == 0 ? "none" : (traffic == 1 ? "sts " : "lds "), tw, cyc, ideal, 100.0 * ideal / cyc, oper / cyc, traffic_bytes / (double)h[0],
         oper / cyc + traffic_bytes / (double)h[0], cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int traffic = 0; traffic <= 2; ++traffic) {
    for (int tw : {4, 8}) {
      if (traffic == 0 && tw != 4) continue;
      run<64, false>("SS", traffic, tw);
      run<128, false>("SS", traffic, tw);
      run<192, false>("SS", traffic, tw);
      run<256, false>("SS", traffic, tw);
      run<64, true>("TS", traffic, tw);
      run<128, true>("TS", traffic, tw);
      run<256, true>("TS", traffic, tw);
    }
  }
  return 0;
}
