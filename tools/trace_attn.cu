// Developer timeline of one attention CTA (run on the GPU box). Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DVMAE_ATT_TRACE -Iinclude \
//        tools/trace_attn.cu ssl_vit_video_analytics_b200/csrc/attention.cu ssl_vit_video_analytics_b200/csrc/api.cu -o tools/bin/trace_attn
// Usage: trace_attn [fwd|dq|dkdv] [B] [L]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../include/vmae.h"

extern "C" int vmae_att_trace_select(int bx, int by, int bz, int mode);
extern "C" int vmae_att_trace_read(long long* host);
extern "C" int vmae_att_trace_select_fused(int bx, int by, int bz);
extern "C" int vmae_att_trace_read_fused(long long* host);

__global__ void fill(__nv_bfloat16* p, long long n, unsigned seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = __float2bfloat16(((x & 0xffff) / 65536.0f - 0.5f) * 2.0f);
  }
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "fwd";
  int B = argc > 2 ? atoi(argv[2]) : 16, L = argc > 3 ? atoi(argv[3]) : 1568, H = 6, hd = 64;
  long long nq = (long long)B * L * 3 * H * hd, no = (long long)B * L * H * hd;
  __nv_bfloat16 *qkv, *out, *dout, *dqkv;
  float *lse, *delta, *dqacc;
  cudaMalloc(&qkv, nq * 2); cudaMalloc(&out, no * 2); cudaMalloc(&dout, no * 2); cudaMalloc(&dqkv, nq * 2);
  cudaMalloc(&lse, (long long)B * H * L * 4); cudaMalloc(&delta, (long long)B * H * L * 4); cudaMalloc(&dqacc, no * 4);
  fill<<<1024, 256>>>(qkv, nq, 1u);
  fill<<<1024, 256>>>(dout, no, 7u);
  float scale = 0.125f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  vmae_att_trace_select(5, 3, B / 2, -1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    int rc = vmae_attn_fwd(qkv, out, lse, B, L, H, hd, scale, 0.f, 0ull, nullptr);
    if (rc) { printf("fwd failed: %s\n", vmae_last_error()); return 1; }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) printf("fwd B=%d L=%d: %.1f us\n", B, L, ms * 1e3);
  }
  if (strcmp(mode, "fwd") != 0) {
    const bool fused = strcmp(mode, "fused") == 0;
    vmae_att_trace_select(5, 3, B / 2, strcmp(mode, "dq") == 0 ? 0 : 1);
    if (fused) vmae_att_trace_select_fused(70, 12, 0);   // persistent kernel: CTA index, work-item number of that CTA
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      int rc = vmae_attn_bwd(qkv, out, dout, lse, delta, fused ? dqacc : nullptr, dqkv, B, L, H, hd, scale, 0.f, 0ull, nullptr);
      if (rc) { printf("bwd failed: %s\n", vmae_last_error()); return 1; }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("bwd B=%d L=%d: %.1f us\n", B, L, ms * 1e3);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> t(2 * 32 * 8);
  if (strcmp(mode, "fused") == 0) vmae_att_trace_read_fused(t.data()); else vmae_att_trace_read(t.data());
  long long base = 0;
  for (long long v : t) if (v > 0 && (base == 0 || v < base)) base = v;
  for (int role = 0; role < 2; ++role) {
    printf("role %d (%s): per iteration, stamps relative to first stamp\n", role, role == 0 ? "row warp" : "MMA thread");
    for (int it = 0; it < 32; ++it) {
      bool any = false;
      for (int s = 0; s < 8; ++s) any |= t[(role * 32 + it) * 8 + s] > 0;
      if (!any) continue;
      printf("  it %2d:", it);
      for (int s = 0; s < 8; ++s) {
        long long v = t[(role * 32 + it) * 8 + s];
        if (v > 0) printf(" %7lld", v - base); else printf("       -");
      }
      printf("\n");
    }
  }
  return 0;
}
