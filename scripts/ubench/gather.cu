// Micro-benchmark: rate of random 8-byte gathers from an L2-resident 48 MB table on B200, as a function
// of resident warps per SM and of independent loads in flight per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int ILP>
__global__ void k_gather(const float2* __restrict__ tab, uint32_t mask, int iters, float* out) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float2 v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      s = s * 1664525u + 1013904223u;
      v[j] = __ldg(tab + ((s >> 8) & mask));
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j].x + v[j].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ILP>
void run(const float2* tab, uint32_t mask, float* out, int threads, int ctas_per_sm, size_t smem) {
  const int grid = 148 * ctas_per_sm, iters = 2048 / ILP;
  cudaFuncSetAttribute(k_gather<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_gather<ILP><<<grid, threads, smem>>>(tab, mask, iters, out);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_gather<ILP><<<grid, threads, smem>>>(tab, mask, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double loads = (double)grid * threads * iters * ILP;
  printf("threads/SM %4d ILP %2d smem %3zu KB: %.3f ms  %.1f G gathers/s  %.2f per clk per SM (1.9 GHz)\n",
         threads * ctas_per_sm, ILP, smem / 1024, ms, loads / (ms * 1e-3) / 1e9, loads / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
  const uint32_t entries = 1u << 22;          // 4 M x 8 B = 32 MB (L2 resident)
  float2* tab; cudaMalloc(&tab, (size_t)entries * 8); cudaMemset(tab, 0, (size_t)entries * 8);
  float* out; cudaMalloc(&out, 148 * 2048 * 4);
  for (size_t smem : {(size_t)0, (size_t)128 * 1024, (size_t)200 * 1024}) {
    run<4>(tab, entries - 1, out, 512, 1, smem);
    run<8>(tab, entries - 1, out, 512, 1, smem);
    run<16>(tab, entries - 1, out, 512, 1, smem);
    run<32>(tab, entries - 1, out, 512, 1, smem);
    run<8>(tab, entries - 1, out, 1024, 1, smem);
    run<16>(tab, entries - 1, out, 1024, 1, smem);
    if (smem == 0) { run<8>(tab, entries - 1, out, 1024, 2, smem); run<16>(tab, entries - 1, out, 1024, 2, smem); }
  }
  return 0;
}
