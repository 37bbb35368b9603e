// Micro-benchmark 2: does the ~1 gather/clk/SM limit count thread requests, distinct sectors, or bytes?
//  A: 8-byte gathers, every lane its own sector            (32 sectors / instruction)
//  B: 16-byte gathers, every lane its own sector           (32 sectors / instruction, 2x bytes)
//  C: 8-byte gathers, lane pairs share a 16-byte pair      (16 sectors / instruction)
//  D: 8-byte gathers, 4 lanes share a 32-byte sector       (8 sectors / instruction)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void k(const float2* __restrict__ tab, uint32_t mask, int iters, float* out) {
  const int lane = threadIdx.x & 31;
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  if (MODE == 2) s = (blockIdx.x * blockDim.x + (threadIdx.x & ~1)) * 2654435761u + 12345u;
  if (MODE == 3) s = (blockIdx.x * blockDim.x + (threadIdx.x & ~3)) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = s * 1664525u + 1013904223u;
      uint32_t idx = (s >> 8) & mask;
      if (MODE == 1) { v[j] = __ldg((const float4*)(tab + (idx & ~1u))); }
      else {
        if (MODE == 2) idx = (idx & ~1u) | (lane & 1);
        if (MODE == 3) idx = (idx & ~3u) | (lane & 3);
        const float2 t = __ldg(tab + idx); v[j] = make_float4(t.x, t.y, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, const float2* tab, uint32_t mask, float* out) {
  const int grid = 148, threads = 512, iters = 256;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  k<MODE><<<grid, threads, 128 * 1024>>>(tab, mask, iters, out); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<grid, threads, 128 * 1024>>>(tab, mask, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double loads = (double)grid * threads * iters * 8;
  printf("%s: %.3f ms  %.1f G thread-requests/s  %.2f per clk per SM\n", name, ms, loads / (ms * 1e-3) / 1e9, loads / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
  const uint32_t entries = 1u << 22;
  float2* tab; cudaMalloc(&tab, (size_t)entries * 8); cudaMemset(tab, 0, (size_t)entries * 8);
  float* out; cudaMalloc(&out, 148 * 2048 * 4);
  run<0>("A  8 B, 32 sectors/instr", tab, entries - 1, out);
  run<1>("B 16 B, 32 sectors/instr", tab, entries - 1, out);
  run<2>("C  8 B, 16 sectors/instr", tab, entries - 1, out);
  run<3>("D  8 B,  8 sectors/instr", tab, entries - 1, out);
  return 0;
}
