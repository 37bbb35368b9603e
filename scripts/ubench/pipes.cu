// Micro-benchmark: issue rate of FFMA, FFMA2 (fma.rn.f32x2) and legacy mma.sync on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[8], aa, bb;
  float2 t = make_float2(a, a); aa = *(unsigned long long*)&t;
  t = make_float2(b, b); bb = *(unsigned long long*)&t;
  for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); x[i] = *(unsigned long long*)&v; }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  }
  float s = 0; for (int i = 0; i < 8; ++i) { float2 v = *(float2*)&x[i]; s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_f16(float* out) {
  unsigned a[4] = {threadIdx.x, 1, 2, 3}, b[2] = {5, threadIdx.x};
  float c[4][4] = {};
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_tf32(float* out) {
  unsigned a[4] = {threadIdx.x, 1, 2, 3}, b[2] = {5, threadIdx.x};
  float c[4][4] = {};
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int grid = 148, block = 512;   // 16 warps / SM = 4 / SMSP
  const double warps = (double)grid * block / 32;
  float ms;
  ms = time_ms([&] { k_ffma<<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("FFMA   : %.3f ms  %.2f warp-inst/clk/SM @1.9GHz  (%.1f TFMA/s)\n", ms, warps * ITER * 8 / (ms * 1e-3) / 148 / 1.9e9, warps * ITER * 8 * 32 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_ffma2<<<grid, block>>>(out, 1.0001f, 0.5f); });
  printf("FFMA2  : %.3f ms  %.2f warp-inst/clk/SM  (%.1f TFMA/s)\n", ms, warps * ITER * 8 / (ms * 1e-3) / 148 / 1.9e9, warps * ITER * 8 * 64 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_mma_f16<<<grid, block>>>(out); });
  printf("HMMA.16816 f16: %.3f ms  %.3f warp-inst/clk/SM  (%.1f TFMA/s)\n", ms, warps * ITER * 4 / (ms * 1e-3) / 148 / 1.9e9, warps * ITER * 4 * 2048 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_mma_tf32<<<grid, block>>>(out); });
  printf("HMMA.1688 tf32: %.3f ms  %.3f warp-inst/clk/SM  (%.1f TFMA/s)\n", ms, warps * ITER * 4 / (ms * 1e-3) / 148 / 1.9e9, warps * ITER * 4 * 1024 / (ms * 1e-3) / 1e12);
  return 0;
}
