// FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/build/ffma_bench tools/ffma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[16], b = threadIdx.x * 1e-3f + 1.0f, c = 0.5f;
  unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = i + threadIdx.x;
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
  unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b), cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters); else k<1><<<148 * 8, 256>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)148 * 8 * 256 * iters * 16;
      printf("%s: %.3f ms  %.1f TFLOP/s (fp32 FMA = 2 flop)\n", mode == 0 ? "FFMA " : "FFMA2", ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
