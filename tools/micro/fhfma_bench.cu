// Throughput of FHFMA (fma.rn.f32.f16: f16*f16+f32) against FFMA and HADD2.F32+FFMA on sm_100a.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float fhfma(uint32_t a, uint32_t b, float c) {
  float r; uint16_t al, ah, bl, bh;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(al), "=h"(ah) : "r"(a));
  asm("mov.b32 {%0, %1}, %2;" : "=h"(bl), "=h"(bh) : "r"(b));
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(ah), "h"(bl), "f"(c)); return r;
}
template <int MODE> __global__ void k(float* out, uint32_t qa, uint32_t qb, float s, int iters) {
  float acc[8]; for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) acc[i] = fmaf(acc[i], s, 1.0f);
      else if (MODE == 1) acc[i] = fhfma(qa + i, qb, acc[i]);
      else { uint32_t d = qa + i + it; float2 f; asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(f.x), "=f"(f.y) : "r"(d)); acc[i] = fmaf(f.x, s, acc[i]) + f.y; }
    }
  }
  float r = 0; for (int i = 0; i < 8; ++i) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, 0x64006400u, 0x3C003C00u, 0.999f, iters);
      if (mode == 1) k<1><<<148 * 8, 256>>>(d, 0x64006400u, 0x3C003C00u, 0.999f, iters);
      if (mode == 2) k<2><<<148 * 8, 256>>>(d, 0x64006400u, 0x3C003C00u, 0.999f, iters);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = 148.0 * 8 * 256 * 8.0 * iters;
    printf("mode %d: %.3f ms, %.1f Gop/s (thread-level)\n", mode, ms, ops / ms / 1e6);
  }
  return 0;
}
