// platform.h — lets the per-thread kernel bodies compile both as sm_100a device code (the product) and
// as plain host C++ (tests/host_emu: a debug harness that runs the same bodies serially so kernel LOGIC can
// be unit-tested in a GPU-less container; it is never part of libanerend_cuda_ddgi.so).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#endif

namespace ddgi {

// Single-rounded fp32 operations that no compiler flag may fuse or approximate.  The radiance kernels (ddgi_shade.cu)
// are built with contraction and approximate div/sqrt; where a result is ill-conditioned (atlas texel addresses: one
// ulp of a coordinate near 2880 is 2e-4 texels; the GGX denominator at low roughness) the oracle's exact sequence is
// kept with these.  Everywhere else in device code compiled with --fmad=false they equal the plain operators.
#ifdef __CUDA_ARCH__
HD float xmul(float a, float b) { return __fmul_rn(a, b); }
HD float xadd(float a, float b) { return __fadd_rn(a, b); }
HD float xsub(float a, float b) { return __fsub_rn(a, b); }
HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
HD float xsqrt(float a) { return __fsqrt_rn(a); }
#else
HD float xmul(float a, float b) { return a * b; }
HD float xadd(float a, float b) { return a + b; }
HD float xsub(float a, float b) { return a - b; }
HD float xdiv(float a, float b) { return a / b; }
HD float xsqrt(float a) { return sqrtf(a); }
#endif

HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
HD int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
// index of the highest set bit (x != 0)
HD int msb32(uint32_t x) { return 31 - clz32(x); }

HD uint32_t atomAdd(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, v);
#else
  uint32_t o = *p; *p = o + v; return o;
#endif
}
HD void atomAdd64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
HD void atomMin(int* p, int v) {
#ifdef __CUDA_ARCH__
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}
HD void atomMin64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
  atomicMin(p, v);              // result unused: a fire-and-forget reduction, no round trip to wait for
#else
  if (v < *p) *p = v;
#endif
}
HD void atomMax64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}
HD void atomMax(int* p, int v) {
#ifdef __CUDA_ARCH__
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}
HD void memFence() {
#ifdef __CUDA_ARCH__
  __threadfence();
#endif
}

// rgba16f storage: round-to-nearest-even both ways (cvt.rn.f16.f32 on the device).
HD uint16_t f2h(float f) {
#ifdef __CUDA_ARCH__
  return __half_as_ushort(__float2half_rn(f));
#else
  uint32_t x = f2u(f), sign = (x >> 16) & 0x8000u, em = x & 0x7FFFFFFFu;
  if (em >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((em > 0x7F800000u) ? (0x0200u | ((em >> 13) & 0x3FFu)) : 0u));
  if (em >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);
  if (em < 0x33000001u) return (uint16_t)sign;
  int e = (int)(em >> 23) - 127;
  uint32_t m = (em & 0x7FFFFFu) | 0x800000u, base = 0; int shift = 13;
  if (e < -14) shift = 13 + (-14 - e); else { base = (uint32_t)(e + 15) << 10; m &= 0x7FFFFFu; }
  uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1), r = base + q;
  if (rem > half || (rem == half && (r & 1u))) r += 1;
  return (uint16_t)(sign | r);
#endif
}
HD float h2f(uint16_t h) {
#ifdef __CUDA_ARCH__
  return __half2float(__ushort_as_half(h));
#else
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { int s = 0; while (!(m & 0x400u)) { m <<= 1; ++s; } m &= 0x3FFu; x = sign | ((uint32_t)(113 - s) << 23) | (m << 13); }
  } else if (e == 31) x = sign | 0x7F800000u | (m << 13);
  else x = sign | ((e + 112) << 23) | (m << 13);
  return u2f(x);
#endif
}

}  // namespace ddgi
