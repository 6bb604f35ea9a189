// ddgi_shade.cu — radiance and convolution kernels of the wavefront path (sm_100a).  Compiled TWICE
// (__graft_entry__.py NVCC_TU_FLAGS):
//   -DDDGI_SHADE_EXACT  --fmad=false                                   -> launch_shade_exact / launch_convolve_exact
//        the oracle's fp32 sequence operation for operation: ProbeRaysIrrTex and the atlas come out bit-identical
//   (fast)              --fmad=true -prec-div=false -prec-sqrt=false   -> launch_shade_fast / launch_convolve_fast
//        ANEREND_DDGI_FLAG_FAST_MATH: contraction and approximate div/sqrt (<= 2 ulp), ill-conditioned spots kept exact
//        with the x* operations of platform.h; within BASELINE.json's 1e-4 relative bar, not bit-identical
//
//   k_shade                 rchit:83-109 (calcLight, calcIndirectDiffuseLight, emissive) + the RGBA16F store of rgen:109,
//                           one thread per HIT ray through the compacted hit list, so warps stay full
//   k_conv_weights          per layer: W[ray][texel] = max(0, dot(half(dir), octDecode(texel)))   (probe_conv.comp:43-58)
//   k_probe_convolve_fast   per probe: out[texel] = sum_ray W[ray][texel] * (rad.rgb, 1), normalise, hysteresis blend,
//                           RGBA16F store (probe_conv.comp:49-89); radiances staged once per block in shared memory as fp32
#ifndef DDGI_HOST_EMU
#include "ddgi_wavefront.h"

#ifdef DDGI_SHADE_EXACT
#define SHADE_NS exact
#else
#define SHADE_NS fast
#endif

#define SHADE_MINB_DEFAULT 8         // 64 registers: occupancy beats the few spills (gpu sweep: 4/5/6/7/8 -> 218/192/185/179/175 us at C3)
#include <cstdlib>

namespace ddgi {
namespace SHADE_NS {

// ------------------------------------------------------------------------------------------------------------
template <int MINB, bool EXT = false, bool TILES = false>
__global__ void __launch_bounds__(128, MINB) k_shade(const FastCtx c) {
  const TraceCtx& t = c.t;
  const uint32_t nHits = c.qcount[FAST_QC_HITS];
  const int P = t.nXZ * t.nXZ;
  const V3 sunL = normalize3x(-t.lightDir);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nHits; i += gridDim.x * blockDim.x) {
    const uint32_t idx = c.hitqSorted ? c.hitqSorted[i] : c.hitq[i];
    int layer = 0, probe = 0; uint32_t ray = 0; V3 origin;
    if (!EXT) {
      const uint32_t slot = idx % (uint32_t)t.N, probeLin = idx / (uint32_t)t.N;
      const int layerLocal = (int)(probeLin / (uint32_t)t.Pl);
      probe = stripe_probe((int)(probeLin % (uint32_t)t.Pl), t.nXZ, t.rowOffset, t.rowStride); layer = t.layerMap[layerLocal];
      ray = t.rayOrder ? (uint32_t)t.rayOrder[slot] : slot;
      origin = mk3((float)(probe % t.nXZ), (float)layer, (float)(probe / t.nXZ)) * t.step + t.camOff;
    } else { const float4 ea = c.extRays[2 * (size_t)idx]; origin = mk3(ea.x, ea.y, ea.z); }      // gl_WorldRayOriginEXT
    const float4 s0 = c.surf[3 * (size_t)idx], s1 = c.surf[3 * (size_t)idx + 1], s2 = c.surf[3 * (size_t)idx + 2];
    const V3 pos = mk3(s0.x, s0.y, s0.z), normal = mk3(s0.w, s1.x, s1.y), albedo = mk3(s1.z, s1.w, s2.x);
    const anerend_material_info_t* m = &t.materials[__float_as_uint(s2.y)];
    const float metallic = s2.z, roughness = s2.w;                 // material factors or texture taps (k_shade_prep)
    V3 emissive = mk3(m->emissive[0], m->emissive[1], m->emissive[2]) * m->emissive[3];
    if (c.emisTex && m->bindless_indices[3] != -1) { const float4 e = c.emisTex[idx]; emissive = mk3(e.x, e.y, e.z); }
    const uint32_t lit = c.lit[idx];
    V3 irradiance = mk3(0.0f);
    if (lit & 1u) irradiance = calc_light(normal, albedo, metallic, roughness, pos, origin, sunL, t.sunIntensity * mk3(1.0f) * 1.0f);
    if (t.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) {
      // every lane walks ITS OWN unoccluded lights in index order (the oracle's accumulation order); looping over all
      // lights with a per-lane skip kept a third of the lanes busy
      for (uint32_t m = lit >> 1; m != 0u; m &= m - 1u) {
        const uint32_t li = (uint32_t)__ffs((int)m) - 1u;
        const anerend_light_t* lt = &t.lights[li];
        const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
        const V3 dv = lp - pos; const float dist = sqrtf(dot3(dv, dv)), range = lt->world_pos[3];
        const V3 L = normalize3x(lp - pos);
        const float dd = fminf(dist, range);
        const float att = (1.0f - dd / range) * (1.0f - dd / range) * 1.0f;
        irradiance = irradiance + calc_light(normal, albedo, metallic, roughness, pos, origin, L, mk3(lt->color[0], lt->color[1], lt->color[2]) * att);
      }
    }
    const uint32_t gate = EXT ? ANEREND_UBO_DDGI_FLAG : ANEREND_UBO_DDGI_MULTI_FLAG;      // mirror_reflections.rchit:105 / rchit:98
    if ((t.uboFlags & gate) == gate)
      irradiance = irradiance + calc_indirect<TILES>(t, normal, albedo, metallic, roughness, pos, origin);      // live atlas (probeTex), tile placement if TILES
    const V3 rad = irradiance + emissive;
    const size_t o = EXT ? (size_t)idx : ((size_t)layer * P + probe) * t.N + ray;
    const uint32_t alpha = EXT ? (uint32_t)f2h(c.extDist[idx]) : 0x3C00u;                 // SpecularGIRT stores the hit distance (rgen:79)
    t.rayIrr[o] = make_uint2((uint32_t)f2h(rad.x) | ((uint32_t)f2h(rad.y) << 16), (uint32_t)f2h(rad.z) | (alpha << 16));
    if (t.records) {
      const float4 h = c.hits[idx];
      anerend_ddgi_ray_record_t rr;
      rr.prim = __float_as_uint(h.w); rr.t = h.x; rr.u = h.y; rr.v = h.z; rr.radiance[0] = rad.x; rr.radiance[1] = rad.y; rr.radiance[2] = rad.z; rr.lit_mask = lit;
      t.records[o] = rr;
    }
  }
}

// the less common instantiations (SpecularGIRT's external rays, scenes with baked tile atlases): one resident wave of blocks
template <bool EXT, bool TILES>
static cudaError_t launch_shade_variant(const FastCtx& f, cudaStream_t s) {
  static int maxBlocks = 0;
  if (!maxBlocks) {
    int dev = 0, sms = 0, nb = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_shade<8, EXT, TILES>, 128, 0);
    maxBlocks = sms * (nb > 0 ? nb : 1);
  }
  uint32_t blocks = (f.totalRays + 127u) / 128u;
  if (blocks > (uint32_t)maxBlocks) blocks = (uint32_t)maxBlocks;
  if (blocks == 0) return cudaSuccess;
  k_shade<8, EXT, TILES><<<blocks, 128, 0, s>>>(f);
  return cudaGetLastError();
}
static cudaError_t launch_shade_ext_impl(const FastCtx& f, cudaStream_t s) {
  return f.t.tex.tileInfos ? launch_shade_variant<true, true>(f, s) : launch_shade_variant<true, false>(f, s);
}
template <int MINB>
static cudaError_t launch_shade_minb(const FastCtx& f, cudaStream_t s) {
  static int maxBlocks = 0;
  if (!maxBlocks) {
    int dev = 0, sms = 0, nb = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_shade<MINB>, 128, 0);
    maxBlocks = sms * (nb > 0 ? nb : 1);
  }
  // the hit count lives on the device: size the grid for the worst case (every ray hits), capped at one resident wave
  uint32_t blocks = (f.totalRays + 127u) / 128u;
  if (blocks > (uint32_t)maxBlocks) blocks = (uint32_t)maxBlocks;
  if (blocks == 0) return cudaSuccess;
  k_shade<MINB><<<blocks, 128, 0, s>>>(f);
  return cudaGetLastError();
}
static cudaError_t launch_shade_impl(const FastCtx& f, cudaStream_t s) {
  if (f.t.tex.tileInfos) return launch_shade_variant<false, true>(f, s);
  static int minb = 0;            // register budget of k_shade: resident blocks per SM the compiler must allow
  if (!minb) { const char* e = getenv("DDGI_SHADE_MINB"); minb = e ? atoi(e) : SHADE_MINB_DEFAULT; }   // sweeps only
  switch (minb) {
    case 4: return launch_shade_minb<4>(f, s);
    case 5: return launch_shade_minb<5>(f, s);
    case 7: return launch_shade_minb<7>(f, s);
    case 6: return launch_shade_minb<6>(f, s);
    case 10: return launch_shade_minb<10>(f, s);
    case 12: return launch_shade_minb<12>(f, s);
    default: return launch_shade_minb<8>(f, s);
  }
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_conv_weights(const ConvFastCtx f) {
  const ConvCtx& c = f.c;
  // one thread per (layer, ray, texel row j): 8 weights each, so even a single layer fills a few SMs
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, idx = tid >> 3, j = tid & 7u;
  if (idx >= (uint32_t)c.numLayers * (uint32_t)c.N) return;
  // position p in the reference's loop order (x outer, y inner; probe_conv.comp:49-50) <-> ray index y*Sx + x
  const uint32_t layerLocal = idx / (uint32_t)c.N, pos = idx % (uint32_t)c.N;
  const uint32_t ray = (pos % (uint32_t)c.Sy) * (uint32_t)c.Sx + pos / (uint32_t)c.Sy;
  const float4 d4 = c.dirs[(uint32_t)c.layerMap[layerLocal] * (uint32_t)c.N + ray];      // [nY][N]
  const V3 d = mk3(h2f(f2h(d4.x)), h2f(f2h(d4.y)), h2f(f2h(d4.z)));  // what the rgba16f ProbeRaysDirTex holds (rgen:108)
  float4* out = (float4*)(f.weights + (size_t)idx * 64);
  float w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float ux = (2.0f * (float)i + 1.0f) / (2.0f * 8.0f), uy = (2.0f * (float)j + 1.0f) / (2.0f * 8.0f);   // pixelToUv
    ux = ux * 2.0f - 1.0f; uy = uy * 2.0f - 1.0f;
    const V3 octDir = oct_decode(ux, uy);
    const float ww = fmaxf(0.0f, dot3(d, octDir));
    w[i] = ww > 1e-6f ? ww : 0.0f;                                // probe_conv.comp:55
  }
  out[2 * j] = make_float4(w[0], w[1], w[2], w[3]);
  out[2 * j + 1] = make_float4(w[4], w[5], w[6], w[7]);
}

#define CONV_CHUNK 128                // rays staged per pass through shared memory

// CONV_PPT probes per thread, GROUPS probe groups per block (64 * GROUPS threads = 64 texels x GROUPS), CONV_G = GROUPS * CONV_PPT probes per block
template <int CONV_PPT, int GROUPS = 4>
__global__ void __launch_bounds__(64 * GROUPS) k_probe_convolve_fast(const ConvFastCtx f) {
  constexpr int CONV_G = GROUPS * CONV_PPT, THREADS = 64 * GROUPS;
  const ConvCtx& c = f.c;
  __shared__ float4 sRad[CONV_G * CONV_CHUNK];                      // 32 KB
  const int P = c.nXZ * c.nXZ;
  const int blocksPerLayer = (c.Pl + CONV_G - 1) / CONV_G;
  const int layerLocal = (int)blockIdx.x / blocksPerLayer, g0 = ((int)blockIdx.x % blocksPerLayer) * CONV_G;
  const int layer = c.layerMap[layerLocal];
  const int texel = (int)threadIdx.x & 63, sub = (int)threadIdx.x >> 6;
  float acc[CONV_PPT][3], accW = 0.0f;          // sum of the weights: the same for every probe of the layer
#pragma unroll
  for (int k = 0; k < CONV_PPT; ++k) { acc[k][0] = 0.0f; acc[k][1] = 0.0f; acc[k][2] = 0.0f; }
  const float* W = f.weights + (size_t)layerLocal * c.N * 64 + texel;
  for (int r0 = 0; r0 < c.N; r0 += CONV_CHUNK) {
    const int nr = min(CONV_CHUNK, c.N - r0);
    int bad = 0;
    __syncthreads();
    for (int e = (int)threadIdx.x; e < CONV_G * CONV_CHUNK; e += THREADS) {
      const int g = e / CONV_CHUNK, r = e % CONV_CHUNK;
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (g0 + g < c.Pl && r < nr) {
        const int pos = r0 + r, ray = (pos % c.Sy) * c.Sx + pos / c.Sy;      // reference loop order
        const uint2 rr = __ldg(&c.rayIrr[((size_t)layer * P + stripe_probe(g0 + g, c.nXZ, c.rowOffset, c.rowStride)) * c.N + ray]);
        v = make_float4(h2f((uint16_t)(rr.x & 0xFFFFu)), h2f((uint16_t)(rr.x >> 16)), h2f((uint16_t)(rr.y & 0xFFFFu)), 1.0f);
      }
      sRad[e] = v;
      bad |= !(fabsf(v.x) <= 3.0e38f) || !(fabsf(v.y) <= 3.0e38f) || !(fabsf(v.z) <= 3.0e38f);
    }
    // probe_conv.comp:54-57 SKIPS rays with weight <= 1e-6; multiplying by the zeroed weight is the same value unless the
    // radiance is Inf/NaN (an fp16 overflow of a very bright hit): 0 * Inf = NaN where the shader adds nothing.  Blocks that
    // staged a non-finite radiance (block-uniform, rare) take the loop with the explicit skip.
    if (__syncthreads_or(bad)) {
      const float4* badRad = sRad + sub * CONV_PPT * CONV_CHUNK;
      for (int r = 0; r < nr; ++r) {
        const float w = __ldg(W + (size_t)(r0 + r) * 64);
        accW += w;
        if (w != 0.0f) {
#pragma unroll
          for (int k = 0; k < CONV_PPT; ++k) {
            const float4 v = badRad[k * CONV_CHUNK + r];
#ifdef DDGI_SHADE_EXACT
            acc[k][0] = acc[k][0] + w * v.x; acc[k][1] = acc[k][1] + w * v.y; acc[k][2] = acc[k][2] + w * v.z;
#else
            acc[k][0] = fmaf(w, v.x, acc[k][0]); acc[k][1] = fmaf(w, v.y, acc[k][1]); acc[k][2] = fmaf(w, v.z, acc[k][2]);
#endif
          }
        }
      }
      continue;
    }
    const float4* myRad = sRad + sub * CONV_PPT * CONV_CHUNK;
#pragma unroll 4
    for (int r = 0; r < nr; ++r) {
      const float w = __ldg(W + (size_t)(r0 + r) * 64);
      accW += w;
#pragma unroll
      for (int k = 0; k < CONV_PPT; ++k) {
        const float4 v = myRad[k * CONV_CHUNK + r];
#ifdef DDGI_SHADE_EXACT
        // the oracle's sequence: result += (w*rad, w) in loop order; w == 0 where the shader skips (adds +0)
        acc[k][0] = acc[k][0] + w * v.x; acc[k][1] = acc[k][1] + w * v.y; acc[k][2] = acc[k][2] + w * v.z;
#else
        acc[k][0] = fmaf(w, v.x, acc[k][0]); acc[k][1] = fmaf(w, v.y, acc[k][1]);
        acc[k][2] = fmaf(w, v.z, acc[k][2]);
#endif
      }
    }
  }
  const int i = texel & 7, j = texel >> 3;
  const float hy = c.hysteresis;
#pragma unroll
  for (int k = 0; k < CONV_PPT; ++k) {
    if (g0 + sub * CONV_PPT + k >= c.Pl) continue;
    const int probe = stripe_probe(g0 + sub * CONV_PPT + k, c.nXZ, c.rowOffset, c.rowStride);
    const int probeX = probe % c.nXZ, probeZ = probe / c.nXZ;
    const int ax = 10 * probeX + 1 + i, ay = 10 * c.nXZ * layer + 10 * probeZ + 1 + j;       // pixOffset + pixel (:42,:64)
    const size_t ti = (size_t)ay * c.W + ax;
    const uint2 ot = c.atlas[ti];
    const float o0 = h2f((uint16_t)(ot.x & 0xFFFFu)), o1 = h2f((uint16_t)(ot.x >> 16)), o2 = h2f((uint16_t)(ot.y & 0xFFFFu)), o3 = h2f((uint16_t)(ot.y >> 16));
    float r0 = acc[k][0], r1 = acc[k][1], r2 = acc[k][2];
    const float r3 = accW;
    float n0, n1, n2, n3;
    if (c.bakeMode) {                       // probe_conv.comp:67-79: running weighted mean, weight sum in alpha
      const float totalWeight = o3 + r3;
      float t0 = r0 + o0 * o3, t1 = r1 + o1 * o3, t2 = r2 + o2 * o3;
      if (totalWeight > 1e-6f) { t0 /= totalWeight; t1 /= totalWeight; t2 /= totalWeight; }
      n0 = t0; n1 = t1; n2 = t2; n3 = totalWeight;
    } else {
#ifdef DDGI_SHADE_EXACT
      if (r3 > 1e-6f) { r0 /= r3; r1 /= r3; r2 /= r3; }
#else
      if (r3 > 1e-6f) { const float inv = 1.0f / r3; r0 *= inv; r1 *= inv; r2 *= inv; }
#endif
      n0 = r0 * (1.0f - hy) + o0 * hy; n1 = r1 * (1.0f - hy) + o1 * hy; n2 = r2 * (1.0f - hy) + o2 * hy; n3 = r3 * (1.0f - hy) + o3 * hy;
    }
    const uint2 nt = make_uint2((uint32_t)f2h(n0) | ((uint32_t)f2h(n1) << 16), (uint32_t)f2h(n2) | ((uint32_t)f2h(n3) << 16));
    c.atlas[ti] = nt;
    for (int p = 0; p < c.numPeers; ++p) c.peers[p][ti] = nt;               // fused all-gather: store into every peer's atlas
    if (c.texelRecords) c.texelRecords[ti] = make_float4(n0, n1, n2, n3);
  }
}

static cudaError_t launch_convolve_impl(const ConvFastCtx& f, cudaStream_t s) {
  const ConvCtx& c = f.c;
  const uint32_t nw = (uint32_t)c.numLayers * (uint32_t)c.N;
  if (nw == 0) return cudaSuccess;
  k_conv_weights<<<(8u * nw + 127u) / 128u, 128, 0, s>>>(f);
  static int ppt = 0, groups = 0;  // probes per thread, probe groups per block: sweeps only (DDGI_CONV_PPT, DDGI_CONV_GROUPS)
  if (!ppt) { const char* e = getenv("DDGI_CONV_PPT"); ppt = e ? atoi(e) : 4; const char* g = getenv("DDGI_CONV_GROUPS"); groups = g ? atoi(g) : 1; }      // 64-thread blocks of 4 probes: 0.062 ms at C3 against 0.069 with 256 threads / 16 probes
  const int pp = ppt == 2 ? 2 : ppt == 1 ? 1 : ppt == 8 ? 8 : 4, gg = groups == 2 ? 2 : groups == 1 ? 1 : 4;
  const int G = gg * pp;
  const uint32_t blocks = (uint32_t)c.numLayers * (uint32_t)((c.Pl + G - 1) / G);
  if (gg == 4) {
    if (pp == 2) k_probe_convolve_fast<2, 4><<<blocks, 256, 0, s>>>(f);
    else if (pp == 1) k_probe_convolve_fast<1, 4><<<blocks, 256, 0, s>>>(f);
    else k_probe_convolve_fast<4, 4><<<blocks, 256, 0, s>>>(f);
  } else if (gg == 2) {
    if (pp == 2) k_probe_convolve_fast<2, 2><<<blocks, 128, 0, s>>>(f);
    else if (pp == 8) k_probe_convolve_fast<8, 2><<<blocks, 128, 0, s>>>(f);
    else k_probe_convolve_fast<4, 2><<<blocks, 128, 0, s>>>(f);
  } else {
    if (pp == 8) k_probe_convolve_fast<8, 1><<<blocks, 64, 0, s>>>(f);
    else k_probe_convolve_fast<4, 1><<<blocks, 64, 0, s>>>(f);
  }
  return cudaGetLastError();
}

}  // namespace SHADE_NS

#ifdef DDGI_SHADE_EXACT
cudaError_t launch_shade_exact(const FastCtx& f, cudaStream_t s) { return exact::launch_shade_impl(f, s); }
cudaError_t launch_shade_ext_exact(const FastCtx& f, cudaStream_t s) { return exact::launch_shade_ext_impl(f, s); }
cudaError_t launch_convolve_exact(const ConvFastCtx& f, cudaStream_t s) { return exact::launch_convolve_impl(f, s); }
#else
cudaError_t launch_shade_fast(const FastCtx& f, cudaStream_t s) { return fast::launch_shade_impl(f, s); }
cudaError_t launch_shade_ext_fast(const FastCtx& f, cudaStream_t s) { return fast::launch_shade_ext_impl(f, s); }
cudaError_t launch_convolve_fast(const ConvFastCtx& f, cudaStream_t s) { return fast::launch_convolve_impl(f, s); }
#endif

}  // namespace ddgi
#endif  // DDGI_HOST_EMU
