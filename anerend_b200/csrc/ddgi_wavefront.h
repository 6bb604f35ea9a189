// ddgi_wavefront.h — shared declarations of the production (wavefront) kernels for sm_100a.
//
// Translation units:
//   ddgi_api.cu   (--fmad=false)  traversal + hit attributes (ddgi_fast.cuh): everything that decides WHICH triangle a
//                                 ray hits and where shadow rays start is the oracle's fp32 sequence bit for bit
//   ddgi_shade.cu                 radiance (rchit:83-109, pbr_light.glsl) and the convolution (probe_conv.comp), built twice:
//                                 exact (--fmad=false, default: bit-identical atlas) and fast (ANEREND_DDGI_FLAG_FAST_MATH:
//                                 contraction + approximate div/sqrt, inside BASELINE.json's 1e-4 relative bar)
#pragma once
#ifndef __CUDACC__
#error "ddgi_wavefront.h is CUDA-only"
#endif
#include "ddgi_kernels.cuh"

namespace ddgi {

// per-triangle shading record, written once by k_flatten_scene in global-triangle-id order (object space, so a refit
// of moving nodes leaves it valid): replaces the 4-deep gather triInstance -> instance -> indices -> 3 x 96 B vertices
// of rchit:27-38 by five independent 16-byte loads
struct ShadeTri {
  float p0[3], p1[3], p2[3];      // object-space positions   (rchit:40-41)
  float n0[3], n1[3], n2[3];      // object-space normals     (rchit:72)
  uint32_t instance;              // index into instances / xforms
  uint32_t flags;                 // bit0: material factor decides the albedo (vertex colours not needed)
};
static_assert(sizeof(ShadeTri) == 80, "ShadeTri is five float4");

// qcount: one counter per 256 bytes (same-line atomics serialise in one L2 slice): [0..31] per-light shadow queue
// lengths, then the fetch cursors of the persistent kernels and the hit count
#define FAST_QC_STRIDE 64
#define FAST_QC(q) ((q) * FAST_QC_STRIDE)
#define FAST_QC_PRIMARY FAST_QC(32)
#define FAST_QC_SHADOW FAST_QC(33)
#define FAST_QC_HITS FAST_QC(34)
#define FAST_QC_BUCKET FAST_QC(35)      // 33 totals of the hit list's classes (number of unoccluded lights of a hit), then 33 scatter cursors
#define FAST_QC_SIZE (FAST_QC(35) + 128)

struct FastCtx {
  TraceCtx t;
  const float4* shadeTris;  // [numTris] ShadeTri
  const float4* shadeUV;    // [numTris][2] texture coordinates of the three vertices, or null (no textured material)
  float4* emisTex;          // [rays] emissive of hits on emissive-textured materials, or null
  float4* hits;             // [rays] t,u,v,prim   (launch order: idx = (layerLocal*P + probe)*N + slot)
  uint32_t* hitInst;        // [rays] instance of the hit: two-level structures only (no triangle -> instance table there)
  float4* surf;             // [rays][3] pos.xyz,n.x | n.yz,albedo.xy | albedo.z,material,metallic,roughness
  uint32_t* lit;            // [rays] visibility bits written by k_trace_shadow
  uint32_t* queues;         // [numQueues][rays] ray indices needing a shadow ray towards light q (0 = sun)
  uint32_t* hitq;           // [rays] compacted indices of the rays that hit (k_shade runs over these with full warps)
  uint32_t* hitqSorted;     // [rays] the same list grouped by the number of unoccluded lights (k_hit_bucket), or null
  uint32_t* qcount;         // FAST_QC_SIZE counters
  uint32_t* primaryCursor;  // fetch cursor of the closest-hit kernel when it runs on the pipeline stream (null: qcount[FAST_QC_PRIMARY])
  uint32_t totalRays, numQueues;
  // external-ray mode (SpecularGIRT: one ray per pixel instead of the probe rays; kernels instantiated with EXT = true)
  const float4* extRays;    // [rays][2] (origin.xyz, tMax; 0 = no ray) (dir.xyz, -)
  float* extDist;           // [rays] distance(hit position, ray origin) of the hits (mirror_reflections.rchit:86)
  uint32_t surfStride;      // float4s per surface record (3; ShadowRT stages positions only: 1)
  uint32_t rawSunDir;       // queue 0 rays run along -lightDir as given (ShadowRT) instead of normalize(-lightDir) (probe rchit)
  uint32_t tune[8];         // loop tunables (ddgi_fast.cuh TUNE_*); [3] = 0x64646464 (planes2)
};

// W[layer][ray][texel] = max(0, dot(half(dir), octDir(texel))), zero where <= 1e-6 (probe_conv.comp:53-58).  The
// rotation is shared by every probe of a layer execution, so the convolution is the contraction
// [probes*4 x N] . [N x 64] — done in fp32 on the CUDA cores (0.9 GFLOP at C3: not tensor-worthy).
struct ConvFastCtx {
  ConvCtx c;
  float* weights;           // [numLayers][N][64], N in the reference's loop order (x outer, y inner)
};

// launchers implemented in ddgi_shade.cu (one pair per build of that file)
cudaError_t launch_shade_exact(const FastCtx& f, cudaStream_t s);
cudaError_t launch_shade_ext_exact(const FastCtx& f, cudaStream_t s);      // EXT = true (SpecularGIRT)
cudaError_t launch_shade_ext_fast(const FastCtx& f, cudaStream_t s);
cudaError_t launch_convolve_exact(const ConvFastCtx& f, cudaStream_t s);
cudaError_t launch_shade_fast(const FastCtx& f, cudaStream_t s);
cudaError_t launch_convolve_fast(const ConvFastCtx& f, cudaStream_t s);

}  // namespace ddgi
