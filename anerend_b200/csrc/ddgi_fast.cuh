// ddgi_fast.cuh — the production (wavefront) kernels of "IrradianceProbeRT" for sm_100a.
//
// The reference runs rgen -> traceRayEXT -> rchit (-> traceRayEXT shadow) -> rmiss per ray inside the RT pipeline
// (irradiance_probe_update.rgen:48-111, .rchit:24-111, .rmiss:17-21).  A one-thread-per-ray transcription of that
// (k_probe_trace in ddgi_kernels.cuh) keeps only ~8 of 32 lanes busy (profiles/r01_*), so the pass is split into
// stages that each keep warps full:
//   k_trace_primary  persistent warps, closest-hit BVH8 traversal, rays re-fetched with __ballot when lanes drain
//   k_shade_prep     hit attributes (rchit:27-42,70-77) -> surface record; shadow rays binned into one queue per light
//   k_trace_shadow   persistent warps, any-hit traversal over the per-light queues (coherent direction / target)
//   k_shade          calcLight + calcIndirectDiffuseLight + emissive (rchit:83-109), RGBA16F store (rgen:109)
// Per-ray arithmetic is the same sequence of fp32 operations as the simple path and the oracle (the box test differs,
// but boxes are conservative and the triangle test decides), so results stay bit-identical.
#pragma once
#ifndef __CUDACC__
#error "ddgi_fast.cuh is CUDA-only"
#endif
#include "ddgi_wavefront.h"

namespace ddgi {

#define FAST_FETCH_THRESHOLD 22     // refill the warp when fewer lanes than this still have a ray

struct RayState {
  float ox, oy, oz, dx, dy, dz, idx, idy, idz, tmin, tmax, hu, hv;
  uint32_t octinv, prim;
  uint2 ng, tg;
  int sp;
};

__device__ __forceinline__ void ray_begin(RayState& r, V3 o, V3 d, float tmin, float tmax) {
  r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
  r.idx = 1.0f / safe_dir(d.x); r.idy = 1.0f / safe_dir(d.y); r.idz = 1.0f / safe_dir(d.z);
  r.tmin = tmin; r.tmax = tmax; r.hu = 0.0f; r.hv = 0.0f; r.prim = DDGI_MISS;
  r.octinv = 7u - ((r.idx < 0.0f ? 1u : 0u) | (r.idy < 0.0f ? 2u : 0u) | (r.idz < 0.0f ? 4u : 0u));   // sign of the SAFE inverse (d = -0 -> negative)
  r.ng = make_uint2(0u, 0x80000000u); r.tg = make_uint2(0u, 0u); r.sp = 0;
}

// 8 quantised child boxes against the ray.  Plane bytes become floats 1+q/32768 with one PRMT each (no I2F);
// near/far planes are picked by the direction sign, so no per-child min/max between lo and hi is needed.
__device__ __forceinline__ void node_step(const uint4* __restrict__ nodes, RayState& r, uint2* stack) {
  const int bit = 31 - __clz((int)r.ng.y);
  r.ng.y &= ~(1u << bit);
  if (r.ng.y & 0xFF000000u) stack[r.sp++] = r.ng;
  const uint32_t slot = ((uint32_t)bit - 24u) ^ r.octinv;
  const uint32_t rel = (uint32_t)__popc(r.ng.y & ~(0xFFFFFFFFu << slot) & 0xFFu);
  const uint4* nd = nodes + 5 * (size_t)(r.ng.x + rel);
  const uint4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3), n4 = __ldg(nd + 4);
  const bool nx = r.idx < 0.0f, ny = r.idy < 0.0f, nz = r.idz < 0.0f;
  // lo planes: x n2.xy, y n2.zw, z n3.xy ; hi planes: x n3.zw, y n4.xy, z n4.zw
  const uint32_t nearx0 = nx ? n3.z : n2.x, nearx1 = nx ? n3.w : n2.y, farx0 = nx ? n2.x : n3.z, farx1 = nx ? n2.y : n3.w;
  const uint32_t neary0 = ny ? n4.x : n2.z, neary1 = ny ? n4.y : n2.w, fary0 = ny ? n2.z : n4.x, fary1 = ny ? n2.w : n4.y;
  const uint32_t nearz0 = nz ? n4.z : n3.x, nearz1 = nz ? n4.w : n3.y, farz0 = nz ? n3.x : n4.z, farz1 = nz ? n3.y : n4.w;
  const float ax = __uint_as_float((n0.w & 0xFFu) << 23) * r.idx * 32768.0f;
  const float ay = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * r.idy * 32768.0f;
  const float az = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * r.idz * 32768.0f;
  const float bx = (__uint_as_float(n0.x) - r.ox) * r.idx - ax;
  const float by = (__uint_as_float(n0.y) - r.oy) * r.idy - ay;
  const float bz = (__uint_as_float(n0.z) - r.oz) * r.idz - az;
  const float kF = 1.0000004f;          // far side widened: conservative against rounding in the slab arithmetic
  const float axf = ax * kF, ayf = ay * kF, azf = az * kF, bxf = bx * kF, byf = by * kF, bzf = bz * kF;
  // meta bytes -> per-child hit bits (internal: 1 << (24 + (slot ^ octinv)); leaf: unary count << first triangle)
  const uint32_t octinv4 = r.octinv * 0x01010101u;
  uint32_t hitmask = 0;
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const uint32_t mw = w ? n1.w : n1.z;
    const uint32_t isInner = (mw & (mw << 1)) & 0x10101010u;
    const uint32_t innerMask = (isInner >> 4) * 0xFFu;
    const uint32_t bitIndex = (mw ^ (octinv4 & innerMask)) & 0x1F1F1F1Fu;
    const uint32_t childBits = (mw >> 5) & 0x07070707u;
    const uint32_t qnx = w ? nearx1 : nearx0, qny = w ? neary1 : neary0, qnz = w ? nearz1 : nearz0;
    const uint32_t qfx = w ? farx1 : farx0, qfy = w ? fary1 : fary0, qfz = w ? farz1 : farz0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t sel = 0x7604u | ((uint32_t)i << 4);
      const float tnx = fmaf(__uint_as_float(__byte_perm(qnx, 0x3F800000u, sel)), ax, bx);
      const float tny = fmaf(__uint_as_float(__byte_perm(qny, 0x3F800000u, sel)), ay, by);
      const float tnz = fmaf(__uint_as_float(__byte_perm(qnz, 0x3F800000u, sel)), az, bz);
      const float tfx = fmaf(__uint_as_float(__byte_perm(qfx, 0x3F800000u, sel)), axf, bxf);
      const float tfy = fmaf(__uint_as_float(__byte_perm(qfy, 0x3F800000u, sel)), ayf, byf);
      const float tfz = fmaf(__uint_as_float(__byte_perm(qfz, 0x3F800000u, sel)), azf, bzf);
      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
      const float tf = fminf(fminf(tfx, tfy), fminf(tfz, r.tmax));
      const uint32_t cb = (childBits >> (8 * i)) & 0xFFu, bi = (bitIndex >> (8 * i)) & 0xFFu;
      if (tn <= tf) hitmask |= cb << bi;        // empty slots are encoded lo=255 > hi=0 and never pass
    }
  }
  r.ng = make_uint2(n1.x, (hitmask & 0xFF000000u) | (n0.w >> 24));
  r.tg = make_uint2(n1.y, hitmask & 0x00FFFFFFu);
}

// One traversal iteration for this lane.  Returns false when the ray is finished.
template <bool ANY, bool CNT>
__device__ __forceinline__ bool trace_iter(const uint4* __restrict__ nodes, const float4* __restrict__ tris, RayState& r, uint2* stack,
                                           uint32_t& cntNodes, uint32_t& cntTris) {
  if (r.ng.y & 0xFF000000u) { node_step(nodes, r, stack); if (CNT) ++cntNodes; }
  else { r.tg = r.ng; r.ng = make_uint2(0u, 0u); }
  while (r.tg.y) {
    const int b = __ffs((int)r.tg.y) - 1;
    r.tg.y &= r.tg.y - 1u;
    const float4* tp = tris + 3 * (size_t)(r.tg.x + (uint32_t)b);
    const float4 a = __ldg(tp), bb = __ldg(tp + 1), cc = __ldg(tp + 2);
    if (CNT) ++cntTris;
    float t, u, v;
    if (tri_intersect(mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), mk3(a.x, a.y, a.z), mk3(a.w, bb.x, bb.y), mk3(bb.z, bb.w, cc.x),
                      r.tmin, ANY ? r.tmax : 3.0e38f, t, u, v)) {
      const uint32_t prim = __float_as_uint(cc.y);
      if (ANY) { r.prim = prim; return false; }
      // closest hit: r.tmax is the best t so far (initially the ray's tMax, exclusive); ties -> lowest triangle id
      if (t < r.tmax || (t == r.tmax && r.prim != DDGI_MISS && prim < r.prim)) { r.tmax = t; r.hu = u; r.hv = v; r.prim = prim; }
    }
  }
  if ((r.ng.y & 0xFF000000u) == 0u) {
    if (r.sp == 0) return false;
    r.ng = stack[--r.sp];
  }
  return true;
}

// ------------------------------------------------------------------------------------------------------------
// k_trace_primary: persistent warps; lanes that finish pick up the next probe rays (rgen:51-106)
// ------------------------------------------------------------------------------------------------------------
template <bool CNT>
__global__ void __launch_bounds__(128) k_trace_primary(const FastCtx c) {
  const TraceCtx& t = c.t;
  const uint32_t lane = threadIdx.x & 31u;
  const int P = t.nXZ * t.nXZ;
  uint2 stack[DDGI_STACK];
  RayState r; r.sp = 0; r.prim = DDGI_MISS; r.ng = make_uint2(0, 0); r.tg = make_uint2(0, 0);
  bool active = false, exhausted = false;
  uint32_t myIdx = 0, cn = 0, ct = 0, nRays = 0, nHits = 0;
  for (;;) {
    const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
    if (idle && !exhausted) {
      uint32_t base = 0;
      if (lane == (uint32_t)(__ffs((int)idle) - 1)) base = atomicAdd(&c.qcount[FAST_QC_PRIMARY], (uint32_t)__popc(idle));
      base = __shfl_sync(0xFFFFFFFFu, base, __ffs((int)idle) - 1);
      if (base + (uint32_t)__popc(idle) >= c.totalRays) exhausted = true;
      if (!active) {
        myIdx = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
        if (myIdx < c.totalRays) {
          const uint32_t slot = myIdx % (uint32_t)t.N, probeLin = myIdx / (uint32_t)t.N;
          const int layerLocal = (int)(probeLin / (uint32_t)P), probe = (int)(probeLin % (uint32_t)P);
          const uint32_t ray = t.rayOrder ? (uint32_t)t.rayOrder[slot] : slot;
          const V3 origin = mk3((float)(probe % t.nXZ), (float)(t.layerBegin + layerLocal), (float)(probe / t.nXZ)) * t.step + t.camOff;
          const float4 d4 = __ldg(&t.dirs[(size_t)layerLocal * t.N + ray]);
          ray_begin(r, origin, mk3(d4.x, d4.y, d4.z), t.tmin, t.tmax);
          active = t.numNodes != 0u;
          if (!active) c.hits[myIdx] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(DDGI_MISS));
          if (CNT) ++nRays;
        }
      }
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) break;
    // traversal until the warp has drained below the refill threshold
    for (;;) {
      if (active) {
        if (!trace_iter<false, CNT>(t.nodes, t.bvhTris, r, stack, cn, ct)) {
          active = false;
          c.hits[myIdx] = make_float4(r.prim == DDGI_MISS ? 0.0f : r.tmax, r.hu, r.hv, __uint_as_float(r.prim));
          if (CNT && r.prim != DDGI_MISS) ++nHits;
        }
      }
      const uint32_t act = __ballot_sync(0xFFFFFFFFu, active);
      if (act == 0u || (!exhausted && __popc(act) < FAST_FETCH_THRESHOLD)) break;
    }
  }
  if (CNT && t.stats) {
    for (int o = 16; o; o >>= 1) { cn += __shfl_xor_sync(0xFFFFFFFFu, cn, o); ct += __shfl_xor_sync(0xFFFFFFFFu, ct, o);
                                   nRays += __shfl_xor_sync(0xFFFFFFFFu, nRays, o); nHits += __shfl_xor_sync(0xFFFFFFFFu, nHits, o); }
    if (lane == 0) { atomicAdd(&t.stats[0], (unsigned long long)nRays); atomicAdd(&t.stats[1], (unsigned long long)nHits);
                     atomicAdd(&t.stats[3], (unsigned long long)cn); atomicAdd(&t.stats[4], (unsigned long long)ct); }
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_shade_prep: rchit:27-42,70-77 (hit attributes + surface) and the decision which shadow rays exist
// ------------------------------------------------------------------------------------------------------------
#define PREP_BLOCK 256
#define PREP_WARPS (PREP_BLOCK / 32)
__global__ void __launch_bounds__(PREP_BLOCK) k_shade_prep(const FastCtx c) {
  const TraceCtx& t = c.t;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
  const bool inRange = idx < c.totalRays;
  const int P = t.nXZ * t.nXZ;
  bool isHit = false;
  V3 pos = mk3(0.0f), normal = mk3(0.0f);
  if (inRange) {
    const float4 h = c.hits[idx];
    const uint32_t hp = __float_as_uint(h.w);
    if (hp != DDGI_MISS) {
      isHit = true;
      // rchit:27-38 through the per-triangle record (five independent 16-byte loads instead of the index/vertex gather)
      const float4* st = c.shadeTris + 5 * (size_t)hp;
      const float4 s0 = __ldg(st), s1 = __ldg(st + 1), s2 = __ldg(st + 2), s3 = __ldg(st + 3), s4 = __ldg(st + 4);
      const uint32_t ii = __float_as_uint(s4.z), material = __float_as_uint(s4.w);
      const InstXform& xf = t.xforms[ii];
      const float b0 = 1.0f - h.y - h.z, b1 = h.y, b2 = h.z;
      pos = mk3(s0.x, s0.y, s0.z) * b0 + mk3(s0.w, s1.x, s1.y) * b1 + mk3(s1.z, s1.w, s2.x) * b2;
      pos = xform_point(xf.o2w, pos);
      normal = normalize3(mk3(s2.y, s2.z, s2.w) * b0 + mk3(s3.x, s3.y, s3.z) * b1 + mk3(s3.w, s4.x, s4.y) * b2);
      normal = normalize3(mk3(normal.x * xf.w2o[0] + normal.y * xf.w2o[1] + normal.z * xf.w2o[2],
                              normal.x * xf.w2o[3] + normal.y * xf.w2o[4] + normal.z * xf.w2o[5],
                              normal.x * xf.w2o[6] + normal.y * xf.w2o[7] + normal.z * xf.w2o[8]));
      const anerend_material_info_t* m = &t.materials[material];
      V3 albedo;
      if (m->base_col_fac[0] + m->base_col_fac[1] + m->base_col_fac[2] > 0.1f) albedo = mk3(m->base_col_fac[0], m->base_col_fac[1], m->base_col_fac[2]);
      else {
        // vertex colour decides (rchit:71,75; bindless.glsl:187-194): rare, so the colours stay in the vertex buffer
        const DInstance in = t.instances[ii];
        const uint32_t prim = hp - in.firstTri;                                                     // gl_PrimitiveID
        const anerend_vertex_t* v0 = &t.vertices[in.vertexOffset + t.indices[in.indexOffset + 3 * prim + 0]];
        const anerend_vertex_t* v1 = &t.vertices[in.vertexOffset + t.indices[in.indexOffset + 3 * prim + 1]];
        const anerend_vertex_t* v2 = &t.vertices[in.vertexOffset + t.indices[in.indexOffset + 3 * prim + 2]];
        const V3 color = mk3(v0->color[0], v0->color[1], v0->color[2]) * b0 + mk3(v1->color[0], v1->color[1], v1->color[2]) * b1 +
                         mk3(v2->color[0], v2->color[1], v2->color[2]) * b2;
        albedo = mk3(to_linear1(color.x), to_linear1(color.y), to_linear1(color.z));
      }
      c.surf[3 * (size_t)idx + 0] = make_float4(pos.x, pos.y, pos.z, normal.x);
      c.surf[3 * (size_t)idx + 1] = make_float4(normal.y, normal.z, albedo.x, albedo.y);
      c.surf[3 * (size_t)idx + 2] = make_float4(albedo.z, __uint_as_float(material), 0.0f, 0.0f);
      c.lit[idx] = 0u;
    } else {
      // rmiss:20 + rgen:109
      const uint32_t slot = idx % (uint32_t)t.N, probeLin = idx / (uint32_t)t.N;
      const int layerLocal = (int)(probeLin / (uint32_t)P), probe = (int)(probeLin % (uint32_t)P);
      const uint32_t ray = t.rayOrder ? (uint32_t)t.rayOrder[slot] : slot;
      const size_t o = ((size_t)(t.layerBegin + layerLocal) * P + probe) * t.N + ray;
      t.rayIrr[o] = make_uint2((uint32_t)f2h(t.sky.x) | ((uint32_t)f2h(t.sky.y) << 16), (uint32_t)f2h(t.sky.z) | (0x3C00u << 16));
      if (t.records) { anerend_ddgi_ray_record_t rr; rr.prim = DDGI_MISS; rr.t = 0.0f; rr.u = 0.0f; rr.v = 0.0f;
                       rr.radiance[0] = t.sky.x; rr.radiance[1] = t.sky.y; rr.radiance[2] = t.sky.z; rr.lit_mask = 0u; t.records[o] = rr; }
    }
  }
  // Which queues does this ray enter?  bit q: shadow ray towards light q (0 = sun, rchit:45-66: every hit; point
  // lights, extension E1: only where the light can contribute); bit 32: the compacted hit list k_shade runs over.
  uint32_t needMask = isHit ? 1u : 0u;
  if (isHit && (t.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS)) {
    for (uint32_t li = 0; li < t.numLights; ++li) {
      const anerend_light_t* lt = &t.lights[li];
      if (lt->color[3] > 0.5f) {
        const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
        const V3 dv = lp - pos; const float dist = sqrtf(dot3(dv, dv));
        if (dist < lt->world_pos[3]) { const V3 L = normalize3(lp - pos); if (dot3(normal, L) > 0.0f) needMask |= 1u << (1u + li); }
      }
    }
  }
  // block-aggregated append: one atomic per queue per block (a per-warp atomic on nine neighbouring counters was the
  // whole cost of this kernel), order inside the block = ray order
  __shared__ uint32_t sCnt[33][PREP_WARPS];
  __shared__ uint32_t sBase[33];
  const uint32_t warp = threadIdx.x >> 5, nq = c.numQueues;
  for (uint32_t q = 0; q <= nq; ++q) {
    const bool need = q < nq ? ((needMask >> q) & 1u) != 0u : isHit;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
    if (lane == 0) sCnt[q][warp] = (uint32_t)__popc(m);
  }
  __syncthreads();
  if (threadIdx.x <= nq) {
    const uint32_t q = threadIdx.x;
    uint32_t sum = 0;
    for (uint32_t w = 0; w < PREP_WARPS; ++w) { const uint32_t v = sCnt[q][w]; sCnt[q][w] = sum; sum += v; }
    sBase[q] = sum ? atomicAdd(&c.qcount[q < nq ? FAST_QC(q) : FAST_QC_HITS], sum) : 0u;
  }
  __syncthreads();
  for (uint32_t q = 0; q <= nq; ++q) {
    const bool need = q < nq ? ((needMask >> q) & 1u) != 0u : isHit;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
    if (need) {
      const uint32_t o = sBase[q] + sCnt[q][warp] + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (q < nq) c.queues[(size_t)q * c.totalRays + o] = idx; else c.hitq[o] = idx;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_trace_shadow: persistent warps over the concatenated per-light queues, any-hit (rchit:48-66)
// ------------------------------------------------------------------------------------------------------------
template <bool CNT>
__global__ void __launch_bounds__(128) k_trace_shadow(const FastCtx c) {
  const TraceCtx& t = c.t;
  __shared__ uint32_t sPrefix[34];
  if (threadIdx.x == 0) { uint32_t s = 0; for (uint32_t q = 0; q < c.numQueues; ++q) { sPrefix[q] = s; s += c.qcount[FAST_QC(q)]; } sPrefix[c.numQueues] = s; }
  __syncthreads();
  const uint32_t total = sPrefix[c.numQueues];
  const uint32_t lane = threadIdx.x & 31u;
  uint2 stack[DDGI_STACK];
  RayState r; r.sp = 0; r.prim = DDGI_MISS; r.ng = make_uint2(0, 0); r.tg = make_uint2(0, 0);
  bool active = false, exhausted = total == 0u;
  uint32_t myRay = 0, myBit = 0, cn = 0, ct = 0, nRays = 0;
  const V3 sunL = normalize3(-t.lightDir);
  for (;;) {
    const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
    if (idle && !exhausted) {
      uint32_t base = 0;
      if (lane == (uint32_t)(__ffs((int)idle) - 1)) base = atomicAdd(&c.qcount[FAST_QC_SHADOW], (uint32_t)__popc(idle));
      base = __shfl_sync(0xFFFFFFFFu, base, __ffs((int)idle) - 1);
      if (base + (uint32_t)__popc(idle) >= total) exhausted = true;
      if (!active) {
        const uint32_t w = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
        if (w < total) {
          uint32_t q = 0; while (q + 1 < c.numQueues && sPrefix[q + 1] <= w) ++q;
          myRay = c.queues[(size_t)q * c.totalRays + (w - sPrefix[q])]; myBit = q;
          const float4 s0 = c.surf[3 * (size_t)myRay];
          const V3 pos = mk3(s0.x, s0.y, s0.z);
          V3 dir = sunL; float tmax = t.tmax;
          if (q > 0) {
            const anerend_light_t* lt = &t.lights[q - 1];
            const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
            const V3 dv = lp - pos; tmax = sqrtf(dot3(dv, dv)); dir = normalize3(lp - pos);
          }
          ray_begin(r, pos, dir, t.tmin, tmax);
          active = true;
          if (CNT) ++nRays;
        }
      }
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) break;
    for (;;) {
      if (active) {
        if (!trace_iter<true, CNT>(t.nodes, t.bvhTris, r, stack, cn, ct)) {
          active = false;
          if (r.prim == DDGI_MISS) atomicOr(&c.lit[myRay], 1u << myBit);      // miss shader: shadow = 0 -> lit
        }
      }
      const uint32_t act = __ballot_sync(0xFFFFFFFFu, active);
      if (act == 0u || (!exhausted && __popc(act) < FAST_FETCH_THRESHOLD)) break;
    }
  }
  if (CNT && t.stats) {
    for (int o = 16; o; o >>= 1) { cn += __shfl_xor_sync(0xFFFFFFFFu, cn, o); ct += __shfl_xor_sync(0xFFFFFFFFu, ct, o); nRays += __shfl_xor_sync(0xFFFFFFFFu, nRays, o); }
    if (lane == 0) { atomicAdd(&t.stats[2], (unsigned long long)nRays); atomicAdd(&t.stats[5], (unsigned long long)cn); atomicAdd(&t.stats[6], (unsigned long long)ct); }
  }
}

}  // namespace ddgi
