// ddgi_fast.cuh — the production (wavefront) kernels of "IrradianceProbeRT" for sm_100a.
//
// The reference runs rgen -> traceRayEXT -> rchit (-> traceRayEXT shadow) -> rmiss per ray inside the RT pipeline
// (irradiance_probe_update.rgen:48-111, .rchit:24-111, .rmiss:17-21).  A one-thread-per-ray transcription of that
// (k_probe_trace in ddgi_kernels.cuh) keeps only ~8 of 32 lanes busy (profiles/r01_*), so the pass is split into
// stages that each keep warps full:
//   k_trace<false>   persistent warps, closest-hit BVH8 traversal of the probe rays
//   k_shade_prep     hit attributes (rchit:27-42,70-77) -> surface record; shadow rays binned into one queue per light
//   k_trace<true>    persistent warps, any-hit traversal over the per-light queues (coherent direction / target)
//   k_shade          calcLight + calcIndirectDiffuseLight + emissive (rchit:83-109), RGBA16F store (rgen:109)  [ddgi_shade.cu]
// Per-ray arithmetic is the same sequence of fp32 operations as the simple path and the oracle (the box test differs,
// but boxes are conservative and the triangle test decides), so results stay bit-identical.
#pragma once
#ifndef __CUDACC__
#error "ddgi_fast.cuh is CUDA-only"
#endif
#include "ddgi_wavefront.h"

namespace ddgi {

// Tunables of the persistent traversal kernels (FastCtx::tune; DDGI_TUNE="a,b,c" overrides them for sweeps)
#define TUNE_NODE_STEPS 3           // box-test steps per loop iteration before the triangle rounds
#define TUNE_REFILL_MIN 4           // refill from the staged batch once this many lanes are idle
#define TUNE_STEP_MIN 12            // (nodeSteps > 1) keep stepping only while this many lanes can
#define TUNE_TRI_MIN 8              // (nodeSteps > 1) keep running triangle rounds only while this many lanes hold triangles
#define FAST_CHUNK 32               // ray indices taken per atomic on the global cursor (one staged batch: shorter tail at the end of the launch)

// Node-phase state of a ray lives in registers; everything only the triangle rounds touch (direction, parked triangle
// groups, hit attributes, ray tag) lives in shared memory, so the box test keeps 8 blocks of 128 threads on an SM.
struct RayState {
  float ox, oy, oz, idx, idy, idz, tmax;
  uint32_t octm;                    // XOR-permutation masks of the ray's octant: 0x55/0x33/0x0F in bytes 0/1/2 where the
                                    // direction is POSITIVE on x/y/z (bit index ^ octinv), octinv itself in bits 24..26
  uint2 ng;                         // current node group: child base | hit-children bits (ray order) << 24 | imask
  int sp;
  uint32_t npend;                   // parked triangle groups (0..2)
  // two-level structures (TWO): the instance whose mesh is being walked (DDGI_MISS: instance level) — ox.. / idx.. / octm then
  // describe the ray in its object space — and the slack of the slab parameters there (ObjRay::pad)
  uint32_t inst; float pad;         // pad: object units (ObjRay::padLen); the slab parameters of axis k move by pad * |1/d'_k|
};

#define LS_DIR 0                    // per-lane shared state, SoA [field][lane] of 16 bytes: (d.xyz, tag)
#define LS_RES 1                    //   (t | hu | hv | prim) closest hit so far;  any-hit: (-,-,-, aux = light bit)
#define LS_PEND_A 2                 //   (triBase, hit bits, triMask, -)
#define LS_PEND_B 3
#define LS_FIELDS 4
#define LS_ORG 4                    //   TWO only: (world-space origin, instance of the closest hit so far)

// staged ray (shared memory, 3 x float4): (o.xyz, tmax) (d.xyz, tag) (1/d.xyz, aux)
__device__ __forceinline__ void ray_stage(float4* s, V3 o, V3 d, float tmax, uint32_t tag, uint32_t aux) {
  s[0] = make_float4(o.x, o.y, o.z, tmax);
  s[1] = make_float4(d.x, d.y, d.z, __uint_as_float(tag));
  s[2] = make_float4(1.0f / safe_dir(d.x), 1.0f / safe_dir(d.y), 1.0f / safe_dir(d.z), __uint_as_float(aux));
}

// Two plane bytes of q -> floats 1024+q: ONE PRMT (alu pipe) builds the half2 {0x64:q_a, 0x64:q_b}, the two
// half->float conversions issue on the fma pipe.  profiles/r02_ncu_trace_v2: the box test was alu-pipe bound
// (alu 68 %, fma 20 % busy) with one PRMT per plane byte.  k64 = 0x64646464 is a kernel ARGUMENT: a literal would be
// folded into the immediate slot and the selector re-materialised in a register for every PRMT.
template <int PAIR>
__device__ __forceinline__ float2 planes2(uint32_t q, uint32_t k64) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(q), "r"(k64), "n"(PAIR ? 0x4342 : 0x4140));
  return __half22float2(*reinterpret_cast<const __half2*>(&d));
}

#define CHILD_CONST(S) ((7u << (3 * (S))) | (1u << (24 + (S))))

// One word of four children per axis side -> hit bits of those four children (slot order).
template <int W>
__device__ __forceinline__ uint32_t four_children(const uint32_t qnx, const uint32_t qny, const uint32_t qnz, const uint32_t qfx, const uint32_t qfy,
                                                  const uint32_t qfz, const float ax, const float ay, const float az, const float bx, const float by,
                                                  const float bz, const float axf, const float ayf, const float azf, const float bxf, const float byf,
                                                  const float bzf, const float tmin, const float tmax, const uint32_t k64) {
  uint32_t acc = 0;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float2 nx = p ? planes2<1>(qnx, k64) : planes2<0>(qnx, k64), ny = p ? planes2<1>(qny, k64) : planes2<0>(qny, k64),
                 nz = p ? planes2<1>(qnz, k64) : planes2<0>(qnz, k64);
    const float2 fx = p ? planes2<1>(qfx, k64) : planes2<0>(qfx, k64), fy = p ? planes2<1>(qfy, k64) : planes2<0>(qfy, k64),
                 fz = p ? planes2<1>(qfz, k64) : planes2<0>(qfz, k64);
    {
      const float tn = fmaxf(fmaxf(fmaf(nx.x, ax, bx), fmaf(ny.x, ay, by)), fmaxf(fmaf(nz.x, az, bz), tmin));
      const float tf = fminf(fminf(fmaf(fx.x, axf, bxf), fmaf(fy.x, ayf, byf)), fminf(fmaf(fz.x, azf, bzf), tmax));
      acc |= (tn <= tf) ? CHILD_CONST(4 * W + 2 * p) : 0u;
    }
    {
      const float tn = fmaxf(fmaxf(fmaf(nx.y, ax, bx), fmaf(ny.y, ay, by)), fmaxf(fmaf(nz.y, az, bz), tmin));
      const float tf = fminf(fminf(fmaf(fx.y, axf, bxf), fmaf(fy.y, ayf, byf)), fminf(fmaf(fz.y, azf, bzf), tmax));
      acc |= (tn <= tf) ? CHILD_CONST(4 * W + 2 * p + 1) : 0u;
    }
  }
  return acc;
}

// bit s of the 8-bit mask x -> bit s ^ octinv (children in the order this ray meets them, Ylitie 2017 §3.2)
__device__ __forceinline__ uint32_t octant_permute(uint32_t x, uint32_t octm) {
  uint32_t t = ((x >> 1) ^ x) & octm & 0x55u; x ^= t | (t << 1);
  t = ((x >> 2) ^ x) & (octm >> 8) & 0x33u; x ^= t | (t << 2);
  t = ((x >> 4) ^ x) & (octm >> 16) & 0x0Fu; x ^= t | (t << 4);
  return x;
}

// 8 quantised child boxes against the ray; near/far planes are picked by the direction sign, so no per-child min/max
// between lo and hi is needed.  Consumes the front child of r.ng (ray order), leaves the hit internal children in r.ng
// and returns (triBase, hit triangle bits, triMask).
template <bool TWO = false>
__device__ __forceinline__ uint4 node_step(const uint4* __restrict__ nodes, RayState& r, uint2* stack, const float tmin, const uint32_t k64) {
  const int bit = 31 - __clz((int)r.ng.y);
  r.ng.y &= ~(1u << bit);
  if (r.ng.y & 0xFF000000u) stack[r.sp++] = r.ng;
  const uint32_t slot = ((uint32_t)bit - 24u) ^ (r.octm >> 24);
  const uint32_t rel = (uint32_t)__popc(r.ng.y & ~(0xFFFFFFFFu << slot) & 0xFFu);
  const uint4* nd = nodes + 5 * (size_t)(r.ng.x + rel);
  const uint4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3), n4 = __ldg(nd + 4);
  const bool nx = r.idx < 0.0f, ny = r.idy < 0.0f, nz = r.idz < 0.0f;
  // lo planes: x n2.xy, y n2.zw, z n3.xy ; hi planes: x n3.zw, y n4.xy, z n4.zw
  const uint32_t nearx0 = nx ? n3.z : n2.x, nearx1 = nx ? n3.w : n2.y, farx0 = nx ? n2.x : n3.z, farx1 = nx ? n2.y : n3.w;
  const uint32_t neary0 = ny ? n4.x : n2.z, neary1 = ny ? n4.y : n2.w, fary0 = ny ? n2.z : n4.x, fary1 = ny ? n2.w : n4.y;
  const uint32_t nearz0 = nz ? n4.z : n3.x, nearz1 = nz ? n4.w : n3.y, farz0 = nz ? n3.x : n4.z, farz1 = nz ? n3.y : n4.w;
  // t(q) = (p + q*2^e - o) / d = (1024 + q) * a + b  with a = 2^e / d,  b = (p - o) / d - 1024 a
  const float ax = __uint_as_float((n0.w & 0xFFu) << 23) * r.idx;
  const float ay = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * r.idy;
  const float az = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * r.idz;
  float bx = fmaf(-1024.0f, ax, (__uint_as_float(n0.x) - r.ox) * r.idx);
  float by = fmaf(-1024.0f, ay, (__uint_as_float(n0.y) - r.oy) * r.idy);
  float bz = fmaf(-1024.0f, az, (__uint_as_float(n0.z) - r.oz) * r.idz);
  const float kF = 1.0000004f;          // far side widened: conservative against rounding in the slab arithmetic
  const float axf = ax * kF, ayf = ay * kF, azf = az * kF;
  float bxf = bx * kF, byf = by * kF, bzf = bz * kF;
  if (TWO) {      // object space: what the transforms may have lost
    const float px = r.pad * fabsf(r.idx), py = r.pad * fabsf(r.idy), pz = r.pad * fabsf(r.idz);
    bx -= px; by -= py; bz -= pz; bxf += px; byf += py; bzf += pz;
  }
  const uint32_t acc = four_children<0>(nearx0, neary0, nearz0, farx0, fary0, farz0, ax, ay, az, bx, by, bz, axf, ayf, azf, bxf, byf, bzf, tmin, r.tmax, k64) |
                       four_children<1>(nearx1, neary1, nearz1, farx1, fary1, farz1, ax, ay, az, bx, by, bz, axf, ayf, azf, bxf, byf, bzf, tmin, r.tmax, k64);
  const uint32_t imask = n0.w >> 24;
  r.ng = make_uint2(n1.x, (octant_permute((acc >> 24) & imask, r.octm) << 24) | imask);
  return make_uint4(n1.y, acc & n1.z, n1.z, 0u);
}

// ------------------------------------------------------------------------------------------------------------
// k_trace<ANY>: persistent warps, one ray per lane.
//   ANY = false  closest hit for the probe rays                 (rgen:51-106)   -> hits[ray]
//   ANY = true   any-hit over the per-light shadow-ray queues   (rchit:48-66)   -> lit[ray] |= 1 << light
// Work distribution: ray set-up runs 32 lanes wide into a shared-memory batch and idle lanes refill from it every
// iteration; one atomic on the global cursor per FAST_CHUNK rays.  Triangles found by the box tests are parked in
// shared memory (two groups per lane) and tested one per lane per iteration in a separate, short section, so lanes
// that found none keep stepping through nodes instead of idling in a per-lane triangle loop.
// ------------------------------------------------------------------------------------------------------------
template <bool ANY, bool CNT, bool EXT = false, bool TWO = false>
__global__ void __launch_bounds__(128, TWO ? 7 : 8) k_trace(const FastCtx c) {
  const TraceCtx& t = c.t;
  __shared__ float4 sRay[4][32][3];
  __shared__ float4 sLane[4][LS_FIELDS + (TWO ? 1 : 0)][32];
  __shared__ uint32_t sPrefix[34];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, ltMask = (1u << lane) - 1u;
  uint32_t total = c.totalRays;
  if (ANY) {
    if (threadIdx.x == 0) { uint32_t s = 0; for (uint32_t q = 0; q < c.numQueues; ++q) { sPrefix[q] = s; s += c.qcount[FAST_QC(q)]; } sPrefix[c.numQueues] = s; }
    __syncthreads();
    total = sPrefix[c.numQueues];
  }
  uint32_t* cursor = (!ANY && c.primaryCursor) ? c.primaryCursor : &c.qcount[ANY ? FAST_QC_SHADOW : FAST_QC_PRIMARY];
  const V3 sunL = c.rawSunDir ? -t.lightDir : normalize3(-t.lightDir);
  const uint32_t nodeSteps = max(1u, c.tune[0]), refillMin = max(1u, ANY ? c.tune[2] : c.tune[1]);
  const uint32_t stepMin = max(1u, c.tune[4]), triMin = max(1u, c.tune[5]);
  const uint32_t k64 = c.tune[3];          // 0x64646464, opaque to ptxas on purpose (see planes2)
  const uint32_t chunk = c.tune[6] ? c.tune[6] : (uint32_t)FAST_CHUNK;
  float4* myDir = &sLane[warp][LS_DIR][lane];
  float4* myRes = &sLane[warp][LS_RES][lane];
  uint4* myPend = reinterpret_cast<uint4*>(&sLane[warp][LS_PEND_A][lane]);      // slot B = myPend + 32
  float4* myOrg = &sLane[warp][TWO ? LS_ORG : LS_DIR][lane];                     // TWO only
  uint2 stack[DDGI_STACK];
  RayState r; r.sp = 0; r.npend = 0; r.ng = make_uint2(0, 0); r.tmax = 0.0f; r.octm = 0;
  r.inst = DDGI_MISS; r.pad = 0.0f;
  r.ox = r.oy = r.oz = 0.0f; r.idx = r.idy = r.idz = 1.0f;
  bool active = false, exhausted = (total == 0u), wantChange = false;      // wantChange (TWO): the top of the lane's stack is a level change
  uint32_t chunkNext = 0, chunkEnd = 0, staged = 0, stagedPos = 0;      // warp-uniform
  uint32_t cn = 0, ct = 0, nRays = 0, nHits = 0, nCand = 0, cs = 0;      // cs: level changes of a two-level walk (instance entries + returns)
  for (;;) {
    // ---------------------------------------------------------------- refill idle lanes from the staged batch
    uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
    if ((uint32_t)__popc(idle) < refillMin) idle = 0u;
    while (idle && !(exhausted && staged == 0u)) {
      if (staged == 0u) {
        if (chunkNext == chunkEnd) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(cursor, chunk);
          base = __shfl_sync(0xFFFFFFFFu, base, 0);
          if (base >= total) { exhausted = true; break; }
          chunkNext = base; chunkEnd = min(base + chunk, total);
        }
        const uint32_t n = min(32u, chunkEnd - chunkNext);
        uint32_t nStaged = n;
        __syncwarp();
        if (ANY) {
          // shadow rays: set-up, then the light map's shortcuts (ddgi_kernels.cuh "Light maps") 32 lanes wide — the far end of a
          // point-light ray is cut back to where geometry can begin, and the texel's triangle (the nearest one to the light in the
          // ray's direction) gets the exact test first: a hit ends the ray here, occluded, before it ever enters the hierarchy
          bool keep = false; V3 pos = mk3(0.0f), dir = mk3(0.0f, 0.0f, 1.0f); float tmax = 0.0f; uint32_t ray = 0u, q = 0u;
          if (lane < n) {
            const uint32_t w = chunkNext + lane;
            while (q + 1 < c.numQueues && sPrefix[q + 1] <= w) ++q;
            ray = c.queues[(size_t)q * c.totalRays + (w - sPrefix[q])];
            const float4 s0 = c.surf[(size_t)c.surfStride * ray];
            pos = mk3(s0.x, s0.y, s0.z); dir = sunL; tmax = t.tmax; keep = true;
            if (q > 0) {
              const anerend_light_t* lt = &t.lights[q - 1];
              const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
              const V3 dv = lp - pos; const float dist = sqrtf(dot3(dv, dv)); dir = normalize3(lp - pos);
              const LmEntry e = lightmap_lookup(t, q - 1, lp, pos);
              tmax = lightmap_tmax(e, dist);
              keep = !lightmap_candidate_occludes(t, e, pos, dir, t.tmin, dist);
            } else if (t.smRes != 0u && !c.rawSunDir) {
              const SmEntry e = sunmap_lookup(t, pos);
              tmax = sunmap_tmax(e, dot3(pos, t.smS), t.tmax);
              keep = !sunmap_candidate_occludes(t, e, pos, dir, t.tmin, t.tmax);
            }
          }
          const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
          if (keep) ray_stage(&sRay[warp][__popc(km & ltMask)][0], pos, dir, tmax, ray, q);
          nStaged = (uint32_t)__popc(km);
          if (CNT) { nRays += (lane < n) ? 1u : 0u; nCand += (lane < n && !keep) ? 1u : 0u; }
        } else
        if (lane < n) {                                             // ray set-up, one ray per lane
          const uint32_t w = chunkNext + lane;
          float4* dst = &sRay[warp][lane][0];
          if (!ANY && EXT) {                                        // one ray per pixel, staged by the caller (SpecularGIRT)
            const float4 ea = __ldg(&c.extRays[2 * (size_t)w]), eb = __ldg(&c.extRays[2 * (size_t)w + 1]);
            ray_stage(dst, mk3(ea.x, ea.y, ea.z), mk3(eb.x, eb.y, eb.z), ea.w, w, 0u);
          } else if (!ANY) {
            const uint32_t slot = w % (uint32_t)t.N, probeLin = w / (uint32_t)t.N;
            const int layerLocal = (int)(probeLin / (uint32_t)t.Pl), probe = stripe_probe((int)(probeLin % (uint32_t)t.Pl), t.nXZ, t.rowOffset, t.rowStride);
            const uint32_t ray = t.rayOrder ? (uint32_t)t.rayOrder[slot] : slot;
            const int layer = __ldg(&t.layerMap[layerLocal]);
            const V3 origin = mk3((float)(probe % t.nXZ), (float)layer, (float)(probe / t.nXZ)) * t.step + t.camOff;
            const float4 d4 = __ldg(&t.dirs[(size_t)layer * t.N + ray]);
            ray_stage(dst, origin, mk3(d4.x, d4.y, d4.z), t.tmax, w, 0u);
          }
        }
        __syncwarp();
        chunkNext += n; staged = nStaged; stagedPos = 0;
      }
      const uint32_t take = min((uint32_t)__popc(idle), staged), rank = (uint32_t)__popc(idle & ltMask);
      if (!active && rank < take) {
        const float4* src = &sRay[warp][stagedPos + rank][0];
        const float4 a = src[0], b = src[1], cc = src[2];
        r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tmax = a.w; r.idx = cc.x; r.idy = cc.y; r.idz = cc.z;
        // sign of the SAFE inverse (d = -0 -> negative); octinv bit k set = direction POSITIVE on axis k
        const uint32_t octinv = 7u - ((cc.x < 0.0f ? 1u : 0u) | (cc.y < 0.0f ? 2u : 0u) | (cc.z < 0.0f ? 4u : 0u));
        r.octm = ((octinv & 1u) ? 0x55u : 0u) | ((octinv & 2u) ? 0x3300u : 0u) | ((octinv & 4u) ? 0x0F0000u : 0u) | (octinv << 24);
        r.ng = make_uint2(0u, 0x80000000u); r.sp = 0; r.npend = 0;
        *myDir = b;                                                 // (d.xyz, tag)
        *myRes = make_float4(0.0f, 0.0f, 0.0f, ANY ? cc.w : __uint_as_float(DDGI_MISS));
        if (TWO) { *myOrg = make_float4(a.x, a.y, a.z, __uint_as_float(DDGI_MISS)); r.inst = DDGI_MISS; r.pad = 0.0f; wantChange = false; }
        if (CNT && !ANY) ++nRays;
        if (!ANY && t.numNodes == 0u) c.hits[__float_as_uint(b.w)] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(DDGI_MISS));
        else active = true;
      }
      stagedPos += take; staged -= take;
      idle = __ballot_sync(0xFFFFFFFFu, !active);
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) { if (exhausted && staged == 0u) break; else continue; }

    // ---------------------------------------------------------------- node phase: up to `nodeSteps` box-test steps for
    // the lanes that have a node to visit and a free parking slot; stops early when too few lanes can still step
    bool finished = false;
    for (uint32_t k = 0;;) {
      const bool hasNode = (r.ng.y & 0xFF000000u) != 0u || r.sp > 0;
      if (active && !finished) {
        if (hasNode && r.npend < 2u && !(TWO && wantChange)) {
          bool step = true;
          if ((r.ng.y & 0xFF000000u) == 0u) {
            const uint2 e = stack[r.sp - 1];
            // two-level structure: an entry without child bits moves the lane between the spaces; it stays on the stack and the
            // lane sits out the rest of this node phase — the level changes of a warp are worked off together below (taken one by
            // one inside the box-test steps they ran at 4 of 32 lanes and were a sixth of the kernel's instructions)
            if (TWO && e.y == 0u) { step = false; wantChange = true; }
            else { r.ng = e; --r.sp; }
          }
          if (step) {
            const uint4 tg = node_step<TWO>(t.nodes, r, stack, t.tmin, k64);
            if (CNT) ++cn;
            if (TWO && r.inst == DDGI_MISS) {
              // instance level: the instances met become stack entries, in front of this node's remaining children
              if (tg.y) {
                if (r.ng.y & 0xFF000000u) stack[r.sp++] = r.ng;
                r.ng.y &= 0x00FFFFFFu;
                for (uint32_t m = tg.y; m != 0u; m &= m - 1u) {
                  const uint32_t bb = (uint32_t)__ffs((int)m) - 1u;
                  const uint32_t rec = tg.x + (uint32_t)__popc(tg.z & ((1u << bb) - 1u));
                  stack[r.sp++] = make_uint2(TWO_INSTANCE | __float_as_uint(__ldg(t.bvhTris + 3 * (size_t)rec).x), 0u);
                }
              }
            } else if (tg.y) { myPend[32u * r.npend] = TWO ? make_uint4(tg.x, tg.y, tg.z, r.inst) : tg; ++r.npend; }
          }
        } else if (!hasNode && r.npend == 0u) finished = true;
      }
      if (++k >= nodeSteps) break;
      const uint32_t can = __ballot_sync(0xFFFFFFFFu, active && !finished && r.npend < 2u && !(TWO && wantChange) && ((r.ng.y & 0xFF000000u) != 0u || r.sp > 0));
      if ((uint32_t)__popc(can) < stepMin) break;
    }
    // ---------------------------------------------------------------- level changes (two-level structures, ddgi_kernels.cuh traverse2):
    // back to the world-space ray, or into an instance's object space
    if (TWO && __ballot_sync(0xFFFFFFFFu, wantChange) != 0u) {
      if (wantChange) {
        wantChange = false;
        if (CNT) ++cs;
        uint2 e = stack[--r.sp];
        // a return directly followed by the next instance: enter it without the detour through the world-space ray
        if (e.x == TWO_RESTORE && r.sp > 0) { const uint2 e2 = stack[r.sp - 1]; if (e2.y == 0u) { e = e2; --r.sp; } }
        const float4 o4 = *myOrg, d4 = *myDir;
        if (e.x == TWO_RESTORE) {
          r.ox = o4.x; r.oy = o4.y; r.oz = o4.z;
          r.idx = 1.0f / safe_dir(d4.x); r.idy = 1.0f / safe_dir(d4.y); r.idz = 1.0f / safe_dir(d4.z);
          r.pad = 0.0f; r.inst = DDGI_MISS;
        } else {
          r.inst = e.x & ~TWO_INSTANCE;
          const InstXform& xf = t.xforms[r.inst];
          const ObjRay q = object_ray(xf, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z), r.tmax);
          r.ox = q.o.x; r.oy = q.o.y; r.oz = q.o.z; r.idx = q.id.x; r.idy = q.id.y; r.idz = q.id.z; r.pad = q.padLen;
          stack[r.sp++] = make_uint2(TWO_RESTORE, 0u);
          r.ng = make_uint2(__float_as_uint(xf._pad[0]), 0x80000000u);
        }
        const uint32_t octinv = 7u - ((r.idx < 0.0f ? 1u : 0u) | (r.idy < 0.0f ? 2u : 0u) | (r.idz < 0.0f ? 4u : 0u));
        r.octm = ((octinv & 1u) ? 0x55u : 0u) | ((octinv & 2u) ? 0x3300u : 0u) | ((octinv & 4u) ? 0x0F0000u : 0u) | (octinv << 24);
      }
    }
    // ---------------------------------------------------------------- triangle rounds: one parked triangle per lane and round
    for (uint32_t pend = __ballot_sync(0xFFFFFFFFu, active && !finished && r.npend != 0u); pend != 0u;) {
      if (active && !finished && r.npend != 0u) {
        uint4 g = myPend[0];
        const int b = __ffs((int)g.y) - 1;
        const float4* tp = t.bvhTris + 3 * (size_t)(g.x + (uint32_t)__popc(g.z & ((1u << b) - 1u)));
        const float4 a = __ldg(tp), bb = __ldg(tp + 1), cc = __ldg(tp + 2);
        const float4 d4 = *myDir;
        g.y &= g.y - 1u;
        if (g.y != 0u) myPend[0].y = g.y;
        else { if (r.npend == 2u) myPend[0] = myPend[32]; --r.npend; }
        if (CNT) ++ct;
        float tt, u, v;
        V3 ro = mk3(r.ox, r.oy, r.oz), v0 = mk3(a.x, a.y, a.z), e1 = mk3(a.w, bb.x, bb.y), e2 = mk3(bb.z, bb.w, cc.x);
        uint32_t prim = __float_as_uint(cc.y);
        if (TWO) {
          // the record holds OBJECT-space vertices: the test runs in world space on the operands of the one-level structure
          const InstXform& xf = t.xforms[g.w];
          world_triangle(xf.o2w, a, bb, cc, v0, e1, e2);
          const float4 o4 = *myOrg; ro = mk3(o4.x, o4.y, o4.z);
          prim += __float_as_uint(xf._pad[1]);
        }
        if (tri_intersect(ro, mk3(d4.x, d4.y, d4.z), v0, e1, e2, t.tmin, ANY ? r.tmax : 3.0e38f, tt, u, v)) {
          if (ANY) { myRes->x = 1.0f; finished = true; }
          else {
            // closest hit: r.tmax is the best t so far (initially the ray's tMax, exclusive); ties -> lowest triangle id
            const uint32_t best = __float_as_uint(myRes->w);
            if (tt < r.tmax || (tt == r.tmax && best != DDGI_MISS && prim < best)) {
              r.tmax = tt; *myRes = make_float4(tt, u, v, __uint_as_float(prim));
              if (TWO) myOrg->w = __uint_as_float(g.w);
            }
          }
        }
        if (!finished) finished = r.npend == 0u && (r.ng.y & 0xFF000000u) == 0u && r.sp == 0;
      }
      if (nodeSteps <= 1u) break;
      pend = __ballot_sync(0xFFFFFFFFu, active && !finished && r.npend != 0u);
      if ((uint32_t)__popc(pend) < triMin &&
          __ballot_sync(0xFFFFFFFFu, active && !finished && r.npend < 2u && ((r.ng.y & 0xFF000000u) != 0u || r.sp > 0)) != 0u) break;
    }
    // ---------------------------------------------------------------- results
    if (active && finished) {
      active = false;
      const float4 res = *myRes;
      const uint32_t tag = __float_as_uint(myDir->w);
      if (!ANY) {
        c.hits[tag] = res;                                           // (t, u, v, prim); (0,0,0,MISS) for a miss
        if (TWO) c.hitInst[tag] = __float_as_uint(myOrg->w);
        if (CNT && __float_as_uint(res.w) != DDGI_MISS) ++nHits;
      } else if (res.x == 0.0f) atomicOr(&c.lit[tag], 1u << __float_as_uint(res.w));      // miss shader: shadow = 0 -> lit
    }
  }
  if (CNT && t.stats) {
    for (int o = 16; o; o >>= 1) { cn += __shfl_xor_sync(0xFFFFFFFFu, cn, o); ct += __shfl_xor_sync(0xFFFFFFFFu, ct, o);
                                   nRays += __shfl_xor_sync(0xFFFFFFFFu, nRays, o); nHits += __shfl_xor_sync(0xFFFFFFFFu, nHits, o);
                                   nCand += __shfl_xor_sync(0xFFFFFFFFu, nCand, o); cs += __shfl_xor_sync(0xFFFFFFFFu, cs, o); }
    if (lane == 0) {
      if (!ANY) { atomicAdd(&t.stats[0], (unsigned long long)nRays); atomicAdd(&t.stats[1], (unsigned long long)nHits);
                  atomicAdd(&t.stats[3], (unsigned long long)cn); atomicAdd(&t.stats[4], (unsigned long long)ct); if (cs) atomicAdd(&t.stats[9], (unsigned long long)cs); }
      else { atomicAdd(&t.stats[2], (unsigned long long)nRays); atomicAdd(&t.stats[5], (unsigned long long)cn); atomicAdd(&t.stats[6], (unsigned long long)ct);
             if (nCand) atomicAdd(&t.stats[8], (unsigned long long)nCand); if (cs) atomicAdd(&t.stats[10], (unsigned long long)cs); }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_shade_prep: rchit:27-42,70-77 (hit attributes + surface) and the decision which shadow rays exist
// ------------------------------------------------------------------------------------------------------------
#define PREP_BLOCK_DEFAULT 256
// getSurfaceDataFromMat with texture taps (bindless.glsl:196-233), uv as rchit:70.  Out of line: factor-only scenes (all
// BASELINE configs) must not pay registers for the sampling code (32 registers = 8 blocks per SM).
// Returns (albedo.rgb, metallic) and writes roughness through the side buffer slot: everything travels in registers.
__device__ __noinline__ float4 prep_textured_surface(const TexCtx tex, const float4* shadeUV, float4* emisTex, const anerend_material_info_t* m,
                                                     uint32_t hp, uint32_t idx, float b1, float b2, float3 color, float* roughnessOut) {
  const float b0 = 1.0f - b1 - b2;
  const float4 ua = __ldg(shadeUV + 2 * (size_t)hp), ub = __ldg(shadeUV + 2 * (size_t)hp + 1);
  const float tu = ua.x * b0 + ua.z * b1 + ub.x * b2, tv = ua.y * b0 + ua.w * b1 + ub.y * b2;
  const Surface sf = surface_from_material(tex, m, tu, tv, mk3(color.x, color.y, color.z));
  if (emisTex && m->bindless_indices[3] != -1) emisTex[idx] = make_float4(sf.emissive.x, sf.emissive.y, sf.emissive.z, 0.0f);
  *roughnessOut = sf.roughness;
  return make_float4(sf.albedo.x, sf.albedo.y, sf.albedo.z, sf.metallic);
}
// One thread per ray; MINT = resident threads per SM the register budget must allow (2048: 32 registers — the kernel waits on
// scattered reads (shading records, light-map texels), so threads in flight count for more than a few spills).
template <int PREP_BLOCK, bool EXT = false, int MINT = 2048, bool BATCH = false, bool TWO = false>
__global__ void __launch_bounds__(PREP_BLOCK, MINT / PREP_BLOCK) k_shade_prep(const FastCtx c) {
  constexpr int PREP_WARPS = PREP_BLOCK / 32;
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
      // two-level structure: the records are per MESH triangle, the instance comes with the hit
      uint32_t rec = hp, hitInst = 0u;
      if (TWO) { hitInst = c.hitInst[idx]; const InstXform& x0 = t.xforms[hitInst]; rec = __float_as_uint(x0._pad[2]) + (hp - __float_as_uint(x0._pad[1])); }
      const float4* st = c.shadeTris + 5 * (size_t)rec;
      const float4 s0 = __ldg(st), s1 = __ldg(st + 1), s2 = __ldg(st + 2), s3 = __ldg(st + 3), s4 = __ldg(st + 4);
      const uint32_t ii = TWO ? hitInst : __float_as_uint(s4.z), material = TWO ? t.instances[hitInst].material : __float_as_uint(s4.w);
      const InstXform& xf = t.xforms[ii];
      const float b0 = 1.0f - h.y - h.z, b1 = h.y, b2 = h.z;
      pos = mk3(s0.x, s0.y, s0.z) * b0 + mk3(s0.w, s1.x, s1.y) * b1 + mk3(s1.z, s1.w, s2.x) * b2;
      pos = xform_point(xf.o2w, pos);
      normal = normalize3(mk3(s2.y, s2.z, s2.w) * b0 + mk3(s3.x, s3.y, s3.z) * b1 + mk3(s3.w, s4.x, s4.y) * b2);
      if (!EXT)
        normal = normalize3(mk3(normal.x * xf.w2o[0] + normal.y * xf.w2o[1] + normal.z * xf.w2o[2],
                                normal.x * xf.w2o[3] + normal.y * xf.w2o[4] + normal.z * xf.w2o[5],
                                normal.x * xf.w2o[6] + normal.y * xf.w2o[7] + normal.z * xf.w2o[8]));
      else {                                                          // mirror_reflections.rchit:74 + :86
        normal = normalize3(mk3(xf.o2w[0] * normal.x + xf.o2w[3] * normal.y + xf.o2w[6] * normal.z,
                                xf.o2w[1] * normal.x + xf.o2w[4] * normal.y + xf.o2w[7] * normal.z,
                                xf.o2w[2] * normal.x + xf.o2w[5] * normal.y + xf.o2w[8] * normal.z));
        const float4 ea = __ldg(&c.extRays[2 * (size_t)idx]);
        const V3 dd = pos - mk3(ea.x, ea.y, ea.z);
        c.extDist[idx] = sqrtf(dot3(dd, dd));
      }
      const anerend_material_info_t* m = &t.materials[material];
      V3 albedo;
      float metallic = m->met_rough[0], roughness = m->met_rough[1];
      bool textureRough = false;
      const bool textured = c.shadeUV != nullptr && material_textured(m);
      if (!textured && m->base_col_fac[0] + m->base_col_fac[1] + m->base_col_fac[2] > 0.1f) albedo = mk3(m->base_col_fac[0], m->base_col_fac[1], m->base_col_fac[2]);
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
        if (textured) {
          const float4 r = prep_textured_surface(t.tex, c.shadeUV, c.emisTex, m, rec, idx, h.y, h.z, make_float3(albedo.x, albedo.y, albedo.z),
                                                 reinterpret_cast<float*>(&c.surf[3 * (size_t)idx + 2]) + 3);
          albedo = mk3(r.x, r.y, r.z); metallic = r.w; textureRough = true;
        }
      }
      c.surf[3 * (size_t)idx + 0] = make_float4(pos.x, pos.y, pos.z, normal.x);
      c.surf[3 * (size_t)idx + 1] = make_float4(normal.y, normal.z, albedo.x, albedo.y);
      if (textureRough) roughness = c.surf[3 * (size_t)idx + 2].w;       // written by prep_textured_surface
      c.surf[3 * (size_t)idx + 2] = make_float4(albedo.z, __uint_as_float(material), metallic, roughness);
      c.lit[idx] = 0u;
    } else if (EXT) {
      // mirror_reflections.rmiss: (sky, 1e6); pixels without a ray are overwritten by the resolve kernel
      t.rayIrr[idx] = make_uint2((uint32_t)f2h(t.sky.x) | ((uint32_t)f2h(t.sky.y) << 16), (uint32_t)f2h(t.sky.z) | ((uint32_t)f2h(1e6f) << 16));
    } else {
      // rmiss:20 + rgen:109
      const uint32_t slot = idx % (uint32_t)t.N, probeLin = idx / (uint32_t)t.N;
      const int layerLocal = (int)(probeLin / (uint32_t)t.Pl), probe = stripe_probe((int)(probeLin % (uint32_t)t.Pl), t.nXZ, t.rowOffset, t.rowStride);
      const uint32_t ray = t.rayOrder ? (uint32_t)t.rayOrder[slot] : slot;
      const size_t o = ((size_t)t.layerMap[layerLocal] * P + probe) * t.N + ray;
      t.rayIrr[o] = make_uint2((uint32_t)f2h(t.sky.x) | ((uint32_t)f2h(t.sky.y) << 16), (uint32_t)f2h(t.sky.z) | (0x3C00u << 16));
      if (t.records) { anerend_ddgi_ray_record_t rr; rr.prim = DDGI_MISS; rr.t = 0.0f; rr.u = 0.0f; rr.v = 0.0f;
                       rr.radiance[0] = t.sky.x; rr.radiance[1] = t.sky.y; rr.radiance[2] = t.sky.z; rr.lit_mask = 0u; t.records[o] = rr; }
    }
  }
  // Which queues does this ray enter?  bit q: shadow ray towards light q (0 = sun, rchit:45-66: every hit; point
  // lights, extension E1: only where the light can contribute); bit 32: the compacted hit list k_shade runs over.
  uint32_t needMask = isHit ? 1u : 0u, preLit = 0u;
  if (isHit && (t.flags & ANEREND_DDGI_FLAG_SKIP_DEAD_SHADOW_RAYS) && !(dot3(normal, normalize3(-t.lightDir)) > 0.0f)) needMask = 0u;
  // the sun's map (ddgi_kernels.cuh): its read is issued here and looked at after the point lights' reads are on their way too
  const bool sunAsk = needMask != 0u && t.smRes != 0u && !c.rawSunDir;
  SmEntry sunE; sunE.inside = false; sunE.top = 0.0f; sunE.cand = 0u;
  if (sunAsk) sunE = sunmap_lookup(t, pos);
  uint32_t askMask = 0u;                         // BATCH: lights whose ray exists (in range and facing), looked up two at a time below
  if (isHit && (t.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS)) {
    for (uint32_t li = 0; li < t.numLights; ++li) {
      const anerend_light_t* lt = &t.lights[li];
      if (lt->color[3] > 0.5f) {
        const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
        const V3 dv = lp - pos; const float d2 = dot3(dv, dv), range = lt->world_pos[3];
        // the decisions below are the oracle's (dist < range, dot(normal, normalize(dv)) > 0); the two filters only
        // skip the sqrt / normalise where rounding cannot change the outcome (margins >> the fp32 error of either side)
        if (d2 > range * range * 1.00001f) continue;
        const float dist = sqrtf(d2);
        if (dist < range) {
          const float sgn = dot3(normal, dv), mag = fabsf(normal.x * dv.x) + fabsf(normal.y * dv.y) + fabsf(normal.z * dv.z);
          bool facing;
          if (d2 > 1e-30f && fabsf(sgn) > 1e-5f * mag) facing = sgn > 0.0f;
          else { const V3 L = normalize3(lp - pos); facing = dot3(normal, L) > 0.0f; }
          // the light map answers the rays that certainly reach the light (ddgi_kernels.cuh "Light maps"): lit without a traversal
          if (facing) {
            if (BATCH) askMask |= 1u << (1u + li);
            else if (lightmap_unoccluded(lightmap_lookup(t, li, lp, pos), dist, t.tmin)) preLit |= 1u << (1u + li); else needMask |= 1u << (1u + li);
          }
        }
      }
    }
  }
  if (BATCH) {
    // two look-ups per round so their (scattered, mostly L2-missing) reads overlap
    if (t.lmRes == 0u) needMask |= askMask;
    else for (uint32_t m = askMask; m != 0u;) {
      const uint32_t b0 = (uint32_t)__ffs((int)m) - 1u; m &= m - 1u;
      const bool two = m != 0u;
      const uint32_t b1 = two ? (uint32_t)__ffs((int)m) - 1u : b0; m &= m - 1u;
      const anerend_light_t* l0 = &t.lights[b0 - 1u]; const anerend_light_t* l1 = &t.lights[b1 - 1u];
      const V3 p0 = mk3(l0->world_pos[0], l0->world_pos[1], l0->world_pos[2]), p1 = mk3(l1->world_pos[0], l1->world_pos[1], l1->world_pos[2]);
      const LmEntry e0 = lightmap_lookup(t, b0 - 1u, p0, pos), e1 = lightmap_lookup(t, b1 - 1u, p1, pos);
      const V3 d0 = p0 - pos, d1 = p1 - pos;
      if (lightmap_unoccluded(e0, sqrtf(dot3(d0, d0)), t.tmin)) preLit |= 1u << b0; else needMask |= 1u << b0;
      if (two) { if (lightmap_unoccluded(e1, sqrtf(dot3(d1, d1)), t.tmin)) preLit |= 1u << b1; else needMask |= 1u << b1; }
    }
  }
  if (sunAsk && sunmap_unoccluded(sunE, dot3(pos, t.smS), t.tmin)) { needMask &= ~1u; preLit |= 1u; }
  if (preLit) {
    c.lit[idx] = preLit;
    if ((t.flags & ANEREND_DDGI_FLAG_COUNTERS) && t.stats) atomicAdd(&t.stats[7], (unsigned long long)__popc(preLit));
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

// Block layout: every thread looks at one ray and the MISSES are finished on the spot (sky, rmiss:20); the hits of the block are
// compacted (ray order kept) and the surface / light work runs over the compacted list, so its warps are full although only about
// half of the probe rays hit anything.
template <int PREP_BLOCK, bool EXT = false>
__global__ void __launch_bounds__(PREP_BLOCK, 1024 / PREP_BLOCK) k_shade_prep_compact(const FastCtx c) {
  constexpr int PREP_WARPS = PREP_BLOCK / 32;
  const TraceCtx& t = c.t;
  const uint32_t ray0 = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int P = t.nXZ * t.nXZ;
  __shared__ uint32_t sHitIdx[PREP_BLOCK];
  __shared__ uint32_t sCnt[33][PREP_WARPS];
  __shared__ uint32_t sBase[33];
  bool hit0 = false;
  if (ray0 < c.totalRays) {
    const uint32_t hp = __float_as_uint(c.hits[ray0].w);
    if (hp != DDGI_MISS) hit0 = true;
    else if (EXT) {
      // mirror_reflections.rmiss: (sky, 1e6); pixels without a ray are overwritten by the resolve kernel
      t.rayIrr[ray0] = make_uint2((uint32_t)f2h(t.sky.x) | ((uint32_t)f2h(t.sky.y) << 16), (uint32_t)f2h(t.sky.z) | ((uint32_t)f2h(1e6f) << 16));
    } else {
      // rmiss:20 + rgen:109
      const uint32_t slot = ray0 % (uint32_t)t.N, probeLin = ray0 / (uint32_t)t.N;
      const int layerLocal = (int)(probeLin / (uint32_t)t.Pl), probe = stripe_probe((int)(probeLin % (uint32_t)t.Pl), t.nXZ, t.rowOffset, t.rowStride);
      const uint32_t ray = t.rayOrder ? (uint32_t)t.rayOrder[slot] : slot;
      const size_t o = ((size_t)t.layerMap[layerLocal] * P + probe) * t.N + ray;
      t.rayIrr[o] = make_uint2((uint32_t)f2h(t.sky.x) | ((uint32_t)f2h(t.sky.y) << 16), (uint32_t)f2h(t.sky.z) | (0x3C00u << 16));
      if (t.records) { anerend_ddgi_ray_record_t rr; rr.prim = DDGI_MISS; rr.t = 0.0f; rr.u = 0.0f; rr.v = 0.0f;
                       rr.radiance[0] = t.sky.x; rr.radiance[1] = t.sky.y; rr.radiance[2] = t.sky.z; rr.lit_mask = 0u; t.records[o] = rr; }
    }
  }
  {   // compaction of the block's hits, ray order kept
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit0);
    if (lane == 0) sCnt[32][warp] = (uint32_t)__popc(m);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (uint32_t w = 0; w < (uint32_t)PREP_WARPS; ++w) { const uint32_t v = sCnt[32][w]; before += w < warp ? v : 0u; total += v; }
    if (hit0) sHitIdx[before + (uint32_t)__popc(m & ((1u << lane) - 1u))] = ray0;
    __syncthreads();
    if (threadIdx.x == 0) sBase[32] = total;        // read below, after the next barrier
  }
  __syncthreads();
  const uint32_t nHitBlock = sBase[32];
  const bool isHit = threadIdx.x < nHitBlock;
  const uint32_t idx = isHit ? sHitIdx[threadIdx.x] : 0u;
  V3 pos = mk3(0.0f), normal = mk3(0.0f);
  if (isHit) {
    const float4 h = c.hits[idx];
    const uint32_t hp = __float_as_uint(h.w);
    {
      // rchit:27-38 through the per-triangle record (five independent 16-byte loads instead of the index/vertex gather)
      const float4* st = c.shadeTris + 5 * (size_t)hp;
      const float4 s0 = __ldg(st), s1 = __ldg(st + 1), s2 = __ldg(st + 2), s3 = __ldg(st + 3), s4 = __ldg(st + 4);
      const uint32_t ii = __float_as_uint(s4.z), material = __float_as_uint(s4.w);
      const InstXform& xf = t.xforms[ii];
      const float b0 = 1.0f - h.y - h.z, b1 = h.y, b2 = h.z;
      pos = mk3(s0.x, s0.y, s0.z) * b0 + mk3(s0.w, s1.x, s1.y) * b1 + mk3(s1.z, s1.w, s2.x) * b2;
      pos = xform_point(xf.o2w, pos);
      normal = normalize3(mk3(s2.y, s2.z, s2.w) * b0 + mk3(s3.x, s3.y, s3.z) * b1 + mk3(s3.w, s4.x, s4.y) * b2);
      if (!EXT)
        normal = normalize3(mk3(normal.x * xf.w2o[0] + normal.y * xf.w2o[1] + normal.z * xf.w2o[2],
                                normal.x * xf.w2o[3] + normal.y * xf.w2o[4] + normal.z * xf.w2o[5],
                                normal.x * xf.w2o[6] + normal.y * xf.w2o[7] + normal.z * xf.w2o[8]));
      else {                                                          // mirror_reflections.rchit:74 + :86
        normal = normalize3(mk3(xf.o2w[0] * normal.x + xf.o2w[3] * normal.y + xf.o2w[6] * normal.z,
                                xf.o2w[1] * normal.x + xf.o2w[4] * normal.y + xf.o2w[7] * normal.z,
                                xf.o2w[2] * normal.x + xf.o2w[5] * normal.y + xf.o2w[8] * normal.z));
        const float4 ea = __ldg(&c.extRays[2 * (size_t)idx]);
        const V3 dd = pos - mk3(ea.x, ea.y, ea.z);
        c.extDist[idx] = sqrtf(dot3(dd, dd));
      }
      const anerend_material_info_t* m = &t.materials[material];
      V3 albedo;
      float metallic = m->met_rough[0], roughness = m->met_rough[1];
      bool textureRough = false;
      const bool textured = c.shadeUV != nullptr && material_textured(m);
      if (!textured && m->base_col_fac[0] + m->base_col_fac[1] + m->base_col_fac[2] > 0.1f) albedo = mk3(m->base_col_fac[0], m->base_col_fac[1], m->base_col_fac[2]);
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
        if (textured) {
          const float4 r = prep_textured_surface(t.tex, c.shadeUV, c.emisTex, m, hp, idx, h.y, h.z, make_float3(albedo.x, albedo.y, albedo.z),
                                                 reinterpret_cast<float*>(&c.surf[3 * (size_t)idx + 2]) + 3);
          albedo = mk3(r.x, r.y, r.z); metallic = r.w; textureRough = true;
        }
      }
      c.surf[3 * (size_t)idx + 0] = make_float4(pos.x, pos.y, pos.z, normal.x);
      c.surf[3 * (size_t)idx + 1] = make_float4(normal.y, normal.z, albedo.x, albedo.y);
      if (textureRough) roughness = c.surf[3 * (size_t)idx + 2].w;       // written by prep_textured_surface
      c.surf[3 * (size_t)idx + 2] = make_float4(albedo.z, __uint_as_float(material), metallic, roughness);
    }
  }
  // Which queues does this ray enter?  bit q: shadow ray towards light q (0 = sun, rchit:45-66: every hit; point
  // lights, extension E1: only where the light can contribute); bit 32: the compacted hit list k_shade runs over.
  uint32_t needMask = isHit ? 1u : 0u, preLit = 0u;
  if (isHit && (t.flags & ANEREND_DDGI_FLAG_SKIP_DEAD_SHADOW_RAYS) && !(dot3(normal, normalize3(-t.lightDir)) > 0.0f)) needMask = 0u;
  // the sun's map (ddgi_kernels.cuh): its read is issued here and looked at after the point lights' reads are on their way too
  const bool sunAsk = needMask != 0u && t.smRes != 0u && !c.rawSunDir;
  SmEntry sunE; sunE.inside = false; sunE.top = 0.0f; sunE.cand = 0u;
  if (sunAsk) sunE = sunmap_lookup(t, pos);
  if (isHit && (t.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS)) {
    uint32_t askMask = 0u;                         // lights whose ray exists: in range and facing
    for (uint32_t li = 0; li < t.numLights; ++li) {
      const anerend_light_t* lt = &t.lights[li];
      if (lt->color[3] > 0.5f) {
        const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
        const V3 dv = lp - pos; const float d2 = dot3(dv, dv), range = lt->world_pos[3];
        // the decisions below are the oracle's (dist < range, dot(normal, normalize(dv)) > 0); the two filters only
        // skip the sqrt / normalise where rounding cannot change the outcome (margins >> the fp32 error of either side)
        if (d2 > range * range * 1.00001f) continue;
        const float dist = sqrtf(d2);
        if (dist < range) {
          const float sgn = dot3(normal, dv), mag = fabsf(normal.x * dv.x) + fabsf(normal.y * dv.y) + fabsf(normal.z * dv.z);
          bool facing;
          if (d2 > 1e-30f && fabsf(sgn) > 1e-5f * mag) facing = sgn > 0.0f;
          else { const V3 L = normalize3(lp - pos); facing = dot3(normal, L) > 0.0f; }
          if (facing) askMask |= 1u << (1u + li);
        }
      }
    }
    // the light maps answer the rays that certainly reach their light (ddgi_kernels.cuh "Light maps"): lit without a traversal.
    // Two look-ups per round so their (scattered, mostly L2-missing) reads overlap.
    if (t.lmRes == 0u) needMask |= askMask;
    else for (uint32_t m = askMask; m != 0u;) {
      const uint32_t b0 = (uint32_t)__ffs((int)m) - 1u; m &= m - 1u;
      const bool two = m != 0u;
      const uint32_t b1 = two ? (uint32_t)__ffs((int)m) - 1u : b0; m &= m - 1u;
      const anerend_light_t* l0 = &t.lights[b0 - 1u]; const anerend_light_t* l1 = &t.lights[b1 - 1u];
      const V3 p0 = mk3(l0->world_pos[0], l0->world_pos[1], l0->world_pos[2]), p1 = mk3(l1->world_pos[0], l1->world_pos[1], l1->world_pos[2]);
      const LmEntry e0 = lightmap_lookup(t, b0 - 1u, p0, pos), e1 = lightmap_lookup(t, b1 - 1u, p1, pos);
      const V3 d0 = p0 - pos, d1 = p1 - pos;
      if (lightmap_unoccluded(e0, sqrtf(dot3(d0, d0)), t.tmin)) preLit |= 1u << b0; else needMask |= 1u << b0;
      if (two) { if (lightmap_unoccluded(e1, sqrtf(dot3(d1, d1)), t.tmin)) preLit |= 1u << b1; else needMask |= 1u << b1; }
    }
  }
  if (sunAsk && sunmap_unoccluded(sunE, dot3(pos, t.smS), t.tmin)) { needMask &= ~1u; preLit |= 1u; }
  if (isHit) {
    c.lit[idx] = preLit;
    if (preLit && (t.flags & ANEREND_DDGI_FLAG_COUNTERS) && t.stats) atomicAdd(&t.stats[7], (unsigned long long)__popc(preLit));
  }
  // block-aggregated append: one atomic per queue per block (a per-warp atomic on nine neighbouring counters was the
  // whole cost of this kernel), order inside the block = ray order
  const uint32_t nq = c.numQueues;
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
// k_hit_bucket<PASS>: the hit list grouped by the number of unoccluded lights of the hit.  k_shade walks each lane's OWN lit lights
// (the oracle's accumulation order), so a warp runs as many rounds as its busiest lane: ~5 against a mean of 2.1 lights per hit with
// the list in ray order (profiles/r02: the light loop is 45 % of k_shade's instructions at 13.8 of 32 lanes).  With hits of equal
// light count next to each other the rounds of a warp are all full.  PASS 0 counts the classes, PASS 1 scatters (order inside a class:
// blocks as they arrive, ray order inside a block — only locality depends on it).
// ------------------------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(256) k_hit_bucket(const FastCtx c) {
  __shared__ uint32_t sCnt[33], sBase[33];
  const uint32_t nHits = c.qcount[FAST_QC_HITS];
  if (blockIdx.x * 256u >= nHits) return;
  if (threadIdx.x < 33u) sCnt[threadIdx.x] = 0u;
  __syncthreads();
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  uint32_t idx = 0u, key = 0u, rank = 0u;
  const bool valid = i < nHits;
  if (valid) { idx = c.hitq[i]; key = (uint32_t)__popc(c.lit[idx]); rank = atomicAdd(&sCnt[key], 1u); }
  __syncthreads();
  uint32_t* bucket = c.qcount + FAST_QC_BUCKET;
  if (PASS == 0) { if (threadIdx.x < 33u && sCnt[threadIdx.x]) atomicAdd(&bucket[threadIdx.x], sCnt[threadIdx.x]); return; }
  if (threadIdx.x < 33u) {
    // classes with MORE lights first: the long warps start early
    uint32_t before = 0u;
    for (uint32_t k = 32u; k > threadIdx.x; --k) before += bucket[k];
    sBase[threadIdx.x] = before + (sCnt[threadIdx.x] ? atomicAdd(&bucket[64u + threadIdx.x], sCnt[threadIdx.x]) : 0u);
  }
  __syncthreads();
  if (valid) c.hitqSorted[sBase[key] + rank] = idx;
}

// ------------------------------------------------------------------------------------------------------------
// ShadowRT on the wavefront kernels: k_shadowrt_stage (world positions + the ray queue, pixel order) -> k_trace<true> ->
// k_shadowrt_resolve
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_shadowrt_stage(const ShadowRtCtx c, float4* pos, uint32_t* lit, uint32_t* queue, uint32_t* qcount) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = (uint32_t)c.width * (uint32_t)c.height, lane = threadIdx.x & 31u;
  bool need = false;
  if (i < n) {
    const float depthSamp = c.depth[i];
    lit[i] = 0u;
    if (depthSamp != 1.0f) {
      const V3 p = shadow_rt_world_pos(c, (int)(i % (uint32_t)c.width), (int)(i / (uint32_t)c.width), depthSamp);
      pos[i] = make_float4(p.x, p.y, p.z, 0.0f);
      need = true;
    }
  }
  __shared__ uint32_t sCnt[8], sBase;
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, need), warp = threadIdx.x >> 5;
  if (lane == 0) sCnt[warp] = (uint32_t)__popc(m);
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t sum = 0; for (int w = 0; w < 8; ++w) { const uint32_t v = sCnt[w]; sCnt[w] = sum; sum += v; } sBase = sum ? atomicAdd(qcount, sum) : 0u; }
  __syncthreads();
  if (need) queue[sBase + sCnt[warp] + (uint32_t)__popc(m & ((1u << lane) - 1u))] = i;
}
__global__ void __launch_bounds__(256) k_shadowrt_resolve(const float* depth, const uint32_t* lit, float* out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (depth[i] == 1.0f || (lit[i] & 1u)) ? 1.0f : 0.0f;
}

// SpecularGIRT on the wavefront kernels: k_specular_stage (one reflected ray per pixel) -> k_trace<false, false, true> ->
// k_shade_prep<., true> -> k_trace<true> -> k_shade<., true> -> k_specular_resolve (pixels of the sky)
__global__ void __launch_bounds__(256) k_specular_stage(const SpecularCtx c, float4* extRays) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = (uint32_t)c.width * (uint32_t)c.height;
  if (i >= n) return;
  const float depthSamp = c.depth[i];
  if (depthSamp == 1.0f) { extRays[2 * (size_t)i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); extRays[2 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 1.0f, 0.0f); return; }
  ShadowRtCtx w{};
  for (int k = 0; k < 16; ++k) w.invViewProj[k] = c.invViewProj[k];
  w.screenW = c.screenW; w.screenH = c.screenH;
  const V3 pos = shadow_rt_world_pos(w, (int)(i % (uint32_t)c.width), (int)(i / (uint32_t)c.width), depthSamp);
  const float4 g0 = c.geom0[i];
  const V3 normal = normalize3(mk3(g0.x, g0.y, g0.z));
  const V3 vv = normalize3(pos - c.cameraPos);
  const float dn = dot3(normal, vv);
  const V3 dir = vv - (2.0f * dn) * normal;
  extRays[2 * (size_t)i] = make_float4(pos.x, pos.y, pos.z, 50.0f);
  extRays[2 * (size_t)i + 1] = make_float4(dir.x, dir.y, dir.z, 0.0f);
}
__global__ void __launch_bounds__(256) k_specular_resolve(const float* depth, uint2* out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && depth[i] == 1.0f) out[i] = make_uint2(0u, 0u);
}

}  // namespace ddgi
