// ddgi_kernels.cuh — per-thread bodies of every kernel on the DDGI probe-update path.
//
// Each `*_body(i, ctx)` is the work of ONE thread of the named kernel; ddgi_launch.h wraps them in
// __global__ kernels (product) or serial loops (tests/host_emu).  Reference citations are relative to the
// AneRend repository (shaders in assets/shaders/).  This translation unit is compiled with --fmad=false:
// arithmetic is evaluated left to right without contraction, the only fused multiply-adds are the explicit
// fmaf calls (triangle test, quantised box test), so hit ids match the CPU oracle bit for bit.
#pragma once
#include "platform.h"
#include "../../include/anerend_cuda_ddgi.h"

namespace ddgi {

// ----------------------------------------------------------------------------------------------------
// small vector maths
// ----------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
HD V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
HD V3 mk3(float s) { return mk3(s, s, s); }
HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
HD V3 operator/(V3 a, V3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
HD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HD V3 normalize3(V3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return v * inv; }
// exact variants (see platform.h xmul..): same values as dot3/normalize3 under --fmad=false, immune to fast-math flags
HD float dot3x(V3 a, V3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
HD V3 normalize3x(V3 v) { float inv = xdiv(1.0f, xsqrt(dot3x(v, v))); return mk3(xmul(v.x, inv), xmul(v.y, inv), xmul(v.z, inv)); }
HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
HD V3 clamp3(V3 v, float lo, float hi) { return mk3(clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)); }
HD V3 min3(V3 a, V3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
HD V3 max3(V3 a, V3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
HD float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
HD V3 mix3(V3 x, V3 y, float a) { return mk3(mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)); }
HD float pow5(float x) { float x2 = x * x; return (x2 * x2) * x; }
HD int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

#define DDGI_PI 3.14159265359f
#define DDGI_MISS 0xFFFFFFFFu
#define PLOC_RADIUS 16            // search window of the PLOC builder (positions either side in Morton order)

// ----------------------------------------------------------------------------------------------------
// device-side scene representation
// ----------------------------------------------------------------------------------------------------
struct DInstance {            // one (renderable, mesh) pair accepted by tlas_update.comp:37-99
  uint32_t renderable, mesh, material, firstTri, numTris, vertexOffset, indexOffset, _pad;
};
struct InstXform {            // gl_ObjectToWorldEXT (4 columns x 3 rows) and the 3x3 of gl_WorldToObjectEXT
  float o2w[12];              // o2w[c*3+r]
  float w2o[9];               // w2o[c*3+r]
  float _pad[3];
};
struct CItem { uint32_t ref, node; };   // BVH8 collapse work item: binary subtree `ref` becomes BVH8 node `node`

struct BuildCtx {
  const anerend_vertex_t* vertices; const uint32_t* indices; const anerend_renderable_t* renderables;
  const DInstance* instances; InstXform* xforms; const uint32_t* triInstance;
  uint32_t numTris, numInstances;
  float4* wtri;               // 3 x float4 per world triangle, global-id order: (v0,e1.x) (e1.yz,e2.xy) (e2.z,id,0,0)
  float4* triLo; float4* triHi;   // padded AABB per triangle
  int* sceneBounds;           // 6 order-preserving ints: min xyz, max xyz
  uint32_t* mortonKeys; uint32_t* triIds;          // unsorted
  uint32_t* sortedKeys; uint32_t* sortedIds;       // after the radix sort
  uint32_t* left; uint32_t* right; uint32_t* parent; uint32_t* rangeFirst; uint32_t* rangeLast; uint32_t* visit;
  float4* blo; float4* bhi;   // boxes of the 2n-1 binary refs (ref >= n-1: leaf at sorted position ref-(n-1))
  uint4* nodes; float4* bvhTris; uint32_t* nodeParent; float4* nodeLo; float4* nodeHi; uint32_t* nodePending;
  uint32_t* counters;         // [0] BVH8 nodes allocated, [1] triangle slots allocated, [2] next queue length
  const CItem* qin; CItem* qout;
  float4* shadeTris;          // [numTris][5] object-space shading record (ddgi_wavefront.h ShadeTri) or null
  float4* shadeUV;            // [numTris][2] (uv0, uv1) (uv2, -, -): only when some material is textured
  // PLOC (parallel locally-ordered clustering) state: clusters in Morton order, double-buffered across the compaction
  uint32_t* clRef; float4* clLo; float4* clHi;              // current clusters: binary ref + box
  uint32_t* clRefOut; float4* clLoOut; float4* clHiOut;     // compacted clusters of the next iteration
  uint32_t* clNN; uint32_t* clKeep; const uint32_t* clPos;  // nearest neighbour, survives-this-round flag, its exclusive scan
  uint32_t* subCount;         // triangles below each of the 2n-1 binary refs
  uint32_t* leafMap; uint32_t* orderedIds;                  // Morton position -> depth-first position; triangle ids in that order
  uint32_t numClusters, plocRadius;      // numClusters: upper bound of the launch; the live count is numTris - counters[3] (read on the device)
  uint32_t* triInstanceOut;              // k_tri_instance: triangle -> index of the instance whose range [firstTri, firstTri + numTris) holds it
  uint32_t level;                        // BVH8 collapse: queue length of this level = counters[CNT_LEVEL0 + level], the next level's is counted beside it
  uint32_t leafMax;                      // primitives per leaf slot (0 = 3, triangles; 1 for the instance level of a two-level structure)
  // refit of moving nodes: which instances' transforms changed since the last build / refit (k_inst_xform compares); null = all
  uint32_t* instChanged;
};
// counters[]: [0] BVH8 nodes allocated, [1] triangle slots allocated, [2] -, [3] PLOC merges so far, [4] clusters of the current PLOC round,
// [CNT_LEVEL0 + L] work items of collapse level L.  Everything the build loops depend on stays on the device: the host launches
// whole batches of rounds / levels sized by an upper bound and reads a counter back once per batch.
#define CNT_ROUND_M 4
#define CNT_LEVEL0 8
#define CNT_LEVELS 64
#define CNT_SIZE (CNT_LEVEL0 + CNT_LEVELS + 1)

// ----------------------------------------------------------------------------------------------------
// instance transforms + flatten (replaces tlas_update.comp's instance write and the driver BLAS/TLAS build inputs)
// ----------------------------------------------------------------------------------------------------
HD void inst_xform_body(uint32_t i, const BuildCtx& c) {
  const float* m = c.renderables[c.instances[i].renderable].transform;   // column-major mat4
  InstXform x;
  for (int col = 0; col < 4; ++col) for (int r = 0; r < 3; ++r) x.o2w[col * 3 + r] = m[col * 4 + r];
  // inverse of the upper 3x3: adjugate / determinant
  float a = m[0], b = m[4], cc = m[8], d = m[1], e = m[5], f = m[9], g = m[2], h = m[6], k = m[10];
  float A = e * k - f * h, B = -(d * k - f * g), C = d * h - e * g;
  float det = a * A + b * B + cc * C;
  float inv = 1.0f / det;
  float r00 = A * inv, r01 = -(b * k - cc * h) * inv, r02 = (b * f - cc * e) * inv;
  float r10 = B * inv, r11 = (a * k - cc * g) * inv, r12 = -(a * f - cc * d) * inv;
  float r20 = C * inv, r21 = -(a * h - b * g) * inv, r22 = (a * e - b * d) * inv;
  x.w2o[0] = r00; x.w2o[1] = r10; x.w2o[2] = r20;
  x.w2o[3] = r01; x.w2o[4] = r11; x.w2o[5] = r21;
  x.w2o[6] = r02; x.w2o[7] = r12; x.w2o[8] = r22;
  x._pad[0] = x._pad[1] = x._pad[2] = 0.0f;
  if (c.instChanged) {
    const InstXform& old = c.xforms[i];
    bool same = true;
    for (int k = 0; k < 12; ++k) same = same && f2u(old.o2w[k]) == f2u(x.o2w[k]);
    c.instChanged[i] = same ? 0u : 1u;
  }
  c.xforms[i] = x;
}

// gl_ObjectToWorldEXT * vec4(p,1): c0*x + c1*y + c2*z + c3, left to right
HD V3 xform_point(const float* o2w, V3 p) {
  return mk3(o2w[0] * p.x + o2w[3] * p.y + o2w[6] * p.z + o2w[9],
             o2w[1] * p.x + o2w[4] * p.y + o2w[7] * p.z + o2w[10],
             o2w[2] * p.x + o2w[5] * p.y + o2w[8] * p.z + o2w[11]);
}

// instances are laid out in order of firstTri (selectInstances): upper-bound search, skipping instances without triangles
HD void tri_instance_body(uint32_t t, const BuildCtx& c) {
  uint32_t lo = 0, hi = c.numInstances;            // last instance with firstTri <= t
  while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (c.instances[mid].firstTri <= t) lo = mid; else hi = mid; }
  c.triInstanceOut[t] = lo;
}
HD int float_to_ordered(float f) { int i = (int)f2u(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
HD float ordered_to_float(int i) { return u2f((uint32_t)(i >= 0 ? i : i ^ 0x7FFFFFFF)); }

HD void flatten_tri_body(uint32_t t, const BuildCtx& c) {
  uint32_t ii = c.triInstance[t];
  if (c.instChanged && !c.instChanged[ii]) return;      // refit: the triangle is where it was
  const DInstance in = c.instances[ii];
  const float* o2w = c.xforms[ii].o2w;
  uint32_t prim = t - in.firstTri;
  V3 w[3];
  float sr[18];
  for (int k = 0; k < 3; ++k) {
    uint32_t vi = in.vertexOffset + c.indices[in.indexOffset + 3 * prim + k];
    const float* p = c.vertices[vi].pos; const float* nn = c.vertices[vi].normal;
    w[k] = xform_point(o2w, mk3(p[0], p[1], p[2]));
    sr[3 * k] = p[0]; sr[3 * k + 1] = p[1]; sr[3 * k + 2] = p[2];
    sr[9 + 3 * k] = nn[0]; sr[9 + 3 * k + 1] = nn[1]; sr[9 + 3 * k + 2] = nn[2];
  }
  if (c.shadeTris) {     // rchit:27-38 inputs, gathered once per build instead of once per hit
    float4* o = c.shadeTris + 5 * (size_t)t;
    o[0] = make_float4(sr[0], sr[1], sr[2], sr[3]); o[1] = make_float4(sr[4], sr[5], sr[6], sr[7]);
    o[2] = make_float4(sr[8], sr[9], sr[10], sr[11]); o[3] = make_float4(sr[12], sr[13], sr[14], sr[15]);
    o[4] = make_float4(sr[16], sr[17], u2f(ii), u2f(in.material));
  }
  if (c.shadeUV) {
    float uv[6];
    for (int k = 0; k < 3; ++k) { const float* q = c.vertices[in.vertexOffset + c.indices[in.indexOffset + 3 * prim + k]].uv; uv[2 * k] = q[0]; uv[2 * k + 1] = q[1]; }
    c.shadeUV[2 * (size_t)t] = make_float4(uv[0], uv[1], uv[2], uv[3]);
    c.shadeUV[2 * (size_t)t + 1] = make_float4(uv[4], uv[5], 0.0f, 0.0f);
  }
  V3 e1 = w[1] - w[0], e2 = w[2] - w[0];
  c.wtri[3 * (size_t)t + 0] = make_float4(w[0].x, w[0].y, w[0].z, e1.x);
  c.wtri[3 * (size_t)t + 1] = make_float4(e1.y, e1.z, e2.x, e2.y);
  c.wtri[3 * (size_t)t + 2] = make_float4(e2.z, u2f(t), 0.0f, 0.0f);
  // AABB of what the triangle test actually sees (v0, v0+e1, v0+e2), padded so that neither the e1/e2 round trip
  // nor slab-test rounding can cull a true hit
  V3 q1 = w[0] + e1, q2 = w[0] + e2;
  V3 lo = min3(min3(w[0], min3(w[1], w[2])), min3(q1, q2)), hi = max3(max3(w[0], max3(w[1], w[2])), max3(q1, q2));
  V3 pad = mk3(1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.x), fabsf(hi.x))), 1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.y), fabsf(hi.y))),
               1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.z), fabsf(hi.z))));
  lo = lo - pad; hi = hi + pad;
  c.triLo[t] = make_float4(lo.x, lo.y, lo.z, 0.0f);
  c.triHi[t] = make_float4(hi.x, hi.y, hi.z, 0.0f);
  if (c.sceneBounds) {
    atomMin(&c.sceneBounds[0], float_to_ordered(lo.x)); atomMin(&c.sceneBounds[1], float_to_ordered(lo.y)); atomMin(&c.sceneBounds[2], float_to_ordered(lo.z));
    atomMax(&c.sceneBounds[3], float_to_ordered(hi.x)); atomMax(&c.sceneBounds[4], float_to_ordered(hi.y)); atomMax(&c.sceneBounds[5], float_to_ordered(hi.z));
  }
}

// ----------------------------------------------------------------------------------------------------
// LBVH: 30-bit Morton codes -> (radix sort) -> Karras 2012 hierarchy -> bottom-up AABB fit
// ----------------------------------------------------------------------------------------------------
HD uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v;
}
HD void morton_body(uint32_t t, const BuildCtx& c) {
  V3 lo = mk3(ordered_to_float(c.sceneBounds[0]), ordered_to_float(c.sceneBounds[1]), ordered_to_float(c.sceneBounds[2]));
  V3 hi = mk3(ordered_to_float(c.sceneBounds[3]), ordered_to_float(c.sceneBounds[4]), ordered_to_float(c.sceneBounds[5]));
  float4 a = c.triLo[t], b = c.triHi[t];
  V3 cen = mk3(0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z));
  V3 ext = max3(hi - lo, mk3(1e-20f));
  V3 u = (cen - lo) / ext;
  uint32_t x = (uint32_t)fminf(fmaxf(u.x * 1024.0f, 0.0f), 1023.0f);
  uint32_t y = (uint32_t)fminf(fmaxf(u.y * 1024.0f, 0.0f), 1023.0f);
  uint32_t z = (uint32_t)fminf(fmaxf(u.z * 1024.0f, 0.0f), 1023.0f);
  c.mortonKeys[t] = (expand_bits10(x) << 2) | (expand_bits10(y) << 1) | expand_bits10(z);
  c.triIds[t] = t;
}

// common-prefix length of sorted keys i and j (duplicates broken by index); -1 outside the array
HD int lbvh_delta(const uint32_t* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint32_t a = keys[i], b = keys[j];
  if (a == b) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
  return clz32(a ^ b);
}
HD void karras_body(uint32_t idx, const BuildCtx& c) {
  const int n = (int)c.numTris, i = (int)idx;
  const uint32_t* keys = c.sortedKeys;
  int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = lbvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2) if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = lbvh_delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
    if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + (d < 0 ? -1 : 0);
  int first = i < j ? i : j, last = i < j ? j : i;
  uint32_t L = (first == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
  uint32_t R = (last == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
  c.left[i] = L; c.right[i] = R; c.parent[L] = (uint32_t)i; c.parent[R] = (uint32_t)i;
  c.rangeFirst[i] = (uint32_t)first; c.rangeLast[i] = (uint32_t)last;
  if (i == 0) c.parent[0] = DDGI_MISS;
}
HD void fit_body(uint32_t leaf, const BuildCtx& c) {
  const uint32_t n = c.numTris;
  uint32_t ref = n - 1 + leaf, gid = c.sortedIds[leaf];
  float4 lo = c.triLo[gid], hi = c.triHi[gid];
  c.blo[ref] = lo; c.bhi[ref] = hi;
  if (n == 1) return;
  memFence();
  uint32_t p = c.parent[ref];
  while (p != DDGI_MISS) {
    if (atomAdd(&c.visit[p], 1u) == 0u) return;     // first child to arrive leaves; the second one merges
    memFence();
    uint32_t l = c.left[p], r = c.right[p];
#ifdef __CUDA_ARCH__
    float4 a0 = __ldcg(&c.blo[l]), a1 = __ldcg(&c.bhi[l]), b0 = __ldcg(&c.blo[r]), b1 = __ldcg(&c.bhi[r]);
#else
    float4 a0 = c.blo[l], a1 = c.bhi[l], b0 = c.blo[r], b1 = c.bhi[r];
#endif
    c.blo[p] = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), 0.0f);
    c.bhi[p] = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), 0.0f);
    memFence();
    p = c.parent[p];
  }
}

// ----------------------------------------------------------------------------------------------------
// PLOC (Meister & Bittner 2018): agglomerative clustering restricted to a window of the Morton order.  Every round each
// cluster finds the neighbour within `plocRadius` positions whose union has the smallest surface area; mutual pairs
// merge into a new binary node; survivors are compacted (order kept) for the next round.  Produces the same arrays as
// the Karras hierarchy (left/right/parent/boxes/ranges) with a markedly lower SAH cost.  Internal refs are handed out
// downwards from n-2, so the last merge — the root — is ref 0, as the collapse expects.
// ----------------------------------------------------------------------------------------------------
HD void ploc_init_body(uint32_t i, const BuildCtx& c) {
  uint32_t gid = c.sortedIds[i], ref = c.numTris - 1u + i;
  c.clRef[i] = ref; c.clLo[i] = c.triLo[gid]; c.clHi[i] = c.triHi[gid];
  c.subCount[ref] = 1u;
}
HD float ploc_union_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
  float ex = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), ey = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), ez = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
  return ex * ey + ey * ez + ez * ex;
}
HD void ploc_nn_body(uint32_t i, const BuildCtx& c) {
  const int m = (int)(c.numTris - c.counters[3]), r = (int)c.plocRadius, ii = (int)i;      // counters[3] is at rest while this kernel runs
  if (i == 0u) c.counters[CNT_ROUND_M] = (uint32_t)m;                                       // ... and is kept for the merge / compaction of the round
  if (ii >= m) return;
  const float4 lo = c.clLo[i], hi = c.clHi[i];
  int b = ii - r < 0 ? 0 : ii - r, e = ii + r > m - 1 ? m - 1 : ii + r;
  float best = 3.0e38f; int bj = ii;
  for (int j = b; j <= e; ++j) {             // ascending j + strict '<': ties go to the lower index, which guarantees a mutual pair
    if (j == ii) continue;
    float a = ploc_union_area(lo, hi, c.clLo[j], c.clHi[j]);
    if (a < best) { best = a; bj = j; }
  }
  c.clNN[i] = (uint32_t)bj;
}
HD void ploc_merge_body(uint32_t i, const BuildCtx& c) {
  if (i >= c.counters[CNT_ROUND_M]) { c.clKeep[i] = 0u; return; }      // beyond the live clusters: nothing for the scan to count
  uint32_t j = c.clNN[i];
  if (j == i || c.clNN[j] != i) { c.clKeep[i] = 1u; return; }
  if (j < i) { c.clKeep[i] = 0u; return; }   // absorbed by the lower partner
  uint32_t a = c.clRef[i], b = c.clRef[j];
  uint32_t id = c.numTris - 2u - atomAdd(&c.counters[3], 1u);
  float4 alo = c.clLo[i], ahi = c.clHi[i], blo = c.clLo[j], bhi = c.clHi[j];
  float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
  float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
  c.left[id] = a; c.right[id] = b; c.parent[a] = id; c.parent[b] = id;
  c.blo[id] = lo; c.bhi[id] = hi; c.subCount[id] = c.subCount[a] + c.subCount[b];
  c.clRef[i] = id; c.clLo[i] = lo; c.clHi[i] = hi; c.clKeep[i] = 1u;
}
HD void ploc_compact_body(uint32_t i, const BuildCtx& c) {
  if (i >= c.counters[CNT_ROUND_M] || !c.clKeep[i]) return;
  uint32_t o = c.clPos[i];
  c.clRefOut[o] = c.clRef[i]; c.clLoOut[o] = c.clLo[i]; c.clHiOut[o] = c.clHi[i];
}
// depth-first position of every ref = triangles in the left siblings of its ancestors; leaves are renumbered in that
// order so that every subtree owns a contiguous triangle range (what the collapse and the leaf copies rely on)
HD void ploc_order_body(uint32_t ref, const BuildCtx& c) {
  const uint32_t n = c.numTris;
  uint32_t off = 0, cur = ref, p = c.parent[cur];
  while (p != DDGI_MISS) {
    if (c.right[p] == cur) off += c.subCount[c.left[p]];
    cur = p; p = c.parent[cur];
  }
  if (ref < n - 1u) { c.rangeFirst[ref] = off; c.rangeLast[ref] = off + c.subCount[ref] - 1u; return; }
  uint32_t s = ref - (n - 1u), gid = c.sortedIds[s];
  c.leafMap[s] = off; c.orderedIds[off] = gid;
  c.blo[n - 1u + off] = c.triLo[gid]; c.bhi[n - 1u + off] = c.triHi[gid];
}
HD void ploc_relabel_body(uint32_t id, const BuildCtx& c) {
  const uint32_t n = c.numTris;
  uint32_t l = c.left[id], r = c.right[id];
  if (l >= n - 1u) c.left[id] = n - 1u + c.leafMap[l - (n - 1u)];
  if (r >= n - 1u) c.right[id] = n - 1u + c.leafMap[r - (n - 1u)];
}

// ----------------------------------------------------------------------------------------------------
// BVH8: 80-byte compressed wide node (layout after Ylitie, Karras, Laine 2017)
//   word0: p.x p.y p.z | ex,ey,ez,imask      p = node box min, e* = biased fp32 exponents of the per-axis grid
//   word1: childBase, triBase, triMask, 0
//          imask bit s: slot s is an internal child, stored at childBase + popc(imask below s)
//          triMask bits 3s..3s+2: slot s is a leaf with 1..3 triangles (unary); triangle (s,j) is stored at
//          triBase + popc(triMask below bit 3s+j).  A slot with neither is empty (planes lo=255 > hi=0 never pass).
//          One compare per child then yields both masks: hit ? (7 << 3s | 1 << 24+s) : 0, AND-ed with triMask / imask.
//   word2: qlo.x[0..7] qlo.y[0..7]   word3: qlo.z[0..7] qhi.x[0..7]   word4: qhi.y[0..7] qhi.z[0..7]
// Child boxes are quantised outwards to 8 bits on the node's grid, so they stay conservative.
// ----------------------------------------------------------------------------------------------------
HD uint32_t grid_exponent(float ext) {
  float s = ext / 254.0f;                    // 254, not 255: margin so that ceil() never needs 256
  uint32_t bits = f2u(s), eb = (bits >> 23) & 0xFFu;
  if (bits & 0x7FFFFFu) eb += 1;
  if (eb < 1u) eb = 1u;
  if (eb > 253u) eb = 253u;
  return eb;
}
HD void encode_node(uint4* out, V3 nlo, V3 nhi, const V3* clo, const V3* chi, uint32_t imask, uint32_t triMask,
                    uint32_t childBase, uint32_t triBase) {
  V3 ext = nhi - nlo;
  uint32_t ex = grid_exponent(ext.x), ey = grid_exponent(ext.y), ez = grid_exponent(ext.z);
  float ix = u2f((254u - ex) << 23), iy = u2f((254u - ey) << 23), iz = u2f((254u - ez) << 23);
  uint32_t q[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // lox lo/hi, loy lo/hi, loz lo/hi, hix.., hiy.., hiz..
  for (int s = 0; s < 8; ++s) {
    uint32_t sh = 8u * (uint32_t)(s & 3); int w = s >> 2;
    if (!((imask >> s) & 1u) && !((triMask >> (3 * s)) & 7u)) {       // empty slot: lo = 255 > hi = 0
      q[0 + w] |= 255u << sh; q[2 + w] |= 255u << sh; q[4 + w] |= 255u << sh;
      continue;
    }
    uint32_t lx = (uint32_t)clampf(floorf((clo[s].x - nlo.x) * ix - 0.02f), 0.0f, 255.0f);   // 0.02 grid steps of slack:
    uint32_t ly = (uint32_t)clampf(floorf((clo[s].y - nlo.y) * iy - 0.02f), 0.0f, 255.0f);   // the PRMT-based decode of the
    uint32_t lz = (uint32_t)clampf(floorf((clo[s].z - nlo.z) * iz - 0.02f), 0.0f, 255.0f);   // fast kernels is exact to ~1/512 step
    uint32_t hx = (uint32_t)clampf(ceilf((chi[s].x - nlo.x) * ix + 0.02f), 0.0f, 255.0f);
    uint32_t hy = (uint32_t)clampf(ceilf((chi[s].y - nlo.y) * iy + 0.02f), 0.0f, 255.0f);
    uint32_t hz = (uint32_t)clampf(ceilf((chi[s].z - nlo.z) * iz + 0.02f), 0.0f, 255.0f);
    q[0 + w] |= lx << sh; q[2 + w] |= ly << sh; q[4 + w] |= lz << sh;
    q[6 + w] |= hx << sh; q[8 + w] |= hy << sh; q[10 + w] |= hz << sh;
  }
  out[0] = make_uint4(f2u(nlo.x), f2u(nlo.y), f2u(nlo.z), ex | (ey << 8) | (ez << 16) | (imask << 24));
  out[1] = make_uint4(childBase, triBase, triMask, 0u);
  out[2] = make_uint4(q[0], q[1], q[2], q[3]);
  out[3] = make_uint4(q[4], q[5], q[6], q[7]);
  out[4] = make_uint4(q[8], q[9], q[10], q[11]);
}

HD V3 f4v(float4 f) { return mk3(f.x, f.y, f.z); }
HD float half_area(V3 lo, V3 hi) { V3 e = hi - lo; return e.x * e.y + e.y * e.z + e.z * e.x; }

// One BVH8 node from one binary subtree (greedy surface-area collapse), children assigned to slots by octant so the
// traversal can order them with a XOR (Ylitie 2017 §3.2); internal children and leaf triangles are stored contiguously.
HD void collapse_body(uint32_t i, const BuildCtx& c) {
  const uint32_t n = c.numTris;
  if (i >= c.counters[CNT_LEVEL0 + c.level]) return;
  const CItem it = c.qin[i];
  auto isInternal = [&](uint32_t r) { return r < n - 1; };
  auto count = [&](uint32_t r) -> uint32_t { return isInternal(r) ? c.rangeLast[r] - c.rangeFirst[r] + 1u : 1u; };
  const uint32_t leafMax = c.leafMax ? c.leafMax : 3u;
  // area and size of every candidate are looked up once, when it enters the list (the loads of a level are one dependent chain)
  uint32_t ch[8], cnt[8]; float area[8]; int nch = 1;
  ch[0] = it.ref; cnt[0] = count(it.ref); area[0] = half_area(f4v(c.blo[it.ref]), f4v(c.bhi[it.ref]));
  for (int phase = 0; phase < 2; ++phase) {
    const uint32_t minCount = phase == 0 ? 4u : 2u;
    while (nch < 8) {
      int best = -1; float bestA = -1.0f;
      for (int k = 0; k < nch; ++k) {
        if (!isInternal(ch[k]) || cnt[k] < minCount) continue;
        if (area[k] > bestA) { bestA = area[k]; best = k; }
      }
      if (best < 0) break;
      const uint32_t r = ch[best], l = c.left[r], rr = c.right[r];
      ch[best] = l; cnt[best] = count(l); area[best] = half_area(f4v(c.blo[l]), f4v(c.bhi[l]));
      ch[nch] = rr; cnt[nch] = count(rr); area[nch] = half_area(f4v(c.blo[rr]), f4v(c.bhi[rr])); ++nch;
    }
  }
  V3 nlo = f4v(c.blo[it.ref]), nhi = f4v(c.bhi[it.ref]);
  V3 ncen = (nlo + nhi) * 0.5f;
  // greedy child -> slot assignment: slot bit k set = child lies towards +axis k
  float cost[8][8]; int slotOf[8]; bool slotUsed[8], childDone[8];
  for (int k = 0; k < 8; ++k) { slotOf[k] = -1; slotUsed[k] = false; childDone[k] = false; }
  for (int k = 0; k < nch; ++k) {
    V3 d = (f4v(c.blo[ch[k]]) + f4v(c.bhi[ch[k]])) * 0.5f - ncen;
    for (int s = 0; s < 8; ++s) cost[k][s] = ((s & 1) ? d.x : -d.x) + ((s & 2) ? d.y : -d.y) + ((s & 4) ? d.z : -d.z);
  }
  for (int round = 0; round < nch; ++round) {
    int bk = -1, bs = -1; float bc = -3.0e38f;
    for (int k = 0; k < nch; ++k) { if (childDone[k]) continue;
      for (int s = 0; s < 8; ++s) { if (slotUsed[s]) continue; if (cost[k][s] > bc) { bc = cost[k][s]; bk = k; bs = s; } } }
    slotOf[bk] = bs; slotUsed[bs] = true; childDone[bk] = true;
  }
  int childAt[8]; for (int s = 0; s < 8; ++s) childAt[s] = -1;
  for (int k = 0; k < nch; ++k) childAt[slotOf[k]] = k;
  uint32_t nInternal = 0, nLeafTris = 0, imask = 0;
  for (int s = 0; s < 8; ++s) { int k = childAt[s]; if (k < 0) continue;
    if (isInternal(ch[k]) && count(ch[k]) > leafMax) { imask |= 1u << s; ++nInternal; } else nLeafTris += count(ch[k]); }
  uint32_t childBase = nInternal ? atomAdd(&c.counters[0], nInternal) : 0u;
  uint32_t triBase = nLeafTris ? atomAdd(&c.counters[1], nLeafTris) : 0u;
  V3 clo[8], chi[8];
  uint32_t rank = 0, toff = 0, triMask = 0;
  for (int s = 0; s < 8; ++s) {
    int k = childAt[s];
    if (k < 0) continue;
    uint32_t r = ch[k];
    clo[s] = f4v(c.blo[r]); chi[s] = f4v(c.bhi[r]);
    if (imask & (1u << s)) {
      uint32_t child = childBase + rank++;
      c.nodeParent[child] = it.node;
      uint32_t qi = atomAdd(&c.counters[CNT_LEVEL0 + c.level + 1u], 1u);
      CItem o; o.ref = r; o.node = child; c.qout[qi] = o;
    } else {
      uint32_t cnt = count(r), first = isInternal(r) ? c.rangeFirst[r] : r - (n - 1);
      triMask |= ((1u << cnt) - 1u) << (3 * s);
      for (uint32_t j = 0; j < cnt; ++j) {
        uint32_t gid = c.sortedIds[first + j]; size_t slot = (size_t)triBase + toff + j;
        c.bvhTris[3 * slot + 0] = c.wtri[3 * (size_t)gid + 0];
        c.bvhTris[3 * slot + 1] = c.wtri[3 * (size_t)gid + 1];
        c.bvhTris[3 * slot + 2] = c.wtri[3 * (size_t)gid + 2];
      }
      toff += cnt;
    }
  }
  encode_node(c.nodes + 5 * (size_t)it.node, nlo, nhi, clo, chi, imask, triMask, childBase, triBase);
  c.nodeLo[it.node] = make_float4(nlo.x, nlo.y, nlo.z, 0.0f);
  c.nodeHi[it.node] = make_float4(nhi.x, nhi.y, nhi.z, 0.0f);
  c.nodePending[it.node] = nInternal;
}

// ----------------------------------------------------------------------------------------------------
// refit (replaces tlas_update.comp + TLAS rebuild for moving nodes): topology kept, boxes recomputed bottom-up
// ----------------------------------------------------------------------------------------------------
HD void refit_tris_body(uint32_t slot, const BuildCtx& c) {
  uint32_t gid = f2u(c.bvhTris[3 * (size_t)slot + 2].y);
  if (c.instChanged && !c.instChanged[c.triInstance[gid]]) return;
  c.bvhTris[3 * (size_t)slot + 0] = c.wtri[3 * (size_t)gid + 0];
  c.bvhTris[3 * (size_t)slot + 1] = c.wtri[3 * (size_t)gid + 1];
  c.bvhTris[3 * (size_t)slot + 2] = c.wtri[3 * (size_t)gid + 2];
}
// thread per BVH8 node; nodes without internal children start, the last child to finish continues with the parent.
// c.visit[node] (zeroed before launch) counts finished internal children.  (Skipping the nodes nothing moved under was measured:
// the launch is as long as the recomputation along ONE path from a mover's leaf to the root, a serial chain through the levels.)
HD void refit_node_body(uint32_t node0, const BuildCtx& c) {
  if (c.nodePending[node0] != 0u) return;
  uint32_t node = node0;
  for (;;) {
    uint4* nd = c.nodes + 5 * (size_t)node;
#ifdef __CUDA_ARCH__
    uint4 w0 = __ldcg(nd), w1 = __ldcg(nd + 1);
#else
    uint4 w0 = nd[0], w1 = nd[1];
#endif
    uint32_t imask = w0.w >> 24, childBase = w1.x, triBase = w1.y, triMask = w1.z;
    V3 clo[8], chi[8];
    V3 nlo = mk3(3.0e38f), nhi = mk3(-3.0e38f);
    for (int s = 0; s < 8; ++s) {
      const uint32_t leafBits = (triMask >> (3 * s)) & 7u;
      if (!leafBits && !(imask & (1u << s))) continue;
      if (imask & (1u << s)) {
        uint32_t child = childBase + (uint32_t)popc32(imask & ((1u << s) - 1u));
#ifdef __CUDA_ARCH__
        clo[s] = f4v(__ldcg(&c.nodeLo[child])); chi[s] = f4v(__ldcg(&c.nodeHi[child]));
#else
        clo[s] = f4v(c.nodeLo[child]); chi[s] = f4v(c.nodeHi[child]);
#endif
      } else {
        uint32_t off = (uint32_t)popc32(triMask & ((1u << (3 * s)) - 1u)), cnt = (uint32_t)popc32(leafBits);
        V3 lo = mk3(3.0e38f), hi = mk3(-3.0e38f);
        for (uint32_t j = 0; j < cnt; ++j) {
          uint32_t gid = f2u(c.bvhTris[3 * ((size_t)triBase + off + j) + 2].y);
          lo = min3(lo, f4v(c.triLo[gid])); hi = max3(hi, f4v(c.triHi[gid]));
        }
        clo[s] = lo; chi[s] = hi;
      }
      nlo = min3(nlo, clo[s]); nhi = max3(nhi, chi[s]);
    }
    encode_node(nd, nlo, nhi, clo, chi, imask, triMask, childBase, triBase);
    c.nodeLo[node] = make_float4(nlo.x, nlo.y, nlo.z, 0.0f);
    c.nodeHi[node] = make_float4(nhi.x, nhi.y, nhi.z, 0.0f);
    if (node == 0u) return;
    memFence();
    uint32_t par = c.nodeParent[node];
    if (atomAdd(&c.visit[par], 1u) + 1u != c.nodePending[par]) return;
    memFence();
    node = par;
  }
}

// ----------------------------------------------------------------------------------------------------
// Two-level structure (ANEREND_DDGI_FLAG_TWO_LEVEL): what the reference builds — one bottom-level structure per
// MESH in object space (VulkanRenderer.cpp:1267-1417), one top-level structure over the instances tlas_update.comp:55-98 writes
// (UpdateTLASPass.cpp:198-228, rebuilt every frame) — instead of one hierarchy over a world-space copy of every instance's
// triangles.  Both levels are BVH8s made by the builder above (the instance level with one instance per leaf slot); they live in
// one node array, the instance level first.  The leaf records of the instance level name the instance; InstXform carries the
// root of its mesh's hierarchy, its first flattened triangle id and the base of its mesh's shading records.
// ----------------------------------------------------------------------------------------------------
struct BlasPrimCtx {
  const anerend_vertex_t* vertices; const uint32_t* indices; uint32_t vertexOffset, indexOffset, numTris;
  float4* wtri;                 // (p0, p1.x) (p1.yz, p2.xy) (p2.z, primitive, 0, 0): OBJECT-space vertices, not (v0, e1, e2)
  float4* triLo; float4* triHi; int* sceneBounds;
  float4* shadeTris; float4* shadeUV;       // per mesh triangle (instance / material words unused)
};
HD void blas_prim_body(uint32_t t, const BlasPrimCtx& c) {
  V3 p[3]; float sr[18];
  for (int k = 0; k < 3; ++k) {
    const anerend_vertex_t& v = c.vertices[c.vertexOffset + c.indices[c.indexOffset + 3 * t + k]];
    p[k] = mk3(v.pos[0], v.pos[1], v.pos[2]);
    sr[3 * k] = v.pos[0]; sr[3 * k + 1] = v.pos[1]; sr[3 * k + 2] = v.pos[2];
    sr[9 + 3 * k] = v.normal[0]; sr[9 + 3 * k + 1] = v.normal[1]; sr[9 + 3 * k + 2] = v.normal[2];
  }
  if (c.shadeTris) {
    float4* o = c.shadeTris + 5 * (size_t)t;
    o[0] = make_float4(sr[0], sr[1], sr[2], sr[3]); o[1] = make_float4(sr[4], sr[5], sr[6], sr[7]);
    o[2] = make_float4(sr[8], sr[9], sr[10], sr[11]); o[3] = make_float4(sr[12], sr[13], sr[14], sr[15]);
    o[4] = make_float4(sr[16], sr[17], 0.0f, 0.0f);
  }
  if (c.shadeUV) {
    float uv[6];
    for (int k = 0; k < 3; ++k) { const float* q = c.vertices[c.vertexOffset + c.indices[c.indexOffset + 3 * t + k]].uv; uv[2 * k] = q[0]; uv[2 * k + 1] = q[1]; }
    c.shadeUV[2 * (size_t)t] = make_float4(uv[0], uv[1], uv[2], uv[3]);
    c.shadeUV[2 * (size_t)t + 1] = make_float4(uv[4], uv[5], 0.0f, 0.0f);
  }
  c.wtri[3 * (size_t)t + 0] = make_float4(p[0].x, p[0].y, p[0].z, p[1].x);
  c.wtri[3 * (size_t)t + 1] = make_float4(p[1].y, p[1].z, p[2].x, p[2].y);
  c.wtri[3 * (size_t)t + 2] = make_float4(p[2].z, u2f(t), 0.0f, 0.0f);
  V3 lo = min3(p[0], min3(p[1], p[2])), hi = max3(p[0], max3(p[1], p[2]));
  V3 pad = mk3(1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.x), fabsf(hi.x))), 1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.y), fabsf(hi.y))),
               1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.z), fabsf(hi.z))));
  lo = lo - pad; hi = hi + pad;
  c.triLo[t] = make_float4(lo.x, lo.y, lo.z, 0.0f);
  c.triHi[t] = make_float4(hi.x, hi.y, hi.z, 0.0f);
  atomMin(&c.sceneBounds[0], float_to_ordered(lo.x)); atomMin(&c.sceneBounds[1], float_to_ordered(lo.y)); atomMin(&c.sceneBounds[2], float_to_ordered(lo.z));
  atomMax(&c.sceneBounds[3], float_to_ordered(hi.x)); atomMax(&c.sceneBounds[4], float_to_ordered(hi.y)); atomMax(&c.sceneBounds[5], float_to_ordered(hi.z));
}
// a mesh's hierarchy was built with indices relative to its own first node / first triangle slot: move them to where it lives
struct BlasRelocCtx { uint4* nodes; uint32_t nodeBase, triBase; };
HD void blas_reloc_body(uint32_t i, const BlasRelocCtx& c) {
  uint4* w1 = c.nodes + 5 * ((size_t)c.nodeBase + i) + 1;
  uint4 v = *w1; v.x += c.nodeBase; v.y += c.triBase; *w1 = v;
}
// the instances as the primitives of the top level: world box of the mesh's root box, leaf record = instance index
struct TlasPrimCtx {
  const DInstance* instances; InstXform* xforms; const uint2* instBlas;      // instBlas[i] = (root node of the mesh's hierarchy, base of its shading records)
  const uint32_t* tlasList;     // primitive j of the instance level -> instance (those with triangles)
  const float4* nodeLo; const float4* nodeHi; uint32_t numInstances;
  float4* wtri; float4* triLo; float4* triHi; int* sceneBounds;
};
HD void tlas_prim_body(uint32_t j, const TlasPrimCtx& c) {
  const uint32_t i = c.tlasList[j];
  const uint2 ib = c.instBlas[i];
  InstXform& xf = c.xforms[i];
  xf._pad[0] = u2f(ib.x); xf._pad[1] = u2f(c.instances[i].firstTri); xf._pad[2] = u2f(ib.y);
  const V3 blo = f4v(c.nodeLo[ib.x]), bhi = f4v(c.nodeHi[ib.x]);
  V3 lo = mk3(3.0e38f), hi = mk3(-3.0e38f);
  for (int k = 0; k < 8; ++k) {
    const V3 w = xform_point(xf.o2w, mk3((k & 1) ? bhi.x : blo.x, (k & 2) ? bhi.y : blo.y, (k & 4) ? bhi.z : blo.z));
    lo = min3(lo, w); hi = max3(hi, w);
  }
  V3 pad = mk3(1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.x), fabsf(hi.x))), 1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.y), fabsf(hi.y))),
               1e-5f * fmaxf(1.0f, fmaxf(fabsf(lo.z), fabsf(hi.z))));
  lo = lo - pad; hi = hi + pad;
  c.wtri[3 * (size_t)j + 0] = make_float4(u2f(i), 0.0f, 0.0f, 0.0f);
  c.wtri[3 * (size_t)j + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  c.wtri[3 * (size_t)j + 2] = make_float4(0.0f, u2f(j), 0.0f, 0.0f);
  c.triLo[j] = make_float4(lo.x, lo.y, lo.z, 0.0f);
  c.triHi[j] = make_float4(hi.x, hi.y, hi.z, 0.0f);
  atomMin(&c.sceneBounds[0], float_to_ordered(lo.x)); atomMin(&c.sceneBounds[1], float_to_ordered(lo.y)); atomMin(&c.sceneBounds[2], float_to_ordered(lo.z));
  atomMax(&c.sceneBounds[3], float_to_ordered(hi.x)); atomMax(&c.sceneBounds[4], float_to_ordered(hi.y)); atomMax(&c.sceneBounds[5], float_to_ordered(hi.z));
}

// ----------------------------------------------------------------------------------------------------
// traversal (replaces traceRayEXT, rgen:96 and rchit:55).  Per-thread short-stack walk of the BVH8.
// ----------------------------------------------------------------------------------------------------
struct Hit { float t, u, v; uint32_t prim; uint32_t inst; };      // inst: two-level structures only

HD float dotF(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
HD V3 crossF(V3 a, V3 b) { return mk3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))); }
// Möller–Trumbore on (v0,e1,e2), two-sided, tmin < t < tmax exclusive — identical association to the oracle's.
HD bool tri_intersect(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tmin, float tmax, float& t, float& u, float& v) {
  V3 p = crossF(d, e2);
  float det = dotF(e1, p);
  if (det == 0.0f) return false;
  float inv = 1.0f / det;
  V3 tv = o - v0;
  u = dotF(tv, p) * inv;
  if (!(u >= 0.0f && u <= 1.0f)) return false;
  V3 q = crossF(tv, e1);
  v = dotF(d, q) * inv;
  if (!(v >= 0.0f && u + v <= 1.0f)) return false;
  t = dotF(e2, q) * inv;
  return t > tmin && t < tmax;
}

#define DDGI_STACK 48
HD float safe_dir(float d) { return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d; }
HD uint32_t byte_of(uint32_t lo, uint32_t hi, int i) { return ((i < 4 ? lo : hi) >> (8 * (i & 3))) & 0xFFu; }

template <bool ANY, bool CNT>
HD bool traverse(const uint4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t numNodes, V3 o, V3 d,
                 float tmin, float tmax, Hit& hit, uint32_t& cntNodes, uint32_t& cntTris) {
  hit.prim = DDGI_MISS; hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f;
  if (numNodes == 0u) return false;
  const float idx = 1.0f / safe_dir(d.x), idy = 1.0f / safe_dir(d.y), idz = 1.0f / safe_dir(d.z);
  const uint32_t octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
  uint2 stack[DDGI_STACK]; int sp = 0;
  uint2 ng = make_uint2(0u, 0x80000000u);
  uint32_t tgAll = 0u;
  for (;;) {
    uint2 tg = make_uint2(0u, 0u);
    {                                            // ng always holds at least one unvisited child here
      const int bit = msb32(ng.y);
      ng.y &= ~(1u << bit);
      if (ng.y & 0xFF000000u) { if (sp < DDGI_STACK) stack[sp++] = ng; }
      const uint32_t slot = ((uint32_t)bit - 24u) ^ octinv;
      const uint32_t rel = (uint32_t)popc32(ng.y & ~(0xFFFFFFFFu << slot) & 0xFFu);
      const uint4* nd = nodes + 5 * (size_t)(ng.x + rel);
      const uint4 n0 = nd[0], n1 = nd[1], n2 = nd[2], n3 = nd[3], n4 = nd[4];
      if (CNT) ++cntNodes;
      const float adjx = u2f((n0.w & 0xFFu) << 23) * idx, adjy = u2f(((n0.w >> 8) & 0xFFu) << 23) * idy, adjz = u2f(((n0.w >> 16) & 0xFFu) << 23) * idz;
      const float orgx = (u2f(n0.x) - o.x) * idx, orgy = (u2f(n0.y) - o.y) * idy, orgz = (u2f(n0.z) - o.z) * idz;
      const float tcur = hit.t;
      uint32_t hitmask = 0;
      const uint32_t imask = n0.w >> 24, triMask = n1.z;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t leafBits = (triMask >> (3 * i)) & 7u;
        if (!leafBits && !((imask >> i) & 1u)) continue;
        const float tlx = fmaf((float)byte_of(n2.x, n2.y, i), adjx, orgx), thx = fmaf((float)byte_of(n3.z, n3.w, i), adjx, orgx);
        const float tly = fmaf((float)byte_of(n2.z, n2.w, i), adjy, orgy), thy = fmaf((float)byte_of(n4.x, n4.y, i), adjy, orgy);
        const float tlz = fmaf((float)byte_of(n3.x, n3.y, i), adjz, orgz), thz = fmaf((float)byte_of(n4.z, n4.w, i), adjz, orgz);
        const float tn = fmaxf(fmaxf(fminf(tlx, thx), fminf(tly, thy)), fmaxf(fminf(tlz, thz), tmin));
        const float tfb = fminf(fminf(fmaxf(tlx, thx), fmaxf(tly, thy)), fmaxf(tlz, thz)) * 1.0000004f;
        const float tf = fminf(tfb, tcur);
        if (tn <= tf) {
          if ((imask >> i) & 1u) hitmask |= 1u << (24u + ((uint32_t)i ^ octinv));
          else hitmask |= leafBits << (3 * i);
        }
      }
      ng = make_uint2(n1.x, (hitmask & 0xFF000000u) | (n0.w >> 24));
      tg = make_uint2(n1.y, hitmask & 0x00FFFFFFu);
      tgAll = triMask;
    }
    while (tg.y) {
      const int b = msb32(tg.y & (0u - tg.y));   // lowest set bit
      tg.y &= tg.y - 1u;
      const float4* tp = tris + 3 * (size_t)(tg.x + (uint32_t)popc32(tgAll & ((1u << b) - 1u)));
      const float4 a = tp[0], bb = tp[1], cc = tp[2];
      if (CNT) ++cntTris;
      float t, u, v;
      if (tri_intersect(o, d, mk3(a.x, a.y, a.z), mk3(a.w, bb.x, bb.y), mk3(bb.z, bb.w, cc.x), tmin, tmax, t, u, v)) {
        if (ANY) { hit.t = t; hit.prim = f2u(cc.y); return true; }
        const uint32_t prim = f2u(cc.y);
        if (t < hit.t || (t == hit.t && prim < hit.prim)) { hit.t = t; hit.u = u; hit.v = v; hit.prim = prim; }
      }
    }
    if ((ng.y & 0xFF000000u) == 0u) {
      if (sp == 0) break;
      ng = stack[--sp];
    }
  }
  return hit.prim != DDGI_MISS;
}

// Two-level traversal (see "Two-level structure" above).  The walk starts in the instance level with the world-space ray; a leaf
// hit there pushes an entry per instance; popping one moves the ray into the instance's object space (same parameter t: the
// direction is not re-normalised) and walks the mesh's hierarchy, a RESTORE entry underneath brings the world-space ray back.
// Boxes only cull: the ray/triangle test runs in WORLD space on the vertices transformed exactly as flatten_tri_body does, so
// (t, u, v), the winning triangle and its flattened id are bit-identical to the one-level structure's and the oracle's.
// Object-space box tests are widened by what fp32 loses in the transforms (TWO_PAD_K x the magnitudes involved).
#define TWO_RESTORE 0x7FFFFFFFu
#define TWO_INSTANCE 0x80000000u
#define TWO_PAD_K 1e-5f
struct ObjRay { V3 o, d, id, pad; float padLen; uint32_t octinv; };      // pad: slack of the slab parameters per axis (t units) = padLen' * |id|, padLen' <= padLen (object units)
HD ObjRay object_ray(const InstXform& xf, V3 o, V3 d, float tmax) {
  ObjRay r;
  const float* m = xf.w2o; const float* c3 = xf.o2w + 9;
  const V3 q = mk3(o.x - c3[0], o.y - c3[1], o.z - c3[2]);
  r.o = mk3(m[0] * q.x + m[3] * q.y + m[6] * q.z, m[1] * q.x + m[4] * q.y + m[7] * q.z, m[2] * q.x + m[5] * q.y + m[8] * q.z);
  r.d = mk3(m[0] * d.x + m[3] * d.y + m[6] * d.z, m[1] * d.x + m[4] * d.y + m[7] * d.z, m[2] * d.x + m[5] * d.y + m[8] * d.z);
  r.id = mk3(1.0f / safe_dir(r.d.x), 1.0f / safe_dir(r.d.y), 1.0f / safe_dir(r.d.z));
  r.octinv = 7u - ((r.id.x < 0.0f ? 1u : 0u) | (r.id.y < 0.0f ? 2u : 0u) | (r.id.z < 0.0f ? 4u : 0u));
  // what the object-space point at parameter t <= tmax can be off by: rounding of the two products, of o - c3, of the world
  // vertices, and the mismatch of w2o with the exact inverse of o2w
  const V3 w = mk3(fabsf(o.x) + fabsf(c3[0]) + tmax * fabsf(d.x), fabsf(o.y) + fabsf(c3[1]) + tmax * fabsf(d.y), fabsf(o.z) + fabsf(c3[2]) + tmax * fabsf(d.z));
  const float own = fmaxf(fmaxf(fabsf(r.o.x), fabsf(r.o.y)), fabsf(r.o.z)) + tmax * fmaxf(fmaxf(fabsf(r.d.x), fabsf(r.d.y)), fabsf(r.d.z));
  const V3 e = mk3(fabsf(m[0]) * w.x + fabsf(m[3]) * w.y + fabsf(m[6]) * w.z + own, fabsf(m[1]) * w.x + fabsf(m[4]) * w.y + fabsf(m[7]) * w.z + own,
                   fabsf(m[2]) * w.x + fabsf(m[5]) * w.y + fabsf(m[8]) * w.z + own);
  r.pad = mk3(TWO_PAD_K * e.x * fabsf(r.id.x), TWO_PAD_K * e.y * fabsf(r.id.y), TWO_PAD_K * e.z * fabsf(r.id.z));
  r.padLen = TWO_PAD_K * fmaxf(fmaxf(e.x, e.y), e.z);
  return r;
}
// world-space (v0, e1, e2) of an object-space triangle record, the operations of flatten_tri_body
HD void world_triangle(const float* o2w, float4 a, float4 bb, float4 cc, V3& v0, V3& e1, V3& e2) {
  const V3 w0 = xform_point(o2w, mk3(a.x, a.y, a.z)), w1 = xform_point(o2w, mk3(a.w, bb.x, bb.y)), w2 = xform_point(o2w, mk3(bb.z, bb.w, cc.x));
  v0 = w0; e1 = w1 - w0; e2 = w2 - w0;
}
template <bool ANY, bool CNT>
HD bool traverse2(const uint4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t numNodes, const InstXform* __restrict__ xforms, V3 o, V3 d,
                  float tmin, float tmax, Hit& hit, uint32_t& cntNodes, uint32_t& cntTris) {
  hit.prim = DDGI_MISS; hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f; hit.inst = DDGI_MISS;
  if (numNodes == 0u) return false;
  ObjRay w;                                      // the world-space ray in the same form
  w.o = o; w.d = d; w.id = mk3(1.0f / safe_dir(d.x), 1.0f / safe_dir(d.y), 1.0f / safe_dir(d.z)); w.pad = mk3(0.0f); w.padLen = 0.0f;
  w.octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
  ObjRay r = w;
  uint32_t inst = DDGI_MISS;                     // MISS: in the instance level
  uint2 stack[DDGI_STACK]; int sp = 0;
  uint2 ng = make_uint2(0u, 0x80000000u);
  for (;;) {
    uint2 tg; uint32_t tgAll;
    {
      const int bit = msb32(ng.y);
      ng.y &= ~(1u << bit);
      if (ng.y & 0xFF000000u) { if (sp < DDGI_STACK) stack[sp++] = ng; }
      const uint32_t slot = ((uint32_t)bit - 24u) ^ r.octinv;
      const uint32_t rel = (uint32_t)popc32(ng.y & ~(0xFFFFFFFFu << slot) & 0xFFu);
      const uint4* nd = nodes + 5 * (size_t)(ng.x + rel);
      const uint4 n0 = nd[0], n1 = nd[1], n2 = nd[2], n3 = nd[3], n4 = nd[4];
      if (CNT) ++cntNodes;
      const float adjx = u2f((n0.w & 0xFFu) << 23) * r.id.x, adjy = u2f(((n0.w >> 8) & 0xFFu) << 23) * r.id.y, adjz = u2f(((n0.w >> 16) & 0xFFu) << 23) * r.id.z;
      const float orgx = (u2f(n0.x) - r.o.x) * r.id.x, orgy = (u2f(n0.y) - r.o.y) * r.id.y, orgz = (u2f(n0.z) - r.o.z) * r.id.z;
      const float tcur = ANY ? tmax : hit.t;
      uint32_t hitmask = 0;
      const uint32_t imask = n0.w >> 24, triMask = n1.z;
      for (int i = 0; i < 8; ++i) {
        const uint32_t leafBits = (triMask >> (3 * i)) & 7u;
        if (!leafBits && !((imask >> i) & 1u)) continue;
        const float tlx = fmaf((float)byte_of(n2.x, n2.y, i), adjx, orgx), thx = fmaf((float)byte_of(n3.z, n3.w, i), adjx, orgx);
        const float tly = fmaf((float)byte_of(n2.z, n2.w, i), adjy, orgy), thy = fmaf((float)byte_of(n4.x, n4.y, i), adjy, orgy);
        const float tlz = fmaf((float)byte_of(n3.x, n3.y, i), adjz, orgz), thz = fmaf((float)byte_of(n4.z, n4.w, i), adjz, orgz);
        const float tn = fmaxf(fmaxf(fminf(tlx, thx) - r.pad.x, fminf(tly, thy) - r.pad.y), fmaxf(fminf(tlz, thz) - r.pad.z, tmin));
        const float tfb = fminf(fminf(fmaxf(tlx, thx) + r.pad.x, fmaxf(tly, thy) + r.pad.y), fmaxf(tlz, thz) + r.pad.z) * 1.0000004f;
        const float tf = fminf(tfb, tcur);
        if (tn <= tf) {
          if ((imask >> i) & 1u) hitmask |= 1u << (24u + ((uint32_t)i ^ r.octinv));
          else hitmask |= leafBits << (3 * i);
        }
      }
      ng = make_uint2(n1.x, (hitmask & 0xFF000000u) | (n0.w >> 24));
      tg = make_uint2(n1.y, hitmask & 0x00FFFFFFu);
      tgAll = triMask;
    }
    if (inst == DDGI_MISS) {
      if (tg.y) {                                // instances met: they come before this node's remaining children
        if (ng.y & 0xFF000000u) { if (sp < DDGI_STACK) stack[sp++] = ng; }
        ng.y &= 0x00FFFFFFu;
        while (tg.y) {
          const int b = msb32(tg.y & (0u - tg.y));
          tg.y &= tg.y - 1u;
          const uint32_t rec = tg.x + (uint32_t)popc32(tgAll & ((1u << b) - 1u));
          if (sp < DDGI_STACK) stack[sp++] = make_uint2(TWO_INSTANCE | f2u(tris[3 * (size_t)rec].x), 0u);
        }
      }
    } else {
      const InstXform& xf = xforms[inst];
      while (tg.y) {
        const int b = msb32(tg.y & (0u - tg.y));
        tg.y &= tg.y - 1u;
        const float4* tp = tris + 3 * (size_t)(tg.x + (uint32_t)popc32(tgAll & ((1u << b) - 1u)));
        const float4 a = tp[0], bb = tp[1], cc = tp[2];
        if (CNT) ++cntTris;
        V3 v0, e1, e2; world_triangle(xf.o2w, a, bb, cc, v0, e1, e2);
        float t, u, v;
        if (tri_intersect(o, d, v0, e1, e2, tmin, tmax, t, u, v)) {
          const uint32_t prim = f2u(xf._pad[1]) + f2u(cc.y);
          if (ANY) { hit.t = t; hit.prim = prim; hit.inst = inst; return true; }
          if (t < hit.t || (t == hit.t && prim < hit.prim)) { hit.t = t; hit.u = u; hit.v = v; hit.prim = prim; hit.inst = inst; }
        }
      }
    }
    while ((ng.y & 0xFF000000u) == 0u) {
      if (sp == 0) return hit.prim != DDGI_MISS;
      const uint2 e = stack[--sp];
      if (e.y != 0u) { ng = e; break; }
      if (e.x == TWO_RESTORE) { r = w; inst = DDGI_MISS; continue; }
      inst = e.x & ~TWO_INSTANCE;
      r = object_ray(xforms[inst], o, d, ANY ? tmax : hit.t);
      if (sp < DDGI_STACK) stack[sp++] = make_uint2(TWO_RESTORE, 0u);
      ng = make_uint2(f2u(xforms[inst]._pad[0]), 0x80000000u);
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// shading: irradiance_probe_update.rchit + pbr_light.glsl + bindless.glsl + octahedron_helpers.glsl
// ----------------------------------------------------------------------------------------------------
// One bindless texture (bindless.glsl:179 `textures[]`): mip 0 in the byte pool, sampled like the reference's default sampler
// (VK_FILTER_LINEAR, REPEAT or CLAMP_TO_EDGE per texture, PipelineUtil.h:98-112) at LOD 0 — ray-tracing stages have no derivatives.
struct DTexture { uint32_t offset, width, height, formatClamp; };       // formatClamp: ANEREND_TEX_* | clamp_to_edge << 8
struct TexCtx { const DTexture* textures; const uint8_t* texPool; const float* srgbLut; uint32_t numTextures;
                const int32_t* tileInfos; uint32_t numTiles; };      // tileInfoBuffer (bindless.glsl binding 14): ddgiAtlasTex per paged tile, or null

struct TraceCtx {
  const uint4* nodes; const float4* bvhTris; uint32_t numNodes;
  uint32_t twoLevel;           // ANEREND_DDGI_FLAG_TWO_LEVEL: instance level + per-mesh hierarchies (traverse2), no triInstance
  const anerend_vertex_t* vertices; const uint32_t* indices; const anerend_material_info_t* materials;
  const DInstance* instances; const InstXform* xforms; const uint32_t* triInstance;
  const anerend_light_t* lights; uint32_t numLights;
  const float4* dirs;          // [nY][N] by absolute layer: mat3(randomRot) * fibonacci dir, computed on the host (rgen:87-88)
  const int* layerMap;         // [numLayers] absolute layer of each launched layer (a slab, or any subset of the grid)
  const uint16_t* rayOrder;    // [N] slot -> ray index (direction-binned order), or null
  const uint2* atlas;          // ProbeRaysConvTex, RGBA16F texels
  uint2* rayIrr;               // ProbeRaysIrrTex replacement: [nY][P][N] RGBA16F
  anerend_ddgi_ray_record_t* records;   // [nY][P][N] or null
  unsigned long long* stats;   // 7 counters or null
  int nXZ, nY, N, W, H;
  int numLayers;
  int Pl, rowOffset, rowStride;   // probes launched per layer: the z-rows rowOffset, rowOffset+rowStride, ... (stripe_probe)
  TexCtx tex;                     // bindless textures (set 0 binding 16)
  V3 cameraPos;                   // ubo.cameraPos.xyz: centre of the tile page (worldPosToTileInfo, bindless.glsl:291-308)
  int clampSampler;               // atlas sampler addressing: 0 = REPEAT (probe RT pass, SpecularGIRT), 1 = CLAMP_TO_EDGE (deferred lighting)
  V3 camOff, step, sky, lightDir;
  float tmin, tmax, sunIntensity;
  uint32_t uboFlags, flags;
  // light maps (below): conservative distance maps around the point lights that answer "certainly unoccluded" shadow rays
  const unsigned long long* lmStatic; const unsigned long long* lmDynamic; const uint32_t* lmNear; uint32_t lmRes;      // lmRes = 0: none
  const float4* lmTris; uint32_t lmNumTris;      // world triangles by global id (the candidates the map entries name)
  uint32_t lmDynShift, smDynShift;               // the dynamic maps are coarser: resolution R >> shift (re-made at every refit, see LM_DYN_SHIFT)
  // the sun's map: orthographic, along normalize(-lightDir); entry = (height above smBase, fp32 bits) << 32 | triangle id, kept maximal
  const unsigned long long* smStatic; const unsigned long long* smDynamic; uint32_t smRes;      // smRes = 0: none
  V3 smA, smB, smS; float smOrgA, smOrgB, smInvCell, smBase;
};

// Row stripe (multi-GPU sharding inside the layers): launched probe `pl` of a layer -> probe index in the layer.  A handle
// updates the probe rows z = rowOffset + k*rowStride of every layer it launches; (0, 1) = all rows.
HD int stripe_probe(int pl, int nXZ, int rowOffset, int rowStride) { return ((pl / nXZ) * rowStride + rowOffset) * nXZ + pl % nXZ; }
template <bool ANY, bool CNT>
HD bool traverse_scene(const TraceCtx& c, V3 o, V3 d, float tmin, float tmax, Hit& hit, uint32_t& cn, uint32_t& ct) {
  if (c.twoLevel) return traverse2<ANY, CNT>(c.nodes, c.bvhTris, c.numNodes, c.xforms, o, d, tmin, tmax, hit, cn, ct);
  return traverse<ANY, CNT>(c.nodes, c.bvhTris, c.numNodes, o, d, tmin, tmax, hit, cn, ct);
}

// ----------------------------------------------------------------------------------------------------
// Light maps: a pre-filter for the point-light shadow rays of extension E1 (rchit:45-66 composed with calcLight's point branch).
//
// Around every enabled point light a cube map (6 faces x R x R) holds, per texel, a LOWER BOUND of the distance from the light
// to any triangle that reaches into the texel's pyramid of directions, and the id of the triangle that set it.  A shadow ray
// pos -> light runs along ONE direction from the light and covers the distances (0, dist - tMin) from it:
//   * if that interval ends before the bound of the direction's texel, no triangle can intersect the ray: it is unoccluded and is
//     never traced (lightmap_unoccluded);
//   * otherwise the texel's triangle — the nearest one to the light in that direction, what a shadow map would hold — is the likely
//     occluder: the ray is put through the exact Moeller-Trumbore test against it first, and a hit ends it (any-hit semantics: WHICH
//     triangle occludes never reaches the result);
//   * the ray's far end is cut back to where geometry can begin (dist - bound), which only prunes box tests.
// Everything else — rays that graze a surface, depth complexity in front of the hit — goes through the BVH as before.  The exact
// triangle test is never replaced, only skipped where its answer is known, so lit masks and radiance stay bit-identical to the
// oracle's; the margins (LM_DELTA on the footprints, rho on the plane bound, the slack in lightmap_unoccluded) are orders of
// magnitude above fp32 rounding of either side.  At C3 (R = 512) 81 % of the unoccluded point-light rays are answered by one read.
//
//   face f = 2 * axis + (d[axis] < 0), axis = major axis of d = p - lightPos; in-face coordinates u = d[axis+1] / |d[axis]|,
//   v = d[axis+2] / |d[axis]| in [-1, 1]; texel X = floor((u + 1) / 2 * R).  Entry = bound (fp32 bits) << 32 | triangle id, kept
//   minimal by 64-bit atomicMin.
//   static map: all triangles at build_accel time.  dynamic map: the triangles of instances that moved since (refit), rasterised
//   again at every refit; the entry is the minimum of both (a stale static entry of a mover is only a bound that is too low, and a
//   candidate that misses).
// ----------------------------------------------------------------------------------------------------
#define LM_EPS_W 1e-4f          // points nearer than this to a face's apex plane are not projected (see LM_NEAR)
#define LM_NEAR 0.02f           // a triangle this close to a light disables that light's map (lmNear): its projection is unbounded
#define LM_DELTA 1e-4f          // footprint growth in face coordinates (texel size at R = 512: 3.9e-3)
#define LM_DYN_SHIFT 2u         // dynamic maps at a quarter of the resolution: movers near a light cover thousands of fine texels each, and their
                                // maps are re-made at every refit (C4: 16 boxes among the lights, 1.1 ms per refit at full resolution, 0.08 at a quarter)
#define LM_FAR_BITS 0x7F7F7F7Fu // "nothing here" (3.39e38; as a triangle id: none), the byte pattern of the clearing memset
struct LmTri { V3 r[3]; V3 n; float h, daabb; bool skinny, nearLight; };
HD float comp3(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
// vertices relative to the light -> plane, box distance; false: cannot occlude a ray of this light (beyond its range, or degenerate)
HD bool lm_tri_setup(V3 r0, V3 r1, V3 r2, float range, LmTri& t) {
  t.r[0] = r0; t.r[1] = r1; t.r[2] = r2;
  const V3 lo = min3(min3(r0, r1), r2), hi = max3(max3(r0, r1), r2);
  const float dx = fmaxf(fmaxf(lo.x, -hi.x), 0.0f), dy = fmaxf(fmaxf(lo.y, -hi.y), 0.0f), dz = fmaxf(fmaxf(lo.z, -hi.z), 0.0f);
  t.daabb = sqrtf(dx * dx + dy * dy + dz * dz) * 0.999999f;        // distance from the light to the triangle's box: a bound that needs no plane
  if (t.daabb > range) return false;                               // rays of this light are shorter than its range
  t.nearLight = t.daabb < LM_NEAR;
  const V3 e1 = r1 - r0, e2 = r2 - r0;
  const V3 n = mk3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
  const float n2 = dot3(n, n);
  if (!(n2 > 0.0f)) return false;                                  // zero area: det == 0 in the triangle test, never hit
  t.skinny = n2 < 1e-8f * dot3(e1, e1) * dot3(e2, e2);             // normal too noisy for the plane bound: box bound only
  const float inv = 1.0f / sqrtf(n2);
  t.n = n * inv; t.h = fabsf(dot3(t.n, r0));
  return true;
}
// Conservative rasterisation inside the footprint rectangle: a texel is touched unless its whole square (grown by the same margin as
// the rectangle) lies outside one edge of the projected triangle.  Edge functions in texel coordinates relative to the rectangle's
// corner (x0, y0): inside <=> A*x + B*y + C >= 0 at the texel centre, the half-texel reach and the margins folded into C.
// `use` is false (rectangle only) where the projection is not a triangle with well-conditioned edges: straddling the apex plane,
// nearly edge-on, or vertices projected far outside the map.
struct LmEdges { float A[3], B[3], C[3]; bool use; };
HD void lm_edges(const float* px, const float* py, float x0, float y0, float slackTexels, float limit, LmEdges& e) {
  const float area2 = (px[1] - px[0]) * (py[2] - py[0]) - (py[1] - py[0]) * (px[2] - px[0]);
  e.use = fabsf(area2) > 1e-3f;
  for (int k = 0; k < 3; ++k) e.use = e.use && fabsf(px[k]) <= limit && fabsf(py[k]) <= limit;
  const float sgn = area2 > 0.0f ? 1.0f : -1.0f;
  for (int i = 0; i < 3; ++i) {
    const int j = i == 2 ? 0 : i + 1;
    const float dx = px[j] - px[i], dy = py[j] - py[i];
    e.A[i] = -dy * sgn; e.B[i] = dx * sgn;
    e.C[i] = (dy * (px[i] - x0) - dx * (py[i] - y0)) * sgn + (0.5f + slackTexels) * (fabsf(dx) + fabsf(dy));
  }
}
HD bool lm_edges_touch(const LmEdges& e, int lx, int ly) {        // lx, ly: texel relative to (x0, y0)
  if (!e.use) return true;
  const float xc = (float)lx + 0.5f, yc = (float)ly + 0.5f;
  return e.A[0] * xc + e.B[0] * yc + e.C[0] >= 0.0f && e.A[1] * xc + e.B[1] * yc + e.C[1] >= 0.0f && e.A[2] * xc + e.B[2] * yc + e.C[2] >= 0.0f;
}
#define LM_EDGE_SLACK 4e-3f      // texels per unit of |dx| + |dy|: above the fp32 rounding of the edge functions for vertices within `limit`
// texel rectangle of the triangle's projection on `face`, grown by LM_DELTA; false: the triangle does not reach the face
HD bool lm_face_rect(const LmTri& t, int face, uint32_t R, int& x0, int& x1, int& y0, int& y1, LmEdges& edges) {
  const int axis = face >> 1, ab = axis == 2 ? 0 : axis + 1, ac = ab == 2 ? 0 : ab + 1;
  const float s = (face & 1) ? -1.0f : 1.0f;
  float w[3], b[3], c[3];
  for (int k = 0; k < 3; ++k) { w[k] = s * comp3(t.r[k], axis); b[k] = comp3(t.r[k], ab); c[k] = comp3(t.r[k], ac); }
  const float wmax = fmaxf(fmaxf(w[0], w[1]), w[2]), wmin = fminf(fminf(w[0], w[1]), w[2]);
  if (!(wmax > LM_EPS_W)) return false;
  float umin, umax, vmin, vmax;
  float px[3] = {0.0f, 0.0f, 0.0f}, py[3] = {0.0f, 0.0f, 0.0f};
  const bool front = wmin > LM_EPS_W;
  edges.use = false;
  if (front) {                    // wholly in front of the apex plane: the projection is the triangle of the projected vertices
    const float u0 = b[0] / w[0], u1 = b[1] / w[1], u2 = b[2] / w[2], v0 = c[0] / w[0], v1 = c[1] / w[1], v2 = c[2] / w[2];
    umin = fminf(fminf(u0, u1), u2); umax = fmaxf(fmaxf(u0, u1), u2); vmin = fminf(fminf(v0, v1), v2); vmax = fmaxf(fmaxf(v0, v1), v2);
    const float half = 0.5f * (float)R;
    px[0] = (u0 + 1.0f) * half; px[1] = (u1 + 1.0f) * half; px[2] = (u2 + 1.0f) * half;
    py[0] = (v0 + 1.0f) * half; py[1] = (v1 + 1.0f) * half; py[2] = (v2 + 1.0f) * half;
  } else {                        // straddles the apex plane: interval arithmetic over the box, w clipped to [LM_EPS_W, wmax]
    const float blo = fminf(fminf(b[0], b[1]), b[2]), bhi = fmaxf(fmaxf(b[0], b[1]), b[2]), clo = fminf(fminf(c[0], c[1]), c[2]), chi = fmaxf(fmaxf(c[0], c[1]), c[2]);
    const float il = 1.0f / LM_EPS_W, ih = 1.0f / wmax;
    umin = fminf(fminf(blo * il, blo * ih), fminf(bhi * il, bhi * ih)); umax = fmaxf(fmaxf(blo * il, blo * ih), fmaxf(bhi * il, bhi * ih));
    vmin = fminf(fminf(clo * il, clo * ih), fminf(chi * il, chi * ih)); vmax = fmaxf(fmaxf(clo * il, clo * ih), fmaxf(chi * il, chi * ih));
  }
  umin -= LM_DELTA; umax += LM_DELTA; vmin -= LM_DELTA; vmax += LM_DELTA;
  if (umax < -1.0f || umin > 1.0f || vmax < -1.0f || vmin > 1.0f) return false;
  const float fr = (float)R, top = (float)(R - 1u);
  x0 = (int)fminf(fmaxf(floorf((umin + 1.0f) * 0.5f * fr), 0.0f), top); x1 = (int)fminf(fmaxf(floorf((umax + 1.0f) * 0.5f * fr), 0.0f), top);
  y0 = (int)fminf(fmaxf(floorf((vmin + 1.0f) * 0.5f * fr), 0.0f), top); y1 = (int)fminf(fmaxf(floorf((vmax + 1.0f) * 0.5f * fr), 0.0f), top);
  if (front) lm_edges(px, py, (float)x0, (float)y0, LM_DELTA * 0.5f * fr + LM_EDGE_SLACK, 3.0f * fr, edges);
  return true;
}
// lower bound of the distance from the light to the triangle inside texel (X, Y) of `face`: the triangle's plane cut by the texel's
// direction nearest to the plane normal (|cos| of the centre direction + the texel's angular radius rho), never below the box distance
HD float lm_texel_bound(V3 n, float h, float daabb, bool skinny, int face, int X, int Y, uint32_t R, float rho) {
  const int axis = face >> 1, ab = axis == 2 ? 0 : axis + 1, ac = ab == 2 ? 0 : ab + 1;
  const float cu = ((float)X + 0.5f) * (2.0f / (float)R) - 1.0f, cv = ((float)Y + 0.5f) * (2.0f / (float)R) - 1.0f;
  const float s = (face & 1) ? -1.0f : 1.0f;
  const float inv = 1.0f / sqrtf(1.0f + cu * cu + cv * cv);
  const float cosn = fabsf(comp3(n, axis) * s + comp3(n, ab) * cu + comp3(n, ac) * cv) * inv;
  const float plane = h / fminf(1.0f, cosn + rho);
  return (skinny ? daabb : fmaxf(plane, daabb)) * 0.999996f;
}
HD float lm_rho(uint32_t R) { return 1.4143f / (float)R * 1.02f + 2e-5f; }
// The entry of the texel the direction light -> pos falls in.
struct LmEntry { float bound; uint32_t cand; };
HD LmEntry lightmap_lookup(const TraceCtx& c, uint32_t li, V3 lp, V3 pos) {
  LmEntry e; e.bound = 0.0f; e.cand = LM_FAR_BITS;
  if (c.lmRes == 0u || c.lmNear[li] != 0u) return e;
  const V3 d = pos - lp;
  const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int face; float w, b, cc;
  if (ax >= ay && ax >= az) { face = d.x < 0.0f ? 1 : 0; w = ax; b = d.y; cc = d.z; }
  else if (ay >= az) { face = d.y < 0.0f ? 3 : 2; w = ay; b = d.z; cc = d.x; }
  else { face = d.z < 0.0f ? 5 : 4; w = az; b = d.x; cc = d.y; }
  if (!(w > 0.0f)) return e;
#ifdef __CUDA_ARCH__
  const float iw = __fdividef(1.0f, w);      // texel choice: any rounding far below LM_DELTA will do
#else
  const float iw = 1.0f / w;
#endif
  const float fr = (float)c.lmRes, top = (float)(c.lmRes - 1u);
  const int X = (int)fminf(fmaxf(floorf((b * iw + 1.0f) * 0.5f * fr), 0.0f), top), Y = (int)fminf(fmaxf(floorf((cc * iw + 1.0f) * 0.5f * fr), 0.0f), top);
  const size_t o = (((size_t)li * 6u + (size_t)face) * c.lmRes + (size_t)Y) * c.lmRes + (size_t)X;
  unsigned long long v = c.lmStatic[o];
  if (c.lmDynamic) {                                                                            // positive floats order like their bit patterns
    const uint32_t rd = c.lmRes >> c.lmDynShift;
    const unsigned long long dv = c.lmDynamic[(((size_t)li * 6u + (size_t)face) * rd + (size_t)(Y >> c.lmDynShift)) * rd + (size_t)(X >> c.lmDynShift)];
    v = dv < v ? dv : v;
  }
  e.bound = u2f((uint32_t)(v >> 32)); e.cand = (uint32_t)v;
  return e;
}
// Is the shadow ray pos -> light (tMin, dist) certainly unoccluded?  `dist` = length(lightPos - pos) as the hit shader computes it.
HD bool lightmap_unoccluded(const LmEntry& e, float dist, float tmin) { return dist - tmin < e.bound * 0.9999f - 1e-3f; }
// no triangle lies nearer to the light than the bound: intersections have t <= dist - bound
HD float lightmap_tmax(const LmEntry& e, float dist) { return fminf(dist, dist - e.bound * 0.9999f + 1e-3f); }
// the exact test against the texel's triangle (same operands as the traversal would hand it)
HD bool lightmap_candidate_occludes(const TraceCtx& c, const LmEntry& e, V3 pos, V3 dir, float tmin, float dist) {
  if (e.cand >= c.lmNumTris) return false;
  const float4 a = c.lmTris[3 * (size_t)e.cand], b = c.lmTris[3 * (size_t)e.cand + 1], cc = c.lmTris[3 * (size_t)e.cand + 2];
  float t, u, v;
  return tri_intersect(pos, dir, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, cc.x), tmin, dist, t, u, v);
}
// Rasterisation, one (triangle, light) pair: every face, every texel of the footprint.  The CUDA kernel (k_lightmap_raster,
// ddgi_api.cu) runs the same setup per thread and spreads the texel loops over the warp.
struct LmBuildCtx {
  const float4* wtri; const uint32_t* triInstance; const uint32_t* moved;      // moved: per instance, or null = every triangle
  const anerend_light_t* lights; uint32_t numTris, numLights, R; float rho;
  unsigned long long* depth; uint32_t* nearFlags;
};
HD bool lm_pair_setup(const LmBuildCtx& c, uint32_t t, uint32_t li, LmTri& tri) {
  if (c.moved && !c.moved[c.triInstance[t]]) return false;
  const anerend_light_t* lt = &c.lights[li];
  if (!(lt->color[3] > 0.5f)) return false;
  const float4 a = c.wtri[3 * (size_t)t], b = c.wtri[3 * (size_t)t + 1], cc = c.wtri[3 * (size_t)t + 2];
  const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
  const V3 v0 = mk3(a.x, a.y, a.z), e1 = mk3(a.w, b.x, b.y), e2 = mk3(b.z, b.w, cc.x);
  return lm_tri_setup(v0 - lp, (v0 + e1) - lp, (v0 + e2) - lp, lt->world_pos[3], tri);
}
HD void lightmap_raster_body(uint32_t i, const LmBuildCtx& c) {
  const uint32_t t = i % c.numTris, li = i / c.numTris;
  LmTri tri;
  if (!lm_pair_setup(c, t, li, tri)) return;
  if (tri.nearLight) c.nearFlags[li] = 1u;
  for (int face = 0; face < 6; ++face) {
    int x0, x1, y0, y1; LmEdges ed;
    if (!lm_face_rect(tri, face, c.R, x0, x1, y0, y1, ed)) continue;
    unsigned long long* dm = c.depth + ((size_t)li * 6u + (size_t)face) * c.R * c.R;
    for (int y = y0; y <= y1; ++y) for (int x = x0; x <= x1; ++x)
      if (lm_edges_touch(ed, x - x0, y - y0)) atomMin64(dm + (size_t)y * c.R + x, ((unsigned long long)f2u(lm_texel_bound(tri.n, tri.h, tri.daabb, tri.skinny, face, x, y, c.R, c.rho)) << 32) | t);
  }
}
// ---- the sun (rchit:45-66): all its shadow rays run along smS = normalize(-lightDir), so the map is an orthographic height field over
// the plane (smA, smB): per texel an UPPER bound of s = dot(p, smS) over every triangle that reaches into the texel's column, and
// the triangle that set it.  A ray from pos covers s in (s(pos) + tMin, s(pos) + tMax): above the bound it is unoccluded.
#define SM_DELTA 2e-4f          // footprint growth in metres (fp32 rounding of the projections: ~1e-5 at scene scale)
struct SunMapCtx {
  const float4* wtri; const uint32_t* triInstance; const uint32_t* moved; uint32_t numTris, R;
  V3 a, b, s; float orgA, orgB, invCell, cell, base;
  unsigned long long* height;
};
struct SmTri { float pa[3], pb[3], ps[3]; float smax, ga, gb; bool flat; };     // flat: plane usable, s = ps[0] + ga * (a - pa[0]) + gb * (b - pb[0])
HD bool sm_tri_setup(const SunMapCtx& c, uint32_t t, SmTri& tri, int& x0, int& x1, int& y0, int& y1, LmEdges& edges) {
  if (c.moved && !c.moved[c.triInstance[t]]) return false;
  const float4 A = c.wtri[3 * (size_t)t], B = c.wtri[3 * (size_t)t + 1], C = c.wtri[3 * (size_t)t + 2];
  const V3 v0 = mk3(A.x, A.y, A.z), e1 = mk3(A.w, B.x, B.y), e2 = mk3(B.z, B.w, C.x);
  const V3 v[3] = {v0, v0 + e1, v0 + e2};
  for (int k = 0; k < 3; ++k) { tri.pa[k] = dot3(v[k], c.a); tri.pb[k] = dot3(v[k], c.b); tri.ps[k] = dot3(v[k], c.s); }
  const float amin = fminf(fminf(tri.pa[0], tri.pa[1]), tri.pa[2]) - SM_DELTA, amax = fmaxf(fmaxf(tri.pa[0], tri.pa[1]), tri.pa[2]) + SM_DELTA;
  const float bmin = fminf(fminf(tri.pb[0], tri.pb[1]), tri.pb[2]) - SM_DELTA, bmax = fmaxf(fmaxf(tri.pb[0], tri.pb[1]), tri.pb[2]) + SM_DELTA;
  tri.smax = fmaxf(fmaxf(tri.ps[0], tri.ps[1]), tri.ps[2]);
  const float fr = (float)c.R;
  const float fx0 = floorf((amin - c.orgA) * c.invCell), fx1 = floorf((amax - c.orgA) * c.invCell), fy0 = floorf((bmin - c.orgB) * c.invCell), fy1 = floorf((bmax - c.orgB) * c.invCell);
  if (fx1 < 0.0f || fy1 < 0.0f || fx0 >= fr || fy0 >= fr) return false;             // wholly outside the map: rays there are not answered by it
  x0 = (int)fmaxf(fx0, 0.0f); x1 = (int)fminf(fx1, fr - 1.0f); y0 = (int)fmaxf(fy0, 0.0f); y1 = (int)fminf(fy1, fr - 1.0f);
  { float px[3], py[3];
    for (int k = 0; k < 3; ++k) { px[k] = (tri.pa[k] - c.orgA) * c.invCell; py[k] = (tri.pb[k] - c.orgB) * c.invCell; }
    lm_edges(px, py, (float)x0, (float)y0, SM_DELTA * c.invCell + LM_EDGE_SLACK, 3.0f * fr, edges); }
  // plane in map coordinates through the three projected vertices: s = k + ga * a + gb * b
  const float ua = tri.pa[1] - tri.pa[0], ub = tri.pb[1] - tri.pb[0], us = tri.ps[1] - tri.ps[0];
  const float va = tri.pa[2] - tri.pa[0], vb = tri.pb[2] - tri.pb[0], vs = tri.ps[2] - tri.ps[0];
  const float det = ua * vb - ub * va;
  const float scale = (fabsf(ua) + fabsf(ub)) * (fabsf(va) + fabsf(vb));
  tri.flat = fabsf(det) > 1e-4f * scale && scale > 0.0f;                            // else: (nearly) edge-on to the sun, highest vertex only
  if (tri.flat) {
    const float inv = 1.0f / det;
    tri.ga = (us * vb - ub * vs) * inv; tri.gb = (ua * vs - us * va) * inv;
  } else { tri.ga = 0.0f; tri.gb = 0.0f; }
  return true;
}
// upper bound of s over the part of the triangle's plane above texel (X, Y), evaluated relative to vertex 0 (no cancellation)
HD float sm_texel_bound(float smax, float ga, float gb, float pa0, float pb0, float ps0, bool flat, int X, int Y, float orgA, float orgB, float cell) {
  if (!flat) return smax;
  const float ca = orgA + ((float)X + 0.5f) * cell, cb = orgB + ((float)Y + 0.5f) * cell;
  const float top = ps0 + ga * (ca - pa0) + gb * (cb - pb0) + (fabsf(ga) + fabsf(gb)) * (0.5f * cell + SM_DELTA);
  return fminf(smax, top + 1e-4f + 1e-5f * fabsf(top));
}
HD unsigned long long sm_pack(float bound, float base, uint32_t t) { return ((unsigned long long)f2u(fmaxf(bound - base, 0.0f)) << 32) | t; }
HD void sunmap_raster_body(uint32_t t, const SunMapCtx& c) {
  SmTri tri; int x0, x1, y0, y1; LmEdges ed;
  if (!sm_tri_setup(c, t, tri, x0, x1, y0, y1, ed)) return;
  for (int y = y0; y <= y1; ++y) for (int x = x0; x <= x1; ++x)
    if (lm_edges_touch(ed, x - x0, y - y0)) atomMax64(c.height + (size_t)y * c.R + x, sm_pack(sm_texel_bound(tri.smax, tri.ga, tri.gb, tri.pa[0], tri.pb[0], tri.ps[0], tri.flat, x, y, c.orgA, c.orgB, c.cell), c.base, t));
}
struct SmEntry { float top; uint32_t cand; bool inside; };
HD SmEntry sunmap_lookup(const TraceCtx& c, V3 pos) {
  SmEntry e; e.top = 0.0f; e.cand = LM_FAR_BITS; e.inside = false;
  if (c.smRes == 0u) return e;
  const float fx = floorf((dot3(pos, c.smA) - c.smOrgA) * c.smInvCell), fy = floorf((dot3(pos, c.smB) - c.smOrgB) * c.smInvCell), fr = (float)c.smRes;
  if (!(fx >= 0.0f && fy >= 0.0f && fx < fr && fy < fr)) return e;
  const size_t o = (size_t)fy * c.smRes + (size_t)fx;
  unsigned long long v = c.smStatic[o];
  if (c.smDynamic) {
    const unsigned long long dv = c.smDynamic[(size_t)((uint32_t)fy >> c.smDynShift) * (c.smRes >> c.smDynShift) + (size_t)((uint32_t)fx >> c.smDynShift)];
    v = dv > v ? dv : v;
  }
  e.inside = true; e.cand = v ? (uint32_t)v : LM_FAR_BITS;                          // 0 = empty column
  e.top = v ? u2f((uint32_t)(v >> 32)) + c.smBase : -3.0e38f;
  return e;
}
HD bool sunmap_unoccluded(const SmEntry& e, float sPos, float tmin) { return e.inside && sPos + tmin > e.top + 1e-3f + 1e-4f * fabsf(e.top); }
HD float sunmap_tmax(const SmEntry& e, float sPos, float tmax) { return e.inside ? fminf(tmax, (e.top - sPos) + 2e-3f + 2e-4f * fabsf(e.top)) : tmax; }
HD bool sunmap_candidate_occludes(const TraceCtx& c, const SmEntry& e, V3 pos, V3 dir, float tmin, float tmax) {
  LmEntry le; le.bound = 0.0f; le.cand = e.cand;
  return lightmap_candidate_occludes(c, le, pos, dir, tmin, tmax);
}

// tlas_update.comp:41-47: a renderable is in the acceleration structure iff it is visible and not culled.  Compares that test on the
// (GPU-written) renderable buffer with the selection the structure was built from.
struct VisCtx { const anerend_renderable_t* renderables; const uint32_t* selected; uint32_t* changed; };
HD void vis_changed_body(uint32_t i, const VisCtx& c) {
  const uint32_t v = c.renderables[i].visible;
  const uint32_t sel = ((v & 1u) != 0u && ((v >> 1) & 1u) == 0u) ? 1u : 0u;
  if (sel != c.selected[i]) *c.changed = 1u;
}

// instances whose transform differs from the one the static map was rasterised with
struct LmMovedCtx { const InstXform* xforms; const InstXform* base; uint32_t* moved; };
HD void lm_moved_body(uint32_t i, const LmMovedCtx& c) {
  bool m = false;
  for (int k = 0; k < 12; ++k) m = m || f2u(c.xforms[i].o2w[k]) != f2u(c.base[i].o2w[k]);
  c.moved[i] = m ? 1u : 0u;
}

HD float sign_not_zero(float k) { return (k >= 0.0f) ? 1.0f : -1.0f; }            // octahedron_helpers.glsl:1-3
HD void oct_encode(V3 v, float& ox, float& oy) {                                    // octahedron_helpers.glsl:10-20
  V3 t = v; v.y = t.z; v.z = t.y;
  float l1 = xadd(xadd(fabsf(v.x), fabsf(v.y)), fabsf(v.z));
  float inv = xdiv(1.0f, l1);
  float rx = xmul(v.x, inv), ry = xmul(v.y, inv);
  if (v.z < 0.0f) { float nx = xmul(xsub(1.0f, fabsf(ry)), sign_not_zero(rx)), ny = xmul(xsub(1.0f, fabsf(rx)), sign_not_zero(ry)); rx = nx; ry = ny; }
  ox = rx; oy = ry;
}
HD V3 oct_decode(float ox, float oy) {                                              // octahedron_helpers.glsl:25-34
  V3 v = mk3(ox, oy, 1.0f - fabsf(ox) - fabsf(oy));
  if (v.z < 0.0f) { float nx = (1.0f - fabsf(v.y)) * sign_not_zero(v.x), ny = (1.0f - fabsf(v.x)) * sign_not_zero(v.y); v.x = nx; v.y = ny; }
  V3 t = v; v.y = t.z; v.z = t.y;
  return normalize3(v);
}
HD float to_linear1(float c) {                                                      // bindless.glsl:240-247
  float higher = powf((c + 0.055f) / 1.055f, 2.4f), lower = c / 12.92f;
  return (c < 0.04045f) ? lower : higher;
}

// texture(textures[idx], uv) at LOD 0.  Defined arithmetic (the reference leaves it to the sampler hardware): unnormalised
// coordinate uv*size - 0.5, floor / fract weights in fp32, texel decode UNORM8 -> c/255, sRGB8 -> linear (both through a host-made
// table; alpha stays linear), missing components (0, 0, 1), blend w00*t00 + w10*t10 + w01*t01 + w11*t11 like the atlas taps.
HD int tex_wrap(int i, int n, bool clampEdge) {
  if (clampEdge) return clampi(i, 0, n - 1);
  i %= n; return i < 0 ? i + n : i;
}
HD void tex_fetch(const TexCtx& c, const DTexture& t, int x, int y, float out[4]) {
  const uint32_t fmt = t.formatClamp & 0xFFu;
  const uint32_t bpp = fmt == ANEREND_TEX_RGBA16F ? 8u : fmt <= ANEREND_TEX_RGBA8_SRGB ? 4u : (fmt == ANEREND_TEX_RG8_UNORM ? 2u : 1u);
  const uint8_t* p = c.texPool + t.offset + ((size_t)y * t.width + (size_t)x) * bpp;
  out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1.0f;
  if (fmt == ANEREND_TEX_RGBA16F) {          // baked DDGI atlases (VulkanRenderer.cpp:2759); texels are 8-byte aligned in the pool
    const uint2 h = *reinterpret_cast<const uint2*>(p);
    out[0] = h2f((uint16_t)(h.x & 0xFFFFu)); out[1] = h2f((uint16_t)(h.x >> 16)); out[2] = h2f((uint16_t)(h.y & 0xFFFFu)); out[3] = h2f((uint16_t)(h.y >> 16));
    return;
  }
  // srgbLut[0..255] = sRGB EOTF of c/255, srgbLut[256..511] = c/255 (both made on the host, one table for oracle and device)
  if (fmt == ANEREND_TEX_RGBA8_SRGB) { out[0] = c.srgbLut[p[0]]; out[1] = c.srgbLut[p[1]]; out[2] = c.srgbLut[p[2]]; out[3] = c.srgbLut[256 + p[3]]; }
  else for (uint32_t k = 0; k < bpp; ++k) out[k] = c.srgbLut[256 + p[k]];
}
HD void sample_texture(const TexCtx& c, int idx, float u, float v, float out[4]) {
  const DTexture t = c.textures[idx];
  const bool clampEdge = (t.formatClamp >> 8) != 0u;
  const float x = xsub(xmul(u, (float)t.width), 0.5f), y = xsub(xmul(v, (float)t.height), 0.5f);
  const float fx0 = floorf(x), fy0 = floorf(y);
  const float fx = xsub(x, fx0), fy = xsub(y, fy0);
  const int ix = (int)fx0, iy = (int)fy0;
  const int x0 = tex_wrap(ix, (int)t.width, clampEdge), x1 = tex_wrap(ix + 1, (int)t.width, clampEdge);
  const int y0 = tex_wrap(iy, (int)t.height, clampEdge), y1 = tex_wrap(iy + 1, (int)t.height, clampEdge);
  float t00[4], t10[4], t01[4], t11[4];
  tex_fetch(c, t, x0, y0, t00); tex_fetch(c, t, x1, y0, t10); tex_fetch(c, t, x0, y1, t01); tex_fetch(c, t, x1, y1, t11);
  const float w00 = xmul(xsub(1.0f, fx), xsub(1.0f, fy)), w10 = xmul(fx, xsub(1.0f, fy)), w01 = xmul(xsub(1.0f, fx), fy), w11 = xmul(fx, fy);
  for (int k = 0; k < 4; ++k) out[k] = xadd(xadd(xadd(xmul(t00[k], w00), xmul(t10[k], w10)), xmul(t01[k], w01)), xmul(t11[k], w11));
}
// getSurfaceDataFromMat (bindless.glsl:181-237) as the probe rchit calls it (rchit:77: TBN and tangent are zero, so the
// normal-map branch never runs; discardPixel is ignored).  `color` = linearised vertex colour.
struct Surface { V3 albedo; float metallic, roughness; V3 emissive; };
HD bool material_textured(const anerend_material_info_t* m) { return m->bindless_indices[0] != -1 || m->bindless_indices[1] != -1 || m->bindless_indices[3] != -1; }
HD Surface surface_from_material(const TexCtx& c, const anerend_material_info_t* m, float u, float v, V3 color) {
  Surface s;
  s.albedo = color;
  const int metRoughIdx = m->bindless_indices[0], albedoIdx = m->bindless_indices[1], emissiveIdx = m->bindless_indices[3];
  float t[4];
  if (albedoIdx != -1) {
    sample_texture(c, albedoIdx, u, v, t);
    s.albedo = mk3(xmul(m->base_col_fac[0], t[0]), xmul(m->base_col_fac[1], t[1]), xmul(m->base_col_fac[2], t[2]));
  } else if (m->base_col_fac[0] + m->base_col_fac[1] + m->base_col_fac[2] > 0.1f) s.albedo = mk3(m->base_col_fac[0], m->base_col_fac[1], m->base_col_fac[2]);
  if (metRoughIdx != -1) { sample_texture(c, metRoughIdx, u, v, t); s.metallic = t[1]; s.roughness = t[0]; }
  else { s.metallic = m->met_rough[0]; s.roughness = m->met_rough[1]; }
  if (emissiveIdx != -1) {
    sample_texture(c, emissiveIdx, u, v, t);
    s.emissive = mk3(xmul(xmul(t[0], m->emissive[0]), m->emissive[3]), xmul(xmul(t[1], m->emissive[1]), m->emissive[3]), xmul(xmul(t[2], m->emissive[2]), m->emissive[3]));
  } else s.emissive = mk3(m->emissive[0], m->emissive[1], m->emissive[2]) * m->emissive[3];
  return s;
}

// pbr_light.glsl:10-43
HD float distribution_ggx(V3 N, V3 Hh, float roughness) {
  // denom cancels to ~a2 when N.H -> 1 (a2 = 0.0016 at roughness 0.2): kept in the oracle's exact sequence
  float a = xmul(roughness, roughness), a2 = xmul(a, a);
  float NdotH = fmaxf(dot3x(N, Hh), 0.0f), NdotH2 = xmul(NdotH, NdotH);
  float num = xadd(a2, 0.001f);
  float denom = xadd(xmul(NdotH2, xsub(a2, 1.0f)), 1.0f);
  denom = xmul(xmul(DDGI_PI, denom), denom);
  return xdiv(num, denom);
}
HD float geometry_schlick_ggx(float NdotV, float roughness) {
  float r = (roughness + 1.0f), k = (r * r) / 8.0f;
  return NdotV / (NdotV * (1.0f - k) + k);
}
HD float geometry_smith(V3 N, V3 V, V3 L, float roughness) {
  float NdotV = fmaxf(dot3(N, V), 0.0f), NdotL = fmaxf(dot3(N, L), 0.0f);
  float ggx2 = geometry_schlick_ggx(NdotV, roughness), ggx1 = geometry_schlick_ggx(NdotL, roughness);
  return ggx1 * ggx2;
}
// pbr_light.glsl:98-166.  directional: the sun (L, radiance given); point light: caller passes L and radiance too.
HD V3 calc_light(V3 N, V3 albedo, float metallic, float roughness, V3 worldPos, V3 viewPos, V3 L, V3 radiance) {
  V3 V = normalize3x(mk3(xsub(viewPos.x, worldPos.x), xsub(viewPos.y, worldPos.y), xsub(viewPos.z, worldPos.z)));
  V3 F0 = mix3(mk3(0.04f), albedo, metallic);
  V3 Hh = normalize3x(mk3(xadd(V.x, L.x), xadd(V.y, L.y), xadd(V.z, L.z)));
  float NDF = distribution_ggx(N, Hh, roughness);
  float G = geometry_smith(N, V, L, roughness);
  V3 F = F0 + (mk3(1.0f) - F0) * pow5(clampf(1.0f - fmaxf(dot3(Hh, V), 0.0f), 0.0f, 1.0f));   // fresnelSchlick :88-91
  V3 kD = mk3(1.0f) - F; kD = kD * (1.0f - metallic);
  float NdotL = fmaxf(dot3(N, L), 0.0f), NdotV = fmaxf(dot3(N, V), 0.0f);
  V3 numerator = NDF * G * F;
  float denominator = 4.0f * NdotV * NdotL + 0.0001f;
  V3 specular = clamp3(numerator / denominator, 0.0f, 1.0f);
  return mk3(0.0f) + (kD * albedo / DDGI_PI + specular) * radiance * NdotL;
}

// texel fetch + manual bilinear: VK_FILTER_LINEAR with REPEAT addressing (PipelineUtil.h:98-112), fp32 weights
HD int wrap_repeat(int x, int n) {           // REPEAT addressing; taps of the path stay within one period of [0, n)
  if ((unsigned)x >= (unsigned)n) { x -= n; if ((unsigned)x >= (unsigned)n) { x %= n; if (x < 0) x += n; } }
  return x;
}
HD V3 atlas_texel(const TraceCtx& c, int x, int y) {     // x, y already wrapped
#ifdef __CUDA_ARCH__
  uint2 t = __ldg(&c.atlas[(size_t)y * c.W + x]);
#else
  uint2 t = c.atlas[(size_t)y * c.W + x];
#endif
  return mk3(h2f((uint16_t)(t.x & 0xFFFFu)), h2f((uint16_t)(t.x >> 16)), h2f((uint16_t)(t.y & 0xFFFFu)));
}
// One axis of sampleSingleProbe (pbr_light.glsl:168-188) + the bilinear footprint along it: `start` = first interior
// texel of the probe on this axis, `size` = atlas extent, o = octEncode(normal)*0.5+0.5 on this axis.
// texel coordinates reach 10*nXZ*nY (2880 at C3): one ulp there is 2.4e-4 texels, so the address math stays exact
struct AxisTap { int i0, i1; float f; };
HD AxisTap probe_axis_tap(int start, int size, float o, int clampSampler = 0) {
  int end = start + 8 - 1;
  float ts = xdiv(xadd(xmul(2.0f, (float)start), 1.0f), xmul(2.0f, (float)size));
  float te = xdiv(xadd(xmul(2.0f, (float)end), 1.0f), xmul(2.0f, (float)size));
  float uv = xadd(ts, xmul(xsub(te, ts), o));
  float x = xsub(xmul(uv, (float)size), 0.5f);
  float f0 = floorf(x);
  AxisTap t; t.f = xsub(x, f0); int i = (int)f0;
  if (clampSampler) { t.i0 = clampi(i, 0, size - 1); t.i1 = clampi(i + 1, 0, size - 1); }      // VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE
  else { t.i0 = wrap_repeat(i, size); t.i1 = wrap_repeat(i + 1, size); }
  return t;
}
HD V3 atlas_bilinear(const TraceCtx& c, const AxisTap& ax, const AxisTap& ay) {
  V3 t00 = atlas_texel(c, ax.i0, ay.i0), t10 = atlas_texel(c, ax.i1, ay.i0), t01 = atlas_texel(c, ax.i0, ay.i1), t11 = atlas_texel(c, ax.i1, ay.i1);
  float fx = ax.f, fy = ay.f;
  float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
  return t00 * w00 + t10 * w10 + t01 * w01 + t11 * w11;
}
// the same footprint on a bindless texture (a baked tile atlas)
HD V3 texture_bilinear_taps(const TexCtx& c, int idx, const AxisTap& ax, const AxisTap& ay) {
  const DTexture t = c.textures[idx];
  float t00[4], t10[4], t01[4], t11[4];
  tex_fetch(c, t, ax.i0, ay.i0, t00); tex_fetch(c, t, ax.i1, ay.i0, t10); tex_fetch(c, t, ax.i0, ay.i1, t01); tex_fetch(c, t, ax.i1, ay.i1, t11);
  const float fx = ax.f, fy = ay.f;
  const float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
  return mk3(t00[0], t00[1], t00[2]) * w00 + mk3(t10[0], t10[1], t10[2]) * w10 + mk3(t01[0], t01[1], t01[2]) * w01 + mk3(t11[0], t11[1], t11[2]) * w11;
}
// pbr_light.glsl:240-363 (live-grid branch :294-308) + weightProbe :190-238 + sampleSingleProbe :168-188.
// octEncode(normalize(normal)) is the same for all 8 probes, and the atlas column of a tap depends only on the corner's
// x, its row only on (y, z): the per-axis address arithmetic is evaluated once per distinct value (2 + 4 instead of 8 + 8);
// every value is produced by the same operations as in the shader, so the result is unchanged.
//
// TILES (a tileInfoBuffer was uploaded): worldPosToTileInfo (bindless.glsl:291-308, only the upper bounds of the page are
// checked, as in the shader) finds the paged tile of worldPos; if it carries a baked atlas (ddgiAtlasTex >= 0) the grid is
// placed on that 32 m tile with step (32/nXZ, 1, 32/nXZ) (pbr_light.glsl:251-273).  WHICH texture is sampled is the caller's
// choice, as in the shaders: the probe and reflection hit shaders always pass the live atlas (srcTex < 0), the deferred pass
// passes the tile's baked atlas (deferred_tiled.comp:264-296).
#define ANEREND_TILE_PAGE_RADIUS 10
#define ANEREND_TILE_SIZE 32
HD int world_pos_to_tile_atlas(const TraceCtx& c, V3 worldPos, int& tx, int& tz) {
  const int w = ANEREND_TILE_PAGE_RADIUS * 2 + 1;
  tx = (int)floorf(worldPos.x / (float)ANEREND_TILE_SIZE); tz = (int)floorf(worldPos.z / (float)ANEREND_TILE_SIZE);
  const int cx = (int)floorf(c.cameraPos.x / (float)ANEREND_TILE_SIZE), cz = (int)floorf(c.cameraPos.z / (float)ANEREND_TILE_SIZE);
  if (tx - cx > ANEREND_TILE_PAGE_RADIUS || tz - cz > ANEREND_TILE_PAGE_RADIUS) return -1;
  const int idx = (tz - cz + ANEREND_TILE_PAGE_RADIUS) * w + (tx - cx + ANEREND_TILE_PAGE_RADIUS);
  if (idx < 0 || (uint32_t)idx >= c.tex.numTiles) return -1;            // the shader would read out of bounds here: treated as "no tile"
  const int atlasIdx = c.tex.tileInfos[idx];
  return (atlasIdx >= 0 && (uint32_t)atlasIdx < c.tex.numTextures) ? atlasIdx : -1;
}
template <bool TILES = false>
HD V3 sample_probe(const TraceCtx& c, V3 worldPos, V3 normal, int srcTex = -1) {
  V3 step = c.step, camOff = c.camOff;
  if (TILES) {
    int tx, tz;
    if (world_pos_to_tile_atlas(c, worldPos, tx, tz) >= 0) {
      const float s = (float)ANEREND_TILE_SIZE / (float)c.nXZ;
      step = mk3(s, 1.0f, s);
      camOff = mk3((float)tx * (float)ANEREND_TILE_SIZE, 0.0f, (float)tz * (float)ANEREND_TILE_SIZE);
    }
  }
  int px = clampi((int)floorf((worldPos.x - camOff.x) / step.x), 0, c.nXZ - 1);
  int py = clampi((int)floorf(worldPos.y / step.y), 0, c.nY - 1);
  int pz = clampi((int)floorf((worldPos.z - camOff.z) / step.z), 0, c.nXZ - 1);
  V3 base = mk3((float)px, (float)py, (float)pz) * step + camOff;
  V3 alpha = clamp3((worldPos - base) / step, 0.0f, 1.0f);
  float o0, o1; oct_encode(normalize3x(normal), o0, o1);
  o0 = xmul(xadd(o0, 1.0f), 0.5f); o1 = xmul(xadd(o1, 1.0f), 0.5f);
  // texel footprint of the taps: in the size of the texture that is sampled (textureSize(probeTex, 0), pbr_light.glsl:170)
  int srcW = c.W, srcH = c.H, srcClamp = c.clampSampler;
  if (TILES && srcTex >= 0) { const DTexture st = c.tex.textures[srcTex]; srcW = (int)st.width; srcH = (int)st.height; srcClamp = (int)(st.formatClamp >> 8); }
  AxisTap tapX[2], tapY[4];
  for (int ox = 0; ox < 2; ++ox) tapX[ox] = probe_axis_tap((8 + 2) * (px + ox) + 1, srcW, o0, srcClamp);
  for (int oy = 0; oy < 2; ++oy) for (int oz = 0; oz < 2; ++oz)
    tapY[oy + 2 * oz] = probe_axis_tap((8 + 2) * c.nXZ * (py + oy) + (8 + 2) * (pz + oz) + 1, srcH, o1, srcClamp);
  float sumW = 0.0f; V3 sum = mk3(0.0f);
  // corner order of the source: 000,111,110,100,011,010,001,101 packed as x | y<<1 | z<<2 per nibble
  const uint32_t order = 0x0u | (0x7u << 4) | (0x3u << 8) | (0x1u << 12) | (0x6u << 16) | (0x2u << 20) | (0x4u << 24) | (0x5u << 28);
  for (int k = 0; k < 8; ++k) {
    uint32_t o = (order >> (4 * k)) & 7u; int ox = (int)(o & 1u), oy = (int)((o >> 1) & 1u), oz = (int)((o >> 2) & 1u);
    V3 probePos = (k == 0) ? base : base + mk3((float)ox, (float)oy, (float)oz) * step;
    V3 tri = mk3(mixf(1.0f - alpha.x, alpha.x, (float)ox), mixf(1.0f - alpha.y, alpha.y, (float)oy), mixf(1.0f - alpha.z, alpha.z, (float)oz));
    float weight = 1.0f;
    V3 toProbe = normalize3(probePos - worldPos);
    float wb = fmaxf(0.0001f, (dot3(toProbe, normal) + 1.0f) * 0.5f);
    weight *= wb * wb + 0.2f;
    weight = fmaxf(0.000001f, weight);
    if (weight < 0.2f) weight *= weight * weight * (1.0f / (0.2f * 0.2f));    // dead code kept verbatim (:221-224)
    weight *= tri.x * tri.y * tri.z;
    V3 irr = (TILES && srcTex >= 0) ? texture_bilinear_taps(c.tex, srcTex, tapX[ox], tapY[oy + 2 * oz]) : atlas_bilinear(c, tapX[ox], tapY[oy + 2 * oz]);
    irr = mk3(sqrtf(irr.x), sqrtf(irr.y), sqrtf(irr.z));
    sum = sum + weight * irr; sumW += weight;
  }
  V3 net = sum / sumW;
  net = net * net;
  return net * .5f * 3.1415926f;
}
template <bool TILES = false>
HD V3 calc_indirect(const TraceCtx& c, V3 normal, V3 albedo, float metallic, float roughness, V3 worldPos, V3 viewPos, int srcTex = -1) {   // :365-388
  V3 V = normalize3(viewPos - worldPos);
  V3 F0 = mix3(mk3(0.04f), albedo, metallic);
  float rr = roughness * roughness;
  V3 kS = F0 + (max3(mk3(1.0f - rr), F0) - F0) * pow5(clampf(1.0f - fmaxf(dot3(normal, V), 0.0f), 0.0f, 1.0f));   // fresnelSchlickRoughness :93-96
  V3 kD = mk3(1.0f) - kS; kD = kD * (1.0f - metallic);
  V3 irradiance = sample_probe<TILES>(c, worldPos, normal, srcTex);
  return kD * (irradiance * albedo);
}

// Atlas consumer: calcIndirectDiffuseLight for a batch of surface points, what deferred_tiled.comp:264-296
// (CLAMP_TO_EDGE sampler, DeferredLightingRenderPass.cpp:192-200) and mirror_reflections.rchit:92-118 (REPEAT) evaluate per
// pixel / per hit against "ProbeRaysConvTex".  Point record: 3 x float4 = (pos.xyz, n.x) (n.yz, albedo.xy) (albedo.z, metallic, roughness, -).
struct ConsumerCtx { TraceCtx t; const float4* points; float4* out; V3 viewPos; };
HD void indirect_diffuse_body(uint32_t i, const ConsumerCtx& c) {
  const float4 a = c.points[3 * (size_t)i], b = c.points[3 * (size_t)i + 1], d = c.points[3 * (size_t)i + 2];
  const V3 wp = mk3(a.x, a.y, a.z);
  V3 r;
  if (c.t.tex.tileInfos) {
    // deferred_tiled.comp:264-296 samples the tile's baked atlas where there is one (the pass with the CLAMP_TO_EDGE probeTex);
    // mirror_reflections.rchit:105-114 always samples the live atlas
    int tx, tz;
    const int src = c.t.clampSampler ? world_pos_to_tile_atlas(c.t, wp, tx, tz) : -1;
    r = calc_indirect<true>(c.t, mk3(a.w, b.x, b.y), mk3(b.z, b.w, d.x), d.y, d.z, wp, c.viewPos, src);
  } else r = calc_indirect<false>(c.t, mk3(a.w, b.x, b.y), mk3(b.z, b.w, d.x), d.y, d.z, wp, c.viewPos);
  c.out[i] = make_float4(r.x, r.y, r.z, 0.0f);
}

// ShadowRT: shadow.rgen:30-82 + shadow.rmiss for one pixel — world position from the depth buffer, one
// any-hit ray towards -lightDir (NOT normalised, as in the shader), r32f result: 1 = lit or sky, 0 = shadowed.
struct ShadowRtCtx {
  const uint4* nodes; const float4* bvhTris; uint32_t numNodes;
  const InstXform* xforms; uint32_t twoLevel;      // two-level structure (traverse2)
  const float* depth; float* out; int width, height;
  float invViewProj[16]; V3 lightDir; float tmin, tmax; float screenW, screenH;
};
// shadow.rgen:20-28 calcWorldPos with pixelToUv (helpers.glsl:1-4); mat4 * vec4 evaluated c0*x + c1*y + c2*z + c3*w
HD V3 shadow_rt_world_pos(const ShadowRtCtx& c, int px, int py, float depthSamp) {
  const float u = (2.0f * (float)px + 1.0f) / (2.0f * c.screenW), v = (2.0f * (float)py + 1.0f) / (2.0f * c.screenH);
  const float cx = u * 2.0f - 1.0f, cy = v * 2.0f - 1.0f, cz = depthSamp;
  const float* m = c.invViewProj;
  const float X = m[0] * cx + m[4] * cy + m[8] * cz + m[12] * 1.0f, Y = m[1] * cx + m[5] * cy + m[9] * cz + m[13] * 1.0f;
  const float Z = m[2] * cx + m[6] * cy + m[10] * cz + m[14] * 1.0f, W = m[3] * cx + m[7] * cy + m[11] * cz + m[15] * 1.0f;
  return mk3(X / W, Y / W, Z / W);
}
HD void shadow_rt_body(uint32_t i, const ShadowRtCtx& c) {
  const int px = (int)(i % (uint32_t)c.width), py = (int)(i / (uint32_t)c.width);
  const float depthSamp = c.depth[i];                      // texel centre of a depth image of the launch size
  if (depthSamp == 1.0f) { c.out[i] = 1.0f; return; }      // sky (shadow.rgen:45-48)
  const V3 origin = shadow_rt_world_pos(c, px, py, depthSamp);
  Hit sh; uint32_t cn = 0, ct = 0;
  const bool occluded = c.twoLevel ? traverse2<true, false>(c.nodes, c.bvhTris, c.numNodes, c.xforms, origin, -c.lightDir, c.tmin, c.tmax, sh, cn, ct)
                                   : traverse<true, false>(c.nodes, c.bvhTris, c.numNodes, origin, -c.lightDir, c.tmin, c.tmax, sh, cn, ct);
  c.out[i] = occluded ? 0.0f : 1.0f;                       // payload starts at 0, shadow.rmiss sets 1
}

// rchit main for one hit (irradiance_probe_update.rchit:24-111; SPECULAR: mirror_reflections.rchit:24-118, which differs in the
// normal transform and in the flag that gates the indirect term).  origin = gl_WorldRayOriginEXT.
template <bool CNT, bool SPECULAR>
HD V3 shade_hit_simple(const TraceCtx& c, const Hit& hit, V3 origin, uint32_t& lit, uint32_t& nShadow, uint32_t& nPre, uint32_t& nCand, uint32_t* cn, uint32_t* ct, V3& hitPos) {
  const uint32_t ii = c.twoLevel ? hit.inst : c.triInstance[hit.prim];
  const DInstance in = c.instances[ii];
  const InstXform& xf = c.xforms[ii];
  const uint32_t prim = hit.prim - in.firstTri;                                            // gl_PrimitiveID
  const anerend_vertex_t* v0 = &c.vertices[in.vertexOffset + c.indices[in.indexOffset + 3 * prim + 0]];
  const anerend_vertex_t* v1 = &c.vertices[in.vertexOffset + c.indices[in.indexOffset + 3 * prim + 1]];
  const anerend_vertex_t* v2 = &c.vertices[in.vertexOffset + c.indices[in.indexOffset + 3 * prim + 2]];
  const float b0 = 1.0f - hit.u - hit.v, b1 = hit.u, b2 = hit.v;
  V3 pos = mk3(v0->pos[0], v0->pos[1], v0->pos[2]) * b0 + mk3(v1->pos[0], v1->pos[1], v1->pos[2]) * b1 + mk3(v2->pos[0], v2->pos[1], v2->pos[2]) * b2;
  pos = xform_point(xf.o2w, pos);                                                          // rchit:42
  V3 normal = normalize3(mk3(v0->normal[0], v0->normal[1], v0->normal[2]) * b0 + mk3(v1->normal[0], v1->normal[1], v1->normal[2]) * b1 +
                         mk3(v2->normal[0], v2->normal[1], v2->normal[2]) * b2);
  if (!SPECULAR)
    normal = normalize3(mk3(normal.x * xf.w2o[0] + normal.y * xf.w2o[1] + normal.z * xf.w2o[2],
                            normal.x * xf.w2o[3] + normal.y * xf.w2o[4] + normal.z * xf.w2o[5],
                            normal.x * xf.w2o[6] + normal.y * xf.w2o[7] + normal.z * xf.w2o[8]));   // rchit:73: normal * gl_WorldToObjectEXT
  else                                                                                           // mirror_reflections.rchit:74: mat3(gl_ObjectToWorldEXT) * normal
    normal = normalize3(mk3(xf.o2w[0] * normal.x + xf.o2w[3] * normal.y + xf.o2w[6] * normal.z,
                            xf.o2w[1] * normal.x + xf.o2w[4] * normal.y + xf.o2w[7] * normal.z,
                            xf.o2w[2] * normal.x + xf.o2w[5] * normal.y + xf.o2w[8] * normal.z));
  V3 color = mk3(v0->color[0], v0->color[1], v0->color[2]) * b0 + mk3(v1->color[0], v1->color[1], v1->color[2]) * b1 +
             mk3(v2->color[0], v2->color[1], v2->color[2]) * b2;
  color = mk3(to_linear1(color.x), to_linear1(color.y), to_linear1(color.z));              // rchit:75
  // getSurfaceDataFromMat (bindless.glsl:181-237); uv as rchit:70
  const anerend_material_info_t* m = &c.materials[in.material];
  const float tu = v0->uv[0] * b0 + v1->uv[0] * b1 + v2->uv[0] * b2, tv = v0->uv[1] * b0 + v1->uv[1] * b1 + v2->uv[1] * b2;
  const Surface sf = surface_from_material(c.tex, m, tu, tv, color);
  const V3 albedo = sf.albedo, emissive = sf.emissive;
  const float metallic = sf.metallic, roughness = sf.roughness;
  // sun shadow ray (rchit:45-66): any-hit, origin = hit position without offset
  const V3 sunL = normalize3(-c.lightDir);
  Hit sh;
  bool occluded = true;
  if (!(c.flags & ANEREND_DDGI_FLAG_SKIP_DEAD_SHADOW_RAYS) || dot3(normal, sunL) > 0.0f) {
    const SmEntry sme = sunmap_lookup(c, pos);
    const float sPos = dot3(pos, c.smS);
    if (sunmap_unoccluded(sme, sPos, c.tmin)) { ++nPre; occluded = false; }        // answered by the sun's map, not traced
    else {
      ++nShadow;
      if (sunmap_candidate_occludes(c, sme, pos, sunL, c.tmin, c.tmax)) ++nCand;
      else occluded = traverse_scene<true, CNT>(c, pos, sunL, c.tmin, sunmap_tmax(sme, sPos, c.tmax), sh, cn[1], ct[1]);
    }
  }
  V3 irradiance = mk3(0.0f);
  if (!occluded) { lit |= 1u; irradiance = calc_light(normal, albedo, metallic, roughness, pos, origin, sunL, c.sunIntensity * mk3(1.0f) * 1.0f); }
  if (c.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) {
    // extension E1: enabled point lights, calcLight(directional = 0) gated by an any-hit ray pos -> light
    for (uint32_t li = 0; li < c.numLights; ++li) {
      const anerend_light_t* lt = &c.lights[li];
      if (!(lt->color[3] > 0.5f)) continue;
      const V3 lp = mk3(lt->world_pos[0], lt->world_pos[1], lt->world_pos[2]);
      const V3 dv = lp - pos; const float dist = sqrtf(dot3(dv, dv)), range = lt->world_pos[3];
      if (!(dist < range)) continue;
      const V3 L = normalize3(lp - pos);
      if (!(dot3(normal, L) > 0.0f)) continue;
      const LmEntry lme = lightmap_lookup(c, li, lp, pos);
      if (lightmap_unoccluded(lme, dist, c.tmin)) ++nPre;      // answered by the light map, not traced
      else {
        ++nShadow;
        if (lightmap_candidate_occludes(c, lme, pos, L, c.tmin, dist)) { ++nCand; continue; }
        if (traverse_scene<true, CNT>(c, pos, L, c.tmin, lightmap_tmax(lme, dist), sh, cn[1], ct[1])) continue;
      }
      lit |= 1u << (1u + li);
      const float dd = fminf(dist, range);
      const float att = (1.0f - dd / range) * (1.0f - dd / range) * 1.0f;
      irradiance = irradiance + calc_light(normal, albedo, metallic, roughness, pos, origin, L, mk3(lt->color[0], lt->color[1], lt->color[2]) * att);
    }
  }
  const uint32_t gate = SPECULAR ? ANEREND_UBO_DDGI_FLAG : ANEREND_UBO_DDGI_MULTI_FLAG;      // mirror_reflections.rchit:105 / rchit:98
  if ((c.uboFlags & gate) == gate)
    irradiance = irradiance + (c.tex.tileInfos ? calc_indirect<true>(c, normal, albedo, metallic, roughness, pos, origin)
                                               : calc_indirect<false>(c, normal, albedo, metallic, roughness, pos, origin));   // rchit:98-107
  hitPos = pos;
  return irradiance + emissive;                                                                      // rchit:109
}

// SpecularGIRT: mirror_reflections.rgen:30-81 + .rchit + .rmiss for one pixel — reflect the view vector about
// the G-buffer normal, closest hit on the path's BVH, shade the hit like the probe rchit (sun shadow ray, calcLight, indirect
// diffuse from the atlas, emissive); rgba16f result (irradiance, hit distance), (sky, 1e6) on a miss, 0 for sky pixels.
struct SpecularCtx {
  TraceCtx t;                         // scene, BVH, atlas, lights, flags; camOff/step as sampleProbe places the grid
  const float* depth; const float4* geom0; uint2* out; int width, height;
  float invViewProj[16]; V3 cameraPos; float screenW, screenH;
};
HD void specular_gi_body(uint32_t i, const SpecularCtx& c) {
  const int px = (int)(i % (uint32_t)c.width), py = (int)(i / (uint32_t)c.width);
  const float depthSamp = c.depth[i];
  if (depthSamp == 1.0f) { c.out[i] = make_uint2(0u, 0u); return; }                              // rgen:41-44
  ShadowRtCtx w{};                                                                               // calcWorldPos is the same function (rgen:22-30)
  for (int k = 0; k < 16; ++k) w.invViewProj[k] = c.invViewProj[k];
  w.screenW = c.screenW; w.screenH = c.screenH;
  const V3 pos = shadow_rt_world_pos(w, px, py, depthSamp);
  const float4 g0 = c.geom0[i];
  const V3 normal = normalize3(mk3(g0.x, g0.y, g0.z));                                           // rgen:46
  const V3 vv = normalize3(pos - c.cameraPos);
  const float dn = dot3(normal, vv);
  const V3 dir = vv - (2.0f * dn) * normal;                                                      // reflect(v, normal)
  Hit hit; uint32_t cn[2] = {0, 0}, ct[2] = {0, 0}, lit = 0, nShadow = 0, nPre = 0, nCand = 0;
  V3 rad; float dist;
  if (traverse_scene<false, false>(c.t, pos, dir, 0.1f, 50.0f, hit, cn[0], ct[0])) {
    V3 hp;
    rad = shade_hit_simple<false, true>(c.t, hit, pos, lit, nShadow, nPre, nCand, cn, ct, hp);
    const V3 dd = hp - pos; dist = sqrtf(dot3(dd, dd));                                          // distance(pos, gl_WorldRayOriginEXT)
  } else { rad = c.t.sky; dist = 1e6f; }                                                         // mirror_reflections.rmiss
  c.out[i] = make_uint2((uint32_t)f2h(rad.x) | ((uint32_t)f2h(rad.y) << 16), (uint32_t)f2h(rad.z) | ((uint32_t)f2h(dist) << 16));
}

// rgen main + rchit main + rmiss main for one probe ray (rgen:48-111, rchit:24-111, rmiss:17-21)
template <bool CNT>
HD void trace_ray_body(uint32_t idx, const TraceCtx& c) {
  const int P = c.nXZ * c.nXZ;
  const uint32_t slot = idx % (uint32_t)c.N, probeLin = idx / (uint32_t)c.N;
  const int layerLocal = (int)(probeLin / (uint32_t)c.Pl), probe = stripe_probe((int)(probeLin % (uint32_t)c.Pl), c.nXZ, c.rowOffset, c.rowStride);
  const int layer = c.layerMap[layerLocal];
  const uint32_t ray = c.rayOrder ? (uint32_t)c.rayOrder[slot] : slot;
  const int probeX = probe % c.nXZ, probeZ = probe / c.nXZ;
  const V3 origin = mk3((float)probeX, (float)layer, (float)probeZ) * c.step + c.camOff;      // rgen:83
  const float4 d4 = c.dirs[(size_t)layer * c.N + ray];
  const V3 d = mk3(d4.x, d4.y, d4.z);
  uint32_t cn[2] = {0, 0}, ct[2] = {0, 0}, nShadow = 0, nPre = 0, nCand = 0, lit = 0;
  Hit hit;
  V3 rad;
  if (traverse_scene<false, CNT>(c, origin, d, c.tmin, c.tmax, hit, cn[0], ct[0])) {
    V3 hitPos;
    rad = shade_hit_simple<CNT, false>(c, hit, origin, lit, nShadow, nPre, nCand, cn, ct, hitPos);
  } else {
    rad = c.sky;                                                                                      // rmiss:20
  }
  const size_t o = ((size_t)layer * P + probe) * c.N + ray;
  c.rayIrr[o] = make_uint2((uint32_t)f2h(rad.x) | ((uint32_t)f2h(rad.y) << 16), (uint32_t)f2h(rad.z) | (0x3C00u << 16));   // rgen:109
  if (c.records) {
    anerend_ddgi_ray_record_t r;
    r.prim = hit.prim; r.t = hit.prim == DDGI_MISS ? 0.0f : hit.t; r.u = hit.u; r.v = hit.v;
    r.radiance[0] = rad.x; r.radiance[1] = rad.y; r.radiance[2] = rad.z; r.lit_mask = lit;
    c.records[o] = r;
  }
  if (CNT && c.stats) {
    atomAdd64(&c.stats[0], 1ull); atomAdd64(&c.stats[1], hit.prim != DDGI_MISS ? 1ull : 0ull); atomAdd64(&c.stats[2], nShadow);
    atomAdd64(&c.stats[3], cn[0]); atomAdd64(&c.stats[4], ct[0]); atomAdd64(&c.stats[5], cn[1]); atomAdd64(&c.stats[6], ct[1]);
    if (nPre) atomAdd64(&c.stats[7], nPre);
    if (nCand) atomAdd64(&c.stats[8], nCand);
  }
}

// ----------------------------------------------------------------------------------------------------
// probe_conv.comp main (probe_conv.comp:21-90), live mode.  Thread per interior atlas texel.
// ----------------------------------------------------------------------------------------------------
struct ConvCtx {
  const uint2* rayIrr; const float4* dirs;   // dirs: [nY][N] fp32 by absolute layer, rounded to fp16 here like the rgba16f store (rgen:108)
  const int* layerMap;                       // [numLayers] absolute layer of each launched layer
  uint2* atlas; float4* texelRecords;        // atlas written in place (or to every peer's copy, see peers)
  uint2* peers[8]; int numPeers;
  int nXZ, nY, N, Sx, Sy, W, H, numLayers;
  int Pl, rowOffset, rowStride;              // row stripe, as in TraceCtx
  int bakeMode;                              // UBO_BAKE_MODE_ON_FLAG: running weighted mean instead of the hysteresis blend (probe_conv.comp:67-79)
  float hysteresis;
};
HD void conv_texel_body(uint32_t idx, const ConvCtx& c) {
  const int P = c.nXZ * c.nXZ;
  const int texel = (int)(idx & 63u); const uint32_t probeLin = idx >> 6;
  const int layerLocal = (int)(probeLin / (uint32_t)c.Pl), probe = stripe_probe((int)(probeLin % (uint32_t)c.Pl), c.nXZ, c.rowOffset, c.rowStride), layer = c.layerMap[layerLocal];
  const int i = texel & 7, j = texel >> 3, probeX = probe % c.nXZ, probeZ = probe / c.nXZ;
  float ux = (2.0f * (float)i + 1.0f) / (2.0f * 8.0f), uy = (2.0f * (float)j + 1.0f) / (2.0f * 8.0f);     // pixelToUv, helpers.glsl:1-4
  ux = ux * 2.0f - 1.0f; uy = uy * 2.0f - 1.0f;
  const V3 octDir = oct_decode(ux, uy);
  const uint2* irr = c.rayIrr + ((size_t)layer * P + probe) * c.N;
  const float4* dirs = c.dirs + (size_t)layer * c.N;
  float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
  for (int x = 0; x < c.Sx; ++x) for (int y = 0; y < c.Sy; ++y) {        // x outer, y inner (probe_conv.comp:49-50)
    const int ray = y * c.Sx + x;
    const uint2 rr = irr[ray]; const float4 d4 = dirs[ray];
    const V3 rad = mk3(h2f((uint16_t)(rr.x & 0xFFFFu)), h2f((uint16_t)(rr.x >> 16)), h2f((uint16_t)(rr.y & 0xFFFFu)));
    const V3 d = mk3(h2f(f2h(d4.x)), h2f(f2h(d4.y)), h2f(f2h(d4.z)));
    const float w = fmaxf(0.0f, dot3(d, octDir));
    if (w > 1e-6f) { r0 += w * rad.x; r1 += w * rad.y; r2 += w * rad.z; r3 += w; }
  }
  const int ax = 10 * probeX + 1 + i, ay = 10 * c.nXZ * layer + 10 * probeZ + 1 + j;       // pixOffset + pixel (:42,:64)
  const size_t ti = (size_t)ay * c.W + ax;
  const uint2 ot = c.atlas[ti];
  const float o0 = h2f((uint16_t)(ot.x & 0xFFFFu)), o1 = h2f((uint16_t)(ot.x >> 16)), o2 = h2f((uint16_t)(ot.y & 0xFFFFu)), o3 = h2f((uint16_t)(ot.y >> 16));
  float n0, n1, n2, n3;
  if (c.bakeMode) {
    const float totalWeight = o3 + r3;
    float t0 = r0 + o0 * o3, t1 = r1 + o1 * o3, t2 = r2 + o2 * o3;
    if (totalWeight > 1e-6f) { t0 /= totalWeight; t1 /= totalWeight; t2 /= totalWeight; }
    n0 = t0; n1 = t1; n2 = t2; n3 = totalWeight;
  } else {
    if (r3 > 1e-6f) { r0 /= r3; r1 /= r3; r2 /= r3; }
    const float hy = c.hysteresis;
    n0 = r0 * (1.0f - hy) + o0 * hy; n1 = r1 * (1.0f - hy) + o1 * hy; n2 = r2 * (1.0f - hy) + o2 * hy; n3 = r3 * (1.0f - hy) + o3 * hy;
  }
  const uint2 nt = make_uint2((uint32_t)f2h(n0) | ((uint32_t)f2h(n1) << 16), (uint32_t)f2h(n2) | ((uint32_t)f2h(n3) << 16));
  c.atlas[ti] = nt;
  for (int p = 0; p < c.numPeers; ++p) c.peers[p][ti] = nt;               // fused all-gather: store into every peer's atlas
  if (c.texelRecords) c.texelRecords[ti] = make_float4(n0, n1, n2, n3);
}

// ----------------------------------------------------------------------------------------------------
// IrradianceProbeTranslationPass.cpp:92-168: second vkCmdCopyImage (TempProbeTex -> atlas, shifted by one probe)
// ----------------------------------------------------------------------------------------------------
struct TransCtx { const uint2* temp; uint2* atlas; int W, nXZ, nY, dx, dz, ew, eh; };
HD void translate_body(uint32_t idx, const TransCtx& c) {
  const int x = (int)(idx % (uint32_t)c.ew); const uint32_t r = idx / (uint32_t)c.ew;
  const int y = (int)(r % (uint32_t)c.eh), layer = (int)(r / (uint32_t)c.eh);
  const int pps = 10, planeY = layer * pps * c.nXZ;
  int sx = 0, sy = planeY, dxo = 0, dyo = planeY;
  if (c.dx > 0) sx = pps; else if (c.dx < 0) dxo = pps;
  if (c.dz > 0) sy += pps; else if (c.dz < 0) dyo += pps;
  c.atlas[(size_t)(dyo + y) * c.W + dxo + x] = c.temp[(size_t)(sy + y) * c.W + sx + x];
}
struct ClearCtx { uint2* atlas; uint2* temp; };
HD void clear_body(uint32_t idx, const ClearCtx& c) {     // clearConvTex: (0,0,0,1), IrradianceProbeConvolvePass.cpp:20-24
  c.atlas[idx] = make_uint2(0u, 0x3C000000u);
  c.temp[idx] = make_uint2(0u, 0x3C000000u);
}

}  // namespace ddgi
