// ddgi_oracle.cpp — CPU oracle for AneRend's DDGI irradiance-probe update.
//
// TEST INFRASTRUCTURE ONLY (see ddgi_oracle.h).  The reference has no golden vectors, tests or CPU path for this chain;
// the oracle is held to the reference's COMPILED shaders (assets/*.spv executed by tests/golden/spirv_interp.py, vectors in
// tests/golden/) for everything those binaries compute.  PARITY UNPINNED for what they leave to the Vulkan driver — the
// ray/triangle intersection and traversal — which is pinned by KATs (tests/test_oracle_kat.py) only.
//
// Every function below restates one reference function and cites it (paths relative to the AneRend repo;
// shaders live in assets/shaders/).  Arithmetic is IEEE fp32 evaluated strictly left to right, compiled with
// -ffp-contract=off; the only fused multiply-adds are the explicit fmaf calls in tri_intersect().  What the
// reference leaves to the Vulkan driver (ray/triangle test, BVH, fp16 store rounding, bilinear weights,
// transcendental ULPs) is DEFINED here and documented in DESIGN.md §"Defined arithmetic".
#include "ddgi_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// tiny vector maths (GLSL semantics, component-wise, no contraction)
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3(float s) { return V3{s, s, s}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// GLSL normalize(): defined as v * (1/sqrt(dot(v,v))) with IEEE sqrt and division.
inline V3 normalize(V3 v) { float inv = 1.0f / std::sqrt(dot(v, v)); return v * inv; }
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline V3 clamp3(V3 v, float lo, float hi) { return V3{clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)}; }
inline V3 max3(V3 a, V3 b) { return V3{std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
// GLSL mix(x, y, a) = x*(1-a) + y*a
inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline V3 mix3(V3 x, V3 y, float a) { return V3{mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)}; }
// pow(x, 5.0) on a value clamped to [0,1]: defined as ((x*x)*(x*x))*x
inline float pow5(float x) { float x2 = x * x; return (x2 * x2) * x; }

const float PI_F = 3.14159265359f;             // pbr_light.glsl:8, rgen:14
const float GOLDEN_RATIO_F = 1.6180339887f;    // rgen:13

// ------------------------------------------------------------------------------------------------
// fp16 storage (rgba16f images): software round-to-nearest-even, independent of F16C
// ------------------------------------------------------------------------------------------------
inline uint16_t f2h(float f) {
  uint32_t x; std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t em = x & 0x7FFFFFFFu;
  if (em >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((em > 0x7F800000u) ? (0x0200u | ((em >> 13) & 0x3FFu)) : 0u));
  if (em >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);               // rounds to >= 65520 -> inf
  if (em < 0x33000001u) return (uint16_t)sign;                            // < 2^-25 (or == 2^-25 tie -> even 0)
  int32_t e = (int32_t)(em >> 23) - 127;
  uint32_t m = (em & 0x7FFFFFu) | 0x800000u;
  int shift; uint32_t base;
  if (e < -14) { shift = 13 + (-14 - e); base = 0; }                      // subnormal half
  else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7FFFFFu; }
  uint32_t q = m >> shift;
  uint32_t rem = m & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  uint32_t r = base + q;
  if (rem > half || (rem == half && (r & 1u))) r += 1;                     // carries propagate into the exponent
  return (uint16_t)(sign | r);
}
inline float h2f(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1Fu, m = h & 0x3FFu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { int s = 0; while (!(m & 0x400u)) { m <<= 1; ++s; } m &= 0x3FFu; x = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | (m << 13); }
  } else if (e == 31) x = sign | 0x7F800000u | (m << 13);
  else x = sign | ((e - 15 + 127) << 23) | (m << 13);
  float f; std::memcpy(&f, &x, 4); return f;
}

// ------------------------------------------------------------------------------------------------
// shader helper functions
// ------------------------------------------------------------------------------------------------
// octahedron_helpers.glsl:1-7
inline float signNotZero(float k) { return (k >= 0.0f) ? 1.0f : -1.0f; }
// octahedron_helpers.glsl:10-20
inline void octEncode(V3 v, float out[2]) {
  V3 tmp = v; v.y = tmp.z; v.z = tmp.y;
  float l1norm = std::fabs(v.x) + std::fabs(v.y) + std::fabs(v.z);
  float inv = 1.0f / l1norm;
  float rx = v.x * inv, ry = v.y * inv;
  if (v.z < 0.0f) {
    float nx = (1.0f - std::fabs(ry)) * signNotZero(rx);
    float ny = (1.0f - std::fabs(rx)) * signNotZero(ry);
    rx = nx; ry = ny;
  }
  out[0] = rx; out[1] = ry;
}
// octahedron_helpers.glsl:25-34
inline V3 octDecode(float ox, float oy) {
  V3 v = v3(ox, oy, 1.0f - std::fabs(ox) - std::fabs(oy));
  if (v.z < 0.0f) {
    float nx = (1.0f - std::fabs(v.y)) * signNotZero(v.x);
    float ny = (1.0f - std::fabs(v.x)) * signNotZero(v.y);
    v.x = nx; v.y = ny;
  }
  V3 tmp = v; v.y = tmp.z; v.z = tmp.y;
  return normalize(v);
}
// bindless.glsl:240-247 (per channel)
inline float toLinear1(float c) {
  float higher = std::pow((c + 0.055f) / 1.055f, 2.4f);
  float lower = c / 12.92f;
  return (c < 0.04045f) ? lower : higher;
}
inline V3 toLinear(V3 c) { return V3{toLinear1(c.x), toLinear1(c.y), toLinear1(c.z)}; }

// rgen:35-46 worldDirFromFibonacciIndex
inline V3 fibonacciDir(uint32_t idx, uint32_t n) {
  float theta = 2.0f * PI_F * float(idx) / GOLDEN_RATIO_F;
  float phi = std::acos(1.0f - 2.0f * (float(idx) + 0.5f) / float(n));
  V3 o = v3(std::cos(theta) * std::sin(phi), std::sin(theta) * std::sin(phi), std::cos(phi));
  return normalize(o);
}

// glm::rotate(mat4(1), radians(deg), axis) — contrib/glm/include/glm/ext/matrix_transform.inl:18-46,
// called from IrradianceProbeRayTracingPass.cpp:133.  Column-major, out[c*4+r].
void makeRotation(int axisIdx, float deg, float out[16]) {
  float a = deg * 0.01745329251994329576923690768489f;   // glm::radians
  float c = std::cos(a), s = std::sin(a);
  float axis[3] = {0, 0, 0}; axis[axisIdx] = 1.0f;
  float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
  float R[3][3];
  R[0][0] = c + temp[0] * axis[0]; R[0][1] = temp[0] * axis[1] + s * axis[2]; R[0][2] = temp[0] * axis[2] - s * axis[1];
  R[1][0] = temp[1] * axis[0] - s * axis[2]; R[1][1] = c + temp[1] * axis[1]; R[1][2] = temp[1] * axis[2] + s * axis[0];
  R[2][0] = temp[2] * axis[0] + s * axis[1]; R[2][1] = temp[2] * axis[1] - s * axis[0]; R[2][2] = c + temp[2] * axis[2];
  // Result[i] = m[0]*R[i][0] + m[1]*R[i][1] + m[2]*R[i][2] with m = identity
  const float m[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int r = 0; r < 4; ++r)
      out[i * 4 + r] = m[0][r] * R[i][0] + m[1][r] * R[i][1] + m[2][r] * R[i][2];
  for (int r = 0; r < 4; ++r) out[12 + r] = m[3][r];
}

// mat3(rot) * d  (rgen:88): c0*d.x + c1*d.y + c2*d.z
inline V3 mulMat3(const float* m, V3 d) {
  return V3{m[0] * d.x + m[4] * d.y + m[8] * d.z,
            m[1] * d.x + m[5] * d.y + m[9] * d.z,
            m[2] * d.x + m[6] * d.y + m[10] * d.z};
}
// gl_ObjectToWorldEXT * vec4(p, 1) with the renderable's column-major transform (rchit:42, tlas_update.comp:69-83)
inline V3 xformPoint(const float* m, V3 p) {
  return V3{m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
            m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
            m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
}

// ------------------------------------------------------------------------------------------------
// Ray/triangle test — the part the reference leaves to the RT hardware (rgen:96, rchit:55).
// Defined arithmetic: Möller–Trumbore on (v0, e1=v1-v0, e2=v2-v0), two-sided, with explicit
// fused multiply-adds in this exact association; accept tmin < t < tmax (both exclusive, as Vulkan does).
// ------------------------------------------------------------------------------------------------
inline float dotF(V3 a, V3 b) { return std::fmaf(a.z, b.z, std::fmaf(a.y, b.y, a.x * b.x)); }
inline V3 crossF(V3 a, V3 b) {
  return V3{std::fmaf(a.y, b.z, -(a.z * b.y)), std::fmaf(a.z, b.x, -(a.x * b.z)), std::fmaf(a.x, b.y, -(a.y * b.x))};
}
inline bool tri_intersect(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tmin, float tmax, float& t, float& u, float& v) {
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

struct WTri { V3 v0, e1, e2; };          // world-space, as flattened
struct Instance {                         // one (renderable, mesh) pair accepted by tlas_update.comp:37-99
  uint32_t renderable, mesh, material, firstTri, numTris;
  float o2w[16];                          // renderable transform, column-major
  float w2o[9];                           // inverse of the upper 3x3, column-major: w2o[c*3+r]
};

// inverse of the upper-left 3x3 (gl_WorldToObjectEXT's 3x3 part) via adjugate / determinant
void invert3x3(const float* m, float* out) {
  float a = m[0], b = m[4], c = m[8];     // row 0
  float d = m[1], e = m[5], f = m[9];     // row 1
  float g = m[2], h = m[6], i = m[10];    // row 2
  float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  float det = a * A + b * B + c * C;
  float inv = 1.0f / det;
  // inverse rows
  float r00 = A * inv, r01 = -(b * i - c * h) * inv, r02 = (b * f - c * e) * inv;
  float r10 = B * inv, r11 = (a * i - c * g) * inv, r12 = -(a * f - c * d) * inv;
  float r20 = C * inv, r21 = -(a * h - b * g) * inv, r22 = (a * e - b * d) * inv;
  out[0] = r00; out[1] = r10; out[2] = r20;   // column 0
  out[3] = r01; out[4] = r11; out[5] = r21;   // column 1
  out[6] = r02; out[7] = r12; out[8] = r22;   // column 2
}

// ------------------------------------------------------------------------------------------------
// the oracle's own BVH (binary, binned SAH).  Traversal order never matters: closest hit = min (t, id).
// ------------------------------------------------------------------------------------------------
struct BNode { float lo[3], hi[3]; uint32_t left, count; };   // count>0: leaf, left = first index into order[]

struct Bvh {
  std::vector<BNode> nodes;
  std::vector<uint32_t> order;
  void build(const std::vector<WTri>& tris) {
    const uint32_t n = (uint32_t)tris.size();
    nodes.clear(); order.resize(n);
    if (n == 0) return;
    std::vector<float> blo(3 * (size_t)n), bhi(3 * (size_t)n), cen(3 * (size_t)n);
    for (uint32_t i = 0; i < n; ++i) {
      const WTri& t = tris[i];
      V3 a = t.v0, b = t.v0 + t.e1, c = t.v0 + t.e2;
      float ax[3] = {a.x, a.y, a.z}, bx[3] = {b.x, b.y, b.z}, cx[3] = {c.x, c.y, c.z};
      for (int k = 0; k < 3; ++k) {
        float lo = std::fmin(ax[k], std::fmin(bx[k], cx[k])), hi = std::fmax(ax[k], std::fmax(bx[k], cx[k]));
        // pad: e1/e2 round-trips (v0+e1 != v1 exactly) and slab-test rounding must never cull a true hit
        float pad = 1e-5f * std::fmax(1.0f, std::fmax(std::fabs(lo), std::fabs(hi)));
        blo[3 * (size_t)i + k] = lo - pad; bhi[3 * (size_t)i + k] = hi + pad; cen[3 * (size_t)i + k] = 0.5f * (lo + hi);
      }
      order[i] = i;
    }
    nodes.reserve(2 * (size_t)n);
    nodes.push_back(BNode{});
    struct Task { uint32_t node, first, count; };
    std::vector<Task> stack; stack.push_back({0, 0, n});
    const int NB = 16;
    while (!stack.empty()) {
      Task tk = stack.back(); stack.pop_back();
      float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f}, clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
      for (uint32_t i = tk.first; i < tk.first + tk.count; ++i) {
        size_t t = order[i];
        for (int k = 0; k < 3; ++k) {
          lo[k] = std::fmin(lo[k], blo[3 * t + k]); hi[k] = std::fmax(hi[k], bhi[3 * t + k]);
          clo[k] = std::fmin(clo[k], cen[3 * t + k]); chi[k] = std::fmax(chi[k], cen[3 * t + k]);
        }
      }
      BNode nd; for (int k = 0; k < 3; ++k) { nd.lo[k] = lo[k]; nd.hi[k] = hi[k]; }
      auto makeLeaf = [&]() { nd.left = tk.first; nd.count = tk.count; nodes[tk.node] = nd; };
      if (tk.count <= 2) { makeLeaf(); continue; }
      // binned SAH over the three axes
      int bestAxis = -1, bestSplit = -1; float bestCost = 1e30f;
      for (int ax = 0; ax < 3; ++ax) {
        float ext = chi[ax] - clo[ax];
        if (!(ext > 0.0f)) continue;
        float scale = NB / ext;
        int cnt[NB] = {0}; float bl[NB][3], bh[NB][3];
        for (int b = 0; b < NB; ++b) for (int k = 0; k < 3; ++k) { bl[b][k] = 1e30f; bh[b][k] = -1e30f; }
        for (uint32_t i = tk.first; i < tk.first + tk.count; ++i) {
          size_t t = order[i];
          int b = std::min(NB - 1, (int)((cen[3 * t + ax] - clo[ax]) * scale));
          cnt[b]++;
          for (int k = 0; k < 3; ++k) { bl[b][k] = std::fmin(bl[b][k], blo[3 * t + k]); bh[b][k] = std::fmax(bh[b][k], bhi[3 * t + k]); }
        }
        float rA[NB]; int rC[NB];
        { float l[3] = {1e30f, 1e30f, 1e30f}, h[3] = {-1e30f, -1e30f, -1e30f}; int c = 0;
          for (int b = NB - 1; b > 0; --b) {
            for (int k = 0; k < 3; ++k) { l[k] = std::fmin(l[k], bl[b][k]); h[k] = std::fmax(h[k], bh[b][k]); }
            c += cnt[b]; float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
            rA[b] = c ? (dx * dy + dy * dz + dz * dx) : 0.0f; rC[b] = c; } }
        { float l[3] = {1e30f, 1e30f, 1e30f}, h[3] = {-1e30f, -1e30f, -1e30f}; int c = 0;
          for (int b = 0; b < NB - 1; ++b) {
            for (int k = 0; k < 3; ++k) { l[k] = std::fmin(l[k], bl[b][k]); h[k] = std::fmax(h[k], bh[b][k]); }
            c += cnt[b]; if (c == 0 || rC[b + 1] == 0) continue;
            float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
            float cost = (dx * dy + dy * dz + dz * dx) * c + rA[b + 1] * rC[b + 1];
            if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = b; } } }
      }
      uint32_t mid;
      if (bestAxis < 0) {
        if (tk.count <= 8) { makeLeaf(); continue; }
        mid = tk.first + tk.count / 2;   // all centroids coincide: split by index
      } else {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        float leafCost = (dx * dy + dy * dz + dz * dx) * tk.count;
        if (tk.count <= 4 && leafCost <= bestCost + 1.0f * (dx * dy + dy * dz + dz * dx)) { makeLeaf(); continue; }
        float scale = NB / (chi[bestAxis] - clo[bestAxis]);
        auto* b = order.data() + tk.first; auto* e = b + tk.count;
        auto* m = std::partition(b, e, [&](uint32_t t) {
          return std::min(NB - 1, (int)((cen[3 * (size_t)t + bestAxis] - clo[bestAxis]) * scale)) <= bestSplit; });
        mid = tk.first + (uint32_t)(m - b);
        if (mid == tk.first || mid == tk.first + tk.count) mid = tk.first + tk.count / 2;
      }
      uint32_t l = (uint32_t)nodes.size();
      nodes.push_back(BNode{}); nodes.push_back(BNode{});
      nd.left = l; nd.count = 0; nodes[tk.node] = nd;
      stack.push_back({l, tk.first, mid - tk.first});
      stack.push_back({l + 1, mid, tk.first + tk.count - mid});
    }
  }
};

struct RayCtx { V3 o, d, inv; };
inline bool boxHit(const BNode& n, const RayCtx& r, float tmin, float tmax, float& tn) {
  const float o[3] = {r.o.x, r.o.y, r.o.z}, iv[3] = {r.inv.x, r.inv.y, r.inv.z};
  float t0 = tmin, t1 = tmax;
  for (int k = 0; k < 3; ++k) {
    float a = (n.lo[k] - o[k]) * iv[k], b = (n.hi[k] - o[k]) * iv[k];
    float nr = std::fmin(a, b), fr = std::fmax(a, b);          // fmin/fmax drop NaN (0*inf)
    t0 = std::fmax(t0, nr); t1 = std::fmin(t1, fr * 1.0000004f);
  }
  tn = t0; return t0 <= t1;
}

struct Hit { uint32_t prim; float t, u, v; };

}  // namespace

// ================================================================================================
struct ddgi_oracle {
  anerend_ddgi_params_t p{};
  std::string err;
  int threads = 1;
  bool brute = false;
  int nXZ = 0, nY = 0, N = 0, Sx = 0, Sy = 0, W = 0, H = 0;

  // scene (bindless set 0)
  std::vector<anerend_vertex_t> vertices;
  std::vector<uint32_t> indices, meshNumIndices, modelIdx, matIdx;
  std::vector<anerend_mesh_info_t> meshes;
  std::vector<anerend_material_info_t> materials;
  std::vector<anerend_renderable_t> renderables;
  std::vector<anerend_light_t> lights;
  anerend_scene_ubo_t ubo{};
  bool haveScene = false, haveUbo = false, haveAccel = false;

  // acceleration structure
  std::vector<Instance> instances;
  std::vector<uint32_t> triInstance;      // global tri id -> instance
  std::vector<WTri> tris;
  Bvh bvh;

  // pass state
  int currentLayer = 0;                   // IrradianceProbeRayTracingPass::_currentLayer
  int lastProbeIdx[2] = {0, 0};           // IrradianceProbeTranslationPass::_lastProbeIdx
  std::vector<uint16_t> atlas, temp;      // ProbeRaysConvTex, TempProbeTex (RGBA16F)
  std::vector<float> fibDirs;             // N*3 host table
  std::vector<std::vector<uint16_t>> rayIrr;   // per layer: [probe][ray][4]  (ProbeRaysIrrTex contents)
  std::vector<std::vector<uint16_t>> rayDir;   // per layer: [probe][ray][4]  (ProbeRaysDirTex contents)
  std::vector<std::vector<anerend_ddgi_ray_record_t>> records;
  std::vector<float> texelRecords;
  std::atomic<uint64_t> st[7];

  template <class F> void parallelFor(int n, F&& f) {
    int nt = std::max(1, std::min(threads, n));
    if (nt == 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; ++t) ts.emplace_back([&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& t : ts) t.join();
  }

  void clearAtlas() {   // clearConvTex, IrradianceProbeConvolvePass.cpp:9-47: (0,0,0,1)
    atlas.assign((size_t)W * H * 4, 0);
    for (size_t i = 0; i < (size_t)W * H; ++i) atlas[4 * i + 3] = 0x3C00;
    temp = atlas;
    if (p.flags & ANEREND_DDGI_FLAG_TEXEL_RECORDS) texelRecords.assign((size_t)W * H * 4, 0.0f);
  }

  // ---- instance selection + flatten: tlas_update.comp:37-99 -------------------------------------
  int buildInstances() {
    instances.clear(); triInstance.clear(); tris.clear();
    uint32_t firstTri = 0;
    for (uint32_t idx = 0; idx < renderables.size(); ++idx) {
      const anerend_renderable_t& r = renderables[idx];
      bool visible = (r.visible & 1u) != 0;            // visibleBitSet, bindless.glsl:260-263
      bool culled = ((r.visible >> 1) & 1u) != 0;      // cullBitSet,    bindless.glsl:249-252
      if (!visible || culled) continue;
      for (uint32_t i = 0; i < r.num_meshes; ++i) {
        uint32_t mIdx = r.model_offset + i;
        if (mIdx >= modelIdx.size()) { err = "model index out of range"; return ANEREND_DDGI_E_INVALID; }
        uint32_t meshIndex = modelIdx[mIdx];
        uint32_t matAbstract = r.first_material_index + i;
        if (matAbstract >= matIdx.size()) { err = "material index out of range"; return ANEREND_DDGI_E_INVALID; }
        uint32_t matIndex = matIdx[matAbstract];
        if (r.dynamic_model_offset > 0) {
          if (r.dynamic_model_offset + i >= modelIdx.size()) { err = "dynamic model index out of range"; return ANEREND_DDGI_E_INVALID; }
          meshIndex = modelIdx[r.dynamic_model_offset + i];
        }
        // customIndex packing: 10 bits material, 14 bits mesh (tlas_update.comp:88-91; rchit:27-28)
        uint32_t custom = matIndex; custom = (custom & ~(0x3FFFu << 10)) | ((meshIndex & 0x3FFFu) << 10);
        Instance in{}; in.renderable = idx; in.material = custom & 0x3FFu; in.mesh = (custom >> 10) & 0x3FFFu;
        if (in.mesh >= meshes.size() || in.material >= materials.size()) { err = "mesh/material id out of range"; return ANEREND_DDGI_E_INVALID; }
        in.firstTri = firstTri; in.numTris = meshNumIndices[in.mesh] / 3;
        std::memcpy(in.o2w, r.transform, 64); invert3x3(in.o2w, in.w2o);
        firstTri += in.numTris; instances.push_back(in);
      }
    }
    tris.resize(firstTri); triInstance.resize(firstTri);
    for (uint32_t ii = 0; ii < instances.size(); ++ii) {
      const Instance& in = instances[ii]; const anerend_mesh_info_t& mi = meshes[in.mesh];
      for (uint32_t pr = 0; pr < in.numTris; ++pr) {
        V3 w[3];
        for (int k = 0; k < 3; ++k) {
          size_t ix = (size_t)mi.index_offset + 3 * (size_t)pr + k;
          if (ix >= indices.size()) { err = "index out of range"; return ANEREND_DDGI_E_INVALID; }
          size_t vi = (size_t)mi.vertex_offset + indices[ix];
          if (vi >= vertices.size()) { err = "vertex out of range"; return ANEREND_DDGI_E_INVALID; }
          const float* ps = vertices[vi].pos; w[k] = xformPoint(in.o2w, v3(ps[0], ps[1], ps[2]));
        }
        tris[in.firstTri + pr] = WTri{w[0], w[1] - w[0], w[2] - w[0]};
        triInstance[in.firstTri + pr] = ii;
      }
    }
    return 0;
  }

  // ---- traversal ----------------------------------------------------------------------------------
  bool closestHit(V3 o, V3 d, float tmin, float tmax, Hit& best, uint64_t& nodesV, uint64_t& trisT) const {
    best.prim = 0xFFFFFFFFu; best.t = tmax; best.u = best.v = 0.0f;
    auto test = [&](uint32_t id) {
      const WTri& tr = tris[id]; float t, u, v; ++trisT;
      if (tri_intersect(o, d, tr.v0, tr.e1, tr.e2, tmin, tmax, t, u, v))
        if (t < best.t || (t == best.t && id < best.prim)) { best.prim = id; best.t = t; best.u = u; best.v = v; }
    };
    if (brute) { for (uint32_t i = 0; i < tris.size(); ++i) test(i); return best.prim != 0xFFFFFFFFu; }
    if (bvh.nodes.empty()) return false;
    RayCtx r{o, d, v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z)};
    struct SE { uint32_t node; float tn; };
    SE stack[128]; int sp = 0; float tn;
    if (!boxHit(bvh.nodes[0], r, tmin, tmax, tn)) return false;
    stack[sp++] = SE{0, tn};
    while (sp) {
      SE e = stack[--sp];
      // best.t is inclusive here so that equal-t ties in other subtrees are still examined
      if (e.tn > best.t) continue;
      const BNode& n = bvh.nodes[e.node]; ++nodesV;
      if (n.count) { for (uint32_t i = 0; i < n.count; ++i) test(bvh.order[n.left + i]); continue; }
      float ta, tb;
      bool ha = boxHit(bvh.nodes[n.left], r, tmin, best.t, ta), hb = boxHit(bvh.nodes[n.left + 1], r, tmin, best.t, tb);
      if (ha && hb) {
        if (ta <= tb) { stack[sp++] = SE{n.left + 1, tb}; stack[sp++] = SE{n.left, ta}; }
        else { stack[sp++] = SE{n.left, ta}; stack[sp++] = SE{n.left + 1, tb}; }
      } else if (ha) stack[sp++] = SE{n.left, ta};
      else if (hb) stack[sp++] = SE{n.left + 1, tb};
    }
    return best.prim != 0xFFFFFFFFu;
  }
  bool anyHit(V3 o, V3 d, float tmin, float tmax, uint64_t& nodesV, uint64_t& trisT) const {
    float t, u, v;
    if (brute) { for (const WTri& tr : tris) { ++trisT; if (tri_intersect(o, d, tr.v0, tr.e1, tr.e2, tmin, tmax, t, u, v)) return true; } return false; }
    if (bvh.nodes.empty()) return false;
    RayCtx r{o, d, v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z)};
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
      const BNode& n = bvh.nodes[stack[--sp]]; float tn; ++nodesV;
      if (!boxHit(n, r, tmin, tmax, tn)) continue;
      if (n.count) { for (uint32_t i = 0; i < n.count; ++i) { const WTri& tr = tris[bvh.order[n.left + i]]; ++trisT;
                       if (tri_intersect(o, d, tr.v0, tr.e1, tr.e2, tmin, tmax, t, u, v)) return true; } }
      else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
    }
    return false;
  }

  // ---- light-space grid for the sun's shadow rays (PROTOTYPE of DESIGN.md §8 item 1; off unless set_sun_grid) ----------------
  // All sun shadow rays share one direction, so in the sun's orthographic space a ray is a POINT (x, y) plus a depth: the triangles it
  // can hit are those whose footprint covers the point and whose depth interval reaches beyond the origin.  A G x G grid over the scene's
  // footprint holds per-cell triangle lists (footprints padded by one cell: the projection is rounded, the exact test is not); a query
  // runs the SAME tri_intersect on the candidates of its cell, so the answer — and with it lit mask and atlas — is bit-identical to the
  // hierarchy's whenever the lists are a superset, which tests/test_sun_grid.py checks.
  struct SunGrid {
    int G = 0; double ax[3], ay[3], az[3]; double x0 = 0, y0 = 0, cw = 1, ch = 1;
    std::vector<uint32_t> start, ids; std::vector<float> zmax;      // CSR lists + per-triangle far depth (towards the sun)
  } sun;
  int sunGridRes = 0;
  void buildSunGrid() {
    sun = SunGrid{};
    if (sunGridRes <= 0 || tris.empty() || !haveUbo) return;
    const int G = sunGridRes;
    double d[3] = {-(double)ubo.light_dir[0], -(double)ubo.light_dir[1], -(double)ubo.light_dir[2]};
    double dl = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); if (!(dl > 0)) return;
    for (int k = 0; k < 3; ++k) sun.az[k] = d[k] / dl;
    double up[3] = {0, 1, 0}; if (std::fabs(sun.az[1]) > 0.99) { up[0] = 1; up[1] = 0; }
    double a[3] = {sun.az[1] * up[2] - sun.az[2] * up[1], sun.az[2] * up[0] - sun.az[0] * up[2], sun.az[0] * up[1] - sun.az[1] * up[0]};
    double al = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    for (int k = 0; k < 3; ++k) sun.ax[k] = a[k] / al;
    sun.ay[0] = sun.az[1] * sun.ax[2] - sun.az[2] * sun.ax[1]; sun.ay[1] = sun.az[2] * sun.ax[0] - sun.az[0] * sun.ax[2]; sun.ay[2] = sun.az[0] * sun.ax[1] - sun.az[1] * sun.ax[0];
    auto proj = [&](V3 p, double& x, double& y, double& z) {
      x = sun.ax[0] * p.x + sun.ax[1] * p.y + sun.ax[2] * p.z; y = sun.ay[0] * p.x + sun.ay[1] * p.y + sun.ay[2] * p.z; z = sun.az[0] * p.x + sun.az[1] * p.y + sun.az[2] * p.z; };
    const size_t n = tris.size();
    std::vector<double> lox(n), loy(n), hix(n), hiy(n); sun.zmax.resize(n);
    double ex0 = 1e300, ey0 = 1e300, ex1 = -1e300, ey1 = -1e300;
    for (size_t i = 0; i < n; ++i) {
      const WTri& t = tris[i]; V3 vtx[3] = {t.v0, t.v0 + t.e1, t.v0 + t.e2};
      double zx = -1e300; lox[i] = loy[i] = 1e300; hix[i] = hiy[i] = -1e300;
      for (int k = 0; k < 3; ++k) { double x, y, z; proj(vtx[k], x, y, z);
        lox[i] = std::min(lox[i], x); hix[i] = std::max(hix[i], x); loy[i] = std::min(loy[i], y); hiy[i] = std::max(hiy[i], y); zx = std::max(zx, z); }
      sun.zmax[i] = (float)zx;
      ex0 = std::min(ex0, lox[i]); ex1 = std::max(ex1, hix[i]); ey0 = std::min(ey0, loy[i]); ey1 = std::max(ey1, hiy[i]);
    }
    sun.G = G; sun.x0 = ex0; sun.y0 = ey0; sun.cw = std::max((ex1 - ex0) / G, 1e-9); sun.ch = std::max((ey1 - ey0) / G, 1e-9);
    auto cellRange = [&](double lo, double hi, double o, double c, int& a0, int& a1) {
      a0 = std::min(std::max((int)std::floor((lo - o) / c) - 1, 0), G - 1); a1 = std::min(std::max((int)std::floor((hi - o) / c) + 1, 0), G - 1); };
    std::vector<uint32_t> count((size_t)G * G + 1, 0);
    for (size_t i = 0; i < n; ++i) { int x0c, x1c, y0c, y1c; cellRange(lox[i], hix[i], sun.x0, sun.cw, x0c, x1c); cellRange(loy[i], hiy[i], sun.y0, sun.ch, y0c, y1c);
      for (int y = y0c; y <= y1c; ++y) for (int x = x0c; x <= x1c; ++x) ++count[(size_t)y * G + x + 1]; }
    for (size_t c = 0; c < (size_t)G * G; ++c) count[c + 1] += count[c];
    sun.start = count; sun.ids.resize(count.back());
    std::vector<uint32_t> fill(sun.start.begin(), sun.start.end() - 1);
    for (size_t i = 0; i < n; ++i) { int x0c, x1c, y0c, y1c; cellRange(lox[i], hix[i], sun.x0, sun.cw, x0c, x1c); cellRange(loy[i], hiy[i], sun.y0, sun.ch, y0c, y1c);
      for (int y = y0c; y <= y1c; ++y) for (int x = x0c; x <= x1c; ++x) sun.ids[fill[(size_t)y * G + x]++] = (uint32_t)i; }   // ascending triangle id per cell
  }
  bool anyHitSun(V3 o, V3 d, float tmin, float tmax, uint64_t& cellsV, uint64_t& trisT) const {
    const double x = sun.ax[0] * o.x + sun.ax[1] * o.y + sun.ax[2] * o.z, y = sun.ay[0] * o.x + sun.ay[1] * o.y + sun.ay[2] * o.z;
    const double z = sun.az[0] * o.x + sun.az[1] * o.y + sun.az[2] * o.z;
    const int cx = std::min(std::max((int)std::floor((x - sun.x0) / sun.cw), 0), sun.G - 1), cy = std::min(std::max((int)std::floor((y - sun.y0) / sun.ch), 0), sun.G - 1);
    const size_t c = (size_t)cy * sun.G + cx; ++cellsV;
    const float zgate = (float)(z + tmin) - 1e-3f * (1.0f + std::fabs((float)z));     // triangles entirely behind the origin + tMin cannot be hit
    float t, u, v;
    for (uint32_t k = sun.start[c]; k < sun.start[c + 1]; ++k) {
      const uint32_t id = sun.ids[k];
      if (sun.zmax[id] < zgate) continue;
      const WTri& tr = tris[id]; ++trisT;
      if (tri_intersect(o, d, tr.v0, tr.e1, tr.e2, tmin, tmax, t, u, v)) return true;
    }
    return false;
  }

  // ---- probe-grid addressing shared by rgen:75-83 and pbr_light.glsl:294-301 -------------------------
  // Bake mode (UBO_BAKE_MODE_ON_FLAG; rgen:61-73, pbr_light.glsl:274-291): the grid covers the tile bakeTileInfo.xy of size
  // bakeTileInfo.z: origin = tile corner, step = (tSize / nXZ, 1.0, tSize / nXZ) — the y step is 1, not PROBE_STEP.y.
  bool bakeMode() const { return (ubo.flags & ANEREND_UBO_BAKE_MODE_ON_FLAG) == ANEREND_UBO_BAKE_MODE_ON_FLAG; }
  V3 gridStep() const {
    if (bakeMode()) { float s = float(ubo.bake_tile_info[2]) / float(nXZ); return v3(s, 1.0f, s); }
    return v3(p.probe_step[0], p.probe_step[1], p.probe_step[2]);
  }
  V3 camOffset() const {
    if (bakeMode()) return v3(float(ubo.bake_tile_info[0]) * float(ubo.bake_tile_info[2]), 0.0f, float(ubo.bake_tile_info[1]) * float(ubo.bake_tile_info[2]));
    int cix = (int)(ubo.camera_pos[0] / p.probe_step[0]);   // C truncation, quirk Q1
    int ciz = (int)(ubo.camera_pos[2] / p.probe_step[2]);
    return v3(float(cix) * p.probe_step[0] - p.probe_step[0] * float(nXZ) / 2.0f, 0.0f,
              float(ciz) * p.probe_step[2] - p.probe_step[2] * float(nXZ) / 2.0f);
  }

  // ---- pbr_light.glsl:10-43,88-96 ---------------------------------------------------------------------
  static float DistributionGGX(V3 N, V3 Hh, float roughness) {
    float a = roughness * roughness, a2 = a * a;
    float NdotH = std::fmax(dot(N, Hh), 0.0f), NdotH2 = NdotH * NdotH;
    float num = a2 + 0.001f;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = PI_F * denom * denom;
    return num / denom;
  }
  static float GeometrySchlickGGX(float NdotV, float roughness) {
    float r = (roughness + 1.0f), k = (r * r) / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
  }
  static float GeometrySmith(V3 N, V3 V, V3 L, float roughness) {
    float NdotV = std::fmax(dot(N, V), 0.0f), NdotL = std::fmax(dot(N, L), 0.0f);
    float ggx2 = GeometrySchlickGGX(NdotV, roughness), ggx1 = GeometrySchlickGGX(NdotL, roughness);
    return ggx1 * ggx2;
  }
  static V3 fresnelSchlick(float cosTheta, V3 F0) { return F0 + (v3(1.0f) - F0) * pow5(clampf(1.0f - cosTheta, 0.0f, 1.0f)); }
  static V3 fresnelSchlickRoughness(float cosTheta, V3 F0, float roughness) {
    return F0 + (max3(v3(1.0f - roughness), F0) - F0) * pow5(clampf(1.0f - cosTheta, 0.0f, 1.0f));
  }
  // pbr_light.glsl:98-166 (useSpecular = true at the only call site on this path, rchit:85-94)
  static V3 calcLight(V3 normal, V3 albedo, float metallic, float roughness, V3 worldPos, V3 viewPos,
                      const anerend_light_t* light, int directional, V3 uboLightDir, float sunIntensity) {
    V3 lightColor = sunIntensity * v3(1.0f);
    V3 N = normal, V = normalize(viewPos - worldPos);
    V3 F0 = mix3(v3(0.04f), albedo, metallic);
    V3 Lo = v3(0.0f);
    const float brightness = 1.0f;
    float attenuation = 0.0f; V3 L = v3(0.0f);
    if (directional == 1) { attenuation = 1.0f; L = normalize(-uboLightDir); }
    else {
      V3 lp = v3(light->world_pos[0], light->world_pos[1], light->world_pos[2]);
      V3 dv = lp - worldPos; float dist = std::sqrt(dot(dv, dv));
      float d = std::fmin(dist, light->world_pos[3]);
      attenuation = (1.0f - d / light->world_pos[3]) * (1.0f - d / light->world_pos[3]) * brightness;
      L = normalize(lp - worldPos);
      lightColor = v3(light->color[0], light->color[1], light->color[2]);
    }
    V3 Hh = normalize(V + L);
    V3 radiance = lightColor * attenuation;
    float NDF = DistributionGGX(N, Hh, roughness);
    float G = GeometrySmith(N, V, L, roughness);
    V3 F = fresnelSchlick(std::fmax(dot(Hh, V), 0.0f), F0);
    V3 kS = F, kD = v3(1.0f) - kS; kD = kD * (1.0f - metallic);
    float NdotL = std::fmax(dot(N, L), 0.0f), NdotV = std::fmax(dot(N, V), 0.0f);
    V3 numerator = NDF * G * F;
    float denominator = 4.0f * NdotV * NdotL + 0.0001f;
    V3 specular = clamp3(numerator / denominator, 0.0f, 1.0f);
    Lo = Lo + (kD * albedo / PI_F + specular) * radiance * NdotL;
    return Lo;
  }

  // ---- bindless textures: texture(textures[idx], uv) in a ray-tracing stage = LOD 0 (no derivatives), VK_FILTER_LINEAR,
  // REPEAT or CLAMP_TO_EDGE per texture (asset/Texture.h:39, PipelineUtil.h:98-112).  Defined arithmetic: unnormalised
  // coordinate uv*size - 0.5, fract weights in fp32, UNORM8 -> c/255, sRGB8 rgb -> linear before filtering (alpha linear),
  // missing components (0, 0, 1), blend w00*t00 + w10*t10 + w01*t01 + w11*t11.
  struct Tex { std::vector<uint8_t> px; uint32_t w = 0, h = 0, fmt = 0; bool clampEdge = false; };
  std::vector<Tex> textures;
  float texLut[512];
  void makeTexLut() {
    for (int c = 0; c < 256; ++c) { float v = (float)c / 255.0f; texLut[c] = toLinear1(v); texLut[256 + c] = v; }
  }
  void texFetch(const Tex& t, int x, int y, float out[4]) const {
    const uint32_t bpp = t.fmt == ANEREND_TEX_RGBA16F ? 8u : t.fmt <= ANEREND_TEX_RGBA8_SRGB ? 4u : (t.fmt == ANEREND_TEX_RG8_UNORM ? 2u : 1u);
    const uint8_t* q = &t.px[((size_t)y * t.w + (size_t)x) * bpp];
    out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1.0f;
    if (t.fmt == ANEREND_TEX_RGBA16F) { uint16_t hh[4]; memcpy(hh, q, 8); for (int k = 0; k < 4; ++k) out[k] = h2f(hh[k]); return; }
    if (t.fmt == ANEREND_TEX_RGBA8_SRGB) { out[0] = texLut[q[0]]; out[1] = texLut[q[1]]; out[2] = texLut[q[2]]; out[3] = texLut[256 + q[3]]; }
    else for (uint32_t k = 0; k < bpp; ++k) out[k] = texLut[256 + q[k]];
  }
  static int texWrap(int i, int n, bool clampEdge) {
    if (clampEdge) return std::min(std::max(i, 0), n - 1);
    i %= n; return i < 0 ? i + n : i;
  }
  void sampleTexture(int idx, float u, float v, float out[4]) const {
    const Tex& t = textures[(size_t)idx];
    float x = u * float(t.w) - 0.5f, y = v * float(t.h) - 0.5f;
    float fx0 = std::floor(x), fy0 = std::floor(y);
    float fx = x - fx0, fy = y - fy0; int ix = (int)fx0, iy = (int)fy0;
    int x0 = texWrap(ix, (int)t.w, t.clampEdge), x1 = texWrap(ix + 1, (int)t.w, t.clampEdge);
    int y0 = texWrap(iy, (int)t.h, t.clampEdge), y1 = texWrap(iy + 1, (int)t.h, t.clampEdge);
    float t00[4], t10[4], t01[4], t11[4];
    texFetch(t, x0, y0, t00); texFetch(t, x1, y0, t10); texFetch(t, x0, y1, t01); texFetch(t, x1, y1, t11);
    float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
    for (int k = 0; k < 4; ++k) out[k] = t00[k] * w00 + t10[k] * w10 + t01[k] * w01 + t11[k] * w11;
  }

  // ---- atlas sampling: sampler2D probeTex with VK_FILTER_LINEAR, REPEAT (PipelineUtil.h:98-112) --------
  // Defined arithmetic: unnormalised coord = uv*size - 0.5, weights = fract, texel fetch with REPEAT wrap,
  // (1-fx)(1-fy)t00 + fx(1-fy)t10 + (1-fx)fy t01 + fx fy t11 in fp32 (Vulkan spec §16.8 formula).
  bool clampSampler = false;     // consumers only: VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE (DeferredLightingRenderPass.cpp:192-200)
  V3 texel(int x, int y) const {
    if (clampSampler) { x = std::min(std::max(x, 0), W - 1); y = std::min(std::max(y, 0), H - 1); }
    x %= W; if (x < 0) x += W; y %= H; if (y < 0) y += H;
    const uint16_t* t = &atlas[4 * ((size_t)y * W + x)];
    return v3(h2f(t[0]), h2f(t[1]), h2f(t[2]));
  }
  V3 textureBilinear(float u, float v) const {
    float x = u * float(W) - 0.5f, y = v * float(H) - 0.5f;
    float fx0 = std::floor(x), fy0 = std::floor(y);
    float fx = x - fx0, fy = y - fy0; int ix = (int)fx0, iy = (int)fy0;
    V3 t00 = texel(ix, iy), t10 = texel(ix + 1, iy), t01 = texel(ix, iy + 1), t11 = texel(ix + 1, iy + 1);
    float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
    return t00 * w00 + t10 * w10 + t01 * w01 + t11 * w11;
  }
  // pbr_light.glsl:168-188
  // probeTex: the live atlas (srcTex < 0) or a bindless texture (a tile's baked atlas, passed by the deferred pass)
  int srcTex = -1;
  V3 sampleSingleProbe(int pxI, int pyI, int pzI, V3 normal) const {
    int startX = (8 + 2) * pxI + 1, startY = (8 + 2) * nXZ * pyI + (8 + 2) * pzI + 1;
    int endX = startX + 8 - 1, endY = startY + 8 - 1;
    const float tw = srcTex >= 0 ? float(textures[(size_t)srcTex].w) : float(W), th = srcTex >= 0 ? float(textures[(size_t)srcTex].h) : float(H);   // textureSize(probeTex, 0)
    float tsx = (2.0f * float(startX) + 1.0f) / (2.0f * tw), tsy = (2.0f * float(startY) + 1.0f) / (2.0f * th);   // pixelToUv, helpers.glsl:1-4
    float tex = (2.0f * float(endX) + 1.0f) / (2.0f * tw), tey = (2.0f * float(endY) + 1.0f) / (2.0f * th);
    float oct[2]; octEncode(normalize(normal), oct);
    oct[0] = (oct[0] + 1.0f) * 0.5f; oct[1] = (oct[1] + 1.0f) * 0.5f;
    float cu = tsx + (tex - tsx) * oct[0], cv = tsy + (tey - tsy) * oct[1];
    if (srcTex >= 0) { float t4[4]; sampleTexture(srcTex, cu, cv, t4); return v3(t4[0], t4[1], t4[2]); }
    return textureBilinear(cu, cv);
  }
  // pbr_light.glsl:190-238
  V4 weightProbe(V3 worldPos, V3 probePos, V3 normal, int ix, int iy, int iz, V3 alpha, int ox, int oy, int oz) const {
    V3 tri = v3(mixf(1.0f - alpha.x, alpha.x, float(ox)), mixf(1.0f - alpha.y, alpha.y, float(oy)), mixf(1.0f - alpha.z, alpha.z, float(oz)));
    float weight = 1.0f;
    V3 trueDirectionToProbe = normalize(probePos - worldPos);
    float wb = std::fmax(0.0001f, (dot(trueDirectionToProbe, normal) + 1.0f) * 0.5f);
    weight *= wb * wb + 0.2f;
    weight = std::fmax(0.000001f, weight);
    const float crushThreshold = 0.2f;
    if (weight < crushThreshold) weight *= weight * weight * (1.0f / (crushThreshold * crushThreshold));   // dead code, quirk Q6
    weight *= tri.x * tri.y * tri.z;
    V3 irr = sampleSingleProbe(ix, iy, iz, normal);
    irr = v3(std::sqrt(irr.x), std::sqrt(irr.y), std::sqrt(irr.z));
    return V4{irr.x, irr.y, irr.z, weight};
  }
  // pbr_light.glsl:240-363, live-grid branch (:294-308); no baked tiles (quirk Q5), bake mode out of scope
  // tileInfoBuffer (bindless.glsl binding 14) + worldPosToTileInfo (bindless.glsl:291-308; only the upper page bounds are checked,
  // as in the shader; an index outside the uploaded buffer is "no tile"): bindless index of the tile's baked atlas, or -1
  std::vector<int32_t> tileInfos;
  int worldPosToTileAtlas(V3 worldPos, int& tx, int& tz) const {
    if (tileInfos.empty()) return -1;
    const int R = 10, TS = 32, w = R * 2 + 1;
    tx = (int)std::floor(worldPos.x / float(TS)); tz = (int)std::floor(worldPos.z / float(TS));
    int cx = (int)std::floor(ubo.camera_pos[0] / float(TS)), cz = (int)std::floor(ubo.camera_pos[2] / float(TS));
    if (tx - cx > R || tz - cz > R) return -1;
    int idx = (tz - cz + R) * w + (tx - cx + R);
    if (idx < 0 || (size_t)idx >= tileInfos.size()) return -1;
    int a = tileInfos[(size_t)idx];
    return (a >= 0 && (size_t)a < textures.size()) ? a : -1;
  }
  V3 sampleProbe(V3 worldPos, V3 normal) const {
    V3 step = gridStep();
    V3 co = camOffset();
    { int tx, tz;                                      // pbr_light.glsl:251-273: a baked tile places the grid on the tile
      if (worldPosToTileAtlas(worldPos, tx, tz) >= 0) { float s = 32.0f / float(nXZ); step = v3(s, 1.0f, s); co = v3(float(tx) * 32.0f, 0.0f, float(tz) * 32.0f); } }
    int px = std::min(std::max((int)std::floor((worldPos.x - co.x) / step.x), 0), nXZ - 1);
    int py = std::min(std::max((int)std::floor(worldPos.y / step.y), 0), nY - 1);
    int pz = std::min(std::max((int)std::floor((worldPos.z - co.z) / step.z), 0), nXZ - 1);
    V3 base = v3(float(px), float(py), float(pz)) * step + co;
    float sumW = 0.0f; V3 sum = v3(0.0f);
    V3 alpha = clamp3((worldPos - base) / step, 0.0f, 1.0f);
    static const int off[8][3] = {{0, 0, 0}, {1, 1, 1}, {1, 1, 0}, {1, 0, 0}, {0, 1, 1}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}};
    for (int c = 0; c < 8; ++c) {
      const int* o = off[c];
      V3 probePos = (c == 0) ? base : base + v3(float(o[0]), float(o[1]), float(o[2])) * step;
      V4 res = weightProbe(worldPos, probePos, normal, px + o[0], py + o[1], pz + o[2], alpha, o[0], o[1], o[2]);
      sum = sum + res.w * v3(res.x, res.y, res.z);
      sumW += res.w;
    }
    V3 net = sum / sumW;
    net = net * net;
    return net * .5f * 3.1415926f;
  }
  // pbr_light.glsl:365-388
  V3 calcIndirectDiffuseLight(V3 normal, V3 albedo, float metallic, float roughness, V3 worldPos, V3 viewPos) const {
    V3 V = normalize(viewPos - worldPos);
    V3 F0 = mix3(v3(0.04f), albedo, metallic);
    V3 kS = fresnelSchlickRoughness(std::fmax(dot(normal, V), 0.0f), F0, roughness * roughness);
    V3 kD = v3(1.0f) - kS; kD = kD * (1.0f - metallic);
    V3 irradiance = sampleProbe(worldPos, normal);
    V3 diffuse = irradiance * albedo;
    return kD * diffuse;
  }

  // ---- rchit main (irradiance_probe_update.rchit:24-111) -----------------------------------------------
  // specular = mirror_reflections.rchit:24-118: the same shader except for the normal transform (:74) and the flag gating the
  // indirect term (:105)
  V3 shadeHit(const Hit& hit, V3 rayOrigin, uint32_t& litMask, uint64_t* cnt, bool specular = false, V3* hitPos = nullptr) const {
    const Instance& in = instances[triInstance[hit.prim]];
    const anerend_mesh_info_t& mi = meshes[in.mesh];
    uint32_t prim = hit.prim - in.firstTri;                         // gl_PrimitiveID
    const anerend_vertex_t* vv[3];
    for (int k = 0; k < 3; ++k) vv[k] = &vertices[(size_t)mi.vertex_offset + indices[(size_t)mi.index_offset + 3 * (size_t)prim + k]];
    float b0 = 1.0f - hit.u - hit.v, b1 = hit.u, b2 = hit.v;
    auto P = [](const float* f) { return v3(f[0], f[1], f[2]); };
    V3 pos = P(vv[0]->pos) * b0 + P(vv[1]->pos) * b1 + P(vv[2]->pos) * b2;
    pos = xformPoint(in.o2w, pos);
    // shadow ray to the sun (rchit:45-66): any-hit semantics, same miss shader -> shadow = 0
    V3 lightDir = v3(ubo.light_dir[0], ubo.light_dir[1], ubo.light_dir[2]);
    V3 shadowDir = normalize(-lightDir);
    float uvx = vv[0]->uv[0] * b0 + vv[1]->uv[0] * b1 + vv[2]->uv[0] * b2, uvy = vv[0]->uv[1] * b0 + vv[1]->uv[1] * b1 + vv[2]->uv[1] * b2;   // rchit:70
    V3 normal = normalize(P(vv[0]->normal) * b0 + P(vv[1]->normal) * b1 + P(vv[2]->normal) * b2);
    // normal * gl_WorldToObjectEXT : row-vector times matrix = dot with each column of W2O's 3x3
    if (!specular)
      normal = normalize(v3(normal.x * in.w2o[0] + normal.y * in.w2o[1] + normal.z * in.w2o[2],
                            normal.x * in.w2o[3] + normal.y * in.w2o[4] + normal.z * in.w2o[5],
                            normal.x * in.w2o[6] + normal.y * in.w2o[7] + normal.z * in.w2o[8]));
    else      // mat3(gl_ObjectToWorldEXT) * normal: c0*n.x + c1*n.y + c2*n.z (o2w is the column-major mat4: column c at o2w[4c])
      normal = normalize(v3(in.o2w[0] * normal.x + in.o2w[4] * normal.y + in.o2w[8] * normal.z,
                            in.o2w[1] * normal.x + in.o2w[5] * normal.y + in.o2w[9] * normal.z,
                            in.o2w[2] * normal.x + in.o2w[6] * normal.y + in.o2w[10] * normal.z));
    if (hitPos) *hitPos = pos;
    // The reference traces the sun shadow ray before it knows the normal (rchit:45-66 precede :72-73); its result only
    // gates calcLight, which multiplies by max(N.L, 0) (pbr_light.glsl:161).  FLAG_SKIP_DEAD_SHADOW_RAYS drops the rays
    // whose result cannot matter; the radiance is the same (tests/test_oracle_kat.py proves it on the oracle itself).
    bool occluded = true;
    if (!(p.flags & ANEREND_DDGI_FLAG_SKIP_DEAD_SHADOW_RAYS) || dot(normal, shadowDir) > 0.0f) {
      cnt[2]++;
      occluded = sun.G ? anyHitSun(pos, shadowDir, p.t_min, p.t_max, cnt[5], cnt[6]) : anyHit(pos, shadowDir, p.t_min, p.t_max, cnt[5], cnt[6]);
    }
    V3 color = P(vv[0]->color) * b0 + P(vv[1]->color) * b1 + P(vv[2]->color) * b2;
    color = toLinear(color);
    // getSurfaceDataFromMat, bindless.glsl:181-237 (normal map never sampled: rchit:77 passes tangent = 0; discardPixel unused)
    const anerend_material_info_t& m = materials[in.material];
    V3 sColor = color;
    float tx[4];
    if (m.bindless_indices[1] != -1) {
      sampleTexture(m.bindless_indices[1], uvx, uvy, tx);
      sColor = v3(m.base_col_fac[0] * tx[0], m.base_col_fac[1] * tx[1], m.base_col_fac[2] * tx[2]);
    } else if (m.base_col_fac[0] + m.base_col_fac[1] + m.base_col_fac[2] > 0.1f) sColor = v3(m.base_col_fac[0], m.base_col_fac[1], m.base_col_fac[2]);
    float metallic = m.met_rough[0], roughness = m.met_rough[1];
    if (m.bindless_indices[0] != -1) { sampleTexture(m.bindless_indices[0], uvx, uvy, tx); metallic = tx[1]; roughness = tx[0]; }
    V3 emissive = v3(m.emissive[0], m.emissive[1], m.emissive[2]) * m.emissive[3];
    if (m.bindless_indices[3] != -1) {
      sampleTexture(m.bindless_indices[3], uvx, uvy, tx);
      emissive = v3(tx[0] * m.emissive[0] * m.emissive[3], tx[1] * m.emissive[1] * m.emissive[3], tx[2] * m.emissive[2] * m.emissive[3]);
    }

    V3 irradiance = v3(0.0f);
    litMask = 0;
    if (!occluded) {
      litMask |= 1u;
      irradiance = calcLight(normal, sColor, metallic, roughness, pos, rayOrigin, nullptr, 1, lightDir, ubo.sun_intensity);
    }
    if (p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) {
      // Extension E1 (not in the reference's rchit, which shades the sun only — quirk Q8): every enabled light,
      // composed from calcLight(directional = 0) and an any-hit shadow ray pos -> light.
      for (size_t li = 0; li < lights.size(); ++li) {
        const anerend_light_t& lt = lights[li];
        if (!(lt.color[3] > 0.5f)) continue;
        V3 lp = v3(lt.world_pos[0], lt.world_pos[1], lt.world_pos[2]);
        V3 dv = lp - pos; float dist = std::sqrt(dot(dv, dv));
        if (!(dist < lt.world_pos[3])) continue;                    // attenuation is exactly 0 at/after range
        V3 L = normalize(lp - pos);
        if (!(dot(normal, L) > 0.0f)) continue;                     // NdotL = 0 -> contribution exactly 0
        cnt[2]++;
        if (anyHit(pos, L, p.t_min, dist, cnt[5], cnt[6])) continue;
        litMask |= 1u << (1 + li);
        irradiance = irradiance + calcLight(normal, sColor, metallic, roughness, pos, rayOrigin, &lt, 0, lightDir, ubo.sun_intensity);
      }
    }
    const uint32_t gate = specular ? ANEREND_UBO_DDGI_FLAG : ANEREND_UBO_DDGI_MULTI_FLAG;
    if ((ubo.flags & gate) == gate)
      irradiance = irradiance + calcIndirectDiffuseLight(normal, sColor, metallic, roughness, pos, rayOrigin);
    irradiance = irradiance + emissive;
    return irradiance;
  }

  // calcWorldPos(pixelToUv(pixel, screen), depth) — shadow.rgen:20-28,38 and mirror_reflections.rgen:38-46,51:
  // mat4 * vec4 = c0*x + c1*y + c2*z + c3*w, then the perspective divide
  V3 screenWorldPos(uint32_t px, uint32_t py, float depthSamp) const {
    const float* m = ubo.inv_view_proj;
    const float sw = (float)ubo.screen_width, sh = (float)ubo.screen_height;
    const float u = (2.0f * (float)px + 1.0f) / (2.0f * sw), v = (2.0f * (float)py + 1.0f) / (2.0f * sh);   // pixelToUv, helpers.glsl:1-4
    const float cx = u * 2.0f - 1.0f, cy = v * 2.0f - 1.0f, cz = depthSamp;
    const float X = m[0] * cx + m[4] * cy + m[8] * cz + m[12] * 1.0f, Y = m[1] * cx + m[5] * cy + m[9] * cz + m[13] * 1.0f;
    const float Z = m[2] * cx + m[6] * cy + m[10] * cz + m[14] * 1.0f, W = m[3] * cx + m[7] * cy + m[11] * cz + m[15] * 1.0f;
    return v3(X / W, Y / W, Z / W);
  }
  // mirror_reflections.rgen:64-70: reflect(normalize(pos - cameraPos), normalize(geom0.rgb)); reflect(I, N) = I - 2 dot(N, I) N
  V3 reflectionDir(V3 pos, V3 geomNormal) const {
    const V3 normal = normalize(geomNormal);
    const V3 vv = normalize(pos - v3(ubo.camera_pos[0], ubo.camera_pos[1], ubo.camera_pos[2]));
    const float dn = dot(normal, vv);
    return vv - (2.0f * dn) * normal;
  }
  // rgen:51-83: probe (probeX, layer, probeZ) of the launch pixel -> ray origin
  V3 probeOrigin(int layer, int probe) const {
    int probeX = probe % nXZ, probeZ = probe / nXZ;
    return v3(float(probeX), float(layer), float(probeZ)) * gridStep() + camOffset();
  }
  // rgen:86-87: mat3(pc.randomRot) * worldDirFromFibonacciIndex(rayIndex), the Fibonacci direction from the host table
  V3 probeRayDir(const float* rot, int ray) const { return mulMat3(rot, v3(fibDirs[3 * ray], fibDirs[3 * ray + 1], fibDirs[3 * ray + 2])); }
  // rmiss:17-21
  V3 missRadiance() const { return toLinear(v3(ubo.sky_intensity * ubo.sky_color[0], ubo.sky_intensity * ubo.sky_color[1], ubo.sky_intensity * ubo.sky_color[2])); }

  // ---- rgen main for every ray of one layer (irradiance_probe_update.rgen:48-111) -------------------------
  int traceLayer(int layer, const float* rot) {
    if (!haveAccel) { err = "trace before build_accel"; return ANEREND_DDGI_E_STATE; }
    if (!haveUbo) { err = "trace before set_scene_ubo"; return ANEREND_DDGI_E_STATE; }
    if (layer < 0 || layer >= nY) { err = "layer out of range"; return ANEREND_DDGI_E_INVALID; }
    const int P = nXZ * nXZ;
    rayIrr[layer].assign((size_t)P * N * 4, 0); rayDir[layer].assign((size_t)P * N * 4, 0);
    bool rec = (p.flags & ANEREND_DDGI_FLAG_RAY_RECORDS) != 0;
    if (rec) records[layer].assign((size_t)P * N, anerend_ddgi_ray_record_t{});
    std::vector<V3> dirs(N);
    for (int i = 0; i < N; ++i) dirs[i] = probeRayDir(rot, i);
    V3 sky = missRadiance();
    parallelFor(P, [&](int probe) {
      V3 origin = probeOrigin(layer, probe);
      uint64_t cnt[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int ray = 0; ray < N; ++ray) {
        V3 d = dirs[ray]; Hit hit; uint32_t lit = 0; V3 rad;
        cnt[0]++;
        if (closestHit(origin, d, p.t_min, p.t_max, hit, cnt[3], cnt[4])) { cnt[1]++; rad = shadeHit(hit, origin, lit, cnt); }
        else rad = sky;
        size_t o = ((size_t)probe * N + ray) * 4;
        rayIrr[layer][o] = f2h(rad.x); rayIrr[layer][o + 1] = f2h(rad.y); rayIrr[layer][o + 2] = f2h(rad.z); rayIrr[layer][o + 3] = 0x3C00;
        rayDir[layer][o] = f2h(d.x); rayDir[layer][o + 1] = f2h(d.y); rayDir[layer][o + 2] = f2h(d.z); rayDir[layer][o + 3] = 0x3C00;
        if (rec) { anerend_ddgi_ray_record_t& r = records[layer][(size_t)probe * N + ray];
                   r.prim = hit.prim; r.t = hit.prim == 0xFFFFFFFFu ? 0.0f : hit.t; r.u = hit.u; r.v = hit.v;
                   r.radiance[0] = rad.x; r.radiance[1] = rad.y; r.radiance[2] = rad.z; r.lit_mask = lit; }
      }
      for (int k = 0; k < 7; ++k) st[k] += cnt[k];
    });
    return 0;
  }

  // ---- probe_conv.comp main (probe_conv.comp:21-90), live mode; quirk Q3 (extra dispatch row) waived ---------
  int convolveLayer(int layer) {
    if (layer < 0 || layer >= nY) { err = "layer out of range"; return ANEREND_DDGI_E_INVALID; }
    if (rayIrr[layer].empty()) { err = "convolve before trace of this layer"; return ANEREND_DDGI_E_STATE; }
    const int P = nXZ * nXZ;
    const float hysteresis = p.hysteresis;
    const bool trec = (p.flags & ANEREND_DDGI_FLAG_TEXEL_RECORDS) != 0;
    parallelFor(P, [&](int probe) {
      int probeX = probe % nXZ, probeZ = probe / nXZ;
      const uint16_t* irr = &rayIrr[layer][(size_t)probe * N * 4];
      const uint16_t* dir = &rayDir[layer][(size_t)probe * N * 4];
      for (int j = 0; j < 8; ++j) for (int i = 0; i < 8; ++i) {
        // pixelToUv(pixelWithinProbe, 8) * 2 - 1 ; helpers.glsl:1-4
        float ux = (2.0f * float(i) + 1.0f) / (2.0f * 8.0f), uy = (2.0f * float(j) + 1.0f) / (2.0f * 8.0f);
        ux = ux * 2.0f - 1.0f; uy = uy * 2.0f - 1.0f;
        V3 octDir = octDecode(ux, uy);
        float r4[4] = {0, 0, 0, 0};
        for (int x = 0; x < Sx; ++x) for (int y = 0; y < Sy; ++y) {           // x is the OUTER loop (probe_conv.comp:49-50)
          int ray = y * Sx + x;
          V3 rad = v3(h2f(irr[4 * ray]), h2f(irr[4 * ray + 1]), h2f(irr[4 * ray + 2]));
          V3 d = v3(h2f(dir[4 * ray]), h2f(dir[4 * ray + 1]), h2f(dir[4 * ray + 2]));
          float weight = std::fmax(0.0f, dot(d, octDir));
          if (weight > 1e-6f) { r4[0] += weight * rad.x; r4[1] += weight * rad.y; r4[2] += weight * rad.z; r4[3] += weight; }
        }
        int ax = 10 * probeX + 1 + i, ay = 10 * nXZ * layer + 10 * probeZ + 1 + j;   // pixOffset + pixel (probe_conv.comp:42,64)
        uint16_t* t = &atlas[4 * ((size_t)ay * W + ax)];
        float old[4] = {h2f(t[0]), h2f(t[1]), h2f(t[2]), h2f(t[3])};
        float baked[4];
        if (bakeMode()) {                      // probe_conv.comp:67-79: running weighted mean, weight sum in alpha
          float totalWeight = old[3] + r4[3];
          float tr[3] = {r4[0] + old[0] * old[3], r4[1] + old[1] * old[3], r4[2] + old[2] * old[3]};
          if (totalWeight > 1e-6f) { tr[0] /= totalWeight; tr[1] /= totalWeight; tr[2] /= totalWeight; }
          baked[0] = tr[0]; baked[1] = tr[1]; baked[2] = tr[2]; baked[3] = totalWeight;
        } else if (r4[3] > 1e-6f) { r4[0] /= r4[3]; r4[1] /= r4[3]; r4[2] /= r4[3]; }
        for (int c = 0; c < 4; ++c) {
          float o = bakeMode() ? baked[c] : r4[c] * (1.0f - hysteresis) + old[c] * hysteresis;
          t[c] = f2h(o);
          if (trec) texelRecords[4 * ((size_t)ay * W + ax) + c] = o;
        }
      }
    });
    return 0;
  }

  // ---- IrradianceProbeTranslationPass.cpp:56-186 --------------------------------------------------------------
  int translate() {
    if (!haveUbo) { err = "translate before set_scene_ubo"; return ANEREND_DDGI_E_STATE; }
    if ((ubo.flags & ANEREND_UBO_DDGI_FLAG) != ANEREND_UBO_DDGI_FLAG) return 0;
    if ((ubo.flags & ANEREND_UBO_BAKE_MODE_ON_FLAG) == ANEREND_UBO_BAKE_MODE_ON_FLAG) return 0;
    int ix = (int)(ubo.camera_pos[0] / p.probe_step[0]), iz = (int)(ubo.camera_pos[2] / p.probe_step[2]);
    int dx = ix - lastProbeIdx[0], dz = iz - lastProbeIdx[1];
    if (dx == 0 && dz == 0) return 0;
    lastProbeIdx[0] = ix; lastProbeIdx[1] = iz;
    temp = atlas;                                             // first vkCmdCopyImage: atlas -> TempProbeTex
    const int pps = 8 + 2, planeH = W;
    for (int i = 0; i < nY; ++i) {
      int planeY = i * pps * nXZ;
      int ew = W, eh = planeH, sx = 0, sy = planeY, dxo = 0, dyo = planeY;
      if (dx > 0) { ew = W - pps; sx = pps; dxo = 0; } else if (dx < 0) { ew = W - pps; sx = 0; dxo = pps; }
      if (dz > 0) { eh = planeH - pps; sy += pps; } else if (dz < 0) { eh = planeH - pps; dyo += pps; }
      for (int y = 0; y < eh; ++y)
        std::memcpy(&atlas[4 * ((size_t)(dyo + y) * W + dxo)], &temp[4 * ((size_t)(sy + y) * W + sx)], (size_t)ew * 8);
    }
    return 0;
  }

  bool ddgiOn() const {   // IrradianceProbeRayTracingPass.cpp:97-98 gating, via the UBO flags the renderer derives from RenderOptions
    return (ubo.flags & ANEREND_UBO_RT_ON_FLAG) && (ubo.flags & ANEREND_UBO_DDGI_FLAG);
  }
};

// ================================================================================================
// C API
// ================================================================================================
extern "C" {

static int rayTile(int n, int& sx, int& sy) {   // S = sqrt(N) when square (rgen:51-58); 2^ceil(log2 N / 2) wide otherwise
  sx = 1; while (sx * sx < n) sx *= 2; sy = n / sx; return (sx * sy == n) ? 0 : -1;
}

int ddgi_oracle_create(const anerend_ddgi_params_t* params, ddgi_oracle_t** out) {
  if (!params || !out || params->struct_size != sizeof(anerend_ddgi_params_t)) return ANEREND_DDGI_E_INVALID;
  const anerend_ddgi_params_t& q = *params;
  if (q.num_probes_xz < 1 || q.num_probes_y < 1 || q.rays_per_probe < 32 || q.rays_per_probe > 1024 || (q.rays_per_probe % 32)) return ANEREND_DDGI_E_INVALID;
  int sx, sy; if (rayTile(q.rays_per_probe, sx, sy)) return ANEREND_DDGI_E_INVALID;
  if (q.slab_begin < 0 || q.slab_end > q.num_probes_y || q.slab_begin >= q.slab_end) return ANEREND_DDGI_E_INVALID;
  ddgi_oracle* h = new ddgi_oracle();
  h->p = q; h->nXZ = q.num_probes_xz; h->nY = q.num_probes_y; h->N = q.rays_per_probe; h->Sx = sx; h->Sy = sy;
  h->W = 10 * h->nXZ; h->H = 10 * h->nXZ * h->nY;
  h->threads = (int)std::max(1u, std::thread::hardware_concurrency());
  h->fibDirs.resize((size_t)h->N * 3);
  for (int i = 0; i < h->N; ++i) { V3 d = fibonacciDir(i, h->N); h->fibDirs[3 * i] = d.x; h->fibDirs[3 * i + 1] = d.y; h->fibDirs[3 * i + 2] = d.z; }
  h->rayIrr.resize(h->nY); h->rayDir.resize(h->nY); h->records.resize(h->nY);
  for (auto& s : h->st) s = 0;
  h->clearAtlas();
  *out = h; return 0;
}
int ddgi_oracle_destroy(ddgi_oracle_t* h) { delete h; return 0; }
int ddgi_oracle_reset(ddgi_oracle_t* h) { h->clearAtlas(); h->currentLayer = 0; h->lastProbeIdx[0] = h->lastProbeIdx[1] = 0; return 0; }
const char* ddgi_oracle_last_error(const ddgi_oracle_t* h) { return h ? h->err.c_str() : "null handle"; }
int ddgi_oracle_set_threads(ddgi_oracle_t* h, int32_t n) { h->threads = n > 0 ? n : (int)std::max(1u, std::thread::hardware_concurrency()); return h->threads; }
int ddgi_oracle_set_brute_force(ddgi_oracle_t* h, int32_t on) { h->brute = on != 0; return 0; }

int ddgi_oracle_upload_scene(ddgi_oracle_t* h,
    const anerend_vertex_t* vertices, uint32_t nv, const uint32_t* indices, uint32_t ni,
    const anerend_mesh_info_t* meshes, const uint32_t* meshNumIndices, uint32_t nm,
    const anerend_material_info_t* materials, uint32_t nmat,
    const anerend_renderable_t* renderables, uint32_t nr,
    const uint32_t* modelIdx, uint32_t nmi, const uint32_t* matIdx, uint32_t nmti,
    const anerend_light_t* lights, uint32_t nl) {
  for (uint32_t i = 0; i < nmat; ++i) for (int k = 0; k < 4; ++k)
    if (k != 2 && materials[i].bindless_indices[k] != -1 && (materials[i].bindless_indices[k] < 0 || (size_t)materials[i].bindless_indices[k] >= h->textures.size())) {
      h->err = "material refers to a texture that was not uploaded (upload_textures first)"; return ANEREND_DDGI_E_INVALID; }
  if ((h->p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) && nl > 31) { h->err = "at most 31 point lights with FLAG_POINT_LIGHTS"; return ANEREND_DDGI_E_INVALID; }
  h->vertices.assign(vertices, vertices + nv); h->indices.assign(indices, indices + ni);
  h->meshes.assign(meshes, meshes + nm); h->meshNumIndices.assign(meshNumIndices, meshNumIndices + nm);
  h->materials.assign(materials, materials + nmat); h->renderables.assign(renderables, renderables + nr);
  h->modelIdx.assign(modelIdx, modelIdx + nmi); h->matIdx.assign(matIdx, matIdx + nmti);
  h->lights.assign(lights, lights + nl);
  h->haveScene = true; h->haveAccel = false; return 0;
}
int ddgi_oracle_upload_textures(ddgi_oracle_t* h, const anerend_texture_desc_t* tex, uint32_t count) {
  std::vector<ddgi_oracle::Tex> out(count);
  for (uint32_t i = 0; i < count; ++i) {
    if (!tex[i].pixels || !tex[i].width || !tex[i].height || tex[i].format > ANEREND_TEX_RGBA16F) { h->err = "bad texture descriptor"; return ANEREND_DDGI_E_INVALID; }
    size_t bpp = tex[i].format == ANEREND_TEX_RGBA16F ? 8 : tex[i].format <= ANEREND_TEX_RGBA8_SRGB ? 4 : (tex[i].format == ANEREND_TEX_RG8_UNORM ? 2 : 1);
    out[i].w = tex[i].width; out[i].h = tex[i].height; out[i].fmt = tex[i].format; out[i].clampEdge = tex[i].clamp_to_edge != 0;
    const uint8_t* p = (const uint8_t*)tex[i].pixels;
    out[i].px.assign(p, p + (size_t)tex[i].width * tex[i].height * bpp);
  }
  h->textures.swap(out); h->makeTexLut();
  return 0;
}
int ddgi_oracle_sample_texture(ddgi_oracle_t* h, int32_t idx, float u, float v, float out[4]) {
  if (idx < 0 || (size_t)idx >= h->textures.size()) return ANEREND_DDGI_E_INVALID;
  h->sampleTexture(idx, u, v, out); return 0;
}
int ddgi_oracle_set_scene_ubo(ddgi_oracle_t* h, const anerend_scene_ubo_t* ubo) {
  const bool sunMoved = !h->haveUbo || std::memcmp(h->ubo.light_dir, ubo->light_dir, 12) != 0;
  h->ubo = *ubo; h->haveUbo = true;
  if (sunMoved && h->haveAccel) h->buildSunGrid();          // the grid lives in the sun's space
  return 0;
}
// Prototype switch: resolution of the light-space grid the sun's shadow rays use instead of the hierarchy (0 = off, the default).
int ddgi_oracle_set_sun_grid(ddgi_oracle_t* h, int32_t resolution) {
  if (resolution < 0 || resolution > 4096) return ANEREND_DDGI_E_INVALID;
  h->sunGridRes = resolution;
  if (h->haveAccel) h->buildSunGrid();
  return 0;
}
int ddgi_oracle_build_accel(ddgi_oracle_t* h) {
  if (!h->haveScene) { h->err = "build_accel before upload_scene"; return ANEREND_DDGI_E_STATE; }
  int rc = h->buildInstances(); if (rc) return rc;
  h->bvh.build(h->tris); h->buildSunGrid(); h->haveAccel = true; return 0;
}
int ddgi_oracle_update_vertices(ddgi_oracle_t* h, const anerend_vertex_t* v, uint32_t first, uint32_t count) {
  if (!v || (size_t)first + count > h->vertices.size()) { h->err = "vertex range out of bounds"; return ANEREND_DDGI_E_INVALID; }
  std::copy(v, v + count, h->vertices.begin() + first);
  return 0;
}
int ddgi_oracle_refit_accel(ddgi_oracle_t* h, const anerend_renderable_t* r, uint32_t n) {
  if (!h->haveAccel) { h->err = "refit before build_accel"; return ANEREND_DDGI_E_STATE; }
  if (r) { if (n != h->renderables.size()) { h->err = "renderable count changed"; return ANEREND_DDGI_E_INVALID; } h->renderables.assign(r, r + n); }
  size_t nt = h->tris.size(), ninst = h->instances.size();
  int rc = h->buildInstances(); if (rc) return rc;
  if (h->tris.size() != nt || h->instances.size() != ninst) { h->err = "visibility changed: rebuild required"; h->haveAccel = false; return ANEREND_DDGI_E_STATE; }
  h->bvh.build(h->tris);   // the oracle simply rebuilds; closest hit does not depend on the tree
  h->buildSunGrid();
  return 0;
}
// cull.comp:241-244,296-303 (withinRangeOfCamera + setCullBit), centre = buildTranslation(model) * vec4(bounds.xyz, 1) (cull.comp:229-239,297)
int ddgi_oracle_apply_camera_cull(ddgi_oracle_t* h, float range, uint32_t* changedOut) {
  if (!h->haveUbo) { h->err = "apply_camera_cull before set_scene_ubo"; return ANEREND_DDGI_E_STATE; }
  uint32_t changed = 0;
  for (anerend_renderable_t& r : h->renderables) {
    if (!(r.visible & 1u)) continue;
    const float trans[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, r.transform[12], r.transform[13], r.transform[14], 1};
    const V3 centre = xformPoint(trans, v3(r.bounds[0], r.bounds[1], r.bounds[2]));
    const V3 d = v3(h->ubo.camera_pos[0], h->ubo.camera_pos[1], h->ubo.camera_pos[2]) - centre;
    const bool within = std::sqrt(dot(d, d)) <= range;
    const uint32_t v = (r.visible & ~2u) | (within ? 0u : 2u);
    if (v != r.visible) { r.visible = v; ++changed; }
  }
  if (changed) h->haveAccel = false;
  if (changedOut) *changedOut = changed;
  return 0;
}
int ddgi_oracle_get_bvh_info(ddgi_oracle_t* h, anerend_ddgi_bvh_info_t* o) {
  std::memset(o, 0, sizeof(*o));
  o->num_triangles = (uint32_t)h->tris.size(); o->num_nodes = (uint32_t)h->bvh.nodes.size(); o->node_bytes = sizeof(BNode); o->tri_bytes = sizeof(WTri);
  o->num_instances = (uint32_t)h->instances.size();
  if (!h->bvh.nodes.empty()) for (int k = 0; k < 3; ++k) { o->bounds_min[k] = h->bvh.nodes[0].lo[k]; o->bounds_max[k] = h->bvh.nodes[0].hi[k]; }
  return 0;
}
int ddgi_oracle_download_instances(ddgi_oracle_t* h, anerend_ddgi_instance_t* out, size_t count) {
  if (!out || count != h->instances.size()) { h->err = "instance count mismatch"; return ANEREND_DDGI_E_INVALID; }
  for (size_t i = 0; i < count; ++i) { const Instance& in = h->instances[i];
    out[i].renderable = in.renderable; out[i].custom_index = in.material | (in.mesh << 10); out[i].first_triangle = in.firstTri; out[i].num_triangles = in.numTris; }
  return 0;
}
int ddgi_oracle_download_world_triangles(ddgi_oracle_t* h, float* out, size_t count) {
  if (count != h->tris.size()) { h->err = "triangle count mismatch"; return ANEREND_DDGI_E_INVALID; }
  for (size_t i = 0; i < count; ++i) { const WTri& t = h->tris[i]; V3 a = t.v0, b = t.v0 + t.e1, c = t.v0 + t.e2;
    float* o = out + 9 * i; o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = c.x; o[7] = c.y; o[8] = c.z; }
  return 0;
}

int ddgi_oracle_translate(ddgi_oracle_t* h) { return h->translate(); }
int ddgi_oracle_trace_layer(ddgi_oracle_t* h, int32_t layer, const float rot[16]) {
  if (h->haveUbo && !h->ddgiOn()) return 0;
  return h->traceLayer(layer, rot);
}
int ddgi_oracle_convolve_layer(ddgi_oracle_t* h, int32_t layer) {
  if (h->haveUbo && !h->ddgiOn()) return 0;
  return h->convolveLayer(layer);
}
int ddgi_oracle_update(ddgi_oracle_t* h, int32_t layer, const float* rots) {
  if (h->haveUbo && !h->ddgiOn()) return 0;
  if (layer >= 0) { int rc = h->traceLayer(layer, rots); return rc ? rc : h->convolveLayer(layer); }
  for (int l = h->p.slab_begin; l < h->p.slab_end; ++l) { int rc = h->traceLayer(l, rots + 16 * (size_t)l); if (rc) return rc; }
  for (int l = h->p.slab_begin; l < h->p.slab_end; ++l) { int rc = h->convolveLayer(l); if (rc) return rc; }
  return 0;
}
int ddgi_oracle_frame(ddgi_oracle_t* h, const float rot[16], int32_t* layerOut) {
  int rc = h->translate(); if (rc) return rc;
  if (!h->ddgiOn()) return 0;
  int layer = h->currentLayer;
  rc = h->traceLayer(layer, rot); if (rc) return rc;
  if (layerOut) *layerOut = layer;
  h->currentLayer = (h->currentLayer + 1 >= h->nY) ? 0 : h->currentLayer + 1;
  return h->convolveLayer(layer);
}

int ddgi_oracle_atlas_dims(const ddgi_oracle_t* h, int32_t* w, int32_t* hh) { *w = h->W; *hh = h->H; return 0; }
int ddgi_oracle_download_atlas(ddgi_oracle_t* h, void* out, size_t bytes) {
  if (bytes != h->atlas.size() * 2) { h->err = "atlas size mismatch"; return ANEREND_DDGI_E_INVALID; }
  std::memcpy(out, h->atlas.data(), bytes); return 0;
}
int ddgi_oracle_upload_atlas(ddgi_oracle_t* h, const void* in, size_t bytes) {
  if (bytes != h->atlas.size() * 2) { h->err = "atlas size mismatch"; return ANEREND_DDGI_E_INVALID; }
  std::memcpy(h->atlas.data(), in, bytes); return 0;
}
int ddgi_oracle_ray_tile(const ddgi_oracle_t* h, int32_t* sx, int32_t* sy) { *sx = h->Sx; *sy = h->Sy; return 0; }

// linear [probe][ray] <-> the reference's image layout: pixel (probeX*Sx+px, probeZ*Sy+py), ray = py*Sx+px
static void toImage(const ddgi_oracle* h, const std::vector<uint16_t>& lin, uint16_t* img) {
  int IW = h->Sx * h->nXZ;
  for (int probe = 0; probe < h->nXZ * h->nXZ; ++probe) for (int ray = 0; ray < h->N; ++ray) {
    int px = ray % h->Sx, py = ray / h->Sx, X = (probe % h->nXZ) * h->Sx + px, Y = (probe / h->nXZ) * h->Sy + py;
    std::memcpy(&img[4 * ((size_t)Y * IW + X)], &lin[4 * ((size_t)probe * h->N + ray)], 8);
  }
}
int ddgi_oracle_download_rays(ddgi_oracle_t* h, int32_t layer, void* irr, void* dir, size_t bytesEach) {
  if (layer < 0 || layer >= h->nY || h->rayIrr[layer].empty()) { h->err = "no rays for this layer"; return ANEREND_DDGI_E_STATE; }
  if (bytesEach != h->rayIrr[layer].size() * 2) { h->err = "ray image size mismatch"; return ANEREND_DDGI_E_INVALID; }
  if (irr) toImage(h, h->rayIrr[layer], (uint16_t*)irr);
  if (dir) toImage(h, h->rayDir[layer], (uint16_t*)dir);
  return 0;
}
int ddgi_oracle_set_rays(ddgi_oracle_t* h, int32_t layer, const void* irr, const void* dir, size_t bytesEach) {
  size_t n = (size_t)h->nXZ * h->nXZ * h->N * 4;
  if (layer < 0 || layer >= h->nY || bytesEach != n * 2) { h->err = "bad set_rays arguments"; return ANEREND_DDGI_E_INVALID; }
  h->rayIrr[layer].resize(n); h->rayDir[layer].resize(n);
  int IW = h->Sx * h->nXZ; const uint16_t* ii = (const uint16_t*)irr; const uint16_t* di = (const uint16_t*)dir;
  for (int probe = 0; probe < h->nXZ * h->nXZ; ++probe) for (int ray = 0; ray < h->N; ++ray) {
    int px = ray % h->Sx, py = ray / h->Sx, X = (probe % h->nXZ) * h->Sx + px, Y = (probe / h->nXZ) * h->Sy + py;
    std::memcpy(&h->rayIrr[layer][4 * ((size_t)probe * h->N + ray)], &ii[4 * ((size_t)Y * IW + X)], 8);
    std::memcpy(&h->rayDir[layer][4 * ((size_t)probe * h->N + ray)], &di[4 * ((size_t)Y * IW + X)], 8);
  }
  return 0;
}
// the product's _upload_ray_radiance: radiance image from the caller, direction image as trace_layer writes it for `rot`
int ddgi_oracle_upload_ray_radiance(ddgi_oracle_t* h, int32_t layer, const void* irr, size_t bytes, const float rot[16]) {
  size_t n = (size_t)h->nXZ * h->nXZ * h->N * 4;
  if (layer < 0 || layer >= h->nY || bytes != n * 2 || !irr || !rot) { h->err = "bad upload_ray_radiance arguments"; return ANEREND_DDGI_E_INVALID; }
  h->rayIrr[layer].resize(n); h->rayDir[layer].resize(n);
  int IW = h->Sx * h->nXZ; const uint16_t* ii = (const uint16_t*)irr;
  for (int probe = 0; probe < h->nXZ * h->nXZ; ++probe) for (int ray = 0; ray < h->N; ++ray) {
    int px = ray % h->Sx, py = ray / h->Sx, X = (probe % h->nXZ) * h->Sx + px, Y = (probe / h->nXZ) * h->Sy + py;
    size_t o = 4 * ((size_t)probe * h->N + ray);
    std::memcpy(&h->rayIrr[layer][o], &ii[4 * ((size_t)Y * IW + X)], 8);
    V3 d = h->probeRayDir(rot, ray);
    h->rayDir[layer][o] = f2h(d.x); h->rayDir[layer][o + 1] = f2h(d.y); h->rayDir[layer][o + 2] = f2h(d.z); h->rayDir[layer][o + 3] = 0x3C00;
  }
  return 0;
}
int ddgi_oracle_download_ray_records(ddgi_oracle_t* h, int32_t layer, anerend_ddgi_ray_record_t* out, size_t count) {
  if (!(h->p.flags & ANEREND_DDGI_FLAG_RAY_RECORDS)) { h->err = "FLAG_RAY_RECORDS not set"; return ANEREND_DDGI_E_STATE; }
  if (layer < 0 || layer >= h->nY || h->records[layer].size() != count) { h->err = "no records for this layer / count mismatch"; return ANEREND_DDGI_E_INVALID; }
  std::memcpy(out, h->records[layer].data(), count * sizeof(*out)); return 0;
}
int ddgi_oracle_download_texel_records(ddgi_oracle_t* h, float* out, size_t bytes) {
  if (!(h->p.flags & ANEREND_DDGI_FLAG_TEXEL_RECORDS)) { h->err = "FLAG_TEXEL_RECORDS not set"; return ANEREND_DDGI_E_STATE; }
  if (bytes != h->texelRecords.size() * 4) { h->err = "size mismatch"; return ANEREND_DDGI_E_INVALID; }
  std::memcpy(out, h->texelRecords.data(), bytes); return 0;
}
int ddgi_oracle_get_stats(ddgi_oracle_t* h, anerend_ddgi_stats_t* o) {
  o->primary_rays = h->st[0]; o->primary_hits = h->st[1]; o->shadow_rays = h->st[2]; o->primary_nodes_visited = h->st[3];
  o->primary_tris_tested = h->st[4]; o->shadow_nodes_visited = h->st[5]; o->shadow_tris_tested = h->st[6]; o->shadow_rays_prefiltered = 0; o->shadow_rays_candidate = 0;
  o->primary_level_changes = 0; o->shadow_level_changes = 0; return 0;
}
int ddgi_oracle_reset_stats(ddgi_oracle_t* h) { for (auto& s : h->st) s = 0; return 0; }

int ddgi_oracle_make_rotation(int32_t axis, float deg, float out[16]) { if (axis < 0 || axis > 2) return ANEREND_DDGI_E_INVALID; makeRotation(axis, deg, out); return 0; }
// any-hit query through the oracle's own hierarchy for an arbitrary ray (planning probes: tests/lightspace_probe.py)
int ddgi_oracle_any_hit(ddgi_oracle_t* h, const float origin[3], const float dir[3], float tmin, float tmax) {
  if (!h || !h->haveAccel) return ANEREND_DDGI_E_INVALID;
  uint64_t a = 0, b = 0;
  return h->anyHit(v3(origin[0], origin[1], origin[2]), v3(dir[0], dir[1], dir[2]), tmin, tmax, a, b) ? 1 : 0;
}
int ddgi_oracle_shade_hit(ddgi_oracle_t* h, uint32_t prim, float u, float v, const float origin[3], int32_t specular, float out[4], uint32_t* lit_mask) {
  if (!h || !h->haveAccel || !h->haveUbo || prim >= h->tris.size()) return ANEREND_DDGI_E_INVALID;
  Hit hit{prim, 0.0f, u, v}; uint32_t lit = 0; uint64_t cnt[7] = {0, 0, 0, 0, 0, 0, 0}; V3 hp;
  const V3 o = v3(origin[0], origin[1], origin[2]);
  V3 rad = h->shadeHit(hit, o, lit, cnt, specular != 0, &hp);
  const V3 dd = hp - o;
  out[0] = rad.x; out[1] = rad.y; out[2] = rad.z; out[3] = std::sqrt(dot(dd, dd));       // distance(pos, gl_WorldRayOriginEXT), mirror_reflections.rchit:90
  if (lit_mask) *lit_mask = lit;
  return 0;
}
int ddgi_oracle_screen_ray(ddgi_oracle_t* h, int32_t specular, uint32_t px, uint32_t py, float depth, const float geom_normal[3], float origin[3], float dir[3]) {
  if (!h || !h->haveUbo) return ANEREND_DDGI_E_INVALID;
  V3 o = h->screenWorldPos(px, py, depth);
  V3 d = specular ? h->reflectionDir(o, v3(geom_normal[0], geom_normal[1], geom_normal[2])) : -v3(h->ubo.light_dir[0], h->ubo.light_dir[1], h->ubo.light_dir[2]);
  origin[0] = o.x; origin[1] = o.y; origin[2] = o.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
  return 0;
}
int ddgi_oracle_probe_ray(ddgi_oracle_t* h, int32_t layer, int32_t probe, int32_t ray, const float rot[16], float origin[3], float dir[3], float miss[3]) {
  if (!h || !h->haveUbo || layer < 0 || layer >= h->nY || probe < 0 || probe >= h->nXZ * h->nXZ || ray < 0 || ray >= h->N) return ANEREND_DDGI_E_INVALID;
  V3 o = h->probeOrigin(layer, probe), d = h->probeRayDir(rot, ray), m = h->missRadiance();
  origin[0] = o.x; origin[1] = o.y; origin[2] = o.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z; miss[0] = m.x; miss[1] = m.y; miss[2] = m.z;
  return 0;
}
int ddgi_oracle_fibonacci_dirs(int32_t n, float* out) {
  if (n < 1) return ANEREND_DDGI_E_INVALID;
  for (int i = 0; i < n; ++i) { V3 d = fibonacciDir(i, n); out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z; } return 0;
}
void ddgi_oracle_oct_encode(const float v[3], float out[2]) { octEncode(v3(v[0], v[1], v[2]), out); }
void ddgi_oracle_oct_decode(const float o[2], float out[3]) { V3 d = octDecode(o[0], o[1]); out[0] = d.x; out[1] = d.y; out[2] = d.z; }
void ddgi_oracle_to_linear(const float s[3], float out[3]) { V3 c = toLinear(v3(s[0], s[1], s[2])); out[0] = c.x; out[1] = c.y; out[2] = c.z; }
uint16_t ddgi_oracle_float_to_half(float f) { return f2h(f); }
float ddgi_oracle_half_to_float(uint16_t hh) { return h2f(hh); }
void ddgi_oracle_calc_light(const float n[3], const float a[3], float metallic, float roughness, const float wp[3], const float vp[3],
                            const anerend_light_t* light, int32_t directional, const float ld[3], float sunIntensity, float out[3]) {
  V3 r = ddgi_oracle::calcLight(v3(n[0], n[1], n[2]), v3(a[0], a[1], a[2]), metallic, roughness, v3(wp[0], wp[1], wp[2]), v3(vp[0], vp[1], vp[2]),
                                light, directional, v3(ld[0], ld[1], ld[2]), sunIntensity);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int ddgi_oracle_calc_indirect(ddgi_oracle_t* h, const float n[3], const float a[3], float metallic, float roughness,
                              const float wp[3], const float vp[3], float out[3]) {
  V3 r = h->calcIndirectDiffuseLight(v3(n[0], n[1], n[2]), v3(a[0], a[1], a[2]), metallic, roughness, v3(wp[0], wp[1], wp[2]), v3(vp[0], vp[1], vp[2]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; return 0;
}
// mirror_reflections.rgen:30-81 + .rchit + .rmiss (frame-graph pass "SpecularGIRT", SpecularGIRTPass.cpp:100-127): per pixel, reflect
// the view vector about the G-buffer normal, closest hit, shade like the probe rchit; out = rgba16f (irradiance, hit distance).
int ddgi_oracle_specular_gi(ddgi_oracle_t* h, const float* depth, const float* geom0, uint32_t width, uint32_t height, uint16_t* out) {
  if (!h->haveAccel) { h->err = "specular_gi before build_accel"; return ANEREND_DDGI_E_STATE; }
  if (!h->haveUbo) { h->err = "specular_gi before set_scene_ubo"; return ANEREND_DDGI_E_STATE; }
  if (!(h->ubo.flags & ANEREND_UBO_RT_ON_FLAG) || !(h->ubo.flags & ANEREND_UBO_SPECULAR_GI_FLAG)) return 0;   // SpecularGIRTPass.cpp:102-103
  const V3 sky = toLinear(v3(h->ubo.sky_intensity * h->ubo.sky_color[0], h->ubo.sky_intensity * h->ubo.sky_color[1], h->ubo.sky_intensity * h->ubo.sky_color[2]));
  h->parallelFor((int)height, [&](int py) {
    uint64_t cnt[7] = {0, 0, 0, 0, 0, 0, 0};
    for (uint32_t px = 0; px < width; ++px) {
      const size_t i = (size_t)py * width + px;
      uint16_t* o = out + 4 * i;
      const float depthSamp = depth[i];
      if (depthSamp == 1.0f) { o[0] = o[1] = o[2] = o[3] = 0; continue; }                        // rgen:41-44
      const V3 pos = h->screenWorldPos(px, py, depthSamp);
      const V3 dir = h->reflectionDir(pos, v3(geom0[4 * i], geom0[4 * i + 1], geom0[4 * i + 2]));
      Hit hit; uint32_t lit = 0; V3 rad; float dist;
      if (h->closestHit(pos, dir, 0.1f, 50.0f, hit, cnt[3], cnt[4])) {
        V3 hp;
        rad = h->shadeHit(hit, pos, lit, cnt, true, &hp);
        const V3 dd = hp - pos; dist = std::sqrt(dot(dd, dd));                                   // distance(pos, gl_WorldRayOriginEXT)
      } else { rad = sky; dist = 1e6f; }                                                         // mirror_reflections.rmiss
      o[0] = f2h(rad.x); o[1] = f2h(rad.y); o[2] = f2h(rad.z); o[3] = f2h(dist);
    }
  });
  return 0;
}
// shadow.rgen:30-82 + shadow.rmiss (frame-graph pass "ShadowRT", ShadowRayTracingPass.cpp:73-81) for a width x height depth buffer
int ddgi_oracle_shadow_rt(ddgi_oracle_t* h, const float* depth, uint32_t width, uint32_t height, float* out) {
  if (!h->haveAccel) { h->err = "shadow_rt before build_accel"; return ANEREND_DDGI_E_STATE; }
  if (!h->haveUbo) { h->err = "shadow_rt before set_scene_ubo"; return ANEREND_DDGI_E_STATE; }
  if (!(h->ubo.flags & ANEREND_UBO_RT_SHADOWS_FLAG)) return 0;                       // shadow.rgen:32-34
  const V3 dir = -v3(h->ubo.light_dir[0], h->ubo.light_dir[1], h->ubo.light_dir[2]);  // shadow.rgen:58, not normalised
  h->parallelFor((int)height, [&](int py) {
    uint64_t cn = 0, ct = 0;
    for (uint32_t px = 0; px < width; ++px) {
      const size_t i = (size_t)py * width + px;
      const float depthSamp = depth[i];                                               // texel centre (pixelToUv, helpers.glsl:1-4)
      if (depthSamp == 1.0f) { out[i] = 1.0f; continue; }                             // sky, shadow.rgen:45-48
      const V3 origin = h->screenWorldPos(px, py, depthSamp);                        // calcWorldPos, shadow.rgen:20-28
      out[i] = h->anyHit(origin, dir, 0.1f, 50.0f, cn, ct) ? 0.0f : 1.0f;           // payload 0, shadow.rmiss sets 1
    }
  });
  return 0;
}
// deferred_tiled.comp:264-296 / mirror_reflections.rchit:92-118: calcIndirectDiffuseLight per surface point (12 floats each)
int ddgi_oracle_indirect_diffuse(ddgi_oracle_t* h, const float* pts, uint32_t count, const float vp[3], int32_t clampToEdge, float* out) {
  h->clampSampler = clampToEdge != 0;
  for (uint32_t i = 0; i < count; ++i) {
    const float* q = pts + 12 * (size_t)i;
    // deferred_tiled.comp:264-296 passes the tile's baked atlas as probeTex where there is one; mirror_reflections.rchit the live atlas
    { int tx, tz; h->srcTex = clampToEdge ? h->worldPosToTileAtlas(v3(q[0], q[1], q[2]), tx, tz) : -1; }
    V3 r = h->calcIndirectDiffuseLight(v3(q[3], q[4], q[5]), v3(q[6], q[7], q[8]), q[9], q[10], v3(q[0], q[1], q[2]), v3(vp[0], vp[1], vp[2]));
    out[4 * (size_t)i] = r.x; out[4 * (size_t)i + 1] = r.y; out[4 * (size_t)i + 2] = r.z; out[4 * (size_t)i + 3] = 0.0f;
  }
  h->clampSampler = false; h->srcTex = -1;
  return 0;
}
int ddgi_oracle_upload_tile_infos(ddgi_oracle_t* h, const anerend_tile_info_t* tiles, uint32_t count) {
  h->tileInfos.resize(count);
  for (uint32_t i = 0; i < count; ++i) h->tileInfos[i] = tiles[i].ddgi_atlas_tex;
  return 0;
}
int ddgi_oracle_intersect(const float o[3], const float d[3], const float a[3], const float b[3], const float c[3],
                          float tmin, float tmax, float* t, float* u, float* v) {
  V3 v0 = v3(a[0], a[1], a[2]), v1 = v3(b[0], b[1], b[2]), v2 = v3(c[0], c[1], c[2]);
  float tt = 0, uu = 0, vv = 0;
  bool hit = tri_intersect(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), v0, v1 - v0, v2 - v0, tmin, tmax, tt, uu, vv);
  if (t) *t = tt; if (u) *u = uu; if (v) *v = vv; return hit ? 1 : 0;
}

}  // extern "C"
