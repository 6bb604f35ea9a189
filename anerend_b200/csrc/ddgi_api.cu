// ddgi_api.cu — the C ABI (include/anerend_cuda_ddgi.h) and the host side of the three passes.
//
// Host-side restatement of the reference's pass code:
//   IrradianceProbeTranslationPass.cpp:56-186  -> translate()
//   IrradianceProbeRayTracingPass.cpp:95-165   -> traceLayers()   (+ rgen/rchit/rmiss in ddgi_kernels.cuh)
//   IrradianceProbeConvolvePass.cpp:109-143    -> convolveLayers() (+ probe_conv.comp)
//   tlas_update.comp:37-99 instance selection  -> selectInstances()
// Compiled by nvcc into libanerend_cuda_ddgi.so.  (With -DDDGI_HOST_EMU the same file builds the serial debug
// harness under tests/host_emu/, exported as anerend_emu_ddgi_*; it is never part of the product library.)
#include "ddgi_launch.h"
#ifndef DDGI_HOST_EMU
#include "ddgi_fast.cuh"
#endif

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

using namespace ddgi;

// ------------------------------------------------------------------------------------------------ kernels
DDGI_KERNEL(k_inst_xform, inst_xform_body, BuildCtx)
DDGI_KERNEL(k_tri_instance, tri_instance_body, BuildCtx)
DDGI_KERNEL(k_flatten_scene, flatten_tri_body, BuildCtx)
DDGI_KERNEL(k_morton_codes, morton_body, BuildCtx)
DDGI_KERNEL(k_lbvh_hierarchy, karras_body, BuildCtx)
DDGI_KERNEL(k_lbvh_fit, fit_body, BuildCtx)
DDGI_KERNEL(k_ploc_init, ploc_init_body, BuildCtx)
DDGI_KERNEL(k_ploc_nn, ploc_nn_body, BuildCtx)
DDGI_KERNEL(k_ploc_merge, ploc_merge_body, BuildCtx)
DDGI_KERNEL(k_ploc_compact, ploc_compact_body, BuildCtx)
DDGI_KERNEL(k_ploc_order, ploc_order_body, BuildCtx)
DDGI_KERNEL(k_ploc_relabel, ploc_relabel_body, BuildCtx)
DDGI_KERNEL(k_bvh8_collapse, collapse_body, BuildCtx)
DDGI_KERNEL(k_refit_tris, refit_tris_body, BuildCtx)
DDGI_KERNEL(k_refit_nodes, refit_node_body, BuildCtx)
DDGI_KERNEL(k_blas_prims, blas_prim_body, BlasPrimCtx)
DDGI_KERNEL(k_blas_reloc, blas_reloc_body, BlasRelocCtx)
DDGI_KERNEL(k_tlas_prims, tlas_prim_body, TlasPrimCtx)
DDGI_KERNEL(k_probe_trace, trace_ray_body<false>, TraceCtx)
DDGI_KERNEL(k_probe_trace_counted, trace_ray_body<true>, TraceCtx)
DDGI_KERNEL(k_probe_convolve, conv_texel_body, ConvCtx)
DDGI_KERNEL(k_probe_translate, translate_body, TransCtx)
DDGI_KERNEL(k_indirect_diffuse, indirect_diffuse_body, ConsumerCtx)
DDGI_KERNEL(k_shadow_rt, shadow_rt_body, ShadowRtCtx)
DDGI_KERNEL(k_specular_gi, specular_gi_body, SpecularCtx)
DDGI_KERNEL(k_atlas_clear, clear_body, ClearCtx)
DDGI_KERNEL(k_lightmap_raster_serial, lightmap_raster_body, LmBuildCtx)
DDGI_KERNEL(k_lightmap_moved, lm_moved_body, LmMovedCtx)
DDGI_KERNEL(k_vis_changed, vis_changed_body, VisCtx)
DDGI_KERNEL(k_sunmap_raster_serial, sunmap_raster_body, SunMapCtx)

// ------------------------------------------------------------------------------------------------ handle
struct anerend_ddgi {
  anerend_ddgi_params_t p{};
  Backend be;
  std::string err;
  int nXZ = 0, nY = 0, N = 0, Sx = 0, Sy = 0, W = 0, H = 0, P = 0;

  // host copies needed for instance selection
  std::vector<anerend_renderable_t> renderables;
  std::vector<anerend_mesh_info_t> meshes;
  std::vector<uint32_t> meshNumIndices, modelIdx, matIdx;
  std::vector<DInstance> instances;
  uint32_t numVertices = 0, numIndices = 0, numMaterials = 0, numLights = 0;
  anerend_scene_ubo_t ubo{};
  bool haveScene = false, haveUbo = false, haveAccel = false;

  // device scene
  anerend_vertex_t* dVertices = nullptr; uint32_t* dIndices = nullptr; anerend_material_info_t* dMaterials = nullptr;
  anerend_renderable_t* dRenderables = nullptr; anerend_light_t* dLights = nullptr;
  DInstance* dInstances = nullptr; InstXform* dXforms = nullptr; uint32_t* dTriInstance = nullptr;
  // acceleration structure
  uint32_t numTris = 0, numNodes = 0, bvhDepth = 0;
  uint32_t capTris = 0, capInstances = 0; bool capTextured = false;      // what the buffers of the structure were allocated for (reused by rebuilds)
  float4* dWtri = nullptr; float4* dTriLo = nullptr; float4* dTriHi = nullptr; int* dSceneBounds = nullptr;
  uint4* dNodes = nullptr; float4* dBvhTris = nullptr; uint32_t* dNodeParent = nullptr; float4* dNodeLo = nullptr; float4* dNodeHi = nullptr;
  uint32_t* dNodePending = nullptr; uint32_t* dVisit = nullptr; uint32_t* dCounters = nullptr;
  uint32_t* dInstChanged = nullptr;      // refit: instances whose transform changed since the last one
  float boundsMin[3] = {0, 0, 0}, boundsMax[3] = {0, 0, 0};
  // two-level structure (ANEREND_DDGI_FLAG_TWO_LEVEL): one hierarchy per distinct mesh behind the instance level's slots of the
  // node / leaf arrays; numTris stays the number of flattened triangle ids, numNodes the nodes of the instance level
  struct MeshBlas { uint32_t mesh, numTris, triBase, root, nodes, depth; };
  std::vector<MeshBlas> blas; std::vector<uint32_t> tlasList;
  uint2* dInstBlas = nullptr; uint32_t* dTlasList = nullptr;
  uint32_t tlasCap = 0, blasTris = 0, blasNodes = 0, tlasDepth = 0, blasDepthMax = 0, primCap = 0; bool twoBuilt = false; float tlasMs = 0.0f;

  // pass resources
  uint2* dAtlas = nullptr; uint2* dTemp = nullptr;        // ProbeRaysConvTex, TempProbeTex
  uint2* dRayIrr = nullptr;                                // ProbeRaysIrrTex for all layers, linear [layer][probe][ray]
  float* dFib = nullptr;                                   // [N][3] unrotated Fibonacci directions
  float4* dDirs = nullptr; float4* hDirs = nullptr;        // [nY][N] rotated directions (ProbeRaysDirTex is derived from it)
  int* dLayerMap = nullptr; int* hLayerMap = nullptr;      // [nY] absolute layers of the current launch (pinned staging + device)
  int numLaunchLayers = 0;
  uint16_t* dRayOrder = nullptr;
  anerend_ddgi_ray_record_t* dRecords = nullptr; float4* dTexelRecords = nullptr; unsigned long long* dStats = nullptr;
  std::vector<float> fib;                                  // N*3
  std::vector<uint8_t> layerTraced;
  uint2* peers[8] = {nullptr}; int numPeers = 0;
  // fused atlas exchange over NVLink peer memory (exchange_export/connect): arrival flags of the device-side barrier
  uint32_t* dXflags = nullptr; uint32_t* dXerr = nullptr;   // own flags [8] (written by the peers), sticky time-out flag
  uint32_t* xflags[8] = {nullptr};                          // flag arrays of all ranks, own included, by rank
  void* xopened[16] = {nullptr}; int numOpened = 0;         // cudaIpc mappings to close
  uint32_t xrank = 0, xworld = 0, xepoch = 0;
#ifndef DDGI_HOST_EMU
  // the push of an update runs on its own stream under the NEXT update's tracing: only the hit shading samples the atlas
  cudaStream_t xStream = nullptr; cudaEvent_t evConv = nullptr, evXDone = nullptr; bool exchangePending = false;
  // Pipelined updates (anerend_cuda_ddgi_set_pipeline): the closest-hit trace of an update reads nothing the previous update writes
  // (ray generation and traversal need the rotation and the acceleration structure; only the hit shading samples the atlas), so it
  // runs on a stream of its own as soon as the call is made — under the tail of the previous update's kernels — with its own copies
  // of the direction table / layer map, its own fetch cursor and a second hit-record buffer; the rest of the update waits for it.
  bool pipeline = false, pipeStateDirty = true, pipeDoneValid[2] = {false, false};
  cudaStream_t pStream = nullptr; cudaEvent_t evTrace[2] = {nullptr, nullptr}, evHitsFree[2] = {nullptr, nullptr}, evState = nullptr;
  float4* dHits2 = nullptr; uint32_t* dHitInst2 = nullptr; float4* dDirsP[2] = {nullptr, nullptr}; int* dLayerMapP[2] = {nullptr, nullptr}; uint32_t* dCursorP = nullptr;
  uint32_t pipeIdx = 0;
#endif
  // wavefront (fast) path scratch: hit records, surface records, visibility masks, per-light shadow queues
  float4* dHits = nullptr; float4* dSurf = nullptr; uint32_t* dLit = nullptr; uint32_t* dQueues = nullptr; uint32_t* dQcount = nullptr; uint32_t* dHitInst = nullptr;
  uint32_t* dHitq = nullptr; uint32_t* dHitqSorted = nullptr; float4* dShadeTris = nullptr; float* dConvW = nullptr;
  // bindless textures: descriptor table, texel pool, decode table (sRGB EOTF + UNORM); per-triangle uv records and per-ray
  // emissive exist only while some material is textured
  int32_t* dTileInfos = nullptr; uint32_t numTiles = 0;     // tileInfoBuffer (baked atlases of the paged tiles)
  DTexture* dTextures = nullptr; uint8_t* dTexPool = nullptr; float* dSrgbLut = nullptr; uint32_t numTextures = 0;
  std::vector<DTexture> textures; bool texturedScene = false, emissiveTextures = false;
  bool verticesDirty = false;      // update_vertices since the last build / refit
  std::vector<std::pair<uint32_t, uint32_t>> dirtyVerts;      // ... and which vertex ranges (two-level: only the meshes they touch are rebuilt)
  std::vector<uint32_t> meshNumVertices;                      // highest index of each mesh + 1
  bool texturesStale = false;      // the texture array was replaced after the scene that refers to it was validated
  float4* dShadeUV = nullptr; float4* dEmisTex = nullptr;
  // ShadowRT scratch (pixels): staged positions, visibility words, ray queue, counters
  float4* dExtRays = nullptr; float* dExtDist = nullptr; uint32_t extPixels = 0;      // SpecularGIRT: staged rays, hit distances
  float4* dSrtPos = nullptr; uint32_t* dSrtLit = nullptr; uint32_t* dSrtQueue = nullptr; uint32_t* dSrtQcount = nullptr; uint32_t srtPixels = 0; int srtGrid = 0;
  void* dBigItems = nullptr; uint32_t* dBigCount = nullptr;           // queue of large light-map footprints (k_lightmap_raster_big)
  uint32_t* dMovedTris = nullptr; uint32_t* dMovedCount = nullptr; uint32_t movedCap = 0;      // refit: triangles of the instances that moved (k_moved_tris)
  uint32_t* dSelected = nullptr; uint32_t* dVisChanged = nullptr;      // per renderable: in the structure (tlas_update.comp:41-47); device flag of refit_accel_device
  // light maps (ddgi_kernels.cuh): static + dynamic distance cube maps per point light, the transforms the static map was made with
  unsigned long long* dLmStatic = nullptr; unsigned long long* dLmDynamic = nullptr; uint32_t* dLmNear = nullptr; InstXform* dLmBase = nullptr; uint32_t* dLmMoved = nullptr;
  uint32_t lmRes = 0, lmLights = 0; bool lmDynamicValid = false; float lmMs = 0.0f;
  uint32_t lmDynShift = LM_DYN_SHIFT;      // the movers' maps stay at 96..128 texels a side (re-made at every refit): shift 3 from R = 768
  unsigned long long* dSmStatic = nullptr; unsigned long long* dSmDynamic = nullptr; uint32_t smRes = 0; bool smValid = false, smDynamicValid = false, lmMovedValid = false;
  float smDir[3] = {0, 0, 0}; SunMapCtx sm{};      // the lightDir the sun's map was made for, its frame
  size_t fastRays = 0; uint32_t fastQueues = 0; int gridPrimary[2] = {0, 0}, gridShadow[2] = {0, 0}; uint32_t tune[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int sms = 0;

  int rowOffset = 0, rowStride = 1, Pl = 0;                // row stripe (set_row_stripe): probes launched per layer = Pl
  int currentLayer = 0;                                    // IrradianceProbeRayTracingPass::_currentLayer
  int lastProbeIdx[2] = {0, 0};                            // IrradianceProbeTranslationPass::_lastProbeIdx
  anerend_ddgi_timings_t timings{};
  bool timed[4] = {false, false, false, false}; bool timedKernels = false;
  float accelMs = 0.0f; int plocRounds = 0;
};

namespace {

int fail(anerend_ddgi* h, int code, const std::string& msg) { h->err = msg; return code; }
int fenceExchange(anerend_ddgi* h);      // multi-GPU: the previous update's rows from the peers have landed (defined with the exchange)
#define BE_CHECK(h, expr) do { if (!(expr)) return fail((h), ANEREND_DDGI_E_CUDA, std::string(#expr ": ") + (h)->be.errStr()); } while (0)

int rayTile(int n, int& sx, int& sy) { sx = 1; while (sx * sx < n) sx *= 2; sy = n / sx; return (sx * sy == n) ? 0 : -1; }

// rgen:35-46 worldDirFromFibonacciIndex, evaluated once on the host in fp32
void fibonacciDir(uint32_t idx, uint32_t n, float out[3]) {
  const float PI = 3.14159265359f, GOLDEN_RATIO = 1.6180339887f;
  float theta = 2.0f * PI * float(idx) / GOLDEN_RATIO;
  float phi = acosf(1.0f - 2.0f * (float(idx) + 0.5f) / float(n));
  float x = cosf(theta) * sinf(phi), y = sinf(theta) * sinf(phi), z = cosf(phi);
  float inv = 1.0f / sqrtf(x * x + y * y + z * z);
  out[0] = x * inv; out[1] = y * inv; out[2] = z * inv;
}

bool ddgiOn(const anerend_ddgi* h) {   // IrradianceProbeRayTracingPass.cpp:97-98 via the UBO flags (VulkanRenderer.cpp:1222-1232)
  return (h->ubo.flags & ANEREND_UBO_RT_ON_FLAG) && (h->ubo.flags & ANEREND_UBO_DDGI_FLAG);
}

// Origin and spacing of the probe grid: live mode follows the camera (rgen:75-83, pbr_light.glsl:295-301; int() truncates toward
// zero, quirk Q1); bake mode (UBO_BAKE_MODE_ON_FLAG, rgen:61-73, pbr_light.glsl:274-291) covers the tile bakeTileInfo.xy of
// size bakeTileInfo.z with step (tSize / nXZ, 1.0, tSize / nXZ) — the y step is 1, not PROBE_STEP.y.
bool bakeMode(const anerend_ddgi* h) { return (h->ubo.flags & ANEREND_UBO_BAKE_MODE_ON_FLAG) == ANEREND_UBO_BAKE_MODE_ON_FLAG; }
void gridFrame(const anerend_ddgi* h, V3& camOff, V3& stepOut) {
  if (bakeMode(h)) {
    const float ts = float(h->ubo.bake_tile_info[2]), s = ts / float(h->nXZ);
    stepOut = mk3(s, 1.0f, s);
    camOff = mk3(float(h->ubo.bake_tile_info[0]) * ts, 0.0f, float(h->ubo.bake_tile_info[1]) * ts);
    return;
  }
  const float* step = h->p.probe_step;
  int cix = (int)(h->ubo.camera_pos[0] / step[0]), ciz = (int)(h->ubo.camera_pos[2] / step[2]);
  camOff = mk3(float(cix) * step[0] - step[0] * float(h->nXZ) / 2.0f, 0.0f, float(ciz) * step[2] - step[2] * float(h->nXZ) / 2.0f);
  stepOut = mk3(step[0], step[1], step[2]);
}

// tlas_update.comp:37-99: which (renderable, mesh) pairs enter the acceleration structure, in renderable order
int selectInstances(anerend_ddgi* h, std::vector<DInstance>& out, uint32_t& numTris) {
  out.clear(); numTris = 0;
  for (uint32_t idx = 0; idx < h->renderables.size(); ++idx) {
    const anerend_renderable_t& r = h->renderables[idx];
    bool visible = (r.visible & 1u) != 0, culled = ((r.visible >> 1) & 1u) != 0;
    if (!visible || culled) continue;
    for (uint32_t i = 0; i < r.num_meshes; ++i) {
      if (r.model_offset + i >= h->modelIdx.size()) return fail(h, ANEREND_DDGI_E_INVALID, "model index out of range");
      uint32_t meshIndex = h->modelIdx[r.model_offset + i];
      if (r.first_material_index + i >= h->matIdx.size()) return fail(h, ANEREND_DDGI_E_INVALID, "material index out of range");
      uint32_t matIndex = h->matIdx[r.first_material_index + i];
      if (r.dynamic_model_offset > 0) {
        if (r.dynamic_model_offset + i >= h->modelIdx.size()) return fail(h, ANEREND_DDGI_E_INVALID, "dynamic model index out of range");
        meshIndex = h->modelIdx[r.dynamic_model_offset + i];
      }
      uint32_t custom = (matIndex & ~(0x3FFFu << 10)) | ((meshIndex & 0x3FFFu) << 10);   // 10 bits material | 14 bits mesh
      DInstance in{};
      in.renderable = idx; in.material = custom & 0x3FFu; in.mesh = (custom >> 10) & 0x3FFFu;
      if (in.mesh >= h->meshes.size() || in.material >= h->numMaterials) return fail(h, ANEREND_DDGI_E_INVALID, "mesh/material id out of range");
      in.firstTri = numTris; in.numTris = h->meshNumIndices[in.mesh] / 3;
      in.vertexOffset = h->meshes[in.mesh].vertex_offset; in.indexOffset = h->meshes[in.mesh].index_offset;
      if ((uint64_t)in.indexOffset + 3ull * in.numTris > h->numIndices) return fail(h, ANEREND_DDGI_E_INVALID, "mesh indices out of range");
      numTris += in.numTris; out.push_back(in);
    }
  }
  return 0;
}

BuildCtx buildCtx(anerend_ddgi* h) {
  BuildCtx c{};
  c.vertices = h->dVertices; c.indices = h->dIndices; c.renderables = h->dRenderables;
  c.instances = h->dInstances; c.xforms = h->dXforms; c.triInstance = h->dTriInstance;
  c.numTris = h->numTris; c.numInstances = (uint32_t)h->instances.size();
  c.wtri = h->dWtri; c.triLo = h->dTriLo; c.triHi = h->dTriHi; c.sceneBounds = h->dSceneBounds;
  c.nodes = h->dNodes; c.bvhTris = h->dBvhTris; c.nodeParent = h->dNodeParent; c.nodeLo = h->dNodeLo; c.nodeHi = h->dNodeHi;
  c.nodePending = h->dNodePending; c.visit = h->dVisit; c.counters = h->dCounters;
  c.shadeTris = nullptr; c.shadeUV = nullptr;
  return c;
}

void freeAccel(anerend_ddgi* h) {
  Backend& be = h->be;
  be.free(h->dInstances); be.free(h->dXforms); be.free(h->dTriInstance); be.free(h->dWtri); be.free(h->dTriLo); be.free(h->dTriHi);
  be.free(h->dNodes); be.free(h->dBvhTris); be.free(h->dNodeParent); be.free(h->dNodeLo); be.free(h->dNodeHi); be.free(h->dNodePending); be.free(h->dVisit);
  be.free(h->dShadeTris); h->dShadeTris = nullptr; be.free(h->dShadeUV); h->dShadeUV = nullptr;
  be.free(h->dInstChanged); h->dInstChanged = nullptr;
  be.free(h->dLmStatic); be.free(h->dLmDynamic); be.free(h->dLmNear); be.free(h->dLmBase); be.free(h->dLmMoved);
  h->dLmStatic = nullptr; h->dLmDynamic = nullptr; h->dLmNear = nullptr; h->dLmBase = nullptr; h->dLmMoved = nullptr; h->lmRes = 0; h->lmLights = 0; h->lmDynamicValid = false;
  be.free(h->dSmStatic); be.free(h->dSmDynamic); h->dSmStatic = nullptr; h->dSmDynamic = nullptr; h->smRes = 0; h->smValid = false; h->smDynamicValid = false; h->lmMovedValid = false;
  h->dInstances = nullptr; h->dXforms = nullptr; h->dTriInstance = nullptr; h->dWtri = nullptr; h->dTriLo = nullptr; h->dTriHi = nullptr;
  h->dNodes = nullptr; h->dBvhTris = nullptr; h->dNodeParent = nullptr; h->dNodeLo = nullptr; h->dNodeHi = nullptr; h->dNodePending = nullptr; h->dVisit = nullptr;
  h->haveAccel = false; h->numTris = 0; h->numNodes = 0; h->capTris = 0; h->capInstances = 0;
  be.free(h->dInstBlas); be.free(h->dTlasList); h->dInstBlas = nullptr; h->dTlasList = nullptr;
  h->blas.clear(); h->tlasList.clear(); h->twoBuilt = false; h->tlasCap = 0; h->blasTris = 0; h->blasNodes = 0; h->primCap = 0;
}
void freeScene(anerend_ddgi* h) {
  Backend& be = h->be;
  be.free(h->dVertices); be.free(h->dIndices); be.free(h->dMaterials); be.free(h->dRenderables); be.free(h->dLights);
  h->dVertices = nullptr; h->dIndices = nullptr; h->dMaterials = nullptr; h->dRenderables = nullptr; h->dLights = nullptr;
  h->haveScene = false;
}

int clearAtlas(anerend_ddgi* h) {
  BE_CHECK(h, h->be.fenceDownload());
  { int rf = fenceExchange(h); if (rf) return rf; }
  ClearCtx c{h->dAtlas, h->dTemp};
  DDGI_LAUNCH(h->be, k_atlas_clear, clear_body, (uint32_t)(h->W * h->H), c, 256);
  if (h->dTexelRecords) BE_CHECK(h, h->be.zero(h->dTexelRecords, (size_t)h->W * h->H * sizeof(float4)));
  BE_CHECK(h, h->be.check());
  return 0;
}

#ifndef DDGI_HOST_EMU
// One thread per (triangle, light of blockIdx.y): the setup of lightmap_raster_body per thread, the texel loops of the
// footprints spread over the warp (a footprint is a few dozen texels at R = 512, but a triangle next to a light covers a
// whole face: per-thread loops would leave one lane running for milliseconds).  The lane whose footprint is being filled
// hands its parameters to the warp through shared memory.
// Footprints of more than LM_BIG_TEXELS texels (a triangle next to a light covers a whole face, a wall made of two triangles
// thousands of texels of the sun's map; a mover among the lights hundreds) are not filled by the warp that found them — it would run for milliseconds while the
// rest of the GPU idles — but appended to a queue that a second launch fills with one block per footprint.
#define LM_BIG_TEXELS 256
#define LM_BIG_CAP 65536u
struct LmItem { int x0, x1, y0, y1; V3 n; float h, daabb; uint32_t skinny, tid, face, light; LmEdges ed; };
struct SmItem { int x0, x1, y0, y1; float smax, ga, gb, pa0, pb0, ps0; uint32_t flat, tid; LmEdges ed; };
struct BigQueue { void* items; uint32_t* count; };
__device__ __forceinline__ void lm_fill(const LmItem& w, const LmBuildCtx& c, uint32_t first, uint32_t step) {
  unsigned long long* dm = c.depth + ((size_t)w.light * 6u + (size_t)w.face) * c.R * c.R;
  const int wdt = w.x1 - w.x0 + 1, cnt = wdt * (w.y1 - w.y0 + 1);
  for (int k = (int)first; k < cnt; k += (int)step) {
    const int lx = k % wdt, ly = k / wdt;
    if (lm_edges_touch(w.ed, lx, ly))
      atomMin64(dm + (size_t)(w.y0 + ly) * c.R + (w.x0 + lx),
                ((unsigned long long)__float_as_uint(lm_texel_bound(w.n, w.h, w.daabb, w.skinny != 0u, (int)w.face, w.x0 + lx, w.y0 + ly, c.R, c.rho)) << 32) | w.tid);
  }
}
__device__ __forceinline__ void lm_raster_thread(const LmBuildCtx& c, const BigQueue& big, uint32_t t, bool valid, uint32_t li, LmItem* sItem) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  LmTri tri;
  tri.n = mk3(0.0f); tri.h = 0.0f; tri.daabb = 0.0f; tri.skinny = false; tri.nearLight = false;
  const bool ok = valid && lm_pair_setup(c, t, li, tri);
  if (ok && tri.nearLight) c.nearFlags[li] = 1u;
  for (int face = 0; face < 6; ++face) {
    LmItem it; it.x0 = 0; it.x1 = -1; it.y0 = 0; it.y1 = -1; it.ed.use = false;
    bool reach = ok && lm_face_rect(tri, face, c.R, it.x0, it.x1, it.y0, it.y1, it.ed);
    if (reach) { it.n = tri.n; it.h = tri.h; it.daabb = tri.daabb; it.skinny = tri.skinny ? 1u : 0u; it.tid = t; it.face = (uint32_t)face; it.light = li; }
    if (reach && (it.x1 - it.x0 + 1) * (it.y1 - it.y0 + 1) > LM_BIG_TEXELS) {
      const uint32_t slot = atomicAdd(big.count, 1u);
      if (slot < LM_BIG_CAP) { reinterpret_cast<LmItem*>(big.items)[slot] = it; reach = false; }      // queue full: filled here after all
    }
    for (uint32_t m = __ballot_sync(0xFFFFFFFFu, reach); m != 0u; m &= m - 1u) {
      const uint32_t src = (uint32_t)__ffs((int)m) - 1u;
      __syncwarp();
      if (lane == src) sItem[warp] = it;
      __syncwarp();
      lm_fill(sItem[warp], c, lane, 32u);
    }
  }
}
// one thread per (triangle, light of blockIdx.y); the dynamic pass runs over all triangles too and the threads of instances that have
// not moved leave at once (a launch over the movers only — one block per instance — was slower: the fill is latency bound and a
// few blocks keep too few warps in flight)
__global__ void __launch_bounds__(256) k_lightmap_raster(const LmBuildCtx c, const BigQueue big) {
  __shared__ LmItem sItem[8];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  lm_raster_thread(c, big, t, t < c.numTris, blockIdx.y, sItem);
}
__global__ void __launch_bounds__(256) k_lightmap_raster_big(const LmBuildCtx c, const BigQueue big) {
  __shared__ LmItem sIt;
  const uint32_t n = min(*big.count, LM_BIG_CAP);
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) sIt = reinterpret_cast<const LmItem*>(big.items)[i];
    __syncthreads();
    lm_fill(sIt, c, threadIdx.x, blockDim.x);
  }
}
__device__ __forceinline__ void sm_fill(const SmItem& w, const SunMapCtx& c, uint32_t first, uint32_t step) {
  const int wdt = w.x1 - w.x0 + 1, cnt = wdt * (w.y1 - w.y0 + 1);
  for (int k = (int)first; k < cnt; k += (int)step) {
    const int lx = k % wdt, ly = k / wdt;
    if (lm_edges_touch(w.ed, lx, ly))
      atomMax64(c.height + (size_t)(w.y0 + ly) * c.R + (w.x0 + lx),
                sm_pack(sm_texel_bound(w.smax, w.ga, w.gb, w.pa0, w.pb0, w.ps0, w.flat != 0u, w.x0 + lx, w.y0 + ly, c.orgA, c.orgB, c.cell), c.base, w.tid));
  }
}
__device__ __forceinline__ void sm_raster_thread(const SunMapCtx& c, const BigQueue& big, uint32_t t, bool valid, SmItem* sItem) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  SmTri tri; SmItem it; it.x0 = 0; it.x1 = -1; it.y0 = 0; it.y1 = -1; it.ed.use = false;
  tri.smax = 0.0f; tri.ga = 0.0f; tri.gb = 0.0f; tri.flat = false; tri.pa[0] = 0.0f; tri.pb[0] = 0.0f; tri.ps[0] = 0.0f;
  bool reach = valid && sm_tri_setup(c, t, tri, it.x0, it.x1, it.y0, it.y1, it.ed);
  if (reach) { it.smax = tri.smax; it.ga = tri.ga; it.gb = tri.gb; it.pa0 = tri.pa[0]; it.pb0 = tri.pb[0]; it.ps0 = tri.ps[0]; it.flat = tri.flat ? 1u : 0u; it.tid = t; }
  if (reach && (it.x1 - it.x0 + 1) * (it.y1 - it.y0 + 1) > LM_BIG_TEXELS) {
    const uint32_t slot = atomicAdd(big.count, 1u);
    if (slot < LM_BIG_CAP) { reinterpret_cast<SmItem*>(big.items)[slot] = it; reach = false; }
  }
  for (uint32_t m = __ballot_sync(0xFFFFFFFFu, reach); m != 0u; m &= m - 1u) {
    const uint32_t src = (uint32_t)__ffs((int)m) - 1u;
    __syncwarp();
    if (lane == src) sItem[warp] = it;
    __syncwarp();
    sm_fill(sItem[warp], c, lane, 32u);
  }
}
__global__ void __launch_bounds__(256) k_sunmap_raster(const SunMapCtx c, const BigQueue big) {
  __shared__ SmItem sItem[8];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  sm_raster_thread(c, big, t, t < c.numTris, sItem);
}
__global__ void __launch_bounds__(256) k_sunmap_raster_big(const SunMapCtx c, const BigQueue big) {
  __shared__ SmItem sIt;
  const uint32_t n = min(*big.count, LM_BIG_CAP);
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) sIt = reinterpret_cast<const SmItem*>(big.items)[i];
    __syncthreads();
    sm_fill(sIt, c, threadIdx.x, blockDim.x);
  }
}
// Refit: the dynamic maps hold the triangles of the instances that moved — a few thousand of a few hundred thousand.  One thread per
// (triangle, light) left all the work to the few warps that hold a mover's triangles, 32 footprints x 6 faces in a row each (C4:
// 117 us for the lights, 25 us for the sun).  Here the movers' triangles are compacted first and every (triangle, light) pair gets a
// thread (a warp per pair was tried: every lane repeats the set-up, 186 M warp instructions, slower): at the dynamic maps' resolution a
// footprint is a handful of texels, which the thread fills itself (up to LM_OWN_TEXELS: a thread that fills 64 texels on several faces
// IS the duration of the launch); larger ones go to the queue that k_lightmap_raster_big fills, here one warp per footprint.
__global__ void __launch_bounds__(256) k_moved_tris(const uint32_t* triInstance, const uint32_t* moved, uint32_t numTris, uint32_t* list, uint32_t* count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
  const bool m = t < numTris && moved[triInstance[t]] != 0u;
  const uint32_t b = __ballot_sync(0xFFFFFFFFu, m);
  if (b == 0u) return;
  const int leader = __ffs((int)b) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(b));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (m) list[base + (uint32_t)__popc(b & ((1u << lane) - 1u))] = t;
}
#define LM_OWN_TEXELS 12
__global__ void __launch_bounds__(128) k_lightmap_raster_list(const LmBuildCtx c, const BigQueue big, const uint32_t* list, const uint32_t* count) {
  const uint32_t n = *count * c.numLights;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const uint32_t t = list[w / c.numLights], li = w % c.numLights;
    LmTri tri;
    tri.n = mk3(0.0f); tri.h = 0.0f; tri.daabb = 0.0f; tri.skinny = false; tri.nearLight = false;
    if (!lm_pair_setup(c, t, li, tri)) continue;
    if (tri.nearLight) c.nearFlags[li] = 1u;
    for (int face = 0; face < 6; ++face) {
      LmItem it; it.x0 = 0; it.x1 = -1; it.y0 = 0; it.y1 = -1; it.ed.use = false;
      if (!lm_face_rect(tri, face, c.R, it.x0, it.x1, it.y0, it.y1, it.ed)) continue;
      it.n = tri.n; it.h = tri.h; it.daabb = tri.daabb; it.skinny = tri.skinny ? 1u : 0u; it.tid = t; it.face = (uint32_t)face; it.light = li;
      if ((it.x1 - it.x0 + 1) * (it.y1 - it.y0 + 1) > LM_OWN_TEXELS) {
        const uint32_t slot = atomicAdd(big.count, 1u);
        if (slot < LM_BIG_CAP) { reinterpret_cast<LmItem*>(big.items)[slot] = it; continue; }      // queue full: filled here after all
      }
      lm_fill(it, c, 0u, 1u);
    }
  }
}
__global__ void __launch_bounds__(128) k_sunmap_raster_list(const SunMapCtx c, const BigQueue big, const uint32_t* list, const uint32_t* count) {
  const uint32_t n = *count;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const uint32_t t = list[w];
    SmTri tri; SmItem it; it.x0 = 0; it.x1 = -1; it.y0 = 0; it.y1 = -1; it.ed.use = false;
    tri.smax = 0.0f; tri.ga = 0.0f; tri.gb = 0.0f; tri.flat = false; tri.pa[0] = 0.0f; tri.pb[0] = 0.0f; tri.ps[0] = 0.0f;
    if (!sm_tri_setup(c, t, tri, it.x0, it.x1, it.y0, it.y1, it.ed)) continue;
    it.smax = tri.smax; it.ga = tri.ga; it.gb = tri.gb; it.pa0 = tri.pa[0]; it.pb0 = tri.pb[0]; it.ps0 = tri.ps[0]; it.flat = tri.flat ? 1u : 0u; it.tid = t;
    if ((it.x1 - it.x0 + 1) * (it.y1 - it.y0 + 1) > LM_OWN_TEXELS) {
      const uint32_t slot = atomicAdd(big.count, 1u);
      if (slot < LM_BIG_CAP) { reinterpret_cast<SmItem*>(big.items)[slot] = it; continue; }
    }
    sm_fill(it, c, 0u, 1u);
  }
}
#endif
#ifndef DDGI_HOST_EMU
__global__ void __launch_bounds__(256) k_zero_words(uint32_t* p, uint32_t n);
int ensureBigQueue(anerend_ddgi* h) {
  if (h->dBigItems) return 0;
  const size_t itemBytes = sizeof(LmItem) > sizeof(SmItem) ? sizeof(LmItem) : sizeof(SmItem);
  BE_CHECK(h, h->be.alloc((uint8_t**)&h->dBigItems, itemBytes * LM_BIG_CAP)); BE_CHECK(h, h->be.alloc(&h->dBigCount, 1));
  if (h->sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, dev); }
  return 0;
}
#endif
bool twoLevel(const anerend_ddgi* h) { return (h->p.flags & ANEREND_DDGI_FLAG_TWO_LEVEL) != 0; }
// the maps are rasterised from the world-space triangle copies, which a two-level structure does not keep
bool mapsEnabled(const anerend_ddgi* h) { return !(h->p.flags & ANEREND_DDGI_FLAG_NO_LIGHT_MAPS) && h->numTris != 0 && !twoLevel(h); }
int rasterSunMap(anerend_ddgi* h, unsigned long long* target, const uint32_t* moved) {
  Backend& be = h->be;
  const uint32_t R = moved ? h->smRes >> LM_DYN_SHIFT : h->smRes;          // the dynamic map is coarser: same frame, cells 2^shift times as wide
  BE_CHECK(h, be.zero(target, (size_t)R * R * sizeof(unsigned long long)));
  SunMapCtx c = h->sm;
  if (moved) { c.cell = h->sm.cell * (float)(1u << LM_DYN_SHIFT); c.invCell = h->sm.invCell / (float)(1u << LM_DYN_SHIFT); }
  c.wtri = h->dWtri; c.triInstance = h->dTriInstance; c.moved = moved; c.numTris = h->numTris; c.R = R; c.height = target;
#ifndef DDGI_HOST_EMU
  { int rq = ensureBigQueue(h); if (rq) return rq; }
  const BigQueue big{h->dBigItems, h->dBigCount};
  if (moved && h->dMovedTris) {          // refit: a thread per triangle of a moved instance (compacted by refitLightMaps)
    k_zero_words<<<1, 32, 0, be.stream>>>(h->dBigCount, 1u);
    k_sunmap_raster_list<<<8 * (h->sms > 0 ? h->sms : 148), 128, 0, be.stream>>>(c, big, h->dMovedTris, h->dMovedCount);
    k_sunmap_raster_big<<<32 * (h->sms > 0 ? h->sms : 148), 32, 0, be.stream>>>(c, big);
    be.launches += 3;
  } else {
  k_zero_words<<<1, 32, 0, be.stream>>>(h->dBigCount, 1u);
  k_sunmap_raster<<<(h->numTris + 255u) / 256u, 256, 0, be.stream>>>(c, big);
  k_sunmap_raster_big<<<4 * (h->sms > 0 ? h->sms : 148), 256, 0, be.stream>>>(c, big);
  be.launches += 3;
  }
#else
  DDGI_LAUNCH(be, k_sunmap_raster_serial, sunmap_raster_body, h->numTris, c, 256);
#endif
  BE_CHECK(h, be.check());
  return 0;
}
uint32_t lightMapResolution(const anerend_ddgi* h) {
  if (!(h->p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) || !mapsEnabled(h) || h->numLights == 0) return 0;
  // 768: C3 0.874 -> 0.846 ms per update against 512 (1.83 instead of 1.74 M shadow rays answered without a traversal), a rebuild
  // 2.5 -> 3.4 ms, 225 instead of 100 MB of maps at 8 lights; 1024 gives another 1 % for 400 MB
  uint32_t r = 768;
  if (const char* e = getenv("DDGI_LM_RES")) r = (uint32_t)atoi(e);          // sweeps only; 0 disables the maps
  r = r > 4096u ? 4096u : r;
  return r & ~7u;      // a multiple of the dynamic maps' coarsening (lmDynShift <= 3)
}
// Rasterises the triangles into the light maps: all of them into the static map (after build_accel, or when vertices
// changed), or only those of moved instances into the dynamic map (refit).
int rasterLightMap(anerend_ddgi* h, unsigned long long* target, const uint32_t* moved) {
  Backend& be = h->be;
  const uint32_t R = moved ? h->lmRes >> h->lmDynShift : h->lmRes;          // the dynamic map is coarser
  const size_t texels = (size_t)h->lmLights * 6u * R * R;
  BE_CHECK(h, be.fill(target, 0x7F, texels * sizeof(unsigned long long)));
  LmBuildCtx c{};
  c.wtri = h->dWtri; c.triInstance = h->dTriInstance; c.moved = moved; c.lights = h->dLights; c.numTris = h->numTris; c.numLights = h->lmLights;
  c.R = R; c.rho = lm_rho(R); c.depth = target; c.nearFlags = h->dLmNear;
#ifndef DDGI_HOST_EMU
  { int rq = ensureBigQueue(h); if (rq) return rq; }
  const BigQueue big{h->dBigItems, h->dBigCount};
  if (moved && h->dMovedTris) {
    k_zero_words<<<1, 32, 0, be.stream>>>(h->dBigCount, 1u);
    k_lightmap_raster_list<<<8 * (h->sms > 0 ? h->sms : 148), 128, 0, be.stream>>>(c, big, h->dMovedTris, h->dMovedCount);
    k_lightmap_raster_big<<<32 * (h->sms > 0 ? h->sms : 148), 32, 0, be.stream>>>(c, big);          // footprints of a few dozen texels: a warp each
    be.launches += 3;
  } else {
  k_zero_words<<<1, 32, 0, be.stream>>>(h->dBigCount, 1u);
  k_lightmap_raster<<<dim3((h->numTris + 255u) / 256u, h->lmLights), 256, 0, be.stream>>>(c, big);
  k_lightmap_raster_big<<<4 * (h->sms > 0 ? h->sms : 148), 256, 0, be.stream>>>(c, big);
  be.launches += 3;
  }
#else
  DDGI_LAUNCH(be, k_lightmap_raster_serial, lightmap_raster_body, h->numTris * h->lmLights, c, 256);
#endif
  BE_CHECK(h, be.check());
  return 0;
}
int buildLightMaps(anerend_ddgi* h) {
  Backend& be = h->be;
  h->lmDynamicValid = false; h->smDynamicValid = false; h->lmMovedValid = false;
  h->smValid = false;                                     // the sun's map is made at the next trace, for the lightDir of that frame (ensureSunMap)
  if (!mapsEnabled(h)) { h->lmRes = 0; return 0; }
  const size_t ni = h->instances.size();
  if (!h->dLmBase) { BE_CHECK(h, be.alloc(&h->dLmBase, (size_t)h->capInstances)); BE_CHECK(h, be.alloc(&h->dLmMoved, (size_t)h->capInstances)); }
  BE_CHECK(h, be.d2d(h->dLmBase, h->dXforms, ni * sizeof(InstXform)));      // the pose both static maps are made with
  const uint32_t R = lightMapResolution(h);
  if (R == 0) { h->lmRes = 0; return 0; }
  const size_t texels = (size_t)h->numLights * 6u * R * R;
  if (h->lmRes != R || h->lmLights != h->numLights || !h->dLmStatic) {
    be.free(h->dLmStatic); be.free(h->dLmDynamic); be.free(h->dLmNear);
    h->dLmStatic = nullptr; h->dLmDynamic = nullptr; h->dLmNear = nullptr;
    BE_CHECK(h, be.alloc(&h->dLmStatic, texels)); BE_CHECK(h, be.alloc(&h->dLmNear, h->numLights));
    h->lmRes = R; h->lmLights = h->numLights; h->lmDynShift = R >= 768u ? 3u : LM_DYN_SHIFT;
  }
  BE_CHECK(h, be.zero(h->dLmNear, h->numLights * sizeof(uint32_t)));
  return rasterLightMap(h, h->dLmStatic, nullptr);
}
// after a refit: instances whose transform left the pose of the static maps go into the dynamic maps at their new place
int refitLightMaps(anerend_ddgi* h, bool verticesChanged) {
  Backend& be = h->be;
  if (!mapsEnabled(h)) return 0;
  if (verticesChanged) return buildLightMaps(h);          // deformed meshes: every triangle may have moved
  const uint32_t ni = (uint32_t)h->instances.size();
  LmMovedCtx mc{h->dXforms, h->dLmBase, h->dLmMoved};
  DDGI_LAUNCH(be, k_lightmap_moved, lm_moved_body, ni, mc, 128);
  h->lmMovedValid = true;
#ifndef DDGI_HOST_EMU
  if (h->lmRes || h->smValid) {
    if (h->movedCap < h->capTris) {
      be.free(h->dMovedTris); h->dMovedTris = nullptr;
      BE_CHECK(h, be.alloc(&h->dMovedTris, h->capTris)); h->movedCap = h->capTris;
      if (!h->dMovedCount) BE_CHECK(h, be.alloc(&h->dMovedCount, 1));
    }
    { int rq = ensureBigQueue(h); if (rq) return rq; }      // also learns the SM count
    k_zero_words<<<1, 32, 0, be.stream>>>(h->dMovedCount, 1u);
    k_moved_tris<<<(h->numTris + 255u) / 256u, 256, 0, be.stream>>>(h->dTriInstance, h->dLmMoved, h->numTris, h->dMovedTris, h->dMovedCount);
    be.launches += 2;
  }
#endif
  if (h->lmRes) {
    if (!h->dLmDynamic) BE_CHECK(h, be.alloc(&h->dLmDynamic, (size_t)h->lmLights * 6u * (h->lmRes >> h->lmDynShift) * (h->lmRes >> h->lmDynShift)));
    h->lmDynamicValid = true;
    int rc = rasterLightMap(h, h->dLmDynamic, h->dLmMoved); if (rc) return rc;
  }
  if (h->smValid) {
    if (!h->dSmDynamic) BE_CHECK(h, be.alloc(&h->dSmDynamic, (size_t)(h->smRes >> LM_DYN_SHIFT) * (h->smRes >> LM_DYN_SHIFT)));
    h->smDynamicValid = true;
    int rc = rasterSunMap(h, h->dSmDynamic, h->dLmMoved); if (rc) return rc;
  }
  return 0;
}
// The sun's map follows ubo.lightDir: (re)made at the first trace after build_accel and whenever the direction changed.
int ensureSunMap(anerend_ddgi* h) {
  Backend& be = h->be;
  if (!mapsEnabled(h) || !h->haveUbo) { h->smValid = false; return 0; }
  uint32_t R = 2048;
  if (const char* e = getenv("DDGI_SM_RES")) R = (uint32_t)atoi(e);          // sweeps only; 0 disables the map
  if (R > 8192u) R = 8192u;
  R &= ~((1u << LM_DYN_SHIFT) - 1u);
  if (R == 0) { h->smValid = false; return 0; }
  const float* ld = h->ubo.light_dir;
  if (h->smValid && h->smRes == R && ld[0] == h->smDir[0] && ld[1] == h->smDir[1] && ld[2] == h->smDir[2]) return 0;
  h->smValid = false;
  const V3 sdir = normalize3(-mk3(ld[0], ld[1], ld[2]));
  if (!(fabsf(sdir.x) <= 1.0f && fabsf(sdir.y) <= 1.0f && fabsf(sdir.z) <= 1.0f)) return 0;      // zero / non-finite lightDir: no map
  // instances away from the pose of the static maps: make that pose the current one first (both static maps then hold
  // every triangle where it is now, and the next refit measures "moved" from here)
  if (h->lmMovedValid) { int rc = buildLightMaps(h); if (rc) return rc; }
  const V3 up = fabsf(sdir.y) < 0.9f ? mk3(0.0f, 1.0f, 0.0f) : mk3(1.0f, 0.0f, 0.0f);
  const V3 a = normalize3(mk3(sdir.y * up.z - sdir.z * up.y, sdir.z * up.x - sdir.x * up.z, sdir.x * up.y - sdir.y * up.x));
  const V3 b = mk3(sdir.y * a.z - sdir.z * a.y, sdir.z * a.x - sdir.x * a.z, sdir.x * a.y - sdir.y * a.x);
  float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
  for (int k = 0; k < 8; ++k) {
    const V3 p = mk3((k & 1) ? h->boundsMax[0] : h->boundsMin[0], (k & 2) ? h->boundsMax[1] : h->boundsMin[1], (k & 4) ? h->boundsMax[2] : h->boundsMin[2]);
    const float q[3] = {dot3(p, a), dot3(p, b), dot3(p, sdir)};
    for (int j = 0; j < 3; ++j) { lo[j] = fminf(lo[j], q[j]); hi[j] = fmaxf(hi[j], q[j]); }
  }
  const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), 1e-3f) * 1.001f + 0.01f;
  if (h->smRes != R || !h->dSmStatic) {
    be.free(h->dSmStatic); be.free(h->dSmDynamic); h->dSmStatic = nullptr; h->dSmDynamic = nullptr;
    BE_CHECK(h, be.alloc(&h->dSmStatic, (size_t)R * R));
    h->smRes = R;
  }
  h->sm = SunMapCtx{};
  h->sm.a = a; h->sm.b = b; h->sm.s = sdir; h->sm.cell = ext / (float)R; h->sm.invCell = (float)R / ext;
  h->sm.orgA = lo[0] - 0.005f; h->sm.orgB = lo[1] - 0.005f; h->sm.base = lo[2] - 1.0f;
  for (int k = 0; k < 3; ++k) h->smDir[k] = ld[k];
  h->smDynamicValid = false;
  int rc = rasterSunMap(h, h->dSmStatic, nullptr); if (rc) return rc;
  h->smValid = true;
  return 0;
}

#ifndef DDGI_HOST_EMU
// The last PLOC rounds in ONE block: once at most PLOC_TAIL clusters are left (after ~24 of ~52 rounds at C3) a round is a few
// microseconds of work behind five launches; here the same round bodies run back to back with block barriers, the compaction
// scan done in shared memory.  Does nothing while more clusters are left, so the host can launch it blindly after every batch.
#define PLOC_TAIL 2048
__global__ void __launch_bounds__(1024) k_ploc_tail(BuildCtx c, uint32_t* refA, float4* loA, float4* hiA, uint32_t* refB, float4* loB, float4* hiB) {
  __shared__ uint32_t sWarp[32];
  __shared__ uint32_t sM;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) sM = c.numTris - c.counters[3];
  __syncthreads();
  if (sM > PLOC_TAIL) return;
  const uint32_t lim = min((uint32_t)PLOC_TAIL, c.numTris);      // the cluster arrays hold numTris entries
  int cur = 0;
  for (uint32_t guard = 0; guard < 4096u; ++guard) {
    const uint32_t m = sM;
    if (m <= 1u) break;
    c.clRef = cur ? refB : refA; c.clLo = cur ? loB : loA; c.clHi = cur ? hiB : hiA;
    c.clRefOut = cur ? refA : refB; c.clLoOut = cur ? loA : loB; c.clHiOut = cur ? hiA : hiB;
    for (uint32_t i = tid; i < m; i += 1024u) ploc_nn_body(i, c);
    __syncthreads();
    for (uint32_t i = tid; i < lim; i += 1024u) ploc_merge_body(i, c);             // beyond m: keep = 0
    __syncthreads();
    // exclusive scan of keep[0 .. PLOC_TAIL): two consecutive entries per thread
    const uint32_t k0 = 2u * tid < lim ? c.clKeep[2u * tid] : 0u, k1 = 2u * tid + 1u < lim ? c.clKeep[2u * tid + 1u] : 0u;
    uint32_t v = k0 + k1;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= (uint32_t)o) v += u; }
    if (lane == 31u) sWarp[warp] = v;
    __syncthreads();
    if (warp == 0) { uint32_t w = sWarp[lane]; for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, w, o); if (lane >= (uint32_t)o) w += u; } sWarp[lane] = w; }
    __syncthreads();
    const uint32_t base = (warp ? sWarp[warp - 1u] : 0u) + v - (k0 + k1);
    if (2u * tid < lim) const_cast<uint32_t*>(c.clPos)[2u * tid] = base;
    if (2u * tid + 1u < lim) const_cast<uint32_t*>(c.clPos)[2u * tid + 1u] = base + k0;
    __syncthreads();
    for (uint32_t i = tid; i < m; i += 1024u) ploc_compact_body(i, c);
    __syncthreads();
    if (tid == 0) sM = c.numTris - c.counters[3];
    cur ^= 1;
    __syncthreads();
  }
}
// The BVH8 collapse of a small hierarchy (the instance level of a two-level structure, a small mesh) in ONE block: the levels follow
// each other behind block barriers instead of one launch each (a level is a few dozen nodes there: the launches were the cost).
#define COLLAPSE_SMALL 1024      // measured: 160 primitives 0.38 -> 0.27 ms per rebuild, 1500 the same either way, 4096 slower in one block
__global__ void __launch_bounds__(1024) k_bvh8_collapse_small(BuildCtx c, CItem* qa, CItem* qb) {
  for (uint32_t level = 0; level < (uint32_t)CNT_LEVELS; ++level) {
    const uint32_t count = c.counters[CNT_LEVEL0 + level];
    if (count == 0u) break;                                  // block-uniform
    c.qin = qa; c.qout = qb; c.level = level;
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) collapse_body(i, c);
    __threadfence();
    __syncthreads();
    CItem* t = qa; qa = qb; qb = t;
  }
}
#endif

// Morton codes -> radix sort -> PLOC (or Karras) hierarchy -> BVH8 collapse over `n` primitives whose boxes (c.triLo / c.triHi), 48-byte
// leaf payloads (c.wtri) and bounds (c.sceneBounds) are in place.  c.nodes, c.bvhTris and the per-node arrays point at the first node /
// leaf slot of THIS hierarchy (indices in the nodes are relative to them): the scene's one hierarchy, or one level of a two-level one.
int buildHierarchy(anerend_ddgi* h, BuildCtx c, uint32_t n, uint32_t* numNodesOut, uint32_t* depthOut) {
  Backend& be = h->be;
  const uint32_t nodeCap = n;   // every BVH8 node has >= 2 primitives below it and >= 2 children
  // build-only scratch
  uint32_t *keys = nullptr, *ids = nullptr, *skeys = nullptr, *sids = nullptr, *left = nullptr, *right = nullptr, *parent = nullptr, *rf = nullptr, *rl = nullptr;
  float4 *blo = nullptr, *bhi = nullptr; CItem *qa = nullptr, *qb = nullptr;
  BE_CHECK(h, be.allocScratch(&keys, n)); BE_CHECK(h, be.allocScratch(&ids, n)); BE_CHECK(h, be.allocScratch(&skeys, n)); BE_CHECK(h, be.allocScratch(&sids, n));
  BE_CHECK(h, be.allocScratch(&left, n)); BE_CHECK(h, be.allocScratch(&right, n)); BE_CHECK(h, be.allocScratch(&parent, 2 * (size_t)n)); BE_CHECK(h, be.allocScratch(&rf, n)); BE_CHECK(h, be.allocScratch(&rl, n));
  BE_CHECK(h, be.allocScratch(&blo, 2 * (size_t)n)); BE_CHECK(h, be.allocScratch(&bhi, 2 * (size_t)n)); BE_CHECK(h, be.allocScratch(&qa, n)); BE_CHECK(h, be.allocScratch(&qb, n));
  auto freeScratch = [&]() { be.freeScratch(keys); be.freeScratch(ids); be.freeScratch(skeys); be.freeScratch(sids); be.freeScratch(left); be.freeScratch(right); be.freeScratch(parent);
                             be.freeScratch(rf); be.freeScratch(rl); be.freeScratch(blo); be.freeScratch(bhi); be.freeScratch(qa); be.freeScratch(qb); };
  c.mortonKeys = keys; c.triIds = ids; c.sortedKeys = skeys; c.sortedIds = sids; c.left = left; c.right = right; c.parent = parent;
  c.rangeFirst = rf; c.rangeLast = rl; c.blo = blo; c.bhi = bhi;
  DDGI_LAUNCH(be, k_morton_codes, morton_body, n, c, 256);
  void* sortTmp = nullptr;         // scratch of the radix sort: freed after the build's final synchronisation
  if (!be.sortPairs(keys, skeys, ids, sids, n, &sortTmp)) { be.freeScratch(sortTmp); freeScratch(); return fail(h, ANEREND_DDGI_E_CUDA, std::string("radix sort: ") + be.errStr()); }
  uint32_t plocRadius = PLOC_RADIUS;
  if (const char* e = getenv("DDGI_PLOC_RADIUS")) plocRadius = (uint32_t)atoi(e);     // builder sweeps only (0 = plain LBVH); not part of the ABI
  if (plocRadius == 0) {
    BE_CHECK(h, be.zero(c.visit, ((size_t)n + 1) * sizeof(uint32_t)));
    if (n > 1) DDGI_LAUNCH(be, k_lbvh_hierarchy, karras_body, n - 1, c, 256);
    DDGI_LAUNCH(be, k_lbvh_fit, fit_body, n, c, 256);
  } else {
    // PLOC over the Morton order; scratch: two cluster arrays, neighbour / keep / scan words, subtree counts
    uint32_t *clRef[2] = {nullptr, nullptr}, *nn = nullptr, *keep = nullptr, *pos = nullptr, *sub = nullptr; float4 *clLo[2] = {nullptr, nullptr}, *clHi[2] = {nullptr, nullptr};
    void* scanTmp = nullptr; const size_t scanBytes = be.scanTempBytes(n);
    bool okA = be.allocScratch(&clRef[0], n) && be.allocScratch(&clRef[1], n) && be.allocScratch(&clLo[0], n) && be.allocScratch(&clLo[1], n) && be.allocScratch(&clHi[0], n) && be.allocScratch(&clHi[1], n) &&
               be.allocScratch(&nn, n) && be.allocScratch(&keep, n) && be.allocScratch(&pos, n) && be.allocScratch(&sub, 2 * (size_t)n) && be.allocScratch((uint8_t**)&scanTmp, scanBytes);
    auto freePloc = [&]() { for (int k = 0; k < 2; ++k) { be.freeScratch(clRef[k]); be.freeScratch(clLo[k]); be.freeScratch(clHi[k]); } be.freeScratch(nn); be.freeScratch(keep); be.freeScratch(pos); be.freeScratch(sub); be.freeScratch(scanTmp); };
    uint32_t zero = 0;      // parent[] = MISS everywhere: the root (ref 0) keeps it
    bool okP = okA && be.fill(parent, 0xFF, 2 * (size_t)n * sizeof(uint32_t)) && be.h2d(h->dCounters + 3, &zero, sizeof(uint32_t));
    c.clNN = nn; c.clKeep = keep; c.clPos = pos; c.subCount = sub; c.plocRadius = plocRadius;
    c.clRef = clRef[0]; c.clLo = clLo[0]; c.clHi = clHi[0];
    if (okP) DDGI_LAUNCH(be, k_ploc_init, ploc_init_body, n, c, 256);
    // Rounds are launched in batches without a host round trip: every kernel takes the live cluster count from the device
    // (numTris - counters[3]) and is launched over the last count the host has seen; rounds past the end of the hierarchy do nothing.
    uint32_t m = n, merged = 0; int cur = 0, rounds = 0;
    while (okP && m > 1) {
#ifndef DDGI_HOST_EMU
      const int batch = m <= (uint32_t)PLOC_TAIL ? 0 : 8;      // few clusters (the instance level, a small mesh): the one-block tail does it all
#else
      const int batch = 8;
#endif
      for (int k = 0; k < batch && okP; ++k) {
        c.numClusters = m;
        c.clRef = clRef[cur]; c.clLo = clLo[cur]; c.clHi = clHi[cur]; c.clRefOut = clRef[cur ^ 1]; c.clLoOut = clLo[cur ^ 1]; c.clHiOut = clHi[cur ^ 1];
        DDGI_LAUNCH(be, k_ploc_nn, ploc_nn_body, m, c, 128);
        DDGI_LAUNCH(be, k_ploc_merge, ploc_merge_body, m, c, 128);
        okP = be.exclusiveSum(scanTmp, scanBytes, keep, pos, m);
        DDGI_LAUNCH(be, k_ploc_compact, ploc_compact_body, m, c, 256);
        cur ^= 1; ++rounds;
      }
#ifndef DDGI_HOST_EMU
      if (okP && n > 1) {           // finishes the hierarchy in one launch once few clusters are left (a no-op before that)
        k_ploc_tail<<<1, 1024, 0, be.stream>>>(c, clRef[cur], clLo[cur], clHi[cur], clRef[cur ^ 1], clLo[cur ^ 1], clHi[cur ^ 1]);
        ++be.launches;
      }
#endif
      if (batch == 0) { okP = okP && be.check(); m = 1; break; }      // the tail runs until one cluster is left: nothing to read back
      uint32_t total = 0;
      okP = okP && be.check() && be.d2h(&total, h->dCounters + 3, sizeof(uint32_t));
      if (total <= merged) { okP = false; break; }     // no mutual pair in a whole batch: cannot happen with a consistent tie-break
      m = n - total; merged = total;
    }
    if (okP) {
      c.leafMap = keys; c.orderedIds = ids;            // the unsorted key/id arrays are free after the sort
      DDGI_LAUNCH(be, k_ploc_order, ploc_order_body, 2 * n - 1, c, 256);
      if (n > 1) DDGI_LAUNCH(be, k_ploc_relabel, ploc_relabel_body, n - 1, c, 256);
      c.sortedIds = ids;
      okP = be.check();
    }
    freePloc();            // stream-ordered: released when the kernels above have run
    if (!okP) { freeScratch(); return fail(h, ANEREND_DDGI_E_CUDA, std::string("ploc build: ") + be.errStr()); }
    h->plocRounds = rounds;
  }
  // collapse, one launch per BVH8 level.  The queue length of every level is counted on the device (counters[CNT_LEVEL0 + level]); the
  // host launches a batch of levels over an upper bound (8^level, at most one item per node) and looks once per batch whether
  // the last level still produced work.
  std::vector<uint32_t> counters(CNT_SIZE, 0u);
  counters[0] = 1u; counters[CNT_LEVEL0] = 1u;
  BE_CHECK(h, be.h2d(h->dCounters, counters.data(), CNT_SIZE * sizeof(uint32_t)));
  CItem root; root.ref = 0u; root.node = 0u;
  BE_CHECK(h, be.h2d(qa, &root, sizeof(root)));
  uint32_t missParent = DDGI_MISS;
  BE_CHECK(h, be.h2d(c.nodeParent, &missParent, sizeof(uint32_t)));
  uint32_t depth = 0;
#ifndef DDGI_HOST_EMU
  if (n <= (uint32_t)COLLAPSE_SMALL) {
    k_bvh8_collapse_small<<<1, 1024, 0, be.stream>>>(c, qa, qb);
    ++be.launches;
    if (!be.check() || !be.d2h(counters.data(), h->dCounters, CNT_SIZE * sizeof(uint32_t))) { freeScratch(); return fail(h, ANEREND_DDGI_E_CUDA, std::string("bvh8 collapse: ") + be.errStr()); }
    depth = (uint32_t)CNT_LEVELS;      // every level has run: the loop below has nothing to launch
  }
#endif
  const bool allLevelsDone = depth != 0u;
  uint32_t levelBatch = 4u;      // levels launched per look at the counters: a hierarchy of n primitives is rarely deeper than log8(n) + 3
  for (uint64_t cap8 = 8; cap8 < n && levelBatch < 12u; cap8 *= 8u) ++levelBatch;
  for (bool more = !allLevelsDone; more;) {
    const uint32_t stop = std::min<uint32_t>(depth + levelBatch, (uint32_t)CNT_LEVELS);
    for (; depth < stop; ++depth) {
      uint64_t bound = 1; for (uint32_t k = 0; k < depth && bound < nodeCap; ++k) bound *= 8u;
      c.qin = qa; c.qout = qb; c.level = depth;
      DDGI_LAUNCH(be, k_bvh8_collapse, collapse_body, (uint32_t)std::min<uint64_t>(bound, nodeCap), c, 64);
      std::swap(qa, qb);
    }
    if (!be.check() || !be.d2h(counters.data(), h->dCounters, CNT_SIZE * sizeof(uint32_t))) { freeScratch(); return fail(h, ANEREND_DDGI_E_CUDA, std::string("bvh8 collapse: ") + be.errStr()); }
    more = counters[CNT_LEVEL0 + depth] != 0u && depth < (uint32_t)CNT_LEVELS;
  }
  if (counters[CNT_LEVEL0 + depth] != 0u) { freeScratch(); return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "BVH8 deeper than the collapse supports"); }
  depth = 0; while (depth < (uint32_t)CNT_LEVELS && counters[CNT_LEVEL0 + depth] != 0u) ++depth;        // levels that held work
  be.freeScratch(sortTmp);
  freeScratch();
  if (counters[1] != n) return fail(h, ANEREND_DDGI_E_CUDA, "bvh8 collapse lost primitives");
  *numNodesOut = counters[0]; *depthOut = depth;
  return 0;
}

// ---- two-level structure (ANEREND_DDGI_FLAG_TWO_LEVEL; ddgi_kernels.cuh "Two-level structure") --------------------------------
// The instance level: transforms from the renderable buffer (tlas_update.comp:93-98), world boxes of the meshes' root boxes, and a
// BVH8 with one instance per leaf slot over them.  This is all a moved, shown, hidden or culled instance costs
// (UpdateTLASPass.cpp:198-228 rebuilds the TLAS every frame too).
// what the instance level is built from, staged once per instance set (a refit keeps it)
int uploadTlasInputs(anerend_ddgi* h) {
  Backend& be = h->be;
  const uint32_t ni = (uint32_t)h->instances.size(), nt = (uint32_t)h->tlasList.size();
  std::vector<uint2> ib(ni);
  std::vector<int> slot(h->meshes.size(), -1);              // mesh -> its hierarchy
  for (size_t u = 0; u < h->blas.size(); ++u) slot[h->blas[u].mesh] = (int)u;
  for (uint32_t i = 0; i < ni; ++i) {
    const int u = slot[h->instances[i].mesh];
    ib[i] = u >= 0 ? make_uint2(h->blas[u].root, h->blas[u].triBase) : make_uint2(0u, 0u);
  }
  BE_CHECK(h, be.h2d(h->dInstances, h->instances.data(), ni * sizeof(DInstance)));
  BE_CHECK(h, be.h2d(h->dInstBlas, ib.data(), ni * sizeof(uint2)));
  BE_CHECK(h, be.h2d(h->dTlasList, h->tlasList.data(), nt * sizeof(uint32_t)));
  BE_CHECK(h, be.sync());          // the staging vectors are pageable
  return 0;
}
int rebuildTlas(anerend_ddgi* h) {
  Backend& be = h->be;
  const uint32_t ni = (uint32_t)h->instances.size(), nt = (uint32_t)h->tlasList.size();
  be.mark(4);
  int initBounds[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  BE_CHECK(h, be.h2d(h->dSceneBounds, initBounds, sizeof(initBounds)));
  BuildCtx c{};
  c.renderables = h->dRenderables; c.instances = h->dInstances; c.xforms = h->dXforms; c.numInstances = ni;
  DDGI_LAUNCH(be, k_inst_xform, inst_xform_body, ni, c, 128);
  h->numNodes = 0; h->tlasDepth = 0;
  if (nt == 0) { h->haveAccel = true; return 0; }
  TlasPrimCtx tp{};
  tp.instances = h->dInstances; tp.xforms = h->dXforms; tp.instBlas = h->dInstBlas; tp.tlasList = h->dTlasList; tp.nodeLo = h->dNodeLo; tp.nodeHi = h->dNodeHi;
  tp.numInstances = nt; tp.wtri = h->dWtri; tp.triLo = h->dTriLo; tp.triHi = h->dTriHi; tp.sceneBounds = h->dSceneBounds;
  DDGI_LAUNCH(be, k_tlas_prims, tlas_prim_body, nt, tp, 128);
  c.numTris = nt; c.wtri = h->dWtri; c.triLo = h->dTriLo; c.triHi = h->dTriHi; c.sceneBounds = h->dSceneBounds;
  c.nodes = h->dNodes; c.bvhTris = h->dBvhTris; c.nodeParent = h->dNodeParent; c.nodeLo = h->dNodeLo; c.nodeHi = h->dNodeHi;
  c.nodePending = h->dNodePending; c.visit = h->dVisit; c.counters = h->dCounters; c.leafMax = 1u;
  uint32_t nodes = 0, depth = 0;
  { int rch = buildHierarchy(h, c, nt, &nodes, &depth); if (rch) return rch; }
  be.mark(5);
  int bounds[6];
  if (!be.d2h(bounds, h->dSceneBounds, sizeof(bounds))) return fail(h, ANEREND_DDGI_E_CUDA, std::string("build: ") + be.errStr());
  for (int k = 0; k < 3; ++k) { h->boundsMin[k] = ordered_to_float(bounds[k]); h->boundsMax[k] = ordered_to_float(bounds[3 + k]); }
  h->numNodes = nodes; h->tlasDepth = depth; h->bvhDepth = depth + h->blasDepthMax;
#ifndef DDGI_HOST_EMU
  h->pipeStateDirty = true;
#endif
  // worst case of the traversal stack: per instance level the remaining siblings and the node's own children, eight instance
  // entries, the RESTORE entry, one entry per level of a mesh's hierarchy
  if (2u * depth + 9u + h->blasDepthMax > (uint32_t)DDGI_STACK) return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "two-level structure deeper than the traversal stack");
  h->tlasMs = be.elapsed(4, 5);
  h->haveAccel = true;
  return 0;
}
int buildAccelTwoLevel(anerend_ddgi* h) {
  Backend& be = h->be;
  std::vector<DInstance> inst; uint32_t n = 0;
  int rc = selectInstances(h, inst, n); if (rc) return rc;
  const uint32_t ni = (uint32_t)inst.size();
  // the distinct meshes, in order of first use
  std::vector<anerend_ddgi::MeshBlas> um; std::vector<uint32_t> list;
  std::vector<uint8_t> seenMesh(h->meshes.size(), 0);
  uint32_t total = 0, maxMesh = 0;
  for (uint32_t i = 0; i < ni; ++i) {
    if (inst[i].numTris == 0) continue;
    list.push_back(i);
    if (seenMesh[inst[i].mesh]) continue;
    seenMesh[inst[i].mesh] = 1;
    anerend_ddgi::MeshBlas b{}; b.mesh = inst[i].mesh; b.numTris = inst[i].numTris; b.triBase = total;
    total += b.numTris; maxMesh = std::max(maxMesh, b.numTris); um.push_back(b);
  }
  // the meshes' hierarchies are kept while the same meshes (unchanged vertices) are in use: then only the instance level is rebuilt
  bool reuse = h->twoBuilt && h->tlasCap >= ni && um.size() == h->blas.size() && h->capTextured == h->texturedScene;
  for (size_t u = 0; reuse && u < um.size(); ++u) reuse = um[u].mesh == h->blas[u].mesh && um[u].numTris == h->blas[u].numTris;
  h->haveAccel = false; h->numNodes = 0;
  // one mesh's hierarchy, into its own slots behind the instance level's (UpdateBlasPass.cpp:60-175 / VulkanRenderer.cpp:1267-1417)
  auto buildBlas = [&](anerend_ddgi::MeshBlas& b, uint32_t capI) -> int {
    const anerend_mesh_info_t& mi = h->meshes[b.mesh];
    int initBounds[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    BE_CHECK(h, be.h2d(h->dSceneBounds, initBounds, sizeof(initBounds)));
    const size_t base = (size_t)capI + b.triBase;
    BlasPrimCtx bp{};
    bp.vertices = h->dVertices; bp.indices = h->dIndices; bp.vertexOffset = mi.vertex_offset; bp.indexOffset = mi.index_offset; bp.numTris = b.numTris;
    bp.wtri = h->dWtri; bp.triLo = h->dTriLo; bp.triHi = h->dTriHi; bp.sceneBounds = h->dSceneBounds;
    bp.shadeTris = h->dShadeTris + 5 * (size_t)b.triBase; bp.shadeUV = h->dShadeUV ? h->dShadeUV + 2 * (size_t)b.triBase : nullptr;
    DDGI_LAUNCH(be, k_blas_prims, blas_prim_body, b.numTris, bp, 256);
    BuildCtx c{};
    c.numTris = b.numTris; c.wtri = h->dWtri; c.triLo = h->dTriLo; c.triHi = h->dTriHi; c.sceneBounds = h->dSceneBounds;
    c.nodes = h->dNodes + 5 * base; c.bvhTris = h->dBvhTris + 3 * base; c.nodeParent = h->dNodeParent + base; c.nodeLo = h->dNodeLo + base; c.nodeHi = h->dNodeHi + base;
    c.nodePending = h->dNodePending + base; c.visit = h->dVisit; c.counters = h->dCounters;
    { int rch = buildHierarchy(h, c, b.numTris, &b.nodes, &b.depth); if (rch) return rch; }
    BlasRelocCtx rl{h->dNodes, (uint32_t)base, (uint32_t)base};
    DDGI_LAUNCH(be, k_blas_reloc, blas_reloc_body, b.nodes, rl, 256);
    b.root = (uint32_t)base;
    return 0;
  };
  if (reuse && h->verticesDirty) {
    // deformed meshes (update_vertices): only the hierarchies whose vertex range was written are rebuilt, in place
    for (auto& b : h->blas) {
      const uint32_t v0 = h->meshes[b.mesh].vertex_offset, v1 = v0 + (b.mesh < h->meshNumVertices.size() ? h->meshNumVertices[b.mesh] : 0u);
      bool touched = false;
      for (const auto& d : h->dirtyVerts) touched = touched || (d.first < v1 && d.first + d.second > v0);
      if (touched) { int rcb = buildBlas(b, h->tlasCap); if (rcb) return rcb; }
    }
    h->blasNodes = 0; h->blasDepthMax = 0;
    for (const auto& b : h->blas) { h->blasNodes += b.nodes; h->blasDepthMax = std::max(h->blasDepthMax, b.depth); }
    BE_CHECK(h, be.check());
    h->verticesDirty = false; h->dirtyVerts.clear();
  }
  if (!reuse) {
    freeAccel(h);
    const uint32_t capI = ni + ni / 8u + 16u; const size_t all = (size_t)capI + total; const uint32_t primCap = std::max(capI, maxMesh);
    BE_CHECK(h, be.alloc(&h->dInstances, capI)); BE_CHECK(h, be.alloc(&h->dXforms, capI)); BE_CHECK(h, be.alloc(&h->dInstBlas, capI)); BE_CHECK(h, be.alloc(&h->dTlasList, capI));
    BE_CHECK(h, be.alloc(&h->dWtri, 3 * (size_t)primCap)); BE_CHECK(h, be.alloc(&h->dTriLo, primCap)); BE_CHECK(h, be.alloc(&h->dTriHi, primCap));
    BE_CHECK(h, be.alloc(&h->dNodes, 5 * all)); BE_CHECK(h, be.alloc(&h->dBvhTris, 3 * all));
    BE_CHECK(h, be.alloc(&h->dNodeParent, all)); BE_CHECK(h, be.alloc(&h->dNodeLo, all)); BE_CHECK(h, be.alloc(&h->dNodeHi, all));
    BE_CHECK(h, be.alloc(&h->dNodePending, all)); BE_CHECK(h, be.alloc(&h->dVisit, (size_t)primCap + 1));
    BE_CHECK(h, be.alloc(&h->dShadeTris, 5 * (size_t)total));
    if (h->texturedScene) BE_CHECK(h, be.alloc(&h->dShadeUV, 2 * (size_t)total));
    h->tlasCap = capI; h->capInstances = capI; h->primCap = primCap; h->blasTris = total; h->capTextured = h->texturedScene;
    if (!h->dVisChanged) BE_CHECK(h, be.alloc(&h->dVisChanged, 1));
    be.free(h->dSelected); h->dSelected = nullptr; BE_CHECK(h, be.alloc(&h->dSelected, h->renderables.size()));
    h->blasNodes = 0; h->blasDepthMax = 0;
    for (auto& b : um) {
      int rcb = buildBlas(b, capI); if (rcb) return rcb;
      h->blasNodes += b.nodes; h->blasDepthMax = std::max(h->blasDepthMax, b.depth);
    }
    BE_CHECK(h, be.check());
    h->blas = um; h->twoBuilt = true; h->verticesDirty = false; h->dirtyVerts.clear();
  }
  h->instances = inst; h->numTris = n; h->tlasList = list;
  { std::vector<uint32_t> sel(h->renderables.size());
    for (size_t i = 0; i < sel.size(); ++i) { const uint32_t v = h->renderables[i].visible; sel[i] = ((v & 1u) && !((v >> 1) & 1u)) ? 1u : 0u; }
    BE_CHECK(h, be.h2d(h->dSelected, sel.data(), sel.size() * sizeof(uint32_t))); BE_CHECK(h, be.sync()); }
  if ((rc = uploadTlasInputs(h))) return rc;
  rc = rebuildTlas(h);
  h->accelMs = h->tlasMs;
  return rc;
}

// flatten + LBVH + BVH8 collapse
int buildAccel(anerend_ddgi* h) {
  if (twoLevel(h)) return buildAccelTwoLevel(h);
  Backend& be = h->be;
  std::vector<DInstance> inst; uint32_t n = 0;
  int rc = selectInstances(h, inst, n); if (rc) return rc;
  const uint32_t ni = (uint32_t)inst.size();
  // A rebuild (visibility or cull bits flipped) of about the same size keeps every buffer, light maps included: no allocation, no
  // device-wide synchronisation inside the frame.  Otherwise everything is released and allocated with some head room.
  const bool reuse = h->capTris >= n && h->capInstances >= ni && h->capTris != 0 && h->capTextured == h->texturedScene;
  if (!reuse) freeAccel(h);
  h->haveAccel = false; h->numNodes = 0;
  h->instances = inst; h->numTris = n;
  { // the selection this structure is built from, for refit_accel_device's check of GPU-written visibility / cull bits
    std::vector<uint32_t> sel(h->renderables.size());
    for (size_t i = 0; i < sel.size(); ++i) { const uint32_t v = h->renderables[i].visible; sel[i] = ((v & 1u) && !((v >> 1) & 1u)) ? 1u : 0u; }
    if (!reuse || !h->dSelected) { be.free(h->dSelected); h->dSelected = nullptr; BE_CHECK(h, be.alloc(&h->dSelected, sel.size())); }
    if (!h->dVisChanged) BE_CHECK(h, be.alloc(&h->dVisChanged, 1));
    BE_CHECK(h, be.h2d(h->dSelected, sel.data(), sel.size() * sizeof(uint32_t))); BE_CHECK(h, be.sync()); }
  const uint32_t capN = reuse ? h->capTris : n + n / 8u + 64u, capI = reuse ? h->capInstances : ni + ni / 8u + 16u;
  if (!reuse) { BE_CHECK(h, be.alloc(&h->dInstances, capI)); BE_CHECK(h, be.alloc(&h->dXforms, capI)); BE_CHECK(h, be.alloc(&h->dTriInstance, capN)); }
  BE_CHECK(h, be.h2d(h->dInstances, inst.data(), ni * sizeof(DInstance)));
  BE_CHECK(h, be.sync());          // `inst` (pageable) has been staged
  if (n == 0) { h->numNodes = 0; h->bvhDepth = 0; h->haveAccel = true; return 0; }
  { BuildCtx tc{}; tc.instances = h->dInstances; tc.numInstances = ni; tc.numTris = n; tc.triInstanceOut = h->dTriInstance;      // triangle -> instance, on the device
    DDGI_LAUNCH(be, k_tri_instance, tri_instance_body, n, tc, 256); }
  if (!reuse) {
    BE_CHECK(h, be.alloc(&h->dWtri, 3 * (size_t)capN)); BE_CHECK(h, be.alloc(&h->dTriLo, capN)); BE_CHECK(h, be.alloc(&h->dTriHi, capN));
    BE_CHECK(h, be.alloc(&h->dNodes, 5 * (size_t)capN)); BE_CHECK(h, be.alloc(&h->dBvhTris, 3 * (size_t)capN));
    BE_CHECK(h, be.alloc(&h->dNodeParent, capN)); BE_CHECK(h, be.alloc(&h->dNodeLo, capN)); BE_CHECK(h, be.alloc(&h->dNodeHi, capN));
    BE_CHECK(h, be.alloc(&h->dNodePending, capN)); BE_CHECK(h, be.alloc(&h->dVisit, (size_t)capN + 1));
    BE_CHECK(h, be.alloc(&h->dShadeTris, 5 * (size_t)capN));
    if (h->texturedScene) BE_CHECK(h, be.alloc(&h->dShadeUV, 2 * (size_t)capN));
    h->capTris = capN; h->capInstances = capI; h->capTextured = h->texturedScene;
  }
  int initBounds[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  BE_CHECK(h, be.h2d(h->dSceneBounds, initBounds, sizeof(initBounds)));
  BuildCtx c = buildCtx(h);
  c.shadeTris = h->dShadeTris; c.shadeUV = h->dShadeUV;
  be.mark(4);
  DDGI_LAUNCH(be, k_inst_xform, inst_xform_body, ni, c, 128);
  DDGI_LAUNCH(be, k_flatten_scene, flatten_tri_body, n, c, 256);
  uint32_t numNodes = 0, depth = 0;
  { int rch = buildHierarchy(h, c, n, &numNodes, &depth); if (rch) return rch; }
  { int rcl = buildLightMaps(h); if (rcl) return rcl; }
  be.mark(5);
  int bounds[6];
  if (!be.d2h(bounds, h->dSceneBounds, sizeof(bounds))) return fail(h, ANEREND_DDGI_E_CUDA, std::string("build: ") + be.errStr());
  for (int k = 0; k < 3; ++k) { h->boundsMin[k] = ordered_to_float(bounds[k]); h->boundsMax[k] = ordered_to_float(bounds[3 + k]); }
  h->numNodes = numNodes; h->bvhDepth = depth;
  if (depth > DDGI_STACK) return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "BVH8 deeper than the traversal stack");
  h->accelMs = be.elapsed(4, 5);
  h->haveAccel = true; h->verticesDirty = false; h->dirtyVerts.clear();
#ifndef DDGI_HOST_EMU
  h->pipeStateDirty = true;
#endif
  return 0;
}

int refitAccel(anerend_ddgi* h) {
  Backend& be = h->be;
  if (h->numTris == 0) return 0;
#ifndef DDGI_HOST_EMU
  h->pipeStateDirty = true;          // the next pipelined trace waits for this refit
#endif
  if (twoLevel(h)) {          // moved instances: the instance level is rebuilt; deformed vertices: their meshes' hierarchies too
    int rc2 = h->verticesDirty ? buildAccelTwoLevel(h) : rebuildTlas(h);
    h->timed[3] = true;
    return rc2;
  }
  BuildCtx c = buildCtx(h);
  c.sceneBounds = nullptr;
  const bool verticesChanged = h->verticesDirty;
  if (h->verticesDirty) { c.shadeTris = h->dShadeTris; c.shadeUV = h->dShadeUV; h->verticesDirty = false; h->dirtyVerts.clear(); }   // deformed meshes: refresh the shading records too
  else {
    // moving nodes only: triangles and leaf copies are recomputed where an instance's transform changed (C4: 16 movers of 161 instances)
    if (!h->dInstChanged) BE_CHECK(h, be.alloc(&h->dInstChanged, (size_t)h->capInstances));
    c.instChanged = h->dInstChanged;
  }
  be.mark(4);
  DDGI_LAUNCH(be, k_inst_xform, inst_xform_body, (uint32_t)h->instances.size(), c, 128);
  DDGI_LAUNCH(be, k_flatten_scene, flatten_tri_body, h->numTris, c, 256);
  DDGI_LAUNCH(be, k_refit_tris, refit_tris_body, h->numTris, c, 256);
  BE_CHECK(h, be.zero(h->dVisit, ((size_t)h->numNodes + 1) * sizeof(uint32_t)));
  DDGI_LAUNCH(be, k_refit_nodes, refit_node_body, h->numNodes, c, 128);
  { int rcl = refitLightMaps(h, verticesChanged); if (rcl) return rcl; }
  be.mark(5);
  BE_CHECK(h, be.check());
  h->timed[3] = true;
  return 0;
}

#ifdef DDGI_HOST_EMU
// mat3(randomRot) * dir for every ray of `layer` into the pinned staging table (rgen:87-88)
void rotateDirs(anerend_ddgi* h, int layer, const float* m) {
  float4* out = h->hDirs + (size_t)layer * h->N;
  for (int i = 0; i < h->N; ++i) {
    float x = h->fib[3 * i], y = h->fib[3 * i + 1], z = h->fib[3 * i + 2];
    out[i] = make_float4(m[0] * x + m[4] * y + m[8] * z, m[1] * x + m[5] * y + m[9] * z, m[2] * x + m[6] * y + m[10] * z, 1.0f);
  }
}
#endif

#ifndef DDGI_HOST_EMU
__global__ void __launch_bounds__(256) k_zero_words(uint32_t* p, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0u;
}
// scratch of the wavefront kernels (hit records, surface records, visibility words, per-light queues, hit list); grows on demand
int ensureFastScratch(anerend_ddgi* h, size_t need, uint32_t nq) {
  Backend& be = h->be;
  if (h->fastRays < need || h->fastQueues < nq) {
    BE_CHECK(h, be.sync());
    be.free(h->dHits); be.free(h->dSurf); be.free(h->dLit); be.free(h->dQueues); be.free(h->dQcount); be.free(h->dHitq);
    h->dHits = nullptr; h->dSurf = nullptr; h->dLit = nullptr; h->dQueues = nullptr; h->dQcount = nullptr; h->dHitq = nullptr; h->fastRays = 0;
    BE_CHECK(h, be.alloc(&h->dHits, need)); BE_CHECK(h, be.alloc(&h->dSurf, 3 * need)); BE_CHECK(h, be.alloc(&h->dLit, need));
    BE_CHECK(h, be.alloc(&h->dQueues, (size_t)nq * need)); BE_CHECK(h, be.alloc(&h->dQcount, FAST_QC_SIZE)); BE_CHECK(h, be.alloc(&h->dHitq, need));
    be.free(h->dHitqSorted); h->dHitqSorted = nullptr; BE_CHECK(h, be.alloc(&h->dHitqSorted, need));
    be.free(h->dHitInst); h->dHitInst = nullptr; if (twoLevel(h)) BE_CHECK(h, be.alloc(&h->dHitInst, need));
    if (h->pStream) BE_CHECK(h, be.ok(cudaStreamSynchronize(h->pStream)));
    be.free(h->dHits2); be.free(h->dHitInst2); h->dHits2 = nullptr; h->dHitInst2 = nullptr; h->pipeDoneValid[0] = h->pipeDoneValid[1] = false;
    h->fastRays = need; h->fastQueues = nq;
    be.free(h->dEmisTex); h->dEmisTex = nullptr;
  }
  if (h->emissiveTextures && !h->dEmisTex) BE_CHECK(h, be.alloc(&h->dEmisTex, h->fastRays));     // per-ray emissive of textured materials
  return 0;
}
int ensureFastGrids(anerend_ddgi* h) {
  Backend& be = h->be;
  if (h->gridPrimary[0] == 0) {
    int dev = 0, sms = 0, nb = 0;
    BE_CHECK(h, be.ok(cudaGetDevice(&dev))); BE_CHECK(h, be.ok(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)));
    h->sms = sms;
    if (twoLevel(h)) {
      BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<false, false, false, true>, 128, 0))); h->gridPrimary[0] = sms * (nb > 0 ? nb : 1);
      BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<false, true, false, true>, 128, 0))); h->gridPrimary[1] = sms * (nb > 0 ? nb : 1);
      BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<true, false, false, true>, 128, 0))); h->gridShadow[0] = sms * (nb > 0 ? nb : 1);
      BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<true, true, false, true>, 128, 0))); h->gridShadow[1] = sms * (nb > 0 ? nb : 1);
    } else {
    BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<false, false>, 128, 0))); h->gridPrimary[0] = sms * (nb > 0 ? nb : 1);
    BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<false, true>, 128, 0))); h->gridPrimary[1] = sms * (nb > 0 ? nb : 1);
    BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<true, false>, 128, 0))); h->gridShadow[0] = sms * (nb > 0 ? nb : 1);
    BE_CHECK(h, be.ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<true, true>, 128, 0))); h->gridShadow[1] = sms * (nb > 0 ? nb : 1);
    }
    h->tune[0] = TUNE_NODE_STEPS; h->tune[1] = TUNE_REFILL_MIN; h->tune[2] = TUNE_REFILL_MIN; h->tune[3] = 0;
    h->tune[4] = TUNE_STEP_MIN; h->tune[5] = TUNE_TRI_MIN; h->tune[6] = 0; h->tune[7] = 0;
    if (const char* e = getenv("DDGI_TUNE")) {          // kernel-tuning sweeps only (tools/); not part of the ABI
      unsigned v[8]; for (int k = 0; k < 8; ++k) v[k] = h->tune[k];
      sscanf(e, "%u,%u,%u,%u,%u,%u,%u,%u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
      for (int k = 0; k < 8; ++k) h->tune[k] = v[k];
    }
  }
  return 0;
}
// Pipelined updates: see the handle.  Off while anything needs the kernels of an update in one stream order (per-kernel or per-pass
// timers, counters, a connected peer exchange).
bool usePipeline(const anerend_ddgi* h) {
  // (with a connected peer exchange too: the closest-hit trace touches neither the atlas nor the barrier flags)
  return h->pipeline && h->be.timerLevel == 0 && !(h->p.flags & (ANEREND_DDGI_FLAG_SIMPLE_KERNELS | ANEREND_DDGI_FLAG_COUNTERS));
}
int ensurePipeline(anerend_ddgi* h) {
  Backend& be = h->be;
  if (!h->pStream) {
    BE_CHECK(h, be.ok(cudaStreamCreateWithFlags(&h->pStream, cudaStreamNonBlocking)));
    for (int k = 0; k < 2; ++k) {
      BE_CHECK(h, be.ok(cudaEventCreateWithFlags(&h->evTrace[k], cudaEventDisableTiming))); BE_CHECK(h, be.ok(cudaEventCreateWithFlags(&h->evHitsFree[k], cudaEventDisableTiming)));
      BE_CHECK(h, be.alloc(&h->dDirsP[k], (size_t)h->nY * h->N)); BE_CHECK(h, be.alloc(&h->dLayerMapP[k], (size_t)h->nY));
    }
    BE_CHECK(h, be.ok(cudaEventCreateWithFlags(&h->evState, cudaEventDisableTiming)));
    BE_CHECK(h, be.alloc(&h->dCursorP, 128));
  }
  if (!h->dHits2) { BE_CHECK(h, be.alloc(&h->dHits2, h->fastRays)); if (twoLevel(h)) BE_CHECK(h, be.alloc(&h->dHitInst2, h->fastRays)); }
  return 0;
}
// Wavefront path: k_trace_primary -> k_shade_prep -> k_trace_shadow -> k_shade (ddgi_fast.cuh)
// pbuf >= 0: pipelined — the closest-hit trace runs on the pipeline stream with buffer set pbuf (traceLayers has put its directions there)
int traceFast(anerend_ddgi* h, const TraceCtx& c, uint32_t n, int pbuf = -1) {
  Backend& be = h->be;
  const uint32_t nq = 1u + ((h->p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) ? h->numLights : 0u);
  { int rcs = ensureFastScratch(h, n, nq); if (rcs) return rcs; }
  { int rcg = ensureFastGrids(h); if (rcg) return rcg; }
  FastCtx f{};
  f.t = c; f.shadeTris = h->dShadeTris; f.shadeUV = h->dShadeUV; f.emisTex = h->dEmisTex; f.hits = h->dHits; f.surf = h->dSurf; f.lit = h->dLit; f.queues = h->dQueues; f.hitq = h->dHitq;
  f.qcount = h->dQcount; f.totalRays = n; f.numQueues = nq; f.hitInst = h->dHitInst;
  if (pbuf == 1) { f.hits = h->dHits2; f.hitInst = h->dHitInst2; }
  f.surfStride = 3u; f.rawSunDir = 0u;
  for (int k = 0; k < 8; ++k) f.tune[k] = h->tune[k];
  f.tune[3] = 0x64646464u;                 // planes2: the PRMT constant travels as a kernel argument so the selector can be the immediate
  // queue counters and cursors: cleared by a kernel, not cudaMemsetAsync (a memset may be scheduled on a copy engine, where
  // it waits behind a background read-back of the atlas)
  k_zero_words<<<(FAST_QC_SIZE + 255u) / 256u, 256, 0, be.stream>>>(h->dQcount, (uint32_t)FAST_QC_SIZE);
  ++be.launches;
  const bool cnt = (h->p.flags & ANEREND_DDGI_FLAG_COUNTERS) != 0;
  static const int prepBlock = getenv("DDGI_PREP_BLOCK") ? atoi(getenv("DDGI_PREP_BLOCK")) : PREP_BLOCK_DEFAULT;     // sweeps only
  const int gp = h->gridPrimary[cnt ? 1 : 0], gs = h->gridShadow[cnt ? 1 : 0];
  const bool two = twoLevel(h);
  be.mark(6);
  if (pbuf >= 0) {
    FastCtx fb = f;
    fb.t.dirs = h->dDirsP[pbuf]; fb.t.layerMap = h->dLayerMapP[pbuf]; fb.primaryCursor = h->dCursorP + 64 * pbuf;
    k_zero_words<<<1, 32, 0, h->pStream>>>(fb.primaryCursor, 1u);
    if (two) k_trace<false, false, false, true><<<gp, 128, 0, h->pStream>>>(fb); else k_trace<false, false><<<gp, 128, 0, h->pStream>>>(fb);
    ++be.launches;
    BE_CHECK(h, be.ok(cudaEventRecord(h->evTrace[pbuf], h->pStream)));
    BE_CHECK(h, be.ok(cudaStreamWaitEvent(be.stream, h->evTrace[pbuf], 0)));
  } else
  if (two) { if (cnt) k_trace<false, true, false, true><<<gp, 128, 0, be.stream>>>(f); else k_trace<false, false, false, true><<<gp, 128, 0, be.stream>>>(f); }
  else if (cnt) k_trace<false, true><<<gp, 128, 0, be.stream>>>(f); else k_trace<false, false><<<gp, 128, 0, be.stream>>>(f);
  be.mark(7);
  // default: 32 registers (2048 resident threads per SM), light-map look-ups two at a time — g7 sweep at C3: 0.125 ms against 0.128
  // (look-ups one by one), 0.136 (40 registers), 0.173 (56 registers), 0.156 / 0.176 (hits compacted per block first, 128 / 256 threads)
  // blocks of 128: the queue append ends in a block barrier, and a block waits for its slowest warp (0.117 ms against 0.123 with 256).
  // A 16-bit copy of the bounds for these look-ups (25 MB instead of 100 MB behind them) was tried: 0.118 ms, the same — the look-ups wait
  // for latency, not for bandwidth or L2 space.
  static const int prepVariant = getenv("DDGI_PREP_VARIANT") ? atoi(getenv("DDGI_PREP_VARIANT")) : 7;                  // sweeps only
  if (two) k_shade_prep<128, false, 2048, true, true><<<(n + 127u) / 128u, 128, 0, be.stream>>>(f);
  else if (prepVariant == 1 && prepBlock == 256) k_shade_prep<256, false, 2048, true><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
  else if (prepVariant == 7) k_shade_prep<128, false, 2048, true><<<(n + 127u) / 128u, 128, 0, be.stream>>>(f);
  else if (prepVariant == 8) k_shade_prep<64, false, 2048, true><<<(n + 63u) / 64u, 64, 0, be.stream>>>(f);
  else if (prepVariant == 2) k_shade_prep<256, false, 1536, false><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
  else if (prepVariant == 3) k_shade_prep<256, false, 1536, true><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
  else if (prepVariant == 4) k_shade_prep<256, false, 1024, true><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
  else if (prepVariant == 5) k_shade_prep_compact<256><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
  else if (prepVariant == 6) k_shade_prep_compact<128><<<(n + 127u) / 128u, 128, 0, be.stream>>>(f);
  else if (prepBlock == 128) k_shade_prep<128><<<(n + 127u) / 128u, 128, 0, be.stream>>>(f);
  else if (prepBlock == 512) k_shade_prep<512><<<(n + 511u) / 512u, 512, 0, be.stream>>>(f);
  else k_shade_prep<256><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
  be.mark(8);
  if (two) { if (cnt) k_trace<true, true, false, true><<<gs, 128, 0, be.stream>>>(f); else k_trace<true, false, false, true><<<gs, 128, 0, be.stream>>>(f); }
  else if (cnt) k_trace<true, true><<<gs, 128, 0, be.stream>>>(f); else k_trace<true, false><<<gs, 128, 0, be.stream>>>(f);
  be.mark(9);
  static const bool bucketHits = !getenv("DDGI_NO_HIT_BUCKETS");          // A/B measurements only
  if (bucketHits && (h->p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) && h->numLights > 1u) {
    f.hitqSorted = h->dHitqSorted;
    k_hit_bucket<0><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
    k_hit_bucket<1><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
    be.launches += 2;
  }
  { int rf = fenceExchange(h); if (rf) return rf; }      // the only kernel of the pass that samples the atlas
  BE_CHECK(h, be.ok((h->p.flags & ANEREND_DDGI_FLAG_FAST_MATH) ? launch_shade_fast(f, be.stream) : launch_shade_exact(f, be.stream)));
  if (pbuf >= 0) { BE_CHECK(h, be.ok(cudaEventRecord(h->evHitsFree[pbuf], be.stream))); h->pipeDoneValid[pbuf] = true; }      // the hit records of this set are free again
  be.mark(10);
  h->timedKernels = be.timerLevel >= 2;
  be.launches += 5;
  BE_CHECK(h, be.check());
  return 0;
}
#endif

#ifndef DDGI_HOST_EMU
// Cross-GPU barrier between the ranks of a connected exchange: thread r announces this rank's arrival in rank r's flag
// array (a store through the NVLink-mapped pointer), then waits until rank r has announced itself here.  All ranks issue
// the same sequence of barriers, so `epoch` is the same number everywhere; flags only grow.  A rank that never arrives
// (a crashed peer) trips the time-out instead of hanging the GPU.
struct PeerBarrierCtx { uint32_t* flags[8]; uint32_t* err; uint32_t rank, world, epoch; unsigned long long timeoutNs; };
__global__ void __launch_bounds__(32) k_peer_barrier(const PeerBarrierCtx b) {
  const uint32_t r = threadIdx.x;
  if (r < b.world) {
    __threadfence_system();                                   // this rank's earlier stores (peer atlases) before the flag
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(b.flags[r] + b.rank), "r"(b.epoch) : "memory");
    const uint32_t* mine = b.flags[b.rank] + r;
    unsigned long long t0, t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - b.epoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > b.timeoutNs) { *b.err = 1u + r; break; }
      __nanosleep(200);
    }
  }
  __syncwarp();
  __threadfence_system();
}
// Push of this rank's freshly convolved layers into every peer's atlas over NVLink: whole layer blocks (10*nXZ atlas rows,
// contiguous) as 16-byte loads and stores, one load feeding all peers.  The per-texel stores of the convolution reach
// the link as 64-byte runs (8 interior texels of a probe row); full lines move ~2x the bytes per unit time.
struct PushCtx { const uint4* src; uint4* peers[8]; const int* layerMap; uint32_t numPeers, chunksPerRow, rows, rowOffset, rowStride, nXZ; };
__global__ void __launch_bounds__(256) k_atlas_push(const PushCtx c) {
  // a probe row z of a layer = 10 atlas rows = chunksPerRow 16-byte chunks, contiguous; this handle owns `rows` of them
  const size_t layerBase = (size_t)c.layerMap[blockIdx.y] * c.nXZ * c.chunksPerRow;
  const uint32_t total = c.rows * c.chunksPerRow;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t zr = i / c.chunksPerRow, w = i % c.chunksPerRow;
    const size_t o = layerBase + (size_t)(c.rowOffset + zr * c.rowStride) * c.chunksPerRow + w;
    const uint4 v = __ldcg(c.src + o);
    for (uint32_t p = 0; p < c.numPeers; ++p) __stcg(c.peers[p] + o, v);
  }
}
int pushLayers(anerend_ddgi* h, int numLayers, cudaStream_t stream = nullptr) {
  if (h->numPeers == 0 || numLayers == 0) return 0;
  if (!stream) stream = h->be.stream;
  PushCtx c{};
  c.src = (const uint4*)h->dAtlas; c.layerMap = h->dLayerMap; c.numPeers = (uint32_t)h->numPeers;
  for (int p = 0; p < h->numPeers; ++p) c.peers[p] = (uint4*)h->peers[p];
  c.chunksPerRow = (uint32_t)((size_t)10 * h->W * sizeof(uint2) / sizeof(uint4));          // W = 10*nXZ is even
  c.rows = (uint32_t)(h->Pl / h->nXZ); c.rowOffset = (uint32_t)h->rowOffset; c.rowStride = (uint32_t)h->rowStride; c.nXZ = (uint32_t)h->nXZ;
  const uint32_t bx = std::min<uint32_t>((c.rows * c.chunksPerRow + 255u) / 256u, (uint32_t)std::max(1, 4 * h->sms / numLayers));
  k_atlas_push<<<dim3(bx, (uint32_t)numLayers), 256, 0, stream>>>(c);
  ++h->be.launches;
  BE_CHECK(h, h->be.check());
  return 0;
}
int peerBarrier(anerend_ddgi* h, cudaStream_t stream = nullptr) {
  if (h->xworld < 2) return 0;
  if (!stream) stream = h->be.stream;
  PeerBarrierCtx b{};
  for (uint32_t r = 0; r < h->xworld; ++r) b.flags[r] = h->xflags[r];
  b.err = h->dXerr; b.rank = h->xrank; b.world = h->xworld; b.epoch = ++h->xepoch; b.timeoutNs = 30ull * 1000000000ull;
  k_peer_barrier<<<1, 32, 0, stream>>>(b);
  ++h->be.launches;
  BE_CHECK(h, h->be.check());
  return 0;
}
int peerCheck(anerend_ddgi* h) {        // after a stream sync: did a barrier time out?
  if (h->xworld < 2) return 0;
  uint32_t e = 0;
  BE_CHECK(h, h->be.d2h(&e, h->dXerr, sizeof(e)));
  if (e) return fail(h, ANEREND_DDGI_E_STATE, "exchange barrier timed out waiting for rank " + std::to_string(e - 1));
  return 0;
}
// Whatever samples or rewrites the whole atlas waits (on the device) for the exchange of the previous update to have landed.
int fenceExchange(anerend_ddgi* h) {
  if (!h->exchangePending) return 0;
  BE_CHECK(h, h->be.ok(cudaStreamWaitEvent(h->be.stream, h->evXDone, 0)));
  return 0;
}
void exchangeDisconnect(anerend_ddgi* h) {
  h->be.sync();
  if (h->xStream) cudaStreamSynchronize(h->xStream);
  h->exchangePending = false;
  for (int i = 0; i < h->numOpened; ++i) if (h->xopened[i]) cudaIpcCloseMemHandle(h->xopened[i]);
  h->numOpened = 0; h->numPeers = 0; h->xworld = 0; h->xrank = 0;
  for (auto& f : h->xflags) f = nullptr;
}
#else
int peerBarrier(anerend_ddgi*) { return 0; }
int fenceExchange(anerend_ddgi*) { return 0; }
int pushLayers(anerend_ddgi*, int) { return 0; }
int peerCheck(anerend_ddgi*) { return 0; }
void exchangeDisconnect(anerend_ddgi*) {}
#endif

// Everything of a TraceCtx that does not depend on which layers are launched
int texturesUsable(anerend_ddgi* h) {
  if (h->texturesStale) return fail(h, ANEREND_DDGI_E_STATE, "upload_textures was called after upload_scene: upload the scene again so its texture indices are re-checked");
  return 0;
}
TraceCtx sceneTraceCtx(anerend_ddgi* h) {
  TraceCtx c{};
  c.nodes = h->dNodes; c.bvhTris = h->dBvhTris; c.numNodes = h->numNodes; c.twoLevel = twoLevel(h) ? 1u : 0u;
  c.vertices = h->dVertices; c.indices = h->dIndices; c.materials = h->dMaterials;
  c.tex.textures = h->dTextures; c.tex.texPool = h->dTexPool; c.tex.srgbLut = h->dSrgbLut; c.tex.numTextures = h->numTextures;
  c.tex.tileInfos = h->dTileInfos; c.tex.numTiles = h->numTiles;
  c.cameraPos = mk3(h->ubo.camera_pos[0], h->ubo.camera_pos[1], h->ubo.camera_pos[2]);
  c.instances = h->dInstances; c.xforms = h->dXforms; c.triInstance = h->dTriInstance;
  c.lights = h->dLights; c.numLights = h->numLights;
  c.dirs = h->dDirs; c.layerMap = h->dLayerMap; c.rayOrder = h->dRayOrder; c.atlas = h->dAtlas; c.rayIrr = h->dRayIrr;
  c.records = h->dRecords; c.stats = h->dStats;
  c.nXZ = h->nXZ; c.nY = h->nY; c.N = h->N; c.W = h->W; c.H = h->H; c.numLayers = 0;
  c.Pl = h->Pl; c.rowOffset = h->rowOffset; c.rowStride = h->rowStride;
  gridFrame(h, c.camOff, c.step);
  // rmiss:20 toLinear(skyIntensity * skyColor), constant per frame
  { float s[3]; for (int k = 0; k < 3; ++k) { float v = h->ubo.sky_intensity * h->ubo.sky_color[k];
      float higher = powf((v + 0.055f) / 1.055f, 2.4f), lower = v / 12.92f; s[k] = (v < 0.04045f) ? lower : higher; }
    c.sky = mk3(s[0], s[1], s[2]); }
  c.lightDir = mk3(h->ubo.light_dir[0], h->ubo.light_dir[1], h->ubo.light_dir[2]);
  c.tmin = h->p.t_min; c.tmax = h->p.t_max; c.sunIntensity = h->ubo.sun_intensity;
  c.uboFlags = h->ubo.flags; c.flags = h->p.flags;
  c.lmRes = h->lmRes; c.lmStatic = h->dLmStatic; c.lmDynamic = h->lmDynamicValid ? h->dLmDynamic : nullptr; c.lmNear = h->dLmNear;
  c.lmTris = h->dWtri; c.lmNumTris = h->numTris; c.lmDynShift = h->lmDynShift; c.smDynShift = LM_DYN_SHIFT;
  // the sun's map only while it is the map of THIS frame's lightDir (ensureSunMap runs in front of the probe trace)
  if (h->smValid && h->ubo.light_dir[0] == h->smDir[0] && h->ubo.light_dir[1] == h->smDir[1] && h->ubo.light_dir[2] == h->smDir[2]) {
    c.smRes = h->smRes; c.smStatic = h->dSmStatic; c.smDynamic = h->smDynamicValid ? h->dSmDynamic : nullptr;
    c.smA = h->sm.a; c.smB = h->sm.b; c.smS = h->sm.s; c.smOrgA = h->sm.orgA; c.smOrgB = h->sm.orgB; c.smInvCell = h->sm.invCell; c.smBase = h->sm.base;
  } else { c.smRes = 0; c.smS = mk3(0.0f, 1.0f, 0.0f); }
  return c;
}

#ifndef DDGI_HOST_EMU
#define DIRS_BATCH 32
struct DirsCtx { float m[DIRS_BATCH][9]; int layer[DIRS_BATCH]; const float* fib; float4* dDirs; int* dLayerMap; uint32_t N, first, count; };
// mat3(randomRot) * worldDirFromFibonacciIndex(i) (rgen:87-88) for the layers of one launch + their entries of the layer map
__global__ void __launch_bounds__(128) k_make_dirs(const DirsCtx c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (i >= c.N) return;
  const float* m = c.m[k];
  const float x = c.fib[3 * i], y = c.fib[3 * i + 1], z = c.fib[3 * i + 2];
  c.dDirs[(size_t)c.layer[k] * c.N + i] = make_float4(m[0] * x + m[3] * y + m[6] * z, m[1] * x + m[4] * y + m[7] * z, m[2] * x + m[5] * y + m[8] * z, 1.0f);
  if (i == 0) c.dLayerMap[c.first + k] = c.layer[k];
}
#endif

// `layers`: the absolute layers of this launch (any subset of the grid, each at most once); rots + 16*rotStride*i is
// the randomRot of layers[i] (rotStride 0: one matrix, 1: packed per launched layer; rotsByLayer: indexed by absolute layer)
int traceLayers(anerend_ddgi* h, const int* layers, int numLayers, const float* rots, bool rotsByLayer) {
  Backend& be = h->be;
  if (!h->haveAccel) return fail(h, ANEREND_DDGI_E_STATE, "trace before build_accel");
  if (!h->haveUbo) return fail(h, ANEREND_DDGI_E_STATE, "trace before set_scene_ubo");
  { int rct = texturesUsable(h); if (rct) return rct; }
  be.waitCopy();            // a pending H2D of the layer map (convolveLayers) must be done before the staging array is rewritten
  for (int l = 0; l < numLayers; ++l) h->hLayerMap[l] = layers[l];
#ifdef DDGI_HOST_EMU
  int loLayer = h->nY, hiLayer = -1;
  for (int l = 0; l < numLayers; ++l) {
    const int layer = layers[l];
    rotateDirs(h, layer, rots + 16 * (size_t)(rotsByLayer ? layer : l));
    loLayer = layer < loLayer ? layer : loLayer; hiLayer = layer > hiLayer ? layer : hiLayer;
  }
#endif
#ifndef DDGI_HOST_EMU
  // Direction table and layer map are made ON the device from the rotation matrices, which travel as kernel arguments:
  // no host memory is read at the start of an update.  (A host->device copy, or a kernel reading page-locked memory,
  // queues up behind a background read-back of the atlas — PCIe reads may not pass the posted writes of that copy — and
  // held the whole update back by the length of the copy, 85 us at C3.)  Same fp32 operations as rotateDirs, no contraction.
  int pbuf = -1;
  if (usePipeline(h) && h->haveAccel) {
    const uint32_t np = (uint32_t)numLayers * (uint32_t)h->Pl * (uint32_t)h->N, nq = 1u + ((h->p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) ? h->numLights : 0u);
    { int rcs = ensureFastScratch(h, np, nq); if (rcs) return rcs; }
    { int rcp = ensurePipeline(h); if (rcp) return rcp; }
    pbuf = (int)(h->pipeIdx++ & 1u);
    if (h->pipeDoneValid[pbuf]) BE_CHECK(h, be.ok(cudaStreamWaitEvent(h->pStream, h->evHitsFree[pbuf], 0)));      // the update before last has read its hit records
    if (h->pipeStateDirty) {      // the acceleration structure was (re)built or refitted on the main stream since the last trace
      BE_CHECK(h, be.ok(cudaEventRecord(h->evState, be.stream))); BE_CHECK(h, be.ok(cudaStreamWaitEvent(h->pStream, h->evState, 0)));
      h->pipeStateDirty = false;
    }
  } else h->pipeStateDirty = true;      // an update whose trace ran on the main stream used hit-record set 0 there: the next pipelined one waits for it
  for (int l0 = 0; l0 < numLayers; l0 += DIRS_BATCH) {
    DirsCtx dc{};
    dc.fib = h->dFib; dc.dDirs = h->dDirs; dc.dLayerMap = h->dLayerMap; dc.N = (uint32_t)h->N; dc.first = (uint32_t)l0;
    dc.count = (uint32_t)std::min(DIRS_BATCH, numLayers - l0);
    for (uint32_t k = 0; k < dc.count; ++k) {
      const int layer = layers[l0 + (int)k];
      const float* m = rots + 16 * (size_t)(rotsByLayer ? layer : l0 + (int)k);
      dc.layer[k] = layer;
      for (int c3 = 0; c3 < 3; ++c3) for (int r3 = 0; r3 < 3; ++r3) dc.m[k][c3 * 3 + r3] = m[c3 * 4 + r3];
    }
    k_make_dirs<<<dim3(((uint32_t)h->N + 127u) / 128u, dc.count), 128, 0, be.stream>>>(dc);
    ++be.launches;
    if (pbuf >= 0) {      // the closest-hit trace's own copy, on its stream
      dc.dDirs = h->dDirsP[pbuf]; dc.dLayerMap = h->dLayerMapP[pbuf];
      k_make_dirs<<<dim3(((uint32_t)h->N + 127u) / 128u, dc.count), 128, 0, h->pStream>>>(dc);
      ++be.launches;
    }
  }
#else
  if (hiLayer >= loLayer)
    BE_CHECK(h, be.h2d(h->dDirs + (size_t)loLayer * h->N, h->hDirs + (size_t)loLayer * h->N, (size_t)(hiLayer - loLayer + 1) * h->N * sizeof(float4)));
  BE_CHECK(h, be.h2d(h->dLayerMap, h->hLayerMap, (size_t)numLayers * sizeof(int)));
#endif
  h->numLaunchLayers = numLayers;
  { int rcs = ensureSunMap(h); if (rcs) return rcs; }
  TraceCtx c = sceneTraceCtx(h); c.numLayers = numLayers;
  const uint32_t n = (uint32_t)numLayers * (uint32_t)h->Pl * (uint32_t)h->N;
#ifndef DDGI_HOST_EMU
  if (!(h->p.flags & ANEREND_DDGI_FLAG_SIMPLE_KERNELS)) { int rc = traceFast(h, c, n, pbuf); if (rc) return rc; } else
#endif
  {
    { int rf = fenceExchange(h); if (rf) return rf; }
    if (h->p.flags & ANEREND_DDGI_FLAG_COUNTERS) DDGI_LAUNCH(be, k_probe_trace_counted, trace_ray_body<true>, n, c, 128);
    else DDGI_LAUNCH(be, k_probe_trace, trace_ray_body<false>, n, c, 128);
    BE_CHECK(h, be.check());
  }
  for (int l = 0; l < numLayers; ++l) h->layerTraced[layers[l]] = 1;
  return 0;
}

bool texelStoreExchange() { static const bool on = getenv("DDGI_EXCHANGE_TEXEL") != nullptr; return on; }      // A/B measurements only
// `push`: store the fresh layers into the peers' atlases right after the convolution (the caller has made sure no peer still reads them)
int convolveLayers(anerend_ddgi* h, const int* layers, int numLayers, bool push = true) {
  Backend& be = h->be;
  BE_CHECK(h, be.fenceDownload());          // a background read-back of the previous atlas must finish before it is overwritten
  for (int l = 0; l < numLayers; ++l) if (!h->layerTraced[layers[l]]) return fail(h, ANEREND_DDGI_E_STATE, "convolve before trace of this layer");
  bool same = numLayers == h->numLaunchLayers;
  for (int l = 0; same && l < numLayers; ++l) same = h->hLayerMap[l] == layers[l];
  if (!same) {              // convolve of a different layer set than the last trace: restage the layer map
    be.waitCopy();
    for (int l = 0; l < numLayers; ++l) h->hLayerMap[l] = layers[l];
    BE_CHECK(h, be.h2d(h->dLayerMap, h->hLayerMap, (size_t)numLayers * sizeof(int)));
    h->numLaunchLayers = numLayers;
    be.markCopy();
  }
  ConvCtx c{};
  c.rayIrr = h->dRayIrr; c.dirs = h->dDirs; c.layerMap = h->dLayerMap; c.atlas = h->dAtlas; c.texelRecords = h->dTexelRecords;
  // peers: whole layer blocks are pushed after the convolution (pushLayers); DDGI_EXCHANGE_TEXEL=1 keeps the first
  // version, per-texel peer stores from inside the convolution, for A/B measurements
  const bool texelStores = texelStoreExchange();
  for (int i = 0; i < h->numPeers; ++i) c.peers[i] = h->peers[i];
  c.numPeers = texelStores ? h->numPeers : 0;
  c.nXZ = h->nXZ; c.nY = h->nY; c.N = h->N; c.Sx = h->Sx; c.Sy = h->Sy; c.W = h->W; c.H = h->H; c.numLayers = numLayers;
  c.Pl = h->Pl; c.rowOffset = h->rowOffset; c.rowStride = h->rowStride;
  c.hysteresis = h->p.hysteresis; c.bakeMode = bakeMode(h) ? 1 : 0;
#ifndef DDGI_HOST_EMU
  if (!(h->p.flags & ANEREND_DDGI_FLAG_SIMPLE_KERNELS)) {
    ConvFastCtx f{}; f.c = c; f.weights = h->dConvW;
    BE_CHECK(h, be.ok((h->p.flags & ANEREND_DDGI_FLAG_FAST_MATH) ? launch_convolve_fast(f, be.stream) : launch_convolve_exact(f, be.stream)));
    be.launches += 2;
  } else
#endif
  DDGI_LAUNCH(be, k_probe_convolve, conv_texel_body, (uint32_t)numLayers * (uint32_t)h->Pl * 64u, c, 128);
  BE_CHECK(h, be.check());
  if (!texelStores && push) return pushLayers(h, numLayers);
  return 0;
}
// The exchange of a sharded update.  The convolution only writes this handle's OWN rows, which no peer reads or writes, so it runs as
// soon as the own trace is done; the first barrier (every rank has finished tracing against the previous atlas) is only needed in
// front of the push into the peers' atlases — the ranks that finish tracing first convolve while they wait for the slowest one.
// The second barrier: every rank's rows have landed here before the next trace samples them.
int convolveAndExchange(anerend_ddgi* h, const int* layers, int numLayers) {
  int rc;
  if (texelStoreExchange()) {             // per-texel peer stores from inside the convolution: the barrier has to come first
    if ((rc = peerBarrier(h))) return rc;
    if ((rc = convolveLayers(h, layers, numLayers))) return rc;
    return peerBarrier(h);
  }
  if ((rc = convolveLayers(h, layers, numLayers, false))) return rc;
  if ((rc = peerBarrier(h))) return rc;
#ifndef DDGI_HOST_EMU
  if (h->xworld > 1 && h->xStream) {
    // The push and the second barrier go to the exchange stream: of the next update only the hit shading (k_shade) samples the
    // atlas, so its ray generation, both traversals and the shadow-ray set-up run while the rows travel (fenceExchange in
    // front of the shading; C5 on 8 GPUs: 92 MB per rank over NVLink, 0.15 ms that used to sit at the end of every update).
    Backend& be = h->be;
    BE_CHECK(h, be.ok(cudaEventRecord(h->evConv, be.stream)));
    BE_CHECK(h, be.ok(cudaStreamWaitEvent(h->xStream, h->evConv, 0)));
    if ((rc = pushLayers(h, numLayers, h->xStream))) return rc;
    if ((rc = peerBarrier(h, h->xStream))) return rc;
    BE_CHECK(h, be.ok(cudaEventRecord(h->evXDone, h->xStream)));
    h->exchangePending = true;
    return 0;
  }
#endif
  if ((rc = pushLayers(h, numLayers))) return rc;
  return peerBarrier(h);
}

int translate(anerend_ddgi* h) {
  Backend& be = h->be;
  if (!h->haveUbo) return fail(h, ANEREND_DDGI_E_STATE, "translate before set_scene_ubo");
  if ((h->ubo.flags & ANEREND_UBO_DDGI_FLAG) != ANEREND_UBO_DDGI_FLAG) return 0;            // :57
  if ((h->ubo.flags & ANEREND_UBO_BAKE_MODE_ON_FLAG) == ANEREND_UBO_BAKE_MODE_ON_FLAG) return 0;   // :58
  int ix = (int)(h->ubo.camera_pos[0] / h->p.probe_step[0]), iz = (int)(h->ubo.camera_pos[2] / h->p.probe_step[2]);   // :64
  int dx = ix - h->lastProbeIdx[0], dz = iz - h->lastProbeIdx[1];
  if (dx == 0 && dz == 0) return 0;
  h->lastProbeIdx[0] = ix; h->lastProbeIdx[1] = iz;
  BE_CHECK(h, be.fenceDownload());
  { int rf = fenceExchange(h); if (rf) return rf; }
  BE_CHECK(h, be.d2d(h->dTemp, h->dAtlas, (size_t)h->W * h->H * sizeof(uint2)));   // first vkCmdCopyImage (:92-99)
  ++be.launches;
  TransCtx c{};
  c.temp = h->dTemp; c.atlas = h->dAtlas; c.W = h->W; c.nXZ = h->nXZ; c.nY = h->nY; c.dx = dx; c.dz = dz;
  c.ew = dx != 0 ? h->W - 10 : h->W; c.eh = dz != 0 ? h->W - 10 : h->W;
  DDGI_LAUNCH(be, k_probe_translate, translate_body, (uint32_t)c.ew * (uint32_t)c.eh * (uint32_t)h->nY, c, 256);
  BE_CHECK(h, be.check());
  return 0;
}

void destroyHandle(anerend_ddgi* h) {
  if (!h) return;
  Backend& be = h->be;
  be.waitDownload();
  be.sync();
  exchangeDisconnect(h);
#ifndef DDGI_HOST_EMU
  if (h->xStream) { cudaStreamDestroy(h->xStream); cudaEventDestroy(h->evConv); cudaEventDestroy(h->evXDone); h->xStream = nullptr; }
  if (h->pStream) {
    cudaStreamSynchronize(h->pStream); cudaStreamDestroy(h->pStream); h->pStream = nullptr;
    for (int k = 0; k < 2; ++k) { cudaEventDestroy(h->evTrace[k]); cudaEventDestroy(h->evHitsFree[k]); be.free(h->dDirsP[k]); be.free(h->dLayerMapP[k]); }
    cudaEventDestroy(h->evState); be.free(h->dCursorP);
  }
  be.free(h->dHits2); be.free(h->dHitInst2);
#endif
  be.free(h->dXflags); be.free(h->dXerr); be.free(h->dSelected); be.free(h->dVisChanged); be.free(h->dBigItems); be.free(h->dBigCount);
  be.free(h->dMovedTris); be.free(h->dMovedCount);
  freeAccel(h); freeScene(h);
  be.free(h->dAtlas); be.free(h->dTemp); be.free(h->dRayIrr); be.free(h->dDirs); be.free(h->dFib); be.freeHostPinned(h->hDirs); be.free(h->dRayOrder);
  be.free(h->dLayerMap); be.freeHostPinned(h->hLayerMap);
  be.free(h->dTextures); be.free(h->dTexPool); be.free(h->dSrgbLut); be.free(h->dEmisTex); be.free(h->dTileInfos);
  be.free(h->dExtRays); be.free(h->dExtDist);
  be.free(h->dSrtPos); be.free(h->dSrtLit); be.free(h->dSrtQueue); be.free(h->dSrtQcount);
  be.free(h->dRecords); be.free(h->dTexelRecords); be.free(h->dStats); be.free(h->dSceneBounds); be.free(h->dCounters);
  be.free(h->dHits); be.free(h->dSurf); be.free(h->dLit); be.free(h->dQueues); be.free(h->dQcount); be.free(h->dHitq); be.free(h->dHitqSorted); be.free(h->dConvW); be.free(h->dHitInst);
  be.destroy();
  delete h;
}

// The per-layer entry points (the frame-graph adapter's) carry no cross-GPU barriers: with peers connected a push would land in an
// atlas another rank is still tracing against.  Sharded updates go through _update / _update_layers, which bracket the exchange.
int refuseWithPeers(anerend_ddgi* h, const char* what) {
  if (h->numPeers > 0 || h->xworld > 1) return fail(h, ANEREND_DDGI_E_STATE, std::string(what) + " while a peer exchange is connected: use _update or _update_layers");
  return 0;
}
void beginCall(anerend_ddgi* h) { h->be.launches = 0; h->timed[0] = h->timed[1] = h->timed[2] = false; h->timedKernels = false; h->be.mark(0); }

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int DDGI_API(create)(const anerend_ddgi_params_t* params, anerend_ddgi_t** out) {
  if (!params || !out || params->struct_size != sizeof(anerend_ddgi_params_t)) return ANEREND_DDGI_E_INVALID;
  const anerend_ddgi_params_t& q = *params;
  if (q.num_probes_xz < 1 || q.num_probes_y < 1 || q.rays_per_probe < 32 || q.rays_per_probe > 1024 || (q.rays_per_probe % 32)) return ANEREND_DDGI_E_INVALID;
  int sx, sy; if (rayTile(q.rays_per_probe, sx, sy)) return ANEREND_DDGI_E_INVALID;
  if (q.slab_begin < 0 || q.slab_end > q.num_probes_y || q.slab_begin >= q.slab_end) return ANEREND_DDGI_E_INVALID;
  if (!(q.probe_step[0] > 0.0f && q.probe_step[1] > 0.0f && q.probe_step[2] > 0.0f)) return ANEREND_DDGI_E_INVALID;
  // ray indices, atlas addresses and queue offsets are 32-bit on the device
  if ((uint64_t)q.num_probes_xz * q.num_probes_xz * q.num_probes_y * q.rays_per_probe >= (1ull << 31) ||
      (uint64_t)q.num_probes_xz * q.num_probes_xz * q.num_probes_y * 100ull >= (1ull << 31)) return ANEREND_DDGI_E_INVALID;
  anerend_ddgi* h = new anerend_ddgi();
  h->p = q; h->nXZ = q.num_probes_xz; h->nY = q.num_probes_y; h->N = q.rays_per_probe; h->Sx = sx; h->Sy = sy;
  h->W = 10 * h->nXZ; h->H = 10 * h->nXZ * h->nY; h->P = h->nXZ * h->nXZ; h->Pl = h->P;
  Backend& be = h->be;
  if (!be.init(q.device)) { delete h; return ANEREND_DDGI_E_CUDA; }   // no CUDA device: there is no CPU fallback

  const size_t texels = (size_t)h->W * h->H, rays = (size_t)h->nY * h->P * h->N;
  bool ok = be.alloc(&h->dAtlas, texels) && be.alloc(&h->dTemp, texels) && be.alloc(&h->dRayIrr, rays) &&
            be.alloc(&h->dDirs, (size_t)h->nY * h->N) && be.allocHostPinned(&h->hDirs, (size_t)h->nY * h->N) &&
            be.alloc(&h->dSceneBounds, 6) && be.alloc(&h->dCounters, CNT_SIZE) && be.alloc(&h->dStats, 16) &&
            be.alloc(&h->dConvW, (size_t)h->nY * h->N * 64) && be.alloc(&h->dLayerMap, (size_t)h->nY) && be.allocHostPinned(&h->hLayerMap, (size_t)h->nY);
  if (ok && (q.flags & ANEREND_DDGI_FLAG_RAY_RECORDS)) ok = be.alloc(&h->dRecords, rays);
  if (ok && (q.flags & ANEREND_DDGI_FLAG_TEXEL_RECORDS)) ok = be.alloc(&h->dTexelRecords, texels);
  if (ok) ok = be.zero(h->dStats, 16 * sizeof(unsigned long long)) && be.zero(h->dRayIrr, rays * sizeof(uint2));
  if (!ok) { destroyHandle(h); return ANEREND_DDGI_E_CUDA; }
  memset(h->hDirs, 0, (size_t)h->nY * h->N * sizeof(float4));      // rows of layers never traced travel with their neighbours
  h->fib.resize((size_t)h->N * 3);
  for (int i = 0; i < h->N; ++i) fibonacciDir((uint32_t)i, (uint32_t)h->N, &h->fib[3 * i]);
  h->layerTraced.assign(h->nY, 0);
  if (!be.alloc(&h->dFib, (size_t)h->N * 3) || !be.h2d(h->dFib, h->fib.data(), (size_t)h->N * 3 * sizeof(float)) || !be.sync()) { destroyHandle(h); return ANEREND_DDGI_E_CUDA; }
  { // ray binning: slot -> ray index, sorted by the Morton code of the octahedral mapping of the (unrotated) direction;
    // a rotation is rigid, so 32 consecutive slots (one warp) stay a compact cone of directions
    std::vector<std::pair<uint32_t, uint16_t>> keyed(h->N);
    for (int i = 0; i < h->N; ++i) {
      float x = h->fib[3 * i], y = h->fib[3 * i + 1], z = h->fib[3 * i + 2];
      float l1 = fabsf(x) + fabsf(y) + fabsf(z); float u = x / l1, v = y / l1;
      if (z < 0.0f) { float nu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f), nv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f); u = nu; v = nv; }
      // hemisphere first (keeps the two z-halves apart), then 2-D Morton inside the square
      uint32_t qu = (uint32_t)fminf(fmaxf((u * 0.5f + 0.5f) * 1024.0f, 0.0f), 1023.0f), qv = (uint32_t)fminf(fmaxf((v * 0.5f + 0.5f) * 1024.0f, 0.0f), 1023.0f);
      uint32_t m = 0; for (int b = 0; b < 10; ++b) m |= ((qu >> b) & 1u) << (2 * b) | ((qv >> b) & 1u) << (2 * b + 1);
      keyed[i] = std::make_pair((z < 0.0f ? (1u << 20) : 0u) | m, (uint16_t)i);
    }
    std::sort(keyed.begin(), keyed.end());
    std::vector<uint16_t> order(h->N); for (int i = 0; i < h->N; ++i) order[i] = keyed[i].second;
    if (!be.alloc(&h->dRayOrder, h->N) || !be.h2d(h->dRayOrder, order.data(), h->N * sizeof(uint16_t)) || !be.sync()) { destroyHandle(h); return ANEREND_DDGI_E_CUDA; }
  }
  if (clearAtlas(h) || !be.sync()) { destroyHandle(h); return ANEREND_DDGI_E_CUDA; }
  *out = h;
  return 0;
}

int DDGI_API(destroy)(anerend_ddgi_t* h) {
  destroyHandle(h);
  return 0;
}

int DDGI_API(reset)(anerend_ddgi_t* h) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  h->currentLayer = 0; h->lastProbeIdx[0] = h->lastProbeIdx[1] = 0;
  std::fill(h->layerTraced.begin(), h->layerTraced.end(), 0);
  return clearAtlas(h);
}
const char* DDGI_API(last_error)(const anerend_ddgi_t* h) { return h ? h->err.c_str() : "null handle"; }
int DDGI_API(sync)(anerend_ddgi_t* h) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  BE_CHECK(h, h->be.sync());
#ifndef DDGI_HOST_EMU
  if (h->xStream) BE_CHECK(h, h->be.ok(cudaStreamSynchronize(h->xStream)));      // the exchange of the last update
  if (h->pStream) BE_CHECK(h, h->be.ok(cudaStreamSynchronize(h->pStream)));
#endif
  return peerCheck(h);
}

static size_t texelBytes(uint32_t format) { return format == ANEREND_TEX_RGBA16F ? 8 : format <= ANEREND_TEX_RGBA8_SRGB ? 4 : (format == ANEREND_TEX_RG8_UNORM ? 2 : 1); }
// tileInfoBuffer (bindless.glsl binding 14, GPUTileInfo): one entry per tile of the (2*10+1)^2 page around the camera; ddgiAtlasTex
// = bindless index of the tile's baked atlas (an ANEREND_TEX_RGBA16F texture of the atlas size) or -1.  count = 0 removes it.
int DDGI_API(upload_tile_infos)(anerend_ddgi_t* h, const anerend_tile_info_t* tiles, uint32_t count) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  Backend& be = h->be;
  if (count && !tiles) return fail(h, ANEREND_DDGI_E_INVALID, "null tile array");
  BE_CHECK(h, be.sync());
  be.free(h->dTileInfos); h->dTileInfos = nullptr; h->numTiles = 0;
  if (count == 0) return 0;
  BE_CHECK(h, be.alloc(&h->dTileInfos, count)); BE_CHECK(h, be.h2d(h->dTileInfos, tiles, count * sizeof(int32_t))); BE_CHECK(h, be.sync());
  h->numTiles = count;
  return 0;
}
int DDGI_API(upload_textures)(anerend_ddgi_t* h, const anerend_texture_desc_t* tex, uint32_t count) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  Backend& be = h->be;
  if (count && !tex) return fail(h, ANEREND_DDGI_E_INVALID, "null texture array");
  std::vector<DTexture> table(count);
  size_t total = 0;
  for (uint32_t i = 0; i < count; ++i) {
    if (!tex[i].pixels || tex[i].width == 0 || tex[i].height == 0 || tex[i].width > 16384u || tex[i].height > 16384u || tex[i].format > ANEREND_TEX_RGBA16F)
      return fail(h, ANEREND_DDGI_E_INVALID, "bad texture descriptor (BC-compressed assets are handed over decoded)");
    const size_t bpp = texelBytes(tex[i].format);
    total = (total + 15) & ~(size_t)15;      // 16-byte aligned images (RGBA16F texels are read as 8-byte words)
    if (total > 0xFFFFFFFFull - (size_t)tex[i].width * tex[i].height * bpp) return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "texture pool larger than 4 GiB");
    table[i].offset = (uint32_t)total; table[i].width = tex[i].width; table[i].height = tex[i].height;
    table[i].formatClamp = tex[i].format | (tex[i].clamp_to_edge ? 1u << 8 : 0u);
    total += (size_t)tex[i].width * tex[i].height * bpp;
  }
  BE_CHECK(h, be.sync());
  be.free(h->dTextures); be.free(h->dTexPool); h->dTextures = nullptr; h->dTexPool = nullptr; h->numTextures = 0; h->textures.clear();
  h->texturesStale = h->haveScene && h->texturedScene;
  if (count == 0) return 0;
  if (!h->dSrgbLut) {          // texel decode table, made once on the host in fp32 (the oracle makes the same one)
    float lut[512];
    for (int c = 0; c < 256; ++c) {
      const float v = (float)c / 255.0f;
      const float higher = powf((v + 0.055f) / 1.055f, 2.4f), lower = v / 12.92f;     // sRGB EOTF, as toLinear (bindless.glsl:240-247)
      lut[c] = (v < 0.04045f) ? lower : higher; lut[256 + c] = v;
    }
    BE_CHECK(h, be.alloc(&h->dSrgbLut, 512)); BE_CHECK(h, be.h2d(h->dSrgbLut, lut, sizeof(lut))); BE_CHECK(h, be.sync());
  }
  BE_CHECK(h, be.alloc(&h->dTextures, count)); BE_CHECK(h, be.alloc(&h->dTexPool, total));
  for (uint32_t i = 0; i < count; ++i) {
    const size_t bpp = texelBytes(tex[i].format);
    BE_CHECK(h, be.h2d(h->dTexPool + table[i].offset, tex[i].pixels, (size_t)tex[i].width * tex[i].height * bpp));
  }
  BE_CHECK(h, be.h2d(h->dTextures, table.data(), count * sizeof(DTexture))); BE_CHECK(h, be.sync());
  h->numTextures = count; h->textures = table;
  h->texturesStale = h->haveScene && h->texturedScene;       // material indices were checked against the previous array
  return 0;
}

int DDGI_API(upload_scene)(anerend_ddgi_t* h,
    const anerend_vertex_t* vertices, uint32_t nv, const uint32_t* indices, uint32_t ni,
    const anerend_mesh_info_t* meshes, const uint32_t* meshNumIndices, uint32_t nm,
    const anerend_material_info_t* materials, uint32_t nmat,
    const anerend_renderable_t* renderables, uint32_t nr,
    const uint32_t* modelIdx, uint32_t nmi, const uint32_t* matIdx, uint32_t nmti,
    const anerend_light_t* lights, uint32_t nl) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  Backend& be = h->be;
  bool textured = false, emisTex = false;
  for (uint32_t i = 0; i < nmat; ++i) for (int k = 0; k < 4; ++k) {
    const int32_t ti = materials[i].bindless_indices[k];
    if (k == 2 || ti == -1) continue;          // [2] = normal map: never sampled on this path (rchit:77 passes a zero tangent)
    if (ti < 0 || (uint32_t)ti >= h->numTextures) return fail(h, ANEREND_DDGI_E_INVALID, "material refers to a texture that was not uploaded (upload_textures first)");
    textured = true; emisTex = emisTex || k == 3;
  }
  if ((h->p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) && nl > 31) return fail(h, ANEREND_DDGI_E_INVALID, "at most 31 point lights with FLAG_POINT_LIGHTS");
  for (uint32_t i = 0; i < nm; ++i) {
    if ((uint64_t)meshes[i].index_offset + meshNumIndices[i] > ni) return fail(h, ANEREND_DDGI_E_INVALID, "mesh index range exceeds the index buffer");
  }
  std::vector<uint32_t> meshVerts;
  { // every index must address a vertex inside the giga vertex buffer
    meshVerts.assign(nm, 0u);
    for (uint32_t m = 0; m < nm; ++m) { const uint32_t vo = meshes[m].vertex_offset;
      for (uint32_t k = 0; k < meshNumIndices[m]; ++k) { const uint32_t ix = indices[meshes[m].index_offset + k];
        if ((uint64_t)vo + ix >= nv) return fail(h, ANEREND_DDGI_E_INVALID, "vertex index out of range");
        meshVerts[m] = std::max(meshVerts[m], ix + 1u); } } }
  BE_CHECK(h, be.sync());
  freeAccel(h); freeScene(h);
  // everything validated: only now does the new scene (and what was learnt about its textures) replace the old one
  h->texturedScene = textured; h->emissiveTextures = emisTex; h->texturesStale = false;
  BE_CHECK(h, be.alloc(&h->dVertices, nv)); BE_CHECK(h, be.alloc(&h->dIndices, ni)); BE_CHECK(h, be.alloc(&h->dMaterials, nmat));
  BE_CHECK(h, be.alloc(&h->dRenderables, nr)); BE_CHECK(h, be.alloc(&h->dLights, nl));
  BE_CHECK(h, be.h2d(h->dVertices, vertices, (size_t)nv * sizeof(*vertices))); BE_CHECK(h, be.h2d(h->dIndices, indices, (size_t)ni * 4));
  BE_CHECK(h, be.h2d(h->dMaterials, materials, (size_t)nmat * sizeof(*materials))); BE_CHECK(h, be.h2d(h->dRenderables, renderables, (size_t)nr * sizeof(*renderables)));
  BE_CHECK(h, be.h2d(h->dLights, lights, (size_t)nl * sizeof(*lights)));
  BE_CHECK(h, be.sync());
  h->renderables.assign(renderables, renderables + nr); h->meshes.assign(meshes, meshes + nm);
  h->meshNumIndices.assign(meshNumIndices, meshNumIndices + nm); h->modelIdx.assign(modelIdx, modelIdx + nmi); h->matIdx.assign(matIdx, matIdx + nmti);
  h->numVertices = nv; h->numIndices = ni; h->numMaterials = nmat; h->numLights = nl;
  h->meshNumVertices = meshVerts; h->dirtyVerts.clear(); h->verticesDirty = false;
  h->haveScene = true;
  return 0;
}

int DDGI_API(set_scene_ubo)(anerend_ddgi_t* h, const anerend_scene_ubo_t* ubo) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!ubo) return fail(h, ANEREND_DDGI_E_INVALID, "null ubo");
  h->ubo = *ubo; h->haveUbo = true; return 0;
}

int DDGI_API(build_accel)(anerend_ddgi_t* h) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveScene) return fail(h, ANEREND_DDGI_E_STATE, "build_accel before upload_scene");
  return buildAccel(h);
}

static int refitCommon(anerend_ddgi_t* h) {
  // visibility must be unchanged: the topology of the acceleration structure is kept
  std::vector<DInstance> inst; uint32_t n = 0;
  int rc = selectInstances(h, inst, n); if (rc) return rc;
  bool same = n == h->numTris && inst.size() == h->instances.size();
  for (size_t i = 0; same && i < inst.size(); ++i) same = inst[i].renderable == h->instances[i].renderable && inst[i].mesh == h->instances[i].mesh && inst[i].material == h->instances[i].material;
  if (!same) { h->haveAccel = false; return fail(h, ANEREND_DDGI_E_STATE, "visibility changed: rebuild required"); }
  return refitAccel(h);
}
// Deforming geometry (skinned meshes: the reference re-skins into the giga vertex buffer every frame and refits the BLAS,
// UpdateBlasPass.cpp:60-175, VulkanRenderer.cpp:967-1143): refresh a vertex range, then _refit_accel — the topology of the
// hierarchy is kept, triangles, boxes and the per-triangle shading records are recomputed from the new vertices.
int DDGI_API(update_vertices)(anerend_ddgi_t* h, const anerend_vertex_t* vertices, uint32_t first, uint32_t count) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveScene) return fail(h, ANEREND_DDGI_E_STATE, "update_vertices before upload_scene");
  if (!vertices || (uint64_t)first + count > h->numVertices) return fail(h, ANEREND_DDGI_E_INVALID, "vertex range out of bounds");
  BE_CHECK(h, h->be.sync());
  BE_CHECK(h, h->be.h2d(h->dVertices + first, vertices, (size_t)count * sizeof(anerend_vertex_t)));
  h->verticesDirty = true; h->dirtyVerts.emplace_back(first, count);
  return 0;
}
int DDGI_API(update_vertices_device)(anerend_ddgi_t* h, const void* dVertices, uint32_t first, uint32_t count) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveScene) return fail(h, ANEREND_DDGI_E_STATE, "update_vertices before upload_scene");
  if (!dVertices || (uint64_t)first + count > h->numVertices) return fail(h, ANEREND_DDGI_E_INVALID, "vertex range out of bounds");
  BE_CHECK(h, h->be.d2d(h->dVertices + first, dVertices, (size_t)count * sizeof(anerend_vertex_t)));
  h->verticesDirty = true; h->dirtyVerts.emplace_back(first, count);
  return 0;
}
int DDGI_API(refit_accel)(anerend_ddgi_t* h, const anerend_renderable_t* renderables, uint32_t nr) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveAccel) return fail(h, ANEREND_DDGI_E_STATE, "refit before build_accel");
  if (renderables) {
    if (nr != h->renderables.size()) return fail(h, ANEREND_DDGI_E_INVALID, "renderable count changed");
    h->renderables.assign(renderables, renderables + nr);
    BE_CHECK(h, h->be.sync());
    BE_CHECK(h, h->be.h2d(h->dRenderables, h->renderables.data(), (size_t)nr * sizeof(anerend_renderable_t)));
  }
  return refitCommon(h);
}
int DDGI_API(refit_accel_device)(anerend_ddgi_t* h, const void* dRenderables, uint32_t nr) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveAccel) return fail(h, ANEREND_DDGI_E_STATE, "refit before build_accel");
  if (!dRenderables || nr != h->renderables.size()) return fail(h, ANEREND_DDGI_E_INVALID, "renderable count changed");
  // device-written RenderableBuffer (cull.comp:296-303 sets the cull bit there every frame, particle_update.comp:84-85 the
  // transforms): the visible / cull bits are compared with the selection the structure was built from (4 bytes come back); where a
  // renderable entered or left, the host copy is refreshed and the structure is rebuilt — what UpdateTLASPass.cpp:198-228 does every
  // frame — otherwise the topology is kept and refitted.
  Backend& be = h->be;
  BE_CHECK(h, be.d2d(h->dRenderables, dRenderables, (size_t)nr * sizeof(anerend_renderable_t)));
  uint32_t changed = 0;
  BE_CHECK(h, be.zero(h->dVisChanged, sizeof(uint32_t)));
  VisCtx vc{h->dRenderables, h->dSelected, h->dVisChanged};
  DDGI_LAUNCH(be, k_vis_changed, vis_changed_body, nr, vc, 128);
  BE_CHECK(h, be.check()); BE_CHECK(h, be.d2h(&changed, h->dVisChanged, sizeof(uint32_t)));
  if (changed) {
    BE_CHECK(h, be.d2h(h->renderables.data(), h->dRenderables, (size_t)nr * sizeof(anerend_renderable_t)));
    return buildAccel(h);
  }
  // transforms of the host copy are not refreshed here: instance selection reads visibility, meshes and materials only
  return refitAccel(h);
}

// cull.comp:241-244,296-303: the renderable culling pass marks every VISIBLE renderable whose bounding-sphere centre — bounds.xyz moved
// by the translation of its transform only (buildTranslation, cull.comp:229-239) — lies farther than `range` (reference: 50 m) from the
// camera; tlas_update.comp then leaves it out of the TLAS.  For hosts that do not run the Vulkan cull pass in front of the chain.
// A changed bit invalidates the acceleration structure (the instance set differs): _build_accel must follow.
int DDGI_API(apply_camera_cull)(anerend_ddgi_t* h, float range, uint32_t* changedOut) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveScene) return fail(h, ANEREND_DDGI_E_STATE, "apply_camera_cull before upload_scene");
  if (!h->haveUbo) return fail(h, ANEREND_DDGI_E_STATE, "apply_camera_cull before set_scene_ubo");
  uint32_t changed = 0;
  for (anerend_renderable_t& r : h->renderables) {
    if (!(r.visible & 1u)) continue;                                   // cull.comp:289-292: hidden renderables are not looked at
    // trans * vec4(bounds.xyz, 1) with trans = identity + translation column: c0*x + c1*y + c2*z + c3*1
    const float cx = ((1.0f * r.bounds[0] + 0.0f * r.bounds[1]) + 0.0f * r.bounds[2]) + r.transform[12] * 1.0f;
    const float cy = ((0.0f * r.bounds[0] + 1.0f * r.bounds[1]) + 0.0f * r.bounds[2]) + r.transform[13] * 1.0f;
    const float cz = ((0.0f * r.bounds[0] + 0.0f * r.bounds[1]) + 1.0f * r.bounds[2]) + r.transform[14] * 1.0f;
    const float dx = h->ubo.camera_pos[0] - cx, dy = h->ubo.camera_pos[1] - cy, dz = h->ubo.camera_pos[2] - cz;
    const bool within = sqrtf((dx * dx + dy * dy) + dz * dz) <= range;                        // distance(cameraPos, centre) <= range
    const uint32_t v = (r.visible & ~2u) | (within ? 0u : 2u);                                // setCullBit(rend, !withinRange), bindless.glsl:254-258
    if (v != r.visible) { r.visible = v; ++changed; }
  }
  if (changed) {
    BE_CHECK(h, h->be.sync());
    BE_CHECK(h, h->be.h2d(h->dRenderables, h->renderables.data(), h->renderables.size() * sizeof(anerend_renderable_t)));
    BE_CHECK(h, h->be.sync());
    h->haveAccel = false;                                              // trace / refit now fail loudly until _build_accel
  }
  if (changedOut) *changedOut = changed;
  return 0;
}
int DDGI_API(get_bvh_info)(anerend_ddgi_t* h, anerend_ddgi_bvh_info_t* o) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  memset(o, 0, sizeof(*o));
  o->num_triangles = h->numTris; o->num_nodes = h->numNodes + (twoLevel(h) ? h->blasNodes : 0u); o->node_bytes = 80; o->tri_bytes = 48; o->num_instances = (uint32_t)h->instances.size();
  for (int k = 0; k < 3; ++k) { o->bounds_min[k] = h->boundsMin[k]; o->bounds_max[k] = h->boundsMax[k]; }
  return 0;
}
int DDGI_API(get_two_level_info)(anerend_ddgi_t* h, anerend_ddgi_two_level_info_t* o) {
  if (!h || !o) return ANEREND_DDGI_E_INVALID;
  memset(o, 0, sizeof(*o));
  if (!twoLevel(h) || !h->twoBuilt) return 0;
  o->num_meshes = (uint32_t)h->blas.size(); o->stored_triangles = h->blasTris; o->blas_nodes = h->blasNodes; o->tlas_nodes = h->numNodes;
  o->tlas_depth = h->tlasDepth; o->blas_depth_max = h->blasDepthMax; o->tlas_build_ms = h->tlasMs;
  const uint64_t ni = h->instances.size();
  o->accel_bytes = 80ull * (h->blasNodes + h->numNodes) + 48ull * (h->blasTris + h->tlasList.size()) + 80ull * h->blasTris + (sizeof(InstXform) + sizeof(DInstance)) * ni;
  return 0;
}
int DDGI_API(download_instances)(anerend_ddgi_t* h, anerend_ddgi_instance_t* out, size_t count) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveAccel) return fail(h, ANEREND_DDGI_E_STATE, "no acceleration structure");
  if (!out || count != h->instances.size()) return fail(h, ANEREND_DDGI_E_INVALID, "instance count mismatch");
  for (size_t i = 0; i < count; ++i) { const DInstance& in = h->instances[i];
    out[i].renderable = in.renderable; out[i].custom_index = in.material | (in.mesh << 10); out[i].first_triangle = in.firstTri; out[i].num_triangles = in.numTris; }
  return 0;
}
int DDGI_API(download_bvh)(anerend_ddgi_t* h, void* nodes, size_t nb, void* tris, size_t tb) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->haveAccel) return fail(h, ANEREND_DDGI_E_STATE, "no acceleration structure");
  if (twoLevel(h)) return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "download_bvh reads a one-level structure");
  if (nb != (size_t)h->numNodes * 80 || tb != (size_t)h->numTris * 48) return fail(h, ANEREND_DDGI_E_INVALID, "size mismatch");
  BE_CHECK(h, h->be.d2h(nodes, h->dNodes, nb)); BE_CHECK(h, h->be.d2h(tris, h->dBvhTris, tb));
  return 0;
}

int DDGI_API(translate)(anerend_ddgi_t* h) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  beginCall(h);
  int rc = translate(h); h->be.mark(1); h->timed[0] = true;
  return rc;
}
int DDGI_API(trace_layer)(anerend_ddgi_t* h, int32_t layer, const float rot[16]) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (h->haveUbo && !ddgiOn(h)) return 0;
  if (layer < 0 || layer >= h->nY || !rot) return fail(h, ANEREND_DDGI_E_INVALID, "layer out of range");
  { int rp = refuseWithPeers(h, "trace_layer"); if (rp) return rp; }
  beginCall(h); h->be.mark(1);
  const int one = layer;
  int rc = traceLayers(h, &one, 1, rot, false); h->be.mark(2); h->timed[1] = true;
  return rc;
}
int DDGI_API(convolve_layer)(anerend_ddgi_t* h, int32_t layer) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (h->haveUbo && !ddgiOn(h)) return 0;
  if (layer < 0 || layer >= h->nY) return fail(h, ANEREND_DDGI_E_INVALID, "layer out of range");
  { int rp = refuseWithPeers(h, "convolve_layer"); if (rp) return rp; }
  beginCall(h); h->be.mark(2);
  const int one = layer;
  int rc = convolveLayers(h, &one, 1); h->be.mark(3); h->timed[2] = true;
  return rc;
}
int DDGI_API(update)(anerend_ddgi_t* h, int32_t layer, const float* rots) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (h->haveUbo && !ddgiOn(h)) return 0;
  if (!rots || layer >= h->nY) return fail(h, ANEREND_DDGI_E_INVALID, "bad update arguments");
  beginCall(h); h->be.mark(1);
  std::vector<int> layers;
  if (layer >= 0) layers.push_back(layer); else for (int l = h->p.slab_begin; l < h->p.slab_end; ++l) layers.push_back(l);
  int rc = traceLayers(h, layers.data(), (int)layers.size(), rots, layer < 0);
  h->be.mark(2); h->timed[1] = true;
  if (rc) return rc;
  rc = convolveAndExchange(h, layers.data(), (int)layers.size());
  h->be.mark(3); h->timed[2] = true;
  return rc;
}
int DDGI_API(update_layers)(anerend_ddgi_t* h, const int32_t* layerList, uint32_t n, const float* rots) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (h->haveUbo && !ddgiOn(h)) return 0;
  if (!rots || !layerList || n == 0 || n > (uint32_t)h->nY) return fail(h, ANEREND_DDGI_E_INVALID, "bad update_layers arguments");
  std::vector<uint8_t> seen(h->nY, 0);
  for (uint32_t i = 0; i < n; ++i) {
    if (layerList[i] < 0 || layerList[i] >= h->nY || seen[layerList[i]]) return fail(h, ANEREND_DDGI_E_INVALID, "layer out of range or listed twice");
    seen[layerList[i]] = 1;
  }
  beginCall(h); h->be.mark(1);
  int rc = traceLayers(h, layerList, (int)n, rots, true);
  h->be.mark(2); h->timed[1] = true;
  if (rc) return rc;
  rc = convolveAndExchange(h, layerList, (int)n);
  h->be.mark(3); h->timed[2] = true;
  return rc;
}
int DDGI_API(frame)(anerend_ddgi_t* h, const float rot[16], int32_t* layerOut) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  { int rp = refuseWithPeers(h, "frame"); if (rp) return rp; }
  beginCall(h);
  int rc = translate(h); h->be.mark(1); h->timed[0] = true;
  if (rc) return rc;
  if (!ddgiOn(h)) return 0;
  const int layer = h->currentLayer;                                   // IrradianceProbeRayTracingPass.cpp:112
  rc = traceLayers(h, &layer, 1, rot, false); h->be.mark(2); h->timed[1] = true;
  if (rc) return rc;
  if (layerOut) *layerOut = layer;                                     // blackboard "ProbeLayer" (:158)
  h->currentLayer = (h->currentLayer + 1 >= h->nY) ? 0 : h->currentLayer + 1;   // :160-164
  rc = convolveLayers(h, &layer, 1); h->be.mark(3); h->timed[2] = true;
  return rc;
}

// ---- ShadowRT on the same BVH ------------------------------------------------------------------
namespace {
int shadowRt(anerend_ddgi* h, const float* dDepth, uint32_t width, uint32_t height, float* dOut) {
  Backend& be = h->be;
  if (!h->haveAccel) return fail(h, ANEREND_DDGI_E_STATE, "shadow_rt before build_accel");
  if (!h->haveUbo) return fail(h, ANEREND_DDGI_E_STATE, "shadow_rt before set_scene_ubo");
  if (!(h->ubo.flags & ANEREND_UBO_RT_SHADOWS_FLAG)) return 0;          // shadow.rgen:32-34: the pass writes nothing
  const uint32_t n = width * height;
  if (n == 0) return 0;
  ShadowRtCtx c{};
  c.nodes = h->dNodes; c.bvhTris = h->dBvhTris; c.numNodes = h->numNodes; c.depth = dDepth; c.out = dOut; c.width = (int)width; c.height = (int)height;
  c.xforms = h->dXforms; c.twoLevel = twoLevel(h) ? 1u : 0u;
  for (int k = 0; k < 16; ++k) c.invViewProj[k] = h->ubo.inv_view_proj[k];
  c.lightDir = mk3(h->ubo.light_dir[0], h->ubo.light_dir[1], h->ubo.light_dir[2]);
  c.tmin = 0.1f; c.tmax = 50.0f;                                        // literals of shadow.rgen:64-65
  c.screenW = (float)h->ubo.screen_width; c.screenH = (float)h->ubo.screen_height;
#ifndef DDGI_HOST_EMU
  if (!(h->p.flags & ANEREND_DDGI_FLAG_SIMPLE_KERNELS) && h->numNodes > 0) {
    // the persistent any-hit kernel over one queue of pixels; scratch of its own (a frame buffer is not a probe layer)
    if (h->srtPixels < n) {
      BE_CHECK(h, be.sync());
      be.free(h->dSrtPos); be.free(h->dSrtLit); be.free(h->dSrtQueue); h->dSrtPos = nullptr; h->dSrtLit = nullptr; h->dSrtQueue = nullptr; h->srtPixels = 0;
      BE_CHECK(h, be.alloc(&h->dSrtPos, n)); BE_CHECK(h, be.alloc(&h->dSrtLit, n)); BE_CHECK(h, be.alloc(&h->dSrtQueue, n));
      if (!h->dSrtQcount) BE_CHECK(h, be.alloc(&h->dSrtQcount, FAST_QC_SIZE));
      h->srtPixels = n;
    }
    { int rcg = ensureFastGrids(h); if (rcg) return rcg; }
    h->srtGrid = h->gridShadow[0];
    k_zero_words<<<(FAST_QC_SIZE + 255u) / 256u, 256, 0, be.stream>>>(h->dSrtQcount, (uint32_t)FAST_QC_SIZE);
    k_shadowrt_stage<<<(n + 255u) / 256u, 256, 0, be.stream>>>(c, h->dSrtPos, h->dSrtLit, h->dSrtQueue, h->dSrtQcount + FAST_QC(0));
    FastCtx f{};
    f.t.nodes = h->dNodes; f.t.bvhTris = h->dBvhTris; f.t.numNodes = h->numNodes; f.t.lightDir = c.lightDir; f.t.tmin = c.tmin; f.t.tmax = c.tmax;
    f.t.xforms = h->dXforms; f.t.twoLevel = c.twoLevel;
    f.surf = h->dSrtPos; f.lit = h->dSrtLit; f.queues = h->dSrtQueue; f.qcount = h->dSrtQcount; f.totalRays = n; f.numQueues = 1u;
    f.surfStride = 1u; f.rawSunDir = 1u;
    f.tune[0] = TUNE_NODE_STEPS; f.tune[1] = f.tune[2] = TUNE_REFILL_MIN; f.tune[3] = 0x64646464u; f.tune[4] = TUNE_STEP_MIN; f.tune[5] = TUNE_TRI_MIN;
    if (twoLevel(h)) k_trace<true, false, false, true><<<h->srtGrid, 128, 0, be.stream>>>(f);
    else k_trace<true, false><<<h->srtGrid, 128, 0, be.stream>>>(f);
    k_shadowrt_resolve<<<(n + 255u) / 256u, 256, 0, be.stream>>>(dDepth, h->dSrtLit, dOut, n);
    be.launches += 4;
    BE_CHECK(h, be.check());
    return 0;
  }
#endif
  DDGI_LAUNCH(be, k_shadow_rt, shadow_rt_body, n, c, 128);
  BE_CHECK(h, be.check());
  return 0;
}
}  // namespace
int DDGI_API(shadow_rt_device)(anerend_ddgi_t* h, const void* dDepth, uint32_t width, uint32_t height, void* dOut) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!dDepth || !dOut) return fail(h, ANEREND_DDGI_E_INVALID, "null argument");
  return shadowRt(h, (const float*)dDepth, width, height, (float*)dOut);
}
int DDGI_API(shadow_rt)(anerend_ddgi_t* h, const float* depth, uint32_t width, uint32_t height, float* out) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!depth || !out) return fail(h, ANEREND_DDGI_E_INVALID, "null argument");
  const size_t n = (size_t)width * height;
  if (n == 0 || !h->haveUbo || !(h->ubo.flags & ANEREND_UBO_RT_SHADOWS_FLAG)) return h->haveUbo ? 0 : fail(h, ANEREND_DDGI_E_STATE, "shadow_rt before set_scene_ubo");
  Backend& be = h->be;
  float *dD = nullptr, *dO = nullptr;
  if (!be.alloc(&dD, n) || !be.alloc(&dO, n)) { be.free(dD); be.free(dO); return fail(h, ANEREND_DDGI_E_CUDA, std::string("shadow_rt: ") + be.errStr()); }
  int rc = be.h2d(dD, depth, n * sizeof(float)) ? 0 : fail(h, ANEREND_DDGI_E_CUDA, std::string("shadow_rt: ") + be.errStr());
  if (!rc) rc = shadowRt(h, dD, width, height, dO);
  if (!rc && !be.d2h(out, dO, n * sizeof(float))) rc = fail(h, ANEREND_DDGI_E_CUDA, std::string("shadow_rt: ") + be.errStr());
  be.sync(); be.free(dD); be.free(dO);
  return rc;
}

// ---- SpecularGIRT on the same BVH and atlas -----------------------------------------------------
namespace {
int specularGi(anerend_ddgi* h, const float* dDepth, const float4* dGeom0, uint32_t width, uint32_t height, uint2* dOut) {
  if (!h->haveAccel) return fail(h, ANEREND_DDGI_E_STATE, "specular_gi before build_accel");
  { int rf = fenceExchange(h); if (rf) return rf; }
  if (!h->haveUbo) return fail(h, ANEREND_DDGI_E_STATE, "specular_gi before set_scene_ubo");
  if (!(h->ubo.flags & ANEREND_UBO_RT_ON_FLAG) || !(h->ubo.flags & ANEREND_UBO_SPECULAR_GI_FLAG)) return 0;    // SpecularGIRTPass.cpp:102-103
  { int rct = texturesUsable(h); if (rct) return rct; }
  const uint32_t n = width * height;
  if (n == 0) return 0;
  SpecularCtx c{};
  c.t = sceneTraceCtx(h);
  c.t.tmin = 0.1f; c.t.tmax = 50.0f; c.t.records = nullptr; c.t.stats = nullptr;      // literals of mirror_reflections.rgen:57-58 / .rchit:49-50
  c.depth = dDepth; c.geom0 = dGeom0; c.out = dOut; c.width = (int)width; c.height = (int)height;
  for (int k = 0; k < 16; ++k) c.invViewProj[k] = h->ubo.inv_view_proj[k];
  c.cameraPos = mk3(h->ubo.camera_pos[0], h->ubo.camera_pos[1], h->ubo.camera_pos[2]);
  c.screenW = (float)h->ubo.screen_width; c.screenH = (float)h->ubo.screen_height;
#ifndef DDGI_HOST_EMU
  static const bool specWavefront = getenv("DDGI_SPECULAR_SIMPLE") == nullptr;      // A/B switch: the one-thread-per-pixel kernel (5.9 ms vs 1.17 ms at 1080p)
  if (specWavefront && !(h->p.flags & ANEREND_DDGI_FLAG_SIMPLE_KERNELS) && h->numNodes > 0) {
    // the wavefront kernels in external-ray mode: one reflected ray per pixel instead of the probe rays
    Backend& be = h->be;
    const uint32_t nq = 1u + ((h->p.flags & ANEREND_DDGI_FLAG_POINT_LIGHTS) ? h->numLights : 0u);
    { int rcs = ensureFastScratch(h, n, nq); if (rcs) return rcs; }
    { int rcg = ensureFastGrids(h); if (rcg) return rcg; }
    h->pipeStateDirty = true;             // this pass uses hit-record set 0 on the main stream
    if (h->extPixels < n) {
      BE_CHECK(h, be.sync());
      be.free(h->dExtRays); be.free(h->dExtDist); h->dExtRays = nullptr; h->dExtDist = nullptr; h->extPixels = 0;
      BE_CHECK(h, be.alloc(&h->dExtRays, 2 * (size_t)n)); BE_CHECK(h, be.alloc(&h->dExtDist, n));
      h->extPixels = n;
    }
    FastCtx f{};
    f.t = c.t; f.t.rayIrr = dOut; f.t.layerMap = nullptr; f.t.rayOrder = nullptr;
    f.shadeTris = h->dShadeTris; f.shadeUV = h->dShadeUV; f.emisTex = h->dEmisTex; f.hits = h->dHits; f.surf = h->dSurf; f.lit = h->dLit;
    f.queues = h->dQueues; f.hitq = h->dHitq; f.qcount = h->dQcount; f.totalRays = n; f.numQueues = nq; f.hitInst = h->dHitInst;
    f.surfStride = 3u; f.rawSunDir = 0u; f.extRays = h->dExtRays; f.extDist = h->dExtDist;
    for (int k = 0; k < 8; ++k) f.tune[k] = h->tune[k];
    f.tune[3] = 0x64646464u;
    k_zero_words<<<(FAST_QC_SIZE + 255u) / 256u, 256, 0, be.stream>>>(h->dQcount, (uint32_t)FAST_QC_SIZE);
    k_specular_stage<<<(n + 255u) / 256u, 256, 0, be.stream>>>(c, h->dExtRays);
    if (twoLevel(h)) {
      k_trace<false, false, true, true><<<h->gridPrimary[0], 128, 0, be.stream>>>(f);
      k_shade_prep<256, true, 2048, false, true><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
      k_trace<true, false, false, true><<<h->gridShadow[0], 128, 0, be.stream>>>(f);
    } else {
      k_trace<false, false, true><<<h->gridPrimary[0], 128, 0, be.stream>>>(f);
      k_shade_prep<256, true><<<(n + 255u) / 256u, 256, 0, be.stream>>>(f);
      k_trace<true, false><<<h->gridShadow[0], 128, 0, be.stream>>>(f);
    }
    BE_CHECK(h, be.ok((h->p.flags & ANEREND_DDGI_FLAG_FAST_MATH) ? launch_shade_ext_fast(f, be.stream) : launch_shade_ext_exact(f, be.stream)));
    k_specular_resolve<<<(n + 255u) / 256u, 256, 0, be.stream>>>(dDepth, dOut, n);
    be.launches += 7;
    BE_CHECK(h, be.check());
    return 0;
  }
#endif
  DDGI_LAUNCH(h->be, k_specular_gi, specular_gi_body, n, c, 128);
  BE_CHECK(h, h->be.check());
  return 0;
}
}  // namespace
int DDGI_API(specular_gi_device)(anerend_ddgi_t* h, const void* dDepth, const void* dGeom0, uint32_t width, uint32_t height, void* dOut) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!dDepth || !dGeom0 || !dOut) return fail(h, ANEREND_DDGI_E_INVALID, "null argument");
  return specularGi(h, (const float*)dDepth, (const float4*)dGeom0, width, height, (uint2*)dOut);
}
int DDGI_API(specular_gi)(anerend_ddgi_t* h, const float* depth, const float* geom0, uint32_t width, uint32_t height, void* outRgba16f) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!depth || !geom0 || !outRgba16f) return fail(h, ANEREND_DDGI_E_INVALID, "null argument");
  if (!h->haveUbo) return fail(h, ANEREND_DDGI_E_STATE, "specular_gi before set_scene_ubo");
  const size_t n = (size_t)width * height;
  if (n == 0 || !(h->ubo.flags & ANEREND_UBO_RT_ON_FLAG) || !(h->ubo.flags & ANEREND_UBO_SPECULAR_GI_FLAG)) return 0;
  Backend& be = h->be;
  float* dD = nullptr; float4* dG = nullptr; uint2* dO = nullptr;
  if (!be.alloc(&dD, n) || !be.alloc(&dG, n) || !be.alloc(&dO, n)) { be.free(dD); be.free(dG); be.free(dO); return fail(h, ANEREND_DDGI_E_CUDA, std::string("specular_gi: ") + be.errStr()); }
  int rc = (be.h2d(dD, depth, n * sizeof(float)) && be.h2d(dG, geom0, n * sizeof(float4))) ? 0 : fail(h, ANEREND_DDGI_E_CUDA, std::string("specular_gi: ") + be.errStr());
  if (!rc) rc = specularGi(h, dD, dG, width, height, dO);
  if (!rc && !be.d2h(outRgba16f, dO, n * sizeof(uint2))) rc = fail(h, ANEREND_DDGI_E_CUDA, std::string("specular_gi: ") + be.errStr());
  be.sync(); be.free(dD); be.free(dG); be.free(dO);
  return rc;
}

// ---- atlas consumer ------------------------------------------------------------------------
namespace {
int indirectDiffuse(anerend_ddgi* h, const float4* dPoints, uint32_t count, const float viewPos[3], int clampSampler, float4* dOut) {
  { int rf = fenceExchange(h); if (rf) return rf; }
  if (!h->haveUbo) return fail(h, ANEREND_DDGI_E_STATE, "indirect_diffuse before set_scene_ubo");
  { int rct = texturesUsable(h); if (rct) return rct; }
  ConsumerCtx c{};
  c.t = sceneTraceCtx(h); c.t.clampSampler = clampSampler ? 1 : 0;
  gridFrame(h, c.t.camOff, c.t.step);        // as sampleProbe places the grid (pbr_light.glsl:274-301), from the CURRENT scene UBO
  c.points = dPoints; c.out = dOut; c.viewPos = mk3(viewPos[0], viewPos[1], viewPos[2]);
  DDGI_LAUNCH(h->be, k_indirect_diffuse, indirect_diffuse_body, count, c, 128);
  BE_CHECK(h, h->be.check());
  return 0;
}
}  // namespace
int DDGI_API(indirect_diffuse_device)(anerend_ddgi_t* h, const void* dPoints, uint32_t count, const float viewPos[3], int32_t clampToEdge, void* dOut) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!dPoints || !dOut || !viewPos) return fail(h, ANEREND_DDGI_E_INVALID, "null argument");
  return indirectDiffuse(h, (const float4*)dPoints, count, viewPos, clampToEdge, (float4*)dOut);
}
int DDGI_API(indirect_diffuse)(anerend_ddgi_t* h, const float* points, uint32_t count, const float viewPos[3], int32_t clampToEdge, float* outRgba) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!points || !outRgba || !viewPos) return fail(h, ANEREND_DDGI_E_INVALID, "null argument");
  if (count == 0) return 0;
  Backend& be = h->be;
  float4 *dP = nullptr, *dO = nullptr;
  if (!be.alloc(&dP, 3 * (size_t)count) || !be.alloc(&dO, count)) { be.free(dP); be.free(dO); return fail(h, ANEREND_DDGI_E_CUDA, std::string("indirect_diffuse: ") + be.errStr()); }
  int rc = be.h2d(dP, points, (size_t)count * 48) ? 0 : fail(h, ANEREND_DDGI_E_CUDA, std::string("indirect_diffuse: ") + be.errStr());
  if (!rc) rc = indirectDiffuse(h, dP, count, viewPos, clampToEdge, dO);
  if (!rc && !be.d2h(outRgba, dO, (size_t)count * 16)) rc = fail(h, ANEREND_DDGI_E_CUDA, std::string("indirect_diffuse: ") + be.errStr());
  be.sync(); be.free(dP); be.free(dO);
  return rc;
}

int DDGI_API(atlas_dims)(const anerend_ddgi_t* h, int32_t* w, int32_t* hh) { if (!h) return ANEREND_DDGI_E_INVALID; *w = h->W; *hh = h->H; return 0; }
int DDGI_API(download_atlas)(anerend_ddgi_t* h, void* out, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (bytes != (size_t)h->W * h->H * 8) return fail(h, ANEREND_DDGI_E_INVALID, "atlas size mismatch");
  { int rf = fenceExchange(h); if (rf) return rf; }
  BE_CHECK(h, h->be.d2h(out, h->dAtlas, bytes)); return 0;
}
// Row stripe: this handle updates only the probe rows z = offset + k*stride of every layer it launches (sharding inside
// the layers: with R ranks, rank r takes (r, R) and every rank carries the same share of every layer's work).
int DDGI_API(set_row_stripe)(anerend_ddgi_t* h, int32_t offset, int32_t stride) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (stride < 1 || offset < 0 || offset >= stride || offset >= h->nXZ) return fail(h, ANEREND_DDGI_E_INVALID, "row stripe out of range");
  h->rowOffset = offset; h->rowStride = stride;
  h->Pl = ((h->nXZ - offset + stride - 1) / stride) * h->nXZ;
  return 0;
}
// Contiguous variant: the probe rows z in [first, first + count) of every layer — a spatially compact share (the handle's rays stay
// in one part of the scene, so its share of the hierarchy stays cached) where the scene's cost is even along z.
int DDGI_API(set_row_block)(anerend_ddgi_t* h, int32_t first, int32_t count) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (first < 0 || count < 1 || first + count > h->nXZ) return fail(h, ANEREND_DDGI_E_INVALID, "row block out of range");
  h->rowOffset = first; h->rowStride = 1; h->Pl = count * h->nXZ;
  return 0;
}
// The atlas rows of this handle's stripe in all layers, packed in (layer, row) order: rows * nY blocks of 10 atlas rows.
int DDGI_API(download_atlas_stripe)(anerend_ddgi_t* h, void* out, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  const size_t rowBytes = (size_t)10 * h->W * 8, rows = (size_t)(h->Pl / h->nXZ);
  if (bytes != rowBytes * rows * h->nY) return fail(h, ANEREND_DDGI_E_INVALID, "stripe size mismatch");
  for (int y = 0; y < h->nY; ++y)
    if (!h->be.d2hStrided((char*)out + rowBytes * rows * y, rowBytes, (const char*)h->dAtlas + rowBytes * ((size_t)y * h->nXZ + h->rowOffset),
                          rowBytes * h->rowStride, rowBytes, rows))
      return fail(h, ANEREND_DDGI_E_CUDA, std::string("download_atlas_stripe: ") + h->be.errStr());
  BE_CHECK(h, h->be.sync()); return 0;
}
// The stripe on the read-back stream (see _download_atlas_async).  Allowed while a peer exchange is connected: the rows of this
// handle's stripe are written by its own convolution only (the peers push THEIR rows), and that convolution waits for the copy.
int DDGI_API(download_atlas_stripe_async)(anerend_ddgi_t* h, void* out, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  const size_t rowBytes = (size_t)10 * h->W * 8, rows = (size_t)(h->Pl / h->nXZ);
  if (!out || bytes != rowBytes * rows * h->nY) return fail(h, ANEREND_DDGI_E_INVALID, "stripe size mismatch");
  BE_CHECK(h, h->be.d2hStridedBackground(out, rowBytes * rows, rowBytes, (const char*)h->dAtlas + rowBytes * h->rowOffset, rowBytes * h->nXZ,
                                         rowBytes * h->rowStride, rowBytes, rows, (size_t)h->nY));
  return 0;
}
int DDGI_API(download_atlas_layers)(anerend_ddgi_t* h, const int32_t* layers, uint32_t n, void* out, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  const size_t layerBytes = (size_t)h->W * h->W * 8;       // 10*nXZ rows of 10*nXZ RGBA16F texels
  if (!layers || bytes != layerBytes * n) return fail(h, ANEREND_DDGI_E_INVALID, "layer block size mismatch");
  { int rf = fenceExchange(h); if (rf) return rf; }
  for (uint32_t i = 0; i < n; ++i) if (layers[i] < 0 || layers[i] >= h->nY) return fail(h, ANEREND_DDGI_E_INVALID, "layer out of range");
  for (uint32_t i = 0; i < n; ++i)
    if (!h->be.d2hAsync((char*)out + layerBytes * i, (const char*)h->dAtlas + layerBytes * (size_t)layers[i], layerBytes))
      return fail(h, ANEREND_DDGI_E_CUDA, std::string("download_atlas_layers: ") + h->be.errStr());
  BE_CHECK(h, h->be.sync()); return 0;
}
// Asynchronous read-back: the copy starts when the work queued so far has finished and runs on its own stream, so the next
// update's tracing overlaps it; the next writer of the atlas (convolution, translation, reset) waits for it on the device.
// `rgba16f` should be page-locked for the overlap to happen.  _download_wait blocks the host until the bytes have landed.
int DDGI_API(download_atlas_async)(anerend_ddgi_t* h, void* out, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!out || bytes != (size_t)h->W * h->H * 8) return fail(h, ANEREND_DDGI_E_INVALID, "atlas size mismatch");
  if (h->xworld > 1 || h->numPeers > 0) return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "asynchronous read-back of the whole atlas while peers write it (use _download_atlas_stripe_async)");
  BE_CHECK(h, h->be.d2hBackground(out, h->dAtlas, bytes)); return 0;
}
int DDGI_API(download_wait)(anerend_ddgi_t* h) { if (!h) return ANEREND_DDGI_E_INVALID; BE_CHECK(h, h->be.waitDownload()); return 0; }
int DDGI_API(upload_atlas)(anerend_ddgi_t* h, const void* in, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (bytes != (size_t)h->W * h->H * 8) return fail(h, ANEREND_DDGI_E_INVALID, "atlas size mismatch");
  { int rf = fenceExchange(h); if (rf) return rf; }
  BE_CHECK(h, h->be.waitDownload()); BE_CHECK(h, h->be.sync()); BE_CHECK(h, h->be.h2d(h->dAtlas, in, bytes)); BE_CHECK(h, h->be.sync()); return 0;
}
int DDGI_API(atlas_device_ptr)(anerend_ddgi_t* h, void** p, size_t* bytes) { if (!h) return ANEREND_DDGI_E_INVALID; *p = h->dAtlas; *bytes = (size_t)h->W * h->H * 8; return 0; }
int DDGI_API(export_atlas_device)(anerend_ddgi_t* h, void* dDst, size_t pitch) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  const size_t rowBytes = (size_t)h->W * 8;
  if (!dDst || pitch < rowBytes) return fail(h, ANEREND_DDGI_E_INVALID, "export_atlas_device: null destination or pitch below the row size");
  { int rf = fenceExchange(h); if (rf) return rf; }
#ifndef DDGI_HOST_EMU
  BE_CHECK(h, h->be.ok(cudaMemcpy2DAsync(dDst, pitch, h->dAtlas, rowBytes, rowBytes, (size_t)h->H, cudaMemcpyDeviceToDevice, h->be.stream)));
#else
  for (int y = 0; y < h->H; ++y) memcpy((char*)dDst + pitch * y, (const char*)h->dAtlas + rowBytes * y, rowBytes);
#endif
  ++h->be.launches;
  return 0;
}
int DDGI_API(set_peer_atlases)(anerend_ddgi_t* h, void* const* peers, uint32_t n) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (n > 8) return fail(h, ANEREND_DDGI_E_INVALID, "at most 8 peers");
  for (uint32_t i = 0; i < n; ++i) h->peers[i] = (uint2*)peers[i];
  h->numPeers = (int)n; return 0;
}
int DDGI_API(exchange_export)(anerend_ddgi_t* h, void* blob, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
#ifndef DDGI_HOST_EMU
  if (!blob || bytes != ANEREND_DDGI_EXCHANGE_BLOB_BYTES) return fail(h, ANEREND_DDGI_E_INVALID, "exchange blob size mismatch");
  static_assert(2 * sizeof(cudaIpcMemHandle_t) == ANEREND_DDGI_EXCHANGE_BLOB_BYTES, "blob = two cudaIpcMemHandle_t");
  Backend& be = h->be;
  if (!h->dXflags) {
    BE_CHECK(h, be.alloc(&h->dXflags, 64)); BE_CHECK(h, be.alloc(&h->dXerr, 1));
    BE_CHECK(h, be.zero(h->dXflags, 64 * sizeof(uint32_t))); BE_CHECK(h, be.zero(h->dXerr, sizeof(uint32_t))); BE_CHECK(h, be.sync());
  }
  cudaIpcMemHandle_t hs[2];
  BE_CHECK(h, be.ok(cudaIpcGetMemHandle(&hs[0], h->dAtlas))); BE_CHECK(h, be.ok(cudaIpcGetMemHandle(&hs[1], h->dXflags)));
  memcpy(blob, hs, sizeof(hs));
  return 0;
#else
  (void)blob; (void)bytes; return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "no peer exchange in the host harness");
#endif
}
int DDGI_API(exchange_connect)(anerend_ddgi_t* h, const void* blobs, uint32_t world, uint32_t rank) {
  if (!h) return ANEREND_DDGI_E_INVALID;
#ifndef DDGI_HOST_EMU
  if (!blobs || world < 2 || world > 8 || rank >= world) return fail(h, ANEREND_DDGI_E_INVALID, "exchange_connect: 2..8 ranks");
  if (!h->dXflags) return fail(h, ANEREND_DDGI_E_STATE, "exchange_connect before exchange_export");
  exchangeDisconnect(h);
  Backend& be = h->be;
  const cudaIpcMemHandle_t* hs = (const cudaIpcMemHandle_t*)blobs;
  int np = 0;
  for (uint32_t r = 0; r < world; ++r) {
    if (r == rank) { h->xflags[r] = h->dXflags; continue; }
    void *pa = nullptr, *pf = nullptr;
    cudaIpcMemHandle_t a, f; memcpy(&a, &hs[2 * r], sizeof(a)); memcpy(&f, &hs[2 * r + 1], sizeof(f));
    if (!be.ok(cudaIpcOpenMemHandle(&pa, a, cudaIpcMemLazyEnablePeerAccess)) || !be.ok(cudaIpcOpenMemHandle(&pf, f, cudaIpcMemLazyEnablePeerAccess))) {
      if (pa) cudaIpcCloseMemHandle(pa);
      exchangeDisconnect(h);
      return fail(h, ANEREND_DDGI_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + be.errStr());
    }
    h->xopened[h->numOpened++] = pa; h->xopened[h->numOpened++] = pf;
    h->peers[np++] = (uint2*)pa; h->xflags[r] = (uint32_t*)pf;
  }
  h->numPeers = np; h->xworld = world; h->xrank = rank;
  if (!h->xStream && !getenv("DDGI_EXCHANGE_SYNC")) {          // DDGI_EXCHANGE_SYNC: push inside the update, on its stream (A/B measurements)
    BE_CHECK(h, be.ok(cudaStreamCreateWithFlags(&h->xStream, cudaStreamNonBlocking)));
    BE_CHECK(h, be.ok(cudaEventCreateWithFlags(&h->evConv, cudaEventDisableTiming))); BE_CHECK(h, be.ok(cudaEventCreateWithFlags(&h->evXDone, cudaEventDisableTiming)));
  }
  return 0;
#else
  (void)blobs; (void)world; (void)rank; return fail(h, ANEREND_DDGI_E_UNSUPPORTED, "no peer exchange in the host harness");
#endif
}
// Makes the handle's stream wait (on the device) until the rows the peers pushed during the last update have landed: for whoever reads
// the atlas on that stream outside this library, and for timing an update sequence to its true end.
int DDGI_API(exchange_fence)(anerend_ddgi_t* h) { if (!h) return ANEREND_DDGI_E_INVALID; return fenceExchange(h); }
int DDGI_API(exchange_disconnect)(anerend_ddgi_t* h) { if (!h) return ANEREND_DDGI_E_INVALID; exchangeDisconnect(h); return 0; }
int DDGI_API(stream)(anerend_ddgi_t* h, void** s) { if (!h) return ANEREND_DDGI_E_INVALID; *s = (void*)h->be.stream; return 0; }

int DDGI_API(ray_tile)(const anerend_ddgi_t* h, int32_t* sx, int32_t* sy) { if (!h) return ANEREND_DDGI_E_INVALID; *sx = h->Sx; *sy = h->Sy; return 0; }
int DDGI_API(download_rays)(anerend_ddgi_t* h, int32_t layer, void* irrOut, void* dirOut, size_t bytesEach) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (layer < 0 || layer >= h->nY || !h->layerTraced[layer]) return fail(h, ANEREND_DDGI_E_STATE, "no rays for this layer");
  const size_t n = (size_t)h->P * h->N;
  if (bytesEach != n * 8) return fail(h, ANEREND_DDGI_E_INVALID, "ray image size mismatch");
  std::vector<uint2> lin(n); std::vector<float4> dirs(h->N);
  BE_CHECK(h, h->be.d2h(lin.data(), h->dRayIrr + (size_t)layer * n, n * 8));
  BE_CHECK(h, h->be.d2h(dirs.data(), h->dDirs + (size_t)layer * h->N, (size_t)h->N * sizeof(float4)));
  const int IW = h->Sx * h->nXZ;
  for (int probe = 0; probe < h->P; ++probe) for (int ray = 0; ray < h->N; ++ray) {
    int px = ray % h->Sx, py = ray / h->Sx, X = (probe % h->nXZ) * h->Sx + px, Y = (probe / h->nXZ) * h->Sy + py;   // rgen:51-58
    size_t o = (size_t)Y * IW + X;
    if (irrOut) ((uint2*)irrOut)[o] = lin[(size_t)probe * h->N + ray];
    if (dirOut) { const float4 d = dirs[ray];                                                                          // rgen:108
      ((uint2*)dirOut)[o] = make_uint2((uint32_t)f2h(d.x) | ((uint32_t)f2h(d.y) << 16), (uint32_t)f2h(d.z) | (0x3C00u << 16)); }
  }
  return 0;
}
// IrradianceProbeConvolve on its own (BASELINE config 1): the caller supplies ProbeRaysIrrTex of `layer`; ProbeRaysDirTex is what
// IrradianceProbeRT would have written for randomRot = rot (rgen:87-88,108), made by the same code as in traceLayers.
int DDGI_API(upload_ray_radiance)(anerend_ddgi_t* h, int32_t layer, const void* irr, size_t bytes, const float rot[16]) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (layer < 0 || layer >= h->nY || !irr || !rot) return fail(h, ANEREND_DDGI_E_INVALID, "layer out of range");
  if (h->rowStride != 1) return fail(h, ANEREND_DDGI_E_STATE, "upload_ray_radiance with a row stripe set");
  const size_t n = (size_t)h->P * h->N;
  if (bytes != n * 8) return fail(h, ANEREND_DDGI_E_INVALID, "ray image size mismatch");
  std::vector<uint2> lin(n);
  const int IW = h->Sx * h->nXZ;
  for (int probe = 0; probe < h->P; ++probe) for (int ray = 0; ray < h->N; ++ray) {
    int px = ray % h->Sx, py = ray / h->Sx, X = (probe % h->nXZ) * h->Sx + px, Y = (probe / h->nXZ) * h->Sy + py;   // rgen:51-58
    lin[(size_t)probe * h->N + ray] = ((const uint2*)irr)[(size_t)Y * IW + X];
  }
  Backend& be = h->be;
  BE_CHECK(h, be.sync());
  BE_CHECK(h, be.h2d(h->dRayIrr + (size_t)layer * n, lin.data(), n * 8));
  be.waitCopy();
  h->hLayerMap[0] = layer; h->numLaunchLayers = 1;
#ifdef DDGI_HOST_EMU
  rotateDirs(h, layer, rot);
  BE_CHECK(h, be.h2d(h->dDirs + (size_t)layer * h->N, h->hDirs + (size_t)layer * h->N, (size_t)h->N * sizeof(float4)));
  BE_CHECK(h, be.h2d(h->dLayerMap, h->hLayerMap, sizeof(int)));
#else
  DirsCtx dc{};
  dc.fib = h->dFib; dc.dDirs = h->dDirs; dc.dLayerMap = h->dLayerMap; dc.N = (uint32_t)h->N; dc.first = 0; dc.count = 1; dc.layer[0] = layer;
  for (int c3 = 0; c3 < 3; ++c3) for (int r3 = 0; r3 < 3; ++r3) dc.m[0][c3 * 3 + r3] = rot[c3 * 4 + r3];
  k_make_dirs<<<dim3(((uint32_t)h->N + 127u) / 128u, 1), 128, 0, be.stream>>>(dc);
  ++be.launches;
#endif
  BE_CHECK(h, be.sync());            // `lin` is pageable: the copy is complete before it goes away
  h->layerTraced[layer] = 1;
  return 0;
}
int DDGI_API(download_ray_records)(anerend_ddgi_t* h, int32_t layer, anerend_ddgi_ray_record_t* out, size_t count) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->dRecords) return fail(h, ANEREND_DDGI_E_STATE, "FLAG_RAY_RECORDS not set");
  if (layer < 0 || layer >= h->nY || count != (size_t)h->P * h->N || !h->layerTraced[layer]) return fail(h, ANEREND_DDGI_E_INVALID, "no records for this layer / count mismatch");
  BE_CHECK(h, h->be.d2h(out, h->dRecords + (size_t)layer * count, count * sizeof(*out))); return 0;
}
int DDGI_API(download_texel_records)(anerend_ddgi_t* h, float* out, size_t bytes) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (!h->dTexelRecords) return fail(h, ANEREND_DDGI_E_STATE, "FLAG_TEXEL_RECORDS not set");
  if (bytes != (size_t)h->W * h->H * 16) return fail(h, ANEREND_DDGI_E_INVALID, "size mismatch");
  BE_CHECK(h, h->be.d2h(out, h->dTexelRecords, bytes)); return 0;
}

int DDGI_API(set_pipeline)(anerend_ddgi_t* h, int32_t on) {
  if (!h) return ANEREND_DDGI_E_INVALID;
#ifndef DDGI_HOST_EMU
  if (!on && h->pStream) BE_CHECK(h, h->be.ok(cudaStreamSynchronize(h->pStream)));
  h->pipeline = on != 0; h->pipeStateDirty = true;
  if (on && h->fastRays) { BE_CHECK(h, h->be.sync()); int rcp = ensurePipeline(h); if (rcp) return rcp; }      // stream, events and buffers now, not inside the first pipelined update
#endif
  return 0;
}
int DDGI_API(set_timers)(anerend_ddgi_t* h, int32_t level) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  if (level < 0 || level > 2) return fail(h, ANEREND_DDGI_E_INVALID, "timer level: 0 none, 1 passes, 2 passes + kernels");
  h->be.timerLevel = level;
  h->timed[0] = h->timed[1] = h->timed[2] = h->timed[3] = false; h->timedKernels = false;
  return 0;
}
int DDGI_API(get_timings)(anerend_ddgi_t* h, anerend_ddgi_timings_t* o) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  BE_CHECK(h, h->be.sync());
  memset(o, 0, sizeof(*o));
  if (h->be.timerLevel == 0) { o->accel_ms = h->accelMs; o->kernel_launches = h->be.launches; return 0; }
  if (h->timed[0]) o->trans_ms = h->be.elapsed(0, 1);
  if (h->timed[1]) o->rt_ms = h->be.elapsed(1, 2);
  if (h->timed[2]) o->convolve_ms = h->be.elapsed(2, 3);
  if (h->timed[1] && h->timedKernels) for (int k = 0; k < 4; ++k) o->rt_kernel_ms[k] = h->be.elapsed(6 + k, 7 + k);
  if (h->timed[3]) { h->accelMs = h->be.elapsed(4, 5); h->timed[3] = false; }
  o->accel_ms = h->accelMs;
  o->total_ms = o->trans_ms + o->rt_ms + o->convolve_ms;
  o->kernel_launches = h->be.launches;
  return 0;
}
int DDGI_API(get_stats)(anerend_ddgi_t* h, anerend_ddgi_stats_t* o) {
  if (!h) return ANEREND_DDGI_E_INVALID;
  unsigned long long s[16];
  BE_CHECK(h, h->be.d2h(s, h->dStats, sizeof(s)));
  o->primary_rays = s[0]; o->primary_hits = s[1]; o->shadow_rays = s[2]; o->primary_nodes_visited = s[3]; o->primary_tris_tested = s[4];
  o->shadow_nodes_visited = s[5]; o->shadow_tris_tested = s[6]; o->shadow_rays_prefiltered = s[7]; o->shadow_rays_candidate = s[8];
  o->primary_level_changes = s[9]; o->shadow_level_changes = s[10];
  return 0;
}
int DDGI_API(reset_stats)(anerend_ddgi_t* h) { if (!h) return ANEREND_DDGI_E_INVALID; BE_CHECK(h, h->be.zero(h->dStats, 16 * sizeof(unsigned long long))); return 0; }

// glm::rotate(mat4(1), radians(deg), axis) (IrradianceProbeRayTracingPass.cpp:133; glm/ext/matrix_transform.inl:18-46)
int DDGI_API(make_rotation)(int32_t axisIdx, float deg, float out[16]) {
  if (axisIdx < 0 || axisIdx > 2 || !out) return ANEREND_DDGI_E_INVALID;
  const float a = deg * 0.01745329251994329576923690768489f;
  const float c = cosf(a), s = sinf(a);
  float axis[3] = {0.0f, 0.0f, 0.0f}; axis[axisIdx] = 1.0f;
  const float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
  float R[3][3];
  R[0][0] = c + temp[0] * axis[0]; R[0][1] = temp[0] * axis[1] + s * axis[2]; R[0][2] = temp[0] * axis[2] - s * axis[1];
  R[1][0] = temp[1] * axis[0] - s * axis[2]; R[1][1] = c + temp[1] * axis[1]; R[1][2] = temp[1] * axis[2] + s * axis[0];
  R[2][0] = temp[2] * axis[0] + s * axis[1]; R[2][1] = temp[2] * axis[1] - s * axis[0]; R[2][2] = c + temp[2] * axis[2];
  const float m[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int i = 0; i < 3; ++i) for (int r = 0; r < 4; ++r) out[i * 4 + r] = m[0][r] * R[i][0] + m[1][r] * R[i][1] + m[2][r] * R[i][2];
  for (int r = 0; r < 4; ++r) out[12 + r] = m[3][r];
  return 0;
}
int DDGI_API(fibonacci_dirs)(int32_t n, float* out) {
  if (n < 1 || !out) return ANEREND_DDGI_E_INVALID;
  for (int i = 0; i < n; ++i) fibonacciDir((uint32_t)i, (uint32_t)n, out + 3 * i);
  return 0;
}
const char* DDGI_API(version)(void) {
#ifdef DDGI_HOST_EMU
  return "anerend_ddgi host-emu debug harness (not the product)";
#else
  return "anerend_cuda_ddgi 0.1 (sm_100a)";
#endif
}

}  // extern "C"
