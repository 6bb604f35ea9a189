// Vulkan-free stub of the engine types the reference's passes use (RenderPass.h:14-35, FrameGraphBuilder.h:22-42,150-185,
// RenderContext.h:85-188, RenderOptions.h:7-28), enough to register the three CUDA DDGI passes into a "frame graph",
// execute it for a number of frames in submission order and print a checksum of the atlas.
//   usage: test_passes <frames> <seed> [scene.bin]     (scene.bin written by tests/test_passes_adapter.py)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/anerend_ddgi_passes.hpp"

namespace stub {
struct RenderOptions { bool raytracingEnabled = true; bool ddgiEnabled = true; bool multiBounceDdgiEnabled = true; bool specularGiEnabled = true; };
struct RenderContext {
  RenderOptions opts; anerend_scene_ubo_t ubo{}; std::map<std::string, int> blackboard; bool baking = false;
  size_t nXZ = 8, nY = 4;
  size_t getNumIrradianceProbesXZ() { return nXZ; }
  size_t getNumIrradianceProbesY() { return nY; }
  RenderOptions& getRenderOptions() { return opts; }
  bool isBaking() { return baking; }
  int blackboardValueInt(const std::string& k) { return blackboard[k]; }
  void setBlackboardValueInt(const std::string& k, int v) { blackboard[k] = v; }
  const anerend_scene_ubo_t* sceneUbo() { return &ubo; }
  std::map<std::string, void*> cuda;            // device pointers of frame-graph images (none in this stub: the passes no-op)
  void* cudaResource(const char* name) { auto it = cuda.find(name); return it == cuda.end() ? nullptr : it->second; }
  size_t numRenderables = 0; uint32_t dynFirst = 0, dynCount = 0;
  size_t getCurrentNumRenderables() { return numRenderables; }
  bool dynamicVertexRange(uint32_t* first, uint32_t* count) { *first = dynFirst; *count = dynCount; return dynCount != 0; }
  uint32_t screenWidth() { return ubo.screen_width; }
  uint32_t screenHeight() { return ubo.screen_height; }
};
struct RenderExeParams { RenderContext* rc; };
struct ResourceUsage { std::string _resourceName; };
struct RenderPassRegisterInfo { std::string _name; std::string _group; std::vector<ResourceUsage> _resourceUsages; };
struct FrameGraphBuilder {
  std::vector<RenderPassRegisterInfo> passes; std::map<std::string, std::function<void(RenderExeParams)>> exes;
  void registerRenderPass(RenderPassRegisterInfo&& i) { passes.emplace_back(std::move(i)); }
  void registerRenderPassExe(const std::string& n, std::function<void(RenderExeParams)> f) { exes[n] = std::move(f); }
  void executeGraph(RenderContext* rc) { for (auto& p : passes) exes[p._name](RenderExeParams{rc}); }   // submission order, FrameGraphBuilder.cpp:1197-1228
  void reset() { passes.clear(); exes.clear(); }
};
struct RenderPass {
  virtual ~RenderPass() = default;
  virtual void registerToGraph(FrameGraphBuilder&, RenderContext*) = 0;
};
struct Engine {
  using RenderPass = stub::RenderPass; using FrameGraphBuilder = stub::FrameGraphBuilder; using RenderContext = stub::RenderContext;
  using RenderExeParams = stub::RenderExeParams; using RenderPassRegisterInfo = stub::RenderPassRegisterInfo; using ResourceUsage = stub::ResourceUsage;
};
}  // namespace stub

template <class T> static std::vector<T> readVec(FILE* f) {
  uint64_t n = 0; if (fread(&n, 8, 1, f) != 1) return {};
  std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) v.clear(); return v;
}

int main(int argc, char** argv) {
  using namespace render::cuda_ddgi;
  int frames = argc > 1 ? atoi(argv[1]) : 4; uint32_t seed = argc > 2 ? (uint32_t)atoi(argv[2]) : 1u;
  stub::RenderContext rc; stub::FrameGraphBuilder fgb;
  auto vol = std::make_shared<Volume>();
  IrradianceProbeTranslationPass<stub::Engine> trans(vol);
  IrradianceProbeRayTracingPass<stub::Engine> rt(vol);
  IrradianceProbeConvolvePass<stub::Engine> conv(vol);
  ShadowRayTracingPass<stub::Engine> shadowRt(vol);
  SpecularGIRTPass<stub::Engine> specRt(vol);
  {  // the other ray passes: names, group and resource usages as the reference declares them; without images they no-op
    stub::FrameGraphBuilder g2;
    shadowRt.registerToGraph(g2, &rc); specRt.registerToGraph(g2, &rc);
    if (g2.passes.size() != 2 || g2.passes[0]._name != "ShadowRT" || g2.passes[0]._resourceUsages.size() != 2 ||
        g2.passes[0]._resourceUsages[1]._resourceName != "RTShadowImage" || g2.passes[1]._name != "SpecularGIRT" ||
        g2.passes[1]._group != "SpecularGI" || g2.passes[1]._resourceUsages.size() != 6 ||
        g2.passes[1]._resourceUsages[3]._resourceName != "ProbeRaysConvTex") { printf("FAIL ray pass names\n"); return 1; }
    g2.executeGraph(&rc);
  }
  UpdateBlasPass<stub::Engine> blas(vol);
  UpdateTLASPass<stub::Engine> tlas(vol);
  {  // the acceleration-structure passes: names and resource usages as the reference declares them
    stub::FrameGraphBuilder g3;
    blas.registerToGraph(g3, &rc); tlas.registerToGraph(g3, &rc);
    if (g3.passes.size() != 2 || g3.passes[0]._name != "BLASUpdate" || g3.passes[0]._resourceUsages.size() != 1 ||
        g3.passes[0]._resourceUsages[0]._resourceName != "GigaVtxBuffer" || g3.passes[1]._name != "TLASUpdate" ||
        g3.passes[1]._resourceUsages.size() != 3 || g3.passes[1]._resourceUsages[0]._resourceName != "TLASInstanceBuf" ||
        g3.passes[1]._resourceUsages[2]._resourceName != "RenderableBuffer") { printf("FAIL accel pass names\n"); return 1; }
    rc.opts.raytracingEnabled = false; g3.executeGraph(&rc); rc.opts.raytracingEnabled = true;       // gated off: no call, no error
    if (!vol->topologyDirty) { printf("FAIL accel gating\n"); return 1; }
  }
  std::vector<stub::RenderPass*> passes = {&trans, &rt, &conv};              // VulkanRenderer.cpp:4473-4475
  for (auto* p : passes) p->registerToGraph(fgb, &rc);
  if (fgb.passes.size() != 3 || fgb.passes[0]._name != "IrradianceProbeTrans" || fgb.passes[1]._name != "IrradianceProbeRT" ||
      fgb.passes[2]._name != "IrradianceProbeConvolve" || fgb.passes[1]._group != "IrradianceProbe") { printf("FAIL names\n"); return 1; }
  if (!vol->handle) { printf("NO_DEVICE passes registered, handle unavailable: %s\n", vol->error.c_str()); return 0; }
  if (argc > 3) {
    FILE* f = fopen(argv[3], "rb"); if (!f) { printf("FAIL open scene\n"); return 1; }
    auto vtx = readVec<anerend_vertex_t>(f); auto idx = readVec<uint32_t>(f); auto meshes = readVec<anerend_mesh_info_t>(f);
    auto mni = readVec<uint32_t>(f); auto mats = readVec<anerend_material_info_t>(f); auto rends = readVec<anerend_renderable_t>(f);
    auto mi = readVec<uint32_t>(f); auto mti = readVec<uint32_t>(f); auto lights = readVec<anerend_light_t>(f); auto ubo = readVec<anerend_scene_ubo_t>(f);
    fclose(f);
    rc.ubo = ubo[0];
    int r = anerend_cuda_ddgi_upload_scene(vol->handle, vtx.data(), (uint32_t)vtx.size(), idx.data(), (uint32_t)idx.size(), meshes.data(), mni.data(),
                                           (uint32_t)meshes.size(), mats.data(), (uint32_t)mats.size(), rends.data(), (uint32_t)rends.size(), mi.data(),
                                           (uint32_t)mi.size(), mti.data(), (uint32_t)mti.size(), lights.data(), (uint32_t)lights.size());
    if (r != 0) { printf("FAIL scene: %s\n", anerend_cuda_ddgi_last_error(vol->handle)); return 1; }
    // the acceleration structure is built by the TLASUpdate pass (topologyDirty), as the frame graph would run it
    rc.numRenderables = rends.size();
    stub::FrameGraphBuilder g4; blas.registerToGraph(g4, &rc); tlas.registerToGraph(g4, &rc);
    g4.executeGraph(&rc);
    anerend_ddgi_bvh_info_t bi{};
    if (vol->topologyDirty || anerend_cuda_ddgi_get_bvh_info(vol->handle, &bi) != 0 || bi.num_triangles == 0) { printf("FAIL TLASUpdate build\n"); return 1; }
    g4.executeGraph(&rc);                                                      // no device RenderableBuffer in the stub: the refit no-ops
  }
  rt.seed(seed);
  for (int fr = 0; fr < frames; ++fr) {
    rc.ubo.camera_pos[0] += 0.37f;                                             // drifts across probe cells: translation fires
    fgb.executeGraph(&rc);
    if (rc.blackboard["ProbeLayer"] != fr % (int)rc.nY) { printf("FAIL layer cadence\n"); return 1; }
  }
  // swap-chain recreation: the same pass objects register again, atlas re-cleared, layer back to 0 (VulkanRenderer.cpp:1962-2000)
  int32_t w = 0, h = 0; anerend_cuda_ddgi_atlas_dims(vol->handle, &w, &h);
  std::vector<uint16_t> atlas((size_t)w * h * 4);
  anerend_cuda_ddgi_download_atlas(vol->handle, atlas.data(), atlas.size() * 2);
  uint64_t sum = 1469598103934665603ull; for (uint16_t v : atlas) { sum ^= v; sum *= 1099511628211ull; }
  printf("OK frames=%d atlas=%dx%d fnv=%016llx\n", frames, w, h, (unsigned long long)sum);
  fgb.reset();
  for (auto* p : passes) p->registerToGraph(fgb, &rc);
  anerend_cuda_ddgi_download_atlas(vol->handle, atlas.data(), atlas.size() * 2);
  for (size_t i = 0; i < atlas.size(); i += 4) if (atlas[i] || atlas[i + 1] || atlas[i + 2] || atlas[i + 3] != 0x3C00) { printf("FAIL re-register clear\n"); return 1; }
  fgb.executeGraph(&rc);
  if (rc.blackboard["ProbeLayer"] != 0) { printf("FAIL layer reset\n"); return 1; }
  printf("OK re-register\n");
  return 0;
}
