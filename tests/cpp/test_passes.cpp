// Vulkan-free stub of the engine types the reference's passes use (RenderPass.h:14-35, FrameGraphBuilder.h:22-42,150-185,
// RenderContext.h:85-188, RenderOptions.h:7-28), enough to register the three CUDA DDGI passes into a "frame graph",
// execute it for a number of frames in submission order and print a checksum of the atlas.
//   usage: test_passes <frames> <seed> [scene.bin]     (scene.bin written by tests/test_passes_adapter.py)
#include <bitset>
#include <cstdint>
#include <cstdio>
#include <optional>
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
// FrameGraphBuilder.h:44-176, the fields the passes set
enum class Access { Read, Write, Count };
enum class Stage { Transfer, Compute, IndirectDraw, Vertex, Fragment, Irrelevant, RayTrace, Count };
enum class Type { ColorAttachment, DepthAttachment, Present, SampledTexture, SampledDepthTexture, UBO, SSBO, Irrelevant, ImageStorage, ImageTransferSrc, ImageTransferDst, Count };
struct ImageInitialCreateInfo {
  std::uint32_t _initialWidth = 0, _initialHeight = 0; int _intialFormat = 0; int _initialLayout = 0;
  std::function<void(RenderContext*, void*&)> _initialDataCb = nullptr;
};
struct ResourceUsage {
  std::bitset<(std::size_t)Access::Count> _access; std::bitset<(std::size_t)Stage::Count> _stage; Type _type = Type::Irrelevant;
  bool _ownedByEngine = false, _multiBuffered = false; std::string _resourceName; std::optional<ImageInitialCreateInfo> _imageCreateInfo;
};
struct RenderPassRegisterInfo { std::string _name; std::string _group; std::vector<ResourceUsage> _resourceUsages; };
struct FrameGraphBuilder {
  std::vector<RenderPassRegisterInfo> passes; std::map<std::string, std::function<void(RenderExeParams)>> exes;
  void registerRenderPass(RenderPassRegisterInfo&& i) { passes.emplace_back(std::move(i)); }
  void registerRenderPassExe(const std::string& n, std::function<void(RenderExeParams)> f) { exes[n] = std::move(f); }
  void executeGraph(RenderContext* rc) { for (auto& p : passes) exes[p._name](RenderExeParams{rc}); }   // submission order, FrameGraphBuilder.cpp:1197-1228
  void reset() { passes.clear(); exes.clear(); }
  // What building the graph checks first (FrameGraphBuilder.cpp, resource creation): every resource a pass touches is either owned by the
  // engine or created by exactly one usage that carries a create info.  Returns the first offending "pass:resource", or "" if the graph is sound.
  std::string validate(RenderContext* rc) {
    std::map<std::string, int> creators;
    for (auto& p : passes) for (auto& u : p._resourceUsages) if (u._imageCreateInfo) {
      creators[u._resourceName]++;
      if (u._imageCreateInfo->_initialWidth == 0 || u._imageCreateInfo->_initialHeight == 0 || u._imageCreateInfo->_intialFormat == 0) return p._name + ":" + u._resourceName + " (empty create info)";
      if (u._imageCreateInfo->_initialDataCb) { void* img = nullptr; u._imageCreateInfo->_initialDataCb(rc, img); }
    }
    for (auto& p : passes) for (auto& u : p._resourceUsages) {
      if (u._ownedByEngine) continue;
      if (u._access.none() || u._stage.none() || u._type == Type::Irrelevant) return p._name + ":" + u._resourceName + " (no access / stage / type)";
      if (creators[u._resourceName] != 1) return p._name + ":" + u._resourceName + (creators[u._resourceName] ? " (created twice)" : " (no producer)");
    }
    return "";
  }
};
struct RenderPass {
  virtual ~RenderPass() = default;
  virtual void registerToGraph(FrameGraphBuilder&, RenderContext*) = 0;
};
static int g_convCleared = 0;
struct Engine {
  using RenderPass = stub::RenderPass; using FrameGraphBuilder = stub::FrameGraphBuilder; using RenderContext = stub::RenderContext;
  using RenderExeParams = stub::RenderExeParams; using RenderPassRegisterInfo = stub::RenderPassRegisterInfo; using ResourceUsage = stub::ResourceUsage;
  using Access = stub::Access; using Stage = stub::Stage; using Type = stub::Type; using ImageInitialCreateInfo = stub::ImageInitialCreateInfo;
  static constexpr int kFormatRGBA16F = 97, kFormatR32F = 100, kLayoutShaderReadOnly = 5;        // the VkFormat / VkImageLayout values
  static std::function<void(RenderContext*, void*&)> clearConvTexCb() { return [](RenderContext*, void*&) { ++g_convCleared; }; }   // clearConvTex stand-in
};
// a consumer of the atlas as the reference declares it (DeferredLightingRenderPass.cpp:192-200): sampled in a compute pass
struct DeferredLightingStub : RenderPass {
  void registerToGraph(FrameGraphBuilder& fgb, RenderContext*) override {
    RenderPassRegisterInfo info{}; info._name = "DeferredLighting";
    ResourceUsage u{}; u._resourceName = "ProbeRaysConvTex"; u._access.set((std::size_t)Access::Read); u._stage.set((std::size_t)Stage::Compute); u._type = Type::SampledTexture;
    info._resourceUsages.emplace_back(std::move(u));
    fgb.registerRenderPass(std::move(info));
    fgb.registerRenderPassExe("DeferredLighting", [](RenderExeParams) {});
  }
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
  if (const char* e = getenv("DDGI_TEST_FLAGS")) vol->flags = (uint32_t)strtoul(e, nullptr, 0);      // e.g. ANEREND_DDGI_FLAG_TWO_LEVEL
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
  {  // frame-graph soundness: with the three passes registered, a consumer of "ProbeRaysConvTex" finds its producer (and the image its
     // create info + clear callback); without them it does not, and the stub graph says so
    stub::DeferredLightingStub deferred;
    stub::FrameGraphBuilder alone; deferred.registerToGraph(alone, &rc);
    if (alone.validate(&rc) != "DeferredLighting:ProbeRaysConvTex (no producer)") { printf("FAIL the stub graph accepts a consumer without a producer\n"); return 1; }
    stub::FrameGraphBuilder g5;
    for (auto* p : passes) p->registerToGraph(g5, &rc);
    deferred.registerToGraph(g5, &rc);
    const std::string bad = g5.validate(&rc);
    if (!bad.empty()) { printf("FAIL frame graph: %s\n", bad.c_str()); return 1; }
    if (stub::g_convCleared != 1) { printf("FAIL clearConvTex not attached to the atlas' create info\n"); return 1; }
    const auto& cu = g5.passes[2]._resourceUsages;                          // IrradianceProbeConvolvePass.cpp:66-98
    if (cu.size() != 3 || cu[2]._resourceName != "ProbeRaysConvTex" || !cu[2]._imageCreateInfo || cu[2]._imageCreateInfo->_initialWidth != 10 * rc.nXZ ||
        cu[2]._imageCreateInfo->_initialHeight != 10 * rc.nXZ * rc.nY || cu[2]._imageCreateInfo->_initialLayout != stub::Engine::kLayoutShaderReadOnly ||
        !cu[2]._access.test((std::size_t)stub::Access::Write) || cu[2]._type != stub::Type::ImageStorage) { printf("FAIL atlas create info\n"); return 1; }
    const auto& ru = g5.passes[1]._resourceUsages;                          // IrradianceProbeRayTracingPass.cpp:42-81
    if (ru.size() != 3 || ru[0]._resourceName != "ProbeRaysIrrTex" || ru[1]._resourceName != "ProbeRaysDirTex" || !ru[0]._imageCreateInfo ||
        ru[0]._imageCreateInfo->_initialWidth != 16 * rc.nXZ || ru[2]._type != stub::Type::SampledTexture) { printf("FAIL ray image usages\n"); return 1; }
    const auto& tu = g5.passes[0]._resourceUsages;                          // IrradianceProbeTranslationPass.cpp:29-51
    if (tu.size() != 2 || tu[0]._resourceName != "TempProbeTex" || tu[0]._type != stub::Type::ImageTransferDst || tu[1]._type != stub::Type::ImageTransferSrc) { printf("FAIL translation usages\n"); return 1; }
    // the whole set the adapter provides, wired to the engine-owned inputs they declare
    stub::FrameGraphBuilder g6;
    blas.registerToGraph(g6, &rc); tlas.registerToGraph(g6, &rc);
    for (auto* p : passes) p->registerToGraph(g6, &rc);
    for (const char* n : {"GeometryDepthImage", "Geometry0Image", "Geometry1Image", "SpecularBRDFLutTex"}) {       // produced by the G-buffer passes
      stub::RenderPassRegisterInfo gi{}; gi._name = std::string("Producer_") + n;
      stub::ResourceUsage u{}; u._resourceName = n; u._access.set((std::size_t)stub::Access::Write); u._stage.set((std::size_t)stub::Stage::Fragment); u._type = stub::Type::ColorAttachment;
      stub::ImageInitialCreateInfo ci{}; ci._initialWidth = 4; ci._initialHeight = 4; ci._intialFormat = 1; u._imageCreateInfo = ci;
      gi._resourceUsages.emplace_back(std::move(u)); g6.registerRenderPass(std::move(gi));
    }
    rc.ubo.screen_width = 64; rc.ubo.screen_height = 32;
    shadowRt.registerToGraph(g6, &rc); specRt.registerToGraph(g6, &rc);
    const std::string bad6 = g6.validate(&rc);
    if (!bad6.empty()) { printf("FAIL full frame graph: %s\n", bad6.c_str()); return 1; }
    rc.ubo.screen_width = 0; rc.ubo.screen_height = 0;
    if (vol->handle) anerend_cuda_ddgi_reset(vol->handle);
  }
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
  // the engine's "ProbeRaysConvTex": device memory of the same size (here: the atlas allocation of a second, otherwise unused handle);
  // the adapter publishes the CUDA atlas into it after every convolution
  anerend_ddgi_t* engineImage = nullptr;
  if (anerend_cuda_ddgi_create(&vol->params, &engineImage) == 0) {
    void* img = nullptr; size_t bytes = 0; anerend_cuda_ddgi_atlas_device_ptr(engineImage, &img, &bytes);
    rc.cuda["ProbeRaysConvTex"] = img;
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
  if (engineImage) {                                                          // what a consumer of the engine's image would sample
    anerend_cuda_ddgi_sync(vol->handle);
    std::vector<uint16_t> pub(atlas.size());
    anerend_cuda_ddgi_download_atlas(engineImage, pub.data(), pub.size() * 2);
    if (pub != atlas) { printf("FAIL the published ProbeRaysConvTex differs from the atlas\n"); return 1; }
    printf("OK published\n");
  }
  fgb.reset();
  for (auto* p : passes) p->registerToGraph(fgb, &rc);
  anerend_cuda_ddgi_download_atlas(vol->handle, atlas.data(), atlas.size() * 2);
  for (size_t i = 0; i < atlas.size(); i += 4) if (atlas[i] || atlas[i + 1] || atlas[i + 2] || atlas[i + 3] != 0x3C00) { printf("FAIL re-register clear\n"); return 1; }
  fgb.executeGraph(&rc);
  if (rc.blackboard["ProbeLayer"] != 0) { printf("FAIL layer reset\n"); return 1; }
  printf("OK re-register\n");
  return 0;
}
