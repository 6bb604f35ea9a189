// anerend_ddgi_passes.hpp — header-only adapter: the three DDGI passes (plus TLASUpdate / BLASUpdate in front of them and ShadowRT /
// SpecularGIRT on the same acceleration structure) as AneRend `RenderPass` subclasses that call
// the C ABI (anerend_cuda_ddgi.h) instead of recording Vulkan ray-tracing/compute/transfer commands.
//
// Mirrors, name for name:
//   src/anerend/render/passes/IrradianceProbeTranslationPass.{h,cpp}   "IrradianceProbeTrans"
//   src/anerend/render/passes/IrradianceProbeRayTracingPass.{h,cpp}    "IrradianceProbeRT"
//   src/anerend/render/passes/IrradianceProbeConvolvePass.{h,cpp}      "IrradianceProbeConvolve"
// The classes are templates over an `Engine` traits struct so the same header compiles against the real engine headers
//   struct Engine { using RenderPass = render::RenderPass; using FrameGraphBuilder = render::FrameGraphBuilder;
//                   using RenderContext = render::RenderContext; using RenderExeParams = render::RenderExeParams;
//                   using RenderPassRegisterInfo = render::RenderPassRegisterInfo; using ResourceUsage = render::ResourceUsage;
//                   using Access = render::Access; using Stage = render::Stage; using Type = render::Type;
//                   using ImageInitialCreateInfo = render::ImageInitialCreateInfo;
//                   static constexpr auto kFormatRGBA16F = VK_FORMAT_R16G16B16A16_SFLOAT; static constexpr auto kFormatR32F = VK_FORMAT_R32_SFLOAT;
//                   static constexpr auto kLayoutShaderReadOnly = VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL;
//                   static auto clearConvTexCb() { return clearConvTex; } };      // IrradianceProbeConvolvePass.cpp:9-47, moved to a header
// or against the Vulkan-free stub used by tests/cpp/test_passes.cpp.
//
// Frame-graph resources.  The three passes declare the SAME four resources, with the same access / stage / type and the same
// ImageInitialCreateInfo, as the reference passes (IrradianceProbeRayTracingPass.cpp:42-81, IrradianceProbeConvolvePass.cpp:66-98,
// IrradianceProbeTranslationPass.cpp:29-51), so the frame graph creates "ProbeRaysConvTex" exactly as before and the passes that
// consume it (DeferredLightingRenderPass.cpp:192-200, SpecularGIRTPass.cpp:66, DebugViewRenderPass.cpp:225) find their producer.
// The CUDA side keeps its own working atlas; after every convolution, translation and (re-)registration the adapter publishes it
// into the engine's image through rc->cudaResource("ProbeRaysConvTex") + anerend_cuda_ddgi_export_atlas_device (see INTEGRATION.md
// for how the engine imports that image's memory and orders the CUDA stream against its queue).  "ProbeRaysIrrTex" / "ProbeRaysDirTex"
// / "TempProbeTex" stay declared — the graph's barriers and the debug view keep their names — and are backed by CUDA buffers.  What the engine must provide is exactly what the
// reference passes already use: rc->getNumIrradianceProbesXZ/Y(), rc->getRenderOptions(), rc->isBaking(),
// rc->blackboardValueInt / setBlackboardValueInt (RenderContext.h:85-188), fgb.registerRenderPass / registerRenderPassExe
// (FrameGraphBuilder.h:178-179), and one addition: rc->sceneUbo() returning the GPUSceneData the renderer already fills
// every frame (VulkanRenderer.cpp:1201-1243).
#pragma once
#include <cstdio>
#include <memory>
#include <random>
#include <string>

#include "anerend_cuda_ddgi.h"

namespace render {
namespace cuda_ddgi {

// Shared by the three passes: the opaque handle replaces the vault-owned images ProbeRaysIrrTex / ProbeRaysDirTex /
// ProbeRaysConvTex / TempProbeTex and the TLAS.
struct Volume {
  anerend_ddgi_t* handle = nullptr;
  anerend_ddgi_params_t params{};
  std::string error;
  bool topologyDirty = true;      // set by the engine when instances appear / disappear / flip visibility: TLASUpdate rebuilds instead of refitting
  uint32_t flags = 0;             // ANEREND_DDGI_FLAG_* for the handle, set before the first registerToGraph; ANEREND_DDGI_FLAG_TWO_LEVEL gives
                                  // the reference's own structure: one hierarchy per mesh (VulkanRenderer.cpp:1267-1417) and an instance level
                                  // that "TLASUpdate" rebuilds every frame (UpdateTLASPass.cpp:198-228)
  ~Volume() { if (handle) anerend_cuda_ddgi_destroy(handle); }
  // Called from registerToGraph: the reference re-registers the same pass objects on swap-chain recreation
  // (VulkanRenderer.cpp:1962-2000), re-creating and re-clearing the atlas and resetting _currentLayer.
  bool ensure(int numProbesXZ, int numProbesY) {
    if (handle && params.num_probes_xz == numProbesXZ && params.num_probes_y == numProbesY) return anerend_cuda_ddgi_reset(handle) == 0;
    if (handle) { anerend_cuda_ddgi_destroy(handle); handle = nullptr; }
    params = anerend_ddgi_params_t{};
    params.struct_size = sizeof(params);
    params.num_probes_xz = numProbesXZ; params.num_probes_y = numProbesY;
    params.rays_per_probe = 256;                                   // probe_helpers.glsl:4
    params.probe_step[0] = 1.0f; params.probe_step[1] = 2.0f; params.probe_step[2] = 1.0f;   // probe_helpers.glsl:3
    params.t_min = 0.1f; params.t_max = 50.0f;                     // rgen:91-92
    params.hysteresis = 0.96f;                                     // probe_conv.comp:85
    params.slab_begin = 0; params.slab_end = numProbesY; params.device = -1; params.flags = flags;
    int rc = anerend_cuda_ddgi_create(&params, &handle);
    if (rc != 0) { handle = nullptr; error = "anerend_cuda_ddgi_create failed (no CUDA device?)"; std::printf("%s\n", error.c_str()); return false; }
    anerend_cuda_ddgi_set_timers(handle, 0);      // FrameGraphBuilder times the exes itself (FrameGraphBuilder.cpp:276-278): no events of our own between the launches
    return true;
  }
  // The engine's "ProbeRaysConvTex" follows the CUDA atlas: called after whatever changed it (convolution, translation, clear).
  template <class RC> void publish(RC* rc) {
    if (!handle) return;
    void* img = rc->cudaResource(ANEREND_RES_CONV);
    if (!img) return;                                              // engine without the import: consumers see the image as the graph cleared it
    if (anerend_cuda_ddgi_export_atlas_device(handle, img, (size_t)(8 + 2) * params.num_probes_xz * 8) != 0) std::printf("%s\n", anerend_cuda_ddgi_last_error(handle));
  }
};

// the reference's resource declarations, shared by the three passes
template <class Engine>
typename Engine::ResourceUsage probeUsage(const char* name, typename Engine::Access access, typename Engine::Stage stage, typename Engine::Type type) {
  typename Engine::ResourceUsage u{};
  u._resourceName = name;
  u._access.set((std::size_t)access);
  u._stage.set((std::size_t)stage);
  u._type = type;
  return u;
}
template <class Engine>
typename Engine::ImageInitialCreateInfo probeImage(uint32_t width, uint32_t height) {
  typename Engine::ImageInitialCreateInfo ci{};
  ci._initialWidth = width; ci._initialHeight = height; ci._intialFormat = Engine::kFormatRGBA16F;
  return ci;
}

template <class Engine>
class IrradianceProbeTranslationPass : public Engine::RenderPass {
public:
  explicit IrradianceProbeTranslationPass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext* rc) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = ANEREND_PASS_TRANS;
    info._group = ANEREND_PASS_GROUP;
    const uint32_t width = (8 + 2) * (uint32_t)rc->getNumIrradianceProbesXZ(), height = width * (uint32_t)rc->getNumIrradianceProbesY();
    {  // IrradianceProbeTranslationPass.cpp:29-44: the copy target, created here
      auto u = probeUsage<Engine>(ANEREND_RES_TEMP, Engine::Access::Write, Engine::Stage::Transfer, Engine::Type::ImageTransferDst);
      u._imageCreateInfo = probeImage<Engine>(width, height);
      info._resourceUsages.emplace_back(std::move(u));
    }
    info._resourceUsages.emplace_back(probeUsage<Engine>(ANEREND_RES_CONV, Engine::Access::Read, Engine::Stage::Transfer, Engine::Type::ImageTransferSrc));   // :45-51
    fgb.registerRenderPass(std::move(info));
    _vol->ensure((int)rc->getNumIrradianceProbesXZ(), (int)rc->getNumIrradianceProbesY());
    _vol->publish(rc);                                             // the cleared atlas (clearConvTex) reaches the engine's image as well
    fgb.registerRenderPassExe(ANEREND_PASS_TRANS, [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().ddgiEnabled) return;   // IrradianceProbeTranslationPass.cpp:57
      if (exeParams.rc->isBaking()) return;                        // :58
      anerend_cuda_ddgi_set_scene_ubo(_vol->handle, exeParams.rc->sceneUbo());
      if (anerend_cuda_ddgi_translate(_vol->handle) != 0) std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
      // no publish here: nothing samples the atlas between this pass and IrradianceProbeConvolve, which publishes it
    });
  }
private:
  std::shared_ptr<Volume> _vol;
};

template <class Engine>
class IrradianceProbeRayTracingPass : public Engine::RenderPass {
public:
  explicit IrradianceProbeRayTracingPass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext* rc) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = ANEREND_PASS_RT;
    info._group = ANEREND_PASS_GROUP;
    {  // IrradianceProbeRayTracingPass.cpp:42-81 (sqrtNumRays = 16: NUM_RAYS_PER_PROBE = 256)
      const uint32_t side = 16u * (uint32_t)rc->getNumIrradianceProbesXZ();
      for (const char* n : {ANEREND_RES_RAYS_IRR, ANEREND_RES_RAYS_DIR}) {
        auto u = probeUsage<Engine>(n, Engine::Access::Write, Engine::Stage::RayTrace, Engine::Type::ImageStorage);
        u._imageCreateInfo = probeImage<Engine>(side, side);
        info._resourceUsages.emplace_back(std::move(u));
      }
      info._resourceUsages.emplace_back(probeUsage<Engine>(ANEREND_RES_CONV, Engine::Access::Read, Engine::Stage::RayTrace, Engine::Type::SampledTexture));
    }
    fgb.registerRenderPass(std::move(info));
    const int numProbesHeight = (int)rc->getNumIrradianceProbesY();
    _currentLayer = 0;                                             // IrradianceProbeRayTracingPass.cpp:39
    fgb.registerRenderPassExe(ANEREND_PASS_RT, [this, numProbesHeight](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().raytracingEnabled) return;   // :97
      if (!exeParams.rc->getRenderOptions().ddgiEnabled) return;         // :98
      uint32_t probeLayer = _currentLayer;                               // :112
      // Random orientation matrix (:115-133); the seed source is the engine's choice, the matrix maths is the ABI's
      std::uniform_real_distribution<> angle(0.0, 360.0);
      std::uniform_int_distribution<> axis(0, 2);
      int ax = axis(_rng);
      float rot[16];
      anerend_cuda_ddgi_make_rotation(ax, float(angle(_rng)), rot);
      anerend_cuda_ddgi_set_scene_ubo(_vol->handle, exeParams.rc->sceneUbo());
      if (anerend_cuda_ddgi_trace_layer(_vol->handle, (int32_t)probeLayer, rot) != 0) std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
      exeParams.rc->setBlackboardValueInt(ANEREND_BLACKBOARD_LAYER, (int)_currentLayer);   // :158
      _currentLayer++;
      if (_currentLayer >= static_cast<uint32_t>(numProbesHeight)) _currentLayer = 0;      // :160-164
    });
  }
  void seed(uint32_t s) { _rng.seed(s); }
private:
  std::shared_ptr<Volume> _vol;
  uint32_t _currentLayer = 0;
  std::mt19937 _rng{std::random_device{}()};
};

template <class Engine>
class IrradianceProbeConvolvePass : public Engine::RenderPass {
public:
  explicit IrradianceProbeConvolvePass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext* rc) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = ANEREND_PASS_CONVOLVE;
    info._group = ANEREND_PASS_GROUP;
    {  // IrradianceProbeConvolvePass.cpp:66-98: the atlas is CREATED by this pass's usage (initial layout shader-read, cleared by clearConvTex)
      info._resourceUsages.emplace_back(probeUsage<Engine>(ANEREND_RES_RAYS_IRR, Engine::Access::Read, Engine::Stage::Compute, Engine::Type::ImageStorage));
      info._resourceUsages.emplace_back(probeUsage<Engine>(ANEREND_RES_RAYS_DIR, Engine::Access::Read, Engine::Stage::Compute, Engine::Type::ImageStorage));
      auto u = probeUsage<Engine>(ANEREND_RES_CONV, Engine::Access::Write, Engine::Stage::Compute, Engine::Type::ImageStorage);
      const uint32_t width = (8 + 2) * (uint32_t)rc->getNumIrradianceProbesXZ();
      auto ci = probeImage<Engine>(width, width * (uint32_t)rc->getNumIrradianceProbesY());
      ci._initialLayout = Engine::kLayoutShaderReadOnly;           // "hack since this is used in a pass before this one" (:90)
      ci._initialDataCb = Engine::clearConvTexCb();
      u._imageCreateInfo = ci;
      info._resourceUsages.emplace_back(std::move(u));
    }
    fgb.registerRenderPass(std::move(info));
    fgb.registerRenderPassExe(ANEREND_PASS_CONVOLVE, [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().raytracingEnabled) return;   // IrradianceProbeConvolvePass.cpp:111
      if (!exeParams.rc->getRenderOptions().ddgiEnabled) return;         // :112
      uint32_t probeLayer = (uint32_t)exeParams.rc->blackboardValueInt(ANEREND_BLACKBOARD_LAYER);   // :126
      if (anerend_cuda_ddgi_convolve_layer(_vol->handle, (int32_t)probeLayer) != 0) std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
      _vol->publish(exeParams.rc);                                 // the consumers' "ProbeRaysConvTex" (deferred lighting, debug view) follows
    });
  }
private:
  std::shared_ptr<Volume> _vol;
};

// ---- the other ray passes on the same acceleration structure ---------------------------------------
// They read and write frame-graph images.  The engine hands those over as CUDA device pointers through ONE more accessor,
//   void* rc->cudaResource(const char* resourceName)   (the image's memory, imported once via VK_KHR_external_memory)
// for the resource names the reference passes declare (ShadowRayTracingPass.cpp:18-36, SpecularGIRTPass.cpp:19-88); a null
// pointer turns the pass into a no-op.  Formats: GeometryDepthImage r32f, RTShadowImage r32f, Geometry0Image rgba32f (normal in
// rgb), ReflectTex rgba16f; all of the swap-chain extent (rc->screenWidth() x rc->screenHeight()).

template <class Engine>
class ShadowRayTracingPass : public Engine::RenderPass {          // src/anerend/render/passes/ShadowRayTracingPass.{h,cpp}, "ShadowRT"
public:
  explicit ShadowRayTracingPass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext* rc) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = "ShadowRT";
    info._resourceUsages.emplace_back(probeUsage<Engine>("GeometryDepthImage", Engine::Access::Read, Engine::Stage::RayTrace, Engine::Type::SampledDepthTexture));   // :22
    {  // :30-41: created here, r32f of the swap-chain extent
      auto u = probeUsage<Engine>("RTShadowImage", Engine::Access::Write, Engine::Stage::RayTrace, Engine::Type::ImageStorage);
      typename Engine::ImageInitialCreateInfo ci{};
      ci._initialWidth = rc->screenWidth(); ci._initialHeight = rc->screenHeight(); ci._intialFormat = Engine::kFormatR32F;
      u._imageCreateInfo = ci;
      info._resourceUsages.emplace_back(std::move(u));
    }
    fgb.registerRenderPass(std::move(info));
    fgb.registerRenderPassExe("ShadowRT", [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      void* depth = exeParams.rc->cudaResource("GeometryDepthImage");
      void* out = exeParams.rc->cudaResource("RTShadowImage");
      if (!depth || !out) return;
      anerend_cuda_ddgi_set_scene_ubo(_vol->handle, exeParams.rc->sceneUbo());      // UBO_RT_SHADOWS_FLAG gates the pass (shadow.rgen:32-34)
      if (anerend_cuda_ddgi_shadow_rt_device(_vol->handle, depth, exeParams.rc->screenWidth(), exeParams.rc->screenHeight(), out) != 0)
        std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
    });
  }
private:
  std::shared_ptr<Volume> _vol;
};

template <class Engine>
class SpecularGIRTPass : public Engine::RenderPass {              // src/anerend/render/passes/SpecularGIRTPass.{h,cpp}, "SpecularGIRT"
public:
  explicit SpecularGIRTPass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext* rc) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = "SpecularGIRT";
    info._group = "SpecularGI";
    {  // :30-46: created here, rgba16f ("needs to be HDR") of the swap-chain extent
      auto u = probeUsage<Engine>("ReflectTex", Engine::Access::Write, Engine::Stage::RayTrace, Engine::Type::ImageStorage);
      u._imageCreateInfo = probeImage<Engine>(rc->screenWidth(), rc->screenHeight());
      info._resourceUsages.emplace_back(std::move(u));
    }
    info._resourceUsages.emplace_back(probeUsage<Engine>("GeometryDepthImage", Engine::Access::Read, Engine::Stage::RayTrace, Engine::Type::SampledDepthTexture));   // :49
    for (const char* n : {"Geometry0Image", ANEREND_RES_CONV, "Geometry1Image", "SpecularBRDFLutTex"})                                                            // :57-85
      info._resourceUsages.emplace_back(probeUsage<Engine>(n, Engine::Access::Read, Engine::Stage::RayTrace, Engine::Type::SampledTexture));
    fgb.registerRenderPass(std::move(info));
    fgb.registerRenderPassExe("SpecularGIRT", [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().raytracingEnabled) return;   // SpecularGIRTPass.cpp:102
      if (!exeParams.rc->getRenderOptions().specularGiEnabled) return;   // :103
      void* depth = exeParams.rc->cudaResource("GeometryDepthImage");
      void* geom0 = exeParams.rc->cudaResource("Geometry0Image");
      void* out = exeParams.rc->cudaResource("ReflectTex");
      if (!depth || !geom0 || !out) return;
      anerend_cuda_ddgi_set_scene_ubo(_vol->handle, exeParams.rc->sceneUbo());
      if (anerend_cuda_ddgi_specular_gi_device(_vol->handle, depth, geom0, exeParams.rc->screenWidth(), exeParams.rc->screenHeight(), out) != 0)
        std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
    });
  }
private:
  std::shared_ptr<Volume> _vol;
};

// -----------------------------------------------------------------------------------------------------------------
// The acceleration-structure passes in front of the chain.  The reference fills the TLAS instance buffer with
// tlas_update.comp and rebuilds the TLAS every frame ("TLASUpdate", UpdateTLASPass.cpp:77-248), and refits the BLASes of skinned
// meshes from the giga vertex buffer ("BLASUpdate", UpdateBlasPass.cpp:14-207).  Here both become calls on the volume's handle:
// the device RenderableBuffer is re-flattened and the hierarchy refitted; deforming vertices are copied first.  The engine
// tells the adapter when the set of instances changed (an asset upload, a visibility flip) through Volume::topologyDirty —
// then the next TLASUpdate rebuilds (anerend_cuda_ddgi_build_accel) instead of refitting.  Two additions to RenderContext next
// to cudaResource(name): getCurrentNumRenderables() (already there, RenderContext.h:135) and dynamicVertexRange(first, count),
// the range of GigaVtxBuffer the skinning pass wrote this frame (false: nothing moved).

template <class Engine>
class UpdateTLASPass : public Engine::RenderPass {                // src/anerend/render/passes/UpdateTLASPass.{h,cpp}, "TLASUpdate"
public:
  explicit UpdateTLASPass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext*) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = "TLASUpdate";
    // UpdateTLASPass.cpp:80-117.  The two buffers the compute shader filled stay declared (other passes and the graph's barriers know
    // their names) but carry no create info here: nothing on the CUDA side needs them, so they are marked engine-owned.
    for (const char* n : {"TLASInstanceBuf", "TLASCounterBuf"}) {
      auto u = probeUsage<Engine>(n, Engine::Access::Write, Engine::Stage::Compute, Engine::Type::SSBO);
      u._ownedByEngine = true;
      info._resourceUsages.emplace_back(std::move(u));
    }
    { auto u = probeUsage<Engine>("RenderableBuffer", Engine::Access::Read, Engine::Stage::Compute, Engine::Type::SSBO);
      u._multiBuffered = true; u._ownedByEngine = true; info._resourceUsages.emplace_back(std::move(u)); }
    fgb.registerRenderPass(std::move(info));
    fgb.registerRenderPassExe("TLASUpdate", [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().raytracingEnabled) return;                       // UpdateTLASPass.cpp:130
      int rc = 0;
      if (_vol->topologyDirty) {                                                              // instances changed: select + flatten + build
        rc = anerend_cuda_ddgi_build_accel(_vol->handle);
        if (rc == 0) _vol->topologyDirty = false;
      } else {
        void* rend = exeParams.rc->cudaResource("RenderableBuffer");                         // GPU-written earlier in the frame (cull.comp:303)
        if (!rend) return;
        rc = anerend_cuda_ddgi_refit_accel_device(_vol->handle, rend, (uint32_t)exeParams.rc->getCurrentNumRenderables());
      }
      if (rc != 0) std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
    });
  }
private:
  std::shared_ptr<Volume> _vol;
};

template <class Engine>
class UpdateBlasPass : public Engine::RenderPass {                // src/anerend/render/passes/UpdateBlasPass.{h,cpp}, "BLASUpdate"
public:
  explicit UpdateBlasPass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext*) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = "BLASUpdate";
    { auto u = probeUsage<Engine>("GigaVtxBuffer", Engine::Access::Read, Engine::Stage::Vertex, Engine::Type::SSBO);               // UpdateBlasPass.cpp:24-32
      u._ownedByEngine = true; info._resourceUsages.emplace_back(std::move(u)); }
    fgb.registerRenderPass(std::move(info));
    fgb.registerRenderPassExe("BLASUpdate", [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().raytracingEnabled) return;                       // UpdateBlasPass.cpp:37
      uint32_t first = 0, count = 0;
      if (!exeParams.rc->dynamicVertexRange(&first, &count) || count == 0) return;           // no dynamic BLAS this frame (:41)
      const char* vtx = static_cast<const char*>(exeParams.rc->cudaResource("GigaVtxBuffer"));
      if (!vtx) return;
      // the following TLASUpdate refits triangles, boxes and shading records from the new vertices
      if (anerend_cuda_ddgi_update_vertices_device(_vol->handle, vtx + (size_t)first * sizeof(anerend_vertex_t), first, count) != 0)
        std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
    });
  }
private:
  std::shared_ptr<Volume> _vol;
};

}  // namespace cuda_ddgi
}  // namespace render
