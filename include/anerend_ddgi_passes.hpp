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
//                   using RenderPassRegisterInfo = render::RenderPassRegisterInfo; using ResourceUsage = render::ResourceUsage; };
// or against the Vulkan-free stub used by tests/cpp/test_passes.cpp.  What the engine must provide is exactly what the
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
    params.slab_begin = 0; params.slab_end = numProbesY; params.device = -1;
    int rc = anerend_cuda_ddgi_create(&params, &handle);
    if (rc != 0) { handle = nullptr; error = "anerend_cuda_ddgi_create failed (no CUDA device?)"; std::printf("%s\n", error.c_str()); return false; }
    return true;
  }
};

template <class Engine>
class IrradianceProbeTranslationPass : public Engine::RenderPass {
public:
  explicit IrradianceProbeTranslationPass(std::shared_ptr<Volume> v) : _vol(std::move(v)) {}
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext* rc) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = ANEREND_PASS_TRANS;
    info._group = ANEREND_PASS_GROUP;
    fgb.registerRenderPass(std::move(info));
    _vol->ensure((int)rc->getNumIrradianceProbesXZ(), (int)rc->getNumIrradianceProbesY());
    fgb.registerRenderPassExe(ANEREND_PASS_TRANS, [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().ddgiEnabled) return;   // IrradianceProbeTranslationPass.cpp:57
      if (exeParams.rc->isBaking()) return;                        // :58
      anerend_cuda_ddgi_set_scene_ubo(_vol->handle, exeParams.rc->sceneUbo());
      if (anerend_cuda_ddgi_translate(_vol->handle) != 0) std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
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
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext*) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = ANEREND_PASS_CONVOLVE;
    info._group = ANEREND_PASS_GROUP;
    fgb.registerRenderPass(std::move(info));
    fgb.registerRenderPassExe(ANEREND_PASS_CONVOLVE, [this](typename Engine::RenderExeParams exeParams) {
      if (!_vol->handle) return;
      if (!exeParams.rc->getRenderOptions().raytracingEnabled) return;   // IrradianceProbeConvolvePass.cpp:111
      if (!exeParams.rc->getRenderOptions().ddgiEnabled) return;         // :112
      uint32_t probeLayer = (uint32_t)exeParams.rc->blackboardValueInt(ANEREND_BLACKBOARD_LAYER);   // :126
      if (anerend_cuda_ddgi_convolve_layer(_vol->handle, (int32_t)probeLayer) != 0) std::printf("%s\n", anerend_cuda_ddgi_last_error(_vol->handle));
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
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext*) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = "ShadowRT";
    { typename Engine::ResourceUsage u{}; u._resourceName = "GeometryDepthImage"; info._resourceUsages.emplace_back(std::move(u)); }   // :22
    { typename Engine::ResourceUsage u{}; u._resourceName = "RTShadowImage"; info._resourceUsages.emplace_back(std::move(u)); }        // :30
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
  void registerToGraph(typename Engine::FrameGraphBuilder& fgb, typename Engine::RenderContext*) override final {
    typename Engine::RenderPassRegisterInfo info{};
    info._name = "SpecularGIRT";
    info._group = "SpecularGI";
    for (const char* n : {"ReflectTex", "GeometryDepthImage", "Geometry0Image", ANEREND_RES_CONV, "Geometry1Image", "SpecularBRDFLutTex"}) {   // :30-82
      typename Engine::ResourceUsage u{}; u._resourceName = n; info._resourceUsages.emplace_back(std::move(u));
    }
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
    for (const char* n : {"TLASInstanceBuf", "TLASCounterBuf", "RenderableBuffer"}) {        // UpdateTLASPass.cpp:80-117
      typename Engine::ResourceUsage u{}; u._resourceName = n; info._resourceUsages.emplace_back(std::move(u));
    }
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
    { typename Engine::ResourceUsage u{}; u._resourceName = "GigaVtxBuffer"; info._resourceUsages.emplace_back(std::move(u)); }   // UpdateBlasPass.cpp:24-32
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
