// ref_host_shim.cpp — oracle/_ref: the parts of the REFERENCE'S OWN HOST CODE on the DDGI path that compile in this image, built from
// the sources where they lie under /root/reference (oracle/Makefile, target _ref/libanerend_ref_host.so; never copied into this repo):
//   * src/anerend/render/GpuBuffers.h and Vertex.h — the structs behind the bindless buffers and the scene UBO;
//   * src/anerend/render/RenderOptions.h — the defaults the synthetic scenes' UBO starts from;
//   * contrib/glm — glm::rotate / glm::radians exactly as IrradianceProbeRayTracingPass.cpp:115-133 calls them (A3).
// TEST INFRASTRUCTURE ONLY: loaded by tests/ and tests/golden/make_ref_host_golden.py; nothing under anerend_b200/ uses it.
// The GLSL stages and the Vulkan host code cannot be built here; their compiled SPIR-V is executed instead.
#include <cstddef>
#include <cstdio>
#include <cstring>

#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>

#include "GpuBuffers.h"
#include "RenderOptions.h"
#include "Vertex.h"

namespace {
struct Field { const char* name; size_t offset, size; };
#define F(S, f) Field{#S "." #f, offsetof(S, f), sizeof(((S*)nullptr)->f)}
#define SZ(S) Field{#S, 0, sizeof(S)}
using namespace render;
using namespace render::gpu;
const Field kFields[] = {
  SZ(Vertex), F(Vertex, pos), F(Vertex, color), F(Vertex, normal), F(Vertex, tangent), F(Vertex, jointWeights), F(Vertex, uv), F(Vertex, padding),
  F(Vertex, jointIds), F(Vertex, padding1), F(Vertex, padding2),
  SZ(GPURenderable), F(GPURenderable, _transform), F(GPURenderable, _bounds), F(GPURenderable, _tint), F(GPURenderable, _modelOffset),
  F(GPURenderable, _numMeshes), F(GPURenderable, _skeletonOffset), F(GPURenderable, _visible), F(GPURenderable, _firstMaterialIndex),
  F(GPURenderable, _dynamicModelOffset), F(GPURenderable, _terrainOffset),
  SZ(GPUMaterialInfo), F(GPUMaterialInfo, _baseColFac), F(GPUMaterialInfo, _emissive), F(GPUMaterialInfo, _bindlessIndices), F(GPUMaterialInfo, _metRough),
  SZ(GPUMeshInfo), F(GPUMeshInfo, _minPos), F(GPUMeshInfo, _maxPos), F(GPUMeshInfo, _vertexOffset), F(GPUMeshInfo, _indexOffset), F(GPUMeshInfo, _blasRef),
  SZ(GPULight), F(GPULight, _worldPos), F(GPULight, _color),
  SZ(GPUTileInfo), F(GPUTileInfo, _ddgiAtlasTex),
  SZ(GPUSceneData), F(GPUSceneData, view), F(GPUSceneData, proj), F(GPUSceneData, invView), F(GPUSceneData, invProj), F(GPUSceneData, invViewProj),
  F(GPUSceneData, directionalShadowMatrixProj), F(GPUSceneData, directionalShadowMatrixView), F(GPUSceneData, cameraPos), F(GPUSceneData, cameraGridPos),
  F(GPUSceneData, lightDir), F(GPUSceneData, bakeTileInfo), F(GPUSceneData, skyColor), F(GPUSceneData, time), F(GPUSceneData, delta),
  F(GPUSceneData, screenWidth), F(GPUSceneData, screenHeight), F(GPUSceneData, sunIntensity), F(GPUSceneData, skyIntensity), F(GPUSceneData, exposure),
  F(GPUSceneData, flags), F(GPUSceneData, bloomThresh), F(GPUSceneData, bloomKnee),
};
struct Flag { const char* name; unsigned value; };
#define FL(x) Flag{#x, (unsigned)x}
const Flag kFlags[] = {FL(UBO_RT_SHADOWS_FLAG), FL(UBO_DDGI_FLAG), FL(UBO_DDGI_MULTI_FLAG), FL(UBO_SPECULAR_GI_FLAG), FL(UBO_RT_ON_FLAG), FL(UBO_BAKE_MODE_ON_FLAG)};
}  // namespace

extern "C" {
// "Struct.field offset size" per line (a struct's own line carries its sizeof), then "flag NAME value" lines; returns the length.
size_t ref_host_layout(char* buf, size_t cap) {
  size_t n = 0;
  for (const Field& f : kFields) n += (size_t)snprintf(buf ? buf + n : nullptr, buf && cap > n ? cap - n : 0, "%s %zu %zu\n", f.name, f.offset, f.size);
  for (const Flag& f : kFlags) n += (size_t)snprintf(buf ? buf + n : nullptr, buf && cap > n ? cap - n : 0, "flag %s %u\n", f.name, f.value);
  return n;
}
// RenderOptions{} as the renderer starts with: sunIntensity, skyIntensity, exposure, skyColor.rgb, bloomThresh, bloomKnee, then the
// switches raytracingEnabled, ddgiEnabled, multiBounceDdgiEnabled, specularGiEnabled, raytracedShadows as 0/1
void ref_host_render_option_defaults(float out[13]) {
  render::RenderOptions o;
  const float v[13] = {o.sunIntensity, o.skyIntensity, o.exposure, o.skyColor.x, o.skyColor.y, o.skyColor.z, o.bloomThresh, o.bloomKnee,
                       (float)o.raytracingEnabled, (float)o.ddgiEnabled, (float)o.multiBounceDdgiEnabled, (float)o.specularGiEnabled, (float)o.raytracedShadows};
  std::memcpy(out, v, sizeof(v));
}
// IrradianceProbeRayTracingPass.cpp:120-133: rotAx from the axis index, glm::rotate(glm::mat4(1.0f), glm::radians(float(angle)), rotAx)
void ref_host_make_rotation(int axis, float angle_deg, float out[16]) {
  const glm::vec3 axes[3] = {glm::vec3(1.0f, 0.0f, 0.0f), glm::vec3(0.0f, 1.0f, 0.0f), glm::vec3(0.0f, 0.0f, 1.0f)};
  const glm::mat4 randomRot = glm::rotate(glm::mat4(1.0f), glm::radians(angle_deg), axes[axis < 0 ? 0 : axis > 2 ? 2 : axis]);
  std::memcpy(out, &randomRot[0][0], 64);
}
}
