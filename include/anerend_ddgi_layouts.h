/*
 * anerend_ddgi_layouts.h — byte-exact C mirrors of the bindless scene buffers the
 * DDGI probe-update path consumes.  These are the INPUT FORMAT of the C ABI in
 * anerend_cuda_ddgi.h; a host engine hands over its existing buffers unchanged.
 *
 * Reference structs mirrored (paths relative to the AneRend repository):
 *   src/anerend/render/Vertex.h:7-19            -> anerend_vertex_t         (96 B)
 *   src/anerend/render/GpuBuffers.h:15-26       -> anerend_renderable_t     (128 B)
 *   src/anerend/render/GpuBuffers.h:29-36       -> anerend_material_info_t  (64 B)
 *   src/anerend/render/GpuBuffers.h:38-44       -> anerend_mesh_info_t      (48 B)
 *   src/anerend/render/GpuBuffers.h:69-72       -> anerend_light_t          (32 B)
 *   src/anerend/render/GpuBuffers.h:74-76       -> anerend_tile_info_t      (4 B)
 *   src/anerend/render/GpuBuffers.h:125-148     -> anerend_scene_ubo_t      (568 B)
 *   GLSL views: assets/shaders/bindless.glsl:7-121, scene_ubo.glsl:14-37
 */
#ifndef ANEREND_DDGI_LAYOUTS_H
#define ANEREND_DDGI_LAYOUTS_H

#include <stdint.h>

#ifdef __cplusplus
#define ANEREND_STATIC_ASSERT(c, m) static_assert(c, m)
extern "C" {
#else
#define ANEREND_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* Vertex.h:7-19; the shader reads it as 6 vec4 ("PackedVertex", bindless.glsl:60-68,98-109). */
typedef struct anerend_vertex {
  float pos[3];          /* dwords 0-2   */
  float color[3];        /* dwords 3-5   */
  float normal[3];       /* dwords 6-8   */
  float tangent[4];      /* dwords 9-12  */
  float joint_weights[4];/* dwords 13-16 */
  float uv[2];           /* dwords 17-18 */
  float padding;         /* dword 19     */
  int16_t joint_ids[4];  /* dwords 20-21 */
  float padding1;        /* dword 22     */
  float padding2;        /* dword 23     */
} anerend_vertex_t;
ANEREND_STATIC_ASSERT(sizeof(anerend_vertex_t) == 96, "Vertex must be 96 bytes");

/* GpuBuffers.h:15-26 (alignas(16)); GLSL Renderable bindless.glsl:7-20. transform is column-major. */
typedef struct anerend_renderable {
  float transform[16];
  float bounds[4];
  float tint[4];
  uint32_t model_offset;
  uint32_t num_meshes;
  uint32_t skeleton_offset;
  uint32_t visible;              /* bit0 visible, bit1 cull bit (bindless.glsl:249-270) */
  uint32_t first_material_index;
  uint32_t dynamic_model_offset;
  int32_t terrain_offset;
  uint32_t _pad;
} anerend_renderable_t;
ANEREND_STATIC_ASSERT(sizeof(anerend_renderable_t) == 128, "GPURenderable must be 128 bytes");

/* GpuBuffers.h:29-36 */
typedef struct anerend_material_info {
  float base_col_fac[4];
  float emissive[4];
  int32_t bindless_indices[4];   /* metallicRoughness, albedo, normal, emissive; -1 = none */
  float met_rough[4];            /* r metallic, g roughness */
} anerend_material_info_t;
ANEREND_STATIC_ASSERT(sizeof(anerend_material_info_t) == 64, "GPUMaterialInfo must be 64 bytes");

/* One entry of the bindless texture array `textures[]` (bindless.glsl:179, VulkanRenderer::addTextureToBindless,
 * VulkanRenderer.cpp:4916-4946): mip 0 of an asset::Texture (asset/Texture.h:10-40), uncompressed.  The array index is
 * the bindless index the materials refer to.  BC-compressed assets are handed over decoded (RGBA8). */
#define ANEREND_TEX_RGBA8_UNORM 0u   /* asset::Texture::Format::RGBA8_UNORM */
#define ANEREND_TEX_RGBA8_SRGB  1u   /* RGBA8_SRGB: rgb decoded to linear before filtering, alpha linear */
#define ANEREND_TEX_RG8_UNORM   2u   /* RG8_UNORM: (r, g, 0, 1) */
#define ANEREND_TEX_R8_UNORM    3u   /* R8_UNORM: (r, 0, 0, 1) */
#define ANEREND_TEX_RGBA16F     4u   /* RGBA16F_SFLOAT: baked DDGI atlases (VulkanRenderer::downloadDDGIAtlas, VulkanRenderer.cpp:2706-2771) */
typedef struct anerend_texture_desc {
  const void* pixels;            /* width*height texels, rows top to bottom, tightly packed */
  uint32_t width, height;
  uint32_t format;               /* ANEREND_TEX_* */
  uint32_t clamp_to_edge;        /* asset::Texture::_clampToEdge; 0 = REPEAT (PipelineUtil.h:102) */
} anerend_texture_desc_t;

/* GpuBuffers.h:38-44 */
typedef struct anerend_mesh_info {
  float min_pos[4];
  float max_pos[4];
  uint32_t vertex_offset;        /* in vertices */
  uint32_t index_offset;         /* in indices  */
  uint64_t blas_ref;             /* ignored by the CUDA path */
} anerend_mesh_info_t;
ANEREND_STATIC_ASSERT(sizeof(anerend_mesh_info_t) == 48, "GPUMeshInfo must be 48 bytes");

/* GpuBuffers.h:69-72 */
typedef struct anerend_light {
  float world_pos[4];            /* w = range */
  float color[4];                /* w > 0.5 = enabled */
} anerend_light_t;
ANEREND_STATIC_ASSERT(sizeof(anerend_light_t) == 32, "GPULight must be 32 bytes");

/* GpuBuffers.h:74-76 */
typedef struct anerend_tile_info {
  int32_t ddgi_atlas_tex;
} anerend_tile_info_t;
ANEREND_STATIC_ASSERT(sizeof(anerend_tile_info_t) == 4, "GPUTileInfo must be 4 bytes");

/* GpuBuffers.h:125-148 / scene_ubo.glsl:14-37 */
typedef struct anerend_scene_ubo {
  float view[16];
  float proj[16];
  float inv_view[16];
  float inv_proj[16];
  float inv_view_proj[16];
  float directional_shadow_matrix_proj[16];
  float directional_shadow_matrix_view[16];
  float camera_pos[4];
  int32_t camera_grid_pos[4];
  float light_dir[4];
  int32_t bake_tile_info[4];
  float sky_color[4];
  float time;
  float delta;
  uint32_t screen_width;
  uint32_t screen_height;
  float sun_intensity;
  float sky_intensity;
  float exposure;
  uint32_t flags;
  float bloom_thresh;
  float bloom_knee;
} anerend_scene_ubo_t;
ANEREND_STATIC_ASSERT(sizeof(anerend_scene_ubo_t) == 568, "GPUSceneData must be 568 bytes");

/* SceneUboFlags, GpuBuffers.h:108-122 / scene_ubo.glsl:1-12 */
#define ANEREND_UBO_SSAO_FLAG                (1u << 0)
#define ANEREND_UBO_FXAA_FLAG                (1u << 1)
#define ANEREND_UBO_DIRECTIONAL_SHADOWS_FLAG (1u << 2)
#define ANEREND_UBO_RT_SHADOWS_FLAG          (1u << 3)
#define ANEREND_UBO_DDGI_FLAG                (1u << 4)
#define ANEREND_UBO_DDGI_MULTI_FLAG          (1u << 5)
#define ANEREND_UBO_SPECULAR_GI_FLAG         (1u << 6)
#define ANEREND_UBO_SS_PROBES_FLAG           (1u << 7)
#define ANEREND_UBO_BS_VIS_FLAG              (1u << 8)
#define ANEREND_UBO_HACK_FLAG                (1u << 9)
#define ANEREND_UBO_RT_ON_FLAG               (1u << 10)
#define ANEREND_UBO_BAKE_MODE_ON_FLAG        (1u << 11)

/* Frame-graph names kept verbatim (IrradianceProbe*Pass.cpp registerToGraph). */
#define ANEREND_PASS_TRANS      "IrradianceProbeTrans"
#define ANEREND_PASS_RT         "IrradianceProbeRT"
#define ANEREND_PASS_CONVOLVE   "IrradianceProbeConvolve"
#define ANEREND_PASS_GROUP      "IrradianceProbe"
#define ANEREND_RES_RAYS_IRR    "ProbeRaysIrrTex"
#define ANEREND_RES_RAYS_DIR    "ProbeRaysDirTex"
#define ANEREND_RES_CONV        "ProbeRaysConvTex"
#define ANEREND_RES_TEMP        "TempProbeTex"
#define ANEREND_BLACKBOARD_LAYER "ProbeLayer"

#ifdef __cplusplus
}
#endif
#endif
