/*
 * anerend_cuda_ddgi.h — C ABI of the B200-native DDGI irradiance-probe update.
 *
 * This is the drop-in boundary for AneRend's frame-graph pass chain
 *     IrradianceProbeTrans -> IrradianceProbeRT -> IrradianceProbeConvolve
 * (src/anerend/render/passes/IrradianceProbe{Translation,RayTracing,Convolve}Pass.cpp).
 * The reference records Vulkan commands (vkCmdCopyImage / vkCmdTraceRaysKHR /
 * vkCmdDispatch); B200 has no RT pipeline, so each pass exe calls one entry point
 * below instead.  INTEGRATION.md shows the pass-side binding.
 *
 * Conventions: every function returns 0 on success or a negative ANEREND_DDGI_E_*
 * code (never throws, no global state besides the opaque handle);
 * anerend_cuda_ddgi_last_error(h) returns a human-readable message for the last
 * failure on that handle.  A NULL handle is ANEREND_DDGI_E_INVALID for every entry
 * point (_destroy(NULL) is a no-op returning 0, _last_error(NULL) returns "null
 * handle").  A handle is not thread-safe; all device work of a
 * handle is issued on one CUDA stream in call order and is asynchronous with
 * respect to the host until _sync or a _download_* call.  Host pointers passed in
 * are only read during the call.  There is NO CPU fallback: every entry point that
 * computes needs a CUDA device and fails with ANEREND_DDGI_E_CUDA otherwise.
 */
#ifndef ANEREND_CUDA_DDGI_H
#define ANEREND_CUDA_DDGI_H

#include <stddef.h>
#include <stdint.h>
#include "anerend_ddgi_layouts.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ANEREND_DDGI_OK            0
#define ANEREND_DDGI_E_INVALID    -1   /* bad argument / unsupported parameter   */
#define ANEREND_DDGI_E_CUDA       -2   /* CUDA runtime error / no device         */
#define ANEREND_DDGI_E_STATE      -3   /* call order (e.g. trace before build)   */
#define ANEREND_DDGI_E_UNSUPPORTED -4  /* scene feature outside the path's scope */

/* params.flags */
#define ANEREND_DDGI_FLAG_POINT_LIGHTS  (1u << 0) /* extension E1: shade enabled point lights, each gated by an any-hit shadow ray */
#define ANEREND_DDGI_FLAG_RAY_RECORDS   (1u << 1) /* keep a 32-byte fp32 debug record per ray (parity tests)                       */
#define ANEREND_DDGI_FLAG_TEXEL_RECORDS (1u << 2) /* keep the fp32 pre-quantisation value of every convolved texel (parity tests)  */
#define ANEREND_DDGI_FLAG_COUNTERS      (1u << 3) /* count BVH nodes visited / triangles tested (roofline bytes); slower kernels    */
#define ANEREND_DDGI_FLAG_SIMPLE_KERNELS (1u << 4) /* one-thread-per-ray reference kernel instead of the wavefront kernels (debug)    */
#define ANEREND_DDGI_FLAG_SKIP_DEAD_SHADOW_RAYS (1u << 6) /* do not trace the sun shadow ray (rchit:45-66) where dot(N, L) <= 0:
                                                     calcLight multiplies by max(N.L, 0) = 0 there (pbr_light.glsl:161), so the
                                                     radiance is unchanged; lit_mask bit 0 reads 0 and shadow_rays counts traced rays */
#define ANEREND_DDGI_FLAG_NO_LIGHT_MAPS (1u << 7) /* trace every point-light shadow ray: do not build the conservative distance maps
                                                     around the lights that answer "certainly unoccluded" rays without a traversal
                                                     (same lit masks either way; A/B measurements and tests)                        */
#define ANEREND_DDGI_FLAG_TWO_LEVEL     (1u << 8) /* acceleration structure as the reference builds it (tlas_update.comp:55-98,
                                                     VulkanRenderer.cpp:1267-1417): one hierarchy per MESH in object space, shared by
                                                     its instances, under a hierarchy over the instances that _refit_accel rebuilds —
                                                     instead of one hierarchy over a world-space copy of every instance's triangles.
                                                     Results are bit-identical to the one-level structure (the ray/triangle test still
                                                     runs in world space on the same operands); memory and rebuild cost follow the
                                                     number of distinct meshes and of instances.  Light maps are not built.          */
#define ANEREND_DDGI_FLAG_FAST_MATH     (1u << 5) /* radiance + convolution with contracted / approximate fp32 (<= 1e-4 relative per
                                                     update instead of bit-identical to the oracle's defined arithmetic); hit ids,
                                                     shadow-ray results and hit positions are unaffected                            */

/*
 * Parameters that are compile-time constants in the reference
 * (assets/shaders/probe_helpers.glsl:1-7, probe_conv.comp:85, irradiance_probe_update.rgen:91-92).
 */
typedef struct anerend_ddgi_params {
  uint32_t struct_size;       /* = sizeof(anerend_ddgi_params_t)                            */
  int32_t  num_probes_xz;     /* NUM_PROBES_PER_PLANE            (reference 64)              */
  int32_t  num_probes_y;      /* NUM_PROBE_PLANES                (reference 8)               */
  int32_t  rays_per_probe;    /* NUM_RAYS_PER_PROBE, multiple of 32 in [32,1024] (ref. 256)  */
  float    probe_step[3];     /* PROBE_STEP                      (reference 1,2,1)           */
  float    t_min;             /* ray tMin                        (reference 0.1)             */
  float    t_max;             /* ray tMax                        (reference 50)              */
  float    hysteresis;        /* probe_conv.comp:85              (reference 0.96)            */
  uint32_t flags;             /* ANEREND_DDGI_FLAG_*                                          */
  int32_t  slab_begin;        /* first Y-layer this handle updates in a full-grid update     */
  int32_t  slab_end;          /* one past the last; (0, num_probes_y) = whole grid           */
  int32_t  device;            /* CUDA device ordinal, -1 = current device                    */
} anerend_ddgi_params_t;

/* 32-byte per-ray debug record (ANEREND_DDGI_FLAG_RAY_RECORDS). */
typedef struct anerend_ddgi_ray_record {
  uint32_t prim;              /* flattened world triangle id, 0xFFFFFFFF = miss              */
  float    t, u, v;           /* hit distance and barycentrics (rchit `bary`)                */
  float    radiance[3];       /* fp32 radiance before the rgba16f store (rgen:109)           */
  uint32_t lit_mask;          /* bit0 sun unoccluded; bit(1+i) point light i unoccluded (E1) */
} anerend_ddgi_ray_record_t;

/* Device-time breakdown of the most recent call, named like the reference's per-pass GPU timers
 * (FrameGraphBuilder.cpp:276-278; group "IrradianceProbe"). */
typedef struct anerend_ddgi_timings {
  float trans_ms;             /* "IrradianceProbeTrans"    */
  float rt_ms;                /* "IrradianceProbeRT"       */
  float convolve_ms;          /* "IrradianceProbeConvolve" */
  float accel_ms;             /* last _build_accel/_refit_accel ("TLASUpdate"+"BLASUpdate")   */
  float total_ms;             /* first to last event of the last call                         */
  uint32_t kernel_launches;   /* kernels launched by the last call                            */
  float rt_kernel_ms[4];      /* inside "IrradianceProbeRT": closest-hit trace, shade_prep, any-hit trace, shade */
} anerend_ddgi_timings_t;

/* Traversal statistics (ANEREND_DDGI_FLAG_COUNTERS); totals since the last _reset_stats. */
typedef struct anerend_ddgi_stats {
  uint64_t primary_rays;
  uint64_t primary_hits;
  uint64_t shadow_rays;            /* any-hit rays actually traced */
  uint64_t primary_nodes_visited;  /* BVH8 nodes fetched by closest-hit traversal */
  uint64_t primary_tris_tested;
  uint64_t shadow_nodes_visited;
  uint64_t shadow_tris_tested;
  uint64_t shadow_rays_prefiltered; /* shadow rays answered by the light maps instead of a traversal (not in shadow_rays) */
  uint64_t shadow_rays_candidate;   /* of shadow_rays: ended by the exact test against the light map's candidate triangle */
  uint64_t primary_level_changes;   /* two-level structure: instance entries + returns of the closest-hit rays (wavefront kernels) */
  uint64_t shadow_level_changes;
} anerend_ddgi_stats_t;

typedef struct anerend_ddgi_bvh_info {
  uint32_t num_triangles;     /* flattened world triangles                                   */
  uint32_t num_nodes;         /* BVH8 nodes (two-level structure: instance level + all meshes) */
  uint32_t node_bytes;        /* bytes per compressed node (80)                              */
  uint32_t tri_bytes;         /* bytes per traversal triangle record (48)                    */
  uint32_t num_instances;     /* (renderable, mesh) pairs that passed tlas_update.comp's test */
  float    bounds_min[3];
  float    bounds_max[3];
} anerend_ddgi_bvh_info_t;

/* What a two-level structure (ANEREND_DDGI_FLAG_TWO_LEVEL) holds; zeros on a one-level handle. */
typedef struct anerend_ddgi_two_level_info {
  uint32_t num_meshes;        /* distinct meshes = bottom-level hierarchies (VulkanRenderer.cpp:1267-1417: one BLAS per mesh) */
  uint32_t stored_triangles;  /* triangles kept, each mesh once (bvh_info.num_triangles counts every instance's)               */
  uint32_t blas_nodes;        /* BVH8 nodes of all bottom-level hierarchies                                                     */
  uint32_t tlas_nodes;        /* BVH8 nodes of the instance level (tlas_update.comp:55-98 + UpdateTLASPass.cpp:198-228)        */
  uint32_t tlas_depth, blas_depth_max;
  uint64_t accel_bytes;       /* nodes + leaf records + shading records + instance transforms                                   */
  float    tlas_build_ms;     /* device time of the last rebuild of the instance level                                          */
  uint32_t _pad;
} anerend_ddgi_two_level_info_t;

typedef struct anerend_ddgi anerend_ddgi_t;

/* ---- lifetime -------------------------------------------------------------------------- */

/* Allocates the atlas "ProbeRaysConvTex" ((8+2)*nXZ x (8+2)*nXZ*nY RGBA16F, cleared to (0,0,0,1) like
 * clearConvTex, IrradianceProbeConvolvePass.cpp:9-47), "TempProbeTex" and the ray buffers that replace
 * "ProbeRaysIrrTex"/"ProbeRaysDirTex" (IrradianceProbeRayTracingPass.cpp:42-74). */
int anerend_cuda_ddgi_create(const anerend_ddgi_params_t* params, anerend_ddgi_t** out);
int anerend_cuda_ddgi_destroy(anerend_ddgi_t* h);
/* Mirrors re-registration on swap-chain recreation (VulkanRenderer.cpp:1962-2000): atlas re-cleared,
 * _currentLayer = 0 (IrradianceProbeRayTracingPass.cpp:39), _lastProbeIdx = (0,0). */
int anerend_cuda_ddgi_reset(anerend_ddgi_t* h);
const char* anerend_cuda_ddgi_last_error(const anerend_ddgi_t* h);
int anerend_cuda_ddgi_sync(anerend_ddgi_t* h);

/* ---- scene (bindless set 0: bindings 2,3,6,7,8,9,10,11; bindless.glsl:123-179) ---------- */

/* mesh_num_indices carries InternalMesh::_numIndices (internal/InternalMesh.h:12), which GPUMeshInfo lacks
 * (primitiveCount = _numIndices/3, VulkanRenderer.cpp:1313).  Materials may refer to bindless textures
 * (bindless_indices[0,1,3]; [2], the normal map, is never sampled on this path): upload the array with
 * _upload_textures FIRST — an index outside it is E_INVALID.  The call validates everything before it replaces
 * the previous scene; a rejected upload leaves the previous scene and its texture state as they were. */
int anerend_cuda_ddgi_upload_scene(anerend_ddgi_t* h,
    const anerend_vertex_t* vertices, uint32_t num_vertices,
    const uint32_t* indices, uint32_t num_indices,
    const anerend_mesh_info_t* meshes, const uint32_t* mesh_num_indices, uint32_t num_meshes,
    const anerend_material_info_t* materials, uint32_t num_materials,
    const anerend_renderable_t* renderables, uint32_t num_renderables,
    const uint32_t* model_indices, uint32_t num_model_indices,      /* binding 8: rendModelBuffer    */
    const uint32_t* material_indices, uint32_t num_material_indices,/* binding 7: rendMatIndexBuffer */
    const anerend_light_t* lights, uint32_t num_lights);

/* Bindless textures (set 0 binding 16): the whole array, index = bindless index.  Call before _upload_scene; materials may
 * then reference them through bindless_indices {metallicRoughness, albedo, normal, emissive} (the normal map is never sampled
 * on this path: rchit:77 passes a zero tangent).  Sampling: LOD 0, bilinear, REPEAT / CLAMP_TO_EDGE per texture.
 * count = 0 removes all textures. */
int anerend_cuda_ddgi_upload_textures(anerend_ddgi_t* h, const anerend_texture_desc_t* textures, uint32_t count);
/* tileInfoBuffer (set 0 binding 14, GPUTileInfo, GpuBuffers.h:74-76): one entry per tile of the 21 x 21 page around the camera
 * (bindless.glsl:4-5, 291-308); ddgi_atlas_tex = bindless index of the tile's baked atlas (ANEREND_TEX_RGBA16F, what
 * _download_atlas returned after baking that tile) or -1.  With it, sampleProbe places the grid on the baked tile where there
 * is one (pbr_light.glsl:251-273) and _indirect_diffuse(clamp_to_edge = 1) samples the baked atlas like the deferred pass
 * (deferred_tiled.comp:264-296).  count = 0 removes it. */
int anerend_cuda_ddgi_upload_tile_infos(anerend_ddgi_t* h, const anerend_tile_info_t* tiles, uint32_t count);

/* Binding 0 (scene_ubo.glsl:14-37).  Fields read: cameraPos, lightDir, skyColor, sunIntensity,
 * skyIntensity, flags (DDGI, DDGI_MULTI, RT_ON, BAKE_MODE_ON). */
int anerend_cuda_ddgi_set_scene_ubo(anerend_ddgi_t* h, const anerend_scene_ubo_t* ubo);

/* Replaces tlas_update.comp + the TLAS/BLAS builds (UpdateTLASPass.cpp:128-248, VulkanRenderer.cpp:1267-1417):
 * instance selection (visible bit set, cull bit clear), flatten to world-space triangles, LBVH, BVH8. */
int anerend_cuda_ddgi_build_accel(anerend_ddgi_t* h);
/* Deforming geometry (skinned meshes; UpdateBlasPass.cpp:60-175, VulkanRenderer.cpp:967-1143): overwrite `count` vertices of
 * the giga vertex buffer starting at `first` (host pointer, or a device pointer for a buffer the skinning pass wrote), then call
 * _refit_accel: the hierarchy's topology is kept, triangles, boxes and shading records follow the new vertices. */
int anerend_cuda_ddgi_update_vertices(anerend_ddgi_t* h, const anerend_vertex_t* vertices, uint32_t first, uint32_t count);
int anerend_cuda_ddgi_update_vertices_device(anerend_ddgi_t* h, const void* d_vertices, uint32_t first, uint32_t count);
/* Moving nodes: new renderable buffer (same count, same visibility) -> re-flatten + bottom-up refit.
 * renderables may be NULL to refit with the current buffer.  `_device` takes a device pointer (the
 * reference's RenderableBuffer is GPU-written earlier in the frame, cull.comp:303). */
int anerend_cuda_ddgi_refit_accel(anerend_ddgi_t* h, const anerend_renderable_t* renderables, uint32_t num_renderables);
int anerend_cuda_ddgi_refit_accel_device(anerend_ddgi_t* h, const void* d_renderables, uint32_t num_renderables);
/* The cull bit the reference's culling pass sets in front of the chain (cull.comp:241-244,296-303): every visible renderable whose
 * bounding-sphere centre (bounds.xyz + the transform's translation) is farther than `range` metres from ubo.cameraPos gets bit 1 of
 * `visible` set, the others get it cleared; tlas_update.comp (and _build_accel) leave culled renderables out.  For hosts that do not
 * run the Vulkan cull pass; the reference uses range = 50.  *changed = number of renderables whose bit flipped; if it is not 0 the
 * acceleration structure is invalid until the next _build_accel. */
int anerend_cuda_ddgi_apply_camera_cull(anerend_ddgi_t* h, float range, uint32_t* changed);
/* Timing events behind _get_timings: level 0 none, 1 (default) around the passes — what FrameGraphBuilder's per-pass GPU timers
 * measure (FrameGraphBuilder.cpp:276-278) — 2 also around every kernel of the RT pass (rt_kernel_ms).  Events between launches cost
 * time (C3: 26 us per update at level 2, and a read-back of the atlas in flight delays them further), so production runs at 0 or 1. */
int anerend_cuda_ddgi_set_timers(anerend_ddgi_t* h, int32_t level);
/* Pipelined updates (off by default).  Ray generation and the closest-hit traversal of an update read the rotation and the
 * acceleration structure, never the atlas (only the hit shading samples it, rchit:98-107), so with on = 1 they run on a stream of
 * their own as soon as _update / _trace_layer / _frame is called — under the tail of the previous update's kernels — and the rest
 * of the update waits for them on the device.  Results are unchanged.  Active only with _set_timers(h, 0) and without counters;
 * builds and refits are waited for. */
int anerend_cuda_ddgi_set_pipeline(anerend_ddgi_t* h, int32_t on);
int anerend_cuda_ddgi_get_bvh_info(anerend_ddgi_t* h, anerend_ddgi_bvh_info_t* out);
int anerend_cuda_ddgi_get_two_level_info(anerend_ddgi_t* h, anerend_ddgi_two_level_info_t* out);
/* The instances _build_accel selected, in the order they were emitted: what tlas_update.comp:54-98 writes to its InstanceBuffer
 * when its invocations run in renderable order (custom index = material | mesh << 10, tlas_update.comp:88-91), plus the range of
 * flattened triangle ids the instance owns.  count must equal bvh_info.num_instances.  Inspection only (tests/test_golden_tlas.py). */
typedef struct anerend_ddgi_instance {
  uint32_t renderable;        /* index into the renderable buffer                            */
  uint32_t custom_index;      /* 24-bit instanceCustomIndex: material (10 bits) | mesh << 10 */
  uint32_t first_triangle;    /* first flattened world triangle id                           */
  uint32_t num_triangles;
} anerend_ddgi_instance_t;
int anerend_cuda_ddgi_download_instances(anerend_ddgi_t* h, anerend_ddgi_instance_t* out, size_t count);
/* Debug/validation: copies the BVH8 nodes (num_nodes*80 B) and traversal triangles (num_triangles*48 B). */
int anerend_cuda_ddgi_download_bvh(anerend_ddgi_t* h, void* nodes, size_t nodes_bytes, void* tris, size_t tris_bytes);

/* ---- the three passes ------------------------------------------------------------------- */

/* "IrradianceProbeTrans" exe (IrradianceProbeTranslationPass.cpp:56-186).  Reads cameraPos from the UBO,
 * keeps _lastProbeIdx in the handle.  No-op when DDGI is off or bake mode is on. */
int anerend_cuda_ddgi_translate(anerend_ddgi_t* h);
/* "IrradianceProbeRT" exe for one layer (IrradianceProbeRayTracingPass.cpp:95-165 + rgen/rchit/rmiss):
 * rot = the push-constant `randomRot` (column-major mat4).  Reads the current atlas (multi-bounce). */
int anerend_cuda_ddgi_trace_layer(anerend_ddgi_t* h, int32_t layer, const float rot[16]);
/* "IrradianceProbeConvolve" exe for one layer (IrradianceProbeConvolvePass.cpp:109-143 + probe_conv.comp). */
int anerend_cuda_ddgi_convolve_layer(anerend_ddgi_t* h, int32_t layer);
/* layer >= 0: trace_layer + convolve_layer (one reference execution).
 * layer  < 0: full-grid update of layers [slab_begin, slab_end): every layer is traced against the SAME
 *             previous atlas, then all are convolved ("jacobi" mode); rots = num_probes_y matrices,
 *             indexed by absolute layer. */
int anerend_cuda_ddgi_update(anerend_ddgi_t* h, int32_t layer, const float* rots);
/* Jacobi update of an arbitrary set of layers (each at most once): the multi-GPU path gives every rank an interleaved
 * set so the work balances (the reference's own time slicing is one layer per execution,
 * IrradianceProbeRayTracingPass.cpp:112,160-164).  rots = num_probes_y matrices indexed by absolute layer. */
int anerend_cuda_ddgi_update_layers(anerend_ddgi_t* h, const int32_t* layers, uint32_t num_layers, const float* rots);
/* One frame at the reference's cadence: translate, trace+convolve layer `_currentLayer`, advance it
 * (IrradianceProbeRayTracingPass.cpp:112,158-164).  *layer_out receives blackboard "ProbeLayer". */
int anerend_cuda_ddgi_frame(anerend_ddgi_t* h, const float rot[16], int32_t* layer_out);

/* ---- results / state -------------------------------------------------------------------- */

int anerend_cuda_ddgi_atlas_dims(const anerend_ddgi_t* h, int32_t* width, int32_t* height);
/* "ProbeRaysConvTex" as row-major RGBA16F, width*height*8 bytes (mirrors downloadDDGIAtlas,
 * VulkanRenderer.cpp:2706-2771). */
int anerend_cuda_ddgi_download_atlas(anerend_ddgi_t* h, void* rgba16f, size_t bytes);
/* ---- atlas consumer ---------------------------------------------------------
 * calcIndirectDiffuseLight (pbr_light.glsl:365-388) for `count` surface points against the current atlas: what
 * deferred_tiled.comp:264-296 evaluates per G-buffer pixel (CLAMP_TO_EDGE sampler, DeferredLightingRenderPass.cpp:192-200;
 * clamp_to_edge = 1) and mirror_reflections.rchit:92-118 per reflection hit (default REPEAT sampler, SpecularGIRTPass.cpp:66;
 * clamp_to_edge = 0).  The probe grid is centred on the camera of the last _set_scene_ubo, like the shader's.
 * points: 12 floats each = pos.xyz, normal.xyz, albedo.rgb, metallic, roughness, unused.  out: 4 floats each (rgb, 0).
 * The _device variant takes device pointers (16-byte aligned) and runs asynchronously on the handle's stream. */
int anerend_cuda_ddgi_indirect_diffuse(anerend_ddgi_t* h, const float* points, uint32_t count, const float view_pos[3],
                                       int32_t clamp_to_edge, float* out_rgba);
int anerend_cuda_ddgi_indirect_diffuse_device(anerend_ddgi_t* h, const void* d_points, uint32_t count, const float view_pos[3],
                                              int32_t clamp_to_edge, void* d_out_rgba);

/* ---- ShadowRT on the same acceleration structure ---------------------------
 * The frame-graph pass "ShadowRT" (ShadowRayTracingPass.cpp:73-81, shadow.rgen:30-82, shadow.rmiss): for every pixel of a
 * width x height depth buffer (float, row-major, the launch size), world position = invViewProj * (uv*2-1, depth, 1) with
 * uv = pixelToUv(pixel, (screenWidth, screenHeight)), one any-hit ray along -lightDir (not normalised), t in (0.1, 50).
 * out: r32f image, 1 = lit (or sky: depth == 1), 0 = shadowed.  Nothing is written unless UBO_RT_SHADOWS_FLAG is set.
 * Reads invViewProj, lightDir, screenWidth/Height and flags from the last _set_scene_ubo. */
int anerend_cuda_ddgi_shadow_rt(anerend_ddgi_t* h, const float* depth, uint32_t width, uint32_t height, float* out_r32f);
int anerend_cuda_ddgi_shadow_rt_device(anerend_ddgi_t* h, const void* d_depth, uint32_t width, uint32_t height, void* d_out_r32f);

/* The frame-graph pass "SpecularGIRT" (SpecularGIRTPass.cpp:100-127, mirror_reflections.rgen/.rchit/.rmiss): per pixel, the view
 * vector reflected about the G-buffer normal (geom0: rgba32f, normal in rgb), closest hit on the path's BVH, the hit shaded like
 * the probe hit shader (sun shadow ray, calcLight, calcIndirectDiffuseLight from the current atlas when UBO_DDGI_FLAG is set,
 * emissive).  out: rgba16f (irradiance.rgb, hit distance); (sky, 1e6) on a miss; 0 for sky pixels (depth == 1).  Nothing is
 * written unless UBO_RT_ON_FLAG and UBO_SPECULAR_GI_FLAG are set.  Reads invViewProj, cameraPos, screenWidth/Height, lightDir,
 * sun/sky and flags from the last _set_scene_ubo. */
int anerend_cuda_ddgi_specular_gi(anerend_ddgi_t* h, const float* depth, const float* geom0_rgba32f, uint32_t width, uint32_t height,
                                  void* out_rgba16f);
int anerend_cuda_ddgi_specular_gi_device(anerend_ddgi_t* h, const void* d_depth, const void* d_geom0_rgba32f, uint32_t width,
                                         uint32_t height, void* d_out_rgba16f);

/* Row stripe (sharding inside the layers): the handle updates only the probe rows z = offset + k*stride of
 * every layer it launches; (0, 1) = all rows (default).  With R ranks, rank r takes (r, R): every rank carries the same
 * share of every layer, so the ranks finish together whatever the scene looks like.  _download_atlas_stripe returns
 * this handle's rows of all layers, packed (layer, row): num_probes_y * rows * 10 * width * 8 bytes. */
int anerend_cuda_ddgi_set_row_stripe(anerend_ddgi_t* h, int32_t offset, int32_t stride);
/* The contiguous variant: the probe rows z in [first, first + count) of every layer (a spatially compact share; _download_atlas_stripe
 * and the peer exchange follow it like a stripe). */
int anerend_cuda_ddgi_set_row_block(anerend_ddgi_t* h, int32_t first, int32_t count);
int anerend_cuda_ddgi_download_atlas_stripe(anerend_ddgi_t* h, void* rgba16f, size_t bytes);
/* The same stripe on the read-back stream (see _download_atlas_async; _download_wait blocks until it has landed).  Allowed while a
 * peer exchange is connected: a handle's own rows are written by its own convolution only, and that convolution waits for the copy. */
int anerend_cuda_ddgi_download_atlas_stripe_async(anerend_ddgi_t* h, void* rgba16f, size_t bytes);
/* The atlas rows of `num_layers` Y-layers (10*nXZ rows each), packed in the order given: what one rank of a sharded
 * update reads back (its own layers); bytes = num_layers * (10*nXZ)^2 * 8. */
int anerend_cuda_ddgi_download_atlas_layers(anerend_ddgi_t* h, const int32_t* layers, uint32_t num_layers, void* rgba16f, size_t bytes);
/* Asynchronous read-back: returns at once; the copy starts when the work queued so far is done and overlaps the next update's
 * tracing (the next writer of the atlas waits for it on the device).  `rgba16f` should be page-locked.  _download_wait blocks
 * until the bytes have landed.  Not available while peers write this atlas (connected exchange, _set_peer_atlases): read the own
 * stripe with _download_atlas_stripe_async instead. */
int anerend_cuda_ddgi_download_atlas_async(anerend_ddgi_t* h, void* rgba16f, size_t bytes);
int anerend_cuda_ddgi_download_wait(anerend_ddgi_t* h);
int anerend_cuda_ddgi_upload_atlas(anerend_ddgi_t* h, const void* rgba16f, size_t bytes);
/* Device pointer of the atlas for consumers on the same device (deferred lighting, NCCL all-gather of slabs).
 * Rows [10*nXZ*slab_begin, 10*nXZ*slab_end) are the rows this handle writes. */
int anerend_cuda_ddgi_atlas_device_ptr(anerend_ddgi_t* h, void** d_ptr, size_t* bytes);
/* Publishes the atlas into device memory the ENGINE owns — the frame-graph image "ProbeRaysConvTex" its consumers sample
 * (DeferredLightingRenderPass.cpp:192-200, DebugViewRenderPass.cpp:225), imported into CUDA as linear RGBA16F rows of `pitch_bytes`
 * (>= width * 8): a device-to-device pitched copy ordered after the work queued so far on the handle's stream.  The pass adapter
 * (anerend_ddgi_passes.hpp) calls it after the convolution, after a translation and after (re-)registration, where the reference
 * clears the image (clearConvTex, IrradianceProbeConvolvePass.cpp:9-47). */
int anerend_cuda_ddgi_export_atlas_device(anerend_ddgi_t* h, void* d_dst, size_t pitch_bytes);
/* Fused atlas all-gather: device pointers to the peers' atlases (cudaIpc-opened or P2P-enabled).  When set,
 * the convolution stores every finished texel into all peers as well, replacing the separate collective. */
int anerend_cuda_ddgi_set_peer_atlases(anerend_ddgi_t* h, void* const* d_peers, uint32_t num_peers);
/* Fused atlas exchange between the ranks of one box (one process per GPU), replacing the NCCL all-gather of the
 * slabs: every rank exports a blob (cudaIpc handles of its atlas and of its barrier flags), the
 * blobs of all ranks — concatenated in rank order by whatever plumbing the host has (torch.distributed here) — are
 * handed to _exchange_connect, which maps the peers' atlases over NVLink.  From then on _update/_update_layers run
 *   trace -> device barrier (all ranks have read the previous atlas) -> convolution storing every finished texel into
 *   the local atlas AND all peers' -> device barrier (all texels have landed),
 * so the full atlas is assembled inside the convolution kernel, with no separate collective.  A barrier that waits
 * longer than 30 s sets a sticky error reported by _sync.  _exchange_disconnect (all ranks, before any _destroy)
 * unmaps the peers. */
#define ANEREND_DDGI_EXCHANGE_BLOB_BYTES 128
int anerend_cuda_ddgi_exchange_export(anerend_ddgi_t* h, void* blob, size_t bytes);
int anerend_cuda_ddgi_exchange_connect(anerend_ddgi_t* h, const void* blobs, uint32_t world, uint32_t rank);
/* With a connected exchange the push of an update's rows into the peers runs on a second stream, under the next update's tracing (only
 * the hit shading samples the atlas, and waits for it).  _exchange_fence makes the handle's stream wait for the rows of the last update:
 * call it before reading the atlas on that stream from outside the library (the library's own readers do), or to time a sequence of
 * updates to its true end.  _sync waits for both streams. */
int anerend_cuda_ddgi_exchange_fence(anerend_ddgi_t* h);
int anerend_cuda_ddgi_exchange_disconnect(anerend_ddgi_t* h);
/* CUDA stream (cudaStream_t) all work of this handle is issued on. */
int anerend_cuda_ddgi_stream(anerend_ddgi_t* h, void** stream);

/* "ProbeRaysIrrTex"/"ProbeRaysDirTex" of the last trace of `layer`, in the reference's image layout:
 * (S_x*nXZ) x (S_y*nXZ) RGBA16F, pixel (probeX*S_x+px, probeZ*S_y+py), rayIndex = py*S_x+px  (rgen:51-58). */
int anerend_cuda_ddgi_ray_tile(const anerend_ddgi_t* h, int32_t* sx, int32_t* sy);
int anerend_cuda_ddgi_download_rays(anerend_ddgi_t* h, int32_t layer, void* irr_rgba16f, void* dir_rgba16f, size_t bytes_each);
/* "IrradianceProbeConvolve" on its own (BASELINE config 1, probe_conv.comp over caller-supplied radiance): ProbeRaysIrrTex of
 * `layer` in the image layout above (host pointer); ProbeRaysDirTex becomes what "IrradianceProbeRT" writes for randomRot = rot
 * (mat3(rot) * Fibonacci direction, rgen:87-88,108).  Follow with _convolve_layer(h, layer).  Not available with a row stripe. */
int anerend_cuda_ddgi_upload_ray_radiance(anerend_ddgi_t* h, int32_t layer, const void* irr_rgba16f, size_t bytes, const float rot[16]);
/* nXZ*nXZ*rays records, index (probeZ*nXZ+probeX)*rays + rayIndex.  Needs FLAG_RAY_RECORDS. */
int anerend_cuda_ddgi_download_ray_records(anerend_ddgi_t* h, int32_t layer, anerend_ddgi_ray_record_t* out, size_t count);
/* fp32 RGBA per atlas texel (same addressing as the atlas) of the last convolution.  Needs FLAG_TEXEL_RECORDS. */
int anerend_cuda_ddgi_download_texel_records(anerend_ddgi_t* h, float* rgba32f, size_t bytes);

int anerend_cuda_ddgi_get_timings(anerend_ddgi_t* h, anerend_ddgi_timings_t* out);
int anerend_cuda_ddgi_get_stats(anerend_ddgi_t* h, anerend_ddgi_stats_t* out);
int anerend_cuda_ddgi_reset_stats(anerend_ddgi_t* h);

/* ---- host helpers shared by pass code --------------------------------------------------- */

/* randomRot = glm::rotate(mat4(1), radians(angle_deg), axis∈{0:x,1:y,2:z})
 * (IrradianceProbeRayTracingPass.cpp:115-133; glm/ext/matrix_transform.inl:18-46). Column-major. */
int anerend_cuda_ddgi_make_rotation(int32_t axis, float angle_deg, float out[16]);
/* worldDirFromFibonacciIndex for i in [0,n) (rgen:35-46); out = n*3 floats.  The same host table is what the
 * kernels consume, so directions do not depend on device sin/cos range reduction. */
int anerend_cuda_ddgi_fibonacci_dirs(int32_t n, float* out);
const char* anerend_cuda_ddgi_version(void);

#ifdef __cplusplus
}
#endif
#endif
