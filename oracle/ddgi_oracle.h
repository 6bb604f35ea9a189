/*
 * ddgi_oracle.h — C API of the CPU oracle for the DDGI probe-update path.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a CPU restatement (multithreaded C++17) of the reference's
 * shaders and pass code; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (anerend_b200/) never links, imports or calls anything in oracle/.
 *
 * PINNING: the reference holds no golden vectors, tests or CPU path for this chain and cannot be built here.
 * The oracle is held to the reference's own COMPILED shaders instead: assets/probe_conv_comp.spv, irradiance_probe_update_rgen /
 * _rchit / _rmiss.spv and tlas_update_comp.spv executed by tests/golden/spirv_interp.py (vectors + generators in tests/golden/,
 * tests/test_golden_*.py).  PARITY UNPINNED for what those binaries leave to the Vulkan driver: the ray/triangle intersection and
 * traversal (defined in DESIGN.md, checked by known-answer tests in tests/test_oracle_kat.py), filtering weights, fp16 rounding.
 *
 * The API mirrors include/anerend_cuda_ddgi.h one to one (same parameter struct, same buffer formats), so a
 * parity test drives both sides with identical calls.
 */
#ifndef DDGI_ORACLE_H
#define DDGI_ORACLE_H

#include "../include/anerend_cuda_ddgi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ddgi_oracle ddgi_oracle_t;

int ddgi_oracle_create(const anerend_ddgi_params_t* params, ddgi_oracle_t** out);
int ddgi_oracle_destroy(ddgi_oracle_t* h);
int ddgi_oracle_reset(ddgi_oracle_t* h);
const char* ddgi_oracle_last_error(const ddgi_oracle_t* h);
/* 0 = hardware_concurrency(); returns the thread count in use. */
int ddgi_oracle_set_threads(ddgi_oracle_t* h, int32_t n);
/* 1 = brute-force ray/triangle loops instead of the oracle's own BVH (validates that BVH). */
int ddgi_oracle_set_brute_force(ddgi_oracle_t* h, int32_t on);

/* PROTOTYPE: answer the sun's shadow rays from a resolution x resolution light-space grid of triangle lists instead of
 * the hierarchy (0 = off).  Same exact triangle test, so results must not change: tests/test_sun_grid.py. */
int ddgi_oracle_set_sun_grid(ddgi_oracle_t* h, int32_t resolution);

int ddgi_oracle_upload_scene(ddgi_oracle_t* h,
    const anerend_vertex_t* vertices, uint32_t num_vertices,
    const uint32_t* indices, uint32_t num_indices,
    const anerend_mesh_info_t* meshes, const uint32_t* mesh_num_indices, uint32_t num_meshes,
    const anerend_material_info_t* materials, uint32_t num_materials,
    const anerend_renderable_t* renderables, uint32_t num_renderables,
    const uint32_t* model_indices, uint32_t num_model_indices,
    const uint32_t* material_indices, uint32_t num_material_indices,
    const anerend_light_t* lights, uint32_t num_lights);
int ddgi_oracle_update_vertices(ddgi_oracle_t* h, const anerend_vertex_t* vertices, uint32_t first, uint32_t count);
int ddgi_oracle_upload_tile_infos(ddgi_oracle_t* h, const anerend_tile_info_t* tiles, uint32_t count);
int ddgi_oracle_sample_texture(ddgi_oracle_t* h, int32_t index, float u, float v, float out_rgba[4]);
int ddgi_oracle_upload_textures(ddgi_oracle_t* h, const anerend_texture_desc_t* textures, uint32_t count);
int ddgi_oracle_set_scene_ubo(ddgi_oracle_t* h, const anerend_scene_ubo_t* ubo);
int ddgi_oracle_build_accel(ddgi_oracle_t* h);
int ddgi_oracle_refit_accel(ddgi_oracle_t* h, const anerend_renderable_t* renderables, uint32_t num_renderables);
int ddgi_oracle_apply_camera_cull(ddgi_oracle_t* h, float range, uint32_t* changed);
int ddgi_oracle_get_bvh_info(ddgi_oracle_t* h, anerend_ddgi_bvh_info_t* out);
int ddgi_oracle_download_instances(ddgi_oracle_t* h, anerend_ddgi_instance_t* out, size_t count);
/* World-space flattened triangles, 9 floats each (v0, v1, v2), in global triangle-id order. */
int ddgi_oracle_download_world_triangles(ddgi_oracle_t* h, float* out, size_t count);

int ddgi_oracle_translate(ddgi_oracle_t* h);
int ddgi_oracle_trace_layer(ddgi_oracle_t* h, int32_t layer, const float rot[16]);
int ddgi_oracle_convolve_layer(ddgi_oracle_t* h, int32_t layer);
int ddgi_oracle_update(ddgi_oracle_t* h, int32_t layer, const float* rots);
int ddgi_oracle_frame(ddgi_oracle_t* h, const float rot[16], int32_t* layer_out);

int ddgi_oracle_atlas_dims(const ddgi_oracle_t* h, int32_t* width, int32_t* height);
int ddgi_oracle_download_atlas(ddgi_oracle_t* h, void* rgba16f, size_t bytes);
int ddgi_oracle_upload_atlas(ddgi_oracle_t* h, const void* rgba16f, size_t bytes);
int ddgi_oracle_ray_tile(const ddgi_oracle_t* h, int32_t* sx, int32_t* sy);
int ddgi_oracle_download_rays(ddgi_oracle_t* h, int32_t layer, void* irr_rgba16f, void* dir_rgba16f, size_t bytes_each);
int ddgi_oracle_download_ray_records(ddgi_oracle_t* h, int32_t layer, anerend_ddgi_ray_record_t* out, size_t count);
int ddgi_oracle_download_texel_records(ddgi_oracle_t* h, float* rgba32f, size_t bytes);
int ddgi_oracle_get_stats(ddgi_oracle_t* h, anerend_ddgi_stats_t* out);
int ddgi_oracle_reset_stats(ddgi_oracle_t* h);

/* Convolution alone (BASELINE config 1): radiance/direction images supplied by the caller in the reference's
 * image layout (RGBA16F, (S_x*nXZ) x (S_y*nXZ)), then probe_conv.comp over `layer`. */
int ddgi_oracle_set_rays(ddgi_oracle_t* h, int32_t layer, const void* irr_rgba16f, const void* dir_rgba16f, size_t bytes_each);

int ddgi_oracle_upload_ray_radiance(ddgi_oracle_t* h, int32_t layer, const void* irr_rgba16f, size_t bytes, const float rot[16]);

/* ---- single shader functions, exported for the known-answer tests ------------------------ */
/* 1 if anything is hit in (tmin, tmax) along the ray (the shadow-ray query of rchit:48-66 for an arbitrary ray), 0 if not, < 0 on error. */
int  ddgi_oracle_any_hit(ddgi_oracle_t* h, const float origin[3], const float dir[3], float tmin, float tmax);
/* The hit shader for one hit (flattened triangle id, barycentrics) of a ray from `origin`: irradiance_probe_update.rchit
 * (specular = 0) or mirror_reflections.rchit (specular = 1) against the handle's current atlas; out = (irradiance, distance from
 * the origin).  The function trace_layer / specular_gi themselves call. */
int  ddgi_oracle_shade_hit(ddgi_oracle_t* h, uint32_t prim, float u, float v, const float origin[3], int32_t specular, float out[4], uint32_t* lit_mask);
/* The ray shadow.rgen (specular = 0) or mirror_reflections.rgen (specular = 1) traces for screen pixel (px, py) with depth sample
 * `depth` (and G-buffer normal geom_normal): the functions the two passes themselves use. */
int  ddgi_oracle_screen_ray(ddgi_oracle_t* h, int32_t specular, uint32_t px, uint32_t py, float depth, const float geom_normal[3], float origin[3], float dir[3]);
/* What rgen hands to traceRayEXT for launch pixel (probe, ray) of `layer` (origin rgen:51-83, direction rgen:86-87), and what rmiss
 * would return for it (rmiss:17-21): the functions trace_layer itself uses. */
int  ddgi_oracle_probe_ray(ddgi_oracle_t* h, int32_t layer, int32_t probe, int32_t ray, const float rot[16], float origin[3], float dir[3], float miss[3]);
int  ddgi_oracle_make_rotation(int32_t axis, float angle_deg, float out[16]);          /* glm::rotate */
int  ddgi_oracle_fibonacci_dirs(int32_t n, float* out);                                /* rgen:35-46 */
void ddgi_oracle_oct_encode(const float v[3], float out[2]);                           /* octahedron_helpers.glsl:10-20 */
void ddgi_oracle_oct_decode(const float o[2], float out[3]);                           /* octahedron_helpers.glsl:25-34 */
void ddgi_oracle_to_linear(const float srgb[3], float out[3]);                         /* bindless.glsl:240-247 */
uint16_t ddgi_oracle_float_to_half(float f);
float    ddgi_oracle_half_to_float(uint16_t h);
/* calcLight (pbr_light.glsl:98-166). directional=1: sun from light_dir/sun_intensity; 0: `light`. */
void ddgi_oracle_calc_light(const float normal[3], const float albedo[3], float metallic, float roughness,
                            const float world_pos[3], const float view_pos[3], const anerend_light_t* light,
                            int32_t directional, const float light_dir[3], float sun_intensity, float out[3]);
/* calcIndirectDiffuseLight against the handle's current atlas/UBO (pbr_light.glsl:365-388). */
int  ddgi_oracle_specular_gi(ddgi_oracle_t* h, const float* depth, const float* geom0_rgba32f, uint32_t width, uint32_t height, uint16_t* out_rgba16f);
int  ddgi_oracle_shadow_rt(ddgi_oracle_t* h, const float* depth, uint32_t width, uint32_t height, float* out_r32f);
int  ddgi_oracle_indirect_diffuse(ddgi_oracle_t* h, const float* points, uint32_t count, const float view_pos[3], int32_t clamp_to_edge, float* out_rgba);
int  ddgi_oracle_calc_indirect(ddgi_oracle_t* h, const float normal[3], const float albedo[3], float metallic,
                               float roughness, const float world_pos[3], const float view_pos[3], float out[3]);
/* Möller–Trumbore exactly as DESIGN.md specifies; returns 1 on hit with t in (tmin,tmax). */
int  ddgi_oracle_intersect(const float orig[3], const float dir[3], const float v0[3], const float v1[3],
                           const float v2[3], float tmin, float tmax, float* t, float* u, float* v);

#ifdef __cplusplus
}
#endif
#endif
