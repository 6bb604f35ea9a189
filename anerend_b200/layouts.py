"""numpy views of the bindless scene buffers (include/anerend_ddgi_layouts.h).

Byte-for-byte the reference's structs: Vertex.h:7-19, GpuBuffers.h:15-44,69-76,125-148.
"""
import numpy as np

VERTEX = np.dtype([
    ("pos", "<f4", 3), ("color", "<f4", 3), ("normal", "<f4", 3), ("tangent", "<f4", 4),
    ("joint_weights", "<f4", 4), ("uv", "<f4", 2), ("padding", "<f4"),
    ("joint_ids", "<i2", 4), ("padding1", "<f4"), ("padding2", "<f4")])
RENDERABLE = np.dtype([
    ("transform", "<f4", 16), ("bounds", "<f4", 4), ("tint", "<f4", 4),
    ("model_offset", "<u4"), ("num_meshes", "<u4"), ("skeleton_offset", "<u4"), ("visible", "<u4"),
    ("first_material_index", "<u4"), ("dynamic_model_offset", "<u4"), ("terrain_offset", "<i4"), ("_pad", "<u4")])
MATERIAL = np.dtype([
    ("base_col_fac", "<f4", 4), ("emissive", "<f4", 4), ("bindless_indices", "<i4", 4), ("met_rough", "<f4", 4)])
MESH_INFO = np.dtype([
    ("min_pos", "<f4", 4), ("max_pos", "<f4", 4), ("vertex_offset", "<u4"), ("index_offset", "<u4"), ("blas_ref", "<u8")])
LIGHT = np.dtype([("world_pos", "<f4", 4), ("color", "<f4", 4)])
SCENE_UBO = np.dtype([
    ("view", "<f4", 16), ("proj", "<f4", 16), ("inv_view", "<f4", 16), ("inv_proj", "<f4", 16),
    ("inv_view_proj", "<f4", 16), ("directional_shadow_matrix_proj", "<f4", 16),
    ("directional_shadow_matrix_view", "<f4", 16),
    ("camera_pos", "<f4", 4), ("camera_grid_pos", "<i4", 4), ("light_dir", "<f4", 4),
    ("bake_tile_info", "<i4", 4), ("sky_color", "<f4", 4),
    ("time", "<f4"), ("delta", "<f4"), ("screen_width", "<u4"), ("screen_height", "<u4"),
    ("sun_intensity", "<f4"), ("sky_intensity", "<f4"), ("exposure", "<f4"), ("flags", "<u4"),
    ("bloom_thresh", "<f4"), ("bloom_knee", "<f4")])
RAY_RECORD = np.dtype([
    ("prim", "<u4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("radiance", "<f4", 3), ("lit_mask", "<u4")])

assert VERTEX.itemsize == 96 and RENDERABLE.itemsize == 128 and MATERIAL.itemsize == 64
assert MESH_INFO.itemsize == 48 and LIGHT.itemsize == 32 and SCENE_UBO.itemsize == 568 and RAY_RECORD.itemsize == 32

# SceneUboFlags (GpuBuffers.h:108-122)
UBO_DDGI_FLAG = 1 << 4
UBO_DDGI_MULTI_FLAG = 1 << 5
UBO_RT_ON_FLAG = 1 << 10
UBO_BAKE_MODE_ON_FLAG = 1 << 11

# anerend_ddgi_params.flags
FLAG_POINT_LIGHTS = 1 << 0
FLAG_RAY_RECORDS = 1 << 1
FLAG_TEXEL_RECORDS = 1 << 2
FLAG_COUNTERS = 1 << 3
FLAG_SIMPLE_KERNELS = 1 << 4
FLAG_FAST_MATH = 1 << 5
FLAG_SKIP_DEAD_SHADOW_RAYS = 1 << 6
FLAG_NO_LIGHT_MAPS = 1 << 7
FLAG_TWO_LEVEL = 1 << 8

MISS = 0xFFFFFFFF

# anerend_texture_desc_t.format (include/anerend_ddgi_layouts.h)
TEX_RGBA8_UNORM, TEX_RGBA8_SRGB, TEX_RG8_UNORM, TEX_R8_UNORM = 0, 1, 2, 3
UBO_RT_SHADOWS_FLAG = 1 << 3
UBO_SPECULAR_GI_FLAG = 1 << 6
TEX_RGBA16F = 4
