"""Two-level acceleration structure (ANEREND_DDGI_FLAG_TWO_LEVEL; SURVEY.md §8f N3, tlas_update.comp:55-98 +
VulkanRenderer.cpp:1267-1417): one hierarchy per distinct mesh in object space under a hierarchy over the instances.
The oracle knows nothing of it — it flattens every instance to world space and traces that — and the results have to be
the same bit for bit: hit ids (flattened triangle ids), (t, u, v), lit masks, radiance, atlas.  CPU part: the product's
build + per-thread traversal on the host harness; the `-m gpu` tests repeat it with the CUDA kernels (both families).
"""
import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

ALL = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_TEXEL_RECORDS | L.FLAG_COUNTERS

SCENES = {
    "cornell": lambda: (scenes.cornell_box(grid=(8, 4, 64), wall_res=10, box_res=3), (8, 4, 64)),
    "soup11": lambda: (scenes.random_soup(11), (4, 3, 64)),
    "soup12_128rays": lambda: (scenes.random_soup(12, ntris=900, grid=(3, 2, 128)), (3, 2, 128)),
    "quad": lambda: (scenes.quad_scene(second_quad=True), (4, 2, 64)),
    "textured": lambda: (scenes.textured_soup(51), (4, 3, 64)),
    "empty": lambda: (scenes.empty_scene(grid=(4, 2, 64)), (4, 2, 64)),
    "instanced": lambda: (scenes.instanced_field(7, instances=40, grid=(4, 3, 64)), (4, 3, 64)),
}


def check_update(o, e, ny, rots, updates=2):
    for _ in range(updates):
        for h in (o, e):
            h.translate()
            h.update(-1, rots)
        for layer in range(ny):
            assert parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer)) == 0
        assert np.array_equal(o.download_atlas(), e.download_atlas()) or parity.compare_atlas(o.download_atlas(), e.download_atlas()) > 0.99


@pytest.mark.parametrize("name", sorted(SCENES))
def test_two_level_update_matches_oracle(name, oracle, make_emu):
    scene, (nxz, ny, nr) = SCENES[name]()
    p = make_params(nxz, ny, nr, flags=ALL)
    o, e = oracle.Oracle(p), make_emu(make_params(nxz, ny, nr, flags=ALL | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, e), seed=3)
    assert o.bvh_info().num_triangles == e.bvh_info().num_triangles == scene.num_triangles
    check_update(o, e, ny, rots)
    if scene.num_triangles:
        info = e.two_level_info()
        assert info.num_meshes >= 1 and info.stored_triangles <= scene.num_triangles
        assert info.tlas_nodes >= 1 and info.blas_nodes >= info.num_meshes
    o.close()
    e.close()


def test_two_level_stores_every_mesh_once(make_emu):
    scene = scenes.instanced_field(9, instances=64, grid=(4, 2, 64))
    e = make_emu(make_params(4, 2, 64, flags=L.FLAG_TWO_LEVEL))
    e.upload_scene(scene)
    e.set_scene_ubo(scene.ubo)
    e.build_accel()
    info = e.two_level_info()
    per_mesh = scene.mesh_num_indices // 3
    used = sorted(set(int(m) for m in scene.model_indices))
    assert info.num_meshes == len(used)
    assert info.stored_triangles == int(per_mesh[used].sum())
    assert e.bvh_info().num_triangles == scene.num_triangles > 10 * info.stored_triangles
    e.close()


def test_two_level_moving_instances_rebuild_the_instance_level_only(oracle, make_emu):
    scene = scenes.sponza_like(seed=4, grid=(4, 2, 64), movers=4, curtain_res=4)
    keep = [0, 1, 2, 3, 4] + [m[0] for m in scene.movers]
    vis = np.zeros(len(scene.renderables), np.uint32)
    vis[keep] = 1
    scene.renderables["visible"] = vis
    o, e = oracle.Oracle(make_params(4, 2, 64, flags=ALL)), make_emu(make_params(4, 2, 64, flags=ALL | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, e), seed=6)
    blas_nodes = e.two_level_info().blas_nodes
    for k in range(3):
        rend = scene.animated_renderables(0.4 * (k + 1))
        for h in (o, e):
            h.refit_accel(rend)
            h.update(-1, rots)
        for layer in range(2):
            assert parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer)) == 0
        parity.compare_atlas(o.download_atlas(), e.download_atlas())
        assert e.two_level_info().blas_nodes == blas_nodes
    # hiding an instance is a rebuild of the instance level (the meshes' hierarchies are kept)
    rend = scene.animated_renderables(2.0)
    rend["visible"][keep[-1]] = 0
    scene.renderables = rend
    for h in (o, e):
        h.upload_scene(scene)
        h.build_accel()
        h.update(-1, rots)
    for layer in range(2):
        assert parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer)) == 0
    o.close()
    e.close()


def test_two_level_shadow_rt_and_specular_gi(oracle, make_emu):
    """The other ray passes on the same structure: images bit-identical to the oracle's."""
    from tests import shadow_inputs as SI
    scene = scenes.random_soup(21, ntris=800)
    o, e = oracle.Oracle(make_params(4, 3, 64, flags=L.FLAG_POINT_LIGHTS)), make_emu(make_params(4, 3, 64, flags=L.FLAG_POINT_LIGHTS | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, e), seed=5)
    for h in (o, e):
        h.update(-1, rots)
    W, H = 40, 24
    depth, geom0 = SI.depth_buffer(W, H, seed=3), SI.gbuffer_normals(W, H, seed=4)
    ivp = SI.perspective_inv_view_proj(2)
    for h in (o, e):
        h.set_scene_ubo(SI.with_shadow_camera(scene.ubo, ivp, W, H))
    ro, re = o.shadow_rt(depth), e.shadow_rt(depth)
    assert np.array_equal(ro, re) and 0.02 < (ro == 0.0).mean() < 0.98
    for h in (o, e):
        h.set_scene_ubo(SI.with_specular_camera(scene.ubo, ivp, W, H))
    assert np.array_equal(o.specular_gi(depth, geom0), e.specular_gi(depth, geom0))
    o.close()
    e.close()


def test_two_level_deforming_mesh_rebuilds_only_its_hierarchy(oracle, make_emu):
    """Skinned / deforming meshes (UpdateBlasPass.cpp:60-175): new vertices into the giga vertex buffer, then refit — the hierarchies
    of the meshes whose vertices were written are rebuilt in place, the others and the shading records of untouched meshes are
    kept; three deformation steps bit-identical to the oracle (which re-flattens everything)."""
    scene, (nxz, ny, nr) = SCENES["textured"]()
    o, e = oracle.Oracle(make_params(nxz, ny, nr, flags=ALL)), make_emu(make_params(nxz, ny, nr, flags=ALL | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, e), seed=3)
    rng = np.random.RandomState(4)
    nv = len(scene.vertices)
    first, count = nv // 4, nv // 2
    for it in range(3):
        v = scene.vertices[first:first + count].copy()
        v["pos"] += rng.uniform(-0.03, 0.03, v["pos"].shape).astype(np.float32) * (it + 1)
        v["normal"] += rng.uniform(-0.2, 0.2, v["normal"].shape).astype(np.float32)
        v["uv"] += np.float32(0.1 * (it + 1))
        for h in (o, e):
            h.update_vertices(v, first)
            h.refit_accel()
            h.update(-1, rots)
        for layer in range(ny):
            assert parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer)) == 0
        assert np.array_equal(o.download_atlas(), e.download_atlas())
    o.close()
    e.close()
