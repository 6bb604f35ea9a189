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


def test_two_level_random_scenes_bit_exact(oracle, make_emu):
    """Seeded fuzz (a slice of tests/test_fuzz_parity.py with the flag set): random soups / textured soups / Cornell boxes / instanced
    fields, random grids, ray counts, cameras and flags — two updates each, everything bit for bit."""
    for seed in range(5000, 5032):
        rng = np.random.RandomState(seed)
        nxz, ny, nr = int(rng.choice([1, 2, 3, 5])), int(rng.choice([1, 2, 3])), int(rng.choice([32, 64, 128, 256]))
        cam = tuple(rng.uniform(-3, 3, 3) * [1, 0.5, 1] + [0, 1.5, 0])
        kind = rng.randint(0, 4)
        if kind == 0:
            s = scenes.random_soup(seed, ntris=int(rng.choice([40, 300, 1500])), extent=float(rng.uniform(2, 12)), grid=(nxz, ny, nr), camera=cam,
                                   nlights=int(rng.randint(0, 4)))
        elif kind == 1:
            s = scenes.textured_soup(seed, ntris=int(rng.choice([60, 400])), grid=(nxz, ny, nr), camera=cam)
        elif kind == 2:
            s = scenes.cornell_box(seed=seed, grid=(nxz, ny, nr), wall_res=int(rng.choice([2, 6])), box_res=2)
            s.ubo = s.with_camera(cam)
        else:
            s = scenes.instanced_field(seed, instances=int(rng.choice([3, 30, 120])), grid=(nxz, ny, nr), extent=float(rng.uniform(4, 16)), seg=(8, 4),
                                       nlights=int(rng.randint(0, 4)))
        if rng.rand() < 0.3:
            s.ubo["flags"] &= ~np.uint32(L.UBO_DDGI_MULTI_FLAG)
        extra = L.FLAG_SKIP_DEAD_SHADOW_RAYS if rng.rand() < 0.5 else 0
        o, e = oracle.Oracle(make_params(nxz, ny, nr, flags=ALL | extra)), make_emu(make_params(nxz, ny, nr, flags=ALL | extra | L.FLAG_TWO_LEVEL))
        rots = parity.setup_pair(s, (o, e), seed=seed)
        for it in range(2):
            for h in (o, e):
                h.translate()
                h.update(-1, rots)
            for layer in range(ny):
                assert parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer)) == 0, seed
            assert np.array_equal(o.download_atlas(), e.download_atlas()), seed
        o.close()
        e.close()


def test_two_level_random_refits_and_vertex_updates_bit_exact(oracle, make_emu):
    """Moving instances (random rigid steps, now and then mirrored) and deforming vertex ranges, refit after each step."""
    for seed in range(6000, 6012):
        rng = np.random.RandomState(seed)
        nxz, ny, nr = int(rng.choice([2, 3])), int(rng.choice([1, 2])), 64
        s = scenes.random_soup(seed, ntris=int(rng.choice([100, 600])), grid=(nxz, ny, nr))
        o, e = oracle.Oracle(make_params(nxz, ny, nr, flags=ALL)), make_emu(make_params(nxz, ny, nr, flags=ALL | L.FLAG_TWO_LEVEL))
        rots = parity.setup_pair(s, (o, e), seed=seed)
        verts = s.vertices.copy()
        for step in range(3):
            rend = s.renderables.copy()
            for i in range(len(rend)):
                if rng.rand() < 0.6:
                    M = rend["transform"][i].reshape(4, 4).T.astype(np.float64)
                    M2 = scenes.translate(rng.uniform(-1, 1, 3)) @ scenes.rotate(int(rng.randint(0, 3)), rng.uniform(0, 360)) @ M
                    if rng.rand() < 0.2:
                        M2 = M2 @ scenes.scale((-1, 1, 1))
                    rend["transform"][i] = M2.T.astype(np.float32).ravel()
            k0 = int(rng.randint(0, len(verts) - 10))
            k1 = min(len(verts), k0 + int(rng.randint(1, 200)))
            verts["pos"][k0:k1] += rng.uniform(-0.05, 0.05, (k1 - k0, 3)).astype(np.float32)
            for h in (o, e):
                h.update_vertices(verts[k0:k1], first=k0)
                h.refit_accel(rend)
                h.update(-1, rots)
            for layer in range(ny):
                assert parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer)) == 0, (seed, step)
            assert np.array_equal(o.download_atlas(), e.download_atlas()), (seed, step)
        o.close()
        e.close()
