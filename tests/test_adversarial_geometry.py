"""Edge cases of the geometry the path has to survive, on the CPU: the oracle's hierarchy against its own brute-force loop, and the
product's kernel logic (flatten, hierarchy build, BVH8 collapse, traversal, shading — host harness) against the oracle.

The scene holds what real assets contain and hierarchies dislike: exactly coincident duplicate surfaces (equal-t ties: the lowest
flattened triangle id wins on both sides), zero-area triangles (repeated vertex, collinear vertices: det == 0, never hit,
and a zero-extent box), needle and sliver triangles, a surface closer to the probes than tMin (ignored) and one just beyond tMax,
huge triangles spanning the whole scene next to tiny ones, a mirrored (negative determinant) transform, a renderable that is hidden
and one whose cull bit is set, and a one-triangle mesh."""
import numpy as np

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

ALL = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_TEXEL_RECORDS | L.FLAG_COUNTERS
GRID = (4, 3, 64)


def adversarial_scene():
    b = scenes.SceneBuilder("adversarial", 77)
    rng = b.rng
    up = lambda n: np.tile(np.array([0, 1, 0], np.float32), (n, 1))
    quad_p = np.array([(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1)], np.float32)
    quad = b.add_mesh(quad_p, up(4), np.array([0, 1, 2, 0, 2, 3], np.uint32))
    grey, red, glow = b.add_material((0.6, 0.6, 0.6)), b.add_material((0.8, 0.1, 0.1), 0.0, 0.4), b.add_material((0.2, 0.2, 0.9), 1.0, 0.3, (0.5, 0.4, 0.1, 2.0))
    # coincident duplicates: the same ceiling three times (two renderables with the same transform, one a two-mesh renderable of the same mesh)
    ceil = scenes.translate((0.3, 5.5, -0.2)) @ scenes.scale((7, 1, 7))
    b.add_renderable(ceil, quad, grey)
    b.add_renderable(ceil, [quad, quad], [red, glow])
    # floor far below: huge, a sliver next to it
    b.add_renderable(scenes.translate((0, -1.25, 0)) @ scenes.scale((40, 1, 40)), quad, grey)
    # zero-area and needle triangles
    deg_p = np.array([(0, 0, 0), (1, 0, 0), (1, 0, 0),            # repeated vertex
                      (0, 1, 0), (1, 1, 0), (2, 1, 0),            # collinear
                      (0, 2, 0), (3, 2, 0), (1.5, 2, 1e-6),       # needle
                      (0, 3, 0), (0, 3, 0), (0, 3, 0)], np.float32)   # a point
    deg = b.add_mesh(deg_p, up(12), np.arange(12, dtype=np.uint32))
    b.add_renderable(scenes.translate((-1.5, 0.5, 0.4)) @ scenes.rotate(2, 33.0), deg, red)
    # a small box around one probe, closer than tMin = 0.1 (never hit) and a shell just beyond tMax = 50 (never hit)
    bp, bn, bi = scenes.box(1)
    box = b.add_mesh(bp, bn, bi)
    b.add_renderable(scenes.translate((0.0, 2.0, 0.0)) @ scenes.scale((0.12, 0.12, 0.12)), box, glow)      # probes sit at integer x, z and even y
    b.add_renderable(scenes.scale((120, 120, 120)), box, grey)
    # mirrored transform (negative determinant), tiny triangles, a one-triangle mesh
    sp, sn, si = scenes.sphere(8, 6)
    sph = b.add_mesh(sp, sn, si)
    b.add_renderable(scenes.translate((1.2, 1.4, 0.9)) @ scenes.scale((-1.5, 1.5, 1.5)), sph, red)
    b.add_renderable(scenes.translate((-0.8, 3.1, -1.1)) @ scenes.scale((0.02, 0.02, 0.02)), sph, glow)
    one = b.add_mesh(np.array([(0, 0, 0), (2, 0, 0), (0, 0, 2)], np.float32), up(3), np.array([0, 1, 2], np.uint32))
    b.add_renderable(scenes.translate((-2, 0.9, -2)) @ scenes.rotate(0, 20.0), one, grey)
    # hidden (bit 0 clear) and culled (bit 1 set) renderables right in front of the probes: must not be hit
    b.add_renderable(scenes.translate((0, 1.0, 0)) @ scenes.scale((3, 1, 3)), quad, red, visible=0)
    b.add_renderable(scenes.translate((0, 1.1, 0)) @ scenes.scale((3, 1, 3)), quad, red, visible=3)
    b.add_light((0.5, 3.0, 0.5), 9.0, (4.0, 3.0, 2.0))
    b.add_light((-1.0, 1.0, 1.5), 6.0, (1.0, 2.0, 5.0))
    return b.finish((0.3, 2.0, -0.2), GRID)


def test_oracle_hierarchy_equals_brute_force_on_adversarial_geometry(oracle):
    scene = adversarial_scene()
    p = make_params(*GRID, flags=ALL)
    a, bf = oracle.Oracle(p), oracle.Oracle(p)
    bf.set_brute_force(True)
    rots = parity.setup_pair(scene, (a, bf), seed=9)
    for h in (a, bf):
        h.update(-1, rots)
    hit_tris = set()
    for layer in range(GRID[1]):
        ra, rb = a.download_ray_records(layer), bf.download_ray_records(layer)
        assert np.array_equal(ra["prim"], rb["prim"]) and np.array_equal(ra["t"], rb["t"]) and np.array_equal(ra["lit_mask"], rb["lit_mask"])
        assert np.array_equal(ra["radiance"].view(np.uint32), rb["radiance"].view(np.uint32))
        hit_tris |= set(int(x) for x in ra["prim"][ra["prim"] != L.MISS])
        t = ra["t"][ra["prim"] != L.MISS]
        assert t.min() > np.float32(0.1) and t.max() < 50.0
    assert np.array_equal(a.download_atlas(), bf.download_atlas())
    # equal-t ties go to the lowest id: the ceiling is hit, and only through its first copy (triangles 0, 1), never through copies 2-5
    assert hit_tris & {0, 1} and not hit_tris & {2, 3, 4, 5}
    # zero-area triangles (ids 8, 9, 11 of the flattened list: after 3 ceilings + floor = 8 triangles) are never hit
    assert not hit_tris & {8, 9, 11}
    a.close(); bf.close()


def test_product_kernel_logic_matches_oracle_on_adversarial_geometry(oracle, make_emu):
    scene = adversarial_scene()
    p = make_params(*GRID, flags=ALL)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=9)
    assert o.bvh_info().num_triangles == e.bvh_info().num_triangles == scene.num_triangles
    assert np.array_equal(o.download_instances(), e.download_instances())
    for it in range(2):
        for h in (o, e):
            h.update(-1, rots)
        for layer in range(GRID[1]):
            parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
        assert np.array_equal(o.download_atlas(), e.download_atlas())
    o.close(); e.close()


def test_empty_and_ragged_inputs(oracle, make_emu):
    """Nothing visible (an acceleration structure of zero triangles), a mesh whose index count is not a multiple of three (the last
    partial triangle is dropped, as primitiveCount = numIndices / 3 does in the reference's BLAS build, VulkanRenderer.cpp:1313),
    indices pointing past the vertex buffer (refused), a grid of one probe."""
    import pytest
    p = make_params(2, 1, 64, flags=ALL)
    # (a) every renderable hidden: zero triangles, every ray is the sky, both sides agree
    scene = scenes.quad_scene(grid=(2, 1, 64))
    scene.renderables["visible"][:] = 0
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=1)
    assert o.bvh_info().num_triangles == e.bvh_info().num_triangles == 0 and e.bvh_info().num_instances == 0
    for h in (o, e):
        h.update(-1, rots)
    assert np.all(e.download_ray_records(0)["prim"] == L.MISS)
    assert np.array_equal(o.download_atlas(), e.download_atlas())
    # (b) ragged index count: 7 indices = 2 triangles + 1 dangling index
    ragged = scenes.quad_scene(grid=(2, 1, 64))
    ragged.indices = np.concatenate([ragged.indices[:6], ragged.indices[:1]]).astype(np.uint32)
    ragged.mesh_num_indices = np.array([7], np.uint32)
    for h in (o, e):
        h.upload_scene(ragged); h.set_scene_ubo(ragged.ubo); h.build_accel()
        assert h.bvh_info().num_triangles == 2
        h.update(-1, rots)
    parity.compare_records(o.download_ray_records(0), e.download_ray_records(0))
    assert np.array_equal(o.download_atlas(), e.download_atlas())
    # (c) an index past the end of the vertex buffer, a mesh reaching past the end of the index buffer: refused by both, handle stays usable
    bad = scenes.quad_scene(grid=(2, 1, 64))
    bad.indices = bad.indices.copy(); bad.indices[4] = len(bad.vertices) + 5
    for h in (o, e):
        with pytest.raises(Exception):
            h.upload_scene(bad); h.build_accel()
    bad2 = scenes.quad_scene(grid=(2, 1, 64))
    bad2.mesh_num_indices = np.array([600], np.uint32)
    for h in (o, e):
        with pytest.raises(Exception):
            h.upload_scene(bad2); h.build_accel()
        h.upload_scene(ragged); h.set_scene_ubo(ragged.ubo); h.build_accel()
        assert h.bvh_info().num_triangles == 2
    o.close(); e.close()
    # (d) the smallest grid there is
    p1 = make_params(1, 1, 32, flags=ALL)
    one = scenes.quad_scene(grid=(1, 1, 32))
    o, e = oracle.Oracle(p1), make_emu(p1)
    rots = parity.setup_pair(one, (o, e), seed=2)
    for h in (o, e):
        h.update(-1, rots)
    parity.compare_records(o.download_ray_records(0), e.download_ray_records(0))
    assert np.array_equal(o.download_atlas(), e.download_atlas()) and o.download_atlas().shape == (10, 10, 4)
    o.close(); e.close()
