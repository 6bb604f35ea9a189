"""CPU checks of the product's kernel LOGIC: the per-thread kernel bodies (csrc/ddgi_kernels.cuh) compiled for the
host by the debug harness tests/host_emu, compared with the oracle.  The harness shares no code with oracle/; it is
the product's own flatten / LBVH / BVH8-collapse / traversal / shading / convolution / translation run serially.
The `-m gpu` tests repeat these comparisons with the real CUDA kernels through the C ABI.
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
}


@pytest.mark.parametrize("extra", [0, L.FLAG_SKIP_DEAD_SHADOW_RAYS], ids=["faithful", "skip_dead_shadow_rays"])
@pytest.mark.parametrize("name", sorted(SCENES))
def test_full_update_matches_oracle(name, extra, oracle, make_emu):
    scene, (nxz, ny, nr) = SCENES[name]()
    p = make_params(nxz, ny, nr, flags=ALL | extra)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=3)
    assert o.bvh_info().num_triangles == e.bvh_info().num_triangles == scene.num_triangles
    for it in range(2):                       # 2nd update reads a non-trivial previous atlas (multi-bounce)
        for h in (o, e):
            h.translate()
            h.update(-1, rots)
        for layer in range(ny):
            parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
            io, do = o.download_rays(layer)
            ie, de = e.download_rays(layer)
            assert np.array_equal(do, de)
            assert parity.half_ulp_distance(io, ie).max() <= 1
        assert parity.compare_atlas(o.download_atlas(), e.download_atlas()) > 0.99
        err = parity.rel_err(e.download_texel_records(), o.download_texel_records(), floor=1e-4)
        assert err.max() <= parity.TEXEL_REL_TOL
    st = e.stats()
    assert st["primary_rays"] == 2 * nxz * nxz * ny * nr and st["primary_hits"] <= st["primary_rays"]
    o.close()
    e.close()


def test_sequential_layer_mode_and_frame_cadence(oracle, make_emu):
    scene = scenes.random_soup(21, ntris=600)
    p = make_params(4, 3, 64, flags=ALL)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=4)
    seen = []
    for f in range(7):                        # reference cadence: one layer per frame, round-robin (RayTracingPass.cpp:112,160-164)
        lo, le = o.frame(rots[f % 3]), e.frame(rots[f % 3])
        assert lo == le == f % 3
        seen.append(le)
        assert parity.compare_atlas(o.download_atlas(), e.download_atlas()) > 0.99
    # separate trace/convolve calls == update(layer)
    e.trace_layer(1, rots[1])
    e.convolve_layer(1)
    o.update(1, rots[1])
    parity.compare_atlas(o.download_atlas(), e.download_atlas())
    o.close()
    e.close()


def test_translation_matches_oracle(oracle, make_emu):
    scene = scenes.random_soup(31, ntris=400)
    p = make_params(4, 2, 64)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=5)
    rng = np.random.RandomState(0)
    cam = np.array([0.3, 2.0, -0.2])
    for step in range(8):
        cam = cam + rng.uniform(-1.3, 1.3, 3) * (1, 0, 1)
        ubo = scene.with_camera(cam)
        for h in (o, e):
            h.set_scene_ubo(ubo)
            h.translate()
            h.update(-1, rots)
        assert parity.compare_atlas(o.download_atlas(), e.download_atlas()) > 0.99
    o.close()
    e.close()


def test_bvh8_structure(make_emu):
    for scene in (scenes.random_soup(41, ntris=700), scenes.cornell_box(grid=(4, 2, 64), wall_res=6, box_res=2)):
        p = make_params(4, 2, 64)
        e = make_emu(p)
        e.upload_scene(scene)
        e.set_scene_ubo(scene.ubo)
        e.build_accel()
        info = e.bvh_info()
        assert info.node_bytes == 80 and info.tri_bytes == 48 and info.num_triangles == scene.num_triangles
        nodes, tris = e.download_bvh()
        parity.validate_bvh8(nodes, tris, info.num_triangles)
        e.close()


def test_refit_moving_nodes(oracle, make_emu):
    scene = scenes.sponza_like(seed=4, grid=(4, 2, 64), movers=4, curtain_res=4)
    # shrink: keep only the movers and a few big static pieces so the CPU harness stays fast
    keep = [0, 1, 2, 3, 4] + [m[0] for m in scene.movers]
    vis = np.zeros(len(scene.renderables), np.uint32)
    vis[keep] = 1
    scene.renderables["visible"] = vis
    p = make_params(4, 2, 64, flags=ALL)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=6)
    for k in range(3):
        rend = scene.animated_renderables(0.4 * (k + 1))
        for h in (o, e):
            h.refit_accel(rend)
            h.update(-1, rots)
        nodes, tris = e.download_bvh()
        parity.validate_bvh8(nodes, tris, e.bvh_info().num_triangles)
        for layer in range(2):
            parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
        parity.compare_atlas(o.download_atlas(), e.download_atlas())
    # a visibility change needs a rebuild, like the reference's per-frame TLAS build would pick up
    rend = scene.animated_renderables(2.0)
    rend["visible"][keep[-1]] = 0
    with pytest.raises(Exception):
        e.refit_accel(rend)
    o.close()
    e.close()


def test_error_behaviour(make_emu):
    p = make_params(4, 2, 64)
    e = make_emu(p)
    scene = scenes.quad_scene()
    with pytest.raises(Exception):
        e.build_accel()                                   # before upload_scene
    e.upload_scene(scene)
    e.set_scene_ubo(scene.ubo)
    with pytest.raises(Exception):
        e.trace_layer(0, np.eye(4, dtype=np.float32))     # before build_accel
    e.build_accel()
    with pytest.raises(Exception):
        e.convolve_layer(1)                               # layer never traced
    with pytest.raises(Exception):
        e.trace_layer(5, np.eye(4, dtype=np.float32))     # layer out of range
    bad = scenes.quad_scene()
    bad.materials["bindless_indices"][0, 1] = 3           # textured material: outside the path's scope
    with pytest.raises(Exception):
        e.upload_scene(bad)
    # DDGI off: the passes are no-ops (IrradianceProbeRayTracingPass.cpp:97-98)
    off = scene.ubo.copy()
    off["flags"] = 0
    e.set_scene_ubo(off)
    before = e.download_atlas()
    e.update(-1, np.tile(np.eye(4, dtype=np.float32).ravel(), 2))
    assert np.array_equal(before, e.download_atlas())
    e.close()
    with pytest.raises(Exception):
        make_emu(make_params(4, 2, 48))                   # rays per probe must tile


@pytest.mark.parametrize("clamp", [False, True], ids=["repeat", "clamp_to_edge"])
def test_atlas_consumer_matches_oracle(clamp, oracle, make_emu):
    """calcIndirectDiffuseLight per surface point (deferred_tiled.comp:264-296 with the CLAMP_TO_EDGE sampler,
    mirror_reflections.rchit:92-118 with REPEAT) against the atlas the update produced."""
    from tests.consumer_points import surface_points
    scene, (nxz, ny, nr) = SCENES["cornell"]()
    p = make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=5)
    for h in (o, e):
        for _ in range(2):
            h.update(-1, rots)
    assert np.array_equal(o.download_atlas(), e.download_atlas())
    pts = surface_points(3000, nxz, ny, seed=17)
    view = scene.ubo["camera_pos"][0][:3] if hasattr(scene.ubo, "dtype") else (0.0, 1.0, 0.0)
    ro, re = o.indirect_diffuse(pts, view, clamp), e.indirect_diffuse(pts, view, clamp)
    assert np.isfinite(ro).all() and (ro[:, :3] >= 0).all() and ro[:, :3].max() > 0
    assert np.array_equal(ro.view(np.uint32), re.view(np.uint32))
    other = o.indirect_diffuse(pts, view, not clamp)
    assert not np.array_equal(other, ro)            # the sampler mode matters at the edge of the grid
    o.close()
    e.close()


def test_textures_change_the_result_and_are_validated(oracle, make_emu):
    """The textured scene really samples its textures (stripping them changes the atlas), materials that point at a missing
    texture are rejected, and the (never sampled) normal-map index is not."""
    scene, (nxz, ny, nr) = SCENES["textured"]()
    p = make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS)
    e = make_emu(p)
    rots = parity.setup_pair(scene, (e,), seed=3)
    e.update(-1, rots)
    with_tex = e.download_atlas()
    import copy
    plain = copy.copy(scene)
    plain.materials = scene.materials.copy()
    plain.materials["bindless_indices"][:] = -1
    plain.textures = []
    e2 = make_emu(p)
    parity.setup_pair(plain, (e2,), seed=3)
    e2.update(-1, rots)
    assert not np.array_equal(with_tex, e2.download_atlas())
    bad = copy.copy(scene)
    bad.textures = scene.textures[:2]
    e3 = make_emu(p)
    with pytest.raises(Exception):
        e3.upload_scene(bad)
    for h in (e, e2, e3):
        h.close()


def test_async_download_entry_points_in_the_host_harness(make_emu):
    scene, (nxz, ny, nr) = SCENES["quad"]()
    e = make_emu(make_params(nxz, ny, nr))
    rots = parity.setup_pair(scene, (e,), seed=2)
    e.update(-1, rots)
    buf = np.empty((e.atlas_height, e.atlas_width, 4), np.uint16)
    e.download_atlas_async(buf)
    e.download_wait()
    assert np.array_equal(buf, e.download_atlas())
    e.close()


def test_shadow_rt_matches_oracle(oracle, make_emu):
    """the "ShadowRT" pass (shadow.rgen / shadow.rmiss) on the path's BVH; bit-exact visibility image."""
    from tests import shadow_inputs as SI
    scene, (nxz, ny, nr) = SCENES["cornell"]()
    p = make_params(nxz, ny, nr)
    o, e = oracle.Oracle(p), make_emu(p)
    parity.setup_pair(scene, (o, e), seed=5)
    W, H = 48, 32
    depth = SI.depth_buffer(W, H, seed=3)
    for ivp in (SI.ortho_inv_view_proj(7.0, 7.0, 14.0), SI.perspective_inv_view_proj(2)):
        ubo = SI.with_shadow_camera(scene.ubo, ivp, W, H)
        for h in (o, e):
            h.set_scene_ubo(ubo)
        ro, re = o.shadow_rt(depth), e.shadow_rt(depth)
        assert np.array_equal(ro, re)
        assert set(np.unique(ro)) <= {0.0, 1.0} and (ro[depth == 1.0] == 1.0).all()
    assert 0.05 < (ro == 0.0).mean() < 0.95                        # both outcomes occur
    off = SI.with_shadow_camera(scene.ubo, ivp, W, H, rt_shadows=False)
    for h in (o, e):
        h.set_scene_ubo(off)
        assert (h.shadow_rt(depth) == -1.0).all()                   # shadow.rgen:32-34: nothing written
    o.close()
    e.close()


def test_bake_mode_matches_oracle(oracle, make_emu):
    """UBO_BAKE_MODE_ON_FLAG — tile-placed grid (rgen:61-73, pbr_light.glsl:274-291), running-mean convolution
    (probe_conv.comp:67-79), translation disabled; three accumulating updates, atlas bit-identical."""
    scene, (nxz, ny, nr) = SCENES["cornell"]()
    p = make_params(nxz, ny, nr, flags=ALL)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=6)
    ubo = scene.ubo.copy()
    ubo["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
    ubo["bake_tile_info"][0] = (-1, 0, 6, 0)
    for h in (o, e):
        h.set_scene_ubo(ubo)
    for it in range(3):
        for h in (o, e):
            h.translate()
            h.update(-1, rots)
        for layer in range(ny):
            parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
        assert np.array_equal(o.download_atlas(), e.download_atlas())
    a = o.download_atlas().view(np.float16).astype(np.float32)
    assert a[..., 3].max() > 40.0                                    # alpha carries the accumulated weight (3 x ~16 at 64 rays)
    o.close()
    e.close()


@pytest.mark.parametrize("name", ["cornell", "textured"])
def test_specular_gi_matches_oracle(name, oracle, make_emu):
    """the "SpecularGIRT" pass (mirror_reflections.rgen/.rchit/.rmiss) on the path's BVH and atlas; the rgba16f
    image is bit-identical."""
    from tests import shadow_inputs as SI
    scene, (nxz, ny, nr) = SCENES[name]()
    p = make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS)
    o, e = oracle.Oracle(p), make_emu(p)
    rots = parity.setup_pair(scene, (o, e), seed=5)
    for h in (o, e):
        h.update(-1, rots)                       # a non-trivial atlas for the indirect term
    W, H = 40, 24
    depth, geom0 = SI.depth_buffer(W, H, seed=3), SI.gbuffer_normals(W, H, seed=4)
    for ivp in (SI.ortho_inv_view_proj(5.0, 5.0, 8.0), SI.perspective_inv_view_proj(2)):
        ubo = SI.with_specular_camera(scene.ubo, ivp, W, H)
        for h in (o, e):
            h.set_scene_ubo(ubo)
        ro, re = o.specular_gi(depth, geom0), e.specular_gi(depth, geom0)
        assert np.array_equal(ro, re)
        assert (ro[depth == 1.0] == 0).all()
    f = ro.view(np.float16).astype(np.float32)
    hit = np.isfinite(f[..., 3]) & (f[..., 3] > 0)
    assert hit.any() and (~hit & (depth != 1.0)).any()                 # hits (finite distance) and misses (1e6 -> inf in fp16)
    off = SI.with_specular_camera(scene.ubo, ivp, W, H, enabled=False)
    for h in (o, e):
        h.set_scene_ubo(off)
        assert (h.specular_gi(depth, geom0) == 0xFFFF).all()          # SpecularGIRTPass.cpp:102-103: nothing written
    o.close()
    e.close()


def test_replacing_textures_after_the_scene_is_refused(make_emu):
    scene, (nxz, ny, nr) = SCENES["textured"]()
    e = make_emu(make_params(nxz, ny, nr))
    rots = parity.setup_pair(scene, (e,), seed=3)
    e.update(-1, rots)
    e.upload_textures(scene.textures[:1])            # material indices now dangle
    with pytest.raises(Exception):
        e.update(-1, rots)
    e.upload_scene(scene)                            # re-uploads the full array, then the scene: valid again
    e.build_accel()
    e.update(-1, rots)
    e.close()


def _bake_tile_atlas(make_handle, scene, grid, tile, updates=3):
    """Bakes tile `tile` (x, z) of size 32 with a handle in bake mode; returns the atlas as an RGBA16F texture."""
    nxz, ny, nr = grid
    h = make_handle(make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS))
    rots = parity.setup_pair(scene, (h,), seed=11)
    u = scene.ubo.copy()
    u["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
    u["bake_tile_info"][0] = (tile[0], tile[1], 32, 0)
    h.set_scene_ubo(u)
    for _ in range(updates):
        h.update(-1, rots)
    atlas = h.download_atlas().copy()
    h.close()
    return atlas


def test_baked_tile_atlases_match_oracle(oracle, make_emu):
    """a tileInfoBuffer with a baked atlas.  sampleProbe places the grid on the baked tile
    (pbr_light.glsl:251-273) for the probe hit shader (which still samples the live atlas), and the deferred-pass consumer
    samples the tile's baked atlas (deferred_tiled.comp:264-296); hit shading, atlas and consumer output bit-identical."""
    import copy
    from tests.consumer_points import surface_points
    scene, grid = SCENES["cornell"]()
    nxz, ny, nr = grid
    baked = _bake_tile_atlas(oracle.Oracle, scene, grid, (0, 0))
    assert np.array_equal(baked, _bake_tile_atlas(make_emu, scene, grid, (0, 0)))
    tiled = copy.copy(scene)
    tiled.textures = list(scene.textures) + [(baked.view(np.uint8).reshape(baked.shape[0], baked.shape[1], 8), L.TEX_RGBA16F, False)]
    tiles = np.full(21 * 21, -1, np.int32)
    tiles[10 * 21 + 10] = len(tiled.textures) - 1             # the camera's own tile (0, 0): x, z in [0, 32)
    p = make_params(nxz, ny, nr, flags=ALL)
    o, e, plain = oracle.Oracle(p), make_emu(p), make_emu(p)
    rots = parity.setup_pair(tiled, (o, e), seed=3)
    parity.setup_pair(scene, (plain,), seed=3)
    for h in (o, e):
        h.upload_tile_infos(tiles)
    for it in range(2):
        for h in (o, e, plain):
            h.update(-1, rots)
        for layer in range(ny):
            parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
        assert np.array_equal(o.download_atlas(), e.download_atlas())
    assert not np.array_equal(e.download_atlas(), plain.download_atlas())          # the tile placement changes the multi-bounce term
    pts = surface_points(2000, nxz, ny, seed=21)
    pts[:, 0] = np.abs(pts[:, 0]); pts[:500, 2] = np.abs(pts[:500, 2])              # many points inside the baked tile
    view = (0.0, 1.0, 0.0)
    for clamp in (True, False):
        ro, re = o.indirect_diffuse(pts, view, clamp), e.indirect_diffuse(pts, view, clamp)
        assert np.array_equal(ro.view(np.uint32), re.view(np.uint32))
        assert not np.array_equal(re, plain.indirect_diffuse(pts, view, clamp))
    for h in (o, e):
        h.upload_tile_infos([])                                                       # removing the buffer restores the live placement
    assert np.array_equal(e.indirect_diffuse(pts, view, False)[:, :3] >= 0, np.ones((2000, 3), bool))
    for h in (o, e, plain):
        h.close()


def test_deforming_mesh_update_vertices_and_refit(oracle, make_emu):
    """Skinned / deforming meshes (the refit half): new vertex positions, normals and uvs into the giga vertex
    buffer, then refit — triangles, boxes AND shading records follow; three deformation steps bit-identical to the oracle."""
    scene, (nxz, ny, nr) = SCENES["textured"]()
    p = make_params(nxz, ny, nr, flags=ALL)
    o, e = oracle.Oracle(p), make_emu(p)
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
            parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
        assert np.array_equal(o.download_atlas(), e.download_atlas())
    with pytest.raises(Exception):
        e.update_vertices(scene.vertices, 1)                      # range runs past the buffer
    o.close()
    e.close()


def test_timer_levels_and_pipeline_switch(make_emu):
    """anerend_cuda_ddgi_set_timers: level 0 reports no pass times, level 1 (the default) the three passes; bad levels are refused.
    set_pipeline is accepted by the host harness (it has one 'stream': nothing to overlap) and leaves the results alone."""
    from anerend_b200.capi import Timings
    import ctypes
    scene, (nxz, ny, nr) = SCENES["quad"]()
    a, b = make_emu(make_params(nxz, ny, nr)), make_emu(make_params(nxz, ny, nr))
    rots = parity.setup_pair(scene, (a, b), seed=3)

    def timings(h):
        t = Timings()
        h._call("get_timings", ctypes.byref(t))
        return t
    a.update(-1, rots)
    assert timings(a).rt_ms > 0.0
    b.set_timers(0)
    b.set_pipeline(True)
    b.update(-1, rots)
    t = timings(b)
    assert t.rt_ms == 0.0 and t.convolve_ms == 0.0 and t.kernel_launches > 0
    assert np.array_equal(a.download_atlas(), b.download_atlas())
    b.set_timers(1)
    b.update(-1, rots)
    assert timings(b).rt_ms > 0.0
    with pytest.raises(Exception):
        b.set_timers(3)
    a.close()
    b.close()
