"""Known-answer tests that pin the oracle's transliteration of the shaders.

The reference has no tests, golden vectors or CPU path for this chain ("parity unpinned"); these KATs are derived
from the shader formulas themselves, each citing the function it pins.
"""
import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes


def test_fibonacci_directions(oracle):
    # irradiance_probe_update.rgen:35-46
    d = oracle.fibonacci_dirs(256)
    np.testing.assert_allclose(d[0], [0.088302, 0.0, 0.996094], atol=2e-6)
    np.testing.assert_allclose(d[1], [-0.112555, -0.103109, 0.988281], atol=2e-6)
    np.testing.assert_allclose(d[255], [-0.071870, -0.051303, -0.996094], atol=2e-6)
    np.testing.assert_allclose(oracle.fibonacci_dirs(64)[0], [0.176085, 0.0, 0.984375], atol=2e-6)
    np.testing.assert_allclose(oracle.fibonacci_dirs(128)[0], [0.124756, 0.0, 0.992188], atol=2e-6)
    for n in (64, 128, 256):
        d = oracle.fibonacci_dirs(n).astype(np.float64)
        np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)
        np.testing.assert_allclose(d[:, 2], 1.0 - 2.0 * (np.arange(n) + 0.5) / n, atol=1e-6)   # cos(phi)
        assert np.abs(d.mean(0)).max() < 0.02                                                   # near-uniform over the sphere


def test_oct_decode_texel_centres_and_round_trip(oracle):
    # octahedron_helpers.glsl:10-34, helpers.glsl:1-4
    want = {(0, 0): (-0.162221, -0.973329, -0.162221), (3, 3): (-0.162221, 0.973329, -0.162221),
            (4, 4): (0.162221, 0.973329, 0.162221), (7, 7): (0.162221, -0.973329, 0.162221)}
    for (i, j), v in want.items():
        o = [((2 * i + 1) / 16.0) * 2 - 1, ((2 * j + 1) / 16.0) * 2 - 1]
        np.testing.assert_allclose(oracle.oct_decode(o), v, atol=2e-6)
    for i in range(8):
        for j in range(8):
            o = np.array([((2 * i + 1) / 16.0) * 2 - 1, ((2 * j + 1) / 16.0) * 2 - 1], np.float32)
            np.testing.assert_allclose(oracle.oct_encode(oracle.oct_decode(o)), o, atol=1e-6)
    rng = np.random.RandomState(0)
    for _ in range(200):
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        np.testing.assert_allclose(oracle.oct_decode(oracle.oct_encode(v)), v, atol=2e-6)


def test_to_linear(oracle):
    # bindless.glsl:240-247 and rmiss:20 with RenderOptions defaults
    np.testing.assert_allclose(oracle.to_linear([0.2, 0.5, 0.7]), [0.033105, 0.214041, 0.447988], atol=1e-6)
    assert np.array_equal(oracle.to_linear([1, 1, 1]), np.ones(3, np.float32))
    lo = oracle.to_linear([0.04044, 0.0, 0.0])[0]
    hi = oracle.to_linear([0.04045, 0.0, 0.0])[0]
    assert lo == pytest.approx(0.04044 / 12.92, rel=1e-6)
    assert hi == pytest.approx(((0.04045 + 0.055) / 1.055) ** 2.4, rel=1e-6)


def test_half_conversion_matches_ieee(oracle):
    rng = np.random.RandomState(1)
    vals = np.concatenate([rng.uniform(-70000, 70000, 2000), rng.uniform(-1, 1, 2000) * 10.0 ** rng.uniform(-9, 0, 2000),
                           [0.0, -0.0, 65504.0, 65519.9, 65520.0, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 1.0, 0.1]]).astype(np.float32)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([oracle.float_to_half(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got, want)
    allh = np.arange(0, 0x7C00, 7, dtype=np.uint16)
    back = np.array([oracle.half_to_float(int(h)) for h in allh], np.float32)
    assert np.array_equal(back, allh.view(np.float16).astype(np.float32))


def test_rotation_matches_glm_rotate(oracle):
    # glm::rotate(mat4(1), radians(a), axis): column-major; axis = z -> columns (c,s,0),(-s,c,0),(0,0,1)
    m = oracle.make_rotation(2, 30.0).reshape(4, 4)     # rows of this view are glm columns
    c, s = np.cos(np.radians(30.0)), np.sin(np.radians(30.0))
    np.testing.assert_allclose(m, [[c, s, 0, 0], [-s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], atol=1e-7)
    mx = oracle.make_rotation(0, 90.0).reshape(4, 4)
    np.testing.assert_allclose(mx, [[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], atol=1e-7)
    for ax in range(3):
        r = oracle.make_rotation(ax, 123.4).reshape(4, 4)[:3, :3].astype(np.float64)
        np.testing.assert_allclose(r @ r.T, np.eye(3), atol=1e-6)


def test_intersection_analytic(oracle):
    # KAT 9: axis-aligned quad; Vulkan interval is exclusive at both ends
    v0, v1, v2 = [0, 0, 5], [4, 0, 5], [0, 4, 5]
    hit, t, u, v = oracle.intersect([1, 1, 0], [0, 0, 1], v0, v1, v2)
    assert hit and t == 5.0 and u == 0.25 and v == 0.25
    assert oracle.intersect([1, 1, 10], [0, 0, -1], v0, v1, v2)[0]                   # two-sided
    assert not oracle.intersect([3, 3, 0], [0, 0, 1], v0, v1, v2)[0]                 # u+v > 1
    assert not oracle.intersect([1, 1, 0], [1, 0, 0], v0, v1, v2)[0]                 # parallel: det == 0
    assert not oracle.intersect([1, 1, 4.95], [0, 0, 1], v0, v1, v2, 0.1, 50)[0]     # t = 0.05 < tMin
    assert not oracle.intersect([1, 1, 0], [0, 0, 1], v0, v1, v2, 0.1, 5.0)[0]       # t == tMax excluded
    assert not oracle.intersect([1, 1, 0], [0, 0, 1], v0, v1, v2, 5.0, 50.0)[0]      # t == tMin excluded


def _conv_setup(oracle, n_xz, n_y, n_rays, seed):
    p = make_params(n_xz, n_y, n_rays, flags=L.FLAG_TEXEL_RECORDS)
    o = oracle.Oracle(p)
    sx, sy = o.ray_tile
    rot = oracle.make_rotation(*scenes.rotation_sequence(seed, 1)[0])
    fib = oracle.fibonacci_dirs(n_rays).astype(np.float32)
    R = rot.reshape(4, 4)[:3, :3].T                       # math matrix from column-major storage
    dirs = (fib @ R.T).astype(np.float32)
    dir_img = np.zeros((sy * n_xz, sx * n_xz, 4), np.float16)
    for ray in range(n_rays):
        px, py = ray % sx, ray // sx
        dir_img[py::sy, px::sx, :3] = dirs[ray]
    dir_img[..., 3] = 1.0
    return o, dir_img, dirs


def test_convolution_sum_of_weights_and_uniform_radiance(oracle):
    # KAT 3 + KAT 4: probe_conv.comp:49-87
    for n_rays, want in ((64, 16.0184), (128, 31.9457), (256, 64.0239)):
        p = make_params(2, 1, n_rays, flags=L.FLAG_TEXEL_RECORDS)
        o = oracle.Oracle(p)
        sx, sy = o.ray_tile
        fib = oracle.fibonacci_dirs(n_rays)
        # identity rotation, fp32 directions -> emulate by computing the expected sum directly
        octdir = oracle.oct_decode([((2 * 3 + 1) / 16.0) * 2 - 1] * 2).astype(np.float64)
        w = np.maximum(0.0, fib.astype(np.float64) @ octdir)
        assert w[w > 1e-6].sum() == pytest.approx(want, abs=2e-3)
        o.close()
    o, dir_img, _ = _conv_setup(oracle, 4, 2, 256, seed=1)
    L0 = np.array([0.75, 1.5, 0.25], np.float16)
    irr = np.zeros_like(dir_img)
    irr[..., :3] = L0
    irr[..., 3] = 1.0
    for layer in range(2):
        o.set_rays(layer, irr.view(np.uint16), dir_img.view(np.uint16))
    mask = np.broadcast_to(__import__("tests.parity", fromlist=["x"]).interior_mask(4, 2)[..., None], (80, 40, 4))
    prev = np.zeros(3)
    for k in range(1, 6):
        for layer in range(2):
            o.convolve_layer(layer)
        rec = o.download_texel_records()
        rgb = rec[..., :3][mask[..., :3]].reshape(-1, 3)
        # normalised result == L exactly for uniform radiance, so rgb follows 0.04*L + 0.96*old on half-rounded olds
        want = L0.astype(np.float64) * (1 - 0.96 ** k)
        np.testing.assert_allclose(rgb, np.broadcast_to(want, rgb.shape), rtol=2e-3)
        assert np.all(rgb.mean(0) > prev)
        prev = rgb.mean(0)
    # borders are never written: still (0,0,0,1)
    a = o.download_atlas()
    border = ~mask[..., 0]
    assert np.all(a[border][:, :3] == 0) and np.all(a[border][:, 3] == 0x3C00)
    o.close()


def test_config1_convolution_runs_single_thread(oracle):
    # BASELINE config 1: 8x8x8 probes x 64 rays, synthetic radiance U[0,4)^3 seed 1, one update of all layers
    o, dir_img, dirs = _conv_setup(oracle, 8, 8, 64, seed=1)
    o.set_threads(1)
    rng = np.random.RandomState(1)
    for layer in range(8):
        irr = np.ones(dir_img.shape, np.float16)
        irr[..., :3] = rng.uniform(0, 4, irr[..., :3].shape)
        o.set_rays(layer, irr.view(np.uint16), dir_img.view(np.uint16))
        o.convolve_layer(layer)
        if layer == 0:
            # independent numpy restatement for probe (0,0), texel (i,j): probe_conv.comp:43-87
            irr0 = irr[:8, :8, :3].reshape(64, 3).astype(np.float64)       # ray = py*8+px
            d16 = dir_img[:8, :8, :3].reshape(64, 3).astype(np.float64)
            rec = o.download_texel_records()
            for (i, j) in ((0, 0), (3, 5), (7, 7)):
                od = oracle.oct_decode([((2 * i + 1) / 16.0) * 2 - 1, ((2 * j + 1) / 16.0) * 2 - 1]).astype(np.float64)
                w = np.maximum(0, d16 @ od)
                w[w <= 1e-6] = 0
                want_rgb = (w[:, None] * irr0).sum(0) / w.sum() * 0.04
                want_a = w.sum() * (1.0 - np.float32(0.96)) + 1.0 * np.float32(0.96)
                np.testing.assert_allclose(rec[1 + j, 1 + i, :3], want_rgb, rtol=2e-5)
                assert rec[1 + j, 1 + i, 3] == pytest.approx(want_a, rel=2e-5)
    a = o.download_atlas().view(np.float16)
    assert np.isfinite(a.astype(np.float32)).all() and a[..., :3].max() > 0
    o.close()


def test_empty_scene_gives_sky_everywhere(oracle):
    # KAT 5: every ray misses -> rmiss:20 -> atlas = 0.04 * sky after one update
    s = scenes.empty_scene(grid=(4, 2, 64))
    p = make_params(4, 2, 64, flags=L.FLAG_RAY_RECORDS | L.FLAG_TEXEL_RECORDS)
    o = oracle.Oracle(p)
    o.upload_scene(s)
    o.set_scene_ubo(s.ubo)
    o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(5, 2))
    o.update(-1, rots)
    sky = oracle.to_linear([0.2, 0.5, 0.7])
    for layer in range(2):
        rec = o.download_ray_records(layer)
        assert np.all(rec["prim"] == L.MISS)
        assert np.array_equal(rec["radiance"], np.broadcast_to(sky, rec["radiance"].shape))
    from tests.parity import interior_mask
    tex = o.download_texel_records()[interior_mask(4, 2)]
    skyh = sky.astype(np.float16).astype(np.float64)
    np.testing.assert_allclose(tex[:, :3], np.broadcast_to(skyh * (1.0 - np.float32(0.96)), tex[:, :3].shape), rtol=1e-5)
    o.close()


def test_atlas_addressing_and_translation(oracle):
    # KAT 7 + KAT 8: probe (x,y,z) texel (i,j) <-> pixel (10x+1+i, 10*nXZ*y+10z+1+j); translation by one probe
    n_xz, n_y = 4, 2
    p = make_params(n_xz, n_y, 64)
    o = oracle.Oracle(p)
    W, H = 10 * n_xz, 10 * n_xz * n_y
    a = np.zeros((H, W, 4), np.uint16)
    yy, xx = np.mgrid[0:H, 0:W]
    a[..., 0] = xx
    a[..., 1] = yy
    s = scenes.empty_scene(grid=(n_xz, n_y, 64))
    o.set_scene_ubo(s.with_camera((0.5, 1.0, 0.5)))
    o.translate()
    o.upload_atlas(a)
    o.set_scene_ubo(s.with_camera((1.5, 1.0, 0.5)))          # dx = +1
    o.translate()
    b = o.download_atlas()
    assert np.array_equal(b[:, :W - 10], a[:, 10:]) and np.array_equal(b[:, W - 10:], a[:, W - 10:])   # last block unchanged (Q7)
    o.set_scene_ubo(s.with_camera((1.5, 1.0, -1.2)))         # dz = int(-1.2) - 0 = -1 (truncation toward zero, Q1)
    o.upload_atlas(a)
    o.translate()
    b = o.download_atlas()
    for layer in range(n_y):
        y0 = layer * 10 * n_xz
        assert np.array_equal(b[y0 + 10:y0 + 10 * n_xz], a[y0:y0 + 10 * n_xz - 10])
        assert np.array_equal(b[y0:y0 + 10], a[y0:y0 + 10])
    o.set_scene_ubo(s.with_camera((1.9, 1.0, -1.9)))         # same integer cell: no-op
    o.upload_atlas(a)
    o.translate()
    assert np.array_equal(o.download_atlas(), a)
    o.set_scene_ubo(s.with_camera((5.5, 1.0, -1.9)))         # |dx| = 4 still shifts by exactly one probe (Q7)
    o.translate()
    b = o.download_atlas()
    assert np.array_equal(b[:, :W - 10], a[:, 10:])
    o.close()


def test_quad_scene_analytic_hits_and_shadow_split(oracle):
    # KAT 9: floor quad at y=-1.5 below a probe at (px, 0, pz): straight-down ray hits at t = 1.5
    s = scenes.quad_scene(grid=(4, 2, 64), second_quad=True)
    p = make_params(4, 2, 64, flags=L.FLAG_RAY_RECORDS)
    o = oracle.Oracle(p)
    o.upload_scene(s)
    o.set_scene_ubo(s.ubo)
    o.build_accel()
    rot = np.eye(4, dtype=np.float32).ravel()
    o.trace_layer(0, rot)
    rec = o.download_ray_records(0)
    fib = oracle.fibonacci_dirs(64)
    down = fib[:, 1] < -0.2
    # with identity rotation the hit distance to the floor plane is -1.5 / d.y for rays that reach the 20 m quad
    for probe in (0, 5, 15):
        r = rec[probe]
        hit_floor = (r["prim"] < 2)
        assert hit_floor[down].mean() > 0.5
        np.testing.assert_allclose(r["t"][hit_floor], (-1.5 / fib[hit_floor, 1]), rtol=1e-5)
    # the second quad (y = 6, x in [-10,0]) shadows the sun (from -x,+y,-z) for part of the floor only
    lit = rec["lit_mask"][rec["prim"] < 2] & 1
    assert 0.05 < lit.mean() < 0.95
    o.close()


def test_bvh_equals_brute_force(oracle):
    for seed in (11, 12):
        s = scenes.random_soup(seed, ntris=1500)
        p = make_params(3, 2, 64, flags=L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS)
        a, b = oracle.Oracle(p), oracle.Oracle(p)
        b.set_brute_force(True)
        for h in (a, b):
            h.upload_scene(s)
            h.set_scene_ubo(s.ubo)
            h.build_accel()
        rots = a.rotations(scenes.rotation_sequence(seed, 2))
        for h in (a, b):
            h.update(-1, rots)
        for layer in range(2):
            ra, rb = a.download_ray_records(layer), b.download_ray_records(layer)
            assert np.array_equal(ra["prim"], rb["prim"]) and np.array_equal(ra["lit_mask"], rb["lit_mask"])
            assert np.array_equal(ra["radiance"], rb["radiance"])
        assert np.array_equal(a.download_atlas(), b.download_atlas())
        a.close()
        b.close()


def test_white_furnace_energy(oracle):
    # KAT 10: closed box, all probes inside, sun off, emissive-free: atlas rgb stays exactly 0
    b = scenes.SceneBuilder("furnace", 0)
    cube = b.add_mesh(*scenes.box(4))
    mat = b.add_material((0.8, 0.8, 0.8))
    b.add_renderable(scenes.translate((0.0, 3.0, 0.0)) @ scenes.scale((30.0, 30.0, 30.0)), cube, mat)
    s = b.finish((0.0, 3.0, 0.0), (4, 2, 64))
    s.ubo["sun_intensity"] = 0.0
    p = make_params(4, 2, 64, flags=L.FLAG_RAY_RECORDS)
    o = oracle.Oracle(p)
    o.upload_scene(s)
    o.set_scene_ubo(s.ubo)
    o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(3, 2))
    for _ in range(3):
        o.update(-1, rots)
    a = o.download_atlas().view(np.float16)
    from tests.parity import interior_mask
    assert np.all(a[interior_mask(4, 2)][:, :3] == 0)
    # with emissive walls and multi-bounce the atlas grows monotonically towards E/(1-rho*c)
    s.materials["emissive"][0] = (1.0, 1.0, 1.0, 0.5)
    o.upload_scene(s)
    o.build_accel()
    o.reset()
    prev = 0.0
    for _ in range(4):
        o.update(-1, rots)
        m = o.download_atlas().view(np.float16)[interior_mask(4, 2)][:, :3].astype(np.float64).mean()
        assert m > prev
        prev = m
    o.close()


def test_skipping_dead_shadow_rays_changes_nothing_but_the_ray_count(oracle):
    """FLAG_SKIP_DEAD_SHADOW_RAYS: the sun shadow ray of rchit:45-66 is not traced where dot(N, L) <= 0.  calcLight
    multiplies by max(N.L, 0) = 0 there (pbr_light.glsl:161), so radiance and atlas must be bit-identical to the
    faithful path; only lit_mask bit 0 (forced to 0) and the number of rays traced differ."""
    for seed in (11, 14):
        s = scenes.random_soup(seed, ntris=1500)
        base = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_COUNTERS
        a, b = oracle.Oracle(make_params(3, 2, 64, flags=base)), oracle.Oracle(make_params(3, 2, 64, flags=base | L.FLAG_SKIP_DEAD_SHADOW_RAYS))
        for h in (a, b):
            h.upload_scene(s)
            h.set_scene_ubo(s.ubo)
            h.build_accel()
        rots = a.rotations(scenes.rotation_sequence(seed, 2))
        for _ in range(3):
            for h in (a, b):
                h.update(-1, rots)
        for layer in range(2):
            ra, rb = a.download_ray_records(layer), b.download_ray_records(layer)
            assert np.array_equal(ra["prim"], rb["prim"]) and np.array_equal(ra["radiance"], rb["radiance"])
            assert np.array_equal(ra["lit_mask"] >> 1, rb["lit_mask"] >> 1)
            assert np.all((rb["lit_mask"] & 1) <= (ra["lit_mask"] & 1))
        assert np.array_equal(a.download_atlas(), b.download_atlas())
        sa, sb = a.stats(), b.stats()
        assert sb["shadow_rays"] < sa["shadow_rays"] and sa["primary_hits"] == sb["primary_hits"]
        a.close()
        b.close()


def test_atlas_consumer_uniform_atlas_known_answer(oracle):
    """KAT for the consumer: with every atlas texel = L, every probe tap returns L, so the sqrt-space blend
    gives irradiance = L * 0.5 * 3.1415926 whatever the weights, and calcIndirectDiffuseLight = kD * irradiance * albedo
    with kD = (1 - kS)(1 - metallic), kS = F0 + (max(1 - r^4... see fresnelSchlickRoughness(cos, F0, roughness^2)."""
    n_xz, n_y = 4, 2
    o = oracle.Oracle(make_params(n_xz, n_y, 64))
    s = scenes.empty_scene(grid=(n_xz, n_y, 64))
    o.set_scene_ubo(s.ubo)
    W, H = 10 * n_xz, 10 * n_xz * n_y
    Lrgb = np.array([0.25, 1.0, 4.0], np.float32)             # exact in fp16, exact square roots
    a = np.zeros((H, W, 4), np.float16); a[..., :3] = Lrgb; a[..., 3] = 1.0
    o.upload_atlas(a.view(np.uint16))
    pts = np.zeros((3, 12), np.float32)
    pts[:, 0:3] = [(0.3, 1.1, -0.4), (-1.2, 0.5, 1.7), (0.9, 2.5, 0.1)]        # inside the grid: taps never leave it
    pts[:, 3:6] = [(0, 1, 0), (0.6, 0.0, 0.8), (-1, 0, 0)]
    pts[:, 6:9] = [(0.8, 0.5, 0.2), (1, 1, 1), (0.3, 0.3, 0.3)]
    pts[:, 9] = [0.0, 0.0, 0.5]; pts[:, 10] = [1.0, 0.5, 0.3]
    view = np.array([0.0, 5.0, 0.0], np.float32)
    for clamp in (False, True):
        got = o.indirect_diffuse(pts, view, clamp)[:, :3]
        for i in range(3):
            pos, n, alb, met, rough = pts[i, 0:3], pts[i, 3:6], pts[i, 6:9], pts[i, 9], pts[i, 10]
            V = (view - pos) / np.linalg.norm(view - pos)
            F0 = 0.04 * (1 - met) + alb * met
            cos = max(float(np.dot(n, V)), 0.0)
            kS = F0 + (np.maximum(1.0 - rough * rough, F0) - F0) * (1.0 - cos) ** 5
            want = (1 - kS) * (1 - met) * (Lrgb * 0.5 * 3.1415926) * alb
            assert np.allclose(got[i], want, rtol=2e-5, atol=1e-7), (i, got[i], want)
    o.close()


def _srgb_to_linear(c):
    v = np.float32(c) / np.float32(255.0)
    return v / np.float32(12.92) if v < np.float32(0.04045) else np.float32(((np.float64(v) + 0.055) / 1.055) ** 2.4)


def test_texture_sampling_known_answers(oracle):
    """KATs for texture(textures[idx], uv) as defined in DESIGN.md: LOD 0, bilinear with texel centres at (i + 0.5) / size,
    REPEAT vs CLAMP_TO_EDGE, UNORM8 = c/255, sRGB8 decoded before filtering, alpha linear, RG8 -> (r, g, 0, 1)."""
    o = oracle.Oracle(make_params(4, 2, 64))
    px = np.zeros((2, 4, 4), np.uint8)
    px[0, :, 0] = (0, 51, 102, 255); px[1, :, 0] = (255, 204, 153, 0); px[..., 1] = 128; px[..., 3] = (64, 255, 0, 128)
    rg = np.array([[[10, 20], [30, 40]]], np.uint8)
    o.upload_textures([(px, L.TEX_RGBA8_UNORM, False), (px, L.TEX_RGBA8_UNORM, True), (px, L.TEX_RGBA8_SRGB, False), (rg, L.TEX_RG8_UNORM, False)])
    # texel centres return the texel
    for (i, j) in ((0, 0), (2, 1), (3, 0)):
        got = o.sample_texture(0, (i + 0.5) / 4, (j + 0.5) / 2)
        assert np.allclose(got, px[j, i] / 255.0, atol=1e-7)
    # halfway between texel (1,0) and (2,0): the mean
    got = o.sample_texture(0, 2.0 / 4, 0.5 / 2)
    assert np.allclose(got[0], (51 + 102) / 2 / 255.0, atol=1e-6)
    # left of the first texel centre: REPEAT blends with the last column, CLAMP_TO_EDGE returns the first
    rep, clp = o.sample_texture(0, 0.0, 0.25), o.sample_texture(1, 0.0, 0.25)
    assert np.allclose(rep[0], (0 + 255) / 2 / 255.0, atol=1e-6) and np.allclose(clp[0], 0.0, atol=1e-7)
    assert np.allclose(o.sample_texture(0, 1.625, 0.25), o.sample_texture(0, 0.625, 0.25), atol=1e-6)     # period 1
    assert np.allclose(o.sample_texture(0, -0.375, 0.25), o.sample_texture(0, 0.625, 0.25), atol=1e-6)
    # sRGB: decode, then filter; alpha is not decoded
    got = o.sample_texture(2, 1.5 / 4, 0.5 / 2)
    assert np.allclose(got[0], _srgb_to_linear(51), rtol=1e-6) and np.allclose(got[1], _srgb_to_linear(128), rtol=1e-6) and np.allclose(got[3], 1.0)
    mid = o.sample_texture(2, 2.0 / 4, 0.5 / 2)
    assert np.allclose(mid[0], (_srgb_to_linear(51) + _srgb_to_linear(102)) / 2, rtol=1e-6)
    assert np.allclose(o.sample_texture(3, 0.25, 0.5), (10 / 255.0, 20 / 255.0, 0.0, 1.0), atol=1e-7)
    o.close()


def test_constant_texture_equals_the_factor_material(oracle):
    """A 1x1 sRGB albedo texture with unit base colour factor shades exactly like the factor-only material holding the
    decoded colour; a constant metallic-roughness texture like the factors it holds."""
    def build(textured):
        b = scenes.SceneBuilder("kat_tex", 3)
        pos, nrm, idx = scenes.box(3)
        if textured:
            ta = b.add_texture(np.array([[[200, 100, 50, 255]]], np.uint8), L.TEX_RGBA8_SRGB)
            tm = b.add_texture(np.array([[[102, 51, 0, 255]]], np.uint8), L.TEX_RGBA8_UNORM)
            mat = b.add_material((1.0, 1.0, 1.0), 0.9, 0.1, tex=(tm, ta, -1, -1))
        else:
            mat = b.add_material(tuple(float(_srgb_to_linear(c)) for c in (200, 100, 50)), 51 / 255.0, 102 / 255.0)
        mesh = b.add_mesh(pos, nrm, idx, uv=np.random.RandomState(1).uniform(-3, 3, (len(pos), 2)).astype(np.float32))
        b.add_renderable(scenes.translate((0.2, 2.0, 0.1)) @ scenes.scale(5.0), mesh, mat)
        b.add_light((0.0, 2.0, 0.0), 6.0, (3.0, 3.0, 3.0))
        return b.finish((0.3, 2.0, -0.2), (4, 3, 64))
    atl = []
    for textured in (True, False):
        s = build(textured)
        o = oracle.Oracle(make_params(4, 3, 64, flags=L.FLAG_POINT_LIGHTS))
        o.upload_scene(s); o.set_scene_ubo(s.ubo); o.build_accel()
        rots = o.rotations(scenes.rotation_sequence(3, 3))
        for _ in range(2):
            o.update(-1, rots)
        atl.append(o.download_atlas().view(np.float16).astype(np.float32))
        o.close()
    assert atl[0][..., :3].max() > 0.01
    # the factor material stores the decoded colour rounded to fp32 once more than the texture path: allow a few fp16 ulps
    assert np.allclose(atl[0], atl[1], rtol=4e-3, atol=1e-5)


def test_shadow_rt_known_answer(oracle):
    """KAT for ShadowRT: a horizontal quad at y = 2 over x, z in [-1, 1], light straight down.  Seen through a top-down
    orthographic inverse view-projection, pixels whose world position lies under the quad (y < 2 - tMin) are shadowed, pixels
    above it or outside its footprint are lit, sky pixels (depth == 1) are lit without a ray."""
    from tests import shadow_inputs as SI
    b = scenes.SceneBuilder("kat_shadow", 1)
    pos = np.array([(-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1)], np.float32)
    mesh = b.add_mesh(pos, np.tile(np.array([0, 1, 0], np.float32), (4, 1)), np.array([0, 1, 2, 0, 2, 3], np.uint32))
    b.add_renderable(np.eye(4), mesh, b.add_material((0.5, 0.5, 0.5)))
    s = b.finish((0.0, 1.0, 0.0), (4, 2, 64))
    o = oracle.Oracle(make_params(4, 2, 64))
    o.upload_scene(s); o.build_accel()
    W, H = 16, 16
    ubo = SI.with_shadow_camera(s.ubo, SI.ortho_inv_view_proj(4.0, 4.0, 4.0), W, H)
    ubo["light_dir"][0] = (0.0, -1.0, 0.0, 0.0)
    o.set_scene_ubo(ubo)
    depth = np.full((H, W), 0.25, np.float32)          # world y = 1
    depth[0, :] = 0.75                                  # world y = 3: above the quad
    depth[1, 0] = 1.0                                   # sky
    got = o.shadow_rt(depth)
    xs = 4.0 * ((2 * np.arange(W) + 1) / (2 * W) * 2 - 1)
    inside = np.abs(xs) < 1.0
    want = np.ones((H, W), np.float32)
    want[np.ix_(inside, inside)] = 0.0                  # rows = z, columns = x
    want[0, :] = 1.0
    assert np.array_equal(got, want)
    o.close()


def test_bake_mode_running_mean_and_tile_grid(oracle):
    """KATs for bake mode.  (a) probe_conv.comp:67-79: with uniform radiance L on all rays, k convolutions from
    the cleared atlas (0,0,0,1) give alpha = 1 + k*S and rgb = L * k*S / (1 + k*S), S = sum of the texel's weights (the cleared
    texel counts as one black sample of weight 1).  (b) rgen:61-73: the grid covers tile bakeTileInfo.xy of size bakeTileInfo.z
    with step (tSize/nXZ, 1, tSize/nXZ): in an empty scene nothing depends on it, with one quad it decides which probes see it."""
    o, dir_img, _ = _conv_setup(oracle, 4, 2, 256, seed=1)
    ubo = scenes.empty_scene(grid=(4, 2, 256)).ubo.copy()
    ubo["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
    ubo["bake_tile_info"][0] = (0, 0, 8, 0)
    o.set_scene_ubo(ubo)
    L0 = np.array([0.75, 1.5, 0.25], np.float16)
    irr = np.zeros_like(dir_img); irr[..., :3] = L0; irr[..., 3] = 1.0
    for layer in range(2):
        o.set_rays(layer, irr.view(np.uint16), dir_img.view(np.uint16))
    from tests import parity as _p
    mask = _p.interior_mask(4, 2)
    S = None
    for k in range(1, 5):
        for layer in range(2):
            o.convolve_layer(layer)
        rec = o.download_texel_records()
        alpha = rec[..., 3][mask]
        if S is None:
            S = alpha - 1.0
            assert np.all((S > 55) & (S < 75))                    # ~N/4 = 64
        np.testing.assert_allclose(alpha, 1.0 + k * S, rtol=3e-3)  # alpha is re-rounded to fp16 every update
        rgb = rec[..., :3][mask]
        want = L0.astype(np.float64)[None, :] * (k * S / (1.0 + k * S))[:, None]
        np.testing.assert_allclose(rgb, want, rtol=3e-3)
    o.close()
    # (b) probe origins: quad straight above tile (1, 0) of size 4 -> only probes of a handle baking that tile see it overhead
    b = scenes.SceneBuilder("kat_bake", 1)
    pos = np.array([(4, 9, 0), (8, 9, 0), (8, 9, 4), (4, 9, 4)], np.float32)
    mesh = b.add_mesh(pos, np.tile(np.array([0, -1, 0], np.float32), (4, 1)), np.array([0, 1, 2, 0, 2, 3], np.uint32))
    b.add_renderable(np.eye(4), mesh, b.add_material((0.5, 0.5, 0.5), emissive=(1.0, 1.0, 1.0, 3.0)))
    s = b.finish((100.0, 1.0, 100.0), (4, 2, 64))                  # the camera is far away: bake mode must not use it
    hits = {}
    for tile in ((1, 0), (5, 5)):
        oo = oracle.Oracle(make_params(4, 2, 64, flags=L.FLAG_RAY_RECORDS))
        oo.upload_scene(s); oo.build_accel()
        u = s.ubo.copy(); u["flags"] |= L.UBO_BAKE_MODE_ON_FLAG; u["bake_tile_info"][0] = (tile[0], tile[1], 4, 0)
        oo.set_scene_ubo(u)
        oo.translate()                                              # no-op in bake mode (IrradianceProbeTranslationPass.cpp:58-59)
        oo.update(-1, oo.rotations(scenes.rotation_sequence(1, 2)))
        hits[tile] = sum(int((oo.download_ray_records(l)["prim"] != 0xFFFFFFFF).sum()) for l in range(2))
        oo.close()
    assert hits[(1, 0)] > 10 and hits[(5, 5)] == 0        # ~1.6 % of 2048 rays point at the quad from 8-9 m below


def test_specular_gi_known_answer(oracle):
    """KAT for SpecularGIRT: a mirror-like floor point at the origin seen from a camera at (0, 3, -3) reflects towards (0, 3, 3)
    (angle of incidence = angle of reflection).  An emissive quad placed there, facing the floor, with the sun and DDGI off,
    gives irradiance = emissive and distance = |(0,3,3)| = 4.2426; a pixel whose reflection leaves the scene gives (sky, inf)."""
    from tests import shadow_inputs as SI
    b = scenes.SceneBuilder("kat_spec", 1)
    c = np.array([0.0, 3.0, 3.0]); n = -c / np.linalg.norm(c)
    t1 = np.cross(n, [1.0, 0.0, 0.0]); t1 /= np.linalg.norm(t1); t2 = np.cross(n, t1)
    pos = np.array([c - t1 - t2, c + t1 - t2, c + t1 + t2, c - t1 + t2], np.float32)
    mesh = b.add_mesh(pos, np.tile(n.astype(np.float32), (4, 1)), np.array([0, 1, 2, 0, 2, 3], np.uint32))
    b.add_renderable(np.eye(4), mesh, b.add_material((0.5, 0.5, 0.5), emissive=(0.5, 1.0, 2.0, 1.5)))
    s = b.finish((0.0, 3.0, -3.0), (4, 2, 64))
    o = oracle.Oracle(make_params(4, 2, 64))
    o.upload_scene(s); o.build_accel()
    W = H = 2
    ivp = np.zeros((4, 4)); ivp[0, 0], ivp[2, 1], ivp[1, 2], ivp[3, 3] = 1e-3, 1e-3, 1.0, 1.0      # every pixel ~ world (0, depth, 0)
    u = SI.with_specular_camera(s.ubo, ivp.T.astype(np.float32).ravel(), W, H)
    u["flags"] &= ~np.uint32(L.UBO_DDGI_FLAG | L.UBO_DDGI_MULTI_FLAG)
    u["sun_intensity"] = 0.0
    o.set_scene_ubo(u)
    depth = np.zeros((H, W), np.float32)                     # world y = 0: the floor point
    geom0 = np.zeros((H, W, 4), np.float32); geom0[..., 1] = 2.0          # normal (0, 1, 0), un-normalised
    geom0[1, 1, :3] = (0.0, 0.0, -1.0)                       # this pixel's reflection goes up and back: leaves the scene
    depth[0, 1] = 1.0                                        # sky pixel
    got = o.specular_gi(depth, geom0).view(np.float16).astype(np.float32)
    want_e = np.array([0.5, 1.0, 2.0]) * 1.5
    for (y, x) in ((0, 0), (1, 0)):
        assert np.allclose(got[y, x, :3], want_e, rtol=2e-3) and abs(got[y, x, 3] - np.sqrt(18.0)) < 0.01
    assert (got[0, 1] == 0).all()
    assert np.isinf(got[1, 1, 3]) and np.allclose(got[1, 1, :3], (0.033105, 0.214041, 0.447988), rtol=2e-3)
    o.close()


def _cook_torrance_f64(N, albedo, metallic, roughness, pos, view, L, radiance):
    """pbr_light.glsl:98-166 restated in float64 from SURVEY.md Appendix A.7 (independent of the oracle's code)."""
    N, albedo, pos, view, L, radiance = (np.asarray(v, np.float64) for v in (N, albedo, pos, view, L, radiance))
    V = (view - pos) / np.linalg.norm(view - pos)
    F0 = 0.04 * (1 - metallic) + albedo * metallic
    H = (V + L) / np.linalg.norm(V + L)
    a2 = (roughness * roughness) ** 2
    NdotH = max(N @ H, 0.0)
    NDF = (a2 + 0.001) / (3.14159265359 * (NdotH * NdotH * (a2 - 1) + 1) ** 2)
    k = (roughness + 1) ** 2 / 8
    NdotV, NdotL = max(N @ V, 0.0), max(N @ L, 0.0)
    G = (NdotL / (NdotL * (1 - k) + k)) * (NdotV / (NdotV * (1 - k) + k))
    F = F0 + (1 - F0) * np.clip(1 - max(H @ V, 0.0), 0, 1) ** 5
    kD = (1 - F) * (1 - metallic)
    spec = np.clip(NDF * G * F / (4 * NdotV * NdotL + 1e-4), 0, 1)
    return (kD * albedo / 3.14159265359 + spec) * radiance * NdotL


def test_calc_light_known_answers(oracle):
    """KATs for calcLight (pbr_light.glsl:98-166): the sun (radiance = sunIntensity) and a point light with the (1 - d/range)^2
    falloff (:132-138), against a float64 restatement of the formulas; plus the closed form at normal incidence."""
    rng = np.random.RandomState(3)
    for it in range(40):
        N = rng.normal(size=3); N /= np.linalg.norm(N)
        pos, view = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3)
        albedo, metallic, rough = rng.uniform(0.05, 0.95, 3), float(rng.rand() < 0.3), rng.uniform(0.15, 1.0)
        ld = rng.normal(size=3); ld /= np.linalg.norm(ld)
        got = oracle.calc_light(N, albedo, metallic, rough, pos, view, ld, 5.0)
        want = _cook_torrance_f64(N, albedo, metallic, rough, pos, view, -ld, np.full(3, 5.0))
        assert np.allclose(got, want, rtol=3e-4, atol=2e-6), (it, got, want)
        light = np.zeros(1, L.LIGHT); lp = pos + rng.uniform(-4, 4, 3); rng_m = 9.0
        light["world_pos"][0] = (*lp, rng_m); light["color"][0] = (2.0, 3.0, 4.0, 1.0)
        d = min(np.linalg.norm(lp - pos), rng_m)
        got = oracle.calc_light(N, albedo, metallic, rough, pos, view, ld, 5.0, light=light)
        want = _cook_torrance_f64(N, albedo, metallic, rough, pos, view, (lp - pos) / np.linalg.norm(lp - pos), np.array([2.0, 3.0, 4.0]) * (1 - d / rng_m) ** 2)
        assert np.allclose(got, want, rtol=3e-4, atol=2e-6), (it, got, want)
    # normal incidence, dielectric, roughness 1: N = V = L -> NdotH = 1, NDF = 1.001/pi, G = 1, F = F0 = 0.04
    got = oracle.calc_light((0, 1, 0), (0.5, 0.5, 0.5), 0.0, 1.0, (0, 0, 0), (0, 2, 0), (0, -1, 0), 3.0)
    spec = min(1.001 / 3.14159265359 * 0.04 / (4 + 1e-4), 1.0)
    assert np.allclose(got, ((1 - 0.04) * 0.5 / 3.14159265359 + spec) * 3.0, rtol=1e-5)
    # light below the horizon: NdotL = 0 -> exactly zero
    assert (oracle.calc_light((0, 1, 0), (0.5, 0.5, 0.5), 0.0, 0.5, (0, 0, 0), (1, 2, 0), (0, 1, 0), 3.0) == 0).all()


def _indirect_f64(atlas, nxz, ny, cam, pos, n, albedo, metallic, rough, view, clamp):
    """calcIndirectDiffuseLight restated in float64 from SURVEY.md Appendix A.8 (sampleProbe, weightProbe, sampleSingleProbe,
    octEncode, fresnelSchlickRoughness), independent of the oracle's code."""
    H, W = atlas.shape[:2]
    step = np.array([1.0, 2.0, 1.0])
    camoff = np.array([int(cam[0] / 1.0) * 1.0 - nxz / 2.0, 0.0, int(cam[2] / 1.0) * 1.0 - nxz / 2.0])
    pos, n, albedo, view = (np.asarray(v, np.float64) for v in (pos, n, albedo, view))
    V = (view - pos) / np.linalg.norm(view - pos)
    F0 = 0.04 * (1 - metallic) + albedo * metallic
    r2 = rough * rough
    kS = F0 + (np.maximum(1 - r2, F0) - F0) * np.clip(1 - max(n @ V, 0.0), 0, 1) ** 5
    kD = (1 - kS) * (1 - metallic)
    p = np.array([np.clip(np.floor((pos[0] - camoff[0]) / 1.0), 0, nxz - 1), np.clip(np.floor(pos[1] / 2.0), 0, ny - 1),
                  np.clip(np.floor((pos[2] - camoff[2]) / 1.0), 0, nxz - 1)])
    base = p * step + camoff
    alpha = np.clip((pos - base) / step, 0, 1)
    nn = n / np.linalg.norm(n)
    v = np.array([nn[0], nn[2], nn[1]])                                  # octEncode swaps y and z first
    v = v / np.abs(v).sum()
    o = v[:2].copy()
    if v[2] < 0:
        sgn = np.where(o >= 0, 1.0, -1.0)
        o = (1 - np.abs(o[::-1])) * sgn
    o = o * 0.5 + 0.5

    def texel(x, y):
        if clamp:
            x, y = min(max(x, 0), W - 1), min(max(y, 0), H - 1)
        return atlas[y % H, x % W, :3].astype(np.float64)

    sw, sm = 0.0, np.zeros(3)
    for off in ((0, 0, 0), (1, 1, 1), (1, 1, 0), (1, 0, 0), (0, 1, 1), (0, 1, 0), (0, 0, 1), (1, 0, 1)):
        off = np.array(off)
        tri = np.where(off == 1, alpha, 1 - alpha)
        dirp = base + off * step - pos
        dirp = dirp / np.linalg.norm(dirp)
        w = max(1e-6, max(1e-4, (dirp @ n + 1) / 2) ** 2 + 0.2) * tri.prod()
        q = (p + off).astype(int)
        sx, sy = 10 * q[0] + 1, 10 * nxz * q[1] + 10 * q[2] + 1
        tsx, tex = (2 * sx + 1) / (2 * W), (2 * (sx + 7) + 1) / (2 * W)
        tsy, tey = (2 * sy + 1) / (2 * H), (2 * (sy + 7) + 1) / (2 * H)
        x, y = (tsx + (tex - tsx) * o[0]) * W - 0.5, (tsy + (tey - tsy) * o[1]) * H - 0.5
        ix, iy = int(np.floor(x)), int(np.floor(y)); fx, fy = x - ix, y - iy
        c = (texel(ix, iy) * (1 - fx) * (1 - fy) + texel(ix + 1, iy) * fx * (1 - fy) + texel(ix, iy + 1) * (1 - fx) * fy + texel(ix + 1, iy + 1) * fx * fy)
        sm += w * np.sqrt(c); sw += w
    irr = (sm / sw) ** 2 * 0.5 * 3.1415926
    return kD * irr * albedo


def test_indirect_diffuse_against_a_float64_restatement(oracle):
    """Pins sampleProbe / weightProbe / sampleSingleProbe / octEncode (pbr_light.glsl:168-388) on a random atlas: the oracle
    against an independent float64 restatement, REPEAT and CLAMP_TO_EDGE samplers, points inside and outside the grid."""
    from tests.consumer_points import surface_points
    nxz, ny = 4, 2
    o = oracle.Oracle(make_params(nxz, ny, 64))
    s = scenes.empty_scene(grid=(nxz, ny, 64))
    cam = (0.3, 1.0, -0.6)
    o.set_scene_ubo(s.with_camera(cam))
    rng = np.random.RandomState(8)
    a = np.zeros((10 * nxz * ny, 10 * nxz, 4), np.float16)
    a[..., :3] = rng.uniform(0.0, 2.0, a.shape[:2] + (3,)); a[..., 3] = 1.0
    o.upload_atlas(a.view(np.uint16))
    pts = surface_points(120, nxz, ny, seed=5, camera=cam)
    view = np.array([0.2, 3.0, 0.1])
    for clamp in (False, True):
        got = o.indirect_diffuse(pts, view, clamp)[:, :3]
        want = np.array([_indirect_f64(a.astype(np.float32), nxz, ny, cam, q[0:3], q[3:6], q[6:9], q[9], q[10], view, clamp) for q in pts])
        close = np.isclose(got, want, rtol=2e-3, atol=1e-5).all(axis=1)
        assert close.mean() >= 0.97, (clamp, close.mean(), got[~close][:2], want[~close][:2])    # a texel-boundary tap may land one texel off in float64
    o.close()


def test_tile_lookup_and_baked_atlas_known_answers(oracle):
    """KATs for the tile page (bindless.glsl:4-5, 291-308) and the baked-atlas branch of sampleProbe (pbr_light.glsl:251-273):
    a uniform baked atlas B on one tile and a uniform live atlas A.  The deferred-pass consumer returns kD*albedo*B*pi/2 for
    points on that tile and kD*albedo*A*pi/2 elsewhere; the REPEAT-sampler consumer (SpecularGIRT) always reads the live atlas."""
    n_xz, n_y = 4, 2
    o = oracle.Oracle(make_params(n_xz, n_y, 64))
    s = scenes.empty_scene(grid=(n_xz, n_y, 64))
    W, H = 10 * n_xz, 10 * n_xz * n_y
    def uniform(v):
        a = np.zeros((H, W, 4), np.float16); a[..., :3] = v; a[..., 3] = 1.0
        return a
    A, B = 0.25, 4.0
    o.upload_textures([(uniform(B).view(np.uint8).reshape(H, W, 8), L.TEX_RGBA16F, False)])
    o.set_scene_ubo(s.with_camera((40.0, 1.0, 10.0)))                # camera tile (1, 0)
    o.upload_atlas(uniform(A).view(np.uint16))
    tiles = np.full(21 * 21, -1, np.int32)
    tiles[(0 - 0 + 10) * 21 + (2 - 1 + 10)] = 0                       # world tile (2, 0): x in [64, 96), z in [0, 32)
    o.upload_tile_infos(tiles)
    pts = np.zeros((3, 12), np.float32)
    pts[:, 0:3] = [(70.0, 1.0, 5.0), (40.5, 1.0, 10.5), (64.0 - 1e-3, 1.0, 5.0)]       # on the baked tile, on the camera's tile, just outside
    pts[:, 3:6] = (0, 1, 0); pts[:, 6:9] = 0.5; pts[:, 10] = 1.0
    view = np.array([40.0, 6.0, 10.0], np.float32)
    def want(i, Lv):
        V = (view - pts[i, 0:3]) / np.linalg.norm(view - pts[i, 0:3])
        kS = 0.04 + (max(1.0 - 1.0, 0.04) - 0.04) * (1.0 - max(float(V[1]), 0.0)) ** 5
        return (1 - kS) * Lv * 0.5 * 3.1415926 * 0.5
    got = o.indirect_diffuse(pts, view, True)[:, 0]
    assert np.allclose(got, [want(0, B), want(1, A), want(2, A)], rtol=2e-3)
    got = o.indirect_diffuse(pts, view, False)[:, 0]
    assert np.allclose(got, [want(0, A), want(1, A), want(2, A)], rtol=2e-3)
    o.close()
