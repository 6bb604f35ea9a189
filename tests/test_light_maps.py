"""Light maps (csrc/ddgi_kernels.cuh "Light maps"): conservative distance cube maps around the point lights that answer
"certainly unoccluded" shadow rays without a traversal.  They must never change a result: lit masks, radiance and atlas are compared
bit for bit with the oracle (which traces every shadow ray) and with the same handle built with FLAG_NO_LIGHT_MAPS — on scenes made to
stress the conservative footprints (many lights inside triangle soups, lights close to geometry, lights on a face diagonal, moving
instances, deforming meshes) — and they must actually answer rays (stats.shadow_rays_prefiltered > 0).

CPU: the product's host-emu harness (same raster / look-up code, serial).  -m gpu: the CUDA kernels through the C ABI."""
import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

FLAGS = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_COUNTERS | L.FLAG_SKIP_DEAD_SHADOW_RAYS


def _lights_everywhere(scene, seed, n=12, near_geometry=False):
    """Replaces the scene's lights by `n` enabled ones spread through it (some on exact axis / diagonal positions relative to
    geometry so that rays fall on cube-face borders)."""
    rng = np.random.RandomState(seed)
    lights = np.zeros(n, scene.lights.dtype)
    for k in range(n):
        pos = rng.uniform(-2.5, 2.5, 3) + (0, 2, 0)
        if k % 4 == 1:
            pos = np.round(pos)                       # integer positions: probe-grid aligned, many rays exactly on face diagonals
        lights[k]["world_pos"] = (*pos, rng.uniform(3, 9))
        lights[k]["color"] = (*rng.uniform(1, 5, 3), 1.0 if k % 5 != 4 else 0.0)     # some disabled
    if near_geometry:
        # put two lights a few millimetres / centimetres off a triangle of the first renderable's mesh
        v = scene.vertices["pos"][scene.indices[:3]]
        M = scene.renderables["transform"][0].reshape(4, 4).T
        w = (np.c_[v, np.ones(3)] @ M.T)[:, :3]
        n_ = np.cross(w[1] - w[0], w[2] - w[0]); n_ /= max(np.linalg.norm(n_), 1e-12)
        lights[0]["world_pos"][:3] = w.mean(0) + 0.004 * n_
        lights[2]["world_pos"][:3] = w.mean(0) - 0.05 * n_
    scene.lights = lights
    return scene


def _run_pair(scene, grid, make_a, make_b, updates=1, flags_b=FLAGS, exact=True):
    """`exact`: radiance and atlas bit for bit (same implementation on both sides, or scenes without vertex-colour `pow`);
    otherwise the parity bar of the other GPU tests (libm vs CUDA `powf` ulps: 1e-4 relative, atlas within one half ulp)."""
    nxz, ny, nr = grid
    a, b = make_a(make_params(nxz, ny, nr, flags=FLAGS)), make_b(make_params(nxz, ny, nr, flags=flags_b))
    rots = parity.setup_pair(scene, (a, b), seed=7)
    for _ in range(updates):
        for h in (a, b):
            h.update(-1, rots)
    for layer in range(ny):
        ra, rb = a.download_ray_records(layer), b.download_ray_records(layer)
        assert np.array_equal(ra["prim"], rb["prim"])
        assert np.array_equal(ra["lit_mask"], rb["lit_mask"]), "a light map answered a ray the traversal finds occluded (or the reverse)"
        if exact:
            assert np.array_equal(ra["radiance"].view(np.uint32), rb["radiance"].view(np.uint32))
        else:
            parity.compare_records(ra, rb)
    if exact:
        assert np.array_equal(a.download_atlas(), b.download_atlas())
    else:
        parity.compare_atlas(a.download_atlas(), b.download_atlas())
    return a, b


def _with(p, flags):
    p.flags |= flags
    return p


CASES = {
    "soup_12_lights": lambda: (_lights_everywhere(scenes.random_soup(31, ntris=1500), 1), (4, 3, 64)),
    "soup_dense_24_lights": lambda: (_lights_everywhere(scenes.random_soup(32, ntris=4000, extent=4.0), 2, n=24), (4, 2, 64)),
    "soup_lights_touching_geometry": lambda: (_lights_everywhere(scenes.random_soup(33, ntris=1200), 3, near_geometry=True), (4, 3, 64)),
    "cornell": lambda: (scenes.cornell_box(grid=(8, 4, 64), wall_res=10, box_res=3), (8, 4, 64)),
    "sponza_small": lambda: (scenes.sponza_like(seed=8, grid=(6, 3, 64), curtain_res=8), (6, 3, 64)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_light_maps_never_change_a_result(name, oracle, make_emu):
    scene, grid = CASES[name]()
    o, e = _run_pair(scene, grid, oracle.Oracle, make_emu, updates=2)
    so, se = o.stats(), e.stats()
    assert se["shadow_rays"] + se["shadow_rays_prefiltered"] == so["shadow_rays"], "every shadow ray is either traced or answered by a map"
    if name != "soup_lights_touching_geometry":
        assert se["shadow_rays_prefiltered"] > 0, "the maps answered nothing: the test does not exercise them"
    print(f"{name}: {se['shadow_rays_prefiltered']} of {so['shadow_rays']} shadow rays answered by the light maps")
    o.close()
    e.close()


def test_no_light_maps_flag_traces_everything(oracle, make_emu):
    scene, grid = CASES["soup_12_lights"]()
    a, b = _run_pair(scene, grid, make_emu, make_emu, flags_b=FLAGS | L.FLAG_NO_LIGHT_MAPS)
    sa, sb = a.stats(), b.stats()
    assert sb["shadow_rays_prefiltered"] == 0 and sa["shadow_rays_prefiltered"] > 0
    assert sa["shadow_rays"] + sa["shadow_rays_prefiltered"] == sb["shadow_rays"]
    a.close()
    b.close()


def _moving_sequence(scene, grid, make_o, make_p, steps=4):
    nxz, ny, nr = grid
    p = make_params(nxz, ny, nr, flags=FLAGS)
    o, g = make_o(p), make_p(p)
    rots = parity.setup_pair(scene, (o, g), seed=9)
    pre = 0
    for k in range(steps):
        rend = scene.animated_renderables(0.35 * (k + 1))
        for h in (o, g):
            h.refit_accel(rend)
            h.update(-1, rots)
        for layer in range(ny):
            ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
            assert np.array_equal(ro["prim"], rg["prim"]) and np.array_equal(ro["lit_mask"], rg["lit_mask"]), f"step {k}: movers and the dynamic map disagree"
        assert np.array_equal(o.download_atlas(), g.download_atlas())
        pre = g.stats()["shadow_rays_prefiltered"]
    assert pre > 0
    # back to the pose of the build: the dynamic map empties again, results still equal
    for h in (o, g):
        h.refit_accel(scene.renderables)
        h.update(-1, rots)
    assert np.array_equal(o.download_atlas(), g.download_atlas())
    o.close()
    g.close()


def test_moving_instances_go_through_the_dynamic_map(oracle, make_emu):
    scene = scenes.sponza_like(seed=4, grid=(6, 3, 64), movers=6, curtain_res=6)
    _moving_sequence(scene, (6, 3, 64), oracle.Oracle, make_emu)


def test_texel_bound_is_below_the_distance_of_every_point_of_the_triangle():
    """Geometry of the bound itself, in float64 against a dense sampling: for random triangles and lights, every sampled point of the
    triangle lies at or beyond the bound of the texel its direction falls in (restated from the header's formulas)."""
    rng = np.random.RandomState(5)
    R = 64
    rho = 1.4143 / R * 1.02 + 2e-5
    worst = np.inf
    for _ in range(300):
        c = rng.uniform(-3, 3, 3)
        tri = c + rng.uniform(-0.8, 0.8, (3, 3))
        n = np.cross(tri[1] - tri[0], tri[2] - tri[0])
        if np.linalg.norm(n) < 1e-6:
            continue
        n /= np.linalg.norm(n)
        h = abs(n @ tri[0])
        bc = rng.dirichlet((1, 1, 1), 400)
        pts = bc @ tri
        d = np.linalg.norm(pts, axis=1)
        a = np.abs(pts).argmax(1)
        for ax in range(3):
            m = a == ax
            if not m.any():
                continue
            q = pts[m]
            w = np.abs(q[:, ax]); s = np.sign(q[:, ax])
            u, v = q[:, (ax + 1) % 3] / w, q[:, (ax + 2) % 3] / w
            X, Y = np.clip(np.floor((u + 1) * 0.5 * R), 0, R - 1), np.clip(np.floor((v + 1) * 0.5 * R), 0, R - 1)
            cu, cv = (X + 0.5) * 2 / R - 1, (Y + 0.5) * 2 / R - 1
            inv = 1 / np.sqrt(1 + cu * cu + cv * cv)
            cosn = np.abs(n[ax] * s + n[(ax + 1) % 3] * cu + n[(ax + 2) % 3] * cv) * inv
            bound = h / np.minimum(1.0, cosn + rho)
            worst = min(worst, float((d[m] - bound).min()))
    assert worst >= -1e-9, f"a triangle point lies {-worst} nearer to the light than its texel's bound"


# ------------------------------------------------------------------------------------------------------------ CUDA
@pytest.mark.gpu
@pytest.mark.parametrize("kernels", ["wavefront", "simple"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_light_maps_never_change_a_result(name, kernels, oracle, make_gpu):
    scene, grid = CASES[name]()

    def make(p):
        p.flags |= L.FLAG_SIMPLE_KERNELS if kernels == "simple" else 0
        return make_gpu(p)
    o, g = _run_pair(scene, grid, oracle.Oracle, make, updates=2, exact=False)
    # the same CUDA handle with the maps off: bit for bit
    a, b = _run_pair(scene, grid, make, lambda p: make_gpu(_with(p, L.FLAG_NO_LIGHT_MAPS | (L.FLAG_SIMPLE_KERNELS if kernels == "simple" else 0))), updates=2)
    a.close()
    b.close()
    so, sg = o.stats(), g.stats()
    assert sg["shadow_rays"] + sg["shadow_rays_prefiltered"] == so["shadow_rays"]
    if name != "soup_lights_touching_geometry":
        assert sg["shadow_rays_prefiltered"] > 0
    o.close()
    g.close()


@pytest.mark.gpu
def test_cuda_moving_instances_go_through_the_dynamic_map(oracle, make_gpu):
    scene = scenes.sponza_like(seed=4, grid=(6, 3, 64), movers=6, curtain_res=6)
    _moving_sequence(scene, (6, 3, 64), oracle.Oracle, make_gpu)


@pytest.mark.gpu
def test_cuda_maps_equal_the_serial_rasteriser(make_gpu, make_emu):
    """The warp-cooperative rasteriser and the serial one of the harness fill the same maps: same rays answered."""
    scene, grid = CASES["soup_dense_24_lights"]()
    e, g = _run_pair(scene, grid, make_emu, make_gpu, exact=False)
    assert e.stats()["shadow_rays_prefiltered"] == g.stats()["shadow_rays_prefiltered"]
    e.close()
    g.close()


# ------------------------------------------------------------------------------------------------------------ the sun's map
def _sun_sequence(scene, grid, make_o, make_p, dirs, move=False):
    """Updates under a changing sun direction (the map is re-made for each), optionally with moving instances in between."""
    nxz, ny, nr = grid
    p = make_params(nxz, ny, nr, flags=FLAGS)
    o, g = make_o(p), make_p(p)
    rots = parity.setup_pair(scene, (o, g), seed=11)
    for k, d in enumerate(dirs):
        ubo = scene.ubo.copy()
        ubo["light_dir"][0][:3] = d
        rend = scene.animated_renderables(0.3 * (k + 1)) if move else None
        for h in (o, g):
            h.set_scene_ubo(ubo)
            if move:
                h.refit_accel(rend)
            h.update(-1, rots)
        for layer in range(ny):
            ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
            assert np.array_equal(ro["prim"], rg["prim"]) and np.array_equal(ro["lit_mask"], rg["lit_mask"]), f"sun direction {d}: lit masks differ"
        assert np.array_equal(o.download_atlas(), g.download_atlas())
    st = g.stats()
    o.close()
    g.close()
    return st


def _sun_only(scene, grid, make_o, make_p):
    """The faithful subset (no point lights, rchit:85-94): only the sun's map is in play."""
    nxz, ny, nr = grid
    fl = L.FLAG_RAY_RECORDS | L.FLAG_COUNTERS | L.FLAG_SKIP_DEAD_SHADOW_RAYS
    o, g = make_o(make_params(nxz, ny, nr, flags=fl)), make_p(make_params(nxz, ny, nr, flags=fl))
    rots = parity.setup_pair(scene, (o, g), seed=12)
    for _ in range(2):
        for h in (o, g):
            h.update(-1, rots)
    for layer in range(ny):
        ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
        assert np.array_equal(ro["prim"], rg["prim"]) and np.array_equal(ro["lit_mask"], rg["lit_mask"])
    assert np.array_equal(o.download_atlas(), g.download_atlas())
    so, sg = o.stats(), g.stats()
    assert sg["shadow_rays"] + sg["shadow_rays_prefiltered"] == so["shadow_rays"] and sg["shadow_rays_prefiltered"] > 0
    o.close()
    g.close()


def test_sun_only_faithful_subset(oracle, make_emu):
    _sun_only(scenes.sponza_like(seed=8, grid=(6, 3, 64), curtain_res=8), (6, 3, 64), oracle.Oracle, make_emu)


@pytest.mark.gpu
def test_cuda_sun_only_faithful_subset(oracle, make_gpu):
    _sun_only(scenes.sponza_like(seed=8, grid=(6, 3, 64), curtain_res=8), (6, 3, 64), oracle.Oracle, make_gpu)


SUN_DIRS = [(0.3, -1.0, 0.2), (0.0, -1.0, 0.0), (-0.7, -0.4, 0.6), (1.0, -0.05, 0.0), (0.2, 1.0, 0.1), (0.0, 0.0, 0.0)]


def test_sun_map_follows_the_light_direction(oracle, make_emu):
    """Steep, vertical (the frame's up-vector switches), grazing, from-below and zero directions; sun-only flags so every answered
    ray is the sun's."""
    scene = scenes.sponza_like(seed=8, grid=(6, 3, 64), curtain_res=8)
    st = _sun_sequence(scene, (6, 3, 64), oracle.Oracle, make_emu, SUN_DIRS)
    assert st["shadow_rays_prefiltered"] > 0


def test_sun_map_with_moving_instances_and_a_moving_sun(oracle, make_emu):
    scene = scenes.sponza_like(seed=4, grid=(6, 3, 64), movers=6, curtain_res=6)
    _sun_sequence(scene, (6, 3, 64), oracle.Oracle, make_emu, [(0.3, -1.0, 0.2), (0.3, -1.0, 0.2), (0.35, -1.0, 0.2), (0.35, -1.0, 0.2), (-0.5, -0.6, 0.1)], move=True)


@pytest.mark.gpu
def test_cuda_sun_map_follows_the_light_direction(oracle, make_gpu):
    scene = scenes.sponza_like(seed=8, grid=(6, 3, 64), curtain_res=8)
    st = _sun_sequence(scene, (6, 3, 64), oracle.Oracle, make_gpu, SUN_DIRS)
    assert st["shadow_rays_prefiltered"] > 0


@pytest.mark.gpu
def test_cuda_sun_map_with_moving_instances_and_a_moving_sun(oracle, make_gpu):
    scene = scenes.sponza_like(seed=4, grid=(6, 3, 64), movers=6, curtain_res=6)
    _sun_sequence(scene, (6, 3, 64), oracle.Oracle, make_gpu, [(0.3, -1.0, 0.2), (0.3, -1.0, 0.2), (0.35, -1.0, 0.2), (0.35, -1.0, 0.2), (-0.5, -0.6, 0.1)], move=True)
