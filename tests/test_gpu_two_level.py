"""Two-level acceleration structure (ANEREND_DDGI_FLAG_TWO_LEVEL) on the GPU: the CUDA kernels behind the C ABI —
wavefront (k_trace<.., TWO>, k_shade_prep<.., TWO>) and the one-thread-per-ray kernel — against the oracle, which flattens every
instance and knows nothing of the structure.  Everything bit-exact: flattened hit ids, (t, u, v), lit masks, radiance, atlas.
Each size test appends what it observed to gpurun_out/r02_parity.txt."""
import os
import time

import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity
from tests.test_two_level import SCENES

pytestmark = pytest.mark.gpu
ALL = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_TEXEL_RECORDS | L.FLAG_COUNTERS
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = {"wavefront": 0, "wavefront_skip_dead": L.FLAG_SKIP_DEAD_SHADOW_RAYS, "simple": L.FLAG_SIMPLE_KERNELS}


def _report(line):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r02_parity.txt"), "a") as f:
        f.write(line + "\n")
    print(line)


def _bit_exact(o, g, ny, radiance_bits=True):
    """radiance_bits=False: where `pow` enters the radiance (toLinear of non-white vertex colours, rchit:75; the one-thread-per-ray
    kernel's own division / pow sequences) the device's result may sit an ulp off libm's — DESIGN.md §2, <= 2e-6 relative; the
    stored RGBA16F values and the atlas are identical all the same."""
    mism = 0
    for layer in range(ny):
        ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
        mism += int((ro["prim"] != rg["prim"]).sum())
        for k in ("t", "u", "v", "lit_mask"):
            assert np.array_equal(ro[k], rg[k]), k
        if radiance_bits:
            assert np.array_equal(ro["radiance"].view(np.uint32), rg["radiance"].view(np.uint32))
        else:
            assert parity.rel_err(rg["radiance"], ro["radiance"], floor=1e-3).max() <= 1e-5
    assert mism == 0
    assert np.array_equal(o.download_atlas(), g.download_atlas())


@pytest.mark.parametrize("kernels", sorted(KERNELS))
@pytest.mark.parametrize("name", sorted(SCENES))
def test_two_level_update_matches_oracle(name, kernels, oracle, make_gpu):
    scene, (nxz, ny, nr) = SCENES[name]()
    o = oracle.Oracle(make_params(nxz, ny, nr, flags=ALL | KERNELS[kernels]))
    g = make_gpu(make_params(nxz, ny, nr, flags=ALL | KERNELS[kernels] | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, g), seed=3)
    assert g.bvh_info().num_triangles == scene.num_triangles
    for _ in range(3):
        for h in (o, g):
            h.translate()
            h.update(-1, rots)
        _bit_exact(o, g, ny, radiance_bits=kernels != "simple" and not name.startswith("soup"))
    so, sg = o.stats(), g.stats()
    assert sg["primary_hits"] == so["primary_hits"] and sg["shadow_rays"] == so["shadow_rays"]      # no light maps: every shadow ray is traced
    o.close()
    g.close()


def test_two_level_equals_one_level_on_the_gpu(make_gpu):
    """The same handle parameters with and without the flag: identical ray records and atlas, far fewer stored triangles."""
    scene = scenes.instanced_field(5, instances=600, grid=(8, 4, 128))
    flags = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS
    a, b = make_gpu(make_params(8, 4, 128, flags=flags)), make_gpu(make_params(8, 4, 128, flags=flags | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (a, b), seed=9)
    for _ in range(2):
        for h in (a, b):
            h.update(-1, rots)
        _bit_exact(a, b, 4)
    info = b.two_level_info()
    assert info.num_meshes == 5 and info.stored_triangles * 50 < scene.num_triangles
    a.close()
    b.close()


@pytest.mark.timeout(600)
def test_two_level_instanced_scene_with_movers(oracle, make_gpu):
    """1024 instances of four props (tiny ones included), 64 of them moving: refit = rebuild of the instance level only; the
    oracle re-flattens.  Hidden and culled instances drop out through refit_accel_device's check of the GPU-written bits."""
    scene = scenes.instanced_field(3, instances=1024, grid=(12, 4, 128))
    flags = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS
    o, g = oracle.Oracle(make_params(12, 4, 128, flags=flags)), make_gpu(make_params(12, 4, 128, flags=flags | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, g), seed=4)
    info0 = g.two_level_info()
    t0 = time.time()
    for k in range(4):
        rend = scene.animated_renderables(0.25 * (k + 1))
        if k == 3:
            rend["visible"][5:40:3] = 0                       # hidden
            rend["visible"][41:90:5] |= 2                     # cull bit (cull.comp:296-303)
        for h in (o, g):
            if k == 3:
                scene.renderables = rend
                h.upload_scene(scene)
                h.build_accel()
            else:
                h.refit_accel(rend)
            h.update(-1, rots)
        _bit_exact(o, g, 4)
    info = g.two_level_info()
    assert info.num_meshes == info0.num_meshes and info.stored_triangles == info0.stored_triangles
    _report(f"two-level, instanced_field (1024 instances, {scene.num_triangles} flattened / {info.stored_triangles} stored triangles, {info.num_meshes} meshes, "
            f"instance level {info.tlas_nodes} nodes, depth {info.tlas_depth}+{info.blas_depth_max}): 4 updates with moving / hidden / culled instances, "
            f"hit-id / lit-mask / radiance / atlas mismatches 0; instance-level rebuild {info.tlas_build_ms:.3f} ms  ({time.time() - t0:.0f} s)")
    o.close()
    g.close()


@pytest.mark.timeout(900)
def test_two_level_config3_scene_bit_exact(oracle, make_gpu):
    """The C3 scene (265k flattened triangles: 161 instances of 10 meshes, 56k triangles stored) through the two-level structure: four layers of the
    full-size grid ray by ray, and the atlas of those layers."""
    scene = scenes.sponza_like()
    flags = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS
    o, g = oracle.Oracle(make_params(24, 12, 256, flags=flags)), make_gpu(make_params(24, 12, 256, flags=flags | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, g), seed=3)
    layers = [0, 3, 6, 11]
    t0 = time.time()
    rays = mism = lit = rad = 0
    for layer in layers:                 # the oracle's per-layer call (sequential semantics, the same on both sides)
        for h in (o, g):
            h.update(layer, rots[layer])
    for layer in layers:
        ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
        rays += ro.size
        mism += int((ro["prim"] != rg["prim"]).sum())
        lit += int((ro["lit_mask"] != rg["lit_mask"]).sum())
        rad += int((ro["radiance"].view(np.uint32) != rg["radiance"].view(np.uint32)).sum())
    texels = int((o.download_atlas() != g.download_atlas()).sum())
    info = g.two_level_info()
    _report(f"two-level, C3 scene ({scene.num_triangles} flattened / {info.stored_triangles} stored triangles, {info.num_meshes} meshes): {rays} rays in layers {layers}: "
            f"hit-id mismatches {mism}, lit-mask mismatches {lit}, radiance words differing {rad}, atlas values differing {texels}  ({time.time() - t0:.0f} s)")
    assert mism == 0 and lit == 0 and rad == 0 and texels == 0
    o.close()
    g.close()


def test_two_level_shadow_rt_and_specular_gi(oracle, make_gpu):
    from tests import shadow_inputs as SI
    scene = scenes.random_soup(21, ntris=800)
    o = oracle.Oracle(make_params(4, 3, 64, flags=L.FLAG_POINT_LIGHTS))
    g = make_gpu(make_params(4, 3, 64, flags=L.FLAG_POINT_LIGHTS | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, g), seed=5)
    for h in (o, g):
        h.update(-1, rots)
    W, H = 64, 40
    depth, geom0 = SI.depth_buffer(W, H, seed=3), SI.gbuffer_normals(W, H, seed=4)
    ivp = SI.perspective_inv_view_proj(2)
    for h in (o, g):
        h.set_scene_ubo(SI.with_shadow_camera(scene.ubo, ivp, W, H))
    assert np.array_equal(o.shadow_rt(depth), g.shadow_rt(depth))
    for h in (o, g):
        h.set_scene_ubo(SI.with_specular_camera(scene.ubo, ivp, W, H))
    assert np.array_equal(o.specular_gi(depth, geom0), g.specular_gi(depth, geom0))
    o.close()
    g.close()


@pytest.mark.parametrize("kernels", ["wavefront", "simple"])
def test_two_level_deforming_mesh(kernels, oracle, make_gpu):
    """update_vertices + refit: only the hierarchies of the meshes written to are rebuilt (UpdateBlasPass.cpp:60-175)."""
    scene, (nxz, ny, nr) = SCENES["textured"]()
    o = oracle.Oracle(make_params(nxz, ny, nr, flags=ALL | KERNELS[kernels]))
    g = make_gpu(make_params(nxz, ny, nr, flags=ALL | KERNELS[kernels] | L.FLAG_TWO_LEVEL))
    rots = parity.setup_pair(scene, (o, g), seed=3)
    rng = np.random.RandomState(4)
    nv = len(scene.vertices)
    first, count = nv // 4, nv // 2
    for it in range(3):
        v = scene.vertices[first:first + count].copy()
        v["pos"] += rng.uniform(-0.03, 0.03, v["pos"].shape).astype(np.float32) * (it + 1)
        v["normal"] += rng.uniform(-0.2, 0.2, v["normal"].shape).astype(np.float32)
        v["uv"] += np.float32(0.1 * (it + 1))
        for h in (o, g):
            h.update_vertices(v, first)
            h.refit_accel()
            h.update(-1, rots)
        _bit_exact(o, g, ny, radiance_bits=False)
    o.close()
    g.close()


def test_two_level_row_stripes_assemble_the_one_level_atlas(make_gpu):
    """The multi-GPU sharding (row stripes) on two-level handles: the merged rows equal the atlas of one one-level
    handle bit for bit, update after update."""
    nxz, ny, nr, world = 8, 4, 64, 4
    scene = scenes.instanced_field(11, instances=200, grid=(nxz, ny, nr), extent=10.0)
    flags = L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS
    full = make_gpu(make_params(nxz, ny, nr, flags=flags))
    parts = [make_gpu(make_params(nxz, ny, nr, flags=flags | L.FLAG_TWO_LEVEL)) for _ in range(world)]
    rots = parity.setup_pair(scene, [full] + parts, seed=8)
    for r, h in enumerate(parts):
        h.set_row_stripe(r, world)
    for it in range(3):
        full.update(-1, rots)
        for h in parts:
            h.update(-1, rots)
        slabs = [h.download_atlas() for h in parts]
        merged = slabs[0].copy()
        for y in range(ny):
            for z in range(nxz):
                r0 = 10 * nxz * y + 10 * z
                merged[r0:r0 + 10] = slabs[z % world][r0:r0 + 10]
        for h in parts:
            h.upload_atlas(merged)
        assert np.array_equal(merged, full.download_atlas())
    for h in [full] + parts:
        h.close()


def test_two_level_random_scenes_on_the_wavefront_kernels(oracle, make_gpu):
    """Seeded fuzz of k_trace<.., TWO> / k_shade_prep<.., TWO>: random soups, Cornell boxes and instanced fields (tiny, rotated,
    non-uniformly scaled instances), random grids / ray counts / cameras / light counts; hit ids, (t, u, v), lit masks and the atlas
    bit for bit, radiance to the pow ulps of DESIGN.md §2."""
    for seed in range(8000, 8024):
        rng = np.random.RandomState(seed)
        nxz, ny, nr = int(rng.choice([2, 3, 5, 8])), int(rng.choice([1, 2, 3])), int(rng.choice([32, 64, 128, 256]))
        cam = tuple(rng.uniform(-3, 3, 3) * [1, 0.5, 1] + [0, 1.5, 0])
        kind = rng.randint(0, 3)
        if kind == 0:
            s = scenes.random_soup(seed, ntris=int(rng.choice([300, 1500, 6000])), extent=float(rng.uniform(2, 12)), grid=(nxz, ny, nr), camera=cam,
                                   nlights=int(rng.randint(0, 4)))
        elif kind == 1:
            s = scenes.instanced_field(seed, instances=int(rng.choice([30, 200, 900])), grid=(nxz, ny, nr), extent=float(rng.uniform(3, 20)),
                                       seg=(int(rng.choice([6, 12])), 4), nlights=int(rng.randint(0, 4)))
            s.ubo = s.with_camera(cam)
        else:
            s = scenes.cornell_box(seed=seed, grid=(nxz, ny, nr), wall_res=int(rng.choice([2, 6, 12])), box_res=2)
            s.ubo = s.with_camera(cam)
        extra = L.FLAG_SKIP_DEAD_SHADOW_RAYS if rng.rand() < 0.5 else 0
        flags = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | extra
        o, g = oracle.Oracle(make_params(nxz, ny, nr, flags=flags)), make_gpu(make_params(nxz, ny, nr, flags=flags | L.FLAG_TWO_LEVEL))
        rots = parity.setup_pair(s, (o, g), seed=seed)
        for it in range(2):
            for h in (o, g):
                h.translate()
                h.update(-1, rots)
            _bit_exact(o, g, ny, radiance_bits=False)
        o.close()
        g.close()
