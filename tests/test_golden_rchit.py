"""The oracle's hit shading against the REFERENCE'S OWN COMPILED HIT SHADER.

tests/golden/rchit_golden.npz holds, for 96 hits of a textured test scene on the reference's native 64 x 8 x 64 probe grid, the
payload that assets/irradiance_probe_update_rchit.spv (the binary the renderer loads) produces when it is executed with the small
SPIR-V interpreter tests/golden/spirv_interp.py (generator: tests/golden/make_rchit_golden.py, run in the build container where
the reference is mounted).  The compiled code covers the vertex fetch and interpolation, the normal transform, toLinear,
getSurfaceDataFromMat with its texture taps, calcLight, calcIndirectDiffuseLight / sampleProbe / weightProbe / sampleSingleProbe /
octEncode against the previous atlas, and the emissive term (and, for the same hits, assets/mirror_reflections_rchit.spv — the
"SpecularGIRT" hit shader — with its other normal transform, distance payload and UBO_DDGI_FLAG gate); the interpreter is handed what the driver would provide (the hit,
the shadow-ray result, bilinear sampling as DESIGN.md defines it).  Multi-bounce is on, so the previous atlas matters.
The CUDA kernels are held to the oracle bit for bit by the -m gpu tests."""
import os

import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rchit_golden.npz"))


def test_oracle_hit_shading_equals_the_compiled_reference_shader(oracle):
    nxz, ny, nr = (int(v) for v in G["grid"])
    scene = scenes.textured_soup(51, grid=(nxz, ny, nr))
    o = oracle.Oracle(make_params(nxz, ny, nr, flags=L.FLAG_RAY_RECORDS))      # the faithful subset: sun only, every sun shadow ray traced
    o.set_threads(0)
    o.upload_scene(scene); o.set_scene_ubo(scene.ubo); o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(3, ny))
    o.update(-1, rots)
    o.update(-1, rots)                          # these hits are shaded against the atlas of the first update
    recs = {}
    got = np.zeros_like(G["irradiance"])
    for k, (layer, probe, ray) in enumerate(G["rays"]):
        if layer not in recs:
            recs[int(layer)] = o.download_ray_records(int(layer))
        r = recs[int(layer)][probe, ray]
        assert (int(r["prim"]), float(r["u"]), float(r["v"])) == tuple(G["prim_uv"][k]) and (int(r["lit_mask"]) & 1) == G["lit"][k]
        got[k] = r["radiance"]
    # the same hits through mirror_reflections.rchit (SpecularGIRT), against the atlas as it is now
    mgot = np.zeros_like(G["mirror"])
    for k in range(len(G["rays"])):
        prim, u, v = G["prim_uv"][k]
        rad, dist, lit = o.shade_hit(int(prim), u, v, G["origins"][k], specular=True)
        mgot[k] = (*rad, dist)
        assert (lit & 1) == G["lit"][k]
    # and through the probe hit shader with baked tiles paged in (quirk Q5 / SURVEY 8f N4): hits on those tiles place the grid on the tile
    shade = lambda k: o.shade_hit(int(G["prim_uv"][k][0]), G["prim_uv"][k][1], G["prim_uv"][k][2], G["origins"][k])[0]
    plain = np.stack([shade(k) for k in range(len(G["rays"]))])
    # bake mode: the grid is placed on tile bakeTileInfo.xy of size bakeTileInfo.z (pbr_light.glsl:274-291)
    bake_ubo = scene.ubo.copy()
    bake_ubo["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
    bake_ubo["bake_tile_info"][0] = G["bake_tile"]
    o.set_scene_ubo(bake_ubo)
    bgot = np.stack([shade(k) for k in range(len(G["rays"]))])
    o.set_scene_ubo(scene.ubo)
    brel = np.abs(bgot - G["baked"]) / np.maximum(np.abs(G["baked"]), 1e-6)
    assert brel.max() <= 2e-6, f"bake mode: max relative difference {brel.max():.3e}"
    assert (np.abs(bgot - plain).max(axis=1) > 0).mean() > 0.9
    o.upload_tile_infos(G["tile_pattern"])
    tgot = np.stack([shade(k) for k in range(len(G["rays"]))])
    o.close()
    twant = G["tiled"]
    trel = np.abs(tgot - twant) / np.maximum(np.abs(twant), 1e-6)
    assert trel.max() <= 2e-6, f"tiled: max relative difference {trel.max():.3e}"
    assert (tgot.view(np.uint32) == twant.view(np.uint32)).mean() > 0.9
    moved = (np.abs(tgot - plain).max(axis=1) > 0).mean()
    assert 0.2 < moved < 0.8, moved                                             # the checkerboard: some hits are on baked tiles, some are not
    mwant = G["mirror"]
    assert np.array_equal(mgot[:, 3], mwant[:, 3])                              # distance(pos, gl_WorldRayOriginEXT): bit-exact
    mrel = np.abs(mgot[:, :3] - mwant[:, :3]) / np.maximum(np.abs(mwant[:, :3]), 1e-6)
    assert mrel.max() <= 2e-6, f"mirror shader: max relative difference {mrel.max():.3e}"
    assert (mgot.view(np.uint32) == mwant.view(np.uint32)).mean() > 0.9
    assert np.abs(mwant[:, :3] - G["irradiance"]).max() > 0.05                  # and it is a different shader: other normal transform, gate and atlas
    want = G["irradiance"]
    assert want.max() > 1.0 and (want > 0).all(axis=1).mean() > 0.9 and len(set(G["lit"])) == 2       # a non-trivial sample: lit and shadowed hits
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    # the only operations not reproduced bit for bit are pow() calls (numpy's powf in the interpreter, libm in the oracle)
    assert rel.max() <= 2e-6, f"max relative difference {rel.max():.3e}"
    assert (got.view(np.uint32) == want.view(np.uint32)).mean() > 0.9


def _probe_hits_against_the_binary(make):
    nxz, ny, nr = (int(v) for v in G["grid"])
    scene = scenes.textured_soup(51, grid=(nxz, ny, nr))
    e = make(make_params(nxz, ny, nr, flags=L.FLAG_RAY_RECORDS))
    e.upload_scene(scene); e.set_scene_ubo(scene.ubo); e.build_accel()
    rots = e.rotations(scenes.rotation_sequence(3, ny))
    e.update(-1, rots)
    e.update(-1, rots)
    recs, got = {}, np.zeros_like(G["irradiance"])
    for k, (layer, probe, ray) in enumerate(G["rays"]):
        if layer not in recs:
            recs[int(layer)] = e.download_ray_records(int(layer))
        r = recs[int(layer)][probe, ray]
        assert (int(r["prim"]), float(r["u"]), float(r["v"])) == tuple(G["prim_uv"][k]) and (int(r["lit_mask"]) & 1) == G["lit"][k]
        got[k] = r["radiance"]
    e.close()
    want = G["irradiance"]
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    assert rel.max() <= 2e-6, f"max relative difference {rel.max():.3e}"
    assert (got.view(np.uint32) == want.view(np.uint32)).mean() > 0.9


def test_product_kernel_logic_hit_shading_equals_the_compiled_reference_shader(make_emu):
    """The same probe hits through the PRODUCT's kernel bodies (k_probe_trace / shade_hit_simple of csrc/, run by the host harness):
    the hit ids, barycentrics and shadow results are the golden's, the radiance is the compiled shader's."""
    _probe_hits_against_the_binary(make_emu)


@pytest.mark.gpu
def test_cuda_hit_shading_equals_the_compiled_reference_shader(make_gpu):
    """And through the CUDA kernels behind the C ABI (the production wavefront path)."""
    _probe_hits_against_the_binary(make_gpu)
