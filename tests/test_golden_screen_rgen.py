"""ShadowRT and SpecularGIRT ray generation against the REFERENCE'S OWN COMPILED SHADERS.

tests/golden/screen_rgen_golden.npz holds, for a 12 x 8 screen (depth buffer with sky pixels, un-normalised G-buffer normals, a
projective invViewProj with w != 1), the rays assets/shadow_rgen.spv and assets/mirror_reflections_rgen.spv hand to traceRayEXT and
the pixels they store, the misses shaded by the compiled shadow_rmiss / mirror_reflections_rmiss — executed with the small SPIR-V
interpreter tests/golden/spirv_interp.py (generator: tests/golden/make_screen_rgen_golden.py, run in the build container where the
reference is mounted).  calcWorldPos, the perspective divide, reflect() and the ray ranges are held bit for bit; the sky goes
through pow() and is held to 2e-6 relative.  The CUDA kernels are held to the oracle by the -m gpu tests."""
import os

import numpy as np

from anerend_b200 import layouts as L, make_params, scenes

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "screen_rgen_golden.npz"))
H, W = G["depth"].shape


def _handle(make, ubo_key):
    ubo = G[ubo_key].view(L.SCENE_UBO).copy()
    s = scenes.empty_scene(grid=(4, 2, 64))
    h = make(make_params(4, 2, 64))
    h.upload_scene(s); h.set_scene_ubo(ubo); h.build_accel()
    return h


def test_oracle_screen_rays_equal_the_compiled_reference_shaders(oracle):
    for specular, key in ((0, "shadow"), (1, "specular")):
        o = _handle(oracle.Oracle, "ubo_" + key)
        rays = G[key + "_rays"]
        n = 0
        for y in range(H):
            for x in range(W):
                if G["depth"][y, x] == 1.0:
                    assert np.isnan(rays[y, x]).all()                          # sky: no ray (shadow.rgen:45-48, mirror_reflections.rgen:57-60)
                    continue
                org, d = o.screen_ray(specular, x, y, float(G["depth"][y, x]), G["geom0"][y, x, :3])
                assert org.tolist() == rays[y, x, 0:3].tolist(), (key, x, y)   # calcWorldPos + perspective divide: bit-exact
                assert d.tolist() == rays[y, x, 3:6].tolist(), (key, x, y)     # -lightDir (not normalised) / reflect(v, normal): bit-exact
                assert rays[y, x, 6] == np.float32(0.1) and rays[y, x, 7] == 50.0
                n += 1
        assert n == 81
        o.close()


def _check_pass_outputs(h):
    depth, sky_px = G["depth"], G["depth"] == 1.0
    # every ray misses in an empty scene: ShadowRT stores what shadow_rmiss returns (1), sky pixels 1 without a ray
    got = h.shadow_rt(depth)
    assert np.array_equal(got, np.ones_like(depth))
    occluded = (np.add.outer(np.arange(H), np.arange(W)) % 3 == 0) & ~sky_px
    assert np.all(G["shadow_image"][occluded] == 0.0) and np.all(G["shadow_image"][~occluded] == 1.0)   # the binary: payload 0 survives a hit


def test_oracle_and_product_pass_outputs_match_the_compiled_miss_shaders(oracle, make_emu):
    for make in (oracle.Oracle, make_emu):
        h = _handle(make, "ubo_shadow")
        _check_pass_outputs(h)
        h.close()
        h = _handle(make, "ubo_specular")
        out = h.specular_gi(G["depth"], G["geom0"]).view(np.float16)
        img, sky_px = G["specular_image"], G["depth"] == 1.0
        miss = ~sky_px & (np.add.outer(np.arange(H), np.arange(W)) % 3 != 0)                       # pixels the generator shaded with the miss binary
        assert np.all(out[sky_px] == 0)                                                            # imageStore(vec4(0)) for the sky
        want = img[miss][0]
        assert np.all(img[miss] == want) and want[3] == 1e6                                         # mirror_reflections.rmiss: dist = 1e6
        rays = ~sky_px                                                                             # here every ray misses
        np.testing.assert_allclose(out[rays][:, :3].astype(np.float32), np.broadcast_to(want[:3], (rays.sum(), 3)), rtol=2.0 ** -10, atol=0)
        assert np.all(np.isinf(out[rays][:, 3]))                                                   # 1e6 does not fit fp16: +inf in the rgba16f image
        h.close()
