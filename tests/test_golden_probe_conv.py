"""The oracle's convolution against the REFERENCE'S OWN COMPILED SHADER.

tests/golden/probe_conv_golden.npz holds inputs and outputs of assets/probe_conv_comp.spv (the binary the renderer loads),
executed invocation by invocation with the small SPIR-V interpreter tests/golden/spirv_interp.py by
tests/golden/make_probe_conv_golden.py (both committed; the fixture is regenerated in the build container, where the
reference is mounted).  Three probes of layer 3 of the reference's native 64 x 8 x 64 grid with 256 rays, random fp16
radiance / direction / previous-atlas values, live mode (hysteresis blend) and bake mode (running weighted mean).
This pins loop order, ray-tile and atlas addressing, the weight threshold, normalisation, blend and constants of the oracle's
probe_conv transliteration to the shipped binary; the CUDA kernels are then held to the oracle bit for bit (tests -m gpu)."""
import os

import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "probe_conv_golden.npz"))
NXZ, NY, N, S = 64, 8, 256, 16


def _run_oracle(oracle, bake):
    o = oracle.Oracle(make_params(NXZ, NY, N, flags=L.FLAG_TEXEL_RECORDS))
    o.set_threads(0)
    ubo = scenes.empty_scene(grid=(NXZ, NY, N)).ubo.copy()
    if bake:
        ubo["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
    o.set_scene_ubo(ubo)
    layer = int(G["layer"])
    irr = np.zeros((S * NXZ, S * NXZ, 4), np.uint16); dirs = np.zeros_like(irr)
    atlas = o.download_atlas()
    for k, (px, pz) in enumerate(G["probes"]):
        irr[pz * S:(pz + 1) * S, px * S:(px + 1) * S] = G["irr"][k]
        dirs[pz * S:(pz + 1) * S, px * S:(px + 1) * S] = G["dir"][k]
        y0, x0 = 10 * NXZ * layer + 10 * pz + 1, 10 * px + 1
        atlas[y0:y0 + 8, x0:x0 + 8] = G["old"][k]
    o.upload_atlas(atlas)
    o.set_rays(layer, irr, dirs)
    o.convolve_layer(layer)
    rec, out = o.download_texel_records(), o.download_atlas()
    o.close()
    got = np.stack([rec[10 * NXZ * layer + 10 * pz + 1:10 * NXZ * layer + 10 * pz + 9, 10 * px + 1:10 * px + 9] for (px, pz) in G["probes"]])
    half = np.stack([out[10 * NXZ * layer + 10 * pz + 1:10 * NXZ * layer + 10 * pz + 9, 10 * px + 1:10 * px + 9] for (px, pz) in G["probes"]])
    return got, half


@pytest.mark.parametrize("mode", ["live", "bake"])
def test_oracle_convolution_equals_the_compiled_reference_shader(mode, oracle):
    want = G["out_" + mode]
    got, half = _run_oracle(oracle, mode == "bake")
    assert got.shape == want.shape == (3, 8, 8, 4)
    # same fp32 operations in the same order: the values agree to the last bit, not just to a tolerance
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"max rel diff {np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6)):.3e}"
    # and the stored RGBA16F texels are the round-to-nearest-even halves of those values
    assert np.array_equal(half.view(np.float16), want.astype(np.float16))


# ---- the same binary with the direction image the ray-generation pass writes (one rotation, shared by all probes): this case can be
# ---- fed to the PRODUCT through _upload_ray_radiance + _convolve_layer, so its convolution is held to the compiled shader directly
def _run_shared(make, bake):
    h = make(make_params(NXZ, NY, N, flags=L.FLAG_TEXEL_RECORDS))
    ubo = scenes.empty_scene(grid=(NXZ, NY, N)).ubo.copy()
    if bake:
        ubo["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
    h.set_scene_ubo(ubo)
    layer = int(G["layer"])
    irr = np.zeros((S * NXZ, S * NXZ, 4), np.uint16)
    atlas = h.download_atlas()
    for k, (px, pz) in enumerate(G["probes"]):
        irr[pz * S:(pz + 1) * S, px * S:(px + 1) * S] = G["irr"][k]
        y0, x0 = 10 * NXZ * layer + 10 * pz + 1, 10 * px + 1
        atlas[y0:y0 + 8, x0:x0 + 8] = G["old"][k]
    h.upload_atlas(atlas)
    h.upload_ray_radiance(layer, irr, G["shared_rot"])
    _, dirs = h.download_rays(layer)
    for (px, pz) in G["probes"]:                                    # ProbeRaysDirTex is the golden's direction image, bit for bit
        assert np.array_equal(dirs[pz * S:(pz + 1) * S, px * S:(px + 1) * S, :3].reshape(N, 3), G["shared_dir"])
    h.convolve_layer(layer)
    rec = h.download_texel_records()
    h.close()
    return np.stack([rec[10 * NXZ * layer + 10 * pz + 1:10 * NXZ * layer + 10 * pz + 9, 10 * px + 1:10 * px + 9] for (px, pz) in G["probes"]])


def _check_shared(make, mode):
    want = G["out_" + mode + "_shared"]
    got = _run_shared(make, mode == "bake")
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"max rel diff {np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6)):.3e}"


@pytest.mark.parametrize("mode", ["live", "bake"])
@pytest.mark.parametrize("side", ["oracle", "product_harness"])
def test_convolution_with_rgen_directions_equals_the_compiled_reference_shader(side, mode, oracle, make_emu):
    _check_shared(oracle.Oracle if side == "oracle" else make_emu, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["live", "bake"])
@pytest.mark.parametrize("kernels", ["wavefront", "simple"])
def test_cuda_convolution_equals_the_compiled_reference_shader(kernels, mode, make_gpu):
    def make(p):
        if kernels == "simple":
            p.flags |= L.FLAG_SIMPLE_KERNELS
        return make_gpu(p)
    _check_shared(make, mode)
