"""The oracle's small shader functions against the REFERENCE'S OWN COMPILED CODE, function by function.

tests/golden/helpers_golden.npz holds inputs and results of toLinear and octEncode as compiled into
assets/irradiance_probe_update_rchit.spv and of octDecode and pixelToUv as compiled into assets/probe_conv_comp.spv, each called
directly with the small SPIR-V interpreter tests/golden/spirv_interp.py (generator: tests/golden/make_helpers_golden.py, run in the
build container where the reference is mounted)."""
import os

import numpy as np

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "helpers_golden.npz"))


def test_to_linear_equals_the_compiled_function(oracle):
    srgb, want = G["srgb"], G["linear"]
    assert np.array_equal(want[:, 3], np.full(len(srgb), 0.25, np.float32))                  # alpha passes through (bindless.glsl:246)
    got = np.array([oracle.to_linear([v, np.float32(v) * np.float32(0.5), np.float32(1.0) - np.float32(v)]) for v in srgb], np.float32)
    lin_branch = np.stack([srgb, srgb * np.float32(0.5), np.float32(1.0) - srgb], 1) < np.float32(0.04045)
    assert lin_branch.any() and (~lin_branch).any()
    assert np.array_equal(got[lin_branch], want[:, :3][lin_branch])                          # x / 12.92: exact, and the same side of the boundary
    np.testing.assert_allclose(got[~lin_branch], want[:, :3][~lin_branch], rtol=2e-6, atol=0)   # pow(., 2.4): numpy powf vs libm


def test_oct_encode_equals_the_compiled_function(oracle):
    got = np.array([oracle.oct_encode(d) for d in G["dirs"]], np.float32)
    assert np.array_equal(got.view(np.uint32), G["oct_encode"].view(np.uint32))
    assert np.abs(G["oct_encode"]).max() <= 1.0


def test_oct_decode_and_texel_centres_equal_the_compiled_functions(oracle):
    uv = G["texel_uv"].reshape(8, 8, 2)
    for j in range(8):
        for i in range(8):                                                                   # pixelToUv, helpers.glsl:1-4
            assert uv[j, i].tolist() == [np.float32((2.0 * i + 1.0) / 16.0), np.float32((2.0 * j + 1.0) / 16.0)]
    got = np.array([oracle.oct_decode(o) for o in G["oct_points"]], np.float32)
    assert np.array_equal(got.view(np.uint32), G["oct_decode"].view(np.uint32))
    np.testing.assert_allclose(np.linalg.norm(G["oct_decode"], axis=1), 1.0, atol=1e-6)
    # the round trip the convolution relies on: encode(decode(texel centre)) lands on the texel centre again
    back = np.array([oracle.oct_encode(d) for d in G["oct_decode"][:64]], np.float32)
    np.testing.assert_allclose(back, G["oct_points"][:64], atol=2e-7)
