"""BASELINE config 1 — probe_conv.comp alone: 8 x 8 x 8 probes x 64 rays, synthetic radiance U[0, 4)^3 (seed 1), Fibonacci directions
under the seed-1 rotation, one convolution of every layer from the cleared atlas.  The reference case is CPU-only (single-thread
transliteration); here the same case also runs through the product: `_upload_ray_radiance` + `_convolve_layer` on the host harness
and (-m gpu) on the CUDA kernels, bit-identical to the oracle."""
import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes

NXZ, NY, NR = 8, 8, 64


def _radiance(layer):
    rng = np.random.RandomState(1000 + layer)
    irr = np.ones((8 * NXZ, 8 * NXZ, 4), np.float16)          # S = 8: 64 rays per probe as an 8 x 8 tile
    irr[..., :3] = rng.uniform(0, 4, irr[..., :3].shape)
    return irr


def _run(make, oracle):
    h = make(make_params(NXZ, NY, NR, flags=L.FLAG_TEXEL_RECORDS))
    h.set_scene_ubo(scenes.empty_scene(grid=(NXZ, NY, NR)).ubo)
    rot = oracle.make_rotation(*scenes.rotation_sequence(1, 1)[0])
    for layer in range(NY):
        h.upload_ray_radiance(layer, _radiance(layer), rot)
        h.convolve_layer(layer)
    atlas, rec = h.download_atlas(), h.download_texel_records()
    _, dirs = h.download_rays(3)
    h.close()
    return atlas, rec, dirs


def test_oracle_upload_path_equals_the_ray_image_path(oracle):
    """upload_ray_radiance(rot) builds the same ProbeRaysDirTex as handing set_rays the rotated Fibonacci image (tests/test_oracle_kat.py)."""
    from tests.test_oracle_kat import _conv_setup
    o, dir_img, _ = _conv_setup(oracle, NXZ, NY, NR, seed=1)
    o.set_scene_ubo(scenes.empty_scene(grid=(NXZ, NY, NR)).ubo)
    for layer in range(NY):
        o.set_rays(layer, _radiance(layer).view(np.uint16), dir_img.view(np.uint16))
        o.convolve_layer(layer)
    want = o.download_atlas()
    o.close()
    got, rec, dirs = _run(oracle.Oracle, oracle)
    assert np.array_equal(got, want) and np.array_equal(dirs, dir_img.view(np.uint16))
    from tests.parity import interior_mask
    m = interior_mask(NXZ, NY)
    assert np.isfinite(rec[m]).all() and rec[m][:, :3].min() > 0.02 and rec[m][:, :3].max() < 0.16      # 0.04 * a weighted mean of U[0, 4)


def test_product_kernel_logic_config1_matches_oracle(oracle, make_emu):
    want, wrec, wdirs = _run(oracle.Oracle, oracle)
    got, rec, dirs = _run(make_emu, oracle)
    assert np.array_equal(got, want) and np.array_equal(rec.view(np.uint32), wrec.view(np.uint32)) and np.array_equal(dirs, wdirs)


@pytest.mark.gpu
@pytest.mark.parametrize("kernels", ["wavefront", "simple"])
def test_cuda_config1_matches_oracle(kernels, oracle, make_gpu):
    want, wrec, wdirs = _run(oracle.Oracle, oracle)

    def make(p):
        if kernels == "simple":
            p.flags |= L.FLAG_SIMPLE_KERNELS
        return make_gpu(p)
    got, rec, dirs = _run(make, oracle)
    assert np.array_equal(got, want) and np.array_equal(dirs, wdirs)
    assert np.array_equal(rec.view(np.uint32), wrec.view(np.uint32))


# ---- non-finite radiance (an fp16 overflow of a very bright hit): probe_conv.comp:54-57 SKIPS rays whose weight is <= 1e-6, so an
# Inf / NaN radiance on a ray facing away from a texel must leave that texel finite; where the ray does contribute, Inf / NaN propagate.
def _radiance_nonfinite(layer):
    irr = _radiance(layer)
    rng = np.random.RandomState(77 + layer)
    flat = irr.reshape(-1, 4)
    idx = rng.choice(len(flat), 24, replace=False)
    flat[idx[:12], rng.randint(0, 3, 12)] = np.float16(np.inf)
    flat[idx[12:], rng.randint(0, 3, 12)] = np.float16(np.nan)
    return irr


def _run_nonfinite(make, oracle):
    h = make(make_params(NXZ, NY, NR, flags=L.FLAG_TEXEL_RECORDS))
    h.set_scene_ubo(scenes.empty_scene(grid=(NXZ, NY, NR)).ubo)
    rot = oracle.make_rotation(*scenes.rotation_sequence(1, 1)[0])
    for layer in range(2):
        h.upload_ray_radiance(layer, _radiance_nonfinite(layer), rot)
        h.convolve_layer(layer)
    atlas = h.download_atlas()
    h.close()
    return atlas[:2 * 10 * NXZ]


def _same_texels(a, b):
    """Equal bit patterns, NaNs compared as NaN (their payload is not defined by the shader)."""
    fa, fb = a.view(np.float16), b.view(np.float16)
    return bool(np.all((a == b) | (np.isnan(fa) & np.isnan(fb))))


def test_nonfinite_radiance_is_skipped_on_zero_weight_rays(oracle, make_emu):
    want = _run_nonfinite(oracle.Oracle, oracle)
    f = want.view(np.float16)
    assert np.isnan(f).any() or np.isinf(f).any(), "the fixture must reach some texels"
    from tests.parity import interior_mask
    m = interior_mask(NXZ, 2)
    assert np.isfinite(f[m]).mean() > 0.3, "texels facing away from the overflowing rays stay finite"
    assert _same_texels(_run_nonfinite(make_emu, oracle), want)


@pytest.mark.gpu
@pytest.mark.parametrize("kernels", ["wavefront", "wavefront_fast", "simple"])
def test_cuda_nonfinite_radiance_matches_oracle(kernels, oracle, make_gpu):
    want = _run_nonfinite(oracle.Oracle, oracle)

    def make(p):
        p.flags |= {"simple": L.FLAG_SIMPLE_KERNELS, "wavefront_fast": L.FLAG_FAST_MATH}.get(kernels, 0)
        return make_gpu(p)
    got = _run_nonfinite(make, oracle)
    if kernels == "wavefront_fast":       # same texels finite / non-finite; values within the fast-math bar
        fw, fg = want.view(np.float16).astype(np.float32), got.view(np.float16).astype(np.float32)
        assert np.array_equal(np.isfinite(fw), np.isfinite(fg))
        ok = np.isfinite(fw)
        assert np.max(np.abs(fw[ok] - fg[ok]) / np.maximum(np.abs(fw[ok]), 1e-3)) < 2e-3
    else:
        assert _same_texels(got, want)
