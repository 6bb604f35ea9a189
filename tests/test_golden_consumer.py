"""The atlas consumer against the REFERENCE'S OWN COMPILED CODE.

tests/golden/consumer_golden.npz holds the results of the function calcIndirectDiffuseLight as compiled into
assets/deferred_tiled_comp.spv (the deferred lighting pass), called directly with the small SPIR-V interpreter
tests/golden/spirv_interp.py for 160 surface points in and around the reference's native 64 x 8 probe grid, the way
deferred_tiled.comp:264-296 calls it: the tile's baked atlas (a bindless texture with its REPEAT sampler) on tiles that have one,
the live atlas through the pass's CLAMP_TO_EDGE sampler elsewhere (generator: tests/golden/make_consumer_golden.py, run in the
build container where the reference is mounted).  _indirect_diffuse(clamp_to_edge = 1) of the oracle and of the product's kernel
logic (host harness) is held to it; the CUDA kernel is held to the oracle bit for bit by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "consumer_golden.npz"))


def _setup(make):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_consumer_golden as gen

    class Binding:                       # gen.setup() only needs .Oracle: the handle factory under test
        Oracle = staticmethod(make)
    scene, h, tiles, pts, _ = gen.setup(Binding)
    assert np.array_equal(tiles, G["tiles"]) and np.array_equal(pts, G["points"])
    return scene, h, pts


def _check(make):
    scene, h, pts = _setup(make)
    got = h.indirect_diffuse(pts, scene.ubo["camera_pos"][0][:3], clamp_to_edge=True)
    h.close()
    want, on_tile = G["out"], G["on_tile"].astype(bool)
    assert 40 < on_tile.sum() < 120 and (want[on_tile].max(axis=1) > 0).sum() > 30 and (want[~on_tile].max(axis=1) > 0).sum() > 30
    assert np.all(got[:, 3] == 0)
    rel = np.abs(got[:, :3] - want) / np.maximum(np.abs(want), 1e-6)
    assert rel.max() <= 2e-6, f"max relative difference {rel.max():.3e}"                   # pow(1 - cosTheta, 5): numpy powf vs the defined product chain
    assert (got[:, :3].view(np.uint32) == want.view(np.uint32)).mean() > 0.9


@pytest.mark.parametrize("side", ["oracle", "product_harness"])
def test_indirect_diffuse_equals_the_compiled_deferred_pass_function(side, oracle, make_emu):
    _check(oracle.Oracle if side == "oracle" else make_emu)


@pytest.mark.gpu
def test_cuda_indirect_diffuse_equals_the_compiled_deferred_pass_function(make_gpu):
    _check(make_gpu)
