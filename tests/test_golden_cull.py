"""The cull bit in front of the chain (cull.comp:241-244,296-303) against the REFERENCE'S OWN COMPILED SHADER.

tests/golden/cull_golden.npz holds the `visible` words assets/cull_comp.spv writes back for 72 renderables around the 50 m boundary
(3-4-5 positions at exactly 50 m, rotated / scaled transforms, non-zero bounding-sphere centres, hidden ones, stale cull bits) —
its main executed with the small SPIR-V interpreter tests/golden/spirv_interp.py (generator: tests/golden/make_cull_golden.py, run in
the build container where the reference is mounted).  _apply_camera_cull of the oracle and of the product (host code of the C ABI, here
through the host harness) must write the same words; the instance selection then leaves the culled ones out."""
import os

import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cull_golden.npz"))


def _scene():
    rend = G["renderables"].view(L.RENDERABLE).copy()
    b = scenes.SceneBuilder("cull", 0)
    pos = np.array([(-1, 0, -1), (1, 0, -1), (1, 0, 1)], np.float32)
    b.add_mesh(pos, np.tile(np.array([0, 1, 0], np.float32), (3, 1)), np.array([0, 1, 2], np.uint32))
    b.add_material((0.5, 0.5, 0.5))
    b.add_renderable(np.eye(4), 0, 0)
    s = b.finish(tuple(float(c) for c in G["camera"]), (4, 2, 64))
    s.renderables = rend                                   # every renderable: mesh 0 (model index 0), material 0
    return s


@pytest.mark.parametrize("side", ["oracle", "product_harness"])
def test_cull_bits_equal_the_compiled_cull_shader(side, oracle, make_emu):
    s = _scene()
    h = (oracle.Oracle if side == "oracle" else make_emu)(make_params(4, 2, 64))
    h.upload_scene(s); h.set_scene_ubo(s.ubo); h.build_accel()
    want = G["visible_out"]
    before = s.renderables["visible"].copy()
    changed = h.apply_camera_cull(50.0)
    assert changed == int((want != before).sum()) and changed > 20
    with pytest.raises(Exception):
        h.trace_layer(0, np.eye(4, dtype=np.float32))      # the instance set changed: the old acceleration structure is refused
    h.build_accel()
    inst = h.download_instances()
    keep = np.nonzero((want & 1).astype(bool) & ~((want >> 1) & 1).astype(bool))[0]
    assert np.array_equal(inst[:, 0], keep)                # exactly the renderables the binary left visible and unculled, in order
    assert h.apply_camera_cull(50.0) == 0                  # idempotent
    # a camera far away culls everything that is visible; coming back restores the golden's words
    far = s.with_camera((1000.0, 0.0, 0.0))
    h.set_scene_ubo(far)
    assert h.apply_camera_cull(50.0) > 0
    h.build_accel()
    assert h.bvh_info().num_instances == 0
    h.set_scene_ubo(s.ubo)
    h.apply_camera_cull(50.0); h.build_accel()
    assert np.array_equal(h.download_instances()[:, 0], keep)
    h.close()
