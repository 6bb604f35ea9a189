"""IrradianceProbeTrans against an independent restatement of the reference's copy regions.

IrradianceProbeTranslationPass.cpp:56-186 is host code that records two vkCmdCopyImage calls: the atlas to TempProbeTex, then one
VkImageCopy per probe layer back, whose source / destination offsets and extent depend on the sign of the camera's probe-cell step
(dx, dz).  It has no binary that could be executed here, so this test restates the REGIONS (cpp:133-172) with numpy slices — a copy
of `extent` texels from `srcOffset` of the old atlas to `dstOffset` of the atlas, everything outside the region left as it was — and
holds the oracle and the product's kernel logic (host harness) to it for all eight directions, a no-op step, and a jump of several
cells (which still shifts by one probe: quirk Q7)."""
import itertools

import numpy as np
import pytest

from anerend_b200 import make_params, scenes

NXZ, NY = 5, 3
W, H = 10 * NXZ, 10 * NXZ * NY


def vk_image_copies(old, dx, dz):
    """What the second vkCmdCopyImage of the pass leaves in the atlas (cpp:133-172)."""
    out = old.copy()                                   # texels outside every region keep their old value
    probe_pix, plane_h = 10, W                         # probePixSize = octPixelSize + 2; planeHeight = width
    for i in range(NY):
        src_x = dst_x = 0
        src_y = dst_y = i * probe_pix * NXZ            # planeYOffset
        ext_w, ext_h = W, plane_h
        if dx > 0:
            ext_w, src_x, dst_x = W - probe_pix, probe_pix, 0
        elif dx < 0:
            ext_w, src_x, dst_x = W - probe_pix, 0, probe_pix
        if dz > 0:
            ext_h = plane_h - probe_pix; src_y += probe_pix
        elif dz < 0:
            ext_h = plane_h - probe_pix; dst_y += probe_pix
        out[dst_y:dst_y + ext_h, dst_x:dst_x + ext_w] = old[src_y:src_y + ext_h, src_x:src_x + ext_w]
    return out


def _atlas():
    rng = np.random.RandomState(4)
    return rng.randint(0, 0x7BFF, (H, W, 4)).astype(np.uint16)       # finite, arbitrary RGBA16F bit patterns


def _run(h, scene, old, start, end):
    h.set_scene_ubo(scene.with_camera(start))
    h.translate()                                      # settles _lastProbeIdx on the start cell
    h.upload_atlas(old)
    h.set_scene_ubo(scene.with_camera(end))
    h.translate()
    return h.download_atlas()


@pytest.mark.parametrize("side", ["oracle", "product_harness"])
def test_translation_equals_the_reference_copy_regions(side, oracle, make_emu):
    make = oracle.Oracle if side == "oracle" else make_emu
    scene = scenes.empty_scene(grid=(NXZ, NY, 64))
    old = _atlas()
    h = make(make_params(NXZ, NY, 64))
    for dx, dz in itertools.product((-1, 0, 1), repeat=2):
        start = (7.5, 1.0, -3.5)
        got = _run(h, scene, old, start, (start[0] + dx, 1.0, start[2] + dz))
        assert np.array_equal(got, vk_image_copies(old, dx, dz)), (dx, dz)
    # across zero: int() truncates toward zero (quirk Q1), so -0.5 -> +0.5 is the same cell and +0.5 -> -1.5 is one step
    assert np.array_equal(_run(h, scene, old, (-0.5, 1.0, 0.5), (0.5, 1.0, -0.5)), old)
    assert np.array_equal(_run(h, scene, old, (0.5, 1.0, 0.5), (-1.5, 1.0, 0.5)), vk_image_copies(old, -1, 0))
    # a jump of several cells still moves the atlas by one probe (Q7)
    assert np.array_equal(_run(h, scene, old, (0.5, 1.0, 0.5), (4.5, 1.0, -6.5)), vk_image_copies(old, 4, -7))
    # bake mode: the pass returns before doing anything (cpp:58-59)
    import anerend_b200.layouts as L
    h.set_scene_ubo(scene.with_camera((0.5, 1.0, 0.5))); h.translate(); h.upload_atlas(old)
    bake = scene.with_camera((3.5, 1.0, 0.5)); bake["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
    h.set_scene_ubo(bake); h.translate()
    assert np.array_equal(h.download_atlas(), old)
    h.close()
