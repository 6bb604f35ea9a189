"""BASELINE.json's configurations at their STATED size, CUDA (through the C ABI) against the oracle, everything bit-exact:
C3 all 12 layers, C4 the whole 120-update scroll on the C3 scene with 16 movers, C5 sampled layers of the 2 M-triangle
volume.  Each test appends the counts it observed to gpurun_out/r02_parity.txt (copied to profiles/ after a GPU run)."""
import os
import time

import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(line):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r02_parity.txt"), "a") as f:
        f.write(line + "\n")
    print(line)


def _compare_layer(o, g, layer):
    """prim / t / u / v / lit mask / radiance of every ray of one layer: returns (#rays, #hit-id mismatches, #lit-mask
    mismatches, #radiance values that differ in any bit)."""
    ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
    mism = int((ro["prim"] != rg["prim"]).sum())
    same = ro["prim"] == rg["prim"]
    assert np.array_equal(ro["t"][same], rg["t"][same]) and np.array_equal(ro["u"][same], rg["u"][same]) and np.array_equal(ro["v"][same], rg["v"][same])
    lit = int((ro["lit_mask"] != rg["lit_mask"]).sum())
    rad = int((ro["radiance"].view(np.uint32) != rg["radiance"].view(np.uint32)).sum())
    return ro.size, mism, lit, rad


@pytest.mark.timeout(900)
def test_config3_full_size_every_layer_bit_exact(oracle, make_gpu):
    """C3 (24x12x24 x 256 over ~262k triangles, 8 lights): every ray of all 12 layers — hit ids, t/u/v, lit masks and radiance —
    and the whole stored atlas equal to the oracle's bit for bit; no allowance."""
    scene = scenes.sponza_like()
    p = make_params(24, 12, 256, flags=L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS)
    o, g = oracle.Oracle(p), make_gpu(p)
    rots = parity.setup_pair(scene, (o, g), seed=3)
    t0 = time.time()
    for h in (o, g):
        h.update(-1, rots)
    tot = [0, 0, 0, 0]
    for layer in range(12):
        for k, v in enumerate(_compare_layer(o, g, layer)):
            tot[k] += v
    ao, ag = o.download_atlas(), g.download_atlas()
    texels = int((ao != ag).any(axis=-1).sum()) if ao.ndim == 3 else int((ao != ag).sum())
    # a second update reads the first one's atlas (multi-bounce): still identical
    for h in (o, g):
        h.update(-1, rots)
    texels2 = int((o.download_atlas() != g.download_atlas()).sum())
    _report(f"C3 full size: {tot[0]} rays in 12 layers: hit-id mismatches {tot[1]}, lit-mask mismatches {tot[2]}, radiance words differing {tot[3]}; "
            f"atlas texels differing after update 1: {texels}, after update 2: {texels2}  ({time.time() - t0:.0f} s)")
    assert tot[1] == 0 and tot[2] == 0 and tot[3] == 0
    assert texels == 0 and texels2 == 0
    o.close()
    g.close()


@pytest.mark.timeout(1800)
def test_config4_full_size_scroll_120_updates(oracle, make_gpu):
    """C4 at full size: the C3-sized scene with 16 moving rigid nodes, 120 updates at 30 Hz, camera x: -8 -> +8 (probe-grid
    translations), previous-frame atlas for multi-bounce, refit before every update; no resync.  The final atlas and the
    last update's hit ids are bit-identical to the oracle's."""
    scene = scenes.sponza_like(seed=4, movers=16)
    p = make_params(24, 12, 256, flags=L.FLAG_POINT_LIGHTS | L.FLAG_RAY_RECORDS | L.FLAG_SKIP_DEAD_SHADOW_RAYS)
    o, g = oracle.Oracle(p), make_gpu(p)
    parity.setup_pair(scene, (o, g), seed=4)
    rots = o.rotations(scenes.rotation_sequence(4, 12 * 120)).reshape(120, 12, 16)
    translations, last_ix = 0, 0
    t0 = time.time()
    for k in range(120):
        t = k / 30.0
        cam = np.array([-8.0 + 16.0 * k / 119.0, 5.0, 0.3 * np.sin(t)])
        ubo = scene.with_camera(cam)
        rend = scene.animated_renderables(t)
        for h in (o, g):
            h.set_scene_ubo(ubo)
            h.refit_accel(rend)
            h.translate()
            h.update(-1, rots[k])
        ix = int(cam[0])
        translations += int(ix != last_ix)
        last_ix = ix
        if k in (0, 59):
            assert np.array_equal(o.download_atlas(), g.download_atlas()), f"atlas differs after update {k + 1}"
    tot = [0, 0, 0, 0]
    for layer in range(12):
        for i, v in enumerate(_compare_layer(o, g, layer)):
            tot[i] += v
    texels = int((o.download_atlas() != g.download_atlas()).sum())
    _report(f"C4 full size: 120 updates, {translations} grid translations, 16 movers refitted every step; last update {tot[0]} rays: hit-id mismatches {tot[1]}, "
            f"lit-mask mismatches {tot[2]}, radiance words differing {tot[3]}; final atlas words differing: {texels}  ({time.time() - t0:.0f} s)")
    assert translations >= 14
    assert tot[1] == 0 and tot[2] == 0 and tot[3] == 0 and texels == 0
    o.close()
    g.close()


@pytest.mark.timeout(1200)
def test_config5_sampled_layers_bit_exact(oracle, make_gpu):
    """C5 (64x32x64 x 256 over 2 M triangles): four layers spread over the height of the volume, ray by ray against the oracle,
    and those layers' atlas blocks after the convolution."""
    scene = scenes.large_volume()
    p = make_params(64, 32, 256, flags=L.FLAG_POINT_LIGHTS | L.FLAG_RAY_RECORDS | L.FLAG_SKIP_DEAD_SHADOW_RAYS)
    o, g = oracle.Oracle(p), make_gpu(p)
    rots = parity.setup_pair(scene, (o, g), seed=5)
    layers = (1, 9, 18, 30)
    t0 = time.time()
    g.update(-1, rots)
    tot = [0, 0, 0, 0]
    for layer in layers:
        o.trace_layer(layer, rots[layer])
        for i, v in enumerate(_compare_layer(o, g, layer)):
            tot[i] += v
    for layer in layers:
        o.convolve_layer(layer)
    ao, ag = o.download_atlas(), g.download_atlas()
    W = 640
    texels = sum(int((ao[l * W:(l + 1) * W] != ag[l * W:(l + 1) * W]).sum()) for l in layers)
    _report(f"C5 sampled layers {layers}: {tot[0]} rays: hit-id mismatches {tot[1]}, lit-mask mismatches {tot[2]}, radiance words differing {tot[3]}; "
            f"atlas words differing in those layers: {texels}  ({time.time() - t0:.0f} s)")
    assert tot[1] == 0 and tot[2] == 0 and tot[3] == 0 and texels == 0
    o.close()
    g.close()
