"""Pipelined updates (anerend_cuda_ddgi_set_pipeline): the closest-hit trace of an update on its own stream, under the tail of the
previous update — results must not change: sequences of updates with translation, refits and rebuilds in between, against the oracle
and against the same handle parameters without the pipeline, bit for bit."""
import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

pytestmark = pytest.mark.gpu
FLAGS = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS


@pytest.mark.parametrize("two_level", [False, True], ids=["one_level", "two_level"])
def test_pipelined_updates_match_the_oracle(two_level, oracle, make_gpu):
    scene = scenes.sponza_like(seed=4, grid=(8, 4, 128), movers=6, curtain_res=12)
    o = oracle.Oracle(make_params(8, 4, 128, flags=FLAGS))
    g = make_gpu(make_params(8, 4, 128, flags=FLAGS | (L.FLAG_TWO_LEVEL if two_level else 0)))
    parity.setup_pair(scene, (o, g), seed=6)
    g.set_timers(0)
    g.set_pipeline(True)
    seq = scenes.rotation_sequence(11, 4 * 12)
    rots = o.rotations(seq).reshape(12, 4, 16)
    for k in range(12):
        ubo = scene.with_camera((0.7 * k - 3.0, 5.0, 0.3 * k))
        for h in (o, g):
            h.set_scene_ubo(ubo)
            if k % 3 == 1:
                h.refit_accel(scene.animated_renderables(0.1 * k))       # the next trace has to wait for this
            if k == 7:
                h.build_accel()
            h.translate()
            h.update(-1, rots[k])
        if k in (0, 1, 5, 8, 11):
            for layer in range(4):
                ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
                assert np.array_equal(ro["prim"], rg["prim"]) and np.array_equal(ro["t"], rg["t"]) and np.array_equal(ro["lit_mask"], rg["lit_mask"]), (k, layer)
            assert np.array_equal(o.download_atlas(), g.download_atlas()), k
    o.close()
    g.close()


def test_pipelined_per_layer_calls_keep_the_sequential_semantics(make_gpu):
    """frame() cadence (one layer per call, in place): with and without the pipeline the same atlas after every call."""
    scene = scenes.cornell_box(grid=(8, 4, 64), wall_res=10, box_res=3)
    a, b = make_gpu(make_params(8, 4, 64, flags=L.FLAG_POINT_LIGHTS)), make_gpu(make_params(8, 4, 64, flags=L.FLAG_POINT_LIGHTS))
    parity.setup_pair(scene, (a, b), seed=2)
    b.set_timers(0)
    b.set_pipeline(True)
    rots = a.rotations(scenes.rotation_sequence(5, 16))
    for k in range(16):
        for h in (a, b):
            h.frame(rots[k])
        assert np.array_equal(a.download_atlas(), b.download_atlas()), k
    a.close()
    b.close()


@pytest.mark.timeout(600)
def test_config3_pipelined_updates_bit_exact(oracle, make_gpu):
    """C3 at full size, four consecutive updates enqueued without a host synchronisation in between (the next update's closest-hit
    trace runs under the tail of the previous one): atlas equal to the oracle's after every update, ray records after the last."""
    scene = scenes.sponza_like()
    flags = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS
    o, g = oracle.Oracle(make_params(24, 12, 256, flags=flags)), make_gpu(make_params(24, 12, 256, flags=flags))
    parity.setup_pair(scene, (o, g), seed=3)
    g.set_timers(0)
    g.set_pipeline(True)
    rots = o.rotations(scenes.rotation_sequence(21, 12 * 4)).reshape(4, 12, 16)
    for k in range(4):
        g.update(-1, rots[k])                      # all four enqueued back to back
    atlases = []
    for k in range(4):
        o.update(-1, rots[k])
        atlases.append(o.download_atlas().copy())
    assert np.array_equal(atlases[3], g.download_atlas())
    for layer in (0, 5, 11):
        ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
        assert np.array_equal(ro["prim"], rg["prim"]) and np.array_equal(ro["t"], rg["t"]) and np.array_equal(ro["lit_mask"], rg["lit_mask"])
        assert np.array_equal(ro["radiance"].view(np.uint32), rg["radiance"].view(np.uint32))
    o.close()
    g.close()


def test_timer_levels_on_the_device(make_gpu):
    """Level 2: a time for every kernel of the RT pass; level 1 (default): the passes only; level 0: nothing — and the same atlas."""
    scene = scenes.cornell_box(grid=(8, 4, 64), wall_res=10, box_res=3)
    a, b = make_gpu(make_params(8, 4, 64, flags=L.FLAG_POINT_LIGHTS)), make_gpu(make_params(8, 4, 64, flags=L.FLAG_POINT_LIGHTS))
    rots = parity.setup_pair(scene, (a, b), seed=2)
    a.set_timers(2)
    a.update(-1, rots)
    t = a.timings()
    assert t["rt_ms"] > 0.0 and all(v > 0.0 for v in t["rt_kernel_ms"]), t
    b.update(-1, rots)
    t = b.timings()
    assert t["rt_ms"] > 0.0 and all(v == 0.0 for v in t["rt_kernel_ms"]), t
    b.set_timers(0)
    b.update(-1, rots)
    a.update(-1, rots)
    t = b.timings()
    assert t["rt_ms"] == 0.0 and t["kernel_launches"] > 0, t
    assert np.array_equal(a.download_atlas(), b.download_atlas())
    a.close()
    b.close()
