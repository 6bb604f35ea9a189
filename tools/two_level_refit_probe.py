#!/usr/bin/env python
"""A few instance-level rebuilds of a two-level structure (for `ncu --metrics gpu__time_duration.sum` launch lists).
    python tools/two_level_refit_probe.py [instances]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import anerend_b200 as A  # noqa: E402
from anerend_b200 import layouts as L, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scene = scenes.instanced_field(17, instances=n, grid=(8, 4, 64), extent=28.0)
h = A.Ddgi(A.make_params(8, 4, 64, flags=L.FLAG_POINT_LIGHTS | L.FLAG_TWO_LEVEL))
h.upload_scene(scene)
h.set_scene_ubo(scene.ubo)
h.build_accel()
for k in range(4):
    h.refit_accel(scene.animated_renderables(k / 30.0))
    h.sync()
    print("refit", k, h.timings()["accel_ms"], "ms", flush=True)
h.close()
