#!/usr/bin/env python
"""A few full-grid C3 updates on the two-level structure (target of `ncu -k regex:k_trace`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import anerend_b200 as A  # noqa: E402
from anerend_b200 import layouts as L, scenes  # noqa: E402
from tools.two_level_bench import workload  # noqa: E402

scene, (nxz, ny, nr), seed = workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
h = A.Ddgi(A.make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS | L.FLAG_TWO_LEVEL))
h.upload_scene(scene)
h.set_scene_ubo(scene.ubo)
h.build_accel()
rots = h.rotations(scenes.rotation_sequence(seed, ny))
for _ in range(6):
    h.update(-1, rots)
h.sync()
h.close()
