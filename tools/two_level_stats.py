#!/usr/bin/env python
"""Nodes visited / triangles tested per ray, one-level against two-level structure (FLAG_COUNTERS), one full-grid update.
    python tools/two_level_stats.py [c3 instanced4096 ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import anerend_b200 as A  # noqa: E402
from anerend_b200 import layouts as L, scenes  # noqa: E402
from tools.two_level_bench import workload  # noqa: E402

for name in (sys.argv[1:] or ["c3"]):
    scene, (nxz, ny, nr), seed = workload(name)
    for label, fl in (("one-level", L.FLAG_NO_LIGHT_MAPS), ("two-level", L.FLAG_TWO_LEVEL)):
        h = A.Ddgi(A.make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS | L.FLAG_COUNTERS | fl))
        h.upload_scene(scene)
        h.set_scene_ubo(scene.ubo)
        h.build_accel()
        h.set_timers(2)
        h.update(-1, h.rotations(scenes.rotation_sequence(seed, ny)))
        s = h.stats()
        pr, sr = max(1, s["primary_rays"]), max(1, s["shadow_rays"])
        t = h.timings()
        print(f"{name} {label}: primary nodes/ray {s['primary_nodes_visited'] / pr:.2f} tris/ray {s['primary_tris_tested'] / pr:.2f}; "
              f"level changes/ray {s['primary_level_changes'] / pr:.2f} (shadow {s['shadow_level_changes'] / sr:.2f}); shadow rays {sr} nodes/ray {s['shadow_nodes_visited'] / sr:.2f} tris/ray {s['shadow_tris_tested'] / sr:.2f}; kernels ms (counted build) {[round(v, 3) for v in t['rt_kernel_ms']]}", flush=True)
        h.close()
