#!/usr/bin/env python
"""Time of anerend_cuda_ddgi_build_accel (flatten + Morton sort + PLOC + BVH8 collapse + light maps) on the current GPU: CUDA events
around the build of a warm handle (the first build of a process also pays module loading and allocations), per workload.
    python tools/build_time.py [c3 c5 ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import anerend_b200 as A  # noqa: E402
from anerend_b200 import layouts as L  # noqa: E402

for name in (sys.argv[1:] or ["c3", "c5"]):
    scene, (nxz, ny, nr), _ = bench.build_workload(name)
    for label, flags in (("with light maps", L.FLAG_POINT_LIGHTS), ("hierarchy only", L.FLAG_POINT_LIGHTS | L.FLAG_NO_LIGHT_MAPS)):
        h = A.Ddgi(A.make_params(nxz, ny, nr, flags=flags))
        h.upload_scene(scene)
        h.set_scene_ubo(scene.ubo)
        ms, wall = [], []
        for k in range(4):
            t0 = time.perf_counter()
            h.build_accel()
            h.sync()
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(h.timings()["accel_ms"])
        info = h.bvh_info()
        print(f"{name} {label}: {info.num_triangles} triangles, {info.num_nodes} nodes; build_accel device ms {['%.2f' % v for v in ms]}, wall ms {['%.2f' % v for v in wall]}")
        h.close()
