#!/usr/bin/env python
"""One-level (flattened) against two-level (ANEREND_DDGI_FLAG_TWO_LEVEL) acceleration structure on the current GPU: memory, build,
the per-frame cost of moving instances (refit of the flattened hierarchy + dynamic light maps against a rebuild of the instance
level — what UpdateTLASPass.cpp:198-228 does every frame) and the full-grid probe update, on scenes with many instances of few
meshes and on the C3 / C5 scenes.  One JSON line per (scene, structure).
    python tools/two_level_bench.py [instanced4096 instanced16384 c3 c5]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import anerend_b200 as A  # noqa: E402
from anerend_b200 import layouts as L, scenes  # noqa: E402


def workload(name):
    if name.startswith("instanced"):
        n = int(name[len("instanced"):])
        grid = (24, 12, 256)
        return scenes.instanced_field(17, instances=n, grid=grid, extent=28.0 if n <= 4096 else 48.0), grid, 5
    return bench.build_workload(name)


def mem_used():
    free, total = torch.cuda.mem_get_info()
    return total - free


def run(name, label, flags, steps=10):
    scene, (nxz, ny, nr), seed = workload(name)
    torch.cuda.synchronize()
    m0 = mem_used()
    h = A.Ddgi(A.make_params(nxz, ny, nr, flags=flags))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.set_timers(2)                 # events around every kernel: the update times below include their cost
    m1 = mem_used()
    build = []
    for _ in range(3):
        t0 = time.perf_counter()
        h.build_accel()
        h.sync()
        build.append((time.perf_counter() - t0) * 1e3)
    m2 = mem_used()
    rots = h.rotations(scenes.rotation_sequence(seed, ny))
    upd, move = [], []
    for k in range(steps + 3):
        if scene.movers:
            rend = scene.animated_renderables(k / 30.0)
            h.refit_accel(rend)
            h.sync()
            ms = h.timings()["accel_ms"]
            if k >= 3:
                move.append(ms)
        h.update(-1, rots)
        h.sync()
        t = h.timings()
        if k >= 3:
            upd.append(t["rt_ms"] + t["convolve_ms"])
    info, two = h.bvh_info(), h.two_level_info()
    out = {"scene": name, "structure": label, "flattened_triangles": int(info.num_triangles), "instances": int(info.num_instances),
           "stored_triangles": int(two.stored_triangles) if two.num_meshes else int(info.num_triangles), "meshes": int(two.num_meshes),
           "accel_device_mb": (m2 - m1) / 2 ** 20, "build_wall_ms_first": build[0], "build_wall_ms_warm": min(build[1:]),
           "moving_instances": len(scene.movers), "per_frame_accel_update_ms": float(np.mean(move)) if move else None,
           "full_grid_update_ms": float(np.mean(upd)), "kernels_ms": [round(float(v), 4) for v in h.timings()["rt_kernel_ms"]]}
    h.close()
    del h
    torch.cuda.synchronize()
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    base = L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS
    for name in (sys.argv[1:] or ["instanced4096", "c3"]):
        run(name, "one-level, light maps", base)
        run(name, "one-level, no light maps", base | L.FLAG_NO_LIGHT_MAPS)
        run(name, "two-level", base | L.FLAG_TWO_LEVEL)
