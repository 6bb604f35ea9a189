#!/usr/bin/env python
"""Golden vectors for the cull bit from the reference's OWN compiled shader: executes the main of assets/cull_comp.spv (the renderable
culling pass in front of the chain) with tests/golden/spirv_interp.py, one invocation per renderable, and records the `visible` word it
writes back: bit 1 = "farther than 50 m from the camera" (cull.comp:296-303, withinRangeOfCamera, buildTranslation, setCullBit), which
tlas_update.comp then honours.  The frustum / occlusion part of the pass is switched off through its own push constants (a far plane
behind the camera: nothing is drawn), so only the code in front of it runs.  Renderables sit around the 50 m boundary — including
3-4-5 positions at exactly 50 m — with rotated and scaled transforms (only the translation may matter), non-zero bounding-sphere centres,
hidden ones (bit 0 clear: not touched) and stale cull bits that must be cleared.

    python tests/golden/make_cull_golden.py [/root/reference/assets/cull_comp.spv]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from spirv_interp import Invocation, Module, f32  # noqa: E402

CAMERA = (3.0, 2.0, -4.0)


def make_renderables(n=72, seed=31):
    from anerend_b200 import layouts as L, scenes
    rng = np.random.RandomState(seed)
    rend = np.zeros(n, L.RENDERABLE)
    cam = np.array(CAMERA)
    exact = [(30, 40, 0), (0, 30, 40), (40, 0, 30), (-30, 0, -40), (50, 0, 0), (0, 0, -50)]          # exactly 50 m: still within (<=)
    for i in range(n):
        if i < len(exact):
            centre = cam + np.array(exact[i], np.float64)
        else:
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            centre = cam + d * rng.choice([rng.uniform(1, 49.9), rng.uniform(49.99, 50.01), rng.uniform(50.1, 90)])
        bounds = rng.uniform(-2, 2, 3) if i % 3 else np.zeros(3)
        t = centre - bounds                                    # centre = bounds.xyz + translation
        M = scenes.translate(t) @ scenes.rotate(i % 3, rng.uniform(0, 360)) @ scenes.scale(rng.uniform(0.2, 5.0, 3))
        rend["transform"][i] = M.T.astype(np.float32).ravel()
        rend["bounds"][i] = (*bounds, rng.uniform(0.5, 3.0))
        rend["tint"][i] = 1.0
        rend["num_meshes"][i] = 1
        rend["terrain_offset"][i] = -1
        rend["visible"][i] = (1, 1, 3, 1, 0, 2, 1 | 8, 3)[i % 8]          # visible / stale cull bit / hidden / noise in the upper bits
    return rend


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/cull_comp.spv"
    m = Module(path)
    rend = make_renderables()
    renderables = [[[[f32(r["transform"][4 * c + k]) for k in range(4)] for c in range(4)], [f32(x) for x in r["bounds"]], [f32(x) for x in r["tint"]],
                    int(r["model_offset"]), int(r["num_meshes"]), int(r["skeleton_offset"]), int(r["visible"]), int(r["first_material_index"]),
                    int(r["dynamic_model_offset"]), int(r["terrain_offset"])] for r in rend]
    ubo_gid = next(g for g, n in m.names.items() if n == "ubo" and g in m.globals)
    ubo_t = m.types[m.globals[ubo_gid][0]][2]
    ubo = m.zero(ubo_t)
    ubo[m.struct_member(ubo_t, "cameraPos")] = [f32(c) for c in CAMERA] + [f32(1)]
    pc_gid = next(g for g, n in m.names.items() if n == "pushConstants" and g in m.globals)
    pc_t = m.types[m.globals[pc_gid][0]][2]
    pc = m.zero(pc_t)
    eye = [[f32(1 if r == c else 0) for r in range(4)] for c in range(4)]
    pc[m.struct_member(pc_t, "view")] = eye
    pc[m.struct_member(pc_t, "farDist")] = f32(-1e30)          # sphereCenterView.z + radius > -farDist is false for everything: no draw
    pc[m.struct_member(pc_t, "nearDist")] = f32(0.1)
    pc[m.struct_member(pc_t, "drawCount")] = len(rend)
    for gid in range(len(rend) + 1):                           # one past drawCount: the bounds check drops it
        inv = Invocation(m, {"gl_GlobalInvocationID": [gid // 32, gid % 32, 0], "gl_NumWorkGroups": [1, 1, 1], "ubo": ubo, "pushConstants": pc,
                             "renderableBuffer": [renderables]})
        inv.run()
    out = np.array([r[6] for r in renderables], np.uint32)
    vis = (rend["visible"] & 1).astype(bool)
    print("visible renderables:", int(vis.sum()), "culled by the binary:", int(((out >> 1) & 1)[vis].sum()), "hidden ones untouched:", bool(np.array_equal(out[~vis], rend["visible"][~vis])))
    dst = os.path.join(HERE, "cull_golden.npz")
    np.savez_compressed(dst, camera=np.array(CAMERA, np.float32), renderables=rend.view(np.uint8), visible_out=out)
    print("wrote", dst)


if __name__ == "__main__":
    main()
