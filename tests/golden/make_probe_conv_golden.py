#!/usr/bin/env python
"""Golden vectors for the convolution pass from the reference's OWN compiled shader: executes assets/probe_conv_comp.spv
(the binary the renderer loads) with tests/golden/spirv_interp.py for a few probes of one layer, in live and in bake mode, on
random fp16 inputs, and writes inputs + outputs to tests/golden/probe_conv_golden.npz.  Run in the build container, where
/root/reference is mounted; tests/test_golden_probe_conv.py only reads the .npz.

    python tests/golden/make_probe_conv_golden.py [/root/reference/assets/probe_conv_comp.spv]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from spirv_interp import Invocation, Module, f32  # noqa: E402

NXZ, S, LAYER = 64, 16, 3                       # probe_helpers.glsl: 64 probes per side, 16 x 16 ray tile
PROBES = [(5, 7), (63, 63), (0, 0)]             # (probeX, probeZ)


class Image:
    def __init__(self, data):                   # dict (x, y) -> 4 floats; reads of untouched pixels are an error
        self.d, self.writes = data, {}

    def read(self, xy):
        return self.d[tuple(xy)]

    def write(self, xy, v):
        self.writes[tuple(xy)] = [float(c) for c in v]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/probe_conv_comp.spv"
    m = Module(path)
    rng = np.random.RandomState(7)
    irr, dirs, old = {}, {}, {}
    irr_in = np.zeros((len(PROBES), S, S, 4), np.float16); dir_in = np.zeros_like(irr_in); old_in = np.zeros((len(PROBES), 8, 8, 4), np.float16)
    for k, (px, pz) in enumerate(PROBES):
        rad = rng.uniform(0, 4, (S, S, 3)).astype(np.float16)
        d = rng.normal(size=(S, S, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
        irr_in[k, ..., :3], irr_in[k, ..., 3] = rad, 1.0
        dir_in[k, ..., :3], dir_in[k, ..., 3] = d.astype(np.float16), 1.0
        old_in[k, ..., :3], old_in[k, ..., 3] = rng.uniform(0, 2, (8, 8, 3)).astype(np.float16), rng.uniform(0.5, 40.0, (8, 8)).astype(np.float16)
        for y in range(S):
            for x in range(S):                                      # image pixel (probeX*16 + x, probeZ*16 + y); arrays are [y][x]
                irr[(px * S + x, pz * S + y)] = [float(c) for c in irr_in[k, y, x]]
                dirs[(px * S + x, pz * S + y)] = [float(c) for c in dir_in[k, y, x]]
        for j in range(8):
            for i in range(8):                                      # atlas texel (10 px + 1 + i, 10*64*layer + 10 pz + 1 + j)
                old[(10 * px + 1 + i, 10 * NXZ * LAYER + 10 * pz + 1 + j)] = [float(c) for c in old_in[k, j, i]]
    # a second direction image, the one the ray-generation shader writes: mat3(randomRot) * Fibonacci direction, the same 256
    # directions for every probe (rgen:87-88,108) — the case the product's _upload_ray_radiance + _convolve_layer can be fed with
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import binding as ob
    shared_rot = ob.make_rotation(1, 211.25)
    fib = ob.fibonacci_dirs(S * S)
    R = shared_rot.reshape(4, 4)[:3, :3].T
    rdirs = np.stack([[np.float32(np.float32(np.float32(R[r, 0] * d[0]) + np.float32(R[r, 1] * d[1])) + np.float32(R[r, 2] * d[2])) for r in range(3)] for d in fib]).astype(np.float16)
    dirs_shared = {}
    for (px, pz) in PROBES:
        for y in range(S):
            for x in range(S):
                dirs_shared[(px * S + x, pz * S + y)] = [float(c) for c in rdirs[y * S + x]] + [1.0]
    ubo_gid = next(g for g, n in m.names.items() if n == "ubo" and g in m.globals)
    ubo_struct = m.types[m.globals[ubo_gid][0]][2]
    flags_member = m.struct_member(ubo_struct, "flags")
    out = {}
    for mode, flags, dirs in (("live", 0, dirs), ("bake", 1 << 11, dirs), ("live_shared", 0, dirs_shared), ("bake_shared", 1 << 11, dirs_shared)):
        ubo = m.zero(ubo_struct); ubo[flags_member] = flags
        res = np.zeros((len(PROBES), 8, 8, 4), np.float32)
        for k, (px, pz) in enumerate(PROBES):
            for j in range(8):
                for i in range(8):
                    img_out = Image(dict(old))
                    inv = Invocation(m, {"irrIm": Image(irr), "dirIm": Image(dirs), "outputIm": img_out, "ubo": ubo, "pc": [LAYER],
                                         "gl_GlobalInvocationID": [8 * px + i, 8 * pz + j, 0]})
                    inv.run()
                    (xy, v), = img_out.writes.items()
                    assert xy == (10 * px + 1 + i, 10 * NXZ * LAYER + 10 * pz + 1 + j), xy       # the atlas addressing of the binary
                    res[k, j, i] = v
        out[mode] = res
        print(mode, "done", res[0, 0, 0], res[-1, 7, 7])
    dst = os.path.join(HERE, "probe_conv_golden.npz")
    np.savez_compressed(dst, probes=np.array(PROBES), layer=LAYER, irr=irr_in.view(np.uint16), dir=dir_in.view(np.uint16), old=old_in.view(np.uint16),
                        out_live=out["live"], out_bake=out["bake"], shared_rot=shared_rot, shared_dir=rdirs.view(np.uint16),
                        out_live_shared=out["live_shared"], out_bake_shared=out["bake_shared"])
    print("wrote", dst)


if __name__ == "__main__":
    main()
