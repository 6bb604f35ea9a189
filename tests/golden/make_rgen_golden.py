#!/usr/bin/env python
"""Golden vectors for ray generation and the miss shader from the reference's OWN compiled shaders: executes
assets/irradiance_probe_update_rgen.spv (launch pixel -> probe and ray index, probe origin in live and bake mode,
worldDirFromFibonacciIndex, mat3(randomRot) * dir, the traceRayEXT arguments, the two imageStores) and
assets/irradiance_probe_update_rmiss.spv (sky radiance) with tests/golden/spirv_interp.py and writes what they produce to
tests/golden/rgen_golden.npz.  The trace call is a callback that records its arguments and returns a marker radiance.

sin / cos / acos in the interpreter are numpy's float32 functions, in the oracle the host libm (the direction table is
an input both sides share), so directions agree to a few ulp, not bit for bit; everything else in these shaders is exact.

    python tests/golden/make_rgen_golden.py [/root/reference/assets]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from spirv_interp import Invocation, Module, f32  # noqa: E402

GRID = (64, 8, 256)                # the grid and ray count the shipped binary is compiled for (probe_helpers.glsl:1-7)
PIXELS_PER_CASE = 48
UBO_BAKE_MODE_ON_FLAG = 1 << 11    # scene_ubo.glsl


class Image:
    def __init__(self):
        self.px = {}

    def write(self, xy, v):
        self.px[tuple(xy)] = v


def cases():
    """(name, camera, flags, bake tile info, layer, rotation (axis, degrees))"""
    return [
        ("live_pos", (5.3, 4.0, 9.8), 0, (0, 0, 0, 0), 0, (0, 37.5)),
        ("live_neg", (-7.6, 1.0, -12.9), 0, (0, 0, 0, 0), 3, (1, 211.25)),          # int() truncates toward zero (quirk Q1)
        ("live_mix", (-0.4, 2.0, 30.2), 0, (0, 0, 0, 0), 7, (2, 359.0)),
        ("bake", (5.3, 4.0, 9.8), UBO_BAKE_MODE_ON_FLAG, (-2, 3, 32, 0), 5, (1, 90.0)),
        ("bake_big", (0.0, 0.0, 0.0), UBO_BAKE_MODE_ON_FLAG, (4, -1, 48, 0), 2, (0, 123.0)),
    ]


def main():
    from oracle import binding as ob
    assets = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets"
    m = Module(os.path.join(assets, "irradiance_probe_update_rgen.spv"))
    nxz, ny, nr = GRID
    ubo_gid = next(g for g, n in m.names.items() if n == "ubo" and g in m.globals)
    ubo_t = m.types[m.globals[ubo_gid][0]][2]
    pay_gid = next(g for g, n in m.names.items() if n == "payload" and g in m.globals)
    rng = np.random.RandomState(7)
    marker = [f32(0.25), f32(1.5), f32(3.0)]
    rows = []
    for ci, (name, cam, flags, tile, layer, (axis, deg)) in enumerate(cases()):
        rot = ob.make_rotation(axis, deg)                      # glm::rotate as the pass builds it (an INPUT of the shader)
        ubo = m.zero(ubo_t)
        for fname, val in (("cameraPos", [f32(x) for x in cam] + [f32(1)]), ("bakeTileInfo", [int(x) for x in tile]), ("flags", int(flags))):
            ubo[m.struct_member(ubo_t, fname)] = val
        pc = [[[f32(rot[4 * c + r]) for r in range(4)] for c in range(4)], layer]
        # corners of the launch, then random pixels
        pix = [(0, 0), (16 * nxz - 1, 16 * nxz - 1), (15, 16), (16, 15)] + [(int(rng.randint(0, 16 * nxz)), int(rng.randint(0, 16 * nxz))) for _ in range(PIXELS_PER_CASE - 4)]
        for (x, y) in pix:
            calls, irr_im, dir_im = [], Image(), Image()

            def trace(org, d, tmin, tmax, payload, _calls=calls):
                _calls.append((org, d, tmin, tmax))
                s = payload.load(); s[0] = list(marker); payload.store(s)

            inv = Invocation(m, {"gl_LaunchIDEXT": [x, y, 0], "ubo": ubo, "pc": pc, "irradianceIm": irr_im, "dirIm": dir_im,
                                 "payload": [[f32(9)] * 3, [f32(0)] * 3, f32(0)], "traceRay": trace})
            inv.run()
            assert len(calls) == 1 and list(irr_im.px) == [(x, y)] and list(dir_im.px) == [(x, y)]
            org, d, tmin, tmax = calls[0]
            assert [float(c) for c in irr_im.px[(x, y)]] == [0.25, 1.5, 3.0, 1.0]            # payload.irradiance, alpha 1 (rgen:109)
            assert [float(c) for c in dir_im.px[(x, y)]] == [float(c) for c in d] + [1.0]      # the traced direction, alpha 1 (rgen:108)
            rows.append((ci, x, y, layer, *[float(c) for c in org], *[float(c) for c in d], float(tmin), float(tmax)))
        print(name, "origin of pixel (0,0):", rows[-PIXELS_PER_CASE][4:7])
    # ---- rmiss ------------------------------------------------------------------------------------------
    mm = Module(os.path.join(assets, "irradiance_probe_update_rmiss.spv"))
    ubo_gid = next(g for g, n in mm.names.items() if n == "ubo" and g in mm.globals)
    ubo_t = mm.types[mm.globals[ubo_gid][0]][2]
    pay_gid = next(g for g, n in mm.names.items() if n == "payload" and g in mm.globals)
    skies = [((0.5, 0.7, 1.0), 1.0), ((0.2, 0.4, 0.9), 0.35), ((1.0, 1.0, 1.0), 2.0), ((0.0, 0.003, 0.04), 1.0), ((0.9, 0.1, 0.5), 0.0)]
    miss = []
    for col, inten in skies:
        ubo = mm.zero(ubo_t)
        ubo[mm.struct_member(ubo_t, "skyColor")] = [f32(c) for c in col] + [f32(1)]
        ubo[mm.struct_member(ubo_t, "skyIntensity")] = f32(inten)
        inv = Invocation(mm, {"ubo": ubo, "payload": [[f32(9)] * 3, [f32(0)] * 3, f32(7)]})
        inv.run()
        res = inv.g[pay_gid].load()
        assert float(res[2]) == 0.0                                                           # payload.shadow = 0 (rmiss:19)
        miss.append([float(c) for c in res[0]])
        print("sky", col, inten, "->", miss[-1])
    dst = os.path.join(HERE, "rgen_golden.npz")
    cs = cases()
    np.savez_compressed(dst, grid=np.array(GRID), rows=np.array(rows, np.float64),
                        case_camera=np.array([c[1] for c in cs], np.float32), case_flags=np.array([c[2] for c in cs], np.int64),
                        case_tile=np.array([c[3] for c in cs], np.int64), case_rot=np.array([c[5] for c in cs], np.float64),
                        sky=np.array([list(c) + [i] for c, i in skies], np.float32), miss=np.array(miss, np.float32))
    print("wrote", dst)


if __name__ == "__main__":
    main()
