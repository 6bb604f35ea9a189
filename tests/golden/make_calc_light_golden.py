#!/usr/bin/env python
"""Golden vectors for calcLight from the reference's OWN compiled code: calls the function calcLight (pbr_light.glsl:98-166) inside
assets/irradiance_probe_update_rchit.spv directly with tests/golden/spirv_interp.py — the directional branch the probe hit shader uses
and the point-light branch (distance attenuation, light colour) that extension E1 (ANEREND_DDGI_FLAG_POINT_LIGHTS) composes its
per-light term from — for random surfaces, and writes inputs and results to tests/golden/calc_light_golden.npz.

    python tests/golden/make_calc_light_golden.py [/root/reference/assets/irradiance_probe_update_rchit.spv]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from spirv_interp import Invocation, Module, Ptr, f32  # noqa: E402

CASES = 96


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/irradiance_probe_update_rchit.spv"
    m = Module(path)
    fid = next(f for f in m.functions if m.names.get(f, "").startswith("calcLight("))
    ubo_gid = next(g for g, n in m.names.items() if n == "ubo" and g in m.globals)
    ubo_t = m.types[m.globals[ubo_gid][0]][2]
    rng = np.random.RandomState(17)
    rows, outs = [], []
    for k in range(CASES):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        albedo = rng.uniform(0.02, 1.0, 3)
        metallic = float(rng.choice([0.0, 1.0, rng.uniform()]))
        roughness = float(rng.uniform(0.05, 1.0))
        wp = rng.uniform(-5, 5, 3)
        vp = wp + rng.normal(size=3) * rng.uniform(0.5, 8.0)
        ld = rng.normal(size=3); ld /= np.linalg.norm(ld)
        sun = float(rng.uniform(0.5, 8.0))
        # point light: position within ~1.3 ranges of the surface (so some cases sit beyond the range: attenuation exactly 0), colour, enabled
        rangev = float(rng.uniform(3.0, 15.0))
        lp = wp + rng.normal(size=3) * rangev * 0.45
        lcol = rng.uniform(0.2, 6.0, 3)
        directional = k % 2
        L = -ld if directional else (lp - wp) / np.linalg.norm(lp - wp)
        if k % 5 != 0 and float(np.dot(n, L)) < 0:            # most surfaces face their light; every fifth case keeps whatever came out
            n = -n
        row = np.array([*n, *albedo, metallic, roughness, *wp, *vp, *ld, sun, *lp, rangev, *lcol, 1.0, directional], np.float32)
        ubo = m.zero(ubo_t)
        ubo[m.struct_member(ubo_t, "lightDir")] = [f32(x) for x in row[14:17]] + [f32(0)]
        ubo[m.struct_member(ubo_t, "sunIntensity")] = f32(row[17])
        inv = Invocation(m, {"ubo": ubo})
        P = lambda v: Ptr([v], (0,))
        v3 = lambda a: [f32(x) for x in a]
        light = [v3(row[18:22]), v3(row[22:26])]
        res = inv.call(fid, [P(v3(row[0:3])), P(v3(row[3:6])), P(f32(row[6])), P(f32(row[7])), P(v3(row[8:11])), P(v3(row[11:14])),
                             P(light), P(int(directional)), P(True)])
        rows.append(row); outs.append([float(c) for c in res])
    rows, outs = np.array(rows, np.float32), np.array(outs, np.float32)
    pt = rows[:, 26] == 0
    print("directional: nonzero", int((outs[~pt].max(axis=1) > 0).sum()), "of", int((~pt).sum()), "| point: nonzero", int((outs[pt].max(axis=1) > 0).sum()), "of", int(pt.sum()))
    dst = os.path.join(HERE, "calc_light_golden.npz")
    np.savez_compressed(dst, inputs=rows, out=outs)
    print("wrote", dst)


if __name__ == "__main__":
    main()
