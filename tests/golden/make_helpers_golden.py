#!/usr/bin/env python
"""Golden vectors for the small shader functions the path is built from, from the reference's OWN compiled code, called one by one with
tests/golden/spirv_interp.py:
  toLinear   (bindless.glsl:240-247)            inside assets/irradiance_probe_update_rchit.spv  — both branches and the 0.04045 boundary
  octEncode  (octahedron_helpers.glsl:10-20)    inside assets/irradiance_probe_update_rchit.spv  — all octants, axis directions, folds
  octDecode  (octahedron_helpers.glsl:25-34)    inside assets/probe_conv_comp.spv               — the 64 texel centres and random points
  pixelToUv  (helpers.glsl:1-4)                 inside assets/probe_conv_comp.spv
Written to tests/golden/helpers_golden.npz.

    python tests/golden/make_helpers_golden.py [/root/reference/assets]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from spirv_interp import Invocation, Module, Ptr, f32  # noqa: E402


def fn(m, prefix):
    return next(f for f in m.functions if m.names.get(f, "").startswith(prefix))


def main():
    assets = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets"
    hit = Module(os.path.join(assets, "irradiance_probe_update_rchit.spv"))
    conv = Module(os.path.join(assets, "probe_conv_comp.spv"))
    P = lambda v: Ptr([v], (0,))
    rng = np.random.RandomState(29)
    inv_hit, inv_conv = Invocation(hit, {}), Invocation(conv, {"gl_GlobalInvocationID": [0, 0, 0]})
    # toLinear: vec4 in, vec4 out (alpha passes through)
    srgb = np.concatenate([[0.0, 0.04045, np.nextafter(np.float32(0.04045), np.float32(0)), np.nextafter(np.float32(0.04045), np.float32(1)), 1.0, 0.5, 2.5, 7.0],
                           rng.uniform(0, 1, 120), rng.uniform(0, 0.06, 40), rng.uniform(1, 8, 24)]).astype(np.float32)
    f_lin = fn(hit, "toLinear(")
    lin = np.array([[float(c) for c in inv_hit.call(f_lin, [P([f32(v), f32(v * f32(0.5)), f32(1.0) - f32(v), f32(0.25)])])] for v in srgb], np.float32)
    # octEncode: unit vectors incl. the axes, the octant diagonals and vectors on the fold z = 0
    dirs = [np.eye(3)[k] * s for k in range(3) for s in (1.0, -1.0)]
    dirs += [np.array([sx, sy, sz]) / np.sqrt(3) for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)]
    dirs += [np.array([np.cos(a), np.sin(a), 0.0]) for a in np.linspace(0.1, 6.2, 12)]
    v = rng.normal(size=(160, 3)); dirs += list(v / np.linalg.norm(v, axis=1, keepdims=True))
    dirs = np.array(dirs, np.float32)
    f_enc = fn(hit, "octEncode(")
    enc = np.array([[float(c) for c in inv_hit.call(f_enc, [P([f32(c) for c in d])])] for d in dirs], np.float32)
    # octDecode: the 64 texel centres of a probe (probe_conv.comp:33-35) and random points of the square
    f_uv, f_dec = fn(conv, "pixelToUv("), fn(conv, "octDecode(")
    uv = np.array([[float(c) for c in inv_conv.call(f_uv, [P([i, j]), P([8, 8])])] for j in range(8) for i in range(8)], np.float32)
    pts = np.concatenate([uv * np.float32(2) - np.float32(1), rng.uniform(-1, 1, (136, 2)).astype(np.float32)]).astype(np.float32)
    dec = np.array([[float(c) for c in inv_conv.call(f_dec, [P([f32(c) for c in o])])] for o in pts], np.float32)
    dst = os.path.join(HERE, "helpers_golden.npz")
    np.savez_compressed(dst, srgb=srgb, linear=lin, dirs=dirs, oct_encode=enc, texel_uv=uv, oct_points=pts, oct_decode=dec)
    print("toLinear", len(srgb), "octEncode", len(dirs), "octDecode", len(pts), "-> wrote", dst)


if __name__ == "__main__":
    main()
