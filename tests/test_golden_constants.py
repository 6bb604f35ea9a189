"""Golden fixture from the reference's own build artefacts: the scalar constants of its precompiled SPIR-V modules
(assets/*.spv at the surveyed commit), extracted by tests/golden/make_spirv_constants.py into tests/golden/spirv_constants.json.
The reference has no test vectors for this path; these constants are the one thing its shipped binaries pin, and they are what
the oracle and the C ABI defaults must carry."""
import json
import os

import numpy as np

from anerend_b200 import make_params

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "spirv_constants.json")))
f32 = np.float32


def floats(module):
    return {f32(float(x)) for x in G[module]["float32"]}


def test_shipped_modules_carry_the_literals_the_path_is_built_on():
    rgen, rchit, rmiss, conv = (floats(m) for m in ("irradiance_probe_update_rgen.spv", "irradiance_probe_update_rchit.spv",
                                                    "irradiance_probe_update_rmiss.spv", "probe_conv_comp.spv"))
    assert {f32(0.1), f32(50.0), f32(1.6180339887), f32(3.14159265359), f32(2.0 * 3.14159265359), f32(256.0)} <= rgen      # tMin, tMax, phi, pi, 2 pi, N
    assert {f32(0.04), f32(1e-4), f32(0.001), f32(0.2), f32(1e-6), f32(3.1415926), f32(3.14159265359), f32(0.5)} <= rchit   # F0, spec eps, GGX, weights, pi/2 factor
    assert {f32(0.04045), f32(0.055), f32(1.055), f32(2.4), f32(12.92)} <= rmiss and {f32(0.04045), f32(12.92)} <= rchit     # toLinear
    assert {f32(0.96), f32(1e-6), f32(1.0), f32(2.0)} <= conv                                                              # hysteresis, weight threshold
    assert {8, 16, 64, 256} <= set(G["probe_conv_comp.spv"]["int32"]) and 640 in G["irradiance_probe_update_rchit.spv"]["int32"]   # texels, ray tile, grid, atlas width
    assert {f32(0.1), f32(50.0)} <= floats("shadow_rgen.spv") and floats("shadow_rmiss.spv") == {f32(1.0)}
    assert f32(1e6) in floats("mirror_reflections_rmiss.spv") and {f32(0.1), f32(50.0)} <= floats("mirror_reflections_rgen.spv")


def test_abi_defaults_are_the_shipped_constants():
    p = make_params(4, 2, 64)
    conv, rgen = floats("probe_conv_comp.spv"), floats("irradiance_probe_update_rgen.spv")
    assert f32(p.hysteresis) in conv and f32(p.hysteresis) == f32(0.96)
    assert f32(p.t_min) in rgen and f32(p.t_max) in rgen and (f32(p.t_min), f32(p.t_max)) == (f32(0.1), f32(50.0))
    assert tuple(p.probe_step) == (1.0, 2.0, 1.0)


def test_oracle_functions_evaluate_with_the_shipped_constants(oracle):
    # worldDirFromFibonacciIndex with the module's own 2*pi and golden-ratio words
    two_pi = [x for x in floats("irradiance_probe_update_rgen.spv") if abs(float(x) - 6.2831853) < 1e-5][0]
    phi = [x for x in floats("irradiance_probe_update_rgen.spv") if abs(float(x) - 1.618034) < 1e-5][0]
    n = 256
    i = np.arange(n, dtype=np.float32)
    theta = (two_pi * i) / phi
    ph = np.arccos(f32(1.0) - (f32(2.0) * (i + f32(0.5))) / f32(n)).astype(np.float32)
    want = np.stack([np.cos(theta) * np.sin(ph), np.sin(theta) * np.sin(ph), np.cos(ph)], 1).astype(np.float32)
    want /= np.linalg.norm(want, axis=1, keepdims=True)
    assert np.abs(oracle.fibonacci_dirs(n) - want).max() < 3e-6        # sin/cos of angles up to ~990 rad: a few fp32 ulps of theta
    # toLinear with the module's words
    c = {round(float(x), 5): x for x in floats("irradiance_probe_update_rmiss.spv")}
    x = np.linspace(0, 1, 101, dtype=np.float32)
    lin = np.where(x < c[0.04045], x / c[12.92], ((x + c[0.055]) / c[1.055]) ** c[2.4]).astype(np.float32)
    got = np.array([oracle.to_linear((v, v, v))[0] for v in x], np.float32)
    assert np.allclose(got, lin, rtol=3e-6, atol=1e-9)
