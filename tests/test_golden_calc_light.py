"""The oracle's calcLight against the REFERENCE'S OWN COMPILED CODE.

tests/golden/calc_light_golden.npz holds inputs and results of the function calcLight (pbr_light.glsl:98-166) as compiled into
assets/irradiance_probe_update_rchit.spv, called directly with the small SPIR-V interpreter tests/golden/spirv_interp.py for 96
random surfaces (generator: tests/golden/make_calc_light_golden.py, run in the build container where the reference is mounted): the
directional branch the probe hit shader uses, and the point-light branch (distance attenuation, light colour) that extension E1
(ANEREND_DDGI_FLAG_POINT_LIGHTS, every BASELINE config with lights) builds its per-light term from.  pow(x, 5) is numpy's powf in the
interpreter and ((x x)(x x)) x in the oracle and the CUDA kernels (defined arithmetic), hence a tolerance of a few ulp."""
import os

import numpy as np

from anerend_b200 import layouts as L

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calc_light_golden.npz"))


def test_oracle_calc_light_equals_the_compiled_reference_function(oracle):
    rows, want = G["inputs"], G["out"]
    got = np.zeros_like(want)
    for k, r in enumerate(rows):
        light = None
        if r[26] == 0:
            light = np.zeros(1, L.LIGHT)
            light["world_pos"][0] = r[18:22]
            light["color"][0] = r[22:26]
        got[k] = oracle.calc_light(r[0:3], r[3:6], float(r[6]), float(r[7]), r[8:11], r[11:14], r[14:17], float(r[17]), light=light)
    assert np.array_equal(got == 0, want == 0)                       # unlit cases (N.L <= 0, beyond the light's range) are exactly 0 on both sides
    for pt in (0, 1):
        sel = rows[:, 26] == pt
        assert (want[sel].max(axis=1) > 0).sum() >= 30              # both branches are exercised
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-12)
    assert rel.max() <= 4e-6, f"max relative difference {rel.max():.3e}"
    assert (got.view(np.uint32) == want.view(np.uint32)).mean() > 0.95          # in fact every value is bit-identical here
