"""The oracle's ray generation and miss shading against the REFERENCE'S OWN COMPILED SHADERS.

tests/golden/rgen_golden.npz holds what assets/irradiance_probe_update_rgen.spv hands to traceRayEXT and stores to its two images
for 240 launch pixels (five cases: live mode with positive / negative / mixed camera positions, two bake-mode tiles; the reference's
native 64 x 8 grid with 256 rays per probe), and what assets/irradiance_probe_update_rmiss.spv returns for five skies — both binaries
executed with the small SPIR-V interpreter tests/golden/spirv_interp.py (generator: tests/golden/make_rgen_golden.py, run in the
build container where the reference is mounted).  Origins, tMin / tMax and the pixel -> (probe, ray) mapping are exact; directions
go through sin / cos / acos (numpy float32 in the interpreter, libm in the oracle's host table) and the sky through pow(), so those
are held to a few ulp.  The CUDA kernels are held to the oracle bit for bit by the -m gpu tests."""
import os

import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rgen_golden.npz"))
NXZ, NY, NR = (int(v) for v in G["grid"])


def _oracle_for_case(oracle, ci, records=False):
    s = scenes.empty_scene(grid=(NXZ, NY, NR), camera=tuple(float(c) for c in G["case_camera"][ci]))
    s.ubo["flags"] |= np.uint32(G["case_flags"][ci])
    s.ubo["bake_tile_info"][0] = G["case_tile"][ci]
    o = oracle.Oracle(make_params(NXZ, NY, NR, flags=L.FLAG_RAY_RECORDS if records else 0))
    o.set_threads(0)
    o.upload_scene(s); o.set_scene_ubo(s.ubo); o.build_accel()
    return o


def test_oracle_ray_generation_equals_the_compiled_reference_shader(oracle):
    p = make_params(NXZ, NY, NR)
    rows = G["rows"]
    assert set(rows[:, 10]) == {np.float64(np.float32(p.t_min))} and set(rows[:, 11]) == {np.float64(p.t_max)} and p.t_max == 50.0   # rgen:91-92
    worst = 0.0
    for ci in range(len(G["case_camera"])):
        o = _oracle_for_case(oracle, ci)
        axis, deg = G["case_rot"][ci]
        rot = oracle.make_rotation(int(axis), float(deg))
        for row in rows[rows[:, 0] == ci]:
            x, y, layer = int(row[1]), int(row[2]), int(row[3])
            probe = (y // 16) * NXZ + x // 16                       # rgen:51-58 with PROBE_IRR_DIR_PIX_SIZE = 16
            ray = (y % 16) * 16 + x % 16
            org, d, _ = o.probe_ray(layer, probe, ray, rot)
            assert org.tolist() == row[4:7].tolist(), (ci, x, y)    # probe origin: bit-exact, live (Q1 truncation) and bake mode
            worst = max(worst, float(np.abs(d.astype(np.float64) - row[7:10]).max()))
        o.close()
    assert worst <= 4e-7, worst                                     # ~3 ulp of a unit-vector component


def test_oracle_ray_images_use_the_reference_pixel_layout(oracle):
    # one whole layer of the 'live_neg' case through trace_layer: ProbeRaysDirTex / ProbeRaysIrrTex at the golden's pixels
    ci = 1
    o = _oracle_for_case(oracle, ci, records=True)
    rows = G["rows"][G["rows"][:, 0] == ci]
    layer = int(rows[0, 3])
    axis, deg = G["case_rot"][ci]
    o.trace_layer(layer, oracle.make_rotation(int(axis), float(deg)))
    irr, dr = o.download_rays(layer)
    sky = oracle.to_linear([0.2, 0.5, 0.7]).astype(np.float16)
    same = 0
    for row in rows:
        x, y = int(row[1]), int(row[2])
        want = row[7:10].astype(np.float32).astype(np.float16)
        got = dr[y, x].view(np.float16)
        assert got[3] == 1.0 and np.abs(got[:3].astype(np.float32) - want.astype(np.float32)).max() <= 2.0 ** -10       # one fp16 ulp below 1
        same += int(np.array_equal(got[:3], want))
        assert np.array_equal(irr[y, x].view(np.float16), np.array([*sky, 1.0], np.float16))     # every ray misses: rmiss via rgen:109
    assert same >= 0.9 * len(rows)
    o.close()


def test_oracle_miss_radiance_equals_the_compiled_miss_shader(oracle):
    for (r, g, b, inten), want in zip(G["sky"], G["miss"]):
        s = scenes.empty_scene(grid=(4, 2, 64))
        s.ubo["sky_color"][0] = (r, g, b, 1.0)
        s.ubo["sky_intensity"] = inten
        o = oracle.Oracle(make_params(4, 2, 64, flags=L.FLAG_RAY_RECORDS))
        o.upload_scene(s); o.set_scene_ubo(s.ubo); o.build_accel()
        _, _, miss = o.probe_ray(0, 0, 0, np.eye(4, dtype=np.float32))
        o.update(-1, o.rotations(scenes.rotation_sequence(1, 2)))
        rec = o.download_ray_records(1)
        o.close()
        assert np.array_equal(rec["radiance"], np.broadcast_to(miss, rec["radiance"].shape))       # what trace_layer writes is missRadiance
        np.testing.assert_allclose(miss, want, rtol=2e-6, atol=0)                                  # pow(x, 2.4): numpy powf vs libm
        assert (want == 0).tolist() == (miss == 0).tolist()


def test_product_kernel_logic_ray_images_match_the_compiled_shader(make_emu):
    # the same check on the PRODUCT's kernel bodies (k_make_dirs, ray set-up, miss path) run by the host harness
    _ray_images_against_the_binary(make_emu)


@pytest.mark.gpu
def test_cuda_ray_images_match_the_compiled_shader(make_gpu):
    _ray_images_against_the_binary(make_gpu)


def _ray_images_against_the_binary(make_emu):
    ci = 1
    s = scenes.empty_scene(grid=(NXZ, NY, NR), camera=tuple(float(c) for c in G["case_camera"][ci]))
    s.ubo["sky_color"][0] = (*G["sky"][0][:3], 1.0)
    s.ubo["sky_intensity"] = G["sky"][0][3]
    sky = G["miss"][0].astype(np.float16).astype(np.float32)
    e = make_emu(make_params(NXZ, NY, NR))
    e.upload_scene(s); e.set_scene_ubo(s.ubo); e.build_accel()
    rows = G["rows"][G["rows"][:, 0] == ci]
    layer = int(rows[0, 3])
    axis, deg = G["case_rot"][ci]
    e.trace_layer(layer, e.make_rotation(int(axis), float(deg)))
    irr, dr = e.download_rays(layer)
    same = 0
    for row in rows:
        x, y = int(row[1]), int(row[2])
        want = row[7:10].astype(np.float32).astype(np.float16)
        got = dr[y, x].view(np.float16)
        assert got[3] == 1.0 and np.abs(got[:3].astype(np.float32) - want.astype(np.float32)).max() <= 2.0 ** -10
        same += int(np.array_equal(got[:3], want))
        np.testing.assert_allclose(irr[y, x].view(np.float16)[:3].astype(np.float32), sky, rtol=2.0 ** -10, atol=0)      # rmiss through rgen:109
    assert same >= 0.9 * len(rows)
    e.close()
