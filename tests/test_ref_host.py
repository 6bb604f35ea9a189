"""The byte layouts and the rotation matrix against the REFERENCE'S OWN HOST CODE, compiled.

oracle/_ref/libanerend_ref_host.so is built (oracle/Makefile; __graft_entry__.build()) from the reference's render/GpuBuffers.h,
render/Vertex.h and contrib/glm where they lie under /root/reference; tests/golden/ref_host_golden.json is what it reports
(generator: tests/golden/make_ref_host_golden.py).  Held to it: include/anerend_ddgi_layouts.h as gcc lays it out, the numpy views
the tests and the bench use, the UBO flag values, and the rotation matrices of the product (anerend_cuda_ddgi_make_rotation, host
code: no device needed) and of the oracle — bit for bit against glm::rotate."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np

from anerend_b200 import layouts as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_host_golden.json")))

# reference struct -> (C struct of include/anerend_ddgi_layouts.h, numpy view)
STRUCTS = {"Vertex": ("anerend_vertex_t", L.VERTEX), "GPURenderable": ("anerend_renderable_t", L.RENDERABLE),
           "GPUMaterialInfo": ("anerend_material_info_t", L.MATERIAL), "GPUMeshInfo": ("anerend_mesh_info_t", L.MESH_INFO),
           "GPULight": ("anerend_light_t", L.LIGHT), "GPUTileInfo": ("anerend_tile_info_t", None), "GPUSceneData": ("anerend_scene_ubo_t", L.SCENE_UBO)}


def _snake(name):
    return re.sub(r"(?<=[a-z0-9])([A-Z])", lambda m: "_" + m.group(1).lower(), name.lstrip("_")).lower()


def test_golden_is_what_the_compiled_reference_headers_report(oracle):
    ref = oracle.ref_layout()
    if ref is None:                       # neither /root/reference nor a built oracle/_ref: the committed fixture stands alone
        import pytest
        pytest.skip("oracle/_ref not available here")
    fields, flags = ref
    assert {k: list(v) for k, v in fields.items()} == G["layout"] and flags == G["flags"]
    for r in G["rotations"][::7]:
        assert [int(x) for x in oracle.ref_make_rotation(r["axis"], r["angle_deg"]).view(np.uint32)] == r["bits"]
    assert [float(x) for x in oracle.ref_render_option_defaults()] == G["render_option_defaults"]


def test_c_header_layouts_equal_the_reference_structs(tmp_path):
    rows = [(ref, c, _snake(k.split(".")[1])) for ref, (c, _) in STRUCTS.items() for k in G["layout"] if k.startswith(ref + ".")]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "anerend_ddgi_layouts.h"', "int main(void) {"]
    for ref, (c, _) in STRUCTS.items():
        src.append(f'  printf("{ref} 0 %zu\\n", sizeof({c}));')
    for ref, c, f in rows:
        src.append(f'  printf("{ref}.{f} %zu %zu\\n", offsetof({c}, {f}), sizeof((({c}*)0)->{f}));')
    for n in G["flags"]:
        src.append(f'  printf("flag.{n} %u 0\\n", (unsigned)(ANEREND_{n}));')
    src += ["  return 0;", "}"]
    cfile, exe = tmp_path / "layout.c", tmp_path / "layout"
    cfile.write_text("\n".join(src))
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(cfile)])
    ours = {a[0]: [int(a[1]), int(a[2])] for a in (line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())}
    want = {(k if "." not in k else k.split(".")[0] + "." + _snake(k.split(".")[1])): v for k, v in G["layout"].items()}
    want.update({"flag." + n: [v, 0] for n, v in G["flags"].items()})
    assert ours == want


def test_numpy_views_equal_the_reference_structs():
    for ref, (_, dt) in STRUCTS.items():
        if dt is None:
            continue
        assert dt.itemsize == G["layout"][ref][1], ref
        for k, (off, size) in G["layout"].items():
            if k.startswith(ref + "."):
                f = _snake(k.split(".")[1])
                assert dt.fields[f][1] == off and dt.fields[f][0].itemsize == size, k
    assert (L.UBO_RT_SHADOWS_FLAG, L.UBO_DDGI_FLAG, L.UBO_DDGI_MULTI_FLAG, L.UBO_SPECULAR_GI_FLAG, L.UBO_RT_ON_FLAG, L.UBO_BAKE_MODE_ON_FLAG) == tuple(
        G["flags"][n] for n in ("UBO_RT_SHADOWS_FLAG", "UBO_DDGI_FLAG", "UBO_DDGI_MULTI_FLAG", "UBO_SPECULAR_GI_FLAG", "UBO_RT_ON_FLAG", "UBO_BAKE_MODE_ON_FLAG"))


def test_rotation_matrices_equal_glm_rotate_bit_for_bit(oracle):
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build_cuda())
    out = np.empty(16, np.float32)
    for r in G["rotations"]:
        want = np.array(r["bits"], np.uint32)
        assert lib.anerend_cuda_ddgi_make_rotation(ctypes.c_int32(r["axis"]), ctypes.c_float(r["angle_deg"]), out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(out.view(np.uint32), want), ("product", r["axis"], r["angle_deg"])
        assert np.array_equal(oracle.make_rotation(r["axis"], r["angle_deg"]).view(np.uint32), want), ("oracle", r["axis"], r["angle_deg"])


def test_synthetic_scene_ubo_starts_from_the_reference_render_option_defaults():
    from anerend_b200 import scenes
    d = G["render_option_defaults"]            # RenderOptions.h:7-28, compiled
    u = scenes.make_ubo((0.0, 0.0, 0.0))[0]
    assert [float(u["sun_intensity"]), float(u["sky_intensity"]), float(u["exposure"])] == d[0:3]
    assert [float(np.float32(c)) for c in u["sky_color"][:3]] == d[3:6] and [float(u["bloom_thresh"]), float(u["bloom_knee"])] == d[6:8]
    # raytracingEnabled, ddgiEnabled, multiBounceDdgiEnabled are on by default (VulkanRenderer.cpp:1222-1232 turns them into flags)
    assert d[8:11] == [1.0, 1.0, 1.0]
    assert int(u["flags"]) & (L.UBO_RT_ON_FLAG | L.UBO_DDGI_FLAG | L.UBO_DDGI_MULTI_FLAG) == L.UBO_RT_ON_FLAG | L.UBO_DDGI_FLAG | L.UBO_DDGI_MULTI_FLAG
