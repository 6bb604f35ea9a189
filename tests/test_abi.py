"""The C-ABI library loads and exports every symbol include/anerend_cuda_ddgi.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "anerend_cuda_ddgi.h")).read()
    return sorted(set(re.findall(r"\b(anerend_cuda_ddgi_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for n in ("create", "destroy", "reset", "upload_scene", "set_scene_ubo", "build_accel", "refit_accel", "translate",
              "trace_layer", "convolve_layer", "update", "frame", "download_atlas", "download_rays", "get_timings"):
        assert "anerend_cuda_ddgi_" + n in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build_cuda())
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"missing exports: {missing}"
    lib.anerend_cuda_ddgi_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.anerend_cuda_ddgi_version()


def test_null_handle_is_an_error_not_a_crash():
    """Header conventions: every entry point taking the handle first returns ANEREND_DDGI_E_INVALID (-1) for NULL (no device needed)."""
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build_cuda())
    src = open(os.path.join(ROOT, "include", "anerend_cuda_ddgi.h")).read()
    decls = re.findall(r"\bint\s+(anerend_cuda_ddgi_[a-z_0-9]+)\s*\(\s*(?:const\s+)?anerend_ddgi_t\s*\*\s*h\b([^;]*)\)\s*;", src)
    assert len(decls) >= 45
    for name, rest in decls:
        fn = getattr(lib, name)
        nargs = 1 + (rest.count(",") if rest.strip() else 0)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p] * nargs         # pointers and small integers all travel as 8-byte zeros here
        rc = fn(*([None] * nargs))
        assert rc == (0 if name.endswith("_destroy") else -1), name
    lib.anerend_cuda_ddgi_last_error.restype = ctypes.c_char_p
    lib.anerend_cuda_ddgi_last_error.argtypes = [ctypes.c_void_p]
    assert lib.anerend_cuda_ddgi_last_error(None) == b"null handle"


def test_oracle_mirrors_the_abi():
    src = open(os.path.join(ROOT, "oracle", "ddgi_oracle.h")).read()
    oracle_names = set(re.findall(r"\bddgi_oracle_([a-z_0-9]+)\s*\(", src))
    for n in ("create", "upload_scene", "set_scene_ubo", "build_accel", "refit_accel", "translate", "trace_layer",
              "convolve_layer", "update", "frame", "download_atlas", "download_rays", "download_ray_records"):
        assert n in oracle_names


def test_struct_layouts_match_the_reference_sizes():
    from anerend_b200 import layouts as L
    from anerend_b200.capi import Params, Timings, Stats, BvhInfo
    assert (L.VERTEX.itemsize, L.RENDERABLE.itemsize, L.MATERIAL.itemsize, L.MESH_INFO.itemsize, L.LIGHT.itemsize,
            L.SCENE_UBO.itemsize) == (96, 128, 64, 48, 32, 568)
    assert L.VERTEX.fields["color"][1] == 12 and L.VERTEX.fields["normal"][1] == 24 and L.VERTEX.fields["uv"][1] == 68
    assert L.SCENE_UBO.fields["camera_pos"][1] == 448 and L.SCENE_UBO.fields["light_dir"][1] == 480
    assert L.SCENE_UBO.fields["sky_color"][1] == 512 and L.SCENE_UBO.fields["flags"][1] == 556
    assert ctypes.sizeof(Params) == 56 and ctypes.sizeof(Timings) == 40 and ctypes.sizeof(Stats) == 88 and ctypes.sizeof(BvhInfo) == 44


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "anerend_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"ddgi_oracle|from\s+oracle|import\s+oracle|oracle/|libddgi_oracle", text), f


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import anerend_b200 as A
    with pytest.raises(A.DdgiError):
        A.Ddgi(A.make_params(4, 2, 64))
