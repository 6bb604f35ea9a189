"""ctypes binding of the CPU oracle (oracle/ddgi_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under anerend_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from anerend_b200.capi import DdgiHandle, make_params  # generic prefix-agnostic wrapper (product -> oracle never)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle with the committed Makefile (g++ only; no dependency on /root/reference)."""
    so = os.path.join(_HERE, "libddgi_oracle.so")
    src = os.path.join(_HERE, "ddgi_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libddgi_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ddgi_oracle_half_to_float.restype = C.c_float
        _LIB.ddgi_oracle_half_to_float.argtypes = [C.c_uint16]
        _LIB.ddgi_oracle_float_to_half.restype = C.c_uint16
        _LIB.ddgi_oracle_float_to_half.argtypes = [C.c_float]
    return _LIB


REF_ROOT = os.environ.get("ANEREND_REFERENCE", "/root/reference")
_REF = None


def build_ref(force=False):
    """oracle/_ref/libanerend_ref_host.so: the reference's own struct headers and glm compiled from /root/reference (oracle/Makefile).
    Returns the path, or None when neither the reference nor a previously built library is there (the GPU box gets the built file)."""
    so = os.path.join(_HERE, "_ref", "libanerend_ref_host.so")
    if os.path.isdir(os.path.join(REF_ROOT, "contrib", "glm", "include")):
        src = os.path.join(_HERE, "ref_host_shim.cpp")
        if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "REF=" + REF_ROOT, "_ref/libanerend_ref_host.so"], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


def ref_lib():
    global _REF
    if _REF is None:
        so = build_ref()
        if so is None:
            return None
        _REF = C.CDLL(so)
        _REF.ref_host_layout.restype = C.c_size_t
        _REF.ref_host_layout.argtypes = [C.c_char_p, C.c_size_t]
    return _REF


def ref_layout():
    """({'Struct.field': (offset, size), 'Struct': (0, sizeof)}, {'FLAG': value}) from the reference's compiled headers, or None."""
    lib_ = ref_lib()
    if lib_ is None:
        return None
    n = lib_.ref_host_layout(None, 0)
    buf = C.create_string_buffer(n + 1)
    lib_.ref_host_layout(buf, n + 1)
    fields, flags = {}, {}
    for line in buf.value.decode().splitlines():
        a = line.split()
        if a[0] == "flag":
            flags[a[1]] = int(a[2])
        else:
            fields[a[0]] = (int(a[1]), int(a[2]))
    return fields, flags


def ref_render_option_defaults():
    out = np.empty(13, np.float32)
    ref_lib().ref_host_render_option_defaults(out.ctypes.data_as(C.c_void_p))
    return out


def ref_make_rotation(axis, deg):
    out = np.empty(16, np.float32)
    ref_lib().ref_host_make_rotation(C.c_int(axis), C.c_float(deg), out.ctypes.data_as(C.c_void_p))
    return out


class Oracle(DdgiHandle):
    def __init__(self, params):
        super().__init__(lib(), "ddgi_oracle_", params)

    def set_threads(self, n):
        return self._fn("set_threads")(self._h, C.c_int32(n))

    def set_sun_grid(self, resolution):
        self._call("set_sun_grid", C.c_int32(resolution))

    def set_brute_force(self, on):
        self._call("set_brute_force", C.c_int32(1 if on else 0))

    def set_rays(self, layer, irr, dr):
        irr = np.ascontiguousarray(irr, np.uint16)
        dr = np.ascontiguousarray(dr, np.uint16)
        self._call("set_rays", C.c_int32(layer), irr.ctypes.data_as(C.c_void_p), dr.ctypes.data_as(C.c_void_p),
                   C.c_size_t(irr.nbytes))

    def world_triangles(self):
        n = self.bvh_info().num_triangles
        out = np.empty((n, 3, 3), np.float32)
        self._call("download_world_triangles", out.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        return out

    def sample_texture(self, index, u, v):
        out = (C.c_float * 4)()
        self._call("sample_texture", C.c_int32(index), C.c_float(u), C.c_float(v), out)
        return np.array(out[:], np.float32)

    def calc_indirect(self, normal, albedo, metallic, roughness, world_pos, view_pos):
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        out = (C.c_float * 3)()
        self._call("calc_indirect", f3(normal), f3(albedo), C.c_float(metallic), C.c_float(roughness), f3(world_pos),
                   f3(view_pos), out)
        return np.array(out[:], np.float32)

    def any_hit(self, origin, direction, tmin=0.1, tmax=50.0):
        rc = self._fn("any_hit")(self._h, (C.c_float * 3)(*[float(x) for x in origin]), (C.c_float * 3)(*[float(x) for x in direction]),
                                 C.c_float(tmin), C.c_float(tmax))
        assert rc >= 0
        return bool(rc)

    def shade_hit(self, prim, u, v, origin, specular=False):
        """(irradiance [3], distance, lit mask) of the probe / mirror-reflection hit shader for one hit."""
        out, lit = (C.c_float * 4)(), C.c_uint32(0)
        self._call("shade_hit", C.c_uint32(prim), C.c_float(u), C.c_float(v), (C.c_float * 3)(*[float(x) for x in origin]),
                   C.c_int32(1 if specular else 0), out, C.byref(lit))
        return np.array(out[:3], np.float32), np.float32(out[3]), int(lit.value)

    def screen_ray(self, specular, px, py, depth, normal=(0.0, 1.0, 0.0)):
        """(origin, direction) shadow.rgen / mirror_reflections.rgen trace for a screen pixel."""
        o, d = (C.c_float * 3)(), (C.c_float * 3)()
        self._call("screen_ray", C.c_int32(1 if specular else 0), C.c_uint32(px), C.c_uint32(py), C.c_float(depth),
                   (C.c_float * 3)(*[float(x) for x in normal]), o, d)
        return np.array(o[:], np.float32), np.array(d[:], np.float32)

    def probe_ray(self, layer, probe, ray, rot):
        """(origin, direction, miss radiance) of one launch pixel, from the functions trace_layer uses (rgen:51-87, rmiss:17-21)."""
        o, d, m = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
        r = np.ascontiguousarray(rot, np.float32).reshape(16)
        self._call("probe_ray", C.c_int32(layer), C.c_int32(probe), C.c_int32(ray), r.ctypes.data_as(C.c_void_p), o, d, m)
        return np.array(o[:], np.float32), np.array(d[:], np.float32), np.array(m[:], np.float32)

# ---- single shader functions (KATs) ----------------------------------------------------------------
def _f(n, v):
    return (C.c_float * n)(*[float(x) for x in v])


def oct_encode(v):
    out = (C.c_float * 2)()
    lib().ddgi_oracle_oct_encode(_f(3, v), out)
    return np.array(out[:], np.float32)


def oct_decode(o):
    out = (C.c_float * 3)()
    lib().ddgi_oracle_oct_decode(_f(2, o), out)
    return np.array(out[:], np.float32)


def to_linear(c):
    out = (C.c_float * 3)()
    lib().ddgi_oracle_to_linear(_f(3, c), out)
    return np.array(out[:], np.float32)


def fibonacci_dirs(n):
    out = np.empty((n, 3), np.float32)
    lib().ddgi_oracle_fibonacci_dirs(C.c_int32(n), out.ctypes.data_as(C.c_void_p))
    return out


def make_rotation(axis, deg):
    out = np.empty(16, np.float32)
    rc = lib().ddgi_oracle_make_rotation(C.c_int32(axis), C.c_float(deg), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def float_to_half(f):
    return int(lib().ddgi_oracle_float_to_half(C.c_float(f)))


def half_to_float(h):
    return float(lib().ddgi_oracle_half_to_float(C.c_uint16(h)))


def calc_light(normal, albedo, metallic, roughness, world_pos, view_pos, light_dir, sun_intensity, light=None):
    out = (C.c_float * 3)()
    lp = None
    directional = 1
    if light is not None:
        arr = np.ascontiguousarray(light)
        lp = arr.ctypes.data_as(C.c_void_p)
        directional = 0
    lib().ddgi_oracle_calc_light(_f(3, normal), _f(3, albedo), C.c_float(metallic), C.c_float(roughness),
                                 _f(3, world_pos), _f(3, view_pos), lp, C.c_int32(directional), _f(3, light_dir),
                                 C.c_float(sun_intensity), out)
    return np.array(out[:], np.float32)


def intersect(orig, d, v0, v1, v2, tmin=0.1, tmax=50.0):
    t, u, v = C.c_float(), C.c_float(), C.c_float()
    hit = lib().ddgi_oracle_intersect(_f(3, orig), _f(3, d), _f(3, v0), _f(3, v1), _f(3, v2), C.c_float(tmin),
                                      C.c_float(tmax), C.byref(t), C.byref(u), C.byref(v))
    return bool(hit), t.value, u.value, v.value
