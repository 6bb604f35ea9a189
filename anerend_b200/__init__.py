"""anerend_b200 — B200-native DDGI irradiance-probe update behind AneRend's pass interface.

The product is the C-ABI library `libanerend_cuda_ddgi.so` (csrc/, include/anerend_cuda_ddgi.h): hand-written
sm_100a CUDA for the chain IrradianceProbeTrans -> IrradianceProbeRT -> IrradianceProbeConvolve.  This package
is the thin host side: ctypes binding (`Ddgi`), synthetic scene generators (`scenes`) and the multi-GPU slab helper
(`sharding`); the reference-shaped pass classes are C++ (`include/anerend_ddgi_passes.hpp`).

There is no CPU fallback: creating a `Ddgi` without the built library or without a CUDA device raises.
"""
import ctypes as _C
import os as _os

from . import layouts, scenes  # noqa: F401
from .capi import DdgiError, DdgiHandle, make_params  # noqa: F401

_HERE = _os.path.dirname(_os.path.abspath(__file__))
LIB_PATH = _os.path.join(_HERE, "libanerend_cuda_ddgi.so")
_LIB = None


def load_library():
    """Loads the CUDA library; fails loudly when it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _LIB
    if _LIB is None:
        if not _os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                               "(there is no CPU fallback for the DDGI probe update)")
        _LIB = _C.CDLL(LIB_PATH)
        _LIB.anerend_cuda_ddgi_version.restype = _C.c_char_p
    return _LIB


class Ddgi(DdgiHandle):
    """The CUDA DDGI probe volume (anerend_cuda_ddgi_* entry points)."""

    def __init__(self, params):
        super().__init__(load_library(), "anerend_cuda_ddgi_", params)

    def timings(self):
        from .capi import Timings
        t = Timings()
        self._call("get_timings", _C.byref(t))
        return {k: (list(getattr(t, k)) if k == "rt_kernel_ms" else getattr(t, k)) for k, _ in t._fields_}

    def atlas_device_ptr(self):
        p, n = _C.c_void_p(), _C.c_size_t()
        self._call("atlas_device_ptr", _C.byref(p), _C.byref(n))
        return p.value, n.value

    def stream(self):
        s = _C.c_void_p()
        self._call("stream", _C.byref(s))
        return s.value or 0

    def export_atlas_device(self, d_dst, pitch_bytes):
        """Publishes the atlas into engine-owned device memory (the frame-graph image "ProbeRaysConvTex", linear rows)."""
        self._call("export_atlas_device", _C.c_void_p(d_dst), _C.c_size_t(pitch_bytes))

    def set_peer_atlases(self, ptrs):
        arr = (_C.c_void_p * len(ptrs))(*ptrs)
        self._call("set_peer_atlases", arr, _C.c_uint32(len(ptrs)))

    def exchange_export(self):
        blob = _C.create_string_buffer(128)
        self._call("exchange_export", blob, _C.c_size_t(128))
        return blob.raw

    def exchange_connect(self, blobs, world, rank):
        """`blobs`: the exchange_export() results of all ranks, in rank order."""
        data = b"".join(blobs)
        assert len(data) == 128 * world
        self._call("exchange_connect", _C.c_char_p(data), _C.c_uint32(world), _C.c_uint32(rank))

    def exchange_fence(self):
        self._call("exchange_fence")

    def exchange_disconnect(self):
        self._call("exchange_disconnect")

    def specular_gi_device(self, d_depth, d_geom0, width, height, d_out):
        self._call("specular_gi_device", _C.c_void_p(d_depth), _C.c_void_p(d_geom0), _C.c_uint32(width), _C.c_uint32(height), _C.c_void_p(d_out))

    def shadow_rt_device(self, d_depth, width, height, d_out):
        self._call("shadow_rt_device", _C.c_void_p(d_depth), _C.c_uint32(width), _C.c_uint32(height), _C.c_void_p(d_out))

    def indirect_diffuse_device(self, d_points, count, view_pos, clamp_to_edge, d_out):
        vp = (_C.c_float * 3)(*[float(x) for x in view_pos])
        self._call("indirect_diffuse_device", _C.c_void_p(d_points), _C.c_uint32(count), vp, _C.c_int32(1 if clamp_to_edge else 0), _C.c_void_p(d_out))

    def refit_accel_device(self, d_ptr, count):
        self._call("refit_accel_device", _C.c_void_p(d_ptr), _C.c_uint32(count))
