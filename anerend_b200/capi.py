"""ctypes binding of the anerend_cuda_ddgi_* C ABI (include/anerend_cuda_ddgi.h).

`DdgiHandle` is prefix-generic (library handle + symbol prefix are constructor arguments), so the test
infrastructure can drive a second implementation of the same ABI with identical calls.  The product
itself only ever binds its own library: `anerend_b200.Ddgi` always binds libanerend_cuda_ddgi.so.
"""
import ctypes as C
import numpy as np

from . import layouts as L


class Params(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("num_probes_xz", C.c_int32), ("num_probes_y", C.c_int32),
                ("rays_per_probe", C.c_int32), ("probe_step", C.c_float * 3), ("t_min", C.c_float),
                ("t_max", C.c_float), ("hysteresis", C.c_float), ("flags", C.c_uint32),
                ("slab_begin", C.c_int32), ("slab_end", C.c_int32), ("device", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [("trans_ms", C.c_float), ("rt_ms", C.c_float), ("convolve_ms", C.c_float),
                ("accel_ms", C.c_float), ("total_ms", C.c_float), ("kernel_launches", C.c_uint32),
                ("rt_kernel_ms", C.c_float * 4)]


class Stats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("primary_hits", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("primary_nodes_visited", C.c_uint64), ("primary_tris_tested", C.c_uint64),
                ("shadow_nodes_visited", C.c_uint64), ("shadow_tris_tested", C.c_uint64), ("shadow_rays_prefiltered", C.c_uint64),
                ("shadow_rays_candidate", C.c_uint64), ("primary_level_changes", C.c_uint64), ("shadow_level_changes", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class BvhInfo(C.Structure):
    _fields_ = [("num_triangles", C.c_uint32), ("num_nodes", C.c_uint32), ("node_bytes", C.c_uint32),
                ("tri_bytes", C.c_uint32), ("num_instances", C.c_uint32),
                ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3)]


class TwoLevelInfo(C.Structure):       # anerend_ddgi_two_level_info_t
    _fields_ = [("num_meshes", C.c_uint32), ("stored_triangles", C.c_uint32), ("blas_nodes", C.c_uint32), ("tlas_nodes", C.c_uint32),
                ("tlas_depth", C.c_uint32), ("blas_depth_max", C.c_uint32), ("accel_bytes", C.c_uint64), ("tlas_build_ms", C.c_float),
                ("_pad", C.c_uint32)]


class TextureDesc(C.Structure):        # anerend_texture_desc_t
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("clamp_to_edge", C.c_uint32)]


class DdgiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def make_params(num_probes_xz, num_probes_y, rays_per_probe, probe_step=(1.0, 2.0, 1.0), t_min=0.1, t_max=50.0,
                hysteresis=0.96, flags=0, slab=None, device=-1):
    """Defaults are the reference's compile-time constants (probe_helpers.glsl:1-7, rgen:91-92, probe_conv.comp:85)."""
    p = Params()
    p.struct_size = C.sizeof(Params)
    p.num_probes_xz, p.num_probes_y, p.rays_per_probe = num_probes_xz, num_probes_y, rays_per_probe
    p.probe_step = (C.c_float * 3)(*probe_step)
    p.t_min, p.t_max, p.hysteresis, p.flags = t_min, t_max, hysteresis, flags
    p.slab_begin, p.slab_end = slab if slab is not None else (0, num_probes_y)
    p.device = device
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class DdgiHandle:
    """One DDGI probe volume: state of the three passes plus the scene/acceleration structure."""

    def __init__(self, lib, prefix, params):
        self._lib, self._pfx = lib, prefix
        self.params = params
        self._h = C.c_void_p()
        rc = self._fn("create")(C.byref(params), C.byref(self._h))
        if rc != 0:
            raise DdgiError(rc, f"{prefix}create failed (invalid parameters or no CUDA device)")
        w, h = C.c_int32(), C.c_int32()
        self._call("atlas_dims", C.byref(w), C.byref(h))
        self.atlas_width, self.atlas_height = w.value, h.value
        sx, sy = C.c_int32(), C.c_int32()
        self._call("ray_tile", C.byref(sx), C.byref(sy))
        self.ray_tile = (sx.value, sy.value)
        self.n_xz, self.n_y, self.n_rays = params.num_probes_xz, params.num_probes_y, params.rays_per_probe

    def _fn(self, name):
        f = getattr(self._lib, self._pfx + name)
        f.restype = C.c_int
        return f

    def _call(self, name, *args):
        rc = self._fn(name)(self._h, *args)
        if rc != 0:
            e = getattr(self._lib, self._pfx + "last_error")
            e.restype = C.c_char_p
            raise DdgiError(rc, (e(self._h) or b"").decode())
        return rc

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- scene ---------------------------------------------------------------------------------
    def upload_textures(self, textures):
        """`textures`: list of (pixels uint8 [h, w, c], ANEREND_TEX_* format, clamp_to_edge); index = bindless index."""
        n = len(textures)
        arr = (TextureDesc * max(n, 1))()
        keep = []
        for i, (px, fmt, clamp) in enumerate(textures):
            px = np.ascontiguousarray(px, np.uint8)
            keep.append(px)
            arr[i].pixels = px.ctypes.data_as(C.c_void_p)
            arr[i].width, arr[i].height, arr[i].format, arr[i].clamp_to_edge = px.shape[1], px.shape[0], int(fmt), 1 if clamp else 0
        self._call("upload_textures", arr, C.c_uint32(n))

    def update_vertices(self, vertices, first=0):
        """Overwrites a range of the giga vertex buffer (deforming / skinned meshes); follow with refit_accel()."""
        v = _c(vertices, L.VERTEX)
        self._call("update_vertices", _ptr(v), C.c_uint32(first), C.c_uint32(len(v)))

    def upload_tile_infos(self, atlas_indices):
        """tileInfoBuffer: bindless index of each paged tile's baked atlas (-1 = none), 21 x 21 entries around the camera."""
        arr = _c(atlas_indices, np.int32).reshape(-1)
        self._call("upload_tile_infos", _ptr(arr) if arr.size else None, C.c_uint32(arr.size))

    def upload_scene(self, scene):
        s = scene
        self._keep = s
        self.upload_textures(getattr(s, "textures", []))
        self._call("upload_scene",
                   _ptr(s.vertices), C.c_uint32(len(s.vertices)), _ptr(s.indices), C.c_uint32(len(s.indices)),
                   _ptr(s.meshes), _ptr(s.mesh_num_indices), C.c_uint32(len(s.meshes)),
                   _ptr(s.materials), C.c_uint32(len(s.materials)),
                   _ptr(s.renderables), C.c_uint32(len(s.renderables)),
                   _ptr(s.model_indices), C.c_uint32(len(s.model_indices)),
                   _ptr(s.material_indices), C.c_uint32(len(s.material_indices)),
                   _ptr(s.lights), C.c_uint32(len(s.lights)))

    def set_scene_ubo(self, ubo):
        ubo = _c(ubo, L.SCENE_UBO)
        self._call("set_scene_ubo", _ptr(ubo))

    def build_accel(self):
        self._call("build_accel")

    def refit_accel(self, renderables=None):
        if renderables is None:
            self._call("refit_accel", None, C.c_uint32(0))
        else:
            r = _c(renderables, L.RENDERABLE)
            self._call("refit_accel", _ptr(r), C.c_uint32(len(r)))

    def apply_camera_cull(self, range_m=50.0):
        """The cull bit of cull.comp:296-303 for the current scene UBO; returns how many renderables flipped (then build_accel again)."""
        n = C.c_uint32(0)
        self._call("apply_camera_cull", C.c_float(range_m), C.byref(n))
        return int(n.value)

    def bvh_info(self):
        b = BvhInfo()
        self._call("get_bvh_info", C.byref(b))
        return b

    def set_pipeline(self, on):
        """Closest-hit trace of an update on its own stream, under the previous update's tail (results unchanged); needs set_timers(0)."""
        self._call("set_pipeline", C.c_int32(1 if on else 0))

    def set_timers(self, level):
        """0 none, 1 around the passes (default), 2 also around every kernel of the RT pass (timings()['rt_kernel_ms'])."""
        self._call("set_timers", C.c_int32(level))

    def two_level_info(self):
        b = TwoLevelInfo()
        self._call("get_two_level_info", C.byref(b))
        return b

    # ---- passes ----------------------------------------------------------------------------------
    def reset(self):
        self._call("reset")

    def translate(self):
        self._call("translate")

    def trace_layer(self, layer, rot):
        rot = _c(rot, np.float32).reshape(16)
        self._call("trace_layer", C.c_int32(layer), _ptr(rot))

    def convolve_layer(self, layer):
        self._call("convolve_layer", C.c_int32(layer))

    def update(self, layer, rots):
        rots = _c(rots, np.float32).reshape(-1)
        need = 16 if layer >= 0 else 16 * self.n_y
        if rots.size != need:
            raise ValueError(f"update(layer={layer}) needs {need} floats of rotation, got {rots.size}")
        self._call("update", C.c_int32(layer), _ptr(rots))

    def update_layers(self, layers, rots):
        """Jacobi update of an arbitrary layer subset; rots = n_y matrices indexed by absolute layer."""
        layers = _c(layers, np.int32).reshape(-1)
        rots = _c(rots, np.float32).reshape(-1)
        if rots.size != 16 * self.n_y:
            raise ValueError(f"update_layers needs {16 * self.n_y} floats of rotation, got {rots.size}")
        self._call("update_layers", _ptr(layers), C.c_uint32(layers.size), _ptr(rots))

    def frame(self, rot):
        rot = _c(rot, np.float32).reshape(16)
        layer = C.c_int32(-1)
        self._call("frame", _ptr(rot), C.byref(layer))
        return layer.value

    def sync(self):
        if hasattr(self._lib, self._pfx + "sync"):
            self._call("sync")

    # ---- results ---------------------------------------------------------------------------------
    def specular_gi(self, depth, geom0):
        """Frame-graph pass "SpecularGIRT": depth float32 [h, w], geom0 float32 [h, w, 4] (normal in rgb) -> uint16 [h, w, 4]
        (RGBA16F bit patterns: irradiance, hit distance)."""
        depth = _c(depth, np.float32)
        geom0 = _c(geom0, np.float32)
        out = np.full(depth.shape + (4,), 0xFFFF, np.uint16)          # untouched when the pass is disabled
        self._call("specular_gi", _ptr(depth), _ptr(geom0), C.c_uint32(depth.shape[1]), C.c_uint32(depth.shape[0]), _ptr(out))
        return out

    def shadow_rt(self, depth):
        """Frame-graph pass "ShadowRT": depth float32 [height, width] -> r32f [height, width] (1 lit / sky, 0 shadowed)."""
        depth = _c(depth, np.float32)
        out = np.full(depth.shape, -1.0, np.float32)          # untouched when UBO_RT_SHADOWS_FLAG is off
        self._call("shadow_rt", _ptr(depth), C.c_uint32(depth.shape[1]), C.c_uint32(depth.shape[0]), _ptr(out))
        return out

    def indirect_diffuse(self, points, view_pos, clamp_to_edge=False):
        """Atlas consumer: calcIndirectDiffuseLight for surface points [n, 12] float32 (pos, normal, albedo, metallic,
        roughness, unused) -> [n, 4] float32 (rgb, 0).  clamp_to_edge: the deferred pass's sampler; False: REPEAT."""
        points = _c(points, np.float32).reshape(-1, 12)
        out = np.empty((points.shape[0], 4), np.float32)
        vp = (C.c_float * 3)(*[float(x) for x in view_pos])
        self._call("indirect_diffuse", _ptr(points), C.c_uint32(points.shape[0]), vp, C.c_int32(1 if clamp_to_edge else 0), _ptr(out))
        return out

    def set_row_stripe(self, offset, stride):
        """This handle updates only the probe rows z = offset + k*stride of every layer (multi-GPU sharding inside the layers)."""
        self._call("set_row_stripe", C.c_int32(offset), C.c_int32(stride))
        self.stripe = (int(offset), int(stride))

    def set_row_block(self, first, count):
        """This handle updates only the probe rows z in [first, first + count) of every layer."""
        self._call("set_row_block", C.c_int32(first), C.c_int32(count))
        self.stripe = (int(first), 1)
        self.stripe_rows = int(count)

    def download_atlas_stripe(self, out=None):
        """Atlas rows of this handle's stripe in all layers: uint16 [n_y, rows, 10, W, 4]."""
        off, stride = getattr(self, "stripe", (0, 1))
        rows = getattr(self, "stripe_rows", None) or (self.n_xz - off + stride - 1) // stride
        if out is None:
            out = np.empty((self.n_y, rows, 10, self.atlas_width, 4), np.uint16)
        self._call("download_atlas_stripe", _ptr(out), C.c_size_t(out.nbytes))
        return out

    def download_atlas_stripe_async(self, out):
        """The same stripe on the read-back stream (download_wait() blocks until it has landed); `out` page-locked."""
        self._call("download_atlas_stripe_async", _ptr(out), C.c_size_t(out.nbytes))

    def download_atlas_layers(self, layers, out=None):
        """Atlas rows of the given Y-layers, packed: uint16 [len(layers), 10*nXZ, W, 4]."""
        layers = _c(layers, np.int32).reshape(-1)
        if out is None:
            out = np.empty((layers.size, self.atlas_width, self.atlas_width, 4), np.uint16)
        self._call("download_atlas_layers", _ptr(layers), C.c_uint32(layers.size), _ptr(out), C.c_size_t(out.nbytes))
        return out

    def download_atlas_async(self, out):
        """Background read-back into `out` (uint16 [H, W, 4], ideally pinned); download_wait() before reading it."""
        self._call("download_atlas_async", _ptr(out), C.c_size_t(out.nbytes))

    def download_wait(self):
        self._call("download_wait")

    def download_atlas(self, out=None):
        """ProbeRaysConvTex as uint16 [H, W, 4] (RGBA16F bit patterns)."""
        if out is None:
            out = np.empty((self.atlas_height, self.atlas_width, 4), np.uint16)
        self._call("download_atlas", _ptr(out), C.c_size_t(out.nbytes))
        return out

    def upload_atlas(self, atlas):
        a = _c(atlas, np.uint16)
        self._call("upload_atlas", _ptr(a), C.c_size_t(a.nbytes))

    def download_rays(self, layer):
        """(ProbeRaysIrrTex, ProbeRaysDirTex) of `layer` as uint16 [S_y*nXZ, S_x*nXZ, 4]."""
        sx, sy = self.ray_tile
        irr = np.empty((sy * self.n_xz, sx * self.n_xz, 4), np.uint16)
        dr = np.empty_like(irr)
        self._call("download_rays", C.c_int32(layer), _ptr(irr), _ptr(dr), C.c_size_t(irr.nbytes))
        return irr, dr

    def upload_ray_radiance(self, layer, irr, rot):
        """Convolution alone (BASELINE config 1): ProbeRaysIrrTex of `layer` as uint16 [S_y*nXZ, S_x*nXZ, 4]; directions from `rot`."""
        irr = np.ascontiguousarray(irr)
        irr = irr.view(np.uint16) if irr.dtype == np.float16 else _c(irr, np.uint16)
        rot = _c(rot, np.float32).reshape(16)
        self._call("upload_ray_radiance", C.c_int32(layer), _ptr(irr), C.c_size_t(irr.nbytes), _ptr(rot))

    def download_ray_records(self, layer):
        out = np.empty(self.n_xz * self.n_xz * self.n_rays, L.RAY_RECORD)
        self._call("download_ray_records", C.c_int32(layer), _ptr(out), C.c_size_t(len(out)))
        return out.reshape(self.n_xz * self.n_xz, self.n_rays)

    def download_texel_records(self):
        out = np.empty((self.atlas_height, self.atlas_width, 4), np.float32)
        self._call("download_texel_records", _ptr(out), C.c_size_t(out.nbytes))
        return out

    def download_instances(self):
        """uint32 [num_instances, 4]: renderable, custom index (material | mesh << 10), first triangle, triangle count."""
        out = np.empty((self.bvh_info().num_instances, 4), np.uint32)
        self._call("download_instances", _ptr(out), C.c_size_t(len(out)))
        return out

    def download_bvh(self):
        """(nodes uint8[num_nodes*80], tris uint8[num_triangles*48]) — the BVH8 exactly as the traversal kernel reads it."""
        info = self.bvh_info()
        nodes = np.empty(info.num_nodes * info.node_bytes, np.uint8)
        tris = np.empty(info.num_triangles * info.tri_bytes, np.uint8)
        self._call("download_bvh", _ptr(nodes), C.c_size_t(nodes.nbytes), _ptr(tris), C.c_size_t(tris.nbytes))
        return nodes, tris

    def stats(self):
        s = Stats()
        self._call("get_stats", C.byref(s))
        return s.as_dict()

    def reset_stats(self):
        self._call("reset_stats")

    # ---- helpers -----------------------------------------------------------------------------------
    def make_rotation(self, axis, angle_deg):
        out = np.empty(16, np.float32)
        f = getattr(self._lib, self._pfx + "make_rotation")
        f.restype = C.c_int
        rc = f(C.c_int32(int(axis)), C.c_float(float(angle_deg)), _ptr(out))
        if rc != 0:
            raise DdgiError(rc, "make_rotation: bad axis")
        return out

    def rotations(self, axis_angle_list):
        return np.stack([self.make_rotation(a, d) for a, d in axis_angle_list])

    def fibonacci_dirs(self, n):
        out = np.empty((n, 3), np.float32)
        f = getattr(self._lib, self._pfx + "fibonacci_dirs")
        f.restype = C.c_int
        f(C.c_int32(n), _ptr(out))
        return out
