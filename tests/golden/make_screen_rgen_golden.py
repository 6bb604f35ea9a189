#!/usr/bin/env python
"""Golden vectors for the two screen-space ray generators that share the path's acceleration structure, from the
reference's OWN compiled shaders: executes assets/shadow_rgen.spv + shadow_rmiss.spv ("ShadowRT") and
assets/mirror_reflections_rgen.spv + mirror_reflections_rmiss.spv ("SpecularGIRT") with tests/golden/spirv_interp.py over a small
screen (depth buffer with sky pixels, G-buffer normals, a projective invViewProj with w != 1) and writes the rays they hand to
traceRayEXT and the pixels they store to tests/golden/screen_rgen_golden.npz.  The trace call is a callback: it records its
arguments and then either leaves the payload alone (a hit whose closest-hit shader is skipped / a marker hit) or runs the
reference's compiled MISS shader on it.

    python tests/golden/make_screen_rgen_golden.py [/root/reference/assets]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from spirv_interp import Invocation, Module, f32  # noqa: E402

W, H = 12, 8
UBO_RT_SHADOWS_FLAG, UBO_SPECULAR_GI_FLAG, UBO_RT_ON_FLAG = 1 << 3, 1 << 6, 1 << 10


class Texel:
    """sampler2D sampled at texel centres (pixelToUv): returns the texel the coordinate falls in."""

    def __init__(self, a):
        self.a = a if a.ndim == 3 else a[..., None]

    def size(self, lod):
        return [self.a.shape[1], self.a.shape[0]]

    def sample_lod(self, uv, lod):
        h, w = self.a.shape[:2]
        x, y = int(np.floor(f32(uv[0] * f32(w)))), int(np.floor(f32(uv[1] * f32(h))))
        t = [f32(c) for c in self.a[y, x]]
        return (t + [f32(0)] * 4)[:4]


class Image:
    def __init__(self):
        self.px = {}

    def write(self, xy, v):
        self.px[tuple(xy)] = [float(c) for c in v]


def fill_ubo(m, gid_name, ubo_np):
    gid = next(g for g, n in m.names.items() if n == gid_name and g in m.globals)
    t = m.types[m.globals[gid][0]][2]
    ubo = m.zero(t)
    u = ubo_np[0]
    mat = lambda a: [[f32(a[4 * c + r]) for r in range(4)] for c in range(4)]
    for name, val in (("invViewProj", mat(u["inv_view_proj"])), ("cameraPos", [f32(x) for x in u["camera_pos"]]), ("lightDir", [f32(x) for x in u["light_dir"]]),
                      ("skyColor", [f32(x) for x in u["sky_color"]]), ("skyIntensity", f32(u["sky_intensity"])),
                      ("screenWidth", int(u["screen_width"])), ("screenHeight", int(u["screen_height"])), ("flags", int(u["flags"]))):
        ubo[m.struct_member(t, name)] = val
    return ubo


def run_miss(mm, ubo_np, payload_name, payload):
    bind = {payload_name: payload}
    if any(n == "ubo" and g in mm.globals for g, n in mm.names.items()):
        bind["ubo"] = fill_ubo(mm, "ubo", ubo_np)
    inv = Invocation(mm, bind)
    inv.run()
    gid = next(g for g, n in mm.names.items() if n == payload_name and g in mm.globals)
    return inv.g[gid].load()


def main():
    import shadow_inputs as SI
    from anerend_b200 import scenes
    assets = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets"
    ivp = SI.perspective_inv_view_proj(5)
    depth = SI.depth_buffer(W, H, 21, sky_fraction=0.15)
    geom0 = SI.gbuffer_normals(W, H, 22)
    base = scenes.make_ubo((1.5, 6.0, -2.25), sky_intensity=0.8, sky_color=(0.3, 0.6, 0.9), light_dir=(0.5, -0.8, 0.3))
    out = {"inv_view_proj": ivp, "depth": depth, "geom0": geom0}

    # ---- ShadowRT --------------------------------------------------------------------------------------
    ubo_np = SI.with_shadow_camera(base, ivp, W, H)
    m, mm = Module(os.path.join(assets, "shadow_rgen.spv")), Module(os.path.join(assets, "shadow_rmiss.spv"))
    ubo = fill_ubo(m, "ubo", ubo_np)
    rays, image = np.full((H, W, 8), np.nan, np.float64), np.full((H, W, 4), np.nan, np.float64)
    for y in range(H):
        for x in range(W):
            calls, im = [], Image()
            occluded = (x + y) % 3 == 0

            def trace(org, d, tmin, tmax, payload, _c=calls, _occ=occluded):
                _c.append([float(c) for c in org] + [float(c) for c in d] + [float(tmin), float(tmax)])
                if not _occ:                                   # terminate-on-first-hit + skip-closest-hit: only a miss touches the payload
                    payload.store(run_miss(mm, ubo_np, "prd", payload.load()))

            Invocation(m, {"gl_LaunchIDEXT": [x, y, 0], "ubo": ubo, "depth": Texel(depth), "outputIm": im, "prd": [f32(7)], "traceRay": trace}).run()
            assert list(im.px) == [(x, y)]
            image[y, x] = im.px[(x, y)]
            if depth[y, x] == 1.0:
                assert not calls and image[y, x, 0] == 1.0
            else:
                assert len(calls) == 1 and image[y, x, 0] == (0.0 if occluded else 1.0)
                rays[y, x] = calls[0]
    # the flag gate (shadow.rgen:32-34): no trace, no store
    off = SI.with_shadow_camera(base, ivp, W, H, rt_shadows=False)
    im = Image()
    Invocation(m, {"gl_LaunchIDEXT": [0, 0, 0], "ubo": fill_ubo(m, "ubo", off), "depth": Texel(depth), "outputIm": im, "prd": [f32(7)],
                   "traceRay": lambda *a: (_ for _ in ()).throw(AssertionError("traced with the flag off"))}).run()
    assert not im.px
    out.update(shadow_rays=rays, shadow_image=image, ubo_shadow=ubo_np.view(np.uint8))
    print("ShadowRT: rays", int(np.isfinite(rays[..., 0]).sum()), "sky pixels", int((depth == 1.0).sum()))

    # ---- SpecularGIRT ----------------------------------------------------------------------------------
    ubo_np = SI.with_specular_camera(base, ivp, W, H)
    m, mm = Module(os.path.join(assets, "mirror_reflections_rgen.spv")), Module(os.path.join(assets, "mirror_reflections_rmiss.spv"))
    ubo = fill_ubo(m, "ubo", ubo_np)
    rays, image = np.full((H, W, 8), np.nan, np.float64), np.full((H, W, 4), np.nan, np.float64)
    geom1 = np.zeros((H, W, 4), np.float32)
    marker = [[f32(0.5), f32(2.0), f32(0.125)], f32(3.5), f32(1.0)]
    for y in range(H):
        for x in range(W):
            calls, im = [], Image()
            hit = (x + y) % 3 == 0

            def trace(org, d, tmin, tmax, payload, _c=calls, _hit=hit):
                _c.append([float(c) for c in org] + [float(c) for c in d] + [float(tmin), float(tmax)])
                payload.store([list(marker[0]), marker[1], marker[2]] if _hit else run_miss(mm, ubo_np, "payload", payload.load()))

            Invocation(m, {"gl_LaunchIDEXT": [x, y, 0], "ubo": ubo, "depthIm": Texel(depth), "geom0Im": Texel(geom0), "geom1Im": Texel(geom1),
                           "outputIm": im, "payload": [[f32(9)] * 3, f32(9), f32(9)], "traceRay": trace}).run()
            assert list(im.px) == [(x, y)]
            image[y, x] = im.px[(x, y)]
            if depth[y, x] == 1.0:
                assert not calls and image[y, x].tolist() == [0.0] * 4
            else:
                assert len(calls) == 1
                rays[y, x] = calls[0]
                if hit:
                    assert image[y, x].tolist() == [0.5, 2.0, 0.125, 3.5]
    out.update(specular_rays=rays, specular_image=image, ubo_specular=ubo_np.view(np.uint8))
    print("SpecularGIRT: rays", int(np.isfinite(rays[..., 0]).sum()), "a miss pixel:", image[0, 1] if depth[0, 1] != 1.0 else image[1, 1])
    dst = os.path.join(HERE, "screen_rgen_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst)


if __name__ == "__main__":
    main()
