#!/usr/bin/env python
"""Golden vectors for the probe hit shader from the reference's OWN compiled shader: executes
assets/irradiance_probe_update_rchit.spv (main + getSurfaceDataFromMat + calcLight + calcIndirectDiffuseLight + sampleProbe + ...)
with tests/golden/spirv_interp.py for a sample of the hits of a textured test scene, and writes the hits and the payload the binary
produces to tests/golden/rchit_golden.npz.  The same hits are also run through assets/mirror_reflections_rchit.spv (the "SpecularGIRT"
hit shader: mat3(gl_ObjectToWorldEXT) normal transform, distance payload, UBO_DDGI_FLAG gate) against the atlas after the second update.  What the driver / hardware would supply is supplied here from the oracle's run of the
same scene: the hit (instance custom index, primitive id, barycentrics, object<->world matrices, ray origin), the result of the shadow
ray (OpTraceRayKHR callback), and texture / atlas sampling (bilinear at LOD 0 as DESIGN.md defines it).  Everything else — vertex
fetch and interpolation, normal transform, toLinear, surface data from the material and its textures, Cook-Torrance, the probe-grid
lookup with its eight weighted taps, emissive — is the reference's compiled code.

    python tests/golden/make_rchit_golden.py [/root/reference/assets/irradiance_probe_update_rchit.spv]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from spirv_interp import Invocation, Module, f32  # noqa: E402

GRID = (64, 8, 64)                 # the grid the shipped binary is compiled for (probe_helpers.glsl:1-2); 64 rays keep the oracle run short
SAMPLES = 96


def invert3x3_f32(m):
    """oracle/ddgi_oracle.cpp invert3x3 (adjugate / determinant, fp32, same association); m = column-major mat4, returns w2o[c*3+r]."""
    a, b, c, d, e, f, g, h, i = (f32(m[k]) for k in (0, 4, 8, 1, 5, 9, 2, 6, 10))
    A, B, C = f32(f32(e * i) - f32(f * h)), f32(-f32(f32(d * i) - f32(f * g))), f32(f32(d * h) - f32(e * g))
    det = f32(f32(f32(a * A) + f32(b * B)) + f32(c * C))
    inv = f32(f32(1.0) / det)
    r00, r01, r02 = f32(A * inv), f32(f32(-f32(f32(b * i) - f32(c * h))) * inv), f32(f32(f32(b * f) - f32(c * e)) * inv)
    r10, r11, r12 = f32(B * inv), f32(f32(f32(a * i) - f32(c * g)) * inv), f32(f32(-f32(f32(a * f) - f32(c * d))) * inv)
    r20, r21, r22 = f32(C * inv), f32(f32(-f32(f32(a * h) - f32(b * g))) * inv), f32(f32(f32(a * e) - f32(b * d)) * inv)
    return [r00, r10, r20, r01, r11, r21, r02, r12, r22]


class Sampler:
    """sampler2D over an [h, w, 4] float32 array: bilinear at LOD 0, REPEAT or CLAMP_TO_EDGE (defined arithmetic)."""

    def __init__(self, texels, clamp=False):
        self.t, self.clamp = texels, clamp

    def size(self, lod):
        return [self.t.shape[1], self.t.shape[0]]

    def _wrap(self, i, n):
        return min(max(i, 0), n - 1) if self.clamp else i % n

    def sample_lod(self, uv, lod):
        h, w = self.t.shape[:2]
        x, y = f32(f32(uv[0] * f32(w)) - f32(0.5)), f32(f32(uv[1] * f32(h)) - f32(0.5))
        fx0, fy0 = f32(np.floor(x)), f32(np.floor(y))
        fx, fy, ix, iy = f32(x - fx0), f32(y - fy0), int(fx0), int(fy0)
        t00, t10 = self.t[self._wrap(iy, h), self._wrap(ix, w)], self.t[self._wrap(iy, h), self._wrap(ix + 1, w)]
        t01, t11 = self.t[self._wrap(iy + 1, h), self._wrap(ix, w)], self.t[self._wrap(iy + 1, h), self._wrap(ix + 1, w)]
        w00, w10, w01, w11 = f32(f32(1 - fx) * f32(1 - fy)), f32(fx * f32(1 - fy)), f32(f32(1 - fx) * fy), f32(fx * fy)
        return [f32(f32(f32(f32(t00[k] * w00) + f32(t10[k] * w10)) + f32(t01[k] * w01)) + f32(t11[k] * w11)) for k in range(4)]


def _copy_nested(x):
    return [_copy_nested(c) for c in x] if isinstance(x, list) else x


def decode_texture(px, fmt, lut):
    from anerend_b200 import layouts as L
    h, w = px.shape[:2]
    out = np.zeros((h, w, 4), np.float32); out[..., 3] = 1.0
    if fmt == L.TEX_RGBA8_SRGB:
        out[..., :3] = lut[px[..., :3]]; out[..., 3] = lut[256 + px[..., 3].astype(int)]
    else:
        c = px.shape[2]
        out[..., :c] = lut[256 + px.astype(int)]
    return out


def main():
    from anerend_b200 import layouts as L, make_params, scenes
    from oracle import binding as ob
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/irradiance_probe_update_rchit.spv"
    m = Module(path)
    nxz, ny, nr = GRID
    scene = scenes.textured_soup(51, grid=GRID)
    o = ob.Oracle(make_params(nxz, ny, nr, flags=L.FLAG_RAY_RECORDS))          # the faithful subset: sun only, every sun shadow ray traced
    if os.environ.get("RCHIT_FLAGS_OFF"):                      # debugging aid: clear UBO flag bits in both runs
        scene.ubo["flags"] &= ~np.uint32(int(os.environ["RCHIT_FLAGS_OFF"], 0))
    o.upload_scene(scene); o.set_scene_ubo(scene.ubo); o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(3, ny))
    o.update(-1, rots)
    prev = o.download_atlas().view(np.float16).astype(np.float32)             # the atlas the second update's hit shader samples
    o.update(-1, rots)
    recs = [o.download_ray_records(l) for l in range(ny)]
    atlas2 = o.download_atlas().view(np.float16).astype(np.float32)           # the atlas mirror_reflections.rchit samples below
    mm = Module(os.path.join(os.path.dirname(path), "mirror_reflections_rchit.spv"))
    mubo_gid = next(g for g, n in mm.names.items() if n == "ubo" and g in mm.globals)
    mubo_t = mm.types[mm.globals[mubo_gid][0]][2]
    msp_gid = next(g for g, n in mm.names.items() if n == "shadowPayload" and g in mm.globals)
    mshadow_member = mm.struct_member(mm.types[mm.globals[msp_gid][0]][2], "shadow")
    mpay_gid = next(g for g, n in mm.names.items() if n == "payload" and g in mm.globals)
    # instances in the order tlas_update.comp emits them (visible, not culled), with their triangle ranges
    inst, first = [], 0
    for r in scene.renderables:
        if not (r["visible"] & 1) or (r["visible"] >> 1) & 1:
            continue
        for i in range(r["num_meshes"]):
            mesh = int(scene.model_indices[r["model_offset"] + i]); mat = int(scene.material_indices[r["first_material_index"] + i])
            ntri = int(scene.mesh_num_indices[mesh]) // 3
            inst.append((first, ntri, mesh, mat, r["transform"].astype(np.float32)))
            first += ntri
    lut = np.array([ob.to_linear((c / np.float32(255.0),) * 3)[0] for c in range(256)] + [np.float32(c) / np.float32(255.0) for c in range(256)], np.float32)
    textures = [Sampler(decode_texture(px, fmt, lut), clamp) for (px, fmt, clamp) in scene.textures]
    verts = scene.vertices.view(np.float32).reshape(-1, 24)
    vint = scene.vertices.view(np.int32).reshape(-1, 24)
    vertices = [[[f32(x) for x in v[0:4]], [f32(x) for x in v[4:8]], [f32(x) for x in v[8:12]], [f32(x) for x in v[12:16]], [f32(x) for x in v[16:20]],
                 [int(x) for x in vi[20:24]]] for v, vi in zip(verts, vint)]
    meshes = [[[f32(x) for x in mm["min_pos"]], [f32(x) for x in mm["max_pos"]], int(mm["vertex_offset"]), int(mm["index_offset"]), 0] for mm in scene.meshes]
    mats = [[[f32(x) for x in mt["base_col_fac"]], [f32(x) for x in mt["emissive"]], [int(x) for x in mt["bindless_indices"]], [f32(x) for x in mt["met_rough"]]]
            for mt in scene.materials]
    ubo_gid = next(g for g, n in m.names.items() if n == "ubo" and g in m.globals)
    ubo_t = m.types[m.globals[ubo_gid][0]][2]
    ubo, mubo = m.zero(ubo_t), mm.zero(mubo_t)
    u = scene.ubo[0]
    for name, val in (("cameraPos", [f32(x) for x in u["camera_pos"]]), ("lightDir", [f32(x) for x in u["light_dir"]]),
                      ("skyColor", [f32(x) for x in u["sky_color"]]), ("bakeTileInfo", [int(x) for x in u["bake_tile_info"]]),
                      ("sunIntensity", f32(u["sun_intensity"])), ("skyIntensity", f32(u["sky_intensity"])), ("flags", int(u["flags"]))):
        ubo[m.struct_member(ubo_t, name)] = val
        mubo[mm.struct_member(mubo_t, name)] = val
    step, cam = (1.0, 2.0, 1.0), u["camera_pos"]
    camoff = (f32(int(cam[0] / 1.0)) * f32(1.0) - f32(1.0) * f32(nxz) / f32(2.0), f32(0.0), f32(int(cam[2] / 1.0)) * f32(1.0) - f32(1.0) * f32(nxz) / f32(2.0))
    sp_gid = next(g for g, n in m.names.items() if n == "shadowPayload" and g in m.globals)
    shadow_member = m.struct_member(m.types[m.globals[sp_gid][0]][2], "shadow")
    rng = np.random.RandomState(1)
    hits = [(l, p, r) for l in range(ny) for p in range(nxz * nxz) for r in range(nr) if recs[l][p, r]["prim"] != 0xFFFFFFFF]
    pick = [hits[k] for k in rng.choice(len(hits), SAMPLES, replace=False)]
    mirror = np.zeros((SAMPLES, 4), np.float32); origins = np.zeros((SAMPLES, 3), np.float32)
    # baked tiles (quirk Q5 / SURVEY 8f N4): a checkerboard of the 21 x 21 paged tiles has a baked atlas (bindless index 0), so hits on those
    # tiles place the probe grid on the 32 m tile (pbr_light.glsl:251-273) while the hit shader keeps sampling the live atlas
    tile_pattern = np.array([0 if (i % 21 + i // 21) % 2 == 0 else -1 for i in range(441)], np.int32)
    o.upload_tile_infos(tile_pattern)
    tiled = np.zeros((SAMPLES, 3), np.float32)
    # bake mode: UBO_BAKE_MODE_ON_FLAG places the grid on tile bakeTileInfo.xy of size bakeTileInfo.z (pbr_light.glsl:274-291)
    bake_tile = (-1, -1, 48, 0)
    baked = np.zeros((SAMPLES, 3), np.float32)
    bubo = _copy_nested(ubo)
    bubo[m.struct_member(ubo_t, "flags")] = int(u["flags"]) | (1 << 11)
    bubo[m.struct_member(ubo_t, "bakeTileInfo")] = [int(x) for x in bake_tile]
    out = np.zeros((SAMPLES, 3), np.float32); meta = np.zeros((SAMPLES, 3), np.int64); prim_uv = np.zeros((SAMPLES, 3), np.float64); lit = np.zeros(SAMPLES, np.int64)
    for k, (layer, probe, ray) in enumerate(pick):
        rec = recs[layer][probe, ray]
        prim = int(rec["prim"])
        first, ntri, mesh, mat, T = next(i for i in inst if i[0] <= prim < i[0] + i[1])
        origin = [f32(f32(f32(probe % nxz) * f32(step[0])) + camoff[0]), f32(f32(f32(layer) * f32(step[1])) + camoff[1]), f32(f32(f32(probe // nxz) * f32(step[2])) + camoff[2])]
        w2o = invert3x3_f32(T)
        o2w_cols = [[f32(T[4 * c + r]) for r in range(3)] for c in range(4)]
        w2o_cols = [[w2o[3 * c + r] for r in range(3)] for c in range(3)] + [[f32(0), f32(0), f32(0)]]
        shadow_calls = []

        def trace(org, d, tmin, tmax, payload, _rec=rec, _calls=shadow_calls):
            _calls.append((org, d, tmin, tmax))
            s = payload.load()
            # the miss shader clears it (rmiss:19); a hit leaves the 1.0 of rchit:46
            s[shadow_member] = f32(0.0) if (int(_rec["lit_mask"]) & 1) else f32(1.0)
            payload.store(s)

        payload = [[f32(0)] * 3, [f32(0)] * 3, f32(0)]
        bind = {"gl_InstanceCustomIndexEXT": (mat & 0x3FF) | ((mesh & 0x3FFF) << 10), "gl_PrimitiveID": prim - first,
                "bary": [f32(rec["u"]), f32(rec["v"])], "gl_ObjectToWorldEXT": o2w_cols, "gl_WorldToObjectEXT": w2o_cols,
                "gl_WorldRayOriginEXT": origin, "meshBuffer": [meshes], "indexBuffer": [[int(x) for x in scene.indices]],
                "vertexBuffer": [vertices], "materialBuffer": [mats], "ubo": ubo, "tileInfoBuffer": [[[-1]] * 441],
                "probeTex": Sampler(prev), "textures": textures, "payload": payload, "shadowPayload": [[f32(0)] * 3, [f32(0)] * 3, f32(0)],
                "traceRay": trace}
        inv = Invocation(m, bind)
        inv.run()
        pay_gid = next(g for g, n in m.names.items() if n == "payload" and g in m.globals)
        res = inv.g[pay_gid].load()
        out[k] = res[0]
        assert len(shadow_calls) == 1 and float(shadow_calls[0][2]) == np.float32(0.1) and float(shadow_calls[0][3]) == 50.0
        # the same hit through mirror_reflections.rchit ("SpecularGIRT"): payload (irradiance, dist, shadow), the live atlas of now
        mcalls = []

        def mtrace(org, d, tmin, tmax, payload, _rec=rec, _calls=mcalls):
            _calls.append((org, d, tmin, tmax))
            sp = payload.load()
            sp[mshadow_member] = f32(0.0) if (int(_rec["lit_mask"]) & 1) else f32(1.0)
            payload.store(sp)

        mbind = dict(bind)
        mbind.update({"ubo": mubo, "probeTex": Sampler(atlas2), "payload": [[f32(0)] * 3, f32(0), f32(0)], "shadowPayload": [[f32(0)] * 3, f32(0), f32(0)],
                      "traceRay": mtrace})
        minv = Invocation(mm, mbind)
        minv.run()
        mres = minv.g[mpay_gid].load()
        mirror[k] = [*mres[0], mres[1]]; origins[k] = origin
        # the probe hit shader again, with the tile-info buffer filled
        tbind = dict(bind)
        tbind.update({"tileInfoBuffer": [[[int(t)] for t in tile_pattern]], "probeTex": Sampler(atlas2), "payload": [[f32(0)] * 3, [f32(0)] * 3, f32(0)],
                      "shadowPayload": [[f32(0)] * 3, [f32(0)] * 3, f32(0)], "traceRay": trace})
        tinv = Invocation(m, tbind)
        tinv.run()
        tiled[k] = tinv.g[pay_gid].load()[0]
        bbind = dict(bind)
        bbind.update({"ubo": bubo, "probeTex": Sampler(atlas2), "payload": [[f32(0)] * 3, [f32(0)] * 3, f32(0)],
                      "shadowPayload": [[f32(0)] * 3, [f32(0)] * 3, f32(0)], "traceRay": trace})
        binv = Invocation(m, bbind)
        binv.run()
        baked[k] = binv.g[pay_gid].load()[0]
        assert len(mcalls) == 1 and [float(c) for c in mcalls[0][0]] == [float(c) for c in shadow_calls[0][0]]      # same shadow ray as the probe shader
        meta[k] = (layer, probe, ray); prim_uv[k] = (prim, rec["u"], rec["v"]); lit[k] = int(rec["lit_mask"]) & 1
        if k % 16 == 0 or os.environ.get("RCHIT_VERBOSE"):
            print(k, "binary:", out[k], "oracle:", rec["radiance"], "lit", lit[k], "material", mat)
            print("   tiled binary:", tiled[k], "oracle:", o.shade_hit(prim, rec["u"], rec["v"], origin)[0])
            print("   mirror binary:", mirror[k])
    dst = os.path.join(HERE, "rchit_golden.npz")
    np.savez_compressed(dst, rays=meta, prim_uv=prim_uv, lit=lit, irradiance=out, grid=np.array(GRID), origins=origins, mirror=mirror, tiled=tiled,
                        tile_pattern=tile_pattern, baked=baked, bake_tile=np.array(bake_tile))
    print("wrote", dst)


if __name__ == "__main__":
    main()
