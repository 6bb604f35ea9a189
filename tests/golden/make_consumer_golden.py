#!/usr/bin/env python
"""Golden vectors for the atlas consumer from the reference's OWN compiled code: calls the function
calcIndirectDiffuseLight (pbr_light.glsl:365-388 -> sampleProbe -> weightProbe -> sampleSingleProbe) inside
assets/deferred_tiled_comp.spv — the deferred lighting pass — directly with tests/golden/spirv_interp.py for surface points in and
around the probe volume, the way deferred_tiled.comp:264-296 calls it: with the tile's baked atlas (a bindless texture, its own
REPEAT sampler) where the tile-info buffer has one, with the live atlas through the pass's CLAMP_TO_EDGE sampler
(DeferredLightingRenderPass.cpp:192-200) elsewhere.  Inputs and results go to tests/golden/consumer_golden.npz; the two atlases are
not stored: the live one is the oracle's after two updates of the test scene, the baked one is tests/consumer_points.py's
synthetic_baked_atlas — both are re-made by the test.

    python tests/golden/make_consumer_golden.py [/root/reference/assets/deferred_tiled_comp.spv]
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from spirv_interp import Invocation, Module, Ptr, f32  # noqa: E402
from make_rchit_golden import Sampler  # noqa: E402

GRID = (64, 8, 64)
POINTS = 160


def setup(ob):
    """Scene, oracle handle after two updates, tile pattern, points: shared with tests/test_golden_consumer.py through the fixture's fields."""
    from anerend_b200 import layouts as L, make_params, scenes
    from tests.consumer_points import surface_points, synthetic_baked_atlas
    nxz, ny, nr = GRID
    scene = copy.copy(scenes.textured_soup(51, grid=GRID))
    baked = synthetic_baked_atlas(10 * nxz, 10 * nxz * ny, 5)
    scene.textures = list(scene.textures) + [(baked.view(np.uint8).reshape(baked.shape[0], baked.shape[1], 8), L.TEX_RGBA16F, False)]
    o = ob.Oracle(make_params(nxz, ny, nr))
    if hasattr(o, "set_threads"):
        o.set_threads(0)
    o.upload_scene(scene); o.set_scene_ubo(scene.ubo); o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(3, ny))
    o.update(-1, rots); o.update(-1, rots)
    tiles = np.array([len(scene.textures) - 1 if (i % 21 + i // 21) % 2 == 0 else -1 for i in range(441)], np.int32)
    o.upload_tile_infos(tiles)
    cam = scene.ubo["camera_pos"][0][:3]
    pts = surface_points(POINTS, nxz, ny, seed=33, camera=tuple(float(c) for c in cam))
    return scene, o, tiles, pts, baked


def main():
    from oracle import binding as ob
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/deferred_tiled_comp.spv"
    m = Module(path)
    fid = next(f for f in m.functions if m.names.get(f, "").startswith("calcIndirectDiffuseLight("))
    tile_fid = next(f for f in m.functions if m.names.get(f, "").startswith("worldPosToTileInfo("))
    scene, o, tiles, pts, baked = setup(ob)
    live = Sampler(o.download_atlas().view(np.float16).astype(np.float32), clamp=True)          # the pass's CLAMP_TO_EDGE sampler
    baked_s = Sampler(baked.view(np.float16).astype(np.float32), clamp=False)                    # a bindless texture: REPEAT
    ubo_gid = next(g for g, n in m.names.items() if n == "ubo" and g in m.globals)
    ubo_t = m.types[m.globals[ubo_gid][0]][2]
    ubo = m.zero(ubo_t)
    u = scene.ubo[0]
    for name, val in (("cameraPos", [f32(x) for x in u["camera_pos"]]), ("bakeTileInfo", [int(x) for x in u["bake_tile_info"]]), ("flags", int(u["flags"]))):
        ubo[m.struct_member(ubo_t, name)] = val
    P = lambda v: Ptr([v], (0,))
    v3 = lambda a: [f32(x) for x in a]
    out, on_tile = np.zeros((len(pts), 3), np.float32), np.zeros(len(pts), np.int64)
    want = o.indirect_diffuse(pts, u["camera_pos"][:3], clamp_to_edge=True)
    for k, q in enumerate(pts):
        inv = Invocation(m, {"ubo": ubo, "tileInfoBuffer": [[[int(t)] for t in tiles]], "gl_GlobalInvocationID": [0, 0, 0]})
        # deferred_tiled.comp:264-268: which texture the caller passes
        idx = P([0, 0])
        tile_buf_idx = inv.call(tile_fid, [P(v3(q[0:3])), P(v3(u["camera_pos"][:3])), idx])
        atlas_idx = tiles[tile_buf_idx] if tile_buf_idx >= 0 else -1
        on_tile[k] = int(tile_buf_idx >= 0 and atlas_idx >= 0)
        res = inv.call(fid, [P(v3(q[3:6])), P(v3(q[6:9])), P(f32(q[9])), P(f32(q[10])), P(v3(q[0:3])), P(v3(u["camera_pos"][:3])), P(baked_s if on_tile[k] else live)])
        out[k] = [float(c) for c in res]
        if k % 32 == 0:
            print(k, "tile" if on_tile[k] else "live", "binary:", out[k], "oracle:", want[k, :3])
    o.close()
    print("points on a baked tile:", int(on_tile.sum()), "of", len(pts))
    dst = os.path.join(HERE, "consumer_golden.npz")
    np.savez_compressed(dst, grid=np.array(GRID), points=pts, tiles=tiles, on_tile=on_tile, out=out)
    print("wrote", dst)


if __name__ == "__main__":
    main()
