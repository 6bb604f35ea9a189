"""Seeded fuzz of the product's kernel logic (host harness) against the oracle: random small scenes (triangle soups, textured soups,
Cornell boxes), random grids (1-5 probes a side, 1-3 layers), ray counts (32-256), cameras, light counts, flag combinations — two
updates each with translation, hit ids / barycentrics / shadow masks and the stored atlas bit for bit; then random G-buffer points
through both consumer samplers and random camera walks through the translation pass.  (A 400-scene run of the same loop at the end of
round 1 found no mismatch; this keeps a slice of it in the suite.)"""
import numpy as np

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

ALL = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_TEXEL_RECORDS | L.FLAG_COUNTERS


def test_random_scenes_bit_exact(oracle, make_emu):
    for seed in range(2000, 2048):
        rng = np.random.RandomState(seed)
        nxz, ny, nr = int(rng.choice([1, 2, 3, 5])), int(rng.choice([1, 2, 3])), int(rng.choice([32, 64, 128, 256]))
        cam = tuple(rng.uniform(-3, 3, 3) * [1, 0.5, 1] + [0, 1.5, 0])
        kind = rng.randint(0, 3)
        if kind == 0:
            s = scenes.random_soup(seed, ntris=int(rng.choice([40, 300, 1500])), extent=float(rng.uniform(2, 12)), grid=(nxz, ny, nr), camera=cam,
                                   nlights=int(rng.randint(0, 4)))
        elif kind == 1:
            s = scenes.textured_soup(seed, ntris=int(rng.choice([60, 400])), grid=(nxz, ny, nr), camera=cam)
        else:
            s = scenes.cornell_box(seed=seed, grid=(nxz, ny, nr), wall_res=int(rng.choice([2, 6])), box_res=2)
            s.ubo = s.with_camera(cam)
        if rng.rand() < 0.3:
            s.ubo["flags"] &= ~np.uint32(L.UBO_DDGI_MULTI_FLAG)
        p = make_params(nxz, ny, nr, flags=ALL | (L.FLAG_SKIP_DEAD_SHADOW_RAYS if rng.rand() < 0.5 else 0))
        o, e = oracle.Oracle(p), make_emu(p)
        rots = parity.setup_pair(s, (o, e), seed=seed)
        for it in range(2):
            for h in (o, e):
                h.translate()
                h.update(-1, rots)
            for layer in range(ny):
                parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
            assert np.array_equal(o.download_atlas(), e.download_atlas()), seed
        o.close(); e.close()


def test_random_consumer_points_and_camera_walks_bit_exact(oracle, make_emu):
    for seed in range(40):
        rng = np.random.RandomState(seed)
        nxz, ny = int(rng.choice([1, 2, 4, 7])), int(rng.choice([1, 2, 5]))
        p = make_params(nxz, ny, 64)
        o, e = oracle.Oracle(p), make_emu(p)
        cam = tuple(rng.uniform(-40, 40, 3))
        s = scenes.empty_scene(grid=(nxz, ny, 64), camera=cam)
        if rng.rand() < 0.3:
            s.ubo["flags"] |= L.UBO_BAKE_MODE_ON_FLAG
            s.ubo["bake_tile_info"][0] = (int(rng.randint(-3, 3)), int(rng.randint(-3, 3)), int(rng.choice([16, 32, 48])), 0)
        atlas = rng.uniform(0, 3, (10 * nxz * ny, 10 * nxz, 4)).astype(np.float16).view(np.uint16)
        pts = np.zeros((200, 12), np.float32)
        pts[:, 0:3] = np.array(cam) + rng.uniform(-1, 1, (200, 3)) * rng.choice([1.0, 5.0, 60.0])
        nrm = rng.normal(size=(200, 3)); pts[:, 3:6] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        pts[:, 6:9] = rng.uniform(0, 1, (200, 3)); pts[:, 9] = rng.rand(200) < 0.3; pts[:, 10] = rng.uniform(0.05, 1, 200)
        for h in (o, e):
            h.set_scene_ubo(s.ubo); h.upload_atlas(atlas)
        for clamp in (False, True):
            assert np.array_equal(o.indirect_diffuse(pts, cam, clamp).view(np.uint32), e.indirect_diffuse(pts, cam, clamp).view(np.uint32)), (seed, clamp)
        for step in range(4):
            u = s.with_camera(tuple(np.array(cam) + rng.uniform(-2.5, 2.5, 3)))
            for h in (o, e):
                h.set_scene_ubo(u); h.translate()
            assert np.array_equal(o.download_atlas(), e.download_atlas()), (seed, step)
        o.close(); e.close()


def test_random_refits_and_vertex_updates_bit_exact(oracle, make_emu):
    """Moving nodes (random rigid steps, now and then mirrored) and deforming vertex ranges, refit after each step."""
    for seed in range(3000, 3016):
        rng = np.random.RandomState(seed)
        nxz, ny, nr = int(rng.choice([2, 3])), int(rng.choice([1, 2])), 64
        s = scenes.random_soup(seed, ntris=int(rng.choice([100, 600])), grid=(nxz, ny, nr))
        p = make_params(nxz, ny, nr, flags=ALL)
        o, e = oracle.Oracle(p), make_emu(p)
        rots = parity.setup_pair(s, (o, e), seed=seed)
        verts = s.vertices.copy()
        for step in range(3):
            rend = s.renderables.copy()
            for i in range(len(rend)):
                if rng.rand() < 0.6:
                    M = rend["transform"][i].reshape(4, 4).T.astype(np.float64)
                    M2 = scenes.translate(rng.uniform(-1, 1, 3)) @ scenes.rotate(int(rng.randint(0, 3)), rng.uniform(0, 360)) @ M
                    if rng.rand() < 0.2:
                        M2 = M2 @ scenes.scale((-1, 1, 1))
                    rend["transform"][i] = M2.T.astype(np.float32).ravel()
            k0 = int(rng.randint(0, len(verts) - 10)); k1 = min(len(verts), k0 + int(rng.randint(1, 200)))
            verts["pos"][k0:k1] += rng.uniform(-0.05, 0.05, (k1 - k0, 3)).astype(np.float32)
            for h in (o, e):
                h.update_vertices(verts[k0:k1], first=k0)
                h.refit_accel(rend)
                h.update(-1, rots)
            for layer in range(ny):
                parity.compare_records(o.download_ray_records(layer), e.download_ray_records(layer))
            assert np.array_equal(o.download_atlas(), e.download_atlas()), (seed, step)
        o.close(); e.close()
