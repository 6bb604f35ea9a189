#!/usr/bin/env python
"""Planning probe for the next round: how much work would a light-space culling structure leave for the sun's
shadow rays?  CPU only, uses the oracle (test infrastructure) on a BASELINE scene:

  * one update with ray records -> the hit points the sun shadow rays start from;
  * all world triangles projected into the sun's orthographic space (footprint box + depth interval);
  * a G x G uniform grid over the footprint with per-cell triangle lists (conservative: bounding boxes);
  * per shadow ray: triangles in its cell, and those whose depth interval lies on the sun side of the hit point beyond tMin —
    the exact Moeller-Trumbore tests a grid walk would still have to do (no early out counted) — to be read next to what the
    product's BVH8 walk does today (bench line: roofline.nodes_per_shadow_ray, tris_per_shadow_ray).

    python tests/lightspace_probe.py [c2|c3] [grid=1024]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from anerend_b200 import layouts as L, make_params, scenes
    from oracle import binding as ob
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    scene, grid, _ = bench.build_workload(name)
    nxz, ny, nr = grid
    o = ob.Oracle(make_params(nxz, ny, nr, flags=L.FLAG_RAY_RECORDS | L.FLAG_COUNTERS | L.FLAG_SKIP_DEAD_SHADOW_RAYS))   # sun only: the faithful subset
    o.set_threads(0)
    o.upload_scene(scene); o.set_scene_ubo(scene.ubo); o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(scene.seed, ny))
    o.update(-1, rots)
    st = o.stats()
    tris = o.world_triangles().astype(np.float64)                      # [n, 3, 3]
    recs = np.concatenate([o.download_ray_records(l).reshape(-1) for l in range(ny)])
    o.close()
    hit = recs[recs["prim"] != L.MISS]
    t = tris[hit["prim"]]
    pos = t[:, 0] + hit["u"][:, None] * (t[:, 1] - t[:, 0]) + hit["v"][:, None] * (t[:, 2] - t[:, 0])
    ld = scene.ubo["light_dir"][0][:3].astype(np.float64)
    sun = -ld / np.linalg.norm(ld)                                     # shadow-ray direction (rchit:48)
    a = np.cross(sun, [0.0, 1.0, 0.0]); a /= np.linalg.norm(a)
    b = np.cross(sun, a)
    P = np.stack([a, b, sun])                                          # rows: light-space x, y, depth (towards the sun)
    tl = tris @ P.T                                                    # [n, 3 verts, 3]
    lo, hi = tl.min(1), tl.max(1)
    pl = pos @ P.T
    ext_lo, ext_hi = lo[:, :2].min(0), hi[:, :2].max(0)
    cell = (ext_hi - ext_lo) / G
    c0 = np.clip(((lo[:, :2] - ext_lo) / cell).astype(int), 0, G - 1)
    c1 = np.clip(((hi[:, :2] - ext_lo) / cell).astype(int), 0, G - 1)
    span = (c1 - c0 + 1)
    entries = int((span[:, 0] * span[:, 1]).sum())
    # CSR grid
    counts = np.zeros(G * G, np.int64)
    tri_ids, cell_ids = [], []
    for k in range(len(tris)):
        xs, ys = np.arange(c0[k, 0], c1[k, 0] + 1), np.arange(c0[k, 1], c1[k, 1] + 1)
        cc = (ys[:, None] * G + xs[None, :]).ravel()
        cell_ids.append(cc); tri_ids.append(np.full(cc.size, k))
    cell_ids, tri_ids = np.concatenate(cell_ids), np.concatenate(tri_ids)
    order = np.argsort(cell_ids, kind="stable")
    cell_ids, tri_ids = cell_ids[order], tri_ids[order]
    counts = np.bincount(cell_ids, minlength=G * G)
    start = np.concatenate([[0], np.cumsum(counts)])
    # queries: a sample of the hit points
    rng = np.random.RandomState(1)
    q = rng.choice(len(pl), min(len(pl), 50000), replace=False)
    qc = np.clip(((pl[q, :2] - ext_lo) / cell).astype(int), 0, G - 1)
    qcell = qc[:, 1] * G + qc[:, 0]
    in_cell = counts[qcell]
    ahead = np.zeros(len(q), np.int64)
    tmin, tmax = 0.1, 50.0
    for i, (cid, z) in enumerate(zip(qcell, pl[q, 2])):
        ids = tri_ids[start[cid]:start[cid + 1]]
        ahead[i] = int(((hi[ids, 2] > z + tmin) & (lo[ids, 2] < z + tmax)).sum())
    shadow = max(1, st["shadow_rays"])
    print(f"{name}: {len(tris)} triangles, {len(hit)} hits, {st['shadow_rays']} sun shadow rays traced (dead ones skipped)")
    print(f"oracle's own binary hierarchy: {st['shadow_nodes_visited'] / shadow:.2f} nodes + {st['shadow_tris_tested'] / shadow:.2f} triangle tests per shadow ray "
          f"(the product's BVH8 figures are in the bench line: roofline.nodes_per_shadow_ray / tris_per_shadow_ray)")
    print(f"light-space grid {G} x {G}: cell {cell[0] * 100:.1f} x {cell[1] * 100:.1f} cm, {entries} list entries ({entries * 4 / 2 ** 20:.1f} MiB), "
          f"{entries / len(tris):.1f} cells per triangle")
    for nm, v in (("triangles in the ray's cell", in_cell), ("of those on the sun side within (tMin, tMax): exact tests left", ahead)):
        print(f"  {nm}: mean {v.mean():.1f}, median {np.median(v):.0f}, p90 {np.percentile(v, 90):.0f}, max {v.max()}")

    # ---- is the culling conservative?  For a sample of the sun rays: the oracle's exact any-hit answer (its own hierarchy, validated against
    # brute force) against "any candidate of the ray's grid cell passes the exact Moeller-Trumbore test" — the same ray on both sides, so
    # every disagreement is a triangle the grid dropped.  The footprints are padded by `pad` cells (fp32 projection vs. the exact test).
    if "--verify" in sys.argv:
        o2 = ob.Oracle(make_params(nxz, ny, nr))
        o2.set_threads(0); o2.upload_scene(scene); o2.set_scene_ubo(scene.ubo); o2.build_accel()
        tri32 = tris.astype(np.float32)
        sun32 = sun.astype(np.float32)
        qs = q[:int(next((a.split("=")[1] for a in sys.argv if a.startswith("--rays=")), 4000))]
        for pad in (0, 1):
            c0p, c1p = np.clip(c0 - pad, 0, G - 1), np.clip(c1 + pad, 0, G - 1)
            miss = 0
            for i in qs:
                org = pos[i].astype(np.float32)
                exact = o2.any_hit(org, sun32)
                cx, cy = np.clip(((pl[i, :2] - ext_lo) / cell).astype(int), 0, G - 1)
                ids = np.nonzero((c0p[:, 0] <= cx) & (cx <= c1p[:, 0]) & (c0p[:, 1] <= cy) & (cy <= c1p[:, 1]) & (hi[:, 2] > pl[i, 2] + tmin - 1e-3))[0]
                culled = any(ob.intersect(org, sun32, tri32[k, 0], tri32[k, 1], tri32[k, 2])[0] for k in ids)
                miss += int(exact != culled)
            print(f"conservativeness, footprints padded by {pad} cell(s): {miss} disagreements in {len(qs)} sun rays")
        o2.close()

    # ---- point lights (extension E1): rays share their END point, so the analogue is a direction-space structure around each light
    # (e.g. a cube map of triangle lists).  Estimated per query with angular bounding cones instead of building the cube map: a triangle is a
    # candidate when the ray's direction from the light falls inside the triangle's cone widened by one cell (face resolution G), and its
    # distance interval from the light reaches in front of the hit point.
    lights = [l for l in scene.lights if l["color"][3] > 0.5]
    cell_angle = (np.pi / 2) / G
    tot, left = [], []
    for l in lights:
        lp, rng_m = l["world_pos"][:3].astype(np.float64), float(l["world_pos"][3])
        rel = tris - lp                                                  # [n, 3, 3]
        dist_v = np.linalg.norm(rel, axis=2)
        cdir = rel.mean(1); cdir /= np.linalg.norm(cdir, axis=1, keepdims=True)
        cosv = np.clip(np.einsum("nvk,nk->nv", rel / np.maximum(dist_v[..., None], 1e-12), cdir), -1, 1)
        rho = np.arccos(cosv.min(1)) + cell_angle
        dmin, dmax = dist_v.min(1), dist_v.max(1)
        d = pos - lp
        dd = np.linalg.norm(d, axis=1)
        sel = np.nonzero(dd < rng_m)[0]
        if sel.size == 0:
            continue
        for i in rng.choice(sel, min(sel.size, 400), replace=False):
            qd = d[i] / dd[i]
            ang = np.arccos(np.clip(cdir @ qd, -1, 1))
            cand = ang <= rho
            tot.append(int(cand.sum()))
            left.append(int((cand & (dmin < dd[i] - tmin)).sum()))
    if tot:
        tot, left = np.array(tot), np.array(left)
        print(f"point lights ({len(lights)} enabled, cube map of {G} x {G} per face, estimated with bounding cones):")
        for nm, v in (("triangles in the ray's direction cell", tot), ("of those nearer to the light than the hit point: exact tests left", left)):
            print(f"  {nm}: mean {v.mean():.1f}, median {np.median(v):.0f}, p90 {np.percentile(v, 90):.0f}, max {v.max()}")


if __name__ == "__main__":
    main()
