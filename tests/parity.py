"""Comparison helpers shared by the parity tests (oracle vs host-emu vs CUDA)."""
import numpy as np

from anerend_b200 import layouts as L

RAY_REL_TOL = 1e-4      # north_star: per-ray radiance within 1e-4 relative
TEXEL_REL_TOL = 1e-4    # north_star: convolved texels within 1e-4 relative after 1 update (fp32, before the rgba16f store)
TEXEL_REL_TOL_120 = 5e-4


def half_ordered(h):
    """uint16 half bit patterns -> integers whose difference is the distance in half ulps."""
    h = h.astype(np.int32)
    return np.where(h & 0x8000, -(h & 0x7FFF), h & 0x7FFF)


def half_ulp_distance(a, b):
    return np.abs(half_ordered(a) - half_ordered(b))


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def setup_pair(scene, handles, seed, build=True):
    rots = None
    for h in handles:
        h.upload_scene(scene)
        h.set_scene_ubo(scene.ubo)
        if build:
            h.build_accel()
    from anerend_b200 import scenes
    seq = scenes.rotation_sequence(seed, handles[0].n_y)
    rots = handles[0].rotations(seq)
    for h in handles[1:]:
        assert np.array_equal(h.rotations(seq), rots), "make_rotation differs between implementations"
    return rots


def compare_records(ref, got, allow_mismatch_frac=0.0):
    """Hit ids bit-exact (except documented equal-t ties), radiance within 1e-4 relative."""
    mism = ref["prim"] != got["prim"]
    n_mism = int(mism.sum())
    if n_mism:
        # the only documented mismatch class: equal-t ties cannot occur (both sides break ties by lowest id), so
        # anything here is a grazing-ray box/triangle disagreement; it must stay below the allowed fraction
        assert n_mism <= allow_mismatch_frac * mism.size, f"{n_mism} of {mism.size} hit ids differ"
    ok = ~mism
    assert np.array_equal(ref["t"][ok], got["t"][ok]) and np.array_equal(ref["u"][ok], got["u"][ok]) \
        and np.array_equal(ref["v"][ok], got["v"][ok]), "t/u/v differ on identical hit ids"
    same_lit = ref["lit_mask"][ok] == got["lit_mask"][ok]
    assert same_lit.mean() >= 1.0 - allow_mismatch_frac, "shadow-ray results differ"
    sel = ok.copy()
    sel[ok] = same_lit
    err = rel_err(got["radiance"][sel], ref["radiance"][sel], floor=1e-3)
    assert err.size == 0 or err.max() <= RAY_REL_TOL, f"per-ray radiance rel err {err.max():.3e}"
    return n_mism


def compare_atlas(ref_u16, got_u16, max_ulps=1):
    """Stored RGBA16F atlas: identical up to `max_ulps` half ulps (one ulp = the rounding step of the store)."""
    d = half_ulp_distance(ref_u16, got_u16)
    assert d.max() <= max_ulps, f"atlas differs by {d.max()} half ulps at {np.argwhere(d == d.max())[0]}"
    return float((d == 0).mean())


def interior_mask(n_xz, n_y):
    """True for texels the convolution writes (8x8 interior of each 10x10 probe tile)."""
    W, H = 10 * n_xz, 10 * n_xz * n_y
    x = (np.arange(W) % 10)
    y = (np.arange(H) % 10)
    return ((y[:, None] >= 1) & (y[:, None] <= 8)) & ((x[None, :] >= 1) & (x[None, :] <= 8))


def decode_bvh8(nodes_u8, tris_u8):
    """Decode the 80-byte BVH8 nodes and 48-byte triangle records (layout: csrc/ddgi_kernels.cuh)."""
    nodes = nodes_u8.reshape(-1, 80)
    w = nodes.view(np.uint32).reshape(-1, 20)
    p = w[:, 0:3].copy().view(np.float32)
    e = np.stack([w[:, 3] & 0xFF, (w[:, 3] >> 8) & 0xFF, (w[:, 3] >> 16) & 0xFF], 1)
    imask = (w[:, 3] >> 24) & 0xFF
    child_base, tri_base, tri_mask = w[:, 4], w[:, 5], w[:, 6]
    q = nodes[:, 32:80].reshape(-1, 6, 8)        # lox, loy, loz, hix, hiy, hiz
    scale = (e.astype(np.uint32) << 23).view(np.float32)
    lo = p[:, None, :] + q[:, 0:3, :].transpose(0, 2, 1).astype(np.float32) * scale[:, None, :]
    hi = p[:, None, :] + q[:, 3:6, :].transpose(0, 2, 1).astype(np.float32) * scale[:, None, :]
    t = tris_u8.view(np.float32).reshape(-1, 12)
    v0 = t[:, 0:3]
    e1 = t[:, 3:6]
    e2 = t[:, 6:9]
    gid = t[:, 9].copy().view(np.uint32)
    return dict(p=p, e=e, imask=imask, child_base=child_base, tri_base=tri_base, tri_mask=tri_mask, lo=lo, hi=hi,
                v0=v0, v1=v0 + e1, v2=v0 + e2, gid=gid)


def validate_bvh8(nodes_u8, tris_u8, num_tris):
    """Structural invariants of the BVH8: every triangle stored once, leaves bound their triangles, children nest."""
    b = decode_bvh8(nodes_u8, tris_u8)
    assert np.array_equal(np.sort(b["gid"]), np.arange(num_tris, dtype=np.uint32)), "triangle set is not a permutation"
    nn = len(b["p"])
    seen_nodes = np.zeros(nn, bool)
    seen_nodes[0] = True
    seen_tris = np.zeros(num_tris, bool)
    true_lo = np.full((nn, 3), np.inf)      # exact bounds of the triangles below each node
    true_hi = np.full((nn, 3), -np.inf)
    for n in range(nn - 1, -1, -1):         # children are always allocated after their parent
        rank = 0
        tm = int(b["tri_mask"][n])
        assert tm < (1 << 24) and int(nodes_u8.reshape(-1, 80).view(np.uint32)[n, 7]) == 0
        for s in range(8):
            leaf_bits = (tm >> (3 * s)) & 7
            inner = (int(b["imask"][n]) >> s) & 1
            lo, hi = b["lo"][n, s], b["hi"][n, s]
            if not leaf_bits and not inner:
                assert np.any(lo > hi), "empty slot must never pass the box test"
                continue
            assert not (leaf_bits and inner) and leaf_bits in (0, 1, 3, 7)
            if inner:
                c = int(b["child_base"][n]) + rank
                rank += 1
                assert c > n and not seen_nodes[c], "node referenced twice / not after its parent"
                seen_nodes[c] = True
                assert np.all(true_lo[c] >= lo) and np.all(true_hi[c] <= hi), f"subtree of node {c} leaks out of slot {s} of node {n}"
                true_lo[n] = np.minimum(true_lo[n], true_lo[c])
                true_hi[n] = np.maximum(true_hi[n], true_hi[c])
            else:
                cnt = bin(leaf_bits).count("1")
                off = bin(tm & ((1 << (3 * s)) - 1)).count("1")
                assert 1 <= cnt <= 3
                for j in range(cnt):
                    ti = int(b["tri_base"][n]) + off + j
                    assert not seen_tris[ti], "triangle slot referenced twice"
                    seen_tris[ti] = True
                    for v in (b["v0"][ti], b["v1"][ti], b["v2"][ti]):
                        assert np.all(v >= lo) and np.all(v <= hi), f"triangle {ti} leaks out of its leaf box in node {n}"
                        true_lo[n] = np.minimum(true_lo[n], v)
                        true_hi[n] = np.maximum(true_hi[n], v)
    assert seen_nodes.all() and seen_tris.all(), "unreachable nodes or triangles"
    return b
