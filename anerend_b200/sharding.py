"""Probe-slab sharding across the GPUs of one box.

Probes are independent within one update, so rank r of R owns the contiguous Y-layers
[r*nY/R, (r+1)*nY/R) — a contiguous row range of the atlas "ProbeRaysConvTex" — traces and convolves only
those, and the atlas is re-assembled on every rank with ONE all-gather per update (next update's
sampleProbe reads neighbouring layers, pbr_light.glsl:318-355).  Scene and BVH are replicated.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing; the collective runs on the
handle's own CUDA stream, in place on the atlas memory.
"""
import numpy as np


def slab_for_rank(num_probes_y, rank, world):
    """Contiguous layer range of `rank`; layers are split as evenly as possible, every rank gets at least one."""
    if world > num_probes_y:
        raise ValueError(f"{world} ranks cannot share {num_probes_y} probe layers")
    base, rem = divmod(num_probes_y, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def layers_for_rank(num_probes_y, rank, world):
    """Interleaved ("zigzag") layer set of `rank` for a balanced multi-GPU update: layers are dealt to the ranks in
    boustrophedon order (0..R-1, R-1..0, 0..R-1, ...).  Work per layer falls with height in any scene that has a
    ground (lower layers hit more geometry and spawn more shadow rays); contiguous slabs gave the bottom rank 1.6x the
    work of the top one at C5, a zigzag deal cancels a linear trend exactly.  Every rank gets floor or ceil(nY/R) layers."""
    if world > num_probes_y:
        raise ValueError(f"{world} ranks cannot share {num_probes_y} probe layers")
    out = []
    for y in range(num_probes_y):
        k, pos = divmod(y, world)
        owner = pos if k % 2 == 0 else world - 1 - pos
        if owner == rank:
            out.append(y)
    return out


def mirrored_slabs_for_rank(num_probes_y, rank, world):
    """Balanced AND contiguous: rank r owns chunk r of the lower half of the layers and chunk R-1-r of the upper half
    (k = nY/(2R) layers each).  Like the zigzag deal it cancels a linear work-vs-height trend, but each half is a plain
    all-gather of contiguous atlas rows (in place for the lower half; into reversed row blocks for the upper half), so
    the exchange needs no per-texel pack/scatter.  Needs nY divisible by 2R."""
    if num_probes_y % (2 * world):
        raise ValueError("mirrored slabs need num_probes_y divisible by twice the world size")
    k = num_probes_y // (2 * world)
    half = num_probes_y // 2
    return list(range(rank * k, (rank + 1) * k)) + list(range(half + (world - 1 - rank) * k, half + (world - rank) * k))


def balanced_layers(costs, world):
    """Measured-cost deal of the layers (longest-processing-time-first): `costs[y]` is the time (any unit) one update of
    layer y takes; layers are handed out in order of falling cost, each to the rank with the least work so far (ties ->
    lowest rank, so every rank derives the same deal from the same costs).  Returns one sorted layer list per rank; every
    rank gets at least one layer.  Jacobi updates do not depend on how the layers are dealt, so the deal can be
    re-measured at any time; it needs an exchange that takes arbitrary layer sets (the fused peer-store exchange)."""
    n = len(costs)
    if world > n:
        raise ValueError(f"{world} ranks cannot share {n} probe layers")
    order = sorted(range(n), key=lambda y: (-float(costs[y]), y))
    load, out = [0.0] * world, [[] for _ in range(world)]
    for i, y in enumerate(order):
        remaining = n - i
        empty = [r for r in range(world) if not out[r]]
        # never leave a rank without a layer: once the layers left only just cover the empty ranks, serve those
        cand = empty if empty and remaining <= len(empty) else range(world)
        r = min(cand, key=lambda k: (load[k], k))
        out[r].append(y)
        load[r] += float(costs[y])
    return [sorted(o) for o in out]


def measure_layer_costs(handle, rots, repeats=2):
    """Per-layer update time in ms (trace + convolve, CUDA events of the handle), best of `repeats`: one single-layer
    update_layers call per layer.  Run it on an UNCONNECTED handle (no peer stores); it advances the atlas like any
    update, so callers reset() afterwards if they need the cleared state."""
    ny = handle.n_y
    costs = [float("inf")] * ny
    for _ in range(repeats):
        for y in range(ny):
            handle.update_layers([y], rots)
            t = handle.timings()
            costs[y] = min(costs[y], t["rt_ms"] + t["convolve_ms"])
    return costs


_STAGING = {}


def _staging(like, numel):
    key = (like.device, like.dtype, numel)
    if key not in _STAGING:
        import torch
        _STAGING[key] = torch.empty(numel, dtype=like.dtype, device=like.device)
    return _STAGING[key]


def all_gather_atlas_mirrored(atlas, num_probes_xz, num_probes_y, rank, world, group=None):
    """Exchange for `mirrored_slabs_for_rank`: an in-place all_gather_into_tensor of the lower half of the atlas rows and
    an all_gather of the upper half whose output blocks are the row ranges in reversed rank order."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return atlas
    a = atlas.view(torch.int64)                   # one RGBA16F texel per element; bit copy
    rows = 10 * num_probes_xz
    k = num_probes_y // (2 * world)
    half_rows = rows * (num_probes_y // 2)
    lo, hi = a[:half_rows], a[half_rows:]
    b = rank * k * rows
    dist.all_gather_into_tensor(lo.reshape(-1), lo[b:b + k * rows].reshape(-1), group=group)
    # upper half: rank r owns block R-1-r; gather into a staging buffer in rank order, then one flipped block copy
    b = (world - 1 - rank) * k * rows
    staging = _staging(a, hi.numel())
    dist.all_gather_into_tensor(staging, hi[b:b + k * rows].reshape(-1), group=group)
    hi.view(world, -1).copy_(staging.view(world, -1).flip(0))
    return atlas


def all_gather_atlas_layers(atlas, num_probes_xz, num_probes_y, rank, world, group=None):
    """All-gather for interleaved layer sets (`layers_for_rank`).  `atlas` is a torch tensor [H, W, 4]; a rank's layers
    are not contiguous rows, so its rows are packed into a send buffer, gathered, and scattered back (two extra
    device copies of the atlas per update).  Needs nY divisible by R (equal chunks)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return atlas
    if num_probes_y % world:
        raise ValueError("interleaved gather needs num_probes_y divisible by the world size")
    rows = 10 * num_probes_xz
    a = atlas.view(torch.int64).view(num_probes_y, rows, atlas.shape[1])              # [nY, rows, W] texels (bit copy)
    owners = [layers_for_rank(num_probes_y, r, world) for r in range(world)]
    idx = torch.as_tensor(owners[rank], device=a.device)
    send = a.index_select(0, idx).contiguous()
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=a.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    order = torch.as_tensor([y for o in owners for y in o], device=a.device)
    a.index_copy_(0, order, recv.view((num_probes_y,) + tuple(send.shape[1:])))
    return atlas


def all_gather_atlas_stripes(atlas, num_probes_xz, num_probes_y, rank, world, group=None):
    """All-gather for row stripes (`_set_row_stripe(rank, world)`): rank r owns the probe rows z = r mod R of every layer, i.e.
    the 10-row blocks r, r+R, ... of the atlas.  Packed send buffer, one all_gather_into_tensor, one strided scatter back.
    `atlas` is a torch tensor [H, W, 4]; needs nXZ divisible by R (equal stripes)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return atlas
    if num_probes_xz % world:
        raise ValueError("stripe gather needs num_probes_xz divisible by the world size")
    per = num_probes_xz // world
    a = atlas.view(torch.int64).view(num_probes_y, per, world, 10, atlas.shape[1])        # [layer][row k][owner][10 atlas rows][W] texels
    send = a[:, :, rank].contiguous()
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=a.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    a.copy_(recv.permute(1, 2, 0, 3, 4))
    return atlas


def slab_rows(num_probes_xz, slab):
    """Atlas row range [begin, end) of a layer slab: 10*nXZ rows per layer (IrradianceProbeConvolvePass.cpp:84-99)."""
    return 10 * num_probes_xz * slab[0], 10 * num_probes_xz * slab[1]


class DeviceArray:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr="<u2"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def all_gather_atlas(atlas, num_probes_xz, num_probes_y, rank, world, group=None):
    """In-place all-gather of the atlas slabs.  `atlas` is a torch tensor [H, W, 4] (any device); every rank
    contributes the rows of its own slab and receives all the others."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return atlas
    full = atlas
    atlas = atlas.view(torch.uint8)          # bit copy; uint8 is supported by both NCCL and gloo
    slabs = [slab_for_rank(num_probes_y, r, world) for r in range(world)]
    rows = [slab_rows(num_probes_xz, s) for s in slabs]
    sizes = {e - b for b, e in rows}
    if len(sizes) == 1:
        # equal slabs: one all_gather_into_tensor, send buffer is a view of the receive buffer (in place)
        b, e = rows[rank]
        dist.all_gather_into_tensor(atlas.view(-1), atlas[b:e].reshape(-1), group=group)
    else:
        # unequal slabs (nY not divisible by R): one in-place broadcast of each rank's rows
        for r, (b, e) in enumerate(rows):
            dist.broadcast(atlas[b:e], src=r, group=group)
    return full


def merge_slabs_numpy(slab_atlases, num_probes_xz, num_probes_y):
    """Host-side equivalent of the all-gather (tests): take each rank's own rows."""
    world = len(slab_atlases)
    out = np.array(slab_atlases[0], copy=True)
    for r in range(world):
        b, e = slab_rows(num_probes_xz, slab_for_rank(num_probes_y, r, world))
        out[b:e] = slab_atlases[r][b:e]
    return out
