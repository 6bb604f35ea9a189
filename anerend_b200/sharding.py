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
