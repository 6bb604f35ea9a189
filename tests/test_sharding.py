"""Host-side logic of the Y-slab sharding on CPU: slab arithmetic, and a world_size-2 gloo run in which
each rank updates only its slab with the host-emu harness, all-gathers the atlas, and ends with the single-rank atlas."""
import os
import socket
import sys

import numpy as np
import pytest

from anerend_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_arithmetic():
    for ny, world in ((8, 1), (8, 2), (12, 4), (32, 8), (12, 8), (5, 2)):
        slabs = [sharding.slab_for_rank(ny, r, world) for r in range(world)]
        assert slabs[0][0] == 0 and slabs[-1][1] == ny
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:])) and all(e > b for b, e in slabs)
        assert max(e - b for b, e in slabs) - min(e - b for b, e in slabs) <= 1
    assert sharding.slab_rows(24, (3, 6)) == (720, 1440)
    with pytest.raises(ValueError):
        sharding.slab_for_rank(4, 0, 8)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import ctypes
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from anerend_b200 import layouts as L, make_params, scenes
    from anerend_b200.capi import DdgiHandle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = ctypes.CDLL(g.build_host_emu())
    nxz, ny, nr = 4, 3, 64                       # 3 layers over 2 ranks: unequal slabs (2 + 1)
    scene = scenes.random_soup(51, ntris=400, grid=(nxz, ny, nr))
    slab = sharding.slab_for_rank(ny, rank, world)
    h = DdgiHandle(emu, "anerend_emu_ddgi_", make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS, slab=slab))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    rots = h.rotations(scenes.rotation_sequence(9, ny))
    for it in range(3):
        h.update(-1, rots)
        atlas = torch.from_numpy(h.download_atlas().view(np.int16).copy())
        sharding.all_gather_atlas(atlas, nxz, ny, rank, world)
        h.upload_atlas(atlas.numpy().view(np.uint16))
    np.save(os.path.join(out_dir, f"atlas_{rank}.npy"), h.download_atlas())
    dist.barrier()
    dist.destroy_process_group()


def _worker_interleaved(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import ctypes
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from anerend_b200 import layouts as L, make_params, scenes
    from anerend_b200.capi import DdgiHandle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = ctypes.CDLL(g.build_host_emu())
    nxz, ny, nr = 4, 4, 64
    scene = scenes.random_soup(52, ntris=400, grid=(nxz, ny, nr))
    h = DdgiHandle(emu, "anerend_emu_ddgi_", make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    rots = h.rotations(scenes.rotation_sequence(9, ny))
    for it in range(4):
        # alternate the two balanced splits: zigzag deal (packed gather) and mirrored slabs (two in-place gathers)
        mirrored = it % 2 == 1
        mine = sharding.mirrored_slabs_for_rank(ny, rank, world) if mirrored else sharding.layers_for_rank(ny, rank, world)
        h.update_layers(mine, rots)
        atlas = torch.from_numpy(h.download_atlas().view(np.int16).copy())
        if mirrored:
            sharding.all_gather_atlas_mirrored(atlas, nxz, ny, rank, world)
        else:
            sharding.all_gather_atlas_layers(atlas, nxz, ny, rank, world)
        h.upload_atlas(atlas.numpy().view(np.uint16))
    np.save(os.path.join(out_dir, f"atlas_i{rank}.npy"), h.download_atlas())
    dist.barrier()
    dist.destroy_process_group()


def test_layers_for_rank_is_a_balanced_partition():
    for ny, world in ((32, 8), (32, 4), (32, 2), (12, 4), (8, 8), (5, 2)):
        sets = [sharding.layers_for_rank(ny, r, world) for r in range(world)]
        assert sorted(sum(sets, [])) == list(range(ny))
        assert max(map(len, sets)) - min(map(len, sets)) <= 1
        if ny % (2 * world) == 0:                 # a linear work profile is balanced exactly
            assert len({sum(s) for s in sets}) == 1
            mir = [sharding.mirrored_slabs_for_rank(ny, r, world) for r in range(world)]
            assert sorted(sum(mir, [])) == list(range(ny)) and len({sum(s) for s in mir}) == 1


def test_two_rank_gloo_interleaved_update_equals_single_rank(tmp_path, make_emu):
    import torch.multiprocessing as mp
    from anerend_b200 import layouts as L, make_params, scenes
    import __graft_entry__ as g
    g.build_host_emu()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_interleaved, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    nxz, ny, nr = 4, 4, 64
    scene = scenes.random_soup(52, ntris=400, grid=(nxz, ny, nr))
    h = make_emu(make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    rots = h.rotations(scenes.rotation_sequence(9, ny))
    for it in range(4):
        h.update(-1, rots)
    want = h.download_atlas()
    a0, a1 = np.load(tmp_path / "atlas_i0.npy"), np.load(tmp_path / "atlas_i1.npy")
    assert np.array_equal(a0, a1) and np.array_equal(a0, want)


def test_two_rank_gloo_slab_update_equals_single_rank(tmp_path, make_emu):
    import torch.multiprocessing as mp
    from anerend_b200 import layouts as L, make_params, scenes
    import __graft_entry__ as g
    g.build_host_emu()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    nxz, ny, nr = 4, 3, 64
    scene = scenes.random_soup(51, ntris=400, grid=(nxz, ny, nr))
    h = make_emu(make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    rots = h.rotations(scenes.rotation_sequence(9, ny))
    for it in range(3):
        h.update(-1, rots)
    want = h.download_atlas()
    a0, a1 = np.load(tmp_path / "atlas_0.npy"), np.load(tmp_path / "atlas_1.npy")
    assert np.array_equal(a0, a1) and np.array_equal(a0, want)
    merged = sharding.merge_slabs_numpy([a0, a1], nxz, ny)
    assert np.array_equal(merged, want)


def test_balanced_layers_is_a_partition_and_beats_the_static_deals():
    rng = np.random.default_rng(5)
    for ny, world in ((32, 8), (32, 4), (12, 2), (12, 5), (8, 8)):
        # cost falls with height, plus a bump (a dense layer of props) and noise: what C5 looks like
        costs = [3.0 * np.exp(-y / (0.3 * ny)) + (1.5 if y == ny // 3 else 0.0) + 0.2 * rng.random() + 0.3 for y in range(ny)]
        deal = sharding.balanced_layers(costs, world)
        assert sorted(y for d in deal for y in d) == list(range(ny)) and all(d for d in deal)
        load = [sum(costs[y] for y in d) for d in deal]
        zig = [sum(costs[y] for y in sharding.layers_for_rank(ny, r, world)) for r in range(world)]
        assert max(load) <= max(zig) + 1e-9
        # LPT bound: the busiest rank carries at most the mean plus one layer
        assert max(load) <= sum(costs) / world + max(costs) + 1e-9
    assert sharding.balanced_layers([1.0, 1.0, 1.0], 3) == [[0], [1], [2]]
    with pytest.raises(ValueError):
        sharding.balanced_layers([1.0, 2.0], 3)


def test_download_atlas_layers_returns_the_layer_rows(make_emu):
    from anerend_b200 import layouts as L, make_params, scenes
    from tests import parity
    scene = scenes.random_soup(31, ntris=300)
    e = make_emu(make_params(4, 3, 64, flags=L.FLAG_POINT_LIGHTS))
    rots = parity.setup_pair(scene, (e,), seed=9)
    e.update(-1, rots)
    full = e.download_atlas()
    rows = 10 * 4
    got = e.download_atlas_layers([2, 0])
    assert got.shape == (2, rows, rows, 4)
    assert np.array_equal(got[0], full[2 * rows:3 * rows]) and np.array_equal(got[1], full[0:rows])
    with pytest.raises(Exception):
        e.download_atlas_layers([3])
    e.close()


def _stripe_merge(handles, nxz, ny, world):
    """What the peer push / an all-gather does for row stripes: probe row z of every layer comes from handle z % world."""
    parts = [h.download_atlas() for h in handles]
    merged = parts[0].copy()
    for y in range(ny):
        for z in range(nxz):
            r0 = 10 * nxz * y + 10 * z
            merged[r0:r0 + 10] = parts[z % world][r0:r0 + 10]
    return merged


def test_row_stripes_assemble_the_full_atlas(make_emu):
    """Sharding inside the layers: handle r updates the probe rows z = r mod R of every layer; after exchanging rows
    the atlas equals the one a single handle computes (jacobi updates), also when nXZ is not a multiple of R."""
    from anerend_b200 import layouts as L, make_params, scenes
    from tests import parity
    nxz, ny, nr, world = 5, 2, 64, 3
    scene = scenes.random_soup(41, ntris=400, grid=(nxz, ny, nr))
    p = make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS)
    full, parts = make_emu(p), [make_emu(p) for _ in range(world)]
    rots = parity.setup_pair(scene, [full] + parts, seed=12)
    for r, h in enumerate(parts):
        h.set_row_stripe(r, world)
    for it in range(2):
        full.update(-1, rots)
        for h in parts:
            h.update(-1, rots)
        merged = _stripe_merge(parts, nxz, ny, world)
        for r, h in enumerate(parts):
            own = h.download_atlas_stripe()
            rows = list(range(r, nxz, world))
            assert own.shape == (ny, len(rows), 10, 10 * nxz, 4)
            for y in range(ny):
                for k, z in enumerate(rows):
                    r0 = 10 * nxz * y + 10 * z
                    assert np.array_equal(own[y, k], merged[r0:r0 + 10])
            h.upload_atlas(merged)
        assert np.array_equal(merged, full.download_atlas())
    with pytest.raises(Exception):
        parts[0].set_row_stripe(3, 3)
    for h in [full] + parts:
        h.close()


def _worker_stripes(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import ctypes
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from anerend_b200 import layouts as L, make_params, scenes
    from anerend_b200.capi import DdgiHandle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = ctypes.CDLL(g.build_host_emu())
    nxz, ny, nr = 4, 3, 64
    scene = scenes.random_soup(53, ntris=400, grid=(nxz, ny, nr))
    h = DdgiHandle(emu, "anerend_emu_ddgi_", make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    h.set_row_stripe(rank, world)
    rots = h.rotations(scenes.rotation_sequence(9, ny))
    for it in range(3):
        h.update(-1, rots)
        atlas = torch.from_numpy(h.download_atlas().view(np.int16).copy())
        sharding.all_gather_atlas_stripes(atlas, nxz, ny, rank, world)
        h.upload_atlas(atlas.numpy().view(np.uint16))
    np.save(os.path.join(out_dir, f"atlas_s{rank}.npy"), h.download_atlas())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_row_stripes_equal_single_rank(tmp_path, make_emu):
    """The default multi-GPU partition (row stripes) with a collective exchange, world size 2 over gloo."""
    import torch.multiprocessing as mp
    from anerend_b200 import layouts as L, make_params, scenes
    import __graft_entry__ as g
    g.build_host_emu()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_stripes, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    nxz, ny, nr = 4, 3, 64
    scene = scenes.random_soup(53, ntris=400, grid=(nxz, ny, nr))
    h = make_emu(make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    rots = h.rotations(scenes.rotation_sequence(9, ny))
    for it in range(3):
        h.update(-1, rots)
    single = h.download_atlas()
    h.close()
    for r in range(2):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), f"atlas_s{r}.npy")), single)
