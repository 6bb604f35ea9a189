"""Fused atlas exchange (anerend_cuda_ddgi_exchange_*): two PROCESSES, each with its own handle and its own set of probe
layers, map each other's atlas through cudaIpc; the convolution stores into both atlases and the device-side barriers
order the updates.  Both ranks must end with the atlas one handle computes alone (jacobi updates do not depend on how
the layers are dealt).  On the one-GPU test box both processes share cuda:0 (the driver time-slices them); the bench
runs the same code with one GPU per rank over NVLink."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRID = (8, 4, 64)
UPDATES = 3


def _scene_and_rots():
    from anerend_b200 import scenes
    scene = scenes.cornell_box(grid=GRID, wall_res=8, box_res=3)
    return scene, scenes.rotation_sequence(7, GRID[1] * UPDATES)


def _worker(rank, world, port, out_dir, pipelined=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import anerend_b200 as A
    from anerend_b200 import layouts as L, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    nxz, ny, nr = GRID
    scene, seq = _scene_and_rots()
    h = A.Ddgi(A.make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS, device=dev))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    blobs = [None] * world
    dist.all_gather_object(blobs, h.exchange_export())
    h.exchange_connect(blobs, world, rank)
    if pipelined:                                # the next update's closest-hit trace under the exchange of this one
        h.set_timers(0)
        h.set_pipeline(True)
    dist.barrier()
    rots = h.rotations(seq).reshape(UPDATES, ny, 16)
    layers = sharding.layers_for_rank(ny, rank, world)
    for k in range(UPDATES):
        h.translate()
        h.update_layers(layers, rots[k])
    h.sync()                                     # raises if a barrier timed out
    dist.barrier()
    np.save(os.path.join(out_dir, f"atlas_{rank}.npy"), h.download_atlas())
    h.exchange_disconnect()
    dist.barrier()
    h.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("pipelined", [False, True], ids=["plain", "pipelined"])
def test_two_processes_assemble_the_atlas_through_peer_stores(pipelined, tmp_path, make_gpu):
    import torch.multiprocessing as mp
    from anerend_b200 import layouts as L, make_params
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), pipelined), nprocs=world, join=True)
    nxz, ny, nr = GRID
    scene, seq = _scene_and_rots()
    h = make_gpu(make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS))
    h.upload_scene(scene)
    h.set_scene_ubo(scene.ubo)
    h.build_accel()
    rots = h.rotations(seq).reshape(UPDATES, ny, 16)
    for k in range(UPDATES):
        h.translate()
        h.update(-1, rots[k])
    single = h.download_atlas()
    h.close()
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"atlas_{r}.npy"))
        assert np.array_equal(got, single), f"rank {r}: {(got != single).sum()} differing atlas words"


@pytest.mark.gpu
def test_exchange_argument_checks(make_gpu):
    import anerend_b200 as A
    from anerend_b200 import make_params
    h = make_gpu(make_params(4, 2, 64))
    with pytest.raises(A.DdgiError):
        h.exchange_connect([b"\0" * 128] * 2, 2, 0)          # connect before export
    blob = h.exchange_export()
    assert len(blob) == 128 and any(blob)
    with pytest.raises(A.DdgiError):
        h.exchange_connect([blob], 1, 0)                     # needs 2..8 ranks
    with pytest.raises(A.DdgiError):
        h.exchange_connect([blob] * 2, 2, 5)                 # rank out of range
    h.exchange_disconnect()                                  # harmless when not connected
    h.close()
