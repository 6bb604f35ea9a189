"""anerend_cuda_ddgi_refit_accel_device: the TLASUpdate entry point of the pass adapter.  The renderable buffer it is handed lives on
the device and is written by the GPU earlier in the frame — transforms by the animation / particle passes, the cull bit by cull.comp
(296-303) whenever a renderable crosses the 50 m range — and the reference rebuilds its TLAS from it every frame (UpdateTLASPass.cpp:
198-228).  Moved transforms must be refitted; a flipped visible / cull bit must be NOTICED (the instance set differs) and lead to a
rebuild, never to a silently stale structure."""
import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

pytestmark = pytest.mark.gpu
FLAGS = L.FLAG_RAY_RECORDS | L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS


def _device_copy(arr):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).cuda()
    torch.cuda.synchronize()
    return t


def _equal(o, g, ny):
    for layer in range(ny):
        ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
        assert np.array_equal(ro["prim"], rg["prim"]) and np.array_equal(ro["lit_mask"], rg["lit_mask"])
    assert np.array_equal(o.download_atlas(), g.download_atlas())


def test_device_renderables_moved_transforms_and_flipped_cull_bits(oracle, make_gpu):
    scene = scenes.sponza_like(seed=4, grid=(6, 3, 64), movers=6, curtain_res=6)
    nxz, ny, nr = 6, 3, 64
    p = make_params(nxz, ny, nr, flags=FLAGS)
    o, g = oracle.Oracle(p), make_gpu(p)
    rots = parity.setup_pair(scene, (o, g), seed=5)
    n0 = g.bvh_info().num_triangles
    base = scene.renderables.copy()
    # 1. transforms only: refit on both sides
    rend = scene.animated_renderables(0.7)
    d = _device_copy(rend)
    o.refit_accel(rend)
    g.refit_accel_device(d.data_ptr(), len(rend))
    for h in (o, g):
        h.update(-1, rots)
    assert g.bvh_info().num_triangles == n0
    _equal(o, g, ny)
    # 2. the GPU culls two renderables (cull bit) and hides one (visible bit): the instance set shrinks.  The oracle is told through
    # its host path (new scene + build); the CUDA handle only sees the device buffer.
    rend2 = scene.animated_renderables(1.1).copy()
    victims = [0, 5, int(scene.movers[0][0])]           # the floor, a wall strip / column, a mover
    rend2["visible"][victims[0]] |= 2
    rend2["visible"][victims[1]] = 0
    rend2["visible"][victims[2]] |= 2
    d2 = _device_copy(rend2)
    g.refit_accel_device(d2.data_ptr(), len(rend2))
    scene.renderables = rend2
    o.upload_scene(scene)
    o.build_accel()
    assert g.bvh_info().num_triangles == o.bvh_info().num_triangles < n0, "a flipped cull / visible bit in the device buffer went unnoticed"
    for h in (o, g):
        h.reset()
        h.update(-1, rots)
    _equal(o, g, ny)
    # 3. they come back (and the movers move on): rebuilt again
    scene.renderables = base
    rend3 = scene.animated_renderables(1.5).copy()
    d3 = _device_copy(rend3)
    g.refit_accel_device(d3.data_ptr(), len(rend3))
    scene.renderables = rend3
    o.upload_scene(scene)
    o.build_accel()
    assert g.bvh_info().num_triangles == o.bvh_info().num_triangles == n0
    for h in (o, g):
        h.reset()
        h.update(-1, rots)
    _equal(o, g, ny)
    o.close()
    g.close()


def _with_cull_bits(rend, cam, range_m):
    """cull.comp:241-244,296-303 restated: the cull bit of every visible renderable whose bounding-sphere centre (bounds.xyz moved by the
    translation of its transform) lies farther than `range_m` from the camera (fp32, like the shader)."""
    out = rend.copy()
    centre = out["bounds"][:, :3].astype(np.float32) + out["transform"][:, 12:15].astype(np.float32)
    d = (cam.astype(np.float32)[None, :] - centre).astype(np.float32)
    dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
    vis = (out["visible"] & 1) != 0
    out["visible"] = np.where(vis, (out["visible"] & ~np.uint32(2)) | np.where(dist <= np.float32(range_m), 0, 2).astype(np.uint32), out["visible"])
    return out


def _walk(scene, o, g, ny, frames, range_m):
    from anerend_b200.capi import DdgiError
    rots = o.rotations(scenes.rotation_sequence(4, ny * frames)).reshape(frames, ny, 16)
    rebuilds, sizes = 0, set()
    for k in range(frames):
        cam = np.array([-10.0 + 20.0 * k / (frames - 1.0), 4.0, 0.5 * np.sin(k / 5.0)])
        ubo = scene.with_camera(cam)
        rend = _with_cull_bits(scene.animated_renderables(k / 30.0), cam, range_m)       # what the engine's RenderableBuffer holds this frame
        rebuilt = []
        for h in (o, g):
            h.set_scene_ubo(ubo)
            try:                      # TLASUpdate: refit, or rebuild when the instance set changed (the adapter's logic)
                h.refit_accel(rend)
                rebuilt.append(False)
            except DdgiError as e:
                assert e.code == -3, e
                h.build_accel()
                rebuilt.append(True)
            h.translate()
            h.update(-1, rots[k])
        # (the oracle re-selects and rebuilds on every refit and only reports a change of the COUNTS; the product compares the instance set)
        rebuilds += int(rebuilt[1])
        sizes.add(g.bvh_info().num_triangles)
        assert g.bvh_info().num_triangles == o.bvh_info().num_triangles
    return rebuilds, sizes


def test_camera_walk_flips_cull_bits_and_rebuilds(oracle, make_gpu):
    """C4-style scroll where the camera's 50 m cull range (cull.comp:296-303; here 9 m so that it bites in a 30 m scene) keeps adding and
    removing renderables: whenever a bit flips the structure is rebuilt (what UpdateTLASPass.cpp:198-228 does every frame), otherwise the
    movers are refitted; translation and multi-bounce as in C4.  Final atlas and hit ids bit-identical to the oracle's."""
    scene = scenes.sponza_like(seed=4, grid=(8, 4, 64), movers=6, curtain_res=6)
    nxz, ny, nr = 8, 4, 64
    p = make_params(nxz, ny, nr, flags=FLAGS)
    o, g = oracle.Oracle(p), make_gpu(p)
    parity.setup_pair(scene, (o, g), seed=4)
    rebuilds, sizes = _walk(scene, o, g, ny, 40, 9.0)
    assert rebuilds >= 5 and len(sizes) >= 4, "the walk must actually change the instance set"
    _equal(o, g, ny)
    o.close()
    g.close()
