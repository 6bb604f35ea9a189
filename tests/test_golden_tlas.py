"""Instance selection against the REFERENCE'S OWN COMPILED SHADER.

tests/golden/tlas_golden.npz holds the VkAccelerationStructureInstanceKHR buffer that assets/tlas_update_comp.spv (the binary the
"UpdateTLAS" pass dispatches) writes for a renderable buffer mixing visible / hidden / culled renderables, multi-mesh models, dynamic
mesh offsets and scaled + rotated + translated transforms — executed with the small SPIR-V interpreter tests/golden/spirv_interp.py,
invocations in renderable order (generator: tests/golden/make_tlas_golden.py, run in the build container where the reference is
mounted).  Held to it: which (renderable, mesh) pairs enter the acceleration structure and in which order, the 24-bit custom index
(material | mesh << 10) the hit shader decodes, and the object-to-world matrix, through the world-space triangles it produces."""
import os

import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tlas_golden.npz"))


def _scene():
    rend = G["renderables"].view(L.RENDERABLE).copy()
    meshes = G["meshes"].view(L.MESH_INFO).copy()
    tri_counts = G["mesh_tri_counts"]
    rng = np.random.RandomState(3)
    verts = np.zeros(16 * len(meshes), L.VERTEX)
    verts["pos"] = rng.uniform(-1, 1, (len(verts), 3)).astype(np.float32)
    verts["normal"][:, 1] = 1.0
    indices = np.concatenate([rng.randint(0, 16, 3 * int(n)) for n in tri_counts]).astype(np.uint32)
    assert np.array_equal(meshes["index_offset"], np.concatenate([[0], np.cumsum(tri_counts * 3)[:-1]]))
    mats = np.zeros(7, L.MATERIAL)
    mats["base_col_fac"] = (0.5, 0.5, 0.5, 1.0); mats["bindless_indices"] = -1; mats["met_rough"][:, 1] = 0.5
    return scenes.Scene("tlas_golden", verts, indices, meshes, (tri_counts * 3).astype(np.uint32), mats, rend, G["model_indices"].copy(),
                        G["material_indices"].copy(), np.zeros(1, L.LIGHT), scenes.make_ubo((0.0, 1.0, 0.0)), (4, 2, 64))


def _check(h, scene, world_triangles=None):
    custom, n = G["custom_index_and_mask"], len(G["custom_index_and_mask"])
    assert h.bvh_info().num_instances == n == 27
    inst = h.download_instances()
    assert np.array_equal(inst[:, 1], custom & 0xFFFFFF)                     # instanceCustomIndex: tlas_update.comp:88-91
    assert np.all(custom >> 24 == 0xFF) and np.all(G["sbt_offset_and_flags"] == 1 << 24)
    mesh = (inst[:, 1] >> 10) & 0x3FFF
    assert np.array_equal(G["blas_ref"], scene.meshes["blas_ref"][mesh])     # the BLAS the binary references is the custom index's mesh
    assert np.array_equal(inst[:, 3], G["mesh_tri_counts"][mesh])
    assert np.array_equal(inst[:, 2], np.concatenate([[0], np.cumsum(inst[:, 3])[:-1]]))
    # renderables in buffer order, each renderable's meshes in order; hidden (bit 0 clear) and culled (bit 1 set) ones absent
    vis = scene.renderables["visible"]
    want_rend = np.concatenate([[i] * int(r["num_meshes"]) for i, r in enumerate(scene.renderables) if (vis[i] & 1) and not (vis[i] >> 1) & 1])
    assert np.array_equal(inst[:, 0], want_rend)
    # the 3x4 row-major matrix the binary wrote is the object-to-world transform of the instance
    for k in range(n):
        assert np.array_equal(G["matrix"][k], scene.renderables["transform"][inst[k, 0]].reshape(4, 4).T[:3])
    if world_triangles is not None:
        f32 = np.float32
        for k in range(n):
            M, mi = G["matrix"][k], scene.meshes[mesh[k]]
            for t in range(int(inst[k, 3])):
                for c in range(3):
                    p = scene.vertices["pos"][int(mi["vertex_offset"]) + int(scene.indices[int(mi["index_offset"]) + 3 * t + c])]
                    w = [f32(f32(f32(f32(M[r, 0] * p[0]) + f32(M[r, 1] * p[1])) + f32(M[r, 2] * p[2])) + M[r, 3]) for r in range(3)]
                    got = world_triangles[int(inst[k, 2]) + t, c]
                    if c == 0:
                        assert got.tolist() == w             # v0 is stored as is; the export rebuilds v1, v2 as v0 + edge
                    else:
                        np.testing.assert_allclose(got, w, rtol=0, atol=1e-6)


def test_oracle_instances_equal_the_compiled_tlas_update_shader(oracle):
    s = _scene()
    o = oracle.Oracle(make_params(4, 2, 64))
    o.upload_scene(s); o.set_scene_ubo(s.ubo); o.build_accel()
    _check(o, s, o.world_triangles())
    o.close()


def test_product_instance_selection_equals_the_compiled_tlas_update_shader(make_emu):
    # selectInstances of csrc/ddgi_api.cu (host code of the product) through the host harness
    s = _scene()
    e = make_emu(make_params(4, 2, 64))
    e.upload_scene(s); e.set_scene_ubo(s.ubo); e.build_accel()
    _check(e, s)
    e.close()


@pytest.mark.gpu
def test_cuda_instance_selection_equals_the_compiled_tlas_update_shader():
    from anerend_b200 import Ddgi
    s = _scene()
    with Ddgi(make_params(4, 2, 64)) as d:
        d.upload_scene(s); d.set_scene_ubo(s.ubo); d.build_accel()
        _check(d, s)
