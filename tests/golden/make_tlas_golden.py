#!/usr/bin/env python
"""Golden vectors for instance selection from the reference's OWN compiled shader: executes assets/tlas_update_comp.spv
(the "UpdateTLAS" compute pass that fills the VkAccelerationStructureInstanceKHR buffer the TLAS is built from) with
tests/golden/spirv_interp.py, one invocation per renderable in buffer order, over a renderable buffer that mixes visible, hidden
and culled renderables, multi-mesh models, dynamic (skinned) mesh offsets and non-trivial transforms, and writes the instance
buffer the binary produces to tests/golden/tlas_golden.npz together with the input buffers.

    python tests/golden/make_tlas_golden.py [/root/reference/assets/tlas_update_comp.spv]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from spirv_interp import Invocation, Module, f32  # noqa: E402


def make_inputs(seed=11, num_renderables=24, num_meshes=9, num_materials=7):
    """Random renderable / model-index / material-index / mesh buffers in the reference's layouts."""
    from anerend_b200 import layouts as L
    rng = np.random.RandomState(seed)
    meshes = np.zeros(num_meshes, L.MESH_INFO)
    tri_counts = rng.randint(1, 6, num_meshes)
    meshes["vertex_offset"] = np.arange(num_meshes) * 16
    meshes["index_offset"] = np.concatenate([[0], np.cumsum(tri_counts * 3)[:-1]])
    meshes["blas_ref"] = (0xABCD00000000 + np.arange(num_meshes) * 0x100).astype(np.uint64)
    meshes["min_pos"][:, :3], meshes["max_pos"][:, :3] = -1.0, 1.0
    model_indices, material_indices = [0], [0]             # slot 0 unused, so that a dynamicModelOffset of 0 means "none"
    rend = np.zeros(num_renderables, L.RENDERABLE)
    for i in range(num_renderables):
        n = int(rng.randint(1, 4))
        rend["model_offset"][i] = len(model_indices)
        rend["num_meshes"][i] = n
        model_indices += [int(x) for x in rng.randint(0, num_meshes, n)]
        rend["first_material_index"][i] = len(material_indices)
        material_indices += [int(x) for x in rng.randint(0, num_materials, n)]
        if i % 5 == 2:                                     # a skinned renderable: the TLAS takes its dynamic mesh copies
            rend["dynamic_model_offset"][i] = len(model_indices)
            model_indices += [int(x) for x in rng.randint(0, num_meshes, n)]
        # bit 0 visible, bit 1 culled (bindless.glsl:249-263); other bits are noise the shader must ignore
        rend["visible"][i] = (0, 1, 1, 3, 1, 2, 1, 1 | 4 | 8)[i % 8]
        t = np.eye(4, dtype=np.float32)
        a = rng.uniform(0, 2 * np.pi)
        t[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32) * np.float32(rng.uniform(0.5, 2.0))
        t[:3, 3] = rng.uniform(-10, 10, 3)
        rend["transform"][i] = t.T.ravel()                 # column-major mat4
        rend["terrain_offset"][i] = -1
    return rend, np.array(model_indices, np.uint32), np.array(material_indices, np.uint32), meshes, tri_counts.astype(np.uint32)


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/tlas_update_comp.spv"
    m = Module(path)
    rend, model_indices, material_indices, meshes, tri_counts = make_inputs()
    renderables = [[[[f32(r["transform"][4 * c + k]) for k in range(4)] for c in range(4)], [f32(x) for x in r["bounds"]], [f32(x) for x in r["tint"]],
                    int(r["model_offset"]), int(r["num_meshes"]), int(r["skeleton_offset"]), int(r["visible"]), int(r["first_material_index"]),
                    int(r["dynamic_model_offset"]), int(r["terrain_offset"])] for r in rend]
    mesh_list = [[[f32(x) for x in mm["min_pos"]], [f32(x) for x in mm["max_pos"]], int(mm["vertex_offset"]), int(mm["index_offset"]), int(mm["blas_ref"])] for mm in meshes]
    inst_gid = next(g for g, n in m.names.items() if n == "instanceBuffer" and g in m.globals)
    inst_t = m.types[m.types[m.types[m.globals[inst_gid][0]][2]][1][0]][1]              # block -> runtime array -> AccelerationStructureInstance
    total = int(rend["num_meshes"].sum())
    instances = [m.zero(inst_t) for _ in range(total)]
    counter = [0]
    # two extra invocations past maxCount: the bounds check (tlas_update.comp:41-43) must drop them
    for idx in range(len(rend) + 2):
        inv = Invocation(m, {"gl_GlobalInvocationID": [idx, 0, 0], "pushConstants": [len(rend)], "renderableBuffer": [renderables],
                             "rendModelBuffer": [[int(x) for x in model_indices]], "rendMatIndexBuffer": [[int(x) for x in material_indices]],
                             "meshBuffer": [mesh_list], "counterBuffer": [counter], "instanceBuffer": [instances]})
        inv.run()
    n = counter[0]
    print("instances written:", n, "of", total, "candidate (renderable, mesh) pairs")
    out_matrix = np.array([[[float(c) for c in row] for row in inst[0][0]] for inst in instances[:n]], np.float32)        # [n, 3, 4] row major
    out_custom = np.array([inst[1] for inst in instances[:n]], np.uint32)
    out_sbt = np.array([inst[2] for inst in instances[:n]], np.uint32)
    out_blas = np.array([inst[3] for inst in instances[:n]], np.uint64)
    dst = os.path.join(HERE, "tlas_golden.npz")
    np.savez_compressed(dst, renderables=rend.view(np.uint8), model_indices=model_indices, material_indices=material_indices,
                        meshes=meshes.view(np.uint8), mesh_tri_counts=tri_counts,
                        matrix=out_matrix, custom_index_and_mask=out_custom, sbt_offset_and_flags=out_sbt, blas_ref=out_blas)
    print("wrote", dst)


if __name__ == "__main__":
    main()
