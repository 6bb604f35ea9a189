#!/usr/bin/env python
"""Fixture from the reference's OWN HOST CODE compiled here (oracle/_ref/libanerend_ref_host.so: render/GpuBuffers.h, render/Vertex.h
and contrib/glm built from /root/reference by oracle/Makefile): the RenderOptions defaults, the byte layout of every struct behind the bindless buffers and the
scene UBO, the SceneUboFlags the path reads, and glm::rotate(mat4(1), radians(angle), axis) — the push-constant matrix of
IrradianceProbeRayTracingPass.cpp:115-133 — for a sweep of axes and angles.  Written to tests/golden/ref_host_golden.json so the
tests can hold the product to it where the reference is not mounted (the GPU box).

    python tests/golden/make_ref_host_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def rotation_cases():
    rng = np.random.RandomState(23)
    angles = [0.0, 90.0, 180.0, 270.0, 359.99, 45.0, 1e-3] + [float(np.float32(a)) for a in rng.uniform(0.0, 360.0, 57)]
    return [(ax, a) for ax in range(3) for a in angles]


def main():
    from oracle import binding as ob
    fields, flags = ob.ref_layout()
    rots = [{"axis": ax, "angle_deg": a, "bits": [int(x) for x in ob.ref_make_rotation(ax, a).view(np.uint32)]} for ax, a in rotation_cases()]
    dst = os.path.join(HERE, "ref_host_golden.json")
    json.dump({"source": "src/anerend/render/GpuBuffers.h, src/anerend/render/Vertex.h, contrib/glm (glm::rotate), compiled by oracle/Makefile",
               "render_option_defaults": [float(x) for x in ob.ref_render_option_defaults()],
               "layout": {k: list(v) for k, v in fields.items()}, "flags": flags, "rotations": rots}, open(dst, "w"), indent=0)
    print("wrote", dst, len(fields), "layout rows,", len(rots), "rotations")


if __name__ == "__main__":
    main()
