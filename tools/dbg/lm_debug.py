import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import anerend_b200 as A
from anerend_b200 import layouts as L, make_params, scenes
from oracle import binding as ob
from tests import parity
from tests.test_light_maps import CASES, FLAGS
for name in ("soup_12_lights", "soup_dense_24_lights"):
    scene, grid = CASES[name]()
    nxz, ny, nr = grid
    for extra, tag in ((0, "maps"), (L.FLAG_NO_LIGHT_MAPS, "nomaps")):
        o, g = ob.Oracle(make_params(nxz, ny, nr, flags=FLAGS)), A.Ddgi(make_params(nxz, ny, nr, flags=FLAGS | extra))
        rots = parity.setup_pair(scene, (o, g), seed=7)
        for h in (o, g):
            h.update(-1, rots)
        for layer in range(ny):
            ro, rg = o.download_ray_records(layer), g.download_ray_records(layer)
            bad = (ro["radiance"].view(np.uint32) != rg["radiance"].view(np.uint32)).any(axis=-1)
            print(name, tag, "layer", layer, "prim mism", int((ro["prim"] != rg["prim"]).sum()), "lit mism", int((ro["lit_mask"] != rg["lit_mask"]).sum()), "rad mism", int(bad.sum()))
            idx = np.argwhere(bad)[:5]
            for i in idx:
                a, b = ro[tuple(i)], rg[tuple(i)]
                print("   ", tuple(i), "prim", a["prim"], "lit %x %x" % (a["lit_mask"], b["lit_mask"]), a["radiance"], b["radiance"], (b["radiance"] - a["radiance"]) / np.maximum(np.abs(a["radiance"]), 1e-9))
        print(name, tag, "stats", g.stats())
        o.close(); g.close()
