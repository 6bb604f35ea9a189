"""PROTOTYPE for the next round, on the oracle only: the sun's shadow rays answered from a light-space grid of
triangle lists instead of the hierarchy.  The grid only CULLS — the candidates of a ray's cell go through the same exact triangle test —
so hit records (lit masks), radiance and the stored atlas must not change by a bit, while the work per sun ray drops from a tree walk to
one short list.  This is the property a CUDA version has to keep; the counters show what it buys."""
import numpy as np
import pytest

from anerend_b200 import layouts as L, make_params, scenes
from tests import parity

FLAGS = L.FLAG_RAY_RECORDS | L.FLAG_COUNTERS | L.FLAG_SKIP_DEAD_SHADOW_RAYS          # sun only (the reference's light): every shadow ray is a sun ray

SCENES = {
    "cornell": lambda: (scenes.cornell_box(grid=(8, 4, 64), wall_res=10, box_res=3), (8, 4, 64)),
    "soup": lambda: (scenes.random_soup(11), (4, 3, 64)),
    "textured": lambda: (scenes.textured_soup(51), (4, 3, 64)),
    "adversarial": lambda: (__import__("tests.test_adversarial_geometry", fromlist=["x"]).adversarial_scene(), (4, 3, 64)),
}


@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("res", [16, 256])
def test_sun_grid_changes_nothing_but_the_work(name, res, oracle):
    scene, (nxz, ny, nr) = SCENES[name]()
    p = make_params(nxz, ny, nr, flags=FLAGS)
    a, b = oracle.Oracle(p), oracle.Oracle(p)
    b.set_sun_grid(res)
    rots = parity.setup_pair(scene, (a, b), seed=3)
    for it in range(2):
        for h in (a, b):
            h.update(-1, rots)
        for layer in range(ny):
            ra, rb = a.download_ray_records(layer), b.download_ray_records(layer)
            assert np.array_equal(ra["lit_mask"], rb["lit_mask"]) and np.array_equal(ra["prim"], rb["prim"])
            assert np.array_equal(ra["radiance"].view(np.uint32), rb["radiance"].view(np.uint32))
        assert np.array_equal(a.download_atlas(), b.download_atlas())
    sa, sb = a.stats(), b.stats()
    assert sa["shadow_rays"] == sb["shadow_rays"] > 0
    assert sb["shadow_nodes_visited"] == sb["shadow_rays"]                     # one cell per sun ray instead of a tree walk
    a.close(); b.close()


def test_sun_grid_follows_the_sun_and_moving_geometry(oracle):
    scene, (nxz, ny, nr) = SCENES["soup"]()
    p = make_params(nxz, ny, nr, flags=FLAGS)
    a, b = oracle.Oracle(p), oracle.Oracle(p)
    b.set_sun_grid(128)
    rots = parity.setup_pair(scene, (a, b), seed=5)
    rng = np.random.RandomState(5)
    for step in range(3):
        ubo = scene.ubo.copy()
        d = rng.normal(size=3); d[1] = -abs(d[1]) - 0.2
        ubo["light_dir"][0][:3] = d / np.linalg.norm(d)                        # a new sun direction: the grid is rebuilt in its space
        rend = scene.renderables.copy()
        for i in range(len(rend)):
            M = rend["transform"][i].reshape(4, 4).T.astype(np.float64)
            rend["transform"][i] = (scenes.translate(rng.uniform(-1, 1, 3)) @ M).T.astype(np.float32).ravel()
        for h in (a, b):
            h.set_scene_ubo(ubo); h.refit_accel(rend); h.update(-1, rots)
        for layer in range(ny):
            ra, rb = a.download_ray_records(layer), b.download_ray_records(layer)
            assert np.array_equal(ra["lit_mask"], rb["lit_mask"])
        assert np.array_equal(a.download_atlas(), b.download_atlas())
    a.close(); b.close()
