"""Synthetic G-buffer for the atlas-consumer tests: surface points in and around the probe volume."""
import numpy as np


def surface_points(n, nxz, ny, seed, camera=(0.0, 0.0, 0.0), step=(1.0, 2.0, 1.0)):
    rng = np.random.RandomState(seed)
    pts = np.zeros((n, 12), np.float32)
    half = 0.5 * nxz * step[0]
    # 80 % inside the grid, 20 % up to two cells outside (probe index clamping, out-of-range taps: REPEAT vs CLAMP_TO_EDGE differ there)
    lo = np.array([camera[0] - half - 2.0, -2.0, camera[2] - half - 2.0]); hi = np.array([camera[0] + half + 2.0, ny * step[1] + 2.0, camera[2] + half + 2.0])
    pts[:, 0:3] = rng.uniform(lo, hi, (n, 3))
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[: n // 16] = np.eye(3)[rng.randint(0, 3, n // 16)] * rng.choice([-1.0, 1.0], (n // 16, 1))      # axis-aligned normals (octahedral folds)
    pts[:, 3:6] = nrm
    pts[:, 6:9] = rng.uniform(0.05, 0.95, (n, 3))
    pts[:, 9] = (rng.rand(n) < 0.2).astype(np.float32)
    pts[:, 10] = rng.uniform(0.1, 1.0, n)
    return pts


def synthetic_baked_atlas(width, height, seed):
    """A deterministic stand-in for a baked tile atlas: RGBA16F bit patterns [height, width, 4] (uint16), radiance in [0, 2), alpha = weight sum."""
    rng = np.random.RandomState(seed)
    a = rng.uniform(0.0, 2.0, (height, width, 4)).astype(np.float16)
    a[..., 3] = np.float16(64.0)
    return a.view(np.uint16)
