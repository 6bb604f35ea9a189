"""Synthetic camera + depth buffer for the ShadowRT tests."""
import numpy as np

from anerend_b200 import layouts as L


def ortho_inv_view_proj(half_x=6.0, half_z=6.0, height=10.0):
    """clip (x, y, depth) -> world (half_x * x, height * depth, half_z * y): a top-down view, w = 1.  Column-major, like glm."""
    m = np.zeros((4, 4), np.float64)           # math convention world = m @ clip
    m[0, 0], m[2, 1], m[1, 2], m[3, 3] = half_x, half_z, height, 1.0
    return m.T.astype(np.float32).ravel()


def perspective_inv_view_proj(seed=1):
    """A well-conditioned projective matrix with w != 1 (exercises the perspective divide)."""
    rng = np.random.RandomState(seed)
    m = np.eye(4) + 0.15 * rng.uniform(-1, 1, (4, 4))
    m[:3, :3] *= 5.0
    m[:3, 3] = (0.3, 4.0, -0.2)
    m[3, :3] = 0.08 * rng.uniform(-1, 1, 3)
    return m.T.astype(np.float32).ravel()


def with_shadow_camera(ubo, inv_view_proj, width, height, rt_shadows=True):
    u = ubo.copy()
    u["inv_view_proj"][0] = inv_view_proj
    u["screen_width"], u["screen_height"] = width, height
    if rt_shadows:
        u["flags"] |= L.UBO_RT_SHADOWS_FLAG
    else:
        u["flags"] &= ~np.uint32(L.UBO_RT_SHADOWS_FLAG)
    return u


def depth_buffer(width, height, seed, sky_fraction=0.1):
    rng = np.random.RandomState(seed)
    d = rng.uniform(0.02, 0.98, (height, width)).astype(np.float32)
    d[rng.rand(height, width) < sky_fraction] = 1.0
    return d


def gbuffer_normals(width, height, seed):
    """geom0 (rgba32f): un-normalised normals in rgb (the shader normalises), albedo.r in a."""
    rng = np.random.RandomState(seed)
    g = np.zeros((height, width, 4), np.float32)
    n = rng.normal(size=(height, width, 3))
    g[..., :3] = n * rng.uniform(0.5, 2.0, (height, width, 1))
    g[..., 3] = rng.uniform(0, 1, (height, width))
    return g


def with_specular_camera(ubo, inv_view_proj, width, height, enabled=True):
    u = with_shadow_camera(ubo, inv_view_proj, width, height, rt_shadows=False)
    if enabled:
        u["flags"] |= L.UBO_SPECULAR_GI_FLAG
    else:
        u["flags"] &= ~np.uint32(L.UBO_SPECULAR_GI_FLAG)
    return u
