"""Synthetic scenes in the reference's bindless buffer formats (BASELINE.json configs).

The reference loads glTF assets through an editor; here every scene is generated in-process with fixed
seeds, and emitted as exactly the buffers `VulkanRenderer` would upload: 96-byte vertices
(Vertex.h:7-19), u32 indices, GPUMeshInfo / GPUMaterialInfo / GPURenderable / GPULight arrays
(GpuBuffers.h:15-72), the renderable->mesh and renderable->material index tables
(VulkanRenderer.cpp:3157-3360) and the 568-byte scene UBO (GpuBuffers.h:125-148).
"""
from dataclasses import dataclass, field

import numpy as np

from . import layouts as L


# --------------------------------------------------------------------------------------------------
# mesh primitives: (positions [n,3] f32, normals [n,3] f32, indices [m] u32)
# --------------------------------------------------------------------------------------------------
def _grid_indices(nu, nv):
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).ravel()
    b, c, d = a + (nv + 1), a + (nv + 1) + 1, a + 1
    return np.stack([a, b, c, a, c, d], 1).astype(np.uint32).ravel()


def grid_plane(nu, nv):
    """Unit square [0,1]^2 in the XZ plane, normal +Y, 2*nu*nv triangles."""
    u, v = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="ij")
    pos = np.stack([u.ravel(), np.zeros(u.size), v.ravel()], 1).astype(np.float32)
    nrm = np.tile(np.array([0, 1, 0], np.float32), (len(pos), 1))
    return pos, nrm, _grid_indices(nu, nv)


def heightfield(nu, nv, fn):
    """[0,1]^2 in XZ with y = fn(x, z); normals from central differences."""
    u, v = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="ij")
    y = fn(u, v)
    e = 1e-3
    dydu = (fn(u + e, v) - fn(u - e, v)) / (2 * e)
    dydv = (fn(u, v + e) - fn(u, v - e)) / (2 * e)
    n = np.stack([-dydu, np.ones_like(y), -dydv], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    pos = np.stack([u.ravel(), y.ravel(), v.ravel()], 1).astype(np.float32)
    return pos, n.reshape(-1, 3).astype(np.float32), _grid_indices(nu, nv)


def box(n):
    """Cube [-0.5,0.5]^3, each face an n x n grid with outward normals: 12*n*n triangles."""
    P, N, I = [], [], []
    base = 0
    axes = [(0, 1, 2), (1, 2, 0), (2, 0, 1)]
    for ax, (a, b, c) in enumerate(axes):
        for sgn in (-1.0, 1.0):
            u, v = np.meshgrid(np.linspace(-0.5, 0.5, n + 1), np.linspace(-0.5, 0.5, n + 1), indexing="ij")
            p = np.zeros((u.size, 3), np.float32)
            p[:, a] = u.ravel()
            p[:, b] = v.ravel()
            p[:, c] = 0.5 * sgn
            nn = np.zeros_like(p)
            nn[:, c] = sgn
            P.append(p)
            N.append(nn)
            I.append(_grid_indices(n, n) + base)
            base += len(p)
    return np.concatenate(P), np.concatenate(N), np.concatenate(I).astype(np.uint32)


def cylinder(nseg, nring, radius=0.5):
    """Open cylinder around +Y, y in [0,1]: 2*nseg*nring triangles."""
    th, y = np.meshgrid(np.linspace(0, 2 * np.pi, nseg + 1), np.linspace(0, 1, nring + 1), indexing="ij")
    n = np.stack([np.cos(th), np.zeros_like(th), np.sin(th)], -1).reshape(-1, 3)
    pos = n * radius
    pos[:, 1] = y.ravel()
    return pos.astype(np.float32), n.astype(np.float32), _grid_indices(nseg, nring)


def sphere(nseg, nring, radius=0.5):
    """Lat-long sphere: 2*nseg*nring triangles (pole triangles are degenerate and never hit)."""
    th, ph = np.meshgrid(np.linspace(0, 2 * np.pi, nseg + 1), np.linspace(0, np.pi, nring + 1), indexing="ij")
    n = np.stack([np.cos(th) * np.sin(ph), np.cos(ph), np.sin(th) * np.sin(ph)], -1).reshape(-1, 3)
    return (n * radius).astype(np.float32), n.astype(np.float32), _grid_indices(nseg, nring)


def torus(nmaj, nmin, R=0.5, r=0.12, arc=2 * np.pi):
    """(Partial) torus in the XY plane around Z: 2*nmaj*nmin triangles. arc=pi gives an arch."""
    a, b = np.meshgrid(np.linspace(0, arc, nmaj + 1), np.linspace(0, 2 * np.pi, nmin + 1), indexing="ij")
    cx, cy = np.cos(a), np.sin(a)
    n = np.stack([cx * np.cos(b), cy * np.cos(b), np.sin(b)], -1).reshape(-1, 3)
    c = np.stack([cx * R, cy * R, np.zeros_like(a)], -1).reshape(-1, 3)
    return (c + n * r).astype(np.float32), n.astype(np.float32), _grid_indices(nmaj, nmin)


# --------------------------------------------------------------------------------------------------
# transforms (math convention: 4x4, p' = M @ [p,1]); stored column-major like glm::mat4
# --------------------------------------------------------------------------------------------------
def translate(t):
    m = np.eye(4)
    m[:3, 3] = t
    return m


def scale(s):
    return np.diag(list(np.broadcast_to(np.asarray(s, float), (3,))) + [1.0])


def rotate(axis, deg):
    a = np.radians(deg)
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    m[i, i], m[i, j], m[j, i], m[j, j] = c, -s, s, c
    return m


def frame(origin, ux, uy, uz):
    """Maps unit axes to the given (possibly scaled) basis vectors."""
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = ux, uy, uz, origin
    return m


# --------------------------------------------------------------------------------------------------
@dataclass
class Scene:
    name: str
    vertices: np.ndarray
    indices: np.ndarray
    meshes: np.ndarray
    mesh_num_indices: np.ndarray
    materials: np.ndarray
    renderables: np.ndarray
    model_indices: np.ndarray
    material_indices: np.ndarray
    lights: np.ndarray
    ubo: np.ndarray
    grid: tuple = (16, 8, 16)         # nXZ, nY (as (nXZ, nY, rays))
    movers: list = field(default_factory=list)   # [(renderable index, base 4x4, centre, amplitude, omega, phase)]
    seed: int = 0
    textures: list = field(default_factory=list)   # [(pixels uint8 [h, w, c], TEX_* format, clamp_to_edge)], index = bindless index

    @property
    def num_triangles(self):
        """Triangles of all visible instances (what tlas_update.comp would put in the TLAS)."""
        total = 0
        for r in self.renderables:
            if (r["visible"] & 1) and not (r["visible"] >> 1) & 1:
                for i in range(r["num_meshes"]):
                    total += int(self.mesh_num_indices[self.model_indices[r["model_offset"] + i]]) // 3
        return total

    def animated_renderables(self, t):
        """Renderable buffer at simulation time t (seconds): rigid movers on sinusoidal paths (config C4)."""
        out = self.renderables.copy()
        for (ri, base, centre, amp, omega, phase) in self.movers:
            p = centre + amp * np.sin(omega * t + phase)
            m = translate(p) @ rotate(1, 40.0 * t + 57.0 * phase[0]) @ base
            out[ri]["transform"] = m.T.astype(np.float32).ravel()
        return out

    def with_camera(self, cam):
        ubo = self.ubo.copy()
        ubo["camera_pos"][0] = (cam[0], cam[1], cam[2], 1.0)
        ubo["camera_grid_pos"][0] = (int(np.floor(cam[0])), int(np.floor(cam[1])), int(np.floor(cam[2])), 1)
        return ubo


class SceneBuilder:
    def __init__(self, name, seed):
        self.name, self.seed = name, seed
        self.rng = np.random.RandomState(seed)
        self.v, self.i, self.meshes, self.mni = [], [], [], []
        self.nv = self.ni = 0
        self.mats, self.rends, self.model_idx, self.mat_idx, self.lights, self.movers = [], [], [], [], [], []
        self.textures = []

    def add_texture(self, pixels, fmt, clamp_to_edge=False):
        self.textures.append((np.ascontiguousarray(pixels, np.uint8), fmt, clamp_to_edge))
        return len(self.textures) - 1

    def add_mesh(self, pos, nrm, idx, color=(1.0, 1.0, 1.0), uv=None):
        v = np.zeros(len(pos), L.VERTEX)
        v["pos"], v["normal"], v["color"] = pos, nrm, color     # GLTFLoader.cpp:343-366 default colour = 1,1,1
        if uv is not None:
            v["uv"] = uv
        m = np.zeros(1, L.MESH_INFO)
        m["min_pos"][0, :3], m["max_pos"][0, :3] = pos.min(0), pos.max(0)
        m["vertex_offset"], m["index_offset"] = self.nv, self.ni
        self.v.append(v)
        self.i.append(idx.astype(np.uint32))
        self.meshes.append(m)
        self.mni.append(len(idx))
        self.nv += len(pos)
        self.ni += len(idx)
        return len(self.meshes) - 1

    def add_material(self, base_col, metallic=0.0, roughness=1.0, emissive=(0, 0, 0, 0), tex=(-1, -1, -1, -1)):
        """tex = bindless indices (metallicRoughness, albedo, normal, emissive), -1 = none (GpuBuffers.h:29-36)."""
        m = np.zeros(1, L.MATERIAL)
        m["base_col_fac"][0] = (*base_col, 1.0)
        m["emissive"][0] = emissive
        m["bindless_indices"][0] = tex
        m["met_rough"][0, :2] = (metallic, roughness)
        self.mats.append(m)
        return len(self.mats) - 1

    def add_renderable(self, M, mesh_ids, mat_ids, visible=1):
        mesh_ids, mat_ids = list(np.atleast_1d(mesh_ids)), list(np.atleast_1d(mat_ids))
        r = np.zeros(1, L.RENDERABLE)
        r["transform"][0] = np.asarray(M, np.float64).T.astype(np.float32).ravel()
        lo = np.min([self.meshes[m]["min_pos"][0, :3] for m in mesh_ids], 0)
        hi = np.max([self.meshes[m]["max_pos"][0, :3] for m in mesh_ids], 0)
        r["bounds"][0] = (*((lo + hi) / 2), float(np.linalg.norm(hi - lo) / 2))
        r["tint"][0] = 1.0
        r["model_offset"], r["num_meshes"] = len(self.model_idx), len(mesh_ids)
        r["first_material_index"] = len(self.mat_idx)
        r["visible"], r["terrain_offset"] = visible, -1
        self.model_idx += mesh_ids
        self.mat_idx += mat_ids
        self.rends.append(r)
        return len(self.rends) - 1

    def add_light(self, pos, rng_m, color, enabled=True):
        l = np.zeros(1, L.LIGHT)
        l["world_pos"][0] = (*pos, rng_m)
        l["color"][0] = (*color, 1.0 if enabled else 0.0)
        self.lights.append(l)

    def add_mover(self, rend, base, centre, amp, omega, phase):
        self.movers.append((rend, np.asarray(base, float), np.asarray(centre, float), np.asarray(amp, float),
                            np.asarray(omega, float), np.asarray(phase, float)))

    def finish(self, camera, grid, multi_bounce=True):
        ubo = make_ubo(camera, multi_bounce=multi_bounce)
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
        return Scene(self.name, cat(self.v, L.VERTEX), cat(self.i, np.uint32), cat(self.meshes, L.MESH_INFO),
                     np.asarray(self.mni, np.uint32), cat(self.mats, L.MATERIAL), cat(self.rends, L.RENDERABLE),
                     np.asarray(self.model_idx, np.uint32), np.asarray(self.mat_idx, np.uint32),
                     cat(self.lights, L.LIGHT), ubo, grid, self.movers, self.seed, list(self.textures))


def make_ubo(camera, multi_bounce=True, sun_intensity=5.0, sky_intensity=1.0, sky_color=(0.2, 0.5, 0.7),
             light_dir=(0.7, -0.7, 0.7), ddgi=True, rt_on=True):
    """Defaults: RenderOptions.h:7-28; sun direction AneditApplication.cpp:34; flags VulkanRenderer.cpp:1222-1232."""
    u = np.zeros(1, L.SCENE_UBO)
    eye = np.eye(4, dtype=np.float32).ravel()
    for k in ("view", "proj", "inv_view", "inv_proj", "inv_view_proj", "directional_shadow_matrix_proj",
              "directional_shadow_matrix_view"):
        u[k][0] = eye
    ld = np.asarray(light_dir, np.float64)
    ld = ld / np.linalg.norm(ld)
    u["camera_pos"][0] = (*camera, 1.0)
    u["camera_grid_pos"][0] = (*[int(np.floor(c)) for c in camera], 1)
    u["light_dir"][0] = (*ld, 0.0)
    u["sky_color"][0] = (*sky_color, 1.0)
    u["screen_width"], u["screen_height"] = 1920, 1080
    u["sun_intensity"], u["sky_intensity"], u["exposure"] = sun_intensity, sky_intensity, 1.0
    u["bloom_thresh"], u["bloom_knee"] = 1.2, 0.1
    flags = 0
    if ddgi:
        flags |= L.UBO_DDGI_FLAG
    if multi_bounce:
        flags |= L.UBO_DDGI_MULTI_FLAG
    if rt_on:
        flags |= L.UBO_RT_ON_FLAG
    u["flags"] = flags
    return u


def rotation_sequence(seed, count):
    """(axis, angle_deg) pairs like IrradianceProbeRayTracingPass.cpp:115-133 draws them, from mt19937(seed)
    instead of std::random_device so oracle and CUDA see identical rotations."""
    rng = np.random.RandomState(seed)
    return [(int(rng.randint(0, 3)), float(np.float32(rng.uniform(0.0, 360.0)))) for _ in range(count)]


# --------------------------------------------------------------------------------------------------
# scenes
# --------------------------------------------------------------------------------------------------
def empty_scene(grid=(8, 8, 64), camera=(0.0, 4.0, 0.0)):
    """No geometry at all: every ray misses (KAT 5).  One invisible renderable keeps the buffers non-empty."""
    b = SceneBuilder("empty", 0)
    m = b.add_mesh(*grid_plane(1, 1))
    mat = b.add_material((0.5, 0.5, 0.5))
    b.add_renderable(np.eye(4), m, mat, visible=0)
    return b.finish(camera, grid)


def quad_scene(grid=(4, 2, 64), camera=(0.0, 1.0, 0.0), second_quad=False):
    """KAT 9: one axis-aligned quad above the probes, optionally a second one shadowing half of it."""
    b = SceneBuilder("quad", 0)
    m = b.add_mesh(*grid_plane(1, 1))
    white = b.add_material((0.8, 0.8, 0.8))
    # floor quad 20 m x 20 m at y = -1.5 (normal +y), centred on the grid
    b.add_renderable(frame((-10, -1.5, -10), (20, 0, 0), (0, 1, 0), (0, 0, 20)), m, white)
    if second_quad:
        b.add_renderable(frame((-10, 6.0, -10), (10, 0, 0), (0, 1, 0), (0, 0, 20)), m, white)
    return b.finish(camera, grid)


def cornell_box(seed=2, grid=(16, 8, 128), wall_res=30, box_res=6):
    """BASELINE config C2: 14 m Cornell box (~10k triangles), red/green/white diffuse, two inner boxes,
    sun through the open -z face plus one point light (extension E1)."""
    b = SceneBuilder("cornell", seed)
    plane = b.add_mesh(*grid_plane(wall_res, wall_res))
    cube = b.add_mesh(*box(box_res))
    white = b.add_material((0.73, 0.73, 0.73))
    red = b.add_material((0.65, 0.05, 0.05))
    green = b.add_material((0.12, 0.45, 0.15))
    x0, x1, y0, y1, z0, z1 = -7.4, 6.6, -0.3, 13.7, -7.4, 6.6   # offset so no probe lies exactly on a wall
    S = 14.0
    b.add_renderable(frame((x0, y0, z0), (S, 0, 0), (0, 1, 0), (0, 0, S)), plane, white)      # floor,   normal +y
    b.add_renderable(frame((x0, y1, z0), (S, 0, 0), (0, -1, 0), (0, 0, S)), plane, white)     # ceiling, normal -y
    b.add_renderable(frame((x0, y0, z1), (S, 0, 0), (0, 0, -1), (0, S, 0)), plane, white)     # back,    normal -z
    b.add_renderable(frame((x0, y0, z0), (0, 0, S), (1, 0, 0), (0, S, 0)), plane, red)        # left,    normal +x
    b.add_renderable(frame((x1, y0, z0), (0, 0, S), (-1, 0, 0), (0, S, 0)), plane, green)     # right,   normal -x
    # inner boxes: rotated, non-uniformly scaled (exercises the inverse-transpose normal transform, rchit:73)
    b.add_renderable(translate((-2.6, y0 + 4.0, 1.9)) @ rotate(1, 17.0) @ scale((4.0, 8.0, 4.0)), cube, white)
    b.add_renderable(translate((2.3, y0 + 2.0, -2.2)) @ rotate(1, -18.0) @ scale((4.0, 4.0, 4.0)), cube, white)
    b.add_light((-0.4, 12.2, -0.4), 15.0, (8.0, 7.0, 6.0))
    return b.finish((0.0, 7.0, 0.0), grid)


def _noise(rng, octaves=4):
    ks = [(rng.uniform(1, 4) * 2 ** o, rng.uniform(1, 4) * 2 ** o, rng.uniform(0, 6.28), rng.uniform(0, 6.28), 0.5 ** o)
          for o in range(octaves)]

    def fn(u, v):
        y = np.zeros_like(u)
        for ku, kv, pu, pv, a in ks:
            y = y + a * np.sin(ku * u * 6.28 + pu) * np.cos(kv * v * 6.28 + pv)
        return y
    return fn


def sponza_like(seed=3, grid=(24, 12, 256), movers=0, curtain_res=53):
    """BASELINE config C3 (and C4 with movers=16): ~262k triangles inside Sponza's world AABB
    x∈[-15.4,14.4], y∈[-1.0,11.4], z∈[-9.5,8.8] (assets/models/sponza-gltf-pbr/sponza.glb, node scale 0.008):
    floor, outer walls, aisle roofs and gallery slabs around an open atrium, two colonnade rows on two
    storeys with arches, hanging curtains (noisy quads), props; 64 factor-only materials, 8 point lights."""
    b = SceneBuilder("sponza_like" if not movers else "sponza_scroll", seed)
    rng = b.rng
    mats = []
    for k in range(64):
        emissive = (0, 0, 0, 0)
        if k % 16 == 7:                                     # 4 emissive materials
            emissive = (*rng.uniform(0.5, 1.0, 3), rng.uniform(1.0, 3.0))
        mats.append(b.add_material(rng.uniform(0.1, 0.9, 3), float(rng.rand() < 0.1), rng.uniform(0.2, 1.0), emissive))
    pick = lambda: mats[rng.randint(0, 64)]
    X0, X1, Y0, Y1, Z0, Z1 = -15.4, 14.4, -1.0, 11.4, -9.5, 8.8
    fy = Y0 + 0.05
    floor = b.add_mesh(*grid_plane(128, 96))
    wall_l = b.add_mesh(*grid_plane(64, 32))
    strip = b.add_mesh(*grid_plane(64, 16))
    col = b.add_mesh(*cylinder(32, 24))
    arch = b.add_mesh(*torus(24, 12, arc=np.pi))
    ball = b.add_mesh(*sphere(32, 16))
    cube = b.add_mesh(*box(9))
    b.add_renderable(frame((X0, fy, Z0), (X1 - X0, 0, 0), (0, 1, 0), (0, 0, Z1 - Z0)), floor, pick())
    Hh = Y1 - fy
    b.add_renderable(frame((X0, fy, Z0), (X1 - X0, 0, 0), (0, 0, 1), (0, Hh, 0)), wall_l, pick())     # z = Z0 wall
    b.add_renderable(frame((X0, fy, Z1), (X1 - X0, 0, 0), (0, 0, -1), (0, Hh, 0)), wall_l, pick())    # z = Z1 wall
    b.add_renderable(frame((X0, fy, Z0), (0, 0, Z1 - Z0), (1, 0, 0), (0, Hh, 0)), wall_l, pick())     # x = X0 wall
    b.add_renderable(frame((X1, fy, Z0), (0, 0, Z1 - Z0), (-1, 0, 0), (0, Hh, 0)), wall_l, pick())    # x = X1 wall
    aisle = 4.3
    for (za, zb) in ((Z0, -aisle), (aisle, Z1)):                # roofs and gallery slabs over the side aisles
        b.add_renderable(frame((X0, Y1 - 0.05, za), (X1 - X0, 0, 0), (0, -1, 0), (0, 0, zb - za)), strip, pick())
        b.add_renderable(frame((X0, 5.2, za), (X1 - X0, 0, 0), (0, -1, 0), (0, 0, zb - za)), strip, pick())
    ncol = 14
    xs = np.linspace(X0 + 1.6, X1 - 1.6, ncol)
    for zc in (-aisle, aisle):
        for storey, (ya, yb) in enumerate(((fy, 5.2), (5.2, Y1 - 0.05))):
            cm = pick()
            for k, x in enumerate(xs):
                b.add_renderable(translate((x, ya, zc)) @ scale((0.7, yb - ya - 1.0, 0.7)), col, cm)
                if k + 1 < ncol:                                 # arch spanning to the next column
                    span = xs[k + 1] - x
                    b.add_renderable(translate((x + span / 2, yb - 1.0, zc)) @ scale((span, 1.9, 1.6)), arch, cm)
    ncurt = 12
    curtain_meshes = [b.add_mesh(*heightfield(curtain_res, curtain_res, lambda u, v, f=_noise(rng): 0.06 * f(u, v)))
                      for _ in range(4)]
    for k in range(ncurt):
        x = X0 + 2.5 + (X1 - X0 - 5.0) * (k % 6) / 5.0
        z = (-aisle + 0.45) if k < 6 else (aisle - 0.45)
        # hangs in the XY plane: local x -> world x, local z -> world -y, local y (noise) -> world z
        b.add_renderable(frame((x - 1.1, 10.6, z), (2.2, 0, 0), (0, 0, 3.0), (0, -4.6, 0)), curtain_meshes[k % 4], pick())
    for k in range(32):                                          # props on the floor and the galleries
        p = (rng.uniform(X0 + 1, X1 - 1), fy + 0.45 if k % 2 == 0 else 5.2 + 0.45, rng.uniform(Z0 + 1, Z1 - 1))
        s = rng.uniform(0.6, 0.9)
        b.add_renderable(translate(p) @ rotate(1, rng.uniform(0, 360)) @ scale((s, s * rng.uniform(0.8, 1.4), s)), ball, pick())
    for k in range(8):
        b.add_light((rng.uniform(X0 + 2, X1 - 2), rng.uniform(1.5, 8.0), rng.uniform(-3.5, 3.5)),
                    rng.uniform(8.0, 15.0), tuple(rng.uniform(2.0, 8.0, 3)))
    for k in range(movers):                                      # C4: rigid boxes on sinusoidal paths in the atrium
        centre = np.array([rng.uniform(X0 + 3, X1 - 3), rng.uniform(1.0, 8.0), rng.uniform(-3.0, 3.0)])
        base = scale(rng.uniform(0.5, 1.2, 3))
        r = b.add_renderable(translate(centre) @ base, cube, pick())
        b.add_mover(r, base, centre, rng.uniform(0.3, 1.5, 3), rng.uniform(0.5, 2.5, 3), rng.uniform(0, 6.28, 3))
    return b.finish((0.0, 5.0, 0.0), grid)


def large_volume(seed=5, grid=(64, 32, 256), terrain_res=707, props=500):
    """BASELINE config C5: ~2.0M triangles in a 64 m cube: a 64 m x 64 m terrain (2*707^2 ≈ 1.0M tris) plus
    500 instanced 2k-triangle props scattered through the volume; 8 point lights."""
    b = SceneBuilder("large_volume", seed)
    rng = b.rng
    mats = [b.add_material(rng.uniform(0.1, 0.9, 3), float(rng.rand() < 0.1), rng.uniform(0.2, 1.0),
                           (*rng.uniform(0.5, 1, 3), 2.0) if k % 16 == 7 else (0, 0, 0, 0)) for k in range(64)]
    f = _noise(rng, 5)
    terrain = b.add_mesh(*heightfield(terrain_res, terrain_res, lambda u, v: 0.05 * f(u, v)))
    b.add_renderable(frame((-32.0, 3.0, -32.0), (64.0, 0, 0), (0, 64.0, 0), (0, 0, 64.0)), terrain, mats[0])
    prop_meshes = [b.add_mesh(*sphere(40, 25)), b.add_mesh(*box(13)), b.add_mesh(*cylinder(40, 25)), b.add_mesh(*torus(50, 20))]
    for k in range(props):
        p = (rng.uniform(-30, 30), rng.uniform(6.0, 60.0), rng.uniform(-30, 30))
        s = rng.uniform(0.5, 3.0, 3)
        M = translate(p) @ rotate(1, rng.uniform(0, 360)) @ rotate(0, rng.uniform(0, 360)) @ scale(s)
        b.add_renderable(M, prop_meshes[k % 4], mats[rng.randint(0, 64)])
    for k in range(8):
        b.add_light((rng.uniform(-25, 25), rng.uniform(8.0, 40.0), rng.uniform(-25, 25)), rng.uniform(8.0, 15.0),
                    tuple(rng.uniform(2.0, 8.0, 3)))
    return b.finish((0.0, 30.0, 0.0), grid)


def instanced_field(seed, instances=1024, grid=(16, 8, 128), extent=24.0, seg=(32, 16), nlights=4):
    """Many instances of a few meshes (what the reference's TLAS is for, tlas_update.comp:55-98): a ground plane and `instances`
    rigid props — spheres, tori, cylinders, boxes of ~1-2k triangles each, random rotations, non-uniform scales down to a few
    centimetres — some of them on sinusoidal paths (`movers`).  Flattened, the scene holds instances x mesh triangles."""
    b = SceneBuilder("instanced_field", seed)
    rng = b.rng
    mats = [b.add_material(rng.uniform(0.1, 0.9, 3), float(rng.rand() < 0.2), rng.uniform(0.2, 1.0),
                           (*rng.uniform(0.5, 1, 3), 1.5) if k % 8 == 5 else (0, 0, 0, 0)) for k in range(16)]
    ground = b.add_mesh(*grid_plane(8, 8))
    props = [b.add_mesh(*sphere(seg[0], seg[1])), b.add_mesh(*torus(seg[0], max(4, seg[1] // 2))), b.add_mesh(*cylinder(seg[0], seg[1])),
             b.add_mesh(*box(max(2, seg[1] // 2)))]
    h = extent / 2
    b.add_renderable(frame((-h, 0.0, -h), (extent, 0, 0), (0, 1, 0), (0, 0, extent)), ground, mats[0])
    for k in range(instances):
        p = (rng.uniform(-h + 1, h - 1), rng.uniform(0.3, 0.6 * h), rng.uniform(-h + 1, h - 1))
        s = rng.uniform(0.3, 1.6, 3) * (0.05 if k % 37 == 36 else 1.0)
        M = translate(p) @ rotate(k % 3, rng.uniform(0, 360)) @ rotate((k + 1) % 3, rng.uniform(0, 360)) @ scale(s)
        r = b.add_renderable(M, props[k % 4], mats[rng.randint(0, 16)])
        if k % 16 == 3:
            base = rotate(k % 3, rng.uniform(0, 360)) @ scale(s)
            b.add_mover(r, base, np.asarray(p), rng.uniform(0.2, 1.0, 3), rng.uniform(0.5, 2.5, 3), rng.uniform(0, 6.28, 3))
    for k in range(nlights):
        b.add_light((rng.uniform(-h / 2, h / 2), rng.uniform(2.0, 0.5 * h), rng.uniform(-h / 2, h / 2)), rng.uniform(8.0, 15.0),
                    tuple(rng.uniform(2.0, 8.0, 3)))
    return b.finish((0.0, 3.0, 0.0), grid)


def random_soup(seed, ntris=2000, extent=6.0, grid=(4, 3, 64), camera=(0.3, 2.0, -0.2), nmat=5, nlights=2):
    """Unstructured triangle soup with several meshes/instances, non-white vertex colours and odd transforms:
    a small adversarial scene for parity tests."""
    b = SceneBuilder("soup%d" % seed, seed)
    rng = b.rng
    mats = [b.add_material(rng.uniform(0.0, 0.9, 3) if k else (0.01, 0.02, 0.03), float(rng.rand() < 0.3),
                           rng.uniform(0.1, 1.0), (*rng.uniform(0, 1, 3), rng.uniform(0, 2)) if k % 3 == 2 else (0, 0, 0, 0))
            for k in range(nmat)]
    meshes = []
    per = max(1, ntris // 4)
    for m in range(4):
        c = rng.uniform(-0.5, 0.5, (per, 1, 3))
        tri = c + rng.uniform(-0.12, 0.12, (per, 3, 3))
        n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
        nrm = np.repeat(n, 3, 0) + rng.uniform(-0.2, 0.2, (per * 3, 3))
        col = rng.uniform(0.2, 1.0, (per * 3, 3)) if m % 2 else (1.0, 1.0, 1.0)
        meshes.append(b.add_mesh(tri.reshape(-1, 3).astype(np.float32), nrm.astype(np.float32),
                                 np.arange(per * 3, dtype=np.uint32), col))
    for k in range(6):
        M = translate(rng.uniform(-extent / 3, extent / 3, 3) + (0, 2, 0)) @ rotate(k % 3, rng.uniform(0, 360)) @ \
            rotate((k + 1) % 3, rng.uniform(0, 360)) @ scale(rng.uniform(extent * 0.3, extent, 3))
        if k == 4:   # a two-mesh renderable
            b.add_renderable(M, [meshes[0], meshes[3]], [mats[k % nmat], mats[(k + 1) % nmat]])
        else:
            b.add_renderable(M, meshes[k % 4], mats[k % nmat], visible=(3 if k == 5 else 1))   # k==5: cull bit set
    for k in range(nlights):
        b.add_light(rng.uniform(-2, 2, 3) + (0, 2, 0), rng.uniform(4, 9), tuple(rng.uniform(1, 5, 3)), enabled=(k != 1 or nlights < 3))
    return b.finish(camera, grid)


def _procedural_texture(rng, w, h, channels):
    """Smooth noise + checker + a few hard edges, as uint8 [h, w, channels]."""
    y, x = np.mgrid[0:h, 0:w]
    t = np.zeros((h, w, channels))
    for c in range(channels):
        fx, fy, ph = rng.uniform(0.5, 4.0), rng.uniform(0.5, 4.0), rng.uniform(0, 6.28)
        t[..., c] = 0.5 + 0.35 * np.sin(2 * np.pi * (fx * x / w + fy * y / h) + ph) + 0.15 * (((x // 4) + (y // 4)) % 2)
    t += rng.uniform(-0.08, 0.08, t.shape)
    return np.clip(t * 255.0, 0, 255).astype(np.uint8)


def textured_soup(seed, ntris=1200, grid=(4, 3, 64), camera=(0.3, 2.0, -0.2)):
    """Triangle soup whose materials sample bindless textures (bindless.glsl:196-233): albedo (sRGB and UNORM, REPEAT and
    CLAMP_TO_EDGE, uvs well outside [0,1]), metallic-roughness (RGBA8 and RG8), emissive, a 1x1 texture, a material whose only
    texture is the (never sampled) normal map, plus factor-only and vertex-colour materials."""
    b = SceneBuilder("texsoup%d" % seed, seed)
    rng = b.rng
    t_alb_srgb = b.add_texture(_procedural_texture(rng, 32, 16, 4), L.TEX_RGBA8_SRGB)
    t_alb_unorm = b.add_texture(_procedural_texture(rng, 8, 8, 4), L.TEX_RGBA8_UNORM, clamp_to_edge=True)
    t_mr = b.add_texture(_procedural_texture(rng, 16, 16, 4), L.TEX_RGBA8_UNORM)
    t_mr_rg = b.add_texture(_procedural_texture(rng, 5, 7, 2), L.TEX_RG8_UNORM)
    t_emis = b.add_texture(_procedural_texture(rng, 16, 4, 4), L.TEX_RGBA8_SRGB)
    t_one = b.add_texture(np.array([[[200, 100, 50, 255]]], np.uint8), L.TEX_RGBA8_SRGB)
    t_r8 = b.add_texture(_procedural_texture(rng, 4, 4, 1), L.TEX_R8_UNORM)
    mats = [
        b.add_material((1.0, 1.0, 1.0), 0.0, 1.0, tex=(-1, t_alb_srgb, -1, -1)),
        b.add_material((0.8, 0.6, 0.9), 0.3, 0.5, tex=(t_mr, t_alb_unorm, -1, -1)),
        b.add_material((0.5, 0.5, 0.5), 0.0, 0.7, (1.0, 0.8, 0.5, 2.0), tex=(t_mr_rg, -1, -1, t_emis)),
        b.add_material((0.9, 0.9, 0.9), 0.1, 0.9, tex=(-1, t_one, 5, -1)),            # normal-map index is ignored on this path
        b.add_material((0.01, 0.02, 0.03), 0.0, 0.6, (0.2, 0.3, 0.1, 1.5), tex=(-1, -1, -1, t_r8)),   # vertex colour + R8 emissive
        b.add_material((0.7, 0.2, 0.2), 0.0, 0.8),                                     # factor only
    ]
    meshes = []
    per = max(1, ntris // 4)
    for m in range(4):
        c = rng.uniform(-0.5, 0.5, (per, 1, 3))
        tri = c + rng.uniform(-0.15, 0.15, (per, 3, 3))
        n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
        nrm = np.repeat(n, 3, 0) + rng.uniform(-0.2, 0.2, (per * 3, 3))
        col = rng.uniform(0.2, 1.0, (per * 3, 3)) if m % 2 else (1.0, 1.0, 1.0)
        uv = rng.uniform(-1.5, 2.5, (per * 3, 2)).astype(np.float32)
        meshes.append(b.add_mesh(tri.reshape(-1, 3).astype(np.float32), nrm.astype(np.float32),
                                 np.arange(per * 3, dtype=np.uint32), col, uv))
    for k in range(6):
        M = translate(rng.uniform(-2, 2, 3) + (0, 2, 0)) @ rotate(k % 3, rng.uniform(0, 360)) @ scale(rng.uniform(2.0, 6.0, 3))
        b.add_renderable(M, meshes[k % 4], mats[k % len(mats)])
    b.add_light((0.5, 3.0, 0.2), 8.0, (3.0, 2.0, 4.0))
    return b.finish(camera, grid)
