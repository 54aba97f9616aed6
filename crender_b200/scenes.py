"""Deterministic procedural scenes for the BASELINE.json configs (the reference has no scene generator and
no headless mode; the model_data-shaped arrays produced here are what its OBJ loader would
hand to scene::add_model, src/util/asset_loader.h:16-30).

Everything is numpy and seeded; the same arrays are fed to the CUDA library and, in tests, to the CPU
oracle. A SceneDesc is backend-neutral: `load(desc, scn)` works with any object exposing the cr::scene
methods (crender_b200.api.scene or the oracle binding used by tests).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .api import GLASS, METAL, RAY_DTYPE, SMOOTH, camera, material, sun


@dataclass
class MeshDesc:
    verts: np.ndarray  # (T,3,3) f32
    mat_idx: np.ndarray  # (T,) u32
    materials: list
    uvs: Optional[np.ndarray] = None  # (T,3,2) f32
    instances: Optional[np.ndarray] = None  # (I,4,4) column-major; None = identity
    name: str = ""


@dataclass
class SceneDesc:
    name: str
    meshes: list
    cam: camera
    sun: sun = field(default_factory=sun)
    sun_enabled: bool = True
    textures: list = field(default_factory=list)
    skybox: Optional[np.ndarray] = None
    skybox_rotation: tuple = (0.0, 0.0)

    @property
    def n_source_tris(self) -> int:
        return sum(m.verts.shape[0] for m in self.meshes)

    @property
    def n_flat_tris(self) -> int:
        return sum(m.verts.shape[0] * (1 if m.instances is None else len(m.instances)) for m in self.meshes)

    def aabb(self):
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        for m in self.meshes:
            v = m.verts.reshape(-1, 3).astype(np.float64)
            xs = [np.eye(4)] if m.instances is None else [np.asarray(t, np.float64).reshape(4, 4).T for t in m.instances]
            for x in xs:
                # bounds of the transformed box corners are enough for ray generation
                w = v[:: max(1, len(v) // 200000)] @ x[:3, :3].T + x[:3, 3]
                lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
        return lo.astype(np.float32), hi.astype(np.float32)


def load(desc: SceneDesc, scn) -> list:
    """Feeds a SceneDesc through the cr::scene-shaped interface. Returns the model ids."""
    tex_ids = [scn.add_texture(t) for t in desc.textures]
    ids = []
    for m in desc.meshes:
        mid = scn.add_mesh(m.verts, m.uvs, m.mat_idx)
        mats = []
        for mm in m.materials:
            c = material(**{**mm.__dict__})
            if c.tex is not None:
                c.tex = tex_ids[c.tex]
            mats.append(c)
        scn.set_materials(mid, mats)
        if m.instances is not None:
            scn.set_instances(mid, m.instances)
        ids.append(mid)
    scn.set_sun(desc.sun)
    scn.set_sun_enabled(desc.sun_enabled)
    if desc.skybox is not None:
        scn.set_skybox(desc.skybox, desc.skybox_rotation)
    scn.set_camera(desc.cam)
    return ids


# ---------------------------------------------------------------------------------------------- helpers
def _quad(a, b, c, d):
    """Two triangles (a,b,c),(a,c,d); geometric normal = (b-a)x(c-a) (Embree's Ng convention)."""
    a, b, c, d = (np.asarray(p, np.float32) for p in (a, b, c, d))
    return np.stack([np.stack([a, b, c]), np.stack([a, c, d])])


def _box(center, half, yaw_deg):
    """Axis-aligned box rotated about +y, outward normals, 12 triangles."""
    cx, cy, cz = center
    hx, hy, hz = half
    c, s = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))

    def P(x, y, z):
        return (cx + c * x * hx + s * z * hz, cy + y * hy, cz - s * x * hx + c * z * hz)

    faces = [
        (P(-1, -1, -1), P(-1, 1, -1), P(1, 1, -1), P(1, -1, -1)),  # -z
        (P(-1, -1, 1), P(1, -1, 1), P(1, 1, 1), P(-1, 1, 1)),  # +z
        (P(-1, -1, -1), P(-1, -1, 1), P(-1, 1, 1), P(-1, 1, -1)),  # -x
        (P(1, -1, -1), P(1, 1, -1), P(1, 1, 1), P(1, -1, 1)),  # +x
        (P(-1, 1, -1), P(-1, 1, 1), P(1, 1, 1), P(1, 1, -1)),  # +y
        (P(-1, -1, -1), P(1, -1, -1), P(1, -1, 1), P(-1, -1, 1)),  # -y
    ]
    return np.concatenate([_quad(*f) for f in faces])


def translation(x, y, z) -> np.ndarray:
    """glm::translate(mat4(1), (x,y,z)) as a (4,4) column-major array m[c][r]."""
    m = np.eye(4, dtype=np.float32)
    m[3, :3] = (x, y, z)
    return m


def rotation_y(deg) -> np.ndarray:
    c, s = np.float32(np.cos(np.radians(deg))), np.float32(np.sin(np.radians(deg)))
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c  # columns: m[c][r]
    return m


def compose(*ms) -> np.ndarray:
    """Matrix product ms[0] * ms[1] * ... of column-major (m[c][r]) matrices, result in the same storage."""
    out = np.eye(4, dtype=np.float32)
    for m in ms:
        out = (np.asarray(m, np.float32) @ out).astype(np.float32)  # storage of (A*B) is B_s @ A_s
    return out


def _hash32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


# ---------------------------------------------------------------------------------------------- C1
def cornell(light_emission: float = 15.0) -> SceneDesc:
    """BASELINE config 1: closed-ish Cornell box, 36 triangles, diffuse + one emissive quad, sun off.
    Walls are wound so that the geometric normal points into the room (the reference does not
    face-forward normals, renderer.cpp:92-98)."""
    white, red, green = (0.73, 0.73, 0.73, 1.0), (0.65, 0.05, 0.05, 1.0), (0.12, 0.45, 0.15, 1.0)
    mats = [
        material(SMOOTH, colour=white, name="white"),
        material(SMOOTH, colour=red, name="red"),
        material(SMOOTH, colour=green, name="green"),
        material(SMOOTH, colour=(1.0, 1.0, 1.0, 1.0), emission=light_emission, name="light"),
    ]
    parts, idx = [], []

    def add(tris, m):
        parts.append(tris)
        idx.extend([m] * len(tris))

    add(_quad((-1, -1, -1), (-1, -1, 1), (1, -1, 1), (1, -1, -1)), 0)  # floor, +y
    add(_quad((-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1)), 0)  # ceiling, -y
    add(_quad((-1, -1, 1), (-1, 1, 1), (1, 1, 1), (1, -1, 1)), 0)  # back, -z
    add(_quad((-1, -1, -1), (-1, 1, -1), (-1, 1, 1), (-1, -1, 1)), 1)  # left, +x
    add(_quad((1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, -1)), 2)  # right, -x
    add(_quad((-0.3, 0.998, -0.3), (0.3, 0.998, -0.3), (0.3, 0.998, 0.3), (-0.3, 0.998, 0.3)), 3)  # light, -y
    add(_box((-0.35, -0.7, -0.25), (0.3, 0.3, 0.3), 18.0), 0)  # short box
    add(_box((0.35, -0.4, 0.35), (0.3, 0.6, 0.3), -17.0), 0)  # tall box
    verts = np.concatenate(parts).astype(np.float32)
    cam = camera(position=(0.0, 0.0, -3.4), fov=40.0)
    return SceneDesc("cornell", [MeshDesc(verts, np.asarray(idx, np.uint32), mats, name="cornell")], cam, sun_enabled=False)


# ---------------------------------------------------------------------------------------------- C2 / C3
def displaced_sphere(nu: int, nv: int, seed: int = 1, radius: float = 1.0, amp: float = 0.12):
    """Closed-ish tessellated surface: nu x nv quads -> 2*nu*nv triangles, outward normals."""
    rs = np.random.RandomState(seed)
    k = 6
    fu, fv = rs.randint(1, 9, k), rs.randint(1, 7, k)
    ph, am = rs.uniform(0, 2 * np.pi, (k, 2)), rs.uniform(0.3, 1.0, k)
    eps = 1e-3
    phi = np.linspace(0.0, 2 * np.pi, nu + 1)
    theta = np.linspace(eps, np.pi - eps, nv + 1)
    P, T = np.meshgrid(phi, theta, indexing="ij")
    disp = np.zeros_like(P)
    for i in range(k):
        disp += am[i] * np.sin(fu[i] * P + ph[i, 0]) * np.sin(fv[i] * T * 2 + ph[i, 1])
    r = radius * (1.0 + amp * disp / am.sum())
    pts = np.stack([r * np.sin(T) * np.cos(P), r * np.cos(T), r * np.sin(T) * np.sin(P)], -1).astype(np.float32)
    p00, p10, p11, p01 = pts[:-1, :-1], pts[1:, :-1], pts[1:, 1:], pts[:-1, 1:]
    t0 = np.stack([p00, p11, p10], -2)  # outward for this parametrisation (checked in tests)
    t1 = np.stack([p00, p01, p11], -2)
    tris = np.stack([t0, t1], 2).reshape(-1, 3, 3)
    # per-quad indices for material patches
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    quad_id = np.stack([iu, iv], -1)
    quad_id = np.repeat(quad_id[:, :, None, :], 2, axis=2).reshape(-1, 2)
    return np.ascontiguousarray(tris), quad_id


def mesh_scene(nu: int = 1000, nv: int = 500, seed: int = 1, with_ground: bool = True, sun_enabled: bool = True) -> SceneDesc:
    """BASELINE config 2/3: 2*nu*nv-triangle displaced sphere (1M at the defaults) with three material
    groups in patches (Lambert / mirror metal / glass, the reference's three shade types) on a ground
    quad, default sun."""
    tris, qid = displaced_sphere(nu, nv, seed)
    patch = (qid[:, 0] // max(1, nu // 40)).astype(np.uint32) * np.uint32(977) + (qid[:, 1] // max(1, nv // 20)).astype(np.uint32)
    h = _hash32(patch + np.uint32(seed * 7919))
    sel = h % np.uint32(100)
    palette = [(0.75, 0.75, 0.75, 1), (0.8, 0.3, 0.25, 1), (0.25, 0.5, 0.8, 1), (0.85, 0.75, 0.3, 1), (0.3, 0.7, 0.4, 1)]
    mats = [material(SMOOTH, colour=c, name=f"lambert{i}") for i, c in enumerate(palette)]
    mats.append(material(METAL, colour=(0.9, 0.9, 0.9, 1), reflectiveness=0.9, name="metal"))
    mats.append(material(GLASS, colour=(0.95, 0.98, 1.0, 1), ior=1.5, name="glass"))
    mat_idx = np.where(sel < 60, (h >> np.uint32(8)) % np.uint32(5), np.where(sel < 85, np.uint32(5), np.uint32(6))).astype(np.uint32)
    meshes = [MeshDesc(tris, mat_idx, mats, name="sphere")]
    if with_ground:
        g = _quad((-8, -1.3, -8), (-8, -1.3, 8), (8, -1.3, 8), (8, -1.3, -8)).astype(np.float32)
        meshes.append(MeshDesc(g, np.zeros(2, np.uint32), [material(SMOOTH, colour=(0.6, 0.6, 0.6, 1), name="ground")], name="ground"))
    cam = camera(position=(0.0, 0.0, -3.5), fov=45.0)
    return SceneDesc(f"mesh{2 * nu * nv}", meshes, cam, sun_enabled=sun_enabled)


def random_rays(lo, hi, n: int, seed: int = 2, inflate: float = 1.5, occlusion: bool = False) -> np.ndarray:
    """BASELINE config 3 ray batch: origins uniform in the scene AABB inflated `inflate`x, directions
    uniform on the sphere (the sampling.h:156-166 mapping), tmin 1e-5, tmax inf (or uniform in
    (0, diagonal) for occlusion queries)."""
    rng = np.random.Generator(np.random.Philox(seed))
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    c, e = 0.5 * (lo + hi), 0.5 * (hi - lo) * inflate
    rays = np.empty(n, dtype=RAY_DTYPE)
    rays["o"] = (c + (rng.random((n, 3)) * 2 - 1) * e).astype(np.float32)
    u = rng.random((n, 2))
    ct = 2 * u[:, 0] - 1
    st = np.sqrt(np.maximum(0.0, 1 - ct * ct))
    ph = 2 * np.pi * u[:, 1]
    rays["d"] = np.stack([st * np.cos(ph), ct, st * np.sin(ph)], -1).astype(np.float32)
    rays["tmin"] = np.float32(1e-5)
    if occlusion:
        rays["tmax"] = (rng.random(n) * np.linalg.norm(hi - lo)).astype(np.float32)
    else:
        rays["tmax"] = np.float32(np.inf)
    return rays


# ---------------------------------------------------------------------------------------------- C4
def terrain_city(n: int = 1000, grid: int = 3, seed: int = 3, n_buildings: int = 64) -> SceneDesc:
    """BASELINE config 4: an n x n heightfield tile (2*n*n triangles) instanced on a grid x grid lattice
    plus a block of box "buildings" instanced per tile. grid=3, n=1000 -> 18M+ flattened triangles."""
    rs = np.random.RandomState(seed)
    xs = np.linspace(-1.0, 1.0, n + 1)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    H = np.zeros_like(X)
    for _ in range(5):
        fx, fz = rs.randint(1, 6, 2)
        H += rs.uniform(0.2, 1.0) * np.sin(np.pi * fx * X + rs.uniform(0, 6.28)) * np.sin(np.pi * fz * Z + rs.uniform(0, 6.28))
    # tileable: sin(pi*k*x) terms repeat with period 2 = the tile size
    Y = 0.06 * H
    pts = np.stack([X, Y, Z], -1).astype(np.float32)
    p00, p10, p11, p01 = pts[:-1, :-1], pts[1:, :-1], pts[1:, 1:], pts[:-1, 1:]
    tris = np.stack([np.stack([p00, p01, p11], -2), np.stack([p00, p11, p10], -2)], 2).reshape(-1, 3, 3)  # +y normals
    iu, iv = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    patch = ((iu // max(1, n // 16)) * 131 + (iv // max(1, n // 16))).astype(np.uint32)
    patch = np.repeat(patch[:, :, None], 2, axis=2).reshape(-1)
    mats = [material(SMOOTH, colour=c, name=f"soil{i}") for i, c in enumerate([(0.45, 0.5, 0.3, 1), (0.55, 0.5, 0.4, 1), (0.35, 0.45, 0.3, 1)])]
    mat_idx = (_hash32(patch) % np.uint32(3)).astype(np.uint32)
    inst_t = np.stack([translation(2.0 * (i - (grid - 1) / 2), 0.0, 2.0 * (j - (grid - 1) / 2)) for i in range(grid) for j in range(grid)])
    terrain = MeshDesc(np.ascontiguousarray(tris), mat_idx, mats, instances=inst_t, name="terrain")
    boxes, bidx = [], []
    for b in range(n_buildings):
        cx, cz = rs.uniform(-0.9, 0.9, 2)
        hh = rs.uniform(0.05, 0.35)
        boxes.append(_box((cx, hh - 0.02, cz), (rs.uniform(0.02, 0.06), hh, rs.uniform(0.02, 0.06)), rs.uniform(0, 90)))
        bidx.extend([b % 3] * 12)
    bm = [
        material(SMOOTH, colour=(0.7, 0.7, 0.72, 1), name="concrete"),
        material(METAL, colour=(0.85, 0.85, 0.9, 1), reflectiveness=0.85, name="steel"),
        material(GLASS, colour=(0.9, 0.95, 1.0, 1), ior=1.45, name="glazing"),
    ]
    city = MeshDesc(np.concatenate(boxes).astype(np.float32), np.asarray(bidx, np.uint32), bm, instances=inst_t.copy(), name="city")
    cam = camera(position=(0.0, 0.9, -float(grid) - 0.4), fov=50.0, rotation=(0.0, 14.0, 0.0))
    return SceneDesc(f"terrain{grid}x{grid}x{2 * n * n}", [terrain, city], cam, sun_enabled=True)


# ---------------------------------------------------------------------------------------------- C5
def lights_scene(nu: int = 1000, nv: int = 500, n_lights: int = 64, seed: int = 5) -> SceneDesc:
    """BASELINE config 5: the config-2 mesh plus n_lights emissive quads. The reference has no
    area-light NEE: emitters are found by path hits only, which is what is reproduced."""
    d = mesh_scene(nu, nv, seed=1, with_ground=True, sun_enabled=True)
    rs = np.random.RandomState(seed)
    quads = []
    for _ in range(n_lights):
        cx, cz = rs.uniform(-3, 3, 2)
        cy, s = rs.uniform(1.6, 2.6), rs.uniform(0.08, 0.2)
        quads.append(_quad((cx - s, cy, cz - s), (cx + s, cy, cz - s), (cx + s, cy, cz + s), (cx - s, cy, cz + s)))  # -y normal
    lm = [material(SMOOTH, colour=(1, 1, 1, 1), emission=20.0, name="emitter")]
    d.meshes.append(MeshDesc(np.concatenate(quads).astype(np.float32), np.zeros(2 * n_lights, np.uint32), lm, name="lights"))
    d.name = f"lights{n_lights}_{d.name}"
    return d


# ---------------------------------------------------------------------------------------------- N2 coverage
def textured_scene(seed: int = 7) -> SceneDesc:
    """Small scene exercising textures, alpha cut-outs, skybox, instancing with rotation, all three
    shade types and the sun (parity coverage for SURVEY.md §8 rows a6, a10, a17 and 'next' N2)."""
    rs = np.random.RandomState(seed)
    tw, th = 16, 8
    tex = rs.uniform(0.1, 1.0, (th, tw, 4)).astype(np.float32)
    tex[..., 3] = 1.0
    cut = rs.uniform(0.2, 1.0, (8, 8, 4)).astype(np.float32)
    cut[..., 3] = ((np.add.outer(np.arange(8), np.arange(8)) % 2) == 0).astype(np.float32)  # checkerboard alpha
    sky = rs.uniform(0.0, 1.5, (16, 32, 4)).astype(np.float32)
    sky[..., 3] = 1.0
    ground = _quad((-4, 0, -4), (-4, 0, 4), (4, 0, 4), (4, 0, -4)).astype(np.float32)
    guv = np.array([[[0, 0], [0, 3], [3, 3]], [[0, 0], [3, 3], [3, 0]]], np.float32)
    m_ground = MeshDesc(ground, np.zeros(2, np.uint32), [material(SMOOTH, tex=0, name="tex-ground")], uvs=guv, name="ground")
    fence = _quad((-1.5, 0, 0.2), (-1.5, 1.2, 0.2), (1.5, 1.2, 0.2), (1.5, 0, 0.2)).astype(np.float32)  # -z facing
    fuv = np.array([[[0, 0], [0, 1], [2, 1]], [[0, 0], [2, 1], [2, 0]]], np.float32)
    m_fence = MeshDesc(fence, np.zeros(2, np.uint32), [material(SMOOTH, tex=1, name="cutout")], uvs=fuv, name="fence")
    box = _box((0, 0.37, 0), (0.3, 0.35, 0.3), 0.0).astype(np.float32)
    inst = np.stack([translation(-1.0, 0, 1.2), compose(translation(1.0, 0.0, 1.4), rotation_y(30.0)), translation(0.0, 0.0, 2.2)])
    bm = [material(SMOOTH, colour=(0.8, 0.4, 0.3, 1), name="clay"), material(METAL, colour=(0.9, 0.9, 0.9, 1), reflectiveness=0.8), material(GLASS, ior=1.4)]
    bidx = np.asarray([0] * 4 + [1] * 4 + [2] * 4, np.uint32)
    m_box = MeshDesc(box, bidx, bm, instances=inst.astype(np.float32), name="boxes")
    cam = camera(position=(0.0, 0.9, -3.2), fov=55.0, rotation=(4.0, 8.0, 0.0))
    return SceneDesc("textured", [m_ground, m_fence, m_box], cam, textures=[tex, cut], skybox=sky, skybox_rotation=(0.1, 0.05), sun_enabled=True)
