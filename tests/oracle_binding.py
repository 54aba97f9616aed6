"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Exposes the same cr::scene / cr::renderer-shaped methods as crender_b200.api so that a SceneDesc can be
loaded into either side. Nothing in the crender_b200 package imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("model", "<u4"), ("inst", "<u4")])
RAW_SUM, PROGRESS, ALBEDO, NORMAL, DEPTH = 0, 1, 2, 3, 4


class OMaterial(C.Structure):
    _fields_ = [("shade_type", C.c_uint32), ("ior", C.c_float), ("roughness", C.c_float), ("reflectiveness", C.c_float),
                ("emission", C.c_float), ("colour", C.c_float * 4), ("tex", C.c_int32)]


class OSun(C.Structure):
    _fields_ = [("size", C.c_float), ("intensity", C.c_float), ("direction", C.c_float * 3), ("colour", C.c_float * 3)]


class OCamera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 3), ("fov", C.c_float), ("scale", C.c_float), ("mode", C.c_uint32)]


class OStats(C.Structure):
    _fields_ = [("total_queries", C.c_uint64), ("ref_rays", C.c_uint64), ("pixel_samples", C.c_uint64), ("passes", C.c_uint64)]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_LIB):
        build()
    L = C.CDLL(ORACLE_LIB)
    P = C.c_void_p
    L.orc_scene_create.restype = P
    L.orc_scene_destroy.argtypes = [P]
    L.orc_scene_add_mesh.restype = C.c_int
    L.orc_scene_add_mesh.argtypes = [P, P, P, P, C.c_uint32]
    L.orc_scene_set_materials.argtypes = [P, C.c_int, C.POINTER(OMaterial), C.c_uint32]
    L.orc_scene_set_instances.argtypes = [P, C.c_int, P, C.c_uint32]
    L.orc_scene_add_texture.argtypes = [P, P, C.c_uint32, C.c_uint32]
    L.orc_scene_set_sun.argtypes = [P, C.POINTER(OSun), C.c_int]
    L.orc_scene_set_skybox.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_float, C.c_float]
    L.orc_scene_set_camera.argtypes = [P, C.POINTER(OCamera)]
    L.orc_scene_commit.restype = C.c_double
    L.orc_scene_commit.argtypes = [P]
    for f in (L.orc_intersect_batch, L.orc_intersect_brute, L.orc_occluded_batch):
        f.argtypes = [P, P, P, C.c_uint64, C.c_int]
    L.orc_render_create.restype = P
    L.orc_render_create.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.orc_render_destroy.argtypes = [P]
    L.orc_render_reset.argtypes = [P]
    L.orc_render_set_extended.argtypes = [P, C.c_int]
    L.orc_kat_scatter_extended.restype = C.c_int
    L.orc_kat_scatter_extended.argtypes = [C.POINTER(OMaterial), P, P, P, C.c_float, C.c_float, P, P, P]
    L.orc_render_set_rows.argtypes = [P, C.c_uint32, C.c_uint32]
    L.orc_render_set_reference_stream.argtypes = [P, C.c_int, C.c_uint64]
    L.orc_render_set_sample_table.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_int]
    L.orc_render_samples.argtypes = [P, C.c_uint32, C.c_uint32, C.c_int]
    L.orc_render_read.argtypes = [P, C.c_int, P]
    L.orc_render_stats.argtypes = [P, C.POINTER(OStats)]
    L.orc_render_primary_hits.argtypes = [P, C.c_uint32, P, C.c_int]
    L.orc_kat_mt19937_randf.argtypes = [C.c_uint32, P]
    L.orc_kat_rng.restype = C.c_float
    L.orc_kat_rng.argtypes = [C.c_uint32] * 4
    L.orc_kat_camera_ray.argtypes = [C.POINTER(OCamera), C.c_float, C.c_float, C.c_float, P, P]
    L.orc_kat_build_local.argtypes = [P, P, P]
    L.orc_kat_sun_transform.argtypes = [P, P]
    L.orc_kat_map_to_solid_angle.argtypes = [C.c_float, C.c_float, C.c_float, P, P]
    L.orc_kat_sphere.argtypes = [C.c_float, C.c_float, P]
    L.orc_kat_process_hit.argtypes = [C.POINTER(OMaterial), P, P, P, C.c_float, C.c_float, P, P, P, C.POINTER(C.c_int)]
    L.orc_kat_resolve.restype = C.c_float
    L.orc_kat_resolve.argtypes = [C.c_float, C.c_uint32]
    L.orc_kat_tri.restype = C.c_int
    L.orc_kat_tri.argtypes = [P, P, P, P, P, C.c_float, C.c_float, P, P, P]
    L.orc_kat_sky_uv.argtypes = [P, P]
    L.orc_kat_image_get_uv_index.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32, P]
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def nthreads_default() -> int:
    return max(1, os.cpu_count() or 1)


def c_material(m) -> OMaterial:
    cm = OMaterial()
    cm.shade_type, cm.ior, cm.roughness, cm.reflectiveness, cm.emission = m.shade_type, m.ior, m.roughness, m.reflectiveness, m.emission
    cm.colour = (C.c_float * 4)(*m.colour)
    cm.tex = -1 if m.tex is None else int(m.tex)
    return cm


def c_camera(cam) -> OCamera:
    return OCamera((C.c_float * 3)(*cam.position), (C.c_float * 3)(*cam.rotation), cam.fov, cam.scale, cam.current_mode)


class scene:
    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.orc_scene_create())
        self._sun_enabled = True
        self.build_ms = 0.0

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_scene_destroy(self._h)
            self._h = None

    def add_mesh(self, verts, uvs=None, mat_idx=None) -> int:
        v = _f32(verts).reshape(-1, 9)
        u = None if uvs is None else _f32(uvs).reshape(-1, 6)
        m = None if mat_idx is None else np.ascontiguousarray(mat_idx, dtype=np.uint32)
        return self._L.orc_scene_add_mesh(self._h, _ptr(v), _ptr(u), _ptr(m), v.shape[0])

    def set_materials(self, mid, mats):
        arr = (OMaterial * len(mats))(*[c_material(m) for m in mats])
        assert self._L.orc_scene_set_materials(self._h, mid, arr, len(mats)) == 0

    def set_instances(self, mid, transforms):
        t = _f32(transforms).reshape(-1, 16)
        assert self._L.orc_scene_set_instances(self._h, mid, _ptr(t), t.shape[0]) == 0

    def add_texture(self, rgba) -> int:
        a = _f32(rgba)
        return self._L.orc_scene_add_texture(self._h, _ptr(a), a.shape[1], a.shape[0])

    def set_sun(self, s):
        cs = OSun(s.size, s.intensity, (C.c_float * 3)(*s.direction), (C.c_float * 3)(*s.colour))
        self._L.orc_scene_set_sun(self._h, C.byref(cs), int(self._sun_enabled))

    def set_sun_enabled(self, v):
        self._sun_enabled = bool(v)
        self._L.orc_scene_set_sun(self._h, None, int(self._sun_enabled))

    def set_skybox(self, rgba, rotation=(0.0, 0.0)):
        if rgba is None:
            self._L.orc_scene_set_skybox(self._h, None, 0, 0, rotation[0], rotation[1])
            return
        a = _f32(rgba)
        self._L.orc_scene_set_skybox(self._h, _ptr(a), a.shape[1], a.shape[0], rotation[0], rotation[1])

    def set_camera(self, cam):
        cc = c_camera(cam)
        self._L.orc_scene_set_camera(self._h, C.byref(cc))

    def commit(self):
        self.build_ms = self._L.orc_scene_commit(self._h)
        return self.build_ms

    def cast_rays(self, rays, brute=False, nthreads=None):
        r = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(r.shape[0], dtype=HIT_DTYPE)
        f = self._L.orc_intersect_brute if brute else self._L.orc_intersect_batch
        f(self._h, _ptr(r), _ptr(hits), r.shape[0], nthreads or nthreads_default())
        return hits

    def occluded(self, rays, nthreads=None):
        r = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.empty(r.shape[0], dtype=np.uint8)
        self._L.orc_occluded_batch(self._h, _ptr(r), _ptr(occ), r.shape[0], nthreads or nthreads_default())
        return occ


class renderer:
    def __init__(self, res_x, res_y, bounces, scn: scene, seed=0, extended=False):
        self._L = lib()
        self._scene = scn
        self._h = C.c_void_p(self._L.orc_render_create(scn._h, res_x, res_y, bounces, seed))
        if extended:
            self._L.orc_render_set_extended(self._h, int(extended))  # 2 = extended without the light list
        self._res = (res_x, res_y)
        self._next = 0

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_render_destroy(self._h)
            self._h = None

    def start(self):
        self._L.orc_render_reset(self._h)
        self._next = 0

    def set_rows(self, y0, y1):
        self._L.orc_render_set_rows(self._h, y0, y1)

    def set_reference_stream(self, on=True, discard=0):
        """The reference's own default-seeded mt19937 consumed in call order by one thread (render with nthreads=1)."""
        self._L.orc_render_set_reference_stream(self._h, int(on), int(discard))

    def set_sample_table(self, table, replay=False):
        """table: float32 [n_samples, h*w, dims], kept alive by this object. Recorded in reference-stream mode, or replayed."""
        assert table.dtype == np.float32 and table.flags["C_CONTIGUOUS"] and table.ndim == 3
        self._table = table
        self._L.orc_render_set_sample_table(self._h, _ptr(table), table.shape[0], table.shape[2], int(replay))

    def render(self, n_spp, first_sample=None, nthreads=None):
        first = self._next if first_sample is None else first_sample
        self._L.orc_render_samples(self._h, first, n_spp, nthreads or nthreads_default())
        self._next = first + n_spp

    def _read(self, kind):
        w, h = self._res
        out = np.empty((h, w, 4), dtype=np.float32)
        self._L.orc_render_read(self._h, kind, _ptr(out))
        return out

    def current_progress(self):
        return self._read(PROGRESS)

    def current_normals(self):
        return self._read(NORMAL)

    def current_albedos(self):
        return self._read(ALBEDO)

    def current_depths(self):
        return self._read(DEPTH)

    def raw_sum(self):
        return self._read(RAW_SUM)

    def current_stats(self) -> OStats:
        st = OStats()
        self._L.orc_render_stats(self._h, C.byref(st))
        return st

    def primary_hits(self, sample=0, nthreads=None):
        w, h = self._res
        hits = np.empty(w * h, dtype=HIT_DTYPE)
        self._L.orc_render_primary_hits(self._h, sample, _ptr(hits), nthreads or nthreads_default())
        return hits
