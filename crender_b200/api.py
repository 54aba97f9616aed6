"""Host-side mirror of CRender's render-facing classes over the C ABI (include/crender_b200.h).

Names, argument meaning and defaults follow the reference:
  cr::material::information  src/render/material/material.h:31-41
  cr::camera                 src/render/camera.h:11-41
  cr::entity::sun            src/render/entities/components.h:23-29
  cr::asset_loader::model_data  src/util/asset_loader.h:16-30
  cr::scene                  src/render/scene.h:19-52  (+ cr::registry)
  cr::renderer               src/render/renderer.h:24-105
This module only marshals arrays; all rendering work happens in libcrender_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import CrbError  # noqa: F401  (re-export)

METAL, SMOOTH, GLASS = 0, 1, 2  # cr::material::type
PERSPECTIVE, ORTHOGRAPHIC = 0, 1  # cr::camera::mode
RAW_SUM, PROGRESS, ALBEDO, NORMAL, DEPTH = 0, 1, 2, 3, 4
K_RAYGEN, K_TRACE, K_SHADE, K_SHADOW, K_ADVANCE, K_ACCUMULATE = range(6)  # crb_stats.kernel_ms index
PARTITION_SPP, PARTITION_TILE = 0, 1  # multi-GPU work split (BASELINE configs 4 / 5)
MERGE_NONE, MERGE_PEER_KERNEL, MERGE_NCCL = 0, 1, 2

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("model", "<u4"), ("inst", "<u4")])


@dataclass
class material:
    shade_type: int = SMOOTH
    ior: float = 1.5
    roughness: float = 0.5
    reflectiveness: float = 1.0
    emission: float = 0.0
    colour: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    tex: Optional[int] = None
    name: str = "ERROR - Report"


@dataclass
class camera:
    position: Sequence[float] = (5.0, 5.0, 0.0)
    fov: float = 75.0
    current_mode: int = PERSPECTIVE
    rotation: Sequence[float] = (0.0, 0.0, 0.0)
    scale: float = 1.0


@dataclass
class sun:
    size: float = 3.14159265359 / 48.0
    intensity: float = 100.0
    direction: Sequence[float] = (0.8 / math.sqrt(1.64), -1.0 / math.sqrt(1.64), 0.0)
    colour: Sequence[float] = (1.0, 0.9, 0.7)


@dataclass
class model_data:
    """cr::asset_loader::model_data (indexed form, as the OBJ loader produces it)."""

    name: str = ""
    vertices: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    texture_coords: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))
    materials: list = field(default_factory=list)
    textures: list = field(default_factory=list)  # RGBA float32 arrays (h, w, 4)
    vertex_indices: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    material_indices: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    texture_indices: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c_material(m: material) -> _capi.Material:
    cm = _capi.Material()
    cm.shade_type, cm.ior, cm.roughness, cm.reflectiveness, cm.emission = m.shade_type, m.ior, m.roughness, m.reflectiveness, m.emission
    cm.colour = (C.c_float * 4)(*m.colour)
    cm.tex = -1 if m.tex is None else int(m.tex)
    return cm


def post_settings(use_bloom=False, bloom_threshold=0.7, bloom_strength=1.0, use_gray_scale=False, use_tonemapping=False, tonemapping_type=0,
                  tonemapping_exposure=1.0, gamma_correction=2.2) -> _capi.PostSettings:
    """cr::post_processor settings with the reference's defaults (post_processor.h:19-39)."""
    return _capi.PostSettings(int(use_bloom), bloom_threshold, bloom_strength, int(use_gray_scale), int(use_tonemapping), tonemapping_type,
                              tonemapping_exposure, gamma_correction)


class scene:
    """cr::scene. Geometry changes take effect at commit() (the rtcCommitScene point)."""

    def __init__(self, lib_path: Optional[str] = None, device: int = -1):
        self._lib = _capi.load(lib_path)
        _capi.check(self._lib, self._lib.crb_set_device(device))
        h = C.c_void_p()
        _capi.check(self._lib, self._lib.crb_scene_create(C.byref(h)))
        self._h = h
        self._sun_enabled = True
        self.build_info: Optional[_capi.BuildInfo] = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.crb_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- geometry
    def add_mesh(self, verts, uvs=None, mat_idx=None) -> int:
        """De-indexed triangle soup: verts (T,3,3) f32, uvs (T,3,2) f32 or None, mat_idx (T,) u32."""
        v = _f32(verts).reshape(-1, 9)
        t = v.shape[0]
        u = None if uvs is None else _f32(uvs).reshape(-1, 6)
        m = None if mat_idx is None else np.ascontiguousarray(mat_idx, dtype=np.uint32)
        if u is not None and u.shape[0] != t:
            raise ValueError("uvs must have one (3,2) entry per triangle")
        if m is not None and m.shape[0] != t:
            raise ValueError("mat_idx must have one entry per triangle")
        mid = C.c_int(-1)
        _capi.check(self._lib, self._lib.crb_scene_add_mesh(self._h, _ptr(v), _ptr(u), _ptr(m), t, C.byref(mid)))
        return mid.value

    def add_model(self, model: model_data) -> int:
        """scene::add_model -> registry::register_model (registry.cpp:51-97): expands the indexed data to a
        flat triangle soup, registers textures, remaps material texture handles."""
        vi = np.asarray(model.vertex_indices, dtype=np.int64)
        verts = _f32(model.vertices)[vi].reshape(-1, 3, 3)
        uvs = None
        if len(model.texture_coords) and len(model.texture_indices):
            uvs = _f32(model.texture_coords)[np.asarray(model.texture_indices, dtype=np.int64)].reshape(-1, 3, 2)
        handles = [self.add_texture(t) for t in model.textures]
        mid = self.add_mesh(verts, uvs, model.material_indices if len(model.material_indices) else None)
        if model.materials:
            mats = []
            for m in model.materials:
                mm = material(**{**m.__dict__})
                if mm.tex is not None:
                    mm.tex = handles[mm.tex]
                mats.append(mm)
            self.set_materials(mid, mats)
        return mid

    def set_materials(self, model_id: int, mats: Sequence[material]):
        arr = (_capi.Material * len(mats))(*[_c_material(m) for m in mats])
        _capi.check(self._lib, self._lib.crb_scene_set_materials(self._h, model_id, arr, len(mats)))

    def set_instances(self, model_id: int, transforms):
        """transforms: (I,4,4) column-major glm::mat4 (i.e. transforms[i][c][r])."""
        t = _f32(transforms).reshape(-1, 16)
        _capi.check(self._lib, self._lib.crb_scene_set_instances(self._h, model_id, _ptr(t), t.shape[0]))

    def add_texture(self, rgba) -> int:
        a = _f32(rgba)
        if a.ndim != 3 or a.shape[2] != 4:
            raise ValueError("texture must be (h, w, 4) float32 RGBA")
        tid = C.c_int(-1)
        _capi.check(self._lib, self._lib.crb_scene_add_texture(self._h, _ptr(a), a.shape[1], a.shape[0], C.byref(tid)))
        return tid.value

    # ---- environment
    def set_sun(self, s: sun):
        cs = _capi.Sun(s.size, s.intensity, (C.c_float * 3)(*s.direction), (C.c_float * 3)(*s.colour))
        _capi.check(self._lib, self._lib.crb_scene_set_sun(self._h, C.byref(cs), int(self._sun_enabled)))

    def set_sun_enabled(self, value: bool):
        self._sun_enabled = bool(value)
        _capi.check(self._lib, self._lib.crb_scene_set_sun(self._h, None, int(self._sun_enabled)))

    def is_sun_enabled(self) -> bool:
        return self._sun_enabled

    def set_skybox(self, rgba, rotation=(0.0, 0.0)):
        if rgba is None:
            _capi.check(self._lib, self._lib.crb_scene_set_skybox(self._h, None, 0, 0, rotation[0], rotation[1]))
            return
        a = _f32(rgba)
        _capi.check(self._lib, self._lib.crb_scene_set_skybox(self._h, _ptr(a), a.shape[1], a.shape[0], rotation[0], rotation[1]))

    def set_camera(self, cam: camera):
        cc = _capi.Camera((C.c_float * 3)(*cam.position), (C.c_float * 3)(*cam.rotation), cam.fov, cam.scale, cam.current_mode)
        _capi.check(self._lib, self._lib.crb_scene_set_camera(self._h, C.byref(cc)))

    def set_flatten_instances(self, on: bool):
        """Instanced scenes: expand the instances into world-space triangles under one BVH (the fast path) instead of
        the default two-level traversal that reproduces the reference's per-instance arithmetic bit for bit."""
        _capi.check(self._lib, self._lib.crb_scene_set_option(self._h, 1, int(bool(on))))

    def commit(self) -> _capi.BuildInfo:
        info = _capi.BuildInfo()
        _capi.check(self._lib, self._lib.crb_scene_commit(self._h, C.byref(info)))
        self.build_info = info
        return info

    # ---- queries
    def cast_rays(self, rays: np.ndarray) -> np.ndarray:
        """Batch scene::cast_ray. rays: structured RAY_DTYPE array (host). Returns HIT_DTYPE array."""
        r = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(r.shape[0], dtype=HIT_DTYPE)
        _capi.check(self._lib, self._lib.crb_intersect_batch(self._h, _ptr(r), _ptr(hits), r.shape[0], 0))
        return hits

    def occluded(self, rays: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.empty(r.shape[0], dtype=np.uint8)
        _capi.check(self._lib, self._lib.crb_occluded_batch(self._h, _ptr(r), _ptr(occ), r.shape[0], 0))
        return occ

    def cast_rays_device(self, rays_ptr: int, hits_ptr: int, n: int):
        """Device-pointer variant (rays: n x 32 B, hits: n x 24 B, both 16-byte aligned)."""
        _capi.check(self._lib, self._lib.crb_intersect_batch(self._h, C.c_void_p(rays_ptr), C.c_void_p(hits_ptr), n, 1))

    def occluded_device(self, rays_ptr: int, occ_ptr: int, n: int):
        _capi.check(self._lib, self._lib.crb_occluded_batch(self._h, C.c_void_p(rays_ptr), C.c_void_p(occ_ptr), n, 1))

    def trace_counters(self, rays: np.ndarray, any_hit: bool = False):
        r = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        a, b = C.c_uint64(0), C.c_uint64(0)
        _capi.check(self._lib, self._lib.crb_trace_counters(self._h, _ptr(r), r.shape[0], 0, int(any_hit), C.byref(a), C.byref(b)))
        return a.value, b.value

    def microbench_read(self, nbytes: int, iters: int) -> float:
        """GB/s of 16-byte reads over a working set of nbytes (L2 vs HBM roofline context)."""
        g = C.c_double(0)
        _capi.check(self._lib, self._lib.crb_microbench_read(self._h, nbytes, iters, C.byref(g)))
        return g.value

    def post_process(self, rgba, settings: _capi.PostSettings) -> np.ndarray:
        """cr::post_processor::process on a host RGBA image, run as CUDA kernels."""
        a = _f32(rgba)
        out = np.empty_like(a)
        _capi.check(self._lib, self._lib.crb_post_process(self._h, _ptr(a), a.shape[1], a.shape[0], C.byref(settings), _ptr(out)))
        return out

    def last_query_ms(self) -> float:
        ms = C.c_double(0)
        _capi.check(self._lib, self._lib.crb_last_query_ms(self._h, C.byref(ms)))
        return ms.value

    def stream(self) -> int:
        p = C.c_void_p()
        _capi.check(self._lib, self._lib.crb_scene_stream(self._h, C.byref(p)))
        return p.value or 0


class renderer:
    """cr::renderer(res_x, res_y, bounces, pool, scene). The thread pool argument has no equivalent: the
    GPU grid replaces it. Progressive passes are requested explicitly with render(n) (the reference's
    management thread issues them in a loop until the target spp, renderer.cpp:116-144)."""

    def __init__(self, res_x: int, res_y: int, bounces: int, scn: scene, seed: int = 0, counters: bool = False, timers: bool = False,
                 material_sort: bool = False, extended: bool = False, gpus=None, partition: int = PARTITION_SPP, comm=None):
        """gpus: None = the scene's GPU; an int N or a list of device ids = one process driving N GPUs
        (crb_render_create_multi: the library replicates the scene and merges the accumulators itself).
        comm = (nccl_id_bytes, rank, nranks): one process per GPU (crb_render_create_rank)."""
        self._lib = scn._lib
        self._scene = scn
        h = C.c_void_p()
        flags = (1 if counters else 0) | (2 if timers else 0) | (4 if material_sort else 0) | (8 if extended else 0)
        if comm is not None:
            ident, rank, nranks = comm
            buf = C.create_string_buffer(bytes(ident), 128) if ident is not None else None
            _capi.check(self._lib, self._lib.crb_render_create_rank(scn._h, buf, rank, nranks, partition, res_x, res_y, bounces, seed, flags, C.byref(h)))
        elif gpus is not None:
            devs = list(range(gpus)) if isinstance(gpus, int) else [int(d) for d in gpus]
            arr = (C.c_int * len(devs))(*devs)
            _capi.check(self._lib, self._lib.crb_render_create_multi(scn._h, arr, len(devs), partition, res_x, res_y, bounces, seed, flags, C.byref(h)))
        else:
            _capi.check(self._lib, self._lib.crb_render_create(scn._h, res_x, res_y, bounces, seed, flags, C.byref(h)))
        self._h = h
        self._res = (res_x, res_y)
        self._spp_target = 0
        self._next_sample = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.crb_render_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- control (renderer.cpp:154-218)
    def start(self) -> bool:
        """Clears the accumulation and, if a target spp is set, renders up to it (the library's own target loop:
        crb_render_set_target_spp + crb_render_run, renderer.cpp:116-144)."""
        _capi.check(self._lib, self._lib.crb_render_reset(self._h))
        self._next_sample = 0
        if self._spp_target:
            total = C.c_uint64(0)
            _capi.check(self._lib, self._lib.crb_render_set_target_spp(self._h, self._spp_target))
            _capi.check(self._lib, self._lib.crb_render_run(self._h, 16, C.byref(total)))
            _capi.check(self._lib, self._lib.crb_render_sync(self._h))
            self._next_sample = int(total.value)
        return True

    def pause(self) -> bool:
        _capi.check(self._lib, self._lib.crb_render_sync(self._h))
        return True

    def update(self, fn: Callable[[], None]):
        """pause -> mutate -> restart from sample 0 (renderer.cpp:185-192)."""
        self.pause()
        fn()
        _capi.check(self._lib, self._lib.crb_render_refresh(self._h))
        self.start()

    def set_resolution(self, x: int, y: int):
        _capi.check(self._lib, self._lib.crb_render_set_resolution(self._h, x, y))
        self._res = (x, y)
        self._next_sample = 0

    def set_max_bounces(self, bounces: int):
        _capi.check(self._lib, self._lib.crb_render_set_max_bounces(self._h, bounces))

    def set_target_spp(self, target: int):
        self._spp_target = int(target)

    def set_rows(self, y0: int, y1: int):
        _capi.check(self._lib, self._lib.crb_render_set_rows(self._h, y0, y1))

    def set_bands(self, band_rows: int, first: int, stride: int, serpentine: bool = False):
        """Interleaved row bands first, first+stride, ... of band_rows rows each, rendered as one launch sequence
        (serpentine: the owner order is reversed in every other period of `stride` bands, as the library's tile partition does)."""
        if serpentine:
            _capi.check(self._lib, self._lib.crb_render_set_bands_ordered(self._h, band_rows, first, stride, 1))
        else:
            _capi.check(self._lib, self._lib.crb_render_set_bands(self._h, band_rows, first, stride))

    def set_sample_table(self, table: Optional[np.ndarray]):
        """A caller-supplied sample table float32 [n_samples, h*w, dims] replaces the hash sampler (ref-exact mode):
        dimensions 0,1 jitter; 2+4i+{0,1} scatter of bounce i; 2+4i+{2,3} sun sample of bounce i. None removes it."""
        if table is None:
            _capi.check(self._lib, self._lib.crb_render_set_sample_table(self._h, None, 0, 0))
            return
        t = np.ascontiguousarray(table, dtype=np.float32)
        w, h = self._res
        if t.ndim != 3 or t.shape[1] != w * h:
            raise ValueError("table must be [n_samples, h*w, dims]")
        _capi.check(self._lib, self._lib.crb_render_set_sample_table(self._h, _ptr(t), t.shape[0], t.shape[2]))

    def flush(self):
        """Multi-GPU handles: start merging the accumulators now (asynchronous; the read calls imply it)."""
        _capi.check(self._lib, self._lib.crb_render_flush(self._h))

    def join_flush(self):
        """The render stream waits for the last merge (for device-side timing of a whole step)."""
        _capi.check(self._lib, self._lib.crb_render_join_flush(self._h))

    def info(self) -> dict:
        a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        _capi.check(self._lib, self._lib.crb_render_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"gpus_local": a.value, "ranks": b.value, "partition": ("spp", "tile")[c.value], "merge": ("none", "peer-kernel", "nccl")[d.value]}

    # ---- rendering
    def render(self, n_spp: int, first_sample: Optional[int] = None, sync: bool = True):
        first = self._next_sample if first_sample is None else first_sample
        _capi.check(self._lib, self._lib.crb_render_samples(self._h, first, n_spp))
        self._next_sample = first + n_spp
        if sync:
            _capi.check(self._lib, self._lib.crb_render_sync(self._h))

    def sync(self):
        _capi.check(self._lib, self._lib.crb_render_sync(self._h))

    # ---- outputs (renderer.cpp:220-238, 386-404)
    def _read(self, kind: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        w, h = self._res
        if out is None:
            out = np.empty((h, w, 4), dtype=np.float32)
        _capi.check(self._lib, self._lib.crb_render_read(self._h, kind, _ptr(out)))
        return out

    def current_progress(self, out=None) -> np.ndarray:
        return self._read(PROGRESS, out)

    def current_progress_async(self, out: np.ndarray, kind: int = PROGRESS) -> int:
        """The reference's UI reads the live buffers while the workers keep rendering (renderer.cpp:220-238):
        queues a snapshot of the buffer after the work submitted so far and its copy into `out` (pinned host
        memory for a truly asynchronous copy); returns a ticket for wait_read(). The next render() can be
        submitted at once, the transfer overlaps it."""
        w, h = self._res
        if out.dtype != np.float32 or out.size != w * h * 4 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous float32 array of h*w*4 elements")
        t = C.c_uint64(0)
        _capi.check(self._lib, self._lib.crb_render_read_async(self._h, kind, _ptr(out), C.byref(t)))
        return int(t.value)

    def wait_read(self, ticket: int):
        _capi.check(self._lib, self._lib.crb_render_read_wait(self._h, ticket))

    def current_normals(self) -> np.ndarray:
        return self._read(NORMAL)

    def current_albedos(self) -> np.ndarray:
        return self._read(ALBEDO)

    def current_depths(self) -> np.ndarray:
        return self._read(DEPTH)

    def raw_sum(self) -> np.ndarray:
        return self._read(RAW_SUM)

    def current_resolution(self):
        return self._res

    def current_sample_count(self) -> int:
        return self.current_stats().passes

    def current_stats(self) -> _capi.Stats:
        st = _capi.Stats()
        _capi.check(self._lib, self._lib.crb_render_stats(self._h, C.byref(st)))
        return st

    def post_process(self, settings: _capi.PostSettings) -> np.ndarray:
        """The post chain applied to the device-resident display buffer (export path, ui.h:567-631)."""
        w, h = self._res
        out = np.empty((h, w, 4), dtype=np.float32)
        _capi.check(self._lib, self._lib.crb_render_post_process(self._h, C.byref(settings), _ptr(out)))
        return out

    # ---- checkpoint / resume
    def checkpoint(self) -> np.ndarray:
        """The accumulation buffer (RGBA, A = pass count): a complete checkpoint of the progressive render."""
        return self.raw_sum()

    def restore(self, raw_sum: np.ndarray):
        a = np.ascontiguousarray(raw_sum, dtype=np.float32)
        passes = int(a[..., 3].max()) if a.size else 0
        _capi.check(self._lib, self._lib.crb_render_restore(self._h, _ptr(a), passes))
        self._next_sample = passes

    # ---- multi-GPU plumbing
    def accum_ptr(self):
        p, n = C.c_void_p(), C.c_uint64(0)
        _capi.check(self._lib, self._lib.crb_render_accum_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_pass_count(self, passes: int):
        _capi.check(self._lib, self._lib.crb_render_set_pass_count(self._h, passes))

    def resolve(self):
        _capi.check(self._lib, self._lib.crb_render_resolve(self._h))

    def stream(self) -> int:
        p = C.c_void_p()
        _capi.check(self._lib, self._lib.crb_render_stream(self._h, C.byref(p)))
        return p.value or 0
