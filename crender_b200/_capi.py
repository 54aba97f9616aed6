"""ctypes declarations for include/crender_b200.h.

The shared library is the product: if it is missing or cannot be loaded this module raises — there is
no Python/CPU fallback of any kind.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libcrender_b200.so")


class Material(C.Structure):
    # cr::material::information, src/render/material/material.h:31-41
    _fields_ = [
        ("shade_type", C.c_uint32),
        ("ior", C.c_float),
        ("roughness", C.c_float),
        ("reflectiveness", C.c_float),
        ("emission", C.c_float),
        ("colour", C.c_float * 4),
        ("tex", C.c_int32),
    ]


class Sun(C.Structure):
    # cr::entity::sun, src/render/entities/components.h:23-29
    _fields_ = [("size", C.c_float), ("intensity", C.c_float), ("direction", C.c_float * 3), ("colour", C.c_float * 3)]


class Camera(C.Structure):
    # cr::camera, src/render/camera.h:11-41
    _fields_ = [
        ("position", C.c_float * 3),
        ("rotation", C.c_float * 3),
        ("fov", C.c_float),
        ("scale", C.c_float),
        ("mode", C.c_uint32),
    ]


class BuildInfo(C.Structure):
    _fields_ = [
        ("build_ms", C.c_double),
        ("upload_ms", C.c_double),
        ("n_triangles", C.c_uint64),
        ("n_nodes", C.c_uint64),
        ("node_bytes", C.c_uint64),
        ("tri_bytes", C.c_uint64),
        ("max_depth", C.c_uint32),
        ("sah_cost", C.c_float),
    ]


class PostSettings(C.Structure):
    # cr::post_processor::*_settings, src/render/post/post_processor.h:19-39
    _fields_ = [
        ("use_bloom", C.c_int32),
        ("bloom_threshold", C.c_float),
        ("bloom_strength", C.c_float),
        ("use_gray_scale", C.c_int32),
        ("use_tonemapping", C.c_int32),
        ("tonemapping_type", C.c_int32),
        ("tonemapping_exposure", C.c_float),
        ("gamma_correction", C.c_float),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("total_queries", C.c_uint64),
        ("ref_rays", C.c_uint64),
        ("pixel_samples", C.c_uint64),
        ("passes", C.c_uint64),
        ("device_ms", C.c_double),
        ("kernel_launches", C.c_uint64),
        ("node_visits", C.c_uint64 * 2),
        ("tri_tests", C.c_uint64 * 2),
        ("closest_queries", C.c_uint64),
        ("shadow_queries", C.c_uint64),
        ("kernel_ms", C.c_double * 8),
        ("kernel_count", C.c_uint64 * 8),
        ("running_time", C.c_double),
        ("rays_per_second", C.c_double),
        ("samples_per_second", C.c_double),
    ]


# every symbol include/crender_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "crb_last_error": (C.c_char_p, []),
    "crb_set_device": (C.c_int, [C.c_int]),
    "crb_device_info": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "crb_scene_create": (C.c_int, [C.POINTER(_P)]),
    "crb_scene_destroy": (C.c_int, [_P]),
    "crb_scene_add_mesh": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.POINTER(C.c_int)]),
    "crb_scene_set_materials": (C.c_int, [_P, C.c_int, C.POINTER(Material), C.c_uint32]),
    "crb_scene_set_instances": (C.c_int, [_P, C.c_int, _P, C.c_uint32]),
    "crb_scene_add_texture": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]),
    "crb_scene_set_sun": (C.c_int, [_P, C.POINTER(Sun), C.c_int]),
    "crb_scene_set_skybox": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_float, C.c_float]),
    "crb_scene_set_camera": (C.c_int, [_P, C.POINTER(Camera)]),
    "crb_scene_commit": (C.c_int, [_P, C.POINTER(BuildInfo)]),
    "crb_scene_set_option": (C.c_int, [_P, C.c_int, C.c_int]),
    "crb_intersect_batch": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int]),
    "crb_occluded_batch": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int]),
    "crb_trace_counters": (C.c_int, [_P, _P, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "crb_last_query_ms": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "crb_microbench_read": (C.c_int, [_P, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "crb_scene_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "crb_render_create": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "crb_render_destroy": (C.c_int, [_P]),
    "crb_render_reset": (C.c_int, [_P]),
    "crb_render_set_resolution": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "crb_render_set_max_bounces": (C.c_int, [_P, C.c_uint32]),
    "crb_render_refresh": (C.c_int, [_P]),
    "crb_render_set_rows": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "crb_render_samples": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "crb_render_sync": (C.c_int, [_P]),
    "crb_render_read": (C.c_int, [_P, C.c_int, _P]),
    "crb_render_read_async": (C.c_int, [_P, C.c_int, _P, C.POINTER(C.c_uint64)]),
    "crb_render_read_wait": (C.c_int, [_P, C.c_uint64]),
    "crb_render_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "crb_render_restore": (C.c_int, [_P, _P, C.c_uint32]),
    "crb_post_process": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.POINTER(PostSettings), _P]),
    "crb_render_post_process": (C.c_int, [_P, C.POINTER(PostSettings), _P]),
    "crb_render_accum_ptr": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "crb_render_set_pass_count": (C.c_int, [_P, C.c_uint32]),
    "crb_render_resolve": (C.c_int, [_P]),
    "crb_render_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "crb_render_set_bands": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32]),
    "crb_render_set_bands_ordered": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "crb_render_create_multi": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "crb_comm_unique_id": (C.c_int, [_P]),
    "crb_render_create_rank": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "crb_render_set_sample_table": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32]),
    "crb_render_set_target_spp": (C.c_int, [_P, C.c_uint64]),
    "crb_render_run": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint64)]),
    "crb_render_flush": (C.c_int, [_P]),
    "crb_render_join_flush": (C.c_int, [_P]),
    "crb_render_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}


class CrbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"crender_b200 error {code}: {message}")
        self.code = code


_libs: dict[str, C.CDLL] = {}


def load(path: str | None = None) -> C.CDLL:
    """Loads libcrender_b200.so (built by `__graft_entry__.build()` / csrc/Makefile) and binds every symbol."""
    path = os.path.abspath(path or os.environ.get("CRENDER_B200_LIB", DEFAULT_LIB))
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). crender_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib


def check(lib: C.CDLL, code: int) -> None:
    if code != 0:
        raise CrbError(code, lib.crb_last_error().decode("utf-8", "replace"))
