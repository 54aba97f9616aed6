"""crender_b200 — B200-native (sm_100a) path-tracing core behind CRender's render-facing API.

The package is a thin host-side mirror of cr::scene / cr::renderer over the C ABI in
include/crender_b200.h; all rendering work is done by hand-written CUDA kernels in
libcrender_b200.so (crender_b200/csrc). There is no CPU path: importing `crender_b200.api` and creating
a scene without the built library or without a CUDA device raises.
"""
from .api import (  # noqa: F401
    ALBEDO, DEPTH, GLASS, HIT_DTYPE, METAL, NORMAL, ORTHOGRAPHIC, PERSPECTIVE, PROGRESS, RAW_SUM, RAY_DTYPE, SMOOTH,
    CrbError, camera, material, model_data, renderer, scene, sun,
)

__all__ = [
    "scene", "renderer", "camera", "material", "sun", "model_data", "CrbError",
    "METAL", "SMOOTH", "GLASS", "PERSPECTIVE", "ORTHOGRAPHIC",
    "RAW_SUM", "PROGRESS", "ALBEDO", "NORMAL", "DEPTH", "RAY_DTYPE", "HIT_DTYPE",
]
