#!/usr/bin/env python
"""Generates tests/golden/reference_v1.npz: outputs of THE REFERENCE ITSELF.

The reference's own sources (renderer.cpp, scene.cpp, model.cpp, registry.cpp, camera.cpp, ... compiled unmodified by
oracle/ref/Makefile against shim headers for the three absent third-party libraries glm / fmt / embree3) are run in THIS
container on the seeded procedural scenes below, with a one-thread pool so that its default-seeded std::mt19937
(renderer.cpp:6-11) is consumed in call order. /root/reference and the compiled binary do not travel to the GPU box; these
fixtures do. Consumers: tests/test_reference_anchor.py (CPU: the oracle in reference-stream mode must reproduce every
array bit for bit; GPU: the CUDA path replaying the same numbers through crb_render_set_sample_table must match the
reference's images within the north-star tolerance).

    python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
import ref_binding as rb  # noqa: E402
from crender_b200 import api, scenes  # noqa: E402

# name -> (w, h, bounces, spp)
CASES = {
    "cornell": (64, 64, 8, 8),
    "mesh": (64, 36, 8, 4),
    "textured": (64, 48, 6, 6),
    "terrain": (64, 36, 5, 4),
    "cornell_rotated": (48, 32, 5, 3),
    "cornell_ortho": (48, 32, 5, 3),
    "mesh_sky": (48, 28, 6, 3),
    "general_instances": (64, 48, 6, 4),
}


def scene_of(name):
    if name == "cornell_rotated":
        d = scenes.cornell()
        d.cam = api.camera(position=(0.2, 0.1, -3.0), fov=50.0, rotation=(6.0, -4.0, 10.0))
        return d
    if name == "cornell_ortho":
        d = scenes.cornell()
        d.cam = api.camera(position=(0.0, 0.0, -3.0), current_mode=api.ORTHOGRAPHIC, scale=1.1)
        return d
    if name == "mesh_sky":
        d = scenes.mesh_scene(60, 30, sun_enabled=False)  # lit by an environment map only
        d.skybox = np.random.RandomState(9).uniform(0.0, 2.0, (8, 16, 4)).astype(np.float32)
        d.skybox_rotation = (0.2, 0.1)
        return d
    if name == "general_instances":
        # instance transforms that are not rigid: non-uniform scale, rotation about a skew axis, shear (exercises
        # glm::inverse's full cofactor arithmetic, the per-instance renormalisation and the world-space re-measuring of
        # model.cpp:107-120)
        d = scenes.textured_scene()
        rs = np.random.RandomState(4)
        inst = []
        for k in range(4):
            A = np.eye(4, dtype=np.float64)
            A[:3, :3] = np.diag(rs.uniform(0.6, 1.6, 3)) @ np.linalg.qr(rs.normal(size=(3, 3)))[0] + rs.uniform(-0.15, 0.15, (3, 3))
            A[:3, 3] = (rs.uniform(-1.5, 1.5), rs.uniform(0.2, 0.6), rs.uniform(0.8, 2.4))
            inst.append(A.T.astype(np.float32))  # column-major storage m[c][r]
        d.meshes[2].instances = np.stack(inst)
        return d
    return common.small_scenes()[name]


def generate():
    out = {}
    for name, (w, h, bounces, spp) in CASES.items():
        ref = rb.render(scene_of(name), w, h, bounces, spp)
        assert ref["samples"] == spp
        out[f"{name}/params"] = np.asarray([w, h, bounces, spp], np.int64)
        out[f"{name}/draws_before"] = np.asarray([ref["draws_before"]], np.uint64)
        out[f"{name}/total_rays"] = np.asarray([ref["total_rays"]], np.uint64)
        for k in ("raw", "progress", "albedo", "normal", "depth"):
            out[f"{name}/{k}"] = ref[k]
        print(f"{name}: {w}x{h} {spp} spp depth {bounces}: draws_before {ref['draws_before']} total_rays {ref['total_rays']} mean {ref['raw'].mean():.5f}")
    return out


if __name__ == "__main__":
    rb.build()
    data = generate()
    path = os.path.join(HERE, "reference_v1.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")
