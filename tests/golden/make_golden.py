#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz from the CPU oracle.

The reference ships no golden vectors and cannot be compiled here, so these fixtures are
outputs of the ORACLE on seeded procedural inputs, generated in this container by this script. They pin the
oracle against regressions (tests/test_golden.py, CPU) and give the CUDA path a second, frozen target besides
the live oracle comparison (tests/test_gpu_parity.py::test_against_golden_fixtures).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
import oracle_binding as ob  # noqa: E402
from crender_b200 import scenes  # noqa: E402

N_RAYS = 2048
IMAGES = {"cornell": (32, 32, 4, 8, 3), "mesh": (48, 32, 2, 5, 3), "textured": (40, 30, 3, 6, 3), "terrain": (40, 24, 2, 4, 3)}  # w,h,spp,bounces,seed
EXT_IMAGES = {"cornell": (32, 32, 4, 8, 3), "textured": (40, 30, 3, 6, 3), "lights": (48, 32, 3, 6, 3)}


def build(name):
    desc = common.small_scenes()[name]
    s = ob.scene()
    scenes.load(desc, s)
    s.commit()
    return desc, s


def generate():
    out = {}
    for name in ("cornell", "mesh", "textured", "terrain"):
        desc, s = build(name)
        rays = common.mixed_rays(desc, N_RAYS, seed=101)
        h = s.cast_rays(rays, nthreads=1)
        for f in ("t", "u", "v", "prim", "model", "inst"):
            out[f"{name}/hit_{f}"] = h[f].copy()
        out[f"{name}/occluded"] = s.occluded(rays, nthreads=1)
        w, hh, spp, bounces, seed = IMAGES[name]
        r = ob.renderer(w, hh, bounces, s, seed=seed)
        r.render(spp, nthreads=1)
        out[f"{name}/raw"] = r.raw_sum()
        out[f"{name}/progress"] = r.current_progress()
        out[f"{name}/albedo"] = r.current_albedos()
        out[f"{name}/normal"] = r.current_normals()
        out[f"{name}/depth"] = r.current_depths()
        st = r.current_stats()
        out[f"{name}/stats"] = np.asarray([st.total_queries, st.ref_rays, st.pixel_samples, st.passes], np.uint64)
    # extended shading mode (oracle-specified, see oracle.cpp "EXTENDED shading mode")
    for name in EXT_IMAGES:
        desc = scenes.lights_scene(40, 20, n_lights=8) if name == "lights" else common.small_scenes()[name]
        s = ob.scene()
        scenes.load(desc, s)
        s.commit()
        w, hh, spp, bounces, seed = EXT_IMAGES[name]
        r = ob.renderer(w, hh, bounces, s, seed=seed, extended=True)
        r.render(spp, nthreads=1)
        out[f"{name}/ext_raw"] = r.raw_sum()
    L = ob.lib()
    out["rng"] = np.asarray([L.orc_kat_rng(s_, p, k, d) for s_, p, k, d in [(0, 0, 0, 0), (0, 0, 0, 1), (3, 12345, 7, 5), (9, 2073599, 255, 33)]], np.float32)
    return out


if __name__ == "__main__":
    data = generate()
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "arrays")
