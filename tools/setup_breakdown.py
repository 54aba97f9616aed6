#!/usr/bin/env python
"""Where the one-off set-up time of the public API goes (scene creation, host->device upload, BVH build,
renderer creation, first render call) for BASELINE config 2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from crender_b200 import api, scenes


class A: nu, nv, width, height, bounces = 1000, 500, 1920, 1080, 8
desc = scenes.mesh_scene(A.nu, A.nv)
for rnd in range(3):
    t = [time.perf_counter()]
    g = api.scene(); t.append(time.perf_counter())
    scenes.load(desc, g); t.append(time.perf_counter())
    info = g.commit(); t.append(time.perf_counter())
    r = api.renderer(A.width, A.height, A.bounces, g, seed=0); t.append(time.perf_counter())
    r.render(16); r.sync(); t.append(time.perf_counter())
    r.render(16, first_sample=16); r.sync(); t.append(time.perf_counter())
    names = ["scene()", "load (python + add_mesh copies)", "commit", "renderer()", "first render(16)", "second render(16)"]
    print("round", rnd, {n: round((t[i + 1] - t[i]) * 1e3, 2) for i, n in enumerate(names)}, "upload_ms", round(info.upload_ms, 2), "build_ms", round(info.build_ms, 2))
    del r, g
