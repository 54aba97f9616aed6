#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): builds a 20k-triangle
scene, traces a ray batch both ways, renders a few samples with every optional flag, runs the post chain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import common
from crender_b200 import api, scenes, _capi

# (the textured scene is instanced: two-level traversal by default; CRB_FLATTEN=1 in the environment runs the flattened path)
for desc in (scenes.mesh_scene(100, 100), scenes.textured_scene(), scenes.terrain_city(24, 2, n_buildings=12)):
    g = api.scene(); scenes.load(desc, g); info = g.commit()
    rays = common.mixed_rays(desc, 20000, seed=3)
    h = g.cast_rays(rays); o = g.occluded(rays)
    assert ((h["prim"] != 0xffffffff) == o.astype(bool)).all()
    for kw in ({}, {"counters": True, "timers": True}, {"material_sort": True}):
        r = api.renderer(96, 64, 4, g, seed=1, **kw); r.render(2); s = r.raw_sum(); r.close()
    s = _capi.PostSettings(); s.use_bloom = 1; s.bloom_threshold = 0.5; s.bloom_strength = 0.5; s.use_tonemapping = 1; s.tonemapping_type = 3; s.tonemapping_exposure = 1.0; s.gamma_correction = 2.2
    g.post_process(np.random.default_rng(0).random((64, 96, 4), dtype=np.float32), s)
    # the multi-GPU handle on one GPU: snapshot, merge kernel, merged reads, both partitions
    for part in (api.PARTITION_SPP, api.PARTITION_TILE):
        m = api.renderer(96, 150, 4, g, seed=1, gpus=[0], partition=part); m.render(2, sync=False); s2 = m.raw_sum(); a = m.current_albedos(); m.close()
    print("ok", info.n_nodes)
