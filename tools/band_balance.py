#!/usr/bin/env python
"""Load balance of the tile partition on ONE GPU: renders each of the 8 ranks' interleaved row bands of the config-5 frame
separately and prints max/mean of the device times per band height (the strong-scaling efficiency the partition can reach).
    python tools/band_balance.py [spp]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from crender_b200 import api, scenes

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
world = 8
desc = scenes.lights_scene(1000, 500, n_lights=64)
g = api.scene(); scenes.load(desc, g); g.commit()
r = api.renderer(3840, 2160, 8, g, seed=0, extended=True)
for band, serp in ((64, False), (16, False), (16, True), (8, False), (8, True), (4, True)):
    ms = []
    for rank in range(world):
        r.start(); r.set_bands(band, rank, world, serpentine=serp); r.render(4, first_sample=0); r.pause()   # warm
        r.start(); r.set_bands(band, rank, world, serpentine=serp)
        s0 = r.current_stats().device_ms
        r.render(spp, first_sample=0); r.pause()
        ms.append(r.current_stats().device_ms - s0)
    ms = np.asarray(ms)
    print("band %2d rows%s: per-rank ms %s  max/mean %.4f -> efficiency bound %.4f" % (band, " serpentine" if serp else "", np.round(ms, 1), ms.max() / ms.mean(), ms.mean() / ms.max()), flush=True)
