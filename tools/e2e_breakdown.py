#!/usr/bin/env python
"""Where the wall-clock time of one end-to-end step (bench.py e2e) goes: launch, device, read-back."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from crender_b200 import api, scenes

desc = scenes.mesh_scene(1000, 500)
w, h, spp, bounces = 1920, 1080, 16, 8
out = torch.empty((h, w, 4), dtype=torch.float32, pin_memory=True).numpy()
g = api.scene(device=0); scenes.load(desc, g); g.commit()
r = api.renderer(w, h, bounces, g, seed=0); r.render(spp); r.current_progress(out); del r, g
t0 = time.perf_counter()
g = api.scene(device=0); scenes.load(desc, g); info = g.commit()
r = api.renderer(w, h, bounces, g, seed=0)
t1 = time.perf_counter()
rows = []
for k in range(8):
    a = time.perf_counter(); g.set_camera(desc.cam)
    b = time.perf_counter(); r.render(spp, first_sample=k * spp, sync=False)
    c = time.perf_counter(); r.pause()
    d = time.perf_counter(); r.current_progress(out)
    e = time.perf_counter()
    rows.append([(b - a) * 1e3, (c - b) * 1e3, (d - c) * 1e3, (e - d) * 1e3])
st = r.current_stats()
print("setup ms", (t1 - t0) * 1e3, "device_ms total", st.device_ms, "per step", st.device_ms / 8)
for row in rows: print("set_camera %.2f  launch %.2f  wait %.2f  read %.2f  | step %.2f" % (*row, sum(row)))
