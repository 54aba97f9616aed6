#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs and for the roofline context. One JSON object
per line on stdout; results are copied into profiles/ per round. (bench.py stays the headline contract.)

  python tools/bench_configs.py rays   [--n 100000000]   config 3: incoherent closest-hit + occlusion batch on the 1M-tri BVH
  python tools/bench_configs.py build                     BVH build time (repeated commits) at 1M / 2M / 18M triangles
  python tools/bench_configs.py c1                        config 1: Cornell 512x512, 64 spp, depth 8
  python tools/bench_configs.py c4     [--spp 64]         config 4 geometry: 18M flattened triangles (instanced terrain+city), 1080p
  python tools/bench_configs.py c5     [--spp 16]         config 5 geometry: 1M mesh + 64 emitters, 3840x2160
  (c1/c2/c4/c5 take --extended: GGX / Fresnel / area-light NEE shading, and --sort: material sort before shading)
  python tools/bench_configs.py mem                       L2 / HBM read bandwidth micro-benchmark
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crender_b200 import api, scenes  # noqa: E402


def device_rays(lo, hi, n, seed, occlusion=False):
    """Config-3 ray batch generated on the device (torch is plumbing here: device memory + RNG)."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    lo_t, hi_t = torch.tensor(lo, device="cuda", dtype=torch.float64), torch.tensor(hi, device="cuda", dtype=torch.float64)
    c, e = 0.5 * (lo_t + hi_t), 0.5 * (hi_t - lo_t) * 1.5
    rays = torch.empty((n, 8), device="cuda", dtype=torch.float32)
    rays[:, 0:3] = (c + (torch.rand((n, 3), device="cuda", generator=g, dtype=torch.float32).double() * 2 - 1) * e).float()
    u = torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32)
    ct = 2 * u[:, 0] - 1
    st = torch.sqrt(torch.clamp(1 - ct * ct, min=0))
    ph = 2 * np.pi * u[:, 1]
    rays[:, 4], rays[:, 5], rays[:, 6] = st * torch.cos(ph), ct, st * torch.sin(ph)
    rays[:, 3] = 1e-5
    if occlusion:
        rays[:, 7] = torch.rand(n, device="cuda", generator=g) * float(np.linalg.norm(np.asarray(hi) - np.asarray(lo)))
    else:
        rays[:, 7] = float("inf")
    return rays


def cmd_rays(a):
    import torch

    desc = scenes.mesh_scene(1000, 500, with_ground=not a.no_ground)
    g = api.scene()
    scenes.load(desc, g)
    info = g.commit()
    lo, hi = desc.aabb()
    n = a.n
    out = {"config": "3: incoherent ray batch on the 1M-triangle BVH" + (" (mesh only: rays drawn in the mesh's own box)" if a.no_ground else " (scene box incl. the 16x16 ground quad)"), "rays": n, "triangles": int(info.n_triangles), "bvh_nodes": int(info.n_nodes)}
    sub = scenes.random_rays(lo, hi, 2_000_000, seed=2)
    for any_hit, name in ((False, "closest"), (True, "occluded")):
        nn, nt = g.trace_counters(sub if not any_hit else scenes.random_rays(lo, hi, 2_000_000, seed=2, occlusion=True), any_hit=any_hit)
        out[f"{name}_nodes_per_ray"], out[f"{name}_tris_per_ray"] = nn / len(sub), nt / len(sub)
    rays = device_rays(lo, hi, n, 2)
    hits = torch.empty((n, 6), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    for _ in range(2):
        g.cast_rays_device(rays.data_ptr(), hits.data_ptr(), n)
    ms = g.last_query_ms()
    out["closest_mrays_s"], out["closest_ms"] = n / ms / 1e3, ms
    out["hit_fraction"] = float((hits[:, 0] != float("inf")).float().mean().item())
    b = 32 + 24 + out["closest_nodes_per_ray"] * 80 + out["closest_tris_per_ray"] * 48
    out["closest_algorithmic_GBs"] = n * b / (ms * 1e-3) / 1e9
    del hits
    rays = device_rays(lo, hi, n, 3, occlusion=True)
    occ = torch.empty(n, device="cuda", dtype=torch.uint8)
    for _ in range(2):
        g.occluded_device(rays.data_ptr(), occ.data_ptr(), n)
    ms = g.last_query_ms()
    out["occluded_mrays_s"], out["occluded_ms"] = n / ms / 1e3, ms
    out["occluded_fraction"] = float(occ.float().mean().item())
    print(json.dumps(out), flush=True)


def cmd_build(a):
    for name, desc in (("1M mesh", scenes.mesh_scene(1000, 500)), ("2M heightfield tile", scenes.terrain_city(1000, 1, n_buildings=8)),
                       ("18M flattened (9 instances of 2M terrain + city)", scenes.terrain_city(1000, 3))):
        g = api.scene()
        scenes.load(desc, g)
        times, up = [], []
        for _ in range(4):
            g.set_instances(0, desc.meshes[0].instances if desc.meshes[0].instances is not None else np.eye(4, dtype=np.float32)[None])  # forces a rebuild
            info = g.commit()
            times.append(info.build_ms), up.append(info.upload_ms)
        print(json.dumps({"scene": name, "triangles": int(info.n_triangles), "nodes": int(info.n_nodes), "depth": int(info.max_depth), "sah": info.sah_cost,
                          "build_ms_first": times[0], "build_ms_min": min(times[1:]), "upload_ms_min": min(up[1:]),
                          "mtris_per_s": info.n_triangles / min(times[1:]) / 1e3}), flush=True)
        del g


def render_rate(desc, w, h, bounces, spp, label, warm=4, extended=False, material_sort=False):
    g = api.scene()
    scenes.load(desc, g)
    info = g.commit()
    r = api.renderer(w, h, bounces, g, seed=0, extended=extended, material_sort=material_sort, timers=extended)
    r.render(warm)
    r.start()
    t0 = time.perf_counter()
    r.render(spp)
    wall = time.perf_counter() - t0
    st = r.current_stats()
    print(json.dumps({"config": label, "triangles": int(info.n_triangles), "nodes": int(info.n_nodes), "build_ms": info.build_ms, "res": [w, h], "spp": spp,
                      "bounces": bounces, "device_ms": st.device_ms, "wall_ms": wall * 1e3, "mrays_s": st.total_queries / st.device_ms / 1e3,
                      "samples_per_s": st.pixel_samples / (st.device_ms * 1e-3), "rays_per_sample": st.total_queries / st.pixel_samples,
                      "bvh_bytes": int(info.node_bytes + info.tri_bytes), "shading": "extended" if extended else "ref-exact", "material_sort": bool(material_sort),
                      **({"kernel_ms": {k: round(st.kernel_ms[i], 3) for i, k in enumerate(("raygen", "trace", "shade", "shadow", "advance", "accumulate"))}} if extended else {})}),
          flush=True)


def cmd_mem(a):
    g = api.scene()
    g.commit()
    out = {}
    for mb in (8, 32, 64, 96, 256, 4096):
        iters = max(2, int(16384 / mb))
        out[f"{mb}MB"] = g.microbench_read(mb << 20, iters)
    print(json.dumps({"read_GBs_by_working_set": out, "note": "16-byte loads, grid-stride, 148x8 CTAs of 256; <=96 MB is L2-resident"}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["rays", "build", "c1", "c2", "c4", "c5", "mem"])
    ap.add_argument("--extended", action="store_true", help="render configs: the extended shading mode (GGX / Fresnel / area-light NEE)")
    ap.add_argument("--sort", action="store_true", help="render configs: material sort before shading")
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--no-ground", action="store_true", help="rays: drop the ground quad so that the ray box is the mesh's own box")
    a = ap.parse_args()
    if a.cmd == "rays":
        cmd_rays(a)
    elif a.cmd == "build":
        cmd_build(a)
    elif a.cmd == "c1":
        render_rate(scenes.cornell(), 512, 512, 8, a.spp or 64, "1: Cornell 512x512 64 spp depth 8", extended=a.extended, material_sort=a.sort)
    elif a.cmd == "c2":
        render_rate(scenes.mesh_scene(1000, 500), 1920, 1080, 8, a.spp or 32, "2: 1M-triangle mesh, 1080p", extended=a.extended, material_sort=a.sort)
    elif a.cmd == "c4":
        render_rate(scenes.terrain_city(1000, 3), 1920, 1080, 8, a.spp or 64, "4 (geometry): 18M flattened triangles, 1080p", extended=a.extended, material_sort=a.sort)
    elif a.cmd == "c5":
        render_rate(scenes.lights_scene(), 3840, 2160, 8, a.spp or 16, "5: 1M mesh + 64 emitters, 3840x2160" if a.extended else "5 (geometry): 1M mesh + 64 emitters, 3840x2160",
                    extended=a.extended, material_sort=a.sort)
    else:
        cmd_mem(a)
