"""Tree-quality counters WITHOUT a GPU (and, per ray class, where the node visits go): the kernel-logic harness (tests/emu, the product's own builder and instrumented
traversal compiled for the CPU) on config 2 at reduced resolution. Usage: [ENV knobs] python tools/tree_quality.py [nu nv w h spp]
Prints nodes / triangles per closest-hit and shadow query (the GPU's counters on the full-size frame agree to ~1 %)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crender_b200 import api, scenes

a = [int(x) for x in sys.argv[1:]]
nu, nv, w, h, spp = (a + [1000, 500, 192, 108, 2][len(a):])[:5]
lib = os.path.join(ROOT, "tests", "emu", "_build", "libcrb_emu.so")
desc = scenes.mesh_scene(nu, nv) if os.environ.get("TQ_SCENE", "c2") == "c2" else scenes.terrain_city(int(os.environ.get("TQ_TERRAIN", "200")), 3)
g = api.scene(lib_path=lib)
scenes.load(desc, g)
t0 = time.time()
info = g.commit()
t1 = time.time()
r = api.renderer(w, h, 8, g, seed=1, counters=True)
r.render(spp)
s = r.current_stats()
cq, sq = s.total_queries - s.shadow_queries, s.shadow_queries
print("nodes %d  build %.1fs  closest %.3f nodes %.3f tris | shadow %.3f nodes %.3f tris | queries %d + %d" % (
    info.n_nodes, t1 - t0, s.node_visits[0] / max(cq, 1), s.tri_tests[0] / max(cq, 1), s.node_visits[1] / max(sq, 1), s.tri_tests[1] / max(sq, 1), cq, sq))

# per-ray classes (harness only: crb_emu_ray_classes, render.cu RayClassProbe)
import ctypes as C
emu = C.CDLL(lib)
if hasattr(emu, "crb_emu_ray_classes"):
    buf = (C.c_ulonglong * 16)()
    emu.crb_emu_ray_classes(buf, 0)
    for kind, kname in enumerate(("closest-hit", "shadow     ")):
        tot = buf[kind * 8 + 0] + buf[kind * 8 + 4]
        for outcome, oname in enumerate(("nothing hit", "hit        ")):
            n, nodes, tris, le1 = (buf[kind * 8 + outcome * 4 + j] for j in range(4))
            if n:
                print("  %s rays, %s: %5.1f %% of the kind, %6.2f node visits, %5.2f triangle tests per ray, %4.1f %% of them done after at most one node" % (
                    kname, oname, 100.0 * n / max(tot, 1), nodes / n, tris / n, 100.0 * le1 / n))
