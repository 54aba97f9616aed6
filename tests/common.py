"""Shared helpers for the parity tests."""
import numpy as np

from crender_b200 import scenes

MISS = 0xFFFFFFFF


def relrmse(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b**2)) + 1e-30))


def hit_agreement(a, b, abs_slack=0.0):
    """Fraction of rays whose (model, inst, prim) agree, and max relative |dt| on the agreeing hits.
    abs_slack: absolute distance forgiven before the relative error is taken (a few float ulps at the
    scene's coordinate magnitude; needed when an instance transform moves the arithmetic to a different
    coordinate frame and t is much smaller than the coordinates)."""
    same = (a["prim"] == b["prim"]) & (a["model"] == b["model"]) & (a["inst"] == b["inst"])
    hit = same & (a["prim"] != MISS)
    if hit.any():
        d = np.maximum(np.abs(a["t"][hit].astype(np.float64) - b["t"][hit]) - abs_slack, 0.0)
        dt = d / np.maximum(np.abs(b["t"][hit]), 1e-30)
        return float(same.mean()), float(dt.max())
    return float(same.mean()), 0.0


def mixed_rays(desc, n, seed=11):
    """Random rays through the inflated scene box plus rays aimed at the scene centre from outside."""
    lo, hi = desc.aabb()
    a = scenes.random_rays(lo, hi, n // 2, seed=seed)
    rng = np.random.Generator(np.random.Philox(seed + 1))
    m = n - n // 2
    c, e = 0.5 * (lo + hi), 0.5 * (hi - lo)
    o = c + (rng.random((m, 3)) * 2 - 1) * e * 3.0
    tgt = c + (rng.random((m, 3)) * 2 - 1) * e * 0.9
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = np.empty(m, dtype=a.dtype)
    b["o"], b["d"], b["tmin"], b["tmax"] = o.astype(np.float32), d.astype(np.float32), np.float32(1e-5), np.float32(np.inf)
    return np.concatenate([a, b])


def small_scenes():
    return {
        "cornell": scenes.cornell(),
        "mesh": scenes.mesh_scene(60, 30),
        "textured": scenes.textured_scene(),
        "terrain": scenes.terrain_city(24, 2, n_buildings=12),
    }
