"""The oracle's own BVH is only trusted because an O(T) brute-force loop over every triangle (same
triangle test) gives identical answers. These tests pin that, plus the edge cases of the query path
('what replaces Embree in the oracle')."""
import numpy as np
import pytest

import common
from crender_b200 import scenes
from crender_b200.api import RAY_DTYPE, material


@pytest.mark.parametrize("name", ["cornell", "mesh", "textured", "terrain"])
def test_bvh_equals_brute_force(oracle, name):
    desc = common.small_scenes()[name]
    s = oracle.scene()
    scenes.load(desc, s)
    s.commit()
    rays = common.mixed_rays(desc, 6000)
    a, b = s.cast_rays(rays), s.cast_rays(rays, brute=True)
    for f in ("t", "u", "v", "prim", "model", "inst"):
        np.testing.assert_array_equal(a[f], b[f], err_msg=f)
    assert (a["prim"] != common.MISS).mean() > 0.2
    occ = s.occluded(rays)
    np.testing.assert_array_equal(occ.astype(bool), a["prim"] != common.MISS)


def _rays(o, d, tmin=1e-5, tmax=np.inf):
    o, d = np.atleast_2d(np.asarray(o, np.float32)), np.atleast_2d(np.asarray(d, np.float32))
    r = np.empty(len(o), dtype=RAY_DTYPE)
    r["o"], r["d"], r["tmin"], r["tmax"] = o, d, tmin, tmax
    return r


def test_empty_scene_and_single_triangle(oracle):
    s = oracle.scene()
    s.commit()
    h = s.cast_rays(_rays([0, 0, 0], [0, 0, 1]))
    assert h["prim"][0] == common.MISS and np.isinf(h["t"][0])
    s = oracle.scene()
    s.add_mesh(np.asarray([[[0, 0, 1], [1, 0, 1], [0, 1, 1]]], np.float32))
    s.commit()
    h = s.cast_rays(_rays([[0.2, 0.2, 0], [2, 2, 0]], [[0, 0, 1], [0, 0, 1]]))
    assert h["prim"][0] == 0 and h["t"][0] == 1.0 and h["prim"][1] == common.MISS


def test_coincident_triangles_tie_goes_to_lowest_prim(oracle):
    tri = np.asarray([[[0, 0, 1], [1, 0, 1], [0, 1, 1]]], np.float32)
    s = oracle.scene()
    s.add_mesh(np.repeat(tri, 40, axis=0))
    s.commit()
    h = s.cast_rays(_rays([0.2, 0.2, 0], [0, 0, 1]))
    assert h["prim"][0] == 0
    # two models with the same triangle: the first model wins the tie (scene.cpp:79-98 uses strict <)
    s = oracle.scene()
    s.add_mesh(tri)
    s.add_mesh(tri)
    s.commit()
    h = s.cast_rays(_rays([0.2, 0.2, 0], [0, 0, 1]))
    assert h["model"][0] == 0


def test_tmin_tmax_window(oracle):
    s = oracle.scene()
    s.add_mesh(np.asarray([[[0, 0, 1], [1, 0, 1], [0, 1, 1]], [[0, 0, 2], [1, 0, 2], [0, 1, 2]]], np.float32))
    s.commit()
    o, d = [0.2, 0.2, 0], [0, 0, 1]
    assert s.cast_rays(_rays(o, d))["prim"][0] == 0
    assert s.cast_rays(_rays(o, d, tmin=1.0))["prim"][0] == 1  # t == tnear is rejected
    assert s.cast_rays(_rays(o, d, tmin=1.0, tmax=2.0))["prim"][0] == 1  # t == tfar is accepted
    assert s.cast_rays(_rays(o, d, tmin=1.0, tmax=1.5))["prim"][0] == common.MISS
    # un-normalised direction: t is in units of |d|
    assert s.cast_rays(_rays(o, [0, 0, 4]))["t"][0] == 0.25


def test_instances_follow_reference_semantics(oracle):
    # model.cpp:99-126: the ray is taken into object space per instance; nearest instance wins
    tri = np.asarray([[[0, 0, 1], [1, 0, 1], [0, 1, 1]]], np.float32)
    s = oracle.scene()
    m = s.add_mesh(tri)
    s.set_materials(m, [material()])
    s.set_instances(m, np.stack([scenes.translation(0, 0, 5), scenes.translation(0, 0, 2), scenes.compose(scenes.translation(0, 0, 3), scenes.rotation_y(0.0))]))
    s.commit()
    h = s.cast_rays(_rays([0.2, 0.2, 0], [0, 0, 1]))
    assert h["inst"][0] == 1 and abs(h["t"][0] - 3.0) < 1e-6
