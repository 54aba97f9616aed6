"""Kernel-logic checks on CPU: the product's kernel bodies (BVH build, wide-node collapse, traversal,
wavefront shading) compiled by g++ in CRB_EMU mode (tests/emu) and compared with the oracle. With the same
libm on both sides the comparison is exact. This is a development/test tool only; the GPU parity tests
(test_gpu_parity.py) are the real gate."""
import pytest

import common
import parity_cases as pc
from crender_b200 import scenes


@pytest.mark.parametrize("name", ["cornell", "mesh", "textured", "terrain"])
def test_hits_match_oracle(oracle, emu_lib, name):
    pc.check_hits(oracle, emu_lib, common.small_scenes()[name], n_rays=8000)


@pytest.mark.parametrize("name,w,h,spp,bounces", [("cornell", 40, 40, 3, 8), ("mesh", 48, 32, 2, 5), ("textured", 48, 36, 3, 6), ("terrain", 40, 24, 2, 4)])
def test_images_match_oracle(oracle, emu_lib, name, w, h, spp, bounces):
    desc = common.small_scenes()[name]
    # exact for instanced scenes too: the two-level traversal runs the reference's per-instance arithmetic
    pc.check_image(oracle, emu_lib, desc, w, h, spp, bounces, exact=True)
    pc.check_primary_hits(oracle, emu_lib, desc, w, h, exact=True)


def test_partition_invariance(emu_lib):
    pc.check_partition_invariance(emu_lib, scenes.mesh_scene(24, 12), 30, 20, 4, 4)


def test_edge_cases(emu_lib):
    pc.check_edge_cases(emu_lib)


def test_larger_build(oracle, emu_lib):
    # 80k triangles: exercises deeper trees, many collapse levels and duplicate Morton keys
    pc.check_hits(oracle, emu_lib, scenes.mesh_scene(200, 200, with_ground=False), n_rays=4000)


def test_update_semantics(oracle, emu_lib):
    pc.check_update_semantics(oracle, emu_lib)


def test_checkpoint_resume(emu_lib):
    pc.check_checkpoint_resume(emu_lib)


def test_async_read(emu_lib):
    pc.check_async_read(emu_lib)


def test_post_chain(emu_lib):
    pc.check_post_chain(emu_lib)


def test_material_sort_is_equivalent(emu_lib):
    pc.check_material_sort_is_equivalent(emu_lib)


def test_query_kinds_consistent(emu_lib):
    pc.check_query_kinds_consistent(emu_lib, common.small_scenes()["terrain"], n_rays=6000, seeds=(21,))


def test_instrumented_render_is_identical(emu_lib):
    pc.check_instrumented_render_is_identical(emu_lib)


def test_recycled_memory_is_clean(emu_lib):
    pc.check_recycled_memory_is_clean(emu_lib)


@pytest.mark.parametrize("n", [1, 3])
def test_multi_gpu_handle(emu_lib, n):
    # the multi-GPU orchestration (replicas, spp / tile partition, merge, AOV gather, restore) executed serially
    pc.check_multi_gpu_handle(emu_lib, list(range(n)))


def test_instance_edits(oracle, emu_lib):
    pc.check_instance_edits(oracle, emu_lib)


def test_target_spp(emu_lib):
    pc.check_target_spp(emu_lib)


def test_two_level_edge_cases(oracle, emu_lib):
    pc.check_two_level_edge_cases(oracle, emu_lib)


def test_tiny_models(oracle, emu_lib):
    pc.check_tiny_models(oracle, emu_lib)


def test_instance_wavefront_equals_loop(emu_lib):
    pc.check_instance_wavefront_equals_loop(emu_lib)


def test_collapse_groups(oracle, emu_lib):
    pc.check_collapse_groups(oracle, emu_lib)
