"""Parity cases shared by the CPU kernel-logic harness (tests/emu, exact comparison: same libm) and the
GPU tests (through libcrender_b200.so; tolerances from BASELINE.json north_star: primary-hit primitive
ids agree on >= 99.99 % of rays with |dt|/t <= 1e-5, images within 1 % relative RMSE at equal spp and
seeds)."""
import ctypes as C
import os

import numpy as np
import pytest

import common
from crender_b200 import _capi, api, scenes
from crender_b200.api import RAY_DTYPE, material

PRIM_AGREE = 0.9999
T_REL = 1e-5
IMG_RELRMSE = 0.01


def build_pair(oracle, lib_path, desc):
    o = oracle.scene()
    scenes.load(desc, o)
    o.commit()
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    info = g.commit()
    assert info.n_triangles == desc.n_flat_tris
    return o, g


def check_hits(oracle, lib_path, desc, n_rays=20000, exact=False):
    o, g = build_pair(oracle, lib_path, desc)
    rays = common.mixed_rays(desc, n_rays)
    ho, hg = o.cast_rays(rays), g.cast_rays(rays)
    identity_only = all(m.instances is None for m in desc.meshes)
    lo, hi = desc.aabb()
    # instanced models: the oracle intersects in object space, the product in world space -> a few ulps
    # at the coordinate magnitude on top of the 1e-5 relative bound
    slack = 0.0 if identity_only else 8 * float(np.finfo(np.float32).eps) * float(max(np.abs(lo).max(), np.abs(hi).max()) * 3.0)
    agree, dt = common.hit_agreement(hg, ho, slack)
    flattened = os.environ.get("CRB_FLATTEN", "0") not in ("", "0")
    if exact or identity_only or not flattened:
        # same triangle test on the same coordinates -> bit-identical t,u,v,prim: identity instances, and — two-level
        # traversal, the default — instanced models in their own object space like the reference (model.cpp:107-112)
        for f in ("t", "u", "v", "prim", "model", "inst"):
            np.testing.assert_array_equal(hg[f], ho[f], err_msg=f)
    else:
        assert agree >= PRIM_AGREE, agree
        assert dt <= T_REL, dt
    # any-hit must agree with closest-hit existence, also with a finite tmax window
    occ = g.occluded(rays)
    np.testing.assert_array_equal(occ.astype(bool), hg["prim"] != common.MISS)
    r2 = scenes.random_rays(lo, hi, n_rays // 2, seed=5, occlusion=True)
    h2, occ2 = g.cast_rays(r2), g.occluded(r2)
    np.testing.assert_array_equal(occ2.astype(bool), h2["prim"] != common.MISS)
    ho2 = o.cast_rays(r2)
    assert common.hit_agreement(h2, ho2)[0] >= PRIM_AGREE
    return agree, dt


def render_pair(oracle, lib_path, desc, w, h, spp, bounces, seed=3, extended=False):
    o, g = build_pair(oracle, lib_path, desc)
    ro = oracle.renderer(w, h, bounces, o, seed=seed, extended=extended)
    rg = api.renderer(w, h, bounces, g, seed=seed, extended=extended)
    ro.render(spp)
    rg.render(spp)
    return ro, rg


def check_primary_hits(oracle, lib_path, desc, w, h, exact=False):
    """BASELINE north_star: primary-ray hit primitive ids agree on >= 99.99 % of rays, t within 1e-5 rel.
    The GPU side is observed through the depth/normal/albedo AOVs of a 1-bounce render (first-hit data,
    renderer.cpp:303-308) and through the batch query fed with the oracle's own camera rays."""
    ro, rg = render_pair(oracle, lib_path, desc, w, h, 1, 1)
    for name, a, b in (("depth", rg.current_depths(), ro.current_depths()), ("normal", rg.current_normals(), ro.current_normals()),
                       ("albedo", rg.current_albedos(), ro.current_albedos())):
        same = np.all(np.abs(a - b) <= 1e-5 * np.maximum(1.0, np.abs(b)), axis=-1)
        if exact:
            np.testing.assert_array_equal(a, b, err_msg=name)
        assert same.mean() >= PRIM_AGREE, (name, same.mean())
    ph = ro.primary_hits(0)
    return ph


def check_image(oracle, lib_path, desc, w, h, spp, bounces, exact=False, extended=False):
    ro, rg = render_pair(oracle, lib_path, desc, w, h, spp, bounces, extended=extended)
    a, b = rg.raw_sum(), ro.raw_sum()
    np.testing.assert_array_equal(a[..., 3], b[..., 3])  # pass count
    err = common.relrmse(a[..., :3], b[..., :3])
    if exact:
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(rg.current_progress(), ro.current_progress())
    assert err <= IMG_RELRMSE, err
    pa, pb = rg.current_progress(), ro.current_progress()
    assert common.relrmse(pa[..., :3], pb[..., :3]) <= IMG_RELRMSE
    np.testing.assert_array_equal(pa[..., 3], pb[..., 3])  # alpha 1 (image.h:123-126)
    so, sg = ro.current_stats(), rg.current_stats()
    assert sg.passes == so.passes == spp
    assert sg.pixel_samples == so.pixel_samples == w * h * spp
    # the reference's ray counter (renderer.cpp:271-272,356): path segments, +1 per path that ran out
    assert abs(int(sg.ref_rays) - int(so.ref_rays)) <= max(2, int(1e-4 * so.ref_rays)), (sg.ref_rays, so.ref_rays)
    return err


def check_partition_invariance(lib_path, desc, w, h, spp, bounces):
    """Size-independent properties: splitting the sample range or the rows must not change a single bit of
    the accumulated sums (the sampler is keyed by global pixel and sample index; accumulation is in
    sample order)."""
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    r = api.renderer(w, h, bounces, g, seed=9)
    r.render(spp)
    whole = r.raw_sum().copy()
    disp = r.current_progress().copy()
    r.start()
    r.render(spp // 2, first_sample=0)
    r.render(spp - spp // 2, first_sample=spp // 2)
    np.testing.assert_array_equal(r.raw_sum(), whole)
    np.testing.assert_array_equal(r.current_progress(), disp)
    r.start()
    cut = h // 3
    r.set_rows(0, cut)
    r.render(spp, first_sample=0)
    r.set_rows(cut, h)
    r.render(spp, first_sample=0)
    # the pass count is per pixel (accum.w), so the bands need no manual set_pass_count / resolve, and a RAW_SUM
    # read after a banded render is a complete checkpoint (alpha == spp everywhere)
    np.testing.assert_array_equal(r.raw_sum(), whole)
    np.testing.assert_array_equal(r.current_progress(), disp)
    assert r.current_stats().passes == spp
    ck = r.checkpoint().copy()
    r2 = api.renderer(w, h, bounces, g, seed=9)
    r2.restore(ck)
    assert r2.current_stats().passes == spp
    r2.render(2)
    r.set_rows(0, h)
    r.render(2, first_sample=spp)
    np.testing.assert_array_equal(r2.raw_sum(), r.raw_sum())
    np.testing.assert_array_equal(r2.current_progress(), r.current_progress())
    del r2
    # interleaved bands in one launch sequence (what a tile-partition rank renders): 3 "ranks" cover the frame
    r.start()
    for first in range(3):
        r.set_bands(7, first, 3)
        r.render(spp, first_sample=0)
    np.testing.assert_array_equal(r.raw_sum(), whole)
    np.testing.assert_array_equal(r.current_progress(), disp)
    # the same with the serpentine owner order of the library's own tile partition (every other period reversed), for rank
    # counts that leave the last period partial and band heights that leave the last band partial
    for band, world in ((7, 3), (5, 4), (h, 2), (1, 5)):
        r.start()
        for first in range(world):
            r.set_bands(band, first, world, serpentine=True)
            r.render(spp, first_sample=0)
        np.testing.assert_array_equal(r.raw_sum(), whole, err_msg=f"serpentine bands {band} x {world}")
    # linearity of the accumulation: rendering spp twice from the same first sample doubles the sum
    r.set_rows(0, h)
    r.start()
    r.render(spp, first_sample=0)
    r.render(spp, first_sample=0)
    np.testing.assert_allclose(r.raw_sum()[..., :3], 2 * whole[..., :3], rtol=1e-6, atol=1e-6)
    return whole


def check_edge_cases(lib_path):
    L = _capi.load(lib_path)
    # empty scene renders black (no skybox) without error
    g = api.scene(lib_path=lib_path)
    g.commit()
    r = api.renderer(8, 6, 3, g)
    r.render(2)
    assert np.all(r.raw_sum()[..., :3] == 0)
    assert np.all(r.current_progress()[..., :3] == 0) and np.all(r.current_progress()[..., 3] == 1)
    rays = np.zeros(3, dtype=RAY_DTYPE)
    rays["d"] = (0, 0, 1)
    rays["tmax"] = np.inf
    h = g.cast_rays(rays)
    assert np.all(h["prim"] == common.MISS) and np.all(np.isinf(h["t"]))
    assert g.cast_rays(rays[:0]).shape == (0,)
    # one triangle / two triangles / coincident duplicates (tie -> lowest prim)
    tri = np.asarray([[[0, 0, 1], [1, 0, 1], [0, 1, 1]]], np.float32)
    for reps in (1, 2, 40):
        g = api.scene(lib_path=lib_path)
        g.add_mesh(np.repeat(tri, reps, axis=0))
        g.commit()
        rays = np.zeros(2, dtype=RAY_DTYPE)
        rays["o"] = [(0.2, 0.2, 0), (2, 2, 0)]
        rays["d"] = (0, 0, 1)
        rays["tmin"], rays["tmax"] = 1e-5, np.inf
        h = g.cast_rays(rays)
        assert h["prim"][0] == 0 and h["t"][0] == 1.0 and h["prim"][1] == common.MISS
    # rendering before commit is an error, not a crash; so are bad ids
    g = api.scene(lib_path=lib_path)
    g.add_mesh(tri)
    r = api.renderer(4, 4, 2, g)
    try:
        r.render(1)
        raise AssertionError("render on an uncommitted scene must fail")
    except api.CrbError as e:
        assert e.code == 33
    for bad in (lambda: g.set_materials(5, [material()]), lambda: g.set_materials(0, []), lambda: g.set_instances(-1, np.eye(4))):
        try:
            bad()
            raise AssertionError("expected CrbError")
        except api.CrbError as e:
            assert e.code == 2
    assert b"" != L.crb_last_error()


def check_first_hit_via_batch(oracle, lib_path, desc, w, h):
    """Feeds the oracle's primary rays (same camera, same jitter) to the batch query and compares ids."""
    o, g = build_pair(oracle, lib_path, desc)
    ro = oracle.renderer(w, h, 1, o, seed=3)
    ph = ro.primary_hits(0)
    return ph, g


def check_update_semantics(oracle, lib_path):
    """renderer::update = pause -> mutate -> restart from sample 0 (renderer.cpp:185-192); material edits
    need no rebuild (ui.h:924-929), instance edits do (ui.h:1183-1192); set_resolution / set_max_bounces
    (renderer.cpp:194-213); un-rendered pixels keep cr::image's FLT_MAX fill (image.h:30-38)."""
    desc = scenes.textured_scene()
    o, g = build_pair(oracle, lib_path, desc)
    w, h = 64, 48
    rg = api.renderer(w, h, 5, g, seed=2)
    rg.render(3)
    before = rg.raw_sum().copy()

    # 1. material edit through update(): restarts from sample 0, no re-commit
    new_mats = [material(api.METAL, colour=(0.2, 0.9, 0.3, 1.0), reflectiveness=0.7), material(api.SMOOTH, colour=(0.9, 0.1, 0.1, 1.0)), material(api.GLASS, ior=1.3)]
    nodes_before = g.build_info.n_nodes
    rg.update(lambda: g.set_materials(2, new_mats))
    assert rg.current_stats().passes == 0
    rg.render(3)
    after = rg.raw_sum().copy()
    assert not np.array_equal(before, after)
    o.set_materials(2, new_mats)
    ro = oracle.renderer(w, h, 5, o, seed=2)
    ro.render(3)
    assert common.relrmse(after[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE
    assert g.build_info.n_nodes == nodes_before

    # 2. instance edit: needs a commit (the rtcCommitScene point); render before commit is an error
    inst = np.stack([scenes.translation(-0.8, 0.0, 1.0), scenes.compose(scenes.translation(0.9, 0.1, 1.6), scenes.rotation_y(-20.0))]).astype(np.float32)
    g.set_instances(2, inst)
    try:
        rg.start()
        rg.render(1)
        raise AssertionError("render after an instance edit without commit must fail")
    except api.CrbError as e:
        assert e.code == 33
    g.commit()
    rg.start()
    rg.render(3)
    o.set_instances(2, inst)
    o.commit()
    ro = oracle.renderer(w, h, 5, o, seed=2)
    ro.render(3)
    assert common.relrmse(rg.raw_sum()[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE

    # 3. set_max_bounces / set_resolution
    rg.set_max_bounces(2)
    rg.start()
    rg.render(2)
    ro = oracle.renderer(w, h, 2, o, seed=2)
    ro.render(2)
    assert common.relrmse(rg.raw_sum()[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE
    rg.set_resolution(40, 30)
    assert rg.current_resolution() == (40, 30)
    rg.render(2)
    ro = oracle.renderer(40, 30, 2, o, seed=2)
    ro.render(2)
    assert rg.raw_sum().shape == (30, 40, 4)
    assert common.relrmse(rg.raw_sum()[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE

    # 4. FLT_MAX fill of pixels never rendered, sun toggle, camera change picked up through update()
    rg.start()
    rg.set_rows(0, 10)
    rg.render(1)
    disp = rg.current_progress()
    flt_max = np.float32(3.4028234663852886e38)
    assert np.all(disp[:20] == flt_max) and np.all(disp[20:, :, 3] == 1.0)  # sample rows 0..9 land in flipped rows 20..29
    rg.set_rows(0, 30)
    cam2 = api.camera(position=(0.3, 1.0, -3.0), fov=60.0, rotation=(-3.0, 10.0, 0.0))

    def mutate():
        g.set_sun_enabled(False)
        g.set_camera(cam2)

    rg.update(mutate)
    rg.render(2)
    o.set_sun_enabled(False)
    o.set_camera(cam2)
    ro = oracle.renderer(40, 30, 2, o, seed=2)
    ro.render(2)
    assert common.relrmse(rg.raw_sum()[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE
    assert rg.current_stats().total_queries == rg.current_stats().closest_queries  # sun off: no shadow rays


def check_checkpoint_resume(lib_path):
    """accumulator checkpoint/resume. Rendering 8 passes in one go and rendering 5,
    checkpointing, restoring into a NEW renderer and rendering 3 more must agree bit for bit."""
    desc = scenes.mesh_scene(40, 20)
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    r = api.renderer(72, 40, 5, g, seed=6)
    r.render(8)
    whole_raw, whole_disp = r.raw_sum().copy(), r.current_progress().copy()
    r2 = api.renderer(72, 40, 5, g, seed=6)
    r2.render(5)
    ckpt = r2.checkpoint().copy()
    del r2
    r3 = api.renderer(72, 40, 5, g, seed=6)
    r3.restore(ckpt)
    assert r3.current_stats().passes == 5
    r3.render(3)  # continues at first_sample = 5
    np.testing.assert_array_equal(r3.raw_sum(), whole_raw)
    np.testing.assert_array_equal(r3.current_progress(), whole_disp)


def check_async_read(lib_path):
    """crb_render_read_async / crb_render_read_wait: a read queued between two render calls returns the image
    as it was after the FIRST (snapshot in stream order), bit-identical to the blocking read, while the second
    call is already submitted; tickets older than the event ring are still waitable; bad tickets fail."""
    desc = scenes.mesh_scene(40, 20)
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    ref = api.renderer(96, 54, 5, g, seed=9)
    want = []
    for k in range(12):
        ref.render(2)
        want.append((ref.current_progress().copy(), ref.raw_sum().copy()))
    r = api.renderer(96, 54, 5, g, seed=9)
    outs = [np.zeros((54, 96, 4), np.float32) for _ in range(12)]
    raws = [np.zeros((54, 96, 4), np.float32) for _ in range(12)]
    tickets = []
    for k in range(12):
        r.render(2, sync=False)
        tickets.append((r.current_progress_async(outs[k]), r.current_progress_async(raws[k], kind=api.RAW_SUM)))
    r.wait_read(tickets[0][0])  # 24 reads issued: older than the 8-deep event ring
    np.testing.assert_array_equal(outs[0], want[0][0])
    for k in range(12):
        r.wait_read(tickets[k][1])
        np.testing.assert_array_equal(outs[k], want[k][0])
        np.testing.assert_array_equal(raws[k], want[k][1])
    assert tickets[-1][1] == 23
    with pytest.raises(Exception):
        r.wait_read(24)
    with pytest.raises(ValueError):
        r.current_progress_async(np.zeros((10, 10, 4), np.float32))
    # sync() covers outstanding reads; a resolution change afterwards re-sizes the staging buffer
    r.render(1, sync=False)
    t = r.current_progress_async(outs[0])
    r.sync()
    r.set_resolution(48, 27)
    r.start()
    r.render(2)
    small = np.zeros((27, 48, 4), np.float32)
    r.wait_read(r.current_progress_async(small))
    np.testing.assert_array_equal(small, r.current_progress())


def check_post_chain(lib_path):
    """the export-time post chain as CUDA kernels against the numpy restatement of the
    reference's GLSL (tests/post_oracle.py). Tolerance 2e-5 relative: powf / GLSL pow differ by ulps."""
    import post_oracle

    g = api.scene(lib_path=lib_path)
    g.commit()
    rs = np.random.RandomState(3)
    for (w, h) in ((32, 32), (40, 27), (24, 48)):  # square, not multiples of 8, taller than wide
        img = rs.uniform(0.0, 1.4, (h, w, 4)).astype(np.float32)
        img[..., 3] = 1.0
        cases = [dict(), dict(use_gray_scale=True), dict(use_bloom=True, bloom_threshold=0.8, bloom_strength=0.6)]
        cases += [dict(use_tonemapping=True, tonemapping_type=t, tonemapping_exposure=1.3, gamma_correction=2.2) for t in range(4)]
        cases += [dict(use_bloom=True, use_gray_scale=True, use_tonemapping=True, tonemapping_type=1)]
        for kw in cases:
            got = g.post_process(img, api.post_settings(**kw))
            ref = post_oracle.process(img, **kw)
            np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6, err_msg=str((w, h, kw)))
    # the renderer-side entry point works on the device-resident display buffer
    desc = scenes.cornell()
    gs = api.scene(lib_path=lib_path)
    scenes.load(desc, gs)
    gs.commit()
    r = api.renderer(48, 40, 4, gs, seed=1)
    r.render(2)
    kw = dict(use_tonemapping=True, tonemapping_type=3, use_bloom=True)
    np.testing.assert_allclose(r.post_process(api.post_settings(**kw)), post_oracle.process(r.current_progress(), **kw), rtol=2e-5, atol=2e-6)


def check_material_sort_is_equivalent(lib_path):
    """The optional material sort (k_classify) only reorders work: accumulated sums are bit-identical."""
    desc = scenes.textured_scene()
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    a = api.renderer(80, 60, 6, g, seed=4, material_sort=False)
    b = api.renderer(80, 60, 6, g, seed=4, material_sort=True)
    a.render(4)
    b.render(4)
    np.testing.assert_array_equal(a.raw_sum(), b.raw_sum())
    np.testing.assert_array_equal(a.current_normals(), b.current_normals())
    sa, sb = a.current_stats(), b.current_stats()
    assert sa.total_queries == sb.total_queries and sa.ref_rays == sb.ref_rays
    assert sb.kernel_launches > sa.kernel_launches  # the sort is an extra pass per bounce


def check_extended_properties(lib_path):
    """Extended mode: it really is a different shading path, the material sort only reorders work, splitting the
    sample range or the rows changes no bit, and the light list follows material edits without a rebuild."""
    desc = scenes.lights_scene(40, 20, n_lights=6)
    g = api.scene(lib_path=lib_path)
    ids = scenes.load(desc, g)
    g.commit()
    w, h, spp, bounces = 48, 32, 4, 6
    ref = api.renderer(w, h, bounces, g, seed=4)
    ref.render(spp)
    a = api.renderer(w, h, bounces, g, seed=4, extended=True)
    a.render(spp)
    whole = a.raw_sum().copy()
    assert common.relrmse(whole[..., :3], ref.raw_sum()[..., :3]) > 0.05
    np.testing.assert_array_equal(a.current_normals(), ref.current_normals())  # AOVs are the same first hits
    b = api.renderer(w, h, bounces, g, seed=4, extended=True, material_sort=True)
    b.render(spp)
    np.testing.assert_array_equal(b.raw_sum(), whole)
    a.start()
    a.render(1, first_sample=0)
    a.render(spp - 1, first_sample=1)
    np.testing.assert_array_equal(a.raw_sum(), whole)
    a.start()
    a.set_rows(0, 11)
    a.render(spp, first_sample=0)
    a.set_rows(11, h)
    a.render(spp, first_sample=0)
    np.testing.assert_array_equal(a.raw_sum()[..., :3], whole[..., :3])
    # switching the emitters off through a material edit empties the light list: darker image, no rebuild
    a.set_rows(0, h)
    mats = [material(api.SMOOTH, colour=(1, 1, 1, 1), emission=0.0, name="emitter-off")]

    def mutate():
        g.set_materials(ids[-1], mats)

    a.update(mutate)
    a.render(spp, first_sample=0)
    assert a.raw_sum()[..., :3].mean() < whole[..., :3].mean()
    # roughness is live in extended mode (dead in the reference)
    rough = []
    for r_ in (0.05, 0.9):
        d2 = scenes.mesh_scene(24, 12)
        for m in d2.meshes:
            for mm in m.materials:
                mm.roughness = r_
        g2 = api.scene(lib_path=lib_path)
        scenes.load(d2, g2)
        g2.commit()
        rr = api.renderer(w, h, bounces, g2, seed=4, extended=True)
        rr.render(2)
        rough.append(rr.raw_sum().copy())
    assert not np.array_equal(rough[0], rough[1])


def check_query_kinds_consistent(lib_path, desc, n_rays=400000, seeds=(21, 22, 23)):
    """Any-hit and closest-hit queries are separate kernel instantiations of the same traversal loop: for
    every ray `occluded` must equal `closest hit exists`. 400k rays x 3 seeds per scene: the ptxas
    miscompile recorded in DESIGN.md section 4 affected 5e-5 .. 6e-3 of the rays of ONE instantiation, so
    small ray counts do not find this class of error. Also checks the counting instantiations: in total, stopping at
    the first hit visits less than finding the closest one; per ray, see below."""
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    for seed in seeds:
        rays = common.mixed_rays(desc, n_rays, seed=seed)
        hit = g.cast_rays(rays)["prim"] != common.MISS
        occ = g.occluded(rays).astype(bool)
        assert int((occ != hit).sum()) == 0, (seed, int((occ != hit).sum()))
    sub = rays[:: max(1, n_rays // 20000)]
    nc, tc = g.trace_counters(sub, any_hit=False)
    na, ta = g.trace_counters(sub, any_hit=True)
    assert 0 < na <= nc and 0 < ta <= tc, (na, nc, ta, tc)
    # per ray: any-hit queries visit a node's hit children back to front (bvh8.cuh CRB_ANY_BACK_FIRST), closest-hit queries
    # front to back, so for a ray that hits something neither count bounds the other; a ray that hits NOTHING is never
    # pruned and never stops early: both kinds must visit exactly the same nodes and triangles
    for i in np.nonzero(~hit[:: max(1, n_rays // 20000)])[0][:16]:
        c, a = g.trace_counters(sub[i : i + 1], any_hit=False), g.trace_counters(sub[i : i + 1], any_hit=True)
        assert a[0] == c[0] and a[1] == c[1] and a[0] >= 1, (int(i), a, c)


def check_instrumented_render_is_identical(lib_path):
    """The counting / timing instantiations of the wavefront kernels (k_trace<true>, k_shadow<true>) must
    produce the same image, bit for bit, as the plain ones, and their query counts must match."""
    desc = scenes.terrain_city(24, 2, n_buildings=12)
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    a = api.renderer(160, 90, 5, g, seed=9)
    b = api.renderer(160, 90, 5, g, seed=9, counters=True, timers=True)
    a.render(6)
    b.render(6)
    np.testing.assert_array_equal(a.raw_sum(), b.raw_sum())
    np.testing.assert_array_equal(a.current_depths(), b.current_depths())
    sb = b.current_stats()
    assert sb.total_queries > 0 and sb.node_visits[0] > 0 and sb.tri_tests[0] > 0


RENDER_HASH_SNIPPET = r"""
import sys, hashlib
sys.path.insert(0, {root!r})
from crender_b200 import api, scenes
desc = scenes.terrain_city(24, 2, n_buildings=12)
g = api.scene(lib_path={lib!r}); scenes.load(desc, g); g.commit()
r = api.renderer(160, 90, 5, g, seed=9); r.render(6)
print("HASH", hashlib.sha256(r.raw_sum().tobytes()).hexdigest())
"""


def check_trace_steps_variants_identical(lib_path):
    """k_trace / k_shadow are instantiated for 1, 2, 4 and 8 traversal steps per refill check (tuning knob
    CRB_TRACE_STEPS, read once per process): every instantiation must give the same accumulated sums."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = RENDER_HASH_SNIPPET.format(root=root, lib=lib_path)
    hashes = {}
    for steps in ("1", "2", "4", "8"):
        env = dict(os.environ, CRB_TRACE_STEPS=steps)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        hashes[steps] = [ln for ln in out.stdout.splitlines() if ln.startswith("HASH")][0]
    assert len(set(hashes.values())) == 1, hashes


def check_recycled_memory_is_clean(lib_path):
    """Device blocks are recycled between handles (platform.cuh DevBlockCache): a renderer that inherits the
    path state and image buffers of a destroyed one must not see any of their contents."""

    def run(desc, seed):
        g = api.scene(lib_path=lib_path)
        scenes.load(desc, g)
        g.commit()
        r = api.renderer(160, 90, 5, g, seed=seed)
        r.render(4)
        out = (r.raw_sum().copy(), r.current_normals().copy(), r.current_depths().copy())
        r.close()
        g.close()
        return out

    a0 = run(scenes.terrain_city(24, 2, n_buildings=12), 1)
    run(scenes.textured_scene(), 2)  # same buffer sizes, different contents
    run(scenes.cornell(), 3)
    a1 = run(scenes.terrain_city(24, 2, n_buildings=12), 1)
    for x, y in zip(a0, a1):
        np.testing.assert_array_equal(x, y)

    # the host copies of the caller's arrays are recycled too (scene.cuh HostBlockCache, blocks of 1 MB and more):
    # a scene that inherits the blocks of a destroyed one of the same size must see only its own geometry
    def hits(desc, rays):
        g = api.scene(lib_path=lib_path)
        scenes.load(desc, g)
        g.commit()
        h = g.cast_rays(rays).copy()
        g.close()
        return h

    da, db = scenes.mesh_scene(200, 80, seed=1), scenes.mesh_scene(200, 80, seed=5)  # 32000 triangles: 1.15 MB of vertices each
    assert da.meshes[0].verts.nbytes >= 1 << 20 and da.meshes[0].verts.shape == db.meshes[0].verts.shape
    lo, hi = da.aabb()
    rays = scenes.random_rays(lo, hi, 4000, seed=9)
    h0 = hits(da, rays)
    h1 = hits(db, rays)
    h2 = hits(da, rays)
    for k in ("t", "prim", "model", "inst"):
        np.testing.assert_array_equal(h0[k], h2[k])
    assert (h0["prim"] != h1["prim"]).any()  # the second scene really was different



def check_multi_gpu_handle(lib_path, gpus, w=72, h=150, spp=6, bounces=5):
    """crb_render_create_multi: one process, the devices in `gpus` (the kernel-logic harness ignores the ids: the
    replicas, the partition and the merge run serially on the host). The tile partition must merge bit-identically
    to a single renderer (row bands are disjoint and per-pixel sample order is unchanged), the spp partition up to
    float summation order across ranks; AOVs, statistics, checkpoint/resume, resolution and scene edits follow."""
    desc = scenes.textured_scene()
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    single = api.renderer(w, h, bounces, g, seed=11)
    single.render(spp)
    want_raw, want_disp = single.raw_sum().copy(), single.current_progress().copy()
    want_aov = [single.current_albedos().copy(), single.current_normals().copy(), single.current_depths().copy()]
    want_q = single.current_stats().total_queries
    n = len(gpus)
    for part in (api.PARTITION_TILE, api.PARTITION_SPP):
        m = api.renderer(w, h, bounces, g, seed=11, gpus=gpus, partition=part)
        info = m.info()
        assert info["gpus_local"] == n and info["ranks"] == n and info["partition"] == ("spp", "tile")[part]
        assert info["merge"] in (("none",) if n == 1 else ("peer-kernel", "nccl"))
        # progressive: two calls, a read in between (flush + merge), one more call
        m.render(spp // 2, sync=False)
        half = m.raw_sum().copy()
        assert np.all(half[..., 3] == spp // 2)
        m.render(spp - spp // 2, sync=False)
        raw, disp = m.raw_sum().copy(), m.current_progress().copy()
        np.testing.assert_array_equal(raw[..., 3], want_raw[..., 3])
        if part == api.PARTITION_TILE or n == 1:
            np.testing.assert_array_equal(raw, want_raw)
            np.testing.assert_array_equal(disp, want_disp)
        else:
            assert common.relrmse(raw[..., :3], want_raw[..., :3]) < 1e-6
            np.testing.assert_allclose(disp, want_disp, rtol=0, atol=1e-5)
        for got, want in zip((m.current_albedos(), m.current_normals(), m.current_depths()), want_aov):
            np.testing.assert_array_equal(got, want)
        st = m.current_stats()
        assert st.passes == spp and st.pixel_samples == w * h * spp and st.total_queries == want_q
        # asynchronous merged read, then checkpoint -> restore into a fresh multi handle -> continue == single
        out = np.zeros((h, w, 4), np.float32)
        m.wait_read(m.current_progress_async(out))
        np.testing.assert_array_equal(out, disp)
        ck = m.checkpoint().copy()
        m2 = api.renderer(w, h, bounces, g, seed=11, gpus=gpus, partition=part)
        m2.restore(ck)
        assert m2.current_stats().passes == spp
        m2.render(2, first_sample=spp)
        single.render(2, first_sample=spp)
        if part == api.PARTITION_TILE or n == 1:
            np.testing.assert_array_equal(m2.raw_sum(), single.raw_sum())
        else:
            assert common.relrmse(m2.raw_sum()[..., :3], single.raw_sum()[..., :3]) < 1e-6
        del m2
        single.start()
        single.render(spp)
        # scene edits reach the replicas through update(): camera (light state) and instances (re-commit everywhere)
        cam2 = api.camera(position=(0.2, 1.1, -3.0), fov=60.0, rotation=(-2.0, 9.0, 0.0))
        inst = np.stack([scenes.translation(-0.8, 0.0, 1.0), scenes.compose(scenes.translation(0.9, 0.1, 1.6), scenes.rotation_y(-20.0))]).astype(np.float32)

        def mutate():
            g.set_camera(cam2)
            g.set_instances(2, inst)
            g.commit()

        m.update(mutate)
        m.render(3)
        s2 = api.renderer(w, h, bounces, g, seed=11)
        s2.render(3)
        if part == api.PARTITION_TILE or n == 1:
            np.testing.assert_array_equal(m.raw_sum(), s2.raw_sum())
        else:
            assert common.relrmse(m.raw_sum()[..., :3], s2.raw_sum()[..., :3]) < 1e-6
        m.set_resolution(40, 70)
        m.render(2)
        s2.set_resolution(40, 70)
        s2.render(2)
        if part == api.PARTITION_TILE or n == 1:
            np.testing.assert_array_equal(m.current_progress(), s2.current_progress())
        else:
            np.testing.assert_allclose(m.current_progress(), s2.current_progress(), rtol=0, atol=1e-5)
        # single-GPU-only calls are refused on a multi handle, with a message
        with pytest.raises(api.CrbError):
            m.set_rows(0, 10)
        del m, s2
        # restore the scene for the next partition
        g.set_camera(desc.cam)
        g.set_instances(2, desc.meshes[2].instances)
        g.commit()
        single = api.renderer(w, h, bounces, g, seed=11)
        single.render(spp)


def check_instance_edits(oracle, lib_path):
    """The reference edits instance transforms without touching Embree (src/ui/ui.h:1183-1192: the transforms are applied
    per ray, model.cpp:107-112). Two-level scenes: set_instances + commit keeps every BLAS and rebuilds the TLAS only, the
    hits stay bit-identical to the oracle for rigid, scaled and sheared transforms, identity and empty instance lists
    included; the flattened alternative agrees within the north-star tolerance."""
    desc = scenes.terrain_city(24, 2, n_buildings=12)
    o, g = build_pair(oracle, lib_path, desc)
    nodes0 = g.build_info.n_nodes
    rs = np.random.RandomState(8)
    rays = common.mixed_rays(desc, 20000, seed=17)
    for trial in range(4):
        inst = []
        for k in range(1 + trial):
            A = np.eye(4, dtype=np.float64)
            if trial >= 2:
                A[:3, :3] = np.diag(rs.uniform(0.7, 1.4, 3)) @ np.linalg.qr(rs.normal(size=(3, 3)))[0] + rs.uniform(-0.1, 0.1, (3, 3))
            else:
                A[:3, :3] = scenes.rotation_y(rs.uniform(0, 90))[:3, :3].T
            A[:3, 3] = rs.uniform(-1.5, 1.5, 3) * (1, 0.2, 1)
            inst.append(A.T.astype(np.float32))
        inst = np.stack(inst)
        g.set_instances(1, inst)
        info = g.commit()
        o.set_instances(1, inst)
        o.commit()
        assert info.n_triangles == desc.meshes[0].verts.shape[0] * len(desc.meshes[0].instances) + desc.meshes[1].verts.shape[0] * len(inst)
        ho, hg = o.cast_rays(rays), g.cast_rays(rays)
        for f in ("prim", "model", "inst", "t", "u", "v"):
            np.testing.assert_array_equal(hg[f], ho[f], err_msg=f"trial {trial} {f}")
        np.testing.assert_array_equal(g.occluded(rays).astype(bool), hg["prim"] != common.MISS)
    assert abs(int(g.build_info.n_nodes) - int(nodes0)) <= 4  # only the top level changed size
    # images through the renderer: exact on the CPU harness is checked by the callers' image tests; here vs the oracle
    w, h = 64, 40
    rg = api.renderer(w, h, 5, g, seed=5)
    rg.render(3)
    ro = oracle.renderer(w, h, 5, o, seed=5)
    ro.render(3)
    assert common.relrmse(rg.raw_sum()[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE
    # the flattened alternative on the same scene
    g.set_flatten_instances(True)
    g.commit()
    hf = g.cast_rays(rays)
    lo, hi = desc.aabb()
    slack = 8 * float(np.finfo(np.float32).eps) * float(max(np.abs(lo).max(), np.abs(hi).max()) * 3.0)
    agree, dt = common.hit_agreement(hf, ho, slack)
    assert agree >= PRIM_AGREE and dt <= T_REL, (agree, dt)
    rf = api.renderer(w, h, 5, g, seed=5)
    rf.render(3)
    assert common.relrmse(rf.raw_sum()[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE
    g.set_flatten_instances(False)
    g.commit()
    np.testing.assert_array_equal(g.cast_rays(rays)["t"], ho["t"])


def check_target_spp(lib_path):
    """renderer::set_target_spp + start (renderer.cpp:116-144,154-170,215-218): the library's own loop renders exactly the
    missing passes up to the target, in calls of 16; the result is the same render as one call; the reference-style
    statistics (passes per second, path segments per second) are filled."""
    desc = scenes.mesh_scene(40, 20)
    g = api.scene(lib_path=lib_path)
    scenes.load(desc, g)
    g.commit()
    r = api.renderer(64, 40, 5, g, seed=2)
    r.render(37)
    want = r.raw_sum().copy()
    r2 = api.renderer(64, 40, 5, g, seed=2)
    r2.set_target_spp(37)
    r2.start()
    st = r2.current_stats()
    assert st.passes == 37 and r2.current_sample_count() == 37
    np.testing.assert_array_equal(r2.raw_sum(), want)
    r2.start()  # restart: renders to the target again from sample 0
    np.testing.assert_array_equal(r2.raw_sum(), want)
    if st.device_ms > 0:
        assert st.running_time > 0 and abs(st.samples_per_second - 37 / st.running_time) < 1e-6 * st.samples_per_second
        assert abs(st.rays_per_second - st.ref_rays / st.running_time) < 1e-6 * st.rays_per_second


def check_two_level_edge_cases(oracle, lib_path):
    """Instanced scenes at the edges: a model with an EMPTY instance list is invisible, a thousand instances of a small
    model (a TLAS several levels deep), an instance list mixing the identity with other transforms, coincident instances
    (ties go to the first in the reference's loop order), all against the oracle, bit for bit."""
    box = scenes._box((0, 0, 0), (0.2, 0.2, 0.2), 15.0).astype(np.float32)
    ground = scenes._quad((-30, -0.5, -30), (-30, -0.5, 30), (30, -0.5, 30), (30, -0.5, -30)).astype(np.float32)
    rs = np.random.RandomState(3)

    def build(instances):
        desc = scenes.SceneDesc("edge", [scenes.MeshDesc(ground, np.zeros(2, np.uint32), [material()], name="ground"),
                                         scenes.MeshDesc(box, np.zeros(12, np.uint32), [material()], instances=instances, name="box")],
                                api.camera(position=(0.0, 3.0, -12.0), fov=60.0, rotation=(0.0, 12.0, 0.0)))
        return desc

    cases = {
        "thousand": np.stack([scenes.compose(scenes.translation(rs.uniform(-20, 20), rs.uniform(0, 3), rs.uniform(-20, 20)), scenes.rotation_y(rs.uniform(0, 360))) for _ in range(1000)]),
        "identity_mixed": np.stack([np.eye(4, dtype=np.float32), scenes.translation(1.0, 0.5, 0.0), np.eye(4, dtype=np.float32)]),  # instances 0 and 2 coincide
        "none": np.zeros((0, 4, 4), np.float32),
    }
    for name, inst in cases.items():
        desc = build(inst.astype(np.float32))
        o = oracle.scene()
        scenes.load(desc, o)
        o.commit()
        g = api.scene(lib_path=lib_path)
        scenes.load(desc, g)
        info = g.commit()
        assert info.n_triangles == 2 + 12 * len(inst), name
        rays = common.mixed_rays(desc, 6000, seed=23)
        ho, hg = o.cast_rays(rays), g.cast_rays(rays)
        for f in ("prim", "model", "inst", "t", "u", "v"):
            np.testing.assert_array_equal(hg[f], ho[f], err_msg=f"{name} {f}")
        if name == "none":
            assert not np.any(hg["model"] == 1)
        np.testing.assert_array_equal(g.occluded(rays).astype(bool), hg["prim"] != common.MISS)
        w, h = 48, 30
        rg = api.renderer(w, h, 4, g, seed=1)
        rg.render(2)
        ro = oracle.renderer(w, h, 4, o, seed=1)
        ro.render(2)
        assert common.relrmse(rg.raw_sum()[..., :3], ro.raw_sum()[..., :3]) <= IMG_RELRMSE, name


def check_tiny_models(oracle, lib_path):
    """Models of 1..8 triangles take the single-node build (bvh_build.cu k_tiny_bvh), 9 and more the full pipeline: hits of
    random small triangle soups against the oracle, every field equal; as a flat scene and as an instanced model."""
    rs = np.random.RandomState(5)
    for n in (1, 2, 3, 5, 7, 8, 9, 17):
        tris = (rs.uniform(-1, 1, (n, 1, 3)) + rs.uniform(-0.6, 0.6, (n, 3, 3))).astype(np.float32)
        for inst in (None, np.stack([np.eye(4, dtype=np.float32), scenes.translation(2.5, 0.0, 0.5)]).astype(np.float32)):
            desc = scenes.SceneDesc(f"tiny{n}", [scenes.MeshDesc(tris, np.zeros(n, np.uint32), [material()], instances=inst, name="soup")],
                                    api.camera(position=(0.0, 0.0, -6.0), fov=50.0))
            o, g = build_pair(oracle, lib_path, desc)
            rays = common.mixed_rays(desc, 4000, seed=31 + n)
            ho, hg = o.cast_rays(rays), g.cast_rays(rays)
            assert (ho["prim"] != common.MISS).sum() > 50, n
            for f in ("prim", "model", "inst", "t", "u", "v"):
                np.testing.assert_array_equal(hg[f], ho[f], err_msg=f"n={n} {f}")
            np.testing.assert_array_equal(g.occluded(rays).astype(bool), hg["prim"] != common.MISS)


def check_collapse_groups(oracle, lib_path):
    """The wide-tree collapse queues its levels in groups without reading anything back (bvh_build.cu, COLLAPSE_GROUP = 12 levels;
    no tree of the test scenes is deeper than one group): with groups of 1 and 3 levels the continuation after a group's
    read-back runs too. Same tree statistics, same hits as the default and as the oracle."""
    desc = scenes.mesh_scene(60, 40, seed=3)
    rays = common.mixed_rays(desc, 6000, seed=17)
    ref, infos = None, []
    old = os.environ.get("CRB_COLLAPSE_GROUP")
    try:
        for grp in (None, "1", "3"):
            if grp is None:
                os.environ.pop("CRB_COLLAPSE_GROUP", None)
            else:
                os.environ["CRB_COLLAPSE_GROUP"] = grp
            o, g = build_pair(oracle, lib_path, desc)
            info = g.build_info
            infos.append((int(info.n_nodes), int(info.max_depth), int(info.n_triangles)))
            hg = g.cast_rays(rays)
            if ref is None:
                ref = hg
                ho = o.cast_rays(rays)
                for f in ("prim", "t", "u", "v"):
                    np.testing.assert_array_equal(hg[f], ho[f], err_msg=f)
            for f in ("prim", "t", "u", "v"):
                np.testing.assert_array_equal(hg[f], ref[f], err_msg=f"group {grp} {f}")
    finally:
        if old is None:
            os.environ.pop("CRB_COLLAPSE_GROUP", None)
        else:
            os.environ["CRB_COLLAPSE_GROUP"] = old
    assert infos[0][1] >= 4 and infos[0] == infos[1] == infos[2], infos


IW_HASH_SNIPPET = r"""
import sys, hashlib
import numpy as np
sys.path.insert(0, {root!r})
from crender_b200 import api, scenes
from crender_b200.api import material
rs = np.random.RandomState(4)
box = scenes._box((0, 0, 0), (0.35, 0.35, 0.35), 15.0).astype(np.float32)
ground = scenes._quad((-30, -0.5, -30), (-30, -0.5, 30), (30, -0.5, 30), (30, -0.5, -30)).astype(np.float32)
def soup(n, spread):
    inst = np.stack([scenes.compose(scenes.translation(rs.uniform(-spread, spread), rs.uniform(0, 2), rs.uniform(-spread, spread)), scenes.rotation_y(rs.uniform(0, 360))) for _ in range(n)]).astype(np.float32)
    return scenes.SceneDesc("soup", [scenes.MeshDesc(ground, np.zeros(2, np.uint32), [material()], name="ground"),
                                     scenes.MeshDesc(box, np.zeros(12, np.uint32), [material()], instances=inst, name="box")],
                            api.camera(position=(0.0, 3.0, -12.0), fov=60.0, rotation=(0.0, 12.0, 0.0)))
cases = dict(terrain=scenes.terrain_city(24, 2, n_buildings=12),     # 8 instances of two models: every ray tests every instance's bounds
             sparse=soup(60, 12.0),                                   # more than 32 instances: the candidates come from a TLAS walk
             crowded=soup(200, 1.2))                                  # rays with more than 8 candidate instances: the fallback queue
for name, desc in cases.items():
    g = api.scene(lib_path={lib!r}); scenes.load(desc, g); g.commit()
    r = api.renderer(96, 54, 5, g, seed=9); r.render(4)
    print("HASH", name, hashlib.sha256(r.raw_sum().tobytes()).hexdigest(), hashlib.sha256(r.current_depths().tobytes()).hexdigest(), r.current_stats().kernel_launches)
"""


def check_instance_wavefront_equals_loop(lib_path):
    """Two-level scenes trace their closest hits as a wavefront over instance visits (k_iw_*) or — CRB_INSTANCE_WAVEFRONT=0,
    read once per process — in the persistent two-level loop: both must give the same bits, for few instances (every ray
    tests every instance), many (TLAS walk) and crowded ones (rays with more candidates than the lists hold: fallback)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = IW_HASH_SNIPPET.format(root=root, lib=lib_path)
    res = {}
    for mode in ("0", "1"):
        env = dict(os.environ, CRB_INSTANCE_WAVEFRONT=mode)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
        res[mode] = {l.split()[1]: l.split()[2:] for l in out.stdout.splitlines() if l.startswith("HASH")}
    assert set(res["0"]) == {"terrain", "sparse", "crowded"}
    for name in res["0"]:
        assert res["0"][name][:2] == res["1"][name][:2], name
        assert int(res["1"][name][2]) > int(res["0"][name][2]), "the wavefront path did not run"
