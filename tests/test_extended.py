"""Extended shading mode (CRB_RENDER_FLAG_EXTENDED): GGX metal, Fresnel dielectric, Lambert, NEE of the sun and of
emissive triangles. The reference only has dead code for it (src/render/brdf.h:10-29, src/util/sampling.h:83-142;
SURVEY.md D5), so the mode is specified by the oracle. These CPU tests pin that specification with checks that do
not depend on the oracle's own code path:

  * the GGX sample weight against an independent numpy quadrature of D * Vis * cos over the hemisphere, with D and
    Vis written from the reference's formulas (sampling.h:94-118);
  * the dielectric against Snell's law and the closed-form normal-incidence reflectance;
  * the area-light NEE estimator against the estimator that finds emitters by path hits only (same expectation);

and then compare the product's kernel bodies (tests/emu) with the oracle, exactly.
"""
import ctypes as C

import numpy as np
import pytest

import common
import parity_cases as pc
from crender_b200 import scenes
from crender_b200.api import GLASS, METAL, SMOOTH, material


def _scatter(oracle, m, n, d, u0, u1):
    L = oracle.lib()
    om = oracle.c_material(m)
    n3 = np.asarray(n, np.float32)
    p3 = np.zeros(3, np.float32)
    d3 = np.asarray(d, np.float32)
    w, o, dr = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    kind = L.orc_kat_scatter_extended(C.byref(om), ptr(n3), ptr(p3), ptr(d3), C.c_float(u0), C.c_float(u1), ptr(w), ptr(o), ptr(dr))
    return kind, w, o, dr


def _ggx_albedo_quadrature(a, cos_v, n=512):
    """integral over the hemisphere of D(h) * Vis(NoV, NoL) * NoL dw with F = 1 (sampling.h:94-118, brdf.h:17-28)."""
    v = np.array([np.sqrt(1 - cos_v * cos_v), cos_v, 0.0])
    mu = (np.arange(n) + 0.5) / n  # cos(theta_l), uniform in cos -> dw = dmu dphi
    phi = (np.arange(2 * n) + 0.5) / (2 * n) * 2 * np.pi
    MU, PH = np.meshgrid(mu, phi, indexing="ij")
    s = np.sqrt(1 - MU * MU)
    l = np.stack([s * np.cos(PH), MU, s * np.sin(PH)], -1)
    h = l + v
    h /= np.linalg.norm(h, axis=-1, keepdims=True)
    noh, nol, nov = h[..., 1], MU, cos_v
    a2 = a * a
    d = (noh * a2 - noh) * noh + 1.0
    D = a2 / (d * d * np.pi)
    vis = 0.5 / (nol * np.sqrt(nov * nov * (1 - a2) + a2) + nov * np.sqrt(nol * nol * (1 - a2) + a2))
    return float((D * vis * nol).sum() * (1.0 / n) * (2 * np.pi / (2 * n)))


@pytest.mark.parametrize("rough,cos_v", [(0.5, 1.0), (0.5, 0.6), (0.25, 0.8), (0.9, 0.7)])
def test_ggx_sample_weight_integrates_to_the_directional_albedo(oracle, rough, cos_v):
    m = material(METAL, colour=(1, 1, 1, 1), reflectiveness=1.0, roughness=rough)
    d = (-np.sqrt(1 - cos_v * cos_v), -cos_v, 0.0)  # travelling towards the surface, normal = +y
    k = 96
    us = (np.arange(k) + 0.5) / k
    tot = 0.0
    for u0 in us:
        for u1 in us:
            kind, w, _, dr = _scatter(oracle, m, (0, 1, 0), d, float(u0), float(u1))
            assert kind in (1, 2)
            if kind == 1:
                assert dr[1] > 0  # reflected into the upper hemisphere
                assert abs(np.linalg.norm(dr) - 1) < 1e-5
                tot += float(w[0])
    est = tot / (k * k)
    ref = _ggx_albedo_quadrature(rough, cos_v)
    assert abs(est - ref) <= 0.01 * ref + 2e-3, (est, ref)


def test_ggx_fresnel_tint(oracle):
    # f0 = colour * reflectiveness; at normal incidence and low roughness the weight is ~ f0
    m = material(METAL, colour=(0.9, 0.5, 0.2, 1), reflectiveness=0.8, roughness=0.02)
    kind, w, o, dr = _scatter(oracle, m, (0, 1, 0), (0, -1, 0), 0.3, 0.7)
    assert kind == 1
    np.testing.assert_allclose(w, np.array([0.72, 0.4, 0.16]), rtol=2e-3)
    np.testing.assert_allclose(dr, [0, 1, 0], atol=0.05)
    assert o[1] > 0  # offset along the shading normal


def test_dielectric_fresnel_and_snell(oracle):
    ior = 1.5
    m = material(GLASS, colour=(0.9, 0.95, 1.0, 1), ior=ior)
    r0 = ((1 - ior) / (1 + ior)) ** 2  # 0.04
    kind, w, o, dr = _scatter(oracle, m, (0, 1, 0), (0, -1, 0), r0 - 1e-3, 0.0)
    assert kind == 1 and dr[1] > 0 and o[1] > 0  # reflected, stays outside
    kind, w, o, dr = _scatter(oracle, m, (0, 1, 0), (0, -1, 0), r0 + 1e-3, 0.0)
    assert kind == 1 and dr[1] < 0 and o[1] < 0  # refracted, starts inside
    np.testing.assert_allclose(w, [0.9, 0.95, 1.0])
    # oblique incidence: Snell sin(t) = sin(i)/ior, and the exact unpolarised Fresnel threshold
    ti = np.radians(50.0)
    d = (np.sin(ti), -np.cos(ti), 0.0)
    tt = np.arcsin(np.sin(ti) / ior)
    rs = (np.cos(ti) - ior * np.cos(tt)) / (np.cos(ti) + ior * np.cos(tt))
    rp = (ior * np.cos(ti) - np.cos(tt)) / (ior * np.cos(ti) + np.cos(tt))
    R = 0.5 * (rs * rs + rp * rp)
    kind, _, _, dr = _scatter(oracle, m, (0, 1, 0), d, R + 2e-3, 0.0)
    np.testing.assert_allclose(dr / np.linalg.norm(dr), [np.sin(tt), -np.cos(tt), 0.0], atol=1e-5)
    kind, _, _, dr = _scatter(oracle, m, (0, 1, 0), d, R - 2e-3, 0.0)
    np.testing.assert_allclose(dr, [np.sin(ti), np.cos(ti), 0.0], atol=1e-5)
    # leaving the glass beyond the critical angle: total internal reflection for every u0
    ti = np.radians(60.0)
    kind, _, o, dr = _scatter(oracle, m, (0, 1, 0), (np.sin(ti), np.cos(ti), 0.0), 0.999, 0.0)
    assert dr[1] < 0 and o[1] < 0  # travelling +y inside, reflected back down, stays inside


def test_lambert_is_face_forwarded(oracle):
    m = material(SMOOTH, colour=(0.5, 0.6, 0.7, 1))
    for u0, u1 in [(0.1, 0.2), (0.9, 0.5), (0.5, 0.99)]:
        kind, w, o, dr = _scatter(oracle, m, (0, 1, 0), (0.3, 1.0, 0.1), u0, u1)  # hits the back face
        assert kind == 0 and dr[1] < 0 and o[1] < 0
        np.testing.assert_allclose(w, [0.5, 0.6, 0.7])


def test_area_light_nee_matches_hit_only_estimator(oracle):
    """Same expectation from two estimators: NEE on the light list vs finding the emitter by path hits only."""
    desc = scenes.cornell()
    o = oracle.scene()
    scenes.load(desc, o)
    o.commit()
    imgs = {}
    for mode, spp in ((1, 96), (2, 512)):
        r = oracle.renderer(32, 32, 8, o, seed=5, extended=mode)
        r.render(spp)
        s = r.raw_sum()
        imgs[mode] = s[..., :3] / s[..., 3:4]
    a, b = imgs[1].mean(), imgs[2].mean()
    assert abs(a / b - 1) < 0.02, (a, b)
    A = imgs[1].reshape(4, 8, 4, 8, 3).mean(axis=(1, 3))
    B = imgs[2].reshape(4, 8, 4, 8, 3).mean(axis=(1, 3))
    assert common.relrmse(A, B) < 0.05


def _ext_scenes():
    s = common.small_scenes()
    s["lights"] = scenes.lights_scene(40, 20, n_lights=8)
    return s


@pytest.mark.parametrize("name,w,h,spp,bounces", [("cornell", 40, 40, 3, 8), ("mesh", 48, 32, 2, 5), ("textured", 48, 36, 3, 6), ("terrain", 40, 24, 2, 4),
                                                  ("lights", 48, 32, 3, 6)])
def test_extended_images_match_oracle(oracle, emu_lib, name, w, h, spp, bounces):
    desc = _ext_scenes()[name]
    identity_only = all(m.instances is None for m in desc.meshes)
    pc.check_image(oracle, emu_lib, desc, w, h, spp, bounces, exact=identity_only, extended=True)


def test_extended_differs_from_ref_exact_and_sort_is_equivalent(emu_lib):
    pc.check_extended_properties(emu_lib)
