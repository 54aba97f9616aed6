"""Known-answer tests that pin the oracle to the reference SOURCE. The reference ships no
tests or golden vectors, so each expected value below is derived by hand from the cited reference lines."""
import ctypes as C
import math

import numpy as np
import pytest

from crender_b200.api import GLASS, METAL, SMOOTH, camera, material


def f3(*v):
    return np.asarray(v, np.float32)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_randf_stream_is_mt19937_default_seed(oracle):
    # renderer.cpp:6-11: thread_local std::mt19937 (seed 5489) through uniform_real_distribution<float>(0,1)
    out = np.zeros(8, np.float32)
    oracle.lib().orc_kat_mt19937_randf(8, p(out))
    raw = [3499211612, 581869302, 3890346734, 3586334585, 545404204, 4161255391, 3922919429, 949333985]
    expect = np.asarray([np.float32(r) / np.float32(4294967296.0) for r in raw], np.float32)
    np.testing.assert_allclose(out, expect, rtol=0, atol=1e-7)
    np.testing.assert_allclose(out[:4], [0.81472367, 0.135477006, 0.905791938, 0.835008562], rtol=0, atol=1e-7)


def test_counter_rng_spec(oracle):
    # DESIGN.md "Sampler": lowbias32 finaliser, two 32-bit keys k1 = mix(mix(mix(seed+phi)^pixel)^sample),
    # k2 = mix(mix(mix(seed+0x85ebca6b)^sample)^pixel), u = mix(mix(k1+dim*phi)^k2), 24-bit mantissa
    def mix(x):
        x &= 0xFFFFFFFF
        x ^= x >> 16
        x = (x * 0x7FEB352D) & 0xFFFFFFFF
        x ^= x >> 15
        x = (x * 0x846CA68B) & 0xFFFFFFFF
        x ^= x >> 16
        return x

    L = oracle.lib()
    for seed, pix, smp, dim in [(0, 0, 0, 0), (3, 12345, 7, 5), (0xFFFFFFFF, 2073599, 255, 33)]:
        k1 = mix(mix(mix((seed + 0x9E3779B9) & 0xFFFFFFFF) ^ pix) ^ smp)
        k2 = mix(mix(mix((seed + 0x85EBCA6B) & 0xFFFFFFFF) ^ smp) ^ pix)
        u = (mix(mix((k1 + dim * 0x9E3779B9) & 0xFFFFFFFF) ^ k2) >> 8) / 16777216.0
        assert L.orc_kat_rng(seed, pix, smp, dim) == np.float32(u)
    us = np.asarray([L.orc_kat_rng(1, i, 0, 0) for i in range(20000)])
    assert 0.0 <= us.min() and us.max() < 1.0 and abs(us.mean() - 0.5) < 0.01


def _cam_ray(oracle, cam, x, y, aspect):
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    cc = oracle.c_camera(cam)
    oracle.lib().orc_kat_camera_ray(C.byref(cc), x, y, aspect, p(o), p(d))
    return o, d


def test_camera_centre_and_corner(oracle):
    # camera.cpp:20-26, camera.h:21: default camera, identity matrix: centre ray = +z from (5,5,0)
    o, d = _cam_ray(oracle, camera(), 0.5, 0.5, 1.0)
    np.testing.assert_array_equal(o, f3(5, 5, 0))
    np.testing.assert_allclose(d, f3(0, 0, 1), atol=1e-7)
    a = 16.0 / 9.0
    o, d = _cam_ray(oracle, camera(), 0.0, 0.0, a)
    w = 1.0 / math.tan(math.radians(75.0) * 0.5)
    e = np.asarray([-a, -1.0, w])
    np.testing.assert_allclose(d, e / np.linalg.norm(e), atol=2e-7)


def test_camera_rotation_order(oracle):
    # camera.cpp:54-68: M = T * Ry(rot.x) * Rx(rot.y) * Rz(rot.z); centre direction = M * (0,0,w)
    cam = camera(position=(1, 2, 3), rotation=(90.0, 0.0, 0.0))
    _, d = _cam_ray(oracle, cam, 0.5, 0.5, 1.0)
    np.testing.assert_allclose(d, f3(1, 0, 0), atol=1e-6)  # +z rotated about +y by +90deg -> +x
    cam = camera(rotation=(0.0, 90.0, 0.0))
    _, d = _cam_ray(oracle, cam, 0.5, 0.5, 1.0)
    np.testing.assert_allclose(d, f3(0, -1, 0), atol=1e-6)  # about +x by +90deg: +z -> -y
    cam = camera(rotation=(90.0, 90.0, 0.0))
    _, d = _cam_ray(oracle, cam, 0.5, 0.5, 1.0)
    np.testing.assert_allclose(d, f3(0, -1, 0), atol=1e-6)  # Rx applied first, then Ry leaves -y alone


def test_camera_orthographic(oracle):
    # camera.cpp:28-37: origin = M*(scale*u, scale*v, 0, 1), direction = M[2]
    cam = camera(position=(1, 2, 3), current_mode=1, scale=2.0)
    o, d = _cam_ray(oracle, cam, 0.75, 0.25, 1.0)
    np.testing.assert_allclose(o, f3(1 + 2 * 0.5, 2 - 2 * 0.5, 3), atol=1e-6)
    np.testing.assert_allclose(d, f3(0, 0, 1), atol=1e-7)


def test_build_local_is_orthonormal(oracle):
    # sampling.h:21-33
    rs = np.random.RandomState(0)
    for _ in range(50):
        n = rs.normal(size=3)
        n = (n / np.linalg.norm(n)).astype(np.float32)
        t, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        oracle.lib().orc_kat_build_local(p(n), p(t), p(b))
        assert abs(np.dot(t, n)) < 1e-5 and abs(np.dot(b, n)) < 1e-5 and abs(np.dot(t, b)) < 1e-5
        assert abs(np.linalg.norm(t) - 1) < 1e-5 and abs(np.linalg.norm(b) - 1) < 1e-5
    n = f3(0, 0, 1)
    t, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
    oracle.lib().orc_kat_build_local(p(n), p(t), p(b))
    np.testing.assert_array_equal(t, f3(1, 0, 0))
    np.testing.assert_array_equal(b, f3(0, 1, 0))


def test_sun_transform_normal_is_column_one(oracle):
    # registry.cpp:44-48: mat3(tangent, normal, bitangent) of -sun.direction  (Y-up local frame)
    d = f3(0.8, -1.0, 0.0) / np.float32(math.sqrt(1.64))
    m = np.zeros(9, np.float32)
    oracle.lib().orc_kat_sun_transform(p(d), p(m))
    np.testing.assert_allclose(m[3:6], -d, atol=1e-7)
    M = m.reshape(3, 3)
    np.testing.assert_allclose(M @ M.T, np.eye(3), atol=1e-6)


def test_cone_sample(oracle):
    # sampling.h:35-47: phi = tau*u.x, cos = 1 - u.y(1-cos(theta_max)); (cos(phi) sin, cos, sin(phi) sin); pdf const
    out, pdf = np.zeros(3, np.float32), np.zeros(1, np.float32)
    th = math.pi / 48
    oracle.lib().orc_kat_map_to_solid_angle(0.25, 0.5, th, p(out), p(pdf))
    ct = 1 - 0.5 * (1 - math.cos(th))
    st = math.sqrt(1 - ct * ct)
    np.testing.assert_allclose(out, [math.cos(math.pi / 2) * st, ct, st], atol=2e-6)  # sqrt(1-cos^2) cancels in f32
    assert abs(pdf[0] - 1 / (2 * math.pi * (1 - math.cos(th)))) / pdf[0] < 1e-4
    oracle.lib().orc_kat_map_to_solid_angle(0.9, 0.0, th, p(out), p(pdf))
    np.testing.assert_allclose(out, [0, 1, 0], atol=1e-7)  # u.y = 0 -> cone axis (local +Y)


def test_sphere_and_diffuse_scatter(oracle):
    # sampling.h:156-172: cos = 2u.x-1, phi = tau*u.y, (sin cos(phi), cos, sin sin(phi)); scatter = normalize(n + sphere)
    out = np.zeros(3, np.float32)
    oracle.lib().orc_kat_sphere(1.0, 0.3, p(out))
    np.testing.assert_allclose(out, [0, 1, 0], atol=1e-7)
    oracle.lib().orc_kat_sphere(0.5, 0.25, p(out))
    np.testing.assert_allclose(out, [0, 0, 1], atol=2e-7)
    m = oracle.c_material(material(SMOOTH, colour=(0.5, 0.25, 1.0, 1.0)))
    n, pt, d = f3(0, 1, 0), f3(1, 2, 3), f3(0, -1, 0)
    o2, d2, alb, alpha = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_int(0)
    oracle.lib().orc_kat_process_hit(C.byref(m), p(n), p(pt), p(d), 0.5, 0.25, p(o2), p(d2), p(alb), C.byref(alpha))
    np.testing.assert_allclose(o2, pt + n * np.float32(0.0001), atol=1e-7)  # renderer.cpp:96
    e = np.asarray([0, 1, 1.0])
    np.testing.assert_allclose(d2, e / np.linalg.norm(e), atol=2e-7)
    np.testing.assert_array_equal(alb, f3(0.5, 0.25, 1.0))
    assert alpha.value == 0


def test_metal_and_glass(oracle):
    L = oracle.lib()
    o2, d2, alb, alpha = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_int(0)
    n, pt = f3(0, 1, 0), f3(0, 0, 0)
    # metal: renderer.cpp:78-91 — perfect mirror, albedo *= reflectiveness, origin + n*1e-4
    m = oracle.c_material(material(METAL, colour=(1, 0.5, 0.25, 1), reflectiveness=0.5))
    d = f3(1, -1, 0) / np.float32(math.sqrt(2))
    L.orc_kat_process_hit(C.byref(m), p(n), p(pt), p(d), 0.1, 0.9, p(o2), p(d2), p(alb), C.byref(alpha))
    np.testing.assert_allclose(d2, f3(1, 1, 0) / math.sqrt(2), atol=2e-7)
    np.testing.assert_allclose(alb, [0.5, 0.25, 0.125], atol=1e-7)
    np.testing.assert_allclose(o2, [0, 1e-4, 0], atol=1e-9)
    # glass entering (d.n < 0): ni_over_nt = 1/ior, origin = p - n*1e-4 (renderer.cpp:47-77), no Fresnel
    m = oracle.c_material(material(GLASS, ior=1.5))
    L.orc_kat_process_hit(C.byref(m), p(n), p(pt), p(d), 0.0, 0.0, p(o2), p(d2), p(alb), C.byref(alpha))
    sin_t = math.sin(math.pi / 4) / 1.5
    np.testing.assert_allclose(d2, [sin_t, -math.sqrt(1 - sin_t**2), 0], atol=3e-7)
    np.testing.assert_allclose(o2, [0, -1e-4, 0], atol=1e-9)
    # glass leaving at a grazing angle: total internal reflection -> reflect(d, n) about the ORIGINAL normal,
    # origin offset along the flipped normal (+n*1e-4 here since out_normal = -n)
    d = f3(0.9, 0.1, 0)
    d = d / np.linalg.norm(d)
    L.orc_kat_process_hit(C.byref(m), p(n), p(pt), p(d.astype(np.float32)), 0.0, 0.0, p(o2), p(d2), p(alb), C.byref(alpha))
    np.testing.assert_allclose(d2, [d[0], -d[1], 0], atol=3e-7)
    np.testing.assert_allclose(o2, [0, 1e-4, 0], atol=1e-9)
    # alpha cut-out: colour.w == 0 (renderer.cpp:37-41)
    m = oracle.c_material(material(SMOOTH, colour=(1, 1, 1, 0)))
    L.orc_kat_process_hit(C.byref(m), p(n), p(pt), p(f3(0, -1, 0)), 0.0, 0.0, p(o2), p(d2), p(alb), C.byref(alpha))
    assert alpha.value == 1


def test_resolve(oracle):
    # renderer.cpp:371-383: pow(clamp(sum/(n+1),0,1), 1/2.2)
    L = oracle.lib()
    assert L.orc_kat_resolve(2.0, 4) == pytest.approx(0.5 ** (1 / 2.2), rel=1e-6)
    assert L.orc_kat_resolve(100.0, 4) == 1.0
    assert L.orc_kat_resolve(-3.0, 4) == 0.0


def test_triangle_conventions(oracle):
    # Embree conventions (external): u weights v1, v weights v2, no back-face culling, tnear < t <= tfar
    L = oracle.lib()
    v0, v1, v2 = f3(0, 0, 0), f3(1, 0, 0), f3(0, 1, 0)
    t, u, v = C.c_float(), C.c_float(), C.c_float()

    def q(o, d, tmin=1e-5, tmax=np.inf):
        return L.orc_kat_tri(p(v0), p(v1), p(v2), p(f3(*o)), p(f3(*d)), tmin, tmax, C.byref(t), C.byref(u), C.byref(v))

    assert q((0.25, 0.5, 1), (0, 0, -1)) == 1 and (t.value, u.value, v.value) == (1.0, 0.25, 0.5)
    assert q((0.25, 0.5, -1), (0, 0, 1)) == 1 and t.value == 1.0  # back face also hits
    assert q((0.25, 0.5, 1), (0, 0, -1), tmax=1.0) == 1  # t == tfar accepted
    assert q((0.25, 0.5, 1), (0, 0, -1), tmin=1.0) == 0  # t == tnear rejected
    assert q((0.75, 0.75, 1), (0, 0, -1)) == 0  # u+v > 1
    assert q((0.25, 0.5, 1), (1, 0, 0)) == 0  # parallel
    assert q((0, 0, 1), (0, 0, -1)) == 1  # vertex / edge inclusive


def test_sky_uv_and_image_lookup(oracle):
    # renderer.cpp:279-283; image.h:104-109 (nearest, modulo, x86 signed conversion for negatives)
    L = oracle.lib()
    uv = np.zeros(2, np.float32)
    L.orc_kat_sky_uv(p(f3(1, 0, 0)), p(uv))
    np.testing.assert_allclose(uv, [0.5, 0.5], atol=1e-7)
    L.orc_kat_sky_uv(p(f3(0, 1, 0)), p(uv))
    np.testing.assert_allclose(uv, [0.5, 0.0], atol=1e-6)
    L.orc_kat_sky_uv(p(f3(0, 0, 1)), p(uv))
    np.testing.assert_allclose(uv, [0.75, 0.5], atol=1e-6)
    xy = np.zeros(2, np.uint32)
    L.orc_kat_image_get_uv_index(0.26, 1.5, 100, 10, p(xy))
    assert tuple(xy) == (26, 5)
    L.orc_kat_image_get_uv_index(-0.25, 0.0, 100, 10, p(xy))
    assert xy[0] == (2**64 - 25) % 100  # = 91: what static_cast<uint64_t>(-25.f) % 100 yields on x86-64
