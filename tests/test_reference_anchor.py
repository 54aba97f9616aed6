"""The anchor to the reference's own code.

oracle/ref compiles the REFERENCE'S sources for the hot path — src/render/{renderer,scene,camera,ray}.cpp,
src/render/entities/registry.cpp, src/objects/{model,thread_pool}.cpp, src/util/asset_loader.cpp ... — unmodified, against
shim headers for the three third-party libraries that are absent from the image (glm and Embree: arithmetic restated with
their published conventions; fmt: log strings). Run with a one-thread pool, the reference's default-seeded mt19937
(renderer.cpp:6-11) is consumed in call order, so its output is reproducible. Three links:

  1. reference (compiled here)  ==  committed fixtures tests/golden/reference_v1.npz       (where /root/reference exists)
  2. oracle, reference-stream mode  ==  fixtures, BIT FOR BIT: raw sums, display, AOVs, ray count   (everywhere, CPU)
  3. CUDA path (or the kernel-logic harness) replaying the same numbers through crb_render_set_sample_table  vs  fixtures:
     bit-identical on the CPU harness for un-instanced scenes, within the north-star tolerance on the GPU.

What stays restated rather than compiled: glm's vector arithmetic and Embree's BVH / triangle test (oracle/ref/shim).
"""
import os
import sys

import numpy as np
import pytest

import common
import ref_binding as rb
from crender_b200 import api, scenes

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_golden as mg  # noqa: E402

GOLD = os.path.join(HERE, "golden", "reference_v1.npz")
HAVE_REFERENCE = os.path.isdir(os.path.join(rb.REFERENCE_SRC, "src"))
IMG_RELRMSE = 0.01


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def case(gold, name):
    w, h, bounces, spp = (int(v) for v in gold[f"{name}/params"])
    return w, h, bounces, spp, int(gold[f"{name}/draws_before"][0])


def oracle_replay(oracle, gold, name, record=True):
    w, h, bounces, spp, draws = case(gold, name)
    o = oracle.scene()
    scenes.load(mg.scene_of(name), o)
    o.commit()
    r = oracle.renderer(w, h, bounces, o, seed=0)
    table = np.zeros((spp, w * h, 2 + 4 * bounces), np.float32)
    if record:
        r.set_sample_table(table)
    r.set_reference_stream(True, discard=draws)
    r.render(spp, nthreads=1)
    return r, table


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference is only present in the development container")
@pytest.mark.parametrize("name", list(mg.CASES))
def test_compiled_reference_reproduces_the_fixtures(name, gold):
    w, h, bounces, spp, _ = case(gold, name)
    ref = rb.render(mg.scene_of(name), w, h, bounces, spp)
    if ref["draws_before"] != int(gold[f"{name}/draws_before"][0]):
        pytest.skip("the reference rendered a different number of empty passes before the scene arrived (a race in its own start-up): other stream offset")
    for k in ("raw", "progress", "albedo", "normal", "depth"):
        np.testing.assert_array_equal(ref[k], gold[f"{name}/{k}"], err_msg=k)
    assert ref["total_rays"] == int(gold[f"{name}/total_rays"][0])


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_equals_reference_bit_for_bit(oracle, gold, name):
    r, _ = oracle_replay(oracle, gold, name)
    np.testing.assert_array_equal(r.raw_sum()[..., :3], gold[f"{name}/raw"])
    np.testing.assert_array_equal(r.current_progress(), gold[f"{name}/progress"])
    np.testing.assert_array_equal(r.current_albedos(), gold[f"{name}/albedo"])
    np.testing.assert_array_equal(r.current_normals(), gold[f"{name}/normal"])
    np.testing.assert_array_equal(r.current_depths(), gold[f"{name}/depth"])
    assert int(r.current_stats().ref_rays) == int(gold[f"{name}/total_rays"][0])  # renderer.cpp:271-272,356


def test_oracle_table_replay_is_the_same_render(oracle, gold):
    # the recorded table fed back as the sampler reproduces the reference-stream render: the table carries every number
    # that feeds the estimator (the draws the reference makes and discards do not)
    name = "textured"
    w, h, bounces, spp, _ = case(gold, name)
    _, table = oracle_replay(oracle, gold, name)
    o = oracle.scene()
    scenes.load(mg.scene_of(name), o)
    o.commit()
    r = oracle.renderer(w, h, bounces, o, seed=0)
    r.set_sample_table(table, replay=True)
    r.render(spp, nthreads=1)
    np.testing.assert_array_equal(r.raw_sum()[..., :3], gold[f"{name}/raw"])


def product_vs_reference(oracle, gold, lib_path, name, exact):
    w, h, bounces, spp, _ = case(gold, name)
    _, table = oracle_replay(oracle, gold, name)
    g = api.scene(lib_path=lib_path)
    scenes.load(mg.scene_of(name), g)
    g.commit()
    r = api.renderer(w, h, bounces, g, seed=0)
    r.set_sample_table(table)
    r.render(spp)
    raw, ref = r.raw_sum(), gold[f"{name}/raw"]
    assert np.all(raw[..., 3] == spp)
    if exact:
        np.testing.assert_array_equal(raw[..., :3], ref)
        np.testing.assert_array_equal(r.current_progress(), gold[f"{name}/progress"])
    assert common.relrmse(raw[..., :3], ref) <= IMG_RELRMSE
    assert common.relrmse(r.current_progress()[..., :3], gold[f"{name}/progress"][..., :3]) <= IMG_RELRMSE
    # first-hit AOVs of the last sample = the reference's (primary-ray agreement)
    for got, k in ((r.current_albedos(), "albedo"), (r.current_normals(), "normal"), (r.current_depths(), "depth")):
        want = gold[f"{name}/{k}"]
        same = np.all(np.abs(got - want) <= 1e-5 * np.maximum(1.0, np.abs(want)), axis=-1)
        if exact:
            np.testing.assert_array_equal(got, want, err_msg=k)
        assert same.mean() >= 0.9999, (k, same.mean())
    assert abs(int(r.current_stats().ref_rays) - int(gold[f"{name}/total_rays"][0])) <= (0 if exact else max(2, int(1e-3 * gold[f"{name}/total_rays"][0])))


@pytest.mark.parametrize("name", list(mg.CASES))
def test_kernel_logic_equals_reference(oracle, gold, emu_lib, name):
    # the product's kernel bodies (CPU harness, same libm as the reference build) replaying the reference's numbers:
    # bit-identical images, AOVs and ray counts — instanced models included (two-level traversal with the reference's own
    # per-instance arithmetic, incl. non-rigid transforms)
    product_vs_reference(oracle, gold, emu_lib, name, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_cuda_path_vs_reference(oracle, gold, product_lib, name):
    # the CUDA path against images THE REFERENCE produced: same numbers, libdevice instead of libm transcendentals
    product_vs_reference(oracle, gold, product_lib, name, exact=False)


# ---------------------------------------------------------------------------------------------- asset I/O
# cr::asset_loader compiled from the reference (asset_loader.cpp + its vendored stb_image / stb_image_write / tinyexr /
# tinyobj) is both the reference writer and the independent DECODER of what the two hosts write.
def _run_ref(args, cwd):
    import subprocess

    r = subprocess.run([rb.build(), *args], cwd=cwd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr + r.stdout


def _decode_with_reference(path, cwd):
    out = os.path.join(cwd, "dec.bin")
    _run_ref(["load_picture", path, out], cwd)
    raw = open(out, "rb").read()
    w, h = np.frombuffer(raw[:8], np.int32)
    return np.frombuffer(raw[8:], np.float32).reshape(h, w, 4)


def _test_image():
    rng = np.random.default_rng(3)
    h, w = 37, 50
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([xx / w, yy / h, 0.5 + 0.5 * np.sin(xx / 5.0), np.ones_like(xx, float)], -1).astype(np.float32)
    img[5:15, 5:20, :3] = rng.random((10, 15, 3)).astype(np.float32)
    img[20:25, 30:40, :3] = 1.7  # over-range: PNG/JPG clamp at 255, HDR/EXR keep it
    img[30:, :8, :3] = 1e-4
    return img


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference is only present in the development container")
@pytest.mark.parametrize("kind", ["PNG", "JPG", "EXR", "HDR"])
def test_exporters_decode_like_the_reference_file(tmp_path, kind):
    import subprocess

    from crender_b200 import assets

    img = _test_image()
    h, w = img.shape[:2]
    cwd = str(tmp_path)
    os.makedirs(os.path.join(cwd, "out"))
    img.tofile(os.path.join(cwd, "img.f32"))
    _run_ref(["export", kind, str(w), str(h), "img.f32", "ref"], cwd)  # ./out/ref.<ext> (asset_loader.cpp:348-377)
    ext = {"PNG": ".png", "JPG": ".jpg", "EXR": ".exr", "HDR": ".hdr"}[kind]
    ref_file = os.path.join(cwd, "out", "ref" + ext)
    assert os.path.exists(ref_file)
    want = _decode_with_reference(ref_file, cwd)
    # Python host
    py_file = assets.export_framebuffer(img, "py", kind, out_dir=os.path.join(cwd, "py"))
    # C++ host (crender_b200/host/assets.hpp through its test front end)
    tool = os.path.join(rb.ROOT, "crender_b200", "host", "assets_tool")
    open(os.path.join(cwd, "im.bin"), "wb").write(np.asarray([w, h], np.int32).tobytes() + img.tobytes())
    r = subprocess.run([tool, "export", os.path.join(cwd, "im.bin"), "cpp", kind, os.path.join(cwd, "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp_file = r.stdout.strip()
    for host, f in (("python", py_file), ("c++", cpp_file)):
        got = _decode_with_reference(f, cwd)
        assert got.shape == want.shape, host
        if kind == "JPG":
            # lossy: three encoders (stb in the reference, libjpeg, the C++ host's own), all at quality 100 / 4:4:4
            # (each encoder is within ~3 grey levels of the source on the noise patch; two of them are compared here)
            diff = np.abs(got[..., :3] - want[..., :3]) * 255
            assert diff.max() <= 6.5 and diff.mean() <= 1.0, (host, diff.max(), diff.mean())
        else:
            np.testing.assert_array_equal(got[..., :3], want[..., :3], err_msg=f"{host} {kind}")


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference is only present in the development container")
def test_obj_loaders_equal_the_reference(tmp_path):
    import struct
    import subprocess

    from PIL import Image

    from crender_b200 import assets

    d = str(tmp_path)
    open(os.path.join(d, "m.obj"), "w").write("mtllib m.mtl\nv -1 -1 0\nv 1 -1 0\nv 1 1 0.25\nv -1 1 0\nv -1 -1 -1\nv 1 -1 -1\nv 1 -1 1\nv -1 -1 1\nv 0 2 0\n"
                                              "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 0.5\nusemtl tex\nf 1/1 2/2 3/3 4/4\nusemtl red\nf 5/1 6/2 7/3 8/4\nf 4/4 3/3 9/5\n")
    open(os.path.join(d, "m.mtl"), "w").write("newmtl tex\nKd 1 1 1\nmap_Kd t.png\nnewmtl red\nKd 0.8 0.2 0.1\n")
    rng = np.random.default_rng(2)
    tex = rng.integers(0, 256, (6, 5, 4), dtype=np.uint8)
    Image.fromarray(tex, "RGBA").save(os.path.join(d, "t.png"))
    # the reference builds the texture path as folder + '\\' + name (asset_loader.cpp:237): on this file system that is a
    # file whose NAME starts with a backslash
    Image.fromarray(tex, "RGBA").save(os.path.join(d, "\\t.png"), format="PNG")
    _run_ref(["load_model", os.path.join(d, "m.obj"), d + "/", os.path.join(d, "ref.bin")], d)
    raw = open(os.path.join(d, "ref.bin"), "rb").read()
    pos = 0

    def take(fmt, n=1):
        nonlocal pos
        a = np.frombuffer(raw, fmt, n, pos)
        pos += a.nbytes
        return a

    nv = int(take("<u8")[0]); verts = take("<f4", nv * 3).reshape(-1, 3)
    nt = int(take("<u8")[0]); uvs = take("<f4", nt * 2).reshape(-1, 2)
    nn = int(take("<u8")[0]); take("<f4", nn * 3)
    idx = {}
    for k in ("v", "m", "t", "n"):
        n = int(take("<u8")[0])
        idx[k] = take("<u4", n)
    nm = int(take("<u8")[0])
    mats = []
    for _ in range(nm):
        st = int(take("<u4")[0]); ior, rough, refl, emis = take("<f4", 4); col = take("<f4", 4); tex_id = int(take("<i4")[0])
        ln = int(take("<u8")[0]); name = raw[pos : pos + ln].decode(); pos += ln
        mats.append((st, float(emis), tuple(col), tex_id, name))
    ntex = int(take("<u8")[0])
    texs = []
    for _ in range(ntex):
        tw, th = (int(v) for v in take("<u8", 2))
        texs.append(take("<f4", tw * th * 4).reshape(th, tw, 4))
    assert ntex == 1 and nm == 2

    def check(md_vertices, md_uvs, vi, mi, ti, materials, textures, who):
        np.testing.assert_array_equal(np.asarray(md_vertices, np.float32).reshape(-1, 3), verts, err_msg=who)
        np.testing.assert_array_equal(np.asarray(md_uvs, np.float32).reshape(-1, 2), uvs, err_msg=who)
        np.testing.assert_array_equal(vi, idx["v"], err_msg=who)
        np.testing.assert_array_equal(mi, idx["m"], err_msg=who)
        np.testing.assert_array_equal(ti, idx["t"], err_msg=who)
        assert len(materials) == nm
        for (st, emis, col, tex_id, name), got in zip(mats, materials):
            assert got[0] == st and got[1] == emis and tuple(np.float32(c) for c in got[2]) == tuple(np.float32(c) for c in col) and got[3] == tex_id and got[4] == name, (who, got)
        np.testing.assert_array_equal(np.asarray(textures[0], np.float32), texs[0], err_msg=who)

    md = assets.load_model(os.path.join(d, "m.obj"))
    check(md.vertices, md.texture_coords, md.vertex_indices, md.material_indices, md.texture_indices,
          [(m.shade_type, m.emission, m.colour, -1 if m.tex is None else m.tex, m.name) for m in md.materials], md.textures, "python host")
    tool = os.path.join(rb.ROOT, "crender_b200", "host", "assets_tool")
    r = subprocess.run([tool, "load", os.path.join(d, "m.obj"), os.path.join(d, "cpp.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c = open(os.path.join(d, "cpp.bin"), "rb").read()
    cnt = np.frombuffer(c, "<u8", 7)
    p = 56
    cv = np.frombuffer(c, "<f4", int(cnt[0]) * 3, p); p += cv.nbytes
    cu = np.frombuffer(c, "<f4", int(cnt[1]) * 2, p); p += cu.nbytes
    cvi = np.frombuffer(c, "<u4", int(cnt[2]), p); p += cvi.nbytes
    cti = np.frombuffer(c, "<u4", int(cnt[3]), p); p += cti.nbytes
    cmi = np.frombuffer(c, "<u4", int(cnt[4]), p); p += cmi.nbytes
    cm = []
    for _ in range(int(cnt[5])):
        rec = np.frombuffer(c, "<f4", 6, p); p += 24
        st, ln = struct.unpack_from("<II", c, p); p += 8
        name = c[p : p + ln].decode(); p += ln
        cm.append((st, float(rec[4]), tuple(rec[:4]), int(rec[5]), name))
    ct = []
    for _ in range(int(cnt[6])):
        tw, th = struct.unpack_from("<QQ", c, p); p += 16
        ct.append(np.frombuffer(c, "<f4", tw * th * 4, p).reshape(th, tw, 4)); p += tw * th * 16
    check(cv, cu, cvi, cmi, cti, cm, ct, "c++ host")
