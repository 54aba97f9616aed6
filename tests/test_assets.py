"""Host-side asset I/O: OBJ -> model_data following asset_loader.cpp:182-303 and the
framebuffer exporters following asset_loader.cpp:89-178,348-377."""
import os

import numpy as np
from PIL import Image

from crender_b200 import assets

OBJ = """# a unit quad and a triangle
mtllib test.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 2 0 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
usemtl red
f 1/1 2/2 3/3 4/4
usemtl tex
f 2/2 5/1 3/3
"""
MTL = """newmtl red
Kd 0.8 0.1 0.2
newmtl tex
Kd 1 1 1
map_Kd tex.png
"""


def test_load_model(tmp_path):
    (tmp_path / "test.obj").write_text(OBJ)
    (tmp_path / "test.mtl").write_text(MTL)
    tex = np.zeros((2, 3, 4), np.uint8)
    tex[0, :, 0] = 255  # top row red
    tex[1, :, 1] = 255  # bottom row green
    tex[..., 3] = 255
    Image.fromarray(tex, "RGBA").save(tmp_path / "tex.png")
    md = assets.load_model(str(tmp_path / "test.obj"))
    assert md.name == "test"
    assert md.vertices.shape == (5, 3) and md.texture_coords.shape == (4, 2)
    # the quad is split like the reference's vendored tinyobj does (shorter diagonal; a square's equal diagonals give
    # (0,1,3) (1,2,3) — verified against the compiled reference in test_reference_anchor.py), then the triangle
    np.testing.assert_array_equal(md.vertex_indices, [0, 1, 3, 1, 2, 3, 1, 4, 2])
    np.testing.assert_array_equal(md.texture_indices, [0, 1, 3, 1, 2, 3, 1, 0, 2])
    np.testing.assert_array_equal(md.material_indices, [0, 0, 1])
    assert [m.name for m in md.materials] == ["red", "tex"]
    assert tuple(md.materials[0].colour) == (0.8, 0.1, 0.2, 1.0) and md.materials[0].shade_type == 1 and md.materials[0].emission == 0.0
    assert md.materials[1].tex == 0 and len(md.textures) == 1
    # flipped vertically on load: row 0 of the cr::image is the file's bottom row (green)
    np.testing.assert_allclose(md.textures[0][0, 0], [0, 1, 0, 1])
    np.testing.assert_allclose(md.textures[0][1, 0], [1, 0, 0, 1])


def test_faces_without_material_get_a_default(tmp_path):
    (tmp_path / "m.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    md = assets.load_model(str(tmp_path / "m.obj"))
    assert len(md.materials) == 1 and md.material_indices.tolist() == [0] and len(md.texture_indices) == 0


def test_export_png_jpg_hdr(tmp_path):
    rs = np.random.RandomState(0)
    buf = rs.uniform(0, 1.2, (5, 7, 4)).astype(np.float32)
    buf[..., 3] = 1.0
    out = str(tmp_path / "out")
    p = assets.export_framebuffer(buf, "frame", assets.PNG, out_dir=out)
    assert p.endswith("frame.png")
    got = np.asarray(Image.open(p))
    expect = np.minimum(buf * np.float32(255.0), np.float32(255.0)).astype(np.uint8)  # asset_loader.cpp:92
    np.testing.assert_array_equal(got, expect)
    # existing file -> " (1)" suffix (asset_loader.cpp:363-368)
    p2 = assets.export_framebuffer(buf, "frame", assets.PNG, out_dir=out)
    assert os.path.basename(p2) == "frame (1).png"
    pj = assets.export_framebuffer(buf, "frame", assets.JPG, out_dir=out)
    assert np.asarray(Image.open(pj)).shape == (5, 7, 3)
    ph = assets.export_framebuffer(buf, "frame", assets.HDR, out_dir=out)
    back = assets.read_hdr(ph)
    expect = np.power(buf[..., :3], 2.2)
    # RGBE shares one exponent per pixel: absolute error up to max_component / 256 (truncating mantissa)
    assert np.all(np.abs(back - expect) <= expect.max(axis=-1, keepdims=True) / 128 + 1e-6)


def test_export_exr(tmp_path):
    # asset_loader.cpp:112-170: three HALF channels named B, G, R; tinyexr default = ZIP blocks of 16 lines
    rng = np.random.default_rng(5)
    h, w = 37, 23  # not a multiple of the 16-line block
    img = rng.random((h, w, 4), dtype=np.float32) * 4.0
    img[0, 0, :3] = (0.0, 65520.0, 1e-9)  # zero, overflow -> inf, underflow -> 0
    img[0, 1, :3] = (np.float32(1.0) + np.float32(2.0**-11), 6e-6, -2.5)  # tie rounds up in magnitude; denormal half; sign
    path = assets.export_framebuffer(img, "frame", assets.EXR, out_dir=str(tmp_path))
    assert path.endswith("frame.exr")
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x76\x2f\x31\x01\x02\x00\x00\x00"
    assert raw.index(b"channels\0chlist\0") == 8
    chl = raw[raw.index(b"chlist\0") + 11 :]
    assert chl[:2] == b"B\0" and chl[18:20] == b"G\0" and chl[36:38] == b"R\0"  # (A)BGR order, pixel type 1 = HALF
    assert chl[2:6] == b"\x01\0\0\0"
    back = assets.read_exr(path)
    assert back.shape == (h, w, 3)
    with np.errstate(over="ignore"):
        want = img[..., :3].astype(np.float16).astype(np.float32)
    # identical to IEEE round-to-nearest except exact ties, which tinyexr rounds away from zero
    ties = (img[..., :3].view(np.uint32) & 0x1FFF) == 0x1000
    np.testing.assert_array_equal(back[~ties], want[~ties])
    assert back[0, 0, 0] == 0.0 and np.isinf(back[0, 0, 1]) and back[0, 0, 2] == 0.0
    assert back[0, 1, 0] == np.float32(1.0) + np.float32(2.0**-10)
    assert back[0, 1, 2] == -2.5 and abs(back[0, 1, 1] - 6e-6) < 6e-8
    # second export of the same name gets the " (1)" suffix like every other format
    assert assets.export_framebuffer(img, "frame", assets.EXR, out_dir=str(tmp_path)).endswith("frame (1).exr")


# ---------------------------------------------------------------------------------------------- C++ host (assets.hpp)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "crender_b200", "host", "assets_tool")


def _tool(*args):
    import subprocess

    import pytest

    if not os.path.exists(TOOL):
        pytest.fail(f"{TOOL} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    r = subprocess.run([TOOL, *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip()


def _read_dump(path):
    raw = open(path, "rb").read()
    c = np.frombuffer(raw[:56], np.uint64).astype(int)
    pos = 56

    def take(n, dt):
        nonlocal pos
        a = np.frombuffer(raw[pos : pos + n * np.dtype(dt).itemsize], dt)
        pos += a.nbytes
        return a

    d = {"vertices": take(c[0] * 3, np.float32).reshape(-1, 3), "texture_coords": take(c[1] * 2, np.float32).reshape(-1, 2),
         "vertex_indices": take(c[2], np.uint32), "texture_indices": take(c[3], np.uint32), "material_indices": take(c[4], np.uint32),
         "materials": [], "textures": []}
    for _ in range(c[5]):
        rec = take(6, np.float32)
        typ, ln = take(2, np.uint32)
        name = raw[pos : pos + ln].decode()
        pos += int(ln)
        d["materials"].append((tuple(rec[:4]), float(rec[4]), int(rec[5]), int(typ), name))
    for _ in range(c[6]):
        w, h = take(2, np.uint64).astype(int)
        d["textures"].append(take(w * h * 4, np.float32).reshape(h, w, 4))
    assert pos == len(raw)
    return d


def _same_model(tmp_path, obj_name):
    md = assets.load_model(str(tmp_path / obj_name))
    name = _tool("load", str(tmp_path / obj_name), str(tmp_path / "dump.bin"))
    d = _read_dump(tmp_path / "dump.bin")
    assert name == md.name
    for k in ("vertices", "texture_coords", "vertex_indices", "texture_indices", "material_indices"):
        np.testing.assert_array_equal(d[k], getattr(md, k), err_msg=k)
    assert len(d["materials"]) == len(md.materials) and len(d["textures"]) == len(md.textures)
    for (col, emission, tex, typ, nm), m in zip(d["materials"], md.materials):
        assert col == tuple(np.float32(x) for x in m.colour) and emission == m.emission and typ == m.shade_type and nm == m.name
        assert tex == (-1 if m.tex is None else m.tex)
    for a, b in zip(d["textures"], md.textures):
        np.testing.assert_array_equal(a, b)
    return md


def test_cpp_load_model_matches_python(tmp_path):
    (tmp_path / "test.obj").write_text(OBJ + "f -5/-4 -4/-3 -3/-2\n")  # + a face with negative (relative) indices
    (tmp_path / "test.mtl").write_text(MTL + "newmtl pal\nKd 0.5 0.5 0.5\nmap_Kd -s 1 1 1 pal.png\nnewmtl again\nmap_Kd tex.png\nnewmtl gone\nmap_Kd missing.png\n")
    rng = np.random.default_rng(4)
    tex = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)  # compressed by PIL with dynamic Huffman blocks and filters
    Image.fromarray(tex, "RGBA").save(tmp_path / "tex.png")
    Image.fromarray(rng.integers(0, 256, (16, 16), dtype=np.uint8), "L").convert("P").save(tmp_path / "pal.png")  # palette PNG
    md = _same_model(tmp_path, "test.obj")
    assert len(md.textures) == 2 and md.materials[3].tex == md.materials[1].tex and md.materials[4].tex is None
    (tmp_path / "m.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    _same_model(tmp_path, "m.obj")
    # gradient images exercise the PNG filters; RGB and grey+alpha colour types
    g = np.add.outer(np.arange(40), np.arange(64)).astype(np.uint8)
    Image.fromarray(np.stack([g, g.T[:40, :64] if False else g[::-1], g // 2], -1), "RGB").save(tmp_path / "tex.png")
    Image.fromarray(np.stack([g, 255 - g], -1), "LA").save(tmp_path / "pal.png")
    _same_model(tmp_path, "test.obj")


def test_cpp_export_matches_python(tmp_path):
    rng = np.random.default_rng(7)
    img = rng.random((19, 23, 4), dtype=np.float32) * 1.3 - 0.1  # some values outside [0, 1]
    img[0, 0] = [0.0, 1.0, 0.5, 1.0]
    img[0, 1] = [1e-4, 65504.0, 1e6, 1.0]  # half denormal / max / overflow
    with open(tmp_path / "img.bin", "wb") as f:
        f.write(np.array([23, 19], np.int32).tobytes() + img.tobytes())
    out = str(tmp_path / "out")
    p_png = _tool("export", str(tmp_path / "img.bin"), "shot", "PNG", out)
    assert p_png.endswith("shot.png")
    np.testing.assert_array_equal(np.asarray(Image.open(p_png)), assets._to_bytes(img))
    assert _tool("export", str(tmp_path / "img.bin"), "shot", "PNG", out).endswith("shot (1).png")  # collision rule
    pos = np.clip(img, 0.0, None)
    with open(tmp_path / "pos.bin", "wb") as f:
        f.write(np.array([23, 19], np.int32).tobytes() + pos.tobytes())
    p_hdr = _tool("export", str(tmp_path / "pos.bin"), "shot", "HDR", out)
    ref_hdr = assets.export_framebuffer(pos, "ref", assets.HDR, out_dir=out)
    a, b = assets.read_hdr(p_hdr), assets.read_hdr(ref_hdr)
    # libm powf vs numpy power may differ in the last ulp, which can move an 8-bit mantissa by one step
    assert np.abs(a - b).max() <= np.maximum(a, b).max(axis=-1, keepdims=True).max() / 128.0 and (a != b).mean() < 0.01
    p_exr = _tool("export", str(tmp_path / "img.bin"), "shot", "EXR", out)
    ref_exr = assets.export_framebuffer(img, "ref", assets.EXR, out_dir=out)
    np.testing.assert_array_equal(assets.read_exr(p_exr), assets.read_exr(ref_exr))
