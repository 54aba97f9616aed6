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
    # the quad is fan-triangulated: (1,2,3) (1,3,4), then the triangle
    np.testing.assert_array_equal(md.vertex_indices, [0, 1, 2, 0, 2, 3, 1, 4, 2])
    np.testing.assert_array_equal(md.texture_indices, [0, 1, 2, 0, 2, 3, 1, 0, 2])
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
