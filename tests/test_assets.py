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
