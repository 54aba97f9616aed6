"""Host-side asset I/O around the core: OBJ -> model_data and framebuffer export.

Mirrors cr::asset_loader (src/util/asset_loader.cpp):
  load_model          :182-303  tinyobj with triangulate=true; Kd -> colour (alpha 1), every material `smooth`,
                                emission 0; map_Kd loaded flipped vertically, /255 (stbi_set_flip_vertically_on_load)
  export_framebuffer  :348-377  ./out/<name><ext>, " (n)" suffix when the file exists
  export_png / jpg    :89-110   byte = min(x*255, 255) on all four channels, JPG quality 100
  export_hdr          :172-178  pow(x, 2.2) then Radiance RGBE
  export_exr          :112-170  3 half channels in B,G,R order, ZIP blocks (own writer; tinyexr is not in the image)
Nothing here is on the measured path; it is plain numpy/PIL host code.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from .api import SMOOTH, material, model_data

PNG, JPG, EXR, HDR = "PNG", "JPG", "EXR", "HDR"
_EXT = {PNG: ".png", JPG: ".jpg", EXR: ".exr", HDR: ".hdr"}


# ---------------------------------------------------------------------------------------------- OBJ
def _parse_mtl(path: str):
    mats, cur = [], None
    if not os.path.exists(path):
        return mats
    for line in open(path, errors="replace"):
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        if tok[0] == "newmtl":
            cur = {"name": " ".join(tok[1:]), "Kd": (0.6, 0.6, 0.6), "map_Kd": ""}  # tinyobj's diffuse default is 0.6
            mats.append(cur)
        elif cur is not None and tok[0] == "Kd" and len(tok) >= 4:
            cur["Kd"] = tuple(float(x) for x in tok[1:4])
        elif cur is not None and tok[0] == "map_Kd":
            cur["map_Kd"] = tok[-1]
    return mats


def _load_texture(path: str) -> Optional[np.ndarray]:
    try:
        from PIL import Image

        im = Image.open(path).convert("RGBA")
    except Exception:
        return None
    a = np.asarray(im, dtype=np.float32) / np.float32(255.0)
    return np.ascontiguousarray(a[::-1])  # stbi_set_flip_vertically_on_load(true)


def load_model(file: str, folder: Optional[str] = None) -> model_data:
    """cr::asset_loader::load_model. Quads are split along the shorter diagonal, larger polygons fan-triangulated
    (what the reference's vendored tinyobj does for quads / convex faces; checked against the compiled reference).
    Faces without a material get an extra default material appended (tinyobj reports id -1, which the
    reference would use to index materials[] out of bounds)."""
    folder = folder if folder is not None else os.path.dirname(os.path.abspath(file))
    md = model_data(name=os.path.splitext(os.path.basename(file))[0])
    verts, uvs, mtl_defs = [], [], []
    vi, ti, mi = [], [], []
    mat_by_name, cur_mat = {}, -1
    for line in open(file, errors="replace"):
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        if tok[0] == "v":
            verts.append([float(x) for x in tok[1:4]])
        elif tok[0] == "vt":
            uvs.append([float(tok[1]), float(tok[2]) if len(tok) > 2 else 0.0])
        elif tok[0] == "mtllib":
            for m in _parse_mtl(os.path.join(folder, " ".join(tok[1:]))):
                mat_by_name[m["name"]] = len(mtl_defs)
                mtl_defs.append(m)
        elif tok[0] == "usemtl":
            cur_mat = mat_by_name.get(" ".join(tok[1:]), -1)
        elif tok[0] == "f":
            corners = []
            for c in tok[1:]:
                parts = c.split("/")
                v = int(parts[0])
                t = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                corners.append((v - 1 if v > 0 else len(verts) + v, (t - 1 if t > 0 else len(uvs) + t) if t else -1))
            tris = [(0, k, k + 1) for k in range(1, len(corners) - 1)]  # fan = tinyobj's ear clipping for convex faces
            if len(corners) == 4 and all(0 <= c[0] < len(verts) for c in corners):
                # tinyobj splits a quad along its SHORTER diagonal, in float (external/tinyobj/tinobj.h:1394-1490)
                p = [np.asarray(verts[c[0]], np.float32) for c in corners]
                a, b = p[2] - p[0], p[3] - p[1]
                sqr02 = np.float32(a[0] * a[0] + a[1] * a[1]) + np.float32(a[2] * a[2])
                sqr13 = np.float32(b[0] * b[0] + b[1] * b[1]) + np.float32(b[2] * b[2])
                tris = [(0, 1, 2), (0, 2, 3)] if sqr02 < sqr13 else [(0, 1, 3), (1, 2, 3)]
            for tri in tris:
                for c in (corners[tri[0]], corners[tri[1]], corners[tri[2]]):
                    vi.append(c[0])
                    ti.append(c[1])
                mi.append(cur_mat)
    md.vertices = np.asarray(verts, np.float32).reshape(-1, 3)
    md.texture_coords = np.asarray(uvs, np.float32).reshape(-1, 2)
    md.vertex_indices = np.asarray(vi, np.uint32)
    already = {}
    for m in mtl_defs:
        mat = material(SMOOTH, colour=(m["Kd"][0], m["Kd"][1], m["Kd"][2], 1.0), emission=0.0, name=m["name"])
        if m["map_Kd"]:
            if m["map_Kd"] in already:
                mat.tex = already[m["map_Kd"]]
            else:
                tex = _load_texture(os.path.join(folder, m["map_Kd"]))
                if tex is not None:
                    md.textures.append(tex)
                    mat.tex = len(md.textures) - 1
                    already[m["map_Kd"]] = mat.tex
        md.materials.append(mat)
    mi = np.asarray(mi, np.int64)
    if len(mi) and (mi < 0).any() or not md.materials:
        md.materials.append(material(SMOOTH, name="default"))
        mi = np.where(mi < 0, len(md.materials) - 1, mi)
    md.material_indices = mi.astype(np.uint32)
    ti = np.asarray(ti, np.int64)
    md.texture_indices = ti.astype(np.uint32) if len(ti) and (ti >= 0).all() and len(uvs) else np.zeros(0, np.uint32)
    return md


# ---------------------------------------------------------------------------------------------- export
def _to_bytes(buffer: np.ndarray) -> np.ndarray:
    # data[i] = glm::min(buffer[i] * 255.f, 255.f) stored to uint8_t (asset_loader.cpp:92,105): truncation
    x = np.minimum(np.asarray(buffer, np.float32) * np.float32(255.0), np.float32(255.0))
    return np.clip(np.nan_to_num(x, nan=0.0), 0.0, 255.0).astype(np.uint8)


def _rgbe(rgb: np.ndarray) -> np.ndarray:
    """Radiance RGBE of an (..., 3) float array, as stb_image_write's stbiw__linear_to_rgbe computes it."""
    m = rgb.max(axis=-1)
    out = np.zeros(rgb.shape[:-1] + (4,), np.uint8)
    ok = m >= 1e-32
    mant, exp = np.frexp(m[ok].astype(np.float32))
    norm = (mant * np.float32(256.0) / m[ok])[..., None]
    out[ok, :3] = (rgb[ok] * norm).astype(np.uint8)
    out[ok, 3] = (exp + 128).astype(np.uint8)
    return out


def _write_hdr(path: str, rgba: np.ndarray):
    h, w = rgba.shape[:2]
    lin = np.power(np.asarray(rgba[..., :3], np.float32), np.float32(2.2))  # asset_loader.cpp:175
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# Written by crender_b200 (flat RGBE; stb_image_write would RLE-compress the same pixels)\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(_rgbe(lin).tobytes())


def read_hdr(path: str) -> np.ndarray:
    """Minimal flat-RGBE reader (test helper for export_framebuffer(HDR))."""
    raw = open(path, "rb").read()
    head, _, rest = raw.partition(b"\n\n")
    dims, _, pix = rest.partition(b"\n")
    tok = dims.split()
    h, w = int(tok[1]), int(tok[3])
    a = np.frombuffer(pix[: h * w * 4], np.uint8).reshape(h, w, 4)
    scale = np.ldexp(np.float32(1.0), a[..., 3].astype(np.int32) - (128 + 8))
    return a[..., :3].astype(np.float32) * scale[..., None]


def _float_to_half_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> IEEE half bit patterns the way tinyexr's float_to_half_full does it (the reference stores
    HALF channels, asset_loader.cpp:158-162): mantissa truncated to 10 bits, +1 when the first dropped bit is
    set (round half up in magnitude, not ties-to-even), overflow -> inf, tiny -> signed zero/denormal."""
    f = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.int64)
    sign = (f >> 31) & 1
    exp = (f >> 23) & 0xFF
    man = f & 0x7FFFFF
    newexp = exp - 127 + 15
    out = np.zeros(f.shape, np.int64)
    infnan = exp == 255
    out = np.where(infnan, (31 << 10) | np.where(man != 0, 0x200, 0), out)
    over = (~infnan) & (exp != 0) & (newexp >= 31)
    out = np.where(over, 31 << 10, out)
    norm = (~infnan) & (exp != 0) & (newexp > 0) & (newexp < 31)
    out = np.where(norm, ((newexp << 10) | (man >> 13)) + ((man >> 12) & 1), out)
    under = (~infnan) & (exp != 0) & (newexp <= 0) & ((14 - newexp) <= 24)
    sh = np.clip(14 - newexp, 0, 62)
    mant = man | 0x800000
    out = np.where(under, (mant >> sh) + ((mant >> np.clip(sh - 1, 0, 62)) & 1), out)
    return ((sign << 15) | (out & 0x7FFF)).astype(np.uint16)


def _exr_attr(name: str, typ: str, value: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + np.int32(len(value)).tobytes() + value


def _write_exr(path: str, rgba: np.ndarray):
    """OpenEXR 2 single-part scanline file, three HALF channels named B, G, R (the reference's order,
    asset_loader.cpp:135-161), ZIP compression in 16-line blocks (tinyexr's default). The pixel values decode
    to exactly what the reference's file decodes to; the deflate stream itself comes from zlib, not miniz."""
    import zlib

    h, w = rgba.shape[:2]
    half = _float_to_half_bits(rgba[..., :3])  # (h, w, 3) as R, G, B
    chl = b""
    for ch in (b"B", b"G", b"R"):
        chl += ch + b"\0" + np.int32(1).tobytes() + b"\0\0\0\0" + np.int32(1).tobytes() + np.int32(1).tobytes()
    chl += b"\0"
    box = np.array([0, 0, w - 1, h - 1], np.int32).tobytes()
    head = b"\x76\x2f\x31\x01" + np.int32(2).tobytes()
    head += _exr_attr("channels", "chlist", chl)
    head += _exr_attr("compression", "compression", bytes([3]))  # ZIP_COMPRESSION
    head += _exr_attr("dataWindow", "box2i", box)
    head += _exr_attr("displayWindow", "box2i", box)
    head += _exr_attr("lineOrder", "lineOrder", bytes([0]))  # INCREASING_Y
    head += _exr_attr("pixelAspectRatio", "float", np.float32(1).tobytes())
    head += _exr_attr("screenWindowCenter", "v2f", np.zeros(2, np.float32).tobytes())
    head += _exr_attr("screenWindowWidth", "float", np.float32(1).tobytes())
    head += b"\0"
    chunks = []
    for y0 in range(0, h, 16):
        rows = half[y0 : y0 + 16]
        raw = np.ascontiguousarray(rows[:, :, ::-1].transpose(0, 2, 1)).astype("<u2").tobytes()  # per line: B row, G row, R row
        a = np.frombuffer(raw, np.uint8)
        re = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)  # even bytes, then odd bytes
        pred = re.copy()
        pred[1:] = (re[1:] - re[:-1] + 128 + 256) & 255
        comp = zlib.compress(pred.astype(np.uint8).tobytes())
        data = comp if len(comp) < len(raw) else raw
        chunks.append(np.int32(y0).tobytes() + np.int32(len(data)).tobytes() + data)
    off = len(head) + 8 * len(chunks)
    table = b""
    for c in chunks:
        table += np.uint64(off).tobytes()
        off += len(c)
    with open(path, "wb") as f:
        f.write(head + table + b"".join(chunks))


def read_exr(path: str) -> np.ndarray:
    """Minimal reader for the files _write_exr produces (test helper): returns (H, W, 3) float32 as R, G, B."""
    import zlib

    raw = open(path, "rb").read()
    assert raw[:4] == b"\x76\x2f\x31\x01"
    pos, attrs = 8, {}
    while raw[pos] != 0:
        e = raw.index(b"\0", pos)
        name = raw[pos:e].decode()
        e2 = raw.index(b"\0", e + 1)
        size = int(np.frombuffer(raw[e2 + 1 : e2 + 5], np.int32)[0])
        attrs[name] = (raw[e + 1 : e2].decode(), raw[e2 + 5 : e2 + 5 + size])
        pos = e2 + 5 + size
    pos += 1
    x0, y0, x1, y1 = np.frombuffer(attrs["dataWindow"][1], np.int32)
    w, h = int(x1 - x0 + 1), int(y1 - y0 + 1)
    names = [c for c in attrs["channels"][1].split(b"\0") if c in (b"B", b"G", b"R")]
    assert names == [b"B", b"G", b"R"] and attrs["compression"][1] == bytes([3])
    nblk = (h + 15) // 16
    offs = np.frombuffer(raw[pos : pos + 8 * nblk], np.uint64)
    out = np.zeros((h, w, 3), np.float32)
    for o in offs:
        o = int(o)
        y, n = (int(v) for v in np.frombuffer(raw[o : o + 8], np.int32))
        lines = min(16, h - y)
        want = lines * w * 3 * 2
        data = raw[o + 8 : o + 8 + n]
        if n < want:
            p = np.frombuffer(zlib.decompress(data), np.uint8).astype(np.int32)
            p[1:] -= 128
            re = (np.cumsum(p) & 255).astype(np.uint8)
            a = np.empty(want, np.uint8)
            a[0::2], a[1::2] = re[: (want + 1) // 2], re[(want + 1) // 2 :]
            data = a.tobytes()
        blk = np.frombuffer(data, "<u2").reshape(lines, 3, w).view(np.float16).astype(np.float32)
        out[y : y + lines] = blk[:, ::-1, :].transpose(0, 2, 1)
    return out


def export_framebuffer(buffer: np.ndarray, path: str, image_type: str = PNG, out_dir: str = "./out/") -> str:
    """cr::asset_loader::export_framebuffer: writes `out_dir/path.ext`, or `path (n).ext` if it exists.
    buffer: (H, W, 4) float32, e.g. renderer.current_progress(). Returns the file written."""
    ext = _EXT[image_type]
    os.makedirs(out_dir, exist_ok=True)
    target = os.path.join(out_dir, path + ext)
    n = 1
    while os.path.exists(target):
        target = os.path.join(out_dir, f"{path} ({n}){ext}")
        n += 1
    if image_type in (PNG, JPG):
        from PIL import Image

        img = Image.fromarray(_to_bytes(buffer), "RGBA")
        if image_type == PNG:
            img.save(target, format="PNG")
        else:
            img.convert("RGB").save(target, format="JPEG", quality=100, subsampling=0)  # stbi_write_jpg at quality 100: no chroma subsampling
    elif image_type == HDR:
        _write_hdr(target, np.asarray(buffer, np.float32))
    else:
        _write_exr(target, np.asarray(buffer, np.float32))
    return target
