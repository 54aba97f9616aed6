"""Driver of oracle/_ref/ref_render — the REFERENCE'S OWN sources (renderer.cpp, scene.cpp, model.cpp, registry.cpp,
camera.cpp, asset_loader.cpp ...) compiled unmodified against shim headers for glm / fmt / embree3 (oracle/ref). TEST
INFRASTRUCTURE ONLY: it exists in the development container, where /root/reference is; the fixtures it produces are
committed under tests/golden/ (make_reference_golden.py) so that the checks also run where the reference is absent."""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "ref")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_render")
REFERENCE_SRC = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "src")) or os.path.exists(REF_BIN)


def build() -> str:
    if os.path.isdir(os.path.join(REFERENCE_SRC, "src")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        subprocess.run(["make", "-s", "-C", REF_DIR], check=True)
    if not os.path.exists(REF_BIN):
        raise FileNotFoundError(REF_BIN)
    return REF_BIN


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32).tobytes()


def scene_bytes(desc, w, h, bounces, spp, threads=1) -> bytes:
    """The binary scene description ref_harness.cpp reads (every value 4 bytes, little endian)."""
    out = [struct.pack("<5I", w, h, bounces, spp, threads)]
    c = desc.cam
    out += [_f(c.position), _f(c.rotation), struct.pack("<ffI", c.fov, c.scale, c.current_mode)]
    s = desc.sun
    out += [struct.pack("<Iff", int(desc.sun_enabled), s.size, s.intensity), _f(s.direction), _f(s.colour)]
    if desc.skybox is not None:
        sk = np.ascontiguousarray(desc.skybox, np.float32)
        out += [struct.pack("<II", sk.shape[1], sk.shape[0]), _f(desc.skybox_rotation), sk.tobytes()]
    else:
        out += [struct.pack("<II", 0, 0), _f((0.0, 0.0))]
    out.append(struct.pack("<I", len(desc.textures)))
    for t in desc.textures:
        t = np.ascontiguousarray(t, np.float32)
        out += [struct.pack("<II", t.shape[1], t.shape[0]), t.tobytes()]
    out.append(struct.pack("<I", len(desc.meshes)))
    for m in desc.meshes:
        v = np.ascontiguousarray(m.verts, np.float32).reshape(-1, 9)
        out += [struct.pack("<I", v.shape[0]), v.tobytes()]
        if m.uvs is not None:
            out += [struct.pack("<I", 1), _f(m.uvs)]
        else:
            out.append(struct.pack("<I", 0))
        out.append(np.ascontiguousarray(m.mat_idx, np.uint32).tobytes())
        out.append(struct.pack("<I", len(m.materials)))
        for mm in m.materials:
            out += [struct.pack("<Iffff", mm.shade_type, mm.ior, mm.roughness, mm.reflectiveness, mm.emission), _f(mm.colour),
                    struct.pack("<i", -1 if mm.tex is None else int(mm.tex))]
        inst = np.eye(4, dtype=np.float32)[None] if m.instances is None else np.ascontiguousarray(m.instances, np.float32).reshape(-1, 4, 4)
        out += [struct.pack("<I", inst.shape[0]), inst.tobytes()]
    return b"".join(out)


def render(desc, w, h, bounces, spp, binary=None, timeout=5, retries=40, threads=1) -> dict:
    """Runs the compiled reference on `desc` with a one-thread pool. Returns raw (h,w,3) sums, the four RGBA images,
    and draws_before (randf() draws of the empty-scene passes the renderer completed before the scene was handed over)."""
    binary = binary or build()
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "scene.bin"), os.path.join(d, "out.bin")
        open(src, "wb").write(scene_bytes(desc, w, h, bounces, spp, threads))
        for attempt in range(retries):
            # (the reference's pause()/thread-pool hand-shakes wait on condition variables without predicates,
            # renderer.cpp:172-183, thread_pool.cpp:12-13,55-56: a lost wake-up hangs it — time out and retry. Measured: 0 - 50 %
            # of the attempts hang depending on the scene while a good run of these small cases takes 0.1 - 0.3 s, hence a
            # short time-out and many retries)
            try:
                r = subprocess.run([binary, "render", src, dst], capture_output=True, text=True, timeout=timeout)
            except subprocess.TimeoutExpired:
                continue
            if r.returncode == 0 and os.path.exists(dst):
                break
        else:
            raise RuntimeError("ref_render did not finish")
        raw = open(dst, "rb").read()
    draws_before, total_rays, samples, calls = struct.unpack_from("<4Q", raw, 0)
    (seconds,) = struct.unpack_from("<d", raw, 32)
    off, n = 40, w * h
    out = {"draws_before": draws_before, "total_rays": total_rays, "samples": samples, "intersect_calls": calls, "seconds": seconds}
    out["raw"] = np.frombuffer(raw, np.float32, n * 3, off).reshape(h, w, 3).copy()
    off += n * 12
    for k in ("progress", "albedo", "normal", "depth"):
        out[k] = np.frombuffer(raw, np.float32, n * 4, off).reshape(h, w, 4).copy()
        off += n * 16
    return out


def oracle_like_reference(oracle, desc, w, h, bounces, spp, draws_before):
    """The oracle on the same scene with the reference's random stream: one thread, default-seeded mt19937, the draws of
    the empty-scene passes skipped."""
    from crender_b200 import scenes

    o = oracle.scene()
    scenes.load(desc, o)
    o.commit()
    r = oracle.renderer(w, h, bounces, o, seed=0)
    r.set_reference_stream(True, discard=draws_before)
    r.render(spp, nthreads=1)
    return r
