"""The drop-in boundary: libcrender_b200.so loads, exports every symbol include/crender_b200.h declares,
and — with no GPU — fails loudly instead of computing anything on the CPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from crender_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "crender_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_table_agree():
    assert declared_functions() == sorted(_capi.SIGNATURES)


def test_integration_guide_names_every_entry_point():
    """INTEGRATION.md maps each reference interface to the entry point that replaces it: none may be left out."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_functions() if n not in doc]
    assert not missing, missing


def test_library_exports_every_declared_symbol(product_lib):
    lib = _capi.load(product_lib)
    for name in declared_functions():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", product_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(crb_[a-z0-9_]+)", out))
    assert set(declared_functions()) <= exported


def test_library_is_cuda_sm100a_and_not_linked_to_the_oracle(product_lib):
    out = subprocess.run(["ldd", product_lib], capture_output=True, text=True).stdout
    assert "oracle" not in out
    sass = subprocess.run(["cuobjdump", "-lelf", product_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass, sass


def test_no_gpu_fails_loudly(product_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _capi.load(product_lib)
    h = C.c_void_p()
    rc = lib.crb_scene_create(C.byref(h))
    assert rc == 10  # CRB_ERR_NO_DEVICE
    assert b"no CPU path" in lib.crb_last_error()
    from crender_b200 import api

    with pytest.raises(api.CrbError):
        api.scene(lib_path=product_lib)


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "crender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_binding" not in text and "liboracle" not in text and "oracle/" not in text.replace("oracle/_", ""), f


def test_header_is_plain_c(tmp_path):
    """The boundary must be bindable from C (cgo/JNI/ctypes-style FFI): the header compiles as C99 and a C
    program that references every entry point links against the library."""
    names = declared_functions()
    src = tmp_path / "use.c"
    body = "\n".join(f"    p[{i}] = (fn) {n};" for i, n in enumerate(names))
    src.write_text(f'#include "crender_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {{\n    fn p[{len(names)}];\n{body}\n'
                   f'    printf("%d\\n", p[0] != 0);\n    crb_scene *s = 0;\n    int rc = crb_scene_create(&s);\n    printf("%d %s\\n", rc, crb_last_error());\n    return 0;\n}}\n')
    exe = tmp_path / "use"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(exe),
                    f"-L{os.path.join(ROOT, 'crender_b200')}", "-lcrender_b200", f"-Wl,-rpath,{os.path.join(ROOT, 'crender_b200')}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0
    import torch

    if not torch.cuda.is_available():
        assert out.stdout.splitlines()[-1].startswith("10 ")  # CRB_ERR_NO_DEVICE, loudly
