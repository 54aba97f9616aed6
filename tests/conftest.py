"""Test configuration.

Markers:  gpu  — needs a CUDA device and the built libcrender_b200.so (run with `-m gpu` on a B200).
Everything else runs on CPU: the oracle, its known-answer tests, the C-ABI load/export checks, the
host-side partition logic (gloo, world_size 2) and the kernel-logic harness (tests/emu, the product's
kernel bodies compiled for serial CPU execution — a test tool, never a fallback of the product).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PRODUCT_LIB = os.path.join(ROOT, "crender_b200", "libcrender_b200.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libcrb_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200) and libcrender_b200.so")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding

    oracle_binding.build()
    oracle_binding.lib()
    return oracle_binding


@pytest.fixture(scope="session")
def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    return EMU_LIB


@pytest.fixture(scope="session")
def product_lib():
    if not os.path.exists(PRODUCT_LIB):
        pytest.fail(f"{PRODUCT_LIB} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return PRODUCT_LIB
