"""pytest configuration: `gpu` marker, library/oracle fixtures.

`-m "not gpu"`: oracle vs its independent restatement and invariants, host logic,
C-ABI symbol check (no compute calls).  `-m gpu`: the parity tests proper, all
through the C ABI on a real B200.
"""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """the product package (hyphenated directory name)"""
    build = importlib.import_module("hycom-src_b200.build")
    build.build_library()
    return importlib.import_module("hycom-src_b200")


@pytest.fixture(scope="session")
def oracle():
    """the CPU oracle (test infrastructure)"""
    odir = os.path.join(ROOT, "oracle")
    lib = os.path.join(odir, "_build", "liboracle.so")
    src = [os.path.join(odir, f) for f in ("tsadvc_oracle.c", "cnuity_oracle.inc.c", "tsadvc_oracle.h", "Makefile")]
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in src):
        subprocess.run(["make", "-C", odir], check=True, capture_output=True)
    import oracle_binding
    return oracle_binding.Oracle(lib)
