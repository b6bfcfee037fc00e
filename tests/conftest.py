"""pytest configuration: registers the `gpu` marker and puts the repo root + the product package on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    n = None
    for item in items:
        if "gpu" in item.keywords:
            if n is None:
                n = _cuda_device_count()
            if n == 0:
                item.add_marker(pytest.mark.skip(reason="no CUDA device here"))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def build_c_client(tmp_path):
    """Compiles tests/c/abi_client.c (plain C99, -pedantic -Werror) against include/b200blur.h and links the built library."""
    import subprocess
    import b200blur
    exe = os.path.join(str(tmp_path), "abi_client")
    libdir = os.path.dirname(b200blur.lib_path())
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_client.c"), "-L", libdir, "-lb200blur", f"-Wl,-rpath,{libdir}", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe
