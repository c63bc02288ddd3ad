import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_built():
    """Build the product library and the test oracle when a fresh checkout has neither (both are git-ignored
    artefacts). nvcc cross-compiles sm_100a without a GPU; on the GPU box the prebuilt files travel with the snapshot."""
    import subprocess
    lib = os.path.join(ROOT, "apex_solver_b200", "csrc", "libapex_gpu.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.dirname(lib), "-j", "8"], stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


_ensure_built()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `-m gpu` on a B200 box)")


def _have_gpu():
    try:
        from apex_solver_b200 import _ffi
        return _ffi.load_library().apex_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
