import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    try:
        return int(importlib.import_module("sph-erosion_b200").capi.lib().sphe_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a CUDA device skips the gpu tests instead of failing them.  On a GPU box a
    missing or stale library is an ERROR, not a skip: the product path has no fallback."""
    if not any("gpu" in it.keywords for it in items):
        return
    if os.path.exists("/dev/nvidiactl") or os.environ.get("SPHE_REQUIRE_GPU") == "1":
        return
    if _cuda_devices() == 0:
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


def product():
    """The product package (directory name has a hyphen, so importlib)."""
    return importlib.import_module("sph-erosion_b200")


@pytest.fixture(scope="session")
def pkg():
    return product()
