import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def cuda_device_count() -> int:
    """Number of usable CUDA devices, asked of the CUDA runtime directly (no torch import)."""
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return int(n.value) if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped, not failed, on a box without a CUDA device (the product itself has no CPU
    fallback and fails loudly there: tests/test_abi.py checks that)."""
    if not any("gpu" in item.keywords for item in items):
        return
    if cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return True
