import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a CUDA device: skip the gpu-marked tests instead of erroring in
    their fixtures.  Only a machine that really has no device skips -- a missing or broken liblbx.so still
    fails loudly -- and LBX_REQUIRE_GPU=1 (set it on the B200 box) turns a missing device into failures."""
    if os.environ.get("LBX_REQUIRE_GPU", "") not in ("", "0"):
        return
    try:
        import ctypes
        from lambrex_b200 import lbx
        n = ctypes.c_int(0)
        lbx.lib().lbx_device_count(ctypes.byref(n))
        ndev = n.value
    except Exception:
        return                      # library missing / not loadable: let the tests fail loudly
    if ndev > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device here (gpu tests run on the B200 box; LBX_REQUIRE_GPU=1 forbids this skip)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def approx_catch2(x, golden):
    """Catch2 v2 Approx with default epsilon = 100*FLT_EPSILON, scale 0, margin 0
    (the rule /root/reference/tests/catch2RegressionTests.cpp applies):
    |x - g| <= eps * |g|  (so g == 0 demands x == 0 exactly)."""
    import numpy as np
    eps = 100.0 * float(np.finfo(np.float32).eps)
    return np.abs(np.asarray(x) - golden) <= eps * np.abs(golden)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def coracle():
    from oracle.lbm_oracle import COracle, build
    build()
    return COracle()
