import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
