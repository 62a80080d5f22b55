"""CPU-side checks of the C-ABI boundary: liblbx.so loads and exports every symbol that
include/lbx.h declares (no compute calls: there is no GPU here), and the ctypes table in
lambrex_b200/lbx.py names exactly that set."""
import os
import re

from lambrex_b200 import lbx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(lbx_[a-z0-9_]+)\s*\(", txt))


def test_library_exports_every_declared_symbol():
    L = lbx.lib()
    names = declared_symbols("lbx.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "liblbx.so does not export " + n
    assert names == set(lbx.SYMBOLS), names ^ set(lbx.SYMBOLS)


def test_no_gpu_means_loud_failure_not_fallback():
    L = lbx.lib()
    import ctypes
    n = ctypes.c_int(0)
    L.lbx_device_count(ctypes.byref(n))
    if n.value == 0:
        assert L.lbx_init(-1) != 0
        assert b"no CUDA device" in L.lbx_last_error()
        assert L.lbx_sync() != 0          # not initialised


def test_tables_query_needs_no_gpu():
    M, Mi, c, w = lbx.tables()
    import numpy as np
    assert np.allclose(M @ Mi, np.eye(15), atol=1e-15)
    assert abs(w.sum() - 1.0) < 1e-15 and (c[0] == 0).all()


def test_host_library_exports_every_symbol_of_lambrex_c_h():
    """liblambrex.so (the C mirror of AmrSim, include/lambrex_c.h): every declared entry point is
    exported and bound by lambrex_b200/amrsim.py."""
    from lambrex_b200 import amrsim
    L = amrsim.lib()
    names = declared_symbols("lambrex_c.h")
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), "liblambrex.so does not export " + n
    assert names == set(amrsim.SYMBOLS), names ^ set(amrsim.SYMBOLS)


def test_product_tables_equal_the_reference_matrices_bit_for_bit():
    """The kernels' moment basis (csrc/d3q15.cuh generators, queried through lbx_d3q15_tables) against the
    matrices parsed from /root/reference/src/AmrSim.cpp:1037-1073 and include/d3q15_bgk.h:14-28
    (tests/golden/mode_matrices.npz): identical doubles, not just close."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "mode_matrices.npz"))
    M, Mi, c, w = lbx.tables()
    assert np.array_equal(M, g["MODE_MATRIX"]) and np.array_equal(Mi, g["MODE_MATRIX_INVERSE"])
    assert np.array_equal(c[:, 0], g["CX"]) and np.array_equal(c[:, 1], g["CY"]) and np.array_equal(c[:, 2], g["CZ"])
    assert np.array_equal(w, g["W"])
