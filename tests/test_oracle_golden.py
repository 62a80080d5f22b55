"""Pin the oracle (numpy + C) to the reference's own golden vectors and tables.

Mirrors /root/reference/tests/catch2RegressionTests.cpp:6-93 ("pulse
Regression"): 10x10x50 periodic pulse, tau = 0.5, compare rho,u at t = 0, 100,
200 with Catch2's Approx rule.
"""
import os

import numpy as np
import pytest

from conftest import approx_catch2
from lambrex_b200 import workloads
from oracle import lbm_oracle as orc

NX, NY, NZ, TAU = 10, 10, 50, 0.5


def _pulse_initial():
    rho = orc.user_to_fab(workloads.pulse_density(NX, NY, NZ), NX, NY, NZ)
    u = np.zeros((3, NZ, NY, NX))
    return rho, u


def _flatten(rho, u):
    # golden order: k outer, j, i inner; velocity component innermost
    return rho.reshape(-1), np.ascontiguousarray(u.transpose(1, 2, 3, 0)).reshape(-1)


def _check(rho, u, g, t):
    r, v = _flatten(rho, u)
    okr = approx_catch2(r, g["RHO_t%d" % t])
    okv = approx_catch2(v, g["VEL_t%d" % t])
    assert okr.all(), "rho t=%d: %d mismatches" % (t, (~okr).sum())
    assert okv.all(), "vel t=%d: %d mismatches" % (t, (~okv).sum())


def test_tables_match_reference_bit_for_bit(golden_dir, coracle):
    g = np.load(os.path.join(golden_dir, "mode_matrices.npz"))
    Mc, Mic, c = coracle.tables()
    for Mx, Mix in ((orc.M, orc.MINV), (Mc, Mic)):
        assert np.array_equal(Mx, g["MODE_MATRIX"])
        assert np.array_equal(Mix, g["MODE_MATRIX_INVERSE"])
    assert np.array_equal(orc.DELTA, g["DELTA"])
    assert np.array_equal(c[:, 0], g["CX"]) and np.array_equal(c[:, 1], g["CY"])
    assert np.array_equal(c[:, 2], g["CZ"])
    assert np.array_equal(orc.W, g["W"])


def test_table_identities():
    # SURVEY.md 8c known-answer identities
    assert np.abs(orc.M @ orc.MINV - np.eye(15)).max() < 4e-16
    assert np.array_equal(orc.M[1], orc.CX) and np.array_equal(orc.M[2], orc.CY)
    assert np.array_equal(orc.M[3], orc.CZ)
    assert np.array_equal(orc.MINV[:, 0], orc.W)
    assert np.array_equal(orc.M[14], orc.CX * orc.CY * orc.CZ)


def test_numpy_oracle_reproduces_pulse_regression(golden_dir):
    g = np.load(os.path.join(golden_dir, "pulse_regression.npz"))
    rho, u = _pulse_initial()
    f = orc.np_equilibrium(rho, u)
    _check(*orc.np_moments(f), g, 0)
    w = workloads.omega(TAU)
    f = orc.np_step(f, w, w, 100)
    _check(*orc.np_moments(f), g, 100)
    f = orc.np_step(f, w, w, 100)
    _check(*orc.np_moments(f), g, 200)


def test_c_oracle_reproduces_pulse_regression_and_equals_numpy(golden_dir, coracle):
    g = np.load(os.path.join(golden_dir, "pulse_regression.npz"))
    rho, u = _pulse_initial()
    f = coracle.equilibrium(rho, u)
    assert np.array_equal(f, orc.np_equilibrium(rho, u))
    _check(*coracle.moments(f), g, 0)
    w = workloads.omega(TAU)
    f100 = coracle.step(f, w, w, 100)
    _check(*coracle.moments(f100), g, 100)
    assert np.array_equal(f100, orc.np_step(f, w, w, 100))   # bit for bit
    f200 = coracle.step(f100, w, w, 100)
    _check(*coracle.moments(f200), g, 200)


def test_c_and_numpy_agree_on_random_state(coracle):
    rng = np.random.default_rng(7)
    rho = 1.0 + 0.1 * rng.standard_normal((6, 5, 7))
    u = 0.05 * rng.standard_normal((3, 6, 5, 7))
    f = orc.np_equilibrium(rho, u) * (1.0 + 0.01 * rng.standard_normal((15, 6, 5, 7)))
    assert np.array_equal(coracle.collide(f, 1.3, 0.9), orc.np_collide(f, 1.3, 0.9))
    assert np.array_equal(coracle.stream(f), orc.np_stream(f))
    r1, u1 = coracle.moments(f)
    r2, u2 = orc.np_moments(f)
    assert np.array_equal(r1, r2) and np.array_equal(u1, u2)


def test_collide_conserves_mass_and_momentum(coracle):
    rng = np.random.default_rng(3)
    f = 0.05 + 0.01 * rng.random((15, 4, 4, 4))
    g = coracle.collide(f, 1.0 / 0.6, 1.2)
    r0, u0 = coracle.moments(f)
    r1, u1 = coracle.moments(g)
    assert np.allclose(r0, r1, rtol=1e-14, atol=0)
    assert np.allclose(r0 * u0, r1 * u1, rtol=0, atol=1e-15)


def test_equilibrium_roundtrips_moments(coracle):
    # tests/catch2InitTests.cpp:5-48 in spirit: rho = 0.63, u = 0.23 survive f_eq -> moments
    rho = np.full((13, 12, 11), 0.63)
    u = np.full((3, 13, 12, 11), 0.23)
    r, v = coracle.moments(coracle.equilibrium(rho, u))
    assert approx_catch2(r, rho).all() and approx_catch2(v, u).all()


def test_reference_pass_structure_equals_periodic_step(coracle):
    """The ghosted-box CPU baseline (FillPatch copy, in-place collide, FillBoundary,
    stream into a fresh fab, swap) is decomposition independent on one level."""
    rng = np.random.default_rng(11)
    nx, ny, nz = 10, 10, 50
    rho = 1.0 + 0.05 * rng.standard_normal((nz, ny, nx))
    u = 0.02 * rng.standard_normal((3, nz, ny, nx))
    f = coracle.equilibrium(rho, u)
    want = coracle.step(f, 1.1, 0.8, 3)
    edges = [[0, 10], [0, 4, 10], [0, 24, 50]]
    for order in (0, 1):
        got, secs = coracle.ref_passes(f, 1.1, 0.8, 3, edges, loop_order=order)
        assert np.array_equal(got, want)
        assert secs >= 0.0


def test_linear_moments_restate_the_mode_matrix_rows(coracle):
    """oracle.linear_moments (generic derived variables, include/derived_var.h:55-91): rows of the
    mode matrix reproduce the oracle's own density / velocity (src/AmrSim.cpp:957-971)."""
    rng = np.random.default_rng(3)
    f = rng.random((15, 4, 5, 6)) + 0.5
    M, _, c = coracle.tables()
    rho, u = coracle.moments(f)
    assert np.max(np.abs(orc.linear_moments(f, M[:1])[0] - rho)) < 1e-14
    assert np.max(np.abs(orc.linear_moments(f, M[1:4], per_unit_density=True) - u)) < 1e-14
    assert np.array_equal(orc.linear_moments(f, np.asarray(c, dtype=np.float64).T), orc.linear_moments(f, M[1:4]))


def test_wall_oracle_reduces_to_periodic_and_conserves_mass():
    """np_stream_walls (addition: half-way bounce-back walls): without walls it IS the periodic stream; with walls the
    total mass is conserved exactly per population pair and the momentum of a closed box decays."""
    import numpy as np
    from oracle import lbm_oracle as orc
    rng = np.random.default_rng(5)
    nz, ny, nx = 6, 5, 7
    rho = 1.0 + 0.01 * rng.standard_normal((nz, ny, nx))
    u = 0.02 * rng.standard_normal((3, nz, ny, nx))
    f = orc.np_equilibrium(rho, u)
    assert np.array_equal(orc.np_stream_walls(f, (0, 0, 0)), orc.np_stream(f))
    for walls in ((0, 0, 1), (1, 1, 1), (1, 0, 0)):
        g = orc.np_step_walls(f, 1.0 / 0.6, 1.0 / 0.6, walls, 40)
        assert abs(g.sum() - f.sum()) < 1e-12 * f.sum()
    p0 = np.abs(orc.np_moments(f)[1] * rho).sum()
    g = orc.np_step_walls(f, 1.0 / 0.6, 1.0 / 0.6, (1, 1, 1), 200)
    r1, u1 = orc.np_moments(g)
    assert np.abs(u1 * r1).sum() < 0.5 * p0          # no-slip walls drain the momentum of a closed box
    assert (orc.OPP[orc.OPP] == np.arange(15)).all()
    assert (orc.CX[orc.OPP] == -orc.CX).all() and (orc.CY[orc.OPP] == -orc.CY).all() and (orc.CZ[orc.OPP] == -orc.CZ).all()
