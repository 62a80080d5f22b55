"""GPU parity tests of the C-ABI kernels (include/lbx.h) against the oracle.

Bars (BASELINE.json north_star): populations and density within 1e-12 relative,
velocity within 1e-12 absolute (lattice units; it legitimately holds 1e-18
noise).  With LBX_OPT_COLLIDE_LITERAL the kernels follow the
reference's operation order without FMA contraction and must be BIT-IDENTICAL to
the non-FMA CPU oracle.
"""
import os

import numpy as np
import pytest

from conftest import approx_catch2
from lambrex_b200 import lbx, workloads
from oracle import lbm_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-12
UATOL = 1e-12


@pytest.fixture(scope="module", autouse=True)
def _ctx():
    lbx.init()           # raises without a CUDA device: no CPU fallback
    yield
    lbx.set_option(lbx.OPT_COLLIDE_LITERAL, 0)


def random_state(shape, seed=0, amp=0.02):
    rng = np.random.default_rng(seed)
    rho = 1.0 + 0.05 * rng.standard_normal(shape)
    u = amp * rng.standard_normal((3,) + shape)
    f = orc.np_equilibrium(rho, u) * (1.0 + 0.01 * rng.standard_normal((15,) + shape))
    return rho, u, f


def relerr(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def periodic_fabs(nx, ny, nz):
    lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
    return lbx.Fab(lo, hi, 15), lbx.Fab(lo, hi, 15), lbx.box(lo, hi), lbx.domain(lo, hi)


@pytest.mark.parametrize("literal", [1, 0])
@pytest.mark.parametrize("shape", [(5, 7, 9), (3, 4, 130), (13, 12, 11)])
def test_equilibrium_moments_collide(coracle, literal, shape):
    lbx.set_option(lbx.OPT_COLLIDE_LITERAL, literal)
    nz, ny, nx = shape
    rho, u, f = random_state(shape, seed=nx)
    lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
    F, R, U = lbx.Fab(lo, hi, 15), lbx.Fab(lo, hi, 1), lbx.Fab(lo, hi, 3)
    bx = lbx.box(lo, hi)
    R.upload(rho[None])
    U.upload(u)
    lbx.equilibrium(F, R, U, bx)
    feq = F.download()
    want = coracle.equilibrium(rho, u)
    if literal:
        assert np.array_equal(feq, want)
    else:
        assert relerr(feq, want) < RTOL
    # moments of a non-equilibrium state
    F.upload(f)
    lbx.moments(F, R, U, bx)
    r1, u1 = coracle.moments(f)
    if literal:
        assert np.array_equal(R.download()[0], r1) and np.array_equal(U.download(), u1)
    else:
        assert relerr(R.download()[0], r1) < RTOL
        assert np.max(np.abs(U.download() - u1)) < UATOL
    # collide in place and out of place
    G = lbx.Fab(lo, hi, 15)
    lbx.collide(F, G, bx, 1.0 / 0.6, 1.3)
    want = coracle.collide(f, 1.0 / 0.6, 1.3)
    got = G.download()
    if literal:
        assert np.array_equal(got, want)
    else:
        assert relerr(got, want) < RTOL
    lbx.collide(F, F, bx, 1.0 / 0.6, 1.3)
    assert np.array_equal(F.download(), got)


@pytest.mark.parametrize("shape", [(4, 5, 6), (3, 3, 131), (2, 2, 2), (1, 1, 1)])
def test_stream_periodic_exact(coracle, shape):
    nz, ny, nx = shape
    _, _, f = random_state(shape, seed=1)
    A, B, bx, dom = periodic_fabs(nx, ny, nz)
    A.upload(f)
    lbx.stream(A, B, bx, dom)
    assert np.array_equal(B.download(), coracle.stream(f))


def test_stream_reads_ghosts_like_reference(coracle):
    """AMR-path streaming: destination = valid grown by 1, sources up to ghost ring 2,
    no wrap (src/AmrSim.cpp:114-116); ring 2 of the destination is never written."""
    nx, ny, nz, g = 6, 5, 4, 2
    lo, hi = (10, 20, 30), (10 + nx - 1, 20 + ny - 1, 30 + nz - 1)
    A, B = lbx.Fab(lo, hi, 15, ng=g), lbx.Fab(lo, hi, 15, ng=g)
    rng = np.random.default_rng(5)
    a = rng.random(A.shape)
    A.upload(a)
    dom = lbx.domain((0, 0, 0), (63, 63, 63), periodic=(0, 0, 0))
    lbx.stream(A, B, A.valid_box(grow=1), dom)
    got = B.download()
    want = np.zeros_like(a)
    for p in range(15):
        cx, cy, cz = int(orc.CX[p]), int(orc.CY[p]), int(orc.CZ[p])
        want[p, 1:-1, 1:-1, 1:-1] = a[p, 1 - cz:a.shape[1] - 1 - cz, 1 - cy:a.shape[2] - 1 - cy,
                                      1 - cx:a.shape[3] - 1 - cx]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("literal", [1, 0])
@pytest.mark.parametrize("shape", [(6, 5, 7), (3, 4, 200), (11, 12, 13)])
def test_fused_step_push_and_pull(coracle, literal, shape):
    lbx.set_option(lbx.OPT_COLLIDE_LITERAL, literal)
    nz, ny, nx = shape
    _, _, f = random_state(shape, seed=2)
    ws, wb = 1.0 / 0.6, 1.0 / 0.9
    A, B, bx, dom = periodic_fabs(nx, ny, nz)
    A.upload(f)
    lbx.collide_stream(A, B, bx, dom, ws, wb, lbx.PUSH)          # F <- S(C(F))
    want = coracle.stream(coracle.collide(f, ws, wb))
    got = B.download()
    assert np.array_equal(got, want) if literal else relerr(got, want) < RTOL
    lbx.collide_stream(A, B, bx, dom, ws, wb, lbx.PULL)          # G <- C(S(G))
    want = coracle.collide(coracle.stream(f), ws, wb)
    got = B.download()
    assert np.array_equal(got, want) if literal else relerr(got, want) < RTOL


def test_masked_collide_zeroes_fine_cells(coracle):
    nz, ny, nx = 4, 6, 8
    _, _, f = random_state((nz, ny, nx), seed=3)
    lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
    F = lbx.Fab(lo, hi, 15)
    Mk = lbx.Fab(lo, hi, 1, ng=2, dtype=lbx.I32)
    mask = np.zeros(Mk.shape, dtype=np.int32)
    mask[0, 2:4, 2:5, 2:6] = 1            # low-index corner in valid coordinates
    Mk.upload(mask)
    F.upload(f)
    lbx.collide(F, F, lbx.box(lo, hi), 1.1, 0.9, mask=Mk, fine_val=1)
    got = F.download()
    want = coracle.collide(f, 1.1, 0.9)
    want[:, 0:2, 0:3, 0:4] = 0.0
    assert relerr(got[want != 0], want[want != 0]) < RTOL
    assert np.all(got[want == 0] == 0.0)


@pytest.mark.parametrize("scheme", [lbx.PUSH, lbx.PULL])
def test_pulse_regression_on_gpu(coracle, golden_dir, scheme):
    """/root/reference/tests/catch2RegressionTests.cpp:6-93 on the GPU path: golden
    vectors under Catch2's Approx AND populations vs the oracle at 1e-12."""
    g = np.load(os.path.join(golden_dir, "pulse_regression.npz"))
    nx, ny, nz, tau = 10, 10, 50, 0.5
    w = workloads.omega(tau)
    rho0 = orc.user_to_fab(workloads.pulse_density(nx, ny, nz), nx, ny, nz)
    u0 = np.zeros((3, nz, ny, nx))
    lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
    A, B, bx, dom = periodic_fabs(nx, ny, nz)
    R, U = lbx.Fab(lo, hi, 1), lbx.Fab(lo, hi, 3)
    R.upload(rho0[None])
    U.upload(u0)
    lbx.equilibrium(A, R, U, bx)
    f_orc = coracle.equilibrium(rho0, u0)

    def check(t):
        lbx.moments(A, R, U, bx)
        r, v = R.download()[0], U.download()
        assert approx_catch2(r.reshape(-1), g["RHO_t%d" % t]).all()
        vflat = np.ascontiguousarray(v.transpose(1, 2, 3, 0)).reshape(-1)
        gv = g["VEL_t%d" % t]
        # golden velocities hold 1e-18 round-off noise that only the literal operation
        # order reproduces digit for digit; elsewhere the absolute bar applies
        assert (approx_catch2(vflat, gv) | (np.abs(vflat - gv) < 1e-15)).all()
        ro, vo = coracle.moments(f_orc)
        assert relerr(A.download(), f_orc) < RTOL
        assert relerr(r, ro) < RTOL and np.max(np.abs(v - vo)) < UATOL

    check(0)
    for t in (100, 200):
        if scheme == lbx.PUSH:
            for _ in range(100):
                lbx.collide_stream(A, B, bx, dom, w, w, lbx.PUSH)
                A, B = B, A
        else:   # S (C S)^(n-1) C : collide, n-1 fused pull steps, stream
            lbx.collide(A, A, bx, w, w)
            for _ in range(99):
                lbx.collide_stream(A, B, bx, dom, w, w, lbx.PULL)
                A, B = B, A
            lbx.stream(A, B, bx, dom)
            A, B = B, A
        f_orc = coracle.step(f_orc, w, w, 100)
        check(t)


def test_shear_wave_64_vs_oracle(coracle):
    n, tau, steps = 64, 0.1, 20
    w = workloads.omega(tau)
    rho_c, u_c = workloads.shear_wave(n, n, n)
    rho0 = orc.user_to_fab(rho_c, n, n, n)
    u0 = orc.user_to_fab(u_c, n, n, n, 3)
    f = coracle.equilibrium(rho0, u0)
    A, B, bx, dom = periodic_fabs(n, n, n)
    A.upload(f)
    for _ in range(steps):
        lbx.collide_stream(A, B, bx, dom, w, w, lbx.PUSH)
        A, B = B, A
    want = coracle.step(f, w, w, steps)
    assert relerr(A.download(), want) < RTOL


def test_full_size_256_properties(coracle):
    """BASELINE.json configs[1] at full size through size-independent properties:
    the planar pulse stays uniform in x,y, so every column must equal the oracle's
    4x4x256 run; total mass and momentum are conserved."""
    n, tau, steps = 256, 0.5, 12
    w = workloads.omega(tau)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    A, B, bx, dom = periodic_fabs(n, n, n)
    R, U = lbx.Fab(lo, hi, 1), lbx.Fab(lo, hi, 3)
    col = orc.user_to_fab(workloads.pulse_density(4, 4, n), 4, 4, n)       # [k, 4, 4]
    rho0 = np.ascontiguousarray(np.broadcast_to(col[:, :1, :1], (n, n, n)))
    R.upload(rho0[None])
    lbx.equilibrium(A, R, U, bx)
    for _ in range(steps):
        lbx.collide_stream(A, B, bx, dom, w, w, lbx.PUSH)
        A, B = B, A
    lbx.moments(A, R, U, bx)
    r, v = R.download()[0], U.download()
    f_small = coracle.step(coracle.equilibrium(col, np.zeros((3, n, 4, 4))), w, w, steps)
    rs, vs = coracle.moments(f_small)
    assert relerr(r, np.broadcast_to(rs[:, :1, :1], r.shape)) < RTOL
    assert np.max(np.abs(v - np.broadcast_to(vs[:, :, :1, :1], v.shape))) < UATOL
    assert abs(r.sum() / rho0.sum() - 1.0) < 1e-13
    assert np.max(np.abs((r * v).sum(axis=(1, 2, 3)))) / r.sum() < 1e-15


def test_errors_are_loud():
    A = lbx.Fab((0, 0, 0), (3, 3, 3), 15)
    B = lbx.Fab((0, 0, 0), (3, 3, 3), 15)
    dom = lbx.domain((0, 0, 0), (7, 7, 7))
    with pytest.raises(lbx.LbxError):     # periodic wrap needs the fab to span the domain
        lbx.collide_stream(A, B, lbx.box((0, 0, 0), (3, 3, 3)), dom, 1.0, 1.0, lbx.PUSH)
    with pytest.raises(lbx.LbxError):     # aliasing
        lbx.stream(A, A, lbx.box((0, 0, 0), (3, 3, 3)), lbx.domain((0, 0, 0), (3, 3, 3)))
    with pytest.raises(lbx.LbxError):     # box outside fab
        lbx.collide(A, B, lbx.box((0, 0, 0), (4, 3, 3)), 1.0, 1.0)


# ------------------------------------------------------------------ walls (addition)
@pytest.mark.parametrize("walls", [(0, 0, 1), (1, 1, 1), (1, 0, 1)])
def test_fused_step_with_bounce_back_walls_matches_oracle(walls):
    """lbx_collide_stream with lbx_domain.periodic[d] = 2: half-way bounce-back walls at both faces of direction d --
    against the numpy restatement (oracle np_step_walls), populations <= 1e-12 relative, mass conserved."""
    from oracle import lbm_oracle as orc
    rng = np.random.default_rng(11)
    nx, ny, nz, steps = 20, 9, 7, 12
    rho = 1.0 + 0.01 * rng.standard_normal((nz, ny, nx))
    u = 0.02 * rng.standard_normal((3, nz, ny, nx))
    f0 = orc.np_equilibrium(rho, u)
    w = 1.0 / 0.6
    lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
    bx = lbx.box(lo, hi)
    dom = lbx.domain(lo, hi, tuple(2 if wl else 1 for wl in walls))
    A, B = lbx.Fab(lo, hi, 15), lbx.Fab(lo, hi, 15)
    A.upload(f0)
    for _ in range(steps):
        lbx.collide_stream(A, B, bx, dom, w, w, lbx.PUSH)
        A, B = B, A
    got = A.download()
    want = orc.np_step_walls(f0, w, w, walls, steps)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-12
    assert abs(got.sum() - f0.sum()) < 1e-12 * f0.sum()
    with pytest.raises(lbx.LbxError):
        lbx.collide_stream(A, B, bx, dom, w, w, lbx.PULL)       # walls exist in the push scheme only
