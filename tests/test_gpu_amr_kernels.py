"""GPU parity tests of the batched AMR-path kernels and gather plans (lbx_mf_*, lbx_plan_*)
against numpy restatements (oracle/amr_oracle.py).  Integer/zero patterns must be exact;
collisions follow the usual 1e-12 bar (bit-exact with LBX_OPT_COLLIDE_LITERAL)."""
import numpy as np
import pytest

from lambrex_b200 import lbx
from oracle import amr_oracle as ao
from oracle import lbm_oracle as orc

pytestmark = pytest.mark.gpu

BOXES = [((0, 0, 0), (5, 6, 3)), ((6, 0, 0), (12, 6, 3)), ((0, 7, 0), (12, 9, 3)), ((-4, -3, 4), (2, 1, 9))]


@pytest.fixture(scope="module", autouse=True)
def _ctx():
    lbx.init()
    yield
    lbx.set_option(lbx.OPT_COLLIDE_LITERAL, 0)


def rand_mf(boxes, ncomp, ng, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    m = ao.MultiFab(boxes, ncomp, ng, dtype=dtype)
    for f in m.fabs:
        f[...] = rng.random(f.shape) + 0.5 if dtype == np.float64 else rng.integers(0, 2, f.shape)
    return m


def to_dev(m, dtype=lbx.F64):
    d = lbx.MF(m.boxes, m.ncomp, m.ng, dtype)
    d.upload(m.fabs)
    return d


def test_mf_roundtrip_and_setval():
    m = rand_mf(BOXES, 3, 2, 0)
    d = to_dev(m)
    for a, b in zip(d.download(), m.fabs):
        assert np.array_equal(a, b)
    d.setval(7.5)
    assert all(np.all(a == 7.5) for a in d.download())
    z = lbx.MF(BOXES, 15, 2)
    assert all(np.all(a == 0.0) for a in z.download())       # NEW_FAB_FILL = 0


@pytest.mark.parametrize("literal", [1, 0])
def test_mf_equilibrium_moments_collide(coracle, literal):
    lbx.set_option(lbx.OPT_COLLIDE_LITERAL, literal)
    rho, u = rand_mf(BOXES, 1, 0, 1), rand_mf(BOXES, 3, 0, 2)
    for f in u.fabs:
        f[...] = 0.05 * (f - 1.0)
    F, R, U = lbx.MF(BOXES, 15, 2), to_dev(rho), to_dev(u)
    lbx.mf_equilibrium(F, R, U)
    got = F.download()
    for i in range(len(BOXES)):
        want = coracle.equilibrium(rho.fabs[i][0], u.fabs[i])
        assert np.all(got[i][:, :2] == 0) and np.all(got[i][:, :, :, -2:] == 0)      # ghosts untouched
        core = got[i][:, 2:-2, 2:-2, 2:-2]
        assert np.array_equal(core, want) if literal else np.max(np.abs(core - want) / want) < 1e-12
    # perturb, then masked collide in place + moments
    f = rand_mf(BOXES, 15, 2, 3)
    for i in range(len(BOXES)):
        f.fabs[i][...] = 0.02 * f.fabs[i] + np.pad(coracle.equilibrium(rho.fabs[i][0], u.fabs[i]),
                                                   ((0, 0), (2, 2), (2, 2), (2, 2)), constant_values=0.1)
    mask = rand_mf(BOXES, 1, 2, 4, dtype=np.int32)
    F.upload(f.fabs)
    M = to_dev(mask, lbx.I32)
    lbx.mf_collide(F, 1.7, 1.2, mask=M, fine_val=1)
    got = F.download()
    for i in range(len(BOXES)):
        want = f.fabs[i].copy()
        v = want[:, 2:-2, 2:-2, 2:-2]
        c = coracle.collide(np.ascontiguousarray(v), 1.7, 1.2)
        c[:, mask.fabs[i][0, 2:-2, 2:-2, 2:-2] == 1] = 0.0
        v[...] = c
        assert np.array_equal(got[i], want) if literal else np.max(np.abs(got[i] - want)) < 1e-12
    lbx.mf_moments(F, R, U)
    r, v = R.download(), U.download()
    for i in range(len(BOXES)):
        core = np.ascontiguousarray(got[i][:, 2:-2, 2:-2, 2:-2])
        keep = mask.fabs[i][0, 2:-2, 2:-2, 2:-2] != 1
        ro, uo = coracle.moments(core)
        assert np.max(np.abs(r[i][0][keep] - ro[keep]) / ro[keep]) < 1e-12
        assert np.max(np.abs(v[i][:, keep] - uo[:, keep])) < 1e-12


def test_mf_stream_zero_invalid_zero_ring():
    f = rand_mf(BOXES, 15, 2, 5)
    S, D = to_dev(f), lbx.MF(BOXES, 15, 2)
    D.setval(9.0)                                   # stale content must be overwritten / zeroed
    lbx.mf_stream(S, D)
    got = D.download()
    for i, b in enumerate(BOXES):
        want = np.zeros_like(f.fabs[i])
        for p in range(15):
            cx, cy, cz = ao.C[p]
            a = f.fabs[i][p]
            want[p, 1:-1, 1:-1, 1:-1] = a[1 - cz:a.shape[0] - 1 - cz, 1 - cy:a.shape[1] - 1 - cy, 1 - cx:a.shape[2] - 1 - cx]
        assert np.array_equal(got[i], want)
    # ZeroInvalidComponents and the comp-0 ring zeroing against the oracle's restatement
    sim = ao.AmrSimOracle(4, 4, 4, 0, 0.5, 0.5, coracle=object())
    sim.levels[0].next_f = rand_mf(BOXES, 15, 2, 6)
    ref = [a.copy() for a in sim.levels[0].next_f.fabs]
    S.upload(ref)
    sim.zero_invalid_components(0)
    lbx.mf_zero_invalid(S)
    for a, w in zip(S.download(), sim.levels[0].next_f.fabs):
        assert np.array_equal(a, w)
    S.upload(ref)
    lbx.mf_zero_ring(S, 1, 0)
    for a, w in zip(S.download(), ref):
        w = w.copy()
        z = w[0]
        z[0] = z[-1] = 0
        z[:, 0] = z[:, -1] = 0
        z[:, :, 0] = z[:, :, -1] = 0
        assert np.array_equal(a, w)


def test_plan_copy_pc_avg_const_and_ordering():
    cb = [((0, 0, 0), (7, 7, 7))]
    fb = [((4, 4, 4), (11, 11, 11)), ((12, 4, 4), (15, 11, 11))]
    crse, fine = rand_mf(cb, 2, 2, 7), rand_mf(fb, 2, 2, 8)
    Cd, Fd = to_dev(crse), to_dev(fine)
    # PC interpolation of the whole grown fine boxes from the (grown) coarse fab, then the
    # neighbour's valid cells copied on top (last match wins)
    descs = []
    for k in range(2):
        g = ao.grow(fb[k], 2)
        descs.append(dict(dst_fab=k, src_set=1, src_fab=0, kind=lbx.G_PC, ratio=2, lo=g[0], hi=g[1]))
        r = ao.isect(g, fb[1 - k])
        descs.append(dict(dst_fab=k, src_set=0, src_fab=1 - k, kind=lbx.G_COPY, lo=r[0], hi=r[1]))
    D = lbx.MF(fb, 2, 2)
    lbx.Plan(descs).apply(D, Fd, Cd, lbx.OP_COPY)
    want = ao.MultiFab(fb, 2, 2)
    for k in range(2):
        g = ao.grow(fb[k], 2)
        cg = ao.coarsen(g, 2)
        big = crse.view(0, cg).repeat(2, 1).repeat(2, 2).repeat(2, 3)
        full = ao.refine(cg, 2)
        sl = tuple(slice(g[0][d] - full[0][d], g[1][d] - full[0][d] + 1) for d in (2, 1, 0))
        want.fabs[k][...] = big[(slice(None),) + sl]
        r = ao.isect(g, fb[1 - k])
        want.view(k, r)[...] = fine.view(1 - k, r)
    for a, w in zip(D.download(), want.fabs):
        assert np.array_equal(a, w)
    # AVG (fine -> coarse, ADD twice from overlapping sources) equals sum_fine_to_coarse
    crse2 = rand_mf(cb, 2, 2, 9)
    Cd.upload(crse2.fabs)
    descs = []
    for i in range(2):
        cf = ao.grow(ao.coarsen(fb[i], 2), 1)
        r = ao.isect(cf, cb[0])
        descs.append(dict(dst_fab=0, src_fab=i, kind=lbx.G_AVG, ratio=2, lo=r[0], hi=r[1]))
    lbx.Plan(descs).apply(Cd, Fd, None, lbx.OP_ADD)
    ao.sum_fine_to_coarse(fine, crse2, (0, 0, 0))
    assert np.max(np.abs(Cd.download()[0] - crse2.fabs[0])) < 1e-15
    assert np.array_equal(Cd.download()[0], crse2.fabs[0])      # same summation order -> same bits
    # CONST on an int set, periodic-shift COPY on a double set
    Md = lbx.MF(cb, 1, 2, lbx.I32)
    lbx.Plan([dict(dst_fab=0, kind=lbx.G_CONST, lo=(-2, -2, -2), hi=(3, 3, 3), value=1)]).apply(Md)
    m = Md.download()[0][0]
    assert m[:6, :6, :6].all() and m.sum() == 216
    Cd.upload(crse.fabs)
    lbx.Plan([dict(dst_fab=0, src_fab=0, kind=lbx.G_COPY, shift=(0, 0, 8), lo=(0, 0, -2), hi=(7, 7, -1))]).apply(Cd, Cd)
    got = Cd.download()[0]
    assert np.array_equal(got[:, 0:2, 2:-2, 2:-2], crse.fabs[0][:, 8:10, 2:-2, 2:-2])


def test_fill_boundary_resolved_table_equals_descriptor_search():
    """FillBoundary plans (same-set COPY into the ghost shell) run from the resolved per-ghost-cell table
    (k_shell_copy); the descriptor kernel (LBX_OPT_ROW_KERNEL = 0) is the same copy -> same bytes.  Partly periodic
    domain: ghost cells no descriptor covers keep their values on both paths."""
    n, b = 16, 8
    boxes = [((i, j, k), (i + b - 1, j + b - 1, k + b - 1)) for k in range(0, n, b) for j in range(0, n, b) for i in range(0, n, b)]
    index = {lo: q for q, (lo, _) in enumerate(boxes)}
    descs = []
    for k, (lo, hi) in enumerate(boxes):
        g = ao.grow((lo, hi), 2)
        for dz in (-b, 0, b):
            for dy in (-b, 0, b):
                for dx in (-b, 0, b):
                    nlo = (lo[0] + dx, lo[1] + dy, lo[2] + dz)
                    if (dx, dy, dz) == (0, 0, 0) or not 0 <= nlo[2] < n:      # z is not periodic
                        continue
                    wrapped = tuple(v % n for v in nlo)
                    r = ao.isect(g, (nlo, tuple(v + b - 1 for v in nlo)))
                    descs.append(dict(dst_fab=k, src_fab=index[wrapped], kind=lbx.G_COPY,
                                      shift=tuple(w - v for w, v in zip(wrapped, nlo)), lo=r[0], hi=r[1]))
    plan = lbx.Plan(descs)
    for ncomp in (15, 3):
        m = rand_mf(boxes, ncomp, 2, 40 + ncomp)
        out = []
        for rows in (1, 0):
            lbx.set_option(lbx.OPT_ROW_KERNEL, rows)
            D = to_dev(m)
            plan.apply(D, D)
            out.append(D.download())
        lbx.set_option(lbx.OPT_ROW_KERNEL, 1)
        for k, (a, w) in enumerate(zip(*out)):
            assert np.array_equal(a, w)
            assert np.array_equal(a[:, 2:-2, 2:-2, 2:-2], m.fabs[k][:, 2:-2, 2:-2, 2:-2])     # valid cells untouched
        lo_z = [k for k, (lo, _) in enumerate(boxes) if lo[2] == 0]
        for k in lo_z:                                   # below the non-periodic face: nobody's ghost cells
            assert np.array_equal(out[0][k][:, :2], m.fabs[k][:, :2])
        # a filled ghost cell: the +x neighbour's first valid column
        k0, k1 = index[(0, 0, 0)], index[(8, 0, 0)]
        assert np.array_equal(out[0][k0][:, 2:-2, 2:-2, -2:], m.fabs[k1][:, 2:-2, 2:-2, 2:4])


def test_plan_validation_is_loud():
    cb = [((0, 0, 0), (7, 7, 7))]
    A, B = lbx.MF(cb, 1, 1), lbx.MF(cb, 1, 1)
    with pytest.raises(lbx.LbxError):      # region outside the destination fab
        lbx.Plan([dict(dst_fab=0, kind=lbx.G_COPY, lo=(-3, 0, 0), hi=(0, 0, 0))]).apply(A, B)
    with pytest.raises(lbx.LbxError):      # mapped source outside the source fab
        lbx.Plan([dict(dst_fab=0, kind=lbx.G_COPY, shift=(5, 0, 0), lo=(4, 0, 0), hi=(8, 0, 0))]).apply(A, B)
    with pytest.raises(lbx.LbxError):      # missing source set
        lbx.Plan([dict(dst_fab=0, kind=lbx.G_COPY, lo=(0, 0, 0), hi=(1, 1, 1))]).apply(A)


# ------------------------------------------------------------------ kernels of the section-8f additions
def test_mf_linear_moments_generic_derived_variables(coracle):
    """lbx_mf_linear_moments: one kernel for every linear-moment derived variable
    (include/derived_var.h:55-91, d3q15_bgk.h:34-55).  Separately rounded, fixed order: bit-identical
    to the numpy statement; density / velocity rows agree with the dedicated moments kernel."""
    f = rand_mf(BOXES, 15, 2, 11)
    F = to_dev(f)
    M, _, c, _ = lbx.tables()
    cases = [("density", np.ones((1, 15)), False), ("momentum", np.asarray(c, dtype=np.float64).T, False),
             ("velocity", np.asarray(c, dtype=np.float64).T, True), ("modes 0-9", np.asarray(M)[:10], False)]
    for name, w, norm in cases:
        out = lbx.MF(BOXES, w.shape[0], 0)
        lbx.mf_linear_moments(F, out, w, norm)
        got = out.download()
        for i in range(len(BOXES)):
            want = orc.linear_moments(f.fabs[i][:, 2:-2, 2:-2, 2:-2], w, norm)
            assert np.array_equal(got[i], want), name
    # the dedicated kernel (CalcHydroVars) and the generic path agree to rounding
    R, U = lbx.MF(BOXES, 1, 0), lbx.MF(BOXES, 3, 0)
    lbx.mf_moments(F, R, U)
    V = lbx.MF(BOXES, 3, 0)
    lbx.mf_linear_moments(F, V, np.asarray(c, dtype=np.float64).T, True)
    for a, b in zip(U.download(), V.download()):
        assert np.max(np.abs(a - b)) < 1e-14
    with pytest.raises(lbx.LbxError):
        lbx.mf_linear_moments(F, lbx.MF(BOXES, 2, 0), np.ones((3, 15)))      # more rows than output components


def test_mf_lincomb_and_tag_gradient():
    x, y = rand_mf(BOXES, 15, 2, 21), rand_mf(BOXES, 15, 2, 22)
    X, Y, D = to_dev(x), to_dev(y), lbx.MF(BOXES, 15, 2)
    lbx.mf_lincomb(D, 0.25, X, 0.75, Y)
    for i, got in enumerate(D.download()):
        want = np.zeros_like(got)                                            # ghost cells are not written
        want[:, 2:-2, 2:-2, 2:-2] = 0.25 * x.fabs[i][:, 2:-2, 2:-2, 2:-2] + 0.75 * y.fabs[i][:, 2:-2, 2:-2, 2:-2]
        assert np.array_equal(got, want)
    rho = rand_mf(BOXES, 1, 1, 23)
    R, T = to_dev(rho), lbx.MF(BOXES, 1, 0, lbx.I32)
    thr = 0.35
    lbx.mf_tag_gradient(R, thr, T, 2)
    ntag = 0
    for i, got in enumerate(T.download()):
        r = rho.fabs[i][0]
        gx, gy, gz = r[1:-1, 1:-1, 2:] - r[1:-1, 1:-1, :-2], r[1:-1, 2:, 1:-1] - r[1:-1, :-2, 1:-1], r[2:, 1:-1, 1:-1] - r[:-2, 1:-1, 1:-1]
        want = np.where(0.25 * ((gx * gx + gy * gy) + gz * gz) > thr * thr, 2, 0)
        assert np.array_equal(got[0], want)
        ntag += int((want == 2).sum())
    assert 0 < ntag < sum(ao.numpts(b) for b in BOXES)
