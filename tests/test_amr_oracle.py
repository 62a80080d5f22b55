"""CPU tests of the AMR oracle (oracle/amr_oracle.py) against what the reference itself pins:
the pulse golden vectors through the box-decomposed pass structure
(/root/reference/tests/catch2RegressionTests.cpp:6-93), the uniform-field, ladder, tag-set
and coverage checks of /root/reference/tests/catch2AMRTests.cpp, and internal consistency
(box-decomposed step == single periodic box, bit for bit)."""
import os

import numpy as np
import pytest

from conftest import approx_catch2
from lambrex_b200 import workloads
from oracle import amr_oracle as ao
from oracle import lbm_oracle as lo


def test_box_decomposed_pulse_reproduces_golden_and_periodic_oracle(golden_dir, coracle):
    g = np.load(os.path.join(golden_dir, "pulse_regression.npz"))
    nx, ny, nz = 10, 10, 50
    sim = ao.AmrSimOracle(nx, ny, nz, 0, 0.5, 0.5, coracle=coracle)
    sim.set_initial_density(workloads.pulse_density(nx, ny, nz))
    sim.set_initial_velocity(0.0)
    sim.init_from_scratch(0.0)
    assert sim.grids[0] == [((0, 0, 0), (9, 9, 23)), ((0, 0, 24), (9, 9, 49))]
    f = coracle.equilibrium(lo.user_to_fab(workloads.pulse_density(nx, ny, nz), nx, ny, nz), np.zeros((3, nz, ny, nx)))
    for t in (100, 200):
        sim.iterate(100)
        sim.calc_hydro_vars(0)
        assert approx_catch2(sim.gather_valid(0, "rho")[0].reshape(-1), g["RHO_t%d" % t]).all()
        f = coracle.step(f, 1.0, 1.0, 100)
        assert np.array_equal(sim.gather_valid(0, "f"), f)
    assert sim.levels[0].time == 200.0 and sim.levels[0].step == 200


def two_level(coracle, tau=0.01):
    sim = ao.AmrSimOracle(48, 24, 12, 1, tau, tau, coracle=coracle)
    sim.set_initial_density(0.8)
    sim.set_initial_velocity(0.1)
    sim.init_from_scratch(0.0)
    return sim


def test_two_level_initialisation_and_ladder(coracle):
    """tests/catch2AMRTests.cpp:118-161, 306-355."""
    sim = two_level(coracle)
    assert sim.finest_level == 0 and sim.levels[1].now_f is None
    ba = sim.grids[0]
    sim.make_new_level_from_coarse(1, sim.levels[0].time, ba)
    assert sim.levels[1].time == 0.0
    for (i, j, k) in [(0, 0, 0), (47, 23, 11), (13, 7, 5)]:
        assert sim.get_density(i, j, k, 1) == pytest.approx(0.8, rel=1e-12)
        for n in range(3):
            assert sim.get_velocity(i, j, k, n, 1) == pytest.approx(0.1, rel=1e-12)
    assert sim.levels[1].delta == 0.5 and sim.mass[1] == 0.5
    assert sim.tau_s[1] == pytest.approx(2 * (0.01 - 0.5) + 0.5)


def test_static_refinement_grids_and_tags(coracle):
    """tests/catch2AMRTests.cpp:209-304 (tag sets) and 385-422 (coverage inequalities)."""
    sim = two_level(coracle)
    nx, ny, nz = sim.n
    lo_c, hi_c = (nx // 4, ny // 4, nz // 4), (3 * nx // 4, 3 * ny // 4, 3 * nz // 4)
    sim.set_static_refinement(0, lo_c, hi_c)
    assert sim.finest_level == 1
    mb = ao.minimal_box(sim.grids[1])
    assert ao.contains_pt(mb, tuple(2 * v for v in lo_c)) and ao.contains_pt(mb, tuple(2 * v for v in hi_c))
    assert ao.numpts(mb) >= 8 * np.prod([h - l for l, h in zip(lo_c, hi_c)])
    # every fine box is a refined coarse box of at most 32 cells per side
    for b in sim.grids[1]:
        assert all(l % 2 == 0 and h % 2 == 1 and h - l + 1 <= 32 for l, h in zip(*b))
    # tags: SET exactly inside the static box
    tags = {"ng": 0, "fabs": [np.full(tuple(h - l + 1 for l, h in zip(*b))[::-1], ao.TAG_SET, dtype=np.uint8)
                              for b in sim.grids[0]]}
    sim.error_est(0, tags)
    for b, t in zip(sim.grids[0], tags["fabs"]):
        zz, yy, xx = np.meshgrid(*[np.arange(b[0][d], b[1][d] + 1) for d in (2, 1, 0)], indexing="ij")
        inside = ((xx >= lo_c[0]) & (xx <= hi_c[0]) & (yy >= lo_c[1]) & (yy <= hi_c[1]) & (zz >= lo_c[2]) & (zz <= hi_c[2]))
        assert np.array_equal(t == ao.TAG_SET, inside) and np.array_equal(t == ao.TAG_CLEAR, ~inside)
    sim.unset_static_refinement(0)
    assert sim.finest_level == 0 and sim.levels[1].now_f is None


def test_full_domain_refinement_covers_everything(coracle):
    sim = two_level(coracle)
    sim.set_static_refinement(0, (0, 0, 0), (47, 23, 11))
    assert sum(ao.numpts(b) for b in sim.grids[1]) == 8 * 48 * 24 * 12
    assert ao.minimal_box(sim.grids[1]) == ((0, 0, 0), (95, 47, 23))


def test_max_size_and_cluster_examples():
    assert ao.max_size([((0, 0, 0), (4, 4, 24))], 16) == [((0, 0, 0), (4, 4, 11)), ((0, 0, 12), (4, 4, 24))]
    assert len(ao.make_base_grids(((0, 0, 0), (255, 255, 255)))) == 512
    # two separated blobs -> two boxes (hole cut)
    pts = np.array([(i, j, k) for i in range(4) for j in range(4) for k in range(4)] +
                   [(i + 10, j, k) for i in range(3) for j in range(4) for k in range(4)])
    boxes, _ = ao.cluster(pts, 0.7)
    assert sorted(boxes) == [((0, 0, 0), (3, 3, 3)), ((10, 0, 0), (12, 3, 3))]


def test_sum_fine_to_coarse_and_fillpatch_semantics():
    """Uniform fine field: every coarse cell under fine valid cells receives exactly the fine
    value; coarse cells under fine GHOST cells also receive it (appendix C)."""
    cb = [((0, 0, 0), (7, 7, 7))]
    fb = [((4, 4, 4), (11, 11, 11))]                   # fine box = coarse cells 2..5
    crse, fine = ao.MultiFab(cb, 1, 2), ao.MultiFab(fb, 1, 2)
    fine.fabs[0][...] = 3.0
    ao.sum_fine_to_coarse(fine, crse, (8, 8, 8))
    v = crse.valid(0)[0]
    assert np.all(v[1:7, 1:7, 1:7] == 3.0)             # coarsen(fine grown by 2) = cells 1..6
    assert np.all(v[0] == 0.0) and np.all(v[7] == 0.0)
    # FillPatchTwoLevels: fine valid from fine, ghosts from coarse
    crse.fabs[0][...] = 5.0
    dst = ao.MultiFab(fb, 1, 2)
    ao.fillpatch_two(dst, crse, fine, (8, 8, 8), (16, 16, 16))
    assert np.all(dst.valid(0) == 3.0)
    d = dst.fabs[0][0].copy()
    d[2:-2, 2:-2, 2:-2] = 5.0
    assert np.all(d == 5.0)


def test_two_level_pulse_runs_and_clocks_advance(coracle):
    nx, ny, nz = 16, 16, 32
    sim = ao.AmrSimOracle(nx, ny, nz, 1, 0.5, 0.5, coracle=coracle)
    sim.set_initial_density(workloads.pulse_density(nx, ny, nz))
    sim.set_initial_velocity(0.0)
    sim.init_from_scratch(0.0)
    sim.set_static_refinement(0, (4, 4, 8), (12, 12, 24))
    sim.iterate(3)
    assert (sim.levels[0].time, sim.levels[0].step) == (3.0, 3)
    assert (sim.levels[1].time, sim.levels[1].step) == (3.0, 6)
    assert np.isfinite(sim.gather_valid(0, "f")).all()


# ------------------------------------------------------------------ conventional subcycling
def test_average_down_inverts_piecewise_constant_interpolation(coracle):
    rng = np.random.default_rng(5)
    cboxes = [((0, 0, 0), (7, 7, 3)), ((0, 0, 4), (7, 7, 7))]
    crse = ao.MultiFab(cboxes, ao.NV, ao.HALO)
    for i in range(len(cboxes)):
        crse.valid(i)[...] = rng.random(crse.valid(i).shape)
    fboxes = [((4, 4, 4), (11, 11, 11)), ((4, 4, 2), (11, 11, 3))]
    fine = ao.MultiFab(fboxes, ao.NV, ao.HALO)
    ao.pc_interp_fill(fine, crse, (8, 8, 8))
    out = ao.MultiFab(cboxes, ao.NV, ao.HALO, fill=-7.0)
    ao.average_down(fine, out)
    covered = 0
    for i, b in enumerate(cboxes):
        for fb in fboxes:
            r = ao.isect(ao.coarsen(fb, 2), b)
            if r is not None:
                # sequential sum of 8 equal values rounds at 3x, 5x, 6x(exact), 7x: a few ulp
                assert np.max(np.abs(out.view(i, r) - crse.view(i, r))) <= 4e-16
                covered += ao.numpts(r)
    assert covered == (8 * 8 * 8 + 8 * 8 * 2) // 8
    assert sum(int((out.valid(i)[0] != -7.0).sum()) for i in range(2)) == covered     # nothing else written


def test_subcycle_uniform_state_is_a_fixed_point_and_clocks(coracle):
    sim = ao.AmrSimOracle(48, 24, 12, 1, 0.3, 0.3, coracle=coracle)
    sim.set_initial_density(0.8)
    sim.set_initial_velocity(0.0)      # at rest the initial equilibrium is a fixed point of the collision
    sim.init_from_scratch(0.0)
    sim.coupling = "subcycle"
    sim.set_static_refinement(0, (12, 6, 3), (35, 17, 8))
    assert sim.finest_level == 1
    ref = sim.levels[1].now_f.valid(0).copy()
    sim.iterate(2)
    assert (sim.levels[0].time, sim.levels[0].step) == (2.0, 2)
    assert (sim.levels[1].time, sim.levels[1].step) == (2.0, 4)
    for lev in (0, 1):
        L = sim.levels[lev]
        for i in range(len(L.boxes)):
            assert np.max(np.abs(L.now_f.valid(i) - ref[:, :1, :1, :1])) < 1e-15


def test_subcycle_time_interpolation_of_coarse_ghost_data(coracle):
    """coarse_state_at: old state at the first fine substep, the half-half LinComb at the second."""
    nx, ny, nz = 16, 8, 8
    sim = ao.AmrSimOracle(nx, ny, nz, 1, 0.4, 0.4, coracle=coracle)
    sim.coupling = "subcycle"
    rho, u = workloads.shear_wave(nx, ny, nz)
    sim.set_initial_density(rho)
    sim.set_initial_velocity(u)
    sim.init_from_scratch(0.0)
    sim.set_static_refinement(0, (4, 2, 2), (11, 5, 5))
    sim.iterate_level(0)
    L0 = sim.levels[0]
    assert sim.coarse_state_at(0, 0.0) is L0.next_f and sim.coarse_state_at(0, 1.0) is L0.now_f
    mid = sim.coarse_state_at(0, 0.5)
    assert np.array_equal(mid.valid(0), 0.5 * L0.next_f.valid(0) + 0.5 * L0.now_f.valid(0))
    sim.coupling = "rohde"
    assert sim.coarse_state_at(0, 0.5) is L0.now_f          # the reference: NOW, whatever its time


# ------------------------------------------------------------------ dynamic refinement
def test_gradient_refinement_follows_the_pulse(coracle):
    nx, ny, nz = 8, 8, 48
    sim = ao.AmrSimOracle(nx, ny, nz, 1, 0.5, 0.5, max_grid_size=16, coracle=coracle)
    sim.coupling = "subcycle"
    sim.set_initial_density(workloads.pulse_density(nx, ny, nz))
    sim.set_initial_velocity(0.0)
    sim.init_from_scratch(0.0)
    sim.set_gradient_refinement(0, 2e-4)
    assert sim.finest_level == 1
    # the planar pulse sits at k = nz/2 - 1: only planes next to it are tagged, over the whole x-y extent
    zs = sorted({(b[0][2], b[1][2]) for b in sim.grids[1]})
    assert zs == [(2 * (nz // 2 - 3), 2 * (nz // 2 + 1) + 1)]
    assert sum(ao.numpts(b) for b in sim.grids[1]) == (2 * nx) * (2 * ny) * (zs[0][1] - zs[0][0] + 1)
    sim.regrid_int = 4
    sim.iterate(8)
    assert sim.num_regrids == 2 and sim.finest_level == 1
    z2 = sorted({(b[0][2], b[1][2]) for b in sim.grids[1]})
    assert z2 != zs and min(z[0] for z in z2) < zs[0][0] and max(z[1] for z in z2) > zs[0][1]   # two pulses moving apart
    sim.calc_hydro_vars(0)
    assert abs(sim.gather_valid(0, "rho")[0].mean() - 1.0) < 1e-4
    # threshold above every gradient: the fine level disappears at the next regrid
    sim.gradient_threshold[0] = 1.0
    sim.iterate(4)
    assert sim.finest_level == 0 and sim.levels[1].now_f is None
    sim.iterate(2)                               # and the run continues on one level
    assert sim.levels[0].step == 14
