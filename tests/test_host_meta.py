"""CPU tests of the C++ host library's grid-generation metadata (liblambrex.so, no GPU):
bit-exact agreement of box lists with the oracle's restatement -- base grids, maxSize,
simplify, complement, Berger-Rigoutsos clustering, and whole regrid sequences with the
reference's static-box tagging -- plus the C-ABI export check for include/lambrex_c.h."""
import os
import re

import numpy as np
import pytest

from lambrex_b200 import amrsim
from oracle import amr_oracle as ao

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = amrsim.lib()
    txt = open(os.path.join(ROOT, "include", "lambrex_c.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(lbx_(?:sim|meta)_[a-z0-9_]+)\s*\(", txt))
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), "liblambrex.so does not export " + n
    assert names == set(amrsim.SYMBOLS), names ^ set(amrsim.SYMBOLS)


def test_sim_needs_initialised_gpu_context():
    from lambrex_b200 import lbx
    import ctypes
    n = ctypes.c_int(0)
    lbx.lib().lbx_device_count(ctypes.byref(n))
    if n.value == 0:
        with pytest.raises(amrsim.LambrexError):
            amrsim.lambrexInit()
        with pytest.raises(amrsim.LambrexError):
            amrsim.AmrSim(4, 4, 4, 0, (1, 1, 1), 0.5, 0.5)


@pytest.mark.parametrize("dims", [(10, 10, 50), (48, 24, 12), (11, 12, 13), (256, 256, 256), (16, 9, 8), (100, 70, 33)])
def test_base_grids_match_oracle(dims):
    want = ao.make_base_grids(((0, 0, 0), tuple(d - 1 for d in dims)))
    assert amrsim.meta_base_grids(dims) == want
    assert sum(ao.numpts(b) for b in want) == int(np.prod(dims))


def test_box_calculus_matches_oracle():
    rng = np.random.default_rng(7)
    for _ in range(40):
        lo = rng.integers(-5, 10, 3)
        hi = lo + rng.integers(0, 70, 3)
        b = ao.bx(lo, hi)
        chunk = int(rng.choice([4, 8, 16, 32]))
        assert amrsim.meta_max_size([b], chunk) == ao.max_size([b], chunk)
        cuts = []
        for _ in range(3):
            cl = lo + rng.integers(-3, 40, 3)
            cuts.append(ao.bx(cl, cl + rng.integers(0, 30, 3)))
        comp = ao.complement_in(b, cuts)
        assert amrsim.meta_complement(b, cuts) == comp
        assert amrsim.meta_simplify(comp) == ao.simplify(comp)
        assert sum(ao.numpts(x) for x in ao.simplify(comp)) == sum(ao.numpts(x) for x in comp)


def test_cluster_matches_oracle():
    rng = np.random.default_rng(3)
    for trial in range(12):
        pts = set()
        for _ in range(int(rng.integers(1, 5))):
            lo = rng.integers(0, 40, 3)
            ext = rng.integers(1, 12, 3)
            for i in range(ext[0]):
                for j in range(ext[1]):
                    for k in range(ext[2]):
                        if rng.random() < 0.9:
                            pts.add((int(lo[0] + i), int(lo[1] + j), int(lo[2] + k)))
        p = np.array(sorted(pts, key=lambda v: (v[2], v[1], v[0])))
        want, _ = ao.cluster(p, 0.7)
        assert amrsim.meta_cluster(p, 0.7) == want, trial


REGRID_CASES = [
    ((48, 24, 12), 1, [("set", 0, (12, 6, 3), (36, 18, 9))]),
    ((48, 24, 12), 1, [("set", 0, (0, 0, 0), (47, 23, 11)), ("unset", 0)]),
    ((16, 16, 32), 1, [("set", 0, (4, 4, 8), (12, 12, 24)), ("set", 0, (2, 2, 2), (9, 9, 20))]),
    ((32, 32, 32), 2, [("set", 0, (8, 8, 8), (23, 23, 23)), ("set", 1, (24, 24, 24), (39, 39, 39))]),
    ((32, 32, 32), 2, [("set", 0, (0, 8, 8), (10, 23, 23)), ("set", 1, (4, 20, 20), (15, 30, 30)),
                       ("set", 0, (4, 8, 8), (14, 23, 23)), ("unset", 1)]),
    ((64, 64, 64), 2, [("set", 0, (10, 12, 14), (50, 41, 30)), ("set", 1, (30, 30, 34), (90, 70, 52))]),
]


@pytest.mark.parametrize("dims,max_level,ops", REGRID_CASES)
def test_regrid_sequences_match_oracle(dims, max_level, ops, coracle):
    mesh = amrsim.MetaMesh(dims, max_level)
    sim = ao.AmrSimOracle(*dims, max_level, 0.5, 0.5, coracle=coracle)
    sim.set_initial_density(1.0)
    sim.set_initial_velocity(0.0)
    sim.init_from_scratch(0.0)
    assert mesh.boxes(0) == sim.grids[0]
    for op in ops:
        if op[0] == "set":
            mesh.set_static(op[1], op[2], op[3])
            sim.set_static_refinement(op[1], op[2], op[3])
        else:
            mesh.unset_static(op[1])
            sim.unset_static_refinement(op[1])
        assert mesh.finest_level() == sim.finest_level, op
        for lev in range(max_level + 1):
            assert mesh.boxes(lev) == sim.grids[lev], (op, lev)
    # proper nesting: every level-(l+1) box, coarsened and grown by n_proper, lies in level l
    for lev in range(1, sim.finest_level + 1):
        for b in sim.grids[lev]:
            c = ao.grow(ao.coarsen(b, 2), 1)
            dom = sim.domain(lev - 1)
            cells = ao.complement_in(c, sim.grids[lev - 1])
            for r in cells:      # anything uncovered must be outside the domain (periodic wrap)
                assert ao.isect(r, dom) is None or lev - 1 == 0


def test_box_ownership_of_a_distributed_level():
    """amrex::DistributionMapping(ba, nprocs) as the distributed paths use it, identical on every
    rank by construction (no communication).  A BoxArray made of whole x-y layers (every level 0): each rank
    owns ONE z-slab -- consecutive layers, plane counts within one layer of the mean -- which the uniform
    path stores as a single fab per GPU.  Any other BoxArray: contiguous chunks of the list by cell count."""
    ba = amrsim.meta_base_grids((256, 256, 256))
    assert len(ba) == 512
    for nprocs in (1, 2, 3, 4, 8):
        own = amrsim.meta_distribution(ba, nprocs)
        assert set(own) == set(range(nprocs))                                 # nobody idle
        planes = []
        for r in range(nprocs):
            mine = [b for b, o in zip(ba, own) if o == r]
            zlo, zhi = min(b[0][2] for b in mine), max(b[1][2] for b in mine)
            assert sum(int(np.prod([h - l + 1 for l, h in zip(*b)])) for b in mine) == 256 * 256 * (zhi - zlo + 1)   # one full slab
            planes.append((zlo, zhi))
        assert sorted(planes) == planes and all(a[1] + 1 == b[0] for a, b in zip(planes, planes[1:]))   # rank order = z order
        sizes = [hi - lo + 1 for lo, hi in planes]
        assert max(sizes) - min(sizes) <= 32
    # ragged boxes: balance by CELLS, not by box count
    ragged = [((0, 0, 0), (63, 63, 63))] + [((64 + 8 * i, 0, 0), (71 + 8 * i, 7, 7)) for i in range(64)]
    own = amrsim.meta_distribution(ragged, 2)
    cells = [0, 0]
    for (lo, hi), r in zip(ragged, own):
        cells[r] += int(np.prod([h - l + 1 for l, h in zip(lo, hi)]))
    assert own[0] == 0 and own[-1] == 1 and cells[0] >= 64 ** 3
    # a refined level of 17 x 17 x 17 boxes (C5's levels 1 and 2: 516^3 cells chopped at 32) on 8 ranks: whole layers
    # would be 3 + 7 x 2 (1.41 x the mean on one rank); the level is cut inside layers instead, in (z, y, x) order
    edges = [0] + list(np.cumsum([30] * 3 + [31] * 2 + [30] * 12))            # 17 pieces
    fine = [((edges[i], edges[j], edges[k]), (edges[i + 1] - 1, edges[j + 1] - 1, edges[k + 1] - 1))
            for i in range(17) for j in range(17) for k in range(17)]           # deliberately NOT in z order
    own = amrsim.meta_distribution(fine, 8)
    cells = np.zeros(8)
    for (lo, hi), r in zip(fine, own):
        cells[r] += np.prod([h - l + 1 for l, h in zip(lo, hi)])
    assert cells.max() <= 1.01 * cells.mean()
    zmin = [min(b[0][2] for b, o in zip(fine, own) if o == r) for r in range(8)]
    assert zmin == sorted(zmin)                                                   # ranks follow z
    # a hierarchy's levels: two interleaved runs per rank, so that the part of a level under the next finer one (the
    # central half) is spread over ALL ranks -- with one share per rank four of eight ranks owned all of it
    for level in (ba, fine):
        own = amrsim.meta_distribution(level, 8, runs_per_rank=2)
        zs = sorted({b[0][2] for b in level})
        zlo, zhi = zs[len(zs) // 4], zs[3 * len(zs) // 4]
        cells, central = np.zeros(8), np.zeros(8)
        for (lo, hi), r in zip(level, own):
            c = np.prod([h - l + 1 for l, h in zip(lo, hi)])
            cells[r] += c
            if zlo <= lo[2] < zhi:
                central[r] += c
        assert cells.max() <= 1.02 * cells.mean()
        assert central.min() > 0 and central.max() <= 1.35 * central.mean()
