"""The AMR configurations of BASELINE.json as one timed function, shared by bench.py's AMR leg (`extra.amr` of
the one JSON line) and tools/amr_bench.py:

  C4 (configs[3]): 2-level pulse, ratio 2, static box = central half of the domain
                   (/root/reference/tests/catch2AMRTests.cpp:386-387), one GPU;
  C5 (configs[4]): 3-level pulse, 512^3 base, nested central boxes translated every `regrid_every` coarse steps
                   (the reference regrids only inside Set/UnsetStaticRefinement, src/AmrSim.cpp:1003,1013),
                   boxes of every level distributed over the ranks.

Everything goes through AmrSim (liblambrex.so).  MLUPS counts sum_l cells_l x substeps_l per coarse step
(SURVEY.md 8d), device-timed with CUDA events on the library's stream, max over ranks.  The roofline block is
the fused per-level kernel k_mf_collide_stream: 240 B x the VALID cells each launch updates / its time, both
measured live (lbx_prof_begin / lbx_prof_end bracket every launch with events), against the measured HBM peak."""
import time

import numpy as np

from . import amrsim, lbx, workloads


def static_boxes(n, levels, shift=0):
    """nested central-half boxes (tests/catch2AMRTests.cpp:386-387), optionally translated"""
    out = []
    lo, hi = n // 4 + shift, 3 * n // 4 - 1 + shift
    for _ in range(levels - 1):
        out.append(((lo,) * 3, (hi,) * 3))
        lo, hi = 2 * lo + (hi - lo + 1) // 2, 2 * lo + (hi - lo + 1) // 2 + (hi - lo)      # central half again
    return out


def _cells(sim, lev):
    return sum(int(np.prod([h - l + 1 for l, h in zip(*b)])) for b in sim.boxArray(lev))


def run_amr_case(n, levels, steps, warmup=3, coupling="rohde", regrid_every=0, max_grid=32, fused=True, gradient=0.0,
                 dist=None, peak_gbs=None):
    """One timed AMR job; returns the result dict on every rank (identical numbers after the reductions)."""
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    sim = amrsim.AmrSim(n, n, n, levels - 1, (1, 1, 1), 0.5, 0.5)
    sim.SetMaxGridSize(max_grid)
    sim.SetRohdeFusion(fused)
    sim.SetCoupling(amrsim.SUBCYCLE if coupling == "subcycle" else amrsim.ROHDE)
    # the planar pulse varies along z only: stated as a profile (no whole-domain host array at 512^3 per rank)
    sim.SetInitialDensityProfile(2, workloads.pulse_density(2, 2, n).reshape(2, 2, n)[0, 0, :])
    sim.SetInitialVelocityProfile(2, np.zeros((n, 3)))
    sim.InitFromScratch(0.0)
    t0 = time.perf_counter()
    if gradient > 0:
        sim.SetGradientRefinement(0, gradient)
        for lev, (lo, hi) in list(enumerate(static_boxes(n, levels)))[1:]:
            sim.SetStaticRefinement(lev, lo, hi)
        sim.SetRegridInterval(regrid_every)
    else:
        for lev, (lo, hi) in enumerate(static_boxes(n, levels)):
            sim.SetStaticRefinement(lev, lo, hi)
    lbx.sync()
    setup_regrid_s = time.perf_counter() - t0
    cells = [_cells(sim, l) for l in range(levels)]
    nbox = [len(sim.boxArray(l)) for l in range(levels)]
    substeps = [1] + [2 ** l for l in range(1, levels)]
    t0 = time.perf_counter()
    sim.Iterate(warmup)             # includes the FLAT -> BOXES re-layout and plan building
    lbx.sync()
    first_s = time.perf_counter() - t0
    if regrid_every > 0 and gradient <= 0:
        # one untimed regrid: the steady state of a periodically regridding run reuses the device blocks (and,
        # distributed, the CUDA-IPC mappings) the previous regrid released
        for lev, (lo, hi) in enumerate(static_boxes(n, levels, shift=2)):
            sim.SetStaticBox(lev, lo, hi)
        sim.Regrid()
        sim.Iterate(1)
        lbx.sync()
    if dist:
        dist.barrier()
    l0 = lbx.launch_count()
    regrids, regrid_host_s = 0, 0.0
    lbx.prof_begin()
    with lbx.Timer() as t:
        if gradient > 0:
            sim.Iterate(steps)          # regrid_int inside Iterate
            regrids = sim.NumRegrids()
        elif regrid_every > 0:
            done = 0
            while done < steps:
                k = min(regrid_every, steps - done)
                sim.Iterate(k)
                done += k
                if done < steps:
                    r0 = time.perf_counter()
                    regrids += 1
                    for lev, (lo, hi) in enumerate(static_boxes(n, levels, shift=regrids % 3)):
                        sim.SetStaticBox(lev, lo, hi)       # every level's box moves, then ONE regrid from level 0
                    sim.Regrid()
                    regrid_host_s += time.perf_counter() - r0
        else:
            sim.Iterate(steps)
    prof = lbx.prof_end()
    launches = lbx.launch_count() - l0
    ms = t.ms
    kern_ms, kern_cells = prof["ms"], prof["valid_cells"]
    per_rank_kernel_ms = [round(kern_ms, 3)]
    if dist:
        import torch
        tt = torch.tensor([ms, regrid_host_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, regrid_host_s = float(tt[0]), float(tt[1])
        # the fused passes' summed time on every rank: tells load imbalance (spread) from exchange cost (all alike)
        kk = torch.zeros(world, dtype=torch.float64)
        kk[rank] = kern_ms
        dist.all_reduce(kk, op=dist.ReduceOp.SUM)
        per_rank_kernel_ms = [round(float(v), 3) for v in kk]
    work = sum(c * s for c, s in zip(cells, substeps))      # cells of the initial grids; regrids change them little
    cells_end = [_cells(sim, l) for l in range(sim.finestLevel() + 1)]
    # a cheap physics check without a whole-domain array: mean density of level 0 from one x-row per (y, z)?  The
    # dense getter is fine up to 256^3; larger bases report the mass of rank 0's view of a sub-sampled field instead
    check = {}
    sim.CalcHydroVars(0)
    if n <= 256:
        check["mean_rho_level0"] = float(sim.GetDensityField(0).mean())
    achieved = 240.0 * kern_cells / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    res = {"metric": "MLUPS (fp64 D3Q15 BGK, sum_l cells_l x substeps_l)", "value": work * steps / (ms * 1e-3) / 1e6, "unit": "MLUPS",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "levels": levels, "coupling": coupling,
           "fused": bool(fused), "base_grid": [n, n, n], "max_grid": max_grid, "cells_per_level": cells,
           "cells_per_level_at_end": cells_end, "boxes_per_level": nbox, "substeps": substeps,
           "launches_per_step": launches / steps, "regrid_every": regrid_every, "regrids_in_timed_region": regrids,
           "regrid_host_seconds_in_timed_region": regrid_host_s,
           "regrid_seconds_per_event": regrid_host_s / regrids if regrids else None,
           "setup_regrid_seconds": setup_regrid_s, "first_%d_steps_seconds" % warmup: first_s,
           "gradient_threshold": gradient, "check": check,
           "roofline": {"bound": "hbm", "kernel": "k_mf_collide_stream (rank 0's boxes)", "achieved": achieved, "peak": peak_gbs,
                        "unit": "GB/s", "frac": achieved / peak_gbs if peak_gbs else None,
                        "launches": prof["launches"], "kernel_ms_total": kern_ms, "kernel_ms_total_per_rank": per_rank_kernel_ms, "share_of_timed_region": kern_ms / t.ms if t.ms else None,
                        "algorithmic_bytes": 240.0 * kern_cells, "not_bracketed": prof["dropped"],
                        "rank0_ms_by_kind": {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in prof["by_kind"].items()},
                        "how": "240 B x valid cells of every launch / summed launch time, CUDA events around each launch on its stream"}}
    sim.close()
    return res
