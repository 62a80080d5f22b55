"""AMR-path throughput (BASELINE.json configs[3] and [4]): multi-level pulse, refinement ratio 2,
the reference's Rohde cycle with subcycling through AmrSim.  One GPU, or -- under torchrun -- the
boxes of every level distributed over the ranks (configs[4]: 3 levels, 512^3 base, periodic regrid).
Prints one JSON line from rank 0: MLUPS counts sum_l cells_l x substeps_l per coarse step
(SURVEY.md 8d), device-timed, max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import amrsim, lbx, workloads   # noqa: E402


def static_boxes(n, levels, shift=0):
    """nested central-half boxes (tests/catch2AMRTests.cpp:386-387), optionally translated"""
    out = []
    lo, hi = n // 4 + shift, 3 * n // 4 - 1 + shift
    for _ in range(levels - 1):
        out.append(((lo,) * 3, (hi,) * 3))
        lo, hi = 2 * lo + (hi - lo + 1) // 2, 2 * lo + (hi - lo + 1) // 2 + (hi - lo)      # central half again
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--levels", type=int, default=2)
    ap.add_argument("--max-grid", type=int, default=32)
    ap.add_argument("--regrid-every", type=int, default=0, help="coarse steps between regrids (0: never)")
    ap.add_argument("--no-fusion", action="store_true", help="Rohde cycle as the literal pass sequence")
    ap.add_argument("--valid-tiling", default=None, choices=["rows", "linear"], help="valid-cell tiles of the fused pass")
    ap.add_argument("--xghost-in-row", type=int, default=0)
    ap.add_argument("--debug-skip", type=int, default=0, help="profiling only: 1 skip valid tiles, 2 skip ghost tiles")
    ap.add_argument("--coupling", default="rohde", choices=["rohde", "subcycle"],
                    help="rohde: the reference's live RohdeCycle; subcycle: conventional subcycling (FillPatch with "
                         "time interpolation, average_down)")
    ap.add_argument("--gradient", type=float, default=0.0,
                    help="> 0: refine level 0 by the density-gradient criterion with this threshold (device tagging) "
                         "instead of static boxes; regrids happen inside Iterate every --regrid-every steps")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        amrsim.lambrexInitParallel()
    else:
        amrsim.lambrexInit()
    if args.valid_tiling:
        lbx.set_option(lbx.OPT_VALID_TILING, 1 if args.valid_tiling == "linear" else 0)
    lbx.set_option(lbx.OPT_XGHOST_IN_ROW, args.xghost_in_row)
    n = args.grid
    sim = amrsim.AmrSim(n, n, n, args.levels - 1, (1, 1, 1), 0.5, 0.5)
    sim.SetMaxGridSize(args.max_grid)
    sim.SetRohdeFusion(not args.no_fusion)
    sim.SetCoupling(amrsim.SUBCYCLE if args.coupling == "subcycle" else amrsim.ROHDE)
    sim.SetInitialDensity(workloads.pulse_density(n, n, n))
    sim.SetInitialVelocity(0.0)
    sim.InitFromScratch(0.0)
    t0 = time.perf_counter()
    if args.gradient > 0:
        sim.SetGradientRefinement(0, args.gradient)
        for lev, (lo, hi) in list(enumerate(static_boxes(n, args.levels)))[1:]:
            sim.SetStaticRefinement(lev, lo, hi)
        sim.SetRegridInterval(args.regrid_every)
    else:
        for lev, (lo, hi) in enumerate(static_boxes(n, args.levels)):
            sim.SetStaticRefinement(lev, lo, hi)
    lbx.sync()
    regrid_s = time.perf_counter() - t0
    cells = [sum(int(np.prod([h - l + 1 for l, h in zip(*b)])) for b in sim.boxArray(l)) for l in range(args.levels)]
    nbox = [len(sim.boxArray(l)) for l in range(args.levels)]
    mine = [sum(1 for b in range(nbox[l]) if sim.Owner(l, b) == rank) for l in range(args.levels)]
    substeps = [1] + [2 ** l for l in range(1, args.levels)]
    t0 = time.perf_counter()
    sim.Iterate(args.warmup)        # includes the FLAT -> BOXES re-layout and plan building
    lbx.sync()
    first_s = time.perf_counter() - t0
    if args.regrid_every > 0 and args.gradient <= 0:
        # one untimed regrid: the steady state of a periodically regridding run reuses the device blocks
        # (and, distributed, the CUDA-IPC mappings) the previous regrid released
        for lev, (lo, hi) in enumerate(static_boxes(n, args.levels, shift=2)):
            sim.SetStaticRefinement(lev, lo, hi)
        sim.Iterate(1)
        lbx.sync()
    if dist:
        dist.barrier()
    lbx.set_option(lbx.OPT_DEBUG_SKIP, args.debug_skip)
    l0 = lbx.launch_count()
    regrids, regrid_in_loop_s = 0, 0.0
    with lbx.Timer() as t:
        if args.gradient > 0:
            sim.Iterate(args.steps)          # regrid_int inside Iterate
            regrids = sim.NumRegrids()
        elif args.regrid_every > 0:
            done = 0
            while done < args.steps:
                k = min(args.regrid_every, args.steps - done)
                sim.Iterate(k)
                done += k
                if done < args.steps:
                    r0 = time.perf_counter()
                    regrids += 1
                    for lev, (lo, hi) in enumerate(static_boxes(n, args.levels, shift=regrids % 3)):
                        sim.SetStaticRefinement(lev, lo, hi)
                    regrid_in_loop_s += time.perf_counter() - r0
        else:
            sim.Iterate(args.steps)
    launches = lbx.launch_count() - l0
    ms = t.ms
    if dist:
        import torch
        tt = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    work = sum(c * s for c, s in zip(cells, substeps))      # (cells of the INITIAL grids; regrids change them little)
    cells_end = [sum(int(np.prod([h - l + 1 for l, h in zip(*b)])) for b in sim.boxArray(l)) for l in range(sim.finestLevel() + 1)]
    sim.CalcHydroVars(0)
    rho = sim.GetDensityField(0)
    if rank == 0:
        print(json.dumps({"metric": "MLUPS (fp64 D3Q15, Rohde cycle)", "value": work * args.steps / (ms * 1e-3) / 1e6,
                          "n_gpus": world, "ms_per_coarse_step": ms / args.steps, "levels": args.levels,
                          "coupling": args.coupling, "gradient_threshold": args.gradient, "cells_per_level_at_end": cells_end,
                          "valid_tiling": args.valid_tiling, "debug_skip": args.debug_skip,
                          "fused": not args.no_fusion, "max_grid": args.max_grid, "base_grid": [n, n, n],
                          "cells_per_level": cells, "boxes_per_level": nbox, "boxes_of_rank0": mine, "substeps": substeps,
                          "launches_per_coarse_step": launches / args.steps, "regrid_seconds": regrid_s,
                          "regrids_in_timed_region": regrids, "regrid_host_seconds_in_timed_region": regrid_in_loop_s,
                          "first_%d_steps_seconds" % args.warmup: first_s,
                          "device_barriers": lbx.par_info()["barriers"],
                          "check_mean_rho_level0": float(rho.mean()),
                          "bytes_per_cell_update_at_roofline": 240}), flush=True)
    sim.close()
    amrsim.lambrexFinalise()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
