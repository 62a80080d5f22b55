"""AMR-path throughput (BASELINE.json configs[3]): 2-level pulse, refinement ratio 2, the
reference's Rohde cycle with subcycling, on one GPU through AmrSim.  Prints one JSON line:
MLUPS counts sum_l cells_l x substeps_l per coarse step."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import amrsim, lbx, workloads   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--levels", type=int, default=2)
    ap.add_argument("--max-grid", type=int, default=32)
    ap.add_argument("--no-fusion", action="store_true", help="Rohde cycle as the literal pass sequence")
    args = ap.parse_args()
    n = args.grid
    amrsim.lambrexInit()
    sim = amrsim.AmrSim(n, n, n, args.levels - 1, (1, 1, 1), 0.5, 0.5)
    sim.SetMaxGridSize(args.max_grid)
    sim.SetRohdeFusion(not args.no_fusion)
    sim.SetInitialDensity(workloads.pulse_density(n, n, n))
    sim.SetInitialVelocity(0.0)
    sim.InitFromScratch(0.0)
    t0 = time.perf_counter()
    lo, hi = n // 4, 3 * n // 4 - 1
    for lev in range(args.levels - 1):
        sim.SetStaticRefinement(lev, (lo,) * 3, (hi,) * 3)
        lo, hi = 2 * lo + (hi - lo + 1) // 2, 2 * lo + (hi - lo + 1) // 2 + (hi - lo)      # central half again
    regrid_s = time.perf_counter() - t0
    cells = [sum(int(np.prod([h - l + 1 for l, h in zip(*b)])) for b in sim.boxArray(l)) for l in range(args.levels)]
    nbox = [len(sim.boxArray(l)) for l in range(args.levels)]
    substeps = [1] + [2 ** l for l in range(1, args.levels)]
    t0 = time.perf_counter()
    sim.Iterate(args.warmup)        # includes the FLAT -> BOXES re-layout and plan building
    lbx.sync()
    first_s = time.perf_counter() - t0
    l0 = lbx.launch_count()
    with lbx.Timer() as t:
        sim.Iterate(args.steps)
    launches = lbx.launch_count() - l0
    work = sum(c * s for c, s in zip(cells, substeps))
    print(json.dumps({"metric": "MLUPS (fp64 D3Q15, Rohde cycle)", "value": work * args.steps / (t.ms * 1e-3) / 1e6,
                      "ms_per_coarse_step": t.ms / args.steps, "levels": args.levels, "fused": not args.no_fusion, "max_grid": args.max_grid, "base_grid": [n, n, n],
                      "cells_per_level": cells, "boxes_per_level": nbox, "substeps": substeps,
                      "launches_per_coarse_step": launches / args.steps, "regrid_seconds": regrid_s,
                      "first_%d_steps_seconds" % args.warmup: first_s,
                      "bytes_per_cell_update_at_roofline": 240}), flush=True)


if __name__ == "__main__":
    main()
