"""AMR-path throughput (BASELINE.json configs[3] and [4]) from the command line: multi-level pulse, refinement
ratio 2, through AmrSim.  One GPU, or -- under torchrun -- the boxes of every level distributed over the ranks.
Prints one JSON line from rank 0 (lambrex_b200/amr_workload.py does the work; bench.py's AMR leg calls the same
function)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import amrsim, lbx   # noqa: E402
from lambrex_b200.amr_workload import run_amr_case   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--levels", type=int, default=2)
    ap.add_argument("--max-grid", type=int, default=32)
    ap.add_argument("--regrid-every", type=int, default=0, help="coarse steps between regrids (0: never)")
    ap.add_argument("--no-fusion", action="store_true", help="Rohde cycle as the literal pass sequence")
    ap.add_argument("--debug-skip", type=int, default=0, help="profiling only: 1 skip valid tiles, 2 skip ghost tiles")
    ap.add_argument("--coupling", default="rohde", choices=["rohde", "subcycle"])
    ap.add_argument("--gradient", type=float, default=0.0,
                    help="> 0: refine level 0 by the density-gradient criterion (device tagging); regrids inside Iterate")
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
    lbx.set_option(lbx.OPT_DEBUG_SKIP, args.debug_skip)
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except Exception:
        peak = 6650.0
    res = run_amr_case(args.grid, args.levels, args.steps, args.warmup, args.coupling, args.regrid_every, args.max_grid,
                       not args.no_fusion, args.gradient, dist, peak)
    if rank == 0:
        print(json.dumps(res), flush=True)
    amrsim.lambrexFinalise()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
