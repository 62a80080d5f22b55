"""torchrun entry (>= 2 GPUs): the DISTRIBUTED AMR path -- boxes of every level owned by ranks,
neighbour boxes read over NVLink through CUDA-IPC, device-side all-rank barriers -- must reproduce
the single-GPU AmrSim BIT FOR BIT: every cell (ghost rings included) of every box of every level,
plus the clocks.  Each rank also runs the whole problem alone (all boxes local) as the reference.
Prints AMR_DIST_CHECK_OK from rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import amrsim, lbx, workloads   # noqa: E402

PER = (1, 1, 1)


def build(nx, ny, nz, max_level, boxes, max_grid, fast=False):
    rho, u = workloads.shear_wave(nx, ny, nz)
    rho = rho * workloads.pulse_density(nx, ny, nz)
    sim = amrsim.AmrSim(nx, ny, nz, max_level, PER, 0.3, 0.4)
    # fast: level 0 alone is stored as ONE ghost-free slab per rank and stepped by the fused kernel with peer
    # stores (the uniform multi-GPU path inside AmrSim); else per-box storage with ghost cells on every path
    sim.SetUniformFastPath(fast)
    if "--subcycle" in sys.argv:         # conventional subcycling + gradient tagging + regrid_int
        sim.SetCoupling(amrsim.SUBCYCLE)
    sim.SetMaxGridSize(max_grid)
    sim.SetInitialDensity(rho)
    sim.SetInitialVelocity(u)
    sim.InitFromScratch(0.0)
    for lev, (lo, hi) in enumerate(boxes):
        sim.SetStaticRefinement(lev, lo, hi)
    if "--subcycle" in sys.argv and boxes:
        sim.SetGradientRefinement(0, 2e-3)
        sim.SetRegridInterval(2)
    return sim


def snapshot(sim, max_level):
    out = []
    for lev in range(max_level + 1):
        boxes = sim.FieldBoxes(lev, amrsim.DISTFN)
        out.append((boxes, [sim.FieldFab(lev, amrsim.DISTFN, b, 2, 15) for b in range(len(boxes))],
                    sim.GetTime(lev), sim.GetTimeStep(lev)))
    rho = [sim.GetDensityField(lev) for lev in range(max_level + 1)]
    return out, rho


def local_io_check(rank, world):
    """Distributed uniform run stated per rank: profile initial conditions and local (slab) arrays in, local
    fields out -- against the whole-domain API on a single GPU, bit for bit."""
    nx, ny, nz, steps = 24, 20, 8 * max(world, 2), 9
    rho3, u3 = workloads.shear_wave(nx, ny, nz)             # rho = 1, u_x(j)
    rho3 = (rho3 * workloads.pulse_density(nx, ny, nz)).reshape(nx, ny, nz)       # rho(k)
    u3 = u3.reshape(nx, ny, nz, 3)
    good = True
    for mode in ("profile", "local arrays"):
        sim = amrsim.AmrSim(nx, ny, nz, 0, PER, 0.3, 0.4)
        sim.SetMaxGridSize(8)
        lo, hi = sim.LocalBox()
        if mode == "profile":
            sim.SetInitialDensityProfile(2, rho3[0, 0, :])
            # a profile states ONE axis: u(y) of the shear wave; the density then carries no y dependence
            sim.SetInitialVelocityProfile(1, u3[0, :, 0, :])
        else:
            sl = (slice(lo[0], hi[0] + 1), slice(lo[1], hi[1] + 1), slice(lo[2], hi[2] + 1))
            rl, ul = np.ascontiguousarray(rho3[sl]), np.ascontiguousarray(u3[sl])
            sim.SetInitialDensityLocalView(rl)
            sim.SetInitialVelocityLocalView(ul)
        sim.InitFromScratch(0.0)
        sim.Iterate(steps)
        sim.CalcHydroVars(0)
        r_loc, u_loc = sim.GetLocalDensityField(0), sim.GetLocalVelocityField(0)
        sim.close()
        dist.barrier()
        amrsim.setParallelView(0, 1)
        ref = amrsim.AmrSim(nx, ny, nz, 0, PER, 0.3, 0.4)
        ref.SetMaxGridSize(8)
        ref.SetInitialDensity(rho3.reshape(-1))
        ref.SetInitialVelocity(u3.reshape(-1))
        ref.InitFromScratch(0.0)
        ref.Iterate(steps)
        ref.CalcHydroVars(0)
        r_ref, u_ref = ref.GetDensityField(0), ref.GetVelocityField(0)
        ref.close()
        amrsim.setParallelView(rank, world)
        sl = (slice(lo[0], hi[0] + 1), slice(lo[1], hi[1] + 1), slice(lo[2], hi[2] + 1))
        same = np.array_equal(r_loc, r_ref[sl]) and np.array_equal(u_loc, u_ref[sl])
        flag = torch.tensor([1 if same else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("amr_dist_check local i/o (%s) world=%d slab of rank 0 %s..%s bit-equal=%s" % (mode, world, lo, hi, bool(flag.item())),
                  flush=True)
        good = good and bool(flag.item())
    return good


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # more ranks than GPUs (debugging the host logic of an 8-rank layout on one GPU): the ranks share devices -- CUDA-IPC
    # and the device-side barriers work across processes of one GPU (time-sliced, slow), NCCL does not
    shared = torch.cuda.device_count() < world
    torch.cuda.set_device(local % torch.cuda.device_count())
    dist.init_process_group("gloo" if shared else "cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    amrsim.lambrexInitParallel()
    ok = True
    cases = [
        ("1 level", (16, 12, 20), 0, [], 8, 3),
        ("2 levels", (16, 12, 20), 1, [((3, 2, 4), (11, 9, 14))], 8, 3),
        ("3 levels", (16, 16, 16), 2, [((3, 3, 3), (12, 12, 12)), ((10, 10, 10), (21, 21, 21))], 8, 2),
    ]
    if "--only-2level" in sys.argv:
        cases = cases[1:2]
    if "--big" in sys.argv:
        cases = [("64^3 L2", (64, 64, 64), 1, [((16, 16, 16), (47, 47, 47))], 16, 3),
                 ("128^3 L2", (128, 128, 128), 1, [((32, 32, 32), (95, 95, 95))], 32, 3),
                 ("128^3 L3", (128, 128, 128), 2, [((32, 32, 32), (95, 95, 95)), ((96, 96, 96), (159, 159, 159))], 32, 2)]
    cases = [c + (False,) for c in cases] + [(c[0] + " (slab level 0)",) + c[1:] + (True,) for c in cases]
    for name, (nx, ny, nz), max_level, boxes, max_grid, steps, fast in cases:
        sim = build(nx, ny, nz, max_level, boxes, max_grid, fast)
        owners = [sorted({sim.Owner(lev, b) for b in range(len(sim.boxArray(lev)))}) for lev in range(max_level + 1)]
        sim.Iterate(steps)
        for lev in range(max_level + 1):
            sim.CalcHydroVars(lev)
        got, rho_got = snapshot(sim, max_level)
        # regrid mid-run: move the finest static box, keep stepping
        if boxes:
            rlev = len(boxes) - 1
            lo, hi = boxes[rlev]
            sim.SetStaticRefinement(rlev, tuple(v + 1 for v in lo), tuple(v + 1 for v in hi))
            sim.Iterate(1)
            got2, _ = snapshot(sim, max_level)
        sim.close()
        dist.barrier()

        amrsim.setParallelView(0, 1)            # the same problem, alone on this GPU
        ref = build(nx, ny, nz, max_level, boxes, max_grid, fast)
        ref.Iterate(steps)
        for lev in range(max_level + 1):
            ref.CalcHydroVars(lev)
        want, rho_want = snapshot(ref, max_level)
        if boxes:
            ref.SetStaticRefinement(rlev, tuple(v + 1 for v in lo), tuple(v + 1 for v in hi))
            ref.Iterate(1)
            want2, _ = snapshot(ref, max_level)
        ref.close()
        amrsim.setParallelView(rank, world)

        def same(a, b):
            good = True
            for lv, ((ba, fa, ta, sa), (bb, fb, tb, sb)) in enumerate(zip(a, b)):
                if not (ba == bb and ta == tb and sa == sb and len(fa) == len(fb)):
                    firstdiff = [(x, y) for x, y in zip(ba, bb) if x != y][:2]
                    print("[rank %d] %s level %d: metadata differs (%d vs %d boxes, t %s/%s, step %s/%s) first %s"
                          % (rank, name, lv, len(ba), len(bb), ta, tb, sa, sb, firstdiff), flush=True)
                    good = False
                    continue
                for bi, (x, y) in enumerate(zip(fa, fb)):
                    if fast and lv == 0 and max_level == 0:
                        # a ghost-free level has no ghost cells to compare: what FieldFab shows around the valid box
                        # depends on the storage (one slab per rank when the layers divide over the ranks, boxes
                        # otherwise -- 3 layers on 8 ranks), the valid cells do not
                        x, y = x[:, 2:-2, 2:-2, 2:-2], y[:, 2:-2, 2:-2, 2:-2]
                        if np.array_equal(x, y):
                            continue
                        print("[rank %d] %s level 0 box %d: valid cells differ" % (rank, name, bi), flush=True)
                        good = False
                        continue
                    if not np.array_equal(x, y):
                        d = np.abs(x - y)
                        dv = float(np.max(d[:, 2:-2, 2:-2, 2:-2]))
                        if good:
                            print("[rank %d] %s level %d box %d %s: max|diff| all %.3e valid %.3e, %d cells differ, nan got %d"
                                  % (rank, name, lv, bi, ba[bi], float(np.nanmax(d)), dv, int(np.count_nonzero(d)),
                                     int(np.isnan(x).sum())), flush=True)
                        good = False
            return good

        good = same(got, want) and all(np.array_equal(x, y) for x, y in zip(rho_got, rho_want))
        if boxes:
            good = good and same(got2, want2)
        flag = torch.tensor([1 if good else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("amr_dist_check %-8s world=%d owners per level=%s bit-equal=%s" % (name, world, owners, bool(flag.item())),
                  flush=True)
        ok = ok and bool(flag.item())
        # every rank must have gone through the same number of device barriers: a rank that skips or adds one pairs
        # its later barriers with the wrong ones of its peers (no deadlock, silently unordered reads)
        bc = torch.zeros(world, dtype=torch.int64)
        bc[rank] = lbx.par_info()["barriers"]
        dist.all_reduce(bc, op=dist.ReduceOp.SUM)
        if int(bc.min()) != int(bc.max()):
            if rank == 0:
                print("amr_dist_check %-8s BARRIER COUNTS DIFFER across ranks: %s" % (name, bc.tolist()), flush=True)
            ok = False
    ok = local_io_check(rank, world) and ok
    info = lbx.par_info()
    if rank == 0:
        print("device barriers executed: %d" % info["barriers"], flush=True)
        if ok:
            print("AMR_DIST_CHECK_OK", flush=True)
    amrsim.lambrexFinalise()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
