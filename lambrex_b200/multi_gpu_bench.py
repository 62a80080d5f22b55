"""bench.py --gpus N (N > 1), one rank per GPU under torchrun: BASELINE.json configs[2], the uniform 1024^3
periodic shear wave, run through the reference-facing class on EVERY rank -- AmrSim (liblambrex.so) after
lambrexInitParallel.  Level 0 is owned as one z-slab per rank (amrex::DistributionMapping by whole x-y layers
of boxes, the reference's only parallelism: box ownership, /root/reference/src/AmrSim.cpp:665-669) and stored
as one ghost-free fab per GPU; a time step is ONE launch per GPU (lbx_mf_collide_stream_slab): boundary-plane
CTAs store the 5 face-crossing populations straight into the neighbour's fab over NVLink (CUDA-IPC peer
pointers) and order themselves against the neighbours' steps inside the kernel.  Rank 0 prints the one JSON
line.  --halo nccl keeps the round-1 baseline transport (Python SlabSim, packed planes over NCCL send/recv)."""
import ctypes
import json
import os
import time

import numpy as np


def run_multi(args, helpers):
    import torch
    import torch.distributed as dist
    from . import amrsim, lbx

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit("bench.py --gpus %d must be launched with torchrun --nproc-per-node %d "
                         "(WORLD_SIZE is %d)" % (args.gpus, args.gpus, world))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    if args.halo != "p2p":
        return run_multi_slabsim(args, helpers, rank, world, local)
    amrsim.lambrexInitParallel()

    n = args.grid_multi
    tau, U = 0.1, 0.01
    # memory guard per local cell: 2 x 15 population planes, rho (now, next) + u, 2 % slack (+ 128 MiB of staging)
    info = lbx.device_info()
    need = lambda edge: (edge * edge * -(-edge // world)) * (240 + 40) * 1.02 + (160 << 20)
    fits = torch.tensor([1 if need(n) < info["free_bytes"] else 0])
    dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    reduced = False
    while not fits.item() and n > 128:
        n //= 2
        reduced = True
        fits = torch.tensor([1 if need(n) < info["free_bytes"] else 0])
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    nx = ny = nz = n
    if args.grid_multi_z:
        nz = args.grid_multi_z
    cells_total = float(nx) * ny * nz
    prof = U * np.sin(2.0 * np.pi * np.arange(ny, dtype=np.float64) / ny)
    u_profile = np.zeros((ny, 3))
    u_profile[:, 0] = prof

    def make_sim():
        sim = amrsim.AmrSim(nx, ny, nz, 0, (1, 1, 1), tau, tau)
        sim.SetMaxGridSize(args.max_grid_multi)
        return sim

    # ---- device-resident leg: separable initial condition (no whole-domain host array exists at 1024^3) -------
    sim = make_sim()
    lo, hi = sim.LocalBox()
    nzl = hi[2] - lo[2] + 1
    sim.SetInitialDensityProfile(2, np.ones(nz))
    sim.SetInitialVelocityProfile(1, u_profile)
    sim.InitFromScratch(0.0)
    sim.Iterate(args.warmup)
    lbx.sync()
    dist.barrier()
    sampler = helpers["ClockSampler"](local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    l0 = lbx.launch_count()
    with lbx.Timer() as t:
        sim.Iterate(args.steps)           # K launches of the fused step + the final neighbour wait
    launches = lbx.launch_count() - l0
    lbx.sync()
    dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([t.ms], dtype=torch.float64)
    per_rank_ms = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(per_rank_ms, tt)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    ln = torch.tensor([launches], dtype=torch.int64)
    dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    ms_step = ms_total / args.steps
    mlups = cells_total * args.steps / (ms_total * 1e-3) / 1e6
    sim.close()

    # ---- end-to-end leg: every rank states ITS slab in pinned host arrays (C order [i][j][k_local]) -> H2D ->
    # equilibrium -> K steps -> moments -> D2H of its slab's rho, u; also the source of the physics checks
    L = lbx.lib()
    ncl = nx * ny * nzl
    hp = ctypes.c_void_p()
    lbx.check(L.lbx_host_alloc(ctypes.byref(hp), 4 * ncl * 8))
    host = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(4 * ncl,))
    rho_h, u_h = host[:ncl].reshape(nx, ny, nzl), host[ncl:].reshape(nx, ny, nzl, 3)
    rho_h[...] = 1.0
    u_h[...] = 0.0
    u_h[..., 0] = prof[None, :, None]
    lbx.sync()
    dist.barrier()
    t0 = time.perf_counter()
    sim = make_sim()
    sim.SetInitialDensityLocalView(host[:ncl])
    sim.SetInitialVelocityLocalView(host[ncl:])
    sim.InitFromScratch(0.0)
    sim.Iterate(args.steps)
    sim.CalcHydroVars(0)
    sim.GetLocalDensityField(0, host[:ncl])
    sim.GetLocalVelocityField(0, host[ncl:])
    e2e_s = time.perf_counter() - t0
    dist.barrier()
    te = torch.tensor([e2e_s], dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    io_bytes = 32.0 * cells_total
    e2e = {"value": cells_total * args.steps / float(te.item()) / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": io_bytes / args.steps, "d2h_bytes_per_step": io_bytes / args.steps,
           "note": "one job on every rank = AmrSim ctor + SetInitialDensity/VelocityLocalView (this rank's slab, pinned host "
                   "arrays) + InitFromScratch (H2D, equilibrium) + Iterate(%d) + CalcHydroVars + GetLocalDensity/VelocityField "
                   "(D2H); max over ranks" % args.steps}
    # physics checks at full size from the returned host arrays: mass conservation and the shear-wave decay
    mass = float(rho_h.sum())
    amp = float((u_h[..., 0].mean(axis=(0, 2)) * np.sin(2.0 * np.pi * np.arange(ny) / ny)).sum() * 2.0 / ny * nzl)
    red = torch.tensor([mass, amp], dtype=torch.float64)
    dist.all_reduce(red, op=dist.ReduceOp.SUM)
    kk = 2.0 * np.pi / ny
    check = {"total_mass_over_cells": float(red[0]) / cells_total,
             "ux_amplitude_over_U": float(red[1]) / nz / U,
             "ux_amplitude_expected": float(np.exp(-(tau / 3.0) * kk * kk * args.steps)),
             "source": "this run's end-to-end job (%d steps), host arrays returned by GetLocal*Field" % args.steps}
    boxes0 = len(sim.boxArray(0))
    sim.close()
    del rho_h, u_h, host
    lbx.check(L.lbx_host_free(hp))

    # ---- AMR leg: BASELINE configs[4] (C5) -- 3-level pulse, boxes of every level distributed over the ranks,
    # static nested boxes translated every 16 coarse steps (regrid); static run beside it for the regrid cost
    amr = None
    if not args.no_amr:
        from .amr_workload import run_amr_case
        lbx.check(L.lbx_arena_release())          # the uniform legs' blocks go back to the driver first
        dist.barrier()
        pk = helpers["measured_peak"]()[0]
        amr = {"config": "C5: 3-level AMR pulse, %d^3 base, nested central-half boxes translated every 16 coarse steps, 32^3 boxes "
                         "distributed over %d B200 (BASELINE configs[4]%s)"
                         % (args.amr_grid, world, "" if args.amr_grid == 512 else "; base REDUCED from 512^3 for %d GPUs" % world)}
        amr["C5_regrid16"] = run_amr_case(args.amr_grid, 3, args.amr_steps, 3, "rohde", 16, 32, True, 0.0, dist, pk)
        amr["C5_static"] = run_amr_case(args.amr_grid, 3, max(8, args.amr_steps // 2), 3, "rohde", 0, 32, True, 0.0, dist, pk)
        amr["C5_subcycle_regrid16"] = run_amr_case(args.amr_grid, 3, args.amr_steps, 3, "subcycle", 16, 32, True, 0.0, dist, pk)
        # not a headline: max_grid_size 64 instead of AMReX's default 32
        amr["C5_static_boxes64"] = run_amr_case(args.amr_grid, 3, max(8, args.amr_steps // 2), 3, "rohde", 0, 64, True, 0.0, dist, pk)

    if rank == 0:
        peak, peak_src = helpers["measured_peak"]()
        achieved = helpers["BYTES_PER_CELL"] * float(nx) * ny * nzl / (ms_step * 1e-3) / 1e9
        cfg = helpers["workload_config"](world)
        cfg.update({"grid": [nx, ny, nz], "halo": "p2p stores fused in the step kernel (one launch per step and GPU)",
                    "api": "AmrSim (liblambrex.so) on every rank after lambrexInitParallel -> lbx_mf_collide_stream_slab",
                    "partition": "box ownership by whole x-y layers: one z-slab of %d planes per rank, %d level-0 boxes of %d^3"
                                 % (nzl, boxes0, args.max_grid_multi),
                    "parallelism": "domain decomposition, %d slabs" % world})
        if reduced:
            cfg["workload"] += " -- REDUCED to %d^3: 1024^3 does not fit %d GPUs' free HBM" % (n, world)
        line = {"metric": helpers["METRIC"], "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(ln.item()),
                "roofline": {"bound": "hbm", "kernel": "k_collide_stream_slab_sync (rank 0's slab; wait, face stores over NVLink and "
                                                       "signal inside the kernel)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "peak_source": peak_src, "traffic": None,
                             "per_rank_ms_per_step": [round(float(x.item()) / args.steps, 4) for x in per_rank_ms]},
                "cpu_baseline": None, "check": check}
        if amr is not None:
            line["extra"] = {"amr": amr}
        print(json.dumps(line), flush=True)
    dist.barrier()
    amrsim.lambrexFinalise()
    dist.destroy_process_group()


def run_multi_slabsim(args, helpers, rank, world, local):
    """--halo nccl: the round-1 baseline transport, kept for comparison -- Python SlabSim (lambrex_b200/slab.py),
    ghost planes exchanged as packed 5-population planes with NCCL send/recv."""
    import torch
    import torch.distributed as dist
    from . import lbx
    from .slab import SlabSim

    lbx.init(local)
    n = args.grid_multi
    tau, U = 0.1, 0.01
    nx = ny = nz = n
    sim = SlabSim(nx, ny, nz, tau, tau, rank=rank, world=world, halo="nccl" if args.halo == "nccl" else "p2p")
    cells_total = float(nx) * ny * nz
    prof = U * torch.sin(2.0 * np.pi * torch.arange(ny, dtype=torch.float64) / ny)
    sim.rho_t.fill_(1.0)
    sim.u_t.zero_()
    sim.u_t[0] += prof.to(sim.dev)[None, :, None]
    torch.cuda.synchronize()
    sim.set_initial(None, None)
    sim.step(args.warmup)
    sim.barrier()
    torch.cuda.synchronize()
    dist.barrier()
    l0 = lbx.launch_count()
    with lbx.Timer() as t:
        sim.step(args.steps)
        sim.finish()
    launches = lbx.launch_count() - l0
    torch.cuda.synchronize()
    dist.barrier()
    tt = torch.tensor([t.ms], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    if rank == 0:
        cfg = helpers["workload_config"](world)
        cfg.update({"grid": [nx, ny, nz], "halo": "nccl send/recv of packed 5-population planes (baseline transport)" if args.halo == "nccl"
                    else "round-1 path: peer stores, separate wait / boundary / signal / interior launches",
                    "api": "SlabSim (Python), not AmrSim"})
        print(json.dumps({"metric": helpers["METRIC"], "value": cells_total * args.steps / (ms_total * 1e-3) / 1e6, "unit": "MLUPS",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": cfg, "gpu_launches": int(launches), "e2e": None, "roofline": None, "cpu_baseline": None}), flush=True)
    sim.close()
    dist.barrier()
    dist.destroy_process_group()
    lbx.finalize()
