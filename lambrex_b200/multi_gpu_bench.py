"""bench.py --gpus N (N > 1), one rank per GPU under torchrun: BASELINE.json configs[2],
the uniform 1024^3 periodic shear wave sharded as z-slabs with the face exchange fused into
the step kernel (peer stores over NVLink) or, with --halo nccl, packed 5-population planes
over NCCL send/recv.  Rank 0 prints the one JSON line."""
import json
import os
import time

import numpy as np


def run_multi(args, helpers):
    import torch
    import torch.distributed as dist
    from . import lbx
    from .slab import SlabSim

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit("bench.py --gpus %d must be launched with torchrun --nproc-per-node %d "
                         "(WORLD_SIZE is %d)" % (args.gpus, args.gpus, world))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    lbx.init(local)

    n = args.grid_multi
    tau, U = 0.1, 0.01
    # memory guard: 2 x 15 population planes + rho,u per local cell
    info = lbx.device_info()
    need = lambda edge: (edge * edge * -(-edge // world)) * (240 + 32 + 8) * 1.02
    fits = torch.tensor([1 if need(n) < info["free_bytes"] else 0])
    dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    reduced = False
    while not fits.item() and n > 128:
        n //= 2
        reduced = True
        fits = torch.tensor([1 if need(n) < info["free_bytes"] else 0])
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    nx = ny = nz = n
    sim = SlabSim(nx, ny, nz, tau, tau, rank=rank, world=world, halo=args.halo, split=not args.no_split)
    klo, khi = sim.layout.slab(rank)
    nzl = khi - klo + 1
    cells_total = float(nx) * ny * nz
    prof = U * torch.sin(2.0 * np.pi * torch.arange(ny, dtype=torch.float64) / ny)

    def device_init():
        sim.rho_t.fill_(1.0)
        sim.u_t.zero_()
        sim.u_t[0] += prof.to(sim.dev)[None, :, None]
        torch.cuda.synchronize()
        sim.set_initial(None, None)

    # ---- device-resident leg ---------------------------------------------------------
    device_init()
    sim.step(args.warmup)
    sim.barrier()
    torch.cuda.synchronize()
    sampler = helpers["ClockSampler"](local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    l0 = lbx.launch_count()
    with lbx.Timer() as t:
        sim.step(args.steps)
        sim.finish()
    launches = lbx.launch_count() - l0
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([t.ms], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    ln = torch.tensor([launches], dtype=torch.int64)
    dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    ms_step = ms_total / args.steps
    mlups = cells_total * args.steps / (ms_total * 1e-3) / 1e6

    # physics checks at full size: mass conservation and the shear-wave decay rate
    sim.moments(None, None)
    sim.sync()
    torch.cuda.synchronize()
    mass = sim.rho_t.sum().cpu()
    amp = (sim.u_t[0].mean(dim=(0, 2)).cpu() * torch.sin(2.0 * np.pi * torch.arange(ny, dtype=torch.float64) / ny)
           ).sum() * 2.0 / ny * nzl
    red = torch.stack([mass, amp])
    dist.all_reduce(red, op=dist.ReduceOp.SUM)
    tsteps = args.warmup + args.steps
    kk = 2.0 * np.pi / ny
    check = {"total_mass_over_cells": float(red[0]) / cells_total,
             "ux_amplitude_over_U": float(red[1]) / nz / U,
             "ux_amplitude_expected": float(np.exp(-(tau / 3.0) * kk * kk * tsteps))}

    # ---- end-to-end leg: pinned host rho,u -> H2D -> equilibrium -> K steps -> moments -> D2H
    rho_h = torch.empty((1, nzl, ny, nx), dtype=torch.float64, pin_memory=True)
    u_h = torch.empty((3, nzl, ny, nx), dtype=torch.float64, pin_memory=True)
    rho_h.fill_(1.0)
    u_h.zero_()
    u_h[0] += prof[None, :, None]
    sim.barrier()
    t0 = time.perf_counter()
    sim.set_initial(rho_h, u_h)
    sim.step(args.steps)
    sim.moments(rho_h, u_h)
    sim.sync()
    e2e_s = time.perf_counter() - t0
    dist.barrier()
    te = torch.tensor([e2e_s], dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    io_bytes = 32.0 * cells_total
    e2e = {"value": cells_total * args.steps / float(te.item()) / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": io_bytes / args.steps, "d2h_bytes_per_step": io_bytes / args.steps,
           "note": "one job = H2D rho,u slabs (pinned) + equilibrium + %d steps + moments + D2H rho,u, all ranks"
                   % args.steps}

    if rank == 0:
        peak, peak_src = helpers["measured_peak"]()
        cells_local = float(nx) * ny * max(h - l + 1 for l, h in sim.layout.slabs)
        achieved = helpers["BYTES_PER_CELL"] * cells_local / (ms_step * 1e-3) / 1e9
        tr = helpers["recorded_traffic"]()
        cfg = helpers["workload_config"](world)
        cfg.update({"grid": [nx, ny, nz], "halo": args.halo, "split": bool(sim.split and args.halo == "p2p"), "partition": "z-slabs %s" % sim.layout.slabs,
                    "parallelism": "domain decomposition, %d slabs" % world})
        if reduced:
            cfg["workload"] += " -- REDUCED to %d^3: 1024^3 does not fit %d GPUs' free HBM" % (n, world)
        line = {"metric": helpers["METRIC"], "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(ln.item()),
                "roofline": {"bound": "hbm", "kernel": ("k_collide_stream<push> on the interior planes + k_collide_stream_slab on the 2 boundary "
                                        "planes (per GPU, largest slab)") if sim.split and args.halo == "p2p" else
                             "k_collide_stream_slab (per GPU, largest slab)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "peak_source": peak_src, "traffic": (tr or {}).get("bytes_per_launch"),
                             "traffic_source": (tr or {}).get("source")},
                "cpu_baseline": None, "check": check}
        print(json.dumps(line), flush=True)
    sim.close()
    dist.barrier()
    dist.destroy_process_group()
    lbx.finalize()
