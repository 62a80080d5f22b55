#!/usr/bin/env python3
"""Benchmark of the D3Q15 fp64 collide-and-stream hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one lattice time step (collide + stream) of the whole grid.
  N = 1 : BASELINE.json configs[1] -- uniform single-level periodic pulse, 256^3, fp64.
  N > 1 : configs[2] -- uniform 1024^3 periodic shear wave through AmrSim on every rank (one z-slab of
          level-0 boxes per GPU, face exchange fused into the step kernel; launched under torchrun).
value   = MLUPS with the populations resident in HBM (device-timed, CUDA events).
e2e     = MLUPS through the host-buffer API: pinned-host rho,u -> H2D -> equilibrium
          -> K steps -> moments -> D2H of rho,u, all inside the timed region.
roofline= fused kernel against the measured HBM copy bandwidth, 240 B/cell/step.
cpu_baseline / --impl reference = the oracle's restatement of the reference's own
          pass structure (ghosted 32^3 boxes, FillPatch copy, in-place dense collide,
          FillBoundary, stream into a fresh fab, swap) timed on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MLUPS (fp64 D3Q15 BGK)"
BYTES_PER_CELL = 240.0       # 15 fp64 loads + 15 fp64 stores


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as fh:
            return json.load(fh)
    except Exception:
        return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax = float(c[2])
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
def base_grid_edges(n, max_grid=32):
    """AMReX MakeBaseGrids chop of one direction -- the
    max_grid_size-32 boxes the reference runs on."""
    from lambrex_b200.boxes import chop_1d
    return chop_1d(n, max_grid)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference(sample_n, steps, warmup, tau=0.5, threads=None, loop_order=0, shear=False):
    """Time the oracle's restatement of the reference's pass structure (CPU).  threads=1 with loop_order=1
    (k innermost, /root/reference/include/amr_help.h:87-92) is the FAITHFUL figure: the reference has no OpenMP
    and its AMReX recipe disables it (amrex_cmake.sh:9-10); all cores with x innermost is the generous one."""
    from lambrex_b200 import workloads
    from oracle import lbm_oracle as orc
    co = orc.COracle()
    all_threads = co.max_threads()
    co.set_threads(threads or all_threads)
    n = sample_n
    w = workloads.omega(tau)
    if shear:
        rho = np.ones((n, n, n))
        u = np.zeros((3, n, n, n))
        u[0] = (0.01 * np.sin(2.0 * np.pi * np.arange(n) / n))[None, :, None]
    else:
        rho = orc.user_to_fab(workloads.pulse_density(n, n, n), n, n, n)
        u = np.zeros((3, n, n, n))
    f = co.equilibrium(rho, u)
    edges = [base_grid_edges(n)] * 3
    if warmup:
        f, _ = co.ref_passes(f, w, w, warmup, edges, loop_order)
    f, secs = co.ref_passes(f, w, w, steps, edges, loop_order)
    co.set_threads(all_threads)
    cells = float(n) ** 3
    return {"mlups": cells * steps / secs / 1e6, "secs": secs, "cores": threads or all_threads,
            "sample": "%d^3 periodic %s, %d steps after %d warm-up, 32^3 ghosted boxes, reference pass structure "
                      "(FillPatch copy, in-place collide, FillBoundary, stream into a fresh fab, swap), %s loop order, "
                      "gcc -O3 -fopenmp, %d thread%s" % (n, "shear wave" if shear else "pulse", steps, warmup,
                                                        "k-innermost (the reference's)" if loop_order else "x-innermost",
                                                        threads or all_threads, "" if (threads or all_threads) == 1 else "s")}


def cpu_baseline_block(sample_n):
    """cpu_baseline of the GPU arm's line: the all-core figure (value) and the faithful 1-thread figure."""
    r = cpu_reference(sample_n, 4, 1)
    one = cpu_reference(min(sample_n, 64), 2, 1, threads=1, loop_order=1)
    return {"value": r["mlups"], "unit": "MLUPS", "cores": r["cores"], "kind": "port", "sample": r["sample"],
            "cpu_model": cpu_model(),
            "faithful_1thread": {"value": one["mlups"], "unit": "MLUPS", "cores": 1, "sample": one["sample"],
                                 "why": "the reference contains no OpenMP and its AMReX recipe sets ENABLE_OMP=OFF "
                                        "(amrex_cmake.sh:9-10); loop order of include/amr_help.h:87-92"}}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (kind "port": the oracle's C
    restatement of its pass structure; the real reference needs AMReX + gfortran and cannot be built here) on all
    host cores.  N = 1: the labelled configuration itself (256^3 pulse), a few steps.  N > 1: 1024^3 does not fit a
    host, so each step is a bounded 256^3 sample of the shear-wave workload -- the label says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                      # under torchrun only rank 0 measures the CPU arm
    steps = max(1, min(args.steps, 4))
    warm = max(0, min(args.warmup, 1))
    n = args.cpu_ref_grid
    shear = args.gpus > 1
    try:
        r = cpu_reference(n, steps, warm, tau=0.1 if shear else 0.5, shear=shear)
    except MemoryError:
        n = 128
        r = cpu_reference(n, steps, warm, tau=0.1 if shear else 0.5, shear=shear)
    cfg = workload_config(args.gpus)
    if shear:
        cfg["workload"] += " -- CPU arm: bounded %d^3 sample of it per step (1024^3 does not fit a host)" % n
    elif n != 256:
        cfg["workload"] += " -- CPU arm: bounded %d^3 sample of it" % n
    cfg["grid"] = [n, n, n]
    one = cpu_reference(64, 2, 1, threads=1, loop_order=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["mlups"], "unit": "MLUPS",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": r["secs"] / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["mlups"], "unit": "MLUPS", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"], "cpu_model": cpu_model(),
                         "faithful_1thread": {"value": one["mlups"], "unit": "MLUPS", "cores": 1, "sample": one["sample"]}},
        "e2e": {"value": r["mlups"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def workload_config(ngpus):
    if ngpus == 1:
        return {"workload": "uniform single-level periodic pulse 256^3 fp64 (BASELINE configs[1])",
                "grid": [256, 256, 256], "tau": 0.5, "l2": "inputs (2 x 2.0 GB) larger than L2"}
    return {"workload": "uniform 1024^3 periodic shear wave, z-slabs over %d GPUs (BASELINE configs[2])" % ngpus,
            "grid": [1024, 1024, 1024], "tau": 0.1, "l2": "inputs larger than L2"}


def run_single(args):
    """N = 1.  Both legs go through the reference-facing API (AmrSim, include/lambrex_c.h):
    value = K x AmrSim::Iterate steps with the populations resident in HBM (CUDA events on the
    library's stream); e2e = SetInitialDensity/Velocity (host arrays) + InitFromScratch +
    Iterate(K) + CalcHydroVars + bulk getters back to host arrays.  --api raw times the bare
    C-ABI kernels instead (kernel comparisons: --scheme push|pull|slab)."""
    from lambrex_b200 import lbx, workloads
    if args.api == "raw":
        return run_single_raw(args)
    from lambrex_b200 import amrsim
    amrsim.lambrexInit()
    n = args.grid
    cfg = workload_config(1)
    cfg["grid"] = [n, n, n]
    cfg["api"] = "AmrSim (liblambrex.so) -> lbx_collide_stream, FLAT level-0 storage"
    cells = float(n) ** 3
    # pinned host buffers: [rho | u] inputs and [rho | u] outputs, C-ordered like the reference's API
    L = lbx.lib()
    ncell = n ** 3
    hp = ctypes.c_void_p()
    lbx.check(L.lbx_host_alloc(ctypes.byref(hp), 8 * ncell * 8))
    host = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(8 * ncell,))
    rho0, u0 = host[:ncell], host[ncell:4 * ncell]
    rho_out, u_out = host[4 * ncell:5 * ncell], host[5 * ncell:]
    rho0[:] = workloads.pulse_density(n, n, n)
    u0[:] = 0.0

    def make_sim():
        sim = amrsim.AmrSim(n, n, n, 0, (1, 1, 1), 0.5, 0.5)
        sim.SetInitialDensityView(rho0)       # read from the pinned arrays at InitFromScratch
        sim.SetInitialVelocityView(u0)
        sim.InitFromScratch(0.0)
        return sim

    # ---- device-resident leg -------------------------------------------------
    sim = make_sim()
    sim.Iterate(args.warmup)
    lbx.sync()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    l0 = lbx.launch_count()
    with lbx.Timer() as t:
        sim.Iterate(args.steps)
    launches = lbx.launch_count() - l0
    clocks = sampler.stop()
    ms_step = t.ms / args.steps
    mlups = cells * args.steps / (t.ms * 1e-3) / 1e6
    sim.CalcHydroVars(0)
    rho = sim.GetDensityField(0, rho_out)
    mass = float(rho.sum())
    sim.close()

    # ---- end-to-end leg ------------------------------------------------------
    # one job through the public API with HOST buffers; repeated, the median job is reported and
    # every job's wall time is listed (host-side phase marks of the median job: no extra syncs)
    jobs = []
    for rep in range(max(1, args.e2e_repeat)):
        lbx.sync()
        marks = [("start", time.perf_counter())]
        sim = make_sim()
        marks.append(("ctor+InitFromScratch", time.perf_counter()))
        sim.Iterate(args.steps)
        marks.append(("Iterate (queued)", time.perf_counter()))
        sim.CalcHydroVars(0)
        rho = sim.GetDensityField(0, rho_out)
        marks.append(("CalcHydroVars+GetDensityField", time.perf_counter()))
        vel = sim.GetVelocityField(0, u_out)
        marks.append(("GetVelocityField", time.perf_counter()))
        jobs.append((marks[-1][1] - marks[0][1],
                     {marks[i][0]: round((marks[i][1] - marks[i - 1][1]) * 1e3, 3) for i in range(1, len(marks))}))
        sim.close()
    order = sorted(range(len(jobs)), key=lambda r: jobs[r][0])
    e2e_s, phases = jobs[order[len(order) // 2]]
    io = 32.0 * cells
    e2e = {"value": cells * args.steps / e2e_s / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": io / args.steps, "d2h_bytes_per_step": io / args.steps,
           "note": "one job = AmrSim ctor + SetInitialDensity/VelocityView (pinned host arrays) + InitFromScratch (H2D, "
                   "equilibrium) + Iterate(%d) + CalcHydroVars + GetDensityField/GetVelocityField (D2H); median of %d jobs"
                   % (args.steps, len(jobs)),
           "jobs_ms": [round(j[0] * 1e3, 3) for j in jobs], "phases_ms": phases,
           "check_rho_minmax": [float(rho.min()), float(rho.max())], "check_u_absmax": float(np.abs(vel).max())}

    peak, peak_src = measured_peak()
    achieved = BYTES_PER_CELL * cells / (ms_step * 1e-3) / 1e9
    tr = recorded_traffic()
    roofline = {"bound": "hbm", "kernel": "k_collide_stream<push>", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": (tr or {}).get("bytes_per_launch"), "traffic_source": (tr or {}).get("source")}
    cpu = None if args.no_cpu else cpu_baseline_block(args.cpu_sample)
    line = {"metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "check": {"total_mass_over_cells": mass / cells}}
    if not args.no_amr:
        line["extra"] = {"amr": amr_leg_single(args, peak)}
    print(json.dumps(line), flush=True)
    del rho, vel, rho0, u0, rho_out, u_out, host
    lbx.check(L.lbx_host_free(hp))


def amr_leg_single(args, peak):
    """N = 1: BASELINE configs[3] (C4) -- 2-level AMR pulse, ratio 2, 256^3 base, static central-half box, both
    couplings: the reference's live Rohde cycle (bug-compatible, diverges physically: see check) and conventional
    subcycling (FillPatch with time interpolation + average_down)."""
    from lambrex_b200.amr_workload import run_amr_case
    out = {"config": "C4: 2-level AMR pulse (ref ratio 2), %d^3 base, static box = central half, 32^3 boxes, 1 B200 "
                     "(BASELINE configs[3])" % args.amr_grid}
    for coupling in ("rohde", "subcycle"):
        out["C4_" + coupling] = run_amr_case(args.amr_grid, 2, args.amr_steps, 3, coupling, 0, 32, True, 0.0, None, peak)
    # not a headline: the same case with AMReX's max_grid_size raised to 64 (what its GPU guidance suggests) -- less ghost
    # shell per valid cell in every pass
    out["C4_rohde_boxes64"] = run_amr_case(args.amr_grid, 2, args.amr_steps, 3, "rohde", 0, 64, True, 0.0, None, peak)
    return out


def run_single_raw(args):
    from lambrex_b200 import lbx, workloads
    lbx.init()
    n = args.grid
    cfg = workload_config(1)
    cfg["grid"] = [n, n, n]
    cfg["scheme"] = args.scheme
    scheme = {"push": lbx.PUSH, "pull": lbx.PULL, "slab": None}[args.scheme]
    w = workloads.omega(0.5)
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    bx, dom = lbx.box(lo, hi), lbx.domain(lo, hi)
    cells = float(n) ** 3
    A, B = lbx.Fab(lo, hi, 15, zero=False), lbx.Fab(lo, hi, 15, zero=False)
    R, U = lbx.Fab(lo, hi, 1), lbx.Fab(lo, hi, 3)

    # pinned host buffers in fab order (x fastest): the host side of the e2e leg
    L = lbx.lib()
    nb_r, nb_u = int(cells) * 8, int(cells) * 24
    hp = ctypes.c_void_p()
    lbx.check(L.lbx_host_alloc(ctypes.byref(hp), nb_r + nb_u))
    host = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(4 * int(cells),))
    from lambrex_b200.layout import user_to_fab
    host[:int(cells)] = user_to_fab(workloads.pulse_density(n, n, n), n, n, n).reshape(-1)
    host[int(cells):] = 0.0

    def upload_init():
        lbx.check(L.lbx_h2d(R.ptr, hp.value, nb_r))
        lbx.check(L.lbx_h2d(U.ptr, hp.value + nb_r, nb_u))
        lbx.equilibrium(A, R, U, bx)

    def steps(k):
        nonlocal A, B
        for _ in range(k):
            if scheme is None:      # the multi-GPU kernel with both neighbours = this GPU
                lbx.collide_stream_slab(A, B, B, B, bx, dom, w, w)
            else:
                lbx.collide_stream(A, B, bx, dom, w, w, scheme)
            A, B = B, A

    def download_moments():
        lbx.moments(A, R, U, bx)
        lbx.check(L.lbx_d2h(hp.value, R.ptr, nb_r))
        lbx.check(L.lbx_d2h(hp.value + nb_r, U.ptr, nb_u))

    # ---- device-resident leg -------------------------------------------------
    upload_init()
    steps(args.warmup)
    lbx.sync()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    l0 = lbx.launch_count()
    with lbx.Timer() as t:
        steps(args.steps)
    launches = lbx.launch_count() - l0
    clocks = sampler.stop()
    ms_step = t.ms / args.steps
    mlups = cells * args.steps / (t.ms * 1e-3) / 1e6
    download_moments()
    lbx.sync()
    mass = float(host[:int(cells)].sum())

    # ---- end-to-end leg ------------------------------------------------------
    host[:int(cells)] = user_to_fab(workloads.pulse_density(n, n, n), n, n, n).reshape(-1)
    host[int(cells):] = 0.0
    lbx.sync()
    t0 = time.perf_counter()
    upload_init()
    steps(args.steps)
    download_moments()
    lbx.sync()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": cells * args.steps / e2e_s / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": (nb_r + nb_u) / args.steps, "d2h_bytes_per_step": (nb_r + nb_u) / args.steps,
           "note": "one job = H2D rho,u (pinned) + equilibrium + %d steps + moments + D2H rho,u" % args.steps}

    peak, peak_src = measured_peak()
    achieved = BYTES_PER_CELL * cells / (ms_step * 1e-3) / 1e9
    tr = recorded_traffic()
    roofline = {"bound": "hbm", "kernel": "k_collide_stream_slab" if scheme is None else
                "k_collide_stream<%s>" % args.scheme, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": (tr or {}).get("bytes_per_launch"), "traffic_source": (tr or {}).get("source")}

    cpu = None if args.no_cpu else cpu_baseline_block(args.cpu_sample)

    line = {"metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "check": {"total_mass_over_cells": mass / cells}}
    print(json.dumps(line), flush=True)
    lbx.check(L.lbx_host_free(hp))


def json_only_stdout():
    """The host library keeps the reference's own stdout chatter ("NX: .. NY: .. NZ: ..",
    src/AmrSim.cpp:776, grid summaries).  The bench contract is ONE JSON line on stdout: move
    file descriptor 1 to stderr for everything native and keep the real stdout for print()."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w")


def main():
    json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="N=1 cubic grid edge (default 256 = configs[1])")
    ap.add_argument("--api", default="amrsim", choices=["amrsim", "raw"],
                    help="N=1: time AmrSim::Iterate (default) or the bare C-ABI kernels")
    ap.add_argument("--scheme", default="push", choices=["push", "pull", "slab"], help="kernel for --api raw")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl", "slabsim-p2p"],
                    help="N>1 face exchange: peer stores fused into the step kernel, or packed NCCL send/recv")
    ap.add_argument("--max-grid-multi", type=int, default=32,
                    help="N>1: max_grid_size of level 0 (AMReX default 32; ownership is by whole x-y layers of these boxes)")
    ap.add_argument("--grid-multi", type=int, default=1024, help="N>1 cubic grid edge (default 1024 = configs[2])")
    ap.add_argument("--grid-multi-z", type=int, default=0, help="N>1 experiments: z extent if not cubic")
    ap.add_argument("--cpu-sample", type=int, default=128, help="edge of the CPU-baseline sample grid (cpu_baseline of the GPU arm)")
    ap.add_argument("--cpu-ref-grid", type=int, default=256, help="--impl reference: cubic grid edge (256 = the N=1 configuration itself)")
    ap.add_argument("--e2e-repeat", type=int, default=3, help="N=1: end-to-end jobs run (median reported)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-amr", action="store_true", help="skip the AMR leg (extra.amr: C4 at N=1, C5 at N>1)")
    ap.add_argument("--amr-grid", type=int, default=0, help="base grid edge of the AMR leg (default: 256 at N=1; 512 at N>=4, 256 at N=2)")
    ap.add_argument("--amr-steps", type=int, default=0, help="coarse steps of the AMR leg (default: 12 at N=1, 32 at N>1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if not args.amr_grid:
        args.amr_grid = 256 if args.gpus < 4 else 512
    if not args.amr_steps:
        args.amr_steps = 12 if args.gpus == 1 else 32
    if args.impl == "reference" or args.gpus == 1:
        # the CPU arm uses every host core it can: torchrun exports OMP_NUM_THREADS=1 to its workers,
        # which would time the OpenMP oracle on one thread (set before libgomp is loaded)
        os.environ["OMP_NUM_THREADS"] = os.environ.get("LBX_BENCH_THREADS") or str(len(os.sched_getaffinity(0)))
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus == 1:
        return run_single(args)
    from lambrex_b200.multi_gpu_bench import run_multi   # torchrun path
    helpers = {"ClockSampler": ClockSampler, "measured_peak": measured_peak, "recorded_traffic": lambda: None,
               "workload_config": workload_config, "METRIC": METRIC, "BYTES_PER_CELL": BYTES_PER_CELL}
    return run_multi(args, helpers)


if __name__ == "__main__":
    main()
