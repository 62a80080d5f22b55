"""Where the end-to-end leg of bench.py (N=1) spends its time: each phase of one AmrSim job
bracketed by a device synchronise.  Prints one JSON line (milliseconds)."""
import argparse
import ctypes
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
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    n, ncell = args.grid, args.grid ** 3
    amrsim.lambrexInit()
    L = lbx.lib()
    hp = ctypes.c_void_p()
    lbx.check(L.lbx_host_alloc(ctypes.byref(hp), 8 * ncell * 8))
    host = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(8 * ncell,))
    rho0, u0 = host[:ncell], host[ncell:4 * ncell]
    rho_out, u_out = host[4 * ncell:5 * ncell], host[5 * ncell:]
    rho0[:] = workloads.pulse_density(n, n, n)
    u0[:] = 0.0
    for rep in range(args.repeat):
        marks = []

        def mark(name):
            lbx.sync()
            marks.append((name, time.perf_counter()))

        mark("start")
        sim = amrsim.AmrSim(n, n, n, 0, (1, 1, 1), 0.5, 0.5)
        sim.SetInitialDensityView(rho0)
        sim.SetInitialVelocityView(u0)
        mark("ctor")
        sim.InitFromScratch(0.0)
        mark("InitFromScratch (alloc, H2D, transpose, equilibrium)")
        sim.Iterate(args.steps)
        mark("Iterate")
        sim.CalcHydroVars(0)
        mark("CalcHydroVars")
        sim.GetDensityField(0, rho_out)
        mark("GetDensityField (transpose, D2H)")
        sim.GetVelocityField(0, u_out)
        mark("GetVelocityField (transpose, D2H)")
        sim.close()
        mark("close")
        out = {marks[i][0]: round((marks[i][1] - marks[i - 1][1]) * 1e3, 3) for i in range(1, len(marks))}
        out["total_ms"] = round((marks[-1][1] - marks[0][1]) * 1e3, 3)
        out["rep"] = rep
        print(json.dumps(out), flush=True)
    # raw PCIe rates for the same buffers
    dev = ctypes.c_void_p()
    lbx.check(L.lbx_malloc(ctypes.byref(dev), 4 * ncell * 8))
    for name, fn, a, b in (("h2d", L.lbx_h2d, dev, hp), ("d2h", L.lbx_d2h, hp, dev)):
        lbx.sync()
        t0 = time.perf_counter()
        lbx.check(fn(a, b, 4 * ncell * 8))
        lbx.sync()
        dt = time.perf_counter() - t0
        print(json.dumps({"copy": name, "GB": 4 * ncell * 8 / 1e9, "GB_per_s": 4 * ncell * 8 / dt / 1e9}), flush=True)
    lbx.check(L.lbx_free(dev))
    lbx.check(L.lbx_host_free(hp))


if __name__ == "__main__":
    main()
