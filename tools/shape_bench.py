"""Uniform fused step on ONE GPU over a box of any shape (nx ny nz): separates the kernel's shape dependence
from the multi-GPU exchange (a 1024 x 1024 x 128 slab on world = 1).  One JSON line per shape.
    python tools/shape_bench.py --dims 1024 1024 128 --dims 256 256 256 [--scheme push|slab] [--steps 20]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import lbx   # noqa: E402


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, action="append", required=True)
    ap.add_argument("--scheme", default="push", choices=["push", "pull", "slab", "slabsync"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    lbx.init()
    pk = peak()
    for nx, ny, nz in args.dims:
        lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
        bx, dom = lbx.box(lo, hi), lbx.domain(lo, hi)
        A, B = lbx.Fab(lo, hi, 15, zero=False), lbx.Fab(lo, hi, 15, zero=False)
        R, U = lbx.Fab(lo, hi, 1), lbx.Fab(lo, hi, 3)
        lbx.check(lbx.lib().lbx_memset(R.ptr, 0, R.nbytes))
        if args.scheme != "slabsync":
            lbx.equilibrium(A, R, U, bx)      # rho = 0 field: timing only

        if args.scheme == "slabsync":      # the distributed step kernel on one GPU (no flags): lbx_mf_collide_stream_slab
            for f in (A, B):
                f.free()
            A, B = lbx.MF([(lo, hi)], 15, 0), lbx.MF([(lo, hi)], 15, 0)

        def steps(k):
            nonlocal A, B
            for _ in range(k):
                if args.scheme == "slabsync":
                    lbx.check(lbx.lib().lbx_mf_collide_stream_slab(A.h, B.h, ctypes.byref(dom), 1.0, 1.0))
                elif args.scheme == "slab":
                    lbx.collide_stream_slab(A, B, B, B, bx, dom, 1.0, 1.0)
                else:
                    lbx.collide_stream(A, B, bx, dom, 1.0, 1.0, lbx.PUSH if args.scheme == "push" else lbx.PULL)
                A, B = B, A
        steps(args.warmup)
        lbx.sync()
        with lbx.Timer() as t:
            steps(args.steps)
        ms = t.ms / args.steps
        cells = float(nx) * ny * nz
        gbs = 240.0 * cells / (ms * 1e-3) / 1e9
        print(json.dumps({"dims": [nx, ny, nz], "scheme": args.scheme, "ms_per_step": round(ms, 4),
                          "MLUPS": round(cells / ms / 1e3, 1), "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / pk, 4)}),
              flush=True)
        for f in (A, B, R, U):
            f.free()



if __name__ == "__main__":
    main()
