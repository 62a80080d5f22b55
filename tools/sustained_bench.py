"""Does the fused step slow down under SUSTAINED load on one GPU?  1024 x 1024 x 128 cells (the per-GPU share of the
1024^3 run on 8 GPUs), the plain single-GPU kernel, timed in windows of 100 steps for --seconds, with nvidia-smi's
clocks / temperatures / power sampled beside it.  One JSON line per window (name the 1 -> 8 limiter)."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import lbx   # noqa: E402


def smi():
    q = "clocks.sm,clocks.mem,temperature.gpu,temperature.memory,power.draw,clocks_event_reasons.active"
    try:
        out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        return dict(zip(q.split(","), [x.strip() for x in out.split(",")]))
    except Exception as e:
        return {"error": str(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=40.0)
    ap.add_argument("--dims", type=int, nargs=3, default=[1024, 1024, 128])
    args = ap.parse_args()
    lbx.init()
    nx, ny, nz = args.dims
    lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
    bx, dom = lbx.box(lo, hi), lbx.domain(lo, hi)
    A, B = lbx.Fab(lo, hi, 15), lbx.Fab(lo, hi, 15)
    t_start = time.time()
    w = 0
    while time.time() - t_start < args.seconds:
        with lbx.Timer() as t:
            for _ in range(100):
                lbx.collide_stream_slab(A, B, B, B, bx, dom, 1.0, 1.0)
                A, B = B, A
        print(json.dumps({"window": w, "t_s": round(time.time() - t_start, 2), "ms_per_step": round(t.ms / 100, 4), "smi": smi()}), flush=True)
        w += 1


if __name__ == "__main__":
    main()
