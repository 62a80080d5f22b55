"""Occupancy sweep of the fused kernels (1 GPU): cap resident CTAs/SM with dynamic shared
memory padding and time push / pull / slab at a given grid.  Prints one JSON line per point."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import lbx   # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    lbx.init()
    lo, hi = (0, 0, 0), (n - 1, n - 1, n - 1)
    bx, dom = lbx.box(lo, hi), lbx.domain(lo, hi)
    A, B = lbx.Fab(lo, hi, 15), lbx.Fab(lo, hi, 15)
    R, U = lbx.Fab(lo, hi, 1), lbx.Fab(lo, hi, 3)
    import numpy as np
    R.upload(np.ones(R.shape))
    lbx.equilibrium(A, R, U, bx)
    for scheme in ("push", "pull", "slab"):
        for pad in (0, 16 * 1024, 20 * 1024, 24 * 1024, 28 * 1024, 32 * 1024, 37 * 1024, 45 * 1024):
            lbx.set_option(lbx.OPT_SMEM_PAD, pad)

            def run(k):
                nonlocal A, B
                for _ in range(k):
                    if scheme == "slab":
                        lbx.collide_stream_slab(A, B, B, B, bx, dom, 1.0, 1.0)
                    else:
                        lbx.collide_stream(A, B, bx, dom, 1.0, 1.0, lbx.PUSH if scheme == "push" else lbx.PULL)
                    A, B = B, A
            run(5)
            lbx.sync()
            with lbx.Timer() as t:
                run(steps)
            ms = t.ms / steps
            print(json.dumps({"grid": n, "scheme": scheme, "smem_pad": pad,
                              "ctas_per_sm_cap": (227 * 1024) // (pad + 1024) if pad else None,
                              "ms_per_step": ms, "MLUPS": n ** 3 / ms / 1e3, "GBs": 240.0 * n ** 3 / ms / 1e6}),
                  flush=True)
    lbx.set_option(lbx.OPT_SMEM_PAD, 0)


if __name__ == "__main__":
    main()
