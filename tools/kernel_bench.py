"""Per-kernel roofline table of the secondary kernels: each kernel timed alone
with CUDA events over a level of N^3 cells cut into B^3 boxes (AMR storage: 2 ghost cells), its
ALGORITHMIC bytes per launch (the per-cell figure of DESIGN.md x the valid cells) divided by the
average launch time, against the measured HBM copy bandwidth.  One JSON line per kernel.
    python tools/kernel_bench.py [--grid 256] [--box 32] [--reps 20]"""
import argparse
import json
import os
import sys

import numpy as np

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
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--box", type=int, default=32)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--xghost-in-row", type=int, default=0)
    ap.add_argument("--only", default="", help="substring filter on kernel names")
    ap.add_argument("--plain-stores", type=int, default=0)
    ap.add_argument("--valid-tiling", type=int, default=0)
    ap.add_argument("--align-rows", type=int, default=0)
    ap.add_argument("--debug-skip", type=int, default=0, help="profiling only: 1 skip valid tiles, 2 skip ghost tiles")
    ap.add_argument("--row-kernel", type=int, default=1, help="1: row-owner kernel (default); 0: round-1 tile kernel")
    args = ap.parse_args()
    lbx.init()
    lbx.set_option(lbx.OPT_XGHOST_IN_ROW, args.xghost_in_row)
    lbx.set_option(lbx.OPT_DEBUG_SKIP, args.debug_skip)
    lbx.set_option(lbx.OPT_PLAIN_STORES, args.plain_stores)
    lbx.set_option(lbx.OPT_VALID_TILING, args.valid_tiling)
    lbx.set_option(lbx.OPT_ALIGN_ROWS, args.align_rows)
    lbx.set_option(lbx.OPT_ROW_KERNEL, args.row_kernel)
    n, b = args.grid, args.box
    boxes = [((i, j, k), (i + b - 1, j + b - 1, k + b - 1))
             for k in range(0, n, b) for j in range(0, n, b) for i in range(0, n, b)]
    cells = float(n) ** 3
    F, G, H = lbx.MF(boxes, 15, 2), lbx.MF(boxes, 15, 2), lbx.MF(boxes, 15, 2)
    R, U = lbx.MF(boxes, 1, 0), lbx.MF(boxes, 3, 0)
    R.setval(1.0)
    U.setval(0.01)
    lbx.mf_equilibrium(F, R, U)
    lbx.mf_equilibrium(G, R, U)
    M, _, c, _ = lbx.tables()
    cw = np.asarray(c, dtype=np.float64).T
    V3, V6, V10 = lbx.MF(boxes, 3, 0), lbx.MF(boxes, 6, 0), lbx.MF(boxes, 10, 0)
    Rg, T = lbx.MF(boxes, 1, 1), lbx.MF(boxes, 1, 0, lbx.I32)
    Rg.setval(1.0)
    # coarse level under F for the averaging kernels (fine boxes = refined coarse boxes)
    cb = b // 2
    cboxes = [((lo[0] // 2, lo[1] // 2, lo[2] // 2), (lo[0] // 2 + cb - 1, lo[1] // 2 + cb - 1, lo[2] // 2 + cb - 1)) for lo, _ in boxes]
    C1 = lbx.MF(cboxes, 15, 1)
    # FillBoundary plan of the periodic level: every ghost cell from the valid cell covering it
    per = [(sx * n, sy * n, sz * n) for sz in (-1, 0, 1) for sy in (-1, 0, 1) for sx in (-1, 0, 1)]
    descs = []
    index = {lo: q for q, (lo, _) in enumerate(boxes)}
    for k, (lo, hi) in enumerate(boxes):
        glo, ghi = tuple(v - 2 for v in lo), tuple(v + 2 for v in hi)
        for dz in (-b, 0, b):
            for dy in (-b, 0, b):
                for dx in (-b, 0, b):
                    if dx == dy == dz == 0:
                        continue
                    nlo = (lo[0] + dx, lo[1] + dy, lo[2] + dz)
                    wrapped = tuple(v % n for v in nlo)
                    src = index[wrapped]
                    shift = tuple(w - v for w, v in zip(wrapped, nlo))       # source = x + shift
                    rlo = tuple(max(g, v) for g, v in zip(glo, nlo))
                    rhi = tuple(min(g, v + b - 1) for g, v in zip(ghi, nlo))
                    descs.append(dict(dst_fab=k, src_fab=src, kind=lbx.G_COPY, shift=shift, lo=rlo, hi=rhi))
    fb = lbx.Plan(descs)
    ghost_cells = len(boxes) * ((b + 4) ** 3 - b ** 3)
    avg = lbx.Plan([dict(dst_fab=k, src_fab=k, kind=lbx.G_AVG, ratio=2, lo=clo, hi=chi) for k, (clo, chi) in enumerate(cboxes)])

    w_s = w_b = 1.0
    cases = [
        ("k_mf_equilibrium (CalcEquilibriumDist :845-936)", 152.0 * cells, lambda: lbx.mf_equilibrium(F, R, U)),
        ("k_mf_moments (CalcHydroVars :938-979)", 152.0 * cells, lambda: lbx.mf_moments(F, R, U)),
        ("k_mf_collide in place (Collide :25-107)", 240.0 * cells, lambda: lbx.mf_collide(F, w_s, w_b)),
        ("k_mf_collide out of place (FillPatch copy + Collide)", 240.0 * cells, lambda: lbx.mf_collide2(F, G, w_s, w_b)),
        ("k_mf_stream (Stream :109-122, valid grown by 1 + ring-2 zero)", 240.0 * cells, lambda: lbx.mf_stream(F, G)),
        ("k_mf_collide_stream, ghosts from own cells (FineCollide+Stream)", 240.0 * cells,
         lambda: lbx.mf_collide_stream(F, F, G, w_s, w_b)),
        ("k_mf_collide_stream + ZeroInvalidComponents", 240.0 * cells,
         lambda: lbx.mf_collide_stream(F, F, G, w_s, w_b, zero_invalid=True)),
        ("k_plan_apply COPY: FillBoundary, 2 ghost cells, 15 comps", 240.0 * ghost_cells, lambda: fb.apply(F, F)),
        ("k_mf_average_down (sum_fine_to_coarse step 1: valid + ghost)", 135.0 * len(boxes) * (b + 4) ** 3,
         lambda: lbx.mf_average_down(F, C1, 2)),
        ("k_plan_apply AVG: amrex::average_down (valid only)", 135.0 * cells, lambda: avg.apply(C1, F)),
        ("k_mf_lincomb (FillPatch time interpolation)", 360.0 * cells, lambda: lbx.mf_lincomb(H, 0.5, F, 0.5, G)),
        ("k_mf_linear_moments, 3 rows (momentum density)", (120.0 + 24.0) * cells, lambda: lbx.mf_linear_moments(F, V3, cw)),
        ("k_mf_linear_moments, 6 rows (momentum flux)", (120.0 + 48.0) * cells,
         lambda: lbx.mf_linear_moments(F, V6, np.asarray(M)[4:10])),
        ("k_mf_linear_moments, 10 rows", (120.0 + 80.0) * cells, lambda: lbx.mf_linear_moments(F, V10, np.asarray(M)[:10])),
        ("k_mf_tag_gradient (ErrorEst criterion)", 12.0 * cells, lambda: lbx.mf_tag_gradient(Rg, 1e-3, T)),
    ]
    pk = peak()
    for name, nbytes, fn in cases:
        if args.only and args.only not in name:
            continue
        for _ in range(3):
            fn()
        lbx.sync()
        with lbx.Timer() as t:
            for _ in range(args.reps):
                fn()
        ms = t.ms / args.reps
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "grid": n, "box": b, "boxes": len(boxes), "ms": round(ms, 4),
                          "algorithmic_bytes": nbytes, "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / pk, 3),
                          "peak_GBps": pk}), flush=True)


if __name__ == "__main__":
    main()
