"""Bisect experiment 2: the uniform push kernel on single fabs that differ only in ghost width / wrap."""
import sys
sys.path.insert(0, "/root/repo")
from lambrex_b200 import lbx
lbx.init()
n = 256
def run(name, lo, hi, ng, periodic, setval=None):
    A, B = lbx.Fab(lo, hi, 15, ng), lbx.Fab(lo, hi, 15, ng)
    bx, dom = lbx.box(lo, hi), lbx.domain(lo, hi, periodic)
    R, U = lbx.Fab(lo, hi, 1), lbx.Fab(lo, hi, 3)
    if setval:
        import numpy as np
        R.upload(np.ones((1,) + tuple(h - l + 1 for l, h in zip(lo, hi))[::-1]))
        lbx.equilibrium(A, R, U, bx)
    for _ in range(3):
        lbx.collide_stream(A, B, bx, dom, 1.0, 1.0, lbx.PUSH)
    lbx.sync()
    with lbx.Timer() as t:
        for _ in range(20):
            lbx.collide_stream(A, B, bx, dom, 1.0, 1.0, lbx.PUSH)
    ms = t.ms / 20
    cells = 1.0
    for l, h in zip(lo, hi): cells *= (h - l + 1)
    print("%-52s %.4f ms  %.0f GB/s" % (name, ms, 240.0 * cells / (ms * 1e-3) / 1e9), flush=True)
    del A, B, R, U
z, m = (0, 0, 0), (n - 1,) * 3
run("256^3 no ghosts periodic, zero data", z, m, 0, (1, 1, 1))
run("256^3 no ghosts periodic, equilibrium data", z, m, 0, (1, 1, 1), True)
run("256^3 + 2 ghosts, periodic wrap (ghosts unused), zeros", z, m, 2, (1, 1, 1))
run("256^3 + 2 ghosts, periodic wrap, equilibrium data", z, m, 2, (1, 1, 1), True)
run("256^3 + 2 ghosts, non-periodic, equilibrium data", z, m, 2, (0, 0, 0), True)
run("256^3 + 2 ghosts, non-periodic, zeros", z, m, 2, (0, 0, 0))
run("256^3 + ghosts (2,0,0) periodic, equilibrium", z, m, (2, 0, 0), (1, 1, 1), True)
run("256^3 + ghosts (0,2,2) periodic, equilibrium", z, m, (0, 2, 2), (1, 1, 1), True)
