"""Bisect experiment 4: x-ghost cells (row pitch 260 for 256 valid cells) on the source only / the
destination only of the uniform push and pull kernels."""
import sys
sys.path.insert(0, "/root/repo")
from lambrex_b200 import lbx
lbx.init()
n = 256
z, m = (0, 0, 0), (n - 1,) * 3
bx, dom = lbx.box(z, m), lbx.domain(z, m, (1, 1, 1))
for scheme, sname in ((lbx.PUSH, "push"), (lbx.PULL, "pull")):
    for sg, dg in (((0, 0, 0), (0, 0, 0)), ((2, 0, 0), (0, 0, 0)), ((0, 0, 0), (2, 0, 0)), ((2, 0, 0), (2, 0, 0)), ((4, 0, 0), (4, 0, 0)), ((16, 0, 0), (16, 0, 0))):
        A, B = lbx.Fab(z, m, 15, sg), lbx.Fab(z, m, 15, dg)
        for _ in range(3):
            lbx.collide_stream(A, B, bx, dom, 1.0, 1.0, scheme)
        lbx.sync()
        with lbx.Timer() as t:
            for _ in range(20):
                lbx.collide_stream(A, B, bx, dom, 1.0, 1.0, scheme)
        ms = t.ms / 20
        print("%s  src ghosts %-10s dst ghosts %-10s %.4f ms  %.0f GB/s" % (sname, sg, dg, ms, 240.0 * n ** 3 / (ms * 1e-3) / 1e9), flush=True)
        del A, B
