"""Bisect experiment: the single-fab uniform kernel (k_collide_stream<push>) run on the storage of a boxed
fab set (one 256^3 box with 2 ghost cells, non-periodic: pushes land in the ghost planes) against the
batched kernel's valid-cell path on the same memory."""
import ctypes, sys
sys.path.insert(0, "/root/repo")
from lambrex_b200 import lbx
lbx.init()
lbx.set_option(lbx.OPT_ALIGN_ROWS, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = 256
boxes = [((0, 0, 0), (n - 1, n - 1, n - 1))]
F, G = lbx.MF(boxes, 15, 2), lbx.MF(boxes, 15, 2)
F.setval(1.0 / 15)
bx, dom = lbx.box((0, 0, 0), (n - 1,) * 3), lbx.domain((0, 0, 0), (n - 1,) * 3, (0, 0, 0))
fa, ga = F.fab(0), G.fab(0)
print("align", sys.argv[1:] or 0, "fab lo", list(fa.lo), "n", list(fa.n), "data %% 128 =", fa.data % 128)
L = lbx.lib()
def uni():
    lbx.check(L.lbx_collide_stream(ctypes.byref(fa), ctypes.byref(ga), ctypes.byref(bx), ctypes.byref(dom), 1.0, 1.0, lbx.PUSH))
def boxed():
    lbx.mf_collide_stream(F, F, G, 1.0, 1.0)
for name, fn, opts in (("uniform kernel on boxed storage", uni, {}),
                       ("batched kernel valid-only rows", boxed, {lbx.OPT_DEBUG_SKIP: 2, lbx.OPT_VALID_TILING: 0}),
                       ("batched kernel valid-only linear", boxed, {lbx.OPT_DEBUG_SKIP: 2, lbx.OPT_VALID_TILING: 1})):
    for k, v in opts.items():
        lbx.set_option(k, v)
    for _ in range(3): fn()
    lbx.sync()
    with lbx.Timer() as t:
        for _ in range(20): fn()
    ms = t.ms / 20
    print("%-36s %.4f ms  %.0f GB/s" % (name, ms, 240.0 * n ** 3 / (ms * 1e-3) / 1e9), flush=True)
