"""Round-1 experiment: cost of ghost width on the valid-only writers (profiles/r01_alignment.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from lambrex_b200 import lbx
    lbx.init()
    n, b = 256, 32
    boxes = [((i, j, k), (i + b - 1, j + b - 1, k + b - 1)) for k in range(0, n, b) for j in range(0, n, b) for i in range(0, n, b)]
    for ng in (0, 1, 2, 3, 4):
        F, G = lbx.MF(boxes, 15, ng), lbx.MF(boxes, 15, ng)
        F.setval(1.0/15)
        for name, fn in (("collide2 out-of-place", lambda: lbx.mf_collide2(F, G, 1.0, 1.0)), ("collide in-place", lambda: lbx.mf_collide(F, 1.0, 1.0))):
            for _ in range(3): fn()
            lbx.sync()
            with lbx.Timer() as t:
                for _ in range(20): fn()
            ms = t.ms / 20
            print("ngrow %d  %-22s %.4f ms  %.0f GB/s" % (ng, name, ms, 240.0 * n**3 / (ms * 1e-3) / 1e9), flush=True)
        F.free(); G.free()


if __name__ == "__main__":
    main()
