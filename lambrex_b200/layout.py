"""Host-side array layout helpers (product code, numpy only).

User-facing arrays are C-ordered [i][j][k][n] (/root/reference/include/AmrSim.h:79-83);
device fabs are x-fastest, component-slowest [n][k][j][i] (FArrayBox order).
"""
import numpy as np


def user_to_fab(a, nx, ny, nz, ncomp=1):
    a = np.asarray(a, dtype=np.float64).reshape(nx, ny, nz, ncomp)
    out = np.ascontiguousarray(a.transpose(3, 2, 1, 0))
    return out[0] if ncomp == 1 else out


def fab_to_user(a, ncomp=1):
    a = np.asarray(a)
    if ncomp == 1 and a.ndim == 3:
        a = a[None]
    return np.ascontiguousarray(a.transpose(3, 2, 1, 0)).reshape(-1)
