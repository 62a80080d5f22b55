"""ORACLE -- test infrastructure only, NOT product code.

Two CPU restatements of LAMBReX's single-level D3Q15 fp64 path:

* ``np_*``   : literal numpy restatement (same accumulation order, numpy never
               contracts a*b+c into an FMA), cell-vectorised.
* ``COracle``: ctypes view of ``oracle/liblbm_oracle.so`` (lbm_oracle.c), the
               C twin used for larger sizes and as the timed CPU baseline.

Both are pinned by tests/test_oracle_golden.py against the reference's golden
vectors (tests/golden/pulse_regression.npz, extracted from
/root/reference/tests/pulseRegression.h) and must agree with each other bit for
bit.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this module.

Array convention: populations f[p, k, j, i] (x fastest, component slowest --
the FArrayBox order); rho[k, j, i]; u[a, k, j, i].
User-facing input arrays are C-ordered [i][j][k] (/root/reference/include/
AmrSim.h:79-83); see ``user_to_fab``.
"""
import ctypes
import os
import subprocess
from fractions import Fraction

import numpy as np

from . import gen_tables

NV = 15
CX = np.array(gen_tables.CX)
CY = np.array(gen_tables.CY)
CZ = np.array(gen_tables.CZ)
_Mfr = gen_tables.mode_matrix()
_Mifr, _Nfr = gen_tables.inverse(_Mfr)
# value = float(num)/float(den): one correctly rounded IEEE division, like the
# reference's literals (src/AmrSim.cpp:1037-1073)
M = np.array([[float(f.numerator) / float(f.denominator) for f in r] for r in _Mfr])
MINV = np.array([[float(f.numerator) / float(f.denominator) for f in r] for r in _Mifr])
DELTA = np.diag([1.0 / NV] * 3)          # sic: 1/NMODES, src/AmrSim.cpp:1033-1035
W = np.array([float(Fraction(w)) for w in gen_tables.W])


# ----------------------------------------------------------------------------
# layout helpers (include/AmrSim.h:79-83, src/AmrSim.cpp:155-165, 233-245)
# ----------------------------------------------------------------------------
def user_to_fab(a, nx, ny, nz, ncomp=1):
    """User C-ordered [i][j][k][n] array -> fab order [n][k][j][i]."""
    a = np.asarray(a, dtype=np.float64).reshape(nx, ny, nz, ncomp)
    out = np.ascontiguousarray(a.transpose(3, 2, 1, 0))
    return out[0] if ncomp == 1 else out


# ----------------------------------------------------------------------------
# numpy restatement
# ----------------------------------------------------------------------------
def np_equilibrium(rho, u):
    """src/AmrSim.cpp:879-927, same expression order."""
    CS2 = 1.0 / 3.0
    rw0 = rho * 2.0 / 9.0
    rw1 = rho / 9.0
    rw2 = rho / 72.0
    u2 = [u[a] * u[a] for a in range(3)]
    uc = [u[a] / CS2 for a in range(3)]
    q = [u2[a] / (2.0 * CS2 * CS2) for a in range(3)]
    uv = uc[0] * uc[1]
    vw = uc[1] * uc[2]
    uw = uc[0] * uc[2]
    ms = (u2[0] + u2[1] + u2[2]) / (2.0 * CS2)
    ms2 = (u2[0] + u2[1] + u2[2]) * (1 - CS2) / (2.0 * CS2 * CS2)
    f = np.empty((NV,) + rho.shape)
    f[0] = rw0 * (1.0 - ms)
    for a in range(3):
        f[1 + 2 * a] = rw1 * (1.0 - ms + uc[a] + q[a])
        f[2 + 2 * a] = rw1 * (1.0 - ms - uc[a] + q[a])
    for p in range(7, NV):
        sx, sy, sz = float(CX[p]), float(CY[p]), float(CZ[p])
        t = 1.0 + sx * uc[0]
        t = t + sy * uc[1]
        t = t + sz * uc[2]
        t = t + (sx * sy) * uv
        t = t + (sy * sz) * vw
        t = t + (sx * sz) * uw
        f[p] = rw2 * (t + ms2)
    return f


def np_modes(f, nrows=NV):
    mode = []
    for m in range(nrows):
        acc = np.zeros(f.shape[1:])
        for p in range(NV):
            acc = acc + f[p] * M[m][p]
        mode.append(acc)
    return mode


def np_collide(f, omega_s, omega_b):
    """src/AmrSim.cpp:28-104, same accumulation order.  Returns a new array."""
    mode = np_modes(f)
    rho = mode[0]
    v = [mode[a + 1] / rho for a in range(3)]
    usq = np.zeros_like(rho)
    for a in range(3):
        usq = usq + v[a] * v[a]
    S = [[mode[4], mode[5], mode[6]], [mode[5], mode[7], mode[8]], [mode[6], mode[8], mode[9]]]
    S = [[x.copy() for x in row] for row in S]
    TrS = np.zeros_like(rho)
    for a in range(3):
        TrS = TrS + S[a][a]
    for a in range(3):
        S[a][a] = S[a][a] - (TrS / 3)
    TrS = TrS - omega_b * (TrS - rho * usq)
    for a in range(3):
        for b in range(3):
            S[a][b] = S[a][b] - omega_s * (S[a][b] - rho * (v[a] * v[b] - usq * DELTA[a][b]))
        S[a][a] = S[a][a] + (TrS / 3)
    mode[4], mode[5], mode[6] = S[0][0], S[0][1], S[0][2]
    mode[7], mode[8], mode[9] = S[1][1], S[1][2], S[2][2]
    for m in range(10, NV):
        mode[m] = np.zeros_like(rho)
    out = np.empty_like(f)
    for p in range(NV):
        acc = np.zeros_like(rho)
        for m in range(NV):
            acc = acc + mode[m] * MINV[p][m]
        out[p] = acc
    return out


def np_stream(f):
    """f'(x,i) = f(x - c_i, i), fully periodic (include/component.h:23-29)."""
    out = np.empty_like(f)
    for p in range(NV):
        out[p] = np.roll(f[p], shift=(int(CZ[p]), int(CY[p]), int(CX[p])), axis=(0, 1, 2))
    return out


OPP = np.array([0, 2, 1, 4, 3, 6, 5, 14, 13, 12, 11, 10, 9, 8, 7])     # c_OPP[p] = -c_p


def np_stream_walls(f, walls):
    """Addition (the reference aborts on non-periodic directions, src/AmrSim.cpp:788-797): streaming with solid
    no-slip walls at both faces of every direction d with walls[d] (x, y, z), half-way bounce-back --
        f'(x, p) = f(x - c_p, p)      if x - c_p lies inside the domain in every walled direction
                 = f(x, OPP[p])       otherwise (the population that left through the wall comes back reversed);
    periodic wrap in the other directions.  f is the POST-collision field [15, nz, ny, nx]."""
    out = np.empty_like(f)
    nz, ny, nx = f.shape[1:]
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    for p in range(NV):
        c = (int(CX[p]), int(CY[p]), int(CZ[p]))
        rolled = np.roll(f[p], shift=(c[2], c[1], c[0]), axis=(0, 1, 2))
        outside = np.zeros(f.shape[1:], dtype=bool)
        for d, (idx, n) in enumerate(((ii, nx), (jj, ny), (kk, nz))):
            if walls[d] and c[d] != 0:
                src = idx - c[d]
                outside |= (src < 0) | (src >= n)
        out[p] = np.where(outside, f[OPP[p]], rolled)
    return out


def np_step_walls(f, omega_s, omega_b, walls, nsteps=1):
    for _ in range(nsteps):
        f = np_stream_walls(np_collide(f, omega_s, omega_b), walls)
    return f


def np_moments(f):
    """src/AmrSim.cpp:957-971: rho = row0 . f ; u_a = (row_{a+1} . f) / rho."""
    mode = np_modes(f, 4)
    rho = mode[0]
    u = np.stack([mode[a + 1] / mode[0] for a in range(3)])
    return rho, u


def np_step(f, omega_s, omega_b, nsteps=1):
    for _ in range(nsteps):
        f = np_stream(np_collide(f, omega_s, omega_b))
    return f


# ----------------------------------------------------------------------------
# C twin
# ----------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblbm_oracle.so")
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _ptr(a):
    return a.ctypes.data_as(_dp)



def linear_moments(f, weights, per_unit_density=False):
    """Generic derived variable (include/derived_var.h:55-91 fill_box over a `calculate` that is a
    linear moment, include/d3q15_bgk.h:34-55): out[c] = sum_p weights[c][p] * f[p], accumulated from
    0.0 in increasing p, multiply and add rounded separately; divided by rho = sum_p f[p] (same
    order) when per_unit_density.  f is [15, ...]; returns [ncomp, ...]."""
    w = np.asarray(weights, dtype=np.float64).reshape(-1, NV)
    out = np.zeros((w.shape[0],) + f.shape[1:])
    for c in range(w.shape[0]):
        acc = np.zeros(f.shape[1:])
        for p in range(NV):
            acc = acc + f[p] * w[c, p]
        out[c] = acc
    if per_unit_density:
        rho = np.zeros(f.shape[1:])
        for p in range(NV):
            rho = rho + f[p]
        out = out / rho
    return out


class COracle:
    def __init__(self):
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.orc_tables.argtypes = [_dp, _dp, _ip]
        L.orc_max_threads.restype = ctypes.c_int
        L.orc_equilibrium.argtypes = [ctypes.c_int64, _dp, _dp, _dp]
        L.orc_collide.argtypes = [ctypes.c_int64, _dp, ctypes.c_double, ctypes.c_double]
        L.orc_moments.argtypes = [ctypes.c_int64, _dp, _dp, _dp]
        L.orc_stream_periodic.argtypes = [ctypes.c_int] * 3 + [_dp, _dp]
        L.orc_step_periodic.argtypes = [ctypes.c_int] * 3 + [_dp, _dp, ctypes.c_double,
                                                            ctypes.c_double, ctypes.c_int]
        L.orc_ref_passes.argtypes = [ctypes.c_int] * 3 + [_ip] * 4 + [
            _dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.orc_ref_passes.restype = ctypes.c_double
        self.L = L

    def max_threads(self):
        return int(self.L.orc_max_threads())

    def set_threads(self, n):
        self.L.orc_set_threads(int(n))

    def tables(self):
        Mm = np.empty((NV, NV))
        Mi = np.empty((NV, NV))
        c = np.empty((NV, 3), dtype=np.int32)
        self.L.orc_tables(_ptr(Mm), _ptr(Mi), c.ctypes.data_as(_ip))
        return Mm, Mi, c

    def equilibrium(self, rho, u):
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        f = np.empty((NV,) + rho.shape)
        self.L.orc_equilibrium(rho.size, _ptr(rho), _ptr(u), _ptr(f))
        return f

    def collide(self, f, omega_s, omega_b):
        f = np.array(f, dtype=np.float64, order="C", copy=True)
        self.L.orc_collide(f[0].size, _ptr(f), omega_s, omega_b)
        return f

    def moments(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        rho = np.empty(f.shape[1:])
        u = np.empty((3,) + f.shape[1:])
        self.L.orc_moments(rho.size, _ptr(f), _ptr(rho), _ptr(u))
        return rho, u

    def stream(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        nz, ny, nx = f.shape[1:]
        out = np.empty_like(f)
        self.L.orc_stream_periodic(nx, ny, nz, _ptr(f), _ptr(out))
        return out

    def step(self, f, omega_s, omega_b, nsteps=1):
        f = np.array(f, dtype=np.float64, order="C", copy=True)
        nz, ny, nx = f.shape[1:]
        tmp = np.empty_like(f)
        self.L.orc_step_periodic(nx, ny, nz, _ptr(f), _ptr(tmp), omega_s, omega_b, nsteps)
        return f

    def ref_passes(self, f, omega_s, omega_b, nsteps, edges, loop_order=0):
        """The reference's pass structure on ghosted boxes; ``edges`` = per
        direction piece boundaries [ex, ey, ez].  Returns (f_out, seconds)."""
        f = np.array(f, dtype=np.float64, order="C", copy=True)
        nz, ny, nx = f.shape[1:]
        e = [np.ascontiguousarray(x, dtype=np.int32) for x in edges]
        npc = np.array([len(x) - 1 for x in e], dtype=np.int32)
        secs = self.L.orc_ref_passes(nx, ny, nz, npc.ctypes.data_as(_ip),
                                     e[0].ctypes.data_as(_ip), e[1].ctypes.data_as(_ip),
                                     e[2].ctypes.data_as(_ip), _ptr(f), omega_s, omega_b,
                                     nsteps, loop_order)
        return f, float(secs)
