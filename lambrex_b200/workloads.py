"""Synthetic initial conditions for the benchmark and parity workloads (host side).

User-facing arrays are C-ordered like the reference's SetInitialDensity /
SetInitialVelocity inputs: rho[(i*NY + j)*NZ + k], u[((i*NY + j)*NZ + k)*3 + n]
(/root/reference/include/AmrSim.h:79-83, tests/catch2InitTests.cpp:84-107).
"""
import numpy as np


def pulse_density(nx, ny, nz, amplitude=0.01):
    """Planar density pulse of /root/reference/examples/amr_pulse.cpp:21-32:
    rho = 1, += amplitude on the plane k = nz/2 - 1, then divided by the mean
    over the z-column at flat offset (numel + ny*nz)/2."""
    numel = nx * ny * nz
    rho = np.ones(numel, dtype=np.float64)
    k = nz // 2
    r3 = rho.reshape(nx, ny, nz)
    r3[:, :, k - 1] += amplitude
    start = (numel + ny * nz) // 2
    z_mean = 0.0
    for kk in range(nz):            # same left-to-right accumulation as the example
        z_mean += rho[start + kk]
    z_mean /= nz
    rho /= z_mean
    return rho


def shear_wave(nx, ny, nz, U=0.01):
    """rho = 1, u_x = U sin(2 pi j / NY), u_y = u_z = 0."""
    rho = np.ones(nx * ny * nz, dtype=np.float64)
    u = np.zeros((nx, ny, nz, 3), dtype=np.float64)
    j = np.arange(ny, dtype=np.float64)
    u[:, :, :, 0] = (U * np.sin(2.0 * np.pi * j / ny))[None, :, None]
    return rho, u.reshape(-1)


def omega(tau):
    """Relaxation rate from the (shifted) relaxation time,
    /root/reference/src/AmrSim.cpp:125-126."""
    return 1.0 / (tau + 0.5)
