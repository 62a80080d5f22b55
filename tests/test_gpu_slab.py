"""GPU parity tests of the slab-decomposed uniform path (lambrex_b200/slab.py,
lbx_collide_stream_slab, lbx_halo_pack/unpack, lbx_peer_*).  Single-GPU tests emulate
several ranks in one process (neighbour fabs are ordinary device pointers); the 2-GPU test
launches tools/slab_check.py under torchrun and needs >= 2 devices."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from lambrex_b200 import lbx, workloads
from lambrex_b200.boxes import slab_partition
from oracle import lbm_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _ctx():
    lbx.init()
    yield
    lbx.set_option(lbx.OPT_COLLIDE_LITERAL, 0)


def random_f(shape, seed):
    rng = np.random.default_rng(seed)
    rho = 1.0 + 0.05 * rng.standard_normal(shape)
    u = 0.02 * rng.standard_normal((3,) + shape)
    return orc.np_equilibrium(rho, u) * (1.0 + 0.01 * rng.standard_normal((15,) + shape))


@pytest.mark.parametrize("literal", [1, 0])
@pytest.mark.parametrize("nslabs,shape", [(1, (6, 5, 7)), (2, (8, 5, 130)), (3, (9, 4, 6)), (4, (4, 3, 5))])
def test_slab_step_equals_single_box_step(coracle, literal, nslabs, shape):
    """K slabs exchanging through neighbour pointers == one periodic box, bit for bit."""
    lbx.set_option(lbx.OPT_COLLIDE_LITERAL, literal)
    nz, ny, nx = shape
    f = random_f(shape, seed=nslabs)
    ws, wb = 1.0 / 0.6, 1.0 / 0.8
    dom = lbx.domain((0, 0, 0), (nx - 1, ny - 1, nz - 1))
    slabs = slab_partition(nz, nslabs)
    A = [lbx.Fab((0, 0, lo), (nx - 1, ny - 1, hi), 15) for lo, hi in slabs]
    B = [lbx.Fab((0, 0, lo), (nx - 1, ny - 1, hi), 15) for lo, hi in slabs]
    for a, (lo, hi) in zip(A, slabs):
        a.upload(f[:, lo:hi + 1])
    steps = 3
    for _ in range(steps):
        for r in range(nslabs):
            lbx.collide_stream_slab(A[r], B[r], B[(r - 1) % nslabs], B[(r + 1) % nslabs],
                                    A[r].valid_box(), dom, ws, wb)
        A, B = B, A
    got = np.concatenate([a.download() for a in A], axis=1)
    # single-box fused kernel: must agree bit for bit (same arithmetic per cell)
    S, T = lbx.Fab((0, 0, 0), (nx - 1, ny - 1, nz - 1), 15), lbx.Fab((0, 0, 0), (nx - 1, ny - 1, nz - 1), 15)
    S.upload(f)
    for _ in range(steps):
        lbx.collide_stream(S, T, S.valid_box(), dom, ws, wb, lbx.PUSH)
        S, T = T, S
    assert np.array_equal(got, S.download())
    want = coracle.step(f, ws, wb, steps)
    if literal:
        assert np.array_equal(got, want)
    else:
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-12


@pytest.mark.parametrize("face", range(6))
def test_halo_pack_unpack_five_crossing_populations(face):
    nx, ny, nz, g = 7, 6, 5, 1
    F = lbx.Fab((0, 0, 0), (nx - 1, ny - 1, nz - 1), 15, ng=g)
    rng = np.random.default_rng(face)
    a = rng.random(F.shape)
    F.upload(a)
    axis, sign = face // 2, (1 if face % 2 == 0 else -1)
    c = [orc.CX, orc.CY, orc.CZ][axis]
    pops = [p for p in range(15) if c[p] == sign]
    assert len(pops) == 5
    lo, hi = [-g, -g, -g], [nx - 1 + g, ny - 1 + g, nz - 1 + g]
    lo[axis] = hi[axis] = (hi[axis] if sign > 0 else lo[axis])      # the ghost plane beyond that face
    reg = lbx.box(lo, hi)
    ncell = int(np.prod([h - l + 1 for l, h in zip(lo, hi)]))
    buf = lbx.Fab((0, 0, 0), (5 * ncell - 1, 0, 0), 1)
    lbx.halo_pack(F, reg, face, buf.ptr)
    sl = tuple(slice(l + g, h + g + 1) for l, h in zip(lo, hi))[::-1]
    want = np.stack([a[p][sl].reshape(-1) for p in pops])
    assert np.array_equal(buf.download().reshape(5, ncell), want)
    # unpack into a zeroed fab touches exactly those 5 populations on that region
    G = lbx.Fab((0, 0, 0), (nx - 1, ny - 1, nz - 1), 15, ng=g)
    lbx.halo_unpack(G, reg, face, buf.ptr)
    exp = np.zeros_like(a)
    for p in pops:
        exp[p][sl] = a[p][sl]
    assert np.array_equal(G.download(), exp)


@pytest.mark.parametrize("halo", ["p2p", "nccl"])
def test_slabsim_world1_matches_oracle(coracle, halo):
    torch = pytest.importorskip("torch")
    from lambrex_b200.slab import SlabSim, shear_slab
    nx, ny, nz, tau, steps = 12, 16, 10, 0.1, 7
    sim = SlabSim(nx, ny, nz, tau, tau, halo=halo)
    rho, u = shear_slab(nx, ny, nz, 0, nz - 1)
    sim.set_initial(torch.from_numpy(rho), torch.from_numpy(u))
    sim.step(steps)
    got = sim.download_f()
    w = workloads.omega(tau)
    want = coracle.step(coracle.equilibrium(rho[0], u), w, w, steps)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-12
    rh, uh = torch.empty(rho.shape, dtype=torch.float64), torch.empty(u.shape, dtype=torch.float64)
    sim.moments(rh, uh)
    sim.sync()
    ro, uo = coracle.moments(want)
    assert np.max(np.abs(rh.numpy()[0] - ro) / ro) < 1e-12 and np.max(np.abs(uh.numpy() - uo)) < 1e-12
    sim.close()


def test_peer_flags_order_and_timeout():
    L = lbx.lib()
    p = ctypes.c_void_p()
    lbx.check(L.lbx_malloc(ctypes.byref(p), 16))
    lbx.check(L.lbx_memset(p, 0, 16))
    lbx.peer_signal(p.value, p.value + 8, 3)
    lbx.peer_wait(p.value, p.value + 8, 3, 1_000_000_000)
    lbx.sync()
    assert L.lbx_peer_error() == 0
    lbx.peer_wait(p.value, None, 4, 20_000_000)        # never signalled: must give up after 20 ms
    with pytest.raises(lbx.LbxError):
        lbx.sync()
    assert L.lbx_peer_error() == 1 and L.lbx_peer_error() == 0
    lbx.sync()
    lbx.check(L.lbx_free(p))


def test_two_gpu_slabs_bit_equal_single_gpu():
    n = ctypes.c_int(0)
    lbx.lib().lbx_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "slab_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SLAB_CHECK_OK" in r.stdout
