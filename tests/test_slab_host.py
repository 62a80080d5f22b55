"""Host-side logic of the multi-GPU uniform path (lambrex_b200/slab.py) on CPU:
slab ownership, neighbour maps, and the halo message pairing over torch.distributed
(gloo, world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest

from lambrex_b200.boxes import chop_1d, slab_partition
from lambrex_b200.slab import SlabLayout, allgather_objects, exchange_z_halos, pulse_slab, shear_slab
from lambrex_b200 import workloads
from lambrex_b200.layout import user_to_fab


def test_slab_partition_covers_domain():
    for nz, w in [(1024, 8), (50, 3), (7, 7), (130, 4)]:
        lay = SlabLayout(8, 8, nz, w)
        planes = []
        for r in range(w):
            lo, hi = lay.slab(r)
            assert hi >= lo
            planes += list(range(lo, hi + 1))
            assert lay.owner(hi + 1) == lay.up(r) and lay.owner(lo - 1) == lay.dn(r)
        assert planes == list(range(nz))
        assert sum(lay.cells(r) for r in range(w)) == 64 * nz
    with pytest.raises(ValueError):
        SlabLayout(4, 4, 2, 3)


def test_base_grid_chop_matches_amrex_examples():
    # 10x10x50 -> z pieces [0,23],[24,49]; 256 -> 8 x 32
    assert chop_1d(50) == [0, 24, 50]
    assert chop_1d(10) == [0, 10]
    assert chop_1d(256) == list(range(0, 257, 32))
    assert slab_partition(10, 3) == [(0, 3), (4, 6), (7, 9)]


def test_slab_initial_conditions_match_global_workloads():
    nx, ny, nz = 6, 5, 20
    rho = user_to_fab(workloads.pulse_density(nx, ny, nz), nx, ny, nz)
    assert np.array_equal(pulse_slab(nx, ny, nz, 4, 11)[0], rho[4:12])
    rc, uc = workloads.shear_wave(nx, ny, nz)
    r, u = shear_slab(nx, ny, nz, 3, 9)
    assert np.array_equal(u, user_to_fab(uc, nx, ny, nz, 3)[:, 3:10])
    assert np.array_equal(r[0], user_to_fab(rc, nx, ny, nz)[3:10])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = SlabLayout(4, 4, 12, world)
        got = allgather_objects({"rank": rank, "slab": lay.slab(rank)})
        assert [g["rank"] for g in got] == list(range(world))
        assert [tuple(g["slab"]) for g in got] == lay.slabs
        for rep in range(3):     # repeated exchanges must keep pairing up
            su = torch.full((5,), 100.0 * rank + 1 + rep, dtype=torch.float64)    # to the rank above
            sd = torch.full((5,), 100.0 * rank + 2 + rep, dtype=torch.float64)    # to the rank below
            rd, ru = torch.zeros(5, dtype=torch.float64), torch.zeros(5, dtype=torch.float64)
            exchange_z_halos(su, sd, rd, ru, rank, lay)
            assert rd[0].item() == 100.0 * lay.dn(rank) + 1 + rep      # what the rank below sent up
            assert ru[0].item() == 100.0 * lay.up(rank) + 2 + rep      # what the rank above sent down
        q.put((rank, "ok"))
    except Exception as e:       # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_message_pairing_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def _allgather_worker(rank, world, port, q):
    import ctypes
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lambrex_b200 import amrsim
        hook = amrsim.make_allgather_hook()
        # the two message kinds of the distributed AMR path: a 64-byte CUDA-IPC handle per allocation and a
        # padded list of 16-byte tag runs per regrid (TagBoxArray::collate)
        for nbytes in (64, 16 * 37):
            send = (ctypes.c_ubyte * nbytes)(*[(7 * rank + i) % 251 for i in range(nbytes)])
            recv = (ctypes.c_ubyte * (nbytes * world))()
            assert hook(ctypes.cast(send, ctypes.c_void_p), nbytes, ctypes.cast(recv, ctypes.c_void_p), None) == 0
            for r in range(world):
                assert list(recv[r * nbytes:(r + 1) * nbytes]) == [(7 * r + i) % 251 for i in range(nbytes)]
        # ownership is computed, not communicated: every rank derives the same map
        ba = amrsim.meta_base_grids((64, 64, 128))
        own = amrsim.meta_distribution(ba, world)
        got = allgather_objects(own)
        assert all(g == own for g in got)
        q.put((rank, "ok"))
    except Exception as e:       # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_amr_allgather_hook_and_ownership_over_gloo():
    """Host plumbing of the distributed AMR path (lambrexInitParallel) on CPU, world_size 2."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_allgather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def _regrid_worker(rank, world, port, q, want):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lambrex_b200 import amrsim
        amrsim.metaParallelInit()
        m = amrsim.MetaMesh((64, 48, 32), 2, 16)
        got = []
        m.set_static(0, (10, 8, 6), (41, 30, 21))
        got.append([m.boxes(l) for l in range(3)])
        m.set_static(1, (40, 30, 20), (75, 55, 37))
        got.append([m.boxes(l) for l in range(3)])
        m.set_static(0, (12, 8, 6), (43, 30, 21))            # moves level 1 (and what nests in it)
        got.append([m.boxes(l) for l in range(3)])
        m.unset_static(1)
        got.append([m.boxes(l) for l in range(3)])
        fin = m.finest_level()
        del m
        amrsim.metaParallelFinalise()
        q.put((rank, "ok" if (got == want["grids"] and fin == want["finest"]) else "grids differ from the single-process run"))
    except Exception as e:       # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_regrid_gives_the_single_process_grids_over_gloo(world):
    """The distributed regrid of the AMR path on CPU: every rank tags only the boxes it owns
    (TagBoxArray without storage for the others), TagBoxArray::collate merges the tag runs through the
    allgather hook, and every rank must arrive at exactly the box lists of a single-process run."""
    import torch.multiprocessing as mp
    from lambrex_b200 import amrsim
    m = amrsim.MetaMesh((64, 48, 32), 2, 16)
    grids = []
    m.set_static(0, (10, 8, 6), (41, 30, 21))
    grids.append([m.boxes(l) for l in range(3)])
    m.set_static(1, (40, 30, 20), (75, 55, 37))
    grids.append([m.boxes(l) for l in range(3)])
    m.set_static(0, (12, 8, 6), (43, 30, 21))
    grids.append([m.boxes(l) for l in range(3)])
    m.unset_static(1)
    grids.append([m.boxes(l) for l in range(3)])
    want = {"grids": grids, "finest": m.finest_level()}
    assert len(grids[1][2]) > 0 and grids[1] != grids[2] and len(grids[3][2]) == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_regrid_worker, args=(r, world, port, q, want)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
