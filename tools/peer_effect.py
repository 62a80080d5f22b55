"""Which part of a multi-GPU run slows the fused step kernel down on EVERY GPU (name the 1->8
limiter with evidence)?  torchrun entry, one rank per GPU; every rank times the SAME single-GPU kernel launch
(lbx_collide_stream_slab with itself as both neighbours, 1024 x 1024 x nz/world cells) in four states:
  t1  alone on its GPU's memory, nothing shared (all ranks run at the same time: shared power / cooling shows here);
  t2  after a small buffer of the neighbour was opened through CUDA IPC (peer access enabled on the device);
  t3  with population buffers allocated AFTER peer access was enabled;
  t4  with the population buffers themselves exported and opened by the neighbour (the state of a real run).
Rank 0 prints one JSON line with the per-rank milliseconds of every state."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import lbx   # noqa: E402


def timed(A, B, bx, dom, steps=20):
    for _ in range(3):
        lbx.collide_stream_slab(A, B, B, B, bx, dom, 1.0, 1.0)
        A, B = B, A
    lbx.sync()
    dist.barrier()
    with lbx.Timer() as t:
        for _ in range(steps):
            lbx.collide_stream_slab(A, B, B, B, bx, dom, 1.0, 1.0)
            A, B = B, A
    return t.ms / steps


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lbx.init(local)
    nx = ny = 1024
    nz = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    lo, hi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
    bx, dom = lbx.box(lo, hi), lbx.domain(lo, hi)
    res = {}
    A, B = lbx.Fab(lo, hi, 15), lbx.Fab(lo, hi, 15)
    res["t1_nothing_shared"] = timed(A, B, bx, dom)
    # a small buffer of the neighbour opened through IPC: enables peer access
    p = ctypes.c_void_p()
    lbx.check(lbx.lib().lbx_malloc(ctypes.byref(p), 4 << 20))
    handles = [None] * world
    dist.all_gather_object(handles, lbx.ipc_get_handle(p.value))
    peer_small = lbx.ipc_open_handle(handles[(rank + 1) % world]) if world > 1 else None
    res["t2_peer_access_enabled"] = timed(A, B, bx, dom)
    A.free(); B.free()
    lbx.check(lbx.lib().lbx_arena_release())
    A, B = lbx.Fab(lo, hi, 15), lbx.Fab(lo, hi, 15)
    res["t3_allocated_after_peer_access"] = timed(A, B, bx, dom)
    hs = [None] * world
    dist.all_gather_object(hs, (lbx.ipc_get_handle(A.ptr), lbx.ipc_get_handle(B.ptr)))
    opened = [lbx.ipc_open_handle(h) for h in hs[(rank + 1) % world]] if world > 1 else []
    res["t4_buffers_exported_and_opened"] = timed(A, B, bx, dom)
    if world > 1:
        # t5: the face-crossing populations of the two boundary planes really go to the neighbour over NVLink (and the
        # neighbour's come in), unsynchronised -- timing only: the domain is two slabs tall, this rank's box the lower one,
        # the neighbour's buffers stand for the upper one
        dom2 = lbx.domain(lo, (nx - 1, ny - 1, 2 * nz - 1))
        peerA = lbx.fab_desc(opened[0], (0, 0, nz), (nx, ny, nz))
        peerB = lbx.fab_desc(opened[1], (0, 0, nz), (nx, ny, nz))
        cur = [A, B]
        peers = [peerA, peerB]

        def run(k):
            for _ in range(k):
                lbx.collide_stream_slab(cur[0], cur[1], peers[1], peers[1], bx, dom2, 1.0, 1.0)
                cur.reverse()
                peers.reverse()
        run(3)
        lbx.sync()
        dist.barrier()
        with lbx.Timer() as t:
            run(20)
        res["t5_face_stores_over_nvlink_unsynchronised"] = t.ms / 20
        # t6 / t7: ONLY the two boundary planes, stores to the neighbour / kept local (domain = own slab): the cost of
        # the NVLink stores by themselves
        planes = [lbx.box((0, 0, 0), (nx - 1, ny - 1, 0)), lbx.box((0, 0, nz - 1), (nx - 1, ny - 1, nz - 1))]
        for key, d, pr in (("t6_two_boundary_planes_remote", dom2, True), ("t7_two_boundary_planes_local", dom, False)):
            def run2(k):
                for _ in range(k):
                    for pb in planes:
                        lbx.collide_stream_slab(A, B, peerB if pr else B, peerB if pr else B, pb, d, 1.0, 1.0)
            run2(3)
            lbx.sync()
            dist.barrier()
            with lbx.Timer() as t:
                run2(20)
            res[key] = t.ms / 20
        # t9: as t5, but ONLY rank 0 stores over NVLink (the others keep their faces local): is it the outgoing stores or
        # the incoming ones that slow the whole launch down?
        def run9(k):
            for _ in range(k):
                if rank == 0:
                    lbx.collide_stream_slab(cur[0], cur[1], peers[1], peers[1], bx, dom2, 1.0, 1.0)
                else:
                    lbx.collide_stream_slab(cur[0], cur[1], cur[1], cur[1], bx, dom, 1.0, 1.0)
                cur.reverse()
                peers.reverse()
        run9(3)
        lbx.sync()
        dist.barrier()
        with lbx.Timer() as t:
            run9(20)
        res["t9_only_rank0_stores_remotely"] = t.ms / 20
        # t10: boundary planes (remote) and interior planes as SEPARATE launches back to back on one stream
        def run10(k):
            for _ in range(k):
                for pb in planes:
                    lbx.collide_stream_slab(A, B, peerB, peerB, pb, dom2, 1.0, 1.0)
                lbx.collide_stream_slab(A, B, B, B, inner_box, dom, 1.0, 1.0)
        inner_box = lbx.box((0, 0, 1), (nx - 1, ny - 1, nz - 2))
        run10(3)
        lbx.sync()
        dist.barrier()
        with lbx.Timer() as t:
            run10(20)
        res["t10_boundary_then_interior_separate_launches"] = t.ms / 20
        # t8: the interior planes alone (nz - 2 planes, all stores local) after the remote traffic above
        inner = lbx.box((0, 0, 1), (nx - 1, ny - 1, nz - 2))
        lbx.sync()
        dist.barrier()
        with lbx.Timer() as t:
            for _ in range(20):
                lbx.collide_stream_slab(A, B, B, B, inner, dom, 1.0, 1.0)
        res["t8_interior_planes_only_local"] = t.ms / 20
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        out = {"dims": [nx, ny, nz], "world": world, "single_gpu_reference_ms": "tools/shape_bench.py --scheme slab on an otherwise idle box"}
        for k in res:
            out[k] = [round(r[k], 4) for r in allres]
        print(json.dumps(out), flush=True)
    lbx.sync()
    dist.barrier()
    for q in opened:
        lbx.ipc_close_handle(q)
    if peer_small:
        lbx.ipc_close_handle(peer_small)
    dist.barrier()
    lbx.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
