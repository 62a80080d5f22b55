"""Uniform single-level periodic path sharded as z-slabs, one per GPU (SURVEY.md 8e,
BASELINE.json configs[2]).  Host-side plumbing only: partition, device memory, CUDA-IPC
handle exchange and step ordering; every cell update happens in liblbx.so.

One reference step (CollideAndStream, /root/reference/include/AmrSim.h:89-94) on rank r:

  halo="p2p"  : lbx_peer_wait(step-1) -> lbx_collide_stream_slab(src, dst, peer_dn.dst,
                peer_up.dst) -> lbx_peer_signal(step).  The boundary-plane CTAs store the 5
                populations crossing each z face directly into the neighbour's fab over
                NVLink (CUDA-IPC peer pointers); there is no separate exchange pass.
  halo="nccl" : fabs carry one ghost plane per z side; the step kernel scatters into them,
                lbx_halo_pack packs the 5 crossing populations of each ghost plane, NCCL
                send/recv (torch.distributed) moves them, lbx_halo_unpack writes them into
                the neighbour's boundary plane.  Baseline transport.

torch is used for torch.distributed and for the rho/u/halo buffers' device memory.
"""
import ctypes

import numpy as np

from . import lbx
from .boxes import slab_partition


class SlabLayout:
    """Pure host logic (no CUDA): who owns which planes and who the neighbours are."""

    def __init__(self, nx, ny, nz, world):
        if world > nz:
            raise ValueError("more ranks than z-planes")
        self.n = (int(nx), int(ny), int(nz))
        self.world = int(world)
        self.slabs = slab_partition(nz, world)          # [(klo, khi)] inclusive

    def slab(self, rank):
        return self.slabs[rank]

    def up(self, rank):          # owner of plane khi+1 (periodic)
        return (rank + 1) % self.world

    def dn(self, rank):          # owner of plane klo-1 (periodic)
        return (rank - 1) % self.world

    def owner(self, k):
        k %= self.n[2]
        for r, (lo, hi) in enumerate(self.slabs):
            if lo <= k <= hi:
                return r
        raise AssertionError

    def cells(self, rank):
        lo, hi = self.slabs[rank]
        return self.n[0] * self.n[1] * (hi - lo + 1)

    def valid(self, rank):
        lo, hi = self.slabs[rank]
        return (0, 0, lo), (self.n[0] - 1, self.n[1] - 1, hi)


def allgather_objects(obj, group=None):
    """obj from every rank (list indexed by rank); identity when torch.distributed is off."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def exchange_z_halos(send_up, send_dn, recv_dn, recv_up, rank, layout, group=None):
    """Send `send_up` to the rank above and `send_dn` to the rank below; receive the matching
    buffers.  Posting order (send_up, recv_dn, send_dn, recv_up) is the same on every rank so
    that with world == 2 (both neighbours are the same peer) messages pair up in order."""
    import torch.distributed as dist
    if layout.world == 1:
        recv_dn.copy_(send_up)
        recv_up.copy_(send_dn)
        return
    up, dn = layout.up(rank), layout.dn(rank)
    ops = [dist.P2POp(dist.isend, send_up, up, group), dist.P2POp(dist.irecv, recv_dn, dn, group),
           dist.P2POp(dist.isend, send_dn, dn, group), dist.P2POp(dist.irecv, recv_up, up, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class SlabSim:
    """Device state of one rank's slab + the step loop."""

    def __init__(self, nx, ny, nz, tau_s, tau_b, rank=0, world=1, halo="p2p", group=None,
                 timeout_ns=20_000_000_000, split=True):
        import torch
        self.torch = torch
        self.layout = SlabLayout(nx, ny, nz, world)
        self.rank, self.world, self.halo, self.group = rank, world, halo, group
        self.timeout_ns = timeout_ns
        self.omega_s, self.omega_b = 1.0 / (tau_s + 0.5), 1.0 / (tau_b + 0.5)
        lo, hi = self.layout.valid(rank)
        self.lo, self.hi = lo, hi
        self.box = lbx.box(lo, hi)
        self.cells = self.layout.cells(rank)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        nzl = hi[2] - lo[2] + 1
        # torch-owned device memory for the hydrodynamic fields (plumbing)
        self.rho_t = torch.zeros((1, nzl, ny, nx), dtype=torch.float64, device=dev)
        self.u_t = torch.zeros((3, nzl, ny, nx), dtype=torch.float64, device=dev)
        self.R = lbx.fab_desc(self.rho_t.data_ptr(), lo, (nx, ny, nzl), 1)
        self.U = lbx.fab_desc(self.u_t.data_ptr(), lo, (nx, ny, nzl), 3)
        self.step_id = 0
        self.cur = 0
        # interior / boundary-plane split of the step (p2p, world > 1, slab at least 3 planes thick)
        self.split = bool(split) and nzl >= 3
        self.box_edges = [lbx.box(lo, (hi[0], hi[1], lo[2])), lbx.box((lo[0], lo[1], hi[2]), hi)]
        self.box_interior = lbx.box((lo[0], lo[1], lo[2] + 1), (hi[0], hi[1], hi[2] - 1))
        glo, ghi = (0, 0, 0), (nx - 1, ny - 1, nz - 1)
        if halo == "p2p":
            self.dom = lbx.domain(glo, ghi, (1, 1, 1))
            self.F = [lbx.Fab(lo, hi, lbx.NV, zero=False), lbx.Fab(lo, hi, lbx.NV, zero=False)]
            self._setup_peers()
        elif halo == "nccl":
            self.dom = lbx.domain(glo, ghi, (1, 1, 0))
            self.F = [lbx.Fab(lo, hi, lbx.NV, ng=(0, 0, 1), zero=False),
                      lbx.Fab(lo, hi, lbx.NV, ng=(0, 0, 1), zero=False)]
            n = 5 * nx * ny
            self.buf = {k: torch.empty(n, dtype=torch.float64, device=dev)
                        for k in ("send_up", "send_dn", "recv_up", "recv_dn")}
            # everything on torch's current stream so NCCL ops are ordered with our kernels
            lbx.check(lbx.lib().lbx_set_stream(torch.cuda.current_stream().cuda_stream or 1))   # 1 = cudaStreamLegacy
        else:
            raise ValueError("halo must be 'p2p' or 'nccl'")

    # ------------------------------------------------------------------ peers (CUDA IPC)
    def _setup_peers(self):
        L = lbx.lib()
        fp = ctypes.c_void_p()
        lbx.check(L.lbx_malloc(ctypes.byref(fp), 16))
        lbx.check(L.lbx_memset(fp, 0, 16))
        lbx.sync()
        self.flags = fp.value            # [0]: written by the rank below, [1]: by the rank above
        self._opened = []
        if self.world == 1:
            self.peer_dn = self.peer_up = [self.F[0].c, self.F[1].c]
            self.sig_dn = self.sig_up = None
            return
        mine = {"handles": [lbx.ipc_get_handle(self.F[0].ptr), lbx.ipc_get_handle(self.F[1].ptr),
                            lbx.ipc_get_handle(self.flags)],
                "alo": self.F[0].alo, "n": self.F[0].n}
        everyone = allgather_objects(mine, self.group)
        opened = {}

        def peer(r):
            if r not in opened:
                info = everyone[r]
                ptrs = [lbx.ipc_open_handle(h) for h in info["handles"]]
                self._opened.extend(ptrs)
                opened[r] = ([lbx.fab_desc(ptrs[0], info["alo"], info["n"]),
                              lbx.fab_desc(ptrs[1], info["alo"], info["n"])], ptrs[2])
            return opened[r]

        self.peer_up, up_flags = peer(self.layout.up(self.rank))
        self.peer_dn, dn_flags = peer(self.layout.dn(self.rank))
        self.sig_up = up_flags + 0       # I am the rank BELOW my upper neighbour -> its flag [0]
        self.sig_dn = dn_flags + 8       # I am the rank ABOVE my lower neighbour -> its flag [1]

    # ------------------------------------------------------------------ init / output
    def set_initial(self, rho_host=None, u_host=None):
        """rho_host [1,nzl,ny,nx], u_host [3,nzl,ny,nx]: torch (pinned) host tensors in fab
        order, or None to keep what rho_t / u_t hold.  Then f <- f_eq(rho, u)."""
        L = lbx.lib()
        self.barrier()      # no neighbour may still read or write the buffers re-initialised here
        if rho_host is not None:
            lbx.check(L.lbx_h2d(self.rho_t.data_ptr(), rho_host.data_ptr(), self.cells * 8))
        if u_host is not None:
            lbx.check(L.lbx_h2d(self.u_t.data_ptr(), u_host.data_ptr(), self.cells * 24))
        self.cur = 0
        lbx.check(L.lbx_equilibrium(self.F[0].ref(), ctypes.byref(self.R), ctypes.byref(self.U),
                                    ctypes.byref(self.box)))

    def moments(self, rho_host=None, u_host=None):
        L = lbx.lib()
        self.finish()
        lbx.check(L.lbx_moments(self.F[self.cur].ref(), ctypes.byref(self.R), ctypes.byref(self.U),
                                ctypes.byref(self.box)))
        if rho_host is not None:
            lbx.check(L.lbx_d2h(rho_host.data_ptr(), self.rho_t.data_ptr(), self.cells * 8))
        if u_host is not None:
            lbx.check(L.lbx_d2h(u_host.data_ptr(), self.u_t.data_ptr(), self.cells * 24))

    # ------------------------------------------------------------------ stepping
    def step(self, n=1):
        for _ in range(n):
            src, dst = self.F[self.cur], self.F[1 - self.cur]
            if self.halo == "p2p":
                if self.world > 1 and self.split:
                    # Only the two boundary planes touch a neighbour (they read the 5 populations it
                    # stored last step and store 5 into its fab), so only they are ordered against
                    # it: wait -> boundary planes -> signal, then the interior planes with the plain
                    # push kernel while the neighbours already work on their next step.
                    lbx.peer_wait(self.flags, self.flags + 8, self.step_id, self.timeout_ns)
                    for bx in self.box_edges:
                        lbx.collide_stream_slab(src, dst, self.peer_dn[1 - self.cur], self.peer_up[1 - self.cur],
                                                bx, self.dom, self.omega_s, self.omega_b)
                    self.step_id += 1
                    lbx.peer_signal(self.sig_dn, self.sig_up, self.step_id)
                    lbx.collide_stream(src, dst, self.box_interior, self.dom, self.omega_s, self.omega_b, lbx.PUSH)
                else:
                    if self.world > 1:
                        lbx.peer_wait(self.flags, self.flags + 8, self.step_id, self.timeout_ns)
                    lbx.collide_stream_slab(src, dst, self.peer_dn[1 - self.cur], self.peer_up[1 - self.cur],
                                            self.box, self.dom, self.omega_s, self.omega_b)
                    self.step_id += 1
                    if self.world > 1:
                        lbx.peer_signal(self.sig_dn, self.sig_up, self.step_id)
            else:
                lbx.collide_stream_slab(src, dst, dst, dst, self.box, self.dom, self.omega_s, self.omega_b)
                self.step_id += 1
                self._nccl_exchange(dst)
            self.cur = 1 - self.cur

    def _nccl_exchange(self, dst):
        (x0, y0, klo), (x1, y1, khi) = self.lo, self.hi
        b = self.buf
        plane = lambda k: lbx.box((x0, y0, k), (x1, y1, k))
        lbx.halo_pack(dst, plane(khi + 1), lbx.FACE_ZP, b["send_up"].data_ptr())
        lbx.halo_pack(dst, plane(klo - 1), lbx.FACE_ZM, b["send_dn"].data_ptr())
        exchange_z_halos(b["send_up"], b["send_dn"], b["recv_dn"], b["recv_up"], self.rank, self.layout,
                         self.group)
        lbx.halo_unpack(dst, plane(klo), lbx.FACE_ZP, b["recv_dn"].data_ptr())
        lbx.halo_unpack(dst, plane(khi), lbx.FACE_ZM, b["recv_up"].data_ptr())

    def finish(self):
        """Make the neighbours' stores of the last step visible before this rank reads."""
        if self.halo == "p2p" and self.world > 1:
            lbx.peer_wait(self.flags, self.flags + 8, self.step_id, self.timeout_ns)

    def sync(self):
        lbx.sync()

    def barrier(self):
        """All ranks: every queued step finished on every GPU (host-side, not on the step path)."""
        self.finish()
        lbx.sync()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def download_f(self):
        """Populations of the valid slab, [15, nzl, ny, nx] (tests)."""
        self.finish()
        return self.F[self.cur].download_valid()

    def close(self):
        lbx.sync()
        for p in getattr(self, "_opened", []):
            lbx.ipc_close_handle(p)
        self._opened = []
        if self.halo == "nccl":
            lbx.check(lbx.lib().lbx_set_stream(None))
        for f in self.F:
            f.free()
        if getattr(self, "flags", None):
            lbx.check(lbx.lib().lbx_free(self.flags))
            self.flags = None


def pulse_slab(nx, ny, nz, klo, khi, amplitude=0.01):
    """rho of the planar pulse (/root/reference/examples/amr_pulse.cpp:21-32) restricted to
    planes klo..khi, fab order [1, nzl, ny, nx] (numpy).  Same arithmetic as
    workloads.pulse_density without materialising the global array."""
    col = np.ones(nz)
    col[nz // 2 - 1] += amplitude
    # the example averages nz consecutive FLAT entries from offset (numel + ny*nz)/2, i.e. a
    # column entered at k0 and continued in the next column: same values, this summation order
    k0 = ((nx * ny * nz + ny * nz) // 2) % nz
    z_mean = 0.0
    for kk in range(nz):
        z_mean += col[(k0 + kk) % nz]
    z_mean /= nz
    col = col / z_mean
    return np.ascontiguousarray(np.broadcast_to(col[klo:khi + 1, None, None], (khi - klo + 1, ny, nx)))[None]


def shear_slab(nx, ny, nz, klo, khi, U=0.01):
    """rho = 1, u_x = U sin(2 pi j / NY) on planes klo..khi: ([1,nzl,ny,nx], [3,nzl,ny,nx])."""
    nzl = khi - klo + 1
    rho = np.ones((1, nzl, ny, nx))
    u = np.zeros((3, nzl, ny, nx))
    j = np.arange(ny, dtype=np.float64)
    u[0] = (U * np.sin(2.0 * np.pi * j / ny))[None, :, None]
    return rho, u
