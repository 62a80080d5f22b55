"""ctypes binding of the C ABI in include/lbx.h (liblbx.so).

Plumbing only: loads the in-tree shared library, mirrors the POD structs, and
offers a small ``Fab`` convenience (device allocation + host<->device copies in
fab order).  All compute happens in the CUDA library; nothing here falls back to
the CPU -- a missing library or device raises ``LbxError``.
"""
import ctypes
import os

import numpy as np

NV, ND, HALO = 15, 3, 2
F64, I32 = 0, 1
PUSH, PULL = 0, 1
OPT_COLLIDE_LITERAL = 1
OPT_VALID_TILING, OPT_DEBUG_SKIP, OPT_ALIGN_ROWS, OPT_XGHOST_IN_ROW, OPT_PLAIN_STORES = 3, 4, 5, 6, 7
DEFAULT_VALID_TILING = 0      # lbx_abi.cu g_valid_linear
DEFAULT_XGHOST_IN_ROW = 0     # lbx_abi.cu g_xghost_in_row
OPT_SMEM_PAD = 2
OPT_ROW_KERNEL = 8
IPC_HANDLE_BYTES = 64
FACE_XP, FACE_XM, FACE_YP, FACE_YM, FACE_ZP, FACE_ZM = range(6)

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBX_LIB_DIR: a build variant of the native libraries next to _lib (tuning experiments: make OUT=../_lib_x RO=...)
LIB_PATH = os.path.join(_HERE, os.environ.get("LBX_LIB_DIR", "_lib"), "liblbx.so")


class LbxError(RuntimeError):
    pass


class lbx_fab(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("lo", ctypes.c_int32 * 3), ("n", ctypes.c_int32 * 3),
                ("ncomp", ctypes.c_int32), ("dtype", ctypes.c_int32)]


class lbx_box(ctypes.Structure):
    _fields_ = [("lo", ctypes.c_int32 * 3), ("hi", ctypes.c_int32 * 3)]


class lbx_domain(ctypes.Structure):
    _fields_ = [("lo", ctypes.c_int32 * 3), ("hi", ctypes.c_int32 * 3), ("periodic", ctypes.c_int32 * 3)]


class lbx_gather(ctypes.Structure):
    _fields_ = [("dst_fab", ctypes.c_int32), ("group", ctypes.c_int32), ("src_set", ctypes.c_int32), ("src_fab", ctypes.c_int32),
                ("kind", ctypes.c_int32), ("ratio", ctypes.c_int32), ("shift", ctypes.c_int32 * 3),
                ("region", lbx_box), ("value", ctypes.c_double)]


G_COPY, G_PC, G_AVG, G_CONST = 0, 1, 2, 3
OP_COPY, OP_ADD = 0, 1

# every symbol include/lbx.h declares: name -> (restype, argtypes)
_vp, _sz, _i, _d = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double
_fp, _bp, _dp = ctypes.POINTER(lbx_fab), ctypes.POINTER(lbx_box), ctypes.POINTER(lbx_domain)
SYMBOLS = {
    "lbx_init": (_i, [_i]),
    "lbx_finalize": (_i, []),
    "lbx_initialized": (_i, []),
    "lbx_last_error": (ctypes.c_char_p, []),
    "lbx_device_count": (_i, [ctypes.POINTER(_i)]),
    "lbx_device_info": (_i, [ctypes.c_char_p, _i, ctypes.POINTER(_i), ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "lbx_set_option": (_i, [_i, _i]),
    "lbx_set_stream": (_i, [_vp]),
    "lbx_sync": (_i, []),
    "lbx_launch_count": (ctypes.c_uint64, []),
    "lbx_malloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "lbx_free": (_i, [_vp]),
    "lbx_arena_release": (_i, []),
    "lbx_par_init": (_i, [_i, _i, _vp, _vp]),
    "lbx_par_info": (_i, [ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_uint64)]),
    "lbx_par_init_host": (_i, [_i, _i, _vp, _vp]), "lbx_par_finalize_host": (_i, []),
    "lbx_par_barrier": (_i, []),
    "lbx_par_allgather": (_i, [_vp, _sz, _vp]),
    "lbx_mf_create_dist": (_i, [_vp, _i, _i, _i, _i, _vp, ctypes.POINTER(_vp)]),
    "lbx_concurrent_begin": (_i, []), "lbx_concurrent_end": (_i, []),
    "lbx_arena_info": (_i, [ctypes.POINTER(_sz), ctypes.POINTER(_sz), ctypes.POINTER(ctypes.c_uint64),
                            ctypes.POINTER(ctypes.c_uint64)]),
    "lbx_memset": (_i, [_vp, _i, _sz]),
    "lbx_host_alloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "lbx_host_free": (_i, [_vp]),
    "lbx_h2d": (_i, [_vp, _vp, _sz]),
    "lbx_d2h": (_i, [_vp, _vp, _sz]),
    "lbx_d2d": (_i, [_vp, _vp, _sz]),
    "lbx_timer_start": (_i, []),
    "lbx_timer_stop": (_i, [ctypes.POINTER(ctypes.c_float)]),
    "lbx_equilibrium": (_i, [_fp, _fp, _fp, _bp]),
    "lbx_moments": (_i, [_fp, _fp, _fp, _bp]),
    "lbx_collide": (_i, [_fp, _fp, _bp, _d, _d, _fp, _i]),
    "lbx_stream": (_i, [_fp, _fp, _bp, _dp]),
    "lbx_collide_stream": (_i, [_fp, _fp, _bp, _dp, _d, _d, _i]),
    "lbx_ipc_get_handle": (_i, [_vp, ctypes.c_char_p]),
    "lbx_ipc_open_handle": (_i, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "lbx_ipc_close_handle": (_i, [_vp]),
    "lbx_peer_signal": (_i, [_vp, _vp, ctypes.c_uint64]),
    "lbx_peer_wait": (_i, [_vp, _vp, ctypes.c_uint64, ctypes.c_uint64]),
    "lbx_peer_error": (_i, []),
    "lbx_collide_stream_slab": (_i, [_fp, _fp, _fp, _fp, _bp, _dp, _d, _d]),
    "lbx_halo_pack": (_i, [_fp, _bp, _i, _vp]),
    "lbx_halo_unpack": (_i, [_fp, _bp, _i, _vp]),
    "lbx_mf_create": (_i, [_bp, _i, _i, _i, _i, ctypes.POINTER(_vp)]),
    "lbx_mf_destroy": (_i, [_vp]),
    "lbx_mf_info": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                         ctypes.POINTER(_sz)]),
    "lbx_mf_fab": (_i, [_vp, _i, _fp, _bp, ctypes.POINTER(_sz)]),
    "lbx_mf_upload": (_i, [_vp, _vp, _sz]),
    "lbx_mf_download": (_i, [_vp, _vp, _sz]),
    "lbx_mf_setval": (_i, [_vp, _d]),
    "lbx_mf_equilibrium": (_i, [_vp, _vp, _vp]),
    "lbx_mf_moments": (_i, [_vp, _vp, _vp]),
    "lbx_mf_collide": (_i, [_vp, _d, _d, _vp, _i]),
    "lbx_mf_collide2": (_i, [_vp, _vp, _d, _d, _vp, _i]),
    "lbx_mf_stream": (_i, [_vp, _vp]),
    "lbx_mf_average_down": (_i, [_vp, _vp, _i]),
    "lbx_mf_collide_stream_fillpatch": (_i, [_vp, _vp, _d, _d, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "lbx_mf_collide_stream": (_i, [_vp, _vp, _vp, _d, _d, _vp, _i, _i]),
    "lbx_mf_zero_invalid": (_i, [_vp]),
    "lbx_mf_zero_ring": (_i, [_vp, _i, _i]),
    "lbx_mf_lincomb": (_i, [_vp, _d, _vp, _d, _vp]),
    "lbx_mf_collide_stream_level": (_i, [_vp, _vp, _d, _d, _vp, _vp, _vp, _d, _vp, _d, _vp]),
    "lbx_mf_linear_moments": (_i, [_vp, _vp, _vp, _i, _i]),
    "lbx_mf_tag_gradient": (_i, [_vp, _d, _vp, _i]),
    "lbx_mf_from_user": (_i, [_vp, _vp, _bp, _i]),
    "lbx_mf_to_user": (_i, [_vp, _vp, _bp, _i]),
    "lbx_mf_to_user_local": (_i, [_vp, _vp, _bp, _i]),
    "lbx_mf_from_user_host": (_i, [_vp, _vp, _bp, _i]),
    "lbx_mf_to_user_host": (_i, [_vp, _vp, _bp, _i, _i, _i, _d]),
    "lbx_fill_f64": (_i, [_vp, _sz, _d]),
    "lbx_mf_fill_profile": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "lbx_mf_collide_stream_slab": (_i, [_vp, _vp, _dp, _d, _d]),
    "lbx_par_step_finish": (_i, []),
    "lbx_prof_begin": (_i, []),
    "lbx_prof_end": (_i, [ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_uint64)]),
    "lbx_prof_breakdown": (_i, [ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_uint64)]),
    "lbx_plan_create": (_i, [ctypes.POINTER(lbx_gather), _i, ctypes.POINTER(_vp)]),
    "lbx_plan_apply": (_i, [_vp, _vp, _vp, _vp, _i]),
    "lbx_plan_destroy": (_i, [_vp]),
    "lbx_d3q15_tables": (None, [ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_int32),
                                ctypes.POINTER(_d)]),
}

_lib = None


def lib():
    """Load liblbx.so (once).  Raises LbxError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LbxError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(no CPU fallback exists)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)      # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise LbxError(lib().lbx_last_error().decode())


def init(device=-1):
    check(lib().lbx_init(device))


def finalize():
    check(lib().lbx_finalize())


def sync():
    check(lib().lbx_sync())


def set_option(key, value):
    check(lib().lbx_set_option(key, int(value)))


def launch_count():
    return int(lib().lbx_launch_count())


def par_info():
    r, w, b = _i(0), _i(1), ctypes.c_uint64(0)
    check(lib().lbx_par_info(ctypes.byref(r), ctypes.byref(w), ctypes.byref(b)))
    return {"rank": r.value, "world": w.value, "barriers": b.value}


def device_info():
    name = ctypes.create_string_buffer(256)
    sm = _i(0)
    tot, fr = _sz(0), _sz(0)
    check(lib().lbx_device_info(name, 256, ctypes.byref(sm), ctypes.byref(tot), ctypes.byref(fr)))
    return {"name": name.value.decode(), "sm_count": sm.value, "total_bytes": tot.value, "free_bytes": fr.value}


def tables():
    M = np.empty((NV, NV))
    Mi = np.empty((NV, NV))
    c = np.empty((NV, 3), dtype=np.int32)
    w = np.empty(NV)
    P = ctypes.POINTER(_d)
    lib().lbx_d3q15_tables(M.ctypes.data_as(P), Mi.ctypes.data_as(P),
                           c.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), w.ctypes.data_as(P))
    return M, Mi, c, w


def box(lo, hi):
    b = lbx_box()
    b.lo[:] = [int(x) for x in lo]
    b.hi[:] = [int(x) for x in hi]
    return b


def domain(lo, hi, periodic=(1, 1, 1)):
    d = lbx_domain()
    d.lo[:] = [int(x) for x in lo]
    d.hi[:] = [int(x) for x in hi]
    d.periodic[:] = [int(x) for x in periodic]
    return d


def prof_begin():
    check(lib().lbx_prof_begin())


def prof_end():
    """{ms, launches, valid_cells, dropped} of the lbx_mf_collide_stream* launches since prof_begin()."""
    ms, n, cells, dr = _d(0), ctypes.c_uint64(0), _d(0), ctypes.c_uint64(0)
    check(lib().lbx_prof_end(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(cells), ctypes.byref(dr)))
    ms4, n4 = (_d * 4)(), (ctypes.c_uint64 * 4)()
    check(lib().lbx_prof_breakdown(ms4, n4))
    return {"ms": ms.value, "launches": int(n.value), "valid_cells": cells.value, "dropped": int(dr.value),
            "by_kind": {"fused_pass": [ms4[0], int(n4[0])], "gather_plans": [ms4[1], int(n4[1])],
                        "average_down": [ms4[2], int(n4[2])]}}


class Timer:
    def __enter__(self):
        check(lib().lbx_timer_start())
        return self

    def __exit__(self, *exc):
        ms = ctypes.c_float(0)
        check(lib().lbx_timer_stop(ctypes.byref(ms)))
        self.ms = float(ms.value)
        return False


class Fab:
    """One device fab: valid box [lo, hi] grown by ``ng`` ghosts per direction,
    ``ncomp`` SoA planes, x fastest.  Host mirror arrays are [comp, z, y, x]."""

    def __init__(self, lo, hi, ncomp, ng=(0, 0, 0), dtype=F64, zero=True):
        if isinstance(ng, int):
            ng = (ng, ng, ng)
        self.valid = (tuple(int(x) for x in lo), tuple(int(x) for x in hi))
        self.ng = tuple(int(g) for g in ng)
        self.ncomp, self.dtype = int(ncomp), dtype
        self.alo = tuple(l - g for l, g in zip(self.valid[0], self.ng))
        self.n = tuple(h - l + 1 + 2 * g for l, h, g in zip(self.valid[0], self.valid[1], self.ng))
        self.itemsize = 8 if dtype == F64 else 4
        self.nbytes = self.ncomp * self.n[0] * self.n[1] * self.n[2] * self.itemsize
        p = _vp()
        check(lib().lbx_malloc(ctypes.byref(p), self.nbytes))
        self.ptr = p.value
        if zero:
            check(lib().lbx_memset(self.ptr, 0, self.nbytes))
        self.c = lbx_fab()
        self.c.data = self.ptr
        self.c.lo[:] = self.alo
        self.c.n[:] = self.n
        self.c.ncomp, self.c.dtype = self.ncomp, dtype

    @property
    def shape(self):            # host mirror shape
        return (self.ncomp, self.n[2], self.n[1], self.n[0])

    @property
    def npdtype(self):
        return np.float64 if self.dtype == F64 else np.int32

    def ref(self):
        return ctypes.byref(self.c)

    def valid_box(self, grow=0):
        return box([l - grow for l in self.valid[0]], [h + grow for h in self.valid[1]])

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.npdtype)
        assert a.shape == self.shape, (a.shape, self.shape)
        check(lib().lbx_h2d(self.ptr, a.ctypes.data, self.nbytes))
        sync()

    def download(self):
        a = np.empty(self.shape, dtype=self.npdtype)
        check(lib().lbx_d2h(a.ctypes.data, self.ptr, self.nbytes))
        sync()
        return a

    def upload_valid(self, a):
        """a: [comp, nz, ny, nx] over the valid box; ghosts keep their content."""
        full = self.download()
        gx, gy, gz = self.ng
        full[:, gz:self.n[2] - gz, gy:self.n[1] - gy, gx:self.n[0] - gx] = a
        self.upload(full)

    def download_valid(self):
        gx, gy, gz = self.ng
        return np.ascontiguousarray(self.download()[:, gz:self.n[2] - gz, gy:self.n[1] - gy, gx:self.n[0] - gx])

    def free(self):
        if self.ptr:
            check(lib().lbx_free(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            if self.ptr and _lib is not None and _lib.lbx_initialized():
                _lib.lbx_free(self.ptr)
        except Exception:
            pass


def equilibrium(f, rho, u, bx):
    check(lib().lbx_equilibrium(f.ref(), rho.ref(), u.ref(), ctypes.byref(bx)))


def moments(f, rho, u, bx):
    check(lib().lbx_moments(f.ref(), rho.ref(), u.ref(), ctypes.byref(bx)))


def collide(src, dst, bx, omega_s, omega_b, mask=None, fine_val=1):
    check(lib().lbx_collide(src.ref(), dst.ref(), ctypes.byref(bx), omega_s, omega_b,
                            mask.ref() if mask is not None else None, fine_val))


def stream(src, dst, bx, dom):
    check(lib().lbx_stream(src.ref(), dst.ref(), ctypes.byref(bx), ctypes.byref(dom)))


def collide_stream_slab(src, dst, dn, up, bx, dom, omega_s, omega_b):
    """src, dst: Fab; dn, up: Fab or lbx_fab descriptor of a peer fab."""
    r = lambda x: x.ref() if isinstance(x, Fab) else ctypes.byref(x)
    check(lib().lbx_collide_stream_slab(src.ref(), dst.ref(), r(dn), r(up), ctypes.byref(bx), ctypes.byref(dom),
                                        omega_s, omega_b))


def halo_pack(f, region, face, buf_ptr):
    check(lib().lbx_halo_pack(f.ref(), ctypes.byref(region), face, buf_ptr))


def halo_unpack(f, region, face, buf_ptr):
    check(lib().lbx_halo_unpack(f.ref(), ctypes.byref(region), face, buf_ptr))


def ipc_get_handle(ptr):
    h = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
    check(lib().lbx_ipc_get_handle(ptr, h))
    return h.raw


def ipc_open_handle(handle):
    p = _vp()
    check(lib().lbx_ipc_open_handle(handle, ctypes.byref(p)))
    return p.value


def ipc_close_handle(ptr):
    check(lib().lbx_ipc_close_handle(ptr))


def peer_signal(flag_a, flag_b, value):
    check(lib().lbx_peer_signal(flag_a, flag_b, value))


def peer_wait(flag_a, flag_b, value, timeout_ns=5_000_000_000):
    check(lib().lbx_peer_wait(flag_a, flag_b, value, timeout_ns))


def fab_desc(ptr, alo, n, ncomp=NV, dtype=F64):
    """lbx_fab descriptor of memory this process did not allocate (a peer's fab)."""
    c = lbx_fab()
    c.data = ptr
    c.lo[:] = [int(x) for x in alo]
    c.n[:] = [int(x) for x in n]
    c.ncomp, c.dtype = ncomp, dtype
    return c


def collide_stream(src, dst, bx, dom, omega_s, omega_b, scheme=PUSH):
    check(lib().lbx_collide_stream(src.ref(), dst.ref(), ctypes.byref(bx), ctypes.byref(dom),
                                   omega_s, omega_b, scheme))


class MF:
    """Device fab set (lbx_mf): all boxes of one level/field.  ``boxes`` = [(lo, hi)] valid
    boxes.  Host mirrors are lists of numpy arrays [comp, z, y, x] over the GROWN boxes."""

    def __init__(self, boxes, ncomp, ngrow=0, dtype=F64):
        self.boxes = [(tuple(int(x) for x in lo), tuple(int(x) for x in hi)) for lo, hi in boxes]
        self.ncomp, self.ngrow, self.dtype = int(ncomp), int(ngrow), dtype
        arr = (lbx_box * len(self.boxes))()
        for a, (lo, hi) in zip(arr, self.boxes):
            a.lo[:] = lo
            a.hi[:] = hi
        h = _vp()
        check(lib().lbx_mf_create(arr, len(self.boxes), self.ncomp, self.ngrow, dtype, ctypes.byref(h)))
        self.h = h.value
        nb = _sz(0)
        check(lib().lbx_mf_info(self.h, None, None, None, None, ctypes.byref(nb)))
        self.nbytes = nb.value
        self.item = 8 if dtype == F64 else 4
        self.npdtype = np.float64 if dtype == F64 else np.int32
        # shapes: the LOGICAL fab (valid + ghosts) the caller sees; alloc / xoff: the allocated fab, which
        # may carry unused alignment cells in x (lbx_mf_fab)
        self.offsets, self.shapes, self.alloc, self.xoff = [], [], [], []
        for i, (lo, hi) in enumerate(self.boxes):
            f, off = lbx_fab(), _sz(0)
            check(lib().lbx_mf_fab(self.h, i, ctypes.byref(f), None, ctypes.byref(off)))
            self.offsets.append(off.value)
            self.alloc.append((self.ncomp, f.n[2], f.n[1], f.n[0]))
            self.xoff.append(lo[0] - self.ngrow - f.lo[0])
            self.shapes.append((self.ncomp,) + tuple(hi[d] - lo[d] + 1 + 2 * self.ngrow for d in (2, 1, 0)))

    def fab(self, i):
        f = lbx_fab()
        check(lib().lbx_mf_fab(self.h, i, ctypes.byref(f), None, None))
        return f

    def upload(self, arrays):
        buf = np.zeros(self.nbytes // self.item, dtype=self.npdtype)
        for a, off, shp, al, xo in zip(arrays, self.offsets, self.shapes, self.alloc, self.xoff):
            a = np.ascontiguousarray(a, dtype=self.npdtype)
            assert a.shape == shp, (a.shape, shp)
            view = buf[off // self.item: off // self.item + int(np.prod(al))].reshape(al)
            view[..., xo:xo + shp[3]] = a
        check(lib().lbx_mf_upload(self.h, buf.ctypes.data, self.nbytes))
        sync()

    def download(self):
        buf = np.empty(self.nbytes // self.item, dtype=self.npdtype)
        check(lib().lbx_mf_download(self.h, buf.ctypes.data, self.nbytes))
        sync()
        return [buf[off // self.item: off // self.item + int(np.prod(al))].reshape(al)[..., xo:xo + shp[3]].copy()
                for off, shp, al, xo in zip(self.offsets, self.shapes, self.alloc, self.xoff)]

    def setval(self, v):
        check(lib().lbx_mf_setval(self.h, float(v)))

    def free(self):
        if self.h:
            check(lib().lbx_mf_destroy(self.h))
            self.h = None

    def __del__(self):
        try:
            if self.h and _lib is not None and _lib.lbx_initialized():
                _lib.lbx_mf_destroy(self.h)
        except Exception:
            pass


def mf_equilibrium(f, rho, u):
    check(lib().lbx_mf_equilibrium(f.h, rho.h, u.h))


def mf_moments(f, rho, u):
    check(lib().lbx_mf_moments(f.h, rho.h, u.h))


def mf_collide(f, omega_s, omega_b, mask=None, fine_val=1):
    check(lib().lbx_mf_collide(f.h, omega_s, omega_b, mask.h if mask is not None else None, fine_val))


def mf_collide2(src, dst, omega_s, omega_b, mask=None, fine_val=1):
    check(lib().lbx_mf_collide2(src.h, dst.h, omega_s, omega_b, mask.h if mask is not None else None, fine_val))


def mf_collide_stream(src_valid, src_ghost, dst, omega_s, omega_b, mask=None, fine_val=1, zero_invalid=False):
    check(lib().lbx_mf_collide_stream(src_valid.h, src_ghost.h if src_ghost is not None else None, dst.h, omega_s, omega_b,
                                      mask.h if mask is not None else None, fine_val, 1 if zero_invalid else 0))


def mf_stream(src, dst):
    check(lib().lbx_mf_stream(src.h, dst.h))


def mf_average_down(fine, crse, ratio=2):
    check(lib().lbx_mf_average_down(fine.h, crse.h, int(ratio)))


def mf_lincomb(dst, a, x, b, y):
    check(lib().lbx_mf_lincomb(dst.h, float(a), x.h, float(b), y.h))


def mf_tag_gradient(rho, threshold, tags, set_val=1):
    check(lib().lbx_mf_tag_gradient(rho.h, float(threshold), tags.h, int(set_val)))


def mf_linear_moments(f, out, weights, per_unit_density=False):
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1, 15))
    check(lib().lbx_mf_linear_moments(f.h, out.h, w.ctypes.data_as(ctypes.c_void_p), w.shape[0], 1 if per_unit_density else 0))


def mf_zero_invalid(f):
    check(lib().lbx_mf_zero_invalid(f.h))


def mf_zero_ring(f, depth, comp):
    check(lib().lbx_mf_zero_ring(f.h, depth, comp))


class Plan:
    """descs: iterable of dicts(dst_fab, src_set, src_fab, kind, ratio, shift, lo, hi, value),
    grouped by ascending dst_fab."""

    def __init__(self, descs):
        descs = list(descs)
        arr = (lbx_gather * max(len(descs), 1))()
        for a, d in zip(arr, descs):
            a.dst_fab, a.src_set, a.src_fab = d["dst_fab"], d.get("src_set", 0), d.get("src_fab", 0)
            a.group = d.get("group", 0)
            a.kind, a.ratio = d.get("kind", G_COPY), d.get("ratio", 1)
            a.shift[:] = [int(x) for x in d.get("shift", (0, 0, 0))]
            a.region.lo[:] = [int(x) for x in d["lo"]]
            a.region.hi[:] = [int(x) for x in d["hi"]]
            a.value = float(d.get("value", 0.0))
        h = _vp()
        check(lib().lbx_plan_create(arr, len(descs), ctypes.byref(h)))
        self.h, self.n = h.value, len(descs)

    def apply(self, dst, src0=None, src1=None, op=OP_COPY):
        check(lib().lbx_plan_apply(self.h, dst.h, src0.h if src0 is not None else None,
                                   src1.h if src1 is not None else None, op))

    def free(self):
        if self.h:
            check(lib().lbx_plan_destroy(self.h))
            self.h = None
