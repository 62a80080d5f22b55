"""ctypes binding of include/lambrex_c.h (liblambrex.so): the reference's ``AmrSim`` surface
(/root/reference/include/AmrSim.h:127-155) as a Python class with the same method names, plus
the white-box accessors of the reference's test subclass (/root/reference/tests/AmrTest.h).

Plumbing only: every method is one call into the C++ host library, which drives the CUDA
kernels of liblbx.so.  No CPU fallback: without the built libraries or without a GPU the
calls raise ``LambrexError``.
"""
import ctypes
import os

import numpy as np

from . import lbx as _lbx

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBX_LIB_DIR: a build variant of the native libraries next to _lib (tuning experiments: make OUT=../_lib_x RO=...)
LIB_PATH = os.path.join(_HERE, os.environ.get("LBX_LIB_DIR", "_lib"), "liblambrex.so")
NL_DENSITY, NL_VELOCITY = -1.0, -3e8
DISTFN, DENSITY, VELOCITY, DISTFN_NEXT, FINE_MASK = range(5)
TAG_CLEAR, TAG_BUF, TAG_SET = 0, 1, 2


ROHDE, SUBCYCLE = 0, 1      # AmrSim::Coupling (include/lambrex_c.h)


class LambrexError(RuntimeError):
    pass


_i, _d, _vp, _sz = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
_ip, _dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
_i3 = ctypes.c_int * 3
SYMBOLS = {
    "lbx_sim_global_init": (_i, []), "lbx_sim_global_finalise": (_i, []),
    "lbx_sim_last_error": (ctypes.c_char_p, []),
    "lbx_sim_create": (_i, [_i, _i, _i, _i, _ip, _d, _d, ctypes.POINTER(_vp)]),
    "lbx_sim_destroy": (_i, [_vp]),
    "lbx_sim_set_max_grid_size": (_i, [_vp, _i]), "lbx_sim_set_uniform_fast_path": (_i, [_vp, _i]),
    "lbx_sim_set_rohde_fusion": (_i, [_vp, _i]), "lbx_sim_set_coupling": (_i, [_vp, _i]),
    "lbx_sim_write_checkpoint": (_i, [_vp, ctypes.c_char_p]), "lbx_sim_read_checkpoint": (_i, [_vp, ctypes.c_char_p]),
    "lbx_sim_get_linear_moment_field": (_i, [_vp, _i, _dp, _i, _i, _d, _dp, _sz]),
    "lbx_sim_set_gradient_refinement": (_i, [_vp, _i, _d]), "lbx_sim_unset_gradient_refinement": (_i, [_vp, _i]),
    "lbx_sim_set_regrid_interval": (_i, [_vp, _i]), "lbx_sim_num_regrids": (_i, [_vp]),
    "lbx_sim_plan_cache_size": (_i, []), "lbx_sim_allow_walls": (_i, [_i]),
    "lbx_sim_write_plotfile": (_i, [_vp, ctypes.c_char_p]),
    "lbx_sim_set_static_box": (_i, [_vp, _i, _ip, _ip]), "lbx_sim_regrid_all": (_i, [_vp]),
    "lbx_sim_set_initial_density_profile": (_i, [_vp, _i, _dp, _sz]),
    "lbx_sim_set_initial_velocity_profile": (_i, [_vp, _i, _dp, _sz]),
    "lbx_sim_local_box": (_i, [_vp, _ip, _ip]),
    "lbx_sim_set_initial_density_local_view": (_i, [_vp, _dp, _sz]),
    "lbx_sim_set_initial_velocity_local_view": (_i, [_vp, _dp, _sz]),
    "lbx_sim_get_local_density_field": (_i, [_vp, _i, _dp, _sz]),
    "lbx_sim_get_local_velocity_field": (_i, [_vp, _i, _dp, _sz]),
    "lbx_sim_global_init_parallel": (_i, [_i, _i, _vp, _vp]), "lbx_sim_set_parallel_view": (_i, [_i, _i]),
    "lbx_sim_owner": (_i, [_vp, _i, _i, ctypes.POINTER(_i)]),
    "lbx_sim_set_initial_density": (_i, [_vp, _dp, _sz]), "lbx_sim_set_initial_velocity": (_i, [_vp, _dp, _sz]),
    "lbx_sim_set_initial_density_view": (_i, [_vp, _dp, _sz]), "lbx_sim_set_initial_velocity_view": (_i, [_vp, _dp, _sz]),
    "lbx_sim_init_from_scratch": (_i, [_vp, _d]), "lbx_sim_regrid": (_i, [_vp, _i, _d]),
    "lbx_sim_iterate": (_i, [_vp, _i]), "lbx_sim_calc_hydro_vars": (_i, [_vp, _i]),
    "lbx_sim_calc_equilibrium_dist": (_i, [_vp, _i]),
    "lbx_sim_get_density": (_i, [_vp, _i, _i, _i, _i, _dp]),
    "lbx_sim_get_velocity": (_i, [_vp, _i, _i, _i, _i, _i, _dp]),
    "lbx_sim_get_density_field": (_i, [_vp, _i, _dp, _sz]), "lbx_sim_get_velocity_field": (_i, [_vp, _i, _dp, _sz]),
    "lbx_sim_get_time": (_i, [_vp, _i, _dp]), "lbx_sim_get_time_step": (_i, [_vp, _i, _ip]),
    "lbx_sim_get_dims": (_i, [_vp, _ip]), "lbx_sim_get_extent": (_i, [_vp, _i, _ip, _ip]),
    "lbx_sim_set_static_refinement": (_i, [_vp, _i, _ip, _ip]), "lbx_sim_unset_static_refinement": (_i, [_vp, _i]),
    "lbx_sim_max_level": (_i, [_vp]), "lbx_sim_finest_level": (_i, [_vp]),
    "lbx_sim_ref_ratio": (_i, [_vp, _i, _ip]),
    "lbx_sim_num_boxes": (_i, [_vp, _i]), "lbx_sim_get_boxes": (_i, [_vp, _i, _ip]),
    "lbx_sim_field_empty": (_i, [_vp, _i, _i]), "lbx_sim_field_num_boxes": (_i, [_vp, _i, _i]),
    "lbx_sim_field_boxes": (_i, [_vp, _i, _i, _ip]),
    "lbx_sim_field_fab": (_i, [_vp, _i, _i, _i, _dp, _sz, _ip]),
    "lbx_sim_get_tau": (_i, [_vp, _i, _dp, _dp]), "lbx_sim_get_mass": (_i, [_vp, _i, _dp]),
    "lbx_sim_get_dt": (_i, [_vp, _i, _dp]), "lbx_sim_num_levels_allocated": (_i, [_vp]),
    "lbx_sim_call_error_est": (_i, [_vp, _i, _ip, _i, _i, ctypes.c_char_p, _sz]),
    "lbx_sim_call_make_new_level_from_scratch": (_i, [_vp, _i, _ip, _i, _d]),
    "lbx_sim_call_make_new_level_from_coarse": (_i, [_vp, _i, _ip, _i]),
    "lbx_sim_call_remake_level": (_i, [_vp, _i, _d, _ip, _i]),
    "lbx_sim_call_clear_level": (_i, [_vp, _i]),
    "lbx_meta_base_grids": (_i, [_ip, _i, _ip, _i]), "lbx_meta_max_size": (_i, [_ip, _i, _i, _ip, _i]),
    "lbx_meta_simplify": (_i, [_ip, _i, _ip, _i]), "lbx_meta_complement": (_i, [_ip, _ip, _i, _ip, _i]),
    "lbx_meta_cluster": (_i, [_ip, _i, _d, _ip, _i]), "lbx_meta_distribution": (_i, [_ip, _i, _i, _ip]),
    "lbx_meta_distribution_runs": (_i, [_ip, _i, _i, _i, _ip]),
    "lbx_meta_parallel_init": (_i, [_i, _i, _vp, _vp]), "lbx_meta_parallel_finalise": (_i, []),
    "lbx_meta_mesh_create": (_i, [_ip, _i, _i, ctypes.POINTER(_vp)]), "lbx_meta_mesh_destroy": (_i, [_vp]),
    "lbx_meta_mesh_set_static": (_i, [_vp, _i, _ip, _ip]), "lbx_meta_mesh_unset_static": (_i, [_vp, _i]),
    "lbx_meta_mesh_finest_level": (_i, [_vp]), "lbx_meta_mesh_boxes": (_i, [_vp, _i, _ip, _i]),
    "lbx_meta_mesh_log": (_i, [_vp, ctypes.c_char_p, _i]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lbx.lib()                     # liblbx.so first (RTLD_GLOBAL), liblambrex.so links against it
        if not os.path.exists(LIB_PATH):
            raise LambrexError("%s not built: run __graft_entry__.build() (no CPU fallback exists)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise LambrexError(lib().lbx_sim_last_error().decode())


def lambrexInit():
    _check(lib().lbx_sim_global_init())


_ALLGATHER_T = ctypes.CFUNCTYPE(_i, _vp, _sz, _vp, _vp)
_par_keep = []


def lambrexInitParallel(group=None):
    """One process per GPU (torchrun): boxes of every level are owned by ranks, neighbours' boxes
    are read over NVLink through CUDA-IPC.  torch.distributed (an initialised process group whose
    CPU backend is gloo) is the host plumbing: it only carries IPC handles.  Collective."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cb = make_allgather_hook(group)
    _check(lib().lbx_sim_global_init_parallel(rank, world, cb, None))


def make_allgather_hook(group=None):
    """The C callback lbx_sim_global_init_parallel takes -- allgather(send, nbytes, recv, user): `nbytes`
    from every rank into `recv`, rank order -- over torch.distributed's CPU backend.  It carries CUDA-IPC
    handles at allocation time and the tag runs of a regrid; no field data."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def allgather(send, nbytes, recv, _user):
        try:
            mine = torch.frombuffer(bytearray(ctypes.string_at(send, nbytes)), dtype=torch.uint8)
            parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            ctypes.memmove(recv, torch.cat(parts).numpy().tobytes(), nbytes * world)
            return 0
        except Exception as exc:            # never unwind through the C frames
            print("lambrex allgather failed:", exc, flush=True)
            return 1

    cb = _ALLGATHER_T(allgather)
    _par_keep.append(cb)
    return cb


def allowWalls(on=True):
    """Non-periodic directions become solid no-slip walls (half-way bounce-back) instead of aborting like the
    reference; uniform single-GPU path only.  Process-wide; call before constructing AmrSim."""
    _check(lib().lbx_sim_allow_walls(int(bool(on))))


def setParallelView(rank, nranks):
    _check(lib().lbx_sim_set_parallel_view(rank, nranks))


def lambrexFinalise():
    _check(lib().lbx_sim_global_finalise())


def _boxes_in(boxes):
    flat = [int(v) for lo, hi in boxes for v in (*lo, *hi)]
    return (ctypes.c_int * max(len(flat), 1))(*flat), len(boxes)


def _boxes_out(arr, n):
    return [((arr[6 * i], arr[6 * i + 1], arr[6 * i + 2]), (arr[6 * i + 3], arr[6 * i + 4], arr[6 * i + 5]))
            for i in range(n)]


class AmrSim:
    """Same constructor and methods as the reference class; user arrays are C-ordered
    rho[(i*NY+j)*NZ+k], u[((i*NY+j)*NZ+k)*3+n] (include/AmrSim.h:79-83)."""

    def __init__(self, nx, ny, nz, max_level, periodicity, tau_s, tau_b):
        h = _vp()
        _check(lib().lbx_sim_create(nx, ny, nz, max_level, _i3(*periodicity), tau_s, tau_b, ctypes.byref(h)))
        self._h = h.value
        self._keep = None

    def close(self):
        if self._h:
            _check(lib().lbx_sim_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.lbx_sim_destroy(self._h)
        except Exception:
            pass

    # ---- reference surface
    def SetMaxGridSize(self, n):
        _check(lib().lbx_sim_set_max_grid_size(self._h, n))

    def SetUniformFastPath(self, on):
        _check(lib().lbx_sim_set_uniform_fast_path(self._h, int(on)))

    def Owner(self, level, box):
        r = _i()
        _check(lib().lbx_sim_owner(self._h, level, box, ctypes.byref(r)))
        return r.value

    def SetRohdeFusion(self, on):
        _check(lib().lbx_sim_set_rohde_fusion(self._h, int(on)))

    def SetCoupling(self, coupling):
        """ROHDE (default, the reference's live path) or SUBCYCLE (conventional subcycling with
        time-interpolated FillPatch and average_down)."""
        _check(lib().lbx_sim_set_coupling(self._h, int(coupling)))

    def WritePlotFile(self, directory):
        """AMReX-format plotfile (rho, ux, uy, uz of every level on its boxes)"""
        _check(lib().lbx_sim_write_plotfile(self._h, str(directory).encode()))

    def SetStaticBox(self, level, lo, hi):
        """record the static box of `level` without regridding (Regrid() then regrids every level once)"""
        _check(lib().lbx_sim_set_static_box(self._h, level, _i3(*lo), _i3(*hi)))

    def Regrid(self):
        _check(lib().lbx_sim_regrid_all(self._h))

    def SetGradientRefinement(self, level, threshold):
        _check(lib().lbx_sim_set_gradient_refinement(self._h, level, float(threshold)))

    def UnsetGradientRefinement(self, level):
        _check(lib().lbx_sim_unset_gradient_refinement(self._h, level))

    def SetRegridInterval(self, n):
        _check(lib().lbx_sim_set_regrid_interval(self._h, int(n)))

    def NumRegrids(self):
        return int(lib().lbx_sim_num_regrids(self._h))

    @staticmethod
    def PlanCacheSize():
        return int(lib().lbx_sim_plan_cache_size())

    def _set(self, fn, v):
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype=np.float64)).reshape(-1))
        _check(fn(self._h, a.ctypes.data_as(_dp), a.size))

    def SetInitialDensity(self, rho):
        self._set(lib().lbx_sim_set_initial_density, rho)

    def SetInitialVelocity(self, u):
        self._set(lib().lbx_sim_set_initial_velocity, u)

    def SetInitialDensityView(self, rho):
        """Zero-copy: `rho` (contiguous float64 array, ideally pinned) is read at InitFromScratch."""
        assert rho.dtype == np.float64 and rho.flags["C_CONTIGUOUS"]
        self._keep = (self._keep or []) + [rho]
        _check(lib().lbx_sim_set_initial_density_view(self._h, rho.ctypes.data_as(_dp), rho.size))

    def SetInitialVelocityView(self, u):
        assert u.dtype == np.float64 and u.flags["C_CONTIGUOUS"]
        self._keep = (self._keep or []) + [u]
        _check(lib().lbx_sim_set_initial_velocity_view(self._h, u.ctypes.data_as(_dp), u.size))

    # ---- distributed runs: each rank states and reads its own part (include/lambrex_c.h)
    def SetInitialDensityProfile(self, axis, rho_of_axis):
        a = np.ascontiguousarray(rho_of_axis, dtype=np.float64).reshape(-1)
        _check(lib().lbx_sim_set_initial_density_profile(self._h, int(axis), a.ctypes.data_as(_dp), a.size))

    def SetInitialVelocityProfile(self, axis, u_of_axis):
        a = np.ascontiguousarray(u_of_axis, dtype=np.float64).reshape(-1)
        _check(lib().lbx_sim_set_initial_velocity_profile(self._h, int(axis), a.ctypes.data_as(_dp), a.size))

    def LocalBox(self):
        lo, hi = _i3(0, 0, 0), _i3(0, 0, 0)
        _check(lib().lbx_sim_local_box(self._h, lo, hi))
        return tuple(lo), tuple(hi)

    def SetInitialDensityLocalView(self, rho):
        assert rho.dtype == np.float64 and rho.flags["C_CONTIGUOUS"]
        self._keep = (self._keep or []) + [rho]
        _check(lib().lbx_sim_set_initial_density_local_view(self._h, rho.ctypes.data_as(_dp), rho.size))

    def SetInitialVelocityLocalView(self, u):
        assert u.dtype == np.float64 and u.flags["C_CONTIGUOUS"]
        self._keep = (self._keep or []) + [u]
        _check(lib().lbx_sim_set_initial_velocity_local_view(self._h, u.ctypes.data_as(_dp), u.size))

    def GetLocalDensityField(self, level, out=None):
        """this rank's cells of the level, [nx, ny, nz_local] (C order); no communication"""
        lo, hi = self.LocalBox()
        dims = tuple(h - l + 1 for l, h in zip(lo, hi))
        if out is None:
            out = np.empty(dims)
        _check(lib().lbx_sim_get_local_density_field(self._h, level, out.ctypes.data_as(_dp), out.size))
        return out.reshape(dims)

    def GetLocalVelocityField(self, level, out=None):
        lo, hi = self.LocalBox()
        dims = tuple(h - l + 1 for l, h in zip(lo, hi)) + (3,)
        if out is None:
            out = np.empty(dims)
        _check(lib().lbx_sim_get_local_velocity_field(self._h, level, out.ctypes.data_as(_dp), out.size))
        return out.reshape(dims)

    def InitFromScratch(self, time=0.0):
        _check(lib().lbx_sim_init_from_scratch(self._h, time))

    def regrid(self, lbase, time):
        _check(lib().lbx_sim_regrid(self._h, lbase, time))

    def Iterate(self, nsteps):
        _check(lib().lbx_sim_iterate(self._h, nsteps))

    def CalcHydroVars(self, level):
        _check(lib().lbx_sim_calc_hydro_vars(self._h, level))

    def CalcEquilibriumDist(self, level):
        _check(lib().lbx_sim_calc_equilibrium_dist(self._h, level))

    def GetDensity(self, i, j, k, level):
        v = _d()
        _check(lib().lbx_sim_get_density(self._h, i, j, k, level, ctypes.byref(v)))
        return v.value

    def GetVelocity(self, i, j, k, n, level):
        v = _d()
        _check(lib().lbx_sim_get_velocity(self._h, i, j, k, n, level, ctypes.byref(v)))
        return v.value

    def _level_dims(self, level):
        return tuple(d * 2 ** level for d in self.GetDims())

    def GetDensityField(self, level, out=None):
        """[NX_l, NY_l, NZ_l] array (C order); sentinel -1.0 where the level has no cell.
        `out`: optional preallocated (e.g. pinned) float64 array of that size."""
        if out is None:
            out = np.empty(self._level_dims(level))
        _check(lib().lbx_sim_get_density_field(self._h, level, out.ctypes.data_as(_dp), out.size))
        return out.reshape(self._level_dims(level))

    def GetVelocityField(self, level, out=None):
        if out is None:
            out = np.empty(self._level_dims(level) + (3,))
        _check(lib().lbx_sim_get_velocity_field(self._h, level, out.ctypes.data_as(_dp), out.size))
        return out.reshape(self._level_dims(level) + (3,))

    def GetLinearMomentField(self, level, weights, per_unit_density=False, sentinel=-3e8):
        """Generic derived variable: rows of `weights` ([ncomp, 15]) applied to the populations of
        every valid cell; dense [nx, ny, nz, ncomp] over the level's domain."""
        w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1, 15))
        dims = self._level_dims(level)
        out = np.empty(tuple(dims) + (w.shape[0],))
        _check(lib().lbx_sim_get_linear_moment_field(self._h, level, w.ctypes.data_as(_dp), w.shape[0], int(per_unit_density),
                                                     float(sentinel), out.ctypes.data_as(_dp), out.size))
        return out

    def WriteCheckpoint(self, path):
        _check(lib().lbx_sim_write_checkpoint(self._h, os.fsencode(path)))

    def ReadCheckpoint(self, path):
        _check(lib().lbx_sim_read_checkpoint(self._h, os.fsencode(path)))

    def GetTime(self, level):
        v = _d()
        _check(lib().lbx_sim_get_time(self._h, level, ctypes.byref(v)))
        return v.value

    def GetTimeStep(self, level):
        v = _i()
        _check(lib().lbx_sim_get_time_step(self._h, level, ctypes.byref(v)))
        return v.value

    def GetDims(self):
        d = _i3()
        _check(lib().lbx_sim_get_dims(self._h, d))
        return tuple(d)

    def GetExtent(self, level):
        lo, hi = _i3(), _i3()
        _check(lib().lbx_sim_get_extent(self._h, level, lo, hi))
        return tuple(lo), tuple(hi)

    def SetStaticRefinement(self, level, lo, hi):
        _check(lib().lbx_sim_set_static_refinement(self._h, level, _i3(*lo), _i3(*hi)))

    def UnsetStaticRefinement(self, level):
        _check(lib().lbx_sim_unset_static_refinement(self._h, level))

    def maxLevel(self):
        return lib().lbx_sim_max_level(self._h)

    def finestLevel(self):
        return lib().lbx_sim_finest_level(self._h)

    def refRatio(self, level):
        r = _i3()
        _check(lib().lbx_sim_ref_ratio(self._h, level, r))
        return tuple(r)

    def boxArray(self, level):
        n = lib().lbx_sim_num_boxes(self._h, level)
        if n < 0:
            raise LambrexError("level out of range")
        arr = (ctypes.c_int * max(6 * n, 1))()
        _check(lib().lbx_sim_get_boxes(self._h, level, arr))
        return _boxes_out(arr, n)

    # ---- white-box (tests/AmrTest.h)
    def FieldEmpty(self, level, field):
        r = lib().lbx_sim_field_empty(self._h, level, field)
        if r < 0:
            raise LambrexError(lib().lbx_sim_last_error().decode())
        return bool(r)

    def DensityEmpty(self, level):
        return self.FieldEmpty(level, DENSITY)

    def VelocityEmpty(self, level):
        return self.FieldEmpty(level, VELOCITY)

    def DistFnEmpty(self, level):
        return self.FieldEmpty(level, DISTFN)

    def FieldBoxes(self, level, field):
        n = lib().lbx_sim_field_num_boxes(self._h, level, field)
        if n < 0:
            raise LambrexError(lib().lbx_sim_last_error().decode())
        arr = (ctypes.c_int * max(6 * n, 1))()
        _check(lib().lbx_sim_field_boxes(self._h, level, field, arr))
        return _boxes_out(arr, n)

    def FieldFab(self, level, field, b, ngrow, ncomp):
        """[comp, z, y, x] over box b grown by the field's ghost width."""
        lo, hi = self.FieldBoxes(level, field)[b]
        n = ncomp * int(np.prod([h - l + 1 + 2 * ngrow for l, h in zip(lo, hi)]))
        out = np.empty(n)
        shape = (ctypes.c_int * 4)()
        _check(lib().lbx_sim_field_fab(self._h, level, field, b, out.ctypes.data_as(_dp), n, shape))
        return out.reshape(tuple(shape))

    def GetTauS(self, level):
        a, b = _d(), _d()
        _check(lib().lbx_sim_get_tau(self._h, level, ctypes.byref(a), ctypes.byref(b)))
        return a.value

    def GetTauB(self, level):
        a, b = _d(), _d()
        _check(lib().lbx_sim_get_tau(self._h, level, ctypes.byref(a), ctypes.byref(b)))
        return b.value

    def GetMass(self, level):
        v = _d()
        _check(lib().lbx_sim_get_mass(self._h, level, ctypes.byref(v)))
        return v.value

    def GetDt(self, level):
        v = _d()
        _check(lib().lbx_sim_get_dt(self._h, level, ctypes.byref(v)))
        return v.value

    def NumLevelsAllocated(self):
        return lib().lbx_sim_num_levels_allocated(self._h)

    def CallErrorEst(self, level, tag_boxes, preset=TAG_SET):
        """TagBoxArray(tag_boxes) preset on boxArray(level), then ErrorEst; returns one uint8
        array [z, y, x] per tag box."""
        arr, n = _boxes_in(tag_boxes)
        total = sum(int(np.prod([h - l + 1 for l, h in zip(lo, hi)])) for lo, hi in tag_boxes)
        buf = ctypes.create_string_buffer(max(total, 1))
        _check(lib().lbx_sim_call_error_est(self._h, level, arr, n, preset, buf, total))
        raw = np.frombuffer(buf.raw[:total], dtype=np.uint8)
        out, q = [], 0
        for lo, hi in tag_boxes:
            shp = tuple(h - l + 1 for l, h in zip(lo, hi))[::-1]
            m = int(np.prod(shp))
            out.append(raw[q:q + m].reshape(shp))
            q += m
        return out

    def CallMakeNewLevelFromScratch(self, level, boxes, time):
        arr, n = _boxes_in(boxes)
        _check(lib().lbx_sim_call_make_new_level_from_scratch(self._h, level, arr, n, time))

    def CallMakeNewLevelFromCoarse(self, level, boxes):
        arr, n = _boxes_in(boxes)
        _check(lib().lbx_sim_call_make_new_level_from_coarse(self._h, level, arr, n))

    def CallRemakeLevel(self, level, time, boxes):
        arr, n = _boxes_in(boxes)
        _check(lib().lbx_sim_call_remake_level(self._h, level, time, arr, n))

    def CallClearLevel(self, level):
        _check(lib().lbx_sim_call_clear_level(self._h, level))


# ---- grid-generation metadata (no GPU needed)
def _meta_call(fn, *args, cap=65536):
    out = (ctypes.c_int * (6 * cap))()
    n = fn(*args, out, cap)
    if n < 0:
        raise LambrexError(lib().lbx_sim_last_error().decode())
    return _boxes_out(out, n)


def meta_base_grids(dims, max_grid_size=32):
    return _meta_call(lib().lbx_meta_base_grids, _i3(*dims), max_grid_size)


def meta_max_size(boxes, chunk):
    arr, n = _boxes_in(boxes)
    return _meta_call(lib().lbx_meta_max_size, arr, n, chunk)


def meta_simplify(boxes):
    arr, n = _boxes_in(boxes)
    return _meta_call(lib().lbx_meta_simplify, arr, n)


def meta_complement(region, boxes):
    arr, n = _boxes_in(boxes)
    reg = (ctypes.c_int * 6)(*region[0], *region[1])
    return _meta_call(lib().lbx_meta_complement, reg, arr, n)


def metaParallelInit(group=None):
    """Distributed grid generation on the CPU (no GPU): MetaMesh objects tag the boxes this rank owns and
    merge the tag runs through torch.distributed -- the regrid path of lambrexInitParallel."""
    import torch.distributed as dist
    _check(lib().lbx_meta_parallel_init(dist.get_rank(group), dist.get_world_size(group), make_allgather_hook(group), None))


def metaParallelFinalise():
    _check(lib().lbx_meta_parallel_finalise())


def meta_distribution(boxes, nprocs, runs_per_rank=1):
    arr, n = _boxes_in(boxes)
    out = (_i * max(n, 1))()
    _check(lib().lbx_meta_distribution_runs(arr, n, int(nprocs), int(runs_per_rank), out))
    return [int(out[i]) for i in range(n)]


def meta_cluster(points, efficiency=0.7):
    p = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 3)
    return _meta_call(lib().lbx_meta_cluster, p.ctypes.data_as(_ip), len(p), efficiency)


class MetaMesh:
    """Field-less AmrCore with the reference's static-box tagging: grids only."""

    def __init__(self, dims, max_level, max_grid_size=32):
        h = _vp()
        _check(lib().lbx_meta_mesh_create(_i3(*dims), max_level, max_grid_size, ctypes.byref(h)))
        self._h = h.value

    def set_static(self, level, lo, hi):
        _check(lib().lbx_meta_mesh_set_static(self._h, level, _i3(*lo), _i3(*hi)))

    def unset_static(self, level):
        _check(lib().lbx_meta_mesh_unset_static(self._h, level))

    def finest_level(self):
        return lib().lbx_meta_mesh_finest_level(self._h)

    def boxes(self, level):
        return _meta_call(lib().lbx_meta_mesh_boxes, self._h, level)

    def log(self):
        buf = ctypes.create_string_buffer(1 << 16)
        lib().lbx_meta_mesh_log(self._h, buf, 1 << 16)
        return buf.value.decode().split()

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.lbx_meta_mesh_destroy(self._h)
        except Exception:
            pass
