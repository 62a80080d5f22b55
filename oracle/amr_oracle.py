"""ORACLE -- test infrastructure only, NOT product code.

CPU restatement (numpy) of LAMBReX's multi-level path on an explicit box hierarchy: the
reference's ``AmrSim`` members (/root/reference/src/AmrSim.cpp, cited per method) together
with the AMReX operations they call, restated from upstream AMReX semantics as recorded in
SURVEY.md appendix C.

PARITY UNPINNED for everything that depends on AMReX: AMReX is an un-vendored, un-pinned
dependency of the reference (CMakeLists.txt:28, ">= 19.08") that is neither under
/root/reference nor installed here, so the box generation (MakeBaseGrids, regrid /
Berger-Rigoutsos clustering), FillPatch*, sum_fine_to_coarse and makeFineMask below are
[AMReX, unverified].  What the reference itself pins is checked in
tests/test_amr_oracle.py: the single-level pulse golden vectors through the box-decomposed
pass structure, uniform-field interpolation, the dt/mass/tau ladder, tag sets and the
coverage inequalities of tests/catch2AMRTests.cpp.  The reference's behaviours that look like
defects are reproduced on purpose; freshly allocated fabs are
zero-filled (NEW_FAB_FILL, appendix B-4).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Conventions: a box is ((ilo,jlo,klo),(ihi,jhi,khi)) inclusive; fab arrays are
[comp, z, y, x] over the box grown by the MultiFab's ghost width.
"""
import numpy as np

from . import lbm_oracle as lo

NV, HALO, REF_RATIO = 15, 2, 2
COARSE_VAL, FINE_VAL = 0, 1          # include/AmrSim.h:69-70
NL_DENSITY, NL_VELOCITY = -1.0, -3e8  # include/AmrSim.h:35-36
TAG_CLEAR, TAG_BUF, TAG_SET = 0, 1, 2
NEW_FAB_FILL = 0.0
C = [(int(lo.CX[p]), int(lo.CY[p]), int(lo.CZ[p])) for p in range(NV)]


# ----------------------------------------------------------------------------- box calculus
def bx(lo_, hi_):
    return (tuple(int(v) for v in lo_), tuple(int(v) for v in hi_))


def ok(b):
    return all(h >= l for l, h in zip(*b))


def numpts(b):
    return int(np.prod([h - l + 1 for l, h in zip(*b)])) if ok(b) else 0


def grow(b, n):
    n = (n, n, n) if isinstance(n, int) else n
    return (tuple(l - g for l, g in zip(b[0], n)), tuple(h + g for h, g in zip(b[1], n)))


def shift(b, s):
    return (tuple(l + d for l, d in zip(b[0], s)), tuple(h + d for h, d in zip(b[1], s)))


def isect(a, b):
    r = (tuple(max(x, y) for x, y in zip(a[0], b[0])), tuple(min(x, y) for x, y in zip(a[1], b[1])))
    return r if ok(r) else None


def contains_pt(b, p):
    return all(l <= x <= h for l, x, h in zip(b[0], p, b[1]))


def contains_box(a, b):
    return contains_pt(a, b[0]) and contains_pt(a, b[1])


def coarsen(b, r):
    return (tuple(l // r for l in b[0]), tuple(h // r for h in b[1]))      # floor division


def refine(b, r):
    return (tuple(l * r for l in b[0]), tuple((h + 1) * r - 1 for h in b[1]))


def minimal_box(boxes):
    return (tuple(min(b[0][d] for b in boxes) for d in range(3)), tuple(max(b[1][d] for b in boxes) for d in range(3)))


def box_diff(b, cut):
    """b minus cut as disjoint boxes; per direction low slab then high slab (AMReX boxDiff)."""
    if isect(b, cut) is None:
        return [b]
    out, lo_, hi_ = [], list(b[0]), list(b[1])
    for d in range(3):
        if lo_[d] < cut[0][d] <= hi_[d]:
            h = list(hi_)
            h[d] = cut[0][d] - 1
            out.append(bx(lo_, h))
            lo_[d] = cut[0][d]
        if lo_[d] <= cut[1][d] < hi_[d]:
            l = list(lo_)
            l[d] = cut[1][d] + 1
            out.append(bx(l, hi_))
            hi_[d] = cut[1][d]
    return out


def complement_in(region, boxes):
    cur = [region]
    for cut in boxes:
        cur = [p for b in cur for p in box_diff(b, cut)]
        if not cur:
            break
    return cur


def simplify(boxes):
    boxes = list(boxes)
    changed = True
    while changed:
        changed = False
        for a in range(len(boxes)):
            for b in range(a + 1, len(boxes)):
                lo_, hi_, join, can = [0] * 3, [0] * 3, 0, True
                for d in range(3):
                    al, ah, bl, bh = boxes[a][0][d], boxes[a][1][d], boxes[b][0][d], boxes[b][1][d]
                    if al == bl and ah == bh:
                        lo_[d], hi_[d] = al, ah
                    elif al <= bl <= ah + 1:
                        lo_[d], hi_[d], join = al, max(ah, bh), join + 1
                    elif bl <= al <= bh + 1:
                        lo_[d], hi_[d], join = bl, max(ah, bh), join + 1
                    else:
                        can = False
                        break
                if can and join <= 1:
                    boxes[b] = bx(lo_, hi_)
                    del boxes[a]
                    changed = True
                    break
            if changed:
                break
    return boxes


def max_size(boxes, chunk):
    """BoxList::maxSize: per direction, pieces are chopped off the HIGH end of each box and
    appended after all boxes of that pass [AMReX, unverified]."""
    boxes = [bx(*b) for b in boxes]
    chunk = (chunk,) * 3 if isinstance(chunk, int) else chunk
    for d in range(3):
        chopped = []
        for n, b in enumerate(boxes):
            ln = b[1][d] - b[0][d] + 1
            if ln <= chunk[d]:
                continue
            ratio, bs, nlen = 1, chunk[d], ln
            while bs % 2 == 0 and nlen % 2 == 0:
                ratio, bs, nlen = ratio * 2, bs // 2, nlen // 2
            numblk = nlen // bs + (1 if nlen % bs else 0)
            sz, extra = nlen // numblk, nlen % numblk
            lo_, hi_ = list(b[0]), list(b[1])
            for k in range(numblk - 1):
                ksize = ((sz + 1) if k < extra else sz) * ratio
                pos = hi_[d] - ksize + 1
                l = list(lo_)
                l[d] = pos
                chopped.append(bx(l, hi_))
                hi_[d] = pos - 1
            boxes[n] = bx(lo_, hi_)
        boxes += chopped
    return boxes


def make_base_grids(domain, max_grid_size=32):
    """AmrMesh::MakeBaseGrids [AMReX, unverified]: coarsen by 2 where the extent is even,
    maxSize(max_grid_size / fac), refine back."""
    fac = tuple(2 if (domain[1][d] - domain[0][d] + 1) % 2 == 0 else 1 for d in range(3))
    cdom = (tuple(l // f for l, f in zip(domain[0], fac)), tuple(h // f for h, f in zip(domain[1], fac)))
    boxes = max_size([cdom], tuple(max_grid_size // f for f in fac))
    return [(tuple(l * f for l, f in zip(b[0], fac)), tuple((h + 1) * f - 1 for h, f in zip(b[1], fac))) for b in boxes]


def periodic_shifts(period):
    """Periodicity::shiftIntVect: all shifts including zero, i outermost [AMReX, unverified]."""
    r = [(-1, 0, 1) if p > 0 else (0,) for p in period]
    return [(i * period[0], j * period[1], k * period[2]) for i in r[0] for j in r[1] for k in r[2]]


# ----------------------------------------------------------------------------- MultiFab
class MultiFab:
    def __init__(self, boxes, ncomp, ng, dtype=np.float64, fill=NEW_FAB_FILL):
        self.boxes, self.ncomp, self.ng = [bx(*b) for b in boxes], ncomp, ng
        self.fabs = [np.full((ncomp,) + tuple(h - l + 1 + 2 * ng for l, h in zip(*b))[::-1], fill, dtype=dtype)
                     for b in self.boxes]

    def empty(self):
        return len(self.boxes) == 0

    def grown(self, i):
        return grow(self.boxes[i], self.ng)

    def view(self, i, region):
        """numpy view [comp, z, y, x] of `region` (global indices) inside fab i."""
        g = self.grown(i)
        assert contains_box(g, region), (g, region)
        s = tuple(slice(region[0][d] - g[0][d], region[1][d] - g[0][d] + 1) for d in (2, 1, 0))
        return self.fabs[i][(slice(None),) + s]

    def valid(self, i):
        return self.view(i, self.boxes[i])


def parallel_copy(dst, src, src_ng, dst_ng, period, add=False, ncomp=None):
    """FabArray::ParallelCopy [AMReX, unverified]: for each source box (ascending), each periodic
    shift, each destination box: dst(x) (=|+=) src(x - shift) on (src box grown + shift) n (dst box grown)."""
    nc = ncomp or dst.ncomp
    for i, sb in enumerate(src.boxes):
        sg = grow(sb, src_ng)
        for s in periodic_shifts(period):
            moved = shift(sg, s)
            for k, db in enumerate(dst.boxes):
                r = isect(moved, grow(db, dst_ng))
                if r is None:
                    continue
                dv = dst.view(k, r)
                sv = src.view(i, shift(r, tuple(-x for x in s)))
                if add:
                    dv[:nc] += sv[:nc]
                else:
                    dv[:nc] = sv[:nc]


def fill_boundary(mf, period):
    """FabArray::FillBoundary: every ghost cell <- the valid cell that covers it (incl. periodic
    images).  Valid cells are not written."""
    for k, db in enumerate(mf.boxes):
        gk = mf.grown(k)
        for i, sb in enumerate(mf.boxes):
            for s in periodic_shifts(period):
                if i == k and s == (0, 0, 0):
                    continue
                r = isect(shift(sb, s), gk)
                if r is None:
                    continue
                mf.view(k, r)[...] = mf.view(i, shift(r, tuple(-x for x in s)))


def fillpatch_single(dst, src, period):
    """FillPatchSingleLevel with one source at the same time (src/AmrSim.cpp:371)."""
    parallel_copy(dst, src, 0, dst.ng, period)


def pc_interp_fill(dst, crse, cperiod, ratio=REF_RATIO):
    """Every cell of dst (valid + ghosts): fine(x) = crse(floor(x / ratio)) where the coarse
    level (mod periodicity) has the cell (InterpFromCoarseLevel + PCInterp, src/AmrSim.cpp:406)."""
    for k in range(len(dst.boxes)):
        g = dst.grown(k)
        cg = coarsen(g, ratio)
        for i, cb in enumerate(crse.boxes):
            for s in periodic_shifts(cperiod):
                rc = isect(shift(cb, s), cg)
                if rc is None:
                    continue
                rf = isect(refine(rc, ratio), g)
                src = crse.view(i, shift(rc, tuple(-x for x in s)))
                big = src.repeat(ratio, axis=1).repeat(ratio, axis=2).repeat(ratio, axis=3)
                full = refine(rc, ratio)
                sl = tuple(slice(rf[0][d] - full[0][d], rf[1][d] - full[0][d] + 1) for d in (2, 1, 0))
                dst.view(k, rf)[...] = big[(slice(None),) + sl]


def fillpatch_two(dst, crse, fine, cperiod, fperiod):
    """FillPatchTwoLevels + PCInterp (src/AmrSim.cpp:385): cells not covered by fine valid data
    come from the coarse level by piecewise-constant interpolation; all others from fine."""
    pc_interp_fill(dst, crse, cperiod)
    fillpatch_single(dst, fine, fperiod)


def sum_fine_to_coarse(fine, crse, cperiod, ratio=REF_RATIO):
    """amrex::sum_fine_to_coarse (src/AmrSim.cpp:598) [AMReX, unverified]: average the fine cells
    (valid AND ghosts) onto coarsen(fine boxes) grown by ng/ratio, then ParallelCopy-ADD into
    the coarse valid cells with periodic wrap."""
    assert fine.ng % ratio == 0
    tmp = MultiFab([coarsen(b, ratio) for b in fine.boxes], fine.ncomp, fine.ng // ratio)
    for i in range(len(fine.boxes)):
        assert refine(tmp.grown(i), ratio) == fine.grown(i), "fine box not aligned to the coarse grid"
        f = fine.fabs[i]
        acc = np.zeros_like(tmp.fabs[i])
        for kr in range(ratio):            # amrex_avgdown: iref fastest
            for jr in range(ratio):
                for ir in range(ratio):
                    acc = acc + f[:, kr::ratio, jr::ratio, ir::ratio]
        tmp.fabs[i][...] = acc * (1.0 / ratio ** 3)
    parallel_copy(crse, tmp, tmp.ng, 0, cperiod, add=True)


def average_down(fine, crse, ratio=REF_RATIO):
    """amrex::average_down [AMReX, unverified]: every coarse VALID cell under fine VALID cells <- mean
    of its ratio^3 fine cells (amrex_avgdown order: iref fastest), overwriting; no ghost cells, no
    periodic images.  Used by the conventional subcycling driver, not by the
    reference's live path."""
    for i, fb in enumerate(fine.boxes):
        cb = coarsen(fb, ratio)
        assert refine(cb, ratio) == fb, "fine box not aligned to the coarse grid"
        f = fine.valid(i)
        acc = np.zeros((fine.ncomp,) + tuple(s // ratio for s in f.shape[1:]))
        for kr in range(ratio):
            for jr in range(ratio):
                for ir in range(ratio):
                    acc = acc + f[:, kr::ratio, jr::ratio, ir::ratio]
        acc = acc * (1.0 / ratio ** 3)
        for k, kb in enumerate(crse.boxes):
            r = isect(cb, kb)
            if r is None:
                continue
            sl = tuple(slice(r[0][d] - cb[0][d], r[1][d] - cb[0][d] + 1) for d in (2, 1, 0))
            crse.view(k, r)[...] = acc[(slice(None),) + sl]


def make_fine_mask(cmf_boxes, ng, fba, ratio=REF_RATIO):
    """amrex::makeFineMask(cmf, fba, ratio, crse, fine) [AMReX, unverified]: int mask on the coarse
    boxes (with cmf's ghosts), FINE_VAL on coarsen(fba) n grown fab box, no periodic images."""
    m = MultiFab(cmf_boxes, 1, ng, dtype=np.int32, fill=COARSE_VAL)
    cf = [coarsen(b, ratio) for b in fba]
    for k in range(len(m.boxes)):
        for c in cf:
            r = isect(c, m.grown(k))
            if r is not None:
                m.view(k, r)[...] = FINE_VAL
    return m


# ----------------------------------------------------------------------------- grid generation
def find_cut(hist):
    """Cluster FindCut [AMReX, unverified].  Returns (offset, status); status 0 hole, 1 steep,
    2 bisect, 3 invalid."""
    n = len(hist)
    if n <= 1:
        return 0, 3
    mid, cut, status = n // 2, -1, 3
    for i in range(n):
        if hist[i] == 0:
            status = 0
            if abs(cut - mid) > abs(i - mid):
                cut = i
                if i > mid:
                    break
    if status == 0:
        return cut, 0
    dh = [0] * n
    for i in range(1, n - 1):
        dh[i] = hist[i + 1] - 2 * hist[i] + hist[i - 1]
    locmax = -1
    for i in range(2, n - 2):
        ip, ic = dh[i - 1], dh[i]
        locdif = abs(ip - ic)
        if ip * ic < 0 and locdif >= locmax:
            if locdif > locmax:
                status, cut, locmax = 1, i, locdif
            elif abs(i - mid) < abs(cut - mid):
                cut = i
    if locmax <= 2:
        return mid, 2
    return cut, status


def cluster(points, eff):
    """ClusterList::chop(eff) [AMReX, unverified]: Berger-Rigoutsos on a point set.  Returns the
    clusters' minimal boxes in list order."""
    def minbox(p):
        return bx(p.min(axis=0), p.max(axis=0))

    lst = [points]
    i = 0
    while i < len(lst):
        p = lst[i]
        b = minbox(p)
        if len(p) / numpts(b) >= eff:
            i += 1
            continue
        cuts, stats = [], []
        for d in range(3):
            h = np.bincount(p[:, d] - b[0][d], minlength=b[1][d] - b[0][d] + 1)
            c, s = find_cut(list(h))
            cuts.append(b[0][d] + c)
            stats.append(s)
        mincut = min(stats)
        dirn = -1
        for d in range(3):
            if stats[d] == mincut and (dirn < 0 or (b[1][d] - b[0][d]) > (b[1][dirn] - b[0][dirn])):
                dirn = d
        low = p[:, dirn] < cuts[dirn]
        if low.all() or not low.any():      # cannot split further
            i += 1
            continue
        lst[i] = p[low]
        lst.append(p[~low])
    return [minbox(p) for p in lst], lst


# ----------------------------------------------------------------------------- the simulation
class Level:
    def __init__(self):
        self.clear()

    def clear(self):
        self.boxes = []
        self.now_f = self.next_f = self.now_rho = self.next_rho = self.vel = None
        self.time, self.delta, self.step = 0.0, 0.0, 0


class AmrSimOracle:
    """Mirror of /root/reference/include/AmrSim.h with the reference's method names in
    snake_case.  n_error_buf = 1, grid_eff = 0.7, n_proper = 1, blocking_factor = 1,
    ref_ratio = 2 (src/AmrSim.cpp:762-769 + AMReX defaults [unverified])."""

    def __init__(self, nx, ny, nz, max_level, tau_s, tau_b, max_grid_size=32, coracle=None):
        self.n = (nx, ny, nz)
        self.max_level, self.finest_level = max_level, 0
        self.max_grid_size, self.n_error_buf, self.grid_eff, self.n_proper = max_grid_size, 1, 0.7, 1
        self.levels = [Level() for _ in range(max_level + 1)]
        self.tau_s = [tau_s] + [0.0] * max_level
        self.tau_b = [tau_b] + [0.0] * max_level
        self.mass = [0.0] * (max_level + 1)
        self.static_tags = [None] * (max_level + 1)
        self.fine_masks = [None] * (max_level + 1)
        self.grids = [[] for _ in range(max_level + 1)]
        self.co = coracle or lo.COracle()
        self.initial_density = self.initial_velocity = None
        self.coupling = "rohde"      # or "subcycle": conventional driver, see subcycle_advance
        # dynamic refinement (not in the reference): gradient criterion per level
        # and AMReX-style regrid interval in coarse steps
        self.gradient_threshold = [0.0] * (max_level + 1)
        self.regrid_int, self.steps_since_regrid, self.num_regrids = 0, 0, 0

    # -- geometry
    def domain(self, level):
        r = REF_RATIO ** level
        return ((0, 0, 0), tuple(n * r - 1 for n in self.n))

    def period(self, level):
        return tuple(n * REF_RATIO ** level for n in self.n)

    # -- user input (src/AmrSim.cpp:804-822); arrays are C-ordered [i][j][k](,[n])
    def set_initial_density(self, rho):
        nx, ny, nz = self.n
        self.initial_density = (np.full(nx * ny * nz, float(rho)) if np.isscalar(rho)
                                else np.asarray(rho, dtype=np.float64).reshape(-1))

    def set_initial_velocity(self, u):
        nx, ny, nz = self.n
        self.initial_velocity = (np.full(3 * nx * ny * nz, float(u)) if np.isscalar(u)
                                 else np.asarray(u, dtype=np.float64).reshape(-1))

    # -- level construction
    def _define(self, level, boxes):
        L = self.levels[level]
        L.boxes = [bx(*b) for b in boxes]
        L.vel = MultiFab(boxes, 3, 0)
        L.now_f, L.next_f = MultiFab(boxes, NV, HALO), MultiFab(boxes, NV, HALO)
        L.now_rho, L.next_rho = MultiFab(boxes, 1, 0), MultiFab(boxes, 1, 0)

    def compute_dt(self, level):
        """src/AmrSim.cpp:297-322."""
        L = self.levels[level]
        if level:
            r = REF_RATIO
            L.delta = self.levels[level - 1].delta / r
            self.mass[level] = self.mass[level - 1] / r
            self.tau_s[level] = r * (self.tau_s[level - 1] - 0.5) + 0.5
            self.tau_b[level] = r * (self.tau_b[level - 1] - 0.5) + 0.5
        else:
            L.delta = 1.0
            self.mass[level] = 1.0

    def make_new_level_from_scratch(self, level, time, boxes):
        """src/AmrSim.cpp:665-687."""
        self._define(level, boxes)
        L = self.levels[level]
        L.time = time
        self.compute_dt(level)
        L.step = 0
        if level == 0:
            nx, ny, nz = self.n
            rho = self.initial_density.reshape(nx, ny, nz)
            u = self.initial_velocity.reshape(nx, ny, nz, 3)
            for i, b in enumerate(L.boxes):
                sl = tuple(slice(b[0][d], b[1][d] + 1) for d in range(3))
                L.now_rho.fabs[i][0] = rho[sl].transpose(2, 1, 0)
                L.vel.fabs[i][...] = u[sl].transpose(3, 2, 1, 0)
            self.calc_equilibrium_dist(level)

    def make_new_level_from_coarse(self, level, time, boxes):
        """src/AmrSim.cpp:689-713."""
        assert level > 0
        self._define(level, boxes)
        L = self.levels[level]
        L.time = time
        self.compute_dt(level)
        L.step = 0
        pc_interp_fill(L.now_f, self.levels[level - 1].now_f, self.period(level - 1))
        self.calc_hydro_vars(level)
        self.make_fine_mask(level - 1)

    def remake_level(self, level, time, boxes):
        """src/AmrSim.cpp:715-744."""
        L = self.levels[level]
        new_f = MultiFab(boxes, NV, HALO)
        self.dist_fn_fill_patch(level, new_f)
        L.boxes = [bx(*b) for b in boxes]
        L.now_f, L.now_rho, L.vel = new_f, MultiFab(boxes, 1, 0), MultiFab(boxes, 3, 0)
        # DEVIATION (documented in DESIGN.md): the reference leaves `next` on the OLD BoxArray
        # here, which breaks the next step whenever the grids really changed; we redefine it.
        L.next_f, L.next_rho = MultiFab(boxes, NV, HALO), MultiFab(boxes, 1, 0)
        L.time = time
        self.calc_hydro_vars(level)
        if level < self.finest_level:
            self.make_fine_mask(level)

    def clear_level(self, level):
        """src/AmrSim.cpp:746-751."""
        self.levels[level].clear()

    def make_fine_mask(self, coarse_level):
        """src/AmrSim.cpp:419-428 -- sic: the level's OWN boxes are passed as the fine BoxArray
        (SURVEY.md B-1)."""
        f = self.levels[coarse_level].now_f
        self.fine_masks[coarse_level] = make_fine_mask(f.boxes, f.ng, f.boxes)

    # -- fills
    def dist_fn_fill_patch(self, level, dest):
        """src/AmrSim.cpp:359-391."""
        if level == 0:
            fillpatch_single(dest, self.levels[0].now_f, self.period(0))
        else:
            fillpatch_two(dest, self.coarse_state_at(level - 1, self.levels[level].time),
                          self.levels[level].now_f, self.period(level - 1), self.period(level))

    def coarse_state_at(self, clev, t):
        """The coarse populations FillPatchTwoLevels interpolates from.  Reference coupling: NOW,
        whatever its time (src/AmrSim.cpp:373-389 passes one state).  "subcycle": the coarse level
        has already advanced, NOW is at t1 and NEXT holds the old state at t0 = t1 - dt (UpdateNow
        swapped them): state 0 / 1 when t is within 1e-3 dt of its time, else the linear
        combination (t1-t)/(t1-t0) old + (t-t0)/(t1-t0) new [AMReX FillPatchTwoLevels, unverified]."""
        Lc = self.levels[clev]
        if self.coupling != "subcycle":
            return Lc.now_f
        t1, dt = Lc.time, Lc.delta
        t0, eps = t1 - dt, 1e-3 * dt
        if abs(t - t1) < eps or Lc.step == 0:
            return Lc.now_f
        assert Lc.next_f.boxes == Lc.now_f.boxes
        if abs(t - t0) < eps:
            return Lc.next_f
        assert t0 - eps <= t <= t1 + eps
        a, b = (t1 - t) / (t1 - t0), (t - t0) / (t1 - t0)
        tmp = MultiFab(Lc.now_f.boxes, NV, HALO)
        for i in range(len(tmp.boxes)):
            tmp.valid(i)[...] = a * Lc.next_f.valid(i) + b * Lc.now_f.valid(i)
        return tmp

    def subcycle_advance(self, level):
        """Conventional subcycling (the corrected form of the reference's dead SubCycle,
        src/AmrSim.cpp:335-344): one step of `level`, `ratio` x the finer levels, average_down."""
        self.iterate_level(level)
        if level < self.finest_level:
            for _ in range(REF_RATIO):
                self.subcycle_advance(level + 1)
            average_down(self.levels[level + 1].now_f, self.levels[level].now_f)

    def update_boundaries(self, level):
        """src/AmrSim.cpp:19-23: FillBoundary on NEXT."""
        fill_boundary(self.levels[level].next_f, self.period(level))

    # -- cell physics on boxes
    def _collide_valid(self, mf, level, mask=None):
        ws, wb = 1.0 / (self.tau_s[level] + 0.5), 1.0 / (self.tau_b[level] + 0.5)
        for i in range(len(mf.boxes)):
            v = mf.valid(i)
            out = self.co.collide(np.ascontiguousarray(v), ws, wb)
            if mask is not None:
                out[:, mask.valid(i)[0] == FINE_VAL] = 0.0
            v[...] = out

    def stream(self, level):
        """src/AmrSim.cpp:109-122: pull into a FRESH fab over valid grown by 1, then swap."""
        L = self.levels[level]
        src = L.next_f
        prop = MultiFab(src.boxes, NV, HALO)
        for i, b in enumerate(src.boxes):
            r = grow(b, 1)
            for p in range(NV):
                prop.view(i, r)[p] = src.view(i, shift(r, tuple(-c for c in C[p])))[p]
        L.next_f = prop

    def calc_equilibrium_dist(self, level):
        """src/AmrSim.cpp:845-936 (valid cells of NOW, then FillBoundary of NEXT, sic B-6)."""
        L = self.levels[level]
        for i in range(len(L.boxes)):
            L.now_f.valid(i)[...] = self.co.equilibrium(np.ascontiguousarray(L.now_rho.fabs[i][0]),
                                                        np.ascontiguousarray(L.vel.fabs[i]))
        self.update_boundaries(level)

    def calc_hydro_vars(self, level):
        """src/AmrSim.cpp:938-979."""
        L = self.levels[level]
        for i in range(len(L.now_f.boxes)):
            r, u = self.co.moments(np.ascontiguousarray(L.now_f.valid(i)))
            L.now_rho.fabs[i][0], L.vel.fabs[i][...] = r, u

    # -- single-level step (src/AmrSim.cpp:124-135, 324-333; include/AmrSim.h:89-94)
    def iterate_level(self, level):
        L = self.levels[level]
        self.dist_fn_fill_patch(level, L.next_f)
        self._collide_valid(L.next_f, level)
        fill_boundary(L.next_f, self.period(level))
        self.stream(level)
        L.now_f, L.next_f = L.next_f, L.now_f            # UpdateNow swaps the whole State
        L.now_rho, L.next_rho = L.next_rho, L.now_rho
        L.time += L.delta
        L.step += 1

    # -- Rohde cycle (src/AmrSim.cpp:430-631)
    def init_post_collision(self, level):
        L = self.levels[level]
        self.dist_fn_fill_patch(level, L.next_f)
        if level != 0:
            for f in L.next_f.fabs:                       # comp 0 only, outermost ring (B-2)
                z = f[0]
                z[0, :, :] = z[-1, :, :] = 0.0
                z[:, 0, :] = z[:, -1, :] = 0.0
                z[:, :, 0] = z[:, :, -1] = 0.0

    def zero_invalid_components(self, level):
        """src/AmrSim.cpp:604-617."""
        mf = self.levels[level].next_f
        for i, b in enumerate(mf.boxes):
            g = mf.grown(i)
            f = mf.fabs[i]
            zz, yy, xx = np.meshgrid(np.arange(g[0][2], g[1][2] + 1), np.arange(g[0][1], g[1][1] + 1),
                                     np.arange(g[0][0], g[1][0] + 1), indexing="ij")

            def inside(x, y, z):
                return ((x >= b[0][0]) & (x <= b[1][0]) & (y >= b[0][1]) & (y <= b[1][1]) &
                        (z >= b[0][2]) & (z <= b[1][2]))
            shell = ~inside(xx, yy, zz)
            for m in range(NV):
                kill = shell & ~inside(xx - 2 * C[m][0], yy - 2 * C[m][1], zz - 2 * C[m][2])
                f[m][kill] = 0.0

    def update_distribution(self, level):
        L = self.levels[level]
        if level and level == self.finest_level:
            L.time += 2 * L.delta
            L.step += 2
        else:
            L.time += L.delta
            L.step += 1
        L.now_f, L.next_f = L.next_f, L.now_f

    def rohde_cycle(self, cl):
        self.init_post_collision(cl)
        self._collide_valid(self.levels[cl].next_f, cl, mask=self.fine_masks[cl])      # CoarseCollide
        if cl + 1 == self.finest_level:
            fl = self.finest_level
            self.init_post_collision(fl)
            self._collide_valid(self.levels[fl].next_f, fl)
            self.stream(fl)
            self._collide_valid(self.levels[fl].next_f, fl)
            self.stream(fl)
            self.zero_invalid_components(fl)
            self.update_distribution(fl)
        else:
            for _ in range(REF_RATIO):
                self.rohde_cycle(cl + 1)
        self.stream(cl)
        sum_fine_to_coarse(self.levels[cl + 1].now_f, self.levels[cl].next_f, self.period(cl))
        self.zero_invalid_components(cl)
        if cl == 0:
            self.update_boundaries(0)
        self.update_distribution(cl)

    def iterate(self, nsteps):
        """src/AmrSim.cpp:981-993."""
        for _ in range(nsteps):
            if self.finest_level == 0:
                self.iterate_level(0)
            elif self.coupling == "subcycle":
                self.subcycle_advance(0)
            else:
                self.rohde_cycle(0)
            if self.regrid_int > 0 and self.max_level > 0:
                self.steps_since_regrid += 1
                if self.steps_since_regrid >= self.regrid_int:
                    self.steps_since_regrid = 0
                    self.num_regrids += 1
                    self.regrid(0, self.levels[0].time)
                    for l in range(self.finest_level):
                        self.make_fine_mask(l)

    # -- tagging and regrid
    def error_est(self, level, tags):
        """src/AmrSim.cpp:633-663: valid cells SET inside static_tags[level], CLEAR elsewhere.
        tags: list of uint8 arrays [z,y,x] over the level's boxes grown by ng (tags_ng)."""
        ng = tags["ng"]
        st = self.static_tags[level]
        for i, b in enumerate(self.levels[level].now_f.boxes):
            t = tags["fabs"][i]
            core = t[ng:t.shape[0] - ng, ng:t.shape[1] - ng, ng:t.shape[2] - ng]
            core[...] = TAG_CLEAR
            r = isect(b, st) if st is not None else None
            if r is not None:
                core[r[0][2] - b[0][2]:r[1][2] - b[0][2] + 1, r[0][1] - b[0][1]:r[1][1] - b[0][1] + 1,
                     r[0][0] - b[0][0]:r[1][0] - b[0][0] + 1] = TAG_SET
        if self.gradient_threshold[level] > 0.0:
            self._gradient_tags(level, tags)

    def _gradient_tags(self, level, tags):
        """Gradient criterion (addition): rho of NOW with one ghost cell, filled like DistFnFillPatch
        fills populations (same level + periodic images, else PC from the coarse level); a valid cell
        is SET when 0.25 * ((rho(x+ex)-rho(x-ex))^2 + (..y..)^2 + (..z..)^2) > threshold^2, every
        operation rounded separately in this order."""
        ng = tags["ng"]
        L = self.levels[level]
        self.calc_hydro_vars(level)
        rho_g = MultiFab(L.now_rho.boxes, 1, 1)
        if level == 0:
            parallel_copy(rho_g, L.now_rho, 0, 1, self.period(0))
        else:
            self.calc_hydro_vars(level - 1)
            pc_interp_fill(rho_g, self.levels[level - 1].now_rho, self.period(level - 1))
            parallel_copy(rho_g, L.now_rho, 0, 1, self.period(level))
        thr2 = self.gradient_threshold[level] * self.gradient_threshold[level]
        for i in range(len(L.now_rho.boxes)):
            r = rho_g.fabs[i][0]
            gx = r[1:-1, 1:-1, 2:] - r[1:-1, 1:-1, :-2]
            gy = r[1:-1, 2:, 1:-1] - r[1:-1, :-2, 1:-1]
            gz = r[2:, 1:-1, 1:-1] - r[:-2, 1:-1, 1:-1]
            g2 = 0.25 * ((gx * gx + gy * gy) + gz * gz)
            t = tags["fabs"][i]
            core = t[ng:t.shape[0] - ng, ng:t.shape[1] - ng, ng:t.shape[2] - ng]
            core[g2 > thr2] = TAG_SET

    def _tag_points(self, levc, extra_boxes):
        """Tagged cells of level levc after ErrorEst, projection of finer new grids, buffering and periodic
        mapping: sorted unique points inside the level's domain.  Order as in AmrMesh::MakeNewGrids [AMReX,
        unverified]: the tag boxes are allocated with n_error_buf + ngrow ghost cells (ngrow = how far the
        level's grids must grow to contain the projection), the projection is SET BEFORE buffering, and the
        buffer width is n_error_buf + ngrow; TagBox::buffer spreads only SET cells of the VALID region."""
        boxes = self.grids[levc]
        ngrow = 0
        if extra_boxes:
            def covered(g):
                grown = [grow(b, g) for b in boxes]
                return all(not complement_in(e, grown) for e in extra_boxes)
            while ngrow < 64 and not covered(ngrow):
                ngrow += 1
        nb = self.n_error_buf + ngrow
        tags = {"ng": nb, "fabs": [np.zeros(tuple(h - l + 1 + 2 * nb for l, h in zip(*b))[::-1], dtype=np.uint8)
                                   for b in boxes]}
        self.error_est(levc, tags)
        pts = []
        per = self.period(levc)
        for b, t in zip(boxes, tags["fabs"]):
            for e in extra_boxes:                  # proper-nesting projection of finer new grids: SET
                r = isect(e, grow(b, nb))
                if r is not None:
                    o = tuple(b[0][d] - nb for d in range(3))
                    t[r[0][2] - o[2]:r[1][2] - o[2] + 1, r[0][1] - o[1]:r[1][1] - o[1] + 1,
                      r[0][0] - o[0]:r[1][0] - o[0] + 1] = TAG_SET
            # TagBox::buffer: SET cells of the valid region spread BUF over +-nb
            setm = np.zeros(t.shape, dtype=bool)
            if nb:
                setm[nb:-nb, nb:-nb, nb:-nb] = (t[nb:-nb, nb:-nb, nb:-nb] == TAG_SET)
            if nb and setm.any():
                grown = setm.copy()
                for dz in range(-nb, nb + 1):
                    for dy in range(-nb, nb + 1):
                        for dx in range(-nb, nb + 1):
                            grown |= np.roll(setm, (dz, dy, dx), axis=(0, 1, 2))   # valid cells are >= nb from the rim
                t[grown & (t == TAG_CLEAR)] = TAG_BUF
            z, y, x = np.nonzero(t)
            if len(x):
                pts.append(np.stack([x + b[0][0] - nb, y + b[0][1] - nb, z + b[0][2] - nb], axis=1))
        if not pts:
            return np.zeros((0, 3), dtype=np.int64)
        p = np.concatenate(pts).astype(np.int64)
        p %= np.array(per, dtype=np.int64)          # mapPeriodic: everything lands inside the domain
        p = np.unique(p, axis=0)
        return p[np.lexsort((p[:, 0], p[:, 1], p[:, 2]))]   # IntVect order: z slowest

    def make_new_grids(self, lbase):
        """AmrMesh::MakeNewGrids(lbase, time, new_finest, new_grids) [AMReX, unverified]."""
        max_crse = min(self.finest_level, self.max_level - 1)
        new_grids = {l: list(self.grids[l]) for l in range(lbase + 1)}
        # proper nesting domains
        p_n_comp, p_n = {}, {}
        dom = self.domain(lbase)
        comp = simplify(complement_in(dom, simplify(self.grids[lbase])))
        comp = [grow(b, self.n_proper) for b in comp]
        comp = self._proj_periodic(comp, lbase)
        p_n_comp[lbase] = comp
        p_n[lbase] = simplify(complement_in(dom, comp))
        for i in range(lbase + 1, max_crse + 1):
            c = [grow(refine(b, REF_RATIO), self.n_proper) for b in simplify(p_n_comp[i - 1])]
            c = self._proj_periodic(c, i)
            p_n_comp[i] = c
            p_n[i] = simplify(complement_in(self.domain(i), c))
        new_finest = lbase
        for levc in range(max_crse, lbase - 1, -1):
            levf = levc + 1
            extra = []
            if levf < new_finest:
                extra = [coarsen(grow(coarsen(b, REF_RATIO), self.n_proper), REF_RATIO) for b in new_grids[levf + 1]]
            pts = self._tag_points(levc, extra)
            if len(pts) and p_n_comp[levc]:
                keep = np.ones(len(pts), dtype=bool)
                for b in p_n_comp[levc]:
                    keep &= ~np.all((pts >= np.array(b[0])) & (pts <= np.array(b[1])), axis=1)
                pts = pts[keep]
            if len(pts) == 0:
                continue
            new_finest = max(new_finest, levf)
            cboxes, _ = cluster(pts, self.grid_eff)
            clipped = []
            for b in cboxes:                         # ClusterList::intersect(p_n)
                if any(contains_box(q, b) for q in p_n[levc]):
                    clipped.append(b)
                else:
                    clipped += [r for r in (isect(b, q) for q in p_n[levc]) if r is not None]
            nb = simplify(clipped)
            nb = max_size(nb, self.max_grid_size // REF_RATIO)
            nb = [refine(b, REF_RATIO) for b in nb]
            new_grids[levf] = nb
        return new_finest, new_grids

    def _proj_periodic(self, boxes, level):
        """ProjPeriodic: add the periodic images that intersect the domain [AMReX, unverified]."""
        dom = self.domain(level)
        out = list(boxes)
        for b in boxes:
            for s in periodic_shifts(self.period(level)):
                if s == (0, 0, 0):
                    continue
                r = isect(shift(b, s), dom)
                if r is not None:
                    out.append(r)
        return out

    def init_from_scratch(self, time=0.0):
        """AmrCore::InitFromScratch -> AmrMesh::MakeNewGrids(time) [AMReX, unverified]."""
        self.finest_level = 0
        self.grids[0] = make_base_grids(self.domain(0), self.max_grid_size)
        self.make_new_level_from_scratch(0, time, self.grids[0])
        while self.finest_level < self.max_level:
            new_finest, new_grids = self.make_new_grids(self.finest_level)
            if new_finest <= self.finest_level:
                break
            self.finest_level = new_finest
            self.make_new_level_from_scratch(new_finest, time, new_grids[new_finest])
            self.grids[new_finest] = new_grids[new_finest]

    def regrid(self, lbase, time):
        """AmrCore::regrid [AMReX, unverified]."""
        if lbase >= self.max_level:
            return
        new_finest, new_grids = self.make_new_grids(lbase)
        coarse_changed = False
        for lev in range(lbase + 1, new_finest + 1):
            if lev <= self.finest_level:
                changed = new_grids[lev] != self.grids[lev]
                if changed or coarse_changed:
                    g = new_grids[lev] if changed else self.grids[lev]
                    self.remake_level(lev, time, g)
                    self.grids[lev] = g
                coarse_changed = changed
            else:
                self.make_new_level_from_coarse(lev, time, new_grids[lev])
                self.grids[lev] = new_grids[lev]
        for lev in range(new_finest + 1, self.finest_level + 1):
            self.clear_level(lev)
            self.grids[lev] = []
        self.finest_level = new_finest

    def set_static_refinement(self, level, lo_corner, hi_corner):
        """src/AmrSim.cpp:995-1009."""
        self.static_tags[level] = bx(lo_corner, hi_corner)
        self.regrid(level, self.levels[level].time)
        self.make_fine_mask(level)

    def unset_static_refinement(self, level):
        """src/AmrSim.cpp:1011-1017."""
        self.static_tags[level] = None
        self.regrid(level, self.levels[level].time)
        self.make_fine_mask(level)

    def set_gradient_refinement(self, level, threshold):
        assert threshold > 0.0
        self.gradient_threshold[level] = threshold
        self.regrid(level, self.levels[level].time)
        self.make_fine_mask(level)

    def unset_gradient_refinement(self, level):
        self.gradient_threshold[level] = 0.0
        self.regrid(level, self.levels[level].time)
        self.make_fine_mask(level)

    # -- output (src/AmrSim.cpp:824-843)
    def get_density(self, i, j, k, level):
        mf = self.levels[level].now_rho
        if mf is None:
            return NL_DENSITY
        for n, b in enumerate(mf.boxes):
            if contains_pt(b, (i, j, k)):
                return float(mf.fabs[n][0, k - b[0][2], j - b[0][1], i - b[0][0]])
        return NL_DENSITY

    def get_velocity(self, i, j, k, n, level):
        mf = self.levels[level].vel
        if mf is None:
            return NL_VELOCITY
        for q, b in enumerate(mf.boxes):
            if contains_pt(b, (i, j, k)):
                return float(mf.fabs[q][n, k - b[0][2], j - b[0][1], i - b[0][0]])
        return NL_VELOCITY

    def gather_valid(self, level, what="f"):
        """Dense array over the level's domain ([comp, z, y, x]); cells the level does not own
        are NaN.  what: 'f' | 'rho' | 'u'."""
        L = self.levels[level]
        mf = {"f": L.now_f, "rho": L.now_rho, "u": L.vel}[what]
        d = self.domain(level)
        out = np.full((mf.ncomp,) + tuple(h + 1 for h in d[1])[::-1], np.nan)
        for i, b in enumerate(mf.boxes):
            out[:, b[0][2]:b[1][2] + 1, b[0][1]:b[1][1] + 1, b[0][0]:b[1][0] + 1] = mf.valid(i)
        return out
