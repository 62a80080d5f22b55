// Row-owner form of the fused per-level collide + Stream (round 2; replaces the tile form k_mf_collide_stream
// on every path that has a ghost source).  Reference passes replaced per launch: InitPostCollision /
// DistFnFillPatch + CoarseCollide / FineCollide + Stream + ZeroInvalidComponents
// (/root/reference/src/AmrSim.cpp:471-617, 109-122) or, as the conventional level step, CollideLevel + Stream
// (:124-135, 109-122).
//
// Why: a boxed destination carries 2 ghost cells per side.  In the tile form valid cells were pushed by one set
// of CTAs and ghost cells by others, later in launch order, so every 32-byte sector at a row end was written
// twice, by different CTAs, microseconds to milliseconds apart -- the round-1 micro-benchmark
// (profiles/r01_sector_bench.log) measured that pattern at -46 %.  Here ONE WARP owns one source row of the GROWN
// box (36 cells for a 32^3 box: 2 ghost + 32 valid + 2 ghost) and produces, for every population p, the WHOLE
// destination row (j + c_y, k + c_z) of plane p in one aligned store instruction:
//   phase 1  the warp gathers its row's 15 post-collision populations into a shared-memory row buffer
//            [15][n0]: valid cells are read from the valid source and collided (masked cells -> 0); ghost cells
//            take their values uncollided from the ghost source -- the fab's own ghost cells, or the cell a
//            FillPatch plan resolves them to (same level / periodic image / coarse cell; a per-ghost-cell table
//            built once per plan by k_plan_resolve, no descriptor search in the step), loading only the
//            populations that have a destination;
//   phase 2  for each p the warp writes destination row (j + c_y, k + c_z), plane p: cell x takes buffer[p][x - c_x]
//            (the x shift happens in shared memory, so global stores stay sector-aligned: 18 lanes x 16 B for a
//            36-cell row), ring-2 cells and rows take 0 (the reference's fresh fab), and
//            ZeroInvalidComponents (:604-617) is a filter on the stored value.  Ring-2 rows of planes whose
//            natural writer (j - c_y, k - c_z) lies outside the grown box are zeroed by the row itself.
// Every (cell, population) of the destination is written exactly once, by exactly one warp, as part of a whole
// row: no partial sectors, no second writer.  Rows need an even allocated length and the tight layout (valid + 2
// ghosts per side); other fab sets use the tile kernel.
#pragma once
#include "kernels.cuh"

namespace lbx {

#ifndef LBX_RO_THREADS
#define LBX_RO_THREADS 128
#endif
constexpr int RO_THREADS = LBX_RO_THREADS;      // threads per CTA = 32 x source rows per CTA (128 x 6 CTAs/SM measured best: profiles/r02_row_kernel.md)

struct ROPlan {                // FillPatch sources resolved per ghost cell
  const int2* tab;             // x = kind | src_fab << 2 (kind: 0 none, 1 same level COPY, 2 coarse PC), y = source cell offset
  const long long* tab_first;  // [nfabs] first table entry of each fab (0 for fabs of other ranks)
  const DFabT* s0;             // same-level source set
  const DFabT* s1;             // coarse source set
  const DFabT* s1b;            // second coarse state (level step, time interpolation) or null
  const DFabT* fb;             // fallback: the fab FillPatch would have filled, or null (0)
  double wa, wb;
};

struct ROArgs {
  const double* vbase;         // valid-cell source allocation (same geometry as the destination)
  double* dbase;               // destination allocation
  const DFabT* dt;             // destination boxes
  const DFabT* mt;             // fine mask (int32, same boxes) or null
  const DFabT* gt;             // MODE 1: the set whose own ghost cells are pushed
  ROPlan plan;                 // MODE 2, 3
  int nfabs, warps, pitch;     // warps per CTA, shared-memory row pitch in doubles
  int fab0;                    // first fab of this launch (grid.z covers at most 65535 fabs)
  double omega_s, omega_b;
  int fine_val, zero_invalid;
};

// shell enumeration of a tight fab (n0 x n1 x n2 grown cells, 2 ghost cells per side): z slabs, y slabs, x cells;
// whole ghost rows are contiguous, the 4 x-ghost cells of a valid row are contiguous
__host__ __device__ inline long long ro_shell_size(int n0, int n1, int n2) {
  return 4ll * n0 * n1 + 4ll * n0 * (n2 - 4) + 4ll * (n1 - 4) * (n2 - 4);
}
__device__ __forceinline__ long long ro_shell_index(int n0, int n1, int n2, int x, int jr, int kr) {
  const long long A = 2ll * n0 * n1, B = 2ll * n0 * (n2 - 4);
  if (kr < 2) return ((long long)kr * n1 + jr) * n0 + x;
  if (kr >= n2 - 2) return A + ((long long)(kr - (n2 - 2)) * n1 + jr) * n0 + x;
  if (jr < 2) return 2 * A + ((long long)(kr - 2) * 2 + jr) * n0 + x;
  if (jr >= n1 - 2) return 2 * A + B + ((long long)(kr - 2) * 2 + (jr - (n1 - 2))) * n0 + x;
  return 2 * A + 2 * B + ((long long)(kr - 2) * (n1 - 4) + (jr - 2)) * 4 + (x < 2 ? x : x - (n0 - 4));
}
// inverse: entry t -> (x, jr, kr) relative to the grown box (a fab holds < 2^31 cells: 32-bit divisions)
__device__ __forceinline__ void ro_shell_cell(int n0, int n1, int n2, long long tt, int& x, int& jr, int& kr) {
  const unsigned A = 2u * n0 * n1, B = 2u * n0 * (n2 - 4), un0 = n0, un1 = n1;
  unsigned t = (unsigned)tt;
  if (t < 2 * A) {
    const unsigned s = t >= A;
    const unsigned r = t - s * A, row = r / un0;
    x = (int)(r - row * un0);
    const unsigned pl = row / un1;
    jr = (int)(row - pl * un1);
    kr = (s ? n2 - 2 : 0) + (int)pl;
    return;
  }
  t -= 2 * A;
  if (t < 2 * B) {
    const unsigned s = t >= B;
    const unsigned r = t - s * B, row = r / un0;
    x = (int)(r - row * un0);
    jr = (s ? n1 - 2 : 0) + (int)(row & 1u);
    kr = 2 + (int)(row >> 1);
    return;
  }
  t -= 2 * B;
  const unsigned q = t & 3u, row = t >> 2, pl = row / (un1 - 4);
  x = q < 2 ? (int)q : n0 - 4 + (int)q;
  jr = 2 + (int)(row - pl * (un1 - 4));
  kr = 2 + (int)pl;
}

// One launch per plan and geometry: the backward descriptor search of the tile kernel, done ONCE per ghost cell
// instead of once per ghost cell per time step.  grid = (tiles over the largest shell, fab).
static __global__ void __launch_bounds__(MFT) k_plan_resolve(const DFabT* __restrict__ dt, int nfabs, const GDst* __restrict__ dsts,
                                                      const int* __restrict__ fab_first, const GDesc* __restrict__ descs,
                                                      const DFabT* __restrict__ s0, const DFabT* __restrict__ s1,
                                                      int2* __restrict__ tab, const long long* __restrict__ tab_first) {
  const int b = mf_local_fab(dt, nfabs);
  if (b < 0) return;
  const DFabT D = dt[b];
  if (!D.local) return;
  const int n0 = D.n[0], n1 = D.n[1], n2 = D.n[2];
  const long long t = (long long)blockIdx.x * MFT + threadIdx.x;
  if (t >= ro_shell_size(n0, n1, n2)) return;
  int x, jr, kr;
  ro_shell_cell(n0, n1, n2, t, x, jr, kr);
  const int i = D.lo[0] + x, j = D.lo[1] + jr, k = D.lo[2] + kr;
  int2 e = make_int2(0, 0);
  bool done = false;
  for (int q = fab_first[b]; q < fab_first[b + 1] && !done; ++q) {
    const GDst G = dsts[q];
    if (i < G.blo[0] || i > G.bhi[0] || j < G.blo[1] || j > G.bhi[1] || k < G.blo[2] || k > G.bhi[2]) continue;
    for (int d = G.count - 1; d >= 0; --d) {                 // the last matching descriptor wins
      const GDesc g = descs[G.first + d];
      if (i < g.lo[0] || i > g.hi[0] || j < g.lo[1] || j > g.hi[1] || k < g.lo[2] || k > g.hi[2]) continue;
      if (g.kind == G_COPY || g.kind == G_PC) {
        const DFabT S = (g.src_set ? s1 : s0)[g.src_fab];
        int si = i, sj = j, sk = k;
        if (g.kind == G_PC) { si = fdiv(i, g.ratio); sj = fdiv(j, g.ratio); sk = fdiv(k, g.ratio); }
        e.x = (g.kind == G_COPY ? 1 : 2) | (g.src_fab << 2);
        e.y = (int)mf_off(S, si + g.shift[0], sj + g.shift[1], sk + g.shift[2]);
      }
      done = true;                                            // a G_NONE hit ends the search too: no source
      break;
    }
  }
  tab[tab_first[b] + t] = e;
}

// FillBoundary through the resolved table: one thread per ghost cell of the shell, NC components each (0: run-time
// count).  The table already says which fab and element the cell copies, so the step is 8 B of table + 2 x 8 B x
// ncomp of data per ghost cell, and whole ghost rows are read and written as contiguous runs.  Plans whose every
// descriptor is a same-set COPY into the ghost shell qualify (lbx_plan_apply checks); entries with no source keep
// their value, as the descriptor kernel leaves cells no region covers.
// The x-ghost cells of a valid row share a 32-byte sector with two valid cells (rows are even and start on a sector
// boundary): one thread per such sector reads its four cells, replaces the ghost pair and stores the WHOLE sector,
// so that L2 never holds a half-written sector it must complete from DRAM (profiles/r02_shell_copy.md).  The valid
// pair is stored back unchanged: a peer that reads it meanwhile sees the same bits.
template <int NC>
__global__ void __launch_bounds__(MFT) k_shell_copy(const DFabT* __restrict__ dt, int nfabs, const DFabT* __restrict__ st,
                                                    const int2* __restrict__ tab, const long long* __restrict__ tab_first,
                                                    int ncomp) {
  const int b = mf_local_fab(dt, nfabs);
  if (b < 0) return;
  const DFabT D = dt[b];
  if (!D.local) return;
  const int n0 = D.n[0], n1 = D.n[1], n2 = D.n[2];
  const unsigned t = blockIdx.x * MFT + threadIdx.x;
  const unsigned base3 = 4u * n0 * n1 + 4u * n0 * (n2 - 4);     // first x-ghost entry
  const long long dsc = mf_stride(D);
  const int nc = NC > 0 ? NC : ncomp;
  if (t >= base3) {
    const unsigned u = t - base3, rows = (unsigned)(n1 - 4) * (n2 - 4);
    if (u >= 2 * rows) return;
    const unsigned row = u >> 1, side = u & 1u, pl = row / (unsigned)(n1 - 4);
    const int jr = 2 + (int)(row - pl * (n1 - 4)), kr = 2 + (int)pl;
    const int4 ee = __ldg(reinterpret_cast<const int4*>(tab + tab_first[b] + base3 + 4 * row + 2 * side));
    double* dp = static_cast<double*>(D.p) + ((side ? n0 - 4 : 0) + (long long)n0 * (jr + (long long)n1 * kr));
    const int go = side ? 2 : 0, vo = 2 - go;                   // ghost pair / valid pair inside the sector
    const bool h0 = (ee.x & 3) != 0, h1 = (ee.z & 3) != 0;
    if (!h0 && !h1) return;
    const DFabT S0 = st[h0 ? ee.x >> 2 : ee.z >> 2], S1 = st[h1 ? ee.z >> 2 : ee.x >> 2];
    const double* s0 = static_cast<const double*>(S0.p) + ee.y;
    const double* s1 = static_cast<const double*>(S1.p) + ee.w;
    const long long sc0 = mf_stride(S0), sc1 = mf_stride(S1);
    const bool pair = h0 && h1 && ee.x == ee.z && ee.w == ee.y + 1 && !(ee.y & 1);
    constexpr int CH = 5;                                       // loads of a chunk in flight together, then its stores
    for (int c0 = 0; c0 < nc; c0 += CH) {
      double2 gv[CH], vv[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int c = min(c0 + q, nc - 1);
        vv[q] = *reinterpret_cast<const double2*>(dp + c * dsc + vo);
        if (pair) gv[q] = *reinterpret_cast<const double2*>(s0 + c * sc0);
        else {
          gv[q] = *reinterpret_cast<const double2*>(dp + c * dsc + go);
          if (h0) gv[q].x = s0[c * sc0];
          if (h1) gv[q].y = s1[c * sc1];
        }
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        if (c0 + q >= nc) break;
        const int c = c0 + q;
        *reinterpret_cast<double2*>(dp + c * dsc + go) = gv[q];
        *reinterpret_cast<double2*>(dp + c * dsc + vo) = vv[q];
      }
    }
    return;
  }
  const int2 e = __ldg(tab + tab_first[b] + t);
  if ((e.x & 3) == 0) return;
  int x, jr, kr;
  ro_shell_cell(n0, n1, n2, t, x, jr, kr);
  const DFabT S = st[e.x >> 2];
  const long long ssc = mf_stride(S);
  const double* sp = static_cast<const double*>(S.p) + e.y;
  double* dp = static_cast<double*>(D.p) + (x + (long long)n0 * (jr + (long long)n1 * kr));
  if (NC > 0) {
    double v[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c) v[c] = sp[c * ssc];
#pragma unroll
    for (int c = 0; c < NC; ++c) dp[c * dsc] = v[c];
  } else {
    for (int c = 0; c < ncomp; ++c) dp[c * dsc] = sp[c * ssc];
  }
}

// ---- separable population masks: bit p of dir_mask(axis, c) is set iff component `axis` of c_p equals c ------------
__host__ __device__ constexpr unsigned dir_mask(int axis, int c) {
  unsigned m = 0;
  for (int p = 0; p < NV; ++p)
    if (cc(axis, p) == c) m |= 1u << p;
  return m;
}
// populations whose step along `axis` from coordinate v lands in [lo, n - lo): one mask from three range tests
// (the 15 populations factorise by direction, so "destination inside a box" is an AND of three such masks)
__device__ __forceinline__ unsigned axis_sel(int axis_is, int v, int n, int lo, bool backwards, int step = 1) {
  const unsigned M = axis_is == 0 ? dir_mask(0, -1) : axis_is == 1 ? dir_mask(1, -1) : dir_mask(2, -1);
  const unsigned Z = axis_is == 0 ? dir_mask(0, 0) : axis_is == 1 ? dir_mask(1, 0) : dir_mask(2, 0);
  const unsigned P = axis_is == 0 ? dir_mask(0, 1) : axis_is == 1 ? dir_mask(1, 1) : dir_mask(2, 1);
  const int vm = backwards ? v + step : v - step, vp = backwards ? v - step : v + step;   // where a c = -1 / +1 population lands
  unsigned m = 0;
  if (vm >= lo && vm < n - lo) m |= M;
  if (v >= lo && v < n - lo) m |= Z;
  if (vp >= lo && vp < n - lo) m |= P;
  return m;
}

// MODE 1: own ghost cells; 2: FillPatch plan (Rohde pair); 3: conventional level step.  ZI: ZeroInvalidComponents folded in.
// Fabs hold fewer than 2^31 / 15 cells (checked on the host): element offsets inside a fab are 32-bit.
// Shared-memory row buffer of a warp, stored ALREADY SHIFTED along x: population p of source cell x sits at
// sm[p * pitch + x + c_x(p)], i.e. at the x of the cell it streams to, so that phase 2 reads one aligned 16-byte pair
// per population and lane (read unshifted, the two 8-byte loads of a lane hit every other bank pair and the
// shared-memory queue became the limiter on long rows).  sm[p * pitch - 1] and sm[p * pitch + n0] take the values that
// leave the row; the slot nobody writes (x = 0 for c_x = +1, x = n0 - 1 for c_x = -1) is zeroed, so that the ring-2
// destination cells x = 0 and x = n0 - 1 come out as 0 without a test: they read that slot, or a ghost value phase 1 has
// already zeroed because it has no destination.
#ifndef LBX_RO_MIN_CTAS
#define LBX_RO_MIN_CTAS 6
#endif
template <class C, int MODE, bool ZI>
__global__ void __launch_bounds__(RO_THREADS, LBX_RO_MIN_CTAS) k_mf_cs_rows(ROArgs a) {
  extern __shared__ double ro_smem[];
  // grid = (row tiles of one z-plane, z-planes of the largest fab, fabs): no division anywhere
  const int b = a.dt[a.fab0 + (int)blockIdx.z].lid;       // a.nfabs counts the boxes this rank owns
  const DFabT D = a.dt[b];
  if (!D.local) return;                                   // a peer's box: its owner streams it
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n0 = D.n[0], n1 = D.n[1], n2 = D.n[2];
  const int jr = (int)blockIdx.x * a.warps + w, kr = (int)blockIdx.y;
  if (w >= a.warps || jr >= n1 || kr >= n2) return;       // warp-uniform
  const bool valid_row = jr >= 2 && jr < n1 - 2 && kr >= 2 && kr < n2 - 2;
  // rows whose 15 destination rows are all rows of valid grown by 1 and -- under ZI -- all valid rows: phase 2 needs no
  // per-population test at all
  const bool deep = ZI ? (jr >= 3 && jr < n1 - 3 && kr >= 3 && kr < n2 - 3) : valid_row;
  const int pitch = a.pitch;
  double* sm = ro_smem + 2 + w * (NV * pitch);
  const unsigned sc = (unsigned)n0 * (unsigned)n1 * (unsigned)n2;       // plane stride (elements)
  const unsigned rowoff = (unsigned)n0 * ((unsigned)jr + (unsigned)n1 * (unsigned)kr);
  double* dfab = static_cast<double*>(D.p);
  if (lane < NV && cx(lane) != 0) sm[lane * pitch + (cx(lane) > 0 ? 0 : n0 - 1)] = 0.0;

  // ------------------------------------------------------------------ phase 1: the row's populations -> sm[p][x]
  if (valid_row) {
    const double* vfab = a.vbase + (dfab - a.dbase);
    const int* mrow = a.mt ? static_cast<const int*>(a.mt[b].p) + rowoff : nullptr;
    for (int x = 2 + lane; x < n0 - 2; x += 32) {
      double f[NV];
      if (mrow && mrow[x] == a.fine_val) {
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = 0.0;
      } else {
        const double* sp = vfab + (rowoff + (unsigned)x);
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = __ldcs(sp + (unsigned)p * sc);
        C::collide(f, a.omega_s, a.omega_b);
      }
      double* s = sm + x;
#pragma unroll
      for (int p = 0; p < NV; ++p) s[p * pitch + cx(p)] = f[p];
    }
  }
  // ghost cells of the row: all of it, or the 2 + 2 cells at the ends of a valid row.  Warp-uniform population masks of
  // this source row (the y, z parts of every per-cell decision): a ghost source pushes p only into valid grown by 1 --
  // under ZI only into valid cells
  const unsigned NEEDROW = axis_sel(1, jr, n1, ZI ? 2 : 1, false) & axis_sel(2, kr, n2, ZI ? 2 : 1, false);
  if (MODE == 1 && !valid_row) {
    // own ghost cells of a whole ghost row: a plain copy of the needed planes, 16 bytes per lane (cells of a needed
    // plane that have no destination are copied too: phase 2's general path tests for them)
    const double2* g2 = reinterpret_cast<const double2*>(static_cast<const double*>(a.gt[b].p) + rowoff);
    const unsigned sc2 = sc >> 1;
    for (int q = lane; q < (n0 >> 1); q += 32) {
      double* s = sm + 2 * q;
#pragma unroll
      for (int p = 0; p < NV; ++p)
        if (NEEDROW >> p & 1u) {
          const double2 v = __ldcs(g2 + (q + (unsigned)p * sc2));
          if (cx(p) == 0) *reinterpret_cast<double2*>(s + p * pitch) = v;
          else { s[p * pitch + cx(p)] = v.x; s[p * pitch + cx(p) + 1] = v.y; }
        }
    }
  } else {
    for (int q = lane; q < (valid_row ? 4 : n0); q += 32) {
      const int x = valid_row ? (q < 2 ? q : n0 - 4 + q) : q;
      const unsigned need = NEEDROW & axis_sel(0, x, n0, ZI ? 2 : 1, false);
      const double* sp = nullptr;
      const double* spb = nullptr;
      unsigned ssc = sc;
      bool collide_src = false;
      if (MODE == 1) {
        sp = static_cast<const double*>(a.gt[b].p) + (rowoff + (unsigned)x);
      } else {
        const int2 e = a.plan.tab[a.plan.tab_first[b] + ro_shell_index(n0, n1, n2, x, jr, kr)];
        const int kind = e.x & 3, fab = e.x >> 2;
        if (kind) {
          const DFabT* S = (kind == 1 ? a.plan.s0 : a.plan.s1) + fab;
          const int4 h0 = reinterpret_cast<const int4*>(S)[0], h1 = reinterpret_cast<const int4*>(S)[1];   // p, lo0, lo1 | lo2, n0, n1, n2
          const double* base = reinterpret_cast<const double*>(((unsigned long long)(unsigned)h0.y << 32) | (unsigned)h0.x);
          ssc = (unsigned)h1.y * (unsigned)h1.z * (unsigned)h1.w;
          sp = base + e.y;
          if (MODE == 3) {
            collide_src = kind == 1;
            if (kind == 2 && a.plan.s1b) spb = static_cast<const double*>(a.plan.s1b[fab].p) + e.y;
          }
        } else if (a.plan.fb) {
          sp = static_cast<const double*>(a.plan.fb[b].p) + (rowoff + (unsigned)x);
        }
      }
      double f[NV];
      if (MODE == 3 && collide_src) {
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = __ldcs(sp + (unsigned)p * ssc);
        C::collide(f, a.omega_s, a.omega_b);
        // the populations without a destination must read as 0 in phase 2's test-free path
#pragma unroll
        for (int p = 0; p < NV; ++p)
          if (!(need >> p & 1u)) f[p] = 0.0;
      } else if (MODE == 3 && spb) {
#pragma unroll
        for (int p = 0; p < NV; ++p)
          f[p] = (need >> p & 1u) ? __dadd_rn(__dmul_rn(a.plan.wa, __ldcs(sp + (unsigned)p * ssc)), __dmul_rn(a.plan.wb, __ldcs(spb + (unsigned)p * ssc))) : 0.0;
      } else {
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = ((need >> p & 1u) && sp) ? __ldcs(sp + (unsigned)p * ssc) : 0.0;
      }
      double* s = sm + x;
#pragma unroll
      for (int p = 0; p < NV; ++p) s[p * pitch + cx(p)] = f[p];
    }
  }
  __syncwarp();

  // ------------------------------------------------------------------ phase 2: whole destination rows, one per population
  const int npair = n0 >> 1;
  const int oY = n0 >> 1, oZ = (n0 >> 1) * n1;            // row / plane pitch of the destination in 16-byte pairs
  const unsigned sc2 = sc >> 1;
  if (deep) {
    for (int q = lane; q < npair; q += 32) {
      double2* pd = reinterpret_cast<double2*>(dfab + rowoff) + q;
      const double* ss = sm + 2 * q;
#pragma unroll
      for (int p = 0; p < NV; ++p) {
        __stcs(pd + (cy(p) * oY + cz(p) * oZ), *reinterpret_cast<const double2*>(ss));
        pd += sc2;
        ss += pitch;
      }
    }
    return;
  }
  // Rows on the rim of the valid y-z range and ghost rows.  A row writes only destination rows of valid grown by 1 (mask
  // NZ); a ring-2 row zeroes its own 15 planes itself (whole rows, whole sectors: the reference's fresh fab).
  const unsigned NZ = axis_sel(1, jr, n1, 1, false) & axis_sel(2, kr, n2, 1, false);
  const bool ring2 = jr == 0 || jr == n1 - 1 || kr == 0 || kr == n2 - 1;
  unsigned TV = 0, SV = 0;
  if (ZI) {
    TV = axis_sel(1, jr, n1, 2, false) & axis_sel(2, kr, n2, 2, false);      // destination row lies in the valid y-z range
    SV = axis_sel(1, jr, n1, 2, true) & axis_sel(2, kr, n2, 2, true);        // (destination - 2 c) lies in the valid y-z range
  }
  for (int q = lane; q < npair; q += 32) {
    const int x0 = 2 * q;
    double2* pd = reinterpret_cast<double2*>(dfab + rowoff) + q;
    const double* ss = sm + x0;
    // populations whose value survives in cell x0 / x0 + 1 of the destination row: not a ring-2 cell in x, and under
    // ZeroInvalidComponents (:604-617) a valid destination cell or one whose (destination - 2 c) is valid
    unsigned K0 = q == 0 ? 0u : NZ, K1 = q == npair - 1 ? 0u : NZ;
    if (ZI) {
      const int x1 = x0 + 1;
      const unsigned all = (1u << NV) - 1;
      const unsigned D0 = (x0 >= 2 && x0 < n0 - 2) ? all : 0u, D1 = (x1 >= 2 && x1 < n0 - 2) ? all : 0u;
      K0 &= (TV & D0) | (SV & axis_sel(0, x0, n0, 2, true, 2));      // (destination - 2 c) valid in x
      K1 &= (TV & D1) | (SV & axis_sel(0, x1, n0, 2, true, 2));
    }
#pragma unroll
    for (int p = 0; p < NV; ++p) {
      const unsigned bit = 1u << p;
      if (NZ & bit) {
        double2 v = *reinterpret_cast<const double2*>(ss);
        if (!(K0 & bit)) v.x = 0.0;
        if (!(K1 & bit)) v.y = 0.0;
        __stcs(pd + (cy(p) * oY + cz(p) * oZ), v);
      }
      if (ring2) __stcs(pd, make_double2(0.0, 0.0));
      pd += sc2;
      ss += pitch;
    }
  }
}

}  // namespace lbx
