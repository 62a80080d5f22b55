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

constexpr int RO_THREADS = 256;

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
// inverse: entry t -> (x, jr, kr) relative to the grown box
__device__ __forceinline__ void ro_shell_cell(int n0, int n1, int n2, long long t, int& x, int& jr, int& kr) {
  const long long A = 2ll * n0 * n1, B = 2ll * n0 * (n2 - 4);
  if (t < 2 * A) {
    const int s = t >= A;
    const long long r = t - s * A;
    x = (int)(r % n0);
    jr = (int)((r / n0) % n1);
    kr = (s ? n2 - 2 : 0) + (int)(r / ((long long)n0 * n1));
    return;
  }
  t -= 2 * A;
  if (t < 2 * B) {
    const int s = t >= B;
    const long long r = t - s * B;
    x = (int)(r % n0);
    jr = (s ? n1 - 2 : 0) + (int)((r / n0) % 2);
    kr = 2 + (int)(r / (2ll * n0));
    return;
  }
  t -= 2 * B;
  const int q = (int)(t % 4);
  x = q < 2 ? q : n0 - 4 + q;
  jr = 2 + (int)((t / 4) % (n1 - 4));
  kr = 2 + (int)(t / (4ll * (n1 - 4)));
}

// One launch per plan and geometry: the backward descriptor search of the tile kernel, done ONCE per ghost cell
// instead of once per ghost cell per time step.  grid = (tiles over the largest shell, fab).
static __global__ void __launch_bounds__(MFT) k_plan_resolve(const DFabT* __restrict__ dt, int nfabs, const GDst* __restrict__ dsts,
                                                      const int* __restrict__ fab_first, const GDesc* __restrict__ descs,
                                                      const DFabT* __restrict__ s0, const DFabT* __restrict__ s1,
                                                      int2* __restrict__ tab, const long long* __restrict__ tab_first) {
  const int b = mf_fab_index();
  if (b >= nfabs) return;
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

// populations a ghost source cell (x, jr, kr relative to the grown box) pushes: bit p set iff dst = x + c_p lies
// inside valid grown by 1 and (under ZeroInvalidComponents) is a VALID cell -- a ghost source reaches a ghost
// destination only with x - c_p outside the valid box, which the filter zeroes
__device__ __forceinline__ unsigned ro_need_mask(int x, int jr, int kr, int n0, int n1, int n2, bool zi) {
  unsigned m = 0;
  const int lo = zi ? 2 : 1;                      // destination range per direction: [lo, n - 1 - lo]
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    const int dx = x + cx(p), dy = jr + cy(p), dz = kr + cz(p);
    if (dx >= lo && dx < n0 - lo && dy >= lo && dy < n1 - lo && dz >= lo && dz < n2 - lo) m |= 1u << p;
  }
  return m;
}

template <class C, int MODE>   // MODE 1: own ghost cells; 2: FillPatch plan (Rohde pair); 3: conventional level step
__global__ void __launch_bounds__(RO_THREADS) k_mf_cs_rows(ROArgs a) {
  extern __shared__ double ro_smem[];
  const int b = mf_fab_index();
  if (b >= a.nfabs) return;
  const DFabT D = a.dt[b];
  if (!D.local) return;                                   // a peer's box: its owner streams it
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n0 = D.n[0], n1 = D.n[1], n2 = D.n[2];
  const int row = (int)blockIdx.x * a.warps + w;
  if (w >= a.warps || row >= n1 * n2) return;             // warp-uniform
  const int jr = row % n1, kr = row / n1;
  const int j = D.lo[1] + jr, k = D.lo[2] + kr;
  const int ey = jr < 2 ? jr - 2 : jr >= n1 - 2 ? jr - (n1 - 3) : 0;
  const int ez = kr < 2 ? kr - 2 : kr >= n2 - 2 ? kr - (n2 - 3) : 0;
  const bool valid_row = ey == 0 && ez == 0;
  double* sm = ro_smem + (size_t)w * NV * a.pitch;
  const long long sc = mf_stride(D);
  const long long rowoff = (long long)n0 * (jr + (long long)n1 * kr);
  const bool zi = a.zero_invalid != 0;

  // ------------------------------------------------------------------ phase 1: the row's populations -> sm[p][x]
  if (valid_row) {
    const double* srow = a.vbase + (static_cast<double*>(D.p) - a.dbase) + rowoff;
    const int* mrow = a.mt ? static_cast<const int*>(a.mt[b].p) + rowoff : nullptr;
    for (int x = 2 + lane; x < n0 - 2; x += 32) {
      double f[NV];
      if (mrow && mrow[x] == a.fine_val) {
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = 0.0;
      } else {
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = __ldcs(srow + p * sc + x);
        C::collide(f, a.omega_s, a.omega_b);
      }
#pragma unroll
      for (int p = 0; p < NV; ++p) sm[p * a.pitch + x] = f[p];
    }
  }
  // ghost cells of the row: all of it, or the 2 + 2 cells at the ends of a valid row
  for (int q = lane; q < (valid_row ? 4 : n0); q += 32) {
    const int x = valid_row ? (q < 2 ? q : n0 - 4 + q) : q;
    const unsigned need = ro_need_mask(x, jr, kr, n0, n1, n2, zi);
    const double* sp = nullptr;
    const double* spb = nullptr;
    long long ssc = sc;
    bool collide_src = false;
    if (MODE == 1) {
      sp = static_cast<const double*>(a.gt[b].p) + rowoff + x;
    } else {
      const int2 e = a.plan.tab[a.plan.tab_first[b] + ro_shell_index(n0, n1, n2, x, jr, kr)];
      const int kind = e.x & 3, fab = e.x >> 2;
      if (kind) {
        const DFabT* S = (kind == 1 ? a.plan.s0 : a.plan.s1) + fab;
        const int4 h0 = reinterpret_cast<const int4*>(S)[0], h1 = reinterpret_cast<const int4*>(S)[1];   // p, lo0, lo1 | lo2, n0, n1, n2
        const double* base = reinterpret_cast<const double*>(((unsigned long long)(unsigned)h0.y << 32) | (unsigned)h0.x);
        ssc = (long long)h1.y * h1.z * h1.w;
        sp = base + e.y;
        if (MODE == 3) {
          collide_src = kind == 1;
          if (kind == 2 && a.plan.s1b) spb = static_cast<const double*>(a.plan.s1b[fab].p) + e.y;
        }
      } else if (a.plan.fb) {
        sp = static_cast<const double*>(a.plan.fb[b].p) + rowoff + x;
      }
    }
    double f[NV];
    if (MODE == 3 && collide_src) {
#pragma unroll
      for (int p = 0; p < NV; ++p) f[p] = __ldcs(sp + p * ssc);
      C::collide(f, a.omega_s, a.omega_b);
    } else if (MODE == 3 && spb) {
#pragma unroll
      for (int p = 0; p < NV; ++p)
        f[p] = (need >> p & 1u) ? __dadd_rn(__dmul_rn(a.plan.wa, __ldcs(sp + p * ssc)), __dmul_rn(a.plan.wb, __ldcs(spb + p * ssc))) : 0.0;
    } else {
#pragma unroll
      for (int p = 0; p < NV; ++p) f[p] = ((need >> p & 1u) && sp) ? __ldcs(sp + p * ssc) : 0.0;
    }
#pragma unroll
    for (int p = 0; p < NV; ++p) sm[p * a.pitch + x] = f[p];
  }
  __syncwarp();

  // ------------------------------------------------------------------ phase 2: whole destination rows, one per population
  double* dfab = static_cast<double*>(D.p);
  const int npair = n0 >> 1;
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    const int tj = jr + cy(p), tk = kr + cz(p);
    if (tj >= 0 && tj < n1 && tk >= 0 && tk < n2) {
      const bool zero_row = tj == 0 || tj == n1 - 1 || tk == 0 || tk == n2 - 1;       // ring 2 of the fresh fab
      double2* drow = reinterpret_cast<double2*>(dfab + p * sc + (long long)n0 * (tj + (long long)n1 * tk));
      if (zero_row) {
        for (int q = lane; q < npair; q += 32) __stcs(drow + q, make_double2(0.0, 0.0));
      } else {
        const double* srow = sm + p * a.pitch - cx(p);
        const bool t_valid = tj >= 2 && tj < n1 - 2 && tk >= 2 && tk < n2 - 2;         // destination row inside the valid y-z range
        const bool s_valid = jr - cy(p) >= 2 && jr - cy(p) < n1 - 2 && kr - cz(p) >= 2 && kr - cz(p) < n2 - 2;   // (dst - 2c) in y, z
        for (int q = lane; q < npair; q += 32) {
          double v[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int x = 2 * q + e;
            if (x == 0 || x == n0 - 1) {
              v[e] = 0.0;                                                              // ring 2 in x
            } else {
              v[e] = srow[x];
              if (zi) {
                const bool dst_valid = t_valid && x >= 2 && x < n0 - 2;
                const int x2 = x - 2 * cx(p);
                if (!dst_valid && !(s_valid && x2 >= 2 && x2 < n0 - 2)) v[e] = 0.0;    // ZeroInvalidComponents
              }
            }
          }
          __stcs(drow + q, make_double2(v[0], v[1]));
        }
      }
    }
    // ring-2 rows: planes whose natural writer (j - c_y, k - c_z) lies outside the grown box are zeroed by the row itself
    const int sj = jr - cy(p), sk = kr - cz(p);
    if ((cy(p) != 0 || cz(p) != 0) && (sj < 0 || sj >= n1 || sk < 0 || sk >= n2)) {
      double2* orow = reinterpret_cast<double2*>(dfab + p * sc + rowoff);
      for (int q = lane; q < npair; q += 32) __stcs(orow + q, make_double2(0.0, 0.0));
    }
  }
}

}  // namespace lbx
