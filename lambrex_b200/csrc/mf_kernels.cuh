// Collision-independent batched kernels of the AMR path (compiled once, in lbx_mf.cu).
//
// kernel            replaces (reference file:line)                               HBM traffic
// k_mf_stream       Stream + PropagatePoint on valid grown by 1, fresh zeroed     240 B/cell
//                   destination (src/AmrSim.cpp:109-122, include/component.h:22-29)
// k_mf_zero_invalid ZeroInvalidComponents (src/AmrSim.cpp:604-617)                ghost shell only
// k_mf_zero_ring    InitPostCollision's outer-ring zeroing (src/AmrSim.cpp:477-482)
// k_plan_apply      AMReX ParallelCopy / FillBoundary / FillPatchSingleLevel /    <= 240 B/cell
//                   FillPatchTwoLevels + PCInterp / InterpFromCoarseLevel /
//                   sum_fine_to_coarse / makeFineMask (src/AmrSim.cpp:21,132,371,385,406,425,598)
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace lbx {

// dst(x,p) = src(x - c_p, p) for x in valid grown by 1; every other cell of dst (ghost ring 2)
// is set to 0: the reference streams into a freshly allocated fab whose ring 2 it never
// writes (policy NEW_FAB_FILL = 0).
__global__ void __launch_bounds__(MFT) k_mf_stream(const DFabT* __restrict__ st, const DFabT* __restrict__ dt,
                                                   int nfabs, int grow_all) {
  const int b = mf_local_fab(dt, nfabs);
  if (b < 0) return;
  const DFabT D = dt[b];
  if (!D.local) return;
  int i, j, k;
  if (!mf_cell(D, grow_all, i, j, k)) return;
  double* dp = static_cast<double*>(D.p) + mf_off(D, i, j, k);
  const long long dsc = mf_stride(D);
  const bool inner = i >= D.vlo[0] - 1 && i <= D.vhi[0] + 1 && j >= D.vlo[1] - 1 && j <= D.vhi[1] + 1 &&
                     k >= D.vlo[2] - 1 && k <= D.vhi[2] + 1;
  if (!inner) {
#pragma unroll
    for (int p = 0; p < NV; ++p) dp[p * dsc] = 0.0;
    return;
  }
  const DFabT S = st[b];
  const double* sp = static_cast<const double*>(S.p);
  const long long ssc = mf_stride(S);
#pragma unroll
  for (int p = 0; p < NV; ++p) dp[p * dsc] = sp[p * ssc + mf_off(S, i - cx(p), j - cy(p), k - cz(p))];
}

// For every cell of the ghost shell (all `grow_all` rings): component m survives only if
// pos - 2 c_m lies in the box's valid region.
__global__ void __launch_bounds__(MFT) k_mf_zero_invalid(const DFabT* __restrict__ ft, int nfabs, int grow_all) {
  const int b = mf_local_fab(ft, nfabs);
  if (b < 0) return;
  const DFabT F = ft[b];
  if (!F.local) return;
  int i, j, k;
  if (!mf_cell(F, grow_all, i, j, k)) return;
  if (mf_in_valid(F, i, j, k)) return;
  double* fp = static_cast<double*>(F.p) + mf_off(F, i, j, k);
  const long long sc = mf_stride(F);
#pragma unroll
  for (int p = 0; p < NV; ++p)
    if (!mf_in_valid(F, i - 2 * cx(p), j - 2 * cy(p), k - 2 * cz(p))) fp[p * sc] = 0.0;
}

// comp <- 0 on the outermost `depth` rings of every (grown) fab box.
__global__ void __launch_bounds__(MFT) k_mf_zero_ring(const DFabT* __restrict__ ft, int nfabs, int grow_all, int depth,
                                                      int comp) {
  const int b = mf_local_fab(ft, nfabs);
  if (b < 0) return;
  const DFabT F = ft[b];
  if (!F.local) return;
  int i, j, k;
  if (!mf_cell(F, grow_all, i, j, k)) return;
  const int g = grow_all - depth;   // cells inside valid grown by g are untouched
  if (i >= F.vlo[0] - g && i <= F.vhi[0] + g && j >= F.vlo[1] - g && j <= F.vhi[1] + g && k >= F.vlo[2] - g &&
      k <= F.vhi[2] + g)
    return;
  static_cast<double*>(F.p)[comp * mf_stride(F) + mf_off(F, i, j, k)] = 0.0;
}

// First half of sum_fine_to_coarse (src/AmrSim.cpp:598; amrex_avgdown): every cell of the
// coarsened fine boxes INCLUDING their ghost ring -- ct(x) = mean of the r^3 fine cells
// ft(r x .. r x + r - 1), summed iref fastest, then jref, kref.  Box b of ct is box b of ft
// coarsened (valid and ghosts).  One thread per coarse cell, 15 populations, every load issued
// before the first use; no descriptor search and no divergence (the ADD into the coarse level is
// then a plain ParallelCopy plan over these patches).
__global__ void __launch_bounds__(MFT) k_mf_average_down(const DFabT* __restrict__ ft, const DFabT* __restrict__ ct, int nfabs,
                                                         int cgrow, int r) {
  const int b = mf_local_fab(ct, nfabs);
  if (b < 0) return;
  const DFabT Cf = ct[b];
  if (!Cf.local) return;
  int i, j, k;
  if (!mf_cell(Cf, cgrow, i, j, k)) return;
  const DFabT S = ft[b];
  const long long ssc = mf_stride(S), sy = S.n[0], sz = (long long)S.n[0] * S.n[1];
  const double* sp = static_cast<const double*>(S.p) + mf_off(S, i * r, j * r, k * r);
  double* dp = static_cast<double*>(Cf.p) + mf_off(Cf, i, j, k);
  const long long dsc = mf_stride(Cf);
  if (r == 2) {
    double v[NV];
    if (((i * r - S.lo[0]) & 1) == 0 && (S.n[0] & 1) == 0) {          // 16-byte aligned pairs
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const double2 a = *reinterpret_cast<const double2*>(sp + c * ssc);
        const double2 bb = *reinterpret_cast<const double2*>(sp + c * ssc + sy);
        const double2 cc = *reinterpret_cast<const double2*>(sp + c * ssc + sz);
        const double2 d = *reinterpret_cast<const double2*>(sp + c * ssc + sz + sy);
        v[c] = ((((((a.x + a.y) + bb.x) + bb.y) + cc.x) + cc.y) + d.x + d.y) * 0.125;
      }
    } else {
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const double* q = sp + c * ssc;
        const double a0 = q[0], a1 = q[1], b0 = q[sy], b1 = q[sy + 1], c0 = q[sz], c1 = q[sz + 1], d0 = q[sz + sy],
                     d1 = q[sz + sy + 1];
        v[c] = ((((((a0 + a1) + b0) + b1) + c0) + c1) + d0 + d1) * 0.125;
      }
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) dp[c * dsc] = v[c];
    return;
  }
  const double w = 1.0 / (double)(r * r * r);
  for (int c = 0; c < NV; ++c) {
    double acc = 0;
    for (int kr = 0; kr < r; ++kr)
      for (int jr = 0; jr < r; ++jr)
        for (int ir = 0; ir < r; ++ir) acc += sp[c * ssc + kr * sz + jr * sy + ir];
    dp[c * dsc] = acc * w;
  }
}

// dst = a * x + b * y on the valid cells of every box, `ncomp` components (MultiFab::LinComb --
// the time interpolation of FillPatchTwoLevels between two coarse states [AMReX]).  Products
// and sum are rounded separately (no FMA contraction): bit-identical to the CPU statement.
__global__ void __launch_bounds__(MFT) k_mf_lincomb(const DFabT* __restrict__ dt, const DFabT* __restrict__ xt,
                                                    const DFabT* __restrict__ yt, int nfabs, int ncomp, double a, double b) {
  const int q = mf_local_fab(dt, nfabs);
  if (q < 0) return;
  const DFabT D = dt[q];
  if (!D.local) return;
  int i, j, k;
  if (!mf_cell(D, 0, i, j, k)) return;
  const DFabT X = xt[q], Y = yt[q];
  const double* xp = static_cast<const double*>(X.p) + mf_off(X, i, j, k);
  const double* yp = static_cast<const double*>(Y.p) + mf_off(Y, i, j, k);
  double* dp = static_cast<double*>(D.p) + mf_off(D, i, j, k);
  const long long xs = mf_stride(X), ys = mf_stride(Y), ds = mf_stride(D);
  for (int c0 = 0; c0 < ncomp; c0 += 5) {          // 10 loads in flight per thread
    double v[5], w[5];
#pragma unroll
    for (int c = 0; c < 5; ++c)
      if (c0 + c < ncomp) { v[c] = __ldcs(xp + (c0 + c) * xs); w[c] = __ldcs(yp + (c0 + c) * ys); }
#pragma unroll
    for (int c = 0; c < 5; ++c)
      if (c0 + c < ncomp) dp[(c0 + c) * ds] = __dadd_rn(__dmul_rn(a, v[c]), __dmul_rn(b, w[c]));
  }
}

// Refinement criterion on the device (replaces the host loop of ErrorEst, src/AmrSim.cpp:633-663, for
// the gradient criterion of SURVEY.md 8f-2): a valid cell is tagged when the squared central
// difference of the density, (|rho(x+e_d) - rho(x-e_d)|^2 summed over d) / 4, exceeds thr2.  rho has at
// least one FILLED ghost cell.  Tagged cells get `set_val`; others are left alone.  Every
// operation is rounded separately, in a fixed order: the tag set is bit-reproducible on the CPU.
__global__ void __launch_bounds__(MFT) k_mf_tag_gradient(const DFabT* __restrict__ rt, const DFabT* __restrict__ tt, int nfabs,
                                                         double thr2, int set_val) {
  const int q = mf_local_fab(tt, nfabs);
  if (q < 0) return;
  const DFabT T = tt[q];
  if (!T.local) return;
  int i, j, k;
  if (!mf_cell(T, 0, i, j, k)) return;
  const DFabT R = rt[q];
  const double* rp = static_cast<const double*>(R.p) + mf_off(R, i, j, k);
  const long long sy = R.n[0], sz = (long long)R.n[0] * R.n[1];
  const double gx = __dsub_rn(rp[1], rp[-1]), gy = __dsub_rn(rp[sy], rp[-sy]), gz = __dsub_rn(rp[sz], rp[-sz]);
  const double g2 = __dmul_rn(0.25, __dadd_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), __dmul_rn(gz, gz)));
  if (g2 > thr2) static_cast<int*>(T.p)[mf_off(T, i, j, k)] = set_val;
}

// Generic derived variable (include/derived_var.h:55-91 fill_box + a per-cell `calculate`): every
// derived variable the reference defines or sketches -- density (d3q15_bgk.h:34-41), momentum
// density and stress (:43-55, commented out) -- is a LINEAR moment of the populations,
//   out_c(x) = sum_p W[c][p] f_p(x),   optionally divided by rho(x) = sum_p f_p(x),
// so one kernel with the weight rows as a launch parameter serves them all.  Accumulated from 0.0
// in increasing p with separately rounded multiply and add (the order of `calculate`'s loop).
constexpr int LM_MAX = 10;
struct LMWeights { double w[LM_MAX][NV]; };
__global__ void __launch_bounds__(MFT) k_mf_linear_moments(const DFabT* __restrict__ ft, const DFabT* __restrict__ ot, int nfabs,
                                                           int ncomp, int normalise, const __grid_constant__ LMWeights W) {
  const int q = mf_local_fab(ot, nfabs);
  if (q < 0) return;
  const DFabT O = ot[q];
  if (!O.local) return;
  int i, j, k;
  if (!mf_cell(O, 0, i, j, k)) return;
  const DFabT F = ft[q];
  const double* fp = static_cast<const double*>(F.p) + mf_off(F, i, j, k);
  const long long fs = mf_stride(F), os = mf_stride(O);
  double f[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = __ldcs(fp + p * fs);
  double inv = 1.0;
  if (normalise) {
    double rho = 0.0;
#pragma unroll
    for (int p = 0; p < NV; ++p) rho = __dadd_rn(rho, f[p]);
    inv = rho;
  }
  double* op = static_cast<double*>(O.p) + mf_off(O, i, j, k);
  for (int c = 0; c < ncomp; ++c) {
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < NV; ++p) acc = __dadd_rn(acc, __dmul_rn(f[p], W.w[c][p]));
    op[c * os] = normalise ? __ddiv_rn(acc, inv) : acc;
  }
}

// User arrays of the reference's API are C-ordered, i slowest, component fastest
// (CLindex, include/AmrSim.h:79-83): user[(((i-d0)*NY + (j-d1))*NZ + (k-d2))*ncomp + n].
// TO_FAB: valid cells of every fab <- user (InitDensity / InitVelocity, src/AmrSim.cpp:138-295);
// else  : user <- valid cells (bulk form of GetDensity / GetVelocity, :824-843).
template <bool TO_FAB>
__global__ void __launch_bounds__(MFT) k_mf_user(const DFabT* __restrict__ ft, int nfabs, double* __restrict__ user,
                                                 int d0, int d1, int d2, int nx, int ny, int nz, int ncomp, int local_only) {
  const int b = mf_fab_index();
  if (b >= nfabs) return;
  const DFabT F = ft[b];
  if ((TO_FAB || local_only) && !F.local) return;   // reading (to user) also takes peers' boxes over NVLink, unless local_only
  int i, j, k;
  if (!mf_cell(F, 0, i, j, k)) return;
  if (i < d0 || i >= d0 + nx) return;               // the user array may cover an x-range of the boxes only (chunked staging)
  double* fp = static_cast<double*>(F.p) + mf_off(F, i, j, k);
  const long long sc = mf_stride(F);
  double* up = user + ((((long long)(i - d0) * ny + (j - d1)) * nz + (k - d2)) * ncomp);
  for (int n = 0; n < ncomp; ++n) {
    if (TO_FAB) fp[n * sc] = up[n];
    else up[n] = fp[n * sc];
  }
}

// Same transposition through shared memory: a CTA moves a 32 (x) x 32 (z) tile of one y-row, all
// components, so that BOTH sides are coalesced -- the fab side along x, the user side along its
// contiguous (k, n) run.  grid = (x-tiles * z-tiles of the largest box, y extent, fab).
constexpr int UT = 32;
template <bool TO_FAB, int NC>
__global__ void __launch_bounds__(UT * 8) k_mf_user_tiled(const DFabT* __restrict__ ft, double* __restrict__ user,
                                                          int tiles_x, int d0, int d1, int d2, int nx, int ny, int nz, int local_only) {
  __shared__ double tile[NC][UT][UT + 1];     // [n][z][x]
  const DFabT F = ft[blockIdx.z];
  if ((TO_FAB || local_only) && !F.local) return;
  const int j = F.vlo[1] + blockIdx.y;
  const int i0 = F.vlo[0] + (blockIdx.x % tiles_x) * UT, k0 = F.vlo[2] + (blockIdx.x / tiles_x) * UT;
  if (j > F.vhi[1] || i0 > F.vhi[0] || k0 > F.vhi[2]) return;           // block-uniform
  if (i0 + UT <= d0 || i0 >= d0 + nx) return;                             // block-uniform: tile outside the user array's x-range
  const int ni = min(UT, F.vhi[0] - i0 + 1), nk = min(UT, F.vhi[2] - k0 + 1);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long sc = mf_stride(F);
  double* fp = static_cast<double*>(F.p);
  if (!TO_FAB) {
    if (tx < ni)
      for (int kk = ty; kk < nk; kk += 8) {
        const long long o = mf_off(F, i0 + tx, j, k0 + kk);
#pragma unroll
        for (int n = 0; n < NC; ++n) tile[n][kk][tx] = fp[n * sc + o];
      }
    __syncthreads();
    for (int ii = ty; ii < ni; ii += 8) {
      if (i0 + ii < d0 || i0 + ii >= d0 + nx) continue;       // outside the x-range the user array covers
      double* up = user + (((long long)(i0 + ii - d0) * ny + (j - d1)) * nz + (k0 - d2)) * NC;
      for (int e = tx; e < nk * NC; e += UT) up[e] = tile[e % NC][e / NC][ii];
    }
  } else {
    for (int ii = ty; ii < ni; ii += 8) {
      if (i0 + ii < d0 || i0 + ii >= d0 + nx) continue;
      const double* up = user + (((long long)(i0 + ii - d0) * ny + (j - d1)) * nz + (k0 - d2)) * NC;
      for (int e = tx; e < nk * NC; e += UT) tile[e % NC][e / NC][ii] = up[e];
    }
    __syncthreads();
    if (tx < ni && i0 + tx >= d0 && i0 + tx < d0 + nx)
      for (int kk = ty; kk < nk; kk += 8) {
        const long long o = mf_off(F, i0 + tx, j, k0 + kk);
#pragma unroll
        for (int n = 0; n < NC; ++n) fp[n * sc + o] = tile[n][kk][tx];
      }
  }
}

__global__ void k_fill_f64(double* __restrict__ p, long long n, double v) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    p[t] = v;
}

// Separable initial condition: f(i,j,k,n) = profile[(x_axis - axis_lo) * ncomp + n] on the valid cells of every
// local box -- a field that varies along ONE axis (the planar pulse varies along z, the shear wave u_x along y),
// stated by a profile of axis-length entries instead of a whole-domain host array (1024^3: 34 GB per rank).
__global__ void __launch_bounds__(MFT) k_mf_fill_profile(const DFabT* __restrict__ ft, int nfabs, int ncomp,
                                                         const double* __restrict__ profile, int axis, int axis_lo) {
  const int b = mf_local_fab(ft, nfabs);
  if (b < 0) return;
  const DFabT F = ft[b];
  if (!F.local) return;
  int i, j, k;
  if (!mf_cell(F, 0, i, j, k)) return;
  const int a = (axis == 0 ? i : axis == 1 ? j : k) - axis_lo;
  double* fp = static_cast<double*>(F.p) + mf_off(F, i, j, k);
  const long long sc = mf_stride(F);
  for (int n = 0; n < ncomp; ++n) fp[n * sc] = profile[(long long)a * ncomp + n];
}

template <class T>
__global__ void __launch_bounds__(MFT) k_mf_setval(const DFabT* __restrict__ ft, int nfabs, int grow_all, int ncomp,
                                                   T value) {
  const int b = mf_local_fab(ft, nfabs);
  if (b < 0) return;
  const DFabT F = ft[b];
  if (!F.local) return;
  int i, j, k;
  if (!mf_cell(F, grow_all, i, j, k)) return;
  T* fp = static_cast<T*>(F.p) + mf_off(F, i, j, k);
  const long long sc = mf_stride(F);
  for (int c = 0; c < ncomp; ++c) fp[c * sc] = value;
}

// ---------------------------------------------------------------------------
// Gather plans.  A plan is a list of descriptors grouped by destination fab; one
// thread per destination cell walks the descriptors of its fab:
//   COPY: the LAST descriptor whose region holds the cell wins (= sequential CPU
//         copies in list order);  ADD: every match is added, in list order
//         (deterministic, no atomics).
// Source index maps (x = destination cell, r = ratio):
//   G_COPY  src(x + shift)                       same-level copies, periodic images
//   G_PC    src(floor(x / r) + shift)            piecewise-constant interpolation (PCInterp)
//   G_AVG   r^-3 * sum src(r x + shift + ref)    fine -> coarse average (amrex_avgdown order:
//                                                iref fastest, then jref, kref)
//   G_CONST value                                setVal on a region (masks)
//   G_NONE  --                                   no source: the cell keeps its value; only widens the
//                                                tiled bounding box of its group to the whole region
// ---------------------------------------------------------------------------
// The value one matching descriptor gives one destination cell, components [c0, c0 + NC), into v.  Every load of
// the cell is issued before the first use, which is what keeps these gather kernels bandwidth- rather than
// latency-bound.  false: the descriptor is G_NONE (the cell keeps its value).
template <class T, int NC>
__device__ __forceinline__ bool g_fetch(const GDesc& g, const DFabT* __restrict__ s0, const DFabT* __restrict__ s1,
                                        int i, int j, int k, int c0, T (&v)[NC]) {
  if (g.kind == G_NONE) return false;
  if (g.kind == G_CONST) {
#pragma unroll
    for (int c = 0; c < NC; ++c) v[c] = (T)g.value;
    return true;
  }
  const DFabT S = (g.src_set ? s1 : s0)[g.src_fab];
  const long long ssc = mf_stride(S);
  if (g.kind == G_AVG) {
    const int r = g.ratio;
    const int i0 = i * r + g.shift[0];
    const T* sp = static_cast<const T*>(S.p) + c0 * ssc + mf_off(S, i0, j * r + g.shift[1], k * r + g.shift[2]);
    const long long sy = S.n[0], sz = (long long)S.n[0] * S.n[1];
    if (sizeof(T) == 8 && r == 2) {
      // ratio 2: the 8 fine cells of every component summed in amrex_avgdown order (iref fastest, then jref, kref)
      const double* dpp = reinterpret_cast<const double*>(sp);
      if (((i0 - S.lo[0]) & 1) == 0 && (S.n[0] & 1) == 0) {        // 16-byte aligned pairs
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double2 a = *reinterpret_cast<const double2*>(dpp + c * ssc);
          const double2 b = *reinterpret_cast<const double2*>(dpp + c * ssc + sy);
          const double2 cc = *reinterpret_cast<const double2*>(dpp + c * ssc + sz);
          const double2 d = *reinterpret_cast<const double2*>(dpp + c * ssc + sz + sy);
          v[c] = (T)(((((((a.x + a.y) + b.x) + b.y) + cc.x) + cc.y) + d.x + d.y) * 0.125);
        }
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double* q = dpp + c * ssc;
          const double a0 = q[0], a1 = q[1], b0 = q[sy], b1 = q[sy + 1], c0_ = q[sz], c1 = q[sz + 1], d0 = q[sz + sy],
                       d1 = q[sz + sy + 1];
          v[c] = (T)(((((((a0 + a1) + b0) + b1) + c0_) + c1) + d0 + d1) * 0.125);
        }
      }
      return true;
    }
    const double w = 1.0 / (double)(r * r * r);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      T acc = 0;
      for (int kr = 0; kr < r; ++kr)          // amrex_avgdown order: iref fastest
        for (int jr = 0; jr < r; ++jr)
          for (int ir = 0; ir < r; ++ir) acc += sp[c * ssc + kr * sz + jr * sy + ir];
      v[c] = (T)(acc * w);
    }
    return true;
  }
  int si = i, sj = j, sk = k;
  if (g.kind == G_PC) { si = fdiv(i, g.ratio); sj = fdiv(j, g.ratio); sk = fdiv(k, g.ratio); }
  const T* sp = static_cast<const T*>(S.p) + c0 * ssc + mf_off(S, si + g.shift[0], sj + g.shift[1], sk + g.shift[2]);
#pragma unroll
  for (int c = 0; c < NC; ++c) v[c] = sp[c * ssc];
  return true;
}

// run-time component count: one component at a time through the same fetch
template <class T, bool ADD>
__device__ __forceinline__ void g_apply_rt(const GDesc& g, const DFabT* __restrict__ s0, const DFabT* __restrict__ s1,
                                           int i, int j, int k, T* __restrict__ dp, long long dsc, int ncomp) {
  for (int c = 0; c < ncomp; ++c) {
    T v[1];
    if (!g_fetch<T, 1>(g, s0, s1, i, j, k, c, v)) return;
    dp[c * dsc] = ADD ? (T)(dp[c * dsc] + v[0]) : v[0];
  }
}

// One CTA = 256 consecutive cells of one destination group's bounding box.  The group's descriptors pass through
// shared memory in chunks; each WARP first keeps the descriptors that meet the bounding box of its 32 cells (one
// descriptor per lane, one ballot per 32 descriptors), then every lane tests only those against its own cell: a
// coarse box under 27 fine boxes costs a lane 1-3 box tests instead of 27 or more.  COPY walks the list backwards (the
// last match wins) and stops at the first hit; ADD walks it forwards and sums in list order IN REGISTERS -- the
// destination is read once before the first contribution and written once at the end, so the result has the bits of
// the sequential read-modify-write loop (amrex::ParallelAdd order) at one RMW per cell.
// CPAR: component-parallel launch -- gridDim.x = tiles * ncomp and a thread handles ONE component of its cell (NC
// must be 1).  Used for the averaging plans, whose 60 loads per cell otherwise leave a 15-component thread
// latency-bound at low occupancy.
template <class T, bool ADD, int NC, bool CPAR>
__global__ void __launch_bounds__(MFT) k_plan_apply(const GDst* __restrict__ dsts, int ndst,
                                                    const GDesc* __restrict__ descs, const DFabT* __restrict__ dt,
                                                    const DFabT* __restrict__ s0, const DFabT* __restrict__ s1,
                                                    int ncomp) {
  __shared__ GDesc sd[PLAN_CHUNK];
  const int q = mf_fab_index();
  if (q >= ndst) return;                       // block-uniform
  const GDst D = dsts[q];
  const unsigned tile = CPAR ? blockIdx.x / (unsigned)ncomp : blockIdx.x;
  const int c0 = CPAR ? (int)(blockIdx.x % (unsigned)ncomp) : 0;
  const unsigned nx = D.bhi[0] - D.blo[0] + 1, ny = D.bhi[1] - D.blo[1] + 1, nz = D.bhi[2] - D.blo[2] + 1;
  const unsigned cells = nx * ny * nz;         // < 2^31 (checked at plan creation)
  if (tile * MFT >= cells) return;             // whole block idle: uniform exit
  unsigned t = tile * MFT + threadIdx.x;
  bool active = t < cells;
  const int i = D.blo[0] + (int)(t % nx);
  t /= nx;
  const int j = D.blo[1] + (int)(t % ny), k = D.blo[2] + (int)(t / ny);
  const DFabT F = dt[D.fab];
  if (!F.local) return;                        // block-uniform: a peer's box is filled by its owner
  const long long dsc = mf_stride(F);
  T* dp = static_cast<T*>(F.p) + c0 * dsc + (active ? mf_off(F, i, j, k) : 0);
  // bounding box of the warp's cells
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int BIG = 0x7fffffff;
  const int lane = threadIdx.x & 31;
  const int w0 = __reduce_min_sync(FULL, active ? i : BIG), w1 = __reduce_max_sync(FULL, active ? i : -BIG);
  const int w2 = __reduce_min_sync(FULL, active ? j : BIG), w3 = __reduce_max_sync(FULL, active ? j : -BIG);
  const int w4 = __reduce_min_sync(FULL, active ? k : BIG), w5 = __reduce_max_sync(FULL, active ? k : -BIG);
  T acc[NC > 0 ? NC : 1];
  bool touched = false;                        // ADD: acc holds the destination plus the contributions so far
  const int nchunks = (D.count + PLAN_CHUNK - 1) / PLAN_CHUNK;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int cb = ADD ? ch * PLAN_CHUNK : max(D.count - (ch + 1) * PLAN_CHUNK, 0);
    const int ce = ADD ? min(cb + PLAN_CHUNK, D.count) : D.count - ch * PLAN_CHUNK;
    const int n = ce - cb;
    {  // cooperative copy of n descriptors as 16-byte words
      const int4* src = reinterpret_cast<const int4*>(descs + D.first + cb);
      int4* dst = reinterpret_cast<int4*>(sd);
      for (int w = threadIdx.x; w < n * (int)(sizeof(GDesc) / 16); w += MFT) dst[w] = src[w];
    }
    __syncthreads();
    for (int part = 0; part < PLAN_CHUNK / 32; ++part) {
      const int base = ADD ? part * 32 : (PLAN_CHUNK / 32 - 1 - part) * 32;
      if (base >= n) continue;                 // block-uniform
      const int mine = base + lane;
      bool meets = false;
      if (mine < n) {
        const GDesc& g = sd[mine];
        meets = g.lo[0] <= w1 && g.hi[0] >= w0 && g.lo[1] <= w3 && g.hi[1] >= w2 && g.lo[2] <= w5 && g.hi[2] >= w4;
      }
      unsigned m = __ballot_sync(FULL, meets);
      while (m) {                              // warp-uniform loop
        const int bit = ADD ? __ffs(m) - 1 : 31 - __clz(m);
        m &= ~(1u << bit);
        if (!active) continue;
        const GDesc& g = sd[base + bit];
        if (i < g.lo[0] || i > g.hi[0] || j < g.lo[1] || j > g.hi[1] || k < g.lo[2] || k > g.hi[2]) continue;
        if (NC > 0) {
          T v[NC > 0 ? NC : 1];
          const bool has = g_fetch<T, (NC > 0 ? NC : 1)>(g, s0, s1, i, j, k, c0, v);
          if (ADD) {
            if (has) {
              if (!touched) {
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[c] = dp[c * dsc];
                touched = true;
              }
#pragma unroll
              for (int c = 0; c < NC; ++c) acc[c] = (T)(acc[c] + v[c]);
            }
          } else {
            if (has) {
#pragma unroll
              for (int c = 0; c < NC; ++c) dp[c * dsc] = v[c];
            }
            active = false;                    // the last matching descriptor wins, a G_NONE hit included
          }
        } else {
          g_apply_rt<T, ADD>(g, s0, s1, i, j, k, dp, dsc, ncomp);
          if (!ADD) active = false;
        }
      }
    }
    __syncthreads();
  }
  if (ADD && NC > 0 && touched) {
#pragma unroll
    for (int c = 0; c < NC; ++c) dp[c * dsc] = acc[c];
  }
}

}  // namespace lbx
