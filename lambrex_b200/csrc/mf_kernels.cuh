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
  const int b = mf_fab_index();
  if (b >= nfabs) return;
  const DFabT D = dt[b];
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
  const int b = mf_fab_index();
  if (b >= nfabs) return;
  const DFabT F = ft[b];
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
  const int b = mf_fab_index();
  if (b >= nfabs) return;
  const DFabT F = ft[b];
  int i, j, k;
  if (!mf_cell(F, grow_all, i, j, k)) return;
  const int g = grow_all - depth;   // cells inside valid grown by g are untouched
  if (i >= F.vlo[0] - g && i <= F.vhi[0] + g && j >= F.vlo[1] - g && j <= F.vhi[1] + g && k >= F.vlo[2] - g &&
      k <= F.vhi[2] + g)
    return;
  static_cast<double*>(F.p)[comp * mf_stride(F) + mf_off(F, i, j, k)] = 0.0;
}

template <class T>
__global__ void __launch_bounds__(MFT) k_mf_setval(const DFabT* __restrict__ ft, int nfabs, int grow_all, int ncomp,
                                                   T value) {
  const int b = mf_fab_index();
  if (b >= nfabs) return;
  const DFabT F = ft[b];
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
// ---------------------------------------------------------------------------
enum { G_COPY = 0, G_PC = 1, G_AVG = 2, G_CONST = 3 };
struct GDesc {
  int lo[3], hi[3];   // destination region (destination index space)
  int shift[3];       // added in SOURCE index space after the map
  int src_set, src_fab, kind, ratio;
  double value;
};
struct GDst {
  int fab, first, count, pad;
  int blo[3], bhi[3];   // bounding box of this fab's regions
};

__device__ __forceinline__ int fdiv(int a, int r) { return a >= 0 ? a / r : -((-a + r - 1) / r); }

template <class T>
__device__ __forceinline__ T g_value(const GDesc& g, const DFabT* s0, const DFabT* s1, int i, int j, int k, int c) {
  if (g.kind == G_CONST) return (T)g.value;
  const DFabT S = (g.src_set ? s1 : s0)[g.src_fab];
  const T* sp = static_cast<const T*>(S.p) + c * mf_stride(S);
  if (g.kind == G_COPY) return sp[mf_off(S, i + g.shift[0], j + g.shift[1], k + g.shift[2])];
  if (g.kind == G_PC)
    return sp[mf_off(S, fdiv(i, g.ratio) + g.shift[0], fdiv(j, g.ratio) + g.shift[1], fdiv(k, g.ratio) + g.shift[2])];
  // G_AVG
  const int r = g.ratio, i0 = i * r + g.shift[0], j0 = j * r + g.shift[1], k0 = k * r + g.shift[2];
  T acc = 0;
  for (int kr = 0; kr < r; ++kr)
    for (int jr = 0; jr < r; ++jr)
      for (int ir = 0; ir < r; ++ir) acc += sp[mf_off(S, i0 + ir, j0 + jr, k0 + kr)];
  return (T)(acc * (1.0 / (double)(r * r * r)));
}

template <class T, bool ADD>
__global__ void __launch_bounds__(MFT) k_plan_apply(const GDst* __restrict__ dsts, int ndst,
                                                    const GDesc* __restrict__ descs, const DFabT* __restrict__ dt,
                                                    const DFabT* __restrict__ s0, const DFabT* __restrict__ s1,
                                                    int ncomp) {
  const int q = mf_fab_index();
  if (q >= ndst) return;
  const GDst D = dsts[q];
  const int nx = D.bhi[0] - D.blo[0] + 1, ny = D.bhi[1] - D.blo[1] + 1, nz = D.bhi[2] - D.blo[2] + 1;
  long long t = (long long)blockIdx.x * MFT + threadIdx.x;
  if (t >= (long long)nx * ny * nz) return;
  const int i = D.blo[0] + (int)(t % nx);
  t /= nx;
  const int j = D.blo[1] + (int)(t % ny), k = D.blo[2] + (int)(t / ny);
  const DFabT F = dt[D.fab];
  T* dp = static_cast<T*>(F.p) + mf_off(F, i, j, k);
  const long long dsc = mf_stride(F);
  if (!ADD) {
    for (int d = D.count - 1; d >= 0; --d) {
      const GDesc& g = descs[D.first + d];
      if (i < g.lo[0] || i > g.hi[0] || j < g.lo[1] || j > g.hi[1] || k < g.lo[2] || k > g.hi[2]) continue;
      for (int c = 0; c < ncomp; ++c) dp[c * dsc] = g_value<T>(g, s0, s1, i, j, k, c);
      return;
    }
  } else {
    for (int d = 0; d < D.count; ++d) {
      const GDesc& g = descs[D.first + d];
      if (i < g.lo[0] || i > g.hi[0] || j < g.lo[1] || j > g.hi[1] || k < g.lo[2] || k > g.hi[2]) continue;
      for (int c = 0; c < ncomp; ++c) dp[c * dsc] += g_value<T>(g, s0, s1, i, j, k, c);
    }
  }
}

}  // namespace lbx
