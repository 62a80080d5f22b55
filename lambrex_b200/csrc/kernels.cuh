// Device kernels of the D3Q15 collide-and-stream path (sm_100a).
//
// Data layout in HBM: one "fab" = a rectangular box of cells (valid region plus
// ghosts), SoA: component planes slowest, then z, y, x fastest -- 15 fp64
// population planes per box.  One thread per cell, threads consecutive along x
// so every plane access of a warp is one contiguous 256 B segment.
//
// Kernel            replaces (reference file:line)                      HBM bytes / cell
// k_collide_stream  CollideLevel + Stream + UpdateNow                    240 (15 ld + 15 st)
//                   (src/AmrSim.cpp:124-135, 109-122; AmrSim.h:89-94)
// k_collide         Collide / CoarseCollide (src/AmrSim.cpp:25-107,      240 (in place)
//                   487-580)
// k_stream          Stream + PropagatePoint (src/AmrSim.cpp:109-122,     240
//                   include/component.h:22-29)
// k_moments         CalcHydroVars (src/AmrSim.cpp:938-979)               152 (15 ld + 4 st)
// k_equilibrium     CalcEquilibriumDist (src/AmrSim.cpp:845-931)         152 (4 ld + 15 st)
//
// This header is compiled twice: kernels_fast.cu (CollideFast, FMA allowed) and
// kernels_literal.cu (CollideLiteral, --fmad=false).
#pragma once
#include <cuda_runtime.h>
#include "d3q15.cuh"

namespace lbx {

struct DFab {          // device view of one fab
  double* p;           // component 0, cell (lo[0],lo[1],lo[2])
  int lo[3];           // lower corner of the ALLOCATED box (valid lo - ghosts)
  int n[3];            // allocated extents
};
struct DBox { int lo[3], hi[3]; };               // inclusive cell range
struct DDom { int lo[3], hi[3], periodic[3]; };  // index domain and wrap flags

__device__ __forceinline__ long long plane_stride(const DFab& f) {
  return (long long)f.n[0] * f.n[1] * f.n[2];
}
// offset of row (j,k) such that element (i,j,k) lives at row_off + i
__device__ __forceinline__ long long row_off(const DFab& f, int j, int k) {
  return (long long)f.n[0] * ((j - f.lo[1]) + (long long)f.n[1] * (k - f.lo[2])) - f.lo[0];
}

constexpr int BX = 128;   // threads per CTA along x

// ---------------------------------------------------------------------------
// Fused collide + stream, one launch = one full reference time step on the box:
//   PUSH=true : read own cell (aligned), collide, scatter f_p to x + c_p   [F <- S(C(F))]
//   PUSH=false: gather f_p from x - c_p, collide, write own cell (aligned) [G <- C(S(G))]
// Directions flagged periodic wrap inside the kernel (no ghost cells needed);
// other directions index straight into ghost planes of the fab.
// ---------------------------------------------------------------------------
template <class C, bool PUSH>
__global__ void __launch_bounds__(BX) k_collide_stream(DFab src, DFab dst, DBox box, DDom dom,
                                                       double omega_s, double omega_b) {
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = box.lo[1] + blockIdx.y;
  const int k = box.lo[2] + blockIdx.z;
  if (i > box.hi[0]) return;

  int ip = i + 1, im = i - 1, jp = j + 1, jm = j - 1, kp = k + 1, km = k - 1;
  if (dom.periodic[0]) { if (ip > dom.hi[0]) ip = dom.lo[0]; if (im < dom.lo[0]) im = dom.hi[0]; }
  if (dom.periodic[1]) { if (jp > dom.hi[1]) jp = dom.lo[1]; if (jm < dom.lo[1]) jm = dom.hi[1]; }
  if (dom.periodic[2]) { if (kp > dom.hi[2]) kp = dom.lo[2]; if (km < dom.lo[2]) km = dom.hi[2]; }

  const DFab& nb = PUSH ? dst : src;   // the fab accessed at neighbour cells
  const DFab& own = PUSH ? src : dst;  // the fab accessed at (i,j,k)
  const long long nsc = plane_stride(nb), osc = plane_stride(own);
  // PUSH: f_p goes to x + c_p ; PULL: f_p comes from x - c_p
  const int ixp = PUSH ? ip : im, ixm = PUSH ? im : ip;   // x index for c_x = +1 / -1
  const int jyp = PUSH ? jp : jm, jym = PUSH ? jm : jp;
  const int kzp = PUSH ? kp : km, kzm = PUSH ? km : kp;
  const long long r00 = row_off(nb, j, k);
  const long long rp0 = row_off(nb, jyp, k), rm0 = row_off(nb, jym, k);
  const long long r0p = row_off(nb, j, kzp), r0m = row_off(nb, j, kzm);
  const long long rpp = row_off(nb, jyp, kzp), rpm = row_off(nb, jyp, kzm);
  const long long rmp = row_off(nb, jym, kzp), rmm = row_off(nb, jym, kzm);
  long long off[NV];
  off[0] = r00 + i;
  off[1] = r00 + ixp;  off[2] = r00 + ixm;
  off[3] = rp0 + i;    off[4] = rm0 + i;
  off[5] = r0p + i;    off[6] = r0m + i;
  off[7] = rpp + ixp;  off[8] = rpm + ixp;  off[9] = rmp + ixp;  off[10] = rmm + ixp;
  off[11] = rpp + ixm; off[12] = rpm + ixm; off[13] = rmp + ixm; off[14] = rmm + ixm;

  const long long o = row_off(own, j, k) + i;
  double f[NV];
  if (PUSH) {
#pragma unroll
    for (int p = 0; p < NV; ++p) f[p] = __ldcs(src.p + p * osc + o);
  } else {
#pragma unroll
    for (int p = 0; p < NV; ++p) f[p] = __ldcs(src.p + p * nsc + off[p]);
  }
  C::collide(f, omega_s, omega_b);
  if (PUSH) {
#pragma unroll
    for (int p = 0; p < NV; ++p) __stcs(dst.p + p * nsc + off[p], f[p]);
  } else {
#pragma unroll
    for (int p = 0; p < NV; ++p) __stcs(dst.p + p * osc + o, f[p]);
  }
}

// ---------------------------------------------------------------------------
// Fused collide + stream + z-face exchange for the slab-decomposed uniform path
// (one z-slab per GPU).  Push form.  Cells are addressed by GLOBAL
// k; a destination plane that is not part of this rank's fab (k+1 above the slab,
// k-1 below it, after the periodic wrap of the global domain) lives in the
// neighbour's fab, reached through a CUDA-IPC peer pointer: the 5 populations
// crossing the face are stored straight into the neighbour's HBM over NVLink by
// the boundary-plane CTAs while the interior CTAs keep the local HBM busy.  This
// replaces CollideLevel + FillBoundary + Stream + UpdateNow
// (src/AmrSim.cpp:124-135, 109-122; include/AmrSim.h:89-94).  With up = dn = dst
// (one GPU) it degenerates to k_collide_stream<.., true>.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool has_plane(const DFab& f, int k) {
  return k >= f.lo[2] && k < f.lo[2] + f.n[2];
}

template <class C>
__global__ void __launch_bounds__(BX) k_collide_stream_slab(DFab src, DFab dst, DFab dn, DFab up, DBox box,
                                                            DDom dom, double omega_s, double omega_b) {
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = box.lo[1] + blockIdx.y;
  const int k = box.lo[2] + blockIdx.z;
  if (i > box.hi[0]) return;

  int ip = i + 1, im = i - 1, jp = j + 1, jm = j - 1, kp = k + 1, km = k - 1;
  if (dom.periodic[0]) { if (ip > dom.hi[0]) ip = dom.lo[0]; if (im < dom.lo[0]) im = dom.hi[0]; }
  if (dom.periodic[1]) { if (jp > dom.hi[1]) jp = dom.lo[1]; if (jm < dom.lo[1]) jm = dom.hi[1]; }
  if (dom.periodic[2]) { if (kp > dom.hi[2]) kp = dom.lo[2]; if (km < dom.lo[2]) km = dom.hi[2]; }

  // block-uniform choice of the fab that owns plane k+1 / k-1
  const DFab& fp = has_plane(dst, kp) ? dst : up;
  const DFab& fm = has_plane(dst, km) ? dst : dn;
  const long long sc0 = plane_stride(dst), scp = plane_stride(fp), scm = plane_stride(fm);
  double* const b0 = dst.p;
  double* const bp = fp.p;
  double* const bm = fm.p;
  const long long r00 = row_off(dst, j, k), rp0 = row_off(dst, jp, k), rm0 = row_off(dst, jm, k);
  const long long r0p = row_off(fp, j, kp), rpp = row_off(fp, jp, kp), rmp = row_off(fp, jm, kp);
  const long long r0m = row_off(fm, j, km), rpm = row_off(fm, jp, km), rmm = row_off(fm, jm, km);

  const long long ssc = plane_stride(src), o = row_off(src, j, k) + i;
  double f[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = __ldcs(src.p + p * ssc + o);
  C::collide(f, omega_s, omega_b);
  // same-plane populations (c_z = 0)
  __stcs(b0 + 0 * sc0 + r00 + i, f[0]);
  __stcs(b0 + 1 * sc0 + r00 + ip, f[1]);
  __stcs(b0 + 2 * sc0 + r00 + im, f[2]);
  __stcs(b0 + 3 * sc0 + rp0 + i, f[3]);
  __stcs(b0 + 4 * sc0 + rm0 + i, f[4]);
  // c_z = +1 : plane k+1 (local or the upper neighbour's fab)
  __stcs(bp + 5 * scp + r0p + i, f[5]);
  __stcs(bp + 7 * scp + rpp + ip, f[7]);
  __stcs(bp + 9 * scp + rmp + ip, f[9]);
  __stcs(bp + 11 * scp + rpp + im, f[11]);
  __stcs(bp + 13 * scp + rmp + im, f[13]);
  // c_z = -1 : plane k-1 (local or the lower neighbour's fab)
  __stcs(bm + 6 * scm + r0m + i, f[6]);
  __stcs(bm + 8 * scm + rpm + ip, f[8]);
  __stcs(bm + 10 * scm + rmm + ip, f[10]);
  __stcs(bm + 12 * scm + rpm + im, f[12]);
  __stcs(bm + 14 * scm + rmm + im, f[14]);
}

// ---------------------------------------------------------------------------
// Collide valid cells (src -> dst, may alias).  Optional int mask (1 comp, any
// ghost width): cells with mask == fine_val get all 15 populations zeroed
// (CoarseCollide, src/AmrSim.cpp:499-500).
// ---------------------------------------------------------------------------
struct DMask { const int* p; int lo[3]; int n[3]; };

template <class C>
__global__ void __launch_bounds__(BX) k_collide(DFab src, DFab dst, DBox box, double omega_s,
                                                double omega_b, DMask mask, int fine_val) {
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = box.lo[1] + blockIdx.y;
  const int k = box.lo[2] + blockIdx.z;
  if (i > box.hi[0]) return;
  const long long ssc = plane_stride(src), dsc = plane_stride(dst);
  const long long so = row_off(src, j, k) + i, dp = row_off(dst, j, k) + i;
  double f[NV];
  bool zero = false;
  if (mask.p) {
    const long long mo = (long long)mask.n[0] * ((j - mask.lo[1]) + (long long)mask.n[1] * (k - mask.lo[2])) +
                         (i - mask.lo[0]);
    zero = (mask.p[mo] == fine_val);
  }
  if (zero) {
#pragma unroll
    for (int p = 0; p < NV; ++p) dst.p[p * dsc + dp] = 0.0;
    return;
  }
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = src.p[p * ssc + so];
  C::collide(f, omega_s, omega_b);
#pragma unroll
  for (int p = 0; p < NV; ++p) dst.p[p * dsc + dp] = f[p];
}

// ---------------------------------------------------------------------------
// Pull streaming dst(x,p) = src(x - c_p, p) for x in box; wraps in periodic
// directions, otherwise reads ghost cells (box = valid grown by 1 on the AMR
// path, src ghosts 2 deep).
// ---------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(BX) k_stream(DFab src, DFab dst, DBox box, DDom dom) {
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = box.lo[1] + blockIdx.y;
  const int k = box.lo[2] + blockIdx.z;
  if (i > box.hi[0]) return;
  int ip = i + 1, im = i - 1, jp = j + 1, jm = j - 1, kp = k + 1, km = k - 1;
  if (dom.periodic[0]) { if (ip > dom.hi[0]) ip = dom.lo[0]; if (im < dom.lo[0]) im = dom.hi[0]; }
  if (dom.periodic[1]) { if (jp > dom.hi[1]) jp = dom.lo[1]; if (jm < dom.lo[1]) jm = dom.hi[1]; }
  if (dom.periodic[2]) { if (kp > dom.hi[2]) kp = dom.lo[2]; if (km < dom.lo[2]) km = dom.hi[2]; }
  const long long ssc = plane_stride(src), dsc = plane_stride(dst);
  const long long d = row_off(dst, j, k) + i;
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    const int ii = cx(p) > 0 ? im : cx(p) < 0 ? ip : i;
    const int jj = cy(p) > 0 ? jm : cy(p) < 0 ? jp : j;
    const int kk = cz(p) > 0 ? km : cz(p) < 0 ? kp : k;
    dst.p[p * dsc + d] = src.p[p * ssc + row_off(src, jj, kk) + ii];
  }
}

// ---------------------------------------------------------------------------
// rho = sum_p f_p ; u = sum_p c_p f_p / rho over box.  PULL=true evaluates the
// moments of the streamed field S(f) without materialising it.
// ---------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(BX) k_moments(DFab f_, DFab rho_, DFab u_, DBox box) {
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = box.lo[1] + blockIdx.y;
  const int k = box.lo[2] + blockIdx.z;
  if (i > box.hi[0]) return;
  const long long fsc = plane_stride(f_), usc = plane_stride(u_);
  const long long fo = row_off(f_, j, k) + i;
  double f[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = f_.p[p * fsc + fo];
  double rho, ux, uy, uz;
  C::moments(f, rho, ux, uy, uz);
  rho_.p[row_off(rho_, j, k) + i] = rho;
  const long long uo = row_off(u_, j, k) + i;
  u_.p[uo] = ux;
  u_.p[usc + uo] = uy;
  u_.p[2 * usc + uo] = uz;
}

template <class C>
__global__ void __launch_bounds__(BX) k_equilibrium(DFab f_, DFab rho_, DFab u_, DBox box) {
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = box.lo[1] + blockIdx.y;
  const int k = box.lo[2] + blockIdx.z;
  if (i > box.hi[0]) return;
  const long long fsc = plane_stride(f_), usc = plane_stride(u_);
  const long long uo = row_off(u_, j, k) + i;
  double f[NV];
  equilibrium_cell(rho_.p[row_off(rho_, j, k) + i], u_.p[uo], u_.p[usc + uo], u_.p[2 * usc + uo], f);
  const long long fo = row_off(f_, j, k) + i;
#pragma unroll
  for (int p = 0; p < NV; ++p) f_.p[p * fsc + fo] = f[p];
}

inline dim3 grid_for(const DBox& b) {
  return dim3((unsigned)((b.hi[0] - b.lo[0] + BX) / BX), (unsigned)(b.hi[1] - b.lo[1] + 1),
              (unsigned)(b.hi[2] - b.lo[2] + 1));
}

}  // namespace lbx
