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
// opposite population: c_opp(p) = -c_p (1<->2, 3<->4, 5<->6, 7<->14, 8<->13, 9<->12, 10<->11)
__host__ __device__ constexpr int opp(int p) { return p == 0 ? 0 : p < 7 ? (p & 1 ? p + 1 : p - 1) : 21 - p; }

// dom.periodic[d]: 1 = periodic wrap, 0 = the fab carries ghost cells in that direction, 2 = solid no-slip WALLS at both
// domain faces (half-way bounce-back; addition -- the reference's constructor aborts on non-periodic directions,
// src/AmrSim.cpp:788-797, and leaves DistFnFillShim :346-357 as the hook): a population that would leave the domain
// through a wall returns to its own cell with the opposite velocity, f'(x, opp(p)) = f*(x, p).  WALLS selects the code
// path at compile time so that the periodic kernel is unchanged.
template <class C, bool PUSH, bool WALLS = false>
__global__ void __launch_bounds__(BX) k_collide_stream(DFab src, DFab dst, DBox box, DDom dom,
                                                       double omega_s, double omega_b) {
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = box.lo[1] + blockIdx.y;
  const int k = box.lo[2] + blockIdx.z;
  if (i > box.hi[0]) return;

  int ip = i + 1, im = i - 1, jp = j + 1, jm = j - 1, kp = k + 1, km = k - 1;
  if (dom.periodic[0] == 1) { if (ip > dom.hi[0]) ip = dom.lo[0]; if (im < dom.lo[0]) im = dom.hi[0]; }
  if (dom.periodic[1] == 1) { if (jp > dom.hi[1]) jp = dom.lo[1]; if (jm < dom.lo[1]) jm = dom.hi[1]; }
  if (dom.periodic[2] == 1) { if (kp > dom.hi[2]) kp = dom.lo[2]; if (km < dom.lo[2]) km = dom.hi[2]; }

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
    if (WALLS) {
      // which faces of this cell are walls
      const bool xl = dom.periodic[0] == 2 && i == dom.lo[0], xh = dom.periodic[0] == 2 && i == dom.hi[0];
      const bool yl = dom.periodic[1] == 2 && j == dom.lo[1], yh = dom.periodic[1] == 2 && j == dom.hi[1];
      const bool zl = dom.periodic[2] == 2 && k == dom.lo[2], zh = dom.periodic[2] == 2 && k == dom.hi[2];
      const long long oo = row_off(dst, j, k) + i;
#pragma unroll
      for (int p = 0; p < NV; ++p) {
        const bool out = (cx(p) > 0 && xh) || (cx(p) < 0 && xl) || (cy(p) > 0 && yh) || (cy(p) < 0 && yl) ||
                         (cz(p) > 0 && zh) || (cz(p) < 0 && zl);
        __stcs(dst.p + (out ? opp(p) * nsc + oo : p * nsc + off[p]), f[p]);
      }
    } else {
#pragma unroll
      for (int p = 0; p < NV; ++p) __stcs(dst.p + p * nsc + off[p], f[p]);
    }
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
  // 15 destination addresses up front (like k_collide_stream's off[]): the register
  // footprint this costs caps residency at 6 CTAs/SM, which measured FASTER than the
  // 10 CTAs/SM of an address-on-the-fly formulation (profiles/r01_occupancy_sweep.md)
  double* d[NV];
  {
    const long long r00 = row_off(dst, j, k), rp0 = row_off(dst, jp, k), rm0 = row_off(dst, jm, k);
    d[0] = dst.p + r00 + i;
    d[1] = dst.p + 1 * sc0 + r00 + ip;
    d[2] = dst.p + 2 * sc0 + r00 + im;
    d[3] = dst.p + 3 * sc0 + rp0 + i;
    d[4] = dst.p + 4 * sc0 + rm0 + i;
    // c_z = +1 : plane k+1 (local or the upper neighbour's fab)
    const long long r0p = row_off(fp, j, kp), rpp = row_off(fp, jp, kp), rmp = row_off(fp, jm, kp);
    d[5] = fp.p + 5 * scp + r0p + i;
    d[7] = fp.p + 7 * scp + rpp + ip;
    d[9] = fp.p + 9 * scp + rmp + ip;
    d[11] = fp.p + 11 * scp + rpp + im;
    d[13] = fp.p + 13 * scp + rmp + im;
    // c_z = -1 : plane k-1 (local or the lower neighbour's fab)
    const long long r0m = row_off(fm, j, km), rpm = row_off(fm, jp, km), rmm = row_off(fm, jm, km);
    d[6] = fm.p + 6 * scm + r0m + i;
    d[8] = fm.p + 8 * scm + rpm + ip;
    d[10] = fm.p + 10 * scm + rmm + ip;
    d[12] = fm.p + 12 * scm + rpm + im;
    d[14] = fm.p + 14 * scm + rmm + im;
  }
  const long long ssc = plane_stride(src), o = row_off(src, j, k) + i;
  double f[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = __ldcs(src.p + p * ssc + o);
  C::collide(f, omega_s, omega_b);
#pragma unroll
  for (int p = 0; p < NV; ++p) __stcs(d[p], f[p]);
}

// ---------------------------------------------------------------------------
// The same step with the cross-GPU ORDERING folded in: ONE launch per time step per rank, no separate
// wait / signal kernels (lbx_mf_collide_stream_slab).  The two boundary planes of the slab are mapped to
// blockIdx.z = 0, 1 so that their CTAs are dispatched first:
//   * a boundary CTA first waits (thread 0 spins with acquire loads at system scope, wall-clock timeout)
//     until both neighbours have published `wait_value` -- their boundary planes of the previous step,
//     which stored 5 populations per face cell into THIS rank's source fab, are complete (RAW), and they
//     no longer read the planes of their fab that this step overwrites (WAR of the ping-pong);
//   * after its stores (5 populations per face cell go straight into the neighbour's fab over NVLink) every
//     thread fences at system scope; the last boundary CTA to finish publishes `sig_value` into the
//     neighbours' flags with a release store.  Interior CTAs never touch a flag: they keep HBM busy while
//     the boundary CTAs wait, store remotely and signal.
// The neighbours' progress never depends on this launch (they wait for the PREVIOUS signal), so spinning
// boundary CTAs cannot deadlock the grid.
// ---------------------------------------------------------------------------
struct SlabSync {
  const unsigned long long* wait_a;   // this rank's flags: written by the rank below / above (null: no wait)
  const unsigned long long* wait_b;
  unsigned long long wait_value;
  unsigned long long* sig_a;          // the rank below's "above" flag, the rank above's "below" flag (peer memory)
  unsigned long long* sig_b;
  unsigned long long sig_value;
  unsigned long long* counter;        // boundary CTAs that have finished (local; reset by the last one)
  unsigned long long nboundary;       // CTAs of the boundary planes
  unsigned long long timeout_ns;
  int* err;                           // host-mapped: set on timeout (sticky)
  int ilv_shift;                      // every 2^ilv_shift-th row at the start of the grid is a boundary row (0: all first)
};

// one cell of the slab step (the body of k_collide_stream_slab), shared by the two code paths below
template <class C>
__device__ __forceinline__ void slab_cell(const DFab& src, const DFab& dst, const DFab& dn, const DFab& up, const DBox& box,
                                          const DDom& dom, double omega_s, double omega_b, int i, int j, int k) {
  int ip = i + 1, im = i - 1, jp = j + 1, jm = j - 1, kp = k + 1, km = k - 1;
  if (dom.periodic[0]) { if (ip > dom.hi[0]) ip = dom.lo[0]; if (im < dom.lo[0]) im = dom.hi[0]; }
  if (dom.periodic[1]) { if (jp > dom.hi[1]) jp = dom.lo[1]; if (jm < dom.lo[1]) jm = dom.hi[1]; }
  if (dom.periodic[2]) { if (kp > dom.hi[2]) kp = dom.lo[2]; if (km < dom.lo[2]) km = dom.hi[2]; }
  const DFab& fp = has_plane(dst, kp) ? dst : up;
  const DFab& fm = has_plane(dst, km) ? dst : dn;
  const long long sc0 = plane_stride(dst), scp = plane_stride(fp), scm = plane_stride(fm);
  double* d[NV];
  {
    const long long r00 = row_off(dst, j, k), rp0 = row_off(dst, jp, k), rm0 = row_off(dst, jm, k);
    d[0] = dst.p + r00 + i;
    d[1] = dst.p + 1 * sc0 + r00 + ip;
    d[2] = dst.p + 2 * sc0 + r00 + im;
    d[3] = dst.p + 3 * sc0 + rp0 + i;
    d[4] = dst.p + 4 * sc0 + rm0 + i;
    const long long r0p = row_off(fp, j, kp), rpp = row_off(fp, jp, kp), rmp = row_off(fp, jm, kp);
    d[5] = fp.p + 5 * scp + r0p + i;
    d[7] = fp.p + 7 * scp + rpp + ip;
    d[9] = fp.p + 9 * scp + rmp + ip;
    d[11] = fp.p + 11 * scp + rpp + im;
    d[13] = fp.p + 13 * scp + rmp + im;
    const long long r0m = row_off(fm, j, km), rpm = row_off(fm, jp, km), rmm = row_off(fm, jm, km);
    d[6] = fm.p + 6 * scm + r0m + i;
    d[8] = fm.p + 8 * scm + rpm + ip;
    d[10] = fm.p + 10 * scm + rmm + ip;
    d[12] = fm.p + 12 * scm + rpm + im;
    d[14] = fm.p + 14 * scm + rmm + im;
  }
  const long long ssc = plane_stride(src), o = row_off(src, j, k) + i;
  double f[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = __ldcs(src.p + p * ssc + o);
  C::collide(f, omega_s, omega_b);
#pragma unroll
  for (int p = 0; p < NV; ++p) __stcs(d[p], f[p]);
}

template <class C>
__global__ void __launch_bounds__(BX) k_collide_stream_slab_sync(DFab src, DFab dst, DFab dn, DFab up, DBox box,
                                                                 DDom dom, double omega_s, double omega_b, SlabSync sy) {
  // Row order of the launch (rows = (j, k) pairs, dispatched in blockIdx.y + ny * blockIdx.z order): the rows of the two
  // boundary planes are INTERLEAVED with interior rows, one in every 2^ilv_shift, over about the first half of the
  // grid instead of all coming first.  Their CTAs store 5 populations per cell over NVLink, which moves those 84 MB
  // (1024^2 planes) at ~300 GB/s: bunched together they hold every CTA slot of the GPU while the link drains and the
  // local HBM idles (0.3 ms per step, profiles/r02_scale.md); spread thin, the link runs at a few per cent duty
  // beside the bandwidth-bound interior rows.  The signal leaves at about half of the step.
  const int ny = box.hi[1] - box.lo[1] + 1, nzl = box.hi[2] - box.lo[2] + 1;
  const int nb = (nzl < 2 ? nzl : 2) * ny;                  // boundary rows
  const int r = (int)blockIdx.y + ny * (int)blockIdx.z;     // position in dispatch order
  const int i = box.lo[0] + blockIdx.x * BX + threadIdx.x;
  int rb = -1, ri;
  if (nzl <= 2) rb = r;                                     // nothing but boundary rows
  else if (sy.ilv_shift == 0) { if (r < nb) rb = r; else ri = r - nb; }   // thin slab: too few interior rows to interleave
  else if (r < (nb << sy.ilv_shift)) {
    if ((r & ((1 << sy.ilv_shift) - 1)) == 0) rb = r >> sy.ilv_shift;
    else ri = r - (r >> sy.ilv_shift) - 1;
  } else ri = r - nb;
  if (rb < 0) {
    // interior planes: exactly the plain slab step (own copy of the code, so that the synchronisation below does not
    // cost the bandwidth-bound path its uniform-datapath address arithmetic -- measured 6 % otherwise)
    if (i > box.hi[0]) return;
    const int kq = ri / ny;
    slab_cell<C>(src, dst, dn, up, box, dom, omega_s, omega_b, i, box.lo[1] + (ri - kq * ny), box.lo[2] + 1 + kq);
    return;
  }
  const int kb = rb / ny;
  const int j = box.lo[1] + (rb - kb * ny);
  const int k = kb == 0 ? box.lo[2] : box.hi[2];
  if (sy.wait_a) {
    if (threadIdx.x == 0) {     // (never READ the host-mapped error word here: a zero-copy read costs a PCIe round trip per CTA)
      unsigned long long t0, t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
      for (int which = 0; which < 2; ++which) {
        const unsigned long long* fl = which ? sy.wait_b : sy.wait_a;
        for (;;) {
          unsigned long long v;
          asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(fl) : "memory");
          if (v >= sy.wait_value) break;
          asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
          if (t - t0 > sy.timeout_ns) { *sy.err = 1; __threadfence_system(); which = 2; break; }
          __nanosleep(100);
        }
      }
    }
    __syncthreads();
  }
  if (i <= box.hi[0]) slab_cell<C>(src, dst, dn, up, box, dom, omega_s, omega_b, i, j, k);
  if (sy.sig_a) {
    // the CTA's stores (local and remote) are ordered before its arrival below: barrier, then ONE system-scope fence
    // by the arriving thread, which is cumulative over the writes it observes through the barrier (the pattern of
    // cooperative-groups grid synchronisation); a fence per thread made every boundary CTA wait out 128 of them
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned long long done = atomicAdd(sy.counter, 1ull) + 1ull;
      if (done == sy.nboundary) {      // the last boundary CTA of this launch: every face store has been fenced
        *sy.counter = 0ull;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sy.sig_a), "l"(sy.sig_value) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sy.sig_b), "l"(sy.sig_value) : "memory");
      }
    }
  }
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

// ===========================================================================
// Batched per-level kernels (AMR path): a level's field is ONE device allocation
// holding all its boxes (each grown by its ghosts) plus a device table of DFabT
// descriptors; one launch covers every box of the level.
// grid = (tiles of MFT cells, fab index); a tile is MFT consecutive cells (x
// fastest) of the fab's operating region = valid box grown by `grow`.
// ===========================================================================
struct DFabT {
  void* p;           // component 0 of the allocated box (a CUDA-IPC peer pointer when !local)
  int lo[3], n[3];   // allocated box: lower corner, extents
  int vlo[3], vhi[3];  // valid box (inclusive)
  int local;         // 1: this rank owns the box; 0: it lives in a peer's HBM (read-only here)
  int lid;           // entry q of a table names the q-th box THIS RANK owns (a property of the table, not of box q)
};
static_assert(sizeof(DFabT) == 64, "DFabT is read as four 16-byte words");
constexpr int MFT = 256;

__device__ __forceinline__ int mf_fab_index() { return blockIdx.y + gridDim.y * blockIdx.z; }
// Launches over "the boxes of a fab set" cover the boxes this rank owns only (`nlocal` of them): on 8 ranks a launch
// over all boxes would start seven CTAs that exit for every one that works.  -1: past the end of the launch.
__device__ __forceinline__ int mf_local_fab(const DFabT* __restrict__ t, int nlocal) {
  const int q = mf_fab_index();
  return q < nlocal ? t[q].lid : -1;
}

__device__ __forceinline__ bool mf_cell(const DFabT& f, int grow, int& i, int& j, int& k) {
  const int nx = f.vhi[0] - f.vlo[0] + 1 + 2 * grow, ny = f.vhi[1] - f.vlo[1] + 1 + 2 * grow,
            nz = f.vhi[2] - f.vlo[2] + 1 + 2 * grow;
  unsigned t = blockIdx.x * MFT + threadIdx.x;      // a fab holds < 2^31 cells (lbx_mf_create)
  if (t >= (unsigned)nx * ny * nz) return false;
  i = f.vlo[0] - grow + (int)(t % (unsigned)nx);
  t /= (unsigned)nx;
  j = f.vlo[1] - grow + (int)(t % (unsigned)ny);
  k = f.vlo[2] - grow + (int)(t / (unsigned)ny);
  return true;
}
__device__ __forceinline__ long long mf_stride(const DFabT& f) { return (long long)f.n[0] * f.n[1] * f.n[2]; }
__device__ __forceinline__ long long mf_off(const DFabT& f, int i, int j, int k) {
  return (i - f.lo[0]) + (long long)f.n[0] * ((j - f.lo[1]) + (long long)f.n[1] * (k - f.lo[2]));
}
__device__ __forceinline__ bool mf_in_valid(const DFabT& f, int i, int j, int k) {
  return i >= f.vlo[0] && i <= f.vhi[0] && j >= f.vlo[1] && j <= f.vhi[1] && k >= f.vlo[2] && k <= f.vhi[2];
}

// Collide / CoarseCollide / FineCollide on the valid cells of every box: dst <- collide(src)
// (src/AmrSim.cpp:25-107, 487-590; in place when src == dst).  mask != null: cells with
// mask == fine_val are zeroed.
template <class C>
__global__ void __launch_bounds__(MFT) k_mf_collide(const DFabT* __restrict__ st, const DFabT* __restrict__ ft,
                                                    const DFabT* __restrict__ mt, int nfabs, double omega_s,
                                                    double omega_b, int fine_val) {
  const int b = mf_local_fab(ft, nfabs);
  if (b < 0) return;
  const DFabT F = ft[b];
  if (!F.local) return;
  int i, j, k;
  if (!mf_cell(F, 0, i, j, k)) return;
  double* fp = static_cast<double*>(F.p) + mf_off(F, i, j, k);
  const long long sc = mf_stride(F);
  if (mt) {
    const DFabT M = mt[b];
    if (static_cast<const int*>(M.p)[mf_off(M, i, j, k)] == fine_val) {
#pragma unroll
      for (int p = 0; p < NV; ++p) fp[p * sc] = 0.0;
      return;
    }
  }
  const DFabT S = st[b];      // source set: same boxes, may be the destination itself
  const double* sp = static_cast<const double*>(S.p) + mf_off(S, i, j, k);
  const long long ssc = mf_stride(S);
  double f[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = sp[p * ssc];
  C::collide(f, omega_s, omega_b);
#pragma unroll
  for (int p = 0; p < NV; ++p) fp[p * sc] = f[p];
}

// Fused Collide/CoarseCollide/FineCollide + Stream of the Rohde cycle on every box of a level
// (src/AmrSim.cpp:487-590 then :109-122), push form, ONE pass over the level instead of two:
//   thread = one SOURCE cell x of the box grown by its 2 ghost rings;
//   valid cell : f <- vt(x), zeroed where mask == fine_val (CoarseCollide, :499-500), else
//                collided;
//   ghost cell : f <- gt(x) UNcollided -- the reference never refreshes ghosts between its
//                collide and its Stream, so Stream pulls FillPatch values there;
//   push       : dst(x + c_p, p) = f_p for destinations inside valid grown by 1 (Stream's
//                iteration space); ring 2 of dst is never written by Stream and holds the fresh
//                fab's fill, 0: the ring-2 thread zeroes its own cell.
// zero_invalid != 0 folds the ZeroInvalidComponents that follows the last Stream of a cycle
// (:604-617) into the stores: a ghost destination y keeps component p only if y - 2 c_p is a
// valid cell; y - 2 c_p = x - c_p.  (Ring 2 is all zeros already, so nothing else remains.)
// The ghost cells are pushed from their own values or, with DistFnFillPatch folded in, from
// wherever FillPatch takes them (k_mf_collide_stream below), loading only the populations that
// have a destination.
// The shell INCLUDES the row-alignment cells of the allocated box in x (LBX_OPT_ALIGN_ROWS: lead-in / tail
// cells beyond the 2 ghost cells): they behave like ghost rings >= 2 -- nothing is pushed from them and
// they zero themselves -- so that a pass writes every sector of the destination completely.
__device__ __forceinline__ bool mf_shell_cell(const DFabT& f, unsigned t, int& i, int& j, int& k) {
  constexpr unsigned h = HALO;
  const unsigned v0 = f.vhi[0] - f.vlo[0] + 1, v1 = f.vhi[1] - f.vlo[1] + 1, v2 = f.vhi[2] - f.vlo[2] + 1;
  const unsigned n0 = f.n[0], n1 = v1 + 2 * h;                           // x: the whole allocated row
  const unsigned hl = f.vlo[0] - f.lo[0], hr = n0 - v0 - hl;             // cells left / right of the valid row
  const unsigned A = n0 * n1 * h, B = n0 * h * v2, Cl = hl * v1 * v2, Cr = hr * v1 * v2;   // a fab holds < 2^31 cells
  if (t < 2 * A) {                      // z-low / z-high slabs: full x-y planes
    const unsigned s = t >= A, r = t - s * A;
    i = f.lo[0] + (int)(r % n0);
    j = f.vlo[1] - (int)h + (int)((r / n0) % n1);
    k = (s ? f.vhi[2] + 1 : f.vlo[2] - (int)h) + (int)(r / (n0 * n1));
    return true;
  }
  t -= 2 * A;
  if (t < 2 * B) {                      // y-low / y-high slabs over the valid z range
    const unsigned s = t >= B, r = t - s * B;
    i = f.lo[0] + (int)(r % n0);
    j = (s ? f.vhi[1] + 1 : f.vlo[1] - (int)h) + (int)((r / n0) % h);
    k = f.vlo[2] + (int)(r / (n0 * h));
    return true;
  }
  t -= 2 * B;
  if (t < Cl) {                         // x-low slab over the valid y-z range
    i = f.lo[0] + (int)(t % hl);
    j = f.vlo[1] + (int)((t / hl) % v1);
    k = f.vlo[2] + (int)(t / (hl * v1));
    return true;
  }
  t -= Cl;
  if (t < Cr) {                         // x-high slab
    i = f.vhi[0] + 1 + (int)(t % hr);
    j = f.vlo[1] + (int)((t / hr) % v1);
    k = f.vlo[2] + (int)(t / (hr * v1));
    return true;
  }
  return false;
}

// Push of ONE ghost source cell x = (i,j,k) of fab D whose 15 populations live at sp[p * ssc]
// (its own ghost cell, or wherever a FillPatch plan says that ghost cell's value comes from):
// dst(x + c_p, p) = f_p for destinations inside valid grown by 1, loading only those populations;
// zero_invalid: a ghost destination keeps component p only if x - c_p is valid; the ring-2 cell
// zeroes itself (fresh-fab fill).
__device__ __forceinline__ void ghost_push(const DFabT& D, int i, int j, int k, const double* __restrict__ sp,
                                           long long ssc, int zero_invalid) {
  double* dp = static_cast<double*>(D.p) + mf_off(D, i, j, k);
  const long long dsc = mf_stride(D), dy = D.n[0], dz = (long long)D.n[0] * D.n[1];
  // signed ring distance of x from the valid box per direction: <0 below, >0 above, 0 inside
  const int ex = i < D.vlo[0] ? i - D.vlo[0] : i > D.vhi[0] ? i - D.vhi[0] : 0;
  const int ey = j < D.vlo[1] ? j - D.vlo[1] : j > D.vhi[1] ? j - D.vhi[1] : 0;
  const int ez = k < D.vlo[2] ? k - D.vlo[2] : k > D.vhi[2] ? k - D.vhi[2] : 0;
  double f[NV];
  bool go[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    // ring of the destination per direction: moving inward from ring |e| lands in |e|-1
    const int rx = ex == 0 ? 0 : (ex > 0 ? ex + cx(p) : -(ex + cx(p)));
    const int ry = ey == 0 ? 0 : (ey > 0 ? ey + cy(p) : -(ey + cy(p)));
    const int rz = ez == 0 ? 0 : (ez > 0 ? ez + cz(p) : -(ez + cz(p)));
    go[p] = rx <= 1 && ry <= 1 && rz <= 1;
    // kept by ZeroInvalidComponents if the destination is valid, or if x - c_p is valid
    if (zero_invalid && go[p] && !mf_in_valid(D, i + cx(p), j + cy(p), k + cz(p)) &&
        !mf_in_valid(D, i - cx(p), j - cy(p), k - cz(p)))
      f[p] = 0.0;
    else
      f[p] = (go[p] && sp) ? __ldcs(sp + p * ssc) : 0.0;
  }
#pragma unroll
  for (int p = 0; p < NV; ++p)
    if (go[p]) __stcs(dp + p * dsc + cx(p) + cy(p) * dy + cz(p) * dz, f[p]);
  if (ex < -1 || ex > 1 || ey < -1 || ey > 1 || ez < -1 || ez > 1) {   // own cell is in ring 2
#pragma unroll
    for (int p = 0; p < NV; ++p) dp[p * dsc] = 0.0;
  }
}

// Ghost push of the conventional level step: as ghost_push, but the source populations are
//   collide_src : collide(sp[.]) -- all 15 loaded, collided, then the ones with a destination stored;
//   else        : sp[p] or, with a second source, wa * sp[p] + wb * spb[p] (separately rounded, like
//                 k_mf_lincomb followed by a copy).
template <class C>
__device__ __forceinline__ void ghost_push_level(const DFabT& D, int i, int j, int k, const double* __restrict__ sp,
                                                 long long ssc, const double* __restrict__ spb, double wa, double wb,
                                                 bool collide_src, double omega_s, double omega_b) {
  double* dp = static_cast<double*>(D.p) + mf_off(D, i, j, k);
  const long long dsc = mf_stride(D), dy = D.n[0], dz = (long long)D.n[0] * D.n[1];
  const int ex = i < D.vlo[0] ? i - D.vlo[0] : i > D.vhi[0] ? i - D.vhi[0] : 0;
  const int ey = j < D.vlo[1] ? j - D.vlo[1] : j > D.vhi[1] ? j - D.vhi[1] : 0;
  const int ez = k < D.vlo[2] ? k - D.vlo[2] : k > D.vhi[2] ? k - D.vhi[2] : 0;
  double f[NV];
  bool go[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    const int rx = ex == 0 ? 0 : (ex > 0 ? ex + cx(p) : -(ex + cx(p)));
    const int ry = ey == 0 ? 0 : (ey > 0 ? ey + cy(p) : -(ey + cy(p)));
    const int rz = ez == 0 ? 0 : (ez > 0 ? ez + cz(p) : -(ez + cz(p)));
    go[p] = rx <= 1 && ry <= 1 && rz <= 1;
  }
  if (collide_src) {
#pragma unroll
    for (int p = 0; p < NV; ++p) f[p] = __ldcs(sp + p * ssc);
    C::collide(f, omega_s, omega_b);
  } else {
#pragma unroll
    for (int p = 0; p < NV; ++p) {
      if (!go[p] || !sp) f[p] = 0.0;
      else if (spb) f[p] = __dadd_rn(__dmul_rn(wa, __ldcs(sp + p * ssc)), __dmul_rn(wb, __ldcs(spb + p * ssc)));
      else f[p] = __ldcs(sp + p * ssc);
    }
  }
#pragma unroll
  for (int p = 0; p < NV; ++p)
    if (go[p]) __stcs(dp + p * dsc + cx(p) + cy(p) * dy + cz(p) * dz, f[p]);
  if (ex < -1 || ex > 1 || ey < -1 || ey > 1 || ez < -1 || ez > 1) {   // own cell is in ring 2
#pragma unroll
    for (int p = 0; p < NV; ++p) dp[p * dsc] = 0.0;
  }
}

// ---- gather-plan records (built by lbx_plan_create, see mf_kernels.cuh for the semantics) ------
enum { G_COPY = 0, G_PC = 1, G_AVG = 2, G_CONST = 3, G_NONE = 4 };
struct alignas(16) GDesc {   // 64 bytes
  int lo[3], hi[3];   // destination region (destination index space)
  int shift[3];       // added in SOURCE index space after the map
  int src_set, src_fab, kind, ratio;
  int pad;
  double value;
};
static_assert(sizeof(GDesc) == 64, "GDesc must be 64 bytes (staged to shared memory as int4 words)");
struct GDst {
  int fab, first, count, pad;
  int blo[3], bhi[3];   // bounding box of this fab's regions
};

__device__ __forceinline__ int fdiv(int a, int r) { return a >= 0 ? a / r : -((-a + r - 1) / r); }

constexpr int PLAN_CHUNK = 64;   // descriptors staged in shared memory at a time (4 KB)

// One launch = one collide + Stream of a level.  grid = (tiles, fab): the CTAs of ONE box are
// adjacent in launch order -- first its valid-cell tiles, then its ghost-cell tiles -- so the
// bandwidth-bound valid CTAs and the latency-bound ghost CTAs of a few boxes are resident together
// and overlap (launched as separate kernels they run back to back: +30 % time).
//  * valid tile  : CSY rows of one z-plane, a warp per row (no per-thread division; a warp's access
//                  to a population plane is one contiguous row segment).  The source set has the
//                  destination's geometry, so only ONE box descriptor is read and the source
//                  address is dst's offset into the other allocation.
//  * ghost tiles : plan == null: the six shell slabs, each ghost cell pushes its OWN value in gt;
//                  plan != null: DistFnFillPatch :359-391 folded in -- the thread of ghost cell x looks
//                  up where FillPatch would take x's value from (COPY: a same-level box / periodic
//                  image; PC: the coarse cell under x; descriptors of a ghosts-only
//                  FillPatchSingleLevel / FillPatchTwoLevels plan, last match wins) and pushes from
//                  there; NONE / no match: the value FillPatch would have left, read from `fb`.
struct CSPlan {                    // ghosts-only FillPatch plan of the level being streamed (or all null)
  const GDst* dsts;                // groups (ghost slabs), sorted by fab
  const int* fab_first;            // [nfabs + 1] first group of each fab
  const GDesc* descs;
  const DFabT* s0;                 // same-level source set (NOW)
  const DFabT* s1;                 // coarse source set (NOW of level - 1)
  const DFabT* fb;                 // fallback: the fab FillPatch would have filled
  int tiles_per_group;
  // conventional level step (k_mf_collide_stream<.., LEVELSTEP = true>): ghost cells FillPatch takes
  // from a same-level valid cell push that cell's COLLIDED populations (CollideLevel's FillBoundary,
  // src/AmrSim.cpp:132, runs after its Collide); ghost cells under the coarse level push
  // wa * s1 (+ wb * s1b): FillPatchTwoLevels' interpolation in time between two coarse states
  const DFabT* s1b;                // second coarse set (same boxes as s1) or null
  double wa, wb;
};
constexpr int CSX = 32, CSY = 8;
// -DLBX_MF_CS_MIN_CTAS=4 (64 registers, 32 warps/SM) measured 4 % SLOWER than the natural 71-75 registers /
// 3 CTAs per SM (profiles/r01_alignment.md); note that a second launch-bound argument of 1 makes ptxas
// spend 96 registers (2 CTAs/SM)
#ifdef LBX_MF_CS_MIN_CTAS
#define LBX_MF_CS_BOUNDS __launch_bounds__(MFT, LBX_MF_CS_MIN_CTAS)
#else
#define LBX_MF_CS_BOUNDS __launch_bounds__(MFT)
#endif
static_assert(CSX * CSY == MFT, "valid and ghost tiles share one CTA shape");

// LINEAR (valid tiles): false = CSY rows of one z-plane per CTA, a warp per row (no per-thread
// division, but lanes beyond the row end and rows beyond the box idle: a 26^3 box uses 66 % of
// the lanes); true = MFT consecutive cells of the box's valid region in x-fastest order (two
// integer divisions per thread, every lane busy; a warp's plane access is 1-3 row segments).
// flags: bit 0 = zero_invalid; bits 8 / 9 (profiling only) skip the valid / ghost tiles' work.
// LEVELSTEP: the conventional level step (CollideLevel + Stream, src/AmrSim.cpp:124-135, 109-122) instead
// of the Rohde pair: needs a plan; see CSPlan and ghost_push_level.
template <class C, bool LINEAR, bool LEVELSTEP = false>
__global__ void LBX_MF_CS_BOUNDS k_mf_collide_stream(const double* __restrict__ vbase, double* __restrict__ dbase,
                                                           const DFabT* __restrict__ dt, const DFabT* __restrict__ mt,
                                                           const DFabT* __restrict__ gt, CSPlan plan, int nfabs,
                                                           int ytiles, int valid_tiles, double omega_s, double omega_b,
                                                           int fine_val, int flags) {
  __shared__ GDesc sd[PLAN_CHUNK];
  const int b = mf_local_fab(dt, nfabs);
  if (b < 0) return;
  const DFabT D = dt[b];
  if (!D.local) return;                                   // a peer's box: its owner streams it
  const int tid = threadIdx.x;
  const int zero_invalid = flags & 1;
  if ((int)blockIdx.x < valid_tiles) {
    if (flags & 0x100) return;
    // ---- valid source cells: collide, push ---------------------------------------------------
    int i0, j, k, istep;
    if (LINEAR) {
      // own-ghost mode with flags bit 10: the rows are tiled INCLUDING their 2 + 2 x-ghost cells, so that
      // consecutive threads cover consecutive memory across row ends and the x-ghost pushes / ring-2
      // zeros that complete a row's end sectors come from the same or the neighbouring warp
      const bool xrows = (flags & 0x400) && !plan.dsts && gt && !LEVELSTEP;
      // with x rows: the whole ALLOCATED row (ghost cells and alignment cells included)
      const unsigned nx = xrows ? (unsigned)D.n[0] : (unsigned)(D.vhi[0] - D.vlo[0] + 1);
      const unsigned ny = D.vhi[1] - D.vlo[1] + 1, nz = D.vhi[2] - D.vlo[2] + 1;
      unsigned t = blockIdx.x * MFT + tid;
      if (t >= nx * ny * nz) return;
      i0 = (xrows ? D.lo[0] : D.vlo[0]) + (int)(t % nx);
      t /= nx;
      j = D.vlo[1] + (int)(t % ny);
      k = D.vlo[2] + (int)(t / ny);
      istep = 1 << 30;                                    // one cell per thread
      if (i0 < D.vlo[0] || i0 > D.vhi[0]) {               // an x-ghost cell of a valid row: push its own value
        const DFabT S = gt[b];
        ghost_push(D, i0, j, k, static_cast<const double*>(S.p) + mf_off(S, i0, j, k), mf_stride(S), zero_invalid);
        return;
      }
    } else {
      const int ty = tid / CSX, tx = tid % CSX;
      j = D.vlo[1] + ((int)blockIdx.x % ytiles) * CSY + ty;
      k = D.vlo[2] + (int)blockIdx.x / ytiles;
      if (j > D.vhi[1] || k > D.vhi[2]) return;
      i0 = D.vlo[0] + tx;
      istep = CSX;
    }
    const int* mp = mt ? static_cast<const int*>(mt[b].p) : nullptr;   // mask: same boxes and ghosts, 1 comp
    const long long sc = mf_stride(D), dy = D.n[0], dz = (long long)D.n[0] * D.n[1];
    const long long row = mf_off(D, D.lo[0], j, k);                 // offset of x = lo[0] in this row
    double* drow = static_cast<double*>(D.p) + row;
    const double* srow = vbase + (static_cast<double*>(D.p) - dbase) + row;
    const bool rim_jk = j == D.vlo[1] || j == D.vhi[1] || k == D.vlo[2] || k == D.vhi[2];
    for (int i = i0; i <= D.vhi[0]; i += istep) {
      const int x = i - D.lo[0];
      double f[NV];
      if (mp && mp[row + x] == fine_val) {
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = 0.0;
      } else {
#pragma unroll
        for (int p = 0; p < NV; ++p) f[p] = __ldcs(srow + p * sc + x);
        C::collide(f, omega_s, omega_b);
      }
      // a valid source on the rim of its box pushes into ghost ring 1: that component survives
      // ZeroInvalidComponents only if x - c_p is valid too (false only for boxes one cell thick)
      if (zero_invalid && (rim_jk || i == D.vlo[0] || i == D.vhi[0])) {
#pragma unroll
        for (int p = 1; p < NV; ++p)
          if (!mf_in_valid(D, i + cx(p), j + cy(p), k + cz(p)) && !mf_in_valid(D, i - cx(p), j - cy(p), k - cz(p)))
            f[p] = 0.0;
      }
      if (flags & 0x800) {           // plain write-back stores: partial end sectors stay in L2 until completed
#pragma unroll
        for (int p = 0; p < NV; ++p) drow[p * sc + x + (cx(p) + cy(p) * dy + cz(p) * dz)] = f[p];
      } else {
#pragma unroll
        for (int p = 0; p < NV; ++p) __stcs(drow + p * sc + x + (cx(p) + cy(p) * dy + cz(p) * dz), f[p]);
      }
    }
    // x-ghost cells of this row pushed by the row's own warp (flags bit 10, own-ghost mode): the partial
    // sectors its valid cells left at the two row ends are completed microseconds later by the same
    // warp instead of by an x-slab ghost CTA much later in launch order (profiles/r01_alignment.md)
    if (!LINEAR && !LEVELSTEP && (flags & 0x400) && !plan.dsts && gt) {
      const int tx = tid % CSX;
      const int hl = D.vlo[0] - D.lo[0], hr = D.n[0] - (D.vhi[0] - D.vlo[0] + 1) - hl;   // ghost + alignment cells left / right
      if (tx < hl + hr) {
        const int gi = tx < hl ? D.lo[0] + tx : D.vhi[0] + 1 + (tx - hl);
        const DFabT S = gt[b];
        ghost_push(D, gi, j, k, static_cast<const double*>(S.p) + mf_off(S, gi, j, k), mf_stride(S), zero_invalid);
      }
    }
    return;
  }
  if (flags & 0x200) return;
  const unsigned gtile = blockIdx.x - valid_tiles;
  int i, j, k;
  if (!plan.dsts) {
    // ---- ghost source cells pushing their own values -------------------------------------------
    if (!gt || !mf_shell_cell(D, gtile * MFT + tid, i, j, k)) return;
    if ((flags & 0x400) && j >= D.vlo[1] && j <= D.vhi[1] && k >= D.vlo[2] && k <= D.vhi[2]) return;   // x slabs: done by the rows
    const DFabT S = gt[b];
    ghost_push(D, i, j, k, static_cast<const double*>(S.p) + mf_off(S, i, j, k), mf_stride(S), zero_invalid);
    return;
  }
  // ---- ghost source cells, FillPatch folded in ---------------------------------------------------
  const int q = plan.fab_first[b] + (int)(gtile / plan.tiles_per_group);
  if (q >= plan.fab_first[b + 1]) return;                  // block-uniform
  const GDst G = plan.dsts[q];
  const unsigned nx = G.bhi[0] - G.blo[0] + 1, ny = G.bhi[1] - G.blo[1] + 1, nz = G.bhi[2] - G.blo[2] + 1;
  const unsigned cells = nx * ny * nz, tile = gtile % plan.tiles_per_group;
  if (tile * MFT >= cells) return;
  unsigned t = tile * MFT + tid;
  const bool in_box = t < cells;
  i = G.blo[0] + (int)(t % nx);
  t /= nx;
  j = G.blo[1] + (int)(t % ny);
  k = G.blo[2] + (int)(t / ny);
  const bool ghost = in_box && !mf_in_valid(D, i, j, k);   // these plans tile ghost slabs only
  bool searching = ghost;
  const double* sp = nullptr;
  const double* spb = nullptr;         // LEVELSTEP: the second coarse state of a PC source
  bool collide_src = false;            // LEVELSTEP: the source is a same-level valid cell
  long long ssc = 0;
  const int nchunks = (G.count + PLAN_CHUNK - 1) / PLAN_CHUNK;
  for (int ch = 0; ch < nchunks; ++ch) {                   // backwards: the last matching descriptor wins
    const int cb = max(G.count - (ch + 1) * PLAN_CHUNK, 0), ce = G.count - ch * PLAN_CHUNK, n = ce - cb;
    {
      const int4* src = reinterpret_cast<const int4*>(plan.descs + G.first + cb);
      int4* dst = reinterpret_cast<int4*>(sd);
      for (int w = tid; w < n * (int)(sizeof(GDesc) / 16); w += MFT) dst[w] = src[w];
    }
    __syncthreads();
    if (searching) {
      for (int d = n - 1; d >= 0; --d) {
        const GDesc& g = sd[d];
        if (i < g.lo[0] || i > g.hi[0] || j < g.lo[1] || j > g.hi[1] || k < g.lo[2] || k > g.hi[2]) continue;
        if (g.kind == G_COPY || g.kind == G_PC) {
          const DFabT S = (g.src_set ? plan.s1 : plan.s0)[g.src_fab];
          int si = i, sj = j, sk = k;
          if (g.kind == G_PC) { si = fdiv(i, g.ratio); sj = fdiv(j, g.ratio); sk = fdiv(k, g.ratio); }
          const long long so = mf_off(S, si + g.shift[0], sj + g.shift[1], sk + g.shift[2]);
          sp = static_cast<const double*>(S.p) + so;
          ssc = mf_stride(S);
          if (LEVELSTEP) {
            collide_src = g.kind == G_COPY;
            if (g.kind == G_PC && plan.s1b) spb = static_cast<const double*>(plan.s1b[g.src_fab].p) + so;
          }
        }
        searching = false;                                  // a G_NONE hit ends the search too: no source
        break;
      }
    }
    __syncthreads();
  }
  if (!ghost) return;
  if (!sp && plan.fb) {
    const DFabT B = plan.fb[b];
    sp = static_cast<const double*>(B.p) + mf_off(B, i, j, k);
    ssc = mf_stride(B);
  }
  if (LEVELSTEP) {
    const bool lerp = spb != nullptr;
    ghost_push_level<C>(D, i, j, k, sp, ssc, spb, lerp ? plan.wa : 1.0, lerp ? plan.wb : 0.0, collide_src, omega_s, omega_b);
  } else {
    ghost_push(D, i, j, k, sp, ssc, zero_invalid);
  }
}

template <class C>
__global__ void __launch_bounds__(MFT) k_mf_moments(const DFabT* __restrict__ ft, const DFabT* __restrict__ rt,
                                                    const DFabT* __restrict__ ut, int nfabs) {
  const int b = mf_local_fab(ft, nfabs);
  if (b < 0) return;
  const DFabT F = ft[b];
  if (!F.local) return;
  int i, j, k;
  if (!mf_cell(F, 0, i, j, k)) return;
  const double* fp = static_cast<const double*>(F.p) + mf_off(F, i, j, k);
  const long long sc = mf_stride(F);
  double f[NV];
#pragma unroll
  for (int p = 0; p < NV; ++p) f[p] = fp[p * sc];
  double rho, ux, uy, uz;
  C::moments(f, rho, ux, uy, uz);
  const DFabT R = rt[b], U = ut[b];
  static_cast<double*>(R.p)[mf_off(R, i, j, k)] = rho;
  double* up = static_cast<double*>(U.p) + mf_off(U, i, j, k);
  const long long usc = mf_stride(U);
  up[0] = ux;
  up[usc] = uy;
  up[2 * usc] = uz;
}

// CalcEquilibriumDist (src/AmrSim.cpp:845-931) on the valid cells of every box.
template <class C>
__global__ void __launch_bounds__(MFT) k_mf_equilibrium(const DFabT* __restrict__ ft, const DFabT* __restrict__ rt,
                                                        const DFabT* __restrict__ ut, int nfabs) {
  const int b = mf_local_fab(ft, nfabs);
  if (b < 0) return;
  const DFabT F = ft[b];
  if (!F.local) return;
  int i, j, k;
  if (!mf_cell(F, 0, i, j, k)) return;
  const DFabT R = rt[b], U = ut[b];
  const double* up = static_cast<const double*>(U.p) + mf_off(U, i, j, k);
  const long long usc = mf_stride(U);
  double f[NV];
  equilibrium_cell(static_cast<const double*>(R.p)[mf_off(R, i, j, k)], up[0], up[usc], up[2 * usc], f);
  double* fp = static_cast<double*>(F.p) + mf_off(F, i, j, k);
  const long long sc = mf_stride(F);
#pragma unroll
  for (int p = 0; p < NV; ++p) fp[p * sc] = f[p];
}

inline dim3 mf_grid(long long max_cells, int nfabs, long long extra_tiles = 0, int tile_mult = 1) {
  if (nfabs < 1) nfabs = 1;                // a rank that owns no box still launches (one row of CTAs that exit)
  const unsigned gy = (unsigned)(nfabs < 65535 ? nfabs : 65535);
  return dim3((unsigned)(((max_cells + MFT - 1) / MFT + extra_tiles) * tile_mult), gy, (unsigned)((nfabs + gy - 1) / gy));
}

inline dim3 grid_for(const DBox& b) {
  return dim3((unsigned)((b.hi[0] - b.lo[0] + BX) / BX), (unsigned)(b.hi[1] - b.lo[1] + 1),
              (unsigned)(b.hi[2] - b.lo[2] + 1));
}

}  // namespace lbx
