// Collision-independent device kernels (compiled once, in lbx_abi.cu): face halo
// pack/unpack and the cross-GPU step-ordering flags of the peer-store exchange.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace lbx {

// ---------------------------------------------------------------------------
// Face halo pack / unpack: only the 5 populations that cross a face travel
// (D3Q15 has 6 axis and 8 body-diagonal velocities, include/d3q15_bgk.h:14-22).
// face: 0 +x, 1 -x, 2 +y, 3 -y, 4 +z, 5 -z  ->  populations with c.n > 0.
// buf layout: [5][cells of region], x fastest.  Replaces the FillBoundary of
// src/AmrSim.cpp:132 on the uniform path (15 comps x 2 ghosts there).
// ---------------------------------------------------------------------------
__host__ __device__ constexpr int face_pop(int face, int q) {
  // q-th population (ascending index) whose velocity has the face's sign along its axis
  int seen = 0;
  for (int p = 1; p < NV; ++p) {
    const int c = (face >> 1) == 0 ? cx(p) : (face >> 1) == 1 ? cy(p) : cz(p);
    if (c == ((face & 1) ? -1 : 1)) {
      if (seen == q) return p;
      ++seen;
    }
  }
  return -1;
}

template <bool PACK>
__global__ void __launch_bounds__(BX) k_halo(DFab f, DBox reg, int face, double* __restrict__ buf) {
  const int i = reg.lo[0] + blockIdx.x * BX + threadIdx.x;
  const int j = reg.lo[1] + blockIdx.y;
  const int k = reg.lo[2] + blockIdx.z;
  if (i > reg.hi[0]) return;
  const long long nx = reg.hi[0] - reg.lo[0] + 1, ny = reg.hi[1] - reg.lo[1] + 1, nz = reg.hi[2] - reg.lo[2] + 1;
  const long long cells = nx * ny * nz;
  const long long c = (i - reg.lo[0]) + nx * ((j - reg.lo[1]) + ny * (long long)(k - reg.lo[2]));
  const long long fsc = plane_stride(f), o = row_off(f, j, k) + i;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    int p = 0;
    switch (face) {
      case 0: p = face_pop(0, q); break;
      case 1: p = face_pop(1, q); break;
      case 2: p = face_pop(2, q); break;
      case 3: p = face_pop(3, q); break;
      case 4: p = face_pop(4, q); break;
      default: p = face_pop(5, q); break;
    }
    if (PACK) buf[q * cells + c] = f.p[p * fsc + o];
    else f.p[p * fsc + o] = buf[q * cells + c];
  }
}

// ---------------------------------------------------------------------------
// Cross-GPU step ordering for the peer-store exchange: 8-byte flags in
// IPC-shared device memory.  k_peer_signal publishes "step s finished" to a
// neighbour (release at system scope, after the step kernel that precedes it
// in stream order); k_peer_wait spins (acquire, system scope) until both
// neighbours have published >= s, with a wall-clock timeout so a dead peer
// cannot hang the GPU (err is host-mapped memory).
// ---------------------------------------------------------------------------
__global__ void k_peer_signal(unsigned long long* flag_a, unsigned long long* flag_b, unsigned long long value) {
  __threadfence_system();
  if (flag_a) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_a), "l"(value) : "memory");
  if (flag_b) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_b), "l"(value) : "memory");
}

__global__ void k_peer_wait(const unsigned long long* flag_a, const unsigned long long* flag_b,
                            unsigned long long value, unsigned long long timeout_ns, int* err) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (int which = 0; which < 2; ++which) {
    const unsigned long long* fl = which ? flag_b : flag_a;
    if (!fl) continue;
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(fl) : "memory");
      if (v >= value) break;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) {
        *err = 1;
        __threadfence_system();
        return;
      }
      __nanosleep(200);
    }
  }
}

// All-rank barrier on the device (distributed AMR path): thread r publishes `epoch` into slot
// [my_rank] of rank r's flag array (CUDA-IPC peer pointer, release at system scope -- ordered after
// every kernel queued before it on this stream), then spins until its own slot [r] shows that
// rank r has arrived at the same barrier.  No host round trip; wall-clock timeout like k_peer_wait.
__global__ void k_par_barrier(unsigned long long* const* __restrict__ peer_flags, const unsigned long long* my_flags,
                              int my_rank, int world, unsigned long long epoch, unsigned long long timeout_ns, int* err) {
  const int r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flags[r] + my_rank), "l"(epoch) : "memory");
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + r) : "memory");
    if (v >= epoch) break;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      *err = 1;
      __threadfence_system();
      return;
    }
    __nanosleep(100);
  }
}

}  // namespace lbx
