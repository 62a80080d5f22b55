// Host-callable launchers, one set per translation unit / collision variant.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace lbx {
struct Launchers {
  void (*equilibrium)(cudaStream_t, DFab f, DFab rho, DFab u, DBox box);
  void (*moments)(cudaStream_t, DFab f, DFab rho, DFab u, DBox box);
  void (*collide)(cudaStream_t, DFab src, DFab dst, DBox box, double ws, double wb, DMask mask, int fine_val);
  void (*stream)(cudaStream_t, DFab src, DFab dst, DBox box, DDom dom);
  void (*collide_stream)(cudaStream_t, DFab src, DFab dst, DBox box, DDom dom, double ws, double wb, int scheme);
};
const Launchers& launchers_fast();
const Launchers& launchers_literal();
}  // namespace lbx
