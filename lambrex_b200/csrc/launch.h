// Host-callable launchers, one set per translation unit / collision variant.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "rows_kernel.cuh"

namespace lbx {
// dynamic shared memory added to the fused kernels' launches purely to cap resident CTAs/SM
// (occupancy tuning knob, LBX_OPT_SMEM_PAD); 0 = no cap
extern int g_smem_pad;
// valid tiles of k_mf_collide_stream: 1 = MFT consecutive cells (default), 0 = a warp per row (LBX_OPT_VALID_TILING)
extern int g_valid_linear;
// 1: fabs with ghost cells are allocated with sector-aligned valid rows (LBX_OPT_ALIGN_ROWS); default 0
extern int g_align_rows;
// 1: the valid-row warps of k_mf_collide_stream also push their row's x-ghost cells (LBX_OPT_XGHOST_IN_ROW)
extern int g_xghost_in_row;
// 1: valid-cell pushes of k_mf_collide_stream use write-back instead of streaming stores (LBX_OPT_PLAIN_STORES)
extern int g_plain_stores;
// profiling only (LBX_OPT_DEBUG_SKIP): bit 0 skips the valid tiles' work, bit 1 the ghost tiles' (results are wrong)
extern int g_debug_skip;
// 1 (default): the row-owner kernel k_mf_cs_rows runs every fused collide + Stream that has a ghost source;
// 0: the round-1 tile kernel k_mf_collide_stream (LBX_OPT_ROW_KERNEL)
extern int g_row_kernel;
struct Launchers {
  void (*equilibrium)(cudaStream_t, DFab f, DFab rho, DFab u, DBox box);
  void (*moments)(cudaStream_t, DFab f, DFab rho, DFab u, DBox box);
  void (*collide)(cudaStream_t, DFab src, DFab dst, DBox box, double ws, double wb, DMask mask, int fine_val);
  void (*stream)(cudaStream_t, DFab src, DFab dst, DBox box, DDom dom);
  void (*collide_stream)(cudaStream_t, DFab src, DFab dst, DBox box, DDom dom, double ws, double wb, int scheme);
  void (*collide_stream_slab)(cudaStream_t, DFab src, DFab dst, DFab dn, DFab up, DBox box, DDom dom, double ws,
                              double wb);
  void (*collide_stream_slab_sync)(cudaStream_t, DFab src, DFab dst, DFab dn, DFab up, DBox box, DDom dom, double ws,
                                   double wb, SlabSync sy);
  void (*mf_collide)(cudaStream_t, const DFabT* src, const DFabT* f, const DFabT* mask, int nfabs, long long max_cells, double ws,
                     double wb, int fine_val);
  void (*mf_collide_stream)(cudaStream_t, const double* vbase, double* dbase, const DFabT* dst, const DFabT* mask,
                            const DFabT* gsrc, CSPlan plan, int nfabs, int max_ny, int max_nz, long long max_valid,
                            long long ghost_tiles, double ws, double wb, int fine_val, int zero_invalid);
  // mode 1 own ghost cells, 2 FillPatch plan, 3 conventional level step; max_n1 / max_n2 = largest grown y / z extent of any fab
  int (*mf_cs_rows)(cudaStream_t, ROArgs a, int mode, int max_n1, int max_n2);
  void (*mf_moments)(cudaStream_t, const DFabT* f, const DFabT* rho, const DFabT* u, int nfabs, long long max_cells);
  void (*mf_equilibrium)(cudaStream_t, const DFabT* f, const DFabT* rho, const DFabT* u, int nfabs,
                         long long max_cells);
};
const Launchers& launchers_fast();
const Launchers& launchers_literal();
}  // namespace lbx
