#include "lambrex.h"

#include "AMReX_FillPatch.H"

void lambrexInit() { amrex::lbx_check(lbx_init(-1), "lambrexInit"); }

void lambrexInitParallel(int rank, int nranks, int (*allgather)(const void*, size_t, void*, void*), void* user) {
  amrex::lbx_check(lbx_init(-1), "lambrexInitParallel");
  amrex::lbx_check(lbx_par_init(rank, nranks, allgather, user), "lambrexInitParallel");
  amrex::DistributionMapping::SetParallel(rank, nranks);
}

void lambrexSetParallelView(int rank, int nranks) { amrex::DistributionMapping::SetParallel(rank, nranks); }

void lambrexFinalise() {
  amrex::ClearPlanCache();
  lbx_finalize();
  amrex::DistributionMapping::SetParallel(0, 1);
}
