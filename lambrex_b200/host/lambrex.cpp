#include "lambrex.h"

#include "AMReX_FillPatch.H"

void lambrexInit() { amrex::lbx_check(lbx_init(-1), "lambrexInit"); }

void lambrexFinalise() {
  amrex::ClearPlanCache();
  lbx_finalize();
}
