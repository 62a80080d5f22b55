// Derived variables on the generic device path (include/derived_var.h LinearMoment): a shear wave,
// then momentum density, velocity and momentum flux of level 0, each ONE launch of
// lbx_mf_linear_moments with the variable's weight rows -- no per-variable kernel.  The reference
// only sketches these variables (/root/reference/include/d3q15_bgk.h:43-55, commented out).
// Usage: derived_vars [n [steps]]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "AmrSim.h"
#include "lambrex.h"

int main(int argc, char** argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 32, steps = argc > 2 ? std::atoi(argv[2]) : 50;
  std::vector<double> u((size_t)n * n * n * 3, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      for (int k = 0; k < n; ++k) u[(((size_t)i * n + j) * n + k) * 3] = 0.01 * std::sin(2.0 * M_PI * j / n);
  lambrexInit();
  {
    AmrSim sim(n, n, n, 0, {{1, 1, 1}}, 0.1, 0.1);
    sim.SetInitialDensity(1.0);
    sim.SetInitialVelocity(u);
    sim.InitFromScratch(0.0);
    sim.Iterate(steps);
    amrex::MultiFab mom, vel, flux;
    sim.CalcDerived<MomentumDensity>(0, mom);
    sim.CalcDerived<Velocity>(0, vel);
    sim.CalcDerived<MomentumFlux>(0, flux);
    const amrex::IntVect probe(0, n / 4, 0);
    std::printf("t = %g: rho u_x = %.6e, u_x = %.6e, Pi_xx = %.6e at j = n/4 (initial u_x amplitude 1e-2)\n", sim.GetTime(0),
                mom.hostValue(0, probe, 0), vel.hostValue(0, probe, 0), flux.hostValue(0, probe, 0));
  }
  lambrexFinalise();
  return 0;
}
