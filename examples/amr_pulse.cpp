// Planar density pulse in a periodic box: the scenario of the reference's only example
// (/root/reference/examples/amr_pulse.cpp) written against lambrex-b200's AmrSim.  Prints the
// density profile along z at three times.  Usage: amr_pulse [nx ny nz [steps]]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "AmrSim.h"
#include "lambrex.h"

static void print_profile(AmrSim& sim, const char* label) {
  const auto dims = sim.GetDims();
  sim.CalcHydroVars(0);
  std::printf("%s (t = %g, step %d)\n", label, sim.GetTime(0), sim.GetTimeStep(0));
  for (int k = 0; k < dims[2]; ++k)
    std::printf("  k=%3d rho=%.6g uz=%.6g\n", k, sim.GetDensity(dims[0] / 2, dims[1] / 2, k, 0),
                sim.GetVelocity(dims[0] / 2, dims[1] / 2, k, 2, 0));
}

int main(int argc, char** argv) {
  const int nx = argc > 3 ? std::atoi(argv[1]) : 10, ny = argc > 3 ? std::atoi(argv[2]) : 10,
            nz = argc > 3 ? std::atoi(argv[3]) : 50;
  const int steps = argc > 4 ? std::atoi(argv[4]) : 100;
  const double tau = 0.5, amplitude = 0.01;

  // rho = 1 with a bump on the plane k = nz/2 - 1, normalised by the mean of nz consecutive
  // entries starting in the middle of the array (as the reference example does)
  std::vector<double> rho((size_t)nx * ny * nz, 1.0);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) rho[((size_t)i * ny + j) * nz + (nz / 2 - 1)] += amplitude;
  const size_t start = (rho.size() + (size_t)ny * nz) / 2;
  double mean = 0.0;
  for (int k = 0; k < nz; ++k) mean += rho[start + k];
  mean /= nz;
  for (double& r : rho) r /= mean;

  lambrexInit();
  {
    AmrSim sim(nx, ny, nz, 0, {{1, 1, 1}}, tau, tau);
    sim.SetInitialDensity(rho);
    sim.SetInitialVelocity(0.0);
    sim.InitFromScratch(0.0);
    print_profile(sim, "initial");
    sim.Iterate(steps);
    print_profile(sim, "after first leg");
    sim.Iterate(steps);
    print_profile(sim, "after second leg");
  }
  lambrexFinalise();
  return 0;
}
