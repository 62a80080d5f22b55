// -*- mode: c++ -*-
// The D3Q15 model: lattice, distribution-function field, density derived variable and the
// level-data aliases.  Public names of /root/reference/include/d3q15_bgk.h:11-58.
#ifndef LBX_D3Q15_BGK_H
#define LBX_D3Q15_BGK_H
#include "component.h"
#include "derived_var.h"
#include "field.h"
#include "multilevel.h"
#include "velocity_set.h"

// rest, 6 axis neighbours (+-x, +-y, +-z), 8 body diagonals; 2 ghost cells (two fine substeps)
struct D3Q15 : VelocitySet<D3Q15, 3, 15, 2> {
  static constexpr int CX[15] = {0, 1, -1, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1};
  static constexpr int CY[15] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
  static constexpr int CZ[15] = {0, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1};
  static constexpr double W[15] = {2.0 / 9.0,  1.0 / 9.0,  1.0 / 9.0,  1.0 / 9.0,  1.0 / 9.0,
                                   1.0 / 9.0,  1.0 / 9.0,  1.0 / 72.0, 1.0 / 72.0, 1.0 / 72.0,
                                   1.0 / 72.0, 1.0 / 72.0, 1.0 / 72.0, 1.0 / 72.0, 1.0 / 72.0};
};

struct DistFn : Component<D3Q15> {};

// rho = sum_i f_i, as a device reduction over the 15 planes (k_mf_moments also yields u)
struct Density : DerivedVar<Density, ScalarTag, DistFn> {
  static void fill(amrex::MultiFab& rho, amrex::MultiFab& u_scratch, const amrex::MultiFab& f) {
    amrex::lbx_check(lbx_mf_moments(f.mf(), rho.mf(), u_scratch.mf()), "Density::fill");
    rho.touch();
    u_scratch.touch();
  }
};

// The derived variables the reference sketches but leaves commented out
// (/root/reference/include/d3q15_bgk.h:43-55), as linear moments on the generic device path.
struct MomentumDensity : LinearMoment<MomentumDensity, 3, DistFn> {      // rho u_a = sum_i f_i c_ia
  static constexpr bool PER_UNIT_DENSITY = false;
  static constexpr double WEIGHTS[3][15] = {{0, 1, -1, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1},
                                            {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
                                            {0, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1}};
};
struct Velocity : LinearMoment<Velocity, 3, DistFn> {                    // u_a = sum_i f_i c_ia / rho
  static constexpr bool PER_UNIT_DENSITY = true;
  static constexpr double WEIGHTS[3][15] = {{0, 1, -1, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1},
                                            {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
                                            {0, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1}};
};
// second moment sum_i f_i c_ia c_ib, components xx, xy, xz, yy, yz, zz
struct MomentumFlux : LinearMoment<MomentumFlux, 6, DistFn> {
  static constexpr bool PER_UNIT_DENSITY = false;
  static constexpr double WEIGHTS[6][15] = {{0, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1},
                                            {0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, -1, -1, 1, 1},
                                            {0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, -1, 1},
                                            {0, 0, 0, 1, 1, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1},
                                            {0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1},
                                            {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}};
};

using SimState = State<DistFn, Density>;
using SimLevelData = LevelData<SimState>;
#endif
