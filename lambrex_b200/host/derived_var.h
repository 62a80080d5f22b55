// -*- mode: c++ -*-
// Derived (output) variables computed from other fields.  Public names of
// /root/reference/include/derived_var.h:12-15, 55-91 (DerivedVar, ScalarTag, NELEM, deps).
// An implementation provides  static void fill(amrex::MultiFab& out, const amrex::MultiFab& dep...)
// as a device operation; there is no per-cell host loop.
#ifndef LBX_DERIVED_VAR_H
#define LBX_DERIVED_VAR_H
#include <cstddef>
#include <tuple>

#include "AMReX_MultiFab.H"

// dimensionality tags of a derived variable (derived_var.h:12-15; the reference sketches ElementTag /
// VectorTag in comments :17-31 -- ElementTag<N> is the N-component form used by LinearMoment below)
struct ScalarTag {
  static constexpr std::size_t RANK = 0;
  static constexpr std::size_t SIZE = 1;
};
template <std::size_t N>
struct ElementTag {
  static_assert(N > 0, "Invalid dimension of zero");
  static constexpr std::size_t RANK = 1;
  static constexpr std::size_t SIZE = N;
};
template <std::size_t N>
using VectorTag = ElementTag<N>;

// CRTP base (derived_var.h:55-91): Impl, a dimensionality tag, the fields it is computed from
template <typename Impl, typename DIM, typename... Deps>
struct DerivedVar {
  using dim_type = DIM;
  using base_type = DerivedVar<Impl, DIM, Deps...>;
  using dependencies = std::tuple<Deps...>;
  static constexpr std::size_t NELEM = DIM::SIZE;
  static constexpr std::size_t HALO = 0;
};

// Generic device path for derived variables that are linear moments of a distribution function
// (SURVEY.md 8f-3): where the reference's DerivedVar::fill_box (:80-90) loops over the cells of a
// box calling Impl::calculate, an implementation here states its per-cell rule as NELEM weight
// rows over the NV populations,
//     static constexpr double WEIGHTS[NELEM][NV];   static constexpr bool PER_UNIT_DENSITY;
// and fill() evaluates out_c = sum_p WEIGHTS[c][p] f_p (divided by rho when PER_UNIT_DENSITY) for
// every valid cell of the level in one launch (lbx_mf_linear_moments) -- a new moment needs no
// new kernel.  Density, MomentumDensity, Velocity and Stress in d3q15_bgk.h are such moments.
template <typename Impl, std::size_t DIM, typename DistT>
struct LinearMoment : DerivedVar<Impl, ElementTag<DIM>, DistT> {
  static void fill(amrex::MultiFab& out, const amrex::MultiFab& f) {
    static_assert(DIM >= 1 && DIM <= 10, "lbx_mf_linear_moments takes 1..10 weight rows");
    if (out.boxArray() != f.boxArray() || out.layout() != f.layout() || out.nComp() < (int)DIM)
      amrex::Abort("LinearMoment::fill: output and source must share boxes and layout");
    amrex::lbx_check(lbx_mf_linear_moments(f.mf(), out.mf(), &Impl::WEIGHTS[0][0], (int)DIM, Impl::PER_UNIT_DENSITY ? 1 : 0),
                     "LinearMoment::fill");
    out.touch();
  }
};
#endif
