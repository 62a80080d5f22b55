// -*- mode: c++ -*-
// A field that streams along a velocity set (the distribution function).  Public names of
// /root/reference/include/component.h:15-30.  The per-cell pull-streaming rule
//     f_prop(x, i) = f(x - C[i], i)
// that the reference implements in Component::PropagatePoint runs on the device here:
// k_mf_stream (AMR path) and k_collide_stream (uniform path) in lambrex_b200/csrc/.
#ifndef LBX_COMPONENT_H
#define LBX_COMPONENT_H
#include "AMReX_MultiFab.H"

template <typename VS>
struct Component {
  using VelocitySet = VS;
  static constexpr auto ND = VS::ND;
  static constexpr auto NV = VS::NV;
  static constexpr auto HALO = VS::HALO;
  // whole-level form of PropagatePoint: dst(x,i) = src(x - C[i], i) on valid grown by one
  static void Propagate(const amrex::MultiFab& src, amrex::MultiFab& dst) {
    amrex::lbx_check(lbx_mf_stream(src.mf(), dst.mf()), "Component::Propagate");
    dst.touch();
  }
};
#endif
