// -*- mode: c++ -*-
// Compile-time traits of a field: how many elements per cell, how many ghost cells, and how
// its level data is allocated.  Public names of /root/reference/include/field.h:46-130.
#ifndef LBX_FIELD_H
#define LBX_FIELD_H
#include <type_traits>

#include "AMReX_MultiFab.H"
#include "component.h"
#include "derived_var.h"

namespace lbx_detail {
template <typename T, typename = void>
struct has_velocity_set : std::false_type {};
template <typename T>
struct has_velocity_set<T, std::void_t<typename T::VelocitySet>> : std::true_type {};
template <typename T, typename = void>
struct has_dv_marker : std::false_type {};
template <typename T>
struct has_dv_marker<T, std::void_t<decltype(T::is_derived_var)>> : std::true_type {};
}  // namespace lbx_detail

template <typename F>
struct is_component : lbx_detail::has_velocity_set<F> {};
template <typename F>
inline constexpr bool is_component_v = is_component<F>::value;
template <typename F>
inline constexpr bool is_derived_var_v = lbx_detail::has_dv_marker<F>::value;

template <typename F, typename = void>
struct field_traits;

// components: NV elements, HALO ghost cells
template <typename F>
struct field_traits<F, std::enable_if_t<is_component_v<F>>> {
  static constexpr int NELEM = (int)F::NV;
  static constexpr int HALO = F::HALO;
  static void DefineLevelData(amrex::MultiFab& mf, const amrex::BoxArray& ba, const amrex::DistributionMapping& dm,
                              amrex::Layout lay = amrex::Layout::BOXES) {
    mf.define(ba, dm, NELEM, HALO, lay);
  }
  static amrex::MultiFab MakeLevelData(const amrex::BoxArray& ba, const amrex::DistributionMapping& dm,
                                       amrex::Layout lay = amrex::Layout::BOXES) {
    return amrex::MultiFab(ba, dm, NELEM, HALO, lay);
  }
};

// derived variables: NELEM elements, no ghost cells
template <typename F>
struct field_traits<F, std::enable_if_t<is_derived_var_v<F>>> {
  static constexpr int NELEM = (int)F::NELEM;
  static constexpr int HALO = 0;
  static void DefineLevelData(amrex::MultiFab& mf, const amrex::BoxArray& ba, const amrex::DistributionMapping& dm,
                              amrex::Layout lay = amrex::Layout::BOXES) {
    mf.define(ba, dm, NELEM, HALO, lay);
  }
  static amrex::MultiFab MakeLevelData(const amrex::BoxArray& ba, const amrex::DistributionMapping& dm,
                                       amrex::Layout lay = amrex::Layout::BOXES) {
    return amrex::MultiFab(ba, dm, NELEM, HALO, lay);
  }
};
#endif
