// -*- mode: c++ -*-
// Compile-time traits of a field: how many elements per cell, how many ghost cells, and how
// its level data is allocated.  Public names of /root/reference/include/field.h:46-130.
#ifndef LBX_FIELD_H
#define LBX_FIELD_H
#include <type_traits>

#include "AMReX_MultiFab.H"
#include "component.h"
#include "derived_var.h"

// category detection, same names as the reference's namespace detail (field.h:49-84; pinned by the
// reference's compile-time test tests/meta_basic.cpp:25-31)
namespace detail {
struct component_tag {};
struct derived_var_tag {};

// T is a Component: it derives from Component<T::VelocitySet>
template <typename T, typename Enable = void>
struct is_component : std::false_type {};
template <typename T>
struct is_component<T, std::enable_if_t<std::is_base_of_v<Component<typename T::VelocitySet>, T>>> : std::true_type {};
template <typename T>
inline constexpr bool is_component_v = is_component<T>::value;

// T is a DerivedVar: a T* converts to a pointer to some DerivedVar<Impl, DIM, deps...> base
template <typename Impl, typename DIM, typename... dependencies>
constexpr bool blah(const DerivedVar<Impl, DIM, dependencies...>*) { return true; }
constexpr bool foo(...) { return false; }
template <typename T>
constexpr auto foo(const T* t = nullptr) -> decltype(blah(t)) { return true; }
template <typename T>
inline constexpr bool is_derived_var_v = foo(static_cast<const T*>(nullptr));
}  // namespace detail

template <typename F>
using is_component = detail::is_component<F>;
template <typename F>
inline constexpr bool is_component_v = detail::is_component_v<F>;
template <typename F>
inline constexpr bool is_derived_var_v = detail::is_derived_var_v<F>;

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
