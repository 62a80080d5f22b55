// -*- mode: c++ -*-
// Derived (output) variables computed from other fields.  Public names of
// /root/reference/include/derived_var.h:12-15, 55-91 (DerivedVar, ScalarTag, NELEM, deps).
// An implementation provides  static void fill(amrex::MultiFab& out, const amrex::MultiFab& dep...)
// as a device operation; there is no per-cell host loop.
#ifndef LBX_DERIVED_VAR_H
#define LBX_DERIVED_VAR_H
#include <cstddef>
#include <tuple>

struct ScalarTag {};

template <typename Impl, std::size_t DIM, typename... Deps>
struct DerivedVar {
  static constexpr std::size_t NELEM = DIM;
  using dependencies = std::tuple<Deps...>;
  static constexpr bool is_derived_var = true;
};
#endif
