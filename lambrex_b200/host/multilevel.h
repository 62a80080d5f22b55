// -*- mode: c++ -*-
// Per-level storage: the fields of one level at one time instant (State), and the now/next
// pair with the level's clock (LevelData).  Public names of
// /root/reference/include/multilevel.h:16-85.  The MultiFabs are device handles.
#ifndef LBX_MULTILEVEL_H
#define LBX_MULTILEVEL_H
#include <array>
#include <utility>

#include "field.h"

struct TimeData {
  amrex::Real current = 0.0;
  amrex::Real delta = 0.0;
  int step = 0;
};

namespace lbx_detail {
template <typename F, typename... Fs>
struct index_of;
template <typename F, typename... Rest>
struct index_of<F, F, Rest...> : std::integral_constant<std::size_t, 0> {};
template <typename F, typename G, typename... Rest>
struct index_of<F, G, Rest...> : std::integral_constant<std::size_t, 1 + index_of<F, Rest...>::value> {};
}  // namespace lbx_detail

template <typename... Fields>
struct State {
  static constexpr std::size_t NFIELDS = sizeof...(Fields);
  std::array<amrex::MultiFab, NFIELDS> fields;

  template <typename F>
  amrex::MultiFab& get() { return fields[lbx_detail::index_of<F, Fields...>::value]; }
  template <typename F>
  const amrex::MultiFab& get() const { return fields[lbx_detail::index_of<F, Fields...>::value]; }

  void Define(const amrex::BoxArray& ba, const amrex::DistributionMapping& dm,
              amrex::Layout lay = amrex::Layout::BOXES) {
    (field_traits<Fields>::DefineLevelData(get<Fields>(), ba, dm, lay), ...);
  }
  void Clear() {
    for (auto& f : fields) f.clear();
  }
  void Relayout(amrex::Layout lay) {
    for (auto& f : fields) f.relayout(lay);
  }
};

template <typename StateT>
struct LevelData {
  void Define(const amrex::BoxArray& ba, const amrex::DistributionMapping& dm,
              amrex::Layout lay = amrex::Layout::BOXES) {
    now.Define(ba, dm, lay);
    next.Define(ba, dm, lay);
  }
  void Clear() {
    now.Clear();
    next.Clear();
    time = TimeData{};
  }
  void UpdateNow() { std::swap(now, next); }   // the whole State: every field changes sides

  TimeData time;
  StateT now;
  StateT next;
};
#endif
