// -*- mode: c++ -*-
// range(n) / range(a, b): a constexpr-friendly half-open integer range for range-for loops
// (`for (auto i : range(D3Q15::NV))`), the loop sugar the reference takes from a third-party
// header of the same name (/root/reference/include/range.hpp, used by tests/meta_basic.cpp:6,16
// and include/d3q15_bgk.h:38).  Written from scratch: value iterator, no step, no containers.
#ifndef LBX_RANGE_HPP
#define LBX_RANGE_HPP

template <typename T>
struct range_span {
  struct iterator {
    T v;
    constexpr T operator*() const { return v; }
    constexpr iterator& operator++() { ++v; return *this; }
    constexpr bool operator!=(const iterator& o) const { return v != o.v; }
    constexpr bool operator==(const iterator& o) const { return v == o.v; }
  };
  T first, last;
  constexpr iterator begin() const { return iterator{first}; }
  constexpr iterator end() const { return iterator{last < first ? first : last}; }
  constexpr T size() const { return last < first ? T(0) : last - first; }
};
template <typename T>
constexpr range_span<T> range(T n) { return range_span<T>{T(0), n}; }
template <typename T>
constexpr range_span<T> range(T a, T b) { return range_span<T>{a, b}; }
#endif
