// -*- mode: c++ -*-
// Compile-time description of a lattice velocity set.  Mirrors the public names of
// /root/reference/include/velocity_set.h:57-108 (ND, NV, HALO, CS2, C, X, Q, kronecker);
// an implementation supplies CX/CY/CZ/W.  The device kernels carry their own constexpr copy
// of the same lattice (lambrex_b200/csrc/d3q15.cuh); tests/test_host_sim.py checks the two agree.
#ifndef LBX_VELOCITY_SET_H
#define LBX_VELOCITY_SET_H
#include <array>
#include <cstddef>

template <typename Impl, std::size_t ND_, std::size_t NV_, int HALO_>
struct VelocitySet {
  static constexpr std::size_t ND = ND_;
  static constexpr std::size_t NV = NV_;
  static constexpr int HALO = HALO_;
  static constexpr double CS2 = 1.0 / 3.0;
  using ivec = std::array<int, ND_>;
  using dvec = std::array<double, ND_>;
  using dmat = std::array<std::array<double, ND_>, ND_>;

  static constexpr int kronecker(std::size_t a, std::size_t b) { return a == b ? 1 : 0; }

 private:
  static constexpr std::array<ivec, NV_> make_c() {
    std::array<ivec, NV_> c{};
    for (std::size_t i = 0; i < NV_; ++i) c[i] = ivec{{Impl::CX[i], Impl::CY[i], Impl::CZ[i]}};
    return c;
  }
  static constexpr std::array<dvec, NV_> make_x() {
    std::array<dvec, NV_> x{};
    for (std::size_t i = 0; i < NV_; ++i)
      x[i] = dvec{{(double)Impl::CX[i], (double)Impl::CY[i], (double)Impl::CZ[i]}};
    return x;
  }
  static constexpr std::array<dmat, NV_> make_q() {
    std::array<dmat, NV_> q{};
    const auto c = make_c();
    for (std::size_t i = 0; i < NV_; ++i)
      for (std::size_t a = 0; a < ND_; ++a)
        for (std::size_t b = 0; b < ND_; ++b) q[i][a][b] = c[i][a] * c[i][b] - CS2 * kronecker(a, b);
    return q;
  }

 public:
  static constexpr std::array<ivec, NV_> C = make_c();   // integer lattice vectors
  static constexpr std::array<dvec, NV_> X = make_x();   // the same as doubles
  static constexpr std::array<dmat, NV_> Q = make_q();   // c_a c_b - cs^2 delta_ab
};
#endif
