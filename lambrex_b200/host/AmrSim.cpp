// AmrSim on the GPU: host control flow of the reference's time stepping
// (/root/reference/src/AmrSim.cpp), every field operation a kernel launch through
// include/lbx.h.  Reference lines are cited per member.
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <sys/stat.h>
#include <cmath>
#include <cstdlib>
#include "AmrSim.h"

#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>

using amrex::Box;
using amrex::BoxArray;
using amrex::DistributionMapping;
using amrex::IntVect;
using amrex::Layout;
using amrex::lbx_check;
using amrex::MultiFab;

namespace {
amrex::AmrInfo make_info(int max_ref_level) {
  amrex::AmrInfo info;
  info.verbose = 1;                                                        // src/AmrSim.cpp:763
  info.max_level = max_ref_level;
  info.ref_ratio.assign((size_t)max_ref_level + 1, IntVect(2, 2, 2));       // :765-766
  info.blocking_factor.assign((size_t)max_ref_level + 1, IntVect(1, 1, 1)); // :767-768
  return info;
}
lbx_box to_lbx(const Box& b) {
  lbx_box r;
  for (int d = 0; d < 3; ++d) { r.lo[d] = b.smallEnd(d); r.hi[d] = b.bigEnd(d); }
  return r;
}
// entry (m, p) of the moment basis / its inverse as the device library defines them
double table_entry(bool inverse, int r, int c) {
  struct Tables {
    double M[NMODES * NMODES], Mi[NMODES * NMODES], w[NMODES];
    int32_t c[NMODES * NDIMS];
    Tables() { lbx_d3q15_tables(M, Mi, c, w); }
  };
  static const Tables t;
  return (inverse ? t.Mi : t.M)[r * NMODES + c];
}
}  // namespace

namespace { bool g_allow_walls = false; }
void AmrSim::AllowWalls(bool on) { g_allow_walls = on; }
bool AmrSim::WallsAllowed() { return g_allow_walls; }

// src/AmrSim.cpp:1033-1073 (static member definitions)
const double AmrSim::DELTA[NDIMS][NDIMS] = {{1.0 / NMODES, 0.0, 0.0}, {0.0, 1.0 / NMODES, 0.0}, {0.0, 0.0, 1.0 / NMODES}};
#define LBX_ROW(I, r) {table_entry(I, r, 0), table_entry(I, r, 1), table_entry(I, r, 2), table_entry(I, r, 3), table_entry(I, r, 4), \
                       table_entry(I, r, 5), table_entry(I, r, 6), table_entry(I, r, 7), table_entry(I, r, 8), table_entry(I, r, 9), \
                       table_entry(I, r, 10), table_entry(I, r, 11), table_entry(I, r, 12), table_entry(I, r, 13), table_entry(I, r, 14)}
#define LBX_ROWS(I) {LBX_ROW(I, 0), LBX_ROW(I, 1), LBX_ROW(I, 2), LBX_ROW(I, 3), LBX_ROW(I, 4), LBX_ROW(I, 5), LBX_ROW(I, 6), LBX_ROW(I, 7), \
                     LBX_ROW(I, 8), LBX_ROW(I, 9), LBX_ROW(I, 10), LBX_ROW(I, 11), LBX_ROW(I, 12), LBX_ROW(I, 13), LBX_ROW(I, 14)}
const double AmrSim::MODE_MATRIX[NMODES][NMODES] = LBX_ROWS(false);
const double AmrSim::MODE_MATRIX_INVERSE[NMODES][NMODES] = LBX_ROWS(true);
#undef LBX_ROWS
#undef LBX_ROW

// src/AmrSim.cpp:753-802
AmrSim::AmrSim(int const nx, int const ny, int const nz, int const max_ref_level,
               const std::array<int, NDIMS>& periodicity, double const tau_s_0, double const tau_b_0)
    : AmrCore(amrex::Geometry(Box(IntVect(0, 0, 0), IntVect(nx - 1, ny - 1, nz - 1)),
                              amrex::RealBox({{0.0, 0.0, 0.0}}, {{1.0, 1.0, 1.0}}), 0, periodicity),
              make_info(max_ref_level)),
      NX(geom[0].Domain().length(0)), NY(geom[0].Domain().length(1)), NZ(geom[0].Domain().length(2)),
      NUMEL(NX * NY * NZ), COORD_SYS(0),
      PERIODICITY{{geom[0].period(0), geom[0].period(1), geom[0].period(2)}}, levels((size_t)max_ref_level + 1) {
  std::cout << "NX: " << NX << " NY: " << NY << " NZ: " << NZ << std::endl;
  const int num_levels = max_level + 1;
  velocity.resize(num_levels);
  stream_scratch.resize(num_levels);
  valid_pending.assign(num_levels, false);
  tau_s.resize(num_levels);
  tau_b.resize(num_levels);
  mass.resize(num_levels);
  static_tags.resize(num_levels);
  gradient_threshold.assign(num_levels, 0.0);
  fine_masks.resize(num_levels);
  for (int d = 0; d < NDIMS; ++d)
    if (!PERIODICITY[d]) {
      if (!WallsAllowed()) amrex::Abort("Currently only periodic boundary conditions allowed.");
      has_walls = true;            // AllowWalls: this direction is closed by bounce-back walls
    }
  tau_s.at(0) = tau_s_0;
  tau_b.at(0) = tau_b_0;
  if (!lbx_initialized()) amrex::Abort("AmrSim: call lambrexInit() first (no CUDA context; there is no CPU path)");
  // default coupling of refined hierarchies for callers that cannot call SetCoupling -- the reference's own,
  // unmodified test and example sources (oracle/Makefile ref_tests): LBX_COUPLING=subcycle|rohde
  if (const char* c = std::getenv("LBX_COUPLING")) {
    if (!std::strcmp(c, "subcycle")) coupling = Coupling::SUBCYCLE;
    else if (!std::strcmp(c, "rohde")) coupling = Coupling::ROHDE;
    else amrex::Abort("LBX_COUPLING must be 'rohde' or 'subcycle'");
  }
}

AmrSim::~AmrSim() {
  if (lbx_initialized()) lbx_sync();
}

// ----------------------------------------------------------------------------- input / output
void AmrSim::SetInitialDensity(double const rho_init) { density_view = nullptr; density_profile_axis = -1; initial_density.assign(NUMEL, rho_init); }
void AmrSim::SetInitialDensity(std::vector<double> rho_init) { density_view = nullptr; density_profile_axis = -1; initial_density = std::move(rho_init); }
void AmrSim::SetInitialVelocity(double const u_init) { velocity_view = nullptr; velocity_profile_axis = -1; initial_velocity.assign(3 * (size_t)NUMEL, u_init); }
void AmrSim::SetInitialVelocity(std::vector<double> u_init) { velocity_view = nullptr; velocity_profile_axis = -1; initial_velocity = std::move(u_init); }
void AmrSim::SetInitialDensityProfile(int const axis, std::vector<double> rho_of_axis) {
  if (axis < 0 || axis >= NDIMS || (int)rho_of_axis.size() != geom[0].Domain().length(axis))
    amrex::Abort("SetInitialDensityProfile: one value per cell along the axis");
  density_view = nullptr;
  density_profile_axis = axis;
  density_profile = std::move(rho_of_axis);
}
void AmrSim::SetInitialVelocityProfile(int const axis, std::vector<double> u_of_axis) {
  if (axis < 0 || axis >= NDIMS || (int)u_of_axis.size() != NDIMS * geom[0].Domain().length(axis))
    amrex::Abort("SetInitialVelocityProfile: three values per cell along the axis");
  velocity_view = nullptr;
  velocity_profile_axis = axis;
  velocity_profile = std::move(u_of_axis);
}

// the box this rank owns of level 0 in a distributed uniform run (its z-slab); the whole domain on one rank
Box AmrSim::LocalBox() {
  const int np = DistributionMapping::NProcs();
  if (np <= 1) return geom[0].Domain();
  const BoxArray ba = grids[0].empty() ? MakeBaseGrids() : grids[0];
  const DistributionMapping dm = (dmap[0].size() == ba.size()) ? dmap[0] : MakeDistributionMap(ba);
  std::vector<Box> slabs;
  if (!amrex::SlabOwnership(ba, dm, np, &slabs)) amrex::Abort("LocalBox: level 0 is not owned as one slab per rank");
  return slabs[DistributionMapping::MyProc()];
}

void AmrSim::upload_profile(MultiFab& mf, const std::vector<double>& profile, int axis, int ncomp) {
  void* dev = nullptr;
  lbx_check(lbx_malloc(&dev, profile.size() * sizeof(double)), "upload_profile");
  int rc = lbx_h2d(dev, profile.data(), profile.size() * sizeof(double));
  if (!rc) rc = lbx_mf_fill_profile(mf.mf(), static_cast<const double*>(dev), axis, geom[0].Domain().smallEnd(axis),
                                    geom[0].Domain().length(axis), ncomp);
  const int rc2 = lbx_free(dev);          // synchronises the stream first
  lbx_check(rc, "upload_profile");
  lbx_check(rc2, "upload_profile");
  mf.touch();
}

// user array (C-ordered, i slowest, component fastest) -> device field in fab order: one
// host->device copy of the raw array, then a transposing kernel.  local: the array covers this rank's
// slab only (SetInitial*LocalView).
void AmrSim::upload_user_field(MultiFab& mf, const double* user, size_t n, int ncomp, bool local) {
  if (local && !mf.isFlat()) amrex::Abort("local initial arrays need level 0 stored as one slab per rank");
  const Box ub = local ? mf.storageValid(mf.localSlab()) : geom[0].Domain();
  const size_t need = (size_t)ub.numPts() * ncomp;
  if (n < need) throw std::out_of_range("AmrSim: initial field has fewer than NX*NY*NZ*ncomp entries");
  // staged through the library's device buffers chunk by chunk (copies overlap the transposing kernels);
  // `user` must stay valid until the stream is drained: vectors owned by this object do, views are the
  // caller's promise (until InitFromScratch returns) -- so drain here
  const lbx_box dom = to_lbx(ub);
  lbx_check(lbx_mf_from_user_host(mf.mf(), user, &dom, ncomp), "upload_user_field");
  lbx_check(lbx_sync(), "upload_user_field");
  mf.touch();
}

// src/AmrSim.cpp:138-214
void AmrSim::InitDensity(int const level) {
  if (level) amrex::Abort("Only level 0 should be initialised from scratch currently.");
  MultiFab& rho = levels.at(level).now.get<Density>();
  if (density_view) upload_user_field(rho, density_view, density_view_n, 1, views_local);
  else if (density_profile_axis >= 0) upload_profile(rho, density_profile, density_profile_axis, 1);
  else upload_user_field(rho, initial_density.data(), initial_density.size(), 1, false);
}
// src/AmrSim.cpp:217-295
void AmrSim::InitVelocity(int const level) {
  if (level) amrex::Abort("Only LEVEL 0 should be initialised from scratch currently.");
  if (velocity_view) upload_user_field(velocity.at(level), velocity_view, velocity_view_n, NDIMS, views_local);
  else if (velocity_profile_axis >= 0) upload_profile(velocity.at(level), velocity_profile, velocity_profile_axis, NDIMS);
  else upload_user_field(velocity.at(level), initial_velocity.data(), initial_velocity.size(), NDIMS, false);
}

// src/AmrSim.cpp:824-843
double AmrSim::GetDensity(int const i, int const j, int const k, int const level) const {
  const IntVect pos(i, j, k);
  const MultiFab& rho = levels[level].now.get<Density>();
  for (amrex::MFIter mfi(rho); mfi.isValid(); ++mfi)
    if (mfi.validbox().contains(pos)) return rho.hostValue(mfi.index(), pos, 0);
  return NL_DENSITY;
}
double AmrSim::GetVelocity(int const i, int const j, int const k, int const n, int const level) const {
  const IntVect pos(i, j, k);
  const MultiFab& u = velocity.at(level);
  for (amrex::MFIter mfi(u); mfi.isValid(); ++mfi)
    if (mfi.validbox().contains(pos)) return u.hostValue(mfi.index(), pos, n);
  return NL_VELOCITY;
}

// dense C-ordered copy of a field into caller memory (device-side transpose, one D2H copy)
void AmrSim::dense_field_into(const MultiFab& mf, int level, double sentinel, double* out, size_t n, bool local) const {
  if (mf.empty()) amrex::Abort("dense field requested on an empty level");
  if (local && !mf.isFlat()) amrex::Abort("local fields need the level stored as one slab per rank");
  const Box domb = local ? mf.storageValid(mf.localSlab()) : geom.at(level).Domain();
  const int nc = mf.nComp() > 0 ? mf.nComp() : 1;
  if (n != (size_t)domb.numPts() * nc) amrex::Abort("dense field: buffer size mismatch");
  const lbx_box dom = to_lbx(domb);
  const bool holes = !local && mf.boxArray().numPts() != domb.numPts();      // cells the level does not hold: sentinel
  lbx_check(lbx_mf_to_user_host(mf.mf(), out, &dom, nc, local ? 1 : 0, holes ? 1 : 0, sentinel), "dense_field");
}
void AmrSim::GetDensityField(int const level, double* out, size_t n) const {
  dense_field_into(levels.at(level).now.get<Density>(), level, NL_DENSITY, out, n);
}
void AmrSim::GetVelocityField(int const level, double* out, size_t n) const {
  dense_field_into(velocity.at(level), level, NL_VELOCITY, out, n);
}
void AmrSim::GetLocalDensityField(int const level, double* out, size_t n) const {
  dense_field_into(levels.at(level).now.get<Density>(), level, NL_DENSITY, out, n, true);
}
void AmrSim::GetLocalVelocityField(int const level, double* out, size_t n) const {
  dense_field_into(velocity.at(level), level, NL_VELOCITY, out, n, true);
}
void AmrSim::GetLinearMomentField(int const level, const double* weights, int const ncomp, bool const per_unit_density,
                                  double const sentinel, double* out, size_t n) const {
  const MultiFab& f = levels.at(level).now.get<DistFn>();
  if (f.empty()) amrex::Abort("GetLinearMomentField: empty level");
  MultiFab dv(f.boxArray(), f.DistributionMap(), ncomp, 0, f.layout());
  lbx_check(lbx_mf_linear_moments(f.mf(), dv.mf(), weights, ncomp, per_unit_density ? 1 : 0), "GetLinearMomentField");
  dv.touch();
  dense_field_into(dv, level, sentinel, out, n);
}
std::vector<double> AmrSim::GetDensityField(int const level) const {
  std::vector<double> out((size_t)geom.at(level).Domain().numPts());
  GetDensityField(level, out.data(), out.size());
  return out;
}
std::vector<double> AmrSim::GetVelocityField(int const level) const {
  std::vector<double> out((size_t)geom.at(level).Domain().numPts() * NDIMS);
  GetVelocityField(level, out.data(), out.size());
  return out;
}

// src/AmrSim.cpp:1019-1030
std::pair<std::array<int, NDIMS>, std::array<int, NDIMS>> AmrSim::GetExtent(int const level) const {
  const Box mb = levels[level].now.get<DistFn>().boxArray().minimalBox();
  return {{{mb.smallEnd(0), mb.smallEnd(1), mb.smallEnd(2)}}, {{mb.bigEnd(0), mb.bigEnd(1), mb.bigEnd(2)}}};
}

// ----------------------------------------------------------------------------- layouts
// Level 0 while it is the only level: ONE ghost-free fab per rank (FLAT) -- the whole periodic domain on a
// single GPU, the rank's z-slab in a distributed run (box ownership by whole x-y layers, SlabOwnership) --
// so that a time step is one fused launch per GPU.  Anything else: per-box storage with ghost cells.
Layout AmrSim::PreferredLayout(int const level) const { return PreferredLayout(level, grids[level], dmap[level]); }
Layout AmrSim::PreferredLayout(int const level, const BoxArray& ba, const DistributionMapping& dm) const {
  if (!(level == 0 && finest_level == 0 && uniform_fast_path)) return Layout::BOXES;
  const int np = DistributionMapping::NProcs();
  if (np == 1) return Layout::FLAT;
  return amrex::SlabOwnership(ba, dm, np) ? Layout::FLAT : Layout::BOXES;
}

void AmrSim::FinishPeerStores() {
  if (!peer_stores_pending) return;
  lbx_check(lbx_par_step_finish(), "FinishPeerStores");
  peer_stores_pending = false;
}

void AmrSim::SetLevelLayout(int const level, Layout lay) {
  auto& lvl = levels.at(level);
  if (lvl.now.get<DistFn>().empty() || lvl.now.get<DistFn>().layout() == lay) return;
  lvl.now.Relayout(lay);                         // keeps the valid cells of f and rho
  velocity.at(level).relayout(lay);
  const BoxArray ba = lvl.now.get<DistFn>().boxArray();
  const DistributionMapping dm = lvl.now.get<DistFn>().DistributionMap();
  lvl.next.Define(ba, dm, lay);                  // next is scratch between steps
  stream_scratch.at(level).clear();
}

// ----------------------------------------------------------------------------- physics
// src/AmrSim.cpp:845-936: f <- f_eq(rho, u) on the valid cells of NOW; then (sic, SURVEY.md
// B-6) the boundary fill goes to NEXT.
void AmrSim::CalcEquilibriumDist(int const level) {
  auto& lvl = levels.at(level);
  MultiFab& f = lvl.now.get<DistFn>();
  lbx_check(lbx_mf_equilibrium(f.mf(), lvl.now.get<Density>().mf(), velocity.at(level).mf()), "CalcEquilibriumDist");
  f.touch();
  UpdateBoundaries(level);
}

// src/AmrSim.cpp:938-979
void AmrSim::CalcHydroVars(int const level) {
  FinishPeerStores();
  auto& state = levels.at(level).now;
  Density::fill(state.get<Density>(), velocity.at(level), state.get<DistFn>());
}

// src/AmrSim.cpp:19-23
void AmrSim::UpdateBoundaries(int const level) {
  amrex::FillBoundary(levels.at(level).next.get<DistFn>(), geom[level].periodicity());
}

// src/AmrSim.cpp:25-107 (valid cells, in place)
void AmrSim::Collide(MultiFab& f, const double omega_s, const double omega_b) {
  lbx_check(lbx_mf_collide(f.mf(), omega_s, omega_b, nullptr, FINE_VAL), "Collide");
  f.touch();
}

// src/AmrSim.cpp:109-122: pull-stream NEXT into a "fresh" fab over valid grown by one, swap
void AmrSim::Stream(int const level) {
  MultiFab& f_nxt = levels[level].next.get<DistFn>();
  if (valid_pending.at(level)) {      // only if a caller streams without colliding first
    amrex::CopyValid(f_nxt, levels[level].now.get<DistFn>());
    valid_pending.at(level) = false;
  }
  MultiFab& f_prop = stream_scratch.at(level);
  if (f_prop.empty() || f_prop.boxArray() != f_nxt.boxArray() || f_prop.layout() != f_nxt.layout())
    f_prop = field_traits<DistFn>::MakeLevelData(f_nxt.boxArray(), f_nxt.DistributionMap(), f_nxt.layout());
  DistFn::Propagate(f_nxt, f_prop);
  std::swap(f_nxt, f_prop);
}

// src/AmrSim.cpp:124-135
void AmrSim::CollideLevel(int const level) {
  const double omega_s = 1.0 / (tau_s.at(level) + 0.5);
  const double omega_b = 1.0 / (tau_b.at(level) + 0.5);
  MultiFab& f_pc = levels.at(level).next.get<DistFn>();
  const MultiFab& f_now = levels.at(level).now.get<DistFn>();
  const bool same_boxes = f_pc.boxArray() == f_now.boxArray() && f_pc.layout() == f_now.layout();
  if (level == 0 && same_boxes) {
    // FillPatch's valid-cell copy fused into the collision (next <- collide(now)); its ghost
    // fill is dead work here because the FillBoundary below rewrites every ghost cell
    lbx_check(lbx_mf_collide2(f_now.mf(), f_pc.mf(), omega_s, omega_b, nullptr, FINE_VAL), "CollideLevel");
    f_pc.touch();
  } else if (same_boxes && !f_now.isFlat()) {
    // finer level: FillPatch writes the ghost cells only (the coarse-fine ones survive the
    // FillBoundary below), the valid-cell copy is again fused into the collision
    FillPatchImpl(level, f_pc, true);
    lbx_check(lbx_mf_collide2(f_now.mf(), f_pc.mf(), omega_s, omega_b, nullptr, FINE_VAL), "CollideLevel");
    f_pc.touch();
  } else {
    DistFnFillPatch(level, f_pc);
    Collide(f_pc, omega_s, omega_b);
  }
  amrex::FillBoundary(f_pc, geom[level].periodicity());
}

// include/AmrSim.h:89-94.  FLAT storage: CollideLevel + Stream collapse into one fused launch
// (15 loads + 15 stores per cell); the result lands in NEXT exactly as after the reference's
// Stream, then UpdateNow swaps the states.
void AmrSim::CollideAndStream(int const level) {
  auto& lvl = levels.at(level);
  MultiFab& now_f = lvl.now.get<DistFn>();
  if (now_f.isFlat()) {
    MultiFab& next_f = lvl.next.get<DistFn>();
    const lbx_box box = to_lbx(geom[level].Domain());
    lbx_domain dom;
    for (int d = 0; d < 3; ++d) {
      dom.lo[d] = box.lo[d];
      dom.hi[d] = box.hi[d];
      dom.periodic[d] = geom[level].isPeriodic(d) ? 1 : 0;
    }
    const double omega_s = 1.0 / (tau_s.at(level) + 0.5), omega_b = 1.0 / (tau_b.at(level) + 0.5);
    if (has_walls) {
      if (now_f.numStorageFabs() > 1) amrex::Abort("walls are not available in a distributed run");
      for (int d = 0; d < 3; ++d)
        if (!geom[level].isPeriodic(d)) dom.periodic[d] = LBX_BC_WALL;
    }
    if (now_f.numStorageFabs() > 1) {
      // one slab per rank: the FillBoundary between boxes of different GPUs (src/AmrSim.cpp:132) is fused into
      // the step -- boundary-plane CTAs store the face-crossing populations into the neighbours' NEXT over
      // NVLink and order themselves against the neighbours' steps; one launch per step and rank
      lbx_check(lbx_mf_collide_stream_slab(now_f.mf(), next_f.mf(), &dom, omega_s, omega_b), "CollideAndStream");
      peer_stores_pending = true;
    } else {
      const lbx_fab src = now_f.fabDesc(0), dst = next_f.fabDesc(0);
      lbx_check(lbx_collide_stream(&src, &dst, &box, &dom, omega_s, omega_b, LBX_PUSH), "CollideAndStream");
    }
    next_f.touch();
  } else if (has_walls) {
    amrex::Abort("walls run on the uniform single-GPU path only (level 0 alone, uniform fast path on)");
  } else if (rohde_fused && CanFuseLevelStep(level)) {
    LevelStepFused(level);
  } else {
    CollideLevel(level);
    Stream(level);
  }
  lvl.UpdateNow();
}

// CollideLevel + Stream in one pass over the level (per-box storage, any level).  Same values per
// cell as the literal sequence: valid cells collide(NOW) pushed; ghost cells FillPatch would take from
// same-level valid cells push those cells' collided values (what FillBoundary leaves after Collide),
// ghost cells under the coarse level push the coarse state, interpolated in time under SUBCYCLE.
bool AmrSim::CanFuseLevelStep(int const level) const {
  const MultiFab& a = levels[level].now.get<DistFn>();
  const MultiFab& b = levels[level].next.get<DistFn>();
  return !(a.empty() || b.empty() || a.isFlat() || b.isFlat() || a.nGrow() != 2 || b.nGrow() != 2 ||
           a.boxArray() != b.boxArray());
}

void AmrSim::LevelStepFused(int const level) {
  MultiFab& f_nxt = levels[level].next.get<DistFn>();
  const MultiFab& f_now = levels[level].now.get<DistFn>();
  MultiFab& f_prop = stream_scratch.at(level);
  if (f_prop.empty() || f_prop.boxArray() != f_nxt.boxArray() || f_prop.layout() != f_nxt.layout())
    f_prop = field_traits<DistFn>::MakeLevelData(f_nxt.boxArray(), f_nxt.DistributionMap(), f_nxt.layout());
  amrex::GhostPush push;
  push.src_valid = &f_now;
  push.fallback = &f_nxt;
  push.omega_s = 1.0 / (tau_s.at(level) + 0.5);
  push.omega_b = 1.0 / (tau_b.at(level) + 0.5);
  push.level_step = true;
  if (!level) {
    amrex::FillPatchSingleLevel(f_prop, f_now, geom[level], true, &push);
  } else {
    const MultiFab* ca = nullptr;
    CoarseStatesAt(level - 1, levels[level].time.current, ca, push.wa, push.crse_b, push.wb);
    amrex::FillPatchTwoLevels(f_prop, *ca, f_now, geom[level - 1], geom[level], refRatio(level - 1), true, &push);
  }
  std::swap(f_nxt, f_prop);          // NEXT = the streamed state, as after the reference's Stream
}

// src/AmrSim.cpp:324-333
void AmrSim::IterateLevel(int const level) {
  CollideAndStream(level);
  auto& time = levels.at(level).time;
  time.current += time.delta;
  ++time.step;
}

// src/AmrSim.cpp:335-344 (never called by the reference either)
void AmrSim::SubCycle(int const base_level, int const num_steps) {
  if (base_level == finest_level) {
    for (int iter = 0; iter < num_steps; ++iter) IterateLevel(base_level);
  } else {
    IterateLevel(base_level);
    SubCycle(base_level + 1, refRatio(base_level)[0]);
  }
}

// src/AmrSim.cpp:297-322
void AmrSim::ComputeDt(int const level) {
  auto& fine_time = levels[level].time;
  if (level) {
    const int r = refRatio(level - 1)[0];
    fine_time.delta = levels[level - 1].time.delta / r;
    mass.at(level) = mass.at(level - 1) / r;
    tau_s.at(level) = r * (tau_s.at(level - 1) - 0.5) + 0.5;
    tau_b.at(level) = r * (tau_b.at(level - 1) - 0.5) + 0.5;
  } else {
    fine_time.delta = 1.0;
    mass.at(level) = 1.0;
  }
}

// src/AmrSim.cpp:359-391
void AmrSim::DistFnFillPatch(int const level, MultiFab& dest) { FillPatchImpl(level, dest, false); }

// ghosts_only: dest's valid cells are left for an out-of-place collision to produce
void AmrSim::FillPatchImpl(int const level, MultiFab& dest, bool ghosts_only, const amrex::GhostPush* push) {
  if (!level) {
    amrex::FillPatchSingleLevel(dest, levels[level].now.get<DistFn>(), geom[level], ghosts_only, push);
  } else {
    amrex::FillPatchTwoLevels(dest, CoarseStateAt(level - 1, levels[level].time.current), levels[level].now.get<DistFn>(),
                              geom[level - 1], geom[level], refRatio(level - 1), ghosts_only, push);
  }
}

// The reference hands FillPatchTwoLevels ONE coarse state, NOW, whatever its time (src/AmrSim.cpp:373-389).
// The conventional driver advances the coarse level first, so at fine time t the coarse level holds
// two states -- NOW at t1 and, in NEXT since UpdateNow's swap, the old one at t0 = t1 - dt -- and
// FillPatchTwoLevels interpolates between them [AMReX: state 0 or 1 when t is within 1e-3 dt of its
// time, else LinComb((t1-t)/(t1-t0), old, (t-t0)/(t1-t0), new)].
void AmrSim::CoarseStatesAt(int const coarse_level, double const t, const MultiFab*& a, double& wa, const MultiFab*& b,
                            double& wb) {
  auto& crse = levels.at(coarse_level);
  const MultiFab& f_new = crse.now.get<DistFn>();
  a = &f_new; wa = 1.0; b = nullptr; wb = 0.0;
  if (coupling != Coupling::SUBCYCLE) return;
  const double t1 = crse.time.current, dt = crse.time.delta, t0 = t1 - dt, eps = 1e-3 * dt;
  if (std::abs(t - t1) < eps || crse.time.step == 0) return;
  const MultiFab& f_old = crse.next.get<DistFn>();
  if (f_old.empty() || f_old.boxArray() != f_new.boxArray() || f_old.layout() != f_new.layout())
    amrex::Abort("CoarseStateAt: the coarse level's old state is not available");
  a = &f_old;
  if (std::abs(t - t0) < eps) return;
  if (t < t0 - eps || t > t1 + eps) amrex::Abort("CoarseStateAt: fine time outside the coarse step");
  wa = (t1 - t) / (t1 - t0);
  b = &f_new;
  wb = (t - t0) / (t1 - t0);
}

const MultiFab& AmrSim::CoarseStateAt(int const coarse_level, double const t) {
  auto& crse = levels.at(coarse_level);
  const MultiFab& f_new = crse.now.get<DistFn>();
  if (coupling != Coupling::SUBCYCLE) return f_new;
  const double t1 = crse.time.current, dt = crse.time.delta, t0 = t1 - dt, eps = 1e-3 * dt;
  if (std::abs(t - t1) < eps || crse.time.step == 0) return f_new;
  const MultiFab& f_old = crse.next.get<DistFn>();
  if (f_old.empty() || f_old.boxArray() != f_new.boxArray() || f_old.layout() != f_new.layout())
    amrex::Abort("CoarseStateAt: the coarse level's old state is not available");
  if (std::abs(t - t0) < eps) return f_old;
  if (t < t0 - eps || t > t1 + eps) amrex::Abort("CoarseStateAt: fine time outside the coarse step");
  if (coarse_interp.size() < levels.size()) coarse_interp.resize(levels.size());
  MultiFab& tmp = coarse_interp[coarse_level];
  if (tmp.empty() || tmp.boxArray() != f_new.boxArray() || tmp.layout() != f_new.layout())
    tmp = field_traits<DistFn>::MakeLevelData(f_new.boxArray(), f_new.DistributionMap(), f_new.layout());
  amrex::LinComb(tmp, (t1 - t) / (t1 - t0), f_old, (t - t0) / (t1 - t0), f_new);
  return tmp;
}

// amrex::average_down of the populations: coarse NOW valid cells under fine valid cells <- mean of
// the ratio^3 fine NOW cells.
void AmrSim::AverageDown(int const coarse_level) {
  amrex::average_down(levels.at(coarse_level + 1).now.get<DistFn>(), levels.at(coarse_level).now.get<DistFn>(), 0, NMODES,
                      refRatio(coarse_level));
}

// Coupling::SUBCYCLE.  Unlike the reference's dead SubCycle (above), an intermediate level takes
// `ratio` steps per step of its parent, and each group of fine steps ends with an average_down.
void AmrSim::SubCycleAdvance(int const level) {
  IterateLevel(level);
  if (level < finest_level) {
    const int r = refRatio(level)[0];
    for (int iter = 0; iter < r; ++iter) SubCycleAdvance(level + 1);
    AverageDown(level);
  }
}

// src/AmrSim.cpp:393-411
void AmrSim::DistFnFillFromCoarse(int const level, MultiFab& fine_mf) {
  if (!level) amrex::Abort("Cannot fill level 0 from coarse.");
  amrex::InterpFromCoarseLevel(fine_mf, levels[level - 1].now.get<DistFn>(), geom[level - 1], geom[level],
                               refRatio(level - 1));
}

// src/AmrSim.cpp:413-417
bool AmrSim::TagCell(int const level, const IntVect& pos) { return static_tags.at(level).contains(pos); }

// src/AmrSim.cpp:419-428 -- sic: the level's OWN BoxArray is handed over as the fine one
// (SURVEY.md B-1), so the mask marks coarsen(level grids), not the cells under level+1.
void AmrSim::MakeFineMask(int const coarse_level) {
  const MultiFab& cmf = levels[coarse_level].now.get<DistFn>();
  const BoxArray& fba = levels[coarse_level].now.get<DistFn>().boxArray();
  fine_masks[coarse_level] = amrex::makeFineMask(cmf, fba, refRatio(coarse_level), COARSE_VAL, FINE_VAL);
}

// ----------------------------------------------------------------------------- Rohde cycle
// src/AmrSim.cpp:430-469
void AmrSim::RohdeCycle(int const coarse_level) {
  if (rohde_fused && CanFuseRohde(coarse_level)) return RohdeCycleFused(coarse_level);
  const int ref_ratio_here = refRatio(coarse_level)[0];
  InitPostCollision(coarse_level);
  CoarseCollide(coarse_level);
  if (coarse_level + 1 == finest_level) {
    InitPostCollision(finest_level);
    FineCollide(finest_level);
    Stream(finest_level);
    FineCollide(finest_level);
    Stream(finest_level);
    ZeroInvalidComponents(finest_level);
    UpdateDistribution(finest_level);
  } else {
    for (int iter = 0; iter < ref_ratio_here; ++iter) RohdeCycle(coarse_level + 1);
  }
  Stream(coarse_level);
  SumFromFine(coarse_level);
  ZeroInvalidComponents(coarse_level);
  if (coarse_level == 0) UpdateBoundaries(coarse_level);
  UpdateDistribution(coarse_level);
}

// The same cycle with each level's collide + Stream pair fused into one pass over the level
// (lbx_mf_collide_stream): every pass of the reference's sequence is still performed on the same
// data in the same order per cell, so the result is identical to the unfused sequence above.
//  * CoarseCollide is deferred to just before Stream(coarse): between the two the reference only
//    works on finer levels, which read this level's NOW (FillPatchTwoLevels :385), never its NEXT.
//  * ZeroInvalidComponents is folded into the stores of the cycle's last Stream (on the coarse
//    level it commutes with SumFromFine, which adds into valid cells only).
//  * InitPostCollision's zeroing of component 0 in the outermost ghost ring (:477-482) is not
//    observable: the rest population never moves, and Stream leaves ring 2 of its fresh
//    destination at the fill value.
bool AmrSim::CanFuseRohde(int const level) const {
  for (int l = level; l <= finest_level; ++l) {
    const MultiFab& a = levels[l].now.get<DistFn>();
    const MultiFab& b = levels[l].next.get<DistFn>();
    if (a.empty() || b.empty() || a.isFlat() || b.isFlat() || a.nGrow() != 2 || b.nGrow() != 2 ||
        a.boxArray() != b.boxArray())
      return false;
  }
  return true;
}

// One collide + Stream of a level in fused form.  Valid cells: read from NOW while InitPostCollision's
// valid-cell copy is still pending (first pass of a cycle), else from NEXT.  Ghost cells:
//   from_fillpatch: the level's DistFnFillPatch is folded in -- each ghost cell pushes straight from
//                   the NOW cell (same level, periodic image, or the coarse cell under it) FillPatch
//                   would copy into it (lbx_mf_collide_stream_fillpatch); NEXT's ghost cells are never written;
//   otherwise     : NEXT's own ghost cells, as left by the previous Stream (second fine pass).
void AmrSim::CollideStreamFused(int const level, bool masked, bool zero_invalid, bool from_fillpatch) {
  MultiFab& f_nxt = levels[level].next.get<DistFn>();
  const MultiFab& f_now = levels[level].now.get<DistFn>();
  const MultiFab& vsrc = valid_pending.at(level) ? f_now : f_nxt;
  MultiFab& f_prop = stream_scratch.at(level);
  if (f_prop.empty() || f_prop.boxArray() != f_nxt.boxArray() || f_prop.layout() != f_nxt.layout())
    f_prop = field_traits<DistFn>::MakeLevelData(f_nxt.boxArray(), f_nxt.DistributionMap(), f_nxt.layout());
  const amrex::iMultiFab* mask = nullptr;
  if (masked) {
    mask = &fine_masks.at(level);
    if (mask->empty()) amrex::Abort("CoarseCollide: no fine mask on this level");
  }
  const double omega_s = 1.0 / (tau_s.at(level) + 0.5), omega_b = 1.0 / (tau_b.at(level) + 0.5);
  if (from_fillpatch) {
    amrex::GhostPush push;
    push.src_valid = &vsrc;
    push.mask = mask;
    push.fallback = &f_nxt;
    push.omega_s = omega_s;
    push.omega_b = omega_b;
    push.fine_val = FINE_VAL;
    push.zero_invalid = zero_invalid;
    FillPatchImpl(level, f_prop, true, &push);
  } else {
    lbx_check(lbx_mf_collide_stream(vsrc.mf(), f_nxt.mf(), f_prop.mf(), omega_s, omega_b, mask ? mask->mf() : nullptr, FINE_VAL,
                                    zero_invalid ? 1 : 0),
              "CollideStreamFused");
  }
  valid_pending.at(level) = false;
  std::swap(f_nxt, f_prop);
  f_nxt.touch();
}

void AmrSim::RohdeCycleFused(int const coarse_level) {
  const int ref_ratio_here = refRatio(coarse_level)[0];
  // InitPostCollision: the valid-cell copy is folded into the collision, the ghost-cell fill into
  // the ghost cells' pushes (CollideStreamFused) -- on every level nothing but that one Stream ever
  // reads the ghost cells FillPatch writes.
  valid_pending.at(coarse_level) = true;
  if (coarse_level + 1 == finest_level) {
    valid_pending.at(finest_level) = true;
    CollideStreamFused(finest_level, false, false, true);    // InitPostCollision + FineCollide + Stream
    CollideStreamFused(finest_level, false, true, false);    // FineCollide + Stream + ZeroInvalidComponents
    UpdateDistribution(finest_level);
  } else {
    for (int iter = 0; iter < ref_ratio_here; ++iter) RohdeCycle(coarse_level + 1);
  }
  // InitPostCollision + CoarseCollide + Stream + ZeroInvalidComponents
  CollideStreamFused(coarse_level, true, true, true);
  SumFromFine(coarse_level);
  // UpdateBoundaries(0) writes level-0 ghost cells that every following cycle overwrites before
  // reading (its FillPatch): only the cycle that ends Iterate() needs to leave them filled
  if (coarse_level == 0 && !defer_boundaries) UpdateBoundaries(coarse_level);
  UpdateDistribution(coarse_level);
}

// src/AmrSim.cpp:471-485
void AmrSim::InitPostCollision(int const level) {
  MultiFab& f_pc = levels[level].next.get<DistFn>();
  // ghost cells now; the valid cells (= NOW's valid cells) are produced by the collision that
  // always follows, out of place (CoarseCollide / FineCollide below)
  const MultiFab& f_now = levels[level].now.get<DistFn>();
  const bool fuse = f_pc.boxArray() == f_now.boxArray() && f_pc.layout() == f_now.layout() && !f_now.isFlat();
  FillPatchImpl(level, f_pc, fuse);
  valid_pending.at(level) = fuse;
  if (level != 0) {
    // sic: component 0 only, outermost ghost ring
    lbx_check(lbx_mf_zero_ring(f_pc.mf(), 1, 0), "InitPostCollision");
    f_pc.touch();
  }
}

// src/AmrSim.cpp:487-580
void AmrSim::CoarseCollide(int const level) {
  MultiFab& f_pc = levels[level].next.get<DistFn>();
  const amrex::iMultiFab& mask = fine_masks.at(level);
  if (mask.empty()) amrex::Abort("CoarseCollide: no fine mask on this level");
  const MultiFab& src = valid_pending.at(level) ? levels[level].now.get<DistFn>() : f_pc;
  lbx_check(lbx_mf_collide2(src.mf(), f_pc.mf(), 1.0 / (tau_s.at(level) + 0.5), 1.0 / (tau_b.at(level) + 0.5), mask.mf(),
                            FINE_VAL),
            "CoarseCollide");
  valid_pending.at(level) = false;
  f_pc.touch();
}

// src/AmrSim.cpp:582-590
void AmrSim::FineCollide(int const level) {
  MultiFab& f_pc = levels.at(level).next.get<DistFn>();
  const MultiFab& src = valid_pending.at(level) ? levels[level].now.get<DistFn>() : f_pc;
  lbx_check(lbx_mf_collide2(src.mf(), f_pc.mf(), 1.0 / (tau_s.at(level) + 0.5), 1.0 / (tau_b.at(level) + 0.5), nullptr,
                            FINE_VAL),
            "FineCollide");
  valid_pending.at(level) = false;
  f_pc.touch();
}

// src/AmrSim.cpp:592-602
void AmrSim::SumFromFine(int const coarse_level) {
  amrex::sum_fine_to_coarse(levels[coarse_level + 1].now.get<DistFn>(), levels[coarse_level].next.get<DistFn>(), 0, NMODES,
                            refRatio(coarse_level), geom[coarse_level], geom[coarse_level + 1]);
}

// src/AmrSim.cpp:604-617
void AmrSim::ZeroInvalidComponents(int const level) {
  MultiFab& f_pc = levels[level].next.get<DistFn>();
  lbx_check(lbx_mf_zero_invalid(f_pc.mf()), "ZeroInvalidComponents");
  f_pc.touch();
}

// src/AmrSim.cpp:619-631
void AmrSim::UpdateDistribution(int const level) {
  auto& lvl = levels[level];
  if (level && level == finest_level) {
    lvl.time.current += 2 * lvl.time.delta;
    lvl.time.step += 2;
  } else {
    lvl.time.current += lvl.time.delta;
    ++lvl.time.step;
  }
  std::swap(lvl.now.get<DistFn>(), lvl.next.get<DistFn>());
}

// src/AmrSim.cpp:981-993 (+ the regrid_int hook)
void AmrSim::Iterate(int const nsteps) {
  for (int t = 0; t < nsteps; ++t) {
    if (!finest_level) {
      SetLevelLayout(0, PreferredLayout(0));
      IterateLevel(0);
    } else {
      if (has_walls) amrex::Abort("walls run on the uniform single-GPU path only: no refined levels");
      for (int l = 0; l <= finest_level; ++l) SetLevelLayout(l, Layout::BOXES);
      if (coupling == Coupling::SUBCYCLE) {
        SubCycleAdvance(0);
      } else {
        defer_boundaries = (t + 1 < nsteps);
        RohdeCycle(0);
        defer_boundaries = false;
      }
    }
    RegridIfDue();
  }
  FinishPeerStores();
}

// AMReX's regrid_int: after every regrid_int-th coarse step, regrid from level 0 up (tags from
// ErrorEst: static boxes and/or the gradient criterion); the fine masks follow the new grids.
void AmrSim::RegridIfDue() {
  if (regrid_int <= 0 || max_level == 0) return;
  if (++steps_since_regrid < regrid_int) return;
  steps_since_regrid = 0;
  ++num_regrids;
  regrid(0, GetTime(0));
  for (int l = 0; l < finest_level; ++l) MakeFineMask(l);
}

// ----------------------------------------------------------------------------- regrid hooks
// src/AmrSim.cpp:633-663
void AmrSim::ErrorEst(int level, amrex::TagBoxArray& tba, double /*time*/, int /*ngrow*/) {
  const MultiFab& f = levels[level].now.get<DistFn>();
  for (amrex::MFIter mfi(f); mfi.isValid(); ++mfi) {
    const Box box = mfi.validbox();
    amrex::TagBox& tagfab = tba[mfi];
    tagfab.setVal(amrex::TagBox::CLEAR, box);
    for (const auto& hit : static_tags.at(level).intersections(box)) tagfab.setVal(amrex::TagBox::SET, hit.second);
  }
  if (gradient_threshold.at(level) > 0.0) GradientTags(level, tba);
}

// Gradient criterion: rho of the level's NOW with one ghost cell filled the way DistFnFillPatch fills
// the populations (same level + periodic images, else piecewise constant from the coarse level),
// one tagging launch, one device->host copy of the int tags.
void AmrSim::GradientTags(int const level, amrex::TagBoxArray& tba) {
  CalcHydroVars(level);
  const MultiFab& rho = levels[level].now.get<Density>();
  MultiFab rho_g(rho.boxArray(), rho.DistributionMap(), 1, 1);
  if (level == 0) {
    amrex::FillPatchSingleLevel(rho_g, rho, geom[0]);
  } else {
    CalcHydroVars(level - 1);
    amrex::FillPatchTwoLevels(rho_g, levels[level - 1].now.get<Density>(), rho, geom[level - 1], geom[level],
                              refRatio(level - 1));
  }
  amrex::iMultiFab tg(rho.boxArray(), rho.DistributionMap(), 1, 0);
  tg.setVal(0);
  lbx_check(lbx_mf_tag_gradient(rho_g.mf(), gradient_threshold[level], tg.mf(), 1), "ErrorEst (gradient)");
  tg.touch();
  const std::vector<int>& m = tg.hostMirror();
  for (amrex::MFIter mfi(tg); mfi.isValid(); ++mfi) {
    amrex::TagBox& tagfab = tba[mfi];
    if (!tagfab.allocated()) continue;           // another rank's box: its owner tags it
    const Box box = mfi.validbox();
    const int* t = m.data() + tg.storageOffset(mfi.index());
    for (int k = box.smallEnd(2); k <= box.bigEnd(2); ++k)
      for (int j = box.smallEnd(1); j <= box.bigEnd(1); ++j)
        for (int i = box.smallEnd(0); i <= box.bigEnd(0); ++i)
          if (*t++) tagfab(IntVect(i, j, k)) = amrex::TagBox::SET;
  }
}

// src/AmrSim.cpp:665-687
void AmrSim::MakeNewLevelFromScratch(int level, double time, const BoxArray& ba, const DistributionMapping& dm) {
  auto& lvl = levels[level];
  const Layout lay = (level == 0) ? PreferredLayout(0, ba, dm) : Layout::BOXES;
  velocity[level].define(ba, dm, NDIMS, 0, lay);
  lvl.Define(ba, dm, lay);
  stream_scratch[level].clear();
  lvl.time.current = time;
  ComputeDt(level);
  lvl.time.step = 0;
  if (!level) {
    InitDensity(level);
    InitVelocity(level);
    CalcEquilibriumDist(level);
  }
}

// src/AmrSim.cpp:689-713
void AmrSim::MakeNewLevelFromCoarse(int level, double time, const BoxArray& ba, const DistributionMapping& dm) {
  if (!level) amrex::Abort("Cannot construct level 0 from a coarser level.");
  velocity[level].define(ba, dm, NDIMS, 0);
  auto& lvl = levels[level];
  lvl.Define(ba, dm);
  stream_scratch[level].clear();
  lvl.time.current = time;
  ComputeDt(level);
  lvl.time.step = 0;
  DistFnFillFromCoarse(level, lvl.now.get<DistFn>());
  CalcHydroVars(level);
  MakeFineMask(level - 1);
}

// src/AmrSim.cpp:715-744.  DEVIATION: the reference swaps in new NOW fabs only and
// leaves NEXT on the old BoxArray, which breaks the following step whenever the grids really
// changed; NEXT is redefined on the new BoxArray here.
void AmrSim::RemakeLevel(int level, double time, const BoxArray& ba, const DistributionMapping& dm) {
  auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!getenv("LBX_HOST_TIMING")) return;
    const auto T1 = std::chrono::steady_clock::now();
    std::cerr << "    [RemakeLevel " << level << "] " << what << " " << std::chrono::duration<double>(T1 - T0).count() << " s\n";
    T0 = T1;
  };
  auto& lvl = levels[level];
  const Layout lay = lvl.now.get<DistFn>().empty() ? Layout::BOXES : lvl.now.get<DistFn>().layout();
  MultiFab new_u(ba, dm, NDIMS, 0, lay);
  MultiFab new_f = field_traits<DistFn>::MakeLevelData(ba, dm, lay);
  MultiFab new_rho = field_traits<Density>::MakeLevelData(ba, dm, lay);
  lap("allocate u, f, rho");
  DistFnFillPatch(level, new_f);
  lap("FillPatch (plan + launch)");
  auto& state = lvl.now;
  std::swap(new_f, state.get<DistFn>());
  std::swap(new_rho, state.get<Density>());
  std::swap(new_u, velocity[level]);
  lvl.next.Define(ba, dm, lay);
  stream_scratch[level].clear();
  lap("swap, redefine NEXT, clear scratch");
  lvl.time.current = time;
  CalcHydroVars(level);
  if (level < finest_level) MakeFineMask(level);
  lap("moments, fine mask");
  new_f.clear();
  new_rho.clear();
  new_u.clear();
  lap("release the old level");
}

// src/AmrSim.cpp:746-751
void AmrSim::ClearLevel(int level) {
  velocity.at(level).clear();
  levels.at(level).Clear();
  stream_scratch.at(level).clear();
}

// src/AmrSim.cpp:995-1009
void AmrSim::SetStaticRefinement(int const level, const std::array<int, NDIMS>& lo_corner,
                                 const std::array<int, NDIMS>& hi_corner) {
  static_tags.at(level).define(Box(IntVect(lo_corner), IntVect(hi_corner)));
  const auto T0 = std::chrono::steady_clock::now();
  regrid(level, GetTime(level));
  const auto T1 = std::chrono::steady_clock::now();
  MakeFineMask(level);
  if (getenv("LBX_HOST_TIMING"))
    std::cerr << "[SetStaticRefinement " << level << "] regrid " << std::chrono::duration<double>(T1 - T0).count()
              << " s, MakeFineMask " << std::chrono::duration<double>(std::chrono::steady_clock::now() - T1).count() << " s\n";
}

void AmrSim::SetStaticBox(int const level, const std::array<int, NDIMS>& lo_corner, const std::array<int, NDIMS>& hi_corner) {
  static_tags.at(level).define(Box(IntVect(lo_corner), IntVect(hi_corner)));
}
void AmrSim::Regrid() {
  if (max_level == 0) return;
  ++num_regrids;
  regrid(0, GetTime(0));
  for (int l = 0; l < finest_level; ++l) MakeFineMask(l);
}

// src/AmrSim.cpp:1011-1017
void AmrSim::UnsetStaticRefinement(int const level) {
  static_tags.at(level).clear();
  regrid(level, GetTime(level));
  MakeFineMask(level);
}

void AmrSim::SetGradientRefinement(int const level, double const threshold) {
  if (!(threshold > 0.0)) amrex::Abort("SetGradientRefinement: threshold must be positive");
  gradient_threshold.at(level) = threshold;
  regrid(level, GetTime(level));
  MakeFineMask(level);
}

void AmrSim::UnsetGradientRefinement(int const level) {
  gradient_threshold.at(level) = 0.0;
  regrid(level, GetTime(level));
  MakeFineMask(level);
}

// ----------------------------------------------------------------------------- checkpoint / restart
namespace {
constexpr char CHK_MAGIC[8] = {'L', 'B', 'X', 'C', 'H', 'K', '2', 0};
constexpr uint32_t CHK_ENDIAN = 0x01020304u;     // written natively after the magic: a byte-swapped reader sees 0x04030201
struct ChkFile {
  std::FILE* f;
  explicit ChkFile(const std::string& path, const char* mode) : f(std::fopen(path.c_str(), mode)) {
    if (!f) amrex::Abort(("checkpoint: cannot open " + path).c_str());
  }
  ~ChkFile() { if (f) std::fclose(f); }
  void put(const void* p, size_t n) { if (std::fwrite(p, 1, n, f) != n) amrex::Abort("checkpoint: short write"); }
  void get(void* p, size_t n) { if (std::fread(p, 1, n, f) != n) amrex::Abort("checkpoint: short read / truncated file"); }
  template <class T> void put(const T& v) { put(&v, sizeof(T)); }
  template <class T> T get() { T v; get(&v, sizeof(T)); return v; }
};
}  // namespace

void AmrSim::WritePlotFile(const std::string& dir) {
  if (DistributionMapping::NProcs() > 1) amrex::Abort("WritePlotFile: single-process runs only");
  auto mkdir_p = [](const std::string& d) {
    if (::mkdir(d.c_str(), 0755) != 0 && errno != EEXIST) amrex::Abort(("WritePlotFile: cannot create " + d).c_str());
  };
  auto boxstr = [](const Box& b) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "((%d,%d,%d) (%d,%d,%d) (0,0,0))", b.smallEnd(0), b.smallEnd(1), b.smallEnd(2), b.bigEnd(0),
                  b.bigEnd(1), b.bigEnd(2));
    return std::string(buf);
  };
  mkdir_p(dir);
  const char* names[4] = {"rho", "ux", "uy", "uz"};
  std::FILE* h = std::fopen((dir + "/Header").c_str(), "w");
  if (!h) amrex::Abort(("WritePlotFile: cannot open " + dir + "/Header").c_str());
  std::fprintf(h, "HyperCLaw-V1.1\n4\n");
  for (const char* n : names) std::fprintf(h, "%s\n", n);
  std::fprintf(h, "3\n%.17g\n%d\n0 0 0\n1 1 1\n", GetTime(0), finest_level);
  for (int l = 0; l < finest_level; ++l) std::fprintf(h, "%d ", refRatio(l)[0]);
  std::fprintf(h, "\n");
  for (int l = 0; l <= finest_level; ++l) std::fprintf(h, "%s ", boxstr(geom[l].Domain()).c_str());
  std::fprintf(h, "\n");
  for (int l = 0; l <= finest_level; ++l) std::fprintf(h, "%d ", GetTimeStep(l));
  std::fprintf(h, "\n");
  for (int l = 0; l <= finest_level; ++l) {
    const Box& d = geom[l].Domain();
    std::fprintf(h, "%.17g %.17g %.17g\n", 1.0 / d.length(0), 1.0 / d.length(1), 1.0 / d.length(2));
  }
  std::fprintf(h, "0\n0\n");
  for (int l = 0; l <= finest_level; ++l) {
    CalcHydroVars(l);
    const MultiFab& rho = levels[l].now.get<Density>();
    const MultiFab& u = velocity[l];
    const std::vector<double>& hr = rho.hostMirror();
    const std::vector<double>& hu = u.hostMirror();
    const int nb = rho.numStorageFabs();
    const Box& d = geom[l].Domain();
    std::fprintf(h, "%d %d %.17g\n%d\n", l, nb, GetTime(l), GetTimeStep(l));
    for (int s = 0; s < nb; ++s) {
      const Box b = rho.storageValid(s);
      for (int a = 0; a < 3; ++a)
        std::fprintf(h, "%.17g %.17g\n", (double)(b.smallEnd(a) - d.smallEnd(a)) / d.length(a),
                     (double)(b.bigEnd(a) + 1 - d.smallEnd(a)) / d.length(a));
    }
    std::fprintf(h, "Level_%d/Cell\n", l);
    const std::string ldir = dir + "/Level_" + std::to_string(l);
    mkdir_p(ldir);
    std::FILE* fd = std::fopen((ldir + "/Cell_D_00000").c_str(), "wb");
    if (!fd) amrex::Abort("WritePlotFile: cannot open the FAB file");
    std::vector<long> offsets(nb);
    std::vector<double> mins((size_t)nb * 4), maxs((size_t)nb * 4);
    for (int s = 0; s < nb; ++s) {
      const Box b = rho.storageValid(s);
      const size_t n = (size_t)b.numPts();
      offsets[s] = std::ftell(fd);
      std::fprintf(fd, "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))%s 4\n", boxstr(b).c_str());
      // density and velocity have no ghost cells: storage fab s = [comp][z][y][x] over its valid box (x rows may be
      // padded when row alignment is on: copy row by row)
      const lbx_fab fr = rho.fabDesc(s), fu = u.fabDesc(s);
      std::vector<double> comp(n);
      for (int c = 0; c < 4; ++c) {
        const std::vector<double>& src = c == 0 ? hr : hu;
        const lbx_fab& f = c == 0 ? fr : fu;
        const size_t base = (c == 0 ? rho.storageOffset(s) : u.storageOffset(s)) + (size_t)(c == 0 ? 0 : c - 1) * f.n[0] * f.n[1] * f.n[2];
        size_t q = 0;
        for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k)
          for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j) {
            const size_t row = base + (size_t)f.n[0] * ((size_t)(j - f.lo[1]) + (size_t)f.n[1] * (size_t)(k - f.lo[2])) + (size_t)(b.smallEnd(0) - f.lo[0]);
            std::memcpy(&comp[q], &src[row], sizeof(double) * (size_t)b.length(0));
            q += (size_t)b.length(0);
          }
        mins[(size_t)s * 4 + c] = *std::min_element(comp.begin(), comp.end());
        maxs[(size_t)s * 4 + c] = *std::max_element(comp.begin(), comp.end());
        if (std::fwrite(comp.data(), sizeof(double), n, fd) != n) amrex::Abort("WritePlotFile: short write");
      }
    }
    std::fclose(fd);
    std::FILE* ch = std::fopen((ldir + "/Cell_H").c_str(), "w");
    if (!ch) amrex::Abort("WritePlotFile: cannot open Cell_H");
    std::fprintf(ch, "1\n1\n4\n0\n(%d 0\n", nb);
    for (int s = 0; s < nb; ++s) std::fprintf(ch, "%s\n", boxstr(rho.storageValid(s)).c_str());
    std::fprintf(ch, ")\n%d\n", nb);
    for (int s = 0; s < nb; ++s) std::fprintf(ch, "FabOnDisk: Cell_D_00000 %ld\n", offsets[s]);
    std::fprintf(ch, "\n%d,4\n", nb);
    for (int s = 0; s < nb; ++s) {
      for (int c = 0; c < 4; ++c) std::fprintf(ch, "%.17g,", mins[(size_t)s * 4 + c]);
      std::fprintf(ch, "\n");
    }
    std::fprintf(ch, "\n%d,4\n", nb);
    for (int s = 0; s < nb; ++s) {
      for (int c = 0; c < 4; ++c) std::fprintf(ch, "%.17g,", maxs[(size_t)s * 4 + c]);
      std::fprintf(ch, "\n");
    }
    std::fclose(ch);
  }
  std::fclose(h);
}

void AmrSim::WriteCheckpoint(const std::string& path) {
  if (DistributionMapping::NProcs() > 1) amrex::Abort("WriteCheckpoint: single-process runs only");
  ChkFile c(path, "wb");
  c.put(CHK_MAGIC, sizeof(CHK_MAGIC));
  c.put<uint32_t>(CHK_ENDIAN);
  for (int v : {NX, NY, NZ, max_level, finest_level, (int)coupling, regrid_int, steps_since_regrid, num_regrids}) c.put<int32_t>(v);
  for (int l = 0; l <= max_level; ++l) {
    c.put<double>(tau_s[l]); c.put<double>(tau_b[l]); c.put<double>(mass[l]);
    c.put<double>(levels[l].time.current); c.put<double>(levels[l].time.delta); c.put<int32_t>(levels[l].time.step);
    c.put<double>(gradient_threshold[l]);
    c.put<int32_t>((int)static_tags[l].size());
    for (long q = 0; q < static_tags[l].size(); ++q)
      for (int d = 0; d < 3; ++d) { c.put<int32_t>(static_tags[l][q].smallEnd(d)); c.put<int32_t>(static_tags[l][q].bigEnd(d)); }
  }
  for (int l = 0; l <= finest_level; ++l) {
    const MultiFab& f = levels[l].now.get<DistFn>();
    const BoxArray& ba = f.boxArray();
    c.put<int32_t>((int)ba.size());
    for (long q = 0; q < ba.size(); ++q)
      for (int d = 0; d < 3; ++d) { c.put<int32_t>(ba[q].smallEnd(d)); c.put<int32_t>(ba[q].bigEnd(d)); }
    // valid cells only, box after box, [comp][z][y][x]: a ghost-free copy on the device, one download
    MultiFab tight(ba, f.DistributionMap(), NMODES, 0);
    amrex::CopyValid(tight, f);
    const std::vector<double>& h = tight.hostMirror();
    c.put<uint64_t>((uint64_t)h.size());
    c.put(h.data(), h.size() * sizeof(double));
  }
}

void AmrSim::ReadCheckpoint(const std::string& path) {
  if (DistributionMapping::NProcs() > 1) amrex::Abort("ReadCheckpoint: single-process runs only");
  ChkFile c(path, "rb");
  char magic[8];
  c.get(magic, sizeof(magic));
  if (std::memcmp(magic, CHK_MAGIC, sizeof(magic)) != 0) amrex::Abort("ReadCheckpoint: not a lambrex-b200 checkpoint (format 2)");
  if (c.get<uint32_t>() != CHK_ENDIAN) amrex::Abort("ReadCheckpoint: the checkpoint was written with another byte order");
  const int nx = c.get<int32_t>(), ny = c.get<int32_t>(), nz = c.get<int32_t>(), ml = c.get<int32_t>();
  if (nx != NX || ny != NY || nz != NZ || ml != max_level)
    amrex::Abort("ReadCheckpoint: the checkpoint was written by a simulation with other extents or max level");
  const int new_finest = c.get<int32_t>();
  if (new_finest < 0 || new_finest > max_level) amrex::Abort("ReadCheckpoint: corrupt file (finest level outside 0..max_level)");
  coupling = c.get<int32_t>() == (int)Coupling::SUBCYCLE ? Coupling::SUBCYCLE : Coupling::ROHDE;
  regrid_int = c.get<int32_t>();
  steps_since_regrid = c.get<int32_t>();
  num_regrids = c.get<int32_t>();
  for (int l = 0; l <= max_level; ++l) {
    tau_s[l] = c.get<double>(); tau_b[l] = c.get<double>(); mass[l] = c.get<double>();
    levels[l].time.current = c.get<double>(); levels[l].time.delta = c.get<double>(); levels[l].time.step = c.get<int32_t>();
    gradient_threshold[l] = c.get<double>();
    const int nst = c.get<int32_t>();
    if (nst < 0 || nst > (1 << 20)) amrex::Abort("ReadCheckpoint: corrupt file (static-tag box count)");
    amrex::BoxList bl;
    for (int q = 0; q < nst; ++q) {
      IntVect lo, hi;
      for (int d = 0; d < 3; ++d) { lo[d] = c.get<int32_t>(); hi[d] = c.get<int32_t>(); }
      if (!Box(lo, hi).ok()) amrex::Abort("ReadCheckpoint: corrupt file (empty static-tag box)");
      bl.push_back(Box(lo, hi));
    }
    if (nst) static_tags[l].define(bl); else static_tags[l].clear();
  }
  amrex::ClearPlanCache();
  for (int l = 0; l <= max_level; ++l) {
    const TimeData keep = levels[l].time;
    if (l > new_finest) {
      if (!levels[l].now.get<DistFn>().empty()) { ClearLevel(l); ClearBoxArray(l); ClearDistributionMap(l); }
      levels[l].time = keep;
      continue;
    }
    const int nb = c.get<int32_t>();
    if (nb < 1 || nb > (1 << 24)) amrex::Abort("ReadCheckpoint: corrupt file (box count of a level)");
    amrex::BoxList bl;
    for (int q = 0; q < nb; ++q) {
      IntVect lo, hi;
      for (int d = 0; d < 3; ++d) { lo[d] = c.get<int32_t>(); hi[d] = c.get<int32_t>(); }
      if (!Box(lo, hi).ok() || !geom[l].Domain().contains(Box(lo, hi)))
        amrex::Abort("ReadCheckpoint: corrupt file (a box is empty or outside its level's domain)");
      bl.push_back(Box(lo, hi));
    }
    const BoxArray ba(bl);
    const DistributionMapping dm(ba);
    velocity[l].define(ba, dm, NDIMS, 0);
    levels[l].Define(ba, dm);                    // per-box storage; Iterate re-lays level 0 out when it is alone
    levels[l].time = keep;
    stream_scratch[l].clear();
    valid_pending[l] = false;
    SetBoxArray(l, ba);
    SetDistributionMap(l, dm);
    MultiFab tight(ba, dm, NMODES, 0);
    const uint64_t n = c.get<uint64_t>();
    if (n != (uint64_t)ba.numPts() * NMODES) amrex::Abort("ReadCheckpoint: corrupt file (payload size does not match the box list)");
    std::vector<double> h((size_t)n);
    c.get(h.data(), h.size() * sizeof(double));
    tight.upload(h);
    amrex::CopyValid(levels[l].now.get<DistFn>(), tight);
    levels[l].now.get<DistFn>().touch();
  }
  finest_level = new_finest;
  for (int l = 0; l <= finest_level; ++l) CalcHydroVars(l);
  for (int l = 0; l < finest_level; ++l) MakeFineMask(l);
}
