// -*- mode: c++ -*-
// AmrSim: the reference's simulation class (/root/reference/include/AmrSim.h:23-155) with the
// same public and protected member names, argument meaning and error behaviour, implemented
// as host control flow over device kernels (include/lbx.h).  Each member cites the reference
// lines it replaces in AmrSim.cpp.
//
// What differs by design:
//  * field data lives in HBM; members launch kernels, the getters read a lazily refreshed host
//    mirror (GetDensity/GetVelocity keep the reference's per-cell signature and sentinels).
//  * while level 0 is the only level it is stored FLAT (one ghost-free fab over the periodic
//    domain) and one Iterate step is ONE fused kernel; with finer levels present every level
//    uses per-box storage and the Rohde cycle runs as batched per-level kernels.
//  * Stream() reuses a third population buffer instead of allocating a fab per call; the
//    buffer's ghost ring 2 is zeroed by the kernel (a "fresh" fab).
#ifndef LBX_AMRSIM_H
#define LBX_AMRSIM_H

#include <array>
#include <string>
#include <utility>
#include <vector>

#include "AMReX_AmrCore.H"
#include "AMReX_FillPatch.H"
#include "d3q15_bgk.h"

#define NDIMS 3
#define NMODES 15

class AmrSim : public amrex::AmrCore {
 public:
  AmrSim(int const nx, int const ny, int const nz, int const max_ref_level,
         const std::array<int, NDIMS>& periodicity, double const tau_s_0, double const tau_b_0);
  ~AmrSim() override;

  // clocks and extents
  double GetTime(int const level) const { return levels.at(level).time.current; }
  int GetTimeStep(int const level) const { return levels.at(level).time.step; }
  std::array<int, NDIMS> GetDims() const { return {{NX, NY, NZ}}; }
  std::pair<std::array<int, NDIMS>, std::array<int, NDIMS>> GetExtent(int const level) const;

  // initial condition: scalars or C-ordered arrays rho[(i*NY+j)*NZ+k], u[((i*NY+j)*NZ+k)*3+n]
  void SetInitialDensity(double const rho_init);
  void SetInitialDensity(std::vector<double> rho_init);
  void SetInitialVelocity(double const u_init);
  void SetInitialVelocity(std::vector<double> u_init);

  // output; cells the level does not hold give the sentinels NL_DENSITY / NL_VELOCITY
  double GetDensity(int const i, int const j, int const k, int const level) const;
  double GetVelocity(int const i, int const j, int const k, int const n, int const level) const;
  bool OnProcessDensity(double const rho) const { return rho != NL_DENSITY; }
  bool OnProcessVelocity(double const u) const { return u != NL_VELOCITY; }

  void CalcEquilibriumDist(int const level);
  void CalcHydroVars(int const level);
  void Iterate(int const nsteps);
  void SetStaticRefinement(int const level, const std::array<int, NDIMS>& lo_corner,
                           const std::array<int, NDIMS>& hi_corner);
  void UnsetStaticRefinement(int const level);

  // ---- additions (not in the reference) -------------------------------------------------
  // dense copies of the level's fields over its index domain, C-ordered like the inputs;
  // cells the level does not hold carry the sentinels.  One device->host copy per call.
  std::vector<double> GetDensityField(int const level) const;
  std::vector<double> GetVelocityField(int const level) const;
  void GetDensityField(int const level, double* out, size_t n) const;    // into caller memory
  void GetVelocityField(int const level, double* out, size_t n) const;
  // generic derived variables (include/derived_var.h LinearMoment): DV::fill on the
  // level's NOW populations, `out` (re)defined on the level's boxes when needed ...
  template <class DV>
  void CalcDerived(int const level, amrex::MultiFab& out) {
    const amrex::MultiFab& f = levels.at(level).now.get<DistFn>();
    if (out.empty() || out.boxArray() != f.boxArray() || out.layout() != f.layout() || out.nComp() != (int)DV::NELEM)
      out.define(f.boxArray(), f.DistributionMap(), (int)DV::NELEM, 0, f.layout());
    DV::fill(out, f);
  }
  // ... and its run-time form: `ncomp` (<= 10) weight rows over the 15 populations, result dense over
  // the level's domain, C-ordered [i][j][k][c]; cells the level does not hold carry `sentinel`
  void GetLinearMomentField(int const level, const double* weights, int const ncomp, bool const per_unit_density,
                            double const sentinel, double* out, size_t n) const;
  // checkpoint / restart (the reference has no I/O).  The file holds the clocks, the tau
  // ladder, the refinement criteria, every level's box list and the VALID-cell populations of NOW;
  // ghost cells, densities, velocities and masks are recomputed.  ReadCheckpoint needs a sim constructed
  // with the same extents and max level (single process); a restarted run continues bit for bit.
  // plotfile (the reference has no I/O): an AMReX-format plotfile directory (HyperCLaw-V1.1 Header,
  // Level_l/Cell_H + Cell_D_00000 with native fp64 FABs) holding rho, ux, uy, uz of every level on its boxes -- what
  // amrex::WriteMultiLevelPlotfile would write for these fields; readable by yt / VisIt / Amrvis.  Single process.
  void WritePlotFile(const std::string& dir);
  void WriteCheckpoint(const std::string& path);
  void ReadCheckpoint(const std::string& path);
  // zero-copy inputs: the arrays (same C ordering) are read at InitFromScratch directly from
  // caller memory -- pinned memory makes that one DMA -- and must stay valid until it returns.
  void SetInitialDensityView(const double* rho_init, size_t n) { density_view = rho_init; density_view_n = n; }
  void SetInitialVelocityView(const double* u_init, size_t n) { velocity_view = u_init; velocity_view_n = n; }
  // ---- distributed runs (lambrexInitParallel): large domains cannot be initialised from, or returned as,
  // whole-domain host arrays (1024^3: 8.6 + 25.8 GB per rank), so each rank states and reads ITS part.
  // Separable initial conditions: the field varies along one axis only (planar pulse: rho(z); shear wave:
  // u(y)); profile[a] resp. profile[a * 3 + n] for a = 0 .. extent(axis) - 1.  Replace SetInitial{Density,Velocity}.
  void SetInitialDensityProfile(int const axis, std::vector<double> rho_of_axis);
  void SetInitialVelocityProfile(int const axis, std::vector<double> u_of_axis);
  // This rank's part of level 0 while it is stored as one slab per rank (uniform path): the box a rank
  // will own, available right after construction (grid generation is deterministic) ...
  amrex::Box LocalBox();
  // ... initial arrays over that box, C-ordered [i][j][k]([n]) like the whole-domain ones (read at
  // InitFromScratch; must stay valid until it returns; pinned memory makes it one DMA) ...
  void SetInitialDensityLocalView(const double* rho_init, size_t n) { density_view = rho_init; density_view_n = n; views_local = true; }
  void SetInitialVelocityLocalView(const double* u_init, size_t n) { velocity_view = u_init; velocity_view_n = n; views_local = true; }
  // ... and the matching bulk getters (this rank's cells only; no communication)
  void GetLocalDensityField(int const level, double* out, size_t n) const;
  void GetLocalVelocityField(int const level, double* out, size_t n) const;
  // Addition: solid no-slip walls.  The reference's constructor aborts on a non-periodic direction
  // (src/AmrSim.cpp:788-797; DistFnFillShim :346-357 is its intended hook).  After AllowWalls(true) -- process-wide,
  // before construction -- a direction with periodicity 0 is closed by walls at both domain faces (half-way
  // bounce-back inside the fused step kernel: a population that would leave through a wall returns to its cell
  // reversed).  Walls run on the uniform single-GPU path only (level 0 alone, one fab); refinement or a distributed
  // run with walls aborts.  Default off: non-periodic input aborts like the reference.
  static void AllowWalls(bool on);
  static bool WallsAllowed();
  // false: run level 0 through the reference's literal pass structure on per-box storage even
  // when it is the only level (FillPatch, collide, FillBoundary, stream, swap).  Default true.
  void SetUniformFastPath(bool on) { uniform_fast_path = on; }
  // false: run the Rohde cycle as the reference's literal sequence of passes (collide, Stream and
  // ZeroInvalidComponents as separate launches).  Default true: one fused pass per collide+Stream.
  void SetRohdeFusion(bool on) { rohde_fused = on; }
  // How Iterate couples the levels of a refined hierarchy.
  //   ROHDE    (default): the reference's live path, RohdeCycle (src/AmrSim.cpp:430-469), bug for bug.
  //   SUBCYCLE : the conventional AMR driver the reference sketches in its dead SubCycle (:335-344):
  //              per level FillPatch (coarse ghost data interpolated in TIME between the coarse
  //              level's old and new states, piecewise constant in space) + collide + FillBoundary +
  //              Stream, `ratio` fine steps per coarse step, then average_down of the populations.
  // Dynamic refinement (the reference tags one static box per level, TagCell :413-417).
  // SetGradientRefinement: ErrorEst also tags the valid cells of `level` where the central-difference
  // density gradient |grad rho| exceeds `threshold` (evaluated on the device); like
  // SetStaticRefinement it regrids at once.  SetRegridInterval(n > 0): Iterate calls regrid(0, t)
  // after every n-th coarse step (AMReX's regrid_int); 0 = only on request (the reference).
  void SetGradientRefinement(int const level, double const threshold);
  void UnsetGradientRefinement(int const level);
  void SetRegridInterval(int const n) { regrid_int = n; }
  // Static boxes of several levels changed together: SetStaticRefinement regrids at once, level by level (the
  // reference, src/AmrSim.cpp:995-1009), so moving the boxes of an L-level hierarchy costs L regrids.
  // SetStaticBox records the box of `level` WITHOUT regridding; Regrid() then regrids every level from 0 once
  // (AmrCore::regrid(0, t)) and rebuilds the fine masks.
  void SetStaticBox(int const level, const std::array<int, NDIMS>& lo_corner, const std::array<int, NDIMS>& hi_corner);
  void Regrid();
  int NumRegrids() const { return num_regrids; }
  enum class Coupling { ROHDE = 0, SUBCYCLE = 1 };
  void SetCoupling(Coupling c) { coupling = c; }
  Coupling GetCoupling() const { return coupling; }

 protected:
  const int NX, NY, NZ, NUMEL, COORD_SYS;
  std::array<int, NDIMS> PERIODICITY;

  constexpr static double CS2 = 1.0 / 3.0;
  constexpr static double NL_DENSITY = -1.0;
  constexpr static double NL_VELOCITY = -3E8;
  // include/AmrSim.h:37-39.  The kernels carry their own constexpr copy (csrc/d3q15.cuh, generated
  // from exact rationals); these host-side members hold the same values (lbx_d3q15_tables) for
  // subclasses that read them.  DELTA is diag(1 / NMODES), sic (src/AmrSim.cpp:1033-1035).
  static const double DELTA[NDIMS][NDIMS];
  static const double MODE_MATRIX[NMODES][NMODES];
  static const double MODE_MATRIX_INVERSE[NMODES][NMODES];
  std::vector<double> initial_density;
  std::vector<double> initial_velocity;
  std::vector<amrex::BoxArray> static_tags;

  std::vector<double> tau_s;
  std::vector<double> tau_b;
  std::vector<double> mass;
  std::vector<amrex::MultiFab> velocity;    // output, 3 comps, no ghosts
  std::vector<SimLevelData> levels;         // now/next x {DistFn, Density} + clock

  const int COARSE_VAL = 0;
  const int FINE_VAL = 1;
  std::vector<amrex::iMultiFab> fine_masks;

  int CLindex(int const i, int const j, int const k, int const n, amrex::IntVect const dims,
              int const n_comps) const {
    return ((i * dims[1] + j) * dims[2] + k) * n_comps + n;
  }

  // single-level step
  void UpdateBoundaries(int const level);
  void Collide(amrex::MultiFab& f, const double omega_s, const double omega_b);
  void Stream(int const level);
  void CollideLevel(int const level);
  void CollideAndStream(int const level);
  void IterateLevel(int const level);
  void SubCycle(int const base_level, int const num_steps);
  // conventional subcycling (Coupling::SUBCYCLE): one step of `level`, then ratio x the finer
  // levels, then AverageDown(level)
  void SubCycleAdvance(int const level);
  void AverageDown(int const coarse_level);

  // set-up
  void InitDensity(int const level);
  void InitVelocity(int const level);
  void ComputeDt(int const level);
  void DistFnFillPatch(int const level, amrex::MultiFab& dest);
  void DistFnFillFromCoarse(int const level, amrex::MultiFab& fine_mf);
  bool TagCell(int const level, const amrex::IntVect& pos);
  void MakeFineMask(int const coarse_level);

  // Rohde et al. (2006) volumetric two-level coupling
  void RohdeCycle(int const coarse_level);
  void InitPostCollision(int const level);
  void CoarseCollide(int const level);
  void FineCollide(int const level);
  void SumFromFine(int const coarse_level);
  void ZeroInvalidComponents(int const level);
  void UpdateDistribution(int const level);

  // amrex::AmrCore hooks
  void ErrorEst(int level, amrex::TagBoxArray& tags, double time, int ngrow) override;
  void MakeNewLevelFromScratch(int level, double time, const amrex::BoxArray& ba,
                               const amrex::DistributionMapping& dm) override;
  void MakeNewLevelFromCoarse(int level, double time, const amrex::BoxArray& ba,
                              const amrex::DistributionMapping& dm) override;
  void RemakeLevel(int level, double time, const amrex::BoxArray& ba, const amrex::DistributionMapping& dm) override;
  void ClearLevel(int level) override;

  // storage layout of a level (see AMReX_MultiFab.H)
  void SetLevelLayout(int const level, amrex::Layout lay);
  amrex::Layout PreferredLayout(int const level) const;
  amrex::Layout PreferredLayout(int const level, const amrex::BoxArray& ba, const amrex::DistributionMapping& dm) const;

 private:
  std::vector<amrex::MultiFab> stream_scratch;   // third population buffer per level
  // InitPostCollision filled only the ghost cells of NEXT; its valid cells are still NOW's and
  // the next collision reads them from there (valid-cell copy fused into the collision)
  std::vector<char> valid_pending;
  void FillPatchImpl(int const level, amrex::MultiFab& dest, bool ghosts_only, const amrex::GhostPush* push = nullptr);
  bool has_walls = false;
  bool uniform_fast_path = true;
  bool rohde_fused = true;
  bool CanFuseRohde(int const level) const;
  void RohdeCycleFused(int const coarse_level);
  void CollideStreamFused(int const level, bool masked, bool zero_invalid, bool from_fillpatch);
  // CollideLevel + Stream of a level on per-box storage as ONE launch (lbx_mf_collide_stream_level)
  bool CanFuseLevelStep(int const level) const;
  void LevelStepFused(int const level);
  bool defer_boundaries = false;
  Coupling coupling = Coupling::ROHDE;
  std::vector<double> gradient_threshold;        // per level; <= 0: criterion off
  int regrid_int = 0, steps_since_regrid = 0, num_regrids = 0;
  void GradientTags(int const level, amrex::TagBoxArray& tags);
  void RegridIfDue();
  // the coarse populations FillPatchTwoLevels reads when filling `fine_level` at time t: NOW in the
  // reference's coupling; under SUBCYCLE the coarse state at t (old, new, or their LinComb)
  const amrex::MultiFab& CoarseStateAt(int const coarse_level, double const t);
  // the same choice without materialising the LinComb: state a with weight wa (+ state b with wb)
  void CoarseStatesAt(int const coarse_level, double const t, const amrex::MultiFab*& a, double& wa,
                      const amrex::MultiFab*& b, double& wb);
  std::vector<amrex::MultiFab> coarse_interp;    // LinComb scratch per coarse level (SUBCYCLE)
  void upload_user_field(amrex::MultiFab& mf, const double* user, size_t n, int ncomp, bool local);
  void upload_profile(amrex::MultiFab& mf, const std::vector<double>& profile, int axis, int ncomp);
  std::vector<double> density_profile, velocity_profile;
  int density_profile_axis = -1, velocity_profile_axis = -1;
  bool views_local = false;
  // distributed uniform path: the neighbours' face stores of the last queued step are not yet known to be
  // complete; FinishPeerStores queues the wait (lbx_par_step_finish) before anything else reads the populations
  bool peer_stores_pending = false;
  void FinishPeerStores();
  const double* density_view = nullptr;
  const double* velocity_view = nullptr;
  size_t density_view_n = 0, velocity_view_n = 0;
  void dense_field_into(const amrex::MultiFab& mf, int level, double sentinel, double* out, size_t n, bool local = false) const;
};

#endif
