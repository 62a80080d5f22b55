// Flat C mirror of AmrSim (include/lambrex_c.h).  Every entry point converts C++ failures
// (amrex::Abort -> AbortException, std::out_of_range, ...) into a status + message.
#include "../../include/lambrex_c.h"

#include <cstring>
#include <string>

#include "AmrSim.h"
#include "lambrex.h"

namespace {
thread_local std::string g_msg;

// the reference's tests reach protected state through a subclass (tests/AmrTest.h); same here
class SimAccess : public AmrSim {
 public:
  using AmrSim::AmrSim;
  using AmrSim::ClearLevel;
  using AmrSim::ErrorEst;
  using AmrSim::fine_masks;
  using AmrSim::levels;
  using AmrSim::MakeNewLevelFromCoarse;
  using AmrSim::MakeNewLevelFromScratch;
  using AmrSim::mass;
  using AmrSim::RemakeLevel;
  using AmrSim::tau_b;
  using AmrSim::tau_s;
  using AmrSim::velocity;
};

template <class F>
int guarded(F&& f) {
  try {
    amrex::SetAbortThrows(true);
    f();
    return 0;
  } catch (const std::exception& e) {
    g_msg = e.what();
    return 1;
  } catch (...) {
    g_msg = "unknown C++ exception";
    return 1;
  }
}
amrex::BoxArray boxes_from(const int* b, int n) {
  amrex::BoxList bl;
  for (int i = 0; i < n; ++i)
    bl.push_back(amrex::Box(amrex::IntVect(b[6 * i], b[6 * i + 1], b[6 * i + 2]),
                            amrex::IntVect(b[6 * i + 3], b[6 * i + 4], b[6 * i + 5])));
  return amrex::BoxArray(bl);
}
void boxes_to(const amrex::BoxArray& ba, int* out) {
  for (long i = 0; i < ba.size(); ++i)
    for (int d = 0; d < 3; ++d) { out[6 * i + d] = ba[i].smallEnd(d); out[6 * i + 3 + d] = ba[i].bigEnd(d); }
}
}  // namespace

struct lbx_sim {
  SimAccess s;
  lbx_sim(int nx, int ny, int nz, int ml, const std::array<int, 3>& per, double ts, double tb)
      : s(nx, ny, nz, ml, per, ts, tb) {}
};

namespace {
const amrex::FabArrayBase* field_of(const lbx_sim* sim, int level, int field) {
  switch (field) {
    case LBX_FIELD_DISTFN: return &sim->s.levels.at(level).now.get<DistFn>();
    case LBX_FIELD_DENSITY: return &sim->s.levels.at(level).now.get<Density>();
    case LBX_FIELD_VELOCITY: return &sim->s.velocity.at(level);
    case LBX_FIELD_DISTFN_NEXT: return &sim->s.levels.at(level).next.get<DistFn>();
    case LBX_FIELD_FINE_MASK: return &sim->s.fine_masks.at(level);
    default: amrex::Abort("unknown field id");
  }
}
}  // namespace

extern "C" {

const char* lbx_sim_last_error(void) { return g_msg.c_str(); }
int lbx_sim_global_init(void) { return guarded([] { lambrexInit(); }); }
int lbx_sim_global_finalise(void) { return guarded([] { lambrexFinalise(); }); }
int lbx_sim_global_init_parallel(int rank, int nranks, int (*allgather)(const void*, size_t, void*, void*), void* user) {
  return guarded([&] { lambrexInitParallel(rank, nranks, allgather, user); });
}
int lbx_sim_set_parallel_view(int rank, int nranks) { return guarded([&] { lambrexSetParallelView(rank, nranks); }); }
int lbx_sim_owner(const lbx_sim* sim, int level, int box, int* rank) {
  return guarded([&] { *rank = sim->s.DistributionMap(level)[box]; });
}

int lbx_sim_create(int nx, int ny, int nz, int max_level, const int per[3], double tau_s, double tau_b, lbx_sim** out) {
  return guarded([&] { *out = new lbx_sim(nx, ny, nz, max_level, {{per[0], per[1], per[2]}}, tau_s, tau_b); });
}
int lbx_sim_destroy(lbx_sim* sim) { return guarded([&] { delete sim; }); }
int lbx_sim_set_max_grid_size(lbx_sim* sim, int n) { return guarded([&] { sim->s.SetMaxGridSize(n); }); }
int lbx_sim_set_uniform_fast_path(lbx_sim* sim, int on) { return guarded([&] { sim->s.SetUniformFastPath(on != 0); }); }

int lbx_sim_set_rohde_fusion(lbx_sim* sim, int on) { return guarded([&] { sim->s.SetRohdeFusion(on != 0); }); }

int lbx_sim_set_gradient_refinement(lbx_sim* sim, int level, double threshold) {
  return guarded([&] { sim->s.SetGradientRefinement(level, threshold); });
}
int lbx_sim_unset_gradient_refinement(lbx_sim* sim, int level) { return guarded([&] { sim->s.UnsetGradientRefinement(level); }); }
int lbx_sim_set_regrid_interval(lbx_sim* sim, int n) { return guarded([&] { sim->s.SetRegridInterval(n); }); }
int lbx_sim_num_regrids(const lbx_sim* sim) { return sim->s.NumRegrids(); }
int lbx_sim_plan_cache_size(void) { return (int)amrex::PlanCacheSize(); }

int lbx_sim_get_linear_moment_field(const lbx_sim* sim, int level, const double* weights, int ncomp, int per_unit_density,
                                    double sentinel, double* out, size_t n) {
  return guarded([&] { sim->s.GetLinearMomentField(level, weights, ncomp, per_unit_density != 0, sentinel, out, n); });
}

int lbx_sim_write_plotfile(lbx_sim* sim, const char* dir) { return guarded([&] { sim->s.WritePlotFile(dir); }); }
int lbx_sim_write_checkpoint(lbx_sim* sim, const char* path) { return guarded([&] { sim->s.WriteCheckpoint(path); }); }
int lbx_sim_read_checkpoint(lbx_sim* sim, const char* path) { return guarded([&] { sim->s.ReadCheckpoint(path); }); }

int lbx_sim_set_coupling(lbx_sim* sim, int coupling) {
  return guarded([&] {
    if (coupling != LBX_COUPLING_ROHDE && coupling != LBX_COUPLING_SUBCYCLE) amrex::Abort("unknown coupling");
    sim->s.SetCoupling(coupling == LBX_COUPLING_SUBCYCLE ? AmrSim::Coupling::SUBCYCLE : AmrSim::Coupling::ROHDE);
  });
}

int lbx_sim_set_initial_density(lbx_sim* sim, const double* rho, size_t n) {
  return guarded([&] {
    if (n == 1) sim->s.SetInitialDensity(rho[0]);
    else sim->s.SetInitialDensity(std::vector<double>(rho, rho + n));     // the one copy the by-value API implies
  });
}
int lbx_sim_set_initial_velocity(lbx_sim* sim, const double* u, size_t n) {
  return guarded([&] {
    if (n == 1) sim->s.SetInitialVelocity(u[0]);
    else sim->s.SetInitialVelocity(std::vector<double>(u, u + n));
  });
}
int lbx_sim_set_initial_density_view(lbx_sim* sim, const double* rho, size_t n) {
  return guarded([&] { sim->s.SetInitialDensityView(rho, n); });
}
int lbx_sim_set_initial_velocity_view(lbx_sim* sim, const double* u, size_t n) {
  return guarded([&] { sim->s.SetInitialVelocityView(u, n); });
}
int lbx_sim_set_initial_density_profile(lbx_sim* sim, int axis, const double* v, size_t n) {
  return guarded([&] { sim->s.SetInitialDensityProfile(axis, std::vector<double>(v, v + n)); });
}
int lbx_sim_set_initial_velocity_profile(lbx_sim* sim, int axis, const double* v, size_t n) {
  return guarded([&] { sim->s.SetInitialVelocityProfile(axis, std::vector<double>(v, v + n)); });
}
int lbx_sim_local_box(lbx_sim* sim, int lo[3], int hi[3]) {
  return guarded([&] {
    const amrex::Box b = sim->s.LocalBox();
    for (int d = 0; d < 3; ++d) { lo[d] = b.smallEnd(d); hi[d] = b.bigEnd(d); }
  });
}
int lbx_sim_set_initial_density_local_view(lbx_sim* sim, const double* rho, size_t n) {
  return guarded([&] { sim->s.SetInitialDensityLocalView(rho, n); });
}
int lbx_sim_set_initial_velocity_local_view(lbx_sim* sim, const double* u, size_t n) {
  return guarded([&] { sim->s.SetInitialVelocityLocalView(u, n); });
}
int lbx_sim_get_local_density_field(const lbx_sim* sim, int level, double* out, size_t n) {
  return guarded([&] { sim->s.GetLocalDensityField(level, out, n); });
}
int lbx_sim_get_local_velocity_field(const lbx_sim* sim, int level, double* out, size_t n) {
  return guarded([&] { sim->s.GetLocalVelocityField(level, out, n); });
}
int lbx_sim_init_from_scratch(lbx_sim* sim, double time) { return guarded([&] { sim->s.InitFromScratch(time); }); }
int lbx_sim_regrid(lbx_sim* sim, int lbase, double time) { return guarded([&] { sim->s.regrid(lbase, time); }); }
int lbx_sim_iterate(lbx_sim* sim, int nsteps) { return guarded([&] { sim->s.Iterate(nsteps); }); }
int lbx_sim_calc_hydro_vars(lbx_sim* sim, int level) { return guarded([&] { sim->s.CalcHydroVars(level); }); }
int lbx_sim_calc_equilibrium_dist(lbx_sim* sim, int level) { return guarded([&] { sim->s.CalcEquilibriumDist(level); }); }

int lbx_sim_get_density(const lbx_sim* sim, int i, int j, int k, int level, double* out) {
  return guarded([&] { *out = sim->s.GetDensity(i, j, k, level); });
}
int lbx_sim_get_velocity(const lbx_sim* sim, int i, int j, int k, int n, int level, double* out) {
  return guarded([&] { *out = sim->s.GetVelocity(i, j, k, n, level); });
}
int lbx_sim_get_density_field(const lbx_sim* sim, int level, double* out, size_t n) {
  return guarded([&] { sim->s.GetDensityField(level, out, n); });
}
int lbx_sim_get_velocity_field(const lbx_sim* sim, int level, double* out, size_t n) {
  return guarded([&] { sim->s.GetVelocityField(level, out, n); });
}
int lbx_sim_get_time(const lbx_sim* sim, int level, double* out) { return guarded([&] { *out = sim->s.GetTime(level); }); }
int lbx_sim_get_time_step(const lbx_sim* sim, int level, int* out) {
  return guarded([&] { *out = sim->s.GetTimeStep(level); });
}
int lbx_sim_get_dims(const lbx_sim* sim, int dims[3]) {
  return guarded([&] { const auto d = sim->s.GetDims(); for (int a = 0; a < 3; ++a) dims[a] = d[a]; });
}
int lbx_sim_get_extent(const lbx_sim* sim, int level, int lo[3], int hi[3]) {
  return guarded([&] {
    const auto e = sim->s.GetExtent(level);
    for (int a = 0; a < 3; ++a) { lo[a] = e.first[a]; hi[a] = e.second[a]; }
  });
}
int lbx_sim_set_static_refinement(lbx_sim* sim, int level, const int lo[3], const int hi[3]) {
  return guarded([&] { sim->s.SetStaticRefinement(level, {{lo[0], lo[1], lo[2]}}, {{hi[0], hi[1], hi[2]}}); });
}
int lbx_sim_set_static_box(lbx_sim* sim, int level, const int lo[3], const int hi[3]) {
  return guarded([&] { sim->s.SetStaticBox(level, {{lo[0], lo[1], lo[2]}}, {{hi[0], hi[1], hi[2]}}); });
}
int lbx_sim_regrid_all(lbx_sim* sim) { return guarded([&] { sim->s.Regrid(); }); }
int lbx_sim_allow_walls(int on) { AmrSim::AllowWalls(on != 0); return 0; }
int lbx_sim_unset_static_refinement(lbx_sim* sim, int level) {
  return guarded([&] { sim->s.UnsetStaticRefinement(level); });
}
int lbx_sim_max_level(const lbx_sim* sim) { return sim->s.maxLevel(); }
int lbx_sim_finest_level(const lbx_sim* sim) { return sim->s.finestLevel(); }
int lbx_sim_ref_ratio(const lbx_sim* sim, int level, int ratio[3]) {
  return guarded([&] { const auto r = sim->s.refRatio().at(level); for (int a = 0; a < 3; ++a) ratio[a] = r[a]; });
}
int lbx_sim_num_boxes(const lbx_sim* sim, int level) {
  if (level < 0 || level > sim->s.maxLevel()) return -1;
  return (int)sim->s.boxArray(level).size();
}
int lbx_sim_get_boxes(const lbx_sim* sim, int level, int* boxes) {
  return guarded([&] { boxes_to(sim->s.boxArray().at(level), boxes); });
}

int lbx_sim_field_empty(const lbx_sim* sim, int level, int field) {
  int r = -1;
  if (guarded([&] { r = field_of(sim, level, field)->empty() ? 1 : 0; })) return -1;
  return r;
}
int lbx_sim_field_num_boxes(const lbx_sim* sim, int level, int field) {
  int r = -1;
  if (guarded([&] { r = (int)field_of(sim, level, field)->size(); })) return -1;
  return r;
}
int lbx_sim_field_boxes(const lbx_sim* sim, int level, int field, int* boxes) {
  return guarded([&] { boxes_to(field_of(sim, level, field)->boxArray(), boxes); });
}
int lbx_sim_field_fab(const lbx_sim* sim, int level, int field, int b, double* out, size_t n, int shape[4]) {
  return guarded([&] {
    const amrex::FabArrayBase* fa = field_of(sim, level, field);
    if (b < 0 || b >= fa->size()) amrex::Abort("lbx_sim_field_fab: box index out of range");
    const amrex::Box g = fa->fabbox(b);
    const int nc = fa->nComp();
    shape[0] = nc; shape[1] = g.length(2); shape[2] = g.length(1); shape[3] = g.length(0);
    const size_t cells = (size_t)g.numPts();
    if (n < cells * nc) amrex::Abort("lbx_sim_field_fab: buffer too small");
    size_t q = 0;
    for (int c = 0; c < nc; ++c)
      for (int k = g.smallEnd(2); k <= g.bigEnd(2); ++k)
        for (int j = g.smallEnd(1); j <= g.bigEnd(1); ++j)
          for (int i = g.smallEnd(0); i <= g.bigEnd(0); ++i, ++q) {
            const amrex::IntVect p(i, j, k);
            if (field == LBX_FIELD_FINE_MASK) {
              out[q] = (double)static_cast<const amrex::iMultiFab*>(fa)->hostValue(b, p, c);
            } else {
              const auto* mf = static_cast<const amrex::MultiFab*>(fa);
              // FLAT storage keeps no ghost cells: report them as 0 like a freshly allocated fab
              out[q] = (mf->isFlat() && !mf->boxArray().minimalBox().contains(p)) ? 0.0 : mf->hostValue(b, p, c);
            }
          }
  });
}
int lbx_sim_get_tau(const lbx_sim* sim, int level, double* ts, double* tb) {
  return guarded([&] { *ts = sim->s.tau_s.at(level); *tb = sim->s.tau_b.at(level); });
}
int lbx_sim_get_mass(const lbx_sim* sim, int level, double* m) { return guarded([&] { *m = sim->s.mass.at(level); }); }
int lbx_sim_get_dt(const lbx_sim* sim, int level, double* dt) {
  return guarded([&] { *dt = sim->s.levels.at(level).time.delta; });
}
int lbx_sim_num_levels_allocated(const lbx_sim* sim) {
  const auto& s = sim->s;
  const size_t n = s.levels.size();
  if (s.velocity.size() != n || s.tau_s.size() != n || s.tau_b.size() != n || s.mass.size() != n) return -1;
  return (int)n;
}

int lbx_sim_call_error_est(lbx_sim* sim, int level, const int* tag_boxes, int ntag, int preset, char* out, size_t n) {
  return guarded([&] {
    const amrex::BoxArray ba = boxes_from(tag_boxes, ntag);
    amrex::TagBoxArray tba(ba, amrex::DistributionMapping(ba));
    tba.setVal(sim->s.boxArray(level), (amrex::TagBox::TagVal)preset);
    sim->s.ErrorEst(level, tba, sim->s.GetTime(level), 1);
    size_t q = 0;
    for (amrex::MFIter mfi(tba); mfi.isValid(); ++mfi) {
      const amrex::Box b = mfi.validbox();
      const amrex::TagBox& t = static_cast<const amrex::TagBoxArray&>(tba)[mfi];      // const read: no storage is forced
      for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k)
        for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
          for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i) {
            if (q >= n) amrex::Abort("lbx_sim_call_error_est: buffer too small");
            out[q++] = t(amrex::IntVect(i, j, k));
          }
    }
  });
}
int lbx_sim_call_make_new_level_from_scratch(lbx_sim* sim, int level, const int* boxes, int nboxes, double time) {
  return guarded([&] {
    const amrex::BoxArray ba = boxes_from(boxes, nboxes);
    sim->s.MakeNewLevelFromScratch(level, time, ba, amrex::DistributionMapping(ba));
  });
}
int lbx_sim_call_make_new_level_from_coarse(lbx_sim* sim, int level, const int* boxes, int nboxes) {
  return guarded([&] {
    const amrex::BoxArray ba = boxes_from(boxes, nboxes);
    sim->s.MakeNewLevelFromCoarse(level, sim->s.GetTime(level - 1), ba, amrex::DistributionMapping(ba));
  });
}
int lbx_sim_call_remake_level(lbx_sim* sim, int level, double time, const int* boxes, int nboxes) {
  return guarded([&] {
    const amrex::BoxArray ba = boxes_from(boxes, nboxes);
    sim->s.RemakeLevel(level, time, ba, amrex::DistributionMapping(ba));
  });
}
int lbx_sim_call_clear_level(lbx_sim* sim, int level) { return guarded([&] { sim->s.ClearLevel(level); }); }

// ----------------------------------------------------------------------------- metadata only
namespace {
int emit(const amrex::BoxList& bl, int* out, int cap) {
  if ((int)bl.size() > cap) { g_msg = "box buffer too small"; return -1; }
  boxes_to(amrex::BoxArray(bl), out);
  return (int)bl.size();
}
// field-less AmrCore with AmrSim's tagging rule
class MetaMesh : public amrex::AmrCore {
 public:
  MetaMesh(const int dims[3], int max_level, int max_grid)
      : AmrCore(amrex::Geometry(amrex::Box(amrex::IntVect(0), amrex::IntVect(dims[0] - 1, dims[1] - 1, dims[2] - 1)),
                                amrex::RealBox(), 0, {{1, 1, 1}}),
                info(max_level, max_grid)),
        static_tags((size_t)max_level + 1), times((size_t)max_level + 1, 0.0) {}
  static amrex::AmrInfo info(int max_level, int max_grid) {
    amrex::AmrInfo i;
    i.max_level = max_level;
    i.ref_ratio.assign((size_t)max_level + 1, amrex::IntVect(2));
    i.blocking_factor.assign((size_t)max_level + 1, amrex::IntVect(1));
    i.max_grid_size.assign((size_t)max_level + 1, amrex::IntVect(max_grid));
    return i;
  }
  void ErrorEst(int lev, amrex::TagBoxArray& tags, double, int) override {
    for (amrex::MFIter mfi(tags); mfi.isValid(); ++mfi) {
      tags[mfi].setVal(amrex::TagBox::CLEAR, mfi.validbox());
      for (const auto& hit : static_tags.at(lev).intersections(mfi.validbox()))
        tags[mfi].setVal(amrex::TagBox::SET, hit.second);
    }
  }
  void MakeNewLevelFromScratch(int l, double, const amrex::BoxArray&, const amrex::DistributionMapping&) override { log += "S" + std::to_string(l) + " "; }
  void MakeNewLevelFromCoarse(int l, double, const amrex::BoxArray&, const amrex::DistributionMapping&) override { log += "C" + std::to_string(l) + " "; }
  void RemakeLevel(int l, double, const amrex::BoxArray&, const amrex::DistributionMapping&) override { log += "R" + std::to_string(l) + " "; }
  void ClearLevel(int l) override { log += "X" + std::to_string(l) + " "; }
  std::vector<amrex::BoxArray> static_tags;
  std::vector<double> times;
  std::string log;
};
}  // namespace
struct lbx_meta_mesh { MetaMesh m; lbx_meta_mesh(const int d[3], int ml, int mg) : m(d, ml, mg) {} };

int lbx_meta_base_grids(const int dims[3], int max_grid_size, int* boxes, int cap) {
  int n = -1;
  if (guarded([&] {
        MetaMesh m(dims, 0, max_grid_size);
        n = emit(m.MakeBaseGrids().boxList(), boxes, cap);
      })) return -1;
  return n;
}
int lbx_meta_max_size(const int* in, int n, int chunk, int* boxes, int cap) {
  int r = -1;
  if (guarded([&] { amrex::BoxList bl = boxes_from(in, n).boxList(); amrex::maxSize(bl, amrex::IntVect(chunk)); r = emit(bl, boxes, cap); })) return -1;
  return r;
}
int lbx_meta_simplify(const int* in, int n, int* boxes, int cap) {
  int r = -1;
  if (guarded([&] { amrex::BoxList bl = boxes_from(in, n).boxList(); amrex::simplify(bl); r = emit(bl, boxes, cap); })) return -1;
  return r;
}
int lbx_meta_complement(const int region[6], const int* in, int n, int* boxes, int cap) {
  int r = -1;
  if (guarded([&] {
        const amrex::Box reg(amrex::IntVect(region[0], region[1], region[2]), amrex::IntVect(region[3], region[4], region[5]));
        r = emit(amrex::complementIn(reg, boxes_from(in, n).boxList()), boxes, cap);
      })) return -1;
  return r;
}
int lbx_meta_cluster(const int* points, int npoints, double eff, int* boxes, int cap) {
  int r = -1;
  if (guarded([&] {
        std::vector<amrex::IntVect> t(npoints);
        for (int i = 0; i < npoints; ++i) t[i] = amrex::IntVect(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        r = emit(amrex::ClusterTags(t, eff), boxes, cap);
      })) return -1;
  return r;
}
int lbx_meta_parallel_init(int rank, int nranks, int (*allgather)(const void*, size_t, void*, void*), void* user) {
  return guarded([&] {
    amrex::lbx_check(lbx_par_init_host(rank, nranks, allgather, user), "lbx_meta_parallel_init");
    amrex::DistributionMapping::SetParallel(rank, nranks);
  });
}
int lbx_meta_parallel_finalise(void) {
  return guarded([&] {
    amrex::lbx_check(lbx_par_finalize_host(), "lbx_meta_parallel_finalise");
    amrex::DistributionMapping::SetParallel(0, 1);
  });
}
int lbx_meta_distribution(const int* in_boxes, int n, int nprocs, int* owners) {
  return lbx_meta_distribution_runs(in_boxes, n, nprocs, 1, owners);
}
int lbx_meta_distribution_runs(const int* in_boxes, int n, int nprocs, int runs_per_rank, int* owners) {
  return guarded([&] {
    amrex::BoxList bl;
    for (int i = 0; i < n; ++i)
      bl.push_back(amrex::Box(amrex::IntVect(in_boxes[6 * i], in_boxes[6 * i + 1], in_boxes[6 * i + 2]),
                              amrex::IntVect(in_boxes[6 * i + 3], in_boxes[6 * i + 4], in_boxes[6 * i + 5])));
    const amrex::DistributionMapping dm(amrex::BoxArray(bl), nprocs, runs_per_rank);
    for (int i = 0; i < n; ++i) owners[i] = dm[i];
  });
}
int lbx_meta_mesh_create(const int dims[3], int max_level, int max_grid_size, lbx_meta_mesh** out) {
  return guarded([&] { *out = new lbx_meta_mesh(dims, max_level, max_grid_size); (*out)->m.InitFromScratch(0.0); });
}
int lbx_meta_mesh_destroy(lbx_meta_mesh* m) { delete m; return 0; }
int lbx_meta_mesh_set_static(lbx_meta_mesh* m, int level, const int lo[3], const int hi[3]) {
  return guarded([&] {
    m->m.static_tags.at(level).define(amrex::Box(amrex::IntVect(lo[0], lo[1], lo[2]), amrex::IntVect(hi[0], hi[1], hi[2])));
    m->m.regrid(level, 0.0);
  });
}
int lbx_meta_mesh_unset_static(lbx_meta_mesh* m, int level) {
  return guarded([&] { m->m.static_tags.at(level).clear(); m->m.regrid(level, 0.0); });
}
int lbx_meta_mesh_finest_level(const lbx_meta_mesh* m) { return m->m.finestLevel(); }
int lbx_meta_mesh_boxes(const lbx_meta_mesh* m, int level, int* boxes, int cap) {
  int r = -1;
  if (guarded([&] { r = emit(m->m.boxArray().at(level).boxList(), boxes, cap); })) return -1;
  return r;
}
int lbx_meta_mesh_log(const lbx_meta_mesh* m, char* out, int cap) {
  if ((int)m->m.log.size() + 1 > cap) return -1;
  std::memcpy(out, m->m.log.c_str(), m->m.log.size() + 1);
  return (int)m->m.log.size();
}

}  // extern "C"
