/*
 * lambrex_c.h -- flat C mirror of the reference's C++ surface (liblambrex.so), for bindings
 * from other languages (ctypes / cgo / JNI).  One function per public member of
 * /root/reference/include/AmrSim.h:127-155 and include/lambrex.h:6-7, plus the inherited
 * amrex::AmrCore members the reference's callers use and the protected members its own test
 * subclass re-exports (/root/reference/tests/AmrTest.h:6-50), so the reference's tests can be
 * restated against this library.
 *
 * Conventions: int functions return 0 on success; on failure lbx_sim_last_error() holds the
 * message (amrex::Abort, std::out_of_range and CUDA errors are all reported this way instead
 * of terminating the process).  Host buffers are owned by the caller.  Levels are 0-based.
 * Boxes are 6 ints: lo[3], hi[3], inclusive.
 */
#ifndef LAMBREX_C_H
#define LAMBREX_C_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lbx_sim lbx_sim;

/* lambrexInit / lambrexFinalise (include/lambrex.h:6-7) */
int lbx_sim_global_init(void);
int lbx_sim_global_finalise(void);
/* addition: distributed start-up (one process per GPU): lbx_init + lbx_par_init + box ownership by
 * rank.  allgather(send, bytes, recv, user) gathers `bytes` from every rank, rank order, returns 0. */
int lbx_sim_global_init_parallel(int rank, int nranks,
                                 int (*allgather)(const void *send, size_t bytes, void *recv, void *user), void *user);
int lbx_sim_set_parallel_view(int rank, int nranks);   /* testing aid: ownership view of this process */
int lbx_sim_owner(const lbx_sim *sim, int level, int box, int *rank);   /* DistributionMap(level)[box] */
const char *lbx_sim_last_error(void);

/* addition: allow non-periodic directions = solid no-slip walls (half-way bounce-back) on the
 * uniform single-GPU path; process-wide, call before lbx_sim_create.  Default 0: non-periodic input fails like the
 * reference's constructor aborts (src/AmrSim.cpp:788-797). */
int lbx_sim_allow_walls(int on);
/* AmrSim::AmrSim (include/AmrSim.h:128-129) */
int lbx_sim_create(int nx, int ny, int nz, int max_level, const int periodicity[3], double tau_s,
                   double tau_b, lbx_sim **out);
int lbx_sim_destroy(lbx_sim *sim);
/* inherited AmrMesh knobs, callable before InitFromScratch */
int lbx_sim_set_max_grid_size(lbx_sim *sim, int n);
int lbx_sim_set_uniform_fast_path(lbx_sim *sim, int on);
/* 0: Rohde cycle as the reference's literal pass sequence; 1 (default): collide+Stream fused per level */
int lbx_sim_set_rohde_fusion(lbx_sim *sim, int on);
/* addition: how Iterate couples refined levels.  ROHDE (default) = the reference's live
 * RohdeCycle (src/AmrSim.cpp:430-469); SUBCYCLE = the conventional driver its dead SubCycle (:335-344)
 * sketches: per-level FillPatch (time-interpolated coarse data) + collide + FillBoundary + Stream,
 * `ratio` fine steps per coarse step, then average_down of the populations. */
enum { LBX_COUPLING_ROHDE = 0, LBX_COUPLING_SUBCYCLE = 1 };
int lbx_sim_set_coupling(lbx_sim *sim, int coupling);
/* addition: dynamic refinement.  Gradient criterion: ErrorEst also tags valid cells of
 * `level` where the central-difference |grad rho| > threshold (device kernel); set/unset regrid at
 * once like Set/UnsetStaticRefinement.  Regrid interval n > 0: Iterate calls regrid(0, t) after every
 * n-th coarse step (0, the default and the reference's behaviour: only on request). */
int lbx_sim_set_gradient_refinement(lbx_sim *sim, int level, double threshold);
int lbx_sim_unset_gradient_refinement(lbx_sim *sim, int level);
int lbx_sim_set_regrid_interval(lbx_sim *sim, int n);
int lbx_sim_num_regrids(const lbx_sim *sim);
/* gather plans currently cached (process-wide): bounded over any number of regrids (generation sweep in
 * AmrCore::regrid + LRU cap); a diagnostic for long dynamic-AMR runs */
int lbx_sim_plan_cache_size(void);

/* SetInitialDensity / SetInitialVelocity (:141-144); n == 1 selects the scalar overloads */
int lbx_sim_set_initial_density(lbx_sim *sim, const double *rho, size_t n);
int lbx_sim_set_initial_velocity(lbx_sim *sim, const double *u, size_t n);
/* zero-copy variants (addition): the arrays are read by lbx_sim_init_from_scratch straight from
 * caller memory (one DMA when it is pinned, see lbx_host_alloc) and must stay valid until then */
int lbx_sim_set_initial_density_view(lbx_sim *sim, const double *rho, size_t n);
int lbx_sim_set_initial_velocity_view(lbx_sim *sim, const double *u, size_t n);
/* ---- additions for distributed runs (one process per GPU): each rank states and reads ITS part ----
 * separable initial conditions -- the field varies along one axis (0 x, 1 y, 2 z): n = extent(axis) values for
 * the density, 3 * extent(axis) (component fastest) for the velocity; replace set_initial_density / velocity */
int lbx_sim_set_initial_density_profile(lbx_sim *sim, int axis, const double *rho_of_axis, size_t n);
int lbx_sim_set_initial_velocity_profile(lbx_sim *sim, int axis, const double *u_of_axis, size_t n);
/* the box of level 0 this rank owns while the level is stored as one slab per rank (whole domain on one rank);
 * available right after lbx_sim_create */
int lbx_sim_local_box(lbx_sim *sim, int lo[3], int hi[3]);
/* zero-copy initial arrays over that box, C-ordered [i][j][k]([n]); read by lbx_sim_init_from_scratch */
int lbx_sim_set_initial_density_local_view(lbx_sim *sim, const double *rho, size_t n);
int lbx_sim_set_initial_velocity_local_view(lbx_sim *sim, const double *u, size_t n);
/* bulk output over that box (this rank's cells, no communication), C-ordered [i][j][k]([n]) */
int lbx_sim_get_local_density_field(const lbx_sim *sim, int level, double *out, size_t n);
int lbx_sim_get_local_velocity_field(const lbx_sim *sim, int level, double *out, size_t n);
/* AmrCore::InitFromScratch, regrid */
int lbx_sim_init_from_scratch(lbx_sim *sim, double time);
int lbx_sim_regrid(lbx_sim *sim, int lbase, double time);
/* Iterate, CalcHydroVars, CalcEquilibriumDist (:147-149) */
int lbx_sim_iterate(lbx_sim *sim, int nsteps);
int lbx_sim_calc_hydro_vars(lbx_sim *sim, int level);
int lbx_sim_calc_equilibrium_dist(lbx_sim *sim, int level);
/* GetDensity / GetVelocity (:145-146): value through *out; sentinels -1.0 / -3e8 off-level */
int lbx_sim_get_density(const lbx_sim *sim, int i, int j, int k, int level, double *out);
int lbx_sim_get_velocity(const lbx_sim *sim, int i, int j, int k, int n, int level, double *out);
/* bulk output (addition): dense over the level's domain, C-ordered [i][j][k]([n]) */
int lbx_sim_get_density_field(const lbx_sim *sim, int level, double *out, size_t n);
int lbx_sim_get_velocity_field(const lbx_sim *sim, int level, double *out, size_t n);
/* addition: generic derived variable = linear moment of the populations
 * (include/derived_var.h:55-91; d3q15_bgk.h:34-55): out[..][c] = sum_p weights[c*15+p] f_p, c < ncomp <= 10,
 * divided by rho when per_unit_density; dense over the level's domain, C-ordered [i][j][k][c], cells the
 * level does not hold = sentinel. */
int lbx_sim_get_linear_moment_field(const lbx_sim *sim, int level, const double *weights, int ncomp,
                                    int per_unit_density, double sentinel, double *out, size_t n);
/* addition (the reference has no I/O): checkpoint of clocks, tau ladder, refinement criteria,
 * box lists and valid-cell populations; read into a sim created with the same extents and max level.  A
 * restarted run continues bit for bit. */
/* addition: AMReX-format plotfile directory (Header, Level_l/Cell_H, Cell_D_00000) with rho, ux, uy, uz of every level */
int lbx_sim_write_plotfile(lbx_sim *sim, const char *dir);
int lbx_sim_write_checkpoint(lbx_sim *sim, const char *path);
int lbx_sim_read_checkpoint(lbx_sim *sim, const char *path);
/* GetTime, GetTimeStep, GetDims, GetExtent (:130-139, 154) */
int lbx_sim_get_time(const lbx_sim *sim, int level, double *out);
int lbx_sim_get_time_step(const lbx_sim *sim, int level, int *out);
int lbx_sim_get_dims(const lbx_sim *sim, int dims[3]);
int lbx_sim_get_extent(const lbx_sim *sim, int level, int lo[3], int hi[3]);
/* SetStaticRefinement / UnsetStaticRefinement (:150-152) */
int lbx_sim_set_static_refinement(lbx_sim *sim, int level, const int lo[3], const int hi[3]);
int lbx_sim_unset_static_refinement(lbx_sim *sim, int level);
/* addition: record the static box of `level` without regridding, then regrid every level once from level 0
 * (moving the boxes of an L-level hierarchy costs one regrid instead of L) */
int lbx_sim_set_static_box(lbx_sim *sim, int level, const int lo[3], const int hi[3]);
int lbx_sim_regrid_all(lbx_sim *sim);
/* AmrCore: maxLevel, finestLevel, refRatio, boxArray(level) */
int lbx_sim_max_level(const lbx_sim *sim);
int lbx_sim_finest_level(const lbx_sim *sim);
int lbx_sim_ref_ratio(const lbx_sim *sim, int level, int ratio[3]);
int lbx_sim_num_boxes(const lbx_sim *sim, int level);             /* AmrCore::boxArray(level).size() */
int lbx_sim_get_boxes(const lbx_sim *sim, int level, int *boxes /* [6 * num_boxes] */);

/* ---- white-box access, as tests/AmrTest.h re-exports it ---- */
enum { LBX_FIELD_DISTFN = 0, LBX_FIELD_DENSITY = 1, LBX_FIELD_VELOCITY = 2, LBX_FIELD_DISTFN_NEXT = 3,
       LBX_FIELD_FINE_MASK = 4 };
int lbx_sim_field_empty(const lbx_sim *sim, int level, int field);    /* 1 empty, 0 not, <0 error */
int lbx_sim_field_num_boxes(const lbx_sim *sim, int level, int field);
int lbx_sim_field_boxes(const lbx_sim *sim, int level, int field, int *boxes);
/* box `b` of the field incl. ghost cells as [comp][z][y][x]; n = capacity of out in elements.
 * Mask data is returned converted to double. */
int lbx_sim_field_fab(const lbx_sim *sim, int level, int field, int b, double *out, size_t n, int shape[4]);
int lbx_sim_get_tau(const lbx_sim *sim, int level, double *tau_s, double *tau_b);
int lbx_sim_get_mass(const lbx_sim *sim, int level, double *mass);
int lbx_sim_get_dt(const lbx_sim *sim, int level, double *dt);
int lbx_sim_num_levels_allocated(const lbx_sim *sim);   /* levels.size() == tau_s.size() == ... */
/* ErrorEst on a TagBoxArray built on the boxes of `tag_level_boxes` (6 ints each), pre-set to
 * `preset` (0 CLEAR, 2 SET) on the level's AmrCore boxArray; out: one char per cell of each box, box after box */
int lbx_sim_call_error_est(lbx_sim *sim, int level, const int *tag_boxes, int ntag_boxes, int preset,
                           char *out, size_t n);
int lbx_sim_call_make_new_level_from_scratch(lbx_sim *sim, int level, const int *boxes, int nboxes, double time);
int lbx_sim_call_make_new_level_from_coarse(lbx_sim *sim, int level, const int *boxes, int nboxes);
int lbx_sim_call_remake_level(lbx_sim *sim, int level, double time, const int *boxes, int nboxes);
int lbx_sim_call_clear_level(lbx_sim *sim, int level);

/* ---- grid-generation metadata only (no GPU needed): the amrex-mini box calculus behind
 * AmrCore::InitFromScratch / regrid.  Functions returning boxes write up to `cap` boxes
 * (6 ints each) and return the number of boxes, or -1 on error. ---- */
int lbx_meta_base_grids(const int dims[3], int max_grid_size, int *boxes, int cap);
int lbx_meta_max_size(const int *in_boxes, int n, int chunk, int *boxes, int cap);
int lbx_meta_simplify(const int *in_boxes, int n, int *boxes, int cap);
int lbx_meta_complement(const int region[6], const int *in_boxes, int n, int *boxes, int cap);
int lbx_meta_cluster(const int *points /* [3 * npoints] */, int npoints, double efficiency, int *boxes, int cap);
/* distributed grid generation WITHOUT a GPU (tests of the host logic over gloo / MPI): after this, the
 * lbx_meta_mesh of every rank tags only the boxes it owns and regrid merges the tag runs through
 * `allgather` -- the same code path lbx_sim_global_init_parallel runs on the GPU box. */
int lbx_meta_parallel_init(int rank, int nranks,
                           int (*allgather)(const void *send, size_t bytes, void *recv, void *user), void *user);
int lbx_meta_parallel_finalise(void);
/* box ownership of a distributed run (amrex::DistributionMapping(ba, nprocs)): owners[i] = rank of box i, the same
 * on every rank by construction.  lbx_meta_distribution: a uniform run (one z-slab per rank where the layers divide,
 * else one contiguous share of the (z, y, x)-ordered boxes balanced by cells).  _runs: runs_per_rank > 1 is what a
 * hierarchy's levels get (AmrMesh::MakeDistributionMap): nprocs * runs_per_rank runs of equal cell count dealt to the
 * ranks in turn. */
int lbx_meta_distribution(const int *in_boxes, int n, int nprocs, int *owners);
int lbx_meta_distribution_runs(const int *in_boxes, int n, int nprocs, int runs_per_rank, int *owners);
/* A field-less AmrCore with the reference's static-box tagging (TagCell, src/AmrSim.cpp:413-417):
 * create, InitFromScratch, then set / unset static boxes (each triggers regrid(level)), query grids. */
typedef struct lbx_meta_mesh lbx_meta_mesh;
int lbx_meta_mesh_create(const int dims[3], int max_level, int max_grid_size, lbx_meta_mesh **out);
int lbx_meta_mesh_destroy(lbx_meta_mesh *m);
int lbx_meta_mesh_set_static(lbx_meta_mesh *m, int level, const int lo[3], const int hi[3]);
int lbx_meta_mesh_unset_static(lbx_meta_mesh *m, int level);
int lbx_meta_mesh_finest_level(const lbx_meta_mesh *m);
int lbx_meta_mesh_boxes(const lbx_meta_mesh *m, int level, int *boxes, int cap);
/* log of hook calls since creation: "S<l>" scratch, "C<l>" from coarse, "R<l>" remake, "X<l>" clear */
int lbx_meta_mesh_log(const lbx_meta_mesh *m, char *out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* LAMBREX_C_H */
