/*
 * lbx.h -- C ABI of the B200-native LAMBReX hot path (liblbx.so).
 *
 * This is the thin device layer the C++ host code (lambrex_b200/host/AmrSim.*)
 * calls; nothing above AmrSim sees it.  The reference has no FFI layer: its
 * boundary is the C++ class API (/root/reference/include/AmrSim.h:127-155,
 * include/lambrex.h:6-7), whose private/protected members do the work these
 * entry points replace.  Each entry point cites the reference member it
 * replaces.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     lbx_last_error() then returns a message (per thread).
 *   - all calls come from one host thread; kernels and copies are queued
 *     asynchronously on the library's stream (or the one given to
 *     lbx_set_stream) unless the name says _sync.
 *   - device memory is owned by the library (lbx_malloc/lbx_free); host
 *     buffers are owned by the caller.
 *   - a "fab" is a rectangular box of cells (valid region + ghosts) stored
 *     SoA: x fastest, then y, z, component slowest -- the FArrayBox order the
 *     reference runs on.  Cells are addressed by GLOBAL integer indices.
 *   - no CPU fallback: without a CUDA device lbx_init fails.
 */
#ifndef LBX_H
#define LBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBX_NV 15   /* populations per cell (NMODES, include/AmrSim.h:21) */
#define LBX_ND 3    /* dimensions          (NDIMS,  include/AmrSim.h:20) */
#define LBX_HALO 2  /* DistFn ghost width  (include/d3q15_bgk.h:11)      */

typedef struct lbx_fab {
  void *data;      /* device pointer: component 0 of cell lo[]           */
  int32_t lo[3];   /* lower corner of the ALLOCATED box (valid - ghosts) */
  int32_t n[3];    /* allocated extents                                  */
  int32_t ncomp;   /* components (15 DistFn, 1 Density, 3 velocity ...)  */
  int32_t dtype;   /* LBX_F64 or LBX_I32                                 */
} lbx_fab;
enum { LBX_F64 = 0, LBX_I32 = 1 };

typedef struct lbx_box { int32_t lo[3], hi[3]; } lbx_box;   /* inclusive */

typedef struct lbx_domain {   /* index domain of a level + per-direction boundary treatment */
  int32_t lo[3], hi[3];
  int32_t periodic[3];        /* 1: periodic wrap; 0: the fab carries ghost cells in that direction;
                                 2 (lbx_collide_stream, push scheme only): solid no-slip walls at both domain faces,
                                 half-way bounce-back -- an addition, the reference aborts on non-periodic directions */
} lbx_domain;
enum { LBX_BC_GHOSTS = 0, LBX_BC_PERIODIC = 1, LBX_BC_WALL = 2 };

/* options for lbx_set_option */
enum {
  LBX_OPT_COLLIDE_LITERAL = 1, /* 1: reference operation order, no FMA (bit-exact vs a
                                  non-FMA CPU build); 0 (default): fast structured form */
  LBX_OPT_SMEM_PAD = 2,        /* bytes of dynamic shared memory added to the fused kernels' launches to
                                  cap resident CTAs per SM (tuning knob; 0 = uncapped, max 49152) */
  LBX_OPT_VALID_TILING = 3,    /* valid-cell tiles of lbx_mf_collide_stream*: 0 = a warp per box row,
                                  1 = 256 consecutive cells of the box (every lane busy)            */
  LBX_OPT_ALIGN_ROWS = 5,      /* fab sets created from now on: 1 = fp64 fabs with ghost cells get sector-
                                  aligned valid rows (x extent padded, see lbx_mf_fab); 0 (default) = tight.
                                  Measured (profiles/r01_alignment.md): 1.5x on kernels that write valid
                                  cells only, 0.7x on the passes that also write ghost cells            */
  LBX_OPT_XGHOST_IN_ROW = 6,   /* lbx_mf_collide_stream (ghosts from own cells): the warp of a valid row also pushes
                                  the row's 4 x-ghost cells, completing the row's partial end sectors at once */
  LBX_OPT_PLAIN_STORES = 7,    /* lbx_mf_collide_stream*: valid-cell pushes as write-back instead of streaming stores */
  LBX_OPT_ROW_KERNEL = 8,      /* lbx_mf_collide_stream*: 1 (default) = the row-owner kernel (a warp per source row of the grown
                                  box writes whole destination rows: every sector completed by the warp that opens it);
                                  0 = the round-1 tile kernel.  Results are bit-identical.                          */
  LBX_OPT_DEBUG_SKIP = 4,      /* PROFILING ONLY (results are wrong): bit 0 skips the valid-cell work of
                                  lbx_mf_collide_stream*, bit 1 the ghost-cell work (tile kernel)               */
};
/* fused step schemes for lbx_collide_stream */
enum {
  LBX_PUSH = 0,  /* dst(x + c_p, p) = collide(src(x, .))_p : F <- S(C(F)), one reference step */
  LBX_PULL = 1,  /* dst(x, .) = collide(src(x - c_p, p))   : G <- C(S(G))                    */
};

/* ---- context: replaces lambrexInit / lambrexFinalise (src/lambrex.cpp:4-14) ---- */
int lbx_init(int device);            /* device < 0: use $LOCAL_RANK or 0 */
int lbx_finalize(void);
int lbx_initialized(void);
const char *lbx_last_error(void);
int lbx_device_count(int *count);
int lbx_device_info(char *name, int name_cap, int *sm_count, size_t *total_bytes, size_t *free_bytes);
int lbx_set_option(int key, int value);
int lbx_set_stream(void *cuda_stream);   /* NULL: back to the library's own stream;
                                            the legacy default stream is cudaStreamLegacy (0x1) */
/* ---- distributed runs: one process per GPU of one NVSwitch box ----
 * Boxes of every level belong to ranks (lbx_mf_create_dist); a rank allocates only its own boxes and
 * reaches the others' through CUDA-IPC peer pointers: gather plans and the fused FillPatch read
 * neighbour boxes straight out of the owner's HBM over NVLink.  Ranks run the same sequence of calls;
 * calls that read peer memory are bracketed by a device-side all-rank barrier (lbx_par_barrier,
 * 8-byte flags in peer memory, no host round trip).  `allgather(send, bytes, recv, user)` gathers
 * `bytes` from every rank into recv in rank order and returns 0 (torch.distributed / MPI plumbing);
 * it only carries IPC handles when fields are created.  lbx_par_init is collective. */
int lbx_par_init(int rank, int world, int (*allgather)(const void *send, size_t bytes, void *recv, void *user),
                 void *user);
/* host-only variant (no CUDA device): rank / world / allgather for the grid-generation metadata of a
 * distributed regrid; device collectives stay unavailable.  Not combinable with lbx_init. */
int lbx_par_init_host(int rank, int world, int (*allgather)(const void *send, size_t bytes, void *recv, void *user),
                      void *user);
int lbx_par_finalize_host(void);
int lbx_par_info(int *rank, int *world, uint64_t *barriers);
int lbx_par_barrier(void);
/* the registered allgather, for host metadata the ranks must agree on (regrid tag lists):
 * recv holds `bytes` from every rank, rank order; a plain copy on a single rank */
int lbx_par_allgather(const void *send, size_t bytes, void *recv);
int lbx_sync(void);
/* Concurrent section: the calls made between begin and end must be independent of one another
 * (they may read the same data but write disjoint data).  Each launch goes to its own auxiliary
 * stream, ordered after everything queued before begin; everything queued after end is ordered
 * behind all of them.  Lets the latency-bound ghost-cell kernels overlap the bandwidth-bound pass
 * over the valid cells.  No host synchronisation; no lbx_sync / lbx_free inside a section. */
int lbx_concurrent_begin(void);
int lbx_concurrent_end(void);
uint64_t lbx_launch_count(void);         /* kernels launched by this library so far */

/* ---- memory: replaces amrex::MultiFab allocation (include/field.h:124-129) ---- */
/* Device memory comes from an arena (the role of AMReX's Arena under every MultiFab): lbx_free /
 * lbx_mf_destroy park the block, the next request of the same (rounded) size reuses it, so
 * rebuilding levels or whole simulations does not pay cudaMalloc/cudaFree again.  Blocks are whole
 * device allocations (CUDA-IPC handles stay valid).  lbx_arena_release returns the parked blocks
 * to the driver (also done by lbx_finalize, and automatically when an allocation fails). */
int lbx_malloc(void **dev_ptr, size_t bytes);
int lbx_free(void *dev_ptr);
int lbx_arena_release(void);
int lbx_arena_info(size_t *in_use_bytes, size_t *cached_bytes, uint64_t *hits, uint64_t *misses);
int lbx_memset(void *dev_ptr, int byte, size_t bytes);
int lbx_host_alloc(void **host_ptr, size_t bytes);   /* pinned */
int lbx_host_free(void *host_ptr);
int lbx_h2d(void *dev_dst, const void *host_src, size_t bytes);
int lbx_d2h(void *host_dst, const void *dev_src, size_t bytes);
int lbx_d2d(void *dev_dst, const void *dev_src, size_t bytes);

/* ---- timing on the launching stream (CUDA events) ---- */
int lbx_timer_start(void);
int lbx_timer_stop(float *milliseconds);   /* synchronises on the stop event */

/* live timing of the AMR path's dominant kernel: between begin and end every lbx_mf_collide_stream* launch
 * (the fused collide + Stream of one level) is bracketed by CUDA events on the stream it is queued on.
 * end synchronises and returns the summed kernel time, the number of launches, the VALID cells of this
 * rank's boxes those launches updated (x 240 B = algorithmic bytes) and the launches not bracketed (event pool
 * exhausted / inside a concurrent section).  For bench.py's roofline of the AMR leg. */
int lbx_prof_begin(void);
int lbx_prof_end(double *ms_total, uint64_t *launches, double *valid_cells, uint64_t *dropped);
/* the same brackets by kind, as of the last lbx_prof_end: [0] fused level passes (what lbx_prof_end returns),
 * [1] gather plans (lbx_plan_apply: FillBoundary, ParallelCopy, the ADD of sum_fine_to_coarse; their device barriers
 * included), [2] lbx_mf_average_down, [3] unused.  ms4 and n4 hold 4 entries each. */
int lbx_prof_breakdown(double *ms4, uint64_t *n4);

/* ---- kernels ---- */
/* CalcEquilibriumDist, src/AmrSim.cpp:845-931: f <- f_eq(rho, u) on box. */
int lbx_equilibrium(const lbx_fab *f, const lbx_fab *rho, const lbx_fab *u, const lbx_box *box);
/* CalcHydroVars, src/AmrSim.cpp:938-979: rho, u <- moments of f on box. */
int lbx_moments(const lbx_fab *f, const lbx_fab *rho, const lbx_fab *u, const lbx_box *box);
/* Collide / CoarseCollide, src/AmrSim.cpp:25-107, 487-580: dst <- collide(src) on box
 * (src may equal dst).  mask != NULL: cells with mask == fine_val get 15 zeros. */
int lbx_collide(const lbx_fab *src, const lbx_fab *dst, const lbx_box *box, double omega_s,
                double omega_b, const lbx_fab *mask, int fine_val);
/* Stream + PropagatePoint, src/AmrSim.cpp:109-122, include/component.h:22-29:
 * dst(x,p) = src(x - c_p, p) for x in box; wraps where dom->periodic, else reads ghosts. */
int lbx_stream(const lbx_fab *src, const lbx_fab *dst, const lbx_box *box, const lbx_domain *dom);
/* CollideAndStream, include/AmrSim.h:89-94 (CollideLevel + Stream + UpdateNow) fused into
 * one pass over the box: 15 loads + 15 stores per cell. */
int lbx_collide_stream(const lbx_fab *src, const lbx_fab *dst, const lbx_box *box,
                       const lbx_domain *dom, double omega_s, double omega_b, int scheme);

/* ---- multi-GPU uniform path: z-slabs, one per GPU ----
 * Replaces the FillBoundary of CollideLevel (src/AmrSim.cpp:132) between boxes owned by
 * different GPUs.  Two transports:
 *   peer stores : lbx_collide_stream_slab writes the 5 populations crossing each z face
 *                 straight into the neighbour's fab through CUDA-IPC peer pointers
 *                 (lbx_ipc_*), ordered per step by lbx_peer_signal / lbx_peer_wait;
 *   packed halos: lbx_halo_pack / lbx_halo_unpack move the 5 crossing populations of a
 *                 face between a fab and a contiguous buffer the caller sends (NCCL). */
#define LBX_IPC_HANDLE_BYTES 64
int lbx_ipc_get_handle(void *dev_ptr, unsigned char *handle /* [LBX_IPC_HANDLE_BYTES] */);
int lbx_ipc_open_handle(const unsigned char *handle, void **peer_ptr);
int lbx_ipc_close_handle(void *peer_ptr);
/* store `value` (release, system scope) to up to two flags (NULL is skipped), after all
 * work queued so far on the stream. */
int lbx_peer_signal(uint64_t *flag_a, uint64_t *flag_b, uint64_t value);
/* block the STREAM (not the host) until both flags hold >= value; after timeout_ns the
 * wait gives up and lbx_peer_error() / lbx_sync() report it. */
int lbx_peer_wait(const uint64_t *flag_a, const uint64_t *flag_b, uint64_t value, uint64_t timeout_ns);
int lbx_peer_error(void);   /* 1 if a wait timed out since the last call; clears the flag */
/* One reference step (CollideAndStream, include/AmrSim.h:89-94) of this rank's slab `box`
 * (global indices).  Destination planes outside `dst` after the periodic wrap of `dom` are
 * taken from `dst_dn` (k-1) / `dst_up` (k+1): peer fabs, or `dst` itself on one GPU. */
int lbx_collide_stream_slab(const lbx_fab *src, const lbx_fab *dst, const lbx_fab *dst_dn,
                            const lbx_fab *dst_up, const lbx_box *box, const lbx_domain *dom,
                            double omega_s, double omega_b);
/* face: 0 +x, 1 -x, 2 +y, 3 -y, 4 +z, 5 -z; buf holds [5][cells of region] doubles. */
/* The same step for a level stored as ONE ghost-free slab per rank (fab sets created with lbx_mf_create_dist:
 * nfabs == world, owner[r] == r, ngrow 0; slab r spans the periodic domain in x and y), with the neighbour
 * ordering folded into the kernel: ONE launch per time step and rank.  Boundary-plane CTAs wait for the
 * neighbours' previous step, store the 5 face-crossing populations into the neighbours' `next` over NVLink and
 * the last of them publishes this step; interior CTAs never synchronise.  Every rank calls it once per step
 * (same sequence on every rank).  Replaces CollideLevel + FillBoundary + Stream + UpdateNow
 * (src/AmrSim.cpp:124-135, 109-122; include/AmrSim.h:89-94) across GPUs.  With one rank: the plain fused step. */
typedef struct lbx_mf lbx_mf;
int lbx_mf_collide_stream_slab(const lbx_mf *now, lbx_mf *next, const lbx_domain *dom, double omega_s, double omega_b);
/* queue a wait (stream, not host) until both neighbours have published the last step queued with
 * lbx_mf_collide_stream_slab: their face stores into this rank's slab are then complete and visible */
int lbx_par_step_finish(void);
int lbx_halo_pack(const lbx_fab *f, const lbx_box *region, int face, double *buf);
int lbx_halo_unpack(const lbx_fab *f, const lbx_box *region, int face, const double *buf);

/* ---- AMR path: batched per-level operations ----
 * A "fab set" (lbx_mf) is every box of one level/field in ONE device allocation, each box
 * grown by `ngrow` ghosts, `ncomp` SoA planes, zero-filled at creation -- the device side of
 * an amrex::MultiFab / iMultiFab (include/field.h:124-129, src/AmrSim.cpp:668).  Every call
 * below is ONE kernel launch over all boxes of the set. */
int lbx_mf_create(const lbx_box *valid, int nfabs, int ncomp, int ngrow, int dtype, lbx_mf **out);
/* distributed (after lbx_par_init, COLLECTIVE, same arguments on every rank): owner[i] = rank that
 * holds box i (amrex::DistributionMapping).  This rank allocates its own boxes only; the others are
 * mapped through CUDA-IPC and can be READ by gather plans / the fused FillPatch; kernels that write
 * a set skip the boxes of other ranks.  owner == NULL: every box local. */
int lbx_mf_create_dist(const lbx_box *valid, int nfabs, int ncomp, int ngrow, int dtype, const int *owner,
                       lbx_mf **out);
int lbx_mf_destroy(lbx_mf *mf);
int lbx_mf_info(const lbx_mf *mf, int *nfabs, int *ncomp, int *ngrow, int *dtype, size_t *bytes);
/* descriptor of fab i (usable with the single-fab kernels above), its valid box, and its
 * byte offset inside the allocation (host mirrors: fab after fab, [comp][z][y][x] over the ALLOCATED
 * box fab->lo, fab->n -- which in x may be wider than valid + ghosts: unused alignment cells) */
int lbx_mf_fab(const lbx_mf *mf, int i, lbx_fab *fab, lbx_box *valid, size_t *byte_offset);
int lbx_mf_upload(lbx_mf *mf, const void *host, size_t bytes);
int lbx_mf_download(const lbx_mf *mf, void *host, size_t bytes);
int lbx_mf_setval(lbx_mf *mf, double value);                       /* all cells incl. ghosts */
/* CalcEquilibriumDist :845-931 / CalcHydroVars :938-979 on the valid cells of every box */
int lbx_mf_equilibrium(lbx_mf *f, const lbx_mf *rho, const lbx_mf *u);
int lbx_mf_moments(const lbx_mf *f, lbx_mf *rho, lbx_mf *u);
/* Collide :25-107, CoarseCollide :487-580 (mask != NULL), FineCollide :582-590: in place */
int lbx_mf_collide(lbx_mf *f, double omega_s, double omega_b, const lbx_mf *mask, int fine_val);
/* same, out of place: dst valid cells <- collide(src valid cells); src and dst hold the same boxes.
 * Fuses the valid-cell copy of FillPatch (src/AmrSim.cpp:129) with the collision that follows it. */
int lbx_mf_collide2(const lbx_mf *src, lbx_mf *dst, double omega_s, double omega_b, const lbx_mf *mask,
                    int fine_val);
/* Stream :109-122 into a "fresh" fab: dst(x,p) = src(x - c_p, p) on valid grown by 1, every
 * other cell of dst zeroed (the reference never writes ghost ring 2 of its new fab) */
int lbx_mf_stream(const lbx_mf *src, lbx_mf *dst);
/* Rohde-cycle collide + Stream in ONE pass (push form) -- CoarseCollide/FineCollide :487-590 followed
 * by Stream :109-122, the pair RohdeCycle :441-457 runs on every level:
 *   valid cells  : read from `src_valid`, zeroed where mask == fine_val, else collided;
 *   ghost cells  : read from `src_ghost` UNcollided (the reference does not refresh ghosts between its
 *                  collide and its Stream); src_ghost == NULL: valid cells only;
 *   dst(x + c_p, p) = f_p(x) for destinations in valid grown by 1; ghost ring 2 of dst = 0 (fresh fab).
 * zero_invalid != 0 also applies the ZeroInvalidComponents :604-617 that follows the cycle's last
 * Stream.  The three sets hold the same boxes with 2 ghost cells; dst must not alias a source. */
int lbx_mf_collide_stream(const lbx_mf *src_valid, const lbx_mf *src_ghost, lbx_mf *dst, double omega_s,
                          double omega_b, const lbx_mf *mask, int fine_val, int zero_invalid);
/* First half of sum_fine_to_coarse :598 (amrex_avgdown): crse box b (valid AND ghosts) <- mean of the
 * ratio^3 fine cells of fine box b; the allocated fine box must be the refinement of the allocated
 * coarse box (fine ghosts = ratio x coarse ghosts).  The ADD into the coarse level is then a COPY-kind
 * plan applied with LBX_OP_ADD over these patches. */
int lbx_mf_average_down(const lbx_mf *fine, lbx_mf *crse, int ratio);
/* dst = a * x + b * y on the valid cells of every box (MultiFab::LinComb; the time interpolation of
 * FillPatchTwoLevels between two coarse states, used by the conventional subcycling driver -- the
 * reference's DistFnFillPatch :359-391 passes one state and so never interpolates).  Products and sum
 * are rounded separately.  The three sets hold the same boxes and component count; dst may alias. */
int lbx_mf_lincomb(lbx_mf *dst, double a, const lbx_mf *x, double b, const lbx_mf *y);
/* Gradient refinement criterion, the device side of an ErrorEst (:633-663 tags a static box only):
 * tags(x) = set_val on valid cells where sum_d (rho(x+e_d) - rho(x-e_d))^2 / 4 > threshold^2; other
 * cells keep their value.  rho: 1 component, >= 1 filled ghost cell; tags: int32, same boxes. */
int lbx_mf_tag_gradient(const lbx_mf *rho, double threshold, lbx_mf *tags, int set_val);
/* Generic derived variable (include/derived_var.h:55-91): out_c = sum_p weights[c*15 + p] * f_p on the valid
 * cells, c < ncomp <= 10, accumulated from 0.0 in increasing p (multiply and add rounded separately);
 * normalise != 0 divides every component by rho = sum_p f_p.  Density = one row of ones, momentum density =
 * the rows c_x, c_y, c_z (d3q15_bgk.h:43-55), velocity = the same normalised, stress = the rows Q_ab.
 * `weights` is HOST memory (copied into the launch parameters). */
int lbx_mf_linear_moments(const lbx_mf *f, lbx_mf *out, const double *weights, int ncomp, int normalise);
/* ZeroInvalidComponents :604-617: in the ghost shell, f_m = 0 unless pos - 2 c_m is valid */
int lbx_mf_zero_invalid(lbx_mf *f);
/* InitPostCollision :477-482: `comp` = 0 on the outermost `depth` rings of every fab box */
int lbx_mf_zero_ring(lbx_mf *f, int depth, int comp);

/* User-facing arrays are C-ordered, i slowest, component fastest (CLindex, include/AmrSim.h:79-83):
 * user[(((i-lo0)*NY + (j-lo1))*NZ + (k-lo2))*ncomp + n] over `domain`.  `user_dev` is DEVICE memory.
 * from_user: valid cells of every box <- user (InitDensity / InitVelocity, src/AmrSim.cpp:138-295);
 * to_user  : user <- valid cells; cells no box holds keep their content (GetDensity / GetVelocity
 *            :824-843 in bulk; pre-fill the sentinel with lbx_fill_f64).
 * `domain` may cover only an x-range (the slowest user index) of the boxes: cells outside it are skipped, so a
 * large array can be staged through a small device buffer chunk by chunk; y and z must cover the boxes. */
int lbx_mf_from_user(lbx_mf *mf, const double *user_dev, const lbx_box *domain, int ncomp);
int lbx_mf_to_user(const lbx_mf *mf, double *user_dev, const lbx_box *domain, int ncomp);
/* distributed runs: only THIS rank's boxes (which must lie inside `domain`, e.g. the rank's slab); not collective */
int lbx_mf_to_user_local(const lbx_mf *mf, double *user_dev, const lbx_box *domain, int ncomp);
/* The same two moves with the user array in HOST memory (pinned memory gives full PCIe speed): staged through
 * two device buffers of the library chunk by chunk along x, on two streams, so that the copy of one chunk
 * overlaps the transposing kernel of the other -- no whole-array device copy is ever allocated.
 * from_user_host is asynchronous like every other call (user_host must stay valid until lbx_sync);
 * to_user_host returns when user_host is complete.  local_only: this rank's boxes only (no communication);
 * fill != 0: cells no box holds receive fill_value (the off-level sentinels of GetDensity / GetVelocity). */
int lbx_mf_from_user_host(lbx_mf *mf, const double *user_host, const lbx_box *domain, int ncomp);
int lbx_mf_to_user_host(const lbx_mf *mf, double *user_host, const lbx_box *domain, int ncomp, int local_only, int fill,
                        double fill_value);
int lbx_fill_f64(double *dev, size_t n, double value);
/* separable initial condition (addition; a 1024^3 domain cannot be initialised from whole-domain host arrays):
 * valid cells of every local box <- profile_dev[(x_axis - axis_lo) * ncomp + n], a DEVICE array of
 * axis_len * ncomp doubles -- fields that vary along one axis (planar pulse along z, shear wave u_x(y)) */
int lbx_mf_fill_profile(lbx_mf *mf, const double *profile_dev, int axis, int axis_lo, int axis_len, int ncomp);

/* Gather plans: the box-intersection metadata of AMReX's ParallelCopy-type calls, computed
 * once on the host, executed as one launch (one thread per destination cell).
 * COPY: the last matching descriptor of a destination fab wins; ADD: all matches are added
 * in list order (deterministic).  Replaces FillBoundary :21,:132, FillPatchSingleLevel :371,
 * FillPatchTwoLevels+PCInterp :385, InterpFromCoarseLevel :406, sum_fine_to_coarse :598,
 * makeFineMask :425. */
enum { LBX_G_COPY = 0,   /* dst(x) = src(x + shift)                                  */
       LBX_G_PC = 1,     /* dst(x) = src(floor(x / ratio) + shift)   (PCInterp)      */
       LBX_G_AVG = 2,    /* dst(x) = mean of the ratio^3 cells src(ratio x + shift..) */
       LBX_G_CONST = 3,  /* dst(x) = value                                           */
       LBX_G_NONE = 4    /* no source: dst(x) keeps its value; the region only widens the tiled
                            bounding box of its group (put it FIRST in the group)          */ };
enum { LBX_OP_COPY = 0, LBX_OP_ADD = 1 };
typedef struct lbx_gather {
  int32_t dst_fab;      /* index into the destination set; descriptors sorted by it   */
  int32_t group;        /* tile group inside dst_fab (ascending): each (dst_fab, group) is
                           tiled over the bounding box of ITS regions only -- e.g. the six
                           slabs of a ghost shell instead of the whole grown box          */
  int32_t src_set;      /* 0 or 1: which of the two source sets given to apply        */
  int32_t src_fab;      /* index into that set                                        */
  int32_t kind;         /* LBX_G_*                                                    */
  int32_t ratio;        /* refinement ratio for PC / AVG                              */
  int32_t shift[3];     /* added in SOURCE index space after the map                  */
  lbx_box region;       /* destination cells                                          */
  double value;         /* LBX_G_CONST                                                */
} lbx_gather;
typedef struct lbx_plan lbx_plan;
int lbx_plan_create(const lbx_gather *g, int n, lbx_plan **out);
int lbx_plan_apply(lbx_plan *plan, lbx_mf *dst, const lbx_mf *src0, const lbx_mf *src1, int op);
/* lbx_mf_collide_stream with the level's DistFnFillPatch :359-391 folded in: `ghost_plan` is a
 * ghosts-only FillPatchSingleLevel / FillPatchTwoLevels plan (COPY / PC / NONE descriptors over the
 * ghost slabs of every box of dst; src0 = same-level NOW, src1 = coarse NOW).  A ghost cell x is not
 * written and read back: its thread looks up FillPatch's source for x and pushes dst(x + c_p, p)
 * from there.  NONE / uncovered cells take x's value in `fallback` (may be NULL: 0).  ONE launch:
 * the CTAs of a box's valid cells and of its ghost cells are interleaved so that they overlap. */
int lbx_mf_collide_stream_fillpatch(const lbx_mf *src_valid, lbx_mf *dst, double omega_s, double omega_b,
                                    const lbx_mf *mask, int fine_val, int zero_invalid, lbx_plan *ghost_plan,
                                    const lbx_mf *src0, const lbx_mf *src1, const lbx_mf *fallback);
/* One CONVENTIONAL level step in one pass -- CollideLevel (DistFnFillPatch, Collide, FillBoundary) :124-135
 * followed by Stream :109-122, as IterateLevel :324-333 runs it (the subcycling driver):
 *   valid cells : dst(x + c_p, p) = collide(now(x))_p;
 *   ghost cells : looked up in `ghost_plan` (a ghosts-only FillPatch plan, as above).  A cell FillPatch copies
 *                 from a same-level valid cell (src0, periodic images included) pushes that cell's COLLIDED
 *                 populations -- what FillBoundary leaves there after Collide; a cell under the coarse level
 *                 pushes wa * crse_a (+ wb * crse_b when crse_b != NULL) UNcollided: FillPatchTwoLevels'
 *                 piecewise-constant interpolation of the coarse state, linear in time between two states;
 *   ghost ring 2 of dst = 0 (fresh fab). */
int lbx_mf_collide_stream_level(const lbx_mf *now, lbx_mf *dst, double omega_s, double omega_b, lbx_plan *ghost_plan,
                                const lbx_mf *src0, const lbx_mf *crse_a, double wa, const lbx_mf *crse_b, double wb,
                                const lbx_mf *fallback);
int lbx_plan_destroy(lbx_plan *plan);

/* lattice constants and moment basis the kernels use (host-side query; no GPU needed):
 * M[15][15], Minv[15][15] row-major, c[15][3], w[15].  src/AmrSim.cpp:1037-1073,
 * include/d3q15_bgk.h:14-28. */
void lbx_d3q15_tables(double *M, double *Minv, int32_t *c, double *w);

#ifdef __cplusplus
}
#endif
#endif /* LBX_H */
