/*
 * ORACLE -- test infrastructure only.  NOT product code.
 *
 * Plain-C CPU restatement of LAMBReX's D3Q15 fp64 moment-space ("BGK")
 * collide-and-stream path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this library; the
 * product (lambrex_b200/) never does.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the
 * reference's own golden vectors (/root/reference/tests/pulseRegression.h,
 * extracted to tests/golden/pulse_regression.npz) under Catch2's Approx rule,
 * and against the mode matrices parsed from the reference source.
 *
 * Build: gcc -O3 -ffp-contract=off -fopenmp (no FMA contraction, so the
 * accumulation order below is the rounding order; the reference's golden
 * vectors were produced by such a build).
 *
 * Memory layout everywhere: x fastest, then y, then z, component slowest
 * (the AMReX FArrayBox order the reference runs on).
 *
 * Reference lines each function follows are cited at the function.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "d3q15_tables.h"

#define NV 15
#define ND 3

static double TM[NV][NV];    /* mode matrix            (src/AmrSim.cpp:1037-1054) */
static double TMI[NV][NV];   /* mode matrix inverse    (src/AmrSim.cpp:1056-1073) */
static double TDELTA[ND][ND];/* diag(1/NMODES) -- 1/15, sic (src/AmrSim.cpp:1033-1035) */
static int tables_ready = 0;

static void init_tables(void) {
  if (tables_ready) return;
  for (int m = 0; m < NV; ++m)
    for (int p = 0; p < NV; ++p) {
      TM[m][p] = (double)ORC_M_NUM[m][p] / (double)ORC_M_DEN[m][p];
      TMI[m][p] = (double)ORC_MINV_NUM[m][p] / (double)ORC_MINV_DEN[m][p];
    }
  for (int a = 0; a < ND; ++a)
    for (int b = 0; b < ND; ++b) TDELTA[a][b] = (a == b) ? 1.0 / NV : 0.0;
  tables_ready = 1;
}

void orc_tables(double *M, double *Minv, int *c) {
  init_tables();
  memcpy(M, TM, sizeof(TM));
  memcpy(Minv, TMI, sizeof(TMI));
  for (int p = 0; p < NV; ++p) {
    c[3 * p + 0] = ORC_CX[p];
    c[3 * p + 1] = ORC_CY[p];
    c[3 * p + 2] = ORC_CZ[p];
  }
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* threads used by the OpenMP loops from now on (bench.py: 1 = the faithful figure, the reference has no OpenMP) */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- per-cell kernels --------------------------------------------------- */

/* Second-order equilibrium, expanded form.  src/AmrSim.cpp:879-927. */
void orc_equilibrium_cell(double rho, const double *u, double *f) {
  const double CS2 = 1.0 / 3.0;                       /* include/AmrSim.h:34 */
  double rw0 = rho * 2.0 / 9.0, rw1 = rho / 9.0, rw2 = rho / 72.0;
  double u2x = u[0] * u[0], u2y = u[1] * u[1], u2z = u[2] * u[2];
  double ucx = u[0] / CS2, ucy = u[1] / CS2, ucz = u[2] / CS2;
  double qx = u2x / (2.0 * CS2 * CS2), qy = u2y / (2.0 * CS2 * CS2), qz = u2z / (2.0 * CS2 * CS2);
  double uv = ucx * ucy, vw = ucy * ucz, uw = ucx * ucz;
  double ms = (u2x + u2y + u2z) / (2.0 * CS2);
  double ms2 = (u2x + u2y + u2z) * (1 - CS2) / (2.0 * CS2 * CS2);
  f[0] = rw0 * (1.0 - ms);
  f[1] = rw1 * (1.0 - ms + ucx + qx);
  f[2] = rw1 * (1.0 - ms - ucx + qx);
  f[3] = rw1 * (1.0 - ms + ucy + qy);
  f[4] = rw1 * (1.0 - ms - ucy + qy);
  f[5] = rw1 * (1.0 - ms + ucz + qz);
  f[6] = rw1 * (1.0 - ms - ucz + qz);
  f[7] = rw2 * (1.0 + ucx + ucy + ucz + uv + vw + uw + ms2);
  f[8] = rw2 * (1.0 + ucx + ucy - ucz + uv - vw - uw + ms2);
  f[9] = rw2 * (1.0 + ucx - ucy + ucz - uv - vw + uw + ms2);
  f[10] = rw2 * (1.0 + ucx - ucy - ucz - uv + vw - uw + ms2);
  f[11] = rw2 * (1.0 - ucx + ucy + ucz - uv + vw - uw + ms2);
  f[12] = rw2 * (1.0 - ucx + ucy - ucz - uv - vw + uw + ms2);
  f[13] = rw2 * (1.0 - ucx - ucy + ucz + uv - vw - uw + ms2);
  f[14] = rw2 * (1.0 - ucx - ucy - ucz + uv + vw + uw + ms2);
}

/* Moment-space two-relaxation-time collision, in place.  src/AmrSim.cpp:28-104
 * (and the identical body in CoarseCollide, :502-575). */
void orc_collide_cell(double *f, double omega_s, double omega_b) {
  double mode[NV];
  for (int m = 0; m < NV; ++m) {
    double acc = 0.0;
    for (int p = 0; p < NV; ++p) acc += f[p] * TM[m][p];
    mode[m] = acc;
  }
  double rho = mode[0], v[ND], usq = 0.0;
  for (int a = 0; a < ND; ++a) {
    v[a] = mode[a + 1] / rho;
    usq += v[a] * v[a];
  }
  double S[ND][ND] = {{mode[4], mode[5], mode[6]},
                      {mode[5], mode[7], mode[8]},
                      {mode[6], mode[8], mode[9]}};
  double TrS = 0.0;
  for (int a = 0; a < ND; ++a) TrS += S[a][a];
  for (int a = 0; a < ND; ++a) S[a][a] -= (TrS / ND);
  TrS -= omega_b * (TrS - rho * usq);
  for (int a = 0; a < ND; ++a) {
    for (int b = 0; b < ND; ++b)
      S[a][b] -= omega_s * (S[a][b] - rho * (v[a] * v[b] - usq * TDELTA[a][b]));
    S[a][a] += (TrS / ND);
  }
  mode[4] = S[0][0]; mode[5] = S[0][1]; mode[6] = S[0][2];
  mode[7] = S[1][1]; mode[8] = S[1][2]; mode[9] = S[2][2];
  for (int m = 10; m < NV; ++m) mode[m] = 0.0;      /* ghost modes, :90-94 */
  for (int p = 0; p < NV; ++p) {
    double fp = 0;
    for (int m = 0; m < NV; ++m) fp += mode[m] * TMI[p][m];
    f[p] = fp;
  }
}

/* Density and velocity moments.  src/AmrSim.cpp:957-971 (all 15 rows are
 * formed there; only rows 0..3 are used, which is all we keep). */
void orc_moments_cell(const double *f, double *rho, double *u) {
  double mode[4];
  for (int m = 0; m < 4; ++m) {
    double acc = 0.0;
    for (int p = 0; p < NV; ++p) acc += f[p] * TM[m][p];
    mode[m] = acc;
  }
  *rho = mode[0];
  for (int a = 0; a < ND; ++a) u[a] = mode[a + 1] / mode[0];
}

/* ---- whole-array ops on SoA planes: f[p*n + cell] ------------------------ */

void orc_equilibrium(int64_t n, const double *rho, const double *u, double *f) {
  init_tables();
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    double uu[3] = {u[c], u[n + c], u[2 * n + c]}, ff[NV];
    orc_equilibrium_cell(rho[c], uu, ff);
    for (int p = 0; p < NV; ++p) f[p * n + c] = ff[p];
  }
}

void orc_collide(int64_t n, double *f, double omega_s, double omega_b) {
  init_tables();
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    double ff[NV];
    for (int p = 0; p < NV; ++p) ff[p] = f[p * n + c];
    orc_collide_cell(ff, omega_s, omega_b);
    for (int p = 0; p < NV; ++p) f[p * n + c] = ff[p];
  }
}

void orc_moments(int64_t n, const double *f, double *rho, double *u) {
  init_tables();
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    double ff[NV], uu[3];
    for (int p = 0; p < NV; ++p) ff[p] = f[p * n + c];
    orc_moments_cell(ff, &rho[c], uu);
    for (int a = 0; a < 3; ++a) u[a * n + c] = uu[a];
  }
}

static inline int wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

/* Pull streaming f'(x,i) = f(x - c_i, i) on one fully periodic box.
 * include/component.h:23-29 + the periodic FillBoundary at src/AmrSim.cpp:132. */
void orc_stream_periodic(int nx, int ny, int nz, const double *src, double *dst) {
  int64_t n = (int64_t)nx * ny * nz;
#pragma omp parallel for collapse(2) schedule(static)
  for (int p = 0; p < NV; ++p)
    for (int k = 0; k < nz; ++k) {
      int ks = wrap(k - ORC_CZ[p], nz);
      for (int j = 0; j < ny; ++j) {
        int js = wrap(j - ORC_CY[p], ny);
        const double *s = src + p * n + ((int64_t)ks * ny + js) * nx;
        double *d = dst + p * n + ((int64_t)k * ny + j) * nx;
        int cx = ORC_CX[p];
        for (int i = 0; i < nx; ++i) d[i] = s[wrap(i - cx, nx)];
      }
    }
}

/* n steps F <- S(C(F)) on one periodic box (net effect of CollideAndStream,
 * include/AmrSim.h:89-94; decomposition independent). */
void orc_step_periodic(int nx, int ny, int nz, double *f, double *tmp, double omega_s,
                       double omega_b, int nsteps) {
  int64_t n = (int64_t)nx * ny * nz;
  for (int t = 0; t < nsteps; ++t) {
    orc_collide(n, f, omega_s, omega_b);
    orc_stream_periodic(nx, ny, nz, f, tmp);
    memcpy(f, tmp, sizeof(double) * NV * n);
  }
}

/* ---- CPU baseline: the reference's pass structure on ghosted boxes -------
 *
 * One level-0 step as the reference executes it (src/AmrSim.cpp:124-135,
 * 109-122, include/AmrSim.h:89-94):
 *   1. FillPatchSingleLevel: next <- now on valid cells + 2-deep periodic ghosts
 *   2. Collide(next) in place on valid cells (dense 15x15 twice)
 *   3. next.FillBoundary(periodic)
 *   4. f_prop = freshly allocated fab; pull-stream over valid grown by 1
 *   5. swap(next.f, f_prop); swap(now, next)
 * on a tensor-product box decomposition (pieces per direction given by the
 * caller, AMReX max_grid_size chop) with HALO=2 ghost cells per box.
 * loop_order 0 = x innermost (cache friendly, generous to the reference),
 *            1 = the reference's for_point_in order: i outer, k inner
 *                (include/amr_help.h:87-92) over x-fastest memory.
 * The global SoA array f (no ghosts) is scattered into the boxes before and
 * gathered after; only the steps are timed by the caller via the return
 * value (seconds, omp_get_wtime).
 */
typedef struct {
  int lo[3], n[3];     /* valid lo, valid extent */
  int64_t sy, sz, sc;  /* strides of the ghosted fab */
  double *now, *next;
} obox;

#define HALO 2

static inline int64_t bidx(const obox *b, int i, int j, int k) {
  /* i,j,k are global indices, may lie in the ghost region */
  return (int64_t)(i - b->lo[0] + HALO) + (int64_t)(j - b->lo[1] + HALO) * b->sy +
         (int64_t)(k - b->lo[2] + HALO) * b->sz;
}

static int find_piece(const int *edges, int np, int g) {
  for (int q = 0; q < np; ++q)
    if (g < edges[q + 1]) return q;
  return np - 1;
}

static void fill_ghosts(obox *boxes, int nb, const int *npc, const int *ex, const int *ey,
                        const int *ez, const int *dom, int from_now, int valid_too) {
  /* dst = next of every box; src = (from_now ? now : next) valid cells of the owner */
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < nb; ++b) {
    obox *B = &boxes[b];
    for (int k = B->lo[2] - HALO; k < B->lo[2] + B->n[2] + HALO; ++k) {
      int kin = (k >= B->lo[2] && k < B->lo[2] + B->n[2]);
      int kg = ((k % dom[2]) + dom[2]) % dom[2];
      int qz = find_piece(ez, npc[2], kg);
      for (int j = B->lo[1] - HALO; j < B->lo[1] + B->n[1] + HALO; ++j) {
        int jin = (j >= B->lo[1] && j < B->lo[1] + B->n[1]);
        int jg = ((j % dom[1]) + dom[1]) % dom[1];
        int qy = find_piece(ey, npc[1], jg);
        for (int i = B->lo[0] - HALO; i < B->lo[0] + B->n[0] + HALO; ++i) {
          int iin = (i >= B->lo[0] && i < B->lo[0] + B->n[0]);
          int is_valid = kin && jin && iin;
          if (is_valid && !valid_too) continue;
          int ig = ((i % dom[0]) + dom[0]) % dom[0];
          int qx = find_piece(ex, npc[0], ig);
          const obox *S = &boxes[(qz * npc[1] + qy) * npc[0] + qx];
          const double *sp = (from_now ? S->now : S->next) + bidx(S, ig, jg, kg);
          double *dp = B->next + bidx(B, i, j, k);
          for (int p = 0; p < NV; ++p) dp[p * B->sc] = sp[p * S->sc];
        }
      }
    }
  }
}

double orc_ref_passes(int nx, int ny, int nz, const int *npc, const int *ex, const int *ey,
                      const int *ez, double *f, double omega_s, double omega_b, int nsteps,
                      int loop_order) {
  init_tables();
  int dom[3] = {nx, ny, nz};
  int nb = npc[0] * npc[1] * npc[2];
  int64_t n = (int64_t)nx * ny * nz;
  obox *boxes = (obox *)calloc(nb, sizeof(obox));
  for (int qz = 0; qz < npc[2]; ++qz)
    for (int qy = 0; qy < npc[1]; ++qy)
      for (int qx = 0; qx < npc[0]; ++qx) {
        obox *B = &boxes[(qz * npc[1] + qy) * npc[0] + qx];
        B->lo[0] = ex[qx]; B->n[0] = ex[qx + 1] - ex[qx];
        B->lo[1] = ey[qy]; B->n[1] = ey[qy + 1] - ey[qy];
        B->lo[2] = ez[qz]; B->n[2] = ez[qz + 1] - ez[qz];
        B->sy = B->n[0] + 2 * HALO;
        B->sz = B->sy * (B->n[1] + 2 * HALO);
        B->sc = B->sz * (B->n[2] + 2 * HALO);
        B->now = (double *)calloc(NV * B->sc, sizeof(double));
        B->next = (double *)calloc(NV * B->sc, sizeof(double));
        for (int p = 0; p < NV; ++p)
          for (int k = 0; k < B->n[2]; ++k)
            for (int j = 0; j < B->n[1]; ++j)
              for (int i = 0; i < B->n[0]; ++i)
                B->now[p * B->sc + bidx(B, B->lo[0] + i, B->lo[1] + j, B->lo[2] + k)] =
                    f[p * n + ((int64_t)(B->lo[2] + k) * ny + (B->lo[1] + j)) * nx + B->lo[0] + i];
      }
#ifdef _OPENMP
  double t0 = omp_get_wtime();
#else
  double t0 = 0.0;
#endif
  for (int t = 0; t < nsteps; ++t) {
    /* 1. FillPatch: next <- now, valid + ghosts */
    fill_ghosts(boxes, nb, npc, ex, ey, ez, dom, 1, 1);
    /* 2. collide valid cells of next, in place */
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nb; ++b) {
      obox *B = &boxes[b];
      if (loop_order == 0) {
        for (int k = 0; k < B->n[2]; ++k)
          for (int j = 0; j < B->n[1]; ++j)
            for (int i = 0; i < B->n[0]; ++i) {
              double *c = B->next + bidx(B, B->lo[0] + i, B->lo[1] + j, B->lo[2] + k), ff[NV];
              for (int p = 0; p < NV; ++p) ff[p] = c[p * B->sc];
              orc_collide_cell(ff, omega_s, omega_b);
              for (int p = 0; p < NV; ++p) c[p * B->sc] = ff[p];
            }
      } else {
        for (int i = 0; i < B->n[0]; ++i)
          for (int j = 0; j < B->n[1]; ++j)
            for (int k = 0; k < B->n[2]; ++k) {
              double *c = B->next + bidx(B, B->lo[0] + i, B->lo[1] + j, B->lo[2] + k), ff[NV];
              for (int p = 0; p < NV; ++p) ff[p] = c[p * B->sc];
              orc_collide_cell(ff, omega_s, omega_b);
              for (int p = 0; p < NV; ++p) c[p * B->sc] = ff[p];
            }
      }
    }
    /* 3. FillBoundary(next): ghosts <- post-collision valid of owners */
    fill_ghosts(boxes, nb, npc, ex, ey, ez, dom, 0, 0);
    /* 4. stream into a freshly allocated fab over valid grown by 1; 5. swaps */
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nb; ++b) {
      obox *B = &boxes[b];
      double *prop = (double *)calloc(NV * B->sc, sizeof(double));
      if (loop_order == 0) {
        for (int p = 0; p < NV; ++p)
          for (int k = -1; k <= B->n[2]; ++k)
            for (int j = -1; j <= B->n[1]; ++j)
              for (int i = -1; i <= B->n[0]; ++i) {
                int gi = B->lo[0] + i, gj = B->lo[1] + j, gk = B->lo[2] + k;
                prop[p * B->sc + bidx(B, gi, gj, gk)] =
                    B->next[p * B->sc + bidx(B, gi - ORC_CX[p], gj - ORC_CY[p], gk - ORC_CZ[p])];
              }
      } else {
        for (int i = -1; i <= B->n[0]; ++i)
          for (int j = -1; j <= B->n[1]; ++j)
            for (int k = -1; k <= B->n[2]; ++k)
              for (int p = 0; p < NV; ++p) {
                int gi = B->lo[0] + i, gj = B->lo[1] + j, gk = B->lo[2] + k;
                prop[p * B->sc + bidx(B, gi, gj, gk)] =
                    B->next[p * B->sc + bidx(B, gi - ORC_CX[p], gj - ORC_CY[p], gk - ORC_CZ[p])];
              }
      }
      free(B->next);
      B->next = B->now;   /* swap(next.f, prop) then swap(now, next) */
      B->now = prop;
    }
  }
#ifdef _OPENMP
  double t1 = omp_get_wtime();
#else
  double t1 = 0.0;
#endif
  for (int b = 0; b < nb; ++b) {
    obox *B = &boxes[b];
    for (int p = 0; p < NV; ++p)
      for (int k = 0; k < B->n[2]; ++k)
        for (int j = 0; j < B->n[1]; ++j)
          for (int i = 0; i < B->n[0]; ++i)
            f[p * n + ((int64_t)(B->lo[2] + k) * ny + (B->lo[1] + j)) * nx + B->lo[0] + i] =
                B->now[p * B->sc + bidx(B, B->lo[0] + i, B->lo[1] + j, B->lo[2] + k)];
    free(B->now);
    free(B->next);
  }
  free(boxes);
  return t1 - t0;
}
