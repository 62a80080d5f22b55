// C ABI of the batched AMR-path operations (lbx_mf_*, lbx_plan_*; declared in include/lbx.h).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/lbx.h"
#include "ctx.h"
#include "launch.h"
#include "mf_kernels.cuh"

struct lbx_mf {
  int nfabs = 0, ncomp = 0, ngrow = 0, dtype = LBX_F64;
  int nlocal = 0;                       // boxes this rank owns: launches cover these (table entry q names the q-th in `lid`)
  size_t bytes = 0;                     // all fabs (host-mirror layout)
  size_t local_bytes = 0;               // what this rank allocated
  char* base = nullptr;                 // one device allocation: this rank's fabs
  lbx::DFabT* table = nullptr;          // device table [nfabs]
  std::vector<lbx::DFabT> host;         // same, host side
  std::vector<size_t> offset;           // byte offset of fab i in the host-mirror layout (all fabs, back to back)
  std::vector<int> owner;               // rank that owns fab i (distributed runs; all 0 otherwise)
  std::vector<size_t> own_off;          // byte offset of fab i inside ITS OWNER's allocation
  bool dist = false;                    // some fabs live in peers' HBM
  long long max_valid = 0;              // largest valid-box cell count
  uint64_t geom = 0;                    // signature of (boxes, ngrow, ncomp, dtype)
  // row-owner kernel (rows_kernel.cuh): every fab is tight (valid + 2 ghost cells per side) with an even row length
  bool rows_ok = false;
  int max_n0 = 0;                       // longest allocated row
  long long max_rows = 0;               // most grown rows (n1 * n2) of any fab
  int max_extent(int dir) const {           // largest valid-box extent along dir
    int m = 0;
    for (const auto& f : host) m = std::max(m, f.vhi[dir] - f.vlo[dir] + 1);
    return m;
  }
  long long max_shell(int grow) const {     // most ghost-shell cells (grown minus valid) of any fab, incl. x alignment cells
    long long m = 0;
    for (const auto& f : host) {
      long long v = 1, c = 1;
      for (int d = 0; d < 3; ++d) {
        v *= (f.vhi[d] - f.vlo[d] + 1);
        c *= d == 0 ? (long long)f.n[0] : (f.vhi[d] - f.vlo[d] + 1 + 2 * grow);
      }
      m = std::max(m, c - v);
    }
    return m;
  }
  long long max_cells(int grow) const {
    long long m = 0;
    for (const auto& f : host) {
      long long c = 1;
      for (int d = 0; d < 3; ++d) c *= (f.vhi[d] - f.vlo[d] + 1 + 2 * grow);
      m = std::max(m, c);
    }
    return m;
  }
};

struct lbx_resolved {                   // FillPatch sources per ghost cell of one destination geometry (k_plan_resolve)
  int2* tab = nullptr;
  long long* first = nullptr;
};
struct lbx_plan {
  std::map<std::tuple<uint64_t, uint64_t, uint64_t>, lbx_resolved> resolved;
  std::vector<lbx::GDesc> descs;
  std::vector<lbx::GDst> dsts;
  lbx::GDesc* d_descs = nullptr;
  lbx::GDst* d_dsts = nullptr;
  long long max_cells = 0;
  bool has_avg = false, has_const = false;
  int* d_fab_first = nullptr;           // [nfabs + 1] first group of each destination fab (lbx_mf_collide_stream_fillpatch)
  int fab_first_n = -1, max_groups = 0;
  std::set<std::tuple<uint64_t, uint64_t, uint64_t>> validated;
  std::map<uint64_t, bool> shell_only;          // per destination geometry: every region lies in the ghost shell
};

namespace {
using lbx::fail;
lbx::Ctx& g = lbx::g_ctx;
const lbx::Launchers& L() { return g.literal ? lbx::launchers_literal() : lbx::launchers_fast(); }

// live timing brackets (lbx_prof_begin / lbx_prof_end): -1 when profiling is off or the bracket cannot be taken
int prof_open(int kind) {
  if (!g.prof) return -1;
  if (g.prof_n >= lbx::Ctx::PROF_MAX || g.conc_next >= 0) { if (kind == 0) ++g.prof_dropped; return -1; }
  const int i = g.prof_n++;
  g.prof_kind[i] = (unsigned char)kind;
  cudaEventRecord(g.prof_ev[2 * i], g.cur);
  return i;
}
void prof_close(int i) {
  if (i >= 0) cudaEventRecord(g.prof_ev[2 * i + 1], g.cur);
}
struct ProfScope {          // brackets everything a function queues, its barriers included
  int i;
  explicit ProfScope(int kind) : i(prof_open(kind)) {}
  ~ProfScope() { prof_close(i); }
};

uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  return h;
}
int same_boxes(const lbx_mf* a, const lbx_mf* b, const char* what) {
  if (!a || !b) return fail(std::string(what) + ": null fab set");
  if (a->nfabs != b->nfabs) return fail(std::string(what) + ": fab sets differ in size");
  for (int i = 0; i < a->nfabs; ++i)
    for (int d = 0; d < 3; ++d)
      if (a->host[i].vlo[d] != b->host[i].vlo[d] || a->host[i].vhi[d] != b->host[i].vhi[d])
        return fail(std::string(what) + ": fab sets have different valid boxes");
  return 0;
}
int need(const lbx_mf* f, int ncomp, int dtype, int ngrow_min, const char* what) {
  if (!f || !f->base) return fail(std::string(what) + ": null or empty fab set");
  if (f->ncomp < ncomp) return fail(std::string(what) + ": too few components");
  if (f->dtype != dtype) return fail(std::string(what) + ": wrong dtype");
  if (f->ngrow < ngrow_min) return fail(std::string(what) + ": too few ghost cells");
  return 0;
}
}  // namespace

extern "C" {

int lbx_mf_create_dist(const lbx_box* valid, int nfabs, int ncomp, int ngrow, int dtype, const int* owner, lbx_mf** out) {
  LBX_NEED_INIT();
  if (!valid || nfabs <= 0 || ncomp <= 0 || ngrow < 0 || !out) return fail("lbx_mf_create: bad arguments");
  if (dtype != LBX_F64 && dtype != LBX_I32) return fail("lbx_mf_create: unknown dtype");
  const int world = g.world, rank = g.rank;
  auto* m = new lbx_mf;
  m->nfabs = nfabs; m->ncomp = ncomp; m->ngrow = ngrow; m->dtype = dtype;
  m->dist = (world > 1 && owner != nullptr);
  const size_t item = dtype == LBX_F64 ? 8 : 4;
  m->host.resize(nfabs);
  m->offset.resize(nfabs);
  m->owner.assign(nfabs, 0);
  m->own_off.resize(nfabs);
  std::vector<size_t> per_rank((size_t)world, 0);      // bytes each rank allocates (same arithmetic on every rank)
  size_t off = 0;
  uint64_t h = mix(mix(mix(0x1234, ncomp), ngrow), dtype);
  for (int i = 0; i < nfabs; ++i) {
    lbx::DFabT& f = m->host[i];
    size_t cells = 1;
    for (int d = 0; d < 3; ++d) {
      if (valid[i].hi[d] < valid[i].lo[d]) { delete m; return fail("lbx_mf_create: empty box"); }
      f.vlo[d] = valid[i].lo[d]; f.vhi[d] = valid[i].hi[d];
      f.lo[d] = f.vlo[d] - ngrow; f.n[d] = f.vhi[d] - f.vlo[d] + 1 + 2 * ngrow;
      if (d == 0 && ngrow > 0 && lbx::g_align_rows) {      // int masks too: the fused pass indexes them with the populations' row offsets
        // 32-byte sector alignment of the VALID rows (profiles/r01_alignment.md): unused lead-in cells
        // put the first valid cell of every row on a sector boundary and the row pitch is a whole
        // number of sectors, so out-of-place row stores are whole sectors instead of partial ones
        // that the L2 has to complete from DRAM (measured: 1.5x on the out-of-place collision).
        const int A = lbx::g_align_rows, lead = (A - ngrow % A) % A;     // A doubles: 4 = a 32 B sector, 8 = a 64 B burst, 16 = a line
        f.lo[0] -= lead;
        f.n[0] = (f.n[0] + lead + A - 1) / A * A;
      }
      cells *= (size_t)f.n[d];
      if (cells >= (size_t(1) << 31)) { delete m; return fail("lbx_mf_create: a fab exceeds 2^31 cells"); }
      h = mix(mix(h, (uint64_t)(uint32_t)f.vlo[d]), (uint64_t)(uint32_t)f.vhi[d]);
      h = mix(mix(h, (uint64_t)(uint32_t)f.lo[d]), (uint64_t)(uint32_t)f.n[d]);      // the layout too (LBX_OPT_ALIGN_ROWS)
    }
    const size_t fab_bytes = (cells * ncomp * item + 255) / 256 * 256;
    m->offset[i] = off;
    off += fab_bytes;
    const int o = m->dist ? owner[i] : 0;
    if (o < 0 || o >= world) { delete m; return fail("lbx_mf_create: owner rank out of range"); }
    m->owner[i] = o;
    m->own_off[i] = per_rank[o];
    per_rank[o] += fab_bytes;
    f.local = (!m->dist || o == rank) ? 1 : 0;
    f.lid = 0;
    if (m->dist) h = mix(h, (uint64_t)o);
  }
  for (int i = 0; i < nfabs; ++i)
    if (m->host[i].local) m->host[m->nlocal++].lid = i;
  m->bytes = off;
  m->geom = h;
  m->max_valid = m->max_cells(0);
  m->rows_ok = (ngrow == 2);
  for (const auto& f : m->host) {
    if (f.n[0] != f.vhi[0] - f.vlo[0] + 5 || (f.n[0] & 1)) m->rows_ok = false;
    m->max_n0 = std::max(m->max_n0, f.n[0]);
    m->max_rows = std::max(m->max_rows, (long long)f.n[1] * f.n[2]);
  }
  const size_t mine = m->dist ? per_rank[rank] : off;
  m->local_bytes = mine;
  cudaError_t e = lbx::arena_alloc(reinterpret_cast<void**>(&m->base), mine ? mine : 256);
  if (e != cudaSuccess) { delete m; return fail(std::string("lbx_mf_create: cudaMalloc: ") + cudaGetErrorString(e)); }
  std::vector<char*> bases((size_t)world, nullptr);
  bases[m->dist ? rank : 0] = m->base;
  if (m->dist) {
    // collective: every rank publishes the CUDA-IPC handle of its allocation and maps the others'
    cudaIpcMemHandle_t hd;
    e = cudaIpcGetMemHandle(&hd, m->base);
    if (e != cudaSuccess) { lbx::arena_free(m->base); delete m; return fail(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    std::vector<unsigned char> all((size_t)world * LBX_IPC_HANDLE_BYTES);
    if (lbx::par_allgather(&hd, LBX_IPC_HANDLE_BYTES, all.data())) { lbx::arena_free(m->base); delete m; return 1; }
    for (int r = 0; r < world; ++r) {
      if (r == rank || per_rank[r] == 0) continue;
      void* b = nullptr;
      if (lbx::ipc_open_cached(all.data() + (size_t)r * LBX_IPC_HANDLE_BYTES, &b)) { lbx::arena_free(m->base); delete m; return 1; }
      bases[r] = static_cast<char*>(b);
    }
  }
  for (int i = 0; i < nfabs; ++i) m->host[i].p = bases[m->owner[i]] + m->own_off[i];
  e = lbx::arena_alloc(reinterpret_cast<void**>(&m->table), sizeof(lbx::DFabT) * nfabs);
  // pageable-host async copy: the driver stages it before returning, so `host` may change later
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->table, m->host.data(), sizeof(lbx::DFabT) * nfabs, cudaMemcpyHostToDevice, g.cur);
  if (e == cudaSuccess && mine) e = cudaMemsetAsync(m->base, 0, mine, g.cur);     // NEW_FAB_FILL = 0
  // a recycled arena block may still be read by a slow peer under its previous identity: nobody
  // proceeds until every rank has got here (and has therefore finished what it queued before).  A rank
  // whose own set-up failed still takes part in the collective, so the peers are not left spinning.
  const int brc = m->dist ? lbx::par_barrier() : 0;
  if (e != cudaSuccess || brc) {
    if (e != cudaSuccess) fail(std::string("lbx_mf_create: ") + cudaGetErrorString(e));
    cudaStreamSynchronize(g.cur);
    lbx::arena_free(m->base);
    if (m->table) lbx::arena_free(m->table);
    delete m;
    return 1;
  }
  *out = m;
  return 0;
}

int lbx_mf_create(const lbx_box* valid, int nfabs, int ncomp, int ngrow, int dtype, lbx_mf** out) {
  return lbx_mf_create_dist(valid, nfabs, ncomp, ngrow, dtype, nullptr, out);
}

int lbx_mf_destroy(lbx_mf* m) {
  if (!m) return 0;
  if (g.ready) {
    if (m->dist) lbx::par_barrier();      // collective: no peer may still be reading these boxes
    cudaStreamSynchronize(g.cur);
    lbx::arena_free(m->base);
    lbx::arena_free(m->table);
  }
  delete m;
  return 0;
}

int lbx_mf_info(const lbx_mf* m, int* nfabs, int* ncomp, int* ngrow, int* dtype, size_t* bytes) {
  if (!m) return fail("lbx_mf_info: null fab set");
  if (nfabs) *nfabs = m->nfabs;
  if (ncomp) *ncomp = m->ncomp;
  if (ngrow) *ngrow = m->ngrow;
  if (dtype) *dtype = m->dtype;
  if (bytes) *bytes = m->bytes;
  return 0;
}

int lbx_mf_fab(const lbx_mf* m, int i, lbx_fab* fab, lbx_box* valid, size_t* byte_offset) {
  if (!m || i < 0 || i >= m->nfabs) return fail("lbx_mf_fab: index out of range");
  const lbx::DFabT& f = m->host[i];
  if (fab) {
    fab->data = f.p;
    for (int d = 0; d < 3; ++d) { fab->lo[d] = f.lo[d]; fab->n[d] = f.n[d]; }
    fab->ncomp = m->ncomp;
    fab->dtype = m->dtype;
  }
  if (valid)
    for (int d = 0; d < 3; ++d) { valid->lo[d] = f.vlo[d]; valid->hi[d] = f.vhi[d]; }
  if (byte_offset) *byte_offset = m->offset[i];
  return 0;
}

// host mirrors hold ALL fabs back to back (lbx_mf_fab's byte_offset); a distributed set moves them
// fab by fab: upload writes this rank's boxes, download also fetches the peers' (over NVLink)
int lbx_mf_upload(lbx_mf* m, const void* host, size_t bytes) {
  LBX_NEED_INIT();
  if (!m || !host || bytes != m->bytes) return fail("lbx_mf_upload: size mismatch");
  if (!m->dist) {
    LBX_CUDA(cudaMemcpyAsync(m->base, host, bytes, cudaMemcpyHostToDevice, g.cur));
    return 0;
  }
  for (int i = 0; i < m->nfabs; ++i) {
    if (!m->host[i].local) continue;
    const size_t n = (i + 1 < m->nfabs ? m->offset[i + 1] : m->bytes) - m->offset[i];
    LBX_CUDA(cudaMemcpyAsync(m->host[i].p, static_cast<const char*>(host) + m->offset[i], n, cudaMemcpyHostToDevice, g.cur));
  }
  return 0;
}
int lbx_mf_download(const lbx_mf* m, void* host, size_t bytes) {
  LBX_NEED_INIT();
  if (!m || !host || bytes != m->bytes) return fail("lbx_mf_download: size mismatch");
  if (!m->dist) {
    LBX_CUDA(cudaMemcpyAsync(host, m->base, bytes, cudaMemcpyDeviceToHost, g.cur));
    return 0;
  }
  if (lbx::par_barrier()) return 1;
  for (int i = 0; i < m->nfabs; ++i) {
    const size_t n = (i + 1 < m->nfabs ? m->offset[i + 1] : m->bytes) - m->offset[i];
    LBX_CUDA(cudaMemcpyAsync(static_cast<char*>(host) + m->offset[i], m->host[i].p, n, cudaMemcpyDeviceToHost, g.cur));
  }
  return lbx::par_barrier();
}

int lbx_mf_setval(lbx_mf* m, double value) {
  LBX_NEED_INIT();
  if (!m) return fail("lbx_mf_setval: null fab set");
  const dim3 grid = lbx::mf_grid(m->max_cells(m->ngrow), m->nlocal);
  if (m->dtype == LBX_F64)
    lbx::k_mf_setval<double><<<grid, lbx::MFT, 0, g.cur>>>(m->table, m->nlocal, m->ngrow, m->ncomp, value);
  else
    lbx::k_mf_setval<int><<<grid, lbx::MFT, 0, g.cur>>>(m->table, m->nlocal, m->ngrow, m->ncomp, (int)value);
  return lbx::after_launch("lbx_mf_setval");
}

int lbx_mf_equilibrium(lbx_mf* f, const lbx_mf* rho, const lbx_mf* u) {
  LBX_NEED_INIT();
  if (need(f, LBX_NV, LBX_F64, 0, "lbx_mf_equilibrium f") || need(rho, 1, LBX_F64, 0, "lbx_mf_equilibrium rho") ||
      need(u, 3, LBX_F64, 0, "lbx_mf_equilibrium u") || same_boxes(f, rho, "lbx_mf_equilibrium") ||
      same_boxes(f, u, "lbx_mf_equilibrium"))
    return 1;
  L().mf_equilibrium(g.cur, f->table, rho->table, u->table, f->nlocal, f->max_valid);
  return lbx::after_launch("lbx_mf_equilibrium");
}

int lbx_mf_moments(const lbx_mf* f, lbx_mf* rho, lbx_mf* u) {
  LBX_NEED_INIT();
  if (need(f, LBX_NV, LBX_F64, 0, "lbx_mf_moments f") || need(rho, 1, LBX_F64, 0, "lbx_mf_moments rho") ||
      need(u, 3, LBX_F64, 0, "lbx_mf_moments u") || same_boxes(f, rho, "lbx_mf_moments") ||
      same_boxes(f, u, "lbx_mf_moments"))
    return 1;
  L().mf_moments(g.cur, f->table, rho->table, u->table, f->nlocal, f->max_valid);
  return lbx::after_launch("lbx_mf_moments");
}

int lbx_mf_collide2(const lbx_mf* src, lbx_mf* dst, double omega_s, double omega_b, const lbx_mf* mask, int fine_val) {
  LBX_NEED_INIT();
  if (need(src, LBX_NV, LBX_F64, 0, "lbx_mf_collide src") || need(dst, LBX_NV, LBX_F64, 0, "lbx_mf_collide dst") ||
      same_boxes(src, dst, "lbx_mf_collide"))
    return 1;
  if (mask && (need(mask, 1, LBX_I32, 0, "lbx_mf_collide mask") || same_boxes(dst, mask, "lbx_mf_collide"))) return 1;
  L().mf_collide(g.cur, src->table, dst->table, mask ? mask->table : nullptr, dst->nlocal, dst->max_valid, omega_s, omega_b,
                 fine_val);
  return lbx::after_launch("lbx_mf_collide");
}
int lbx_mf_collide(lbx_mf* f, double omega_s, double omega_b, const lbx_mf* mask, int fine_val) {
  return lbx_mf_collide2(f, f, omega_s, omega_b, mask, fine_val);
}

static int collide_stream_common(const lbx_mf* src_valid, const lbx_mf* src_ghost, lbx_mf* dst, double omega_s, double omega_b,
                                 const lbx_mf* mask, int fine_val, int zero_invalid, lbx_plan* plan, const lbx_mf* src0,
                                 const lbx_mf* src1, const lbx_mf* fallback, bool level_step = false,
                                 const lbx_mf* src1b = nullptr, double wa = 1.0, double wb = 0.0);

int lbx_mf_collide_stream(const lbx_mf* src_valid, const lbx_mf* src_ghost, lbx_mf* dst, double omega_s, double omega_b,
                          const lbx_mf* mask, int fine_val, int zero_invalid) {
  return collide_stream_common(src_valid, src_ghost, dst, omega_s, omega_b, mask, fine_val, zero_invalid, nullptr, nullptr,
                               nullptr, nullptr);
}

int lbx_mf_stream(const lbx_mf* src, lbx_mf* dst) {
  LBX_NEED_INIT();
  if (need(src, LBX_NV, LBX_F64, 2, "lbx_mf_stream src") || need(dst, LBX_NV, LBX_F64, 1, "lbx_mf_stream dst") ||
      same_boxes(src, dst, "lbx_mf_stream"))
    return 1;
  if (src->base == dst->base) return fail("lbx_mf_stream: src and dst must not alias");
  lbx::k_mf_stream<<<lbx::mf_grid(dst->max_cells(dst->ngrow), dst->nlocal), lbx::MFT, 0, g.cur>>>(
      src->table, dst->table, dst->nlocal, dst->ngrow);
  return lbx::after_launch("lbx_mf_stream");
}

int lbx_mf_average_down(const lbx_mf* fine, lbx_mf* crse, int ratio) {
  LBX_NEED_INIT();
  if (need(fine, LBX_NV, LBX_F64, 0, "lbx_mf_average_down fine") || need(crse, LBX_NV, LBX_F64, 0, "lbx_mf_average_down crse")) return 1;
  if (ratio < 1 || fine->nfabs != crse->nfabs || fine->ncomp != LBX_NV || crse->ncomp != LBX_NV)
    return fail("lbx_mf_average_down: sets differ in size or components");
  for (int i = 0; i < fine->nfabs; ++i)
    for (int d = 0; d < 3; ++d) {
      const lbx::DFabT &f = fine->host[i], &c = crse->host[i];
      // fine box with its ghosts = refine(coarse box with its ghosts)
      if (f.vlo[d] - fine->ngrow != (c.vlo[d] - crse->ngrow) * ratio ||
          f.vhi[d] + fine->ngrow != (c.vhi[d] + crse->ngrow) * ratio + ratio - 1)
        return fail("lbx_mf_average_down: fine box (with ghosts) is not the refinement of the coarse box (with ghosts)");
    }
  const ProfScope timed_scope(2);
  lbx::k_mf_average_down<<<lbx::mf_grid(crse->max_cells(crse->ngrow), crse->nlocal), lbx::MFT, 0, g.cur>>>(
      fine->table, crse->table, crse->nlocal, crse->ngrow, ratio);
  return lbx::after_launch("lbx_mf_average_down");
}

int lbx_mf_lincomb(lbx_mf* dst, double a, const lbx_mf* x, double b, const lbx_mf* y) {
  LBX_NEED_INIT();
  if (need(dst, 1, LBX_F64, 0, "lbx_mf_lincomb dst") || need(x, 1, LBX_F64, 0, "lbx_mf_lincomb x") ||
      need(y, 1, LBX_F64, 0, "lbx_mf_lincomb y") || same_boxes(dst, x, "lbx_mf_lincomb") || same_boxes(dst, y, "lbx_mf_lincomb"))
    return 1;
  if (x->ncomp != dst->ncomp || y->ncomp != dst->ncomp) return fail("lbx_mf_lincomb: component counts differ");
  lbx::k_mf_lincomb<<<lbx::mf_grid(dst->max_valid, dst->nlocal), lbx::MFT, 0, g.cur>>>(dst->table, x->table, y->table, dst->nlocal,
                                                                                     dst->ncomp, a, b);
  return lbx::after_launch("lbx_mf_lincomb");
}

int lbx_mf_tag_gradient(const lbx_mf* rho, double threshold, lbx_mf* tags, int set_val) {
  LBX_NEED_INIT();
  if (need(rho, 1, LBX_F64, 1, "lbx_mf_tag_gradient rho") || need(tags, 1, LBX_I32, 0, "lbx_mf_tag_gradient tags") ||
      same_boxes(rho, tags, "lbx_mf_tag_gradient"))
    return 1;
  if (!(threshold >= 0.0)) return fail("lbx_mf_tag_gradient: threshold must be >= 0");
  lbx::k_mf_tag_gradient<<<lbx::mf_grid(tags->max_valid, tags->nlocal), lbx::MFT, 0, g.cur>>>(rho->table, tags->table, tags->nlocal,
                                                                                            threshold * threshold, set_val);
  return lbx::after_launch("lbx_mf_tag_gradient");
}

int lbx_mf_linear_moments(const lbx_mf* f, lbx_mf* out, const double* weights, int ncomp, int normalise) {
  LBX_NEED_INIT();
  if (need(f, LBX_NV, LBX_F64, 0, "lbx_mf_linear_moments f") || need(out, 1, LBX_F64, 0, "lbx_mf_linear_moments out") ||
      same_boxes(f, out, "lbx_mf_linear_moments"))
    return 1;
  if (!weights || ncomp < 1 || ncomp > lbx::LM_MAX || out->ncomp < ncomp)
    return fail("lbx_mf_linear_moments: 1..10 weight rows, at most the output's component count");
  lbx::LMWeights W;
  memset(&W, 0, sizeof(W));
  memcpy(W.w, weights, sizeof(double) * (size_t)ncomp * LBX_NV);
  lbx::k_mf_linear_moments<<<lbx::mf_grid(out->max_valid, out->nlocal), lbx::MFT, 0, g.cur>>>(f->table, out->table, out->nlocal,
                                                                                            ncomp, normalise ? 1 : 0, W);
  return lbx::after_launch("lbx_mf_linear_moments");
}

int lbx_mf_zero_invalid(lbx_mf* f) {
  LBX_NEED_INIT();
  if (need(f, LBX_NV, LBX_F64, 1, "lbx_mf_zero_invalid")) return 1;
  lbx::k_mf_zero_invalid<<<lbx::mf_grid(f->max_cells(f->ngrow), f->nlocal), lbx::MFT, 0, g.cur>>>(f->table, f->nlocal,
                                                                                               f->ngrow);
  return lbx::after_launch("lbx_mf_zero_invalid");
}

int lbx_mf_zero_ring(lbx_mf* f, int depth, int comp) {
  LBX_NEED_INIT();
  if (need(f, 1, LBX_F64, 1, "lbx_mf_zero_ring")) return 1;
  if (comp < 0 || comp >= f->ncomp || depth < 1 || depth > f->ngrow) return fail("lbx_mf_zero_ring: bad comp/depth");
  lbx::k_mf_zero_ring<<<lbx::mf_grid(f->max_cells(f->ngrow), f->nlocal), lbx::MFT, 0, g.cur>>>(f->table, f->nlocal,
                                                                                            f->ngrow, depth, comp);
  return lbx::after_launch("lbx_mf_zero_ring");
}

static int user_common(const lbx_mf* m, const void* user, const lbx_box* dom, int ncomp, const char* what, bool local_only) {
  if (need(m, ncomp, LBX_F64, 0, what)) return 1;
  if (!user || !dom) return fail(std::string(what) + ": null argument");
  for (const auto& f : m->host) {
    if (local_only && !f.local) continue;
    // x (the slowest user index) may be covered partially: cells outside [lo0, hi0] are skipped (chunked staging)
    for (int d = 1; d < 3; ++d)
      if (f.vlo[d] < dom->lo[d] || f.vhi[d] > dom->hi[d]) return fail(std::string(what) + ": a box lies outside the user array's domain");
  }
  return 0;
}
}  // extern "C"
// tiled launch when the shape allows it (1 or 3 components, grid limits), else the plain kernel
template <bool TO_FAB>
static bool user_tiled(const lbx_mf* m, double* user, const lbx_box* dom, int ncomp, int local_only = 0) {
  int mx = 0, my = 0, mz = 0;
  for (const auto& f : m->host) {
    mx = std::max(mx, f.vhi[0] - f.vlo[0] + 1);
    my = std::max(my, f.vhi[1] - f.vlo[1] + 1);
    mz = std::max(mz, f.vhi[2] - f.vlo[2] + 1);
  }
  if ((ncomp != 1 && ncomp != 3) || m->nfabs > 65535 || my > 65535) return false;
  const int tx = (mx + lbx::UT - 1) / lbx::UT, tz = (mz + lbx::UT - 1) / lbx::UT;
  const dim3 grid((unsigned)(tx * tz), (unsigned)my, (unsigned)m->nfabs), block(lbx::UT, 8);
  const int nx = dom->hi[0] - dom->lo[0] + 1, ny = dom->hi[1] - dom->lo[1] + 1, nz = dom->hi[2] - dom->lo[2] + 1;
  if (ncomp == 1)
    lbx::k_mf_user_tiled<TO_FAB, 1><<<grid, block, 0, g.cur>>>(m->table, user, tx, dom->lo[0], dom->lo[1], dom->lo[2], nx, ny, nz, local_only);
  else
    lbx::k_mf_user_tiled<TO_FAB, 3><<<grid, block, 0, g.cur>>>(m->table, user, tx, dom->lo[0], dom->lo[1], dom->lo[2], nx, ny, nz, local_only);
  return true;
}
extern "C" {
int lbx_mf_from_user(lbx_mf* m, const double* user_dev, const lbx_box* dom, int ncomp) {
  LBX_NEED_INIT();
  if (user_common(m, user_dev, dom, ncomp, "lbx_mf_from_user", true)) return 1;      // writes this rank's boxes only
  if (!user_tiled<true>(m, const_cast<double*>(user_dev), dom, ncomp))
    lbx::k_mf_user<true><<<lbx::mf_grid(m->max_valid, m->nfabs), lbx::MFT, 0, g.cur>>>(
        m->table, m->nfabs, const_cast<double*>(user_dev), dom->lo[0], dom->lo[1], dom->lo[2], dom->hi[0] - dom->lo[0] + 1,
        dom->hi[1] - dom->lo[1] + 1, dom->hi[2] - dom->lo[2] + 1, ncomp, 0);
  return lbx::after_launch("lbx_mf_from_user");
}
int lbx_mf_to_user_local(const lbx_mf* m, double* user_dev, const lbx_box* dom, int ncomp) {
  LBX_NEED_INIT();
  if (user_common(m, user_dev, dom, ncomp, "lbx_mf_to_user_local", true)) return 1;
  if (!user_tiled<false>(m, user_dev, dom, ncomp, 1))
    lbx::k_mf_user<false><<<lbx::mf_grid(m->max_valid, m->nfabs), lbx::MFT, 0, g.cur>>>(
        m->table, m->nfabs, user_dev, dom->lo[0], dom->lo[1], dom->lo[2], dom->hi[0] - dom->lo[0] + 1,
        dom->hi[1] - dom->lo[1] + 1, dom->hi[2] - dom->lo[2] + 1, ncomp, 1);
  return lbx::after_launch("lbx_mf_to_user_local");
}
int lbx_mf_to_user(const lbx_mf* m, double* user_dev, const lbx_box* dom, int ncomp) {
  LBX_NEED_INIT();
  if (user_common(m, user_dev, dom, ncomp, "lbx_mf_to_user", false)) return 1;
  if (m->dist && lbx::par_barrier()) return 1;          // reads every rank's boxes
  if (!user_tiled<false>(m, user_dev, dom, ncomp))
    lbx::k_mf_user<false><<<lbx::mf_grid(m->max_valid, m->nfabs), lbx::MFT, 0, g.cur>>>(
        m->table, m->nfabs, user_dev, dom->lo[0], dom->lo[1], dom->lo[2], dom->hi[0] - dom->lo[0] + 1,
        dom->hi[1] - dom->lo[1] + 1, dom->hi[2] - dom->lo[2] + 1, ncomp, 0);
  if (lbx::after_launch("lbx_mf_to_user")) return 1;
  return m->dist ? lbx::par_barrier() : 0;
}
int lbx_mf_fill_profile(lbx_mf* m, const double* profile_dev, int axis, int axis_lo, int axis_len, int ncomp) {
  LBX_NEED_INIT();
  if (need(m, ncomp, LBX_F64, 0, "lbx_mf_fill_profile")) return 1;
  if (!profile_dev || axis < 0 || axis > 2 || axis_len < 1 || ncomp < 1) return fail("lbx_mf_fill_profile: bad arguments");
  for (const auto& f : m->host)
    if (f.local && (f.vlo[axis] < axis_lo || f.vhi[axis] >= axis_lo + axis_len))
      return fail("lbx_mf_fill_profile: a box reaches outside the profile");
  lbx::k_mf_fill_profile<<<lbx::mf_grid(m->max_valid, m->nlocal), lbx::MFT, 0, g.cur>>>(m->table, m->nlocal, ncomp, profile_dev, axis,
                                                                                      axis_lo);
  return lbx::after_launch("lbx_mf_fill_profile");
}

// ----------------------------------------------------------------------------- distributed uniform path
namespace {
lbx::DFab dfab_of(const lbx::DFabT& t) {
  lbx::DFab d;
  d.p = static_cast<double*>(t.p);
  for (int a = 0; a < 3; ++a) { d.lo[a] = t.lo[a]; d.n[a] = t.n[a]; }
  return d;
}
// rank whose slab holds plane k (same x-y extent as `mine`), or -1
int slab_owner(const lbx_mf* m, const lbx::DFabT& mine, int k) {
  for (int r = 0; r < m->nfabs; ++r) {
    const lbx::DFabT& f = m->host[r];
    if (k >= f.vlo[2] && k <= f.vhi[2] && f.vlo[0] == mine.vlo[0] && f.vhi[0] == mine.vhi[0] && f.vlo[1] == mine.vlo[1] &&
        f.vhi[1] == mine.vhi[1])
      return r;
  }
  return -1;
}
}  // namespace

int lbx_mf_collide_stream_slab(const lbx_mf* now, lbx_mf* next, const lbx_domain* dom, double omega_s, double omega_b) {
  LBX_NEED_INIT();
  const char* what = "lbx_mf_collide_stream_slab";
  if (!dom) return fail(std::string(what) + ": null domain");
  if (need(now, LBX_NV, LBX_F64, 0, what) || need(next, LBX_NV, LBX_F64, 0, what)) return 1;
  if (now->geom != next->geom || now->ngrow != 0 || now->ncomp != LBX_NV) return fail(std::string(what) + ": now and next must be ghost-free 15-component sets over the same slabs");
  if (now->base == next->base) return fail(std::string(what) + ": next aliases now");
  if (now->nfabs != g.world) return fail(std::string(what) + ": one slab per rank expected");
  for (int r = 0; r < now->nfabs; ++r)
    if (now->owner[r] != (now->dist ? r : 0)) return fail(std::string(what) + ": slab r must belong to rank r");
  const lbx::DFabT& S = now->host[g.rank];
  const lbx::DFabT& D = next->host[g.rank];
  lbx::DBox box;
  lbx::DDom dd;
  for (int a = 0; a < 3; ++a) {
    box.lo[a] = S.vlo[a]; box.hi[a] = S.vhi[a];
    dd.lo[a] = dom->lo[a]; dd.hi[a] = dom->hi[a]; dd.periodic[a] = dom->periodic[a];
  }
  for (int a = 0; a < 2; ++a)
    if (box.lo[a] != dd.lo[a] || box.hi[a] != dd.hi[a] || !dd.periodic[a])
      return fail(std::string(what) + ": a slab spans the periodic domain in x and y");
  if (box.hi[1] - box.lo[1] >= 65535 || box.hi[2] - box.lo[2] >= 65535) return fail(std::string(what) + ": slab exceeds 65535 rows/planes per launch");
  int kp = box.hi[2] + 1, km = box.lo[2] - 1;
  if (dd.periodic[2]) { if (kp > dd.hi[2]) kp = dd.lo[2]; if (km < dd.lo[2]) km = dd.hi[2]; }
  else return fail(std::string(what) + ": the domain must be periodic in z");
  const int up = slab_owner(next, D, kp), dn = slab_owner(next, D, km);
  if (up < 0 || dn < 0) return fail(std::string(what) + ": no slab holds the plane above / below this rank's slab");
  lbx::SlabSync sy;
  memset(&sy, 0, sizeof(sy));
  if (g.world > 1) {
    // flags: [0] is written by the rank below, [1] by the rank above (ctx.h); this rank is "above" its lower
    // neighbour and "below" its upper neighbour
    // the wait for the neighbours' previous step is a one-thread launch ahead of the step (they publish at the START of
    // their step, so it returns at once); folded into the kernel it made each of the 2 x 8192 boundary CTAs at 1024^2
    // poll system-scope flags and cost 0.4 ms per step (profiles/r02_scale.md).  The SIGNAL stays in the kernel.
    if (lbx::step_wait_launch(g.step_epoch)) return 1;
    sy.wait_a = nullptr;
    sy.wait_b = nullptr;
    sy.wait_value = g.step_epoch;
    sy.sig_a = g.peer_flag_base[dn] + g.step_off + 1;
    sy.sig_b = g.peer_flag_base[up] + g.step_off + 0;
    sy.sig_value = ++g.step_epoch;
    sy.counter = g.step_flags + 2;
    sy.timeout_ns = 30000000000ull;
    LBX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&sy.err), g.peer_err, 0));
  }
  L().collide_stream_slab_sync(g.cur, dfab_of(S), dfab_of(D), dfab_of(next->host[dn]), dfab_of(next->host[up]), box, dd, omega_s,
                               omega_b, sy);
  return lbx::after_launch(what);
}

// ----------------------------------------------------------------------------- staged host transfers
}  // extern "C"
namespace {
int stage_setup(size_t plane_bytes) {
  const size_t want = std::max(lbx::Ctx::STAGE_BYTES, plane_bytes);
  for (int a = 0; a < 2; ++a) {
    if (!g.xfer[a]) LBX_CUDA(cudaStreamCreateWithFlags(&g.xfer[a], cudaStreamNonBlocking));
    if (!g.xfer_done[a]) LBX_CUDA(cudaEventCreateWithFlags(&g.xfer_done[a], cudaEventDisableTiming));
    if (g.stage[a] && g.stage_bytes < want) {
      LBX_CUDA(cudaStreamSynchronize(g.xfer[a]));
      LBX_CUDA(cudaFree(g.stage[a]));
      g.stage[a] = nullptr;
    }
    if (!g.stage[a]) LBX_CUDA(cudaMalloc(&g.stage[a], want));
  }
  g.stage_bytes = want;
  if (!g.xfer_fork) LBX_CUDA(cudaEventCreateWithFlags(&g.xfer_fork, cudaEventDisableTiming));
  return 0;
}
// kernels of one chunk on stream `st`: user_dev covers the x-range [c0, c1] of dom
template <bool TO_FAB>
int stage_kernel(const lbx_mf* m, double* user_dev, const lbx_box* dom, int c0, int c1, int ncomp, int local_only, cudaStream_t st) {
  lbx_box cb = *dom;
  cb.lo[0] = c0;
  cb.hi[0] = c1;
  cudaStream_t keep = g.cur;
  g.cur = st;
  if (!user_tiled<TO_FAB>(m, user_dev, &cb, ncomp, local_only))
    lbx::k_mf_user<TO_FAB><<<lbx::mf_grid(m->max_valid, m->nfabs), lbx::MFT, 0, st>>>(
        m->table, m->nfabs, user_dev, cb.lo[0], cb.lo[1], cb.lo[2], cb.hi[0] - cb.lo[0] + 1, cb.hi[1] - cb.lo[1] + 1,
        cb.hi[2] - cb.lo[2] + 1, ncomp, local_only);
  g.cur = keep;
  return lbx::after_launch("staged user transfer");
}
}  // namespace
extern "C" {

int lbx_mf_from_user_host(lbx_mf* m, const double* user_host, const lbx_box* dom, int ncomp) {
  LBX_NEED_INIT();
  if (user_common(m, user_host, dom, ncomp, "lbx_mf_from_user_host", true)) return 1;
  if (g.conc_next >= 0) return fail("lbx_mf_from_user_host: must not be called inside a concurrent section");
  const size_t plane = (size_t)(dom->hi[1] - dom->lo[1] + 1) * (dom->hi[2] - dom->lo[2] + 1) * ncomp * sizeof(double);
  if (stage_setup(plane)) return 1;
  const int per = (int)std::max<size_t>(1, g.stage_bytes / plane);
  LBX_CUDA(cudaEventRecord(g.xfer_fork, g.cur));
  for (int a = 0; a < 2; ++a) LBX_CUDA(cudaStreamWaitEvent(g.xfer[a], g.xfer_fork, 0));
  int c = 0;
  for (int x0 = dom->lo[0]; x0 <= dom->hi[0]; x0 += per, ++c) {
    const int x1 = std::min(dom->hi[0], x0 + per - 1), s = c & 1;
    const char* src = reinterpret_cast<const char*>(user_host) + (size_t)(x0 - dom->lo[0]) * plane;
    LBX_CUDA(cudaMemcpyAsync(g.stage[s], src, (size_t)(x1 - x0 + 1) * plane, cudaMemcpyHostToDevice, g.xfer[s]));
    if (stage_kernel<true>(m, static_cast<double*>(g.stage[s]), dom, x0, x1, ncomp, 0, g.xfer[s])) return 1;
  }
  for (int a = 0; a < 2; ++a) {
    LBX_CUDA(cudaEventRecord(g.xfer_done[a], g.xfer[a]));
    LBX_CUDA(cudaStreamWaitEvent(g.cur, g.xfer_done[a], 0));
  }
  return 0;
}

int lbx_mf_to_user_host(const lbx_mf* m, double* user_host, const lbx_box* dom, int ncomp, int local_only, int fill,
                        double fill_value) {
  LBX_NEED_INIT();
  if (user_common(m, user_host, dom, ncomp, "lbx_mf_to_user_host", local_only != 0)) return 1;
  if (g.conc_next >= 0) return fail("lbx_mf_to_user_host: must not be called inside a concurrent section");
  const size_t plane = (size_t)(dom->hi[1] - dom->lo[1] + 1) * (dom->hi[2] - dom->lo[2] + 1) * ncomp * sizeof(double);
  if (stage_setup(plane)) return 1;
  const int per = (int)std::max<size_t>(1, g.stage_bytes / plane);
  const bool remote = m->dist && !local_only;            // reads every rank's boxes
  if (remote && lbx::par_barrier()) return 1;
  LBX_CUDA(cudaEventRecord(g.xfer_fork, g.cur));
  for (int a = 0; a < 2; ++a) LBX_CUDA(cudaStreamWaitEvent(g.xfer[a], g.xfer_fork, 0));
  int c = 0;
  for (int x0 = dom->lo[0]; x0 <= dom->hi[0]; x0 += per, ++c) {
    const int x1 = std::min(dom->hi[0], x0 + per - 1), s = c & 1;
    const size_t bytes = (size_t)(x1 - x0 + 1) * plane;
    if (fill) {
      const size_t n = bytes / sizeof(double);
      lbx::k_fill_f64<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, g.xfer[s]>>>(static_cast<double*>(g.stage[s]),
                                                                                               (long long)n, fill_value);
      if (lbx::after_launch("lbx_mf_to_user_host")) return 1;
    }
    if (stage_kernel<false>(m, static_cast<double*>(g.stage[s]), dom, x0, x1, ncomp, local_only ? 1 : 0, g.xfer[s])) return 1;
    char* dst = reinterpret_cast<char*>(user_host) + (size_t)(x0 - dom->lo[0]) * plane;
    LBX_CUDA(cudaMemcpyAsync(dst, g.stage[s], bytes, cudaMemcpyDeviceToHost, g.xfer[s]));
  }
  for (int a = 0; a < 2; ++a) {
    LBX_CUDA(cudaEventRecord(g.xfer_done[a], g.xfer[a]));
    LBX_CUDA(cudaStreamWaitEvent(g.cur, g.xfer_done[a], 0));
  }
  if (remote && lbx::par_barrier()) return 1;
  LBX_CUDA(cudaStreamSynchronize(g.cur));               // the caller reads user_host next
  return 0;
}

// ----------------------------------------------------------------------------- live kernel timing
int lbx_prof_begin(void) {
  LBX_NEED_INIT();
  if (!g.prof_ev) {
    g.prof_ev = new cudaEvent_t[2 * lbx::Ctx::PROF_MAX];
    for (int i = 0; i < 2 * lbx::Ctx::PROF_MAX; ++i) LBX_CUDA(cudaEventCreate(&g.prof_ev[i]));
    g.prof_kind = new unsigned char[lbx::Ctx::PROF_MAX];
  }
  g.prof = true;
  g.prof_n = 0;
  g.prof_cells = 0.0;
  g.prof_dropped = 0;
  return 0;
}
int lbx_prof_end(double* ms_total, uint64_t* launches, double* valid_cells, uint64_t* dropped) {
  LBX_NEED_INIT();
  if (!g.prof) return fail("lbx_prof_end: lbx_prof_begin was not called");
  g.prof = false;
  LBX_CUDA(cudaStreamSynchronize(g.cur));
  for (int k = 0; k < 4; ++k) { g.prof_kind_ms[k] = 0.0; g.prof_kind_n[k] = 0; }
  for (int i = 0; i < g.prof_n; ++i) {
    float t = 0.f;
    LBX_CUDA(cudaEventElapsedTime(&t, g.prof_ev[2 * i], g.prof_ev[2 * i + 1]));
    g.prof_kind_ms[g.prof_kind[i] & 3] += t;
    ++g.prof_kind_n[g.prof_kind[i] & 3];
  }
  if (ms_total) *ms_total = g.prof_kind_ms[0];
  if (launches) *launches = g.prof_kind_n[0];
  if (valid_cells) *valid_cells = g.prof_cells;
  if (dropped) *dropped = g.prof_dropped;
  return 0;
}

int lbx_prof_breakdown(double* ms4, uint64_t* n4) {
  LBX_NEED_INIT();
  if (!ms4 || !n4) return fail("lbx_prof_breakdown: null output");
  for (int k = 0; k < 4; ++k) { ms4[k] = g.prof_kind_ms[k]; n4[k] = g.prof_kind_n[k]; }
  return 0;
}

int lbx_fill_f64(double* dev, size_t n, double value) {
  LBX_NEED_INIT();
  if (!dev) return fail("lbx_fill_f64: null pointer");
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
  lbx::k_fill_f64<<<blocks, 256, 0, g.cur>>>(dev, (long long)n, value);
  return lbx::after_launch("lbx_fill_f64");
}

// ----------------------------------------------------------------------------- gather plans
int lbx_plan_create(const lbx_gather* gs, int n, lbx_plan** out) {
  LBX_NEED_INIT();
  if (!out || n < 0 || (n > 0 && !gs)) return fail("lbx_plan_create: bad arguments");
  auto* p = new lbx_plan;
  p->descs.resize(n);
  for (int i = 0; i < n; ++i) {
    const lbx_gather& a = gs[i];
    lbx::GDesc& d = p->descs[i];
    if (i > 0 && (a.dst_fab < gs[i - 1].dst_fab || (a.dst_fab == gs[i - 1].dst_fab && a.group < gs[i - 1].group))) {
      delete p;
      return fail("lbx_plan_create: descriptors must be sorted by (dst_fab, group)");
    }
    if (a.kind < LBX_G_COPY || a.kind > LBX_G_NONE) { delete p; return fail("lbx_plan_create: unknown kind"); }
    if ((a.kind == LBX_G_PC || a.kind == LBX_G_AVG) && a.ratio < 1) { delete p; return fail("lbx_plan_create: ratio must be >= 1"); }
    if (a.src_set != 0 && a.src_set != 1) { delete p; return fail("lbx_plan_create: src_set must be 0 or 1"); }
    for (int k = 0; k < 3; ++k) {
      if (a.region.hi[k] < a.region.lo[k]) { delete p; return fail("lbx_plan_create: empty region"); }
      d.lo[k] = a.region.lo[k]; d.hi[k] = a.region.hi[k]; d.shift[k] = a.shift[k];
    }
    if (a.kind == LBX_G_AVG) p->has_avg = true;
    if (a.kind == LBX_G_CONST) p->has_const = true;
    d.src_set = a.src_set; d.src_fab = a.src_fab; d.kind = a.kind; d.ratio = a.ratio > 0 ? a.ratio : 1; d.value = a.value;
    if (p->dsts.empty() || p->dsts.back().fab != a.dst_fab || gs[i - 1].group != a.group) {
      lbx::GDst t;
      t.fab = a.dst_fab; t.first = i; t.count = 0; t.pad = 0;
      for (int k = 0; k < 3; ++k) { t.blo[k] = d.lo[k]; t.bhi[k] = d.hi[k]; }
      p->dsts.push_back(t);
    }
    lbx::GDst& t = p->dsts.back();
    ++t.count;
    for (int k = 0; k < 3; ++k) { t.blo[k] = std::min(t.blo[k], d.lo[k]); t.bhi[k] = std::max(t.bhi[k], d.hi[k]); }
  }
  for (const auto& t : p->dsts) {
    long long c = 1;
    for (int k = 0; k < 3; ++k) c *= (t.bhi[k] - t.blo[k] + 1);
    p->max_cells = std::max(p->max_cells, c);
  }
  if (p->max_cells >= (1ll << 31)) { delete p; return fail("lbx_plan_create: a destination region exceeds 2^31 cells"); }
  if (n > 0) {
    LBX_CUDA(lbx::arena_alloc(reinterpret_cast<void**>(&p->d_descs), sizeof(lbx::GDesc) * n));
    LBX_CUDA(lbx::arena_alloc(reinterpret_cast<void**>(&p->d_dsts), sizeof(lbx::GDst) * p->dsts.size()));
    LBX_CUDA(cudaMemcpyAsync(p->d_descs, p->descs.data(), sizeof(lbx::GDesc) * n, cudaMemcpyHostToDevice, g.cur));
    LBX_CUDA(cudaMemcpyAsync(p->d_dsts, p->dsts.data(), sizeof(lbx::GDst) * p->dsts.size(), cudaMemcpyHostToDevice, g.cur));
    LBX_CUDA(cudaStreamSynchronize(g.cur));
  }
  *out = p;
  return 0;
}

int lbx_plan_destroy(lbx_plan* p) {
  if (!p) return 0;
  if (g.ready) {
    cudaStreamSynchronize(g.cur);
    lbx::arena_free(p->d_descs);
    lbx::arena_free(p->d_dsts);
    lbx::arena_free(p->d_fab_first);
    for (auto& kv : p->resolved) {
      lbx::arena_free(kv.second.tab);
      lbx::arena_free(kv.second.first);
    }
  }
  delete p;
  return 0;
}

// first group of every destination fab (groups are sorted by fab): device array [nfabs + 1]
static int plan_fab_first(lbx_plan* plan, const lbx_mf* dst) {
  if (plan->fab_first_n == dst->nfabs) return 0;
  std::vector<int> first((size_t)dst->nfabs + 1, 0);
  int maxg = 0;
  for (const auto& t : plan->dsts) ++first[(size_t)t.fab + 1];
  for (int f = 0; f < dst->nfabs; ++f) { maxg = std::max(maxg, first[f + 1]); first[f + 1] += first[f]; }
  if (plan->d_fab_first) lbx::arena_free(plan->d_fab_first);
  LBX_CUDA(lbx::arena_alloc(reinterpret_cast<void**>(&plan->d_fab_first), sizeof(int) * first.size()));
  LBX_CUDA(cudaMemcpyAsync(plan->d_fab_first, first.data(), sizeof(int) * first.size(), cudaMemcpyHostToDevice, g.cur));
  LBX_CUDA(cudaStreamSynchronize(g.cur));
  plan->fab_first_n = dst->nfabs;
  plan->max_groups = maxg;
  return 0;
}

static int validate_plan(lbx_plan* p, const lbx_mf* dst, const lbx_mf* s0, const lbx_mf* s1) {
  const auto key = std::make_tuple(dst->geom, s0 ? s0->geom : 0, s1 ? s1->geom : 0);
  if (p->validated.count(key)) return 0;
  // inside valid + ghosts (the allocated box may be wider in x: alignment cells are nobody's data)
  auto inside = [](const lbx::DFabT& f, int ngrow, const int* lo, const int* hi) {
    for (int d = 0; d < 3; ++d)
      if (lo[d] < f.vlo[d] - ngrow || hi[d] > f.vhi[d] + ngrow) return false;
    return true;
  };
  for (const auto& t : p->dsts) {
    if (t.fab < 0 || t.fab >= dst->nfabs) return fail("lbx_plan_apply: destination fab index out of range");
    for (int q = t.first; q < t.first + t.count; ++q) {
      const lbx::GDesc& d = p->descs[q];
      if (!inside(dst->host[t.fab], dst->ngrow, d.lo, d.hi)) return fail("lbx_plan_apply: region outside the destination fab");
      if (d.kind == lbx::G_CONST || d.kind == lbx::G_NONE) continue;
      const lbx_mf* s = d.src_set ? s1 : s0;
      if (!s) return fail("lbx_plan_apply: plan needs a source set that was not given");
      if (d.src_fab < 0 || d.src_fab >= s->nfabs) return fail("lbx_plan_apply: source fab index out of range");
      if (s->dtype != dst->dtype || s->ncomp < dst->ncomp) return fail("lbx_plan_apply: source dtype/components mismatch");
      int lo[3], hi[3];
      for (int k = 0; k < 3; ++k) {
        auto fd = [](int a, int r) { return a >= 0 ? a / r : -((-a + r - 1) / r); };
        if (d.kind == lbx::G_COPY) { lo[k] = d.lo[k] + d.shift[k]; hi[k] = d.hi[k] + d.shift[k]; }
        else if (d.kind == lbx::G_PC) { lo[k] = fd(d.lo[k], d.ratio) + d.shift[k]; hi[k] = fd(d.hi[k], d.ratio) + d.shift[k]; }
        else { lo[k] = d.lo[k] * d.ratio + d.shift[k]; hi[k] = d.hi[k] * d.ratio + d.shift[k] + d.ratio - 1; }
      }
      if (!inside(s->host[d.src_fab], s->ngrow, lo, hi)) return fail("lbx_plan_apply: mapped source region outside the source fab");
    }
  }
  p->validated.insert(key);
  return 0;
}

static int plan_resolved(lbx_plan* plan, const lbx_mf* dst, const lbx_mf* s0, const lbx_mf* s1, lbx_resolved* out);

// every descriptor a set-0 COPY (or NONE) into the 2-cell ghost shell of tight boxes: the plan can run from the
// resolved per-ghost-cell table (k_shell_copy) instead of the descriptor search
static bool plan_shell_only(lbx_plan* p, const lbx_mf* dst) {
  auto it = p->shell_only.find(dst->geom);
  if (it != p->shell_only.end()) return it->second;
  bool ok = dst->rows_ok && !p->has_avg && !p->has_const;
  for (const auto& f : dst->host)
    if (f.local && (f.n[1] != f.vhi[1] - f.vlo[1] + 5 || f.n[2] != f.vhi[2] - f.vlo[2] + 5)) ok = false;
  for (size_t q = 0; ok && q < p->dsts.size(); ++q) {
    const lbx::GDst& t = p->dsts[q];
    if (t.fab < 0 || t.fab >= dst->nfabs) { ok = false; break; }
    const lbx::DFabT& f = dst->host[t.fab];
    for (int d = t.first; d < t.first + t.count; ++d) {
      const lbx::GDesc& r = p->descs[d];
      if (r.kind == lbx::G_NONE) continue;
      if (r.kind != lbx::G_COPY || r.src_set != 0) { ok = false; break; }
      bool meets_valid = true;
      for (int k = 0; k < 3; ++k)
        if (r.hi[k] < f.vlo[k] || r.lo[k] > f.vhi[k]) meets_valid = false;
      if (meets_valid) { ok = false; break; }
    }
  }
  p->shell_only[dst->geom] = ok;
  return ok;
}

int lbx_plan_apply(lbx_plan* p, lbx_mf* dst, const lbx_mf* src0, const lbx_mf* src1, int op) {
  LBX_NEED_INIT();
  if (!p || !dst) return fail("lbx_plan_apply: null plan or destination");
  if (op != LBX_OP_COPY && op != LBX_OP_ADD) return fail("lbx_plan_apply: unknown op");
  const bool remote = (src0 && src0->dist) || (src1 && src1->dist);   // sources may sit in peers' HBM
  if (p->descs.empty()) {
    // a rank that owns no box of the destination still takes part in the two all-rank barriers
    if (remote && (lbx::par_barrier() || lbx::par_barrier())) return 1;
    return 0;
  }
  if (validate_plan(p, dst, src0, src1)) return 1;
  const ProfScope timed_scope(1);
  if (op == LBX_OP_COPY && dst->dtype == LBX_F64 && src0 && lbx::g_row_kernel && !lbx::g_debug_skip && dst->rows_ok &&
      plan_shell_only(p, dst)) {
    lbx_resolved res;
    if (plan_fab_first(p, dst) || plan_resolved(p, dst, src0, nullptr, &res)) return 1;
    if (remote && lbx::par_barrier()) return 1;
    const dim3 gs = lbx::mf_grid(dst->max_shell(2), dst->nlocal);
    if (dst->ncomp == LBX_NV)
      lbx::k_shell_copy<LBX_NV><<<gs, lbx::MFT, 0, g.cur>>>(dst->table, dst->nlocal, src0->table, res.tab, res.first, dst->ncomp);
    else
      lbx::k_shell_copy<0><<<gs, lbx::MFT, 0, g.cur>>>(dst->table, dst->nlocal, src0->table, res.tab, res.first, dst->ncomp);
    if (lbx::after_launch("lbx_plan_apply (shell)")) return 1;
    return remote ? lbx::par_barrier() : 0;
  }
  if (remote && lbx::par_barrier()) return 1;
  const dim3 grid = lbx::mf_grid(p->max_cells, (int)p->dsts.size());
  const lbx::DFabT* t0 = src0 ? src0->table : nullptr;
  const lbx::DFabT* t1 = src1 ? src1->table : nullptr;
  const int nd = (int)p->dsts.size();
#define LBX_PLAN_LAUNCH(T, ADD, NC) \
  lbx::k_plan_apply<T, ADD, NC, false><<<grid, lbx::MFT, 0, g.cur>>>(p->d_dsts, nd, p->d_descs, dst->table, t0, t1, dst->ncomp)
  if (dst->dtype == LBX_F64 && p->has_avg) {
    // averaging plans (sum_fine_to_coarse fused, average_down): 8 fine cells per coarse value -- component-parallel
    // launch, one thread per (cell, component), or a 15-component thread sits on 60 dependent 16-byte loads
    const dim3 gridc = lbx::mf_grid(p->max_cells, nd, 0, dst->ncomp);
    if (op == LBX_OP_COPY)
      lbx::k_plan_apply<double, false, 1, true><<<gridc, lbx::MFT, 0, g.cur>>>(p->d_dsts, nd, p->d_descs, dst->table, t0, t1, dst->ncomp);
    else
      lbx::k_plan_apply<double, true, 1, true><<<gridc, lbx::MFT, 0, g.cur>>>(p->d_dsts, nd, p->d_descs, dst->table, t0, t1, dst->ncomp);
  } else if (dst->dtype == LBX_F64) {
    if (dst->ncomp == LBX_NV) {          // the populations: compile-time component count
      if (op == LBX_OP_COPY) LBX_PLAN_LAUNCH(double, false, LBX_NV);
      else LBX_PLAN_LAUNCH(double, true, LBX_NV);
    } else {
      if (op == LBX_OP_COPY) LBX_PLAN_LAUNCH(double, false, 0);
      else LBX_PLAN_LAUNCH(double, true, 0);
    }
  } else {
    if (op == LBX_OP_COPY) LBX_PLAN_LAUNCH(int, false, 0);
    else LBX_PLAN_LAUNCH(int, true, 0);
  }
#undef LBX_PLAN_LAUNCH
  if (lbx::after_launch("lbx_plan_apply")) return 1;
  return remote ? lbx::par_barrier() : 0;
}

int lbx_mf_collide_stream_fillpatch(const lbx_mf* src_valid, lbx_mf* dst, double omega_s, double omega_b, const lbx_mf* mask,
                                    int fine_val, int zero_invalid, lbx_plan* ghost_plan, const lbx_mf* src0,
                                    const lbx_mf* src1, const lbx_mf* fallback) {
  if (!ghost_plan) return fail("lbx_mf_collide_stream_fillpatch: null plan");
  return collide_stream_common(src_valid, nullptr, dst, omega_s, omega_b, mask, fine_val, zero_invalid, ghost_plan, src0, src1,
                               fallback);
}

int lbx_mf_collide_stream_level(const lbx_mf* now, lbx_mf* dst, double omega_s, double omega_b, lbx_plan* ghost_plan,
                                const lbx_mf* src0, const lbx_mf* crse_a, double wa, const lbx_mf* crse_b, double wb,
                                const lbx_mf* fallback) {
  if (!ghost_plan) return fail("lbx_mf_collide_stream_level: null plan");
  if (crse_b && (!crse_a || crse_a->geom != crse_b->geom))
    return fail("lbx_mf_collide_stream_level: the two coarse states must hold the same boxes");
  return collide_stream_common(now, nullptr, dst, omega_s, omega_b, nullptr, 0, 0, ghost_plan, src0, crse_a, fallback, true, crse_b,
                               wa, wb);
}

// the plan's FillPatch sources resolved per ghost cell of dst's boxes: built once per (plan, geometries)
static int plan_resolved(lbx_plan* plan, const lbx_mf* dst, const lbx_mf* s0, const lbx_mf* s1, lbx_resolved* out) {
  const auto key = std::make_tuple(dst->geom, s0 ? s0->geom : 0, s1 ? s1->geom : 0);
  auto it = plan->resolved.find(key);
  if (it != plan->resolved.end()) { *out = it->second; return 0; }
  std::vector<long long> first((size_t)dst->nfabs, 0);
  long long total = 0, most = 0;
  for (int b = 0; b < dst->nfabs; ++b) {
    const lbx::DFabT& f = dst->host[b];
    if (!f.local) continue;
    first[b] = total;
    const long long n = lbx::ro_shell_size(f.n[0], f.n[1], f.n[2]);
    total += n;
    most = std::max(most, n);
  }
  lbx_resolved r;
  LBX_CUDA(lbx::arena_alloc(reinterpret_cast<void**>(&r.tab), sizeof(int2) * (size_t)std::max<long long>(total, 1)));
  LBX_CUDA(lbx::arena_alloc(reinterpret_cast<void**>(&r.first), sizeof(long long) * first.size()));
  LBX_CUDA(cudaMemcpyAsync(r.first, first.data(), sizeof(long long) * first.size(), cudaMemcpyHostToDevice, g.cur));
  LBX_CUDA(cudaStreamSynchronize(g.cur));          // `first` is a local
  if (total > 0) {
    lbx::k_plan_resolve<<<lbx::mf_grid(most, dst->nlocal), lbx::MFT, 0, g.cur>>>(dst->table, dst->nlocal, plan->d_dsts, plan->d_fab_first,
                                                                              plan->d_descs, s0 ? s0->table : nullptr,
                                                                              s1 ? s1->table : nullptr, r.tab, r.first);
    if (lbx::after_launch("k_plan_resolve")) return 1;
  }
  plan->resolved[key] = r;
  *out = r;
  return 0;
}

static int collide_stream_common(const lbx_mf* src_valid, const lbx_mf* src_ghost, lbx_mf* dst, double omega_s, double omega_b,
                                 const lbx_mf* mask, int fine_val, int zero_invalid, lbx_plan* plan, const lbx_mf* src0,
                                 const lbx_mf* src1, const lbx_mf* fallback, bool level_step, const lbx_mf* src1b, double wa,
                                 double wb) {
  LBX_NEED_INIT();
  const char* what = "lbx_mf_collide_stream";
  if (need(src_valid, LBX_NV, LBX_F64, 2, what) || need(dst, LBX_NV, LBX_F64, 2, what)) return 1;
  if (src_valid->geom != dst->geom || dst->ngrow != 2 || dst->ncomp != LBX_NV)
    return fail("lbx_mf_collide_stream: src_valid and dst must hold the same boxes, 15 components, 2 ghost cells");
  for (const lbx_mf* s : {src_ghost, fallback})
    if (s && s->geom != dst->geom) return fail("lbx_mf_collide_stream: ghost source and dst differ in geometry");
  for (const lbx_mf* s : {src_valid, src_ghost, src0, src1, src1b, fallback})
    if (s && s->base == dst->base) return fail("lbx_mf_collide_stream: dst aliases a source");
  if (mask && (need(mask, 1, LBX_I32, 2, what) || mask->ngrow != 2 || mask->ncomp != 1 || same_boxes(dst, mask, what))) return 1;
  if (dst->max_extent(2) > 65535) return fail("lbx_mf_collide_stream: box extents exceed the launch grid");
  lbx::CSPlan cp;
  memset(&cp, 0, sizeof(cp));
  long long ghost_tiles = 0;
  if (plan) {
    if (plan->has_avg || plan->has_const) return fail("lbx_mf_collide_stream_fillpatch: only COPY / PC / NONE descriptors can be pushed");
    if (plan->descs.empty()) {      // this rank owns no box of the level: only the collective part remains
      const bool rem = (src0 && src0->dist) || (src1 && src1->dist) || (src1b && src1b->dist);
      if (!dst->dist) return fail("lbx_mf_collide_stream_fillpatch: empty plan");
      if (rem && (lbx::par_barrier() || lbx::par_barrier())) return 1;
      return 0;
    }
    if (validate_plan(plan, dst, src0, src1)) return 1;
    if (plan_fab_first(plan, dst)) return 1;
    cp.dsts = plan->d_dsts;
    cp.fab_first = plan->d_fab_first;
    cp.descs = plan->d_descs;
    cp.s0 = src0 ? src0->table : nullptr;
    cp.s1 = src1 ? src1->table : nullptr;
    cp.fb = fallback ? fallback->table : nullptr;
    cp.s1b = src1b ? src1b->table : nullptr;
    cp.wa = wa;
    cp.wb = wb;
    cp.tiles_per_group = (int)((plan->max_cells + lbx::MFT - 1) / lbx::MFT);
    ghost_tiles = (long long)cp.tiles_per_group * plan->max_groups;
  } else if (src_ghost) {
    ghost_tiles = (dst->max_shell(2) + lbx::MFT - 1) / lbx::MFT;
  }
  const bool remote = plan && ((src0 && src0->dist) || (src1 && src1->dist) || (src1b && src1b->dist));
  if (remote && lbx::par_barrier()) return 1;
  // the row-owner kernel needs a ghost source, tight fabs with even rows, and a row buffer that fits shared memory
  int ro_warps = lbx::RO_THREADS / 32;
  const int ro_pitch = dst->max_n0 + 4;          // even; the row buffer keeps a zero pad on either side of the row
  while (ro_warps > 1 && (size_t)ro_warps * LBX_NV * ro_pitch * sizeof(double) > 199 * 1024) ro_warps >>= 1;
  bool small = true;                 // 32-bit element offsets inside every fab the kernel touches
  for (const lbx_mf* s : {(const lbx_mf*)dst, src0, src1, src1b})
    if (s && (double)s->max_cells(s->ngrow) * LBX_NV >= 2147483648.0) small = false;
  const bool use_rows = lbx::g_row_kernel && (plan || src_ghost) && dst->rows_ok && small && !lbx::g_debug_skip &&
                        (size_t)ro_warps * LBX_NV * ro_pitch * sizeof(double) <= 199 * 1024;
  lbx_resolved res;
  if (use_rows && plan && plan_resolved(plan, dst, src0, src1, &res)) return 1;
  const int pi = prof_open(0);
  const bool timed = pi >= 0;
  if (use_rows) {
    lbx::ROArgs a;
    memset(&a, 0, sizeof(a));
    a.vbase = reinterpret_cast<const double*>(src_valid->base);
    a.dbase = reinterpret_cast<double*>(dst->base);
    a.dt = dst->table;
    a.mt = mask ? mask->table : nullptr;
    a.gt = src_ghost ? src_ghost->table : nullptr;
    if (plan) {
      a.plan.tab = res.tab;
      a.plan.tab_first = res.first;
      a.plan.s0 = cp.s0; a.plan.s1 = cp.s1; a.plan.s1b = cp.s1b; a.plan.fb = cp.fb;
      a.plan.wa = wa; a.plan.wb = wb;
    }
    a.nfabs = dst->nlocal; a.warps = ro_warps; a.pitch = ro_pitch;
    a.omega_s = omega_s; a.omega_b = omega_b; a.fine_val = fine_val; a.zero_invalid = zero_invalid ? 1 : 0;
    if (L().mf_cs_rows(g.cur, a, level_step ? 3 : plan ? 2 : 1, dst->max_extent(1) + 4, dst->max_extent(2) + 4)) return fail("k_mf_cs_rows: cannot raise the shared-memory limit");
  } else {
    L().mf_collide_stream(g.cur, reinterpret_cast<const double*>(src_valid->base), reinterpret_cast<double*>(dst->base), dst->table,
                          mask ? mask->table : nullptr, src_ghost ? src_ghost->table : nullptr, cp, dst->nlocal, dst->max_extent(1),
                          dst->max_extent(2), dst->max_valid, ghost_tiles, omega_s, omega_b, fine_val,
                          (zero_invalid ? 1 : 0) | (level_step ? 2 : 0));
  }
  if (timed) {
    prof_close(pi);
    for (const auto& f : dst->host)
      if (f.local) g.prof_cells += (double)(f.vhi[0] - f.vlo[0] + 1) * (f.vhi[1] - f.vlo[1] + 1) * (f.vhi[2] - f.vlo[2] + 1);
  }
  if (lbx::after_launch(what)) return 1;
  return remote ? lbx::par_barrier() : 0;
}

}  // extern "C"
