// C ABI of liblbx.so (declared in include/lbx.h).  Thin: argument checks,
// POD -> device-view conversion, one kernel launch per call.  No CPU fallback.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <string>
#include <vector>
#include <unordered_map>

#include "../../include/lbx.h"
#include "ctx.h"
#include "launch.h"
#include "kernels_util.cuh"

namespace lbx {
int g_smem_pad = 0;
int g_valid_linear = 0;
int g_align_rows = 0;
int g_xghost_in_row = 0;
int g_plain_stores = 0;
int g_debug_skip = 0;
int g_row_kernel = 1;
Ctx g_ctx;
thread_local std::string g_err;
int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
// ---- arena -------------------------------------------------------------------------------
namespace {
std::unordered_map<void*, size_t> a_live;        // block -> rounded size
std::multimap<size_t, void*> a_cache;            // rounded size -> free block
size_t a_in_use = 0, a_cached = 0;
uint64_t a_hits = 0, a_misses = 0;
size_t a_round(size_t b) {
  // 2 MiB pages for big blocks; in a distributed run EVERY block is a whole number of 2 MiB pages so
  // that it is an allocation of its own that a CUDA-IPC handle maps at offset 0
  const size_t g = (b >= (size_t(2) << 20) || g_ctx.world > 1) ? (size_t(2) << 20) : 512;
  return (b + g - 1) / g * g;
}
}  // namespace
void arena_release() {
  for (auto& kv : a_cache) cudaFree(kv.second);
  a_cache.clear();
  a_cached = 0;
}
cudaError_t arena_alloc(void** p, size_t bytes) {
  size_t rb = a_round(bytes ? bytes : 1);
  // best fit with bounded slack: after a regrid the levels' allocations change size by a few per
  // cent; reusing a slightly larger parked block saves the cudaMalloc and, in a distributed run,
  // the CUDA-IPC export + the peers' cudaIpcOpenMemHandle of a new multi-GB block (~0.1 s each)
  auto it = a_cache.lower_bound(rb);
  if (it != a_cache.end() && it->first <= rb + rb / 4 + (size_t(2) << 20)) {
    *p = it->second;
    rb = it->first;
    a_cache.erase(it);
    a_cached -= rb;
    ++a_hits;
  } else {
    cudaError_t e = cudaMalloc(p, rb);
    if (e != cudaSuccess) {          // give the cached blocks back to the driver and retry once
      cudaGetLastError();
      arena_release();
      e = cudaMalloc(p, rb);
      if (e != cudaSuccess) return e;
    }
    ++a_misses;
  }
  a_live[*p] = rb;
  a_in_use += rb;
  return cudaSuccess;
}
void arena_free(void* p) {
  if (!p) return;
  auto it = a_live.find(p);
  if (it == a_live.end()) { cudaFree(p); return; }     // not ours (defensive)
  const size_t rb = it->second;
  a_live.erase(it);
  a_in_use -= rb;
  a_cache.emplace(rb, p);
  a_cached += rb;
  // keep at most half of the device's memory parked in the cache: evict the largest blocks first
  // (not in a distributed run: peers may still have the block mapped)
  if (g_ctx.world > 1) return;
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return; }
  while (a_cached > tot / 2 && !a_cache.empty()) {
    auto last = std::prev(a_cache.end());
    cudaFree(last->second);
    a_cached -= last->first;
    a_cache.erase(last);
  }
}
void arena_stats(size_t* in_use, size_t* cached, uint64_t* hits, uint64_t* misses) {
  if (in_use) *in_use = a_in_use;
  if (cached) *cached = a_cached;
  if (hits) *hits = a_hits;
  if (misses) *misses = a_misses;
}

// ---- distributed helpers -------------------------------------------------------------------
namespace {
std::map<std::array<unsigned char, LBX_IPC_HANDLE_BYTES>, void*> ipc_opened;
}
int ipc_open_cached(const unsigned char* handle, void** base) {
  std::array<unsigned char, LBX_IPC_HANDLE_BYTES> key;
  memcpy(key.data(), handle, key.size());
  auto it = ipc_opened.find(key);
  if (it != ipc_opened.end()) { *base = it->second; return 0; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  ipc_opened[key] = *base;
  return 0;
}
int par_allgather(const void* send, size_t bytes, void* recv) {
  Ctx& g = g_ctx;
  if (g.world == 1) { memcpy(recv, send, bytes); return 0; }
  if (!g.allgather) return fail("lbx: distributed run without an allgather callback (lbx_par_init)");
  if (g.allgather(send, bytes, recv, g.allgather_user) != 0) return fail("lbx: allgather callback failed");
  return 0;
}
int par_barrier() {
  Ctx& g = g_ctx;
  if (g.world == 1) return 0;
  if (*g.peer_err) return fail("par_barrier: an earlier peer wait or barrier timed out; the run is no longer synchronised");
  int* derr = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&derr), g.peer_err, 0);
  if (e != cudaSuccess) return fail("par_barrier: cudaHostGetDevicePointer");
  ++g.bar_epoch;
  ++g.barriers;
  k_par_barrier<<<1, 32, 0, g.cur>>>(g.d_bar_peers, g.bar_flags, g.rank, g.world, g.bar_epoch, 30000000000ull, derr);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("par_barrier launch: ") + cudaGetErrorString(e));
  return 0;
}

int after_launch(const char* what) {
  ++g_ctx.launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string(what) + " launch: " + cudaGetErrorString(e));
  // STICKY peer error: once a peer wait / all-rank barrier has timed out, the work queued behind it ran on
  // unsynchronised peer data; every later call fails (host-mapped flag, no synchronisation needed) until
  // lbx_peer_error() acknowledges it, so a caller's step loop cannot run on silently
  if (g_ctx.peer_err && *g_ctx.peer_err)
    return fail(std::string(what) + ": a peer wait or all-rank barrier timed out earlier (a neighbour GPU did not signal); results since then are invalid");
  if (g_ctx.conc_next >= 0) {      // concurrent section: the next launch gets the next auxiliary stream
    Ctx& g = g_ctx;
    g.conc_next = (g.conc_next + 1) % Ctx::NAUX;
    if (g.conc_used < Ctx::NAUX) {
      ++g.conc_used;
      cudaStreamWaitEvent(g.aux[g.conc_next], g.fork_ev, 0);
    }
    g.cur = g.aux[g.conc_next];
  }
  return 0;
}
}  // namespace lbx

namespace {
using lbx::after_launch;
using lbx::fail;
using lbx::g_err;
lbx::Ctx& g = lbx::g_ctx;

int check_fab(const lbx_fab* f, int ncomp, int dtype, const char* what) {
  if (!f || !f->data) return fail(std::string(what) + ": null fab");
  if (f->ncomp < ncomp) return fail(std::string(what) + ": too few components");
  if (f->dtype != dtype) return fail(std::string(what) + ": wrong dtype");
  if (f->n[0] <= 0 || f->n[1] <= 0 || f->n[2] <= 0) return fail(std::string(what) + ": empty fab");
  return 0;
}
// every cell of box grown by `grow` (clipped by wrap in periodic directions) must exist in f
int check_cover(const lbx_fab* f, const lbx_box* b, int grow, const lbx_domain* dom, const char* what) {
  for (int d = 0; d < 3; ++d) {
    if (b->hi[d] < b->lo[d]) return fail(std::string(what) + ": empty box");
    int lo = b->lo[d] - grow, hi = b->hi[d] + grow;
    if (dom && dom->periodic[d]) {      // 1 periodic wrap, 2 walls: neighbours stay inside the domain either way
      // wrapped neighbours stay inside the domain
      if (lo < dom->lo[d]) lo = dom->lo[d];
      if (hi > dom->hi[d]) hi = dom->hi[d];
      if (dom->periodic[d] == 1 && grow && (b->lo[d] == dom->lo[d] || b->hi[d] == dom->hi[d])) {
        // wrap reaches the opposite side of the domain: fab must span it
        if (f->lo[d] > dom->lo[d] || f->lo[d] + f->n[d] - 1 < dom->hi[d])
          return fail(std::string(what) + ": periodic wrap needs a fab spanning the domain");
      }
    }
    if (lo < f->lo[d] || hi > f->lo[d] + f->n[d] - 1)
      return fail(std::string(what) + ": box (+stencil) not covered by fab");
  }
  if ((long long)(b->hi[1] - b->lo[1]) >= 65535 || (long long)(b->hi[2] - b->lo[2]) >= 65535)
    return fail(std::string(what) + ": box exceeds 65535 rows/planes per launch");
  return 0;
}

lbx::DFab dfab(const lbx_fab* f) {
  lbx::DFab d;
  d.p = static_cast<double*>(f->data);
  for (int a = 0; a < 3; ++a) { d.lo[a] = f->lo[a]; d.n[a] = f->n[a]; }
  return d;
}
lbx::DBox dbox(const lbx_box* b) {
  lbx::DBox d;
  for (int a = 0; a < 3; ++a) { d.lo[a] = b->lo[a]; d.hi[a] = b->hi[a]; }
  return d;
}
lbx::DDom ddom(const lbx_domain* b) {
  lbx::DDom d;
  for (int a = 0; a < 3; ++a) { d.lo[a] = b->lo[a]; d.hi[a] = b->hi[a]; d.periodic[a] = b->periodic[a]; }
  return d;
}
const lbx::Launchers& L() { return g.literal ? lbx::launchers_literal() : lbx::launchers_fast(); }

}  // namespace

extern "C" {

const char* lbx_last_error(void) { return g_err.c_str(); }

int lbx_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
  *count = n;
  return 0;
}

int lbx_init(int device) {
  if (g.ready) return 0;
  int n = 0;
  if (lbx_device_count(&n) || n == 0)
    return fail("lbx_init: no CUDA device available (this library has no CPU fallback)");
  if (device < 0) {
    const char* lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) % n : 0;
  }
  if (device >= n) return fail("lbx_init: device index out of range");
  LBX_CUDA(cudaSetDevice(device));
  LBX_CUDA(cudaStreamCreateWithFlags(&g.own, cudaStreamNonBlocking));
  LBX_CUDA(cudaEventCreate(&g.t0));
  LBX_CUDA(cudaEventCreate(&g.t1));
  LBX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g.peer_err), sizeof(int), cudaHostAllocMapped));
  *g.peer_err = 0;
  for (int a = 0; a < lbx::Ctx::NAUX; ++a) {
    LBX_CUDA(cudaStreamCreateWithFlags(&g.aux[a], cudaStreamNonBlocking));
    LBX_CUDA(cudaEventCreateWithFlags(&g.join_ev[a], cudaEventDisableTiming));
  }
  LBX_CUDA(cudaEventCreateWithFlags(&g.fork_ev, cudaEventDisableTiming));
  g.cur = g.own;
  g.device = device;
  g.ready = true;
  // LBX_COLLIDE=literal: start in the reference's operation order without FMA contraction (LBX_OPT_COLLIDE_LITERAL)
  // -- for callers that cannot call lbx_set_option, e.g. the reference's own unmodified test binaries, whose
  // golden velocities carry 1e-18 round-off noise that only bit-compatible arithmetic reproduces
  if (const char* c = getenv("LBX_COLLIDE")) {
    if (!strcmp(c, "literal")) g.literal = true;
    else if (!strcmp(c, "fast")) g.literal = false;
    else return fail("lbx_init: LBX_COLLIDE must be 'literal' or 'fast'");
  }
  return 0;
}

/* Launches made between begin and end are independent of one another (the caller's promise):
 * each goes to its own auxiliary stream, ordered after everything queued before begin; end
 * orders everything queued after it behind all of them.  Small latency-bound kernels (ghost-cell
 * pushes, gather plans) then overlap with the bandwidth-bound pass over the valid cells. */
int lbx_concurrent_end(void);
int lbx_concurrent_begin(void) {
  LBX_NEED_INIT();
  if (g.conc_next >= 0 && lbx_concurrent_end()) return 1;    // a section left open by a failed call: close it
  LBX_CUDA(cudaEventRecord(g.fork_ev, g.cur));
  g.main_saved = g.cur;
  g.conc_next = 0;
  g.conc_used = 1;
  LBX_CUDA(cudaStreamWaitEvent(g.aux[0], g.fork_ev, 0));
  g.cur = g.aux[0];
  return 0;
}
int lbx_concurrent_end(void) {
  LBX_NEED_INIT();
  if (g.conc_next < 0) return fail("lbx_concurrent_end: not inside a concurrent section");
  g.cur = g.main_saved;
  for (int a = 0; a < g.conc_used; ++a) {
    LBX_CUDA(cudaEventRecord(g.join_ev[a], g.aux[a]));
    LBX_CUDA(cudaStreamWaitEvent(g.cur, g.join_ev[a], 0));
  }
  g.conc_next = -1;
  g.conc_used = 0;
  return 0;
}

int lbx_finalize(void) {
  if (!g.ready) return 0;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.own);
  if (g.world > 1) {
    // every rank stops using its peers' memory before anyone unmaps or frees it
    lbx::par_barrier();
    cudaStreamSynchronize(g.cur);
    for (auto& kv : lbx::ipc_opened) cudaIpcCloseMemHandle(kv.second);
    lbx::ipc_opened.clear();
    unsigned char tok = 0;
    std::vector<unsigned char> toks(g.world);
    lbx::par_allgather(&tok, 1, toks.data());
  }
  for (int a = 0; a < 2; ++a) {
    if (g.xfer[a]) { cudaStreamSynchronize(g.xfer[a]); cudaStreamDestroy(g.xfer[a]); }
    if (g.xfer_done[a]) cudaEventDestroy(g.xfer_done[a]);
    if (g.stage[a]) cudaFree(g.stage[a]);
  }
  if (g.xfer_fork) cudaEventDestroy(g.xfer_fork);
  for (int a = 0; a < lbx::Ctx::NAUX; ++a) {
    if (g.aux[a]) { cudaStreamSynchronize(g.aux[a]); cudaStreamDestroy(g.aux[a]); }
    if (g.join_ev[a]) cudaEventDestroy(g.join_ev[a]);
  }
  if (g.fork_ev) cudaEventDestroy(g.fork_ev);
  lbx::arena_release();
  if (g.prof_ev) {
    for (int i = 0; i < 2 * lbx::Ctx::PROF_MAX; ++i) cudaEventDestroy(g.prof_ev[i]);
    delete[] g.prof_ev;
    delete[] g.prof_kind;
    g.prof_kind = nullptr;
  }
  cudaEventDestroy(g.t0);
  cudaEventDestroy(g.t1);
  cudaStreamDestroy(g.own);
  if (g.peer_err) cudaFreeHost(g.peer_err);
  g = lbx::Ctx();
  return 0;
}

int lbx_initialized(void) { return g.ready ? 1 : 0; }

/* One process per GPU.  `allgather(send, bytes, recv, user)` must gather `bytes` from every rank
 * into recv (rank order) and return 0 -- the host plumbing (torch.distributed / MPI); it is used
 * only to exchange CUDA-IPC handles when device fields are created.  Collective. */
int lbx_par_init(int rank, int world, int (*allgather)(const void*, size_t, void*, void*), void* user) {
  LBX_NEED_INIT();
  if (world < 1 || rank < 0 || rank >= world || world > 32) return fail("lbx_par_init: bad rank/world (1..32 ranks)");
  if (g.world > 1) return fail("lbx_par_init: already initialised");
  if (world == 1) return 0;
  if (!allgather) return fail("lbx_par_init: null allgather callback");
  g.rank = rank; g.world = world; g.allgather = allgather; g.allgather_user = user;
  g.step_off = (world + 3) / 4 * 4;                  // step-ordering words behind the barrier slots, 32 B aligned
  const size_t flag_words = (size_t)g.step_off + lbx::Ctx::STEP_WORDS;
  LBX_CUDA(lbx::arena_alloc(reinterpret_cast<void**>(&g.bar_flags), sizeof(unsigned long long) * flag_words));
  LBX_CUDA(cudaMemset(g.bar_flags, 0, sizeof(unsigned long long) * flag_words));
  g.step_flags = g.bar_flags + g.step_off;
  g.step_epoch = 0;
  cudaIpcMemHandle_t mine;
  LBX_CUDA(cudaIpcGetMemHandle(&mine, g.bar_flags));
  std::vector<unsigned char> all((size_t)world * LBX_IPC_HANDLE_BYTES);
  if (lbx::par_allgather(&mine, LBX_IPC_HANDLE_BYTES, all.data())) return 1;
  std::vector<unsigned long long*> peers(world);
  for (int r = 0; r < world; ++r) {
    if (r == rank) { peers[r] = g.bar_flags; continue; }
    void* base = nullptr;
    if (lbx::ipc_open_cached(all.data() + (size_t)r * LBX_IPC_HANDLE_BYTES, &base)) return 1;
    peers[r] = static_cast<unsigned long long*>(base);
  }
  for (int r = 0; r < world; ++r) g.peer_flag_base[r] = peers[r];
  LBX_CUDA(lbx::arena_alloc(reinterpret_cast<void**>(&g.d_bar_peers), sizeof(void*) * world));
  LBX_CUDA(cudaMemcpy(g.d_bar_peers, peers.data(), sizeof(void*) * world, cudaMemcpyHostToDevice));
  // nobody may signal before every rank has zeroed and published its flags
  unsigned char tok = 0;
  std::vector<unsigned char> toks(world);
  return lbx::par_allgather(&tok, 1, toks.data());
}
/* Host-only variant for the grid-generation metadata (no CUDA device needed): registers rank, world and the
 * allgather callback so that lbx_par_allgather works -- what the distributed regrid (tag runs merged across
 * ranks) needs.  No device barrier flags are set up: device collectives stay unavailable. */
int lbx_par_init_host(int rank, int world, int (*allgather)(const void*, size_t, void*, void*), void* user) {
  if (g.ready) return fail("lbx_par_init_host: a device context exists; use lbx_par_init");
  if (world < 1 || rank < 0 || rank >= world || world > 32) return fail("lbx_par_init_host: bad rank/world (1..32 ranks)");
  if (world > 1 && !allgather) return fail("lbx_par_init_host: null allgather callback");
  g.rank = rank; g.world = world; g.allgather = world > 1 ? allgather : nullptr; g.allgather_user = user;
  return 0;
}
int lbx_par_finalize_host(void) {
  if (g.ready) return fail("lbx_par_finalize_host: a device context exists; use lbx_finalize");
  g.rank = 0; g.world = 1; g.allgather = nullptr; g.allgather_user = nullptr;
  return 0;
}
int lbx_par_info(int* rank, int* world, uint64_t* barriers) {
  if (rank) *rank = g.rank;
  if (world) *world = g.world;
  if (barriers) *barriers = g.barriers;
  return 0;
}
int lbx_par_allgather(const void* send, size_t bytes, void* recv) {
  if (!g.ready && !g.allgather && g.world > 1) return fail("lbx_par_allgather: not initialised");
  return lbx::par_allgather(send, bytes, recv);
}
int lbx_par_barrier(void) {
  LBX_NEED_INIT();
  return lbx::par_barrier();
}

int lbx_device_info(char* name, int name_cap, int* sm_count, size_t* total_bytes, size_t* free_bytes) {
  LBX_NEED_INIT();
  cudaDeviceProp p;
  LBX_CUDA(cudaGetDeviceProperties(&p, g.device));
  if (name && name_cap > 0) { strncpy(name, p.name, name_cap - 1); name[name_cap - 1] = 0; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  size_t fr = 0, tot = 0;
  LBX_CUDA(cudaMemGetInfo(&fr, &tot));
  if (total_bytes) *total_bytes = tot;
  if (free_bytes) *free_bytes = fr;
  return 0;
}

int lbx_set_option(int key, int value) {
  switch (key) {
    case LBX_OPT_COLLIDE_LITERAL: g.literal = (value != 0); return 0;
    case LBX_OPT_SMEM_PAD:
      if (value < 0 || value > 48 * 1024) return fail("lbx_set_option: LBX_OPT_SMEM_PAD must be 0..49152");
      lbx::g_smem_pad = value;
      return 0;
    case LBX_OPT_XGHOST_IN_ROW: lbx::g_xghost_in_row = (value != 0); return 0;
    case LBX_OPT_PLAIN_STORES: lbx::g_plain_stores = (value != 0); return 0;
    case LBX_OPT_ALIGN_ROWS:
      if (value != 0 && value != 1 && value != 4 && value != 8 && value != 16) return fail("lbx_set_option: LBX_OPT_ALIGN_ROWS takes 0, 4, 8 or 16 (doubles)");
      lbx::g_align_rows = value == 1 ? 4 : value;
      return 0;
    case LBX_OPT_VALID_TILING: lbx::g_valid_linear = (value != 0); return 0;
    case LBX_OPT_DEBUG_SKIP: lbx::g_debug_skip = value & 3; return 0;
    case LBX_OPT_ROW_KERNEL: lbx::g_row_kernel = (value != 0); return 0;
    default: return fail("lbx_set_option: unknown key");
  }
}

int lbx_set_stream(void* s) {
  LBX_NEED_INIT();
  // NULL: the library's own stream; pass cudaStreamLegacy (0x1) / cudaStreamPerThread (0x2) for the default streams
  g.cur = s ? static_cast<cudaStream_t>(s) : g.own;
  return 0;
}

int lbx_sync(void) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaStreamSynchronize(g.cur));
  if (*g.peer_err) return fail("lbx_sync: a peer wait timed out (neighbour GPU did not signal)");
  return 0;
}

uint64_t lbx_launch_count(void) { return g.launches; }

int lbx_malloc(void** p, size_t bytes) {
  LBX_NEED_INIT();
  LBX_CUDA(lbx::arena_alloc(p, bytes));
  return 0;
}
int lbx_free(void* p) {
  if (!p) return 0;
  LBX_NEED_INIT();
  LBX_CUDA(cudaStreamSynchronize(g.cur));
  lbx::arena_free(p);
  return 0;
}
int lbx_arena_release(void) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaStreamSynchronize(g.cur));
  lbx::arena_release();
  return 0;
}
int lbx_arena_info(size_t* in_use_bytes, size_t* cached_bytes, uint64_t* hits, uint64_t* misses) {
  lbx::arena_stats(in_use_bytes, cached_bytes, hits, misses);
  return 0;
}
int lbx_memset(void* p, int byte, size_t bytes) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaMemsetAsync(p, byte, bytes, g.cur));
  return 0;
}
int lbx_host_alloc(void** p, size_t bytes) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
int lbx_host_free(void* p) {
  if (!p) return 0;
  LBX_CUDA(cudaFreeHost(p));
  return 0;
}
int lbx_h2d(void* d, const void* h, size_t bytes) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, g.cur));
  return 0;
}
int lbx_d2h(void* h, const void* d, size_t bytes) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, g.cur));
  return 0;
}
int lbx_d2d(void* dst, const void* src, size_t bytes) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g.cur));
  return 0;
}

int lbx_timer_start(void) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaEventRecord(g.t0, g.cur));
  return 0;
}
int lbx_timer_stop(float* ms) {
  LBX_NEED_INIT();
  LBX_CUDA(cudaEventRecord(g.t1, g.cur));
  LBX_CUDA(cudaEventSynchronize(g.t1));
  LBX_CUDA(cudaEventElapsedTime(ms, g.t0, g.t1));
  return 0;
}

int lbx_equilibrium(const lbx_fab* f, const lbx_fab* rho, const lbx_fab* u, const lbx_box* box) {
  LBX_NEED_INIT();
  if (check_fab(f, LBX_NV, LBX_F64, "lbx_equilibrium f") || check_fab(rho, 1, LBX_F64, "lbx_equilibrium rho") ||
      check_fab(u, 3, LBX_F64, "lbx_equilibrium u"))
    return 1;
  if (check_cover(f, box, 0, nullptr, "lbx_equilibrium f") || check_cover(rho, box, 0, nullptr, "lbx_equilibrium rho") ||
      check_cover(u, box, 0, nullptr, "lbx_equilibrium u"))
    return 1;
  L().equilibrium(g.cur, dfab(f), dfab(rho), dfab(u), dbox(box));
  return after_launch("lbx_equilibrium");
}

int lbx_moments(const lbx_fab* f, const lbx_fab* rho, const lbx_fab* u, const lbx_box* box) {
  LBX_NEED_INIT();
  if (check_fab(f, LBX_NV, LBX_F64, "lbx_moments f") || check_fab(rho, 1, LBX_F64, "lbx_moments rho") ||
      check_fab(u, 3, LBX_F64, "lbx_moments u"))
    return 1;
  if (check_cover(f, box, 0, nullptr, "lbx_moments f") || check_cover(rho, box, 0, nullptr, "lbx_moments rho") ||
      check_cover(u, box, 0, nullptr, "lbx_moments u"))
    return 1;
  L().moments(g.cur, dfab(f), dfab(rho), dfab(u), dbox(box));
  return after_launch("lbx_moments");
}

int lbx_collide(const lbx_fab* src, const lbx_fab* dst, const lbx_box* box, double omega_s, double omega_b,
                const lbx_fab* mask, int fine_val) {
  LBX_NEED_INIT();
  if (check_fab(src, LBX_NV, LBX_F64, "lbx_collide src") || check_fab(dst, LBX_NV, LBX_F64, "lbx_collide dst")) return 1;
  if (check_cover(src, box, 0, nullptr, "lbx_collide src") || check_cover(dst, box, 0, nullptr, "lbx_collide dst")) return 1;
  lbx::DMask m;
  memset(&m, 0, sizeof(m));
  if (mask) {
    if (check_fab(mask, 1, LBX_I32, "lbx_collide mask") || check_cover(mask, box, 0, nullptr, "lbx_collide mask")) return 1;
    m.p = static_cast<const int*>(mask->data);
    for (int a = 0; a < 3; ++a) { m.lo[a] = mask->lo[a]; m.n[a] = mask->n[a]; }
  }
  L().collide(g.cur, dfab(src), dfab(dst), dbox(box), omega_s, omega_b, m, fine_val);
  return after_launch("lbx_collide");
}

int lbx_stream(const lbx_fab* src, const lbx_fab* dst, const lbx_box* box, const lbx_domain* dom) {
  LBX_NEED_INIT();
  if (!dom) return fail("lbx_stream: null domain");
  if (check_fab(src, LBX_NV, LBX_F64, "lbx_stream src") || check_fab(dst, LBX_NV, LBX_F64, "lbx_stream dst")) return 1;
  if (src->data == dst->data) return fail("lbx_stream: src and dst must not alias");
  if (check_cover(src, box, 1, dom, "lbx_stream src") || check_cover(dst, box, 0, nullptr, "lbx_stream dst")) return 1;
  L().stream(g.cur, dfab(src), dfab(dst), dbox(box), ddom(dom));
  return after_launch("lbx_stream");
}

int lbx_collide_stream(const lbx_fab* src, const lbx_fab* dst, const lbx_box* box, const lbx_domain* dom,
                       double omega_s, double omega_b, int scheme) {
  LBX_NEED_INIT();
  if (!dom) return fail("lbx_collide_stream: null domain");
  if (scheme != LBX_PUSH && scheme != LBX_PULL) return fail("lbx_collide_stream: unknown scheme");
  for (int d = 0; d < 3; ++d) {
    if (dom->periodic[d] < 0 || dom->periodic[d] > 2) return fail("lbx_collide_stream: periodic[] takes 0 (ghost cells), 1 (periodic) or 2 (walls)");
    if (dom->periodic[d] == 2 && scheme != LBX_PUSH) return fail("lbx_collide_stream: walls need the push scheme");
  }
  if (check_fab(src, LBX_NV, LBX_F64, "lbx_collide_stream src") ||
      check_fab(dst, LBX_NV, LBX_F64, "lbx_collide_stream dst"))
    return 1;
  if (src->data == dst->data) return fail("lbx_collide_stream: src and dst must not alias");
  const bool push = (scheme == LBX_PUSH);
  if (check_cover(src, box, push ? 0 : 1, push ? nullptr : dom, "lbx_collide_stream src") ||
      check_cover(dst, box, push ? 1 : 0, push ? dom : nullptr, "lbx_collide_stream dst"))
    return 1;
  L().collide_stream(g.cur, dfab(src), dfab(dst), dbox(box), ddom(dom), omega_s, omega_b, scheme);
  return after_launch("lbx_collide_stream");
}

int lbx_ipc_get_handle(void* p, unsigned char* handle) {
  LBX_NEED_INIT();
  static_assert(sizeof(cudaIpcMemHandle_t) == LBX_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  LBX_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle, &h, sizeof(h));
  return 0;
}
int lbx_ipc_open_handle(const unsigned char* handle, void** p) {
  LBX_NEED_INIT();
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  LBX_CUDA(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int lbx_ipc_close_handle(void* p) {
  if (!p) return 0;
  LBX_NEED_INIT();
  LBX_CUDA(cudaStreamSynchronize(g.cur));
  LBX_CUDA(cudaIpcCloseMemHandle(p));
  return 0;
}

int lbx_peer_signal(uint64_t* a, uint64_t* b, uint64_t value) {
  LBX_NEED_INIT();
  lbx::k_peer_signal<<<1, 1, 0, g.cur>>>(reinterpret_cast<unsigned long long*>(a),
                                         reinterpret_cast<unsigned long long*>(b), value);
  return after_launch("lbx_peer_signal");
}
int lbx_peer_wait(const uint64_t* a, const uint64_t* b, uint64_t value, uint64_t timeout_ns) {
  LBX_NEED_INIT();
  int* derr = nullptr;
  LBX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&derr), g.peer_err, 0));
  lbx::k_peer_wait<<<1, 1, 0, g.cur>>>(reinterpret_cast<const unsigned long long*>(a),
                                       reinterpret_cast<const unsigned long long*>(b), value, timeout_ns, derr);
  return after_launch("lbx_peer_wait");
}
}  // extern "C"
// queue a one-thread wait until both neighbours have published `value` (lbx_mf_collide_stream_slab, lbx_mf.cu)
int lbx::step_wait_launch(unsigned long long value) {
  int* derr = nullptr;
  LBX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&derr), g.peer_err, 0));
  lbx::k_peer_wait<<<1, 1, 0, g.cur>>>(g.step_flags, g.step_flags + 1, value, 30000000000ull, derr);
  return after_launch("lbx_mf_collide_stream_slab (neighbour wait)");
}
extern "C" {
int lbx_par_step_finish(void) {
  LBX_NEED_INIT();
  if (g.world == 1) return 0;
  int* derr = nullptr;
  LBX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&derr), g.peer_err, 0));
  lbx::k_peer_wait<<<1, 1, 0, g.cur>>>(g.step_flags, g.step_flags + 1, g.step_epoch, 30000000000ull, derr);
  return after_launch("lbx_par_step_finish");
}
int lbx_peer_error(void) {   // returns the flag and clears it
  if (!g.ready || !g.peer_err) return 0;
  const int e = *g.peer_err;
  *g.peer_err = 0;
  return e;
}

int lbx_collide_stream_slab(const lbx_fab* src, const lbx_fab* dst, const lbx_fab* dn, const lbx_fab* up,
                            const lbx_box* box, const lbx_domain* dom, double omega_s, double omega_b) {
  LBX_NEED_INIT();
  if (!dom) return fail("lbx_collide_stream_slab: null domain");
  if (check_fab(src, LBX_NV, LBX_F64, "lbx_collide_stream_slab src") ||
      check_fab(dst, LBX_NV, LBX_F64, "lbx_collide_stream_slab dst") ||
      check_fab(dn, LBX_NV, LBX_F64, "lbx_collide_stream_slab dst_dn") ||
      check_fab(up, LBX_NV, LBX_F64, "lbx_collide_stream_slab dst_up"))
    return 1;
  if (src->data == dst->data) return fail("lbx_collide_stream_slab: src and dst must not alias");
  if (check_cover(src, box, 0, nullptr, "lbx_collide_stream_slab src")) return 1;
  // x,y: the destination must hold box grown by 1 (wrapped where periodic); z: every plane
  // k+-1 must exist in dst or in the neighbour given for that side
  lbx_box xy = *box;
  xy.lo[2] = xy.hi[2] = box->lo[2];
  for (const lbx_fab* f : {dst, dn, up}) {
    lbx_box b = xy;
    b.lo[2] = b.hi[2] = f->lo[2];
    lbx_fab flat = *f;           // check x,y coverage only
    flat.n[2] = 1;
    if (check_cover(&flat, &b, 0, nullptr, "lbx_collide_stream_slab dst") ) return 1;
    for (int d = 0; d < 2; ++d) {
      const int lo = box->lo[d] - 1, hi = box->hi[d] + 1;
      const bool wrap = dom->periodic[d] && box->lo[d] == dom->lo[d] && box->hi[d] == dom->hi[d];
      if (!wrap && (lo < f->lo[d] || hi > f->lo[d] + f->n[d] - 1))
        return fail("lbx_collide_stream_slab: destination fab lacks the x/y stencil rim");
    }
  }
  auto has = [](const lbx_fab* f, int k) { return k >= f->lo[2] && k < f->lo[2] + f->n[2]; };
  if (!has(dst, box->lo[2]) || !has(dst, box->hi[2])) return fail("lbx_collide_stream_slab: box not inside dst in z");
  int kp = box->hi[2] + 1, km = box->lo[2] - 1;
  if (dom->periodic[2]) { if (kp > dom->hi[2]) kp = dom->lo[2]; if (km < dom->lo[2]) km = dom->hi[2]; }
  if (!has(dst, kp) && !has(up, kp)) return fail("lbx_collide_stream_slab: plane above the slab is in neither dst nor dst_up");
  if (!has(dst, km) && !has(dn, km)) return fail("lbx_collide_stream_slab: plane below the slab is in neither dst nor dst_dn");
  L().collide_stream_slab(g.cur, dfab(src), dfab(dst), dfab(dn), dfab(up), dbox(box), ddom(dom), omega_s, omega_b);
  return after_launch("lbx_collide_stream_slab");
}

static int halo_common(const lbx_fab* f, const lbx_box* reg, int face, const void* buf, const char* what) {
  if (check_fab(f, LBX_NV, LBX_F64, what) || check_cover(f, reg, 0, nullptr, what)) return 1;
  if (face < 0 || face > 5) return fail(std::string(what) + ": face must be 0..5");
  if (!buf) return fail(std::string(what) + ": null buffer");
  return 0;
}
int lbx_halo_pack(const lbx_fab* f, const lbx_box* reg, int face, double* buf) {
  LBX_NEED_INIT();
  if (halo_common(f, reg, face, buf, "lbx_halo_pack")) return 1;
  lbx::k_halo<true><<<lbx::grid_for(dbox(reg)), lbx::BX, 0, g.cur>>>(dfab(f), dbox(reg), face, buf);
  return after_launch("lbx_halo_pack");
}
int lbx_halo_unpack(const lbx_fab* f, const lbx_box* reg, int face, const double* buf) {
  LBX_NEED_INIT();
  if (halo_common(f, reg, face, buf, "lbx_halo_unpack")) return 1;
  lbx::k_halo<false><<<lbx::grid_for(dbox(reg)), lbx::BX, 0, g.cur>>>(dfab(f), dbox(reg), face,
                                                                     const_cast<double*>(buf));
  return after_launch("lbx_halo_unpack");
}

void lbx_d3q15_tables(double* M, double* Minv, int32_t* c, double* w) {
  for (int m = 0; m < LBX_NV; ++m)
    for (int p = 0; p < LBX_NV; ++p) {
      M[m * LBX_NV + p] = lbx::mode_entry(m, p);
      Minv[p * LBX_NV + m] = lbx::inv_entry(p, m);
    }
  for (int p = 0; p < LBX_NV; ++p) {
    c[3 * p + 0] = lbx::cx(p);
    c[3 * p + 1] = lbx::cy(p);
    c[3 * p + 2] = lbx::cz(p);
    w[p] = (double)lbx::w_num(p) / (double)lbx::w_den(p);
  }
}

}  // extern "C"
