// Library-internal context shared by the translation units of liblbx.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

namespace lbx {
struct Ctx {
  bool ready = false;
  int device = -1;
  cudaStream_t own = nullptr;      // the library's stream
  cudaStream_t cur = nullptr;      // stream kernels are queued on (own or external)
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool literal = false;
  int* peer_err = nullptr;         // host-mapped: set by k_peer_wait on timeout
  uint64_t launches = 0;
  // concurrent section (lbx_concurrent_begin/end): every launch goes to its own auxiliary stream
  static constexpr int NAUX = 4;
  cudaStream_t aux[NAUX] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[NAUX] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t main_saved = nullptr;
  // ranks of a distributed run (lbx_par_init): one process per GPU, peers reached through CUDA-IPC
  int rank = 0, world = 1;
  int (*allgather)(const void* send, size_t bytes, void* recv, void* user) = nullptr;
  void* allgather_user = nullptr;
  unsigned long long* bar_flags = nullptr;           // [world] slots written by the peers
  unsigned long long** d_bar_peers = nullptr;        // device array [world]: every rank's flag array
  unsigned long long bar_epoch = 0;
  uint64_t barriers = 0;
  // neighbour ordering of the distributed uniform path (lbx_mf_collide_stream_slab): four 8-byte words per rank
  // behind the barrier slots of bar_flags -- [0] written by the rank that owns the planes below this rank's
  // slab, [1] by the rank above, [2] the boundary-CTA counter of the step kernel, [3] spare
  static constexpr int STEP_WORDS = 4;
  unsigned long long* step_flags = nullptr;          // = bar_flags + step_off
  unsigned long long* peer_flag_base[32] = {};       // every rank's bar_flags (own pointer for this rank)
  int step_off = 0;
  unsigned long long step_epoch = 0;                 // steps queued so far (the same on every rank)
  // staged host <-> fab-set transfers (lbx_mf_from_user_host / lbx_mf_to_user_host): two device staging
  // buffers on two streams, so that the PCIe copy of one chunk overlaps the transposing kernel of the other
  static constexpr size_t STAGE_BYTES = size_t(64) << 20;
  cudaStream_t xfer[2] = {nullptr, nullptr};
  cudaEvent_t xfer_done[2] = {nullptr, nullptr}, xfer_fork = nullptr;
  void* stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;          // current size of each staging buffer (>= STAGE_BYTES, grown to one x-plane if larger)
  // live profiling of the AMR path's dominant kernel (lbx_prof_begin / lbx_prof_end): every
  // lbx_mf_collide_stream* launch is bracketed by CUDA events on the stream it is queued on
  bool prof = false;
  static constexpr int PROF_MAX = 8192;
  cudaEvent_t* prof_ev = nullptr;       // [2 * PROF_MAX], created on first use
  int prof_n = 0;
  unsigned char* prof_kind = nullptr;   // [PROF_MAX] what each bracket holds: 0 fused level pass, 1 gather plan, 2 average_down
  double prof_kind_ms[4] = {0, 0, 0, 0};   // last lbx_prof_end, per kind
  uint64_t prof_kind_n[4] = {0, 0, 0, 0};
  double prof_cells = 0.0;              // valid cells of this rank's boxes over the bracketed launches
  uint64_t prof_dropped = 0;
  int conc_next = -1;              // >= 0: inside a concurrent section; index of the stream in use
  int conc_used = 0;
};
extern Ctx g_ctx;
// Device-memory arena (the role AMReX's Arena plays under every MultiFab): freed blocks are kept
// and handed out again for requests of the same rounded size, so rebuilding a level or a whole
// simulation does not go back to cudaMalloc/cudaFree (milliseconds per GB).  Blocks are whole
// cudaMalloc allocations, never sub-allocated, so CUDA-IPC handles of arena memory stay valid.
// arena_free expects the caller to have drained the stream that last used the block.
cudaError_t arena_alloc(void** p, size_t bytes);
void arena_free(void* p);
void arena_release();                                    // cudaFree every cached block
void arena_stats(size_t* in_use, size_t* cached, uint64_t* hits, uint64_t* misses);
// distributed helpers (lbx_abi.cu)
int step_wait_launch(unsigned long long value);         // one-thread wait for both neighbours' step flags (lbx_abi.cu)
int par_barrier();                                       // device-side all-rank barrier on the current stream
int par_allgather(const void* send, size_t bytes, void* recv);   // host, through the registered callback
int ipc_open_cached(const unsigned char* handle, void** base);   // cudaIpcOpenMemHandle, once per handle
int fail(const std::string& msg);            // records the message for lbx_last_error(); returns 1
int after_launch(const char* what);          // counts the launch, reports launch errors
}  // namespace lbx

#define LBX_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return lbx::fail(std::string(#expr) + ": " + cudaGetErrorString(e_));              \
  } while (0)
#define LBX_NEED_INIT() \
  if (!lbx::g_ctx.ready) return lbx::fail("lbx: not initialised (call lbx_init / lambrexInit first)")
