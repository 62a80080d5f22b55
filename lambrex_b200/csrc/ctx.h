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
};
extern Ctx g_ctx;
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
