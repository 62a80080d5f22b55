// sector_bench -- stand-alone micro-benchmark for round 2 (profiles/r01_alignment.md): what does it cost on a
// B200 when a streaming pass writes a destination made of short rows with ghost cells?
//
// Layout: R rows of `pitch` doubles; cells [off, off + n) of every row are "valid", the others are "ghost".
// A pass copies src -> dst (one load and one store per valid cell, the shape of the collide-stream kernels
// without the arithmetic) and treats the ghost cells of the DESTINATION in one of five ways:
//   0 none     : ghost cells are never written (sectors shared with valid cells stay partially written)
//   1 samewarp : the warp that wrote a row's valid cells zeroes the row's ghost cells right after
//   2 samecta  : the last warp of the CTA zeroes the ghost cells of all rows of the CTA
//   3 latecta  : extra CTAs at the END of the grid zero the ghost cells (the boxed kernel's ghost tiles)
//   4 kernel2  : a second kernel zeroes them
// plus the reference point `tight` (pitch == n, off == 0: no ghost cells at all).
// Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sector_bench tools/sector_bench.cu
// Run :   ./sector_bench [n=32] [ghost=2] [lead=0] [mb=4096]     (prints GB/s of valid-cell bytes, read + write)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)

constexpr int ROWS_PER_CTA = 8;     // 8 warps, one row each (n <= 32) -- the boxed kernel's valid tile

// mode 0..2: grid = rows / 8;  mode 3: grid = rows / 8 + ghost CTAs
__global__ void __launch_bounds__(256) k_pass(const double* __restrict__ src, double* __restrict__ dst, long long rows, int n, int pitch,
                                              int off, int mode, long long valid_ctas) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if ((long long)blockIdx.x < valid_ctas) {
    const long long r = (long long)blockIdx.x * ROWS_PER_CTA + w;
    if (r < rows) {
      const long long base = r * pitch;
      for (int x = lane; x < n; x += 32) __stcs(dst + base + off + x, __ldcs(src + base + off + x) + 1.0);
      if (mode == 1) {                                         // ghost cells of this row, same warp
        const int g = pitch - n;
        if (lane < g) { const int x = lane < off ? lane : n + lane; dst[base + x] = 0.0; }
      }
    }
    if (mode == 2) {
      __syncthreads();
      const int g = pitch - n;                                 // ghost cells of the CTA's 8 rows, by the last warp
      if (w == ROWS_PER_CTA - 1)
        for (int q = lane; q < g * ROWS_PER_CTA; q += 32) {
          const long long rr = (long long)blockIdx.x * ROWS_PER_CTA + q / g;
          const int c = q % g, x = c < off ? c : n + c;
          if (rr < rows) dst[rr * pitch + x] = 0.0;
        }
    }
    return;
  }
  // mode 3: ghost CTAs after all valid CTAs; one thread per ghost cell
  const int g = pitch - n;
  const long long q = ((long long)blockIdx.x - valid_ctas) * 256 + threadIdx.x;
  if (q < rows * g) {
    const long long rr = q / g;
    const int c = (int)(q % g), x = c < off ? c : n + c;
    dst[rr * pitch + x] = 0.0;
  }
}

__global__ void __launch_bounds__(256) k_ghost(double* __restrict__ dst, long long rows, int n, int pitch, int off) {
  const int g = pitch - n;
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  if (q < rows * g) {
    const long long rr = q / g;
    const int c = (int)(q % g), x = c < off ? c : n + c;
    dst[rr * pitch + x] = 0.0;
  }
}

static double run(const double* src, double* dst, long long rows, int n, int pitch, int off, int mode, int reps) {
  const long long vc = (rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  const long long gc = (mode == 3) ? (rows * (pitch - n) + 255) / 256 : 0;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int it = 0; it < reps + 3; ++it) {
    if (it == 3) CK(cudaEventRecord(a));
    k_pass<<<(unsigned)(vc + gc), 256>>>(src, dst, rows, n, pitch, off, mode, vc);
    if (mode == 4 && pitch > n) k_ghost<<<(unsigned)((rows * (pitch - n) + 255) / 256), 256>>>(dst, rows, n, pitch, off);
  }
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaEventDestroy(a));
  CK(cudaEventDestroy(b));
  return 16.0 * (double)rows * n * reps / (ms * 1e-3) / 1e9;      // valid-cell bytes, read + write
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 32, ghost = argc > 2 ? std::atoi(argv[2]) : 2,
            lead = argc > 3 ? std::atoi(argv[3]) : 0;
  const long long mb = argc > 4 ? std::atoll(argv[4]) : 4096;
  if (n < 1 || ghost < 0 || lead < 0) { std::printf("bad arguments\n"); return 1; }
  const int pitch = lead + ghost + n + ghost, off = lead + ghost;
  const long long rows = mb * 1024 * 1024 / 2 / (8LL * pitch);      // two buffers of mb/2 MiB
  double *src = nullptr, *dst = nullptr;
  CK(cudaMalloc(&src, sizeof(double) * rows * pitch));
  CK(cudaMalloc(&dst, sizeof(double) * rows * pitch));
  CK(cudaMemset(src, 0, sizeof(double) * rows * pitch));
  CK(cudaMemset(dst, 0, sizeof(double) * rows * pitch));
  std::printf("rows of %d valid doubles, %d ghost cells each side, lead-in %d: pitch %d doubles (%d B), valid offset %d B, %lld rows\n",
              n, ghost, lead, pitch, pitch * 8, off * 8, rows);
  const long long trows = mb * 1024 * 1024 / 2 / (8LL * n);
  std::printf("  tight (no ghost cells, pitch = n)      : %8.1f GB/s\n", run(src, dst, trows, n, n, 0, 0, 20));
  const char* names[5] = {"ghost cells never written             ", "ghost cells by the row's own warp      ",
                          "ghost cells by the last warp of the CTA", "ghost cells by CTAs at the grid's end  ",
                          "ghost cells by a second kernel         "};
  for (int mode = 0; mode < 5; ++mode)
    std::printf("  %s: %8.1f GB/s\n", names[mode], run(src, dst, rows, n, pitch, off, mode, 20));
  CK(cudaFree(src));
  CK(cudaFree(dst));
  return 0;
}
