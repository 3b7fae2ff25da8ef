// Diagnostics library (libgpar_b200_debug.so; include/gpar_b200_debug.h): raw fp64 issue-rate probes (DMMA m8n8k4,
// DFMA) and single-warp latency probes.  Not linked into the product library.
#include <cstdarg>
#include <cstdio>

#include "../../include/gpar_b200_debug.h"
#include "common.cuh"

namespace gpar {
// common.cuh declares these; the debug library carries its own minimal copies (it is a separate shared object)
static thread_local char g_dbg_err[256] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_dbg_err, sizeof(g_dbg_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -1000 - (int)e;
  }
  return 0;
}

// ---- fp64 issue-rate probes -------------------------------------------------------------
__global__ void __launch_bounds__(256) probe_dmma_kernel(int64_t iters, double* sink) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(int64_t iters, double* sink) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-12;
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) sink[0] = s;
}

// Single-warp dependent-chain latency probes (cycles per op), for designing the Cholesky critical path.
__global__ void probe_latency_kernel(double* out, double seed) {
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = seed + lane; sm[lane + 32] = 1.0;
  __syncthreads();
  const int N = 256;
  long long t0, t1;
  double r[10];
  // 0: dependent DMMA chain (one accumulator)
  { double c0 = 0, c1 = 0, a = seed, b = 1.0; t0 = clock64();
    for (int i = 0; i < N; ++i) dmma884(c0, c1, a, b);
    t1 = clock64(); r[0] = (double)(t1 - t0) / N; sm[0] += c0 + c1; }
  // 1: 4 accumulators round-robin (k-step structure of strip_mma): per k-step
  { double c[4][2] = {{0,0},{0,0},{0,0},{0,0}}; double a = seed, b = 1.0; t0 = clock64();
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int t = 0; t < 4; ++t) dmma884(c[t][0], c[t][1], a, b); }
    t1 = clock64(); r[1] = (double)(t1 - t0) / N; sm[1] += c[0][0] + c[1][0] + c[2][1] + c[3][1]; }
  // 2: dependent DFMA chain
  { double x = seed; t0 = clock64();
    for (int i = 0; i < N; ++i) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64(); r[2] = (double)(t1 - t0) / N; sm[2] += x; }
  // 3: double shuffle chain
  { double x = seed + lane; t0 = clock64();
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64(); r[3] = (double)(t1 - t0) / N; sm[3] += x; }
  // 4: rsqrt chain
  { double x = seed + 2.0; t0 = clock64();
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
    t1 = clock64(); r[4] = (double)(t1 - t0) / N; sm[4] += x; }
  // 5: sqrt chain
  { double x = seed + 2.0; t0 = clock64();
    for (int i = 0; i < N; ++i) x = sqrt(x) + 1.5;
    t1 = clock64(); r[5] = (double)(t1 - t0) / N; sm[5] += x; }
  // 6: division chain
  { double x = seed + 2.0; t0 = clock64();
    for (int i = 0; i < N; ++i) x = 3.0 / x + 1.5;
    t1 = clock64(); r[6] = (double)(t1 - t0) / N; sm[6] += x; }
  // 7: shared-memory store -> syncwarp -> load round trip chain
  { double x = seed; t0 = clock64();
    for (int i = 0; i < N; ++i) { sm[32 + lane] = x; __syncwarp(); x = sm[32 + ((lane + 1) & 31)] + 1.0; __syncwarp(); }
    t1 = clock64(); r[7] = (double)(t1 - t0) / N; sm[7] += x; }
  // 8: fast reciprocal-sqrt seed (MUFU.RSQ64H) + one Newton step
  { double x = seed + 2.0; t0 = clock64();
    for (int i = 0; i < N; ++i) {
      double y = __longlong_as_double(((long long)__double2hiint(x)) << 32);  // placeholder dependency
      float xf = (float)x; float yf = rsqrtf(xf); y = (double)yf;
      double e = fma(-x * y, y, 1.0); y = fma(y * e, fma(e, 0.375, 0.5), y);
      e = fma(-x * y, y, 1.0); y = fma(y * e, fma(e, 0.375, 0.5), y);
      x = y + 1.5; }
    t1 = clock64(); r[8] = (double)(t1 - t0) / N; sm[8] += x; }
  if (lane == 0) for (int i = 0; i < 9; ++i) out[i] = r[i];
  if (sm[lane & 7] == 12345.678) out[20] = sm[lane];
}

}  // namespace gpar

using namespace gpar;

extern "C" int gpar_debug_latency_probe(double* out, void* stream) {
  probe_latency_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(out, 1.25);
  return check_launch("gpar_debug_latency_probe");
}

extern "C" int gpar_fp64_probe(int mode, int64_t iters, double* sink, double* flops, void* stream) {
  const int blocks = 148 * 4;
  if (mode == 0) {
    probe_dmma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (flops) *flops = (double)blocks * 8 /*warps*/ * (double)iters * 16.0 * (2.0 * 8 * 8 * 4);
  } else {
    probe_dfma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (flops) *flops = (double)blocks * 256 * (double)iters * 16.0 * 2.0;
  }
  return check_launch("gpar_fp64_probe");
}
