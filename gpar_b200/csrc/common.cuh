// Shared device primitives for the gpar_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpar_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gpar_b200 kernels are written for sm_100a"
#endif

namespace gpar {

constexpr int TILE = GPAR_TILE;  // 128
constexpr int WARP = 32;

__host__ __device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

// ---- error plumbing (host) -------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

// ---- cp.async (LDGSTS), 16-byte, L2-only, zero-filling beyond src_bytes ------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- fp64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8) ---------------------------
// lane = 4*gid + tig.  a = A[gid][tig], b = B[tig][gid], c0/c1 = C[gid][2*tig + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, UBLKCP) -------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(b), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}

// Warp index broadcast from lane 0: tells the compiler the value is warp-uniform, which keeps
// branches on it (role dispatch, triangular skipping) free of WARPSYNC / BSSY divergence code.
__device__ __forceinline__ int canonical_warp() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }

// ---- warp reductions -------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum; result valid in thread 0.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? scratch[lane] : 0.0;
    t = warp_sum(t);
    return t;
  }
  return 0.0;
}

// ---- kernel evaluation on feature vectors --------------------------------------
// fx/fy point at the per-row feature arrays with stride `fs` between features.
__device__ __forceinline__ double eval_terms(const gpar_kernel_spec_t& spec, const double* fx, const double* fy,
                                             int fs) {
  double val = 0.0;
  for (int t = 0; t < spec.n_terms; ++t) {
    const gpar_term_t& T = spec.terms[t];
    if (T.type == GPAR_TERM_CONST) {
      val += T.variance;
      continue;
    }
    double acc = 0.0;
    if (T.type == GPAR_TERM_LINEAR) {
      for (int f = T.f_begin; f < T.f_end; ++f) acc = fma(fx[f * fs], fy[f * fs], acc);
      val = fma(T.variance, acc, val);
    } else {
      for (int f = T.f_begin; f < T.f_end; ++f) {
        double d = fx[f * fs] - fy[f * fs];
        acc = fma(d, d, acc);
      }
      if (T.type == GPAR_TERM_EQ)
        val = fma(T.variance, exp(-0.5 * acc), val);
      else
        val = fma(T.variance, pow(1.0 + acc / (2.0 * T.alpha), -T.alpha), val);
    }
  }
  return val;
}

__device__ __forceinline__ double eval_feature(const gpar_kernel_spec_t& spec, int f, const double* xrow) {
  double x = xrow[spec.feat_col[f]];
  int op = spec.feat_op[f];
  double a = spec.feat_a[f];
  if (op == GPAR_FEAT_SCALE) return x * a;
  double ang = x * spec.feat_b[f];
  return a * (op == GPAR_FEAT_SIN ? sin(ang) : cos(ang));
}

}  // namespace gpar
