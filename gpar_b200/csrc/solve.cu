// K3 / K6 / K7 / K9: the HBM- and latency-bound companions of the Cholesky sweep: backward
// triangular solve, log-determinant + quadratic form, GEMV, joint Gaussian draws with injected
// normals, row gather / column scatter, and small elementwise helpers.
#include "common.cuh"

namespace gpar {

// ---- backward solve, one 128-block step -------------------------------------------------
// work holds the running right-hand side.  Step bt: alpha_b = Linv_bb^T work_b (every CTA
// recomputes it: 16K MACs), CTA c == bt publishes alpha_b, CTA c < bt applies
// work_c -= L[b, c]^T alpha_b.
__global__ void __launch_bounds__(128)
backsolve_step_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ ws,
                      double* __restrict__ work, double* __restrict__ alpha, int bt) {
  __shared__ double ub[TILE];
  __shared__ double ab[TILE];
  const int c = blockIdx.x, t = threadIdx.x;
  const int64_t r0 = (int64_t)bt * TILE;
  const int kb = static_cast<int>(min64(TILE, n - r0));
  ub[t] = (t < kb) ? work[r0 + t] : 0.0;
  __syncthreads();
  const double* Linv = ws + (int64_t)bt * TILE * TILE;
  double s = 0.0;
  for (int r = t; r < kb; ++r) s = fma(Linv[r * TILE + t], ub[r], s);
  ab[t] = s;
  __syncthreads();
  if (c == bt) {
    if (t < kb) alpha[r0 + t] = s;
    return;
  }
  const int64_t c0 = (int64_t)c * TILE;
  const double* Lb = L + r0 * ldl + c0 + t;
  double acc0 = 0.0, acc1 = 0.0;
  int r = 0;
  for (; r + 1 < kb; r += 2) {
    acc0 = fma(Lb[(int64_t)r * ldl], ab[r], acc0);
    acc1 = fma(Lb[(int64_t)(r + 1) * ldl], ab[r + 1], acc1);
  }
  if (r < kb) acc0 = fma(Lb[(int64_t)r * ldl], ab[r], acc0);
  work[c0 + t] -= (acc0 + acc1);
}

__global__ void __launch_bounds__(1024)
logdet_quad_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ u,
                   double* __restrict__ out2) {
  __shared__ double red[32];
  double ld = 0.0, q = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    ld += log(L[i * ldl + i]);
    if (u) {
      const double v = u[i];
      q = fma(v, v, q);
    }
  }
  double a = block_sum(ld, red);
  double b = block_sum(q, red);
  if (threadIdx.x == 0) {
    out2[0] = 2.0 * a;
    out2[1] = b;
  }
}

__global__ void __launch_bounds__(256)
gemv_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, const double* __restrict__ x,
            double* __restrict__ y) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  const double* a = A + row * lda;
  double s0 = 0.0, s1 = 0.0;
  int64_t j = lane;
  for (; j + 32 < n; j += 64) {
    s0 = fma(a[j], x[j], s0);
    s1 = fma(a[j + 32], x[j + 32], s1);
  }
  if (j < n) s0 = fma(a[j], x[j], s0);
  const double s = warp_sum(s0 + s1);
  if (lane == 0) y[row] = s;
}

// ---- joint draws: out = mean + tril(C) z (+ sd * z2) -----------------------------------------
constexpr int ST_I = 64;  // rows per CTA
constexpr int ST_S = 16;  // samples per CTA
__global__ void __launch_bounds__(256)
sample_affine_kernel(const double* __restrict__ C, int64_t ldc, int64_t n, int64_t strideC,
                     const double* __restrict__ mean, const double* __restrict__ sd, const double* __restrict__ Z,
                     const double* __restrict__ Z2, int64_t ns, double* __restrict__ out) {
  __shared__ double Cs[ST_I][ST_I + 1];
  __shared__ double Zs[ST_S][ST_I];
  const int b = blockIdx.z;
  const int64_t i0 = (int64_t)blockIdx.x * ST_I, s0 = (int64_t)blockIdx.y * ST_S;
  const double* Cb = C + (int64_t)b * strideC;
  const int il = threadIdx.x & 63, sg = threadIdx.x >> 6;  // 4 sample groups x 4 samples
  const int64_t gi = i0 + il;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t j0 = 0; j0 <= i0; j0 += ST_I) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < ST_I * ST_I; idx += 256) {
      const int r = idx >> 6, cc = idx & 63;
      const int64_t gr = i0 + r, gc = j0 + cc;
      Cs[r][cc] = (gr < n && gc <= gr) ? Cb[gr * ldc + gc] : 0.0;
    }
    for (int idx = threadIdx.x; idx < ST_S * ST_I; idx += 256) {
      const int s = idx >> 6, cc = idx & 63;
      const int64_t gs = s0 + s, gc = j0 + cc;
      Zs[s][cc] = (gs < ns && gc < n) ? Z[((int64_t)b * ns + gs) * n + gc] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int cc = 0; cc < ST_I; ++cc) {
      const double cv = Cs[il][cc];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma(cv, Zs[sg * 4 + q][cc], acc[q]);
    }
  }
  if (gi < n) {
    const double mu = mean ? mean[(int64_t)b * n + gi] : 0.0;
    const double sdv = (sd && Z2) ? sd[(int64_t)b * n + gi] : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t gs = s0 + sg * 4 + q;
      if (gs < ns) {
        const int64_t o = ((int64_t)b * ns + gs) * n + gi;
        double v = mu + acc[q];
        if (sd && Z2) v = fma(sdv, Z2[o], v);
        out[o] = v;
      }
    }
  }
}

__global__ void gather_rows_kernel(const double* __restrict__ src, int64_t lds, const int64_t* __restrict__ idx,
                                   int64_t n_out, int64_t ncols, double* __restrict__ dst, int64_t ldd) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_out * ncols) return;
  const int64_t r = e / ncols, c = e % ncols;
  const int64_t sr = idx ? idx[r] : r;
  dst[r * ldd + c] = src[sr * lds + c];
}

__global__ void scatter_col_kernel(double* __restrict__ dst, int64_t ldd, int64_t col, const int64_t* __restrict__ idx,
                                   const double* __restrict__ src, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t dr = idx ? idx[r] : r;
  dst[dr * ldd + col] = src[r];
}

__global__ void mean_identity_kernel(const double* __restrict__ y, const double* __restrict__ d, double eps,
                                     const double* __restrict__ alpha, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = y[i] - (d[i] + eps) * alpha[i];
}

__global__ void mean_axis0_kernel(const double* __restrict__ in, int64_t ns, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int64_t k = 0; k < ns; ++k) s += in[k * n + i];
  out[i] = s / (double)ns;
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

}  // namespace gpar

using namespace gpar;

extern "C" int gpar_backsolve(const double* L, int64_t ldl, int64_t n, const double* ws, const double* u,
                              double* alpha, double* work, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!L || ldl < n) { set_error("gpar_backsolve: bad L"); return -1; }
  if (!ws) return -4;
  if (!u || !alpha || !work) { set_error("gpar_backsolve: null vector"); return -5; }
  if (n <= 0) return 0;
  cudaMemcpyAsync(work, u, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream);
  const int nt = (int)((n + TILE - 1) / TILE);
  for (int bt = nt - 1; bt >= 0; --bt)
    backsolve_step_kernel<<<bt + 1, 128, 0, stream>>>(L, ldl, n, ws, work, alpha, bt);
  return check_launch("gpar_backsolve");
}

extern "C" int gpar_logdet_quad(const double* L, int64_t ldl, int64_t n, const double* u, double* out2,
                                void* stream) {
  if (!L || !out2) return -1;
  logdet_quad_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(L, ldl, n, u, out2);
  return check_launch("gpar_logdet_quad");
}

extern "C" int gpar_gemv(const double* A, int64_t lda, int64_t m, int64_t n, const double* x, double* y,
                         void* stream) {
  if (m <= 0) return 0;
  if (!A || !x || !y) return -1;
  gemv_kernel<<<(unsigned)((m + 7) / 8), 256, 0, (cudaStream_t)stream>>>(A, lda, m, n, x, y);
  return check_launch("gpar_gemv");
}

extern "C" int gpar_sample_affine(const double* C, int64_t ldc, int64_t n, int64_t strideC, const double* mean,
                                  const double* sd, const double* Z, const double* Z2, int64_t ns, int64_t batch,
                                  double* out, void* stream) {
  if (n <= 0 || ns <= 0 || batch <= 0) return 0;
  if (!C || !Z || !out) return -1;
  if (batch > 65535 || (ns + ST_S - 1) / ST_S > 65535) { set_error("gpar_sample_affine: grid too large"); return -9; }
  dim3 grid((unsigned)((n + ST_I - 1) / ST_I), (unsigned)((ns + ST_S - 1) / ST_S), (unsigned)batch);
  sample_affine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, ldc, n, strideC, mean, sd, Z, Z2, ns, out);
  return check_launch("gpar_sample_affine");
}

extern "C" int gpar_gather_rows(const double* src, int64_t lds, const int64_t* idx, int64_t n_out, int64_t ncols,
                                double* dst, int64_t ldd, void* stream) {
  const int64_t total = n_out * ncols;
  if (total <= 0) return 0;
  gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, lds, idx, n_out, ncols,
                                                                                      dst, ldd);
  return check_launch("gpar_gather_rows");
}

extern "C" int gpar_scatter_col(double* dst, int64_t ldd, int64_t col, const int64_t* idx, const double* src,
                                int64_t n, void* stream) {
  if (n <= 0) return 0;
  scatter_col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, ldd, col, idx, src, n);
  return check_launch("gpar_scatter_col");
}

extern "C" int gpar_mean_identity(const double* y, const double* d, double eps, const double* alpha, int64_t n,
                                  double* out, void* stream) {
  if (n <= 0) return 0;
  mean_identity_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, d, eps, alpha, n, out);
  return check_launch("gpar_mean_identity");
}

extern "C" int gpar_mean_axis0(const double* in, int64_t ns, int64_t n, double* out, void* stream) {
  if (n <= 0 || ns <= 0) return 0;
  mean_axis0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ns, n, out);
  return check_launch("gpar_mean_axis0");
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
