// K3 / K6 / K7 / K9: the HBM- and latency-bound companions of the Cholesky sweep: backward
// triangular solve, log-determinant + quadratic form, GEMV, joint Gaussian draws with injected
// normals, row gather / column scatter, and small elementwise helpers.
#include "common.cuh"

namespace gpar {

// ---- backward solve alpha = L^-T u in ONE launch -------------------------------------------
// CTA c owns the 128-block b = nt-1-c of alpha (block index descending with blockIdx, so every
// dependency points at a CTA that was scheduled earlier: no co-residency requirement).  It
// accumulates s = sum_{k>b} L[k, b]^T alpha_k tile by tile as the alpha_k are published
// (acquire/release flags), then alpha_b = Linv_bb^T (u_b - s).  Loads are 16-way unrolled
// (memory-level parallelism); the chain per block is one tile GEMV + one 128x128 matvec.
__device__ __forceinline__ int ld_acquire_i(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 512 threads: column c = t & 127 of the block, row quarter q = t >> 7 (32 rows each).  The 32
// tile values a thread needs for step k are loaded BEFORE it waits for alpha_k (they only depend
// on the factor), so the dependent chain per block is flag -> 32 FMAs -> shared reduce -> 32 FMAs.
__global__ void __launch_bounds__(512)
backsolve_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ ws,
                 const double* __restrict__ u, double* __restrict__ alpha, int* __restrict__ ready, int nt) {
  __shared__ double v_s[TILE];
  __shared__ double part[4][TILE];
  const int b = nt - 1 - blockIdx.x;
  const int t = threadIdx.x, c = t & 127, q = t >> 7;
  const int64_t c0 = (int64_t)b * TILE;
  const int kb = static_cast<int>(min64(TILE, n - c0));
  const bool col_ok = c < kb;
  // this thread's quarter of column c of Linv_bb (rows 32 q .. 32 q + 31), kept in registers
  const double* Linv = ws + (int64_t)b * TILE * TILE;
  double li[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int r = 32 * q + i;
    li[i] = (col_ok && r < kb && r >= (c & ~7)) ? __ldcg(Linv + r * TILE + c) : 0.0;
  }
  double s = 0.0;  // threads with q == 0 hold the running sum for column c0 + c
  for (int k = nt - 1; k > b; --k) {
    const int64_t r0 = (int64_t)k * TILE;
    const int kr = static_cast<int>(min64(TILE, n - r0));
    double x[32];
    const double* Lt = L + (r0 + 32 * q) * ldl + c0 + c;
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (col_ok && 32 * q + i < kr) ? __ldcg(Lt + (int64_t)i * ldl) : 0.0;
    while (ld_acquire_i(ready + k) == 0) __nanosleep(20);
    __syncthreads();  // previous step's readers of v_s / part are done
    if (t < TILE) v_s[t] = (t < kr) ? __ldcg(alpha + r0 + t) : 0.0;
    __syncthreads();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      a0 = fma(x[i], v_s[32 * q + i], a0);
      a1 = fma(x[i + 1], v_s[32 * q + i + 1], a1);
      a2 = fma(x[i + 2], v_s[32 * q + i + 2], a2);
      a3 = fma(x[i + 3], v_s[32 * q + i + 3], a3);
    }
    part[q][c] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (q == 0) s += (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]);
  }
  __syncthreads();
  if (q == 0) v_s[c] = col_ok ? u[c0 + c] - s : 0.0;
  __syncthreads();
  // alpha_b[c] = sum_r Linv[r][c] v[r]
  {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      a0 = fma(li[i], v_s[32 * q + i], a0);
      a1 = fma(li[i + 1], v_s[32 * q + i + 1], a1);
      a2 = fma(li[i + 2], v_s[32 * q + i + 2], a2);
      a3 = fma(li[i + 3], v_s[32 * q + i + 3], a3);
    }
    part[q][c] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (q == 0 && col_ok) alpha[c0 + c] = (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]);
  __threadfence();
  __syncthreads();
  if (t == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(ready + b), "r"(1) : "memory");
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
                     const double* __restrict__ mean, const double* __restrict__ sd, int64_t strideSd,
                     const double* __restrict__ Z, const double* __restrict__ Z2, int64_t ns,
                     double* __restrict__ out) {
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
    const double sdv = (sd && Z2) ? sd[(int64_t)b * strideSd + gi] : 0.0;
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

// ---- VFE helpers -------------------------------------------------------------
// dst (cols x rows, ldd) = transpose(src (rows x cols, lds)) with row r of src scaled by scale[r].
__global__ void __launch_bounds__(256)
transpose_scale_kernel(const double* __restrict__ src, int64_t lds, int64_t rows, int64_t cols,
                       const double* __restrict__ scale, double* __restrict__ dst, int64_t ldd) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    double v = 0.0;
    if (r < rows && c < cols) v = src[r * lds + c] * (scale ? scale[r] : 1.0);
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) dst[c * ldd + r] = tile[tx][i];
  }
}

// out[0] = sum_j [ (k(x_j, x_j) - ||Bt_j||^2) / sigma_j + log(2 pi sigma_j) + y_j^2 / sigma_j ]
// (trace, normaliser and data terms of the Titsias bound; Bt = K_xz L_z^-T is n x M).  One CTA:
// deterministic reduction.
__global__ void __launch_bounds__(1024)
vfe_rowterms_kernel(const __grid_constant__ gpar_kernel_spec_t spec, const double* __restrict__ X, int64_t ldx,
                    int64_t n, const double* __restrict__ Bt, int64_t ldb, int64_t M,
                    const double* __restrict__ sigma, const double* __restrict__ y, const double* __restrict__ Pm,
                    const double* __restrict__ Pp, int64_t ldp, int64_t Mp, double* __restrict__ out) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  double acc = 0.0;
  for (int64_t j = warp; j < n; j += nw) {
    const double* b = Bt + j * ldb;
    double ss = 0.0;
    for (int64_t c = lane; c < M; c += 32) ss = fma(b[c], b[c], ss);
    if (Pm) {  // the prior of the bound is a sparse posterior: k_jj - |Pm_j|^2 + |Pp_j|^2
      const double* pm = Pm + j * ldp;
      const double* pp = Pp + j * ldp;
      for (int64_t c = lane; c < Mp; c += 32) ss += pm[c] * pm[c] - pp[c] * pp[c];
    }
    ss = warp_sum(ss);
    if (lane == 0) {
      double kjj = 0.0;
      const double* xr = X + j * ldx;
      for (int t = 0; t < spec.n_terms; ++t) {
        const gpar_term_t& T = spec.terms[t];
        if (T.type == GPAR_TERM_LINEAR) {
          double a = 0.0;
          for (int f = T.f_begin; f < T.f_end; ++f) {
            const double ph = eval_feature(spec, f, xr);
            a = fma(ph, ph, a);
          }
          kjj = fma(T.variance, a, kjj);
        } else {
          kjj += T.variance;  // EQ / RQ / const at zero distance
        }
      }
      const double sg = sigma[j], yy = y[j];
      acc += (kjj - ss) / sg + log(6.283185307179586 * sg) + yy * yy / sg;
    }
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) out[blockIdx.x] = tot;  // per-CTA partial; vfe_rowterms_reduce_kernel adds them in order
}

__global__ void __launch_bounds__(256)
vfe_rowterms_reduce_kernel(const double* __restrict__ partial, int np, double* __restrict__ out) {
  __shared__ double red[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) v += partial[i];
  v = block_sum(v, red);
  if (threadIdx.x == 0) out[0] = v;
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
  const int nt = (int)((n + TILE - 1) / TILE);
  // `work` (n doubles) hosts the nt ready flags.
  cudaMemsetAsync(work, 0, sizeof(int) * (size_t)nt, stream);
  backsolve_kernel<<<nt, 512, 0, stream>>>(L, ldl, n, ws, u, alpha, reinterpret_cast<int*>(work), nt);
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
                                  const double* sd, int64_t strideSd, const double* Z, const double* Z2, int64_t ns,
                                  int64_t batch, double* out, void* stream) {
  if (n <= 0 || ns <= 0 || batch <= 0) return 0;
  if (!C || !Z || !out) return -1;
  if (batch > 65535 || (ns + ST_S - 1) / ST_S > 65535) { set_error("gpar_sample_affine: grid too large"); return -9; }
  dim3 grid((unsigned)((n + ST_I - 1) / ST_I), (unsigned)((ns + ST_S - 1) / ST_S), (unsigned)batch);
  sample_affine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, ldc, n, strideC, mean, sd, strideSd, Z, Z2, ns, out);
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

// inout[i] += sum_k in[k][i], k in fixed order: combines the K-slices of a split SYRK deterministically.
__global__ void sum_axis0_add_kernel(const double* __restrict__ in, int64_t ns, int64_t n, double* __restrict__ inout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = inout[i];
  for (int64_t k = 0; k < ns; ++k) s += in[k * n + i];
  inout[i] = s;
}

extern "C" int gpar_sum_axis0_add(const double* in, int64_t ns, int64_t n, double* inout, void* stream) {
  if (n <= 0 || ns <= 0) return 0;
  if (!in || !inout) return -1;
  sum_axis0_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ns, n, inout);
  return check_launch("gpar_sum_axis0_add");
}

// Two percentiles over the sample axis (regression.py:593-594: np.percentile(samples, q, axis=0), numpy's
// default "linear" method): for every entry i the order statistics j and j + 1 of in[:, i] are found by
// rank counting (no scratch, no writes; the S x n block stays in L2) and combined with numpy's _lerp.
__device__ __forceinline__ double numpy_lerp(double a, double b, double t) {
  const double d = b - a;
  return (t >= 0.5) ? b - d * (1.0 - t) : a + d * t;
}

__global__ void __launch_bounds__(128)
percentile2_axis0_kernel(const double* __restrict__ in, int64_t ns, int64_t n, int64_t j_lo, double g_lo, int64_t j_hi,
                         double g_hi, double* __restrict__ out_lo, double* __restrict__ out_hi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t jl1 = (j_lo + 1 < ns) ? j_lo + 1 : ns - 1, jh1 = (j_hi + 1 < ns) ? j_hi + 1 : ns - 1;
  double al = 0.0, bl = 0.0, ah = 0.0, bh = 0.0;
  for (int64_t k = 0; k < ns; ++k) {
    const double x = in[k * n + i];
    int64_t rank = 0;
    for (int64_t l = 0; l < ns; ++l) {
      const double v = in[l * n + i];
      rank += (v < x) || (v == x && l < k);
    }
    if (rank == j_lo) al = x;
    if (rank == jl1) bl = x;
    if (rank == j_hi) ah = x;
    if (rank == jh1) bh = x;
  }
  out_lo[i] = numpy_lerp(al, bl, g_lo);
  out_hi[i] = numpy_lerp(ah, bh, g_hi);
}

extern "C" int gpar_percentile2_axis0(const double* in, int64_t ns, int64_t n, int64_t j_lo, double g_lo, int64_t j_hi,
                                      double g_hi, double* out_lo, double* out_hi, void* stream) {
  if (n <= 0 || ns <= 0) return 0;
  if (!in) return -1;
  if (j_lo < 0 || j_lo >= ns || j_hi < 0 || j_hi >= ns) { set_error("gpar_percentile2_axis0: order statistic out of range"); return -4; }
  if (!out_lo || !out_hi) return -8;
  percentile2_axis0_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(in, ns, n, j_lo, g_lo, j_hi,
                                                                                          g_hi, out_lo, out_hi);
  return check_launch("gpar_percentile2_axis0");
}

extern "C" int gpar_transpose_scale(const double* src, int64_t lds, int64_t rows, int64_t cols, const double* scale,
                                    double* dst, int64_t ldd, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!src || !dst) return -1;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  if (grid.y > 65535) { set_error("gpar_transpose_scale: too many rows"); return -3; }
  transpose_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, cols, scale, dst, ldd);
  return check_launch("gpar_transpose_scale");
}

extern "C" int gpar_vfe_rowterms(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t n,
                                 const double* Bt, int64_t ldb, int64_t M, const double* sigma, const double* y,
                                 const double* Pm, const double* Pp, int64_t ldp, int64_t Mp, double* workspace,
                                 double* out, void* stream) {
  if (!spec || !out) return -1;
  if (!workspace) { set_error("gpar_vfe_rowterms: workspace of GPAR_VFE_ROWTERMS_WS doubles required"); return -10; }
  // one CTA used to stream the whole n x M block (1.7 ms at n = 16384, M = 512: 40 GB/s); now the rows are
  // spread over the grid and the per-CTA partials are combined in a fixed order
  if ((Pm == nullptr) != (Pp == nullptr)) { set_error("gpar_vfe_rowterms: Pm and Pp go together"); return -10; }
  vfe_rowterms_kernel<<<GPAR_VFE_ROWTERMS_WS, 256, 0, (cudaStream_t)stream>>>(*spec, X, ldx, n, Bt, ldb, M, sigma, y,
                                                                             Pm, Pp, ldp, Mp, workspace);
  vfe_rowterms_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(workspace, GPAR_VFE_ROWTERMS_WS, out);
  return check_launch("gpar_vfe_rowterms");
}

__global__ void untransform_kernel(double* __restrict__ a, int64_t total, int64_t p, const double* __restrict__ scale,
                                   const double* __restrict__ shift, int kind) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t j = e % p;
  double v = a[e];
  if (scale) v = v * scale[j] + shift[j];  // y * std + mean, rounded like numpy (no fma)
  if (kind == GPAR_TRANSFORM_LOG) {
    v = exp(v);
  } else if (kind == GPAR_TRANSFORM_SQUISH) {
    const double m = exp(fabs(v)) - 1.0;
    v = (v > 0.0) ? m : ((v < 0.0) ? -m : 0.0 * v);  // sign(v) * (exp|v| - 1); NaN stays NaN
  }
  a[e] = v;
}

extern "C" int gpar_untransform(double* a, int64_t rows, int64_t p, const double* scale, const double* shift,
                                int kind, void* stream) {
  if (rows <= 0 || p <= 0) return 0;
  if (!a) return -1;
  if ((scale == nullptr) != (shift == nullptr)) { set_error("gpar_untransform: scale and shift go together"); return -4; }
  if (kind < 0 || kind > 2) { set_error("gpar_untransform: unknown transform kind"); return -6; }
  const int64_t total = rows * p;
  untransform_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, total, p, scale, shift, kind);
  return check_launch("gpar_untransform");
}

// out[i] = sum_c A[i][c]^2: one warp per row (fixed lane order: deterministic).
__global__ void __launch_bounds__(256)
row_sqnorm_kernel(const double* __restrict__ A, int64_t lda, int64_t n, int64_t k, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const double* a = A + row * lda;
  double s = 0.0;
  for (int64_t c = lane; c < k; c += 32) s = fma(a[c], a[c], s);
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

extern "C" int gpar_row_sqnorm(const double* A, int64_t lda, int64_t n, int64_t k, double* out, void* stream) {
  if (n <= 0) return 0;
  if (!A || !out || lda < k) { set_error("gpar_row_sqnorm: bad arguments"); return -1; }
  row_sqnorm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(A, lda, n, k, out);
  return check_launch("gpar_row_sqnorm");
}

__global__ void axpy_kernel(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fma(a, x[i], y[i]);
}

extern "C" int gpar_axpy(int64_t n, double a, const double* x, double* y, void* stream) {
  if (n <= 0) return 0;
  if (!x || !y) return -3;
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, a, x, y);
  return check_launch("gpar_axpy");
}

