// K1: Gram matrix of the GPAR kernel family (feature-map normal form), and the fused
// cross-covariance x vector product.  HBM-bound on the n^2 * 8 B output; the input rows are
// staged with the bulk-copy (TMA) engine, features (scaling, sin/cos maps) are built once per
// tile in shared memory, and each thread produces a 4x4 register block of kernel values.
#include "common.cuh"

namespace gpar {

constexpr int GT = 64;  // Gram tile edge

__device__ __forceinline__ void stage_rows(double* raw, const double* src, int64_t ld, int rows, uint64_t* bar,
                                           uint32_t& parity) {
  // rows x ld doubles, contiguous in global memory.
  uint32_t bytes = static_cast<uint32_t>(rows * ld * 8);
  bool bulk_ok = (bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && bytes > 0;
  if (bulk_ok) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, bytes);
      tma_bulk_g2s(raw, src, bytes, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1;
  } else {
    for (int i = threadIdx.x; i < rows * ld; i += blockDim.x) raw[i] = src[i];
    __syncthreads();
  }
}

__device__ __forceinline__ void build_features(const gpar_kernel_spec_t& spec, double* feat, const double* raw,
                                               int64_t ld, int rows) {
  int F = spec.n_feats;
  for (int i = threadIdx.x; i < F * GT; i += blockDim.x) {
    int f = i / GT, r = i % GT;
    feat[i] = (r < rows) ? eval_feature(spec, f, raw + r * ld) : 0.0;
  }
}

__global__ void __launch_bounds__(256)
gram_kernel(const __grid_constant__ gpar_kernel_spec_t spec, const double* __restrict__ X, int64_t ldx, int64_t nx,
            const double* __restrict__ Y, int64_t ldy, int64_t ny, const double* __restrict__ diag_add, double eps,
            int sym, int lower_only, double* __restrict__ out, int64_t ldo, int64_t strideX, int64_t strideY,
            int64_t strideD, int64_t strideO) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (sym && lower_only && bj > bi) return;
  X += (int64_t)blockIdx.z * strideX;
  Y += (int64_t)blockIdx.z * strideY;
  out += (int64_t)blockIdx.z * strideO;
  if (diag_add) diag_add += (int64_t)blockIdx.z * strideD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = spec.n_feats;
  // layout: [bar (16 B)] [raw_x GT*ldx] [raw_y GT*ldy] [fx F*GT] [fy F*GT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* raw_x = reinterpret_cast<double*>(smem_raw + 16);
  double* raw_y = raw_x + GT * ldx;
  double* fx = raw_y + GT * ldy;
  double* fy = fx + F * GT;

  const int rows_x = static_cast<int>(min64(GT, nx - (int64_t)bi * GT));
  const int rows_y = static_cast<int>(min64(GT, ny - (int64_t)bj * GT));
  const bool diag_tile = sym && (bi == bj);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t parity = 0;
  stage_rows(raw_x, X + (int64_t)bi * GT * ldx, ldx, rows_x, bar, parity);
  if (!diag_tile) stage_rows(raw_y, Y + (int64_t)bj * GT * ldy, ldy, rows_y, bar, parity);
  build_features(spec, fx, raw_x, ldx, rows_x);
  if (!diag_tile) build_features(spec, fy, raw_y, ldy, rows_y);
  __syncthreads();
  const double* fyy = diag_tile ? fx : fy;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double val[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) val[r][c] = 0.0;

  for (int t = 0; t < spec.n_terms; ++t) {
    const int type = spec.terms[t].type;
    const double var = spec.terms[t].variance;
    if (type == GPAR_TERM_CONST) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) val[r][c] += var;
      continue;
    }
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    const int f0 = spec.terms[t].f_begin, f1 = spec.terms[t].f_end;
    if (type == GPAR_TERM_LINEAR) {
      for (int f = f0; f < f1; ++f) {
        double xv[4], yv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xv[r] = fx[f * GT + ty + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) yv[c] = fyy[f * GT + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fma(xv[r], yv[c], acc[r][c]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) val[r][c] = fma(var, acc[r][c], val[r][c]);
    } else {
      for (int f = f0; f < f1; ++f) {
        double xv[4], yv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xv[r] = fx[f * GT + ty + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) yv[c] = fyy[f * GT + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double d = xv[r] - yv[c];
            acc[r][c] = fma(d, d, acc[r][c]);
          }
      }
      if (type == GPAR_TERM_EQ) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) val[r][c] = fma(var, exp(-0.5 * acc[r][c]), val[r][c]);
      } else {
        const double alpha = spec.terms[t].alpha;
        const double inv2a = 1.0 / (2.0 * alpha);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) val[r][c] = fma(var, pow(fma(acc[r][c], inv2a, 1.0), -alpha), val[r][c]);
      }
    }
  }

#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int lr = ty + 16 * r;
    if (lr >= rows_x) continue;
    const int64_t gr = (int64_t)bi * GT + lr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int lc = tx + 16 * c;
      if (lc >= rows_y) continue;
      const int64_t gc = (int64_t)bj * GT + lc;
      double v = val[r][c];
      if (sym && gr == gc) v += (diag_add ? diag_add[gr] : 0.0) + eps;
      out[gr * ldo + gc] = v;
    }
  }
}

// out[j] = sum_i k(Xq[j], Xa[i]) v[i].  One CTA per QT query rows; the a-rows stream through
// shared memory in chunks of 256 (features built cooperatively).  QT = 4 keeps the grid above the SM
// count for the few hundred missing rows of an imputation call and lets three CTAs share an SM
// (the fp64 exp chains are latency-bound at 8 warps per SM: 0.27 -> 0.1 ms per call at C3).
constexpr int QT = 4;
constexpr int AC = 256;

__global__ void __launch_bounds__(AC, 3)
gram_gemv_kernel(const __grid_constant__ gpar_kernel_spec_t spec, const double* __restrict__ Xq, int64_t ldq,
                 int64_t nq, const double* __restrict__ Xa, int64_t lda, int64_t na, const double* __restrict__ v,
                 double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = spec.n_feats;
  double* fq = reinterpret_cast<double*>(smem_raw);  // [F][QT]
  double* fa = fq + F * QT;                           // [F][AC]
  double* red = fa + F * AC;                          // [32]
  const int64_t q0 = (int64_t)blockIdx.x * QT;
  const int nqt = static_cast<int>(min64(QT, nq - q0));
  for (int i = threadIdx.x; i < F * QT; i += blockDim.x) {
    int f = i / QT, r = i % QT;
    fq[i] = (r < nqt) ? eval_feature(spec, f, Xq + (q0 + r) * ldq) : 0.0;
  }
  double sums[QT];
#pragma unroll
  for (int q = 0; q < QT; ++q) sums[q] = 0.0;
  for (int64_t a0 = 0; a0 < na; a0 += AC) {
    __syncthreads();
    const int64_t ia = a0 + threadIdx.x;
    const bool live = ia < na;
    for (int f = 0; f < F; ++f) fa[f * AC + threadIdx.x] = live ? eval_feature(spec, f, Xa + ia * lda) : 0.0;
    __syncthreads();
    if (live) {
      const double vi = v[ia];
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        double val = 0.0;
        for (int t = 0; t < spec.n_terms; ++t) {
          const gpar_term_t& T = spec.terms[t];
          if (T.type == GPAR_TERM_CONST) {
            val += T.variance;
            continue;
          }
          double acc = 0.0;
          if (T.type == GPAR_TERM_LINEAR) {
            for (int f = T.f_begin; f < T.f_end; ++f) acc = fma(fq[f * QT + q], fa[f * AC + threadIdx.x], acc);
            val = fma(T.variance, acc, val);
          } else {
            for (int f = T.f_begin; f < T.f_end; ++f) {
              double d = fq[f * QT + q] - fa[f * AC + threadIdx.x];
              acc = fma(d, d, acc);
            }
            if (T.type == GPAR_TERM_EQ)
              val = fma(T.variance, exp(-0.5 * acc), val);
            else
              val = fma(T.variance, pow(fma(acc, 1.0 / (2.0 * T.alpha), 1.0), -T.alpha), val);
          }
        }
        sums[q] = fma(val, vi, sums[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < QT; ++q) {
    double s = block_sum(sums[q], red);
    if (threadIdx.x == 0 && q < nqt) out[q0 + q] = s;
  }
}

// ---- gradient of the log-marginal w.r.t. the kernel spec (fit) -------------------------
// d LML / d theta = sum_ij W_ij dA_ij / d theta with W = 1/2 (alpha alpha^T - A^-1).  One pass over the
// lower triangle of A^-1 (HBM-bound: 8 B per pair) accumulates, for every term t and feature f of the
// spec, the raw sums from which the host finishes the chain rule (gpar_b200/spec.py: spec_gradient):
//   out[2 t]      Sv_t   = sum W k_t / v_t                    (d/dv_t;  CONST: sum W)
//   out[2 t + 1]  Sal_t  = sum W k_t (-ln(1+u) + u/(1+u))     (RQ only: d/dalpha, u = r^2 / (2 alpha))
//   out[16 + 2 f]     S1_f = sum W g_t (phi_f(x) - phi_f(y))^2        (EQ / RQ: d/da_f = -S1_f / a_f)
//                          = sum W v_t phi_f(x) phi_f(y)              (LINEAR:  d/da_f = 2 S1_f / a_f)
//   out[16 + 2 f + 1] S2_f = sum W g_t (phi_f(x) - phi_f(y)) (psi_f(x) - psi_f(y))   (sin / cos features:
//                            psi = d phi / d b;  d/db_f = -S2_f)
//   out[208]      Sd     = sum_i W_ii dvec_i                  (diagonal term: noise / w)
// with g_t = k_t (EQ) or v_t (1+u)^(-alpha-1) (RQ), sums over ALL pairs (the lower triangle is visited,
// off-diagonal pairs weigh 2).  Every CTA (one 64 x 64 tile) writes its partial sums to `partials`;
// grad_reduce_kernel adds them in a fixed order: the gradient is bitwise reproducible.
constexpr int GRAD_NP = 2 * GPAR_MAX_TERMS + 2 * GPAR_MAX_FEATS + 1;

__device__ __forceinline__ double eval_feature_db(const gpar_kernel_spec_t& spec, int f, const double* xrow) {
  const int op = spec.feat_op[f];
  if (op == GPAR_FEAT_SCALE) return 0.0;
  const double x = xrow[spec.feat_col[f]];
  const double ang = x * spec.feat_b[f];
  return spec.feat_a[f] * x * (op == GPAR_FEAT_SIN ? cos(ang) : -sin(ang));
}

// RECT (gpar_gram_wgrad): the same sums for a rectangular block k(x_i, y_j), i < n, j < ny, with explicit
// weights W_ij = ux_i uy_j + sx_i G_ij (every pair counted once; ux / uy / sx / G optional) -- the
// sum_{m,j} G_zx[m, j] dK(z_m, x_j) and sum G_zz dK(z, z') terms of the VFE bound's gradient.
struct RectW {
  const double* Y; int64_t ldy; int64_t ny;
  const double* G; int64_t ldg;
  const double* ux; const double* uy; const double* sx;
};

template <bool RECT>
__global__ void __launch_bounds__(256)
gram_grad_kernel(const __grid_constant__ gpar_kernel_spec_t spec, const double* __restrict__ X, int64_t ldx, int64_t n,
                 const double* __restrict__ alpha, const double* __restrict__ Ainv, int64_t lda,
                 const double* __restrict__ dvec, const RectW rw, double* __restrict__ partials) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (!RECT && bj > bi) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = spec.n_feats;
  // layout: [bar 16 B] [raw_x GT*ldx] [raw_y GT*ldx] [fx F*GT] [fy F*GT] [px F*GT] [py F*GT] [ax GT] [ay GT] [accw 8*NP]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* raw_x = reinterpret_cast<double*>(smem_raw + 16);
  const int64_t ldy = RECT ? rw.ldy : ldx;
  double* raw_y = raw_x + GT * ldx;
  double* fx = raw_y + GT * ldy;
  double* fy = fx + F * GT;
  double* px = fy + F * GT;
  double* py = px + F * GT;
  double* ax = py + F * GT;
  double* ay = ax + GT;
  double* accw = ay + GT;
  const int rows_x = static_cast<int>(min64(GT, n - (int64_t)bi * GT));
  const int rows_y = static_cast<int>(min64(GT, (RECT ? rw.ny : n) - (int64_t)bj * GT));
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 8 * GRAD_NP; i += blockDim.x) accw[i] = 0.0;
  __syncthreads();
  uint32_t parity = 0;
  stage_rows(raw_x, X + (int64_t)bi * GT * ldx, ldx, rows_x, bar, parity);
  stage_rows(raw_y, (RECT ? rw.Y : X) + (int64_t)bj * GT * ldy, ldy, rows_y, bar, parity);
  build_features(spec, fx, raw_x, ldx, rows_x);
  build_features(spec, fy, raw_y, ldy, rows_y);
  for (int i = threadIdx.x; i < F * GT; i += blockDim.x) {
    const int f = i / GT, r = i % GT;
    px[i] = (r < rows_x) ? eval_feature_db(spec, f, raw_x + r * ldx) : 0.0;
    py[i] = (r < rows_y) ? eval_feature_db(spec, f, raw_y + r * ldy) : 0.0;
  }
  if (threadIdx.x < GT) {
    if (RECT) {
      ax[threadIdx.x] = (threadIdx.x < rows_x && rw.ux) ? rw.ux[(int64_t)bi * GT + threadIdx.x] : 0.0;
      ay[threadIdx.x] = (threadIdx.x < rows_y && rw.uy) ? rw.uy[(int64_t)bj * GT + threadIdx.x] : 0.0;
    } else {
      ax[threadIdx.x] = (threadIdx.x < rows_x) ? alpha[(int64_t)bi * GT + threadIdx.x] : 0.0;
      ay[threadIdx.x] = (threadIdx.x < rows_y) ? alpha[(int64_t)bj * GT + threadIdx.x] : 0.0;
    }
  }
  __syncthreads();

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // W = m * 1/2 (alpha_i alpha_j - Ainv_ij), m = 1 on the diagonal, 2 below, 0 above / out of range
  double W[4][4];
  double sdiag = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int lr = ty + 16 * r;
    const int64_t gr = (int64_t)bi * GT + lr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int lc = tx + 16 * c;
      const int64_t gc = (int64_t)bj * GT + lc;
      double w = 0.0;
      if (RECT) {
        if (lr < rows_x && lc < rows_y) {
          w = ax[lr] * ay[lc];
          if (rw.G) w = fma(rw.sx ? rw.sx[gr] : 1.0, __ldcg(rw.G + gr * rw.ldg + gc), w);
        }
      } else if (lr < rows_x && lc < rows_y && gc <= gr) {
        w = 0.5 * (ax[lr] * ay[lc] - __ldcg(Ainv + gr * lda + gc));
        if (gc == gr) {
          if (dvec) sdiag = fma(w, dvec[gr], sdiag);
        } else {
          w *= 2.0;
        }
      }
      W[r][c] = w;
    }
  }
  auto reduce_add = [&](double v, int slot) {
    v = warp_sum(v);
    if (lane == 0) accw[warp * GRAD_NP + slot] += v;
  };
  reduce_add(sdiag, GRAD_NP - 1);

  for (int t = 0; t < spec.n_terms; ++t) {
    const int type = spec.terms[t].type;
    const double var = spec.terms[t].variance;
    if (type == GPAR_TERM_CONST) {
      double sv = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) sv += W[r][c];
      reduce_add(sv, 2 * t);
      continue;
    }
    const int f0 = spec.terms[t].f_begin, f1 = spec.terms[t].f_end;
    if (type == GPAR_TERM_LINEAR) {
      double sv = 0.0;
      for (int f = f0; f < f1; ++f) {
        double xv[4], yv[4], s1 = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) xv[r] = fx[f * GT + ty + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) yv[c] = fy[f * GT + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) s1 = fma(W[r][c], xv[r] * yv[c], s1);
        sv += s1;
        reduce_add(var * s1, 2 * GPAR_MAX_TERMS + 2 * f);
      }
      reduce_add(sv, 2 * t);
      continue;
    }
    // EQ / RQ: squared distance over the term's features
    double r2[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) r2[r][c] = 0.0;
    for (int f = f0; f < f1; ++f) {
      double xv[4], yv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) xv[r] = fx[f * GT + ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) yv[c] = fy[f * GT + tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double d = xv[r] - yv[c];
          r2[r][c] = fma(d, d, r2[r][c]);
        }
    }
    double Wg[4][4];  // W * g_t
    double sv = 0.0, sal = 0.0;
    if (type == GPAR_TERM_EQ) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double e = exp(-0.5 * r2[r][c]);
          sv = fma(W[r][c], e, sv);
          Wg[r][c] = W[r][c] * var * e;
        }
    } else {
      const double al = spec.terms[t].alpha;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double u = r2[r][c] / (2.0 * al);
          const double base = pow(1.0 + u, -al);
          sv = fma(W[r][c], base, sv);
          sal = fma(W[r][c] * var * base, u / (1.0 + u) - log1p(u), sal);
          Wg[r][c] = W[r][c] * var * base / (1.0 + u);
        }
    }
    reduce_add(sv, 2 * t);
    if (type == GPAR_TERM_RQ) reduce_add(sal, 2 * t + 1);
    for (int f = f0; f < f1; ++f) {
      double xv[4], yv[4], s1 = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) xv[r] = fx[f * GT + ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) yv[c] = fy[f * GT + tx + 16 * c];
      const bool trig = spec.feat_op[f] != GPAR_FEAT_SCALE;
      double s2 = 0.0;
      if (trig) {
        double pxv[4], pyv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) pxv[r] = px[f * GT + ty + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) pyv[c] = py[f * GT + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double d = xv[r] - yv[c];
            s1 = fma(Wg[r][c], d * d, s1);
            s2 = fma(Wg[r][c], d * (pxv[r] - pyv[c]), s2);
          }
        reduce_add(s2, 2 * GPAR_MAX_TERMS + 2 * f + 1);
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double d = xv[r] - yv[c];
            s1 = fma(Wg[r][c], d * d, s1);
          }
      }
      reduce_add(s1, 2 * GPAR_MAX_TERMS + 2 * f);
    }
  }
  __syncthreads();
  double* dst = partials + ((int64_t)bi * gridDim.x + bj) * GRAD_NP;
  for (int q = threadIdx.x; q < GRAD_NP; q += blockDim.x) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += accw[w * GRAD_NP + q];
    dst[q] = v;
  }
}

// out[q] = sum over the lower tiles (bi >= bj), row-major tile order: fixed summation order.
__global__ void __launch_bounds__(256)
grad_reduce_kernel(const double* __restrict__ partials, int nb, double* __restrict__ out) {
  __shared__ double red[32];
  const int q = blockIdx.x;
  double v = 0.0;
  const int ntile = nb * (nb + 1) / 2;
  // thread k owns the tiles k, k + 256, ... of the lower-triangular enumeration; block_sum is deterministic
  for (int idx = threadIdx.x; idx < ntile; idx += blockDim.x) {
    int bi = static_cast<int>((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= idx) ++bi;
    while (bi * (bi + 1) / 2 > idx) --bi;
    const int bj = idx - bi * (bi + 1) / 2;
    v += partials[((int64_t)bi * nb + bj) * GRAD_NP + q];
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) out[q] = v;
}

// out[q] = sum over ALL tiles of a rectangular pass, row-major tile order (fixed summation order).
__global__ void __launch_bounds__(256)
grad_reduce_rect_kernel(const double* __restrict__ partials, int ntile, double* __restrict__ out) {
  __shared__ double red[32];
  const int q = blockIdx.x;
  double v = 0.0;
  for (int idx = threadIdx.x; idx < ntile; idx += blockDim.x) v += partials[(int64_t)idx * GRAD_NP + q];
  v = block_sum(v, red);
  if (threadIdx.x == 0) out[q] = v;
}

}  // namespace gpar

using namespace gpar;

extern "C" int gpar_gram_batched(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t nx,
                                 int64_t strideX, const double* Y, int64_t ldy, int64_t ny, int64_t strideY,
                                 const double* diag_add, int64_t strideD, double eps, int lower_only, double* out,
                                 int64_t ldo, int64_t strideO, int64_t batch, void* stream) {
  if (!spec || spec->n_feats < 0 || spec->n_feats > GPAR_MAX_FEATS || spec->n_terms < 0 ||
      spec->n_terms > GPAR_MAX_TERMS) {
    set_error("gpar_gram: bad spec");
    return -1;
  }
  if (!X || ldx <= 0) { set_error("gpar_gram: bad X"); return -2; }
  if (nx < 0) return -4;
  const int sym = (Y == nullptr);
  if (sym) { Y = X; ldy = ldx; ny = nx; strideY = strideX; }
  if (batch <= 0) return 0;
  if (batch > 65535) { set_error("gpar_gram: batch > 65535"); return -17; }
  if (ny < 0) return -7;
  if (!out || ldo < ny) { set_error("gpar_gram: bad out/ldo"); return -11; }
  if (nx == 0 || ny == 0) return 0;
  size_t smem = 16 + sizeof(double) * (size_t)(GT * ldx + GT * ldy + 2 * (size_t)spec->n_feats * GT);
  if (smem > 227 * 1024) { set_error("gpar_gram: ldx/ldy/features too large for shared memory (%zu B)", smem); return -3; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  dim3 grid((unsigned)((ny + GT - 1) / GT), (unsigned)((nx + GT - 1) / GT), (unsigned)batch);
  gram_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(*spec, X, ldx, nx, Y, ldy, ny, diag_add, eps, sym,
                                                         lower_only, out, ldo, strideX, strideY, strideD, strideO);
  return check_launch("gpar_gram");
}

extern "C" int gpar_gram(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t nx, const double* Y,
                         int64_t ldy, int64_t ny, const double* diag_add, double eps, int lower_only, double* out,
                         int64_t ldo, void* stream) {
  return gpar_gram_batched(spec, X, ldx, nx, 0, Y, ldy, ny, 0, diag_add, 0, eps, lower_only, out, ldo, 0, 1, stream);
}

extern "C" int gpar_gram_gemv(const gpar_kernel_spec_t* spec, const double* Xq, int64_t ldq, int64_t nq,
                              const double* Xa, int64_t lda, int64_t na, const double* v, double* out, void* stream) {
  if (!spec || spec->n_feats > GPAR_MAX_FEATS || spec->n_terms > GPAR_MAX_TERMS) { set_error("gpar_gram_gemv: bad spec"); return -1; }
  if (nq <= 0) return 0;
  size_t smem = sizeof(double) * ((size_t)spec->n_feats * (QT + AC) + 32);
  if (smem > 227 * 1024) { set_error("gpar_gram_gemv: too many features"); return -1; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gram_gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  unsigned grid = (unsigned)((nq + QT - 1) / QT);
  gram_gemv_kernel<<<grid, AC, smem, (cudaStream_t)stream>>>(*spec, Xq, ldq, nq, Xa, lda, na, v, out);
  return check_launch("gpar_gram_gemv");
}

extern "C" size_t gpar_gram_grad_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  const size_t nb = (size_t)((n + GT - 1) / GT);
  return nb * nb * GRAD_NP * sizeof(double);
}

extern "C" int gpar_gram_grad(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t n,
                              const double* alpha, const double* Ainv, int64_t lda, const double* dvec,
                              double* workspace, double* out, void* stream) {
  if (!spec || spec->n_feats < 0 || spec->n_feats > GPAR_MAX_FEATS || spec->n_terms < 0 ||
      spec->n_terms > GPAR_MAX_TERMS) { set_error("gpar_gram_grad: bad spec"); return -1; }
  if (!X || ldx <= 0) { set_error("gpar_gram_grad: bad X"); return -2; }
  if (!alpha) return -5;
  if (!Ainv || lda < n) { set_error("gpar_gram_grad: bad Ainv"); return -6; }
  if (!workspace) return -9;
  if (!out) return -10;
  if (n <= 0) { cudaMemsetAsync(out, 0, sizeof(double) * GRAD_NP, (cudaStream_t)stream); return 0; }
  const size_t smem = 16 + sizeof(double) * ((size_t)2 * GT * ldx + 4 * (size_t)spec->n_feats * GT + 2 * GT + 8 * GRAD_NP);
  if (smem > 227 * 1024) { set_error("gpar_gram_grad: ldx/features too large for shared memory (%zu B)", smem); return -3; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gram_grad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const unsigned nb = (unsigned)((n + GT - 1) / GT);
  RectW none = {nullptr, 0, 0, nullptr, 0, nullptr, nullptr, nullptr};
  gram_grad_kernel<false><<<dim3(nb, nb), 256, smem, (cudaStream_t)stream>>>(*spec, X, ldx, n, alpha, Ainv, lda, dvec,
                                                                            none, workspace);
  grad_reduce_kernel<<<GRAD_NP, 256, 0, (cudaStream_t)stream>>>(workspace, (int)nb, out);
  return check_launch("gpar_gram_grad");
}

extern "C" size_t gpar_gram_wgrad_workspace_bytes(int64_t nx, int64_t ny) {
  if (nx <= 0 || ny <= 0) return 0;
  return (size_t)((nx + GT - 1) / GT) * (size_t)((ny + GT - 1) / GT) * GRAD_NP * sizeof(double);
}

extern "C" int gpar_gram_wgrad(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t nx,
                               const double* Y, int64_t ldy, int64_t ny, const double* G, int64_t ldg,
                               const double* sx, const double* ux, const double* uy, double* workspace, double* out,
                               void* stream) {
  if (!spec || spec->n_feats < 0 || spec->n_feats > GPAR_MAX_FEATS || spec->n_terms < 0 ||
      spec->n_terms > GPAR_MAX_TERMS) { set_error("gpar_gram_wgrad: bad spec"); return -1; }
  if (!X || ldx <= 0 || !Y || ldy <= 0) { set_error("gpar_gram_wgrad: bad X / Y"); return -2; }
  if (G && ldg < ny) { set_error("gpar_gram_wgrad: bad G"); return -8; }
  if ((ux == nullptr) != (uy == nullptr)) { set_error("gpar_gram_wgrad: ux and uy go together"); return -11; }
  if (!workspace) return -13;
  if (!out) return -14;
  if (nx <= 0 || ny <= 0) { cudaMemsetAsync(out, 0, sizeof(double) * GRAD_NP, (cudaStream_t)stream); return 0; }
  const size_t smem = 16 + sizeof(double) * ((size_t)GT * ldx + (size_t)GT * ldy + 4 * (size_t)spec->n_feats * GT + 2 * GT +
                                             8 * GRAD_NP);
  if (smem > 227 * 1024) { set_error("gpar_gram_wgrad: ldx/features too large for shared memory (%zu B)", smem); return -3; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gram_grad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const unsigned nbx = (unsigned)((nx + GT - 1) / GT), nby = (unsigned)((ny + GT - 1) / GT);
  if (nbx > 65535) { set_error("gpar_gram_wgrad: more than 65535 row tiles"); return -4; }
  RectW rw = {Y, ldy, ny, G, ldg, ux, uy, sx};
  gram_grad_kernel<true><<<dim3(nby, nbx), 256, smem, (cudaStream_t)stream>>>(*spec, X, ldx, nx, nullptr, nullptr, 0,
                                                                            nullptr, rw, workspace);
  grad_reduce_rect_kernel<<<GRAD_NP, 256, 0, (cudaStream_t)stream>>>(workspace, (int)(nbx * nby), out);
  return check_launch("gpar_gram_wgrad");
}
