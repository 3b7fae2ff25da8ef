// K2 / K4 / K5: blocked Cholesky with appended row blocks, triangular solve of row blocks, and
// the symmetric rank-k downdate -- all on the DMMA GEMM core (gemm_core.cuh).
//
// Right-looking tile algorithm, tile = 128:
//   for k:  diag  : L_kk = chol(A_kk) and Linv_kk = L_kk^-1 (one CTA, latency bound: warp-level
//                   32x32 factor with shuffle pivots + DMMA block updates out of shared memory)
//           panel : A_ik <- A_ik Linv_kk^T          (NT GEMM, K = 128)
//           update: A_ij <- A_ij - A_ik A_jk^T      (NT GEMM, K = 128, lower tiles only)
// Appended rows B (nb x n) ride along as extra row tiles, so B <- B L^-T falls out of the same
// sweep: with B = y^T this is the forward solve of the log-marginal; with the joint matrix
// [[K_aa, .], [K_*a, K_**]] the sweep yields L, V^T = K_*a L^-T and chol(K_** - V^T V) at once.
#include <cstdlib>

#include "gemm_core.cuh"

namespace gpar {

// --------------------------------------------------------------------------------------
// Diagonal tile: factor + invert, entirely in shared memory.
// --------------------------------------------------------------------------------------
constexpr int DLD = 129;  // row stride of the tile in shared memory
constexpr int ILD = 33;   // row stride of a 32x32 inverse block
constexpr int NB32 = TILE / 32;
constexpr int NINV = NB32 * (NB32 + 1) / 2;  // 10 lower blocks
constexpr double REFINE_KAPPA = 1.0e3;  // kappa_inf(L_kk) above which the tile solves are refined
constexpr size_t DIAG_SMEM_BYTES = sizeof(double) * (TILE * DLD + NINV * 32 * ILD + TILE) + 16;

__device__ __forceinline__ int blk(int b, int a) { return b * (b + 1) / 2 + a; }

// Warp-level Cholesky of the 32x32 block at Ls (row stride DLD).  Lane r owns row r in
// registers; finished rows are broadcast through shared memory.  Returns the 1-based index of
// the first non-positive pivot (0 if none) in every lane.
__device__ __forceinline__ int potf2_32(double* Ls, double* rdiag, int lane) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? Ls[lane * DLD + c] : 0.0;
  int bad = 0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int t = 0; t + 3 < c; t += 4) {
      s0 = fma(a[t], Ls[c * DLD + t], s0);
      s1 = fma(a[t + 1], Ls[c * DLD + t + 1], s1);
      s2 = fma(a[t + 2], Ls[c * DLD + t + 2], s2);
      s3 = fma(a[t + 3], Ls[c * DLD + t + 3], s3);
    }
#pragma unroll
    for (int t = (c / 4) * 4; t < c; ++t) s0 = fma(a[t], Ls[c * DLD + t], s0);
    const double v = a[c] - ((s0 + s1) + (s2 + s3));
    double piv = __shfl_sync(0xffffffffu, v, c);
    if (!(piv > 0.0)) {
      if (bad == 0) bad = c + 1;
      piv = 1.0;
    }
    const double rs = rsqrt(piv);
    const double l = (lane == c) ? piv * rs : v * rs;
    a[c] = l;
    if (lane >= c) Ls[lane * DLD + c] = l;
    if (lane == c) rdiag[c] = rs;
    __syncwarp();
  }
  return bad;
}

// Warp-level inverse of the lower-triangular 32x32 block at Ls: lane j solves L x = e_j.
__device__ __forceinline__ void trtri_32(const double* Ls, const double* rdiag, double* inv, int lane) {
  double x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k + 3 < i; k += 4) {
      s0 = fma(-Ls[i * DLD + k], x[k], s0);
      s1 = fma(-Ls[i * DLD + k + 1], x[k + 1], s1);
      s2 = fma(-Ls[i * DLD + k + 2], x[k + 2], s2);
      s3 = fma(-Ls[i * DLD + k + 3], x[k + 3], s3);
    }
#pragma unroll
    for (int k = (i / 4) * 4; k < i; ++k) s0 = fma(-Ls[i * DLD + k], x[k], s0);
    x[i] = ((s0 + s1) + (s2 + s3)) * rdiag[i];
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) inv[i * ILD + lane] = x[i];
}

// acc[t] += sum_k A(gid, k0+tig) * B(k0+tig, t*8+gid) over K (multiple of 4): one 8 x (8*NTN) strip.
template <int NTN, class FA, class FB>
__device__ __forceinline__ void strip_mma(double (&c)[NTN][2], int K, FA a, FB b, int gid, int tig) {
  for (int k0 = 0; k0 < K; k0 += 4) {
    const double av = a(gid, k0 + tig);
#pragma unroll
    for (int t = 0; t < NTN; ++t) {
      const double bv = b(k0 + tig, t * 8 + gid);
      dmma884(c[t][0], c[t][1], av, bv);
    }
  }
}

// Factor one diagonal tile in shared memory (all 256 threads of the CTA): Atile points at
// A[j0][j0]; on return the tile holds L_kk (strictly upper part zeroed), ws_tile its inverse,
// *flag_out the refinement flag, *info_b the first bad pivot (if none was recorded before).
__device__ __forceinline__ void diag_factor_tile(unsigned char* smem_raw, double* __restrict__ Atile, int64_t lda,
                                                 int kb, int64_t j0, double* __restrict__ ws,
                                                 double* __restrict__ flag_out, int32_t* __restrict__ info_b) {
  double* Ls = reinterpret_cast<double*>(smem_raw);
  double* Inv = Ls + TILE * DLD;
  double* rdiag = Inv + NINV * 32 * ILD;
  int* s_bad = reinterpret_cast<int*>(rdiag + TILE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;

  for (int idx = tid; idx < TILE * TILE; idx += 256) {
    const int r = idx >> 7, c = idx & 127;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < kb && c <= r) v = __ldcg(Atile + (int64_t)r * lda + c);
    Ls[r * DLD + c] = v;
  }
  if (tid == 0) *s_bad = 0;
  __syncthreads();

  for (int a = 0; a < NB32; ++a) {
    const int o = 32 * a;
    if (warp == 0) {
      int bad = potf2_32(Ls + o * DLD + o, rdiag + o, lane);
      if (bad && lane == 0 && *s_bad == 0) *s_bad = o + bad;
      __syncwarp();
      trtri_32(Ls + o * DLD + o, rdiag + o, Inv + blk(a, a) * 32 * ILD, lane);
    }
    __syncthreads();
    const int rem = TILE - (o + 32);
    if (rem == 0) break;
    // panel: X = T L_aa^-T, T = rows [o+32, 128) x cols [o, o+32) (in place, a warp owns whole rows).
    // X0 = T inv^T, then one step of iterative refinement against L_aa itself
    // (R = T - X0 L_aa^T, X = X0 + R inv^T) so that the solve stays backward stable when the
    // block is ill-conditioned (explicit inverses alone lose a factor cond(L_aa)).
    {
      const double* inv = Inv + blk(a, a) * 32 * ILD;
      const double* Laa = Ls + o * DLD + o;
      for (int strip = warp; strip < rem / 8; strip += 8) {
        double* T = Ls + (o + 32 + strip * 8) * DLD + o;
        auto fa = [&](int m, int k) { return T[m * DLD + k]; };
        auto finv = [&](int k, int nn) { return inv[nn * ILD + k]; };
        auto flt = [&](int k, int nn) { return (k <= nn) ? Laa[nn * DLD + k] : 0.0; };
        double t0[4][2], x0[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          t0[t][0] = T[gid * DLD + t * 8 + 2 * tig];
          t0[t][1] = T[gid * DLD + t * 8 + 2 * tig + 1];
        }
        strip_mma<4>(x0, 32, fa, finv, gid, tig);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          T[gid * DLD + t * 8 + 2 * tig] = x0[t][0];
          T[gid * DLD + t * 8 + 2 * tig + 1] = x0[t][1];
        }
        __syncwarp();
        double xl[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        strip_mma<4>(xl, 32, fa, flt, gid, tig);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          T[gid * DLD + t * 8 + 2 * tig] = t0[t][0] - xl[t][0];
          T[gid * DLD + t * 8 + 2 * tig + 1] = t0[t][1] - xl[t][1];
        }
        __syncwarp();
        strip_mma<4>(x0, 32, fa, finv, gid, tig);  // x0 += R inv^T
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          T[gid * DLD + t * 8 + 2 * tig] = x0[t][0];
          T[gid * DLD + t * 8 + 2 * tig + 1] = x0[t][1];
        }
      }
    }
    __syncthreads();
    // trailing: T2[m][nn] -= sum_k X[m][k] X[nn][k], lower 8x8 tiles only
    {
      const double* X = Ls + (o + 32) * DLD + o;
      double* T2 = Ls + (o + 32) * DLD + (o + 32);
      const int ntm = rem / 8;
      // work units: (tm, group g of 4 column tiles) with 4*g <= tm
      int unit = 0;
      for (int tm = 0; tm < ntm; ++tm) {
        for (int g = 0; g * 4 <= tm; ++g, ++unit) {
          if ((unit & 7) != warp) continue;
          double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
          const double* Xa = X + (tm * 8) * DLD;
          const double* Xb = X + (g * 32) * DLD;
          strip_mma<4>(
              c, 32, [&](int m, int k) { return Xa[m * DLD + k]; }, [&](int k, int nn) { return Xb[nn * DLD + k]; },
              gid, tig);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int col = g * 32 + t * 8 + 2 * tig;
            if (g * 4 + t <= tm) {
              double* p = T2 + (tm * 8 + gid) * DLD + col;
              p[0] -= c[t][0];
              p[1] -= c[t][1];
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- assemble the off-diagonal blocks of the inverse, block row by block row --------------
  for (int bb = 1; bb < NB32; ++bb) {
    // stage 1: M_ba = sum_{c=a}^{bb-1} L[bb][c] Inv[c][a]  -> stored in Inv[bb][a]
    for (int unit = warp; unit < bb * 4; unit += 8) {
      const int a = unit >> 2, tm = unit & 3;
      double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      for (int cc = a; cc < bb; ++cc) {
        const double* Lb = Ls + (32 * bb + tm * 8) * DLD + 32 * cc;
        const double* Ic = Inv + blk(cc, a) * 32 * ILD;
        strip_mma<4>(
            c, 32, [&](int m, int k) { return Lb[m * DLD + k]; }, [&](int k, int nn) { return Ic[k * ILD + nn]; }, gid,
            tig);
      }
      double* M = Inv + blk(bb, a) * 32 * ILD + (tm * 8) * ILD;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        M[gid * ILD + t * 8 + 2 * tig] = c[t][0];
        M[gid * ILD + t * 8 + 2 * tig + 1] = c[t][1];
      }
    }
    __syncthreads();
    // stage 2: Inv[bb][a] = -Inv[bb][bb] M_ba, in place: a warp owns 8 whole columns.
    for (int unit = warp; unit < bb * 4; unit += 8) {
      const int a = unit >> 2, tn = unit & 3;
      double* M = Inv + blk(bb, a) * 32 * ILD + tn * 8;
      const double* Ib = Inv + blk(bb, bb) * 32 * ILD;
      double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      // out[m][nn] over 4 row tiles: treat row tiles as the "N" strips by transposing roles:
      // out^T[nn][m] = sum_k M^T[nn][k] Ib^T[k][m]  -> A(nn,k) = M[k][nn], B(k,m) = Ib[m][k]
      strip_mma<4>(
          c, 32, [&](int nn, int k) { return M[k * ILD + nn]; }, [&](int k, int m) { return Ib[m * ILD + k]; }, gid,
          tig);
      __syncwarp();
      // c[t][e] = out^T[gid][t*8 + 2*tig + e] = out[m = t*8+2*tig+e][nn = gid]
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        M[(t * 8 + 2 * tig) * ILD + gid] = -c[t][0];
        M[(t * 8 + 2 * tig + 1) * ILD + gid] = -c[t][1];
      }
    }
    __syncthreads();
  }

  // ---- write back L (valid block; its strictly upper part is zeroed so the tile can serve as a
  // GEMM operand) and Linv (dense 128x128, zero padded); kappa_inf(L_kk) decides refinement -----
  double rowL = 0.0, rowI = 0.0;  // thread t < 128 accumulates the abs row sums of row t
  if (tid < kb) {
    for (int c = 0; c <= tid; ++c) {
      rowL += fabs(Ls[tid * DLD + c]);
      rowI += fabs(Inv[blk(tid >> 5, c >> 5) * 32 * ILD + (tid & 31) * ILD + (c & 31)]);
    }
  }
  // max over the block via shared memory (reuse rdiag as scratch after a barrier)
  __syncthreads();
  double* red = rdiag;
  if (tid < TILE) red[tid] = 0.0;
  __syncthreads();
  {
    double mL = rowL, mI = rowI;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mL = fmax(mL, __shfl_xor_sync(0xffffffffu, mL, off));
      mI = fmax(mI, __shfl_xor_sync(0xffffffffu, mI, off));
    }
    if (lane == 0) {
      red[warp] = mL;
      red[8 + warp] = mI;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double mL = 0.0, mI = 0.0;
    for (int i = 0; i < 8; ++i) {
      mL = fmax(mL, red[i]);
      mI = fmax(mI, red[8 + i]);
    }
    const double kappa = mL * mI;
    *flag_out = (kappa > REFINE_KAPPA || !(kappa == kappa)) ? 1.0 : 0.0;
  }
  for (int idx = tid; idx < TILE * TILE; idx += 256) {
    const int r = idx >> 7, c = idx & 127;
    if (r < kb && c < kb) Atile[(int64_t)r * lda + c] = (c <= r) ? Ls[r * DLD + c] : 0.0;
    double v = 0.0;
    if (r < kb && c <= r) v = Inv[blk(r >> 5, c >> 5) * 32 * ILD + (r & 31) * ILD + (c & 31)];
    ws[idx] = v;
  }
  if (tid == 0 && *s_bad != 0 && *info_b == 0) *info_b = static_cast<int32_t>(j0) + *s_bad;
}

__global__ void __launch_bounds__(256, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int64_t strideA, int kt, int nt_total, int64_t n,
                  double* __restrict__ ws, int64_t strideWs, int32_t* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int64_t j0 = (int64_t)kt * TILE;
  const int kb = static_cast<int>(min64(TILE, n - j0));
  double* wsb = ws + (int64_t)b * strideWs;
  diag_factor_tile(smem_raw, A + (int64_t)b * strideA + j0 * lda + j0, lda, kb, j0, wsb + (int64_t)kt * TILE * TILE,
                   wsb + (int64_t)nt_total * TILE * TILE + kt, info + b);
}

// --------------------------------------------------------------------------------------
// T (rows x kb) <- T Linv^T for every 128-row tile of a row block.
// --------------------------------------------------------------------------------------
// X = T L_kk^-T for one 128-row tile: X0 = T Linv^T; when the diagonal block is flagged
// ill-conditioned, one refinement step R = T - X0 L_kk^T, X = X0 + R Linv^T (X0 parked in a
// per-CTA scratch tile) restores backward stability.
__device__ __forceinline__ void tile_solve(GemmStage* stages, double* __restrict__ T, int64_t ldt, int valid, int kb,
                                           const double* __restrict__ Linv, const double* __restrict__ Lkk,
                                           int64_t ldl, bool refine, double* __restrict__ scratch) {
  Acc acc;
  acc_zero(acc);
  gemm_nt_mainloop(stages, T, ldt, valid, Linv, TILE, kb, kb, acc);
  if (!refine) {
    store_tile<0>(T, ldt, valid, kb, acc, false);
    return;
  }
  store_tile<0>(scratch, TILE, valid, kb, acc, false);  // X0
  __threadfence();
  __syncthreads();
  acc_zero(acc);
  gemm_nt_mainloop(stages, scratch, TILE, valid, Lkk, ldl, kb, kb, acc);  // X0 L_kk^T
  store_tile<1>(T, ldt, valid, kb, acc, false);                            // T <- R = T - X0 L_kk^T
  __threadfence();
  __syncthreads();
  acc_zero(acc);
  gemm_nt_mainloop(stages, T, ldt, valid, Linv, TILE, kb, kb, acc);  // R Linv^T
  store_tile_add(T, ldt, scratch, TILE, valid, kb, acc);              // T <- X0 + R Linv^T
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
trsm_tile_kernel(double* __restrict__ T1, int64_t ldt1, int64_t rows1, int64_t strideT1, int nt1,
                 double* __restrict__ T2, int64_t ldt2, int64_t rows2, int64_t strideT2, int kb,
                 const double* __restrict__ Linv, int64_t strideW, const double* __restrict__ Lkk, int64_t ldl,
                 int64_t strideL, const double* __restrict__ flag, double* __restrict__ scratch,
                 int64_t strideScratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  int ti = blockIdx.x;
  const int b = blockIdx.y;
  double* T;
  int64_t ldt, rows;
  if (ti < nt1) {
    T = T1 + (int64_t)b * strideT1; ldt = ldt1; rows = rows1;
  } else {
    ti -= nt1;
    T = T2 + (int64_t)b * strideT2; ldt = ldt2; rows = rows2;
  }
  T += (int64_t)ti * TILE * ldt;
  const int valid = static_cast<int>(min64(TILE, rows - (int64_t)ti * TILE));
  const bool refine = flag[(int64_t)b * strideW] != 0.0;
  tile_solve(stages, T, ldt, valid, kb, Linv + (int64_t)b * strideW, Lkk + (int64_t)b * strideL, ldl, refine,
             scratch + (int64_t)b * strideScratch + (int64_t)blockIdx.x * TILE * TILE);
}

// --------------------------------------------------------------------------------------
// C(ti, tj) -= Aop(ti) Bop(tj)^T, K columns.  lower != 0: only tiles tj <= ti, and only
// col <= row on the diagonal tiles.
// --------------------------------------------------------------------------------------
struct SubArgs {
  // primary row space (row tiles [0, nt_rows1)): symmetric/lower part when lower != 0
  double* C; int64_t ldc; int64_t c_rows; int64_t c_cols; int64_t strideC;
  const double* Aop; int64_t lda; int64_t strideA;
  const double* Bop; int64_t ldb; int64_t strideB;
  int K; int lower; int nt_rows1;
  // secondary row space (appended rows, row tiles >= nt_rows1): all column tiles
  double* C2; int64_t ldc2; int64_t c_rows2; int64_t strideC2;
  const double* Aop2; int64_t lda2; int64_t strideA2;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_sub_kernel(const SubArgs p) {
  const int tj = blockIdx.x, b = blockIdx.z;
  int ti = blockIdx.y;
  const bool second = ti >= p.nt_rows1;
  if (!second && p.lower && tj > ti) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  const int cols = static_cast<int>(min64(TILE, p.c_cols - (int64_t)tj * TILE));
  const double* Bp = p.Bop + (int64_t)b * p.strideB + (int64_t)tj * TILE * p.ldb;
  const double* Ap;
  double* C;
  int64_t lda, ldc;
  int rows;
  bool lower_diag = false;
  if (!second) {
    rows = static_cast<int>(min64(TILE, p.c_rows - (int64_t)ti * TILE));
    lda = p.lda; ldc = p.ldc;
    Ap = p.Aop + (int64_t)b * p.strideA + (int64_t)ti * TILE * lda;
    C = p.C + (int64_t)b * p.strideC + (int64_t)ti * TILE * ldc + (int64_t)tj * TILE;
    lower_diag = p.lower && ti == tj;
  } else {
    ti -= p.nt_rows1;
    rows = static_cast<int>(min64(TILE, p.c_rows2 - (int64_t)ti * TILE));
    lda = p.lda2; ldc = p.ldc2;
    Ap = p.Aop2 + (int64_t)b * p.strideA2 + (int64_t)ti * TILE * lda;
    C = p.C2 + (int64_t)b * p.strideC2 + (int64_t)ti * TILE * ldc + (int64_t)tj * TILE;
  }
  Acc acc;
  acc_zero(acc);
  gemm_nt_mainloop(stages, Ap, lda, rows, Bp, p.ldb, cols, p.K, acc);
  store_tile<1>(C, ldc, rows, cols, acc, lower_diag);
}

// --------------------------------------------------------------------------------------
// B (nb x n) <- B L^-T with L already factored: every CTA owns one 128-row tile of B and sweeps
// the column tiles left to right (left-looking), so no inter-CTA dependency exists.
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1)
trsm_rows_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ ws,
                 double* __restrict__ B, int64_t ldb, int64_t nb, double* __restrict__ scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  const int ti = blockIdx.x;
  double* Brow = B + (int64_t)ti * TILE * ldb;
  const int valid = static_cast<int>(min64(TILE, nb - (int64_t)ti * TILE));
  const int nt = static_cast<int>((n + TILE - 1) / TILE);
  const double* flags = ws + (int64_t)nt * TILE * TILE;
  for (int j = 0; j < nt; ++j) {
    const int kb = static_cast<int>(min64(TILE, n - (int64_t)j * TILE));
    if (j > 0) {
      Acc acc;
      acc_zero(acc);
      gemm_nt_mainloop(stages, Brow, ldb, valid, L + (int64_t)j * TILE * ldl, ldl, kb, j * TILE, acc);
      store_tile<1>(Brow + (int64_t)j * TILE, ldb, valid, kb, acc, false);
      __threadfence();
      __syncthreads();
    }
    tile_solve(stages, Brow + (int64_t)j * TILE, ldb, valid, kb, ws + (int64_t)j * TILE * TILE,
               L + (int64_t)j * TILE * ldl + (int64_t)j * TILE, ldl, flags[j] != 0.0,
               scratch + (int64_t)ti * TILE * TILE);
    __threadfence();
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------------
// v2: persistent left-looking tile-dataflow Cholesky.  One CTA per SM pulls tile tasks (i, j)
// from a global ticket counter in column-major order (every dependency of a task has a smaller
// ticket, so a spinning CTA only ever waits on work that is already running or done: no
// deadlock, no co-residency requirement).  Task (i, j):
//     acc = sum_{k<j} L_ik L_jk^T   one long-K DMMA GEMM, waiting per 128-column block on the
//                                   "ready" flags of the tiles it streams
//     T   = A_ij - acc              accumulators meet the tile exactly once (HBM traffic n^2 per
//                                   sweep instead of n^3 / (3 * 128) for the right-looking form)
//     i == j:  L_jj = chol(T), Linv_jj, refinement flag   (diag_factor_tile)
//     i >  j:  L_ij = T L_jj^-T                           (tile_solve, after L_jj is ready)
// Appended row tiles (B) and batched matrices are just more tasks of the same list.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_ready(const int* flag) {
  while (ld_acquire(flag) == 0) __nanosleep(40);
}

// gemm_nt_mainloop over K = 128 * ktiles with per-k-tile dependency waits on readyA[kt], readyB[kt].
__device__ __forceinline__ void gemm_nt_mainloop_dep(GemmStage* stages, const double* __restrict__ Ap, int64_t lda,
                                                     int validA, const double* __restrict__ Bp, int64_t ldb,
                                                     int validB, int K, Acc& acc, const int* readyA,
                                                     const int* readyB) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const int nchunks = K / BK;
  constexpr int CPT = TILE / BK;  // chunks per k-tile
  auto issue = [&](int nc) {
    if (nc % CPT == 0) {
      wait_ready(readyA + nc / CPT);
      if (readyB != readyA) wait_ready(readyB + nc / CPT);
    }
    load_chunk(stages[nc % STAGES], Ap, lda, validA, Bp, ldb, validB, nc * BK, K);
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) issue(s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    mma_chunk(stages[c % STAGES], acc, wm, wn, gid, tig);
    const int nc = c + STAGES - 1;
    if (nc < nchunks) issue(nc);
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncthreads();
}

constexpr size_t DF_SMEM_BYTES = DIAG_SMEM_BYTES > GEMM_SMEM_BYTES ? DIAG_SMEM_BYTES : GEMM_SMEM_BYTES;
constexpr int DF_POOL_TILES = 192;  // scratch tiles for the persistent grid (>= SM count)

struct DfArgs {
  double* A; int64_t lda; int64_t n; int64_t strideA;
  double* B; int64_t ldb; int64_t nb; int64_t strideB;
  int batch; int nt; int nbt; int total_tasks;
  double* ws; int64_t strideWs;   // per matrix: inverse tiles, refine flags (see ws_* helpers)
  double* pool;                   // gridDim.x scratch tiles
  int32_t* info;
  int* ticket; int* ready;        // ready[(b * (nt + nbt) + i) * nt + j]
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) potrf_dataflow_kernel(const DfArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_task;
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  const int tid = threadIdx.x;
  const int rows_total = p.nt + p.nbt;
  double* scratch = p.pool + (int64_t)blockIdx.x * TILE * TILE;
  for (;;) {
    if (tid == 0) s_task = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= p.total_tasks) break;
    // decode ticket -> (column j, matrix b, row tile i); columns outermost, diagonal task first
    int j = 0, rem = t;
    for (;; ++j) {
      const int per_col = p.batch * (rows_total - j);
      if (rem < per_col) break;
      rem -= per_col;
    }
    const int rows_in_col = rows_total - j;
    const int b = rem / rows_in_col;
    const int i = j + rem % rows_in_col;

    double* Ab = p.A + (int64_t)b * p.strideA;
    const int kb = static_cast<int>(min64(TILE, p.n - (int64_t)j * TILE));
    const double* rowj = Ab + (int64_t)j * TILE * p.lda;
    double* rowi;
    int64_t ldi;
    int valid;
    if (i < p.nt) {
      rowi = Ab + (int64_t)i * TILE * p.lda; ldi = p.lda;
      valid = static_cast<int>(min64(TILE, p.n - (int64_t)i * TILE));
    } else {
      rowi = p.B + (int64_t)b * p.strideB + (int64_t)(i - p.nt) * TILE * p.ldb; ldi = p.ldb;
      valid = static_cast<int>(min64(TILE, p.nb - (int64_t)(i - p.nt) * TILE));
    }
    int* ready_b = p.ready + (int64_t)b * rows_total * p.nt;
    const int* ready_i = ready_b + (int64_t)i * p.nt;
    const int* ready_j = ready_b + (int64_t)j * p.nt;
    double* wsb = p.ws + (int64_t)b * p.strideWs;
    double* T = rowi + (int64_t)j * TILE;

    if (j > 0) {
      Acc acc;
      acc_zero(acc);
      gemm_nt_mainloop_dep(stages, rowi, ldi, valid, rowj, p.lda, kb, j * TILE, acc, ready_i, ready_j);
      store_tile<1>(T, ldi, valid, kb, acc, i == j);
      __threadfence();
      __syncthreads();
    }
    if (i == j) {
      diag_factor_tile(smem_raw, T, p.lda, kb, (int64_t)j * TILE, wsb + (int64_t)j * TILE * TILE,
                       wsb + (int64_t)p.nt * TILE * TILE + j, p.info + b);
    } else {
      wait_ready(ready_j + j);
      const bool refine = __ldcg(wsb + (int64_t)p.nt * TILE * TILE + j) != 0.0;
      tile_solve(stages, T, ldi, valid, kb, wsb + (int64_t)j * TILE * TILE, rowj + (int64_t)j * TILE, p.lda, refine,
                 scratch);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(ready_b + (int64_t)i * p.nt + j, 1);
  }
}

static void set_smem_attrs() {
  static bool done = false;
  if (done) return;
  cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM_BYTES);
  cudaFuncSetAttribute(trsm_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
  cudaFuncSetAttribute(gemm_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
  cudaFuncSetAttribute(trsm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
  cudaFuncSetAttribute(potrf_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM_BYTES);
  done = true;
}

}  // namespace gpar

using namespace gpar;

// Workspace layout per matrix (doubles): [nt inverse tiles][nt flags, padded to even]
// [(nt + nbt) scratch tiles for the refined solves].
static int64_t ws_flags_off(int64_t nt) { return nt * TILE * TILE; }
static int64_t ws_scratch_off(int64_t nt) { return nt * TILE * TILE + ((nt + 1) & ~(int64_t)1); }
static int64_t ws_stride(int64_t nt, int64_t nbt) { return ws_scratch_off(nt) + (nt + nbt) * TILE * TILE; }

// After the per-matrix regions: [DF_POOL_TILES scratch tiles][int region: ticket (2 ints) + ready flags].
static int64_t ws_ready_ints(int64_t nt, int64_t nbt, int64_t batch) { return 2 + batch * (nt + nbt) * nt; }

extern "C" size_t gpar_potrf_workspace_bytes(int64_t n, int64_t nb, int64_t batch) {
  if (n <= 0 || batch <= 0) return 0;
  const int64_t nt = (n + TILE - 1) / TILE, nbt = nb > 0 ? (nb + TILE - 1) / TILE : 0;
  const int64_t doubles = batch * ws_stride(nt, nbt) + (int64_t)DF_POOL_TILES * TILE * TILE +
                          (ws_ready_ints(nt, nbt, batch) + 1) / 2 + 2;
  return (size_t)doubles * sizeof(double);
}

extern "C" size_t gpar_trsm_rows_scratch_bytes(int64_t nb) {
  if (nb <= 0) return 0;
  return (size_t)((nb + TILE - 1) / TILE) * TILE * TILE * sizeof(double);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int gpar_potrf(double* A, int64_t lda, int64_t n, int64_t strideA, double* B, int64_t ldb, int64_t nb,
                          int64_t strideB, int64_t batch, double* ws, int32_t* info, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!A || !aligned16(A)) { set_error("gpar_potrf: A null or not 16-byte aligned"); return -1; }
  if (lda < n || (lda & 1)) { set_error("gpar_potrf: lda must be even and >= n"); return -2; }
  if (n < 0) return -3;
  if (batch > 1 && (strideA & 1)) { set_error("gpar_potrf: strideA must be even"); return -4; }
  if (nb > 0 && (!B || !aligned16(B) || ldb < n || (ldb & 1) || (batch > 1 && (strideB & 1)))) {
    set_error("gpar_potrf: bad appended row block");
    return -5;
  }
  if (batch <= 0) return -9;
  if (!ws || !aligned16(ws)) { set_error("gpar_potrf: bad workspace"); return -10; }
  if (!info) return -11;
  if (batch > 65535) { set_error("gpar_potrf: batch > 65535"); return -9; }
  set_smem_attrs();
  cudaMemsetAsync(info, 0, sizeof(int32_t) * batch, stream);
  if (n == 0) return 0;
  const int nt = (int)((n + TILE - 1) / TILE);
  const int nbt = nb > 0 ? (int)((nb + TILE - 1) / TILE) : 0;
  const int64_t strideWs = ws_stride(nt, nbt);
  double* scratch = ws + ws_scratch_off(nt);
  static const bool use_v1 = (getenv("GPAR_POTRF_V1") != nullptr);
  if (!use_v1) {
    static int num_sms = 0;
    if (num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    DfArgs p;
    p.A = A; p.lda = lda; p.n = n; p.strideA = strideA;
    p.B = B; p.ldb = ldb; p.nb = nb; p.strideB = strideB;
    p.batch = (int)batch; p.nt = nt; p.nbt = nbt;
    int64_t total = 0;
    for (int j = 0; j < nt; ++j) total += batch * (int64_t)(nt + nbt - j);
    if (total > 0x7fffffff) { set_error("gpar_potrf: too many tile tasks"); return -9; }
    p.total_tasks = (int)total;
    p.ws = ws; p.strideWs = strideWs;
    p.pool = ws + batch * strideWs;
    p.info = info;
    int* ints = reinterpret_cast<int*>(p.pool + (int64_t)DF_POOL_TILES * TILE * TILE);
    p.ticket = ints; p.ready = ints + 2;
    cudaMemsetAsync(ints, 0, sizeof(int) * (size_t)ws_ready_ints(nt, nbt, batch), stream);
    int grid = num_sms < DF_POOL_TILES ? num_sms : DF_POOL_TILES;
    if ((int64_t)grid > total) grid = (int)total;
    potrf_dataflow_kernel<<<grid, GEMM_THREADS, DF_SMEM_BYTES, stream>>>(p);
    return check_launch("gpar_potrf");
  }
  for (int k = 0; k < nt; ++k) {
    const int64_t j0 = (int64_t)k * TILE;
    const int kb = (int)((n - j0 < TILE) ? (n - j0) : TILE);
    potrf_diag_kernel<<<(unsigned)batch, 256, DIAG_SMEM_BYTES, stream>>>(A, lda, strideA, k, nt, n, ws, strideWs,
                                                                       info);
    const int64_t below = n - (j0 + TILE);
    const double* Linv = ws + (int64_t)k * TILE * TILE;
    const int ntr = below > 0 ? (int)((below + TILE - 1) / TILE) : 0;
    if (ntr + nbt > 0) {
      dim3 g((unsigned)(ntr + nbt), (unsigned)batch);
      trsm_tile_kernel<<<g, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(
          below > 0 ? A + (j0 + TILE) * lda + j0 : A, lda, below > 0 ? below : 0, strideA, ntr,
          nb > 0 ? B + j0 : B, ldb, nb, strideB, kb, Linv, strideWs, A + j0 * lda + j0, lda, strideA,
          ws + ws_flags_off(nt) + k, scratch, strideWs);
    }
    if (below > 0) {
      SubArgs p;
      p.C = A + (j0 + TILE) * lda + (j0 + TILE); p.ldc = lda; p.c_rows = below; p.c_cols = below; p.strideC = strideA;
      p.Aop = A + (j0 + TILE) * lda + j0; p.lda = lda; p.strideA = strideA;
      p.Bop = p.Aop; p.ldb = lda; p.strideB = strideA;
      p.K = kb; p.lower = 1; p.nt_rows1 = ntr;
      p.C2 = nb > 0 ? B + (j0 + TILE) : nullptr; p.ldc2 = ldb; p.c_rows2 = nb; p.strideC2 = strideB;
      p.Aop2 = nb > 0 ? B + j0 : nullptr; p.lda2 = ldb; p.strideA2 = strideB;
      gemm_sub_kernel<<<dim3((unsigned)ntr, (unsigned)(ntr + nbt), (unsigned)batch), GEMM_THREADS, GEMM_SMEM_BYTES,
                        stream>>>(p);
    }
  }
  return check_launch("gpar_potrf");
}

extern "C" int gpar_trsm_rows(const double* L, int64_t ldl, int64_t n, const double* ws, double* B, int64_t ldb,
                              int64_t nb, double* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!L || !aligned16(L) || (ldl & 1) || ldl < n) { set_error("gpar_trsm_rows: bad L"); return -1; }
  if (!ws || !aligned16(ws)) return -4;
  if (nb <= 0 || n <= 0) return 0;
  if (!B || !aligned16(B) || (ldb & 1) || ldb < n) { set_error("gpar_trsm_rows: bad B"); return -5; }
  if (!scratch || !aligned16(scratch)) { set_error("gpar_trsm_rows: bad scratch"); return -8; }
  set_smem_attrs();
  unsigned g = (unsigned)((nb + TILE - 1) / TILE);
  trsm_rows_kernel<<<g, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(L, ldl, n, ws, B, ldb, nb, scratch);
  return check_launch("gpar_trsm_rows");
}

extern "C" int gpar_syrk_sub(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw,
                             int64_t k, int64_t strideW, int64_t batch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!C || !aligned16(C) || (ldc & 1) || ldc < n) { set_error("gpar_syrk_sub: bad C"); return -1; }
  if (!W || !aligned16(W) || (ldw & 1) || ldw < k) { set_error("gpar_syrk_sub: bad W"); return -5; }
  if (batch > 1 && ((strideC & 1) || (strideW & 1))) { set_error("gpar_syrk_sub: odd batch stride"); return -4; }
  if (n <= 0 || k <= 0 || batch <= 0) return 0;
  if (batch > 65535) { set_error("gpar_syrk_sub: batch > 65535"); return -9; }
  set_smem_attrs();
  const unsigned nt = (unsigned)((n + TILE - 1) / TILE);
  SubArgs p;
  p.C = C; p.ldc = ldc; p.c_rows = n; p.c_cols = n; p.strideC = strideC;
  p.Aop = W; p.lda = ldw; p.strideA = strideW;
  p.Bop = W; p.ldb = ldw; p.strideB = strideW;
  p.K = (int)k; p.lower = 1; p.nt_rows1 = (int)nt;
  p.C2 = nullptr; p.ldc2 = 0; p.c_rows2 = 0; p.strideC2 = 0; p.Aop2 = nullptr; p.lda2 = 0; p.strideA2 = 0;
  gemm_sub_kernel<<<dim3(nt, nt, (unsigned)batch), GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
  return check_launch("gpar_syrk_sub");
}
