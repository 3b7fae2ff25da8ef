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
constexpr int DLD = 132;  // row stride of the tile in shared memory: = 4 (mod 16) doubles makes the DMMA
                          // fragment loads bank-conflict free; rows stay 16-byte aligned for cp.async
constexpr double REFINE_KAPPA = 1.0e3;  // kappa_inf(L_kk) above which the tile solves are refined
constexpr size_t DIAG_SMEM_BYTES = sizeof(double) * (TILE * DLD + TILE + 32) + 16;

// Factor one diagonal tile in shared memory (all 256 threads of the CTA).  Atile points at
// A[j0][j0]; on return the tile holds L_kk (strictly upper part zeroed), ws its inverse,
// *flag_out the refinement flag, *info_b the first bad pivot (if none was recorded before).
//
// Right-looking sweep over 16 column blocks of width 8 on ONE 128 x 129 shared array:
//   lower triangle  = L (in place);
//   strict upper    = the appended identity rows of the sweep, i.e. (I L^-T)[r][c] = Linv[c][r]
//                     (the diagonal of Linv is rdiag = 1 / L_rr), so the inverse costs no extra
//                     phase and no extra storage.
// Per block: (a) warps 0-3: every thread factors the 8x8 diagonal block redundantly in registers
// (chain = 8 x (rsqrt + mul + fma), no communication), (b) thread r solves row r against it by
// substitution (backward stable: no refinement needed), (c) all warps apply the rank-8 trailing
// update with DMMA (independent 8x8 sub-tiles, no accumulate chains).
__device__ __forceinline__ void diag_factor_tile(unsigned char* smem_raw, double* __restrict__ Atile, int64_t lda,
                                                 int kb, int64_t j0, double* __restrict__ ws,
                                                 double* __restrict__ flag_out, int32_t* __restrict__ info_b,
                                                 long long* prof = nullptr) {
#define GPAR_PROF(i) do { if (prof && threadIdx.x == 0) prof[i] = clock64(); } while (0)
  GPAR_PROF(0);
  double* Ls = reinterpret_cast<double*>(smem_raw);
  double* rdiag = Ls + TILE * DLD;
  double* red = rdiag + TILE;  // 32 doubles of reduction scratch
  int* s_bad = reinterpret_cast<int*>(red + 32);
  const int tid = threadIdx.x, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int warp = canonical_warp();

  // ---- load the lower triangle with cp.async (16-byte, all in flight), then fix up: zeros above
  // the diagonal, identity padding beyond kb ------------------------------------------------------
  {
    const bool vec_ok = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(Atile) & 15) == 0);
    if (vec_ok) {
      // zero-filling copies: 16 B below the diagonal, 8 B on it, 0 B (pure zero fill) above / beyond kb
#pragma unroll 8
      for (int q = 0; q < 32; ++q) {
        const int idx = tid + q * 256;  // 0..8191: row r, double2 column c
        const int r = idx >> 6, c = (idx & 63) * 2;
        const int bytes = (r < kb) ? ((c + 1 <= r) ? 16 : ((c == r) ? 8 : 0)) : 0;
        cp_async16(&Ls[r * DLD + c], bytes ? (Atile + (int64_t)r * lda + c) : Atile, bytes);
      }
      cp_async_commit();
      cp_async_wait<0>();
    } else {
      for (int idx = tid; idx < TILE * TILE; idx += 256) {
        const int r = idx >> 7, c = idx & 127;
        Ls[r * DLD + c] = (r < kb && c <= r) ? __ldcg(Atile + (int64_t)r * lda + c) : 0.0;
      }
    }
    __syncthreads();
    if (tid >= kb && tid < TILE) Ls[tid * DLD + tid] = 1.0;  // identity padding
  }
  if (tid == 0) *s_bad = 0;
  __syncthreads();
  GPAR_PROF(1);

  // Warp 0: factor the 8x8 diagonal block kblk in registers (all lanes redundantly: the chain is
  // 8 x (rsqrt + mul + fma) with no communication); lanes 0-7 publish row m of L8 and 1 / L_mm.
  auto factor_block = [&](int kblk) {
    const int c0 = 8 * kblk;
    double d[8][8], rd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) d[i][j] = Ls[(c0 + i) * DLD + c0 + j];
    int bad = 0;
    if (kblk == 1) GPAR_PROF(13);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double piv = d[j][j];
      if (!(piv > 0.0)) {
        if (bad == 0) bad = j + 1;
        piv = 1.0;
      }
      const double rs = rsqrt(piv);
      rd[j] = rs;
      d[j][j] = piv * rs;
#pragma unroll
      for (int i = j + 1; i < 8; ++i) d[i][j] *= rs;
#pragma unroll
      for (int i = j + 1; i < 8; ++i)
#pragma unroll
        for (int l = j + 1; l <= i; ++l) d[i][l] = fma(-d[i][j], d[l][j], d[i][l]);
    }
    if (kblk == 1) GPAR_PROF(14);
    // publish without divergent branches: predicated stores, lane mm writes row mm
#pragma unroll
    for (int mm = 0; mm < 8; ++mm) {
#pragma unroll
      for (int j = 0; j <= mm; ++j) {
        const double v = d[mm][j];
        if (lane == mm) Ls[(c0 + mm) * DLD + c0 + j] = v;
      }
      const double rv = rd[mm];
      if (lane == mm) rdiag[c0 + mm] = rv;
    }
    if (bad && lane == 0 && *s_bad == 0) *s_bad = c0 + bad;
  };

  if (warp == 0) factor_block(0);
#pragma unroll 1
  for (int k = 0; k < TILE / 8; ++k) {
    const int c0 = 8 * k;
    __syncthreads();  // L8(k) published; trailing update of step k-1 complete
    if (k == 0 || k == 8) GPAR_PROF(2 + (k ? 6 : 0));
    if (warp < 4) {
      // row r = tid: x L8^T = t by substitution against the published block (backward stable)
      double d[8][8], rd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rd[i] = rdiag[c0 + i];
#pragma unroll
        for (int j = 0; j < i; ++j) d[i][j] = Ls[(c0 + i) * DLD + c0 + j];
      }
      const int r = tid;
      const bool in_block = (r >= c0 && r < c0 + 8);
      double x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = in_block ? ((r - c0 == j) ? 1.0 : 0.0) : Ls[r * DLD + c0 + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double sacc = x[j];
#pragma unroll
        for (int i = 0; i < j; ++i) sacc = fma(-x[i], d[j][i], sacc);
        x[j] = sacc * rd[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)  // in-block rows keep L8 in their lower part: only the strict upper is theirs
        if (!in_block || j > r - c0) Ls[r * DLD + c0 + j] = x[j];
    }
    if (k == 0 || k == 8) GPAR_PROF(3 + (k ? 6 : 0));
    __syncthreads();
    if (k == 0 || k == 8) GPAR_PROF(4 + (k ? 6 : 0));
    if (k == TILE / 8 - 1) break;
    // rank-8 trailing update C -= X X^T.  Column tile ct > k meets row tiles rt in [0, k] (inverse
    // rows, strict-upper region) and rt in [ct, 15] (L rows).  Warp 0 takes the next diagonal
    // sub-tile and factors it right away (look-ahead); warps 1-7 share the other row tiles and
    // process four column tiles at a time (independent accumulators).
    if (warp == 0) {
      const int row = (k + 1) * 8 + gid;
      const double a0 = Ls[row * DLD + c0 + tig], a1 = Ls[row * DLD + c0 + 4 + tig];
      double* pc = Ls + row * DLD + (k + 1) * 8 + 2 * tig;
      double c0v = pc[0], c1v = pc[1];
      dmma884(c0v, c1v, -a0, a0);
      dmma884(c0v, c1v, -a1, a1);
      if (2 * tig <= gid) pc[0] = c0v;
      if (2 * tig + 1 <= gid) pc[1] = c1v;
      __syncwarp();
      if (k == 0 || k == 8) GPAR_PROF(5 + (k ? 6 : 0));
      factor_block(k + 1);
      if (k == 0 || k == 8) GPAR_PROF(6 + (k ? 6 : 0));
    } else {
      if (prof && k == 0 && tid == 32) prof[15] = clock64();
      int unit = 0;
      for (int rt = 0; rt < TILE / 8; ++rt) {
        if (rt == k + 1) continue;
        const int ct_lo = k + 1, ct_hi = (rt <= k) ? (TILE / 8 - 1) : rt;  // inclusive
        const int nunits = (ct_hi - ct_lo + 4) / 4;
        // does this warp own any (rt, batch) unit?  units are dealt round-robin to warps 1..7
        const int first = (warp - 1 - unit % 7 + 7) % 7;  // offset of my first unit within this row tile
        const int ubase = unit;
        unit += nunits;
        if (first >= nunits) continue;
        (void)ubase;
        const int row = rt * 8 + gid;
        double a0, a1;
        if (rt == k) {  // inverse rows of the current block: strict upper stored, diagonal = rdiag
          a0 = (tig > gid) ? Ls[row * DLD + c0 + tig] : ((tig == gid) ? rdiag[row] : 0.0);
          a1 = (tig + 4 > gid) ? Ls[row * DLD + c0 + 4 + tig] : ((tig + 4 == gid) ? rdiag[row] : 0.0);
        } else {
          a0 = Ls[row * DLD + c0 + tig];
          a1 = Ls[row * DLD + c0 + 4 + tig];
        }
        a0 = -a0;
        a1 = -a1;
        for (int ctb = ct_lo + 4 * first; ctb <= ct_hi; ctb += 28) {
          double b0[4], b1[4], cv[4][2];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ct = min(ctb + u, ct_hi);
            b0[u] = Ls[(ct * 8 + gid) * DLD + c0 + tig];
            b1[u] = Ls[(ct * 8 + gid) * DLD + c0 + 4 + tig];
            const double* pc = Ls + row * DLD + ct * 8 + 2 * tig;
            cv[u][0] = pc[0];
            cv[u][1] = pc[1];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma884(cv[u][0], cv[u][1], a0, b0[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma884(cv[u][0], cv[u][1], a1, b1[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ct = ctb + u;
            if (ct > ct_hi) continue;
            double* pc = Ls + row * DLD + ct * 8 + 2 * tig;
            if (rt == ct) {  // diagonal sub-tile: its strict upper part belongs to the inverse rows
              if (2 * tig <= gid) pc[0] = cv[u][0];
              if (2 * tig + 1 <= gid) pc[1] = cv[u][1];
            } else {
              pc[0] = cv[u][0];
              pc[1] = cv[u][1];
            }
          }
        }
      }
      if (prof && k == 0 && tid == 32) prof[16] = clock64();
    }
  }
  GPAR_PROF(19);

  // ---- kappa_inf(L_kk) = ||L||_inf ||Linv||_inf decides refinement: two threads per row ------
  {
    const int r = tid >> 1, h = tid & 1;
    double rowL = 0.0, rowI = 0.0;
    if (r < kb) {
      // row r of L: columns [0, r]; row r of Linv: Linv[r][c] = Ls[c][r] for c < r, rdiag[r] at c = r
      const int lo = h ? (r + 1) / 2 : 0, hi = h ? r : (r + 1) / 2;
#pragma unroll 8
      for (int c = lo; c < hi; ++c) {
        rowL += fabs(Ls[r * DLD + c]);
        rowI += fabs(Ls[c * DLD + r]);
      }
      if (h) {
        rowL += fabs(Ls[r * DLD + r]);
        rowI += fabs(rdiag[r]);
      }
    }
    rowL += __shfl_xor_sync(0xffffffffu, rowL, 1);
    rowI += __shfl_xor_sync(0xffffffffu, rowI, 1);
#pragma unroll
    for (int off = 16; off > 1; off >>= 1) {
      rowL = fmax(rowL, __shfl_xor_sync(0xffffffffu, rowL, off));
      rowI = fmax(rowI, __shfl_xor_sync(0xffffffffu, rowI, off));
    }
    if (lane == 0) {
      red[warp] = rowL;
      red[8 + warp] = rowI;
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
  }
  // ---- write back L and Linv (row major, ld = 128).  Only the lower triangles are stored, plus
  // zeros in the strictly upper part of each 32x32 diagonal block: every consumer (MODE 2 GEMMs
  // skip whole 32-column halves per k-chunk, backsolve starts at the 8x8 diagonal block) stays
  // inside that region, so nothing else is ever read. -----------------------------------------------
  {
    const bool vec_ok = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(Atile) & 15) == 0);
#pragma unroll 4
    for (int idx = tid; idx < TILE * (TILE / 2); idx += 256) {  // L: row-wise, conflict free
      const int r = idx >> 6, c = (idx & 63) * 2;
      if (r >= kb || c > (r | 31) || c >= kb) continue;  // keep zeros in the 32x32 diagonal blocks
      double2 lv = make_double2(0.0, 0.0);
      if (c <= r) lv.x = Ls[r * DLD + c];
      if (c + 1 <= r) lv.y = Ls[r * DLD + c + 1];
      double* dst = Atile + (int64_t)r * lda + c;
      if (vec_ok && c + 1 < kb) {
        *reinterpret_cast<double2*>(dst) = lv;
      } else {
        dst[0] = lv.x;
        if (c + 1 < kb) dst[1] = lv.y;
      }
    }
    // Linv[r][c] = Ls[c][r] (c < r): transposed read.  A warp moves an 8 (r) x 4 (c) block per step
    // with lanes laid out so that (4 c + r) mod 16 is distinct within each half warp (conflict free
    // at stride 132); the 4 consecutive c of a row form one 32-byte sector of the output.
    {
      const int rr = (lane & 3) + 4 * (lane >> 4), cc = (lane >> 2) & 3;
#pragma unroll 4
      for (int blk = warp; blk < (TILE / 8) * (TILE / 4); blk += 8) {
        const int R0 = (blk >> 5) * 8, C0 = (blk & 31) * 4;
        if (C0 > (R0 | 31)) continue;
        const int r = R0 + rr, c = C0 + cc;
        if (r >= kb) continue;
        double v = 0.0;
        if (c < r) v = Ls[c * DLD + r]; else if (c == r) v = rdiag[r];
        ws[r * TILE + c] = v;
      }
    }
  }
  if (tid == 0 && *s_bad != 0 && *info_b == 0) *info_b = static_cast<int32_t>(j0) + *s_bad;
  GPAR_PROF(20);
#undef GPAR_PROF
}

__global__ void __launch_bounds__(256, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int64_t strideA, int kt, int nt_total, int64_t n,
                  double* __restrict__ ws, int64_t strideWs, int32_t* __restrict__ info, long long* prof) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int64_t j0 = (int64_t)kt * TILE;
  const int kb = static_cast<int>(min64(TILE, n - j0));
  double* wsb = ws + (int64_t)b * strideWs;
  diag_factor_tile(smem_raw, A + (int64_t)b * strideA + j0 * lda + j0, lda, kb, j0, wsb + (int64_t)kt * TILE * TILE,
                   wsb + (int64_t)nt_total * TILE * TILE + kt, info + b, prof);
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
  gemm_nt_mainloop<2>(stages, T, ldt, valid, Linv, TILE, kb, kb, acc);
  if (!refine) {
    store_tile<0>(T, ldt, valid, kb, acc, false);
    return;
  }
  store_tile<0>(scratch, TILE, valid, kb, acc, false);  // X0
  __threadfence();
  __syncthreads();
  acc_zero(acc);
  gemm_nt_mainloop<2>(stages, scratch, TILE, valid, Lkk, ldl, kb, kb, acc);  // X0 L_kk^T
  store_tile<1>(T, ldt, valid, kb, acc, false);                            // T <- R = T - X0 L_kk^T
  __threadfence();
  __syncthreads();
  acc_zero(acc);
  gemm_nt_mainloop<2>(stages, T, ldt, valid, Linv, TILE, kb, kb, acc);  // R Linv^T
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
  int K; int lower; int nt_rows1; int add;  // add != 0: C += A B^T instead of C -= A B^T
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
  gemm_nt_mainloop<0>(stages, Ap, lda, rows, Bp, p.ldb, cols, p.K, acc);
  if (p.add) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[i][j][0] = -acc[i][j][0];
        acc[i][j][1] = -acc[i][j][1];
      }
  }
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
template <int MODE>
__device__ __forceinline__ void gemm_nt_mainloop_dep(GemmStage* stages, const double* __restrict__ Ap, int64_t lda,
                                                     int validA, const double* __restrict__ Bp, int64_t ldb,
                                                     int validB, int K, Acc& acc, const int* readyA,
                                                     const int* readyB) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const int nchunks = K / BK;
  const unsigned mask = block_mask<MODE>(wm, wn, validA);
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
    mma_chunk<MODE>(stages[c % STAGES], acc, wm, wn, gid, tig, c * BK, mask);
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
  long long* prof;                // debug: globaltimer stamps of the tasks around column nt/2 (or null)
};

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

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
    // debug stamps: tasks (jp, jp), (jp+1, jp), (jp+1, jp+1) with jp = nt/2
    long long* pf = nullptr;
    if (p.prof && b == 0 && tid == 0) {
      const int jp = p.nt / 2;
      if (j == jp && i == jp) pf = p.prof;
      else if (j == jp && i == jp + 1) pf = p.prof + 8;
      else if (j == jp + 1 && i == jp + 1) pf = p.prof + 16;
    }
    if (pf) pf[0] = globaltimer_ns();

    if (j > 0) {
      Acc acc;
      acc_zero(acc);
      // (the masked lower-triangular MODE 1 path is slower than the full tile: its branches break
      //  the DMMA/LDS software pipeline; measured 2x per chunk)
      gemm_nt_mainloop_dep<0>(stages, rowi, ldi, valid, rowj, p.lda, kb, j * TILE, acc, ready_i, ready_j);
      if (pf) pf[1] = globaltimer_ns();
      store_tile<1>(T, ldi, valid, kb, acc, i == j);
      __threadfence();
      __syncthreads();
      if (pf) pf[2] = globaltimer_ns();
    }
    if (i == j) {
      diag_factor_tile(smem_raw, T, p.lda, kb, (int64_t)j * TILE, wsb + (int64_t)j * TILE * TILE,
                       wsb + (int64_t)p.nt * TILE * TILE + j, p.info + b);
    } else {
      wait_ready(ready_j + j);
      if (pf) pf[3] = globaltimer_ns();
      const bool refine = __ldcg(wsb + (int64_t)p.nt * TILE * TILE + j) != 0.0;
      tile_solve(stages, T, ldi, valid, kb, wsb + (int64_t)j * TILE * TILE, rowj + (int64_t)j * TILE, p.lda, refine,
                 scratch);
    }
    if (pf) pf[4] = globaltimer_ns();
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(ready_b + (int64_t)i * p.nt + j, 1);
    if (pf) pf[5] = globaltimer_ns();
  }
}

static long long* g_df_prof = nullptr;  // debug hook (gpar_debug_set_dataflow_prof)

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
    p.prof = g_df_prof;
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
                                                                       info, nullptr);
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
      p.K = kb; p.lower = 1; p.nt_rows1 = ntr; p.add = 0;
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

static int syrk_impl(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw, int64_t k,
                     int64_t strideW, int64_t batch, int add, void* stream_);

extern "C" int gpar_syrk_sub(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw,
                             int64_t k, int64_t strideW, int64_t batch, void* stream_) {
  return syrk_impl(C, ldc, n, strideC, W, ldw, k, strideW, batch, 0, stream_);
}

extern "C" int gpar_syrk_add(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw,
                             int64_t k, int64_t strideW, int64_t batch, void* stream_) {
  return syrk_impl(C, ldc, n, strideC, W, ldw, k, strideW, batch, 1, stream_);
}

static int syrk_impl(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw, int64_t k,
                     int64_t strideW, int64_t batch, int add, void* stream_) {
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
  p.K = (int)k; p.lower = 1; p.nt_rows1 = (int)nt; p.add = add;
  p.C2 = nullptr; p.ldc2 = 0; p.c_rows2 = 0; p.strideC2 = 0; p.Aop2 = nullptr; p.lda2 = 0; p.strideA2 = 0;
  gemm_sub_kernel<<<dim3(nt, nt, (unsigned)batch), GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
  return check_launch("gpar_syrk_sub");
}

// Debug: phase timestamps (clock64) of the diagonal-tile factor on the leading 128 block of A.
extern "C" int gpar_debug_diag_profile(double* A, int64_t lda, int64_t n, double* ws, int32_t* info, long long* prof,
                                       void* stream_) {
  set_smem_attrs();
  const int nt = (int)((n + TILE - 1) / TILE);
  potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, (cudaStream_t)stream_>>>(A, lda, 0, 0, nt, n, ws, ws_stride(nt, 0), info,
                                                                        prof);
  return check_launch("gpar_debug_diag_profile");
}

extern "C" int gpar_debug_set_dataflow_prof(long long* prof) {
  gpar::g_df_prof = prof;
  return 0;
}
