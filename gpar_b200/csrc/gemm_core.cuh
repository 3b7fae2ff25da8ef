// fp64 tensor-core (DMMA m8n8k4) GEMM core shared by potrf / trsm / syrk:
//   acc(128x128) += sum_k Arows[r][k] * Brows[c][k]       ("NT": both operands K-contiguous)
// Operand tiles stream global -> shared through a 4-stage cp.async (LDGSTS, L2-only) ring of
// 128 x 16 chunks with a padded row stride of 20 doubles, which makes every 64-bit fragment
// load bank-conflict free.  8 warps in a 4 (m) x 2 (n) grid, each owning 4 x 8 DMMA tiles (four
// interleaved 8-row groups x 64 columns), 64 accumulator doubles per thread; 12 shared loads
// feed 32 DMMAs per k-step of 4.
//
// tcgen05 has no f64 kind and TMEM holds no fp64 accumulators, so on sm_100a the fp64 dense
// contractions run on the warp-level DMMA path (SASS: DMMA.8x8x4).
#pragma once
#include "common.cuh"

namespace gpar {

constexpr int BK = 16;     // k-chunk
constexpr int LDSM = 20;   // padded shared row stride (doubles): 160 B keeps 16 B alignment
#ifndef GPAR_STAGES
#define GPAR_STAGES 4
#endif
constexpr int STAGES = GPAR_STAGES;
constexpr int GEMM_THREADS = 256;

struct __align__(16) GemmStage {
  double a[TILE * LDSM];
  double b[TILE * LDSM];
};
constexpr size_t GEMM_SMEM_BYTES = STAGES * sizeof(GemmStage);  // 163840

typedef double Acc[4][8][2];

__device__ __forceinline__ void acc_zero(Acc& acc) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// Issue the cp.async copies of one k-chunk [k0, k0+16) of both operands.  Rows >= valid and
// columns >= K are zero-filled (src_bytes < 16), so callers never need padded matrices.
__device__ __forceinline__ void load_chunk(GemmStage& st, const double* __restrict__ Ap, int64_t lda, int validA,
                                           const double* __restrict__ Bp, int64_t ldb, int validB, int k0, int K) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * GEMM_THREADS;  // 0..1023
    const int r = idx >> 3, seg = idx & 7;
    const int k = k0 + seg * 2;
    int kb = (K - k) * 8;
    kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
    const int ba = (r < validA) ? kb : 0;
    const int bb = (r < validB) ? kb : 0;
    const double* sa = ba ? (Ap + (int64_t)r * lda + k) : Ap;
    const double* sb = bb ? (Bp + (int64_t)r * ldb + k) : Bp;
    cp_async16(&st.a[r * LDSM + seg * 2], sa, ba);
    cp_async16(&st.b[r * LDSM + seg * 2], sb, bb);
  }
}

// Warp (wm, wn) owns the 8-row groups {(4 i + wm) * 8} (interleaved over the four m-warps, so that
// triangular tiles load all four SM sub-partitions evenly) and the columns wn*64 + j*8.
// MODE 0: full tile.  MODE 1: lower-triangular OUTPUT (SYRK on a diagonal tile): 8x8 sub-tiles
// strictly above the diagonal are skipped.  MODE 2: lower-triangular B OPERAND (B[c][k] = 0 for
// k > c, i.e. X Linv^T / X L^T): k-steps that only meet zeros are skipped.  Row groups at or
// beyond `valid_rows` are skipped in every mode (ragged tiles, the single y row).
__device__ __forceinline__ int acc_row(int wm, int i) { return (4 * i + wm) * 8; }

// Block mask of a warp for one task: bit (2 i + h) enables the 8 x 32 block made of row group i
// and column half h (columns wn*64 + 32 h .. + 31).  MODE 1 (lower-triangular OUTPUT, SYRK on a
// diagonal tile) drops blocks strictly above the diagonal; every mode drops row groups at or
// beyond `valid_rows` (ragged tiles, the single y row).  Warp-uniform, computed once per task.
template <int MODE>
__device__ __forceinline__ unsigned block_mask(int wm, int wn, int valid_rows) {
  unsigned m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      bool on = acc_row(wm, i) < valid_rows;
      if (MODE == 1) on = on && (wn * 64 + 32 * h <= acc_row(wm, i) + 7);
      if (on) m |= 1u << (2 * i + h);
    }
  return m;
}

// MODE 2 (lower-triangular B OPERAND, B[c][k] = 0 for k > c: X Linv^T, X L^T): a k-chunk starting
// at k0 only meets zeros in column half h when k0 > last column of the half -- one uniform test
// per chunk and half.  Blocks are straight-line groups of 4 DMMAs, so skipping costs one uniform
// branch per block and k-step instead of one per DMMA.
template <int MODE>
__device__ __forceinline__ void mma_chunk(const GemmStage& st, Acc& acc, int wm, int wn, int gid, int tig, int k0,
                                          unsigned mask) {
  if (MODE == 2) {
    if (k0 > wn * 64 + 31) mask &= 0xAAu;   // drop h = 0 blocks (bits 0, 2, 4, 6)
    if (k0 > wn * 64 + 63) mask = 0;
  }
  if (mask == 0) return;
#pragma unroll
  for (int kk = 0; kk < BK; kk += 4) {
    double a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = st.a[(acc_row(wm, i) + gid) * LDSM + kk + tig];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = st.b[(wn * 64 + j * 8 + gid) * LDSM + kk + tig];
    if (mask == 0xFFu) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    } else if (mask == 0xAAu) {  // right column half only (the common MODE 2 case): straight line
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 4; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!(mask & (1u << (2 * i + h)))) continue;
#pragma unroll
          for (int j = 4 * h; j < 4 * h + 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
  }
}

// Full pipelined mainloop.  All 256 threads must call it.  On return every cp.async has
// landed and all warps have passed a barrier (the stage ring may be reused immediately).
template <int MODE = 0>
__device__ __forceinline__ void gemm_nt_mainloop(GemmStage* stages, const double* __restrict__ Ap, int64_t lda,
                                                 int validA, const double* __restrict__ Bp, int64_t ldb, int validB,
                                                 int K, Acc& acc) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const int nchunks = (K + BK - 1) / BK;
  const unsigned mask = block_mask<MODE>(wm, wn, validA);
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_chunk(stages[s], Ap, lda, validA, Bp, ldb, validB, s * BK, K);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    mma_chunk<MODE>(stages[c % STAGES], acc, wm, wn, gid, tig, c * BK, mask);
    // copies after the DMMAs: issued first they queue ahead of the fragment loads (see potrf.cu)
    const int nc = c + STAGES - 1;
    if (nc < nchunks) load_chunk(stages[nc % STAGES], Ap, lda, validA, Bp, ldb, validB, nc * BK, K);
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncthreads();
}

// Epilogue helpers.  Element (i, j, e) of Acc is C[acc_row(wm, i) + gid][wn*64 + j*8 + 2*tig + e].
// mode 0: C = acc;  mode 1: C -= acc.  `lower_diag`: only write col <= row (tile on the diagonal).
template <int MODE>
__device__ __forceinline__ void store_tile(double* __restrict__ C, int64_t ldc, int rows, int cols, const Acc& acc,
                                           bool lower_diag) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const bool vec = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  // MODE 1 is a read-modify-write: per row group issue all 8 loads first (independent; interleaving
  // loads and stores would serialise 64 L2 round trips), then subtract and store.
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = acc_row(wm, i) + gid;
    double* crow = C + (int64_t)r * ldc;
    double2 old[8];
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = wn * 64 + j * 8 + 2 * tig;
        const bool ok0 = (r < rows) && (c < cols) && (!lower_diag || c <= r);
        const bool ok1 = (r < rows) && (c + 1 < cols) && (!lower_diag || c + 1 <= r);
        old[j] = make_double2(0.0, 0.0);
        if (ok0 && ok1 && vec) {
          old[j] = __ldcg(reinterpret_cast<const double2*>(crow + c));
        } else {
          if (ok0) old[j].x = __ldcg(crow + c);
          if (ok1) old[j].y = __ldcg(crow + c + 1);
        }
      }
    }
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = wn * 64 + j * 8 + 2 * tig;
      const bool ok0 = (c < cols) && (!lower_diag || c <= r);
      const bool ok1 = (c + 1 < cols) && (!lower_diag || c + 1 <= r);
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      if (MODE == 1) {
        v0 = old[j].x - v0;
        v1 = old[j].y - v1;
      }
      if (ok0 && ok1 && vec) {
        *reinterpret_cast<double2*>(crow + c) = make_double2(v0, v1);
      } else {
        if (ok0) crow[c] = v0;
        if (ok1) crow[c + 1] = v1;
      }
    }
  }
}

// acc += C2 (a dense tile with leading dimension ldc2, written earlier by store_tile<0> of the
// same thread layout: every thread re-reads exactly the elements it stored).
__device__ __forceinline__ void acc_add_tile(Acc& acc, const double* __restrict__ C2, int64_t ldc2, int rows,
                                             int cols) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = acc_row(wm, i) + gid;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = wn * 64 + j * 8 + 2 * tig;
      if (c < cols) acc[i][j][0] += C2[(int64_t)r * ldc2 + c];
      if (c + 1 < cols) acc[i][j][1] += C2[(int64_t)r * ldc2 + c + 1];
    }
  }
}

}  // namespace gpar
