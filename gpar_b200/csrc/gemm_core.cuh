// fp64 tensor-core (DMMA m8n8k4) GEMM core shared by potrf / trsm / syrk:
//   acc(128x128) += sum_k Arows[r][k] * Brows[c][k]       ("NT": both operands K-contiguous)
// Operand tiles stream global -> shared through a 4-stage cp.async (LDGSTS, L2-only) ring of
// 128 x 16 chunks with a padded row stride of 20 doubles, which makes every 64-bit fragment
// load bank-conflict free.  8 warps in a 4 (m) x 2 (n) grid, warp tile 32 x 64 = 4 x 8 DMMA
// tiles, 64 accumulator doubles per thread; 12 shared loads feed 32 DMMAs per k-step of 4.
//
// tcgen05 has no f64 kind and TMEM holds no fp64 accumulators, so on sm_100a the fp64 dense
// contractions run on the warp-level DMMA path (SASS: DMMA.8x8x4).
#pragma once
#include "common.cuh"

namespace gpar {

constexpr int BK = 16;     // k-chunk
constexpr int LDSM = 20;   // padded shared row stride (doubles): 160 B keeps 16 B alignment
constexpr int STAGES = 4;
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

__device__ __forceinline__ void mma_chunk(const GemmStage& st, Acc& acc, int wm, int wn, int gid, int tig) {
#pragma unroll
  for (int kk = 0; kk < BK; kk += 4) {
    double a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = st.a[(wm * 32 + i * 8 + gid) * LDSM + kk + tig];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = st.b[(wn * 64 + j * 8 + gid) * LDSM + kk + tig];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// Full pipelined mainloop.  All 256 threads must call it.  On return every cp.async has
// landed and all warps have passed a barrier (the stage ring may be reused immediately).
__device__ __forceinline__ void gemm_nt_mainloop(GemmStage* stages, const double* __restrict__ Ap, int64_t lda,
                                                 int validA, const double* __restrict__ Bp, int64_t ldb, int validB,
                                                 int K, Acc& acc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const int nchunks = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_chunk(stages[s], Ap, lda, validA, Bp, ldb, validB, s * BK, K);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nc = c + STAGES - 1;
    if (nc < nchunks) load_chunk(stages[nc % STAGES], Ap, lda, validA, Bp, ldb, validB, nc * BK, K);
    cp_async_commit();
    mma_chunk(stages[c % STAGES], acc, wm, wn, gid, tig);
  }
  cp_async_wait<0>();
  __syncthreads();
}

// Epilogue helpers.  Element (i, j, e) of Acc is C[wm*32 + i*8 + gid][wn*64 + j*8 + 2*tig + e].
// mode 0: C = acc;  mode 1: C -= acc.  `lower_diag`: only write col <= row (tile on the diagonal).
template <int MODE>
__device__ __forceinline__ void store_tile(double* __restrict__ C, int64_t ldc, int rows, int cols, const Acc& acc,
                                           bool lower_diag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = wm * 32 + i * 8 + gid;
    if (r >= rows) continue;
    double* crow = C + (int64_t)r * ldc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = wn * 64 + j * 8 + 2 * tig;
      const bool ok0 = (c < cols) && (!lower_diag || c <= r);
      const bool ok1 = (c + 1 < cols) && (!lower_diag || c + 1 <= r);
      if (ok0 && ok1) {
        double2* p = reinterpret_cast<double2*>(crow + c);
        if (MODE == 0) {
          *p = make_double2(acc[i][j][0], acc[i][j][1]);
        } else {
          double2 v = *p;
          v.x -= acc[i][j][0];
          v.y -= acc[i][j][1];
          *p = v;
        }
      } else {
        if (ok0) crow[c] = (MODE == 0) ? acc[i][j][0] : crow[c] - acc[i][j][0];
        if (ok1) crow[c + 1] = (MODE == 0) ? acc[i][j][1] : crow[c + 1] - acc[i][j][1];
      }
    }
  }
}

// C = C2 + acc (C2 is a dense tile with leading dimension ldc2).
__device__ __forceinline__ void store_tile_add(double* __restrict__ C, int64_t ldc, const double* __restrict__ C2,
                                               int64_t ldc2, int rows, int cols, const Acc& acc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = wm * 32 + i * 8 + gid;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = wn * 64 + j * 8 + 2 * tig;
      if (c < cols) C[(int64_t)r * ldc + c] = C2[(int64_t)r * ldc2 + c] + acc[i][j][0];
      if (c + 1 < cols) C[(int64_t)r * ldc + c + 1] = C2[(int64_t)r * ldc2 + c + 1] + acc[i][j][1];
    }
  }
}

}  // namespace gpar
