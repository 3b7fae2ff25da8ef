// fp64 tensor-core (DMMA m8n8k4) GEMM core shared by potrf / trsm / syrk:
//   acc(128x128) += sum_k Arows[r][k] * Brows[c][k]       ("NT": both operands K-contiguous)
// Operand tiles stream global -> shared through a 5-stage cp.async (LDGSTS, L2-only) ring of
// 128 x 16 chunks with a padded row stride of 20 doubles, which makes every 64-bit fragment
// load bank-conflict free; stage hand-over runs on mbarriers (no CTA-wide barrier per chunk).  8 warps in a 4 (m) x 2 (n) grid, each owning 4 x 8 DMMA tiles (four
// interleaved 8-row groups x 64 columns), 64 accumulator doubles per thread; 12 shared loads
// feed 32 DMMAs per k-step of 4.
//
// tcgen05 has no f64 kind and TMEM holds no fp64 accumulators, so on sm_100a the fp64 dense
// contractions run on the warp-level DMMA path (SASS: DMMA.8x8x4).
#pragma once
#include "common.cuh"

namespace gpar {

#ifndef GPAR_BK
#define GPAR_BK 16
#endif
constexpr int BK = GPAR_BK;      // k-chunk
constexpr int LDSM = BK + 4;     // padded shared row stride (doubles), = 4 (mod 16): conflict-free fragment loads,
                                 // rows stay 16-byte aligned
#ifndef GPAR_STAGES
#define GPAR_STAGES 5
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
  constexpr int SEGS = BK / 2;  // 16-byte segments per row
#pragma unroll
  for (int it = 0; it < TILE * SEGS / GEMM_THREADS; ++it) {
    const int idx = threadIdx.x + it * GEMM_THREADS;
    const int r = idx / SEGS, seg = idx % SEGS;
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
// MODE 0: full tile (MODE 4: the same with straight-line bodies for 32 / 64 / 96 valid rows).  MODE 1: lower-triangular OUTPUT (SYRK on a diagonal tile): 8x8 sub-tiles
// strictly above the diagonal are skipped.  MODE 2: lower-triangular B OPERAND (B[c][k] = 0 for
// k > c, i.e. X Linv^T / X L^T): k-chunks that only meet zeros are skipped, and the accumulator
// columns are PERMUTED (acc_col<true>): every consumer of a MODE 2 result passes PERM = true.  Row
// groups at or beyond `valid_rows` are skipped in every mode (ragged tiles, the single y row).
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
      // MODE 3 (diagonal-tile SYRK at 64 x 64 granularity): the quadrant rows 0..63 x columns 64..127 lies above
      // the diagonal -- the right n-warps skip their first two row groups and run a straight-line half body
      // (unlike the 8 x 32 masks of MODE 1, whose per-block branches break the DMMA / LDS software pipeline)
      if (MODE == 3) on = on && !(wn == 1 && i < 2);
      if (on) m |= 1u << (2 * i + h);
    }
  return m;
}

// Column of accumulator slot j of n-warp wn.  Standard layout: wn owns the contiguous half
// wn*64 + j*8.  PERM (used by MODE 2 and everything that consumes its accumulators): the 8-column
// tiles are dealt alternately, (2 j + wn) * 8 -- with a lower-triangular B operand the K-extent of a
// column tile grows with its index, so the contiguous split leaves the left warp idle for the second
// half of the K loop (36 vs 100 block-units) while the alternating split gives both warps of a
// sub-partition 36 chunk-tiles each.
template <bool PERM>
__device__ __forceinline__ int acc_col(int wn, int j) { return PERM ? (2 * j + wn) * 8 : wn * 64 + j * 8; }

// MODE 2 (lower-triangular B OPERAND, B[c][k] = 0 for k > c: X Linv^T, X L^T): the k-chunk starting at
// k0 only meets zeros in column tile ct = 2 j + wn when k0 > 8 ct + 7, i.e. the active tiles are the
// suffix j >= J0 of the warp's slots: one uniform switch per chunk selects a straight-line body.
template <int J0>
__device__ __forceinline__ void mma_tri_steps(const GemmStage& st, Acc& acc, int wm, int wn, int gid, int tig,
                                              unsigned rowmask) {
#pragma unroll
  for (int kk = 0; kk < BK; kk += 4) {
    double a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = st.a[(acc_row(wm, i) + gid) * LDSM + kk + tig];
#pragma unroll
    for (int j = J0; j < 8; ++j) b[j] = st.b[(acc_col<true>(wn, j) + gid) * LDSM + kk + tig];
    if (rowmask == 0xFu) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = J0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    } else {  // ragged tiles, 32-row operands
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!(rowmask & (1u << i))) continue;
#pragma unroll
        for (int j = J0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
  }
}

// MODE 4 bodies: the first NI row groups of every warp, all 8 column tiles.
template <int NI>
__device__ __forceinline__ void mma_rows_steps(const GemmStage& st, Acc& acc, int wm, int wn, int gid, int tig) {
#pragma unroll
  for (int kk = 0; kk < BK; kk += 4) {
    double a[NI], b[8];
#pragma unroll
    for (int i = 0; i < NI; ++i) a[i] = st.a[(acc_row(wm, i) + gid) * LDSM + kk + tig];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = st.b[(wn * 64 + j * 8 + gid) * LDSM + kk + tig];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

template <int MODE>
__device__ __forceinline__ void mma_chunk(const GemmStage& st, Acc& acc, int wm, int wn, int gid, int tig, int k0,
                                          unsigned mask) {
  if (MODE == 2) {
    unsigned rowmask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (mask & (3u << (2 * i))) rowmask |= 1u << i;
    const int t = k0 - 8 * wn - 7;
    const int j0 = t <= 0 ? 0 : (t + 15) / 16;
    if (rowmask == 0 || j0 > 7) return;
    switch (j0) {
      case 0: mma_tri_steps<0>(st, acc, wm, wn, gid, tig, rowmask); break;
      case 1: mma_tri_steps<1>(st, acc, wm, wn, gid, tig, rowmask); break;
      case 2: mma_tri_steps<2>(st, acc, wm, wn, gid, tig, rowmask); break;
      case 3: mma_tri_steps<3>(st, acc, wm, wn, gid, tig, rowmask); break;
      case 4: mma_tri_steps<4>(st, acc, wm, wn, gid, tig, rowmask); break;
      case 5: mma_tri_steps<5>(st, acc, wm, wn, gid, tig, rowmask); break;
      case 6: mma_tri_steps<6>(st, acc, wm, wn, gid, tig, rowmask); break;
      default: mma_tri_steps<7>(st, acc, wm, wn, gid, tig, rowmask); break;
    }
    return;
  }
  if (mask == 0) return;
  if (MODE == 4) {
    // row blocks of 32 / 64 / 96 / 128 rows (trsm_rows_kernel's tail blocks): the interleaved row groups
    // keep all eight warps busy with 1 .. 4 groups each; one uniform switch selects a straight-line body
    switch (mask) {
      case 0x03u: mma_rows_steps<1>(st, acc, wm, wn, gid, tig); return;
      case 0x0Fu: mma_rows_steps<2>(st, acc, wm, wn, gid, tig); return;
      case 0x3Fu: mma_rows_steps<3>(st, acc, wm, wn, gid, tig); return;
      case 0xFFu: mma_rows_steps<4>(st, acc, wm, wn, gid, tig); return;
      default: break;  // ragged last block: the per-block branches below
    }
  }
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
    } else if (mask == 0xF0u) {  // row groups 2, 3 only (MODE 3, right n-warps)
#pragma unroll
      for (int i = 2; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    } else if (mask == 0x03u) {  // first row group only (32-row operands: the L^-T sweep of gpar_potri)
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884(acc[0][j][0], acc[0][j][1], a[0], b[j]);
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

// ---- mbarrier pipeline: no CTA-wide barrier per chunk ------------------------------------------
// full[s]  (count 256): armed by cp.async.mbarrier.arrive.noinc of every thread after its copies of
//                       the chunk; a warp starts the chunk's DMMAs as soon as the phase completes.
// empty[s] (count 8):   one arrival per warp when it is done reading the stage; a thread refills
//                       stage (c-1) % S only after that phase, i.e. warps may drift by up to a chunk
//                       instead of marching in lock step.
// The barriers live for the whole (persistent) kernel; `n` counts the chunks consumed so far so
// that stage and phase parity carry over from one mainloop call to the next.
struct PipeShared {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint32_t n;
};
__device__ __forceinline__ PipeShared& pipe_shared() {
  __shared__ __align__(8) PipeShared ps;
  return ps;
}
__device__ __forceinline__ void pipe_init() {
  PipeShared& ps = pipe_shared();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&ps.full[s], GEMM_THREADS);
      mbar_init(&ps.empty[s], GEMM_THREADS / 32);
    }
    ps.n = 0;
    fence_mbar_init();
  }
  __syncthreads();
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(b) : "memory");
}

// WAIT(kt) is called by every thread before it issues the first chunk of k-tile kt.
template <int MODE, typename WaitFn>
__device__ __forceinline__ void gemm_nt_pipe(GemmStage* stages, const double* __restrict__ Ap, int64_t lda, int validA,
                                             const double* __restrict__ Bp, int64_t ldb, int validB, int K, Acc& acc,
                                             WaitFn wait_tile) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const int nchunks = (K + BK - 1) / BK;
  const unsigned mask = block_mask<MODE>(wm, wn, validA);
  constexpr int CPT = TILE / BK;
#ifdef GPAR_SANITIZE_SYNC
  // Sanitizer build (scripts/gpu_sanitize.sh): the same ring driven by cp.async commit groups and one CTA
  // barrier per chunk -- the classic multistage pipeline that compute-sanitizer's racecheck models.  The
  // product build replaces the per-chunk barrier by the mbarrier hand-over below (6 % faster; racecheck
  // does not model cp.async.mbarrier.arrive / mbarrier.try_wait as synchronisation and reports the ring).
  {
#pragma unroll
    for (int q = 0; q < STAGES - 1; ++q) {
      if (q < nchunks) {
        if (q % CPT == 0) wait_tile(q / CPT);
        load_chunk(stages[q % STAGES], Ap, lda, validA, Bp, ldb, validB, q * BK, K);
      }
      cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int q = c + STAGES - 1;
      if (q < nchunks) {
        if (q % CPT == 0) wait_tile(q / CPT);
        load_chunk(stages[q % STAGES], Ap, lda, validA, Bp, ldb, validB, q * BK, K);
      }
      cp_async_commit();
      mma_chunk<MODE>(stages[c % STAGES], acc, wm, wn, gid, tig, c * BK, mask);
    }
    cp_async_wait<0>();
    __syncthreads();
    return;
  }
#endif
  PipeShared& ps = pipe_shared();
  const uint32_t base = ps.n;
  auto issue = [&](int q) {
    if (q % CPT == 0) wait_tile(q / CPT);
    const uint32_t g = base + q;
    load_chunk(stages[g % STAGES], Ap, lda, validA, Bp, ldb, validB, q * BK, K);
    cp_async_mbar_arrive(&ps.full[g % STAGES]);
  };
#pragma unroll
  for (int q = 0; q < STAGES - 1; ++q)
    if (q < nchunks) issue(q);
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t g = base + c;
    mbar_wait(&ps.full[g % STAGES], (g / STAGES) & 1);
    mma_chunk<MODE>(stages[g % STAGES], acc, wm, wn, gid, tig, c * BK, mask);
    __syncwarp();
    if (lane == 0) mbar_arrive(&ps.empty[g % STAGES]);
    const int q = c + STAGES - 1;
    if (q < nchunks) {
      if (c >= 1) mbar_wait(&ps.empty[(g - 1) % STAGES], ((g - 1) / STAGES) & 1);
      issue(q);
    }
  }
  __syncthreads();  // every warp is done with the ring (callers reuse it), all copies have landed
  if (threadIdx.x == 0) ps.n = base + nchunks;
  __syncthreads();
}

template <int MODE = 0>
__device__ __forceinline__ void gemm_nt_mainloop(GemmStage* stages, const double* __restrict__ Ap, int64_t lda,
                                                 int validA, const double* __restrict__ Bp, int64_t ldb, int validB,
                                                 int K, Acc& acc) {
  gemm_nt_pipe<MODE>(stages, Ap, lda, validA, Bp, ldb, validB, K, acc, [](int) {});
}
// Epilogue helpers.  Element (i, j, e) of Acc is C[acc_row(wm, i) + gid][acc_col<PERM>(wn, j) + 2*tig + e].
// mode 0: C = acc;  mode 1: C -= acc.  `lower_diag`: only write col <= row (tile on the diagonal).
template <int MODE, bool PERM = false>
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
        const int c = acc_col<PERM>(wn, j) + 2 * tig;
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
      const int c = acc_col<PERM>(wn, j) + 2 * tig;
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
template <bool PERM = false>
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
      const int c = acc_col<PERM>(wn, j) + 2 * tig;
      if (c < cols) acc[i][j][0] += C2[(int64_t)r * ldc2 + c];
      if (c + 1 < cols) acc[i][j][1] += C2[(int64_t)r * ldc2 + c + 1];
    }
  }
}

}  // namespace gpar
