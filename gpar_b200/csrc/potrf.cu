// K2 / K4 / K5 / K10: blocked Cholesky with appended row blocks (one GPU or sharded over the GPUs of an
// NVLink domain), triangular solve of row blocks, symmetric rank-k update, and the inverse from the
// factor -- all on the DMMA GEMM core (gemm_core.cuh).  Tile = 128.
//
// Map of this file:
//   diag_factor_core        128 x 128 diagonal tile: L_kk, L_kk^-1, conditioning flag (warp-shuffle 32-panels,
//                           substitution row solves, rank-32 DMMA updates), + peer pushes when sharded
//   tile_solve              X = T L_kk^-T through the inverse tile (+ one refinement step when flagged)
//   potrf_dataflow_kernel   the product path: persistent left-looking tile dataflow (ticketed D0 / HEAD / PLAIN /
//                           PRE tasks, ready flags, split-K tail); <MULTI> adds row-block ownership and in-kernel
//                           NVLink pushes (gpar_potrf_multi)
//   gemm_sub_kernel         C -/+= A B^T (SYRK / GEMM, batched; triangular-operand mode for gpar_potri)
//   trsm_rows_kernel        B <- B L^-T for extra row blocks after the fact
//   trtri_rows_kernel       L^-T by the same sweep over the identity (gpar_potri)
//   potrf_diag / trsm_tile  v1 right-looking sweep, three launches per 128 columns (GPAR_POTRF_V1=1; kept as
//                           the measured baseline of DESIGN.md section 3)
// Appended rows B (nb x n) ride along as extra row tiles, so B <- B L^-T falls out of the same
// sweep: with B = y^T this is the forward solve of the log-marginal; with the joint matrix
// [[K_aa, .], [K_*a, K_**]] the sweep yields L, V^T = K_*a L^-T and chol(K_** - V^T V) at once.
#include <cstdlib>
#include <cstring>

#include "gemm_core.cuh"

namespace gpar {

// --------------------------------------------------------------------------------------
// Diagonal tile: factor + invert, entirely in shared memory.
// --------------------------------------------------------------------------------------
constexpr int DLD = 132;  // row stride of the tile in shared memory: = 4 (mod 16) doubles makes the DMMA
                          // fragment loads bank-conflict free; rows stay 16-byte aligned for cp.async
constexpr int PANEL = 32; // panel width of the in-tile blocked factorisation
constexpr double REFINE_KAPPA = 1.0e3;  // kappa_inf(L_kk) above which the tile solves are refined
// shared layout: Ls[TILE][DLD] | rdiag[TILE] | LT[PANEL][PANEL] | red[32] | rowL[TILE] | isumP[8][TILE] | ints
constexpr size_t DIAG_SMEM_BYTES = sizeof(double) * (TILE * DLD + TILE + PANEL * PANEL + 32 + TILE + 8 * TILE) + 16;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
#ifdef GPAR_SANITIZE_SYNC
  // sanitizer build: the producer warp blocks too (bar.sync) -- racecheck does not order bar.arrive / bar.sync pairs
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
#else
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// ---- multi-GPU (one process per GPU): every rank holds a full copy of the matrix and workspace at
// identical offsets of a peer-mapped allocation; a finished tile is pushed into every peer's copy
// over NVLink by plain stores at `local address + delta[r]`, followed by a system-scope fence and
// the peers' ready flags.  world == 1: nothing of this is touched.
struct Peers {
  int rank, world;
  int row_block;                    // tile rows are dealt to the ranks in blocks of this many (block-cyclic)
  long long delta[GPAR_MAX_PEERS];  // byte distance from this rank's allocation to rank r's (0 for r == rank)
};
template <typename T>
__device__ __forceinline__ T* peer_ptr(T* p, long long d) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(p) + d);
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// Publication targets of a finished tile.  The rank that owns the next tile row sits on the critical
// chain (its HEAD task consumes the tile next), so it is served -- and flagged -- first; the other
// peers follow.  PEERS_FIRST = this rank + the next one, PEERS_REST = everybody else.
enum { PEERS_ALL = 0, PEERS_FIRST = 1, PEERS_REST = 2 };
__device__ __forceinline__ bool peer_in_set(const Peers* pe, int r, int set) {
  const int next = (pe->rank + 1) % pe->world;
  if (set == PEERS_ALL) return true;
  const bool first = (r == pe->rank) || (r == next);
  return set == PEERS_FIRST ? first : !first;
}
// Called by one thread after the CTA's stores were fenced: raise a flag on the ranks of `set`.
__device__ __forceinline__ void publish_flag(int* flag, const Peers* pe, int set = PEERS_ALL) {
  if (pe == nullptr || pe->world == 1) {
    st_release(flag, 1);
    return;
  }
  for (int r = 0; r < pe->world; ++r)
    if (r != pe->rank && peer_in_set(pe, r, set)) st_release_sys(peer_ptr(flag, pe->delta[r]), 1);
  if (peer_in_set(pe, pe->rank, set)) st_release_sys(flag, 1);
}
__device__ __forceinline__ void fence_publish(const Peers* pe) {
  if (pe == nullptr || pe->world == 1) __threadfence(); else __threadfence_system();
}

// Load the lower triangle of a diagonal tile into Ls (zeros above the diagonal, identity padding
// beyond kb).  All 256 threads; ends with a barrier.
__device__ __forceinline__ void diag_load(unsigned char* smem_raw, const double* __restrict__ Atile, int64_t lda,
                                          int kb) {
  double* Ls = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x;
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
  __syncthreads();
}

// Explicit shared-space accesses for the latency-critical panel routines: with generic pointers the
// compiler re-derives the shared window base (S2UR SR_CgaCtaId) after every barrier / __syncwarp,
// which sits right on the pivot chain.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_v2f64(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(a), "d"(x), "d"(y) : "memory");
}
// 1 / sqrt(x) for a normal positive x: MUFU.RSQ64H seed (2^-22) + one third-order correction, the
// same arithmetic as the CUDA library routine without its special-case branch (pivots that are
// not positive normal numbers were replaced before this is called).
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y0) : "d"(x));
  const double e = fma(-(y0 * y0), x, 1.0);
  const double pp = fma(e, 0.375, 0.5);
  const double ye = y0 * e;
  return fma(pp, ye, y0);
}

// Warp 0 of a panel step: Cholesky of the 32 x 32 diagonal block at (c0, c0), one ROW per lane in
// registers, right-looking.  Per column: the pivot travels by one shuffle, every lane takes the
// rsqrt redundantly, scales its entry, and the rank-1 update reads L[l][j] back from the published
// column (broadcast shared loads; shuffling every operand is scoreboard-bound, 3x slower).
// The next pivot is updated first (own-lane operands only), so the dependent chain per column is
// shuffle -> rsqrt -> mul -> fma.  Column j is published transposed (LT[j][i] = L[i][j], conflict
// free) with rdiag[c0 + j] = 1 / L_jj; every 8 columns the warp arrives on named barrier 1 + j/8,
// on which the row-solving warps wait.  Returns the 1-based index of the first bad pivot (or 0).
__device__ __forceinline__ int factor32_warp(double* __restrict__ Ls, double* __restrict__ rdiag,
                                             double* __restrict__ LT, double* __restrict__ rowL, int c0, int lane) {
  double a[PANEL];
  const uint32_t row = smem_u32(Ls + (c0 + lane) * DLD + c0);
  const uint32_t lt = smem_u32(LT);
  const uint32_t rd = smem_u32(rdiag + c0);
#pragma unroll
  for (int q = 0; q < PANEL / 2; ++q) {
    const double2 v = lds_v2f64(row + 16 * q);
    a[2 * q] = v.x;
    a[2 * q + 1] = v.y;
  }
  int bad = 0;
  double piv = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
  for (int j = 0; j < PANEL; ++j) {
    if (!(piv > 1.0e-300) || piv > 1.0e300) {
      if (bad == 0) bad = j + 1;
      piv = 1.0;
    }
    const double rs = rsqrt_pos(piv);
    const double lij = ((lane == j) ? piv : a[j]) * rs;  // lane j: sqrt(piv); lanes > j: L[i][j]
    sts_f64(lt + 8 * (j * PANEL + lane), lij);
    sts_f64(row + 8 * j, lij);  // L[i][j] into the tile.  Lanes < j drop garbage into the strict upper part of
                                // the block, which the identity-row warp overwrites after the last barrier.
    if (lane == j) sts_f64(rd + 8 * j, rs);
    if (j + 1 < PANEL) {
      const double dn = fma(-lij, lij, a[j + 1]);  // next pivot (meaningful in lane j + 1)
      piv = __shfl_sync(0xffffffffu, dn, j + 1);
    }
    __syncwarp();  // column j is visible to the whole warp
    if (j + 1 < PANEL) {
      // rank-1 update: L[l][j] comes back from the published column as broadcast loads (the pivot
      // chain above does not wait for them)
      const uint32_t col = lt + 8 * (j * PANEL);
      int l = j + 1;
      if (l & 1) {
        a[l] = fma(-lij, lds_f64(col + 8 * l), a[l]);
        ++l;
      }
#pragma unroll
      for (; l < PANEL; l += 2) {
        const double2 t = lds_v2f64(col + 8 * l);
        a[l] = fma(-lij, t.x, a[l]);
        a[l + 1] = fma(-lij, t.y, a[l + 1]);
      }
    }
    a[j] = lij;
    if ((j & 7) == 7) {
      __syncwarp();
      named_bar_arrive(1 + (j >> 3), 160);
    }
  }
  // ||L||_inf bookkeeping: this row's entries inside the block (columns <= row)
  double sabs = 0.0;
#pragma unroll
  for (int j = 0; j < PANEL; ++j) sabs += (j <= lane) ? fabs(a[j]) : 0.0;
  rowL[c0 + lane] += sabs;
  return bad;
}

// Warps 1-4 of a panel step: row r of the tile (one row per lane) is solved against the 32 x 32
// block, x L32^T = t, by right-looking substitution as the columns are published (backward
// stable, no inverse involved).  Rows above the panel are inverse rows (their entries live in the
// strict upper triangle), rows inside the panel start as identity rows (the inverse of the block
// itself falls out), rows below are rows of L.
__device__ __forceinline__ void panel_solve_warp(double* __restrict__ Ls, const double* __restrict__ rdiag,
                                                 const double* __restrict__ LT, double* __restrict__ rowL, int c0,
                                                 int r) {
  double x[PANEL];
  const uint32_t row = smem_u32(Ls + r * DLD + c0);
  const uint32_t lt = smem_u32(LT);
  const uint32_t rd = smem_u32(rdiag + c0);
  const bool inblk = (r >= c0) && (r < c0 + PANEL);
  if (!inblk) {
#pragma unroll
    for (int q = 0; q < PANEL / 2; ++q) {
      const double2 v = lds_v2f64(row + 16 * q);
      x[2 * q] = v.x;
      x[2 * q + 1] = v.y;
    }
  } else {
    // rows of the panel start as identity rows; their slots in the tile are being written by the
    // factoring warp right now and must not be read here
#pragma unroll
    for (int j = 0; j < PANEL; ++j) x[j] = (j == r - c0) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int j = 0; j < PANEL; ++j) {
    if ((j & 7) == 0) named_bar_sync(1 + (j >> 3), 160);
    const double xj = x[j] * lds_f64(rd + 8 * j);
    x[j] = xj;
    const uint32_t col = lt + 8 * (j * PANEL);  // col[l] = L[l][j]
    int l = j + 1;
    if (l & 1) {
      if (l < PANEL) x[l] = fma(-xj, lds_f64(col + 8 * l), x[l]);
      ++l;
    }
#pragma unroll
    for (; l < PANEL; l += 2) {
      const double2 v = lds_v2f64(col + 8 * l);
      x[l] = fma(-xj, v.x, x[l]);
      x[l + 1] = fma(-xj, v.y, x[l + 1]);
    }
  }
  if (!inblk) {
#pragma unroll
    for (int q = 0; q < PANEL / 2; ++q) sts_v2f64(row + 16 * q, x[2 * q], x[2 * q + 1]);
    if (r >= c0 + PANEL) {  // a row of L: its 32 new entries enter ||L||_inf (same thread every panel: no race)
      double sabs = 0.0;
#pragma unroll
      for (int j = 0; j < PANEL; ++j) sabs += fabs(x[j]);
      rowL[r] += sabs;
    }
  } else {
    // the diagonal entry is rdiag (already there) and the lower part holds L32: strict upper only
#pragma unroll
    for (int j = 0; j < PANEL; ++j)
      if (j > r - c0) sts_f64(row + 8 * j, x[j]);
  }
}

// Rank-32 trailing update with the results of panel p: C -= X_panel L_panel^T on the inverse rows
// (rows < c0 + 32, strict-upper region) and on the lower part of the L rows, for the 32-column blocks
// cb_lo .. cb_hi.  Work unit = 16 rows x 32 columns (2 x 4 DMMA tiles, K = 32), dealt round-robin to
// the `nw` warps wfirst .. wfirst + nw - 1.  The caller splits the update in two: the block right
// behind the panel (needed by the next panel step) by all warps, the blocks beyond it by the three
// warps that idle during the next panel step.
__device__ __forceinline__ void rank32_update(double* __restrict__ Ls, const double* __restrict__ rdiag, int p,
                                              int cb_lo, int cb_hi, int wfirst, int nw, int warp, int gid,
                                              int tig) {
  const int c0 = PANEL * p;
  if (warp < wfirst || warp >= wfirst + nw) return;
  int u = 0;
  for (int rb = 0; rb < TILE / 16; ++rb) {
    const int R0 = 16 * rb;
    const bool lrows = R0 >= c0 + PANEL;
    const bool inblk = (R0 >= c0) && !lrows;
    for (int cb = cb_lo; cb <= cb_hi; ++cb) {
      if (lrows && PANEL * cb > R0 + 15) continue;
      if ((u++ % nw) != warp - wfirst) continue;
      const int C0 = PANEL * cb;
      double af[2][8], bf[4][8], cf[2][4][2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int r = R0 + 8 * t + gid;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const int kk = 4 * ks + tig;
          double v = Ls[r * DLD + c0 + kk];
          if (inblk) v = (kk > r - c0) ? v : ((kk == r - c0) ? rdiag[r] : 0.0);
          af[t][ks] = -v;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) bf[q][ks] = Ls[(C0 + 8 * q + gid) * DLD + c0 + 4 * ks + tig];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 v = *reinterpret_cast<const double2*>(Ls + (R0 + 8 * t + gid) * DLD + C0 + 8 * q + 2 * tig);
          cf[t][q][0] = v.x;
          cf[t][q][1] = v.y;
        }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) dmma884(cf[t][q][0], cf[t][q][1], af[t][ks], bf[q][ks]);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rt = 2 * rb + t, ct = 4 * cb + q;
          double* pc = Ls + (R0 + 8 * t + gid) * DLD + C0 + 8 * q + 2 * tig;
          if (!lrows || ct < rt) {
            *reinterpret_cast<double2*>(pc) = make_double2(cf[t][q][0], cf[t][q][1]);
          } else if (ct == rt) {  // diagonal sub-tile: its strict upper part belongs to the inverse rows
            if (2 * tig <= gid) pc[0] = cf[t][q][0];
            if (2 * tig + 1 <= gid) pc[1] = cf[t][q][1];
          }
        }
    }
  }
}

// Factor the diagonal tile held in Ls (see diag_load / assemble_diag) -- all 256 threads.  On
// return the tile L_kk (strictly upper part of each 32 x 32 diagonal block zeroed) is in Atile, its
// inverse in ws, *flag_out holds the refinement flag, *info_b the first bad pivot (if none was
// recorded before).  When ready_flag != nullptr it is released as soon as everything a consumer
// reads is in global memory: after Linv for well-conditioned tiles (only the refined solves read
// L_kk), after L_kk otherwise.
//
// One 128 x 132 shared array: lower triangle = L (in place); strict upper = the appended identity
// rows of the sweep, (I L^-T)[r][c] = Linv[c][r] (the diagonal of Linv is rdiag = 1 / L_rr), so the
// inverse costs no extra phase.  Four panel steps of width 32: warp 0 factors the diagonal block
// (shuffle pivots, registers), warps 1-4 solve the 128 rows against it as its columns appear, then
// all warps apply the rank-32 DMMA update to the columns beyond the panel.
//
// panel_flags (single GPU, optional): four flags of this tile.  Row panel p of the results -- rows [32 p, 32 p + 32)
// of L_kk and of L_kk^-1 -- is final as soon as panel step p is over; it is written out right then (by the warps
// that idle during the next panel step) and panel_flags[p] is released, so that the HEAD task of the next
// column can run its 32-column blocks behind the factorisation instead of after it (head_pipelined).
template <bool MULTI>
__device__ __noinline__ void diag_factor_core(unsigned char* smem_raw, double* __restrict__ Atile, int64_t lda,
                                                 int kb, int64_t j0, double* __restrict__ ws,
                                                 double* __restrict__ flag_out, int32_t* __restrict__ info_b,
                                                 int* ready_flag, const Peers* pe, long long* prof,
                                                 int* panel_flags = nullptr) {
#define GPAR_PROF(i) do { if (prof && threadIdx.x == 0) prof[i] = clock64(); } while (0)
  double* Ls = reinterpret_cast<double*>(smem_raw);
  double* rdiag = Ls + TILE * DLD;
  double* LT = rdiag + TILE;
  double* red = LT + PANEL * PANEL;  // 32 doubles of reduction scratch
  double* rowL = red + 32;           // running row sums of |L| (accumulated panel by panel)
  double* isumP = rowL + TILE;       // [8][TILE]: per-warp partial row sums of |Linv| (panel-wise write-out)
  int* s_int = reinterpret_cast<int*>(isumP + 8 * TILE);  // [0] first bad pivot, [1] refine
  const int tid = threadIdx.x, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int warp = canonical_warp();
  const bool multi = MULTI && (pe != nullptr) && (pe->world > 1);  // MULTI = false: all peer code compiles out
  if (tid == 0) s_int[0] = 0;
  if (tid < TILE) rowL[tid] = 0.0;
  // Panel-wise write-out of the LOCAL copy (also with several GPUs: the HEAD of the next column usually lives on
  // this rank -- tile rows are dealt in blocks of row_block -- and pipelines behind the local panel flags; the
  // peers get the finished tile and the ready flag at the end, as before).
  const bool panelwise = panel_flags != nullptr;
  const bool panel_end = panelwise && !multi;  // single GPU: the last panel is written the same way, no peer pass
  if (panelwise)
    for (int q = tid; q < 8 * TILE; q += 256) isumP[q] = 0.0;
  const bool vec_l = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(Atile) & 15) == 0);
  // Write row panel pp of Linv (to ws) and of L (to Atile) with the warps wfirst .. wfirst + nw - 1.
  // Linv[r][c] = Ls[c][r] (c < r), rdiag[r] at c = r, zeros up to the end of the diagonal 32-block; lanes run
  // along r (conflict-free), a lane assembles 4 consecutive c.  L rows: lanes along the columns.
  auto write_panel = [&](int pp, int wfirst, int nw) {
    if (warp < wfirst || warp >= wfirst + nw) return;
    const int wl = warp - wfirst;
    const int r = PANEL * pp + lane;
    double isum = 0.0;
    for (int cg = wl; cg < 8 * (pp + 1); cg += nw) {
      const int c = 4 * cg;
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double off = Ls[(c + q) * DLD + r];
        v[q] = (c + q < r) ? off : ((c + q == r) ? rdiag[r] : 0.0);
      }
      if (r < kb) {
        double* dst = ws + r * TILE + c;
        *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
        isum += (fabs(v[0]) + fabs(v[1])) + (fabs(v[2]) + fabs(v[3]));
      }
    }
    isumP[warp * TILE + r] += isum;  // slot (warp, r) belongs to this lane alone
    const int c = 4 * lane;
    for (int rr = PANEL * pp + wl; rr < PANEL * pp + PANEL && rr < kb; rr += nw) {
      if (c > (rr | 31) || c >= kb) continue;
      const double2 v0 = *reinterpret_cast<const double2*>(Ls + rr * DLD + c);
      const double2 v1 = *reinterpret_cast<const double2*>(Ls + rr * DLD + c + 2);
      double v[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c + q > rr) v[q] = 0.0;
      double* pd = Atile + (int64_t)rr * lda + c;
      if (vec_l && c + 3 < kb) {
        *reinterpret_cast<double2*>(pd) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(pd + 2) = make_double2(v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c + q < kb) pd[q] = v[q];
      }
    }
  };
  GPAR_PROF(1);
#pragma unroll 1
  for (int p = 0; p < TILE / PANEL; ++p) {
    const int c0 = PANEL * p;
    __syncthreads();  // the tile (or the previous rank-32 update) is complete
    if (warp == 0) {
      const int bad = factor32_warp(Ls, rdiag, LT, rowL, c0, lane);
      if (bad && lane == 0 && s_int[0] == 0) s_int[0] = c0 + bad;
      if (prof && lane == 0) prof[16 + p] = clock64();
    } else if (warp <= 4) {
      panel_solve_warp(Ls, rdiag, LT, rowL, c0, (warp - 1) * 32 + lane);
      if (prof && lane == 0 && (warp == 1 || warp == 4)) prof[(warp == 1 ? 20 : 24) + p] = clock64();
    } else if (p >= 1) {
      // warps 5-7 idle through the panel step.  Row panel p - 1 of L and Linv is final: out it goes, flag up
      // (everything it reads -- columns < 32 p of all rows, rdiag -- is left alone by this panel step).
      if (panelwise) {
        write_panel(p - 1, 5, 3);
        __threadfence();
        named_bar_sync(6, 96);
        if (warp == 5 && lane == 0) st_release(panel_flags + (p - 1), 1);
      }
      // then they finish the update of panel p - 1 on the columns beyond this panel (disjoint from
      // everything the panel step touches)
      rank32_update(Ls, rdiag, p - 1, p + 1, TILE / PANEL - 1, 5, 3, warp, gid, tig);
    }
    __syncthreads();
    GPAR_PROF(2 + 2 * p);
    // the 32 columns right behind the panel: what the next panel step reads
    if (p + 1 < TILE / PANEL) rank32_update(Ls, rdiag, p, p + 1, p + 1, 0, 8, warp, gid, tig);
    GPAR_PROF(3 + 2 * p);
  }
  __syncthreads();
  GPAR_PROF(10);

  // ---- ||L||_inf from the row sums gathered during the sweep (rows >= kb are identity padding) --------
  if (tid < TILE) {
    double v = (tid < kb) ? rowL[tid] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if (lane == 0) red[warp] = v;
  }
  GPAR_PROF(13);
  // ---- Linv (row major, ld = 128) and its row sums.  Linv[r][c] = Ls[c][r] (c < r), rdiag[r] at
  // c = r, zeros up to the end of the 32 x 32 diagonal block.  Lanes run along r (two lanes per
  // shared bank: conflict free at stride 132), every lane assembles 4 consecutive c = one 32-byte
  // sector of the output.  Unit = (32-row block rb, column group cg), 80 units over the 8 warps. --
  // set: PEERS_FIRST (single GPU: just the local copy) accumulates the row sums; PEERS_REST re-emits
  // the tile for the remaining peers of a multi-GPU run.
  auto write_linv = [&](int set) {
    double isum[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int it = 0; it < 10; ++it) {
      const int u = warp + 8 * it;
      const int rb = (u < 8) ? 0 : ((u < 24) ? 1 : ((u < 48) ? 2 : 3));
      const int cg = u - ((rb == 0) ? 0 : ((rb == 1) ? 8 : ((rb == 2) ? 24 : 48)));
      const int r = 32 * rb + lane, c = 4 * cg;
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double off = Ls[(c + q) * DLD + r];
        v[q] = (c + q < r) ? off : ((c + q == r) ? rdiag[r] : 0.0);
      }
      if (r < kb) {
        double* dst = ws + r * TILE + c;
        for (int pr = 0; pr < (multi ? pe->world : 1); ++pr) {
          if (multi && !peer_in_set(pe, pr, set)) continue;
          double* pd = multi ? peer_ptr(dst, pe->delta[pr]) : dst;  // delta[rank] == 0: the local copy
          *reinterpret_cast<double2*>(pd) = make_double2(v[0], v[1]);
          *reinterpret_cast<double2*>(pd + 2) = make_double2(v[2], v[3]);
        }
        const double s = (fabs(v[0]) + fabs(v[1])) + (fabs(v[2]) + fabs(v[3]));
        if (rb == 0) isum[0] += s; else if (rb == 1) isum[1] += s; else if (rb == 2) isum[2] += s; else isum[3] += s;
      }
    }
    if (set != PEERS_REST) {  // partial row sums of this warp -> LT scratch [8][128]
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) LT[warp * TILE + 32 * rb + lane] = isum[rb];
    }
  };
  if (panel_end) write_panel(TILE / PANEL - 1, 0, 8); else write_linv(PEERS_FIRST);
  GPAR_PROF(14);
  __syncthreads();
  if (tid < TILE) {
    double s = 0.0;
    const double* part = panel_end ? isumP : LT;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += (tid < kb) ? part[w * TILE + tid] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, off));
    if (lane == 0) red[8 + warp] = s;
  }
  __syncthreads();
  if (tid == 0) {
    double mL = 0.0, mI = 0.0;
    for (int i = 0; i < 4; ++i) mL = fmax(mL, red[i]);
    for (int i = 0; i < 4; ++i) mI = fmax(mI, red[8 + i]);
    const double kappa = mL * mI;
    const int refine = (kappa > REFINE_KAPPA || !(kappa == kappa)) ? 1 : 0;
    *flag_out = refine ? 1.0 : 0.0;
    if (multi)
      for (int pr = 0; pr < pe->world; ++pr)
        if (pr != pe->rank) *peer_ptr(flag_out, pe->delta[pr]) = refine ? 1.0 : 0.0;
    s_int[1] = refine;
    if (s_int[0] != 0 && *info_b == 0) *info_b = static_cast<int32_t>(j0) + s_int[0];
  }
  fence_publish(pe);
  __syncthreads();
  const bool refine = s_int[1] != 0;
  if (panel_end) {  // every row panel of L and Linv is out (the last one just now), the refine flag is set
    if (tid == 0) {
      st_release(panel_flags + (TILE / PANEL - 1), 1);
      if (ready_flag) st_release(ready_flag, 1);
    }
    GPAR_PROF(11);
    GPAR_PROF(12);
    return;
  }
  // Well-conditioned tile on a single GPU: consumers that only read Linv could be released before L_kk is
  // written.  The HEAD task of the next column reads the row panels of L_kk (head_pipelined), so with several
  // GPUs (no panel flags) the flag goes up once both are out.
  const bool early = ready_flag && !refine && !multi;
  if (early && tid == 0) publish_flag(ready_flag, pe, PEERS_FIRST);
  if (multi && pe->world > 2) write_linv(PEERS_REST);
  GPAR_PROF(11);
  // ---- L (row major).  Only the lower triangle is stored, plus zeros in the strictly upper part of
  // each 32 x 32 diagonal block: every consumer (MODE 2 GEMMs skip whole 32-column halves per
  // k-chunk, backsolve starts at the 8 x 8 diagonal block) stays inside that region. ----------------
  {
    const bool vec_ok = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(Atile) & 15) == 0);
    const int c = 4 * lane;
#pragma unroll 4
    for (int r = warp; r < kb; r += 8) {
      if (c > (r | 31) || c >= kb) continue;
      const double2 v0 = *reinterpret_cast<const double2*>(Ls + r * DLD + c);
      const double2 v1 = *reinterpret_cast<const double2*>(Ls + r * DLD + c + 2);
      double v[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c + q > r) v[q] = 0.0;
      double* dst = Atile + (int64_t)r * lda + c;
      for (int pr = 0; pr < (multi ? pe->world : 1); ++pr) {
        double* pd = multi ? peer_ptr(dst, pe->delta[pr]) : dst;  // delta[rank] == 0: the local copy
        if (vec_ok && c + 3 < kb) {
          *reinterpret_cast<double2*>(pd) = make_double2(v[0], v[1]);
          *reinterpret_cast<double2*>(pd + 2) = make_double2(v[2], v[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (c + q < kb) pd[q] = v[q];
        }
      }
    }
  }
  if (ready_flag && (!early || (multi && pe->world > 2))) {
    fence_publish(pe);
    __syncthreads();
    if (tid == 0) publish_flag(ready_flag, pe, early ? PEERS_REST : PEERS_ALL);
  }
  if (panelwise) {  // (several GPUs) the local copy is complete: last local panel flag
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(panel_flags + (TILE / PANEL - 1), 1);
  }
  GPAR_PROF(12);
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
  double* Atile = A + (int64_t)b * strideA + j0 * lda + j0;
  if (prof && threadIdx.x == 0) prof[0] = clock64();
  diag_load(smem_raw, Atile, lda, kb);
  diag_factor_core<false>(smem_raw, Atile, lda, kb, j0, wsb + (int64_t)kt * TILE * TILE,
                   wsb + (int64_t)nt_total * TILE * TILE + kt, info + b, nullptr, nullptr, prof);
}

// --------------------------------------------------------------------------------------
// T (rows x kb) <- T Linv^T for every 128-row tile of a row block.
// --------------------------------------------------------------------------------------
// X = T L_kk^-T for one 128-row tile: X0 = T Linv^T; when the diagonal block is flagged
// ill-conditioned, one refinement step R = T - X0 L_kk^T, X = X0 + R Linv^T (X0 parked in a
// per-CTA scratch tile) restores backward stability.  On return T holds X and so does `acc` -- in the
// PERMUTED column layout of MODE 2 (acc_col<true>): consumers pass PERM = true.
__device__ __forceinline__ void tile_solve(GemmStage* stages, double* __restrict__ T, int64_t ldt, int valid, int kb,
                                           const double* __restrict__ Linv, const double* __restrict__ Lkk,
                                           int64_t ldl, bool refine, double* __restrict__ scratch, Acc& acc) {
  acc_zero(acc);
  gemm_nt_mainloop<2>(stages, T, ldt, valid, Linv, TILE, kb, kb, acc);
  if (!refine) {
    store_tile<0, true>(T, ldt, valid, kb, acc, false);
    return;
  }
  store_tile<0, true>(scratch, TILE, valid, kb, acc, false);  // X0
  __threadfence();
  __syncthreads();
  acc_zero(acc);
  gemm_nt_mainloop<2>(stages, scratch, TILE, valid, Lkk, ldl, kb, kb, acc);  // X0 L_kk^T
  store_tile<1, true>(T, ldt, valid, kb, acc, false);                           // T <- R = T - X0 L_kk^T
  __threadfence();
  __syncthreads();
  acc_zero(acc);
  gemm_nt_mainloop<2>(stages, T, ldt, valid, Linv, TILE, kb, kb, acc);  // R Linv^T
  acc_add_tile<true>(acc, scratch, TILE, valid, kb);                         // + X0 (each thread re-reads its own stores)
  store_tile<0, true>(T, ldt, valid, kb, acc, false);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
trsm_tile_kernel(double* __restrict__ T1, int64_t ldt1, int64_t rows1, int64_t strideT1, int nt1,
                 double* __restrict__ T2, int64_t ldt2, int64_t rows2, int64_t strideT2, int kb,
                 const double* __restrict__ Linv, int64_t strideW, const double* __restrict__ Lkk, int64_t ldl,
                 int64_t strideL, const double* __restrict__ flag, double* __restrict__ scratch,
                 int64_t strideScratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  pipe_init();
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
  Acc acc;
  tile_solve(stages, T, ldt, valid, kb, Linv + (int64_t)b * strideW, Lkk + (int64_t)b * strideL, ldl, refine,
             scratch + (int64_t)b * strideScratch + (int64_t)blockIdx.x * TILE * TILE, acc);
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
  int tri_k;  // != 0: both operands are UPPER triangular (zero for k < row): tile (ti, tj) starts at k = ti * 128
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
  pipe_init();
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
  const int kstart = (p.tri_k && !second) ? ti * TILE : 0;
  gemm_nt_mainloop<0>(stages, Ap + kstart, lda, rows, Bp + kstart, p.ldb, cols, p.K - kstart, acc);
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
// Row blocks: whole waves of 128-row blocks (one CTA per SM), then the rows that are left as blocks of
// 32 / 64 / 96 rows -- the smallest height that still fits them in ONE more wave -- instead of a last wave of
// 128-row blocks that leaves most SMs idle (C5 on 8 GPUs: 512 blocks = 3.46 waves -> 3 waves + 136 blocks of
// 64 rows).  Rows are independent, so the partition changes no bit of the result.
struct RowPlan {
  int64_t full_blocks;  // 128-row blocks [0, full_blocks)
  int64_t tail_blocks;  // then blocks of tail_h rows
  int tail_h;
};
__host__ __device__ inline RowPlan trsm_row_plan(int64_t nb, int sms) {
  RowPlan p;
  const int64_t T = (nb + TILE - 1) / TILE;
  const int64_t r = T % sms;
  p.full_blocks = T - r;
  p.tail_blocks = 0;
  p.tail_h = TILE;
  if (r > 0) {
    const int64_t rows = nb - p.full_blocks * TILE;
    int64_t h = 32 * ((rows + 32 * (int64_t)sms - 1) / (32 * (int64_t)sms));
    if (h > TILE) h = TILE;
    p.tail_h = (int)h;
    p.tail_blocks = (rows + h - 1) / h;
  }
  return p;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
trsm_rows_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ ws,
                 double* __restrict__ B, int64_t ldb, int64_t nb, double* __restrict__ scratch, int64_t full_blocks,
                 int tail_h) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  pipe_init();
  const int ti = blockIdx.x;
  const int64_t row0 = ti < full_blocks ? (int64_t)ti * TILE : full_blocks * TILE + (ti - full_blocks) * (int64_t)tail_h;
  double* Brow = B + row0 * ldb;
  const int valid = static_cast<int>(min64(ti < full_blocks ? TILE : tail_h, nb - row0));
  const int nt = static_cast<int>((n + TILE - 1) / TILE);
  const double* flags = ws + (int64_t)nt * TILE * TILE;
  for (int j = 0; j < nt; ++j) {
    const int kb = static_cast<int>(min64(TILE, n - (int64_t)j * TILE));
    if (j > 0) {
      Acc acc;
      acc_zero(acc);
      gemm_nt_mainloop<4>(stages, Brow, ldb, valid, L + (int64_t)j * TILE * ldl, ldl, kb, j * TILE, acc);
      store_tile<1>(Brow + (int64_t)j * TILE, ldb, valid, kb, acc, false);
      __threadfence();
      __syncthreads();
    }
    Acc xacc;
    tile_solve(stages, Brow + (int64_t)j * TILE, ldb, valid, kb, ws + (int64_t)j * TILE * TILE,
               L + (int64_t)j * TILE * ldl + (int64_t)j * TILE, ldl, flags[j] != 0.0,
               scratch + (int64_t)ti * TILE * TILE, xacc);
    __threadfence();
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------------
// U = L^-T (upper triangular, row major) for the inverse A^-1 = U U^T of the fit gradients: the
// sweep of trsm_rows_kernel applied to the identity, with the structure exploited -- row block ti
// of I L^-T is zero left of column block ti, so the sweep and every K-loop start there (n^3 / 3
// instead of n^3 flops) -- and 32-row CTAs (4 per 128-row block: the rows of a block are
// independent) so that 4 nt CTAs share the work instead of nt.
// --------------------------------------------------------------------------------------
constexpr int TRTRI_ROWS = 32;
__global__ void __launch_bounds__(GEMM_THREADS, 1)
trtri_rows_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ ws,
                  double* __restrict__ U, int64_t ldu, double* __restrict__ scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  pipe_init();
  constexpr int SUB = TILE / TRTRI_ROWS;
  // heaviest row blocks first: block 0 sweeps every column, the last block only its own
  const int ti = blockIdx.x / SUB, sub = blockIdx.x % SUB;
  const int64_t row0 = (int64_t)ti * TILE + sub * TRTRI_ROWS;
  if (row0 >= n) return;
  double* Brow = U + row0 * ldu;
  const int valid = static_cast<int>(min64(TRTRI_ROWS, n - row0));
  const int nt = static_cast<int>((n + TILE - 1) / TILE);
  const double* flags = ws + (int64_t)nt * TILE * TILE;
  for (int j = ti; j < nt; ++j) {
    const int kb = static_cast<int>(min64(TILE, n - (int64_t)j * TILE));
    if (j > ti) {
      Acc acc;
      acc_zero(acc);
      gemm_nt_mainloop(stages, Brow + (int64_t)ti * TILE, ldu, valid,
                       L + (int64_t)j * TILE * ldl + (int64_t)ti * TILE, ldl, kb, (j - ti) * TILE, acc);
      store_tile<1>(Brow + (int64_t)j * TILE, ldu, valid, kb, acc, false);
      __threadfence();
      __syncthreads();
    }
    Acc xacc;
    tile_solve(stages, Brow + (int64_t)j * TILE, ldu, valid, kb, ws + (int64_t)j * TILE * TILE,
               L + (int64_t)j * TILE * ldl + (int64_t)j * TILE, ldl, flags[j] != 0.0,
               scratch + (int64_t)blockIdx.x * TILE * TILE, xacc);
    __threadfence();
    __syncthreads();
  }
}

__global__ void set_identity_kernel(double* __restrict__ U, int64_t ldu, int64_t n) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ldu) return;
  const int64_t r = idx / ldu, c = idx % ldu;
  U[idx] = (r == c) ? 1.0 : 0.0;
}

// --------------------------------------------------------------------------------------
// v3: persistent left-looking tile-dataflow Cholesky.  One CTA per SM pulls tile tasks from a
// global ticket counter.  Four kinds of task (tile = 128, k = tile row, j = tile column):
//   PLAIN (i, j), i >= j + 2 or appended rows:
//       acc = sum_{l<j} L_il L_jl^T  (one long-K DMMA GEMM that consumes the k-tiles as their
//       "ready" flags come up), T = A_ij - acc,
//       then, once L_jj is ready, L_ij = T L_jj^-T.
//   PRE (k):   A_kk -= sum_{l<k-1} L_kl L_kl^T      (everything of the diagonal update that does
//              not depend on the previous column), published through pready[k].
//   HEAD (k):  the whole critical chain of column k-1 -> k in ONE CTA, operands staying on chip:
//              T = A_{k,k-1} - sum_{l<k-1} ..., wait for L_{k-1,k-1}, X = T L^-T (published as
//              L_{k,k-1}), X parked in shared memory, S = X X^T (lower) straight out of shared
//              memory, tile = A_kk(PRE) - S assembled in shared memory, diagonal factor + inverse,
//              published.  Per column the chain is solve + syrk + factor on one SM with no global
//              round trip in between (v2: three tasks on three SMs, 125 us per column).
//   D0:        factor of the first diagonal tile.
// Ticket order: [D0, HEAD(1)] then for every column g: HEAD(g+2), PLAIN(g+2, g), PRE(g+2),
// PLAIN(g+3.., g) (in the tail each of them preceded by its split-K parts, see df_split).  Every
// dependency of a task has a smaller ticket, except that HEAD(g+2) needs the two tiles ticketed right
// behind it; since tickets are handed out in order those are always held by a running CTA (or the
// next free one), so a spinning CTA only ever waits on running or finished work: deadlock-free for
// any grid >= 3 * GPAR_SPLIT_MAX + 1 (3 tiles x parts + 1) without a co-residency requirement; smaller grids only occur
// for matrices too small to be split.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void wait_ready(const int* flag, bool sys = false) {
  if (sys) {
    while (ld_acquire_sys(flag) == 0) __nanosleep(40);  // the flag (and its tile) may have come from a peer GPU
  } else {
    while (ld_acquire(flag) == 0) __nanosleep(40);
  }
}

__device__ __forceinline__ void wait_count(const int* counter, int target) {
  while (ld_acquire(counter) < target) __nanosleep(40);
}

// gemm_nt_mainloop over K = 128 * ktiles with per-k-tile dependency waits on readyA[kt], readyB[kt]:
// every thread acquires the flags before it issues its copies of the first chunk of a k-tile.
// (Measured alternatives: polling the flags in the background so that landed chunks are never
// held up by a late flag was 4-12 % slower in the throughput-bound regime and no faster in the
// chain-bound one; a CTA-wide barrier per chunk instead of the mbarrier ring costs 6 %.)
#ifdef GPAR_DF_PROF
#define g_dfp_wait dfp_wait_ptr()
__device__ __forceinline__ long long* dfp_wait_ptr() {
  __shared__ long long s_wait;
  return &s_wait;
}
#endif
// Ctile (optional): the tile the epilogue will read-modify-write.  Its lines are pulled into L2 when the last
// k-tile starts (~17 us ahead), so that the epilogue's loads do not pay the HBM latency four times in a row.
__device__ __forceinline__ void prefetch_tile_l2(const double* T, int64_t ldt, int rows) {
  // 128 rows x 1 KB: 8 lines of 128 B per row, 4 lines per thread
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = threadIdx.x + q * GEMM_THREADS;  // 0..1023
    const int r = idx >> 3, seg = idx & 7;
    if (r < rows) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(T + (int64_t)r * ldt + seg * 16));
  }
}
template <int MODE>
__device__ __forceinline__ void gemm_nt_mainloop_dep(GemmStage* stages, const double* __restrict__ Ap, int64_t lda,
                                                     int validA, const double* __restrict__ Bp, int64_t ldb,
                                                     int validB, int K, Acc& acc, const int* readyA,
                                                     const int* readyB, bool sys, const double* Ctile = nullptr,
                                                     int64_t ldc = 0) {
  const int last_kt = (K + TILE - 1) / TILE - 1;
  gemm_nt_pipe<MODE>(stages, Ap, lda, validA, Bp, ldb, validB, K, acc, [&](int kt) {
#ifndef GPAR_NO_CPREFETCH
    if (Ctile != nullptr && kt == last_kt) prefetch_tile_l2(Ctile, ldc, validA);
#endif
#ifdef GPAR_DF_PROF
    const long long t0_ = clock64();
#endif
    wait_ready(readyA + kt, sys);
    if (readyB != readyA) wait_ready(readyB + kt, sys);
#ifdef GPAR_DF_PROF
    if (threadIdx.x == 0 && g_dfp_wait) *g_dfp_wait += clock64() - t0_;
#endif
  });
}
// X (accumulator layout of tile_solve: permuted columns) -> shared operand tile Xs[128][DLD].
__device__ __forceinline__ void acc_to_smem(double* __restrict__ Xs, const Acc& acc) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<double2*>(Xs + (acc_row(wm, i) + gid) * DLD + acc_col<true>(wn, j) + 2 * tig) =
          make_double2(acc[i][j][0], acc[i][j][1]);
}

// acc += lower part of Xs[:, k_lo:k_hi] Xs[:, k_lo:k_hi]^T, both operands straight from the shared tile.  Only
// the 8 x 32 blocks that touch the lower triangle are computed (static per-warp block lists).
__device__ __forceinline__ void syrk_from_smem_range(const double* __restrict__ Xs, Acc& acc, int k_lo, int k_hi) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const double* pa = Xs + (wm * 8 + gid) * DLD + tig;
  const double* pb = Xs + (wn * 64 + gid) * DLD + tig;
  if (wn == 0) {  // columns 0..63: every row group, except (i = 0, columns 32..63)
#pragma unroll 2
    for (int kk = k_lo; kk < k_hi; kk += 4) {
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = pa[i * 32 * DLD + kk];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = pb[j * 8 * DLD + kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#pragma unroll
      for (int i = 1; i < 4; ++i)
#pragma unroll
        for (int j = 4; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  } else {  // columns 64..127: row groups i = 2 (columns 64..95) and i = 3 (all)
#pragma unroll 2
    for (int kk = k_lo; kk < k_hi; kk += 4) {
      double a[2], b[8];
      a[0] = pa[2 * 32 * DLD + kk];
      a[1] = pa[3 * 32 * DLD + kk];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = pb[j * 8 * DLD + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[2][j][0], acc[2][j][1], a[0], b[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884(acc[3][j][0], acc[3][j][1], a[1], b[j]);
    }
  }
}
__device__ __forceinline__ void syrk_from_smem(const double* __restrict__ Xs, Acc& acc) {
  acc_zero(acc);
  syrk_from_smem_range(Xs, acc, 0, TILE);
}

// ---- HEAD task, pipelined behind the diagonal factor of the previous column ------------------------
// X L_jj^T = T by blocked forward substitution over four 32-column blocks, in place in the shared tile Ts:
//     X_cb = (T_cb - sum_{cb' < cb} X_cb' L_{cb,cb'}^T) Linv_{cb,cb}^T,
// block cb needing only row panel cb of L_jj and the diagonal 32-block of L_jj^-1 -- exactly what
// diag_factor_core publishes after its panel step cb (panel_flags).  S += X_cb X_cb^T follows each block, so
// that when the last panel flag arrives one block solve and a quarter of the SYRK are left instead of
// solve + SYRK of the whole tile (17 + 11 us on the column chain).
typedef double AccP[4][2][2];  // 128 x 32 output: row groups acc_row(wm, i), column tiles wn * 16 + 8 j

// accp += As[:, 0:K] Bs[:, 0:K]^T; As: 128 rows, stride DLD; Bs: 32 rows, stride ldb (= 4 mod 16 doubles).
__device__ __forceinline__ void mma_panel(const double* __restrict__ As, const double* __restrict__ Bs, int ldb, int K,
                                          AccP& acc, int wm, int wn, int gid, int tig) {
  const double* pa = As + (wm * 8 + gid) * DLD + tig;
  const double* pb = Bs + (wn * 16 + gid) * ldb + tig;
#pragma unroll 4
  for (int kk = 0; kk < K; kk += 4) {
    double a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = pa[i * 32 * DLD + kk];
#pragma unroll
    for (int j = 0; j < 2; ++j) b[j] = pb[j * 8 * ldb + kk];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

constexpr int LDD = PANEL + 4;  // stride of the 32 x 32 diagonal inverse block in shared memory
// shared layout of the pipelined HEAD: Ts[TILE][DLD] | Lp[PANEL][DLD] | Ld[PANEL][LDD]
constexpr size_t HEADP_SMEM_BYTES = sizeof(double) * (TILE * DLD + PANEL * DLD + PANEL * LDD);

// T (global, rows x 128) -> Ts (zeros beyond `rows`).  All threads; ends with a barrier.
__device__ __forceinline__ void load_tile_to_smem(double* __restrict__ Ts, const double* __restrict__ T, int64_t ldt,
                                                  int rows) {
#pragma unroll 8
  for (int q = 0; q < 32; ++q) {
    const int idx = threadIdx.x + q * GEMM_THREADS;  // row r, double2 column c
    const int r = idx >> 6, c = (idx & 63) * 2;
    cp_async16(&Ts[r * DLD + c], (r < rows) ? (T + (int64_t)r * ldt + c) : T, (r < rows) ? 16 : 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
}

// Ts (shared) -> dst (global tile, rows x 128).  All threads; no barrier.
__device__ __forceinline__ void store_tile_from_smem(double* __restrict__ dst, int64_t ldd, int rows,
                                                     const double* __restrict__ Ts) {
#pragma unroll 8
  for (int q = 0; q < 32; ++q) {
    const int idx = threadIdx.x + q * GEMM_THREADS;
    const int r = idx >> 6, c = (idx & 63) * 2;
    if (r < rows) *reinterpret_cast<double2*>(dst + (int64_t)r * ldd + c) = *reinterpret_cast<const double2*>(Ts + r * DLD + c);
  }
}

// Ts (shared) = T (global tile, rows x 128) - acc (accumulator layout of MODE 0); zeros beyond `rows`.
// All threads; no barrier.
__device__ __forceinline__ void tile_sub_to_smem(double* __restrict__ Ts, const double* __restrict__ T, int64_t ldt,
                                                 int rows, const Acc& acc) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int ih = 0; ih < 2; ++ih) {
    double2 old[2][8];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int r = acc_row(wm, 2 * ih + ii) + gid;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        old[ii][j] = (r < rows) ? __ldcg(reinterpret_cast<const double2*>(T + (int64_t)r * ldt + wn * 64 + j * 8 + 2 * tig))
                                : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int i = 2 * ih + ii;
      const int r = acc_row(wm, i) + gid;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<double2*>(Ts + r * DLD + wn * 64 + j * 8 + 2 * tig) =
            (r < rows) ? make_double2(old[ii][j].x - acc[i][j][0], old[ii][j].y - acc[i][j][1]) : make_double2(0.0, 0.0);
    }
  }
}

// acc = -T (global tile, rows x 128; zeros beyond `rows`), accumulator layout of MODE 0.
__device__ __forceinline__ void acc_load_neg(Acc& acc, const double* __restrict__ T, int64_t ldt, int rows) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = acc_row(wm, i) + gid;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double2 v = (r < rows) ? __ldcg(reinterpret_cast<const double2*>(T + (int64_t)r * ldt + wn * 64 + j * 8 + 2 * tig))
                                   : make_double2(0.0, 0.0);
      acc[i][j][0] = -v.x;
      acc[i][j][1] = -v.y;
    }
  }
}
// Ts (shared) = -acc.
__device__ __forceinline__ void acc_neg_to_smem(double* __restrict__ Ts, const Acc& acc) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<double2*>(Ts + (acc_row(wm, i) + gid) * DLD + wn * 64 + j * 8 + 2 * tig) =
          make_double2(-acc[i][j][0], -acc[i][j][1]);
}

// The four blocks.  Ts holds T on entry and X on return; S (zeroed here) holds the lower part of X X^T.
// Ljj: tile (j, j) of the matrix (global, ld ldl); Linv: its inverse tile (global, ld 128).
// wait_panel(cb) blocks until row panel cb of both is in global memory.
template <typename WaitFn>
__device__ __forceinline__ void head_blocks(unsigned char* smem_raw, const double* __restrict__ Ljj, int64_t ldl,
                                            const double* __restrict__ Linv, Acc& S, WaitFn wait_panel) {
  double* Ts = reinterpret_cast<double*>(smem_raw);
  double* Lp = Ts + TILE * DLD;
  double* Ld = Lp + PANEL * DLD;
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  acc_zero(S);
#pragma unroll 1
  for (int cb = 0; cb < TILE / PANEL; ++cb) {
    const int c0 = PANEL * cb;
    wait_panel(cb);
    // row panel cb of L_jj (columns < c0) and the diagonal 32-block of Linv (zeros above its diagonal)
    for (int idx = threadIdx.x; idx < PANEL * (c0 / 2); idx += GEMM_THREADS) {
      const int r = idx / (c0 / 2), c = (idx % (c0 / 2)) * 2;
      cp_async16(&Lp[r * DLD + c], Ljj + (int64_t)(c0 + r) * ldl + c, 16);
    }
    for (int idx = threadIdx.x; idx < PANEL * (PANEL / 2); idx += GEMM_THREADS) {
      const int r = idx / (PANEL / 2), c = (idx % (PANEL / 2)) * 2;
      cp_async16(&Ld[r * LDD + c], Linv + (int64_t)(c0 + r) * TILE + c0 + c, 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    AccP u;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) u[i][j][0] = u[i][j][1] = 0.0;
    if (cb > 0) mma_panel(Ts, Lp, DLD, c0, u, wm, wn, gid, tig);  // sum_{cb' < cb} X_cb' L_{cb,cb'}^T
    // V = T_cb - U, in place (nobody reads the columns of block cb during the update above)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double2* pv = reinterpret_cast<double2*>(Ts + (acc_row(wm, i) + gid) * DLD + c0 + wn * 16 + 8 * j + 2 * tig);
        double2 v = *pv;
        v.x -= u[i][j][0];
        v.y -= u[i][j][1];
        *pv = v;
      }
    __syncthreads();
    AccP x;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) x[i][j][0] = x[i][j][1] = 0.0;
    mma_panel(Ts + c0, Ld, LDD, PANEL, x, wm, wn, gid, tig);  // X_cb = V Linv_{cb,cb}^T
    __syncthreads();  // every warp has read V
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        *reinterpret_cast<double2*>(Ts + (acc_row(wm, i) + gid) * DLD + c0 + wn * 16 + 8 * j + 2 * tig) =
            make_double2(x[i][j][0], x[i][j][1]);
    __syncthreads();
    syrk_from_smem_range(Ts, S, c0, c0 + PANEL);
  }
}

// Ls (shared, diagonal-factor layout) = lower(T - acc), zeros above the diagonal, identity
// padding beyond kb.  T is the diagonal tile in global memory (ld = ldt).
__device__ __forceinline__ void assemble_diag(double* __restrict__ Ls, const double* __restrict__ T, int64_t ldt,
                                              int kb, const Acc& acc) {
  const int warp = canonical_warp(), lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2, gid = lane >> 2, tig = lane & 3;
  const bool vec = ((ldt & 1) == 0) && ((reinterpret_cast<uintptr_t>(T) & 15) == 0);
#pragma unroll
  for (int ih = 0; ih < 2; ++ih) {
    double2 old[2][8];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int r = acc_row(wm, 2 * ih + ii) + gid;
      const double* trow = T + (int64_t)r * ldt;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = wn * 64 + j * 8 + 2 * tig;
        const bool ok0 = (r < kb) && (c <= r), ok1 = (r < kb) && (c + 1 <= r);
        old[ii][j] = make_double2(0.0, 0.0);
        if (ok0 && ok1 && vec) {
          old[ii][j] = __ldcg(reinterpret_cast<const double2*>(trow + c));
        } else if (ok0) {
          old[ii][j].x = __ldcg(trow + c);
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int i = 2 * ih + ii;
      const int r = acc_row(wm, i) + gid;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = wn * 64 + j * 8 + 2 * tig;
        double v0 = 0.0, v1 = 0.0;
        if (r < kb) {
          if (c <= r) v0 = old[ii][j].x - acc[i][j][0];
          if (c + 1 <= r) v1 = old[ii][j].y - acc[i][j][1];
        } else {
          if (c == r) v0 = 1.0;
          if (c + 1 == r) v1 = 1.0;
        }
        *reinterpret_cast<double2*>(Ls + r * DLD + c) = make_double2(v0, v1);
      }
    }
  }
}

// The phases of a tile task as out-of-line functions, each with its own accumulators and its own register
// allocation.  Inlined into one kernel body they compete for the 255 registers of the persistent kernel and the
// main K-loop pays for every line added elsewhere (18.5 instead of 17.75 us per k-tile with the pipelined HEAD
// inlined).  No accumulator crosses a call: an Acc passed by reference would pin the caller's accumulators to
// local memory.
struct TaskCtx {
  double* rowi; int64_t ldi; int valid;      // tile row i of the matrix (or of the appended rows)
  const double* rowj; int64_t lda; int kb;   // tile row j
  int i, j, k0, k1a, k1, part;               // K-range of this part: [k0, k1a) through the ring (+ [k1a, k1) for a HEAD)
  bool pre, head, multi;
  const int* ready_i; const int* ready_j;    // ready flags of the two tile rows
  int* ready_ij;                             // flag of tile (i, j)
  int* pcount;                               // K-parts already subtracted from tile (i, j)
  const int* pfl;                            // panel flags of tile (j, j) (single GPU)
  const double* Linv; const double* refine_flag;
  const int* pre_count; int pre_parts;       // HEAD: K-part counter of PRE(i) and its final value (0: no PRE)
  double* scratch;                           // per-CTA scratch tile (refined solves)
};

// The whole chain of a HEAD task after its K-loop, out of line.  Pipelined path: newest k-tile into the shared tile,
// four solve + SYRK blocks behind the panel flags of L_jj, X stored to T (and to the next owner's T) and
// published, diagonal tile assembled in shared memory (ready for diag_factor_core).  When L_jj turns out
// ill-conditioned (or with GPAR_HEAD_UNPIPELINED) the round-1 path: T completed in global memory, refined solve
// through the inverse tile, X X^T out of shared memory, assembly.
template <bool MULTI>
__device__ __noinline__ void head_chain(unsigned char* smem_raw, const TaskCtx& h, const Peers* pe) {
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  double* Xs = reinterpret_cast<double*>(smem_raw);
  double* T = h.rowi + (int64_t)h.j * TILE;
  const bool multi = MULTI && h.multi;
  Acc acc;
#ifndef GPAR_HEAD_UNPIPELINED
  if (h.k1 > h.k1a) {
    // the newest k-tile: T - L_{i,k1-1} L_{j,k1-1}^T goes straight into the shared tile (no global round trip).  The
    // accumulators start at -T (loaded before the wait for L_{j,k1-1}: off the chain), so that what is left after
    // the K = 128 loop is a negation on the way into shared memory.
    acc_load_neg(acc, T, h.ldi, h.valid);
    gemm_nt_mainloop_dep<0>(stages, h.rowi + (int64_t)h.k1a * TILE, h.ldi, h.valid, h.rowj + (int64_t)h.k1a * TILE, h.lda, h.kb,
                            TILE, acc, h.ready_i + h.k1a, h.ready_j + h.k1a, h.multi);
    acc_neg_to_smem(Xs, acc);
    __syncthreads();
  } else {
    load_tile_to_smem(Xs, T, h.ldi, h.valid);
  }
  // several GPUs: panel flags exist on the rank that factored L_jj only; behind a remote factor the blocks wait
  // for the tile's (pushed) ready flag -- same arithmetic either way
  const bool remote = multi && ((h.j / pe->row_block) % pe->world) != pe->rank;
  head_blocks(smem_raw, h.rowj + (int64_t)h.j * TILE, h.lda, h.Linv, acc, [&](int cb) {
    if (remote) {
      if (cb == 0) wait_ready(h.ready_j + h.j, true);
    } else {
      wait_ready(h.pfl + cb);
    }
  });
  if (__ldcg(h.refine_flag) == 0.0) {
    store_tile_from_smem(T, h.ldi, h.valid, Xs);
    if (multi) store_tile_from_smem(peer_ptr(T, pe->delta[(pe->rank + 1) % pe->world]), h.ldi, h.valid, Xs);
    fence_publish(pe);
    __syncthreads();  // (also: every warp is done reading Xs)
    if (threadIdx.x == 0) publish_flag(h.ready_ij, pe, PEERS_FIRST);
    if (h.pre_parts > 0) wait_count(h.pre_count, h.pre_parts);  // every K-part of PRE(k) has been subtracted
    assemble_diag(Xs, h.rowi + (int64_t)h.i * TILE, h.lda, h.valid, acc);
    __syncthreads();
    return;
  }
  // ill-conditioned L_jj: nothing has been stored or published.  The newest k-tile has to reach T in global memory.
  __syncthreads();
  if (h.k1 > h.k1a) {
    acc_zero(acc);
    gemm_nt_mainloop_dep<0>(stages, h.rowi + (int64_t)h.k1a * TILE, h.ldi, h.valid, h.rowj + (int64_t)h.k1a * TILE, h.lda, h.kb,
                            TILE, acc, h.ready_i + h.k1a, h.ready_j + h.k1a, h.multi);
    store_tile<1>(T, h.ldi, h.valid, h.kb, acc, false);
    __threadfence();
    __syncthreads();
  }
#endif
  wait_ready(h.ready_j + h.j, multi);
  const bool refine = __ldcg(h.refine_flag) != 0.0;
  tile_solve(stages, T, h.ldi, h.valid, h.kb, h.Linv, h.rowj + (int64_t)h.j * TILE, h.lda, refine, h.scratch, acc);
  if (multi) store_tile<0, true>(peer_ptr(T, pe->delta[(pe->rank + 1) % pe->world]), h.ldi, h.valid, h.kb, acc, false);
  acc_to_smem(Xs, acc);  // the ring is idle: park X as the SYRK operand
  fence_publish(pe);
  __syncthreads();
  if (threadIdx.x == 0) publish_flag(h.ready_ij, pe, PEERS_FIRST);
  syrk_from_smem(Xs, acc);
  if (h.pre_parts > 0) wait_count(h.pre_count, h.pre_parts);
  __syncthreads();  // every warp is done reading Xs
  assemble_diag(Xs, h.rowi + (int64_t)h.i * TILE, h.lda, h.valid, acc);
  __syncthreads();
}

constexpr size_t DF_SMEM_BYTES = DIAG_SMEM_BYTES > GEMM_SMEM_BYTES ? DIAG_SMEM_BYTES : GEMM_SMEM_BYTES;
static_assert(HEADP_SMEM_BYTES <= DF_SMEM_BYTES, "pipelined HEAD tiles must fit the dataflow kernel's shared memory");
constexpr int DF_POOL_TILES = 192;  // scratch tiles for the persistent grid (>= SM count)

struct DfArgs {
  double* A; int64_t lda; int64_t n; int64_t strideA;
  double* B; int64_t ldb; int64_t nb; int64_t strideB;
  int batch; int nt; int nbt; int total_tasks;
  double* ws; int64_t strideWs;   // per matrix: inverse tiles, refine flags (see ws_* helpers)
  double* pool;                   // gridDim.x scratch tiles
  int32_t* info;
  int* ticket; int* ready;        // ready[(b * (nt + nbt) + i) * nt + j]
  int* pcount;                    // pcount[(b * (nt + nbt) + i) * nt + j]: K-parts already subtracted from tile (i, j)
                                  // (tile (k, k): parts of PRE(k))
  int* pflag;                     // pflag[(b * nt + k) * 4 + p]: row panel p of L_kk / L_kk^-1 is out (single GPU)
  int grid;                       // CTAs of the launch (enters the split-K rule)
  long long* prof;                // debug: globaltimer stamps of HEAD(nt/2), HEAD(nt/2 + 1) (or null)
  Peers peers;                    // multi-GPU: rank, world and the peers' address deltas (world == 1: unused)
};

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

enum { TASK_D0 = 0, TASK_HEAD = 1, TASK_PLAIN = 2, TASK_PRE = 3 };

// Split-K of the tail.  Towards the end of the sweep a column has fewer tiles than there are SMs while
// every tile still carries a K-loop over all earlier columns (1.2 ms at n = 8424): SMs run dry.  From the
// column where fewer than 1.5 tile tasks per SM remain, the K-range of every tile of the group is cut
// into S <= 4 parts; parts 0 .. S-2 are tasks of their own that accumulate their range and subtract it
// from the tile (in part order, sequenced by a per-tile counter, so the result does not depend on
// timing), the last part is the PLAIN / HEAD task itself.  Same function on host (ticket count) and device.
struct DfShape {
  int nt, nbt, batch;
  int grid;   // CTAs of the launch
  int world;  // ranks sharing the sweep (1 on a single GPU)
};
#ifndef GPAR_SPLIT_WANT_X10
#define GPAR_SPLIT_WANT_X10 15  // split once fewer than 1.5 tile tasks per SM remain
#endif
#ifndef GPAR_SPLIT_MAX
#define GPAR_SPLIT_MAX 4
#endif
#ifndef GPAR_SPLIT_MINK
#define GPAR_SPLIT_MINK 6       // k-tiles per part at least
#endif
__host__ __device__ __forceinline__ int df_split(const DfShape& sh, int g) {
  if (g < 4 * GPAR_SPLIT_MINK) return 1;  // short parts are not worth a read-modify-write of the tile
  const long r = sh.nt + sh.nbt - g;
  // (deliberately NOT divided by sh.world: the K-partition of a tile fixes its rounding, and keeping it
  //  independent of the number of ranks makes the sharded factor bit-identical to the single-GPU one)
  const long W = r * (r - 1) / 2 * sh.batch;  // tile tasks left
  const long want = (long)sh.grid * GPAR_SPLIT_WANT_X10 / 10;
  if (W >= want) return 1;
  long s = (want + W - 1) / (W > 0 ? W : 1);
  if (s > GPAR_SPLIT_MAX) s = GPAR_SPLIT_MAX;
  while (s > 1 && g / s < GPAR_SPLIT_MINK) --s;
  return static_cast<int>(s);
}

// Tasks per matrix: prologue and column group g (see the ticket order above).
__host__ __device__ __forceinline__ int df_prologue_tasks(int nt) { return nt > 1 ? 2 : 1; }
__host__ __device__ __forceinline__ int df_group_tiles(int nt, int rows_total, int g) {
  const int first = (g + 1 < nt) ? g + 2 : g + 1;
  const int nplain = rows_total - first;
  return (nplain > 0 ? nplain : 0) + ((g + 2 < nt) ? 2 : 0);
}
__host__ __device__ __forceinline__ int df_group_tasks(const DfShape& sh, int g) {
  return df_group_tiles(sh.nt, sh.nt + sh.nbt, g) * df_split(sh, g);
}
__host__ __device__ __forceinline__ long long df_total_tasks(const DfShape& sh) {
  long long total = df_prologue_tasks(sh.nt);
  for (int g = 0; g < sh.nt; ++g) total += df_group_tasks(sh, g);
  return total * sh.batch;
}

// ticket -> (kind, matrix b, tile row i, tile column j, K-part, number of parts); HEAD(k) comes back as
// (i = k, j = k - 1), PRE(k) as (k, k).  Parts of a tile carry consecutive tickets, the last part last.
// `cur` (optional): forward cursor {column group, tickets before it}.  Tickets handed to one CTA only grow, so
// the scan over the column groups resumes where the previous decode stopped (the scan from group 0 cost 3 us
// per task at nt = 66, 0.8 % of the sweep).
struct DfCursor { int g; int before; };
__host__ __device__ __forceinline__ void df_decode(int t, const DfShape& sh, int& kind, int& b, int& i, int& j,
                                                   int& part, int& nparts, DfCursor* cur = nullptr) {
  const int nt = sh.nt, batch = sh.batch;
  const int rows_total = nt + sh.nbt;
  const int npro = df_prologue_tasks(nt);
  part = 0;
  nparts = 1;
  if (t < batch * npro) {
    b = t / npro;
    if (t % npro == 0) { kind = TASK_D0; i = 0; j = 0; } else { kind = TASK_HEAD; i = 1; j = 0; }
    return;
  }
  int rem = t - batch * npro, g = 0, cnt = 0;
  if (cur != nullptr && rem >= cur->before) { g = cur->g; rem -= cur->before; }
  const int rem0 = rem, g0 = g;
  for (;; ++g) {
    cnt = df_group_tasks(sh, g);
    if (rem < batch * cnt) break;
    rem -= batch * cnt;
  }
  if (cur != nullptr) { cur->before += (g == g0) ? 0 : (rem0 - rem); cur->g = g; }
  b = rem / cnt;
  nparts = df_split(sh, g);
  const int idx = (rem % cnt) / nparts;
  part = (rem % cnt) % nparts;
  j = g;
  if (g + 2 < nt) {
    if (idx == 0) { kind = TASK_HEAD; i = g + 2; j = g + 1; }
    else if (idx == 1) { kind = TASK_PLAIN; i = g + 2; }
    else if (idx == 2) { kind = TASK_PRE; i = g + 2; j = g + 2; }
    else {
      // the remaining tiles of the column: rows g + 3 .. rows_total - 1, ascending in even columns and descending
      // in odd ones (boustrophedon): the row panels L_i,0..j streamed last by column g are the first ones column
      // g + 1 asks for, so that an L2-sized part of them is still resident (a fixed direction re-streams the
      // whole panel set from HBM once it exceeds the 126 MB L2: columns 24..42 at nt = 66)
      kind = TASK_PLAIN;
      i = (g & 1) ? (rows_total - 1) - (idx - 3) : g + idx;
    }
  } else {
    kind = TASK_PLAIN;
    i = ((g + 1 < nt) ? g + 2 : g + 1) + idx;
  }
}


// Per-CTA cycle accounting of the dataflow kernel (debug builds only: -DGPAR_DF_PROF, scripts/prof_budget.py):
// thread 0 charges the cycles since its previous mark to a category; the counters land in
// prof[64 + 16 * blockIdx.x ...] (gpar_debug_set_dataflow_prof).
#ifdef GPAR_DF_PROF
#define DFP_MARK(cat) do { if (threadIdx.x == 0) { const long long now_ = clock64(); s_dfp[cat] += now_ - s_dfp_t; s_dfp_t = now_; } } while (0)
#define DFP_ADD(cat, v) do { if (threadIdx.x == 0) s_dfp[cat] += (v); } while (0)
#else
#define DFP_MARK(cat) do { } while (0)
#define DFP_ADD(cat, v) do { } while (0)
#endif

template <bool MULTI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) potrf_dataflow_kernel(const DfArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_task;
  __shared__ Peers s_peers;
#ifdef GPAR_DF_PROF
  __shared__ long long s_dfp[16];
  __shared__ long long s_dfp_t;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 16; ++q) s_dfp[q] = 0;
    s_dfp[15] = globaltimer_ns();
    s_dfp_t = clock64();
    *dfp_wait_ptr() = 0;
  }
#endif
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) s_peers = p.peers;
  const bool multi = MULTI && p.peers.world > 1;  // MULTI = false: all peer code compiles out
  const Peers* pe = multi ? &s_peers : nullptr;
  const int rows_total = p.nt + p.nbt;
  double* scratch = p.pool + (int64_t)blockIdx.x * TILE * TILE;
  pipe_init();
  DfCursor cursor = {0, 0};
  for (;;) {
    if (tid == 0) s_task = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= p.total_tasks) break;
    int kind, b, i, j, part, nparts;
    const DfShape sh = {p.nt, p.nbt, p.batch, p.grid, p.peers.world};
    df_decode(t, sh, kind, b, i, j, part, nparts, &cursor);
    DFP_MARK(0);
    DFP_ADD(12, 1);
    // multi-GPU: tile rows are dealt block-cyclically; a rank only runs the tasks of its own rows
    if (multi && ((i / p.peers.row_block) % p.peers.world) != p.peers.rank) continue;
    double* Ab = p.A + (int64_t)b * p.strideA;
    int* ready_b = p.ready + (int64_t)b * rows_total * p.nt;
    double* wsb = p.ws + (int64_t)b * p.strideWs;
    double* flags = wsb + (int64_t)p.nt * TILE * TILE;
    long long* pf = nullptr;  // debug stamps of HEAD(nt/2) and HEAD(nt/2 + 1)
    if (p.prof && b == 0 && tid == 0 && kind == TASK_HEAD && (i == p.nt / 2 || i == p.nt / 2 + 1))
      pf = p.prof + 8 * (i - p.nt / 2);
    if (pf) pf[0] = globaltimer_ns();

    if (kind == TASK_D0) {
      const int kb = static_cast<int>(min64(TILE, p.n));
      diag_load(smem_raw, Ab, p.lda, kb);
      diag_factor_core<MULTI>(smem_raw, Ab, p.lda, kb, 0, wsb, flags, p.info + b, ready_b, pe, nullptr,
                              p.pflag + (int64_t)b * p.nt * 4);
      DFP_MARK(11);
    } else {
      // PRE (k): A_kk -= sum_{l<k-1} L_kl L_kl^T.  PLAIN (i, j) and HEAD (k = i, j = k - 1): T = A_ij - sum_{l<j}
      // L_il L_jl^T, then the solve.  The K-range [0, nk) of the tile is cut into `nparts` parts (1 except in
      // the tail of the sweep); this task accumulates part `part` and subtracts it from the tile once the
      // earlier parts have done so.
      const bool pre = kind == TASK_PRE;
      const int nk = pre ? i - 1 : j;
      const int k0 = (int)((long)part * nk / nparts), k1 = (int)((long)(part + 1) * nk / nparts);
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
      const int* ready_i = ready_b + (int64_t)i * p.nt;
      const int* ready_j = ready_b + (int64_t)j * p.nt;
      double* T = rowi + (int64_t)j * TILE;
      int* pcount = p.pcount + ((int64_t)b * rows_total + i) * p.nt + j;
      const bool head = kind == TASK_HEAD && part + 1 == nparts;  // (the last K-part of its tile is the HEAD itself)
#ifndef GPAR_HEAD_UNPIPELINED
      // A pipelined HEAD keeps the newest k-tile (the one that needs L_{k,k-1}, published by the HEAD of the previous
      // column a moment ago) out of this K-loop: it is applied last, straight into the shared tile (head_chain).
      const int k1a = (head && k1 > k0) ? k1 - 1 : k1;
#else
      const int k1a = k1;
#endif
      if (k1a > k0) {
        Acc acc;
        acc_zero(acc);
        // (the 8 x 32 masks of MODE 1 are slower than the full tile: their branches break the DMMA / LDS software
        //  pipeline; MODE 3 drops the upper-right 64 x 64 quadrant of a diagonal tile with straight-line bodies)
#ifndef GPAR_PRE_FULL
        if (pre && valid == TILE)
          gemm_nt_mainloop_dep<3>(stages, rowi + (int64_t)k0 * TILE, ldi, valid, rowj + (int64_t)k0 * TILE, p.lda, kb,
                                  (k1a - k0) * TILE, acc, ready_i + k0, ready_j + k0, multi, T, ldi);
        else
#endif
        gemm_nt_mainloop_dep<0>(stages, rowi + (int64_t)k0 * TILE, ldi, valid, rowj + (int64_t)k0 * TILE, p.lda, kb,
                                (k1a - k0) * TILE, acc, ready_i + k0, ready_j + k0, multi, T, ldi);
        DFP_MARK(1);
        DFP_ADD(13, k1a - k0);
        if (part > 0) wait_count(pcount, part);  // parts subtract in order: the rounding does not depend on timing
        store_tile<1>(T, ldi, valid, kb, acc, pre);
        __threadfence();
        __syncthreads();
      } else if (part > 0) {
        wait_count(pcount, part);
      }
      DFP_MARK(3);
      if (pre || part + 1 < nparts) {
        if (tid == 0) st_release(pcount, part + 1);
        continue;
      }
      if (pf) pf[1] = globaltimer_ns();
      if (head) {
        // The chain of column j -> j + 1 (head_chain, out of line): ends with the diagonal tile assembled in shared
        // memory.  With several GPUs there are no panel flags: the same arithmetic (the factor stays bit-identical
        // to the single-GPU one) behind the tile's ready flag.
        TaskCtx c;
        c.rowi = rowi; c.ldi = ldi; c.valid = valid; c.rowj = rowj; c.lda = p.lda; c.kb = kb;
        c.i = i; c.j = j; c.k0 = k0; c.k1a = k1a; c.k1 = k1; c.part = part;
        c.pre = false; c.head = true; c.multi = multi;
        c.ready_i = ready_i; c.ready_j = ready_j; c.ready_ij = ready_b + (int64_t)i * p.nt + j; c.pcount = pcount;
        c.pfl = p.pflag + ((int64_t)b * p.nt + j) * 4;
        c.Linv = wsb + (int64_t)j * TILE * TILE; c.refine_flag = flags + j;
        c.pre_count = p.pcount + ((int64_t)b * rows_total + i) * p.nt + i;
        c.pre_parts = (i >= 2) ? df_split(sh, i - 2) : 0;  // PRE(i) is ticketed in column group i - 2
        c.scratch = scratch;
        head_chain<MULTI>(smem_raw, c, pe);
        DFP_ADD(13, k1 - k1a);
        DFP_MARK(5);
        if (pf) pf[2] = pf[3] = pf[4] = globaltimer_ns();
      } else {
        Acc acc;
        wait_ready(ready_j + j, multi);
        DFP_MARK(4);
        const bool refine = __ldcg(flags + j) != 0.0;
        tile_solve(stages, T, ldi, valid, kb, wsb + (int64_t)j * TILE * TILE, rowj + (int64_t)j * TILE, p.lda, refine,
                   scratch, acc);
        DFP_MARK(5);
        // multi-GPU: L_ij goes into the peers' copies of the matrix by NVLink stores straight from the
        // accumulators -- first to the rank that owns the next tile row (+ flags), then to the others
        if (multi) store_tile<0, true>(peer_ptr(T, p.peers.delta[(p.peers.rank + 1) % p.peers.world]), ldi, valid, kb, acc, false);
        fence_publish(pe);
        __syncthreads();
        if (tid == 0) publish_flag(ready_b + (int64_t)i * p.nt + j, pe, PEERS_FIRST);
        if (multi && p.peers.world > 2) {
          for (int pr = 0; pr < p.peers.world; ++pr)
            if (peer_in_set(pe, pr, PEERS_REST)) store_tile<0, true>(peer_ptr(T, p.peers.delta[pr]), ldi, valid, kb, acc, false);
          __threadfence_system();
          __syncthreads();
          if (tid == 0) publish_flag(ready_b + (int64_t)i * p.nt + j, pe, PEERS_REST);
        }
        DFP_MARK(6);
      }
      if (head) {
        const int k = i;
        double* Tkk = rowi + (int64_t)k * TILE;
        DFP_MARK(9);
        if (pf) pf[5] = globaltimer_ns();
        diag_factor_core<MULTI>(smem_raw, Tkk, p.lda, valid, (int64_t)k * TILE, wsb + (int64_t)k * TILE * TILE, flags + k,
                         p.info + b, ready_b + (int64_t)k * p.nt + k, pe, nullptr,
                         p.pflag + ((int64_t)b * p.nt + k) * 4);
        DFP_MARK(10);
#ifdef GPAR_DF_PROF
        if (threadIdx.x == 0 && p.prof && b == 0 && k < 1024) p.prof[64 + 16 * 256 + k] = globaltimer_ns();
#endif
        if (pf) pf[6] = globaltimer_ns();
        if (multi && p.peers.world > 2) {  // L_{k,k-1} for the remaining peers, re-read from the local copy
          for (int idx = tid; idx < TILE * (TILE / 2); idx += GEMM_THREADS) {
            const int r = idx >> 6, c = (idx & 63) * 2;
            if (r >= valid) continue;
            const double2 v = __ldcg(reinterpret_cast<const double2*>(T + (int64_t)r * ldi + c));
            for (int pr = 0; pr < p.peers.world; ++pr)
              if (peer_in_set(pe, pr, PEERS_REST))
                *reinterpret_cast<double2*>(peer_ptr(T, p.peers.delta[pr]) + (int64_t)r * ldi + c) = v;
          }
          __threadfence_system();
          __syncthreads();
          if (tid == 0) publish_flag(ready_b + (int64_t)i * p.nt + j, pe, PEERS_REST);
        }
      }
    }
  }
#ifdef GPAR_DF_PROF
  if (threadIdx.x == 0 && p.prof) {
    s_dfp[2] = *dfp_wait_ptr();
    s_dfp[14] = globaltimer_ns();
    for (int q = 0; q < 16; ++q) p.prof[64 + 16 * (long long)blockIdx.x + q] = s_dfp[q];
  }
#endif
}

// --------------------------------------------------------------------------------------
// U = L^-T as a tile dataflow (gpar_potri): tile (r, j), j >= r, of the upper-triangular U is
//     U_rj = (delta_rj I - sum_{k=r}^{j-1} U_rk L_jk^T) L_jj^-T,
// i.e. the sweep of trtri_rows_kernel with every tile a task of its own: a persistent grid pulls tickets
// (column-major: all tiles of column j before column j + 1, longest K first), the K-loop of a tile
// consumes the tiles U_rk as their ready flags come up, and the chain of a row block shrinks from
// sum_j (K-loop of j tiles) -- 11.5 ms of 15.9 ms at n = 7424 with one CTA per 32-row block -- to one
// k-tile + one tile solve per column.  Every dependency has a smaller ticket: deadlock-free.
// --------------------------------------------------------------------------------------
struct TrtriArgs {
  const double* L; int64_t ldl; int64_t n; const double* ws;
  double* U; int64_t ldu;
  double* pool;            // gridDim.x scratch tiles (refined solves)
  int* ticket; int* ready; // ready[r * nt + j]
  int nt; int total;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) trtri_dataflow_kernel(const TrtriArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_task;
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem_raw);
  const int tid = threadIdx.x;
  double* scratch = p.pool + (int64_t)blockIdx.x * TILE * TILE;
  const double* flags = p.ws + (int64_t)p.nt * TILE * TILE;
  pipe_init();
  for (;;) {
    if (tid == 0) s_task = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= p.total) break;
    int j = static_cast<int>((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((j + 1) * (j + 2) / 2 <= t) ++j;
    while (j * (j + 1) / 2 > t) --j;
    const int r = t - j * (j + 1) / 2;
    const int valid = static_cast<int>(min64(TILE, p.n - (int64_t)r * TILE));
    const int kb = static_cast<int>(min64(TILE, p.n - (int64_t)j * TILE));
    double* Urow = p.U + (int64_t)r * TILE * p.ldu;
    double* T = Urow + (int64_t)j * TILE;  // holds delta_rj I (set_identity_kernel)
    const double* Lrow = p.L + (int64_t)j * TILE * p.ldl;
    if (j > r) {
      Acc acc;
      acc_zero(acc);
      const int* rdy = p.ready + (int64_t)r * p.nt + r;
      gemm_nt_mainloop_dep<0>(stages, Urow + (int64_t)r * TILE, p.ldu, valid, Lrow + (int64_t)r * TILE, p.ldl, kb,
                              (j - r) * TILE, acc, rdy, rdy, false, T, p.ldu);
      store_tile<1>(T, p.ldu, valid, kb, acc, false);
      __threadfence();
      __syncthreads();
    }
    Acc xacc;
    tile_solve(stages, T, p.ldu, valid, kb, p.ws + (int64_t)j * TILE * TILE, Lrow + (int64_t)j * TILE, p.ldl,
               __ldcg(flags + j) != 0.0, scratch, xacc);
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(p.ready + (int64_t)r * p.nt + j, 1);
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
  cudaFuncSetAttribute(trtri_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
  cudaFuncSetAttribute(trtri_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
  cudaFuncSetAttribute(potrf_dataflow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM_BYTES);
  cudaFuncSetAttribute(potrf_dataflow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM_BYTES);
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
static int64_t ws_ready_ints(int64_t nt, int64_t nbt, int64_t batch) {
  return 2 + 2 * batch * (nt + nbt) * nt + 4 * batch * nt;  // ticket, tile ready flags, K-part counters, panel flags
}

extern "C" size_t gpar_potrf_workspace_bytes(int64_t n, int64_t nb, int64_t batch) {
  if (n <= 0 || batch <= 0) return 0;
  const int64_t nt = (n + TILE - 1) / TILE, nbt = nb > 0 ? (nb + TILE - 1) / TILE : 0;
  const int64_t doubles = batch * ws_stride(nt, nbt) + (int64_t)DF_POOL_TILES * TILE * TILE +
                          (ws_ready_ints(nt, nbt, batch) + 1) / 2 + 2;
  return (size_t)doubles * sizeof(double);
}

static int device_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

extern "C" size_t gpar_trsm_rows_scratch_bytes(int64_t nb) {
  if (nb <= 0) return 0;
  const RowPlan p = trsm_row_plan(nb, device_sms());
  return (size_t)(p.full_blocks + p.tail_blocks) * TILE * TILE * sizeof(double);  // one tile per row block
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Launch of the persistent dataflow kernel (single GPU: peers.world == 1).  reset != 0 zeroes the
// ticket / ready flags first (the multi-GPU driver does that in a separate call, followed by a
// barrier over the ranks, because peers write flags into this workspace).
static int df_reset(double* ws, int64_t n, int64_t nb, int64_t batch, cudaStream_t stream) {
  const int64_t nt = (n + TILE - 1) / TILE, nbt = nb > 0 ? (nb + TILE - 1) / TILE : 0;
  double* pool = ws + batch * ws_stride(nt, nbt);
  int* ints = reinterpret_cast<int*>(pool + (int64_t)DF_POOL_TILES * TILE * TILE);
  cudaMemsetAsync(ints, 0, sizeof(int) * (size_t)ws_ready_ints(nt, nbt, batch), stream);
  return 0;
}

static int launch_dataflow(double* A, int64_t lda, int64_t n, int64_t strideA, double* B, int64_t ldb, int64_t nb,
                           int64_t strideB, int64_t batch, double* ws, int32_t* info, const Peers& peers, bool reset,
                           cudaStream_t stream) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int nt = (int)((n + TILE - 1) / TILE);
  const int nbt = nb > 0 ? (int)((nb + TILE - 1) / TILE) : 0;
  DfArgs p;
  p.A = A; p.lda = lda; p.n = n; p.strideA = strideA;
  p.B = B; p.ldb = ldb; p.nb = nb; p.strideB = strideB;
  p.batch = (int)batch; p.nt = nt; p.nbt = nbt;
  int grid = num_sms < DF_POOL_TILES ? num_sms : DF_POOL_TILES;
  const DfShape sh = {nt, nbt, (int)batch, grid, peers.world};
  const long long total = df_total_tasks(sh);
  if (total > 0x7fffffff) { set_error("gpar_potrf: too many tile tasks"); return -9; }
  p.total_tasks = (int)total;
  p.grid = grid;
  p.ws = ws; p.strideWs = ws_stride(nt, nbt);
  p.pool = ws + batch * p.strideWs;
  p.info = info;
  p.prof = g_df_prof;
  p.peers = peers;
  int* ints = reinterpret_cast<int*>(p.pool + (int64_t)DF_POOL_TILES * TILE * TILE);
  p.ticket = ints; p.ready = ints + 2;
  p.pcount = p.ready + batch * (int64_t)(nt + nbt) * nt;
  p.pflag = p.pcount + batch * (int64_t)(nt + nbt) * nt;
  if (reset) df_reset(ws, n, nb, batch, stream);
  if ((long long)grid > total) grid = (int)total;  // (p.grid keeps the nominal size: it only feeds the split rule)
  if (peers.world > 1)
    potrf_dataflow_kernel<true><<<grid, GEMM_THREADS, DF_SMEM_BYTES, stream>>>(p);
  else
    potrf_dataflow_kernel<false><<<grid, GEMM_THREADS, DF_SMEM_BYTES, stream>>>(p);
  return check_launch("gpar_potrf");
}

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
    Peers solo;
    solo.rank = 0; solo.world = 1; solo.row_block = 1;
    for (int r = 0; r < GPAR_MAX_PEERS; ++r) solo.delta[r] = 0;
    return launch_dataflow(A, lda, n, strideA, B, ldb, nb, strideB, batch, ws, info, solo, true, stream);
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
      p.K = kb; p.lower = 1; p.nt_rows1 = ntr; p.add = 0; p.tri_k = 0;
      p.C2 = nb > 0 ? B + (j0 + TILE) : nullptr; p.ldc2 = ldb; p.c_rows2 = nb; p.strideC2 = strideB;
      p.Aop2 = nb > 0 ? B + j0 : nullptr; p.lda2 = ldb; p.strideA2 = strideB;
      gemm_sub_kernel<<<dim3((unsigned)ntr, (unsigned)(ntr + nbt), (unsigned)batch), GEMM_THREADS, GEMM_SMEM_BYTES,
                        stream>>>(p);
    }
  }
  return check_launch("gpar_potrf");
}

// ---- multi-GPU Cholesky: one process per GPU, tile rows dealt round-robin ---------
extern "C" int gpar_potrf_multi_reset(double* ws, int64_t n, int64_t nb, int32_t* info, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ws || !aligned16(ws)) { set_error("gpar_potrf_multi_reset: bad workspace"); return -1; }
  if (!info) return -4;
  cudaMemsetAsync(info, 0, sizeof(int32_t), stream);
  if (n <= 0) return 0;
  df_reset(ws, n, nb, 1, stream);
  return check_launch("gpar_potrf_multi_reset");
}

extern "C" int gpar_potrf_multi(double* A, int64_t lda, int64_t n, double* B, int64_t ldb, int64_t nb, double* ws,
                                int32_t* info, int rank, int world, const int64_t* peer_delta_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!A || !aligned16(A)) { set_error("gpar_potrf_multi: A null or not 16-byte aligned"); return -1; }
  if (lda < n || (lda & 1)) { set_error("gpar_potrf_multi: lda must be even and >= n"); return -2; }
  if (n < 0) return -3;
  if (nb > 0 && (!B || !aligned16(B) || ldb < n || (ldb & 1))) { set_error("gpar_potrf_multi: bad appended rows"); return -4; }
  if (!ws || !aligned16(ws)) { set_error("gpar_potrf_multi: bad workspace"); return -7; }
  if (!info) return -8;
  if (world < 1 || world > GPAR_MAX_PEERS || rank < 0 || rank >= world) { set_error("gpar_potrf_multi: bad rank/world"); return -9; }
  if (world > 1 && !peer_delta_bytes) return -11;
  set_smem_attrs();
  if (n == 0) return 0;
  Peers pe;
  pe.rank = rank; pe.world = world; pe.row_block = GPAR_ROW_BLOCK;
  for (int r = 0; r < GPAR_MAX_PEERS; ++r) pe.delta[r] = (r < world && r != rank) ? (long long)peer_delta_bytes[r] : 0;
  for (int r = 0; r < world; ++r)
    if (pe.delta[r] & 15) { set_error("gpar_potrf_multi: peer deltas must be multiples of 16 bytes"); return -11; }
  return launch_dataflow(A, lda, n, 0, B, ldb, nb, 0, 1, ws, info, pe, false, stream);
}

// Peer-mappable device memory (CUDA IPC): every rank allocates the same size, exports its handle,
// and opens its peers' handles; `gpar_potrf_multi` is handed the address differences.
extern "C" int gpar_ipc_alloc(size_t bytes, void** ptr) {
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) { set_error("gpar_ipc_alloc: %s", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
  return 0;
}
extern "C" int gpar_ipc_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? 0 : -1; }
extern "C" int gpar_ipc_export(void* ptr, unsigned char* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { set_error("gpar_ipc_export: %s", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int gpar_ipc_open(const unsigned char* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("gpar_ipc_open: %s", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
  return 0;
}
extern "C" int gpar_ipc_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? 0 : -1; }

extern "C" int gpar_trsm_rows(const double* L, int64_t ldl, int64_t n, const double* ws, double* B, int64_t ldb,
                              int64_t nb, double* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!L || !aligned16(L) || (ldl & 1) || ldl < n) { set_error("gpar_trsm_rows: bad L"); return -1; }
  if (!ws || !aligned16(ws)) return -4;
  if (nb <= 0 || n <= 0) return 0;
  if (!B || !aligned16(B) || (ldb & 1) || ldb < n) { set_error("gpar_trsm_rows: bad B"); return -5; }
  if (!scratch || !aligned16(scratch)) { set_error("gpar_trsm_rows: bad scratch"); return -8; }
  set_smem_attrs();
  const RowPlan p = trsm_row_plan(nb, device_sms());
  unsigned g = (unsigned)(p.full_blocks + p.tail_blocks);
  trsm_rows_kernel<<<g, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(L, ldl, n, ws, B, ldb, nb, scratch, p.full_blocks,
                                                                  p.tail_h);
  return check_launch("gpar_trsm_rows");
}

static int syrk_impl(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw, int64_t k,
                     int64_t strideW, int64_t batch, int add, void* stream_, int tri_k = 0);

extern "C" int gpar_syrk_sub(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw,
                             int64_t k, int64_t strideW, int64_t batch, void* stream_) {
  return syrk_impl(C, ldc, n, strideC, W, ldw, k, strideW, batch, 0, stream_);
}

extern "C" int gpar_syrk_add(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw,
                             int64_t k, int64_t strideW, int64_t batch, void* stream_) {
  return syrk_impl(C, ldc, n, strideC, W, ldw, k, strideW, batch, 1, stream_);
}

static int syrk_impl(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw, int64_t k,
                     int64_t strideW, int64_t batch, int add, void* stream_, int tri_k) {
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
  p.K = (int)k; p.lower = 1; p.nt_rows1 = (int)nt; p.add = add; p.tri_k = tri_k;
  p.C2 = nullptr; p.ldc2 = 0; p.c_rows2 = 0; p.strideC2 = 0; p.Aop2 = nullptr; p.lda2 = 0; p.strideA2 = 0;
  gemm_sub_kernel<<<dim3(nt, nt, (unsigned)batch), GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
  return check_launch("gpar_syrk_sub");
}

extern "C" int gpar_gemm_nt(double* C, int64_t ldc, int64_t m, int64_t n, const double* A, int64_t lda,
                            const double* B, int64_t ldb, int64_t k, int add, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  if (!C || !aligned16(C) || (ldc & 1) || ldc < n) { set_error("gpar_gemm_nt: bad C"); return -1; }
  if (!A || !aligned16(A) || (lda & 1) || lda < k) { set_error("gpar_gemm_nt: bad A"); return -5; }
  if (!B || !aligned16(B) || (ldb & 1) || ldb < k) { set_error("gpar_gemm_nt: bad B"); return -7; }
  const int64_t mt = (m + TILE - 1) / TILE, ntc = (n + TILE - 1) / TILE;
  if (mt > 65535) { set_error("gpar_gemm_nt: more than 65535 row tiles"); return -3; }
  set_smem_attrs();
  SubArgs p;
  p.C = C; p.ldc = ldc; p.c_rows = m; p.c_cols = n; p.strideC = 0;
  p.Aop = A; p.lda = lda; p.strideA = 0;
  p.Bop = B; p.ldb = ldb; p.strideB = 0;
  p.K = (int)k; p.lower = 0; p.nt_rows1 = (int)mt; p.add = add; p.tri_k = 0;
  p.C2 = nullptr; p.ldc2 = 0; p.c_rows2 = 0; p.strideC2 = 0; p.Aop2 = nullptr; p.lda2 = 0; p.strideA2 = 0;
  gemm_sub_kernel<<<dim3((unsigned)ntc, (unsigned)mt, 1), GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
  return check_launch("gpar_gemm_nt");
}

// K10 -- A^-1 from the factor (gradients of the log-marginal): U <- L^-T, Ainv (lower) <- U U^T.
// scratch of gpar_potri: [DF_POOL_TILES scratch tiles][ticket (2 ints) + nt * nt ready flags]
extern "C" size_t gpar_potri_scratch_bytes(int64_t n) {
  if (n <= 0) return 0;
  const int64_t nt = (n + TILE - 1) / TILE;
  return (size_t)((int64_t)DF_POOL_TILES * TILE * TILE + (2 + nt * nt + 1) / 2 + 2) * sizeof(double);
}

extern "C" int gpar_potri(const double* L, int64_t ldl, int64_t n, const double* ws, double* U, int64_t ldu,
                          double* Ainv, int64_t lda, double* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!L || !aligned16(L) || (ldl & 1) || ldl < n) { set_error("gpar_potri: bad L"); return -1; }
  if (!ws || !aligned16(ws)) return -4;
  if (n <= 0) return 0;
  if (!U || !aligned16(U) || (ldu & 1) || ldu < n) { set_error("gpar_potri: bad U"); return -5; }
  if (!Ainv || !aligned16(Ainv) || (lda & 1) || lda < n) { set_error("gpar_potri: bad Ainv"); return -7; }
  if (!scratch || !aligned16(scratch)) { set_error("gpar_potri: bad scratch"); return -9; }
  set_smem_attrs();
  const int64_t total = n * ldu;
  set_identity_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(U, ldu, n);
  const unsigned nt = (unsigned)((n + TILE - 1) / TILE);
  static const bool use_rows = (getenv("GPAR_TRTRI_ROWS") != nullptr);  // round-1 sweep, kept as the measured baseline
  if (use_rows) {
    if ((size_t)nt * (TILE / TRTRI_ROWS) > (size_t)DF_POOL_TILES) { set_error("gpar_potri: GPAR_TRTRI_ROWS needs n <= %d", DF_POOL_TILES / 4 * TILE); return -9; }
    trtri_rows_kernel<<<nt * (TILE / TRTRI_ROWS), GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(L, ldl, n, ws, U, ldu, scratch);
  } else {
    static int num_sms = 0;
    if (num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    TrtriArgs p;
    p.L = L; p.ldl = ldl; p.n = n; p.ws = ws; p.U = U; p.ldu = ldu;
    p.pool = scratch;
    int* ints = reinterpret_cast<int*>(scratch + (int64_t)DF_POOL_TILES * TILE * TILE);
    p.ticket = ints; p.ready = ints + 2;
    p.nt = (int)nt; p.total = (int)(nt * (nt + 1) / 2);
    cudaMemsetAsync(ints, 0, sizeof(int) * (size_t)(2 + (size_t)nt * nt), stream);
    int grid = num_sms < DF_POOL_TILES ? num_sms : DF_POOL_TILES;
    if (grid > p.total) grid = p.total;
    trtri_dataflow_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
  }
  cudaMemsetAsync(Ainv, 0, sizeof(double) * (size_t)n * lda, stream);
  int rc = syrk_impl(Ainv, lda, n, 0, U, ldu, n, 0, 1, 1, stream_, 1);
  if (rc) return rc;
  return check_launch("gpar_potri");
}

// Debug / tests: the task list of the dataflow kernel, decoded on the host by the same function the
// kernel uses.  out6 = {kind (0 D0, 1 HEAD, 2 PLAIN, 3 PRE), matrix, tile row, tile column, K-part, parts};
// returns the total number of tickets of a launch with `grid` CTAs (t < 0: only the count).
extern "C" int gpar_debug_decode_ticket(int64_t n, int64_t nb, int64_t batch, int64_t grid, int64_t t, int32_t* out6) {
  const int nt = (int)((n + TILE - 1) / TILE), nbt = nb > 0 ? (int)((nb + TILE - 1) / TILE) : 0;
  const DfShape sh = {nt, nbt, (int)batch, (int)grid, 1};
  const long long total = df_total_tasks(sh);
  if (t >= 0 && t < total && out6) {
    int kind, b, i, j, part, nparts;
    df_decode((int)t, sh, kind, b, i, j, part, nparts);
    out6[0] = kind; out6[1] = b; out6[2] = i; out6[3] = j; out6[4] = part; out6[5] = nparts;
  }
  return (int)total;
}

// Debug: the row-block plan of gpar_trsm_rows for nb rows on `sms` SMs (host only).
extern "C" int gpar_debug_trsm_row_plan(int64_t nb, int sms, int64_t* out3) {
  if (nb <= 0 || sms <= 0 || !out3) return -1;
  const RowPlan p = trsm_row_plan(nb, sms);
  out3[0] = p.full_blocks; out3[1] = p.tail_blocks; out3[2] = p.tail_h;
  return 0;
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
