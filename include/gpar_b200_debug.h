/* gpar_b200_debug.h -- diagnostics of the gpar_b200 kernels.  NOT part of the drop-in C ABI (gpar_b200.h).
 *
 * Two groups:
 *  - libgpar_b200_debug.so (csrc/debug.cu): raw fp64 issue-rate and latency probes.  bench.py uses
 *    gpar_fp64_probe to state the DMMA / DFMA issue rates next to the cuBLAS DGEMM roofline denominator.
 *  - hooks into the Cholesky dataflow kernel, exported by libgpar_b200.so itself because they need its
 *    internals: the task-list decoder (used by the CPU tests to check coverage / dependency order of the ticket
 *    list with the kernel's own function) and two profiling hooks used by scripts/prof_*.py.
 */
#ifndef GPAR_B200_DEBUG_H
#define GPAR_B200_DEBUG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- libgpar_b200_debug.so ---------------------------------------------------------------- */
/* Diagnostics: raw fp64 issue-rate probes used by bench.py to state the roofline
 * denominators next to cuBLAS DGEMM.  mode 0 = DMMA m8n8k4, 1 = DFMA.  Returns the
 * number of flops executed per launch through *flops. */
int gpar_fp64_probe(int mode, int64_t iters, double* sink, double* flops, void* stream);

/* Debug: single-warp dependent-chain latencies in cycles (out: >= 32 doubles). */
int gpar_debug_latency_probe(double* out, void* stream);

/* ---- hooks exported by libgpar_b200.so ---------------------------------------------------- */
/* Debug: when non-null, gpar_potrf records globaltimer stamps (24 values) of the tile tasks
 * around column nt/2 of matrix 0 into prof. */
int gpar_debug_set_dataflow_prof(long long* prof);
/* Debug / tests: ticket t of the dataflow kernel's task list for an n x n matrix with nb appended rows, decoded on
 * the host by the function the kernel uses, for a launch of `grid` CTAs: out6 = {kind (0 first diagonal tile,
 * 1 head = sub-diagonal solve + diagonal factor of tile row out6[2], 2 plain tile, 3 diagonal pre-update), matrix,
 * tile row, tile column, K-part, number of K-parts of the tile (split-K in the tail of the sweep)}.
 * Returns the number of tickets. */
int gpar_debug_decode_ticket(int64_t n, int64_t nb, int64_t batch, int64_t grid, int64_t t, int32_t* out6);
/* Debug: row-block plan of gpar_trsm_rows for nb rows on `sms` SMs: out3 = {128-row blocks, tail blocks, tail
 * block height}.  Host only. */
int gpar_debug_trsm_row_plan(int64_t nb, int sms, int64_t* out3);
/* Debug: clock64 phase timestamps (21 values) of the diagonal-tile factor on A[0:128, 0:128]. */
int gpar_debug_diag_profile(double* A, int64_t lda, int64_t n, double* ws, int32_t* info, long long* prof,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif
