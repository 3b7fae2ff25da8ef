/*
 * gpar_b200 -- C ABI of the B200-native (sm_100a) per-layer GP hot path of GPAR.
 *
 * The reference (wesselb/gpar) has no FFI: its hot path is the stheno object
 * protocol driven from gpar/model.py.  Each entry point
 * below names the protocol element / call site it replaces.  The caller is the
 * host p-loop in gpar_b200/model.py (ctypes); every pointer is caller-owned
 * DEVICE memory (torch CUDA tensors), `stream` is a cudaStream_t passed as
 * void*.  Nothing is retained after a call returns; work is ordered on
 * `stream`.  All matrices are row-major fp64 with a leading dimension in
 * ELEMENTS; symmetric / triangular matrices keep the LOWER triangle
 * authoritative (entries above the diagonal are never read and, unless stated,
 * never written).  Base pointers must be 16-byte aligned and leading
 * dimensions even (the tile loaders use 16-byte cp.async).
 *
 * Return value: 0 = ok, <0 = -(index of the offending argument, 1-based) or a
 * CUDA launch failure (see gpar_last_error()).  Numerical failure (first
 * non-positive pivot, LAPACK `info` style, 1-based) is written to the caller's
 * device `info` word and read back only when the caller chooses to sync.
 */
#ifndef GPAR_B200_H
#define GPAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPAR_ABI_VERSION 1
#define GPAR_TILE 128        /* Cholesky tile edge; workspace is sized in tiles */
#define GPAR_MAX_TERMS 8
#define GPAR_MAX_FEATS 96
#define GPAR_MAX_PEERS 8       /* ranks of one NVLink domain that gpar_potrf_multi spans */
#define GPAR_ROW_BLOCK 4       /* gpar_potrf_multi deals tile rows to the ranks in blocks of 4 (block-cyclic) */

/* Kernel terms after lowering to a feature map (see gpar_kernel_spec_t). */
enum { GPAR_TERM_EQ = 0, GPAR_TERM_RQ = 1, GPAR_TERM_LINEAR = 2, GPAR_TERM_CONST = 3 };
/* Feature ops: phi = a * x[col], a * sin(b * x[col]), a * cos(b * x[col]). */
enum { GPAR_FEAT_SCALE = 0, GPAR_FEAT_SIN = 1, GPAR_FEAT_COS = 2 };

typedef struct {
  int32_t type;     /* GPAR_TERM_* */
  int32_t f_begin;  /* features [f_begin, f_end) belong to this term */
  int32_t f_end;
  int32_t _pad;
  double variance;  /* multiplier; for CONST the constant itself */
  double alpha;     /* RQ shape */
} gpar_term_t;

/*
 * Closed kernel family of gpar/regression.py:92-180 (`_model_generator`) in
 * feature-map normal form: k(x, y) = sum_t k_t(phi_t(x), phi_t(y)) with
 *   EQ:     var * exp(-1/2 ||phi(x) - phi(y)||^2)
 *   RQ:     var * (1 + ||phi(x) - phi(y)||^2 / (2 alpha))^(-alpha)
 *   LINEAR: var * <phi(x), phi(y)>
 *   CONST:  var
 * `.stretch(s)` becomes a = 1/s; `.select(cols)` becomes feat_col;
 * `EQ().stretch(s).periodic(T) * EQ().stretch(d)` becomes one EQ term over the
 * 3m features [sin(2 pi x/T)/s_c, cos(2 pi x/T)/s_{m+c}, x/d_c].
 */
typedef struct {
  int32_t n_terms;
  int32_t n_feats;
  gpar_term_t terms[GPAR_MAX_TERMS];
  int32_t feat_col[GPAR_MAX_FEATS];
  int32_t feat_op[GPAR_MAX_FEATS];
  double feat_a[GPAR_MAX_FEATS];
  double feat_b[GPAR_MAX_FEATS];
} gpar_kernel_spec_t;

int gpar_abi_version(void);
const char* gpar_last_error(void);

/* K1 -- Gram matrix.  Replaces mlkernels kernel evaluation `k(x, y)` driven from
 * regression.py:94-179 and the `+ noise / w` of `f(x, noise / w)` (model.py:287-289).
 * Y == NULL: symmetric K(X, X); then diag_add (length nx, may be NULL) and eps are
 * added on the diagonal (K_ii += diag_add[i] + eps) and, if lower_only != 0, tiles
 * strictly above the diagonal are skipped.  out is nx x ny. */
int gpar_gram(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t nx,
              const double* Y, int64_t ldy, int64_t ny, const double* diag_add, double eps,
              int lower_only, double* out, int64_t ldo, void* stream);

/* Batched form: matrix b uses X + b*strideX, Y + b*strideY, diag_add + b*strideD, out + b*strideO
 * (strides in elements; a stride of 0 shares the operand across the batch). */
int gpar_gram_batched(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t nx, int64_t strideX,
                      const double* Y, int64_t ldy, int64_t ny, int64_t strideY, const double* diag_add,
                      int64_t strideD, double eps, int lower_only, double* out, int64_t ldo, int64_t strideO,
                      int64_t batch, void* stream);

/* K2 -- blocked Cholesky (matrix.cholesky -> torch.linalg.cholesky in the reference),
 * batched, with optional appended row blocks.  For each b < batch:
 *   A_b (n x n, lower) <- L_b with L_b L_b^T = A_b;   B_b (nb x n) <- B_b L_b^-T.
 * B may be NULL (nb = 0).  `ws` receives the inverses of the diagonal GPAR_TILE
 * blocks of L, a per-block conditioning flag and scratch tiles
 * (gpar_potrf_workspace_bytes(n, nb, batch) bytes, 16-byte aligned); keep it for
 * gpar_trsm_rows / gpar_backsolve.  Solves against a diagonal block with
 * kappa_inf(L_kk) > 1e3 get one step of iterative refinement (backward stable for
 * near-singular covariances).  The strictly upper part of each diagonal 128-tile of A
 * is zeroed.  info[b] = 0 or 1-based index of the first non-positive pivot. */
size_t gpar_potrf_workspace_bytes(int64_t n, int64_t nb, int64_t batch);
int gpar_potrf(double* A, int64_t lda, int64_t n, int64_t strideA, double* B, int64_t ldb, int64_t nb,
               int64_t strideB, int64_t batch, double* ws, int32_t* info, void* stream);

/* K2, multi-GPU (BASELINE config 5): the same factorisation spread over `world` GPUs of one
 * NVLink domain, one process per GPU.  Every rank holds a full-size copy of A (n x n, lower: all ranks build the
 * same Gram matrix), of B and of the workspace, at IDENTICAL offsets inside one peer-mapped allocation
 * (gpar_ipc_* below, or any other means of mapping peer memory); peer_delta_bytes[r] is the address of rank r's
 *  allocation minus the address of this rank's (ignored for r == rank).  Tile rows are dealt block-cyclically
 * (GPAR_ROW_BLOCK consecutive tile rows per turn, so that the critical chain -- sub-diagonal solve, diagonal
 * update and factor of consecutive tile rows -- crosses NVLink only once per block): rank r factors / solves the
 * tiles of its rows and pushes every finished tile, the inverses
 * of the diagonal tiles and the ready flags into all peers' copies by plain NVLink stores from inside the
 * kernel, so operand streaming stays local and the transfer overlaps the math tile by tile.  On return (after
 * the caller has synchronised the stream AND passed a barrier over the ranks) every rank holds the complete
 * L, B L^-T and workspace.  Protocol:  gpar_potrf_multi_reset on every rank -> stream sync + barrier over the
 * ranks -> gpar_potrf_multi on every rank -> stream sync + barrier.  info: as gpar_potrf, set on the rank that
 * owns the offending diagonal tile (reduce with max over ranks). */
int gpar_potrf_multi_reset(double* ws, int64_t n, int64_t nb, int32_t* info, void* stream);
int gpar_potrf_multi(double* A, int64_t lda, int64_t n, double* B, int64_t ldb, int64_t nb, double* ws,
                     int32_t* info, int rank, int world, const int64_t* peer_delta_bytes, void* stream);

/* Peer-mappable device memory over CUDA IPC: allocate, export a 64-byte handle, open a peer's handle. */
int gpar_ipc_alloc(size_t bytes, void** ptr);
int gpar_ipc_free(void* ptr);
int gpar_ipc_export(void* ptr, unsigned char* handle64);
int gpar_ipc_open(const unsigned char* handle64, void** ptr);
int gpar_ipc_close(void* ptr);

/* K4 -- B (nb x n) <- B L^-T given L and the `ws` of its gpar_potrf.  This is
 * (L^-1 K(x_a, x_))^T of PosteriorMean / PosteriorKernel.
 * `scratch`: gpar_trsm_rows_scratch_bytes(nb) bytes on the CURRENT device (one tile per row block; the
 * number of row blocks depends on the SM count: whole waves of 128-row blocks, then blocks of 32 / 64 / 96
 * rows).  Rows are solved independently: the result does not depend on how they fall into blocks. */
size_t gpar_trsm_rows_scratch_bytes(int64_t nb);
int gpar_trsm_rows(const double* L, int64_t ldl, int64_t n, const double* ws, double* B, int64_t ldb,
                   int64_t nb, double* scratch, void* stream);

/* K5 -- C_b (n x n, lower) <- C_b - W_b W_b^T, W_b is n x k.  Posterior covariance
 * K** - V^T V. */
int gpar_syrk_sub(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw, int64_t k,
                  int64_t strideW, int64_t batch, void* stream);

/* C_b <- C_b + W_b W_b^T (lower): the `I + B Sigma^-1 B^T` and `+ B_x^T A^-1 B_x'` terms of the VFE
 * path. */
int gpar_syrk_add(double* C, int64_t ldc, int64_t n, int64_t strideC, const double* W, int64_t ldw, int64_t k,
                  int64_t strideW, int64_t batch, void* stream);

/* C (m x n) <- C -/+ A B^T with A (m x k), B (n x k), all row major (add != 0: plus).  Rectangular
 * counterpart of gpar_syrk_sub / gpar_syrk_add: the cross block k~(x, z) = k(x, z) - B_x B_z^T + D_x D_z^T of a
 * sparse posterior's kernel (logpdf under a conditioned model, regression.py:495-499). */
int gpar_gemm_nt(double* C, int64_t ldc, int64_t m, int64_t n, const double* A, int64_t lda, const double* B,
                 int64_t ldb, int64_t k, int add, void* stream);

/* K10 -- gradients of the dense log-marginal for `fit` (regression.py:434-459; the reference differentiates
 * through Gram + Cholesky + solve with torch autograd):  d LML / d theta = sum_ij W_ij dA_ij / d theta,
 * W = 1/2 (alpha alpha^T - A^-1).
 * gpar_potri: U (n x n work matrix) <- L^-T, Ainv (n x n, lower) <- U U^T = A^-1, given L and the `ws` of its
 * gpar_potrf; scratch = gpar_potri_scratch_bytes(n).
 * gpar_gram_grad: one pass over the lower triangle of Ainv accumulates the raw sums of the chain rule for every
 * term and feature of `spec` into out[GPAR_GRAD_NP] (layout: out[2t] variance of term t, out[2t+1] RQ alpha,
 * out[16+2f] / out[16+2f+1] the a- / b-parameter of feature f, out[208] the diagonal term sum_i W_ii dvec_i;
 * exact definitions at gram_grad_kernel in csrc/gram.cu, chain rule in gpar_b200/spec.py).  Partial sums are
 * combined in a fixed order (bitwise reproducible).  workspace = gpar_gram_grad_workspace_bytes(n). */
#define GPAR_GRAD_NP (2 * GPAR_MAX_TERMS + 2 * GPAR_MAX_FEATS + 1)
size_t gpar_potri_scratch_bytes(int64_t n);
int gpar_potri(const double* L, int64_t ldl, int64_t n, const double* ws, double* U, int64_t ldu, double* Ainv,
               int64_t lda, double* scratch, void* stream);
/* gpar_gram_wgrad: the same raw sums for a rectangular block k(x_i, y_j) (i < nx, j < ny) under explicit weights
 * W_ij = ux_i uy_j + sx_i G_ij (G nx x ny row major with leading dimension ldg; G, sx, ux/uy optional; every pair
 * counted once; out[208] is not touched by it = 0) -- the sum G_zx dK_zx and sum G_zz dK_zz terms of the
 * gradient of the VFE bound (inducing-point layers of `fit`).  gpar_row_sqnorm: out[i] = sum_c A[i][c]^2. */
size_t gpar_gram_wgrad_workspace_bytes(int64_t nx, int64_t ny);
int gpar_gram_wgrad(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t nx, const double* Y,
                    int64_t ldy, int64_t ny, const double* G, int64_t ldg, const double* sx, const double* ux,
                    const double* uy, double* workspace, double* out, void* stream);
int gpar_row_sqnorm(const double* A, int64_t lda, int64_t n, int64_t k, double* out, void* stream);
size_t gpar_gram_grad_workspace_bytes(int64_t n);
int gpar_gram_grad(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t n, const double* alpha,
                   const double* Ainv, int64_t lda, const double* dvec, double* workspace, double* out,
                   void* stream);

/* K8 -- VFE (Titsias) helpers for `PseudoObs` (model.py:286-287).
 * gpar_transpose_scale: dst (cols x rows) = (diag(scale) src)^T, scale may be NULL.
 * gpar_vfe_rowterms: out[0] = sum_j (k_jj - ||Bt_j||^2)/sigma_j + log(2 pi sigma_j) + y_j^2/sigma_j
 * with Bt = K_xz L_z^-T (n x M), the trace / normaliser / data terms of the bound.  When the prior of the
 * bound is itself a sparse posterior, k_jj is its variance k(x_j, x_j) - ||Pm_j||^2 + ||Pp_j||^2 with the row
 * blocks Pm, Pp (n x Mp, leading dimension ldp; NULL for a prior GP). */
int gpar_transpose_scale(const double* src, int64_t lds, int64_t rows, int64_t cols, const double* scale,
                         double* dst, int64_t ldd, void* stream);
#define GPAR_VFE_ROWTERMS_WS 592 /* doubles of workspace (one partial sum per CTA, combined in a fixed order) */
int gpar_vfe_rowterms(const gpar_kernel_spec_t* spec, const double* X, int64_t ldx, int64_t n, const double* Bt,
                      int64_t ldb, int64_t M, const double* sigma, const double* y, const double* Pm,
                      const double* Pp, int64_t ldp, int64_t Mp, double* workspace, double* out, void* stream);
/* Output transforms on the device (regression.py:22-28, 553-554): every sample is un-normalised and
 * un-transformed BEFORE the S-axis reduction (quirk Q8: mean of exp, not exp of mean).  a is (rows x p) row
 * major, in place: a[r][j] = T^-1(a[r][j] * scale[j] + shift[j]); scale / shift may be NULL (no normalisation).
 * kind: 0 identity, 1 log_transform (T^-1 = exp), 2 squishing_transform (T^-1 = sign(v) (exp|v| - 1)). */
#define GPAR_TRANSFORM_IDENTITY 0
#define GPAR_TRANSFORM_LOG 1
#define GPAR_TRANSFORM_SQUISH 2
int gpar_untransform(double* a, int64_t rows, int64_t p, const double* scale, const double* shift, int kind,
                     void* stream);
/* y <- y + a x (n doubles): combines weight vectors of nested posteriors on the device. */
int gpar_axpy(int64_t n, double a, const double* x, double* y, void* stream);

/* K3 -- alpha <- L^-T u (u is read only; `work` is n doubles of scratch);
 * out2[0] = 2 sum log L_ii, out2[1] = ||u||^2.  Normal.logpdf tail / iqf. */
int gpar_backsolve(const double* L, int64_t ldl, int64_t n, const double* ws, const double* u, double* alpha,
                   double* work, void* stream);
int gpar_logdet_quad(const double* L, int64_t ldl, int64_t n, const double* u, double* out2, void* stream);

/* K6 -- y <- A x (A m x n row major) and the fused cross-covariance product
 * out[j] = sum_i k(Xq[j], Xa[i]) v[i] (posterior mean K(x_, x_a) alpha without
 * materialising K; model.py:298-301). */
int gpar_gemv(const double* A, int64_t lda, int64_t m, int64_t n, const double* x, double* y, void* stream);
int gpar_gram_gemv(const gpar_kernel_spec_t* spec, const double* Xq, int64_t ldq, int64_t nq, const double* Xa,
                   int64_t lda, int64_t na, const double* v, double* out, void* stream);

/* K7 -- joint Gaussian draws with injected normals (Normal.sample; model.py:264-270):
 * for b < batch, s < ns:  out[b][s][i] = mean_b[i] + sum_{j<=i} C_b[i][j] Z[b][s][j]
 *                                        (+ sd_b[i] * Z2[b][s][i] when Z2 != NULL).
 * Z/out are (batch*ns) x n row major; mean (stride n) and sd (stride strideSd; 0 = one vector shared by
 * the batch) may be NULL. */
int gpar_sample_affine(const double* C, int64_t ldc, int64_t n, int64_t strideC, const double* mean,
                       const double* sd, int64_t strideSd, const double* Z, const double* Z2, int64_t ns,
                       int64_t batch, double* out, void* stream);

/* K9 -- row gather / column scatter used by per_output masks and `_update_inputs`
 * (model.py:165,220,291-322).  idx are int64 row indices (device). */
int gpar_gather_rows(const double* src, int64_t lds, const int64_t* idx, int64_t n_out, int64_t ncols, double* dst,
                     int64_t ldd, void* stream);
int gpar_scatter_col(double* dst, int64_t ldd, int64_t col, const int64_t* idx, const double* src, int64_t n,
                     void* stream);
/* out[i] = y[i] - (d[i] + eps) * alpha[i]: posterior mean at the training rows via the
 * identity K alpha = y - (Sigma + eps I) alpha. */
int gpar_mean_identity(const double* y, const double* d, double eps, const double* alpha, int64_t n, double* out,
                       void* stream);

/* Device-side reductions over the sample axis (regression.py:589): out[i] = mean_s in[s][i]. */
int gpar_mean_axis0(const double* in, int64_t ns, int64_t n, double* out, void* stream);

/* inout[i] += sum_k in[k][i] (k ascending): combines the K-slices of a split-K SYRK (the M x M matrix
 * A = I + B Sigma^-1 B^T of the VFE path has few tiles and a very long K) in a fixed order. */
int gpar_sum_axis0_add(const double* in, int64_t ns, int64_t n, double* inout, void* stream);

/* Two percentiles over the sample axis with numpy's default ("linear") interpolation -- the credible bounds of
 * regression.py:593-594.  The caller passes, per percentile, the lower order statistic j = floor(h) and the
 * weight g = h - j of numpy's virtual index h (computed on the host with numpy's own formula, so the weights
 * are bit-identical); out = lerp(sorted[j], sorted[min(j + 1, ns - 1)], g) per entry.  NaNs are not supported. */
int gpar_percentile2_axis0(const double* in, int64_t ns, int64_t n, int64_t j_lo, double g_lo, int64_t j_hi,
                           double g_hi, double* out_lo, double* out_hi, void* stream);

/* Diagnostics (probes, profiling hooks, the host decode of the kernel's task list) are NOT part of the drop-in
 * ABI: they are declared in gpar_b200_debug.h; the raw issue-rate probes live in their own library
 * (libgpar_b200_debug.so, csrc/debug.cu). */


#ifdef __cplusplus
}
#endif
#endif /* GPAR_B200_H */
