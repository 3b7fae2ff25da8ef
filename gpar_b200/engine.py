"""Device engine: thin Python wrappers over the C ABI plus the ``Factor`` object
(Cholesky of stacked observation blocks with optional appended test rows) that
the GPAR host loop in :mod:`gpar_b200.model` drives.

torch is plumbing only (device memory, streams, host<->device copies); every
arithmetic step on the hot path is one of our sm_100a kernels."""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import TILE, check

__all__ = ["Engine", "Factor"]

F64 = torch.float64


def _even(n):
    return n + (n & 1)


def wave_efficiency(tiles, sms):
    """Fraction of the SMs' time a pass of ``tiles`` 128-row blocks keeps busy in ``gpar_trsm_rows``: whole
    waves of one block per SM, then the rest as blocks of 32 / 64 / 96 / 128 rows in one more wave
    (``trsm_row_plan`` in csrc/potrf.cu) that costs about 0.4 / 0.6 / 0.8 / 1.0 of a full one (measured on a
    B200)."""
    w, r = divmod(int(tiles), int(sms))
    tail = 0.0 if r == 0 else (0.4, 0.6, 0.8, 1.0)[min(-(-(r * 4) // sms), 4) - 1]
    return tiles / float((w + tail) * sms)


class Engine:
    """Owns the library handle, the device and the launch counter."""

    def __init__(self, device=None, epsilon=1e-12, group=None, shard_min_n=4096):
        """``group``: a torch.distributed process group (one process per GPU of one NVLink domain).
        When set, every factorisation of at least ``shard_min_n`` rows is spread over its ranks
        (:func:`gpar_b200.dist.potrf_sharded`); all ranks must then drive the engine with the same
        calls and the same data (SPMD)."""
        if not torch.cuda.is_available():
            raise _lib.GparError("gpar_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.epsilon = float(epsilon)
        self.launches = 0  # kernels of ours launched so far
        self.flops = 0.0  # algorithmic flops of the dense contractions issued (bookkeeping for bench.py)
        self._infos = []  # device `info` words of the factorizations issued since the last check
        self.group, self.shard_min_n = group, int(shard_min_n)
        self._peer_bufs = []  # peer-mapped allocations of the sharded factorisations still alive
        self._peer_pool = {}  # nbytes -> free peer-mapped allocations (reused by later factorisations)

    def sharded(self, n):
        """True if a factorisation of n rows is spread over the ranks of ``self.group``."""
        if self.group is None or n < self.shard_min_n:
            return False
        import torch.distributed as dist

        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def peer_buffer(self, nbytes):
        """Peer-mapped buffer of ``nbytes`` for a sharded factorisation: reused from the pool when one of
        that size is free (cudaMalloc + IPC exchange only happen the first time).  Collective and
        deterministic: every rank makes the same calls in the same order, so pool positions line up."""
        from .dist import PeerBuffer

        free = self._peer_pool.setdefault(int(nbytes), [])
        buf = free.pop() if free else PeerBuffer(self, nbytes, self.group)
        self._peer_bufs.append(buf)
        return buf

    def free_peer_buffers(self):
        """Hand the peer-mapped buffers of earlier sharded factorisations back to the pool (their Factor
        objects become invalid).  Called by GPARRegressor at the start of every public call; local
        bookkeeping only, but every rank must call it at the same point."""
        for b in self._peer_bufs:
            self._peer_pool.setdefault(b.nbytes, []).append(b)
        self._peer_bufs = []

    def release_peer_buffer(self, buf):
        """Return ONE peer-mapped buffer to the pool (Factor.release).  SPMD: every rank releases the
        same factor at the same point of the program, so the pools stay aligned.  Reuse is safe without a
        device sync here: the next sharded factorisation starts with reset -> synchronize -> barrier."""
        for k, b in enumerate(self._peer_bufs):
            if b is buf:
                del self._peer_bufs[k]
                self._peer_pool.setdefault(b.nbytes, []).append(b)
                return

    def free_bytes(self):
        """Device memory a new allocation can still get: free on the device plus what torch's caching
        allocator holds but does not use."""
        free, _ = torch.cuda.mem_get_info(self.device)
        return int(free + torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device))

    def chain_chunk(self, S, bytes_per_chain, tiles_per_chain=1, frac=0.6):
        """Number of diverged chains to process per pass (model.py:557-564 runs them one by one; the
        arithmetic is independent per chain): as many as fit into ``frac`` of the free device memory
        (or ``GPAR_CHAIN_CHUNK_BYTES``), trimmed so that the row blocks of a pass fill the SMs: whole
        waves of 128-row blocks plus one wave of 32 / 64 / 96-row blocks for the rest (``trsm_row_plan`` in
        csrc/potrf.cu), which cost about 0.4 / 0.6 / 0.8 of a full wave (measured)."""
        budget = int(os.environ.get("GPAR_CHAIN_CHUNK_BYTES", 0)) or int(frac * self.free_bytes())
        hi = int(max(1, min(S, budget // max(int(bytes_per_chain), 1))))
        if hi >= S:
            return int(S)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        best, best_eff = hi, 0.0
        for c in range(hi, max(hi * 3 // 4, 1) - 1, -1):
            eff = wave_efficiency(c * max(int(tiles_per_chain), 1), sms)
            if eff > best_eff + 1e-9:
                best, best_eff = c, eff
        return int(best)

    def standard_normal_host(self, n):
        """n standard normals on the host for the ``sample_missing`` draws (model.py:229-237).  With a
        process group the engine runs SPMD -- every rank must build the same next-layer inputs -- so
        rank 0 of the group draws and broadcasts."""
        if self.group is None:
            return np.random.standard_normal(int(n))
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return np.random.standard_normal(int(n))
        t = torch.empty(int(n), dtype=F64, device=self.device)
        if dist.get_rank(self.group) == 0:
            t.copy_(torch.as_tensor(np.random.standard_normal(int(n))))
        dist.broadcast(t, src=dist.get_global_rank(self.group, 0), group=self.group)
        return t.cpu().numpy()

    def close_peer_buffers(self):
        """Collective: unmap and free every pooled peer buffer."""
        self.free_peer_buffers()
        pool, self._peer_pool = self._peer_pool, {}
        for bufs in pool.values():
            for b in bufs:
                b.close()

    def check_infos(self):
        """Read back the pivot status of every factorization issued since the last call (one
        device->host copy) and raise if any matrix was not positive definite -- the counterpart
        of the LinAlgError a failed torch Cholesky raises in the reference."""
        infos, self._infos = self._infos, []
        if not infos:
            return
        vals = torch.cat([i.reshape(-1) for i in infos]).cpu().numpy()
        if (vals != 0).any():
            k = int(vals[vals != 0][0])
            raise _lib.GparError(f"Cholesky failed: matrix not positive definite (first non-positive pivot {k})")

    # -- plumbing -----------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, *shape, dtype=F64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=F64):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def to_device(self, a, dtype=F64):
        t = torch.as_tensor(np.ascontiguousarray(a))
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True)

    @staticmethod
    def addr(t, offset=0):
        return C.c_void_p(t.data_ptr() + 8 * int(offset))

    # -- K1 -------------------------------------------------------------------
    def gram(self, spec, X, ldx, nx, out, ldo, Y=None, ldy=0, ny=0, diag=None, lower_only=True, x_off=0, y_off=0,
             out_off=0):
        rc = self.lib.gpar_gram(C.byref(spec), self.addr(X, x_off), ldx, nx, None if Y is None else self.addr(Y, y_off),
                                ldy, ny, None if diag is None else self.addr(diag), self.epsilon, int(lower_only),
                                self.addr(out, out_off), ldo, self.stream)
        check(rc, "gpar_gram")
        self.launches += 1

    def gram_batched(self, spec, X, ldx, nx, strideX, out, ldo, strideO, batch, diag=None, strideD=0, lower_only=True):
        rc = self.lib.gpar_gram_batched(C.byref(spec), self.addr(X), ldx, nx, strideX, None, 0, 0, 0,
                                        None if diag is None else self.addr(diag), strideD, self.epsilon,
                                        int(lower_only), self.addr(out), ldo, strideO, batch, self.stream)
        check(rc, "gpar_gram_batched")
        self.launches += 1

    # -- K2 -------------------------------------------------------------------
    def potrf(self, A, lda, n, B=None, ldb=0, nb=0, batch=1, strideA=0, strideB=0, a_off=0):
        nbytes = self.lib.gpar_potrf_workspace_bytes(n, nb, batch)
        ws = self.empty(max(nbytes // 8, 2))
        info = torch.zeros(batch, dtype=torch.int32, device=self.device)
        rc = self.lib.gpar_potrf(self.addr(A, a_off), lda, n, strideA, None if B is None else self.addr(B), ldb, nb,
                                 strideB, batch, self.addr(ws), C.c_void_p(info.data_ptr()), self.stream)
        check(rc, "gpar_potrf")
        nt = (n + TILE - 1) // TILE
        if os.environ.get("GPAR_POTRF_V1"):
            self.launches += nt + max(nt - 1, 0) + (nt if nb > 0 else max(nt - 1, 0))
        else:
            self.launches += 1  # one persistent dataflow kernel
        self.flops += batch * (n ** 3 / 3.0 + nb * float(n) ** 2)
        self._infos.append(info)
        return ws, info

    def trsm_rows(self, L, ldl, n, ws, B, ldb, nb):
        scratch = self.empty(max(self.lib.gpar_trsm_rows_scratch_bytes(nb) // 8, 2))
        rc = self.lib.gpar_trsm_rows(self.addr(L), ldl, n, self.addr(ws), self.addr(B), ldb, nb, self.addr(scratch),
                                     self.stream)
        check(rc, "gpar_trsm_rows")
        self.launches += 1
        self.flops += nb * float(n) ** 2

    def syrk_sub(self, Cm, ldc, n, W, ldw, k, batch=1, strideC=0, strideW=0, c_off=0, w_off=0):
        rc = self.lib.gpar_syrk_sub(self.addr(Cm, c_off), ldc, n, strideC, self.addr(W, w_off), ldw, k, strideW, batch,
                                    self.stream)
        check(rc, "gpar_syrk_sub")
        self.launches += 1
        self.flops += batch * float(n) ** 2 * k

    def syrk_add(self, Cm, ldc, n, W, ldw, k, batch=1, strideC=0, strideW=0):
        rc = self.lib.gpar_syrk_add(self.addr(Cm), ldc, n, strideC, self.addr(W), ldw, k, strideW, batch, self.stream)
        check(rc, "gpar_syrk_add")
        self.launches += 1
        self.flops += batch * float(n) ** 2 * k

    # -- K10 ------------------------------------------------------------------
    def potri(self, L, ldl, n, ws, return_U=False):
        """A^-1 (lower triangle, leading dimension ldl) from the factor L and the workspace of its potrf;
        ``return_U``: also the upper-triangular U = L^-T it is built from."""
        U = self.empty(max(n, 1) * ldl)
        Ainv = self.empty(max(n, 1) * ldl)
        scratch = self.empty(max(self.lib.gpar_potri_scratch_bytes(n) // 8, 2))
        rc = self.lib.gpar_potri(self.addr(L), ldl, n, self.addr(ws), self.addr(U), ldl, self.addr(Ainv), ldl,
                                 self.addr(scratch), self.stream)
        check(rc, "gpar_potri")
        self.launches += 3
        self.flops += 2.0 * n ** 3 / 3.0
        return (Ainv, U) if return_U else Ainv

    def gram_wgrad(self, spec, X, ldx, nx, Y, ldy, ny, G=None, ldg=0, sx=None, ux=None, uy=None):
        """Raw chain-rule sums of sum_ij (ux_i uy_j + sx_i G_ij) d k(x_i, y_j) / d spec (device, GRAD_NP)."""
        wsg = self.empty(max(self.lib.gpar_gram_wgrad_workspace_bytes(nx, ny) // 8, 2))
        out = self.empty(_lib.GRAD_NP)
        rc = self.lib.gpar_gram_wgrad(C.byref(spec), self.addr(X), ldx, nx, self.addr(Y), ldy, ny,
                                      None if G is None else self.addr(G), ldg, None if sx is None else self.addr(sx),
                                      None if ux is None else self.addr(ux), None if uy is None else self.addr(uy),
                                      self.addr(wsg), self.addr(out), self.stream)
        check(rc, "gpar_gram_wgrad")
        self.launches += 2
        return out

    def row_sqnorm(self, A, lda, n, k):
        out = self.empty(max(n, 1))
        rc = self.lib.gpar_row_sqnorm(self.addr(A), lda, n, k, self.addr(out), self.stream)
        check(rc, "gpar_row_sqnorm")
        self.launches += 1 if n > 0 else 0
        return out

    def gram_grad(self, spec, X, ldx, n, alpha, Ainv, lda, dvec=None):
        """Raw chain-rule sums of d LML / d spec (device tensor of _lib.GRAD_NP doubles)."""
        wsg = self.empty(max(self.lib.gpar_gram_grad_workspace_bytes(n) // 8, 2))
        out = self.empty(_lib.GRAD_NP)
        rc = self.lib.gpar_gram_grad(C.byref(spec), self.addr(X), ldx, n, self.addr(alpha), self.addr(Ainv), lda,
                                     None if dvec is None else self.addr(dvec), self.addr(wsg), self.addr(out),
                                     self.stream)
        check(rc, "gpar_gram_grad")
        self.launches += 2
        return out

    # -- K8 -------------------------------------------------------------------
    def transpose_scale(self, src, lds, rows, cols, scale, dst, ldd):
        rc = self.lib.gpar_transpose_scale(self.addr(src), lds, rows, cols, None if scale is None else self.addr(scale),
                                           self.addr(dst), ldd, self.stream)
        check(rc, "gpar_transpose_scale")
        self.launches += 1

    def gemm_nt(self, Cm, ldc, m, n, A, lda, B, ldb, k, add=False):
        """C (m x n) -= A B^T (add: +=), A is m x k, B is n x k."""
        rc = self.lib.gpar_gemm_nt(self.addr(Cm), ldc, m, n, self.addr(A), lda, self.addr(B), ldb, k, int(add),
                                   self.stream)
        check(rc, "gpar_gemm_nt")
        self.launches += 1 if m * n * k > 0 else 0
        self.flops += 2.0 * m * n * k

    def axpy(self, n, a, x, y):
        rc = self.lib.gpar_axpy(n, float(a), self.addr(x), self.addr(y), self.stream)
        check(rc, "gpar_axpy")
        self.launches += 1 if n > 0 else 0

    def vfe_rowterms(self, spec, X, ldx, n, Bt, ldb, M, sigma, y, out, out_off=0, Pm=None, Pp=None, ldp=0, Mp=0):
        wsr = self.empty(592)  # GPAR_VFE_ROWTERMS_WS
        rc = self.lib.gpar_vfe_rowterms(C.byref(spec), self.addr(X), ldx, n, self.addr(Bt), ldb, M, self.addr(sigma),
                                        self.addr(y), None if Pm is None else self.addr(Pm),
                                        None if Pp is None else self.addr(Pp), ldp, Mp, self.addr(wsr),
                                        self.addr(out, out_off), self.stream)
        check(rc, "gpar_vfe_rowterms")
        self.launches += 2

    # -- K3 -------------------------------------------------------------------
    def backsolve(self, L, ldl, n, ws, u):
        alpha = self.empty(max(n, 1))
        work = self.empty(max(n, 1))
        rc = self.lib.gpar_backsolve(self.addr(L), ldl, n, self.addr(ws), self.addr(u), self.addr(alpha),
                                     self.addr(work), self.stream)
        check(rc, "gpar_backsolve")
        self.launches += 1
        return alpha

    def logdet_quad(self, L, ldl, n, u, out2, l_off=0, u_off=0, out_off=0):
        rc = self.lib.gpar_logdet_quad(self.addr(L, l_off), ldl, n, None if u is None else self.addr(u, u_off),
                                       self.addr(out2, out_off), self.stream)
        check(rc, "gpar_logdet_quad")
        self.launches += 1

    # -- K6 -------------------------------------------------------------------
    def gemv(self, A, lda, m, n, x, y, a_off=0):
        rc = self.lib.gpar_gemv(self.addr(A, a_off), lda, m, n, self.addr(x), self.addr(y), self.stream)
        check(rc, "gpar_gemv")
        self.launches += 1 if m > 0 else 0

    def gram_gemv(self, spec, Xq, ldq, nq, Xa, lda, na, v, out):
        rc = self.lib.gpar_gram_gemv(C.byref(spec), self.addr(Xq), ldq, nq, self.addr(Xa), lda, na, self.addr(v),
                                     self.addr(out), self.stream)
        check(rc, "gpar_gram_gemv")
        self.launches += 1 if nq > 0 else 0

    # -- K7 -------------------------------------------------------------------
    def sample_affine(self, Cm, ldc, n, Z, out, ns, batch=1, strideC=0, mean=None, sd=None, Z2=None, c_off=0,
                      strideSd=None):
        """``strideSd``: distance between the noise-sd vectors of consecutive matrices of the batch
        (default n; 0 = one vector shared by the whole batch)."""
        rc = self.lib.gpar_sample_affine(self.addr(Cm, c_off), ldc, n, strideC, None if mean is None else self.addr(mean),
                                         None if sd is None else self.addr(sd), n if strideSd is None else strideSd,
                                         self.addr(Z), None if Z2 is None else self.addr(Z2), ns, batch,
                                         self.addr(out), self.stream)
        check(rc, "gpar_sample_affine")
        self.launches += 1

    # -- K9 -------------------------------------------------------------------
    def gather_rows(self, src, lds, idx, n_out, ncols, dst, ldd, src_off=0, dst_off=0):
        rc = self.lib.gpar_gather_rows(self.addr(src, src_off), lds, None if idx is None else C.c_void_p(idx.data_ptr()),
                                       n_out, ncols, self.addr(dst, dst_off), ldd, self.stream)
        check(rc, "gpar_gather_rows")
        self.launches += 1 if n_out * ncols > 0 else 0

    def scatter_col(self, dst, ldd, col, idx, src, n, src_off=0):
        rc = self.lib.gpar_scatter_col(self.addr(dst), ldd, col, None if idx is None else C.c_void_p(idx.data_ptr()),
                                       self.addr(src, src_off), n, self.stream)
        check(rc, "gpar_scatter_col")
        self.launches += 1 if n > 0 else 0

    def mean_identity(self, y, d, alpha, n, out, off=0, out_off=0):
        rc = self.lib.gpar_mean_identity(self.addr(y, off), self.addr(d, off), self.epsilon, self.addr(alpha, off), n,
                                         self.addr(out, out_off), self.stream)
        check(rc, "gpar_mean_identity")
        self.launches += 1 if n > 0 else 0

    def untransform(self, a, rows, p, scale, shift, kind):
        """In place: a[r][j] = T^-1(a[r][j] * scale[j] + shift[j]) (kind 0 identity, 1 exp, 2 inverse squish)."""
        rc = self.lib.gpar_untransform(self.addr(a), rows, p, None if scale is None else self.addr(scale),
                                       None if shift is None else self.addr(shift), int(kind), self.stream)
        check(rc, "gpar_untransform")
        self.launches += 1 if rows * p > 0 else 0

    def mean_axis0(self, inp, ns, n, out):
        rc = self.lib.gpar_mean_axis0(self.addr(inp), ns, n, self.addr(out), self.stream)
        check(rc, "gpar_mean_axis0")
        self.launches += 1

    def sum_axis0_add(self, inp, ns, n, inout):
        rc = self.lib.gpar_sum_axis0_add(self.addr(inp), ns, n, self.addr(inout), self.stream)
        check(rc, "gpar_sum_axis0_add")
        self.launches += 1

    @staticmethod
    def numpy_virtual_index(ns, q):
        """(j, gamma) of np.percentile(..., q) with the default linear method for ns samples, computed with
        numpy's own floating-point formula (numpy/lib/_function_base_impl.py: virtual index
        ``(n - 1) * quantiles`` with ``quantiles = q / 100``, gamma = index - floor(index))."""
        quant = np.true_divide(q, 100.0)
        vi = (ns - 1) * quant
        j = int(np.floor(vi))
        j = min(max(j, 0), ns - 1)
        return j, float(vi - j)

    def percentile2_axis0(self, inp, ns, n, q_lo, q_hi):
        """Two percentiles over axis 0 of an (ns, n) device block; returns two device vectors."""
        lo, hi = self.empty(max(n, 1)), self.empty(max(n, 1))
        jl, gl = self.numpy_virtual_index(ns, q_lo)
        jh, gh = self.numpy_virtual_index(ns, q_hi)
        rc = self.lib.gpar_percentile2_axis0(self.addr(inp), ns, n, jl, gl, jh, gh, self.addr(lo), self.addr(hi),
                                             self.stream)
        check(rc, "gpar_percentile2_axis0")
        self.launches += 1
        return lo, hi

    def fp64_probe(self, mode, iters):
        """Raw DMMA (mode 0) / DFMA (mode 1) issue-rate probe of the diagnostics library
        (libgpar_b200_debug.so, include/gpar_b200_debug.h); returns the flops of the launch."""
        sink = self.zeros(2)
        flops = C.c_double(0.0)
        rc = _lib.load_debug().gpar_fp64_probe(mode, iters, self.addr(sink), C.byref(flops), self.stream)
        if rc != 0:
            raise _lib.GparError(f"gpar_fp64_probe failed with code {rc}")
        return flops.value


class Factor:
    """Cholesky of ``K(X, X) + diag(d) + eps I`` over the stacked rows ``X`` =
    [observation rows (n_obs); appended test rows (n_ext)] with the right-hand
    side ``y`` riding along as an appended row of the sweep.  One gpar_potrf call
    yields:

    * ``L``  = lower factor of the whole joint matrix (in ``J``), whose blocks are
      ``L_oo`` (observations), ``W = K_*o L_oo^-T`` and ``C = chol(K_** + D_* + eps I - W W^T)``;
    * ``u``  = ``L_oo^-1 y`` (first ``n_obs`` entries of the appended row).
    """

    def __init__(self, eng, spec, X, ldx, d, y, n_obs, n_ext=0):
        self.eng, self.spec = eng, spec
        self.X, self.ldx, self.d, self.y = X, ldx, d, y
        self.n_obs, self.n_ext = int(n_obs), int(n_ext)
        n = self.n = self.n_obs + self.n_ext
        self.ld = ld = _even(max(n, 2))
        self._alpha = None
        self._peer_buf = None
        if eng.sharded(n):
            # multi-GPU: the joint matrix lives in peer-mapped memory, tile rows are dealt to the ranks
            from .dist import potrf_layout, potrf_sharded

            lay = potrf_layout(eng, n, 1)
            assert lay["ld"] == ld
            buf = self._peer_buf = eng.peer_buffer(lay["bytes"])
            self.J = buf.view(lay["a"], n * ld)
            self.u = buf.view(lay["b"], ld)
            self.u.zero_()
            if self.n_obs > 0:
                self.u[: self.n_obs].copy_(y[: self.n_obs])
            eng.gram(spec, X, ldx, n, self.J, ld, diag=d, lower_only=True)
            potrf_sharded(eng, buf, n, 1, eng.group)
            self.ws = buf.view(lay["ws"], eng.lib.gpar_potrf_workspace_bytes(n, 1, 1) // 8)
            self.info = torch.zeros(1, dtype=torch.int32, device=eng.device)
            return
        self.J = eng.empty(max(n, 1) * ld)
        self.u = eng.zeros(ld)
        if self.n_obs > 0:
            self.u[: self.n_obs].copy_(y[: self.n_obs])
        if n > 0:
            eng.gram(spec, X, ldx, n, self.J, ld, diag=d, lower_only=True)
            self.ws, self.info = eng.potrf(self.J, ld, n, B=self.u, ldb=ld, nb=1)
        else:
            self.ws, self.info = eng.empty(2), torch.zeros(1, dtype=torch.int32, device=eng.device)

    def release(self):
        """Drop the factor: a sharded factor hands its peer-mapped buffer back to the engine's pool (the
        next sharded factorisation reuses it); single-GPU factors just drop their references."""
        if self._peer_buf is not None:
            self.eng.release_peer_buffer(self._peer_buf)
            self._peer_buf = None
        self.J = self.u = self.ws = None

    def logdet_quad(self, out2, out_off, r0, r1):
        """out2[out_off:out_off+2] = (2 sum_{r0<=i<r1} log L_ii, sum u_i^2)."""
        self.eng.logdet_quad(self.J, self.ld, r1 - r0, self.u, out2, l_off=r0 * self.ld + r0, u_off=r0, out_off=out_off)

    def alpha(self):
        """alpha = (K + D + eps I)^-1 y over the observation rows."""
        if self._alpha is None:
            self._alpha = self.eng.backsolve(self.J, self.ld, self.n_obs, self.ws, self.u)
        return self._alpha

    def mean_obs(self, out, r0, r1, out_off=0):
        """Posterior mean at observation rows [r0, r1): y - (d + eps) alpha, the exact
        identity K alpha = y - (D + eps I) alpha (no n^2 work)."""
        self.eng.mean_identity(self.y, self.d, self.alpha(), r1 - r0, out, off=r0, out_off=out_off)

    def mean_at(self, Xq, ldq, nq, out):
        """Posterior mean at arbitrary rows: K(Xq, X_obs) alpha, fused (no cross-Gram in HBM)."""
        if self.n_obs == 0:
            out[:nq].zero_()
            return
        self.eng.gram_gemv(self.spec, Xq, ldq, nq, self.X, self.ldx, self.n_obs, self.alpha(), out)

    def ext_mean(self, out):
        """Posterior mean at the appended rows: W u."""
        if self.n_obs == 0:
            out[: self.n_ext].zero_()
            return
        self.eng.gemv(self.J, self.ld, self.n_ext, self.n_obs, self.u, out, a_off=self.n_obs * self.ld)

    def ext_sample(self, Z, ns, out, mean=None, sd=None, Z2=None):
        """out[s] = mean + C Z[s] (+ sd * Z2[s]) with C the factor of the posterior
        covariance at the appended rows (joint draw)."""
        self.eng.sample_affine(self.J, self.ld, self.n_ext, Z, out, ns, mean=mean, sd=sd, Z2=Z2,
                               c_off=self.n_obs * self.ld + self.n_obs)
