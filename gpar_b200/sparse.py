"""Inducing-point (VFE / Titsias) path of GPAR -- the counterpart of stheno's ``PseudoObs``
as driven by gpar/model.py:286-287, 303-305.

Per layer, with inducing inputs z (M x d), observed rows x (n x d), noise diagonal sigma = noise / w
and targets y (prior mean zero):

    L_z  = chol(K_zz + eps I)                      B^T = K_xz L_z^-T          (one gpar_potrf call: K_xz
                                                                              rides along as appended rows)
    A    = I + B Sigma^-1 B^T                      c   = B Sigma^-1 y
    ELBO = -1/2 [ sum_j (k_jj - |b_j|^2)/sigma_j + sum_j log(2 pi sigma_j) + logdet A
                  + sum_j y_j^2/sigma_j - c^T A^-1 c ]
    mean(x_)      = K(x_, z) beta,   beta = L_z^-T A^-1 c
    cov(x_, x_')  = k - B_x^T B_x' + (L_A^-1 B_x)^T (L_A^-1 B_x')

Every step is one of the engine's kernels; the host only prepares sigma-derived vectors.
"""
import math

import numpy as np
import torch

from .engine import Factor, F64, _even
from .model import DevMat, ObsBlock, _stack, construct_model, last, per_output
from .spec import LayerModel

__all__ = ["SparseFactor", "logpdf_sparse", "condition_sparse", "sample_sparse"]


class SparseBlock:
    """Observations of a sparse posterior layer: inducing inputs, observed rows, targets, noise."""

    def __init__(self, Z, X, y_host, sig_host):
        self.Z, self.X, self.y_host, self.sig_host = Z, X, y_host, sig_host
        self.n = X.n
        self.factor = None


class SparseFactor:
    """VFE quantities of one layer.  ``prior``: the :class:`SparseFactor` of the sparse posterior this layer
    was conditioned on before (``logpdf(posterior=True)``, regression.py:495-499 -> model.py:286-287 with ``f``
    a posterior GP): the bound is then taken under that posterior, i.e. with

        m~(a)    = K(a, z_t) beta_t
        k~(a, b) = k(a, b) - P_a P_b^T + Q_a Q_b^T,   P_a = K(a, z_t) L_zt^-T,  Q_a = P_a L_At^-T

    in the place of the zero mean and the prior kernel."""

    def __init__(self, eng, spec, Z, X, y_host, sig_host, prior=None):
        self.eng, self.spec, self.Z, self.X, self.prior = eng, spec, Z, X, prior
        M, n = Z.n, X.n
        self.M, self.n = M, n
        ldz = self.ldz = _even(max(M, 2))
        sig = eng.to_device(sig_host)
        Px = Qx = None
        if prior is not None:
            Mt, ldt = prior.M, prior.ldz
            self.Pz, self.Qz = prior._rows_to_factors(Z, M)  # M x Mt
            if n:
                Px, Qx = prior._rows_to_factors(X, n)
                m_x = eng.empty(n)
                prior.mean_at(X.t, X.ld, n, m_x)
                y_host = np.asarray(y_host, dtype=np.float64) - m_x.cpu().numpy()  # y - m~(x)
        yv = eng.to_device(y_host)
        # L_z and B^T = K_xz L_z^-T in one sweep
        self.Jz = eng.empty(M * ldz)
        eng.gram(spec, Z.t, Z.ld, M, self.Jz, ldz, lower_only=True)  # + eps I on the diagonal
        if prior is not None:
            eng.syrk_sub(self.Jz, ldz, M, self.Pz, ldt, Mt)
            eng.syrk_add(self.Jz, ldz, M, self.Qz, ldt, Mt)
        self.Bt = eng.empty(max(n, 1) * ldz)
        if n:
            eng.gram(spec, X.t, X.ld, n, self.Bt, ldz, Y=Z.t, ldy=Z.ld, ny=M, lower_only=False)
            if prior is not None:
                eng.gemm_nt(self.Bt, ldz, n, M, Px, ldt, self.Pz, ldt, Mt, add=False)
                eng.gemm_nt(self.Bt, ldz, n, M, Qx, ldt, self.Qz, ldt, Mt, add=True)
        self.ws_z, self.info_z = eng.potrf(self.Jz, ldz, M, B=self.Bt if n else None, ldb=ldz, nb=n)
        # A = I + C^T C with C = diag(sigma^-1/2) B^T, formed as an NT SYRK on C^T (M x n)
        ldn = _even(max(n, 2))
        self.JA = eng.zeros(M * ldz)
        # identity (+ the jitter every reference Cholesky carries, matrix.cholesky): diagonal = stride ld + 1
        eng.scatter_col(self.JA, ldz + 1, 0, None, eng.to_device(np.full(M, 1.0 + eng.epsilon)), M)
        self.c = eng.zeros(ldz)
        self.terms = eng.zeros(4)  # [row terms, logdet A, |v|^2]
        if n:
            # The M x M SYRK has few tiles (10 at M = 512) and a very long K = n: split K into slices run as
            # one batched SYRK into per-slice partials (zero padding beyond n), summed in slice order.
            nsl = int(min(16, max(1, n // 2048)))
            Ks = -(-n // nsl)
            Ks += Ks & 1
            ldn = nsl * Ks
            Ct = eng.zeros(M * ldn) if ldn > n else eng.empty(M * ldn)
            eng.transpose_scale(self.Bt, ldz, n, M, eng.to_device(1.0 / np.sqrt(sig_host)), Ct, ldn)
            if nsl == 1:
                eng.syrk_add(self.JA, ldz, M, Ct, ldn, n)
            else:
                part = eng.zeros(nsl * M * ldz)
                eng.syrk_add(part, ldz, M, Ct, ldn, Ks, batch=nsl, strideC=M * ldz, strideW=Ks)
                eng.sum_axis0_add(part, nsl, M * ldz, self.JA)
            eng.gemv(Ct, ldn, M, n, eng.to_device(y_host / np.sqrt(sig_host)), self.c)
            if prior is not None:
                eng.vfe_rowterms(spec, X.t, X.ld, n, self.Bt, ldz, M, sig, yv, self.terms, Pm=Px, Pp=Qx, ldp=ldt,
                                 Mp=Mt)
            else:
                eng.vfe_rowterms(spec, X.t, X.ld, n, self.Bt, ldz, M, sig, yv, self.terms)
        # v = L_A^-1 c rides along as an appended row
        self.ws_A, self.info_A = eng.potrf(self.JA, ldz, M, B=self.c, ldb=ldz, nb=1)
        eng.logdet_quad(self.JA, ldz, M, self.c, self.terms, out_off=1)
        t = self.t = eng.backsolve(self.JA, ldz, M, self.ws_A, self.c)  # A^-1 c
        self.beta = eng.backsolve(self.Jz, ldz, M, self.ws_z, t)
        self.y_host, self.sig_host = np.asarray(y_host, dtype=np.float64), np.asarray(sig_host, dtype=np.float64)
        if prior is not None:
            # posterior-of-posterior mean: m~(q) + k~(q, z) beta = K(q, z_t) (beta_t - h1 + h2) + K(q, z) beta with
            # h1 = L_zt^-T (P_z^T beta), h2 = L_zt^-T L_At^-T (Q_z^T beta)
            PzT, QzT = eng.empty(Mt * ldz), eng.empty(Mt * ldz)
            eng.transpose_scale(self.Pz, ldt, M, Mt, None, PzT, ldz)
            eng.transpose_scale(self.Qz, ldt, M, Mt, None, QzT, ldz)
            g1, g2 = eng.zeros(ldt), eng.zeros(ldt)
            eng.gemv(PzT, ldz, Mt, M, self.beta, g1)
            eng.gemv(QzT, ldz, Mt, M, self.beta, g2)
            h1 = eng.backsolve(prior.Jz, ldt, Mt, prior.ws_z, g1)
            h2 = eng.backsolve(prior.Jz, ldt, Mt, prior.ws_z, eng.backsolve(prior.JA, ldt, Mt, prior.ws_A, g2))
            self.wt = eng.zeros(ldt)
            eng.axpy(Mt, 1.0, prior.weights_t(), self.wt)
            eng.axpy(Mt, -1.0, h1, self.wt)
            eng.axpy(Mt, 1.0, h2, self.wt)

    def elbo_grad_raw(self, w_host):
        """Gradient of this layer's bound w.r.t. its kernel spec and noise (VFE analogue of
        gpar_potri + gpar_gram_grad; the reference differentiates PseudoObs with autograd,
        regression.py:434-459).  With R = B^T = K_xz L_z^-T, U_A = L_A^-T, U_z = L_z^-T:

            beta = Sigma^-1 (y - R A^-1 c),  T^T = R L_z^-1,  G2 = Sigma^-1 R (I - A^-1) L_z^-1      (n x M)
            d ELBO = sum_jm (beta_j (T beta)_m + G2_jm) dk(x_j, z_m) - 1/2 sum_mm' ((T beta)(T beta)^T + G2^T T^T)_mm' dk(z_m, z_m')
                     - 1/2 sum_j dk(x_j, x_j) / sigma_j + sum_j g_sigma_j dsigma_j

        (weights pinned against finite differences in oracle/vfe_grad.py).  The two weighted Gram-derivative
        sums run on the device (gpar_gram_wgrad), the O(n) diagonal terms on the host.  Returns the raw
        chain-rule vector (numpy, layout of gpar_gram_grad; last entry = d ELBO / d noise)."""
        if self.prior is not None:
            raise NotImplementedError("gradients of a bound under a sparse posterior are not implemented")
        from . import _lib

        eng, M, n, ldz, spec = self.eng, self.M, self.n, self.ldz, self.spec
        raw = np.zeros(_lib.GRAD_NP)
        if n == 0:
            return raw
        sig, y, R = self.sig_host, self.y_host, self.Bt
        _, Uz = eng.potri(self.Jz, ldz, M, self.ws_z, return_U=True)
        _, UA = eng.potri(self.JA, ldz, M, self.ws_A, return_U=True)
        UAt = eng.empty(M * ldz)
        eng.transpose_scale(UA, ldz, M, M, None, UAt, ldz)  # L_A^-1
        RU = eng.zeros(n * ldz)
        eng.gemm_nt(RU, ldz, n, M, R, ldz, UAt, ldz, M, add=True)  # R U_A: rows (L_A^-1 b_j)^T
        S1 = eng.empty(n * ldz)
        eng.gather_rows(R, ldz, None, n, M, S1, ldz)
        eng.gemm_nt(S1, ldz, n, M, RU, ldz, UA, ldz, M, add=False)  # R (I - A^-1)
        G2 = eng.zeros(n * ldz)
        eng.gemm_nt(G2, ldz, n, M, S1, ldz, Uz, ldz, M, add=True)  # ... L_z^-1 (row scale 1 / sigma applied below)
        Tt = eng.zeros(n * ldz)
        eng.gemm_nt(Tt, ldz, n, M, R, ldz, Uz, ldz, M, add=True)  # T^T = R L_z^-1
        Rt = eng.empty(n)
        eng.gemv(R, ldz, n, M, self.t, Rt)
        beta = (y - Rt.cpu().numpy()) / sig
        beta_d, inv_sig = eng.to_device(beta), eng.to_device(1.0 / sig)
        ldn = _even(max(n, 2))
        RT = eng.empty(M * ldn)
        eng.transpose_scale(R, ldz, n, M, None, RT, ldn)
        Bbeta = eng.zeros(ldz)
        eng.gemv(RT, ldn, M, n, beta_d, Bbeta)
        Tbeta = eng.backsolve(self.Jz, ldz, M, self.ws_z, Bbeta)
        raw_zx = eng.gram_wgrad(spec, self.X.t, self.X.ld, n, self.Z.t, self.Z.ld, M, G=G2, ldg=ldz, sx=inv_sig,
                                ux=beta_d, uy=Tbeta)
        G2T, TT = eng.empty(M * ldn), RT  # RT is free again
        eng.transpose_scale(G2, ldz, n, M, inv_sig, G2T, ldn)
        eng.transpose_scale(Tt, ldz, n, M, None, TT, ldn)
        Gzz = eng.zeros(M * ldz)
        eng.gemm_nt(Gzz, ldz, M, M, G2T, ldn, TT, ldn, n, add=True)  # G2^T T^T
        mhalf = eng.to_device(np.full(M, -0.5))
        uxz = eng.zeros(ldz)
        eng.axpy(M, -0.5, Tbeta, uxz)
        raw_zz = eng.gram_wgrad(spec, self.Z.t, self.Z.ld, M, self.Z.t, self.Z.ld, M, G=Gzz, ldg=ldz, sx=mhalf,
                                ux=uxz, uy=Tbeta)
        b2 = eng.row_sqnorm(R, ldz, n, M).cpu().numpy()
        v2 = eng.row_sqnorm(RU, ldz, n, M).cpu().numpy()
        raw += raw_zx.cpu().numpy() + raw_zz.cpu().numpy()
        # O(n) diagonal terms: -1/2 sum_j dk(x_j, x_j) / sigma_j and the noise derivative
        kdiag, raw_diag = _diag_raw(spec, self.X.to_host(), -0.5 / sig)
        raw += raw_diag
        P = 1.0 / sig - v2 / sig ** 2
        g_sigma = 0.5 * (beta ** 2 - P) + 0.5 * (kdiag - b2) / sig ** 2
        raw[_lib.GRAD_NP - 1] = float(np.sum(g_sigma / np.asarray(w_host, dtype=np.float64)))
        return raw

    def weights_t(self):
        """Weights of K(., z) in this factor's own posterior mean (a prior-level factor: beta)."""
        return self.beta

    def elbo_slot(self):
        """Device triple (row terms, logdet A, |v|^2): ELBO = -1/2 (t0 + t1 - t2)."""
        return self.terms

    def mean_at(self, Xq, ldq, nq, out):
        self.eng.gram_gemv(self.spec, Xq, ldq, nq, self.Z.t, self.Z.ld, self.M, self.beta, out)
        if self.prior is not None and nq > 0:
            tmp = self.eng.empty(nq)
            self.eng.gram_gemv(self.spec, Xq, ldq, nq, self.prior.Z.t, self.prior.Z.ld, self.prior.M, self.wt, tmp)
            self.eng.axpy(nq, 1.0, tmp, out)

    def _rows_to_factors(self, Xq, nq):
        """Bs = K_qz L_z^-T and Ds = Bs L_A^-T for nq query rows."""
        eng, ldz, M = self.eng, self.ldz, self.M
        Bs = eng.empty(max(nq, 1) * ldz)
        eng.gram(self.spec, Xq.t, Xq.ld, nq, Bs, ldz, Y=self.Z.t, ldy=self.Z.ld, ny=M, lower_only=False)
        eng.trsm_rows(self.Jz, ldz, M, self.ws_z, Bs, ldz, nq)
        Ds = eng.empty(max(nq, 1) * ldz)
        eng.gather_rows(Bs, ldz, None, nq, M, Ds, ldz)
        eng.trsm_rows(self.JA, ldz, M, self.ws_A, Ds, ldz, nq)
        return Bs, Ds

    def sample_rows(self, Xs, d_s, Z, S, batch, ns, sd=None, Z2=None):
        """Joint draws at ``batch`` row sets of ``ns`` rows each (stacked in Xs): returns
        (f_col, y_col, mean) with ``S`` draws per set (S = 1 when batch > 1).  Row sets (diverged
        chains) are processed in passes sized from the free device memory."""
        eng, ldz, M = self.eng, self.ldz, self.M
        N = batch * ns
        ldc = _even(max(ns, 2))
        mean = eng.empty(max(N, 1))
        f_col = eng.empty(max(S * N, 1))
        y_col = eng.empty(max(S * N, 1)) if sd is not None else f_col
        per_set = 8 * ns * (ldc + 2 * ldz) + eng.lib.gpar_potrf_workspace_bytes(ns, 0, 2) // 2
        chunk = eng.chain_chunk(batch, per_set, -(-ns // 128)) if batch > 1 else 1
        for b0 in range(0, batch, chunk):
            B = min(batch, b0 + chunk) - b0
            r0 = b0 * ns
            Xc = DevMat(eng, Xs.t[r0 * Xs.ld:], B * ns, Xs.d, Xs.ld)
            Cs = eng.empty(B * ns * ldc)
            eng.gram_batched(self.spec, Xc.t, Xc.ld, ns, ns * Xc.ld, Cs, ldc, ns * ldc, B, diag=d_s, strideD=0)
            Bs, Ds = self._rows_to_factors(Xc, B * ns)
            eng.syrk_sub(Cs, ldc, ns, Bs, ldz, M, batch=B, strideC=ns * ldc, strideW=ns * ldz)
            eng.syrk_add(Cs, ldc, ns, Ds, ldz, M, batch=B, strideC=ns * ldc, strideW=ns * ldz)
            eng.potrf(Cs, ldc, ns, batch=B, strideA=ns * ldc)
            self.mean_at(Xc.t, Xc.ld, B * ns, mean[r0:])
            # batch == 1: S draws from one matrix; batch > 1: one draw per row set (Z rows b0 .. b0 + B)
            Zc = Z if batch == 1 else Z[b0:b0 + B]
            eng.sample_affine(Cs, ldc, ns, Zc, f_col[S * r0:], S, batch=B, strideC=ns * ldc, mean=mean[r0:])
            if sd is not None:
                Z2c = Z2 if batch == 1 else Z2[b0:b0 + B]
                eng.sample_affine(Cs, ldc, ns, Zc, y_col[S * r0:], S, batch=B, strideC=ns * ldc, mean=mean[r0:],
                                  sd=sd, Z2=Z2c, strideSd=0)
        return f_col, y_col, mean


def _diag_raw(spec, Xh, g):
    """k(x_j, x_j) per row and the raw chain-rule sums of sum_j g_j dk(x_j, x_j) / d spec (layout of
    gpar_gram_grad): EQ / RQ / const terms contribute their variance only (zero distance); a linear term
    v sum_f phi_f(x)^2 also feeds the S1 slot of its features (d / d a_f = 2 S1_f / a_f)."""
    from . import _lib

    raw = np.zeros(_lib.GRAD_NP)
    kdiag = np.zeros(Xh.shape[0])
    base = 2 * _lib.MAX_TERMS
    for t in range(spec.n_terms):
        T = spec.terms[t]
        if T.type == _lib.TERM_LINEAR:
            acc = np.zeros(Xh.shape[0])
            for f in range(T.f_begin, T.f_end):
                phi2 = (Xh[:, spec.feat_col[f]] * spec.feat_a[f]) ** 2
                raw[base + 2 * f] += T.variance * float(np.sum(g * phi2))
                acc += phi2
            raw[2 * t] += float(np.sum(g * acc))
            kdiag += T.variance * acc
        else:
            raw[2 * t] += float(np.sum(g))
            kdiag += T.variance
    return kdiag, raw


def _layer(model):
    layer = model()
    return layer if isinstance(layer, LayerModel) else layer[0]


def _train_step(gpar, layer, xd, zd, y_i, w_i, is_last, prior=None, sample_missing=False, normals=None):
    """One layer of the training-side chain (model.py:165-174 / 220-240 with PseudoObs): returns the
    factor and the inputs of the next layer.  ``sample_missing`` (model.py:229-237): the missing rows of
    this output are filled with one joint draw from the sparse posterior ``(f | obs)(x[missing], noise /
    w[missing])`` (``normals``: list of host arrays, one per layer with missing rows, consumed here)."""
    eng = gpar.engine
    avail = ~np.isnan(y_i[:, 0])
    idx = np.flatnonzero(avail)
    Xa = xd if len(idx) == xd.n else xd.copy_rows(eng.to_device(idx, torch.int64), len(idx))
    y_a = y_i[avail, 0]
    sig = layer.noise / w_i[avail]
    fac = SparseFactor(eng, layer.spec, zd, Xa, y_a, sig, prior=prior)
    block = SparseBlock(DevMat(eng, zd.t, zd.n, zd.d, zd.ld), DevMat(eng, Xa.t, Xa.n, Xa.d, Xa.ld), y_a, sig)
    block.factor = fac
    if is_last:
        return fac, block, xd, zd
    # inducing inputs of the next layer (model.py:304-305)
    est_z = eng.empty(zd.n)
    fac.mean_at(zd.t, zd.ld, zd.n, est_z)
    zd_next = zd.with_col(est_z)
    # y column of the next layer (model.py:307-320)
    n_i = xd.n
    col = eng.to_device(y_i[:, 0])
    miss = ~avail
    if sample_missing and miss.any():
        n_m = int(miss.sum())
        idx_m = eng.to_device(np.flatnonzero(miss), torch.int64)
        Xm = xd.copy_rows(idx_m, n_m)
        z = normals.pop(0) if normals is not None else eng.standard_normal_host(n_m)
        Z = eng.to_device(np.asarray(z, dtype=np.float64).reshape(1, n_m))
        y_m, _, _ = fac.sample_rows(Xm, eng.to_device(layer.noise / w_i[miss]), Z, 1, 1, n_m)
        eng.scatter_col(col, 1, 0, idx_m, y_m, n_m)
        # after the merge every row counts as observed (model.py:237, 292)
        avail, miss = np.ones(n_i, dtype=bool), np.zeros(n_i, dtype=bool)
    if gpar.impute and gpar.replace:
        need = np.ones(n_i, dtype=bool)
    else:
        need = np.zeros(n_i, dtype=bool)
        if gpar.impute:
            need |= miss
        if gpar.replace:
            need |= avail
    if need.any():
        ridx = eng.to_device(np.flatnonzero(need), torch.int64)
        Xq = xd if need.all() else xd.copy_rows(ridx, int(need.sum()))
        est = eng.empty(Xq.n)
        fac.mean_at(Xq.t, Xq.ld, Xq.n, est)
        eng.scatter_col(col, 1, 0, None if need.all() else ridx, est, Xq.n)
    return fac, block, xd.with_col(col), zd_next


def _zd(gpar, x_ind, p):
    if isinstance(x_ind, DevMat):
        return DevMat(x_ind.eng, x_ind.t, x_ind.n, x_ind.d, x_ind.ld, frozen=True)
    return DevMat.from_host(gpar.engine, x_ind, spare=p + 1)


def logpdf_sparse(gpar, x, y, w, only_last_layer, return_inputs, x_ind, outputs, sample_missing=False,
                  normals=None, grad_out=None):
    eng = gpar.engine
    if not isinstance(y, dict):
        y = np.asarray(y, dtype=np.float64)
        w = np.asarray(w, dtype=np.float64)
    p = len(gpar.layers)
    xd = gpar._as_devmat(x, spare=p + 1)
    zd = _zd(gpar, gpar.x_ind if x_ind is None else x_ind, p)
    slots = []
    normals = list(normals) if normals is not None else None
    for is_last, ((y_i, w_i, mask), model) in last(
        zip(per_output(y, w, keep=gpar.impute or sample_missing), gpar.layers), select=outputs
    ):
        xd = xd.take_rows(mask)
        layer = _layer(model)
        prior = None
        if layer.block is not None:
            # the layer is a sparse posterior (model | data): the bound is taken under it (regression.py:495-499)
            if sample_missing:
                raise NotImplementedError("sample_missing under a sparse posterior is not supported")
            blk = layer.block
            if blk.factor is None:
                blk.factor = SparseFactor(eng, layer.spec, blk.Z, blk.X, blk.y_host, blk.sig_host)
            prior = blk.factor
        fac, _, xd, zd = _train_step(gpar, layer, xd, zd, y_i, w_i, is_last, prior=prior,
                                     sample_missing=sample_missing, normals=normals)
        if (not only_last_layer) or is_last:
            slots.append((fac.elbo_slot(), fac.n))
            if grad_out is not None and is_last and fac.n > 0:
                avail = ~np.isnan(y_i[:, 0])
                grad_out["raw"] = fac.elbo_grad_raw(w_i[avail])
                grad_out["layer"] = layer
    eng.check_infos()
    if return_inputs:
        return xd, zd
    total = 0.0
    if slots:
        vals = np.stack([s.cpu().numpy() for s, _ in slots])  # (device -> host reads of 4 doubles per layer)
        for (t0, t1, t2, _), (_, n) in zip(vals, slots):
            if n > 0:
                total += -0.5 * (t0 + t1 - t2)
    return float(total)


def condition_sparse(gpar, out, x, y, w):
    y = np.asarray(y, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    p = y.shape[1]
    xd = gpar._as_devmat(x, spare=p + 1)
    zd = _zd(gpar, gpar.x_ind, p)
    for is_last, ((y_i, w_i, mask), model) in last(zip(per_output(y, w, keep=gpar.impute), gpar.layers)):
        xd = xd.take_rows(mask)
        layer = _layer(model)
        _, block, xd, zd = _train_step(gpar, layer, xd, zd, y_i, w_i, is_last)
        out.layers.append(construct_model(layer.conditioned(block), layer.noise))
    gpar.engine.check_infos()
    return out


def sample_sparse(gpar, x, w, latent, num_samples, normals, train, return_device, generator=None):
    """model.py:245-277 with sparse posteriors; same contract as :meth:`GPAR.sample`."""
    eng = gpar.engine
    S = int(num_samples)
    p = len(gpar.layers)
    w = np.asarray(w, dtype=np.float64)
    xs = gpar._as_devmat(x, spare=p + 1)
    ns = xs.n
    if normals is None:
        Zall = torch.randn(S, p, ns, dtype=F64, device=eng.device, generator=generator)
        Z2all = torch.randn(S, p, ns, dtype=F64, device=eng.device, generator=generator) if latent else None
    else:
        Zall = eng.to_device(np.asarray(normals["Z"], dtype=np.float64).reshape(S, p, ns))
        Z2all = eng.to_device(np.asarray(normals["Z2"], dtype=np.float64).reshape(S, p, ns)) if latent else None
    out = eng.empty(max(S * ns * p, 1))
    shared, xs_all = True, None
    train_iter = None
    if train is not None:
        xt, yt, wt = train
        yt = np.asarray(yt, dtype=np.float64)
        wt = np.asarray(wt, dtype=np.float64)
        xd = gpar._as_devmat(xt, spare=p + 1)
        zd = _zd(gpar, gpar.x_ind, p)
        train_iter = per_output(yt, wt, keep=gpar.impute)
    for i, (is_last, model) in enumerate(last(gpar.layers)):
        layer = _layer(model)
        noise = layer.noise
        sd = eng.to_device(np.sqrt(noise / w[:, i])) if latent else None
        d_s = eng.zeros(max(ns, 1)) if latent else eng.to_device(noise / w[:, i])
        Zi = Zall[:, i, :].contiguous()
        Z2i = Z2all[:, i, :].contiguous() if latent else None
        fac = None
        if train_iter is not None:
            y_i, w_i, mask = next(train_iter)
            xd = xd.take_rows(mask)
            fac, _, xd, zd = _train_step(gpar, layer, xd, zd, y_i, w_i, is_last)
        elif layer.block is not None:
            blk = layer.block
            if blk.factor is None:
                blk.factor = SparseFactor(eng, layer.spec, blk.Z, blk.X, blk.y_host, blk.sig_host)
            fac = blk.factor
        if fac is None:
            # prior layer: dense joint draw (no observations to sparsify)
            if shared:
                pf = Factor(eng, layer.spec, xs.t, xs.ld, d_s, eng.zeros(1), 0, ns)
                f_col = eng.empty(max(S * ns, 1))
                pf.ext_sample(Zi, S, f_col)
                y_col = f_col
                if latent:
                    y_col = eng.empty(max(S * ns, 1))
                    pf.ext_sample(Zi, S, y_col, sd=sd, Z2=Z2i)
                mean = eng.zeros(max(ns, 1))
            else:
                f_col, y_col = gpar._chains_layer(layer, None, xs_all, S, ns, d_s, sd, Zi, Z2i, latent)
                mean = gpar._chain_means
        elif shared:
            f_col, y_col, mean = fac.sample_rows(xs, d_s, Zi, S, 1, ns, sd=sd, Z2=Z2i)
        else:
            f_col, y_col, mean = fac.sample_rows(xs_all, d_s, Zi, 1, S, ns, sd=sd, Z2=Z2i)
        eng.scatter_col(out, p, i, None, f_col, S * ns)
        if not is_last:
            if gpar.replace:
                if shared:
                    xs = xs.with_col(mean)
                else:
                    xs_all = xs_all.with_col(mean)
            else:
                if shared:
                    rep = np.tile(np.arange(ns, dtype=np.int64), S)
                    xs_all = xs.copy_rows(eng.to_device(rep, torch.int64), S * ns, spare=p + 1)
                    shared = False
                xs_all = xs_all.with_col(y_col)
    eng.check_infos()
    res = out.reshape(S, ns, p)
    return res if return_device else res.cpu().numpy()
