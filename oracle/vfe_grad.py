"""CPU oracle for the gradient of the VFE (Titsias) bound -- TEST INFRASTRUCTURE ONLY.

The reference differentiates the ELBO of ``PseudoObs`` (gpar/model.py:286-287, regression.py:434-459) with
torch autograd.  The device path will need the explicit weights below (planned for the next round, DESIGN.md
section 7 item 2); this module states them in numpy and tests/test_oracle_relations.py checks them against
finite differences, so the kernels can be validated against a pinned formula.

With K_zz + eps I = L_z L_z^T, B = L_z^-1 K_zx, Sigma = diag(sigma), A = I + B Sigma^-1 B^T, c = B Sigma^-1 y:

    ELBO = -1/2 [ sum_j (k_jj - |b_j|^2)/sigma_j + sum_j log(2 pi sigma_j) + logdet A + y^T Sigma^-1 y - c^T A^-1 c ]

    d ELBO = sum_{m,j} G_zx[m, j] dK(z_m, x_j) + sum_{m,m'} G_zz[m, m'] dK(z_m, z_m') + sum_j g_kk[j] dk(x_j, x_j)
             + sum_j g_sigma[j] dsigma_j

    beta    = Sigma^-1 y - Sigma^-1 B^T A^-1 c                       (= (Q + Sigma)^-1 y,  Q = B^T B)
    T       = K_zz^-1 K_zx = L_z^-T B
    G_zx    = (T beta) beta^T + L_z^-T (I - A^-1) B Sigma^-1         (M x n)
    G_zz    = -1/2 G_zx T^T                                          (M x M; symmetric part is what matters)
    g_kk    = -1/2 / sigma
    g_sigma = 1/2 (beta^2 - P_jj) + 1/2 (k_jj - |b_j|^2) / sigma^2,  P_jj = 1/sigma_j - |L_A^-1 b_j|^2 / sigma_j^2
"""
import numpy as np
import scipy.linalg as sla

__all__ = ["vfe_elbo", "vfe_elbo_weights"]


def _parts(Kzz, Kzx, sigma, y, eps):
    M = Kzz.shape[0]
    Lz = sla.cholesky(Kzz + eps * np.eye(M), lower=True)
    B = sla.solve_triangular(Lz, Kzx, lower=True)
    A = np.eye(M) + (B / sigma) @ B.T
    LA = sla.cholesky(A, lower=True)
    c = B @ (y / sigma)
    return Lz, B, A, LA, c


def vfe_elbo(Kzz, Kzx, kdiag, sigma, y, eps=1e-12):
    Lz, B, A, LA, c = _parts(Kzz, Kzx, sigma, y, eps)
    v = sla.solve_triangular(LA, c, lower=True)
    t0 = np.sum((kdiag - np.sum(B * B, axis=0)) / sigma) + np.sum(np.log(2 * np.pi * sigma)) + np.sum(y * y / sigma)
    return -0.5 * (t0 + 2 * np.sum(np.log(np.diag(LA))) - v @ v)


def vfe_elbo_weights(Kzz, Kzx, kdiag, sigma, y, eps=1e-12):
    """(G_zx, G_zz, g_kk, g_sigma) of the module docstring."""
    Lz, B, A, LA, c = _parts(Kzz, Kzx, sigma, y, eps)
    M = Kzz.shape[0]
    Ainv_c = sla.cho_solve((LA, True), c)
    beta = y / sigma - (B.T @ Ainv_c) / sigma
    T = sla.solve_triangular(Lz.T, B, lower=False)
    H = np.eye(M) - sla.cho_solve((LA, True), np.eye(M))
    G_zx = np.outer(T @ beta, beta) + sla.solve_triangular(Lz.T, H @ B, lower=False) / sigma
    G_zz = -0.5 * G_zx @ T.T
    g_kk = -0.5 / sigma
    V = sla.solve_triangular(LA, B, lower=True)
    P_jj = 1.0 / sigma - np.sum(V * V, axis=0) / sigma ** 2
    g_sigma = 0.5 * (beta ** 2 - P_jj) + 0.5 * (kdiag - np.sum(B * B, axis=0)) / sigma ** 2
    return G_zx, G_zz, g_kk, g_sigma
