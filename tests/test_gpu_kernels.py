"""Kernel-level parity (through the C ABI) against numpy/scipy restatements.  fp64
tolerances follow SURVEY.md 8(d): Gram entries |d| <= 1e-14 (x magnitude), Cholesky factor
relative Frobenius error <= 1e-12 (1 + log n), solves / products rel <= 1e-10."""
import math

import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import gpar_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from gpar_b200.engine import Engine

    return Engine()


def dev(eng, a):
    return eng.to_device(np.ascontiguousarray(a, dtype=np.float64))


def spd(n, rng, cond_noise=0.1, d=2):
    x = rng.uniform(0, 1, (n, d))
    K = O.kernel_matrix([dict(type="eq", variance=1.0, cols=list(range(d)), scales=[0.25] * d)], x, x)
    return K + cond_noise * np.eye(n)


TERMS_ALL = [
    dict(type="eq", variance=1.3, cols=[0, 1], scales=[0.25, 0.5]),
    dict(type="periodic", variance=0.7, cols=[0, 1], scales=[1.0, 1.5, 0.8, 1.2], periods=[1.0, 0.5], decays=[10.0, 5.0]),
    dict(type="linear", variance=1.0, cols=[0, 1], scales=[3.0, 2.0]),
    dict(type="const", variance=0.4),
    dict(type="linear", variance=1.0, cols=[2, 3], scales=[10.0, 7.0]),
    dict(type="rq", variance=0.9, cols=[2, 3], scales=[1.0, 2.0], alpha=0.5),
]
TERMS_EQ = [dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25, 0.25])]


@pytest.mark.parametrize("terms,d", [(TERMS_EQ, 2), (TERMS_ALL, 4)])
@pytest.mark.parametrize("n", [1, 63, 64, 200, 333])
def test_gram_symmetric(eng, terms, d, n):
    from gpar_b200.spec import lower_terms

    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 2, (n, d + 1))  # one unused column: exercises ldx > d and select
    diag = rng.uniform(0.1, 1.0, n)
    ref = O.kernel_matrix(terms, x, x) + np.diag(diag + 1e-12)
    ld = n + (n & 1) + 2
    out = eng.zeros(n * ld)
    eng.gram(lower_terms(terms), dev(eng, x).reshape(-1), d + 1, n, out, ld, diag=dev(eng, diag), lower_only=True)
    got = out.cpu().numpy().reshape(n, ld)[:, :n]
    scale = np.abs(ref).max()
    assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= 4e-14 * scale
    # tiles strictly above the diagonal are skipped (never written)
    if n > 128:
        assert np.all(got[:64, 128:] == 0)


@pytest.mark.parametrize("nx,ny", [(5, 7), (130, 64), (64, 257)])
def test_gram_cross(eng, nx, ny):
    from gpar_b200.spec import lower_terms

    rng = np.random.default_rng(nx * 1000 + ny)
    x, y = rng.uniform(0, 1, (nx, 4)), rng.uniform(0, 1, (ny, 6))
    ref = O.kernel_matrix(TERMS_ALL, x, y)
    out = eng.zeros(nx * ny)
    eng.gram(lower_terms(TERMS_ALL), dev(eng, x).reshape(-1), 4, nx, out, ny, Y=dev(eng, y).reshape(-1), ldy=6, ny=ny,
             lower_only=False)
    got = out.cpu().numpy().reshape(nx, ny)
    assert np.max(np.abs(got - ref)) <= 4e-14 * np.abs(ref).max()


@pytest.mark.parametrize("n,nb", [(1, 0), (31, 1), (128, 3), (129, 0), (257, 0), (300, 130), (385, 1), (777, 1), (1024, 64),
                                  (1153, 70), (2049, 1), (3500, 1)])
def test_potrf_with_appended_rows(eng, n, nb):
    rng = np.random.default_rng(n + nb)
    A = spd(n, rng)
    B = rng.standard_normal((nb, n))
    ld = n + (n & 1)
    Ap = np.zeros((n, ld)); Ap[:, :n] = np.tril(A) + np.triu(np.full((n, n), np.nan), 1)  # upper must never be read
    Bp = np.zeros((max(nb, 1), ld)); Bp[:nb, :n] = B
    Ad, Bd = dev(eng, Ap).reshape(-1), dev(eng, Bp).reshape(-1)
    ws, info = eng.potrf(Ad, ld, n, B=Bd if nb else None, ldb=ld, nb=nb)
    assert int(info.cpu()[0]) == 0
    L = np.tril(Ad.cpu().numpy().reshape(n, ld)[:, :n])
    Lref = sla.cholesky(A, lower=True)
    tol = 1e-12 * (1 + math.log(max(n, 2)))
    assert np.linalg.norm(L - Lref) / np.linalg.norm(Lref) <= tol
    if nb:
        got = Bd.cpu().numpy().reshape(-1, ld)[:nb, :n]
        ref = sla.solve_triangular(Lref, B.T, lower=True).T
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-10
    # workspace = inverses of the diagonal 128-blocks
    nt = (n + 127) // 128
    W = ws.cpu().numpy()[: nt * 128 * 128].reshape(nt, 128, 128)
    for k in range(nt):
        kb = min(128, n - 128 * k)
        blk = Lref[128 * k : 128 * k + kb, 128 * k : 128 * k + kb]
        assert np.max(np.abs(np.tril(W[k, :kb, :kb]) @ blk - np.eye(kb))) <= 1e-9


def test_potrf_reports_first_bad_pivot(eng):
    n = 200
    A = spd(n, np.random.default_rng(0))
    A[150, 150] = -1.0
    Ad = dev(eng, np.tril(A)).reshape(-1)
    _, info = eng.potrf(Ad, n, n)
    assert int(info.cpu()[0]) == 151
    from gpar_b200._lib import GparError

    with pytest.raises(GparError):
        eng.check_infos()


def test_potrf_near_singular_matches_lapack(eng):
    """Joint [obs; dense grid] EQ matrix with only the 1e-12 jitter on the grid block (the latent
    posterior covariance of SURVEY hard part 4): LAPACK factors it; so must we, with a backward
    error at rounding level (refined triangular solves)."""
    n = 200
    x = np.linspace(0, 1, n)[:, None]
    X = np.vstack([x[::8], x])
    d = np.concatenate([np.full(25, 0.1), np.zeros(n)]) + 1e-12
    K = O.kernel_matrix([dict(type="eq", variance=1.0, cols=[0], scales=[0.1])], X, X) + np.diag(d)
    N = K.shape[0]
    ld = N + (N & 1)
    Ap = np.zeros((N, ld)); Ap[:, :N] = np.tril(K)
    Ad = dev(eng, Ap).reshape(-1)
    _, info = eng.potrf(Ad, ld, N)
    assert int(info.cpu()[0]) == 0
    L = np.tril(Ad.cpu().numpy().reshape(N, ld)[:, :N])
    Lref = sla.cholesky(K, lower=True)
    assert np.abs(L @ L.T - K).max() <= 20 * np.abs(Lref @ Lref.T - K).max() + 1e-15


def test_potrf_near_singular_interior_tiles(eng):
    """Ill-conditioned diagonal tiles in the interior of the sweep (rows 256..639 are a tight cluster with noise
    1e-9: kappa_inf(L_kk) >> 1e3 for tile 2; the Schur complements of tiles 3, 4 are noise-dominated): the pipelined HEAD
    task of column 3 finds the refine flag set after its four blocks, pushes its newest k-tile to global memory and
    falls back to the refined solve
    (potrf.cu head_chain).  Backward error at LAPACK's level."""
    rng = np.random.default_rng(3)
    n = 1300
    X = rng.uniform(0, 1, (n, 2))
    d = rng.uniform(0.05, 0.2, n)
    X[256:640] = X[256] + 2e-3 * rng.uniform(-1, 1, (384, 2))
    d[256:640] = 1e-9
    K = O.kernel_matrix([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25, 0.25])], X, X) + np.diag(d)
    ld = n
    Ap = np.zeros((n, ld)); Ap[:, :n] = np.tril(K)
    Ad = dev(eng, Ap).reshape(-1)
    ws, info = eng.potrf(Ad, ld, n)
    assert int(info.cpu()[0]) == 0
    nt = (n + 127) // 128
    flags = ws.cpu().numpy()[nt * 128 * 128 : nt * 128 * 128 + nt]
    assert flags[2] != 0 and not flags[:2].any()  # the refined path really ran (HEAD of column 3), the plain one too
    L = np.tril(Ad.cpu().numpy().reshape(n, ld)[:, :n])
    Lref = sla.cholesky(K, lower=True)
    assert np.abs(L @ L.T - K).max() <= 20 * np.abs(Lref @ Lref.T - K).max() + 1e-15


def test_potrf_batched_multi_tile(eng):
    """Batched factorisation of multi-tile matrices (panel flags are per matrix and per diagonal tile)."""
    rng = np.random.default_rng(8)
    n, batch = 700, 3
    mats = [spd(n, rng, cond_noise=0.05 * (b + 1)) for b in range(batch)]
    Ad = dev(eng, np.stack([np.tril(a) for a in mats])).reshape(-1)
    _, info = eng.potrf(Ad, n, n, batch=batch, strideA=n * n)
    assert np.all(info.cpu().numpy() == 0)
    got = Ad.cpu().numpy().reshape(batch, n, n)
    for b in range(batch):
        Lref = sla.cholesky(mats[b], lower=True)
        assert np.linalg.norm(np.tril(got[b]) - Lref) / np.linalg.norm(Lref) <= 1e-11


def test_potrf_batched(eng):
    rng = np.random.default_rng(5)
    n, batch = 200, 5
    mats = [spd(n, rng, cond_noise=0.05 * (b + 1)) for b in range(batch)]
    Ad = dev(eng, np.stack([np.tril(a) for a in mats])).reshape(-1)
    _, info = eng.potrf(Ad, n, n, batch=batch, strideA=n * n)
    assert np.all(info.cpu().numpy() == 0)
    got = Ad.cpu().numpy().reshape(batch, n, n)
    for b in range(batch):
        Lref = sla.cholesky(mats[b], lower=True)
        assert np.linalg.norm(np.tril(got[b]) - Lref) / np.linalg.norm(Lref) <= 1e-11


# nb beyond one wave of 128-row blocks (148 SMs): the rows left after the whole waves run as 32 / 64 / 96-row
# blocks (trsm_row_plan): 100 -> 3 x 32 + 4; 27904 -> one wave + 140 x 64; 31761 -> one wave + 133 x 96 + 49
@pytest.mark.parametrize("n,nb", [(100, 5), (300, 260), (513, 129), (300, 100), (200, 27904), (130, 31761),
                                  (129, 148 * 128)])
def test_trsm_rows_and_backsolve(eng, n, nb):
    rng = np.random.default_rng(n * 7 + nb)
    A = spd(n, rng)
    ld = n + (n & 1)
    Ap = np.zeros((n, ld)); Ap[:, :n] = np.tril(A)
    Ad = dev(eng, Ap).reshape(-1)
    u = rng.standard_normal(n)
    ud = dev(eng, np.concatenate([u, np.zeros(ld - n)]))
    ws, info = eng.potrf(Ad, ld, n)
    Lref = sla.cholesky(A, lower=True)
    B = rng.standard_normal((nb, n))
    Bp = np.zeros((nb, ld)); Bp[:, :n] = B
    Bd = dev(eng, Bp).reshape(-1)
    eng.trsm_rows(Ad, ld, n, ws, Bd, ld, nb)
    got = Bd.cpu().numpy().reshape(nb, ld)[:, :n]
    ref = sla.solve_triangular(Lref, B.T, lower=True).T
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-10
    alpha = eng.backsolve(Ad, ld, n, ws, ud).cpu().numpy()[:n]
    ref = sla.solve_triangular(Lref.T, u, lower=False)
    assert np.linalg.norm(alpha - ref) / np.linalg.norm(ref) <= 1e-10
    out2 = eng.zeros(2)
    eng.logdet_quad(Ad, ld, n, ud, out2)
    ld_ref, q_ref = 2 * np.sum(np.log(np.diag(Lref))), float(u @ u)
    got2 = out2.cpu().numpy()
    assert abs(got2[0] - ld_ref) <= 1e-11 * max(1, abs(ld_ref)) and abs(got2[1] - q_ref) <= 1e-12 * q_ref


def test_trsm_rows_refined_tail_blocks(eng):
    """Ill-conditioned diagonal tiles (refine flags set) with a tail of 32-row blocks: every block has its own
    scratch tile for the refinement step; backward error at rounding level, rows bit-identical whatever block
    they fall into (the same rows solved alone and inside a larger batch)."""
    rng = np.random.default_rng(77)
    n, nb = 300, 200
    A = spd(n, rng, cond_noise=1e-9)
    Ad = dev(eng, np.tril(A)).reshape(-1)
    ws, info = eng.potrf(Ad, n, n)
    assert int(info.cpu()[0]) == 0
    assert ws.cpu().numpy()[3 * 128 * 128:3 * 128 * 128 + 3].any()  # refine flags set
    L = np.tril(Ad.cpu().numpy().reshape(n, n))
    B = rng.standard_normal((nb, n))
    Bd = dev(eng, B).reshape(-1)
    eng.trsm_rows(Ad, n, n, ws, Bd, n, nb)
    X = Bd.cpu().numpy().reshape(nb, n)
    assert np.linalg.norm(X @ L.T - B) / (np.linalg.norm(X) * np.linalg.norm(L)) <= 1e-13
    big = np.concatenate([rng.standard_normal((148 * 128 + 40, n)), B[:70]])  # same rows, other block heights
    Gd = dev(eng, big).reshape(-1)
    eng.trsm_rows(Ad, n, n, ws, Gd, n, big.shape[0])
    assert np.array_equal(Gd.cpu().numpy().reshape(-1, n)[-70:], X[:70])


@pytest.mark.parametrize("n,k,batch", [(64, 10, 1), (300, 257, 1), (130, 520, 3)])
def test_syrk_sub(eng, n, k, batch):
    rng = np.random.default_rng(n + k)
    ldc, ldw = n + (n & 1), k + (k & 1)
    Cm = rng.standard_normal((batch, n, ldc))
    Wm = rng.standard_normal((batch, n, ldw)); Wm[:, :, k:] = np.nan  # padding must not be read
    Cd, Wd = dev(eng, Cm).reshape(-1), dev(eng, Wm).reshape(-1)
    eng.syrk_sub(Cd, ldc, n, Wd, ldw, k, batch=batch, strideC=n * ldc, strideW=n * ldw)
    got = Cd.cpu().numpy().reshape(batch, n, ldc)
    for b in range(batch):
        W = Wm[b, :, :k]
        ref = Cm[b, :, :n] - W @ W.T
        g = got[b, :, :n]
        assert np.max(np.abs(np.tril(g) - np.tril(ref))) <= 1e-11 * np.abs(ref).max()
        assert np.array_equal(np.triu(g, 1), np.triu(Cm[b, :, :n], 1))  # strictly upper untouched


def test_gemv_gram_gemv_sample_affine_gather(eng):
    from gpar_b200.spec import lower_terms

    rng = np.random.default_rng(1)
    m, n = 77, 301
    A, x = rng.standard_normal((m, n)), rng.standard_normal(n)
    y = eng.zeros(m)
    eng.gemv(dev(eng, A).reshape(-1), n, m, n, dev(eng, x), y)
    np.testing.assert_allclose(y.cpu().numpy(), A @ x, rtol=1e-12, atol=1e-12)

    xq, xa, v = rng.uniform(0, 1, (37, 4)), rng.uniform(0, 1, (530, 4)), rng.standard_normal(530)
    out = eng.zeros(37)
    eng.gram_gemv(lower_terms(TERMS_ALL), dev(eng, xq).reshape(-1), 4, 37, dev(eng, xa).reshape(-1), 4, 530, dev(eng, v), out)
    ref = O.kernel_matrix(TERMS_ALL, xq, xa) @ v
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-11, atol=1e-11)

    nn, ns, batch = 150, 21, 3
    Cm = rng.standard_normal((batch, nn, nn)); mean = rng.standard_normal((batch, nn)); sd = rng.uniform(size=(batch, nn))
    Z, Z2 = rng.standard_normal((batch, ns, nn)), rng.standard_normal((batch, ns, nn))
    outd = eng.zeros(batch * ns * nn)
    eng.sample_affine(dev(eng, Cm).reshape(-1), nn, nn, dev(eng, Z).reshape(-1), outd, ns, batch=batch, strideC=nn * nn,
                      mean=dev(eng, mean).reshape(-1), sd=dev(eng, sd).reshape(-1), Z2=dev(eng, Z2).reshape(-1))
    ref = mean[:, None, :] + np.einsum("bij,bsj->bsi", np.tril(Cm), Z) + sd[:, None, :] * Z2
    np.testing.assert_allclose(outd.cpu().numpy().reshape(batch, ns, nn), ref, rtol=1e-12, atol=1e-12)

    src = rng.standard_normal((50, 6)); idx = rng.integers(0, 50, 33)
    dst = eng.zeros(33 * 8)
    eng.gather_rows(dev(eng, src).reshape(-1), 6, eng.to_device(idx, torch.int64), 33, 5, dst, 8)
    assert np.array_equal(dst.cpu().numpy().reshape(33, 8)[:, :5], src[idx, :5])
    col = rng.standard_normal(33)
    perm = rng.permutation(33)
    eng.scatter_col(dst, 8, 7, eng.to_device(perm, torch.int64), dev(eng, col), 33)
    got = dst.cpu().numpy().reshape(33, 8)[:, 7]
    assert np.array_equal(got[perm], col)
    inp = rng.standard_normal((9, 40)); o = eng.zeros(40)
    eng.mean_axis0(dev(eng, inp).reshape(-1), 9, 40, o)
    np.testing.assert_allclose(o.cpu().numpy(), inp.mean(axis=0), rtol=1e-14, atol=1e-15)


def test_factor_joint_identities(eng):
    """Joint factor [obs; ext]: u, alpha, identity mean, W u and chol of the Schur complement."""
    from gpar_b200.engine import Factor
    from gpar_b200.spec import lower_terms

    rng = np.random.default_rng(3)
    n_o, n_e, d = 333, 150, 3
    terms = [dict(type="eq", variance=1.0, cols=[0, 1, 2], scales=[0.3, 0.3, 0.5])]
    X = rng.uniform(0, 1, (n_o + n_e, d)); dvec = rng.uniform(0.05, 0.2, n_o + n_e); y = rng.standard_normal(n_o)
    fac = Factor(eng, lower_terms(terms), dev(eng, np.hstack([X, np.zeros((n_o + n_e, 1))])).reshape(-1), 4, dev(eng, dvec),
                 dev(eng, y), n_o, n_e)
    assert int(fac.info.cpu()[0]) == 0
    K = O.kernel_matrix(terms, X, X) + np.diag(dvec + 1e-12)
    Koo, Keo, Kee = K[:n_o, :n_o], K[n_o:, :n_o], K[n_o:, n_o:]
    L = sla.cholesky(Koo, lower=True)
    u = sla.solve_triangular(L, y, lower=True)
    np.testing.assert_allclose(fac.u.cpu().numpy()[:n_o], u, rtol=1e-9, atol=1e-11)
    alpha = np.linalg.solve(Koo, y)
    np.testing.assert_allclose(fac.alpha().cpu().numpy()[:n_o], alpha, rtol=1e-8, atol=1e-10)
    m = eng.zeros(n_o); fac.mean_obs(m, 0, n_o)
    np.testing.assert_allclose(m.cpu().numpy(), (Koo - np.diag(dvec[:n_o] + 1e-12)) @ alpha, rtol=1e-8, atol=1e-9)
    me = eng.zeros(n_e); fac.ext_mean(me)
    np.testing.assert_allclose(me.cpu().numpy(), Keo @ alpha, rtol=1e-8, atol=1e-10)
    Cref = sla.cholesky(Kee - Keo @ np.linalg.solve(Koo, Keo.T), lower=True)
    J = fac.J.cpu().numpy().reshape(fac.n, fac.ld)
    Cgot = np.tril(J[n_o:, n_o : n_o + n_e])
    assert np.linalg.norm(Cgot - Cref) / np.linalg.norm(Cref) <= 1e-9
    out2 = eng.zeros(2); fac.logdet_quad(out2, 0, 0, n_o)
    lp = -0.5 * (out2.cpu().numpy()[0] + n_o * math.log(2 * math.pi) + out2.cpu().numpy()[1])
    from scipy.stats import multivariate_normal

    assert abs(lp - multivariate_normal(np.zeros(n_o), Koo).logpdf(y)) <= 1e-8 * abs(lp)
