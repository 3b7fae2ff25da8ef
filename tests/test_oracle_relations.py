"""Anchors the oracle's floating-point path on every relational known answer of
the reference's tests (reference tests/test_model.py:118-293,
tests/test_regression.py:92-208) and on independent scipy / textbook
identities.  The reference holds no stored fp vectors: parity is unpinned
beyond these relations (see oracle/gpar_oracle.py header)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose
from scipy.stats import multivariate_normal

from oracle import gpar_oracle as O

EQ = [dict(type="eq", variance=1.0, cols=[0], scales=[1.0])]


def eq_terms(d):
    return [dict(type="eq", variance=1.0, cols=list(range(d)), scales=[1.0] * d)]


def lin_terms(d):
    return [dict(type="linear", variance=1.0, cols=list(range(d)), scales=[1.0] * d)]


@pytest.fixture(params=[1, 2])
def x(request):
    return np.random.default_rng(request.param).standard_normal((10, request.param))


@pytest.fixture()
def w():
    return np.random.default_rng(7).uniform(size=(10, 2)) + 1e-2


def test_kernel_against_scipy_logpdf(x):
    # dense logpdf == scipy multivariate_normal.
    d = x.shape[1]
    f = O.GP(eq_terms(d))
    noise = 0.1 / (np.random.default_rng(3).uniform(size=10) + 1e-2)
    y = np.random.default_rng(4).standard_normal(10)
    K = O.kernel_matrix(eq_terms(d), x, x) + np.diag(noise) + O.EPSILON * np.eye(10)
    ref = multivariate_normal(mean=np.zeros(10), cov=K).logpdf(y)
    assert_allclose(f(x, noise).logpdf(y), ref, rtol=0, atol=1e-10)


def test_obs_dense_and_sparse(x):
    # reference tests/test_model.py:118-149
    d = x.shape[1]
    f = O.GP(eq_terms(d))
    noise = 0.1
    rng = np.random.default_rng(5)
    w = rng.uniform(size=10) + 1e-2
    y = f(x, 0.1).sample(O.Normals(rng=rng))
    y_missing = y.copy()
    y_missing[::2] = np.nan
    expect = f(x[1::2], noise / w[1::2]).logpdf(y[1::2])

    obs = O.GPAR()._obs(x, None, y_missing, w, f, noise)
    assert isinstance(obs, O.Obs)
    assert_allclose(f.logpdf(obs), expect, atol=1e-6)

    obs = O.GPAR(x_ind=x)._obs(x, x, y_missing, w, f, noise)
    assert isinstance(obs, O.PseudoObs)
    assert_allclose(f.logpdf(obs), expect, atol=1e-6)


def test_update_inputs_known_answers():
    # reference tests/test_model.py:152-218
    f = O.GP(EQ)
    x = np.array([[1.0], [2.0], [3.0]])
    y = np.array([[4.0], [5.0], [6.0]])
    res = np.concatenate([x, y], axis=1)
    x_ind = np.array([[6.0], [7.0]])
    res_ind = np.array([[6.0, 0], [7.0, 0]])

    def check(got, want):
        for g, h in zip(got, want):
            assert_allclose(g, h, rtol=1e-7, atol=1e-9)

    check(O.GPAR(x_ind=x_ind)._update_inputs(x, x_ind, y, f, None), (res, res_ind))

    this_y = y.copy(); this_y[1] = np.nan
    this_res = res.copy(); this_res[1, 1] = 0
    check(O.GPAR(impute=True, x_ind=x_ind)._update_inputs(x, x_ind, this_y, f, None), (this_res, res_ind))

    this_res = res.copy(); this_res[0, 1] = 0; this_res[1, 1] = np.nan; this_res[2, 1] = 0
    check(O.GPAR(replace=True, x_ind=x_ind)._update_inputs(x, x_ind, this_y, f, None), (this_res, res_ind))

    this_res = res.copy(); this_res[:, 1] = 0
    check(O.GPAR(impute=True, replace=True, x_ind=x_ind)._update_inputs(x, x_ind, y, f, None), (this_res, res_ind))

    obs = O.Obs(f(np.array([1.0, 2, 3, 6, 7])), np.array([9.0, 10, 11, 12, 13]))
    res_ind = np.array([[6.0, 12], [7.0, 13]])

    this_res = res.copy(); this_res[1, 1] = 10
    check(O.GPAR(impute=True, x_ind=x_ind)._update_inputs(x, x_ind, this_y, f, obs), (this_res, res_ind))

    this_res = res.copy(); this_res[0, 1] = 9; this_res[1, 1] = np.nan; this_res[2, 1] = 11
    check(O.GPAR(replace=True, x_ind=x_ind)._update_inputs(x, x_ind, this_y, f, obs), (this_res, res_ind))

    this_res = res.copy(); this_res[0, 1] = 9; this_res[1, 1] = 10; this_res[2, 1] = 11
    check(O.GPAR(impute=True, replace=True, x_ind=x_ind)._update_inputs(x, x_ind, y, f, obs), (this_res, res_ind))


def test_conditioning(x, w):
    # reference tests/test_model.py:221-241
    d = x.shape[1]
    f1, noise1 = O.GP(eq_terms(d)), 1e-10
    f2, noise2 = O.GP(eq_terms(d)), 2e-10
    gpar = O.GPAR().add_layer(lambda: (f1, noise1)).add_layer(lambda: (f2, noise2))
    nrm = O.Normals(rng=np.random.default_rng(0))
    y = np.concatenate([f1(x, noise1).sample(nrm), f2(x, noise2).sample(nrm)], axis=1)
    gpar = gpar | (x, y, w)
    f1_post, n1 = gpar.layers[0]()
    f2_post, n2 = gpar.layers[1]()
    assert n1 == noise1 and n2 == noise2
    assert_allclose(f1_post.mean(x), y[:, 0:1], atol=1e-3)
    # Note: the reference's second layer kernel is EQ over *all* d+1 columns.


def test_logpdf_additivity_and_resume(x, w):
    # reference tests/test_model.py:244-272
    d = x.shape[1]
    f1, noise1 = O.GP(eq_terms(d)), 2e-1
    f2, noise2 = O.GP(lin_terms(d + 1)), 1e-1
    gpar = O.GPAR().add_layer(lambda: (f1, noise1)).add_layer(lambda: (f2, noise2))
    y = gpar.sample(x, w, latent=True, normals=O.Normals(rng=np.random.default_rng(1)))
    x2 = np.concatenate([x, y[:, 0:1]], axis=1)
    logpdf1 = f1(x, noise1 / w[:, 0]).logpdf(y[:, 0])
    logpdf2 = f2(x2, noise2 / w[:, 1]).logpdf(y[:, 1])
    assert gpar.logpdf(x, y, w) == logpdf1 + logpdf2
    assert gpar.logpdf(x, y, w, only_last_layer=True) == logpdf2
    x_partial, x_ind_partial = gpar.logpdf(x, y, w, return_inputs=True, outputs=[0])
    assert gpar.logpdf(x_partial, y, w, x_ind=x_ind_partial, outputs=[1]) == logpdf2
    y[1, 0] = np.nan
    a = gpar.logpdf(x, y, w, sample_missing=True, normals=O.Normals(rng=np.random.default_rng(2)))
    b = gpar.logpdf(x, y, w, sample_missing=True, normals=O.Normals(rng=np.random.default_rng(3)))
    assert abs(a - b) > 1e-6


def test_sample_posterior_reproduces_data(x, w):
    # reference tests/test_model.py:275-293
    d = x.shape[1]
    f1, noise1 = O.GP(eq_terms(d)), 1e-10
    f2, noise2 = O.GP(eq_terms(d + 1)), 2e-10
    gpar = O.GPAR().add_layer(lambda: (f1, noise1)).add_layer(lambda: (f2, noise2))
    nrm = O.Normals(rng=np.random.default_rng(11))
    y = gpar.sample(x, w, latent=True, normals=nrm)
    post = gpar | (x, y, w)
    assert_allclose(post.sample(x, w, normals=nrm), y, atol=1e-3)
    assert_allclose(post.sample(x, w, latent=True, normals=nrm), y, atol=1e-3)


def test_vfe_against_textbook_inverse():
    # mu, posterior mean and covariance vs inv-based Titsias forms.
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 3, (12, 1)); z = rng.uniform(0, 3, (5, 1)); xs = rng.uniform(0, 3, (7, 1))
    y = np.sin(x) + 0.1 * rng.standard_normal((12, 1))
    sig = 0.1 / (rng.uniform(size=12) + 0.5)
    f = O.GP(EQ)
    obs = O.PseudoObs(f(z), f(x, sig), y)
    post = f | obs
    Kzz = O.kernel_matrix(EQ, z, z) + O.EPSILON * np.eye(5)
    Kzx = O.kernel_matrix(EQ, z, x); Ksz = O.kernel_matrix(EQ, xs, z)
    Su = np.linalg.inv(Kzz + Kzx @ np.diag(1 / sig) @ Kzx.T)
    mean = Ksz @ Su @ Kzx @ (y / sig[:, None])
    cov = O.kernel_matrix(EQ, xs, xs) - Ksz @ np.linalg.inv(Kzz) @ Ksz.T + Ksz @ Su @ Ksz.T
    assert_allclose(post.mean(xs), mean, rtol=1e-6, atol=1e-8)
    assert_allclose(post.kernel(xs, xs), cov, rtol=1e-6, atol=1e-8)
    assert_allclose(post.kernel_diag(xs), np.diag(cov), rtol=1e-6, atol=1e-8)
    # ELBO <= exact log marginal, and tight at z = x.
    exact = f(x, sig).logpdf(y)
    assert f.logpdf(obs) <= exact + 1e-9
    assert_allclose(f.logpdf(O.PseudoObs(f(x), f(x, sig), y)), exact, atol=1e-6)


def test_regressor_logpdf_prior_and_posterior():
    # reference tests/test_regression.py:92-137 (x of shape (10, 2), with weights)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((10, 2)); w = rng.uniform(size=(10, 2)) + 1
    reg = O.OracleRegressor(replace=False, impute=False, nonlinear=True, nonlinear_scale=0.1, linear=True,
                            linear_scale=10.0, noise=1e-2, normalise_y=False)
    y = reg.sample(x, w, p=2, latent=True, normals=O.Normals(rng=rng))
    gpar = reg._construct_gpar(2, 2)
    f1, noise1 = gpar.layers[0](); f2, noise2 = gpar.layers[1]()
    noise1 = noise1 / w[:, 0]; noise2 = noise2 / w[:, 1]
    x2 = np.concatenate([x, y[:, 0:1]], axis=1)
    lp = f1(x, noise1).logpdf(y[:, 0]) + f2(x2, noise2).logpdf(y[:, 1])
    assert_allclose(reg.logpdf(x, y, w), lp, atol=1e-6)
    f1_post = f1 | O.Obs(f1(x, noise1), y[:, 0]); f2_post = f2 | O.Obs(f2(x2, noise2), y[:, 1])
    lp = f1_post(x, noise1).logpdf(y[:, 0]) + f2_post(x2, noise2).logpdf(y[:, 1])
    with pytest.raises(RuntimeError):
        reg.logpdf(x, y, w, posterior=True)
    reg.condition(x, y, w)
    assert_allclose(reg.logpdf(x, y, w, posterior=True), lp, atol=1e-6)


def test_regressor_sample_and_predict():
    # reference tests/test_regression.py:161-208
    rng = np.random.default_rng(1)
    x = rng.standard_normal((10, 1)); w = rng.uniform(size=(10, 2)) + 1
    reg = O.OracleRegressor(replace=False, impute=False, linear=True, linear_scale=1.0, nonlinear=False,
                            noise=1e-8, normalise_y=False, transform_y=O.squishing_transform)
    with pytest.raises(ValueError):
        reg.sample(x, w)
    with pytest.raises(RuntimeError):
        reg.sample(x, w, posterior=True)
    nrm = O.Normals(rng=rng)
    assert isinstance(reg.sample(x, w, p=2, normals=nrm), np.ndarray)
    assert isinstance(reg.sample(x, w, p=2, num_samples=2, normals=nrm), list)
    y = reg.sample(x, w, p=2, normals=nrm)
    reg.condition(x, y, w)
    assert_allclose(y, np.mean(reg.sample(x, w, posterior=True, num_samples=100, normals=nrm), axis=0), atol=5e-2)
    assert_allclose(y, reg.predict(x, w, num_samples=100, latent=True, normals=nrm), atol=5e-2)
    _, lo, up = reg.predict(x, w, num_samples=100, credible_bounds=True, normals=nrm)
    assert_allclose(up, lo, atol=5e-2)


def test_regressor_condition_normalisation():
    # reference tests/test_regression.py:211-243 (normalisation part)
    rng = np.random.default_rng(2)
    x = rng.standard_normal((10, 2))
    reg = O.OracleRegressor(replace=False, impute=False, normalise_y=True, transform_y=O.squishing_transform)
    y = reg.sample(x, None, p=2, normals=O.Normals(rng=rng))
    reg.condition(x, y)
    assert_allclose(np.mean(reg.y, axis=0), 0, atol=1e-12)
    assert_allclose(np.std(reg.y, axis=0), 1, rtol=1e-7)
    yp = y.copy(); yp[:, 0] = 1
    reg.condition(x, yp)
    assert not np.any(np.isnan(reg.y))


def test_features_kernel_family_runs():
    # reference tests/test_regression.py:246-265 (kernel options; no fit)
    reg = O.OracleRegressor(replace=True, scale=1.0, per=True, per_period=1.0, per_decay=10.0, input_linear=True,
                            input_linear_scale=0.1, linear=True, linear_scale=1.0, nonlinear=True,
                            nonlinear_scale=1.0, rq=True, noise=0.1)
    x = np.stack([np.linspace(0, 10, 20), np.linspace(10, 20, 20)], axis=1)
    y = reg.sample(x, p=2, normals=O.Normals(rng=np.random.default_rng(0)))
    assert y.shape == (20, 2) and np.all(np.isfinite(y))
    assert np.isfinite(reg.logpdf(x, y))
    reg2 = O.OracleRegressor(scale_tie=True)
    reg2.sample(x, p=2, normals=O.Normals(rng=np.random.default_rng(0)))
    vs = reg2.get_variables()
    assert "0/input/scales" in vs and "1/input/scales" not in vs


def test_vfe_gradient_weights_match_finite_differences():
    """oracle/vfe_grad.py: d ELBO along random directions of (K_zz, K_zx, k_jj, sigma) equals the weighted sums."""
    from oracle.vfe_grad import vfe_elbo, vfe_elbo_weights

    rng = np.random.default_rng(3)
    M, n = 7, 25
    z = rng.uniform(0, 1, (M, 2)); x = rng.uniform(0, 1, (n, 2))
    terms = [dict(type="eq", variance=1.2, cols=[0, 1], scales=[0.4, 0.6])]
    Kzz, Kzx = O.kernel_matrix(terms, z, z), O.kernel_matrix(terms, z, x)
    kdiag = np.full(n, 1.2)
    sigma = rng.uniform(0.05, 0.3, n)
    y = rng.standard_normal(n)
    G_zx, G_zz, g_kk, g_sigma = vfe_elbo_weights(Kzz, Kzx, kdiag, sigma, y)
    for trial in range(4):
        dzz = rng.standard_normal((M, M)); dzz = dzz + dzz.T
        dzx = rng.standard_normal((M, n))
        dkk = rng.standard_normal(n)
        dsg = rng.standard_normal(n) * 0.01
        h = 1e-6
        f1 = vfe_elbo(Kzz + h * dzz, Kzx + h * dzx, kdiag + h * dkk, sigma + h * dsg, y)
        f0 = vfe_elbo(Kzz - h * dzz, Kzx - h * dzx, kdiag - h * dkk, sigma - h * dsg, y)
        fd = (f1 - f0) / (2 * h)
        an = np.sum(G_zx * dzx) + np.sum(G_zz * dzz) + g_kk @ dkk + g_sigma @ dsg
        assert abs(fd - an) <= 1e-5 * max(1.0, abs(fd)), (trial, fd, an)
