"""Inducing-point (VFE) path on the GPU vs the oracle's PseudoObs restatement (SURVEY 8a row a9;
reference tests/test_model.py:118-149 sparse part, BASELINE configs[3] scaled down).  ELBO rel <= 1e-7."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import bench
from oracle import gpar_oracle as O
from tests.test_gpu_model import both

pytestmark = pytest.mark.gpu

LAYERS = [
    ([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.3, 0.4])], 0.05),
    ([dict(type="eq", variance=0.8, cols=[0, 1], scales=[0.3, 0.4]),
      dict(type="linear", variance=1.0, cols=[2], scales=[5.0]),
      dict(type="eq", variance=0.6, cols=[2], scales=[1.0])], 0.08),
    ([dict(type="eq", variance=1.1, cols=[0, 1], scales=[0.5, 0.2]),
      dict(type="linear", variance=1.0, cols=[2, 3], scales=[5.0, 3.0]),
      dict(type="rq", variance=0.6, cols=[2, 3], scales=[1.0, 2.0], alpha=0.7)], 0.1),
]


def test_sparse_obs_tight_at_z_equals_x():
    # reference tests/test_model.py:139-149: with x_ind = x the bound equals the exact log-marginal
    rng = np.random.default_rng(5)
    x = rng.standard_normal((10, 2)); wv = rng.uniform(size=(10, 1)) + 1e-2
    terms = [dict(type="eq", variance=1.0, cols=[0, 1], scales=[1.0, 1.0])]
    y = O.GP(terms)(x, 0.1).sample(O.Normals(rng=rng))
    ym = y.copy(); ym[::2] = np.nan
    g, _ = both([(terms, 0.1)], x_ind=x)
    expect = O.GP(terms)(x[1::2], 0.1 / wv[1::2, 0]).logpdf(y[1::2])
    assert_allclose(g.logpdf(x, ym, wv), expect, atol=1e-6)


@pytest.mark.parametrize("replace,impute", [(False, False), (True, True), (False, True)])
def test_sparse_chain_vs_oracle(replace, impute):
    rng = np.random.default_rng(7)
    n, ns, m, p, S, M = 150, 23, 2, 3, 3, 20
    x = rng.uniform(0, 1, (n, m)); xs = rng.uniform(0, 1, (ns, m)); z = rng.uniform(0, 1, (M, m))
    g, o = both(LAYERS, replace=replace, impute=impute, x_ind=z)
    w = rng.uniform(0.5, 2.0, (n, p)); ws = rng.uniform(0.5, 2.0, (ns, p))
    gen = O.GPAR()
    for t, nz in LAYERS:
        gen = gen.add_layer(lambda t=t, nz=nz: (O.GP(t), nz))
    y = gen.sample(x, w, normals=O.Normals(rng=rng))
    y[rng.uniform(size=(n, p)) < 0.15] = np.nan
    a, b = g.logpdf(x, y, w), o.logpdf(x, y, w)
    assert abs(a - b) <= 1e-7 * abs(b)
    for latent in (False, True):
        Z = rng.standard_normal((S, p, ns)); Z2 = rng.standard_normal((S, p, ns))
        queue = []
        for s in range(S):
            for i in range(p):
                queue.append(Z[s, i])
                if latent:
                    queue.append(Z2[s, i])
        opost = o | (x, y, w)
        nrm = O.Normals(queue=queue)
        ref = np.stack([opost.sample(xs, ws, latent=latent, normals=nrm) for _ in range(S)])
        got = g.sample(xs, ws, latent=latent, num_samples=S, normals={"Z": Z, "Z2": Z2}, train=(x, y, w))
        assert_allclose(got, ref, rtol=1e-6, atol=1e-7)
        got2 = (g | (x, y, w)).sample(xs, ws, latent=latent, num_samples=S, normals={"Z": Z, "Z2": Z2})
        assert_allclose(got2, ref, rtol=1e-6, atol=1e-7)


def test_regressor_c4_small():
    """BASELINE configs[3] (inducing points, linear + nonlinear, replace + impute) at oracle size."""
    from gpar_b200 import GPARRegressor

    data = bench.make_data(n=500, m=2, p=4, ns=60, S=3, missing=0.1)
    z = np.random.default_rng(4).uniform(0, 1, (48, 2))
    kw = dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
              replace=True, impute=True, normalise_y=True, x_ind=z)
    reg, ora = GPARRegressor(**kw), O.OracleRegressor(**kw)
    reg.condition(data["x"], data["y"]); ora.condition(data["x"], data["y"])
    a, b = reg.logpdf(data["x"], data["y"]), ora.logpdf(data["x"], data["y"])
    assert abs(a - b) <= 1e-7 * abs(b)
    S, p = 3, 4
    mean = reg.predict(data["xs"], num_samples=S, normals={"Z": data["Z"]})
    queue = [data["Z"][s, i] for s in range(S) for i in range(p)]
    ref = ora.predict(data["xs"], num_samples=S, normals=O.Normals(queue=queue))
    assert np.max(np.abs(mean - ref)) <= 1e-5 * np.max(np.abs(ref))
    assert reg.x_ind.ndim == 2


def test_regressor_sparse_split_k_syrk():
    """n large enough that A = I + B Sigma^-1 B^T is formed by the split-K SYRK (n // 2048 >= 2 slices,
    ragged last slice): ELBO against the oracle, rel <= 1e-7."""
    from gpar_b200 import GPARRegressor

    data = bench.make_data(n=4500, m=2, p=2, ns=20, S=1, missing=0.05)
    z = np.random.default_rng(4).uniform(0, 1, (40, 2))
    kw = dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
              replace=True, impute=True, normalise_y=True, x_ind=z)
    reg, ora = GPARRegressor(**kw), O.OracleRegressor(**kw)
    reg.condition(data["x"], data["y"]); ora.condition(data["x"], data["y"])
    a, b = reg.logpdf(data["x"], data["y"]), ora.logpdf(data["x"], data["y"])
    assert abs(a - b) <= 1e-7 * abs(b)


def test_sparse_logpdf_sample_missing_vs_oracle():
    """logpdf(sample_missing=True) with inducing points (model.py:229-237 on PseudoObs): the missing rows of
    every non-final output are filled with a joint draw from the sparse posterior; same injected normals as
    the oracle => same value (rel 1e-6), other normals => another value."""
    rng = np.random.default_rng(11)
    n = 60
    x = rng.uniform(0, 1, (n, 2)); wv = rng.uniform(size=(n, 3)) + 0.5
    y = rng.standard_normal((n, 3))
    y[rng.uniform(size=n) < 0.25, 0] = np.nan
    y[rng.uniform(size=n) < 0.25, 1] = np.nan
    z_ind = rng.uniform(0, 1, (12, 2))
    for replace in (False, True):
        g, o = both(LAYERS, x_ind=z_ind, replace=replace, impute=False)
        nm = [int(np.isnan(y[:, 0]).sum())]
        zs = [rng.standard_normal(n) for _ in range(2)]
        # per_output(keep=True) drops nothing here (every row has some later observation or is kept): the
        # oracle consumes one normal vector per layer with missing rows, sized by its own missing count
        a = g.logpdf(x, y, wv, sample_missing=True, normals=[z.copy() for z in zs_for(g, o, x, y, wv, zs)])
        b = o.logpdf(x, y, wv, sample_missing=True, normals=O.Normals(queue=[z.copy() for z in zs_for(g, o, x, y, wv, zs)]))
        assert abs(a - b) <= 1e-6 * abs(b), (replace, a, b)


def zs_for(g, o, x, y, w, zs):
    """Normal vectors cut to the number of missing rows each layer will ask for (layers 0 and 1)."""
    from oracle.gpar_oracle import per_output

    out = []
    for li, (y_i, w_i, mask) in enumerate(per_output(y, w, keep=True)):
        if li >= 2:
            break
        out.append(zs[li][: int(np.isnan(y_i[:, 0]).sum())])
    return [z for z in out if len(z) > 0]


def test_regressor_c4_multi_tile_inducing_points():
    """BASELINE configs[3] with its real number of inducing points (M = 512: four Cholesky tiles of K_zz and of
    A = I + B Sigma^-1 B^T, multi-tile back-substitutions, split-K SYRK over ~4200 observed rows in two slices, the second ragged) at a
    size the oracle still handles: ELBO rel <= 1e-7, predictive means rel <= 1e-5, inducing inputs of the next
    layer (x_ind propagation, model.py:304-305) rel <= 1e-6."""
    from gpar_b200 import GPARRegressor

    data = bench.make_data(n=4700, m=2, p=3, ns=100, S=2, missing=0.1)
    z = np.random.default_rng(4).uniform(0, 1, (512, 2))
    kw = dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
              replace=True, impute=True, normalise_y=True, x_ind=z)
    reg, ora = GPARRegressor(**kw), O.OracleRegressor(**kw)
    reg.condition(data["x"], data["y"]); ora.condition(data["x"], data["y"])
    a, b = reg.logpdf(data["x"], data["y"]), ora.logpdf(data["x"], data["y"])
    assert abs(a - b) <= 1e-7 * abs(b), (a, b)
    S, p = 2, 3
    mean = reg.predict(data["xs"], num_samples=S, normals={"Z": data["Z"]})
    queue = [data["Z"][s, i] for s in range(S) for i in range(p)]
    ref = ora.predict(data["xs"], num_samples=S, normals=O.Normals(queue=queue))
    assert np.max(np.abs(mean - ref)) <= 1e-5 * np.max(np.abs(ref))
    # the chain's inputs after the last layer: training rows [x, est_1, est_2] and inducing rows
    from gpar_b200.regression import _construct_gpar

    g = _construct_gpar(reg, reg.vs, reg.m, reg.p)
    xd, zd = g.logpdf(reg.x, reg.y, reg.w, return_inputs=True)
    og = ora._construct_gpar(ora.m, ora.p)
    xo, zo = og.logpdf(ora.x, ora.y, ora.w, return_inputs=True)
    assert_allclose(zd.to_host(), zo, rtol=1e-6, atol=1e-8)
    assert_allclose(xd.to_host(), xo, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("replace,impute", [(False, True), (True, True)])
def test_logpdf_under_sparse_posterior_vs_oracle(replace, impute):
    """``GPARRegressor.logpdf(posterior=True)`` with inducing points (regression.py:495-499 -> model.py:148-176,
    286-287): every layer is a sparse posterior and the new bound is taken under its mean and kernel
    (k - P P^T + Q Q^T); the chain's inputs / inducing inputs are extended with posterior-of-posterior means.
    Against the oracle's nested PseudoObs, rel <= 1e-6 (the posterior kernel at the inducing points is formed by
    cancellation, cond ~ 1e6)."""
    from gpar_b200 import GPARRegressor

    data = bench.make_data(n=400, m=2, p=3, ns=10, S=1, missing=0.1)
    new = bench.make_data(n=150, m=2, p=3, ns=10, S=1, missing=0.15, seed=50)
    z = np.random.default_rng(4).uniform(0, 1, (24, 2))
    kw = dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
              replace=replace, impute=impute, normalise_y=True, x_ind=z)
    reg, ora = GPARRegressor(**kw), O.OracleRegressor(**kw)
    reg.condition(data["x"], data["y"]); ora.condition(data["x"], data["y"])
    a = reg.logpdf(new["x"], new["y"], posterior=True)
    b = ora.logpdf(new["x"], new["y"], posterior=True)
    assert abs(a - b) <= 1e-6 * abs(b), (a, b)
    # and it differs from the prior bound (the conditioning matters)
    assert abs(a - reg.logpdf(new["x"], new["y"])) > 1e-3 * abs(a)
