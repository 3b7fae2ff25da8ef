"""Gradients of the dense log-marginal for `fit` (reference: torch autograd through Gram +
Cholesky + solve, regression.py:434-459).  CPU part: the chain rule of gpar_b200/spec.py against finite
differences of the oracle's kernel matrices, with the device kernel's raw sums restated in numpy.  GPU part:
gpar_potri / gpar_gram_grad through the C ABI against the same numpy restatement, and fit() end to end."""
import math

import numpy as np
import pytest
import scipy.linalg as sla

from gpar_b200 import _lib
from gpar_b200.spec import lower_terms, named_gradients, term_gradients
from oracle import gpar_oracle as O

TERMS = {
    "eq": [dict(type="eq", variance=1.3, cols=[0, 1], scales=np.array([0.4, 0.7]), names=dict(variance="v", scales="s"))],
    "rq": [dict(type="rq", variance=0.8, cols=[0, 1], scales=np.array([0.5, 0.9]), alpha=0.7,
                names=dict(variance="v", scales="s", alpha="a"))],
    "lin+const": [dict(type="linear", variance=1.0, cols=[0, 1], scales=np.array([2.0, 3.0]), names=dict(scales="ls")),
                  dict(type="const", variance=0.6, names=dict(variance="c")),
                  dict(type="eq", variance=1.0, cols=[1], scales=np.array([0.5]), names=dict(variance="v", scales="s"))],
    "periodic": [dict(type="periodic", variance=0.9, cols=[0, 1], scales=np.array([1.0, 1.4, 0.8, 1.2]),
                      periods=np.array([0.7, 1.3]), decays=np.array([3.0, 5.0]),
                      names=dict(variance="pv", scales="ps", periods="pp", decays="pd")),
                 dict(type="eq", variance=0.5, cols=[0], scales=np.array([0.3]), names=dict(variance="v", scales="s"))],
    "tied": [dict(type="eq", variance=1.0, cols=[0, 1], scales=np.array([0.4, 0.6]), names=dict(variance="v", scales="s")),
             dict(type="eq", variance=0.7, cols=[0, 1], scales=np.array([0.4, 0.6]), names=dict(variance="v2", scales="s"))],
}


def features(spec, X):
    """phi (n, F) and psi = d phi / d b (n, F) of the lowered spec, in numpy."""
    F = spec.n_feats
    phi, psi = np.zeros((X.shape[0], F)), np.zeros((X.shape[0], F))
    for f in range(F):
        x = X[:, spec.feat_col[f]]
        a, b, op = spec.feat_a[f], spec.feat_b[f], spec.feat_op[f]
        if op == _lib.FEAT_SCALE:
            phi[:, f] = a * x
        elif op == _lib.FEAT_SIN:
            phi[:, f], psi[:, f] = a * np.sin(b * x), a * x * np.cos(b * x)
        else:
            phi[:, f], psi[:, f] = a * np.cos(b * x), -a * x * np.sin(b * x)
    return phi, psi


def raw_sums_numpy(spec, X, alpha, Ainv, dvec):
    """What gram_grad_kernel + grad_reduce_kernel produce (layout of include/gpar_b200.h)."""
    W = 0.5 * (np.outer(alpha, alpha) - Ainv)
    phi, psi = features(spec, X)
    raw = np.zeros(_lib.GRAD_NP)
    base = 2 * _lib.MAX_TERMS
    for t in range(spec.n_terms):
        T = spec.terms[t]
        if T.type == _lib.TERM_CONST:
            raw[2 * t] = W.sum()
            continue
        fs = range(T.f_begin, T.f_end)
        if T.type == _lib.TERM_LINEAR:
            for f in fs:
                s1 = (W * np.outer(phi[:, f], phi[:, f])).sum()
                raw[2 * t] += s1
                raw[base + 2 * f] = T.variance * s1
            continue
        d = phi[:, None, T.f_begin:T.f_end] - phi[None, :, T.f_begin:T.f_end]
        r2 = (d ** 2).sum(-1)
        if T.type == _lib.TERM_EQ:
            e = np.exp(-0.5 * r2)
            raw[2 * t] = (W * e).sum()
            g = T.variance * e
        else:
            u = r2 / (2 * T.alpha)
            b = (1 + u) ** (-T.alpha)
            raw[2 * t] = (W * b).sum()
            raw[2 * t + 1] = (W * T.variance * b * (u / (1 + u) - np.log1p(u))).sum()
            g = T.variance * b / (1 + u)
        for k, f in enumerate(fs):
            raw[base + 2 * f] = (W * g * d[:, :, k] ** 2).sum()
            raw[base + 2 * f + 1] = (W * g * d[:, :, k] * (psi[:, None, f] - psi[None, :, f])).sum()
    raw[-1] = (np.diag(W) * dvec).sum()
    return raw


def lml(terms, X, y, noise, w):
    K = O.kernel_matrix(terms, X, X) + np.diag(noise / w + 1e-12)
    L = sla.cholesky(K, lower=True)
    u = sla.solve_triangular(L, y, lower=True)
    return -0.5 * (2 * np.log(np.diag(L)).sum() + len(y) * math.log(2 * math.pi) + u @ u)


def problem(terms, n=40, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, (n, 2))
    y = rng.standard_normal(n)
    w = rng.uniform(0.5, 2.0, n)
    noise = 0.3
    K = O.kernel_matrix(terms, X, X) + np.diag(noise / w + 1e-12)
    Ainv = np.linalg.inv(K)
    return X, y, w, noise, Ainv, Ainv @ y


def fd_named(terms, X, y, noise, w, eps=1e-6):
    """Central finite differences of the numpy log-marginal w.r.t. every named variable."""
    import copy

    names = {}
    for t in terms:
        for field, nm in t.get("names", {}).items():
            names.setdefault(nm, []).append(field)
    out = {}
    for nm in names:
        ref = next(np.atleast_1d(np.asarray(t[f], dtype=float)) for t in terms for f, n2 in t.get("names", {}).items() if n2 == nm)
        g = np.zeros(ref.size)
        for k in range(ref.size):
            vals = []
            for sgn in (1, -1):
                tt = copy.deepcopy(terms)
                for t in tt:
                    for f, n2 in t.get("names", {}).items():
                        if n2 == nm:
                            v = np.atleast_1d(np.asarray(t[f], dtype=float)).copy()
                            v[k] += sgn * eps
                            t[f] = v if np.ndim(t[f]) else float(v[0])
                vals.append(lml(tt, X, y, noise, w))
            g[k] = (vals[0] - vals[1]) / (2 * eps)
        out[nm] = g
    out["noise"] = np.array([(lml(terms, X, y, noise + eps, w) - lml(terms, X, y, noise - eps, w)) / (2 * eps)])
    return out


@pytest.mark.parametrize("which", sorted(TERMS))
def test_chain_rule_matches_finite_differences(which):
    terms = TERMS[which]
    X, y, w, noise, Ainv, alpha = problem(terms)
    raw = raw_sums_numpy(lower_terms(terms), X, alpha, Ainv, 1.0 / w)
    got = named_gradients(terms, raw, noise_name="noise")
    ref = fd_named(terms, X, y, noise, w)
    assert set(got) == set(ref)
    for nm in ref:
        np.testing.assert_allclose(np.atleast_1d(got[nm]), ref[nm], rtol=2e-5, atol=1e-6, err_msg=nm)
    assert len(term_gradients(terms, raw)) == len(terms)


def test_latent_gradient_matches_transform():
    from gpar_b200.spec import Vars

    vs = Vars()
    vs.bnd("a", np.array([0.5, 2.0]))
    vs.get("b", 1.5)
    vs.bnd("c", 0.3, lower=1e-8)
    names = ["a", "b", "c"]
    f = lambda: float(np.sum(np.sin(vs["a"])) + vs["b"] ** 2 + np.log(vs["c"]))
    grads = {"a": np.cos(vs["a"]), "b": 2 * vs["b"], "c": 1 / vs["c"]}
    gz = vs.latent_gradient(names, grads)
    z0 = vs.get_latent_vector(names)
    fd = np.zeros_like(z0)
    for k in range(z0.size):
        for sgn in (1, -1):
            z = z0.copy(); z[k] += sgn * 1e-6
            vs.set_latent_vector(names, z)
            fd[k] += sgn * f() / 2e-6
    vs.set_latent_vector(names, z0)
    np.testing.assert_allclose(gz, fd, rtol=1e-6, atol=1e-9)


# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng():
    from gpar_b200.engine import Engine

    return Engine()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [50, 128, 300, 700])
def test_potri_matches_inverse(eng, n):
    rng = np.random.default_rng(n)
    X = rng.uniform(0, 1, (n, 2))
    K = O.kernel_matrix(TERMS["eq"], X, X) + 0.1 * np.eye(n)
    ld = n + (n & 1)
    Ap = np.zeros((n, ld)); Ap[:, :n] = np.tril(K)
    Ad = eng.to_device(Ap).reshape(-1)
    ws, info = eng.potrf(Ad, ld, n)
    Ainv = eng.potri(Ad, ld, n, ws).cpu().numpy().reshape(n, ld)[:, :n]
    ref = np.linalg.inv(K)
    assert np.abs(np.tril(Ainv) - np.tril(ref)).max() <= 1e-9 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("which", sorted(TERMS))
@pytest.mark.parametrize("n", [40, 200])
def test_gram_grad_matches_numpy(eng, which, n):
    terms = TERMS[which]
    X, y, w, noise, Ainv, alpha = problem(terms, n=n, seed=n)
    spec = lower_terms(terms)
    ld = n + (n & 1)
    Ap = np.zeros((n, ld)); Ap[:, :n] = np.tril(Ainv)
    raw = eng.gram_grad(spec, eng.to_device(X).reshape(-1), 2, n, eng.to_device(alpha), eng.to_device(Ap).reshape(-1),
                        ld, eng.to_device(1.0 / w)).cpu().numpy()
    ref = raw_sums_numpy(spec, X, alpha, Ainv, 1.0 / w)
    scale = np.abs(ref).max()
    assert np.abs(raw - ref).max() <= 1e-10 * scale


@pytest.mark.gpu
def test_fit_analytic_gradient_matches_fd_and_improves(eng):
    import bench
    from gpar_b200 import GPARRegressor
    from gpar_b200.regression import _construct_gpar

    data = bench.make_data(n=150, m=2, p=2, ns=10, S=1, missing=0.1)
    kw = dict(scale=0.5, noise=0.2, linear=True, linear_scale=5.0, nonlinear=True, nonlinear_scale=1.0, input_linear=True,
              replace=True, impute=True, normalise_y=True)
    reg = GPARRegressor(engine=eng, **kw)
    reg.condition(data["x"], data["y"])
    lp0 = reg.logpdf(data["x"], data["y"])
    # gradient of the layer-1 objective at the initial point: device vs central differences of the device logpdf
    from gpar_b200.model import per_output
    from gpar_b200.spec import named_gradients as ng

    y_cached = {k: list(per_output(reg.y, reg.w, keep=k)) for k in [True, False]}
    pi = 1
    gp = _construct_gpar(reg, reg.vs, reg.m, pi + 1)
    fx, fxi = gp.logpdf(reg.x, y_cached, None, only_last_layer=True, outputs=list(range(pi)), return_inputs=True)
    for ctor in _construct_gpar(reg, reg.vs, reg.m, pi + 1).layers:
        ctor()
    names = reg.vs.match([f"{pi}/*"])
    z0 = reg.vs.get_latent_vector(names)

    def val_grad(z, want_grad):
        reg.vs.set_latent_vector(names, z)
        g = {} if want_grad else None
        v = _construct_gpar(reg, reg.vs, reg.m, pi + 1).logpdf(fx, y_cached, None, only_last_layer=True, outputs=[pi],
                                                             x_ind=fxi, grad_out=g)
        if not want_grad:
            return v
        return v, reg.vs.latent_gradient(names, ng(g["layer"].terms, g["raw"].cpu().numpy(), noise_name=f"{pi}/noise"))

    v, gz = val_grad(z0, True)
    fd = np.zeros_like(z0)
    for k in range(z0.size):
        zp, zm = z0.copy(), z0.copy()
        zp[k] += 1e-5; zm[k] -= 1e-5
        fd[k] = (val_grad(zp, False) - val_grad(zm, False)) / 2e-5
    reg.vs.set_latent_vector(names, z0)
    np.testing.assert_allclose(gz, fd, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(fd).max()))
    # fit end to end: the objective goes up
    reg.fit(data["x"], data["y"], iters=15)
    lp1 = reg.logpdf(data["x"], data["y"])
    assert lp1 > lp0


@pytest.mark.gpu
@pytest.mark.parametrize("M,n", [(24, 150), (160, 700)])
def test_vfe_gradient_matches_finite_differences(eng, M, n):
    """Inducing-point layers of ``fit`` (regression.py:434-459 through PseudoObs): the device gradient of the VFE
    bound (SparseFactor.elbo_grad_raw: gpar_gram_wgrad over K_xz and K_zz + host diagonal terms; weights of
    oracle/vfe_grad.py) against central differences of the device bound, layer 1 of a two-layer model
    (EQ + input-linear + output-linear + output-EQ kernel, per-row weights).  M = 160 spans two Cholesky tiles."""
    import bench
    from gpar_b200 import GPARRegressor
    from gpar_b200.model import per_output
    from gpar_b200.regression import _construct_gpar
    from gpar_b200.spec import named_gradients as ng

    data = bench.make_data(n=n, m=2, p=2, ns=10, S=1, missing=0.1)
    z = np.random.default_rng(4).uniform(0, 1, (M, 2))
    kw = dict(scale=0.5, noise=0.2, linear=True, linear_scale=5.0, nonlinear=True, nonlinear_scale=1.0, input_linear=True,
              replace=True, impute=True, normalise_y=True, x_ind=z)
    reg = GPARRegressor(engine=eng, **kw)
    w = np.random.default_rng(9).uniform(0.5, 2.0, data["y"].shape)
    reg.condition(data["x"], data["y"], w)
    y_cached = {k: list(per_output(reg.y, reg.w, keep=k)) for k in [True, False]}
    pi = 1
    gp = _construct_gpar(reg, reg.vs, reg.m, pi + 1)
    fx, fxi = gp.logpdf(reg.x, y_cached, None, only_last_layer=True, outputs=list(range(pi)), return_inputs=True)
    for ctor in _construct_gpar(reg, reg.vs, reg.m, pi + 1).layers:
        ctor()
    names = reg.vs.match([f"{pi}/*"])
    z0 = reg.vs.get_latent_vector(names)

    def val_grad(zv, want_grad):
        reg.vs.set_latent_vector(names, zv)
        g = {} if want_grad else None
        v = _construct_gpar(reg, reg.vs, reg.m, pi + 1).logpdf(fx, y_cached, None, only_last_layer=True, outputs=[pi],
                                                             x_ind=fxi, grad_out=g)
        if not want_grad:
            return v
        return v, reg.vs.latent_gradient(names, ng(g["layer"].terms, np.asarray(g["raw"]), noise_name=f"{pi}/noise"))

    v, gz = val_grad(z0, True)
    fd = np.zeros_like(z0)
    for k in range(z0.size):
        zp, zm = z0.copy(), z0.copy()
        zp[k] += 1e-5; zm[k] -= 1e-5
        fd[k] = (val_grad(zp, False) - val_grad(zm, False)) / 2e-5
    reg.vs.set_latent_vector(names, z0)
    np.testing.assert_allclose(gz, fd, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(fd).max()))


@pytest.mark.gpu
def test_fit_with_inducing_points_and_joint_objective_use_analytic_gradients(eng, monkeypatch):
    """fit(x_ind=...) and fit(fix=False) on data-only inputs run L-BFGS with jac=True (no finite differences):
    the number of bound evaluations stays at the optimiser's own count, and the objective improves."""
    import bench
    import gpar_b200.regression as R
    from gpar_b200 import GPARRegressor

    calls = []
    real_minimize = R.minimize

    def spy(fun, x0, jac=None, **kw):
        calls.append(bool(jac))
        return real_minimize(fun, x0, jac=jac, **kw)

    monkeypatch.setattr(R, "minimize", spy)
    data = bench.make_data(n=300, m=2, p=2, ns=10, S=1, missing=0.1)
    z = np.random.default_rng(4).uniform(0, 1, (40, 2))
    kw = dict(scale=0.5, noise=0.2, linear=True, linear_scale=5.0, nonlinear=True, nonlinear_scale=1.0, replace=True,
              impute=True, normalise_y=True)
    reg = GPARRegressor(engine=eng, x_ind=z, **kw)
    reg.condition(data["x"], data["y"])
    lp0 = reg.logpdf(data["x"], data["y"])
    reg.fit(data["x"], data["y"], iters=10)
    assert calls == [True, True] and reg.logpdf(data["x"], data["y"]) > lp0
    # joint objective, no replace / no missing data: inputs are data => analytic for all layers at once
    calls.clear()
    full = bench.make_data(n=200, m=2, p=3, ns=10, S=1, missing=0.0)
    reg2 = GPARRegressor(engine=eng, **{**kw, "replace": False, "impute": True, "scale_tie": True})
    reg2.condition(full["x"], full["y"])
    lp0 = reg2.logpdf(full["x"], full["y"])
    reg2.fit(full["x"], full["y"], fix=False, iters=8)
    assert calls == [True, True, True] and reg2.logpdf(full["x"], full["y"]) > lp0
    # joint objective with replace: posterior means feed later layers => finite differences (stated in fit)
    calls.clear()
    reg3 = GPARRegressor(engine=eng, **kw)
    reg3.fit(data["x"], data["y"], fix=False, iters=2)
    assert calls == [False, False]
