"""oracle/torch_ref.py (the reference op sequence on torch tensors: the timed CPU baseline and, on CUDA, the
cuSOLVER/cuBLAS secondary bar of bench.py) against the pinned numpy oracle, on CPU."""
import numpy as np
import pytest

import bench
from oracle.gpar_oracle import Normals, OracleRegressor
from oracle.torch_ref import TorchNormals, TorchRegressor, kernel_matrix, timed_step
from oracle import gpar_oracle as O

import torch

KW = [
    dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0, markov=2,
         replace=True, impute=True, normalise_y=True),
    dict(scale=0.25, noise=0.1, linear=False, nonlinear=True, nonlinear_scale=1.0, replace=False, impute=False,
         normalise_y=True),
    dict(scale=0.5, noise=0.05, linear=True, nonlinear=True, rq=True, per=True, input_linear=True, replace=False,
         impute=True, normalise_y=False),
]


def test_kernel_matrix_matches_oracle():
    rng = np.random.default_rng(0)
    X, Y = rng.uniform(0, 1, (40, 3)), rng.uniform(0, 1, (31, 3))
    terms = [dict(type="eq", variance=1.3, cols=[0, 1], scales=[0.3, 0.4]),
             dict(type="rq", variance=0.7, cols=[2], scales=[0.5], alpha=0.8),
             dict(type="linear", variance=1.0, cols=[0, 2], scales=[2.0, 3.0]),
             dict(type="const", variance=0.4),
             dict(type="periodic", variance=0.9, cols=[0, 1], scales=[1.0, 1.1, 1.2, 1.3], periods=[0.7, 0.9],
                  decays=[5.0, 6.0])]
    K = kernel_matrix(terms, torch.as_tensor(X), torch.as_tensor(Y)).numpy()
    np.testing.assert_allclose(K, O.kernel_matrix(terms, X, Y), rtol=0, atol=5e-13)


@pytest.mark.parametrize("kw", KW)
def test_logpdf_and_predict_match_oracle(kw):
    missing = 0.1 if kw["impute"] else 0.0
    cfg = dict(n=120, m=2, p=3, ns=30, S=3, missing=missing)
    data = bench.make_data(**cfg)
    ora, ref = OracleRegressor(**kw), TorchRegressor(**kw)
    ora.condition(data["x"], data["y"])
    ref.condition(data["x"], data["y"])
    lp0, lp1 = ora.logpdf(data["x"], data["y"]), ref.logpdf(data["x"], data["y"])
    assert abs(lp0 - lp1) <= 1e-9 * abs(lp0)
    q = [data["Z"][s, i] for s in range(cfg["S"]) for i in range(cfg["p"])]
    m0 = ora.predict(data["xs"], num_samples=cfg["S"], normals=Normals(queue=q))
    m1 = ref.predict(data["xs"], num_samples=cfg["S"], normals=TorchNormals(queue=q))
    np.testing.assert_allclose(m1, m0, rtol=1e-7, atol=1e-9)


def test_timed_step_budget_and_phases():
    kw = KW[0]
    cfg = dict(n=150, m=2, p=3, ns=40, S=4, missing=0.1)
    data = bench.make_data(**cfg)
    full = timed_step(kw, data, cfg["S"])
    assert full["chains"] == 4 and full["t_logpdf"] > 0 and full["t_condition"] > 0
    ora = OracleRegressor(**kw)
    ora.condition(data["x"], data["y"])
    assert abs(full["logpdf"] - ora.logpdf(data["x"], data["y"])) <= 1e-9 * abs(full["logpdf"])
    cut = timed_step(kw, data, cfg["S"], budget_s=0.0)  # always runs at least one chain
    assert cut["chains"] == 1
