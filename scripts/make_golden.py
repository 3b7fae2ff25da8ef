"""Writes tests/golden/oracle_small.npz: outputs of the CPU oracle (oracle/gpar_oracle.py) on small seeded
workloads.  The reference's own stack (stheno / lab / matrix / varz) is not importable in the build
container, so these are ORACLE outputs, not reference outputs: they pin the oracle against
drift (tests/test_oracle_golden.py) and give the GPU parity tests a committed target next to the live
oracle run.  Re-generate with:  python scripts/make_golden.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle.gpar_oracle import Normals, OracleRegressor

CASES = {
    # name: (data kwargs, regressor kwargs)
    "c3_small": (dict(n=160, m=2, p=3, ns=40, S=3, missing=0.1),
                 dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
                      markov=2, replace=True, impute=True, normalise_y=True)),
    "c2_small": (dict(n=120, m=2, p=2, ns=30, S=3, missing=0.0),
                 dict(scale=0.25, noise=0.1, linear=False, nonlinear=True, nonlinear_scale=1.0, replace=False,
                      impute=False, normalise_y=True)),
    "rq_per_small": (dict(n=90, m=1, p=2, ns=20, S=2, missing=0.15),
                     dict(scale=0.4, noise=0.05, rq=True, per=True, per_period=0.7, input_linear=True, linear=True,
                          nonlinear=True, replace=False, impute=True, normalise_y=True)),
}


def run(name):
    data_kw, reg_kw = CASES[name]
    data = bench.make_data(**data_kw)
    ora = OracleRegressor(**reg_kw)
    ora.condition(data["x"], data["y"])
    lp = ora.logpdf(data["x"], data["y"])
    S, p = data_kw["S"], data_kw["p"]
    queue = [data["Z"][s, i] for s in range(S) for i in range(p)]
    mean = ora.predict(data["xs"], num_samples=S, normals=Normals(queue=queue))
    return {f"{name}/logpdf": np.float64(lp), f"{name}/mean": mean}


if __name__ == "__main__":
    out = {}
    for name in CASES:
        out.update(run(name))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_small.npz")
    np.savez(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})
