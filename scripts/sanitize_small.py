"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck): the persistent dataflow
Cholesky with appended rows (multi-tile: HEAD / PRE / PLAIN tasks, ready flags), backsolve, Gram kernels, the
diverged-chain batch path and one regressor logpdf + predict.  Sizes are small: the sanitizer slows kernels
down by 10-100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from gpar_b200 import GPARRegressor
from gpar_b200.engine import Engine
from gpar_b200.spec import lower_terms

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
eng = Engine()
spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25] * 2)])
X = torch.rand(n * 2, dtype=torch.float64, device=eng.device)
d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
ld = n + (n & 1)
J = eng.empty(n * ld)
B = torch.rand(130 * ld, dtype=torch.float64, device=eng.device)
eng.gram(spec, X, 2, n, J, ld, diag=d, lower_only=True)
ws, info = eng.potrf(J, ld, n, B=B, ldb=ld, nb=130)
alpha = eng.backsolve(J, ld, n, ws, B)
torch.cuda.synchronize()
print("potrf info", int(info.cpu()[0]), "alpha finite", bool(torch.isfinite(alpha).all()))
for name in ("tiny",):
    data_kw, reg_kw = bench.CONFIGS[name]
    data = bench.make_data(**data_kw)
    reg = GPARRegressor(engine=eng, **reg_kw)
    reg.condition(data["x"], data["y"])
    lp = reg.logpdf(data["x"], data["y"])
    mean = reg.predict(data["xs"], num_samples=4)
    print(name, "logpdf", float(lp), "mean finite", bool(np.isfinite(mean).all()))
data_kw, reg_kw = bench.CONFIGS["c2"]
data = bench.make_data(**{**data_kw, "n": 300, "ns": 40, "S": 3})
reg = GPARRegressor(engine=eng, **reg_kw)
reg.condition(data["x"], data["y"])
print("diverged chains", bool(np.isfinite(reg.predict(data["xs"], num_samples=3)).all()))
