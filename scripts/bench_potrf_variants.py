"""Time gpar_potrf (one appended row) at several n for the library selected by GPAR_B200_LIB.
   python scripts/bench_potrf_variants.py 1024 4096 8424 16384"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpar_b200.engine import Engine  # noqa: E402
from gpar_b200.spec import lower_terms  # noqa: E402

eng = Engine()
out = {"lib": os.environ.get("GPAR_B200_LIB", "default")}
for n in [int(a) for a in sys.argv[1:]] or [8424]:
    spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4)])
    X = torch.rand(n * 4, dtype=torch.float64, device="cuda")
    d = torch.full((n,), 0.1, dtype=torch.float64, device="cuda")
    ld = n + (n & 1)
    J = eng.empty(n * ld)
    u = eng.zeros(ld)
    ts = []
    for it in range(8):
        eng.gram(spec, X, 4, n, J, ld, diag=d, lower_only=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.potrf(J, ld, n, B=u, ldb=ld, nb=1)
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(a.elapsed_time(b))
    out[n] = {"ms_mean": float(np.mean(ts)), "ms_min": float(np.min(ts)),
              "tflops": (n ** 3 / 3 + n ** 2) / np.mean(ts) / 1e9}
    chk = float(J[:: ld + 1][:n].sum().cpu())
    out[n]["diag_sum"] = chk
print(json.dumps(out))
