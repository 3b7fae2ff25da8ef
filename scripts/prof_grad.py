"""Cost of one analytic gradient of the dense log-marginal (potrf + backsolve + potri + gram_grad) vs the
forward pass alone, i.e. vs what one finite-difference component costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpar_b200.engine import Engine, Factor
from gpar_b200.spec import lower_terms
eng = Engine()
terms = [dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4),
         dict(type="linear", variance=1.0, cols=[4, 5], scales=[10.0, 10.0]),
         dict(type="eq", variance=1.0, cols=[4, 5], scales=[1.0, 1.0])]
spec = lower_terms(terms)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for n in [int(a) for a in (sys.argv[1:] or ["1024", "2048", "4096", "7424"])]:
    X = torch.rand(n, 6, dtype=torch.float64, device=eng.device)
    d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
    y = torch.randn(n, dtype=torch.float64, device=eng.device)
    fac = Factor(eng, spec, X.reshape(-1), 6, d, y, n, 0)
    alpha = fac.alpha()
    t_fwd = timeit(lambda: Factor(eng, spec, X.reshape(-1), 6, d, y, n, 0))
    t_inv = timeit(lambda: eng.potri(fac.J, fac.ld, n, fac.ws))
    Ainv = eng.potri(fac.J, fac.ld, n, fac.ws)
    t_grad = timeit(lambda: eng.gram_grad(spec, X.reshape(-1), 6, n, alpha, Ainv, fac.ld, d))
    nparam = 1 + 4 + 2 + 1 + 2 + 1
    print(f"n={n:6d}  forward (gram+potrf) {t_fwd:8.3f} ms   potri {t_inv:8.3f} ms ({2*n**3/3/t_inv/1e9:5.1f} TF/s)   gram_grad {t_grad:7.3f} ms"
          f"   analytic total {t_fwd+t_inv+t_grad:8.3f} ms  vs finite differences ({nparam} params, central) {2*nparam*t_fwd:8.3f} ms", flush=True)
