"""Timing of the Cholesky sweep and its companions at several sizes (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpar_b200.engine import Engine
from gpar_b200.spec import lower_terms
eng = Engine()
terms = [dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4)]
spec = lower_terms(terms)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
sizes = [int(s) for s in (sys.argv[1:] or "128 256 512 1024 2048 4096 6144 7424 8192 9216".split())]
for n in sizes:
    X = torch.rand(n, 4, dtype=torch.float64, device=eng.device).reshape(-1)
    d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
    ld = n + (n & 1)
    J = eng.empty(n * ld); u = eng.zeros(ld)
    def gram(): eng.gram(spec, X, 4, n, J, ld, diag=d, lower_only=True)
    tg = timeit(gram)
    def both():
        gram(); eng.potrf(J, ld, n, B=u, ldb=ld, nb=1)
    tb = timeit(both)
    tp = tb - tg
    gram(); ws, info = eng.potrf(J, ld, n, B=u, ldb=ld, nb=1)
    tbs = timeit(lambda: eng.backsolve(J, ld, n, ws, u))
    Kt = torch.empty(n, n, dtype=torch.float64, device=eng.device)
    eng.gram(spec, X, 4, n, Kt.reshape(-1), n, diag=d, lower_only=False)
    Kt = torch.tril(Kt) + torch.tril(Kt, -1).T
    tcs = timeit(lambda: torch.linalg.cholesky(Kt)) if n >= 1024 else float("nan")
    print(f"   cusolver potrf {tcs:8.3f} ms ({n**3/3/max(tcs,1e-9)/1e9:6.2f} TF/s)")
    print(f"n={n:6d} gram {tg:7.3f} ms  potrf {tp:8.3f} ms ({n**3/3/tp/1e9:6.2f} TF/s)  per-col {1e3*tp/((n+127)//128):7.1f} us  backsolve {tbs:7.3f} ms", flush=True)
# batched small: many independent 128 tiles => diag_factor_tile cost
for n, batch in ((128, 148), (128, 1480), (256, 148), (1024, 100)):
    ld = n
    base = torch.rand(n, 4, dtype=torch.float64, device=eng.device).reshape(-1)
    J = eng.empty(batch * n * ld)
    def fill():
        for b in range(1):
            pass
    d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
    def run():
        eng.gram_batched(spec, base, 4, n, 0, J, ld, n * ld, batch, diag=d, strideD=0)
        eng.potrf(J, ld, n, batch=batch, strideA=n * ld)
    def g():
        eng.gram_batched(spec, base, 4, n, 0, J, ld, n * ld, batch, diag=d, strideD=0)
    t = timeit(run) - timeit(g)
    print(f"batched n={n} batch={batch}: potrf {t:.3f} ms -> {1e3*t/(-(-batch//148)):.1f} us per wave")
