import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpar_b200.engine import Engine
from gpar_b200.spec import lower_terms
eng = Engine()
spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4)])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 7424
X = torch.rand(n, 4, dtype=torch.float64, device=eng.device).reshape(-1)
d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
J = eng.empty(n * n); u = eng.zeros(n)
prof = torch.zeros(64 + 16 * 256 + 1024, dtype=torch.int64, device=eng.device)  # PROF builds also write per-CTA counters
eng.lib.gpar_debug_set_dataflow_prof(C.c_void_p(prof.data_ptr()))
for _ in range(2):
    eng.gram(spec, X, 4, n, J, n, diag=d, lower_only=True)
    eng.potrf(J, n, n, B=u, ldb=n, nb=1)
    torch.cuda.synchronize()
eng.lib.gpar_debug_set_dataflow_prof(None)
p = prof.cpu().numpy().astype(np.int64)
t0 = p[0]
lab = ["task start", "accumulate done, T stored", "4 solve/SYRK blocks done (panel flags)", "X stored + published", "(SYRK: in the blocks)",
       "diagonal tile assembled", "factor + inverse published"]
for base, name in ((0, "HEAD(nt/2)"), (8, "HEAD(nt/2+1)")):
    print(name)
    for i in range(7):
        if p[base + i]: print(f"   {lab[i]:32s} t = {(p[base+i]-t0)/1e3:9.2f} us")
print("diag-to-diag period:", (p[8 + 6] - p[6]) / 1e3, "us")
nt = (n + 127) // 128
tl = p[64 + 16 * 256:][:nt]
if tl[2:].all():
    print("L_kk published, k = 2..: deltas", " ".join(f"{(tl[k] - tl[k-1]) / 1e3:.0f}" for k in range(3, nt)))
