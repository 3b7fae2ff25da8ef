"""One C3-layer joint factorisation (for ncu --set full on potrf_dataflow_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpar_b200.engine import Engine
from gpar_b200.spec import lower_terms
eng = Engine()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8424
spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4)])
X = torch.rand(n * 4, dtype=torch.float64, device=eng.device)
d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
J = eng.empty(n * n); u = eng.zeros(n)
for _ in range(3):
    eng.gram(spec, X, 4, n, J, n, diag=d, lower_only=True)
    eng.potrf(J, n, n, B=u, ldb=n, nb=1)
torch.cuda.synchronize()
print("done", int(eng._infos[-1].cpu()[0]))
