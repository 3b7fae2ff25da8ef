"""One pass of trsm_rows_kernel at the C5 shape (n = 32768 factor, rows of a pass of diverged chains) for
ncu --set full.  usage: one_trsm_rows.py [n] [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpar_b200.engine import Engine
from gpar_b200.spec import lower_terms
eng = Engine()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 32 * 2048
spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25] * 2)])
X = torch.rand(n * 2, dtype=torch.float64, device=eng.device)
d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
ld = n
J = eng.empty(n * ld)
eng.gram(spec, X, 2, n, J, ld, diag=d, lower_only=True)
ws, info = eng.potrf(J, ld, n)
E = torch.rand(nb * ld, dtype=torch.float64, device=eng.device)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.trsm_rows(J, ld, n, ws, E, ld, nb)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(f"trsm_rows n={n} rows={nb}: {ms:.1f} ms, {nb * float(n) ** 2 / ms / 1e9:.2f} TFLOP/s, info {int(info.cpu()[0])}")
