import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpar_b200.engine import Engine
from gpar_b200.spec import lower_terms
eng = Engine()
spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4)])
n = 128
X = torch.rand(n, 4, dtype=torch.float64, device=eng.device).reshape(-1)
d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
J = eng.empty(n * n)
ws = eng.empty(eng.lib.gpar_potrf_workspace_bytes(n, 0, 1) // 8)
info = torch.zeros(1, dtype=torch.int32, device=eng.device)
prof = torch.zeros(32, dtype=torch.int64, device=eng.device)
for it in range(3):
    eng.gram(spec, X, 4, n, J, n, diag=d, lower_only=True)
    eng.lib.gpar_debug_diag_profile(eng.addr(J), n, n, eng.addr(ws), C.c_void_p(info.data_ptr()), C.c_void_p(prof.data_ptr()), eng.stream)
    torch.cuda.synchronize()
p = prof.cpu().numpy()
names = {0: "start", 1: "tile loaded", 10: "sweep done", 11: "norms + Linv written, flag", 12: "L written"}
for q in range(4):
    names[2 + 2 * q] = f"panel {q}: factor32 + row solves"
    names[3 + 2 * q] = f"panel {q}: rank-32 update"
names[13] = "  ||L|| pass"; names[14] = "  Linv pass"
for q in range(4):
    names[16 + q] = f"  panel {q}: warp 0 (factor32) done"; names[20 + q] = f"  panel {q}: warp 1 (row solve) done"; names[24 + q] = f"  panel {q}: warp 4 (row solve) done"
order = sorted([i for i in range(28) if p[i] != 0], key=lambda i: p[i])
prev = p[0]
for i in order:
    print(f"{i:2d} {names.get(i,''):34s} +{p[i]-prev:8d} cyc  (t={p[i]-p[0]})"); prev = p[i]
print("total cycles", p[12] - p[0], "=> us @1.965GHz", (p[12] - p[0]) / 1965.0)
