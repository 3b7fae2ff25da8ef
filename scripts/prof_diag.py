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
names = {0: "start", 1: "load", 19: "16-step sweep (L and Linv)", 20: "norms + writeback"}
for b, k in ((2, 0), (8, 8)):
    names[b] = f"k={k}: barrier A"; names[b+1] = f"k={k}: row solve (thread 0)"; names[b+2] = f"k={k}: barrier B"
    names[b+3] = f"k={k}: w0 diag update"; names[b+4] = f"k={k}: w0 factor_block"
names[13]="  fb(1): after loads"; names[14]="  fb(1): after factor chain"; names[15]="  warp1 k=0 bulk start"; names[16]="  warp1 k=0 bulk end"
order = sorted([i for i in range(21) if p[i] != 0], key=lambda i: p[i])
prev = p[0]
for i in order:
    print(f"{i:2d} {names.get(i,''):30s} +{p[i]-prev:8d} cyc  (t={p[i]-p[0]})"); prev = p[i]
sys.exit()
prev = p[0]
for i in range(21):
    if p[i] == 0: continue
    print(f"{i:2d} {names.get(i,''):22s} +{p[i]-prev:8d} cyc  (t={p[i]-p[0]})")
    prev = p[i]
print("total cycles", p[20] - p[0], "=> us @1.965GHz", (p[20] - p[0]) / 1965.0)
