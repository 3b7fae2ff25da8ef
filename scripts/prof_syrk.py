"""Profiling driver: SYRK update at the C3 trailing size, short K (right-looking step) and long K
(left-looking column) -- run under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpar_b200.engine import Engine
eng = Engine()
n = 8192
for k in (128, 2048):
    Cm = torch.zeros(n * n, dtype=torch.float64, device=eng.device)
    Wm = torch.randn(n * k, dtype=torch.float64, device=eng.device)
    for _ in range(3):
        eng.syrk_sub(Cm, n, n, Wm, k, k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        eng.syrk_sub(Cm, n, n, Wm, k, k)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"syrk n={n} k={k}: {ms:.3f} ms, {n*(n+1)*k/ms/1e9:.2f} TFLOP/s")
