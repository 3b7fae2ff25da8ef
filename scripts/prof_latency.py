import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpar_b200.engine import Engine
eng = Engine()
out = eng.zeros(32)
for _ in range(2):
    __import__('gpar_b200')._lib.load_debug().gpar_debug_latency_probe(eng.addr(out), eng.stream); torch.cuda.synchronize()
names = ["DMMA dependent", "4xDMMA k-step", "DFMA dependent", "SHFL double", "rsqrt", "sqrt", "div", "smem st->ld roundtrip", "rsqrtf+2 Newton"]
for n, v in zip(names, out.cpu().numpy()): print(f"{n:26s} {v:8.1f} cycles")
