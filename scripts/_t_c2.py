import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gpar_b200 import GPARRegressor
from gpar_b200.engine import Engine
eng = Engine()
data_kw, reg_kw = bench.CONFIGS["c2"]
data = bench.make_data(**data_kw)
reg = GPARRegressor(engine=eng, **reg_kw)
reg.condition(data["x"], data["y"])
for it in range(4):
    torch.cuda.synchronize(); s0 = torch.cuda.memory_stats()
    l0 = eng.launches
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mean = reg.predict(data["xs"], num_samples=100)
    e1.record(); torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print(f"predict wall {1e3*(time.perf_counter()-t0):.1f} ms, events {e0.elapsed_time(e1):.1f} ms, launches {eng.launches-l0}, "
          f"cudaMalloc calls {s1['num_device_alloc']-s0['num_device_alloc']}, frees {s1['num_device_free']-s0['num_device_free']}, reserved {s1['reserved_bytes.all.current']/2**30:.1f} GiB")
