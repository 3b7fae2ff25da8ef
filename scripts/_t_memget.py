import time, torch, sys, os
sys.path.insert(0, "/root/repo")
x = torch.zeros(1, device="cuda"); torch.cuda.synchronize()
big = torch.empty(int(3.3e9 // 8), dtype=torch.float64, device="cuda")
for _ in range(3):
    t0 = time.perf_counter(); f = torch.cuda.mem_get_info(); t1 = time.perf_counter()
    print("mem_get_info %.3f ms" % ((t1 - t0) * 1e3))
t0 = time.perf_counter(); p = torch.cuda.get_device_properties(0).multi_processor_count; print("props %.3f ms" % ((time.perf_counter() - t0) * 1e3))
