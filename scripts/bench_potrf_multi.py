"""Timing of the sharded Cholesky (gpar_potrf_multi) under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_potrf_multi.py 16384 32768
Device time (CUDA events, max over ranks) of gram + factorisation of one layer; rank 0 prints one JSON
line per size with the achieved fp64 TFLOP/s of the whole job."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ctypes as C
from gpar_b200.engine import Engine
from gpar_b200.spec import lower_terms
from gpar_b200.dist import PeerBuffer, potrf_layout, _check, _sync_ranks

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine()
spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25, 0.25])])
for n in [int(a) for a in (sys.argv[1:] or ["8192", "16384"])]:
    rng = np.random.default_rng(0)
    X = eng.to_device(rng.uniform(0, 1, (n, 2))).reshape(-1)
    d = eng.to_device(np.full(n, 0.1))
    lay = potrf_layout(eng, n, 1)
    buf = PeerBuffer(eng, lay["bytes"], None if world == 1 else dist.group.WORLD)
    J = buf.view(lay["a"], n * lay["ld"])
    A, B, ws, info = (buf.base + 8 * lay[k] for k in ("a", "b", "ws", "info"))
    ts = []
    for it in range(4):
        eng.gram(spec, X, 2, n, J, lay["ld"], diag=d, lower_only=True)
        buf.view(lay["b"], lay["ld"]).fill_(1.0)
        _check(eng.lib.gpar_potrf_multi_reset(C.c_void_p(ws), n, 1, C.c_void_p(info), eng.stream), "reset")
        _sync_ranks(None if world == 1 else dist.group.WORLD)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _check(eng.lib.gpar_potrf_multi(C.c_void_p(A), lay["ld"], n, C.c_void_p(B), lay["ld"], 1, C.c_void_p(ws),
                                        C.c_void_p(info), buf.rank, buf.world, buf.deltas, eng.stream), "multi")
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        _sync_ranks(None if world == 1 else dist.group.WORLD)
        ts.append(float(t[0]))
    ms = min(ts[1:])
    chk = float(torch.log(J[:: lay["ld"] + 1][:n]).sum())  # sum log L_ii: identical on every rank
    if rank == 0:
        print(json.dumps({"n": n, "n_gpus": world, "potrf_ms": ms, "tflops": (n ** 3 / 3 + n * n) / ms / 1e9,
                          "sum_log_diag": chk, "all_ms": ts}), flush=True)
    buf.close()
if world > 1:
    dist.destroy_process_group()
