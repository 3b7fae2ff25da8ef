"""Two-rank NCCL test of the chain-sharded predict / sample (needs >= 2 visible GPUs; skipped on a
one-GPU box).  Each rank conditions redundantly and runs its slice of the chains with the same
injected normals; results must equal the single-rank run bit for bit (same kernels, same order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import bench
    from gpar_b200 import GPARRegressor
    from gpar_b200.dist import predict_sharded, sample_sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    data_kw, reg_kw = bench.CONFIGS["c2"]
    data = bench.make_data(**{**data_kw, "n": 600, "ns": 100, "S": 5})
    reg = GPARRegressor(**reg_kw)
    reg.condition(data["x"], data["y"])
    normals = {"Z": data["Z"]}
    mean = predict_sharded(reg, data["xs"], num_samples=5, normals=normals)
    smp = np.stack(sample_sharded(reg, data["xs"], num_samples=5, normals=normals))
    if rank == 0:
        ref = np.stack(reg.sample(data["xs"], num_samples=5, posterior=True, normals=normals))
        q.put((float(np.abs(smp - ref).max()), float(np.abs(mean - ref.mean(axis=0)).max())))
    dist.destroy_process_group()


def test_sharded_chains_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    err_smp, err_mean = q.get()
    assert err_smp <= 1e-12 and err_mean <= 1e-12
