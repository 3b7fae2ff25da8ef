"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chain partition, sharded predict
(all_reduce of the sample sum) and sharded sample (all_gather, chain order) -- SURVEY 8(e)-1.  The
per-rank sampler is the oracle here (test infrastructure), so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from gpar_b200.dist import chain_slice, predict_sharded, sample_sharded, shard_normals


def test_chain_slice_partitions_exactly():
    for S in (0, 1, 2, 7, 100, 256):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                a, b = chain_slice(S, r, world)
                assert 0 <= a <= b <= S
                cover.extend(range(a, b))
            assert cover == list(range(S))
            sizes = [chain_slice(S, r, world)[1] - chain_slice(S, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    Z = {"Z": np.arange(24).reshape(6, 2, 2), "Z2": np.arange(24).reshape(6, 2, 2) + 100}
    sl = shard_normals(Z, 2, 5)
    assert sl["Z"].shape == (3, 2, 2) and sl["Z2"][0, 0, 0] == 108 and shard_normals(None, 0, 1) is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, S, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.gpar_oracle import Normals, OracleRegressor

    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (30, 1)); xs = rng.uniform(0, 1, (9, 1))
    ora = OracleRegressor(replace=False, impute=False, linear=True, nonlinear=True, noise=0.05, scale=0.3)
    y = ora.sample(x, p=2, normals=Normals(rng=np.random.default_rng(1)))
    ora.condition(x, y)
    Z = np.random.default_rng(2).standard_normal((S, 2, 9))

    def local(start, stop):
        queue = [Z[s, i] for s in range(start, stop) for i in range(2)]
        smp = ora.sample(xs, posterior=True, num_samples=stop - start, normals=Normals(queue=queue))
        return np.stack(smp if isinstance(smp, list) else [smp])

    mean = predict_sharded(None, xs, num_samples=S, local_sampler=local)
    allsmp = np.stack(sample_sharded(None, xs, num_samples=S, local_sampler=local))
    if rank == 0:
        full = local(0, S)
        q.put((np.abs(mean - full.mean(axis=0)).max(), np.abs(allsmp - full).max(), allsmp.shape))
    dist.destroy_process_group()


@pytest.mark.parametrize("S", [5, 8])
def test_sharded_predict_and_sample_gloo_world2(S):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    err_mean, err_smp, shape = q.get()
    assert shape == (S, 9, 2)
    assert err_mean <= 1e-13 and err_smp == 0.0


def test_potrf_layout_is_aligned_and_disjoint():
    """Layout of the peer-mapped buffer of the sharded Cholesky (identical on every rank): regions do not
    overlap, bases are 16-byte aligned, the leading dimension is even (C ABI requirements)."""
    import types

    from gpar_b200 import _lib
    from gpar_b200.dist import potrf_layout, tile_row_owner

    eng = types.SimpleNamespace(lib=_lib.load())
    for n, nb in ((1, 0), (127, 1), (128, 1), (1000, 130), (8424, 1)):
        lay = potrf_layout(eng, n, nb)
        assert lay["ld"] >= n and lay["ld"] % 2 == 0
        ws_doubles = eng.lib.gpar_potrf_workspace_bytes(n, nb, 1) // 8
        regions = [(lay["a"], n * lay["ld"]), (lay["b"], max(nb, 0) * lay["ld"]), (lay["ws"], ws_doubles),
                   (lay["info"], 2)]
        end = 0
        for off, size in regions:
            assert off >= end and off % 2 == 0  # doubles: even offset = 16-byte aligned
            end = off + size
        assert lay["bytes"] >= 8 * end
    assert [tile_row_owner(i, 2) for i in range(10)] == [0, 0, 0, 0, 1, 1, 1, 1, 0, 0]


def _rng_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch

    from gpar_b200.dist import chain_generator

    torch.manual_seed(1234)  # SPMD programs seed every rank identically
    S = 6
    start, stop = chain_slice(S, rank, world)
    gen = chain_generator("cpu", start)
    mine = torch.randn(stop - start, 5, generator=gen, dtype=torch.float64)
    # a second call draws a new base seed on rank 0 and broadcasts it: with the same chain offset every rank
    # gets the same stream, i.e. the base really is shared and only the offset separates the ranks
    other = chain_generator("cpu", 7)
    theirs = torch.randn(3, 5, generator=other, dtype=torch.float64)
    q.put((rank, mine.numpy(), theirs.numpy()))
    dist.destroy_process_group()


def test_chain_rng_streams_differ_across_identically_seeded_ranks_gloo_world2():
    """dist.chain_generator: ranks seeded identically must not draw the same chains (the gathered set would hold
    S / world distinct ones); the base seed is shared (drawn on rank 0, broadcast), the offset is the first chain."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_rng_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        rank, mine, theirs = q.get()
        got[rank] = (mine, theirs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert not np.array_equal(got[0][0], got[1][0])           # distinct chains on the two ranks
    assert np.array_equal(got[0][1], got[1][1])               # same (base, offset) -> same stream on both ranks
    assert not np.array_equal(got[0][1], got[0][0][:3])       # and a fresh base per call
