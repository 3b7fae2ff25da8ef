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


# ---------------------------------------------------------------------------------------------
# Sharded Cholesky (gpar_potrf_multi): world 1 through the same entry point on any GPU box,
# world 2 over CUDA IPC / NVLink when two GPUs are visible.
# ---------------------------------------------------------------------------------------------
def _sharded_case(eng, n, nb, group, seed=0, near_singular=False):
    """Factor the same SPD matrix with gpar_potrf (one GPU) and potrf_sharded; return the max abs
    differences of L, B L^-T and the inverse tiles.  ``near_singular``: rows 256..639 are a tight cluster with
    noise 1e-9, so the diagonal tiles 2..4 are ill-conditioned (kappa_inf(L_kk) >> 1e3): their refinement flags
    must reach the peers together with the tiles (potrf.cu: flag_out pushes) -- a k > 0 tile in multi-GPU mode."""
    import scipy.linalg as sla

    from gpar_b200.dist import PeerBuffer, potrf_layout, potrf_sharded
    from gpar_b200.spec import lower_terms

    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, (n, 2))
    d = rng.uniform(0.05, 0.2, n)
    if near_singular:
        X[256:640] = X[256] + 2e-3 * rng.uniform(-1, 1, (384, 2))
        d[256:640] = 1e-9
    Bh = rng.standard_normal((max(nb, 1), n))
    spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25, 0.25])])
    Xd, dd = eng.to_device(X).reshape(-1), eng.to_device(d)
    lay = potrf_layout(eng, n, nb)
    ld = lay["ld"]
    Bp = np.zeros((max(nb, 1), ld)); Bp[:, :n] = Bh
    # single-GPU reference through gpar_potrf
    J1 = eng.empty(n * ld); B1 = eng.to_device(Bp).reshape(-1)
    eng.gram(spec, Xd, 2, n, J1, ld, diag=dd, lower_only=True)
    ws1, info1 = eng.potrf(J1, ld, n, B=B1 if nb else None, ldb=ld, nb=nb)
    # sharded
    buf = PeerBuffer(eng, lay["bytes"], group)
    J2 = buf.view(lay["a"], n * ld)
    eng.gram(spec, Xd, 2, n, J2, ld, diag=dd, lower_only=True)
    if nb:
        buf.view(lay["b"], nb * ld).copy_(eng.to_device(Bp).reshape(-1)[: nb * ld])
    potrf_sharded(eng, buf, n, nb, group)
    L1 = np.tril(J1.cpu().numpy().reshape(n, ld)[:, :n])
    L2 = np.tril(J2.cpu().numpy().reshape(n, ld)[:, :n])
    K = np.exp(-0.5 * ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) / 0.25 ** 2) + np.diag(d + 1e-12)
    Lref = sla.cholesky(K, lower=True)
    err_ref = np.linalg.norm(L2 - Lref) / np.linalg.norm(Lref)
    dL = float(np.abs(L1 - L2).max())
    dB = 0.0
    if nb:
        dB = float(np.abs(B1.cpu().numpy()[: nb * ld] - buf.view(lay["b"], nb * ld).cpu().numpy()).max())
    nt = (n + 127) // 128
    W1 = ws1.cpu().numpy()[: nt * 128 * 128]
    W2 = buf.view(lay["ws"], nt * 128 * 128).cpu().numpy()
    dW = 0.0  # inverse tiles: only the kb x kb lower triangle of each is defined
    for k in range(nt):
        kb = min(128, n - 128 * k)
        a1 = np.tril(W1.reshape(nt, 128, 128)[k, :kb, :kb]); a2 = np.tril(W2.reshape(nt, 128, 128)[k, :kb, :kb])
        dW = max(dW, float(np.abs(a1 - a2).max()))
    buf.close()
    return dL, dB, dW, float(err_ref)


@pytest.mark.parametrize("n,nb,sing", [(100, 1, False), (640, 0, False), (1000, 130, False), (1300, 70, True)])
def test_potrf_sharded_world1_equals_potrf(n, nb, sing):
    from gpar_b200.engine import Engine

    eng = Engine()
    dL, dB, dW, err_ref = _sharded_case(eng, n, nb, None, near_singular=sing)
    assert dL == 0.0 and dB == 0.0 and dW == 0.0  # same kernel, same tile arithmetic
    assert err_ref <= (1e-3 if sing else 1e-12 * (1 + np.log(n)))


def _worker_potrf(rank, world, port, q):
    import torch.distributed as dist

    from gpar_b200.dist import layer_logpdf_sharded
    from gpar_b200.engine import Engine
    from gpar_b200.spec import lower_terms

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = Engine()
    res = [_sharded_case(eng, n, nb, None if world == 1 else dist.group.WORLD, seed=n) for n, nb in
           ((300, 1), (1500, 0), (2100, 200))]
    res.append(_sharded_case(eng, 1300, 70, None if world == 1 else dist.group.WORLD, seed=3, near_singular=True))
    # the sharded dense log-marginal of one layer against scipy
    rng = np.random.default_rng(5)
    n = 1800
    X = rng.uniform(0, 1, (n, 2)); d = np.full(n, 0.1); y = rng.standard_normal(n)
    spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25, 0.25])])
    lp = layer_logpdf_sharded(eng, spec, eng.to_device(X), eng.to_device(d), eng.to_device(y), dist.group.WORLD)
    if rank == 0:
        import scipy.linalg as sla

        K = np.exp(-0.5 * ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) / 0.25 ** 2) + np.diag(d + 1e-12)
        Lr = sla.cholesky(K, lower=True)
        u = sla.solve_triangular(Lr, y, lower=True)
        ref = -0.5 * (2 * np.log(np.diag(Lr)).sum() + n * np.log(2 * np.pi) + u @ u)
        q.put((res, float(abs(lp - ref) / abs(ref))))
    dist.destroy_process_group()


def test_potrf_sharded_world2_nvlink():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_potrf, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res, rel_lp = q.get()
    for dL, dB, dW, err_ref in res[:3]:
        assert dL == 0.0 and dB == 0.0 and dW == 0.0  # tile arithmetic does not depend on the owner
        assert err_ref <= 1e-11
    dL, dB, dW, err_ref = res[3]  # near-singular tiles at k = 2..4: refined solves on both ranks, same bits
    assert dL == 0.0 and dB == 0.0 and dW == 0.0
    assert err_ref <= 1e-3  # forward error of the factor is conditioning-limited (kappa ~ 1e11); bits match above
    assert rel_lp <= 1e-10


def _worker_regressor(rank, world, port, q):
    import torch.distributed as dist

    import bench
    from gpar_b200 import GPARRegressor
    from gpar_b200.dist import predict_sharded
    from gpar_b200.engine import Engine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    data_kw, reg_kw = bench.CONFIGS["c2"]
    data = bench.make_data(**{**data_kw, "n": 1500, "ns": 200, "S": 6})
    normals = {"Z": data["Z"]}
    # engine with a process group: conditioning factorisations (n >= 1024 here) are spread over the
    # ranks, the chains are partitioned over the ranks
    reg = GPARRegressor(engine=Engine(group=dist.group.WORLD, shard_min_n=1024), **reg_kw)
    reg.condition(data["x"], data["y"])
    lp = reg.logpdf(data["x"], data["y"])
    mean = predict_sharded(reg, data["xs"], num_samples=6, normals=normals)
    lp_again = reg.logpdf(data["x"], data["y"])  # peer buffers come back from the pool
    assert lp_again == lp
    # fit on a sharded engine: every L-BFGS evaluation builds fresh factors; the peer buffers must be recycled
    fit = GPARRegressor(engine=reg._engine, **reg_kw)
    fit.fit(data["x"][:1100], data["y"][:1100, :2], iters=4)
    pooled = sum(len(v) for v in reg._engine._peer_pool.values()) + len(reg._engine._peer_bufs)
    assert pooled <= 4, pooled
    vs = fit.get_variables()
    assert all(np.all(np.isfinite(v)) for v in vs.values())
    reg._engine.close_peer_buffers()
    if rank == 0:
        ref = GPARRegressor(engine=Engine(), **reg_kw)
        ref.condition(data["x"], data["y"])
        lp_ref = ref.logpdf(data["x"], data["y"])
        mean_ref = ref.predict(data["xs"], num_samples=6, normals=normals)
        q.put((float(abs(lp - lp_ref) / abs(lp_ref)), float(np.abs(mean - mean_ref).max())))
    dist.barrier()
    dist.destroy_process_group()


def test_regressor_sharded_cholesky_and_chains_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_regressor, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    rel_lp, err_mean = q.get()
    assert rel_lp <= 1e-13 and err_mean <= 1e-12
