"""Multi-GPU sharding of the GPAR hot path: one process per GPU, torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests) for the plumbing.

What shards: the S Monte-Carlo chains of ``sample`` / ``predict`` are independent given the
conditioned layers (model.py:245-277, regression.py:557-563).  Each rank conditions redundantly
(one joint factorisation per layer -- cheaper than shipping 8.6 GB factors at n = 32768) and runs a
contiguous slice of the chains; the only exchange is the reduction of the (n*, p) sample sum
(``predict``) or the gather of the samples (``sample`` / credible bounds).  With ``replace=True`` all
chains share their inputs (U_i = 1) and there is nothing to shard: replicas only.
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["chain_slice", "shard_normals", "predict_sharded", "sample_sharded"]


def chain_slice(num_samples, rank, world):
    """Contiguous, balanced partition of chains [0, S) over ranks: the first S % world ranks get
    one extra chain.  Returns (start, stop)."""
    base, extra = divmod(int(num_samples), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_normals(normals, start, stop):
    """Slice injected normals {"Z": (S, p, n), "Z2": ...} to the chains of this rank."""
    if normals is None:
        return None
    return {k: np.asarray(v)[start:stop] for k, v in normals.items()}


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _local_samples(reg, x, w, start, stop, latent, normals, local_sampler):
    """(S_loc, n*, p) tensor of this rank's chains (already un-normalised / un-transformed)."""
    if stop <= start:
        return None
    if local_sampler is not None:
        return torch.as_tensor(np.asarray(local_sampler(start, stop)))
    dev = reg._sample_device(x, w, None, True, stop - start, latent, shard_normals(normals, start, stop))
    # un-normalise / un-transform per sample on the host (regression.py:553-562), then back to the device
    smp = dev.cpu().numpy()
    smp = np.stack([reg._untransform_y(reg._unnormalise_y(smp[s])) for s in range(smp.shape[0])])
    return torch.as_tensor(smp, device=dev.device)


def predict_sharded(reg, x, w=None, num_samples=100, latent=False, normals=None, group=None, local_sampler=None):
    """``GPARRegressor.predict`` with the chains partitioned over the ranks of ``group``; every
    rank returns the same (n*, p) mean.  ``local_sampler(start, stop)`` overrides the engine (used by
    the gloo tests to exercise the partition / reduction logic without a GPU)."""
    rank, world = _world(group)
    start, stop = chain_slice(num_samples, rank, world)
    smp = _local_samples(reg, x, w, start, stop, latent, normals, local_sampler)
    if world == 1:
        return smp.double().mean(dim=0).cpu().numpy()
    # shapes are needed on ranks that own no chain
    shape = torch.zeros(2, dtype=torch.int64, device=smp.device if smp is not None else _dev(group))
    if smp is not None:
        shape[0], shape[1] = smp.shape[1], smp.shape[2]
    dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    total = torch.zeros(int(shape[0]), int(shape[1]), dtype=torch.float64, device=shape.device)
    if smp is not None:
        total += smp.double().sum(dim=0)
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return (total / float(num_samples)).cpu().numpy()


def sample_sharded(reg, x, w=None, num_samples=1, latent=False, normals=None, group=None, local_sampler=None):
    """Posterior samples with the chains partitioned over the ranks; every rank returns all
    ``num_samples`` samples (list of (n*, p) arrays, chain order preserved)."""
    rank, world = _world(group)
    start, stop = chain_slice(num_samples, rank, world)
    smp = _local_samples(reg, x, w, start, stop, latent, normals, local_sampler)
    if world == 1:
        return [a for a in smp.cpu().numpy()]
    device = smp.device if smp is not None else _dev(group)
    shape = torch.zeros(2, dtype=torch.int64, device=device)
    if smp is not None:
        shape[0], shape[1] = smp.shape[1], smp.shape[2]
    dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    ns, p = int(shape[0]), int(shape[1])
    # equal-size buffers (base + 1 chains) so that all_gather works on every backend
    cap = -(-int(num_samples) // world)
    buf = torch.zeros(cap, ns, p, dtype=torch.float64, device=device)
    if smp is not None:
        buf[: smp.shape[0]] = smp.double()
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = []
    for r in range(world):
        a, b = chain_slice(num_samples, r, world)
        out.extend(parts[r][: b - a].cpu().numpy())
    return out


def _dev(group):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
