"""Multi-GPU sharding of the GPAR hot path: one process per GPU, torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests) for the plumbing.

What shards: the S Monte-Carlo chains of ``sample`` / ``predict`` are independent given the
conditioned layers (model.py:245-277, regression.py:557-563).  Each rank conditions redundantly
(one joint factorisation per layer -- cheaper than shipping 8.6 GB factors at n = 32768) and runs a
contiguous slice of the chains; the only exchange is the reduction of the (n*, p) sample sum
(``predict``) or the gather of the samples (``sample`` / credible bounds).  With ``replace=True`` all
chains share their inputs (U_i = 1) and there is nothing to shard: replicas only.

The blocked Cholesky of one large layer also shards (BASELINE config 5): tile rows are
dealt block-cyclically to the ranks, every rank keeps a full copy of the matrix in peer-mapped memory
(:class:`PeerBuffer`, CUDA IPC) and the persistent dataflow kernel pushes each finished tile into all
peers' copies over NVLink from inside the kernel (``gpar_potrf_multi``) -- see :func:`potrf_sharded`
and :func:`layer_logpdf_sharded`.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["chain_slice", "shard_normals", "predict_sharded", "sample_sharded", "tile_row_owner", "PeerBuffer",
           "potrf_sharded", "layer_logpdf_sharded"]


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


def chain_generator(device, start, group=None):
    """CUDA generator for the chains [start, ...) of this rank.  Ranks of an SPMD program are usually
    seeded identically (``torch.manual_seed(k)`` everywhere): drawing from the default generator would
    make every rank sample the SAME chains and the gathered set would hold S / world distinct ones.
    A base seed is drawn on group rank 0, broadcast, and offset by the first chain index of the rank."""
    base = torch.zeros(1, dtype=torch.int64, device=device)
    rank, world = _world(group)
    if rank == 0:
        base[0] = int(torch.randint(0, 2 ** 62, (1,)).item())
    if world > 1:
        dist.broadcast(base, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(base.item()) + 1000003 * int(start))
    return gen


def _local_samples(reg, x, w, start, stop, latent, normals, local_sampler, group=None):
    """(S_loc, n*, p) tensor of this rank's chains (already un-normalised / un-transformed)."""
    if local_sampler is not None:
        if stop <= start:
            return None
        return torch.as_tensor(np.asarray(local_sampler(start, stop)))
    gen = None
    if normals is None:  # collective: every rank takes part, also one that owns no chain
        gen = chain_generator(_dev(group) if dist.is_initialized() else torch.device("cuda", torch.cuda.current_device()),
                              start, group)
    if stop <= start:
        eng = getattr(reg, "_engine", None)
        if eng is not None and eng.group is not None:
            # a rank without chains still takes part in the sharded factorisations of the conditioning
            reg._sample_device(x, w, None, True, 1, latent, None, generator=gen)
        return None
    dev = reg._sample_device(x, w, None, True, stop - start, latent, shard_normals(normals, start, stop),
                             generator=gen)
    # un-normalise / un-transform per sample (regression.py:553-562): on the device for the transforms it
    # knows, else through the host callables
    from .regression import _transform_kind

    kind = _transform_kind(reg._untransform_y)
    if kind is not None:
        reg._untransform_device(reg._engine_of(dev), dev, kind)
        return dev
    smp = dev.cpu().numpy()
    smp = np.stack([reg._untransform_y(reg._unnormalise_y(smp[s])) for s in range(smp.shape[0])])
    return torch.as_tensor(smp, device=dev.device)


def predict_sharded(reg, x, w=None, num_samples=100, latent=False, normals=None, group=None, local_sampler=None):
    """``GPARRegressor.predict`` with the chains partitioned over the ranks of ``group``; every
    rank returns the same (n*, p) mean.  ``local_sampler(start, stop)`` overrides the engine (used by
    the gloo tests to exercise the partition / reduction logic without a GPU)."""
    rank, world = _world(group)
    start, stop = chain_slice(num_samples, rank, world)
    smp = _local_samples(reg, x, w, start, stop, latent, normals, local_sampler, group)
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
    smp = _local_samples(reg, x, w, start, stop, latent, normals, local_sampler, group)
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


# ---------------------------------------------------------------------------------------------
# Sharded Cholesky
# ---------------------------------------------------------------------------------------------
ROW_BLOCK = 4  # GPAR_ROW_BLOCK of include/gpar_b200.h


def tile_row_owner(i, world, row_block=ROW_BLOCK):
    """Rank that factors / solves the tiles of tile row ``i``: block-cyclic deal, ``row_block``
    consecutive tile rows per turn (the critical chain crosses NVLink once per block)."""
    return (int(i) // int(row_block)) % int(world)


class PeerBuffer:
    """``nbytes`` of device memory on every rank of ``group``, each rank's allocation mapped into all
    the others (CUDA IPC handles exchanged through torch.distributed).  ``deltas[r]`` is the address
    of rank r's allocation minus this rank's -- what ``gpar_potrf_multi`` consumes; ``view(off, n)``
    is a float64 torch view of the local allocation (plumbing only)."""

    def __init__(self, eng, nbytes, group=None):
        self.eng, self.group = eng, group
        self.rank, self.world = _world(group)
        self.nbytes = int(nbytes)
        lib = eng.lib
        base = C.c_void_p()
        _check(lib.gpar_ipc_alloc(self.nbytes, C.byref(base)), "gpar_ipc_alloc")
        self.base = int(base.value)
        self.peers = [self.base] * self.world
        self._opened = []
        if self.world > 1:
            handle = C.create_string_buffer(64)
            _check(lib.gpar_ipc_export(C.c_void_p(self.base), handle), "gpar_ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                ptr = C.c_void_p()
                _check(lib.gpar_ipc_open(C.create_string_buffer(h, 64), C.byref(ptr)), "gpar_ipc_open")
                self.peers[r] = int(ptr.value)
                self._opened.append(int(ptr.value))
        self.deltas = (C.c_int64 * max(self.world, 1))(*[q - self.base for q in self.peers])

    def view(self, offset_doubles, numel):
        """float64 tensor over [offset, offset + numel) doubles of the LOCAL allocation."""
        iface = {"shape": (int(numel),), "typestr": "<f8", "version": 3,
                 "data": (self.base + 8 * int(offset_doubles), False)}
        holder = type("_CudaArray", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.eng.device)

    def close(self):
        lib = self.eng.lib
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        for q in self._opened:
            lib.gpar_ipc_close(C.c_void_p(q))
        self._opened = []
        if self.base:
            lib.gpar_ipc_free(C.c_void_p(self.base))
            self.base = 0


def _check(rc, what):
    from ._lib import check

    check(rc, what)


def _sync_ranks(group):
    torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)


def potrf_layout(eng, n, nb):
    """Offsets (in doubles) of A (n x ld), B (nb x ld), the workspace and the info word inside a
    :class:`PeerBuffer`, and the total size in bytes.  Identical on every rank by construction."""
    ld = n + (n & 1)
    ws_doubles = eng.lib.gpar_potrf_workspace_bytes(n, nb, 1) // 8
    off_a = 0
    off_b = off_a + n * ld
    off_ws = off_b + max(nb, 0) * ld
    off_ws += off_ws & 1
    off_info = off_ws + ws_doubles
    off_info += off_info & 1
    total = off_info + 2
    return dict(ld=ld, a=off_a, b=off_b, ws=off_ws, info=off_info, bytes=8 * total)


def potrf_sharded(eng, buf, n, nb=0, group=None):
    """In-place Cholesky of the matrix that every rank has written into ``buf`` (layout of
    :func:`potrf_layout`), spread over the ranks of ``group``.  Every rank returns with the complete
    factor, ``B L^-T`` and workspace in its own copy.  Raises on a non-positive pivot (any rank)."""
    lay = potrf_layout(eng, n, nb)
    lib = eng.lib
    A, B, ws = buf.base + 8 * lay["a"], buf.base + 8 * lay["b"], buf.base + 8 * lay["ws"]
    info = buf.base + 8 * lay["info"]
    _check(lib.gpar_potrf_multi_reset(C.c_void_p(ws), n, nb, C.c_void_p(info), eng.stream), "gpar_potrf_multi_reset")
    _sync_ranks(group)  # every rank's matrix and flags are in place before any peer writes into them
    _check(lib.gpar_potrf_multi(C.c_void_p(A), lay["ld"], n, C.c_void_p(B) if nb > 0 else None, lay["ld"], nb,
                                C.c_void_p(ws), C.c_void_p(info), buf.rank, buf.world, buf.deltas, eng.stream),
           "gpar_potrf_multi")
    eng.launches += 1
    eng.flops += (n ** 3 / 3.0 + nb * float(n) ** 2) / buf.world
    _sync_ranks(group)  # all pushes have landed everywhere
    bad = buf.view(lay["info"], 1).view(torch.int32)[:1].clone().to(torch.int64)
    if buf.world > 1:
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
    if int(bad[0]) != 0:
        from ._lib import GparError

        raise GparError(f"Cholesky failed: matrix not positive definite (first non-positive pivot {int(bad[0])})")
    return lay


def layer_logpdf_sharded(eng, spec, X, d, y, group=None, buf=None):
    """Dense log-marginal likelihood of ONE layer (model.py:226 for a dense ``Obs``) with the
    Cholesky spread over the ranks: K = k(X, X) + diag(d) + eps I built redundantly on every rank
    (HBM-bound, cheap), factored by :func:`potrf_sharded` with y riding along as an appended row,
    then -1/2 (logdet + n log 2 pi + ||L^-1 y||^2) from the local copy.  X (n, dcols), d (n,), y (n,)
    are device tensors holding the same values on every rank."""
    import math

    n, dcols = int(X.shape[0]), int(X.shape[1])
    lay = potrf_layout(eng, n, 1)
    own = buf is None
    if own:
        buf = PeerBuffer(eng, lay["bytes"], group)
    try:
        J = buf.view(lay["a"], n * lay["ld"])
        u = buf.view(lay["b"], lay["ld"])
        u.zero_()
        u[:n].copy_(y)
        eng.gram(spec, X.reshape(-1), dcols, n, J, lay["ld"], diag=d, lower_only=True)
        potrf_sharded(eng, buf, n, 1, group)
        out2 = eng.empty(2)
        eng.logdet_quad(J, lay["ld"], n, u, out2)
        ld_q = out2.cpu().numpy()
        return -0.5 * (ld_q[0] + n * math.log(2.0 * math.pi) + ld_q[1])
    finally:
        if own:
            buf.close()
