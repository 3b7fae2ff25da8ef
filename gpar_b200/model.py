"""GPAR model loop on the device engine -- the counterpart of gpar/model.py.

The autoregressive loop over the p outputs stays on the host (north_star); the
closed-downwards masks (``per_output``), ``merge`` and ``last`` are host integer
logic reproduced bit-exactly; every floating-point step is a CUDA kernel behind
the C ABI.  A layer is a :class:`~gpar_b200.spec.LayerModel` (kernel spec +
noise), optionally carrying the observation block it was conditioned on.
"""
import math

import numpy as np
import torch

from .engine import Engine, Factor, F64, _even
from .spec import LayerModel

__all__ = ["GPAR", "merge", "last", "per_output", "construct_model", "DevMat", "ObsBlock"]


# ---------------------------------------------------------------------------
# Host index / mask logic (bit-exact contract; gpar/model.py:14-93, 325-368)
# ---------------------------------------------------------------------------


def merge(x, updates, to_update):
    """Replace ``x[to_update]`` by ``updates`` (in order), keep the rest
    (gpar/model.py:14-44; vectorised scatter with the identical result)."""
    x = np.asarray(x)
    to_update = np.asarray(to_update, dtype=bool)
    out = np.array(x, copy=True)
    if out.dtype != np.result_type(out.dtype, np.asarray(updates).dtype):
        out = out.astype(np.result_type(out.dtype, np.asarray(updates).dtype))
    out[to_update] = np.asarray(updates)
    return out


def construct_model(f, noise):
    """gpar/model.py:47-57."""
    return lambda: (f, noise)


def last(xs, select=None):
    """Zip with an is-last flag; ``select`` filters by index while the flag still
    refers to the unfiltered sequence (gpar/model.py:60-93)."""
    if select is not None:
        select = set(select)
    it = iter(xs)
    try:
        prev = next(it)
    except StopIteration:
        return
    i = 0
    for cur in it:
        if select is None or i in select:
            yield False, prev
        prev = cur
        i += 1
    if select is None or i in select:
        yield True, prev


def per_output(y, w, keep=False):
    """Per-output closed-downwards split (gpar/model.py:325-368): yields
    ``(y[mask, i:i+1], w[mask, i], mask)`` with each mask relative to the previous
    layer's survivors; ``keep`` also retains rows observed in a later output.
    ``y`` may be a dict cache ``{keep: [tuples]}`` (model.py:365-368)."""
    if isinstance(y, dict):
        yield from y[keep]
        return
    y = np.asarray(y)
    w = np.asarray(w)
    p = y.shape[1]
    available = ~np.isnan(y)
    for i in range(p):
        mask = available[:, i].copy()
        if keep and i < p - 1:
            mask |= available[:, i + 1 :].any(axis=1)
        yield y[mask, i : i + 1], w[mask, i], mask
        y, w, available = y[mask], w[mask], available[mask]


# ---------------------------------------------------------------------------
# Device containers
# ---------------------------------------------------------------------------


class DevMat:
    """Row-major fp64 matrix on the device: ``n`` rows, ``d`` logical columns in a
    buffer with leading dimension ``ld >= d`` (spare columns let ``with_col``
    append the next layer's input in place, model.py:320)."""

    def __init__(self, eng, t, n, d, ld, frozen=False):
        self.eng, self.t, self.n, self.d, self.ld = eng, t, int(n), int(d), int(ld)
        self.frozen = frozen  # buffer is shared with a caller: appending a column must copy

    @staticmethod
    def from_host(eng, a, spare=0):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a[:, None]
        n, d = a.shape
        ld = _even(d + spare)
        buf = np.zeros((max(n, 1), ld))
        buf[:n, :d] = a
        return DevMat(eng, eng.to_device(buf).reshape(-1), n, d, ld)

    def take_rows(self, mask):
        """``x[mask]`` (model.py:165,220).  ``mask``: host bool array."""
        mask = np.asarray(mask, dtype=bool)
        if mask.all():
            return self
        idx = np.flatnonzero(mask)
        out = self.eng.empty(max(len(idx), 1) * self.ld)
        if len(idx):
            self.eng.gather_rows(self.t, self.ld, self.eng.to_device(idx, torch.int64), len(idx), self.d, out, self.ld)
        return DevMat(self.eng, out, len(idx), self.d, self.ld)

    def copy_rows(self, idx_dev, n_out, spare=0):
        ld = _even(self.d + spare)
        out = self.eng.empty(max(n_out, 1) * ld)
        if n_out:
            self.eng.gather_rows(self.t, self.ld, idx_dev, n_out, self.d, out, ld)
        return DevMat(self.eng, out, n_out, self.d, ld)

    def with_col(self, vec):
        """``concat(x, y, axis=1)`` (model.py:320): writes ``vec`` (device, n) as column d."""
        if self.d < self.ld and not self.frozen:
            m = DevMat(self.eng, self.t, self.n, self.d + 1, self.ld)
        else:
            m = self.copy_rows(None, self.n, spare=8)
            m.d += 1
        self.eng.scatter_col(m.t, m.ld, self.d, None, vec, self.n)
        return m

    def to_host(self):
        if self.n == 0:
            return np.zeros((0, self.d))
        return self.t[: self.n * self.ld].reshape(self.n, self.ld)[:, : self.d].cpu().numpy()


class ObsBlock:
    """Observations a layer was conditioned on: inputs ``X`` (n x d), targets ``y``
    and the noise diagonal ``d = noise / w`` (all on the device)."""

    def __init__(self, X, y, d, n):
        self.X, self.y, self.d, self.n = X, y, d, int(n)


def _stack(eng, mats, spare=0):
    """Row-stack DevMats with equal ``d`` into one contiguous buffer."""
    mats = [m for m in mats if m is not None and m.n > 0]
    if len(mats) == 1 and spare == 0:
        return mats[0]
    d = mats[0].d if mats else 1
    ld = _even(d + spare)
    n = sum(m.n for m in mats)
    out = eng.empty(max(n, 1) * ld)
    r = 0
    for m in mats:
        assert m.d == d
        eng.gather_rows(m.t, m.ld, None, m.n, d, out, ld, dst_off=r * ld)
        r += m.n
    return DevMat(eng, out, n, d, ld)


def _cat(eng, vecs):
    """Concatenate device vectors (own copy kernel; torch only allocates)."""
    vecs = [v for v in vecs if v is not None and v.numel() > 0]
    if not vecs:
        return eng.zeros(1)
    if len(vecs) == 1:
        return vecs[0]
    out = eng.empty(sum(int(v.numel()) for v in vecs))
    off = 0
    for v in vecs:
        eng.gather_rows(v, 1, None, int(v.numel()), 1, out, 1, dst_off=off)
        off += int(v.numel())
    return out


_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


class GPAR:
    """Basic GPAR model (gpar/model.py:96-322) on the device engine.

    Args:
        replace (bool): condition on the predictive mean instead of the data.
        impute (bool): impute missing data with the predictive mean.
        x_ind (array, optional): inducing-point locations (VFE path).
        engine (:class:`Engine`, optional): device engine.
    """

    def __init__(self, replace=False, impute=False, x_ind=None, engine=None):
        self.replace = replace
        self.impute = impute
        self.layers = []
        self.sparse = x_ind is not None
        self.x_ind = None if x_ind is None else x_ind
        self._engine = engine

    @property
    def engine(self):
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    def copy(self):
        """New model with the same configuration and no layers (model.py:125-132)."""
        return GPAR(replace=self.replace, impute=self.impute, x_ind=self.x_ind, engine=self._engine)

    def add_layer(self, model_constructor):
        gpar = self.copy()
        gpar.layers = list(self.layers) + [model_constructor]
        return gpar

    # -- helpers ------------------------------------------------------------
    def _as_devmat(self, x, spare):
        if isinstance(x, DevMat):
            return DevMat(x.eng, x.t, x.n, x.d, x.ld, frozen=True)
        return DevMat.from_host(self.engine, x, spare=spare)

    def _factor(self, layer, xd, y_i, w_i, avail, ext=None, ext_d=None):
        """Observations of one layer (model.py:279-289) as a joint factorisation:
        stacked rows = [conditioning block of a posterior layer; available rows of
        this call; optional appended rows ``ext``]."""
        eng = self.engine
        blk = layer.block
        idx = np.flatnonzero(avail)
        if len(idx) == xd.n:
            Xa = xd
        else:
            Xa = xd.copy_rows(eng.to_device(idx, torch.int64), len(idx))
        y_a = eng.to_device(y_i[avail, 0])
        d_a = eng.to_device(layer.noise / w_i[avail])
        n_blk = 0 if blk is None else blk.n
        n_a = len(idx)
        n_ext = 0 if ext is None else ext.n
        if blk is None and ext is None:
            X = Xa
            d_all, y_all = d_a, y_a
        else:
            X = _stack(eng, [None if blk is None else blk.X, Xa, ext])
            d_all = _cat(eng, [None if blk is None else blk.d, d_a, ext_d])
            y_all = _cat(eng, [None if blk is None else blk.y, y_a])
        fac = Factor(eng, layer.spec, X.t, X.ld, d_all, y_all, n_blk + n_a, n_ext)
        fac.n_blk, fac.n_a = n_blk, n_a
        fac.new_block = ObsBlock(Xa, y_a, d_a, n_a)
        return fac

    def _update_inputs_dev(self, xd, y_i, avail, fac, sampled=None):
        """``_update_inputs`` (model.py:291-322) for truthy observations: builds the
        column fed to the next layer from data / posterior means / sampled values and
        appends it.  ``sampled``: device vector for the missing rows (sample_missing)."""
        eng = self.engine
        n_i = xd.n
        miss = ~avail
        n_a, n_m = int(avail.sum()), int(miss.sum())
        col = eng.empty(max(n_i, 1))
        both = self.impute and self.replace
        obs_est = self.replace
        if both:
            miss_mode = "est"
        elif sampled is not None:
            miss_mode = "est" if self.replace else "sampled"
        else:
            miss_mode = "est" if self.impute else "nan"
        idx_a = None if n_m == 0 else eng.to_device(np.flatnonzero(avail), torch.int64)
        # rows with observations
        if n_a:
            if obs_est:
                est = eng.empty(n_a)
                fac.mean_obs(est, fac.n_blk, fac.n_blk + n_a)
                eng.scatter_col(col, 1, 0, idx_a, est, n_a)
            else:
                eng.scatter_col(col, 1, 0, idx_a, fac.new_block.y, n_a)
        # rows without
        if n_m:
            idx_m = eng.to_device(np.flatnonzero(miss), torch.int64)
            if miss_mode == "est":
                Xm = xd.copy_rows(idx_m, n_m)
                est = eng.empty(n_m)
                fac.mean_at(Xm.t, Xm.ld, n_m, est)
                eng.scatter_col(col, 1, 0, idx_m, est, n_m)
            elif miss_mode == "sampled":
                eng.scatter_col(col, 1, 0, idx_m, sampled, n_m)
            else:
                eng.scatter_col(col, 1, 0, idx_m, eng.to_device(np.full(n_m, np.nan)), n_m)
        return xd.with_col(col)

    def _needs_factor(self, avail):
        return self.replace or (self.impute and not avail.all())

    # -- conditioning ---------------------------------------------------------
    def __or__(self, x_y_w):
        """Condition on data (model.py:148-176).  Returns a GPAR whose layers carry
        their observation blocks; the noise of each layer is unchanged."""
        x, y, w = x_y_w
        gpar = self.copy()
        if self.sparse:
            from .sparse import condition_sparse

            return condition_sparse(self, gpar, x, y, w)
        y = np.asarray(y, dtype=np.float64)
        w = np.asarray(w, dtype=np.float64)
        xd = self._as_devmat(x, spare=y.shape[1] + 1)
        for is_last, ((y_i, w_i, mask), model) in last(zip(per_output(y, w, keep=self.impute), self.layers)):
            xd = xd.take_rows(mask)
            layer = model()
            if not isinstance(layer, LayerModel):
                layer = layer[0]
            avail = ~np.isnan(y_i[:, 0])
            need = (not is_last) and self._needs_factor(avail)
            if need or layer.block is not None:
                fac = self._factor(layer, xd, y_i, w_i, avail)
                blk = ObsBlock(_stack(self.engine, [None if layer.block is None else layer.block.X, fac.new_block.X]),
                               fac.y, fac.d, fac.n_obs)
            else:
                fac = None
                idx = np.flatnonzero(avail)
                Xa = xd if len(idx) == xd.n else xd.copy_rows(self.engine.to_device(idx, torch.int64), len(idx))
                blk = ObsBlock(DevMat(self.engine, Xa.t, Xa.n, Xa.d, Xa.ld), self.engine.to_device(y_i[avail, 0]),
                               self.engine.to_device(layer.noise / w_i[avail]), len(idx))
            post = layer.conditioned(blk)
            gpar.layers.append(construct_model(post, layer.noise))
            if not is_last:
                if fac is not None:
                    xd = self._update_inputs_dev(xd, y_i, avail, fac)
                else:
                    xd = xd.with_col(self.engine.to_device(y_i[:, 0]))
            if fac is not None:
                fac.release()
        self.engine.check_infos()
        return gpar

    # -- logpdf -----------------------------------------------------------------
    def logpdf(self, x, y, w, only_last_layer=False, sample_missing=False, return_inputs=False, x_ind=None,
               outputs=None, normals=None, grad_out=None):
        """Log-density of ``y`` (model.py:178-243).  ``normals``: list of host arrays,
        one per layer that has missing rows, consumed when ``sample_missing``.  ``grad_out``: a dict
        that receives, for the LAST layer evaluated, the raw chain-rule sums of the gradient of its
        log-marginal w.r.t. its kernel spec and noise (``raw``, device tensor; ``layer``) -- what
        ``GPARRegressor.fit`` feeds to L-BFGS instead of the reference's autograd pass."""
        if self.sparse:
            from .sparse import logpdf_sparse

            return logpdf_sparse(self, x, y, w, only_last_layer, return_inputs, x_ind, outputs,
                                 sample_missing=sample_missing, normals=normals, grad_out=grad_out)
        eng = self.engine
        if not isinstance(y, dict):
            y = np.asarray(y, dtype=np.float64)
            w = np.asarray(w, dtype=np.float64)
        p = len(self.layers)
        xd = self._as_devmat(x, spare=p + 1)
        out2 = eng.zeros(2 * max(p, 1))
        counts = []
        normals = list(normals) if normals is not None else None
        li = -1
        for is_last, ((y_i, w_i, mask), model) in last(
            zip(per_output(y, w, keep=self.impute or sample_missing), self.layers), select=outputs
        ):
            li += 1
            xd = xd.take_rows(mask)
            layer = model()
            if not isinstance(layer, LayerModel):
                layer = layer[0]
            avail = ~np.isnan(y_i[:, 0])
            miss = ~avail
            want_lp = (not only_last_layer) or is_last
            do_sample = (not is_last) and sample_missing and miss.any()
            need_fac = want_lp or do_sample or ((not is_last) and self._needs_factor(avail))
            fac = None
            sampled = None
            if need_fac:
                ext = ext_d = None
                if do_sample:
                    idx_m = eng.to_device(np.flatnonzero(miss), torch.int64)
                    ext = xd.copy_rows(idx_m, int(miss.sum()))
                    ext_d = eng.to_device(layer.noise / w_i[miss])
                fac = self._factor(layer, xd, y_i, w_i, avail, ext=ext, ext_d=ext_d)
                if want_lp:
                    fac.logdet_quad(out2, 2 * li, fac.n_blk, fac.n_obs)
                    counts.append((li, fac.n_a))
                    if grad_out is not None and (is_last or grad_out.get("every_layer")) and fac.n_a > 0:
                        if fac.n_blk != 0 or fac.n_ext != 0:
                            raise NotImplementedError("gradients are implemented for prior layers only")
                        Ainv = eng.potri(fac.J, fac.ld, fac.n_obs, fac.ws)
                        dvec = eng.to_device(1.0 / w_i[avail])
                        raw = eng.gram_grad(layer.spec, fac.X, fac.ldx, fac.n_obs, fac.alpha(), Ainv, fac.ld, dvec)
                        if grad_out.get("every_layer"):
                            grad_out.setdefault("per_layer", {})[li] = (raw, layer)
                        else:
                            grad_out["raw"], grad_out["layer"] = raw, layer
                if do_sample:
                    n_m = ext.n
                    z = normals.pop(0) if normals is not None else eng.standard_normal_host(n_m)
                    Z = eng.to_device(np.asarray(z, dtype=np.float64).reshape(1, n_m))
                    mean = eng.empty(n_m)
                    fac.ext_mean(mean)
                    sampled = eng.empty(n_m)
                    fac.ext_sample(Z, 1, sampled, mean=mean)
            if not is_last:
                if fac is not None:
                    xd = self._update_inputs_dev(xd, y_i, avail, fac, sampled=sampled)
                else:
                    xd = xd.with_col(eng.to_device(y_i[:, 0]))
            if fac is not None:
                fac.release()
        eng.check_infos()
        if return_inputs:
            return xd, x_ind
        vals = out2.cpu().numpy()
        total = 0.0
        for slot, n_a in counts:
            if n_a > 0:
                total += -0.5 * (vals[2 * slot] + n_a * math.log(2.0 * math.pi) + vals[2 * slot + 1])
        return float(total)

    # -- sampling -----------------------------------------------------------------
    def sample(self, x, w, latent=False, num_samples=1, normals=None, train=None, return_device=False,
               generator=None):
        """Ancestral samples at ``x`` (model.py:245-277) for ``num_samples`` independent
        chains at once.

        ``normals``: ``None`` (draw on the device) or a dict with ``"Z"`` of shape
        (S, p, n) and, when ``latent``, ``"Z2"`` (S, p, n) -- the injected standard
        normals in the reference's draw order (layer-major; latent draw first, then
        the noise draw).  ``train=(x, y, w)`` fuses conditioning with sampling: every
        layer is factored once jointly over [training rows; test rows].  ``generator``: torch
        CUDA generator for the device normals (chain sharding gives every rank its own stream).
        Returns an array (S, n, p).
        """
        if self.sparse:
            from .sparse import sample_sparse

            return sample_sparse(self, x, w, latent, num_samples, normals, train, return_device, generator)
        eng = self.engine
        S = int(num_samples)
        p = len(self.layers)
        w = np.asarray(w, dtype=np.float64)
        xs = self._as_devmat(x, spare=p + 1)
        ns = xs.n
        if normals is None:
            Zall = torch.randn(S, p, ns, dtype=F64, device=eng.device, generator=generator)
            Z2all = torch.randn(S, p, ns, dtype=F64, device=eng.device, generator=generator) if latent else None
        else:
            Zall = eng.to_device(np.asarray(normals["Z"], dtype=np.float64).reshape(S, p, ns))
            Z2all = eng.to_device(np.asarray(normals["Z2"], dtype=np.float64).reshape(S, p, ns)) if latent else None
        out = eng.empty(max(S * ns * p, 1))  # (S*ns) x p row-major
        shared = True  # all chains still share their inputs
        xs_all = None  # (S*ns) x d inputs once chains diverge

        if train is not None:
            xt, yt, wt = train
            yt = np.asarray(yt, dtype=np.float64)
            wt = np.asarray(wt, dtype=np.float64)
            xd = self._as_devmat(xt, spare=p + 1)
            train_iter = per_output(yt, wt, keep=self.impute)
        else:
            train_iter = None

        for i, (is_last, model) in enumerate(last(self.layers)):
            layer = model()
            if not isinstance(layer, LayerModel):
                layer = layer[0]
            noise = layer.noise
            sd = eng.to_device(np.sqrt(noise / w[:, i])) if latent else None
            d_s = eng.zeros(max(ns, 1)) if latent else eng.to_device(noise / w[:, i])
            Zi = Zall[:, i, :].contiguous()
            Z2i = Z2all[:, i, :].contiguous() if latent else None

            # training side of this layer
            fac_obs = None
            if train_iter is not None:
                y_i, w_i, mask = next(train_iter)
                xd = xd.take_rows(mask)
                avail = ~np.isnan(y_i[:, 0])
            f_col = eng.empty(max(S * ns, 1))  # recorded sample (latent f or y)
            y_col = f_col

            owned = True  # the factor of this layer is dropped at the end of the iteration
            if shared:
                # one joint factorisation over [block / training rows; test rows]
                if train_iter is not None:
                    fac = self._factor(layer, xd, y_i, w_i, avail, ext=xs, ext_d=d_s)
                elif layer.block is not None:
                    blk = layer.block
                    X = _stack(eng, [blk.X, xs])
                    fac = Factor(eng, layer.spec, X.t, X.ld, _cat(eng, [blk.d, d_s[:ns]]), blk.y, blk.n, ns)
                    fac.n_blk, fac.n_a = blk.n, 0
                else:
                    fac = Factor(eng, layer.spec, xs.t, xs.ld, d_s, eng.zeros(1), 0, ns)
                    fac.n_blk = fac.n_a = 0
                mean = eng.empty(max(ns, 1))
                fac.ext_mean(mean)
                fac.ext_sample(Zi, S, f_col, mean=mean)
                if latent:
                    y_col = eng.empty(max(S * ns, 1))
                    fac.ext_sample(Zi, S, y_col, mean=mean, sd=sd, Z2=Z2i)
                fac_obs = fac
            else:
                # chains have diverged: factor the observations once, then batch the chains
                if train_iter is not None:
                    fac = self._factor(layer, xd, y_i, w_i, avail)
                elif layer.block is not None:
                    blk = layer.block
                    fac = getattr(blk, "factor", None)
                    if fac is None:
                        fac = Factor(eng, layer.spec, blk.X.t, blk.X.ld, blk.d, blk.y, blk.n, 0)
                        fac.n_blk, fac.n_a = blk.n, 0
                        blk.factor = fac
                    owned = False  # cached on the observation block for later calls
                else:
                    fac = None
                f_col, y_col = self._chains_layer(layer, fac, xs_all, S, ns, d_s, sd, Zi, Z2i, latent)
                fac_obs = fac

            eng.scatter_col(out, p, i, None, f_col, S * ns)

            if not is_last:
                # inputs of the next layer (model.py:273-275, obs=None => estimate = f.mean)
                if self.replace:
                    if shared:
                        xs = xs.with_col(mean)
                    else:
                        xs_all = xs_all.with_col(self._chain_means)
                else:
                    if shared:
                        # broadcast x over the chains, then append each chain's own sample
                        rep = np.tile(np.arange(ns, dtype=np.int64), S)
                        xs_all = xs.copy_rows(eng.to_device(rep, torch.int64), S * ns, spare=p + 1)
                        shared = False
                    xs_all = xs_all.with_col(y_col)
                if train_iter is not None:
                    if self._needs_factor(avail):
                        xd = self._update_inputs_dev(xd, y_i, avail, fac_obs)
                    else:
                        xd = xd.with_col(eng.to_device(y_i[:, 0]))
            if owned and fac_obs is not None:
                fac_obs.release()
        eng.check_infos()
        if return_device:
            return out.reshape(S, ns, p)
        return out.reshape(S, ns, p).cpu().numpy()

    def _chains_layer(self, layer, fac, xs_all, S, ns, d_s, sd, Zi, Z2i, latent):
        """One layer for S diverged chains (the reference runs them one after the other,
        regression.py:557-564; their arithmetic is independent): per chain s,
        W_s = K(x*_s, X_a) L^-T (one TRSM over the rows of all chains of a pass), Sigma_s = K_** - W_s W_s^T
        (batched SYRK), C_s = chol (batched), draws.  The chains are processed in passes sized from
        the free device memory (Engine.chain_chunk): at C5 (S = 256, n* = 2048, n = 32768) the
        cross-covariance of all chains at once would be 137 GB."""
        eng = self.engine
        spec = layer.spec
        N = S * ns
        ldc = _even(max(ns, 2))
        ldx = xs_all.ld
        has_obs = fac is not None and fac.n_obs > 0
        n_a, ld = (fac.n_obs, fac.ld) if has_obs else (0, 0)
        mean_all = eng.zeros(max(N, 1))
        f_col = eng.empty(max(N, 1))
        y_col = eng.empty(max(N, 1)) if latent else f_col
        tiles = -(-ns // 128)
        per_chain = 8 * ns * (ldc + ld) + eng.lib.gpar_potrf_workspace_bytes(ns, 0, 2) // 2
        chunk = eng.chain_chunk(S, per_chain, tiles)
        for c0 in range(0, S, chunk):
            B = min(S, c0 + chunk) - c0
            r0 = c0 * ns
            Xc = xs_all.t[r0 * ldx:]
            mean_c = mean_all[r0:]
            Cs = eng.empty(B * ns * ldc)
            eng.gram_batched(spec, Xc, ldx, ns, ns * ldx, Cs, ldc, ns * ldc, B, diag=d_s, strideD=0)
            if has_obs:
                E = eng.empty(B * ns * ld)
                eng.gram(spec, Xc, ldx, B * ns, E, ld, Y=fac.X, ldy=fac.ldx, ny=n_a, lower_only=False)
                eng.trsm_rows(fac.J, ld, n_a, fac.ws, E, ld, B * ns)
                eng.gemv(E, ld, B * ns, n_a, fac.u, mean_c)
                eng.syrk_sub(Cs, ldc, ns, E, ld, n_a, batch=B, strideC=ns * ldc, strideW=ns * ld)
                del E
            eng.potrf(Cs, ldc, ns, batch=B, strideA=ns * ldc)
            eng.sample_affine(Cs, ldc, ns, Zi[c0:c0 + B], f_col[r0:], 1, batch=B, strideC=ns * ldc, mean=mean_c)
            if latent:
                eng.sample_affine(Cs, ldc, ns, Zi[c0:c0 + B], y_col[r0:], 1, batch=B, strideC=ns * ldc, mean=mean_c,
                                  sd=sd, Z2=Z2i[c0:c0 + B], strideSd=0)
            del Cs
        self._chain_means = mean_all
        return f_col, y_col

    # -- reference-shaped helper kept for the known-answer tests ----------------------
    def _update_inputs(self, x, x_ind, y, f, obs):
        """Host-array front end of ``_update_inputs`` (model.py:291-322).  ``f`` is a
        :class:`LayerModel`; ``obs`` is ``None`` (prior mean 0) or a tuple
        ``(x_obs, y_obs, noise_vec)`` of dense observations."""
        eng = self.engine
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        available = ~np.isnan(y[:, 0])
        fac = None
        if obs is not None:
            xo, yo, no = obs
            Xo = DevMat.from_host(eng, xo)
            fac = Factor(eng, f.spec, Xo.t, Xo.ld, eng.to_device(np.broadcast_to(no, (Xo.n,)).copy()),
                         eng.to_device(np.asarray(yo, dtype=np.float64).reshape(-1)), Xo.n, 0)

        def estimate(x_):
            x_ = np.asarray(x_, dtype=np.float64)
            if fac is None or x_.shape[0] == 0:
                return np.zeros((x_.shape[0], 1))
            Xq = DevMat.from_host(eng, x_)
            out = eng.empty(Xq.n)
            fac.mean_at(Xq.t, Xq.ld, Xq.n, out)
            return out.cpu().numpy()[:, None]

        if self.sparse:
            x_ind = np.concatenate([np.asarray(x_ind, dtype=np.float64), estimate(x_ind)], axis=1)
        if self.impute and self.replace:
            y = estimate(x)
        else:
            if self.impute and np.any(~available):
                y = merge(y, estimate(x[~available]), ~available)
            if self.replace and np.any(available):
                y = merge(y, estimate(x[available]), available)
        return np.concatenate([x, y], axis=1), x_ind
