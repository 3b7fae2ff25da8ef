"""Reference op sequence of the GPAR dense hot path on torch tensors.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference (wesselb/gpar) runs its arithmetic through lab -> torch: `torch.linalg.cholesky`,
`torch.linalg.solve_triangular`, `matmul`, `exp` on fp64 CPU tensors (gpar/regression.py:314, 62-64).
This module restates that op sequence on torch tensors of a chosen device:

* ``device="cpu"``  -- what the reference dispatches to (MKL LAPACK/BLAS on the host cores): the timed
  CPU baseline of ``bench.py`` (``--impl reference`` and ``cpu_baseline``; kind "port": the reference's
  own stack stheno/lab/matrix/mlkernels is not installable here).
* ``device="cuda"`` -- the secondary bar of SURVEY.md 2a / BASELINE.md 3: the same restatement on torch CUDA
  tensors (cuSOLVER potrf, cuBLAS trsm/gemm), i.e. the library-call GPU path the reference would get "for
  free".  Library code, timed by ``bench.py`` only; never on the product path.

It follows ``oracle/gpar_oracle.py`` (the pinned numpy restatement) function by function and is checked
against it in ``tests/test_oracle_torch_ref.py``.  Differences, all reference-faithful [UPSTREAM-RECALL]:
squared distances use lab's ``pw_dists2`` form (one column: (a - b^T)^2; otherwise |a|^2 + |b|^2 - 2 a b^T,
not clipped) instead of the oracle's direct differences (agree to ~1e-13 on the test data).
Only dense observations (``Obs``) are restated; inducing points stay with the numpy oracle.

Control flow cites gpar/model.py and gpar/regression.py like the oracle does.  Masks / indices are
computed with the oracle's own (golden-vector-pinned) host functions.
"""
import math
import time

import numpy as np
import torch

from .gpar_oracle import EPSILON, determine_indices, last, model_terms, per_output, _Vars, _uprank  # noqa: F401

__all__ = ["TorchRegressor", "kernel_matrix"]

F64 = torch.float64


def _t(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=F64)
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device=device)


def _pw_dists2(a, b):
    """lab ``pw_dists2`` [UPSTREAM-RECALL]: one column -> (a - b^T)^2, else norm expansion (not clipped)."""
    if a.shape[1] == 1 and b.shape[1] == 1:
        return (a - b.T) ** 2
    na = torch.sum(a * a, dim=1)[:, None]
    nb = torch.sum(b * b, dim=1)[None, :]
    return na + nb - 2.0 * (a @ b.T)


def kernel_matrix(terms, X, Y):
    """K[i, j] = sum_t k_t(X[i], Y[j]) for the kernel family of gpar/regression.py:92-180 (mlkernels
    EQ / RQ / Linear / const / locally periodic with ``.stretch`` and ``.select``)."""
    K = torch.zeros((X.shape[0], Y.shape[0]), dtype=F64, device=X.device)
    for t in terms:
        kind = t["type"]
        var = float(t.get("variance", 1.0))
        if kind == "const":
            K += var
            continue
        cols = list(t.get("cols", []))
        inv = torch.as_tensor(1.0 / np.asarray(t["scales"], dtype=np.float64), device=X.device)
        if kind in ("eq", "rq", "linear"):
            xs, ys = X[:, cols] * inv[None, :], Y[:, cols] * inv[None, :]
            if kind == "eq":
                K += var * torch.exp(-0.5 * _pw_dists2(xs, ys))
            elif kind == "rq":
                alpha = float(t["alpha"])
                K += var * (1.0 + _pw_dists2(xs, ys) / (2.0 * alpha)) ** (-alpha)
            else:
                K += var * (xs @ ys.T)
        elif kind == "periodic":
            m = len(cols)
            freq = torch.as_tensor(2.0 * np.pi / np.asarray(t["periods"], dtype=np.float64), device=X.device)
            inv_decay = torch.as_tensor(1.0 / np.asarray(t["decays"], dtype=np.float64), device=X.device)
            ax, ay = X[:, cols] * freq[None, :], Y[:, cols] * freq[None, :]
            ux = torch.cat([torch.sin(ax), torch.cos(ax)], dim=1) * inv[None, : 2 * m]
            uy = torch.cat([torch.sin(ay), torch.cos(ay)], dim=1) * inv[None, : 2 * m]
            K += var * torch.exp(-0.5 * _pw_dists2(ux, uy)) * torch.exp(
                -0.5 * _pw_dists2(X[:, cols] * inv_decay[None, :], Y[:, cols] * inv_decay[None, :]))
        else:
            raise ValueError(f"unknown kernel term {kind!r}")
    return K


def _chol(a):
    """matrix.cholesky(Dense): chol(a + eps I), lower -> torch.linalg.cholesky (MKL dpotrf / cuSOLVER)."""
    a = a.clone()
    a.diagonal().add_(EPSILON)
    return torch.linalg.cholesky(a)


def _solve_lower(L, b):
    return torch.linalg.solve_triangular(L, b, upper=False)


class GP:
    """Zero-mean prior (term list) or posterior ``f | obs`` -- oracle.GP on torch."""

    def __init__(self, terms=None, parent=None, obs=None):
        self.terms, self.parent, self.obs = terms, parent, obs

    def mean(self, x):
        if self.parent is None:
            return torch.zeros((x.shape[0], 1), dtype=F64, device=x.device)
        return self.parent.mean(x) + self.obs.posterior_mean_correction(self.parent, x)

    def kernel(self, x, y):
        if self.parent is None:
            return kernel_matrix(self.terms, x, y)
        return self.parent.kernel(x, y) - self.obs.posterior_kernel_correction(self.parent, x, y)

    def __call__(self, x, noise=None):
        return FDD(self, x, noise)

    def __or__(self, obs):
        return GP(parent=self, obs=obs)

    def logpdf(self, obs):
        return obs.logpdf(self)


class FDD:
    def __init__(self, f, x, noise=None):
        self.f, self.x = f, x
        n = x.shape[0]
        if noise is None:
            self.noise = torch.zeros(n, dtype=F64, device=x.device)
        else:
            self.noise = torch.broadcast_to(torch.as_tensor(noise, dtype=F64, device=x.device), (n,)).clone()

    def var(self):
        K = self.f.kernel(self.x, self.x)
        K.diagonal().add_(self.noise)
        return K

    def sample(self, normals):
        """Normal.sample: mean + chol(var + eps I) z (joint draw)."""
        n = self.x.shape[0]
        if n == 0:
            return torch.zeros((0, 1), dtype=F64, device=self.x.device)
        L = _chol(self.var())
        return self.f.mean(self.x) + L @ normals(n, self.x.device)


class Obs:
    """Dense observations; posterior per SURVEY 8(c)-3 with a fresh triangular solve per mean / kernel
    call (no alpha cache) -- the reference-faithful op sequence; the Cholesky is cached on the object."""

    def __init__(self, fdd, y):
        self.fdd, self.y, self._L = fdd, y.reshape(-1, 1), None

    def __bool__(self):
        return True

    def _chol_of(self, f):
        if self._L is None:
            K = f.kernel(self.fdd.x, self.fdd.x)
            K.diagonal().add_(self.fdd.noise)
            self._L = _chol(K)
        return self._L

    def logpdf(self, f):
        n = self.y.shape[0]
        if n == 0:
            return 0.0
        L = self._chol_of(f)
        u = _solve_lower(L, self.y - f.mean(self.fdd.x))
        logdet = 2.0 * torch.sum(torch.log(torch.diagonal(L)))
        return float(-0.5 * (logdet + n * math.log(2.0 * math.pi) + torch.sum(u * u)))

    def posterior_mean_correction(self, f, x):
        if self.y.shape[0] == 0:
            return torch.zeros((x.shape[0], 1), dtype=F64, device=x.device)
        L = self._chol_of(f)
        A = _solve_lower(L, f.kernel(self.fdd.x, x))
        b = _solve_lower(L, self.y - f.mean(self.fdd.x))
        return A.T @ b

    def posterior_kernel_correction(self, f, x, y):
        if self.y.shape[0] == 0:
            return torch.zeros((x.shape[0], y.shape[0]), dtype=F64, device=x.device)
        L = self._chol_of(f)
        A = _solve_lower(L, f.kernel(self.fdd.x, x))
        Bm = A if y is x else _solve_lower(L, f.kernel(self.fdd.x, y))
        return A.T @ Bm


class TorchNormals:
    """Replays injected standard normals in the reference's draw order (oracle.Normals), or draws."""

    def __init__(self, queue=None, seed=0):
        self.queue = None if queue is None else list(queue)
        self.pos = 0
        self.gen = torch.Generator().manual_seed(seed)

    def __call__(self, n, device):
        if self.queue is not None:
            z = self.queue[self.pos]
            self.pos += 1
            return _t(np.asarray(z).reshape(n, 1), device)
        return torch.randn(n, 1, dtype=F64, generator=self.gen).to(device)


def _merge(x, updates, mask):
    """gpar/model.py:14-44 on torch (mask: host bool array)."""
    out = x.clone()
    out[torch.as_tensor(mask, device=x.device)] = updates
    return out


class GPAR:
    """gpar/model.py:96-322 on torch tensors; masks on the host through the oracle's ``per_output``.
    ``tick(phase, layer)`` (optional) is called after every layer of every loop -- bench.py uses it to
    time the phases and to stop a run that would exceed its wall-clock budget."""

    def __init__(self, replace=False, impute=False, device="cpu", tick=None):
        self.replace, self.impute, self.device = replace, impute, device
        self.layers = []
        self.tick = tick or (lambda phase, layer: None)

    def copy(self):
        return GPAR(self.replace, self.impute, self.device, self.tick)

    def add_layer(self, model_constructor):
        g = self.copy()
        g.layers = list(self.layers) + [model_constructor]
        return g

    def _obs(self, x, y, w, f, noise):
        avail = ~np.isnan(y[:, 0])
        m = torch.as_tensor(avail, device=self.device)
        return Obs(f(x[m], _t(noise / w[avail], self.device)), _t(y[avail], self.device))

    def _update_inputs(self, x, y, f, obs):
        """model.py:291-322 (dense)."""
        avail = ~np.isnan(y[:, 0])
        yt = _t(y, self.device)

        def estimate(x_):
            return (f | obs).mean(x_) if obs else f.mean(x_)

        if self.impute and self.replace:
            yt = estimate(x)
        else:
            if self.impute and np.any(~avail):
                yt = _merge(yt, estimate(x[torch.as_tensor(~avail, device=self.device)]), ~avail)
            if self.replace and np.any(avail):
                yt = _merge(yt, estimate(x[torch.as_tensor(avail, device=self.device)]), avail)
        return torch.cat([x, yt], dim=1)

    def __or__(self, x_y_w):
        x, y, w = x_y_w
        gpar = self.copy()
        i = 0
        for is_last, ((y_i, w_i, mask), model) in last(zip(per_output(y, w, keep=self.impute), self.layers)):
            x = x[torch.as_tensor(mask, device=self.device)]
            f, noise = model()
            obs = self._obs(x, y_i, w_i, f, noise)
            gpar.layers.append((lambda post, nz: (lambda: (post, nz)))(f | obs, noise))
            if not is_last:
                x = self._update_inputs(x, y_i, f, obs)
            self.tick("condition", i)
            i += 1
        return gpar

    def logpdf(self, x, y, w):
        total, i = 0.0, 0
        for is_last, ((y_i, w_i, mask), model) in last(zip(per_output(y, w, keep=self.impute), self.layers)):
            x = x[torch.as_tensor(mask, device=self.device)]
            f, noise = model()
            obs = self._obs(x, y_i, w_i, f, noise)
            total = total + f.logpdf(obs)
            if not is_last:
                x = self._update_inputs(x, y_i, f, obs)
            self.tick("logpdf", i)
            i += 1
        return total

    def sample(self, x, w, latent=False, normals=None):
        """model.py:245-277."""
        sample = torch.zeros((x.shape[0], 0), dtype=F64, device=self.device)
        for i, (is_last, model) in enumerate(last(self.layers)):
            f, noise = model()
            if latent:
                f_sample = f(x).sample(normals)
                stds = _t(np.sqrt(noise / w[:, i : i + 1]), self.device)
                y_sample = f_sample + stds * normals(f_sample.shape[0], self.device)
                sample = torch.cat([sample, f_sample], dim=1)
            else:
                y_sample = f(x, _t(noise / w[:, i], self.device)).sample(normals)
                sample = torch.cat([sample, y_sample], dim=1)
            if not is_last:
                # obs = None: estimate = f.mean (the posterior mean, f is already conditioned)
                yt = y_sample
                if self.replace:  # no NaNs in a sample: impute never fires, replace overwrites every row
                    yt = f.mean(x)
                x = torch.cat([x, yt], dim=1)
            self.tick("sample", i)
        return sample


class TorchRegressor:
    """gpar/regression.py:200-597 (condition / logpdf / sample / predict; dense layers) on torch."""

    def __init__(self, device="cpu", tick=None, replace=False, impute=True, scale=1.0, scale_tie=False, per=False,
                 per_period=1.0, per_scale=1.0, per_decay=10.0, input_linear=False, input_linear_scale=100.0,
                 linear=True, linear_scale=100.0, nonlinear=False, nonlinear_scale=1.0, rq=False, markov=None,
                 noise=0.1, normalise_y=True):
        self.device, self.tick = torch.device(device), tick
        self.replace, self.impute = replace, impute
        self.model_config = dict(scale=scale, scale_tie=scale_tie, per=per, per_period=per_period,
                                 per_scale=per_scale, per_decay=per_decay, input_linear=input_linear,
                                 input_linear_scale=input_linear_scale, linear=linear, linear_scale=linear_scale,
                                 nonlinear=nonlinear, nonlinear_scale=nonlinear_scale, rq=rq, markov=markov,
                                 noise=noise)
        self.vs = _Vars()
        self.is_conditioned = False
        self.normalise_y = normalise_y
        self._unnormalise_y, self._normalise_y = (lambda y: y), (lambda y: y)

    def _construct_gpar(self, m, p):
        gpar = GPAR(self.replace, self.impute, self.device, self.tick)
        for pi in range(p):
            def model(pi=pi):
                terms, noise = model_terms(self.vs, m, pi, **self.model_config)
                return GP(terms), noise
            gpar = gpar.add_layer(model)
        return gpar

    def condition(self, x, y, w=None):
        """regression.py:339-389 (host statistics, like the oracle)."""
        self.x, self.y = _uprank(x), _uprank(y)
        self.w = np.ones_like(self.y) if w is None else _uprank(w)
        self.n, self.m = self.x.shape
        self.p = self.y.shape[1]
        if self.normalise_y:
            means = np.array([np.mean(self.y[~np.isnan(self.y[:, i]), i]) for i in range(self.p)])[None, :]
            stds = np.array([np.std(self.y[~np.isnan(self.y[:, i]), i]) for i in range(self.p)])[None, :]
            stds = np.where(stds > 0, stds, 1.0)
            self._normalise_y = lambda y_: (y_ - means) / stds
            self._unnormalise_y = lambda y_: y_ * stds + means
            self.y = self._normalise_y(self.y)
        self.is_conditioned = True

    def logpdf(self, x, y, w=None):
        x = _uprank(x)
        y = self._unnormalise_y(_uprank(y))  # quirk Q1
        w = np.ones_like(y) if w is None else _uprank(w)
        gpar = self._construct_gpar(x.shape[1], y.shape[1])
        return gpar.logpdf(_t(x, self.device), y, w)

    def conditioned(self):
        """``gpar | (x, y, w)`` (regression.py:546-547); every Cholesky is forced here so that the
        conditioning cost is not attributed to the first chain (stheno conditions lazily)."""
        gpar = self._construct_gpar(self.m, self.p) | (_t(self.x, self.device), self.y, self.w)
        for model in gpar.layers:
            f, _ = model()
            f.obs._chol_of(f.parent)
        return gpar

    def sample_chain(self, gpar, x, w=None, latent=False, normals=None):
        """One chain of regression.py:557-563 (un-normalised, on the host)."""
        x = _uprank(x)
        w = np.ones((x.shape[0], self.p)) if w is None else _uprank(w)
        s = gpar.sample(_t(x, self.device), w, latent=latent, normals=normals or TorchNormals())
        return self._unnormalise_y(s.cpu().numpy())

    def predict(self, x, num_samples=100, latent=False, normals=None):
        gpar = self.conditioned()
        normals = normals or TorchNormals()
        samples = [self.sample_chain(gpar, x, latent=latent, normals=normals) for _ in range(num_samples)]
        return np.mean(samples, axis=0)


def timed_step(reg_kw, data, S, device="cpu", budget_s=None, sync=None):
    """One (condition, logpdf, predict) pass of the reference op sequence with per-phase wall times.

    Runs the full logpdf and the full conditioning; then chains one after the other until all ``S`` ran or
    the wall-clock ``budget_s`` would be exceeded by the next one.  Returns a dict with the measured seconds
    per phase, the number of chains actually run and the results (logpdf, mean over the chains run).
    ``sync``: callable that drains the device (CUDA) before a clock is read."""
    sync = sync or (lambda: None)
    t_start = time.perf_counter()
    reg = TorchRegressor(device=device, **reg_kw)
    reg.condition(data["x"], data["y"])
    sync()
    t0 = time.perf_counter()
    lp = reg.logpdf(data["x"], data["y"])
    sync()
    t1 = time.perf_counter()
    gpar = reg.conditioned()
    sync()
    t2 = time.perf_counter()
    p = data["y"].shape[1]
    chains, t_chain, acc = 0, [], None
    for s in range(S):
        if budget_s is not None and chains >= 1:
            if (time.perf_counter() - t_start) + float(np.mean(t_chain)) > budget_s:
                break
        c0 = time.perf_counter()
        smp = reg.sample_chain(gpar, data["xs"], normals=TorchNormals(queue=[data["Z"][s, i] for i in range(p)]))
        sync()
        t_chain.append(time.perf_counter() - c0)
        acc = smp if acc is None else acc + smp
        chains += 1
    return {"t_logpdf": t1 - t0, "t_condition": t2 - t1, "t_chains": float(np.sum(t_chain)), "chains": chains,
            "t_chain_mean": float(np.mean(t_chain)), "logpdf": float(lp), "mean": acc / max(chains, 1),
            "t_total": time.perf_counter() - t_start}
