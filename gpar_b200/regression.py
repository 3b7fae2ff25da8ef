"""GPARRegressor with the API of gpar/regression.py:200-597, running the per-layer
GP arithmetic on the B200 engine.  Host code here is data preparation only
(transforms, normalisation, hyper-parameter bookkeeping, the optimiser loop)."""
import numpy as np
from scipy.optimize import minimize

from .model import GPAR, per_output
from .spec import LayerModel, Vars, model_terms, named_gradients

__all__ = ["GPARRegressor", "log_transform", "squishing_transform"]

#: Log transform for the data (regression.py:22).
log_transform = (np.log, np.exp)

#: Squishing transform for the data (regression.py:25-28).
squishing_transform = (
    lambda x: np.sign(x) * np.log(1 + np.abs(x)),
    lambda x: np.sign(x) * (np.exp(np.abs(x)) - 1),
)


def _identity(x):
    return x


def _transform_kind(untransform):
    """Enum tag of an un-transform the device knows (include/gpar_b200.h GPAR_TRANSFORM_*), else None."""
    if untransform is _identity:
        return 0
    if untransform is log_transform[1]:
        return 1
    if untransform is squishing_transform[1]:
        return 2
    return None


def _uprank(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        a = a[None]
    return a[:, None] if a.ndim == 1 else a


def _construct_gpar(reg, vs, m, p):
    """Layer-by-layer construction (regression.py:185-190); each constructor re-reads
    the hyper-parameters from ``vs`` when called, like ``_model_generator``'s closure."""
    gpar = GPAR(replace=reg.replace, impute=reg.impute, x_ind=reg.x_ind, engine=reg._engine)
    for pi in range(p):
        def model(pi=pi):
            terms, noise = model_terms(vs, m, pi, **reg.model_config)
            return LayerModel(terms, noise)
        gpar = gpar.add_layer(model)
    return gpar


def _init_weights(w, y):
    return np.ones_like(y) if w is None else _uprank(w)


class GPARRegressor:
    """GPAR regressor; arguments, attributes and error behaviour follow
    gpar/regression.py:200-326 (code defaults win over the docstring, quirk Q2).

    Extra args:
        engine: :class:`gpar_b200.engine.Engine` to run on (default: shared engine
            on the current CUDA device).
    """

    def __init__(self, replace=False, impute=True, scale=1.0, scale_tie=False, per=False, per_period=1.0,
                 per_scale=1.0, per_decay=10.0, input_linear=False, input_linear_scale=100.0, linear=True,
                 linear_scale=100.0, nonlinear=False, nonlinear_scale=1.0, rq=False, markov=None, noise=0.1,
                 x_ind=None, normalise_y=True, transform_y=(_identity, _identity), engine=None):
        self.replace = replace
        self.impute = impute
        self.sparse = x_ind is not None
        self.x_ind = None if x_ind is None else _uprank(x_ind)
        self.model_config = {
            "scale": scale, "scale_tie": scale_tie, "per": per, "per_period": per_period, "per_scale": per_scale,
            "per_decay": per_decay, "input_linear": input_linear, "input_linear_scale": input_linear_scale,
            "linear": linear, "linear_scale": linear_scale, "nonlinear": nonlinear,
            "nonlinear_scale": nonlinear_scale, "rq": rq, "markov": markov, "noise": noise,
        }
        self.vs = Vars()
        self.is_conditioned = False
        self.x = self.y = self.w = None
        self.n = self.m = self.p = None
        self.normalise_y = normalise_y
        self._unnormalise_y, self._normalise_y = _identity, _identity
        self._norm = None  # (means, stds) per output once conditioned with normalise_y
        self._transform_y, self._untransform_y = transform_y
        self._engine = engine

    def get_variables(self):
        """Dictionary of all hyper-parameters (regression.py:328-337)."""
        return {name: np.array(self.vs[name]) for name in self.vs.names}

    def condition(self, x, y, w=None):
        """Store (transformed, normalised) data; no GP arithmetic (regression.py:339-389).
        The per-output std is the population std (ddof = 0)."""
        self.x = _uprank(x)
        self.y = self._transform_y(_uprank(y))
        self.w = _init_weights(w, self.y)
        self.n, self.m = self.x.shape
        self.p = self.y.shape[1]
        self._check_limits(self.m, self.p)
        if self.normalise_y:
            means, stds = [], []
            for i in range(self.p):
                y_i = self.y[~np.isnan(self.y[:, i]), i]
                means.append(np.mean(y_i))
                std = np.std(y_i)
                stds.append(std if std > 0 else 1.0)
            means, stds = np.array(means)[None, :], np.array(stds)[None, :]
            self._normalise_y = lambda y_: (y_ - means) / stds
            self._unnormalise_y = lambda y_: y_ * stds + means
            self._norm = (means.reshape(-1).copy(), stds.reshape(-1).copy())
            self.y = self._normalise_y(self.y)
        else:
            self._unnormalise_y, self._normalise_y, self._norm = _identity, _identity, None
        self.is_conditioned = True

    def _check_limits(self, m, p):
        """The device kernel spec is a fixed-size POD (include/gpar_b200.h: GPAR_MAX_FEATS = 96 features,
        GPAR_MAX_TERMS = 8 terms) and the Gram kernels stage 64-row input tiles in shared memory (about 200
        input columns).  The reference has no such limits: fail here, with the numbers, rather than deep inside a
        layer constructor."""
        from ._lib import MAX_FEATS

        cfg = self.model_config
        markov = cfg["markov"]
        p_num = (p - 1) if markov is None else min(p - 1, markov)
        feats = m + (3 * m if cfg["per"] else 0) + (m if cfg["input_linear"] else 0)
        feats += (p_num if cfg["linear"] else 0) + (p_num if cfg["nonlinear"] else 0)
        if feats > MAX_FEATS:
            raise ValueError(f"the last layer's kernel needs {feats} input features (m={m}, previous outputs used="
                             f"{p_num}); the device kernel spec holds {MAX_FEATS}: set markov=k to cap the number of "
                             "previous outputs per layer")
        if m + p > 200:
            raise ValueError(f"{m + p} input columns (m + p) exceed what the Gram kernels stage in shared memory (~200)")

    def fit(self, x, y, w=None, greedy=False, fix=True, **kw_args):
        """Layer-wise maximum likelihood (regression.py:391-459).  ``iters`` and other keyword arguments go to
        the L-BFGS-B driver.  Gradients are analytic (device kernels) whenever the inputs of the
        layers being optimised do not depend on the hyper-parameters being optimised: the default layer-wise
        objective (``fix=True``; dense layers: ``gpar_potri`` + ``gpar_gram_grad``, inducing-point layers: the
        VFE weights through ``gpar_gram_wgrad``), and the joint objective (``fix=False``) when neither
        ``replace`` nor imputation feeds posterior means into later layers.  Otherwise (joint objective with
        ``replace`` / imputed rows / inducing points, where the reference backpropagates through the
        posterior means) finite differences of the device log-marginal are used."""
        self.condition(x, y, w)
        if greedy:
            raise NotImplementedError("Greedy search is not implemented yet.")
        y_cached = {k: list(per_output(self.y, self.w, keep=k)) for k in [True, False]}
        iters = int(kw_args.pop("iters", 1000))
        want_analytic = bool(kw_args.pop("analytic", True))
        kw_args.pop("trace", None)  # varz option without a scipy counterpart
        from ._lib import GparError

        for pi in range(self.p):
            if fix:
                gpar = _construct_gpar(self, self.vs, self.m, pi + 1)
                fixed_x, fixed_x_ind = gpar.logpdf(self.x, y_cached, None, only_last_layer=True,
                                                   outputs=list(range(pi)), return_inputs=True)
            # Instantiate the variables of the layers being optimised.
            for ctor in _construct_gpar(self, self.vs, self.m, pi + 1).layers:
                ctor()
            names = self.vs.match([f"{pi}/*"] if fix else [f"{i}/*" for i in range(pi + 1)])

            # Analytic gradients: d LML / d theta = sum_ij W_ij dA_ij / d theta with
            # W = 1/2 (alpha alpha^T - A^-1) (dense) or the VFE weights (inducing points), evaluated on the
            # device and pushed through the bound transform here.  Valid when the layers' inputs are fixed.
            inputs_are_data = (self.x_ind is None and not self.replace
                               and not (self.impute and np.isnan(self.y).any()))
            analytic = want_analytic and (fix or inputs_are_data)

            def objective(z):
                # every evaluation builds fresh factors: hand the peer-mapped buffers of the previous
                # one back to the pool (multi-GPU engines; SPMD-safe, every rank evaluates the same z)
                self._release_sharded()
                self.vs.set_latent_vector(names, z)
                gpar = _construct_gpar(self, self.vs, self.m, pi + 1)
                g = ({} if fix else {"every_layer": True}) if analytic else None
                try:
                    if fix:
                        val = -gpar.logpdf(fixed_x, y_cached, None, only_last_layer=True, outputs=[pi],
                                           x_ind=fixed_x_ind, grad_out=g)
                    else:
                        val = -gpar.logpdf(self.x, y_cached, None, only_last_layer=False, grad_out=g)
                except GparError as e:
                    # a non-positive pivot is a legitimate "infeasible point" for the line search; anything
                    # else (ABI misuse, resource limits, unsupported configuration) must surface
                    if "not positive definite" not in str(e):
                        raise
                    return (1e300, np.zeros_like(z)) if analytic else 1e300
                if not np.isfinite(val):
                    return (1e300, np.zeros_like(z)) if analytic else 1e300
                if not analytic:
                    return val
                per_layer = g.get("per_layer", {pi: (g["raw"], g["layer"])} if "raw" in g else {})
                grads = {}  # layers without observations contribute nothing
                for li, (raw, layer) in per_layer.items():
                    raw = raw.cpu().numpy() if hasattr(raw, "cpu") else np.asarray(raw)
                    for name, val_ in named_gradients(layer.terms, raw, noise_name=f"{li}/noise").items():
                        grads[name] = grads.get(name, 0.0) + val_  # tied scales add up over the layers
                gz = -self.vs.latent_gradient(names, grads)
                return val, np.where(np.isfinite(gz), gz, 0.0)

            z0 = self.vs.get_latent_vector(names)
            res = minimize(objective, z0, jac=bool(analytic), method="L-BFGS-B",
                           options={"maxiter": iters, **kw_args})
            self.vs.set_latent_vector(names, res.x)

    def logpdf(self, x, y, w=None, sample_missing=False, posterior=False, normals=None):
        """Log-density of observations (regression.py:461-506).  As in the reference the
        incoming ``y`` goes through ``_unnormalise_y`` (quirk Q1)."""
        self._release_sharded()
        x = _uprank(x)
        y = self._unnormalise_y(self._transform_y(_uprank(y)))
        w = _init_weights(w, y)
        m, p = x.shape[1], y.shape[1]
        if posterior and not self.is_conditioned:
            raise RuntimeError("Must condition or fit model before computing the logpdf under the posterior.")
        gpar = _construct_gpar(self, self.vs, m, p)
        if posterior:
            gpar = gpar | (self.x, self.y, self.w)
        return np.float64(gpar.logpdf(x, y, w, only_last_layer=False, sample_missing=sample_missing,
                                      normals=normals))

    def _release_sharded(self):
        """Multi-GPU engines: free the peer-mapped factor buffers of the previous public call
        (collective; every rank makes the same calls)."""
        if self._engine is not None and self._engine.group is not None:
            self._engine.free_peer_buffers()

    def _sample_device(self, x, w, p, posterior, num_samples, latent, normals, generator=None):
        self._release_sharded()
        x = _uprank(x)
        if posterior and not self.is_conditioned:
            raise RuntimeError("Must condition or fit model before sampling from the posterior.")
        elif not posterior and p is None:
            raise ValueError("Must specify number of outputs to sample.")
        if w is None:
            w = np.ones((x.shape[0], self.p if posterior else p))
        else:
            w = _uprank(w)
        if posterior:
            # The reference re-conditions on every call (regression.py:546-547); here the
            # conditioning is fused into the sampling sweep (one joint factor per layer).
            gpar = _construct_gpar(self, self.vs, self.m, self.p)
            return gpar.sample(x, w, latent=latent, num_samples=num_samples, normals=normals,
                               train=(self.x, self.y, self.w), return_device=True, generator=generator)
        gpar = _construct_gpar(self, self.vs, x.shape[1], p)
        return gpar.sample(x, w, latent=latent, num_samples=num_samples, normals=normals, return_device=True,
                           generator=generator)

    def sample(self, x, w=None, p=None, posterior=False, num_samples=1, latent=False, normals=None):
        """Sample from the prior or posterior (regression.py:508-564).  ``normals``
        injects the standard normals (see :meth:`GPAR.sample`)."""
        dev = self._sample_device(x, w, p, posterior, num_samples, latent, normals)
        host = dev.cpu().numpy()
        samples = [self._untransform_y(self._unnormalise_y(host[s])) for s in range(host.shape[0])]
        return samples[0] if num_samples == 1 else samples

    def predict(self, x, w=None, num_samples=100, latent=False, credible_bounds=False, normals=None):
        """Predictive means (and 95% credible bounds) from posterior samples
        (regression.py:566-597).  With the identity transform the sample mean and the credible
        bounds are reduced on the device and only (n*, p) values come back."""
        kind = _transform_kind(self._untransform_y)
        if kind is not None:
            # identity / log / squishing transforms: every sample is un-normalised and un-transformed on the
            # device (regression.py:553-554), then the mean and the 2.5 / 97.5 percentiles (numpy's linear
            # interpolation) are reduced over the S axis there; only (n*, p) values come back.
            dev = self._sample_device(x, w, None, True, num_samples, latent, normals)
            S, ns, p = dev.shape
            eng = self._engine_of(dev)
            self._untransform_device(eng, dev, kind)
            out = eng.empty(max(ns * p, 1))
            eng.mean_axis0(dev.reshape(-1), S, ns * p, out)
            mean = out.cpu().numpy().reshape(ns, p)
            if not credible_bounds:
                return mean
            lo, hi = eng.percentile2_axis0(dev.reshape(-1), S, ns * p, 2.5, 100 - 2.5)
            return mean, lo.cpu().numpy().reshape(ns, p), hi.cpu().numpy().reshape(ns, p)
        samples = self.sample(x, w, num_samples=num_samples, latent=latent, posterior=True, normals=normals)
        if num_samples == 1:
            samples = [samples]
        mean = np.mean(samples, axis=0)
        if credible_bounds:
            lowers = np.percentile(samples, 2.5, axis=0)
            uppers = np.percentile(samples, 100 - 2.5, axis=0)
            return mean, lowers, uppers
        return mean

    def _untransform_device(self, eng, dev, kind):
        """Un-normalise and un-transform (S, n*, p) device samples in place (gpar_untransform)."""
        S, ns, p = dev.shape
        scale = shift = None
        if self._norm is not None:
            shift, scale = eng.to_device(self._norm[0]), eng.to_device(self._norm[1])
        if scale is not None or kind != 0:
            eng.untransform(dev.reshape(-1), S * ns, p, scale, shift, kind)

    def _engine_of(self, _):
        from .model import default_engine

        return self._engine if self._engine is not None else default_engine()
