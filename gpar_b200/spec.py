"""Host-side kernel recipe of GPAR (gpar/regression.py:31-59, 72-182) and its
lowering to the feature-map kernel spec consumed by the CUDA Gram kernels."""
import fnmatch
import math

import numpy as np

from . import _lib

__all__ = ["determine_indices", "vector_from_init", "Vars", "model_terms", "lower_terms", "term_gradients",
           "named_gradients", "LayerModel"]


def determine_indices(m, pi, markov):
    """Column indices feeding layer ``pi`` (gpar/regression.py:49-59): all ``m``
    inputs plus the previous outputs, truncated to the last ``markov`` ones."""
    p_last = pi - 1
    p_start = 0 if markov is None else max(p_last - (markov - 1), 0)
    p_num = p_last - p_start + 1
    m_inds = list(range(m))
    p_inds = list(range(m + p_start, m + p_last + 1))
    return m_inds, p_inds, p_num


def vector_from_init(init, length):
    """Broadcast / truncate a hyper-parameter initialiser (gpar/regression.py:31-46)."""
    if np.size(init) == 1:
        return init * np.ones(length)
    init_squeezed = np.squeeze(init)
    if np.ndim(init_squeezed) != 1:
        raise ValueError("Incorrect shape {} of hyperparameters.".format(np.shape(init)))
    if np.size(init_squeezed) < length:
        raise ValueError("Not enough hyperparameters specified.")
    return np.array(init_squeezed)[:length]


class Vars:
    """Named hyper-parameter store standing in for ``varz.Vars`` (regression.py:314):
    ``bnd``/``get`` return the current value, creating it from ``init`` on first
    use.  Bounded variables are optimised in the unconstrained space
    ``z = logit((v - lower) / (upper - lower))`` by :meth:`GPARRegressor.fit`."""

    def __init__(self):
        self._values = {}
        self._bounds = {}

    @property
    def names(self):
        return list(self._values.keys())

    def __contains__(self, name):
        return name in self._values

    def __getitem__(self, name):
        return self._values[name]

    def bnd(self, name, init, lower=1e-4, upper=1e4):
        if name not in self._values:
            self._values[name] = np.clip(np.array(init, dtype=np.float64), lower, upper)
            self._bounds[name] = (lower, upper)
        return self._values[name]

    def get(self, name, init):
        if name not in self._values:
            self._values[name] = np.array(init, dtype=np.float64)
            self._bounds[name] = None
        return self._values[name]

    def assign(self, name, value):
        self._values[name] = np.array(value, dtype=np.float64).reshape(np.shape(self._values[name]))

    def match(self, patterns):
        return [n for n in self._values if any(fnmatch.fnmatch(n, p) for p in patterns)]

    def copy(self):
        vs = Vars()
        vs._values = {k: np.array(v) for k, v in self._values.items()}
        vs._bounds = dict(self._bounds)
        return vs

    # -- packing for the optimiser ------------------------------------
    def get_latent_vector(self, names):
        out = []
        for n in names:
            v = np.atleast_1d(self._values[n]).astype(np.float64)
            b = self._bounds[n]
            if b is None:
                out.append(v)
            else:
                lo, hi = b
                t = np.clip((v - lo) / (hi - lo), 1e-15, 1 - 1e-15)
                out.append(np.log(t) - np.log1p(-t))
        return np.concatenate(out) if out else np.zeros(0)

    def latent_gradient(self, names, grads):
        """Chain rule through the bound transform: ``grads`` maps variable name -> d f / d value
        (missing names count as zero); returns d f / d z in the packing order of
        :meth:`get_latent_vector`."""
        out = []
        for n in names:
            v = np.atleast_1d(self._values[n]).astype(np.float64)
            g = np.atleast_1d(np.asarray(grads.get(n, np.zeros_like(v)), dtype=np.float64)).reshape(v.shape)
            b = self._bounds[n]
            if b is not None:
                lo, hi = b
                g = g * (v - lo) * (hi - v) / (hi - lo)
            out.append(g.reshape(-1))
        return np.concatenate(out) if out else np.zeros(0)

    def set_latent_vector(self, names, z):
        i = 0
        for n in names:
            size = int(np.size(self._values[n]))
            zi = np.asarray(z[i : i + size], dtype=np.float64)
            i += size
            b = self._bounds[n]
            if b is None:
                v = zi
            else:
                lo, hi = b
                v = lo + (hi - lo) / (1.0 + np.exp(-zi))
            self._values[n] = v.reshape(np.shape(self._values[n]))


def model_terms(vs, m, pi, scale, scale_tie, per, per_period, per_scale, per_decay, input_linear,
                input_linear_scale, linear, linear_scale, nonlinear, nonlinear_scale, rq, markov, noise):
    """The composite kernel of layer ``pi`` as a list of terms plus its noise
    variance -- the recipe of ``_model_generator`` (gpar/regression.py:92-180),
    including the variable names, initial values and bounds it registers."""
    m_inds, p_inds, p_num = determine_indices(m, pi, markov)
    terms = []
    variance = vs.bnd(name=f"{pi}/input/var", init=1.0)
    scales = vs.bnd(name=f"{0 if scale_tie else pi}/input/scales", init=vector_from_init(scale, m))
    if rq:
        alpha = vs.bnd(name=f"{pi}/input/alpha", init=1e-2, lower=1e-3, upper=1e3)
        terms.append(dict(type="rq", variance=variance, cols=m_inds, scales=scales, alpha=alpha,
                          names=dict(variance=f"{pi}/input/var", scales=f"{0 if scale_tie else pi}/input/scales",
                                     alpha=f"{pi}/input/alpha")))
    else:
        terms.append(dict(type="eq", variance=variance, cols=m_inds, scales=scales,
                          names=dict(variance=f"{pi}/input/var", scales=f"{0 if scale_tie else pi}/input/scales")))
    if per:
        variance = vs.bnd(name=f"{pi}/input/per/var", init=1.0)
        scales = vs.bnd(name=f"{pi}/input/per/scales", init=vector_from_init(per_scale, 2 * m))
        periods = vs.bnd(name=f"{pi}/input/per/pers", init=vector_from_init(per_period, m))
        decays = vs.bnd(name=f"{pi}/input/per/decay", init=vector_from_init(per_decay, m))
        terms.append(dict(type="periodic", variance=variance, cols=m_inds, scales=scales, periods=periods,
                          decays=decays,
                          names=dict(variance=f"{pi}/input/per/var", scales=f"{pi}/input/per/scales",
                                     periods=f"{pi}/input/per/pers", decays=f"{pi}/input/per/decay")))
    if input_linear:
        scales = vs.bnd(name=f"{pi}/input/lin/scales", init=vector_from_init(input_linear_scale, m))
        const = vs.get(name=f"{pi}/input/lin/const", init=1.0)
        terms.append(dict(type="linear", variance=1.0, cols=m_inds, scales=scales,
                          names=dict(scales=f"{pi}/input/lin/scales")))
        terms.append(dict(type="const", variance=const, names=dict(variance=f"{pi}/input/lin/const")))
    if linear and pi > 0:
        scales = vs.bnd(name=f"{pi}/output/lin/scales", init=vector_from_init(linear_scale, p_num))
        terms.append(dict(type="linear", variance=1.0, cols=p_inds, scales=scales,
                          names=dict(scales=f"{pi}/output/lin/scales")))
    if nonlinear and pi > 0:
        variance = vs.bnd(name=f"{pi}/output/nonlin/var", init=1.0)
        scales = vs.bnd(name=f"{pi}/output/nonlin/scales", init=vector_from_init(nonlinear_scale, p_num))
        if rq:
            alpha = vs.bnd(name=f"{pi}/output/nonlin/alpha", init=1e-2, lower=1e-3, upper=1e3)
            terms.append(dict(type="rq", variance=variance, cols=p_inds, scales=scales, alpha=alpha,
                              names=dict(variance=f"{pi}/output/nonlin/var", scales=f"{pi}/output/nonlin/scales",
                                         alpha=f"{pi}/output/nonlin/alpha")))
        else:
            terms.append(dict(type="eq", variance=variance, cols=p_inds, scales=scales,
                              names=dict(variance=f"{pi}/output/nonlin/var", scales=f"{pi}/output/nonlin/scales")))
    noise_variance = vs.bnd(name=f"{pi}/noise", init=vector_from_init(noise, pi + 1)[pi], lower=1e-8)
    return terms, float(noise_variance)


_TYPE = {"eq": _lib.TERM_EQ, "rq": _lib.TERM_RQ, "linear": _lib.TERM_LINEAR, "const": _lib.TERM_CONST}


def lower_terms(terms):
    """Lower a term list to the C-ABI ``gpar_kernel_spec_t`` (feature-map normal
    form, include/gpar_b200.h).  ``periodic`` becomes one EQ term over the 3m
    features [sin(2 pi x / T) / s_c, cos(2 pi x / T) / s_{m+c}, x / d_c]; terms
    with no columns (e.g. ``markov=0`` output kernels) are dropped (ZeroKernel)."""
    spec = _lib.KernelSpec()
    nf = nt = 0

    def feat(col, op, a, b=0.0):
        nonlocal nf
        if nf >= _lib.MAX_FEATS:
            raise ValueError("kernel needs more than GPAR_MAX_FEATS features")
        spec.feat_col[nf], spec.feat_op[nf], spec.feat_a[nf], spec.feat_b[nf] = int(col), op, float(a), float(b)
        nf += 1

    for t in terms:
        kind = t["type"]
        cols = list(t.get("cols", []))
        if kind != "const" and len(cols) == 0:
            continue
        if nt >= _lib.MAX_TERMS:
            raise ValueError("kernel needs more than GPAR_MAX_TERMS terms")
        T = spec.terms[nt]
        T.f_begin = nf
        T.variance = float(t.get("variance", 1.0))
        T.alpha = float(t.get("alpha", 1.0))
        if kind == "periodic":
            T.type = _lib.TERM_EQ
            m = len(cols)
            scales = np.asarray(t["scales"], dtype=np.float64).reshape(-1)
            periods = np.asarray(t["periods"], dtype=np.float64).reshape(-1)
            decays = np.asarray(t["decays"], dtype=np.float64).reshape(-1)
            for j, c in enumerate(cols):
                feat(c, _lib.FEAT_SIN, 1.0 / scales[j], 2.0 * math.pi / periods[j])
            for j, c in enumerate(cols):
                feat(c, _lib.FEAT_COS, 1.0 / scales[m + j], 2.0 * math.pi / periods[j])
            for j, c in enumerate(cols):
                feat(c, _lib.FEAT_SCALE, 1.0 / decays[j])
        else:
            T.type = _TYPE[kind]
            if kind != "const":
                scales = np.asarray(t["scales"], dtype=np.float64).reshape(-1)
                for j, c in enumerate(cols):
                    feat(c, _lib.FEAT_SCALE, 1.0 / scales[j])
        T.f_end = nf
        nt += 1
    spec.n_terms, spec.n_feats = nt, nf
    return spec


def term_gradients(terms, raw):
    """Finish the chain rule of ``gpar_gram_grad``: ``raw`` is its output vector (layout in
    include/gpar_b200.h), ``terms`` the term list the spec was lowered from.  Returns one dict per term
    with the derivative of the log-marginal w.r.t. each of its fields (``variance``, ``scales``,
    ``alpha``, ``periods``, ``decays``), walking the terms exactly as :func:`lower_terms` does."""
    raw = np.asarray(raw, dtype=np.float64)
    base = 2 * _lib.MAX_TERMS
    out = []
    nf = nt = 0
    for t in terms:
        kind = t["type"]
        cols = list(t.get("cols", []))
        g = {}
        if kind != "const" and len(cols) == 0:
            out.append(g)
            continue
        g["variance"] = raw[2 * nt]
        if kind == "rq":
            g["alpha"] = raw[2 * nt + 1]

        def d_da(f, a):  # d LML / d a_f
            s1 = raw[base + 2 * f]
            return 2.0 * s1 / a if kind == "linear" else -s1 / a

        if kind == "periodic":
            m = len(cols)
            scales = np.asarray(t["scales"], dtype=np.float64).reshape(-1)
            periods = np.asarray(t["periods"], dtype=np.float64).reshape(-1)
            decays = np.asarray(t["decays"], dtype=np.float64).reshape(-1)
            gs, gp, gd = np.zeros(2 * m), np.zeros(m), np.zeros(m)
            for j in range(m):  # sin block, cos block, decay block
                for blk in range(2):
                    f = nf + blk * m + j
                    a = 1.0 / scales[blk * m + j]
                    gs[blk * m + j] = d_da(f, a) * (-a * a)           # a = 1 / s
                    gp[j] += -raw[base + 2 * f + 1] * (-2.0 * math.pi / periods[j] ** 2)  # b = 2 pi / T
                f = nf + 2 * m + j
                a = 1.0 / decays[j]
                gd[j] = d_da(f, a) * (-a * a)
            g["scales"], g["periods"], g["decays"] = gs, gp, gd
            nf += 3 * m
        elif kind != "const":
            scales = np.asarray(t["scales"], dtype=np.float64).reshape(-1)
            gs = np.zeros(len(cols))
            for j in range(len(cols)):
                a = 1.0 / scales[j]
                gs[j] = d_da(nf + j, a) * (-a * a)
            g["scales"] = gs
            nf += len(cols)
        out.append(g)
        nt += 1
    return out


def named_gradients(terms, raw, noise_name=None):
    """d LML / d (named hyper-parameter) from the raw sums: fields of several terms that share a
    variable (``scale_tie``) add up; the noise variance takes the diagonal term (dvec = 1 / w)."""
    grads = {}
    for t, g in zip(terms, term_gradients(terms, raw)):
        for field, name in t.get("names", {}).items():
            if field in g:
                grads[name] = grads.get(name, 0.0) + np.asarray(g[field], dtype=np.float64)
    if noise_name is not None:
        grads[noise_name] = np.asarray(raw, dtype=np.float64)[_lib.GRAD_NP - 1]
    return grads


class LayerModel:
    """What a GPAR layer constructor returns here: the kernel (term list, lowered
    lazily to the device spec) and the noise variance -- the counterpart of the
    reference's ``(f, noise)`` tuple (gpar/model.py:47-57, regression.py:176-180)."""

    def __init__(self, terms, noise, block=None):
        self.terms = terms
        self.noise = float(noise)
        self.block = block  # observation block of a posterior layer (``f | obs``), else None
        self._spec = None

    def conditioned(self, block):
        """``f | obs``: the same kernel carrying the observations it is conditioned on."""
        post = LayerModel(self.terms, self.noise, block=block)
        post._spec = self._spec
        return post

    @property
    def spec(self):
        if self._spec is None:
            self._spec = lower_terms(self.terms)
        return self._spec

    def __iter__(self):  # allows ``f, noise = model()``
        return iter((self, self.noise))
