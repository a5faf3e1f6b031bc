"""marginal / Grid / quantile / cdf: host-side mirror of reference src/marginal_posterior.jl and the
Grid part of src/interp.jl.

    marginal(jp, f) -> marginal with fields wv, μ, σ, itp       reference src/marginal_posterior.jl:3-8,117-123
    marginal(jp, f, Normal) -> the same with itp::NestedPolyGLM  reference src/marginal_posterior.jl:124-129,
                                                                 src/interp.jl:13-17,365-384
    quantile(m, p), cdf(m, x)                                    reference src/marginal_posterior.jl:140-148,
                                                                 src/interp.jl:458-481
    show(m)                                                      reference src/marginal_posterior.jl:150-155

f may be a coordinate index, a selector such as `lambda Θ: Θ.p[0]` (recognised and evaluated on the
device as a zero-copy view of the Theta array), or any function of the parameter blocks (evaluated on
the host over the downloaded Theta and uploaded as values -- the path an arbitrary Julia closure takes).
"""
import ctypes as C

import numpy as np

from ._lib import GRID_KNOTS, SmoothCDF, check, f64, lib, ptr
from .params import probe_coordinate


class Grid:
    """100-knot piecewise-linear CDF (reference struct Grid, src/interp.jl:9-12; field order weights, values)."""

    def __init__(self, weights, values):
        self.weights = f64(weights)
        self.values = f64(values)


class weights_values:
    """Sorted (weights, values) pair (reference src/interp.jl:5-8 after simultaneous_sort!, :21-26)."""

    def __init__(self, weights, values, cum_weights=None):
        self.weights, self.values, self.cum_weights = weights, values, cum_weights


class marginal_result:
    """reference struct marginal{Ω,T}: wv, μ, σ, itp (src/marginal_posterior.jl:3-8)."""

    def __init__(self, jp, k, mu, sigma, itp):
        self._jp, self._k = jp, k
        self._batch = getattr(jp, "_marginal_batch", 0) if jp is not None else 0
        self.mu, self.sigma, self.itp = float(mu), float(sigma), itp
        self._wv = None

    μ = property(lambda self: self.mu)
    σ = property(lambda self: self.sigma)

    @property
    def wv(self):
        """Sorted weights/values; fetched from the device on first access (only valid until the next
        marginal call on the same posterior)."""
        if self._wv is None:
            if self._jp is None:
                raise RuntimeError("sorted weights/values are not available for this marginal")
            if getattr(self._jp, "_marginal_batch", 0) != self._batch:
                raise RuntimeError("the sorted weights/values of this marginal were replaced by a later marginal / fit call on "
                                   "the same posterior; ask for `wv` before the next call (or compute the marginal again)")
            M = self._jp.n_nodes
            sv, sw, cw = np.zeros(M), np.zeros(M), np.zeros(M)
            check(lib().jp_marginal_sorted(self._jp.handle, C.c_int(self._k), ptr(sv), ptr(sw), ptr(cw)))
            self._wv = weights_values(sw, sv, cw)
        return self._wv

    def quantile(self, p):
        return quantile(self, p)

    def cdf(self, x):
        return cdf(self, x)

    def pdf(self, x):
        return pdf(self, x)

    def __repr__(self):   # reference src/marginal_posterior.jl:150-155
        q = [quantile(self, p) for p in (.025, .25, .5, .75, .975)]
        return "Marginal parameter\nμ: %r\nσ: %r\nQuantiles: [%s]" % (self.mu, self.sigma, " ".join("%.6g" % v for v in q))


class Normal:
    """Tag of the smooth CDF family, as `Normal` in `marginal(jp, f, Normal)` (reference src/marginal_posterior.jl:124-129)."""


class NestedPolyGLM:
    """Smooth CDF Phi(P(Q((x - mu) / sigma))) with nested monotone cubics (reference struct, src/interp.jl:13-17: beta,
    theta, d = Normal(mu, sigma)); fitted by the library (jp_marginal_smooth).  `info` holds the optimiser's report."""

    def __init__(self, c):
        self._c = c
        self.beta, self.theta, self.phi = np.array(c.beta), np.array(c.theta), np.array(c.phi)
        self.β, self.θ = self.beta, self.theta
        self.mu, self.sigma = c.mu, c.sigma
        self.info = dict(objective=c.objective, grad_inf_norm=c.grad_inf_norm, iterations=c.iterations,
                         evaluations=c.evaluations, converged=bool(c.converged),
                         criterion={0: None, 1: "g_tol", 2: "newton_decrement"}.get(int(c.converged)))

    def _call(self, fn, x):
        if np.ndim(x) > 0:
            return np.array([fn(C.byref(self._c), float(v)) for v in np.ravel(x)]).reshape(np.shape(x))
        return fn(C.byref(self._c), float(x))

    def cdf(self, x):          # reference src/interp.jl:365-367
        return self._call(lib().jp_smooth_cdf_eval, x)

    def pdf(self, x):          # reference src/interp.jl:371-374
        return self._call(lib().jp_smooth_pdf_eval, x)

    def quantile(self, p):     # reference src/interp.jl:368-370
        return self._call(lib().jp_smooth_quantile_eval, p)


def quantile(m, p):
    """quantile(::Grid, p) (reference src/interp.jl:467-478) / quantile(::NestedPolyGLM, p) (:368-370); vectorised over p."""
    itp = m.itp if isinstance(m, marginal_result) else m
    if isinstance(itp, NestedPolyGLM):
        return itp.quantile(p)
    if np.ndim(p) > 0:
        return np.array([quantile(itp, float(x)) for x in np.ravel(p)]).reshape(np.shape(p))
    return lib().jp_quantile(ptr(itp.weights), ptr(itp.values), C.c_int(len(itp.weights)), C.c_double(float(p)))


def cdf(m, x):
    """cdf(::Grid, x) (reference src/interp.jl:458-466) / cdf(::NestedPolyGLM, x) (:365-367); vectorised over x."""
    itp = m.itp if isinstance(m, marginal_result) else m
    if isinstance(itp, NestedPolyGLM):
        return itp.cdf(x)
    if np.ndim(x) > 0:
        return np.array([cdf(itp, float(v)) for v in np.ravel(x)]).reshape(np.shape(x))
    return lib().jp_cdf(ptr(itp.weights), ptr(itp.values), C.c_int(len(itp.weights)), C.c_double(float(x)))


def pdf(m, x):
    """pdf(m, x) (reference src/marginal_posterior.jl:143-145): defined for the smooth CDF (src/interp.jl:371-374)."""
    itp = m.itp if isinstance(m, marginal_result) else m
    if not isinstance(itp, NestedPolyGLM):
        raise TypeError("pdf is defined for marginal(jp, f, Normal) only, as in the reference (a Grid has no pdf method)")
    return itp.pdf(x)


def _classify(jp, fs):
    """Split the requested functions into device coordinate selectors and host-evaluated closures."""
    coords, host = [], []
    for i, f in enumerate(fs):
        if isinstance(f, (int, np.integer)):
            coords.append((i, int(f)))
            continue
        k = probe_coordinate(f, jp.M.blocks)
        if k is not None:
            coords.append((i, k))
        else:
            host.append((i, f))
    return coords, host


def marginals(jp, fs):
    """Batched marginal(jp, f) for a list of functions: one library call per kind (device / host)."""
    fs = list(fs)
    out = [None] * len(fs)
    jp._marginal_batch = getattr(jp, "_marginal_batch", 0) + 1      # results of earlier batches lose their device-side sorted arrays
    coords, host = _classify(jp, fs)
    L = lib()
    if coords:
        K = len(coords)
        cs = np.array([k for _, k in coords], dtype=np.int32)
        mu, sg = np.zeros(K), np.zeros(K)
        vn, wn = np.zeros((K, GRID_KNOTS)), np.zeros((K, GRID_KNOTS))
        check(L.jp_marginal_coords(jp.handle, C.c_int(K), ptr(cs), ptr(mu), ptr(sg), ptr(vn), ptr(wn)))
        for j, (i, _) in enumerate(coords):
            out[i] = marginal_result(jp if not host else None, j, mu[j], sg[j], Grid(wn[j].copy(), vn[j].copy()))
    if host:
        K = len(host)
        view = jp.view()
        vals = np.zeros((K, jp.n_nodes))
        for j, (_, f) in enumerate(host):
            v = np.asarray(f(view), dtype=np.float64)
            if v.size == jp.n_nodes:
                v = v.reshape(-1)
            vals[j] = np.broadcast_to(v, (jp.n_nodes,))
        mu, sg = np.zeros(K), np.zeros(K)
        vn, wn = np.zeros((K, GRID_KNOTS)), np.zeros((K, GRID_KNOTS))
        check(L.jp_marginal_values(jp.handle, C.c_int(K), ptr(vals), ptr(mu), ptr(sg), ptr(vn), ptr(wn)))
        for j, (i, _) in enumerate(host):
            out[i] = marginal_result(jp, j, mu[j], sg[j], Grid(wn[j].copy(), vn[j].copy()))
    return out


class MarginalBuffer:
    """What update_MarginalBuffer!(jp, f) leaves in the reference's MarginalBuffer (src/marginal_posterior.jl:10-67):
    ind (stable sort permutation, 0-based node indices), w (cumulative weights in sorted order), V (10 x M, column i =
    powers 0..9 of the standardised value), mu, sigma -- the node-touching input of the smooth-CDF fit."""

    def __init__(self, ind, w, V, mu, sigma):
        self.ind, self.w, self.V, self.mu, self.sigma = ind, w, V, mu, sigma
        self.μ, self.σ = mu, sigma


def marginal_buffer(jp, f):
    """update_MarginalBuffer!(jp, f): evaluated on the GPU (sort, cumulative weights, Vandermonde columns); the download
    is for inspection -- marginal(jp, f, Normal) consumes the same buffers on the device."""
    marginals(jp, [f])                        # moments + value pointers of f on the device
    M = jp.n_nodes
    ind = np.zeros(M, dtype=np.int64)
    w, V = np.zeros(M), np.zeros((M, 10))
    mu, sg = C.c_double(), C.c_double()
    check(lib().jp_marginal_buffer(jp.handle, C.c_int(0), ptr(ind), ptr(w), ptr(V), C.byref(mu), C.byref(sg)))
    return MarginalBuffer(ind, w, V.T, mu.value, sg.value)


def marginal_smooth(jp, f, init=None, max_iter=0, g_tol=0.0):
    """marginal(jp, f, Normal): update_MarginalBuffer! then NestedPolyGLM(m, Normal(mu, sigma)) (reference
    src/marginal_posterior.jl:10-16,124-129; src/interp.jl:377-384), both in the library: the sort, cumulative weights,
    design matrix and every objective / score evaluation of the BFGS iteration run on the GPU over all nodes."""
    c = SmoothCDF()
    x0 = None if init is None else f64(init)
    if x0 is not None and x0.shape != (9,):
        raise ValueError("init: 9 unconstrained parameters expected")
    # M.MarginalBuffers: one entry per marginal function, as `get!(..., jp.M.MarginalBuffers, f)` of the reference
    # (src/marginal_posterior.jl:10,71); the entry is the key under which the library keeps f's sorted design on the device
    table = jp.M.__dict__.setdefault("MarginalBuffers", {})
    fk = ("coord", int(f)) if isinstance(f, (int, np.integer)) else f
    key = table.setdefault(fk, len(table) + 1)
    hit = C.c_int(0)
    args = (ptr(x0), C.c_int(int(max_iter)), C.c_double(float(g_tol)), C.byref(c), C.byref(hit))
    check(lib().jp_marginal_smooth_keyed(jp.handle, C.c_int(-1), C.c_longlong(key), *args))      # a design kept since the last fit?
    if not hit.value:
        marginals(jp, [f])                        # moments + value pointers of f on the device
        check(lib().jp_marginal_smooth_keyed(jp.handle, C.c_int(0), C.c_longlong(key), *args))
    res = marginal_result(jp if not hit.value else None, 0, c.mu, c.sigma, NestedPolyGLM(c))
    res.buffer_reused = bool(hit.value)
    return res


def marginal(jp, f, kind=Grid, **kw):
    """marginal(jp, f[, Grid]) (reference src/marginal_posterior.jl:116-123) and marginal(jp, f, Normal) (:124-129)."""
    if kind is Grid:
        return marginals(jp, [f])[0]
    if kind is Normal:
        return marginal_smooth(jp, f, **kw)
    raise NotImplementedError("Currently unsupported.")      # the reference throws the same for Gamma (:130-133)
