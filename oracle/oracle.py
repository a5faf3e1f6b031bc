"""ctypes/numpy front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  See the header of oracle/jp_oracle.cpp for what is restated from where and how the
oracle is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAMILY = {"binomial_mixture": 0, "logistic": 1, "poisson": 2, "hier_normal": 3, "normal_linear": 4}
RULE = {"GenzKeister": 0, "KronrodPatterson": 1}
REAL, POSITIVE, PROBABILITY = 0, 1, 2

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libjporacle.so")
    src = os.path.join(_HERE, "jp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libjporacle.so"], stdout=subprocess.DEVNULL)
    return so


_FLAVOUR = "parity"   # "parity": -O2 -ffp-contract=off (tests); "fast": -O3 -march=native (bench.py's CPU baseline only)


def build_fast():
    """The CPU-baseline build SURVEY section 8d specifies (-O3 -march=native), compiled ON the machine that times it
    (so `native` is that machine's ISA) into oracle/_fast/ (git-ignored).  The parity tests never use it: contraction
    into FMAs changes rounding."""
    out = os.path.join(_HERE, "_fast")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libjporacle_fast.so")
    src = os.path.join(_HERE, "jp_oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call([os.environ.get("CXX", "g++"), "-O3", "-march=native", "-pthread", "-fPIC", "-std=c++17",
                               "-shared", "-o", so, src])
    return so


def use_fast():
    """Switch this process to the -O3 -march=native build (bench.py only).  Returns the flags string for the report."""
    global _LIB, _FLAVOUR
    try:
        build_fast()
    except (OSError, subprocess.CalledProcessError):
        return "-O2 -march=x86-64-v3 -ffp-contract=off (the -O3 -march=native build failed)"
    _LIB, _FLAVOUR = None, "fast"
    return "-O3 -march=native"


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libjporacle.so")
        if _FLAVOUR == "fast":
            so = os.path.join(_HERE, "_fast", "libjporacle_fast.so")
        elif not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_log_density.restype = C.c_double
        L.orc_log_density_unc.restype = C.c_double
        L.orc_log_density_unc_precise.restype = C.c_double
        L.orc_quantile.restype = C.c_double
        L.orc_cdf.restype = C.c_double
        L.orc_glm_grad_hess.restype = C.c_double
        L.orc_smolyak_build.restype = C.c_longlong
        _LIB = L
    return _LIB


def set_precise(on=True):
    """Accumulate the observation sums of the GLM families in long double (tests); the default plain loop is what bench.py times."""
    lib().orc_set_precise(C.c_int(1 if on else 0))


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------- scale matrix
def _colmajor(A):
    return np.asfortranarray(np.array(A, dtype=np.float64))


def chol(S):
    S = _colmajor(S)
    d = S.shape[0]
    U = np.zeros((d, d), order="F")
    lib().orc_chol(_ptr(U), _ptr(S), C.c_int(d))
    return U


def try_chol(S):
    S = _colmajor(S)
    d = S.shape[0]
    U = np.zeros((d, d), order="F")
    ok = lib().orc_try_chol(_ptr(U), _ptr(S), C.c_int(d))
    return bool(ok), U


def inv_upper(U):
    U = _colmajor(U).copy(order="F")
    lib().orc_inv_upper(_ptr(U), C.c_int(U.shape[0]))
    return U


def inv_chol(H):
    H = _colmajor(H)
    d = H.shape[0]
    U = np.zeros((d, d), order="F")
    lib().orc_inv_chol(_ptr(U), _ptr(H), C.c_int(d))
    return U


def reduce_dimensions(H, max_rank=0):
    H = _colmajor(H)
    d = H.shape[0]
    out = np.zeros((d, d), order="F")
    p = lib().orc_reduce_dimensions(_ptr(H), C.c_int(d), C.c_int(max_rank), _ptr(out))
    return np.asfortranarray(out[:, :p])


def reduce_dimensions_ldr(H, g):
    H = _colmajor(H)
    d = H.shape[0]
    out = np.zeros((d, d), order="F")
    p = lib().orc_reduce_dimensions_ldr(_ptr(H), C.c_int(d), C.c_double(g), _ptr(out))
    return np.asfortranarray(out[:, :p])


def deduce_scale_dynamic(H):
    H = _colmajor(H)
    d = H.shape[0]
    U = np.zeros((d, d), order="F")
    p = lib().orc_deduce_scale_dynamic(_ptr(H), C.c_int(d), _ptr(U))
    return np.asfortranarray(U[:, :p])


# ---------------------------------------------------------------- stage 1
def rule_info(rule=0):
    lv, nm = C.c_int(), C.c_int()
    lib().orc_rule_info(C.c_int(rule), C.byref(lv), C.byref(nm), None, None, None)
    npts = np.zeros(lv.value, dtype=np.int32)
    nodes = np.zeros(nm.value)
    weights = np.zeros((lv.value, nm.value))
    lib().orc_rule_info(C.c_int(rule), C.byref(lv), C.byref(nm), _ptr(npts), _ptr(nodes), _ptr(weights))
    return npts, nodes, weights


def rule_level_nodes(rule, level):
    """Master indices of the nodes of 1-D level `level` (1-based), generation order."""
    n = lib().orc_rule_level_nodes(C.c_int(rule), C.c_int(level), None)
    out = np.zeros(max(n, 0), dtype=np.int32)
    lib().orc_rule_level_nodes(C.c_int(rule), C.c_int(level), _ptr(out))
    return out


def smolyak_sizes(rule, d, L):
    a, b = C.c_longlong(), C.c_longlong()
    lib().orc_smolyak_sizes(C.c_int(rule), C.c_int(d), C.c_int(L), C.byref(a), C.byref(b))
    return a.value, b.value


def smolyak(rule, d, L):
    """-> (idx uint8 [M, d], w float64 [M]) in ascending lexicographic key order."""
    M = lib().orc_smolyak_build(C.c_int(rule), C.c_int(d), C.c_int(L), None, None, C.c_longlong(0))
    idx = np.zeros((M, d), dtype=np.uint8)
    w = np.zeros(M)
    r = lib().orc_smolyak_build(C.c_int(rule), C.c_int(d), C.c_int(L), _ptr(idx), _ptr(w), C.c_longlong(M))
    assert r == M
    return idx, w


# ---------------------------------------------------------------- stages 2-4
def transform(code, x):
    code = np.ascontiguousarray(code, dtype=np.int32)
    x = _d(x)
    th = np.zeros_like(x)
    lj = C.c_double()
    lib().orc_transform(_ptr(code), C.c_int(len(x)), _ptr(x), _ptr(th), C.byref(lj))
    return th, lj.value


def log_density(family, theta, obs, hyper):
    theta, obs, hyper = _d(theta), _d(obs), _d(hyper)
    N = obs.shape[0]
    return lib().orc_log_density(C.c_int(family), _ptr(theta), C.c_int(len(theta)), _ptr(obs), C.c_longlong(N), _ptr(hyper))


def log_density_unc(family, code, x, obs, hyper):
    code = np.ascontiguousarray(code, dtype=np.int32)
    x, obs, hyper = _d(x), _d(obs), _d(hyper)
    return lib().orc_log_density_unc(C.c_int(family), _ptr(code), _ptr(x), C.c_int(len(x)), _ptr(obs),
                                     C.c_longlong(obs.shape[0]), _ptr(hyper), None)


def log_density_unc_precise(family, code, x, obs, hyper):
    """log_density_unc with the observation sum of the GLM families in long double (arbiter at N = 1e7)."""
    code = np.ascontiguousarray(code, dtype=np.int32)
    x, obs, hyper = _d(x), _d(obs), _d(hyper)
    return lib().orc_log_density_unc_precise(C.c_int(family), _ptr(code), _ptr(x), C.c_int(len(x)), _ptr(obs),
                                             C.c_longlong(obs.shape[0]), _ptr(hyper))


def eval_grid(rule, family, code, idx, w, mu_hat, U, neg_min, obs, hyper, m0=0, m1=None, threads=0, want_theta=True):
    """-> dict(theta [d, M], logdens [M], density [M]).  U is d x p (any memory order)."""
    code = np.ascontiguousarray(code, dtype=np.int32)
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    w, mu_hat, obs, hyper = _d(w), _d(mu_hat), _d(obs), _d(hyper)
    U = _colmajor(U)
    d, p = U.shape
    M = idx.shape[0]
    assert idx.shape[1] == p and len(mu_hat) == d
    if m1 is None:
        m1 = M
    theta = np.zeros((d, M)) if want_theta else None
    ld = np.zeros(M)
    dens = np.zeros(M)
    lib().orc_eval_grid(C.c_int(rule), C.c_int(family), _ptr(code), C.c_int(d), C.c_int(p), _ptr(idx), _ptr(w),
                        C.c_longlong(M), C.c_longlong(m0), C.c_longlong(m1), _ptr(mu_hat), _ptr(U),
                        C.c_double(neg_min), _ptr(obs), C.c_longlong(obs.shape[0]), _ptr(hyper),
                        _ptr(theta), _ptr(ld), _ptr(dens), C.c_int(threads))
    return dict(theta=theta, logdens=ld, density=dens)


# ---------------------------------------------------------------- stage 5
def marginal(values, weights, want_sorted=False):
    values, weights = _d(values), _d(weights)
    M = len(values)
    mu, sg = C.c_double(), C.c_double()
    vn, wn = np.zeros(100), np.zeros(100)
    sv = np.zeros(M) if want_sorted else None
    sw = np.zeros(M) if want_sorted else None
    cw = np.zeros(M) if want_sorted else None
    lib().orc_marginal(_ptr(values), _ptr(weights), C.c_longlong(M), C.byref(mu), C.byref(sg), _ptr(vn), _ptr(wn),
                       _ptr(sv), _ptr(sw), _ptr(cw))
    out = dict(mu=mu.value, sigma=sg.value, value_nodes=vn, weight_nodes=wn)
    if want_sorted:
        out.update(sorted_values=sv, sorted_weights=sw, cum_weights=cw)
    return out


def marginal_buffer(values, weights):
    values, weights = _d(values), _d(weights)
    M = len(values)
    ind = np.zeros(M, dtype=np.int64)
    cw, V = np.zeros(M), np.zeros((M, 10))
    mu, sg = C.c_double(), C.c_double()
    lib().orc_marginal_buffer(_ptr(values), _ptr(weights), C.c_longlong(M), _ptr(ind), _ptr(cw), _ptr(V), C.byref(mu),
                              C.byref(sg))
    return dict(ind=ind, cum_weights=cw, V=V, mu=mu.value, sigma=sg.value)


def quantile(weight_nodes, value_nodes, p):
    wn, vn = _d(weight_nodes), _d(value_nodes)
    return lib().orc_quantile(_ptr(wn), _ptr(vn), C.c_int(len(wn)), C.c_double(p))


def cdf(weight_nodes, value_nodes, x):
    wn, vn = _d(weight_nodes), _d(value_nodes)
    return lib().orc_cdf(_ptr(wn), _ptr(vn), C.c_int(len(wn)), C.c_double(x))


# ---------------------------------------------------------------- smooth CDF (marginal(jp, f, Normal))
SMOOTH_INIT = np.zeros(9)      # MarginalBuffer.init lives in the absent LogDensities; zeros: a = c = m = 1, b = l = n = d = 0


def smooth_objective(V, cum_weights, phi, want_grad=True):
    """ntl_likelihood! / ntscore!, reference src/interp.jl:81-111.  Returns f, grad[9], beta[10], theta[7]."""
    V, cw, phi = _d(V), _d(cum_weights), _d(phi)
    f = C.c_double()
    g = np.zeros(9) if want_grad else None
    beta, theta = np.zeros(10), np.zeros(7)
    lib().orc_smooth_objective(_ptr(V), _ptr(cw), C.c_longlong(len(cw)), _ptr(phi), C.byref(f), _ptr(g), _ptr(beta),
                               _ptr(theta))
    return f.value, g, beta, theta


def smooth_fit(V, cum_weights, mu, sigma, init=None, gtol=1e-8, maxiter=4000):
    """NestedPolyGLM(m, Normal(mu, sigma)), reference src/interp.jl:377-384: BFGS on (f, score) from `init`.  The reference
    uses Optim's BFGS with a backtracking line search; here scipy's (test side only)."""
    from scipy.optimize import minimize

    def fg(phi):
        f, g, _, _ = smooth_objective(V, cum_weights, phi)
        if not np.isfinite(f):
            return 1e300, np.zeros(9)
        return f, g

    r = minimize(fg, SMOOTH_INIT.copy() if init is None else _d(init), jac=True, method="BFGS",
                 options=dict(gtol=gtol, maxiter=maxiter))
    f, g, beta, theta = smooth_objective(V, cum_weights, r.x)
    return dict(phi=r.x, f=f, grad=g, beta=beta, theta=theta, mu=mu, sigma=sigma, iterations=r.nit)


def smooth_cdf(fit, x):
    """reference src/interp.jl:365-367 with polyexpreval (:338-346)."""
    from scipy.special import erf
    z = (x - fit["mu"]) / fit["sigma"]
    b = fit["beta"]
    zi, out = z, b[0] + z * b[1]
    for i in range(2, len(b)):
        zi = zi * z
        out = out + zi * b[i]
    return (1 + erf(out / np.sqrt(2))) / 2


def smooth_pdf(fit, x):
    """reference src/interp.jl:371-374 with d-polyexpreval (:348-361)."""
    z = (x - fit["mu"]) / fit["sigma"]
    b = fit["beta"]
    fx, dfx, zi = b[0], 0.0, 1.0
    for i in range(1, len(b)):
        dfx = dfx + zi * i * b[i]
        zi = zi * z
        fx = fx + zi * b[i]
    return np.exp(-fx * fx / 2) * (2 * np.pi) ** -0.5 * dfx / fit["sigma"]


def _one_cubic_root(a, c, b, d):
    """reference src/interp.jl:388-394 (argument order a, c, b, d as there)."""
    D0 = b * b - 3 * a * c
    D1 = 2 * b ** 3 - 9 * a * b * c + 27 * a * a * d
    Cc = np.cbrt((D1 + np.sqrt(D1 * D1 - 4 * D0 ** 3)) / 2)
    return -(b + Cc + D0 / Cc) / (3 * a)


def smooth_quantile(fit, p):
    """reference src/interp.jl:368-370 and nested_root (:402-405)."""
    from scipy.special import erfinv
    t = fit["theta"]
    r1 = _one_cubic_root(t[0], t[1], t[2], t[3] - np.sqrt(2) * erfinv(2 * p - 1))
    return _one_cubic_root(1.0, t[4], t[5], t[6] - r1) * fit["sigma"] + fit["mu"]


# ---------------------------------------------------------------- GLM mode (test side)
def glm_grad_hess(family, beta, obs, hyper):
    beta, obs, hyper = _d(beta), _d(obs), _d(hyper)
    d = len(beta)
    g = np.zeros(d)
    H = np.zeros((d, d), order="F")
    ll = lib().orc_glm_grad_hess(C.c_int(family), _ptr(beta), C.c_int(d), _ptr(obs), C.c_longlong(obs.shape[0]),
                                 _ptr(hyper), _ptr(g), _ptr(H))
    return ll, g, H


def glm_mode(family, obs, hyper, d, iters=50, tol=1e-13):
    """Newton iteration for the posterior mode; returns (beta_hat, Hneg, log posterior at the mode)."""
    beta = np.zeros(d)
    for _ in range(iters):
        ll, g, H = glm_grad_hess(family, beta, obs, hyper)
        step = np.linalg.solve(H, g)
        # damp while the objective does not increase
        t = 1.0
        while t > 1e-8:
            ll2, _, _ = glm_grad_hess(family, beta + t * step, obs, hyper)
            if ll2 >= ll - 1e-12 * abs(ll):
                break
            t *= 0.5
        beta = beta + t * step
        if np.max(np.abs(t * step)) < tol * (1 + np.max(np.abs(beta))):
            break
    ll, g, H = glm_grad_hess(family, beta, obs, hyper)
    return beta, H, ll


def num_threads():
    return lib().orc_num_threads()
