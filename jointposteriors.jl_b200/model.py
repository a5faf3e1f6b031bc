"""Model / mode / fit: the host-side mirror of reference src/joint_posterior.jl.

    Model(ParamStruct | tuple-of-blocks [, build])          reference README.md:39, test/runtests.jl:45
    fit(model, data [, n])  -> JointPosterior                reference src/joint_posterior.jl:177-188
    mode(model, data)       -> (mu_hat, U, neg_min)          reference src/joint_posterior.jl:164-168

All numerical work on the path (grid construction, node -> theta map, log-densities, weight
normalisation) happens in libjpcuda.so on the GPU; there is no CPU fallback.  `mode` itself is
upstream of the accelerated path: GLM families use a Newton iteration on the GPU
score/information sums, other families a Newton iteration on finite differences of the
GPU-evaluated log-density.
"""
import ctypes as C

import numpy as np

from . import _lib, linalg
from ._lib import FitArgs, JPError, check, colmajor, f64, lib, ptr
from .data import FAM_LOGISTIC, FAM_POISSON, Data
from .params import ParamView, blocks_of, transform_codes


# ----------------------------------------------------------------------------- rule / build / rank types
class GenzKeister:
    rule_id = 0


class KronrodPatterson:
    rule_id = 1


class _Build:
    raw = False

    def __init__(self, rule=GenzKeister):
        self.rule = rule

    def __class_getitem__(cls, rule):   # Smolyak[GenzKeister] mirrors the reference's b{q}
        return cls(rule)


class Smolyak(_Build):
    """CacheBuild: fit returns Theta per node (reference src/joint_posterior.jl:177-182)."""
    raw = False


class SmolyakRaw(_Build):
    """RawBuild: fit keeps the unconstrained node cache (reference src/joint_posterior.jl:183-188)."""
    raw = True


class Dynamic:
    """Cholesky if positive definite, else eigen fallback (reference src/joint_posterior.jl:136-138)."""


class Full:
    """Always Cholesky (reference src/joint_posterior.jl:139-141)."""


class FixedRank:
    """Keep at most p eigen directions (reference src/joint_posterior.jl:120-134)."""

    def __init__(self, p):
        self.p = int(p)


class LDR:
    """Keep the leading eigen directions that carry the fraction g of the posterior variance
    (reference src/joint_posterior.jl:78-95,111-119)."""

    def __init__(self, g):
        self.g = float(g)
        if not 0.0 < self.g < 1.0:
            raise ValueError("LDR: need 0 < g < 1")       # @assert, reference :79


def default(build=None):
    """Default Smolyak level `default(B)` (reference src/joint_posterior.jl:177); the value lives in the
    absent SparseQuadratureGrids package -- 5 here, the level SURVEY/BASELINE quote configs on."""
    return 5


# ----------------------------------------------------------------------------- context
class Context:
    """One GPU + one stream (jp_ctx).  Created lazily; shared by default."""
    _default = {}

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib().jp_ctx_create(C.c_int(device), C.byref(h)))
        self.handle = h
        self.device = device
        self._data_cache = {}

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = 0
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def use_stream(self, cuda_stream_ptr):
        check(lib().jp_ctx_set_stream(self.handle, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        check(lib().jp_ctx_sync(self.handle))

    def launches(self):
        return int(lib().jp_ctx_launch_count(self.handle))

    def last_kernel_ms(self):
        """Device time of the last node x observation log-density kernel launch (CUDA events, blocking)."""
        ms = C.c_float()
        check(lib().jp_ctx_last_kernel_ms(self.handle, C.byref(ms)))
        return float(ms.value)

    def trace(self, on=True):
        """Start (and clear) / stop the stage trace of this context (jp_ctx_trace)."""
        check(lib().jp_ctx_trace(self.handle, C.c_int(1 if on else 0)))

    def trace_dump(self):
        """[(name, microseconds since the first mark)] of the events recorded since trace(True)."""
        buf = C.create_string_buffer(1 << 16)
        check(lib().jp_ctx_trace_dump(self.handle, buf, C.c_int(len(buf))))
        out = []
        for line in buf.value.decode().splitlines():
            name, us = line.split("\t")
            out.append((name, float(us)))
        return out

    def grid(self, rule_id, d_eff, level):
        g = C.c_void_p()
        check(lib().jp_grid_get(self.handle, C.c_int(rule_id), C.c_int(d_eff), C.c_int(level), C.byref(g)))
        return g

    def upload(self, data):
        return DeviceData(self, data)

    def upload_sharded(self, data, group=None):
        """Every rank of the box ends up with ALL observation records, but each uploads only its 1/world row slice
        over PCIe; the slices are exchanged GPU->GPU by one all_gather (NCCL over NVLink).  See distributed.gather_rows."""
        from . import distributed as D
        import torch
        obs, hyper = data.records()
        dev = torch.device("cuda", self.device)
        # the gather runs on torch's current stream; the library's kernels must be ordered behind it
        self.use_stream(torch.cuda.current_stream(dev).cuda_stream)
        full = D.gather_rows(obs, dev, group)
        return DeviceData(self, data, device_obs=full)


class DeviceData:
    """Observations resident on the GPU (jp_data)."""

    def __init__(self, ctx, data, device_obs=None, rows=None):
        obs, hyper = data.records()
        if rows is not None:             # this rank's observation slice only (observation-sharded fits)
            obs = obs[int(rows[0]):int(rows[1])]
        obs = f64(obs)
        hyper = f64(hyper)
        h = C.c_void_p()
        if device_obs is None:
            check(lib().jp_data_upload(ctx.handle, C.c_int(data.family), C.c_longlong(obs.shape[0]), C.c_int(obs.shape[1]),
                                       ptr(obs), ptr(hyper), C.c_int(len(hyper)), C.byref(h)))
        else:
            # a torch CUDA tensor holding (at least) the N x ncols records, borrowed by the library: keep it alive
            self._device_obs = device_obs
            check(lib().jp_data_adopt_device(ctx.handle, C.c_int(data.family), C.c_longlong(obs.shape[0]),
                                             C.c_int(obs.shape[1]), C.c_void_p(device_obs.data_ptr()), ptr(hyper),
                                             C.c_int(len(hyper)), C.byref(h)))
        self.ctx, self.handle, self.family = ctx, h, data.family
        self.N, self.ncols = obs.shape
        self.nbytes = obs.nbytes

    def free(self):
        if self.handle:
            lib().jp_data_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ----------------------------------------------------------------------------- model
class Model:
    """Model(ParamStruct) / Model(tuple, build): parameter layout + grid build + rank policy.

    The reference's Model also owns mutable buffers (diff_buffer, Grid.U, MarginalBuffers); here the
    device-resident buffers belong to the JointPosterior returned by fit.
    """

    def __init__(self, params, build=None, rank=Dynamic, device=None, hessian_scale=2.0):
        self.blocks = blocks_of(params)
        self.params = params
        self.d = int(sum(b.n for _, b in self.blocks))
        self.transform = transform_codes(self.blocks)
        if build is None:
            build = Smolyak(GenzKeister)
        if isinstance(build, type):
            build = build()
        self.build = build
        self.rank = rank
        self.device = device
        # reference src/joint_posterior.jl:167 passes 2 * Hessian to deduce_scale!
        self.hessian_scale = float(hessian_scale)

    @property
    def ctx(self):
        return Context.get(self.device)


def _as_device_data(M, data):
    if isinstance(data, DeviceData):
        return data
    if not isinstance(data, Data):
        raise TypeError("data must be a jointposteriors Data object (one per registered likelihood family)")
    return M.ctx.upload(data)


def log_density_unc(M, ddata, X):
    """Unconstrained log-density (with log-Jacobian) at the rows of X, evaluated on the GPU."""
    X = f64(np.atleast_2d(X))
    K, d = X.shape
    out = np.zeros(K)
    code = np.ascontiguousarray(M.transform, dtype=np.int32)
    check(lib().jp_log_density_points(ddata.ctx.handle, ddata.handle, C.c_int(d), ptr(code), C.c_longlong(K),
                                      ptr(X), ptr(out)))
    return out


def _native_mode(M, ddata, x0, glm):
    """jp_mode: the Newton iterations run in native host code inside libjpcuda (csrc/jp_hostlinalg.cpp), one batched
    GPU evaluation per iteration -- analytic score / information for GLMs, saddle-free Newton on Richardson finite
    differences of the plugin log-density otherwise."""
    d = M.d
    x = np.array(x0, dtype=np.float64)
    H = np.zeros((d, d), order="F")
    neg_min = C.c_double()
    evals = C.c_int()
    code = np.ascontiguousarray(M.transform, dtype=np.int32)
    check(lib().jp_mode(ddata.ctx.handle, ddata.handle, C.c_int(d), ptr(code), C.c_int(1 if glm else 0), ptr(x), ptr(H),
                        C.byref(neg_min), C.byref(evals)))
    _warn_unless_converged("mode")
    return x, np.array(H), neg_min.value


def _warn_unless_converged(who):
    """jp_mode_report: a search that stopped at its iteration cap or after a failed line search still returns its last point;
    say so instead of silently centring the grid there."""
    import warnings
    g, it, ok = C.c_double(), C.c_int(), C.c_int()
    lib().jp_mode_report(C.byref(g), C.byref(it), C.byref(ok))
    if not ok.value:
        warnings.warn("%s: the mode search did not converge (%d iterations, gradient infinity norm %.3g); the grid is centred on "
                      "its last point" % (who, it.value, g.value), RuntimeWarning, stacklevel=3)


def deduce_scale(M, H):
    """deduce_scale!(M, H, R) (reference src/joint_posterior.jl:136-144)."""
    R = M.rank
    if R is Full or isinstance(R, Full):
        return linalg.inv_chol(H)
    if isinstance(R, FixedRank):
        return linalg.reduce_dimensions(H, R.p)
    if isinstance(R, LDR):
        return linalg.reduce_dimensions_ldr(H, R.g)
    return linalg.deduce_scale_dynamic(H)


def mode(M, data, x0=None):
    """(mu_hat, U, neg_min): unconstrained posterior mode, scale matrix from hessian_scale x Hessian of
    the negative log-density, and the minimised objective (reference src/joint_posterior.jl:164-168)."""
    ddata = _as_device_data(M, data)
    if x0 is None:
        x0 = np.zeros(M.d)
    glm = ddata.family in (FAM_LOGISTIC, FAM_POISSON) and np.all(M.transform == 0)
    x, H, neg_min = _native_mode(M, ddata, x0, glm)
    U = deduce_scale(M, M.hessian_scale * H)
    return x, U, neg_min


# ----------------------------------------------------------------------------- joint posterior
class JointPosterior:
    """Result of fit (reference struct JointPosterior, src/joint_posterior.jl:2-8).

    Fields: M (model), Θ / Theta (constrained parameters, [d, nodes]; device-resident, downloaded on
    access), density (normalised, signed weights), μ_hat / mu_hat, U.
    """

    def __init__(self, M, ddata, grid, mu_hat, U, neg_min, path=_lib.PATH_AUTO, node_range=None, raw=False):
        self.M = M
        self.data = ddata
        self.grid = grid
        self.mu_hat = f64(mu_hat)
        self.U = colmajor(U)
        self.neg_min = float(neg_min)
        self.ctx = ddata.ctx
        d, p = self.U.shape
        self._code = np.ascontiguousarray(M.transform, dtype=np.int32)
        self._args = FitArgs()
        a = self._args
        a.d, a.p = d, p
        a.h_transform = self._code.ctypes.data_as(C.POINTER(C.c_int))
        a.h_mu_hat = self.mu_hat.ctypes.data_as(C.POINTER(C.c_double))
        a.h_U = self.U.ctypes.data_as(C.POINTER(C.c_double))
        a.neg_min = self.neg_min
        a.path = int(path)
        a.node_begin, a.node_end = (0, -1) if node_range is None else node_range
        a.raw = 1 if raw else 0
        h = C.c_void_p()
        check(lib().jp_posterior_create(self.ctx.handle, grid, ddata.handle, C.byref(a), C.byref(h)))
        self.handle = h
        self.n_nodes = int(lib().jp_posterior_size(h))
        self._theta = self._density = None
        self._theta_gen = 0

    def update(self, mu_hat, U, neg_min):
        """Re-point the posterior at a new (mu_hat, U, neg_min) of the same shape (buffers are reused)."""
        self.mu_hat[:] = mu_hat
        self.U[:] = U
        self.neg_min = float(neg_min)
        self._args.neg_min = self.neg_min
        self._theta = self._density = None

    def evaluate(self):
        """Stages 2-4 on the GPU (asynchronous)."""
        check(lib().jp_fit(self.handle, C.byref(self._args)))
        self._theta = self._density = None
        self._theta_gen += 1
        self._marginal_batch = getattr(self, "_marginal_batch", 0) + 1
        return self

    @property
    def path_used(self):
        return int(lib().jp_fit_path_used(self.handle))

    @property
    def diagnostics(self):
        """A-priori error figures of the tensor-core path (jp_fit_diagnostics)."""
        out = np.zeros(8)
        check(lib().jp_fit_diagnostics(self.handle, ptr(out)))
        return dict(max_delta_eta=out[0], truncation_bound=out[1], rounding_estimate=out[2], series_terms=int(out[3]),
                    rounding_worst_case=out[4], economised=bool(out[5]))

    @property
    def Theta(self):
        if self._theta is None:
            out = np.zeros((self._args.d, self.n_nodes))
            check(lib().jp_get_theta(self.handle, ptr(out)))
            self._theta = out
        return self._theta

    Θ = Theta

    @property
    def density(self):
        if self._density is None:
            out = np.zeros(self.n_nodes)
            check(lib().jp_get_density(self.handle, ptr(out)))
            self._density = out
        return self._density

    @property
    def logdens(self):
        out = np.zeros(self.n_nodes)
        check(lib().jp_get_logdens(self.handle, ptr(out)))
        return out

    μ_hat = property(lambda self: self.mu_hat)

    def view(self):
        """Theta as named blocks (what marginal functions receive)."""
        return ParamView(self.M.blocks, self.Theta)

    def free(self):
        if getattr(self, "handle", None):
            lib().jp_posterior_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _RawGrid:
    """The `grid` field of the reference's JointPosteriorRaw (src/joint_posterior.jl:9-14): `cache` = the d x M matrix of
    UNCONSTRAINED node coordinates mu_hat + U z (src/marginal_posterior.jl:69,107), `density` = the normalised weights."""

    def __init__(self, jp):
        self._jp = jp
        self._cache = None

    @property
    def cache(self):
        if self._cache is None or self._jp._theta_gen != self._gen:
            out = np.zeros((self._jp._args.d, self._jp.n_nodes))
            check(lib().jp_get_cache(self._jp.handle, ptr(out)))
            self._cache, self._gen = out, self._jp._theta_gen
        return self._cache

    _gen = -1

    @property
    def density(self):
        return self._jp.density


class JointPosteriorRaw(JointPosterior):
    """Result of fit for a RawBuild model (`Model(params, SmolyakRaw[rule])`; reference struct JointPosteriorRaw,
    src/joint_posterior.jl:9-14,183-188): fields M, grid (.cache, .density), mu_hat, U.  The device keeps the unconstrained
    node matrix; the constrained parameters a marginal function needs are constructed from it on the device at every call
    (update!(Theta) per node, src/marginal_posterior.jl:86-90,106-115), never stored."""

    def __init__(self, M, ddata, grid, mu_hat, U, neg_min, path=_lib.PATH_AUTO, node_range=None):
        super().__init__(M, ddata, grid, mu_hat, U, neg_min, path=path, node_range=node_range, raw=True)
        self.grid_handle = self.grid
        self.grid = _RawGrid(self)


def fit(M, data, n=None, path=_lib.PATH_AUTO, mode_result=None):
    """fit(M, data[, n]): mode -> scale -> grid evaluation -> normalised weights
    (reference src/joint_posterior.jl:177-188).  n is the Smolyak level (default(B))."""
    ddata = _as_device_data(M, data)
    if n is None:
        n = default(M.build)
    mu_hat, U, neg_min = mode(M, ddata) if mode_result is None else mode_result
    U = colmajor(U)
    grid = M.ctx.grid(M.build.rule.rule_id, U.shape[1], int(n))   # cache key: (rule, rank, level), cf. index() :157-162
    cls = JointPosteriorRaw if M.build.raw else JointPosterior      # RawBuild / CacheBuild dispatch of the reference's two fit methods
    jp = cls(M, ddata, grid, mu_hat, U, neg_min, path=path)
    jp.evaluate()
    return jp
