"""Parameter blocks and their constraint transforms (host-side descriptors).

Mirrors the ConstrainedParameters types the reference re-exports (reference
src/JointPosteriors.jl:22-26) as used in its two model-declaration styles:
  * tuple API     `model = (ProbabilityVector(3),)`                   reference test/runtests.jl:5
  * struct API    `struct BinaryClassification{T} <: parameter{T}; x::Vector{T}; p::ProbabilityVector{3,T}; end`
                                                                       reference README.md:30-33
The transforms themselves run on the GPU (csrc/jp_fit.cu, jp_transform); this module only produces
the per-coordinate transform codes of include/jpcuda.h and gives names to slices of Theta.
"""
import numpy as np

T_REAL, T_POSITIVE, T_PROBABILITY, T_NONCENTRED, T_SIMPLEX, T_COVMAT = 0, 1, 2, 3, 4, 5


class _Block:
    code = T_REAL

    def __init__(self, n):
        n = int(n)
        if n < 1:
            raise ValueError("parameter block length must be >= 1")
        self.n = n

    def __repr__(self):
        return "%s(%d)" % (type(self).__name__, self.n)


class RealVector(_Block):
    """theta = x."""
    code = T_REAL


class PositiveVector(_Block):
    """theta = exp(x), log|J| = x."""
    code = T_POSITIVE


class ProbabilityVector(_Block):
    """theta = logistic(x), log|J| = -log(2 + e^x + e^-x) (sign convention of reference src/interp.jl:321-324)."""
    code = T_PROBABILITY


class NonCentredVector(_Block):
    """theta_k = theta[loc] + theta[scale] * x_k, log|J| = n log(theta[scale]): the non-centred
    parameterisation of a hierarchical block.  `loc` and `scale` are flat indices of EARLIER constrained
    coordinates.  (Not a ConstrainedParameters type: the centred eight-schools posterior of BASELINE
    config 2 has no joint mode -- its density is unbounded as tau -> 0 -- so a Laplace-centred grid needs it.)"""
    code = T_NONCENTRED

    def __init__(self, n, loc=0, scale=1):
        super().__init__(n)
        self.loc, self.scale = int(loc), int(scale)

    @property
    def code_word(self):
        return T_NONCENTRED | (self.loc << 8) | (self.scale << 16)


class Simplex(_Block):
    """A point of the n-simplex (ConstrainedParameters Simplex, reference src/JointPosteriors.jl:26): n - 1
    unconstrained coordinates, theta_k = e^{x_k} / (1 + sum_j e^{x_j}) for the first n - 1 components, the last one
    implied (1 - sum); log|J| = sum of the logs of all n components.  The block occupies n - 1 coordinates of Theta;
    marginal functions see all n components (ParamView appends the implied one)."""
    code = T_SIMPLEX

    def __init__(self, n):
        n = int(n)
        if n < 2:
            raise ValueError("a simplex has at least 2 components")
        super().__init__(n - 1)
        self.components = n

    def __repr__(self):
        return "Simplex(%d)" % self.components


class CovarianceMatrix(_Block):
    """A p x p covariance matrix (ConstrainedParameters CovarianceMatrix, reference src/JointPosteriors.jl:22):
    p (p + 1) / 2 unconstrained coordinates = the log-Cholesky factor (lower triangle row by row, diagonal on the log scale);
    constrained = the lower triangle of Sigma = L L' in the same order; log|J| = p log 2 + sum_i (p - i + 1) x_ii.
    Marginal functions receive the full symmetric p x p matrix (ParamView expands the triangle)."""
    code = T_COVMAT

    def __init__(self, p):
        p = int(p)
        if p < 1 or p > 10:
            raise ValueError("CovarianceMatrix: 1 <= p <= 10")
        super().__init__(p * (p + 1) // 2)
        self.p = p

    def __repr__(self):
        return "CovarianceMatrix(%d)" % self.p


def tril_index(p):
    """(rows, cols) of the packed lower triangle, row by row: (0,0), (1,0), (1,1), (2,0), .."""
    r = np.array([i for i in range(p) for j in range(i + 1)])
    c = np.array([j for i in range(p) for j in range(i + 1)])
    return r, c


class parameter:
    """Base class for the struct API: subclasses list their blocks as class attributes, in order.

        class BinaryClassification(parameter):
            p = ProbabilityVector(3)
    (the reference's mandatory first field `x::Vector{T}` is the unconstrained storage; here it is implicit)
    """


def blocks_of(spec):
    """Normalise a model specification to an ordered list of (name, block)."""
    if isinstance(spec, type) and issubclass(spec, parameter):
        out = [(k, v) for k, v in vars(spec).items() if isinstance(v, _Block)]
        if not out:
            raise ValueError("parameter struct %s declares no parameter blocks" % spec.__name__)
        return out
    if isinstance(spec, _Block):
        spec = (spec,)
    if isinstance(spec, (tuple, list)) and all(isinstance(b, _Block) for b in spec) and len(spec) > 0:
        return [("p%d" % (i + 1), b) for i, b in enumerate(spec)]
    raise TypeError("model must be a parameter subclass or a tuple of RealVector/PositiveVector/ProbabilityVector/Simplex/CovarianceMatrix/NonCentredVector")


def transform_codes(blocks):
    out, o = [], 0
    for _, b in blocks:
        if isinstance(b, Simplex):
            word = T_SIMPLEX | (o << 8) | (b.n << 16)      # JP_T_SIMPLEX_CODE(first, len) of include/jpcuda.h
        elif isinstance(b, CovarianceMatrix):
            word = T_COVMAT | (o << 8) | (b.n << 16)       # JP_T_COVMAT_CODE(first, len)
        else:
            word = getattr(b, "code_word", b.code)
        out.append(np.full(b.n, word, dtype=np.int32))
        o += b.n
    return np.concatenate(out)


class ParamView:
    """What a marginal function f(Theta) receives: named blocks of the constrained parameters.

    Field k is an array of shape (n_k, M) (or (n_k,) for a single node); the tuple API's positional
    splat `f(j.Theta...)` (reference src/marginal_posterior.jl:102) corresponds to `view.blocks`.
    """

    def __init__(self, blocks, theta):
        self._names = []
        self.blocks = []
        o = 0
        for name, b in blocks:
            v = theta[o:o + b.n]
            if isinstance(b, Simplex):      # all n components: the implied last one is 1 - sum of the stored ones
                v = np.concatenate([v, 1.0 - np.sum(v, axis=0, keepdims=True)], axis=0)
            if isinstance(b, CovarianceMatrix):      # the symmetric p x p matrix (x nodes) from its stored lower triangle
                r, c = tril_index(b.p)
                full = np.zeros((b.p, b.p) + v.shape[1:], dtype=v.dtype)
                full[r, c] = v
                full[c, r] = v
                v = full
            setattr(self, name, v)
            self._names.append(name)
            self.blocks.append(v)
            o += b.n

    def __getitem__(self, i):  # tuple API with one block: p[1] style access falls through to the block
        if len(self.blocks) == 1:
            return self.blocks[0][i]
        return self.blocks[i]


class _Tracer:
    """Stand-in for one constrained coordinate while a marginal function is probed.  It supports NO arithmetic, comparison or
    conversion: a function that does anything to a parameter except hand it back raises and is treated as a general closure."""
    __slots__ = ("index",)

    def __init__(self, index):
        self.index = index

    def __array__(self, *a, **k):      # numpy must not silently turn it into a number either
        raise TypeError("not a number")


def probe_coordinate(f, blocks):
    """Return the flat coordinate index if f(Theta) merely SELECTS one coordinate, else None.

    f is called once on a ParamView-like object whose entries are opaque tracer objects: only a function that returns one
    of them untouched is a selector (and its marginal is then a zero-copy column of the Theta array on the device).  Any
    arithmetic, comparison (`abs`, `max(p[0], 0)`, `np.clip`, a branch on the value ...) fails on a tracer, so such an f is
    evaluated on the host at every node, as the reference does for every f (src/marginal_posterior.jl:98-105)."""
    view = ParamView.__new__(ParamView)
    view._names, view.blocks = [], []
    o = 0
    for name, b in blocks:
        v = np.empty(b.n + (1 if isinstance(b, Simplex) else 0), dtype=object)
        for i in range(b.n):
            v[i] = _Tracer(o + i)
        if isinstance(b, Simplex):
            v[b.n] = _Tracer(None)      # the implied last component is not a stored coordinate
        if isinstance(b, CovarianceMatrix):      # Sigma[i, j] and Sigma[j, i] are the same stored coordinate
            r, c = tril_index(b.p)
            full = np.empty((b.p, b.p), dtype=object)
            full[r, c] = v
            full[c, r] = v
            v = full
        setattr(view, name, v)
        view._names.append(name)
        view.blocks.append(v)
        o += b.n
    try:
        r = f(view)
    except Exception:
        return None
    if isinstance(r, np.ndarray) and r.dtype == object and r.size == 1:
        r = r.reshape(-1)[0]
    if isinstance(r, _Tracer) and r.index is not None:
        return int(r.index)
    return None
