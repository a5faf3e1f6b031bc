"""Synthetic inputs of the BASELINE.json configurations.

Generators are counter-based (numpy Philox, keyed by the config's seed) so every rank and both
bench arms rebuild identical data without communication.
"""
import numpy as np

from .data import BinaryClassificationData, HierNormalData, LogisticData, PoissonData
from .params import NonCentredVector, PositiveVector, ProbabilityVector, RealVector


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=seed))


def cfg1_binary_classification():
    """README Example 1 (reference README.md:82-86)."""
    data = BinaryClassificationData([0, 1, 2, 3, 4, 7, 8, 9], [10, 2, 2, 1, 2, 3, 2, 16], 9, βm=2, βp=2)
    return dict(name="cfg1", params=(ProbabilityVector(3),), data=data, level=5, d=3)


def cfg2_eight_schools():
    data = HierNormalData([28, 8, -3, 7, -1, 1, 18, 12], [15, 10, 16, 11, 9, 11, 10, 18], tau_scale=25.0)
    return dict(name="cfg2", params=(RealVector(1), PositiveVector(1), NonCentredVector(8, loc=0, scale=1)), data=data, level=5, d=10)


def _glm(seed, d, N, kind, xscale=1.0, chunk=1 << 20):
    rng = _rng(seed)
    beta = rng.standard_normal(d) / np.sqrt(d)
    X = np.empty((N, d))
    y = np.empty(N)
    for b in range(0, N, chunk):
        e = min(N, b + chunk)
        Xb = rng.standard_normal((e - b, d)) * xscale
        Xb[:, 0] = 1.0
        eta = Xb @ beta
        if kind == "logistic":
            yb = (rng.random(e - b) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
        else:
            yb = rng.poisson(np.exp(eta)).astype(np.float64)
        X[b:e] = Xb
        y[b:e] = yb
    return X, y, beta


def cfg3_logistic(N=100_000, d=10, level=6):
    X, y, beta = _glm(3, d, N, "logistic")
    return dict(name="cfg3", params=(RealVector(d),), data=LogisticData(X, y, 10.0), level=level, d=d, beta_true=beta)


def cfg4_poisson(N=1_000_000, d=20, level=5):
    X, y, beta = _glm(4, d, N, "poisson", xscale=0.3)
    return dict(name="cfg4", params=(RealVector(d),), data=PoissonData(X, y, 10.0), level=level, d=d, beta_true=beta)


def cfg5_logistic(N=10_000_000, d=30, level=4):
    X, y, beta = _glm(5, d, N, "logistic")
    return dict(name="cfg5", params=(RealVector(d),), data=LogisticData(X, y, 10.0), level=level, d=d, beta_true=beta)


WORKLOADS = {
    "cfg1": cfg1_binary_classification,
    "cfg2": cfg2_eight_schools,
    "cfg3": cfg3_logistic,
    "cfg4": cfg4_poisson,
    "cfg5": cfg5_logistic,
}
