"""B200-native posterior integration behind the JointPosteriors.jl API (Model / fit / marginal).

Everything numerical runs in libjpcuda.so (hand-written sm_100a CUDA, C ABI in include/jpcuda.h);
this package is the host-side mirror of the reference's Julia interface and binds the library with
ctypes.  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from ._lib import JPError, NotPositiveDefinite, PATH_AUTO, PATH_FP64, PATH_TC, lib
from .data import (BinaryClassificationData, Data, HierNormalData, LogisticData, MultinomialData, MvNormalCovData,
                   NormalLinearData, PoissonData, TwoFactorANOVAData)
from .linalg import chol, deduce_scale_dynamic, inv_chol, inv_upper, reduce_dimensions, reduce_dimensions_ldr, try_chol
from .marginals import (Grid, MarginalBuffer, NestedPolyGLM, Normal, cdf, marginal, marginal_buffer, marginal_smooth, marginals,
                        pdf, quantile)
from .model import (Context, DeviceData, Dynamic, FixedRank, Full, LDR, GenzKeister, JointPosterior, JointPosteriorRaw,
                    KronrodPatterson, Model, Smolyak, SmolyakRaw, default, fit, log_density_unc, mode)
from .distributed import Comm, ObsShardedPosterior, ShardedPosterior, fit_distributed, fit_obs_sharded, mode_p2p
from .params import CovarianceMatrix, NonCentredVector, PositiveVector, ProbabilityVector, RealVector, Simplex, parameter

__all__ = [
    "Model", "fit", "marginal", "marginals", "marginal_buffer", "MarginalBuffer", "mode", "quantile", "cdf", "pdf", "Grid", "Normal", "NestedPolyGLM", "marginal_smooth", "JointPosterior", "JointPosteriorRaw",
    "parameter", "RealVector", "PositiveVector", "ProbabilityVector", "Simplex", "NonCentredVector", "Data", "BinaryClassificationData",
    "LogisticData", "PoissonData", "HierNormalData", "NormalLinearData", "MultinomialData", "Smolyak", "SmolyakRaw", "GenzKeister",
    "KronrodPatterson", "Dynamic", "Full", "FixedRank", "LDR", "default", "Context", "DeviceData", "chol", "try_chol",
    "inv_upper", "inv_chol", "reduce_dimensions", "reduce_dimensions_ldr", "deduce_scale_dynamic", "JPError", "NotPositiveDefinite",
    "PATH_AUTO", "PATH_FP64", "PATH_TC", "log_density_unc", "fit_distributed", "ShardedPosterior", "ObsShardedPosterior",
    "fit_obs_sharded", "mode_p2p", "Comm", "CovarianceMatrix", "MvNormalCovData", "TwoFactorANOVAData",
]
