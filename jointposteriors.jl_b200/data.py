"""Data containers: one class per registered likelihood family (csrc/jp_family.cuh).

In the reference the data struct and `log_density(Theta, data)` are user code (reference
README.md:42-72); on the GPU the likelihood must be a registered device-function plugin, so the data
class names the family and lays the observations out as the N x ncols row-major record array the
plugin reads.
"""
import numpy as np

FAM_BINOMIAL_MIXTURE, FAM_LOGISTIC, FAM_POISSON, FAM_HIER_NORMAL, FAM_NORMAL_LINEAR, FAM_MULTINOMIAL = range(6)


class Data:
    family = None
    sample_size_order = 1   # reference src/joint_posterior.jl:157 (part of the grid-cache key)

    def records(self):
        """-> (obs float64 [N, ncols] C-contiguous, hyper float64 [n_hyper])"""
        raise NotImplementedError


class BinaryClassificationData(Data):
    """reference README.md:42-52,76-79 / test/runtests.jl:7-30: X successes out of n, freq parts each."""
    family = FAM_BINOMIAL_MIXTURE

    def __init__(self, X, freq, n, αm=1.0, βm=1.0, αp=1.0, βp=1.0, ατ=1.0, βτ=1.0, **kw):
        # ASCII aliases: am, bm, ap, bp, at, bt
        αm = kw.pop("am", αm); βm = kw.pop("bm", βm); αp = kw.pop("ap", αp)
        βp = kw.pop("bp", βp); ατ = kw.pop("at", ατ); βτ = kw.pop("bt", βτ)
        if kw:
            raise TypeError("unexpected keyword(s): %s" % ", ".join(kw))
        self.X = np.asarray(X, dtype=np.int64)
        self.freq = np.asarray(freq, dtype=np.int64)
        if self.X.shape != self.freq.shape or self.X.ndim != 1 or self.X.size == 0:
            raise ValueError("X and freq must be non-empty 1-D arrays of equal length")
        self.NmX = int(n) - self.X
        self.prior_m1 = np.array([αm - 1, βm - 1, αp - 1, βp - 1, ατ - 1, βτ - 1], dtype=np.float64)

    def records(self):
        obs = np.stack([self.X, self.freq, self.NmX], axis=1).astype(np.float64)
        return np.ascontiguousarray(obs), self.prior_m1.copy()


class _GLMData(Data):
    def __init__(self, X, y, prior_sd=10.0):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if X.ndim != 2 or y.ndim != 1 or X.shape[0] != y.shape[0] or X.shape[0] == 0:
            raise ValueError("X must be N x d and y of length N, N >= 1")
        self.X, self.y, self.prior_sd = X, y, float(prior_sd)
        self._obs = None

    def records(self):
        if self._obs is None:
            self._obs = np.ascontiguousarray(np.concatenate([self.X, self.y[:, None]], axis=1))
        return self._obs, np.array([self.prior_sd])


class LogisticData(_GLMData):
    """y_n ~ Bernoulli(logistic(x_n . beta)), beta_k ~ N(0, prior_sd^2)."""
    family = FAM_LOGISTIC


class PoissonData(_GLMData):
    """y_n ~ Poisson(exp(x_n . beta)), beta_k ~ N(0, prior_sd^2)."""
    family = FAM_POISSON


class HierNormalData(Data):
    """y_j ~ N(theta_j, s_j^2), theta_j ~ N(mu, tau^2), flat mu, tau ~ half-Cauchy(0, tau_scale)."""
    family = FAM_HIER_NORMAL

    def __init__(self, y, s, tau_scale=25.0):
        self.y = np.asarray(y, dtype=np.float64)
        self.s = np.asarray(s, dtype=np.float64)
        if self.y.shape != self.s.shape or self.y.ndim != 1 or self.y.size == 0:
            raise ValueError("y and s must be non-empty 1-D arrays of equal length")
        self.tau_scale = float(tau_scale)

    def records(self):
        return np.ascontiguousarray(np.stack([self.y, self.s], axis=1)), np.array([self.tau_scale])


class NormalLinearData(Data):
    """reference README.md:250-258 (HiWorld): y ~ N(X beta, sigma), beta ~ N(0, sd_beta), sigma ~ N(0, sd_sigma)."""
    family = FAM_NORMAL_LINEAR

    def __init__(self, X, y, sd_beta=10.0, sd_sigma=1.0):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if X.ndim != 2 or y.ndim != 1 or X.shape[0] != y.shape[0] or X.shape[0] == 0:
            raise ValueError("X must be N x p and y of length N, N >= 1")
        self.X, self.y = X, y
        self.sd_beta, self.sd_sigma = float(sd_beta), float(sd_sigma)

    def records(self):
        obs = np.ascontiguousarray(np.concatenate([self.X, self.y[:, None]], axis=1))
        return obs, np.array([self.sd_beta, self.sd_sigma])


class MultinomialData(Data):
    """Category counts c_1..c_n with theta on the n-simplex (a Simplex(n) block) and a symmetric Dirichlet(alpha)
    prior; the posterior is Dirichlet(alpha + c) in closed form, which is what pins the simplex transform."""
    family = FAM_MULTINOMIAL

    def __init__(self, counts, alpha=1.0):
        self.counts = np.asarray(counts, dtype=np.float64)
        if self.counts.ndim != 1 or self.counts.size < 2 or np.any(self.counts < 0):
            raise ValueError("counts must be a 1-D array of at least two non-negative numbers")
        self.alpha = float(alpha)
        if self.alpha <= 0:
            raise ValueError("alpha must be positive")

    def records(self):
        return np.ascontiguousarray(self.counts[:, None]), np.array([self.alpha - 1.0])
