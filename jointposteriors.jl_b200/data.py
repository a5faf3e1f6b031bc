"""Data containers: one class per registered likelihood family (csrc/jp_family.cuh).

In the reference the data struct and `log_density(Theta, data)` are user code (reference
README.md:42-72); on the GPU the likelihood must be a registered device-function plugin, so the data
class names the family and lays the observations out as the N x ncols row-major record array the
plugin reads.
"""
import numpy as np

FAM_BINOMIAL_MIXTURE, FAM_LOGISTIC, FAM_POISSON, FAM_HIER_NORMAL, FAM_NORMAL_LINEAR, FAM_MULTINOMIAL, FAM_MVN_COV, FAM_ANOVA2 = range(8)


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


class MvNormalCovData(Data):
    """Zero-mean multivariate normal observations y_i in R^p with unknown covariance Sigma (a CovarianceMatrix(p) block) and
    an inverse-Wishart(nu0, psi0 I) prior (lpdf_InverseWishart: reference src/JointPosteriors.jl:36).  The records are the p
    rows of the scatter matrix S = sum_i y_i y_i' (sufficient statistic); the posterior is inverse-Wishart(nu0 + n, S + psi0 I)
    in closed form, which is what pins the covariance-matrix transform."""
    family = FAM_MVN_COV

    def __init__(self, Y, nu0=None, psi0=1.0):
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim != 2 or Y.shape[0] < 1 or not 1 <= Y.shape[1] <= 10:
            raise ValueError("Y must be n x p with 1 <= p <= 10")
        self.n, self.p = Y.shape
        self.S = Y.T @ Y
        self.nu0 = float(self.p + 2 if nu0 is None else nu0)
        self.psi0 = float(psi0)

    def records(self):
        return np.ascontiguousarray(self.S), np.array([float(self.n), self.nu0, self.psi0])

    def posterior_mean(self):
        """E[Sigma | y] of the inverse-Wishart posterior."""
        return (self.S + self.psi0 * np.eye(self.p)) / (self.nu0 + self.n - self.p - 1)


class TwoFactorANOVAData(Data):
    """Balanced two-factor random-effects ANOVA (README Example 3, reference README.md:416-470: `TF_RE_ANOVA_Data(y, yp, yo)`
    of the absent LogDensities package): y[i] measured on part yp[i] by operator yo[i], R replicates per cell.  Parameters
    (mu, s2_P, s2_O, s2_PO, s2_R) = RealVector(1) + PositiveVector(4); the random effects are integrated out, the records are
    the four ANOVA sums of squares with their degrees of freedom and the grand mean; folded-Cauchy(cauchy_scale) prior on the
    operator standard deviation, improper flat priors elsewhere (as the README describes the model)."""
    family = FAM_ANOVA2

    def __init__(self, y, yp, yo, cauchy_scale=20.0):
        y = np.asarray(y, dtype=np.float64)
        yp = np.asarray(yp, dtype=np.int64)
        yo = np.asarray(yo, dtype=np.int64)
        if not (y.ndim == yp.ndim == yo.ndim == 1 and len(y) == len(yp) == len(yo) and len(y) > 0):
            raise ValueError("y, yp, yo must be 1-D arrays of equal length")
        ps, os_ = np.unique(yp), np.unique(yo)
        P, Oo = len(ps), len(os_)
        cells = np.zeros((P, Oo))
        counts = np.zeros((P, Oo), dtype=np.int64)
        pi, oi = np.searchsorted(ps, yp), np.searchsorted(os_, yo)
        np.add.at(cells, (pi, oi), y)
        np.add.at(counts, (pi, oi), 1)
        R = int(counts[0, 0])
        if R < 2 or np.any(counts != R) or P < 2 or Oo < 2:
            raise ValueError("TwoFactorANOVAData needs a balanced design with >= 2 parts, operators and replicates")
        cm = cells / R
        gm = cm.mean()
        pm, om = cm.mean(axis=1), cm.mean(axis=0)
        ss_p = Oo * R * float(np.sum((pm - gm) ** 2))
        ss_o = P * R * float(np.sum((om - gm) ** 2))
        ss_po = R * float(np.sum((cm - pm[:, None] - om[None, :] + gm) ** 2))
        ss_e = float(np.sum((y - cm[pi, oi]) ** 2))
        self.P, self.O, self.R, self.cauchy_scale = P, Oo, R, float(cauchy_scale)
        self._obs = np.array([[ss_p, P - 1.0], [ss_o, Oo - 1.0], [ss_po, (P - 1.0) * (Oo - 1.0)], [ss_e, P * Oo * (R - 1.0)],
                              [gm, float(P * Oo * R)]])

    def records(self):
        return self._obs, np.array([float(self.P), float(self.O), float(self.R), self.cauchy_scale])
