// jp_family.cuh -- likelihood families as CUDA device-function plugins.
//
// In the reference the likelihood is the user's Julia method `log_density(Theta, data)`, called once
// per grid node from the closures at reference src/joint_posterior.jl:147-154.  Julia closures cannot
// run on the device, so a family here is a struct of static __device__ functions that the generic
// node x observation kernel (jp_fit.cu) is instantiated with, and a registry entry mapping the
// family id of the C ABI to the instantiated launcher.  Adding a family = one struct + one
// JP_REGISTER_FAMILY line; nothing else in the library changes.
//
// Plugin concept (all static):
//   kId                      family id of include/jpcuda.h
//   kName
//   __host__ bool shape_ok(d, ncols, N)                 argument validation at fit time
//   __device__ double prior(th, d, N, hyper)            terms that do not loop over observations
//   __device__ double obs(th, d, rec, n, hyper)         log-density term of observation n;
//                                                       rec -> its ncols doubles in SHARED memory
// `th` is the constrained parameter vector of the calling thread's node, a register array of
// compile-time size DPAD >= d with zeros beyond d.
#pragma once
#include "jp_common.cuh"

#define JP_LOG_2PI 1.8378770664093454835606594728112

__device__ __forceinline__ double jp_softplus(double x) { return fmax(x, 0.0) + log1p(exp(-fabs(x))); }
__device__ __forceinline__ double jp_lpdf_normal(double x, double mu, double sd) {
  double z = (x - mu) / sd;
  return -0.5 * z * z - log(sd) - 0.5 * JP_LOG_2PI;
}

// README Example 1 (reference README.md:62-72, test/runtests.jl:19-26):
// theta = (tau, theta_minus, theta_plus); record (X, freq, NmX).
struct FamBinomialMixture {
  static constexpr int kId = JP_FAM_BINOMIAL_MIXTURE;
  static constexpr const char* kName = "binomial_mixture";
  static bool shape_ok(int d, int ncols, long long) { return d == 3 && ncols == 3; }
  template <int DPAD>
  __device__ static double prior(const double (&p)[DPAD], int, long long, const double* h) {
    return h[0] * log(p[1]) + h[1] * log(1 - p[1]) + h[2] * log(p[2]) + h[3] * log(1 - p[2]) + h[4] * log(p[0]) +
           h[5] * log(1 - p[0]);
  }
  template <int DPAD>
  __device__ static double obs(const double (&p)[DPAD], int, const double* r, long long, const double*) {
    return r[1] * log(p[0] * pow(1 - p[1], r[0]) * pow(p[1], r[2]) + (1 - p[0]) * pow(p[2], r[0]) * pow(1 - p[2], r[2]));
  }
};

// shared linear predictor: eta = sum_k x_k th_k over the padded width (th is zero beyond d)
template <int DPAD>
__device__ __forceinline__ double jp_eta(const double (&th)[DPAD], int d, const double* r) {
  double eta = 0;
#pragma unroll
  for (int k = 0; k < DPAD; ++k)
    if (k < d) eta += r[k] * th[k];
  return eta;
}

// logistic regression, beta_k ~ N(0, hyper[0]^2); record (x_0..x_{d-1}, y)
struct FamLogistic {
  static constexpr int kId = JP_FAM_LOGISTIC;
  static constexpr const char* kName = "logistic";
  static bool shape_ok(int d, int ncols, long long) { return ncols == d + 1; }
  template <int DPAD>
  __device__ static double prior(const double (&b)[DPAD], int d, long long, const double* h) {
    double lp = 0;
#pragma unroll
    for (int k = 0; k < DPAD; ++k)
      if (k < d) lp += jp_lpdf_normal(b[k], 0.0, h[0]);
    return lp;
  }
  template <int DPAD>
  __device__ static double obs(const double (&b)[DPAD], int d, const double* r, long long, const double*) {
    double eta = jp_eta(b, d, r);
    return r[d] * eta - jp_softplus(eta);
  }
};

// Poisson regression (log link), beta_k ~ N(0, hyper[0]^2); the theta-independent -lgamma(y+1) is dropped
struct FamPoisson {
  static constexpr int kId = JP_FAM_POISSON;
  static constexpr const char* kName = "poisson";
  static bool shape_ok(int d, int ncols, long long) { return ncols == d + 1; }
  template <int DPAD>
  __device__ static double prior(const double (&b)[DPAD], int d, long long, const double* h) {
    double lp = 0;
#pragma unroll
    for (int k = 0; k < DPAD; ++k)
      if (k < d) lp += jp_lpdf_normal(b[k], 0.0, h[0]);
    return lp;
  }
  template <int DPAD>
  __device__ static double obs(const double (&b)[DPAD], int d, const double* r, long long, const double*) {
    double eta = jp_eta(b, d, r);
    return r[d] * eta - exp(eta);
  }
};

// hierarchical normal ("eight schools"): theta = (mu, tau, theta_1..theta_J); record (y_j, s_j);
// y_j ~ N(theta_j, s_j^2), theta_j ~ N(mu, tau^2), flat mu, tau ~ half-Cauchy(0, hyper[0])
struct FamHierNormal {
  static constexpr int kId = JP_FAM_HIER_NORMAL;
  static constexpr const char* kName = "hier_normal";
  static bool shape_ok(int d, int ncols, long long N) { return ncols == 2 && d == (int)N + 2; }
  template <int DPAD>
  __device__ static double prior(const double (&t)[DPAD], int, long long, const double* h) {
    double r = t[1] / h[0];
    return log(2.0 / (M_PI * h[0])) - log1p(r * r);
  }
  template <int DPAD>
  __device__ static double obs(const double (&t)[DPAD], int, const double* r, long long n, const double*) {
    double tj = t[2 + n];
    return jp_lpdf_normal(r[0], tj, r[1]) + jp_lpdf_normal(tj, t[0], t[1]);
  }
};

// README Example 2 "HiWorld" (reference README.md:245-258): theta = (beta_0..beta_{p-1}, sigma);
// record (x_0..x_{p-1}, y); hyper = (sd of beta prior, sd of sigma prior)
struct FamNormalLinear {
  static constexpr int kId = JP_FAM_NORMAL_LINEAR;
  static constexpr const char* kName = "normal_linear";
  static bool shape_ok(int d, int ncols, long long) { return ncols == d; }
  template <int DPAD>
  __device__ static double prior(const double (&t)[DPAD], int d, long long, const double* h) {
    double lp = 0, sigma = 0;
#pragma unroll
    for (int k = 0; k < DPAD; ++k) {
      if (k < d - 1) lp += jp_lpdf_normal(t[k], 0.0, h[0]);
      if (k == d - 1) sigma = t[k];
    }
    return jp_lpdf_normal(sigma, 0.0, h[1]) + lp;
  }
  template <int DPAD>
  __device__ static double obs(const double (&t)[DPAD], int d, const double* r, long long, const double*) {
    double eta = 0, sigma = 0;
#pragma unroll
    for (int k = 0; k < DPAD; ++k) {
      if (k < d - 1) eta += r[k] * t[k];
      if (k == d - 1) sigma = t[k];
    }
    return jp_lpdf_normal(r[d - 1], eta, sigma);
  }
};

// category counts on a Simplex block: theta = first d = n - 1 components of a point of the n-simplex (the last is
// 1 - sum); record (count_k), one per category; symmetric Dirichlet(alpha) prior, hyper[0] = alpha - 1.
// The posterior is Dirichlet(alpha + counts) in closed form, which pins the simplex transform in the tests.
struct FamMultinomial {
  static constexpr int kId = JP_FAM_MULTINOMIAL;
  static constexpr const char* kName = "multinomial";
  static bool shape_ok(int d, int ncols, long long N) { return ncols == 1 && N == (long long)d + 1; }
  template <int DPAD>
  __device__ static double prior(const double (&t)[DPAD], int, long long, const double*) { return 0.0; }
  template <int DPAD>
  __device__ static double obs(const double (&t)[DPAD], int d, const double* r, long long n, const double* h) {
    double tn = 0, rest = 1.0;
#pragma unroll
    for (int k = 0; k < DPAD; ++k)
      if (k < d) {
        rest -= t[k];
        if (k == n) tn = t[k];
      }
    if (n >= d) tn = rest;
    return (r[0] + h[0]) * log(tn);
  }
};

// zero-mean multivariate normal with unknown covariance on a CovarianceMatrix block, inverse-Wishart(nu0, psi0 I) prior
// (lpdf_InverseWishart is among the reference's exports, src/JointPosteriors.jl:36): theta = lower triangle of Sigma row by
// row; the records are the p rows of the scatter matrix S = sum y y' (sufficient statistic); hyper = (n_obs, nu0, psi0).
//   log p = -(n + nu0 + p + 1) / 2 log|Sigma| - 1/2 tr((S + psi0 I) Sigma^-1):   Sigma | y ~ inverse-Wishart(nu0 + n, S + psi0 I)
// Every call factorises Sigma again (p <= 10, p records): this family exists to pin the transform, not for speed.
struct FamMvnCov {
  static constexpr int kId = JP_FAM_MVN_COV;
  static constexpr const char* kName = "mvn_cov";
  static constexpr int kPmax = 10;
  static bool shape_ok(int d, int ncols, long long N) { return N >= 1 && N <= kPmax && ncols == (int)N && d == (int)(N * (N + 1) / 2); }
  // Cholesky factor C (packed lower triangle) of Sigma; returns log|Sigma|
  template <int DPAD>
  __device__ static double factor(const double (&t)[DPAD], int p, double* C) {
    const int len = p * (p + 1) / 2;
#pragma unroll
    for (int j = 0; j < DPAD; ++j)
      if (j < len) C[j] = t[j];
    double logdet = 0;
    for (int i = 0; i < p; ++i)
      for (int j = 0; j <= i; ++j) {
        double v = C[i * (i + 1) / 2 + j];
        for (int k = 0; k < j; ++k) v -= C[i * (i + 1) / 2 + k] * C[j * (j + 1) / 2 + k];
        if (i == j) {
          v = sqrt(v);
          logdet += 2.0 * log(v);
        } else {
          v /= C[j * (j + 1) / 2 + j];
        }
        C[i * (i + 1) / 2 + j] = v;
      }
    return logdet;
  }
  template <int DPAD>
  __device__ static double prior(const double (&t)[DPAD], int, long long N, const double* h) {
    double C[kPmax * (kPmax + 1) / 2];
    const int p = (int)N;
    return -0.5 * (h[0] + h[1] + p + 1) * factor<DPAD>(t, p, C);
  }
  template <int DPAD>
  __device__ static double obs(const double (&t)[DPAD], int d, const double* r, long long n, const double* h) {
    // row n of Sigma^-1 against row n of S + psi0 I:  Sigma^-1 e_n by two triangular solves
    double C[kPmax * (kPmax + 1) / 2], y[kPmax];
    int p = 0;
    while ((p + 1) * (p + 2) / 2 <= d) ++p;
    factor<DPAD>(t, p, C);
    for (int i = 0; i < p; ++i) {            // C y = e_n
      double v = (i == (int)n) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) v -= C[i * (i + 1) / 2 + k] * y[k];
      y[i] = v / C[i * (i + 1) / 2 + i];
    }
    for (int i = p - 1; i >= 0; --i) {       // C' x = y
      double v = y[i];
      for (int k = i + 1; k < p; ++k) v -= C[k * (k + 1) / 2 + i] * y[k];
      y[i] = v / C[i * (i + 1) / 2 + i];
    }
    double row = 0;
    for (int j = 0; j < p; ++j) row += y[j] * (r[j] + (j == (int)n ? h[2] : 0.0));
    return -0.5 * row;
  }
};

// balanced two-factor random-effects ANOVA with the random effects integrated out (README Example 3, reference
// README.md:416-470; `TF_RE_ANOVA` of the absent LogDensities package): theta = (mu, s2_P, s2_O, s2_PO, s2_R); records
// (SS_k, df_k) for parts, operators, interaction, error and (grand mean, P O R); hyper = (P, O, R, folded-Cauchy scale of
// the operator standard deviation).  Expected mean squares and the variance of the grand mean as in oracle/jp_oracle.cpp.
struct FamAnova2 {
  static constexpr int kId = JP_FAM_ANOVA2;
  static constexpr const char* kName = "anova2";
  static bool shape_ok(int d, int ncols, long long N) { return d == 5 && ncols == 2 && N == 5; }
  template <int DPAD>
  __device__ static double lambda(const double (&t)[DPAD], int k, const double* h) {
    const double base = t[4] + h[2] * t[3];
    if (k == 0) return base + h[1] * h[2] * t[1];
    if (k == 1) return base + h[0] * h[2] * t[2];
    if (k == 2) return base;
    return t[4];
  }
  template <int DPAD>
  __device__ static double prior(const double (&t)[DPAD], int, long long, const double* h) {
    const double so = sqrt(t[2]) / h[3];
    return -log1p(so * so) - 0.5 * log(t[2]);
  }
  template <int DPAD>
  __device__ static double obs(const double (&t)[DPAD], int, const double* r, long long n, const double* h) {
    if (n < 4) {
      const double lam = lambda<DPAD>(t, (int)n, h);
      return -0.5 * r[1] * log(lam) - 0.5 * r[0] / lam;
    }
    const double vm = (lambda<DPAD>(t, 0, h) + lambda<DPAD>(t, 1, h) - lambda<DPAD>(t, 2, h)) / r[1];
    return -0.5 * log(vm) - 0.5 * (r[0] - t[0]) * (r[0] - t[0]) / vm;
  }
};

// ------------------------------------------------------------------------------------ registry
struct JpFitLaunchParams;   // defined in jp_fit.cu
typedef int (*jp_family_launcher)(jp_posterior* post, const JpFitLaunchParams& lp);
typedef bool (*jp_family_shape_ok)(int d, int ncols, long long N);
struct JpFamilyEntry {
  int id;
  const char* name;
  jp_family_launcher launch;
  jp_family_shape_ok shape_ok;
};
void jp_register_family(const JpFamilyEntry& e);
const JpFamilyEntry* jp_find_family(int id);
