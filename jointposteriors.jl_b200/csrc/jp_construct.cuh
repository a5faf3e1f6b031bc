// jp_construct.cuh -- the constraint transforms of stage 2 as device functions: construct / update! of ConstrainedParameters
// (reference src/joint_posterior.jl:148,152; the types it re-exports, src/JointPosteriors.jl:22-26).  Shared by the node kernel
// of the FP64 path (jp_fit.cu) and the device-resident mode finder (jp_mode_dev.cu).
#pragma once
#include "jp_common.cuh"

#define JP_COVMAT_MAX_LEN 55          // lower triangle of a 10 x 10 covariance matrix (JP_MAX_D = 64 coordinates in all)

__device__ __forceinline__ double jp_transform(int code, double x, double& lj) {
  if (code == JP_T_POSITIVE) {
    lj += x;
    return exp(x);
  }
  if (code == JP_T_PROBABILITY) {
    double ex = exp(x);
    lj -= log(2.0 + ex + 1.0 / ex);     // sign convention of nlogit_lj, reference src/interp.jl:321-324
    return 1.0 / (1.0 + exp(-x));
  }
  return x;
}

// construct / update! of ConstrainedParameters (reference src/joint_posterior.jl:148,152): unconstrained -> constrained in
// place (register array, compile-time indices), returns log|J|.  Simplex blocks first, then coordinate by coordinate.
template <int DPAD>
__device__ __forceinline__ double jp_construct(double (&th)[DPAD], int d, const int* s_code) {
  double lj = 0.0;
  // simplex blocks first (each depends only on its own unconstrained coordinates): a run-time loop over block
  // heads, compile-time indices into the register array inside
#pragma unroll 1
  for (int k0 = 0; k0 < d; ++k0) {
    const int code = s_code[k0];
    if (JP_T_KIND(code) != JP_T_SIMPLEX || JP_T_LOC(code) != k0) continue;
    const int hi = k0 + JP_T_SCALE(code);
    double mx = 0.0;                               // the implied last coordinate has x = 0
#pragma unroll
    for (int j = 0; j < DPAD; ++j)
      if (j >= k0 && j < hi) mx = fmax(mx, th[j]);
    double S = exp(-mx);
#pragma unroll
    for (int j = 0; j < DPAD; ++j)
      if (j >= k0 && j < hi) S += exp(th[j] - mx);
    const double logS = log(S);
    lj += -mx - logS;                              // log of the implied last component
#pragma unroll
    for (int j = 0; j < DPAD; ++j)
      if (j >= k0 && j < hi) {
        const double l = th[j] - mx - logS;
        lj += l;
        th[j] = exp(l);
      }
  }
  // covariance-matrix blocks: Sigma = L L' from the log-Cholesky coordinates.  The block is copied to a small local array
  // (run-time indices; only models that declare such a block ever execute this loop body).
#pragma unroll 1
  for (int k0 = 0; k0 < d; ++k0) {
    const int code = s_code[k0];
    if (JP_T_KIND(code) != JP_T_COVMAT || JP_T_LOC(code) != k0) continue;
    const int len = JP_T_SCALE(code);
    int p = 0;
    while ((p + 1) * (p + 2) / 2 <= len) ++p;
    double Lm[JP_COVMAT_MAX_LEN];
#pragma unroll
    for (int j = 0; j < DPAD; ++j)
      if (j >= k0 && j < k0 + len) Lm[j - k0] = th[j];
    for (int i = 0; i < p; ++i) {
      const int e = i * (i + 1) / 2 + i;
      lj += (p - i + 1) * Lm[e];
      Lm[e] = exp(Lm[e]);
    }
    lj += p * 0.69314718055994530942;
    double Sg[JP_COVMAT_MAX_LEN];
    for (int i = 0, e = 0; i < p; ++i)
      for (int j = 0; j <= i; ++j, ++e) {
        double v = 0;
        for (int k = 0; k <= j; ++k) v += Lm[i * (i + 1) / 2 + k] * Lm[j * (j + 1) / 2 + k];
        Sg[e] = v;
      }
#pragma unroll
    for (int j = 0; j < DPAD; ++j)
      if (j >= k0 && j < k0 + len) th[j] = Sg[j - k0];
  }
  int last_scale = -1;        // a hierarchical block's coordinates share one scale: its logarithm is taken once
  double last_log = 0.0;
#pragma unroll
  for (int k = 0; k < DPAD; ++k) {
    if (k < d) {
      const int code = s_code[k];
      if (JP_T_KIND(code) == JP_T_SIMPLEX || JP_T_KIND(code) == JP_T_COVMAT) {
        // transformed above
      } else if (JP_T_KIND(code) == JP_T_NONCENTRED) {
        // theta_k = theta_loc + theta_scale * x_k with loc, scale < k already transformed; the register
        // array is searched with compile-time indices so that it never spills to local memory
        const int il = JP_T_LOC(code), is = JP_T_SCALE(code);
        double loc = 0.0, sc = 1.0;
#pragma unroll
        for (int j = 0; j < k; ++j) {
          if (j == il) loc = th[j];
          if (j == is) sc = th[j];
        }
        th[k] = fma(sc, th[k], loc);
        if (is != last_scale) {
          last_log = log(sc);
          last_scale = is;
        }
        lj += last_log;
      } else {
        th[k] = jp_transform(code, th[k], lj);
      }
    }
  }
  return lj;
}
