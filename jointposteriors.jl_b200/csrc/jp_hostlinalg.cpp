// jp_hostlinalg.cpp -- host-side scale-matrix helpers of libjpcuda.so (d <= 64, negligible cost).
// These are the d x d pieces of `mode`/`deduce_scale!` that the reference keeps on the host as
// well (reference src/joint_posterior.jl:15-144).  Column-major storage throughout.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/jpcuda.h"

static thread_local char g_err[1024] = "";

void jp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

namespace {
struct ColMajor {
  double* a;
  int ld;
  double& operator()(int r, int c) const { return a[(size_t)c * ld + r]; }
};
struct ConstColMajor {
  const double* a;
  int ld;
  double operator()(int r, int c) const { return a[(size_t)c * ld + r]; }
};

// Upper Cholesky factor, one column at a time (the recurrence of chol!/try_chol!,
// reference src/joint_posterior.jl:15-43): for column c, rows r < c are obtained by forward
// substitution against the already finished columns, and the pivot loses the squares of the
// column's off-diagonal entries as they are produced.
bool upper_cholesky(ColMajor U, ConstColMajor S, int d, bool check) {
  for (int c = 0; c < d; ++c) {
    double pivot = S(c, c);
    for (int r = 0; r < c; ++r) {
      double v = S(r, c);
      for (int k = 0; k < r; ++k) v -= U(k, c) * U(k, r);
      v /= U(r, r);
      U(r, c) = v;
      pivot -= v * v;
    }
    if (check && !(pivot > 0)) {
      U(c, c) = pivot;
      return false;
    }
    U(c, c) = std::sqrt(pivot);
  }
  return true;
}

// In-place inverse of an upper-triangular matrix, row by row (inv!, reference
// src/joint_posterior.jl:56-68): entries to the left in the same row are already inverted when
// entry (r, c) is formed.
void upper_inverse(ColMajor U, int d) {
  for (int r = 0; r < d; ++r) {
    double dinv = 1.0 / U(r, r);
    U(r, r) = dinv;
    for (int c = r + 1; c < d; ++c) {
      double v = U(r, c) * dinv;
      for (int k = r + 1; k < c; ++k) v += U(k, c) * U(r, k);
      U(r, c) = v / -U(c, c);
    }
  }
}

// Symmetric eigen-decomposition by the cyclic Jacobi method (the role LAPACK plays behind
// eigfact!(Symmetric(H)), reference src/joint_posterior.jl:99).  Returns eigenvalues ascending and
// the matching eigenvectors as columns; every eigenvector is sign-normalised so that its entry of
// largest magnitude is positive (eigenvector signs are otherwise arbitrary).
void symmetric_eigen(const double* H, int d, std::vector<double>& lambda, std::vector<double>& vec) {
  std::vector<double> a(H, H + (size_t)d * d);
  vec.assign((size_t)d * d, 0.0);
  ColMajor A{a.data(), d}, V{vec.data(), d};
  for (int i = 0; i < d; ++i) V(i, i) = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int c = 0; c < d; ++c) {
      diag += A(c, c) * A(c, c);
      for (int r = 0; r < c; ++r) off += A(r, c) * A(r, c);
    }
    if (off <= 1e-32 * diag || off < 1e-300) break;
    for (int r = 0; r < d - 1; ++r)
      for (int c = r + 1; c < d; ++c) {
        double arc = A(r, c);
        if (arc == 0.0) continue;
        double tau = (A(c, c) - A(r, r)) / (2.0 * arc);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::hypot(1.0, tau));
        double cs = 1.0 / std::hypot(1.0, t), sn = t * cs;
        for (int k = 0; k < d; ++k) {  // columns r, c of A and V
          double x = A(k, r), y = A(k, c);
          A(k, r) = cs * x - sn * y;
          A(k, c) = sn * x + cs * y;
          x = V(k, r), y = V(k, c);
          V(k, r) = cs * x - sn * y;
          V(k, c) = sn * x + cs * y;
        }
        for (int k = 0; k < d; ++k) {  // rows r, c of A
          double x = A(r, k), y = A(c, k);
          A(r, k) = cs * x - sn * y;
          A(c, k) = sn * x + cs * y;
        }
      }
  }
  // selection sort of the (few) eigenpairs into ascending order
  lambda.resize(d);
  for (int i = 0; i < d; ++i) lambda[i] = A(i, i);
  for (int i = 0; i < d; ++i) {
    int best = i;
    for (int j = i + 1; j < d; ++j)
      if (lambda[j] < lambda[best]) best = j;
    if (best != i) {
      std::swap(lambda[i], lambda[best]);
      for (int k = 0; k < d; ++k) std::swap(V(k, i), V(k, best));
    }
    int big = 0;
    for (int k = 1; k < d; ++k)
      if (std::fabs(V(k, i)) > std::fabs(V(big, i))) big = k;
    if (V(big, i) < 0)
      for (int k = 0; k < d; ++k) V(k, i) = -V(k, i);
  }
}
}  // namespace

extern "C" {

const char* jp_last_error(void) { return g_err; }
int jp_version(void) { return 100; }

int jp_chol(double* U, const double* S, int d) {
  if (!U || !S || d < 1) {
    jp_set_error("jp_chol: bad argument");
    return JP_ERR_BAD_ARG;
  }
  upper_cholesky(ColMajor{U, d}, ConstColMajor{S, d}, d, false);
  return JP_OK;
}

int jp_try_chol(double* U, const double* S, int d) {
  if (!U || !S || d < 1) {
    jp_set_error("jp_try_chol: bad argument");
    return JP_ERR_BAD_ARG;
  }
  if (!upper_cholesky(ColMajor{U, d}, ConstColMajor{S, d}, d, true)) {
    jp_set_error("jp_try_chol: matrix is not positive definite");
    return JP_ERR_NOT_PD;
  }
  return JP_OK;
}

int jp_inv_upper(double* U, int d) {
  if (!U || d < 1) {
    jp_set_error("jp_inv_upper: bad argument");
    return JP_ERR_BAD_ARG;
  }
  upper_inverse(ColMajor{U, d}, d);
  return JP_OK;
}

int jp_inv_chol(double* U, const double* H, int d) {
  if (!U || !H || d < 1) {
    jp_set_error("jp_inv_chol: bad argument");
    return JP_ERR_BAD_ARG;
  }
  std::memset(U, 0, sizeof(double) * d * d);
  upper_cholesky(ColMajor{U, d}, ConstColMajor{H, d}, d, false);
  upper_inverse(ColMajor{U, d}, d);
  return JP_OK;
}

int jp_reduce_dimensions(const double* H, int d, int max_rank, double* out, int* rank) {
  if (!H || !out || !rank || d < 1) {
    jp_set_error("jp_reduce_dimensions: bad argument");
    return JP_ERR_BAD_ARG;
  }
  std::vector<double> lambda, vec;
  symmetric_eigen(H, d, lambda, vec);
  std::memset(out, 0, sizeof(double) * d * d);
  int kept = 0;
  for (int i = 0; i < d; ++i) {
    if (lambda[i] < 1e-11) continue;              // src/joint_posterior.jl:103,125
    if (max_rank > 0 && kept >= max_rank) break;  // src/joint_posterior.jl:127-128
    double s = std::sqrt(lambda[i]);
    for (int k = 0; k < d; ++k) out[(size_t)kept * d + k] = vec[(size_t)i * d + k] / s;  // :107
    ++kept;
  }
  *rank = kept;
  return JP_OK;
}

// reduce_dimensions!(M, H, LDR{g}), reference src/joint_posterior.jl:78-95,111-119: keep the leading (largest
// variance 1/lambda) admissible eigen directions until they carry the fraction g of the total variance.  The
// reference's `count` reads `total_energy` before defining it (:84, an UndefVarError); it is initialised to zero here.
int jp_reduce_dimensions_ldr(const double* H, int d, double g, double* out, int* rank) {
  if (!H || !out || !rank || d < 1 || !(g > 0.0 && g < 1.0)) {     // @assert 0.0 < g < 1.0 (:79)
    jp_set_error("jp_reduce_dimensions_ldr: bad argument (need 0 < g < 1)");
    return JP_ERR_BAD_ARG;
  }
  std::vector<double> lambda, vec;
  symmetric_eigen(H, d, lambda, vec);
  std::memset(out, 0, sizeof(double) * d * d);
  int inadmissible = 0;
  for (int i = 0; i < d; ++i)
    if (1.0 / lambda[i] >= 1e11 || !(lambda[i] > 0.0)) ++inadmissible;   // :81 (non-positive eigenvalues too)
  double total = 0.0;
  for (int i = inadmissible; i < d; ++i) total += 1.0 / lambda[i];       // :83-85
  const double limit = g * total;
  int p = 0;
  double cum = 0.0;
  while (cum < limit && inadmissible + p < d) {                          // :89-93
    cum += 1.0 / lambda[inadmissible + p];
    ++p;
  }
  for (int c = 0; c < p; ++c) {                                          // :115-117
    const int i = inadmissible + c;
    const double s = std::sqrt(lambda[i]);
    for (int k = 0; k < d; ++k) out[(size_t)c * d + k] = vec[(size_t)i * d + k] / s;
  }
  *rank = p;
  if (p == 0) {
    jp_set_error("jp_reduce_dimensions_ldr: no admissible eigen direction");
    return JP_ERR_NOT_PD;
  }
  return JP_OK;
}

int jp_deduce_scale_dynamic(const double* H, int d, double* U, int* rank) {
  if (!H || !U || !rank || d < 1) {
    jp_set_error("jp_deduce_scale_dynamic: bad argument");
    return JP_ERR_BAD_ARG;
  }
  std::memset(U, 0, sizeof(double) * d * d);
  if (upper_cholesky(ColMajor{U, d}, ConstColMajor{H, d}, d, true)) {
    upper_inverse(ColMajor{U, d}, d);
    *rank = d;
    return JP_OK;
  }
  return jp_reduce_dimensions(H, d, 0, U, rank);
}

// reference src/interp.jl:467-481: Julia's searchsortedfirst/searchsortedlast bisections are
// reproduced verbatim because weight_nodes need not be monotone (signed Smolyak weights).
static int bisect_first(const double* v, int n, double x) {
  int lo = 0, hi = n + 1;
  while (lo < hi - 1) {
    int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
    if (v[mid - 1] < x) lo = mid; else hi = mid;
  }
  return hi;
}
static int bisect_last(const double* v, int n, double x) {
  int lo = 0, hi = n + 1;
  while (lo < hi - 1) {
    int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
    if (x < v[mid - 1]) hi = mid; else lo = mid;
  }
  return lo;
}
static double lerp_between(const double* x, const double* y, int i /*1-based upper knot*/, double z) {
  return y[i - 2] + (z - x[i - 2]) * (y[i - 1] - y[i - 2]) / (x[i - 1] - x[i - 2]);
}

double jp_quantile(const double* wn, const double* vn, int n, double p) {
  if (p <= 0) return -INFINITY;
  if (p >= 1) return INFINITY;
  int i = (p < 0.5) ? bisect_first(wn, n, p) : bisect_last(wn, n, p) + 1;
  if (i < 2) i = 2;   // the reference would index out of bounds here (@inbounds-free BoundsError)
  if (i > n) i = n;
  return lerp_between(wn, vn, i, p);
}

double jp_cdf(const double* wn, const double* vn, int n, double x) {
  if (x < vn[0]) return 0.0;
  if (x > vn[n - 1]) return 1.0;
  int i = bisect_first(vn, n, x);
  if (i < 2) i = 2;   // x == value_nodes[1]: the reference indexes x[0] (BoundsError); clamp instead
  return lerp_between(vn, wn, i, x);
}

}  // extern "C"
