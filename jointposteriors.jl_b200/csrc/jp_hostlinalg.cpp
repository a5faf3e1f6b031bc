// jp_hostlinalg.cpp -- host-side scale-matrix helpers of libjpcuda.so (d <= 64, negligible cost).
// These are the d x d pieces of `mode`/`deduce_scale!` that the reference keeps on the host as
// well (reference src/joint_posterior.jl:15-144).  Column-major storage throughout.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/jpcuda.h"
// score / information sums of a GLM, optionally added over the ranks of a communicator (csrc/jp_glm.cu)
int jp_glm_grad_hess_comm(jp_ctx* ctx, const jp_data* data, jp_comm* comm, int d, const double* h_beta, double* h_g, double* h_Hneg,
                          double* h_logpost);

int jp_mode_dev_try(jp_ctx* ctx, const jp_data* data, int d, const int* h_transform, double* h_x, double* h_H, double* fx,
                    int iters, int* evals, int* iterations, int* converged, double* grad, int* used);      // jp_mode_dev.cu
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


// ---------------------------------------------------------------------------------------------------------
// mode(M, data): reference src/joint_posterior.jl:164-168 (optBFGS! + ForwardDiff Hessian there).  Here: Newton
// iterations whose every function / derivative evaluation is one batched GPU call, driven from native host code
// (a Newton iteration costs one kernel round trip; the loop itself is d x d arithmetic).
//   * GLM families with unconstrained coefficients: analytic score and information from jp_glm_grad_hess.
//   * every other family: saddle-free Newton on Richardson (h, h/2) central differences of jp_log_density_points.
//     The stencil is evaluated AT THE CANDIDATE x + t step, so its centre value decides the line search and its
//     derivatives start the next iteration.  The Hessian's eigenvalues are replaced by their absolute values so
//     that indefinite regions are descended; at a stationary point with negative curvature the iterate leaves
//     along the most negative eigenvector.
// h_x: in = start, out = mode (unconstrained); h_H: Hessian of the NEGATIVE log-density at the mode, d x d symmetric;
// *neg_min: the minimised objective; *evals: batched GPU evaluations used.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct FdStencil {
  int d = 0, n = 0;                 // n points per step size (without the centre)
  std::vector<double> off;          // n x d offsets in units of h
  explicit FdStencil(int d_) : d(d_) {
    n = 2 * d + 2 * d * (d - 1);
    off.assign((size_t)n * d, 0.0);
    for (int i = 0; i < d; ++i) {
      off[(size_t)(2 * i) * d + i] = 1.0;
      off[(size_t)(2 * i + 1) * d + i] = -1.0;
    }
    int q = 2 * d;
    static const double sg[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
    for (int i = 0; i < d; ++i)
      for (int j = i + 1; j < d; ++j)
        for (int k = 0; k < 4; ++k, ++q) {
          off[(size_t)q * d + i] = sg[k][0];
          off[(size_t)q * d + j] = sg[k][1];
        }
  }
  // O(h^2) gradient / Hessian from the values f[0 .. n) on the stencil
  void derivs(double f0, const double* f, double h, double* g, double* H) const {
    for (int i = 0; i < d; ++i) {
      const double fp = f[2 * i], fm = f[2 * i + 1];
      g[i] = (fp - fm) / (2 * h);
      H[(size_t)i * d + i] = (fp - 2 * f0 + fm) / (h * h);
    }
    int q = 2 * d;
    for (int i = 0; i < d; ++i)
      for (int j = i + 1; j < d; ++j, q += 4) {
        const double v = (f[q] - f[q + 1] - f[q + 2] + f[q + 3]) / (4 * h * h);
        H[(size_t)j * d + i] = H[(size_t)i * d + j] = v;
      }
  }
};

struct ModeProblem {
  jp_ctx* ctx;
  const jp_data* data;
  int d;
  const int* code;
  int evals = 0;
  FdStencil st;
  std::vector<double> pts, val, g1, H1, g2, H2;
  ModeProblem(jp_ctx* c, const jp_data* dt, int d_, const int* cd) : ctx(c), data(dt), d(d_), code(cd), st(d_) {
    pts.resize((size_t)(1 + 2 * st.n) * d);
    val.resize(1 + 2 * st.n);
    g1.resize(d); g2.resize(d); H1.resize((size_t)d * d); H2.resize((size_t)d * d);
  }
  // negative log-density at K points
  int values(const double* x, int K, double* f) {
    int s = jp_log_density_points(ctx, data, d, code, K, x, f);
    ++evals;
    for (int k = 0; k < K; ++k) f[k] = -f[k];
    return s;
  }
  // f(x) and O(h^4) derivatives from the stencils at h and h/2: one batched call
  int richardson(const double* x, double h, double* f0, double* g, double* H) {
    const int n = st.n;
    for (int k = 0; k < d; ++k) pts[k] = x[k];
    for (int s = 0; s < 2; ++s) {
      const double hh = s ? h / 2 : h;
      double* p = pts.data() + (size_t)(1 + s * n) * d;
      for (int q = 0; q < n; ++q)
        for (int k = 0; k < d; ++k) p[(size_t)q * d + k] = x[k] + hh * st.off[(size_t)q * d + k];
    }
    int s = values(pts.data(), 1 + 2 * n, val.data());
    if (s != JP_OK) return s;
    *f0 = val[0];
    st.derivs(val[0], val.data() + 1, h, g1.data(), H1.data());
    st.derivs(val[0], val.data() + 1 + n, h / 2, g2.data(), H2.data());
    for (int k = 0; k < d; ++k) g[k] = (4 * g2[k] - g1[k]) / 3;
    for (int k = 0; k < d * d; ++k) H[k] = (4 * H2[k] - H1[k]) / 3;
    return JP_OK;
  }
};

double max_abs(const double* v, int n) {
  double m = 0;
  for (int i = 0; i < n; ++i) m = std::fmax(m, std::fabs(v[i]));
  return m;
}

// what the last mode search of this thread ended with (jp_mode_report): infinity norm of the gradient of the negative
// log-density at the returned point, and whether the iteration stopped on its own criterion rather than at the cap / a failed
// line search
thread_local double g_mode_grad = 0.0;
thread_local int g_mode_converged = 0, g_mode_iterations = 0;

int mode_generic(ModeProblem& P, double* x, double* H, double* fx_out, int iters) {
  const int d = P.d;
  const double h = 2e-3;
  std::vector<double> g(d), step(d), xn(d), gn(d), Hn((size_t)d * d), lam, V, cands((size_t)4 * d), fc(4);
  double fx;
  g_mode_converged = 0;
  g_mode_iterations = 0;
  int s = P.richardson(x, h, &fx, g.data(), H);
  if (s != JP_OK) return s;
  for (int it = 0; it < iters; ++it) {
    g_mode_iterations = it + 1;
    symmetric_eigen(H, d, lam, V);                       // ascending; V[i * d + k] = component k of eigenvector i
    double scale = 0;
    for (int i = 0; i < d; ++i) scale = std::fmax(scale, std::fabs(lam[i]));
    scale = std::fmax(scale, 1e-300);
    const double gnorm = max_abs(g.data(), d);
    if (lam[0] < -1e-8 * scale && gnorm < 1e-6 * scale) {
      static const double sgn[4] = {1, 1, -1, -1}, tt[4] = {1.0, 0.25, 1.0, 0.25};
      for (int c = 0; c < 4; ++c)
        for (int k = 0; k < d; ++k) cands[(size_t)c * d + k] = x[k] + tt[c] * sgn[c] * V[k];
      s = P.values(cands.data(), 4, fc.data());
      if (s != JP_OK) return s;
      int best = -1;
      for (int c = 0; c < 4; ++c)
        if (std::isfinite(fc[c]) && (best < 0 || fc[c] < fc[best])) best = c;
      if (best < 0 || !(fc[best] < fx)) break;
      for (int k = 0; k < d; ++k) x[k] = cands[(size_t)best * d + k];
      s = P.richardson(x, h, &fx, g.data(), H);
      if (s != JP_OK) return s;
      continue;
    }
    // step = -V (V' g / |lambda|)
    for (int k = 0; k < d; ++k) step[k] = 0;
    for (int i = 0; i < d; ++i) {
      double c = 0;
      for (int k = 0; k < d; ++k) c += V[(size_t)i * d + k] * g[k];
      c /= std::fmax(std::fabs(lam[i]), 1e-10 * scale);
      for (int k = 0; k < d; ++k) step[k] -= V[(size_t)i * d + k] * c;
    }
    double nrm = max_abs(step.data(), d);
    if (nrm > 10.0)
      for (int k = 0; k < d; ++k) step[k] *= 10.0 / nrm;
    if (lam[0] > 0 && max_abs(step.data(), d) < 1e-9 * (1 + max_abs(x, d))) {         // below the finite-difference resolution
      g_mode_converged = 1;
      break;
    }
    double t = 1.0, fn = 0;
    bool ok = false;
    while (t > 1e-8) {
      for (int k = 0; k < d; ++k) xn[k] = x[k] + t * step[k];
      s = P.richardson(xn.data(), h, &fn, gn.data(), Hn.data());
      if (s != JP_OK) return s;
      if (std::isfinite(fn) && fn <= fx + 1e-13 * (1 + std::fabs(fx))) {
        ok = true;
        break;
      }
      t *= 0.25;
    }
    if (!ok) break;
    for (int k = 0; k < d; ++k) { x[k] = xn[k]; g[k] = gn[k]; }
    std::memcpy(H, Hn.data(), sizeof(double) * d * d);
    fx = fn;
  }
  *fx_out = fx;
  g_mode_grad = max_abs(g.data(), d);
  // a line search that finds no further decrease at the resolution of the finite differences, with a gradient that is tiny next
  // to the objective, is the normal way this iteration ends
  if (!g_mode_converged && g_mode_grad <= 1e-5 * (1 + std::fabs(fx))) g_mode_converged = 1;
  return JP_OK;
}

// solve A y = b (A symmetric positive definite in practice; partial pivoting keeps it safe otherwise)
bool solve_dense(std::vector<double> A, std::vector<double>& b, int d) {
  for (int c = 0; c < d; ++c) {
    int piv = c;
    for (int r = c + 1; r < d; ++r)
      if (std::fabs(A[(size_t)c * d + r]) > std::fabs(A[(size_t)c * d + piv])) piv = r;
    if (A[(size_t)c * d + piv] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < d; ++k) std::swap(A[(size_t)k * d + c], A[(size_t)k * d + piv]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < d; ++r) {
      const double f = A[(size_t)c * d + r] / A[(size_t)c * d + c];
      if (f == 0.0) continue;
      for (int k = c; k < d; ++k) A[(size_t)k * d + r] -= f * A[(size_t)k * d + c];
      b[r] -= f * b[c];
    }
  }
  for (int r = d - 1; r >= 0; --r) {
    double v = b[r];
    for (int k = r + 1; k < d; ++k) v -= A[(size_t)k * d + r] * b[k];
    b[r] = v / A[(size_t)r * d + r];
  }
  return true;
}

int mode_glm(jp_ctx* ctx, const jp_data* data, jp_comm* comm, int d, double* x, double* H, double* neg_min, int* evals, int iters) {
  std::vector<double> g(d), g2(d), H2((size_t)d * d), step(d), xn(d);
  double lp = 0, lp2 = 0;
  int s = jp_glm_grad_hess_comm(ctx, data, comm, d, x, g.data(), H, &lp);
  ++*evals;
  if (s != JP_OK) return s;
  g_mode_converged = 0;
  g_mode_iterations = 0;
  for (int it = 0; it < iters; ++it) {
    g_mode_iterations = it + 1;
    step = g;
    if (!solve_dense(std::vector<double>(H, H + (size_t)d * d), step, d)) break;
    double t = 1.0;
    bool ok = false;
    while (t > 1e-10) {
      for (int k = 0; k < d; ++k) xn[k] = x[k] + t * step[k];
      s = jp_glm_grad_hess_comm(ctx, data, comm, d, xn.data(), g2.data(), H2.data(), &lp2);
      ++*evals;
      if (s != JP_OK) return s;
      if (std::isfinite(lp2) && lp2 >= lp - 1e-13 * std::fabs(lp)) {
        ok = true;
        break;
      }
      t *= 0.5;
    }
    if (!ok) break;
    double moved = 0;
    for (int k = 0; k < d; ++k) moved = std::fmax(moved, std::fabs(xn[k] - x[k]));
    for (int k = 0; k < d; ++k) { x[k] = xn[k]; g[k] = g2[k]; }
    std::memcpy(H, H2.data(), sizeof(double) * d * d);
    lp = lp2;
    if (moved < 1e-13 * (1 + max_abs(x, d))) {
      g_mode_converged = 1;
      break;
    }
  }
  g_mode_grad = max_abs(g.data(), d);
  // a Newton step that no longer moves the log-posterior (failed line search at the resolution of the sums) with a gradient that
  // is tiny next to the curvature is a converged fit as well
  if (!g_mode_converged && g_mode_grad <= 1e-6 * (1 + std::fabs(lp))) g_mode_converged = 1;
  *neg_min = -lp;
  return JP_OK;
}
}  // namespace

int jp_mode(jp_ctx* ctx, const jp_data* data, int d, const int* h_transform, int glm, double* h_x, double* h_H,
            double* neg_min, int* evals) {
  if (!ctx || !data || !h_transform || !h_x || !h_H || !neg_min || d < 1 || d > 64) {
    jp_set_error("jp_mode: bad argument");
    return JP_ERR_BAD_ARG;
  }
  int n_eval = 0, s;
  if (glm) {
    for (int k = 0; k < d; ++k)
      if (h_transform[k] != JP_T_REAL) {
        jp_set_error("jp_mode: the GLM Newton iteration needs unconstrained coefficients (coordinate %d is constrained)", k);
        return JP_ERR_BAD_ARG;
      }
    s = mode_glm(ctx, data, nullptr, d, h_x, h_H, neg_min, &n_eval, 60);
  } else {
    // small models: the whole search as one kernel launch (jp_mode_dev.cu); otherwise, or if that did not converge, the same
    // iteration driven from here with one batched evaluation per step
    int used = 0;
    s = jp_mode_dev_try(ctx, data, d, h_transform, h_x, h_H, neg_min, 100, &n_eval, &g_mode_iterations, &g_mode_converged,
                        &g_mode_grad, &used);
    if (s == JP_OK && !used) {
      ModeProblem P(ctx, data, d, h_transform);
      s = mode_generic(P, h_x, h_H, neg_min, 100);
      n_eval += P.evals;
    }
  }
  if (evals) *evals = n_eval;
  return s;
}

// jp_mode for a GLM whose observations are sharded over the ranks of `comm` (data = this rank's rows): the Newton iteration of
// jp_mode with every score / information evaluation summed over the ranks in rank order inside the library -- all ranks walk
// the same iterates and return bit-identical (x, H, neg_min).  Reference src/joint_posterior.jl:164-168.
int jp_mode_p2p(jp_ctx* ctx, const jp_data* data, jp_comm* comm, int d, const int* h_transform, double* h_x, double* h_H,
                double* neg_min, int* evals) {
  if (!ctx || !data || !comm || !h_transform || !h_x || !h_H || !neg_min || d < 1 || d > 64) {
    jp_set_error("jp_mode_p2p: bad argument");
    return JP_ERR_BAD_ARG;
  }
  for (int k = 0; k < d; ++k)
    if (h_transform[k] != JP_T_REAL) {
      jp_set_error("jp_mode_p2p: the GLM Newton iteration needs unconstrained coefficients (coordinate %d is constrained)", k);
      return JP_ERR_BAD_ARG;
    }
  int n_eval = 0;
  const int s = mode_glm(ctx, data, comm, d, h_x, h_H, neg_min, &n_eval, 60);
  if (evals) *evals = n_eval;
  return s;
}

// How the last jp_mode / jp_mode_p2p call of the calling thread ended: infinity norm of the gradient of the negative log-density
// at the returned point, iterations used, and whether the search stopped on its own criterion (1) or at the iteration cap / a
// failed line search (0) -- the reference's Optim call reports the same through its result object, and a grid centred on a
// point that is not the mode is a worse quadrature, not an error.
int jp_mode_report(double* grad_inf_norm, int* iterations, int* converged) {
  if (grad_inf_norm) *grad_inf_norm = g_mode_grad;
  if (iterations) *iterations = g_mode_iterations;
  if (converged) *converged = g_mode_converged;
  return JP_OK;
}

}  // extern "C"
