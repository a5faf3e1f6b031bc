// jp_smooth.cu -- the smooth CDF of marginal(jp, f, Normal): NestedPolyGLM, reference src/interp.jl:33-446 and
// src/marginal_posterior.jl:124-129.
//
// Model (update_ab!, :203-235): F(x) = Phi(P(Q(z))), z = (x - mu) / sigma, Q(z) = z^3 + l z^2 + m z + n, P(y) = a y^3 + b y^2 + c y + d,
// both cubics monotone by construction (b = sqrt(3ac) tanh(phi_3 / 2), l = sqrt(3m) tanh(phi_6 / 2)); beta = the ten coefficients
// of the composition.  The nine unconstrained parameters phi minimise
//     f(phi) = (Q - logdet T(rho) - 2 lj(phi) + phi_1^2) / n + phi_8,   Q = delta' T(rho) delta / sigma2              (:108-111)
// with delta_i = Phi(V_i . beta) - w_i over ALL M sorted nodes (w = cumulative weights), T = tridiag(1, -rho)   (:56-64, :128-143).
//
// Split of the work: everything that touches the M nodes -- the design matrix (jp_marginal_design_device), the residuals, the two
// quadratic sums and the ten score sums  sum_i V_i pdf_i (delta_i - rho (delta_{i-1} + delta_{i+1}))  (mul_tstd_x!, :66-76,
// ntscore!, :86-90) -- is ONE kernel launch per evaluation on data that never leaves the device; the host owns the 9-parameter
// maps, their chain rule and the BFGS iteration (the reference runs Optim's BFGS with a backtracking line search, :380).
// Where the reference multiplies by its tabulated Jacobian alpha (:236-289) and takes d logdet / d rho by a complex step (:102),
// the same derivatives are applied analytically; logdet is accumulated as a sum of logs of pivot ratios, which is the same
// number at the reference's sizes and does not underflow at M = 1e5 (the reference's running determinant does).
#include "jp_common.cuh"

#include <cmath>
#include <cstring>

namespace {

constexpr int SM_THREADS = 256;
constexpr int SM_TILE = SM_THREADS - 2;      // interior points per tile: one halo residual on either side
constexpr int SM_SUMS = 12;                  // sum delta^2, sum delta_i delta_{i-1}, ten score sums
constexpr int SM_MAX_BLOCKS = 296;           // 2 per SM; 296 x 12 partials fit ctx->d_bpart

struct SmoothArgs {
  double beta[10];
  double rho;
};

// One pass over the sorted nodes.  Thread t of a tile computes the residual of point (tile start + t - 1) into shared memory;
// the interior threads then own one point each with both neighbours at hand.  Points outside [0, M) contribute residual 0,
// which is exactly the boundary rule of mul_tstd_x! (:69, :75) and of the lagged product (:131-137).
__global__ void __launch_bounds__(SM_THREADS)
jp_smooth_sums_kernel(const double* __restrict__ V, const double* __restrict__ cw, long long M, SmoothArgs a,
                      double* __restrict__ bpart, unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double s_delta[SM_THREADS];
  __shared__ double s_red[SM_SUMS][8];
  const int t = threadIdx.x;
  double acc[SM_SUMS];
#pragma unroll
  for (int k = 0; k < SM_SUMS; ++k) acc[k] = 0.0;
  const long long n_tiles = (M + SM_TILE - 1) / SM_TILE;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long j = tile * SM_TILE + t - 1;
    double v[10], delta = 0.0, pdf = 0.0;
    const bool in = j >= 0 && j < M;
    if (in) {
      const double2* row = reinterpret_cast<const double2*>(V + (size_t)j * 10);      // 80-byte rows: 16-byte aligned
      double eta = 0.0;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double2 p = row[k];
        v[2 * k] = p.x;
        v[2 * k + 1] = p.y;
      }
#pragma unroll
      for (int k = 0; k < 10; ++k) eta += v[k] * a.beta[k];                         // At_mul_B!(V beta), :60
      delta = (1.0 + erf(eta * 0.70710678118654752440)) / 2.0 - cw[j];              // unconstrained_cdf, :49-51, :61
      pdf = exp(-eta * eta / 2.0) * 0.39894228040143267794;                         // unconstrained_pdf, :52-54
    }
    __syncthreads();                                                                // previous tile's readers are done
    s_delta[t] = delta;
    __syncthreads();
    if (in && t >= 1 && t <= SM_TILE) {
      const double prev = s_delta[t - 1], next = s_delta[t + 1];
      acc[0] += delta * delta;                                                      // :133
      acc[1] += delta * prev;                                                       // :134
      const double zz = (delta - a.rho * (prev + next)) * pdf;                      // :66-76, :89
#pragma unroll
      for (int k = 0; k < 10; ++k) acc[2 + k] += v[k] * zz;                         // A_mul_B!(grad beta, V, .), :90
    }
  }
  // block reduction in a fixed order, then the last block to arrive combines the partials in block order (deterministic)
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int k = 0; k < SM_SUMS; ++k) {
    const double s = jp_warp_sum(acc[k]);
    if (lane == 0) s_red[k][w] = s;
  }
  __syncthreads();
  if (t < SM_SUMS) {
    double s = 0.0;
    for (int i = 0; i < SM_THREADS / 32; ++i) s += s_red[t][i];
    bpart[(size_t)blockIdx.x * SM_SUMS + t] = s;
  }
  __syncthreads();
  if (jp_last_block(counter, gridDim.x)) {
    if (t < SM_SUMS) {
      double s = 0.0;
      for (unsigned int b = 0; b < gridDim.x; ++b) s += bpart[(size_t)b * SM_SUMS + t];
      out[t] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------- host: the 9-parameter maps
struct SmoothCoef {
  double beta[10], theta[7], J[70], sigma2, rho;
};

// update_ab! (:203-235) and update_bt! (:161-200): beta as the composition P(Q(z)), and d beta / d phi_1..7 by the chain rule.
void smooth_coefficients(const double* phi, SmoothCoef* c) {
  const double a = std::exp(phi[0]), cc = std::exp(phi[1]);
  const double e3 = std::exp(phi[2]), sb = std::sqrt(3 * a * cc), b = sb * (e3 - 1) / (e3 + 1);
  const double d = phi[3], m = std::exp(phi[4]);
  const double e6 = std::exp(phi[5]), sl = std::sqrt(3 * m), l = sl * (e6 - 1) / (e6 + 1);
  const double n = phi[6];
  c->sigma2 = std::exp(phi[7]);
  c->rho = 1.0 / (4.0 * (1.0 + std::exp(-phi[8])));
  const double q1[4] = {n, m, l, 1.0};
  double q2[7] = {0}, q3[10] = {0}, g[7];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) q2[i + j] += q1[i] * q1[j];
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 4; ++j) q3[i + j] += q2[i] * q1[j];
  for (int k = 0; k < 10; ++k) c->beta[k] = a * q3[k] + (k < 7 ? b * q2[k] : 0.0) + (k < 4 ? cc * q1[k] : 0.0);
  c->beta[0] += d;
  const double th[7] = {a, cc, b, d, m, l, n};
  std::memcpy(c->theta, th, sizeof th);
  for (int k = 0; k < 7; ++k) g[k] = 3 * a * q2[k] + (k < 4 ? 2 * b * q1[k] : 0.0);     // P'(Q(z)) as a polynomial in z
  g[0] += cc;
  const double db3 = sb * 2 * e3 / ((e3 + 1) * (e3 + 1)), dl6 = sl * 2 * e6 / ((e6 + 1) * (e6 + 1));
  for (int k = 0; k < 10; ++k) {
    const double da = q3[k], db = k < 7 ? q2[k] : 0.0, dc = k < 4 ? q1[k] : 0.0;
    const double dn = k < 7 ? g[k] : 0.0, dm = (k >= 1 && k < 8) ? g[k - 1] : 0.0, dl = (k >= 2 && k < 9) ? g[k - 2] : 0.0;
    double* r = c->J + 7 * k;
    r[0] = a * da + 0.5 * b * db;
    r[1] = cc * dc + 0.5 * b * db;
    r[2] = db3 * db;
    r[3] = k == 0 ? 1.0 : 0.0;
    r[4] = m * dm + 0.5 * l * dl;
    r[5] = dl6 * dl;
    r[6] = dn;
  }
}

double nlogit_lj(double x) {      // :318-321
  const double e = std::exp(x);
  return std::log(2 + e + 1 / e);
}
double dnlogit_lj(double x) {     // :326-330
  const double e = std::exp(x), e2 = e * e;
  return (1 - e2) / (2 * e + e2 + 1);
}

// log det tridiag(1, -rho) of order n (log_det_tstd, :147-156) and its derivative in rho, as sums over the pivot ratios
// r_k = D_k / D_{k-1} = 1 - rho^2 / r_{k-1}.  The ratios converge geometrically (rho <= 1/4); once they are stationary to the
// last bit the remaining terms are added in closed form.
void log_det_tstd(double rho, long long n, double* logdet, double* dlogdet) {
  const double r2 = rho * rho;
  double r = 1.0, dr = 0.0, s = 0.0, ds = 0.0;
  for (long long k = 2; k <= n; ++k) {
    const double rn = 1.0 - r2 / r, drn = -2.0 * rho / r + r2 * dr / (r * r);
    const bool stationary = rn == r && drn == dr;
    r = rn;
    dr = drn;
    if (stationary) {
      s += (double)(n - k + 1) * std::log(r);
      ds += (double)(n - k + 1) * dr / r;
      break;
    }
    s += std::log(r);
    ds += dr / r;
  }
  *logdet = s;
  *dlogdet = ds;
}

struct SmoothProblem {
  jp_ctx* ctx;
  const double* d_V;
  const double* d_cw;
  long long M;
  long long evaluations = 0;
};

// ntl_likelihood! and ntscore! (:81-111) at phi: one kernel launch, 12 doubles back.  A non-finite objective (overflowing
// parameters during a line search) is reported as +inf with a zero gradient.
int smooth_eval(SmoothProblem* P, const double* phi, double* f, double* g9, SmoothCoef* coef_out) {
  SmoothCoef c;
  smooth_coefficients(phi, &c);
  if (coef_out) *coef_out = c;
  bool finite = std::isfinite(c.sigma2) && c.sigma2 > 0 && std::isfinite(c.rho);
  for (int k = 0; k < 10; ++k) finite = finite && std::isfinite(c.beta[k]);
  for (int k = 0; k < 70; ++k) finite = finite && std::isfinite(c.J[k]);
  if (!finite) {
    *f = INFINITY;
    if (g9) std::memset(g9, 0, 9 * sizeof(double));
    return JP_OK;
  }
  SmoothArgs a;
  std::memcpy(a.beta, c.beta, sizeof a.beta);
  a.rho = c.rho;
  jp_ctx* ctx = P->ctx;
  const long long tiles = (P->M + SM_TILE - 1) / SM_TILE;
  const unsigned blocks = (unsigned)std::max(1LL, std::min((long long)SM_MAX_BLOCKS, tiles));
  double* d_out = ctx->d_scratch;
  jp_smooth_sums_kernel<<<blocks, SM_THREADS, 0, ctx->stream>>>(P->d_V, P->d_cw, P->M, a, ctx->d_bpart, ctx->d_counters, d_out);
  JP_CHECK_LAUNCH(ctx);
  JP_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_out, SM_SUMS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  JP_CUDA(cudaStreamSynchronize(ctx->stream));
  P->evaluations++;
  const double* s = ctx->h_pinned;
  const double n = (double)P->M;
  const double Q = (s[0] - 2 * c.rho * s[1]) / c.sigma2;                                  // :138-140
  double logdet, dlogdet;
  log_det_tstd(c.rho, P->M, &logdet, &dlogdet);
  const double lj = 3 * (phi[0] + phi[1] + phi[4]) / 2 - nlogit_lj(phi[2]) - nlogit_lj(phi[5]) + phi[7] - nlogit_lj(phi[8]);   // :293-295
  *f = (Q - logdet - 2 * lj + phi[0] * phi[0]) / n + phi[7];                              // :108-111
  if (!std::isfinite(*f)) *f = INFINITY;
  if (!g9) return JP_OK;
  double g[9] = {0};
  for (int j = 0; j < 7; ++j)
    for (int k = 0; k < 10; ++k) g[j] += c.J[7 * k + j] * (2 * s[2 + k] / c.sigma2);     // :91-95
  g[0] += 2 * phi[0];                                                                     // :100
  g[7] = -Q + n;                                                                          // :101
  const double er = std::exp(phi[8]);
  g[8] = (-2 * s[1] / c.sigma2 - dlogdet) * (1 / (2 + er + 1 / er)) / 4;                  // :102-103
  g[0] -= 3.0;                                                                            // nlj_grad!, :296-305
  g[1] -= 3.0;
  g[2] -= 2 * dnlogit_lj(phi[2]);
  g[4] -= 3.0;
  g[5] -= 2 * dnlogit_lj(phi[5]);
  g[7] -= 2.0;
  g[8] -= 2 * dnlogit_lj(phi[8]);
  for (int j = 0; j < 9; ++j) {
    g9[j] = g[j] / n;                                                                     // :105
    if (!std::isfinite(g9[j])) *f = INFINITY;
  }
  return JP_OK;
}

// BFGS on the inverse Hessian with a backtracking (Armijo, quadratic interpolation clipped to [0.1, 0.5]) line search: the
// iteration the reference asks of Optim (`BFGS(; linesearch = BackTracking())`, :380; g_tol 1e-8 on the infinity norm and
// 1000 iterations are Optim's defaults).
int smooth_bfgs(SmoothProblem* P, double* x, int max_iter, double g_tol, double* f_out, double* g_out, int* iters, int* converged) {
  const int n = 9;
  double f, g[n], H[n][n], p[n], xn[n], gn[n], fn;
  JP_TRY(smooth_eval(P, x, &f, g, nullptr));
  JP_REQUIRE(std::isfinite(f), "jp_marginal_smooth: the objective is not finite at the starting point");
  auto reset = [&]() {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) H[i][j] = i == j ? 1.0 : 0.0;
  };
  reset();
  bool fresh = true;
  int it = 0;
  *converged = 0;
  for (; it < max_iter; ++it) {
    double gmax = 0;
    for (int i = 0; i < n; ++i) gmax = std::max(gmax, std::fabs(g[i]));
    if (gmax <= g_tol) {
      *converged = 1;
      break;
    }
    double slope = 0;
    for (int i = 0; i < n; ++i) {
      p[i] = 0;
      for (int j = 0; j < n; ++j) p[i] -= H[i][j] * g[j];
      slope += p[i] * g[i];
    }
    if (!(slope < 0)) {      // not a descent direction: restart from steepest descent
      reset();
      fresh = true;
      slope = 0;
      for (int i = 0; i < n; ++i) p[i] = -g[i], slope -= g[i] * g[i];
    }
    double alpha = 1.0;
    bool ok = false;
    for (int ls = 0; ls < 60; ++ls) {
      for (int i = 0; i < n; ++i) xn[i] = x[i] + alpha * p[i];
      JP_TRY(smooth_eval(P, xn, &fn, gn, nullptr));
      if (std::isfinite(fn) && fn <= f + 1e-4 * alpha * slope) {
        ok = true;
        break;
      }
      double shrink = 0.5;
      if (std::isfinite(fn)) {
        const double aq = -slope * alpha * alpha / (2 * (fn - f - slope * alpha));      // minimiser of the interpolating parabola
        if (std::isfinite(aq)) shrink = std::min(0.5, std::max(0.1, aq / alpha));
      }
      alpha *= shrink;
    }
    if (!ok) {
      if (fresh) break;      // no decrease along steepest descent: a stationary point to working precision
      reset();
      fresh = true;
      continue;
    }
    double s[n], y[n], sy = 0, ss = 0, yy = 0;
    for (int i = 0; i < n; ++i) {
      s[i] = xn[i] - x[i];
      y[i] = gn[i] - g[i];
      sy += s[i] * y[i];
      ss += s[i] * s[i];
      yy += y[i] * y[i];
    }
    const double df = f - fn;
    std::memcpy(x, xn, sizeof xn);
    std::memcpy(g, gn, sizeof gn);
    f = fn;
    if (sy > 1e-10 * std::sqrt(ss * yy)) {
      if (fresh) {           // scale the first update (Nocedal & Wright 6.20)
        const double sc = sy / yy;
        for (int i = 0; i < n; ++i) H[i][i] = sc;
      }
      double Hy[n], yHy = 0;
      for (int i = 0; i < n; ++i) {
        Hy[i] = 0;
        for (int j = 0; j < n; ++j) Hy[i] += H[i][j] * y[j];
      }
      for (int i = 0; i < n; ++i) yHy += y[i] * Hy[i];
      const double r = 1.0 / sy;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H[i][j] += -r * (Hy[i] * s[j] + s[i] * Hy[j]) + r * r * (sy + yHy) * s[i] * s[j];
      fresh = false;
    }
    if (df <= 1e-15 * std::max(1.0, std::fabs(f)) && std::sqrt(ss) <= 1e-12) break;      // no measurable progress
  }
  *f_out = f;
  std::memcpy(g_out, g, sizeof g);
  *iters = it;
  return JP_OK;
}

// Phi^-1 by Halley iterations on erfc from a logistic starting value (there is no erfinv in the C++ library; the reference's
// quantile uses sqrt(2) erfinv(2p - 1), :369).
double norm_quantile(double p) {
  if (!(p > 0.0)) return p == 0.0 ? -INFINITY : NAN;
  if (!(p < 1.0)) return p == 1.0 ? INFINITY : NAN;
  const bool upper = p > 0.5;
  const double q = upper ? 1.0 - p : p;                     // solve in the lower tail, mirror
  double z = -std::sqrt(-2.0 * std::log(q)) * 0.8 - 0.2;    // crude start, left of the root for small q
  if (q > 0.1) z = (std::log(q / (1.0 - q))) * 0.6;
  for (int i = 0; i < 100; ++i) {
    const double F = 0.5 * std::erfc(-z * 0.70710678118654752440) - q;
    const double pdf = std::exp(-z * z / 2.0) * 0.39894228040143267794;
    const double u = F / pdf, dz = u / (1.0 + 0.5 * z * u);
    z -= dz;
    if (std::fabs(dz) <= 4e-16 * std::max(1.0, std::fabs(z))) break;
  }
  return upper ? -z : z;
}

// one_cubic_root(a, c, b, d), :388-394 (argument order as there): the single real root of a y^3 + b y^2 + c y + d when b^2 < 3ac
double one_cubic_root(double a, double c, double b, double d) {
  const double D0 = b * b - 3 * a * c;
  const double D1 = 2 * b * b * b - 9 * a * b * c + 27 * a * a * d;
  const double C = std::cbrt((D1 + std::sqrt(D1 * D1 - 4 * D0 * D0 * D0)) / 2);
  return -(b + C + D0 / C) / (3 * a);
}

}  // namespace

extern "C" {

int jp_smooth_objective(jp_posterior* post, int k, const double* phi, double* f, double* grad9, double* beta10, double* theta7) {
  JP_REQUIRE(post && phi && f, "jp_smooth_objective: null argument");
  JP_ENTER_CTX(post->ctx);
  double* d_V = nullptr;
  JP_TRY(jp_marginal_design_device(post, k, &d_V, nullptr, nullptr, nullptr));
  SmoothProblem P{post->ctx, d_V, post->d_cw + (size_t)k * post->M, post->M};
  SmoothCoef c;
  const int st = smooth_eval(&P, phi, f, grad9, &c);
  jp_dfree(post->ctx, d_V);
  if (beta10) std::memcpy(beta10, c.beta, sizeof c.beta);
  if (theta7) std::memcpy(theta7, c.theta, sizeof c.theta);
  return st;
}

int jp_marginal_smooth(jp_posterior* post, int k, const double* phi_init, int max_iter, double g_tol, jp_smooth_cdf* out) {
  JP_REQUIRE(post && out, "jp_marginal_smooth: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_REQUIRE(max_iter >= 0 && g_tol >= 0, "jp_marginal_smooth: negative iteration cap or tolerance");
  double* d_V = nullptr;
  double mu, sigma;
  JP_TRY(jp_marginal_design_device(post, k, &d_V, nullptr, &mu, &sigma));
  // sigma^2 = E[v^2] - mu^2 is only known to ~1e-16 E[v^2]: below that the marginal is constant to working precision
  if (!(sigma > 1e-7 * std::sqrt(sigma * sigma + mu * mu)) || !std::isfinite(sigma)) {
    jp_dfree(post->ctx, d_V);
    jp_set_error("jp_marginal_smooth: the marginal has no positive variance (sigma = %g)", sigma);
    return JP_ERR_BAD_ARG;
  }
  SmoothProblem P{post->ctx, d_V, post->d_cw + (size_t)k * post->M, post->M};
  double x[9] = {0};      // MarginalBuffer.init lives in the absent LogDensities: a = c = m = 1, b = d = l = n = 0, sigma2 = 1, rho = 1/8
  if (phi_init) std::memcpy(x, phi_init, sizeof x);
  double f = 0, g[9];
  int iters = 0, conv = 0;
  int st = smooth_bfgs(&P, x, max_iter ? max_iter : 1000, g_tol > 0 ? g_tol : 1e-8, &f, g, &iters, &conv);
  SmoothCoef c;
  if (st == JP_OK) st = smooth_eval(&P, x, &f, g, &c);      // update_bt! at the minimiser, :381
  jp_dfree(post->ctx, d_V);
  JP_TRY(st);
  std::memcpy(out->beta, c.beta, sizeof c.beta);
  std::memcpy(out->theta, c.theta, sizeof c.theta);
  std::memcpy(out->phi, x, sizeof x);
  out->mu = mu;
  out->sigma = sigma;
  out->objective = f;
  out->grad_inf_norm = 0;
  for (int i = 0; i < 9; ++i) out->grad_inf_norm = std::max(out->grad_inf_norm, std::fabs(g[i]));
  out->iterations = iters;
  out->evaluations = (int)P.evaluations;
  out->converged = conv;
  return JP_OK;
}

// polyexpreval, :338-346, and the cdf of :365-367
double jp_smooth_cdf_eval(const jp_smooth_cdf* s, double x) {
  const double z = (x - s->mu) / s->sigma;
  double zi = z, out = s->beta[0] + z * s->beta[1];
  for (int i = 2; i < 10; ++i) {
    zi *= z;
    out += zi * s->beta[i];
  }
  return (1.0 + std::erf(out * 0.70710678118654752440)) / 2.0;
}

// d-polyexpreval, :348-361, and the pdf of :371-374
double jp_smooth_pdf_eval(const jp_smooth_cdf* s, double x) {
  const double z = (x - s->mu) / s->sigma;
  double fx = s->beta[0], dfx = 0.0, zi = 1.0;
  for (int i = 1; i < 10; ++i) {
    dfx += zi * i * s->beta[i];
    zi *= z;
    fx += zi * s->beta[i];
  }
  return std::exp(-fx * fx / 2.0) * 0.39894228040143267794 * dfx / s->sigma;
}

// quantile, :368-370, with nested_root, :402-405
double jp_smooth_quantile_eval(const jp_smooth_cdf* s, double p) {
  const double* t = s->theta;
  const double zq = norm_quantile(p);
  if (std::isinf(zq)) return zq;      // p = 0 / 1: the composition is increasing (the reference's formula gives -Inf / NaN)
  const double r1 = one_cubic_root(t[0], t[1], t[2], t[3] - zq);
  return one_cubic_root(1.0, t[4], t[5], t[6] - r1) * s->sigma + s->mu;
}

}  // extern "C"
