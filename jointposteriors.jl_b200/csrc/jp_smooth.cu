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
#include <cstdlib>
#include <cstring>

namespace {

constexpr int SM_THREADS = 256;
constexpr int SM_TILE = SM_THREADS - 2;      // interior points per tile: one halo residual on either side
constexpr int SM_SUMS = 12;                  // sum delta^2, sum delta_i delta_{i-1}, ten score sums
constexpr int SM_MAX_BLOCKS = 296;           // 2 per SM (a single parameter set)
constexpr int SM_MAX_BATCH = 10;             // parameter sets per launch: phi and its nine forward-difference neighbours
constexpr int SM_BATCH_BLOCKS = 148;         // blocks per set of a batched launch

struct SmoothArgs {
  double beta[10];
  double rho;
};
struct SmoothBatch {
  SmoothArgs set[SM_MAX_BATCH];
};

// One pass over the sorted nodes.  Thread t of a tile computes the residual of point (tile start + t - 1) into shared memory;
// the interior threads then own one point each with both neighbours at hand.  Points outside [0, M) contribute residual 0,
// which is exactly the boundary rule of mul_tstd_x! (:69, :75) and of the lagged product (:131-137).
__global__ void __launch_bounds__(SM_THREADS)
jp_smooth_sums_kernel(const double* __restrict__ V, const double* __restrict__ cw, long long M, const SmoothBatch batch,
                      double* __restrict__ bpart_all, unsigned int* __restrict__ counter_all, double* __restrict__ out_all) {
  // blockIdx.y = parameter set: the objective / score sums of up to SM_MAX_BATCH parameter vectors in ONE launch (a Newton
  // iteration needs the score at phi and at its nine forward-difference neighbours)
  const SmoothArgs& a = batch.set[blockIdx.y];
  double* bpart = bpart_all + (size_t)blockIdx.y * gridDim.x * SM_SUMS;
  unsigned int* counter = counter_all + blockIdx.y;
  double* out = out_all + (size_t)blockIdx.y * SM_SUMS;
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
    // warp w adds sums w, w + 8: lane l takes blocks l, l + 32, .. in ascending order, then the fixed shuffle tree
    for (int k = w; k < SM_SUMS; k += SM_THREADS / 32) {
      double s = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(bpart + (size_t)b * SM_SUMS + k);
      s = jp_warp_sum(s);
      if (lane == 0) out[k] = s;
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

// ntl_likelihood! and ntscore! (:81-111) at `nset` parameter vectors: ONE kernel launch, 12 doubles back per vector.  A
// non-finite objective (overflowing parameters during a line search) is reported as +inf with a zero gradient.
int smooth_eval_batch(SmoothProblem* P, int nset, const double (*phi)[9], double* f, double (*g9)[9], SmoothCoef* coef_out) {
  JP_REQUIRE(nset >= 1 && nset <= SM_MAX_BATCH, "smooth_eval_batch: %d parameter sets", nset);
  SmoothCoef c[SM_MAX_BATCH];
  bool finite[SM_MAX_BATCH];
  SmoothBatch batch;
  for (int q = 0; q < nset; ++q) {
    smooth_coefficients(phi[q], &c[q]);
    finite[q] = std::isfinite(c[q].sigma2) && c[q].sigma2 > 0 && std::isfinite(c[q].rho);
    for (int k = 0; k < 10; ++k) finite[q] = finite[q] && std::isfinite(c[q].beta[k]);
    for (int k = 0; k < 70; ++k) finite[q] = finite[q] && std::isfinite(c[q].J[k]);
    for (int k = 0; k < 10; ++k) batch.set[q].beta[k] = finite[q] ? c[q].beta[k] : 0.0;
    batch.set[q].rho = finite[q] ? c[q].rho : 0.0;
  }
  if (coef_out) *coef_out = c[0];
  jp_ctx* ctx = P->ctx;
  const long long tiles = (P->M + SM_TILE - 1) / SM_TILE;
  const unsigned blocks = (unsigned)std::max(1LL, std::min((long long)(nset == 1 ? SM_MAX_BLOCKS : SM_BATCH_BLOCKS), tiles));
  static_assert(128 + (size_t)SM_MAX_BATCH * SM_MAX_BLOCKS * SM_SUMS <= JP_SCRATCH_DOUBLES, "scratch too small for the smooth-CDF partials");
  // zero-copy result: the last block writes the (at most 120) sums straight into the context's pinned buffer, which the device
  // addresses at its host address (unified addressing) -- launch + synchronise, no copy operation in between
  JP_CUDA(jp_pinned_acquire(ctx));
  double* d_out = ctx->h_pinned;                  // [nset][12]
  double* d_part = ctx->d_scratch + 128;          // [nset][blocks][12]
  jp_smooth_sums_kernel<<<dim3(blocks, nset), SM_THREADS, 0, ctx->stream>>>(P->d_V, P->d_cw, P->M, batch, d_part, ctx->d_counters, d_out);
  JP_CHECK_LAUNCH(ctx);
  JP_CUDA(cudaStreamSynchronize(ctx->stream));
  P->evaluations += nset;
  const double n = (double)P->M;
  for (int q = 0; q < nset; ++q) {
    if (!finite[q]) {
      f[q] = INFINITY;
      if (g9) std::memset(g9[q], 0, 9 * sizeof(double));
      continue;
    }
    const double* s = ctx->h_pinned + (size_t)q * SM_SUMS;
    const double* ph = phi[q];
    const SmoothCoef& cq = c[q];
    const double Q = (s[0] - 2 * cq.rho * s[1]) / cq.sigma2;                                  // :138-140
    double logdet, dlogdet;
    log_det_tstd(cq.rho, P->M, &logdet, &dlogdet);
    const double lj = 3 * (ph[0] + ph[1] + ph[4]) / 2 - nlogit_lj(ph[2]) - nlogit_lj(ph[5]) + ph[7] - nlogit_lj(ph[8]);   // :293-295
    f[q] = (Q - logdet - 2 * lj + ph[0] * ph[0]) / n + ph[7];                                // :108-111
    if (!std::isfinite(f[q])) f[q] = INFINITY;
    if (!g9) continue;
    double g[9] = {0};
    for (int j = 0; j < 7; ++j)
      for (int k = 0; k < 10; ++k) g[j] += cq.J[7 * k + j] * (2 * s[2 + k] / cq.sigma2);     // :91-95
    g[0] += 2 * ph[0];                                                                       // :100
    g[7] = -Q + n;                                                                           // :101
    const double er = std::exp(ph[8]);
    g[8] = (-2 * s[1] / cq.sigma2 - dlogdet) * (1 / (2 + er + 1 / er)) / 4;                  // :102-103
    g[0] -= 3.0;                                                                             // nlj_grad!, :296-305
    g[1] -= 3.0;
    g[2] -= 2 * dnlogit_lj(ph[2]);
    g[4] -= 3.0;
    g[5] -= 2 * dnlogit_lj(ph[5]);
    g[7] -= 2.0;
    g[8] -= 2 * dnlogit_lj(ph[8]);
    for (int j = 0; j < 9; ++j) {
      g9[q][j] = g[j] / n;                                                                   // :105
      if (!std::isfinite(g9[q][j])) f[q] = INFINITY;
    }
  }
  return JP_OK;
}
int smooth_eval(SmoothProblem* P, const double* phi, double* f, double* g9, SmoothCoef* coef_out) {
  double ph[1][9], g[1][9];
  std::memcpy(ph[0], phi, sizeof ph[0]);
  JP_TRY(smooth_eval_batch(P, 1, ph, f, g9 ? g : nullptr, coef_out));
  if (g9) std::memcpy(g9, g[0], sizeof g[0]);
  return JP_OK;
}

// Saddle-free Newton with a trust region on the nine parameters: the analytic score at phi and at its nine forward-difference
// neighbours comes from ONE batched launch, the 9 x 9 Hessian is their difference quotient (symmetrised), its eigenvalues are
// replaced by their absolute values (floored), and the step is clipped to the trust radius, which grows after full steps and
// shrinks after rejected ones.  The reference asks Optim for BFGS (:380); on this objective -- nearly flat along two
// directions of the nested cubics -- BFGS alone needs > 1000 iterations for Optim's g_tol of 1e-8.  jp_marginal_smooth runs
// a short BFGS phase from the starting point (robust far from the minimum) and finishes here: the minimiser is the same point.
void jacobi9(double (*A)[9], double* lam, double (*V)[9]) {      // A is destroyed; V[:, i] = eigenvector of lam[i]
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int i = 0; i < 9; ++i)
      for (int j = i + 1; j < 9; ++j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < 9; ++p)
      for (int q = p + 1; q < 9; ++q) {
        if (A[p][q] == 0.0) continue;
        const double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
        const double c = 1 / std::sqrt(t * t + 1), sn = t * c;
        for (int k = 0; k < 9; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - sn * akq;
          A[k][q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - sn * aqk;
          A[q][k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - sn * vkq;
          V[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 9; ++i) lam[i] = A[i][i];
}
int smooth_newton(SmoothProblem* P, double* x, int max_iter, double g_tol, double* f_out, double* g_out, int* iters, int* converged,
                  const double* f_start = nullptr, const double* g_start = nullptr) {
  const int n = 9;
  double f, g[n];
  if (f_start && g_start) {
    f = *f_start;
    std::memcpy(g, g_start, sizeof g);
  } else {
    JP_TRY(smooth_eval(P, x, &f, g, nullptr));
  }
  JP_REQUIRE(std::isfinite(f), "jp_marginal_smooth: the objective is not finite at the starting point");
  *converged = 0;
  int it = 0;
  double radius = 1.0;
  // score at the nine forward-difference neighbours of x (gq[j] at x + h[j] e_j).  A trial point is evaluated TOGETHER with
  // its own nine neighbours (one launch of ten parameter sets): when the step is accepted -- nearly always -- the next
  // iteration's Hessian needs no launch of its own, so an iteration is one round trip to the device instead of two.
  double fq[SM_MAX_BATCH], gq[SM_MAX_BATCH][9], h[n];
  bool have_neighbours = false;
  auto step_of = [](double xj) { return 1e-6 * std::max(1.0, std::fabs(xj)); };
  for (; it < max_iter; ++it) {
    double gmax = 0;
    for (int i = 0; i < n; ++i) gmax = std::max(gmax, std::fabs(g[i]));
    if (gmax <= g_tol) {
      *converged = 1;
      break;
    }
    double ph[SM_MAX_BATCH][9], H[n][n], lam[n], V[n][n];
    if (!have_neighbours) {
      for (int j = 0; j < n; ++j) {
        std::memcpy(ph[j], x, sizeof ph[j]);
        h[j] = step_of(x[j]);
        ph[j][j] += h[j];
      }
      JP_TRY(smooth_eval_batch(P, n, ph, fq, gq, nullptr));
      have_neighbours = true;
    }
    bool ok_h = true;
    for (int j = 0; j < n; ++j) ok_h = ok_h && std::isfinite(fq[j]);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) H[i][j] = ok_h ? (gq[j][i] - g[i]) / h[j] : (i == j ? 1.0 : 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) H[i][j] = H[j][i] = 0.5 * (H[i][j] + H[j][i]);
    jacobi9(H, lam, V);
    double scale = 0;
    for (int i = 0; i < n; ++i) scale = std::max(scale, std::fabs(lam[i]));
    scale = std::max(scale, 1e-300);
    double coef[n], decrement = 0;      // Newton step in the eigenbasis: coef[i] = -(v_i . g) / |lambda_i|
    for (int i = 0; i < n; ++i) {
      double c = 0;
      for (int k = 0; k < n; ++k) c += V[k][i] * g[k];
      const double l = std::max(std::fabs(lam[i]), 1e-10 * scale);
      coef[i] = -c / l;
      decrement += 0.5 * c * c / l;
    }
    // The curvatures of this objective span ten orders of magnitude (the nested cubics degenerate towards a -> 0, where the
    // shift of the inner cubic and the constant of the outer one become one parameter): the score's infinity norm never
    // reaches 1e-8 -- scipy's BFGS (3000 iterations) and trust-exact (500) stop at 0.2 and 0.02 on the reference's own test
    // posterior -- while the decrease a full Newton step can still buy is far below the objective's own rounding.  That
    // predicted decrease is the scale-free stationarity test used here (reported as converged = 2).
    if (decrement <= 1e-8 * std::max(1.0, std::fabs(f))) {
      *converged = 2;
      break;
    }
    bool stepped = false;
    for (int attempt = 0; attempt < 30; ++attempt) {
      // the trust radius clips every eigen-direction on its own: a nearly flat direction that asks for a huge move is cut
      // back without shortening the full Newton steps of the stiff ones
      double pm = 0, xn[n], slope = 0, p[n] = {0}, hn[n], fb[SM_MAX_BATCH], gb[SM_MAX_BATCH][9];
      bool clipped = false;
      for (int i = 0; i < n; ++i) {
        double ci = coef[i];
        pm = std::max(pm, std::fabs(ci));
        if (std::fabs(ci) > radius) {
          ci = ci > 0 ? radius : -radius;
          clipped = true;
        }
        for (int k = 0; k < n; ++k) p[k] += V[k][i] * ci;
      }
      const double sc = clipped ? 0.5 : 1.0;
      for (int k = 0; k < n; ++k) {
        xn[k] = x[k] + p[k];
        slope += p[k] * g[k];
      }
      std::memcpy(ph[0], xn, sizeof xn);      // set 0: the trial point; sets 1 .. 9: its neighbours
      for (int j = 0; j < n; ++j) {
        std::memcpy(ph[1 + j], xn, sizeof xn);
        hn[j] = step_of(xn[j]);
        ph[1 + j][j] += hn[j];
      }
      JP_TRY(smooth_eval_batch(P, n + 1, ph, fb, gb, nullptr));
      const double fn = fb[0];
      if (std::isfinite(fn) && fn <= f + 1e-4 * slope) {
        std::memcpy(x, xn, sizeof xn);
        std::memcpy(g, gb[0], sizeof gb[0]);
        f = fn;
        for (int j = 0; j < n; ++j) {
          fq[j] = fb[1 + j];
          std::memcpy(gq[j], gb[1 + j], sizeof gq[j]);
          h[j] = hn[j];
        }
        if (sc == 1.0) radius = std::min(4.0 * radius, 16.0);
        else radius = std::min(2.0 * radius, 16.0);
        stepped = true;
        break;
      }
      radius = 0.25 * std::min(radius, pm);
      if (radius < 1e-14) break;
    }
    if (!stepped) break;      // no decrease inside any trust radius: stationary to working precision
  }
  *f_out = f;
  std::memcpy(g_out, g, sizeof g);
  *iters = it;
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

// the fit itself on a design matrix / cumulative-weight column already on the device
static int smooth_fit_on(jp_posterior* post, const double* d_V, const double* d_cw, double mu, double sigma, const double* phi_init,
                         int max_iter, double g_tol, jp_smooth_cdf* out) {
  SmoothProblem P{post->ctx, d_V, d_cw, post->M};
  double x[9] = {0};      // MarginalBuffer.init lives in the absent LogDensities: a = c = m = 1, b = d = l = n = 0, sigma2 = 1, rho = 1/8
  if (phi_init) std::memcpy(x, phi_init, sizeof x);
  double f = 0, g[9];
  int iters = 0, conv = 0;
  static const bool use_bfgs = getenv("JP_SMOOTH_BFGS") != nullptr;      // the reference's optimiser (:380), for comparison
  const double tol = g_tol > 0 ? g_tol : 1e-8;
  int st;
  if (use_bfgs) {
    st = smooth_bfgs(&P, x, max_iter ? max_iter : 1000, tol, &f, g, &iters, &conv);
  } else {
    // phase 1: BFGS from the starting point until the score is small (robust far from the minimum, one launch per step);
    // phase 2: saddle-free Newton on the batched score (fast near it).  max_iter caps the sum of both.
    const int cap = max_iter ? max_iter : 300;
    int it1 = 0, it2 = 0;
    st = smooth_bfgs(&P, x, std::min(cap, 60), std::max(tol, 1e-4), &f, g, &it1, &conv);
    if (st == JP_OK && !(conv && tol >= 1e-4) && cap > it1) {
      conv = 0;
      st = smooth_newton(&P, x, cap - it1, tol, &f, g, &it2, &conv, &f, g);
    }
    iters = it1 + it2;
  }
  SmoothCoef c;
  if (st == JP_OK) st = smooth_eval(&P, x, &f, g, &c);      // update_bt! at the minimiser, :381
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

static int smooth_check_sigma(double mu, double sigma) {
  // sigma^2 = E[v^2] - mu^2 is only known to ~1e-16 E[v^2]: below that the marginal is constant to working precision
  if (!(sigma > 1e-7 * std::sqrt(sigma * sigma + mu * mu)) || !std::isfinite(sigma)) {
    jp_set_error("jp_marginal_smooth: the marginal has no positive variance (sigma = %g)", sigma);
    return JP_ERR_BAD_ARG;
  }
  return JP_OK;
}

int jp_marginal_smooth(jp_posterior* post, int k, const double* phi_init, int max_iter, double g_tol, jp_smooth_cdf* out) {
  JP_REQUIRE(post && out, "jp_marginal_smooth: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_REQUIRE(max_iter >= 0 && g_tol >= 0, "jp_marginal_smooth: negative iteration cap or tolerance");
  double* d_V = nullptr;
  double mu, sigma;
  JP_TRY(jp_marginal_design_device(post, k, &d_V, nullptr, &mu, &sigma));
  int st = smooth_check_sigma(mu, sigma);
  if (st == JP_OK) st = smooth_fit_on(post, d_V, post->d_cw + (size_t)k * post->M, mu, sigma, phi_init, max_iter, g_tol, out);
  jp_dfree(post->ctx, d_V);
  return st;
}

// The reference keeps one MarginalBuffer per marginal function in M.MarginalBuffers (get!(..., f), src/marginal_posterior.jl:10,71)
// so that a repeated marginal(jp, f, Normal) reuses its storage.  Device analogue: per posterior, up to JP_DESIGN_CACHE
// entries (key chosen by the host, one per function) of the sorted design matrix + cumulative weights; an entry made since the
// last fit is reused as is -- no sort, no Vandermonde pass.  k < 0: only look the key up (*cache_hit = 0 and nothing else
// happens on a miss); k >= 0: marginal k of the last jp_marginal_* call provides the values on a miss.
int jp_marginal_smooth_keyed(jp_posterior* post, int k, long long key, const double* phi_init, int max_iter, double g_tol,
                             jp_smooth_cdf* out, int* cache_hit) {
  JP_REQUIRE(post && out && cache_hit, "jp_marginal_smooth_keyed: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_REQUIRE(max_iter >= 0 && g_tol >= 0, "jp_marginal_smooth_keyed: negative iteration cap or tolerance");
  jp_ctx* ctx = post->ctx;
  *cache_hit = 0;
  JpDesignCache* hit = nullptr;
  for (auto& e : post->design_cache)
    if (e.key == key && e.gen == post->fit_gen) hit = &e;
  if (!hit) {
    if (k < 0) return JP_OK;
    double* d_V = nullptr;
    double mu, sigma;
    JP_TRY(jp_marginal_design_device(post, k, &d_V, nullptr, &mu, &sigma));
    double* d_cw = nullptr;
    cudaError_t e = jp_dmalloc(ctx, &d_cw, (size_t)post->M * 8);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_cw, post->d_cw + (size_t)k * post->M, (size_t)post->M * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) {
      jp_dfree(ctx, d_V);
      jp_dfree(ctx, d_cw);
      JP_CUDA(e);
    }
    // stale entries (made before the last fit) go first, then the least recently used
    size_t victim = post->design_cache.size();
    for (size_t i = 0; i < post->design_cache.size(); ++i)
      if (post->design_cache[i].gen != post->fit_gen || post->design_cache[i].key == key) victim = i;
    if (victim == post->design_cache.size() && post->design_cache.size() >= JP_DESIGN_CACHE) {
      victim = 0;
      for (size_t i = 1; i < post->design_cache.size(); ++i)
        if (post->design_cache[i].stamp < post->design_cache[victim].stamp) victim = i;
    }
    if (victim < post->design_cache.size()) {
      jp_dfree(ctx, post->design_cache[victim].d_V);
      jp_dfree(ctx, post->design_cache[victim].d_cw);
      post->design_cache.erase(post->design_cache.begin() + (long)victim);
    }
    post->design_cache.push_back(JpDesignCache{key, post->fit_gen, d_V, d_cw, mu, sigma, 0});
    hit = &post->design_cache.back();
  } else {
    *cache_hit = 1;
  }
  hit->stamp = ++post->cache_clock;
  JP_TRY(smooth_check_sigma(hit->mu, hit->sigma));
  return smooth_fit_on(post, hit->d_V, hit->d_cw, hit->mu, hit->sigma, phi_init, max_iter, g_tol, out);
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
