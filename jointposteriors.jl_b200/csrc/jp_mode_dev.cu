// jp_mode_dev.cu -- the mode search of small models as ONE kernel launch.
//
// mode(M, data) (reference src/joint_posterior.jl:164-168: optBFGS! + ForwardDiff Hessian) is, for every family that is not
// a GLM with analytic derivatives, a saddle-free Newton iteration on Richardson finite differences of the plugin log-density
// (jp_hostlinalg.cpp, mode_generic).  Driven from the host, each of its 10-20 iterations costs a launch and a
// synchronisation (~45 us on the bench boxes) -- at the reference's own model sizes (README Example 1: 8 data rows, 3
// parameters) that was 80 % of the whole fit.  Here the SAME iteration runs inside one launch of ONE thread-block cluster
// (8 CTAs): the finite-difference stencil (1 + 4 d^2 points) x the observations are spread over the cluster's threads, every
// CTA receives all stencil values through distributed shared memory and repeats the (deterministic) d x d linear algebra
// and the control flow for itself -- so the CTAs stay in lock step with one cluster barrier per stencil evaluation and no
// broadcast of decisions -- and the host sees one launch and one synchronisation for the whole search.
//
// Gate (jp_mode_dev_try): d <= 16, the records fit one shared-memory tile, stencil x observations small enough for one
// block.  The result is taken only if the search converged with finite values; anything else falls back to the host-driven
// iteration from the same start (JP_MODE_HOST=1 forces that path, for A/B tests).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cooperative_groups.h>
#include "jp_common.cuh"
#include "jp_family.cuh"
#include "jp_construct.cuh"

int jp_check_transform_codes(const char* who, const int* code, int d);      // jp_fit.cu

namespace cg = cooperative_groups;

#define JP_MD_THREADS 256
#define JP_MD_CTAS 8                  // portable cluster size
#define JP_MD_DMAX 16
#define JP_MD_OBS_DOUBLES 4096

struct JpModeDevParams {
  int d, ncols, iters, K, n, S;       // n = stencil points per step size, K = 1 + 2 n, S = observation slices per point
  int split;                          // 1: independent coordinate transforms only, shared out over a point's S lanes
  long long N;
  double h;
  const double* obs;                  // N x ncols (device)
  const double* x0;                   // d (zero-copy pinned)
  const int* tcode;                   // d (zero-copy pinned)
  double* out;                        // x d | H d*d | fx | grad | iterations | converged | evals | finite | cycle counts  (zero-copy pinned)
  double hyper[JP_MAX_HYPER];
};

namespace {

struct MdShared {
  double x[JP_MD_DMAX], xn[JP_MD_DMAX], g[JP_MD_DMAX], gn[JP_MD_DMAX], step[JP_MD_DMAX], lam[JP_MD_DMAX];
  double H[JP_MD_DMAX * JP_MD_DMAX], Hn[JP_MD_DMAX * JP_MD_DMAX], A[JP_MD_DMAX * JP_MD_DMAX], V[JP_MD_DMAX * JP_MD_DMAX];
  double cand[4 * JP_MD_DMAX], fc[4];
  double fx, fn, t;
  int code[JP_MD_DMAX];
  unsigned char pi[JP_MD_DMAX * (JP_MD_DMAX - 1) / 2], pj[JP_MD_DMAX * (JP_MD_DMAX - 1) / 2];
  double rot_a[JP_MD_DMAX], rot_b[JP_MD_DMAX];      // one round of disjoint Jacobi rotations: new col_j = a_j col_j + b_j col_partner(j)
  int rot_p[JP_MD_DMAX];
  int action, evals, n_eigen, n_values, imin, split_construct;
  long long cyc_values, cyc_linalg, cyc_derivs, cyc_chol, cyc_eigen;
  int n_sweeps;
  long long cyc_r1, cyc_r2, cyc_r3, cyc_chk;
};

enum { MD_STEP = 0, MD_ESCAPE = 1, MD_DONE = 2 };

// offset (in units of the step size) of stencil point o in [0, n): +-e_i first, then (+-e_i +- e_j) for i < j
__device__ __forceinline__ void md_offset(const MdShared& s, int d, int o, int& i, double& si, int& j, double& sj) {
  if (o < 2 * d) {
    i = o >> 1; si = (o & 1) ? -1.0 : 1.0; j = -1; sj = 0.0;
  } else {
    const int q = o - 2 * d, pr = q >> 2, k = q & 3;
    i = s.pi[pr]; j = s.pj[pr];
    si = (k < 2) ? 1.0 : -1.0;
    sj = (k & 1) ? -1.0 : 1.0;
  }
}

// NEGATIVE log-density (with log-Jacobian) at `count` points: kind 0 = the (h, h/2) stencils around xc (point 0 = xc itself),
// kind 1 = the rows of s.cand.  Item = (point, observation slice): the S slices of a point are adjacent lanes of one warp;
// the items are dealt to the cluster's warps 32 at a time, round-robin over the CTAs.  The value of a point is stored into the
// value buffer of EVERY CTA (distributed shared memory); two buffers alternate from call to call, so a fast CTA writing the
// next evaluation's values never touches what a slower one is still reading.  Returns this call's buffer.
template <class F, int DPAD>
__device__ __noinline__ double* md_values(const JpModeDevParams& P, MdShared& s, const double* s_obs, double* s_val2, int kind, const double* xc,
                             int count) {
  cg::cluster_group cluster = cg::this_cluster();
  const int S = P.S, W = count * S, d = P.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cta = (int)cluster.block_rank(), n_cta = (int)cluster.num_blocks();
  double* s_val = s_val2 + (s.n_values & 1) * (P.K + 1);
  for (int chunk = cta + n_cta * warp; 32 * chunk < W; chunk += n_cta * (JP_MD_THREADS / 32)) {
    const int item = 32 * chunk + lane, slice = item & (S - 1);
    const bool live = item < W;
    const int point = live ? item / S : 0;      // lanes past the end shadow point 0: they take part in the shuffles only
    double th[DPAD];
#pragma unroll
    for (int k = 0; k < DPAD; ++k) th[k] = 0.0;
    if (kind == 0) {
      int i = -1, j = -1;
      double si = 0, sj = 0, hh = 0;
      if (point > 0) {
        const int o = point - 1;
        hh = (o < P.n) ? P.h : 0.5 * P.h;
        md_offset(s, d, o < P.n ? o : o - P.n, i, si, j, sj);
      }
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < d) th[k] = xc[k] + (k == i ? hh * si : (k == j ? hh * sj : 0.0));
    } else {
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < d) th[k] = s.cand[point * JP_MD_DMAX + k];
    }
    double lj;
    if (s.split_construct) {
      // independent coordinate transforms (real / positive / probability only): the S lanes of a point take the coordinates in
      // turn and exchange the results -- the critical path is one transform instead of d of them
      double ljp = 0.0;
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < d && (k & (S - 1)) == slice) th[k] = jp_transform(s.code[k], th[k], ljp);
      const int first = lane & ~(S - 1);
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < d) th[k] = __shfl_sync(0xffffffffu, th[k], first + (k & (S - 1)));
      for (int o = S >> 1; o > 0; o >>= 1) ljp += __shfl_xor_sync(0xffffffffu, ljp, o);
      lj = ljp;
    } else {
      lj = jp_construct<DPAD>(th, d, s.code);
    }
    double sum = 0.0;
    if (live) {
      for (long long n = slice; n < P.N; n += S) sum += F::template obs<DPAD>(th, d, s_obs + n * P.ncols, n, P.hyper);
      if (slice == 0) sum += lj;
      if (slice == S - 1) sum += F::template prior<DPAD>(th, d, P.N, P.hyper);      // not on the lane that adds the Jacobian: shorter critical path
    }
    for (int o = S >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (live && slice == 0)
      for (int r = 0; r < n_cta; ++r) cluster.map_shared_rank(s_val, r)[point] = -sum;
  }
  cluster.sync();
  if (threadIdx.x == 0) s.n_values += 1;      // read again only after the block barriers that follow every call
  return s_val;
}

// f(xc) and O(h^4) derivatives: the stencils at h and h/2 (FdStencil::derivs + the Richardson combination of jp_hostlinalg.cpp)
template <class F, int DPAD>
__device__ __noinline__ void md_richardson(const JpModeDevParams& P, MdShared& s, const double* s_obs, double* s_val2, const double* xc,
                              double* f0, double* g, double* H) {
  const long long c_begin = clock64();
  const double* s_val = md_values<F, DPAD>(P, s, s_obs, s_val2, 0, xc, P.K);
  const long long c_mid = clock64();
  const int d = P.d, n = P.n, npairs = d * (d - 1) / 2;
  const double h = P.h, h2 = 0.5 * P.h, c0 = s_val[0];
  const double* f1 = s_val + 1;
  const double* f2 = s_val + 1 + n;
  for (int w = threadIdx.x; w < d + npairs; w += JP_MD_THREADS) {
    if (w < d) {
      const int i = w;
      const double g1 = (f1[2 * i] - f1[2 * i + 1]) / (2 * h), g2 = (f2[2 * i] - f2[2 * i + 1]) / (2 * h2);
      const double H1 = (f1[2 * i] - 2 * c0 + f1[2 * i + 1]) / (h * h), H2 = (f2[2 * i] - 2 * c0 + f2[2 * i + 1]) / (h2 * h2);
      g[i] = (4 * g2 - g1) / 3;
      H[i * d + i] = (4 * H2 - H1) / 3;
    } else {
      const int pr = w - d, q = 2 * d + 4 * pr, i = s.pi[pr], j = s.pj[pr];
      const double v1 = (f1[q] - f1[q + 1] - f1[q + 2] + f1[q + 3]) / (4 * h * h);
      const double v2 = (f2[q] - f2[q + 1] - f2[q + 2] + f2[q + 3]) / (4 * h2 * h2);
      const double v = (4 * v2 - v1) / 3;
      H[j * d + i] = v;
      H[i * d + j] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *f0 = c0;
    s.evals += 1;
    s.cyc_values += c_mid - c_begin;
    s.cyc_derivs += clock64() - c_mid;
  }
  __syncthreads();
}

__device__ __forceinline__ double md_max_abs(const double* v, int n) {
  double m = 0;
  for (int i = 0; i < n; ++i) m = fmax(m, fabs(v[i]));
  return m;
}

// Newton step by warp 0: Cholesky H = L L' with lane r holding row r, then the two triangular solves with the right-hand side in
// registers; step = -H^-1 g.  false (uniformly) unless every pivot is comfortably positive.
__device__ __noinline__ bool md_chol_step(MdShared& s, int d) {
  const int lane = threadIdx.x;
  const double* H = s.H;
  double* L = s.A;                      // L(r, k) = L[k * d + r]
  double dmax = lane < d ? fabs(H[lane * d + lane]) : 0.0;
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  double inv_diag = 0.0;                // lane c: 1 / L(c, c)
  for (int c = 0; c < d; ++c) {
    double v = 0.0;
    if (lane >= c && lane < d) {
      v = H[c * d + lane];
      for (int k = 0; k < c; ++k) v -= L[k * d + lane] * L[k * d + c];
    }
    const double piv = __shfl_sync(0xffffffffu, v, c);
    if (!(piv > 1e-8 * dmax)) return false;
    const double inv = rsqrt(piv);      // one reciprocal square root per column, no division
    if (lane == c) inv_diag = inv;
    else if (lane > c && lane < d) L[c * d + lane] = v * inv;
    __syncwarp();
  }
  double v = lane < d ? s.g[lane] : 0.0;
  for (int k = 0; k < d; ++k) {         // L z = g
    const double yk = __shfl_sync(0xffffffffu, v, k) * __shfl_sync(0xffffffffu, inv_diag, k);
    if (lane == k) v = yk;
    else if (lane > k && lane < d) v -= L[k * d + lane] * yk;
  }
  for (int k = d - 1; k >= 0; --k) {    // L' y = z
    const double yk = __shfl_sync(0xffffffffu, v, k) * __shfl_sync(0xffffffffu, inv_diag, k);
    if (lane == k) v = yk;
    else if (lane < k) v -= L[lane * d + k] * yk;
  }
  if (lane < d) s.step[lane] = -v;
  __syncwarp();
  return true;
}

// Jacobi eigen-decomposition by warp 0 (the role of symmetric_eigen in jp_hostlinalg.cpp) in the PARALLEL ordering: a sweep is
// m - 1 rounds of a round-robin tournament over the m = d (+1 if odd) indices, the floor(d / 2) disjoint rotations of a round are
// applied together -- every element of A' = J' A J and V' = V J from four (two) old elements, the lanes share the d^2 elements.
// Eigenvalue i in s.lam[i] (unsorted; s.imin = index of the smallest), its eigenvector in V[i * d + k].
__device__ __noinline__ void md_eigen_warp(MdShared& s, int d) {
  const int lane = threadIdx.x;
  double* A = s.A;                      // A(r, c) = A[c * d + r]
  double* V = s.V;
  const int dd = d * d, m = (d + 1) & ~1;
  constexpr int kPerLane = JP_MD_DMAX * JP_MD_DMAX / 32;
  int er[kPerLane], ec[kPerLane];       // (row, column) of this lane's elements
#pragma unroll
  for (int e = 0; e < kPerLane; ++e) {
    const int idx = lane + 32 * e;
    ec[e] = idx / d;
    er[e] = idx - ec[e] * d;
    if (idx < dd) {
      A[idx] = s.H[idx];
      V[idx] = er[e] == ec[e] ? 1.0 : 0.0;
    }
  }
  // this lane's pair in round 0 of the tournament; from round to round both players move up by one (mod m - 1), lane 0
  // keeps player m - 1 fixed
  int tp = lane == 0 ? m - 1 : lane % (m - 1), tq = lane == 0 ? 0 : (m - 1 - lane) % (m - 1);
  __syncwarp();
  for (int sweep = 0; sweep < 64; ++sweep) {
    const long long t_chk = clock64();
    double off = 0.0, diag = 0.0;
    if (d <= 4) {                       // a handful of elements: every lane adds them up itself, cheaper than ten shuffles
      for (int c = 0; c < d; ++c) {
        diag += A[c * d + c] * A[c * d + c];
        for (int r = 0; r < c; ++r) off += A[c * d + r] * A[c * d + r];
      }
    } else {
      if (lane < d) {
        diag = A[lane * d + lane] * A[lane * d + lane];
        for (int r = 0; r < lane; ++r) off += A[lane * d + r] * A[lane * d + r];
      }
      for (int o = 16; o > 0; o >>= 1) {
        off += __shfl_xor_sync(0xffffffffu, off, o);
        diag += __shfl_xor_sync(0xffffffffu, diag, o);
      }
    }
    // off-diagonal norm below 1e-6 of the diagonal's: eigenvalues to ~1e-12 of the largest -- ample for a step direction in a
    // region where the Hessian is indefinite (near the mode the step comes from the Cholesky factor), and for the signs
    if (off <= 1e-12 * diag || off < 1e-300) break;
    if (lane == 0) {
      s.n_sweeps += 1;
      s.cyc_chk += clock64() - t_chk;
    }
    for (int round = 0; round < m - 1; ++round) {
      const long long t_r0 = clock64();
      if (lane < m / 2) {
        int p = min(tp, tq), q = max(tp, tq);
        if (q < d) {
          double cs = 1.0, sn = 0.0;
          const double apq = A[q * d + p];
          if (apq != 0.0) {
            // A dependent FP64 operation costs this single warp ~40 cycles, a single-precision one a fraction of that.  The angle only
            // has to be CLOSE to the annihilating one (the next sweep removes what is left), so the tangent and the first guess of the
            // cosine are single-precision arithmetic on the rounded entries; one Newton step in double makes cos^2 + sin^2 = 1 to 1e-14.
            const float fpq = (float)apq, tau = __fdividef((float)A[q * d + q] - (float)A[p * d + p], 2.0f * fpq);
            float tf = __fdividef(copysignf(1.0f, tau), fabsf(tau) + __fsqrt_rn(fmaf(tau, tau, 1.0f)));
            if (!(fabsf(tf) <= 1.0f)) tf = 0.0f;      // overflow / 0 / 0 in single precision: the element is negligible or the pair degenerate
            const double c0 = (double)rsqrtf(fmaf(tf, tf, 1.0f)), t = (double)tf;
            const double u = fma(t, t, 1.0);
            cs = c0 * fma(-0.5 * u, c0 * c0, 1.5);
            sn = t * cs;
          }
          s.rot_a[p] = cs; s.rot_b[p] = -sn; s.rot_p[p] = q;
          s.rot_a[q] = cs; s.rot_b[q] = sn;  s.rot_p[q] = p;
        } else {                        // odd d: index p sits this round out
          s.rot_a[p] = 1.0; s.rot_b[p] = 0.0; s.rot_p[p] = p;
        }
        if (lane == 0) {
          tq = tq + 1 == m - 1 ? 0 : tq + 1;
        } else {
          tp = tp + 1 == m - 1 ? 0 : tp + 1;
          tq = tq + 1 == m - 1 ? 0 : tq + 1;
        }
      }
      __syncwarp();
      const long long t_r1 = clock64();
      double na[kPerLane], nv[kPerLane];
#pragma unroll
      for (int e = 0; e < kPerLane; ++e) {
        const int idx = lane + 32 * e;
        if (idx < dd) {
          const int j = ec[e], i = er[e], pi = s.rot_p[i], pj = s.rot_p[j];
          const double ai = s.rot_a[i], bi = s.rot_b[i], aj = s.rot_a[j], bj = s.rot_b[j];
          na[e] = ai * (aj * A[j * d + i] + bj * A[pj * d + i]) + bi * (aj * A[j * d + pi] + bj * A[pj * d + pi]);
          nv[e] = aj * V[j * d + i] + bj * V[pj * d + i];
        }
      }
      __syncwarp();
      const long long t_r2 = clock64();
#pragma unroll
      for (int e = 0; e < kPerLane; ++e) {
        const int idx = lane + 32 * e;
        if (idx < dd) {
          A[idx] = na[e];
          V[idx] = nv[e];
        }
      }
      __syncwarp();
      if (lane == 0) {
        s.cyc_r1 += t_r1 - t_r0;
        s.cyc_r2 += t_r2 - t_r1;
        s.cyc_r3 += clock64() - t_r2;
      }
    }
  }
  __syncwarp();
  // eigenvalues in the order they ended up on the diagonal (nobody needs them sorted, nor the eigenvectors' signs: the escape
  // tries both directions); the smallest is found by every lane for itself
  if (lane < d) s.lam[lane] = A[lane * d + lane];
  __syncwarp();
  if (lane == 0) {
    int imin = 0;
    for (int i = 1; i < d; ++i)
      if (s.lam[i] < s.lam[imin]) imin = i;
    s.imin = imin;
  }
  __syncwarp();
}

template <class F, int DPAD>
__global__ void __launch_bounds__(JP_MD_THREADS, 1) jp_mode_dev_kernel(const JpModeDevParams P) {
  extern __shared__ double dyn[];
  __shared__ MdShared s;
  double* s_obs = dyn;                                  // N x ncols
  double* s_val = dyn + (size_t)P.N * P.ncols;          // 2 x (K + 1): the two value buffers of md_values
  const int d = P.d, tid = threadIdx.x;
  for (long long i = tid; i < P.N * P.ncols; i += JP_MD_THREADS) s_obs[i] = P.obs[i];
  if (tid < d) {
    s.x[tid] = P.x0[tid];
    s.code[tid] = P.tcode[tid];
  }
  if (tid == 0) {
    int q = 0;
    for (int i = 0; i < d; ++i)
      for (int j = i + 1; j < d; ++j, ++q) { s.pi[q] = (unsigned char)i; s.pj[q] = (unsigned char)j; }
    s.split_construct = P.split;
    s.evals = 0;
    s.n_eigen = 0;
    s.n_values = 0;
    s.cyc_values = s.cyc_linalg = s.cyc_derivs = s.cyc_chol = s.cyc_eigen = 0;
    s.n_sweeps = 0;
    s.cyc_r1 = s.cyc_r2 = s.cyc_r3 = s.cyc_chk = 0;
  }
  cg::this_cluster().sync();      // every CTA of the cluster is running (and initialised) before anyone stores into its shared memory
  md_richardson<F, DPAD>(P, s, s_obs, s_val, s.x, &s.fx, s.g, s.H);
  int converged = 0, iterations = 0;
  for (int it = 0; it < P.iters; ++it) {
    iterations = it + 1;
    // ---- the step: Newton by Cholesky where the Hessian is positive definite, else the saddle-free step from its eigenpairs
    if (tid < 32) {
      const long long c_la = clock64();
      const bool pd = md_chol_step(s, d);
      double lam_min = 1.0;
      const long long c_ch = clock64();
      if (tid == 0) s.cyc_chol += c_ch - c_la;
      int action = MD_STEP;
      if (!pd) {
        md_eigen_warp(s, d);
        if (tid == 0) s.cyc_eigen += clock64() - c_ch;
        // (every lane repeats the O(d) scalars for itself; the O(d^2) products are shared out, lane i owning eigenpair / coordinate i)
        double scale = 0;
        for (int i = 0; i < d; ++i) scale = fmax(scale, fabs(s.lam[i]));
        scale = fmax(scale, 1e-300);
        const double gnorm = md_max_abs(s.g, d);
        const int imin = s.imin;
        lam_min = s.lam[imin];
        if (lam_min < -1e-8 * scale && gnorm < 1e-6 * scale) {
          // a stationary point with negative curvature: leave along the most negative eigenvector
          if (tid < d) {
            const double v = s.V[imin * d + tid], xk = s.x[tid];
            s.cand[0 * JP_MD_DMAX + tid] = xk + v;
            s.cand[1 * JP_MD_DMAX + tid] = xk + 0.25 * v;
            s.cand[2 * JP_MD_DMAX + tid] = xk - v;
            s.cand[3 * JP_MD_DMAX + tid] = xk - 0.25 * v;
          }
          action = MD_ESCAPE;
        } else {
          double c = 0.0;             // lane i: (v_i . g) / |lambda_i|
          if (tid < d) {
            for (int k = 0; k < d; ++k) c += s.V[tid * d + k] * s.g[k];
            c /= fmax(fabs(s.lam[tid]), 1e-10 * scale);
          }
          double st = 0.0;            // lane k: -sum_i v_i[k] c_i
          for (int i = 0; i < d; ++i) {
            const double ci = __shfl_sync(0xffffffffu, c, i);
            if (tid < d) st -= s.V[i * d + tid] * ci;
          }
          if (tid < d) s.step[tid] = st;
        }
        __syncwarp();
      }
      if (tid == 0) {
        if (action == MD_STEP) {
          const double nrm = md_max_abs(s.step, d);
          if (nrm > 10.0)
            for (int k = 0; k < d; ++k) s.step[k] *= 10.0 / nrm;
          if ((pd || lam_min > 0) && md_max_abs(s.step, d) < 1e-9 * (1 + md_max_abs(s.x, d))) action = MD_DONE;   // below the FD resolution
        }
        s.action = action;
        s.t = 1.0;
        s.cyc_linalg += clock64() - c_la;
        s.n_eigen += pd ? 0 : 1;
      }
    }
    __syncthreads();
    const int action = s.action;
    if (action == MD_DONE) {
      converged = 1;
      break;
    }
    if (action == MD_ESCAPE) {
      const double* fcv = md_values<F, DPAD>(P, s, s_obs, s_val, 1, nullptr, 4);
      if (tid < 4) s.fc[tid] = fcv[tid];
      __syncthreads();
      if (tid == 0) {
        s.evals += 1;
        int best = -1;
        for (int c = 0; c < 4; ++c)
          if (isfinite(s.fc[c]) && (best < 0 || s.fc[c] < s.fc[best])) best = c;
        if (best < 0 || !(s.fc[best] < s.fx)) {
          s.action = MD_DONE;
        } else {
          for (int k = 0; k < d; ++k) s.x[k] = s.cand[best * JP_MD_DMAX + k];
        }
      }
      __syncthreads();
      if (s.action == MD_DONE) break;
      md_richardson<F, DPAD>(P, s, s_obs, s_val, s.x, &s.fx, s.g, s.H);
      continue;
    }
    // ---- line search; the stencil at the candidate gives the derivatives of the next iteration
    bool ok = false;
    while (true) {
      const double t = s.t;
      if (!(t > 1e-8)) break;
      if (tid < d) s.xn[tid] = s.x[tid] + t * s.step[tid];
      __syncthreads();
      md_richardson<F, DPAD>(P, s, s_obs, s_val, s.xn, &s.fn, s.gn, s.Hn);
      const double fn = s.fn, fx = s.fx;
      if (isfinite(fn) && fn <= fx + 1e-13 * (1 + fabs(fx))) {
        ok = true;
        break;
      }
      __syncthreads();
      if (tid == 0) s.t = t * 0.25;
      __syncthreads();
    }
    if (!ok) break;
    __syncthreads();
    if (tid < d) {
      s.x[tid] = s.xn[tid];
      s.g[tid] = s.gn[tid];
    }
    for (int i = tid; i < d * d; i += JP_MD_THREADS) s.H[i] = s.Hn[i];
    if (tid == 0) s.fx = s.fn;
    __syncthreads();
  }
  cg::this_cluster().sync();      // nobody leaves while a peer may still address its shared memory
  if (cg::this_cluster().block_rank() != 0) return;
  // ---- result (written straight into the host's pinned buffer)
  if (tid < d) P.out[tid] = s.x[tid];
  for (int i = tid; i < d * d; i += JP_MD_THREADS) P.out[d + i] = s.H[i];
  if (tid == 0) {
    const double grad = md_max_abs(s.g, d);
    if (!converged && grad <= 1e-5 * (1 + fabs(s.fx))) converged = 1;
    bool finite = isfinite(s.fx) && isfinite(grad);
    for (int k = 0; k < d; ++k) finite = finite && isfinite(s.x[k]);
    for (int k = 0; k < d * d; ++k) finite = finite && isfinite(s.H[k]);
    double* o = P.out + d + d * d;
    o[0] = s.fx;
    o[1] = grad;
    o[2] = iterations;
    o[3] = converged;
    o[4] = s.evals;
    o[5] = finite ? 1.0 : 0.0;
    o[6] = (double)s.cyc_values;
    o[7] = (double)s.cyc_derivs;
    o[8] = (double)s.cyc_linalg;
    o[9] = s.n_eigen;
    o[10] = (double)s.cyc_chol;
    o[11] = (double)s.cyc_eigen;
    o[12] = s.n_sweeps;
    o[13] = (double)s.cyc_r1; o[14] = (double)s.cyc_r2; o[15] = (double)s.cyc_r3; o[16] = (double)s.cyc_chk;
  }
}

template <class F, int DPAD>
cudaError_t md_launch(const JpModeDevParams& P, size_t smem, cudaStream_t st) {
  // one CTA when the stencil x slices fit its threads (README Example 1), the 8-CTA cluster otherwise
  const unsigned n_cta = (long long)P.K * P.S <= JP_MD_THREADS ? 1u : (unsigned)JP_MD_CTAS;
  if (smem > 32 * 1024) {      // static (MdShared) + dynamic beyond the 48 KB default
    cudaError_t e = cudaFuncSetAttribute(jp_mode_dev_kernel<F, DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_cta);
  cfg.blockDim = dim3(JP_MD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = n_cta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, jp_mode_dev_kernel<F, DPAD>, P);
}
// Only the padded widths a family's shape_ok admits are instantiated (a kernel per (family, width) takes seconds to compile).
template <class F>
cudaError_t md_launch_family(const JpModeDevParams& P, size_t smem, cudaStream_t st) {
  if (P.d <= 4) return md_launch<F, 4>(P, smem, st);
  if (P.d <= 8) return md_launch<F, 8>(P, smem, st);
  if (P.d <= 12) return md_launch<F, 12>(P, smem, st);
  return md_launch<F, 16>(P, smem, st);
}
template <>
cudaError_t md_launch_family<FamBinomialMixture>(const JpModeDevParams& P, size_t smem, cudaStream_t st) {      // d = 3
  return md_launch<FamBinomialMixture, 4>(P, smem, st);
}
template <>
cudaError_t md_launch_family<FamAnova2>(const JpModeDevParams& P, size_t smem, cudaStream_t st) {               // d = 5
  return md_launch<FamAnova2, 8>(P, smem, st);
}

}  // namespace

// Try the one-launch mode search.  *used = 1: (h_x, h_H, *fx, report fields) hold its converged result.  *used = 0: not
// applicable or not converged -- the caller runs the host-driven iteration from its own start (h_x is untouched then).
int jp_mode_dev_try(jp_ctx* ctx, const jp_data* data, int d, const int* h_transform, double* h_x, double* h_H, double* fx,
                    int iters, int* evals, int* iterations, int* converged, double* grad, int* used) {
  *used = 0;
  if (std::getenv("JP_MODE_HOST")) return JP_OK;
  const int n = 2 * d + 2 * d * (d - 1), K = 1 + 2 * n;
  if (d > JP_MD_DMAX || data->N < 1 || data->N * data->ncols > JP_MD_OBS_DOUBLES || (long long)K * data->N > (1 << 17)) return JP_OK;
  const JpFamilyEntry* fam = jp_find_family(data->family);
  if (!fam || !fam->shape_ok(d, data->ncols, data->N)) return JP_OK;      // the host path reports the error
  JP_TRY(jp_check_transform_codes("jp_mode", h_transform, d));
  JP_CUDA(cudaSetDevice(ctx->device));
  JpModeDevParams P;
  P.d = d; P.ncols = data->ncols; P.iters = iters; P.K = K; P.n = n; P.N = data->N; P.h = 2e-3;
  // observation slices per stencil point: a power of two, at most a warp, the cluster's threads / points, the observations
  int S = 1;
  while (S < 32 && 2 * S * K <= JP_MD_THREADS * JP_MD_CTAS && 2 * S <= data->N) S *= 2;
  P.S = S;
  P.split = S > 1 ? 1 : 0;
  for (int k = 0; k < d; ++k)
    if (h_transform[k] != JP_T_REAL && h_transform[k] != JP_T_POSITIVE && h_transform[k] != JP_T_PROBABILITY) P.split = 0;
  P.obs = data->d_obs;
  for (int i = 0; i < JP_MAX_HYPER; ++i) P.hyper[i] = data->hyper[i];
  // zero-copy through the context's pinned buffer, as jp_log_density_points does for the host-driven iteration
  double* hp = ctx->h_pinned;
  JP_CUDA(jp_pinned_acquire(ctx));
  JP_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(hp, h_x, sizeof(double) * d);
  std::memcpy(hp + d, h_transform, sizeof(int) * d);
  const size_t code_dbl = ((size_t)d + 1) / 2;
  P.x0 = hp;
  P.tcode = reinterpret_cast<const int*>(hp + d);
  P.out = hp + d + code_dbl;
  const size_t smem = sizeof(double) * ((size_t)data->N * data->ncols + 2 * ((size_t)K + 1));
  cudaError_t e = cudaErrorInvalidValue;
  switch (data->family) {
    case JP_FAM_BINOMIAL_MIXTURE: e = md_launch_family<FamBinomialMixture>(P, smem, ctx->stream); break;
    case JP_FAM_HIER_NORMAL: e = md_launch_family<FamHierNormal>(P, smem, ctx->stream); break;
    case JP_FAM_NORMAL_LINEAR: e = md_launch_family<FamNormalLinear>(P, smem, ctx->stream); break;
    case JP_FAM_MULTINOMIAL: e = md_launch_family<FamMultinomial>(P, smem, ctx->stream); break;
    case JP_FAM_MVN_COV: e = md_launch_family<FamMvnCov>(P, smem, ctx->stream); break;
    case JP_FAM_ANOVA2: e = md_launch_family<FamAnova2>(P, smem, ctx->stream); break;
    default: return JP_OK;      // the GLM families have their analytic Newton iteration (mode_glm); constrained GLM coefficients
                                // take the host-driven search
  }
  ctx->launches++;
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    jp_set_error("jp_mode (device-resident search): %s", cudaGetErrorString(e));
    return JP_ERR_CUDA;
  }
  const double* o = P.out + d + (size_t)d * d;
  if (evals) *evals += (int)o[4];
  if (std::getenv("JP_MODE_TRACE"))
    std::fprintf(stderr, "jp_mode (one launch): d=%d K=%d S=%d N=%lld: %d iterations, %d stencil evaluations, converged %d; cycles: values %.0f, "
                 "derivatives %.0f, linear algebra %.0f (Cholesky %.0f; %d eigen-decompositions %.0f, %d sweeps: angles %.0f, elements %.0f, stores %.0f, convergence checks %.0f)\n", d, K, S, data->N, (int)o[2], (int)o[4], (int)o[3],
                 o[6], o[7], o[8], o[10], (int)o[9], o[11], (int)o[12], o[13], o[14], o[15], o[16]);
  if (o[5] != 1.0 || o[3] != 1.0) return JP_OK;        // not converged / not finite: the host-driven iteration decides
  std::memcpy(h_x, P.out, sizeof(double) * d);
  std::memcpy(h_H, P.out + d, sizeof(double) * d * d);
  *fx = o[0];
  *grad = o[1];
  *iterations = (int)o[2];
  *converged = 1;
  *used = 1;
  return JP_OK;
}
