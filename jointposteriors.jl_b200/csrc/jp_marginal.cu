// jp_marginal.cu -- STAGE 5: marginal(jp, f): weighted mean / sigma and the 100-knot Grid CDF.
//
// Replaces weights_values + marginal + Grid (reference src/marginal_posterior.jl:98-123,
// src/interp.jl:21-31,448-457), batched over K functions f:
//   mu = dot(w, v); sigma = sqrt(dot(w, v.^2) - mu^2)            marginal_posterior.jl:120-121
//   stable sort of v carrying w (simultaneous_sort!)              interp.jl:21-26
//   c = cumsum(w_sorted); itp = linear interpolant of (v_sorted, c)   interp.jl:28-31
//   value_nodes = linspace(v_min, v_max, 100); weight_nodes = (0, itp[value_nodes[2:99]], 1)   interp.jl:448-457
// Single-GPU path: 8-bit LSD radix sort of a permutation (8 passes over the order-preserving
// 64-bit image of the doubles), gather + inclusive scan, 98 binary searches.  Multi-GPU path: a
// sort-free splitter pass producing per-knot (mass below, predecessor, successor) candidates that
// the host combines across ranks; both reproduce the same interpolant, including its tie rule
// (left knot = LAST duplicate <= x, right knot = first element of the next tie group).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "jp_common.cuh"
#include "jp_sort.cuh"

#define JP_MOUT_STRIDE (2 + 2 * JP_GRID_KNOTS + 2)   // mu, sigma, value_nodes, weight_nodes, min, max

__device__ __forceinline__ double jp_knot_value(double vmin, double vmax, int i) {
  // i-th of 100 equispaced knots, end knots exact (linspace semantics, interp.jl:450).  Julia builds the
  // range in twice precision, so interior knots are the correctly rounded vmin + (i/99)(vmax - vmin);
  // double-double arithmetic reproduces that here.
  if (i <= 0) return vmin;
  if (i >= JP_GRID_KNOTS - 1) return vmax;
  const double n = (double)(JP_GRID_KNOTS - 1);
  // d = vmax - vmin as (dh, dl)
  double dh = vmax - vmin;
  double bv = dh - vmax;
  double dl = (vmax - (dh - bv)) + (-vmin - bv);
  // t = i / 99 as (th, tl)
  double th = (double)i / n;
  double tl = fma(-th, n, (double)i) / n;
  // p = t * d
  double ph = th * dh;
  double pl = fma(th, dh, -ph) + (th * dl + tl * dh);
  // vmin + p
  double sh = vmin + ph;
  double bs = sh - vmin;
  double sl = (vmin - (sh - bs)) + (ph - bs);
  return sh + (sl + pl);
}

// (sum w v, sum w v^2, min v, max v) of marginal blockIdx.y: a block owns a contiguous slice (coalesced reads, thread t
// adds its elements in ascending order, fixed tree across threads), the last block to arrive adds the block partials in
// block order -> bitwise reproducible
#define JP_MOM_THREADS 256
__global__ void __launch_bounds__(JP_MOM_THREADS)
jp_moments_kernel(const double* const* __restrict__ vptr, const double* __restrict__ w, long long M,
                  double* __restrict__ bpart /* [K][gridDim.x][4] */, unsigned int* __restrict__ counters /* [K] */,
                  double* __restrict__ out, int out_stride) {
  const int k = blockIdx.y;
  const double* v = vptr[k];
  const long long per = (M + gridDim.x - 1) / gridDim.x, b0 = (long long)blockIdx.x * per, b1 = min(M, b0 + per);
  double s1 = 0, s2 = 0, mn = INFINITY, mx = -INFINITY;
  for (long long i = b0 + threadIdx.x; i < b1; i += JP_MOM_THREADS) {
    double x = v[i], wi = w[i];
    s1 += wi * x;
    s2 += wi * (x * x);
    mn = fmin(mn, x);
    mx = fmax(mx, x);
  }
  // the four block reductions share one barrier: shuffle trees inside the warps, then thread 0 combines the warps in order
  // (the same order, hence the same bits, as jp_block_sum / jp_block_min / jp_block_max)
  __shared__ double sw[4][JP_MOM_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  s1 = jp_warp_sum(s1);
  s2 = jp_warp_sum(s2);
  mn = jp_warp_min(mn);
  mx = jp_warp_max(mx);
  if (lane == 0) { sw[0][wid] = s1; sw[1][wid] = s2; sw[2][wid] = mn; sw[3][wid] = mx; }
  __syncthreads();
  double* bp = bpart + (size_t)k * gridDim.x * 4;
  if (threadIdx.x == 0) {
    s1 = 0; s2 = 0; mn = sw[2][0]; mx = sw[3][0];
    for (int i = 0; i < JP_MOM_THREADS / 32; ++i) { s1 += sw[0][i]; s2 += sw[1][i]; }
    for (int i = 1; i < JP_MOM_THREADS / 32; ++i) { mn = fmin(mn, sw[2][i]); mx = fmax(mx, sw[3][i]); }
    double* o = bp + (size_t)blockIdx.x * 4;
    o[0] = s1; o[1] = s2; o[2] = mn; o[3] = mx;
  }
  if (!jp_last_block(counters + k, gridDim.x)) return;
  if (threadIdx.x < 32) {      // warp 0: lane l takes blocks l, l + 32, .. in ascending order, then the fixed shuffle tree
    s1 = 0; s2 = 0; mn = INFINITY; mx = -INFINITY;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      s1 += __ldcg(bp + 4 * b); s2 += __ldcg(bp + 4 * b + 1);
      mn = fmin(mn, __ldcg(bp + 4 * b + 2)); mx = fmax(mx, __ldcg(bp + 4 * b + 3));
    }
    s1 = jp_warp_sum(s1);
    s2 = jp_warp_sum(s2);
    mn = jp_warp_min(mn);
    mx = jp_warp_max(mx);
    if (lane == 0) {
      double* o = out + (size_t)k * out_stride;
      o[0] = s1; o[1] = s2; o[2] = mn; o[3] = mx;
    }
  }
}

struct ValueDigit {
  const double* const* vptr;
  int shift;
  __device__ __forceinline__ unsigned operator()(int batch, uint32_t src) const {
    return (unsigned)((jp_sortable(vptr[batch][src]) >> shift) & 0xFFull);
  }
};

// one block per marginal: gather by the sorted permutation and inclusive-scan the weights
__global__ void __launch_bounds__(1024)
jp_gather_scan_kernel(const double* const* __restrict__ vptr, const double* __restrict__ w,
                      const uint32_t* __restrict__ perm, long long M, double* __restrict__ sv,
                      double* __restrict__ sw, double* __restrict__ cw) {
  __shared__ double wsum[32];
  __shared__ double carry_s;
  const int k = blockIdx.x;
  const double* v = vptr[k];
  const uint32_t* pm = perm + (size_t)k * M;
  double* osv = sv + (size_t)k * M;
  double* osw = sw + (size_t)k * M;
  double* ocw = cw + (size_t)k * M;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  for (long long base = 0; base < M; base += 1024) {
    long long i = base + threadIdx.x;
    double x = 0, wi = 0;
    if (i < M) {
      uint32_t src = pm[i];
      x = v[src];
      wi = w[src];
      osv[i] = x;
      osw[i] = wi;
    }
    // inclusive warp scan
    double s = wi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) wsum[wid] = s;
    __syncthreads();
    if (wid == 0) {
      double t = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      wsum[lane] = t;   // inclusive prefix of warp totals
    }
    __syncthreads();
    double pre = carry_s + (wid > 0 ? wsum[wid - 1] : 0.0);
    if (i < M) ocw[i] = pre + s;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = pre + s;
    __syncthreads();
  }
}

// one block per marginal: the 100 knots from the sorted arrays
__global__ void __launch_bounds__(128)
jp_knots_kernel(const double* __restrict__ sv, const double* __restrict__ cw, long long M,
                const double* __restrict__ mom, int mom_stride, double* __restrict__ mout) {
  const int k = blockIdx.x;
  const double* v = sv + (size_t)k * M;
  const double* c = cw + (size_t)k * M;
  double* o = mout + (size_t)k * JP_MOUT_STRIDE;
  const double vmin = v[0], vmax = v[M - 1];
  int i = threadIdx.x;
  if (i == 0) {
    double s1 = mom[(size_t)k * mom_stride + 0], s2 = mom[(size_t)k * mom_stride + 1];
    o[0] = s1;
    o[1] = sqrt(s2 - s1 * s1);      // no clamp: a negative argument gives NaN, as in the reference
    o[2 + 2 * JP_GRID_KNOTS] = vmin;
    o[3 + 2 * JP_GRID_KNOTS] = vmax;
  }
  if (i >= JP_GRID_KNOTS) return;
  double x = jp_knot_value(vmin, vmax, i);
  o[2 + i] = x;
  double wn;
  if (i == 0) {
    wn = 0.0;                       // interp.jl:451
  } else if (i == JP_GRID_KNOTS - 1) {
    wn = 1.0;                       // interp.jl:452
  } else {
    // searchsortedlast: last (1-based) index with v <= x
    long long lo = 0, hi = M + 1;
    while (lo < hi - 1) {
      long long mid = (lo + hi) >> 1;
      if (x < v[mid - 1]) hi = mid; else lo = mid;
    }
    long long ix = min(max(lo, 1LL), M - 1);
    double k0 = v[ix - 1], k1 = v[ix];
    double fx = (x - k0) / (k1 - k0);
    wn = c[ix - 1] * (1.0 - fx) + c[ix] * fx;
  }
  o[2 + JP_GRID_KNOTS + i] = wn;
}

// ---- sort-free path: one pass over the values bins them between the 100 knots
//
// itp[x_i] of the reference (interp.jl:28-31,453-455) only needs, per interior knot x_i,
//     S_i    = sum of the weights of all values <= x_i                (cumulative weight of the LAST element <= x_i)
//     pred_i = largest value <= x_i, succ_i = smallest value > x_i
//     the lowest-index element attaining succ_i and its weight        (first member of the next tie group)
// so the sort + scan is replaced by binning: value v falls in bin b(v) = #{i in 1..98 : x_i < v} (0..98), a bin keeps
// (sum w, max v, min v with its lowest node index and weight), and knot i reads bins < i and >= i.  Everything is
// summed in a fixed order (lane order inside a warp step, warp order, block order), hence bitwise reproducible.
#define JP_NBINS (JP_GRID_KNOTS - 1)     // 99
#define JP_BIN_THREADS 256
#define JP_BIN_WARPS (JP_BIN_THREADS / 32)
#define JP_BIN_STRIDE 5                  // per bin: W, max, min, index of min (as double), weight at min

// The binning pass of one block over its slice of marginal k (all JP_BIN_THREADS threads): fills s_bin and writes the block's
// bin table to `o`.  s_x holds the 100 knots of [vmin, vmax].
typedef double JpBinTable[JP_BIN_WARPS][JP_NBINS][JP_BIN_STRIDE];
__device__ __forceinline__ void jp_bin_slice(const double* __restrict__ v, const double* __restrict__ w, long long M, long long m0,
                                             double vmin, double vmax, const double* s_x, JpBinTable& s_bin, double* __restrict__ o,
                                             double* mom_s1 = nullptr, double* mom_s2 = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < JP_BIN_WARPS * JP_NBINS; i += JP_BIN_THREADS) {
    double* bn = &s_bin[0][0][0] + (size_t)i * JP_BIN_STRIDE;
    bn[0] = 0.0; bn[1] = -INFINITY; bn[2] = INFINITY; bn[3] = INFINITY; bn[4] = 0.0;
  }
  __syncthreads();
  // contiguous slice per block, contiguous sub-slice per warp: node indices increase along the visiting order
  const long long per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long b0 = (long long)blockIdx.x * per_block, b1 = min(M, b0 + per_block);
  const long long per_warp = ((per_block + JP_BIN_WARPS - 1) / JP_BIN_WARPS + 31) / 32 * 32;
  const long long w0 = b0 + (long long)warp * per_warp, w1 = min(b1, w0 + per_warp);
  const double scale = (vmax > vmin) ? (double)(JP_GRID_KNOTS - 1) / (vmax - vmin) : 0.0;
  double t1 = 0, t2 = 0;      // this lane's share of sum w v, sum w v^2 (only when the caller asks for the moments)
  for (long long base = w0; base < w1; base += 32) {
    const long long j = base + lane;
    const bool live = j < w1;
    double vj = live ? v[j] : 0.0, wj = live ? w[j] : 0.0;
    if (mom_s1) {
      t1 += wj * vj;
      t2 += wj * (vj * vj);
    }
    int bin = -1;
    if (live) {
      int g = (int)floor((vj - vmin) * scale);          // guess, then make exact against the knot table
      g = max(0, min(JP_NBINS - 1, g));
      while (g < JP_NBINS - 1 && s_x[g + 1] < vj) ++g;  // b(v) = #{i >= 1 : x_i < v}
      while (g > 0 && !(s_x[g] < vj)) --g;
      bin = g;
    }
    // Every group of lanes with equal bins is folded into its lowest lane by a binary tree over the group's members:
    // in each of five rounds the holders of even rank absorb the next holder above them (fixed order: reproducible).
    // The minimum keeps the lower lane on ties, i.e. the lower node index.
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    const unsigned lt = (1u << lane) - 1u;
    const bool leader = live && (peers & lt) == 0;
    const bool ordinary = vj < INFINITY;                // NaN and +inf never become a minimum (nor did they in a serial fold)
    double sW = wj, mx = fmax((double)-INFINITY, vj), mn = ordinary ? vj : (double)INFINITY;
    int ml = ordinary ? lane : -1;                      // lane that holds the minimum
    unsigned act = peers;                               // members of my group that still hold a partial
#pragma unroll
    for (int round = 0; round < 5; ++round) {
      if (__all_sync(0xffffffffu, (act & (act - 1u)) == 0u)) break;      // every group is down to one holder
      const unsigned above = act & ~(lt | (1u << lane));
      const bool keep = ((act >> lane) & 1u) && (__popc(act & lt) & 1) == 0;
      const int src = above ? (__ffs(above) - 1) : lane;
      const double pW = __shfl_sync(0xffffffffu, sW, src), pmx = __shfl_sync(0xffffffffu, mx, src);
      const double pmn = __shfl_sync(0xffffffffu, mn, src);
      const int pml = __shfl_sync(0xffffffffu, ml, src);
      if (keep && above) {
        sW += pW;
        mx = fmax(mx, pmx);
        if (pmn < mn) { mn = pmn; ml = pml; }
      }
      act = __ballot_sync(0xffffffffu, keep) & peers;
    }
    const double mw_l = __shfl_sync(0xffffffffu, wj, max(ml, 0));
    const double mw = ml >= 0 ? mw_l : 0.0;
    const double mi = ml >= 0 ? (double)(m0 + base + ml) : INFINITY;
    if (leader) {
      double* bn = s_bin[warp][bin];
      bn[0] += sW;
      bn[1] = fmax(bn[1], mx);
      if (mn < bn[2]) { bn[2] = mn; bn[3] = mi; bn[4] = mw; }   // equal minima: the earlier (lower index) entry stays
    }
    __syncwarp();
  }
  if (mom_s1) {
    *mom_s1 = t1;
    *mom_s2 = t2;
  }
  __syncthreads();
  for (int bq = threadIdx.x; bq < JP_NBINS; bq += JP_BIN_THREADS) {
    double W = 0.0, mx = -INFINITY, mn = INFINITY, mi = INFINITY, mw = 0.0;
    for (int ww = 0; ww < JP_BIN_WARPS; ++ww) {
      const double* bn = s_bin[ww][bq];
      W += bn[0];
      mx = fmax(mx, bn[1]);
      if (bn[2] < mn) { mn = bn[2]; mi = bn[3]; mw = bn[4]; }
    }
    double* ob = o + (size_t)bq * JP_BIN_STRIDE;
    ob[0] = W; ob[1] = mx; ob[2] = mn; ob[3] = mi; ob[4] = mw;
  }
}

__global__ void __launch_bounds__(JP_BIN_THREADS)
jp_bins_kernel(const double* const* __restrict__ vptr, const double* __restrict__ w, long long M, long long m0,
               const double* __restrict__ minmax, int minmax_stride, int minmax_off,
               double* __restrict__ out /* [K][gridDim.x][JP_NBINS][JP_BIN_STRIDE] */) {
  __shared__ double s_x[JP_GRID_KNOTS];
  __shared__ JpBinTable s_bin;
  const int k = blockIdx.y;
  const double vmin = minmax[(size_t)k * minmax_stride + minmax_off], vmax = minmax[(size_t)k * minmax_stride + minmax_off + 1];
  for (int i = threadIdx.x; i < JP_GRID_KNOTS; i += JP_BIN_THREADS) s_x[i] = jp_knot_value(vmin, vmax, i);
  jp_bin_slice(vptr[k], w, M, m0, vmin, vmax, s_x, s_bin, out + ((size_t)k * gridDim.x + blockIdx.x) * JP_NBINS * JP_BIN_STRIDE);
}

// blocks -> bins -> knots of one marginal, by all threads of a block (any block size >= 64).  FINAL: the 100-knot Grid +
// (mu, sigma) into `out` (one JP_MOUT_STRIDE record); otherwise the per-knot candidates (S, pred, succ, index, weight, x)
// for the cross-rank combine.
#define JP_COMBINE_THREADS 512
struct JpCombineSmem {
  double sb[JP_NBINS][JP_BIN_STRIDE];
  int argmin[JP_NBINS];
  double sS[JP_GRID_KNOTS], sP[JP_GRID_KNOTS], sSucc[JP_GRID_KNOTS][3];
};
template <bool FINAL>
__device__ __forceinline__ void jp_combine_bins(JpCombineSmem& S, const double* __restrict__ bins_k, int nblocks, double vmin, double vmax,
                                                double s1, double s2, double* __restrict__ out) {
  const int nt = blockDim.x;
  // one thread per (bin, field): consecutive threads read consecutive doubles of a block's bin table; block order =
  // ascending node index.  The index and weight of the minimum (fields 3, 4) are fetched from the block that won.
  for (int t = threadIdx.x; t < JP_NBINS * JP_BIN_STRIDE; t += nt) {
    const int bin = t / JP_BIN_STRIDE, field = t % JP_BIN_STRIDE;
    if (field > 2) continue;
    const double* col = bins_k + t;
    double acc = field == 0 ? 0.0 : (field == 1 ? -INFINITY : INFINITY);
    int arg = -1;
#pragma unroll 8
    for (int b = 0; b < nblocks; ++b) {
      const double x = __ldcg(col + (size_t)b * JP_NBINS * JP_BIN_STRIDE);
      if (field == 0) acc += x;
      else if (field == 1) acc = fmax(acc, x);
      else if (x < acc) { acc = x; arg = b; }
    }
    S.sb[bin][field] = acc;
    if (field == 2) S.argmin[bin] = arg;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < JP_NBINS * JP_BIN_STRIDE; t += nt) {
    const int bin = t / JP_BIN_STRIDE, field = t % JP_BIN_STRIDE;
    if (field < 3) continue;
    const int arg = S.argmin[bin];
    S.sb[bin][field] = arg >= 0 ? __ldcg(bins_k + t + (size_t)arg * JP_NBINS * JP_BIN_STRIDE) : (field == 3 ? INFINITY : 0.0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // 99 bins: sequential prefix (mass, predecessor) ...
    double Sm = 0.0, pr = -INFINITY;
#pragma unroll 7
    for (int i = 1; i <= JP_NBINS - 1; ++i) {        // knot i reads bins 0 .. i-1
      Sm += S.sb[i - 1][0];
      pr = fmax(pr, S.sb[i - 1][1]);
      S.sS[i] = Sm;
      S.sP[i] = pr;
    }
  } else if (threadIdx.x == 32) {   // ... and, in another warp, suffix (successor)
    double mn = INFINITY, mi = INFINITY, mw = 0.0;
#pragma unroll 7
    for (int i = JP_NBINS - 1; i >= 1; --i) {        // knot i reads bins i .. 98; ties keep the lower bin's entry
      if (S.sb[i][2] <= mn && S.sb[i][2] < INFINITY) { mn = S.sb[i][2]; mi = S.sb[i][3]; mw = S.sb[i][4]; }
      S.sSucc[i][0] = mn; S.sSucc[i][1] = mi; S.sSucc[i][2] = mw;
    }
  }
  __syncthreads();
  if (FINAL) {
    if (threadIdx.x == 0) {
      out[0] = s1;
      out[1] = sqrt(s2 - s1 * s1);      // no clamp: a negative argument gives NaN, as in the reference
      out[2 + 2 * JP_GRID_KNOTS] = vmin;
      out[3 + 2 * JP_GRID_KNOTS] = vmax;
    }
    for (int t = threadIdx.x; t < JP_GRID_KNOTS; t += nt) {
      const double x = jp_knot_value(vmin, vmax, t);
      out[2 + t] = x;
      double wn;
      if (t == 0) wn = 0.0;                              // interp.jl:451
      else if (t == JP_GRID_KNOTS - 1) wn = 1.0;         // interp.jl:452
      else {
        const double k0 = S.sP[t], k1 = S.sSucc[t][0], c0 = S.sS[t], c1 = S.sS[t] + S.sSucc[t][2];
        const double fx = (x - k0) / (k1 - k0);
        wn = c0 * (1.0 - fx) + c1 * fx;
      }
      out[2 + JP_GRID_KNOTS + t] = wn;
    }
  } else {
    for (int t = 1 + threadIdx.x; t <= JP_GRID_KNOTS - 2; t += nt) {
      double* o = out + (size_t)(t - 1) * 6;
      o[0] = S.sS[t]; o[1] = S.sP[t]; o[2] = S.sSucc[t][0]; o[3] = S.sSucc[t][1]; o[4] = S.sSucc[t][2];
      o[5] = jp_knot_value(vmin, vmax, t);
    }
  }
}

template <bool FINAL>
__global__ void __launch_bounds__(JP_COMBINE_THREADS)
jp_bins_combine_kernel(const double* __restrict__ bins, int nblocks, const double* __restrict__ minmax, int minmax_stride,
                       int minmax_off, const double* __restrict__ mom, int mom_stride, double* __restrict__ out) {
  __shared__ JpCombineSmem S;
  const int k = blockIdx.x;
  const double vmin = minmax[(size_t)k * minmax_stride + minmax_off], vmax = minmax[(size_t)k * minmax_stride + minmax_off + 1];
  const double s1 = FINAL ? mom[(size_t)k * mom_stride + 0] : 0.0, s2 = FINAL ? mom[(size_t)k * mom_stride + 1] : 0.0;
  jp_combine_bins<FINAL>(S, bins + (size_t)k * nblocks * JP_NBINS * JP_BIN_STRIDE, nblocks, vmin, vmax, s1, s2,
                         out + (FINAL ? (size_t)k * JP_MOUT_STRIDE : (size_t)k * (JP_GRID_KNOTS - 2) * 6));
}

// ---- single-GPU default path after a fit: ONE launch per batch of coordinate marginals.  Stage 4 (jp_fit.cu) leaves, per
// stage-4 block and coordinate, (sum w theta, sum w theta^2, min theta, max theta): every block here combines the extrema of
// its coordinate itself (block order), bins its slice, and the last block of a marginal to finish (arrival counter) turns the
// block tables into the 100-knot Grid -- no separate moments pass, no second and third launch.
__global__ void __launch_bounds__(JP_BIN_THREADS)
jp_marginal_onepass_kernel(const int* __restrict__ coords, const double* __restrict__ theta, const double* __restrict__ w, long long M,
                           long long m0, const double* __restrict__ cmom /* [nb4][d][4] */, int nb4, int d,
                           double* __restrict__ bins /* [K][gridDim.x][JP_NBINS][JP_BIN_STRIDE] */, unsigned int* __restrict__ counters,
                           double* __restrict__ mout) {
  __shared__ double s_x[JP_GRID_KNOTS];
  __shared__ union U { JpBinTable bin; JpCombineSmem comb; U() {} } sh;
  __shared__ double s_red[4][JP_BIN_WARPS];
  __shared__ double s_mm[4];
  const int k = blockIdx.y, c = coords[k], lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // extrema and moments of coordinate c from the stage-4 block partials, in block order
  {
    // all threads load (thread t takes stage-4 blocks t, t + 256, .. in ascending order), fixed shuffle tree, warps in warp
    // order: reproducible, and no single thread walks the L2-resident partials alone
    double mn = INFINITY, mx = -INFINITY, s1 = 0, s2 = 0;
    for (int b = threadIdx.x; b < nb4; b += JP_BIN_THREADS) {
      const double2* q = reinterpret_cast<const double2*>(cmom + ((size_t)b * d + c) * 4);
      const double2 s12 = __ldcg(q), mm = __ldcg(q + 1);
      s1 += s12.x; s2 += s12.y;
      mn = fmin(mn, mm.x);
      mx = fmax(mx, mm.y);
    }
    mn = jp_warp_min(mn);
    mx = jp_warp_max(mx);
    s1 = jp_warp_sum(s1);
    s2 = jp_warp_sum(s2);
    if (lane == 0) { s_red[0][wid] = s1; s_red[1][wid] = s2; s_red[2][wid] = mn; s_red[3][wid] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      s1 = s_red[0][0]; s2 = s_red[1][0]; mn = s_red[2][0]; mx = s_red[3][0];
      for (int i = 1; i < JP_BIN_WARPS; ++i) {
        s1 += s_red[0][i]; s2 += s_red[1][i];
        mn = fmin(mn, s_red[2][i]); mx = fmax(mx, s_red[3][i]);
      }
      s_mm[0] = s1; s_mm[1] = s2; s_mm[2] = mn; s_mm[3] = mx;
    }
    __syncthreads();
  }
  const double vmin = s_mm[2], vmax = s_mm[3];
  for (int i = threadIdx.x; i < JP_GRID_KNOTS; i += JP_BIN_THREADS) s_x[i] = jp_knot_value(vmin, vmax, i);
  double* bins_k = bins + (size_t)k * gridDim.x * JP_NBINS * JP_BIN_STRIDE;
  jp_bin_slice(theta + (size_t)c * M, w, M, m0, vmin, vmax, s_x, sh.bin, bins_k + (size_t)blockIdx.x * JP_NBINS * JP_BIN_STRIDE);
  if (!jp_last_block(counters + k, gridDim.x)) return;
  jp_combine_bins<true>(sh.comb, bins_k, gridDim.x, vmin, vmax, s_mm[0], s_mm[1], mout + (size_t)k * JP_MOUT_STRIDE);
}

// ---- cross-rank combine on the device (node-sharded posterior): gathered moments [world][K][4], candidates
// [world][K][98][6]; every rank runs the same kernel on the same gathered bits in rank order
__global__ void jp_minmax_gathered_kernel(const double* __restrict__ gm, int world, int K, double* __restrict__ minmax) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double mn = INFINITY, mx = -INFINITY;
  for (int r = 0; r < world; ++r) {
    mn = fmin(mn, gm[((size_t)r * K + k) * 4 + 2]);
    mx = fmax(mx, gm[((size_t)r * K + k) * 4 + 3]);
  }
  minmax[2 * k] = mn;
  minmax[2 * k + 1] = mx;
}

__global__ void __launch_bounds__(128)
jp_combine_gathered_kernel(const double* __restrict__ gm, const double* __restrict__ gc, int world, int K,
                           double* __restrict__ mout) {
  const int k = blockIdx.x, t = threadIdx.x;
  double* o = mout + (size_t)k * JP_MOUT_STRIDE;
  double vmin = INFINITY, vmax = -INFINITY;
  for (int r = 0; r < world; ++r) {
    vmin = fmin(vmin, gm[((size_t)r * K + k) * 4 + 2]);
    vmax = fmax(vmax, gm[((size_t)r * K + k) * 4 + 3]);
  }
  if (t == 0) {
    double s1 = 0, s2 = 0;
    for (int r = 0; r < world; ++r) {
      s1 += gm[((size_t)r * K + k) * 4 + 0];
      s2 += gm[((size_t)r * K + k) * 4 + 1];
    }
    o[0] = s1;
    o[1] = sqrt(s2 - s1 * s1);      // no clamp, as in the reference
    o[2 + 2 * JP_GRID_KNOTS] = vmin;
    o[3 + 2 * JP_GRID_KNOTS] = vmax;
    o[2] = vmin;
    o[2 + JP_GRID_KNOTS - 1] = vmax;
    o[2 + JP_GRID_KNOTS] = 0.0;                                  // interp.jl:451
    o[2 + 2 * JP_GRID_KNOTS - 1] = 1.0;                          // interp.jl:452
  }
  if (t >= 1 && t <= JP_GRID_KNOTS - 2) {
    // left knot: the LAST element <= x over all ranks, cumulative weight = total mass <= x; right knot: the first
    // member of the next tie group = lowest global index among the smallest values > x (interp.jl:28-31)
    double tot = 0, pred = -INFINITY, succ = INFINITY, sidx = INFINITY, sw = 0, x = 0;
    for (int r = 0; r < world; ++r) {
      const double* c = gc + (((size_t)r * K + k) * (JP_GRID_KNOTS - 2) + (t - 1)) * 6;
      tot += c[0];
      pred = fmax(pred, c[1]);
      if (c[2] < succ || (c[2] == succ && c[3] < sidx)) { succ = c[2]; sidx = c[3]; sw = c[4]; }
      if (r == 0) x = c[5];
    }
    const double fx = (x - pred) / (succ - pred);
    o[2 + t] = x;
    o[2 + JP_GRID_KNOTS + t] = tot * (1.0 - fx) + (tot + sw) * fx;
  }
}

#define JP_BIN_SLICE 4096
#define JP_BIN_BLOCKS_MAX 128
// ------------------------------------------------------------------------------------ host side
// blocks per marginal of the binning kernels: slices of JP_BIN_SLICE nodes (JP_BINS_SLICE in the environment overrides, for
// experiments), at most JP_BIN_BLOCKS_MAX block tables for the last block to combine
static int bins_blocks_for(long long M) {
  static const long long slice = [] {
    const char* e = std::getenv("JP_BINS_SLICE");
    const long long v = e ? std::atoll(e) : 0;
    return v >= 256 ? v : (long long)JP_BIN_SLICE;
  }();
  return (int)std::max(1LL, std::min((long long)JP_BIN_BLOCKS_MAX, (M + slice - 1) / slice));
}

// moments of the K value columns in post->d_vptr into d_out[K][4]; partials live in the ctx scratch behind K x 4
static int launch_moments(jp_posterior* post, int K, double* d_out) {
  jp_ctx* ctx = post->ctx;
  const int nb = bins_blocks_for(post->M);
  JP_REQUIRE(K <= JP_COUNTERS && (size_t)K * 4 * (1 + nb) <= JP_SCRATCH_DOUBLES, "marginal: K=%d too large for one call", K);
  dim3 g(nb, K);
  jp_moments_kernel<<<g, JP_MOM_THREADS, 0, ctx->stream>>>(post->d_vptr, post->d_density, post->M, ctx->d_scratch + (size_t)K * 4,
                                                           ctx->d_counters, d_out, 4);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}

// light buffers of the default (sort-free) path
static int ensure_marginal_buffers(jp_posterior* post, int K) {
  if (K <= post->K_cap) return JP_OK;
  jp_ctx* c = post->ctx;
  jp_dfree(c, (void*)post->d_vptr); jp_dfree(c, post->d_bins); jp_dfree(c, post->d_mout);
  post->d_vptr = nullptr; post->d_bins = nullptr; post->d_mout = nullptr;
  post->vptr_host.clear();
  post->K_cap = 0;
  post->bins_blocks = bins_blocks_for(post->M);
  JP_CUDA(jp_dmalloc(c, (void**)&post->d_vptr, (size_t)K * sizeof(double*)));
  JP_CUDA(jp_dmalloc(c, &post->d_bins, (size_t)K * post->bins_blocks * JP_NBINS * JP_BIN_STRIDE * 8));
  JP_CUDA(jp_dmalloc(c, &post->d_mout, (size_t)K * JP_MOUT_STRIDE * 8));
  post->K_cap = K;
  return JP_OK;
}
// uploaded value columns of host closures
static int ensure_value_buffer(jp_posterior* post, int K) {
  if (K <= post->K_cap_vals) return JP_OK;
  jp_dfree(post->ctx, post->d_vals);
  post->d_vals = nullptr;
  post->K_cap_vals = 0;
  JP_CUDA(jp_dmalloc(post->ctx, &post->d_vals, (size_t)K * post->M * 8));
  post->K_cap_vals = K;
  return JP_OK;
}
// buffers of the explicit sort (jp_marginal_sorted: the `wv` field of the reference's marginal struct)
static int ensure_sort_buffers(jp_posterior* post, int K) {
  if (K <= post->K_cap_sort) return JP_OK;
  jp_ctx* c = post->ctx;
  jp_dfree(c, post->d_perm_a); jp_dfree(c, post->d_perm_b); jp_dfree(c, post->d_hist);
  jp_dfree(c, post->d_sv); jp_dfree(c, post->d_sw); jp_dfree(c, post->d_cw);
  post->d_perm_a = post->d_perm_b = post->d_hist = nullptr;
  post->d_sv = post->d_sw = post->d_cw = nullptr;
  post->K_cap_sort = 0;
  size_t KM = (size_t)K * post->M;
  int nb = jp_sort_blocks(post->M);
  JP_CUDA(jp_dmalloc(c, &post->d_perm_a, KM * 4));
  JP_CUDA(jp_dmalloc(c, &post->d_perm_b, KM * 4));
  JP_CUDA(jp_dmalloc(c, &post->d_hist, (size_t)K * JP_SORT_BINS * nb * 4));
  JP_CUDA(jp_dmalloc(c, &post->d_sv, KM * 8));
  JP_CUDA(jp_dmalloc(c, &post->d_sw, KM * 8));
  JP_CUDA(jp_dmalloc(c, &post->d_cw, KM * 8));
  post->K_cap_sort = K;
  return JP_OK;
}

// fills post->d_vptr with the K value-column pointers
static int set_value_pointers(jp_posterior* post, int K, const int* h_coords, const double* d_values) {
  jp_ctx* ctx = post->ctx;
  JP_REQUIRE(K >= 1 && K <= 4096, "marginal: K=%d out of range", K);
  JP_REQUIRE((h_coords != nullptr) != (d_values != nullptr), "marginal: give exactly one of coords / values");
  std::vector<const double*> want((size_t)K);
  if (h_coords && post->raw) {
    // RawBuild: d_theta is the unconstrained cache.  Identity coordinates are their own columns; as soon as one requested
    // coordinate is constrained, the K columns are constructed on the device (update!(Theta) of the reference,
    // src/marginal_posterior.jl:86-90) into the value buffer and used from there.
    bool constrained = false;
    for (int k = 0; k < K; ++k) {
      JP_REQUIRE(h_coords[k] >= 0 && h_coords[k] < post->d, "marginal: coordinate %d out of range [0,%d)", h_coords[k], post->d);
      constrained = constrained || post->tcode_host[(size_t)h_coords[k]] != JP_T_REAL;
    }
    if (constrained) {
      JP_TRY(ensure_value_buffer(post, K));
      int* d_c = nullptr;
      JP_CUDA(jp_dmalloc(ctx, &d_c, (size_t)K * sizeof(int)));
      JP_CUDA(jp_pinned_acquire(ctx));
      int* hc = reinterpret_cast<int*>(ctx->h_pinned);
      std::copy(h_coords, h_coords + K, hc);
      JP_CUDA(cudaMemcpyAsync(d_c, hc, (size_t)K * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      JP_CUDA(jp_pinned_publish(ctx));
      JP_TRY(jp_construct_columns(post, K, d_c, post->d_vals));
      jp_dfree(ctx, d_c);
      h_coords = nullptr;
      d_values = post->d_vals;
    }
  }
  for (int k = 0; k < K; ++k) {
    if (h_coords) {
      JP_REQUIRE(h_coords[k] >= 0 && h_coords[k] < post->d, "marginal: coordinate %d out of range [0,%d)", h_coords[k], post->d);
      want[k] = post->d_theta + (size_t)h_coords[k] * post->M;
    } else {
      want[k] = d_values + (size_t)k * post->M;
    }
  }
  // the same columns as last time (every coordinate after each fit, say): the table on the device is already right
  if (post->vptr_host.size() < (size_t)K || !std::equal(want.begin(), want.end(), post->vptr_host.begin())) {
    JP_CUDA(jp_pinned_acquire(ctx));
    const double** hp = reinterpret_cast<const double**>(ctx->h_pinned);
    std::copy(want.begin(), want.end(), hp);
    JP_CUDA(cudaMemcpyAsync((void*)post->d_vptr, hp, (size_t)K * sizeof(double*), cudaMemcpyHostToDevice, ctx->stream));
    JP_CUDA(jp_pinned_publish(ctx));
    post->vptr_host = want;
  }
  post->sorted_valid = false;
  post->K_last = 0;
  return JP_OK;
}

// default path: moments -> bins -> knots, three launches for any K
static int run_marginals(jp_posterior* post, int K, double* h_mu, double* h_sigma, double* h_vn, double* h_wn) {
  jp_ctx* ctx = post->ctx;
  const long long M = post->M;
  JP_REQUIRE(M >= 2, "marginal: need at least 2 nodes");
  JP_REQUIRE((size_t)K * JP_MOUT_STRIDE <= JP_PINNED_DOUBLES - JP_PINNED_TAIL_DOUBLES, "marginal: K=%d too large for one call", K);
  JP_REQUIRE((size_t)K * 4 <= JP_SCRATCH_DOUBLES, "marginal: K=%d too large for one call", K);
  cudaStream_t st = ctx->stream;
  double* d_mom = ctx->d_scratch;   // K x 4: sum w v, sum w v^2, min, max
  JP_MARK(ctx, "marginals:start");
  JP_TRY(launch_moments(post, K, d_mom));      // counts its own launch
  JP_MARK(ctx, "marginals:moments");
  dim3 gb(post->bins_blocks, K);
  jp_bins_kernel<<<gb, JP_BIN_THREADS, 0, st>>>(post->d_vptr, post->d_density, M, post->m0, d_mom, 4, 2, post->d_bins);
  JP_CHECK_LAUNCH(ctx);
  JP_MARK(ctx, "marginals:bins");
  jp_bins_combine_kernel<true><<<K, JP_COMBINE_THREADS, 0, st>>>(post->d_bins, post->bins_blocks, d_mom, 4, 2, d_mom, 4, post->d_mout);
  JP_CHECK_LAUNCH(ctx);
  JP_MARK(ctx, "marginals:combine");
  JP_CUDA(cudaMemcpyAsync(ctx->h_pinned, post->d_mout, (size_t)K * JP_MOUT_STRIDE * 8, cudaMemcpyDeviceToHost, st));
  JP_MARK(ctx, "marginals:d2h");
  JP_TRY(jp_fit_tc_verify_prefetch(post));
  JP_CUDA(cudaStreamSynchronize(st));
  JP_TRY(jp_fit_tc_verify(post));      // a series length decided on the device is read back here (no-op otherwise)
  for (int k = 0; k < K; ++k) {
    const double* o = ctx->h_pinned + (size_t)k * JP_MOUT_STRIDE;
    if (h_mu) h_mu[k] = o[0];
    if (h_sigma) h_sigma[k] = o[1];
    if (h_vn) std::copy(o + 2, o + 2 + JP_GRID_KNOTS, h_vn + (size_t)k * JP_GRID_KNOTS);
    if (h_wn) std::copy(o + 2 + JP_GRID_KNOTS, o + 2 + 2 * JP_GRID_KNOTS, h_wn + (size_t)k * JP_GRID_KNOTS);
  }
  post->K_last = K;
  return JP_OK;
}

// coordinate marginals right after a single-GPU fit: the moments and extrema come from stage 4, one launch does the rest
static int run_marginals_onepass(jp_posterior* post, int K, const int* h_coords, double* h_mu, double* h_sigma, double* h_vn,
                                 double* h_wn) {
  jp_ctx* ctx = post->ctx;
  const long long M = post->M;
  JP_REQUIRE(M >= 2, "marginal: need at least 2 nodes");
  JP_REQUIRE((size_t)K * JP_MOUT_STRIDE <= JP_PINNED_DOUBLES - JP_PINNED_TAIL_DOUBLES && K <= JP_COUNTERS, "marginal: K=%d too large for one call", K);
  cudaStream_t st = ctx->stream;
  if (post->coords_host.size() != (size_t)K || !std::equal(h_coords, h_coords + K, post->coords_host.begin())) {
    jp_dfree(ctx, post->d_coords);
    post->d_coords = nullptr;
    post->coords_host.clear();
    JP_CUDA(jp_dmalloc(ctx, &post->d_coords, (size_t)K * sizeof(int)));
    JP_CUDA(jp_pinned_acquire(ctx));
    int* hp = reinterpret_cast<int*>(ctx->h_pinned);
    std::copy(h_coords, h_coords + K, hp);
    JP_CUDA(cudaMemcpyAsync(post->d_coords, hp, (size_t)K * sizeof(int), cudaMemcpyHostToDevice, st));
    JP_CUDA(jp_pinned_publish(ctx));
    post->coords_host.assign(h_coords, h_coords + K);
  }
  JP_MARK(ctx, "marginals:start");
  dim3 gb(post->bins_blocks, K);
  jp_marginal_onepass_kernel<<<gb, JP_BIN_THREADS, 0, st>>>(post->d_coords, post->d_theta, post->d_density, M, post->m0, post->d_cmom,
                                                            post->cmom_blocks, post->d, post->d_bins, ctx->d_counters, post->d_mout);
  JP_CHECK_LAUNCH(ctx);
  JP_MARK(ctx, "marginals:onepass");
  JP_CUDA(cudaMemcpyAsync(ctx->h_pinned, post->d_mout, (size_t)K * JP_MOUT_STRIDE * 8, cudaMemcpyDeviceToHost, st));
  JP_TRY(jp_fit_tc_verify_prefetch(post));
  JP_CUDA(cudaStreamSynchronize(st));
  JP_TRY(jp_fit_tc_verify(post));      // a series length decided on the device is read back here (no-op otherwise)
  for (int k = 0; k < K; ++k) {
    const double* o = ctx->h_pinned + (size_t)k * JP_MOUT_STRIDE;
    if (h_mu) h_mu[k] = o[0];
    if (h_sigma) h_sigma[k] = o[1];
    if (h_vn) std::copy(o + 2, o + 2 + JP_GRID_KNOTS, h_vn + (size_t)k * JP_GRID_KNOTS);
    if (h_wn) std::copy(o + 2 + JP_GRID_KNOTS, o + 2 + 2 * JP_GRID_KNOTS, h_wn + (size_t)k * JP_GRID_KNOTS);
  }
  post->K_last = K;
  return JP_OK;
}

// explicit stable sort + cumulative weights of the K_last marginals of the last call (simultaneous_sort! +
// cumsum, reference src/interp.jl:21-31); the knots computed from it (jp_knots_kernel) cross-check the bins
// Vandermonde!(m, density, mu, sigma), reference src/marginal_posterior.jl:44-67: per sorted element the standardised
// value z = (v - mu) / sigma and column (1, z, z^2, .., z^9) of the 10 x M design matrix of the smooth-CDF fit, with the
// reference's own multiplication order; ind = the sort permutation as 0-based GLOBAL node indices.
__global__ void jp_vandermonde_kernel(const double* __restrict__ sv, const uint32_t* __restrict__ perm, long long M, long long m0,
                                      const double* __restrict__ mom, double* __restrict__ V, long long* __restrict__ ind) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double mu = mom[0], sigma = sqrt(mom[1] - mom[0] * mom[0]);      // calc_mu_sigma, :79-86
  const double z = (sv[i] - mu) / sigma;      // :52
  const double z2 = z * z, z4 = z2 * z2;      // :53-54
  double* o = V + (size_t)i * 10;
  o[0] = 1.0;
  o[1] = z;                                   // :55
  o[2] = z2;
  o[3] = z * z2;
  o[4] = z4;
  o[5] = z4 * z;
  o[6] = z4 * z2;
  o[7] = z4 * z2 * z;
  o[8] = z4 * z4;                             // vi4^2
  o[9] = z4 * z4 * z;
  ind[i] = m0 + (long long)perm[i];
}

static int run_sort(jp_posterior* post) {
  jp_ctx* ctx = post->ctx;
  const long long M = post->M;
  const int K = post->K_last;
  JP_TRY(ensure_sort_buffers(post, K));
  cudaStream_t st = ctx->stream;
  dim3 gi((unsigned)((M + 255) / 256), K);
  jp_iota_kernel<<<gi, 256, 0, st>>>(post->d_perm_a, M, M);
  JP_CHECK_LAUNCH(ctx);
  uint32_t *pin = post->d_perm_a, *pout = post->d_perm_b;
  for (int b = 0; b < 8; ++b) {
    ValueDigit f{post->d_vptr, 8 * b};
    JP_TRY(jp_radix_pass(ctx, f, pin, pout, M, M, post->d_hist, K));
    std::swap(pin, pout);
  }
  jp_gather_scan_kernel<<<K, 1024, 0, st>>>(post->d_vptr, post->d_density, pin, M, post->d_sv, post->d_sw, post->d_cw);
  JP_CHECK_LAUNCH(ctx);
  post->sorted_valid = true;
  return JP_OK;
}

// Device-resident design of the smooth-CDF fit for marginal k of the last call: sorts if needed, fills d_V (10 x M
// column-major, caller frees with jp_dfree) and optionally the permutation; the cumulative weights are post->d_cw + k * M.
int jp_marginal_design_device(jp_posterior* post, int k, double** d_V_out, long long** d_ind_out, double* h_mu, double* h_sigma) {
  JP_REQUIRE(post && k >= 0 && k < post->K_last, "jp_marginal_buffer: marginal %d was not computed by the last call", k);
  JP_ENTER_CTX(post->ctx);
  JP_REQUIRE(post->M == post->grid->M, "jp_marginal_buffer: the posterior holds a node shard (%lld of %lld nodes)", post->M,
             post->grid->M);
  jp_ctx* ctx = post->ctx;
  if (!post->sorted_valid) JP_TRY(run_sort(post));      // 8 radix passes: the final permutation is back in d_perm_a
  const long long M = post->M;
  const size_t off = (size_t)k * M;
  cudaStream_t st = ctx->stream;
  double* d_V = nullptr;
  long long* d_ind = nullptr;
  JP_CUDA(jp_dmalloc(ctx, &d_V, (size_t)M * 10 * 8));
  JP_CUDA(jp_dmalloc(ctx, &d_ind, (size_t)M * 8));
  double* d_mom = ctx->d_scratch;                       // K x 4: sum w v, sum w v^2, min, max (calc_mu_sigma, :79-86)
  int status = launch_moments(post, post->K_last, d_mom);
  if (status == JP_OK) {
    jp_vandermonde_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(post->d_sv + off, post->d_perm_a + off, M, post->m0,
                                                                         d_mom + 4 * k, d_V, d_ind);
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_pinned, d_mom + 4 * k, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      jp_set_error("jp_marginal_buffer: %s", cudaGetErrorString(e));
      status = JP_ERR_CUDA;
    }
  }
  if (status != JP_OK) {      // the buffers go back to the pool on every path
    jp_dfree(ctx, d_V);
    jp_dfree(ctx, d_ind);
    return status;
  }
  const double m1 = ctx->h_pinned[0], m2 = ctx->h_pinned[1];
  if (h_mu) *h_mu = m1;
  if (h_sigma) *h_sigma = std::sqrt(m2 - m1 * m1);
  *d_V_out = d_V;
  if (d_ind_out) *d_ind_out = d_ind;
  else jp_dfree(ctx, d_ind);
  return JP_OK;
}

extern "C" {

int jp_marginal_coords(jp_posterior* post, int K, const int* h_coords, double* h_mu, double* h_sigma,
                       double* h_value_nodes, double* h_weight_nodes) {
  JP_REQUIRE(post && h_coords, "jp_marginal_coords: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(ensure_marginal_buffers(post, K));
  JP_TRY(set_value_pointers(post, K, h_coords, nullptr));
  static const bool no_onepass = getenv("JP_NO_ONEPASS") != nullptr;      // A/B aid: the three-launch path
  if (post->cmom_valid && !no_onepass) return run_marginals_onepass(post, K, h_coords, h_mu, h_sigma, h_value_nodes, h_weight_nodes);
  return run_marginals(post, K, h_mu, h_sigma, h_value_nodes, h_weight_nodes);
}

int jp_marginal_values(jp_posterior* post, int K, const double* h_values, double* h_mu, double* h_sigma,
                       double* h_value_nodes, double* h_weight_nodes) {
  JP_REQUIRE(post && h_values, "jp_marginal_values: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(ensure_marginal_buffers(post, K));
  JP_TRY(ensure_value_buffer(post, K));
  JP_CUDA(cudaMemcpyAsync(post->d_vals, h_values, (size_t)K * post->M * 8, cudaMemcpyHostToDevice, post->ctx->stream));
  JP_TRY(set_value_pointers(post, K, nullptr, post->d_vals));
  return run_marginals(post, K, h_mu, h_sigma, h_value_nodes, h_weight_nodes);
}

int jp_marginal_sorted(jp_posterior* post, int k, double* h_sv, double* h_sw, double* h_cw) {
  JP_REQUIRE(post && k >= 0 && k < post->K_last, "jp_marginal_sorted: marginal %d was not computed by the last call", k);
  JP_ENTER_CTX(post->ctx);
  if (!post->sorted_valid) JP_TRY(run_sort(post));
  size_t off = (size_t)k * post->M, bytes = (size_t)post->M * 8;
  cudaStream_t st = post->ctx->stream;
  if (h_sv) JP_CUDA(cudaMemcpyAsync(h_sv, post->d_sv + off, bytes, cudaMemcpyDeviceToHost, st));
  if (h_sw) JP_CUDA(cudaMemcpyAsync(h_sw, post->d_sw + off, bytes, cudaMemcpyDeviceToHost, st));
  if (h_cw) JP_CUDA(cudaMemcpyAsync(h_cw, post->d_cw + off, bytes, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  return JP_OK;
}

int jp_marginal_buffer(jp_posterior* post, int k, long long* h_ind, double* h_cum_w, double* h_V, double* h_mu,
                       double* h_sigma) {
  double* d_V = nullptr;
  long long* d_ind = nullptr;
  JP_TRY(jp_marginal_design_device(post, k, &d_V, &d_ind, h_mu, h_sigma));
  jp_ctx* ctx = post->ctx;
  const long long M = post->M;
  cudaStream_t st = ctx->stream;
  if (h_V) JP_CUDA(cudaMemcpyAsync(h_V, d_V, (size_t)M * 10 * 8, cudaMemcpyDeviceToHost, st));
  if (h_ind) JP_CUDA(cudaMemcpyAsync(h_ind, d_ind, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
  if (h_cum_w) JP_CUDA(cudaMemcpyAsync(h_cum_w, post->d_cw + (size_t)k * M, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  jp_dfree(ctx, d_V);
  jp_dfree(ctx, d_ind);
  return JP_OK;
}

int jp_marginal_knots_from_sort(jp_posterior* post, int K, double* h_value_nodes, double* h_weight_nodes) {
  JP_REQUIRE(post && K >= 1 && K <= post->K_last, "jp_marginal_knots_from_sort: no marginal batch of %d to sort", K);
  JP_ENTER_CTX(post->ctx);
  jp_ctx* ctx = post->ctx;
  if (!post->sorted_valid) JP_TRY(run_sort(post));
  cudaStream_t st = ctx->stream;
  double* d_mom = ctx->d_scratch;
  JP_TRY(launch_moments(post, K, d_mom));      // counts its own launch
  jp_knots_kernel<<<K, 128, 0, st>>>(post->d_sv, post->d_cw, post->M, d_mom, 4, post->d_mout);
  JP_CHECK_LAUNCH(ctx);
  JP_CUDA(cudaMemcpyAsync(ctx->h_pinned, post->d_mout, (size_t)K * JP_MOUT_STRIDE * 8, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < K; ++k) {
    const double* o = ctx->h_pinned + (size_t)k * JP_MOUT_STRIDE;
    if (h_value_nodes) std::copy(o + 2, o + 2 + JP_GRID_KNOTS, h_value_nodes + (size_t)k * JP_GRID_KNOTS);
    if (h_weight_nodes) std::copy(o + 2 + JP_GRID_KNOTS, o + 2 + 2 * JP_GRID_KNOTS, h_weight_nodes + (size_t)k * JP_GRID_KNOTS);
  }
  return JP_OK;
}

int jp_marginal_local_moments(jp_posterior* post, int K, const int* h_coords, const double* d_values, double* d_out) {
  JP_REQUIRE(post && d_out, "jp_marginal_local_moments: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(ensure_marginal_buffers(post, K));
  JP_TRY(set_value_pointers(post, K, h_coords, d_values));
  JP_TRY(launch_moments(post, K, d_out));      // counts its own launch
  post->K_last = K;
  return JP_OK;
}

int jp_marginal_local_knots(jp_posterior* post, int K, const int* h_coords, const double* d_values,
                            const double* d_minmax, double* d_out) {
  JP_REQUIRE(post && d_minmax && d_out, "jp_marginal_local_knots: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(ensure_marginal_buffers(post, K));
  if (h_coords || d_values) JP_TRY(set_value_pointers(post, K, h_coords, d_values));
  else JP_REQUIRE(K == post->K_last, "jp_marginal_local_knots: no value columns given and the last call had %d, not %d", post->K_last, K);
  dim3 gb(post->bins_blocks, K);
  jp_bins_kernel<<<gb, JP_BIN_THREADS, 0, post->ctx->stream>>>(post->d_vptr, post->d_density, post->M, post->m0, d_minmax, 2,
                                                                0, post->d_bins);
  JP_CHECK_LAUNCH(post->ctx);
  jp_bins_combine_kernel<false><<<K, JP_COMBINE_THREADS, 0, post->ctx->stream>>>(post->d_bins, post->bins_blocks, d_minmax, 2, 0, nullptr, 0,
                                                                   d_out);
  JP_CHECK_LAUNCH(post->ctx);
  post->K_last = K;
  return JP_OK;
}

int jp_marginal_local_knots_gathered(jp_posterior* post, int K, const int* h_coords, const double* d_values,
                                     const double* d_gathered_moments, int world, double* d_out) {
  JP_REQUIRE(post && d_gathered_moments && d_out && world >= 1, "jp_marginal_local_knots_gathered: bad argument");
  JP_ENTER_CTX(post->ctx);
  JP_REQUIRE((size_t)K * 2 <= JP_BPART_DOUBLES, "marginal: K=%d too large for one call", K);
  double* d_minmax = post->ctx->d_bpart;    // K x 2 (no reduction of this ctx is in flight: one stream)
  jp_minmax_gathered_kernel<<<(K + 127) / 128, 128, 0, post->ctx->stream>>>(d_gathered_moments, world, K, d_minmax);
  JP_CHECK_LAUNCH(post->ctx);
  return jp_marginal_local_knots(post, K, h_coords, d_values, d_minmax, d_out);
}

int jp_marginal_combine_gathered(jp_posterior* post, int K, int world, const double* d_gathered_moments,
                                 const double* d_gathered_cands, double* h_mu, double* h_sigma, double* h_vn, double* h_wn) {
  JP_REQUIRE(post && d_gathered_moments && d_gathered_cands && world >= 1, "jp_marginal_combine_gathered: bad argument");
  JP_ENTER_CTX(post->ctx);
  jp_ctx* ctx = post->ctx;
  JP_REQUIRE((size_t)K * JP_MOUT_STRIDE <= JP_PINNED_DOUBLES - JP_PINNED_TAIL_DOUBLES, "marginal: K=%d too large for one call", K);
  JP_TRY(ensure_marginal_buffers(post, K));
  cudaStream_t st = ctx->stream;
  jp_combine_gathered_kernel<<<K, 128, 0, st>>>(d_gathered_moments, d_gathered_cands, world, K, post->d_mout);
  JP_CHECK_LAUNCH(ctx);
  JP_CUDA(cudaMemcpyAsync(ctx->h_pinned, post->d_mout, (size_t)K * JP_MOUT_STRIDE * 8, cudaMemcpyDeviceToHost, st));
  JP_TRY(jp_fit_tc_verify_prefetch(post));
  JP_CUDA(cudaStreamSynchronize(st));
  JP_TRY(jp_fit_tc_verify(post));      // a series length decided on the device is read back here (no-op otherwise)
  for (int k = 0; k < K; ++k) {
    const double* o = ctx->h_pinned + (size_t)k * JP_MOUT_STRIDE;
    if (h_mu) h_mu[k] = o[0];
    if (h_sigma) h_sigma[k] = o[1];
    if (h_vn) std::copy(o + 2, o + 2 + JP_GRID_KNOTS, h_vn + (size_t)k * JP_GRID_KNOTS);
    if (h_wn) std::copy(o + 2 + JP_GRID_KNOTS, o + 2 + 2 * JP_GRID_KNOTS, h_wn + (size_t)k * JP_GRID_KNOTS);
  }
  return JP_OK;
}

// marginal(jp, f) of K coordinates of a node-sharded posterior with the two exchanges inside the library (csrc/jp_comm.cu):
// local moments -> exchange -> knot candidates against the global extrema -> exchange -> combine in rank order -> host
int jp_marginal_coords_p2p(jp_posterior* post, jp_comm* comm, int K, const int* h_coords, double* h_mu, double* h_sigma,
                           double* h_vn, double* h_wn) {
  JP_REQUIRE(post && comm && h_coords, "jp_marginal_coords_p2p: null argument");
  JP_REQUIRE(comm->ctx == post->ctx, "jp_marginal_coords_p2p: the communicator belongs to another context");
  JP_REQUIRE(K >= 1, "jp_marginal_coords_p2p: K=%d", K);
  JP_ENTER_CTX(post->ctx);
  jp_ctx* ctx = post->ctx;
  cudaStream_t st = ctx->stream;
  const int world = comm->world;
  for (int k0 = 0; k0 < K; k0 += JP_COMM_KMAX) {       // batches of at most JP_COMM_KMAX marginals per exchange
    const int Kb = std::min(JP_COMM_KMAX, K - k0);
    JP_REQUIRE(post->M >= 1, "marginal: this rank owns no node");
    JP_TRY(ensure_marginal_buffers(post, Kb));
    JP_TRY(set_value_pointers(post, Kb, h_coords + k0, nullptr));
    if (Kb > post->K_cap_cand) {
      jp_dfree(ctx, post->d_cand);
      post->d_cand = nullptr;
      post->K_cap_cand = 0;
      JP_CUDA(jp_dmalloc(ctx, &post->d_cand, (size_t)Kb * (JP_GRID_KNOTS - 2) * 6 * 8));
      post->K_cap_cand = Kb;
    }
    JP_MARK(ctx, "marginals:start");
    double* d_mom = ctx->d_scratch;
    JP_TRY(launch_moments(post, Kb, d_mom));
    const double *gm = nullptr, *gc = nullptr;
    JP_TRY(jp_comm_exchange(comm, JP_CH_MOM, d_mom, Kb * 4, &gm));
    JP_MARK(ctx, "marginals:moments_exchanged");
    double* d_minmax = ctx->d_bpart;
    jp_minmax_gathered_kernel<<<(Kb + 127) / 128, 128, 0, st>>>(gm, world, Kb, d_minmax);
    JP_CHECK_LAUNCH(ctx);
    dim3 gb(post->bins_blocks, Kb);
    jp_bins_kernel<<<gb, JP_BIN_THREADS, 0, st>>>(post->d_vptr, post->d_density, post->M, post->m0, d_minmax, 2, 0, post->d_bins);
    JP_CHECK_LAUNCH(ctx);
    jp_bins_combine_kernel<false><<<Kb, JP_COMBINE_THREADS, 0, st>>>(post->d_bins, post->bins_blocks, d_minmax, 2, 0, nullptr, 0, post->d_cand);
    JP_CHECK_LAUNCH(ctx);
    JP_TRY(jp_comm_exchange(comm, JP_CH_KNOTS, post->d_cand, Kb * (JP_GRID_KNOTS - 2) * 6, &gc));
    JP_MARK(ctx, "marginals:knots_exchanged");
    jp_combine_gathered_kernel<<<Kb, 128, 0, st>>>(gm, gc, world, Kb, post->d_mout);
    JP_CHECK_LAUNCH(ctx);
    JP_CUDA(jp_pinned_acquire(ctx));
    JP_CUDA(cudaMemcpyAsync(ctx->h_pinned, post->d_mout, (size_t)Kb * JP_MOUT_STRIDE * 8, cudaMemcpyDeviceToHost, st));
    JP_TRY(jp_fit_tc_verify_prefetch(post));
    JP_CUDA(cudaStreamSynchronize(st));
    JP_TRY(jp_fit_tc_verify(post));
    for (int k = 0; k < Kb; ++k) {
      const double* o = ctx->h_pinned + (size_t)k * JP_MOUT_STRIDE;
      if (h_mu) h_mu[k0 + k] = o[0];
      if (h_sigma) h_sigma[k0 + k] = o[1];
      if (h_vn) std::copy(o + 2, o + 2 + JP_GRID_KNOTS, h_vn + (size_t)(k0 + k) * JP_GRID_KNOTS);
      if (h_wn) std::copy(o + 2 + JP_GRID_KNOTS, o + 2 + 2 * JP_GRID_KNOTS, h_wn + (size_t)(k0 + k) * JP_GRID_KNOTS);
    }
    post->K_last = Kb;
  }
  return JP_OK;
}

}  // extern "C"
