// jp_fit.cu -- STAGES 2-4 (FP64 path): node -> theta map, node x observation log-density through
// the family plugins, max / log-sum-exp weight normalisation.
//
// Replaces eval_grid! (reference call sites src/joint_posterior.jl:180,186) and the per-node closures
// log_density! / log_density_cache (reference src/joint_posterior.jl:147-154):
//     x_m = mu_hat + U z_m                       (z_m from the integer node keys of stage 1)
//     theta_m = transform(x_m), lj_m = log|J|    (ConstrainedParameters' construct/update!)
//     ld_m = log_density(theta_m, data) + lj_m + neg_min
//     a_m  = ld_m + |z_m|^2/2 ;  density_m = w_m exp(a_m - max a) / sum_m' w_m' exp(a_m' - max a)
//
// Kernel layout of the log-density kernel: one thread per grid node (theta lives in registers),
// a block of JP_FIT_THREADS nodes streams the observation records through shared memory in
// coalesced contiguous tiles (the row-major record array is read exactly once per block), and
// gridDim.y splits the observations when there are too few node blocks to fill 148 SMs.  The
// per-node sums are sequential in n inside a split and the splits are combined in order, so the
// result is deterministic.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cooperative_groups.h>
#include "jp_common.cuh"
#include "jp_family.cuh"
#include "jp_construct.cuh"
namespace cg = cooperative_groups;

#define JP_FIT_THREADS 128
#define JP_FIT_TILE_DOUBLES 4096     // shared-memory tile of observation records (32 KB)

struct JpFitLaunchParams {
  int d, p, ncols, rule;
  long long N, M, m0;          // M = local node count, m0 = first global node of the shard
  long long M_grid;            // global node count (stride of the SoA key array)
  int splits;                  // gridDim.y
  long long obs_per_split;
  double neg_min;
  const uint8_t* idx;          // grid keys SoA [p][M_grid]
  const double* mu;            // d
  const double* U;             // d x p column-major
  const int* tcode;            // d
  const double* obs;           // N x ncols row-major
  double hyper[JP_MAX_HYPER];
  double* theta;               // [d][M]
  double* lj_prior;            // [M]: log-Jacobian + prior(theta)
  double* part;                // [splits][M] observation sums
  const double* xpts;          // non-null: explicit unconstrained points [M][d] instead of grid keys
  int raw;                     // RawBuild: theta receives the unconstrained coordinates
};

__constant__ double c_fit_nodes[2][JP_RULE_NMAX];

template <class F, int DPAD>
__global__ void __launch_bounds__(JP_FIT_THREADS)
jp_fit_nodes_kernel(const JpFitLaunchParams P) {
  extern __shared__ double tile[];
  double* s_U = tile + JP_FIT_TILE_DOUBLES;          // d x p
  double* s_mu = s_U + P.d * P.p;                     // d
  int* s_code = reinterpret_cast<int*>(s_mu + P.d);   // d
  for (int i = threadIdx.x; i < P.d; i += JP_FIT_THREADS) {
    s_mu[i] = P.mu ? P.mu[i] : 0.0;
    s_code[i] = P.tcode[i];
  }
  if (P.U)
    for (int i = threadIdx.x; i < P.d * P.p; i += JP_FIT_THREADS) s_U[i] = P.U[i];
  __syncthreads();

  const long long m = (long long)blockIdx.x * JP_FIT_THREADS + threadIdx.x;   // local node
  const bool live = m < P.M;
  double th[DPAD];
#pragma unroll
  for (int k = 0; k < DPAD; ++k) th[k] = 0.0;
  double ljp = 0.0;
  if (live) {
    // ---- stage 2: affine map from the integer key, then the constraint transforms
    if (P.xpts) {
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < P.d) th[k] = P.xpts[(size_t)m * P.d + k];
    } else {
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < P.d) th[k] = s_mu[k];
    }
    for (int j = 0; j < (P.xpts ? 0 : P.p); ++j) {
      int key = P.idx[(size_t)j * P.M_grid + (P.m0 + m)];
      if (key != 0) {   // key 0 is the centre node z = 0: most coordinates of a sparse-grid node
        double z = c_fit_nodes[P.rule][key];
#pragma unroll
        for (int k = 0; k < DPAD; ++k)
          if (k < P.d) th[k] += s_U[j * P.d + k] * z;
      }
    }
    // RawBuild (reference src/joint_posterior.jl:183-188): the result keeps the UNCONSTRAINED node (the grid's `cache`)
    if (P.raw && blockIdx.y == 0) {
#pragma unroll
      for (int k = 0; k < DPAD; ++k)
        if (k < P.d) P.theta[(size_t)k * P.M + m] = th[k];
    }
    const double lj = jp_construct<DPAD>(th, P.d, s_code);
    if (blockIdx.y == 0) {
      if (!P.raw) {
#pragma unroll
        for (int k = 0; k < DPAD; ++k)
          if (k < P.d) P.theta[(size_t)k * P.M + m] = th[k];
      }
      ljp = lj + F::template prior<DPAD>(th, P.d, P.N, P.hyper);
      P.lj_prior[m] = ljp;
    }
  }
  // ---- stage 3: stream this split's observation records through shared memory
  const long long n_begin = (long long)blockIdx.y * P.obs_per_split;
  const long long n_end = min(P.N, n_begin + P.obs_per_split);
  const int tile_obs = JP_FIT_TILE_DOUBLES / P.ncols;
  // Summation: the records of one tile are added in order, the tile sums go into a Kahan-compensated total.  A plain
  // running sum over N = 1e7 observations (what a user's Julia loop does) carries ~1e-6 of rounding error in a log-likelihood
  // of magnitude 5e6; this keeps the FP64 path at ~1e-10 there.  Up to one tile (a few hundred records) the order is the
  // plain loop's.
  double acc = 0.0, comp = 0.0;
  for (long long base = n_begin; base < n_end; base += tile_obs) {
    const int cnt = (int)min((long long)tile_obs, n_end - base);
    const double* src = P.obs + (size_t)base * P.ncols;
    const int nd = cnt * P.ncols;
    __syncthreads();
    for (int i = threadIdx.x; i < nd; i += JP_FIT_THREADS) tile[i] = __ldg(src + i);
    __syncthreads();
    if (live) {
      double ts = 0.0;
      for (int n = 0; n < cnt; ++n) ts += F::template obs<DPAD>(th, P.d, tile + n * P.ncols, base + n, P.hyper);
      const double y = ts - comp, t = acc + y;
      comp = (t - acc) - y;
      acc = t;
    }
  }
  if (live) P.part[(size_t)blockIdx.y * P.M + m] = acc;
}

// ld = lj + prior + sum_s part[s] + neg_min ; a = ld + |z|^2/2
__global__ void jp_fit_finish_kernel(long long M, long long m0, int splits, const double* __restrict__ part,
                                     const double* __restrict__ lj_prior, const double* __restrict__ hzz,
                                     double neg_min, double* __restrict__ logdens, double* __restrict__ a) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double s = 0;
  for (int k = 0; k < splits; ++k) s += part[(size_t)k * M + m];
  double ld = (s + lj_prior[m]) + neg_min;
  logdens[m] = ld;
  a[m] = ld + hzz[m0 + m];
}

// Stage 4 reductions over the (L2-resident) node arrays: multi-block, coalesced, second stage by the last block to
// arrive (jp_last_block) in block order -> bitwise reproducible.
__global__ void __launch_bounds__(256) jp_reduce_max_kernel(const double* __restrict__ a, long long M, double* __restrict__ bpart,
                                                            unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double sm[33];
  const long long per = (M + gridDim.x - 1) / gridDim.x, b0 = (long long)blockIdx.x * per, b1 = min(M, b0 + per);
  double v = -INFINITY;
  for (long long i = b0 + threadIdx.x; i < b1; i += 256) v = fmax(v, a[i]);
  v = jp_block_max(v, sm);
  if (threadIdx.x == 0) bpart[blockIdx.x] = v;
  if (!jp_last_block(counter, gridDim.x)) return;
  v = -INFINITY;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += 256) v = fmax(v, __ldcg(bpart + b));
  v = jp_block_max(v, sm);
  if (threadIdx.x == 0) out[0] = v;
}
// e = w exp(a - max) and its sum: thread t of a block adds its elements in ascending order, the block combines by the
// fixed tree of jp_block_sum, the last block adds the block sums in block order
__global__ void __launch_bounds__(256) jp_expw_sum_kernel(long long M, long long m0, const double* __restrict__ a,
                                                          const double* __restrict__ w, const double* __restrict__ gmax,
                                                          double* __restrict__ e, double* __restrict__ bpart,
                                                          unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double sm[33];
  const long long per = (M + gridDim.x - 1) / gridDim.x, b0 = (long long)blockIdx.x * per, b1 = min(M, b0 + per);
  const double mx = gmax[0];
  double s = 0;
  for (long long i = b0 + threadIdx.x; i < b1; i += 256) {
    const double v = w[m0 + i] * exp(a[i] - mx);
    e[i] = v;
    s += v;
  }
  s = jp_block_sum(s, sm);
  if (threadIdx.x == 0) bpart[blockIdx.x] = s;
  if (!jp_last_block(counter, gridDim.x)) return;
  double t = 0;      // thread t takes blocks t, t + 256, .. in ascending order, then the fixed tree
  for (int b = threadIdx.x; b < (int)gridDim.x; b += 256) t += __ldcg(bpart + b);
  t = jp_block_sum(t, sm);
  if (threadIdx.x == 0) out[0] = t;
}
__global__ void jp_scale_kernel(long long M, double* __restrict__ e, const double* __restrict__ gsum) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) e[m] = e[m] / gsum[0];
}

// density_m = e_m exp(m_rank - M) / S with M = max_r m_r and S = sum_r s_r exp(m_r - M) from the gathered (m_r, s_r)
__global__ void jp_scale_gathered_kernel(long long M, double* __restrict__ e, const double* __restrict__ g, int world, int rank) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double mx = -INFINITY;
  for (int r = 0; r < world; ++r) mx = fmax(mx, g[2 * r]);
  double S = 0;
  for (int r = 0; r < world; ++r) S += (g[2 * r + 1] == 0.0) ? 0.0 : g[2 * r + 1] * exp(g[2 * r] - mx);   // empty shards: (-inf, 0)
  e[m] = e[m] * exp(g[2 * rank] - mx) / S;
}

// Stage 4 in ONE cooperative launch (replaces finish + max + sum + scale = four launches of a few microseconds of work each,
// whose launch gaps were a sixth of the cfg3 step): every block owns a contiguous slice of the node block and keeps it across
// the phases (L1 / L2 resident), grid-wide barriers separate them:
//   A  per node: log-density from the partials of the log-density kernel (the deferred "finish"), a = ld + |z|^2 / 2, block max
//   B  global max (every block reduces the block maxima itself, same order -> same bits), e = w exp(a - max), block sums
//   C  (one GPU) total in block order, density = e / total.  Several GPUs stop after B with (max, sum) for the all_gather.
// Thread t of a block adds its elements in ascending order, blocks combine by the fixed tree of jp_block_sum, the block sums
// are added in block order: bitwise reproducible for a given node count.
#define JP_S4_THREADS 256
struct Stage4Params {
  long long M, m0;
  JpFinish fin;
  const double* hzz;
  const double* w;
  double* logdens;
  double* a;
  double* e;          // density
  double* bmax;       // [gridDim.x]
  double* bsum;       // [gridDim.x]
  double* stats;      // [2]: max, sum
  int normalise;
  // single-GPU fits also leave the per-block, per-coordinate (sum w theta, sum w theta^2, min, max) for stage 5
  const double* theta;   // [d][M]
  int d;
  double* cmom;          // [gridDim.x][d][4] or null
};
// per-coordinate block reduction of two values per thread: warp shuffles, one barrier, warps combined in warp order
#define JP_S4_WARPS (JP_S4_THREADS / 32)
__global__ void __launch_bounds__(JP_S4_THREADS) jp_stage4_kernel(const Stage4Params P) {
  __shared__ double sm[33];
  cg::grid_group grid = cg::this_grid();
  const long long per = (P.M + gridDim.x - 1) / gridDim.x, b0 = (long long)blockIdx.x * per, b1 = min(P.M, b0 + per);
  // ---- A
  double mx = -INFINITY;
  for (long long m = b0 + threadIdx.x; m < b1; m += JP_S4_THREADS) {
    double ld;
    if (P.fin.path == JP_PATH_TC) {
      const long long gm = P.m0 + m, j = ((gm + 1) >> 1) - P.fin.j_lo;
      const double sgn = ((gm & 1) || gm == 0) ? 1.0 : -1.0;
      double s = 0;
      for (int c = 0; c < P.fin.chunks; ++c) {
        const double2 eo = *reinterpret_cast<const double2*>(P.fin.tc_part + ((size_t)c * P.fin.P + j) * 2);
        s += eo.x + sgn * eo.y;
      }
      ld = (P.fin.quad[m] - s) + P.fin.neg_min;
    } else if (P.fin.path == JP_PATH_FP64) {
      double s = 0;
      for (int k = 0; k < P.fin.splits; ++k) s += P.fin.part[(size_t)k * P.M + m];
      ld = (s + P.fin.lj_prior[m]) + P.fin.neg_min;
    } else {
      ld = P.logdens[m];      // already finished by the path's own kernel
    }
    const double a = ld + P.hzz[P.m0 + m];
    if (P.fin.path != 0) P.logdens[m] = ld;
    P.a[m] = a;
    mx = fmax(mx, a);
  }
  mx = jp_block_max(mx, sm);
  if (threadIdx.x == 0) P.bmax[blockIdx.x] = mx;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  grid.sync();
  // ---- B
  double g = -INFINITY;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += JP_S4_THREADS) g = fmax(g, __ldcg(P.bmax + b));
  g = jp_block_max(g, sm);
  double s = 0;
  for (long long m = b0 + threadIdx.x; m < b1; m += JP_S4_THREADS) {
    const double v = P.w[P.m0 + m] * exp(P.a[m] - g);
    P.e[m] = v;
    s += v;
  }
  s = jp_block_sum(s, sm);
  if (threadIdx.x == 0) P.bsum[blockIdx.x] = s;
  grid.sync();
  // ---- C: every block adds the block sums itself, all threads loading (thread t takes blocks t, t + 256, .. in ascending
  // order, then the fixed tree of jp_block_sum: the same bits in every block; one thread walking the few hundred L2-resident
  // partials alone cost ~35 us of dependent load latency per launch)
  double tot = 0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += JP_S4_THREADS) tot += __ldcg(P.bsum + b);
  tot = jp_block_sum(tot, sm);
  if (blockIdx.x == 0 && threadIdx.x == 0) { P.stats[0] = g; P.stats[1] = tot; }
  if (!P.normalise) return;
  for (long long m = b0 + threadIdx.x; m < b1; m += JP_S4_THREADS) P.e[m] = P.e[m] / tot;
  if (!P.cmom) return;
  // Per-coordinate (sum w theta, sum w theta^2, min theta, max theta) over this block's slice, for stage 5: WARP w takes
  // coordinates w, w + 8, ..; its lanes walk the slice of the coordinate's column (coalesced; the weights were just written by
  // this block), then one shuffle tree per quantity -- the warp's result is the block's, no shared memory, no barrier per
  // coordinate.  (One butterfly per thread-owned node and coordinate cost 8.5 M of this kernel's 13 M warp instructions.)
  __syncthreads();
  for (int k = wid; k < P.d; k += JP_S4_WARPS) {
    const double* th = P.theta + (size_t)k * P.M;
    double s1 = 0, s2 = 0, mn = INFINITY, mxk = -INFINITY;
    for (long long m = b0 + lane; m < b1; m += 32) {
      const double v = th[m], wv = P.e[m];
      s1 += wv * v;
      s2 += wv * (v * v);
      mn = fmin(mn, v);
      mxk = fmax(mxk, v);
    }
    s1 = jp_warp_sum(s1);
    s2 = jp_warp_sum(s2);
    mn = jp_warp_min(mn);
    mxk = jp_warp_max(mxk);
    if (lane == 0) {
      double* o = P.cmom + ((size_t)blockIdx.x * P.d + k) * 4;
      o[0] = s1; o[1] = s2; o[2] = mn; o[3] = mxk;
    }
  }
}

// RawBuild consumers (reference src/marginal_posterior.jl:68-77,106-115: update!(Theta) on every column of grid.cache before
// f is evaluated): constrained coordinates `coords` (null: all d) of every node from the unconstrained cache x[d][M].
template <int DPAD>
__global__ void __launch_bounds__(128) jp_construct_kernel(int d, long long M, const int* __restrict__ tcode, const double* __restrict__ x,
                                                           int K, const int* __restrict__ coords, double* __restrict__ out) {
  __shared__ int s_code[JP_MAX_D];
  for (int i = threadIdx.x; i < d; i += blockDim.x) s_code[i] = tcode[i];
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double th[DPAD];
#pragma unroll
  for (int k = 0; k < DPAD; ++k) th[k] = (k < d) ? x[(size_t)k * M + m] : 0.0;
  jp_construct<DPAD>(th, d, s_code);
  for (int j = 0; j < K; ++j) {
    const int c = coords ? coords[j] : j;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < DPAD; ++k)
      if (k == c) v = th[k];
    out[(size_t)j * M + m] = v;
  }
}

int jp_construct_columns(jp_posterior* post, int K, const int* d_coords, double* d_out) {
  jp_ctx* ctx = post->ctx;
  const unsigned gb = (unsigned)((post->M + 127) / 128);
  if (post->d <= 16)
    jp_construct_kernel<16><<<gb, 128, 0, ctx->stream>>>(post->d, post->M, post->d_tcode, post->d_theta, K, d_coords, d_out);
  else
    jp_construct_kernel<JP_MAX_D><<<gb, 128, 0, ctx->stream>>>(post->d, post->M, post->d_tcode, post->d_theta, K, d_coords, d_out);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}

// ------------------------------------------------------------------------------------ registry
static JpFamilyEntry g_families[16];
static int g_nfamilies = 0;
void jp_register_family(const JpFamilyEntry& e) {
  if (g_nfamilies < 16) g_families[g_nfamilies++] = e;
}
const JpFamilyEntry* jp_find_family(int id) {
  for (int i = 0; i < g_nfamilies; ++i)
    if (g_families[i].id == id) return &g_families[i];
  return nullptr;
}

template <class F, int DPAD>
static int launch_nodes(jp_posterior* post, const JpFitLaunchParams& lp) {
  dim3 grid((unsigned)((lp.M + JP_FIT_THREADS - 1) / JP_FIT_THREADS), lp.splits);
  size_t smem = (size_t)(JP_FIT_TILE_DOUBLES + lp.d * lp.p + lp.d) * sizeof(double) + (size_t)lp.d * sizeof(int);
  if (smem > 48 * 1024)
    JP_CUDA(cudaFuncSetAttribute(jp_fit_nodes_kernel<F, DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEventRecord(post->ctx->ev_k0, post->ctx->stream);
  jp_fit_nodes_kernel<F, DPAD><<<grid, JP_FIT_THREADS, smem, post->ctx->stream>>>(lp);
  cudaEventRecord(post->ctx->ev_k1, post->ctx->stream);
  post->ctx->ev_valid = true;
  JP_CHECK_LAUNCH(post->ctx);
  return JP_OK;
}
template <class F>
static int launch_family(jp_posterior* post, const JpFitLaunchParams& lp) {
  int d = lp.d;
  if (d <= 4) return launch_nodes<F, 4>(post, lp);
  if (d <= 8) return launch_nodes<F, 8>(post, lp);
  if (d <= 12) return launch_nodes<F, 12>(post, lp);
  if (d <= 16) return launch_nodes<F, 16>(post, lp);
  if (d <= 20) return launch_nodes<F, 20>(post, lp);
  if (d <= 24) return launch_nodes<F, 24>(post, lp);
  if (d <= 32) return launch_nodes<F, 32>(post, lp);
  return launch_nodes<F, JP_MAX_D>(post, lp);
}
// families whose shape_ok admits a single dimension need a single padded width (every (family, width) pair is a kernel to compile)
template <>
int launch_family<FamBinomialMixture>(jp_posterior* post, const JpFitLaunchParams& lp) {      // d = 3
  return launch_nodes<FamBinomialMixture, 4>(post, lp);
}
template <>
int launch_family<FamAnova2>(jp_posterior* post, const JpFitLaunchParams& lp) {               // d = 5
  return launch_nodes<FamAnova2, 8>(post, lp);
}
#define JP_REGISTER_FAMILY(F)                                                                  \
  static struct Reg##F {                                                                        \
    Reg##F() { jp_register_family(JpFamilyEntry{F::kId, F::kName, &launch_family<F>, &F::shape_ok}); } \
  } g_reg_##F;

JP_REGISTER_FAMILY(FamBinomialMixture)
JP_REGISTER_FAMILY(FamLogistic)
JP_REGISTER_FAMILY(FamPoisson)
JP_REGISTER_FAMILY(FamHierNormal)
JP_REGISTER_FAMILY(FamNormalLinear)
JP_REGISTER_FAMILY(FamMultinomial)
JP_REGISTER_FAMILY(FamMvnCov)
JP_REGISTER_FAMILY(FamAnova2)

// ------------------------------------------------------------------------------------ host side
int jp_upload_fit_consts(jp_posterior* post, const jp_fit_args* args) {
  jp_ctx* ctx = post->ctx;
  if (!ctx->fit_nodes_uploaded) {
    for (int r = 0; r < 2; ++r) {
      JpRule R = jp_get_rule(r);
      double nodes[JP_RULE_NMAX] = {0};
      for (int j = 0; j < R.nmax; ++j) nodes[j] = R.nodes[j];
      JP_CUDA(cudaMemcpyToSymbol(c_fit_nodes, nodes, sizeof nodes, sizeof(double) * JP_RULE_NMAX * r));
    }
    ctx->fit_nodes_uploaded = true;
  }
  // stage through pinned memory so the copies are truly asynchronous
  double* hp = ctx->h_pinned;
  int d = args->d, p = args->p;
  JP_CUDA(jp_pinned_acquire(ctx));               // the staging area is reused across calls
  for (int i = 0; i < d; ++i) hp[i] = args->h_mu_hat[i];
  for (int i = 0; i < d * p; ++i) hp[d + i] = args->h_U[i];
  int* hc = reinterpret_cast<int*>(hp + d + d * p);
  for (int i = 0; i < d; ++i) hc[i] = args->h_transform[i];
  post->tcode_host.assign(args->h_transform, args->h_transform + d);
  // one copy: the device block of jp_posterior_create has the staging area's layout (mu | U | codes)
  JP_CUDA(cudaMemcpyAsync(post->d_mu, hp, sizeof(double) * (d + d * p) + sizeof(int) * d, cudaMemcpyHostToDevice, ctx->stream));
  JP_CUDA(jp_pinned_publish(ctx));
  return JP_OK;
}

int jp_check_transform_codes(const char* who, const int* code, int d) {
  for (int k = 0; k < d; ++k) {
    int kind = JP_T_KIND(code[k]);
    JP_REQUIRE(kind >= 0 && kind <= JP_T_COVMAT && (kind >= JP_T_NONCENTRED || code[k] == kind),
               "%s: unknown transform code %d at coordinate %d", who, code[k], k);
    if (kind == JP_T_COVMAT) {
      const int first = JP_T_LOC(code[k]), len = JP_T_SCALE(code[k]);
      int p = 0;
      while ((p + 1) * (p + 2) / 2 <= len) ++p;
      JP_REQUIRE(len >= 1 && len <= JP_COVMAT_MAX_LEN && p * (p + 1) / 2 == len && first <= k && k < first + len && first + len <= d &&
                 (code[k] >> 24) == 0, "%s: covariance-matrix coordinate %d: block [%d, %d) is not the lower triangle of a p x p matrix, p <= 10",
                 who, k, first, first + len);
      JP_REQUIRE(code[first] == code[k], "%s: covariance-matrix block [%d, %d) does not carry one code word (coordinate %d)", who, first,
                 first + len, k);
    }
    if (kind == JP_T_SIMPLEX) {
      const int first = JP_T_LOC(code[k]), len = JP_T_SCALE(code[k]);
      JP_REQUIRE(len >= 1 && first <= k && k < first + len && first + len <= d && (code[k] >> 24) == 0,
                 "%s: simplex coordinate %d lies outside its block [%d, %d)", who, k, first, first + len);
      JP_REQUIRE(code[first] == code[k], "%s: simplex block [%d, %d) does not carry one code word (coordinate %d)", who, first,
                 first + len, k);
    }
    if (kind == JP_T_NONCENTRED)
      JP_REQUIRE(JP_T_LOC(code[k]) < k && JP_T_SCALE(code[k]) < k && (code[k] >> 24) == 0,
                 "%s: non-centred coordinate %d must refer to earlier coordinates (loc %d, scale %d)", who, k,
                 JP_T_LOC(code[k]), JP_T_SCALE(code[k]));
  }
  return JP_OK;
}

int jp_fit_check_args(const jp_posterior* post, const jp_fit_args* args) {
  JP_REQUIRE(post && args, "jp_fit: null argument");
  JP_REQUIRE(args->d == post->d && args->p == post->p, "jp_fit: (d,p)=(%d,%d) differs from the posterior's (%d,%d)",
             args->d, args->p, post->d, post->p);
  JP_REQUIRE(args->h_transform && args->h_mu_hat && args->h_U, "jp_fit: null host array");
  JP_TRY(jp_check_transform_codes("jp_fit", args->h_transform, args->d));
  const_cast<jp_posterior*>(post)->raw = args->raw != 0;
  const_cast<jp_posterior*>(post)->fit_gen += 1;
  return JP_OK;
}

int jp_fit_fp64_launch(jp_posterior* post, const jp_fit_args* args, bool finish) {
  jp_ctx* ctx = post->ctx;
  const jp_data* data = post->data;
  const JpFamilyEntry* fam = jp_find_family(data->family);
  JP_REQUIRE(fam != nullptr, "jp_fit: family %d is not registered", data->family);
  JP_REQUIRE(fam->shape_ok(args->d, data->ncols, data->N), "jp_fit: family %s does not accept d=%d ncols=%d N=%lld",
             fam->name, args->d, data->ncols, data->N);
  JP_REQUIRE(data->ncols <= JP_FIT_TILE_DOUBLES / 2, "jp_fit: %d columns per observation is too many", data->ncols);
  JP_TRY(jp_upload_fit_consts(post, args));
  JpFitLaunchParams lp;
  lp.d = args->d; lp.p = args->p; lp.ncols = data->ncols; lp.rule = post->grid->rule;
  lp.N = data->N; lp.M = post->M; lp.m0 = post->m0; lp.M_grid = post->grid->M;
  lp.neg_min = args->neg_min;
  lp.idx = post->grid->d_idx; lp.mu = post->d_mu; lp.U = post->d_U; lp.tcode = post->d_tcode;
  lp.obs = data->d_obs;
  for (int i = 0; i < JP_MAX_HYPER; ++i) lp.hyper[i] = data->hyper[i];
  lp.theta = post->d_theta;
  // observation splits: aim for >= 4 blocks per SM, bounded by the partial buffer and by tiles
  long long node_blocks = (post->M + JP_FIT_THREADS - 1) / JP_FIT_THREADS;
  int tile_obs = JP_FIT_TILE_DOUBLES / data->ncols;
  long long max_by_tiles = std::max(1LL, data->N / (4LL * tile_obs));
  long long want = (4LL * ctx->sm_count + node_blocks - 1) / node_blocks;
  long long cap = std::max(1LL, (long long)JP_POST_PART_SPLITS);
  int splits = (int)std::max(1LL, std::min(std::min(want, max_by_tiles), cap));
  lp.splits = splits;
  lp.obs_per_split = (data->N + splits - 1) / splits;
  // round the split length to whole tiles so tiles never straddle a split boundary unevenly
  lp.obs_per_split = ((lp.obs_per_split + tile_obs - 1) / tile_obs) * tile_obs;
  lp.xpts = nullptr;
  lp.raw = post->raw ? 1 : 0;
  lp.part = post->d_part;
  lp.lj_prior = post->d_part + (size_t)JP_POST_PART_SPLITS * post->M;
  JP_TRY(fam->launch(post, lp));
  post->path_used = JP_PATH_FP64;
  post->fin = JpFinish();
  post->fin.path = JP_PATH_FP64; post->fin.splits = splits; post->fin.part = lp.part; post->fin.lj_prior = lp.lj_prior;
  post->fin.neg_min = args->neg_min;
  if (finish) {
    unsigned gb = (unsigned)((post->M + 255) / 256);
    jp_fit_finish_kernel<<<gb, 256, 0, ctx->stream>>>(post->M, post->m0, splits, lp.part, lp.lj_prior,
                                                       post->grid->d_hzz, args->neg_min, post->d_logdens, post->d_a);
    JP_CHECK_LAUNCH(ctx);
    post->fin.path = 0;
  }
  return JP_OK;
}

__global__ void jp_points_finish_kernel(long long K, int splits, const double* __restrict__ part,
                                        const double* __restrict__ lj_prior, double* __restrict__ out) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= K) return;
  double s = 0;
  for (int k = 0; k < splits; ++k) s += part[(size_t)k * K + m];
  out[m] = s + lj_prior[m];
}

// stage 4 after a log-density launch made with finish = false; d_stats[2] receives (max, sum)
static int jp_stage4_launch(jp_posterior* post, bool normalise, double* d_stats) {
  jp_ctx* ctx = post->ctx;
  Stage4Params P;
  P.M = post->M; P.m0 = post->m0; P.fin = post->fin;
  P.hzz = post->grid->d_hzz; P.w = post->grid->d_w;
  P.logdens = post->d_logdens; P.a = post->d_a; P.e = post->d_density;
  // one node per thread up to six blocks per SM (the phases are latency chains: parallelism, not bytes, is what they need);
  // co-resident by a wide margin (256 threads, 8 KB of shared memory)
  static int per_sm = 0;      // co-resident blocks per SM of this kernel (a property of the build, not of the device instance)
  if (per_sm == 0) {
    JP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jp_stage4_kernel, JP_S4_THREADS, 0));
    per_sm = std::max(1, std::min(per_sm, 6));
  }
  const int nb = (int)std::max<long long>(1, std::min<long long>((long long)per_sm * ctx->sm_count, (post->M + JP_S4_THREADS - 1) / JP_S4_THREADS));
  JP_REQUIRE(2 * nb <= JP_BPART_DOUBLES, "stage 4: %d blocks exceed the partial buffer", nb);
  P.bmax = ctx->d_bpart; P.bsum = ctx->d_bpart + nb;
  P.stats = d_stats;
  P.normalise = normalise ? 1 : 0;
  P.theta = post->d_theta; P.d = post->d; P.cmom = nullptr;
  post->cmom_valid = false;
  if (normalise && post->M >= 2 && !post->raw) {      // (a RawBuild's theta array holds unconstrained coordinates)
    if (nb > post->cmom_cap) {
      jp_dfree(ctx, post->d_cmom);
      post->d_cmom = nullptr;
      post->cmom_cap = 0;
      JP_CUDA(jp_dmalloc(ctx, &post->d_cmom, (size_t)nb * post->d * 4 * 8));
      post->cmom_cap = nb;
    }
    P.cmom = post->d_cmom;
    post->cmom_blocks = nb;
    post->cmom_valid = true;
  }
  void* kargs[] = {(void*)&P};
  JP_CUDA(cudaLaunchCooperativeKernel((const void*)jp_stage4_kernel, dim3(nb), dim3(JP_S4_THREADS), kargs, 0, ctx->stream));
  JP_CHECK_LAUNCH(ctx);
  JP_MARK(ctx, "fit:stage4");
  post->fin.path = 0;
  return JP_OK;
}

// stages 2-3 by the path the arguments select, the finish deferred to stage 4
static int jp_fit_launch_path(jp_posterior* post, const jp_fit_args* args, bool finish) {
  post->sorted_valid = false;      // the sorted arrays / marginals of an earlier fit no longer describe this posterior
  post->K_last = 0;
  post->cmom_valid = false;
  // AUTO: GLM families take the tensor-core path when its a-priori error bounds hold for this
  // (data, U, grid); otherwise, and for every other family, the FP64 plugin kernel runs.
  int path = args->path;
  if (path == JP_PATH_AUTO) path = jp_fit_tc_supported(post, args) ? JP_PATH_TC : JP_PATH_FP64;
  if (path == JP_PATH_TC) {
    int st = jp_fit_tc_launch(post, args, finish);
    if (st == JP_ERR_UNSUPPORTED && args->path == JP_PATH_AUTO) path = JP_PATH_FP64;
    else JP_TRY(st);
  }
  if (path == JP_PATH_FP64) JP_TRY(jp_fit_fp64_launch(post, args, finish));
  return JP_OK;
}

extern "C" {

int jp_log_density_points(jp_ctx* ctx, const jp_data* data, int d, const int* h_transform, long long K,
                          const double* h_x, double* h_ld) {
  JP_REQUIRE(ctx && data && h_transform && h_x && h_ld, "jp_log_density_points: null argument");
  JP_REQUIRE(d >= 1 && d <= JP_MAX_D && K >= 1, "jp_log_density_points: bad shape d=%d K=%lld", d, K);
  const JpFamilyEntry* fam = jp_find_family(data->family);
  JP_REQUIRE(fam != nullptr, "jp_log_density_points: family %d is not registered", data->family);
  JP_REQUIRE(fam->shape_ok(d, data->ncols, data->N), "jp_log_density_points: family %s does not accept d=%d ncols=%d N=%lld",
             fam->name, d, data->ncols, data->N);
  JP_TRY(jp_check_transform_codes("jp_log_density_points", h_transform, d));
  JP_CUDA(cudaSetDevice(ctx->device));
  const int splits = 1;
  double *d_x = nullptr, *d_theta = nullptr, *d_part = nullptr, *d_out = nullptr;
  int* d_code = nullptr;
  jp_posterior tmp;   // only ctx is used by the launcher
  tmp.ctx = ctx;
  int st = JP_OK;
  cudaError_t e = cudaSuccess;
  // Small batches (the mode finder's finite-difference stencils: a few hundred points, called tens of times per fit)
  // live in the context's persistent scratch and are staged through its pinned buffer: no allocation, no pageable copy.
  const size_t code_dbl = ((size_t)d + 1) / 2, need_dev = (size_t)K * (2 * d + splits + 2) + code_dbl;
  const size_t need_pin = (size_t)K * (d + 1) + code_dbl;
  const bool small = need_dev <= JP_SCRATCH_DOUBLES && need_pin <= JP_PINNED_DOUBLES;
  double* hp = ctx->h_pinned;
  if (small) {
    // Zero-copy: with unified addressing the pinned buffer is device-accessible at its host address, so the kernels read
    // the (few KB of) points and codes straight from it and write the results back into it -- no copy calls, one
    // launch pair and one synchronisation per evaluation (a Newton iteration of jp_mode is one such call).
    e = jp_pinned_acquire(ctx);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);    // zero-copy: no kernel of an earlier call may still read / write it
    std::memcpy(hp, h_x, (size_t)K * d * 8);
    std::memcpy(hp + (size_t)K * d, h_transform, (size_t)d * 4);
    d_x = hp;
    d_code = reinterpret_cast<int*>(hp + (size_t)K * d);
    d_out = hp + (size_t)K * d + code_dbl;
    d_theta = ctx->d_scratch;
    d_part = d_theta + (size_t)K * d;
  } else {
    e = jp_dmalloc(ctx, &d_x, (size_t)K * d * 8);
    if (e == cudaSuccess) e = jp_dmalloc(ctx, &d_theta, (size_t)K * d * 8);
    if (e == cudaSuccess) e = jp_dmalloc(ctx, &d_part, (size_t)K * (splits + 1) * 8);
    if (e == cudaSuccess) e = jp_dmalloc(ctx, &d_out, (size_t)K * 8);
    if (e == cudaSuccess) e = jp_dmalloc(ctx, &d_code, (size_t)d * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, h_x, (size_t)K * d * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_code, h_transform, (size_t)d * 4, cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess) {
    JpFitLaunchParams lp;
    lp.d = d; lp.p = 0; lp.ncols = data->ncols; lp.rule = 0;
    lp.N = data->N; lp.M = K; lp.m0 = 0; lp.M_grid = K; lp.neg_min = 0.0;
    lp.idx = nullptr; lp.mu = nullptr; lp.U = nullptr; lp.tcode = d_code; lp.obs = data->d_obs;
    for (int i = 0; i < JP_MAX_HYPER; ++i) lp.hyper[i] = data->hyper[i];
    lp.theta = d_theta; lp.splits = splits; lp.obs_per_split = data->N;
    lp.part = d_part; lp.lj_prior = d_part + (size_t)splits * K; lp.xpts = d_x; lp.raw = 0;
    st = fam->launch(&tmp, lp);
    if (st == JP_OK) {
      jp_points_finish_kernel<<<(unsigned)((K + 255) / 256), 256, 0, ctx->stream>>>(K, splits, lp.part, lp.lj_prior, d_out);
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (st == JP_OK && e == cudaSuccess && !small)
      e = cudaMemcpyAsync(h_ld, d_out, (size_t)K * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (st == JP_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (st == JP_OK && e == cudaSuccess && small) std::memcpy(h_ld, d_out, (size_t)K * 8);
  }
  if (!small) {
    jp_dfree(ctx, d_x); jp_dfree(ctx, d_theta); jp_dfree(ctx, d_part); jp_dfree(ctx, d_out); jp_dfree(ctx, d_code);
  }
  if (st != JP_OK) return st;
  if (e != cudaSuccess) {
    jp_set_error("jp_log_density_points: %s", cudaGetErrorString(e));
    return JP_ERR_CUDA;
  }
  return JP_OK;
}

int jp_fit_local(jp_posterior* post, const jp_fit_args* args, double* d_local_max) {
  JP_REQUIRE(post, "jp_fit_local: null posterior");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  JP_REQUIRE(d_local_max, "jp_fit_local: null output");
  JP_TRY(jp_fit_launch_path(post, args, true));
  jp_reduce_max_kernel<<<jp_red_blocks(post->M), 256, 0, post->ctx->stream>>>(post->d_a, post->M, post->ctx->d_bpart,
                                                                              post->ctx->d_counters, d_local_max);
  JP_CHECK_LAUNCH(post->ctx);
  return JP_OK;
}

int jp_fit_local_sum(jp_posterior* post, const double* d_global_max, double* d_local_sum) {
  JP_REQUIRE(post && d_global_max && d_local_sum, "jp_fit_local_sum: null argument");
  JP_ENTER_CTX(post->ctx);
  jp_expw_sum_kernel<<<jp_red_blocks(post->M), 256, 0, post->ctx->stream>>>(post->M, post->m0, post->d_a, post->grid->d_w,
                                                                            d_global_max, post->d_density, post->ctx->d_bpart,
                                                                            post->ctx->d_counters, d_local_sum);
  JP_CHECK_LAUNCH(post->ctx);
  return JP_OK;
}

int jp_fit_normalise(jp_posterior* post, const double* d_global_sum) {
  JP_REQUIRE(post && d_global_sum, "jp_fit_normalise: null argument");
  JP_ENTER_CTX(post->ctx);
  unsigned gb = (unsigned)((post->M + 255) / 256);
  jp_scale_kernel<<<gb, 256, 0, post->ctx->stream>>>(post->M, post->d_density, d_global_sum);
  JP_CHECK_LAUNCH(post->ctx);
  return JP_OK;
}

int jp_fit_local_stats(jp_posterior* post, const jp_fit_args* args, double* d_stats) {
  JP_REQUIRE(post && d_stats, "jp_fit_local_stats: null argument");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  JP_TRY(jp_fit_launch_path(post, args, false));
  return jp_stage4_launch(post, false, d_stats);      // finish + local max + sum relative to it: one launch
}

// ---- node-sharded fit with the O(N) prep of the tensor-core path sharded by OBSERVATION (csrc/jp_glm_tc.cu)
int jp_fit_prep_len(int d) { return jp_fit_tc_prep_len(d); }

int jp_fit_prep_local(jp_posterior* post, const jp_fit_args* args, int rank, int world, double* d_out) {
  JP_REQUIRE(post, "jp_fit_prep_local: null posterior");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  JP_REQUIRE(args->path != JP_PATH_FP64, "jp_fit_prep_local: the sharded prep belongs to the tensor-core path");
  return jp_fit_tc_prep_local(post, args, rank, world, d_out);
}

int jp_fit_prep_gathered(jp_posterior* post, const jp_fit_args* args, const double* d_gathered, int world, int rank,
                         int* n_rows) {
  JP_REQUIRE(post, "jp_fit_prep_gathered: null posterior");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  return jp_fit_tc_prep_gathered(post, args, d_gathered, world, rank, n_rows);
}

int jp_fit_coef_slab(jp_posterior* post, int n_rows, void** d_local, void** d_all, long long* count) {
  JP_REQUIRE(post && d_local && d_all && count, "jp_fit_coef_slab: null argument");
  JP_ENTER_CTX(post->ctx);
  float *pl = nullptr, *pa = nullptr;
  JP_TRY(jp_fit_tc_coef_slab(post, n_rows, &pl, &pa, count));
  *d_local = pl;
  *d_all = pa;
  return JP_OK;
}

int jp_fit_local_stats_prepared(jp_posterior* post, const jp_fit_args* args, double* d_stats) {
  JP_REQUIRE(post, "jp_fit_local_stats_prepared: null posterior");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  JP_REQUIRE(d_stats, "jp_fit_local_stats_prepared: null output");
  JP_TRY(jp_fit_tc_run_prepared(post, args, false));
  return jp_stage4_launch(post, false, d_stats);
}

int jp_fit_normalise_gathered(jp_posterior* post, const double* d_gathered, int world, int rank) {
  JP_REQUIRE(post && d_gathered && world >= 1 && rank >= 0 && rank < world, "jp_fit_normalise_gathered: bad argument");
  JP_ENTER_CTX(post->ctx);
  if (post->M == 0) return JP_OK;
  unsigned gb = (unsigned)((post->M + 255) / 256);
  jp_scale_gathered_kernel<<<gb, 256, 0, post->ctx->stream>>>(post->M, post->d_density, d_gathered, world, rank);
  JP_CHECK_LAUNCH(post->ctx);
  return JP_OK;
}

// The node-sharded fit with the exchanges inside the library (csrc/jp_comm.cu): one asynchronous queue per rank.
int jp_fit_p2p(jp_posterior* post, const jp_fit_args* args, jp_comm* comm) {
  JP_REQUIRE(post && comm, "jp_fit_p2p: null argument");
  JP_REQUIRE(comm->ctx == post->ctx, "jp_fit_p2p: the communicator belongs to another context");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  post->sorted_valid = false;
  post->K_last = 0;
  post->cmom_valid = false;
  int path = args->path;
  if (path == JP_PATH_AUTO) path = jp_fit_tc_supported(post, args) ? JP_PATH_TC : JP_PATH_FP64;
  if (path == JP_PATH_TC) JP_TRY(jp_fit_tc_launch_dev(post, args, comm, 0));
  else JP_TRY(jp_fit_fp64_launch(post, args, false));
  if (comm->world == 1) {
    JP_TRY(jp_stage4_launch(post, true, post->d_stats));
    return getenv("JP_NO_DENSITY_PREFETCH") ? JP_OK : jp_density_prefetch(post);
  }
  JP_TRY(jp_stage4_launch(post, false, post->d_stats));      // finish + local max + sum relative to it
  const double* g = nullptr;
  JP_TRY(jp_comm_exchange(comm, JP_CH_STATS, post->d_stats, 2, &g));
  if (post->M > 0) {
    jp_scale_gathered_kernel<<<(unsigned)((post->M + 255) / 256), 256, 0, post->ctx->stream>>>(post->M, post->d_density, g, comm->world,
                                                                                              comm->rank);
    JP_CHECK_LAUNCH(post->ctx);
  }
  JP_MARK(post->ctx, "fit:normalised");
  return getenv("JP_NO_DENSITY_PREFETCH") ? JP_OK : jp_density_prefetch(post);
}

// OBSERVATION-sharded fit (the alternative to node sharding): `post` covers ALL grid nodes, its data handle holds
// only this rank's observations.  Tensor-core GLM path only (the remainder sums are additive over observations).  Every
// rank ends up with the complete, bit-identical posterior: stage 5 needs no exchange at all.
int jp_fit_p2p_obs(jp_posterior* post, const jp_fit_args* args, jp_comm* comm) {
  JP_REQUIRE(post && comm, "jp_fit_p2p_obs: null argument");
  JP_REQUIRE(comm->ctx == post->ctx, "jp_fit_p2p_obs: the communicator belongs to another context");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  JP_REQUIRE(post->m0 == 0 && post->M == post->grid->M, "jp_fit_p2p_obs: the posterior must cover all %lld grid nodes (it is the "
             "observations that are sharded)", post->grid->M);
  if (args->path == JP_PATH_FP64 || !jp_fit_tc_supported(post, args)) {
    jp_set_error("jp_fit_p2p_obs: observation sharding needs the tensor-core GLM path (family %d, path %d); shard the nodes instead "
                 "(jp_fit_p2p)", post->data->family, args->path);
    return JP_ERR_UNSUPPORTED;
  }
  post->sorted_valid = false;
  post->K_last = 0;
  post->cmom_valid = false;
  JP_TRY(jp_fit_tc_launch_dev(post, args, comm, 1));
  JP_TRY(jp_stage4_launch(post, true, post->d_stats));
  return getenv("JP_NO_DENSITY_PREFETCH") ? JP_OK : jp_density_prefetch(post);      // as jp_fit: the weights start their way to the host
}

int jp_fit_p2p_check(jp_posterior* post) {
  JP_REQUIRE(post, "jp_fit_p2p_check: null posterior");
  JP_ENTER_CTX(post->ctx);
  JP_CUDA(cudaStreamSynchronize(post->ctx->stream));
  return jp_fit_tc_verify(post);
}

int jp_fit(jp_posterior* post, const jp_fit_args* args) {
  JP_REQUIRE(post, "jp_fit: null posterior");
  JP_ENTER_CTX(post->ctx);
  JP_TRY(jp_fit_check_args(post, args));
  JP_TRY(jp_fit_launch_path(post, args, false));
  JP_TRY(jp_stage4_launch(post, true, post->d_stats));   // finish + max + sum + scale: one cooperative launch
  // `density` is a host vector in the reference's result (src/joint_posterior.jl:5,181): its download starts now, beside
  // whatever follows on the main stream
  static const bool no_prefetch = getenv("JP_NO_DENSITY_PREFETCH") != nullptr;
  return no_prefetch ? JP_OK : jp_density_prefetch(post);
}

}  // extern "C"
