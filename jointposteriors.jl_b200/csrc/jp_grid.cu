// jp_grid.cu -- STAGE 1: Smolyak sparse-grid construction on the GPU.
//
// Replaces the grid build hidden behind eval_grid!/GridVessel in SparseQuadratureGrids (call sites:
// reference src/joint_posterior.jl:177,180,183,186; cache key :157-162).  Pipeline (all kernels on
// the ctx stream, integer work except the weight products):
//   1. multi-index enumeration: thread r unranks the r-th multi-index of the index set
//        I = { i : |i|_1 <= L + d - 1, 1 <= i_k <= cap }   (ascending |i|_1, lexicographic inside a class)
//      from a table of restricted-composition counts, and evaluates its combination coefficient
//        c_i = (-1)^J C(n-1, J),  n = #{k : i_k < cap},  J = min(n, q - |i|_1)   (0 when J >= n)
//   2. exclusive scan of the per-multi-index tensor-product sizes
//   3. expansion: one thread per pre-merge point -> integer key (one master-node index per
//      dimension, nested rules share indices) + weight c_i * prod_k w^{(i_k)}_{j_k}
//   4. duplicate merge: stable LSD radix sort of a permutation on the d key bytes (jp_sort.cuh),
//      head flags, and a sequential per-segment weight sum in stable (= generation) order, which is
//      what makes the weight table bit-identical to the CPU restatement.
#include <algorithm>
#include <cmath>
#include "jp_common.cuh"
#include "jp_sort.cuh"
#include "jp_rule_tables.h"

JpRule jp_get_rule(int rule) {
  if (rule == JP_RULE_KRONROD_PATTERSON)
    return JpRule{JP_KP_LEVELS, JP_KP_NMAX, jp_kp_npts, jp_kp_znodes, &jp_kp_weights[0][0], &jp_kp_index[0][0]};
  return JpRule{JP_GK_LEVELS, JP_GK_NMAX, jp_gk_npts, jp_gk_nodes, &jp_gk_weights[0][0], &jp_gk_index[0][0]};
}

// rule tables in constant memory: [rule][...]
#define JP_RULE_LMAX 8
static_assert(JP_GK_NMAX <= JP_RULE_NMAX && JP_KP_NMAX <= JP_RULE_NMAX, "master node table too small");
static_assert(JP_GK_LEVELS <= JP_RULE_LMAX && JP_KP_LEVELS <= JP_RULE_LMAX, "level table too small");
__constant__ double c_rule_nodes[2][JP_RULE_NMAX];
__constant__ double c_rule_weights[2][JP_RULE_LMAX][JP_RULE_NMAX];   // by master index
__constant__ unsigned char c_rule_index[2][JP_RULE_LMAX][JP_RULE_NMAX];   // position in a level -> master index
__constant__ int c_rule_npts[2][JP_RULE_LMAX];

static int upload_rules(jp_ctx* ctx) {
  if (ctx->rules_uploaded) return JP_OK;
  for (int r = 0; r < 2; ++r) {
    JpRule R = jp_get_rule(r);
    double nodes[JP_RULE_NMAX] = {0};
    double weights[JP_RULE_LMAX][JP_RULE_NMAX] = {{0}};
    int npts[JP_RULE_LMAX] = {0};
    unsigned char index[JP_RULE_LMAX][JP_RULE_NMAX] = {{0}};
    for (int j = 0; j < R.nmax; ++j) nodes[j] = R.nodes[j];
    for (int l = 0; l < R.levels; ++l) {
      npts[l] = R.npts[l];
      for (int j = 0; j < R.nmax; ++j) {
        weights[l][j] = R.weights[(size_t)l * R.nmax + j];
        index[l][j] = R.index[(size_t)l * R.nmax + j];
      }
    }
    JP_CUDA(cudaMemcpyToSymbol(c_rule_index, index, sizeof index, sizeof(unsigned char) * JP_RULE_LMAX * JP_RULE_NMAX * r));
    JP_CUDA(cudaMemcpyToSymbol(c_rule_nodes, nodes, sizeof nodes, sizeof(double) * JP_RULE_NMAX * r));
    JP_CUDA(cudaMemcpyToSymbol(c_rule_weights, weights, sizeof weights, sizeof(double) * JP_RULE_LMAX * JP_RULE_NMAX * r));
    JP_CUDA(cudaMemcpyToSymbol(c_rule_npts, npts, sizeof npts, sizeof(int) * JP_RULE_LMAX * r));
    JP_CUDA(cudaMalloc(&ctx->d_rule_nodes[r], sizeof(double) * JP_RULE_NMAX));
    JP_CUDA(cudaMemcpy(ctx->d_rule_nodes[r], nodes, sizeof nodes, cudaMemcpyHostToDevice));
  }
  ctx->rules_uploaded = true;
  return JP_OK;
}
const double* jp_rule_nodes_dev(const jp_ctx* ctx, int rule) { return ctx->d_rule_nodes[rule ? 1 : 0]; }

// ------------------------------------------------------------------------------------ kernels
// In-place exclusive scan of `len` 32-bit counts by ONE block of 32 warps: warp w owns the contiguous chunk w and walks it 32
// entries at a time (coalesced), first for its total, then -- after the 32 totals are scanned -- again with a running carry
// and a shuffle scan per step.  (One thread per chunk, each walking its own entries with strided loads, cost 29 us per radix
// pass at the BASELINE grids.)
__device__ __forceinline__ void jp_block_scan_u32(uint32_t* __restrict__ h, int len) {
  __shared__ uint32_t s_tot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk = (((len + 31) / 32) + 31) & ~31;          // per warp, a multiple of 32
  const int b = w * chunk, e = min(b + chunk, len);
  uint32_t s = 0;
  for (int i = b + lane; i < e; i += 32) s += h[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) s_tot[w] = s;
  __syncthreads();
  if (w == 0) {
    uint32_t v = s_tot[lane], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += u;
    }
    s_tot[lane] = x - v;                                     // exclusive prefix of the warp totals
  }
  __syncthreads();
  uint32_t run = s_tot[w];
  for (int i0 = b; i0 < e; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t v = (i < e) ? h[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += u;
    }
    if (i < e) h[i] = run + x - v;
    run += __shfl_sync(0xffffffffu, x, 31);
  }
}
__global__ void __launch_bounds__(1024) jp_radix_scan_kernel(uint32_t* __restrict__ hist, int nblocks) {
  jp_block_scan_u32(hist + (size_t)blockIdx.x * JP_SORT_BINS * nblocks, JP_SORT_BINS * nblocks);
}

// Multi-block exclusive scan of n 32-bit values (head flags -> segment ids), three small launches: tile sums, scan of the tile
// sums (one block, above), tile-local scan + offset.  out[n] receives the total.
#define JP_SCAN_TILE 4096
__global__ void __launch_bounds__(256) jp_scan_tile_sums_kernel(const uint32_t* __restrict__ in, unsigned long long n,
                                                                uint32_t* __restrict__ tsum) {
  __shared__ uint32_t sw[8];
  const unsigned long long base = (unsigned long long)blockIdx.x * JP_SCAN_TILE;
  uint32_t s = 0;
  for (int j = threadIdx.x; j < JP_SCAN_TILE; j += 256) {
    const unsigned long long i = base + j;
    if (i < n) s += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < 8; ++i) t += sw[i];
    tsum[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) jp_scan_tile_offsets_kernel(uint32_t* __restrict__ tsum, int ntiles) {
  // exclusive scan of the tile sums in place; the grand total goes behind them
  uint32_t last = (ntiles > 0) ? tsum[ntiles - 1] : 0u;
  __syncthreads();
  jp_block_scan_u32(tsum, ntiles);
  __syncthreads();
  if (threadIdx.x == 0) tsum[ntiles] = (ntiles > 0) ? tsum[ntiles - 1] + last : 0u;
}
__global__ void __launch_bounds__(256) jp_scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                            unsigned long long n, const uint32_t* __restrict__ toff, int ntiles) {
  // thread t owns 16 consecutive values of the tile (a 64-byte line): local sums, block scan of the 256 thread sums, write
  __shared__ uint32_t s_w[8];
  const unsigned long long base = (unsigned long long)blockIdx.x * JP_SCAN_TILE + (unsigned long long)threadIdx.x * 16;
  uint32_t v[16], s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0u;
    s += v[j];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += u;
  }
  if (lane == 31) s_w[w] = x;
  __syncthreads();
  uint32_t wbase = 0;
  for (int i = 0; i < w; ++i) wbase += s_w[i];
  uint32_t run = toff[blockIdx.x] + wbase + x - s;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = toff[ntiles];
}

__global__ void jp_iota_kernel(uint32_t* __restrict__ perm, long long n, long long stride) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) perm[(size_t)blockIdx.y * stride + i] = (uint32_t)i;
}

// comp[(k) * (smax+1) + s] = number of compositions of s into k parts, each in [1, cap]
struct EnumParams {
  int d, cap, q, smin, rule;
  int n_class;
  unsigned long long class_off[JP_MAX_D + 2];   // prefix offsets of the classes s = smin .. q
};

__global__ void jp_enum_kernel(EnumParams P, const unsigned long long* __restrict__ comp, int smax,
                               unsigned long long n_mi, uint8_t* __restrict__ mi, double* __restrict__ coef,
                               unsigned long long* __restrict__ npts) {
  unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_mi) return;
  int cls = 0;
  while (cls + 1 < P.n_class && r >= P.class_off[cls + 1]) ++cls;
  unsigned long long rr = r - P.class_off[cls];
  int rem = P.smin + cls;           // remaining sum
  int nfree = 0;
  unsigned long long prod = 1;
  for (int k = 0; k < P.d; ++k) {
    int left = P.d - 1 - k;         // parts after k
    int lo = max(1, rem - left * P.cap), hi = min(P.cap, rem - left);
    int v = lo;
    for (; v <= hi; ++v) {
      unsigned long long c = (left == 0) ? 1ull : comp[(size_t)left * (smax + 1) + (rem - v)];
      if (rr < c) break;
      rr -= c;
    }
    mi[r * P.d + k] = (uint8_t)v;
    rem -= v;
    if (v < P.cap) ++nfree;
    prod *= (unsigned long long)c_rule_npts[P.rule][v - 1];
  }
  // combination coefficient (see file header); binomial via the exact multiplicative recurrence
  int s = P.smin + cls;
  double c;
  if (nfree == 0) {
    c = 1.0;
  } else {
    int J = min(nfree, P.q - s);
    if (J >= nfree) {
      c = 0.0;
    } else {
      double b = 1.0;
      int nn = nfree - 1;
      for (int i = 1; i <= J; ++i) b = __ddiv_rn(__dmul_rn(b, (double)(nn - J + i)), (double)i);
      b = rint(b);
      c = (J & 1) ? -b : b;
    }
  }
  coef[r] = c;
  npts[r] = (c == 0.0) ? 0ull : prod;
}

// single-block exclusive scan of n 64-bit counts; total written to out[n]
__global__ void __launch_bounds__(1024) jp_scan64_kernel(const unsigned long long* __restrict__ in,
                                                         unsigned long long* __restrict__ out, unsigned long long n) {
  __shared__ unsigned long long tot[1024];
  unsigned long long chunk = (n + 1023) / 1024;
  unsigned long long b = threadIdx.x * chunk, e = min(b + chunk, n);
  unsigned long long s = 0;
  for (unsigned long long i = b; i < e; ++i) s += in[i];
  tot[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    unsigned long long v = (threadIdx.x >= o) ? tot[threadIdx.x - o] : 0ull;
    __syncthreads();
    tot[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned long long run = tot[threadIdx.x] - s;
  for (unsigned long long i = b; i < e; ++i) {
    unsigned long long v = in[i];
    out[i] = run;
    run += v;
  }
  if (threadIdx.x == 1023) out[n] = tot[1023];
}

__global__ void jp_expand_kernel(int d, int rule, unsigned long long n_mi, const uint8_t* __restrict__ mi,
                                 const double* __restrict__ coef, const unsigned long long* __restrict__ off,
                                 unsigned long long P, uint8_t* __restrict__ keys, uint8_t* __restrict__ flip,
                                 double* __restrict__ wt) {
  unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P) return;
  // last multi-index a with off[a] <= t
  unsigned long long lo = 0, hi = n_mi;
  while (hi - lo > 1) {
    unsigned long long mid = (lo + hi) >> 1;
    if (off[mid] <= t) lo = mid; else hi = mid;
  }
  const uint8_t* m = mi + lo * d;
  unsigned long long local = t - off[lo];
  // mixed radix, LAST dimension fastest (generation order of the CPU restatement)
  uint8_t j[JP_MAX_D];
  for (int k = d - 1; k >= 0; --k) {
    unsigned np = (unsigned)c_rule_npts[rule][m[k] - 1];
    j[k] = c_rule_index[rule][m[k] - 1][local % np];     // position in the level -> master index (the key byte)
    local /= np;
  }
  double w = coef[lo];
  uint8_t first = 0;     // first non-zero master index: even = negative node = this point is the mirrored member
  for (int k = 0; k < d; ++k) {
    w = __dmul_rn(w, c_rule_weights[rule][m[k] - 1][j[k]]);
    keys[(size_t)k * P + t] = j[k];
    if (first == 0) first = j[k];
  }
  flip[t] = (first != 0 && (first & 1) == 0) ? 1 : 0;
  wt[t] = w;
}

// Mirror order (see the CPU restatement, oracle/jp_oracle.cpp orc_smolyak_build): sort by the CANONICAL key -- the
// point's own key when its first non-zero coordinate is a positive node (odd master index), else the key of -z
// (master indices 2j-1 <-> 2j swapped) -- and, least significant, by the mirrored flag.  Node 0 is then the
// origin and nodes 2j-1, 2j are a mirror pair (z, -z): the tensor-core GLM path evaluates a pair from one
// contraction.  Duplicates share raw keys, hence canonical keys: the stable sort keeps their generation order.
struct KeyDigit {
  const uint8_t* keys;   // column of the current dimension: keys + k * P
  const uint8_t* flip;
  __device__ __forceinline__ unsigned operator()(int, uint32_t src) const {
    const unsigned j = keys[src];
    return (flip[src] && j) ? (((j - 1u) ^ 1u) + 1u) : j;
  }
};
struct FlipDigit {
  const uint8_t* flip;
  __device__ __forceinline__ unsigned operator()(int, uint32_t src) const { return flip[src]; }
};

// head[i] = 1 if sorted element i starts a new key
__global__ void jp_heads_kernel(int d, unsigned long long P, const uint8_t* __restrict__ keys,
                                const uint32_t* __restrict__ perm, uint32_t* __restrict__ head) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t h = 1;
  if (i > 0) {
    uint32_t a = perm[i], b = perm[i - 1];
    h = 0;
    for (int k = 0; k < d; ++k)
      if (keys[(size_t)k * P + a] != keys[(size_t)k * P + b]) { h = 1; break; }
  }
  head[i] = h;
}

// seg_start[seg] = sorted position of the head of segment seg
__global__ void jp_segstart_kernel(unsigned long long P, const uint32_t* __restrict__ head,
                                   const uint32_t* __restrict__ segid, uint32_t* __restrict__ seg_start) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  if (head[i]) seg_start[segid[i]] = (uint32_t)i;
}

// one thread per merged node: sequential weight sum over its run (stable order), key + |z|^2/2 out
__global__ void jp_merge_kernel(int d, int rule, unsigned long long P, long long M, const uint8_t* __restrict__ keys,
                                const double* __restrict__ wt, const uint32_t* __restrict__ perm,
                                const uint32_t* __restrict__ seg_start, uint8_t* __restrict__ idx,
                                double* __restrict__ w, double* __restrict__ hzz) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = m < M;
  unsigned long long b = 0, e = 0;
  if (live) {
    b = seg_start[m];
    e = (m + 1 < M) ? seg_start[m + 1] : P;
  }
  // Runs longer than 32 (the origin collects one point of every multi-index: thousands) are summed by the whole warp: the 32
  // lanes gather 32 weights at once, lane 0's order of additions is unchanged (position order = generation order, the sum
  // is bit-identical to the sequential loop), only the load latency is no longer paid once per element.
  const int lane = threadIdx.x & 31;
  unsigned longmask = __ballot_sync(0xffffffffu, live && (e - b) > 32);
  double s = 0;
  bool done = false;
  while (longmask) {
    const int src_lane = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const unsigned long long lb = __shfl_sync(0xffffffffu, b, src_lane), le = __shfl_sync(0xffffffffu, e, src_lane);
    double acc = 0;
    bool first = true;
    for (unsigned long long i0 = lb; i0 < le; i0 += 32) {
      const unsigned long long i = i0 + lane;
      const double v = (i < le) ? wt[perm[i]] : 0.0;
      const int cnt = (int)min(32ull, le - i0);
      for (int j = 0; j < cnt; ++j) {
        const double vj = __shfl_sync(0xffffffffu, v, j);
        acc = first ? vj : __dadd_rn(acc, vj);
        first = false;
      }
    }
    if (lane == src_lane) {
      s = acc;
      done = true;
    }
  }
  if (!live) return;
  if (!done) {
    s = wt[perm[b]];
    for (unsigned long long i = b + 1; i < e; ++i) s = __dadd_rn(s, wt[perm[i]]);
  }
  w[m] = s;
  uint32_t src = perm[b];
  double zz = 0;
  for (int k = 0; k < d; ++k) {
    uint8_t j = keys[(size_t)k * P + src];
    idx[(size_t)k * M + m] = j;
    double z = c_rule_nodes[rule][j];
    zz += z * z;
  }
  hzz[m] = 0.5 * zz;
}

// max of a positive array by one block (the grid's largest |z|^2 / 2)
__global__ void __launch_bounds__(1024) jp_grid_max_kernel(const double* __restrict__ v, long long n, double* __restrict__ out) {
  __shared__ double sm[33];
  double m = 0;
  for (long long i = threadIdx.x; i < n; i += 1024) m = fmax(m, v[i]);
  m = jp_block_max(m, sm);
  if (threadIdx.x == 0) out[0] = m;
}

// ------------------------------------------------------------------------------------ host driver
int jp_grid_build(jp_ctx* ctx, int rule, int d, int level, jp_grid* g) {
  JP_REQUIRE(rule == 0 || rule == 1, "jp_grid_get: unknown rule %d", rule);
  JP_REQUIRE(d >= 1 && d <= JP_MAX_D, "jp_grid_get: d_eff=%d out of range [1,%d]", d, JP_MAX_D);
  JP_REQUIRE(level >= 1 && level <= 32, "jp_grid_get: level=%d out of range", level);
  JP_TRY(upload_rules(ctx));
  JpRule R = jp_get_rule(rule);
  const int cap = std::min(level, R.levels);
  const int q = level + d - 1;
  // non-zero combination coefficients need |i|_1 >= q - d + 1, or the all-cap index when the level
  // outruns the rule table (then the grid is the full tensor product of the highest rules)
  int smin = std::max(d, q - d + 1);
  int smax = std::min(q, d * cap);
  if (smax < smin) smin = smax = d * cap;
  // restricted-composition counts comp[k][s], k parts in [1,cap]
  std::vector<unsigned long long> comp((size_t)(d + 1) * (smax + 1), 0ull);
  comp[0] = 1;  // zero parts, sum zero
  for (int k = 1; k <= d; ++k)
    for (int s = k; s <= smax; ++s) {
      unsigned long long c = 0;
      for (int v = 1; v <= cap && v <= s; ++v) c += comp[(size_t)(k - 1) * (smax + 1) + (s - v)];
      comp[(size_t)k * (smax + 1) + s] = c;
    }
  EnumParams P;
  P.d = d; P.cap = cap; P.q = q; P.smin = smin; P.rule = rule;
  P.n_class = smax - smin + 1;
  unsigned long long n_mi = 0;
  for (int c = 0; c < P.n_class; ++c) {
    P.class_off[c] = n_mi;
    n_mi += comp[(size_t)d * (smax + 1) + (smin + c)];
  }
  P.class_off[P.n_class] = n_mi;
  JP_REQUIRE(n_mi > 0 && n_mi < (1ull << 31), "jp_grid_get: %llu multi-indices is out of range", n_mi);

  cudaStream_t st = ctx->stream;
  unsigned long long *d_comp = nullptr, *d_npts = nullptr, *d_off = nullptr;
  uint8_t* d_mi = nullptr;
  double* d_coef = nullptr;
  JP_CUDA(jp_dmalloc(ctx, &d_comp, comp.size() * 8));
  JP_CUDA(cudaMemcpyAsync(d_comp, comp.data(), comp.size() * 8, cudaMemcpyHostToDevice, st));
  JP_CUDA(jp_dmalloc(ctx, &d_mi, n_mi * d));
  JP_CUDA(jp_dmalloc(ctx, &d_coef, n_mi * 8));
  JP_CUDA(jp_dmalloc(ctx, &d_npts, n_mi * 8));
  JP_CUDA(jp_dmalloc(ctx, &d_off, (n_mi + 1) * 8));
  jp_enum_kernel<<<(unsigned)((n_mi + 255) / 256), 256, 0, st>>>(P, d_comp, smax, n_mi, d_mi, d_coef, d_npts);
  JP_CHECK_LAUNCH(ctx);
  jp_scan64_kernel<<<1, 1024, 0, st>>>(d_npts, d_off, n_mi);
  JP_CHECK_LAUNCH(ctx);
  unsigned long long Ptot = 0;
  JP_CUDA(cudaMemcpyAsync(&Ptot, d_off + n_mi, 8, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  JP_REQUIRE(Ptot > 0 && Ptot < (1ull << 31), "jp_grid_get: %llu pre-merge points is out of range", Ptot);
  // count multi-indices with non-zero coefficient (reporting only)
  {
    std::vector<double> hc(n_mi);
    JP_CUDA(cudaMemcpy(hc.data(), d_coef, n_mi * 8, cudaMemcpyDeviceToHost));
    long long nz = 0;
    for (double c : hc) nz += (c != 0.0);
    g->n_multi = nz;
  }
  g->n_premerge = (long long)Ptot;

  uint8_t *d_keys = nullptr, *d_flip = nullptr;
  double* d_wt = nullptr;
  uint32_t *d_pa = nullptr, *d_pb = nullptr, *d_hist = nullptr, *d_head = nullptr, *d_seg = nullptr, *d_start = nullptr;
  int nb = jp_sort_blocks((long long)Ptot);
  JP_CUDA(jp_dmalloc(ctx, &d_keys, Ptot * d));
  JP_CUDA(jp_dmalloc(ctx, &d_flip, Ptot));
  JP_CUDA(jp_dmalloc(ctx, &d_wt, Ptot * 8));
  JP_CUDA(jp_dmalloc(ctx, &d_pa, Ptot * 4));
  JP_CUDA(jp_dmalloc(ctx, &d_pb, Ptot * 4));
  JP_CUDA(jp_dmalloc(ctx, &d_hist, (size_t)JP_SORT_BINS * nb * 4));
  JP_CUDA(jp_dmalloc(ctx, &d_head, Ptot * 4));
  JP_CUDA(jp_dmalloc(ctx, &d_seg, (Ptot + 1) * 4));
  uint32_t* d_tsum = nullptr;
  JP_CUDA(jp_dmalloc(ctx, &d_tsum, ((Ptot + JP_SCAN_TILE - 1) / JP_SCAN_TILE + 2) * 4));
  unsigned gp = (unsigned)((Ptot + 255) / 256);
  jp_expand_kernel<<<gp, 256, 0, st>>>(d, rule, n_mi, d_mi, d_coef, d_off, Ptot, d_keys, d_flip, d_wt);
  JP_CHECK_LAUNCH(ctx);
  jp_iota_kernel<<<dim3(gp, 1), 256, 0, st>>>(d_pa, (long long)Ptot, (long long)Ptot);
  JP_CHECK_LAUNCH(ctx);
  uint32_t *pin = d_pa, *pout = d_pb;
  {   // least significant of all: the mirrored flag
    FlipDigit f{d_flip};
    JP_TRY(jp_radix_pass(ctx, f, pin, pout, (long long)Ptot, (long long)Ptot, d_hist, 1));
    std::swap(pin, pout);
  }
  for (int k = d - 1; k >= 0; --k) {   // LSD: least significant dimension first
    KeyDigit f{d_keys + (size_t)k * Ptot, d_flip};
    JP_TRY(jp_radix_pass(ctx, f, pin, pout, (long long)Ptot, (long long)Ptot, d_hist, 1));
    std::swap(pin, pout);
  }
  jp_heads_kernel<<<gp, 256, 0, st>>>(d, Ptot, d_keys, pin, d_head);
  JP_CHECK_LAUNCH(ctx);
  {
    const int ntiles = (int)((Ptot + JP_SCAN_TILE - 1) / JP_SCAN_TILE);
    jp_scan_tile_sums_kernel<<<ntiles, 256, 0, st>>>(d_head, Ptot, d_tsum);
    JP_CHECK_LAUNCH(ctx);
    jp_scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(d_tsum, ntiles);
    JP_CHECK_LAUNCH(ctx);
    jp_scan_apply_kernel<<<ntiles, 256, 0, st>>>(d_head, d_seg, Ptot, d_tsum, ntiles);
    JP_CHECK_LAUNCH(ctx);
  }
  uint32_t M32 = 0;
  JP_CUDA(cudaMemcpyAsync(&M32, d_seg + Ptot, 4, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  long long M = M32;
  JP_CUDA(jp_dmalloc(ctx, &d_start, (size_t)M * 4));
  jp_segstart_kernel<<<gp, 256, 0, st>>>(Ptot, d_head, d_seg, d_start);
  JP_CHECK_LAUNCH(ctx);
  JP_CUDA(jp_dmalloc(ctx, &g->d_idx, (size_t)M * d));
  JP_CUDA(jp_dmalloc(ctx, &g->d_w, (size_t)M * 8));
  JP_CUDA(jp_dmalloc(ctx, &g->d_hzz, (size_t)M * 8));
  jp_merge_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(d, rule, Ptot, M, d_keys, d_wt, pin, d_start,
                                                                 g->d_idx, g->d_w, g->d_hzz);
  JP_CHECK_LAUNCH(ctx);
  JP_CUDA(cudaStreamSynchronize(st));
  // largest |z|^2 over the grid (needed by the TC path's series-length bound): every coordinate can
  // sit at most at the largest node of the highest 1-D level in use, but the sum constraint couples
  // them; read it back from the device table instead of bounding it.
  {
    jp_grid_max_kernel<<<1, 1024, 0, st>>>(g->d_hzz, M, ctx->d_scratch);
    JP_CHECK_LAUNCH(ctx);
    double mx = 0;
    JP_CUDA(cudaMemcpyAsync(&mx, ctx->d_scratch, 8, cudaMemcpyDeviceToHost, st));
    JP_CUDA(cudaStreamSynchronize(st));
    g->zmax2 = 2.0 * mx;
  }
  g->ctx = ctx; g->rule = rule; g->d = d; g->level = level; g->M = M;
  jp_dfree(ctx, d_comp); jp_dfree(ctx, d_mi); jp_dfree(ctx, d_coef); jp_dfree(ctx, d_npts); jp_dfree(ctx, d_off);
  jp_dfree(ctx, d_keys); jp_dfree(ctx, d_flip); jp_dfree(ctx, d_wt); jp_dfree(ctx, d_pa); jp_dfree(ctx, d_pb); jp_dfree(ctx, d_hist);
  jp_dfree(ctx, d_head); jp_dfree(ctx, d_seg); jp_dfree(ctx, d_start); jp_dfree(ctx, d_tsum);
  return JP_OK;
}
