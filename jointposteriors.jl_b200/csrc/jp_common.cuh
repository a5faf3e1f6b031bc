// jp_common.cuh -- shared internals of libjpcuda.so (sm_100a only).
#pragma once
#include <algorithm>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <tuple>

#include "../../include/jpcuda.h"

#define JP_NUM_SMS_B200 148
#define JP_MAX_D 64          // maximum number of unconstrained coordinates
#define JP_MAX_HYPER 8
#define JP_RULE_NMAX 128    // master 1-D nodes per rule family (Genz-Keister: 99, Kronrod-Patterson: 63)

// ---------------------------------------------------------------------------- errors
void jp_set_error(const char* fmt, ...);

#define JP_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      jp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));          \
      return JP_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define JP_CHECK_LAUNCH(ctx)                                                                       \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) {                                                                       \
      jp_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));      \
      return JP_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define JP_REQUIRE(cond, ...)                                                                      \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      jp_set_error(__VA_ARGS__);                                                                   \
      return JP_ERR_BAD_ARG;                                                                       \
    }                                                                                              \
  } while (0)

// every public entry point that launches work makes its context's GPU current first (several GPUs in one process)
#define JP_ENTER_CTX(c) JP_CUDA(cudaSetDevice((c)->device))

#define JP_TRY(expr)                                                                               \
  do {                                                                                             \
    int _s = (expr);                                                                               \
    if (_s != JP_OK) return _s;                                                                    \
  } while (0)

// ---------------------------------------------------------------------------- handles
struct jp_grid {
  jp_ctx* ctx = nullptr;
  int rule = 0, d = 0, level = 0;
  long long M = 0;
  long long n_multi = 0, n_premerge = 0;
  uint8_t* d_idx = nullptr;   // SoA keys: d_idx[k * M + m]
  double* d_w = nullptr;      // merged quadrature weights
  double* d_hzz = nullptr;    // 0.5 * |z_m|^2 (importance correction)
  double zmax2 = 0;           // max_m |z_m|^2
};

// Stage tracing (the reference has none): when switched on (jp_ctx_trace), entry points drop CUDA events
// at their phase boundaries on whichever stream the phase runs on; jp_ctx_trace_dump turns them into a timeline.
struct JpTrace {
  bool on = false;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
};

struct jp_ctx {
  int device = 0;
  JpTrace trace;
  int sm_count = JP_NUM_SMS_B200;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long launches = 0;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // bracket the last launch of the dominant (log-density) kernel
  bool ev_valid = false;
  std::map<std::tuple<int, int, int>, jp_grid*> grids;
  // scratch for small reductions / scalars (device) and pinned host staging
  double* d_scratch = nullptr;     // JP_SCRATCH_DOUBLES doubles
  double* h_pinned = nullptr;      // JP_PINNED_DOUBLES doubles
  unsigned int* d_counters = nullptr;   // JP_COUNTERS arrival counters of multi-block reductions (zero between kernels)
  double* d_bpart = nullptr;       // JP_BPART_DOUBLES per-block partials of those reductions
  // __constant__ tables and the device copy of the master node tables are PER DEVICE: every context uploads its own
  // (several GPUs in one process each get theirs; a second context on the same device re-uploads identical bytes)
  cudaStream_t side = nullptr, side2 = nullptr;   // further streams of the context: the independent O(N) passes of a fit's
                                                  // preparation overlap on them (forked from / joined into `stream`)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
  cudaEvent_t ev_pinned = nullptr;   // recorded after the last asynchronous copy OUT of h_pinned (see jp_pinned_acquire)
  bool pinned_busy = false;
  // The reference's fit returns `density` as a host vector: jp_fit queues its download on `side` right behind stage 4, into this
  // pinned buffer, so that it overlaps whatever the caller does next on the device (the marginals); jp_get_density then only waits
  // for that copy.  One buffer per context: valid for the posterior / fit generation recorded beside it.
  double* h_density = nullptr;
  size_t h_density_cap = 0;
  const void* h_density_owner = nullptr;
  long long h_density_gen = -1;
  cudaEvent_t ev_density = nullptr, ev_stage4 = nullptr;
  bool rules_uploaded = false, fit_nodes_uploaded = false, tc_tables_uploaded = false;
  double* d_rule_nodes[2] = {nullptr, nullptr};
};
#define JP_COUNTERS 1024
#define JP_BPART_DOUBLES 4096
#define JP_RED_BLOCKS_MAX 512
// Stream-ordered device memory from the device's pool (cudaMallocAsync with an unbounded release
// threshold, set in jp_ctx_create): the buffers of freed posteriors / data sets are recycled by the next
// allocation without a device synchronisation, which is what keeps repeated fit() calls cheap.
template <class T>
inline cudaError_t jp_dmalloc(jp_ctx* ctx, T** p, size_t bytes) {
  return cudaMallocAsync((void**)p, bytes ? bytes : 16, ctx->stream);
}
inline void jp_dfree(jp_ctx* ctx, const void* p) {
  if (p) cudaFreeAsync(const_cast<void*>(p), ctx->stream);
}
// The pinned staging buffer is shared by every entry point of a context.  Before the HOST writes into it, wait until the last
// asynchronous host-to-device copy that reads it has completed (an event, not a stream synchronisation: a fit queued behind
// another fit only waits for that fit's three small uploads); after queueing copies out of it, publish them.
inline cudaError_t jp_pinned_acquire(jp_ctx* ctx) {
  if (!ctx->pinned_busy) return cudaSuccess;
  ctx->pinned_busy = false;
  return cudaEventSynchronize(ctx->ev_pinned);
}
inline cudaError_t jp_pinned_publish(jp_ctx* ctx) {
  ctx->pinned_busy = true;
  return cudaEventRecord(ctx->ev_pinned, ctx->stream);
}
inline void jp_trace_mark(jp_ctx* ctx, const char* name, cudaStream_t st) {
  if (!ctx->trace.on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  ctx->trace.marks.emplace_back(name, e);
}
#define JP_MARK(ctx, name) jp_trace_mark((ctx), (name), (ctx)->stream)
#define JP_MARK_SIDE(ctx, name) jp_trace_mark((ctx), (name), (ctx)->side)
#define JP_SCRATCH_DOUBLES (1 << 16)
#define JP_PINNED_TAIL_DOUBLES 16      // last doubles of the pinned buffer: the device-side series decision on its way to the host
#define JP_PINNED_DOUBLES (1 << 18)   // 2 MB: result vectors of up to 262 144 nodes are downloaded through it
// Fixed sub-regions of the pinned staging buffer.  [0, JP_PINNED_CONSTS_END): the per-fit constants (mu d, U d x p, transform
// codes d: at most 64 + 4096 + 32 doubles) staged by jp_upload_fit_consts; the bounds of the tensor-core path's a-priori gate
// are read back behind them.  Host writes into the constants region wait for the event recorded after the last copy out of it.
#define JP_PINNED_CONSTS_END 4352
#define JP_PINNED_BOUNDS_OFF JP_PINNED_CONSTS_END
static_assert(JP_MAX_D + JP_MAX_D * JP_MAX_D + JP_MAX_D / 2 + 1 <= JP_PINNED_CONSTS_END, "pinned constants region too small");

struct jp_data {
  jp_ctx* ctx = nullptr;
  int family = 0;
  long long N = 0;
  int ncols = 0;
  double* d_obs = nullptr;      // N x ncols row-major
  bool owns_obs = true;         // false: d_obs belongs to the caller (jp_data_adopt_device)
  double hyper[JP_MAX_HYPER] = {0};
  int n_hyper = 0;
  void* tc_state = nullptr;     // per-data state of the GLM tensor-core path (jp_glm_tc.cu): the operand
                                // [x_hi | x_lo | x_hi] in TF32-representable FP32, its TMA map, per-obs coefficients
};

// what the per-node "finish" of a log-density launch needs (deferred into the fused stage-4 kernel, jp_fit.cu)
struct JpFinish {
  int path = 0;                   // JP_PATH_FP64 / JP_PATH_TC
  int splits = 0;                 // FP64 path: observation splits; part [splits][M], lj_prior [M]
  const double* part = nullptr;
  const double* lj_prior = nullptr;
  int chunks = 0;                 // TC path: observation chunks; tc_part [chunks][P][2] (even, odd), quad [M]
  long long P = 0, j_lo = 0;
  const double* tc_part = nullptr;
  const double* quad = nullptr;
  double neg_min = 0;
};

// one cached smooth-CDF design (jp_marginal_smooth_keyed): sorted design matrix and cumulative weights of one marginal function
struct JpDesignCache {
  long long key, gen;
  double* d_V;
  double* d_cw;
  double mu, sigma;
  long long stamp;
};
#define JP_DESIGN_CACHE 4

struct jp_posterior {
  jp_ctx* ctx = nullptr;
  const jp_grid* grid = nullptr;
  const jp_data* data = nullptr;
  int d = 0, p = 0;
  long long m0 = 0, m1 = 0, M = 0;   // shard [m0, m1), M = m1 - m0
  int path_used = 0;
  std::vector<int> tcode_host;   // transform codes of the last fit (host copy)
  long long fit_gen = 0;         // counts the fits of this posterior: what was derived from an earlier fit is stale
  long long cache_clock = 0;
  std::vector<JpDesignCache> design_cache;
  bool raw = false;              // RawBuild: d_theta holds the UNCONSTRAINED node coordinates (the reference's grid.cache)
  JpFinish fin;                  // pending finish of the last log-density launch (consumed by jp_stage4_launch)
  double* d_theta = nullptr;     // SoA constrained parameters [d][M]
  double* d_a = nullptr;         // log-density + neg_min + 0.5|z|^2
  double* d_logdens = nullptr;   // log-density + neg_min
  double* d_density = nullptr;   // normalised weights
  double* d_part = nullptr;      // (JP_POST_PART_SPLITS + 1) x M doubles, see below
  double* d_stats = nullptr;     // [0]=max, [1]=sum (device scalars for the single-GPU path)
  // per stage-4 block and coordinate: (sum w theta, sum w theta^2, min theta, max theta), left by the fused stage-4 kernel of a
  // single-GPU fit; coordinate marginals then need ONE launch (jp_marginal_onepass_kernel)
  double* d_cmom = nullptr;
  int cmom_blocks = 0, cmom_cap = 0;
  bool cmom_valid = false;
  int* d_coords = nullptr;       // coordinates of the last one-pass marginal batch on the device
  std::vector<int> coords_host;
  void* tc_state = nullptr;      // TC path: node operand, tensor map, quadratic part (jp_glm_tc.cu)
  double tc_bounds[8] = {0};     // TC path diagnostics of the last fit: max|Delta|, truncation bound, rounding
                                 // estimate, series coefficients used, worst-case rounding bound
  // per-fit constants on the device
  double* d_mu = nullptr;        // d
  double* d_U = nullptr;         // d x p column-major
  int* d_tcode = nullptr;        // d
  // stage-5 work buffers (allocated on first use)
  int K_cap = 0;                 // capacity of the default (sort-free) path: d_vptr, d_bins, d_mout
  int K_cap_vals = 0;            // capacity of d_vals
  int K_cap_sort = 0;            // capacity of the explicit sort: d_perm_*, d_hist, d_sv, d_sw, d_cw
  int K_last = 0;                // marginals of the last call
  bool sorted_valid = false;     // d_sv / d_sw / d_cw hold the sort of the last call's marginals
  int bins_blocks = 0;
  double* d_vals = nullptr;      // uploaded values K x M (host closures)
  const double** d_vptr = nullptr;  // K device pointers to the value columns
  std::vector<const double*> vptr_host;   // what d_vptr holds (a repeated request skips the upload and its wait)
  double* d_bins = nullptr;      // [K][bins_blocks][99][5] per-block bins of the sort-free path
  uint32_t* d_perm_a = nullptr;  // K x M
  uint32_t* d_perm_b = nullptr;  // K x M
  uint32_t* d_hist = nullptr;    // radix histograms
  double* d_sv = nullptr;        // sorted values K x M
  double* d_sw = nullptr;        // sorted weights K x M
  double* d_cw = nullptr;        // cumulative weights K x M
  double* d_mout = nullptr;      // K x (2 + 200 + 2) results
  double* d_cand = nullptr;      // K x 98 x 6 knot candidates on their way to the peers (jp_marginal_coords_p2p)
  int K_cap_cand = 0;
};
#define JP_POST_PART_SPLITS 32   // d_part holds [splits <= 32][M] observation partial sums + [M] (lj + prior)

// ---------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ double jp_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double jp_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double jp_warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Deterministic block sum: fixed shuffle tree inside warps, fixed order across warps.
// All threads must call; result valid in every thread.  smem: >= 33 doubles.
__device__ __forceinline__ double jp_block_sum(double v, double* smem) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = jp_warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < nw; ++i) s += smem[i];
    smem[32] = s;
  }
  __syncthreads();
  return smem[32];
}
__device__ __forceinline__ double jp_block_max(double v, double* smem) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = jp_warp_max(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = smem[0];
    for (int i = 1; i < nw; ++i) s = fmax(s, smem[i]);
    smem[32] = s;
  }
  __syncthreads();
  return smem[32];
}
__device__ __forceinline__ double jp_block_min(double v, double* smem) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = jp_warp_min(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = smem[0];
    for (int i = 1; i < nw; ++i) s = fmin(s, smem[i]);
    smem[32] = s;
  }
  __syncthreads();
  return smem[32];
}
// asynchronous global -> shared copies (LDGSTS): many bytes in flight per thread without register staging
__device__ __forceinline__ void jp_cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void jp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void jp_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// Multi-block reduction, second stage in the same launch: every block publishes its partial, the block that arrives
// last (arrival counter, self-resetting) combines the partials in block order -- deterministic -- and writes the result.
// Call with all threads after the block's partial has been written by thread 0; true in the last block only.
__device__ __forceinline__ bool jp_last_block(unsigned int* counter, unsigned int nblocks) {
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == nblocks - 1u);
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  const bool last = s_last != 0u;
  if (last) __threadfence();
  return last;
}
static inline int jp_red_blocks(long long M) { return (int)std::max(1LL, std::min((long long)JP_RED_BLOCKS_MAX, (M + 1023) / 1024)); }
// order-preserving map double -> uint64 (ascending), -0.0 < +0.0; NaNs sort last/first by sign
__device__ __forceinline__ unsigned long long jp_sortable(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
#endif

// ---------------------------------------------------------------------------- in-library exchange over NVLink peer memory
// One MAILBOX per rank (plain cudaMalloc memory, exported through CUDA IPC and mapped by every peer): small per-channel
// slots that the peers WRITE into with ordinary stores over NVLink, one 8-byte flag per (channel, parity, sender) that
// carries the sequence number of the exchange, and a bulk region for the coefficient rows of the tensor-core path.  An
// exchange is one kernel: block p copies this rank's payload into peer p's slot, publishes it (fence.sys + st.release.sys of
// the sequence number) and then waits for peer p's payload to arrive in the own mailbox (ld.acquire.sys).  No NCCL, no host
// round trip; slots alternate by sequence parity, and every consumer of a slot precedes the rank's next send on that channel
// in stream order, so a sender can never overwrite data its peer still reads.
#define JP_COMM_MAX_WORLD 8
enum { JP_CH_PREP = 0, JP_CH_STATS, JP_CH_MOM, JP_CH_KNOTS, JP_CH_BULK, JP_CH_USER, JP_COMM_NCHAN };
#define JP_COMM_HEADER_BYTES 4096      // error word + flags
#define JP_COMM_KMAX 64                // marginals per exchange of the knot candidates
struct JpCommDev {                     // what the kernels need, by value
  unsigned char* const* peer;          // device array [world]: mailbox base of every rank as mapped HERE (own one at [rank])
  unsigned char* self;
  int rank, world;
  long long timeout_ns;
};
struct jp_comm {
  jp_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  unsigned char* mailbox = nullptr;
  size_t bytes = 0, bulk_off = 0, bulk_bytes = 0;
  unsigned char* peer_h[JP_COMM_MAX_WORLD] = {nullptr};
  bool ipc_opened[JP_COMM_MAX_WORLD] = {false};
  bool connected = false;
  unsigned char** d_peer = nullptr;
  unsigned long long seq[JP_COMM_NCHAN] = {0};
  size_t data_off[JP_COMM_NCHAN][2] = {{0}};
  size_t cap_doubles[JP_COMM_NCHAN] = {0};     // capacity of one (channel, parity) region, all ranks together
  unsigned int* d_counter = nullptr;           // arrival counter of the bulk push
  long long timeout_ns = 60000000000ll;
};
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long* jp_comm_flag(unsigned char* mailbox, int chan, int parity, int sender) {
  return reinterpret_cast<unsigned long long*>(mailbox + 64) + ((size_t)(chan * 2 + parity) * JP_COMM_MAX_WORLD + sender);
}
__device__ __forceinline__ void jp_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long jp_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long jp_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until the flag carries `seq` (or a later exchange's number); on timeout record the channel in the mailbox's error word
__device__ __forceinline__ void jp_comm_wait_flag(const JpCommDev& c, int chan, int parity, int sender, unsigned long long seq) {
  const unsigned long long* f = jp_comm_flag(c.self, chan, parity, sender);
  const long long t0 = jp_globaltimer();
  while (jp_ld_acquire_sys(f) < seq) {
    __nanosleep(64);
    if (jp_globaltimer() - t0 > c.timeout_ns) {
      atomicExch(reinterpret_cast<int*>(c.self), 1 + chan + 16 * sender);
      break;
    }
  }
}
#endif
JpCommDev jp_comm_dev(const jp_comm* c);
// all_gather of n doubles per rank on channel `chan` (stream-ordered, asynchronous): *d_gathered = [world][n] in the own mailbox
int jp_comm_exchange(jp_comm* c, int chan, const double* d_src, int n, const double** d_gathered);
// the bulk region (coefficient rows): its flags follow the same sequence protocol on JP_CH_BULK, single-buffered
int jp_comm_bulk_begin(jp_comm* c, unsigned long long* seq);                 // next sequence number of the bulk channel
int jp_comm_wait(jp_comm* c, int chan, int parity, unsigned long long seq);  // one tiny kernel: all senders' flags >= seq

// ---------------------------------------------------------------------------- internal entry points
int jp_grid_build(jp_ctx* ctx, int rule, int d, int level, jp_grid* g);
// finish = false: the per-node finish (sum of the partials -> log-density) is left to the fused stage-4 kernel (post->fin)
int jp_fit_fp64_launch(jp_posterior* post, const jp_fit_args* args, bool finish = true);   // stages 2-3, generic plugin kernel
int jp_fit_tc_launch(jp_posterior* post, const jp_fit_args* args, bool finish = true);     // stages 2-3, GLM tensor-core path
bool jp_fit_tc_supported(const jp_posterior* post, const jp_fit_args* args);
// the tensor-core path phase by phase (observation-sharded prep of a node-sharded fit)
int jp_fit_tc_prep_len(int d);
int jp_fit_tc_prep_local(jp_posterior* post, const jp_fit_args* args, int rank, int world, double* d_out);
int jp_fit_tc_prep_gathered(jp_posterior* post, const jp_fit_args* args, const double* d_gathered, int world, int rank,
                            int* n_rows);
int jp_fit_tc_coef_slab(jp_posterior* post, int n_rows, float** d_local, float** d_all, long long* count);
int jp_fit_tc_run_prepared(jp_posterior* post, const jp_fit_args* args, bool finish = true);
int jp_fit_tc_launch_dev(jp_posterior* post, const jp_fit_args* args, jp_comm* comm, int obs_sharded);   // device-side series-length decision
int jp_fit_tc_verify_prefetch(jp_posterior* post);   // queue the read-back in front of the caller's own synchronisation
int jp_fit_tc_verify(jp_posterior* post);      // read that decision back (first blocking call after the fit)
void jp_tc_data_free(jp_data* data);
void jp_tc_post_free(jp_posterior* post);
int jp_construct_columns(jp_posterior* post, int K, const int* d_coords, double* d_out);   // RawBuild: constrained columns from the cache
int jp_glm_grad_hess_comm(jp_ctx* ctx, const jp_data* data, jp_comm* comm, int d, const double* h_beta, double* h_g, double* h_Hneg,
                          double* h_logpost);
int jp_density_prefetch(jp_posterior* post);      // queue the download of the normalised weights behind stage 4 (jp_api.cu)
int jp_upload_fit_consts(jp_posterior* post, const jp_fit_args* args);   // mu_hat, U, transform codes -> device
int jp_marginal_design_device(jp_posterior* post, int k, double** d_V, long long** d_ind, double* h_mu, double* h_sigma);   // jp_marginal.cu
const double* jp_rule_nodes_dev(const jp_ctx* ctx, int rule);   // device copy of the master z-node table (this context's GPU)

struct JpRule {
  int levels, nmax;
  const int* npts;
  const double* nodes;
  const double* weights;         // [levels][nmax] by master index
  const unsigned char* index;    // [levels][nmax]: master index of the pos-th node of a level
};
JpRule jp_get_rule(int rule);
