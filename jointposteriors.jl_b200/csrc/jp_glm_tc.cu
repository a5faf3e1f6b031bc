// jp_glm_tc.cu -- STAGES 2-3 for GLM families on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// The hot loop of the reference is M grid nodes x N observations of the user's log_density
// (reference src/joint_posterior.jl:147-154 called from eval_grid!, :180,186).  For a GLM that is the dense
// contraction  eta = X Theta'  followed by a link-function term per (node, observation) pair.
//
// Centred form (mirrors the reference's own mode-centring `+ neg_min`, :149,153).  With theta_m = mu_hat +
// delta_m, delta_m = U z_m, eta_hat_i = x_i . mu_hat and Delta_im = x_i . delta_m, a GLM log-likelihood
// sum_i [ y_i eta_i - b(eta_i) ] expands EXACTLY into
//     L_hat + g . delta_m - 1/2 delta_m' H delta_m - sum_i R_i(Delta_im)
//     L_hat = sum_i y_i eta_hat_i - b(eta_hat_i),  g = X'(y - b'(eta_hat)),  H = X' diag(b''(eta_hat)) X      (FP64)
//     R_i(D) = b(eta_hat_i + D) - b(eta_hat_i) - b'(eta_hat_i) D - 1/2 b''(eta_hat_i) D^2 = sum_{k>=3} c_{k,i} D^k
// The FP64 pieces cost O(N d^2 + M d^2).  Only the third-order remainder needs the node x observation
// product, and it is O(|Delta|^3) small, so FP32 arithmetic on a 3xTF32 contraction is ample: an error of
// 1e-6 RELATIVE to R_i is ~1e-12 absolute per observation (see jp_tc_choose_order for the bounds that gate
// the path; outside them the FP64 plugin kernel of jp_fit.cu is used).
//
// Kernel (one persistent CTA per SM, warp-specialised, 14 warps):
//   warp 0      TMA producer: the mirror-pair operand (96 pairs x K) once per work item into one of two buffers,
//               observation tiles (128 obs x K) plus their coefficient rows through a ring of shared-memory stages;
//               128-byte swizzle, K-major
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::tf32, M = 128 (observations on TMEM lanes),
//               N = 96 (mirror pairs of grid nodes on TMEM columns), K = 8 per instruction; five accumulator buffers in
//               TMEM (5 x 96 columns); tcgen05.commit releases stages / publishes accumulators.  Both warps run their
//               loops converged and predicate the asynchronous instructions on one elected lane.
//   warps 2-13  epilogue (3 per SM sub-partition, each one TMEM lane quarter x 32 columns): tcgen05.ld 8 columns at a
//               time, software-pipelined across tiles; a column is a PAIR of nodes (z, -z): with D = x_i . delta(z),
//               R_i(+-D) = E_i(D) +- O_i(D), E = D^4 (c4 + c6 D^2 + ..), O = D^3 (c3 + c5 D^2 + ..) in packed FFMA2 /
//               FMUL2 (two columns per instruction, NC + 3 operations per column) with the thread's own observation
//               coefficients from shared memory; accumulated per (thread = observation lane, column) in FP32
//               registers over all observation tiles of the item; per item one shuffle transpose-reduce +
//               shared-memory combine in FP64 into per-(chunk, pair) (E, O) partials.
// The NC coefficients are the ECONOMISED ones when that passes the gate (tools/gen_fold.py, tc_fold_kernel): the next
// odd / even Taylor order folded into the kept ones by its minimax approximation on the observation's own interval.
// 3xTF32: operand rows are [x_hi | x_lo | x_hi] and [d_hi | d_hi | d_lo] (TF32-representable FP32), so one
// K = 3d contraction gives x_hi d_hi + x_lo d_hi + x_hi d_lo with FP32 accumulation in TMEM.  When that needs three
// 128-byte K atoms (21 < d <= 32) the SPLIT layout is used instead: observation rows [x_hi | x_lo] with each part
// padded to its own atom, pair rows [d_hi | d_hi | d_lo] likewise, and the third product re-reads the observation
// row's FIRST atom from shared memory -- the streamed operand (the kernel is bound by L2 -> SM traffic of the
// observation tiles there) shrinks by a third, the contraction is unchanged.
#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <type_traits>
#include "jp_common.cuh"
#include "jp_fold_tables.h"

int jp_glm_sums_device(jp_ctx* ctx, const jp_data* data, int d, const double* d_beta, double* d_out, double* d_work,
                       int nblocks);
int jp_glm_sums_device_range_on(jp_ctx* ctx, cudaStream_t stream, const jp_data* data, int d, const double* d_beta, double* d_out,
                                double* d_work, int nblocks, long long o0, long long o1);
int jp_glm_num_blocks(const jp_ctx* ctx, long long N);

#define TC_OBS_TILE 128          // MMA M: observations per tile (TMEM lanes)
#define TC_PAIR_TILE 96          // MMA N: mirror pairs of grid nodes per tile (TMEM columns)
#define TC_TMEM_STRIDE 96        // columns between accumulator buffers
#ifndef TC_RING_UNROLL
#define TC_RING_UNROLL 1         // epilogue tile loop unrolled over the accumulator ring (compile-time buffer indices)
#endif
#if TC_RING_UNROLL
#define TC_NBUF 4                // accumulator buffers in TMEM (4 x 96 of the 512 columns): a power of two, see the epilogue
#else
#define TC_NBUF 5                // accumulator buffers in TMEM (5 x 96 of the 512 columns)
#endif
#define TC_COLS_PER_WARP 32      // each epilogue warp owns one lane quarter x one third of the pair columns
#define TC_EPI_WARPS (4 * TC_PAIR_TILE / TC_COLS_PER_WARP)   // 12: three per SM sub-partition
#define TC_KATOM 32              // fp32 elements per 128-byte swizzle atom
#define TC_NCMAX 12              // stored Taylor coefficients per observation: orders 3 .. 14
#define TC_ORDER_MAX 16          // highest derivative order tabulated (tail bounds need two more than used)
#define TC_EPI_THREADS (32 * TC_EPI_WARPS)
#define TC_THREADS (64 + TC_EPI_THREADS)   // warp 0 TMA, warp 1 MMA, warps 2-13 epilogue
#ifndef TC_MAX_STAGES
#define TC_MAX_STAGES 6          // observation-tile ring depth (shared memory permitting)
#endif
#ifndef TC_LDW
#define TC_LDW 8                 // columns per tcgen05.ld of the epilogue (8 or 16)
#endif
#ifndef TC_EARLY_PEEK
#define TC_EARLY_PEEK 1          // epilogue peeks at the next tile's accumulator barrier one tile early
#endif
#define TC_PREP_BLOCKS 444       // large N: 3 per SM (two 32 KB record tiles each at d = 30), every block streams many tiles
#define TC_PREP_BLOCKS_MAX 1184  // small N: one block per 128-observation tile up to 8 per SM (a single round, no tile loop)
#define TC_NORD 5                // series lengths NC = 4, 6, 8, 10, 12 (orders 6 .. 14)
#define TC_NBOUND (14 + 2 * TC_NFOLD)   // per-block bound partials, see tc_obs_prep_kernel
#define TC_COEF_ROWS (TC_NCMAX + 1)     // coefficient rows per observation + the row of t_i = |U' x_i|

// derivative polynomials of softplus in s = sigmoid(eta): f_1 = s, f_{k+1} = f_k'(s) (s - s^2)
__constant__ double c_sp_poly[TC_ORDER_MAX + 1][TC_ORDER_MAX + 2];
__constant__ double c_inv_fact[TC_ORDER_MAX + 2];
// economisation constants (jp_fold_tables.h): [0] odd kappa, [1] even kappa; eps / growth factors likewise
__constant__ double c_fold_kappa[2][TC_NFOLD][5];
__constant__ double c_fold_eps[2][TC_NFOLD];
__constant__ double c_fold_grow[2][TC_NFOLD];

struct TcDataState {
  int d = 0, kp = 0, ka = 0;   // observation operand: row length (floats) and 128-byte atoms per row
  int split = 0, kp_b = 0;     // split layout (see header) and the pair operand's row length
  int world = 1, rank = 0;     // observation slices of the sharded prep (one per rank) and the slice this rank prepares; N_pad = world * n_loc
  long long n_loc = 0;         // observations per slice, a multiple of the tile
  long long N = 0, N_pad = 0;
  float* d_xs = nullptr;       // [N_pad][kp]  (x_hi | x_lo | x_hi | 0)
  float* d_coef = nullptr;     // [TC_COEF_ROWS][n_loc]: coefficient k of every observation of THIS rank's slice (all observations on
                               // one GPU), contiguous per 128-observation tile; last row t_i
  float* d_coef_all = nullptr; // world > 1: [world][nc_all][n_loc], the first nc_all rows of every rank's slab after ONE all_gather
  int nc_all = 0;
  double* d_sums = nullptr;    // packed (g[d], upper H[d(d+1)/2], L_hat)
  double* d_work = nullptr;    // partials of the GLM sums
  double* d_bounds = nullptr;  // [TC_PREP_BLOCKS_MAX][TC_NBOUND]
  int prep_blocks = TC_PREP_BLOCKS;   // blocks of tc_obs_prep_kernel for a slice of n_loc observations
  double* d_comb = nullptr;    // [TC_NBOUND]: bounds combined over the ranks (sharded prep)
  double* d_loc = nullptr;     // [L]: this rank's (slice sums | slice bounds) on their way to the peers (jp_fit_p2p)
  int glm_blocks = 0;
  cudaEvent_t ev_bounds = nullptr;   // the bounds have reached the host (the fit waits on it, not on the whole stream)
  CUtensorMap tmA;
};

// what tc_decide_kernel leaves: the series length (0: bounds not met), whether the economised coefficients are used, and the
// diagnostics the host-side decision records in post->tc_bounds
struct TcDecision {
  int NC, fold, pad0, pad1;
  double diag[6];
};

struct TcPostState {
  long long P = 0, P_pad = 0, j_lo = 0;   // mirror pairs touched by the local node range; first pair index
  int kp = 0;
  float* d_ds = nullptr;       // [P_pad][kp]  (d_hi | d_hi | d_lo | 0) of each pair's first member
  double* d_quad = nullptr;    // [M]
  double* d_part = nullptr;    // [part_chunks][P][2] per-chunk even / odd remainder sums
  int part_chunks = 0;
  bool node_prep_queued = false;   // sharded prep: tc_node_prep already runs under the host's wait for the bounds
  TcDecision* d_dec = nullptr;   // series length decided on the device (jp_fit_tc_launch_dev)
  bool dec_pending = false;        // ... and not yet read back by the host
  bool dec_prefetched = false;     // ... its copy into the pinned tail is queued (jp_fit_tc_verify_prefetch)
  CUtensorMap tmB;
};


// ------------------------------------------------------------------------------------ small device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r & 0xFFFFE000u);
}
__device__ __forceinline__ void tf32_split(double x, float& hi, float& lo) {
  hi = tf32_round((float)x);
  lo = tf32_round((float)(x - (double)hi));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking peek at a phase (the epilogue looks at the NEXT tile's accumulator barrier a tile early, so that the
// barrier's read latency is off the critical path when the tile boundary comes)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// epilogue-side wait: latency matters, spin on try_wait
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// producer / MMA-issuer wait: these single threads have thousands of cycles of slack (the kernel is bound by
// the epilogue), and a spin loop steals issue slots from the epilogue warps of their SM sub-partition.  try_wait
// with a suspend-time hint parks the thread in hardware until the phase completes (or the hint expires).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// one lane of a converged warp (the same lane every time: the lowest)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, 128-byte swizzle: 8-row groups 1024 bytes apart (SBO), LBO unused (1), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with the descriptors given by their low words (the high word 0x40004040 -- SBO 1024 bytes, version 1,
// 128-byte swizzle -- is a constant): the low word of a descriptor is affine in the shared-memory address,
// lo(addr + off) = lo(addr) + (off >> 4), so the issuer advances it with one add per operand and instruction.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_tf32_lo(uint32_t tmem_d, uint32_t lo_a, uint32_t lo_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(lo_a), "r"(lo_b), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t (&v)[16], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// the registers are operands of the wait so that their uses cannot be scheduled above it
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t (&v)[8], uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld(uint32_t (&v)[8], uint32_t taddr) { tmem_ld8(v, taddr); }
__device__ __forceinline__ void tmem_ld(uint32_t (&v)[16], uint32_t taddr) { tmem_ld16(v, taddr); }
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[8]) { tmem_ld_wait8(v); }
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) { tmem_ld_wait16(v); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }

// ------------------------------------------------------------------------------------ operand preparation
// X' = [x_hi | x_lo | x_hi | 0]: depends on the data only, built once per jp_data.  HBM-bound streaming kernel
// (8 N (d + 1) bytes in, 4 N_pad kp out): threadIdx.x is the operand column, threadIdx.y the row inside the block's
// slab, so the stores of the [N_pad][kp] operand are coalesced, the three reads of a record hit L1, and no thread
// divides.
#define TC_SPLIT_ROWS 8
#define TC_SPLIT_UNROLL 4        // independent record loads in flight per thread (Little: ~800 ns x 6.4 TB/s / 148 SMs)
__global__ void tc_split_x_kernel(int d, int ncols, int kp, int split, long long N, long long N_pad,
                                  const double* __restrict__ obs, float* __restrict__ xs) {
  const int c = threadIdx.x;                     // 0 .. kp - 1
  // compact: [x_hi | x_lo | x_hi] back to back; split: x_hi in atom 0, x_lo in atom 1
  const int part = split ? (c / TC_KATOM) : ((c >= 2 * d) ? 2 : (c >= d ? 1 : 0));
  const int k = split ? (c % TC_KATOM) : (c - part * d);
  const bool live = split ? (k < d) : (c < 3 * d);
  const long long step = (long long)gridDim.x * TC_SPLIT_ROWS;
  for (long long i0 = (long long)blockIdx.x * TC_SPLIT_ROWS + threadIdx.y; i0 < N_pad; i0 += step * TC_SPLIT_UNROLL) {
    double x[TC_SPLIT_UNROLL];
#pragma unroll
    for (int u = 0; u < TC_SPLIT_UNROLL; ++u) {
      const long long i = i0 + u * step;
      x[u] = (live && i < N) ? __ldg(obs + (size_t)i * ncols + k) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < TC_SPLIT_UNROLL; ++u) {
      const long long i = i0 + u * step;
      if (i >= N_pad) break;
      float hi, lo;
      tf32_split(x[u], hi, lo);
      xs[(size_t)i * kp + c] = (part == 1) ? lo : hi;      // rows past N and the padding columns are zero: split(0) = (0, 0)
    }
  }
}

// Per observation: eta_hat, Taylor coefficients c_3 .. c_14 of the link remainder, t = |U' x| (so that
// |Delta| <= t |z|), and the ingredients of the order / eligibility decision reduced per block:
//   [0] max_i t_i   [1] sum |c_3| t^3   [2+j], [7+j] (j = 0..4 <-> NC = 4, 6, 8, 10, 12): truncation tail bounds of
//   the series stopped at order NC + 2, summed over observations at |z| = z_ref and |z| = z_max
//   [12] sum_i |R_i| and [13] sum_i R_i^2 at |z| = z_ref (majorants)
//   [14+j], [14+TC_NFOLD+j] (j <-> NC = 4, 6, 8, 10): the same two bounds for the ECONOMISED series of NC coefficients
//   (orders NC + 3 and NC + 4 folded into the kept ones over |D| <= t_i z_ref, tc_fold_kernel)
#define TC_PREP_THREADS 128      // = observations per staged tile
__global__ void __launch_bounds__(TC_PREP_THREADS)
tc_obs_prep_kernel(int family, int d, int p, int ncols, long long N, long long N_pad, const double* __restrict__ obs,
                   const double* __restrict__ mu, const double* __restrict__ U, double z_ref, double z_max,
                   float* __restrict__ coef, long long coef_stride, double* __restrict__ bounds) {
  // N, N_pad, obs, coef describe the slice of observations this launch covers (all of them on one GPU); coef_stride is
  // the row stride of the whole coefficient array
  // HBM-bound streaming kernel (8 N (d + 1) bytes in, 48 N_pad out): a block copies a tile of 128 records to shared
  // memory with coalesced loads (row stride padded to an odd number of doubles: conflict-free column access), then a
  // thread owns one record.  t^2 = |U' x|^2 takes U four columns at a time (two 16-byte broadcast loads per record
  // element and four FMAs), so the loop is bound by the FP64 pipe, not by shared-memory traffic.
  extern __shared__ __align__(16) double sh[];
  const int p4 = (p + 3) & ~3, rs = ncols | 1;
  double* s_mu = sh;                       // d
  double* s_U = sh + ((d + 1) & ~1);       // d x p4 row-major: s_U[k * p4 + j] = U[k + j * d]  (16-byte aligned rows)
  double* s_tile = s_U + (size_t)d * p4;   // 2 x TC_PREP_THREADS x rs
  __shared__ double red[TC_PREP_THREADS / 32][TC_NBOUND];
  __shared__ int s_kend[16];               // per block of four columns of U: one past its last non-zero row
  for (int k = threadIdx.x; k < d; k += blockDim.x) s_mu[k] = mu[k];
  for (int e = threadIdx.x; e < d * p4; e += blockDim.x) {
    const int k = e / p4, j = e - k * p4;
    s_U[e] = (j < p) ? U[(size_t)j * d + k] : 0.0;
  }
  // the Cholesky scale matrix is upper triangular (reference src/joint_posterior.jl:56-76): its zero rows are skipped.
  // (Scanned from the shared copy: forty dependent global loads by a handful of threads used to open every block.)
  __syncthreads();
  if (threadIdx.x < p4 / 4) {
    int kend = 0;
    for (int k = 0; k < d; ++k) {
      const double* u = s_U + (size_t)k * p4 + 4 * threadIdx.x;
      if (u[0] != 0.0 || u[1] != 0.0 || u[2] != 0.0 || u[3] != 0.0) kend = k + 1;
    }
    s_kend[threadIdx.x] = kend;
  }
  double b_tmax = 0, b_a1 = 0, b_ref[TC_NORD] = {0}, b_max[TC_NORD] = {0}, b_r = 0, b_r2 = 0;
  double f_ref[TC_NFOLD] = {0}, f_max[TC_NFOLD] = {0};
  // two record tiles: the asynchronous copies (cp.async) of the next tile are in flight while the current one is consumed
  // flat coalesced copy; (row, column) of element e advance by (128 / ncols, 128 % ncols) per step: no division
  auto issue_tile = [&](long long base, int b) {
    const long long cnt = min((long long)TC_PREP_THREADS, N - base);
    if (cnt > 0) {
      const double* src = obs + (size_t)base * ncols;
      double* tl = s_tile + (size_t)b * TC_PREP_THREADS * rs;
      int row = threadIdx.x / ncols, col = threadIdx.x - row * ncols;
      const int drow = TC_PREP_THREADS / ncols, dcol = TC_PREP_THREADS - drow * ncols;
      for (int e = threadIdx.x; e < (int)cnt * ncols; e += TC_PREP_THREADS) {
        jp_cp_async8(tl + row * rs + col, src + e);
        row += drow;
        col += dcol;
        if (col >= ncols) { col -= ncols; ++row; }
      }
    }
    jp_cp_async_commit();
  };
  const long long step = (long long)gridDim.x * TC_PREP_THREADS;
  issue_tile((long long)blockIdx.x * TC_PREP_THREADS, 0);
  int cur = 0;
  for (long long base = (long long)blockIdx.x * TC_PREP_THREADS; base < N_pad; base += step, cur ^= 1) {
    jp_cp_async_wait_all();
    __syncthreads();     // tile `cur` has landed; the previous tile is consumed (and s_mu / s_U are visible)
    issue_tile(base + step, cur ^ 1);
    const long long i = base + threadIdx.x;
    float* o = coef + i;     // o[k * coef_stride]
    if (i >= N) {
      for (int k = 0; k < TC_COEF_ROWS; ++k) o[(size_t)k * coef_stride] = 0.f;
      continue;
    }
    const double* r = s_tile + ((size_t)cur * TC_PREP_THREADS + threadIdx.x) * rs;
    double eta = 0;
    for (int k = 0; k < d; ++k) eta += r[k] * s_mu[k];
    double t2 = 0;
    for (int j = 0; j < p4; j += 4) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const int kend = s_kend[j >> 2];
      for (int k = 0; k < kend; ++k) {
        const double xk = r[k];
        const double2 u01 = *reinterpret_cast<const double2*>(s_U + (size_t)k * p4 + j);
        const double2 u23 = *reinterpret_cast<const double2*>(s_U + (size_t)k * p4 + j + 2);
        a0 = fma(u01.x, xk, a0); a1 = fma(u01.y, xk, a1); a2 = fma(u23.x, xk, a2); a3 = fma(u23.y, xk, a3);
      }
      t2 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const double t = sqrt(t2);
    double c[TC_ORDER_MAX + 1];   // c[k], k = 3 .. TC_ORDER_MAX
    if (family == JP_FAM_LOGISTIC) {
      const double s = 1.0 / (1.0 + exp(-eta));
      for (int k = 3; k <= TC_ORDER_MAX; ++k) {
        double v = c_sp_poly[k][k];
        for (int j = k - 1; j >= 0; --j) v = fma(v, s, c_sp_poly[k][j]);
        c[k] = v * c_inv_fact[k];
      }
    } else {
      const double m = exp(eta);
      for (int k = 3; k <= TC_ORDER_MAX; ++k) c[k] = m * c_inv_fact[k];
    }
    for (int k = 0; k < TC_NCMAX; ++k) o[(size_t)k * coef_stride] = (float)c[3 + k];
    o[(size_t)TC_NCMAX * coef_stride] = __double2float_ru(t);   // rounded up: the fold interval may only grow
    // bounds
    b_tmax = fmax(b_tmax, t);
    b_a1 += fabs(c[3]) * t * t * t;
    const double tr = t * z_ref, tm = t * z_max;
    const double rho = (family == JP_FAM_LOGISTIC) ? 3.14159265358979323846 : 1e300;   // radius of convergence
    double sr = 0;
    for (int k = TC_ORDER_MAX; k >= 3; --k) sr = (sr + fabs(c[k])) * tr;
    sr *= tr * tr;                                     // majorant of |R_i| at |z| = z_ref
    b_r += sr;
    b_r2 += sr * sr;
    // tail <= |c_{K+1}| t^{K+1} + |c_{K+2}| t^{K+2} G(t): G = 1 / (1 - t / rho) majorises the remaining terms of the
    // logistic series (radius of convergence rho >= pi), e^t those of the Poisson series; powers by recurrence
    double gr, gm;
    if (family == JP_FAM_LOGISTIC) {
      gr = tr < rho ? 1.0 / (1.0 - tr / rho) : 1e300;
      gm = tm < rho ? 1.0 / (1.0 - tm / rho) : 1e300;
    } else {
      gr = exp(tr);
      gm = exp(tm);
    }
    double pr = tr * tr, pm = tm * tm;      // t^2
    pr *= pr * pr * tr;                     // t^7
    pm *= pm * pm * tm;
#pragma unroll
    for (int j = 0; j < TC_NORD; ++j) {
      const int K = 2 * j + 6;              // last order kept (NC = K - 2 coefficients); pr = tr^{K+1}
      b_ref[j] += fabs(c[K + 1]) * pr + fabs(c[K + 2]) * pr * tr * gr;
      b_max[j] += fabs(c[K + 1]) * pm + fabs(c[K + 2]) * pm * tm * gm;
      if (j < TC_NFOLD) {
        // economised: the two folded orders leave their minimax error (eps inside the fold interval, the growth
        // factor x^n beyond it), the orders after them are dropped as before
        const double r3 = pr * tr * tr, m3 = pm * tm * tm;   // t^{K+3}
        f_ref[j] += fabs(c[K + 1]) * pr * c_fold_eps[0][j] + fabs(c[K + 2]) * pr * tr * c_fold_eps[1][j] +
                    fabs(c[K + 3]) * r3 + fabs(c[K + 4]) * r3 * tr * gr;
        f_max[j] += fabs(c[K + 1]) * pm * c_fold_grow[0][j] + fabs(c[K + 2]) * pm * tm * c_fold_grow[1][j] +
                    fabs(c[K + 3]) * m3 + fabs(c[K + 4]) * m3 * tm * gm;
      }
      pr *= tr * tr;
      pm *= tm * tm;
    }
  }
  // block partials of all TC_NBOUND quantities in ONE pass: shuffle trees inside the warps (independent, so they pipeline), one
  // barrier, then thread q adds the warps' values of quantity q in warp order -- the same tree and order as 22 calls of
  // jp_block_sum / jp_block_max (bit-identical bounds), without their 66 barriers at the end of every block
  double vals[TC_NBOUND];
  vals[0] = b_tmax;
  vals[1] = b_a1;
#pragma unroll
  for (int j = 0; j < TC_NORD; ++j) {
    vals[2 + j] = b_ref[j];
    vals[7 + j] = b_max[j];
  }
  vals[12] = b_r;
  vals[13] = b_r2;
#pragma unroll
  for (int j = 0; j < TC_NFOLD; ++j) {
    vals[14 + j] = f_ref[j];
    vals[14 + TC_NFOLD + j] = f_max[j];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  vals[0] = jp_warp_max(vals[0]);
#pragma unroll
  for (int q = 1; q < TC_NBOUND; ++q) vals[q] = jp_warp_sum(vals[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < TC_NBOUND; ++q) red[wid][q] = vals[q];
  }
  __syncthreads();
  if (threadIdx.x < TC_NBOUND) {
    const int q = threadIdx.x;
    double r = (q == 0) ? red[0][0] : 0.0 + red[0][q];
    for (int w = 1; w < TC_PREP_THREADS / 32; ++w) r = (q == 0) ? fmax(r, red[w][q]) : r + red[w][q];
    bounds[(size_t)blockIdx.x * TC_NBOUND + q] = r;
  }
}

// Economisation of the chosen series length (once NC is known): per observation, with a = t_i z_ref,
//   c_k' = c_k + kappa_k c_n a^(n-k),   n = NC + 3 for the odd orders k = 3, 5, .., NC + 1 (rows 0, 2, ..),
//                                       n = NC + 4 for the even orders k = 4, 6, .., NC + 2 (rows 1, 3, ..)
// (tools/gen_fold.py: x^n ~ sum_k kappa_k x^k on |x| <= 1 in the span the kernel can evaluate).  In place on the
// first NC coefficient rows; streams (NC + 3) rows in and NC rows out, 4 bytes per observation and row.
__device__ __forceinline__ void tc_fold_body(int NC, long long n, long long N_pad, double z_ref, float* __restrict__ coef) {
  const int j = (NC - 4) >> 1, h = NC >> 1;
  // n observations starting at coef; N_pad is the row stride of the whole array
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float* o = coef + i;
    const double a = (double)o[(size_t)TC_NCMAX * N_pad] * z_ref, a2 = a * a;
    const double cn_o = o[(size_t)NC * N_pad], cn_e = o[(size_t)(NC + 1) * N_pad];
    double pw = a2;                      // a^(NC - 2m), m descending from NC/2 - 1
    for (int m = h - 1; m >= 0; --m) {
      float* ro = o + (size_t)(2 * m) * N_pad;
      float* re = o + (size_t)(2 * m + 1) * N_pad;
      *ro = (float)((double)*ro + c_fold_kappa[0][j][m] * cn_o * pw);
      *re = (float)((double)*re + c_fold_kappa[1][j][m] * cn_e * pw);
      pw *= a2;
    }
  }
}
__global__ void __launch_bounds__(256)
tc_fold_kernel(int NC, long long n, long long N_pad, double z_ref, float* __restrict__ coef) {
  tc_fold_body(NC, n, N_pad, z_ref, coef);
}
// the same with the series length (and whether to fold at all) read from the device-side decision
__global__ void __launch_bounds__(256)
tc_fold_dev_kernel(const TcDecision* __restrict__ dec, long long n, long long N_pad, double z_ref, float* __restrict__ coef) {
  if (dec->NC < 4 || !dec->fold) return;
  tc_fold_body(dec->NC, n, N_pad, z_ref, coef);
}

// Per node: delta = U z, theta = mu + delta (all transforms are the identity on this path), in two kernels that run on
// different streams (both rebuild delta from the integer keys; it costs a few FMAs per non-zero coordinate):
//   PART 0  theta, and per mirror pair (grid nodes 2j-1, 2j; pair 0 is the origin) the operand row [d_hi | d_hi | d_lo | 0] of
//           the pair's FIRST member -- or, when the local range starts on a second member, of that member with the sign flipped
//           (delta(-z) = -delta(z) exactly: the sums are sign-symmetric in IEEE arithmetic, and so is the TF32 split).
//           Needs (mu, U) only: runs while the O(N) passes over the observations are still in flight.
//   PART 1  the FP64 quadratic part  L_hat + g.delta - 1/2 delta' H delta + prior(theta): needs the sums (g, H, L_hat).
template <int PART>
__global__ void __launch_bounds__(128)
tc_node_prep_kernel(int d, int p, int kp, int seg, int rule, long long M, long long m0, long long M_grid, long long j_lo,
                    const uint8_t* __restrict__ idx, const double* __restrict__ znodes, const double* __restrict__ mu,
                    const double* __restrict__ U, const double* __restrict__ sums, double prior_sd,
                    double* __restrict__ theta, double* __restrict__ quad, float* __restrict__ ds) {
  extern __shared__ double sh[];
  double* s_mu = sh;                  // d
  double* s_U = s_mu + d;             // d x p
  double* s_g = s_U + d * p;          // d
  double* s_H = s_g + d;              // d x d (full, symmetric)
  double* s_z = s_H + d * d;          // JP_RULE_NMAX
  double* s_dl = s_z + JP_RULE_NMAX;  // blockDim.x x d  (delta of each thread, strided by thread)
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    s_mu[k] = mu[k];
    if (PART == 1) s_g[k] = sums[k];
  }
  for (int k = threadIdx.x; k < d * p; k += blockDim.x) s_U[k] = U[k];
  for (int k = threadIdx.x; k < JP_RULE_NMAX; k += blockDim.x) s_z[k] = znodes[k];
  if (PART == 1)
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
      int r = e % d, c = e / d;
      int rr = min(r, c), cc = max(r, c);
      s_H[e] = sums[d + cc * (cc + 1) / 2 + rr];     // packed upper triangle, column by column
    }
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double* dl = s_dl + threadIdx.x;       // dl[k * blockDim.x]
  if (m < M) {
    for (int k = 0; k < d; ++k) dl[k * blockDim.x] = 0.0;
    for (int j = 0; j < p; ++j) {
      int key = idx[(size_t)j * M_grid + (m0 + m)];
      if (key != 0) {
        double z = s_z[key];
        for (int k = 0; k < d; ++k) dl[k * blockDim.x] += s_U[(size_t)j * d + k] * z;   // same order as the FP64 path
      }
    }
    if (PART == 0) {
      for (int k = 0; k < d; ++k) theta[(size_t)k * M + m] = s_mu[k] + dl[k * blockDim.x];
    } else {
      const double L_hat = sums[d + d * (d + 1) / 2];
      double lin = 0, qf = 0, prior = 0;
      for (int k = 0; k < d; ++k) {
        const double dk = dl[k * blockDim.x];
        const double th = s_mu[k] + dk;
        lin += s_g[k] * dk;
        double hk = 0;
        for (int l = 0; l < d; ++l) hk += s_H[(size_t)k * d + l] * dl[l * blockDim.x];
        qf += dk * hk;
        const double zz = th / prior_sd;
        prior += -0.5 * zz * zz - log(prior_sd) - 0.5 * 1.8378770664093454835606594728112;
      }
      quad[m] = (L_hat + lin - 0.5 * qf) + prior;
    }
  }
  if (PART == 1) return;
  // The pair operand rows, written by WHOLE WARPS: a row is kp floats (one to three 128-byte lines), lane c takes column
  // c, c + 32, ..; one store instruction is one coalesced line.  (One thread writing its own row cost 3 d scattered
  // 4-byte stores per node, eight times the bytes in 32-byte sectors.)  Warp w serves the nodes of its own 32 threads.
  __syncthreads();
  const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;
  for (int r = w0; r < w0 + 32; ++r) {
    const long long mr = (long long)blockIdx.x * blockDim.x + r;
    if (mr >= M) break;
    const long long gm = m0 + mr;
    const bool first_member = (gm & 1) || gm == 0;
    if (!(first_member || mr == 0)) continue;      // a second member writes only when the local range starts on it
    const float sgn = first_member ? 1.f : -1.f;
    float* o = ds + (size_t)(((gm + 1) >> 1) - j_lo) * kp;
    for (int c = lane; c < kp; c += 32) {
      const int part = c / seg, k = c - part * seg;          // seg = d (compact rows) or one K atom (split rows)
      float v = 0.f;
      if (part < 3 && k < d) {
        float hi, lo;
        tf32_split(s_dl[(size_t)k * blockDim.x + r], hi, lo);
        v = sgn * (part == 2 ? lo : hi);
      }
      o[c] = v;
    }
  }
}

// ld = quad - sum_c (E[c] +- O[c]) + neg_min ; a = ld + |z|^2/2.  The first member of a mirror pair takes E + O,
// the second E - O (R(-D) = E(D) - O(D)).
__global__ void tc_finish_kernel(long long M, long long m0, long long j_lo, long long P, int chunks,
                                 const double* __restrict__ part, const double* __restrict__ quad,
                                 const double* __restrict__ hzz, double neg_min, double* __restrict__ logdens,
                                 double* __restrict__ a) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long gm = m0 + m, j = ((gm + 1) >> 1) - j_lo;
  const double sgn = ((gm & 1) || gm == 0) ? 1.0 : -1.0;
  double s = 0;
  for (int c = 0; c < chunks; ++c) {
    const double2 eo = *reinterpret_cast<const double2*>(part + ((size_t)c * P + j) * 2);
    s += eo.x + sgn * eo.y;
  }
  double ld = (quad[m] - s) + neg_min;
  logdens[m] = ld;
  a[m] = ld + hzz[m0 + m];
}

// ------------------------------------------------------------------------------------ the tensor-core kernel
struct TcKernelParams {
  int ka;                 // 128-byte K atoms per observation operand row
  int kb;                 // ... per pair operand row (= products of one contraction); kb > ka: split layout
  int stages;             // observation-tile ring depth
  int nbbuf;              // pair-operand buffers (2: the next item's operand loads under the current item)
  int n_pair_tiles;
  int chunks;             // observation chunks (work items = n_pair_tiles x chunks)
  int tiles_per_chunk;
  int n_obs_tiles;
  long long P;            // local mirror pairs
  const float* coef;      // coefficient row k of tile t: coef + slice * coef_slice_stride + k * coef_n_loc + (tile in slice) * 128,
  long long coef_n_loc;   //   slice = t / tiles_per_slice (one GPU: a single slice = the slab itself)
  long long coef_slice_stride;
  int tiles_per_slice;
  double* part;           // [chunks][P][2]: even and odd part of the remainder sum of a pair
  const int* nc_sel;      // non-null: the series length chosen on the device; an instantiation with another NC exits at once
  long long* dbg;         // profiling builds, MODE 5: per-tile clock64 stamps of CTA 0 ([6][TC_DBG_TILES]), else null
};
#define TC_DBG_TILES 2048
// MODE 5 timeline rows: 0 producer issues the tile's TMA, 1 MMA issuer starts waiting for the tile's accumulator buffer,
// 2 MMA issued, 3 epilogue warp 2 starts waiting for the tile's accumulator, 4 its wait returns
#define TC_STAMP(row, n) do { if (MODE == 5 && P.dbg && blockIdx.x == 0 && (n) < TC_DBG_TILES) P.dbg[(row) * TC_DBG_TILES + (n)] = clock64(); } while (0)

// Packed FP32 arithmetic (Blackwell FFMA2 / FMUL2): one instruction works on two adjacent pair columns, which
// halves the issue slots of the epilogue (the FMA pipe itself retires 32 lanes x 2 per two cycles either way).
__device__ __forceinline__ uint64_t f32x2_pack(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f32x2_bcast(float x) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void f32x2_fma_acc(uint64_t& acc, uint64_t a, uint64_t b) {   // acc += a * b, in place
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint64_t f32x2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Mirror pairs.  A column of the accumulator tile is a PAIR of grid nodes (z, -z) (adjacent in the grid's mirror
// order, jp_grid.cu), represented by D = x_i . delta(z); the partner sees -D.  Splitting the remainder series by
// parity,   R_i(+-D) = E_i(D) +- O_i(D),   E = D^4 (c_4 + c_6 D^2 + ...),   O = D^3 (c_3 + c_5 D^2 + ...),
// one pass over D yields both nodes: NC + 3 packed operations per column instead of 2 (NC + 2), and half the
// contraction, TMEM reads and operand traffic.  accE / accO: W columns = W / 2 packed pairs each; c[k] = c_{3+k}.
template <int NC, int MODE, int W>
__device__ __forceinline__ void tc_accumulate(const uint32_t (&v)[W], const float (&c)[NC], uint64_t* accE, uint64_t* accO) {
  // coefficients stay scalar: ptxas folds the (c, c) pack into FFMA2's broadcast operand form (Rx.F32), which
  // reads one register instead of a pair
  if (MODE == 1) {   // profiling aid: TMEM traffic without the series arithmetic (results are meaningless)
#pragma unroll
    for (int j = 0; j < W / 2; ++j) accE[j] = f32x2_fma(f32x2_pack(v[2 * j], v[2 * j + 1]), f32x2_bcast(c[0]), accE[j]);
    return;
  }
#pragma unroll
  for (int j = 0; j < W / 2; ++j) {
    const uint64_t D = f32x2_pack(v[2 * j], v[2 * j + 1]);
    const uint64_t D2 = f32x2_mul(D, D);
    uint64_t pe = f32x2_bcast(c[NC - 1]), po = f32x2_bcast(c[NC - 2]);   // NC even: c_{NC+2} is an even order
#pragma unroll
    for (int m = NC / 2 - 2; m >= 0; --m) {
      pe = f32x2_fma(pe, D2, f32x2_bcast(c[2 * m + 1]));
      po = f32x2_fma(po, D2, f32x2_bcast(c[2 * m]));
    }
    f32x2_fma_acc(accE[j], f32x2_mul(D2, D2), pe);
    f32x2_fma_acc(accO[j], f32x2_mul(D2, D), po);
  }
}

// 32 x 32 transpose-reduce: on exit v[0] of lane l holds sum over lanes of the entry v[l]
__device__ __forceinline__ float tc_transpose_reduce32(float* v, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = up ? v[j] : v[j + s];
      const float keep = up ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// MODE 0 is the product; 1 (no series arithmetic) and 2 (no TMEM loads) exist only to attribute time to the
// two resources the epilogue contends for, 3 and 4 repeat 0 and 2 with two extra (overwritten) K-atom products per
// tile to measure how MMA time composes with epilogue time (JP_TC_DEBUG_MODE, profiling builds of bench.py only)
template <int NC, int MODE>
__device__ __forceinline__ void tc_kernel_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcKernelParams& P) {
  extern __shared__ uint8_t smem_raw[];
  // carve-up (1024-byte aligned for the 128-byte swizzle): pair operands, observation ring, reduction buffer, barriers
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = base;                                              // nbbuf x ka x (96 x 128 B)
  const uint32_t b_bytes = (uint32_t)P.kb * TC_PAIR_TILE * 128u;
  const uint32_t sA = sB + (uint32_t)P.nbbuf * b_bytes;                  // stages x ka x (128 x 128 B)
  const uint32_t a_bytes = (uint32_t)P.ka * TC_OBS_TILE * 128u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // per-observation series coefficients of the tiles in flight, [slot][k][128 observations]: a tile's slot lives from
  // its TMA load until the epilogue has consumed its accumulator, i.e. across the observation ring AND the TMEM
  // buffers, hence stages + TC_NBUF slots (a slot is reused only after empty[stage] of a tile `stages` later,
  // which the MMA issuer signals after waiting for tempty of a tile TC_NBUF earlier still)
#if TC_RING_UNROLL
  const int n_cslots = ((P.stages + TC_NBUF + 3) / 4) * 4;      // whole groups of four (the epilogue's unrolled ring)
#else
  const int n_cslots = P.stages + TC_NBUF;
#endif
  const uint32_t c_bytes = (uint32_t)NC * TC_OBS_TILE * 4u;
  const uint32_t sC = sA + (uint32_t)P.stages * a_bytes;
  double* red = reinterpret_cast<double*>(gen + (size_t)P.nbbuf * b_bytes + (size_t)P.stages * a_bytes +
                                          (size_t)n_cslots * c_bytes);   // [4][96][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 4 * TC_PAIR_TILE * 2);
  const uint32_t bar_full = smem_u32(bars);                 // [TC_MAX_STAGES]
  const uint32_t bar_empty = bar_full + 8u * TC_MAX_STAGES; // [TC_MAX_STAGES]
  const uint32_t bar_bfull = bar_empty + 8u * TC_MAX_STAGES;// [2]
  const uint32_t bar_bempty = bar_bfull + 16u;              // [2]
  const uint32_t bar_tfull = bar_bempty + 16u;              // [TC_NBUF]
  const uint32_t bar_tempty = bar_tfull + 8u * TC_NBUF;     // [TC_NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 4 + 2 * TC_NBUF);

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps role / address arithmetic
  // on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_bfull + 8u * b, 1);
      mbar_init(bar_bempty + 8u * b, 1);
    }
    for (int b = 0; b < TC_NBUF; ++b) {
      mbar_init(bar_tfull + 8u * b, 1);
      mbar_init(bar_tempty + 8u * b, TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns (TC_NBUF accumulator buffers of TC_PAIR_TILE columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = P.n_pair_tiles * P.chunks;
  // The producer and the issuer run their loops with the WHOLE warp converged and only predicate the asynchronous
  // instructions on one elected lane: addresses, descriptors, ring indices and phases are then provably warp-uniform
  // and live on the uniform datapath.  (Inside an `if (lane == 0)` region the compiler wraps every tcgen05.mma / TMA
  // in an elect + R2UR + vote loop, ~100 cycles of dependent latency each: the issuer then needs a full tile time for
  // its 4 .. 12 MMAs, runs in lockstep with the epilogue instead of TC_NBUF tiles ahead, and the MMA time shows up on
  // the critical path -- measured with the MODE 5 timeline, profiles/r01_tc_attribution.txt.)
  if (warp == 0) {
    // ===================================================== TMA producer
    const bool leader = elect_one();
    int stage = 0, bb = 0, cslot = 0, ntile = 0;
    uint32_t phase = 0, bphase = 0;   // bphase: bit b = parity of pair-operand buffer b
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / P.n_pair_tiles, pair_tile = item % P.n_pair_tiles;
      const int t0 = chunk * P.tiles_per_chunk, t1 = min(P.n_obs_tiles, t0 + P.tiles_per_chunk);
      mbar_wait_relaxed(bar_bempty + 8u * bb, ((bphase >> bb) & 1u) ^ 1u);
      if (leader) {
        mbar_expect_tx(bar_bfull + 8u * bb, b_bytes);
        for (int a = 0; a < P.kb; ++a)
          tma_load_2d(sB + (uint32_t)bb * b_bytes + (uint32_t)a * TC_PAIR_TILE * 128u, &tmB, bar_bfull + 8u * bb,
                      a * TC_KATOM, pair_tile * TC_PAIR_TILE);
      }
      bphase ^= 1u << bb;
      if (++bb == P.nbbuf) bb = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait_relaxed(bar_empty + 8u * stage, phase ^ 1);
        if (leader) {
          TC_STAMP(0, ntile);
          mbar_expect_tx(bar_full + 8u * stage, a_bytes + c_bytes);
          for (int a = 0; a < P.ka; ++a)
            tma_load_2d(sA + (uint32_t)stage * a_bytes + (uint32_t)a * TC_OBS_TILE * 128u, &tmA, bar_full + 8u * stage,
                        a * TC_KATOM, t * TC_OBS_TILE);
          const int sl = t / P.tiles_per_slice;
          // gathered rows are laid out [slice][NC][n_loc] with the series length actually in use (= this body's NC)
          const size_t slice_stride = P.coef_slice_stride ? (size_t)NC * (size_t)P.coef_n_loc : 0;
          const float* crow = P.coef + (size_t)sl * slice_stride + (size_t)(t - sl * P.tiles_per_slice) * TC_OBS_TILE;
#pragma unroll
          for (int k = 0; k < NC; ++k)
            bulk_load_1d(sC + (uint32_t)cslot * c_bytes + (uint32_t)k * TC_OBS_TILE * 4u, crow + (size_t)k * P.coef_n_loc,
                         TC_OBS_TILE * 4u, bar_full + 8u * stage);
        }
        ++ntile;
        if (++cslot == n_cslots) cslot = 0;
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    const bool leader = elect_one();
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 96 pairs, M = 128 observations
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_PAIR_TILE >> 3) << 17) |
                           ((uint32_t)(TC_OBS_TILE >> 4) << 24);
    int stage = 0, buf = 0, bb = 0, ntile = 0;
    uint32_t phase = 0, bphase = 0, tphase = 0;   // tphase: bit b = parity of accumulator buffer b
    const uint32_t loA0 = umma_desc_lo(sA), loB0 = umma_desc_lo(sB), a_step = a_bytes >> 4, b_step = b_bytes >> 4;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / P.n_pair_tiles;
      const int t0 = chunk * P.tiles_per_chunk, t1 = min(P.n_obs_tiles, t0 + P.tiles_per_chunk);
      mbar_wait_relaxed(bar_bfull + 8u * bb, (bphase >> bb) & 1u);
      const uint32_t loB = loB0 + (uint32_t)bb * b_step;
      for (int t = t0; t < t1; ++t) {
        if (leader) TC_STAMP(1, ntile);
        mbar_wait_relaxed(bar_tempty + 8u * buf, ((tphase >> buf) & 1u) ^ 1u);
        if (leader) TC_STAMP(5, ntile);      // row 5: accumulator buffer free, now waiting for the observation tile
        mbar_wait_relaxed(bar_full + 8u * stage, phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * TC_TMEM_STRIDE;
        const uint32_t loA = loA0 + (uint32_t)stage * a_step;
        if (leader) {
          TC_STAMP(2, ntile);
          if (MODE == 3 || MODE == 4) {   // two dummy atom products, overwritten by the real contraction below
            for (int r = 0; r < 8; ++r) umma_tf32_lo(d_tmem, loA + 2u * (r & 3), loB + 2u * (r & 3), idesc, 0u);
          }
#pragma unroll 1
          for (int a = 0; a < P.kb; ++a) {
            // split layout: the last product (x_hi . d_lo) takes the observation row's first atom again
            const uint32_t la = loA + ((a < P.ka) ? (uint32_t)a * (TC_OBS_TILE * 128u >> 4) : 0u);
            const uint32_t lb = loB + (uint32_t)a * (TC_PAIR_TILE * 128u >> 4);
#pragma unroll
            for (int j = 0; j < 4; ++j)   // K = 8 tf32 = 32 bytes (2 descriptor units) per instruction inside the atom
              umma_tf32_lo(d_tmem, la + 2u * j, lb + 2u * j, idesc, (a | j) ? 1u : 0u);
          }
          tc_commit(bar_empty + 8u * stage);        // frees the observation stage when the MMAs have read it
          tc_commit(bar_tfull + 8u * buf);          // publishes the accumulator buffer
        }
        ++ntile;
        tphase ^= 1u << buf;
        if (++buf == TC_NBUF) buf = 0;
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      if (leader) tc_commit(bar_bempty + 8u * bb);  // pair operand may be overwritten once every MMA has retired
      bphase ^= 1u << bb;
      if (++bb == P.nbbuf) bb = 0;
    }
  } else {
    // ===================================================== epilogue warps
#if TC_RING_UNROLL
    // The accumulator ring has FOUR buffers and the tile loop is unrolled by four, entered at the buffer the CTA's running tile
    // count points at (Duff's device): TMEM addresses, barrier addresses and the coefficient-set ping-pong are compile-time in
    // every copy of the step, the phase parity flips and the coefficient-slot group advances once per round of four tiles.
    // (With five buffers and run-time ring indices the step carried ~20 uniform-datapath instructions per tile and warp; every
    // one of them costs an issue slot the packed FP32 operations cannot hide.)
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;           // column third handled by this warp
    const int et = threadIdx.x - 64;         // 0 .. TC_EPI_THREADS - 1
    uint64_t accE[TC_COLS_PER_WARP / 2], accO[TC_COLS_PER_WARP / 2];   // FP32 sums, packed in pairs of adjacent columns
#pragma unroll
    for (int j = 0; j < TC_COLS_PER_WARP / 2; ++j) accE[j] = accO[j] = 0ull;
    int ntile = 0;                           // tiles this CTA has consumed: buffer = ntile & 3, round parity = (ntile >> 2) & 1
    uint32_t par = 0;                        // phase parity of the current round of accumulator buffers
    uint32_t cgrp = 0;                       // byte offset of the current group of four coefficient slots
    const uint32_t cgrp_end = (uint32_t)(n_cslots / 4) * 4u * c_bytes;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * TC_COLS_PER_WARP);
    const uint32_t coef_addr = sC + (uint32_t)(q * 32 + lane) * 4u;   // this thread's observation: tile row = TMEM lane
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / P.n_pair_tiles, pair_tile = item % P.n_pair_tiles;
      const int t0 = chunk * P.tiles_per_chunk, t1 = min(P.n_obs_tiles, t0 + P.tiles_per_chunk);
      uint32_t va[TC_LDW], vb[TC_LDW];
      float ca[NC], cb[NC];
      auto load_coef = [&](float (&dst)[NC], uint32_t slot_addr) {
#pragma unroll
        for (int k = 0; k < NC; ++k)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dst[k]) : "r"(slot_addr + (uint32_t)k * TC_OBS_TILE * 4u));
      };
      int t = t0;
      // one tile in accumulator buffer B (compile-time); cc = its coefficients, cn receives the next tile's
      auto tile_step = [&](auto Bc, const float (&cc)[NC], float (&cn)[NC]) {
        constexpr int B = decltype(Bc)::value, NB = (B + 1) & 3;
        const bool more = t + 1 < t1;
        const uint32_t taddr = lane_addr + (uint32_t)(B * TC_TMEM_STRIDE);
        const uint32_t npar = (B == 3) ? (par ^ 1u) : par;                     // parity of the next tile's phase
        const uint32_t ncg = (B == 3) ? ((cgrp + 4u * c_bytes == cgrp_end) ? 0u : cgrp + 4u * c_bytes) : cgrp;
#if TC_EARLY_PEEK
        const bool ready = more && mbar_test_wait(bar_tfull + 8u * NB, npar);
#else
        const bool ready = false;
#endif
#pragma unroll
        for (int c = 0; c < TC_COLS_PER_WARP / TC_LDW; ++c) {
          uint32_t(&cur)[TC_LDW] = (c & 1) ? vb : va;
          uint32_t(&nxt)[TC_LDW] = (c & 1) ? va : vb;
          tmem_ld_wait(cur);
          if (c + 1 < TC_COLS_PER_WARP / TC_LDW) {
            if (MODE != 2 && MODE != 4) tmem_ld(nxt, taddr + (uint32_t)(TC_LDW * (c + 1)));
          } else {
            // every tcgen05.ld of this tile has completed: hand the accumulator buffer back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8u * B);
            if (more) {
              if (warp == 2 && lane == 0) TC_STAMP(3, ntile + 1);
              if (!ready) mbar_wait(bar_tfull + 8u * NB, npar);
              if (warp == 2 && lane == 0) TC_STAMP(4, ntile + 1);
              tc_fence_after();
              if (MODE != 2 && MODE != 4) tmem_ld(nxt, lane_addr + (uint32_t)(NB * TC_TMEM_STRIDE));
              load_coef(cn, coef_addr + ncg + (uint32_t)NB * c_bytes);
            }
          }
          tc_accumulate<NC, MODE, TC_LDW>(cur, cc, accE + c * (TC_LDW / 2), accO + c * (TC_LDW / 2));
        }
        if (B == 3) {
          par ^= 1u;
          cgrp = ncg;
        }
        ++ntile;
        ++t;
      };
      // first tile of the item
      {
        const int b0 = ntile & 3;
        if (warp == 2 && lane == 0) TC_STAMP(3, ntile);
        mbar_wait(bar_tfull + 8u * b0, par);
        if (warp == 2 && lane == 0) TC_STAMP(4, ntile);
        tc_fence_after();
        tmem_ld(va, lane_addr + (uint32_t)(b0 * TC_TMEM_STRIDE));
        if (b0 & 1) load_coef(cb, coef_addr + cgrp + (uint32_t)b0 * c_bytes);
        else load_coef(ca, coef_addr + cgrp + (uint32_t)b0 * c_bytes);
        switch (b0) {
          for (;;) {
            case 0: tile_step(std::integral_constant<int, 0>(), ca, cb); if (t == t1) break;
            case 1: tile_step(std::integral_constant<int, 1>(), cb, ca); if (t == t1) break;
            case 2: tile_step(std::integral_constant<int, 2>(), ca, cb); if (t == t1) break;
            case 3: tile_step(std::integral_constant<int, 3>(), cb, ca); if (t == t1) break;
          }
        }
      }
      // flush the item: sum over the 32 observation lanes by transpose-reduce, over the 4 lane quarters in
      // shared memory (FP64), one (even, odd) partial per (chunk, pair)
      {
        float col[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) f32x2_unpack(accE[j], col[2 * j], col[2 * j + 1]);
        const float sE = tc_transpose_reduce32(col, lane);
#pragma unroll
        for (int j = 0; j < 16; ++j) f32x2_unpack(accO[j], col[2 * j], col[2 * j + 1]);
        const float sO = tc_transpose_reduce32(col, lane);
        double* r = red + ((size_t)q * TC_PAIR_TILE + h * TC_COLS_PER_WARP + lane) * 2;
        r[0] = (double)sE;
        r[1] = (double)sO;
      }
      epi_bar_sync();
      if (et < 2 * TC_PAIR_TILE) {
        const double s = (red[et] + red[2 * TC_PAIR_TILE + et]) + (red[4 * TC_PAIR_TILE + et] + red[6 * TC_PAIR_TILE + et]);
        const long long pair = (long long)pair_tile * TC_PAIR_TILE + (et >> 1);
        if (pair < P.P) P.part[((size_t)chunk * P.P + pair) * 2 + (et & 1)] = s;
      }
      epi_bar_sync();
#pragma unroll
      for (int j = 0; j < TC_COLS_PER_WARP / 2; ++j) accE[j] = accO[j] = 0ull;
    }
  }
#else
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;           // column third handled by this warp
    const int et = threadIdx.x - 64;         // 0 .. TC_EPI_THREADS - 1
    uint64_t accE[TC_COLS_PER_WARP / 2], accO[TC_COLS_PER_WARP / 2];   // FP32 sums, packed in pairs of adjacent columns
#pragma unroll
    for (int j = 0; j < TC_COLS_PER_WARP / 2; ++j) accE[j] = accO[j] = 0ull;
    int buf = 0, cslot = 0, ntile = 0;
    uint32_t tphase = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * TC_COLS_PER_WARP);
    const uint32_t coef_addr = sC + (uint32_t)(q * 32 + lane) * 4u;   // this thread's observation: tile row = TMEM lane
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / P.n_pair_tiles, pair_tile = item % P.n_pair_tiles;
      const int t0 = chunk * P.tiles_per_chunk, t1 = min(P.n_obs_tiles, t0 + P.tiles_per_chunk);
      // Software pipeline over (tile, 16-column chunk): the tcgen05.ld of the next chunk -- across a tile
      // boundary too, the next accumulator buffer is normally complete long before -- and the coefficient
      // loads of the next tile are in flight while the current chunk is evaluated.  The tile loop is unrolled
      // by two so that the two coefficient sets alternate without register copies.
      uint32_t va[TC_LDW], vb[TC_LDW];
      float ca[NC], cb[NC];
      // coefficients of the tile whose accumulator was just seen full (their TMA copy completed before its MMA ran)
      auto load_coef = [&](float (&dst)[NC]) {
        const uint32_t a = coef_addr + (uint32_t)cslot * c_bytes;
#pragma unroll
        for (int k = 0; k < NC; ++k)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dst[k]) : "r"(a + (uint32_t)k * TC_OBS_TILE * 4u));
        if (++cslot == n_cslots) cslot = 0;
      };
      auto tile_step = [&](int t, const float (&cc)[NC], float (&cn)[NC]) {
        const bool more = t + 1 < t1;
        const int nbuf = (buf + 1 == TC_NBUF) ? 0 : buf + 1;
        const uint32_t taddr = lane_addr + (uint32_t)buf * TC_TMEM_STRIDE;
#if TC_EARLY_PEEK
        // The MMA issuer runs one to two tiles ahead: the next accumulator is normally complete when this tile starts.
        // Peeking now (result in a register) takes the ~100-cycle barrier read off the tile boundary, where it used to
        // sit in front of the last chunk's arithmetic in all three warps of the sub-partition at once.
        const bool ready = more && mbar_test_wait(bar_tfull + 8u * nbuf, (tphase >> nbuf) & 1u);
#else
        const bool ready = false;
#endif
#pragma unroll
        for (int c = 0; c < TC_COLS_PER_WARP / TC_LDW; ++c) {
          uint32_t(&cur)[TC_LDW] = (c & 1) ? vb : va;
          uint32_t(&nxt)[TC_LDW] = (c & 1) ? va : vb;
          tmem_ld_wait(cur);
          if (c + 1 < TC_COLS_PER_WARP / TC_LDW) {
            if (MODE != 2 && MODE != 4) tmem_ld(nxt, taddr + (uint32_t)(TC_LDW * (c + 1)));
          } else {
            // every tcgen05.ld of this tile has completed: hand the accumulator buffer back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8u * buf);
            if (more) {
              if (warp == 2 && lane == 0) TC_STAMP(3, ntile + 1);
              if (!ready) mbar_wait(bar_tfull + 8u * nbuf, (tphase >> nbuf) & 1u);
              if (warp == 2 && lane == 0) TC_STAMP(4, ntile + 1);
              tphase ^= 1u << nbuf;
              tc_fence_after();
              if (MODE != 2 && MODE != 4) tmem_ld(nxt, lane_addr + (uint32_t)nbuf * TC_TMEM_STRIDE);
              load_coef(cn);
            }
          }
          tc_accumulate<NC, MODE, TC_LDW>(cur, cc, accE + c * (TC_LDW / 2), accO + c * (TC_LDW / 2));
        }
        buf = nbuf;
        ++ntile;
      };
      int t = t0;
      const bool odd = ((t1 - t0) & 1) != 0;
      if (warp == 2 && lane == 0) TC_STAMP(3, ntile);
      mbar_wait(bar_tfull + 8u * buf, (tphase >> buf) & 1u);
      if (warp == 2 && lane == 0) TC_STAMP(4, ntile);
      tphase ^= 1u << buf;
      tc_fence_after();
      tmem_ld(va, lane_addr + (uint32_t)buf * TC_TMEM_STRIDE);
      if (odd) load_coef(cb); else load_coef(ca);
      if (odd) {       // peel one tile so that the main loop is two straight-line steps
        tile_step(t, cb, ca);
        ++t;
      }
      for (; t < t1; t += 2) {
        tile_step(t, ca, cb);
        tile_step(t + 1, cb, ca);
      }
      // flush the item: sum over the 32 observation lanes by transpose-reduce, over the 4 lane quarters in
      // shared memory (FP64), one (even, odd) partial per (chunk, pair)
      {
        float col[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) f32x2_unpack(accE[j], col[2 * j], col[2 * j + 1]);
        const float sE = tc_transpose_reduce32(col, lane);
#pragma unroll
        for (int j = 0; j < 16; ++j) f32x2_unpack(accO[j], col[2 * j], col[2 * j + 1]);
        const float sO = tc_transpose_reduce32(col, lane);
        double* r = red + ((size_t)q * TC_PAIR_TILE + h * TC_COLS_PER_WARP + lane) * 2;
        r[0] = (double)sE;
        r[1] = (double)sO;
      }
      epi_bar_sync();
      if (et < 2 * TC_PAIR_TILE) {
        const double s = (red[et] + red[2 * TC_PAIR_TILE + et]) + (red[4 * TC_PAIR_TILE + et] + red[6 * TC_PAIR_TILE + et]);
        const long long pair = (long long)pair_tile * TC_PAIR_TILE + (et >> 1);
        if (pair < P.P) P.part[((size_t)chunk * P.P + pair) * 2 + (et & 1)] = s;
      }
      epi_bar_sync();
#pragma unroll
      for (int j = 0; j < TC_COLS_PER_WARP / 2; ++j) accE[j] = accO[j] = 0ull;
    }
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int NC, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
jp_glm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcKernelParams P) {
  if (P.nc_sel != nullptr && *P.nc_sel != NC) return;   // device-side series-length decision: uniform over the grid
  tc_kernel_body<NC, MODE>(tmA, tmB, P);
}

// Device-side decision, second launch: whichever of the longer series (NC = 6 .. 12) the device chose -- ONE launch instead
// of four that exit at once (an empty launch of this 448-thread, 150+ KB kernel costs ~6 us).  Its ring geometry is the one
// computed for NC = 12 (the largest coefficient slots), valid for the shorter series too.
__global__ void __launch_bounds__(TC_THREADS, 1)
jp_glm_tc_rest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcKernelParams P) {
  const int nc = *P.nc_sel;
  if (nc == 6) tc_kernel_body<6, 0>(tmA, tmB, P);
  else if (nc == 8) tc_kernel_body<8, 0>(tmA, tmB, P);
  else if (nc == 10) tc_kernel_body<10, 0>(tmA, tmB, P);
  else if (nc == 12) tc_kernel_body<12, 0>(tmA, tmB, P);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tensor_map(CUtensorMap* map, float* base, long long rows, int kp, int box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    JP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    if (!p || qr != cudaDriverEntryPointSuccess) {
      jp_set_error("cuTensorMapEncodeTiled is not available from this driver");
      return JP_ERR_UNSUPPORTED;
    }
    fn = (PFN_encodeTiled)p;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)kp * sizeof(float)};
  cuuint32_t box[2] = {TC_KATOM, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    jp_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld kp=%d box_rows=%d)", (int)r, rows, kp, box_rows);
    return JP_ERR_CUDA;
  }
  return JP_OK;
}

static int upload_tables(jp_ctx* ctx) {
  if (ctx->tc_tables_uploaded) return JP_OK;
  double poly[TC_ORDER_MAX + 1][TC_ORDER_MAX + 2] = {{0}};
  double inv_fact[TC_ORDER_MAX + 2];
  // f_1 = s ; f_{k+1} = f_k'(s) (s - s^2), exact in integers (|coefficients| < 2^53 up to order 16)
  std::vector<long long> f = {0, 1};
  for (int k = 1; k <= TC_ORDER_MAX; ++k) {
    for (size_t j = 0; j < f.size(); ++j) poly[k][j] = (double)f[j];
    std::vector<long long> df(f.size() - 1);
    for (size_t j = 1; j < f.size(); ++j) df[j - 1] = (long long)j * f[j];
    std::vector<long long> nx(df.size() + 2, 0);
    for (size_t j = 0; j < df.size(); ++j) {
      nx[j + 1] += df[j];
      nx[j + 2] -= df[j];
    }
    f.swap(nx);
  }
  double fct = 1;
  inv_fact[0] = 1;
  for (int k = 1; k <= TC_ORDER_MAX + 1; ++k) {
    fct *= k;
    inv_fact[k] = 1.0 / fct;
  }
  JP_CUDA(cudaMemcpyToSymbol(c_sp_poly, poly, sizeof poly));
  JP_CUDA(cudaMemcpyToSymbol(c_inv_fact, inv_fact, sizeof inv_fact));
  double kappa[2][TC_NFOLD][5], eps[2][TC_NFOLD], grow[2][TC_NFOLD];
  for (int j = 0; j < TC_NFOLD; ++j) {
    for (int m = 0; m < 5; ++m) {
      kappa[0][j][m] = k_fold_odd[j][m];
      kappa[1][j][m] = k_fold_even[j][m];
    }
    eps[0][j] = k_fold_eps_odd[j]; eps[1][j] = k_fold_eps_even[j];
    grow[0][j] = k_fold_grow_odd[j]; grow[1][j] = k_fold_grow_even[j];
  }
  JP_CUDA(cudaMemcpyToSymbol(c_fold_kappa, kappa, sizeof kappa));
  JP_CUDA(cudaMemcpyToSymbol(c_fold_eps, eps, sizeof eps));
  JP_CUDA(cudaMemcpyToSymbol(c_fold_grow, grow, sizeof grow));
  ctx->tc_tables_uploaded = true;
  return JP_OK;
}

static bool tc_static_ok(const jp_posterior* post, const jp_fit_args* args) {
  const jp_data* data = post->data;
  if (data->family != JP_FAM_LOGISTIC && data->family != JP_FAM_POISSON) {
    jp_set_error("tensor-core path: family %d is not a GLM", data->family);
    return false;
  }
  if (3 * args->d > 3 * TC_KATOM) {   // three 128-byte K atoms per operand row fit the shared-memory budget
    jp_set_error("tensor-core path: d=%d exceeds %d", args->d, TC_KATOM);
    return false;
  }
  for (int k = 0; k < args->d; ++k)
    if (args->h_transform[k] != JP_T_REAL) {
      jp_set_error("tensor-core path: coordinate %d is constrained", k);
      return false;
    }
  if (data->ncols != args->d + 1) {
    jp_set_error("tensor-core path: %d columns for d=%d", data->ncols, args->d);
    return false;
  }
  return true;
}

bool jp_fit_tc_supported(const jp_posterior* post, const jp_fit_args* args) { return tc_static_ok(post, args); }

void jp_tc_data_free(jp_data* data) {
  TcDataState* s = static_cast<TcDataState*>(data->tc_state);
  if (!s) return;
  jp_dfree(data->ctx, s->d_xs); jp_dfree(data->ctx, s->d_coef); jp_dfree(data->ctx, s->d_coef_all); jp_dfree(data->ctx, s->d_sums);
  jp_dfree(data->ctx, s->d_work); jp_dfree(data->ctx, s->d_bounds); jp_dfree(data->ctx, s->d_comb); jp_dfree(data->ctx, s->d_loc);
  if (s->ev_bounds) cudaEventDestroy(s->ev_bounds);
  delete s;
  data->tc_state = nullptr;
}
void jp_tc_post_free(jp_posterior* post) {
  TcPostState* s = static_cast<TcPostState*>(post->tc_state);
  if (!s) return;
  jp_dfree(post->ctx, s->d_ds); jp_dfree(post->ctx, s->d_quad); jp_dfree(post->ctx, s->d_part); jp_dfree(post->ctx, s->d_dec);
  delete s;
  post->tc_state = nullptr;
}

static int ensure_data_state(jp_ctx* ctx, jp_data* data, int d, int world, int rank) {
  if (data->tc_state) {
    const TcDataState* o = static_cast<const TcDataState*>(data->tc_state);
    if (o->world != world || o->rank != rank) jp_tc_data_free(data);   // re-sliced
  }
  if (data->tc_state) return JP_OK;
  TcDataState* s = new TcDataState();
  data->tc_state = s;
  JP_CUDA(cudaEventCreateWithFlags(&s->ev_bounds, cudaEventDisableTiming));
  s->d = d;
  s->split = (3 * d > 2 * TC_KATOM) ? 1 : 0;     // three atoms compact -> two atoms split (d <= TC_KATOM is checked)
  s->kp = s->split ? 2 * TC_KATOM : ((3 * d + TC_KATOM - 1) / TC_KATOM) * TC_KATOM;
  s->ka = s->kp / TC_KATOM;
  s->kp_b = s->split ? 3 * TC_KATOM : s->kp;
  s->N = data->N;
  // observation slices of the sharded prep: world equal slices of whole tiles (the last ones may be short or empty)
  s->world = world;
  s->rank = rank;
  const long long tiles = (data->N + TC_OBS_TILE - 1) / TC_OBS_TILE;
  s->n_loc = ((tiles + world - 1) / world) * TC_OBS_TILE;
  s->N_pad = s->n_loc * world;
  const int nE = d + d * (d + 1) / 2;
  s->glm_blocks = jp_glm_num_blocks(ctx, data->N);
  JP_CUDA(jp_dmalloc(ctx, &s->d_xs, (size_t)s->N_pad * s->kp * sizeof(float)));
  JP_CUDA(jp_dmalloc(ctx, &s->d_coef, (size_t)s->n_loc * TC_COEF_ROWS * sizeof(float)));
  JP_CUDA(jp_dmalloc(ctx, &s->d_comb, (size_t)TC_NBOUND * 8));
  JP_CUDA(jp_dmalloc(ctx, &s->d_sums, (size_t)(nE + 1) * 8));
  JP_CUDA(jp_dmalloc(ctx, &s->d_work, (size_t)s->glm_blocks * (nE + 1) * 8));
  JP_CUDA(jp_dmalloc(ctx, &s->d_bounds, (size_t)TC_PREP_BLOCKS_MAX * TC_NBOUND * 8));
  {
    const long long slice_tiles = s->n_loc / TC_OBS_TILE;
    s->prep_blocks = (int)(slice_tiles <= TC_PREP_BLOCKS_MAX ? std::max<long long>(1, slice_tiles) : TC_PREP_BLOCKS);
  }
  const unsigned split_blocks = (unsigned)std::min<long long>((s->N_pad + TC_SPLIT_ROWS - 1) / TC_SPLIT_ROWS, (long long)ctx->sm_count * 32);
  tc_split_x_kernel<<<split_blocks, dim3(s->kp, TC_SPLIT_ROWS), 0, ctx->stream>>>(d, data->ncols, s->kp, s->split, s->N, s->N_pad,
                                                                                  data->d_obs, s->d_xs);
  JP_CHECK_LAUNCH(ctx);
  JP_TRY(make_tensor_map(&s->tmA, s->d_xs, s->N_pad, s->kp, TC_OBS_TILE));
  return JP_OK;
}

static int ensure_post_state(jp_posterior* post, int kp) {
  if (post->tc_state) return JP_OK;
  TcPostState* s = new TcPostState();
  post->tc_state = s;
  s->kp = kp;
  // mirror pairs (grid nodes 2j-1, 2j; pair 0 = the origin) touched by the local node range [m0, m0 + M)
  s->j_lo = (post->m0 + 1) >> 1;
  s->P = ((post->m0 + post->M) >> 1) - s->j_lo + 1;
  s->P_pad = ((s->P + TC_PAIR_TILE - 1) / TC_PAIR_TILE) * TC_PAIR_TILE;
  JP_CUDA(jp_dmalloc(post->ctx, &s->d_ds, (size_t)s->P_pad * kp * sizeof(float)));
  JP_CUDA(cudaMemsetAsync(s->d_ds, 0, (size_t)s->P_pad * kp * sizeof(float), post->ctx->stream));
  JP_CUDA(jp_dmalloc(post->ctx, &s->d_quad, (size_t)post->M * 8));
  JP_TRY(make_tensor_map(&s->tmB, s->d_ds, s->P_pad, kp, TC_PAIR_TILE));
  return JP_OK;
}

// Order / eligibility decision from the reduced bounds (see tc_obs_prep_kernel).  Returns NC in {4, 6, .. 12}
// or 0 when the series is not trustworthy for this (data, U, grid) and the FP64 kernel must be used; *fold says
// whether the economised coefficients (tc_fold_kernel) are to be used for that NC.
//   * series convergence (rigorous): max_i |Delta_i| <= t_max z_max must stay well inside the radius pi
//   * truncation (rigorous): error of the evaluated polynomial summed over ALL observations <= 2.5e-8 at
//     |z| <= z_ref = min(z_max, 6) (40 x below the 1e-6 tolerance of this path) and <= 1e-4 up to z_max (nodes
//     beyond |z| = 6 carry < e^-9 of the peak density, so 1e-4 relative on them is < 1e-8 of the largest weight).
//     The economised series of NC coefficients is tried before the plain Taylor truncation of the same length.
//   * rounding (statistical): FP32 Horner and the 3xTF32 contraction perturb each R_i by ~2e-6 |R_i| with
//     pseudo-random sign, so the log-density error is ~2e-6 sqrt(sum_i R_i^2); a factor 8 of margin is
//     required below 2e-7.  (The worst case with every error aligned, 2e-6 sum_i |R_i|, is reported too.)
#define TC_TRUNC_REF 2.5e-8
#define TC_TRUNC_MAX 1e-4
static int jp_tc_choose_order(const double* b, double* err_trunc, double* err_round, int* fold) {
  const double tdelta = b[0];
  *fold = 0;
  if (!(tdelta <= 2.0)) return 0;
  *err_round = 2e-6 * std::sqrt(b[13]);
  if (!(8.0 * *err_round <= 2e-7)) return 0;
  static const bool no_fold = getenv("JP_TC_NO_FOLD") != nullptr;   // A/B aid: plain Taylor truncation only
  for (int j = 0; j < TC_NORD; ++j) {
    // 1 + 1e-5: t_i is stored rounded up to FP32 for the fold, the interval grows by <= 1.2e-7 relative
    if (!no_fold && j < TC_NFOLD && b[14 + j] * (1 + 1e-5) <= TC_TRUNC_REF && b[14 + TC_NFOLD + j] * (1 + 1e-5) <= TC_TRUNC_MAX) {
      *err_trunc = b[14 + j] * (1 + 1e-5);
      *fold = 1;
      return 2 * j + 4;
    }
    if (b[2 + j] <= TC_TRUNC_REF && b[7 + j] <= TC_TRUNC_MAX) {
      *err_trunc = b[2 + j];
      return 2 * j + 4;
    }
  }
  return 0;
}

template <int NC, int MODE>
static int launch_tc_mode(jp_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcKernelParams& kp, size_t smem) {
  JP_CUDA(cudaFuncSetAttribute(jp_glm_tc_kernel<NC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min(ctx->sm_count, kp.n_pair_tiles * kp.chunks);
  if (!kp.nc_sel) cudaEventRecord(ctx->ev_k0, ctx->stream);      // guarded launches are bracketed as a group by the caller
  jp_glm_tc_kernel<NC, MODE><<<grid, TC_THREADS, smem, ctx->stream>>>(tmA, tmB, kp);
  if (!kp.nc_sel) {
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    ctx->ev_valid = true;
  }
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}

template <int NC>
static int launch_tc(jp_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcKernelParams& kp, size_t smem) {
#ifdef JP_TC_PROFILING_MODES
  static const int mode = getenv("JP_TC_DEBUG_MODE") ? atoi(getenv("JP_TC_DEBUG_MODE")) : 0;
  if (mode == 1) return launch_tc_mode<NC, 1>(ctx, tmA, tmB, kp, smem);
  if (mode == 2) return launch_tc_mode<NC, 2>(ctx, tmA, tmB, kp, smem);
  if (mode == 3) return launch_tc_mode<NC, 3>(ctx, tmA, tmB, kp, smem);
  if (mode == 4) return launch_tc_mode<NC, 4>(ctx, tmA, tmB, kp, smem);
  if (mode == 5) {   // per-tile timeline of CTA 0, dumped as 6 x TC_DBG_TILES int64 to $JP_TC_TIMELINE after every launch
    static long long* d_dbg = nullptr;
    if (!d_dbg) JP_CUDA(cudaMalloc(&d_dbg, sizeof(long long) * 6 * TC_DBG_TILES));
    JP_CUDA(cudaMemsetAsync(d_dbg, 0, sizeof(long long) * 6 * TC_DBG_TILES, ctx->stream));
    TcKernelParams k2 = kp;
    k2.dbg = d_dbg;
    JP_TRY((launch_tc_mode<NC, 5>)(ctx, tmA, tmB, k2, smem));
    std::vector<long long> h(6 * TC_DBG_TILES);
    JP_CUDA(cudaMemcpyAsync(h.data(), d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    JP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (const char* path = getenv("JP_TC_TIMELINE")) {
      if (FILE* f = fopen(path, "wb")) {
        fwrite(h.data(), sizeof(long long), h.size(), f);
        fclose(f);
      }
    }
    return JP_OK;
  }
#endif
  return launch_tc_mode<NC, 0>(ctx, tmA, tmB, kp, smem);
}

// ---- the fit in pieces.  One GPU: setup, prep of all observations, decide, fold, run.  Several ranks (node-sharded
// posterior, observations replicated): the O(N) FP64 prep is SHARDED BY OBSERVATION -- rank r prepares slice r, the
// ranks exchange (sums, bounds) and then the coefficient rows -- so that the replicated work per rank stays O(N / world).
static int tc_setup(jp_posterior* post, const jp_fit_args* args, int world, int rank) {
  jp_ctx* ctx = post->ctx;
  jp_data* data = const_cast<jp_data*>(post->data);
  if (!tc_static_ok(post, args)) return JP_ERR_UNSUPPORTED;
  JP_REQUIRE(post->grid->M % 2 == 1, "tensor-core path: the grid is not in mirror order (even node count %lld)", post->grid->M);
  JP_TRY(upload_tables(ctx));
  JP_TRY(ensure_data_state(ctx, data, args->d, world, rank));
  TcDataState* ds = static_cast<TcDataState*>(data->tc_state);
  JP_REQUIRE(ds->d == args->d, "tensor-core path: data was prepared for d=%d", ds->d);
  JP_TRY(ensure_post_state(post, ds->kp_b));
  JP_TRY(jp_upload_fit_consts(post, args));
  return JP_OK;
}

// FP64 sums (g, H, L_hat) over slice `rank` into d_sums, per-observation coefficients of the slice and the block
// partials of the bounds into ds->d_bounds
// what: 1 = fork + coefficient pass (`side`), 2 = sums (`side2`; after a call with 1), 3 = both
static int tc_prep_slice(jp_posterior* post, const jp_fit_args* args, int rank, double* d_sums, int what = 3) {
  jp_ctx* ctx = post->ctx;
  const jp_data* data = post->data;
  TcDataState* ds = static_cast<TcDataState*>(data->tc_state);
  const int d = args->d, p = args->p;
  const long long o0 = std::min(data->N, (long long)rank * ds->n_loc), o1 = std::min(data->N, o0 + ds->n_loc);
  // The two O(N) passes over the slice are independent of each other (sums g, H, L_hat | per-observation coefficients and
  // bounds) and of the node operand: the coefficient pass runs on the context's stream `side`, the sums on `side2`, both
  // forked here; the main stream stays free for the node operand.  The caller joins (tc_join_side) before anything reads
  // their results on the main stream.
  if (what & 1) {
    JP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    JP_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    JP_CUDA(cudaStreamWaitEvent(ctx->side2, ctx->ev_fork, 0));
    const double z_max = std::sqrt(post->grid->zmax2), z_ref = std::min(z_max, 6.0);
    const size_t sm_obs = (size_t)(((d + 1) & ~1) + d * ((p + 3) & ~3) + 2 * TC_PREP_THREADS * (data->ncols | 1)) * 8;
    if (sm_obs > 48 * 1024)
      JP_CUDA(cudaFuncSetAttribute(tc_obs_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_obs));
    tc_obs_prep_kernel<<<ds->prep_blocks, TC_PREP_THREADS, sm_obs, ctx->side>>>(
        data->family, d, p, data->ncols, o1 - o0, ds->n_loc, data->d_obs + (size_t)o0 * data->ncols, post->d_mu, post->d_U, z_ref,
        z_max, ds->d_coef, ds->n_loc, ds->d_bounds);
    JP_CHECK_LAUNCH(ctx);
  }
  if (what & 2) JP_TRY(jp_glm_sums_device_range_on(ctx, ctx->side2, data, d, post->d_mu, d_sums, ds->d_work, ds->glm_blocks, o0, o1));
  return JP_OK;
}

// the main stream waits for everything queued on the two side streams so far
static int tc_join_side(jp_ctx* ctx) {
  JP_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
  JP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  JP_CUDA(cudaEventRecord(ctx->ev_join2, ctx->side2));
  JP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));
  return JP_OK;
}
// The contraction kernel needs the coefficient pass (`side`) only; the sums and the FP64 quadratic part of every node (`side2`)
// are read by the finish / stage 4 BEHIND the kernel, so the main stream picks them up there (tc_join_quad) and the kernel's
// start no longer waits for the sums -> quadratic-part chain.
static int tc_join_coef(jp_ctx* ctx) {
  JP_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
  JP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  JP_CUDA(cudaEventRecord(ctx->ev_join2, ctx->side2));       // waited for behind the kernel
  return JP_OK;
}
static int tc_join_quad(jp_ctx* ctx) {
  JP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));
  return JP_OK;
}

// series length from the reduced bounds b[TC_NBOUND] (b[0] still without the factor z_max); records the diagnostics
static int tc_decide(jp_posterior* post, double* b, int* NC_out, int* fold_out) {
  const double z_max = std::sqrt(post->grid->zmax2);
  b[0] *= z_max;
  double err_trunc = 0, err_round = 0;
  int fold = 0;
  const int NC = jp_tc_choose_order(b, &err_trunc, &err_round, &fold);
  post->tc_bounds[0] = b[0]; post->tc_bounds[1] = err_trunc; post->tc_bounds[2] = err_round; post->tc_bounds[3] = NC;
  post->tc_bounds[4] = 2e-6 * b[12]; post->tc_bounds[5] = fold;
  if (NC == 0) {
    jp_set_error("tensor-core path: series bounds not met (max |Delta| %.3g, rounding estimate %.3g, truncation bounds %.3g/%.3g "
                 "at order 14); use the FP64 path", b[0], 2e-6 * std::sqrt(b[13]), b[6], b[11]);
    return JP_ERR_UNSUPPORTED;
  }
  *NC_out = NC;
  *fold_out = fold;
  return JP_OK;
}

static int tc_fold_slice(jp_posterior* post, int NC) {
  jp_ctx* ctx = post->ctx;
  TcDataState* ds = static_cast<TcDataState*>(post->data->tc_state);
  const double z_max = std::sqrt(post->grid->zmax2), z_ref = std::min(z_max, 6.0);
  tc_fold_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(NC, ds->n_loc, ds->n_loc, z_ref, ds->d_coef);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}

// node operand, theta, FP64 quadratic part: independent of the series length, so the single-GPU fit queues it BEFORE it
// waits for the bounds (the device works on it while the host decides)
template <int PART>
static int tc_node_prep_part(jp_posterior* post, const jp_fit_args* args, cudaStream_t st) {
  jp_ctx* ctx = post->ctx;
  const jp_data* data = post->data;
  TcDataState* ds = static_cast<TcDataState*>(data->tc_state);
  TcPostState* ps = static_cast<TcPostState*>(post->tc_state);
  const int d = args->d, p = args->p;
  size_t sm_node = (size_t)(d + d * p + d + d * d + JP_RULE_NMAX + 128 * d) * 8;
  if (sm_node > 48 * 1024)
    JP_CUDA(cudaFuncSetAttribute(tc_node_prep_kernel<PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_node));
  tc_node_prep_kernel<PART><<<(unsigned)((post->M + 127) / 128), 128, sm_node, st>>>(
      d, p, ds->kp_b, ds->split ? TC_KATOM : d, post->grid->rule, post->M, post->m0, post->grid->M, ps->j_lo, post->grid->d_idx,
      jp_rule_nodes_dev(ctx, post->grid->rule), post->d_mu, post->d_U, ds->d_sums, data->hyper[0], post->d_theta, ps->d_quad,
      ps->d_ds);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}
// theta + pair operand on the main stream (needs mu, U only)
static int tc_node_operand(jp_posterior* post, const jp_fit_args* args) { return tc_node_prep_part<0>(post, args, post->ctx->stream); }
// the quadratic part on `st`, behind the sums it needs
static int tc_node_quad(jp_posterior* post, const jp_fit_args* args, cudaStream_t st) { return tc_node_prep_part<1>(post, args, st); }

// the tensor-core kernel and the per-node finish (after tc_node_prep)
static int tc_run_kernel(jp_posterior* post, const jp_fit_args* args, int NC, bool finish, const int* nc_sel = nullptr,
                         const float* coef_gathered = nullptr, bool rest = false, bool join_quad = false) {
  // rest: the single launch that serves NC = 6 .. 12 under a device-side decision (geometry of NC = TC_NCMAX; the
  // coefficient slice stride of the gathered rows is NC-dependent and read from the decision by the kernel)
  jp_ctx* ctx = post->ctx;
  const jp_data* data = post->data;
  TcDataState* ds = static_cast<TcDataState*>(data->tc_state);
  TcPostState* ps = static_cast<TcPostState*>(post->tc_state);
  cudaStream_t st = ctx->stream;
  // work decomposition: node tiles x observation chunks on a persistent grid
  TcKernelParams kp;
  kp.ka = ds->ka;
  kp.kb = ds->kp_b / TC_KATOM;
  kp.n_pair_tiles = (int)(ps->P_pad / TC_PAIR_TILE);
  kp.n_obs_tiles = (int)((ds->N + TC_OBS_TILE - 1) / TC_OBS_TILE);   // tiles past the data (slice padding) hold zeros: skipped
  const size_t b_bytes = (size_t)kp.kb * TC_PAIR_TILE * 128, a_bytes = (size_t)kp.ka * TC_OBS_TILE * 128;
  const size_t c_bytes = (size_t)NC * TC_OBS_TILE * 4;      // coefficient slot of one tile
  const size_t budget = 220 * 1024, misc = 1024 + 4 * TC_PAIR_TILE * 2 * 8 + 256 + (TC_NBUF + 3) * c_bytes;   // (+3: slot count rounded up to a multiple of four)
  // two pair-operand buffers unless that would push the observation ring below three stages
  kp.nbbuf = ((budget - misc - 2 * b_bytes) / (a_bytes + c_bytes) >= 3) ? 2 : 1;
  const size_t fixed = misc + (size_t)kp.nbbuf * b_bytes;
  kp.stages = (int)std::max<size_t>(2, std::min<size_t>(TC_MAX_STAGES, (budget - fixed) / (a_bytes + c_bytes)));
  const size_t smem = fixed + (size_t)kp.stages * (a_bytes + c_bytes);
  // Work items are (observation chunk, node tile) pairs in CHUNK-MAJOR order on the persistent grid: all CTAs sweep
  // the same chunk of observation tiles at about the same time, so a chunk is read from HBM once and then served
  // from the 126 MB L2 to every node tile.  A chunk is therefore sized to ~24 MB of operand rows; beyond that the
  // chunk count is nudged until the last round of the static schedule is >= 97 % full.
  const long long l2_tiles = std::max<long long>(8, (24ll << 20) / (long long)a_bytes);
  const int c_min = (int)((kp.n_obs_tiles + l2_tiles - 1) / l2_tiles);
  int best_c = c_min;
  double best_eff = 0;
  for (int c = c_min; c <= c_min + 48; ++c) {
    if (c > c_min && kp.n_obs_tiles / c < 8) break;
    long long items = (long long)kp.n_pair_tiles * c;
    long long rounds = (items + ctx->sm_count - 1) / ctx->sm_count;
    double eff = (double)items / (double)(rounds * ctx->sm_count);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_c = c; }
    if (eff >= 0.97) break;
  }
  kp.chunks = best_c;
  kp.tiles_per_chunk = (kp.n_obs_tiles + kp.chunks - 1) / kp.chunks;
  kp.chunks = (kp.n_obs_tiles + kp.tiles_per_chunk - 1) / kp.tiles_per_chunk;   // no empty chunk
  if (kp.chunks > ps->part_chunks) {
    jp_dfree(ctx, ps->d_part);
    ps->d_part = nullptr;
    ps->part_chunks = 0;
    JP_CUDA(jp_dmalloc(ctx, &ps->d_part, (size_t)kp.chunks * ps->P * 2 * 8));
    ps->part_chunks = kp.chunks;
  }
  kp.P = ps->P;
  if (ds->world > 1 && coef_gathered) {      // [world][NC][n_loc] in the mailbox's bulk region (jp_fit_p2p)
    kp.coef = coef_gathered;
    kp.coef_slice_stride = (long long)NC * ds->n_loc;
  } else if (ds->world > 1) {
    JP_REQUIRE(ds->d_coef_all && ds->nc_all == NC, "tensor-core path: the coefficient rows of the other ranks have not been gathered "
               "(jp_fit_coef_slab + one all_gather, then jp_fit_local_stats_prepared)");
    kp.coef = ds->d_coef_all;
    kp.coef_slice_stride = (long long)NC * ds->n_loc;
  } else {
    kp.coef = ds->d_coef;
    kp.coef_slice_stride = 0;
  }
  kp.coef_n_loc = ds->n_loc;
  kp.tiles_per_slice = (int)(ds->n_loc / TC_OBS_TILE);
  kp.part = ps->d_part;
  kp.nc_sel = nc_sel;
  kp.dbg = nullptr;
  int stc;
  if (rest) {
    JP_CUDA(cudaFuncSetAttribute(jp_glm_tc_rest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jp_glm_tc_rest_kernel<<<std::min(ctx->sm_count, kp.n_pair_tiles * kp.chunks), TC_THREADS, smem, ctx->stream>>>(ds->tmA, ps->tmB, kp);
    JP_CHECK_LAUNCH(ctx);
    stc = JP_OK;
  } else if (NC == 4) stc = launch_tc<4>(ctx, ds->tmA, ps->tmB, kp, smem);
  else if (NC == 6) stc = launch_tc<6>(ctx, ds->tmA, ps->tmB, kp, smem);
  else if (NC == 8) stc = launch_tc<8>(ctx, ds->tmA, ps->tmB, kp, smem);
  else if (NC == 10) stc = launch_tc<10>(ctx, ds->tmA, ps->tmB, kp, smem);
  else stc = launch_tc<12>(ctx, ds->tmA, ps->tmB, kp, smem);
  JP_TRY(stc);
  JP_MARK(ctx, "fit:tc_kernel");
  if (join_quad) JP_TRY(tc_join_quad(ctx));      // the finish / stage 4 read the quadratic part
  post->path_used = JP_PATH_TC;
  post->fin = JpFinish();
  post->fin.path = JP_PATH_TC; post->fin.chunks = kp.chunks; post->fin.P = ps->P; post->fin.j_lo = ps->j_lo;
  post->fin.tc_part = ps->d_part; post->fin.quad = ps->d_quad; post->fin.neg_min = args->neg_min;
  if (finish) {
    tc_finish_kernel<<<(unsigned)((post->M + 255) / 256), 256, 0, st>>>(post->M, post->m0, ps->j_lo, ps->P, kp.chunks, ps->d_part,
                                                                         ps->d_quad, post->grid->d_hzz, args->neg_min, post->d_logdens,
                                                                         post->d_a);
    JP_CHECK_LAUNCH(ctx);
    post->fin.path = 0;
  }
  return JP_OK;
}

static int tc_run(jp_posterior* post, const jp_fit_args* args, int NC, bool finish) {
  TcPostState* ps = static_cast<TcPostState*>(post->tc_state);
  if (!ps->node_prep_queued) {
    JP_TRY(tc_node_operand(post, args));
    JP_TRY(tc_node_quad(post, args, post->ctx->stream));
  }
  ps->node_prep_queued = false;
  return tc_run_kernel(post, args, NC, finish);
}

// reduce the block partials of the bounds on the device: one warp per bound, lane l adds blocks l, l + 32, .. in ascending
// order, then the fixed shuffle tree (deterministic); [0] is a maximum.  Launch with TC_NBOUND warps.
__global__ void __launch_bounds__(32 * TC_NBOUND) tc_bounds_reduce_kernel(const double* __restrict__ blocks, int nblocks,
                                                                          double* __restrict__ out) {
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = 0;
  for (int blk = lane; blk < nblocks; blk += 32) {
    const double o = blocks[(size_t)blk * TC_NBOUND + j];
    v = (j == 0) ? fmax(v, o) : v + o;
  }
  v = (j == 0) ? jp_warp_max(v) : jp_warp_sum(v);
  if (lane == 0) out[j] = v;
}

int jp_fit_tc_launch(jp_posterior* post, const jp_fit_args* args, bool finish) {
  jp_ctx* ctx = post->ctx;
  JP_MARK(ctx, "fit:start");
  static const bool dev_decision = getenv("JP_TC_DEVICE_DECISION") != nullptr;   // A/B aid: no host round trip inside the fit
  if (dev_decision && !finish) return jp_fit_tc_launch_dev(post, args, nullptr, 0);
  JP_TRY(tc_setup(post, args, 1, 0));
  JP_MARK(ctx, "fit:setup+consts");
  TcDataState* ds = static_cast<TcDataState*>(post->data->tc_state);
  // diagnostic only (bench.py attribution of the host round trip): JP_TC_ASSUME=NC,fold skips the bounds read-back
  static const char* assume = getenv("JP_TC_ASSUME");
  if (assume) {
    int nc = 4, fd = 1;
    sscanf(assume, "%d,%d", &nc, &fd);
    post->tc_bounds[3] = nc; post->tc_bounds[5] = fd;
    JP_TRY(tc_prep_slice(post, args, 0, ds->d_sums));
    JP_TRY(tc_node_quad(post, args, ctx->side2));
    JP_TRY(tc_node_operand(post, args));
    JP_TRY(tc_join_side(ctx));
    if (fd) JP_TRY(tc_fold_slice(post, nc));
    return tc_run_kernel(post, args, nc, finish);
  }
  // `side`: block bounds -> 22 numbers -> pinned host memory; `side2`: sums, then the quadratic part of every node; main
  // stream: theta and the pair operand.  The host's wait for the bounds is hidden under the three (measured: skipping it
  // with JP_TC_ASSUME changes the fit time by < 5 us).
  // Enqueue order = urgency: the chain the kernel's start hangs on first (coefficient pass -> bounds -> host), then the node
  // operand (needed by the kernel), last the sums and the quadratic part (needed behind the kernel only).
  double* hb = ctx->h_pinned + JP_PINNED_BOUNDS_OFF;   // away from the constants staged by jp_upload_fit_consts
  JP_TRY(tc_prep_slice(post, args, 0, ds->d_sums, 1));
  JP_MARK_SIDE(ctx, "fit:obs_prep(side)");
  tc_bounds_reduce_kernel<<<1, 32 * TC_NBOUND, 0, ctx->side>>>(ds->d_bounds, ds->prep_blocks, ds->d_comb);
  JP_CHECK_LAUNCH(ctx);
  JP_CUDA(cudaMemcpyAsync(hb, ds->d_comb, (size_t)TC_NBOUND * 8, cudaMemcpyDeviceToHost, ctx->side));
  JP_CUDA(cudaEventRecord(ds->ev_bounds, ctx->side));
  JP_MARK_SIDE(ctx, "fit:bounds_d2h(side)");
  JP_TRY(tc_node_operand(post, args));
  JP_MARK(ctx, "fit:node_operand");
  JP_TRY(tc_prep_slice(post, args, 0, ds->d_sums, 2));
  jp_trace_mark(ctx, "fit:glm_sums(side2)", ctx->side2);
  JP_TRY(tc_node_quad(post, args, ctx->side2));       // behind the sums on their stream
  jp_trace_mark(ctx, "fit:node_quad(side2)", ctx->side2);
  JP_CUDA(cudaEventSynchronize(ds->ev_bounds));
  JP_TRY(tc_join_coef(ctx));
  double b[TC_NBOUND];
  for (int j = 0; j < TC_NBOUND; ++j) b[j] = hb[j];
  int NC = 0, fold = 0;
  const int st_dec = tc_decide(post, b, &NC, &fold);
  if (st_dec != JP_OK) {
    tc_join_quad(ctx);      // leave no work of this fit un-joined behind an error return
    return st_dec;
  }
  if (fold) JP_TRY(tc_fold_slice(post, NC));
  JP_MARK(ctx, "fit:host_decision+fold");
  return tc_run_kernel(post, args, NC, finish, nullptr, nullptr, false, true);
}

// ---- sharded prep, phase by phase (extern "C" wrappers in jp_fit.cu).  L = nE + 1 + TC_NBOUND doubles per rank.
int jp_fit_tc_prep_len(int d) { return d + d * (d + 1) / 2 + 1 + TC_NBOUND; }

// phase 1: this rank's observation slice -> d_out[L] = (local g, H, L_hat | local bounds); asynchronous
int jp_fit_tc_prep_local(jp_posterior* post, const jp_fit_args* args, int rank, int world, double* d_out) {
  JP_REQUIRE(world >= 1 && rank >= 0 && rank < world && d_out, "jp_fit_prep_local: bad rank / world / output");
  JP_TRY(tc_setup(post, args, world, rank));
  TcDataState* ds = static_cast<TcDataState*>(post->data->tc_state);
  const int nE1 = args->d + args->d * (args->d + 1) / 2 + 1;
  JP_TRY(tc_prep_slice(post, args, rank, d_out));
  tc_bounds_reduce_kernel<<<1, 32 * TC_NBOUND, 0, post->ctx->side>>>(ds->d_bounds, ds->prep_blocks, d_out + nE1);
  JP_CHECK_LAUNCH(post->ctx);
  JP_TRY(tc_node_operand(post, args));  // theta and the pair operand of this rank's node block, under the O(N) passes
  JP_TRY(tc_join_side(post->ctx));      // the caller's collective reads d_out on the main stream
  return JP_OK;
}

// gathered [world][L] -> global sums (g, H, L_hat) and bounds, combined in RANK ORDER on the device: every rank runs the same
// kernel on the same gathered bits, so all ranks hold identical sums -- and no host arithmetic sits between the collectives
__global__ void tc_combine_gathered_kernel(const double* __restrict__ g, int world, int L, int nE1, double* __restrict__ sums,
                                           double* __restrict__ bounds) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < L; e += gridDim.x * blockDim.x) {
    double v = 0;
    if (e == nE1) {
      for (int r = 0; r < world; ++r) v = fmax(v, g[(size_t)r * L + e]);     // [0] of the bounds is a maximum
    } else {
      for (int r = 0; r < world; ++r) v += g[(size_t)r * L + e];
    }
    if (e < nE1) sums[e] = v; else bounds[e - nE1] = v;
  }
}

// phase 2: the gathered [world][L] buffer -> global sums on the device, the series length, the fold of this rank's slice.
// The host needs the 22 combined bounds to pick the series length (it selects the kernel instantiation and the size of the
// one collective that follows); the node operand / theta / quadratic part of this rank's node block (tc_node_prep, ~25 us)
// is queued BEFORE the host waits, so the device works through the round trip.  *n_rows = coefficient rows to exchange.
int jp_fit_tc_prep_gathered(jp_posterior* post, const jp_fit_args* args, const double* d_gathered, int world, int rank,
                            int* n_rows) {
  jp_ctx* ctx = post->ctx;
  TcDataState* ds = static_cast<TcDataState*>(post->data->tc_state);
  TcPostState* ps = static_cast<TcPostState*>(post->tc_state);
  JP_REQUIRE(ds && ps && ds->world == world && ds->rank == rank && d_gathered && n_rows, "jp_fit_prep_gathered: call jp_fit_prep_local first");
  const int nE1 = args->d + args->d * (args->d + 1) / 2 + 1, L = nE1 + TC_NBOUND;
  tc_combine_gathered_kernel<<<(L + 127) / 128, 128, 0, ctx->stream>>>(d_gathered, world, L, nE1, ds->d_sums, ds->d_comb);
  JP_CHECK_LAUNCH(ctx);
  double* hb = ctx->h_pinned + JP_PINNED_BOUNDS_OFF;
  JP_CUDA(cudaMemcpyAsync(hb, ds->d_comb, (size_t)TC_NBOUND * 8, cudaMemcpyDeviceToHost, ctx->stream));
  JP_CUDA(cudaEventRecord(ds->ev_bounds, ctx->stream));
  JP_TRY(tc_node_quad(post, args, ctx->stream));      // needs the combined sums; the operand was queued by jp_fit_prep_local
  ps->node_prep_queued = true;
  JP_CUDA(cudaEventSynchronize(ds->ev_bounds));
  double b[TC_NBOUND];
  for (int j = 0; j < TC_NBOUND; ++j) b[j] = hb[j];
  int NC = 0, fold = 0;
  JP_TRY(tc_decide(post, b, &NC, &fold));
  if (fold) JP_TRY(tc_fold_slice(post, NC));
  *n_rows = NC;
  return JP_OK;
}

// The exchange of the coefficient rows is ONE all_gather: this rank contributes the first n_rows rows of its slab (contiguous,
// n_rows x n_loc floats at *d_local) and receives everybody's into *d_all = [world][n_rows][n_loc].
int jp_fit_tc_coef_slab(jp_posterior* post, int n_rows, float** d_local, float** d_all, long long* count) {
  TcDataState* ds = post && post->data ? static_cast<TcDataState*>(post->data->tc_state) : nullptr;
  JP_REQUIRE(ds && d_local && d_all && count, "jp_fit_coef_slab: call jp_fit_prep_local first");
  JP_REQUIRE(n_rows >= 1 && n_rows <= TC_NCMAX, "jp_fit_coef_slab: n_rows=%d out of range", n_rows);
  if (ds->nc_all != n_rows || !ds->d_coef_all) {
    jp_dfree(post->ctx, ds->d_coef_all);
    ds->d_coef_all = nullptr;
    ds->nc_all = 0;
    JP_CUDA(jp_dmalloc(post->ctx, &ds->d_coef_all, (size_t)ds->world * n_rows * ds->n_loc * sizeof(float)));
    ds->nc_all = n_rows;
  }
  *d_local = ds->d_coef;
  *d_all = ds->d_coef_all;
  *count = (long long)n_rows * ds->n_loc;
  return JP_OK;
}

// phase 3 (after the rows are exchanged): stages 2-3 of this rank's node block
int jp_fit_tc_run_prepared(jp_posterior* post, const jp_fit_args* args, bool finish) {
  JP_REQUIRE(post->tc_bounds[3] >= 4, "jp_fit_run_prepared: no series length has been decided (call the prep phases first)");
  return tc_run(post, args, (int)post->tc_bounds[3], finish);
}

// ---- the whole tensor-core fit as ONE asynchronous queue: series length decided on the device ------------------------------
// (jp_fit_p2p on every rank of a node-sharded fit; one GPU with JP_TC_DEVICE_DECISION=1 or under stream capture)
__global__ void tc_decide_kernel(const double* __restrict__ comb, double z_max, int no_fold, TcDecision* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // jp_tc_choose_order / tc_decide, statement by statement
  const double tdelta = comb[0] * z_max;
  const double err_round = 2e-6 * sqrt(comb[13]);
  double err_trunc = 0;
  int NC = 0, fold = 0;
  if (tdelta <= 2.0 && 8.0 * err_round <= 2e-7) {
    for (int j = 0; j < TC_NORD && NC == 0; ++j) {
      if (!no_fold && j < TC_NFOLD && comb[14 + j] * (1 + 1e-5) <= TC_TRUNC_REF && comb[14 + TC_NFOLD + j] * (1 + 1e-5) <= TC_TRUNC_MAX) {
        err_trunc = comb[14 + j] * (1 + 1e-5);
        fold = 1;
        NC = 2 * j + 4;
      } else if (comb[2 + j] <= TC_TRUNC_REF && comb[7 + j] <= TC_TRUNC_MAX) {
        err_trunc = comb[2 + j];
        NC = 2 * j + 4;
      }
    }
  }
  out->NC = NC; out->fold = fold; out->pad0 = out->pad1 = 0;
  out->diag[0] = tdelta; out->diag[1] = err_trunc; out->diag[2] = err_round; out->diag[3] = NC;
  out->diag[4] = 2e-6 * comb[12]; out->diag[5] = fold;
}

// This rank's first NC coefficient rows (contiguous in its slab) into slot `rank` of every peer's bulk region
// [world][NC][n_loc], 16 bytes per store over NVLink; the block that finishes last raises this rank's flag on every peer.
__global__ void __launch_bounds__(256)
tc_coef_push_kernel(const JpCommDev c, const TcDecision* __restrict__ dec, const float* __restrict__ coef, long long n_loc,
                    size_t bulk_off, unsigned int* __restrict__ counter, unsigned long long seq) {
  const int NC = dec->NC;
  const long long count4 = (long long)NC * n_loc / 4;      // n_loc is a multiple of the 128-observation tile
  const float4* src = reinterpret_cast<const float4*>(coef);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    for (int p = 0; p < c.world; ++p) reinterpret_cast<float4*>(c.peer[p] + bulk_off)[(long long)c.rank * count4 + i] = v;
  }
  __shared__ unsigned int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1u);
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < c.world) {
    __threadfence_system();
    jp_st_release_sys(jp_comm_flag(c.peer[threadIdx.x], JP_CH_BULK, 0, c.rank), seq);
  }
}

// Observation-sharded fit: every rank's partial (E, O) sums over ITS observations, summed over the chunks, into slot `rank`
// of every peer's bulk region [world][P][2]; the per-node finish then adds the slots in rank order (JpFinish with
// chunks = world), so every rank holds bit-identical log-densities of ALL nodes.
__global__ void __launch_bounds__(256)
tc_part_push_kernel(const JpCommDev c, const double* __restrict__ part, int chunks, long long P, size_t bulk_off,
                    unsigned int* __restrict__ counter, unsigned long long seq) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (long long)gridDim.x * blockDim.x) {
    double2 sum = make_double2(0.0, 0.0);
    for (int ch = 0; ch < chunks; ++ch) {
      const double2 eo = *reinterpret_cast<const double2*>(part + ((size_t)ch * P + j) * 2);
      sum.x += eo.x;
      sum.y += eo.y;
    }
    for (int p = 0; p < c.world; ++p) reinterpret_cast<double2*>(c.peer[p] + bulk_off)[(long long)c.rank * P + j] = sum;
  }
  __shared__ unsigned int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1u);
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < c.world) {
    __threadfence_system();
    jp_st_release_sys(jp_comm_flag(c.peer[threadIdx.x], JP_CH_BULK, 0, c.rank), seq);
  }
}

// obs_sharded = 0: NODE sharding -- the posterior is this rank's node block, the data handle holds ALL observations, the O(N)
//   prep is sharded by observation slice and the coefficient rows are exchanged.
// obs_sharded = 1: OBSERVATION sharding -- the data handle holds this rank's observations only, the posterior
//   covers ALL nodes; the ranks exchange the (sums, bounds) and the per-pair partial sums, nothing else.
int jp_fit_tc_launch_dev(jp_posterior* post, const jp_fit_args* args, jp_comm* comm, int obs_sharded) {
  jp_ctx* ctx = post->ctx;
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  const bool obs = obs_sharded && world > 1;
  JP_MARK(ctx, "fit:start");
  JP_TRY(tc_setup(post, args, obs ? 1 : world, obs ? 0 : rank));
  TcDataState* ds = static_cast<TcDataState*>(post->data->tc_state);
  TcPostState* ps = static_cast<TcPostState*>(post->tc_state);
  const int d = args->d, nE1 = d + d * (d + 1) / 2 + 1, L = nE1 + TC_NBOUND;
  if (!ps->d_dec) JP_CUDA(jp_dmalloc(ctx, &ps->d_dec, sizeof(TcDecision)));
  if (world > 1) {
    if (!ds->d_loc) JP_CUDA(jp_dmalloc(ctx, &ds->d_loc, (size_t)L * 8));
    const size_t need = obs ? (size_t)world * (size_t)ps->P * 16 : (size_t)world * TC_NCMAX * (size_t)ds->n_loc * sizeof(float);
    JP_REQUIRE(comm->bulk_bytes >= need, "jp_fit_p2p: the communicator's bulk region holds %zu bytes, this fit (%lld observations, %lld "
               "nodes, %d ranks) may need %zu (jp_comm_create)", comm->bulk_bytes, ds->N, post->M, world, need);
  }
  const double z_max = std::sqrt(post->grid->zmax2), z_ref = std::min(z_max, 6.0);
  // side: coefficients + bounds of this rank's observations; side2: their sums (one GPU: then the quadratic part); main: node operand
  JP_TRY(tc_prep_slice(post, args, obs ? 0 : rank, world > 1 ? ds->d_loc : ds->d_sums));
  tc_bounds_reduce_kernel<<<1, 32 * TC_NBOUND, 0, ctx->side>>>(ds->d_bounds, ds->prep_blocks, world > 1 ? ds->d_loc + nE1 : ds->d_comb);
  JP_CHECK_LAUNCH(ctx);
  if (world == 1) JP_TRY(tc_node_quad(post, args, ctx->side2));
  JP_TRY(tc_node_operand(post, args));
  if (world == 1) JP_TRY(tc_join_coef(ctx));      // the quadratic part is joined behind the kernel
  else JP_TRY(tc_join_side(ctx));
  JP_MARK(ctx, "fit:prep_joined");
  if (world > 1) {
    const double* g = nullptr;
    JP_TRY(jp_comm_exchange(comm, JP_CH_PREP, ds->d_loc, L, &g));
    tc_combine_gathered_kernel<<<(L + 127) / 128, 128, 0, ctx->stream>>>(g, world, L, nE1, ds->d_sums, ds->d_comb);
    JP_CHECK_LAUNCH(ctx);
    JP_MARK(ctx, "fit:prep_exchanged");
    // the quadratic part of every node needs the combined sums only: on side2, beside the decision / fold / coefficient push
    JP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    JP_CUDA(cudaStreamWaitEvent(ctx->side2, ctx->ev_fork, 0));
    JP_TRY(tc_node_quad(post, args, ctx->side2));
    JP_CUDA(cudaEventRecord(ctx->ev_join2, ctx->side2));
  }
  static const bool no_fold = getenv("JP_TC_NO_FOLD") != nullptr;
  tc_decide_kernel<<<1, 32, 0, ctx->stream>>>(ds->d_comb, z_max, no_fold ? 1 : 0, ps->d_dec);
  JP_CHECK_LAUNCH(ctx);
  tc_fold_dev_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ps->d_dec, ds->n_loc, ds->n_loc, z_ref, ds->d_coef);
  JP_CHECK_LAUNCH(ctx);
  const float* coef_gathered = nullptr;
  if (world > 1 && !obs) {
    unsigned long long seq = 0;
    JP_TRY(jp_comm_bulk_begin(comm, &seq));
    const int pb = (int)std::max<long long>(1, std::min<long long>(2LL * ctx->sm_count, ds->n_loc / 256));
    tc_coef_push_kernel<<<pb, 256, 0, ctx->stream>>>(jp_comm_dev(comm), ps->d_dec, ds->d_coef, ds->n_loc, comm->bulk_off,
                                                     comm->d_counter, seq);
    JP_CHECK_LAUNCH(ctx);
    JP_TRY(jp_comm_wait(comm, JP_CH_BULK, 0, seq));
    coef_gathered = reinterpret_cast<const float*>(comm->mailbox + comm->bulk_off);
    JP_MARK(ctx, "fit:coef_exchanged");
  }
  // two launches: the shortest series (the usual choice) and one kernel for all longer ones; whichever the device did not
  // choose returns before touching anything
  JP_CUDA(cudaEventRecord(ctx->ev_k0, ctx->stream));
  JP_TRY(tc_run_kernel(post, args, 4, false, &ps->d_dec->NC, coef_gathered));
  JP_TRY(tc_run_kernel(post, args, TC_NCMAX, false, &ps->d_dec->NC, coef_gathered, true));
  JP_CUDA(cudaEventRecord(ctx->ev_k1, ctx->stream));
  ctx->ev_valid = true;
  JP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));      // the quadratic part, computed beside the kernel (stage 4 reads it)
  if (obs) {
    unsigned long long seq = 0;
    JP_TRY(jp_comm_bulk_begin(comm, &seq));
    const int pb = (int)std::max<long long>(1, std::min<long long>(2LL * ctx->sm_count, (ps->P + 255) / 256));
    tc_part_push_kernel<<<pb, 256, 0, ctx->stream>>>(jp_comm_dev(comm), ps->d_part, post->fin.chunks, ps->P, comm->bulk_off,
                                                     comm->d_counter, seq);
    JP_CHECK_LAUNCH(ctx);
    JP_TRY(jp_comm_wait(comm, JP_CH_BULK, 0, seq));
    post->fin.chunks = world;       // the finish adds the ranks' slots in rank order
    post->fin.tc_part = reinterpret_cast<const double*>(comm->mailbox + comm->bulk_off);
    JP_MARK(ctx, "fit:partials_exchanged");
  }
  ps->dec_pending = true;
  ps->dec_prefetched = false;
  post->tc_bounds[3] = -1;       // not known to the host until jp_fit_tc_verify
  return JP_OK;
}

// The host reads the device's decision at its first blocking call after the fit (the result downloads synchronise anyway).
// JP_ERR_UNSUPPORTED: the bounds were not met and no instantiation ran -- the results of this fit are garbage.
// Queue the read-back behind whatever the caller is about to synchronise on (the decision then rides in the same wait as
// the results: no second synchronisation, no pageable copy).  The slot is the tail of the context's pinned buffer.
int jp_fit_tc_verify_prefetch(jp_posterior* post) {
  TcPostState* ps = post ? static_cast<TcPostState*>(post->tc_state) : nullptr;
  if (!ps || !ps->dec_pending || ps->dec_prefetched) return JP_OK;
  static_assert(sizeof(TcDecision) <= JP_PINNED_TAIL_DOUBLES * 8, "decision record does not fit the pinned tail");
  JP_CUDA(cudaMemcpyAsync(post->ctx->h_pinned + (JP_PINNED_DOUBLES - JP_PINNED_TAIL_DOUBLES), ps->d_dec, sizeof(TcDecision),
                          cudaMemcpyDeviceToHost, post->ctx->stream));
  ps->dec_prefetched = true;
  return JP_OK;
}

int jp_fit_tc_verify(jp_posterior* post) {
  TcPostState* ps = post ? static_cast<TcPostState*>(post->tc_state) : nullptr;
  if (!ps || !ps->dec_pending) return JP_OK;
  ps->dec_pending = false;
  TcDecision dec;
  if (ps->dec_prefetched) {      // the caller has synchronised the stream since jp_fit_tc_verify_prefetch
    ps->dec_prefetched = false;
    std::memcpy(&dec, post->ctx->h_pinned + (JP_PINNED_DOUBLES - JP_PINNED_TAIL_DOUBLES), sizeof dec);
  } else {
    JP_CUDA(cudaStreamSynchronize(post->ctx->stream));
    JP_CUDA(cudaMemcpy(&dec, ps->d_dec, sizeof dec, cudaMemcpyDeviceToHost));
  }
  for (int i = 0; i < 6; ++i) post->tc_bounds[i] = dec.diag[i];
  if (dec.NC == 0) {
    post->path_used = 0;
    jp_set_error("tensor-core path: series bounds not met (max |Delta| %.3g, rounding estimate %.3g); use the FP64 path",
                 dec.diag[0], dec.diag[2]);
    return JP_ERR_UNSUPPORTED;
  }
  return JP_OK;
}

