// jp_glm.cu -- GLM score / observed information on the GPU (FP64).
//
// Upstream of the five stages: this is what `mode` needs from the data (reference
// src/joint_posterior.jl:164-168: BFGS + ForwardDiff Hessian there; Newton/IRLS here because for a
// GLM the Hessian is the closed form X' W X).  The same sums (X'(y - mu_hat), X' W X at the mode) are
// what the tensor-core log-density path centres its expansion on (jp_glm_tc.cu).
//
// Layout: a block stages a tile of observation records in shared memory (coalesced loads), computes the
// per-record residual and weight once, then updates X' W X as a register-blocked rank update: a thread owns
// one BS x BS block of the upper block triangle and one slice of the tile's observations, so a record costs it
// 2 BS + 1 shared-memory doubles for BS^2 + BS FP64 operations (the FP64 pipe, not shared-memory bandwidth,
// bounds the loop).  Slices are summed in slice order at the end of the block, per-block partials in block
// order by a second kernel, so the result is deterministic.  At N = 1e7, d = 30 the kernel streams the 2.5 GB
// of records once; its floor is max(HBM time, N (d^2/2 + ~150) FP64 operations).
#include <algorithm>
#include <cstring>
#include <vector>
#include "jp_common.cuh"

#define JP_GLM_TILE 128    // observations per tile (upper bound: a tile holds slices x iterations <= 128)

// BS = 4: 256 threads; BS = 8: 128 threads (64 FP64 accumulators = 128 registers per thread); two blocks per SM
// either way, so that one block's record loads overlap the other's arithmetic
template <int BS, int JP_GLM_THREADS>
__global__ void __launch_bounds__(JP_GLM_THREADS, 2)
jp_glm_partials_kernel(int family, int d, long long N, const double* __restrict__ obs, const double* __restrict__ beta,
                       int nE, double* __restrict__ part /* [gridDim.x][nE + 1] */) {
  extern __shared__ __align__(16) double sh[];
  const int ncols = d + 1;
  const int nb = (d + BS - 1) / BS, npairs = nb * (nb + 1) / 2;
  const int rs = nb * BS + 2;                      // row stride: = 2 (mod 4) doubles -> conflict-free 16-byte loads
  const int S = min(JP_GLM_THREADS / npairs, JP_GLM_TILE), iters = JP_GLM_TILE / S, T = S * iters;
  // two record tiles: the asynchronous copies (cp.async, 8 bytes each, no register staging) of tile i + 1 are in flight
  // while tile i is consumed -- with a load -> store loop the kernel was bound by the latency of ~1 KB in flight per warp
  double* tile0 = sh;                              // 2 x JP_GLM_TILE x rs (columns >= d stay zero)
  double* ybuf0 = tile0 + 2 * JP_GLM_TILE * rs;    // 2 x JP_GLM_TILE: the responses of a tile
  double* res = ybuf0 + 2 * JP_GLM_TILE;           // JP_GLM_TILE
  double* wgt = res + JP_GLM_TILE;                 // JP_GLM_TILE
  double* s_beta = wgt + JP_GLM_TILE;              // d
  __shared__ double red[33];
  for (int k = threadIdx.x; k < d; k += JP_GLM_THREADS) s_beta[k] = beta[k];
  for (int k = threadIdx.x; k < 2 * JP_GLM_TILE * rs; k += JP_GLM_THREADS) tile0[k] = 0.0;
  // this thread's block (br <= bc) of the upper block triangle and its observation slice
  const int pair = threadIdx.x / S, slice = threadIdx.x - pair * S;
  const bool active = pair < npairs;
  int br = 0, bc = 0;
  if (active) {
    int t = pair;
    while (t >= bc + 1) { t -= bc + 1; ++bc; }     // packed by block column, like the entries themselves
    br = t;
  }
  const bool diag = active && br == bc;
  double acc[BS][BS], gacc[BS];
#pragma unroll
  for (int i = 0; i < BS; ++i) {
    gacc[i] = 0.0;
#pragma unroll
    for (int j = 0; j < BS; ++j) acc[i][j] = 0.0;
  }
  double ll = 0.0;
  // flat coalesced copy of a tile's records; (row, column) advance by (threads / ncols, threads % ncols) per step
  auto issue_tile = [&](long long base, int b) {
    if (base < N) {
      const int cnt = (int)min((long long)T, N - base);
      const double* src = obs + (size_t)base * ncols;
      double* tl = tile0 + (size_t)b * JP_GLM_TILE * rs;
      double* yb = ybuf0 + b * JP_GLM_TILE;
      int row = threadIdx.x / ncols, col = threadIdx.x - row * ncols;
      const int drow = JP_GLM_THREADS / ncols, dcol = JP_GLM_THREADS - drow * ncols;
      for (int e = threadIdx.x; e < cnt * ncols; e += JP_GLM_THREADS) {
        jp_cp_async8((col < d) ? (tl + row * rs + col) : (yb + row), src + e);
        row += drow;
        col += dcol;
        if (col >= ncols) { col -= ncols; ++row; }
      }
    }
    jp_cp_async_commit();
  };
  __syncthreads();       // the zero fill precedes the first asynchronous copies
  const long long step = (long long)gridDim.x * T;
  issue_tile((long long)blockIdx.x * T, 0);
  int cur = 0;
  for (long long base = (long long)blockIdx.x * T; base < N; base += step, cur ^= 1) {
    const int cnt = (int)min((long long)T, N - base);
    jp_cp_async_wait_all();
    __syncthreads();     // tile `cur` has landed for every thread; the previous tile (in the other buffer) is consumed
    issue_tile(base + step, cur ^ 1);
    const double* tile = tile0 + (size_t)cur * JP_GLM_TILE * rs;
    const double* yb = ybuf0 + cur * JP_GLM_TILE;
    for (int n = threadIdx.x; n < T; n += JP_GLM_THREADS) {
      double rv = 0.0, wv = 0.0;
      if (n < cnt) {
        const double* r = tile + n * rs;
        const double y = yb[n];
        double eta = 0;
        for (int k = 0; k < d; ++k) eta += r[k] * s_beta[k];
        double mu;
        if (family == JP_FAM_LOGISTIC) {
          mu = 1.0 / (1.0 + exp(-eta));
          wv = mu * (1.0 - mu);
          ll += y * eta - (fmax(eta, 0.0) + log1p(exp(-fabs(eta))));
        } else {
          mu = exp(eta);
          wv = mu;
          ll += y * eta - mu;
        }
        rv = y - mu;
      }
      res[n] = rv;     // records past the end of the data contribute nothing
      wgt[n] = wv;
    }
    __syncthreads();
    if (active) {
      const double* xr = tile + br * BS;
      const double* xc = tile + bc * BS;
      for (int it = 0; it < iters; ++it) {
        const int n = slice + it * S;
        const double w = wgt[n];
        double a[BS], b[BS];
#pragma unroll
        for (int i = 0; i < BS; i += 2) {
          const double2 va = *reinterpret_cast<const double2*>(xr + n * rs + i);
          const double2 vb = *reinterpret_cast<const double2*>(xc + n * rs + i);
          a[i] = va.x; a[i + 1] = va.y;
          b[i] = vb.x; b[i + 1] = vb.y;
        }
        if (diag) {
          const double r = res[n];
#pragma unroll
          for (int i = 0; i < BS; ++i) gacc[i] = fma(r, a[i], gacc[i]);
        }
#pragma unroll
        for (int i = 0; i < BS; ++i) {
          const double wa = w * a[i];
#pragma unroll
          for (int j = 0; j < BS; ++j) acc[i][j] = fma(wa, b[j], acc[i][j]);
        }
      }
    }
  }
  // slices -> block partial: one block pair at a time through shared memory (the tile is free now)
  double* o = part + (size_t)blockIdx.x * (nE + 1);
  constexpr int NV = BS * BS + BS;
  jp_cp_async_wait_all();
  double* buf = tile0;                             // [S][NV]
  for (int pp = 0; pp < npairs; ++pp) {
    __syncthreads();
    if (active && pair == pp) {
      double* w = buf + slice * NV;
#pragma unroll
      for (int i = 0; i < BS; ++i) {
#pragma unroll
        for (int j = 0; j < BS; ++j) w[i * BS + j] = acc[i][j];
        w[BS * BS + i] = gacc[i];
      }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NV; v += JP_GLM_THREADS) {
      int pc = 0, t = pp;
      while (t >= pc + 1) { t -= pc + 1; ++pc; }
      const int pr = t;
      double sum = 0;
      for (int sl = 0; sl < S; ++sl) sum += buf[sl * NV + v];
      if (v < BS * BS) {
        const int r = pr * BS + v / BS, c = pc * BS + v % BS;
        if (r <= c && c < d) o[d + c * (c + 1) / 2 + r] = sum;     // packed upper triangle, column by column
      } else if (pr == pc) {
        const int r = pr * BS + (v - BS * BS);
        if (r < d) o[r] = sum;                                     // gradient entry
      }
    }
  }
  ll = jp_block_sum(ll, red);
  if (threadIdx.x == 0) o[nE] = ll;
}

// one warp per entry: lane l adds blocks l, l + 32, ... in ascending order, then the fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(256) jp_glm_combine_kernel(int nblocks, int nE1, const double* __restrict__ part,
                                                             double* __restrict__ out) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= nE1) return;
  double s = 0;
  for (int b = lane; b < nblocks; b += 32) s += part[(size_t)b * nE1 + e];
  s = jp_warp_sum(s);
  if (lane == 0) out[e] = s;
}

// device-side entry used by both the C ABI and the TC path: packed sums into d_out[nE + 1]
template <int BS, int JP_GLM_THREADS>
static int launch_partials(jp_ctx* ctx, cudaStream_t stream, const jp_data* data, int d, const double* d_beta, double* d_work,
                           int nblocks, int nE, long long o0, long long o1) {
  const int nb = (d + BS - 1) / BS, npairs = nb * (nb + 1) / 2, rs = nb * BS + 2;
  const int S = std::min(JP_GLM_THREADS / npairs, JP_GLM_TILE);
  // the slice buffer of the final reduction reuses the tiles
  size_t tile_doubles = std::max<size_t>((size_t)2 * JP_GLM_TILE * rs, (size_t)S * (BS * BS + BS));
  size_t smem = (tile_doubles + 4 * JP_GLM_TILE + d) * sizeof(double);
  if (smem > 48 * 1024)
    JP_CUDA(cudaFuncSetAttribute(jp_glm_partials_kernel<BS, JP_GLM_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jp_glm_partials_kernel<BS, JP_GLM_THREADS><<<nblocks, JP_GLM_THREADS, smem, stream>>>(
      data->family, d, o1 - o0, data->d_obs + (size_t)o0 * data->ncols, d_beta, nE, d_work);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}

// sums over the observation rows [o0, o1) (the whole data set: 0, N)
int jp_glm_sums_device_range_on(jp_ctx* ctx, cudaStream_t stream, const jp_data* data, int d, const double* d_beta, double* d_out,
                                double* d_work, int nblocks, long long o0, long long o1) {
  int nE = d + d * (d + 1) / 2;
  if (o1 <= o0) {      // an empty slice (more ranks than observation tiles) contributes zeros
    JP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (nE + 1), stream));
    return JP_OK;
  }
  // 8 x 8 register blocks unless the padding to a multiple of 8 would waste more than it saves
  const int st = (d > 16) ? (launch_partials<8, 128>)(ctx, stream, data, d, d_beta, d_work, nblocks, nE, o0, o1)
                          : (launch_partials<4, 256>)(ctx, stream, data, d, d_beta, d_work, nblocks, nE, o0, o1);
  JP_TRY(st);
  jp_glm_combine_kernel<<<(nE + 1 + 7) / 8, 256, 0, stream>>>(nblocks, nE + 1, d_work, d_out);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}
int jp_glm_sums_device_range(jp_ctx* ctx, const jp_data* data, int d, const double* d_beta, double* d_out, double* d_work,
                             int nblocks, long long o0, long long o1) {
  return jp_glm_sums_device_range_on(ctx, ctx->stream, data, d, d_beta, d_out, d_work, nblocks, o0, o1);
}

int jp_glm_sums_device(jp_ctx* ctx, const jp_data* data, int d, const double* d_beta, double* d_out, double* d_work,
                       int nblocks) {
  return jp_glm_sums_device_range(ctx, data, d, d_beta, d_out, d_work, nblocks, 0, data->N);
}

int jp_glm_num_blocks(const jp_ctx* ctx, long long N) {
  long long tiles = (N + JP_GLM_TILE - 1) / JP_GLM_TILE;
  return (int)std::max(1LL, std::min(tiles, (long long)ctx->sm_count * 2));
}

// sum over the ranks of a gathered [world][n] buffer, in rank order (the same bits on every rank)
__global__ void jp_sum_gathered_kernel(const double* __restrict__ g, int world, int n, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double v = 0;
  for (int r = 0; r < world; ++r) v += g[(size_t)r * n + e];
  out[e] = v;
}

// comm != null: `data` holds this rank's observations only; the sums are exchanged and added in rank order, so every rank
// gets the score / information of ALL observations (jp_mode_p2p, observation-sharded fits)
int jp_glm_grad_hess_comm(jp_ctx* ctx, const jp_data* data, jp_comm* comm, int d, const double* h_beta, double* h_g,
                          double* h_Hneg, double* h_logpost) {
  JP_REQUIRE(ctx && data && h_beta, "jp_glm_grad_hess: null argument");
  JP_REQUIRE(data->family == JP_FAM_LOGISTIC || data->family == JP_FAM_POISSON,
             "jp_glm_grad_hess: family %d is not a GLM", data->family);
  JP_REQUIRE(d >= 1 && d <= JP_MAX_D && data->ncols == d + 1, "jp_glm_grad_hess: d=%d does not match %d columns", d,
             data->ncols);
  JP_ENTER_CTX(ctx);
  int nE = d + d * (d + 1) / 2;
  int nb = jp_glm_num_blocks(ctx, data->N);
  // A Newton iteration of the mode finder is one such call: no allocation when the block partials fit the context's scratch,
  // beta staged through the pinned buffer, and -- single GPU -- the combined sums written by the last kernel straight into the
  // pinned buffer (zero-copy), so the call is copy + two launches + one synchronisation.
  double *d_beta = nullptr, *d_out = nullptr, *d_work = nullptr;
  const bool exchange = comm && comm->world > 1;
  const size_t need = (size_t)JP_MAX_D + (size_t)(nE + 1) + (size_t)nb * (nE + 1);
  const bool in_scratch = need <= JP_SCRATCH_DOUBLES;
  double* hp = ctx->h_pinned;
  double* out = hp + JP_MAX_D;     // nE + 1 <= 64 + 64 * 65 / 2 + 1 doubles
  if (in_scratch) {
    d_beta = ctx->d_scratch;
    d_out = ctx->d_scratch + JP_MAX_D;
    d_work = d_out + (nE + 1);
  } else {
    JP_CUDA(jp_dmalloc(ctx, &d_beta, sizeof(double) * d));
    JP_CUDA(jp_dmalloc(ctx, &d_out, sizeof(double) * (nE + 1)));
    JP_CUDA(jp_dmalloc(ctx, &d_work, sizeof(double) * (size_t)nb * (nE + 1)));
  }
  JP_CUDA(jp_pinned_acquire(ctx));
  std::memcpy(hp, h_beta, sizeof(double) * d);
  JP_CUDA(cudaMemcpyAsync(d_beta, hp, sizeof(double) * d, cudaMemcpyHostToDevice, ctx->stream));
  int st = jp_glm_sums_device(ctx, data, d, d_beta, exchange ? d_out : out, d_work, nb);
  if (st == JP_OK && exchange) {
    const double* g = nullptr;
    st = jp_comm_exchange(comm, JP_CH_USER, d_out, nE + 1, &g);
    if (st == JP_OK) {
      jp_sum_gathered_kernel<<<(nE + 1 + 127) / 128, 128, 0, ctx->stream>>>(g, comm->world, nE + 1, out);
      ctx->launches++;
    }
  }
  if (st == JP_OK) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      jp_set_error("jp_glm_grad_hess: %s", cudaGetErrorString(e));
      st = JP_ERR_CUDA;
    }
  }
  if (!in_scratch) {
    jp_dfree(ctx, d_beta); jp_dfree(ctx, d_out); jp_dfree(ctx, d_work);
  }
  JP_TRY(st);
  // add the N(0, s^2) prior and unpack
  double s2 = data->hyper[0] * data->hyper[0];
  double lp = out[nE];
  for (int k = 0; k < d; ++k) {
    double z = h_beta[k] / data->hyper[0];
    lp += -0.5 * z * z - std::log(data->hyper[0]) - 0.5 * 1.8378770664093454835606594728112;
    if (h_g) h_g[k] = out[k] - h_beta[k] / s2;
  }
  if (h_Hneg) {
    int e = d;
    for (int c = 0; c < d; ++c)
      for (int r = 0; r <= c; ++r, ++e) {
        double h = out[e] + (r == c ? 1.0 / s2 : 0.0);
        h_Hneg[(size_t)c * d + r] = h;
        h_Hneg[(size_t)r * d + c] = h;
      }
  }
  if (h_logpost) *h_logpost = lp;
  return JP_OK;
}

extern "C" int jp_glm_grad_hess(jp_ctx* ctx, const jp_data* data, int d, const double* h_beta, double* h_g, double* h_Hneg,
                                double* h_logpost) {
  return jp_glm_grad_hess_comm(ctx, data, nullptr, d, h_beta, h_g, h_Hneg, h_logpost);
}
