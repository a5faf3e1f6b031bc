// jp_glm.cu -- GLM score / observed information on the GPU (FP64).
//
// Upstream of the five stages: this is what `mode` needs from the data (reference
// src/joint_posterior.jl:164-168: BFGS + ForwardDiff Hessian there; Newton/IRLS here because for a
// GLM the Hessian is the closed form X' W X).  The same sums (X'(y - mu_hat), X' W X at the mode) are
// what the tensor-core log-density path centres its expansion on (jp_glm_tc.cu).
//
// Layout: a block stages a tile of observation records in shared memory, computes the per-record
// residual and weight once, then each thread owns a few entries of the packed (gradient, upper
// Hessian) vector and accumulates them over the tile.  Per-block partials are combined in block
// order by a second kernel, so the result is deterministic.
#include <algorithm>
#include <vector>
#include "jp_common.cuh"

#define JP_GLM_THREADS 256
#define JP_GLM_TILE 128    // observations per tile

__global__ void __launch_bounds__(JP_GLM_THREADS)
jp_glm_partials_kernel(int family, int d, long long N, const double* __restrict__ obs, const double* __restrict__ beta,
                       int nE, double* __restrict__ part /* [gridDim.x][nE + 1] */) {
  extern __shared__ double sh[];
  const int ncols = d + 1;
  double* tile = sh;                               // JP_GLM_TILE x ncols
  double* res = tile + JP_GLM_TILE * ncols;        // JP_GLM_TILE
  double* wgt = res + JP_GLM_TILE;                 // JP_GLM_TILE
  double* s_beta = wgt + JP_GLM_TILE;              // d
  __shared__ double red[33];
  for (int k = threadIdx.x; k < d; k += JP_GLM_THREADS) s_beta[k] = beta[k];
  // entries owned by this thread: e = threadIdx.x + j * JP_GLM_THREADS; at most 9 for d = 64
  double acc[9];
  int er[9], ec[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    acc[j] = 0.0;
    int e = threadIdx.x + j * JP_GLM_THREADS;
    er[j] = -1; ec[j] = -1;
    if (e < d) {
      er[j] = e; ec[j] = -2;                        // gradient entry
    } else if (e < nE) {
      int t = e - d, c = 0;                          // packed upper triangle, column by column
      while (t >= c + 1) { t -= c + 1; ++c; }
      er[j] = t; ec[j] = c;
    }
  }
  double ll = 0.0;
  for (long long base = (long long)blockIdx.x * JP_GLM_TILE; base < N; base += (long long)gridDim.x * JP_GLM_TILE) {
    int cnt = (int)min((long long)JP_GLM_TILE, N - base);
    __syncthreads();
    const double* src = obs + (size_t)base * ncols;
    for (int i = threadIdx.x; i < cnt * ncols; i += JP_GLM_THREADS) tile[i] = __ldg(src + i);
    __syncthreads();
    if (threadIdx.x < cnt) {
      const double* r = tile + threadIdx.x * ncols;
      double eta = 0;
      for (int k = 0; k < d; ++k) eta += r[k] * s_beta[k];
      double mu, wv;
      if (family == JP_FAM_LOGISTIC) {
        mu = 1.0 / (1.0 + exp(-eta));
        wv = mu * (1.0 - mu);
        ll += r[d] * eta - (fmax(eta, 0.0) + log1p(exp(-fabs(eta))));
      } else {
        mu = exp(eta);
        wv = mu;
        ll += r[d] * eta - mu;
      }
      res[threadIdx.x] = r[d] - mu;
      wgt[threadIdx.x] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      if (er[j] < 0) continue;
      double a = acc[j];
      if (ec[j] == -2) {
        for (int n = 0; n < cnt; ++n) a += res[n] * tile[n * ncols + er[j]];
      } else {
        for (int n = 0; n < cnt; ++n) a += wgt[n] * tile[n * ncols + er[j]] * tile[n * ncols + ec[j]];
      }
      acc[j] = a;
    }
  }
  double* o = part + (size_t)blockIdx.x * (nE + 1);
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    int e = threadIdx.x + j * JP_GLM_THREADS;
    if (e < nE) o[e] = acc[j];
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
int jp_glm_sums_device(jp_ctx* ctx, const jp_data* data, int d, const double* d_beta, double* d_out, double* d_work,
                       int nblocks) {
  int nE = d + d * (d + 1) / 2;
  size_t smem = (size_t)(JP_GLM_TILE * (d + 1) + 2 * JP_GLM_TILE + d) * sizeof(double);
  if (smem > 48 * 1024) {
    JP_CUDA(cudaFuncSetAttribute(jp_glm_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  jp_glm_partials_kernel<<<nblocks, JP_GLM_THREADS, smem, ctx->stream>>>(data->family, d, data->N, data->d_obs, d_beta,
                                                                           nE, d_work);
  JP_CHECK_LAUNCH(ctx);
  jp_glm_combine_kernel<<<(nE + 1 + 7) / 8, 256, 0, ctx->stream>>>(nblocks, nE + 1, d_work, d_out);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}

int jp_glm_num_blocks(const jp_ctx* ctx, long long N) {
  long long tiles = (N + JP_GLM_TILE - 1) / JP_GLM_TILE;
  return (int)std::max(1LL, std::min(tiles, (long long)ctx->sm_count * 4));
}

extern "C" int jp_glm_grad_hess(jp_ctx* ctx, const jp_data* data, int d, const double* h_beta, double* h_g,
                                double* h_Hneg, double* h_logpost) {
  JP_REQUIRE(ctx && data && h_beta, "jp_glm_grad_hess: null argument");
  JP_REQUIRE(data->family == JP_FAM_LOGISTIC || data->family == JP_FAM_POISSON,
             "jp_glm_grad_hess: family %d is not a GLM", data->family);
  JP_REQUIRE(d >= 1 && d <= JP_MAX_D && data->ncols == d + 1, "jp_glm_grad_hess: d=%d does not match %d columns", d,
             data->ncols);
  int nE = d + d * (d + 1) / 2;
  int nb = jp_glm_num_blocks(ctx, data->N);
  double *d_beta = nullptr, *d_out = nullptr, *d_work = nullptr;
  JP_CUDA(jp_dmalloc(ctx, &d_beta, sizeof(double) * d));
  JP_CUDA(jp_dmalloc(ctx, &d_out, sizeof(double) * (nE + 1)));
  JP_CUDA(jp_dmalloc(ctx, &d_work, sizeof(double) * (size_t)nb * (nE + 1)));
  JP_CUDA(cudaMemcpyAsync(d_beta, h_beta, sizeof(double) * d, cudaMemcpyHostToDevice, ctx->stream));
  int st = jp_glm_sums_device(ctx, data, d, d_beta, d_out, d_work, nb);
  std::vector<double> out(nE + 1);
  if (st == JP_OK) {
    cudaError_t e = cudaMemcpyAsync(out.data(), d_out, sizeof(double) * (nE + 1), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      jp_set_error("jp_glm_grad_hess: %s", cudaGetErrorString(e));
      st = JP_ERR_CUDA;
    }
  }
  jp_dfree(ctx, d_beta); jp_dfree(ctx, d_out); jp_dfree(ctx, d_work);
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
