// jp_sort.cuh -- stable LSD radix sort building block (8-bit digits) used by stage 1 (duplicate
// node merge on integer keys) and stage 5 (sort of f(theta) values for the weighted CDF).
//
// One pass = histogram -> exclusive scan -> stable scatter of a permutation array.  The digit of
// element i is produced by a functor from the SOURCE index perm_in[i], so the keys themselves never
// move (they stay L2-resident at the sizes this library sees: <= a few 10^6 elements).
// blockIdx.y selects an independent batch (one marginal) so K sorts share every launch.
//
// Stability: a block owns a contiguous tile, a warp owns a contiguous sub-tile of it and walks it
// in order 32 elements at a time; ranks inside a 32-element step come from __match_any_sync.
#pragma once
#include "jp_common.cuh"

#define JP_SORT_THREADS 256
#define JP_SORT_ITEMS 8
#define JP_SORT_TILE (JP_SORT_THREADS * JP_SORT_ITEMS)
#define JP_SORT_WARPS (JP_SORT_THREADS / 32)
#define JP_SORT_BINS 256

static inline int jp_sort_blocks(long long n) { return (int)((n + JP_SORT_TILE - 1) / JP_SORT_TILE); }

#ifdef __CUDACC__
// counts digits of the calling warp's sub-tile into cnt[JP_SORT_BINS] (shared, zeroed by caller)
template <class DigitFn>
__device__ __forceinline__ void jp_sort_warp_count(const DigitFn& f, int batch, const uint32_t* perm_in,
                                                   long long n, long long warp_begin, uint32_t* cnt) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 0; s < JP_SORT_ITEMS; ++s) {
    long long i = warp_begin + s * 32 + lane;
    unsigned dg = (i < n) ? f(batch, perm_in[i]) : 0xFFFFu;
    unsigned peers = __match_any_sync(0xffffffffu, dg);
    if (dg != 0xFFFFu && (peers & ((1u << lane) - 1)) == 0) cnt[dg] += __popc(peers);
    __syncwarp();
  }
}

template <class DigitFn>
__global__ void __launch_bounds__(JP_SORT_THREADS)
jp_radix_hist_kernel(DigitFn f, const uint32_t* __restrict__ perm_in, long long n, long long perm_stride,
                     uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t cnt[JP_SORT_WARPS][JP_SORT_BINS];
  const int batch = blockIdx.y;
  const int w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < JP_SORT_WARPS * JP_SORT_BINS; i += JP_SORT_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t* pin = perm_in + (size_t)batch * perm_stride;
  long long warp_begin = (long long)blockIdx.x * JP_SORT_TILE + (long long)w * (32 * JP_SORT_ITEMS);
  jp_sort_warp_count(f, batch, pin, n, warp_begin, cnt[w]);
  __syncthreads();
  for (int b = threadIdx.x; b < JP_SORT_BINS; b += JP_SORT_THREADS) {
    uint32_t s = 0;
#pragma unroll
    for (int ww = 0; ww < JP_SORT_WARPS; ++ww) s += cnt[ww][b];
    hist[(size_t)batch * JP_SORT_BINS * nblocks + (size_t)b * nblocks + blockIdx.x] = s;
  }
}

// exclusive scan over the bin-major histogram (length 256 * nblocks) of one batch; one block per batch
__global__ void __launch_bounds__(1024) jp_radix_scan_kernel(uint32_t* __restrict__ hist, int nblocks);

template <class DigitFn>
__global__ void __launch_bounds__(JP_SORT_THREADS)
jp_radix_scatter_kernel(DigitFn f, const uint32_t* __restrict__ perm_in, uint32_t* __restrict__ perm_out,
                        long long n, long long perm_stride, const uint32_t* __restrict__ offs, int nblocks) {
  __shared__ uint32_t cnt[JP_SORT_WARPS][JP_SORT_BINS];   // per-warp counts, then per-warp bases
  __shared__ uint32_t run[JP_SORT_WARPS][JP_SORT_BINS];   // running counters during the scatter
  const int batch = blockIdx.y;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < JP_SORT_WARPS * JP_SORT_BINS; i += JP_SORT_THREADS) {
    (&cnt[0][0])[i] = 0;
    (&run[0][0])[i] = 0;
  }
  __syncthreads();
  const uint32_t* pin = perm_in + (size_t)batch * perm_stride;
  uint32_t* pout = perm_out + (size_t)batch * perm_stride;
  long long warp_begin = (long long)blockIdx.x * JP_SORT_TILE + (long long)w * (32 * JP_SORT_ITEMS);
  jp_sort_warp_count(f, batch, pin, n, warp_begin, cnt[w]);
  __syncthreads();
  for (int b = threadIdx.x; b < JP_SORT_BINS; b += JP_SORT_THREADS) {
    uint32_t base = offs[(size_t)batch * JP_SORT_BINS * nblocks + (size_t)b * nblocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < JP_SORT_WARPS; ++ww) {
      uint32_t c = cnt[ww][b];
      cnt[ww][b] = base;
      base += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < JP_SORT_ITEMS; ++s) {
    long long i = warp_begin + s * 32 + lane;
    uint32_t src = (i < n) ? pin[i] : 0u;
    unsigned dg = (i < n) ? f(batch, src) : 0xFFFFu;
    unsigned peers = __match_any_sync(0xffffffffu, dg);
    unsigned below = peers & ((1u << lane) - 1);
    if (dg != 0xFFFFu) {
      uint32_t pos = cnt[w][dg] + run[w][dg] + __popc(below);
      pout[pos] = src;
    }
    __syncwarp();
    if (dg != 0xFFFFu && below == 0) run[w][dg] += __popc(peers);
    __syncwarp();
  }
}

__global__ void jp_iota_kernel(uint32_t* __restrict__ perm, long long n, long long stride);

// one stable pass over `nbatch` independent arrays of length n
template <class DigitFn>
static int jp_radix_pass(jp_ctx* ctx, DigitFn f, const uint32_t* perm_in, uint32_t* perm_out, long long n,
                         long long perm_stride, uint32_t* hist, int nbatch) {
  int nb = jp_sort_blocks(n);
  dim3 grid(nb, nbatch);
  jp_radix_hist_kernel<<<grid, JP_SORT_THREADS, 0, ctx->stream>>>(f, perm_in, n, perm_stride, hist, nb);
  JP_CHECK_LAUNCH(ctx);
  jp_radix_scan_kernel<<<nbatch, 1024, 0, ctx->stream>>>(hist, nb);
  JP_CHECK_LAUNCH(ctx);
  jp_radix_scatter_kernel<<<grid, JP_SORT_THREADS, 0, ctx->stream>>>(f, perm_in, perm_out, n, perm_stride, hist, nb);
  JP_CHECK_LAUNCH(ctx);
  return JP_OK;
}
#endif
