// tmem_ld.cu -- tcgen05.ld (TMEM -> registers) throughput on B200, the read floor of jp_glm_tc_kernel's epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
// One CTA per SM allocates all 512 TMEM columns; W warps (W/4 per sub-partition, each on its lane quarter) issue
// tcgen05.ld.32x32b.xN in a loop with D loads in flight before a tcgen05.wait::ld.  Prints bytes / clock / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> struct Ld;
template <> struct Ld<8> {
  static __device__ __forceinline__ void ld(uint32_t (&v)[8], uint32_t a) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a) : "memory");
  }
};
template <> struct Ld<16> {
  static __device__ __forceinline__ void ld(uint32_t (&v)[16], uint32_t a) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a) : "memory");
  }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ void ld(uint32_t (&v)[32], uint32_t a) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(a) : "memory");
  }
};

// N columns per load, D loads in flight per wait, FMA2 packed fma.rn.f32x2 per loaded column pair (0: pure read)
template <int N, int D, int FMA2>
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[D][N];
  unsigned long long acc[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) acc[j] = 0ull;
  uint32_t x = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < D; ++k) Ld<N>::ld(v[k], base + (uint32_t)(((it * D + k) * N) & 511 & ~(N - 1)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < D; ++k) {
      if (FMA2 == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) x ^= v[k][j];
      } else {
#pragma unroll
        for (int j = 0; j < N / 2; ++j) {
          unsigned long long Dv;
          asm("mov.b64 %0, {%1, %2};" : "=l"(Dv) : "r"(v[k][2 * j]), "r"(v[k][2 * j + 1]));
#pragma unroll
          for (int f = 0; f < FMA2; ++f) asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc[j]) : "l"(Dv));
        }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = __uint_as_float(x);
#pragma unroll
  for (int j = 0; j < N / 2; ++j) s += (float)(acc[j] & 0xffff);
  if (s == 123.456f) sink[0] = s;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int N, int D, int FMA2>
static void run(int warps, long long* d_cyc, float* d_sink) {
  const int iters = 4096 / D;
  tmem_kernel<N, D, FMA2><<<148, warps * 32>>>(iters, d_cyc, d_sink);
  tmem_kernel<N, D, FMA2><<<148, warps * 32>>>(iters, d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i] / 148;
  const double bytes = (double)warps * iters * D * N * 128.0;
  printf("x%-2d  in flight %d  warps %2d (%d / sub-partition)  FFMA2 per col pair %d   %7.1f B/clk/SM   %6.1f clk per warp-load   %s\n", N, D,
         warps, warps / 4, FMA2, bytes / mean, mean / (iters * D), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_cyc;
  float* d_sink;
  cudaMalloc(&d_cyc, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  for (int w = 4; w <= 16; w += 4) {
    run<8, 1, 0>(w, d_cyc, d_sink);
    run<8, 2, 0>(w, d_cyc, d_sink);
    run<8, 4, 0>(w, d_cyc, d_sink);
    run<16, 1, 0>(w, d_cyc, d_sink);
    run<16, 2, 0>(w, d_cyc, d_sink);
    run<32, 1, 0>(w, d_cyc, d_sink);
    run<32, 2, 0>(w, d_cyc, d_sink);
  }
  // read + the epilogue's arithmetic density (7 packed operations per column pair), 12 warps as in the kernel
  for (int w = 8; w <= 16; w += 4) {
    run<8, 2, 7>(w, d_cyc, d_sink);
    run<16, 2, 7>(w, d_cyc, d_sink);
    run<8, 2, 4>(w, d_cyc, d_sink);
  }
  return 0;
}
