// Microbenchmark: the GLM epilogue's per-column-pair instruction sequence (NC = 6) on opaque register inputs,
// without TMEM / barriers: what FMA-pipe rate does the sequence itself reach?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t f32x2_pack(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ uint64_t f32x2_bcast(float x) { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t f32x2_mul(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float lo32(uint64_t v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// VAR 0: packed, as in the kernel.  VAR 1: scalar FFMA (same arithmetic, 2x instructions)
template <int NC, int VAR>
__device__ __forceinline__ void acc8(const uint32_t (&v)[8], const float (&c)[NC], uint64_t* accE, uint64_t* accO) {
  if (VAR == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t D = f32x2_pack(v[2 * j], v[2 * j + 1]);
      const uint64_t D2 = f32x2_mul(D, D);
      uint64_t pe = f32x2_bcast(c[NC - 1]), po = f32x2_bcast(c[NC - 2]);
#pragma unroll
      for (int m = NC / 2 - 2; m >= 0; --m) {
        pe = f32x2_fma(pe, D2, f32x2_bcast(c[2 * m + 1]));
        po = f32x2_fma(po, D2, f32x2_bcast(c[2 * m]));
      }
      accE[j] = f32x2_fma(f32x2_mul(D2, D2), pe, accE[j]);
      accO[j] = f32x2_fma(f32x2_mul(D2, D), po, accO[j]);
    }
  } else {
    float* aE = reinterpret_cast<float*>(accE);
    float* aO = reinterpret_cast<float*>(accO);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float D = __uint_as_float(v[j]);
      const float D2 = D * D;
      float pe = c[NC - 1], po = c[NC - 2];
#pragma unroll
      for (int m = NC / 2 - 2; m >= 0; --m) {
        pe = fmaf(pe, D2, c[2 * m + 1]);
        po = fmaf(po, D2, c[2 * m]);
      }
      aE[j] = fmaf(D2 * D2, pe, aE[j]);
      aO[j] = fmaf(D2 * D, po, aO[j]);
    }
  }
}

#define ITERS 2048
template <int NC, int VAR>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, const float* coef) {
  uint64_t accE[16], accO[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) accE[j] = accO[j] = 0ull;
  float c[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) c[i] = coef[threadIdx.x * 12 + i];
  uint32_t va[8], vb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { va[i] = __float_as_uint(1e-2f * threadIdx.x + i); vb[i] = __float_as_uint(2e-2f * threadIdx.x - i); }
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < 4 * 8 * 512; i += blockDim.x) sm[i] = __float_as_uint(1e-3f * (i % 977));
  __syncthreads();
  auto lds8 = [&](uint32_t (&v)[8], int ch) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(sm + ((size_t)ch * 512 + threadIdx.x) * 8);
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a));
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a + 16));
  };
  const long long t0 = clock64();
  lds8(va, 0);
  for (int it = 0; it < ITERS; ++it) {
    // one "tile": 32 columns in four chunks of 8, inputs prefetched one chunk ahead (as the TMEM loads are)
    lds8(vb, 1);
    acc8<NC, VAR>(va, c, accE, accO);
    lds8(va, 2);
    acc8<NC, VAR>(vb, c, accE + 4, accO + 4);
    lds8(vb, 3);
    acc8<NC, VAR>(va, c, accE + 8, accO + 8);
    lds8(va, 0);
    acc8<NC, VAR>(vb, c, accE + 12, accO + 12);
  }
  const long long t1 = clock64();
  float res = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) res += lo32(accE[j]) + lo32(accO[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = res;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NC, int VAR>
void run(const char* name) {
  float *out, *coef; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&coef, 1024 * 12 * 4);
  cudaMemset(coef, 0, 1024 * 12 * 4);
  cudaFuncSetAttribute(k<NC, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int threads = 128; threads <= 512; threads += 128) {
    k<NC, VAR><<<148, threads, 100 * 1024>>>(out, cyc, coef);
    cudaDeviceSynchronize();
    k<NC, VAR><<<148, threads, 100 * 1024>>>(out, cyc, coef);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
    const double w = threads / 128.0;
    const double lane_ops = (double)ITERS * 32 /*columns*/ * (NC + 3) * 32 /*lanes*/ * w;   // per SMSP
    printf("%-34s NC %d warps/SMSP %.0f cycles %9.0f  FP32 lane-ops/clk/SMSP %6.2f (%5.1f%% of 32) %s\n", name, NC, w, cy,
           lane_ops / cy, lane_ops / cy / 32 * 100, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}
int main() {
  run<6, 0>("packed FFMA2 sequence");
  run<6, 1>("scalar FFMA sequence");
  run<4, 0>("packed FFMA2 sequence");
  run<4, 1>("scalar FFMA sequence");
  return 0;
}
