// Microbenchmark: issue rate of packed FFMA2 / scalar FFMA on sm_100a for operand patterns like the GLM epilogue.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run: ./ffma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t pack(float lo, float hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ float lo32(uint64_t v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

#define ITERS 4096
// MODE 0: scalar FFMA, 16 independent chains (acc = acc * x + y), x,y per-chain registers
// MODE 1: FFMA2 acc[i] = x[i] * y[i] + acc[i]   (3 distinct 64-bit operands)
// MODE 2: FFMA2 acc[i] = acc[i] * X + c (X shared 64-bit, c scalar broadcast)
// MODE 3: the epilogue's 9-op column-pair sequence (NC = 6), 4 column pairs interleaved
// MODE 4: same arithmetic as 3 with scalar FFMA/FMUL (two columns = 18 ops)
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  const int tid = threadIdx.x;
  float f = seed + tid * 1e-3f;
  long long t0 = 0, t1 = 0;
  float res = 0;
  if (MODE == 0) {
    float a[16], x[16], y[16];
    for (int i = 0; i < 16; ++i) { a[i] = f + i; x[i] = 1.0f + 1e-6f * i; y[i] = 1e-3f * i; }
    t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x[i]), "f"(y[i]));
    }
    t1 = clock64();
    for (int i = 0; i < 16; ++i) res += a[i];
  } else if (MODE == 1) {
    uint64_t a[8], x[8], y[8];
    for (int i = 0; i < 8; ++i) { a[i] = pack(f + i, f - i); x[i] = pack(1.0f + 1e-6f * i, 1.0f - 1e-6f * i); y[i] = pack(1e-3f * i, 1e-4f * i); }
    t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma2(x[i], y[i], a[i]);
    }
    t1 = clock64();
    for (int i = 0; i < 8; ++i) res += lo32(a[i]);
  } else if (MODE == 2) {
    uint64_t a[8];
    uint64_t X = pack(1.0f + 1e-6f * f, 1.0f - 1e-6f * f);
    float c = 1e-3f * f;
    for (int i = 0; i < 8; ++i) a[i] = pack(f + i, f - i);
    t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], X, pack(c, c));
    }
    t1 = clock64();
    for (int i = 0; i < 8; ++i) res += lo32(a[i]);
  } else if (MODE == 3) {
    uint64_t accE[4], accO[4], D[4];
    float c[6];
    for (int i = 0; i < 6; ++i) c[i] = 1e-2f * (i + 1) + 1e-5f * f;
    for (int i = 0; i < 4; ++i) { accE[i] = pack(0.f, 0.f); accO[i] = pack(0.f, 0.f); D[i] = pack(1e-2f * f + i, 1e-2f * f - i); }
    t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint64_t D2 = mul2(D[i], D[i]);
        uint64_t pe = fma2(pack(c[5], c[5]), D2, pack(c[3], c[3]));
        uint64_t po = fma2(pack(c[4], c[4]), D2, pack(c[2], c[2]));
        pe = fma2(pe, D2, pack(c[1], c[1]));
        po = fma2(po, D2, pack(c[0], c[0]));
        accE[i] = fma2(mul2(D2, D2), pe, accE[i]);
        accO[i] = fma2(mul2(D2, D[i]), po, accO[i]);
      }
    }
    t1 = clock64();
    for (int i = 0; i < 4; ++i) res += lo32(accE[i]) + lo32(accO[i]);
  } else {
    float accE[8], accO[8], D[8];
    float c[6];
    for (int i = 0; i < 6; ++i) c[i] = 1e-2f * (i + 1) + 1e-5f * f;
    for (int i = 0; i < 8; ++i) { accE[i] = 0.f; accO[i] = 0.f; D[i] = 1e-2f * f + i; }
    t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float D2, pe, po, D4, D3;
        asm volatile("mul.rn.f32 %0, %1, %1;" : "=f"(D2) : "f"(D[i]));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(pe) : "f"(c[5]), "f"(D2), "f"(c[3]));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(po) : "f"(c[4]), "f"(D2), "f"(c[2]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(pe) : "f"(D2), "f"(c[1]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(po) : "f"(D2), "f"(c[0]));
        asm volatile("mul.rn.f32 %0, %1, %1;" : "=f"(D4) : "f"(D2));
        asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(D3) : "f"(D2), "f"(D[i]));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(accE[i]) : "f"(D4), "f"(pe));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(accO[i]) : "f"(D3), "f"(po));
      }
    }
    t1 = clock64();
    for (int i = 0; i < 8; ++i) res += accE[i] + accO[i];
  }
  out[blockIdx.x * blockDim.x + tid] = res;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double lane_ops_per_thread_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int threads = 128; threads <= 768; threads += 128) {
    k<MODE><<<148, threads, 100 * 1024>>>(out, cyc, 1.0f);   // 100 KB smem: one CTA per SM
    cudaDeviceSynchronize();
    k<MODE><<<148, threads, 100 * 1024>>>(out, cyc, 1.0f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double warps_per_smsp = threads / 128.0;
    // FP32 lane-operations per cycle per SMSP (peak 32)
    double rate = lane_ops_per_thread_iter * ITERS * 32.0 * warps_per_smsp / c;
    printf("%-44s warps/SMSP %.0f  cycles %9.0f  lane-ops/clk/SMSP %6.2f (%5.1f%% of 32)  %s\n", name, warps_per_smsp, c, rate, rate / 32 * 100,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  run<0>("scalar FFMA, 16 chains", 16);
  run<1>("FFMA2 x*y+acc, 3 distinct 64-bit operands", 32);
  run<2>("FFMA2 acc*X+c (shared X, scalar c)", 32);
  run<3>("epilogue sequence, FFMA2/FMUL2 (4 col pairs)", 4 * 9 * 2);
  run<4>("epilogue sequence, scalar FFMA/FMUL (8 cols)", 8 * 9);
  return 0;
}
