// Diagnostic micro-benchmarks for the softmax side of the attention kernel (not on the product path):
//   1. MUFU ex2 throughput (f32, f16x2) per SM by warp count
//   2. a register-resident softmax inner loop (scale-sub FFMA, ex2 or FMA-pipe polynomial, f16x2 pack) by the
//      fraction of exponentials on the polynomial path: elements per clock per SM
//   3. tcgen05.ld bandwidth (32x32b.x32) by warp count
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/sm_probe.bin tools/sm_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../vista_b200/csrc/ptx.cuh"

using namespace vb;

__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// cheaper variant: no clamp (caller guarantees x >= -126), floor split through the magic add
__device__ __forceinline__ float exp2_poly2(float x) {
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t p;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
  return p;
}

template <int MODE>
__global__ void ex2_kernel(int iters, float* out, long long* cyc) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = ex2_f(v[i]) - 1.0f;                       // 1 MUFU + 1 FADD
      if (MODE == 1) {                                                // f16x2 MUFU
        uint32_t u = __float_as_uint(v[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
        v[i] = __uint_as_float(u);
      }
      if (MODE == 2) v[i] = exp2_poly2(v[i]) - 1.0f;
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// softmax inner loop on 64 register-resident "scores" per thread per iteration; POLY of every 8 on the FMA pipe
template <int POLY, int VARIANT>
__global__ void softmax_kernel(int iters, float scale, uint32_t* out, long long* cyc) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = -0.01f * ((threadIdx.x * 7 + i * 3) & 255);
  uint32_t acc = 0;
  float m = 0.5f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      asm("max.f32 %0, %0, %1, %2;" : "+f"(mx0) : "f"(s[i]), "f"(s[i + 1]));
      asm("max.f32 %0, %0, %1, %2;" : "+f"(mx1) : "f"(s[i + 2]), "f"(s[i + 3]));
    }
    m = fmaxf(m, fmaxf(mx0, mx1) * scale) * 0.999f;
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      const float x0 = fmaf(s[i], scale, -m), x1 = fmaf(s[i + 1], scale, -m);
      float p0, p1;
      if (VARIANT == 0) {
        p0 = ((i & 7) < POLY) ? exp2_poly(x0) : ex2_f(x0);
        p1 = (((i + 1) & 7) < POLY) ? exp2_poly(x1) : ex2_f(x1);
      } else {
        p0 = ((i & 7) < POLY) ? exp2_poly2(x0) : ex2_f(x0);
        p1 = (((i + 1) & 7) < POLY) ? exp2_poly2(x1) : ex2_f(x1);
      }
      const uint32_t w = pack_h2(p0, p1);
      acc ^= w;
      s[i] = s[i] * 0.9999f;    // keeps the loop from being hoisted; 0.5 extra FMUL per element
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(m);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void tmem_ld_kernel(int iters, uint32_t* out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(base + ((c * 32 + (warp >> 2) * 128) & 511), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 8) acc ^= v[i];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(slot);
  }
}

static double avg_cycles(long long* d, int n) {
  static long long h[1024];
  cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < n; ++i) s += (double)h[i];
  return s / n;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * 1024 * 4);
  cudaMalloc(&cyc, sms * 8);
  const int iters = 2000;
  printf("# sm_probe (%d SMs); all figures per SM\n\n## ex2 throughput (16 independent chains per thread)\n\n", sms);
  printf("| mode | warps/SM | elements/clk/SM |\n|---|---|---|\n");
  for (int warps : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) ex2_kernel<0><<<sms, warps * 32>>>(iters, out, cyc);
        if (mode == 1) ex2_kernel<1><<<sms, warps * 32>>>(iters, out, cyc);
        if (mode == 2) ex2_kernel<2><<<sms, warps * 32>>>(iters, out, cyc);
        cudaDeviceSynchronize();
      }
      const double c = avg_cycles(cyc, sms);
      const double el = (double)iters * 16 * warps * 32 * (mode == 1 ? 2 : 1);
      printf("| %s | %d | %.1f |\n", mode == 0 ? "ex2.f32 (+FADD)" : mode == 1 ? "ex2.f16x2" : "poly2 (+FADD)", warps, el / c);
    }
  }
  printf("\n## softmax inner loop (max3 + FFMA + exp + f16x2 pack), 64 elements per thread per iteration\n\n");
  printf("| poly of 8 | variant | warps/SM | elements/clk/SM | cycles per 128x128 block |\n|---|---|---|---|---|\n");
#define RUN_SM(P, V)                                                                             \
  for (int warps : {8, 16}) {                                                                    \
    for (int rep = 0; rep < 2; ++rep) {                                                          \
      softmax_kernel<P, V><<<sms, warps * 32>>>(400, 0.18f, (uint32_t*)out, cyc);                \
      cudaDeviceSynchronize();                                                                   \
    }                                                                                            \
    const double c = avg_cycles(cyc, sms);                                                       \
    const double el = 400.0 * 64 * warps * 32;                                                   \
    printf("| %d | %d | %d | %.1f | %.0f |\n", P, V, warps, el / c, 16384.0 / (el / c));         \
  }
  RUN_SM(0, 0) RUN_SM(1, 0) RUN_SM(2, 0) RUN_SM(3, 0) RUN_SM(4, 0)
  RUN_SM(1, 1) RUN_SM(2, 1) RUN_SM(3, 1) RUN_SM(4, 1)
  printf("\n## tcgen05.ld 32x32b.x32 (+wait) throughput\n\n| warps/SM | bytes/clk/SM |\n|---|---|\n");
  for (int warps : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) {
      tmem_ld_kernel<<<sms, warps * 32>>>(2000, (uint32_t*)out, cyc);
      cudaDeviceSynchronize();
    }
    const double c = avg_cycles(cyc, sms);
    printf("| %d | %.1f |\n", warps, 2000.0 * 4 * 32 * 32 * 4 * warps / c);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
