// Issue rate of packed float32 (FFMA2 / FADD2) against scalar FFMA on sm_100a:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/bench_f32x2 tools/bench_f32x2.cu && /tmp/bench_f32x2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[16];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  const float m = seed * 0.999f;
  const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, seed);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, m2);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 4096;
  for (int warps = 4; warps <= 16; warps *= 2) {
    float c0, c1;
    k<0><<<148, warps * 32>>>(d, iters, 1.0f); cudaMemcpy(&c0, d, 4, cudaMemcpyDeviceToHost);
    k<1><<<148, warps * 32>>>(d, iters, 1.0f); cudaMemcpy(&c1, d, 4, cudaMemcpyDeviceToHost);
    // per scheduler: warps/4 warps, each 16 scalar or 8 packed instructions per iteration
    printf("warps/SM %2d: FFMA %.3f inst/clk/SMSP (%.1f flop-lanes), FFMA2 %.3f inst/clk/SMSP (%.1f flop-lanes)\n", warps,
           16.0 * iters * (warps / 4) / c0, 16.0 * iters * (warps / 4) / c0 * 32, 8.0 * iters * (warps / 4) / c1,
           8.0 * iters * (warps / 4) / c1 * 64);
  }
  return 0;
}
