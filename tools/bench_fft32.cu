// Micro-benchmark: how fast does the register fft32 of the fused STFT kernel issue
// on its own?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I soundml_b200/csrc
//   tools/bench_fft32.cu -o gpurun_out/bench_fft32 && gpurun_out/bench_fft32
#include <cstdio>
#include "fft32.cuh"
using namespace smb::fft32impl;

template <int PACKED>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
  float2 a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  for (int it = 0; it < iters; ++it) {
    if (PACKED) fft32_packed(a); else fft32_scalar(a);
#pragma unroll
    for (int i = 0; i < 32; ++i) { a[i].x *= 0.17f; a[i].y *= 0.17f; }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 512 * 4);
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int packed = 0; packed < 2; ++packed) {
    if (packed) k<1><<<148, 512>>>(d, 10); else k<0><<<148, 512>>>(d, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    if (packed) k<1><<<148, 512>>>(d, iters); else k<0><<<148, 512>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // one fft32 per thread per iteration
    const double ffts = 148.0 * 512 * iters;
    printf("fft32 %s: %.3f ms, %.1f ns per warp-fft32, %.1f cycles per warp-fft32 per SMSP at 1.965 GHz\n",
           packed ? "packed (f32x2)" : "scalar", ms, ms * 1e6 / (ffts / 32),
           ms * 1e-3 * 1.965e9 / (ffts / 32 / (148 * 4)));
  }
  return 0;
}
