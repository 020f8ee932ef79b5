// soundml-io's layout pass on the device: a decoded block
// arrives interleaved, [frames][channels], exactly as sf_readf_float/double fills
// the reference's staging block, and leaves planar (channel c at out + c * total)
// or downmixed to mono -- soundml_io_stubs.c:832-872 (soundml_io_read_planar_*).
// The arithmetic of the downmix is the reference's: channels added in order in the
// sample type, then one multiply by 1 / channels, no contraction.
#include "kernels.h"

namespace smb {

namespace {

template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }

// One thread per frame: its `channels` samples are contiguous, so a warp reads a
// contiguous run of the block and writes `channels` contiguous runs (planar) or one.
template <typename T, bool DOWNMIX>
__global__ void ingest_layout_kernel(const T* __restrict__ in, long long frames, int channels,
                                     T* __restrict__ out, long long out_total) {
  const T inv = (T)1 / (T)channels;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < frames;
       i += (long long)gridDim.x * blockDim.x) {
    const T* fr = in + i * channels;
    if (DOWNMIX) {
      T acc = (T)0;
      for (int c = 0; c < channels; ++c) acc = add_rn(acc, fr[c]);
      out[i] = mul_rn(acc, inv);
    } else {
      for (int c = 0; c < channels; ++c) out[(long long)c * out_total + i] = fr[c];
    }
  }
}

}  // namespace

cudaError_t launch_ingest_layout(const void* in, int dtype, long long frames, int channels,
                                 int downmix, void* out, long long out_total, cudaStream_t st) {
  if (frames == 0) return cudaSuccess;
  const int threads = 256;
  long long want = (frames + threads - 1) / threads;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  if (dtype == 0) {
    if (downmix) ingest_layout_kernel<float, true><<<grid, threads, 0, st>>>((const float*)in, frames, channels, (float*)out, out_total);
    else ingest_layout_kernel<float, false><<<grid, threads, 0, st>>>((const float*)in, frames, channels, (float*)out, out_total);
  } else {
    if (downmix) ingest_layout_kernel<double, true><<<grid, threads, 0, st>>>((const double*)in, frames, channels, (double*)out, out_total);
    else ingest_layout_kernel<double, false><<<grid, threads, 0, st>>>((const double*)in, frames, channels, (double*)out, out_total);
  }
  ++g_launch_count;
  return cudaGetLastError();
}

}  // namespace smb
