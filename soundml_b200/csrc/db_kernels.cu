// Decibel conversion and MFCC epilogue kernels (what ML
// consumers ingest downstream of the mel spectrogram).
//
//   to_db   convert.ml:20-56   floor at amin, log / scale / offset in the input's
//           own dtype, optional whole-tensor top_db clamp (one max reduction)
//   mfcc    soundml.ml:50-95   log-mel with the 80 dB clamp, raw type-II DCT along
//           the mel axis, orthonormal scales, lifter -- double interior, one
//           rounding
#include "kernels.h"

namespace smb {

namespace {

// order-preserving map of doubles onto unsigned 64-bit keys (for atomicMax)
__device__ __forceinline__ unsigned long long key_of(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ void block_max_to(double v, unsigned long long* slot) {
  unsigned long long k = key_of(v);
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(slot, k);
}

template <typename T> __device__ __forceinline__ T log_t(T v);
template <> __device__ __forceinline__ float log_t<float>(float v) { return logf(v); }
template <> __device__ __forceinline__ double log_t<double>(double v) { return log(v); }

// db = log(max(|x| or x, amin)) * scale - offset, every operation in T; the
// running maximum of db goes to *max_slot when the clamp is requested.
template <typename T>
__global__ void to_db_kernel(const T* __restrict__ x, long long count, int magnitude, T amin,
                             T scale, T offset, unsigned long long* max_slot,
                             T* __restrict__ out) {
  double local = -1.0e308;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    T v = x[i];
    if (magnitude) v = v < T(0) ? -v : v;
    v = v > amin ? v : amin;
    const T db = log_t<T>(v) * scale - offset;
    out[i] = db;
    local = fmax(local, (double)db);
  }
  if (max_slot) block_max_to(local, max_slot);
}

template <typename T>
__global__ void clamp_db_kernel(T* __restrict__ out, long long count,
                                const unsigned long long* max_slot, double range) {
  const T floor_db = (T)(value_of(*max_slot) - range);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const T v = out[i];
    out[i] = v > floor_db ? v : floor_db;
  }
}

template <typename T>
__global__ void max_value_kernel(const T* __restrict__ x, long long count,
                                 unsigned long long* max_slot) {
  double local = -1.0e308;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    local = fmax(local, (double)x[i]);
  block_max_to(local, max_slot);
}

constexpr int kMfccFrames = 32;     // frame columns per CTA
constexpr int kMfccThreads = 256;

// mel [batch, n_mels, frames] (T) -> out [batch, n_mfcc, frames] (T).  One CTA per
// (32-frame tile, signal): the log-mel tile is built once in shared memory in
// double (with the 80 dB clamp below the whole-tensor maximum), then thread
// (coefficient, frame) walks the mel axis against the DCT table.
template <typename T>
__global__ void __launch_bounds__(kMfccThreads)
mfcc_kernel(const T* __restrict__ mel, int n_mels, long long frames, int n_mfcc,
            const double* __restrict__ dct, const unsigned long long* max_mel_slot, double amin,
            double scale, double offset, double range, T* __restrict__ out) {
  extern __shared__ __align__(8) double sDb[];           // [n_mels][kMfccFrames]
  const long long b = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * kMfccFrames;
  const int nf = (int)min((long long)kMfccFrames, frames - p0);
  // the maximum of db is db of the maximum (the map is monotone)
  const double vmax = fmax(value_of(*max_mel_slot), amin);
  const double floor_db = (log(vmax) * scale - offset) - range;
  const T* src = mel + b * n_mels * frames + p0;
  for (int i = threadIdx.x; i < n_mels * kMfccFrames; i += blockDim.x) {
    const int m = i / kMfccFrames, f = i - m * kMfccFrames;
    double db = 0.0;
    if (f < nf) {
      const double v = fmax((double)src[(long long)m * frames + f], amin);
      db = fmax(log(v) * scale - offset, floor_db);
    }
    sDb[i] = db;
  }
  __syncthreads();
  const int f = threadIdx.x & (kMfccFrames - 1);
  T* dst = out + b * n_mfcc * frames + p0 + f;
  for (int c = threadIdx.x / kMfccFrames; c < n_mfcc; c += kMfccThreads / kMfccFrames) {
    const double* row = dct + (long long)c * n_mels;
    double acc = 0.0;
    for (int m = 0; m < n_mels; ++m) acc = fma(__ldg(row + m), sDb[m * kMfccFrames + f], acc);
    if (f < nf) dst[(long long)c * frames] = (T)acc;
  }
}

int grid_for(long long count) {
  long long g = (count + 255) / 256;
  return (int)(g < 148 * 8 ? (g < 1 ? 1 : g) : 148 * 8);
}

}  // namespace

cudaError_t launch_to_db(const void* x, long long count, int dtype, int magnitude, double amin,
                         double scale, double offset, bool clamp, double range,
                         unsigned long long* max_slot, void* out, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  if (clamp) {
    cudaError_t e = cudaMemsetAsync(max_slot, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
  }
  const int grid = grid_for(count);
  if (dtype == 0)
    to_db_kernel<float><<<grid, 256, 0, st>>>((const float*)x, count, magnitude, (float)amin,
                                              (float)scale, (float)offset,
                                              clamp ? max_slot : nullptr, (float*)out);
  else
    to_db_kernel<double><<<grid, 256, 0, st>>>((const double*)x, count, magnitude, amin, scale,
                                               offset, clamp ? max_slot : nullptr, (double*)out);
  ++g_launch_count;
  if (clamp) {
    if (dtype == 0) clamp_db_kernel<float><<<grid, 256, 0, st>>>((float*)out, count, max_slot, range);
    else clamp_db_kernel<double><<<grid, 256, 0, st>>>((double*)out, count, max_slot, range);
    ++g_launch_count;
  }
  return cudaGetLastError();
}

cudaError_t launch_mfcc(const void* mel, int dtype, long long batch, int n_mels, long long frames,
                        int n_mfcc, const double* dct, unsigned long long* max_slot, double amin,
                        double scale, double offset, double range, void* out, cudaStream_t st) {
  if (batch == 0 || frames == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(max_slot, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  const long long count = batch * n_mels * frames;
  if (dtype == 0) max_value_kernel<float><<<grid_for(count), 256, 0, st>>>((const float*)mel, count, max_slot);
  else max_value_kernel<double><<<grid_for(count), 256, 0, st>>>((const double*)mel, count, max_slot);
  ++g_launch_count;
  const size_t smem = (size_t)n_mels * kMfccFrames * sizeof(double);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  const long long tiles = (frames + kMfccFrames - 1) / kMfccFrames;
  for (long long b0 = 0; b0 < batch; b0 += 65535) {
    const long long nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((unsigned)tiles, (unsigned)nb);
    if (dtype == 0) {
      e = cudaFuncSetAttribute(mfcc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      mfcc_kernel<float><<<grid, kMfccThreads, smem, st>>>(
          (const float*)mel + b0 * n_mels * frames, n_mels, frames, n_mfcc, dct, max_slot, amin,
          scale, offset, range, (float*)out + b0 * n_mfcc * frames);
    } else {
      e = cudaFuncSetAttribute(mfcc_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      mfcc_kernel<double><<<grid, kMfccThreads, smem, st>>>(
          (const double*)mel + b0 * n_mels * frames, n_mels, frames, n_mfcc, dct, max_slot, amin,
          scale, offset, range, (double*)out + b0 * n_mfcc * frames);
    }
    ++g_launch_count;
  }
  return cudaGetLastError();
}

}  // namespace smb
