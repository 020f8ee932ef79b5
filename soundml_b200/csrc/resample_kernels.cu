// Polyphase resampler / FIR kernels (sm_100a).
//
// One stage of the reference's resampler is, whichever executor the planner
// tags it with, the same filter (resample.ml:22-52): output i of a stage with
// factors l/m and group delay k is
//
//     out[i] = sum_{s=0}^{2k} xz[floor(i*m/l) - k + s] * bank[(i*m) mod l][s]
//
// with xz = x extended by zeros on both sides (resample_stubs.c:127-143;
// independent evaluator soundml/test/resample/resample_kernel.ml:144-163).
// Phase and window start are exact integers -- no float time accumulator.
#include "kernels.h"

namespace smb {

namespace {

constexpr int kOutPerCta = 256;

// Direct form: a CTA produces kOutPerCta consecutive outputs of one signal from
// an input span staged once in shared memory; each thread owns one output and
// walks its phase's bank row.
template <typename T>
__global__ void __launch_bounds__(kOutPerCta)
polyphase_direct_kernel(const T* __restrict__ x, long long n,
                        const T* __restrict__ bank, int l, int m, int k,
                        long long n_out, T* __restrict__ out, int span_cap) {
  extern __shared__ __align__(8) unsigned char sRaw[];
  T* sIn = reinterpret_cast<T*>(sRaw);
  const long long c = blockIdx.y;
  const long long i0 = (long long)blockIdx.x * kOutPerCta;
  const int cnt = (int)min((long long)kOutPerCta, n_out - i0);
  const T* xs = x + c * n;
  const int taps = 2 * k + 1;
  const long long lo = (i0 * m) / l - k;                    // first input read
  const long long hi = ((i0 + cnt - 1) * m) / l + k;        // last input read
  const int span = (int)(hi - lo + 1);
  for (int j = threadIdx.x; j < span; j += blockDim.x) {
    const long long s = lo + j;
    sIn[j] = (s >= 0 && s < n) ? __ldg(xs + s) : T(0);
  }
  __syncthreads();
  if ((int)threadIdx.x < cnt) {
    const long long i = i0 + threadIdx.x;
    const long long t = i * m;
    const int phase = (int)(t % l);
    const int off = (int)(t / l - k - lo);
    const T* h = bank + (long long)phase * taps;
    const T* v = sIn + off;
    T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int s = 0;
    for (; s + 4 <= taps; s += 4) {
      a0 = fma(v[s], __ldg(h + s), a0);
      a1 = fma(v[s + 1], __ldg(h + s + 1), a1);
      a2 = fma(v[s + 2], __ldg(h + s + 2), a2);
      a3 = fma(v[s + 3], __ldg(h + s + 3), a3);
    }
    for (; s < taps; ++s) a0 = fma(v[s], __ldg(h + s), a0);
    out[c * n_out + i] = (a0 + a1) + (a2 + a3);
  }
}

}  // namespace

template <typename T>
static cudaError_t launch_direct_t(const T* x, long long batch, long long n, const T* bank, int l,
                                   int m, int k, long long n_out, T* out, cudaStream_t st) {
  if (batch == 0 || n_out == 0) return cudaSuccess;
  const long long span_cap = ((long long)(kOutPerCta - 1) * m) / l + 2 + 2LL * k + 1;
  const size_t smem = (size_t)span_cap * sizeof(T);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(polyphase_direct_kernel<T>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long tiles = (n_out + kOutPerCta - 1) / kOutPerCta;
  for (long long b0 = 0; b0 < batch; b0 += 65535) {
    const long long nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((unsigned)tiles, (unsigned)nb);
    polyphase_direct_kernel<T><<<grid, kOutPerCta, smem, st>>>(x + b0 * n, n, bank, l, m, k, n_out,
                                                               out + b0 * n_out, (int)span_cap);
    ++g_launch_count;
  }
  return cudaGetLastError();
}

cudaError_t launch_polyphase_direct(const float* x, long long batch, long long n,
                                    const float* bank, int l, int m, int k,
                                    long long n_out, float* out, cudaStream_t st) {
  return launch_direct_t<float>(x, batch, n, bank, l, m, k, n_out, out, st);
}

cudaError_t launch_polyphase_direct_f64(const double* x, long long batch, long long n,
                                        const double* bank, int l, int m, int k,
                                        long long n_out, double* out, cudaStream_t st) {
  return launch_direct_t<double>(x, batch, n, bank, l, m, k, n_out, out, st);
}

}  // namespace smb
