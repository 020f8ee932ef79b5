// Least-squares STFT synthesis (Stft.invert, stft.ml:693-939) on the GPU.
//
//   x[m] = sum_p w[q - p hop] y_p[q - p hop] / sum_p w^2[q - p hop],   q = m + left
//
// with y_p the inverse real transform of frame p.  The reference materialises
// every windowed frame and overlap-adds ceil(fft/hop) shifted planes; here a CTA
// owns a run of output positions and gathers: it inverts, in double, the frames
// that reach its run (frame index descending -- the order in which the
// reference's block planes arrive at one position), adds their windowed taps
// into a shared-memory accumulator, divides by the envelope and writes the run
// once.  No intermediate tensor, no atomics, deterministic.
//
// The envelope follows stft.ml:846-894: positions every residue class reaches
// completely take the folded squared window (j ascending), the partially
// covered borders are summed tap by tap (p ascending); an exact zero is
// replaced by one (stft.ml:840-844).
#include "kernels.h"

namespace smb {

namespace {

template <typename C> struct ComplexIn;
template <> struct ComplexIn<float2> {
  __device__ static double2 load(const float2* p) { const float2 v = *p; return make_double2(v.x, v.y); }
};
template <> struct ComplexIn<double2> {
  __device__ static double2 load(const double2* p) { return *p; }
};

__device__ __forceinline__ long long ceil_div_ll(long long a, long long b) {   // b > 0
  return a >= 0 ? (a + b - 1) / b : -((-a) / b);
}

template <typename CIN, typename TOUT>
__global__ void istft_kernel(const IstftArgs a, int log2n, int seg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.fft, bins = N / 2 + 1;
  double2* buf = reinterpret_cast<double2*>(smem_raw);        // N (pow2) or bins (direct) entries
  double* acc = reinterpret_cast<double*>(buf + (log2n >= 0 ? N : bins));
  const long long b = blockIdx.y;
  const long long span = (a.count - 1) * a.hop + N;
  const long long m0 = (long long)blockIdx.x * seg;
  const long long m1 = min(m0 + seg, a.out_len);
  const long long q0 = m0 + a.left;
  const long long q1 = min(m1 + a.left, span);               // positions [q0, q1) receive taps
  const CIN* z = reinterpret_cast<const CIN*>(a.z) + b * bins * a.frames;
  TOUT* out = reinterpret_cast<TOUT*>(a.out) + b * a.out_len;
  const double inv_n = 1.0 / (double)N;

  for (int i = threadIdx.x; i < seg; i += blockDim.x) acc[i] = 0.0;
  if (q1 > q0) {
    const long long p_hi = min(a.count - 1, (q1 - 1) / a.hop);
    const long long p_lo = max(0LL, ceil_div_ll(q0 - N + 1, a.hop));
    for (long long p = p_hi; p >= p_lo; --p) {
      __syncthreads();
      if (log2n >= 0) {
        // conj of the Hermitian-extended spectrum, bit-reversed: a forward
        // transform of it is the conjugate of the inverse, whose real part we keep
        for (int k = threadIdx.x; k < bins; k += blockDim.x) {
          double2 v = ComplexIn<CIN>::load(z + (long long)k * a.frames + p);
          if (k == 0 || 2 * k == N) v.y = 0.0;                // irfft ignores them
          const int d0 = log2n > 0 ? (int)(__brev((unsigned)k) >> (32 - log2n)) : 0;
          buf[d0] = make_double2(v.x, -v.y);
          if (k > 0 && 2 * k < N) {
            const int d1 = (int)(__brev((unsigned)(N - k)) >> (32 - log2n));
            buf[d1] = make_double2(v.x, v.y);
          }
        }
        __syncthreads();
        for (int s = 1; s <= log2n; ++s) {
          const int half = 1 << (s - 1);
          const int stride = N >> s;
          for (int idx = threadIdx.x; idx < N / 2; idx += blockDim.x) {
            const int j = idx & (half - 1);
            const int base = (idx >> (s - 1)) << s;
            const double2 w = a.twiddle[j * stride];
            const double2 u = buf[base + j];
            const double2 c = buf[base + j + half];
            const double tr = w.x * c.x - w.y * c.y;
            const double ti = w.x * c.y + w.y * c.x;
            buf[base + j] = make_double2(u.x + tr, u.y + ti);
            buf[base + j + half] = make_double2(u.x - tr, u.y - ti);
          }
          __syncthreads();
        }
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
          const long long q = p * a.hop + j;
          if (q >= q0 && q < q1) acc[q - q0] += (buf[j].x * inv_n) * a.window[j];
        }
      } else {
        for (int k = threadIdx.x; k < bins; k += blockDim.x) {
          double2 v = ComplexIn<CIN>::load(z + (long long)k * a.frames + p);
          if (k == 0 || 2 * k == N) v.y = 0.0;
          buf[k] = v;
        }
        __syncthreads();
        // direct inverse: y[j] = (X0 + (-1)^j X_{N/2} + 2 sum_k Re(X_k e^{+2 pi i jk/N})) / N
        const int kmax = (N - 1) / 2;
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
          const long long q = p * a.hop + j;
          if (q < q0 || q >= q1) continue;
          double sum = 0.0;
          int idx = 0;
          for (int k = 1; k <= kmax; ++k) {
            idx += j;
            if (idx >= N) idx -= N;
            const double2 w = a.twiddle[idx];                 // (cos, -sin)
            sum += buf[k].x * w.x + buf[k].y * w.y;
          }
          double y = buf[0].x + 2.0 * sum;
          if ((N & 1) == 0) y += (j & 1) ? -buf[N / 2].x : buf[N / 2].x;
          acc[q - q0] += (y * inv_n) * a.window[j];
        }
      }
    }
  }
  __syncthreads();
  // ---- envelope division, trim, zero extension
  const long long head = min(span, (long long)(N - a.hop));
  const long long stop = max(head, min(span, a.count * (long long)a.hop));
  for (long long m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
    const long long q = m + a.left;
    double v = 0.0;
    if (q < span) {
      double e;
      if (q >= head && q < stop) {
        e = a.folded[q % a.hop];
      } else {
        const long long first = max(0LL, ceil_div_ll(q - N + 1, a.hop));
        const long long last = min(a.count - 1, q / a.hop);
        e = 0.0;
        for (long long p = first; p <= last; ++p) {
          const double w = a.window[q - p * a.hop];
          e += w * w;
        }
      }
      if (e == 0.0) e = 1.0;
      v = acc[m - m0] / e;
    }
    out[m] = (TOUT)v;
  }
}

int ilog2_exact(int n) {
  if (n < 1 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}

}  // namespace

cudaError_t launch_istft(const IstftArgs& a, long long batch, cudaStream_t st) {
  if (batch == 0 || a.out_len == 0) return cudaSuccess;
  const int N = a.fft, bins = N / 2 + 1;
  const int log2n = ilog2_exact(N);
  const size_t fixed = (size_t)(log2n >= 0 ? N : bins) * sizeof(double2);
  const size_t budget = 200 * 1024;
  if (fixed + 64 * sizeof(double) > budget) return cudaErrorInvalidConfiguration;
  long long seg = 4LL * N;
  if (seg > 8192) seg = 8192;
  if (seg < N) seg = N;
  if ((size_t)seg * sizeof(double) > budget - fixed) seg = (long long)((budget - fixed) / sizeof(double));
  if (seg > a.out_len) seg = a.out_len;
  const size_t smem = fixed + (size_t)seg * sizeof(double);
  const long long segs = (a.out_len + seg - 1) / seg;
  if (segs > 2147483647LL) return cudaErrorInvalidConfiguration;
  const int threads = N >= 512 ? 256 : (N >= 128 ? 128 : 64);
  for (long long b0 = 0; b0 < batch; b0 += 65535) {
    const long long nb = batch - b0 < 65535 ? batch - b0 : 65535;
    IstftArgs s = a;
    s.z = (const char*)a.z + (size_t)b0 * bins * a.frames * (a.in_f64 ? 16 : 8);
    s.out = (char*)a.out + (size_t)b0 * a.out_len * (a.out_f64 ? 8 : 4);
    dim3 grid((unsigned)segs, (unsigned)nb);
#define SMB_LAUNCH_ISTFT(CIN, TOUT)                                                        \
  do {                                                                                     \
    cudaError_t e = cudaFuncSetAttribute(istft_kernel<CIN, TOUT>,                          \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                         (int)smem);                                       \
    if (e != cudaSuccess) return e;                                                        \
    istft_kernel<CIN, TOUT><<<grid, threads, smem, st>>>(s, log2n, (int)seg);              \
  } while (0)
    if (a.in_f64) {
      if (a.out_f64) SMB_LAUNCH_ISTFT(double2, double);
      else SMB_LAUNCH_ISTFT(double2, float);
    } else {
      if (a.out_f64) SMB_LAUNCH_ISTFT(float2, double);
      else SMB_LAUNCH_ISTFT(float2, float);
    }
#undef SMB_LAUNCH_ISTFT
    ++g_launch_count;
  }
  return cudaGetLastError();
}

}  // namespace smb

// ---- Griffin-Lim projections (stft.ml:941-1025) --------------------------------
namespace smb {

namespace {

// spec = magnitudes * angles with angles = unit(rebuilt - beta * previous)
// (unit: divide by |.| + the smallest positive normal double, stft.ml:959-962);
// first = 1 builds the initial spectrum from the starting phase instead (unit
// phase when phase == nullptr).
template <typename T>
__global__ void gl_project_kernel(const T* __restrict__ mags, const T* __restrict__ phase,
                                  const double2* __restrict__ rebuilt,
                                  const double2* __restrict__ previous, double beta, int first,
                                  long long count, double2* __restrict__ spec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const double m = (double)mags[i];
  double2 a;
  if (first) {
    if (phase) {
      const double p = (double)phase[i];
      a = make_double2(cos(p), sin(p));
    } else {
      a = make_double2(1.0, 0.0);
    }
  } else {
    double2 e = rebuilt[i];
    if (previous) {
      const double2 q = previous[i];
      e.x -= q.x * beta;
      e.y -= q.y * beta;
    }
    const double d = hypot(e.x, e.y) + 2.2250738585072014e-308;
    a = make_double2(e.x / d, e.y / d);
  }
  spec[i] = make_double2(m * a.x, m * a.y);
}

}  // namespace

cudaError_t launch_gl_project(const void* mags, const void* phase, int dtype,
                              const double2* rebuilt, const double2* previous, double beta,
                              int first, long long count, double2* spec, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  const int threads = 256;
  const long long blocks = (count + threads - 1) / threads;
  if (blocks > 2147483647LL) return cudaErrorInvalidConfiguration;
  if (dtype == 0)
    gl_project_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(
        (const float*)mags, (const float*)phase, rebuilt, previous, beta, first, count, spec);
  else
    gl_project_kernel<double><<<(unsigned)blocks, threads, 0, st>>>(
        (const double*)mags, (const double*)phase, rebuilt, previous, beta, first, count, spec);
  ++g_launch_count;
  return cudaGetLastError();
}

}  // namespace smb
