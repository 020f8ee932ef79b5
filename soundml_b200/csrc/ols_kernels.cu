// Overlap-save FIR / resampler stage on the GPU (sm_100a).
//
// Follows the reference's OLS executor geometry (resample.ml:279-300,
// 1309-1319, 1456-1599; spectrum shaping resample_stubs.c:329-372): with plan
// constants (N, B, K, delta), block b transforms stage inputs
// [b*B - 2K - delta, +N) (zeros outside the signal), multiplies by the plan
// spectrum on the inverse transform's half grid -- periodic extension for a
// xL stage (W = N*L), alias fold for a /M stage (W = N/M), plain product
// otherwise (W = N) -- inverts at length W and keeps the wrap-free run that
// extends the output to hi(b).  The reference's frequency path is complex128;
// here it is float32 (the output tolerance is 1e-5 of peak), with all scalar
// weights (1/M, 1/W) folded into the plan spectrum on the host, in double.
//
// One CTA per (block, signal): the real transforms run as half-length complex
// Stockham FFTs ping-ponging between two shared-memory buffers.
#include "kernels.h"

namespace smb {

namespace {

constexpr int kOlsThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// exp(-2 pi i idx / tn) from the half-circle table (idx < tn): W^(k + tn/2) = -W^k.
__device__ __forceinline__ float2 twiddle(const float2* tw, int idx, int tn) {
  const int half = tn >> 1;
  if (idx < half) return tw[idx];
  const float2 w = tw[idx - half];
  return make_float2(-w.x, -w.y);
}

// Forward complex FFT of `n` points (power of two), Stockham autosort: radix-4
// passes, one radix-2 pass when log2 n is odd.  `tw` holds exp(-2 pi i j / tn)
// for j < tn/2, tn a multiple of n.  Returns the buffer with the natural-order
// result.
__device__ float2* stockham_forward(float2* x, float2* y, int n, const float2* tw, int tn) {
  int ncur = n, s = 1;
  for (; ncur >= 4; ncur >>= 2, s <<= 2) {
    const int m = ncur >> 2;
    const int step = tn / ncur;
    for (int idx = threadIdx.x; idx < (n >> 2); idx += blockDim.x) {
      const int p = idx / s, q = idx - p * s;
      const float2 a = x[q + s * p];
      const float2 b = x[q + s * (p + m)];
      const float2 c = x[q + s * (p + 2 * m)];
      const float2 d = x[q + s * (p + 3 * m)];
      const float2 apc = make_float2(a.x + c.x, a.y + c.y), amc = make_float2(a.x - c.x, a.y - c.y);
      const float2 bpd = make_float2(b.x + d.x, b.y + d.y);
      const float2 jbmd = make_float2(-(b.y - d.y), b.x - d.x);          // i (b - d)
      float2* o = y + q + s * (4 * p);
      o[0] = make_float2(apc.x + bpd.x, apc.y + bpd.y);
      const float2 t1 = make_float2(amc.x - jbmd.x, amc.y - jbmd.y);
      const float2 t2 = make_float2(apc.x - bpd.x, apc.y - bpd.y);
      const float2 t3 = make_float2(amc.x + jbmd.x, amc.y + jbmd.y);
      if (p == 0) {
        o[s] = t1;
        o[2 * s] = t2;
        o[3 * s] = t3;
      } else {
        o[s] = cmul(t1, tw[p * step]);
        o[2 * s] = cmul(t2, twiddle(tw, 2 * p * step, tn));
        o[3 * s] = cmul(t3, twiddle(tw, 3 * p * step, tn));
      }
    }
    __syncthreads();
    float2* t = x;
    x = y;
    y = t;
  }
  if (ncur == 2) {
    for (int q = threadIdx.x; q < s; q += blockDim.x) {
      const float2 a = x[q], b = x[q + s];
      y[q] = make_float2(a.x + b.x, a.y + b.y);
      y[q + s] = make_float2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
    float2* t = x;
    x = y;
    y = t;
  }
  return x;
}

__global__ void __launch_bounds__(kOlsThreads)
ols_kernel(const OlsArgs a) {
  extern __shared__ __align__(16) float2 sm2[];
  const int half_n = a.N >> 1, half_w = a.W >> 1;
  const int cap = half_n > half_w ? half_n : half_w;
  const int tn = a.N > a.W ? a.N : a.W;
  float2* bufA = sm2;                   // two ping-pong buffers of cap + 1 complex values
  float2* bufB = bufA + cap + 1;
  float2* tw = bufB + cap + 1;          // tn/2 twiddles

  const long long b = blockIdx.x;
  const long long c = blockIdx.y;
  const float* xs = a.x + c * a.n;
  float* out = a.out + c * a.n_out;

  for (int j = threadIdx.x; j < (tn >> 1); j += blockDim.x) tw[j] = a.tw[j];
  // block b reads stage inputs [b*B - 2K - delta, +N), zeros outside the signal
  const long long start = b * a.B - 2LL * a.K - a.delta;
  float* in = reinterpret_cast<float*>(bufA);
  for (int j = threadIdx.x; j < a.N; j += blockDim.x) {
    const long long s = start + j;
    in[j] = (s >= 0 && s < a.n) ? __ldg(xs + s) : 0.0f;
  }
  __syncthreads();
  float2* z = stockham_forward(bufA, bufB, half_n, tw, tn);   // z = FFT of x[2n] + i x[2n+1]
  float2* Xs = z == bufA ? bufB : bufA;  // N/2 + 1 bins, in the buffer z does not occupy
  float2* Ys = z;                         // W/2 + 1 bins, over z once X is complete

  // half spectrum X[0 .. N/2] of the real block
  const int sN = tn / a.N;
  for (int k = threadIdx.x; k <= half_n; k += blockDim.x) {
    const float2 zk = z[k == half_n ? 0 : k];
    const float2 zn = cconj(z[(half_n - k) % half_n]);
    const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
    const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y - zn.y));
    // W_N^k, k in [0, N/2]: the table stops at tn/2 - 1, W^(N/2) = -1
    const float2 w = k == half_n ? make_float2(-1.f, 0.f) : tw[k * sN];
    const float2 t = cmul(make_float2(d.y, -d.x), w);          // -i d W
    Xs[k] = make_float2(e.x + t.x, e.y + t.y);
  }
  __syncthreads();

  // plan spectrum on the inverse transform's half grid (resample_stubs.c:329-372)
  for (int k = threadIdx.x; k <= half_w; k += blockDim.x) {
    float2 yk;
    if (a.L > 1) {                       // periodic extension, then * H[k]
      const int j = k % a.N;
      const float2 xf = j <= half_n ? Xs[j] : cconj(Xs[a.N - j]);
      yk = cmul(xf, __ldg(a.H + k));
    } else if (a.M > 1) {                // product on the N grid, alias fold onto W bins
      yk = cmul(Xs[k], __ldg(a.H + k));
      int j = k;
      for (int r = 1; r < a.M; ++r) {
        j += a.W;
        float2 pj;
        if (j <= half_n) pj = cmul(Xs[j], __ldg(a.H + j));
        else pj = cconj(cmul(Xs[a.N - j], __ldg(a.H + (a.N - j))));
        yk.x += pj.x;
        yk.y += pj.y;
      }
    } else {
      yk = cmul(Xs[k], __ldg(a.H + k));
    }
    Ys[k] = yk;
  }
  __syncthreads();

  // inverse real transform of length W through a forward half-length FFT:
  // Zinv[k] = (Y[k] + conj Y[W/2-k]) + i W_W^-k (Y[k] - conj Y[W/2-k]); the FFT
  // of conj(Zinv) is the conjugate of the time signal z[n] = y[2n] + i y[2n+1].
  const int sW = tn / a.W;
  for (int k = threadIdx.x; k < half_w; k += blockDim.x) {
    const float2 yk = Ys[k];
    const float2 yn = cconj(Ys[half_w - k]);
    const float2 e = make_float2(yk.x + yn.x, yk.y + yn.y);
    const float2 d = make_float2(yk.x - yn.x, yk.y - yn.y);
    const float2 w = cconj(tw[k * sW]);                        // W_W^-k, k < W/2
    const float2 o = cmul(d, w);
    const float2 zi = make_float2(e.x - o.y, e.y + o.x);       // e + i o
    Xs[k] = cconj(zi);                                         // X is dead: reuse its buffer
  }
  __syncthreads();
  float2* r = stockham_forward(Xs, Ys, half_w, tw, tn);

  // block b extends the output run to hi(b) (resample.ml:1309-1319)
  long long hi, lo;
  if (a.L > 1) {
    hi = (long long)a.L * (b * a.B + a.N - 3LL * a.K) - 1;
    lo = b == 0 ? 0 : (long long)a.L * ((b - 1) * a.B + a.N - 3LL * a.K);
  } else {
    const long long num = b * a.B + a.N - 3LL * a.K - a.delta - 1;
    hi = num >= 0 ? num / a.M : -1;
    const long long prev = (b - 1) * a.B + a.N - 3LL * a.K - a.delta - 1;
    lo = (b == 0 || prev < 0) ? 0 : prev / a.M + 1;
  }
  if (hi >= a.n_out) hi = a.n_out - 1;
  for (long long i = lo + threadIdx.x; i <= hi; i += blockDim.x) {
    long long pos;
    if (a.L > 1) pos = i + (long long)a.L * (3LL * a.K - b * a.B);
    else pos = (i * a.M + 3LL * a.K + a.delta - b * a.B) / a.M;
    const float2 v = r[pos >> 1];
    out[i] = (pos & 1) ? -v.y : v.x;
  }
}

}  // namespace

size_t ols_smem_bytes(int N, int W) {
  const int cap = (N > W ? N : W) / 2;
  const int tn = N > W ? N : W;
  return (size_t)(2 * (cap + 1) + tn / 2) * sizeof(float2);
}

cudaError_t launch_ols(const OlsArgs& a, long long batch, cudaStream_t st) {
  if (batch == 0 || a.n_out == 0 || a.blocks == 0) return cudaSuccess;
  const size_t smem = ols_smem_bytes(a.N, a.W);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(ols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  for (long long b0 = 0; b0 < batch; b0 += 65535) {
    const long long nb = batch - b0 < 65535 ? batch - b0 : 65535;
    OlsArgs s = a;
    s.x = a.x + b0 * a.n;
    s.out = a.out + b0 * a.n_out;
    dim3 grid((unsigned)a.blocks, (unsigned)nb);
    ols_kernel<<<grid, kOlsThreads, smem, st>>>(s);
    ++g_launch_count;
  }
  return cudaGetLastError();
}

}  // namespace smb
