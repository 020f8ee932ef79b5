// 32-point forward complex FFT held entirely in registers (one thread, natural
// order in and out): n = 8a + b, k = c + 4d -> eight 4-point DFTs over a, twiddle
// W32^(bc) with compile-time constants, four 8-point DFTs over b.  Shared by the
// fused STFT kernel and its micro-benchmark.
#pragma once
#include <cuda_runtime.h>

namespace smb {
namespace fft32impl {

__device__ constexpr float kW32C[32] = {
    1.0f, 0.9807852804032304f, 0.9238795325112867f, 0.8314696123025452f,
    0.7071067811865476f, 0.5555702330196023f, 0.38268343236508984f, 0.19509032201612833f,
    0.0f, -0.1950903220161282f, -0.3826834323650897f, -0.555570233019602f,
    -0.7071067811865475f, -0.8314696123025453f, -0.9238795325112867f, -0.9807852804032304f,
    -1.0f, -0.9807852804032304f, -0.9238795325112868f, -0.8314696123025455f,
    -0.7071067811865477f, -0.5555702330196022f, -0.38268343236509034f, -0.19509032201612866f,
    0.0f, 0.1950903220161283f, 0.38268343236509f, 0.5555702330196018f,
    0.7071067811865474f, 0.8314696123025452f, 0.9238795325112865f, 0.9807852804032303f};
__device__ constexpr float kW32S[32] = {
    0.0f, -0.19509032201612825f, -0.3826834323650898f, -0.5555702330196022f,
    -0.7071067811865475f, -0.8314696123025452f, -0.9238795325112867f, -0.9807852804032304f,
    -1.0f, -0.9807852804032304f, -0.9238795325112867f, -0.8314696123025455f,
    -0.7071067811865476f, -0.5555702330196022f, -0.3826834323650899f, -0.1950903220161286f,
    0.0f, 0.19509032201612836f, 0.38268343236508967f, 0.555570233019602f,
    0.7071067811865475f, 0.8314696123025452f, 0.9238795325112865f, 0.9807852804032303f,
    1.0f, 0.9807852804032304f, 0.9238795325112866f, 0.8314696123025455f,
    0.7071067811865477f, 0.5555702330196022f, 0.3826834323650904f, 0.19509032201612872f};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// v * W32^E with E a compile-time exponent: trivial rotations cost no multiply.
template <int E>
__device__ __forceinline__ float2 rot32(float2 v) {
  if constexpr (E == 0) return v;
  else if constexpr (E == 8) return make_float2(v.y, -v.x);
  else if constexpr (E == 16) return make_float2(-v.x, -v.y);
  else if constexpr (E == 24) return make_float2(-v.y, v.x);
  else if constexpr (E == 4) {
    const float r = 0.7071067811865476f;
    return make_float2((v.x + v.y) * r, (v.y - v.x) * r);
  } else if constexpr (E == 12) {
    const float r = 0.7071067811865476f;
    return make_float2((v.y - v.x) * r, -(v.x + v.y) * r);
  } else {
    constexpr float c = kW32C[E], s = kW32S[E];
    return make_float2(v.x * c - v.y * s, v.x * s + v.y * c);
  }
}

__device__ __forceinline__ void fft4(float2 a0, float2 a1, float2 a2, float2 a3,
                                     float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
  const float2 d = csub(a1, a3);
  const float2 t3 = make_float2(d.y, -d.x);     // -i (a1 - a3)
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = cadd(t1, t3);
  x3 = csub(t1, t3);
}

// 8-point forward DFT of v[0..7], natural order in and out.
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
  float2 e0, e1, e2, e3, o0, o1, o2, o3;
  fft4(v[0], v[2], v[4], v[6], e0, e1, e2, e3);
  fft4(v[1], v[3], v[5], v[7], o0, o1, o2, o3);
  o1 = rot32<4>(o1);
  o2 = rot32<8>(o2);
  o3 = rot32<12>(o3);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <int B>
__device__ __forceinline__ void fft32_column(const float2 (&x)[32], float2 (&y)[32]) {
  float2 r0, r1, r2, r3;
  fft4(x[B], x[8 + B], x[16 + B], x[24 + B], r0, r1, r2, r3);
  y[B * 4 + 0] = r0;
  y[B * 4 + 1] = rot32<(B * 1) % 32>(r1);
  y[B * 4 + 2] = rot32<(B * 2) % 32>(r2);
  y[B * 4 + 3] = rot32<(B * 3) % 32>(r3);
}

template <int C>
__device__ __forceinline__ void fft32_row(const float2 (&y)[32], float2 (&x)[32]) {
  float2 v[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) v[b] = y[b * 4 + C];
  fft8(v);
#pragma unroll
  for (int d = 0; d < 8; ++d) x[C + 4 * d] = v[d];
}

// 32-point forward DFT in registers, natural order in and out:
// n = 8a + b, k = c + 4d  ->  4-point DFTs over a, twiddle W32^(bc), 8-point over b.
__device__ __forceinline__ void fft32(float2 (&x)[32]) {
  float2 y[32];
  fft32_column<0>(x, y); fft32_column<1>(x, y); fft32_column<2>(x, y); fft32_column<3>(x, y);
  fft32_column<4>(x, y); fft32_column<5>(x, y); fft32_column<6>(x, y); fft32_column<7>(x, y);
  fft32_row<0>(y, x); fft32_row<1>(y, x); fft32_row<2>(y, x); fft32_row<3>(y, x);
}

}  // namespace fft32impl
}  // namespace smb
