// 32-point forward complex FFT of TWO independent sequences at once, held entirely
// in registers (one thread): every value is a float32x2 register pair carrying the
// same quantity of sequence A (low half) and sequence B (high half), so that every
// real operation of the transform is one packed instruction (FADD2 / FMUL2 / FFMA2
// on sm_100a) and every twiddle constant, being the same for both halves, is an
// immediate or a broadcast scalar operand.  The fused STFT kernel (stft2048p.cu)
// runs two overlapping frames of one clip through it side by side.
//
// Same decomposition as fft32.cuh: n = 8a + b, k = c + 4d -> eight 4-point DFTs
// over a, twiddle W32^(bc), four 8-point DFTs over b; natural order in and out.
// Multiplications by a constant are folded into the additions that follow them
// wherever one exists (a + c*b is one FFMA2).
#pragma once
#include <cuda_runtime.h>

#include "fft32.cuh"   // kW32C / kW32S

namespace smb {
namespace fft32x2 {

typedef unsigned long long pk_t;   // (A, B): two float32 in one 64-bit register pair

__device__ __forceinline__ pk_t pk(float a, float b) {
  pk_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ pk_t pk1(float s) { return pk(s, s); }
__device__ __forceinline__ float pk_lo(pk_t v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float pk_hi(pk_t v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ pk_t padd(pk_t a, pk_t b) {
  pk_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk_t psub(pk_t a, pk_t b) {
  pk_t r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk_t pmul(pk_t a, pk_t b) {
  pk_t r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk_t pfma(pk_t a, pk_t b, pk_t c) {
  pk_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// the scalar is the same for both halves: ptxas folds the duplicate into a broadcast
// (or immediate) operand
__device__ __forceinline__ pk_t pmuls(pk_t a, float s) { return pmul(a, pk1(s)); }
__device__ __forceinline__ pk_t pfmas(pk_t a, float s, pk_t c) { return pfma(a, pk1(s), c); }

struct CP { pk_t re, im; };   // one complex value of each of the two sequences

__device__ __forceinline__ CP cadd(CP a, CP b) { return CP{padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ CP csub(CP a, CP b) { return CP{psub(a.re, b.re), psub(a.im, b.im)}; }
// a + (-i) b and a - (-i) b:  -i (x + i y) = y - i x
__device__ __forceinline__ CP cadd_mi(CP a, CP b) { return CP{padd(a.re, b.im), psub(a.im, b.re)}; }
__device__ __forceinline__ CP csub_mi(CP a, CP b) { return CP{psub(a.re, b.im), padd(a.im, b.re)}; }

// v * W32^E, E a compile-time exponent that is not a multiple of 8: four packed
// operations (the constants of the whole circle are tabulated, so W32^18 = -W32^2
// costs what W32^2 does)
template <int E>
__device__ __forceinline__ CP rot(CP v) {
  static_assert(E > 0 && E < 32 && E % 8 != 0, "non-trivial rotation");
  if constexpr (E == 4) {
    const float r = 0.7071067811865476f;
    return CP{pmuls(padd(v.re, v.im), r), pmuls(psub(v.im, v.re), r)};
  } else if constexpr (E == 12) {
    const float r = 0.7071067811865476f;
    return CP{pmuls(psub(v.im, v.re), r), pmuls(padd(v.re, v.im), -r)};
  } else {
    constexpr float c = fft32impl::kW32C[E], s = fft32impl::kW32S[E];
    return CP{pfmas(v.im, -s, pmuls(v.re, c)), pfmas(v.im, c, pmuls(v.re, s))};
  }
}

// column B of the 4 x 8 decomposition: 4-point DFT over a of x[8a + B], outputs
// twiddled by W32^(B c) into y[4B + c]
template <int B>
__device__ __forceinline__ void column(const CP (&x)[32], CP (&y)[32]) {
  const CP t0 = cadd(x[B], x[16 + B]), t1 = csub(x[B], x[16 + B]);
  const CP t2 = cadd(x[8 + B], x[24 + B]), d = csub(x[8 + B], x[24 + B]);
  y[4 * B + 0] = cadd(t0, t2);
  const CP r1 = cadd_mi(t1, d), r3 = csub_mi(t1, d);
  if constexpr (B == 0) {
    y[1] = r1;
    y[2] = csub(t0, t2);
    y[3] = r3;
  } else {
    y[4 * B + 1] = rot<B>(r1);
    if constexpr (B == 4) {
      // W32^8 = -i: (t0 - t2) * (-i) written out as swapped subtractions
      y[4 * B + 2] = CP{psub(t0.im, t2.im), psub(t2.re, t0.re)};
    } else {
      y[4 * B + 2] = rot<2 * B>(csub(t0, t2));
    }
    y[4 * B + 3] = rot<3 * B>(r3);
  }
}

// 8-point forward DFT of v[0..7] (already twiddled), natural order in and out.  The
// rotations of the odd half by W8 and W8^3 are folded into the last additions.
__device__ __forceinline__ void fft8(CP (&v)[8]) {
  const CP t0 = cadd(v[0], v[4]), t1 = csub(v[0], v[4]), t2 = cadd(v[2], v[6]), d2 = csub(v[2], v[6]);
  const CP e0 = cadd(t0, t2), e2 = csub(t0, t2), e1 = cadd_mi(t1, d2), e3 = csub_mi(t1, d2);
  const CP u0 = cadd(v[1], v[5]), u1 = csub(v[1], v[5]), u2 = cadd(v[3], v[7]), d3 = csub(v[3], v[7]);
  const CP o0 = cadd(u0, u2), o2 = csub(u0, u2), o1 = cadd_mi(u1, d3), o3 = csub_mi(u1, d3);
  const float r = 0.7071067811865476f;
  // o1 W8 = r ((x + y), (y - x));  o3 W8^3 = r ((y - x), -(x + y))
  const pk_t s1 = padd(o1.re, o1.im), q1 = psub(o1.im, o1.re);
  const pk_t s3 = padd(o3.re, o3.im), q3 = psub(o3.im, o3.re);
  v[0] = cadd(e0, o0);
  v[4] = csub(e0, o0);
  v[1] = CP{pfmas(s1, r, e1.re), pfmas(q1, r, e1.im)};
  v[5] = CP{pfmas(s1, -r, e1.re), pfmas(q1, -r, e1.im)};
  v[2] = cadd_mi(e2, o2);
  v[6] = csub_mi(e2, o2);
  v[3] = CP{pfmas(q3, r, e3.re), pfmas(s3, -r, e3.im)};
  v[7] = CP{pfmas(q3, -r, e3.re), pfmas(s3, r, e3.im)};
}

template <int C>
__device__ __forceinline__ void row(const CP (&y)[32], CP (&x)[32]) {
  CP v[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) v[b] = y[4 * b + C];
  fft8(v);
#pragma unroll
  for (int d = 0; d < 8; ++d) x[C + 4 * d] = v[d];
}

// x[k] <- sum_n x[n] W32^(nk) for both sequences
__device__ __forceinline__ void fft32(CP (&x)[32]) {
  CP y[32];
  column<0>(x, y); column<1>(x, y); column<2>(x, y); column<3>(x, y);
  column<4>(x, y); column<5>(x, y); column<6>(x, y); column<7>(x, y);
  row<0>(y, x); row<1>(y, x); row<2>(y, x); row<3>(y, x);
}

}  // namespace fft32x2
}  // namespace smb
