// Order-preserving map of doubles onto unsigned 64-bit keys, so that a whole-tensor
// maximum is one atomicMax per warp (Convert.power_to_db's top_db clamp, convert.ml:49-52,
// "a whole-tensor reduction").  A zeroed slot is below every key.
#pragma once
#include <cuda_runtime.h>

namespace smb {

__device__ __forceinline__ unsigned long long key_of(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}
// the warp's maximum to *slot (all 32 lanes call)
__device__ __forceinline__ void warp_max_to(double v, unsigned long long* slot) {
  unsigned long long k = key_of(v);
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(slot, k);
}

}  // namespace smb
