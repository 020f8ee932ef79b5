// Round-trip time of the small tcgen05.mma batches the tensor-core STFT kernel issues
// (issue -> tcgen05.commit -> mbarrier wait), one CTA per SM:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bench_umma_batch.bin tools/bench_umma_batch.cu
// Variants: MMAs per batch, N per MMA, and whether the other warps of the CTA keep the
// shared-memory pipe busy meanwhile.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: 12 MMAs N = 64 (the kernel's batch); 1: 4 x (N = 128) + 4 x (N = 64); 2: 4 MMAs N = 64;
// 3: 1 MMA N = 64; 4: 12 MMAs N = 64 into three accumulators (no dependent chain)
template <int MODE, bool BUSY>
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* smem = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  const uint32_t a_hi = base, a_lo = base + 16384, f_hi = base + 32768, f_lo = base + 32768 + 8192;
  constexpr uint32_t id64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  constexpr uint32_t id128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    long long total = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      const long long t0 = clock64();
      if (MODE == 0 || MODE == 4) {
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t off = ks * 32;
          umma_f16(tmem, umma_desc(a_hi + off), umma_desc(f_hi + off), id64, ks > 0);
          umma_f16(tmem + (MODE == 4 ? 64 : 0), umma_desc(a_lo + off), umma_desc(f_hi + off), id64, MODE == 4 ? ks > 0 : 1);
          umma_f16(tmem + (MODE == 4 ? 128 : 0), umma_desc(a_hi + off), umma_desc(f_lo + off), id64, MODE == 4 ? ks > 0 : 1);
        }
      } else if (MODE == 1) {
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t off = ks * 32;
          umma_f16(tmem, umma_desc(a_hi + off), umma_desc(f_hi + off), id128, ks > 0);   // F_hi | F_lo contiguous
          umma_f16(tmem, umma_desc(a_lo + off), umma_desc(f_hi + off), id64, 1);
        }
      } else if (MODE == 2) {
        for (int ks = 0; ks < 4; ++ks) umma_f16(tmem, umma_desc(a_hi + ks * 32), umma_desc(f_hi + ks * 32), id64, ks > 0);
      } else {
        umma_f16(tmem, umma_desc(a_hi), umma_desc(f_hi), id64, 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(smem_u32(&bar), phase & 1);
      ++phase;
      total += clock64() - t0;
    }
    out[blockIdx.x] = total;
    reinterpret_cast<volatile uint32_t*>(smem)[49152 / 4] = 1;     // stop flag
  } else if (BUSY && threadIdx.x >= 128) {
    // 12 warps streaming 16-byte shared loads and stores until thread 0 is done
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem) + 49152 / 4;
    uint4* buf = reinterpret_cast<uint4*>(smem + 50176);
    uint4 acc = make_uint4(0, 0, 0, 0);
    while (*flag == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 v = buf[(threadIdx.x + 64 * j) & 511];
        acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
      }
      buf[512 + (threadIdx.x & 255)] = acc;
    }
    if (acc.x == 0x12345678u) out[1000] = acc.y;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE, bool BUSY>
static void run(const char* name, long long* d) {
  const int iters = 2000, smem = 80 * 1024;
  cudaFuncSetAttribute(k<MODE, BUSY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE, BUSY><<<148, 512, smem>>>(d, iters);
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  printf("%-46s %s %8.1f cycles per batch%s\n", name, BUSY ? "(busy smem)" : "(idle smem)", s / 148 / iters,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 2048 * sizeof(long long));
  run<0, false>("12 x (128x64x16) one accumulator", d);
  run<4, false>("12 x (128x64x16) three accumulators", d);
  run<1, false>("4 x (128x128x16) + 4 x (128x64x16)", d);
  run<2, false>("4 x (128x64x16)", d);
  run<3, false>("1 x (128x64x16)", d);
  run<0, true>("12 x (128x64x16) one accumulator", d);
  run<1, true>("4 x (128x128x16) + 4 x (128x64x16)", d);
  run<3, true>("1 x (128x64x16)", d);
  return 0;
}
