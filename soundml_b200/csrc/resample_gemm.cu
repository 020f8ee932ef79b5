// Phase-rich polyphase stage as a banded matrix product on the 5th-generation
// tensor cores (tcgen05 + TMEM), sm_100a.
//
// The reference runs such stages (L >= 64 phases, e.g. 44.1 -> 16 kHz: L/M =
// 160/441, 523 taps) as [R; P] x [P; L] through Nx.matmul
// (resample.ml:430-476, 1608-1698, gemm_bank :207-226): block-row rho holds the
// L outputs of one phase cycle and reads the P = 2K + 1 + floor((L-1) M / L)
// consecutive inputs starting at rho*M - K; column r of G is the phase-(r M mod L)
// bank row shifted down by floor(r M / L).  This is a true dense contraction
// (3.8 TFLOP for BASELINE config 4), so it belongs on the tensor pipe:
//
//   D[128 rows x L] (TMEM, fp32)  +=  A[128 x 32] (smem)  *  B[L x 32]^T (smem)
//
// per K-chunk of 32 inputs, kind::tf32.  float32 accuracy (1e-5 of peak) is kept
// by operand splitting: a = a_hi + a_lo, g = g_hi + g_lo with 11-bit pieces and
// D += a_hi g_hi + a_hi g_lo + a_lo g_hi  (the dropped term is 2^-22).  The two
// small terms go to their own accumulator and the main term alternates between
// two accumulators per chunk, which keeps each accumulation chain short.
//
//   A: gathered by four producer warps (a warp takes 32 block-rows, lane = sample
//      within the chunk: one coalesced 128-byte request per row), split, written
//      in the canonical K-major SWIZZLE_128B layout (16-byte chunk index XOR row
//      mod 8); loads run two chunks ahead in registers.
//   B: pre-split, pre-swizzled images of G built once per plan on the host; one
//      cp.async.bulk (TMA) per chunk lands hi+lo with an mbarrier.
//   MMA: a fifth warp's elected lane issues 4 K-steps x 3 products per chunk;
//      tcgen05.commit releases the stage; three stages, mbarriers only.
//   Epilogue: tcgen05.ld (32 lanes x 16 columns per warp), sum of the three
//      accumulators, transposed through shared memory so global stores are
//      row-contiguous (a block-row is L consecutive outputs).
#include "kernels.h"

#include <cstdint>

namespace smb {

namespace {

constexpr int kRows = 128;            // block-rows per CTA (UMMA M)
constexpr int kChunk = 32;            // tf32 elements per 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kABytes = kRows * 128;  // one split of the A chunk
constexpr int kProducerWarps = 8;      // warps 0-3 double as the epilogue
constexpr int kProducers = kProducerWarps * 32;
constexpr int kRowsPerWarp = kRows / kProducerWarps;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kThreads = kProducers + 32;  // producer warps + 1 TMA/MMA warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);         // start address, 16-byte units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Warps 0-3: producers (gather + split + swizzled store of A) and, at the end,
// the epilogue (a warp may only read its own 32 TMEM lanes).  Warp 4: loads the
// B images by TMA and issues the MMAs.  Stages hand over through mbarriers only:
//   a_full[s] (128 arrivals)  b_full[s] (TMA bytes)  ->  MMA  ->  empty[s] (commit)
__global__ void __launch_bounds__(kThreads, 1)
resample_gemm_kernel(const GemmResampleArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int b_bytes = a.n_pad * 128;                      // one split of the B chunk
  const int stage_bytes = 2 * kABytes + 2 * b_bytes;
  __shared__ __align__(8) uint64_t bars[3 * kStages + 1];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const long long c = blockIdx.y;
  const long long row0 = (long long)blockIdx.x * kRows;
  const float* xs = a.x + c * a.n;
  float* out = a.out + c * a.n_out;

  const uint32_t bar_a = smem_u32(&bars[0]);                // [stage]: A tile written
  const uint32_t bar_b = smem_u32(&bars[kStages]);          // [stage]: B bytes landed
  const uint32_t bar_empty = smem_u32(&bars[2 * kStages]);  // [stage]: MMAs reading the stage done
  const uint32_t bar_done = smem_u32(&bars[3 * kStages]);   // all MMAs done
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_a + 8 * s, kProducers);
      mbar_init(bar_b + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_slot)), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_slot;
  const uint32_t acc_main0 = tmem, acc_main1 = tmem + a.n_pad;
  const uint32_t acc_corr = tmem + 2 * a.n_pad;

  if (warp == kMmaWarp) {
    // ===== B loader + MMA issuer (one elected lane) =====
    if (lane == 0) {
      // instruction descriptor: D f32, A/B tf32, both K-major, M = 128; N per chunk
      const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 4) << 24);
      uint32_t used_main0 = 0, used_main1 = 0, used_corr = 0;
      const uint8_t* images = reinterpret_cast<const uint8_t*>(a.b_images);
      // G is banded: chunk ch only touches columns [col0, col0 + ncols) of the
      // accumulators, and only that slice of the image exists (chunk_meta).
      auto load_b = [&](int ch) {
        const int s = ch % kStages;
        const int4 meta = __ldg(a.chunk_meta + ch);
        const uint32_t bytes = 2u * (uint32_t)meta.z * 128u;
        mbar_expect_tx(bar_b + 8 * s, bytes);
        bulk_g2s(smem_u32(smem + s * stage_bytes + 2 * kABytes), images + meta.x, bytes,
                 bar_b + 8 * s);
      };
      for (int ch = 0; ch < kStages && ch < a.chunks; ++ch) load_b(ch);
      for (int ch = 0; ch < a.chunks; ++ch) {
        const int s = ch % kStages;
        const uint32_t parity = (ch / kStages) & 1;
        mbar_wait(bar_a + 8 * s, parity);
        mbar_wait(bar_b + 8 * s, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int4 meta = __ldg(a.chunk_meta + ch);
        const uint32_t idesc = idesc0 | ((uint32_t)(meta.z >> 3) << 17);
        const uint32_t a_hi = smem_u32(smem + s * stage_bytes), a_lo = a_hi + kABytes;
        const uint32_t b_hi = a_hi + 2 * kABytes, b_lo = b_hi + (uint32_t)meta.z * 128u;
        const uint32_t acc = ((ch & 1) ? acc_main1 : acc_main0) + (uint32_t)meta.y;
        const uint32_t acc_c = acc_corr + (uint32_t)meta.y;
        uint32_t& used = (ch & 1) ? used_main1 : used_main0;
#pragma unroll
        for (int ks = 0; ks < kChunk / 8; ++ks) {
          const uint32_t off = ks * 32;                             // 8 tf32 = 32 bytes along K
          umma_tf32(acc, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, used);
          used = 1;
          umma_tf32(acc_c, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, used_corr);
          used_corr = 1;
          umma_tf32(acc_c, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, 1);
        }
        umma_commit(bar_empty + 8 * s);
        if (ch == a.chunks - 1) umma_commit(bar_done);
        // refill this stage's B as soon as its MMAs have drained
        if (ch + kStages < a.chunks) {
          mbar_wait(bar_empty + 8 * s, parity);
          load_b(ch + kStages);
        }
      }
    }
    __syncwarp();
  } else {
    // ===== producers: a warp gathers 32 block-rows, lane = sample within the
    // chunk, so every row is one coalesced 128-byte request.  Loads run two
    // chunks ahead of the stores (registers), stores one to three chunks ahead
    // of the tensor core (stages).
    float v0[kRowsPerWarp], v1[kRowsPerWarp];
    const long long first_row = row0 + warp * kRowsPerWarp;
    // x index of (row rr of this warp, this lane) in chunk 0; chunk ch adds 32 ch
    const long long base0 = first_row * a.m + lane - a.k;
    // the warp's whole footprint over all chunks is interior for most tiles:
    // no per-element bounds checks then
    const bool interior = first_row * a.m - a.k >= 0 &&
                          (first_row + kRowsPerWarp - 1) * a.m - a.k + (long long)a.chunks * kChunk <= a.n;
    auto load_chunk = [&](int ch, float (&v)[kRowsPerWarp]) {
      if (ch >= a.chunks) return;
      const long long b0 = base0 + (long long)ch * kChunk;
      if (interior) {
        const float* p = xs + b0;
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) v[rr] = __ldg(p + rr * a.m);
      } else {
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
          const long long si = b0 + (long long)rr * a.m;
          v[rr] = (si >= 0 && si < a.n) ? __ldg(xs + si) : 0.0f;
        }
      }
    };
    auto store_chunk = [&](int ch, const float (&v)[kRowsPerWarp]) {
      const int s = ch % kStages;
      if (ch >= kStages) mbar_wait(bar_empty + 8 * s, ((ch / kStages) - 1) & 1);
      // element (row, k) lands at 16-byte chunk (k/4) XOR (row mod 8) of its
      // 128-byte row (K-major SWIZZLE_128B)
      float* ahi = reinterpret_cast<float*>(smem + s * stage_bytes) + warp * kRowsPerWarp * 32;
      float* alo = ahi + kABytes / 4;
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const float h = __uint_as_float(__float_as_uint(v[rr]) & 0xFFFFE000u);
        const int cell = rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3));
        ahi[cell] = h;
        alo[cell] = v[rr] - h;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> async proxy
      mbar_arrive(bar_a + 8 * s);
    };
    load_chunk(0, v0);
    load_chunk(1, v1);
    for (int ch = 0; ch < a.chunks; ch += 2) {
      store_chunk(ch, v0);
      load_chunk(ch + 2, v0);
      if (ch + 1 < a.chunks) {
        store_chunk(ch + 1, v1);
        load_chunk(ch + 3, v1);
      }
    }
    if (warp >= 4) goto teardown;             // warps 0-3 own the TMEM lanes

    // ===== epilogue: TMEM -> registers (thread = block-row) -> shared tile ->
    // row-contiguous global stores (a block-row is L consecutive outputs)
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* tile = reinterpret_cast<float*>(smem);            // the stages are drained
    const int pitch = a.n_pad + 4;
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const bool two_main = a.chunks > 1;
    for (int col = 0; col < a.n_pad; col += 16) {
      float d0[16], d1[16], dc[16];
      tmem_ld16(acc_main0 + lane_base + col, d0);
      tmem_ld16(acc_corr + lane_base + col, dc);
      if (two_main) tmem_ld16(acc_main1 + lane_base + col, d1);
      float4* dst = reinterpret_cast<float4*>(tile + tid * pitch + col);
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        float4 o;
        o.x = (two_main ? d0[e] + d1[e] : d0[e]) + dc[e];
        o.y = (two_main ? d0[e + 1] + d1[e + 1] : d0[e + 1]) + dc[e + 1];
        o.z = (two_main ? d0[e + 2] + d1[e + 2] : d0[e + 2]) + dc[e + 2];
        o.w = (two_main ? d0[e + 3] + d1[e + 3] : d0[e + 3]) + dc[e + 3];
        dst[e >> 2] = o;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kRows) : "memory");   // producers/epilogue warps only
    // block-row r is the L_total consecutive outputs of one phase cycle; this
    // launch owns columns [col_begin, col_begin + l) of it
    const long long rows_total = (a.n_out + a.l_total - 1) / a.l_total;
    const int rows_here = (int)min((long long)kRows, rows_total - row0);
    const bool vec = ((a.l | a.l_total | a.col_begin) & 3) == 0 &&
                     (reinterpret_cast<size_t>(out) & 15) == 0;
    if (vec) {
      const int per_row = a.l >> 2;
      for (int j = tid; j < rows_here * per_row; j += kRows) {
        const int row = j / per_row, c4 = j - row * per_row;
        const long long g = (row0 + row) * a.l_total + a.col_begin + 4 * c4;
        const float4 val = *reinterpret_cast<const float4*>(tile + row * pitch + 4 * c4);
        if (g + 3 < a.n_out) {
          *reinterpret_cast<float4*>(out + g) = val;
        } else {                                              // the signal ends inside this float4
          const float e[4] = {val.x, val.y, val.z, val.w};
          for (int t = 0; t < 4 && g + t < a.n_out; ++t) out[g + t] = e[t];
        }
      }
    } else {
      for (int j = tid; j < rows_here * a.l; j += kRows) {
        const int row = j / a.l, cc = j - row * a.l;
        const long long g = (row0 + row) * a.l_total + a.col_begin + cc;
        if (g < a.n_out) out[g] = tile[row * pitch + cc];
      }
    }
  }
teardown:
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"(a.tmem_cols) : "memory");
}

}  // namespace

size_t resample_gemm_smem_bytes(int n_pad) {
  return (size_t)kStages * (2 * kABytes + 2 * (size_t)n_pad * 128) + 1024;
}

cudaError_t launch_resample_gemm(const GemmResampleArgs& a, long long batch, cudaStream_t st) {
  if (batch == 0 || a.n_out == 0) return cudaSuccess;
  const size_t smem = resample_gemm_smem_bytes(a.n_pad);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(resample_gemm_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long rows = (a.n_out + a.l_total - 1) / a.l_total;
  const long long tiles = (rows + kRows - 1) / kRows;
  for (long long b0 = 0; b0 < batch; b0 += 65535) {
    const long long nb = batch - b0 < 65535 ? batch - b0 : 65535;
    GemmResampleArgs s = a;
    s.x = a.x + b0 * a.n;
    s.out = a.out + b0 * a.n_out;
    dim3 grid((unsigned)tiles, (unsigned)nb);
    resample_gemm_kernel<<<grid, kThreads, smem, st>>>(s);
    ++g_launch_count;
  }
  return cudaGetLastError();
}

}  // namespace smb
