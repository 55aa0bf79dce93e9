// Helpers shared by the depthwise tensor-core kernels (dwconv_tc.cu: channels-last gather producers; dwconv_cm.cu:
// TMA-fed channel-major series).  sm_100a only.
#pragma once
#include "common.cuh"

namespace lasr {

constexpr int DT_CG = 16;        // channels per CTA
constexpr int DT_ROWS = 128;     // MMA M
constexpr int DT_MAX_KS = 112;   // longest Toeplitz factor (K-steps of 16 series frames)

// shared-memory matrix descriptors (Blackwell descriptor version 1)
__device__ __forceinline__ uint64_t umma_desc_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;  // LBO unused: one K-step is exactly the 32-byte swizzle span
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;  // SWIZZLE_32B
  return d;
}
// MN-major operand with the 32-byte swizzle: lbo = byte distance between 16-element groups along M/N, sbo = byte
// distance between groups of 8 along K
__device__ __forceinline__ uint64_t umma_desc_sw32_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}

__device__ __forceinline__ void umma_bf16_first(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x2(uint32_t taddr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ldg_v8(const void* p, uint32_t (&a)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t (&a)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]),
               "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// One 16-byte chunk of a channel's Toeplitz factor B[t, s] = w[s - t - delta] (t < 16 output frames of a row, s < KS series
// positions) in the K-major no-swizzle core-matrix order [s / 8][t / 8], element (t % 8, s % 8): chunk `rem` of the
// channel = (core = rem >> 3, row r = rem & 7); taps given un-flipped, `flip` reads them reversed (the data gradient).
__device__ __forceinline__ uint4 toeplitz_chunk(const float* __restrict__ wc, int K, int delta, int flip, int rem) {
  const int core = rem >> 3, r = rem & 7;
  const int j0 = 8 * (core >> 1) - (8 * (core & 1) + r) - delta;
  uint32_t o[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int j = j0 + 2 * m;
    const float a = (j >= 0 && j < K) ? wc[flip ? K - 1 - j : j] : 0.f;
    const float b = (j + 1 >= 0 && j + 1 < K) ? wc[flip ? K - 2 - j : j + 1] : 0.f;
    o[m] = f32x2_to_bf16x2(a, b);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
// plain (non-tensor) bulk copy global -> shared, completion counted on an mbarrier; size % 16 == 0
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
int lasr_cm_ks_host(int K);

// debug timeline (tools/trace_dw.py): [CTA][8 slots][16 stamps] of %globaltimer, normally NULL
unsigned long long* dw_trace_buffer();
__device__ __forceinline__ unsigned long long dw_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define DW_TRACE(buf, slot, idx)                                                                  \
  do {                                                                                            \
    if ((buf) != nullptr && (idx) < 16) (buf)[(blockIdx.x * 8 + (slot)) * 16 + (idx)] = dw_gtimer(); \
  } while (0)

}  // namespace lasr
