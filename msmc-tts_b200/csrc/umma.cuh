// tcgen05 / mbarrier / shared-memory-descriptor helpers shared by the tensor-core kernels (conv_umma.cu, vq_umma.cu).
#pragma once
#include "common.cuh"

namespace msmc {
namespace {

constexpr int UM_BM = 128;
constexpr int UM_BK = 32;       // tf32 per stage row (128 B)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  // watchdog: a protocol bug must surface as a launch failure (trap), never as a hung GPU.  try_wait suspends the
  // thread for a bounded, implementation-defined time per probe, so 2^24 failed probes is seconds, not a step.
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && ++spins == (1u << 24)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// producer-side "stage is full" signal: every thread publishes its shared-memory writes to the async proxy, the
// warp converges, and ONE lane arrives (256 per-thread arrivals on one mbarrier serialise for ~1k cycles)
__device__ __forceinline__ void publish_and_arrive_warp(uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 = 1 | [32,46) SBO>>4 = 64 (8 rows x 128 B) | [46,48) version = 1 | [61,64) layout = 2
// The MMA-issuing thread is instruction-bound at N <= 64 (a 128 x 64 x 8 TF32 MMA occupies the tensor pipe for only
// 32 cycles), so descriptors are built from a per-stage low word plus compile-time offsets: the high word
// (SBO, version, layout) is an immediate, the low word (start >> 4 | LBO << 16) advances by plain 32-bit adds.
constexpr uint32_t DESC_HI_K = 64u | (1u << 14) | (2u << 29);             // K-major SWIZZLE_128B, SBO = 1024 B
constexpr uint32_t DESC_HI_MN = (512u >> 4) | (1u << 14) | (1u << 29);    // MN-major SWIZZLE_128B_BASE32B, SBO = 512 B
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFF) >> 4) | ((4096u >> 4) << 16);
}
template <uint32_t HI>
__device__ __forceinline__ void umma_tf32_lo(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "n"(HI)
      : "memory");
}
// Same instruction issued from CONVERGENT code: every lane of the warp executes the asm with identical operands and
// `elected` is non-zero in exactly one lane (elect_one).  Inside `if (lane == 0) { ... }` ptxas guards every
// tcgen05.mma with its own ELECT / BRA.U.ANY loop (~90 cycles of issue latency per MMA, which -- not the tensor pipe --
// paced every N <= 128 kernel of round 1); predicated straight-line code issues them back to back.
template <uint32_t HI>
__device__ __forceinline__ void umma_tf32_pred(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                               uint32_t accumulate, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "n"(HI), "r"(elected)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit_pred(uint64_t* bar, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// MN-major TF32 operands only exist in the SWIZZLE_128B_BASE32B layout (layout type 1, cute
// Layout_MN_SW128_32B_Atom: 128-byte rows of 32 MN elements, 4 K-rows per atom, 32-byte chunks XORed with
// row & 3).  LBO = 4096 B between 32-element MN blocks, SBO = 512 B between 4-row K groups (desc_lo_mn / DESC_HI_MN).
// byte offset of 16-byte chunk `c16` of K-row `p` inside one 32-channel block of an MN-major operand tile
__device__ __forceinline__ uint32_t mn_off(int p, int c16) {
  return (uint32_t)p * 128u + (uint32_t)(((((c16 >> 1) ^ (p & 3)) << 1) | (c16 & 1)) << 4);
}

}  // namespace
}  // namespace msmc
