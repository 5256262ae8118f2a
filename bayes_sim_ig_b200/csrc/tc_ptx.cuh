// PTX wrappers shared by the tcgen05 kernels (gemm_tc.cu, corr_layer.cu): mbarriers, TMA tile
// loads, UMMA shared-memory descriptors, tcgen05.mma (SS and TS forms), commit, TMEM ld/st.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace bsig {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps inside the instruction until the phase
// completes (or the hint expires) instead of burning issue slots that the producer /
// converter warps on the same scheduler need
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// K-major, 128B-swizzled operand tile: 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);           // start address
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
// MN-major operand tile as four TMA boxes of [32 (MN, contiguous: 128 B) x BK (K rows)].
// For 32-bit operands the ONLY MN-major layout the tensor core accepts is
// SWIZZLE_128B_BASE32B (cute: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over 32 MN x 4 K):
// 32-byte chunks of a 128-byte row are XOR-ed with (row mod 4) -- the TMA counterpart is
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Atoms (4 K-rows = 512 B) follow each other along K
// every 512 B (SBO) and along MN every BK * 128 B (LBO = one box); an MMA (K = 8) consumes
// two K-atoms.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);           // start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16; // leading byte offset: next 32 MN elements
  d |= (uint64_t)(512 >> 4) << 32;                  // stride byte offset: next 4 K rows
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;                           // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row of the tile, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}


// 32 consecutive 32-bit TMEM columns of this warp's lane quadrant <-> one register array
#define BSIG_TMEM_LD32(R, ADDR)                                                                   \
  asm volatile(                                                                                   \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                   \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                   \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"   \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]),       \
        "=r"(R[7]), "=r"(R[8]), "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]),   \
        "=r"(R[14]), "=r"(R[15]), "=r"(R[16]), "=r"(R[17]), "=r"(R[18]), "=r"(R[19]),             \
        "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]), "=r"(R[25]),             \
        "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])              \
      : "r"(ADDR))
#define BSIG_TMEM_ST32(ADDR, R)                                                                  \
  asm volatile(                                                                                  \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                            \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                 \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(  \
          ADDR),                                                                                 \
      "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]), "r"(R[7]),    \
      "r"(R[8]), "r"(R[9]), "r"(R[10]), "r"(R[11]), "r"(R[12]), "r"(R[13]), "r"(R[14]),          \
      "r"(R[15]), "r"(R[16]), "r"(R[17]), "r"(R[18]), "r"(R[19]), "r"(R[20]), "r"(R[21]),        \
      "r"(R[22]), "r"(R[23]), "r"(R[24]), "r"(R[25]), "r"(R[26]), "r"(R[27]), "r"(R[28]),        \
      "r"(R[29]), "r"(R[30]), "r"(R[31])                                                         \
      : "memory")
#define BSIG_TMEM_ST16(ADDR, R)                                                                  \
  asm volatile(                                                                                  \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                            \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(ADDR),    \
      "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]), "r"(R[7]),    \
      "r"(R[8]), "r"(R[9]), "r"(R[10]), "r"(R[11]), "r"(R[12]), "r"(R[13]), "r"(R[14]),          \
      "r"(R[15])                                                                                 \
      : "memory")
#define BSIG_TMEM_LD16(R, ADDR)                                                                   \
  asm volatile(                                                                                   \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                   \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"            \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]),       \
        "=r"(R[7]), "=r"(R[8]), "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]),   \
        "=r"(R[14]), "=r"(R[15])                                                                  \
      : "r"(ADDR))
#define BSIG_TMEM_ST8(ADDR, R)                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"( \
                   ADDR),                                                                        \
               "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]),      \
               "r"(R[7])                                                                         \
               : "memory")

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();   // cuTensorMapEncodeTiled through the runtime's driver entry point

}  // namespace tc
}  // namespace bsig
