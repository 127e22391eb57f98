// 5th-generation tensor core plumbing for sm_100a: tensor memory (TMEM) allocation, shared-memory
// matrix descriptors, the single-thread tcgen05.mma issue, commit-to-mbarrier and the TMEM -> register
// load of the epilogue, tensor-map stores.  Only what ppk_policy.cuh needs (dense fp16 x fp16 -> fp32, one
// CTA per MMA, K-major operands without swizzle).
#pragma once
#include <cstdint>

#include "ppk_async.cuh"

namespace ppk {
namespace tc {

// ---- tensor memory ---------------------------------------------------------------------------------
// One warp allocates `cols` (power of two >= 32) columns and publishes the base address in smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core operand reads, bulk stores)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, no swizzle: the tile is a grid of 8-row x 16-byte
// core matrices (128 contiguous bytes each).  `lbo` = byte distance between the two core matrices that
// are adjacent along K inside one MMA (K = 16 halves = 2 x 16 B), `sbo` = byte distance between
// 8-row groups.  Bits: [0,14) start>>4, [16,30) lbo>>4, [32,46) sbo>>4, [46,48) version = 1 (sm_100),
// [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::f16: D fp32 (bits [4,6) = 1), A and B fp16 (format 0), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc_f16_f32(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// The mbarrier gets one arrival when every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- epilogue --------------------------------------------------------------------------------------
// 32 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier helpers beyond ppk_async.cuh -----------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

// 2-D tensor store shared -> global through a tensor map (TMA); rows / columns outside the tensor are clipped
__device__ __forceinline__ void tensor_store_2d(const void* tmap, int c0, int c1, const void* ssrc) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
               "r"(smem_u32(ssrc))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the committed bulk stores have finished READING shared memory (the buffer may be rewritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tc
}  // namespace ppk
