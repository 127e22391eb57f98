// Thin wrappers over the sm_90+/sm_100a asynchronous-copy machinery used for staging:
// mbarrier (transaction-count completion) and cp.async.bulk (the 1-D form of TMA; SASS: UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

// make the initialised barrier visible to the async proxy before any bulk copy names it
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// one arrival + announce `bytes` of asynchronous traffic that will complete on this barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    // a transaction-count mismatch would otherwise hang the GPU: fail loudly instead
    if (++spins > (1u << 26)) __trap();
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// for waits that are long by design (a producer far ahead of its consumer): back off between polls so
// the spinning warp does not compete for issue slots with the warps doing the work
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 24)) __trap();
    __nanosleep(200);
  }
}

// one lane of the (fully converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Named barriers 1..15 (0 is __syncthreads): `threads` = arrivers + waiters.  A bar.sync completes only when EVERY expected
// thread has arrived, so a barrier shared by several waiting warps makes each of them wait for the others: give every
// waiting warp its own barrier id.
__device__ __forceinline__ void bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void bar_wait(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// Programmatic dependent launch (griddepcontrol): a kernel launched with the programmatic-stream-serialization attribute
// may START while the kernel before it in the stream is still draining; gdc_wait() blocks until that kernel has
// completed and its memory is visible (a no-op for a plain launch), gdc_launch_dependents() lets the NEXT kernel of the
// stream start early once every CTA of this one has called it (or exited).
__device__ __forceinline__ void gdc_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gdc_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tiled tensor copy global -> shared (SASS: UTMALDG): box of the tensor map at element coordinates (c0, c1);
// dst 128-byte aligned; the inner coordinate c0 must start on a 16-byte boundary (measured: otherwise the
// instruction faults, profiles/r2_staging_probe.md).
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tensor_map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst_smem)),
      "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// shared -> global bulk copy (SASS: UBLKCP.G.S); dst, src 16-byte aligned, bytes a multiple of 16.  The writes to the
// source made through the generic proxy must be fenced (fence_proxy_async) and ordered (barrier) before the issue.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
// ---- the step kernels' streams with an L2 eviction-priority hint (A/B builds: -DPPK_L2_HINT=1 loads, 2 stores, 3 both):
// every input byte is read once and every output byte is written once per step, so evict-first would be the natural
// policy.  Measured (profiles/r2_l2_hint.log): on the loads it COSTS 4-7 % (TILT 65 536 envs 14.7 -> 15.7 us, 1 M envs
// 187 -> 195 us, ADOF 28.3 -> 29.4 us), on the stores it changes nothing: the default stays without a hint
#ifndef PPK_L2_HINT
#define PPK_L2_HINT 0
#endif
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_in(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  if (PPK_L2_HINT & 1)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_evict_first_policy())
                 : "memory");
  else bulk_g2s(dst_smem, src_gmem, bytes, bar);
}
__device__ __forceinline__ void tma_load_2d_in(void* dst_smem, const void* tensor_map, int c0, int c1, uint64_t* bar) {
  if (PPK_L2_HINT & 1)
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst_smem)),
        "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(l2_evict_first_policy())
        : "memory");
  else tma_load_2d(dst_smem, tensor_map, c0, c1, bar);
}
__device__ __forceinline__ void bulk_s2g_out(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  if (PPK_L2_HINT & 2)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes), "l"(l2_evict_first_policy())
                 : "memory");
  else bulk_s2g(dst_gmem, src_smem, bytes);
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread may not exit (shared memory is released) before the source has been read
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace ppk
