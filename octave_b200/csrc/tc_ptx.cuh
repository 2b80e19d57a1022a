// Inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 alloc / mma / commit / ld, and the UMMA shared-memory + instruction descriptors.
// Bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (wake-up is immediate) or the
// hint (~10 ms) expires, instead of returning after the short default window — a waiting role then issues a handful of
// instructions per wait, not a spin loop that competes with the epilogue warps for issue slots (ncu: the spin loops of the
// five waiting warps were ~25 % of all issued instructions of a narrow convolution).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU (2^10 tries of up to ~10 ms each).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 10); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}

// One lane of a fully converged warp (all 32 lanes must execute this).  Role loops run warp-uniform — every lane walks
// the tiles and waits on the barriers — and only the issuing instruction (TMA / MMA / commit) sits under this predicate:
// descriptors and coordinates then live in uniform registers.  Issuing from inside an `if (lane == 0)` region instead
// makes ptxas move every operand through R2UR + a BRA.U.ANY loop (~15 SASS instructions per tcgen05.mma, measured).
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

// ---------------------------------------------------------------- TMA loads (tile mode, mbarrier completion)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// ---------------------------------------------------------------- TMA stores (shared -> global, bulk-group completion)
// Out-of-bounds parts of the box are clipped by the hardware.  The issuing thread commits the group and later waits for the
// READ of the shared-memory source to finish before the staging buffer is overwritten (wait_group.read), and for full
// completion before the CTA exits.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// same box, but global += shared (element-wise add in the tensor map's data type, done by the L2): the accumulate-into-dx
// epilogue of a residual block's second data gradient needs no read-modify-write in the SM
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// named barrier over `nthreads` threads (a multiple of 32) of the CTA; id 0 is __syncthreads()
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 bit, 16 consecutive columns -> 16 registers per thread (lane i of the warp reads TMEM lane
// base_lane + i; a warp may only touch the lane quarter 32*(warp_id % 4)).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor.  swizzle_bytes in {128, 64, 32}.
//  K-major : rows of `swizzle_bytes` bytes, 8-row atoms; SBO = 8 * swizzle_bytes; LBO unused (1).
//  MN-major: rows (one K index) of `swizzle_bytes` bytes holding the MN run, 8-row atoms;
//            SBO = 8 * swizzle_bytes (next 8 K indices); LBO = byte distance to the next MN run.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t swizzle_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);               // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;          // leading byte offset, bits [16,30)
  d |= (uint64_t)(((8u * swizzle_bytes) >> 4) & 0x3FFFu) << 32;  // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                     // descriptor version (Blackwell)
  uint64_t layout = swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6);
  d |= layout << 61;                                          // swizzle mode, bits [61,64)
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_advance(uint64_t desc, uint32_t bytes) {
  // only the 14-bit start-address field moves; callers stay inside one 16 KB-aligned window
  return desc + (uint64_t)(bytes >> 4);
}
// Instruction descriptor for kind::f16 with BF16 A/B and FP32 D.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // D format = F32
         | (1u << 7)                     // A format = BF16
         | (1u << 10)                    // B format = BF16
         | ((uint32_t)a_mn_major << 15)  // A major: 0 = K, 1 = MN
         | ((uint32_t)b_mn_major << 16)  // B major
         | ((uint32_t)(N >> 3) << 17)    // N / 8
         | ((uint32_t)(M >> 4) << 24);   // M / 16
}

}  // namespace tc
