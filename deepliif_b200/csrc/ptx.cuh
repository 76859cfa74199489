// Thin inline-PTX wrappers for the sm_100a features used by the kernels in this directory:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
// No CUTLASS/CuTe dependency: every instruction is spelled out so SASS can be audited
// (UTCHMMA / LDTM / UTMALDG, see profiles/).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dlb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a kernel that would otherwise hang (a strike on the shared GPU pool) traps
// instead after ~seconds, turning a protocol bug into a reported launch failure.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) { __trap(); }
  }
}

// Long waits (an epilogue warp waiting for a whole tile of MMAs, a converter warp waiting for a strip buffer): the
// try_wait carries a suspend-time hint, so the warp sleeps in hardware instead of polling — a polling warp is always
// eligible and takes issue slots from the warps that do work (ncu on the fused conv: 3.4 M poll iterations per launch).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (++spins > (1u << 20)) { __trap(); }
  }
}

// ----------------------------------------------------------------------------------------
// TMA: 5-D tiled tensor load into shared memory, completion on an mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}

// TMA store: one box of a tiled tensor map, shared -> global (bulk-group completion).  The shared-memory writes that filled
// the box must be made visible to the async proxy first (fence_proxy_async by the writers, then a barrier).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// at most ONE of this thread's committed bulk stores may still be reading shared memory
__device__ __forceinline__ void bulk_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's committed bulk stores have finished READING shared memory (the buffer may be rewritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 1-D bulk copy global -> shared (contiguous bytes, 16-byte aligned, size a multiple of 16), completion on an mbarrier.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: tensor memory + 5th-gen tensor core MMA
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {  // whole warp
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16 and bf16 operands, fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent issue: the whole warp executes the call (so the compiler may keep descriptors in uniform registers and
// needs no vector->uniform moves in front of every MMA), one elected lane issues.  A single-thread issue loop spends
// ~60-70 cycles per MMA on those moves, which bounds layers whose MMAs are short (N <= 64).
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One K = 64 chunk (four K = 16 steps: start-address fields + 2 = 32 bytes each) issued from ONE asm block: the operands cross
// from vector to uniform registers once per block instead of once per MMA, and the descriptor increments stay in the
// uniform datapath.  x3: split precision (lo*hi, hi*lo, hi*hi per step, small terms first); x1: single plane.
__device__ __forceinline__ void umma_f16_k64x3_elect(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
      uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e, t;\n\t.reg .b64 ah, al, bh, bl;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 ah, %1;\n\tmov.b64 al, %2;\n\tmov.b64 bh, %3;\n\tmov.b64 bl, %4;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, bh, %5, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 bh, bh, 2;\n\tadd.s64 bl, bl, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, bh, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 bh, bh, 2;\n\tadd.s64 bl, bl, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, bh, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 bh, bh, 2;\n\tadd.s64 bl, bl, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, bh, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm_k64x3_elect(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
      uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e, t;\n\t.reg .b64 ah, al, bh, bl;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 ah, %1;\n\tmov.b64 al, %2;\n\tmov.b64 bh, %3;\n\tmov.b64 bl, %4;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], al, bh, %5, p;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 bh, bh, 2;\n\tadd.s64 bl, bl, 2;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], al, bh, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 bh, bh, 2;\n\tadd.s64 bl, bl, 2;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], al, bh, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 bh, bh, 2;\n\tadd.s64 bl, bl, 2;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], al, bh, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bl, %5, t;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bh, %5, t;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_k64_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a, b;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 a, %1;\n\tmov.b64 b, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, p;\n\t"
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm_k64_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a, b;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 a, %1;\n\tmov.b64 b, %2;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t"
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, t;\n\t"
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, t;\n\t"
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, t;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------
// CTA pair (cta_group::2): two CTAs of a cluster run one M = 256 MMA; each holds its own 128 A rows and half of B.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [rem];\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the even (leader) CTA
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {  // whole warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued cta_group::2 MMAs arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      :
      : "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void umma_f16_2sm_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\t.reg .pred e;\n\tmov.b16 m, 3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      :
      : "r"(smem_u32(bar))
      : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B (64 x 16-bit), 8-row atoms 1024 B apart.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused for SW128 K-major: 1)
//   bits [32,46) stride byte offset >> 4   bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Same layout with an explicit stride between 8-row groups (SBO).  The 128B swizzle is a function of the ABSOLUTE shared-
// memory address bits (chunk bits [4,7) ^= row bits [7,10)), so a descriptor may start at any 128-byte row of a swizzled
// region and step between row groups by any multiple of 128 B with base_offset = 0 (measured on B200 with
// tools/umma_window_probe.cu: every start row 0..10 x SBO in {1024, 1280, 2048, 2304} reproduces the reference product).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16 instruction descriptor: fp32 accumulate, both operands K-major, no negate/sparsity.
//   [4,6) c_format=1 (F32)  [7,10) a_format  [10,13) b_format (0 = F16, 1 = BF16)
//   [17,23) N >> 3          [24,29) M >> 4
__host__ __device__ inline uint32_t make_idesc_f16(int m, int n, int is_bf16) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= static_cast<uint32_t>(is_bf16 ? 1 : 0) << 7;
  d |= static_cast<uint32_t>(is_bf16 ? 1 : 0) << 10;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

}  // namespace dlb
