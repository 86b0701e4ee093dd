// Inline-PTX wrappers shared by the tcgen05 convolution kernels (conv_tc.cu, conv_flat.cu):
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 alloc / mma / commit / ld, UMMA descriptors.
#pragma once

#include <cudaTypedefs.h>

#include "fx_common.cuh"

namespace fx {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (ptxas then knows the guarded region is single-threaded and feeds
// UTCHMMA / UTMALDG from uniform registers directly instead of a per-instruction broadcast loop).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Parity wait with a watchdog: a protocol bug must end in a trap (the launch fails with an
// error) rather than in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (spin == 64) t0 = global_ns();
        if (spin > 64 && (spin & 63) == 0 && global_ns() - t0 > 2000000000ull) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by one thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Weight-stationary form: the B operand is latched in collector buffer b<ID> (fill) and re-used by later
// MMAs (use / lastuse) without being fetched from shared memory again.
#define FX_UMMA_WS(NAME, SUFFIX)                                                                                          \
    __device__ __forceinline__ void NAME(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) { \
        asm volatile(                                                                                                     \
            "{\n\t.reg .pred p;\n\t"                                                                                     \
            "setp.ne.b32 p, %4, 0;\n\t"                                                                                   \
            "tcgen05.mma.ws.cta_group::1.kind::f16.collector::" SUFFIX " [%0], %1, %2, %3, p;\n\t}"                        \
            :                                                                                                             \
            : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)                                                 \
            : "memory");                                                                                                  \
    }
FX_UMMA_WS(umma_ws_b0_fill, "b0::fill")
FX_UMMA_WS(umma_ws_b1_fill, "b1::fill")
FX_UMMA_WS(umma_ws_b2_fill, "b2::fill")
FX_UMMA_WS(umma_ws_b3_fill, "b3::fill")
FX_UMMA_WS(umma_ws_b0_use, "b0::use")
FX_UMMA_WS(umma_ws_b1_use, "b1::use")
FX_UMMA_WS(umma_ws_b2_use, "b2::use")
FX_UMMA_WS(umma_ws_b3_use, "b3::use")
FX_UMMA_WS(umma_ws_b0_discard, "b0::discard")
#undef FX_UMMA_WS
template <int ID, bool FILL>
__device__ __forceinline__ void umma_ws(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (ID == 0) FILL ? umma_ws_b0_fill(d, da, db, idesc, acc) : umma_ws_b0_use(d, da, db, idesc, acc);
    if (ID == 1) FILL ? umma_ws_b1_fill(d, da, db, idesc, acc) : umma_ws_b1_use(d, da, db, idesc, acc);
    if (ID == 2) FILL ? umma_ws_b2_fill(d, da, db, idesc, acc) : umma_ws_b2_use(d, da, db, idesc, acc);
    if (ID == 3) FILL ? umma_ws_b3_fill(d, da, db, idesc, acc) : umma_ws_b3_use(d, da, db, idesc, acc);
}

// mbarrier arrive once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand whose K extent is exactly one swizzle atom
// (BK*2 bytes = 128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B).  8-row groups are `8*BK*2` bytes apart.
template <int BK>
__host__ __device__ constexpr uint64_t make_smem_desc(uint32_t saddr) {
    constexpr uint64_t sbo = (8 * BK * 2) >> 4;
    constexpr uint64_t layout = BK == 64 ? 2 : 4;  // SWIZZLE_128B : SWIZZLE_64B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// Instruction descriptor: c=f32, a=b=bf16, both K-major, M=128, N=BN.
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


// ---- CTA-pair (cta_group::2) forms: two CTAs of a cluster share one M=256 MMA; the even CTA issues ----
// In a cluster the CTA-in-pair index is bit 24 of a shared-window address; clearing it names the same
// offset in the even (leader) CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// TMA loads into THIS CTA's shared memory whose bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(dst), "l"(desc), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_2cta(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(dst), "l"(desc), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(desc), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the barrier at this offset in the leader CTA (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of both CTAs: 2 x 128 rows] * B[smem of both CTAs: 2 x N/2 rows]^T
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once the issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}

// Generic K-major shared-memory matrix descriptor: rows of ROWB bytes (= the swizzle span:
// 128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B, 32 -> SWIZZLE_32B), 8-row groups 8*ROWB bytes apart.
// The start address may sit any whole number of rows into a TMA-written tile: the swizzle XOR is
// a function of the absolute shared-memory address (pinned by fx_debug_umma_shift).
template <int ROWB>
__host__ __device__ constexpr uint64_t make_smem_desc_rowb(uint32_t saddr) {
    constexpr uint64_t sbo = (8 * ROWB) >> 4;
    constexpr uint64_t layout = ROWB == 128 ? 2 : (ROWB == 64 ? 4 : 6);
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// host helper shared by both files
int tc_encode_map(fx_engine* e, CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* estr, CUtensorMapSwizzle sw, const char* what);

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace fx
