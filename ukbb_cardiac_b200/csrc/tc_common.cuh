// sm_100a building blocks: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld) wrappers and UMMA descriptor construction.  Inline PTX only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace ukbb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "     elect.sync %%rx|%%px, 0xffffffff;\n"
        "@%%px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a lost TMA transaction or a descriptor fault must not hang the GPU (the box is shared); after 2^20 failed probes
// (a probe suspends the warp for a hardware-defined time, tens of cycles measured) the kernel traps and the launch returns an error.
// A suspend-time hint on the probe and a __nanosleep(32 / 128 ns) back-off after a failed probe were measured and changed nothing
// (4567 / 4599 / 4594 vs 4591 us per subject): a third of the head's issued instructions are probes, but the busy warps are bound
// by their own dependency chains, not by issue slots.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spins > (1u << 20)) {
#ifdef UKBB_DEBUG_MBAR
            printf("mbar timeout: bar %x parity %u block %d thread %d\n", bar, parity, (int)blockIdx.x, (int)threadIdx.x);
#endif
            __trap();
        }
    }
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every tensor-core kernel of the forward is launched with programmatic stream serialization: its CTAs may start
// (barrier init, TMEM allocation, weight loads) while the previous kernel drains its last tiles.  griddep_wait()
// blocks until the previous kernel has completed and its writes are visible; it must precede the first access to
// activation memory.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---------------------------------------------------------------- thread-block clusters (TMA multicast of shared operands)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one load, delivered to the same shared-memory offset (and signalled on the same mbarrier offset) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// arrive on the mbarrier at the same offset in every CTA of cta_mask once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
                 : "memory");
}

// ---------------------------------------------------------------- CTA pairs (tcgen05 cta_group::2)
// Two CTAs of a cluster execute ONE tcgen05.mma with M = 256: each CTA supplies its own 128 rows of A and HALF of the N rows of B
// from the same shared-memory offsets, and receives its 128 accumulator rows in its own tensor memory.  The leader (cluster rank 0)
// issues; TMA loads of both CTAs signal the LEADER's mbarrier (its shared::cluster address), tcgen05.commit multicasts to both.
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const void* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// kind: 0 = kind::f16 accumulate / overwrite, 1 = kind::f8f6f4, 2 = kind::f16 with D <- A.B + D * 2^-15
template <int KIND>
__device__ __forceinline__ void umma_pair_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 2) {
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p, 15;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    } else if (KIND == 1) {
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
            "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc),
            "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc),
            "r"(accumulate)
            : "memory");
    }
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, BF16 operands, FP32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, with the descriptors given as (lo, hi) 32-bit halves so that a K / tap / window advance
// is a single 32-bit add on the low word (start address field, 16-byte units).
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- "x2" scheme (UKBB_MODE_FP16X2): the FP16 main term plus BOTH correction terms of the split product in ONE FP8 pass.
// Every 16-channel group of the lo plane holds [e4m3(lo * 2^11) x 16 | e4m3(hi) x 16] (32 bytes, the size of 16 FP16 lo values, so
// every layout, descriptor and store of the x3 scheme is unchanged) and the weights' lo plane [e4m3(w_hi * 2^4) x 16 | e4m3(w_lo * 2^15) x 16]:
// one kind::f8f6f4 instruction (K = 32) per group accumulates (a_lo w_hi + a_hi w_lo) * 2^15.  All FP8 instructions of an
// accumulator are issued first; the FIRST kind::f16 instruction then rescales it with scale-input-d = 15 (D <- A.B + D * 2^-15).
// Measured on B200 (experiments/x2f8_probe.cu): bit-identical to this model, relative rms error of a product 9.6e-6.
constexpr int X2_SA = 11, X2_SW_HI = 4, X2_SW_LO = 15;
// instruction descriptor kind::f8f6f4: E4M3 x E4M3 -> F32 (a_format = b_format = 0), K-major operands
__host__ __device__ constexpr uint32_t make_idesc_e4m3(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f8_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D <- A.B + D * 2^-15 (kind::f16 with scale-input-d)
__device__ __forceinline__ void umma_f16_lohi_rescale(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, 1, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p, 15;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
}
// same two with the A operand in tensor memory
__device__ __forceinline__ void umma_ts_f8_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_lohi_rescale(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "setp.ne.b32 p, 1, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p, 15;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
}

// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T : A is [128 lanes][K = 16 as 8 columns of two 16-bit values]
__device__ __forceinline__ void umma_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Instruction descriptor, kind::f16: BF16 x BF16 -> FP32, both operands K-major, M x N tile.
// (bit layout: c_format [4,6) = 1 (F32); a_format [7,10) = 1 (BF16); b_format [10,13) = 1;
//  a_major bit 15 = 0, b_major bit 16 = 0 (K-major); n_dim [17,23) = N >> 3; m_dim [24,29) = M >> 4)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// same with FP16 operands (a_format = b_format = 0): identical tensor-core rate, 11-bit significand
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Shared-memory matrix descriptor for a K-major operand tile whose rows are `row_bytes` wide
// (32 / 64 / 128 -> SWIZZLE_32B / 64B / 128B, the same mode the TMA tensor map uses): rows are
// stored contiguously, 8-row groups are row_bytes*8 apart (SBO); LBO is unused for swizzled
// K-major layouts (set to 1 like CUTLASS); version = 1 (Blackwell); base_offset = 0 (tiles are
// 1024-byte aligned).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : row_bytes == 64 ? 4ull : 6ull;
    const uint64_t sbo = (uint64_t)(row_bytes * 8) >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// 16-bit operand format of the tensor-core path: 0 = BF16, 1 = FP16 (values clamped to +-65504)
__device__ __forceinline__ uint32_t pack16(float a, float b, int fp16) {
    if (fp16) {
        __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
        return *reinterpret_cast<uint32_t*>(&h);
    }
    return pack_bf16(a, b);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16t(float a, float b) {
    if (F16) {
        __half2 h = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));      // inputs are >= 0 (post-ReLU) or small
        return *reinterpret_cast<uint32_t*>(&h);
    }
    return pack_bf16(a, b);
}
template <bool F16>
__device__ __forceinline__ float2 unpack16t(uint32_t u) {
    if (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// split-operand ("x3") modes: a pair of FP32 values -> 16-bit hi pieces and 16-bit lo pieces, v ~ hi + lo with hi = rn16(v) and
// lo = rn16(v - hi) (v - hi is exact in FP32).  FP16: hi is clamped to the finite range; inputs are >= 0 (post-ReLU) or small.
template <bool F16>
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
    if (F16) {
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        uint64_t x, hh;
        float l0, l1;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
        asm("mov.b64 %0, {%1, %2};" : "=l"(hh) : "f"(h.x), "f"(h.y));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(x), "l"(hh));                       // one packed subtract for the pair
        asm("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(l1) : "l"(x));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
    } else {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - __uint_as_float(hi & 0xffff0000u)), "f"(a - __uint_as_float(hi << 16)));
    }
}
// max(v, 0) -> (hi, lo) with the ReLU folded into the conversions (FP16): hi = rz16(max(v, 0)) by cvt.rz.relu, so v - hi lies in [0, ulp)
// for v >= 0 and equals v < 0 otherwise, and lo = cvt.rn.relu(v - hi) is right in both cases -- two FMNMX less per pair than ReLU first,
// at the price of one bit (hi truncated instead of rounded: 21 instead of 22 significant bits).  BF16: ReLU, then split_pack.
template <bool F16>
__device__ __forceinline__ void split_pack_relu(float a, float b, uint32_t& hi, uint32_t& lo) {
    if (F16) {
        asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        uint64_t x, hh;
        float l0, l1;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
        asm("mov.b64 %0, {%1, %2};" : "=l"(hh) : "f"(h.x), "f"(h.y));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(x), "l"(hh));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(l1) : "l"(x));
        asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
    } else {
        split_pack<false>(fmaxf(a, 0.f), fmaxf(b, 0.f), hi, lo);
    }
}
// Four consecutive channels (quad q = 0..3 of a 16-channel group) -> the group's output words: oh[2q], oh[2q + 1] = the 16-bit hi pieces;
// ol[2q], ol[2q + 1] = the 16-bit lo pieces, or (F8, x2 scheme) ol[q] = four e4m3 bytes of (x - hi) * 2^11 and ol[4 + q] = four e4m3 bytes
// of hi, i.e. the eight words of a group are [lo8 x 16 | hi8 x 16] with no shuffle afterwards.  hi8 is converted from the packed FP16
// pair (cvt.e4m3x2.f16x2), the scaling of the lo pieces is one packed multiply per pair.
// RELU: the inputs have NOT been through the ReLU yet (applied here: explicitly for the FP8 layout, folded into the conversions otherwise).
template <bool F16, bool F8, bool RELU = false>
__device__ __forceinline__ void split_pack4(float a0, float a1, float a2, float a3, uint32_t* oh, uint32_t* ol, int q) {
    if (F8) {
        if (RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
        uint32_t h01, h23;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(a1), "f"(a0));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(a3), "f"(a2));
        const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01)), f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
        uint64_t x01, x23, g01, g23, k;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x01) : "f"(a0), "f"(a1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(x23) : "f"(a2), "f"(a3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(g01) : "f"(f01.x), "f"(f01.y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(g23) : "f"(f23.x), "f"(f23.y));
        asm("mov.b64 %0, {%1, %1};" : "=l"(k) : "f"(2048.f));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(x01) : "l"(x01), "l"(g01));                   // x - hi is exact in FP32, and so is * 2^11
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(x23) : "l"(x23), "l"(g23));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(x01) : "l"(x01), "l"(k));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(x23) : "l"(x23), "l"(k));
        float l0, l1, l2, l3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(l1) : "l"(x01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(l2), "=f"(l3) : "l"(x23));
        uint16_t p01, p23, q01, q23;
        asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(p01) : "f"(l1), "f"(l0));
        asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(p23) : "f"(l3), "f"(l2));
        asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(q01) : "r"(h01));
        asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(q23) : "r"(h23));
        oh[2 * q] = h01; oh[2 * q + 1] = h23;
        ol[q] = (uint32_t)p01 | ((uint32_t)p23 << 16);
        ol[4 + q] = (uint32_t)q01 | ((uint32_t)q23 << 16);
    } else {
        if (RELU) {
            split_pack_relu<F16>(a0, a1, oh[2 * q], ol[2 * q]);
            split_pack_relu<F16>(a2, a3, oh[2 * q + 1], ol[2 * q + 1]);
        } else {
            split_pack<F16>(a0, a1, oh[2 * q], ol[2 * q]);
            split_pack<F16>(a2, a3, oh[2 * q + 1], ol[2 * q + 1]);
        }
    }
}
__device__ __forceinline__ float2 unpack16(uint32_t u, int fp16) {
    if (fp16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}

}  // namespace tc
}  // namespace ukbb
