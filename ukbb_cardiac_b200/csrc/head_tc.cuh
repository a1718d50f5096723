// Fused head, version 3 (north_star (b) + (c)): same algebra as head_mma.cuh -- same_dim0, the
// commuted fc0 with the bilinear upsampling as constant U_l matrices on the tensor cores, fc1 -- with
// the epilogues rebuilt around what the ncu captures of head_mma showed
// (profiles/r1_ncu_full_stage3_summary.txt, experiments/src_top.py): its class-score epilogue ran
// ~760 dependent SASS instructions per tile on ONE warp per SM sub-partition (64 x n_class FFMAs fed
// by 96 shared-memory loads, precise softmax, four ballots) and was busy 91 % of the time while the
// tensor pipe was 18 % active.
//
//   S0  D0[128x32] = b0_tile[128x16] . Wsd0^T                      (1 UMMA,  N = 32)      MMA warp 13
//   E0  A0 = relu(D0 + shift_sd0) -> 16 bit, K-major 64 B rows                            warps 0-3
//   S1  D1[128x64] = A0 . W_0^T + sum_l U_l . t_l patch             (2 + 7 UMMAs, N = 64)  MMA warp 14
//   E1  A2 = relu(D1 + shift_fc0) -> 16 bit, 128 B rows                                   warps 0-3
//   S2  D2[128x64] = A2 . W_fc1^T                                   (4 UMMAs)              MMA warp 15
//   E2  f = relu(D2 + shift_fc1) in FP32, class scores = f . W_logits + bias in FP32,      warps 4-11
//       softmax / argmax / crop / class counts (train_network.py:198-199, deploy_network.py:114-130)
// * The BN scales of same_dim0 / fc0 / fc1 are folded into the 16-bit weights (w * scale, then
//   rounded); shifts, class-score weights and bias travel BY VALUE in the kernel parameters, so
//   the epilogues read them as constant-bank / uniform-register operands: an E0 / E1 element is
//   one FADD plus half an F2FP.RELU, an E2 element is FADD + FMNMX + n_class FFMAs -- no shared-
//   memory loads, no registers spent on coefficients.
// * E2 runs on EIGHT warps: warps 4-7 own the even tiles (accumulator D2[0]), warps 8-11 the odd
//   tiles (D2[1]), so one warp has two tile periods for its ~450 instructions.
// * Labels: argmax over the FP32 softmax, lowest index on ties.  When the two largest scores differ
//   by more than 1e-5 the softmax cannot tie (exp(-1e-5) is 84 ulp below 1), so the argmax over the
//   scores is taken directly; closer calls (and calls that want prob / logits written) run the full
//   softmax.  Class counts: one REDUX per warp (a byte lane per class) instead of n_class ballots.
#pragma once
#include "tc_common.cuh"
#include "conv_group.cuh"      // tmem_ld32, TileWalk
#include "head_mma.cuh"        // HM_* layout constants, HeadMmaMaps, HeadParams

namespace ukbb {

namespace tc {
// (a0 + s0, a1 + s1) -> ReLU -> two 16-bit values; s0 / s1 are meant to be constant-bank operands
template <bool F16>
__device__ __forceinline__ uint32_t add_relu_pack(uint32_t a0, uint32_t a1, float s0, float s1) {
    uint64_t a, sh, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sh) : "f"(s0), "f"(s1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(sh));
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d));
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
}  // namespace tc

constexpr int H3_THREADS = 512;
constexpr int H3_STAGES = 6;                                      // input stages (= same_dim0 accumulator stages): HBM latency cover
constexpr int H3_SMEM = H3_STAGES * HM_IN_BYTES + 2 * HM_A0 + 2 * HM_A2 + HM_U_BYTES + HM_W0 + HM_W1 + HM_WSD + 1024 /*align*/ +
                        512 /*barriers*/ + 32 * 8 * 8 + 32 * 8 /*class-score weights and fc1 shifts as pairs*/;
constexpr int H3_D0 = 0, H3_D1 = 192, H3_D2 = 320;                // TMEM columns: 6 x 32, 2 x 64, 2 x 64

template <int NC, bool F16>
__global__ void __launch_bounds__(H3_THREADS, 1)
head_tc_kernel(const __grid_constant__ HeadMmaMaps maps, const __grid_constant__ HeadParams p) {
    using namespace tc;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t in_base = smem_base;
    const uint32_t a0_base = in_base + H3_STAGES * HM_IN_BYTES;
    const uint32_t a2_base = a0_base + 2 * HM_A0;
    const uint32_t u_base = a2_base + 2 * HM_A2;
    const uint32_t w0_base = u_base + HM_U_BYTES;
    const uint32_t w1_base = w0_base + HM_W0;
    const uint32_t wsd_base = w1_base + HM_W1;
    const uint32_t bar_base = wsd_base + HM_WSD;
    auto BAR = [&](int i) { return bar_base + 8u * i; };
    // One wait per MMA stage: the barrier an issuer waits on collects BOTH its operand ("full") and the
    // release of the accumulator it is about to overwrite:
    //   IN_FULL[s]  = TMA bytes of tile i            + 4 x E0(i-6) "D0[s] drained"      (count 1 + 4)
    //   A0_FULL[b]  = 4 x E0(i) "A0[b] written"       + 4 x E1(i-2) "D1[b] drained"      (count 8)
    //   A2_FULL[b]  = 4 x E1(i) "A2[b] written"       + 4 x E2(i-2) "D2[b] drained"      (count 8)
    // (the missing predecessors of the first tiles are pre-arrived once at start-up)
    enum { WFULL = 0, IN_FULL = 1, IN_EMPTY = IN_FULL + H3_STAGES, D0_FULL = IN_EMPTY + H3_STAGES, A0_FULL = D0_FULL + H3_STAGES,
           A0_EMPTY = A0_FULL + 2, D1_FULL = A0_EMPTY + 2, A2_FULL = D1_FULL + 2, A2_EMPTY = A2_FULL + 2, D2_FULL = A2_EMPTY + 2,
           TSLOT = D2_FULL + 2 };
    const uint32_t tmem_slot = BAR(TSLOT);
    float2* s_wl2 = reinterpret_cast<float2*>(smem_gen + (bar_base - smem_base) + 512);      // [32 channel pairs][8 classes]
    float2* s_sh2 = s_wl2 + 32 * 8;                                                            // [32] fc1 shift pairs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 12 && lane == 0) {
        const CUtensorMap* m = &maps.s0;
        for (int i = 0; i < 12; ++i) tma_prefetch_desc(m + i);
    }
    if (warp == 13 && lane == 0) {
        mbar_init(BAR(WFULL), 1);
        for (int s3 = 0; s3 < H3_STAGES; ++s3) { mbar_init(BAR(IN_FULL + s3), 5); mbar_init(BAR(IN_EMPTY + s3), 1); mbar_init(BAR(D0_FULL + s3), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(A0_FULL + b), 8); mbar_init(BAR(A0_EMPTY + b), 1); mbar_init(BAR(D1_FULL + b), 1);
            mbar_init(BAR(A2_FULL + b), 8); mbar_init(BAR(A2_EMPTY + b), 1); mbar_init(BAR(D2_FULL + b), 1);
        }
        fence_barrier_init();
    }
    if (warp == 15) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // zero the input stages once: the K-padding rows of the t_l patches are never written by TMA and
    // must be finite (they meet zero columns of U_l)
    for (int i = threadIdx.x; i < H3_STAGES * HM_IN_BYTES / 16; i += H3_THREADS)
        reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 32 * 8; i += H3_THREADS) s_wl2[i] = p.c_wl2[i >> 3][i & 7];
    if (threadIdx.x < 32) s_sh2[threadIdx.x] = make_float2(p.c_shift1[2 * threadIdx.x], p.c_shift1[2 * threadIdx.x + 1]);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 12) griddep_wait();                      // the producer waits after it has issued the weight loads
    // pre-arrivals standing for the "accumulator drained" signals that the first tiles have no predecessor for
    if (warp < 4 && lane == 0) {
        for (int s3 = 0; s3 < H3_STAGES; ++s3) mbar_arrive(BAR(IN_FULL + s3));
        mbar_arrive(BAR(A0_FULL + 0)); mbar_arrive(BAR(A0_FULL + 1));
    }
    if (warp >= 4 && warp < 12 && lane == 0) mbar_arrive(BAR(A2_FULL + ((warp - 4) >> 2)));
    const int my_tiles = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto LO = [](uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | (1u << 16); };
    constexpr uint32_t HI32 = (uint32_t)((8 * 32) >> 4) | (1u << 14) | (6u << 29);
    constexpr uint32_t HI64 = (uint32_t)((8 * 64) >> 4) | (1u << 14) | (4u << 29);
    constexpr uint32_t HI128 = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);

    if (warp == 12) {
        // ===================== TMA producer =====================
        // The five box loads of a tile (b0 tile + four t_l patches) are ONE warp instruction: lane l < 5 loads level l with
        // its own tensor map / destination / coordinates (measured: a tensor load costs its issuing thread ~330 cycles
        // whatever the box size, profiles/r1_rate_probe.log, so five back-to-back loads from one lane bound the tile period).
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(WFULL), HM_U_BYTES + HM_W0 + HM_W1 + HM_WSD);
            tma_load_2d(wsd_base, &maps.wsd, BAR(WFULL), 0, 0);
            tma_load_2d(u_base, &maps.u1, BAR(WFULL), 0, 0);
            tma_load_2d(u_base + HM_U1, &maps.u2, BAR(WFULL), 0, 0);
            tma_load_2d(u_base + HM_U1 + HM_U2, &maps.u3, BAR(WFULL), 0, 0);
            tma_load_2d(u_base + HM_U1 + HM_U2 + HM_U3, &maps.u4, BAR(WFULL), 0, 0);
            tma_load_2d(u_base + HM_U1 + HM_U2 + HM_U3 + 128 * 32, &maps.u4, BAR(WFULL), 0, 128);
            tma_load_2d(w0_base, &maps.w0, BAR(WFULL), 0, 0);
            tma_load_2d(w1_base, &maps.w1, BAR(WFULL), 0, 0);
        }
        griddep_wait();
        __syncwarp();
        TileWalk w;
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        const int l = lane < 5 ? lane : 0;
        const CUtensorMap* my_map = &maps.s0 + l;                                   // s0, t1, t2, t3, t4 are adjacent
        const uint32_t my_off = l == 0 ? 0u : l == 1 ? (uint32_t)HM_IN_S0 : l == 2 ? (uint32_t)(HM_IN_S0 + HM_IN_P1)
                                : l == 3 ? (uint32_t)(HM_IN_S0 + HM_IN_P1 + HM_IN_P2) : (uint32_t)(HM_IN_S0 + HM_IN_P1 + HM_IN_P2 + HM_IN_P3);
        const int pb = ((1 << l) - 1) >> 1, back = l > 0 ? 1 : 0;                   // level l patch origin: ((x0 + pb) >> l) - 1
        int s3 = 0;
        uint32_t ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int y0 = w.ty * 8, x0 = w.tx * 16, n = w.n;
            mbar_wait(BAR(IN_EMPTY + s3), ph ^ 1);
            const uint32_t dst = in_base + s3 * HM_IN_BYTES;
            const uint32_t fullb = BAR(IN_FULL + s3);
            if (lane == 0) mbar_arrive_expect_tx(fullb, HM_IN_TX);
            __syncwarp();
            if (lane < 5) tma_load_4d(dst + my_off, my_map, fullb, 0, ((x0 + pb) >> l) - back, ((y0 + pb) >> l) - back, n);
            __syncwarp();
            if (++s3 == H3_STAGES) { s3 = 0; ph ^= 1; }
            w.next();
        }
    } else if (warp == 13) {
        // ===================== MMA issuer 0: same_dim0 (S0) =====================
        const bool leader = elect_one();
        const uint32_t idesc_sd = F16 ? make_idesc_f16(128, 32) : make_idesc_bf16(128, 32);
        const uint32_t wsd_lo = LO(wsd_base);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        int s3 = 0;
        uint32_t ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(BAR(IN_FULL + s3), ph);
            tc_fence_after();
            if (leader) {
                umma_bf16_lohi(tmem_base + H3_D0 + s3 * 32, LO(in_base + s3 * HM_IN_BYTES), HI32, wsd_lo, HI32, idesc_sd, 0u);
                umma_commit(BAR(D0_FULL + s3));
            }
            __syncwarp();
            if (++s3 == H3_STAGES) { s3 = 0; ph ^= 1; }
        }
    } else if (warp == 14) {
        // ===================== MMA issuer 1: fc0 with the upsample terms (S1) =====================
        const bool leader = elect_one();
        const uint32_t idesc_kk = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        const uint32_t idesc_kmn = idesc_kk | (1u << 16);            // B operand MN-major (pixel-major t_l patch)
        const uint32_t u1_lo = LO(u_base), u2_lo = LO(u_base + HM_U1), u3_lo = LO(u_base + HM_U1 + HM_U2),
                       u4_lo = LO(u_base + HM_U1 + HM_U2 + HM_U3);
        const uint32_t w0_lo = LO(w0_base);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        TileWalk w;                                                 // needs the tile-row parity for U_4
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        int s3 = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            const uint32_t v = (uint32_t)(w.ty & 1);                 // tile-row parity selects the U_4 variant
            mbar_wait(BAR(A0_FULL + b), ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + H3_D1 + b * 64;
            const uint32_t in_lo = LO(in_base + s3 * HM_IN_BYTES);
            const uint32_t a0_lo = LO(a0_base + b * HM_A0);
            if (leader) {
                umma_bf16_lohi(d, a0_lo, HI64, w0_lo, HI64, idesc_kk, 0u);
                umma_bf16_lohi(d, a0_lo + 2, HI64, w0_lo + 2, HI64, idesc_kk, 1u);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    umma_bf16_lohi(d, u1_lo + 2 * k, HI128, in_lo + ((HM_IN_S0 + k * 2048) >> 4), HI128, idesc_kmn, 1u);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_bf16_lohi(d, u2_lo + 2 * k, HI64, in_lo + ((HM_IN_S0 + HM_IN_P1 + k * 2048) >> 4), HI128, idesc_kmn, 1u);
                umma_bf16_lohi(d, u3_lo, HI32, in_lo + ((HM_IN_S0 + HM_IN_P1 + HM_IN_P2) >> 4), HI128, idesc_kmn, 1u);
                umma_bf16_lohi(d, u4_lo + v * ((128 * 32) >> 4), HI32, in_lo + ((HM_IN_S0 + HM_IN_P1 + HM_IN_P2 + HM_IN_P3) >> 4), HI128,
                               idesc_kmn, 1u);
                umma_commit(BAR(IN_EMPTY + s3));
                umma_commit(BAR(A0_EMPTY + b));
                umma_commit(BAR(D1_FULL + b));
            }
            __syncwarp();
            if (++s3 == H3_STAGES) s3 = 0;
            w.next();
        }
    } else if (warp == 15) {
        // ===================== MMA issuer 2: fc1 (S2) =====================
        const bool leader = elect_one();
        const uint32_t idesc_kk = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        const uint32_t w1_lo = LO(w1_base);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            mbar_wait(BAR(A2_FULL + b), ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + H3_D2 + b * 64;
            const uint32_t a_lo = LO(a2_base + b * HM_A2);
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, a_lo + 2 * k, HI128, w1_lo + 2 * k, HI128, idesc_kk, k != 0 ? 1u : 0u);
                umma_commit(BAR(A2_EMPTY + b));
                umma_commit(BAR(D2_FULL + b));
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ===================== E0 (D0 -> A0, two tiles ahead) and E1 (D1 -> A2), warps 0-3 =====================
        const int q = warp;
        const int r = q * 32 + lane;
        // E1 trails E0 by TWO tiles: S1(i) (9 UMMAs) is issued at the end of iteration i and its accumulator is first
        // needed in iteration i + 2, so this warp never sits waiting for the tensor pipe between its two roles.
        for (int it = 0; it <= my_tiles + 1; ++it) {
            if (it >= 2) {
                const int i = it - 2, b = i & 1;
                const uint32_t ph = ((uint32_t)i >> 1) & 1u;
                mbar_wait(BAR(A2_EMPTY + b), ph ^ 1);              // satisfied long before the accumulator is: off the critical path
                mbar_wait(BAR(D1_FULL + b), ph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + H3_D1 + b * 64;
                uint32_t v[64];
                tmem_ld32(taddr, v);
                tmem_ld32(taddr + 32, v + 32);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(A0_FULL + b));       // D1[b] drained: half of the arrivals S1(i + 2) waits for
                const uint32_t row = a2_base + b * HM_A2 + r * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t o[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) o[u] = add_relu_pack<F16>(v[8 * j + 2 * u], v[8 * j + 2 * u + 1], p.c_shift0[8 * j + 2 * u],
                                                                          p.c_shift0[8 * j + 2 * u + 1]);
                    const uint32_t dst = row + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(A2_FULL + b));
            }
            if (it < my_tiles) {
                const int i = it, b = i & 1, s3 = i % H3_STAGES;
                const uint32_t ph = ((uint32_t)i >> 1) & 1u;
                mbar_wait(BAR(A0_EMPTY + b), ph ^ 1);              // satisfied long before the accumulator is: off the critical path
                mbar_wait(BAR(D0_FULL + s3), (uint32_t)(i / H3_STAGES) & 1u);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + H3_D0 + s3 * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(IN_FULL + s3));       // D0[s3] drained: one of the 5 arrivals S0(i + H3_STAGES) waits for
                const uint32_t row = a0_base + b * HM_A0 + r * 64;
                const uint32_t sw = ((uint32_t)r >> 1) & 3u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t o[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) o[u] = add_relu_pack<F16>(v[8 * j + 2 * u], v[8 * j + 2 * u + 1], p.c_shift_sd0[8 * j + 2 * u],
                                                                          p.c_shift_sd0[8 * j + 2 * u + 1]);
                    const uint32_t dst = row + (((uint32_t)j ^ sw) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(A0_FULL + b));
            }
        }
    } else if (warp < 12) {
        // ===================== E2: FP32 class scores -> labels; warps 4-7 even tiles, warps 8-11 odd tiles =====================
        const int b = (warp - 4) >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int ty = r >> 4, tx = r & 15;
        TileWalk w;
        w.init(blockIdx.x + b * gridDim.x, 2 * gridDim.x, p.tiles_x, p.tiles_y);
        uint32_t ph = 0;
        for (int i = b; i < my_tiles; i += 2, ph ^= 1) {
            mbar_wait(BAR(D2_FULL + b), ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + H3_D2 + b * 64;
            uint32_t v[64];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A2_FULL + b));           // D2[b] drained: half of the arrivals S2(i + 2) waits for
            // class scores in FP32 on plain FFMAs whose second operand is a CONSTANT-BANK word (the weights, shifts and
            // bias travel by value in the kernel parameters): no shared-memory loads and no register-pair packing moves --
            // the head is bound by shared-memory bandwidth (UMMA operand reads + epilogue stores + TMA writes), so the
            // epilogue must not add broadcast LDS traffic of its own.
            float lg[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) lg[c] = p.c_bias[c];                                   // -inf for c >= n_class
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                const float f = fmaxf(__uint_as_float(v[k]) + p.c_shift1[k], 0.f);
#pragma unroll
                for (int c = 0; c < NC; ++c) lg[c] = fmaf(f, (k & 1) ? p.c_wl2[k >> 1][c].y : p.c_wl2[k >> 1][c].x, lg[c]);
            }
            const int n = w.n, y = w.ty * 8 + ty, x = w.tx * 16 + tx;
            float m1 = lg[0], m2 = -INFINITY;
            int arg = 0;
#pragma unroll
            for (int c = 1; c < NC; ++c) {
                if (lg[c] > m1) { m2 = m1; m1 = lg[c]; arg = c; }
                else m2 = fmaxf(m2, lg[c]);
            }
            const bool full = p.prob != nullptr || p.logits != nullptr;
            if (full || __any_sync(0xffffffffu, !(m1 - m2 > 1e-5f))) {
                float e[NC], ssum = 0.f;
#pragma unroll
                for (int c = 0; c < NC; ++c) { e[c] = expf(lg[c] - m1); ssum += e[c]; }
                float best = -1.f;
                arg = 0;
                const size_t pix = ((size_t)n * p.h + y) * p.w + x;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float pr = e[c] / ssum;
                    if (pr > best) { best = pr; arg = c; }
                    if (full && c < p.nc) {
                        if (p.prob) p.prob[pix * p.nc + c] = pr;
                        if (p.logits) p.logits[pix * p.nc + c] = lg[c];
                    }
                }
            }
            const int yy = y - p.y_pre, xx = x - p.x_pre;
            const bool inside = yy >= 0 && yy < p.y && xx >= 0 && xx < p.x;
            if (inside) p.labels[((size_t)n * p.y + yy) * p.x + xx] = (uint8_t)arg;
            if (p.counts) {
                // per-class counts of the warp's 32 pixels in one REDUX: one byte lane per class
                const unsigned lo4 = __reduce_add_sync(0xffffffffu, (inside && arg < 4) ? (1u << (8 * arg)) : 0u);
                unsigned hi4 = 0;
                if (NC > 4) hi4 = __reduce_add_sync(0xffffffffu, (inside && arg >= 4) ? (1u << (8 * (arg - 4))) : 0u);
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const unsigned cnt = ((c < 4 ? lo4 : hi4) >> (8 * (c & 3))) & 0xffu;
                        if (cnt && c < p.nc) atomicAdd(&p.counts[(size_t)n * p.nc + c], (unsigned long long)cnt);
                    }
                }
            }
            w.next();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 15) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace ukbb
