// BF16 tensor-core path of the FCN forward (north_star (a)-(c)), sm_100a only.
//
// conv_tc_kernel: implicit-GEMM convolution (3x3 stride 1/2 or 1x1) + folded BatchNorm + ReLU
//   (common/network.py:19-25) on tcgen05 tensor cores:
//     M = 128 output pixels of a (bn x bh x bw) box of the NHWC activation tensor,
//     N = Cout, K = taps x Cin, accumulated in TMEM (FP32), operands staged by TMA.
//   For every K block (one filter tap x CC input channels) the producer issues
//     * one 4-D tiled TMA load of the activation box shifted by the tap offset -- the
//       out-of-bounds zero fill of TMA IS the TF 'SAME' padding, and the traversal stride of
//       the tensor map IS the conv stride -- landing as a K-major [128][CC] bf16 tile, and
//     * one 2-D TMA load of the [Cout][CC] weight slice,
//   both with the 32/64/128-byte swizzle that the UMMA shared-memory descriptors name.
//   Warp roles (256 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA
//   issuer (one elected lane), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld ->
//   scale/shift/ReLU -> bf16 -> global NHWC).  Two accumulator stages in TMEM let the epilogue
//   of tile i overlap the MMAs of tile i+1.
#include "tc_plan.cuh"
#include <stdlib.h>

namespace ukbb {

using namespace tc;

// SPLIT (the "x3" modes): every activation and weight is the sum of two 16-bit pieces v = hi + lo (hi = rn16(v), lo = rn16(v - hi));
// the product is accumulated as hi.hi + lo.hi + hi.lo in the same FP32 accumulator (the lo.lo term, ~2^-22 relative in FP16,
// ~2^-16 in BF16, is dropped).  A stage then holds four operand tiles: A_hi | A_lo | B_hi | B_lo.
// F8 (x2 scheme, tc_common.cuh): the K blocks are walked TWICE -- first the lo planes (FP8 correction operands, kind::f8f6f4), then the hi
// planes (kind::f16, the first instruction rescales the accumulator) -- so a stage holds one plane pair A | B like the plain kernel.
template <int CC, int COUT, bool SPLIT = false, bool F8 = false>
struct ConvTcCfg {
    static constexpr int A_BYTES = 128 * CC * 2;
    static constexpr int B_BYTES = COUT * CC * 2;
    static constexpr int B_PAD = (B_BYTES + 1023) / 1024 * 1024;
    static constexpr int PLANES = (SPLIT && !F8) ? 2 : 1;                              // operand planes per stage
    static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_PAD);                     // all tiles 1024-aligned
    static constexpr int MAX_STAGES = (200 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
    static constexpr int TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;   // power of two for COUT in {16..256}
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int CC, int COUT, bool F16, bool SPLIT = false, bool F8 = false>
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const ConvTcParams p) {
    using Cfg = ConvTcCfg<CC, COUT, SPLIT, F8>;
    constexpr bool X3 = SPLIT && !F8;
    constexpr int STAGES = Cfg::STAGES;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = p.cin / CC;
    const int kblocks = p.taps * chunks;
    const int kb_total = F8 ? 2 * kblocks : kblocks;          // x2: [0, kblocks) = lo planes, [kblocks, 2 kblocks) = hi planes

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 0) griddep_wait();                       // the producer waits after it has issued the weight loads

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            griddep_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tn = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.bw * p.stride - p.pad_left, y0 = ty * p.bh * p.stride - p.pad_top, n0 = tn * p.bn;
                for (int kbt = 0; kbt < kb_total; ++kbt) {
                    const int kb = (F8 && kbt >= kblocks) ? kbt - kblocks : kbt;
                    const bool lo_pass = F8 && kbt < kblocks;
                    const int tap = kb / chunks, ch = kb - tap * chunks;
                    const int ky = tap / p.ks, kx = tap - ky * p.ks;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + Cfg::PLANES * Cfg::A_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), Cfg::PLANES * (Cfg::A_BYTES + Cfg::B_BYTES));
                    tma_load_4d(a_dst, &map_a, full_bar(stage), ch * CC, x0 + kx, y0 + ky, (lo_pass ? p.lo_n : 0) + n0);
                    tma_load_2d(b_dst, &map_b, full_bar(stage), p.kofs + tap * p.cin + ch * CC, lo_pass ? COUT : 0);
                    if (X3) {                    // lo planes: slices [lo_n, ...) of the activation map, rows [COUT, 2 COUT) of the weight map
                        tma_load_4d(a_dst + Cfg::A_BYTES, &map_a, full_bar(stage), ch * CC, x0 + kx, y0 + ky, p.lo_n + n0);
                        tma_load_2d(b_dst + Cfg::B_PAD, &map_b, full_bar(stage), p.kofs + tap * p.cin + ch * CC, COUT);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, COUT) : make_idesc_bf16(128, COUT);
        const uint32_t idesc8 = make_idesc_e4m3(128, COUT);
        constexpr uint32_t RB = CC * 2;
        constexpr uint32_t HI = (uint32_t)((8 * RB) >> 4) | (1u << 14) | ((RB == 128 ? 2u : RB == 64 ? 4u : 6u) << 29);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + acc * COUT;
            for (int kb = 0; kb < kb_total; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_lo = (((smem_base + stage * Cfg::STAGE_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
                const uint32_t b_lo = a_lo + ((Cfg::PLANES * Cfg::A_BYTES) >> 4);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < CC / 16; ++k) {
                        if (F8) {
                            if (kb < kblocks) umma_f8_lohi(d, a_lo + 2 * k, HI, b_lo + 2 * k, HI, idesc8, (kb | k) != 0 ? 1u : 0u);
                            else if (kb == kblocks && k == 0) umma_f16_lohi_rescale(d, a_lo, HI, b_lo, HI, idesc);
                            else umma_bf16_lohi(d, a_lo + 2 * k, HI, b_lo + 2 * k, HI, idesc, 1u);
                            continue;
                        }
                        umma_bf16_lohi(d, a_lo + 2 * k, HI, b_lo + 2 * k, HI, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (X3) {
                            umma_bf16_lohi(d, a_lo + (Cfg::A_BYTES >> 4) + 2 * k, HI, b_lo + 2 * k, HI, idesc, 1u);                  // lo . hi
                            umma_bf16_lohi(d, a_lo + 2 * k, HI, b_lo + (Cfg::B_PAD >> 4) + 2 * k, HI, idesc, 1u);                    // hi . lo
                        }
                    }
                    umma_commit(empty_bar(stage));          // frees the smem stage when the MMAs retire
                    if (kb == kb_total - 1) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;                              // TMEM lane quarter of this warp
        const int r = q * 32 + lane;                         // row of the 128-pixel tile
        const int rx = r % p.bw, ry = (r / p.bw) % p.bh, rn = r / (p.bw * p.bh);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tn = tile / (p.tiles_x * p.tiles_y);
            const int ox = tx * p.bw + rx, oy = ty * p.bh + ry, on = tn * p.bn + rn;
            const bool live = on < p.n && oy < p.ho && ox < p.wo;
            __nv_bfloat16* dst = p.out + (((size_t)on * p.ho + oy) * p.wo + ox) * COUT;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * COUT;
#pragma unroll 1
            for (int c = 0; c < COUT; c += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c, v);
                tmem_ld_wait();
                uint32_t o[8], ol[8];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c) + j4);
                    const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + c) + j4);
                    float a0 = fmaf(__uint_as_float(v[4 * j4]), sc.x, sh.x), a1 = fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y);
                    float a2 = fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), a3 = fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w);
                    if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
                    else if (F16) { a0 = fmaxf(a0, -65504.f); a1 = fmaxf(a1, -65504.f); a2 = fmaxf(a2, -65504.f); a3 = fmaxf(a3, -65504.f); }
                    if (SPLIT) {
                        split_pack4<F16, F8>(a0, a1, a2, a3, o, ol, j4);
                    } else {
                        o[2 * j4] = pack16t<F16>(a0, a1);
                        o[2 * j4 + 1] = pack16t<F16>(a2, a3);
                    }
                }
                if (live) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst + c);
                    d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    if (SPLIT) {
                        uint4* l4 = reinterpret_cast<uint4*>(dst + p.out_lo + c);
                        l4[0] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                        l4[1] = make_uint4(ol[4], ol[5], ol[6], ol[7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// Launchers of the three convolution kernel families
// ------------------------------------------------------------------------------------------
template <int CC, int COUT, bool F16, bool SPLIT, bool F8 = false>
static int launch_tc2(const TcLayerPlan& P, int sms, cudaStream_t st) {
    using Cfg = ConvTcCfg<CC, COUT, SPLIT, F8>;
    static bool attr_set[kMaxDevices] = {};          // cudaFuncSetAttribute is per device (one process may drive several: SplitEngine)
    if (first_use_on_device(attr_set)) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CC, COUT, F16, SPLIT, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    const int grid = P.p.n_tiles < sms ? P.p.n_tiles : sms;
    UKBB_CUDA(launch_pdl(conv_tc_kernel<CC, COUT, F16, SPLIT, F8>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.p));
    return UKBB_OK;
}
template <int CC, int COUT>
static int launch_tc(const TcLayerPlan& P, int fp16, int sms, cudaStream_t st) {
    return fp16 ? launch_tc2<CC, COUT, true, false>(P, sms, st) : launch_tc2<CC, COUT, false, false>(P, sms, st);
}
template <int CC, int COUT>
static int launch_tc_x3(const TcLayerPlan& P, int fp16, int sms, cudaStream_t st) {
    return fp16 ? launch_tc2<CC, COUT, true, true>(P, sms, st) : launch_tc2<CC, COUT, false, true>(P, sms, st);
}

// levels 3 / 4: streamed weights, clusters of two CTAs share every weight tile through TMA multicast (conv_halo.cuh)
template <int CC, int COUT, bool F16, bool SPLIT, bool F8 = false, bool PAIR = false>
static int launch_halo2(const TcLayerPlan& P, int sms, cudaStream_t st) {
    using Cfg = ConvHaloCfg<CC, COUT, false, 0, SPLIT, F8, PAIR>;
    static bool attr_set[kMaxDevices] = {};          // cudaFuncSetAttribute is per device (one process may drive several: SplitEngine)
    if (first_use_on_device(attr_set)) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<CC, COUT, false, 0, F16, true, SPLIT, F8, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    const int pairs = (P.hp.n_tiles + 1) / 2;
    const int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);
    UKBB_CUDA(launch_pdl<2>(conv_halo_kernel<CC, COUT, false, 0, F16, true, SPLIT, F8, PAIR>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.hp));
    return UKBB_OK;
}
template <int CC, int COUT, bool SPLIT>
static int launch_halo(const TcLayerPlan& P, int fp16, int sms, cudaStream_t st) {
    return fp16 ? launch_halo2<CC, COUT, true, SPLIT>(P, sms, st) : launch_halo2<CC, COUT, false, SPLIT>(P, sms, st);
}

template <int CC, int COUT, int STRIDE, bool F16, bool SPLIT, bool F8 = false, bool PAIR = false>
static int launch_group2(const TcLayerPlan& P, int sms, cudaStream_t st) {
    using Cfg = ConvGroupCfg<CC, COUT, STRIDE, SPLIT, F8, PAIR>;
    static bool attr_set[kMaxDevices] = {};          // cudaFuncSetAttribute is per device (one process may drive several: SplitEngine)
    if (first_use_on_device(attr_set)) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_group_kernel<CC, COUT, STRIDE, F16, SPLIT, F8, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    if (PAIR) {                      // clusters of two CTAs, a pair of tiles per cluster and round
        const int pairs = (P.gp.n_tiles + 1) / 2;
        const int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);
        UKBB_CUDA(launch_pdl<2>(conv_group_kernel<CC, COUT, STRIDE, F16, SPLIT, F8, PAIR>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.map_out, P.gp));
        return UKBB_OK;
    }
    const int grid = P.gp.n_tiles < sms ? P.gp.n_tiles : sms;
    UKBB_CUDA(launch_pdl(conv_group_kernel<CC, COUT, STRIDE, F16, SPLIT, F8, PAIR>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.map_out, P.gp));
    return UKBB_OK;
}
template <int CC, int COUT, int STRIDE, bool SPLIT>
static int launch_group(const TcLayerPlan& P, int fp16, int sms, cudaStream_t st) {
    return fp16 ? launch_group2<CC, COUT, STRIDE, true, SPLIT>(P, sms, st) : launch_group2<CC, COUT, STRIDE, false, SPLIT>(P, sms, st);
}

int launch_plan(const TcLayerPlan& P, int fp16, int sms, cudaStream_t st) {
    if (P.f8) {                     // x2 scheme (FP16 pieces + FP8 corrections): same plans and tile shapes as the x3 instances
        if (P.kind == 2) {
            if (P.pair && P.p.cin == 64 && P.cout == 64 && P.p.stride == 1) return launch_group2<64, 64, 1, true, true, true, true>(P, sms, st);
#define GCASE(CIN, COUTV, S) if (!P.pair && P.p.cin == CIN && P.cout == COUTV && P.p.stride == S) return launch_group2<CIN, COUTV, S, true, true, true>(P, sms, st)
            GCASE(16, 16, 1); GCASE(32, 32, 1); GCASE(16, 32, 2); GCASE(32, 64, 2);
#undef GCASE
        } else if (P.kind == 1) {
            if (P.pair && P.cc == 32 && P.cout == 128) return launch_halo2<32, 128, true, true, true, true>(P, sms, st);
            if (P.pair && P.cc == 32 && P.cout == 256) return launch_halo2<32, 256, true, true, true, true>(P, sms, st);
        } else {
            if (P.cc == 64 && P.cout == 128) return launch_tc2<64, 128, true, true, true>(P, sms, st);
            if (P.cc == 64 && P.cout == 256) return launch_tc2<64, 256, true, true, true>(P, sms, st);
        }
        set_error("x2 scheme: no kernel instance for kind %d, chunk %d, %d -> %d stride %d", P.kind, P.cc, P.p.cin, P.cout, P.p.stride);
        return UKBB_E_UNSUPPORTED;
    }
    if (P.kind == 2) {
#define GCASE(CIN, COUTV, S)                                                                                         \
        if (P.p.cin == CIN && P.cout == COUTV && P.p.stride == S)                                                    \
            return P.split ? launch_group<CIN, COUTV, S, true>(P, fp16, sms, st) : launch_group<CIN, COUTV, S, false>(P, fp16, sms, st)
        GCASE(16, 16, 1); GCASE(32, 32, 1); GCASE(64, 64, 1); GCASE(16, 32, 2); GCASE(32, 64, 2);
#undef GCASE
        set_error("conv_group: no kernel instance for %d -> %d stride %d", P.p.cin, P.cout, P.p.stride);
        return UKBB_E_UNSUPPORTED;
    }
    if (P.kind == 1) {
        if (!P.split && P.cc == 64 && P.cout == 128) return launch_halo<64, 128, false>(P, fp16, sms, st);
        if (!P.split && P.cc == 64 && P.cout == 256) return launch_halo<64, 256, false>(P, fp16, sms, st);
        if (P.split && P.cc == 32 && P.cout == 128) return launch_halo<32, 128, true>(P, fp16, sms, st);
        if (P.split && P.cc == 32 && P.cout == 256) return launch_halo<32, 256, true>(P, fp16, sms, st);
        set_error("conv_halo: no kernel instance for chunk %d x %d, cout %d", P.cc, P.hp.chunks, P.cout);
        return UKBB_E_UNSUPPORTED;
    }
    if (P.split) {
        if (P.cc == 64 && P.cout == 128) return launch_tc_x3<64, 128>(P, fp16, sms, st);
        if (P.cc == 64 && P.cout == 256) return launch_tc_x3<64, 256>(P, fp16, sms, st);
        set_error("conv_tc (split operands): no kernel instance for chunk %d, cout %d", P.cc, P.cout);
        return UKBB_E_UNSUPPORTED;
    }
#define CASE(CCV, COUTV) if (P.cc == CCV && P.cout == COUTV) return launch_tc<CCV, COUTV>(P, fp16, sms, st)
    CASE(16, 32); CASE(32, 32); CASE(32, 64);
    CASE(64, 32); CASE(64, 64); CASE(64, 128); CASE(64, 256);
#undef CASE
    set_error("conv_tc: no kernel instance for chunk %d, cout %d", P.cc, P.cout);
    return UKBB_E_UNSUPPORTED;
}

}  // namespace ukbb
