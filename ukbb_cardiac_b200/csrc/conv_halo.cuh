// 3x3 stride-1 convolution + folded BN + ReLU on tcgen05 with HALO REUSE (north_star (a)).
//
// One tile = 16 x 16 output pixels of one slice = two M=128 UMMA sub-tiles (left / right 8
// columns, 16 rows each).  Per input-channel chunk (CC channels) ONE TMA box load brings the
// 18 x 18 x CC halo patch into shared memory as [py][px][CC] rows (K-major, 32/64/128-byte
// swizzle); the nine filter taps are nine DESCRIPTOR WINDOWS into that patch:
//     start = patch + ((ky*18 + kx + 8*h) * row_bytes),  stride between 8-row groups = 18 rows
// (an 8-row group = 8 horizontally adjacent pixels; the next group is the next image row).
// That is legal because the UMMA swizzle is a pure XOR of address bits (measured:
// profiles/r1_umma_probe.log), so a window may start at any row of a TMA-written patch.
// Activation bytes fetched per output pixel drop from 9x (one load per tap) to 1.27x.
// TMA zero fill outside the image is TF's SAME padding (pad 1 on every side for 3x3 stride 1).
//
// Weights ([Cout][tap*Cin] K-major, one [Cout][CC] tile per (chunk, tap)) are either RESIDENT
// in shared memory for the whole kernel (loaded once per CTA: layers up to 64->64) or STREAMED
// through a ring (128- and 256-channel layers).  Accumulators: TMEM, 2 sub-tiles x Cout columns
// per stage, two stages when they fit (Cout <= 128) so the epilogue overlaps the next tile.
#pragma once
#include "tc_common.cuh"

namespace ukbb {

struct ConvHaloParams {
    int cin, chunks;
    int tiles_x, tiles_y, n_tiles;
    int ho, wo, n;
    int relu, fp16;
    const float* scale;
    const float* shift;
    __nv_bfloat16* out;             // [n][ho][wo][COUT]
};

template <int CC, int COUT, bool RESIDENT, int NKB /* 9 * chunks when RESIDENT */>
struct ConvHaloCfg {
    static constexpr int RB = CC * 2;                                   // bytes per patch row (pixel)
    static constexpr int PATCH_BYTES = (324 * RB + 1023) / 1024 * 1024;
    static constexpr int B_TILE = (COUT * RB + 1023) / 1024 * 1024;
    static constexpr int B_TILES = RESIDENT ? NKB : (COUT >= 256 ? 3 : 4);
    static constexpr int B_BYTES = B_TILES * B_TILE;
    static constexpr int A_MAX = (200 * 1024 - B_BYTES) / PATCH_BYTES;
    static constexpr int A_STAGES = A_MAX > 4 ? 4 : A_MAX;
    static constexpr int ACC_STAGES = COUT <= 128 ? 2 : 1;
    static constexpr int TMEM_COLS_RAW = ACC_STAGES * 2 * COUT;
    static constexpr int TMEM_COLS = TMEM_COLS_RAW < 32 ? 32 : TMEM_COLS_RAW;  // 64..512, powers of two here
    static constexpr int SMEM_BYTES = A_STAGES * PATCH_BYTES + B_BYTES + 1024 + 512 + 2 * COUT * 4;
    static_assert(A_STAGES >= 2, "need at least two patch stages");
};

template <int CC, int COUT, bool RESIDENT, int NKB>
__global__ void __launch_bounds__(256, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const ConvHaloParams p) {
    using namespace tc;
    using Cfg = ConvHaloCfg<CC, COUT, RESIDENT, NKB>;
    constexpr int RB = Cfg::RB, AST = Cfg::A_STAGES, BST = Cfg::B_TILES, ACC = Cfg::ACC_STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + AST * Cfg::PATCH_BYTES;
    const uint32_t bar_base = b_base + Cfg::B_BYTES;
    // barriers: a_full[AST] a_empty[AST] b_full[BST] b_empty[BST] tfull[2] tempty[2] wfull
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (AST + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (2 * AST + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (2 * AST + 4 + s); };
    auto tfull = [&](int a) { return bar_base + 8u * (2 * AST + 8 + a); };
    auto tempty = [&](int a) { return bar_base + 8u * (2 * AST + 10 + a); };
    const uint32_t wfull = bar_base + 8u * (2 * AST + 12);
    const uint32_t tmem_slot = bar_base + 8u * (2 * AST + 13);
    const uint32_t ss_base = bar_base + 512;                 // scale[COUT], shift[COUT] as float
    const float* s_scale = reinterpret_cast<const float*>(smem_raw + (ss_base - smem_u32(smem_raw)));
    const float* s_shift = s_scale + COUT;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < AST; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    if (warp == 3) {
        float* ss = const_cast<float*>(s_scale);
        for (int c = lane; c < COUT; c += 32) { ss[c] = p.scale[c]; ss[COUT + c] = p.shift[c]; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int tiles_per_slice = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (RESIDENT) {
                mbar_arrive_expect_tx(wfull, NKB * COUT * RB);
                for (int kb = 0; kb < NKB; ++kb) {
                    const int ch = kb / 9, tap = kb - ch * 9;
                    tma_load_2d(b_base + kb * Cfg::B_TILE, &map_b, wfull, tap * p.cin + ch * CC, 0);
                }
            }
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int n = tile / tiles_per_slice, t2 = tile - n * tiles_per_slice;
                const int y0 = (t2 / p.tiles_x) * 16, x0 = (t2 % p.tiles_x) * 16;
                for (int ch = 0; ch < p.chunks; ++ch) {
                    mbar_wait(a_empty(as), aph ^ 1);
                    mbar_arrive_expect_tx(a_full(as), 324 * RB);
                    tma_load_4d(smem_base + as * Cfg::PATCH_BYTES, &map_a, a_full(as), ch * CC, x0 - 1, y0 - 1, n);
                    if (++as == AST) { as = 0; aph ^= 1; }
                    if (!RESIDENT) {
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(b_empty(bs), bph ^ 1);
                            mbar_arrive_expect_tx(b_full(bs), COUT * RB);
                            tma_load_2d(b_base + bs * Cfg::B_TILE, &map_b, b_full(bs), tap * p.cin + ch * CC, 0);
                            if (++bs == BST) { bs = 0; bph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = p.fp16 ? make_idesc_f16(128, COUT) : make_idesc_bf16(128, COUT);
            constexpr uint64_t SBO_MASK = ~(0x3FFFull << 32);
            constexpr uint64_t SBO_PATCH = (uint64_t)((18 * RB) >> 4) << 32;
            if (RESIDENT) { mbar_wait(wfull, 0); tc_fence_after(); }
            int as = 0, bs = 0, acc = 0;
            uint32_t aph = 0, bph = 0, acc_ph = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(tempty(acc), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * (2 * COUT);
                for (int ch = 0; ch < p.chunks; ++ch) {
                    mbar_wait(a_full(as), aph);
                    tc_fence_after();
                    const uint32_t patch = smem_base + as * Cfg::PATCH_BYTES;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        uint32_t b_addr;
                        if (RESIDENT) {
                            b_addr = b_base + (ch * 9 + tap) * Cfg::B_TILE;
                        } else {
                            mbar_wait(b_full(bs), bph);
                            tc_fence_after();
                            b_addr = b_base + bs * Cfg::B_TILE;
                        }
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const uint32_t win = patch + (ky * 18 + kx) * RB;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
#pragma unroll
                            for (int k = 0; k < CC / 16; ++k) {
                                const uint64_t ad = (make_smem_desc(win + h * 8 * RB + k * 32, RB) & SBO_MASK) | SBO_PATCH;
                                const uint64_t bd = make_smem_desc(b_addr + k * 32, RB);
                                umma_bf16(d0 + h * COUT, ad, bd, idesc, (ch | tap | k) != 0 ? 1u : 0u);
                            }
                        }
                        if (!RESIDENT) {
                            umma_commit(b_empty(bs));
                            if (++bs == BST) { bs = 0; bph ^= 1; }
                        }
                    }
                    umma_commit(a_empty(as));
                    if (++as == AST) { as = 0; aph ^= 1; }
                }
                umma_commit(tfull(acc));
                if (++acc == ACC) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;
        const int r = q * 32 + lane;
        const int ty = r >> 3, txl = r & 7;
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int n = tile / tiles_per_slice, t2 = tile - n * tiles_per_slice;
            const int oy = (t2 / p.tiles_x) * 16 + ty, x0 = (t2 % p.tiles_x) * 16;
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int ox = x0 + txl + 8 * h;
                const bool live = oy < p.ho && ox < p.wo;
                __nv_bfloat16* dst = p.out + (((size_t)n * p.ho + oy) * p.wo + ox) * COUT;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (2 * COUT) + h * COUT;
#pragma unroll 1
                for (int c = 0; c < COUT; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
                    tmem_ld_wait();
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float a = fmaf(__uint_as_float(v[2 * j]), s_scale[c + 2 * j], s_shift[c + 2 * j]);
                        float b = fmaf(__uint_as_float(v[2 * j + 1]), s_scale[c + 2 * j + 1], s_shift[c + 2 * j + 1]);
                        if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        o[j] = pack16(a, b, p.fp16);
                    }
                    if (live) {
                        uint4* d4 = reinterpret_cast<uint4*>(dst + c);
                        d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (++acc == ACC) { acc = 0; acc_ph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace ukbb
