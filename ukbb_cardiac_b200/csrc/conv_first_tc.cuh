// conv0_0 + conv0_1 in ONE launch, conv0_0 ON THE TENSOR PIPE (north_star (a); network.py:19-25, :184-188).
//
// conv_first.cuh computes conv0_0 (Cin = 1, K = 9) in FP32 on the CUDA cores: 103.7 kFMA per 512-pixel tile, which
// bounds that kernel at ~2800 cycles per tile (the FP32 pipe gives 115 FMA/clk/SM, experiments/ffma_probe.cu) while its
// 18 UMMAs need 864.  Here conv0_0 is a GEMM too, without giving up the FP32 input:
//     one patch GROUP ROW = 4 horizontally adjacent pixels x 16 channels (the 128-byte row conv0_1 reads) needs the
//     3 x 6 image window around it: 18 FP32 values x_v.  Each is split x = hi + lo (two 16-bit values, x - hi is exact in
//     FP32), each weight (BN scale folded in) likewise w = w_hi + w_lo, and
//         sum_v x_v w_v  ~=  sum_v hi_v w_hi + lo_v w_hi + hi_v w_lo            (error ~2^-16 relative: x_lo w_lo is dropped)
//     is ONE GEMM row: A0[group][K = 64] = [hi(18) | lo(18) | hi(18) | 0(10)],  B0[N = 64 = 4 px x 16 ch][K = 64] holds
//     w_hi / w_hi / w_lo at (pixel s, column c) -> tap kx = c - s, zero where that tap does not exist.
//     D0[group][4 px x 16 ch] (+ shift, ReLU, 16 bit) IS the 128-byte patch row, so the epilogue of this stage writes it
//     with eight swizzled STS.128 exactly where the TMA box load of conv_group would have put it.
// 180 group rows per tile = two M = 128 blocks (the second 52 rows), 2 x 4 UMMAs (N = 64, K = 16) with A0 in TENSOR
// MEMORY (32 cycles each, profiles/r1_ts_probe.log).  CUDA-core work per group row drops from ~430 to ~180 instructions.
//   warp 0        TMA: FP32 image boxes (20 x 48, zero fill = SAME padding of conv0_0), weights once
//   warp 1        MMA issuer conv0_1 (18 UMMAs per tile, as conv_group<16,16,1>)
//   warp 2        TMEM allocator, MMA issuer conv0_0 (8 UMMAs per tile, A from TMEM)
//   warps 4-7     epilogue of conv0_1 (folded BN + ReLU, 16-bit pack, staging tile, TMA store)
//   warps 8-13    builders, one thread per group row (warps 8-11: rows 0-127, warps 12-13: rows 128-179):
//                 build A0(i + 1) [18 LDS, hi/lo split, tcgen05.st], then finish tile i [tcgen05.ld D0, shift + ReLU + pack,
//                 8 STS.128; groups outside the image are written as ZEROS: SAME padding of conv0_1 pads a0, not the image].
//                 (Separate builder / finisher warps, 22 warps at 80 registers, spilled and measured slower: 343 vs 266 us.)
// TMEM columns: conv0_1 accumulators 2 x 64 | D0 2 stages x 2 blocks x 64 | A0 2 stages x 2 blocks x 32 = 512.
//
// SPLIT (x3 modes): the a0 patch is written as a hi and a lo patch (a0 = hi + lo), conv0_1 runs three UMMAs per K-slice against the
// hi and lo weight sets, and b0 leaves as two planes through direct 256-bit stores (as conv_group.cuh).  conv0_0 itself is unchanged:
// its hi/lo split of the FP32 image and weights is what the x3 modes do everywhere else.
#pragma once
#include "tc_common.cuh"
#include "conv_group.cuh"
#include "head_common.cuh"     // add_relu_pack / add_relu_split

namespace ukbb {

struct ConvFirstParams {
    int tiles_x, tiles_y, n_tiles;
    int h, w4;                      // image rows; image columns / 4 (pixel groups per row)
    const float* scale;             // conv0_1 folded BN [16]
    const float* shift;
    float shift0[16];               // conv0_0 folded BN shift (the scale is folded into its weights)
    uint32_t* out;                  // split modes: b0 hi plane [n][h][w4 groups][64 elements] as 32-bit words
    long long out_lo;               //              offset of the lo plane in 32-bit words
    int lo_n;                       //              slice index of the lo plane in the output tensor map
};

namespace tc {
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
}  // namespace tc

template <bool SPLIT = false, bool F8 = false>
struct ConvFirstTcCfg {
    static constexpr bool MERGE = SPLIT && !F8;                 // x3: N = 128 merge of the hi.hi / hi.lo products; x2 (F8): FP8 correction pass
    static constexpr int PU = 10, PR = 18, J = 6, N = 64;
    static constexpr int GROUPS = PU * PR;                      // 180 patch group rows per tile
    static constexpr int BUILDERS = 6;                          // builder warps with work (8 launched: TMEM lane quarters)
    static constexpr int PLANE_BYTES = (GROUPS * 128 + 1023) / 1024 * 1024;
    static constexpr int PATCH_BYTES = (SPLIT ? 2 : 1) * PLANE_BYTES;     // hi patch | lo patch
    static constexpr int A_STAGES = SPLIT ? 2 : 4;
    static constexpr int B_TILE = 2048;
    static constexpr int NB_TILES = 3 * J;
    static constexpr int B_SET = NB_TILES * B_TILE;
    static constexpr int B_BYTES = (SPLIT ? 2 : 1) * B_SET;     // hi tiles | lo tiles
    static constexpr int B0_BYTES = 64 * 128;                   // conv0_0 weights [64 rows][64 K] 16-bit, 128 B swizzle
    static constexpr int OUT_BYTES = 128 * 128;                // staging tiles of the TMA store: two tiles (plain) / hi + lo tile of one (SPLIT)
    static constexpr int IMG_W = 48, IMG_H = 20;                // FP32 box: columns x0 - 8 .. x0 + 39, rows y0 - 2 .. y0 + 17 (TMA needs the
                                                                // box origin 16-byte aligned in the inner dimension: experiments/tma_probe_img.cu)
    static constexpr int IMG_TX = IMG_W * IMG_H * 4;
    static constexpr int IMG_BYTES = (IMG_TX + 127) / 128 * 128;
    static constexpr int IMG_STAGES = 4;
    static constexpr int THREADS = 512;
    static constexpr int TMEM_COLS = 512;
    // TMEM columns.  plain: conv0_1 accumulators 2 x 64 | D0 2 stages x 2 blocks x 64 | A0 2 stages x 2 blocks x 32.
    // SPLIT: conv0_1 accumulators 2 x 128 (columns [0, 64) = hi.hi + lo.hi, [64, 128) = hi.lo: the N = 128 merge of conv_group.cuh) |
    //        D0 ONE stage x 2 blocks x 64 (the finisher loads all 64 columns and releases it before converting) | A0 as before.
    static constexpr int ACC_COLS = MERGE ? 128 : 64;
    static constexpr int D0_STAGES = MERGE ? 1 : 2;
    static constexpr int COL_ACC = 0, COL_D0 = 2 * ACC_COLS, COL_A0 = 384;
    static_assert(COL_D0 + D0_STAGES * 2 * 64 <= COL_A0, "TMEM budget");
    static constexpr int SMEM_BYTES = A_STAGES * PATCH_BYTES + B_BYTES + B0_BYTES + 2 * OUT_BYTES + IMG_STAGES * IMG_BYTES + 256 /*barriers*/ +
                                      2 * N * 4 + 1024 /*align*/;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <bool F16, bool SPLIT = false, bool F8 = false>
__global__ void __launch_bounds__(ConvFirstTcCfg<>::THREADS, 1)
conv_first_tc_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_b0, const __grid_constant__ CUtensorMap map_out,
                     const __grid_constant__ ConvFirstParams p) {
    using namespace tc;
    using Cfg = ConvFirstTcCfg<SPLIT, F8>;
    constexpr bool MERGE = Cfg::MERGE;
    constexpr int AST = Cfg::A_STAGES, IST = Cfg::IMG_STAGES, J = Cfg::J, PU = Cfg::PU, N = Cfg::N;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t b_base = smem_base + AST * Cfg::PATCH_BYTES;
    const uint32_t b0_base = b_base + Cfg::B_BYTES;
    const uint32_t out_base = b0_base + Cfg::B0_BYTES;
    const uint32_t img_base = out_base + 2 * Cfg::OUT_BYTES;
    const uint32_t bar_base = img_base + IST * Cfg::IMG_BYTES;
    auto BAR = [&](int i) { return bar_base + 8u * i; };
    enum { A_FULL = 0, A_EMPTY = A_FULL + AST, IMG_FULL = A_EMPTY + AST, IMG_EMPTY = IMG_FULL + IST, A0_FULL = IMG_EMPTY + IST, A0_EMPTY = A0_FULL + 2,
           D0_FULL = A0_EMPTY + 2, D0_EMPTY = D0_FULL + 2, TFULL = D0_EMPTY + 2, TEMPTY = TFULL + 2, WFULL = TEMPTY + 2, TSLOT = WFULL + 1 };
    static_assert((TSLOT + 1) * 8 <= 256, "barrier area");
    auto a_full = [&](int s) { return BAR(A_FULL + s); };
    auto a_empty = [&](int s) { return BAR(A_EMPTY + s); };
    auto tfull = [&](int a) { return BAR(TFULL + a); };
    auto tempty = [&](int a) { return BAR(TEMPTY + a); };
    const uint32_t wfull = BAR(WFULL);
    const uint32_t tmem_slot = BAR(TSLOT);
    float* s_scale = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 256);        // [N] expanded (column -> channel)
    float* s_shift = s_scale + N;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_img); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_b0); tma_prefetch_desc(&map_out); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < AST; ++s) { mbar_init(a_full(s), Cfg::BUILDERS); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < IST; ++s) { mbar_init(BAR(IMG_FULL + s), 1); mbar_init(BAR(IMG_EMPTY + s), Cfg::BUILDERS); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(BAR(A0_FULL + a), Cfg::BUILDERS); mbar_init(BAR(A0_EMPTY + a), 1);
            mbar_init(BAR(D0_FULL + a), 1); mbar_init(BAR(D0_EMPTY + a), Cfg::BUILDERS);
            mbar_init(tfull(a), 1); mbar_init(tempty(a), 4);
        }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    if (warp == 3) {
        for (int c = lane; c < N; c += 32) { s_scale[c] = p.scale[c & 15]; s_shift[c] = p.shift[c & 15]; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 0) griddep_wait();                       // the producer waits after it has issued the weight loads
    const int my_tiles = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0) {
        // ===================== TMA producer: weights once, one FP32 image box per tile =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(wfull, (SPLIT ? 2 : 1) * Cfg::NB_TILES * Cfg::B_TILE + Cfg::B0_BYTES);
            for (int t = 0; t < (SPLIT ? 2 : 1) * Cfg::NB_TILES; ++t) {
                // global: hi tiles, then lo tiles.  Shared (SPLIT): tile t = [hi 64 rows | lo 64 rows], one N = 128 B operand
                const int tt = t % Cfg::NB_TILES, pl = t / Cfg::NB_TILES;
                tma_load_2d(b_base + (MERGE ? tt * 2 * Cfg::B_TILE + pl * Cfg::B_TILE : t * Cfg::B_TILE), &map_b, wfull, 0, t * N);
            }
            tma_load_2d(b0_base, &map_b0, wfull, 0, 0);
            griddep_wait();
            TileWalk w;
            w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
            int is = 0;
            uint32_t iph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                mbar_wait(BAR(IMG_EMPTY + is), iph ^ 1);
                mbar_arrive_expect_tx(BAR(IMG_FULL + is), Cfg::IMG_TX);
                tma_load_3d(img_base + is * Cfg::IMG_BYTES, &map_img, BAR(IMG_FULL + is), w.tx * 32 - 8, w.ty * 16 - 2, w.n);
                if (++is == IST) { is = 0; iph ^= 1; }
                w.next();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (conv0_1, as conv_group<16, 16, 1>) =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, N) : make_idesc_bf16(128, N);
        const uint32_t idesc2 = F16 ? make_idesc_f16(128, 2 * N) : make_idesc_bf16(128, 2 * N);    // x3: A_hi . [w_hi ; w_lo]^T
        const uint32_t idesc8 = make_idesc_e4m3(128, N);                                           // x2: FP8 correction pass
        constexpr uint32_t a_hi = (uint32_t)((PU * 128) >> 4) | (1u << 14) | (2u << 29);       // 8-row groups one patch row apart, 128 B swizzle
        constexpr uint32_t b_hi = (uint32_t)((8 * 32) >> 4) | (1u << 14) | (6u << 29);         // weights: 32-byte rows, 32 B swizzle
        const uint32_t b_lo = ((b_base & 0x3FFFF) >> 4) | (1u << 16);
        mbar_wait(wfull, 0);
        tc_fence_after();
        int as = 0, acc = 0;
        uint32_t aph = 0, acc_ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(tempty(acc), acc_ph ^ 1);
            mbar_wait(a_full(as), aph);
            tc_fence_after();
            const uint32_t d = tmem_base + Cfg::COL_ACC + acc * Cfg::ACC_COLS;
            const uint32_t a_lo = (((smem_base + as * Cfg::PATCH_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
            if (leader) {
#pragma unroll
                for (int pass = F8 ? 0 : 1; pass < 2; ++pass)                    // x2: pass 0 = FP8 corrections, pass 1 = FP16 main term
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int jj = j - 1;                                    // input pixel of this K-slice relative to the group
                        const int ro = jj < 0 ? -1 : jj / 4;
                        const int sub = jj - ro * 4;
                        const int arow = ky * PU + 1 + ro;
                        const uint32_t ao = (uint32_t)((arow * 128 + sub * 32) >> 4), bo = (uint32_t)(((ky * J + j) * (MERGE ? 2 : 1) * Cfg::B_TILE) >> 4);
                        if (F8) {
                            if (pass == 0) umma_f8_lohi(d, a_lo + (Cfg::PLANE_BYTES >> 4) + ao, a_hi, b_lo + (Cfg::B_SET >> 4) + bo, b_hi, idesc8, (ky | j) != 0 ? 1u : 0u);
                            else if ((ky | j) == 0) umma_f16_lohi_rescale(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc);
                            else umma_bf16_lohi(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc, 1u);
                        } else if (SPLIT) {
                            umma_bf16_lohi(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc2, (ky | j) != 0 ? 1u : 0u);                    // hi . [hi ; lo]
                            umma_bf16_lohi(d, a_lo + (Cfg::PLANE_BYTES >> 4) + ao, a_hi, b_lo + bo, b_hi, idesc, 1u);                // lo . hi
                        } else {
                            umma_bf16_lohi(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc, (ky | j) != 0 ? 1u : 0u);
                        }
                    }
                umma_commit(a_empty(as));
                umma_commit(tfull(acc));
            }
            __syncwarp();
            if (++as == AST) { as = 0; aph ^= 1; }
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    } else if (warp == 2) {
        // ===================== MMA issuer (conv0_0): D0[s][m] = A0[s][m] (TMEM) . B0^T, m = 0, 1 =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, N) : make_idesc_bf16(128, N);
        constexpr uint32_t b_hi = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t b_lo = ((b0_base & 0x3FFFF) >> 4) | (1u << 16);
        mbar_wait(wfull, 0);
        tc_fence_after();
        for (int i = 0; i < my_tiles; ++i) {
            const int s = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            const int ds = MERGE ? 0 : s;                                            // D0 stage (x3: one stage, phase = tile parity)
            const uint32_t dph = MERGE ? ((uint32_t)i & 1u) : ph;
            mbar_wait(BAR(D0_EMPTY + ds), dph ^ 1);
            mbar_wait(BAR(A0_FULL + s), ph);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_ts_lohi(tmem_base + Cfg::COL_D0 + (ds * 2 + m) * 64, tmem_base + Cfg::COL_A0 + (s * 2 + m) * 32 + 8 * k, b_lo + 2 * k, b_hi,
                                     idesc, k != 0 ? 1u : 0u);
                umma_commit(BAR(A0_EMPTY + s));
                umma_commit(BAR(D0_FULL + ds));
            }
            __syncwarp();
        }
    } else if (warp >= 8 && warp < 14) {
        // ===================== builders: one thread per patch group row =====================
        const int bw = warp - 8;
        const int m = bw >> 2, q = bw & 3;                                       // UMMA row block, TMEM lane quarter (= warp % 4)
        const int g = m * 128 + q * 32 + lane;                                   // patch group row = patch row py, group pg
        const bool active = g < Cfg::GROUPS;
        const int py = g / PU, pg = g - py * PU;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        TileWalk w;
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        float sh0[16];                                                           // conv0_0 shifts: registers, not one LDC per use
#pragma unroll
        for (int c = 0; c < 16; ++c) sh0[c] = p.shift0[c];
        auto build = [&](int j) {                                                // A0 of tile j -> TMEM stage j & 1
            const int s = j & 1, is = j % IST;
            mbar_wait(BAR(IMG_FULL + is), (uint32_t)(j / IST) & 1u);
            float v[18];
            if (active) {
                // group pg needs image columns x0 - 5 + 4 pg .. + 5 = box columns 4 pg + 3 .. 4 pg + 8, rows py .. py + 2
                const float* src = reinterpret_cast<const float*>(smem_gen + (img_base - smem_base) + is * Cfg::IMG_BYTES) + py * Cfg::IMG_W + 4 * pg;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4 a = *reinterpret_cast<const float4*>(src + r * Cfg::IMG_W + 4);
                    v[6 * r] = src[r * Cfg::IMG_W + 3];
                    v[6 * r + 1] = a.x; v[6 * r + 2] = a.y; v[6 * r + 3] = a.z; v[6 * r + 4] = a.w;
                    v[6 * r + 5] = src[r * Cfg::IMG_W + 8];
                }
            } else {
#pragma unroll
                for (int c = 0; c < 18; ++c) v[c] = 0.f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(IMG_EMPTY + is));                     // image values are in registers
            uint32_t a[32];
#pragma unroll
            for (int pr = 0; pr < 9; ++pr) {
                const float x0 = v[2 * pr], x1 = v[2 * pr + 1];
                const uint32_t hi = pack16t<F16>(x0, x1);
                const float2 hf = unpack16t<F16>(hi);
                const uint32_t lo = pack16t<F16>(x0 - hf.x, x1 - hf.y);          // x - hi is exact in FP32
                a[pr] = hi; a[9 + pr] = lo; a[18 + pr] = hi;
            }
#pragma unroll
            for (int c = 27; c < 32; ++c) a[c] = 0u;
            mbar_wait(BAR(A0_EMPTY + s), (((uint32_t)j >> 1) & 1u) ^ 1u);
            tc_fence_after();
            tmem_st32(lane_base + Cfg::COL_A0 + (s * 2 + m) * 32, a);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A0_FULL + s));
        };
        if (my_tiles > 0) build(0);
        for (int i = 0; i < my_tiles; ++i) {
            if (i + 1 < my_tiles) build(i + 1);
            // ---- finish tile i: D0 -> shift + ReLU -> 16 bit -> the 128-byte patch row of conv0_1
            const int s = i & 1, as = i % AST;
            if (!MERGE) {
                mbar_wait(BAR(D0_FULL + s), ((uint32_t)i >> 1) & 1u);
                tc_fence_after();
            }
            // SAME padding of conv0_1 pads a0 with zeros: groups outside the image are zero, not conv0_0 of a zero image
            const int y = w.ty * 16 - 1 + py, gx = w.tx * 8 - 1 + pg;
            const bool inside = y >= 0 && y < p.h && gx >= 0 && gx < p.w4;
            if (SPLIT) {
                // D0 has ONE stage here: load all 64 accumulator columns, release it, then convert (one pixel = 16 columns at a time)
                // into the hi and lo rows of the patch, two swizzled STS.128 each
                uint32_t d[64];
                if (MERGE) {
                    mbar_wait(BAR(D0_FULL), (uint32_t)i & 1u);
                    tc_fence_after();
                }
                const int dstage = MERGE ? 0 : s;
                tmem_ld32(lane_base + Cfg::COL_D0 + (dstage * 2 + m) * 64, d);
                tmem_ld32(lane_base + Cfg::COL_D0 + (dstage * 2 + m) * 64 + 32, d + 32);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(D0_EMPTY + dstage));
                mbar_wait(a_empty(as), ((uint32_t)(i / AST) & 1u) ^ 1u);
                const uint32_t row = smem_base + as * Cfg::PATCH_BYTES + (uint32_t)g * 128u;
#pragma unroll
                for (int qt = 0; qt < 4; ++qt) {
                    uint32_t oh[8], ol[8];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        add_relu_split4<F16, F8>(d + 16 * qt + 4 * c, sh0 + 4 * c, oh, ol, c);
                    if (!inside) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) { oh[c] = 0u; ol[c] = 0u; }
                    }
                    if (active) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const uint32_t dst = row + (((uint32_t)(2 * qt + c) ^ ((uint32_t)g & 7u)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(oh[4 * c]), "r"(oh[4 * c + 1]), "r"(oh[4 * c + 2]),
                                         "r"(oh[4 * c + 3]) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + Cfg::PLANE_BYTES), "r"(ol[4 * c]), "r"(ol[4 * c + 1]),
                                         "r"(ol[4 * c + 2]), "r"(ol[4 * c + 3]) : "memory");
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full(as));
                w.next();
                continue;
            }
            uint32_t o[32];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t d[32];
                tmem_ld32(lane_base + Cfg::COL_D0 + (s * 2 + m) * 64 + 32 * hf, d);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    o[16 * hf + c] = add_relu_pack<F16>(d[2 * c], d[2 * c + 1], p.shift0[(2 * c) & 15], p.shift0[(2 * c + 1) & 15]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(D0_EMPTY + s));
            mbar_wait(a_empty(as), ((uint32_t)(i / AST) & 1u) ^ 1u);
            if (active) {
                const uint32_t row = smem_base + as * Cfg::PATCH_BYTES + (uint32_t)g * 128u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t o0 = o[4 * c], o1 = o[4 * c + 1], o2 = o[4 * c + 2], o3 = o[4 * c + 3];
                    if (!inside) { o0 = 0u; o1 = 0u; o2 = 0u; o3 = 0u; }
                    const uint32_t dst = row + (((uint32_t)c ^ ((uint32_t)g & 7u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(as));
            w.next();
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== epilogue (conv0_1): as conv_group =====================
        const int q = warp - 4;
        const int r = q * 32 + lane;                         // TMEM lane = tile row * 8 + group
        const bool issuer = threadIdx.x == 128;
        TileWalk w;
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        int acc = 0;
        uint32_t acc_ph = 0;
        float4 scv[4], shv[4];                               // conv0_1 folded BN of the 16 channels (column c -> channel c & 15)
#pragma unroll
        for (int c = 0; c < 4; ++c) { scv[c] = *reinterpret_cast<const float4*>(s_scale + 4 * c); shv[c] = *reinterpret_cast<const float4*>(s_shift + 4 * c); }
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::COL_ACC + acc * Cfg::ACC_COLS;
            if (SPLIT) {
                // The tile leaves through a swizzled staging tile per plane and two TMA stores: a thread owns one 128-byte row, and 32-byte
                // stores from 32 threads to 32 different lines cost 32 LSU wavefronts each (the LSU data pipe was 82 % busy, the bound of
                // this kernel: profiles/r2_ncu_summary.txt); STS.128 are conflict-free (4 wavefronts) and the async proxy does the rest.
                const uint32_t row = out_base + r * 128;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint32_t v[16], v2[16], oh[8], ol[8];
                    tmem_ld16(taddr + 16 * c8, v);
                    if (MERGE) tmem_ld16(taddr + 64 + 16 * c8, v2);
                    tmem_ld_wait();
                    if (c8 == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty(acc));
                    }
                    if (MERGE) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(v2[c]));
                    }
#pragma unroll
                    for (int c = 0; c < 16; c += 4) bn_relu_split4<F16, F8>(v + c, scv[c / 4], shv[c / 4], oh, ol, c / 4);
                    if (c8 == 0) {                                   // the stores of the previous tile have read the staging tiles
                        if (issuer) bulk_wait_read<0>();
                        named_bar_sync(1, 128);
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint32_t dst = row + ((uint32_t)((2 * c8 + k) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(oh[4 * k]), "r"(oh[4 * k + 1]), "r"(oh[4 * k + 2]),
                                     "r"(oh[4 * k + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + Cfg::OUT_BYTES), "r"(ol[4 * k]), "r"(ol[4 * k + 1]),
                                     "r"(ol[4 * k + 2]), "r"(ol[4 * k + 3]) : "memory");
                    }
                }
                fence_proxy_async();
                named_bar_sync(1, 128);
                if (issuer) {
                    tma_store_4d(&map_out, out_base, 0, w.tx * 8, w.ty * 16, w.n);
                    tma_store_4d(&map_out, out_base + Cfg::OUT_BYTES, 0, w.tx * 8, w.ty * 16, w.n + p.lo_n);
                    bulk_commit();
                }
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                w.next();
                continue;
            }
            uint32_t v[64];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            uint32_t o[32];
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
                const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
                o[c / 2] = bn_relu_pack<F16>(v[c], v[c + 1], make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                o[c / 2 + 1] = bn_relu_pack<F16>(v[c + 2], v[c + 3], make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
            }
            const uint32_t row = out_base + (i & 1) * Cfg::OUT_BYTES + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t dst = row + ((uint32_t)(j ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[4 * j]), "r"(o[4 * j + 1]), "r"(o[4 * j + 2]),
                             "r"(o[4 * j + 3])
                             : "memory");
            }
            fence_proxy_async();
            if (issuer) bulk_wait_read<0>();
            named_bar_sync(1, 128);
            if (issuer) {
                tma_store_4d(&map_out, out_base + (i & 1) * Cfg::OUT_BYTES, 0, w.tx * 8, w.ty * 16, w.n);
                bulk_commit();
            }
            w.next();
        }
        if (issuer) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace ukbb
