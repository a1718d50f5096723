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
#include "conv_group.cuh"      // tmem_ld32, bn_relu_pack

namespace ukbb {

struct ConvHaloParams {
    int cin, chunks;
    int tiles_x, tiles_y, n_tiles;
    int ho, wo, n;
    int relu, fp16;
    int lo_n;                       // split mode: slice index of the lo plane in the input tensor map
    long long out_lo;               // split mode: element offset of the lo plane of `out`
    const float* scale;
    const float* shift;
    __nv_bfloat16* out;             // [n][ho][wo][COUT]
};

// SPLIT (x3 modes, streamed weights only): a patch stage holds the hi and the lo patch of a chunk, and the weight stream alternates
// hi / lo tiles: per (chunk, tap) the issuer takes the hi tile (UMMAs A_hi.B_hi and A_lo.B_hi), then the lo tile (A_hi.B_lo).  The
// split instances use 32-channel chunks (64-byte rows) so that two (hi, lo) patch stages and an 8-deep weight ring still fit.
// F8 (x2 scheme, tc_common.cuh): the chunks of a tile are walked TWICE -- lo planes first (patch + weight tiles holding FP8 correction
// operands, kind::f8f6f4), then the hi planes (kind::f16; the first instruction of each sub-tile rescales its accumulator) -- so a patch stage
// holds one plane and the weight stream is one tile per (pass, chunk, tap).
// PAIR (x2 scheme only): the two CTAs of a cluster execute every UMMA together (tcgen05 cta_group::2, M = 256 = sub-tile h of both CTAs'
// tiles): each CTA keeps only ITS half of the Cout rows of a weight tile, so an instruction reads 4 KB of A + Cout/2 * 32 B of B per
// CTA instead of 4 KB + Cout * 32 B -- the shared-memory operand port, not the tensor pipe, bounds these layers.
// The leader CTA issues all UMMAs and commits; the loads of both CTAs signal the leader's barriers; the peer's MMA warp is idle.
template <int CC, int COUT, bool RESIDENT, int NKB /* 9 * chunks when RESIDENT */, bool SPLIT = false, bool F8 = false, bool PAIR = false>
struct ConvHaloCfg {
    static_assert(!SPLIT || !RESIDENT, "the split variant streams its weights");
    static_assert(!F8 || SPLIT, "the FP8 correction scheme is a split-operand scheme");
    static_assert(!PAIR || F8, "the CTA-pair variant exists for the x2 scheme");
    static constexpr int RB = CC * 2;                                   // bytes per patch row (pixel)
    static constexpr int PLANE_BYTES = (324 * RB + 1023) / 1024 * 1024;
    static constexpr int PATCH_BYTES = ((SPLIT && !F8) ? 2 : 1) * PLANE_BYTES;
    static constexpr int B_ROWS = PAIR ? COUT / 2 : COUT;                // weight rows held by one CTA
    static constexpr int B_TILE = (B_ROWS * RB + 1023) / 1024 * 1024;
    // streamed weights: the ring must cover the ~2000-cycle latency of an L2 fetch -- with four 16 KB stages (512 cycles of UMMAs
    // each) the tensor pipe waited for weights half of the time (7350 cycles per chunk for 3456 cycles of UMMAs); two patch stages
    // leave room for eight weight stages (Cout = 128) / four 32 KB stages (Cout = 256)
    static constexpr int A_STREAM = PAIR ? 4 : 2;
    static constexpr int B_FIT = (222 * 1024 - A_STREAM * PATCH_BYTES) / B_TILE;
    static constexpr int B_TILES = RESIDENT ? NKB : (B_FIT > 8 ? 8 : B_FIT);
    static constexpr int B_BYTES = B_TILES * B_TILE;
    static constexpr int A_MAX = (200 * 1024 - B_BYTES) / PATCH_BYTES;
    static constexpr int A_STAGES = RESIDENT ? (A_MAX > 4 ? 4 : A_MAX) : A_STREAM;
    static constexpr int ACC_STAGES = COUT <= 128 ? 2 : 1;
    static constexpr int TMEM_COLS_RAW = ACC_STAGES * 2 * COUT;
    static constexpr int TMEM_COLS = TMEM_COLS_RAW < 32 ? 32 : TMEM_COLS_RAW;  // 64..512, powers of two here
    static constexpr int SMEM_BYTES = A_STAGES * PATCH_BYTES + B_BYTES + 1024 + 512 + 2 * COUT * 4;
    static_assert(A_STAGES >= 2, "need at least two patch stages");
    static_assert(RESIDENT || B_TILES >= 3, "weight ring too shallow");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// CL = true (streamed weights only): the kernel runs in clusters of two CTAs that walk the same (chunk, tap) weight sequence on two
// different tiles.  Every weight tile is fetched from L2 ONCE per cluster -- the CTA whose rank equals the tile's parity issues a TMA
// multicast that lands at the same shared-memory offset in both CTAs and signals both b_full barriers -- and a ring slot is refilled
// only after BOTH tensor pipes have consumed it (tcgen05.commit multicast on both b_empty barriers, count 2).  The level-3/4 layers
// stream 0.29 / 1.18 MB of weights per 16x16-pixel tile and ran at the L2 roofline (6.7 TB/s of L2 reads).
template <int CC, int COUT, bool RESIDENT, int NKB, bool F16, bool CL = false, bool SPLIT = false, bool F8 = false, bool PAIR = false>
__global__ void __launch_bounds__(256, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const ConvHaloParams p) {
    using namespace tc;
    using Cfg = ConvHaloCfg<CC, COUT, RESIDENT, NKB, SPLIT, F8, PAIR>;
    static_assert(!PAIR || CL, "CTA pairs are clusters of two");
    constexpr bool X3 = SPLIT && !F8;
    const int n_chunk_steps = F8 ? 2 * p.chunks : p.chunks;     // x2: [0, chunks) = lo planes, [chunks, 2 chunks) = hi planes
    constexpr int RB = Cfg::RB, AST = Cfg::A_STAGES, BST = Cfg::B_TILES, ACC = Cfg::ACC_STAGES;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + AST * Cfg::PATCH_BYTES;
    const uint32_t bar_base = b_base + Cfg::B_BYTES;
    // barriers: a_full[AST] a_empty[AST] b_full[8] b_empty[8] tfull[2] tempty[2] wfull
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (AST + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (2 * AST + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (2 * AST + 8 + s); };
    auto tfull = [&](int a) { return bar_base + 8u * (2 * AST + 16 + a); };
    auto tempty = [&](int a) { return bar_base + 8u * (2 * AST + 18 + a); };
    const uint32_t wfull = bar_base + 8u * (2 * AST + 20);
    const uint32_t tmem_slot = bar_base + 8u * (2 * AST + 21);
    const uint32_t ss_base = bar_base + 512;                 // scale[COUT], shift[COUT] as float
    const float* s_scale = reinterpret_cast<const float*>(smem_raw + (ss_base - smem_u32(smem_raw)));
    const float* s_shift = s_scale + COUT;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < AST; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < 8; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), (CL && !PAIR) ? 2 : 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), PAIR ? 8 : 4); }   // PAIR: the leader's barrier, both epilogues
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    if (warp == 3) {
        float* ss = const_cast<float*>(s_scale);
        for (int c = lane; c < COUT; c += 32) { ss[c] = p.scale[c]; ss[COUT + c] = p.shift[c]; }
    }
    tc_fence_before();
    __syncthreads();
    // both CTAs' barriers are initialised before any remote arrive / multicast; PAIR: the cta_group::2 allocation is collective and writes
    // the tensor-memory address into BOTH CTAs' slots, so the slot is read after the cluster barrier, not just the CTA barrier
    if (CL) cluster_sync();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 0) griddep_wait();                       // the producer waits after it has issued the weight loads
    // Tile order: x-tile index slowest (tile = (tx * n + slice) * tiles_y + ty).  The two CTAs of a cluster walk the weight stream in
    // lockstep on tiles 2c and 2c + 1; a border tile whose right sub-tile lies outside the image (24 columns = 1.5 tiles at level 3) has
    // half the UMMAs, and with the x-tile index fastest every cluster paired a full tile with a half tile (rank 1 idled a quarter of the
    // time).  Now the pair almost always holds two tiles of the same kind.
    const int tiles_per_tx = p.tiles_y * p.n;
    // tile walk: plain round-robin, or (CL) pairs of tiles dealt to clusters; the odd CTA of the last pair may get a tile beyond
    // n_tiles: it still walks the weight stream (TMA zero-fills its patch, its epilogue stores nothing)
    const uint32_t crank = CL ? cluster_ctarank() : 0u;
    const int tile_first = CL ? 2 * ((int)blockIdx.x >> 1) + (int)crank : (int)blockIdx.x;
    const int tile_step = CL ? ((int)gridDim.x >> 1) * 2 : (int)gridDim.x;
    const int tile_end = CL ? ((p.n_tiles + 1) >> 1) * 2 : p.n_tiles;

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
            griddep_wait();
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            int bt = 0;                                  // running weight-tile index (same sequence in both CTAs of a cluster)
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                const int txi = tile / tiles_per_tx, t2 = tile - txi * tiles_per_tx;
                const int n = t2 / p.tiles_y, y0 = (t2 - n * p.tiles_y) * 16, x0 = txi * 16;
                for (int cs = 0; cs < n_chunk_steps; ++cs) {
                    const int ch = (F8 && cs >= p.chunks) ? cs - p.chunks : cs;
                    const bool lo_pass = F8 && cs < p.chunks;
                    mbar_wait(a_empty(as), aph ^ 1);
                    if (PAIR) {                              // both patches complete on the leader's barrier
                        if (crank == 0) mbar_arrive_expect_tx(a_full(as), 2 * 324 * RB);
                        tma_load_4d_pair(smem_base + as * Cfg::PATCH_BYTES, &map_a, mapa_rank(a_full(as), 0), ch * CC, x0 - 1, y0 - 1, (lo_pass ? p.lo_n : 0) + n);
                    } else {
                    mbar_arrive_expect_tx(a_full(as), (X3 ? 2 : 1) * 324 * RB);
                    tma_load_4d(smem_base + as * Cfg::PATCH_BYTES, &map_a, a_full(as), ch * CC, x0 - 1, y0 - 1, (lo_pass ? p.lo_n : 0) + n);
                    }
                    if (X3) tma_load_4d(smem_base + as * Cfg::PATCH_BYTES + Cfg::PLANE_BYTES, &map_a, a_full(as), ch * CC, x0 - 1, y0 - 1, p.lo_n + n);
                    if (++as == AST) { as = 0; aph ^= 1; }
                    if (!RESIDENT) {
                        for (int tap = 0; tap < (X3 ? 18 : 9); ++tap) {         // x3: (tap, hi), (tap, lo), ... ; lo tiles are rows [COUT, 2 COUT) of the weight map
                            const int tp = X3 ? tap >> 1 : tap, wrow = X3 ? (tap & 1) * COUT : (lo_pass ? COUT : 0);
                            mbar_wait(b_empty(bs), bph ^ 1);
                            if (PAIR) {                      // this CTA's half of the rows (the weight map's box is Cout / 2 rows)
                                if (crank == 0) mbar_arrive_expect_tx(b_full(bs), COUT * RB);
                                tma_load_2d_pair(b_base + bs * Cfg::B_TILE, &map_b, mapa_rank(b_full(bs), 0), tp * p.cin + ch * CC, wrow + (int)crank * (COUT / 2));
                                ++bt;
                                if (++bs == BST) { bs = 0; bph ^= 1; }
                                continue;
                            }
                            mbar_arrive_expect_tx(b_full(bs), COUT * RB);
                            if (!CL) tma_load_2d(b_base + bs * Cfg::B_TILE, &map_b, b_full(bs), tp * p.cin + ch * CC, wrow);
                            else if ((uint32_t)(bt & 1) == crank) tma_load_2d_mc(b_base + bs * Cfg::B_TILE, &map_b, b_full(bs), tp * p.cin + ch * CC, wrow, (uint16_t)3);
                            ++bt;
                            if (++bs == BST) { bs = 0; bph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs the control flow (uniform registers, no divergence); one elected lane
        // issues the tcgen05.mma / tcgen05.commit instructions.  Descriptors are (lo, hi) pairs: hi is
        // constant per operand kind, lo = base + compile-time offset of the (tap, sub-tile, k) window.
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, COUT) : make_idesc_bf16(128, COUT);
        const uint32_t idesc8 = make_idesc_e4m3(128, COUT);
        const uint32_t layout = RB == 128 ? 2u : RB == 64 ? 4u : 6u;
        const uint32_t a_hi = (uint32_t)((18 * RB) >> 4) | (1u << 14) | (layout << 29);     // SBO = 18 patch rows
        const uint32_t b_hi = (uint32_t)((8 * RB) >> 4) | (1u << 14) | (layout << 29);      // SBO = 8 rows
        if (RESIDENT) { mbar_wait(wfull, 0); tc_fence_after(); }
        int as = 0, bs = 0, acc = 0;
        uint32_t aph = 0, bph = 0, acc_ph = 0;
        if (PAIR) {
            // ---- CTA pair: the leader issues M = 256 UMMAs over its own and the peer's sub-tile h; the peer's MMA warp has nothing to do
            const uint32_t idesc_p = make_idesc_f16(256, COUT), idesc8_p = make_idesc_e4m3(256, COUT);
            if (crank == 0)
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                mbar_wait(tempty(acc), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * (2 * COUT);
                // right sub-tiles: needed if either tile of the pair has one (the other CTA then computes on zero-filled patch rows)
                const int nh = (((tile / tiles_per_tx) * 16 + 8 < p.wo) || (((tile + 1) / tiles_per_tx) * 16 + 8 < p.wo)) ? 2 : 1;
                for (int ch = 0; ch < n_chunk_steps; ++ch) {
                    const bool lo_pass = ch < p.chunks;
                    const bool first_main = ch == p.chunks;
                    mbar_wait(a_full(as), aph);
                    tc_fence_after();
                    const uint32_t a_lo = (((smem_base + as * Cfg::PATCH_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(b_full(bs), bph);
                        tc_fence_after();
                        const uint32_t b_lo = (((b_base + bs * Cfg::B_TILE) & 0x3FFFF) >> 4) | (1u << 16);
                        const uint32_t w_lo = a_lo + ((((tap / 3) * 18 + (tap % 3)) * RB) >> 4);
                        if (leader) {
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                if (h < nh)
#pragma unroll
                                for (int k = 0; k < CC / 16; ++k) {
                                    if (lo_pass) umma_pair_lohi<1>(d0 + h * COUT, w_lo + ((8 * h * RB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4), b_hi, idesc8_p,
                                                                   (ch | tap | k) != 0 ? 1u : 0u);
                                    else if (first_main && tap == 0 && k == 0) umma_pair_lohi<2>(d0 + h * COUT, w_lo + ((8 * h * RB) >> 4), a_hi, b_lo, b_hi, idesc_p, 1u);
                                    else umma_pair_lohi<0>(d0 + h * COUT, w_lo + ((8 * h * RB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4), b_hi, idesc_p, 1u);
                                }
                            umma_commit_pair(b_empty(bs), (uint16_t)3);
                        }
                        __syncwarp();
                        if (++bs == BST) { bs = 0; bph ^= 1; }
                    }
                    if (leader) umma_commit_pair(a_empty(as), (uint16_t)3);
                    __syncwarp();
                    if (++as == AST) { as = 0; aph ^= 1; }
                }
                if (leader) umma_commit_pair(tfull(acc), (uint16_t)3);
                __syncwarp();
                if (++acc == ACC) { acc = 0; acc_ph ^= 1; }
            }
        } else
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            mbar_wait(tempty(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + acc * (2 * COUT);
            // the right 8-column sub-tile of a border tile may lie entirely outside the image (24 columns at level 3 = 1.5 tiles):
            // its UMMAs are skipped, and so is its epilogue
            const int nh = ((tile / tiles_per_tx) * 16 + 8 < p.wo) ? 2 : 1;
            for (int ch = 0; ch < n_chunk_steps; ++ch) {
                const bool lo_pass = F8 && ch < p.chunks;
                const bool first_main = F8 && ch == p.chunks;        // x2: the first FP16 instruction of each sub-tile rescales its accumulator
                mbar_wait(a_full(as), aph);
                tc_fence_after();
                const uint32_t a_lo = (((smem_base + as * Cfg::PATCH_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
                if (RESIDENT) {
                    const uint32_t b_lo = (((b_base + ch * 9 * Cfg::B_TILE) & 0x3FFFF) >> 4) | (1u << 16);
                    if (leader) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                if (h < nh)
#pragma unroll
                                for (int k = 0; k < CC / 16; ++k)
                                    umma_bf16_lohi(d0 + h * COUT, a_lo + ((((tap / 3) * 18 + (tap % 3) + 8 * h) * RB + k * 32) >> 4), a_hi,
                                                   b_lo + ((tap * Cfg::B_TILE + k * 32) >> 4), b_hi, idesc,
                                                   (tap | k) != 0 ? 1u : (ch != 0 ? 1u : 0u));
                    }
                } else {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(b_full(bs), bph);
                        tc_fence_after();
                        const uint32_t b_lo = (((b_base + bs * Cfg::B_TILE) & 0x3FFFF) >> 4) | (1u << 16);
                        const uint32_t w_lo = a_lo + ((((tap / 3) * 18 + (tap % 3)) * RB) >> 4);
                        if (leader) {
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                if (h < nh)
#pragma unroll
                                for (int k = 0; k < CC / 16; ++k) {
                                    if (F8) {
                                        if (lo_pass) umma_f8_lohi(d0 + h * COUT, w_lo + ((8 * h * RB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4), b_hi, idesc8,
                                                                  (ch | tap | k) != 0 ? 1u : 0u);
                                        else if (first_main && tap == 0 && k == 0) umma_f16_lohi_rescale(d0 + h * COUT, w_lo + ((8 * h * RB) >> 4), a_hi, b_lo, b_hi, idesc);
                                        else umma_bf16_lohi(d0 + h * COUT, w_lo + ((8 * h * RB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4), b_hi, idesc, 1u);
                                        continue;
                                    }
                                    umma_bf16_lohi(d0 + h * COUT, w_lo + ((8 * h * RB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4), b_hi,
                                                   idesc, (ch | tap | k) != 0 ? 1u : 0u);
                                    if (X3)                                                          // lo patch . hi weights
                                        umma_bf16_lohi(d0 + h * COUT, w_lo + ((Cfg::PLANE_BYTES + 8 * h * RB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4),
                                                       b_hi, idesc, 1u);
                                }
                            if (CL) umma_commit_mc(b_empty(bs), (uint16_t)3); else umma_commit(b_empty(bs));
                        }
                        __syncwarp();
                        if (++bs == BST) { bs = 0; bph ^= 1; }
                        if (X3) {                                                                    // hi patch . lo weights (next ring slot)
                            mbar_wait(b_full(bs), bph);
                            tc_fence_after();
                            const uint32_t bl_lo = (((b_base + bs * Cfg::B_TILE) & 0x3FFFF) >> 4) | (1u << 16);
                            if (leader) {
#pragma unroll
                                for (int h = 0; h < 2; ++h)
                                    if (h < nh)
#pragma unroll
                                    for (int k = 0; k < CC / 16; ++k)
                                        umma_bf16_lohi(d0 + h * COUT, w_lo + ((8 * h * RB + k * 32) >> 4), a_hi, bl_lo + ((k * 32) >> 4), b_hi, idesc, 1u);
                                if (CL) umma_commit_mc(b_empty(bs), (uint16_t)3); else umma_commit(b_empty(bs));
                            }
                            __syncwarp();
                            if (++bs == BST) { bs = 0; bph ^= 1; }
                        }
                    }
                }
                if (leader) umma_commit(a_empty(as));
                __syncwarp();
                if (++as == AST) { as = 0; aph ^= 1; }
            }
            if (leader) umma_commit(tfull(acc));
            __syncwarp();
            if (++acc == ACC) { acc = 0; acc_ph ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;
        const int r = q * 32 + lane;
        const int ty = r >> 3, txl = r & 7;
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            const int txi = tile / tiles_per_tx, t2 = tile - txi * tiles_per_tx;
            const int n = t2 / p.tiles_y, oy = (t2 - n * p.tiles_y) * 16 + ty, x0 = txi * 16;
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                if (x0 + 8 * h >= p.wo) break;                   // sub-tile outside the image: never computed
                const int ox = x0 + txl + 8 * h;
                const bool live = oy < p.ho && ox < p.wo && n < p.n;
                __nv_bfloat16* dst = p.out + (((size_t)n * p.ho + oy) * p.wo + ox) * COUT;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (2 * COUT) + h * COUT;
                static_assert(!SPLIT || COUT % 64 == 0, "split halo epilogue handles 64 channels per round trip");
                if (SPLIT || (p.relu && COUT % 64 == 0)) {
                    // 64 accumulator columns per round trip (two tcgen05.ld.x32, one wait), packed FFMA2 + F2FP.RELU, 128 contiguous
                    // bytes per pixel: the x16 loop below spent ~200 cycles per 16 columns and made the epilogue slower than the UMMAs
#pragma unroll 1
                    for (int c = 0; c < COUT; c += 64) {
                        uint32_t v[64];
                        tmem_ld32(taddr + c, v);
                        tmem_ld32(taddr + c + 32, v + 32);
                        tmem_ld_wait();
                        if (SPLIT) {
#pragma unroll
                            for (int c8 = 0; c8 < 4; ++c8) {
                                uint32_t oh[8], ol[8];
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const float4 sc = *reinterpret_cast<const float4*>(s_scale + c + 16 * c8 + j);
                                    const float4 sh = *reinterpret_cast<const float4*>(s_shift + c + 16 * c8 + j);
                                    bn_relu_split4<F16, F8>(v + 16 * c8 + j, sc, sh, oh, ol, j / 4);
                                }
                                if (live) { stg256(dst + c + 16 * c8, oh); stg256(dst + p.out_lo + c + 16 * c8, ol); }
                            }
                            continue;
                        }
                        uint32_t o[32];
#pragma unroll
                        for (int j = 0; j < 64; j += 4) {
                            const float4 sc = *reinterpret_cast<const float4*>(s_scale + c + j);
                            const float4 sh = *reinterpret_cast<const float4*>(s_shift + c + j);
                            o[j / 2] = bn_relu_pack<F16>(v[j], v[j + 1], make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                            o[j / 2 + 1] = bn_relu_pack<F16>(v[j + 2], v[j + 3], make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                        }
                        if (live) {
                            uint4* d4 = reinterpret_cast<uint4*>(dst + c);
#pragma unroll
                            for (int j = 0; j < 8; ++j) d4[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                        }
                    }
                    continue;
                }
#pragma unroll 1
                for (int c = 0; c < COUT; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
                    tmem_ld_wait();
                    uint32_t o[8];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 sc = *reinterpret_cast<const float4*>(s_scale + c + 4 * j4);
                        const float4 sh = *reinterpret_cast<const float4*>(s_shift + c + 4 * j4);
                        float a0 = fmaf(__uint_as_float(v[4 * j4]), sc.x, sh.x), a1 = fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y);
                        float a2 = fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), a3 = fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w);
                        if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
                        else if (F16) { a0 = fmaxf(a0, -65504.f); a1 = fmaxf(a1, -65504.f); a2 = fmaxf(a2, -65504.f); a3 = fmaxf(a3, -65504.f); }
                        o[2 * j4] = pack16t<F16>(a0, a1);
                        o[2 * j4 + 1] = pack16t<F16>(a2, a3);
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
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(mapa_rank(tempty(acc), 0));    // the leader's barrier counts both epilogues
                else mbar_arrive(tempty(acc));
            }
            if (++acc == ACC) { acc = 0; acc_ph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL) cluster_sync();                              // the peer may still multicast into this CTA's ring / arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace ukbb
