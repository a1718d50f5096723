// Fused head, version 4 (north_star (b) + (c)): the algebra and warp roles of head_tc.cuh with every
// A operand that the kernel itself produces or that is constant living in TENSOR MEMORY instead of
// shared memory.
//
// Measured (profiles/r1_ts_probe.log, experiments/ts_probe.cu): a kind::f16 UMMA with M = 128, N = 64, K = 16 costs
// 48 cycles when A comes from shared memory (the tensor pipe waits for 4 KB of A + 2 KB of B through the
// 128 B/cycle shared-memory port: 48 wavefronts, which is also what ncu counts per UMMA) and 32 cycles
// -- the full tensor rate for N = 64 -- when A comes from TMEM.  head_tc was bound by that port
// (operand reads + epilogue STS of A0 / A2 + TMA writes ~ 120 KB per 128-pixel tile).  Here
//   * the interpolation matrices U_1..U_4 (constant, 128 x 128 K-values in total) are written to 64 TMEM
//     columns once per CTA with tcgen05.st and stay there,
//   * E0 / E1 write the 16-bit activations A0 (same_dim0 output) / A2 (fc0 output) with tcgen05.st into
//     TMEM columns (lane = pixel, column = channel pair) instead of swizzled STS.128 -- no shared-memory
//     stores, no proxy fence,
//   * S1 / S2 issue tcgen05.mma with [tmem] A operands; only the B operands (weights, TMA-written t_l
//     patches) and the b0 tile of S0 are read from shared memory: ~40 KB per tile instead of ~120 KB.
//   * the level-0 input b0 (16 channels) never touches shared memory either: the E0 warps read their pixel's 32 bytes
//     with two LDG.128 two tiles ahead and tcgen05.st them as the A operand of S0 (the 128-row TMA box it replaces was the
//     most expensive of the five loads per tile: TMA moves about one box row per cycle, profiles/r1_tma_probe3.log),
// TMEM columns (512): D0 2 x 32 | B0 3 x 8 | D1 2 x 64 | D2 2 x 64 | A0 2 x 16 | A2 2 x 32 | U 64.
//
// SPLIT (x3 modes): every 16-bit operand that carries rounding error is a (hi, lo) pair -- b0, A0, A2, the t_l patches and the
// weights -- and every product is hi.hi + lo.hi + hi.lo in the same accumulator (U_l is exact, so the upsample terms are
// U.t_hi + U.t_lo): 3 + 20 + 12 UMMAs per tile instead of 1 + 9 + 4.  TMEM (512): D0 2 x 32 | B0 2 x (8 + 8) | D1 2 x 64 |
// D2 2 x 64 | A0 2 x (16 + 16) | A2 1 x (32 + 32) | U 64 with D0 single-buffered: the chain E0 -> A0 -> S1 (20 UMMAs) is the long
// one (a single A0 buffer measured 2350 cycles per tile), S0 -> D0 -> E0's tcgen05.ld is short.
//
//   S0  D0[128x32] = b0_tile[128x16] . Wsd0^T                      (1 UMMA,  N = 32, A from TMEM)   MMA warp 17
//   E0  A0 = relu(D0 + shift_sd0) -> 16 bit -> TMEM; b0 of tile i + 2 -> TMEM                          warps 0-3
//   S1  D1[128x64] = A0 . W_0^T + sum_l U_l . t_l patch             (2 + 7 UMMAs, A from TMEM)       MMA warp 18
//   E1  A2 = relu(D1 + shift_fc0) -> 16 bit -> TMEM                                                  warps 4-7
//   S2  D2[128x64] = A2 . W_fc1^T                                   (4 UMMAs, A from TMEM)           MMA warp 19
//   E2  f = relu(D2 + shift_fc1) in FP32, class scores in FP32, softmax / argmax / crop / counts      warps 8-15
//       (train_network.py:198-199, deploy_network.py:114-130) -- unchanged from head_tc.cuh
#pragma once
#include "tc_common.cuh"
#include "conv_group.cuh"      // tmem_ld32, TileWalk
#include "head_common.cuh"     // HM_* layout constants, HeadMaps, HeadParams, add_relu_pack / add_relu_split

namespace ukbb {

constexpr int H4_THREADS = 928;                                   // 29 warps: 0-3 E0, 4-7 E1, 8-19 E2 (3 sets), 20 TMA, 21 S0, 22 / 23 S1 even / odd tiles, 24 S2, 25-28 b0 loaders
                                                                  // (SPLIT: warps 16-19 are a second E1 set -- the hi / lo split of 64 channels is
                                                                  //  ~450 instructions per pixel row -- and E2 runs on two sets)
template <bool SPLIT>
struct HeadTsCfg {
    static constexpr int D0S = SPLIT ? 1 : 2;                     // same_dim0 accumulator stages
    // TMEM columns
    static constexpr int D0 = 0, A0 = SPLIT ? 32 : 352, B0 = SPLIT ? 96 : 64, D1 = SPLIT ? 128 : 96, D2 = SPLIT ? 256 : 224, A2 = 384, U1 = 448, U2 = 472,
                         U3 = 488, U4 = 496;
    static constexpr int STAGES = SPLIT ? 7 : 10;                 // input stages (t_l patches): ~2000 cycles of TMA latency at ~700 cycles per tile
    static constexpr int IN_BYTES = (SPLIT ? 2 : 1) * HM_IN_PATCHES;   // hi patches | lo patches
    static constexpr int B0S = SPLIT ? 2 : 3;                     // b0 operand stages in TMEM (8 columns per plane)
    static constexpr int W_BYTES = (SPLIT ? 2 : 1) * (HM_W0 + HM_W1 + HM_WSD);
    static constexpr int SMEM = STAGES * IN_BYTES + W_BYTES + 1024 /*align*/ + 512 /*barriers*/;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int NC, bool F16, bool SPLIT = false, bool F8 = false>
__global__ void __launch_bounds__(H4_THREADS, 1)
head_ts_kernel(const __grid_constant__ HeadMaps maps, const __grid_constant__ HeadParams p) {
    using namespace tc;
    using Cfg = HeadTsCfg<SPLIT>;
    constexpr int H4_STAGES = Cfg::STAGES, H4_B0S = Cfg::B0S, IN_BYTES = Cfg::IN_BYTES, H4_D0S = Cfg::D0S;
    constexpr int H4_D0 = Cfg::D0, H4_B0 = Cfg::B0, H4_D1 = Cfg::D1, H4_D2 = Cfg::D2, H4_A0 = Cfg::A0, H4_A2 = Cfg::A2, H4_U1 = Cfg::U1, H4_U2 = Cfg::U2,
                  H4_U3 = Cfg::U3, H4_U4 = Cfg::U4;
    constexpr int PL = SPLIT ? 2 : 1;                              // operand planes
    constexpr int H4_E2SETS = SPLIT ? 2 : 3;                       // E2 warp sets (4 warps each), tile i -> set i % H4_E2SETS
    constexpr int E1SETS = SPLIT ? 2 : 1;                          // E1 warp sets: set e converts accumulator columns [32 e, 32 e + 32)
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t in_base = smem_base;
    const uint32_t w0_base = in_base + H4_STAGES * IN_BYTES;       // hi [64][32] | lo
    const uint32_t w1_base = w0_base + PL * HM_W0;                 // hi [64][64] | lo
    const uint32_t wsd_base = w1_base + PL * HM_W1;                // hi [32][16] | lo
    const uint32_t bar_base = wsd_base + PL * HM_WSD;
    auto BAR = [&](int i) { return bar_base + 8u * i; };
    enum { WFULL = 0, UFULL = 1, IN_FULL = 2, IN_EMPTY = IN_FULL + 10, D0_FULL = IN_EMPTY + 10, D0_EMPTY = D0_FULL + 2,
           B0_FULL = D0_EMPTY + 2, B0_EMPTY = B0_FULL + 3, A0_FULL = B0_EMPTY + 3, A0_EMPTY = A0_FULL + 2, D1_FULL = A0_EMPTY + 2, D1_EMPTY = D1_FULL + 2, A2_FULL = D1_EMPTY + 2,
           A2_EMPTY = A2_FULL + 2, D2_FULL = A2_EMPTY + 2, D2_EMPTY = D2_FULL + 6, TSLOT = D2_EMPTY + 6 };
    // D2 has two TMEM buffers (tile i -> i & 1) but SIX barrier pairs (tile i -> i % 6): with three E2 warp sets (tile i -> set i % 3) a
    // barrier must belong to one set only -- a parity wait can tell the current phase from the previous one, not from the one before
    static_assert((TSLOT + 1) * 8 <= 512, "barrier area");
    const uint32_t tmem_slot = BAR(TSLOT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 20 && lane == 0) {
        const CUtensorMap* m = &maps.t1;
        for (int i = 0; i < 7; ++i) tma_prefetch_desc(m + i);
    }
    if (warp == 21 && lane == 0) {
        mbar_init(BAR(WFULL), 1);
        mbar_init(BAR(UFULL), 4);
        for (int s = 0; s < H4_STAGES; ++s) { mbar_init(BAR(IN_FULL + s), 1); mbar_init(BAR(IN_EMPTY + s), 1); }
        for (int d = 0; d < H4_D0S; ++d) { mbar_init(BAR(D0_FULL + d), 1); mbar_init(BAR(D0_EMPTY + d), 4); }
        for (int d = 0; d < H4_B0S; ++d) { mbar_init(BAR(B0_FULL + d), 4); mbar_init(BAR(B0_EMPTY + d), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(A0_FULL + b), 4); mbar_init(BAR(A0_EMPTY + b), 1); mbar_init(BAR(D1_FULL + b), 1); mbar_init(BAR(D1_EMPTY + b), 4 * E1SETS);
            mbar_init(BAR(A2_FULL + b), 4 * E1SETS); mbar_init(BAR(A2_EMPTY + b), 1);
        }
        for (int k = 0; k < 6; ++k) { mbar_init(BAR(D2_FULL + k), 1); mbar_init(BAR(D2_EMPTY + k), 4); }
        fence_barrier_init();
    }
    if (warp == 24) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // zero the input stages once: the K-padding rows of the t_l patches are never written by TMA and
    // must be finite (they meet zero columns of U_l)
    for (int i = threadIdx.x; i < H4_STAGES * IN_BYTES / 16; i += H4_THREADS)
        reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 20) griddep_wait();                      // the producer waits after it has issued the weight loads
    const int my_tiles = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto LO = [](uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | (1u << 16); };
    constexpr uint32_t HI32 = (uint32_t)((8 * 32) >> 4) | (1u << 14) | (6u << 29);
    constexpr uint32_t HI64 = (uint32_t)((8 * 64) >> 4) | (1u << 14) | (4u << 29);
    constexpr uint32_t HI128 = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);

    if (warp == 20) {
        // ===================== TMA producer: the t_l patch loads of a tile (4, or 8 with the lo planes) are one warp instruction =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(WFULL), PL * (HM_W0 + HM_W1 + HM_WSD));
            for (int pl = 0; pl < PL; ++pl) {                                       // lo weights = rows [cout, 2 cout) of the same maps
                tma_load_2d(wsd_base + pl * HM_WSD, &maps.wsd, BAR(WFULL), 0, pl * 32);
                tma_load_2d(w0_base + pl * HM_W0, &maps.w0, BAR(WFULL), 0, pl * 64);
                tma_load_2d(w1_base + pl * HM_W1, &maps.w1, BAR(WFULL), 0, pl * 64);
            }
        }
        griddep_wait();
        __syncwarp();
        TileWalk w;
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        const bool loader = lane >= 1 && lane <= 4 * PL;
        const int l = loader ? ((lane - 1) & 3) + 1 : 1, plane = loader ? (lane - 1) >> 2 : 0;
        const CUtensorMap* my_map = &maps.t1 + (l - 1);                             // t1, t2, t3, t4 are adjacent
        const uint32_t my_off = (uint32_t)(plane * HM_IN_PATCHES) +
                                (l == 1 ? 0u : l == 2 ? (uint32_t)HM_IN_P1 : l == 3 ? (uint32_t)(HM_IN_P1 + HM_IN_P2) : (uint32_t)(HM_IN_P1 + HM_IN_P2 + HM_IN_P3));
        const int pb = ((1 << l) - 1) >> 1;                                         // level l patch origin: ((x0 + pb) >> l) - 1
        const int n_off = plane * p.lo_n;
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int y0 = w.ty * 8, x0 = w.tx * 16, n = w.n;
            mbar_wait(BAR(IN_EMPTY + s), ph ^ 1);
            const uint32_t dst = in_base + s * IN_BYTES;
            const uint32_t fullb = BAR(IN_FULL + s);
            if (lane == 0) mbar_arrive_expect_tx(fullb, PL * HM_IN_PATCH_TX);       // t_l patches only: b0 goes through registers
            __syncwarp();
            if (loader) tma_load_4d(dst + my_off, my_map, fullb, 0, ((x0 + pb) >> l) - 1, ((y0 + pb) >> l) - 1, n_off + n);
            __syncwarp();
            if (++s == H4_STAGES) { s = 0; ph ^= 1; }
            w.next();
        }
    } else if (warp == 21) {
        // ===================== MMA issuer 0: same_dim0 (S0), A = b0 tile in TMEM (written by the E0 warps) =====================
        const bool leader = elect_one();
        const uint32_t idesc_sd = F16 ? make_idesc_f16(128, 32) : make_idesc_bf16(128, 32);
        const uint32_t wsd_lo = LO(wsd_base);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        int s = 0, d = 0;
        uint32_t ph = 0, dph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(BAR(D0_EMPTY + d), dph ^ 1);
            mbar_wait(BAR(B0_FULL + s), ph);
            tc_fence_after();
            if (leader) {
                if (F8) {    // x2 scheme: the lo plane of b0 and of Wsd0 are FP8 correction operands (one K = 32 step), then the FP16 term rescales
                    umma_ts_f8_lohi(tmem_base + H4_D0 + d * 32, tmem_base + H4_B0 + s * 16 + 8, wsd_lo + (HM_WSD >> 4), HI32, make_idesc_e4m3(128, 32), 0u);
                    umma_ts_lohi_rescale(tmem_base + H4_D0 + d * 32, tmem_base + H4_B0 + s * 16, wsd_lo, HI32, idesc_sd);
                } else {
                umma_ts_lohi(tmem_base + H4_D0 + d * 32, tmem_base + H4_B0 + s * 8 * PL, wsd_lo, HI32, idesc_sd, 0u);
                }
                if (SPLIT && !F8) {
                    umma_ts_lohi(tmem_base + H4_D0 + d * 32, tmem_base + H4_B0 + s * 16 + 8, wsd_lo, HI32, idesc_sd, 1u);                   // lo . hi
                    umma_ts_lohi(tmem_base + H4_D0 + d * 32, tmem_base + H4_B0 + s * 16, wsd_lo + (HM_WSD >> 4), HI32, idesc_sd, 1u);       // hi . lo
                }
                umma_commit(BAR(B0_EMPTY + s));
                umma_commit(BAR(D0_FULL + d));
            }
            __syncwarp();
            if (++s == H4_B0S) { s = 0; ph ^= 1; }
            if (++d == H4_D0S) { d = 0; dph ^= 1; }
        }
    } else if (warp == 22 || warp == 23) {
        // ===================== MMA issuers 1a / 1b: fc0 with the upsample terms (S1), A operands in TMEM =====================
        // Two warps, even / odd tiles: one issuer needed ~1000 cycles per tile (150 dependent scalar instructions of loop, barrier and
        // descriptor bookkeeping at ~4.4 cycles each + 12 queue-limited tcgen05 issues) and was what every other role waited for.
        const int par = warp - 22;
        const bool leader = elect_one();
        const uint32_t idesc_kk = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        const uint32_t idesc_kmn = idesc_kk | (1u << 16);            // B operand MN-major (pixel-major t_l patch)
        const uint32_t w0_lo = LO(w0_base);
        mbar_wait(BAR(WFULL), 0);
        mbar_wait(BAR(UFULL), 0);
        tc_fence_after();
        TileWalk w;                                                 // needs the tile-row parity for U_4
        w.init(blockIdx.x + par * gridDim.x, 2 * gridDim.x, p.tiles_x, p.tiles_y);
        int s = par;
        uint32_t sph = 0;
        for (int i = par; i < my_tiles; i += 2) {
            const int b = par;
            const uint32_t bph = ((uint32_t)i >> 1) & 1u;
            const uint32_t v = (uint32_t)(w.ty & 1);                 // tile-row parity selects the U_4 variant
            mbar_wait(BAR(D1_EMPTY + b), bph ^ 1);
            mbar_wait(BAR(A0_FULL + b), bph);
            mbar_wait(BAR(IN_FULL + s), sph);
            tc_fence_after();
            const uint32_t d = tmem_base + H4_D1 + b * 64;
            const uint32_t in_lo = LO(in_base + s * IN_BYTES);
            const uint32_t a0 = tmem_base + H4_A0 + b * 16 * PL;
            if (leader) {
                umma_ts_lohi(d, a0, w0_lo, HI64, idesc_kk, 0u);
                umma_ts_lohi(d, a0 + 8, w0_lo + 2, HI64, idesc_kk, 1u);
                if (SPLIT) {
                    umma_ts_lohi(d, a0 + 16, w0_lo, HI64, idesc_kk, 1u);                                  // A0_lo . W0_hi
                    umma_ts_lohi(d, a0 + 24, w0_lo + 2, HI64, idesc_kk, 1u);
                    umma_ts_lohi(d, a0, w0_lo + (HM_W0 >> 4), HI64, idesc_kk, 1u);                        // A0_hi . W0_lo
                    umma_ts_lohi(d, a0 + 8, w0_lo + (HM_W0 >> 4) + 2, HI64, idesc_kk, 1u);
                }
#pragma unroll
                for (int pl = 0; pl < PL; ++pl) {                                                         // U_l is exact: U . t_hi + U . t_lo
                    const uint32_t t_lo = in_lo + ((pl * HM_IN_PATCHES) >> 4);
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        umma_ts_lohi(d, tmem_base + H4_U1 + 8 * k, t_lo + ((k * 2048) >> 4), HI128, idesc_kmn, 1u);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_ts_lohi(d, tmem_base + H4_U2 + 8 * k, t_lo + ((HM_IN_P1 + k * 2048) >> 4), HI128, idesc_kmn, 1u);
                    umma_ts_lohi(d, tmem_base + H4_U3, t_lo + ((HM_IN_P1 + HM_IN_P2) >> 4), HI128, idesc_kmn, 1u);
                    umma_ts_lohi(d, tmem_base + H4_U4 + 8 * v, t_lo + ((HM_IN_P1 + HM_IN_P2 + HM_IN_P3) >> 4), HI128, idesc_kmn, 1u);
                }
                umma_commit(BAR(IN_EMPTY + s));
                umma_commit(BAR(A0_EMPTY + b));
                umma_commit(BAR(D1_FULL + b));
            }
            __syncwarp();
            s += 2;
            if (s >= H4_STAGES) { s -= H4_STAGES; sph ^= 1; }
            w.next();
        }
    } else if (warp == 24) {
        // ===================== MMA issuer 2: fc1 (S2), A operand in TMEM =====================
        const bool leader = elect_one();
        const uint32_t idesc_kk = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        const uint32_t w1_lo = LO(w1_base);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            const uint32_t bph = ((uint32_t)i >> 1) & 1u;
            if (i >= 2) mbar_wait(BAR(D2_EMPTY + (i - 2) % 6), (uint32_t)((i - 2) / 6) & 1u);   // the tile that used D2[b] before
            if (SPLIT) mbar_wait(BAR(A2_FULL), (uint32_t)i & 1u); else mbar_wait(BAR(A2_FULL + b), bph);
            tc_fence_after();
            const uint32_t d = tmem_base + H4_D2 + b * 64;
            const uint32_t a2 = tmem_base + H4_A2 + (SPLIT ? 0 : b * 32);
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_ts_lohi(d, a2 + 8 * k, w1_lo + 2 * k, HI128, idesc_kk, k != 0 ? 1u : 0u);
                    if (SPLIT) {
                        umma_ts_lohi(d, a2 + 32 + 8 * k, w1_lo + 2 * k, HI128, idesc_kk, 1u);                 // A2_lo . W1_hi
                        umma_ts_lohi(d, a2 + 8 * k, w1_lo + (HM_W1 >> 4) + 2 * k, HI128, idesc_kk, 1u);       // A2_hi . W1_lo
                    }
                }
                umma_commit(BAR(A2_EMPTY + (SPLIT ? 0 : b)));
                umma_commit(BAR(D2_FULL + i % 6));
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ===================== U_l -> TMEM once; then b0 staging and E0 (D0 -> A0), warps 0-3 =====================
        const int q = warp;
        const int r = q * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        {
            uint32_t u[32];
            const uint4* g1 = reinterpret_cast<const uint4*>(p.u_glob[1] + (size_t)r * 32);          // U_1 row: 64 values, 48 used
#pragma unroll
            for (int j = 0; j < 6; ++j) { const uint4 t = __ldg(g1 + j); u[4 * j] = t.x; u[4 * j + 1] = t.y; u[4 * j + 2] = t.z; u[4 * j + 3] = t.w; }
            tmem_st16(lane_base + H4_U1, u);
            tmem_st8(lane_base + H4_U1 + 16, u + 16);
            const uint4* g2 = reinterpret_cast<const uint4*>(p.u_glob[2] + (size_t)r * 16);          // U_2 row: 32 values
#pragma unroll
            for (int j = 0; j < 4; ++j) { const uint4 t = __ldg(g2 + j); u[4 * j] = t.x; u[4 * j + 1] = t.y; u[4 * j + 2] = t.z; u[4 * j + 3] = t.w; }
            tmem_st16(lane_base + H4_U2, u);
            const uint4* g3 = reinterpret_cast<const uint4*>(p.u_glob[3] + (size_t)r * 8);           // U_3 row: 16 values
            const uint4* g4 = reinterpret_cast<const uint4*>(p.u_glob[4] + (size_t)r * 8);           // U_4 rows: 2 variants x 16 values
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint4 t = __ldg(g3 + j); u[4 * j] = t.x; u[4 * j + 1] = t.y; u[4 * j + 2] = t.z; u[4 * j + 3] = t.w;
                const uint4 a = __ldg(g4 + j); u[8 + 4 * j] = a.x; u[8 + 4 * j + 1] = a.y; u[8 + 4 * j + 2] = a.z; u[8 + 4 * j + 3] = a.w;
                const uint4 c = __ldg(g4 + 128 * 2 + j); u[16 + 4 * j] = c.x; u[16 + 4 * j + 1] = c.y; u[16 + 4 * j + 2] = c.z; u[16 + 4 * j + 3] = c.w;
            }
            tmem_st8(lane_base + H4_U3, u);
            tmem_st16(lane_base + H4_U4, u + 8);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(UFULL));
        }
        // E0 only: E1 runs on warps 4-7 and the b0 operand is staged by warps 20-23.  A single warp executes a dependent
        // instruction stream at ~4 cycles per instruction and every tcgen05.ld / st round trip costs > 100 cycles, so one role per
        // warp keeps each per-tile loop short (timeline: experiments/trace_head.py).
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1, d = i % H4_D0S;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            mbar_wait(BAR(A0_EMPTY + b), ph ^ 1);
            mbar_wait(BAR(D0_FULL + d), (uint32_t)(i / H4_D0S) & 1u);
            tc_fence_after();
            if (SPLIT) {
                uint32_t v[32];
                tmem_ld32(lane_base + H4_D0 + d * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(D0_EMPTY + d));        // D0 is single-buffered here: release it before the conversion
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t oh[8], ol[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        add_relu_split<F16>(v[16 * hf + 2 * j], v[16 * hf + 2 * j + 1], p.c_shift_sd0[16 * hf + 2 * j], p.c_shift_sd0[16 * hf + 2 * j + 1], oh[j], ol[j]);
                    tmem_st8(lane_base + H4_A0 + b * 32 + 8 * hf, oh);
                    tmem_st8(lane_base + H4_A0 + b * 32 + 16 + 8 * hf, ol);
                }
            } else {
                uint32_t v[32];
                tmem_ld32(lane_base + H4_D0 + d * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(D0_EMPTY + d));
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = add_relu_pack<F16>(v[2 * j], v[2 * j + 1], p.c_shift_sd0[2 * j], p.c_shift_sd0[2 * j + 1]);
                tmem_st16(lane_base + H4_A0 + b * 16, o);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A0_FULL + b));
        }
    } else if (warp >= 25) {
        // ===================== b0 loaders (warps 25-28): the A operand of S0 goes global -> registers -> TMEM =====================
        // This thread's pixel (tile row r / 16, column r % 16), 16 channels = 32 bytes = two LDG.128.  b0 comes from HBM (0.64 GB per
        // subject, written by conv_first long before): the loads run three tiles ahead in a register ring, the TMEM stages another three.
        const int q = (warp - 25) & 3;                               // TMEM lane quarter = warp % 4: warps 25..28 -> 1, 2, 3, 0
        const int qq = warp & 3;
        const int r = qq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(qq * 32) << 16);
        (void)q;
        const int pty = r >> 4, ptx = r & 15;
        TileWalk wb;
        wb.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        auto b0_addr = [&]() { return p.b0 + (((size_t)wb.n * p.h + wb.ty * 8 + pty) * p.w + wb.tx * 16 + ptx) * 2; };
        constexpr int RING = SPLIT ? 2 : 3;                           // register ring = TMEM stages (tiles ahead)
        uint4 ring[RING][2 * PL];
#pragma unroll
        for (int u = 0; u < RING; ++u) {
#pragma unroll
            for (int e = 0; e < 2 * PL; ++e) ring[u][e] = make_uint4(0, 0, 0, 0);
            if (u < my_tiles) {
                const uint4* g = b0_addr();
                ring[u][0] = __ldg(g); ring[u][1] = __ldg(g + 1);
                if (SPLIT) { ring[u][2] = __ldg(g + p.b0_lo); ring[u][3] = __ldg(g + p.b0_lo + 1); }
                wb.next();
            }
        }
        for (int j0 = 0; j0 < my_tiles; j0 += RING) {
#pragma unroll
            for (int u = 0; u < RING; ++u) {
                const int j = j0 + u;
                if (j >= my_tiles) break;
                const int sb = u;                                    // j % RING
                mbar_wait(BAR(B0_EMPTY + sb), ((uint32_t)(j / H4_B0S) & 1u) ^ 1u);
                tc_fence_after();
#pragma unroll
                for (int pl = 0; pl < PL; ++pl) {
                    const uint32_t t8[8] = {ring[u][2 * pl].x, ring[u][2 * pl].y, ring[u][2 * pl].z, ring[u][2 * pl].w,
                                            ring[u][2 * pl + 1].x, ring[u][2 * pl + 1].y, ring[u][2 * pl + 1].z, ring[u][2 * pl + 1].w};
                    tmem_st8(lane_base + H4_B0 + sb * 8 * PL + 8 * pl, t8);
                }
                if (j + RING < my_tiles) {
                    const uint4* g = b0_addr();
                    ring[u][0] = __ldg(g); ring[u][1] = __ldg(g + 1);
                    if (SPLIT) { ring[u][2] = __ldg(g + p.b0_lo); ring[u][3] = __ldg(g + p.b0_lo + 1); }
                    wb.next();
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B0_FULL + sb));
            }
        }
    } else if (warp < 8 || (SPLIT && warp >= 16 && warp < 20)) {
        // ===================== E1 (D1 -> A2), warps 4-7 (SPLIT: and warps 16-19 for the upper 32 columns) =====================
        const int q = warp & 3;
        const int e1 = warp < 8 ? 0 : 1;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            if (!SPLIT) mbar_wait(BAR(A2_EMPTY + b), ph ^ 1);
            mbar_wait(BAR(D1_FULL + b), ph);
            tc_fence_after();
            if (SPLIT) {
                // A2 is single-buffered: convert this set's 32 accumulator columns into registers FIRST, wait for S2 of the previous
                // tile to release A2 only then, and store -- the loop S2 -> A2_EMPTY -> E1 -> A2_FULL -> S2 then holds four tcgen05.st
                // instead of the whole conversion.
                uint32_t oh[16], ol[16];
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {                      // 16 accumulator columns at a time (register budget: 64)
                    const int qt = 2 * e1 + qq;
                    uint32_t v[16];
                    tmem_ld16(lane_base + H4_D1 + b * 64 + 16 * qt, v);
                    tmem_ld_wait();
                    if (qq == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(D1_EMPTY + b));
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        add_relu_split<F16>(v[2 * j], v[2 * j + 1], p.c_shift0[16 * qt + 2 * j], p.c_shift0[16 * qt + 2 * j + 1], oh[8 * qq + j], ol[8 * qq + j]);
                }
                mbar_wait(BAR(A2_EMPTY), ((uint32_t)i & 1u) ^ 1u);
                tc_fence_after();
                tmem_st16(lane_base + H4_A2 + 16 * e1, oh);
                tmem_st16(lane_base + H4_A2 + 32 + 16 * e1, ol);
            } else {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                      // 32 accumulator columns at a time (register budget of a 29-warp CTA: 64)
                    uint32_t v[32];
                    tmem_ld32(lane_base + H4_D1 + b * 64 + 32 * hf, v);
                    tmem_ld_wait();
                    if (hf == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(D1_EMPTY + b));
                    }
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = add_relu_pack<F16>(v[2 * j], v[2 * j + 1], p.c_shift0[32 * hf + 2 * j], p.c_shift0[32 * hf + 2 * j + 1]);
                    tmem_st16(lane_base + H4_A2 + b * 32 + 16 * hf, o);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A2_FULL + (SPLIT ? 0 : b)));
        }
    } else if (warp < 20) {
        // ===================== E2: FP32 class scores -> labels; warp sets 8-11, 12-15 (and 16-19 when not SPLIT), tile i -> set i % H4_E2SETS =====================
        const int set = (warp - 8) >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int ty = r >> 4, tx = r & 15;
        TileWalk w;
        w.init(blockIdx.x + set * gridDim.x, H4_E2SETS * gridDim.x, p.tiles_x, p.tiles_y);
        for (int i = set; i < my_tiles; i += H4_E2SETS) {
            const int b = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            mbar_wait(BAR(D2_FULL + i % 6), (uint32_t)(i / 6) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + H4_D2 + b * 64;
            // class scores in FP32: relu(d + s) . w = max(d, -s) . w + s . w, and the constant s . w is part of the bias
            // (host, float64), so an input channel costs one FMNMX and one packed FFMA2 per PAIR of classes; weights,
            // negated shifts and bias travel by value in the kernel parameters (constant bank / uniform registers).
            // The accumulator is read in two halves of 32 columns (register budget of a 25-warp CTA: 80).
            constexpr int NC2 = (NC + 1) / 2;
            uint64_t lg2[NC2];
#pragma unroll
            for (int j = 0; j < NC2; ++j) asm("mov.b64 %0, {%1, %2};" : "=l"(lg2[j]) : "f"(p.c_bias2[2 * j]), "f"(p.c_bias2[2 * j + 1]));
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[32];
                tmem_ld32(taddr + 32 * hf, v);
                tmem_ld_wait();
                if (hf == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(D2_EMPTY + i % 6));  // D2[b] drained: S2(i + 2) may overwrite it
                }
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) {
                    const int k = 32 * hf + kk;
                    const float f = fmaxf(__uint_as_float(v[kk]), p.c_nshift1[k]);
                    uint64_t ff;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(ff) : "f"(f));
#pragma unroll
                    for (int j = 0; j < NC2; ++j) {
                        uint64_t wj;
                        asm("mov.b64 %0, {%1, %2};" : "=l"(wj) : "f"(p.c_wlc[k][2 * j]), "f"(p.c_wlc[k][2 * j + 1]));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(lg2[j]) : "l"(ff), "l"(wj));
                    }
                }
            }
            float lg[NC];
#pragma unroll
            for (int j = 0; j < NC2; ++j) {
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(lg2[j]));
                lg[2 * j] = lo;
                if (2 * j + 1 < NC) lg[2 * j + 1] = hi;
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
    if (warp == 24) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace ukbb
