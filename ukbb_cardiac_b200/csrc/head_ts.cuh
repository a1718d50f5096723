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
// TMEM columns (512): D0 3 x 32 | D1 2 x 64 | D2 2 x 64 | A0 2 x 16 | A2 2 x 32 | U 64.
//
//   S0  D0[128x32] = b0_tile[128x16] . Wsd0^T                      (1 UMMA,  N = 32, A from smem)   MMA warp 13
//   E0  A0 = relu(D0 + shift_sd0) -> 16 bit -> TMEM                                                  warps 0-3
//   S1  D1[128x64] = A0 . W_0^T + sum_l U_l . t_l patch             (2 + 7 UMMAs, A from TMEM)       MMA warp 14
//   E1  A2 = relu(D1 + shift_fc0) -> 16 bit -> TMEM                                                  warps 0-3
//   S2  D2[128x64] = A2 . W_fc1^T                                   (4 UMMAs, A from TMEM)           MMA warp 15
//   E2  f = relu(D2 + shift_fc1) in FP32, class scores in FP32, softmax / argmax / crop / counts      warps 4-11
//       (train_network.py:198-199, deploy_network.py:114-130) -- unchanged from head_tc.cuh
#pragma once
#include "tc_common.cuh"
#include "conv_group.cuh"      // tmem_ld32, TileWalk
#include "head_mma.cuh"        // HM_* layout constants, HeadMmaMaps, HeadParams
#include "head_tc.cuh"         // add_relu_pack

namespace ukbb {

namespace tc {
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
}  // namespace tc

constexpr int H4_THREADS = 512;
constexpr int H4_STAGES = 6;                                      // input stages
constexpr int H4_D0S = 3;                                         // same_dim0 accumulator stages
constexpr int H4_SMEM = H4_STAGES * HM_IN_BYTES + HM_W0 + HM_W1 + HM_WSD + 1024 /*align*/ + 512 /*barriers*/;
// TMEM columns
constexpr int H4_D0 = 0, H4_D1 = 96, H4_D2 = 224, H4_A0 = 352, H4_A2 = 384, H4_U1 = 448, H4_U2 = 472, H4_U3 = 488, H4_U4 = 496;

template <int NC, bool F16>
__global__ void __launch_bounds__(H4_THREADS, 1)
head_ts_kernel(const __grid_constant__ HeadMmaMaps maps, const __grid_constant__ HeadParams p) {
    using namespace tc;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t in_base = smem_base;
    const uint32_t w0_base = in_base + H4_STAGES * HM_IN_BYTES;
    const uint32_t w1_base = w0_base + HM_W0;
    const uint32_t wsd_base = w1_base + HM_W1;
    const uint32_t bar_base = wsd_base + HM_WSD;
    auto BAR = [&](int i) { return bar_base + 8u * i; };
    enum { WFULL = 0, UFULL = 1, IN_FULL = 2, IN_EMPTY = IN_FULL + H4_STAGES, D0_FULL = IN_EMPTY + H4_STAGES, D0_EMPTY = D0_FULL + H4_D0S,
           A0_FULL = D0_EMPTY + H4_D0S, A0_EMPTY = A0_FULL + 2, D1_FULL = A0_EMPTY + 2, D1_EMPTY = D1_FULL + 2, A2_FULL = D1_EMPTY + 2,
           A2_EMPTY = A2_FULL + 2, D2_FULL = A2_EMPTY + 2, D2_EMPTY = D2_FULL + 2, TSLOT = D2_EMPTY + 2 };
    static_assert((TSLOT + 1) * 8 <= 512, "barrier area");
    const uint32_t tmem_slot = BAR(TSLOT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 12 && lane == 0) {
        const CUtensorMap* m = &maps.s0;
        for (int i = 0; i < 12; ++i) tma_prefetch_desc(m + i);
    }
    if (warp == 13 && lane == 0) {
        mbar_init(BAR(WFULL), 1);
        mbar_init(BAR(UFULL), 4);
        for (int s = 0; s < H4_STAGES; ++s) { mbar_init(BAR(IN_FULL + s), 1); mbar_init(BAR(IN_EMPTY + s), 1); }
        for (int d = 0; d < H4_D0S; ++d) { mbar_init(BAR(D0_FULL + d), 1); mbar_init(BAR(D0_EMPTY + d), 4); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(A0_FULL + b), 4); mbar_init(BAR(A0_EMPTY + b), 1); mbar_init(BAR(D1_FULL + b), 1); mbar_init(BAR(D1_EMPTY + b), 4);
            mbar_init(BAR(A2_FULL + b), 4); mbar_init(BAR(A2_EMPTY + b), 1); mbar_init(BAR(D2_FULL + b), 1); mbar_init(BAR(D2_EMPTY + b), 4);
        }
        fence_barrier_init();
    }
    if (warp == 15) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // zero the input stages once: the K-padding rows of the t_l patches are never written by TMA and
    // must be finite (they meet zero columns of U_l)
    for (int i = threadIdx.x; i < H4_STAGES * HM_IN_BYTES / 16; i += H4_THREADS)
        reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 12) griddep_wait();                      // the producer waits after it has issued the weight loads
    const int my_tiles = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto LO = [](uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | (1u << 16); };
    constexpr uint32_t HI32 = (uint32_t)((8 * 32) >> 4) | (1u << 14) | (6u << 29);
    constexpr uint32_t HI64 = (uint32_t)((8 * 64) >> 4) | (1u << 14) | (4u << 29);
    constexpr uint32_t HI128 = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);

    if (warp == 12) {
        // ===================== TMA producer (as head_tc: five box loads per tile in one warp instruction) =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(WFULL), HM_W0 + HM_W1 + HM_WSD);
            tma_load_2d(wsd_base, &maps.wsd, BAR(WFULL), 0, 0);
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
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int y0 = w.ty * 8, x0 = w.tx * 16, n = w.n;
            mbar_wait(BAR(IN_EMPTY + s), ph ^ 1);
            const uint32_t dst = in_base + s * HM_IN_BYTES;
            const uint32_t fullb = BAR(IN_FULL + s);
            if (lane == 0) mbar_arrive_expect_tx(fullb, HM_IN_TX);
            __syncwarp();
            if (lane < 5) tma_load_4d(dst + my_off, my_map, fullb, 0, ((x0 + pb) >> l) - back, ((y0 + pb) >> l) - back, n);
            __syncwarp();
            if (++s == H4_STAGES) { s = 0; ph ^= 1; }
            w.next();
        }
    } else if (warp == 13) {
        // ===================== MMA issuer 0: same_dim0 (S0), A = b0 tile in shared memory =====================
        const bool leader = elect_one();
        const uint32_t idesc_sd = F16 ? make_idesc_f16(128, 32) : make_idesc_bf16(128, 32);
        const uint32_t wsd_lo = LO(wsd_base);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        int s = 0, d = 0;
        uint32_t ph = 0, dph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(BAR(D0_EMPTY + d), dph ^ 1);
            mbar_wait(BAR(IN_FULL + s), ph);
            tc_fence_after();
            if (leader) {
                umma_bf16_lohi(tmem_base + H4_D0 + d * 32, LO(in_base + s * HM_IN_BYTES), HI32, wsd_lo, HI32, idesc_sd, 0u);
                umma_commit(BAR(D0_FULL + d));
            }
            __syncwarp();
            if (++s == H4_STAGES) { s = 0; ph ^= 1; }
            if (++d == H4_D0S) { d = 0; dph ^= 1; }
        }
    } else if (warp == 14) {
        // ===================== MMA issuer 1: fc0 with the upsample terms (S1), A operands in TMEM =====================
        const bool leader = elect_one();
        const uint32_t idesc_kk = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        const uint32_t idesc_kmn = idesc_kk | (1u << 16);            // B operand MN-major (pixel-major t_l patch)
        const uint32_t w0_lo = LO(w0_base);
        mbar_wait(BAR(WFULL), 0);
        mbar_wait(BAR(UFULL), 0);
        tc_fence_after();
        TileWalk w;                                                 // needs the tile-row parity for U_4
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        int s = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            const uint32_t bph = ((uint32_t)i >> 1) & 1u;
            const uint32_t v = (uint32_t)(w.ty & 1);                 // tile-row parity selects the U_4 variant
            mbar_wait(BAR(D1_EMPTY + b), bph ^ 1);
            mbar_wait(BAR(A0_FULL + b), bph);                       // implies IN_FULL[s] of this tile (S0 ran on it)
            tc_fence_after();
            const uint32_t d = tmem_base + H4_D1 + b * 64;
            const uint32_t in_lo = LO(in_base + s * HM_IN_BYTES);
            const uint32_t a0 = tmem_base + H4_A0 + b * 16;
            if (leader) {
                umma_ts_lohi(d, a0, w0_lo, HI64, idesc_kk, 0u);
                umma_ts_lohi(d, a0 + 8, w0_lo + 2, HI64, idesc_kk, 1u);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    umma_ts_lohi(d, tmem_base + H4_U1 + 8 * k, in_lo + ((HM_IN_S0 + k * 2048) >> 4), HI128, idesc_kmn, 1u);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_ts_lohi(d, tmem_base + H4_U2 + 8 * k, in_lo + ((HM_IN_S0 + HM_IN_P1 + k * 2048) >> 4), HI128, idesc_kmn, 1u);
                umma_ts_lohi(d, tmem_base + H4_U3, in_lo + ((HM_IN_S0 + HM_IN_P1 + HM_IN_P2) >> 4), HI128, idesc_kmn, 1u);
                umma_ts_lohi(d, tmem_base + H4_U4 + 8 * v, in_lo + ((HM_IN_S0 + HM_IN_P1 + HM_IN_P2 + HM_IN_P3) >> 4), HI128, idesc_kmn, 1u);
                umma_commit(BAR(IN_EMPTY + s));
                umma_commit(BAR(A0_EMPTY + b));
                umma_commit(BAR(D1_FULL + b));
            }
            __syncwarp();
            if (++s == H4_STAGES) s = 0;
            w.next();
        }
    } else if (warp == 15) {
        // ===================== MMA issuer 2: fc1 (S2), A operand in TMEM =====================
        const bool leader = elect_one();
        const uint32_t idesc_kk = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        const uint32_t w1_lo = LO(w1_base);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            const uint32_t bph = ((uint32_t)i >> 1) & 1u;
            mbar_wait(BAR(D2_EMPTY + b), bph ^ 1);
            mbar_wait(BAR(A2_FULL + b), bph);
            tc_fence_after();
            const uint32_t d = tmem_base + H4_D2 + b * 64;
            const uint32_t a2 = tmem_base + H4_A2 + b * 32;
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ts_lohi(d, a2 + 8 * k, w1_lo + 2 * k, HI128, idesc_kk, k != 0 ? 1u : 0u);
                umma_commit(BAR(A2_EMPTY + b));
                umma_commit(BAR(D2_FULL + b));
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ===================== U_l -> TMEM once; then E0 (D0 -> A0) and E1 (D1 -> A2, two tiles behind), warps 0-3 =====================
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
        // E1 trails E0 by TWO tiles: S1(i) (9 UMMAs) is issued at the end of iteration i and its accumulator is first
        // needed in iteration i + 2, so this warp never sits waiting for the tensor pipe between its two roles.
        for (int it = 0; it <= my_tiles + 1; ++it) {
            if (it >= 2) {
                const int i = it - 2, b = i & 1;
                const uint32_t ph = ((uint32_t)i >> 1) & 1u;
                mbar_wait(BAR(A2_EMPTY + b), ph ^ 1);              // satisfied long before the accumulator is: off the critical path
                mbar_wait(BAR(D1_FULL + b), ph);
                tc_fence_after();
                uint32_t v[64];
                tmem_ld32(lane_base + H4_D1 + b * 64, v);
                tmem_ld32(lane_base + H4_D1 + b * 64 + 32, v + 32);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(D1_EMPTY + b));
                uint32_t o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = add_relu_pack<F16>(v[2 * j], v[2 * j + 1], p.c_shift0[2 * j], p.c_shift0[2 * j + 1]);
                tmem_st32(lane_base + H4_A2 + b * 32, o);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(A2_FULL + b));
            }
            if (it < my_tiles) {
                const int i = it, b = i & 1, d = i % H4_D0S;
                const uint32_t ph = ((uint32_t)i >> 1) & 1u;
                mbar_wait(BAR(A0_EMPTY + b), ph ^ 1);              // satisfied long before the accumulator is: off the critical path
                mbar_wait(BAR(D0_FULL + d), (uint32_t)(i / H4_D0S) & 1u);
                tc_fence_after();
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
                tmem_st_wait();
                tc_fence_before();
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
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + H4_D2 + b * 64;
            uint32_t v[64];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(D2_EMPTY + b));          // D2[b] drained: S2(i + 2) may overwrite it
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
