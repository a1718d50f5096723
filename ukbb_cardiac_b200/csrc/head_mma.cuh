// Fused head, version 2: the bilinear upsampling runs ON THE TENSOR CORES.
//
// network.py:207-229 computes  fc0( concat_l( up_l(s_l) ) ).  Both the fixed bilinear transposed
// convolution up_l (network.py:138-167) and the 1x1 convolution fc0 are linear, and they act on
// different axes (pixels vs channels), so they commute:
//     W_fc0 . concat_l(up_l(s_l)) = W_0 . s_0 + sum_{l=1..4} up_l( W_l . s_l ),      W_fc0 = [W_0 | ... | W_4]
// t_l = W_l . s_l is a 32->64 1x1 convolution at the LOW resolution of level l (conv_tc_kernel, no
// BN / ReLU), and up_l restricted to one tile of 8 x 16 output pixels is a constant matrix
// U_l [128 pixels x S_l source pixels] (products of the dyadic bilinear weights, exact in BF16 /
// FP16; source pixels outside the image are zero-filled by TMA = the tapered borders of the
// transposed convolution).  The fc0 accumulator of a tile is therefore
//     D1[128 x 64] = s0_tile[128 x 32] . W_0^T  +  sum_l U_l[128 x S_l] . t_l_patch[S_l x 64]
// i.e. 2 + (3 + 2 + 1 + 1) UMMAs.  The t_l patches are used exactly as TMA writes them (pixel-major
// rows of 64 channels = an "MN-major" B operand, measured in profiles/r1_umma_probe_mn.log), so
// no thread ever touches the upsampled tensor: FLOPs executed on the tensor pipe for the upsample +
// fc0 stage are 2*128*64*(32 + 48 + 32 + 16 + 16) per tile instead of 2*128*64*160, and the
// 4-tap gather on CUDA cores disappears.  (Declared shortcut, SURVEY 8d: algorithmic FLOPs for the
// roofline stay 3.1374 GFLOP / slice.)
//
// The rest is as in head_fused.cuh: BN + ReLU -> 16-bit A operand in shared memory -> fc1 GEMM ->
// BN + ReLU (FP32) -> logits / softmax / argmax / crop / class counts on CUDA cores.
// Warps: 0-3 epilogue 1 (D1 -> A2), 4-7 epilogue 2 (D2 -> labels), 8 TMA producer, 9 MMA issuer.
#pragma once
#include "tc_common.cuh"
#include "head_fused.cuh"      // HeadParams

namespace ukbb {

constexpr int HM_THREADS = 320;
constexpr int HM_PW[5] = {0, 9, 6, 4, 3};          // source patch width  (x) per level
constexpr int HM_PH[5] = {0, 5, 4, 3, 3};          // source patch height (y) per level
constexpr int HM_KPAD[5] = {32, 64, 32, 16, 16};   // padded K of the U_l matrices (level 0: s0 channels)
constexpr int HM_KSTEPS[5] = {2, 3, 2, 1, 1};      // UMMA K-steps (16 each) actually issued
// shared-memory layout (bytes)
constexpr int HM_IN_S0 = 4096;                                    // b0 tile [128 pixels][16 ch = 32 B] (input of same_dim0)
constexpr int HM_IN_P1 = 48 * 128, HM_IN_P2 = 32 * 128, HM_IN_P3 = 16 * 128, HM_IN_P4 = 16 * 128;
constexpr int HM_IN_BYTES = HM_IN_S0 + HM_IN_P1 + HM_IN_P2 + HM_IN_P3 + HM_IN_P4;
constexpr int HM_IN_TX = HM_IN_S0 + (45 + 24 + 12 + 9) * 128;     // bytes TMA actually delivers per tile
constexpr int HM_A0 = 128 * 64;                                   // same_dim0 output tile as fc0 A operand [128][64 B]
constexpr int HM_A2 = 128 * 128;
constexpr int HM_U1 = 128 * 128, HM_U2 = 128 * 64, HM_U3 = 128 * 32, HM_U4 = 2 * 128 * 32;
constexpr int HM_U_BYTES = HM_U1 + HM_U2 + HM_U3 + HM_U4;         // 36864
constexpr int HM_W0 = 64 * 64, HM_W1 = 64 * 128, HM_WSD = 1024;     // fc0 level-0 slice, fc1, same_dim0 [32][16]
constexpr int HM_IN_STAGES = 3;
constexpr int HM_SMEM = HM_IN_STAGES * HM_IN_BYTES + 2 * HM_A0 + 2 * HM_A2 + HM_U_BYTES + HM_W0 + HM_W1 + HM_WSD + 1024 + 256 +
                        (4 * 64 + 64 * 8 + 8 + 64) * 4;

struct HeadMmaMaps {
    CUtensorMap s0 /* b0: conv0 output, 16 ch */, t1, t2, t3, t4, u1, u2, u3, u4, w0, w1, wsd;
};

template <int NC, bool F16>
__global__ void __launch_bounds__(HM_THREADS, 1)
head_mma_kernel(const __grid_constant__ HeadMmaMaps maps, const HeadParams p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t in_base = smem_base;                           // 2 stages of inputs
    const uint32_t a0_base = in_base + HM_IN_STAGES * HM_IN_BYTES;   // 2 x 8 KB
    const uint32_t a2_base = a0_base + 2 * HM_A0;                 // 2 x 16 KB
    const uint32_t u_base = a2_base + 2 * HM_A2;
    const uint32_t w0_base = u_base + HM_U_BYTES;
    const uint32_t w1_base = w0_base + HM_W0;
    const uint32_t wsd_base = w1_base + HM_W1;
    const uint32_t bar_base = wsd_base + HM_WSD;
    auto BAR = [&](int i) { return bar_base + 8u * i; };
    // 0 wfull | 1,2 in_full | 3,4 in_empty | 5,6 d1_full | 7,8 d1_empty | 9,10 a2_full | 11,12 a2_empty |
    // 13,14 d2_full | 15,16 d2_empty | 17,18 d0_full | 19,20 d0_empty | 21,22 a0_full | 23,24 a0_empty | 25 tmem slot
    // 26,27,28 in_full (3 input stages) | 29,30,31 in_empty      (slots 1-4 unused)
    const uint32_t tmem_slot = BAR(25);
    float* s_f = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 256);
    float* s_sc0 = s_f; float* s_sh0 = s_f + 64; float* s_sc1 = s_f + 128; float* s_sh1 = s_f + 192;
    float* s_wl = s_f + 256;
    float* s_bl = s_f + 256 + 512;
    float* s_scd = s_f + 256 + 512 + 8;       // same_dim0 BN fold: scale[32], shift[32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 8 && lane == 0) {
        const CUtensorMap* m = &maps.s0;
        for (int i = 0; i < 12; ++i) tma_prefetch_desc(m + i);
    }
    if (warp == 9 && lane == 0) {
        mbar_init(BAR(0), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(1 + b), 1);  mbar_init(BAR(3 + b), 1);  mbar_init(BAR(5 + b), 1);  mbar_init(BAR(7 + b), 4);
            mbar_init(BAR(9 + b), 4);  mbar_init(BAR(11 + b), 1); mbar_init(BAR(13 + b), 1); mbar_init(BAR(15 + b), 4);
            mbar_init(BAR(17 + b), 1); mbar_init(BAR(19 + b), 4); mbar_init(BAR(21 + b), 4); mbar_init(BAR(23 + b), 1);
        }
        for (int s3 = 0; s3 < HM_IN_STAGES; ++s3) { mbar_init(BAR(26 + s3), 1); mbar_init(BAR(29 + s3), 1); }
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // zero the input stages once: the K-padding rows of the t_l patches are never written by TMA and
    // must be finite (they meet zero columns of U_l)
    for (int i = threadIdx.x; i < HM_IN_STAGES * HM_IN_BYTES / 16; i += HM_THREADS)
        reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 64; i += HM_THREADS) {
        s_sc0[i] = p.scale0[i]; s_sh0[i] = p.shift0[i]; s_sc1[i] = p.scale1[i]; s_sh1[i] = p.shift1[i];
    }
    for (int i = threadIdx.x; i < 64 * 8; i += HM_THREADS) {
        const int k = i >> 3, c = i & 7;
        s_wl[i] = c < p.nc ? p.wlog[k * p.nc + c] : 0.f;
    }
    if (threadIdx.x < 8) s_bl[threadIdx.x] = threadIdx.x < p.nc ? p.blog[threadIdx.x] : -INFINITY;
    if (threadIdx.x < 32) { s_scd[threadIdx.x] = p.scale_sd0[threadIdx.x]; s_scd[32 + threadIdx.x] = p.shift_sd0[threadIdx.x]; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int tiles_per_slice = p.tiles_x * p.tiles_y;
    const int my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(0), HM_U_BYTES + HM_W0 + HM_W1 + HM_WSD);
            tma_load_2d(wsd_base, &maps.wsd, BAR(0), 0, 0);
            tma_load_2d(u_base, &maps.u1, BAR(0), 0, 0);
            tma_load_2d(u_base + HM_U1, &maps.u2, BAR(0), 0, 0);
            tma_load_2d(u_base + HM_U1 + HM_U2, &maps.u3, BAR(0), 0, 0);
            tma_load_2d(u_base + HM_U1 + HM_U2 + HM_U3, &maps.u4, BAR(0), 0, 0);
            tma_load_2d(u_base + HM_U1 + HM_U2 + HM_U3 + 128 * 32, &maps.u4, BAR(0), 0, 128);
            tma_load_2d(w0_base, &maps.w0, BAR(0), 0, 0);
            tma_load_2d(w1_base, &maps.w1, BAR(0), 0, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = blockIdx.x + i * gridDim.x;
                const int s3 = i % HM_IN_STAGES;
                const uint32_t ph = (uint32_t)(i / HM_IN_STAGES) & 1u;
                const int n = tile / tiles_per_slice, t2 = tile - n * tiles_per_slice;
                const int y0 = (t2 / p.tiles_x) * 8, x0 = (t2 % p.tiles_x) * 16;
                mbar_wait(BAR(29 + s3), ph ^ 1);
                const uint32_t dst = in_base + s3 * HM_IN_BYTES;
                const uint32_t fullb = BAR(26 + s3);
                mbar_arrive_expect_tx(fullb, HM_IN_TX);
                tma_load_4d(dst, &maps.s0, fullb, 0, x0, y0, n);
                // level l patch origin: ((x0 + pb) >> l) - 1, pb = (2^l - 1) / 2
                tma_load_4d(dst + HM_IN_S0, &maps.t1, fullb, 0, (x0 >> 1) - 1, (y0 >> 1) - 1, n);
                tma_load_4d(dst + HM_IN_S0 + HM_IN_P1, &maps.t2, fullb, 0, ((x0 + 1) >> 2) - 1, ((y0 + 1) >> 2) - 1, n);
                tma_load_4d(dst + HM_IN_S0 + HM_IN_P1 + HM_IN_P2, &maps.t3, fullb, 0, ((x0 + 3) >> 3) - 1, ((y0 + 3) >> 3) - 1, n);
                tma_load_4d(dst + HM_IN_S0 + HM_IN_P1 + HM_IN_P2 + HM_IN_P3, &maps.t4, fullb, 0, ((x0 + 7) >> 4) - 1,
                            ((y0 + 7) >> 4) - 1, n);
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t idesc_kk = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        const uint32_t idesc_kmn = idesc_kk | (1u << 16);            // B operand MN-major (pixel-major t_l patch)
        constexpr uint32_t HI32 = (uint32_t)((8 * 32) >> 4) | (1u << 14) | (6u << 29);
        constexpr uint32_t HI64 = (uint32_t)((8 * 64) >> 4) | (1u << 14) | (4u << 29);
        constexpr uint32_t HI128 = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);
        auto LO = [](uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | (1u << 16); };
        const uint32_t u1_lo = LO(u_base), u2_lo = LO(u_base + HM_U1), u3_lo = LO(u_base + HM_U1 + HM_U2),
                       u4_lo = LO(u_base + HM_U1 + HM_U2 + HM_U3);
        const uint32_t w0_lo = LO(w0_base), w1_lo = LO(w1_base), wsd_lo = LO(wsd_base);
        const uint32_t idesc_sd = F16 ? make_idesc_f16(128, 32) : make_idesc_bf16(128, 32);
        mbar_wait(BAR(0), 0);
        tc_fence_after();
        // same_dim0 (network.py:201-204, level 0): D0[128 x 32] = b0_tile[128 x 16] . Wsd0^T
        auto issue0 = [&](int i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const int s3 = i % HM_IN_STAGES;
            mbar_wait(BAR(26 + s3), (uint32_t)(i / HM_IN_STAGES) & 1u);
            mbar_wait(BAR(19 + b), ph ^ 1);
            tc_fence_after();
            if (leader) {
                umma_bf16_lohi(tmem_base + 256 + b * 32, LO(in_base + s3 * HM_IN_BYTES), HI32, wsd_lo, HI32, idesc_sd, 0u);
                umma_commit(BAR(17 + b));
            }
            __syncwarp();
        };
        auto issue1 = [&](int i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const int tile = blockIdx.x + i * gridDim.x;
            const int t2 = tile % tiles_per_slice;
            const uint32_t v = (uint32_t)((t2 / p.tiles_x) & 1);                 // tile-row parity selects the U_4 variant
            mbar_wait(BAR(21 + b), ph);
            mbar_wait(BAR(7 + b), ph ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + b * 64;
            const int s3 = i % HM_IN_STAGES;
            const uint32_t in_lo = LO(in_base + s3 * HM_IN_BYTES);
            const uint32_t a0_lo = LO(a0_base + b * HM_A0);
            if (leader) {
                // level 0: same_dim0 tile (written by epilogue 0, K-major, 64 B rows) x W_0 (K-major)
                umma_bf16_lohi(d, a0_lo, HI64, w0_lo, HI64, idesc_kk, 0u);
                umma_bf16_lohi(d, a0_lo + 2, HI64, w0_lo + 2, HI64, idesc_kk, 1u);
                // level 1: U_1 [128][64] (128 B rows), 3 K-steps; B = patch rows 16k.. (128 B per source pixel)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    umma_bf16_lohi(d, u1_lo + 2 * k, HI128, in_lo + ((HM_IN_S0 + k * 2048) >> 4), HI128, idesc_kmn, 1u);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_bf16_lohi(d, u2_lo + 2 * k, HI64, in_lo + ((HM_IN_S0 + HM_IN_P1 + k * 2048) >> 4), HI128, idesc_kmn, 1u);
                umma_bf16_lohi(d, u3_lo, HI32, in_lo + ((HM_IN_S0 + HM_IN_P1 + HM_IN_P2) >> 4), HI128, idesc_kmn, 1u);
                umma_bf16_lohi(d, u4_lo + v * ((128 * 32) >> 4), HI32, in_lo + ((HM_IN_S0 + HM_IN_P1 + HM_IN_P2 + HM_IN_P3) >> 4), HI128,
                               idesc_kmn, 1u);
                umma_commit(BAR(29 + s3));
                umma_commit(BAR(23 + b));
                umma_commit(BAR(5 + b));
            }
            __syncwarp();
        };
        auto issue2 = [&](int i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(BAR(9 + b), ph);
            mbar_wait(BAR(15 + b), ph ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + 128 + b * 64;
            const uint32_t a_lo = LO(a2_base + b * HM_A2);
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, a_lo + 2 * k, HI128, w1_lo + 2 * k, HI128, idesc_kk, k != 0 ? 1u : 0u);
                umma_commit(BAR(11 + b));
                umma_commit(BAR(13 + b));
            }
            __syncwarp();
        };
        if (my_tiles > 0) issue0(0);
        if (my_tiles > 1) issue0(1);
        if (my_tiles > 0) issue1(0);
        for (int i = 0; i < my_tiles; ++i) {
            if (i + 2 < my_tiles) issue0(i + 2);
            if (i + 1 < my_tiles) issue1(i + 1);
            issue2(i);
        }
    } else if (warp < 4) {
        // ===================== epilogue 1: D1 -> BN + ReLU -> 16-bit A2 (128 B swizzle) =====================
        const int q = warp;
        const int r = q * 32 + lane;
        // epilogue 0: D0 -> same_dim0 BN + ReLU -> 16-bit A0 tile (64 B rows, 64-byte swizzle)
        auto ep0 = [&](int i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(BAR(17 + b), ph);
            mbar_wait(BAR(23 + b), ph ^ 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + b * 32;
            const uint32_t row = a0_base + b * HM_A0 + r * 64;
#pragma unroll
            for (int c = 0; c < 32; c += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c, v);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_scd + c + 4 * j4);
                    const float4 sh = *reinterpret_cast<const float4*>(s_scd + 32 + c + 4 * j4);
                    o[2 * j4] = pack16t<F16>(fmaxf(fmaf(__uint_as_float(v[4 * j4]), sc.x, sh.x), 0.f),
                                             fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y), 0.f));
                    o[2 * j4 + 1] = pack16t<F16>(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), 0.f),
                                                 fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w), 0.f));
                }
                const uint32_t j0 = (uint32_t)(c >> 3), sw = ((uint32_t)r >> 1) & 3u;
                const uint32_t d0 = row + ((j0 ^ sw) << 4), d1 = row + (((j0 + 1) ^ sw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d1), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(BAR(19 + b)); mbar_arrive(BAR(21 + b)); }
        };
        if (my_tiles > 0) ep0(0);
        for (int i = 0; i < my_tiles; ++i) {
            if (i + 1 < my_tiles) ep0(i + 1);
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(BAR(5 + b), ph);
            mbar_wait(BAR(11 + b), ph ^ 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 64;
            const uint32_t row = a2_base + b * HM_A2 + r * 128;
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c, v);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_sc0 + c + 4 * j4);
                    const float4 sh = *reinterpret_cast<const float4*>(s_sh0 + c + 4 * j4);
                    o[2 * j4] = pack16t<F16>(fmaxf(fmaf(__uint_as_float(v[4 * j4]), sc.x, sh.x), 0.f),
                                             fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y), 0.f));
                    o[2 * j4 + 1] = pack16t<F16>(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), 0.f),
                                                 fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w), 0.f));
                }
                const uint32_t j0 = (uint32_t)(c >> 3);
                const uint32_t d0 = row + ((j0 ^ ((uint32_t)r & 7u)) << 4), d1 = row + (((j0 + 1) ^ ((uint32_t)r & 7u)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d1), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(BAR(7 + b)); mbar_arrive(BAR(9 + b)); }
        }
    } else {
        // ===================== epilogue 2 (warps 4-7): D2 -> BN + ReLU -> logits -> labels =====================
        const int q = warp - 4;
        const int r = q * 32 + lane;
        const int ty = r >> 4, tx = r & 15;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const int n = tile / tiles_per_slice, t2 = tile - n * tiles_per_slice;
            const int y = (t2 / p.tiles_x) * 8 + ty, x = (t2 % p.tiles_x) * 16 + tx;
            mbar_wait(BAR(13 + b), ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 128 + b * 64;
            float lg[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) lg[c] = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_sc1 + c0 + 4 * j4);
                    const float4 sh = *reinterpret_cast<const float4*>(s_sh1 + c0 + 4 * j4);
                    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * j4 + u;
                        const float fv = fmaxf(fmaf(__uint_as_float(v[j]), scv[u], shv[u]), 0.f);
                        const float4 wa = *reinterpret_cast<const float4*>(s_wl + (c0 + j) * 8);
                        const float4 wb = *reinterpret_cast<const float4*>(s_wl + (c0 + j) * 8 + 4);
                        const float wr[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int c = 0; c < NC; ++c) lg[c] = fmaf(fv, wr[c], lg[c]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(15 + b));
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < NC; ++c) { lg[c] += s_bl[c]; mx = fmaxf(mx, lg[c]); }
            float e[NC], ssum = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) { e[c] = expf(lg[c] - mx); ssum += e[c]; }
            float best = -1.f;
            int arg = 0;
            const size_t pix = ((size_t)n * p.h + y) * p.w + x;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float pr = e[c] / ssum;
                if (pr > best) { best = pr; arg = c; }
                if (c < p.nc) {
                    if (p.prob) p.prob[pix * p.nc + c] = pr;
                    if (p.logits) p.logits[pix * p.nc + c] = lg[c];
                }
            }
            const int yy = y - p.y_pre, xx = x - p.x_pre;
            const bool inside = yy >= 0 && yy < p.y && xx >= 0 && xx < p.x;
            if (inside) p.labels[((size_t)n * p.y + yy) * p.x + xx] = (uint8_t)arg;
            if (p.counts) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const unsigned bal = __ballot_sync(0xffffffffu, inside && arg == c);
                    if (lane == 0 && bal && c < p.nc) atomicAdd(&p.counts[(size_t)n * p.nc + c], (unsigned long long)__popc(bal));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace ukbb
