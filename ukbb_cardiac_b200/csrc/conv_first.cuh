// conv0_0 + conv0_1 of build_FCN in ONE launch (north_star (a); network.py:19-25, :184-188):
//     a0 = relu(bn(conv3x3(image,  1 -> 16)))      FP32 on the CUDA cores (packed FFMA2), input stays FP32
//     b0 = relu(bn(conv3x3(a0,    16 -> 16)))      tcgen05, pixel-group rows (conv_group.cuh, N = 4 px x 16 ch = 64)
// Before, conv0_0 was its own kernel: it wrote a0 (1.28 MB / slice, 16-bit NHWC) to HBM and
// conv_group<16,16,1> read it back -- 164 MB each way per 128-slice sub-batch, more than L2 holds, so both
// launches ran at the HBM roofline (87.6 + 57.4 us, profiles/r1_launch_summary_stage5_side.txt).
// Here a0 never leaves the SM: "builder" warps compute the 18 x 40-pixel halo patch of a0 that one
// conv0_1 tile (16 rows x 32 pixels) needs, straight into the 128-byte-swizzled shared-memory rows the UMMA
// descriptors of conv_group read ([patch row][group of 4 pixels][4 px x 16 ch], same addresses as the TMA
// box load it replaces).  HBM traffic per slice drops from 4.0 MB to 1.44 MB (FP32 image in, b0 out).
//   warp 0        TMA: FP32 image boxes (20 rows x 48 columns, zero fill outside the image = SAME padding of
//                 conv0_0) into a 4-deep ring; conv0_1 weights once
//   warps 8-15    builders: one thread = one patch group (4 pixels x 16 channels): 18 image values, 9 taps x
//                 4 px x 8 FFMA2 with the weights read as broadcast LDS.128, ReLU + 16-bit pack, 8 swizzled
//                 STS.128.  180 groups per tile = 6 warp-rounds, dealt round-robin over the 8 warps.
//                 Patch groups outside the image are written as ZEROS (SAME padding of conv0_1 pads a0, not
//                 the image).
//   warp 1        MMA issuer: 18 UMMAs (M = 128, N = 64, K = 16) per tile, as conv_group<16,16,1>
//   warps 4-7     epilogue: tcgen05.ld, folded BN + ReLU, 16-bit pack, swizzled staging tile, one TMA store
#pragma once
#include "tc_common.cuh"
#include "conv_group.cuh"

namespace ukbb {

struct ConvFirstParams {
    int tiles_x, tiles_y, n_tiles;
    int h, w4;                      // image rows; image columns / 4 (pixel groups per row)
    const float* scale;             // conv0_1 folded BN [16]
    const float* shift;
    float w0[9][16];                // conv0_0: [tap = dy * 3 + dx][cout] FP32, BN scale folded in
    float shift0[16];
};

struct ConvFirstCfg {
    static constexpr int PU = 10, PR = 18, J = 6, N = 64;
    static constexpr int GROUPS = PU * PR;                      // 180 patch groups per tile
    static constexpr int ROUNDS = (GROUPS + 31) / 32;           // 6 warp-rounds
    static constexpr int PATCH_BYTES = (GROUPS * 128 + 1023) / 1024 * 1024;
    static constexpr int A_STAGES = 4;
    static constexpr int B_TILE = 2048;                         // [64 rows][16 cin] 16-bit
    static constexpr int NB_TILES = 3 * J;
    static constexpr int B_BYTES = NB_TILES * B_TILE;
    static constexpr int OUT_BYTES = 128 * 128;
    static constexpr int IMG_W = 48, IMG_H = 20;                // FP32 box: columns x0 - 8 .. x0 + 39, rows y0 - 2 .. y0 + 17 (TMA needs the
                                                                // box origin 16-byte aligned in the inner dimension: experiments/tma_probe_img.cu)
    static constexpr int IMG_TX = IMG_W * IMG_H * 4;
    static constexpr int IMG_BYTES = (IMG_TX + 127) / 128 * 128;
    static constexpr int IMG_STAGES = 4;
    static constexpr int BUILDERS = 8;                          // builder warps
    static constexpr int THREADS = 256 + 32 * BUILDERS;
    static constexpr int TMEM_COLS = 128;
    static constexpr int SMEM_BYTES = A_STAGES * PATCH_BYTES + B_BYTES + 2 * OUT_BYTES + IMG_STAGES * IMG_BYTES + 1024 /*w0, shift0*/ +
                                      256 /*barriers*/ + 2 * N * 4 + 1024 /*align*/;
};

namespace tc {
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t dup2(float v) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}
template <bool F16>
__device__ __forceinline__ uint32_t relu_pack2(uint64_t v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
}  // namespace tc

template <bool F16>
__global__ void __launch_bounds__(ConvFirstCfg::THREADS, 1)
conv_first_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_out, const __grid_constant__ ConvFirstParams p) {
    using namespace tc;
    using Cfg = ConvFirstCfg;
    constexpr int AST = Cfg::A_STAGES, IST = Cfg::IMG_STAGES, J = Cfg::J, PU = Cfg::PU, N = Cfg::N;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t b_base = smem_base + AST * Cfg::PATCH_BYTES;
    const uint32_t out_base = b_base + Cfg::B_BYTES;
    const uint32_t img_base = out_base + 2 * Cfg::OUT_BYTES;
    const uint32_t w0_base = img_base + IST * Cfg::IMG_BYTES;
    const uint32_t bar_base = w0_base + 1024;
    // barriers: a_full[AST] a_empty[AST] img_full[IST] img_empty[IST] tfull[2] tempty[2] wfull | tmem slot
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (AST + s); };
    auto img_full = [&](int s) { return bar_base + 8u * (2 * AST + s); };
    auto img_empty = [&](int s) { return bar_base + 8u * (2 * AST + IST + s); };
    auto tfull = [&](int a) { return bar_base + 8u * (2 * AST + 2 * IST + a); };
    auto tempty = [&](int a) { return bar_base + 8u * (2 * AST + 2 * IST + 2 + a); };
    const uint32_t wfull = bar_base + 8u * (2 * AST + 2 * IST + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * AST + 2 * IST + 5);
    static_assert((2 * AST + 2 * IST + 6) * 8 <= 256, "barrier area");
    float* s_w0 = reinterpret_cast<float*>(smem_gen + (w0_base - smem_base));                  // [9][16] then shift0[16]
    float* s_scale = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 256);        // [N] expanded (column -> channel)
    float* s_shift = s_scale + N;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_img); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_out); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < AST; ++s) { mbar_init(a_full(s), Cfg::ROUNDS); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < IST; ++s) { mbar_init(img_full(s), 1); mbar_init(img_empty(s), Cfg::ROUNDS); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    if (warp == 3) {
        for (int c = lane; c < N; c += 32) { s_scale[c] = p.scale[c & 15]; s_shift[c] = p.shift[c & 15]; }
        for (int c = lane; c < 160; c += 32) s_w0[c] = c < 144 ? p.w0[c >> 4][c & 15] : p.shift0[c - 144];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 0) griddep_wait();                       // the producer waits after it has issued the weight loads
    const int my_tiles = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0) {
        // ===================== TMA producer: conv0_1 weights once, one FP32 image box per tile =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(wfull, Cfg::NB_TILES * Cfg::B_TILE);
            for (int t = 0; t < Cfg::NB_TILES; ++t) tma_load_2d(b_base + t * Cfg::B_TILE, &map_b, wfull, 0, t * N);
            griddep_wait();
            TileWalk w;
            w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
            int is = 0;
            uint32_t iph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                mbar_wait(img_empty(is), iph ^ 1);
                mbar_arrive_expect_tx(img_full(is), Cfg::IMG_TX);
                tma_load_3d(img_base + is * Cfg::IMG_BYTES, &map_img, img_full(is), w.tx * 32 - 8, w.ty * 16 - 2, w.n);
                if (++is == IST) { is = 0; iph ^= 1; }
                w.next();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (conv0_1, as conv_group<16, 16, 1>) =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, N) : make_idesc_bf16(128, N);
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
            const uint32_t d = tmem_base + acc * N;
            const uint32_t a_lo = (((smem_base + as * Cfg::PATCH_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
            if (leader) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int jj = j - 1;                                    // input pixel of this K-slice relative to the group
                        const int ro = jj < 0 ? -1 : jj / 4;
                        const int sub = jj - ro * 4;
                        const int arow = ky * PU + 1 + ro;
                        umma_bf16_lohi(d, a_lo + ((arow * 128 + sub * 32) >> 4), a_hi, b_lo + (((ky * J + j) * Cfg::B_TILE) >> 4), b_hi, idesc,
                                       (ky | j) != 0 ? 1u : 0u);
                    }
                umma_commit(a_empty(as));
                umma_commit(tfull(acc));
            }
            __syncwarp();
            if (++as == AST) { as = 0; aph ^= 1; }
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    } else if (warp >= 8) {
        // ===================== builders: conv0_0 on the CUDA cores, straight into the UMMA patch layout =====================
        const int bw = warp - 8;
        TileWalk w;
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        int wi = 0;                                                              // tile the walker stands on
        const int total_rounds = my_tiles * Cfg::ROUNDS;
        for (int R = bw; R < total_rounds; R += Cfg::BUILDERS) {
            const int i = R / Cfg::ROUNDS, r = R - i * Cfg::ROUNDS;
            while (wi < i) { w.next(); ++wi; }
            const int as = i % AST, is = i % IST;
            const int g = r * 32 + lane;                                         // patch group = patch row py, group pg
            const bool active = g < Cfg::GROUPS;
            const int py = g / PU, pg = g - py * PU;
            mbar_wait(img_full(is), (uint32_t)(i / IST) & 1u);
            float v[3][6];
            if (active) {
                // group pg needs image columns x0 - 5 + 4 pg .. + 5 = box columns 4 pg + 3 .. 4 pg + 8
                const float* src = reinterpret_cast<const float*>(smem_gen + (img_base - smem_base) + is * Cfg::IMG_BYTES) + py * Cfg::IMG_W + 4 * pg;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float4 a = *reinterpret_cast<const float4*>(src + ky * Cfg::IMG_W + 4);
                    v[ky][0] = src[ky * Cfg::IMG_W + 3];
                    v[ky][1] = a.x; v[ky][2] = a.y; v[ky][3] = a.z; v[ky][4] = a.w;
                    v[ky][5] = src[ky * Cfg::IMG_W + 8];
                }
            } else {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int c = 0; c < 6; ++c) v[ky][c] = 0.f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(img_empty(is));                           // image values are in registers
            uint64_t acc[4][8];
            {
                const ulonglong2* sh = reinterpret_cast<const ulonglong2*>(s_w0 + 144);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const ulonglong2 s2 = sh[q];
#pragma unroll
                    for (int px = 0; px < 4; ++px) { acc[px][2 * q] = s2.x; acc[px][2 * q + 1] = s2.y; }
                }
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int ky = t / 3, kx = t - 3 * ky;
                uint64_t wp[8];
                const ulonglong2* wt = reinterpret_cast<const ulonglong2*>(s_w0 + t * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) { const ulonglong2 w2 = wt[q]; wp[2 * q] = w2.x; wp[2 * q + 1] = w2.y; }
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const uint64_t a = dup2(v[ky][px + kx]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[px][j] = ffma2(a, wp[j], acc[px][j]);
                }
            }
            // SAME padding of conv0_1 pads a0 with zeros: groups outside the image are zero, not conv0_0 of a zero image
            const int y = w.ty * 16 - 1 + py, gx = w.tx * 8 - 1 + pg;
            const bool inside = y >= 0 && y < p.h && gx >= 0 && gx < p.w4;
            mbar_wait(a_empty(as), ((uint32_t)(i / AST) & 1u) ^ 1u);
            if (active) {
                const uint32_t row = smem_base + as * Cfg::PATCH_BYTES + (uint32_t)g * 128u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int px = c >> 1, hf = (c & 1) * 4;
                    uint32_t o0 = relu_pack2<F16>(acc[px][hf]), o1 = relu_pack2<F16>(acc[px][hf + 1]), o2 = relu_pack2<F16>(acc[px][hf + 2]),
                             o3 = relu_pack2<F16>(acc[px][hf + 3]);
                    if (!inside) { o0 = 0u; o1 = 0u; o2 = 0u; o3 = 0u; }
                    const uint32_t dst = row + (((uint32_t)c ^ ((uint32_t)g & 7u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(as));
        }
    } else if (warp >= 4) {
        // ===================== epilogue (conv0_1): as conv_group =====================
        const int q = warp - 4;
        const int r = q * 32 + lane;                         // TMEM lane = tile row * 8 + group
        const bool issuer = threadIdx.x == 128;
        TileWalk w;
        w.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * N;
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
