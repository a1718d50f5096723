// Side branches of build_FCN for levels 1..4 in ONE persistent launch (north_star (b)):
//     s_l = relu(bn(W_sd_l . x_l))                    same_dim 1x1 conv, Cin_l -> 32     (network.py:201-204)
//     t_l = W_l . s_l                                 the level-l column block of fc0, 32 -> 64, moved in front
//                                                     of the (linear) upsampling (head_mma.cuh, network.py:207-227)
// A 1x1 convolution has no spatial structure, so a level is a plain [pixels][Cin] matrix and a
// tile is 128 consecutive pixels.  Per tile:
//     S0  D0[128x32] = X[128 x Cin] . W_sd^T          Cin/16 UMMAs, input streamed in 64-channel chunks   MMA warp 1
//     E0  A1 = relu(D0 * scale + shift) -> 16 bit, K-major 64 B rows in shared memory                     warps 4-7
//     S1  D1[128x64] = A1 . W_l^T                     2 UMMAs                                              MMA warp 2
//     E1  t = D1 -> 16 bit -> swizzled staging tile -> ONE TMA store (coalesced 16 KB)                     warps 8-11
// s_l never reaches HBM (it was written and re-read by two conv_tc launches per level before), and
// the four levels share one launch: tiles are ordered level 4, 3, 2, 1 and dealt round-robin, so the
// few deep-level tiles (large K) overlap with the many shallow ones.  All weights (46 KB) are resident.
// HBM traffic per SA slice: 0.5 MB read + 0.85 MB written; the kernel is bound by that, not by the tensor pipe.
//
// SPLIT (x3 modes): inputs, weights, s_l and t_l are (hi, lo) pairs of 16-bit values and every product is hi.hi + lo.hi + hi.lo.
// Shared memory then holds both weight sets (96 KB) and a ring of four (hi, lo) input slots, so A1 = s_l moves to TENSOR MEMORY
// (tcgen05.st, [tmem] A operand of S1, as in head_ts.cuh) and t_l is written with direct 256-bit stores: one pixel's 64 channels
// are 128 contiguous bytes per plane and consecutive threads own consecutive pixels.
#pragma once
#include "tc_common.cuh"
#include "conv_group.cuh"      // tmem_ld32, bn_relu_pack, bulk-group helpers

namespace ukbb {

struct SideMaps {
    CUtensorMap in[4];          // [pixels][Cin_l], box 128 x 64 channels (32 for level 1); index = level - 1
    CUtensorMap out[4];         // t_l: [pixels][64], box 128 x 64
    CUtensorMap wsd[4];         // same_dim weights [32][Cin_l], box 32 x 64 (32 for level 1)
    CUtensorMap w0;             // fc0 weights [64][160] (BN scale folded in), box 64 x 32
};

struct SideParams {
    int tile_start[5];          // k-th level in processing order (level 4 - k) owns tiles [tile_start[k], tile_start[k+1])
    const float* scale[4];      // same_dim BN scale / shift, index = level - 1
    const float* shift[4];
    // split modes
    int rows[4];                // pixels of level l in this call (index = level - 1)
    int lo_row[4];              // row of the lo plane in the input map of level l (= pixels of the plan's capacity)
    uint32_t* t[4];             // t_l hi plane [pixels][64] as 32-bit words
    long long t_lo[4];          // offset of the lo plane in 32-bit words
};

constexpr int SD_THREADS = 384;
constexpr int SD_SLOT = 128 * 128;                      // one 64-channel chunk of 128 pixels
constexpr int SD_SLOTS = 7;
constexpr int SD_SLOTS_SPLIT = 4;                       // (hi, lo) slot pairs
__host__ __device__ constexpr int sd_wsd_off(int l) { return l == 1 ? 0 : 4096 << (l - 2); }   // offsets 0, 4 K, 8 K, 16 K.  W_sd_l: 32 rows x Cin_l, as Cin_l/64 chunks of [32][128 B] (level 1: [32][64 B])
constexpr int SD_WSD_BYTES = 32768;
constexpr int SD_WL_BYTES = 4 * 4096;                   // W_l: [64][32] each
constexpr int SD_A1 = 128 * 64, SD_OUT = 128 * 128;
constexpr int SD_SMEM = SD_SLOTS * SD_SLOT + SD_WSD_BYTES + SD_WL_BYTES + 2 * SD_A1 + 2 * SD_OUT + 1024 /*align*/ + 256 /*barriers*/ +
                        4 * 64 * 4 /*scale, shift*/;
constexpr int SD_SMEM_SPLIT = SD_SLOTS_SPLIT * 2 * SD_SLOT + 2 * (SD_WSD_BYTES + SD_WL_BYTES) + 1024 + 256 + 4 * 64 * 4;
static_assert(SD_SMEM_SPLIT <= 227 * 1024, "shared memory budget");

namespace tc {
__device__ __forceinline__ void tma_store_2d(const void* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
// two FP32 values -> 16 bit, no activation (FP16: clamped to the finite range)
template <bool F16>
__device__ __forceinline__ uint32_t plain_pack(uint32_t a0, uint32_t a1) {
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(a1)), "f"(__uint_as_float(a0)));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(a1)), "f"(__uint_as_float(a0)));
    return r;
}
// two FP32 values -> hi and lo 16-bit pieces, no activation
template <bool F16>
__device__ __forceinline__ void plain_split(uint32_t a0, uint32_t a1, uint32_t& hi, uint32_t& lo) {
    hi = plain_pack<F16>(a0, a1);
    const float2 h = unpack16t<F16>(hi);
    lo = plain_pack<F16>(__float_as_uint(__uint_as_float(a0) - h.x), __float_as_uint(__uint_as_float(a1) - h.y));
}
}  // namespace tc

// F8 (x2 scheme, tc_common.cuh): the lo planes of the INPUTS and of the same_dim weights hold FP8 correction operands.  S0 streams its K
// chunks, so the "corrections first, then rescale" order of the convolution kernels is not available: the FP8 products go to a second
// accumulator (D0C) and E0 adds D0C * 2^-15.  s_l (A1) and the fc0 block stay (hi, lo) FP16 pairs: A1 is produced in this kernel.
template <bool F16, bool SPLIT = false, bool F8 = false>
__global__ void __launch_bounds__(SD_THREADS, 1)
side_tc_kernel(const __grid_constant__ SideMaps maps, const __grid_constant__ SideParams p) {
    using namespace tc;
    constexpr int SLOTS = SPLIT ? SD_SLOTS_SPLIT : SD_SLOTS, SLOT_BYTES = (SPLIT ? 2 : 1) * SD_SLOT, PL = SPLIT ? 2 : 1;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring_base = smem_base;
    const uint32_t wsd_base = ring_base + SLOTS * SLOT_BYTES;            // hi set | lo set
    const uint32_t wl_base = wsd_base + PL * SD_WSD_BYTES;               // hi set | lo set
    const uint32_t a1_base = wl_base + PL * SD_WL_BYTES;                 // (not SPLIT) s_l tile as the A operand of S1
    const uint32_t out_base = a1_base + (SPLIT ? 0 : 2 * SD_A1);         // (not SPLIT) staging tiles of the TMA store
    const uint32_t bar_base = out_base + (SPLIT ? 0 : 2 * SD_OUT);
    auto BAR = [&](int i) { return bar_base + 8u * i; };
    enum { WFULL = 0, IN_FULL = 1, IN_EMPTY = IN_FULL + SD_SLOTS, D0_FULL = IN_EMPTY + SD_SLOTS, D0_EMPTY = D0_FULL + 2,   // (SLOTS <= SD_SLOTS)
           A1_FULL = D0_EMPTY + 2, A1_EMPTY = A1_FULL + 2, D1_FULL = A1_EMPTY + 2, D1_EMPTY = D1_FULL + 2, TSLOT = D1_EMPTY + 2 };
    static_assert((TSLOT + 1) * 8 <= 256, "barrier area");
    const uint32_t tmem_slot = BAR(TSLOT);
    float* s_scale = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 256);      // [4][32]
    float* s_shift = s_scale + 128;
    constexpr int D0_COL = 0, D1_COL = 64, A1_COL = 192;  // TMEM columns: 2 x 32, 2 x 64, (SPLIT) A1 2 x (16 hi + 16 lo)
    constexpr int D0C_COL = 256, TMEM_COLS = F8 ? 512 : 256;             // (F8) correction accumulators 2 x 32

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = p.tile_start[4];

    if (warp == 0 && lane == 0) {
        const CUtensorMap* m = &maps.in[0];
        for (int i = 0; i < 13; ++i) tma_prefetch_desc(m + i);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(BAR(WFULL), 1);
        for (int s = 0; s < SLOTS; ++s) { mbar_init(BAR(IN_FULL + s), 1); mbar_init(BAR(IN_EMPTY + s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(D0_FULL + b), 1); mbar_init(BAR(D0_EMPTY + b), 4);
            mbar_init(BAR(A1_FULL + b), 4); mbar_init(BAR(A1_EMPTY + b), 1);
            mbar_init(BAR(D1_FULL + b), 1); mbar_init(BAR(D1_EMPTY + b), 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    if (warp == 3) {
        for (int i = lane; i < 128; i += 32) { s_scale[i] = p.scale[i >> 5][i & 31]; s_shift[i] = p.shift[i >> 5][i & 31]; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 0) griddep_wait();                       // the producer waits after it has issued the weight loads
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto LO = [](uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | (1u << 16); };
    constexpr uint32_t HI64 = (uint32_t)((8 * 64) >> 4) | (1u << 14) | (4u << 29);
    constexpr uint32_t HI128 = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);
    // level (1..4) of a tile: processing order is 4, 3, 2, 1
    auto level_of = [&](int tile) { return 4 - ((tile >= p.tile_start[1]) + (tile >= p.tile_start[2]) + (tile >= p.tile_start[3])); };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(WFULL), PL * (32 * 64 + 32 * 128 * 7 + SD_WL_BYTES));
            for (int pl = 0; pl < PL; ++pl) {                // lo weights = rows [cout, 2 cout) of the same maps
                tma_load_2d(wsd_base + pl * SD_WSD_BYTES + sd_wsd_off(1), &maps.wsd[0], BAR(WFULL), 0, 32 * pl);
                for (int l = 2; l <= 4; ++l)
                    for (int c = 0; c < (1 << (l - 2)); ++c)
                        tma_load_2d(wsd_base + pl * SD_WSD_BYTES + sd_wsd_off(l) + c * 4096, &maps.wsd[l - 1], BAR(WFULL), c * 64, 32 * pl);
                for (int l = 1; l <= 4; ++l) tma_load_2d(wl_base + pl * SD_WL_BYTES + (l - 1) * 4096, &maps.w0, BAR(WFULL), 32 * l, 64 * pl);
            }
            griddep_wait();
            int slot = 0;
            uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = blockIdx.x + i * gridDim.x;
                const int l = level_of(tile);
                const int row0 = (tile - p.tile_start[4 - l]) * 128;
                const int chunks = l == 1 ? 1 : 1 << (l - 2);
                const uint32_t bytes = l == 1 ? 128 * 64 : 128 * 128;
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(BAR(IN_EMPTY + slot), ph ^ 1);
                    mbar_arrive_expect_tx(BAR(IN_FULL + slot), PL * bytes);
                    tma_load_2d(ring_base + slot * SLOT_BYTES, &maps.in[l - 1], BAR(IN_FULL + slot), c * 64, row0);
                    if (SPLIT) tma_load_2d(ring_base + slot * SLOT_BYTES + SD_SLOT, &maps.in[l - 1], BAR(IN_FULL + slot), c * 64, p.lo_row[l - 1] + row0);
                    if (++slot == SLOTS) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer 0: same_dim (S0) =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, 32) : make_idesc_bf16(128, 32);
        const uint32_t idesc8 = make_idesc_e4m3(128, 32);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        int slot = 0;
        uint32_t ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int l = level_of(tile);
            const int chunks = l == 1 ? 1 : 1 << (l - 2);
            const int ksteps = l == 1 ? 2 : 4;
            const uint32_t hi = l == 1 ? HI64 : HI128;
            const int b = i & 1;
            mbar_wait(BAR(D0_EMPTY + b), (((uint32_t)i >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + D0_COL + b * 32;
            const uint32_t w_lo = LO(wsd_base + sd_wsd_off(l));
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(BAR(IN_FULL + slot), ph);
                tc_fence_after();
                const uint32_t a_lo = LO(ring_base + slot * SLOT_BYTES);
                if (leader) {
                    for (int k = 0; k < ksteps; ++k) {
                        umma_bf16_lohi(d, a_lo + 2 * k, hi, w_lo + c * (4096 >> 4) + 2 * k, hi, idesc, (c | k) != 0 ? 1u : 0u);
                        if (F8) {
                            umma_f8_lohi(d + (D0C_COL - D0_COL), a_lo + (SD_SLOT >> 4) + 2 * k, hi, w_lo + (SD_WSD_BYTES >> 4) + c * (4096 >> 4) + 2 * k, hi,
                                         idesc8, (c | k) != 0 ? 1u : 0u);
                        } else if (SPLIT) {
                            umma_bf16_lohi(d, a_lo + (SD_SLOT >> 4) + 2 * k, hi, w_lo + c * (4096 >> 4) + 2 * k, hi, idesc, 1u);              // lo . hi
                            umma_bf16_lohi(d, a_lo + 2 * k, hi, w_lo + (SD_WSD_BYTES >> 4) + c * (4096 >> 4) + 2 * k, hi, idesc, 1u);        // hi . lo
                        }
                    }
                    umma_commit(BAR(IN_EMPTY + slot));
                    if (c == chunks - 1) umma_commit(BAR(D0_FULL + b));
                }
                __syncwarp();
                if (++slot == SLOTS) { slot = 0; ph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ===================== MMA issuer 1: fc0 column block (S1) =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        mbar_wait(BAR(WFULL), 0);
        tc_fence_after();
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int l = level_of(tile);
            const int b = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            mbar_wait(BAR(D1_EMPTY + b), ph ^ 1u);
            mbar_wait(BAR(A1_FULL + b), ph);
            tc_fence_after();
            const uint32_t d = tmem_base + D1_COL + b * 64;
            const uint32_t a_lo = LO(a1_base + b * SD_A1), w_lo = LO(wl_base + (l - 1) * 4096);
            if (leader) {
                if (SPLIT) {                                    // A1 = (hi 16 columns | lo 16 columns) in tensor memory
                    const uint32_t a1 = tmem_base + A1_COL + b * 32;
                    for (int k = 0; k < 2; ++k) {
                        umma_ts_lohi(d, a1 + 8 * k, w_lo + 2 * k, HI64, idesc, k != 0 ? 1u : 0u);
                        umma_ts_lohi(d, a1 + 16 + 8 * k, w_lo + 2 * k, HI64, idesc, 1u);
                        umma_ts_lohi(d, a1 + 8 * k, w_lo + (SD_WL_BYTES >> 4) + 2 * k, HI64, idesc, 1u);
                    }
                } else {
                umma_bf16_lohi(d, a_lo, HI64, w_lo, HI64, idesc, 0u);
                umma_bf16_lohi(d, a_lo + 2, HI64, w_lo + 2, HI64, idesc, 1u);
                }
                umma_commit(BAR(A1_EMPTY + b));
                umma_commit(BAR(D1_FULL + b));
            }
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== E0: D0 -> BN + ReLU -> A1 =====================
        const int q = warp - 4;
        const int r = q * 32 + lane;
        float4 scv[8], shv[8];                               // (SPLIT) folded BN of the level's 32 channels in registers, reloaded when the level changes
        int cur_l = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int l = level_of(tile);
            const int b = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            if (SPLIT && l != cur_l) {
                cur_l = l;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    scv[j] = *reinterpret_cast<const float4*>(s_scale + (l - 1) * 32 + 4 * j);
                    shv[j] = *reinterpret_cast<const float4*>(s_shift + (l - 1) * 32 + 4 * j);
                }
            }
            mbar_wait(BAR(D0_FULL + b), ph);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + D0_COL + b * 32, v);
            if (F8) {
                uint32_t vc[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + D0C_COL + b * 32, vc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(vc[j]), 1.f / 32768.f, __uint_as_float(v[j])));
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(D0_EMPTY + b));
            mbar_wait(BAR(A1_EMPTY + b), ph ^ 1u);          // S1 of tile i - 2 has consumed A1[b]
            const float* sc = s_scale + (l - 1) * 32;
            const float* sh = s_shift + (l - 1) * 32;
            if (SPLIT) {
                tc_fence_after();
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t oh[8], ol[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) bn_relu_split4<F16, false>(v + 16 * hf + j, scv[4 * hf + j / 4], shv[4 * hf + j / 4], oh, ol, j / 4);
                    tmem_st8(tmem_base + ((uint32_t)(q * 32) << 16) + A1_COL + b * 32 + 8 * hf, oh);
                    tmem_st8(tmem_base + ((uint32_t)(q * 32) << 16) + A1_COL + b * 32 + 16 + 8 * hf, ol);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(A1_FULL + b));
                continue;
            }
            const uint32_t row = a1_base + b * SD_A1 + r * 64;
            const uint32_t sw = ((uint32_t)r >> 1) & 3u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t o[4];
                const float4 sc0 = *reinterpret_cast<const float4*>(sc + 8 * j), sc1 = *reinterpret_cast<const float4*>(sc + 8 * j + 4);
                const float4 sh0 = *reinterpret_cast<const float4*>(sh + 8 * j), sh1 = *reinterpret_cast<const float4*>(sh + 8 * j + 4);
                o[0] = bn_relu_pack<F16>(v[8 * j], v[8 * j + 1], make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
                o[1] = bn_relu_pack<F16>(v[8 * j + 2], v[8 * j + 3], make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
                o[2] = bn_relu_pack<F16>(v[8 * j + 4], v[8 * j + 5], make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
                o[3] = bn_relu_pack<F16>(v[8 * j + 6], v[8 * j + 7], make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
                const uint32_t dst = row + (((uint32_t)j ^ sw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A1_FULL + b));
        }
    } else if (warp >= 8) {
        // ===================== E1: D1 -> 16 bit -> staging -> TMA store =====================
        const int q = warp - 8;
        const int r = q * 32 + lane;
        const bool issuer = threadIdx.x == 256;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int l = level_of(tile);
            const int row0 = (tile - p.tile_start[4 - l]) * 128;
            const int b = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            mbar_wait(BAR(D1_FULL + b), ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + D1_COL + b * 64;
            uint32_t v[64];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(D1_EMPTY + b));
            if (SPLIT) {
                const int prow = row0 + r;
                uint32_t* dst = p.t[l - 1] + (size_t)prow * 32;
                const bool live = prow < p.rows[l - 1];
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint32_t oh[8], ol[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) plain_split<F16>(v[16 * c8 + 2 * j], v[16 * c8 + 2 * j + 1], oh[j], ol[j]);
                    if (live) { stg256(dst + 8 * c8, oh); stg256(dst + p.t_lo[l - 1] + 8 * c8, ol); }
                }
                continue;
            }
            // staging set b was read by the store of tile i - 2, which the issuer waited for before the last barrier
            const uint32_t row = out_base + b * SD_OUT + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t o0 = plain_pack<F16>(v[8 * j], v[8 * j + 1]), o1 = plain_pack<F16>(v[8 * j + 2], v[8 * j + 3]);
                const uint32_t o2 = plain_pack<F16>(v[8 * j + 4], v[8 * j + 5]), o3 = plain_pack<F16>(v[8 * j + 6], v[8 * j + 7]);
                const uint32_t dst = row + ((uint32_t)(j ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
            }
            fence_proxy_async();
            if (issuer) bulk_wait_read<0>();                 // store of tile i - 1 has left its staging set
            named_bar_sync(2, 128);
            if (issuer) {
                tma_store_2d(&maps.out[l - 1], out_base + b * SD_OUT, 0, row0);
                bulk_commit();
            }
        }
        if (issuer && !SPLIT) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace ukbb
