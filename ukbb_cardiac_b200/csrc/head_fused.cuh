// Fused head of build_FCN on tcgen05 (north_star (b) + (c)):
//   bilinear transposed-conv upsampling of the four coarse 32-channel maps (network.py:138-167,
//   207-211), channel concat (:214-218), 1x1 conv 160->64 + BN + ReLU, 1x1 conv 64->64 + BN + ReLU
//   (:227-228), 1x1 conv 64->n_class + bias (:229), softmax, argmax (train_network.py:198-199),
//   crop to the un-padded image (deploy_network.py:114-116) and per-slice class counts (:127-130)
// in ONE kernel.  The 160-channel tensor and both 64-channel tensors never touch global memory.
//
// One tile = 128 full-resolution pixels (8 rows x 16 columns of one slice).  Per tile:
//   * warp 4 TMA-loads the level-0 chunk (32 channels) straight into the A operand buffer;
//   * warps 6-13 (256 threads) compute the 4-tap bilinear values of levels 1-4 on CUDA cores and
//     write them into the same K-major, 64-byte-swizzled A buffer (chunks 1-4) -- the concat is
//     just the K order of the GEMM;
//   * warp 5 issues  D1[128x64] = A[128x160] . Wfc0^T   (10 UMMAs, TMEM accumulator)
//   * warps 0-3 read D1 from TMEM, apply BN scale/shift + ReLU, round to 16 bit and write the
//     result as the 128-byte-swizzled A operand of the next GEMM (shared memory only);
//   * warp 5 issues  D2[128x64] = A2[128x64] . Wfc1^T   (4 UMMAs)
//   * warps 0-3 read D2, apply BN + ReLU in FP32 and finish on CUDA cores: logits (64 x n_class
//     FMAs per pixel, FP32 features), softmax, argmax (first maximal index), crop, counts.
// All buffers (A, A2, D1, D2) are double-buffered so tile i+1's gather / first GEMM overlap tile
// i's epilogues.  Weights (fc0: 20 KB, fc1: 8 KB) are loaded into shared memory once per CTA.
#pragma once
#include "tc_common.cuh"

namespace ukbb {

struct HeadParams {
    const __nv_bfloat16* s[5];      // same_dim outputs, level l: [n][h>>l][w>>l][32]
    int n, h, w;                    // slices, padded rows (Y2), padded columns (X2)
    int tiles_x, tiles_y, n_tiles;
    int x_pre, y_pre, x, y;         // crop
    int fp16, nc;
    const float* scale0; const float* shift0;     // fc0 BN fold
    const float* scale1; const float* shift1;     // fc1 BN fold
    const float* scale_sd0; const float* shift_sd0;   // same_dim0 BN fold (head_mma only)
    const float* wlog;              // [64][nc] fp32
    const float* blog;              // [nc]
    uint8_t* labels;                // [n][y][x]
    float* logits;                  // optional [n][h][w][nc]
    float* prob;                    // optional
    unsigned long long* counts;     // optional [n][nc]
    // head_tc only: BN scales are folded into the 16-bit weights, the shifts travel by value so that the
    // epilogues read them as constant-bank operands (no shared-memory loads, no registers)
    float c_shift_sd0[32], c_shift0[64], c_shift1[64], c_bias[8];
    float2 c_wl2[32][8];            // class-score weights FP32 as input-channel pairs: [k / 2][class] = (w[k][c], w[k + 1][c])
    const uint32_t* u_glob[5];      // head_ts only: interpolation matrices U_l in global memory, rows of HM_KPAD[l] 16-bit values
    const uint4* b0;                // head_ts only: conv0_1 output [n][h][w][16] 16-bit (level-0 input of same_dim0)
    float c_nshift1[64];            // head_ts only: -shift of fc1;  relu(d + s) . w = max(d, -s) . w + s . w
    float c_bias2[8];               //   bias + sum_k shift1[k] * w[k][c]  (-inf for c >= n_class)
    float c_wlc[64][8];             //   class-score weights [k][class], zero for c >= n_class
    int dbg;                        // experiments (UKBB_HEAD_DBG): 1 short E2, 2 no patch loads, 4 no U terms, 8 no b0 loads, 16 timeline trace
    long long* trace;               // [12 events][64 tiles] clock64 of CTA 0 (dbg & 16)
};

constexpr int HEAD_THREADS = 448;
constexpr int HEAD_A_BYTES = 5 * 8192;          // 5 chunks x [128 rows][64 B]
constexpr int HEAD_A2_BYTES = 128 * 128;        // [128 rows][128 B]
constexpr int HEAD_W0_BYTES = 5 * 4096;         // fc0: 5 chunks x [64 rows][64 B]
constexpr int HEAD_W1_BYTES = 64 * 128;         // fc1: [64 rows][128 B]
constexpr int HEAD_SMEM = 2 * HEAD_A_BYTES + 2 * HEAD_A2_BYTES + HEAD_W0_BYTES + HEAD_W1_BYTES + 1024 /*align*/ +
                          256 /*barriers*/ + (4 * 64 + 64 * 8 + 8) * 4 /*scale/shift, wlog, blog*/;

template <int NC, bool F16>
__global__ void __launch_bounds__(HEAD_THREADS, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap map_s0, const __grid_constant__ CUtensorMap map_w0,
                  const __grid_constant__ CUtensorMap map_w1, const HeadParams p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t a_base = smem_base;                               // 2 x 40 KB
    const uint32_t a2_base = a_base + 2 * HEAD_A_BYTES;              // 2 x 16 KB
    const uint32_t w0_base = a2_base + 2 * HEAD_A2_BYTES;            // 20 KB
    const uint32_t w1_base = w0_base + HEAD_W0_BYTES;                // 8 KB
    const uint32_t bar_base = w1_base + HEAD_W1_BYTES;
    auto BAR = [&](int i) { return bar_base + 8u * i; };
    // 0 wfull | 1,2 a0_full | 3,4 ag_full | 5,6 a_empty | 7,8 d1_full | 9,10 d1_empty | 11,12 a2_full |
    // 13,14 a2_empty | 15,16 d2_full | 17,18 d2_empty | 19: tmem slot
    const uint32_t tmem_slot = BAR(19);
    float* s_f = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 256);
    float* s_sc0 = s_f; float* s_sh0 = s_f + 64; float* s_sc1 = s_f + 128; float* s_sh1 = s_f + 192;
    float* s_wl = s_f + 256;            // [64][8]
    float* s_bl = s_f + 256 + 512;      // [8]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 4 && lane == 0) { tma_prefetch_desc(&map_s0); tma_prefetch_desc(&map_w0); tma_prefetch_desc(&map_w1); }
    if (warp == 5 && lane == 0) {
        mbar_init(BAR(0), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(1 + b), 1);  mbar_init(BAR(3 + b), 8);  mbar_init(BAR(5 + b), 1);
            mbar_init(BAR(7 + b), 1);  mbar_init(BAR(9 + b), 4);  mbar_init(BAR(11 + b), 4);
            mbar_init(BAR(13 + b), 1); mbar_init(BAR(15 + b), 1); mbar_init(BAR(17 + b), 4);
        }
        fence_barrier_init();
    }
    if (warp == 6) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    for (int i = threadIdx.x; i < 64; i += HEAD_THREADS) {
        s_sc0[i] = p.scale0[i]; s_sh0[i] = p.shift0[i]; s_sc1[i] = p.scale1[i]; s_sh1[i] = p.shift1[i];
    }
    for (int i = threadIdx.x; i < 64 * 8; i += HEAD_THREADS) {
        const int k = i >> 3, c = i & 7;
        s_wl[i] = c < p.nc ? p.wlog[k * p.nc + c] : 0.f;
    }
    if (threadIdx.x < 8) s_bl[threadIdx.x] = threadIdx.x < p.nc ? p.blog[threadIdx.x] : -INFINITY;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int tiles_per_slice = p.tiles_x * p.tiles_y;
    const int my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(0), HEAD_W0_BYTES + HEAD_W1_BYTES);
            for (int c = 0; c < 5; ++c) tma_load_2d(w0_base + c * 4096, &map_w0, BAR(0), c * 32, 0);
            tma_load_2d(w1_base, &map_w1, BAR(0), 0, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = blockIdx.x + i * gridDim.x;
                const int b = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                const int n = tile / tiles_per_slice, t2 = tile - n * tiles_per_slice;
                const int y0 = (t2 / p.tiles_x) * 8, x0 = (t2 % p.tiles_x) * 16;
                mbar_wait(BAR(5 + b), ph ^ 1);
                mbar_arrive_expect_tx(BAR(1 + b), 8192);
                tma_load_4d(a_base + b * HEAD_A_BYTES, &map_s0, BAR(1 + b), 0, x0, y0, n);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (whole warp runs the control flow, one elected lane issues) ============
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, 64) : make_idesc_bf16(128, 64);
        constexpr uint32_t HI64 = (uint32_t)((8 * 64) >> 4) | (1u << 14) | (4u << 29);      // SWIZZLE_64B, SBO = 8 rows
        constexpr uint32_t HI128 = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);    // SWIZZLE_128B
        const uint32_t w0_lo = ((w0_base & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t w1_lo = ((w1_base & 0x3FFFF) >> 4) | (1u << 16);
        mbar_wait(BAR(0), 0);
        tc_fence_after();
        auto issue1 = [&](int i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(BAR(1 + b), ph);
            mbar_wait(BAR(3 + b), ph);
            mbar_wait(BAR(9 + b), ph ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + b * 64;
            const uint32_t a_lo = (((a_base + b * HEAD_A_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
            if (leader) {
#pragma unroll
                for (int c = 0; c < 5; ++c)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_bf16_lohi(d, a_lo + ((c * 8192 + k * 32) >> 4), HI64, w0_lo + ((c * 4096 + k * 32) >> 4), HI64, idesc,
                                       (c | k) != 0 ? 1u : 0u);
                umma_commit(BAR(5 + b));
                umma_commit(BAR(7 + b));
            }
            __syncwarp();
        };
        auto issue2 = [&](int i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(BAR(11 + b), ph);
            mbar_wait(BAR(17 + b), ph ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + 128 + b * 64;
            const uint32_t a_lo = (((a2_base + b * HEAD_A2_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_lohi(d, a_lo + ((k * 32) >> 4), HI128, w1_lo + ((k * 32) >> 4), HI128, idesc, k != 0 ? 1u : 0u);
                umma_commit(BAR(13 + b));
                umma_commit(BAR(15 + b));
            }
            __syncwarp();
        };
        if (my_tiles > 0) issue1(0);
        for (int i = 0; i < my_tiles; ++i) {
            if (i + 1 < my_tiles) issue1(i + 1);
            issue2(i);
        }
    } else if (warp >= 6) {
        // ===================== bilinear gather (levels 1-4 -> A chunks 1-4) =====================
        const int g = threadIdx.x - 192;            // 0..255
        const int m = g & 127;                      // pixel row of the tile
        const int half = g >> 7;                    // 0: levels 1,2   1: levels 3,4
        const int ty = m >> 4, tx = m & 15;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const int n = tile / tiles_per_slice, t2 = tile - n * tiles_per_slice;
            const int y = (t2 / p.tiles_x) * 8 + ty, x = (t2 % p.tiles_x) * 16 + tx;
            mbar_wait(BAR(5 + b), ph ^ 1);
#pragma unroll 1
            for (int li = 0; li < 2; ++li) {
                const int l = 1 + half * 2 + li;
                const int f = 1 << l, pb = (f - 1) >> 1;
                const int hl = p.h >> l, wl = p.w >> l;
                const int ry = (y + pb) & (f - 1), y1 = (y + pb) >> l, y0 = y1 - 1;
                const int rx = (x + pb) & (f - 1), x1 = (x + pb) >> l, x0 = x1 - 1;
                const float inv = 1.f / (float)f;
                const float wy1 = (float)(ry + 1) * inv, wy0 = 1.f - wy1;
                const float wx1 = (float)(rx + 1) * inv, wx0 = 1.f - wx1;
                float acc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0.f;
                const __nv_bfloat16* base = p.s[l] + (size_t)n * hl * wl * 32;
                auto tap = [&](int yy, int xx, float wgt) {
                    if (yy < 0 || yy >= hl || xx < 0 || xx >= wl || wgt == 0.f) return;
                    const uint4* src = reinterpret_cast<const uint4*>(base + ((size_t)yy * wl + xx) * 32);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint4 v = __ldg(src + q4);
                        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f2 = unpack16t<F16>(u[j]);
                            acc[q4 * 8 + 2 * j] = fmaf(f2.x, wgt, acc[q4 * 8 + 2 * j]);
                            acc[q4 * 8 + 2 * j + 1] = fmaf(f2.y, wgt, acc[q4 * 8 + 2 * j + 1]);
                        }
                    }
                };
                tap(y0, x0, wy0 * wx0); tap(y0, x1, wy0 * wx1); tap(y1, x0, wy1 * wx0); tap(y1, x1, wy1 * wx1);
                // row m of chunk l: 64 bytes = 4 x 16 B, 64-byte swizzle: chunk j -> j ^ ((m >> 1) & 3)
                const uint32_t row = a_base + b * HEAD_A_BYTES + l * 8192 + m * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t dst = row + (((uint32_t)j ^ ((uint32_t)(m >> 1) & 3u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                                 "r"(pack16t<F16>(acc[8 * j], acc[8 * j + 1])), "r"(pack16t<F16>(acc[8 * j + 2], acc[8 * j + 3])),
                                 "r"(pack16t<F16>(acc[8 * j + 4], acc[8 * j + 5])), "r"(pack16t<F16>(acc[8 * j + 6], acc[8 * j + 7]))
                                 : "memory");
                }
            }
            fence_proxy_async();                    // generic-proxy smem writes -> visible to UMMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(3 + b));
        }
    } else {
        // ===================== epilogues (warps 0-3) =====================
        const int q = warp;
        const int r = q * 32 + lane;
        const int ty = r >> 4, tx = r & 15;
        auto ep1 = [&](int i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(BAR(7 + b), ph);
            mbar_wait(BAR(13 + b), ph ^ 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 64;
            const uint32_t row = a2_base + b * HEAD_A2_BYTES + r * 128;
#pragma unroll 1
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
                // 128-byte swizzle: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
                const uint32_t j0 = (uint32_t)(c >> 3);
                const uint32_t d0 = row + ((j0 ^ ((uint32_t)r & 7u)) << 4), d1 = row + (((j0 + 1) ^ ((uint32_t)r & 7u)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d1), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(BAR(9 + b)); mbar_arrive(BAR(11 + b)); }
        };
        auto ep2 = [&](int i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const int n = tile / tiles_per_slice, t2 = tile - n * tiles_per_slice;
            const int y = (t2 / p.tiles_x) * 8 + ty, x = (t2 % p.tiles_x) * 16 + tx;
            mbar_wait(BAR(15 + b), ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 128 + b * 64;
            float lg[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) lg[c] = 0.f;
#pragma unroll 1
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
            if (lane == 0) mbar_arrive(BAR(17 + b));          // D2 drained: the next GEMM may overwrite it
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
        };
        if (my_tiles > 0) ep1(0);
        for (int i = 0; i < my_tiles; ++i) {
            if (i + 1 < my_tiles) ep1(i + 1);
            ep2(i);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 6) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace ukbb
