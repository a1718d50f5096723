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
#include "engine.cuh"
#include "tc_common.cuh"
#include "conv_halo.cuh"
#include "conv_group.cuh"
#include "conv_first.cuh"
#include "conv_first_tc.cuh"
#include "head_fused.cuh"
#include "head_mma.cuh"
#include "head_tc.cuh"
#include "head_ts.cuh"
#include "side_tc.cuh"
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <string.h>

namespace ukbb {

using namespace tc;

struct ConvTcParams {
    int taps, ks, stride, cin;
    int kofs;                       // first K column of the weight matrix (sub-matrix selection)
    int pad_top, pad_left;
    int bw, bh, bn;                 // output box of one tile: bw * bh * bn == 128
    int tiles_x, tiles_y, n_tiles;
    int ho, wo, n;                  // output height / width / slices actually valid
    int relu;
    int fp16;                       // operand format: 0 = BF16, 1 = FP16
    const float* scale;
    const float* shift;
    __nv_bfloat16* out;             // [n][ho][wo][COUT]
};

template <int CC, int COUT>
struct ConvTcCfg {
    static constexpr int A_BYTES = 128 * CC * 2;
    static constexpr int B_BYTES = COUT * CC * 2;
    static constexpr int B_PAD = (B_BYTES + 1023) / 1024 * 1024;
    static constexpr int STAGE_BYTES = A_BYTES + B_PAD;           // both 1024-aligned
    static constexpr int MAX_STAGES = (200 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
    static constexpr int TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;   // power of two for COUT in {16..256}
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int CC, int COUT, bool F16>
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const ConvTcParams p) {
    using Cfg = ConvTcCfg<CC, COUT>;
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
                for (int kb = 0; kb < kblocks; ++kb) {
                    const int tap = kb / chunks, ch = kb - tap * chunks;
                    const int ky = tap / p.ks, kx = tap - ky * p.ks;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), Cfg::A_BYTES + Cfg::B_BYTES);
                    tma_load_4d(a_dst, &map_a, full_bar(stage), ch * CC, x0 + kx, y0 + ky, n0);
                    tma_load_2d(b_dst, &map_b, full_bar(stage), p.kofs + tap * p.cin + ch * CC, 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, COUT) : make_idesc_bf16(128, COUT);
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
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_lo = (((smem_base + stage * Cfg::STAGE_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
                const uint32_t b_lo = a_lo + (Cfg::A_BYTES >> 4);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < CC / 16; ++k)
                        umma_bf16_lohi(d, a_lo + 2 * k, HI, b_lo + 2 * k, HI, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar(stage));          // frees the smem stage when the MMAs retire
                    if (kb == kblocks - 1) umma_commit(tfull_bar(acc));   // accumulator complete -> epilogue
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
                uint32_t o[8];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c) + j4);
                    const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + c) + j4);
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
// conv0_0 (Cin = 1, K = 9): FP32 image in, 16-bit NHWC out, CUDA cores (11.5 MFLOP / slice; the
// kernel is bound by its 1.3 MB / slice output write).  One thread = one pixel x 16 channels.
// The 144 weights (BN scale folded in, FP32) and the 16 shifts travel BY VALUE in the kernel
// parameters, so every FFMA takes its weight as a constant-bank operand: the first version read
// them from shared memory (144 broadcast LDS per pixel) and ran at a third of the HBM write rate.
// ------------------------------------------------------------------------------------------
struct Conv0Params {
    float w[9][16];                 // [tap = dy * 3 + dx][cout], scale folded in
    float shift[16];
};

template <bool F16>
__global__ void __launch_bounds__(256)
conv0_bf16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, long long total, int h, int w,
                  const __grid_constant__ Conv0Params cp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % w), y = (int)((idx / w) % h);
    const float* base = img + idx;
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            v[ky * 3 + kx] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(base + (ky - 1) * w + (kx - 1)) : 0.f;
        }
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = cp.shift[j];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(v[t], cp.w[t][j], acc[j]);
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(o[j]) : "f"(acc[2 * j + 1]), "f"(acc[2 * j]));
        else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o[j]) : "f"(acc[2 * j + 1]), "f"(acc[2 * j]));
    }
    uint4* d4 = reinterpret_cast<uint4*>(out + idx * 16);
    d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
    d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ------------------------------------------------------------------------------------------
// Bilinear upsample + concat in BF16 (same arithmetic as upsample_concat_fp32_kernel, FP32
// interpolation, one rounding to BF16).  One thread = one pixel x 8 channels of one level.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample_concat_bf16_kernel(const __nv_bfloat16* __restrict__ s0, const __nv_bfloat16* __restrict__ s1,
                            const __nv_bfloat16* __restrict__ s2, const __nv_bfloat16* __restrict__ s3,
                            const __nv_bfloat16* __restrict__ s4, __nv_bfloat16* __restrict__ out, long long total,
                            int h, int w, int fp16) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int q = (int)(idx % 20);
    const long long pix = idx / 20;
    const int x = (int)(pix % w), y = (int)((pix / w) % h);
    const long long n = pix / ((long long)w * h);
    const int l = q / 4, c8 = q % 4;
    uint4 r;
    if (l == 0) {
        r = reinterpret_cast<const uint4*>(s0 + ((n * h + y) * w + x) * 32)[c8];
    } else {
        const __nv_bfloat16* src = l == 1 ? s1 : l == 2 ? s2 : l == 3 ? s3 : s4;
        const int f = 1 << l, pb = (f - 1) / 2;
        const int hl = h >> l, wl = w >> l;
        const int ry = (y + pb) % f, y1 = (y + pb) / f, y0 = y1 - 1;
        const int rx = (x + pb) % f, x1 = (x + pb) / f, x0 = x1 - 1;
        const float wy1 = (float)(ry + 1) / (float)f, wy0 = 1.f - wy1;
        const float wx1 = (float)(rx + 1) / (float)f, wx0 = 1.f - wx1;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const __nv_bfloat16* base = src + n * hl * wl * 32;
        auto tap = [&](int yy, int xx, float wgt) {
            if (yy < 0 || yy >= hl || xx < 0 || xx >= wl || wgt == 0.f) return;
            const uint4 v = reinterpret_cast<const uint4*>(base + ((long long)yy * wl + xx) * 32)[c8];
            const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f2 = unpack16(h2[j], fp16);
                acc[2 * j] = fmaf(f2.x, wgt, acc[2 * j]);
                acc[2 * j + 1] = fmaf(f2.y, wgt, acc[2 * j + 1]);
            }
        };
        tap(y0, x0, wy0 * wx0); tap(y0, x1, wy0 * wx1); tap(y1, x0, wy1 * wx0); tap(y1, x1, wy1 * wx1);
        r = make_uint4(pack16(acc[0], acc[1], fp16), pack16(acc[2], acc[3], fp16), pack16(acc[4], acc[5], fp16),
                       pack16(acc[6], acc[7], fp16));
    }
    reinterpret_cast<uint4*>(out + pix * 160)[q] = r;
}

// ------------------------------------------------------------------------------------------
// Classifier on BF16 features: 1x1 conv 64 -> n_class (+bias), FP32 softmax, argmax (first
// maximal index), crop, per-slice class counts.  Same contract as classifier_fp32_kernel.
// ------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256)
classifier_bf16_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ wt,
                       const float* __restrict__ bias, int h2, int w2, int x_pre, int y_pre, int x, int y,
                       uint8_t* __restrict__ labels, float* __restrict__ logits, float* __restrict__ prob,
                       unsigned long long* __restrict__ counts, int fp16) {
    __shared__ float s_w[64 * NC];
    __shared__ float s_b[NC];
    for (int e = threadIdx.x; e < 64 * NC; e += blockDim.x) s_w[e] = wt[e];
    if (threadIdx.x < NC) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < h2 * w2;
    int label = -1;
    bool inside = false;
    if (live) {
        const uint4* f4 = reinterpret_cast<const uint4*>(feat + ((size_t)n * h2 * w2 + p) * 64);
        float lg[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) lg[c] = 0.f;
#pragma unroll 2
        for (int k8 = 0; k8 < 8; ++k8) {
            const uint4 v = __ldg(f4 + k8);
            const uint32_t* h2p = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 f2 = unpack16(h2p[u], fp16);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    lg[c] = fmaf(f2.x, s_w[(k8 * 8 + 2 * u) * NC + c], lg[c]);
                    lg[c] = fmaf(f2.y, s_w[(k8 * 8 + 2 * u + 1) * NC + c], lg[c]);
                }
            }
        }
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < NC; ++c) { lg[c] += s_b[c]; m = fmaxf(m, lg[c]); }
        float e[NC], s = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) { e[c] = expf(lg[c] - m); s += e[c]; }
        float best = -1.f;
        int arg = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float pr = e[c] / s;
            if (pr > best) { best = pr; arg = c; }
            if (prob) prob[((size_t)n * h2 * w2 + p) * NC + c] = pr;
            if (logits) logits[((size_t)n * h2 * w2 + p) * NC + c] = lg[c];
        }
        const int yy = p / w2 - y_pre, xx = p % w2 - x_pre;
        inside = yy >= 0 && yy < y && xx >= 0 && xx < x;
        if (inside) {
            labels[((size_t)n * y + yy) * x + xx] = (uint8_t)arg;
            label = arg;
        }
    }
    if (counts) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const unsigned b = __ballot_sync(0xffffffffu, inside && label == c);
            if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counts[(size_t)n * NC + c], (unsigned long long)__popc(b));
        }
    }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcLayerPlan {
    CUtensorMap map_a, map_b;
    ConvTcParams p;
    ConvHaloParams hp;
    ConvGroupParams gp;
    CUtensorMap map_out;
    int cc, cout;
    int kind = 0;                   // 0 = per-tap implicit GEMM (conv_tc_kernel), 1 = halo reuse (conv_halo_kernel),
                                    // 2 = pixel-group rows (conv_group_kernel)
    bool valid = false;
};

struct Bf16State {
    EncodeTiledFn encode = nullptr;
    __nv_bfloat16* w[UKBB_N_CONV] = {};      // [cout][taps*cin] bf16, K-major
    __nv_bfloat16* wg[UKBB_N_CONV] = {};     // pixel-group layers: expanded [3 * J tiles][64 rows][cin] (conv_group.cuh)
    Conv0Params c0;                          // conv0_0: FP32 weights with the BN scale folded in + shifts (kernel parameters)
    __nv_bfloat16* cat = nullptr;            // [nb][h][w][160]
    __nv_bfloat16* f0 = nullptr;             // [nb][h][w][64]
    __nv_bfloat16* f1 = nullptr;
    TcLayerPlan plan[UKBB_N_CONV];
    int plan_nb = 0, plan_h = 0, plan_w = 0;
    int fp16 = 0;
    int fused_head = 4;                      // 0 = unfused, 1 = gather head (head_fused), 2 = tensor-core upsample (head_mma),
                                             // 3 = head_mma algebra, 4-stage pipeline, constant-bank epilogues (head_tc),
                                             // 4 = head_tc with the A operands (U_l, A0, A2) in tensor memory (head_ts)
    __nv_bfloat16* wf[UKBB_N_CONV] = {};     // head_tc: weights of same_dim0 / fc0 / fc1 with the BN scale folded in before rounding
    float h_shift[UKBB_N_CONV][64] = {};     // host copies of the folded-BN shifts of those layers (constant-bank operands)
    float h_bias[8] = {};
    float h_wl[64 * 8] = {};                 // class-score weights [k][8] FP32
    CUtensorMap map_s0, map_w0, map_w1;      // fused head operands
    __nv_bfloat16* t[5] = {};                // t_l = W_l . s_l at level l (64 channels), l = 1..4
    __nv_bfloat16* u[5] = {};                // interpolation matrices U_l (16-bit, exact)
    float* ones = nullptr;                   // [64] ones then [64] zeros
    TcLayerPlan tplan[5];
    HeadMmaMaps hm;
    SideMaps sm;                             // side_tc_kernel: same_dim_l + fc0 column block of levels 1..4 in one launch
    int side = 1;
    int first = 2;                           // conv0_0 + conv0_1 in one launch: 1 = conv0_0 in FP32 on the CUDA cores (conv_first.cuh),
                                             // 2 = conv0_0 on the tensor pipe, hi/lo split of the FP32 image (conv_first_tc.cuh)
    __nv_bfloat16* wb0 = nullptr;            // conv_first_tc: expanded conv0_0 weights [64][64]
    CUtensorMap map_b0;
};

static CUtensorMapSwizzle swizzle_for(int cc) {
    return cc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : cc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

static int chunk_for(int cin) { return cin % 64 == 0 ? 64 : cin % 32 == 0 ? 32 : 16; }

static int make_plan(Engine* h, int li, const __nv_bfloat16* in, __nv_bfloat16* out, int nb, int hi, int wi,
                     int level_out) {
    Bf16State* S = h->tc;
    const ConvLayer& L = h->layers[li];
    TcLayerPlan& P = S->plan[li];
    const int cc = chunk_for(L.cin);
    P.cc = cc; P.cout = L.cout;
    const int s = L.stride, ks = L.ksize;
    const int ho = (hi + s - 1) / s, wo = (wi + s - 1) / s;
    int pt = (ho - 1) * s + ks - hi; if (pt < 0) pt = 0; pt /= 2;
    int pl = (wo - 1) * s + ks - wi; if (pl < 0) pl = 0; pl /= 2;
    // output box: bw | wo, bh | ho by construction (padded sizes are multiples of 16 at level 0)
    int bw = 16 >> level_out; if (bw < 1) bw = 1;
    int bh = 8; while (bh > 1 && (ho % bh != 0 || bw * bh > 128)) bh >>= 1;
    while (wo % bw != 0 && bw > 1) bw >>= 1;
    const int bn = 128 / (bw * bh);
    ConvTcParams& p = P.p;
    p.taps = ks * ks; p.ks = ks; p.stride = s; p.cin = L.cin; p.kofs = 0; p.pad_top = pt; p.pad_left = pl;
    p.bw = bw; p.bh = bh; p.bn = bn;
    p.tiles_x = wo / bw; p.tiles_y = ho / bh;
    p.ho = ho; p.wo = wo; p.n = nb; p.relu = L.relu; p.scale = L.scale; p.shift = L.shift; p.out = out;
    p.fp16 = S->fp16;
    const CUtensorMapDataType dt16 = S->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    p.n_tiles = p.tiles_x * p.tiles_y * ((nb + bn - 1) / bn);
    P.kind = (ks == 3 && s == 1 && L.cin >= 16 && !getenv("UKBB_NO_HALO")) ? 1 : 0;
    if (S->wg[li] && wi % (64 / L.cin) == 0 && !getenv("UKBB_NO_GROUP") && !getenv("UKBB_NO_HALO")) P.kind = 2;
    if (P.kind == 2) {
        const int g = 64 / L.cin, gout = s == 1 ? g : g / 2;
        const int pu = s == 1 ? 10 : 9, pr = s == 1 ? 18 : 33, jn = s == 1 ? g + 2 : g + 1;
        ConvGroupParams& gp = P.gp;
        gp.tiles_x = (wo / gout + 7) / 8; gp.tiles_y = (ho + 15) / 16; gp.n_tiles = gp.tiles_x * gp.tiles_y * nb;
        gp.scale = L.scale; gp.shift = L.shift;
        cuuint32_t e4[4] = {1, 1, 1, 1}, e2[2] = {1, 1};
        {   // input: rows of g pixels (128 bytes), box = halo patch
            cuuint64_t dims[4] = {64, (cuuint64_t)(wi / g), (cuuint64_t)hi, (cuuint64_t)nb};
            cuuint64_t strides[3] = {128, (cuuint64_t)wi * L.cin * 2, (cuuint64_t)hi * wi * L.cin * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)pu, (cuuint32_t)pr, 1};
            CUresult r = S->encode(&P.map_a, dt16, 4, (void*)in, dims, strides, box, e4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(group patch, layer %d) failed: %d", li, (int)r); return UKBB_E_CUDA; }
        }
        {   // expanded weights: [3 * J tiles * 64 rows][cin]
            cuuint64_t dims[2] = {(cuuint64_t)L.cin, (cuuint64_t)(3 * jn * 64)};
            cuuint64_t strides[1] = {(cuuint64_t)L.cin * 2};
            cuuint32_t box[2] = {(cuuint32_t)L.cin, 64};
            CUresult r = S->encode(&P.map_b, dt16, 2, (void*)S->wg[li], dims, strides, box, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   swizzle_for(L.cin), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(group weights, layer %d) failed: %d", li, (int)r); return UKBB_E_CUDA; }
        }
        {   // output: rows of gout pixels x cout channels = 64 elements, box = 16 rows x 8 groups
            cuuint64_t dims[4] = {64, (cuuint64_t)(wo / gout), (cuuint64_t)ho, (cuuint64_t)nb};
            cuuint64_t strides[3] = {128, (cuuint64_t)wo * L.cout * 2, (cuuint64_t)ho * wo * L.cout * 2};
            cuuint32_t box[4] = {64, 8, 16, 1};
            CUresult r = S->encode(&P.map_out, dt16, 4, (void*)out, dims, strides, box, e4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(group output, layer %d) failed: %d", li, (int)r); return UKBB_E_CUDA; }
        }
        P.valid = true;
        return UKBB_OK;
    }
    if (P.kind == 1) {
        ConvHaloParams& hp = P.hp;
        hp.cin = L.cin; hp.chunks = L.cin / cc;
        hp.tiles_x = (wo + 15) / 16; hp.tiles_y = (ho + 15) / 16; hp.n_tiles = hp.tiles_x * hp.tiles_y * nb;
        hp.ho = ho; hp.wo = wo; hp.n = nb; hp.relu = L.relu; hp.fp16 = S->fp16;
        hp.scale = L.scale; hp.shift = L.shift; hp.out = out;
        cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)wi, (cuuint64_t)hi, (cuuint64_t)nb};
        cuuint64_t strides[3] = {(cuuint64_t)L.cin * 2, (cuuint64_t)wi * L.cin * 2, (cuuint64_t)hi * wi * L.cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)cc, 18, 18, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = S->encode(&P.map_a, dt16, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle_for(cc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(halo patch, layer %d) failed: %d", li, (int)r); return UKBB_E_CUDA; }
    } else
    // activation map: dims (C, W, H, N), box (cc, bw*s, bh*s, bn), traversal strides (1, s, s, 1)
    {
        cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)wi, (cuuint64_t)hi, (cuuint64_t)nb};
        cuuint64_t strides[3] = {(cuuint64_t)L.cin * 2, (cuuint64_t)wi * L.cin * 2, (cuuint64_t)hi * wi * L.cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)cc, (cuuint32_t)(bw * s), (cuuint32_t)(bh * s), (cuuint32_t)bn};
        cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
        CUresult r = S->encode(&P.map_a, dt16, 4, (void*)in, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation, layer %d) failed: %d", li, (int)r); return UKBB_E_CUDA; }
    }
    {
        const int ktot = p.taps * L.cin;
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)L.cout};
        cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)cc, (cuuint32_t)L.cout};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = S->encode(&P.map_b, dt16, 2, (void*)S->w[li], dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights, layer %d) failed: %d", li, (int)r); return UKBB_E_CUDA; }
    }
    P.valid = true;
    return UKBB_OK;
}

// Launch with programmatic stream serialization (tc_common.cuh: griddep_launch / griddep_wait): the next kernel's
// CTAs start their prologue on an SM as soon as the previous kernel's CTA there has exited, instead of after the
// whole grid has drained.  UKBB_NO_PDL=1 falls back to plain stream order.
template <int CLUSTER = 1, typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
    static const bool pdl = getenv("UKBB_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (CLUSTER > 1) {                                   // thread-block clusters along x (TMA multicast of shared operands)
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = CLUSTER; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int CC, int COUT, bool F16>
static int launch_tc2(const TcLayerPlan& P, int sms, cudaStream_t st) {
    using Cfg = ConvTcCfg<CC, COUT>;
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CC, COUT, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int grid = P.p.n_tiles < sms ? P.p.n_tiles : sms;
    UKBB_CUDA(launch_pdl(conv_tc_kernel<CC, COUT, F16>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.p));
    return UKBB_OK;
}
template <int CC, int COUT>
static int launch_tc(const TcLayerPlan& P, int sms, cudaStream_t st) {
    return P.p.fp16 ? launch_tc2<CC, COUT, true>(P, sms, st) : launch_tc2<CC, COUT, false>(P, sms, st);
}

template <int CC, int COUT, bool RESIDENT, int NKB, bool F16>
static int launch_halo2(const TcLayerPlan& P, int sms, cudaStream_t st) {
    using Cfg = ConvHaloCfg<CC, COUT, RESIDENT, NKB>;
    const bool use_cluster = !RESIDENT && getenv("UKBB_NO_CLUSTER") == nullptr;
    if (!RESIDENT && use_cluster) {
        // clusters of two CTAs share every weight tile through TMA multicast (conv_halo.cuh)
        static bool attr_set_cl = false;
        if (!attr_set_cl) {
            UKBB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<CC, COUT, RESIDENT, NKB, F16, !RESIDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
            attr_set_cl = true;
        }
        const int pairs = (P.hp.n_tiles + 1) / 2;
        const int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);
        UKBB_CUDA(launch_pdl<2>(conv_halo_kernel<CC, COUT, RESIDENT, NKB, F16, !RESIDENT>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.hp));
        return UKBB_OK;
    }
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<CC, COUT, RESIDENT, NKB, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int grid = P.hp.n_tiles < sms ? P.hp.n_tiles : sms;
    UKBB_CUDA(launch_pdl(conv_halo_kernel<CC, COUT, RESIDENT, NKB, F16>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.hp));
    return UKBB_OK;
}
template <int CC, int COUT, bool RESIDENT, int NKB>
static int launch_halo(const TcLayerPlan& P, int sms, cudaStream_t st) {
    return P.hp.fp16 ? launch_halo2<CC, COUT, RESIDENT, NKB, true>(P, sms, st) : launch_halo2<CC, COUT, RESIDENT, NKB, false>(P, sms, st);
}

template <int CC, int COUT, int STRIDE, bool F16>
static int launch_group2(const TcLayerPlan& P, int sms, cudaStream_t st) {
    using Cfg = ConvGroupCfg<CC, COUT, STRIDE>;
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_group_kernel<CC, COUT, STRIDE, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int grid = P.gp.n_tiles < sms ? P.gp.n_tiles : sms;
    UKBB_CUDA(launch_pdl(conv_group_kernel<CC, COUT, STRIDE, F16>, grid, 256, Cfg::SMEM_BYTES, st, P.map_a, P.map_b, P.map_out, P.gp));
    return UKBB_OK;
}
template <int CC, int COUT, int STRIDE>
static int launch_group(const TcLayerPlan& P, int sms, cudaStream_t st) {
    return P.p.fp16 ? launch_group2<CC, COUT, STRIDE, true>(P, sms, st) : launch_group2<CC, COUT, STRIDE, false>(P, sms, st);
}

static int launch_plan(const TcLayerPlan& P, int sms, cudaStream_t st) {
    if (P.kind == 2) {
        if (P.p.cin == 16 && P.cout == 16 && P.p.stride == 1) return launch_group<16, 16, 1>(P, sms, st);
        if (P.p.cin == 32 && P.cout == 32 && P.p.stride == 1) return launch_group<32, 32, 1>(P, sms, st);
        if (P.p.cin == 64 && P.cout == 64 && P.p.stride == 1) return launch_group<64, 64, 1>(P, sms, st);
        if (P.p.cin == 16 && P.cout == 32 && P.p.stride == 2) return launch_group<16, 32, 2>(P, sms, st);
        if (P.p.cin == 32 && P.cout == 64 && P.p.stride == 2) return launch_group<32, 64, 2>(P, sms, st);
        set_error("conv_group: no kernel instance for %d -> %d stride %d", P.p.cin, P.cout, P.p.stride);
        return UKBB_E_UNSUPPORTED;
    }
    if (P.kind == 1) {
        if (P.cc == 16 && P.cout == 16 && P.hp.chunks == 1) return launch_halo<16, 16, true, 9>(P, sms, st);
        if (P.cc == 32 && P.cout == 32 && P.hp.chunks == 1) return launch_halo<32, 32, true, 9>(P, sms, st);
        if (P.cc == 64 && P.cout == 64 && P.hp.chunks == 1) return launch_halo<64, 64, true, 9>(P, sms, st);
        if (P.cc == 64 && P.cout == 128) return launch_halo<64, 128, false, 0>(P, sms, st);
        if (P.cc == 64 && P.cout == 256) return launch_halo<64, 256, false, 0>(P, sms, st);
        set_error("conv_halo: no kernel instance for chunk %d x %d, cout %d", P.cc, P.hp.chunks, P.cout);
        return UKBB_E_UNSUPPORTED;
    }
#define CASE(CCV, COUTV) if (P.cc == CCV && P.cout == COUTV) return launch_tc<CCV, COUTV>(P, sms, st)
    CASE(16, 16); CASE(16, 32);
    CASE(32, 32); CASE(32, 64);
    CASE(64, 32); CASE(64, 64); CASE(64, 128); CASE(64, 256);
#undef CASE
    set_error("conv_tc: no kernel instance for chunk %d, cout %d", P.cc, P.cout);
    return UKBB_E_UNSUPPORTED;
}

int bf16_prepare(Engine* h, const ukbb_fcn_weights* w) {
    Bf16State* S = new Bf16State();
    h->tc = S;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    UKBB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return UKBB_E_CUDA; }
    S->encode = (EncodeTiledFn)fn;
    S->fp16 = h->mode == UKBB_MODE_FP16 ? 1 : 0;
    S->fused_head = getenv("UKBB_NO_FUSED_HEAD") ? 0 : (getenv("UKBB_HEAD_GATHER") ? 1 : (getenv("UKBB_HEAD_V2") ? 2 : (getenv("UKBB_HEAD_V3") ? 3 : 4)));
    S->side = (S->fused_head >= 3 && !getenv("UKBB_NO_SIDE")) ? 1 : 0;
    S->first = getenv("UKBB_NO_FIRST") ? 0 : (getenv("UKBB_FIRST_FP32") ? 1 : 2);
    {   // class-score layer of head_tc: FP32 weights [k][8] and bias, passed by value (constant bank)
        const ukbb_conv_weights& c = w->conv[UKBB_N_CONV - 1];
        for (int k = 0; k < 64; ++k)
            for (int co = 0; co < 8; ++co) S->h_wl[k * 8 + co] = co < c.cout ? c.kernel[(size_t)k * c.cout + co] : 0.f;
        for (int co = 0; co < 8; ++co) S->h_bias[co] = co < c.cout ? c.bias[co] : -INFINITY;
    }
    {   // conv0_0 (CUDA cores): device tap (dy, dx) <- TF kernel[kh = dx][kw = dy], BN scale folded into the FP32 weights
        const ukbb_conv_weights& c = w->conv[0];
        for (int co = 0; co < 16; ++co) {
            const double sc = (double)c.gamma[co] / sqrt((double)c.moving_variance[co] + (double)w->bn_eps);
            S->c0.shift[co] = (float)((double)c.beta[co] - (double)c.moving_mean[co] * sc);
            for (int dy = 0; dy < 3; ++dy)
                for (int dx = 0; dx < 3; ++dx) S->c0.w[dy * 3 + dx][co] = (float)((double)c.kernel[(size_t)(dx * 3 + dy) * 16 + co] * sc);
        }
    }
    {   // conv_first_tc: B0[n = pixel s * 16 + co][k = part * 18 + row r * 6 + column c] = w_hi | w_hi | w_lo of tap (r, kx = c - s)
        std::vector<__nv_bfloat16> b0(64 * 64);
        auto r16 = [&](float v, float* back) {
            __nv_bfloat16 out;
            if (S->fp16) { const __half hv = __float2half_rn(v); memcpy(&out, &hv, 2); *back = __half2float(hv); }
            else { out = __float2bfloat16(v); *back = __bfloat162float(out); }
            return out;
        };
        for (int sp = 0; sp < 4; ++sp)
            for (int co = 0; co < 16; ++co)
                for (int k = 0; k < 64; ++k) {
                    float dummy;
                    __nv_bfloat16 val = r16(0.f, &dummy);
                    if (k < 54) {
                        const int part = k / 18, r = (k % 18) / 6, c = k % 6, kx = c - sp;
                        if (kx >= 0 && kx <= 2) {
                            const float wv = S->c0.w[r * 3 + kx][co];
                            float hi_f, lo_f;
                            const __nv_bfloat16 hi = r16(wv, &hi_f);
                            const __nv_bfloat16 lo = r16(wv - hi_f, &lo_f);
                            val = part < 2 ? hi : lo;
                        }
                    }
                    b0[(size_t)(sp * 16 + co) * 64 + k] = val;
                }
        UKBB_CUDA(cudaMalloc(&S->wb0, b0.size() * 2));
        UKBB_CUDA(cudaMemcpy(S->wb0, b0.data(), b0.size() * 2, cudaMemcpyHostToDevice));
        cuuint64_t d[2] = {64, 64}; cuuint64_t st1[1] = {128}; cuuint32_t bx[2] = {64, 64}; cuuint32_t e2[2] = {1, 1};
        CUresult r = S->encode(&S->map_b0, S->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, S->wb0, d, st1, bx, e2,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(conv0_0 weights) failed: %d", (int)r); return UKBB_E_CUDA; }
    }
    for (int li : {13, 18, 19}) {          // 1x1 layers of the head: fold gamma / sqrt(var + eps) into the weights (head_tc.cuh)
        const ukbb_conv_weights& c = w->conv[li];
        std::vector<__nv_bfloat16> wb((size_t)c.cout * c.cin);
        for (int co = 0; co < c.cout; ++co) {
            const double sc = (double)c.gamma[co] / sqrt((double)c.moving_variance[co] + (double)w->bn_eps);
            S->h_shift[li][co] = (float)((double)c.beta[co] - (double)c.moving_mean[co] * sc);
            for (int ci = 0; ci < c.cin; ++ci) {
                const float v = (float)((double)c.kernel[(size_t)ci * c.cout + co] * sc);
                __nv_bfloat16& d = wb[(size_t)co * c.cin + ci];
                if (S->fp16) { const __half hv = __float2half_rn(v); memcpy(&d, &hv, 2); }
                else d = __float2bfloat16(v);
            }
        }
        UKBB_CUDA(cudaMalloc(&S->wf[li], wb.size() * 2));
        UKBB_CUDA(cudaMemcpy(S->wf[li], wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    }
    {   // interpolation matrices of the tensor-core upsample (see head_mma.cuh) and identity scale / zero shift
        std::vector<float> oz(128, 0.f);
        for (int i = 0; i < 64; ++i) oz[i] = 1.f;
        UKBB_CUDA(cudaMalloc(&S->ones, 128 * sizeof(float)));
        UKBB_CUDA(cudaMemcpy(S->ones, oz.data(), 128 * sizeof(float), cudaMemcpyHostToDevice));
        for (int l = 1; l <= 4; ++l) {
            const int f = 1 << l, pb = (f - 1) / 2, kpad = HM_KPAD[l], pw = HM_PW[l], nv = l == 4 ? 2 : 1;
            std::vector<__nv_bfloat16> U((size_t)nv * 128 * kpad);
            std::vector<float> Uf((size_t)nv * 128 * kpad, 0.f);
            for (int v = 0; v < nv; ++v)
                for (int m = 0; m < 128; ++m) {
                    const int ty = m >> 4, tx = m & 15;
                    const int Y = (l == 4 ? 8 * v : 0) + ty + pb, X = tx + pb;
                    const int ry = Y & (f - 1), py1 = (Y >> l) + 1, rx = X & (f - 1), px1 = (X >> l) + 1;
                    const float wy1 = (float)(ry + 1) / (float)f, wy0 = 1.f - wy1, wx1 = (float)(rx + 1) / (float)f, wx0 = 1.f - wx1;
                    float* row = &Uf[((size_t)v * 128 + m) * kpad];
                    row[(py1 - 1) * pw + (px1 - 1)] += wy0 * wx0;
                    row[(py1 - 1) * pw + px1] += wy0 * wx1;
                    row[py1 * pw + (px1 - 1)] += wy1 * wx0;
                    row[py1 * pw + px1] += wy1 * wx1;
                }
            for (size_t i = 0; i < U.size(); ++i) {
                if (S->fp16) { const __half hv = __float2half_rn(Uf[i]); memcpy(&U[i], &hv, 2); }
                else U[i] = __float2bfloat16(Uf[i]);
            }
            UKBB_CUDA(cudaMalloc(&S->u[l], U.size() * 2));
            UKBB_CUDA(cudaMemcpy(S->u[l], U.data(), U.size() * 2, cudaMemcpyHostToDevice));
        }
    }
    for (int i = 1; i < UKBB_N_CONV - 1; ++i) {
        const ukbb_conv_weights& c = w->conv[i];
        const int taps = c.ksize * c.ksize, ktot = taps * c.cin;
        std::vector<__nv_bfloat16> wb((size_t)c.cout * ktot);
        for (int dy = 0; dy < c.ksize; ++dy)
            for (int dx = 0; dx < c.ksize; ++dx)
                for (int ci = 0; ci < c.cin; ++ci)
                    for (int co = 0; co < c.cout; ++co)      // device tap (dy,dx) <- TF kernel[kh=dx][kw=dy]
                    {
                        const float v = c.kernel[((size_t)(dx * c.ksize + dy) * c.cin + ci) * c.cout + co];
                        __nv_bfloat16& dstw = wb[(size_t)co * ktot + (dy * c.ksize + dx) * c.cin + ci];
                        if (S->fp16) { const __half hv = __float2half_rn(v); memcpy(&dstw, &hv, 2); }
                        else dstw = __float2bfloat16(v);
                    }
        UKBB_CUDA(cudaMalloc(&S->w[i], wb.size() * sizeof(__nv_bfloat16)));
        UKBB_CUDA(cudaMemcpy(S->w[i], wb.data(), wb.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
        // pixel-group layers (conv_group.cuh): N = gout * cout = 64.  Tile (ky, j) row (s, co) holds tap (ky, kx) with
        // kx = j - s (stride 1) or j - 2 s (stride 2), zero where that tap does not exist.
        if (c.ksize == 3 && c.cin <= 64 && 64 % c.cin == 0) {
            const int g = 64 / c.cin, gout = c.stride == 1 ? g : g / 2;
            if (gout >= 1 && gout * c.cout == 64) {
                const int jn = c.stride == 1 ? g + 2 : g + 1;
                std::vector<__nv_bfloat16> we((size_t)3 * jn * 64 * c.cin);
                for (int ky = 0; ky < 3; ++ky)
                    for (int j = 0; j < jn; ++j)
                        for (int sp = 0; sp < gout; ++sp)
                            for (int co = 0; co < c.cout; ++co)
                                for (int ci = 0; ci < c.cin; ++ci) {
                                    const int kx = c.stride == 1 ? j - sp : j - 2 * sp;
                                    __nv_bfloat16 v16;
                                    const uint16_t zero = 0;
                                    memcpy(&v16, &zero, 2);
                                    if (kx >= 0 && kx <= 2) v16 = wb[(size_t)co * ktot + (ky * 3 + kx) * c.cin + ci];
                                    we[((size_t)(ky * jn + j) * 64 + sp * c.cout + co) * c.cin + ci] = v16;
                                }
                UKBB_CUDA(cudaMalloc(&S->wg[i], we.size() * sizeof(__nv_bfloat16)));
                UKBB_CUDA(cudaMemcpy(S->wg[i], we.data(), we.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
            }
        }
    }
    return UKBB_OK;
}

void bf16_release(Engine* h) {
    Bf16State* S = h->tc;
    if (!S) return;
    for (int i = 0; i < UKBB_N_CONV; ++i) { cudaFree(S->w[i]); cudaFree(S->wg[i]); }
    cudaFree(S->cat); cudaFree(S->f0); cudaFree(S->f1);
    for (int l = 0; l < 5; ++l) { cudaFree(S->t[l]); cudaFree(S->u[l]); }
    cudaFree(S->ones); cudaFree(S->wb0);
    for (int i = 0; i < UKBB_N_CONV; ++i) cudaFree(S->wf[i]);
    delete S;
    h->tc = nullptr;
}

static const int kNBlockTc[5] = {2, 2, 3, 3, 3};
static const int kNFilterTc[5] = {16, 32, 64, 128, 256};

static int ensure_plans(Engine* h, int nb, int h2, int w2) {
    Bf16State* S = h->tc;
    if (S->plan_nb == nb && S->plan_h == h2 && S->plan_w == w2) return UKBB_OK;
    UKBB_CUDA(cudaDeviceSynchronize());
    // activation workspace (bf16): encoder ping/pong + same_dim per level, concat + fc buffers
    for (int l = 0; l < 5; ++l) {
        cudaFree(h->ws.a[l]); cudaFree(h->ws.b[l]); cudaFree(h->ws.s[l]);
        h->ws.a[l] = h->ws.b[l] = h->ws.s[l] = nullptr;
        const size_t px = (size_t)nb * (h2 >> l) * (w2 >> l);
        UKBB_CUDA(cudaMalloc(&h->ws.a[l], px * kNFilterTc[l] * 2));
        UKBB_CUDA(cudaMalloc(&h->ws.b[l], px * kNFilterTc[l] * 2));
        UKBB_CUDA(cudaMalloc(&h->ws.s[l], px * 32 * 2));
    }
    cudaFree(S->cat); cudaFree(S->f0); cudaFree(S->f1);
    S->cat = S->f0 = S->f1 = nullptr;
    const size_t px = (size_t)nb * h2 * w2;
    if (S->fused_head == 0) {
        UKBB_CUDA(cudaMalloc(&S->cat, px * 160 * 2));
        UKBB_CUDA(cudaMalloc(&S->f0, px * 64 * 2));
        UKBB_CUDA(cudaMalloc(&S->f1, px * 64 * 2));
    }
    for (int l = 1; l <= 4; ++l) {
        cudaFree(S->t[l]); S->t[l] = nullptr;
        if (S->fused_head >= 2) UKBB_CUDA(cudaMalloc(&S->t[l], (size_t)nb * (h2 >> l) * (w2 >> l) * 64 * 2));
    }
    h->ws.nb = nb; h->ws.h = h2; h->ws.w = w2;
    int li = 0, rc;
    const __nv_bfloat16* cur = nullptr;
    const __nv_bfloat16* level_out[5];
    int hi = h2, wi = w2;
    for (int l = 0; l < 5; ++l) {
        for (int b = 0; b < kNBlockTc[l]; ++b, ++li) {
            __nv_bfloat16* dst = (__nv_bfloat16*)((b & 1) ? h->ws.b[l] : h->ws.a[l]);
            if (li > 0) {
                rc = make_plan(h, li, cur, dst, nb, hi, wi, l);
                if (rc) return rc;
                hi = S->plan[li].p.ho; wi = S->plan[li].p.wo;
            }
            cur = dst;
        }
        level_out[l] = cur;
    }
    for (int l = 0; l < 5; ++l, ++li) {
        rc = make_plan(h, li, level_out[l], (__nv_bfloat16*)h->ws.s[l], nb, h2 >> l, w2 >> l, l);
        if (rc) return rc;
    }
    if (S->fused_head == 0) {
        rc = make_plan(h, 18, S->cat, S->f0, nb, h2, w2, 0);
        if (rc) return rc;
        rc = make_plan(h, 19, S->f0, S->f1, nb, h2, w2, 0);
        if (rc) return rc;
    }
    if (S->fused_head >= 2) {
        const CUtensorMapDataType dt16 = S->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        const CUtensorMapSwizzle sw128 = CU_TENSOR_MAP_SWIZZLE_128B;
        cuuint32_t e4[4] = {1, 1, 1, 1}, e2[2] = {1, 1};
        CUtensorMap* tm[5] = {nullptr, &S->hm.t1, &S->hm.t2, &S->hm.t3, &S->hm.t4};
        CUtensorMap* um[5] = {nullptr, &S->hm.u1, &S->hm.u2, &S->hm.u3, &S->hm.u4};
        for (int l = 1; l <= 4; ++l) {
            // t_l = W_l . s_l : 1x1 conv 32 -> 64 on the same_dim output of level l, columns [32 l, 32 l + 32) of W_fc0
            TcLayerPlan saved = S->plan[18];
            const ConvLayer keep = h->layers[18];
            ConvLayer& L = h->layers[18];
            L.cin = 32; L.relu = 0; L.scale = S->ones; L.shift = S->ones + 64;
            rc = make_plan(h, 18, (const __nv_bfloat16*)h->ws.s[l], S->t[l], nb, h2 >> l, w2 >> l, l);
            h->layers[18] = keep;
            if (rc) { S->plan[18] = saved; return rc; }
            S->tplan[l] = S->plan[18];
            S->plan[18] = saved;
            S->tplan[l].p.kofs = 32 * l;
            {   // weight map must span all 160 K columns: rebuild it (make_plan used cin = 32)
                cuuint64_t d0[2] = {160, 64}; cuuint64_t st0[1] = {320}; cuuint32_t b0[2] = {32, 64};
                CUresult r = S->encode(&S->tplan[l].map_b, dt16, 2, S->fused_head >= 3 ? S->wf[18] : S->w[18], d0, st0, b0, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(t-layer weights) failed: %d", (int)r); return UKBB_E_CUDA; }
            }
            {   // t_l patches for the head: [nb][h_l][w_l][64], box (64, PW, PH, 1)
                cuuint64_t dims[4] = {64, (cuuint64_t)(w2 >> l), (cuuint64_t)(h2 >> l), (cuuint64_t)nb};
                cuuint64_t strides[3] = {128, (cuuint64_t)(w2 >> l) * 128, (cuuint64_t)(h2 >> l) * (w2 >> l) * 128};
                cuuint32_t box[4] = {64, (cuuint32_t)HM_PW[l], (cuuint32_t)HM_PH[l], 1};
                CUresult r = S->encode(tm[l], dt16, 4, S->t[l], dims, strides, box, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, sw128,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(t patch %d) failed: %d", l, (int)r); return UKBB_E_CUDA; }
            }
            {   // U_l: [nv * 128][kpad]
                const int kpad = HM_KPAD[l], nv = l == 4 ? 2 : 1;
                cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)(nv * 128)};
                cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
                cuuint32_t box[2] = {(cuuint32_t)kpad, 128};
                CUresult r = S->encode(um[l], dt16, 2, S->u[l], dims, strides, box, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       swizzle_for(kpad), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(U %d) failed: %d", l, (int)r); return UKBB_E_CUDA; }
            }
        }
    }
    if (S->side) {
        const CUtensorMapDataType dt16 = S->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        cuuint32_t e2[2] = {1, 1};
        CUresult r = CUDA_SUCCESS;
        for (int l = 1; l <= 4 && r == CUDA_SUCCESS; ++l) {
            const int cin = kNFilterTc[l], kc = cin < 64 ? cin : 64;
            const cuuint64_t rows = (cuuint64_t)nb * (h2 >> l) * (w2 >> l);
            const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
            cuuint64_t di[2] = {(cuuint64_t)cin, rows}; cuuint64_t si[1] = {(cuuint64_t)cin * 2}; cuuint32_t bi[2] = {(cuuint32_t)kc, 128};
            r = S->encode(&S->sm.in[l - 1], dt16, 2, (void*)level_out[l], di, si, bi, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t dout[2] = {64, rows}; cuuint64_t so[1] = {128}; cuuint32_t bo[2] = {64, 128};
            if (r == CUDA_SUCCESS)
                r = S->encode(&S->sm.out[l - 1], dt16, 2, S->t[l], dout, so, bo, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t dw[2] = {(cuuint64_t)cin, 32}; cuuint64_t sw1[1] = {(cuuint64_t)cin * 2}; cuuint32_t bw2[2] = {(cuuint32_t)kc, 32};
            if (r == CUDA_SUCCESS)
                r = S->encode(&S->sm.wsd[l - 1], dt16, 2, S->w[13 + l], dw, sw1, bw2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        cuuint64_t d0[2] = {160, 64}; cuuint64_t st0[1] = {320}; cuuint32_t b0[2] = {32, 64};
        if (r == CUDA_SUCCESS)
            r = S->encode(&S->sm.w0, dt16, 2, S->wf[18], d0, st0, b0, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(side branches) failed: %d", (int)r); return UKBB_E_CUDA; }
    }
    {   // fused head operands: s0 tiles (8 rows x 16 columns x 32 channels), fc0 / fc1 weights
        const CUtensorMapDataType dt16 = S->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        cuuint64_t dims[4] = {32, (cuuint64_t)w2, (cuuint64_t)h2, (cuuint64_t)nb};
        cuuint64_t strides[3] = {64, (cuuint64_t)w2 * 64, (cuuint64_t)h2 * w2 * 64};
        cuuint32_t box[4] = {32, 16, 8, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = S->encode(&S->map_s0, dt16, 4, h->ws.s[0], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t d0[2] = {160, 64}; cuuint64_t st0[1] = {320}; cuuint32_t b0[2] = {32, 64}; cuuint32_t e2[2] = {1, 1};
        if (r == CUDA_SUCCESS)
            r = S->encode(&S->map_w0, dt16, 2, S->fused_head >= 3 ? S->wf[18] : S->w[18], d0, st0, b0, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t d1[2] = {64, 64}; cuuint64_t st1[1] = {128}; cuuint32_t b1[2] = {64, 64};
        if (r == CUDA_SUCCESS)
            r = S->encode(&S->map_w1, dt16, 2, S->fused_head >= 3 ? S->wf[19] : S->w[19], d1, st1, b1, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(fused head) failed: %d", (int)r); return UKBB_E_CUDA; }
        S->hm.s0 = S->map_s0; S->hm.w0 = S->map_w0; S->hm.w1 = S->map_w1;
        if (S->fused_head >= 2) {
            // head_mma computes same_dim0 itself: its level-0 input is the conv0_1 output (16 channels)
            cuuint64_t dimsb[4] = {16, (cuuint64_t)w2, (cuuint64_t)h2, (cuuint64_t)nb};
            cuuint64_t stridesb[3] = {32, (cuuint64_t)w2 * 32, (cuuint64_t)h2 * w2 * 32};
            cuuint32_t boxb[4] = {16, 16, 8, 1};
            r = S->encode(&S->hm.s0, dt16, 4, h->ws.b[0], dimsb, stridesb, boxb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t dsd[2] = {16, 32}; cuuint64_t ssd[1] = {32}; cuuint32_t bsd[2] = {16, 32};
            if (r == CUDA_SUCCESS)
                r = S->encode(&S->hm.wsd, dt16, 2, S->fused_head >= 3 ? S->wf[13] : S->w[13], dsd, ssd, bsd, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(head same_dim0) failed: %d", (int)r); return UKBB_E_CUDA; }
        }
    }
    S->plan_nb = nb; S->plan_h = h2; S->plan_w = w2;
    return UKBB_OK;
}

template <int NC, bool F16>
static int launch_head2(const Bf16State* S, const HeadParams& hp, int sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(head_fused_kernel<NC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM));
        attr_set = true;
    }
    const int grid = hp.n_tiles < sms ? hp.n_tiles : sms;
    head_fused_kernel<NC, F16><<<grid, HEAD_THREADS, HEAD_SMEM, st>>>(S->map_s0, S->map_w0, S->map_w1, hp);
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}
template <int NC, bool F16>
static int launch_head_mma2(const Bf16State* S, const HeadParams& hp, int sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(head_mma_kernel<NC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HM_SMEM));
        attr_set = true;
    }
    const int grid = hp.n_tiles < sms ? hp.n_tiles : sms;
    head_mma_kernel<NC, F16><<<grid, HM_THREADS, HM_SMEM, st>>>(S->hm, hp);
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}
template <int NC, bool F16>
static int launch_head_tc2(const Bf16State* S, const HeadParams& hp, int sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(head_tc_kernel<NC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM));
        attr_set = true;
    }
    const int grid = hp.n_tiles < sms ? hp.n_tiles : sms;
    UKBB_CUDA(launch_pdl(head_tc_kernel<NC, F16>, grid, H3_THREADS, H3_SMEM, st, S->hm, hp));
    return UKBB_OK;
}
template <int NC, bool F16>
static int launch_head_ts2(const Bf16State* S, const HeadParams& hp, int sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(head_ts_kernel<NC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, H4_SMEM));
        attr_set = true;
    }
    const int grid = hp.n_tiles < sms ? hp.n_tiles : sms;
    UKBB_CUDA(launch_pdl(head_ts_kernel<NC, F16>, grid, H4_THREADS, H4_SMEM, st, S->hm, hp));
    return UKBB_OK;
}
template <bool F16>
static int launch_side(const Bf16State* S, const SideParams& sp, int sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(side_tc_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SD_SMEM));
        attr_set = true;
    }
    const int n_tiles = sp.tile_start[4];
    const int grid = n_tiles < sms ? n_tiles : sms;
    UKBB_CUDA(launch_pdl(side_tc_kernel<F16>, grid, SD_THREADS, SD_SMEM, st, S->sm, sp));
    return UKBB_OK;
}
template <int NC>
static int launch_head(const Bf16State* S, const HeadParams& hp, int sms, cudaStream_t st) {
    if (S->fused_head == 4) return S->fp16 ? launch_head_ts2<NC, true>(S, hp, sms, st) : launch_head_ts2<NC, false>(S, hp, sms, st);
    if (S->fused_head == 3) return S->fp16 ? launch_head_tc2<NC, true>(S, hp, sms, st) : launch_head_tc2<NC, false>(S, hp, sms, st);
    if (S->fused_head == 2) return S->fp16 ? launch_head_mma2<NC, true>(S, hp, sms, st) : launch_head_mma2<NC, false>(S, hp, sms, st);
    return S->fp16 ? launch_head2<NC, true>(S, hp, sms, st) : launch_head2<NC, false>(S, hp, sms, st);
}

template <bool F16>
static int launch_first2(const Bf16State* S, const TcLayerPlan& P1, const CUtensorMap& map_img, int nb, int h2, int w2, int sms, cudaStream_t st) {
    using Cfg = ConvFirstCfg;
    static bool attr_set = false;
    if (!attr_set) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_first_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    ConvFirstParams fp;
    fp.tiles_x = P1.gp.tiles_x; fp.tiles_y = P1.gp.tiles_y; fp.n_tiles = fp.tiles_x * fp.tiles_y * nb;
    fp.h = h2; fp.w4 = w2 / 4;
    fp.scale = P1.gp.scale; fp.shift = P1.gp.shift;
    memcpy(fp.w0, S->c0.w, sizeof(fp.w0));
    memcpy(fp.shift0, S->c0.shift, sizeof(fp.shift0));
    const int grid = fp.n_tiles < sms ? fp.n_tiles : sms;
    if (S->first == 2) {
        using CfgT = ConvFirstTcCfg;
        static bool attr_set_tc = false;
        if (!attr_set_tc) {
            UKBB_CUDA(cudaFuncSetAttribute(conv_first_tc_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgT::SMEM_BYTES));
            attr_set_tc = true;
        }
        UKBB_CUDA(launch_pdl(conv_first_tc_kernel<F16>, grid, CfgT::THREADS, CfgT::SMEM_BYTES, st, map_img, P1.map_b, S->map_b0, P1.map_out, fp));
        return UKBB_OK;
    }
    UKBB_CUDA(launch_pdl(conv_first_kernel<F16>, grid, Cfg::THREADS, Cfg::SMEM_BYTES, st, map_img, P1.map_b, P1.map_out, fp));
    return UKBB_OK;
}

// Test hook: run ONE tensor-core conv layer of the engine on a caller-provided BF16 NHWC tensor.
int debug_conv_bf16(Engine* h, int li, const void* in, int n, int hi, int wi, int level_out, void* out, cudaStream_t st) {
    UKBB_REQUIRE(h->tc, "debug_conv: engine is not in BF16 mode");
    UKBB_REQUIRE(li >= 1 && li < UKBB_N_CONV - 1, "debug_conv: layer %d has no tensor-core kernel", li);
    Bf16State* S = h->tc;
    TcLayerPlan saved = S->plan[li];
    int rc = make_plan(h, li, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, hi, wi, level_out);
    if (!rc) rc = launch_plan(S->plan[li], h->sms, st);
    if (!rc) { cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) { set_error("debug_conv: %s", cudaGetErrorString(e)); rc = UKBB_E_CUDA; } }
    S->plan[li] = saved;
    h->launches++;
    return rc;
}

int forward_bf16(Engine* h, const float* image, int n, int x2, int y2, int x_pre, int y_pre, int x, int y,
                 uint8_t* labels, float* logits, float* prob, unsigned long long* counts, cudaStream_t st) {
    Bf16State* S = h->tc;
    const int w2 = x2, h2 = y2;
    UKBB_REQUIRE((w2 >> 4) >= 1 && (h2 >> 4) >= 1, "forward_bf16: image too small");
    // sub-batch: a multiple of 128 slices would be ideal for the deepest level (bn = 128); keep the
    // activation working set bounded instead
    int cap = 500;
    if (const char* e = getenv("UKBB_SUBBATCH")) { const int v = atoi(e); if (v > 0) cap = v; }
    // equal sub-batches (1000 slices -> 2 x 500 rather than 500 + 500 + 0); one SA subject (500 slices) is ONE sub-batch:
    // measured 151k -> 165k -> 172k slices/s for caps 125 / 250 / 500 (fewer launches, fuller last waves)
    const int parts = (n + cap - 1) / cap;
    int NB = (n + parts - 1) / parts;
    // plans (tensor maps, workspace) built for a larger sub-batch of the same image size serve any smaller one
    int rc = (S->plan_nb >= NB && S->plan_h == h2 && S->plan_w == w2) ? UKBB_OK : ensure_plans(h, NB, h2, w2);
    if (rc) return rc;
    for (int n0 = 0; n0 < n; n0 += NB) {
        const int nb = n - n0 < NB ? n - n0 : NB;
        const bool first = S->first && S->plan[1].kind == 2;
        if (first) {
            // conv0_0 + conv0_1 in one launch; the image box map names this call's image pointer
            CUtensorMap map_img;
            cuuint64_t dims[3] = {(cuuint64_t)w2, (cuuint64_t)h2, (cuuint64_t)nb};
            cuuint64_t strides[2] = {(cuuint64_t)w2 * 4, (cuuint64_t)h2 * w2 * 4};
            cuuint32_t box[3] = {ConvFirstCfg::IMG_W, ConvFirstCfg::IMG_H, 1}, e3[3] = {1, 1, 1};
            CUresult r = S->encode(&map_img, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)(image + (size_t)n0 * h2 * w2), dims, strides, box, e3,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(image boxes) failed: %d", (int)r); return UKBB_E_CUDA; }
            rc = S->fp16 ? launch_first2<true>(S, S->plan[1], map_img, nb, h2, w2, h->sms, st)
                         : launch_first2<false>(S, S->plan[1], map_img, nb, h2, w2, h->sms, st);
            if (rc) return rc;
            h->launches++;
        } else {
            const long long total = (long long)nb * h2 * w2;
            const unsigned grid0 = (unsigned)((total + 255) / 256);
            if (S->fp16) conv0_bf16_kernel<true><<<grid0, 256, 0, st>>>(image + (size_t)n0 * h2 * w2, (__nv_bfloat16*)h->ws.a[0], total, h2, w2, S->c0);
            else conv0_bf16_kernel<false><<<grid0, 256, 0, st>>>(image + (size_t)n0 * h2 * w2, (__nv_bfloat16*)h->ws.a[0], total, h2, w2, S->c0);
            UKBB_CUDA(cudaGetLastError());
            h->launches++;
        }
        for (int li = first ? 2 : 1; li < 18; ++li) {
            if (li == 13 && S->fused_head >= 2) continue;        // same_dim0 lives inside head_mma_kernel
            if (li > 13 && S->side) break;                       // same_dim 1..4 live inside side_tc_kernel
            TcLayerPlan P = S->plan[li];
            P.p.n = nb;
            P.p.n_tiles = P.p.tiles_x * P.p.tiles_y * ((nb + P.p.bn - 1) / P.p.bn);
            P.hp.n = nb;
            P.hp.n_tiles = P.hp.tiles_x * P.hp.tiles_y * nb;
            P.gp.n_tiles = P.gp.tiles_x * P.gp.tiles_y * nb;
            rc = launch_plan(P, h->sms, st);
            if (rc) return rc;
            h->launches++;
        }
        if (S->side) {
            SideParams sp;
            int acc_tiles = 0;
            for (int k = 0; k < 4; ++k) {
                const int l = 4 - k;
                sp.tile_start[k] = acc_tiles;
                acc_tiles += (int)(((long long)nb * (h2 >> l) * (w2 >> l) + 127) / 128);
                sp.scale[l - 1] = h->layers[13 + l].scale; sp.shift[l - 1] = h->layers[13 + l].shift;
            }
            sp.tile_start[4] = acc_tiles;
            rc = S->fp16 ? launch_side<true>(S, sp, h->sms, st) : launch_side<false>(S, sp, h->sms, st);
            if (rc) return rc;
            h->launches++;
        } else if (S->fused_head >= 2) {
            for (int l = 1; l <= 4; ++l) {
                TcLayerPlan P = S->tplan[l];
                P.p.n = nb;
                P.p.n_tiles = P.p.tiles_x * P.p.tiles_y * ((nb + P.p.bn - 1) / P.p.bn);
                rc = launch_plan(P, h->sms, st);
                if (rc) return rc;
                h->launches++;
            }
        }
        if (S->fused_head) {
            HeadParams hp;
            for (int l = 0; l < 5; ++l) hp.s[l] = (const __nv_bfloat16*)h->ws.s[l];
            hp.n = nb; hp.h = h2; hp.w = w2;
            hp.tiles_x = w2 / 16; hp.tiles_y = h2 / 8; hp.n_tiles = hp.tiles_x * hp.tiles_y * nb;
            hp.x_pre = x_pre; hp.y_pre = y_pre; hp.x = x; hp.y = y;
            hp.fp16 = S->fp16; hp.nc = h->n_class;
            hp.scale0 = h->layers[18].scale; hp.shift0 = h->layers[18].shift;
            hp.scale1 = h->layers[19].scale; hp.shift1 = h->layers[19].shift;
            hp.scale_sd0 = h->layers[13].scale; hp.shift_sd0 = h->layers[13].shift;
            hp.wlog = h->layers[20].w_f32; hp.blog = h->layers[20].shift;
            memcpy(hp.c_shift_sd0, S->h_shift[13], sizeof(hp.c_shift_sd0));
            memcpy(hp.c_shift0, S->h_shift[18], sizeof(hp.c_shift0));
            memcpy(hp.c_shift1, S->h_shift[19], sizeof(hp.c_shift1));
            memcpy(hp.c_bias, S->h_bias, sizeof(hp.c_bias));
            for (int l = 0; l < 5; ++l) hp.u_glob[l] = (const uint32_t*)S->u[l];
            hp.b0 = (const uint4*)h->ws.b[0];
            hp.dbg = getenv("UKBB_HEAD_DBG") ? atoi(getenv("UKBB_HEAD_DBG")) : 0;
            hp.trace = nullptr;
            static long long* d_trace = nullptr;
            if (hp.dbg & 16) {
                if (!d_trace) { UKBB_CUDA(cudaMalloc(&d_trace, 12 * 64 * sizeof(long long))); }
                UKBB_CUDA(cudaMemsetAsync(d_trace, 0, 12 * 64 * sizeof(long long), st));
                hp.trace = d_trace;
            }
            for (int k = 0; k < 64; ++k) {
                hp.c_nshift1[k] = -S->h_shift[19][k];
                for (int c = 0; c < 8; ++c) hp.c_wlc[k][c] = S->h_wl[k * 8 + c];
            }
            for (int c = 0; c < 8; ++c) {
                double acc = 0.0;
                for (int k = 0; k < 64; ++k) acc += (double)S->h_shift[19][k] * (double)S->h_wl[k * 8 + c];
                hp.c_bias2[c] = c < h->n_class ? (float)((double)S->h_bias[c] + acc) : -INFINITY;
            }
            for (int k2 = 0; k2 < 32; ++k2)
                for (int c = 0; c < 8; ++c) hp.c_wl2[k2][c] = make_float2(S->h_wl[(2 * k2) * 8 + c], S->h_wl[(2 * k2 + 1) * 8 + c]);
            const size_t po = (size_t)n0 * h2 * w2 * h->n_class;
            hp.labels = labels + (size_t)n0 * x * y;
            hp.logits = logits ? logits + po : nullptr;
            hp.prob = prob ? prob + po : nullptr;
            hp.counts = counts ? counts + (size_t)n0 * h->n_class : nullptr;
            cudaEvent_t kt0 = nullptr, kt1 = nullptr;
            if (h->ktimer) {                                        // ukbb_fcn_kernel_timer: events on the launching stream
                UKBB_CUDA(cudaEventCreate(&kt0)); UKBB_CUDA(cudaEventCreate(&kt1));
                UKBB_CUDA(cudaEventRecord(kt0, st));
            }
            switch (h->n_class) {
                case 2: rc = launch_head<2>(S, hp, h->sms, st); break;
                case 3: rc = launch_head<3>(S, hp, h->sms, st); break;
                case 4: rc = launch_head<4>(S, hp, h->sms, st); break;
                case 5: case 6: rc = launch_head<6>(S, hp, h->sms, st); break;
                default: rc = launch_head<8>(S, hp, h->sms, st); break;
            }
            if (h->ktimer && !rc) { UKBB_CUDA(cudaEventRecord(kt1, st)); h->ktimer_ev.emplace_back(kt0, kt1); }
            if (rc) return rc;
            h->launches++;
            if ((hp.dbg & 16) && getenv("UKBB_HEAD_TRACE")) {       // experiment: dump the timeline of CTA 0
                long long ht[12 * 64];
                UKBB_CUDA(cudaStreamSynchronize(st));
                UKBB_CUDA(cudaMemcpy(ht, hp.trace, sizeof(ht), cudaMemcpyDeviceToHost));
                if (FILE* f = fopen(getenv("UKBB_HEAD_TRACE"), "wb")) { fwrite(ht, 1, sizeof(ht), f); fclose(f); }
            }
            continue;
        }
        {
            const long long total = (long long)nb * h2 * w2 * 20;
            upsample_concat_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                (const __nv_bfloat16*)h->ws.s[0], (const __nv_bfloat16*)h->ws.s[1], (const __nv_bfloat16*)h->ws.s[2],
                (const __nv_bfloat16*)h->ws.s[3], (const __nv_bfloat16*)h->ws.s[4], S->cat, total, h2, w2, S->fp16);
            UKBB_CUDA(cudaGetLastError());
            h->launches++;
        }
        for (int li = 18; li < 20; ++li) {
            TcLayerPlan P = S->plan[li];
            P.p.n = nb;
            P.p.n_tiles = P.p.tiles_x * P.p.tiles_y * ((nb + P.p.bn - 1) / P.p.bn);
            rc = launch_plan(P, h->sms, st);
            if (rc) return rc;
            h->launches++;
        }
        {
            const ConvLayer& L = h->layers[20];
            dim3 block(256), grid((h2 * w2 + 255) / 256, nb);
            const size_t po = (size_t)n0 * h2 * w2 * h->n_class;
            uint8_t* lab = labels + (size_t)n0 * x * y;
            float* lg = logits ? logits + po : nullptr;
            float* pr = prob ? prob + po : nullptr;
            unsigned long long* cn = counts ? counts + (size_t)n0 * h->n_class : nullptr;
#define LAUNCH(NC) classifier_bf16_kernel<NC><<<grid, block, 0, st>>>(S->f1, L.w_f32, L.shift, h2, w2, x_pre, y_pre, x, y, lab, lg, pr, cn, S->fp16)
            switch (h->n_class) {
                case 2: LAUNCH(2); break;
                case 3: LAUNCH(3); break;
                case 4: LAUNCH(4); break;
                case 5: LAUNCH(5); break;
                case 6: LAUNCH(6); break;
                case 7: LAUNCH(7); break;
                case 8: LAUNCH(8); break;
                default: set_error("classifier: n_class=%d unsupported", h->n_class); return UKBB_E_UNSUPPORTED;
            }
#undef LAUNCH
            UKBB_CUDA(cudaGetLastError());
            h->launches++;
        }
    }
    return UKBB_OK;
}

}  // namespace ukbb
