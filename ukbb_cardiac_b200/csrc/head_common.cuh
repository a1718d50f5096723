// Shared declarations of the fused head (head_ts.cuh) and its feeder (side_tc.cuh): the algebra, the layout of the
// per-tile inputs in shared memory, the kernel parameters and the packed epilogue helper.
//
// network.py:207-229 computes  fc0( concat_l( up_l(s_l) ) ).  Both the fixed bilinear transposed
// convolution up_l (network.py:138-167) and the 1x1 convolution fc0 are linear, and they act on
// different axes (pixels vs channels), so they commute:
//     W_fc0 . concat_l(up_l(s_l)) = W_0 . s_0 + sum_{l=1..4} up_l( W_l . s_l ),      W_fc0 = [W_0 | ... | W_4]
// t_l = W_l . s_l is a 32->64 1x1 convolution at the LOW resolution of level l (side_tc_kernel, no
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
#pragma once
#include "tc_common.cuh"

namespace ukbb {

constexpr int HM_PW[5] = {0, 9, 6, 4, 3};          // source patch width  (x) per level
constexpr int HM_PH[5] = {0, 5, 4, 3, 3};          // source patch height (y) per level
constexpr int HM_KPAD[5] = {32, 64, 32, 16, 16};   // padded K of the U_l matrices (level 0: s0 channels)
// shared-memory layout of one input stage (bytes): t_l patches as TMA writes them, rows of 64 channels = 128 B
constexpr int HM_IN_P1 = 48 * 128, HM_IN_P2 = 32 * 128, HM_IN_P3 = 16 * 128, HM_IN_P4 = 16 * 128;
constexpr int HM_IN_PATCHES = HM_IN_P1 + HM_IN_P2 + HM_IN_P3 + HM_IN_P4;     // one plane of the four patches
constexpr int HM_IN_PATCH_TX = (45 + 24 + 12 + 9) * 128;                     // bytes TMA actually delivers per plane and tile
constexpr int HM_W0 = 64 * 64, HM_W1 = 64 * 128, HM_WSD = 1024;              // fc0 level-0 slice [64][32], fc1 [64][64], same_dim0 [32][16]

struct HeadMaps {
    CUtensorMap t1, t2, t3, t4, w0, w1, wsd;       // t_l: [n][h_l][w_l][64], box (64, PW, PH, 1)
};

struct HeadParams {
    int n, h, w;                    // slices, padded rows (Y2), padded columns (X2)
    int tiles_x, tiles_y, n_tiles;
    int x_pre, y_pre, x, y;         // crop
    int nc;
    int lo_n;                       // split modes: slice index of the lo plane in the t_l tensor maps
    long long b0_lo;                // split modes: offset of the lo plane of b0 in uint4 units
    uint8_t* labels;                // [n][y][x]
    float* logits;                  // optional [n][h][w][nc]
    float* prob;                    // optional
    unsigned long long* counts;     // optional [n][nc]
    // BN scales are folded into the 16-bit weights, the shifts travel by value so that the
    // epilogues read them as constant-bank operands (no shared-memory loads, no registers)
    float c_shift_sd0[32], c_shift0[64];
    const uint32_t* u_glob[5];      // interpolation matrices U_l in global memory, rows of HM_KPAD[l] 16-bit values
    const uint4* b0;                // conv0_1 output [n][h][w][16] 16-bit (level-0 input of same_dim0)
    float c_nshift1[64];            // -shift of fc1;  relu(d + s) . w = max(d, -s) . w + s . w
    float c_bias2[8];               //   bias + sum_k shift1[k] * w[k][c]  (-inf for c >= n_class)
    float c_wlc[64][8];             //   class-score weights [k][class], zero for c >= n_class
};

namespace tc {
// (a0 + s0, a1 + s1) -> ReLU -> two 16-bit values; s0 / s1 are meant to be constant-bank operands
template <bool F16>
__device__ __forceinline__ uint32_t add_relu_pack(uint32_t a0, uint32_t a1, float s0, float s1) {
    uint64_t a, sh, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sh) : "f"(s0), "f"(s1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(sh));
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d));
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// same, split into hi and lo 16-bit pieces (x3 modes)
template <bool F16>
__device__ __forceinline__ void add_relu_split(uint32_t a0, uint32_t a1, float s0, float s1, uint32_t& hi, uint32_t& lo) {
    uint64_t a, sh, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sh) : "f"(s0), "f"(s1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(sh));
    float x0, x1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d));
    split_pack_relu<F16>(x0, x1, hi, lo);
}
// four channels (quad q of a 16-channel group): + shift, ReLU, split_pack4 (tc_common.cuh)
template <bool F16, bool F8>
__device__ __forceinline__ void add_relu_split4(const uint32_t* a, const float* s, uint32_t* oh, uint32_t* ol, int q) {
    uint64_t x, sh, d0, d1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a[0]), "r"(a[1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sh) : "f"(s[0]), "f"(s[1]));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d0) : "l"(x), "l"(sh));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a[2]), "r"(a[3]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sh) : "f"(s[2]), "f"(s[3]));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d1) : "l"(x), "l"(sh));
    float x0, x1, x2, x3;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(d1));
    split_pack4<F16, F8, true>(x0, x1, x2, x3, oh, ol, q);
}
}  // namespace tc

}  // namespace ukbb
