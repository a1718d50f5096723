// Host-side state of the tensor-core path (BF16 / FP16 and their split-operand "x3" variants): per-layer launch plans
// (tensor maps + kernel parameters), device copies of the re-laid-out weights, and the launchers of each kernel family.
// The kernels are compiled in separate translation units (tc_conv.cu, tc_first.cu, tc_head.cu, tc_head_x3.cu); tc_forward.cu
// builds the plans and strings the launches together.
#pragma once
#include "engine.cuh"
#include "tc_common.cuh"
#include "conv_group.cuh"
#include "conv_halo.cuh"
#include "conv_first_tc.cuh"
#include "side_tc.cuh"
#include "head_common.cuh"

namespace ukbb {

struct ConvTcParams {
    int taps, ks, stride, cin;
    int kofs;                       // first K column of the weight matrix (sub-matrix selection)
    int pad_top, pad_left;
    int bw, bh, bn;                 // output box of one tile: bw * bh * bn == 128
    int tiles_x, tiles_y, n_tiles;
    int ho, wo, n;                  // output height / width / slices actually valid
    int relu;
    int fp16;                       // operand format: 0 = BF16, 1 = FP16
    int lo_n;                       // split mode: slice index of the lo plane in the input tensor map (= plan capacity)
    long long out_lo;               // split mode: element offset of the lo plane of `out`
    const float* scale;
    const float* shift;
    __nv_bfloat16* out;             // [n][ho][wo][COUT]  (split mode: hi plane, lo plane at out + out_lo)
};

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
    int kind = 0;                   // 0 = per-tap implicit GEMM (conv_tc_kernel), 1 = halo reuse, streamed weights (conv_halo_kernel),
                                    // 2 = pixel-group rows (conv_group_kernel)
    int split = 0;
    int f8 = 0;                     // split scheme with FP8 correction operands in the lo planes (UKBB_MODE_FP16X2)
    int pair = 0;                   // conv_halo as CTA pairs (tcgen05 cta_group::2): the weight map's box holds Cout / 2 rows
    bool valid = false;
};

struct TcState {
    EncodeTiledFn encode = nullptr;
    int fp16 = 0;                            // 16-bit operand format: 0 = BF16, 1 = FP16
    int split = 0;                           // x3 / x2 modes: every operand is a (hi, lo) pair, second plane right after the first
    int f8 = 0;                              // x2 mode: the lo plane holds FP8 correction operands (tc_common.cuh: split_pack4<.., true>)
    __nv_bfloat16* w[UKBB_N_CONV] = {};      // [planes][cout][taps*cin], K-major
    __nv_bfloat16* wg[UKBB_N_CONV] = {};     // pixel-group layers: expanded [planes][3 * J tiles][64 rows][cin] (conv_group.cuh)
    __nv_bfloat16* wf[UKBB_N_CONV] = {};     // same_dim0 / fc0 / fc1 with the BN scale folded in before rounding: [planes][cout][cin]
    float h_shift[UKBB_N_CONV][64] = {};     // host copies of the folded-BN shifts of those layers (constant-bank operands)
    float h_bias[8] = {};
    float h_wl[64 * 8] = {};                 // class-score weights [k][8] FP32
    float c0_shift[16] = {};                 // conv0_0 folded-BN shift
    __nv_bfloat16* wb0 = nullptr;            // conv_first_tc: expanded conv0_0 weights [64][64] (hi | lo | hi split along K)
    CUtensorMap map_b0;
    __nv_bfloat16* t[5] = {};                // t_l = W_l . s_l at level l (64 channels), l = 1..4: [planes][nb][h_l][w_l][64]
    __nv_bfloat16* u[5] = {};                // interpolation matrices U_l (16-bit, exact)
    TcLayerPlan plan[UKBB_N_CONV];
    int plan_nb = 0, plan_h = 0, plan_w = 0;
    HeadMaps hm;
    SideMaps sm;                             // side_tc_kernel: same_dim_l + fc0 column block of levels 1..4 in one launch
};

// Launch with programmatic stream serialization (tc_common.cuh: griddep_launch / griddep_wait): the next kernel's
// CTAs start their prologue on an SM as soon as the previous kernel's CTA there has exited, instead of after the
// whole grid has drained.
template <int CLUSTER = 1, typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (CLUSTER > 1) {                                   // thread-block clusters along x (TMA multicast of shared operands)
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = CLUSTER; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// tc_conv.cu
int launch_plan(const TcLayerPlan& P, int fp16, int sms, cudaStream_t st);
// tc_first.cu
int launch_first(const TcState* S, const TcLayerPlan& P1, const CUtensorMap& map_img, const ConvFirstParams& fp, int sms, cudaStream_t st);
// per-device "done once" flags of the launchers (function attributes belong to a device)
constexpr int kMaxDevices = 64;
inline bool first_use_on_device(bool (&seen)[kMaxDevices]) {
    int d = 0;
    cudaGetDevice(&d);
    d &= kMaxDevices - 1;
    if (seen[d]) return false;
    seen[d] = true;
    return true;
}

// tc_head.cu (16-bit operands) / tc_head_x3.cu (split operands)
int launch_side_16(const TcState* S, const SideParams& sp, int sms, cudaStream_t st);
int launch_head_16(const TcState* S, const HeadParams& hp, int n_class, int sms, cudaStream_t st);
int launch_side_x3(const TcState* S, const SideParams& sp, int sms, cudaStream_t st);
int launch_head_x3(const TcState* S, const HeadParams& hp, int n_class, int sms, cudaStream_t st);
// tc_head_x2.cu (FP16 pieces + FP8 correction operands)
int launch_side_x2(const TcState* S, const SideParams& sp, int sms, cudaStream_t st);
int launch_head_x2(const TcState* S, const HeadParams& hp, int n_class, int sms, cudaStream_t st);

}  // namespace ukbb
