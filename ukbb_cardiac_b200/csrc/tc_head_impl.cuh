// Launchers of side_tc_kernel and head_ts_kernel for one value of SPLIT (included by tc_head.cu and tc_head_x3.cu so that
// the two sets of instances compile in parallel).
#pragma once
#include "tc_plan.cuh"
#include "head_ts.cuh"

namespace ukbb {

template <int NC, bool F16, bool SPLIT, bool F8 = false>
static int launch_head_ts2(const TcState* S, const HeadParams& hp, int sms, cudaStream_t st) {
    using Cfg = HeadTsCfg<SPLIT>;
    static bool attr_set[kMaxDevices] = {};          // cudaFuncSetAttribute is per device (one process may drive several: SplitEngine)
    if (first_use_on_device(attr_set)) {
        UKBB_CUDA(cudaFuncSetAttribute(head_ts_kernel<NC, F16, SPLIT, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    }
    const int grid = hp.n_tiles < sms ? hp.n_tiles : sms;
    UKBB_CUDA(launch_pdl(head_ts_kernel<NC, F16, SPLIT, F8>, grid, H4_THREADS, Cfg::SMEM, st, S->hm, hp));
    return UKBB_OK;
}

template <int NC, bool SPLIT, bool F8 = false>
static int launch_head_nc(const TcState* S, const HeadParams& hp, int sms, cudaStream_t st) {
    if (F8) return launch_head_ts2<NC, true, true, F8>(S, hp, sms, st);          // the FP8 correction scheme exists for FP16 pieces only
    return S->fp16 ? launch_head_ts2<NC, true, SPLIT>(S, hp, sms, st) : launch_head_ts2<NC, false, SPLIT>(S, hp, sms, st);
}

template <bool SPLIT, bool F8 = false>
static int launch_head_any(const TcState* S, const HeadParams& hp, int n_class, int sms, cudaStream_t st) {
    switch (n_class) {
        case 2: return launch_head_nc<2, SPLIT, F8>(S, hp, sms, st);
        case 3: return launch_head_nc<3, SPLIT, F8>(S, hp, sms, st);
        case 4: return launch_head_nc<4, SPLIT, F8>(S, hp, sms, st);
        case 5: case 6: return launch_head_nc<6, SPLIT, F8>(S, hp, sms, st);
        default: return launch_head_nc<8, SPLIT, F8>(S, hp, sms, st);
    }
}

template <bool F16, bool SPLIT, bool F8 = false>
static int launch_side2(const TcState* S, const SideParams& sp, int sms, cudaStream_t st) {
    constexpr int SMEM = SPLIT ? SD_SMEM_SPLIT : SD_SMEM;
    static bool attr_set[kMaxDevices] = {};          // cudaFuncSetAttribute is per device (one process may drive several: SplitEngine)
    if (first_use_on_device(attr_set)) {
        UKBB_CUDA(cudaFuncSetAttribute(side_tc_kernel<F16, SPLIT, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    }
    const int n_tiles = sp.tile_start[4];
    const int grid = n_tiles < sms ? n_tiles : sms;
    UKBB_CUDA(launch_pdl(side_tc_kernel<F16, SPLIT, F8>, grid, SD_THREADS, SMEM, st, S->sm, sp));
    return UKBB_OK;
}

template <bool SPLIT, bool F8 = false>
static int launch_side_any(const TcState* S, const SideParams& sp, int sms, cudaStream_t st) {
    if (F8) return launch_side2<true, true, F8>(S, sp, sms, st);
    return S->fp16 ? launch_side2<true, SPLIT>(S, sp, sms, st) : launch_side2<false, SPLIT>(S, sp, sms, st);
}

}  // namespace ukbb
