// conv0_0 + conv0_1 in one launch (conv_first_tc.cuh): launcher of the (format, split scheme) instances.
#include "tc_plan.cuh"

namespace ukbb {

template <bool F16, bool SPLIT, bool F8 = false>
static int launch_first2(const TcState* S, const TcLayerPlan& P1, const CUtensorMap& map_img, const ConvFirstParams& fp, int sms, cudaStream_t st) {
    using Cfg = ConvFirstTcCfg<SPLIT, F8>;
    static bool attr_set[kMaxDevices] = {};          // cudaFuncSetAttribute is per device (one process may drive several: SplitEngine)
    if (first_use_on_device(attr_set)) {
        UKBB_CUDA(cudaFuncSetAttribute(conv_first_tc_kernel<F16, SPLIT, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    const int grid = fp.n_tiles < sms ? fp.n_tiles : sms;
    UKBB_CUDA(launch_pdl(conv_first_tc_kernel<F16, SPLIT, F8>, grid, Cfg::THREADS, Cfg::SMEM_BYTES, st, map_img, P1.map_b, S->map_b0, P1.map_out, fp));
    return UKBB_OK;
}

int launch_first(const TcState* S, const TcLayerPlan& P1, const CUtensorMap& map_img, const ConvFirstParams& fp, int sms, cudaStream_t st) {
    if (S->f8) return launch_first2<true, true, true>(S, P1, map_img, fp, sms, st);
    if (S->split) return S->fp16 ? launch_first2<true, true>(S, P1, map_img, fp, sms, st) : launch_first2<false, true>(S, P1, map_img, fp, sms, st);
    return S->fp16 ? launch_first2<true, false>(S, P1, map_img, fp, sms, st) : launch_first2<false, false>(S, P1, map_img, fp, sms, st);
}

}  // namespace ukbb
