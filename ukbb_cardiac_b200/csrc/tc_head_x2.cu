// side_tc_kernel + head_ts_kernel, instances of the FP8-correction split scheme (UKBB_MODE_FP16X2).
#include "tc_head_impl.cuh"

namespace ukbb {
int launch_side_x2(const TcState* S, const SideParams& sp, int sms, cudaStream_t st) { return launch_side_any<true, true>(S, sp, sms, st); }
int launch_head_x2(const TcState* S, const HeadParams& hp, int n_class, int sms, cudaStream_t st) { return launch_head_any<true, true>(S, hp, n_class, sms, st); }
}  // namespace ukbb
