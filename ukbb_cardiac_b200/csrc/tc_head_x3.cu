// side_tc_kernel + head_ts_kernel, split-operand instances (UKBB_MODE_BF16X3 / UKBB_MODE_FP16X3).
#include "tc_head_impl.cuh"

namespace ukbb {
int launch_side_x3(const TcState* S, const SideParams& sp, int sms, cudaStream_t st) { return launch_side_any<true>(S, sp, sms, st); }
int launch_head_x3(const TcState* S, const HeadParams& hp, int n_class, int sms, cudaStream_t st) { return launch_head_any<true>(S, hp, n_class, sms, st); }
}  // namespace ukbb
