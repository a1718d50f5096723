// side_tc_kernel + head_ts_kernel, plain 16-bit operand instances (UKBB_MODE_BF16 / UKBB_MODE_FP16).
#include "tc_head_impl.cuh"

namespace ukbb {
int launch_side_16(const TcState* S, const SideParams& sp, int sms, cudaStream_t st) { return launch_side_any<false>(S, sp, sms, st); }
int launch_head_16(const TcState* S, const HeadParams& hp, int n_class, int sms, cudaStream_t st) { return launch_head_any<false>(S, hp, n_class, sms, st); }
}  // namespace ukbb
