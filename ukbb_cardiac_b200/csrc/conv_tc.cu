// BF16 tcgen05 path (placeholder until the tensor-core kernels land).
#include "engine.cuh"
namespace ukbb {
int bf16_prepare(Engine*, const ukbb_fcn_weights*) {
    set_error("BF16 tensor-core mode is not built yet");
    return UKBB_E_UNSUPPORTED;
}
void bf16_release(Engine*) {}
int forward_bf16(Engine*, const float*, int, int, int, int, int, int, int, uint8_t*, float*, float*,
                 unsigned long long*, cudaStream_t) {
    set_error("BF16 tensor-core mode is not built yet");
    return UKBB_E_UNSUPPORTED;
}
}  // namespace ukbb
