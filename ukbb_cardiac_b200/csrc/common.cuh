// Shared declarations of the ukbb_fcn library (internal).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/ukbb_fcn.h"

namespace ukbb {

void set_error(const char* fmt, ...);

#define UKBB_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ukbb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,               \
                            cudaGetErrorString(_e));                                         \
            return UKBB_E_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define UKBB_REQUIRE(cond, ...)                                                              \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ukbb::set_error(__VA_ARGS__);                                                    \
            return UKBB_E_INVALID;                                                           \
        }                                                                                    \
    } while (0)

// One conv layer on the device.  Weights are re-laid out as [tap][cin][cout] with the
// spatial taps TRANSPOSED relative to TF's HWIO (device rows are Y, columns are X; TF's
// H is X and W is Y): tap = dy*ks + dx holds W_tf[kh=dx][kw=dy].
struct ConvLayer {
    int ksize, cin, cout, stride;
    float* w_f32 = nullptr;       // [ks*ks][cin][cout]
    float* scale = nullptr;       // [cout]  gamma * rsqrt(var + eps)   (1 for logits)
    float* shift = nullptr;       // [cout]  beta - mean * scale        (bias for logits)
    int relu;
};

// ---- FP32 CUDA-core kernels (kernels_fp32.cu) ----
int launch_conv_fp32(const float* in, float* out, const ConvLayer& L, int n, int hi, int wi,
                     int ho, int wo, int pad_top, int pad_left, cudaStream_t st);
int launch_upsample_concat_fp32(const float* const src[5], float* out, int n, int h, int w,
                                cudaStream_t st);
int launch_classifier_fp32(const float* feat, const ConvLayer& L, int n_class, int n, int h2, int w2,
                           int x_pre, int y_pre, int x, int y, uint8_t* labels, float* logits,
                           float* prob, unsigned long long* counts, cudaStream_t st);

// ---- preprocessing (preprocess.cu) ----
struct PreprocWorkspace {
    unsigned int* hist = nullptr;      // 3 passes x 4 ranks x 4096 bins
    unsigned long long* state = nullptr;  // per-rank prefix / residual rank
    double* vlvh = nullptr;            // device (vl, vh)
    float* sel = nullptr;              // 4 selected order statistics
    float* lut = nullptr;              // integer fast path: rescaled value per level (65536 + 2)
    int no_int_path = 0;               // test hook (ukbb_fcn_debug_flags bit 0): always take the generic radix select
};
int preproc_alloc(PreprocWorkspace& ws);
void preproc_free(PreprocWorkspace& ws);
int launch_preprocess(PreprocWorkspace& ws, float* vol, long long n_slices, int x, int y,
                      double q_lo, double q_hi, int x2, int y2, int x_pre, int y_pre, float* out,
                      double* vl_vh_out, int clip_in_place, cudaStream_t st, long long* launches);
int launch_rescale_given(PreprocWorkspace& ws, float* vol, long long n_slices, int x, int y, double vl, double vh, int x2, int y2,
                         int x_pre, int y_pre, float* out, int clip_in_place, cudaStream_t st, long long* launches);

}  // namespace ukbb
