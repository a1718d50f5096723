// Engine state behind the opaque ukbb_fcn handle (internal).
#pragma once
#include "common.cuh"
#include <vector>
#include <utility>

namespace ukbb {

struct Workspace {          // activation buffers of one sub-batch of slices
    void* a[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // encoder ping
    void* b[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // encoder pong
    void* s[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // same_dim outputs (32 ch)
    void* cat = nullptr;    // FP32 mode only: materialised 160-channel concat
    void* f0 = nullptr;     // FP32 mode only: fc0 / fc1 outputs
    void* f1 = nullptr;
    int nb = 0, h = 0, w = 0;
};

struct TcState;             // tc_plan.cuh

struct Engine {
    int device = 0, mode = 0, n_class = 0, sms = 148;
    ConvLayer layers[UKBB_N_CONV];
    Workspace ws;
    PreprocWorkspace pre;
    TcState* tc = nullptr;
    unsigned long long* d_counts = nullptr;
    int counts_cap = 0, counts_n = 0;
    long long launches = 0;
    bool ktimer = false;                                  // ukbb_fcn_kernel_timer
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ktimer_ev;
    // whole-subject staging (ukbb_fcn_segment_host): two slots so H2D / compute / D2H overlap
    float* st_vol[2] = {nullptr, nullptr};
    uint8_t* st_labels[2] = {nullptr, nullptr};
    double* st_vlvh[2] = {nullptr, nullptr};
    long long* st_counts[2] = {nullptr, nullptr};
    size_t st_cap[2] = {0, 0};
    float* st_pad[2] = {nullptr, nullptr};               // rescaled + padded volume per slot
    size_t st_pad_cap[2] = {0, 0};
    // preprocessing of subject s + 1 (own stream) overlaps the forward of subject s (caller's stream)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_pre = nullptr;
    cudaEvent_t ev_ws = nullptr;                          // last use of the (single) selection workspace `pre`, on whichever stream
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr}, ev_pre[2] = {nullptr, nullptr};
};

// tc_forward.cu: tcgen05 path (BF16 / FP16 and the split-operand x3 modes)
int tc_prepare(Engine* h, const ukbb_fcn_weights* w);
void tc_release(Engine* h);
int forward_tc(Engine* h, const float* image, int n, int x2, int y2, int x_pre, int y_pre, int x, int y,
               uint8_t* labels, float* logits, float* prob, unsigned long long* counts, cudaStream_t st);

int debug_conv_tc(Engine* h, int li, const void* in, int n, int hi, int wi, int level_out, void* out, cudaStream_t st);
int debug_read_tc(Engine* h, int which, int level, float* out, long long n_elems, cudaStream_t st);

}  // namespace ukbb
