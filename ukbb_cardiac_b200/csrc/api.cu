// C-ABI entry points of libukbb_fcn (see include/ukbb_fcn.h for the contract and the
// reference interfaces each one replaces).
#include "common.cuh"
#include "engine.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <new>

namespace ukbb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static void same_pad(int in, int k, int s, int* out, int* before) {
    *out = (in + s - 1) / s;
    int total = (*out - 1) * s + k - in;
    if (total < 0) total = 0;
    *before = total / 2;
}

// Topology of build_FCN (network.py:170-230, train_network.py:174-195) used for validation.
static const int kNFilter[5] = {16, 32, 64, 128, 256};
static const int kNBlock[5] = {2, 2, 3, 3, 3};

static int expected_layer(int i, int n_class, int* ks, int* cin, int* cout, int* stride) {
    int idx = 0, c = 1;
    for (int l = 0; l < 5; ++l)
        for (int b = 0; b < kNBlock[l]; ++b, ++idx) {
            if (idx == i) { *ks = 3; *cin = c; *cout = kNFilter[l]; *stride = (l > 0 && b == 0) ? 2 : 1; return 0; }
            c = kNFilter[l];
        }
    for (int l = 0; l < 5; ++l, ++idx)
        if (idx == i) { *ks = 1; *cin = kNFilter[l]; *cout = 32; *stride = 1; return 0; }
    if (i == idx) { *ks = 1; *cin = 160; *cout = 64; *stride = 1; return 0; }
    if (i == idx + 1) { *ks = 1; *cin = 64; *cout = 64; *stride = 1; return 0; }
    if (i == idx + 2) { *ks = 1; *cin = 64; *cout = n_class; *stride = 1; return 0; }
    return -1;
}

static int upload_layer(ConvLayer& L, const ukbb_conv_weights& w, float eps, bool last) {
    L.ksize = w.ksize; L.cin = w.cin; L.cout = w.cout; L.stride = w.stride; L.relu = last ? 0 : 1;
    const int taps = w.ksize * w.ksize;
    std::vector<float> wt((size_t)taps * w.cin * w.cout), sc(w.cout), sh(w.cout);
    // device tap (dy, dx) <- TF kernel[kh = dx][kw = dy]  (device rows are Y = TF's W axis)
    for (int dy = 0; dy < w.ksize; ++dy)
        for (int dx = 0; dx < w.ksize; ++dx)
            memcpy(&wt[(size_t)(dy * w.ksize + dx) * w.cin * w.cout],
                   w.kernel + (size_t)(dx * w.ksize + dy) * w.cin * w.cout, sizeof(float) * w.cin * w.cout);
    for (int c = 0; c < w.cout; ++c) {
        if (w.gamma) {
            const double s = (double)w.gamma[c] / sqrt((double)w.moving_variance[c] + (double)eps);
            sc[c] = (float)s;
            sh[c] = (float)((double)w.beta[c] - (double)w.moving_mean[c] * s);
        } else {
            sc[c] = 1.f;
            sh[c] = w.bias ? w.bias[c] : 0.f;
        }
    }
    UKBB_CUDA(cudaMalloc(&L.w_f32, wt.size() * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&L.scale, sc.size() * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&L.shift, sh.size() * sizeof(float)));
    UKBB_CUDA(cudaMemcpy(L.w_f32, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    UKBB_CUDA(cudaMemcpy(L.scale, sc.data(), sc.size() * sizeof(float), cudaMemcpyHostToDevice));
    UKBB_CUDA(cudaMemcpy(L.shift, sh.data(), sh.size() * sizeof(float), cudaMemcpyHostToDevice));
    return UKBB_OK;
}

// ---------------------------------------------------------------------------------------------
// FP32 workspace + forward
// ---------------------------------------------------------------------------------------------
static void free_ws(Engine* h) {
    for (int l = 0; l < 5; ++l) {
        cudaFree(h->ws.a[l]); cudaFree(h->ws.b[l]); cudaFree(h->ws.s[l]);
        h->ws.a[l] = h->ws.b[l] = h->ws.s[l] = nullptr;
    }
    cudaFree(h->ws.cat); cudaFree(h->ws.f0); cudaFree(h->ws.f1);
    h->ws.cat = h->ws.f0 = h->ws.f1 = nullptr;
    h->ws.nb = h->ws.h = h->ws.w = 0;
}

static int ensure_ws(Engine* h, int nb, int hh, int ww) {
    if (h->ws.nb >= nb && h->ws.h == hh && h->ws.w == ww) return UKBB_OK;
    UKBB_CUDA(cudaDeviceSynchronize());
    free_ws(h);
    const size_t esz = 4;                                 // FP32 mode only (the tensor-core modes own their workspace: tc_forward.cu)
    for (int l = 0; l < 5; ++l) {
        const size_t px = (size_t)nb * (hh >> l) * (ww >> l);
        UKBB_CUDA(cudaMalloc(&h->ws.a[l], px * kNFilter[l] * esz));
        UKBB_CUDA(cudaMalloc(&h->ws.b[l], px * kNFilter[l] * esz));
        UKBB_CUDA(cudaMalloc(&h->ws.s[l], px * 32 * esz));
    }
    const size_t px = (size_t)nb * hh * ww;
    UKBB_CUDA(cudaMalloc(&h->ws.cat, px * 160 * esz));
    UKBB_CUDA(cudaMalloc(&h->ws.f0, px * 64 * esz));
    UKBB_CUDA(cudaMalloc(&h->ws.f1, px * 64 * esz));
    h->ws.nb = nb; h->ws.h = hh; h->ws.w = ww;
    return UKBB_OK;
}

static int forward_fp32(Engine* h, const float* image, int n, int w2, int h2, int x_pre, int y_pre, int x,
                        int y, uint8_t* labels, float* logits, float* prob, unsigned long long* counts,
                        cudaStream_t st) {
    const int NB = n < 16 ? n : 16;
    int rc = ensure_ws(h, NB, h2, w2);
    if (rc) return rc;
    for (int n0 = 0; n0 < n; n0 += NB) {
        const int nb = n - n0 < NB ? n - n0 : NB;
        const float* cur = image + (size_t)n0 * h2 * w2;
        int hi = h2, wi = w2, li = 0;
        const float* level_out[5];
        for (int l = 0; l < 5; ++l) {
            for (int b = 0; b < kNBlock[l]; ++b, ++li) {
                const ConvLayer& L = h->layers[li];
                int ho, wo, pt, pl;
                same_pad(hi, L.ksize, L.stride, &ho, &pt);
                same_pad(wi, L.ksize, L.stride, &wo, &pl);
                float* dst = (float*)((b & 1) ? h->ws.b[l] : h->ws.a[l]);
                rc = launch_conv_fp32(cur, dst, L, nb, hi, wi, ho, wo, pt, pl, st);
                if (rc) return rc;
                h->launches++;
                cur = dst; hi = ho; wi = wo;
            }
            level_out[l] = cur;
        }
        const float* sd[5];
        for (int l = 0; l < 5; ++l, ++li) {
            rc = launch_conv_fp32(level_out[l], (float*)h->ws.s[l], h->layers[li], nb, h2 >> l, w2 >> l, h2 >> l,
                                  w2 >> l, 0, 0, st);
            if (rc) return rc;
            h->launches++;
            sd[l] = (const float*)h->ws.s[l];
        }
        rc = launch_upsample_concat_fp32(sd, (float*)h->ws.cat, nb, h2, w2, st);
        if (rc) return rc;
        rc = launch_conv_fp32((float*)h->ws.cat, (float*)h->ws.f0, h->layers[18], nb, h2, w2, h2, w2, 0, 0, st);
        if (rc) return rc;
        rc = launch_conv_fp32((float*)h->ws.f0, (float*)h->ws.f1, h->layers[19], nb, h2, w2, h2, w2, 0, 0, st);
        if (rc) return rc;
        const size_t po = (size_t)n0 * h2 * w2 * h->n_class;
        rc = launch_classifier_fp32((float*)h->ws.f1, h->layers[20], h->n_class, nb, h2, w2, x_pre, y_pre, x, y,
                                    labels + (size_t)n0 * x * y, logits ? logits + po : nullptr,
                                    prob ? prob + po : nullptr, counts ? counts + (size_t)n0 * h->n_class : nullptr, st);
        if (rc) return rc;
        h->launches += 4;
    }
    return UKBB_OK;
}

static int ensure_counts(Engine* h, int n) {
    if (h->counts_cap >= n) return UKBB_OK;
    UKBB_CUDA(cudaDeviceSynchronize());
    cudaFree(h->d_counts);
    h->d_counts = nullptr;
    UKBB_CUDA(cudaMalloc(&h->d_counts, (size_t)n * UKBB_MAX_CLASS * sizeof(unsigned long long)));
    h->counts_cap = n;
    return UKBB_OK;
}

static int forward_any(Engine* h, const float* image, int n, int x2, int y2, int x_pre, int y_pre, int x, int y,
                       uint8_t* labels, float* logits, float* prob, cudaStream_t st) {
    UKBB_REQUIRE(h && image && labels, "forward: null handle / image / labels");
    UKBB_REQUIRE(n > 0 && x2 > 0 && y2 > 0 && x2 % 16 == 0 && y2 % 16 == 0,
                 "forward: padded size %dx%d must be positive multiples of 16 (n=%d)", x2, y2, n);
    UKBB_REQUIRE(x > 0 && y > 0 && x_pre >= 0 && y_pre >= 0 && x_pre + x <= x2 && y_pre + y <= y2,
                 "forward: crop (%d,%d)+(%d,%d) outside padded %dx%d", x_pre, y_pre, x, y, x2, y2);
    UKBB_CUDA(cudaSetDevice(h->device));
    int rc = ensure_counts(h, n);
    if (rc) return rc;
    UKBB_CUDA(cudaMemsetAsync(h->d_counts, 0, (size_t)n * h->n_class * sizeof(unsigned long long), st));
    h->counts_n = n;
    if (h->mode == UKBB_MODE_FP32)
        return forward_fp32(h, image, n, x2, y2, x_pre, y_pre, x, y, labels, logits, prob, h->d_counts, st);
    return forward_tc(h, image, n, x2, y2, x_pre, y_pre, x, y, labels, logits, prob, h->d_counts, st);
}

}  // namespace ukbb

using namespace ukbb;

extern "C" {

const char* ukbb_last_error(void) { return g_err; }
const char* ukbb_version(void) { return "ukbb_fcn 0.1 (sm_100a)"; }

int ukbb_fcn_create(const ukbb_fcn_weights* w, int n_class, int device, int mode, ukbb_fcn** out) {
    UKBB_REQUIRE(w && out, "create: null argument");
    *out = nullptr;
    UKBB_REQUIRE(w->n_conv == UKBB_N_CONV && w->conv, "create: expected %d conv layers, got %d", UKBB_N_CONV, w->n_conv);
    UKBB_REQUIRE(n_class >= 2 && n_class <= UKBB_MAX_CLASS, "create: n_class=%d not in [2,%d]", n_class, UKBB_MAX_CLASS);
    UKBB_REQUIRE(mode == UKBB_MODE_FP32 || mode == UKBB_MODE_BF16 || mode == UKBB_MODE_FP16 || mode == UKBB_MODE_BF16X3 || mode == UKBB_MODE_FP16X3 ||
                 mode == UKBB_MODE_FP16X2,
                 "create: unknown mode %d", mode);
    for (int i = 0; i < UKBB_N_CONV; ++i) {
        int ks, cin, cout, stride;
        expected_layer(i, n_class, &ks, &cin, &cout, &stride);
        const ukbb_conv_weights& c = w->conv[i];
        UKBB_REQUIRE(c.kernel, "create: layer %d has no kernel", i);
        UKBB_REQUIRE(c.ksize == ks && c.cin == cin && c.cout == cout && c.stride == stride,
                     "create: layer %d is %dx%d %d->%d stride %d, build_FCN expects %dx%d %d->%d stride %d", i,
                     c.ksize, c.ksize, c.cin, c.cout, c.stride, ks, ks, cin, cout, stride);
        const bool last = i == UKBB_N_CONV - 1;
        UKBB_REQUIRE(last ? (c.bias && !c.gamma) : (c.gamma && c.beta && c.moving_mean && c.moving_variance),
                     "create: layer %d %s", i, last ? "needs a bias and no batch norm" : "needs batch-norm parameters");
    }
    int ndev = 0;
    UKBB_CUDA(cudaGetDeviceCount(&ndev));
    UKBB_REQUIRE(device >= 0 && device < ndev, "create: device %d out of range (%d CUDA devices)", device, ndev);
    UKBB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    UKBB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (mode != UKBB_MODE_FP32 && prop.major != 10) {
        set_error("create: tensor-core mode needs an sm_100 device, device %d is sm_%d%d", device, prop.major, prop.minor);
        return UKBB_E_UNSUPPORTED;
    }
    Engine* h = new (std::nothrow) Engine();
    if (!h) { set_error("create: out of host memory"); return UKBB_E_NOMEM; }
    h->device = device; h->mode = mode; h->n_class = n_class; h->sms = prop.multiProcessorCount;
    int rc = UKBB_OK;
    for (int i = 0; i < UKBB_N_CONV && !rc; ++i) rc = upload_layer(h->layers[i], w->conv[i], w->bn_eps, i == UKBB_N_CONV - 1);
    if (!rc) rc = preproc_alloc(h->pre);
    if (!rc && mode != UKBB_MODE_FP32) rc = tc_prepare(h, w);
    if (!rc) {
        cudaError_t e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_pre, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_ws, cudaEventDisableTiming);
        for (int s = 0; s < 2 && e == cudaSuccess; ++s) {
            e = cudaEventCreateWithFlags(&h->ev_h2d[s], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_compute[s], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_d2h[s], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pre[s], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) { set_error("create: stream/event creation failed: %s", cudaGetErrorString(e)); rc = UKBB_E_CUDA; }
    }
    if (rc) { ukbb_fcn_destroy(reinterpret_cast<ukbb_fcn*>(h)); return rc; }
    *out = reinterpret_cast<ukbb_fcn*>(h);
    return UKBB_OK;
}

void ukbb_fcn_destroy(ukbb_fcn* hh) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < UKBB_N_CONV; ++i) {
        cudaFree(h->layers[i].w_f32); cudaFree(h->layers[i].scale); cudaFree(h->layers[i].shift);
    }
    tc_release(h);
    free_ws(h);
    preproc_free(h->pre);
    cudaFree(h->d_counts);
    for (int s = 0; s < 2; ++s) {
        cudaFree(h->st_vol[s]); cudaFree(h->st_labels[s]); cudaFree(h->st_vlvh[s]); cudaFree(h->st_counts[s]);
        if (h->ev_h2d[s]) cudaEventDestroy(h->ev_h2d[s]);
        if (h->ev_compute[s]) cudaEventDestroy(h->ev_compute[s]);
        if (h->ev_d2h[s]) cudaEventDestroy(h->ev_d2h[s]);
        if (h->ev_pre[s]) cudaEventDestroy(h->ev_pre[s]);
        cudaFree(h->st_pad[s]);
    }
    if (h->ev_ws) cudaEventDestroy(h->ev_ws);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->s_pre) cudaStreamDestroy(h->s_pre);
    delete h;
}

int ukbb_fcn_forward(ukbb_fcn* hh, const float* image, int n, int x2, int y2, int x_pre, int y_pre, int x, int y,
                     uint8_t* labels, float* logits, float* prob, void* stream) {
    return forward_any(reinterpret_cast<Engine*>(hh), image, n, x2, y2, x_pre, y_pre, x, y, labels, logits, prob,
                       (cudaStream_t)stream);
}

int ukbb_fcn_preprocess(ukbb_fcn* hh, float* vol, long long n_slices, int x, int y, double q_lo, double q_hi, int x2,
                        int y2, int x_pre, int y_pre, float* out, double* vl_vh, int clip_in_place, void* stream) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h && vol && out, "preprocess: null argument");
    UKBB_REQUIRE(n_slices > 0 && x > 0 && y > 0, "preprocess: empty volume (%lld slices of %dx%d)", n_slices, x, y);
    UKBB_CUDA(cudaSetDevice(h->device));
    // the handle owns ONE selection workspace (histograms, scan state, lookup table): order this call after its previous user,
    // whichever stream that was (another caller stream, or the internal stream of ukbb_fcn_segment_host)
    UKBB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_ws, 0));
    const int rc = launch_preprocess(h->pre, vol, n_slices, x, y, q_lo, q_hi, x2, y2, x_pre, y_pre, out, vl_vh, clip_in_place,
                                     (cudaStream_t)stream, &h->launches);
    UKBB_CUDA(cudaEventRecord(h->ev_ws, (cudaStream_t)stream));
    return rc;
}

int ukbb_fcn_rescale(ukbb_fcn* hh, float* vol, long long n_slices, int x, int y, double vl, double vh, int x2, int y2, int x_pre,
                     int y_pre, float* out, int clip_in_place, void* stream) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h && vol && out, "rescale: null argument");
    UKBB_CUDA(cudaSetDevice(h->device));
    UKBB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_ws, 0));       // the selection workspace holds the thresholds
    const int rc = launch_rescale_given(h->pre, vol, n_slices, x, y, vl, vh, x2, y2, x_pre, y_pre, out, clip_in_place, (cudaStream_t)stream,
                                        &h->launches);
    UKBB_CUDA(cudaEventRecord(h->ev_ws, (cudaStream_t)stream));
    return rc;
}

int ukbb_fcn_class_counts(ukbb_fcn* hh, long long* counts, int n, void* stream) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h && counts, "class_counts: null argument");
    UKBB_REQUIRE(n > 0 && n <= h->counts_n, "class_counts: n=%d but the last forward had %d slices", n, h->counts_n);
    UKBB_CUDA(cudaSetDevice(h->device));
    UKBB_CUDA(cudaMemcpyAsync(counts, h->d_counts, (size_t)n * h->n_class * sizeof(long long), cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream));
    return UKBB_OK;
}

int ukbb_fcn_segment_host(ukbb_fcn* hh, const float* vol, int x, int y, int z, int t, double q_lo, double q_hi,
                          uint8_t* labels, double* vl_vh, long long* counts, int slot, void* stream) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h && vol && labels, "segment_host: null argument");
    UKBB_REQUIRE(x > 0 && y > 0 && z > 0 && t > 0, "segment_host: empty volume %dx%dx%dx%d", x, y, z, t);
    UKBB_REQUIRE(slot == 0 || slot == 1, "segment_host: slot must be 0 or 1");
    UKBB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)z * t;
    const size_t vox = (size_t)n * x * y;
    const int x2 = (x + 15) / 16 * 16, y2 = (y + 15) / 16 * 16;
    const int x_pre = (x2 - x) / 2, y_pre = (y2 - y) / 2;       // deploy_network.py:98
    const size_t padvox = (size_t)n * x2 * y2;
    if (h->st_cap[slot] < vox) {
        UKBB_CUDA(cudaDeviceSynchronize());
        cudaFree(h->st_vol[slot]); cudaFree(h->st_labels[slot]); cudaFree(h->st_counts[slot]);
        h->st_vol[slot] = nullptr; h->st_labels[slot] = nullptr; h->st_counts[slot] = nullptr; h->st_cap[slot] = 0;
        UKBB_CUDA(cudaMalloc(&h->st_vol[slot], vox * sizeof(float)));
        UKBB_CUDA(cudaMalloc(&h->st_labels[slot], vox));
        UKBB_CUDA(cudaMalloc(&h->st_counts[slot], (size_t)n * UKBB_MAX_CLASS * sizeof(long long)));
        if (!h->st_vlvh[slot]) UKBB_CUDA(cudaMalloc(&h->st_vlvh[slot], 2 * sizeof(double)));
        h->st_cap[slot] = vox;
    }
    if (h->st_pad_cap[slot] < padvox) {
        UKBB_CUDA(cudaDeviceSynchronize());
        cudaFree(h->st_pad[slot]); h->st_pad[slot] = nullptr; h->st_pad_cap[slot] = 0;
        UKBB_CUDA(cudaMalloc(&h->st_pad[slot], padvox * sizeof(float)));
        h->st_pad_cap[slot] = padvox;
    }
    // H2D once the previous occupant of this slot's raw volume has been preprocessed
    UKBB_CUDA(cudaStreamWaitEvent(h->s_h2d, h->ev_pre[slot], 0));
    UKBB_CUDA(cudaMemcpyAsync(h->st_vol[slot], vol, vox * sizeof(float), cudaMemcpyHostToDevice, h->s_h2d));
    UKBB_CUDA(cudaEventRecord(h->ev_h2d[slot], h->s_h2d));
    // preprocessing on its own stream (calls are serialised there: one selection workspace per handle); it may run while the
    // caller's stream is still busy with the forward of the previous subject (other slot)
    UKBB_CUDA(cudaStreamWaitEvent(h->s_pre, h->ev_h2d[slot], 0));
    UKBB_CUDA(cudaStreamWaitEvent(h->s_pre, h->ev_compute[slot], 0));    // forward of this slot's previous occupant has read st_pad[slot]
    UKBB_CUDA(cudaStreamWaitEvent(h->s_pre, h->ev_d2h[slot], 0));        // ... and its (vl, vh) have been read back
    UKBB_CUDA(cudaStreamWaitEvent(h->s_pre, h->ev_ws, 0));               // a public ukbb_fcn_preprocess call may be using the workspace
    int rc = launch_preprocess(h->pre, h->st_vol[slot], n, x, y, q_lo, q_hi, x2, y2, x_pre, y_pre, h->st_pad[slot],
                               h->st_vlvh[slot], 0, h->s_pre, &h->launches);
    if (rc) return rc;
    UKBB_CUDA(cudaEventRecord(h->ev_ws, h->s_pre));
    UKBB_CUDA(cudaEventRecord(h->ev_pre[slot], h->s_pre));
    UKBB_CUDA(cudaStreamWaitEvent(st, h->ev_pre[slot], 0));
    UKBB_CUDA(cudaStreamWaitEvent(st, h->ev_d2h[slot], 0));      // label staging of this slot drained
    rc = forward_any(h, h->st_pad[slot], (int)n, x2, y2, x_pre, y_pre, x, y, h->st_labels[slot], nullptr, nullptr, st);
    if (rc) return rc;
    if (counts)
        UKBB_CUDA(cudaMemcpyAsync(h->st_counts[slot], h->d_counts, (size_t)n * h->n_class * sizeof(long long),
                                  cudaMemcpyDeviceToDevice, st));
    UKBB_CUDA(cudaEventRecord(h->ev_compute[slot], st));
    UKBB_CUDA(cudaStreamWaitEvent(h->s_d2h, h->ev_compute[slot], 0));
    UKBB_CUDA(cudaMemcpyAsync(labels, h->st_labels[slot], vox, cudaMemcpyDeviceToHost, h->s_d2h));
    if (vl_vh) UKBB_CUDA(cudaMemcpyAsync(vl_vh, h->st_vlvh[slot], 2 * sizeof(double), cudaMemcpyDeviceToHost, h->s_d2h));
    if (counts)
        UKBB_CUDA(cudaMemcpyAsync(counts, h->st_counts[slot], (size_t)n * h->n_class * sizeof(long long),
                                  cudaMemcpyDeviceToHost, h->s_d2h));
    UKBB_CUDA(cudaEventRecord(h->ev_d2h[slot], h->s_d2h));
    return UKBB_OK;
}

int ukbb_fcn_join(ukbb_fcn* hh, void* stream) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h, "join: null handle");
    UKBB_CUDA(cudaSetDevice(h->device));
    for (int s = 0; s < 2; ++s) UKBB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_d2h[s], 0));
    return UKBB_OK;
}

int ukbb_fcn_sync(ukbb_fcn* hh) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h, "sync: null handle");
    UKBB_CUDA(cudaSetDevice(h->device));
    UKBB_CUDA(cudaDeviceSynchronize());
    return UKBB_OK;
}

int ukbb_fcn_debug_conv(ukbb_fcn* hh, int layer, const void* in_bf16, int n, int hi, int wi, int level_out,
                        void* out_bf16, void* stream) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h && in_bf16 && out_bf16, "debug_conv: null argument");
    UKBB_REQUIRE(h->mode != UKBB_MODE_FP32, "debug_conv: engine is not in a tensor-core mode");
    UKBB_CUDA(cudaSetDevice(h->device));
    return debug_conv_tc(h, layer, in_bf16, n, hi, wi, level_out, out_bf16, (cudaStream_t)stream);
}

int ukbb_fcn_debug_read(ukbb_fcn* hh, int which, int level, float* out_f32, long long n_elems, void* stream) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h && out_f32, "debug_read: null argument");
    UKBB_REQUIRE(h->mode != UKBB_MODE_FP32, "debug_read: engine is not in a tensor-core mode");
    UKBB_CUDA(cudaSetDevice(h->device));
    return debug_read_tc(h, which, level, out_f32, n_elems, (cudaStream_t)stream);
}

int ukbb_fcn_debug_flags(ukbb_fcn* hh, int flags) {
    Engine* h = reinterpret_cast<Engine*>(hh);
    UKBB_REQUIRE(h, "debug_flags: null handle");
    h->pre.no_int_path = (flags & 1) ? 1 : 0;
    return UKBB_OK;
}

int ukbb_fcn_kernel_timer(ukbb_fcn* hh, int enable) {
    ukbb::Engine* h = reinterpret_cast<ukbb::Engine*>(hh);
    if (!h) { ukbb::set_error("kernel_timer: null handle"); return UKBB_E_INVALID; }
    h->ktimer = enable != 0;
    return UKBB_OK;
}

int ukbb_fcn_kernel_timer_read(ukbb_fcn* hh, double* total_ms, long long* launches) {
    ukbb::Engine* h = reinterpret_cast<ukbb::Engine*>(hh);
    if (!h || !total_ms || !launches) { ukbb::set_error("kernel_timer_read: null argument"); return UKBB_E_INVALID; }
    UKBB_CUDA(cudaSetDevice(h->device));
    UKBB_CUDA(cudaDeviceSynchronize());
    double ms = 0.0;
    for (auto& e : h->ktimer_ev) {
        float t = 0.f;
        UKBB_CUDA(cudaEventElapsedTime(&t, e.first, e.second));
        ms += t;
        cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    *total_ms = ms;
    *launches = (long long)h->ktimer_ev.size();
    h->ktimer_ev.clear();
    return UKBB_OK;
}

long long ukbb_fcn_launch_count(const ukbb_fcn* hh) { return hh ? reinterpret_cast<const Engine*>(hh)->launches : 0; }
int ukbb_fcn_mode(const ukbb_fcn* hh) { return hh ? reinterpret_cast<const Engine*>(hh)->mode : -1; }
int ukbb_fcn_n_class(const ukbb_fcn* hh) { return hh ? reinterpret_cast<const Engine*>(hh)->n_class : -1; }

}  // extern "C"
