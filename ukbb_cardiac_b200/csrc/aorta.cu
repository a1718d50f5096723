// Aortic cine segmentation: UNet + bidirectional ConvLSTM deploy path (BASELINE config C5), FP32 CUDA cores.
//   common/network_ao.py:18-64      UNet (conv + BN + ReLU blocks, learned 3x3 stride-2 transposed convs, skip concat)
//   common/network_ao.py:255-319    BiConv_LSTM (tf.contrib.rnn.Conv2DLSTMCell forward + backward, 1x1 output conv)
//   common/deploy_network_ao.py:130-183  circular window of 2R-1 frames around every frame, weighted overlap-add of the
//                                   window probabilities, argmax
// Three things make this a different program from the reference loop (which feeds every 9-frame window through the whole graph):
//   * the UNet features of a frame do not depend on the window it appears in: they are computed ONCE per frame (the reference
//     recomputes them 9 times: 10.1 -> 5.0 TFLOP per 100-frame sequence);
//   * the ConvLSTM's convolution of concat([x, h]) is split into conv(x) + conv(h): conv(x) (+ bias) is evaluated once per frame and
//     direction, only conv(h) runs inside the recurrence (another 2.2 TFLOP saved), and the first step (h = 0) needs no conv at all;
//   * all T windows advance through the recurrence in lockstep as one batch (T images per launch instead of 1), and the
//     window softmax, the weighted overlap-add in the reference's accumulation order and arithmetic (float32 accumulator, float64
//     products), the normalisation and the argmax / crop are one gather kernel per frame: no atomics, deterministic.
// This path is an exactness-mode implementation (FP32 CUDA cores, conv_fp32_kernel): correct first; tensor-core kernels are the
// next step for it.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <vector>
#include <new>

namespace ukbb {

// ---------------------------------------------------------------------------------------------------- kernels
// 3x3 stride-2 transposed convolution (tf.layers.conv2d_transpose, SAME: big[y] = sum_{i,k: 2i+k=y} small[i] w[k]) + folded BN + ReLU.
// One thread = one output pixel x 16 output channels; weights [tap][cin][cout].
__global__ void __launch_bounds__(128)
convT_fp32_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wt, const float* __restrict__ scale,
                  const float* __restrict__ shift, int cin, int cout, int hi, int wi) {
    const int ho = 2 * hi, wo = 2 * wi;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cb0 = blockIdx.y * 16;
    const int n = blockIdx.z;
    if (pix >= (long long)ho * wo) return;
    const int oy = (int)(pix / wo), ox = (int)(pix % wo);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int dy = (oy & 1); dy < 3; dy += 2) {
        const int iy = (oy - dy) >> 1;
        if (oy - dy < 0 || iy >= hi) continue;
        for (int dx = (ox & 1); dx < 3; dx += 2) {
            const int ix = (ox - dx) >> 1;
            if (ox - dx < 0 || ix >= wi) continue;
            const float* src = in + (((size_t)n * hi + iy) * wi + ix) * cin;
            const float* w = wt + (size_t)(dy * 3 + dx) * cin * cout + cb0;
            for (int c = 0; c < cin; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(src + c);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4* w4 = reinterpret_cast<const float4*>(w + (size_t)(c + u) * cout);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 ww = __ldg(w4 + q);
                        acc[4 * q + 0] = fmaf(vv[u], ww.x, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(vv[u], ww.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(vv[u], ww.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(vv[u], ww.w, acc[4 * q + 3]);
                    }
                }
            }
        }
    }
    float* o = out + (((size_t)n * ho + oy) * wo + ox) * cout + cb0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 r;
        r.x = fmaxf(fmaf(acc[4 * q + 0], scale[cb0 + 4 * q + 0], shift[cb0 + 4 * q + 0]), 0.f);
        r.y = fmaxf(fmaf(acc[4 * q + 1], scale[cb0 + 4 * q + 1], shift[cb0 + 4 * q + 1]), 0.f);
        r.z = fmaxf(fmaf(acc[4 * q + 2], scale[cb0 + 4 * q + 2], shift[cb0 + 4 * q + 2]), 0.f);
        r.w = fmaxf(fmaf(acc[4 * q + 3], scale[cb0 + 4 * q + 3], shift[cb0 + 4 * q + 3]), 0.f);
        reinterpret_cast<float4*>(o)[q] = r;
    }
}

// out[p] = [a[p] (ca channels) | b[p] (cb channels)]   (tf.concat([net['conv_l'], x], axis=-1), network_ao.py:52)
__global__ void concat2_kernel(const float4* __restrict__ a, int ca4, const float4* __restrict__ b, int cb4, float4* __restrict__ out, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4 = ca4 + cb4, q = (int)(idx % c4);
    const long long p = idx / c4;
    out[idx] = q < ca4 ? a[p * ca4 + q] : b[p * cb4 + (q - ca4)];
}

// One ConvLSTM step for all windows: gates = Gx[frame of (window, step)] (+ conv(h_prev) when given); Conv2DLSTMCell gate order
// (i, j, f, o), forget_bias = 1: c' = sigmoid(f + 1) c + sigmoid(i) tanh(j), h' = tanh(c') sigmoid(o).  One thread = 4 hidden channels.
__global__ void lstm_point_kernel(const float* __restrict__ gh, const float* __restrict__ gx, float* __restrict__ c, float* __restrict__ h_out,
                                  int n_win, long long hw, int nh, int frame_off, int n_frames, int first) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q4 = nh / 4;
    const long long total = (long long)n_win * hw * q4;
    if (idx >= total) return;
    const int q = (int)(idx % q4);
    const long long p = (idx / q4) % hw;
    const int j = (int)(idx / (q4 * hw));
    int f = (j + frame_off) % n_frames;
    if (f < 0) f += n_frames;
    const float* gxp = gx + ((size_t)f * hw + p) * 4 * nh + 4 * q;
    const float* ghp = gh + ((size_t)j * hw + p) * 4 * nh + 4 * q;
    float4 g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        g[k] = *reinterpret_cast<const float4*>(gxp + k * nh);
        if (!first) {
            const float4 t = *reinterpret_cast<const float4*>(ghp + k * nh);
            g[k].x += t.x; g[k].y += t.y; g[k].z += t.z; g[k].w += t.w;
        }
    }
    float* cp = c + ((size_t)j * hw + p) * nh + 4 * q;
    float4 cv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(cp);
    auto sg = [](float v) { return 1.f / (1.f + expf(-v)); };
    float4 hv;
    cv.x = sg(g[2].x + 1.f) * cv.x + sg(g[0].x) * tanhf(g[1].x); hv.x = tanhf(cv.x) * sg(g[3].x);
    cv.y = sg(g[2].y + 1.f) * cv.y + sg(g[0].y) * tanhf(g[1].y); hv.y = tanhf(cv.y) * sg(g[3].y);
    cv.z = sg(g[2].z + 1.f) * cv.z + sg(g[0].z) * tanhf(g[1].z); hv.z = tanhf(cv.z) * sg(g[3].z);
    cv.w = sg(g[2].w + 1.f) * cv.w + sg(g[0].w) * tanhf(g[1].w); hv.w = tanhf(cv.w) * sg(g[3].w);
    *reinterpret_cast<float4*>(cp) = cv;
    *reinterpret_cast<float4*>(h_out + ((size_t)j * hw + p) * nh + 4 * q) = hv;
}

// Per frame f and pixel: the 2R-1 windows that contain f (window j holds frame f at position s = f - j + rad), in ascending window
// order like the reference loop (deploy_network_ao.py:146-177): logits = W_out [h_fw(s, j) | h_bw(s, j)] + b, softmax (float32),
// prob[f] += p * w_s with the reference's arithmetic (float32 accumulator, float64 product and sum), prob /= sum of weights,
// argmax (first maximum), crop.
template <int NC>
__global__ void ao_output_kernel(const float* __restrict__ hf, const float* __restrict__ hb, const float* __restrict__ wout,
                                 const float* __restrict__ bout, const double* __restrict__ wwin, int tw, int n_frames, int h2, int w2, int nh,
                                 int x_pre, int y_pre, int x, int y, uint8_t* __restrict__ labels, float* __restrict__ prob) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long hw = (long long)h2 * w2;
    if (idx >= (long long)n_frames * hw) return;
    const int f = (int)(idx / hw);
    const long long p = idx % hw;
    const int py = (int)(p / w2), px = (int)(p % w2);
    const int yy = py - y_pre, xx = px - x_pre;
    if (yy < 0 || yy >= y || xx < 0 || xx >= x) return;
    const int rad = (tw - 1) / 2;
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.f;
    double wsum = 0.0;
    // windows containing f: j = f - rad .. f + rad (mod T); ascending j means: first the wrapped-around small indices
    for (int pass = 0; pass < 3; ++pass)
        for (int d = -rad; d <= rad; ++d) {
            const int jr = f + d;                                   // un-wrapped window index
            const int seg = jr >= n_frames ? 0 : (jr >= 0 ? 1 : 2); // wrapped to [0, rad) / in range / wrapped to the top
            if (seg != pass) continue;
            const int j = jr >= n_frames ? jr - n_frames : (jr < 0 ? jr + n_frames : jr);
            const int s = rad - d;                                  // position of frame f inside window j
            const float* a = hf + (((size_t)s * n_frames + j) * hw + p) * nh;
            const float* b = hb + (((size_t)s * n_frames + j) * hw + p) * nh;
            float lg[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) lg[c] = 0.f;
            for (int k = 0; k < nh; ++k) {
                const float va = a[k], vb = b[k];
#pragma unroll
                for (int c = 0; c < NC; ++c) lg[c] = fmaf(va, wout[k * NC + c], fmaf(vb, wout[(nh + k) * NC + c], lg[c]));
            }
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < NC; ++c) { lg[c] += bout[c]; m = fmaxf(m, lg[c]); }
            float e[NC], ssum = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) { e[c] = expf(lg[c] - m); ssum += e[c]; }
            const double wv = wwin[s];
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] = (float)((double)acc[c] + (double)(e[c] / ssum) * wv);
            wsum += wv;
        }
    int arg = 0;
    float best = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const float pr = (float)((double)acc[c] / wsum);
        if (pr > best) { best = pr; arg = c; }
        if (prob) prob[(((size_t)f * y + yy) * x + xx) * NC + c] = pr;
    }
    labels[((size_t)f * y + yy) * x + xx] = (uint8_t)arg;
}

// ---------------------------------------------------------------------------------------------------- engine
struct AoEngine {
    int device = 0, n_class = 3, nh = 16, n_level = 5;
    int nf[8] = {};
    ConvLayer down[16], upt[8], up[16], lx[2], lh[2];
    float* w_out = nullptr;     // [2 nh][n_class]
    float* b_out = nullptr;
    double* d_wwin = nullptr;   // window weights (up to 64)
    // workspace
    int cap_t = 0, cap_h = 0, cap_w = 0;
    float* lvl[8] = {};         // net['conv_l']
    float* tmp0 = nullptr;      // ping
    float* tmp1 = nullptr;      // pong / concat
    float* gx[2] = {};          // conv_x(features) + bias per direction [T][hw][4 nh]
    float* gh = nullptr;        // conv_h(h_prev) [T windows][hw][4 nh]
    float* cst = nullptr;       // cell state [T][hw][nh]
    float* hs[2] = {};          // hidden outputs per direction [tw][T][hw][nh]
    int cap_tw = 0;
    long long launches = 0;
};

static int ao_upload_conv(ConvLayer& L, const float* kernel_hwio, int ks, int cin, int cin_off, int cin_total, int cout, int stride,
                          const std::vector<float>& sc, const std::vector<float>& sh, int relu, bool transposed_layout) {
    L.ksize = ks; L.cin = cin; L.cout = cout; L.stride = stride; L.relu = relu;
    std::vector<float> wt((size_t)ks * ks * cin * cout);
    // device tap (dy, dx) <- TF kernel[kh = dx][kw = dy]  (device rows are Y = TF's W axis).  conv kernels are [kh][kw][cin][cout];
    // conv2d_transpose kernels are [kh][kw][cout][cin].
    for (int dy = 0; dy < ks; ++dy)
        for (int dx = 0; dx < ks; ++dx)
            for (int ci = 0; ci < cin; ++ci)
                for (int co = 0; co < cout; ++co) {
                    const size_t tf_tap = (size_t)(dx * ks + dy);
                    const float v = transposed_layout ? kernel_hwio[(tf_tap * cout + co) * cin_total + cin_off + ci]
                                                      : kernel_hwio[(tf_tap * cin_total + cin_off + ci) * cout + co];
                    wt[((size_t)(dy * ks + dx) * cin + ci) * cout + co] = v;
                }
    UKBB_CUDA(cudaMalloc(&L.w_f32, wt.size() * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&L.scale, cout * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&L.shift, cout * sizeof(float)));
    UKBB_CUDA(cudaMemcpy(L.w_f32, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    UKBB_CUDA(cudaMemcpy(L.scale, sc.data(), cout * sizeof(float), cudaMemcpyHostToDevice));
    UKBB_CUDA(cudaMemcpy(L.shift, sh.data(), cout * sizeof(float), cudaMemcpyHostToDevice));
    return UKBB_OK;
}

static void bn_fold(const ukbb_conv_weights& c, float eps, std::vector<float>& sc, std::vector<float>& sh) {
    sc.resize(c.cout); sh.resize(c.cout);
    for (int k = 0; k < c.cout; ++k) {
        const double s = (double)c.gamma[k] / sqrt((double)c.moving_variance[k] + (double)eps);
        sc[k] = (float)s;
        sh[k] = (float)((double)c.beta[k] - (double)c.moving_mean[k] * s);
    }
}

static void ao_free_layer(ConvLayer& L) { cudaFree(L.w_f32); cudaFree(L.scale); cudaFree(L.shift); L.w_f32 = L.scale = L.shift = nullptr; }

static void ao_free_ws(AoEngine* h) {
    for (int l = 0; l < 8; ++l) { cudaFree(h->lvl[l]); h->lvl[l] = nullptr; }
    cudaFree(h->tmp0); cudaFree(h->tmp1); cudaFree(h->gh); cudaFree(h->cst);
    h->tmp0 = h->tmp1 = h->gh = h->cst = nullptr;
    for (int d = 0; d < 2; ++d) { cudaFree(h->gx[d]); cudaFree(h->hs[d]); h->gx[d] = h->hs[d] = nullptr; }
    h->cap_t = h->cap_h = h->cap_w = h->cap_tw = 0;
}

static int ao_ensure_ws(AoEngine* h, int t, int hh, int ww, int tw) {
    if (h->cap_t >= t && h->cap_h == hh && h->cap_w == ww && h->cap_tw >= tw) return UKBB_OK;
    UKBB_CUDA(cudaDeviceSynchronize());
    ao_free_ws(h);
    const size_t hw = (size_t)hh * ww;
    for (int l = 0; l < h->n_level; ++l) UKBB_CUDA(cudaMalloc(&h->lvl[l], (size_t)t * (hw >> (2 * l)) * h->nf[l] * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&h->tmp0, (size_t)t * hw * h->nf[0] * 2 * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&h->tmp1, (size_t)t * hw * h->nf[0] * 2 * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&h->gh, (size_t)t * hw * 4 * h->nh * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&h->cst, (size_t)t * hw * h->nh * sizeof(float)));
    for (int d = 0; d < 2; ++d) {
        UKBB_CUDA(cudaMalloc(&h->gx[d], (size_t)t * hw * 4 * h->nh * sizeof(float)));
        UKBB_CUDA(cudaMalloc(&h->hs[d], (size_t)tw * t * hw * h->nh * sizeof(float)));
    }
    h->cap_t = t; h->cap_h = hh; h->cap_w = ww; h->cap_tw = tw;
    return UKBB_OK;
}

static void same_pad_ao(int in, int k, int s, int* out, int* before) {
    *out = (in + s - 1) / s;
    int total = (*out - 1) * s + k - in;
    if (total < 0) total = 0;
    *before = total / 2;
}

static int ao_conv(AoEngine* h, const float* in, float* out, const ConvLayer& L, int n, int hi, int wi, cudaStream_t st) {
    int ho, wo, pt, pl;
    same_pad_ao(hi, L.ksize, L.stride, &ho, &pt);
    same_pad_ao(wi, L.ksize, L.stride, &wo, &pl);
    h->launches++;
    return launch_conv_fp32(in, out, L, n, hi, wi, ho, wo, pt, pl, st);
}

}  // namespace ukbb

using namespace ukbb;

extern "C" {

int ukbb_ao_create(const ukbb_ao_weights* w, int device, ukbb_ao** out) {
    UKBB_REQUIRE(w && out, "ao_create: null argument");
    *out = nullptr;
    UKBB_REQUIRE(w->n_level >= 2 && w->n_level <= 6, "ao_create: n_level=%d not in [2, 6]", w->n_level);
    UKBB_REQUIRE(w->n_class >= 2 && w->n_class <= 4, "ao_create: n_class=%d not in [2, 4]", w->n_class);
    UKBB_REQUIRE(w->n_hidden >= 8 && w->n_hidden % 8 == 0 && w->n_hidden <= 64, "ao_create: n_hidden=%d must be a multiple of 8 in [8, 64]", w->n_hidden);
    UKBB_REQUIRE(w->down && w->up_transpose && w->up && w->lstm_kernel[0] && w->lstm_kernel[1] && w->lstm_bias[0] && w->lstm_bias[1] && w->out_kernel &&
                 w->out_bias, "ao_create: missing weights");
    const int nl = w->n_level;
    const int f0 = w->down[0].cout;
    UKBB_REQUIRE(f0 >= 16 && f0 % 16 == 0, "ao_create: first-level filter count %d must be a multiple of 16", f0);
    // topology of network_ao.py:18-64 with n_block = 2 per level (train_network_ao.py:284)
    for (int l = 0, cin = 1; l < nl; ++l)
        for (int b = 0; b < 2; ++b) {
            const ukbb_conv_weights& c = w->down[2 * l + b];
            const int cout = f0 << l, stride = (l > 0 && b == 0) ? 2 : 1;
            UKBB_REQUIRE(c.kernel && c.gamma && c.beta && c.moving_mean && c.moving_variance && c.ksize == 3 && c.cin == cin && c.cout == cout && c.stride == stride,
                         "ao_create: encoder conv %d of level %d is %dx%d %d->%d stride %d, UNet expects 3x3 %d->%d stride %d", b, l, c.ksize, c.ksize,
                         c.cin, c.cout, c.stride, cin, cout, stride);
            cin = cout;
        }
    for (int i = 0; i < nl - 1; ++i) {
        const int l = nl - 2 - i, cout = f0 << l;
        const ukbb_conv_weights& t = w->up_transpose[i];
        UKBB_REQUIRE(t.kernel && t.gamma && t.ksize == 3 && t.cin == 2 * cout && t.cout == cout && t.stride == 2,
                     "ao_create: transposed conv of level %d is %dx%d %d->%d stride %d, UNet expects 3x3 %d->%d stride 2", l, t.ksize, t.ksize, t.cin,
                     t.cout, t.stride, 2 * cout, cout);
        for (int b = 0; b < 2; ++b) {
            const ukbb_conv_weights& c = w->up[2 * i + b];
            const int cin = b == 0 ? 2 * cout : cout;
            UKBB_REQUIRE(c.kernel && c.gamma && c.ksize == 3 && c.cin == cin && c.cout == cout && c.stride == 1,
                         "ao_create: decoder conv %d of level %d is %dx%d %d->%d, UNet expects 3x3 %d->%d", b, l, c.ksize, c.ksize, c.cin, c.cout, cin, cout);
        }
    }
    int ndev = 0;
    UKBB_CUDA(cudaGetDeviceCount(&ndev));
    UKBB_REQUIRE(device >= 0 && device < ndev, "ao_create: device %d out of range (%d CUDA devices)", device, ndev);
    UKBB_CUDA(cudaSetDevice(device));
    AoEngine* h = new (std::nothrow) AoEngine();
    if (!h) { set_error("ao_create: out of host memory"); return UKBB_E_NOMEM; }
    h->device = device; h->n_class = w->n_class; h->nh = w->n_hidden; h->n_level = nl;
    for (int l = 0; l < nl; ++l) h->nf[l] = f0 << l;
    int rc = UKBB_OK;
    std::vector<float> sc, sh;
    for (int i = 0; i < 2 * nl && !rc; ++i) {
        const ukbb_conv_weights& c = w->down[i];
        bn_fold(c, w->bn_eps, sc, sh);
        rc = ao_upload_conv(h->down[i], c.kernel, 3, c.cin, 0, c.cin, c.cout, c.stride, sc, sh, 1, false);
    }
    for (int i = 0; i < nl - 1 && !rc; ++i) {
        const ukbb_conv_weights& t = w->up_transpose[i];
        bn_fold(t, w->bn_eps, sc, sh);
        rc = ao_upload_conv(h->upt[i], t.kernel, 3, t.cin, 0, t.cin, t.cout, 2, sc, sh, 1, true);
        for (int b = 0; b < 2 && !rc; ++b) {
            const ukbb_conv_weights& c = w->up[2 * i + b];
            bn_fold(c, w->bn_eps, sc, sh);
            rc = ao_upload_conv(h->up[2 * i + b], c.kernel, 3, c.cin, 0, c.cin, c.cout, 1, sc, sh, 1, false);
        }
    }
    const int nh = h->nh;
    for (int d = 0; d < 2 && !rc; ++d) {
        // Conv2DLSTMCell kernel [3][3][f0 + nh][4 nh]: the x rows with the bias, the h rows without
        std::vector<float> one(4 * nh, 1.f), bias(w->lstm_bias[d], w->lstm_bias[d] + 4 * nh), zero(4 * nh, 0.f);
        rc = ao_upload_conv(h->lx[d], w->lstm_kernel[d], 3, f0, 0, f0 + nh, 4 * nh, 1, one, bias, 0, false);
        if (!rc) rc = ao_upload_conv(h->lh[d], w->lstm_kernel[d], 3, nh, f0, f0 + nh, 4 * nh, 1, one, zero, 0, false);
    }
    if (!rc) {
        cudaError_t e = cudaMalloc(&h->w_out, (size_t)2 * nh * h->n_class * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&h->b_out, h->n_class * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&h->d_wwin, 64 * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(h->w_out, w->out_kernel, (size_t)2 * nh * h->n_class * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->b_out, w->out_bias, h->n_class * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error("ao_create: %s", cudaGetErrorString(e)); rc = UKBB_E_CUDA; }
    }
    if (rc) { ukbb_ao_destroy(reinterpret_cast<ukbb_ao*>(h)); return rc; }
    *out = reinterpret_cast<ukbb_ao*>(h);
    return UKBB_OK;
}

void ukbb_ao_destroy(ukbb_ao* hh) {
    AoEngine* h = reinterpret_cast<AoEngine*>(hh);
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 16; ++i) { ao_free_layer(h->down[i]); ao_free_layer(h->up[i]); }
    for (int i = 0; i < 8; ++i) ao_free_layer(h->upt[i]);
    for (int d = 0; d < 2; ++d) { ao_free_layer(h->lx[d]); ao_free_layer(h->lh[d]); }
    cudaFree(h->w_out); cudaFree(h->b_out); cudaFree(h->d_wwin);
    ao_free_ws(h);
    delete h;
}

long long ukbb_ao_launch_count(const ukbb_ao* hh) { return hh ? reinterpret_cast<const AoEngine*>(hh)->launches : 0; }

int ukbb_ao_segment(ukbb_ao* hh, const float* image, int n_frames, int x2, int y2, int x_pre, int y_pre, int x, int y, int weight_R, double weight_r,
                    uint8_t* labels, float* prob, void* stream) {
    AoEngine* h = reinterpret_cast<AoEngine*>(hh);
    UKBB_REQUIRE(h && image && labels, "ao_segment: null argument");
    const int nl = h->n_level, down_f = 1 << (nl - 1);
    UKBB_REQUIRE(n_frames > 0 && x2 > 0 && y2 > 0 && x2 % down_f == 0 && y2 % down_f == 0, "ao_segment: padded size %dx%d must be a positive multiple of %d",
                 x2, y2, down_f);
    UKBB_REQUIRE(x > 0 && y > 0 && x_pre >= 0 && y_pre >= 0 && x_pre + x <= x2 && y_pre + y <= y2, "ao_segment: crop (%d,%d)+(%d,%d) outside padded %dx%d",
                 x_pre, y_pre, x, y, x2, y2);
    const int tw = 2 * weight_R - 1, rad = (tw - 1) / 2;
    UKBB_REQUIRE(weight_R >= 1 && tw <= 63, "ao_segment: weight_R=%d out of range", weight_R);
    UKBB_REQUIRE(n_frames >= tw, "ao_segment: %d frames are fewer than the time window of %d (the reference's fancy-indexed += would drop duplicates)",
                 n_frames, tw);
    UKBB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int T = n_frames, hh2 = y2, ww2 = x2;          // device rows are Y, columns X
    int rc = ao_ensure_ws(h, T, hh2, ww2, tw);
    if (rc) return rc;
    {   // window weights, deploy_network_ao.py:134-143 (float64)
        double wv[64];
        for (int t = 0; t < tw; ++t) {
            const int d = abs(t - rad);
            wv[t] = d <= weight_R ? pow(1.0 - (double)d / (double)weight_R, weight_r) : 0.0;
        }
        UKBB_CUDA(cudaMemcpyAsync(h->d_wwin, wv, tw * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    const size_t hw = (size_t)hh2 * ww2;
    // ---- UNet, all T frames as one batch (network_ao.py:18-56)
    const float* cur = image;
    int ch = hh2, cw = ww2;
    for (int l = 0; l < nl; ++l) {
        if (l > 0) { ch /= 2; cw /= 2; }
        rc = ao_conv(h, cur, h->tmp0, h->down[2 * l], T, l > 0 ? ch * 2 : ch, l > 0 ? cw * 2 : cw, st);
        if (!rc) rc = ao_conv(h, h->tmp0, h->lvl[l], h->down[2 * l + 1], T, ch, cw, st);
        if (rc) return rc;
        cur = h->lvl[l];
    }
    const float* upv = h->lvl[nl - 1];
    for (int i = 0; i < nl - 1; ++i) {
        const int l = nl - 2 - i, f = h->nf[l];
        const int hi = hh2 >> (l + 1), wi = ww2 >> (l + 1), ho = 2 * hi, wo = 2 * wi;
        {
            dim3 grid((unsigned)(((long long)ho * wo + 127) / 128), f / 16, T);
            convT_fp32_kernel<<<grid, 128, 0, st>>>(upv, h->tmp0, h->upt[i].w_f32, h->upt[i].scale, h->upt[i].shift, 2 * f, f, hi, wi);
            UKBB_CUDA(cudaGetLastError());
            h->launches++;
        }
        {
            const long long total = (long long)T * ho * wo * (2 * f / 4);
            concat2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float4*)h->lvl[l], f / 4, (const float4*)h->tmp0, f / 4, (float4*)h->tmp1, total);
            UKBB_CUDA(cudaGetLastError());
            h->launches++;
        }
        rc = ao_conv(h, h->tmp1, h->tmp0, h->up[2 * i], T, ho, wo, st);
        if (!rc) rc = ao_conv(h, h->tmp0, h->lvl[l], h->up[2 * i + 1], T, ho, wo, st);      // net['conv_l_up'] overwrites net['conv_l'] (no longer needed)
        if (rc) return rc;
        upv = h->lvl[l];
    }
    const float* feat = h->lvl[0];                                                           // [T][hw][f0]
    // ---- ConvLSTM: conv_x(features) + bias once per frame and direction; the recurrence over the window positions for all T windows
    const int nh = h->nh;
    for (int d = 0; d < 2; ++d) {
        rc = ao_conv(h, feat, h->gx[d], h->lx[d], T, hh2, ww2, st);
        if (rc) return rc;
        for (int k = 0; k < tw; ++k) {
            const int s = d == 0 ? k : tw - 1 - k;                                           // window position handled at step k
            float* h_out = h->hs[d] + (size_t)s * T * hw * nh;
            if (k > 0) {
                const int sp = d == 0 ? s - 1 : s + 1;
                rc = ao_conv(h, h->hs[d] + (size_t)sp * T * hw * nh, h->gh, h->lh[d], T, hh2, ww2, st);
                if (rc) return rc;
            }
            const long long total = (long long)T * hw * (nh / 4);
            lstm_point_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(h->gh, h->gx[d], h->cst, h_out, T, (long long)hw, nh, s - rad, T, k == 0 ? 1 : 0);
            UKBB_CUDA(cudaGetLastError());
            h->launches++;
        }
    }
    {
        const long long total = (long long)T * hw;
        const unsigned grid = (unsigned)((total + 255) / 256);
#define AO_OUT(NC) ao_output_kernel<NC><<<grid, 256, 0, st>>>(h->hs[0], h->hs[1], h->w_out, h->b_out, h->d_wwin, tw, T, hh2, ww2, nh, x_pre, y_pre, x, y, labels, prob)
        switch (h->n_class) {
            case 2: AO_OUT(2); break;
            case 3: AO_OUT(3); break;
            default: AO_OUT(4); break;
        }
#undef AO_OUT
        UKBB_CUDA(cudaGetLastError());
        h->launches++;
    }
    return UKBB_OK;
}

}  // extern "C"
