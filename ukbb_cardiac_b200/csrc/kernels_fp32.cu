// FP32 CUDA-core kernels: the exactness mode of the FCN forward (north_star (a) "FP32
// CUDA-core mode for exactness checks").  Restates, per layer,
//   common/network.py:19-25   conv (TF SAME, no bias) -> BN(inference) -> ReLU
//   common/network.py:138-167 fixed bilinear transposed-conv upsampling
//   common/network.py:214-218 channel concat
//   common/network.py:229 + common/train_network.py:198-199  logits, softmax, argmax
// on NHWC activations whose rows are Y and columns are X (NIfTI slice order).
#include "common.cuh"

namespace ukbb {

// ------------------------------------------------------------------------------------------
// Direct convolution, register-tiled: one thread = a 2 x 2 block of output pixels x CB output channels.
// Block = 16 x 32 output pixels (128 threads); the input patch of CC channels is staged in shared memory
// channel-major; weights of the chunk are staged [tap][c][CB] and read as broadcast float4.  Per input
// channel a thread reads its (S + KS) x (S + KS) input window once (16 or 25 LDS) and the KS x KS x 4 weight
// vectors once for 4 x KS x KS x CB FMAs: ~13 FMAs per shared-memory load (the first version of this kernel,
// one pixel per thread, issued 5 loads per 16 FMAs and ran at 12 TFLOP/s).  Every output accumulates its
// products in the same order as before (chunk, ky, kx, c), so results are bit-identical to that version.
// ------------------------------------------------------------------------------------------
constexpr int TH = 8, TW = 16, CB = 16;          // threads per block: TH x TW, each owning 2 x 2 pixels
constexpr int OH = 2 * TH, OW = 2 * TW;          // output pixels per block: 16 x 32

template <int KS, int S, int CC>
__global__ void __launch_bounds__(TH* TW)
conv_fp32_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wt,
                 const float* __restrict__ scale, const float* __restrict__ shift, int cin, int cout,
                 int hi, int wi, int ho, int wo, int pad_top, int pad_left, int tiles_x, int relu) {
    constexpr int PH = (OH - 1) * S + KS, PW = (OW - 1) * S + KS;
    constexpr int WIN = S + KS;                  // input window of a 2 x 2 output block (per axis)
    __shared__ float s_in[CC][PH * PW];
    __shared__ __align__(16) float s_w[KS * KS][CC][CB];

    const int tid = threadIdx.x;
    const int px = tid % TW, py = tid / TW;
    const int tile = blockIdx.x;
    const int ox0 = (tile % tiles_x) * OW, oy0 = (tile / tiles_x) * OH;
    const int cb0 = blockIdx.y * CB;
    const int n = blockIdx.z;
    const int iy0 = oy0 * S - pad_top, ix0 = ox0 * S - pad_left;

    float acc[4][CB];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < CB; ++j) acc[q][j] = 0.f;

    const float* in_n = in + (size_t)n * hi * wi * cin;
    for (int c0 = 0; c0 < cin; c0 += CC) {
        for (int e = tid; e < PH * PW * CC; e += TH * TW) {
            const int c = e % CC, pix = e / CC;
            const int gy = iy0 + pix / PW, gx = ix0 + pix % PW;
            float v = 0.f;
            if (gy >= 0 && gy < hi && gx >= 0 && gx < wi) v = in_n[((size_t)gy * wi + gx) * cin + c0 + c];
            s_in[c][pix] = v;
        }
        for (int e = tid; e < KS * KS * CC * CB; e += TH * TW) {
            const int j = e % CB, c = (e / CB) % CC, tap = e / (CB * CC);
            s_w[tap][c][j] = wt[((size_t)tap * cin + c0 + c) * cout + cb0 + j];
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < CC; ++c) {
            float v[WIN][WIN];
            const float* base = &s_in[c][(2 * py * S) * PW + 2 * px * S];
#pragma unroll
            for (int r = 0; r < WIN; ++r)
#pragma unroll
                for (int k = 0; k < WIN; ++k) v[r][k] = base[r * PW + k];
#pragma unroll
            for (int ky = 0; ky < KS; ++ky)
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const float4* w4 = reinterpret_cast<const float4*>(&s_w[ky * KS + kx][c][0]);
#pragma unroll
                    for (int q4 = 0; q4 < CB / 4; ++q4) {
                        const float4 w = w4[q4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float x = v[(q >> 1) * S + ky][(q & 1) * S + kx];
                            acc[q][4 * q4 + 0] = fmaf(x, w.x, acc[q][4 * q4 + 0]);
                            acc[q][4 * q4 + 1] = fmaf(x, w.y, acc[q][4 * q4 + 1]);
                            acc[q][4 * q4 + 2] = fmaf(x, w.z, acc[q][4 * q4 + 2]);
                            acc[q][4 * q4 + 3] = fmaf(x, w.w, acc[q][4 * q4 + 3]);
                        }
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int oy = oy0 + 2 * py + (q >> 1), ox = ox0 + 2 * px + (q & 1);
        if (oy < ho && ox < wo) {
            float* o = out + (((size_t)n * ho + oy) * wo + ox) * cout + cb0;
#pragma unroll
            for (int q4 = 0; q4 < CB / 4; ++q4) {
                float4 r;
                r.x = fmaf(acc[q][4 * q4 + 0], scale[cb0 + 4 * q4 + 0], shift[cb0 + 4 * q4 + 0]);
                r.y = fmaf(acc[q][4 * q4 + 1], scale[cb0 + 4 * q4 + 1], shift[cb0 + 4 * q4 + 1]);
                r.z = fmaf(acc[q][4 * q4 + 2], scale[cb0 + 4 * q4 + 2], shift[cb0 + 4 * q4 + 2]);
                r.w = fmaf(acc[q][4 * q4 + 3], scale[cb0 + 4 * q4 + 3], shift[cb0 + 4 * q4 + 3]);
                if (relu) {
                    r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
                }
                reinterpret_cast<float4*>(o)[q4] = r;
            }
        }
    }
}

// The first version of the kernel: one thread = ONE output pixel x CB channels, block = 8 x 16 pixels.  Kept for small feature maps
// (levels 3 / 4 of a short-axis slice are 24 x 26 and 12 x 13 pixels: a 16 x 32 tile would be mostly padding there).
template <int KS, int S, int CC>
__global__ void __launch_bounds__(TH* TW)
conv_fp32_px_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wt,
                 const float* __restrict__ scale, const float* __restrict__ shift, int cin, int cout,
                 int hi, int wi, int ho, int wo, int pad_top, int pad_left, int tiles_x, int relu) {
    constexpr int PH = (TH - 1) * S + KS, PW = (TW - 1) * S + KS;
    __shared__ float s_in[CC][PH * PW];
    __shared__ __align__(16) float s_w[KS * KS][CC][CB];

    const int tid = threadIdx.x;
    const int px = tid % TW, py = tid / TW;
    const int tile = blockIdx.x;
    const int ox0 = (tile % tiles_x) * TW, oy0 = (tile / tiles_x) * TH;
    const int cb0 = blockIdx.y * CB;
    const int n = blockIdx.z;
    const int iy0 = oy0 * S - pad_top, ix0 = ox0 * S - pad_left;

    float acc[CB];
#pragma unroll
    for (int j = 0; j < CB; ++j) acc[j] = 0.f;

    const float* in_n = in + (size_t)n * hi * wi * cin;
    for (int c0 = 0; c0 < cin; c0 += CC) {
        for (int e = tid; e < PH * PW * CC; e += TH * TW) {
            const int c = e % CC, pix = e / CC;
            const int gy = iy0 + pix / PW, gx = ix0 + pix % PW;
            float v = 0.f;
            if (gy >= 0 && gy < hi && gx >= 0 && gx < wi) v = in_n[((size_t)gy * wi + gx) * cin + c0 + c];
            s_in[c][pix] = v;
        }
        for (int e = tid; e < KS * KS * CC * CB; e += TH * TW) {
            const int j = e % CB, c = (e / CB) % CC, tap = e / (CB * CC);
            s_w[tap][c][j] = wt[((size_t)tap * cin + c0 + c) * cout + cb0 + j];
        }
        __syncthreads();
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    const float v = s_in[c][(py * S + ky) * PW + px * S + kx];
                    const float4* w4 = reinterpret_cast<const float4*>(&s_w[ky * KS + kx][c][0]);
#pragma unroll
                    for (int q = 0; q < CB / 4; ++q) {
                        const float4 w = w4[q];
                        acc[4 * q + 0] = fmaf(v, w.x, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(v, w.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(v, w.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(v, w.w, acc[4 * q + 3]);
                    }
                }
        __syncthreads();
    }
    const int oy = oy0 + py, ox = ox0 + px;
    if (oy < ho && ox < wo) {
        float* o = out + (((size_t)n * ho + oy) * wo + ox) * cout + cb0;
#pragma unroll
        for (int q = 0; q < CB / 4; ++q) {
            float4 r;
            r.x = fmaf(acc[4 * q + 0], scale[cb0 + 4 * q + 0], shift[cb0 + 4 * q + 0]);
            r.y = fmaf(acc[4 * q + 1], scale[cb0 + 4 * q + 1], shift[cb0 + 4 * q + 1]);
            r.z = fmaf(acc[4 * q + 2], scale[cb0 + 4 * q + 2], shift[cb0 + 4 * q + 2]);
            r.w = fmaf(acc[4 * q + 3], scale[cb0 + 4 * q + 3], shift[cb0 + 4 * q + 3]);
            if (relu) {
                r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
            }
            reinterpret_cast<float4*>(o)[q] = r;
        }
    }
}

int launch_conv_fp32(const float* in, float* out, const ConvLayer& L, int n, int hi, int wi, int ho,
                     int wo, int pad_top, int pad_left, cudaStream_t st) {
    UKBB_REQUIRE(L.cout % CB == 0, "conv_fp32: cout=%d is not a multiple of %d", L.cout, CB);
    UKBB_REQUIRE(L.cin == 1 || L.cin % 8 == 0, "conv_fp32: cin=%d must be 1 or a multiple of 8", L.cin);
    const bool big = wo >= 48 && ho >= 24;            // 2 x 2 pixels per thread, 16 x 32 tiles
    const int tw = big ? OW : TW, th = big ? OH : TH;
    const int tiles_x = (wo + tw - 1) / tw, tiles_y = (ho + th - 1) / th;
    dim3 grid(tiles_x * tiles_y, L.cout / CB, n), block(TH * TW);
#define LAUNCH(KERN, KS, S, CC)                                                                   \
    KERN<KS, S, CC><<<grid, block, 0, st>>>(in, out, L.w_f32, L.scale, L.shift, L.cin, L.cout, hi, wi, ho, wo, pad_top, pad_left, tiles_x, L.relu)
    if (L.ksize == 3 && L.stride == 1 && L.cin == 1) { if (big) LAUNCH(conv_fp32_kernel, 3, 1, 1); else LAUNCH(conv_fp32_px_kernel, 3, 1, 1); }
    else if (L.ksize == 3 && L.stride == 1) { if (big) LAUNCH(conv_fp32_kernel, 3, 1, 8); else LAUNCH(conv_fp32_px_kernel, 3, 1, 8); }
    else if (L.ksize == 3 && L.stride == 2) { if (big) LAUNCH(conv_fp32_kernel, 3, 2, 4); else LAUNCH(conv_fp32_px_kernel, 3, 2, 8); }
    else if (L.ksize == 1 && L.stride == 1) { if (big) LAUNCH(conv_fp32_kernel, 1, 1, 8); else LAUNCH(conv_fp32_px_kernel, 1, 1, 8); }
    else {
        set_error("conv_fp32: unsupported ksize=%d stride=%d cin=%d", L.ksize, L.stride, L.cin);
        return UKBB_E_UNSUPPORTED;
    }
#undef LAUNCH
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}

// ------------------------------------------------------------------------------------------
// Bilinear transposed-conv upsampling (network.py:138-167) + concat (network.py:214-218).
// For factor f = 2^l the transposed conv with the (2f-1)-tap hat filter and TF 'SAME'
// cropping (pad_before = (f-1)//2) reduces, per axis, to
//   r = (y + pb) mod f, i1 = (y + pb) div f, i0 = i1 - 1, w1 = (r+1)/f, w0 = 1 - w1,
// taps that fall outside the low-resolution map contribute 0 (tapered borders).
// One thread = one output pixel x 4 channels of one level.
// ------------------------------------------------------------------------------------------
__global__ void upsample_concat_fp32_kernel(const float* __restrict__ s0, const float* __restrict__ s1,
                                            const float* __restrict__ s2, const float* __restrict__ s3,
                                            const float* __restrict__ s4, float* __restrict__ out,
                                            long long total, int h, int w) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int q = (int)(idx % 40);             // float4 index inside the 160 channels
    const long long pix = idx / 40;
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const long long n = pix / ((long long)w * h);
    const int l = q / 8, c4 = q % 8;
    float4 r;
    if (l == 0) {
        r = reinterpret_cast<const float4*>(s0 + ((n * h + y) * w + x) * 32)[c4];
    } else {
        const float* src = l == 1 ? s1 : l == 2 ? s2 : l == 3 ? s3 : s4;
        const int f = 1 << l, pb = (f - 1) / 2;
        const int hl = h >> l, wl = w >> l;
        const int ry = (y + pb) % f, y1 = (y + pb) / f, y0 = y1 - 1;
        const int rx = (x + pb) % f, x1 = (x + pb) / f, x0 = x1 - 1;
        const float wy1 = (float)(ry + 1) / (float)f, wy0 = 1.f - wy1;
        const float wx1 = (float)(rx + 1) / (float)f, wx0 = 1.f - wx1;
        r = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = src + n * hl * wl * 32;
        auto tap = [&](int yy, int xx, float wgt) {
            if (yy < 0 || yy >= hl || xx < 0 || xx >= wl || wgt == 0.f) return;
            const float4 v = reinterpret_cast<const float4*>(base + ((long long)yy * wl + xx) * 32)[c4];
            r.x = fmaf(v.x, wgt, r.x); r.y = fmaf(v.y, wgt, r.y);
            r.z = fmaf(v.z, wgt, r.z); r.w = fmaf(v.w, wgt, r.w);
        };
        tap(y0, x0, wy0 * wx0);
        tap(y0, x1, wy0 * wx1);
        tap(y1, x0, wy1 * wx0);
        tap(y1, x1, wy1 * wx1);
    }
    reinterpret_cast<float4*>(out + pix * 160)[q] = r;
}

int launch_upsample_concat_fp32(const float* const src[5], float* out, int n, int h, int w,
                                cudaStream_t st) {
    const long long total = (long long)n * h * w * 40;
    const int block = 256;
    const long long grid = (total + block - 1) / block;
    upsample_concat_fp32_kernel<<<(unsigned)grid, block, 0, st>>>(src[0], src[1], src[2], src[3], src[4],
                                                                   out, total, h, w);
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}

// ------------------------------------------------------------------------------------------
// Classifier: 1x1 conv 64 -> n_class with bias (network.py:229), softmax and argmax
// (train_network.py:198-199: argmax over the FP32 softmax, first maximal index), crop to
// the un-padded image (deploy_network.py:114-116), per-slice class counts.
// One thread = one padded pixel; one block row = one slice (blockIdx.y).
// ------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256)
classifier_fp32_kernel(const float* __restrict__ feat, const float* __restrict__ wt,
                       const float* __restrict__ bias, int h2, int w2, int x_pre, int y_pre, int x, int y,
                       uint8_t* __restrict__ labels, float* __restrict__ logits, float* __restrict__ prob,
                       unsigned long long* __restrict__ counts) {
    __shared__ float s_w[64 * NC];
    __shared__ float s_b[NC];
    for (int e = threadIdx.x; e < 64 * NC; e += blockDim.x) s_w[e] = wt[e];
    if (threadIdx.x < NC) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < h2 * w2;
    int label = -1;
    bool inside = false;
    if (live) {
        const float4* f4 = reinterpret_cast<const float4*>(feat + ((size_t)n * h2 * w2 + p) * 64);
        float lg[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) lg[c] = 0.f;
#pragma unroll 4
        for (int k4 = 0; k4 < 16; ++k4) {
            const float4 v = f4[k4];
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < NC; ++c) lg[c] = fmaf(vv[u], s_w[(k4 * 4 + u) * NC + c], lg[c]);
        }
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < NC; ++c) { lg[c] += s_b[c]; m = fmaxf(m, lg[c]); }
        float e[NC], s = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) { e[c] = expf(lg[c] - m); s += e[c]; }
        float best = -1.f;
        int arg = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float pr = e[c] / s;
            if (pr > best) { best = pr; arg = c; }
            if (prob) prob[((size_t)n * h2 * w2 + p) * NC + c] = pr;
            if (logits) logits[((size_t)n * h2 * w2 + p) * NC + c] = lg[c];
        }
        const int yy = p / w2 - y_pre, xx = p % w2 - x_pre;
        inside = yy >= 0 && yy < y && xx >= 0 && xx < x;
        if (inside) {
            labels[((size_t)n * y + yy) * x + xx] = (uint8_t)arg;
            label = arg;
        }
    }
    if (counts) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const unsigned b = __ballot_sync(0xffffffffu, inside && label == c);
            if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counts[(size_t)n * NC + c], (unsigned long long)__popc(b));
        }
    }
}

int launch_classifier_fp32(const float* feat, const ConvLayer& L, int n_class, int n, int h2, int w2,
                           int x_pre, int y_pre, int x, int y, uint8_t* labels, float* logits, float* prob,
                           unsigned long long* counts, cudaStream_t st) {
    dim3 block(256), grid((h2 * w2 + 255) / 256, n);
#define LAUNCH(NC)                                                                                   \
    classifier_fp32_kernel<NC><<<grid, block, 0, st>>>(feat, L.w_f32, L.shift, h2, w2, x_pre, y_pre, x, \
                                                       y, labels, logits, prob, counts)
    switch (n_class) {
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        case 4: LAUNCH(4); break;
        case 5: LAUNCH(5); break;
        case 6: LAUNCH(6); break;
        case 7: LAUNCH(7); break;
        case 8: LAUNCH(8); break;
        default:
            set_error("classifier: n_class=%d not in [2, %d]", n_class, UKBB_MAX_CLASS);
            return UKBB_E_UNSUPPORTED;
    }
#undef LAUNCH
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}

}  // namespace ukbb
