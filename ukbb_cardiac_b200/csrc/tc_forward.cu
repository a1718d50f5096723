// Tensor-core path of the FCN forward (north_star (a)-(c)), host side: weight re-layout, launch plans (tensor maps),
// and the launch sequence of one sub-batch:
//     conv_first_tc (conv0_0 + conv0_1) -> 11 encoder convolutions (conv_group / conv_tc / conv_halo)
//     -> side_tc (same_dim_l + fc0 column block, levels 1..4) -> head_ts (same_dim0, upsampling, fc0, fc1, class scores, labels).
// Modes: BF16 / FP16 (one 16-bit value per operand) and BF16X3 / FP16X3 (split operands: every activation and weight is a
// (hi, lo) pair of 16-bit values and every product is hi.hi + lo.hi + hi.lo with FP32 accumulation).  In the split modes a tensor
// is stored as two planes of the plain layout, the lo plane right after the hi plane; tensor maps span both planes (the lo
// plane is reached through a slice / row offset), so each kernel needs one map per tensor.
#include "tc_plan.cuh"
#include <cuda_fp8.h>
#include <math.h>
#include <string.h>
#include <vector>

namespace ukbb {

static const int kNBlockTc[5] = {2, 2, 3, 3, 3};
static const int kNFilterTc[5] = {16, 32, 64, 128, 256};

static uint16_t enc16(float v, int fp16, float* back) {
    uint16_t u;
    if (fp16) { const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); memcpy(&u, &h, 2); *back = __half2float(h); }
    else { const __nv_bfloat16 h = __float2bfloat16(v); memcpy(&u, &h, 2); *back = __bfloat162float(h); }
    return u;
}

static uint8_t enc_e4m3(float v) {
    const __nv_fp8_e4m3 q(v);                 // round to nearest even, saturating
    uint8_t u;
    memcpy(&u, &q, 1);
    return u;
}

// float values -> device array of 16-bit values: [n] (plain) or [2][n] = hi plane | lo plane (split = 1).
// split = 2 (x2 scheme, weights side): the lo plane is the FP8 correction operand -- per group of 16 consecutive K elements (32 bytes,
// the bytes an FP16 K = 16 step reads) [e4m3(w_hi * 2^4) x 16 | e4m3(w_lo * 2^15) x 16], the partner of the activations'
// [e4m3(a_lo * 2^11) x 16 | e4m3(a_hi) x 16] (tc_common.cuh).  Every weight array here has K (the innermost extent) a multiple of 16.
static int upload16(const std::vector<float>& v, int fp16, int split, __nv_bfloat16** dev) {
    const size_t n = v.size();
    std::vector<uint16_t> h((split ? 2 : 1) * n);
    if (split == 2 && n % 16 != 0) { set_error("upload16: x2 weights need K groups of 16"); return UKBB_E_INVALID; }
    uint8_t* lo8 = reinterpret_cast<uint8_t*>(h.data() + (split ? n : 0));
    for (size_t i = 0; i < n; ++i) {
        float hf, lf;
        h[i] = enc16(v[i], fp16, &hf);
        if (split == 1) h[n + i] = enc16(v[i] - hf, fp16, &lf);
        if (split == 2) {
            const size_t g = i / 16, e = i % 16;
            lo8[g * 32 + e] = enc_e4m3(ldexpf(hf, tc::X2_SW_HI));
            lo8[g * 32 + 16 + e] = enc_e4m3(ldexpf(v[i] - hf, tc::X2_SW_LO));
        }
    }
    UKBB_CUDA(cudaMalloc(dev, h.size() * 2));
    UKBB_CUDA(cudaMemcpy(*dev, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return UKBB_OK;
}

static CUtensorMapSwizzle swizzle_for(int cc) {
    return cc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : cc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

static int chunk_for(int cin) { return cin % 64 == 0 ? 64 : cin % 32 == 0 ? 32 : 16; }

#define UKBB_ENCODE(what, ...)                                                                              \
    do {                                                                                                    \
        CUresult _r = S->encode(__VA_ARGS__, CU_TENSOR_MAP_INTERLEAVE_NONE, _sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,     \
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);                                         \
        if (_r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed: %d", what, (int)_r); return UKBB_E_CUDA; }   \
    } while (0)

// Plan of conv layer li reading `in` [planes][nb][hi][wi][cin] and writing `out` [planes][nb][ho][wo][cout].
static int make_plan(Engine* h, int li, const __nv_bfloat16* in, __nv_bfloat16* out, int nb, int hi, int wi, int level_out) {
    TcState* S = h->tc;
    const ConvLayer& L = h->layers[li];
    TcLayerPlan& P = S->plan[li];
    const int split = S->split, planes = split ? 2 : 1;
    const int s = L.stride, ks = L.ksize;
    const int ho = (hi + s - 1) / s, wo = (wi + s - 1) / s;
    int pt = (ho - 1) * s + ks - hi; if (pt < 0) pt = 0; pt /= 2;
    int pl = (wo - 1) * s + ks - wi; if (pl < 0) pl = 0; pl /= 2;
    P.split = split; P.f8 = S->f8; P.cout = L.cout;
    P.kind = (ks == 3 && L.cin <= 64 && S->wg[li] && wi % (64 / L.cin) == 0) ? 2 : (ks == 3 && s == 1 && L.cin >= 128) ? 1 : 0;
    const int cc = (P.kind == 1 && split) ? 32 : chunk_for(L.cin);
    P.cc = cc;
    // x2: the halo layers and the pixel-group layers run as CTA pairs (tcgen05 cta_group::2), half of the weight rows of a tile per CTA
    // Of the pixel-group layers only 64 -> 64 gains (206 -> 168 us: it is bound by the operand port and its weights filled the shared
    // memory); the HBM-bound 16 -> 32, 32 -> 32 and 32 -> 64 instances lose 30-40 % when two CTAs advance in lockstep (measured).
    P.pair = (S->f8 && (P.kind == 1 || (P.kind == 2 && L.cin == 64))) ? 1 : 0;
    // output box of the per-tap kernel: bw | wo, bh | ho by construction (padded sizes are multiples of 16 at level 0)
    int bw = 16 >> level_out; if (bw < 1) bw = 1;
    int bh = 8; while (bh > 1 && (ho % bh != 0 || bw * bh > 128)) bh >>= 1;
    while (wo % bw != 0 && bw > 1) bw >>= 1;
    const int bn = 128 / (bw * bh);
    ConvTcParams& p = P.p;
    p.taps = ks * ks; p.ks = ks; p.stride = s; p.cin = L.cin; p.kofs = 0; p.pad_top = pt; p.pad_left = pl;
    p.bw = bw; p.bh = bh; p.bn = bn;
    p.tiles_x = wo / bw; p.tiles_y = ho / bh;
    p.ho = ho; p.wo = wo; p.n = nb; p.relu = L.relu; p.scale = L.scale; p.shift = L.shift; p.out = out;
    p.fp16 = S->fp16;
    p.lo_n = nb; p.out_lo = (long long)nb * ho * wo * L.cout;
    p.n_tiles = p.tiles_x * p.tiles_y * ((nb + bn - 1) / bn);
    const CUtensorMapDataType dt16 = S->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint32_t e4[4] = {1, 1, 1, 1}, e2[2] = {1, 1};
    const cuuint64_t n_dim = (cuuint64_t)planes * nb;               // the lo plane = slices [nb, 2 nb)
    if (P.kind == 2) {
        const int g = 64 / L.cin, gout = s == 1 ? g : g / 2;
        const int trows = (split && L.cin == 64 && s == 1 && !P.pair) ? 14 : 16;   // ConvGroupCfg::TROWS
        const int pu = s == 1 ? 10 : 9, pr = s == 1 ? trows + 2 : 2 * trows + 1, jn = s == 1 ? g + 2 : g + 1;
        ConvGroupParams& gp = P.gp;
        gp.tiles_x = (wo / gout + 7) / 8; gp.tiles_y = (ho + trows - 1) / trows; gp.n_tiles = gp.tiles_x * gp.tiles_y * nb;
        gp.scale = L.scale; gp.shift = L.shift;
        gp.lo_n = nb; gp.out = (uint32_t*)out; gp.out_lo = p.out_lo / 2; gp.ho = ho; gp.wog = wo / gout;
        {   // input: rows of g pixels (128 bytes), box = halo patch
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_128B;
            cuuint64_t dims[4] = {64, (cuuint64_t)(wi / g), (cuuint64_t)hi, n_dim};
            cuuint64_t strides[3] = {128, (cuuint64_t)wi * L.cin * 2, (cuuint64_t)hi * wi * L.cin * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)pu, (cuuint32_t)pr, 1};
            UKBB_ENCODE("group patch", &P.map_a, dt16, 4, (void*)in, dims, strides, box, e4);
        }
        {   // expanded weights: [planes * 3 * J tiles * 64 rows][cin]
            const CUtensorMapSwizzle _sw = swizzle_for(L.cin);
            cuuint64_t dims[2] = {(cuuint64_t)L.cin, (cuuint64_t)(planes * 3 * jn * 64)};
            cuuint64_t strides[1] = {(cuuint64_t)L.cin * 2};
            cuuint32_t box[2] = {(cuuint32_t)L.cin, (cuuint32_t)(P.pair ? 32 : 64)};
            UKBB_ENCODE("group weights", &P.map_b, dt16, 2, (void*)S->wg[li], dims, strides, box, e2);
        }
        {   // output (plain modes): rows of gout pixels x cout channels = 64 elements, box = 16 rows x 8 groups
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_128B;
            cuuint64_t dims[4] = {64, (cuuint64_t)(wo / gout), (cuuint64_t)ho, n_dim};
            cuuint64_t strides[3] = {128, (cuuint64_t)wo * L.cout * 2, (cuuint64_t)ho * wo * L.cout * 2};
            cuuint32_t box[4] = {64, 8, 16, 1};
            UKBB_ENCODE("group output", &P.map_out, dt16, 4, (void*)out, dims, strides, box, e4);
        }
        P.valid = true;
        return UKBB_OK;
    }
    const CUtensorMapSwizzle _sw = swizzle_for(cc);
    if (P.kind == 1) {
        ConvHaloParams& hp = P.hp;
        hp.cin = L.cin; hp.chunks = L.cin / cc;
        hp.tiles_x = (wo + 15) / 16; hp.tiles_y = (ho + 15) / 16; hp.n_tiles = hp.tiles_x * hp.tiles_y * nb;
        hp.ho = ho; hp.wo = wo; hp.n = nb; hp.relu = L.relu; hp.fp16 = S->fp16;
        hp.lo_n = nb; hp.out_lo = p.out_lo;
        hp.scale = L.scale; hp.shift = L.shift; hp.out = out;
        cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)wi, (cuuint64_t)hi, n_dim};
        cuuint64_t strides[3] = {(cuuint64_t)L.cin * 2, (cuuint64_t)wi * L.cin * 2, (cuuint64_t)hi * wi * L.cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)cc, 18, 18, 1};
        UKBB_ENCODE("halo patch", &P.map_a, dt16, 4, (void*)in, dims, strides, box, e4);
    } else {
        // activation map: dims (C, W, H, N), box (cc, bw*s, bh*s, bn), traversal strides (1, s, s, 1)
        cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)wi, (cuuint64_t)hi, n_dim};
        cuuint64_t strides[3] = {(cuuint64_t)L.cin * 2, (cuuint64_t)wi * L.cin * 2, (cuuint64_t)hi * wi * L.cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)cc, (cuuint32_t)(bw * s), (cuuint32_t)(bh * s), (cuuint32_t)bn};
        cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
        UKBB_ENCODE("activation", &P.map_a, dt16, 4, (void*)in, dims, strides, box, estr);
    }
    {   // weights [planes * cout][taps * cin]: the lo plane = rows [cout, 2 cout)
        const int ktot = p.taps * L.cin;
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)(planes * L.cout)};
        cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)cc, (cuuint32_t)(P.pair ? L.cout / 2 : L.cout)};
        UKBB_ENCODE("weights", &P.map_b, dt16, 2, (void*)S->w[li], dims, strides, box, e2);
    }
    P.valid = true;
    return UKBB_OK;
}

int tc_prepare(Engine* h, const ukbb_fcn_weights* w) {
    TcState* S = new TcState();
    h->tc = S;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    UKBB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return UKBB_E_CUDA; }
    S->encode = (EncodeTiledFn)fn;
    S->fp16 = (h->mode == UKBB_MODE_FP16 || h->mode == UKBB_MODE_FP16X3) ? 1 : 0;
    S->fp16 = S->fp16 || h->mode == UKBB_MODE_FP16X2;
    S->f8 = h->mode == UKBB_MODE_FP16X2 ? 1 : 0;
    S->split = (h->mode == UKBB_MODE_BF16X3 || h->mode == UKBB_MODE_FP16X3 || S->f8) ? 1 : 0;
    const int fp16 = S->fp16, split = S->split;
    const int wsplit = S->f8 ? 2 : split;      // weights of the layers that read tensors from HBM: FP8 correction plane in the x2 scheme
    int rc;
    {   // class-score layer: FP32 weights [k][8] and bias, passed by value (constant bank)
        const ukbb_conv_weights& c = w->conv[UKBB_N_CONV - 1];
        for (int k = 0; k < 64; ++k)
            for (int co = 0; co < 8; ++co) S->h_wl[k * 8 + co] = co < c.cout ? c.kernel[(size_t)k * c.cout + co] : 0.f;
        for (int co = 0; co < 8; ++co) S->h_bias[co] = co < c.cout ? c.bias[co] : -INFINITY;
    }
    {   // conv0_0 on the tensor pipe (conv_first_tc.cuh): device tap (dy, dx) <- TF kernel[kh = dx][kw = dy], BN scale folded into the
        // FP32 weights; B0[n = pixel s * 16 + co][k = part * 18 + row r * 6 + column c] = w_hi | w_hi | w_lo of tap (r, kx = c - s)
        const ukbb_conv_weights& c = w->conv[0];
        float w0[9][16];
        for (int co = 0; co < 16; ++co) {
            const double sc = (double)c.gamma[co] / sqrt((double)c.moving_variance[co] + (double)w->bn_eps);
            S->c0_shift[co] = (float)((double)c.beta[co] - (double)c.moving_mean[co] * sc);
            for (int dy = 0; dy < 3; ++dy)
                for (int dx = 0; dx < 3; ++dx) w0[dy * 3 + dx][co] = (float)((double)c.kernel[(size_t)(dx * 3 + dy) * 16 + co] * sc);
        }
        std::vector<uint16_t> b0(64 * 64);
        for (int sp = 0; sp < 4; ++sp)
            for (int co = 0; co < 16; ++co)
                for (int k = 0; k < 64; ++k) {
                    float dummy;
                    uint16_t val = enc16(0.f, fp16, &dummy);
                    if (k < 54) {
                        const int part = k / 18, r = (k % 18) / 6, cx = k % 6, kx = cx - sp;
                        if (kx >= 0 && kx <= 2) {
                            const float wv = w0[r * 3 + kx][co];
                            float hi_f, lo_f;
                            const uint16_t hi = enc16(wv, fp16, &hi_f);
                            const uint16_t lo = enc16(wv - hi_f, fp16, &lo_f);
                            val = part < 2 ? hi : lo;
                        }
                    }
                    b0[(size_t)(sp * 16 + co) * 64 + k] = val;
                }
        UKBB_CUDA(cudaMalloc(&S->wb0, b0.size() * 2));
        UKBB_CUDA(cudaMemcpy(S->wb0, b0.data(), b0.size() * 2, cudaMemcpyHostToDevice));
        const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_128B;
        cuuint64_t d[2] = {64, 64}; cuuint64_t st1[1] = {128}; cuuint32_t bx[2] = {64, 64}; cuuint32_t e2[2] = {1, 1};
        UKBB_ENCODE("conv0_0 weights", &S->map_b0, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, S->wb0, d, st1, bx, e2);
    }
    for (int li : {13, 18, 19}) {          // 1x1 layers of the head: gamma / sqrt(var + eps) folded into the weights before rounding
        const ukbb_conv_weights& c = w->conv[li];
        std::vector<float> wb((size_t)c.cout * c.cin);
        for (int co = 0; co < c.cout; ++co) {
            const double sc = (double)c.gamma[co] / sqrt((double)c.moving_variance[co] + (double)w->bn_eps);
            S->h_shift[li][co] = (float)((double)c.beta[co] - (double)c.moving_mean[co] * sc);
            for (int ci = 0; ci < c.cin; ++ci) wb[(size_t)co * c.cin + ci] = (float)((double)c.kernel[(size_t)ci * c.cout + co] * sc);
        }
        rc = upload16(wb, fp16, li == 13 ? wsplit : split, &S->wf[li]);      // fc0 / fc1 multiply operands made in the kernel: (hi, lo) FP16
        if (rc) return rc;
    }
    for (int l = 1; l <= 4; ++l) {         // interpolation matrices of the tensor-core upsample (head_common.cuh); exact in 16 bits
        const int f = 1 << l, pb = (f - 1) / 2, kpad = HM_KPAD[l], pw = HM_PW[l], nv = l == 4 ? 2 : 1;
        std::vector<float> Uf((size_t)nv * 128 * kpad, 0.f);
        for (int v = 0; v < nv; ++v)
            for (int m = 0; m < 128; ++m) {
                const int ty = m >> 4, tx = m & 15;
                const int Y = (l == 4 ? 8 * v : 0) + ty + pb, X = tx + pb;
                const int ry = Y & (f - 1), py1 = (Y >> l) + 1, rx = X & (f - 1), px1 = (X >> l) + 1;
                const float wy1 = (float)(ry + 1) / (float)f, wy0 = 1.f - wy1, wx1 = (float)(rx + 1) / (float)f, wx0 = 1.f - wx1;
                float* row = &Uf[((size_t)v * 128 + m) * kpad];
                row[(py1 - 1) * pw + (px1 - 1)] += wy0 * wx0;
                row[(py1 - 1) * pw + px1] += wy0 * wx1;
                row[py1 * pw + (px1 - 1)] += wy1 * wx0;
                row[py1 * pw + px1] += wy1 * wx1;
            }
        rc = upload16(Uf, fp16, 0, &S->u[l]);
        if (rc) return rc;
    }
    for (int i = 1; i < UKBB_N_CONV - 1; ++i) {
        const ukbb_conv_weights& c = w->conv[i];
        const int taps = c.ksize * c.ksize, ktot = taps * c.cin;
        std::vector<float> wb((size_t)c.cout * ktot);
        for (int dy = 0; dy < c.ksize; ++dy)
            for (int dx = 0; dx < c.ksize; ++dx)
                for (int ci = 0; ci < c.cin; ++ci)
                    for (int co = 0; co < c.cout; ++co)      // device tap (dy,dx) <- TF kernel[kh=dx][kw=dy]
                        wb[(size_t)co * ktot + (dy * c.ksize + dx) * c.cin + ci] = c.kernel[((size_t)(dx * c.ksize + dy) * c.cin + ci) * c.cout + co];
        rc = upload16(wb, fp16, wsplit, &S->w[i]);
        if (rc) return rc;
        // pixel-group layers (conv_group.cuh): N = gout * cout = 64.  Tile (ky, j) row (s, co) holds tap (ky, kx) with
        // kx = j - s (stride 1) or j - 2 s (stride 2), zero where that tap does not exist.
        if (c.ksize == 3 && c.cin <= 64 && 64 % c.cin == 0) {
            const int g = 64 / c.cin, gout = c.stride == 1 ? g : g / 2;
            if (gout >= 1 && gout * c.cout == 64) {
                const int jn = c.stride == 1 ? g + 2 : g + 1;
                std::vector<float> we((size_t)3 * jn * 64 * c.cin, 0.f);
                for (int ky = 0; ky < 3; ++ky)
                    for (int j = 0; j < jn; ++j)
                        for (int sp = 0; sp < gout; ++sp)
                            for (int co = 0; co < c.cout; ++co)
                                for (int ci = 0; ci < c.cin; ++ci) {
                                    const int kx = c.stride == 1 ? j - sp : j - 2 * sp;
                                    if (kx >= 0 && kx <= 2)
                                        we[((size_t)(ky * jn + j) * 64 + sp * c.cout + co) * c.cin + ci] = wb[(size_t)co * ktot + (ky * 3 + kx) * c.cin + ci];
                                }
                rc = upload16(we, fp16, wsplit, &S->wg[i]);
                if (rc) return rc;
            }
        }
    }
    return UKBB_OK;
}

void tc_release(Engine* h) {
    TcState* S = h->tc;
    if (!S) return;
    for (int i = 0; i < UKBB_N_CONV; ++i) { cudaFree(S->w[i]); cudaFree(S->wg[i]); cudaFree(S->wf[i]); }
    for (int l = 0; l < 5; ++l) { cudaFree(S->t[l]); cudaFree(S->u[l]); }
    cudaFree(S->wb0);
    delete S;
    h->tc = nullptr;
}

// Workspace + plans for sub-batches of up to nb slices of h2 x w2 pixels.
static int ensure_plans(Engine* h, int nb, int h2, int w2) {
    TcState* S = h->tc;
    if (S->plan_nb == nb && S->plan_h == h2 && S->plan_w == w2) return UKBB_OK;
    UKBB_CUDA(cudaDeviceSynchronize());
    const int planes = S->split ? 2 : 1;
    // activation workspace (16-bit, `planes` planes each): encoder ping / pong per level, t_l per level
    for (int l = 0; l < 5; ++l) {
        cudaFree(h->ws.a[l]); cudaFree(h->ws.b[l]); cudaFree(S->t[l]);
        h->ws.a[l] = h->ws.b[l] = nullptr; S->t[l] = nullptr;
        const size_t px = (size_t)nb * (h2 >> l) * (w2 >> l);
        if (l > 0) UKBB_CUDA(cudaMalloc(&h->ws.a[l], planes * px * kNFilterTc[l] * 2));     // a[0] (conv0_0 output) never exists: conv_first_tc
        UKBB_CUDA(cudaMalloc(&h->ws.b[l], planes * px * kNFilterTc[l] * 2));
        if (l > 0) UKBB_CUDA(cudaMalloc(&S->t[l], planes * px * 64 * 2));
    }
    h->ws.nb = nb; h->ws.h = h2; h->ws.w = w2;
    int li = 0, rc;
    const __nv_bfloat16* cur = nullptr;
    const __nv_bfloat16* level_out[5];
    int hi = h2, wi = w2;
    for (int l = 0; l < 5; ++l) {
        for (int b = 0; b < kNBlockTc[l]; ++b, ++li) {
            __nv_bfloat16* dst = (__nv_bfloat16*)((b & 1) ? h->ws.b[l] : h->ws.a[l]);
            if (li > 0) {
                if (li == 1) cur = (const __nv_bfloat16*)h->ws.b[0];   // conv0_1's input a0 only ever exists in shared memory (conv_first_tc): placeholder of the same shape
                rc = make_plan(h, li, cur, dst, nb, hi, wi, l);
                if (rc) return rc;
                hi = S->plan[li].p.ho; wi = S->plan[li].p.wo;
            }
            cur = dst;
        }
        level_out[l] = cur;
    }
    if (S->plan[1].kind != 2) { set_error("forward: conv0_1 has no pixel-group plan for %d x %d", w2, h2); return UKBB_E_UNSUPPORTED; }
    const CUtensorMapDataType dt16 = S->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint32_t e4[4] = {1, 1, 1, 1}, e2[2] = {1, 1};
    {   // side_tc_kernel: level outputs as [pixels][Cin] matrices, same_dim weights, fc0 column blocks
        for (int l = 1; l <= 4; ++l) {
            const int cin = kNFilterTc[l], kc = cin < 64 ? cin : 64;
            const cuuint64_t rows = (cuuint64_t)nb * (h2 >> l) * (w2 >> l);
            const CUtensorMapSwizzle _sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
            cuuint64_t di[2] = {(cuuint64_t)cin, planes * rows}; cuuint64_t si[1] = {(cuuint64_t)cin * 2}; cuuint32_t bi[2] = {(cuuint32_t)kc, 128};
            UKBB_ENCODE("side input", &S->sm.in[l - 1], dt16, 2, (void*)level_out[l], di, si, bi, e2);
            cuuint64_t dw[2] = {(cuuint64_t)cin, (cuuint64_t)(planes * 32)}; cuuint64_t sw1[1] = {(cuuint64_t)cin * 2}; cuuint32_t bw2[2] = {(cuuint32_t)kc, 32};
            UKBB_ENCODE("same_dim weights", &S->sm.wsd[l - 1], dt16, 2, S->w[13 + l], dw, sw1, bw2, e2);
        }
        for (int l = 1; l <= 4; ++l) {
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_128B;
            const cuuint64_t rows = (cuuint64_t)nb * (h2 >> l) * (w2 >> l);
            cuuint64_t dout[2] = {64, planes * rows}; cuuint64_t so[1] = {128}; cuuint32_t bo[2] = {64, 128};
            UKBB_ENCODE("side output", &S->sm.out[l - 1], dt16, 2, S->t[l], dout, so, bo, e2);
        }
        const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_64B;
        cuuint64_t d0[2] = {160, (cuuint64_t)(planes * 64)}; cuuint64_t st0[1] = {320}; cuuint32_t b0[2] = {32, 64};
        UKBB_ENCODE("fc0 column blocks", &S->sm.w0, dt16, 2, S->wf[18], d0, st0, b0, e2);
    }
    {   // head_ts_kernel: t_l patches [planes * nb][h_l][w_l][64] with box (64, PW, PH, 1); weights of same_dim0 / fc0 (level 0) / fc1
        CUtensorMap* tm[5] = {nullptr, &S->hm.t1, &S->hm.t2, &S->hm.t3, &S->hm.t4};
        for (int l = 1; l <= 4; ++l) {
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_128B;
            cuuint64_t dims[4] = {64, (cuuint64_t)(w2 >> l), (cuuint64_t)(h2 >> l), (cuuint64_t)planes * nb};
            cuuint64_t strides[3] = {128, (cuuint64_t)(w2 >> l) * 128, (cuuint64_t)(h2 >> l) * (w2 >> l) * 128};
            cuuint32_t box[4] = {64, (cuuint32_t)HM_PW[l], (cuuint32_t)HM_PH[l], 1};
            UKBB_ENCODE("t patch", tm[l], dt16, 4, S->t[l], dims, strides, box, e4);
        }
        {
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_64B;
            cuuint64_t d0[2] = {160, (cuuint64_t)(planes * 64)}; cuuint64_t st0[1] = {320}; cuuint32_t b0[2] = {32, 64};
            UKBB_ENCODE("fc0 level-0 block", &S->hm.w0, dt16, 2, S->wf[18], d0, st0, b0, e2);
        }
        {
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_128B;
            cuuint64_t d1[2] = {64, (cuuint64_t)(planes * 64)}; cuuint64_t st1[1] = {128}; cuuint32_t b1[2] = {64, 64};
            UKBB_ENCODE("fc1 weights", &S->hm.w1, dt16, 2, S->wf[19], d1, st1, b1, e2);
        }
        {
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_32B;
            cuuint64_t dsd[2] = {16, (cuuint64_t)(planes * 32)}; cuuint64_t ssd[1] = {32}; cuuint32_t bsd[2] = {16, 32};
            UKBB_ENCODE("same_dim0 weights", &S->hm.wsd, dt16, 2, S->wf[13], dsd, ssd, bsd, e2);
        }
    }
    S->plan_nb = nb; S->plan_h = h2; S->plan_w = w2;
    return UKBB_OK;
}

// Test hook: run ONE tensor-core conv layer of the engine on a caller-provided NHWC tensor ([planes][n][hi][wi][cin] 16-bit).
int debug_conv_tc(Engine* h, int li, const void* in, int n, int hi, int wi, int level_out, void* out, cudaStream_t st) {
    UKBB_REQUIRE(h->tc, "debug_conv: engine is not in a tensor-core mode");
    TcState* S = h->tc;
    UKBB_REQUIRE(li >= 1 && li < (S->split ? 13 : UKBB_N_CONV - 1), "debug_conv: layer %d has no stand-alone tensor-core kernel in this mode", li);
    TcLayerPlan saved = S->plan[li];
    int rc = make_plan(h, li, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, hi, wi, level_out);
    if (!rc) rc = launch_plan(S->plan[li], S->fp16, h->sms, st);
    if (!rc) { cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) { set_error("debug_conv: %s", cudaGetErrorString(e)); rc = UKBB_E_CUDA; } }
    S->plan[li] = saved;
    h->launches++;
    return rc;
}

// lo plane formats: 0 = none, 1 = 16-bit values, 2 = x2 scheme (per 16 elements [e4m3(lo * 2^11) x 16 | e4m3(hi) x 16])
__global__ void widen16_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ out, long long n, int fp16, int lo_fmt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto cv = [&](uint16_t u) { return fp16 ? __half2float(__ushort_as_half(u)) : __uint_as_float((uint32_t)u << 16); };
    float l = 0.f;
    if (lo_fmt == 1) l = cv(lo[i]);
    if (lo_fmt == 2) {
        const __nv_fp8_storage_t b = reinterpret_cast<const uint8_t*>(lo)[(i / 16) * 32 + (i % 16)];
        l = __half2float(__half(__nv_cvt_fp8_to_halfraw(b, __NV_E4M3))) * (1.f / (float)(1 << tc::X2_SA));
    }
    out[i] = cv(hi[i]) + l;
}

// Test hook: copy an intermediate tensor of the most recent forward out as FP32 (hi + lo in the split modes).
// which: 0 = encoder ping buffer a[level], 1 = pong buffer b[level], 2 = t[level].
int debug_read_tc(Engine* h, int which, int level, float* out, long long n_elems, cudaStream_t st) {
    UKBB_REQUIRE(h->tc, "debug_read: engine is not in a tensor-core mode");
    TcState* S = h->tc;
    UKBB_REQUIRE(level >= 0 && level < 5 && which >= 0 && which <= 2, "debug_read: no tensor %d at level %d", which, level);
    const uint16_t* src = (const uint16_t*)(which == 0 ? h->ws.a[level] : which == 1 ? h->ws.b[level] : (void*)S->t[level]);
    UKBB_REQUIRE(src, "debug_read: tensor %d of level %d is not materialised", which, level);
    const size_t plane = (size_t)S->plan_nb * (S->plan_h >> level) * (S->plan_w >> level) * (which == 2 ? 64 : kNFilterTc[level]);
    UKBB_REQUIRE(n_elems > 0 && (size_t)n_elems <= plane, "debug_read: %lld elements requested, the tensor has %zu", n_elems, plane);
    const int lo_fmt = !S->split ? 0 : (S->f8 && which != 2) ? 2 : 1;          // t_l stays a (hi, lo) FP16 pair in the x2 scheme
    widen16_kernel<<<(unsigned)((n_elems + 255) / 256), 256, 0, st>>>(src, src + plane, out, n_elems, S->fp16, lo_fmt);
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}

int forward_tc(Engine* h, const float* image, int n, int x2, int y2, int x_pre, int y_pre, int x, int y,
               uint8_t* labels, float* logits, float* prob, unsigned long long* counts, cudaStream_t st) {
    TcState* S = h->tc;
    const int w2 = x2, h2 = y2;
    UKBB_REQUIRE((w2 >> 4) >= 1 && (h2 >> 4) >= 1, "forward: image too small");
    // One SA subject (500 slices) is ONE sub-batch: measured 151k -> 165k -> 172k slices/s for caps 125 / 250 / 500 (fewer launches,
    // fuller last waves); larger calls are split into equal parts (1000 slices -> 2 x 500 rather than 500 + 500 + 0).
    const int cap = 500;
    const int parts = (n + cap - 1) / cap;
    const int NB = (n + parts - 1) / parts;
    // plans (tensor maps, workspace) built for a larger sub-batch of the same image size serve any smaller one
    int rc = (S->plan_nb >= NB && S->plan_h == h2 && S->plan_w == w2) ? UKBB_OK : ensure_plans(h, NB, h2, w2);
    if (rc) return rc;
    const int cap_nb = S->plan_nb;
    for (int n0 = 0; n0 < n; n0 += NB) {
        const int nb = n - n0 < NB ? n - n0 : NB;
        {   // conv0_0 + conv0_1 in one launch; the image box map names this call's image pointer
            CUtensorMap map_img;
            const CUtensorMapSwizzle _sw = CU_TENSOR_MAP_SWIZZLE_NONE;
            cuuint64_t dims[3] = {(cuuint64_t)w2, (cuuint64_t)h2, (cuuint64_t)nb};
            cuuint64_t strides[2] = {(cuuint64_t)w2 * 4, (cuuint64_t)h2 * w2 * 4};
            cuuint32_t box[3] = {ConvFirstTcCfg<>::IMG_W, ConvFirstTcCfg<>::IMG_H, 1}, e3[3] = {1, 1, 1};
            UKBB_ENCODE("image boxes", &map_img, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)(image + (size_t)n0 * h2 * w2), dims, strides, box, e3);
            const TcLayerPlan& P1 = S->plan[1];
            ConvFirstParams fp;
            fp.tiles_x = P1.gp.tiles_x; fp.tiles_y = P1.gp.tiles_y; fp.n_tiles = fp.tiles_x * fp.tiles_y * nb;
            fp.h = h2; fp.w4 = w2 / 4;
            fp.scale = P1.gp.scale; fp.shift = P1.gp.shift;
            memcpy(fp.shift0, S->c0_shift, sizeof(fp.shift0));
            fp.out = P1.gp.out; fp.out_lo = P1.gp.out_lo; fp.lo_n = P1.gp.lo_n;
            rc = launch_first(S, P1, map_img, fp, h->sms, st);
            if (rc) return rc;
            h->launches++;
        }
        for (int li = 2; li < 13; ++li) {
            TcLayerPlan P = S->plan[li];
            P.p.n = nb;
            P.p.n_tiles = P.p.tiles_x * P.p.tiles_y * ((nb + P.p.bn - 1) / P.p.bn);
            P.hp.n = nb;
            P.hp.n_tiles = P.hp.tiles_x * P.hp.tiles_y * nb;
            P.gp.n_tiles = P.gp.tiles_x * P.gp.tiles_y * nb;
            rc = launch_plan(P, S->fp16, h->sms, st);
            if (rc) return rc;
            h->launches++;
        }
        {   // same_dim 1..4 + fc0 column blocks
            SideParams sp;
            int acc_tiles = 0;
            for (int k = 0; k < 4; ++k) {
                const int l = 4 - k;
                const long long rows = (long long)nb * (h2 >> l) * (w2 >> l), cap_rows = (long long)cap_nb * (h2 >> l) * (w2 >> l);
                sp.tile_start[k] = acc_tiles;
                acc_tiles += (int)((rows + 127) / 128);
                sp.scale[l - 1] = h->layers[13 + l].scale; sp.shift[l - 1] = h->layers[13 + l].shift;
                sp.rows[l - 1] = (int)rows; sp.lo_row[l - 1] = (int)cap_rows;
                sp.t[l - 1] = (uint32_t*)S->t[l]; sp.t_lo[l - 1] = cap_rows * 32;
            }
            sp.tile_start[4] = acc_tiles;
            rc = S->f8 ? launch_side_x2(S, sp, h->sms, st) : S->split ? launch_side_x3(S, sp, h->sms, st) : launch_side_16(S, sp, h->sms, st);
            if (rc) return rc;
            h->launches++;
        }
        {
            HeadParams hp;
            hp.n = nb; hp.h = h2; hp.w = w2;
            hp.tiles_x = w2 / 16; hp.tiles_y = h2 / 8; hp.n_tiles = hp.tiles_x * hp.tiles_y * nb;
            hp.x_pre = x_pre; hp.y_pre = y_pre; hp.x = x; hp.y = y;
            hp.nc = h->n_class;
            hp.lo_n = cap_nb;
            hp.b0_lo = (long long)cap_nb * h2 * w2 * 2;             // 16 channels x 2 bytes = 2 uint4 per pixel
            memcpy(hp.c_shift_sd0, S->h_shift[13], sizeof(hp.c_shift_sd0));
            memcpy(hp.c_shift0, S->h_shift[18], sizeof(hp.c_shift0));
            for (int l = 0; l < 5; ++l) hp.u_glob[l] = (const uint32_t*)S->u[l];
            hp.b0 = (const uint4*)h->ws.b[0];
            for (int k = 0; k < 64; ++k) {
                hp.c_nshift1[k] = -S->h_shift[19][k];
                for (int c = 0; c < 8; ++c) hp.c_wlc[k][c] = S->h_wl[k * 8 + c];
            }
            for (int c = 0; c < 8; ++c) {
                double acc = 0.0;
                for (int k = 0; k < 64; ++k) acc += (double)S->h_shift[19][k] * (double)S->h_wl[k * 8 + c];
                hp.c_bias2[c] = c < h->n_class ? (float)((double)S->h_bias[c] + acc) : -INFINITY;
            }
            const size_t po = (size_t)n0 * h2 * w2 * h->n_class;
            hp.labels = labels + (size_t)n0 * x * y;
            hp.logits = logits ? logits + po : nullptr;
            hp.prob = prob ? prob + po : nullptr;
            hp.counts = counts ? counts + (size_t)n0 * h->n_class : nullptr;
            cudaEvent_t kt0 = nullptr, kt1 = nullptr;
            if (h->ktimer) {                                        // ukbb_fcn_kernel_timer: events on the launching stream
                UKBB_CUDA(cudaEventCreate(&kt0)); UKBB_CUDA(cudaEventCreate(&kt1));
                UKBB_CUDA(cudaEventRecord(kt0, st));
            }
            rc = S->f8 ? launch_head_x2(S, hp, h->n_class, h->sms, st)
                 : S->split ? launch_head_x3(S, hp, h->n_class, h->sms, st) : launch_head_16(S, hp, h->n_class, h->sms, st);
            if (h->ktimer && !rc) { UKBB_CUDA(cudaEventRecord(kt1, st)); h->ktimer_ev.emplace_back(kt0, kt1); }
            if (rc) return rc;
            h->launches++;
        }
    }
    return UKBB_OK;
}

}  // namespace ukbb
