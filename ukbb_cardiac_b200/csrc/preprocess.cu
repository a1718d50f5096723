// Preprocessing kernels (north_star (d)): exact percentile intensity rescale + pad to x16.
// Restates common/image_utils.py:70-77 (rescale_intensity) as it executes under numpy 2:
//   vl, vh = np.percentile(image, (1, 99))      method 'linear', float64 result; for float32
//                                               input numpy lerps a + f32(b - a) * t in double
//                                               (b - f32(b - a) * (1 - t) when t >= 0.5)
//   image[image < vl] = vl; image[image > vh] = vh      (in place, rounded to float32)
//   out = (float32(image) - vl) / (vh - vl)             (double arithmetic)
// and common/deploy_network.py:97-107 (zero pad to multiples of 16, float32 cast).
//
// The order statistics are found EXACTLY by a 3-pass most-significant-digit radix select
// (12 + 12 + 8 bits) over the order-preserving integer image of the float32 bit pattern;
// all four ranks (floor/ceil of both percentiles) are resolved together so the volume is
// read three times for the select and once for the rescale (an SA volume, 80 MB, stays
// resident in the 126 MB L2 between passes).  HBM-bound integer/byte work: coalesced
// 128-bit loads, shared-memory histograms, no tensor cores.
//
// INTEGER FAST PATH.  DICOM-derived volumes (and the synthetic stacks) are integer valued: when every voxel is an integer in
// [0, 65535] the four order statistics follow EXACTLY from one 65536-bin counting pass (int_hist_kernel + int_scan_kernel)
// instead of three radix passes, and the rescale becomes a table lookup (lut_kernel evaluates the reference's double-precision
// expression once per integer level, so the result is bit-identical to the generic path's per-voxel FP64 division, which is what
// bounds rescale_pad_kernel on this GPU).  The decision is taken on the device (SelState::nonint / done): the generic kernels
// are always enqueued and return at once when the fast path has succeeded, so the call stays asynchronous.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace ukbb {

constexpr int NRANK = 4;
constexpr int BINS0 = 4096, BINS1 = 4096, BINS2 = 256;
// hist layout: [pass0: 4096][pass1: 4 x 4096][pass2: 4 x 256]
constexpr int HIST_WORDS = BINS0 + NRANK * BINS1 + NRANK * BINS2;
constexpr int HIST_INT = HIST_WORDS;       // offset of the 65536-bin integer histogram

struct SelState {                 // device-resident select state
    unsigned long long rank[NRANK];     // residual rank inside the current prefix
    unsigned int prefix[NRANK];         // key prefix resolved so far (12 then 24 then 32 bits)
    unsigned int nonint;                // set by int_hist_kernel when a voxel is not an integer in [0, 65535]
    unsigned int done;                  // set by int_scan_kernel: prefix[] already holds the four order statistics
    unsigned int maxlevel;              // largest integer level seen by int_hist_kernel (bounds the scan)
};
constexpr int INT_BINS = 65536, INT_SH = 4096;
constexpr int LUT_WORDS = INT_BINS + 2;  // [level] -> rescaled value; then the values of voxels clipped to vl / vh
constexpr int LUT_THR = LUT_WORDS;       // then t_lo, t_hi (float32 thresholds with the float64 truth table), fl, fh
constexpr int LUT_ALLOC = LUT_WORDS + 4;

__device__ __forceinline__ unsigned int f2key(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
    const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__global__ void sel_init_kernel(SelState* st, unsigned int* hist, unsigned long long r0,
                                unsigned long long r1, unsigned long long r2, unsigned long long r3) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HIST_WORDS + INT_BINS; i += gridDim.x * blockDim.x) hist[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->rank[0] = r0; st->rank[1] = r1; st->rank[2] = r2; st->rank[3] = r3;
        for (int r = 0; r < NRANK; ++r) st->prefix[r] = 0;
        st->nonint = 0; st->done = 0; st->maxlevel = 0;
    }
}

// INTEGER FAST PATH, counting pass: levels below 4096 (all of a 12-bit MR image) go through a shared-memory histogram with the
// same run folding as pass 0, higher levels straight to global atomics.
__global__ void __launch_bounds__(512)
int_hist_kernel(const float* __restrict__ vol, long long n, unsigned int* __restrict__ hist, SelState* __restrict__ st) {
    __shared__ unsigned int sh[INT_SH];
    for (int i = threadIdx.x; i < INT_SH; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned int* h16 = hist + HIST_INT;
    bool bad = false;
    int top = 0;
    auto level = [&](float f) -> int {                       // integer level of the voxel, or -1
        const int iv = __float2int_rz(f);
        const bool ok = f >= 0.f && f < 65536.f && (float)iv == f && __float_as_uint(f) != 0x80000000u;
        bad = bad || !ok;
        if (ok) top = max(top, iv);
        return ok ? iv : -1;
    };
    auto count = [&](int b, unsigned int c) {
        if (b < 0) return;
        if (b < INT_SH) atomicAdd(&sh[b], c); else atomicAdd(&h16[b], c);
    };
    const long long n4 = n >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(vol);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(v4 + i);
        const int b0 = level(v.x), b1 = level(v.y), b2 = level(v.z), b3 = level(v.w);
        if (b0 == b1 && b1 == b2 && b2 == b3) count(b0, 4u);
        else { count(b0, 1u); count(b1, 1u); count(b2, 1u); count(b3, 1u); }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) count(level(vol[(n4 << 2) + threadIdx.x]), 1u);
    if (bad) st->nonint = 1u;
    top = __reduce_max_sync(0xffffffffu, top);
    if ((threadIdx.x & 31) == 0 && top > 0) atomicMax(&st->maxlevel, (unsigned int)top);
    __syncthreads();
    for (int i = threadIdx.x; i < INT_SH; i += blockDim.x)
        if (sh[i]) atomicAdd(&h16[i], sh[i]);
}

// INTEGER FAST PATH, selection: one 1024-thread block scans the 65536 counters (64 per thread) and finds the level holding each of the
// four ranks; the levels are stored as float keys in prefix[] exactly where the third radix pass would have left them.
__global__ void __launch_bounds__(1024)
int_scan_kernel(SelState* st, const unsigned int* __restrict__ hist) {
    __shared__ unsigned long long s_scan[1024];
    if (st->nonint) return;                                  // uniform: the generic passes will run
    const unsigned int* h = hist + HIST_INT;
    const int t = threadIdx.x;
    const int per = (int)(st->maxlevel / 1024u) + 1;        // bins per thread: 4 for a 12-bit image, 64 at most
    unsigned long long local = 0;
    for (int i = 0; i < per; ++i) local += h[t * per + i];
    s_scan[t] = local;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const unsigned long long v = t >= off ? s_scan[t - off] : 0ull;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    const unsigned long long incl = s_scan[t], excl = incl - local;
    for (int r = 0; r < NRANK; ++r) {
        const unsigned long long want = st->rank[r];
        if (want >= excl && want < incl) {
            unsigned long long cum = excl;
            int b = t * per;
            for (int i = 0; i < per; ++i) {
                const unsigned int c = h[t * per + i];
                if (cum + c > want) { b = t * per + i; break; }
                cum += c;
            }
            st->prefix[r] = f2key((float)b);
        }
    }
    __syncthreads();
    if (t == 0) st->done = 1u;
}

// INTEGER FAST PATH, rescale table: the reference's expression (clip against the float64 thresholds, subtract and divide in double,
// round to float32) evaluated once per integer level -- identical operations, hence identical bits, to rescale_pad_kernel's slow path.
__global__ void lut_kernel(const SelState* __restrict__ st, const double* __restrict__ vlvh, float* __restrict__ lut) {
    if (!st->done) return;
    const double vl = vlvh[0], vh = vlvh[1];
    const float fl = (float)vl, fh = (float)vh;
    const double den = vh - vl;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < LUT_WORDS; i += gridDim.x * blockDim.x) {
        float v = i < INT_BINS ? (float)i : (i == INT_BINS ? fl : fh);
        if (i < INT_BINS) {
            if ((double)v < vl) v = fl;
            if ((double)v > vh) v = fh;
        }
        lut[i] = (float)(((double)v - vl) / den);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        lut[LUT_THR] = (double)fl >= vl ? fl : nextafterf(fl, INFINITY);
        lut[LUT_THR + 1] = (double)fh <= vh ? fh : nextafterf(fh, -INFINITY);
        lut[LUT_THR + 2] = fl;
        lut[LUT_THR + 3] = fh;
    }
}

// INTEGER FAST PATH, rescale + zero-pad for rows that are whole float4s (x and x_pre multiples of 4: every SA volume): one thread =
// one aligned float4 of the padded output, no index divisions (3-D grid), thresholds precomputed, four table lookups.
__global__ void __launch_bounds__(256)
rescale_lut_kernel(float* __restrict__ vol, float* __restrict__ out, const SelState* __restrict__ st, const float* __restrict__ lut,
                   int x, int y, int x2, int y2, int x_pre, int y_pre, int clip_in_place) {
    if (!st->done) return;                                   // rescale_pad_kernel does the work
    const int q = blockIdx.x * blockDim.x + threadIdx.x;     // float4 index inside the padded row
    const int oy = blockIdx.y * blockDim.y + threadIdx.y;
    const long long n = blockIdx.z;
    if (q >= (x2 >> 2) || oy >= y2) return;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    const int sy = oy - y_pre, sx0 = q * 4 - x_pre;
    if (sy >= 0 && sy < y && sx0 >= 0 && sx0 < x) {
        const float t_lo = lut[LUT_THR], t_hi = lut[LUT_THR + 1], fl = lut[LUT_THR + 2], fh = lut[LUT_THR + 3];
        float4* src = reinterpret_cast<float4*>(vol + (n * y + sy) * (long long)x + sx0);
        float4 v = *src;
        bool changed = false;
        auto one = [&](float& f) -> float {
            int li = (int)f;
            if (f < t_lo) { f = fl; li = INT_BINS; changed = true; }
            if (f > t_hi) { f = fh; li = INT_BINS + 1; changed = true; }
            return __ldg(lut + li);
        };
        r.x = one(v.x); r.y = one(v.y); r.z = one(v.z); r.w = one(v.w);
        if (clip_in_place && changed) *src = v;
    }
    reinterpret_cast<float4*>(out)[(n * y2 + oy) * (long long)(x2 >> 2) + q] = r;
}

// PASS 0: 4096-bin histogram of key >> 20, shared-memory privatised.  Each thread folds runs of
// equal bins among its own consecutive voxels before touching shared memory (neighbouring
// voxels of an MR image mostly share the top 12 key bits).
__global__ void __launch_bounds__(512)
sel_hist0_kernel(const float* __restrict__ vol, long long n, unsigned int* __restrict__ hist, const SelState* __restrict__ st) {
    __shared__ unsigned int sh[BINS0];
    if (st->done) return;                                    // integer fast path succeeded
    for (int i = threadIdx.x; i < BINS0; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const long long n4 = n >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(vol);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(v4 + i);
        const unsigned int b0 = f2key(v.x) >> 20, b1 = f2key(v.y) >> 20, b2 = f2key(v.z) >> 20,
                           b3 = f2key(v.w) >> 20;
        if (b0 == b1 && b1 == b2 && b2 == b3) {
            atomicAdd(&sh[b0], 4u);
        } else {
            atomicAdd(&sh[b0], 1u); atomicAdd(&sh[b1], 1u); atomicAdd(&sh[b2], 1u); atomicAdd(&sh[b3], 1u);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) atomicAdd(&sh[f2key(vol[(n4 << 2) + threadIdx.x]) >> 20], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < BINS0; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// PASS 1 / 2: voxels whose resolved prefix matches one of the four ranks vote on the next digit.
// Matches are ~1/4096 of the voxels for natural images, so they go to global atomics directly.
template <int PASS>
__global__ void __launch_bounds__(512)
sel_histn_kernel(const float* __restrict__ vol, long long n, const SelState* __restrict__ st,
                 unsigned int* __restrict__ hist) {
    if (st->done) return;                                    // integer fast path succeeded
    unsigned int pre[NRANK];
#pragma unroll
    for (int r = 0; r < NRANK; ++r) pre[r] = st->prefix[r];
    constexpr int SHIFT = PASS == 1 ? 20 : 8;
    auto vote = [&](float f) {
        const unsigned int k = f2key(f);
        const unsigned int p = k >> SHIFT;
#pragma unroll
        for (int r = 0; r < NRANK; ++r) {
            if (p == pre[r]) {
                // identical prefixes share the work: later ranks with the same prefix read rank r's bins
                bool first = true;
#pragma unroll
                for (int q = 0; q < r; ++q) first = first && (pre[q] != pre[r]);
                if (first) {
                    if (PASS == 1) atomicAdd(&hist[BINS0 + r * BINS1 + ((k >> 8) & 0xFFFu)], 1u);
                    else atomicAdd(&hist[BINS0 + NRANK * BINS1 + r * BINS2 + (k & 0xFFu)], 1u);
                }
            }
        }
    };
    const long long n4 = n >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(vol);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(v4 + i);
        vote(v.x); vote(v.y); vote(v.z); vote(v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) vote(vol[(n4 << 2) + threadIdx.x]);
}

// After each pass: locate, for every rank, the bin holding it.  One 1024-thread block: every thread
// owns bins/1024 consecutive bins, a shared-memory inclusive scan finds the owning thread.  Ranks that
// share a prefix share the histogram of the first such rank.
template <int PASS>
__global__ void __launch_bounds__(1024)
sel_scan_kernel(SelState* st, const unsigned int* __restrict__ hist) {
    __shared__ unsigned long long s_scan[1024];
    __shared__ unsigned int s_pre[NRANK];
    __shared__ unsigned long long s_newrank[NRANK];
    __shared__ unsigned int s_newpre[NRANK];
    const int t = threadIdx.x;
    if (st->done) return;                                    // integer fast path succeeded (uniform)
    if (t < NRANK) s_pre[t] = st->prefix[t];
    __syncthreads();
    constexpr int bins = PASS == 2 ? BINS2 : BINS0;
    constexpr int per = bins >= 1024 ? bins / 1024 : 1;
    for (int r = 0; r < NRANK; ++r) {
        int src = r;
        if (PASS > 0)
            for (int q = r - 1; q >= 0; --q)
                if (s_pre[q] == s_pre[r]) src = q;
        const unsigned int* h = PASS == 0 ? hist : PASS == 1 ? hist + BINS0 + src * BINS1
                                                             : hist + BINS0 + NRANK * BINS1 + src * BINS2;
        unsigned int c[per];
        unsigned long long local = 0;
#pragma unroll
        for (int i = 0; i < per; ++i) {
            const int bin = t * per + i;
            c[i] = bin < bins ? h[bin] : 0u;
            local += c[i];
        }
        s_scan[t] = local;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const unsigned long long v = t >= off ? s_scan[t - off] : 0ull;
            __syncthreads();
            s_scan[t] += v;
            __syncthreads();
        }
        const unsigned long long incl = s_scan[t], excl = incl - local, want = st->rank[r];
        const unsigned long long total = s_scan[1023];
        const bool owner = (want >= excl && want < incl) || (want >= total && t == 1023 && local == 0 && false);
        if (owner) {
            unsigned long long cum = excl;
            int b = t * per;
#pragma unroll
            for (int i = 0; i < per; ++i) {
                if (cum + c[i] > want) { b = t * per + i; break; }
                cum += c[i];
            }
            s_newrank[r] = want - cum;
            s_newpre[r] = PASS == 0 ? (unsigned)b : PASS == 1 ? ((s_pre[r] << 12) | (unsigned)b) : ((s_pre[r] << 8) | (unsigned)b);
        }
        __syncthreads();
    }
    if (t < NRANK) { st->rank[t] = s_newrank[t]; st->prefix[t] = s_newpre[t]; }
}

__global__ void sel_final_kernel(const SelState* __restrict__ st, double t_lo, double t_hi,
                                 double* __restrict__ vlvh, float* __restrict__ sel, double* vlvh_user) {
    if (threadIdx.x != 0) return;
    float v[NRANK];
    for (int r = 0; r < NRANK; ++r) { v[r] = key2f(st->prefix[r]); sel[r] = v[r]; }
    // numpy _lerp on float32 neighbours: diff in float32, interpolation in float64
    auto lerp = [](float a, float b, double t) {
        const float diff = b - a;
        return t >= 0.5 ? (double)b - (double)diff * (1.0 - t) : (double)a + (double)diff * t;
    };
    const double vl = lerp(v[0], v[1], t_lo), vh = lerp(v[2], v[3], t_hi);
    vlvh[0] = vl; vlvh[1] = vh;
    if (vlvh_user) { vlvh_user[0] = vl; vlvh_user[1] = vh; }
}

// Rescale + zero-pad: one thread = 4 consecutive output pixels along X of the padded slice.
__global__ void __launch_bounds__(256)
rescale_pad_kernel(float* __restrict__ vol, float* __restrict__ out, const double* __restrict__ vlvh,
                   long long total4, int x, int y, int x2, int y2, int x_pre, int y_pre, int clip_in_place,
                   const SelState* __restrict__ st, const float* __restrict__ lut, int lut_kernel_enqueued) {
    if (lut_kernel_enqueued && st->done) return;             // rescale_lut_kernel has written the output
    for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += (long long)gridDim.x * blockDim.x) {
    const double vl = vlvh[0], vh = vlvh[1];
    const float fl = (float)vl, fh = (float)vh;     // what `image[image < vl] = vl` stores
    const double den = vh - vl;
    // float32-vs-float64 comparisons as float32 comparisons with the same truth table: (double)v < vl  <=>  v < t_lo with t_lo the
    // smallest float >= vl, and (double)v > vh  <=>  v > t_hi with t_hi the largest float <= vh (FP64 compares per voxel made the
    // kernel issue-bound)
    const float t_lo = (double)fl >= vl ? fl : nextafterf(fl, INFINITY);
    const float t_hi = (double)fh <= vh ? fh : nextafterf(fh, -INFINITY);
    const bool fast = st->done != 0;                         // integer volume: table lookup instead of an FP64 division per voxel
    const unsigned int xq = (unsigned int)x2 >> 2;
    int ox, oy;
    long long n;
    if (total4 < (1ll << 31)) {                              // 32-bit index arithmetic (64-bit divisions cost more than the rescale)
        const unsigned int i = (unsigned int)i4, t = i / xq;
        ox = (int)(i - t * xq) * 4;
        const unsigned int nn = t / (unsigned int)y2;
        oy = (int)(t - nn * (unsigned int)y2);
        n = nn;
    } else {
        ox = (int)(i4 % xq) * 4;
        oy = (int)((i4 / xq) % y2);
        n = i4 / ((long long)xq * y2);
    }
    const int sy = oy - y_pre;
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (sy >= 0 && sy < y) {
        float* row = vol + (n * y + sy) * (long long)x;
        const int sx0 = ox - x_pre;
        float vin[4] = {0.f, 0.f, 0.f, 0.f};
        const bool vec = ((x | x_pre) & 3) == 0 && sx0 >= 0 && sx0 < x;     // the four pixels are one aligned float4 (SA volumes)
        if (vec) {
            const float4 q4 = *reinterpret_cast<const float4*>(row + sx0);
            vin[0] = q4.x; vin[1] = q4.y; vin[2] = q4.z; vin[3] = q4.w;
        }
        bool changed = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int sx = sx0 + u;
            if (sx >= 0 && sx < x) {
                float v = vec ? vin[u] : row[sx];
                // comparisons against the float64 thresholds, like numpy's float32-vs-float64 compare
                int li = fast ? (int)v : 0;
                if (v < t_lo) { v = fl; li = INT_BINS; changed = true; }
                if (v > t_hi) { v = fh; li = INT_BINS + 1; changed = true; }
                if (clip_in_place && !vec) row[sx] = v;
                vin[u] = v;
                r[u] = fast ? __ldg(lut + li) : (float)(((double)v - vl) / den);
            }
        }
        if (clip_in_place && vec && changed) *reinterpret_cast<float4*>(row + sx0) = make_float4(vin[0], vin[1], vin[2], vin[3]);
    }
    reinterpret_cast<float4*>(out)[i4] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

int preproc_alloc(PreprocWorkspace& ws) {
    UKBB_CUDA(cudaMalloc(&ws.hist, (HIST_WORDS + INT_BINS) * sizeof(unsigned int)));
    UKBB_CUDA(cudaMalloc(&ws.lut, LUT_ALLOC * sizeof(float)));
    UKBB_CUDA(cudaMalloc(&ws.state, sizeof(SelState)));
    UKBB_CUDA(cudaMalloc(&ws.vlvh, 2 * sizeof(double)));
    UKBB_CUDA(cudaMalloc(&ws.sel, NRANK * sizeof(float)));
    return UKBB_OK;
}

void preproc_free(PreprocWorkspace& ws) {
    cudaFree(ws.hist); cudaFree(ws.state); cudaFree(ws.vlvh); cudaFree(ws.sel); cudaFree(ws.lut);
    ws = PreprocWorkspace();
}

int launch_preprocess(PreprocWorkspace& ws, float* vol, long long n_slices, int x, int y, double q_lo,
                      double q_hi, int x2, int y2, int x_pre, int y_pre, float* out, double* vl_vh_out,
                      int clip_in_place, cudaStream_t st, long long* launches) {
    const long long n = n_slices * x * y;
    UKBB_REQUIRE(n > 0, "preprocess: empty volume");
    UKBB_REQUIRE(x2 % 16 == 0 && y2 % 16 == 0 && x2 >= x + x_pre && y2 >= y + y_pre && x_pre >= 0 && y_pre >= 0,
                 "preprocess: bad padded size %dx%d for %dx%d pre (%d,%d)", x2, y2, x, y, x_pre, y_pre);
    UKBB_REQUIRE(((uintptr_t)vol & 15) == 0 && ((uintptr_t)out & 15) == 0, "preprocess: buffers must be 16-byte aligned");
    // numpy 'linear': virtual index (n-1)*q/100, neighbours floor / floor+1 (clipped), gamma = frac
    double t[2];
    unsigned long long rk[4];
    const double qs[2] = {q_lo, q_hi};
    for (int i = 0; i < 2; ++i) {
        UKBB_REQUIRE(qs[i] >= 0.0 && qs[i] <= 100.0, "preprocess: percentile %g outside [0,100]", qs[i]);
        const double vi = (double)(n - 1) * (qs[i] / 100.0);
        const double fl = floor(vi);
        t[i] = vi - fl;
        unsigned long long lo = (unsigned long long)fl;
        if (lo > (unsigned long long)(n - 1)) lo = n - 1;
        unsigned long long hi = lo + 1 > (unsigned long long)(n - 1) ? (unsigned long long)(n - 1) : lo + 1;
        rk[2 * i] = lo; rk[2 * i + 1] = hi;
    }
    SelState* state = reinterpret_cast<SelState*>(ws.state);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int hgrid = sms * 4;       // 4 resident 512-thread CTAs per SM
    sel_init_kernel<<<16, 256, 0, st>>>(state, ws.hist, rk[0], rk[1], rk[2], rk[3]);
    if (!ws.no_int_path) {
        int_hist_kernel<<<hgrid, 512, 0, st>>>(vol, n, ws.hist, state);
        int_scan_kernel<<<1, 1024, 0, st>>>(state, ws.hist);
        if (launches) *launches += 2;
    }
    sel_hist0_kernel<<<hgrid, 512, 0, st>>>(vol, n, ws.hist, state);
    sel_scan_kernel<0><<<1, 1024, 0, st>>>(state, ws.hist);
    sel_histn_kernel<1><<<hgrid, 512, 0, st>>>(vol, n, state, ws.hist);
    sel_scan_kernel<1><<<1, 1024, 0, st>>>(state, ws.hist);
    sel_histn_kernel<2><<<hgrid, 512, 0, st>>>(vol, n, state, ws.hist);
    sel_scan_kernel<2><<<1, 1024, 0, st>>>(state, ws.hist);
    sel_final_kernel<<<1, 32, 0, st>>>(state, t[0], t[1], ws.vlvh, ws.sel, vl_vh_out);
    lut_kernel<<<32, 256, 0, st>>>(state, ws.vlvh, ws.lut);
    const long long total4 = n_slices * y2 * (x2 / 4);
    const int vec_rows = ((x | x_pre) & 3) == 0 && n_slices <= 65535 && !ws.no_int_path;
    if (vec_rows) {
        const int xq = x2 >> 2, bx = xq >= 64 ? 64 : (xq + 15) / 16 * 16, by = 256 / bx;
        dim3 block(bx, by), grid((xq + bx - 1) / bx, (y2 + by - 1) / by, (unsigned)n_slices);
        rescale_lut_kernel<<<grid, block, 0, st>>>(vol, out, state, ws.lut, x, y, x2, y2, x_pre, y_pre, clip_in_place);
        if (launches) *launches += 1;
    }
    // grid-stride with a bounded grid: when the table-lookup kernel has done the work this launch is 2368 blocks that return at once
    const long long pad_blocks = (total4 + 255) / 256, pad_cap = (long long)sms * 16;
    rescale_pad_kernel<<<(unsigned)(pad_blocks < pad_cap ? pad_blocks : pad_cap), 256, 0, st>>>(vol, out, ws.vlvh, total4, x, y, x2, y2,
                                                                          x_pre, y_pre, clip_in_place, state, ws.lut, vec_rows);
    if (launches) *launches += 10;
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}

// Rescale + pad with thresholds the caller already has (a block of a sequence whose percentiles were taken over the WHOLE sequence
// elsewhere: SURVEY 8(e), one subject split over several GPUs).  Same arithmetic as the generic path of launch_preprocess.
__global__ void set_thresholds_kernel(SelState* st, double* vlvh, double vl, double vh) {
    vlvh[0] = vl; vlvh[1] = vh;
    st->done = 0;                     // no lookup table: every voxel goes through the float64 formula
}

int launch_rescale_given(PreprocWorkspace& ws, float* vol, long long n_slices, int x, int y, double vl, double vh, int x2, int y2,
                         int x_pre, int y_pre, float* out, int clip_in_place, cudaStream_t st, long long* launches) {
    UKBB_REQUIRE(n_slices > 0 && x > 0 && y > 0, "rescale: empty volume");
    UKBB_REQUIRE(x2 % 16 == 0 && y2 % 16 == 0 && x2 >= x + x_pre && y2 >= y + y_pre && x_pre >= 0 && y_pre >= 0,
                 "rescale: bad padded size %dx%d for %dx%d pre (%d,%d)", x2, y2, x, y, x_pre, y_pre);
    UKBB_REQUIRE(((uintptr_t)vol & 15) == 0 && ((uintptr_t)out & 15) == 0, "rescale: buffers must be 16-byte aligned");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    SelState* state = reinterpret_cast<SelState*>(ws.state);
    set_thresholds_kernel<<<1, 1, 0, st>>>(state, ws.vlvh, vl, vh);
    const long long total4 = n_slices * y2 * (x2 / 4);
    const long long pad_blocks = (total4 + 255) / 256, pad_cap = (long long)sms * 16;
    rescale_pad_kernel<<<(unsigned)(pad_blocks < pad_cap ? pad_blocks : pad_cap), 256, 0, st>>>(vol, out, ws.vlvh, total4, x, y, x2, y2,
                                                                          x_pre, y_pre, clip_in_place, state, ws.lut, 0);
    if (launches) *launches += 2;
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}

}  // namespace ukbb
