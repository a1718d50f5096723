// Rate probes for the conv kernels' building blocks on B200 (standalone, not part of the library):
//   1. TMA tiled box-load throughput per SM for the box shapes the kernels use (halo patches with
//      32 / 64 / 128-byte rows, "quad-folded" 128-byte rows, head tiles), at several in-flight
//      depths, with the source tensor L2-resident or streamed from HBM;
//   2. tcgen05.mma issue rate for M = 128, N in {16..256}, K = 16 from shared memory, with dense
//      (SBO = 8 rows) and halo-window (SBO = 18 rows) A descriptors.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ukbb_cardiac_b200/csrc experiments/rate_probe.cu -o experiments/bin/rate_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace ukbb::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmaCase {
    int xstep, xofs, ystep, yofs;   // box origin = (tx * xstep + xofs, ty * ystep + yofs)
    int tiles_x, tiles_y, nb;
    int bytes;                      // bytes per load (whole box)
    int stage_bytes;                // 1024-aligned
    int iters, depth, dshift;
};

__global__ void __launch_bounds__(128, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap map, const TmaCase c, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + c.depth * c.stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < c.depth; ++s) mbar_init(bar + 8 * s, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // tile walk without divisions in the timed loop (a 64-bit div costs hundreds of cycles on one thread)
        int tx = blockIdx.x % c.tiles_x, ty = (blockIdx.x / c.tiles_x) % c.tiles_y, n = (blockIdx.x / (c.tiles_x * c.tiles_y)) % c.nb;
        const int dx = gridDim.x % c.tiles_x, dy = (gridDim.x / c.tiles_x) % c.tiles_y, dn = (gridDim.x / (c.tiles_x * c.tiles_y)) % c.nb;
        const long long t0 = clock64();
        for (int i = 0; i < c.iters; ++i) {
            const int s = i & (c.depth - 1);
            if (i >= c.depth) mbar_wait(bar + 8 * s, (uint32_t)((i >> c.dshift) - 1) & 1u);
            mbar_arrive_expect_tx(bar + 8 * s, c.bytes);
            tma_load_4d(base + s * c.stage_bytes, &map, bar + 8 * s, 0, tx * c.xstep + c.xofs, ty * c.ystep + c.yofs, n);
            tx += dx; if (tx >= c.tiles_x) { tx -= c.tiles_x; ++ty; }
            ty += dy; if (ty >= c.tiles_y) { ty -= c.tiles_y; ++n; }
            n += dn; if (n >= c.nb) n -= c.nb;
        }
        for (int i = c.iters; i < c.iters + c.depth; ++i) {
            const int s = i & (c.depth - 1);
            if (i >= c.depth) mbar_wait(bar + 8 * s, (uint32_t)((i >> c.dshift) - 1) & 1u);
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

static EncodeTiledFn g_enc;

static void run_tma(const char* name, void* buf, int C, int W, int H, int nb_total, int fold /*pixels per row*/, int bx, int by,
                    int xstep, int xofs, int ystep, int yofs, CUtensorMapSwizzle sw, bool l2_resident) {
    const int inner = C * fold;
    const int nb = l2_resident ? 4 : nb_total;
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)(W / fold), (cuuint64_t)H, (cuuint64_t)nb};
    cuuint64_t strides[3] = {(cuuint64_t)inner * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMap map;
    CUresult r = g_enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("%s: encode failed %d\n", name, (int)r); return; }
    long long* dout; CK(cudaMalloc(&dout, 148 * 8));
    for (int depth : {1, 2, 4, 8}) {
        if (depth & (depth - 1)) continue;
        TmaCase c;
        c.xstep = xstep; c.xofs = xofs; c.ystep = ystep; c.yofs = yofs;
        c.tiles_x = W / (xstep * fold);
        c.tiles_y = H / ystep; c.nb = nb;
        c.bytes = inner * 2 * bx * by;
        c.stage_bytes = (c.bytes + 1023) / 1024 * 1024;
        if ((size_t)depth * c.stage_bytes > 200 * 1024) continue;
        const long long tiles_total = (long long)c.tiles_x * c.tiles_y * nb_total;
        c.iters = l2_resident ? 600 : (int)(tiles_total / 148);
        c.depth = depth; c.dshift = depth == 1 ? 0 : depth == 2 ? 1 : depth == 4 ? 2 : 3;
        const int smem = depth * c.stage_bytes + 1024 + 128;
        CK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        if (l2_resident) { TmaCase w = c; w.iters = 100; tma_rate_kernel<<<148, 128, smem>>>(map, w, dout); }
        CK(cudaEventRecord(e0));
        tma_rate_kernel<<<148, 128, smem>>>(map, c, dout);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148]; CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
        printf("%-34s %-4s depth %d: %7.0f cyc/load  %6.1f B/cyc/SM  box %6d B (%3d rows x %3d B)  agg %7.1f GB/s\n", name,
               l2_resident ? "L2" : "HBM", depth, avg / c.iters, (double)c.bytes * c.iters / avg, c.bytes, bx * by, inner * 2,
               (double)c.bytes * c.iters * 148 / (ms * 1e-3) / 1e9);
    }
    cudaFree(dout);
}

// ------------------------------------------------------------------------------------------ MMA rate
template <int N, int RB, bool BMN = false>
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int sbo_rows, int n_mma, int windows, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a_base = base;                       // 324 rows x RB (+ slack)
    const uint32_t b_base = base + 48 * 1024;           // 9 tiles of N rows x RB
    constexpr uint32_t B_TILE = (N * RB + 1023) / 1024 * 1024;
    constexpr int NBT = 9 * B_TILE <= 144 * 1024 ? 9 : (144 * 1024) / B_TILE;
    const uint32_t bar = b_base + NBT * B_TILE;
    const uint32_t slot = bar + 8;
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (48 * 1024 + NBT * B_TILE) / 16; i += 128)
        reinterpret_cast<uint4*>(raw + (base - smem_u32(raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N) | (BMN ? (1u << 16) : 0u);
        constexpr uint32_t layout = RB == 128 ? 2u : RB == 64 ? 4u : 6u;
        const uint32_t a_hi = (uint32_t)((sbo_rows * RB) >> 4) | (1u << 14) | (layout << 29);
        const uint32_t b_hi = (uint32_t)((8 * RB) >> 4) | (1u << 14) | (layout << 29);
        const uint32_t a_lo = ((a_base & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t b_lo = ((b_base & 0x3FFFF) >> 4) | (1u << 16);
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; i += 9 * (RB / 32)) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                for (int k = 0; k < RB / 32; ++k) {
                    const int wtap = windows ? tap : 0;
                    umma_bf16_lohi(tmem + (N <= 128 ? ((i / 9) & 1) * N : 0), a_lo + ((((wtap / 3) * 18 + (wtap % 3)) * RB + k * 32) >> 4), a_hi,
                                   BMN ? b_lo + (((tap % 4) * 2048) >> 4) : b_lo + (((tap % NBT) * B_TILE + k * 32) >> 4),
                                   BMN ? ((uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29)) : b_hi, idesc, 1u);
                }
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int RB, bool BMN = false>
static void run_mma(int sbo_rows, int windows) {
    long long* dout; CK(cudaMalloc(&dout, 148 * 8));
    const int bt = (N * RB + 1023) / 1024 * 1024;
    const int smem = 48 * 1024 + (9 * bt <= 144 * 1024 ? 9 : (144 * 1024) / bt) * bt + 1024 + 64;
    CK(cudaFuncSetAttribute(mma_rate_kernel<N, RB, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int n_mma = 9 * (RB / 32) * 200;
    mma_rate_kernel<N, RB, BMN><<<148, 128, smem>>>(sbo_rows, n_mma, windows, dout);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    printf("MMA%s M=128 N=%3d K=16 row %3d B  SBO %2d rows windows=%d : %6.1f cyc/MMA  (floor N/2 = %d)  %5.1f%% of tensor peak\n", BMN ? " B-MN-major" : "", N, RB,
           sbo_rows, windows, avg / n_mma, N / 2, 100.0 * (N / 2.0) / (avg / n_mma));
    cudaFree(dout);
}

int main(int argc, char** argv) {
    const bool only_mn = argc > 1;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_enc = (EncodeTiledFn)fn;
    const int W = 192, H = 208, NB = 128;
    void* buf; CK(cudaMalloc(&buf, (size_t)NB * W * H * 64 * 2));          // up to 64 channels at level-0 size
    CK(cudaMemset(buf, 0, (size_t)NB * W * H * 64 * 2));
    if (only_mn) {
        void* fn2 = nullptr; cudaDriverEntryPointQueryResult q2; CK(cudaFree(0));
        run_mma<64, 128, true>(8, 0); run_mma<64, 64, true>(8, 0); run_mma<64, 32, true>(8, 0); run_mma<64, 128, false>(8, 0); run_mma<64, 64, false>(8, 0);
        return 0;
    }
    for (int l2 = 1; l2 >= 0; --l2) {
        run_tma("halo 18x18 C=16 (32 B rows)", buf, 16, W, H, NB, 1, 18, 18, 16, -1, 16, -1, CU_TENSOR_MAP_SWIZZLE_32B, l2);
        run_tma("halo 18x6quad C=16 (128 B rows)", buf, 16, W, H, NB, 4, 6, 18, 4, -1, 16, -1, CU_TENSOR_MAP_SWIZZLE_128B, l2);
        run_tma("halo 18x18 C=32 (64 B rows)", buf, 32, W, H, NB, 1, 18, 18, 16, -1, 16, -1, CU_TENSOR_MAP_SWIZZLE_64B, l2);
        run_tma("halo 18x10pair C=32 (128 B rows)", buf, 32, W, H, NB, 2, 10, 18, 8, -1, 16, -1, CU_TENSOR_MAP_SWIZZLE_128B, l2);
        run_tma("halo 18x18 C=64 (128 B rows)", buf, 64, W, H, NB / 2, 1, 18, 18, 16, -1, 16, -1, CU_TENSOR_MAP_SWIZZLE_128B, l2);
        run_tma("tile 8x16 C=64 (128 B rows)", buf, 64, W, H, NB / 2, 1, 16, 8, 16, 0, 8, 0, CU_TENSOR_MAP_SWIZZLE_128B, l2);
        run_tma("tile 8x16 C=16 (32 B rows)", buf, 16, W, H, NB, 1, 16, 8, 16, 0, 8, 0, CU_TENSOR_MAP_SWIZZLE_32B, l2);
        run_tma("tile 8x4quad C=16 (128 B rows)", buf, 16, W, H, NB, 4, 4, 8, 4, 0, 8, 0, CU_TENSOR_MAP_SWIZZLE_128B, l2);
        run_tma("patch 5x9 C=64 (head t1)", buf, 64, W / 2, H / 2, NB, 1, 9, 5, 8, -1, 4, -1, CU_TENSOR_MAP_SWIZZLE_128B, l2);
        run_tma("s2 patch 33x17pair C=16 (64 B)", buf, 16, W, H, NB, 2, 17, 33, 16, 0, 32, 0, CU_TENSOR_MAP_SWIZZLE_64B, l2);
        run_tma("s2 patch 33x9quad C=16 (128 B)", buf, 16, W, H, NB, 4, 9, 33, 8, 0, 32, 0, CU_TENSOR_MAP_SWIZZLE_128B, l2);
    }
    for (int win = 0; win < 2; ++win) {
        const int sbo = win ? 18 : 8;
        run_mma<16, 32>(sbo, win); run_mma<32, 32>(sbo, win); run_mma<16, 128>(sbo, win);
        run_mma<32, 64>(sbo, win); run_mma<64, 64>(sbo, win); run_mma<32, 128>(sbo, win);
        run_mma<64, 128>(sbo, win); run_mma<128, 128>(sbo, win); run_mma<256, 128>(sbo, win);
    }
    return 0;
}
