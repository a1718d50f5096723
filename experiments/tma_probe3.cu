// What bounds small TMA tensor loads at ~326 cycles per load (profiles/r1_rate_probe.log)?  Same 4 KB box
// (8 x 16 pixels x 16 channels, L2-resident source), depth-4 ring per issuer, with the loads issued by
//   mode 0: one lane;  mode 1: K lanes of ONE warp in one instruction;  mode 2: lane 0 of K different warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ukbb_cardiac_b200/csrc experiments/tma_probe3.cu -o experiments/bin/tma_probe3
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace ukbb::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int DEPTH = 4, BOX = 4096, ITERS = 400;

__global__ void __launch_bounds__(256, 1)
probe(const __grid_constant__ CUtensorMap map, int mode, int k, int prefetch, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 8 * DEPTH * BOX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int s = 0; s < 8 * DEPTH; ++s) mbar_init(bar + 8 * s, 1); fence_barrier_init(); if (prefetch) tma_prefetch_desc(&map); }
    __syncthreads();
    int me = -1;                         // issuer index
    if (mode == 0) me = threadIdx.x == 0 ? 0 : -1;
    if (mode == 1) me = (warp == 0 && lane < k) ? lane : -1;
    if (mode == 2) me = (lane == 0 && warp < k) ? warp : -1;
    const long long t0 = clock64();
    if (mode == 1 ? warp == 0 : me >= 0) {
        const int who = me < 0 ? 0 : me;
        for (int i = 0; i < ITERS; ++i) {
            const int s = i & (DEPTH - 1);
            const uint32_t b = bar + 8 * (who * DEPTH + s);
            if (me >= 0) {
                if (i >= DEPTH) mbar_wait(b, (uint32_t)((i >> 2) - 1) & 1u);
                mbar_arrive_expect_tx(b, BOX);
            }
            if (mode == 1) __syncwarp();
            if (me >= 0) tma_load_4d(base + (who * DEPTH + s) * BOX, &map, b, 0, ((i * 7 + who) % 12) * 16, ((i * 3 + blockIdx.x) % 26) * 8, (i + who) & 3);
            if (mode == 1) __syncwarp();
        }
        if (me >= 0)
            for (int i = ITERS; i < ITERS + DEPTH; ++i) mbar_wait(bar + 8 * (who * DEPTH + (i & (DEPTH - 1))), (uint32_t)((i >> 2) - 1) & 1u);
    }
    const long long t1 = clock64();
    if (me == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int C = 16, W = 192, H = 208, NB = 4;
    void* buf; CK(cudaMalloc(&buf, (size_t)NB * H * W * C * 2)); CK(cudaMemset(buf, 0, (size_t)NB * H * W * C * 2));
    cuuint64_t dims[4] = {C, W, H, NB}; cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {C, 16, 8, 1}, es[4] = {1, 1, 1, 1};
    CUtensorMap map;
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return 1; }
    long long* dout; CK(cudaMalloc(&dout, 148 * 8));
    const int smem = 8 * DEPTH * BOX + 1024 + 512;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int prefetch : {0, 1})
        for (int mode : {0, 1, 2})
            for (int k : {1, 2, 4, 5, 8}) {
                if (mode == 0 && k > 1) continue;
                probe<<<148, 256, smem>>>(map, mode, k, prefetch, dout);      // warm
                probe<<<148, 256, smem>>>(map, mode, k, prefetch, dout);
                CK(cudaDeviceSynchronize());
                long long h[148]; CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
                double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
                const int issuers = mode == 0 ? 1 : k;
                printf("prefetch %d mode %d issuers %d: %7.1f cycles per loop iteration, %7.1f cycles per load per SM\n", prefetch, mode, issuers,
                       avg / ITERS, avg / ITERS / issuers);
            }
    return 0;
}
