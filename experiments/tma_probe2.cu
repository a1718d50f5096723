// TMA issue/latency probe 2: why does a stream of box loads cost ~905 cycles per load per SM?
// Variants: wait style (try_wait loop / test_wait spin), producer-consumer split across warps,
// descriptor prefetch, several loads per barrier.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ukbb_cardiac_b200/csrc experiments/tma_probe2.cu -o experiments/bin/tma_probe2
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

__device__ __forceinline__ void spin_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

struct P { int iters, depth, split, bytes, stage_bytes, rows_per_load, mode, prefetch, tiles_x, tiles_y, nb, ystep; };

// mode 0: one thread issues + try_wait on full;  1: one thread, test_wait spin;
// mode 2: producer (warp 0) waits on empty (arrived by consumer thread), consumer (warp 1) try_waits on full
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map, const P p, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + p.depth * p.stage_bytes;          // full[depth], empty[depth]
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * p.depth; ++s) mbar_init(bar + 8 * s, 1);
        fence_barrier_init();
        if (p.prefetch) tma_prefetch_desc(&map);
    }
    __syncthreads();
    const int per = p.tiles_x * p.tiles_y;
    auto issue = [&](int i, int s) {
        long long t = (long long)blockIdx.x + (long long)i * gridDim.x;
        const int n = (int)((t / per) % p.nb);
        const int t2 = (int)(t % per);
        const int tx = t2 % p.tiles_x, ty = t2 / p.tiles_x;
        mbar_arrive_expect_tx(bar + 8 * s, p.bytes);
        for (int j = 0; j < p.split; ++j)
            tma_load_4d(base + s * p.stage_bytes + j * (p.bytes / p.split), &map, bar + 8 * s, 0, tx * 16, ty * p.ystep + j * p.rows_per_load, n);
    };
    if (p.mode < 2) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            for (int i = 0; i < p.iters + p.depth; ++i) {
                const int s = i % p.depth;
                if (i >= p.depth) { if (p.mode == 0) mbar_wait(bar + 8 * s, (uint32_t)((i / p.depth) - 1) & 1u); else spin_wait(bar + 8 * s, (uint32_t)((i / p.depth) - 1) & 1u); }
                if (i < p.iters) issue(i, s);
            }
            out[blockIdx.x] = clock64() - t0;
        }
    } else {
        if (threadIdx.x == 0) {
            for (int i = 0; i < p.iters; ++i) {
                const int s = i % p.depth;
                mbar_wait(bar + 8 * (p.depth + s), ((uint32_t)(i / p.depth) & 1u) ^ 1u);
                issue(i, s);
            }
        } else if (threadIdx.x == 32) {
            const long long t0 = clock64();
            for (int i = 0; i < p.iters; ++i) {
                const int s = i % p.depth;
                mbar_wait(bar + 8 * s, (uint32_t)(i / p.depth) & 1u);
                mbar_arrive(bar + 8 * (p.depth + s));
            }
            out[blockIdx.x] = clock64() - t0;
        }
    }
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int W = 192, H = 208, NB = 4, C = 64;
    void* buf; CK(cudaMalloc(&buf, (size_t)NB * W * H * C * 2)); CK(cudaMemset(buf, 0, (size_t)NB * W * H * C * 2));
    long long* dout; CK(cudaMalloc(&dout, 148 * 8));
    for (int grid : {148, 1})
    for (int by : {8, 2}) for (int split : {1, 2, 4}) {
        if (by % split) continue;
        if (grid == 1 && split > 1) continue;
        const int rows = by / split;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)C, 16, (cuuint32_t)rows, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUtensorMap map;
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode failed %d\n", (int)r); return 1; }
        for (int mode = 0; mode < 3; ++mode) for (int prefetch = 0; prefetch < 2; ++prefetch) for (int depth : {2, 4, 8}) {
            if (grid == 1 && (prefetch || depth == 2)) continue;
            P p; p.iters = 400; p.depth = depth; p.split = split; p.bytes = C * 2 * 16 * by; p.stage_bytes = (p.bytes + 1023) / 1024 * 1024;
            p.rows_per_load = rows; p.mode = mode; p.prefetch = prefetch; p.tiles_x = W / 16; p.tiles_y = H / by; p.nb = NB; p.ystep = by;
            const int smem = depth * p.stage_bytes + 1024 + 256;
            CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            probe<<<grid, 128, smem>>>(map, p, dout);
            probe<<<grid, 128, smem>>>(map, p, dout);
            CK(cudaDeviceSynchronize());
            long long h[148]; CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
            double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
            printf("grid %3d box %5d B  split %d  mode %d  prefetch %d  depth %d : %7.1f cyc/stage  %6.1f B/cyc/SM\n", grid, p.bytes, split, mode,
                   prefetch, depth, avg / p.iters, (double)p.bytes * p.iters / avg);
        }
    }
    return 0;
}
