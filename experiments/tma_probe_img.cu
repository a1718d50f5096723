// Which FP32 image boxes can cp.async.bulk.tensor.3d load (SWIZZLE_NONE)?  One config per process
// (an illegal instruction kills the context).  usage: tma_probe_img boxw boxh x y [align]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ukbb_cardiac_b200/csrc experiments/tma_probe_img.cu -o experiments/bin/tma_probe_img
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
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int n, int bytes, int off, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = ((smem_u32(raw) + 1023u) & ~1023u) + off;
    const uint32_t bar = ((smem_u32(raw) + 1023u) & ~1023u) + 32768;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(base),
                     "l"(&map), "r"(bar), "r"(x), "r"(y), "r"(n) : "memory");
    }
    mbar_wait(bar, 0);
    const float* s = reinterpret_cast<const float*>(raw + (base - smem_u32(raw)));
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
    const int bw = atoi(argv[1]), bh = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]), off = argc > 5 ? atoi(argv[5]) : 0;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int W = 96, H = 64, NB = 3;
    std::vector<float> h((size_t)NB * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100000) + 1.f;
    float* buf; CK(cudaMalloc(&buf, h.size() * 4)); CK(cudaMemcpy(buf, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cuuint64_t dims[3] = {W, H, NB}; cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUtensorMap map;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("box %dx%d: encode failed %d\n", bw, bh, (int)r); return 1; }
    float* dout; CK(cudaMalloc(&dout, 65536));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
    probe<<<1, 128, 40000>>>(map, x, y, 1, bw * bh * 4, off, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("box %dx%d at (%d,%d) off %d: %s\n", bw, bh, x, y, off, cudaGetErrorString(e)); return 1; }
    std::vector<float> o(bw * bh); CK(cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int j = 0; j < bh; ++j)
        for (int i = 0; i < bw; ++i) {
            const int xx = x + i, yy = y + j;
            const float want = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? h[((size_t)1 * H + yy) * W + xx] : 0.f;
            if (o[j * bw + i] != want) ++bad;
        }
    printf("box %dx%d at (%d,%d) off %d: OK, %d mismatches\n", bw, bh, x, y, off, bad);
    return 0;
}
