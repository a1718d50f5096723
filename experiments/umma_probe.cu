// Probe: can a UMMA K-major swizzled A descriptor start at an arbitrary 128B/64B/32B-row offset
// inside a TMA-written patch, with an SBO that is not 8 rows?  (halo reuse for 3x3 convs:
// one patch load, 9 shifted windows).  D = A_window * I  -> D rows reveal which rows were read.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ukbb_cardiac_b200/csrc experiments/umma_probe.cu -o experiments/bin/umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace ukbb::tc;

struct Case { int j, sbo_rows, base_off; };

template <int CC>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_patch, const __grid_constant__ CUtensorMap map_b, int R,
             const Case* cases, int ncases, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t patch = base;                        // R rows x CC*2 bytes
    const uint32_t bsm = base + 256 * CC * 2;           // identity CC x CC
    const uint32_t bar = bsm + 8192;
    const uint32_t bar2 = bar + 8;
    const uint32_t slot = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 64); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, R * CC * 2 + CC * CC * 2);
        tma_load_2d(patch, &map_patch, bar, 0, 0);
        tma_load_2d(bsm, &map_b, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t ph = 0;
    for (int ci = 0; ci < ncases; ++ci) {
        const Case c = cases[ci];
        if (threadIdx.x == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, CC);
#pragma unroll
            for (int k = 0; k < CC / 16; ++k) {
                uint64_t ad = make_smem_desc(patch + c.j * CC * 2 + k * 32, CC * 2);
                ad &= ~(0x3FFFull << 32);
                ad |= (uint64_t)((c.sbo_rows * CC * 2) >> 4) << 32;
                ad |= (uint64_t)(c.base_off & 7) << 49;
                const uint64_t bd = make_smem_desc(bsm + k * 32, CC * 2);
                umma_bf16(tmem, ad, bd, idesc, k != 0);
            }
            umma_commit(bar2);
        }
        mbar_wait(bar2, ph); ph ^= 1;
        tc_fence_after();
        const int r = warp * 32 + lane;
        for (int c0 = 0; c0 < CC; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int i = 0; i < 16; ++i) out[((size_t)ci * 128 + r) * CC + c0 + i] = __uint_as_float(v[i]);
        }
        tc_fence_before(); __syncthreads(); tc_fence_after();
    }
    if (warp == 1) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int CC>
void run(EncodeTiledFn enc) {
    const int R = 200;
    std::vector<__nv_bfloat16> hp(R * CC), hb(CC * CC);
    for (int r = 0; r < R; ++r) for (int c = 0; c < CC; ++c) hp[r * CC + c] = __float2bfloat16((c & 1) ? (float)c : (float)r);
    for (int n = 0; n < CC; ++n) for (int k = 0; k < CC; ++k) hb[n * CC + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dp, *db; CK(cudaMalloc(&dp, hp.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2));
    CK(cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMapSwizzle sw = CC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUtensorMap mp, mb;
    { cuuint64_t d[2] = {(cuuint64_t)CC, (cuuint64_t)R}; cuuint64_t s[1] = {(cuuint64_t)CC * 2}; cuuint32_t b[2] = {(cuuint32_t)CC, (cuuint32_t)R}; cuuint32_t e[2] = {1, 1};
      CUresult r = enc(&mp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dp, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encode patch failed %d\n", (int)r); exit(1); } }
    { cuuint64_t d[2] = {(cuuint64_t)CC, (cuuint64_t)CC}; cuuint64_t s[1] = {(cuuint64_t)CC * 2}; cuuint32_t b[2] = {(cuuint32_t)CC, (cuuint32_t)CC}; cuuint32_t e[2] = {1, 1};
      CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encode b failed %d\n", (int)r); exit(1); } }
    std::vector<Case> cases;
    for (int sbo : {8, 10, 18}) for (int j = 0; j < 12; ++j) for (int bo : {0, -1}) {
        if (bo == -1 && (j & 7) == 0) continue;
        cases.push_back({j, sbo, bo == 0 ? 0 : (j & 7)});
    }
    Case* dc; CK(cudaMalloc(&dc, cases.size() * sizeof(Case))); CK(cudaMemcpy(dc, cases.data(), cases.size() * sizeof(Case), cudaMemcpyHostToDevice));
    float* dout; CK(cudaMalloc(&dout, cases.size() * 128 * CC * 4));
    const int smem = 256 * CC * 2 + 8192 + 1024 + 64;
    CK(cudaFuncSetAttribute(probe_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<CC><<<1, 128, smem>>>(mp, mb, R, dc, (int)cases.size(), dout);
    CK(cudaDeviceSynchronize());
    std::vector<float> ho(cases.size() * 128 * CC);
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
    printf("=== CC=%d (row %d bytes) ===\n", CC, CC * 2);
    for (size_t ci = 0; ci < cases.size(); ++ci) {
        const Case c = cases[ci];
        int bad_rows = 0, bad_cols = 0;
        for (int m = 0; m < 128; ++m) {
            const int expect = (m / 8) * c.sbo_rows + c.j + m % 8;
            if (expect >= R) continue;
            for (int col = 0; col < CC; ++col) {
                const float v = ho[(ci * 128 + m) * CC + col];
                if (col & 1) { if (v != (float)col) bad_cols++; } else if (v != (float)expect) bad_rows++;
            }
        }
        printf("j=%2d sbo_rows=%2d base_off=%d : %s (bad row-ids %d, bad col-ids %d)", c.j, c.sbo_rows, c.base_off,
               (bad_rows == 0 && bad_cols == 0) ? "OK  " : "FAIL", bad_rows, bad_cols);
        if (bad_rows || bad_cols) {
            printf("  rows0-9 got:");
            for (int m = 0; m < 10; ++m) printf(" %g", ho[(ci * 128 + m) * CC + 0]);
            printf(" | row0 cols:");
            for (int col = 0; col < 16 && col < CC; ++col) printf(" %g", ho[(ci * 128 + 0) * CC + col]);
        }
        printf("\n");
    }
    cudaFree(dp); cudaFree(db); cudaFree(dc); cudaFree(dout);
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    run<64>(enc); run<32>(enc); run<16>(enc);
    return 0;
}
