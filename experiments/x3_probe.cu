// Probe for the split-operand ("x3") tensor-core mode (round 2):
//   (1) does tcgen05.mma kind::f16 honour FP16 SUBNORMAL operands (the lo parts of weights ~2^-5 are ~2^-17)?
//   (2) may the A and B formats of one kind::f16 instruction differ (a_format = F16, b_format = BF16)?
//   (3) accuracy of D = Ahi.Bhi + Alo.Bhi + Ahi.Blo against a float64 product of the FP32 operands, for
//       FP16 and BF16 pieces, with activation-like A (|a| ~ 1) and weight-like B (|b| ~ 0.05).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ukbb_cardiac_b200/csrc experiments/x3_probe.cu -o experiments/bin/x3_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace ukbb::tc;

// operands: 4 tiles in global memory, each [rows][64] 16-bit: A_hi, A_lo (128 rows), B_hi, B_lo (64 rows)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, uint32_t idesc_hh, uint32_t idesc_lh,
             uint32_t idesc_hl, int terms, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768, b_lo = base + 32768 + 8192;
    const uint32_t bar = base + 49152, bar2 = bar + 8, slot = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 64); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 2 * 16384 + 2 * 8192);
        tma_load_2d(a_hi, &map_a, bar, 0, 0);
        tma_load_2d(a_lo, &map_a, bar, 0, 128);
        tma_load_2d(b_hi, &map_b, bar, 0, 0);
        tma_load_2d(b_lo, &map_b, bar, 0, 64);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            umma_bf16(tmem, make_smem_desc(a_hi + k * 32, 128), make_smem_desc(b_hi + k * 32, 128), idesc_hh, k != 0);
            if (terms >= 2) umma_bf16(tmem, make_smem_desc(a_lo + k * 32, 128), make_smem_desc(b_hi + k * 32, 128), idesc_lh, 1);
            if (terms >= 3) umma_bf16(tmem, make_smem_desc(a_hi + k * 32, 128), make_smem_desc(b_lo + k * 32, 128), idesc_hl, 1);
        }
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(size_t)r * 64 + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static uint16_t enc16(float v, bool f16, float* back) {
    uint16_t u;
    if (f16) { const __half h = __float2half_rn(v); memcpy(&u, &h, 2); *back = __half2float(h); }
    else { const __nv_bfloat16 h = __float2bfloat16(v); memcpy(&u, &h, 2); *back = __bfloat162float(h); }
    return u;
}

static EncodeTiledFn g_enc;

// hi_f16 / lo_f16: format of the hi and lo pieces (true = FP16, false = BF16)
static void run(const char* name, bool hi_f16, bool lo_f16, int terms, float a_scale, float b_scale) {
    std::vector<float> A(128 * 64), B(64 * 64);
    srand(7);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& v : A) v = fabsf(rnd()) * a_scale * (rand() % 8 == 0 ? 0.01f : 1.f);     // post-ReLU-like, some small values
    for (auto& v : B) v = rnd() * b_scale;
    std::vector<uint16_t> a16(2 * 128 * 64), b16(2 * 64 * 64);
    std::vector<double> Aeff(128 * 64), Beff(64 * 64);
    int sub_a = 0, sub_b = 0;
    for (int i = 0; i < 128 * 64; ++i) {
        float h, l;
        a16[i] = enc16(A[i], hi_f16, &h);
        a16[128 * 64 + i] = enc16(A[i] - h, lo_f16, &l);
        Aeff[i] = (double)h + (terms >= 2 ? (double)l : 0.0);
        if (lo_f16 && l != 0.f && fabsf(l) < 6.1035e-5f) ++sub_a;
    }
    for (int i = 0; i < 64 * 64; ++i) {
        float h, l;
        b16[i] = enc16(B[i], hi_f16, &h);
        b16[64 * 64 + i] = enc16(B[i] - h, lo_f16, &l);
        Beff[i] = (double)h + (terms >= 3 ? (double)l : 0.0);
        if (lo_f16 && l != 0.f && fabsf(l) < 6.1035e-5f) ++sub_b;
    }
    uint16_t *da, *db; CK(cudaMalloc(&da, a16.size() * 2)); CK(cudaMalloc(&db, b16.size() * 2));
    CK(cudaMemcpy(da, a16.data(), a16.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b16.data(), b16.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap ma, mb;
    { cuuint64_t d[2] = {64, 256}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, 128}; cuuint32_t e[2] = {1, 1};
      CUresult r = g_enc(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, da, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encode a failed %d\n", (int)r); exit(1); } }
    { cuuint64_t d[2] = {64, 128}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, 64}; cuuint32_t e[2] = {1, 1};
      CUresult r = g_enc(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, db, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encode b failed %d\n", (int)r); exit(1); } }
    // kind::f16 instruction descriptor: a_format bits [7,10), b_format bits [10,13): 0 = F16, 1 = BF16
    auto idesc = [](bool a_f16, bool b_f16) { return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
    float* dout; CK(cudaMalloc(&dout, 128 * 64 * 4));
    const int smem = 49152 + 64 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(ma, mb, idesc(hi_f16, hi_f16), idesc(lo_f16, hi_f16), idesc(hi_f16, lo_f16), terms, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s : kernel failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> D(128 * 64);
    CK(cudaMemcpy(D.data(), dout, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_eff = 0, max_true = 0, ref_rms = 0, err_true_rms = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double se = 0, st = 0;
            for (int k = 0; k < 64; ++k) {
                se += Aeff[m * 64 + k] * Beff[n * 64 + k];
                st += (double)A[m * 64 + k] * (double)B[n * 64 + k];
            }
            max_eff = fmax(max_eff, fabs(D[m * 64 + n] - se));
            max_true = fmax(max_true, fabs(D[m * 64 + n] - st));
            ref_rms += st * st; err_true_rms += (D[m * 64 + n] - st) * (D[m * 64 + n] - st);
        }
    ref_rms = sqrt(ref_rms / (128 * 64)); err_true_rms = sqrt(err_true_rms / (128 * 64));
    printf("%-34s : subnormal lo pieces A %5d B %5d | max |D - pieces product| %.3e | vs FP32 operands: max %.3e rms %.3e (ref rms %.3e) -> rel rms %.3e\n",
           name, sub_a, sub_b, max_eff, max_true, err_true_rms, ref_rms, err_true_rms / ref_rms);
    cudaFree(da); cudaFree(db); cudaFree(dout);
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_enc = (EncodeTiledFn)fn;
    run("fp16 1 term", true, true, 1, 1.f, 0.05f);
    run("bf16 1 term", false, false, 1, 1.f, 0.05f);
    run("fp16 hi + fp16 lo, 3 terms", true, true, 3, 1.f, 0.05f);
    run("bf16 hi + bf16 lo, 3 terms", false, false, 3, 1.f, 0.05f);
    run("fp16 3 terms, tiny weights 1e-3", true, true, 3, 1.f, 0.001f);
    run("fp16 3 terms, tiny acts 1e-3", true, true, 3, 0.001f, 0.05f);
    run("fp16 hi + bf16 lo (mixed), 3 terms", true, false, 3, 1.f, 0.05f);
    return 0;
}
