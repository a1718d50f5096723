// Probe 2: MN-major B operand taken straight from a TMA-written pixel-major patch.
//   D[128 x N] = A[128 x K] * B[K x N],  A K-major (selector matrix), B = patch rows [k][N channels]
//   (N contiguous = "MN-major"), SWIZZLE_128B (N = 64) and SWIZZLE_64B (N = 32); K = 16 per MMA,
//   two MMAs (K = 32) to check the K advance (+16 rows).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace ukbb::tc;

template <int NCH>
__global__ void __launch_bounds__(128, 1)
probe_mn(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int variant, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t a_sm = base;                 // A: [128][32] bf16 K-major SW64 (64 B rows) = 8 KB
    const uint32_t b_sm = base + 8192;          // B patch: [32 rows k][NCH*2 B]
    const uint32_t bar = b_sm + 8192, bar2 = bar + 8, slot = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 64); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 8192 + 32 * NCH * 2);
        tma_load_2d(a_sm, &map_a, bar, 0, 0);
        tma_load_2d(b_sm, &map_b, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (threadIdx.x == 0) {
        constexpr uint32_t RBB = NCH * 2;
        // instruction descriptor with b_major = MN (bit 16)
        const uint32_t idesc = make_idesc_bf16(128, NCH) | (1u << 16);
        const uint32_t a_hi = (uint32_t)((8 * 64) >> 4) | (1u << 14) | (4u << 29);
        const uint32_t a_lo = ((a_sm & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t layout = RBB == 128 ? 2u : 4u;
        // variant 0: SBO = 8 rows * row bytes (K-group stride), LBO = 1
        // variant 1: LBO = 8 rows * row bytes, SBO = 1  (in case the roles are swapped for MN-major)
        uint32_t b_hi, b_lo = ((b_sm & 0x3FFFF) >> 4);
        if (variant == 0) { b_hi = (uint32_t)((8 * RBB) >> 4) | (1u << 14) | (layout << 29); b_lo |= (1u << 16); }
        else { b_hi = 1u | (1u << 14) | (layout << 29); b_lo |= ((uint32_t)((8 * RBB) >> 4) << 16); }
        for (int k = 0; k < 2; ++k)
            umma_bf16_lohi(tmem, a_lo + ((k * 32) >> 4), a_hi, b_lo + ((k * 16 * RBB) >> 4), b_hi, idesc, k != 0);
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < NCH; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(size_t)r * NCH + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int NCH>
void run(EncodeTiledFn enc, int pattern) {
    // A[m][k] = 1 if k == m % 32 else 0  (row m selects source row m % 32);  plus A[m][(m+1)%32] = 0.5 for m >= 64
    std::vector<__nv_bfloat16> ha(128 * 32), hb(32 * NCH);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) {
        float v = (k == m % 32) ? 1.f : 0.f;
        if (m >= 64 && k == (m + 1) % 32) v = 0.5f;
        ha[m * 32 + k] = __float2bfloat16(v);
    }
    for (int k = 0; k < 32; ++k) for (int n = 0; n < NCH; ++n) hb[k * NCH + n] = __float2bfloat16((float)(k * 4) + (float)(n % 4) * 0.25f + (float)(n / 4) * 128.f >= 256.f ? (float)(k + n) : (float)(k * 4 + n % 4));
    // simpler exact pattern: B[k][n] = k + 32 * (n % 8)   (<= 255, exact in bf16), checked with n via second run
    // pattern 0: B[k][n] = k + 32 * (n % 8) (position inside a 16-byte chunk); pattern 1: k + 32 * (n / 8) (chunk order)
    auto bval = [&](int k, int n) { return (float)(k + 32 * (pattern == 0 ? (n % 8) : (n / 8))); };
    for (int k = 0; k < 32; ++k) for (int n = 0; n < NCH; ++n) hb[k * NCH + n] = __float2bfloat16(bval(k, n));
    __nv_bfloat16 *da, *db; CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap ma, mb;
    { cuuint64_t d[2] = {32, 128}; cuuint64_t s[1] = {64}; cuuint32_t b[2] = {32, 128}; cuuint32_t e[2] = {1, 1};
      CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc a %d\n", (int)r); exit(1); } }
    { cuuint64_t d[2] = {(cuuint64_t)NCH, 32}; cuuint64_t s[1] = {(cuuint64_t)NCH * 2}; cuuint32_t b[2] = {(cuuint32_t)NCH, 32}; cuuint32_t e[2] = {1, 1};
      CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, NCH == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc b %d\n", (int)r); exit(1); } }
    float* dout; CK(cudaMalloc(&dout, 128 * NCH * 4));
    const int smem = 8192 + 8192 + 1024 + 64;
    CK(cudaFuncSetAttribute(probe_mn<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int variant = 0; variant < 2; ++variant) {
        CK(cudaMemset(dout, 0, 128 * NCH * 4));
        probe_mn<NCH><<<1, 128, smem>>>(ma, mb, variant, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("NCH=%d variant %d: kernel error %s\n", NCH, variant, cudaGetErrorString(e)); exit(1); }
        std::vector<float> ho(128 * NCH);
        CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < NCH; ++n) {
            float exp = bval(m % 32, n);
            if (m >= 64) exp += 0.5f * bval((m + 1) % 32, n);
            if (ho[m * NCH + n] != exp) bad++;
        }
        printf("pattern %d NCH=%d variant %d (%s): %s, %d bad of %d; row0: ", pattern, NCH, variant, variant == 0 ? "SBO=8rows,LBO=1" : "LBO=8rows,SBO=1", bad ? "FAIL" : "OK", bad, 128 * NCH);
        for (int n = 0; n < 12; ++n) printf("%g ", ho[n]);
        printf("| row1: "); for (int n = 0; n < 6; ++n) printf("%g ", ho[NCH + n]);
        printf("| row17: "); for (int n = 0; n < 6; ++n) printf("%g ", ho[17 * NCH + n]);
        printf("\n");
    }
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    for (int pat = 0; pat < 2; ++pat) { run<64>((EncodeTiledFn)fn, pat); run<32>((EncodeTiledFn)fn, pat); }
    return 0;
}
