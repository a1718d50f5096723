// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16, M = 128, K = 16 per instruction).
//  1. layout: A[m][k] written with tcgen05.st.32x32b (lane m, 32-bit column k/2 holding elements k, k+1) -- D = A . I must return A
//  2. rate: cycles per MMA (N = 64) with A from TMEM vs A from shared memory (same B descriptor)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ukbb_cardiac_b200/csrc experiments/ts_probe.cu -o experiments/bin/ts_probe
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

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}



constexpr int A_COL = 256;          // TMEM columns of A: [256, 288); accumulators of the rate test: 4 x 64 columns from 0

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __nv_bfloat16* a_glob, float* out,
      long long* cyc, int iters) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t asm_ = base;                         // A in shared memory too: [128][64] 128 B swizzle (16 KB)
    const uint32_t bsm = base + 16384;                  // identity [64][64] (8 KB)
    const uint32_t bar = bsm + 8192, bar2 = bar + 8, slot = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 16384 + 8192);
        tma_load_2d(asm_, &map_a, bar, 0, 0);
        tma_load_2d(bsm, &map_b, bar, 0, 0);
    }
    // A row of this thread -> TMEM
    const int m = threadIdx.x;
    uint32_t v[32];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_glob + (size_t)m * 64);
    for (int i = 0; i < 32; ++i) v[i] = src[i];
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + A_COL, v);
    tmem_st_wait();
    mbar_wait(bar, 0);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    uint32_t ph = 0;
    // ---- 1. layout check: D[:, 0:64] = A(tmem) . I ; D[:, 64:128] = A(smem) . I
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) umma_ts(tmem, tmem + A_COL + 8 * k, make_smem_desc(bsm + k * 32, 128), idesc, k != 0);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + 64, make_smem_desc(asm_ + k * 32, 128), make_smem_desc(bsm + k * 32, 128), idesc, k != 0);
        umma_commit(bar2);
    }
    mbar_wait(bar2, ph); ph ^= 1;
    tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t d[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, d);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(size_t)m * 128 + c0 + i] = __uint_as_float(d[i]);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    // ---- 2. rate (warp-uniform issue path: descriptors live in uniform registers, one elected lane issues)
    for (int mode = 0; mode < 2; ++mode) {
        if (warp == 0) {
            const bool leader = elect_one();
            const uint32_t a_lo = ((asm_ & 0x3FFFF) >> 4) | (1u << 16), b_lo = ((bsm & 0x3FFFF) >> 4) | (1u << 16);
            constexpr uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                if (leader) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (mode == 0) {
                                asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
                                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(tmem + 64 * r),
                                             "r"(tmem + A_COL + 8 * k), "r"(b_lo + 2 * k), "r"(hi), "r"(idesc), "r"(1u)
                                             : "memory");
                            } else {
                                umma_bf16_lohi(tmem + 64 * r, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, 1u);
                            }
                        }
                }
                __syncwarp();
            }
            if (leader) umma_commit(bar2);
            __syncwarp();
            mbar_wait(bar2, ph);
            if (lane == 0) cyc[mode] = clock64() - t0;
        }
        ph ^= 1;
        tc_fence_before(); __syncthreads(); tc_fence_after();
    }
    // ---- 3. latency: n MMAs + commit -> barrier completion seen by the issuing warp (cycles)
    for (int mode = 0; mode < 2; ++mode)
        for (int n = 1; n <= 9; n += 4) {
            if (warp == 0) {
                const bool leader = elect_one();
                const uint32_t a_lo = ((asm_ & 0x3FFFF) >> 4) | (1u << 16), b_lo = ((bsm & 0x3FFFF) >> 4) | (1u << 16);
                constexpr uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
                long long best = 1ll << 60;
                for (int rep = 0; rep < 8; ++rep) {
                    const long long t0 = clock64();
                    if (leader) {
                        for (int k = 0; k < n; ++k) {
                            if (mode == 0)
                                asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
                                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(tmem), "r"(tmem + A_COL + 8 * (k & 3)),
                                             "r"(b_lo + 2 * (k & 3)), "r"(hi), "r"(idesc), "r"(1u)
                                             : "memory");
                            else umma_bf16_lohi(tmem, a_lo + 2 * (k & 3), hi, b_lo + 2 * (k & 3), hi, idesc, 1u);
                        }
                        umma_commit(bar2);
                    }
                    __syncwarp();
                    mbar_wait(bar2, ph); ph ^= 1;
                    const long long dt = clock64() - t0;
                    if (dt < best) best = dt;
                }
                if (lane == 0) cyc[2 + mode * 3 + n / 4] = best;
            } else {
                for (int rep = 0; rep < 8; ++rep) ph ^= 1;
            }
            tc_fence_before(); __syncthreads(); tc_fence_after();
        }
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    std::vector<__nv_bfloat16> ha(128 * 64), hb(64 * 64);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) ha[m * 64 + k] = __float2bfloat16((float)((m * 3 + k * 5) % 251));
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *da, *db; CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap ma, mb;
    { cuuint64_t d[2] = {64, 128}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, 128}; cuuint32_t e[2] = {1, 1};
      if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode a failed\n"); return 1; } }
    { cuuint64_t d[2] = {64, 64}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, 64}; cuuint32_t e[2] = {1, 1};
      if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode b failed\n"); return 1; } }
    float* dout; long long* dc; CK(cudaMalloc(&dout, 128 * 128 * 4)); CK(cudaMalloc(&dc, 64));
    const int smem = 16384 + 8192 + 1024 + 64, iters = 256;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<<<1, 128, smem>>>(ma, mb, da, dout, dc, iters);
    CK(cudaDeviceSynchronize());
    std::vector<float> ho(128 * 128); long long hc[8];
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc, dc, 64, cudaMemcpyDeviceToHost));
    int bad_ts = 0, bad_ss = 0;
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) {
        const float want = (float)((m * 3 + k * 5) % 251);
        if (ho[m * 128 + k] != want) ++bad_ts;
        if (ho[m * 128 + 64 + k] != want) ++bad_ss;
    }
    printf("A from TMEM: %d mismatches; A from shared memory: %d mismatches\n", bad_ts, bad_ss);
    if (bad_ts) { printf("row 1, TS:"); for (int k = 0; k < 16; ++k) printf(" %g", ho[128 + k]); printf("\nwant      :"); for (int k = 0; k < 16; ++k) printf(" %d", (3 + k * 5) % 251); printf("\n"); }
    printf("cycles per MMA (M=128, N=64, K=16): A from TMEM %.1f, A from shared memory %.1f\n", (double)hc[0] / (iters * 16), (double)hc[1] / (iters * 16));
    printf("latency issue -> commit seen (cycles): A from TMEM n=1 %lld, n=5 %lld, n=9 %lld; A from smem n=1 %lld, n=5 %lld, n=9 %lld\n", hc[2], hc[3], hc[4], hc[5], hc[6], hc[7]);
    return 0;
}
