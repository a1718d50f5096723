// Issue rate of FP32 FMA forms on one SM sub-partition (cycles per warp instruction):
//   0: FFMA 3-register   1: FFMA with a constant-bank operand   2: FFMA2 (fma.rn.f32x2) 3-register
//   3: FFMA2 with a scalar-broadcast operand ({v, v})   4: FFMA2 {v, v} x uniform-register weights   5: FFMA x uniform-register weights
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 experiments/ffma_probe.cu -o experiments/bin/ffma_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
struct W { float w[32]; };
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE>
__global__ void probe(const __grid_constant__ W cw, const float* in, float* out, long long* cyc, int iters) {
    float v = in[threadIdx.x], u = in[threadIdx.x + 1024];
    float acc[32];
    uint64_t acc2[16];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = in[j];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc2[j] = pk(in[j], in[j + 16]);
    float wr[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) wr[j] = in[64 + j + (MODE == 4 || MODE == 5 ? 0 : threadIdx.x)];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(v, cw.w[j], acc[j]);
        } else if (MODE == 2) {
            const uint64_t a = pk(v, u);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc2[j] = ffma2(a, pk(wr[2 * j], wr[2 * j + 1]), acc2[j]);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc2[j] = ffma2(a, pk(wr[2 * j], wr[2 * j + 1]), acc2[j]);
        } else if (MODE == 5) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
        } else {
            const uint64_t a = pk(v, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc2[j] = ffma2(a, pk(wr[2 * j], wr[2 * j + 1]), acc2[j]);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc2[j] = ffma2(a, pk(wr[2 * j], wr[2 * j + 1]), acc2[j]);
        }
        v += 1e-7f;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += acc[j];
#pragma unroll
    for (int j = 0; j < 16; ++j) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[j])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 8192 * 4); cudaMemset(in, 0, 8192 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    W cw; for (int j = 0; j < 32; ++j) cw.w[j] = 0.5f;
    const int iters = 2000;
    for (int mode = 0; mode < 6; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<148, warps * 32>>>(cw, in, out, cyc, iters);
                if (mode == 1) probe<1><<<148, warps * 32>>>(cw, in, out, cyc, iters);
                if (mode == 2) probe<2><<<148, warps * 32>>>(cw, in, out, cyc, iters);
                if (mode == 3) probe<3><<<148, warps * 32>>>(cw, in, out, cyc, iters);
                if (mode == 4) probe<4><<<148, warps * 32>>>(cw, in, out, cyc, iters);
                if (mode == 5) probe<5><<<148, warps * 32>>>(cw, in, out, cyc, iters);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
            const double winst_per_smsp = (double)iters * 32 * (warps / 4.0);     // warp instructions per sub-partition
            printf("mode %d warps/SM %2d: %.2f cycles per warp instruction per SMSP, %.1f FMA/clk/SM\n", mode, warps, avg / winst_per_smsp,
                   (double)iters * 32 * warps * 32 * (mode >= 2 ? 2 : 1) / avg);
        }
    return 0;
}
