// micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue/throughput on sm_100a, and LDS rate.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { x[i].x = fmaf(x[i].x, a2.x, b2.x); x[i].y = fmaf(x[i].y, a2.y, b2.y); }
                else x[i] = __ffma2_rn(x[i], a2, b2);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int mode = 0; mode < 2; mode++) for (int bpsm = 1; bpsm <= 2; bpsm++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * bpsm, 512>>>(out, iters, 1.0001f, 0.5f); else k<1><<<148 * bpsm, 512>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * bpsm * 512 * (double)iters * 4 * 8 * 2;
        printf("mode %s blocks/SM %d: %.3f ms  %.2f TFMA/s (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ", bpsm, ms, fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
