// micro-benchmark (sm_100a): do FFMA2 (fma.rn.f32x2), FFMA and LDS overlap, and at which rates?  Answers the question the
// window tile kernel raises (profiles/r2_window_kernel_log.md: its time follows the instruction count although neither the
// FMA pipe nor the LSU data pipe is above 60 %).       nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: FFMA only          (16 independent chains)
// MODE 1: FFMA2 only         (8 independent chains of packed pairs)
// MODE 2: FFMA2 + FFMA mixed (8 packed chains + 8 scalar chains, interleaved 1:1)
// (modes 3 / 4, FFMA2 beside LDS and LDS alone, are kept in the source but not run: the compiler drops loads whose value is not
//  consumed, so their first version measured nothing -- see profiles/r2_ubench_pipes.txt for what WAS measured)
// MODE 5: FFMA2 with a scalar broadcast weight (the window kernel's operand form) + FADD2
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, int iters, float a, float b) {
    __shared__ float sm[4352];
    for (int i = threadIdx.x; i < 4352; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    float2 x[8];
    float y[8];
    for (int i = 0; i < 8; i++) { x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i); y[i] = threadIdx.x * 0.003f + i; }
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    float acc = 0.f;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { x[i].x = fmaf(x[i].x, a2.x, b2.x); x[i].y = fmaf(x[i].y, a2.y, b2.y); }
                if (MODE == 1 || MODE == 2 || MODE == 3) x[i] = __ffma2_rn(x[i], a2, b2);
                if (MODE == 2) y[i] = fmaf(y[i], a, b);
                if (MODE == 5) { x[i] = (i & 3) == 3 ? __fadd2_rn(x[i], b2) : __ffma2_rn(make_float2(a, a), x[i], b2); }
                if ((MODE == 3 && (i % 3) == 0) || MODE == 4) {
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr + (unsigned)((u * 8 + i) * 512)));
                    acc += (MODE == 4) ? v : 0.f;
                    if (MODE == 3) y[i & 7] = v;   // consumed outside the loop only
                }
            }
        }
    }
    float s = acc;
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *out, int ctas_per_sm, double fma_per_thread_iter, double lds_per_thread_iter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192;
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<148 * ctas_per_sm, 128>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double thr = 148.0 * ctas_per_sm * 128, clk = ms * 1e-3 * 1.965e9;
    printf("%-28s warps/SMSP %d: %.3f ms | %.1f FMA/clk/SM | %.2f LDS warp-instr/clk/SM\n", name, ctas_per_sm, ms,
           thr * iters * fma_per_thread_iter / clk / 148, thr / 32 * iters * lds_per_thread_iter / clk / 148);
}

int main() {
    float *out; cudaMalloc(&out, 148 * 16 * 128 * 4);
    for (int c = 4; c <= 16; c *= 2) {
        run<0>("FFMA", out, c, 64, 0);
        run<1>("FFMA2", out, c, 64, 0);
        run<2>("FFMA2 + FFMA 1:1", out, c, 96, 0);
        run<5>("FFMA2 bcast w + FADD2 (3:1)", out, c, 64, 0);
    }
    return 0;
}
