// Micro-benchmark: fp32 FMA issue rate per SM on B200 -- scalar FFMA vs packed
// fma.rn.f32x2 (FFMA2).  Decides how the Pearson main loop should be written.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, int iters) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float x0 = a, x1 = b;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], (r & 1) ? x0 : x1, (r & 1) ? x1 : x0);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long d, A, B, Cc;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(acc[i]), "f"(acc[i + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"((r & 1) ? x0 : x1), "f"((r & 1) ? x0 : x1));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(Cc) : "f"((r & 1) ? x1 : x0), "f"((r & 1) ? x1 : x0));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(A), "l"(B), "l"(Cc));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name) {
    float *out;
    int grid = 148 * 8, iters = 20000;
    cudaMalloc(&out, grid * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 0.999f, 0.001f, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 0.999f, 0.001f, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)grid * 256 * iters * 8.0 * 16.0;
    printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s), %.1f FMA/clk/SM @1.965GHz\n", name, ms,
           fma / ms * 1e-9, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main() {
    run<0>("FFMA  (scalar)");
    run<1>("FFMA2 (f32x2) ");
    run<0>("FFMA  (scalar)");
    run<1>("FFMA2 (f32x2) ");
    return 0;
}
