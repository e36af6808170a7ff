// FP64 FMA peak microbenchmark (SURVEY 8d asks for a measured FP64 denominator): 8 independent DFMA chains per
// thread, 1024 threads x (8 x SMs) CTAs. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, blocks = sms * 2, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    dfma_kernel<<<blocks, threads>>>(out, 100, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double fmas = (double)threads * blocks * iters * 8.0;
    printf("{\"sms\": %d, \"ms\": %.3f, \"fp64_tflops\": %.2f, \"dfma_per_clk_per_sm_at_1965MHz\": %.1f}\n", sms, best,
           2.0 * fmas / (best * 1e-3) / 1e12, fmas / (best * 1e-3) / sms / 1.965e9);
    return 0;
}
