// FP64 pipe micro-benchmark (B200): DFMA / DMUL / DADD issue rate for different operand mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double b_in, double c_in, int iters) {
    double a[8];
    const double t = threadIdx.x * 1e-9;
    double b = b_in + t, c = c_in + t;          // per-thread registers
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + j * 0.125 + t;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = fma(a[j], b, c);            // 3 distinct register operands
            if (MODE == 1) a[j] = fma(a[j], b_in, c);         // one uniform/constant operand
            if (MODE == 2) a[j] = fma(a[j], a[j], c);         // 2 distinct registers
            if (MODE == 3) a[j] = a[j] * b;                   // DMUL
            if (MODE == 4) a[j] = a[j] + c;                   // DADD
            if (MODE == 5) a[j] = fma(a[j], b_in, c_in);      // two uniform operands
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double* d, int sms) {
    const int iters = 4096, blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(d, 1.0000001, 1e-9, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(d, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double inst = (double)blocks * 256 * iters * 8;
    printf("%-44s %8.3f ms  %8.2f Ginst/s  (as FMA: %6.2f TFLOP/s)\n", name, best, inst / best / 1e6, 2 * inst / best / 1e9);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz -> nominal 64 DFMA/clk/SM = %.2f Ginst/s\n", p.name, p.multiProcessorCount,
           clk / 1000, p.multiProcessorCount * 64.0 * clk / 1e6);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(double));
    run<0>("DFMA a=fma(a,b,c)  3 distinct registers", d, p.multiProcessorCount);
    run<1>("DFMA a=fma(a,U,c)  1 uniform operand", d, p.multiProcessorCount);
    run<2>("DFMA a=fma(a,a,c)  2 distinct registers", d, p.multiProcessorCount);
    run<5>("DFMA a=fma(a,U,U)  2 uniform operands", d, p.multiProcessorCount);
    run<3>("DMUL a=a*b", d, p.multiProcessorCount);
    run<4>("DADD a=a+c", d, p.multiProcessorCount);
    cudaFree(d);
    return 0;
}
