// How do FP64 and integer/ALU instructions share the issue port on B200?
// Each iteration: 8 independent DFMA + R*8 independent integer ops (LOP3/IADD3 on 32-bit regs).
#include <cstdio>
#include <cuda_runtime.h>
template <int R>
__global__ void __launch_bounds__(256) k(double* out, unsigned* iout, double b_in, double c_in, int iters, unsigned seed) {
    double a[8]; unsigned u[8];
    const double t = threadIdx.x * 1e-9;
    double c = c_in + t;
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = 1.0 + j * 0.125 + t; u[j] = seed + j * 977u + threadIdx.x; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = fma(a[j], b_in, c);
#pragma unroll
            for (int r = 0; r < R; ++r) u[j] = (u[j] ^ (u[j] >> 3)) + (unsigned)(0x9e3779b9u + r);
        }
    }
    double s = 0; unsigned q = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += a[j]; q ^= u[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = q;
}
template <int R>
void run(double* d, unsigned* di, int sms) {
    const int iters = 2048, blocks = sms * 8;
    k<R><<<blocks, 256>>>(d, di, 1.0000001, 1e-9, 16, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<R><<<blocks, 256>>>(d, di, 1.0000001, 1e-9, iters, 7); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double dfma = (double)blocks * 256 * iters * 8;
    printf("8 DFMA + %2d x (SHF,LOP,IADD) per iter: %8.3f ms  DFMA %8.2f Ginst/s (%.0f%% of 18.4 peak)\n", 8 * R, best,
           dfma / best / 1e6, 100 * dfma / best / 1e6 / 18400.0);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; unsigned* di;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8); cudaMalloc(&di, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    run<0>(d, di, p.multiProcessorCount); run<1>(d, di, p.multiProcessorCount); run<2>(d, di, p.multiProcessorCount);
    run<3>(d, di, p.multiProcessorCount); run<4>(d, di, p.multiProcessorCount);
    return 0;
}
