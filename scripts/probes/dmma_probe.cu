// Throughput probe: fp64 FMA pipe vs fp64 tensor MMA (mma.sync.m8n8k4.f64) vs both interleaved.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0 = DFMA only, 1 = DMMA only, 2 = both interleaved
__global__ void __launch_bounds__(256) probe(double* out, int iters, double s) {
    double acc[8], c[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
    double a = s + threadIdx.x * 1e-9, b = 1.0 - s;
    for (int it = 0; it < iters; ++it) {
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
        }
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += acc[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) r += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double* d, int blocks, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(d, iters, 0.5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(d, iters, 0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * 8;
    const double fma_flops = MODE != 1 ? warps * iters * 8.0 * 32 * 2 : 0;
    const double mma_flops = MODE != 0 ? warps * iters * 8.0 * 256 * 2 : 0;
    printf("%-12s %8.3f ms  DFMA %7.2f TFLOP/s  DMMA %7.2f TFLOP/s\n", name, ms,
           fma_flops / ms / 1e9, mma_flops / ms / 1e9);
}

int main() {
    double* d; cudaMalloc(&d, 148 * 8 * 256 * 8 * sizeof(double));
    const int blocks = 148 * 8, iters = 20000;
    run<0>("dfma", d, blocks, iters);
    run<1>("dmma", d, blocks, iters);
    run<2>("both", d, blocks, iters);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
