// How fast can ONE warp per scheduler issue independent DMMA.8x8x4 (and DFMA in between)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_issue_probe dmma_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// NT independent accumulator tiles per trip, NF independent DFMA chains interleaved per DMMA
template <int NT, int NF>
__global__ void probe(double* out, int iters, double s, long long* cyc) {
    double c[2 * NT], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < 2 * NT; ++i) c[i] = i * 1e-3;
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x * 1e-3 + i;
    double a = s + threadIdx.x * 1e-9, b = 1.0 - s;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            dmma(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
            for (int j = 0; j < NF; ++j) f[j] = fma(f[j], a, b);
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < 2 * NT; ++i) r += c[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) r += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NT, int NF>
void run(int warps_per_smsp, double* d, long long* dc) {
    const int iters = 2000;
    probe<NT, NF><<<148, 128 * warps_per_smsp>>>(d, iters, 0.5, dc);
    cudaDeviceSynchronize();
    probe<NT, NF><<<148, 128 * warps_per_smsp>>>(d, iters, 0.5, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("tiles %2d  dfma/dmma %d  warps/SMSP %d : %7.2f cycles per DMMA per warp, %6.2f per SMSP\n", NT, NF,
           warps_per_smsp, (double)c / iters / NT, (double)c / iters / NT / warps_per_smsp);
}

int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    long long* dc; cudaMalloc(&dc, 8);
    for (int w = 1; w <= 4; w *= 2) {
        run<1, 0>(w, d, dc);
        run<2, 0>(w, d, dc);
        run<4, 0>(w, d, dc);
        run<14, 0>(w, d, dc);
        run<14, 2>(w, d, dc);
        run<14, 4>(w, d, dc);
        run<14, 8>(w, d, dc);
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
