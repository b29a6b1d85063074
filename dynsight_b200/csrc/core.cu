// core.cu -- error reporting and build identification of libdynsight_b200 (sm_100a).
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace dsg {

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

// fp64 FMA throughput probe: 16 independent chains per thread, `iters` rounds.
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(double* out, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = seed + (double)(threadIdx.x + k);
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keep the work alive
}

}  // namespace dsg

extern "C" int dsg_fp64_fma_probe(double* d_scratch, int blocks, int iters, void* stream,
                                  double* flops_out) {
    if (!d_scratch || blocks < 1 || iters < 1)
        return dsg::set_error(DSG_ERR_INVALID_ARGUMENT, "bad probe arguments");
    dsg::fp64_fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_scratch, iters, 1.0);
    DSG_LAUNCH_CHECK("fp64_fma_probe_kernel");
    if (flops_out) *flops_out = 2.0 * 16.0 * 256.0 * (double)blocks * (double)iters;
    return DSG_OK;
}

extern "C" const char* dsg_version(void) { return "dynsight_b200 0.1.0 (sm_100a)"; }

extern "C" const char* dsg_last_error(void) { return dsg::last_error_buffer(); }

extern "C" int dsg_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess)
        return dsg::set_error(DSG_ERR_CUDA, "cannot query the current CUDA device");
    return major * 10 + minor;
}

extern "C" long long dsg_launch_count(void) { return dsg::launches(); }
