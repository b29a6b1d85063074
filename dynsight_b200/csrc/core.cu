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

}  // namespace dsg

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
