// common.cuh -- shared helpers of libdynsight_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/dynsight_b200.h"

namespace dsg {

// ---- error reporting (thread-local message, no exceptions across the C ABI) ----------------
char* last_error_buffer();
int set_error(int code, const char* fmt, ...);

#define DSG_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::dsg::set_error(DSG_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__,  \
                                    __LINE__, cudaGetErrorString(_e));                        \
    } while (0)

void count_launch();

#define DSG_LAUNCH_CHECK(name)                                                                \
    do {                                                                                      \
        ::dsg::count_launch();                                                                \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return ::dsg::set_error(DSG_ERR_CUDA, "launch of %s failed at %s:%d: %s", name,   \
                                    __FILE__, __LINE__, cudaGetErrorString(_e));              \
    } while (0)

// ---- reference arithmetic shared by host and device ---------------------------------------

// Python / numba float floor division a // b: exact floor of the true quotient (via fmod).
// Used for ncell_k = max(3, int(box_k // r_cut)), lens.py:43-45.
__host__ __device__ inline double py_floordiv(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
    if (div == 0.0) return copysign(0.0, a / b);
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
}

// int(x / box * ncell) % ncell with truncation toward zero then Python modulo (lens.py:52-54).
__host__ __device__ inline int cell_of(double x, double box, int ncell) {
    double v = x / box * (double)ncell;
    long long t = (long long)v;  // truncation toward zero, like numba's int()
    // an atom inside the box (almost every atom) needs no modulo: no 64-bit division on its path
    if ((unsigned long long)t < (unsigned long long)ncell) return (int)t;
    long long r = t % ncell;
    if (r < 0) r += ncell;
    return (int)r;
}

// LENS of one centre from the two row lengths and the size of their intersection (lens.py:167-188)
__host__ __device__ inline double lens_value(int a, int b, int inter) {
    const int denom = a + b;                 // lens.py:170
    if (denom <= 0) return 0.0;              // lens.py:171-173
    const int numer = denom - 2 * inter;     // lens.py:187
    return (double)numer / (double)denom;    // lens.py:188 (IEEE float64 division)
}

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// timeSOAP of one pair from the squared norms and the dot product of the two SOAP vectors
// (timesoap.py:54-58 + :118-127).  A zero-norm vector normalises to the zero vector => cos = 0.
__device__ __forceinline__ double soap_dist_from_sums(double n1sq, double n2sq, double dot) {
    double c = 0.0;
    if (n1sq > 0.0 && n2sq > 0.0) c = dot / (sqrt(n1sq) * sqrt(n2sq));
    c = fmin(1.0, fmax(-1.0, c));
    return sqrt(fmax(0.0, 2.0 - 2.0 * c));
}
#endif

__host__ __device__ inline int wrap_cell(int c, int n) {
    c %= n;
    return c < 0 ? c + n : c;
}

// Exclusive scan of n_seg uniform segments of int32 (neighbors.cu).  out has seg_len+1 entries per
// segment (prefix + total); tile_tmp needs n_seg*ceil(seg_len/4096) ints; seg_totals (int64, one
// per segment) may be NULL.
int run_scan(const int* in, long long seg_len, int n_seg, int* out, int* tile_tmp,
             long long* seg_totals, cudaStream_t st);
constexpr int kScanTileItems = 4096;

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

}  // namespace dsg
