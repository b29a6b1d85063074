// tcf.cu -- self time-correlation function of a per-particle time series (sm_100a).
//
// Replaces analysis/time_correlations.py:60-92 (self_time_correlation): for every particle the
// mean-subtracted series x_i(t) is correlated with itself, c_i(t') = sum_t x_i(t) x_i(t+t'), and
// the mean and standard deviation of c_i(t') over the particles are returned per lag
// (SURVEY.md section 8f row 4).  The reference is a Python double loop of np.dot calls.
//
// tcf_dot_kernel: one block per particle.  The centred series sits in shared memory with one
// padding slot every 16 doubles; a thread owns 8 consecutive lags and slides an 8-entry register
// window over t, so every step is one broadcast load + one conflict-free load for 8 FMAs:
// fp64-FMA-bound, F*D - D^2/2 FMAs per particle.
// tcf_stats_kernel: one block per lag, two-pass mean / standard deviation over the particles
// (the arithmetic of np.mean / np.std).
#include "common.cuh"

namespace dsg {

constexpr int kTcfThreads = 256;
constexpr int kTcfLags = 8;  // lags per thread
constexpr int kTcfTail = 24;  // zero slots after the series (window look-ahead)

__host__ __device__ __forceinline__ int tcf_pad(int i) { return i + (i >> 4); }

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

// CROSS = false: block b correlates row b with itself.  CROSS = true: block b is the ordered pair
// (i, j), i != j, b = i (N-1) + (j < i ? j : j - 1); x_i(t) is multiplied with x_j(t + t').
template <bool SMEM, bool CROSS>
__global__ void __launch_bounds__(kTcfThreads)
tcf_dot_kernel(const double* __restrict__ data, long long stride_i, int n_part, int n_rows,
               int n_frames, int lag0, int n_lags, double* __restrict__ c) {
    extern __shared__ __align__(16) double tcf_smem[];
    __shared__ double red[kTcfThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    int ia = b, ib = b;
    if (CROSS) {
        ia = b / (n_part - 1);
        const int jj = b - ia * (n_part - 1);
        ib = jj < ia ? jj : jj + 1;
    }
    const double* __restrict__ row_a = data + (long long)ia * stride_i;
    const double* __restrict__ row_b = data + (long long)ib * stride_i;
    double part = 0.0;
    for (int t = tid; t < n_frames; t += kTcfThreads) part += row_a[t];
    const double mean_a = block_sum(part, red) / (double)n_frames;
    double mean_b = mean_a;
    if (CROSS) {
        part = 0.0;
        for (int t = tid; t < n_frames; t += kTcfThreads) part += row_b[t];
        mean_b = block_sum(part, red) / (double)n_frames;
    }
    const int padded = tcf_pad(n_frames + kTcfTail) + 2;
    double* xa = tcf_smem;
    double* xb = CROSS ? tcf_smem + padded : tcf_smem;
    if (SMEM) {
        for (int t = tid; t < n_frames + kTcfTail; t += kTcfThreads) {
            xa[tcf_pad(t)] = t < n_frames ? row_a[t] - mean_a : 0.0;
            if (CROSS) xb[tcf_pad(t)] = t < n_frames ? row_b[t] - mean_b : 0.0;
        }
        __syncthreads();
    }
    auto a_at = [&](int t) -> double {
        if (SMEM) return xa[tcf_pad(t)];
        return t < n_frames ? row_a[t] - mean_a : 0.0;
    };
    auto b_at = [&](int t) -> double {
        if (SMEM) return xb[tcf_pad(t)];
        return t < n_frames ? row_b[t] - mean_b : 0.0;
    };
    for (int l_rel = tid * kTcfLags; l_rel < n_lags; l_rel += kTcfThreads * kTcfLags) {
        const int lag = lag0 + l_rel;
        double acc[kTcfLags], w[kTcfLags];
#pragma unroll
        for (int k = 0; k < kTcfLags; ++k) {
            acc[k] = 0.0;
            w[k] = b_at(min(lag + k, n_frames + kTcfTail - 1));
        }
        const int t_end = n_frames - lag;  // terms t >= t_end multiply zero padding
        for (int t = 0; t < t_end; t += kTcfLags) {
#pragma unroll
            for (int u = 0; u < kTcfLags; ++u) {
                const double a = a_at(t + u);
#pragma unroll
                for (int k = 0; k < kTcfLags; ++k)
                    acc[k] = fma(a, w[(k + u) & (kTcfLags - 1)], acc[k]);  // w slot = x(t+u+lag+k)
                w[u] = b_at(min(t + u + lag + kTcfLags, n_frames + kTcfTail - 1));
            }
        }
#pragma unroll
        for (int k = 0; k < kTcfLags; ++k)
            if (l_rel + k < n_lags) c[(size_t)(l_rel + k) * n_rows + b] = acc[k];
    }
}

__global__ void __launch_bounds__(kTcfThreads)
tcf_stats_kernel(const double* __restrict__ c, int n_part, int n_lags, double* __restrict__ mean_out,
                 double* __restrict__ std_out) {
    __shared__ double red[kTcfThreads / 32];
    const int l = blockIdx.x;
    if (l >= n_lags) return;
    const double* __restrict__ v = c + (size_t)l * n_part;
    double s = 0.0;
    for (int i = threadIdx.x; i < n_part; i += kTcfThreads) s += v[i];
    const double mean = block_sum(s, red) / (double)n_part;
    double q = 0.0;
    for (int i = threadIdx.x; i < n_part; i += kTcfThreads) {
        const double d = v[i] - mean;
        q = fma(d, d, q);
    }
    const double var = block_sum(q, red) / (double)n_part;
    if (threadIdx.x == 0) {
        mean_out[l] = mean;
        std_out[l] = sqrt(var);
    }
}

constexpr size_t kTcfChunkBytes = (size_t)1 << 30;

static int tcf_chunk_lags(int n_part, int max_delay) {
    long long lags = (long long)(kTcfChunkBytes / (8 * (size_t)n_part));
    lags = lags / kTcfLags * kTcfLags;
    if (lags < kTcfLags) lags = kTcfLags;
    const long long need = ((long long)max_delay + kTcfLags - 1) / kTcfLags * kTcfLags;
    return (int)(lags < need ? lags : need);
}

}  // namespace dsg

using namespace dsg;

static int tcf_impl(bool cross, const double* d_data, int64_t stride_i, int n_part, int n_frames,
                    int max_delay, void* d_workspace, size_t workspace_bytes, double* d_mean_out,
                    double* d_std_out, cudaStream_t st, size_t* bytes_only) {
    if (n_part < (cross ? 2 : 1) || n_frames < 1 || max_delay < 1 || max_delay > n_frames)
        return set_error(DSG_ERR_INVALID_ARGUMENT,
                         "n_part=%d, n_frames=%d, max_delay=%d: need 1 <= max_delay <= n_frames%s",
                         n_part, n_frames, max_delay, cross ? " and n_part >= 2" : "");
    const long long rows_ll = cross ? (long long)n_part * (n_part - 1) : n_part;
    if (rows_ll > 0x7fffffffLL) return set_error(DSG_ERR_OVERFLOW, "too many particle pairs");
    const int n_rows = (int)rows_ll;
    const int chunk = tcf_chunk_lags(n_rows, max_delay);
    const size_t need = 8 * (size_t)n_rows * (size_t)chunk;
    if (bytes_only) {
        *bytes_only = need;
        return DSG_OK;
    }
    if (!d_data || !d_workspace || !d_mean_out || !d_std_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (workspace_bytes < need)
        return set_error(DSG_ERR_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed",
                         workspace_bytes, need);
    double* c = (double*)d_workspace;
    const size_t smem = 8 * (size_t)(tcf_pad(n_frames + kTcfTail) + 2) * (cross ? 2 : 1);
    const bool use_smem = smem <= 200 * 1024;
    if (use_smem) {
        if (cross)
            DSG_CUDA_CHECK(cudaFuncSetAttribute(tcf_dot_kernel<true, true>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
        else
            DSG_CUDA_CHECK(cudaFuncSetAttribute(tcf_dot_kernel<true, false>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
    }
    for (int lag0 = 0; lag0 < max_delay; lag0 += chunk) {
        const int n_lags = max_delay - lag0 < chunk ? max_delay - lag0 : chunk;
        if (use_smem && cross)
            tcf_dot_kernel<true, true><<<n_rows, kTcfThreads, smem, st>>>(
                d_data, stride_i, n_part, n_rows, n_frames, lag0, n_lags, c);
        else if (use_smem)
            tcf_dot_kernel<true, false><<<n_rows, kTcfThreads, smem, st>>>(
                d_data, stride_i, n_part, n_rows, n_frames, lag0, n_lags, c);
        else if (cross)
            tcf_dot_kernel<false, true><<<n_rows, kTcfThreads, 0, st>>>(
                d_data, stride_i, n_part, n_rows, n_frames, lag0, n_lags, c);
        else
            tcf_dot_kernel<false, false><<<n_rows, kTcfThreads, 0, st>>>(
                d_data, stride_i, n_part, n_rows, n_frames, lag0, n_lags, c);
        DSG_LAUNCH_CHECK("tcf_dot_kernel");
        tcf_stats_kernel<<<n_lags, kTcfThreads, 0, st>>>(c, n_rows, n_lags, d_mean_out + lag0,
                                                        d_std_out + lag0);
        DSG_LAUNCH_CHECK("tcf_stats_kernel");
    }
    return DSG_OK;
}

extern "C" int dsg_tcf_workspace_bytes(int n_part, int n_frames, int max_delay, size_t* bytes_out) {
    if (!bytes_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "bytes_out is NULL");
    return tcf_impl(false, nullptr, 0, n_part, n_frames, max_delay, nullptr, 0, nullptr, nullptr,
                    nullptr, bytes_out);
}

extern "C" int dsg_self_time_correlation(const double* d_data, int64_t stride_i, int n_part,
                                         int n_frames, int max_delay, void* d_workspace,
                                         size_t workspace_bytes, double* d_mean_out,
                                         double* d_std_out, void* stream) {
    return tcf_impl(false, d_data, stride_i, n_part, n_frames, max_delay, d_workspace,
                    workspace_bytes, d_mean_out, d_std_out, (cudaStream_t)stream, nullptr);
}

extern "C" int dsg_cross_tcf_workspace_bytes(int n_part, int n_frames, int max_delay,
                                             size_t* bytes_out) {
    if (!bytes_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "bytes_out is NULL");
    return tcf_impl(true, nullptr, 0, n_part, n_frames, max_delay, nullptr, 0, nullptr, nullptr,
                    nullptr, bytes_out);
}

extern "C" int dsg_cross_time_correlation(const double* d_data, int64_t stride_i, int n_part,
                                          int n_frames, int max_delay, void* d_workspace,
                                          size_t workspace_bytes, double* d_mean_out,
                                          double* d_std_out, void* stream) {
    return tcf_impl(true, d_data, stride_i, n_part, n_frames, max_delay, d_workspace,
                    workspace_bytes, d_mean_out, d_std_out, (cudaStream_t)stream, nullptr);
}
