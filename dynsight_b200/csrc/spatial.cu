// spatial.cu -- spatial average of a descriptor over each particle's neighbourhood (sm_100a).
//
// Replaces analysis/spatial_average.py:31-63 (processframe: FastNS pair search, then for every
// particle np.mean over its neighbours and itself) -- SURVEY.md section 8f row 1.  The neighbour
// rows come from dsg_neigh_rows in inclusive-cutoff mode on positions wrapped into the box (what
// FastNS does before it bins); this file holds the wrap and the segmented mean.
//
// HBM/L2-bound gathers: per (particle, frame) 4 (k+1) bytes of row + 8 (k+1) C bytes of values
// read (C = components, 1 for a scalar descriptor), 8 C written.
#include "common.cuh"

namespace dsg {

constexpr unsigned kFullSp = 0xffffffffu;
constexpr int kSpWarps = 8;

// x - floor(x / box) * box in the input precision, folded back if rounding lands on the box edge
template <typename T>
__global__ void __launch_bounds__(256)
wrap_kernel(const T* __restrict__ pos, const double* __restrict__ box, int n_atoms,
            T* __restrict__ out) {
    const int f = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_atoms * 3) return;
    const int k = (int)(i % 3);
    const T b = (T)box[3 * f + k];
    const size_t g = (size_t)f * n_atoms * 3 + i;
    T v = pos[g];
    if (b > (T)0) {
        v -= floor(v / b) * b;
        if (v >= b) v -= b;
        if (v < (T)0) v = (T)0;
    }
    out[g] = v;
}

// Scalar descriptor: G lanes per (particle, frame) (4 for rows of ~10 entries ... 32 for long rows);
// the lanes stride over the row, shuffle reduction inside the group.
template <int G>
__global__ void __launch_bounds__(kSpWarps * 32)
spatial_mean_scalar_kernel(const int* __restrict__ counts, const long long* __restrict__ row_start,
                           const int* __restrict__ rows, int n_frames, int n_cent,
                           const double* __restrict__ val, long long v_stride_i,
                           long long v_stride_f, double* __restrict__ out, long long o_stride_i,
                           long long o_stride_f) {
    const int lane = threadIdx.x & (G - 1);
    long long w = ((long long)blockIdx.x * (kSpWarps * 32) + threadIdx.x) / G;
    const bool active = w < (long long)n_frames * n_cent;
    if (!active) w = 0;  // stay for the group shuffles
    const int f = (int)(w / n_cent), i = (int)(w % n_cent);
    const size_t fi = (size_t)f * n_cent + i;
    const int n = active ? counts[fi] : 0;
    const long long rs = row_start[fi];
    const double* __restrict__ vf = val + (long long)f * v_stride_f;
    double s = 0.0;
    if (rs >= 0) {
        const int* __restrict__ row = rows + rs;
        for (int t = lane; t < n; t += G) s += vf[(long long)row[t] * v_stride_i];
    }
    if (lane == 0) s += vf[(long long)i * v_stride_i];  // the particle itself
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(kFullSp, s, off);
    if (lane == 0 && active)
        out[(long long)i * o_stride_i + (long long)f * o_stride_f] = s / (double)(n + 1);
}

// Vector descriptor (components contiguous): one warp per (particle, frame); lanes own components
// (coalesced reads of each neighbour's vector), neighbours are summed in row order, four at a time.
__global__ void __launch_bounds__(kSpWarps * 32)
spatial_mean_vector_kernel(const int* __restrict__ counts, const long long* __restrict__ row_start,
                           const int* __restrict__ rows, int n_frames, int n_cent, int n_comp,
                           const double* __restrict__ val, long long v_stride_i,
                           long long v_stride_f, double* __restrict__ out, long long o_stride_i,
                           long long o_stride_f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kSpWarps + warp;
    if (w >= (long long)n_frames * n_cent) return;
    const int f = (int)(w / n_cent), i = (int)(w % n_cent);
    const size_t fi = (size_t)f * n_cent + i;
    const int n = counts[fi];
    const long long rs = row_start[fi];
    const double* __restrict__ vf = val + (long long)f * v_stride_f;
    double* __restrict__ o = out + (long long)i * o_stride_i + (long long)f * o_stride_f;
    const double inv = 1.0 / (double)(n + 1);
    const int* __restrict__ row = rs >= 0 ? rows + rs : nullptr;
    for (int c0 = 0; c0 < n_comp; c0 += 128) {
        // four components per lane and four neighbours per step: 16 independent gathers in flight;
        // the row's indices are loaded coalesced (lane t holds row[t0 + t]) and broadcast
        double s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + lane + 32 * u;
            s[u] = c < n_comp ? vf[(long long)i * v_stride_i + c] : 0.0;
        }
        const int n_row = row ? n : 0;
        for (int t0 = 0; t0 < n_row; t0 += 32) {
            const int mine = t0 + lane < n_row ? row[t0 + lane] : 0;
            const int m = min(32, n_row - t0);
            int t = 0;
            for (; t + 4 <= m; t += 4) {
                const double* __restrict__ v0 = vf + (long long)__shfl_sync(kFullSp, mine, t) * v_stride_i;
                const double* __restrict__ v1 = vf + (long long)__shfl_sync(kFullSp, mine, t + 1) * v_stride_i;
                const double* __restrict__ v2 = vf + (long long)__shfl_sync(kFullSp, mine, t + 2) * v_stride_i;
                const double* __restrict__ v3 = vf + (long long)__shfl_sync(kFullSp, mine, t + 3) * v_stride_i;
                double x[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + lane + 32 * u;
                    const bool ok = c < n_comp;
                    x[0][u] = ok ? v0[c] : 0.0;
                    x[1][u] = ok ? v1[c] : 0.0;
                    x[2][u] = ok ? v2[c] : 0.0;
                    x[3][u] = ok ? v3[c] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) s[u] += x[0][u], s[u] += x[1][u], s[u] += x[2][u], s[u] += x[3][u];
            }
            for (; t < m; ++t) {
                const double* __restrict__ vj = vf + (long long)__shfl_sync(kFullSp, mine, t) * v_stride_i;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + lane + 32 * u;
                    if (c < n_comp) s[u] += vj[c];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + lane + 32 * u;
            if (c < n_comp) o[c] = s[u] * inv;
        }
    }
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_wrap_positions(const void* d_pos, int pos_dtype, int n_frames, int n_atoms,
                                  const double* d_box, void* d_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_pos || !d_box || !d_out || n_frames < 1 || n_atoms < 1 || n_frames > 65535)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL pointer or bad sizes");
    dim3 grid((unsigned)(((size_t)n_atoms * 3 + 255) / 256), n_frames);
    if (pos_dtype == DSG_F32)
        wrap_kernel<float><<<grid, 256, 0, st>>>((const float*)d_pos, d_box, n_atoms, (float*)d_out);
    else if (pos_dtype == DSG_F64)
        wrap_kernel<double><<<grid, 256, 0, st>>>((const double*)d_pos, d_box, n_atoms,
                                                  (double*)d_out);
    else
        return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                         pos_dtype);
    DSG_LAUNCH_CHECK("wrap_kernel");
    return DSG_OK;
}

extern "C" int dsg_spatial_average_rows(const int32_t* d_counts, const int64_t* d_row_start,
                                        const int32_t* d_rows, int n_frames, int n_cent,
                                        int n_comp, const double* d_values, int64_t v_stride_i,
                                        int64_t v_stride_f, double* d_out, int64_t o_stride_i,
                                        int64_t o_stride_f, int expected_row_len, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_counts || !d_row_start || !d_rows || !d_values || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_frames < 1 || n_cent < 1 || n_comp < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d, n_cent=%d, n_comp=%d must be >= 1",
                         n_frames, n_cent, n_comp);
    const long long warps = (long long)n_frames * n_cent;
    const unsigned blocks = (unsigned)((warps + kSpWarps - 1) / kSpWarps);
    if (n_comp == 1) {
        const int g = expected_row_len <= 0 ? 32
                      : expected_row_len <= 12 ? 4
                      : expected_row_len <= 24 ? 8
                      : expected_row_len <= 48 ? 16 : 32;
        const unsigned gb = (unsigned)((warps * g + kSpWarps * 32 - 1) / (kSpWarps * 32));
#define DSG_SPM(G)                                                                               \
    spatial_mean_scalar_kernel<G><<<gb, kSpWarps * 32, 0, st>>>(                                 \
        d_counts, (const long long*)d_row_start, d_rows, n_frames, n_cent, d_values, v_stride_i, \
        v_stride_f, d_out, o_stride_i, o_stride_f)
        if (g == 4) DSG_SPM(4); else if (g == 8) DSG_SPM(8);
        else if (g == 16) DSG_SPM(16); else DSG_SPM(32);
#undef DSG_SPM
    } else {
        spatial_mean_vector_kernel<<<blocks, kSpWarps * 32, 0, st>>>(
            d_counts, (const long long*)d_row_start, d_rows, n_frames, n_cent, n_comp, d_values,
            v_stride_i, v_stride_f, d_out, o_stride_i, o_stride_f);
    }
    DSG_LAUNCH_CHECK("spatial_mean_kernel");
    return DSG_OK;
}
