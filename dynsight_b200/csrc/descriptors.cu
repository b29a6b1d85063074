// descriptors.cu -- per-centre reductions over the neighbour CSR rows (sm_100a):
// orientational order parameter |psi_n| and velocity alignment phi.
//
// Replaces the Python double loops of descriptors/misc.py:15-95 (orientational_order_param) and
// :98-219 (compute_mean_alignment + velocity_alignment).  One warp owns one (atom, frame); lanes
// stride over the row, partial sums are reduced with shuffles.  Arithmetic: coordinate differences
// in the input dtype (float32 for MDAnalysis positions, like the reference), everything after in
// float64; e^{i n theta} is the n-th power of the unit vector (no atan2).  HBM-bound row reads.
#include "common.cuh"

namespace dsg {

constexpr unsigned kFullD = 0xffffffffu;
constexpr int kDescWarps = 8;

// sum over the G consecutive lanes that share one (atom, frame)
template <int G>
__device__ __forceinline__ double group_sum_d(double v) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(kFullD, v, off);
    return v;
}

// psi (misc.py:76-93): rows with <= 1 entries give 0; entries j == i are skipped in the sum but
// counted in the divisor len(neighbors); (x, y) differences are raw (no minimum image).
template <typename T, int G>
__global__ void __launch_bounds__(kDescWarps * 32)
psi_kernel(const T* __restrict__ pos, const int* __restrict__ indptr,
           const long long* __restrict__ frame_off, const int* __restrict__ indices, int n_frames,
           int n_atoms, int order, double* __restrict__ out, long long ld_out, long long col0) {
    // G lanes per (atom, frame): short rows leave few lanes idle (G = 4 for ~10 neighbours)
    const int lane = threadIdx.x & (G - 1);
    long long w = ((long long)blockIdx.x * (kDescWarps * 32) + threadIdx.x) / G;
    const bool active = w < (long long)n_frames * n_atoms;
    if (!active) w = 0;  // stay for the group shuffles
    const int f = (int)(w / n_atoms), i = (int)(w % n_atoms);
    const int* ip = indptr + (size_t)f * (n_atoms + 1) + i;
    const int s = ip[0], n = active ? ip[1] - ip[0] : 0;
    double re = 0.0, im = 0.0;
    if (n > 1) {
        const T* __restrict__ pf = pos + (size_t)f * n_atoms * 3;
        const T xi = pf[(size_t)i * 3], yi = pf[(size_t)i * 3 + 1];
        const int* __restrict__ row = indices + frame_off[f] + s;
        for (int t = lane; t < n; t += G) {
            const int j = row[t];
            if (j == i) continue;
            const double x = (double)(T)(pf[(size_t)j * 3] - xi);
            const double y = (double)(T)(pf[(size_t)j * 3 + 1] - yi);
            const double r = sqrt(x * x + y * y);
            const double c = r > 0.0 ? x / r : 1.0, sn = r > 0.0 ? y / r : 0.0;  // arctan2(0,0) = 0
            double pr = 1.0, pi = 0.0;
            for (int k = 0; k < order; ++k) {
                const double t2 = pr * c - pi * sn;
                pi = pr * sn + pi * c;
                pr = t2;
            }
            re += pr;
            im += pi;
        }
    }
    re = group_sum_d<G>(re);
    im = group_sum_d<G>(im);
    if (n > 1) {
        re /= (double)n;
        im /= (double)n;
    }
    if (lane == 0 && active) out[(long long)i * ld_out + col0 + f] = sqrt(re * re + im * im);
}

// phi (misc.py:116-133): mean over neighbours j != i with a non-zero vector of the cosine
// similarity (clipped to [-1, 1] like scipy's cosine distance), 0 if v_i == 0 or no valid neighbour.
// displacement != 0: v(f, a) = pos[f + 1][a] - pos[0][a]  (the reference never advances r_0,
// misc.py:207-213) with the rows of frame f; otherwise vec holds one vector per (frame, atom).
template <typename T, int G>
__global__ void __launch_bounds__(kDescWarps * 32)
phi_kernel(const T* __restrict__ vec, int displacement, const int* __restrict__ indptr,
           const long long* __restrict__ frame_off, const int* __restrict__ indices, int n_out,
           int n_atoms, double* __restrict__ out, long long ld_out, long long col0) {
    const int lane = threadIdx.x & (G - 1);
    long long w = ((long long)blockIdx.x * (kDescWarps * 32) + threadIdx.x) / G;
    const bool active = w < (long long)n_out * n_atoms;
    if (!active) w = 0;  // stay for the group shuffles
    const int f = (int)(w / n_atoms), i = (int)(w % n_atoms);
    const T* __restrict__ v1 = vec + (size_t)(displacement ? f + 1 : f) * n_atoms * 3;
    const T* __restrict__ v0 = vec;  // frame 0 (displacement mode)
    auto load = [&](int a, double* v) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T x = v1[(size_t)a * 3 + k];
            if (displacement) x = (T)(x - v0[(size_t)a * 3 + k]);
            v[k] = (double)x;
        }
    };
    double vi[3];
    load(i, vi);
    const int* ip = indptr + (size_t)f * (n_atoms + 1) + i;
    const int s = ip[0], n = active ? ip[1] - ip[0] : 0;
    double sum = 0.0, cnt = 0.0;
    const double uu = vi[0] * vi[0] + vi[1] * vi[1] + vi[2] * vi[2];
    if (n > 0 && (vi[0] != 0.0 || vi[1] != 0.0 || vi[2] != 0.0)) {
        const int* __restrict__ row = indices + frame_off[f] + s;
        for (int t = lane; t < n; t += G) {
            const int j = row[t];
            if (j == i) continue;
            double vj[3];
            load(j, vj);
            if (vj[0] == 0.0 && vj[1] == 0.0 && vj[2] == 0.0) continue;
            const double uv = vi[0] * vj[0] + vi[1] * vj[1] + vi[2] * vj[2];
            const double vv = vj[0] * vj[0] + vj[1] * vj[1] + vj[2] * vj[2];
            double dist = 1.0 - uv / sqrt(uu * vv);
            dist = fmax(0.0, fmin(dist, 2.0));
            sum += 1.0 - dist;
            cnt += 1.0;
        }
    }
    sum = group_sum_d<G>(sum);
    cnt = group_sum_d<G>(cnt);
    if (lane == 0 && active) out[(long long)i * ld_out + col0 + f] = cnt > 0.0 ? sum / cnt : 0.0;
}

// lanes that share one (atom, frame): enough to cover the expected row in about two passes
static int group_lanes(int expected_row_len) {
    if (expected_row_len <= 0) return 32;
    if (expected_row_len <= 12) return 4;
    if (expected_row_len <= 24) return 8;
    if (expected_row_len <= 48) return 16;
    return 32;
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_orientational_op(const void* d_pos, int pos_dtype, const int32_t* d_indptr,
                                    const int64_t* d_frame_off, const int32_t* d_indices,
                                    int n_frames, int n_atoms, int order, double* d_out,
                                    int64_t ld_out, int64_t col0, int expected_row_len,
                                    void* stream) {
    if (!d_pos || !d_indptr || !d_frame_off || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_frames < 1 || n_atoms < 1 || order < 0)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d, n_atoms=%d, order=%d out of range",
                         n_frames, n_atoms, order);
    const int g = group_lanes(expected_row_len);
    const long long items = (long long)n_frames * n_atoms;
    const long long blocks = (items * g + kDescWarps * 32 - 1) / (kDescWarps * 32);
    if (blocks > 0x7fffffffLL) return set_error(DSG_ERR_OVERFLOW, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (pos_dtype != DSG_F32 && pos_dtype != DSG_F64)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                         pos_dtype);
#define DSG_PSI(T, G)                                                                            \
    psi_kernel<T, G><<<(unsigned)blocks, kDescWarps * 32, 0, st>>>(                              \
        (const T*)d_pos, d_indptr, (const long long*)d_frame_off, d_indices, n_frames, n_atoms,  \
        order, d_out, ld_out, col0)
    if (pos_dtype == DSG_F32) {
        if (g == 4) DSG_PSI(float, 4); else if (g == 8) DSG_PSI(float, 8);
        else if (g == 16) DSG_PSI(float, 16); else DSG_PSI(float, 32);
    } else {
        if (g == 4) DSG_PSI(double, 4); else if (g == 8) DSG_PSI(double, 8);
        else if (g == 16) DSG_PSI(double, 16); else DSG_PSI(double, 32);
    }
#undef DSG_PSI
    DSG_LAUNCH_CHECK("psi_kernel");
    return DSG_OK;
}

extern "C" int dsg_velocity_alignment(const void* d_vec, int dtype, int displacement,
                                      const int32_t* d_indptr, const int64_t* d_frame_off,
                                      const int32_t* d_indices, int n_out, int n_atoms,
                                      double* d_out, int64_t ld_out, int64_t col0,
                                      int expected_row_len, void* stream) {
    if (!d_vec || !d_indptr || !d_frame_off || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_out < 1 || n_atoms < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_out=%d, n_atoms=%d out of range", n_out,
                         n_atoms);
    const int g = group_lanes(expected_row_len);
    const long long items = (long long)n_out * n_atoms;
    const long long blocks = (items * g + kDescWarps * 32 - 1) / (kDescWarps * 32);
    if (blocks > 0x7fffffffLL) return set_error(DSG_ERR_OVERFLOW, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != DSG_F32 && dtype != DSG_F64)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "dtype=%d is neither DSG_F32 nor DSG_F64", dtype);
#define DSG_PHI(T, G)                                                                            \
    phi_kernel<T, G><<<(unsigned)blocks, kDescWarps * 32, 0, st>>>(                              \
        (const T*)d_vec, displacement, d_indptr, (const long long*)d_frame_off, d_indices, n_out,\
        n_atoms, d_out, ld_out, col0)
    if (dtype == DSG_F32) {
        if (g == 4) DSG_PHI(float, 4); else if (g == 8) DSG_PHI(float, 8);
        else if (g == 16) DSG_PHI(float, 16); else DSG_PHI(float, 32);
    } else {
        if (g == 4) DSG_PHI(double, 4); else if (g == 8) DSG_PHI(double, 8);
        else if (g == 16) DSG_PHI(double, 16); else DSG_PHI(double, 32);
    }
#undef DSG_PHI
    DSG_LAUNCH_CHECK("phi_kernel");
    return DSG_OK;
}
