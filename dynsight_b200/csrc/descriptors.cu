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

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFullD, v, off);
    return v;
}

// psi (misc.py:76-93): rows with <= 1 entries give 0; entries j == i are skipped in the sum but
// counted in the divisor len(neighbors); (x, y) differences are raw (no minimum image).
template <typename T>
__global__ void __launch_bounds__(kDescWarps * 32)
psi_kernel(const T* __restrict__ pos, const int* __restrict__ indptr,
           const long long* __restrict__ frame_off, const int* __restrict__ indices, int n_frames,
           int n_atoms, int order, double* __restrict__ out, long long ld_out, long long col0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kDescWarps + warp;
    if (w >= (long long)n_frames * n_atoms) return;
    const int f = (int)(w / n_atoms), i = (int)(w % n_atoms);
    const int* ip = indptr + (size_t)f * (n_atoms + 1) + i;
    const int s = ip[0], n = ip[1] - ip[0];
    double re = 0.0, im = 0.0;
    if (n > 1) {
        const T* __restrict__ pf = pos + (size_t)f * n_atoms * 3;
        const T xi = pf[(size_t)i * 3], yi = pf[(size_t)i * 3 + 1];
        const int* __restrict__ row = indices + frame_off[f] + s;
        for (int t = lane; t < n; t += 32) {
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
        re = warp_sum_d(re);
        im = warp_sum_d(im);
        re /= (double)n;
        im /= (double)n;
    }
    if (lane == 0) out[(long long)i * ld_out + col0 + f] = sqrt(re * re + im * im);
}

// phi (misc.py:116-133): mean over neighbours j != i with a non-zero vector of the cosine
// similarity (clipped to [-1, 1] like scipy's cosine distance), 0 if v_i == 0 or no valid neighbour.
// displacement != 0: v(f, a) = pos[f + 1][a] - pos[0][a]  (the reference never advances r_0,
// misc.py:207-213) with the rows of frame f; otherwise vec holds one vector per (frame, atom).
template <typename T>
__global__ void __launch_bounds__(kDescWarps * 32)
phi_kernel(const T* __restrict__ vec, int displacement, const int* __restrict__ indptr,
           const long long* __restrict__ frame_off, const int* __restrict__ indices, int n_out,
           int n_atoms, double* __restrict__ out, long long ld_out, long long col0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kDescWarps + warp;
    if (w >= (long long)n_out * n_atoms) return;
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
    const int s = ip[0], n = ip[1] - ip[0];
    double sum = 0.0, cnt = 0.0;
    const double uu = vi[0] * vi[0] + vi[1] * vi[1] + vi[2] * vi[2];
    if (n > 0 && (vi[0] != 0.0 || vi[1] != 0.0 || vi[2] != 0.0)) {
        const int* __restrict__ row = indices + frame_off[f] + s;
        for (int t = lane; t < n; t += 32) {
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
        sum = warp_sum_d(sum);
        cnt = warp_sum_d(cnt);
    }
    if (lane == 0) out[(long long)i * ld_out + col0 + f] = cnt > 0.0 ? sum / cnt : 0.0;
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_orientational_op(const void* d_pos, int pos_dtype, const int32_t* d_indptr,
                                    const int64_t* d_frame_off, const int32_t* d_indices,
                                    int n_frames, int n_atoms, int order, double* d_out,
                                    int64_t ld_out, int64_t col0, void* stream) {
    if (!d_pos || !d_indptr || !d_frame_off || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_frames < 1 || n_atoms < 1 || order < 0)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d, n_atoms=%d, order=%d out of range",
                         n_frames, n_atoms, order);
    const long long warps = (long long)n_frames * n_atoms;
    const long long blocks = (warps + kDescWarps - 1) / kDescWarps;
    if (blocks > 0x7fffffffLL) return set_error(DSG_ERR_OVERFLOW, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (pos_dtype == DSG_F32)
        psi_kernel<float><<<(unsigned)blocks, kDescWarps * 32, 0, st>>>(
            (const float*)d_pos, d_indptr, (const long long*)d_frame_off, d_indices, n_frames,
            n_atoms, order, d_out, ld_out, col0);
    else if (pos_dtype == DSG_F64)
        psi_kernel<double><<<(unsigned)blocks, kDescWarps * 32, 0, st>>>(
            (const double*)d_pos, d_indptr, (const long long*)d_frame_off, d_indices, n_frames,
            n_atoms, order, d_out, ld_out, col0);
    else
        return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                         pos_dtype);
    DSG_LAUNCH_CHECK("psi_kernel");
    return DSG_OK;
}

extern "C" int dsg_velocity_alignment(const void* d_vec, int dtype, int displacement,
                                      const int32_t* d_indptr, const int64_t* d_frame_off,
                                      const int32_t* d_indices, int n_out, int n_atoms,
                                      double* d_out, int64_t ld_out, int64_t col0, void* stream) {
    if (!d_vec || !d_indptr || !d_frame_off || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_out < 1 || n_atoms < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_out=%d, n_atoms=%d out of range", n_out,
                         n_atoms);
    const long long warps = (long long)n_out * n_atoms;
    const long long blocks = (warps + kDescWarps - 1) / kDescWarps;
    if (blocks > 0x7fffffffLL) return set_error(DSG_ERR_OVERFLOW, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DSG_F32)
        phi_kernel<float><<<(unsigned)blocks, kDescWarps * 32, 0, st>>>(
            (const float*)d_vec, displacement, d_indptr, (const long long*)d_frame_off, d_indices,
            n_out, n_atoms, d_out, ld_out, col0);
    else if (dtype == DSG_F64)
        phi_kernel<double><<<(unsigned)blocks, kDescWarps * 32, 0, st>>>(
            (const double*)d_vec, displacement, d_indptr, (const long long*)d_frame_off, d_indices,
            n_out, n_atoms, d_out, ld_out, col0);
    else
        return set_error(DSG_ERR_INVALID_ARGUMENT, "dtype=%d is neither DSG_F32 nor DSG_F64", dtype);
    DSG_LAUNCH_CHECK("phi_kernel");
    return DSG_OK;
}
