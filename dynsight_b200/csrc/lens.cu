// lens.cu -- LENS from sorted CSR rows + coordination numbers (sm_100a).
//
// Replaces lens/lens.py:150-189 (lens_from_two_csr: serial two-pointer merge per centre) and the
// Python double loop of trajectory/trajectory.py:168-184.  One warp owns one (centre, frame-pair):
// both sorted rows are loaded coalesced, one element per lane, and the intersection size comes from
// lane-parallel equality tests against shuffled copies of the other row (a merge without a serial
// pointer chase).  Integer work is exact; the single float64 division is IEEE (bit-exact).
#include "common.cuh"

namespace dsg {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kLensWarps = 8;

// |A n B| for two ascending rows of distinct int32, cooperatively by one warp.
__device__ __forceinline__ int warp_intersect(const int* __restrict__ ra, int na,
                                              const int* __restrict__ rb, int nb, int lane) {
    int inter = 0;
    for (int a0 = 0; a0 < na; a0 += 32) {
        const int ia = a0 + lane;
        const int va = ia < na ? ra[ia] : -1;           // ids are >= 0: -1 never matches
        const int a_lo = __shfl_sync(kFullMask, va, 0);
        const int a_last = min(31, na - a0 - 1);
        const int a_hi = __shfl_sync(kFullMask, va, a_last);
        bool hit = false;
        for (int b0 = 0; b0 < nb; b0 += 32) {
            const int ib = b0 + lane;
            const int vb = ib < nb ? rb[ib] : -2;
            const int b_last = min(31, nb - b0 - 1);
            const int b_lo = __shfl_sync(kFullMask, vb, 0);
            const int b_hi = __shfl_sync(kFullMask, vb, b_last);
            if (b_hi < a_lo) continue;  // chunk of B entirely below this chunk of A
            if (b_lo > a_hi) break;     // sorted: nothing further can match
            for (int j = 0; j <= b_last; ++j) hit |= (va == __shfl_sync(kFullMask, vb, j));
        }
        inter += __popc(__ballot_sync(kFullMask, hit));
    }
    return inter;
}

// |A n B| for two UNSORTED rows of distinct int32 (rows path): all chunk pairs are compared.
__device__ __forceinline__ int warp_intersect_unsorted(const int* __restrict__ ra, int na,
                                                       const int* __restrict__ rb, int nb,
                                                       int lane) {
    int inter = 0;
    for (int a0 = 0; a0 < na; a0 += 32) {
        const int ia = a0 + lane;
        const int va = ia < na ? ra[ia] : -1;
        bool hit = false;
        for (int b0 = 0; b0 < nb; b0 += 32) {
            const int ib = b0 + lane;
            const int vb = ib < nb ? rb[ib] : -2;
            const int b_last = min(31, nb - b0 - 1);
            for (int j = 0; j <= b_last; ++j) hit |= (va == __shfl_sync(kFullMask, vb, j));
        }
        inter += __popc(__ballot_sync(kFullMask, hit));
    }
    return inter;
}


// grid: (ceil(n_pairs / kLensWarps), n_cent); warp w of a block handles pair blockIdx.x*W + w of
// centre blockIdx.y, so one block writes kLensWarps consecutive doubles of an output row and the
// row shared by pairs k and k+delay is served from L1.
__global__ void __launch_bounds__(kLensWarps * 32)
lens_pairs_kernel(const int* __restrict__ indptr, const long long* __restrict__ frame_off,
                  const int* __restrict__ indices, int n_pairs, int n_cent, int delay,
                  double* __restrict__ out, long long ld_out, long long col0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * kLensWarps + warp;
    const int i = blockIdx.y;
    if (k >= n_pairs) return;
    const size_t stride = (size_t)n_cent + 1;
    const int* p1 = indptr + (size_t)k * stride + i;
    const int* p2 = indptr + (size_t)(k + delay) * stride + i;
    const int s1 = p1[0], e1 = p1[1], s2 = p2[0], e2 = p2[1];
    const int na = e1 - s1, nb = e2 - s2;
    int inter = 0;
    if (na > 0 && nb > 0)
        inter = warp_intersect(indices + frame_off[k] + s1, na, indices + frame_off[k + delay] + s2,
                               nb, lane);
    if (lane == 0) out[(long long)i * ld_out + col0 + k] = lens_value(na, nb, inter);
}

// rows path: per-centre (row_start, count) instead of a canonical CSR; rows are unsorted
__global__ void __launch_bounds__(kLensWarps * 32)
lens_pairs_rows_kernel(const int* __restrict__ counts, const long long* __restrict__ row_start,
                       const int* __restrict__ rows, int n_pairs, int n_cent, int delay,
                       double* __restrict__ out, long long ld_out, long long col0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * kLensWarps + warp;
    const int i = blockIdx.y;
    if (k >= n_pairs) return;
    const size_t i1 = (size_t)k * n_cent + i, i2 = (size_t)(k + delay) * n_cent + i;
    const int na = counts[i1], nb = counts[i2];
    int inter = 0;
    if (na > 0 && nb > 0) {
        const long long s1 = row_start[i1], s2 = row_start[i2];
        if (s1 >= 0 && s2 >= 0) inter = warp_intersect_unsorted(rows + s1, na, rows + s2, nb, lane);
    }
    if (lane == 0) out[(long long)i * ld_out + col0 + k] = lens_value(na, nb, inter);
}

// rows path, delay == 1: one warp owns one centre and a run of kRunPairs consecutive pairs.  The
// kRunPairs+1 rows are loaded once (all loads in flight together), consecutive rows are compared
// in registers, and the run's results leave as one 64-byte store.
constexpr int kRunPairs = 8;
__global__ void __launch_bounds__(kLensWarps * 32)
lens_rows_run_kernel(const int* __restrict__ counts, const long long* __restrict__ row_start,
                     const int* __restrict__ rows, int n_pairs, int n_cent,
                     double* __restrict__ out, long long ld_out, long long col0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.y * kLensWarps + warp;
    const int k0 = blockIdx.x * kRunPairs;
    if (i >= n_cent || k0 >= n_pairs) return;
    const int np = min(kRunPairs, n_pairs - k0);
    int cnt = 0;
    long long st = 0;
    if (lane <= np) {
        const size_t idx = (size_t)(k0 + lane) * n_cent + i;
        cnt = counts[idx];
        st = row_start[idx];
        if (st < 0) cnt = 0;
    }
    int maxc = cnt;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) maxc = max(maxc, __shfl_xor_sync(kFullMask, maxc, off));
    double res = 0.0;
    if (maxc <= 32) {
        int rowv[kRunPairs + 1];
#pragma unroll
        for (int j = 0; j <= kRunPairs; ++j) {
            const int cj = __shfl_sync(kFullMask, cnt, j);
            const long long sj = __shfl_sync(kFullMask, st, j);
            rowv[j] = (j <= np && lane < cj) ? rows[sj + lane] : -1 - j;  // distinct sentinels
        }
#pragma unroll
        for (int k = 0; k < kRunPairs; ++k) {
            if (k < np) {
                const int ca = __shfl_sync(kFullMask, cnt, k), cb = __shfl_sync(kFullMask, cnt, k + 1);
                bool hit = false;
                for (int j = 0; j < cb; ++j) hit |= (rowv[k] == __shfl_sync(kFullMask, rowv[k + 1], j));
                const int inter = __popc(__ballot_sync(kFullMask, hit));
                if (lane == k) res = lens_value(ca, cb, inter);
            }
        }
    } else {
        for (int k = 0; k < np; ++k) {
            const int ca = __shfl_sync(kFullMask, cnt, k), cb = __shfl_sync(kFullMask, cnt, k + 1);
            const long long sa = __shfl_sync(kFullMask, st, k), sb = __shfl_sync(kFullMask, st, k + 1);
            int inter = 0;
            if (ca > 0 && cb > 0) inter = warp_intersect_unsorted(rows + sa, ca, rows + sb, cb, lane);
            if (lane == k) res = lens_value(ca, cb, inter);
        }
    }
    if (lane < np) out[(long long)i * ld_out + col0 + k0 + lane] = res;
}

// rows path, delay == 1, short rows: ONE THREAD per centre and a run of kRunPairs consecutive pairs.
// Both rows of a pair sit in registers (32 values each, loaded in chunks of 8 up to the longest
// row of the warp, every row is read once and then moves from the "later" to the "earlier" role);
// every element of the earlier row is compared with the registers of the later one -- a few
// hundred thread-level instructions per pair instead of a warp-wide shuffle loop that leaves two
// thirds of the lanes idle on rows of ~10 entries.  A warp that meets a row longer than 32
// handles its 32 centres with the warp-cooperative intersection.
constexpr int kLtThreads = 128;
constexpr int kLtCap = 32;

__device__ __forceinline__ void lt_load_row(int (&r)[kLtCap], const int* __restrict__ rows,
                                            long long start, int cnt, int warp_max, int fill) {
#pragma unroll
    for (int c8 = 0; c8 < kLtCap / 8; ++c8) {
        const bool need = c8 * 8 < warp_max && warp_max <= kLtCap;   // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j)
            r[c8 * 8 + j] = (need && c8 * 8 + j < cnt) ? rows[start + c8 * 8 + j] : fill;
    }
}

__global__ void __launch_bounds__(kLtThreads)
lens_rows_thread_kernel(const int* __restrict__ counts, const long long* __restrict__ row_start,
                        const int* __restrict__ rows, int n_pairs, int n_cent,
                        double* __restrict__ out, long long ld_out, long long col0) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.y * kLtThreads + threadIdx.x;
    const int k0 = blockIdx.x * kRunPairs;
    if (k0 >= n_pairs) return;
    const bool active = i < n_cent;   // inactive threads stay for the warp-wide operations
    const int np = min(kRunPairs, n_pairs - k0);
    int ca = 0;
    long long sa = -1;
    if (active) {
        const size_t idx = (size_t)k0 * n_cent + i;
        ca = counts[idx];
        sa = row_start[idx];
        if (sa < 0) ca = 0;
    }
    int maxa = __reduce_max_sync(kFullMask, ca);
    int a[kLtCap], b[kLtCap];
    lt_load_row(a, rows, sa, ca, maxa, -2);
    for (int k = 0; k < np; ++k) {
        int cb = 0;
        long long sb = -1;
        if (active) {
            const size_t idx = (size_t)(k0 + k + 1) * n_cent + i;
            cb = counts[idx];
            sb = row_start[idx];
            if (sb < 0) cb = 0;
        }
        const int maxb = __reduce_max_sync(kFullMask, cb);
        lt_load_row(b, rows, sb, cb, maxb, -1);
        int inter = 0;
        if (maxa <= kLtCap && maxb <= kLtCap) {
#pragma unroll
            for (int ca8 = 0; ca8 < kLtCap / 8; ++ca8) {
                if (ca8 * 8 < maxa) {
#pragma unroll
                    for (int ia = 0; ia < 8; ++ia) {
                        const int av = a[ca8 * 8 + ia];
                        bool hit = false;
#pragma unroll
                        for (int cb8 = 0; cb8 < kLtCap / 8; ++cb8) {
                            if (cb8 * 8 < maxb) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) hit |= (av == b[cb8 * 8 + j]);
                            }
                        }
                        inter += hit ? 1 : 0;
                    }
                }
            }
        } else {
            for (int src = 0; src < 32; ++src) {
                const int ca_s = __shfl_sync(kFullMask, ca, src), cb_s = __shfl_sync(kFullMask, cb, src);
                const long long sa_s = __shfl_sync(kFullMask, sa, src);
                const long long sb_s = __shfl_sync(kFullMask, sb, src);
                int v = 0;
                if (ca_s > 0 && cb_s > 0)
                    v = warp_intersect_unsorted(rows + sa_s, ca_s, rows + sb_s, cb_s, lane);
                if (lane == src) inter = v;
            }
        }
        if (active) out[(long long)i * ld_out + col0 + k0 + k] = lens_value(ca, cb, inter);
#pragma unroll
        for (int j = 0; j < kLtCap; ++j) a[j] = b[j] == -1 ? -2 : b[j];
        ca = cb;
        sa = sb;
        maxa = maxb;
    }
}

__global__ void __launch_bounds__(kLensWarps * 32)
lens_two_csr_kernel(const int* __restrict__ indptr1, const int* __restrict__ indices1,
                    const int* __restrict__ indptr2, const int* __restrict__ indices2, int n_cent,
                    double* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * kLensWarps + warp;
    if (i >= n_cent) return;
    const int s1 = indptr1[i], e1 = indptr1[i + 1], s2 = indptr2[i], e2 = indptr2[i + 1];
    const int na = e1 - s1, nb = e2 - s2;
    int inter = 0;
    if (na > 0 && nb > 0) inter = warp_intersect(indices1 + s1, na, indices2 + s2, nb, lane);
    if (lane == 0) out[i] = lens_value(na, nb, inter);
}

// counts (F, M) int32 -> out[i*ld + col0 + f] float64 through a 32x32 shared-memory tile so that
// both the read and the write are coalesced.
__global__ void __launch_bounds__(256)
counts_to_f64_kernel(const int* __restrict__ counts, int n_frames, int n_cent,
                     double* __restrict__ out, long long ld_out, long long col0) {
    __shared__ int tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int i0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r, i = i0 + tx;
        tile[r][tx] = (f < n_frames && i < n_cent) ? counts[(size_t)f * n_cent + i] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, f = f0 + tx;
        if (i < n_cent && f < n_frames) out[(long long)i * ld_out + col0 + f] = (double)tile[tx][r];
    }
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_lens_pairs(const int32_t* d_indptr, const int64_t* d_frame_off,
                              const int32_t* d_indices, int n_frames, int n_cent, int delay,
                              double* d_out, int64_t ld_out, int64_t col0, void* stream) {
    if (!d_indptr || !d_frame_off || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_cent < 1 || n_cent > 65535 * 1024)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_cent=%d out of range", n_cent);
    if (delay < 1 || delay >= n_frames)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "delay=%d must be in [1, n_frames=%d)", delay,
                         n_frames);
    const int n_pairs = n_frames - delay;
    // grid.y is limited to 65535: fold centres into chunks
    for (int i0 = 0; i0 < n_cent; i0 += 65535) {
        const int ni = n_cent - i0 < 65535 ? n_cent - i0 : 65535;
        dim3 grid((n_pairs + kLensWarps - 1) / kLensWarps, ni);
        lens_pairs_kernel<<<grid, kLensWarps * 32, 0, (cudaStream_t)stream>>>(
            d_indptr + i0, (const long long*)d_frame_off, d_indices, n_pairs, n_cent, delay,
            d_out + (long long)i0 * ld_out, ld_out, col0);
        DSG_LAUNCH_CHECK("lens_pairs_kernel");
    }
    return DSG_OK;
}

extern "C" int dsg_lens_from_two_csr(const int32_t* d_indptr1, const int32_t* d_indices1,
                                     const int32_t* d_indptr2, const int32_t* d_indices2,
                                     int n_cent, double* d_out, void* stream) {
    if (!d_indptr1 || !d_indptr2 || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_cent < 1) return set_error(DSG_ERR_INVALID_ARGUMENT, "n_cent=%d must be >= 1", n_cent);
    const int blocks = (n_cent + kLensWarps - 1) / kLensWarps;
    lens_two_csr_kernel<<<blocks, kLensWarps * 32, 0, (cudaStream_t)stream>>>(
        d_indptr1, d_indices1, d_indptr2, d_indices2, n_cent, d_out);
    DSG_LAUNCH_CHECK("lens_two_csr_kernel");
    return DSG_OK;
}

extern "C" int dsg_counts_to_f64(const int32_t* d_counts, int n_frames, int n_cent, double* d_out,
                                 int64_t ld_out, int64_t col0, void* stream) {
    if (!d_counts || !d_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer");
    if (n_frames < 1 || n_cent < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d, n_cent=%d must be >= 1", n_frames,
                         n_cent);
    dim3 grid((n_cent + 31) / 32, (n_frames + 31) / 32);
    counts_to_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_counts, n_frames, n_cent, d_out,
                                                                ld_out, col0);
    DSG_LAUNCH_CHECK("counts_to_f64_kernel");
    return DSG_OK;
}

extern "C" int dsg_lens_pairs_rows(const int32_t* d_counts, const int64_t* d_row_start,
                                   const int32_t* d_rows, int n_frames, int n_cent, int delay,
                                   double* d_out, int64_t ld_out, int64_t col0, void* stream) {
    if (!d_counts || !d_row_start || !d_rows || !d_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_cent < 1) return set_error(DSG_ERR_INVALID_ARGUMENT, "n_cent=%d out of range", n_cent);
    // bit 30 of delay: dense rows expected -> warp-per-centre kernel also for delay 1
    const int delay_flags = (delay >> 30) & 1;
    delay &= ~(1 << 30);
    if (delay < 1 || delay >= n_frames)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "delay=%d must be in [1, n_frames=%d)", delay,
                         n_frames);
    const int n_pairs = n_frames - delay;
    if (delay == 1 && (delay_flags & 1) == 0) {
        const int per_launch = 65535 * kLtThreads;
        for (int i0 = 0; i0 < n_cent; i0 += per_launch) {
            const int ni = n_cent - i0 < per_launch ? n_cent - i0 : per_launch;
            dim3 grid((n_pairs + kRunPairs - 1) / kRunPairs, (ni + kLtThreads - 1) / kLtThreads);
            lens_rows_thread_kernel<<<grid, kLtThreads, 0, (cudaStream_t)stream>>>(
                d_counts + i0, (const long long*)d_row_start + i0, d_rows, n_pairs, n_cent,
                d_out + (long long)i0 * ld_out, ld_out, col0);
            DSG_LAUNCH_CHECK("lens_rows_thread_kernel");
        }
        return DSG_OK;
    }
    if (delay == 1) {
        const int rows_per_launch = 65535 * kLensWarps;
        for (int i0 = 0; i0 < n_cent; i0 += rows_per_launch) {
            const int ni = n_cent - i0 < rows_per_launch ? n_cent - i0 : rows_per_launch;
            dim3 grid((n_pairs + kRunPairs - 1) / kRunPairs, (ni + kLensWarps - 1) / kLensWarps);
            lens_rows_run_kernel<<<grid, kLensWarps * 32, 0, (cudaStream_t)stream>>>(
                d_counts + i0, (const long long*)d_row_start + i0, d_rows, n_pairs, n_cent,
                d_out + (long long)i0 * ld_out, ld_out, col0);
            DSG_LAUNCH_CHECK("lens_rows_run_kernel");
        }
        return DSG_OK;
    }
    for (int i0 = 0; i0 < n_cent; i0 += 65535) {
        const int ni = n_cent - i0 < 65535 ? n_cent - i0 : 65535;
        dim3 grid((n_pairs + kLensWarps - 1) / kLensWarps, ni);
        lens_pairs_rows_kernel<<<grid, kLensWarps * 32, 0, (cudaStream_t)stream>>>(
            d_counts + i0, (const long long*)d_row_start + i0, d_rows, n_pairs, n_cent, delay,
            d_out + (long long)i0 * ld_out, ld_out, col0);
        DSG_LAUNCH_CHECK("lens_pairs_rows_kernel");
    }
    return DSG_OK;
}
