// timesoap.cu -- fused streaming timeSOAP (sm_100a).
//
// Replaces timesoap/timesoap.py:13-179 (normalize_soap + soap_distance + timesoap: >= 4 full-size
// numpy temporaries) with one pass: a warp owns one particle and a run of frames, streams the
// SOAP vectors once (coalesced 256-byte warp loads), keeps the previous vector and its squared
// norm in registers (a ring of `delay` vectors for delay <= 4) and emits
// sqrt(max(0, 2 - 2 clip(cos))) per frame pair.
// HBM-bound: 8*C bytes read + 8 bytes written per particle-frame.
#include "common.cuh"

namespace dsg {

constexpr int kTsWarps = 4;

struct TsArgs {
    const double* soap;
    int n_particles, n_frames, n_comp;
    long long stride_i, stride_f;
    int delay;
    double* out;
    long long ld_out, col0;
    int seg_len, n_seg;
};

// delay D <= 4, n_comp <= 32*KV: the D previous vectors (and their squared norms) live in
// registers as a ring whose rotation is resolved at compile time (the frame loop is unrolled by D),
// so every vector is read once whatever the delay.
template <int KV, int D>
__global__ void __launch_bounds__(kTsWarps * 32)
timesoap_stream_kernel(const TsArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kTsWarps + warp;
    if (w >= (long long)a.n_particles * a.n_seg) return;
    const int i = (int)(w / a.n_seg), seg = (int)(w % a.n_seg);
    const int n_out = a.n_frames - D;
    const int f_begin = seg * a.seg_len;
    const int f_end = min(n_out, f_begin + a.seg_len);
    const double* __restrict__ base = a.soap + (long long)i * a.stride_i;
    double prev[D][KV];
    double psq[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        // frames f_begin .. f_begin + D - 1 all exist: f_begin < n_out = n_frames - D
        const double* __restrict__ v = base + (long long)(f_begin + d) * a.stride_f;
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < KV; ++k) {
            const int c = lane + 32 * k;
            prev[d][k] = c < a.n_comp ? v[c] : 0.0;
            sq = fma(prev[d][k], prev[d][k], sq);
        }
        psq[d] = warp_sum(sq);
    }
    for (int f0 = f_begin; f0 < f_end; f0 += D) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int f = f0 + d;   // ring slot d holds frame f
            if (f < f_end) {
                const double* __restrict__ v = base + (long long)(f + D) * a.stride_f;
                double cur[KV];
#pragma unroll
                for (int k = 0; k < KV; ++k) {
                    const int c = lane + 32 * k;
                    cur[k] = c < a.n_comp ? v[c] : 0.0;
                }
                double csq = 0.0, dot = 0.0;
#pragma unroll
                for (int k = 0; k < KV; ++k) {
                    csq = fma(cur[k], cur[k], csq);
                    dot = fma(cur[k], prev[d][k], dot);
                    prev[d][k] = cur[k];
                }
                csq = warp_sum(csq);
                dot = warp_sum(dot);
                if (lane == 0)
                    a.out[(long long)i * a.ld_out + a.col0 + f] =
                        soap_dist_from_sums(psq[d], csq, dot);
                psq[d] = csq;
            }
        }
    }
}

// general delay / long vectors: both vectors are read per output (the second read of a vector
// happens `delay` steps later from the same SM and is served by L1/L2).
__global__ void __launch_bounds__(kTsWarps * 32)
timesoap_general_kernel(const TsArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kTsWarps + warp;
    if (w >= (long long)a.n_particles * a.n_seg) return;
    const int i = (int)(w / a.n_seg), seg = (int)(w % a.n_seg);
    const int n_out = a.n_frames - a.delay;
    const int f_begin = seg * a.seg_len;
    const int f_end = min(n_out, f_begin + a.seg_len);
    const double* __restrict__ base = a.soap + (long long)i * a.stride_i;
    for (int f = f_begin; f < f_end; ++f) {
        const double* __restrict__ v1 = base + (long long)f * a.stride_f;
        const double* __restrict__ v2 = base + (long long)(f + a.delay) * a.stride_f;
        double s1 = 0.0, s2 = 0.0, dot = 0.0;
        for (int c = lane; c < a.n_comp; c += 32) {
            const double x = v1[c], y = v2[c];
            s1 = fma(x, x, s1);
            s2 = fma(y, y, s2);
            dot = fma(x, y, dot);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        dot = warp_sum(dot);
        if (lane == 0) a.out[(long long)i * a.ld_out + a.col0 + f] = soap_dist_from_sums(s1, s2, dot);
    }
}

// normalize_soap (timesoap.py:54-58): v / |v| where |v| > 0, else the zero vector.
__global__ void __launch_bounds__(kTsWarps * 32)
normalize_kernel(const double* __restrict__ in, double* __restrict__ out, long long n_vec,
                 int n_comp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long v = (long long)blockIdx.x * kTsWarps + warp;
    if (v >= n_vec) return;
    const double* __restrict__ src = in + v * n_comp;
    double* __restrict__ dst = out + v * n_comp;
    double s = 0.0;
    for (int c = lane; c < n_comp; c += 32) s = fma(src[c], src[c], s);
    s = warp_sum(s);
    const double norm = sqrt(s);
    for (int c = lane; c < n_comp; c += 32) dst[c] = norm > 0.0 ? src[c] / norm : 0.0;
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_normalize_soap(const double* d_in, double* d_out, int64_t n_vectors,
                                  int n_comp, void* stream) {
    if (!d_in || !d_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer");
    if (n_vectors < 0 || n_comp < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_vectors=%lld, n_comp=%d out of range",
                         (long long)n_vectors, n_comp);
    if (n_vectors == 0) return DSG_OK;
    const long long blocks = (n_vectors + kTsWarps - 1) / kTsWarps;
    if (blocks > 0x7fffffffLL) return set_error(DSG_ERR_OVERFLOW, "grid too large");
    normalize_kernel<<<(unsigned)blocks, kTsWarps * 32, 0, (cudaStream_t)stream>>>(
        d_in, d_out, n_vectors, n_comp);
    DSG_LAUNCH_CHECK("normalize_kernel");
    return DSG_OK;
}

extern "C" int dsg_timesoap(const double* d_soap, int n_particles, int n_frames, int n_comp,
                            int64_t stride_i, int64_t stride_f, int delay, double* d_out,
                            int64_t ld_out, int64_t col0, void* stream) {
    if (!d_soap || !d_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer");
    if (n_particles < 1 || n_comp < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_particles=%d, n_comp=%d must be >= 1",
                         n_particles, n_comp);
    if (delay < 1 || delay >= n_frames)  // timesoap.py:175-177
        return set_error(DSG_ERR_INVALID_ARGUMENT, "delay value outside bounds");
    TsArgs a;
    a.soap = d_soap;
    a.n_particles = n_particles;
    a.n_frames = n_frames;
    a.n_comp = n_comp;
    a.stride_i = stride_i;
    a.stride_f = stride_f;
    a.delay = delay;
    a.out = d_out;
    a.ld_out = ld_out;
    a.col0 = col0;
    const int n_out = n_frames - delay;
    // enough warps to fill 148 SMs several times over, segments not shorter than 8 outputs
    const long long target_warps = 8LL * kNumSMs * 32;
    long long n_seg = (target_warps + n_particles - 1) / n_particles;
    if (n_seg < 1) n_seg = 1;
    int seg_len = (int)((n_out + n_seg - 1) / n_seg);
    if (seg_len < 8) seg_len = 8;
    if (seg_len > n_out) seg_len = n_out;
    a.seg_len = seg_len;
    a.n_seg = (n_out + seg_len - 1) / seg_len;
    const long long warps = (long long)n_particles * a.n_seg;
    const long long blocks = (warps + kTsWarps - 1) / kTsWarps;
    if (blocks > 0x7fffffffLL) return set_error(DSG_ERR_OVERFLOW, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int kv = (n_comp + 31) / 32;
    if (delay <= 4 && kv <= (delay == 1 ? 16 : 12)) {
#define DSG_TS(KV, D) timesoap_stream_kernel<KV, D><<<(unsigned)blocks, kTsWarps * 32, 0, st>>>(a)
#define DSG_TS_KV(D)                                                                             \
    do {                                                                                         \
        if (kv <= 2) DSG_TS(2, D); else if (kv <= 4) DSG_TS(4, D);                               \
        else if (kv <= 8) DSG_TS(8, D); else DSG_TS(12, D);                                      \
    } while (0)
        if (delay == 1) {
            if (kv <= 12) DSG_TS_KV(1); else DSG_TS(16, 1);
        } else if (delay == 2) DSG_TS_KV(2);
        else if (delay == 3) DSG_TS_KV(3);
        else DSG_TS_KV(4);
#undef DSG_TS_KV
#undef DSG_TS
        DSG_LAUNCH_CHECK("timesoap_stream_kernel");
    } else {
        timesoap_general_kernel<<<(unsigned)blocks, kTsWarps * 32, 0, st>>>(a);
        DSG_LAUNCH_CHECK("timesoap_general_kernel");
    }
    return DSG_OK;
}
