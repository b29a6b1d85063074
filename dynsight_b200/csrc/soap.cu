// soap.cu -- SOAP power spectrum (GTO radial basis, sigma = 1) batched over frames (sm_100a).
//
// Replaces the DScribe call of soapify/saponify.py:98-122.  One warp owns one centre:
//
//   gather   every periodic image within R = r_cut + sqrt(-2 ln 1e-3) is found through a cell list
//            (cell edge >= R; axes shorter than 3R fall back to an explicit image loop; a
//            non-periodic frame uses an open bounding-box grid); hits are ballot-compacted into a
//            32-entry displacement list in shared memory
//   phase B  two lanes per neighbour evaluate the un-normalised regular solid harmonics
//            r^l Y_lm (polynomial recursions: no sqrt, no division, exact at r = 0)
//   phase C  lane (g, k) owns radial index k and a load-balanced group g of angular momenta; it
//            evaluates its Gaussians exp(-gamma_lk r^2) and accumulates G_lkm in registers
//   epilogue Loewdin mixing c = (beta * prefactor) G in shared memory, then one lane per output
//            contracts over m, scales and writes the spectrum coalesced.
//
// float64 throughout: the Loewdin matrices have entries of order 1e3 and cancel, so primitive
// sums need double precision to meet 1e-5 relative (SURVEY.md section 7 "hard parts").
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace dsg {

constexpr unsigned kFullS = 0xffffffffu;
constexpr int kMaxL = 12;
constexpr int kMaxN = 32;
constexpr int kMaxSpecies = 16;
constexpr int kKBlock = 8;       // radial indices handled per pass (lanes: 4 groups x 8)
constexpr int kGroups = 4;
constexpr int kMaxItems = 4;
constexpr int kHalfChunk = 16;   // neighbours per phase-B/C pass
constexpr int kOpenCellsPerAxis = 32;

enum AxisMode : int { kWide = 0, kNarrow = 1, kOpen = 2 };

struct SoapFrame {
    double box[3];
    double origin[3];
    double inv_edge[3];   // cells per unit length
    int nc[3];
    int mode[3];
    int nimg[3];
    int cell_base;
    unsigned lo_bits[3];  // open mode: encoded min / max (atomics)
    unsigned hi_bits[3];
};

struct SoapPlan {
    int n_max, l_max, n_species, n_feat;
    int n_lm;          // (l_max+1)^2
    int row_stride;    // P: doubles per neighbour in the harmonics buffer (odd)
    int kblocks;
    int item_l[kGroups][kMaxItems];  // -1 = unused
};

struct SoapArgs {
    const double* wpos;     // (F, N, 3) wrapped / shifted coordinates
    const double* sx;       // cell-sorted SoA
    const double* sy;
    const double* sz;
    const int* sspec;
    const int* cell_start;
    const SoapFrame* frames;
    const int* centres;
    int n_atoms, n_cent;
    double r2max;
    const double* gamma;    // (L+1, n)
    const double* bmat;     // (L+1, n, n)  beta * prefactor
    const double* wlm;      // (n_lm)
    const int* lm_l;        // (n_lm) -> l
    const int* decode;      // (n_feat)
    double* out;
    long long stride_i, stride_f;
    SoapPlan plan;
};

// ---- orderable encoding of floats for atomicMin/Max ---------------------------------------
__device__ __forceinline__ unsigned enc_float(float v) {
    unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_float(unsigned e) {
    unsigned b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
    return __uint_as_float(b);
}

template <typename T>
__global__ void __launch_bounds__(256)
soap_bounds_kernel(const T* __restrict__ pos, int n_atoms, SoapFrame* __restrict__ fr) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (i < n_atoms) {
        const size_t g = ((size_t)f * n_atoms + i) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double x = (double)pos[g + k];
            lo[k] = __double2float_rd(x);
            hi[k] = __double2float_ru(x);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(kFullS, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFullS, hi[k], off));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&fr[f].lo_bits[k], enc_float(lo[k]));
            atomicMax(&fr[f].hi_bits[k], enc_float(hi[k]));
        }
    }
}

__global__ void soap_open_setup_kernel(SoapFrame* __restrict__ fr, int n_frames, double r_cell) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    SoapFrame* p = fr + f;
    for (int k = 0; k < 3; ++k) {
        const double lo = (double)dec_float(p->lo_bits[k]);
        const double hi = (double)dec_float(p->hi_bits[k]);
        const double ext = hi - lo;
        int nc = 1;
        if (ext > 0.0) {
            const double q = floor(ext / r_cell);
            nc = q < 1.0 ? 1 : (q > (double)kOpenCellsPerAxis ? kOpenCellsPerAxis : (int)q);
        }
        p->origin[k] = lo;
        p->box[k] = 0.0;
        p->nc[k] = nc;
        p->inv_edge[k] = ext > 0.0 ? (double)nc / ext : 0.0;
        p->mode[k] = kOpen;
        p->nimg[k] = 0;
    }
}

__device__ __forceinline__ void soap_locate(const SoapFrame* __restrict__ p, double x, double y,
                                            double z, double* w, int* c) {
    const double in[3] = {x, y, z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = in[k];
        int ck;
        if (p->mode[k] == kOpen) {
            ck = (int)((v - p->origin[k]) * p->inv_edge[k]);
        } else {
            const double b = p->box[k];
            v -= floor(v / b) * b;          // wrap into [0, box)
            if (v >= b) v -= b;
            if (v < 0.0) v = 0.0;
            ck = (int)(v * p->inv_edge[k]);
        }
        const int nc = p->nc[k];
        ck = ck < 0 ? 0 : (ck >= nc ? nc - 1 : ck);
        w[k] = v;
        c[k] = ck;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
soap_bin_kernel(const T* __restrict__ pos, int n_atoms, const SoapFrame* __restrict__ fr,
                double* __restrict__ wpos, unsigned* __restrict__ cell_id,
                int* __restrict__ cell_cnt) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const SoapFrame* p = fr + f;
    const size_t a = (size_t)f * n_atoms + i;
    double w[3];
    int c[3];
    soap_locate(p, (double)pos[a * 3], (double)pos[a * 3 + 1], (double)pos[a * 3 + 2], w, c);
    wpos[a * 3] = w[0];
    wpos[a * 3 + 1] = w[1];
    wpos[a * 3 + 2] = w[2];
    const int cell = (c[0] * p->nc[1] + c[1]) * p->nc[2] + c[2];
    cell_id[a] = (unsigned)cell;
    atomicAdd(&cell_cnt[p->cell_base + cell], 1);
}

__global__ void __launch_bounds__(256)
soap_scatter_kernel(const double* __restrict__ wpos, const int* __restrict__ species, int n_atoms,
                    const SoapFrame* __restrict__ fr, const unsigned* __restrict__ cell_id,
                    int* __restrict__ cell_cnt, const int* __restrict__ cell_start,
                    double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz,
                    int* __restrict__ sspec) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const size_t a = (size_t)f * n_atoms + i;
    const int gc = fr[f].cell_base + (int)cell_id[a];
    const int slot = cell_start[gc] + atomicSub(&cell_cnt[gc], 1) - 1;
    sx[slot] = wpos[a * 3];
    sy[slot] = wpos[a * 3 + 1];
    sz[slot] = wpos[a * 3 + 2];
    sspec[slot] = species[i];
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
__constant__ double c_inv_int[2 * kMaxL + 2];  // 1/i

template <int N>
struct Acc {
    double v[N > 0 ? N : 1];
};

// Phase B: un-normalised regular solid harmonics of one displacement, m of one parity.
//   Rt[l*l + l + m] = S_l^m(z, r2) * A_m(x, y),   Rt[l*l + l - m] = S_l^m * B_m   (m > 0)
//   A_m + i B_m = (x + i y)^m ;  S_m^m = (2m-1)!!, S_{m+1}^m = (2m+1) z S_m^m,
//   S_l^m = ((2l-1) z S_{l-1}^m - (l+m-1) r2 S_{l-2}^m) / (l-m)
__device__ __forceinline__ void solid_harmonics(double x, double y, double z, double r2, int l_max,
                                                int parity, double* __restrict__ rt) {
    double am = 1.0, bm = 0.0, dfact = 1.0;
    for (int m = 0; m <= l_max; ++m) {
        if (m > 0) {
            const double an = am * x - bm * y;
            bm = am * y + bm * x;
            am = an;
            dfact *= (double)(2 * m - 1);
        }
        if ((m & 1) != parity) continue;
        double s0 = dfact;  // S_m^m
        rt[m * m + 2 * m] = s0 * am;
        if (m > 0) rt[m * m] = s0 * bm;
        if (m + 1 > l_max) continue;
        double s1 = (double)(2 * m + 1) * z * s0;  // S_{m+1}^m
        {
            const int l = m + 1, o = l * l + l;
            rt[o + m] = s1 * am;
            if (m > 0) rt[o - m] = s1 * bm;
        }
        for (int l = m + 2; l <= l_max; ++l) {
            const double s2 =
                ((double)(2 * l - 1) * z * s1 - (double)(l + m - 1) * r2 * s0) * c_inv_int[l - m];
            const int o = l * l + l;
            rt[o + m] = s2 * am;
            if (m > 0) rt[o - m] = s2 * bm;
            s0 = s1;
            s1 = s2;
        }
    }
}

template <int S0, int S1, int S2, int S3>
__global__ void __launch_bounds__(128)
soap_kernel(const SoapArgs a) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    const int ci = blockIdx.x * warps + warp;
    const int f = blockIdx.y;
    if (ci >= a.n_cent) return;  // warps are independent: only __syncwarp below

    const SoapPlan& pl = a.plan;
    const int n = pl.n_max, L = pl.l_max, P = pl.row_stride, n_lm = pl.n_lm;
    const int c_doubles = pl.n_species * n_lm * n;
    const int per_warp = 32 * 4 + kHalfChunk * P + c_doubles;
    double* nbuf = smem + (size_t)warp * per_warp;   // [32][4] dx dy dz r2
    double* rbuf = nbuf + 32 * 4;                      // [16][P]
    double* cbuf = rbuf + kHalfChunk * P;              // [S][n_lm][n]

    const SoapFrame* __restrict__ fr = a.frames + f;
    const int atom = a.centres[ci];
    const double* cp = a.wpos + ((size_t)f * a.n_atoms + atom) * 3;
    const double xc = cp[0], yc = cp[1], zc = cp[2];
    int cc[3];
    {
        double w[3];
        // the centre's coordinates are already wrapped: locate only recomputes the cell
        soap_locate(fr, xc, yc, zc, w, cc);
    }
    int ne[3], mode[3], nc[3], nimg[3];
    double box[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mode[k] = fr->mode[k];
        nc[k] = fr->nc[k];
        nimg[k] = fr->nimg[k];
        box[k] = fr->box[k];
        ne[k] = mode[k] == kNarrow ? 2 * nimg[k] + 1 : 3;
    }
    const int cell_base = fr->cell_base;

    const int g = lane >> 3, kk = lane & 7;
    const unsigned lt_mask = (1u << lane) - 1u;
    int il[kMaxItems], ioff[kMaxItems];
#pragma unroll
    for (int t = 0; t < kMaxItems; ++t) {
        il[t] = pl.item_l[g][t];
        ioff[t] = il[t] >= 0 ? il[t] * il[t] : 0;
    }

    for (int z_sp = 0; z_sp < pl.n_species; ++z_sp) {
        for (int kb = 0; kb < pl.kblocks; ++kb) {
            const int k = kb * kKBlock + kk;
            const bool kact = k < n;
            double gam[kMaxItems];
#pragma unroll
            for (int t = 0; t < kMaxItems; ++t)
                gam[t] = (kact && il[t] >= 0) ? a.gamma[il[t] * n + k] : 0.0;
            Acc<S0> a0;
            Acc<S1> a1;
            Acc<S2> a2;
            Acc<S3> a3;
#pragma unroll
            for (int s = 0; s < S0; ++s) a0.v[s] = 0.0;
#pragma unroll
            for (int s = 0; s < S1; ++s) a1.v[s] = 0.0;
#pragma unroll
            for (int s = 0; s < S2; ++s) a2.v[s] = 0.0;
#pragma unroll
            for (int s = 0; s < S3; ++s) a3.v[s] = 0.0;

            auto process = [&](int fill) {
                for (int h0 = 0; h0 < fill; h0 += kHalfChunk) {
                    const int nh = min(kHalfChunk, fill - h0);
                    {   // phase B
                        const int jj = lane >> 1;
                        if (jj < nh) {
                            const double* d = nbuf + (h0 + jj) * 4;
                            solid_harmonics(d[0], d[1], d[2], d[3], L, lane & 1, rbuf + jj * P);
                        }
                    }
                    __syncwarp();
                    for (int jj = 0; jj < nh; ++jj) {   // phase C
                        const double r2 = nbuf[(h0 + jj) * 4 + 3];
                        const double* __restrict__ rt = rbuf + jj * P;
                        if (S0 > 0) {
                            const double e = exp(-gam[0] * r2);
                            const double* __restrict__ r = rt + ioff[0];
#pragma unroll
                            for (int s = 0; s < S0; ++s) a0.v[s] = fma(e, r[s], a0.v[s]);
                        }
                        if (S1 > 0) {
                            const double e = exp(-gam[1] * r2);
                            const double* __restrict__ r = rt + ioff[1];
#pragma unroll
                            for (int s = 0; s < S1; ++s) a1.v[s] = fma(e, r[s], a1.v[s]);
                        }
                        if (S2 > 0) {
                            const double e = exp(-gam[2] * r2);
                            const double* __restrict__ r = rt + ioff[2];
#pragma unroll
                            for (int s = 0; s < S2; ++s) a2.v[s] = fma(e, r[s], a2.v[s]);
                        }
                        if (S3 > 0) {
                            const double e = exp(-gam[3] * r2);
                            const double* __restrict__ r = rt + ioff[3];
#pragma unroll
                            for (int s = 0; s < S3; ++s) a3.v[s] = fma(e, r[s], a3.v[s]);
                        }
                    }
                    __syncwarp();
                }
            };

            // ---- gather ----
            int fill = 0;
            for (int ex = 0; ex < ne[0]; ++ex) {
                int cx;
                double shx;
                if (mode[0] == kNarrow) { cx = 0; shx = (double)(ex - nimg[0]) * box[0]; }
                else {
                    cx = cc[0] + ex - 1; shx = 0.0;
                    if (mode[0] == kWide) {
                        if (cx < 0) { cx += nc[0]; shx = -box[0]; }
                        else if (cx >= nc[0]) { cx -= nc[0]; shx = box[0]; }
                    } else if (cx < 0 || cx >= nc[0]) continue;
                }
                for (int ey = 0; ey < ne[1]; ++ey) {
                    int cy;
                    double shy;
                    if (mode[1] == kNarrow) { cy = 0; shy = (double)(ey - nimg[1]) * box[1]; }
                    else {
                        cy = cc[1] + ey - 1; shy = 0.0;
                        if (mode[1] == kWide) {
                            if (cy < 0) { cy += nc[1]; shy = -box[1]; }
                            else if (cy >= nc[1]) { cy -= nc[1]; shy = box[1]; }
                        } else if (cy < 0 || cy >= nc[1]) continue;
                    }
                    for (int ez = 0; ez < ne[2]; ++ez) {
                        int cz;
                        double shz;
                        if (mode[2] == kNarrow) { cz = 0; shz = (double)(ez - nimg[2]) * box[2]; }
                        else {
                            cz = cc[2] + ez - 1; shz = 0.0;
                            if (mode[2] == kWide) {
                                if (cz < 0) { cz += nc[2]; shz = -box[2]; }
                                else if (cz >= nc[2]) { cz -= nc[2]; shz = box[2]; }
                            } else if (cz < 0 || cz >= nc[2]) continue;
                        }
                        const int cell = cell_base + (cx * nc[1] + cy) * nc[2] + cz;
                        const int qb = a.cell_start[cell], qe = a.cell_start[cell + 1];
                        for (int q0 = qb; q0 < qe; q0 += 32) {
                            const int q = q0 + lane;
                            bool hit = false;
                            double dx = 0.0, dy = 0.0, dz = 0.0, r2 = 0.0;
                            if (q < qe) {
                                dx = a.sx[q] + shx - xc;
                                dy = a.sy[q] + shy - yc;
                                dz = a.sz[q] + shz - zc;
                                r2 = dx * dx + dy * dy + dz * dz;
                                hit = r2 <= a.r2max && a.sspec[q] == z_sp;
                            }
                            const unsigned m = __ballot_sync(kFullS, hit);
                            if (m == 0u) continue;
                            const int cnt = __popc(m), rank = __popc(m & lt_mask);
                            const int room = 32 - fill;
                            if (hit && rank < room) {
                                double* d = nbuf + (fill + rank) * 4;
                                d[0] = dx; d[1] = dy; d[2] = dz; d[3] = r2;
                            }
                            if (cnt >= room) {
                                __syncwarp();
                                process(32);
                                if (hit && rank >= room) {
                                    double* d = nbuf + (rank - room) * 4;
                                    d[0] = dx; d[1] = dy; d[2] = dz; d[3] = r2;
                                }
                                fill = cnt - room;
                            } else {
                                fill += cnt;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (fill > 0) process(fill);

            // ---- registers -> shared G[z][lm][k] ----
            if (kact) {
                double* cz = cbuf + (size_t)z_sp * n_lm * n;
                if (S0 > 0 && il[0] >= 0) {
#pragma unroll
                    for (int s = 0; s < S0; ++s)
                        if (s < 2 * il[0] + 1) cz[(ioff[0] + s) * n + k] = a0.v[s];
                }
                if (S1 > 0 && il[1] >= 0) {
#pragma unroll
                    for (int s = 0; s < S1; ++s)
                        if (s < 2 * il[1] + 1) cz[(ioff[1] + s) * n + k] = a1.v[s];
                }
                if (S2 > 0 && il[2] >= 0) {
#pragma unroll
                    for (int s = 0; s < S2; ++s)
                        if (s < 2 * il[2] + 1) cz[(ioff[2] + s) * n + k] = a2.v[s];
                }
                if (S3 > 0 && il[3] >= 0) {
#pragma unroll
                    for (int s = 0; s < S3; ++s)
                        if (s < 2 * il[3] + 1) cz[(ioff[3] + s) * n + k] = a3.v[s];
                }
            }
            __syncwarp();
        }
    }

    // ---- Loewdin mixing in place: c[z][lm][n'] = sum_k bmat[l][n'][k] G[z][lm][k] ----
    for (int col = lane; col < pl.n_species * n_lm; col += 32) {
        const int lm = col % n_lm;
        const int l = a.lm_l[lm];
        double* v = cbuf + (size_t)col * n;
        double gk[kMaxN];
#pragma unroll
        for (int k = 0; k < kMaxN; ++k) gk[k] = k < n ? v[k] : 0.0;
        const double* __restrict__ bm = a.bmat + (size_t)l * n * n;
        for (int np = 0; np < n; ++np) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kMaxN; ++k)
                if (k < n) s = fma(bm[np * n + k], gk[k], s);
            v[np] = s;
        }
    }
    __syncwarp();

    // ---- power spectrum ----
    double* __restrict__ out = a.out + (long long)ci * a.stride_i + (long long)f * a.stride_f;
    for (int c = lane; c < pl.n_feat; c += 32) {
        const int code = a.decode[c];
        const int zi = code & 15, zj = (code >> 4) & 15, l = (code >> 8) & 15;
        const int na = (code >> 12) & 31, nb = (code >> 17) & 31;
        const double* __restrict__ ca = cbuf + ((size_t)zi * n_lm + l * l) * n + na;
        const double* __restrict__ cb = cbuf + ((size_t)zj * n_lm + l * l) * n + nb;
        const double* __restrict__ w = a.wlm + l * l;
        double s = 0.0;
        for (int mi = 0; mi < 2 * l + 1; ++mi) s = fma(w[mi] * ca[mi * n], cb[mi * n], s);
        out[c] = s;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct SoapLayout {
    long long total_cells = 0;
    bool periodic = false;
    size_t off_frames = 0, off_wpos = 0, off_cell_id = 0, off_cnt = 0, off_start = 0;
    size_t off_sx = 0, off_sy = 0, off_sz = 0, off_sspec = 0, off_tile = 0;
    size_t off_gamma = 0, off_bmat = 0, off_wlm = 0, off_lml = 0, off_decode = 0;
    size_t bytes = 0;
};

static double soap_radius(double r_cut) { return r_cut + std::sqrt(-2.0 * std::log(1e-3)); }

static int soap_layout(int n_frames, int n_atoms, int n_cent, int n_species, int n_max, int l_max,
                       const double* h_box, double r_cut, SoapLayout* L,
                       std::vector<SoapFrame>* frames) {
    if (n_frames < 1 || n_frames > 65535)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d must be in [1, 65535] per batch",
                         n_frames);
    if (n_atoms < 1 || n_cent < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_atoms=%d, n_cent=%d must be >= 1", n_atoms,
                         n_cent);
    if (n_max < 1 || n_max > kMaxN || l_max < 0 || l_max > kMaxL)
        return set_error(DSG_ERR_UNSUPPORTED, "n_max=%d (1..%d), l_max=%d (0..%d) out of range",
                         n_max, kMaxN, l_max, kMaxL);
    if (n_species < 1 || n_species > kMaxSpecies)
        return set_error(DSG_ERR_UNSUPPORTED, "n_species=%d must be in [1, %d]", n_species,
                         kMaxSpecies);
    if (!(r_cut > 1.0))
        return set_error(DSG_ERR_INVALID_ARGUMENT, "r_cut=%g must be > 1 (DScribe GTO basis)",
                         r_cut);
    if ((long long)n_frames * n_atoms >= 0x7fffffffLL)
        return set_error(DSG_ERR_OVERFLOW, "n_frames*n_atoms must stay below 2^31 per batch");
    const double r_cell = soap_radius(r_cut) * (1.0 + 1e-9);
    L->periodic = h_box != nullptr;
    long long total = 0;
    if (frames) frames->assign(n_frames, SoapFrame{});
    for (int f = 0; f < n_frames; ++f) {
        long long cells = 1;
        SoapFrame fr{};
        if (h_box) {
            for (int k = 0; k < 3; ++k) {
                const double b = h_box[3 * f + k];
                if (!(b > 0.0) || !std::isfinite(b))
                    return set_error(DSG_ERR_INVALID_ARGUMENT,
                                     "box[%d][%d]=%g: a periodic frame needs a positive box", f, k,
                                     b);
                const double q = std::floor(b / r_cell);
                fr.box[k] = b;
                fr.origin[k] = 0.0;
                if (q >= 3.0) {
                    fr.mode[k] = kWide;
                    fr.nc[k] = (int)std::min(q, 1024.0);
                    fr.nimg[k] = 0;
                } else {
                    fr.mode[k] = kNarrow;
                    fr.nc[k] = 1;
                    fr.nimg[k] = (int)std::floor(soap_radius(r_cut) / b) + 1;
                }
                fr.inv_edge[k] = (double)fr.nc[k] / b;
                cells *= fr.nc[k];
            }
        } else {
            cells = (long long)kOpenCellsPerAxis * kOpenCellsPerAxis * kOpenCellsPerAxis;
            for (int k = 0; k < 3; ++k) {
                fr.lo_bits[k] = 0xffffffffu;
                fr.hi_bits[k] = 0u;
            }
        }
        fr.cell_base = (int)total;
        total += cells;
        if (total >= 0x7fffffffLL)
            return set_error(DSG_ERR_OVERFLOW, "cell grid too large; use fewer frames per batch");
        if (frames) (*frames)[f] = fr;
    }
    L->total_cells = total;
    const int n_lm = (l_max + 1) * (l_max + 1);
    const int n_feat = dsg_soap_n_features(n_species, n_max, l_max);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += align_up(bytes);
        return r;
    };
    const size_t fa = (size_t)n_frames * n_atoms;
    L->off_frames = take(sizeof(SoapFrame) * n_frames);
    L->off_wpos = take(8 * 3 * fa);
    L->off_cell_id = take(4 * fa);
    L->off_cnt = take(4 * (size_t)total);
    L->off_start = take(4 * ((size_t)total + 1));
    L->off_sx = take(8 * fa);
    L->off_sy = take(8 * fa);
    L->off_sz = take(8 * fa);
    L->off_sspec = take(4 * fa);
    L->off_tile = take(4 * (((size_t)total + kScanTileItems - 1) / kScanTileItems + 1));
    L->off_gamma = take(8 * (size_t)(l_max + 1) * n_max);
    L->off_bmat = take(8 * (size_t)(l_max + 1) * n_max * n_max);
    L->off_wlm = take(8 * (size_t)n_lm);
    L->off_lml = take(4 * (size_t)n_lm);
    L->off_decode = take(4 * (size_t)n_feat);
    L->bytes = o;
    return DSG_OK;
}

static void make_plan(int n_species, int n_max, int l_max, int s0, SoapPlan* pl) {
    pl->n_max = n_max;
    pl->l_max = l_max;
    pl->n_species = n_species;
    pl->n_feat = dsg_soap_n_features(n_species, n_max, l_max);
    pl->n_lm = (l_max + 1) * (l_max + 1);
    int p = pl->n_lm + s0;  // over-read padding: item rows are read S_t wide
    if ((p & 1) == 0) ++p;  // odd stride: conflict-free phase-B stores
    pl->row_stride = p;
    pl->kblocks = (n_max + kKBlock - 1) / kKBlock;
    // longest-processing-time packing of l = l_max..0 (size 2l+1) into 4 lane groups
    int load[kGroups] = {0, 0, 0, 0}, cnt[kGroups] = {0, 0, 0, 0};
    for (int g = 0; g < kGroups; ++g)
        for (int t = 0; t < kMaxItems; ++t) pl->item_l[g][t] = -1;
    for (int l = l_max; l >= 0; --l) {
        int best = 0;
        for (int g = 1; g < kGroups; ++g)
            if (load[g] < load[best]) best = g;
        pl->item_l[best][cnt[best]++] = l;
        load[best] += 2 * l + 1;
    }
}

template <typename T>
static int soap_impl(const T* d_pos, const int32_t* d_species, const int32_t* d_centres,
                     int n_frames, int n_atoms, int n_cent, int n_species, const double* h_box,
                     double r_cut, int n_max, int l_max, const double* h_alphas,
                     const double* h_betas, void* d_ws, size_t ws_bytes, double* d_out,
                     int64_t stride_i, int64_t stride_f, cudaStream_t st) {
    SoapLayout L;
    std::vector<SoapFrame> frames;
    int rc = soap_layout(n_frames, n_atoms, n_cent, n_species, n_max, l_max, h_box, r_cut, &L,
                         &frames);
    if (rc) return rc;
    if (!d_ws || ws_bytes < L.bytes)
        return set_error(DSG_ERR_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed",
                         ws_bytes, L.bytes);
    if (!d_pos || !d_species || !d_centres || !d_out || !h_alphas || !h_betas)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    char* ws = (char*)d_ws;
    SoapFrame* d_frames = (SoapFrame*)(ws + L.off_frames);
    double* wpos = (double*)(ws + L.off_wpos);
    unsigned* cell_id = (unsigned*)(ws + L.off_cell_id);
    int* cnt = (int*)(ws + L.off_cnt);
    int* start = (int*)(ws + L.off_start);
    double* sx = (double*)(ws + L.off_sx);
    double* sy = (double*)(ws + L.off_sy);
    double* sz = (double*)(ws + L.off_sz);
    int* sspec = (int*)(ws + L.off_sspec);
    int* tile = (int*)(ws + L.off_tile);

    // ---- basis tables (host float64) ----
    const double eta = 0.5;
    const int n_lm = (l_max + 1) * (l_max + 1);
    std::vector<double> gamma((size_t)(l_max + 1) * n_max), bmat((size_t)(l_max + 1) * n_max * n_max),
        wlm(n_lm);
    std::vector<int> lml(n_lm);
    const double pi = 3.14159265358979323846;
    for (int l = 0; l <= l_max; ++l) {
        for (int k = 0; k < n_max; ++k) {
            const double al = h_alphas[l * n_max + k];
            gamma[l * n_max + k] = al * eta / (al + eta);
            const double pref = std::pow(pi, 1.5) * std::pow(eta / (eta + al), l) *
                                std::pow(eta + al, -1.5);
            for (int np = 0; np < n_max; ++np)
                bmat[((size_t)l * n_max + np) * n_max + k] =
                    h_betas[((size_t)l * n_max + np) * n_max + k] * pref;
        }
        for (int mi = 0; mi <= 2 * l; ++mi) {
            const int am = std::abs(mi - l);
            double ratio = 1.0;  // (l-|m|)! / (l+|m|)!
            for (int i = l - am + 1; i <= l + am; ++i) ratio /= (double)i;
            const double n2 = (2 * l + 1) / (4.0 * pi) * ratio * (am > 0 ? 2.0 : 1.0);
            wlm[l * l + mi] = pi * std::sqrt(8.0 / (2 * l + 1)) * n2;
            lml[l * l + mi] = l;
        }
    }
    const int n_feat = dsg_soap_n_features(n_species, n_max, l_max);
    std::vector<int> decode(n_feat);
    {
        int c = 0;
        for (int zi = 0; zi < n_species; ++zi)
            for (int zj = zi; zj < n_species; ++zj)
                for (int l = 0; l <= l_max; ++l)
                    for (int na = 0; na < n_max; ++na)
                        for (int nb = (zi == zj ? na : 0); nb < n_max; ++nb)
                            decode[c++] = zi | (zj << 4) | (l << 8) | (na << 12) | (nb << 17);
        if (c != n_feat) return set_error(DSG_ERR_INVALID_ARGUMENT, "feature count mismatch");
    }
    double inv_int[2 * kMaxL + 2];
    inv_int[0] = 0.0;
    for (int i = 1; i < 2 * kMaxL + 2; ++i) inv_int[i] = 1.0 / (double)i;
    DSG_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_inv_int, inv_int, sizeof(inv_int), 0,
                                           cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_gamma, gamma.data(), 8 * gamma.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_bmat, bmat.data(), 8 * bmat.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_wlm, wlm.data(), 8 * wlm.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_lml, lml.data(), 4 * lml.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_decode, decode.data(), 4 * decode.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(d_frames, frames.data(), sizeof(SoapFrame) * n_frames,
                                   cudaMemcpyHostToDevice, st));

    // ---- cell list ----
    dim3 agrid((n_atoms + 255) / 256, n_frames);
    const double radius = soap_radius(r_cut);
    if (!L.periodic) {
        soap_bounds_kernel<T><<<agrid, 256, 0, st>>>(d_pos, n_atoms, d_frames);
        DSG_LAUNCH_CHECK("soap_bounds_kernel");
        soap_open_setup_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(d_frames, n_frames,
                                                                      radius * (1.0 + 1e-9));
        DSG_LAUNCH_CHECK("soap_open_setup_kernel");
    }
    DSG_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 4 * (size_t)L.total_cells, st));
    soap_bin_kernel<T><<<agrid, 256, 0, st>>>(d_pos, n_atoms, d_frames, wpos, cell_id, cnt);
    DSG_LAUNCH_CHECK("soap_bin_kernel");
    rc = run_scan(cnt, L.total_cells, 1, start, tile, nullptr, st);
    if (rc) return rc;
    soap_scatter_kernel<<<agrid, 256, 0, st>>>(wpos, d_species, n_atoms, d_frames, cell_id, cnt,
                                               start, sx, sy, sz, sspec);
    DSG_LAUNCH_CHECK("soap_scatter_kernel");

    // ---- main kernel ----
    SoapArgs a;
    a.wpos = wpos;
    a.sx = sx;
    a.sy = sy;
    a.sz = sz;
    a.sspec = sspec;
    a.cell_start = start;
    a.frames = d_frames;
    a.centres = d_centres;
    a.n_atoms = n_atoms;
    a.n_cent = n_cent;
    a.r2max = radius * radius;
    a.gamma = (const double*)(ws + L.off_gamma);
    a.bmat = (const double*)(ws + L.off_bmat);
    a.wlm = (const double*)(ws + L.off_wlm);
    a.lm_l = (const int*)(ws + L.off_lml);
    a.decode = (const int*)(ws + L.off_decode);
    a.out = d_out;
    a.stride_i = stride_i;
    a.stride_f = stride_f;

    // template instantiation by l_max (slot widths of the packed items, see make_plan)
    int s0;
    if (l_max <= 4) s0 = 9;
    else if (l_max <= 6) s0 = 13;
    else if (l_max <= 8) s0 = 17;
    else if (l_max <= 10) s0 = 21;
    else s0 = 25;
    make_plan(n_species, n_max, l_max, s0, &a.plan);
    const size_t per_warp =
        8 * ((size_t)32 * 4 + (size_t)kHalfChunk * a.plan.row_stride +
             (size_t)n_species * a.plan.n_lm * n_max);
    int warps = 4;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    const size_t smem = per_warp * warps;
    if (smem > 227 * 1024)
        return set_error(DSG_ERR_UNSUPPORTED,
                         "n_species*n_max*(l_max+1)^2 needs %zu bytes of shared memory per warp",
                         per_warp);
    dim3 grid((n_cent + warps - 1) / warps, n_frames);
#define DSG_SOAP_LAUNCH(A, B, C, D)                                                              \
    do {                                                                                         \
        DSG_CUDA_CHECK(cudaFuncSetAttribute(soap_kernel<A, B, C, D>,                             \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                            (int)smem));                                         \
        soap_kernel<A, B, C, D><<<grid, warps * 32, smem, st>>>(a);                              \
    } while (0)
    if (l_max <= 4) DSG_SOAP_LAUNCH(9, 1, 0, 0);
    else if (l_max <= 6) DSG_SOAP_LAUNCH(13, 5, 0, 0);
    else if (l_max <= 8) DSG_SOAP_LAUNCH(17, 9, 1, 0);
    else if (l_max <= 10) DSG_SOAP_LAUNCH(21, 13, 5, 0);
    else DSG_SOAP_LAUNCH(25, 17, 9, 1);
#undef DSG_SOAP_LAUNCH
    DSG_LAUNCH_CHECK("soap_kernel");
    return DSG_OK;
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_soap_n_features(int n_species, int n_max, int l_max) {
    if (n_species < 1 || n_max < 1 || l_max < 0) return DSG_ERR_INVALID_ARGUMENT;
    const long long sn = (long long)n_species * n_max;
    const long long nf = sn * (sn + 1) / 2 * (l_max + 1);
    return nf > 0x7fffffffLL ? DSG_ERR_OVERFLOW : (int)nf;
}

extern "C" int dsg_soap_workspace_bytes(int n_frames, int n_atoms, int n_cent, int n_species,
                                        int n_max, int l_max, const double* h_box, double r_cut,
                                        size_t* bytes_out) {
    if (!bytes_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "bytes_out is NULL");
    SoapLayout L;
    int rc = soap_layout(n_frames, n_atoms, n_cent, n_species, n_max, l_max, h_box, r_cut, &L,
                         nullptr);
    if (rc) return rc;
    *bytes_out = L.bytes;
    return DSG_OK;
}

extern "C" int dsg_soap(const void* d_pos, int pos_dtype, const int32_t* d_species,
                        const int32_t* d_centres, int n_frames, int n_atoms, int n_cent,
                        int n_species, const double* h_box, double r_cut, int n_max, int l_max,
                        const double* h_alphas, const double* h_betas, void* d_workspace,
                        size_t workspace_bytes, double* d_out, int64_t stride_i, int64_t stride_f,
                        void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (pos_dtype == DSG_F32)
        return soap_impl<float>((const float*)d_pos, d_species, d_centres, n_frames, n_atoms,
                                n_cent, n_species, h_box, r_cut, n_max, l_max, h_alphas, h_betas,
                                d_workspace, workspace_bytes, d_out, stride_i, stride_f, st);
    if (pos_dtype == DSG_F64)
        return soap_impl<double>((const double*)d_pos, d_species, d_centres, n_frames, n_atoms,
                                 n_cent, n_species, h_box, r_cut, n_max, l_max, h_alphas, h_betas,
                                 d_workspace, workspace_bytes, d_out, stride_i, stride_f, st);
    return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                     pos_dtype);
}
