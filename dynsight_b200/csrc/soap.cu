// soap.cu -- SOAP power spectrum (GTO radial basis, sigma = 1) batched over frames (sm_100a).
//
// Replaces the DScribe call of soapify/saponify.py:98-122 and, fused with the timeSOAP distance,
// the chain trajectory.py:267-320 -> insight.py:136-155 -> timesoap.py:130-179.  One warp owns one
// centre:
//
//   gather   every periodic image within R = r_cut + sqrt(-2 ln 1e-3) is found through a cell list
//            (cell edge >= R; axes shorter than 3R fall back to an explicit image loop; a
//            non-periodic frame uses an open bounding-box grid).  List mode (periodic, >= 3 cells
//            per axis): soap_list_kernel searches with one thread per centre and the warp fetches
//            32 listed neighbours per round trip; otherwise hits are ballot-compacted into a
//            32-entry displacement list in shared memory (next batch prefetched)
//   phase B  two lanes per neighbour (even / odd orders m) evaluate the un-normalised regular solid
//            harmonics r^l Y_lm of 16 neighbours with fully unrolled polynomial recursions (no sqrt,
//            no division, exact at r = 0) into dense shared-memory rows
//   phase C  tensor path: per l, G_l^T (cols x 8 radial) += Y_l^T (cols x 4 neighbours) E_l (4 x 8)
//            as DMMA.8x8x4 (mma.sync.m8n8k4.f64); lane (k = lane >> 2, j = lane & 3) evaluates the
//            Gaussians exp(-gamma_lk r_j^2) -- each needed exponential exactly once, branch-free
//            float64 exp, 8 operations -- straight into its B fragment
//   epilogue Loewdin mixing D = Bm G and the power spectrum P = sum_m w^2 D D^T as DMMA products on
//            the C fragments of phase C (nothing leaves the registers); the lane writes its two
//            elements of every P_l and / or folds them into the timeSOAP sums against the spectrum
//            the same warp computed `delay` frames earlier (ring in L2).
//   Several species / n_max > 8 (split mode): the primitive sums G go to a scratch and
//   soap_spectrum_mma_kernel (tensor path, cp.async staging) or soap_spectrum_kernel finishes.
//
// The scalar-FMA form of phase C and the shared-memory epilogue of round 1 (lane (g, k) = radial
// index k x a group g of angular momenta, 20 FMAs + ten 128-bit shared loads per neighbour) stay
// selectable with -DDSG_SOAP_MMA=0 for A/B measurements.
//
// float64 throughout: the Loewdin matrices have entries of order 1e3..1e6 and cancel, so primitive
// sums need double precision to meet 1e-5 relative (SURVEY.md section 7 "hard parts").
#include <algorithm>
#include <cmath>
#include <cstring>
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
#ifndef DSG_SOAP_UNROLL
#define DSG_SOAP_UNROLL 1
#endif
#ifndef DSG_SOAP_CHUNK
#define DSG_SOAP_CHUNK 16
#endif
constexpr int kUnrollC = DSG_SOAP_UNROLL;  // phase-C neighbours unrolled together
constexpr int kChunk = DSG_SOAP_CHUNK;  // neighbours per phase-B/C pass (one lane each in phase B)
constexpr int kOpenCellsPerAxis = 32;
#ifndef DSG_SOAP_MIN_BLOCKS
#define DSG_SOAP_MIN_BLOCKS 6
#endif
#ifndef DSG_SOAP_MIN_BLOCKS_L10
#define DSG_SOAP_MIN_BLOCKS_L10 4
#endif
// Build-time switches of the measured variants (defaults = what ships)
#ifndef DSG_SOAP_MMA_BLOCKS_L8
#define DSG_SOAP_MMA_BLOCKS_L8 3   // resident blocks per SM of the l_max <= 8 tensor-path kernels (168 registers)
#endif
#ifndef DSG_SOAP_MMA_BLOCKS_L10
#define DSG_SOAP_MMA_BLOCKS_L10 2  // resident blocks per SM of the l_max = 9..12 tensor-path kernels
#endif
#ifndef DSG_SOAP_MMA
#define DSG_SOAP_MMA 1             // phase C on the fp64 tensor path (DMMA.8x8x4), see soap_kernel
#endif
#ifndef DSG_SOAP_PACK
#define DSG_SOAP_PACK 1            // l_max = 8: packed rows, 20 accumulator slots per lane
#endif
#ifndef DSG_SOAP_LIST
#define DSG_SOAP_LIST 1            // neighbour rows from soap_list_kernel where the frames allow it
#endif
#ifndef DSG_SOAP_LIST_WARPS
#define DSG_SOAP_LIST_WARPS 4      // warps per block of the list-mode kernel ...
#endif
#ifndef DSG_SOAP_LIST_BLOCKS
#if DSG_SOAP_MMA
#define DSG_SOAP_LIST_BLOCKS 3     // ... and its resident blocks per SM (168 registers)
#else
#define DSG_SOAP_LIST_BLOCKS 4     // ... and its resident blocks per SM (128 registers)
#endif
#endif
#ifndef DSG_SOAP_TILE_SPECTRUM
#define DSG_SOAP_TILE_SPECTRUM 1   // fused epilogue: 2x2 register tiles for n_max = 8
#endif
#ifndef DSG_SOAP_EXTRA_SMEM
#define DSG_SOAP_EXTRA_SMEM 0      // occupancy probe: dummy dynamic shared memory per block
#endif

constexpr int kListCap = 32;     // displacement list entries per warp
constexpr int kCellTab = 32;     // (cell range, shift) entries resolved per pass
constexpr int kBatchCap = 64;    // 32-candidate batches listed per pass
constexpr int kBatchDoubles = kBatchCap * 3 / 2;  // int2 + int per batch
__host__ __device__ constexpr int soap_warp_fixed(bool list) {
    return list ? kListCap * 4 + kCellTab * 3 : kListCap * 4 + kCellTab * 4 + kBatchDoubles;
}

enum AxisMode : int { kWide = 0, kNarrow = 1, kOpen = 2 };

struct SoapFrame {
    double box[3];
    double origin[3];
    double inv_edge[3];   // cells per unit length
    int nc[3];
    int mode[3];
    int nimg[3];
    int cell_base;        // first (species 0) cell of the frame in the batch-wide cell arrays
    int spec_stride;      // cells per species block: the sort key is species * spec_stride + cell
    unsigned lo_bits[3];  // open mode: encoded min / max (atomics)
    unsigned hi_bits[3];
};

struct SoapPlan {
    int n_max, l_max, n_species, n_feat;
    int n_lm;          // (l_max+1)^2
    int row_stride;    // P: doubles per neighbour in the harmonics buffer
    int rbuf_doubles;  // per-warp harmonics buffer (also holds the mixed coefficients)
    int kblocks;
    int item_l[kGroups][kMaxItems];  // -1 = unused
    int item_off[kGroups][kMaxItems];  // start of that l's block in a harmonics row
};

struct SoapArgs {
    const double* sx;       // cell-sorted SoA (wrapped / shifted coordinates)
    const double* sy;
    const double* sz;
    const int* sid;         // original atom index of a sorted slot
    const int* cell_start;
    const SoapFrame* frames;
    const int* centre_of_atom;  // (N) output row of an atom, -1 if it is not a centre
    int n_atoms, n_cent;
    int slots_per_warp;
    double r2max;
    const double* gamma;    // (L+1, n)
    const double* bmat;     // (L+1, n, n)  beta * prefactor
    const double* exp_tab;  // (kExpTab) bit patterns, see exp_scaled
    int bmat_smem_doubles;  // size of bmat if it is staged in shared memory, else 0
    int epi_smem_doubles;   // fused epilogue tables (swlm, lm_l, decode) in shared memory
    const double* swlm;     // (n_lm)  sqrt of the m-contraction weights
    const int* lm_l;        // (n_lm) -> l
    const int* decode;      // (n_feat)
    double* out;
    long long stride_i, stride_f;
    double* gscr;           // split mode: primitive sums G (n_frames, n_cent, S, n_lm, n)
    // neighbour rows built by soap_list_kernel (list mode): row of sorted slot s at rows + s * row_cap,
    // entry = slot of the neighbour inside its frame | (cell-table entry 0..26) << kListEntryShift
    int* rows;
    int* rowcnt;
    int row_cap;
    int* list_flag;         // != 0: some row overflowed row_cap -> the in-kernel gather runs instead
    int list_role;          // 0 no lists; 1 consumer of the lists; 2 fallback (runs iff the flag is set)
    // frame-sequential mode (SEQ kernels, dsg_soap_timesoap): a warp owns one centre and a run of
    // seq_run frame pairs; the spectrum of frame f is compared with the one of frame f - delay,
    // which the same warp left in its ring (global memory, rewritten every `delay` frames, so it
    // lives in L2), and only the timeSOAP value is written (the spectra too if out != NULL)
    const int* slot_of_atom;   // (F, N) sorted slot of an atom
    const int* centres;        // (n_cent) atom index of a centre (unused by the kernels so far)
    int n_frames, delay, seq_run;
    double* ring;              // (n_runs, n_cent, delay, n_feat + 1): vectors + squared norms
    double* ts_out;
    long long ts_ld, ts_col0;
    SoapPlan plan;
};
constexpr int kListEntryShift = 26;

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

// the cell of a coordinate that soap_locate has already wrapped (same arithmetic minus the wrap,
// which leaves such a value unchanged: no float64 division per centre)
__device__ __forceinline__ void soap_cell_of_wrapped(const SoapFrame* __restrict__ p, double x,
                                                     double y, double z, int* c) {
    const double in[3] = {x, y, z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v = p->mode[k] == kOpen ? in[k] - p->origin[k] : in[k];
        int ck = (int)(v * p->inv_edge[k]);
        const int nc = p->nc[k];
        c[k] = ck < 0 ? 0 : (ck >= nc ? nc - 1 : ck);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
soap_bin_kernel(const T* __restrict__ pos, const int* __restrict__ species, int n_atoms,
                const SoapFrame* __restrict__ fr, double* __restrict__ wpos,
                unsigned* __restrict__ cell_id, int* __restrict__ cell_cnt) {
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
    // species-major key: every species has its own complete cell-sorted array, so a gather for
    // one species reads only that species' atoms and x-adjacent cells stay contiguous
    const int cell = species[i] * p->spec_stride + (c[0] * p->nc[1] + c[1]) * p->nc[2] + c[2];
    cell_id[a] = (unsigned)cell;
    atomicAdd(&cell_cnt[p->cell_base + cell], 1);
}

__global__ void __launch_bounds__(256)
soap_scatter_kernel(const double* __restrict__ wpos, int n_atoms,
                    const SoapFrame* __restrict__ fr, const unsigned* __restrict__ cell_id,
                    int* __restrict__ cell_cnt, const int* __restrict__ cell_start,
                    double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz,
                    int* __restrict__ sid, int* __restrict__ slot_of) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const size_t a = (size_t)f * n_atoms + i;
    const int gc = fr[f].cell_base + (int)cell_id[a];
    const int slot = cell_start[gc] + atomicSub(&cell_cnt[gc], 1) - 1;
    sx[slot] = wpos[a * 3];
    sy[slot] = wpos[a * 3 + 1];
    sz[slot] = wpos[a * 3 + 2];
    sid[slot] = i;
    if (slot_of) slot_of[a] = slot;
}

__global__ void __launch_bounds__(256)
soap_centre_map_kernel(const int* __restrict__ centres, int n_cent, int* __restrict__ map) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_cent) map[centres[i]] = i;
}


// ------------------------------------------------------------------------------------------
// list mode: neighbour rows of every centre, one thread per sorted slot
// ------------------------------------------------------------------------------------------
// The distance search of the main kernel costs a warp 27 dependent rounds of candidate loads per
// centre (about a fifth of its time, all exposed latency at 3 resident warps per scheduler).  For
// periodic frames with >= 3 cells per axis this kernel does the search instead, one *thread* per
// centre at full occupancy: consecutive threads sit in the same cell, so their candidate loads are
// warp-wide broadcasts.  It keeps, per centre, the list of (neighbour slot, cell-table entry) codes
// in the order the in-kernel gather would have found them; the main kernel then fetches 32
// neighbours per round trip, one per lane, and recomputes the displacements itself.
__global__ void __launch_bounds__(128)
soap_list_kernel(const SoapArgs a) {
    const int f = blockIdx.y;
    const int slot_local = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot_local >= a.n_atoms) return;
    const size_t slot = (size_t)f * a.n_atoms + slot_local;
    if (a.centre_of_atom[a.sid[slot]] < 0) return;
    const SoapFrame* __restrict__ fr = a.frames + f;
    const double xc = a.sx[slot], yc = a.sy[slot], zc = a.sz[slot];
    int cc[3];
    soap_cell_of_wrapped(fr, xc, yc, zc, cc);
    const int nc0 = fr->nc[0], nc1 = fr->nc[1], nc2 = fr->nc[2];
    const double b0 = fr->box[0], b1 = fr->box[1], b2 = fr->box[2];
    const int frame_first = f * a.n_atoms;
    int* __restrict__ row = a.rows + slot * (size_t)a.row_cap;
    int* __restrict__ off = a.rowcnt + slot * (size_t)(a.plan.n_species + 1);
    const int cap = a.row_cap;
    const double r2max = a.r2max;
    int cnt = 0;
    off[0] = 0;
    // species after species (each has its own cell grid): the row is species-major and off[z + 1]
    // is the end of species z, so a (species, radial block) pass of the main kernel reads one run
    for (int z_sp = 0; z_sp < a.plan.n_species; ++z_sp) {
    const int cell_base = fr->cell_base + z_sp * fr->spec_stride;
    for (int i0 = 0; i0 < 3; ++i0) {
        int c0 = cc[0] + i0 - 1;
        double s0 = -xc;
        if (c0 < 0) { c0 += nc0; s0 -= b0; } else if (c0 >= nc0) { c0 -= nc0; s0 += b0; }
        for (int i1 = 0; i1 < 3; ++i1) {
            int c1 = cc[1] + i1 - 1;
            double s1 = -yc;
            if (c1 < 0) { c1 += nc1; s1 -= b1; } else if (c1 >= nc1) { c1 -= nc1; s1 += b1; }
            for (int i2 = 0; i2 < 3; ++i2) {
                int c2 = cc[2] + i2 - 1;
                double s2 = -zc;
                if (c2 < 0) { c2 += nc2; s2 -= b2; } else if (c2 >= nc2) { c2 -= nc2; s2 += b2; }
                const int cell = cell_base + (c0 * nc1 + c1) * nc2 + c2;
                const int qb = a.cell_start[cell], qe = a.cell_start[cell + 1];
                const int ecode = ((i0 * 3 + i1) * 3 + i2) << kListEntryShift;
                for (int q = qb; q < qe; ++q) {
                    const double dx = a.sx[q] + s0, dy = a.sy[q] + s1, dz = a.sz[q] + s2;
                    const double r2 = dx * dx + dy * dy + dz * dz;
                    if (r2 <= r2max) {
                        if (cnt < cap) row[cnt] = (q - frame_first) | ecode;
                        ++cnt;
                    }
                }
            }
        }
    }
    off[z_sp + 1] = cnt;
    }
    if (cnt > cap) *a.list_flag = 1;
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
template <int N>
struct Acc {
    double v[N > 0 ? N : 1];
};

// exp(x) for x in [-690, 0], argument given pre-scaled: t = x * 128 log2(e).  2^n * 2^(j/128) * e^u
// with 128 n + j = rint(t); 2^(j/128) from a shared-memory table, e^u by a degree-4 Taylor
// polynomial in v = t - rint(t) (coefficients (ln2/128)^i / i! folded in; |u| <= ln2/256,
// truncation 1.2e-15).  Relative error < 4e-15.  Branch-free, 8 fp64 operations (3 for the range
// reduction, 4 Horner steps with constant-bank coefficients, 1 product): the kernel is bound by the
// fp64 pipe, which the tensor-path products share, so the exponential is
// built for the fewest fp64 instructions.  The table is replicated 16 times, entry j of copy r at
// tab[16 j + r]: lane `lane` reads copy lane & 15, i.e. its own bank pair, so a lookup costs the
// two wavefronts of any 64-bit warp load whatever the 32 indices are (the 32-entry table of round
// 1 cost 2 to 4).  The table holds bit patterns, not values: the high word of entry j is that of
// 2^(j/128) minus j << 13, so adding (128 n + j) << 13 = (n << 20) + (j << 13) puts n into the
// exponent without masking j out first.
// The caller guarantees t >= -690 * kExpScale (2^n stays normal): soap_layout rejects cutoffs
// for which gamma * r^2 could exceed 690.
#ifndef DSG_SOAP_EXP_DEG3
#define DSG_SOAP_EXP_DEG3 0   // measured variant: 256-entry table in 8 copies, economised cubic (1.8e-14)
#endif
#if DSG_SOAP_EXP_DEG3
constexpr int kExpTab = 256;
constexpr int kExpRep = 8;
constexpr int kExpHiShift = 12;
constexpr double kExpScale = 369.3299304675746;      // 256 * log2(e)
__constant__ double kExpPoly[4] = {   // Chebyshev-economised cubic of e^u on |u| <= ln2/512, in v
    2.7076061740622624e-03, 3.6655661567589036e-06, 3.308303059503886e-09, 0.9999999999999825,
};
#else
constexpr int kExpTab = 128;
#ifndef DSG_SOAP_EXP_REP
#define DSG_SOAP_EXP_REP 16
#endif
constexpr int kExpRep = DSG_SOAP_EXP_REP;
constexpr int kExpHiShift = 13;                      // 20 - log2(kExpTab)
constexpr double kExpScale = 184.6649652337873;      // 128 * log2(e)
__constant__ double kExpPoly[4] = {
    5.4152123481245725e-03,   // (ln2/128)^1 / 1!
    1.4662262387640423e-05,   // (ln2/128)^2 / 2!
    2.6466421444330967e-08,   // (ln2/128)^3 / 3!
    3.583032305400251e-11,    // (ln2/128)^4 / 4!
};
#endif
constexpr int kExpTabDoubles = kExpTab * kExpRep;   // 16 KB per block
// g = gamma * kExpScale (negative), s = r^2, tab = the lane's copy of the table.  The product
// enters only through two FMAs: rint(g s) by the magic-constant trick and the remainder
// v = g s - rint(g s) with a single rounding, so the argument is never rounded to float64 (one
// operation fewer and v exact to 1e-17).
__device__ __forceinline__ double exp_scale_factor(int kq, const double* __restrict__ tab) {
    const double tv = tab[(kq & (kExpTab - 1)) * kExpRep];
    return __hiloint2double(__double2hiint(tv) + (kq << kExpHiShift), __double2loint(tv));
}
__device__ __forceinline__ double exp_scaled(double g, double s, const double* __restrict__ tab) {
    const double magic = 6755399441055744.0;                     // 1.5 * 2^52
    const double tn = fma(g, s, magic);
    const double sc = exp_scale_factor(__double2loint(tn), tab);  // 2^n 2^(j/128), 128 n + j = rint(g s)
    const double v = fma(g, s, magic - tn);                      // [-1/2, 1/2]
#if DSG_SOAP_EXP_DEG3
    double p = fma(kExpPoly[2], v, kExpPoly[1]);
    p = fma(p, v, kExpPoly[0]);
    p = fma(p, v, kExpPoly[3]);
#else
    double p = fma(kExpPoly[3], v, kExpPoly[2]);
    p = fma(p, v, kExpPoly[1]);
    p = fma(p, v, kExpPoly[0]);
    p = fma(p, v, 1.0);
#endif
    return sc * p;
}

// NE exponentials of the same s, written stage by stage: ptxas keeps roughly the source order, so
// the NE dependency chains advance together (each stage is NE independent fp64 operations) instead
// of two at a time -- the tensor-path loop is bound by the latency of these chains, not by a pipe.
template <int NE>
__device__ __forceinline__ void exp_scaled_batch(const double (&g)[NE], double s,
                                                 const double* __restrict__ tab, double (&e)[NE]) {
    const double magic = 6755399441055744.0;
    double tn[NE], v[NE], p[NE], sc[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) tn[i] = fma(g[i], s, magic);
#pragma unroll
    for (int i = 0; i < NE; ++i) v[i] = fma(g[i], s, magic - tn[i]);
#pragma unroll
    for (int i = 0; i < NE; ++i) sc[i] = exp_scale_factor(__double2loint(tn[i]), tab);
#if DSG_SOAP_EXP_DEG3
#pragma unroll
    for (int i = 0; i < NE; ++i) p[i] = fma(kExpPoly[2], v[i], kExpPoly[1]);
#pragma unroll
    for (int i = 0; i < NE; ++i) p[i] = fma(p[i], v[i], kExpPoly[0]);
#pragma unroll
    for (int i = 0; i < NE; ++i) p[i] = fma(p[i], v[i], kExpPoly[3]);
#else
#pragma unroll
    for (int i = 0; i < NE; ++i) p[i] = fma(kExpPoly[3], v[i], kExpPoly[2]);
#pragma unroll
    for (int i = 0; i < NE; ++i) p[i] = fma(p[i], v[i], kExpPoly[1]);
#pragma unroll
    for (int i = 0; i < NE; ++i) p[i] = fma(p[i], v[i], kExpPoly[0]);
#pragma unroll
    for (int i = 0; i < NE; ++i) p[i] = fma(p[i], v[i], 1.0);
#endif
#pragma unroll
    for (int i = 0; i < NE; ++i) e[i] = sc[i] * p[i];
}

// Layout of one neighbour's harmonics row: the block of angular momentum l (2l+1 values, padded
// to even length) starts at harm_off(LT, l); element (l, m) sits at harm_off + l + m.  l = 0 is
// not used (its single value is 1 and that channel is accumulated separately).  For LT = 8 the
// blocks of l = 1 and l = 7 change places and l = 0 is dropped: with the natural order the four
// lane groups of phase C (which read the blocks of four different l with the same running offset)
// hit the same shared-memory banks for one pair of l; with this order the four block starts are
// distinct modulo 128 bytes for every l_max <= 8 (checked exhaustively against make_plan's
// packing; +6 % on the l_max = 8 kernel).  The other instantiations keep the natural order
// l(l+1): the same change made the l_max = 10 kernel slower (ncu: half the bank conflicts but 6 %
// more instructions and a worse schedule at its 255-register limit).
__host__ __device__ constexpr int harm_off(int lt, int l) {
    if (lt != 8) return l * (l + 1);
    switch (l) {  // block order 7, 2, 3, 4, 5, 6, 1, 8, 9, 10, 11, 12 (sizes 2l+2)
        case 1: return 66;
        case 2: return 16;
        case 3: return 22;
        case 4: return 30;
        case 5: return 40;
        case 6: return 52;
        case 7: return 0;
        case 8: return 70;
        case 9: return 88;
        case 10: return 108;
        case 11: return 130;
        case 12: return 154;
        default: return 0;
    }
}
__host__ __device__ constexpr int harm_len(int lt) {
    return lt != 8 ? (lt + 1) * (lt + 2) : harm_off(lt, lt) + 2 * lt + 2;
}

// Packed rows for l_max = 8 exactly (soap_kernel<8, 20, 0, 0, 0>): the longest-processing-time
// packing gives lane group g the angular momenta {8 - g, 1 + g} with (17 - 2g) + (3 + 2g) = 20
// values, so each group's values are laid out as one contiguous run of 20 doubles at g * 20 (no
// padding; the four runs start 160 bytes apart = 8 banks, so the 128-bit broadcast reads of the four
// groups never share a bank) and every lane accumulates exactly 20 slots instead of 17 + 9.
constexpr int kPackSlots = 20;
__host__ __device__ constexpr int pack_size0(int g) { return 17 - 2 * g; }
// Row layouts: natural / bank-ordered blocks (kRowBlocks), packed l_max = 8 rows (kRowPacked), and
// the dense rows of the tensor path (kRowDense: block l at l^2 - 1, no padding, l = 0 not stored)
enum RowLayout : int { kRowBlocks = 0, kRowPacked = 1, kRowDense = 2 };
template <int LT, int LAY>
__host__ __device__ constexpr int hoff(int l) {
    if (LAY == kRowDense) return l * l - 1;
    if (LAY == kRowBlocks) return harm_off(LT, l);
    return l >= 5 ? (8 - l) * kPackSlots : (l - 1) * kPackSlots + pack_size0(l - 1);
}
// tensor path: 8-column tiles of the (2l+1) values of one l
// l = 4, 8, 12 have 8 t + 1 columns: a tile for one column would be 7/8 padding, so their last
// column (m = +l) is "peeled": accumulated like l = 0 by a scalar FMA per lane (k, j) and handed to
// the epilogue as a one-column tile (two tiles out of 14 less per four neighbours at l_max = 8)
__host__ __device__ constexpr bool mma_peeled(int l) { return l > 0 && l % 4 == 0; }
__host__ __device__ constexpr int mma_cols(int l) { return mma_peeled(l) ? 2 * l : 2 * l + 1; }
__host__ __device__ constexpr int mma_tiles(int l) { return (mma_cols(l) + 7) / 8; }
__host__ __device__ constexpr int mma_tile_base(int l) {
    int b = 0;
    for (int i = 1; i < l; ++i) b += mma_tiles(i);
    return b;
}
__host__ __device__ constexpr int mma_row_stride(int lt) { return (lt * (lt + 2)) | 1; }

// Phase B: un-normalised regular solid harmonics of one displacement:
//   rt[off(l) + l + m] = S_l^m(z, r2) * A_m(x, y),   rt[off(l) + l - m] = S_l^m * B_m   (m > 0)
//   A_m + i B_m = (x + i y)^m ;  S_m^m = (2m-1)!!, S_{m+1}^m = (2m+1) z S_m^m,
//   S_l^m = ((2l-1) z S_{l-1}^m - (l+m-1) r2 S_{l-2}^m) / (l-m)
template <int LT, int LAY>
__device__ __forceinline__ void solid_harmonics(double x, double y, double z, double r2,
                                                double* __restrict__ rt) {
    double am = 1.0, bm = 0.0, dfact = 1.0;
#pragma unroll
    for (int m = 0; m <= LT; ++m) {
        if (m > 0) {
            const double an = am * x - bm * y;
            bm = fma(am, y, bm * x);
            am = an;
            dfact *= (double)(2 * m - 1);
        }
        double s0 = dfact;  // S_m^m
        if (m > 0) {
            rt[hoff<LT, LAY>(m) + 2 * m] = s0 * am;
            rt[hoff<LT, LAY>(m)] = s0 * bm;
        }
        if (m + 1 <= LT) {
            double s1 = (double)(2 * m + 1) * z * s0;  // S_{m+1}^m
            {
                const int o = hoff<LT, LAY>(m + 1) + m + 1;
                rt[o + m] = s1 * am;
                if (m > 0) rt[o - m] = s1 * bm;
            }
#pragma unroll
            for (int l = m + 2; l <= LT; ++l) {
                const double inv = 1.0 / (double)(l - m);  // compile-time constant after unrolling
                const double s2 =
                    ((double)(2 * l - 1) * z * s1 - (double)(l + m - 1) * r2 * s0) * inv;
                const int o = hoff<LT, LAY>(l) + l;
                rt[o + m] = s2 * am;
                if (m > 0) rt[o - m] = s2 * bm;
                s0 = s1;
                s1 = s2;
            }
        }
    }
}

// Phase B of the tensor path on all 32 lanes: two lanes share one neighbour, lane half h = 0 / 1
// takes the even / odd orders m = 2 i + h.  Written in d = l - m, the recursion coefficients of
// the two halves differ only by terms linear in h -- (2l-1) = (4i+2d-1) + 2h, (l+m-1) = (4i+d-1) + 2h,
// 1/(l-m) = 1/d -- which fold into FMAs with the lane constants 2 h z and 2 h r2: the same
// instruction stream for both halves, no selects.  (x + i y)^m advances by (x + i y)^2.  A lane
// runs 16 recursion steps and 50 products instead of 36 and 80.  Dense rows (kRowDense); the odd
// half has nothing to store where l = 2 i + d + 1 would exceed LT.
template <int LT>
__device__ __forceinline__ void solid_harmonics_halves(double x, double y, double z, double r2,
                                                       int h, double* __restrict__ rt) {
    const double hd = (double)h;
    const int h2 = 2 * h;
    const double z2h = 2.0 * hd * z, r22h = 2.0 * hd * r2;
    double zl[LT + 2];
#pragma unroll
    for (int q = 1; q <= LT; ++q) zl[q] = fma((double)(2 * q - 1), z, z2h);      // (2l-1) z, l = q + h
    const double w2r = x * x - y * y, w2i = 2.0 * x * y;
    double am = h ? x : 1.0, bm = h ? y : 0.0, smm = 1.0;   // (x + i y)^h, (2h-1)!! = 1
#pragma unroll
    for (int i = 0; 2 * i <= LT; ++i) {
        if (i > 0) {
            const double an = am * w2r - bm * w2i;
            bm = fma(am, w2i, bm * w2r);
            am = an;
            // (2m-1)!! = (2m-5)!! (2m-3)(2m-1) with m = 2i + h
            const double ce = (double)((4 * i - 3) * (4 * i - 1));   // constants after unrolling
            const double co = (double)((4 * i - 1) * (4 * i + 1));
            smm *= fma(co - ce, hd, ce);
        }
        double s0 = smm, s1 = 0.0;
        // d = 0: l = m = 2i + h -> (2i)^2 + 2i - 1 +- 2i, plus h (4i + 3) resp. h (4i + 1)
        double* pp = rt + (4 * i * i + 4 * i - 1) + h * (4 * i + 3);
        double* pm = rt + (4 * i * i - 1) + h * (4 * i + 1);
#pragma unroll
        for (int d = 0; 2 * i + d <= LT; ++d) {
            const int q = 2 * i + d;   // l = q + h
            double sv;
            if (d == 0) {
                sv = s0;
            } else if (d == 1) {
                s1 = zl[q] * s0;       // S_{m+1}^m = (2m+1) z S_m^m, 2m+1 = 2(q+h)-1
                sv = s1;
            } else {
                const double rm = fma((double)(4 * i + d - 1), r2, r22h);   // (l+m-1) r2
                const double s2 = (zl[q] * s1 - rm * s0) * (1.0 / (double)d);
                s0 = s1;
                s1 = s2;
                sv = s2;
            }
            // element (l, +m) at l^2 - 1 + l + m, (l, -m) at l^2 - 1 + l - m; from l to l + 1 both
            // move by 2l + 2 = 2q + 2 + 2h: walking pointers, two registers instead of 50 offsets
            const bool ok = (q < LT || h == 0) && (q > 0 || h);   // l <= LT, and l = 0 is not stored
            if (ok) {
                *pp = sv * am;
                if (i > 0 || h) *pm = sv * bm;
            }
            pp += 2 * q + 2 + h2;
            pm += 2 * q + 2 + h2;
        }
    }
}

// acc[0..S) += e * r[0..S), r 16-byte aligned, read as 128-bit shared loads
template <int S>
__device__ __forceinline__ void accumulate(Acc<S>& a, double e, const double* __restrict__ r) {
    if (S <= 0) return;
    const double2* __restrict__ r2p = reinterpret_cast<const double2*>(r);
#pragma unroll
    for (int s = 0; s + 1 < S; s += 2) {
        const double2 v = r2p[s >> 1];
        a.v[s] = fma(e, v.x, a.v[s]);
        a.v[s + 1] = fma(e, v.y, a.v[s + 1]);
    }
    if (S & 1) a.v[S - 1] = fma(e, r[S - 1], a.v[S - 1]);
}

// packed l_max = 8 rows: slot s of lane group g belongs to l = 8 - g while s < 17 - 2g, to l = 1 + g
// after that; e0 / e1 are the two Gaussians of the lane.  Slots 11..16 change owner with g: three
// selects per neighbour instead of six padded FMAs and three padded 128-bit loads.
__device__ __forceinline__ void accumulate_packed(Acc<kPackSlots>& a, double e0, double e1, int g,
                                                  const double* __restrict__ r) {
    const double2* __restrict__ r2p = reinterpret_cast<const double2*>(r);
    const double ea = g < 3 ? e0 : e1, eb = g < 2 ? e0 : e1, ec = g < 1 ? e0 : e1;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const double2 v = r2p[s];
        a.v[2 * s] = fma(e0, v.x, a.v[2 * s]);
        a.v[2 * s + 1] = fma(e0, v.y, a.v[2 * s + 1]);
    }
    double2 v = r2p[5];
    a.v[10] = fma(e0, v.x, a.v[10]);
    a.v[11] = fma(ea, v.y, a.v[11]);
    v = r2p[6];
    a.v[12] = fma(ea, v.x, a.v[12]);
    a.v[13] = fma(eb, v.y, a.v[13]);
    v = r2p[7];
    a.v[14] = fma(eb, v.x, a.v[14]);
    a.v[15] = fma(ec, v.y, a.v[15]);
    v = r2p[8];
    a.v[16] = fma(ec, v.x, a.v[16]);
    a.v[17] = fma(e1, v.y, a.v[17]);
    v = r2p[9];
    a.v[18] = fma(e1, v.x, a.v[18]);
    a.v[19] = fma(e1, v.y, a.v[19]);
}

// C (8x8) += A (8x4) * B (4x8) on the fp64 tensor path.  Fragments (PTX m8n8k4): A[row = lane >> 2]
// [col = lane & 3], B[row = lane & 3][col = lane >> 2], C[row = lane >> 2][cols 2 (lane & 3) + {0, 1}].
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ void store_packed(const Acc<kPackSlots>& a, int g, int k, int n,
                                             double* __restrict__ cz) {
    const int size0 = pack_size0(g), la = 8 - g, lb = 1 + g;
#pragma unroll
    for (int s = 0; s < kPackSlots; ++s) {
        const int lm = s < size0 ? la * la + s : lb * lb + s - size0;
        cz[lm * n + k] = a.v[s];
    }
}

template <int S>
__device__ __forceinline__ void store_item(const Acc<S>& a, int l, int k, int n,
                                           double* __restrict__ cz) {
    if constexpr (S > 0) {
        if (l < 0) return;
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (s < 2 * l + 1) cz[(l * l + s) * n + k] = a.v[s];
    }
}

// Loewdin mixing + power spectrum of one centre.  G [S][n_lm][n] may sit in shared or global
// memory; the mixed coefficients c[z][lm][n'] = sqrt(w_lm) * sum_k bmat[l][n'][k] G[z][lm][k] go to
// the shared buffer `mix`, then one lane per output contracts over m and the warp writes the
// spectrum coalesced.
// Where the features of one (centre, frame) go: the spectrum row (out, may be NULL) and/or the
// timeSOAP sums against the vector the same warp computed `delay` frames earlier (ring, may be
// NULL; it is overwritten with the new vector).  Lane `lane` handles the features c = lane (mod 32)
// in ascending order with the same fma sequence as timesoap_stream_kernel, so the fused value equals
// the two-kernel one.
struct SpectrumSink {
    double* out;
    double* ring;
    double csq, dot;
    // staged spectrum (shared memory): the ring loads of a lane are independent of its ring
    // stores only if they all come first -- one L2 round trip per frame instead of eleven
    __device__ __forceinline__ void put_staged(const double* __restrict__ stage, int n_feat,
                                               int lane) {
        if (ring) {
            for (int c = lane; c < n_feat; c += 32) {
                const double v = stage[c];
                csq = fma(v, v, csq);
                dot = fma(v, ring[c], dot);
            }
            for (int c = lane; c < n_feat; c += 32) ring[c] = stage[c];
        }
        if (out)
            for (int c = lane; c < n_feat; c += 32) out[c] = stage[c];
    }
    __device__ __forceinline__ void put(int c, double v) {
        if (out) out[c] = v;
        if (ring) {
            csq = fma(v, v, csq);
            dot = fma(v, ring[c], dot);
            ring[c] = v;
        }
    }
};

__device__ __forceinline__ void mix_and_spectrum(const SoapArgs& a, const double* __restrict__ gsum,
                                                 double* __restrict__ mix,
                                                 const double* __restrict__ bmat,
                                                 const double* __restrict__ swlm,
                                                 const unsigned char* __restrict__ lm_l,
                                                 const unsigned short* __restrict__ decode,
                                                 SpectrumSink& sink, int lane) {
    const SoapPlan& pl = a.plan;
    const int n = pl.n_max, n_lm = pl.n_lm;
    const int n_out = pl.n_species * n_lm * n;
    if (n == 8) {
        // one column (z, lm) per lane: G[0..8) in registers, 8x8 matrix rows as 128-bit broadcasts
        for (int col = lane; col < pl.n_species * n_lm; col += 32) {
            const int lm = col % n_lm;
            const int l = lm_l[lm];
            const double w = swlm[lm];
            const double2* __restrict__ v2 = reinterpret_cast<const double2*>(gsum + (size_t)col * 8);
            const double2 g0 = v2[0], g1 = v2[1], g2 = v2[2], g3 = v2[3];
            const double2* __restrict__ b2 = reinterpret_cast<const double2*>(bmat + (size_t)l * 64);
            double2* __restrict__ o2 = reinterpret_cast<double2*>(mix + (size_t)col * 8);
#pragma unroll
            for (int np = 0; np < 8; np += 2) {
                double r[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double2 b0 = b2[(np + h) * 4], b1 = b2[(np + h) * 4 + 1];
                    const double2 b3 = b2[(np + h) * 4 + 2], b4 = b2[(np + h) * 4 + 3];
                    double sa = b0.x * g0.x, sb = b0.y * g0.y;
                    sa = fma(b1.x, g1.x, sa); sb = fma(b1.y, g1.y, sb);
                    sa = fma(b3.x, g2.x, sa); sb = fma(b3.y, g2.y, sb);
                    sa = fma(b4.x, g3.x, sa); sb = fma(b4.y, g3.y, sb);
                    r[h] = (sa + sb) * w;
                }
                o2[np >> 1] = make_double2(r[0], r[1]);
            }
        }
    } else {
        int col = lane / n, np = lane - col * n;   // one division per lane, then increments
        int lm = col % n_lm;
        const int dcol = 32 / n, dnp = 32 - dcol * n;
        for (int idx = lane; idx < n_out; idx += 32) {
            const int l = lm_l[lm];
            const double* __restrict__ v = gsum + (size_t)col * n;
            const double* __restrict__ bm = bmat + ((size_t)l * n + np) * n;
            double s0 = 0.0, s1 = 0.0;
            int k = 0;
            for (; k + 1 < n; k += 2) {
                s0 = fma(bm[k], v[k], s0);
                s1 = fma(bm[k + 1], v[k + 1], s1);
            }
            if (k < n) s0 = fma(bm[k], v[k], s0);
            mix[idx] = (s0 + s1) * swlm[lm];
            col += dcol;
            lm += dcol;
            np += dnp;
            if (np >= n) { np -= n; ++col; ++lm; }
            while (lm >= n_lm) lm -= n_lm;
        }
    }
    __syncwarp();

    if (DSG_SOAP_TILE_SPECTRUM && n == 8) {
        // 2x2 register tiles: one lane per (l, tile of the upper triangle of the 8x8 Gram matrix
        // C_l^T C_l), two 128-bit shared loads per four FMAs, l descending so that the lanes of a
        // round have similar trip counts.  The results are staged over G (dead after the mixing) and
        // written out coalesced.
        double* __restrict__ stage = const_cast<double*>(gsum);
        const int n_tasks = (pl.l_max + 1) * 10;
        for (int t = lane; t < n_tasks; t += 32) {
            const int lq = t / 10, tile = t - lq * 10, l = pl.l_max - lq;
            const int na0 = 2 * (int)((0x3221110000ULL >> (4 * tile)) & 15ULL);
            const int nb0 = 2 * (int)((0x3323213210ULL >> (4 * tile)) & 15ULL);
            const double2* __restrict__ ca =
                reinterpret_cast<const double2*>(mix + (size_t)(l * l) * 8 + na0);
            const double2* __restrict__ cb =
                reinterpret_cast<const double2*>(mix + (size_t)(l * l) * 8 + nb0);
            double o00 = 0.0, o01 = 0.0, o10 = 0.0, o11 = 0.0;
            for (int mi = 0; mi < 2 * l + 1; ++mi) {
                const double2 u = ca[mi * 4], v = cb[mi * 4];
                o00 = fma(u.x, v.x, o00);
                o01 = fma(u.x, v.y, o01);
                o10 = fma(u.y, v.x, o10);
                o11 = fma(u.y, v.y, o11);
            }
            double* __restrict__ ob = stage + l * 36;
            const int r0 = na0 * 8 - (na0 * (na0 - 1)) / 2 - na0;        // row na0: index = r0 + nb
            const int r1 = (na0 + 1) * 8 - ((na0 + 1) * na0) / 2 - (na0 + 1);
            ob[r0 + nb0 + 1] = o01;
            ob[r1 + nb0 + 1] = o11;
            if (nb0 >= na0) ob[r0 + nb0] = o00;       // always (tiles of the upper triangle)
            if (nb0 > na0) ob[r1 + nb0] = o10;        // below the diagonal inside a diagonal tile
        }
        __syncwarp();
        sink.put_staged(stage, pl.n_feat, lane);
        __syncwarp();
        return;
    }
    for (int c = lane; c < pl.n_feat; c += 32) {
        const int code = decode[c];   // single species: l | na << 4 | nb << 9
        const int l = code & 15, na = (code >> 4) & 31, nb = (code >> 9) & 31;
        const double* __restrict__ ca = mix + (size_t)(l * l) * n + na;
        const double* __restrict__ cb = mix + (size_t)(l * l) * n + nb;
        // 2l+1 terms: one leading term, then pairs on two independent chains
        double s0 = ca[0] * cb[0], s1 = 0.0;
        for (int mi = 1; mi < 2 * l + 1; mi += 2) {
            s0 = fma(ca[mi * n], cb[mi * n], s0);
            s1 = fma(ca[(mi + 1) * n], cb[(mi + 1) * n], s1);
        }
        sink.put(c, s0 + s1);
    }
    __syncwarp();
}

// SPLIT = false: one species and n_max <= 8 -- the primitive sums of a centre are complete after
// one pass, so mixing and the power spectrum follow in the same kernel (G never leaves the SM).
// SPLIT = true: several species and/or radial blocks -- the warp walks its run of centres once per
// (species, radial block) pass and writes that slice of G to a global scratch;
// soap_spectrum_kernel finishes the centres.  No per-warp coefficient buffer is needed, so the
// occupancy is the single-species one whatever the number of species.
// List mode needs neither the prefetch registers nor the batch tables of the in-kernel search, so
// it also fits 4 blocks of 4 warps per SM at 128 registers.  With the 8-operation exponential
// (round 2) that is the faster shape: 4.68e7 centre-frames/s for 4 x 4 against 4.52e7 for 6 x 2 at
// 168 registers on the 100 000-atom box (round 1, 13-operation exponential: 4.30e7 vs 4.38e7).
constexpr int kListWarps = DSG_SOAP_LIST_WARPS;
// SEQ = frame-sequential: warp w of the grid owns centre w and the frame pairs
// [blockIdx.y * seq_run, ...) -- it walks the frames of its run in order (plus the `delay` frames
// before its first output), finds its centre in each frame's sorted order through slot_of_atom,
// and hands every spectrum to the timeSOAP sums against its ring.  Not combined with SPLIT (there
// soap_spectrum_kernel is the frame-sequential one).
template <int LT, int S0, int S1, int S2, int S3, bool SPLIT, bool LIST, bool SEQ = false>
__global__ void __launch_bounds__(
    DSG_SOAP_MMA ? 32 * kListWarps : ((LIST && S0 <= 20) ? 32 * kListWarps : 64),
    // tensor path: four warps share the 16 KB exponential table; 3 blocks per SM (168 registers)
    // up to l_max = 8, 2 blocks (255 registers: 40 / 56 accumulator registers more) beyond
    DSG_SOAP_MMA ? (LT <= 8 ? DSG_SOAP_MMA_BLOCKS_L8 : DSG_SOAP_MMA_BLOCKS_L10)
                 : ((S0 <= 20) ? (LIST ? DSG_SOAP_LIST_BLOCKS : DSG_SOAP_MIN_BLOCKS)
                               : (S0 <= 21 ? DSG_SOAP_MIN_BLOCKS_L10 : 4)))
soap_kernel(const SoapArgs a) {
    static_assert(!(SEQ && SPLIT), "frame-sequential accumulation is only the fused kernel");
    // MM: phase C as tensor-core products G_l (8 radial x 2l+1) += E_l (8 x 4 neighbours) * Y_l
    // (4 x 2l+1): lane (k = lane >> 2, j = lane & 3) evaluates the Gaussians of radial index k on
    // neighbour j of each group of four -- every needed exponential exactly once, as before --
    // and they go straight into the A fragment; the harmonics are the B fragment (one 64-bit
    // shared load per 8x8x4 tile and lane instead of one 128-bit load per two FMAs: the scalar
    // loop moved 20 doubles per lane and neighbour through the shared-memory pipe, which ran at
    // 82 % of its peak).  2l+1 is padded to 8-column tiles (the pad columns read the next block
    // of the row and are discarded), l = 0 stays a scalar sum.
    constexpr bool MM = DSG_SOAP_MMA != 0;
    [[maybe_unused]] constexpr bool PK = !MM && LT == 8 && S0 == kPackSlots;  // packed rows (see hoff)
    constexpr int kTiles = mma_tile_base(LT + 1);
    constexpr int kPeel = LT / 4;   // peeled columns (l = 4, 8, 12)
    // list mode: the consumer runs unless a row overflowed, the fallback only if one did
    if (a.list_role != 0 && ((*a.list_flag != 0) != (a.list_role == 2))) return;
    extern __shared__ __align__(16) double smem_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    // block-wide tables: 2^(j/128) (16 copies, the lane reads its own) and the Loewdin matrices
    double* exp_tab_all = smem_all;
    const double* exp_tab = exp_tab_all + (lane & (kExpRep - 1));
    double* bmat_s = smem_all + kExpTabDoubles;
    const int bmat_stage = SPLIT ? 0 : a.bmat_smem_doubles;
    // fused epilogue: m-weights, lm -> l and the feature decode table also sit in shared memory
    // (their global loads were exposed latency in every centre's epilogue)
    const int epi_stage = SPLIT ? 0 : a.epi_smem_doubles;
    double* swlm_s = bmat_s + bmat_stage;
    unsigned short* decode_s = reinterpret_cast<unsigned short*>(swlm_s + a.plan.n_lm);
    unsigned char* lm_l_s = reinterpret_cast<unsigned char*>(decode_s + a.plan.n_feat);
    for (int i = threadIdx.x; i < kExpTabDoubles; i += blockDim.x)
        exp_tab_all[i] = a.exp_tab[i / kExpRep];
    if (MM && !SPLIT) {
        // tensor-path epilogue: Loewdin matrices padded to 8 x 8 (zero rows / columns behind
        // n_max), and per 8-column tile the squared m-weights (zero in the pad columns)
        const int nm = a.plan.n_max;
        for (int i = threadIdx.x; i < bmat_stage; i += blockDim.x) {
            const int l = i >> 6, np = (i >> 3) & 7, kq = i & 7;
            bmat_s[i] = (l <= a.plan.l_max && np < nm && kq < nm)
                            ? a.bmat[((size_t)l * nm + np) * nm + kq] : 0.0;
        }
        for (int i = threadIdx.x; i < 8 * (kTiles + 1 + kPeel); i += blockDim.x) {
            // tile 0 = l = 0, then the tiles of l = 1, 2, ..., then one single-column tile per
            // peeled l = 4, 8, ...
            const int tt = i >> 3, c = i & 7;
            int l = 0, col = c;
            bool valid = c == 0;
            if (tt > kTiles) {
                l = 4 * (tt - kTiles);
                col = 2 * l;
            } else if (tt > 0) {
                l = 1;
                while (l < LT && mma_tile_base(l + 1) <= tt - 1) ++l;
                col = 8 * (tt - 1 - mma_tile_base(l)) + c;
                valid = col < mma_cols(l);
            }
            const double w = (l <= a.plan.l_max && valid) ? a.swlm[l * l + col] : 0.0;
            swlm_s[i] = w * w;
        }
    } else {
    for (int i = threadIdx.x; i < bmat_stage; i += blockDim.x) bmat_s[i] = a.bmat[i];
    if (!SPLIT) {
        for (int i = threadIdx.x; i < a.plan.n_lm; i += blockDim.x) {
            swlm_s[i] = a.swlm[i];
            lm_l_s[i] = (unsigned char)a.lm_l[i];
        }
        for (int i = threadIdx.x; i < a.plan.n_feat; i += blockDim.x) {
            const int code = a.decode[i];   // zi = zj = 0 in the fused (single-species) kernel
            decode_s[i] = (unsigned short)(((code >> 8) & 15) | (((code >> 12) & 31) << 4) |
                                           (((code >> 17) & 31) << 9));
        }
    }
    }
    __syncthreads();
    const double* __restrict__ bmat = bmat_stage > 0 ? bmat_s : a.bmat;
    double* smem = smem_all + kExpTabDoubles + bmat_stage + epi_stage;
    // non-SEQ: each warp walks a run of consecutive slots of the frame's (species, cell)-sorted
    // order: consecutive centres share their candidate cells (L1 reuse, cached cell table) and
    // the block-wide tables are loaded once
    int first = 0, last = 0, seq_t0 = 0, seq_frames = 1, seq_first = 0;
    if (SEQ) {
        // warp w takes the atom that sits at sorted slot w of the run's FIRST frame (neighbours in
        // space, so the warps of a block keep sharing lines).
        // Measured on 100 000 atoms x 33 frames: groups of 1 / 2 / 4 / 8 / 16 / 32 centres per warp
        // give 69.3 / 72.1 / 73.6 / 76.1 / 77.6 / 80.2 ms -- more, shorter-lived warps balance the
        // SMs better than longer walks reuse L1, so a warp owns ONE centre (69.0 ms + 1.3 ms for
        // the two-kernel path).
        seq_first = blockIdx.x * warps + warp;
        if (seq_first >= a.n_atoms) return;   // warps are independent from here: only __syncwarp below
        seq_t0 = blockIdx.y * a.seq_run;                       // first frame pair of the run
        const int t1 = min(a.n_frames - a.delay, seq_t0 + a.seq_run);
        seq_frames = t1 - seq_t0 + a.delay;                    // frames seq_t0 .. t1 + delay - 1
        first = 0;
        last = 1;
    } else {
        first = (blockIdx.x * warps + warp) * a.slots_per_warp;
        last = min(a.n_atoms, first + a.slots_per_warp);
        if (first >= last) return;
    }

    const SoapPlan& pl = a.plan;
    const int n = pl.n_max, n_lm = pl.n_lm;
    [[maybe_unused]] const int P = pl.row_stride;
    const int c_doubles = pl.n_species * n_lm * n;
    // per-warp fixed part: displacement list + cell table (+ batch list of the in-kernel search)
    constexpr int kWarpFixed = soap_warp_fixed(LIST);
    const int per_warp = (kWarpFixed + pl.rbuf_doubles + 1) & ~1;
    double* nbuf = smem + (size_t)warp * per_warp;     // [32][4] dx dy dz r2
    double* tab_sh = nbuf + kListCap * 4;                // [3][32] image shifts of the cell table
    int2* tab_q = reinterpret_cast<int2*>(tab_sh + 3 * kCellTab);  // [32] candidate ranges
    int2* bl_q = reinterpret_cast<int2*>(tab_sh + kCellTab * 4);   // [64] batch (start, count)
    int* bl_e = reinterpret_cast<int*>(bl_q + kBatchCap);             // [64] batch -> table entry
    double* rbuf = nbuf + kWarpFixed;  // [kChunk][P] harmonics; later G / mixed
    // tensor path: the pad columns of the last row's last tile read up to 8 doubles behind the rows;
    // they must be finite (their products are multiplied by a zero weight)
    if (MM) {
        if (lane < 8) rbuf[kChunk * mma_row_stride(LT) + lane] = 0.0;
        // ... and so must the words between the LT (LT + 2) values of a row and its odd stride
        for (int i = lane; i < kChunk * (mma_row_stride(LT) - LT * (LT + 2)); i += 32) {
            constexpr int kPad = mma_row_stride(LT) - LT * (LT + 2);
            rbuf[(i / kPad) * mma_row_stride(LT) + LT * (LT + 2) + i % kPad] = 0.0;
        }
    }

    const int g = lane >> 3, kk = MM ? lane >> 2 : lane & 7;
    const int ji = lane & 3;   // tensor path: neighbour of the lane inside a group of four
    const unsigned lt_mask = (1u << lane) - 1u;
    int il[kMaxItems], ioff[kMaxItems];
#pragma unroll
    for (int t = 0; t < kMaxItems; ++t) {
        il[t] = pl.item_l[g][t];
        ioff[t] = pl.item_off[g][t];
    }

    const int n_pass = SPLIT ? pl.n_species * pl.kblocks : 1;
    // SEQ: the warp's atom and output row are fixed; its sorted slot in frame f + 1 is fetched while
    // frame f is processed (one of the three dependent round trips at the start of a frame)
    int seq_ci = 0, seq_atom = 0, seq_slot_next = 0;
    if (SEQ) {
        seq_atom = a.sid[(size_t)seq_t0 * a.n_atoms + seq_first];
        seq_ci = a.centre_of_atom[seq_atom];
        if (seq_ci < 0) return;
        seq_slot_next = a.slot_of_atom[(size_t)seq_t0 * a.n_atoms + seq_atom];
    }
    double gm[LT + 1];   // tensor path: exponents of the lane (fixed per pass)
    // (also requesting the centre's coordinates and first row codes of frame f + 1 before the
    // epilogue of frame f made the fused kernel 5 % slower: 168 registers are all in use)
    for (int fo = 0; fo < seq_frames; ++fo) {   // one trip unless SEQ
    const int f = SEQ ? seq_t0 + fo : (int)blockIdx.y;
    const int seq_slot = seq_slot_next;
    if (SEQ && fo + 1 < seq_frames)
        seq_slot_next = a.slot_of_atom[(size_t)(f + 1) * a.n_atoms + seq_atom];
    const SoapFrame* __restrict__ fr = a.frames + f;
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
    const int spec_stride = fr->spec_stride;
    for (int pass = 0; pass < n_pass; ++pass) {
    const int z_sp = SPLIT ? pass / pl.kblocks : 0;
    const int kb = SPLIT ? pass - z_sp * pl.kblocks : 0;
    const int cell_base = fr->cell_base + z_sp * spec_stride;  // the species' own cell grid
    const int k = kb * kKBlock + kk;
    const bool kact = k < n;
    int cached_cell = -1, n_batches = -1;
    if (MM && (SPLIT || fo == 0)) {
#pragma unroll
        for (int l = 0; l <= LT; ++l)
            gm[l] = (kact && l <= pl.l_max) ? -a.gamma[l * n + k] * kExpScale : 0.0;
    }

    for (int walk = first; walk < last; ++walk) {
        // SEQ: the warp's atom, wherever this frame's order has put it
        const size_t slot = SEQ ? (size_t)seq_slot : (size_t)f * a.n_atoms + walk;
        const int ci = SEQ ? seq_ci : a.centre_of_atom[a.sid[slot]];
        if (ci < 0) continue;
        const double xc = a.sx[slot], yc = a.sy[slot], zc = a.sz[slot];
        int cc[3];
        soap_cell_of_wrapped(fr, xc, yc, zc, cc);
        double gam[kMaxItems];
#pragma unroll
        for (int t = 0; t < kMaxItems; ++t)
            gam[t] = (!MM && kact && il[t] >= 0) ? -a.gamma[il[t] * n + k] * kExpScale : 0.0;
        double ct[2 * kTiles];   // tensor path: C fragments
        [[maybe_unused]] double accp[kPeel > 0 ? kPeel : 1];   // peeled columns of l = 4, 8, 12
#pragma unroll
        for (int i = 0; i < kPeel; ++i) accp[i] = 0.0;
#pragma unroll
        for (int t = 0; t < 2 * kTiles; ++t) ct[t] = 0.0;
        // l = 0: lane (jq, k) sums exp(-gamma_0k r^2) over the neighbours jj = jq (mod 4)
        [[maybe_unused]] const double gam0 = kact ? -a.gamma[k] * kExpScale : 0.0;
        double acc00 = 0.0;
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
            if constexpr (MM) {
                constexpr int PM = mma_row_stride(LT);
                for (int h0 = 0; h0 < fill; h0 += kChunk) {
                    const int nh = min(kChunk, fill - h0);
                    {   // phase B, two lanes per neighbour; the rows behind the last one are zero
                        const int nbi = lane & (kChunk - 1);
                        const double* d = nbuf + (h0 + nbi) * 4;
                        const bool on = nbi < nh;
                        const double x = on ? d[0] : 0.0, y = on ? d[1] : 0.0;
                        const double z = on ? d[2] : 0.0, r2 = on ? d[3] : 0.0;
                        // neighbour 4 q + i -> row 4 i + q: with an odd stride the 16 stores of a
                        // half-warp here and its 16 (neighbour, column) loads below each hit 16
                        // different bank pairs
                        solid_harmonics_halves<LT>(x, y, z, r2, lane >> 4,
                                                   rbuf + ((nbi & 3) * 4 + (nbi >> 2)) * PM);
                    }
                    __syncwarp();
                    const int ngrp = (nh + 3) >> 2;
                    for (int q = 0; q < ngrp; ++q) {   // phase C, four neighbours per trip
                        const bool on = 4 * q + ji < nh;
                        const double r2 = on ? nbuf[(h0 + 4 * q + ji) * 4 + 3] : 0.0;
                        const double* __restrict__ rt = rbuf + (ji * 4 + q) * PM + (lane >> 2);
                        // no branch on l <= l_max: one basic block (the results for l > l_max are
                        // dropped at the end)
                        double e[LT + 1];
                        exp_scaled_batch<LT + 1>(gm, r2, exp_tab, e);
                        acc00 += on ? e[0] : 0.0;
#pragma unroll
                        for (int l = 1; l <= LT; ++l) {
#pragma unroll
                            for (int t = 0; t < mma_tiles(l); ++t)
                                dmma_8x8x4(ct[2 * (mma_tile_base(l) + t)],
                                           ct[2 * (mma_tile_base(l) + t) + 1],
                                           rt[l * l - 1 + 8 * t], e[l]);
                            // peeled column (m = +l) of the lane's own neighbour (zero row if absent)
                            if (mma_peeled(l))
                                accp[l / 4 - 1] = fma(rt[l * l - 1 + 2 * l - (lane >> 2)], e[l],
                                                      accp[l / 4 - 1]);
                        }
                    }
                    __syncwarp();
                }
            } else {
            for (int h0 = 0; h0 < fill; h0 += kChunk) {
                const int nh = min(kChunk, fill - h0);
                if (lane < nh) {   // phase B
                    const double* d = nbuf + (h0 + lane) * 4;
                    solid_harmonics<LT, PK ? kRowPacked : kRowBlocks>(d[0], d[1], d[2], d[3],
                                                                        rbuf + lane * P);
                }
                __syncwarp();
                double r2_next = nbuf[h0 * 4 + 3];   // one neighbour ahead: hides the shared load
#pragma unroll kUnrollC
                for (int jj = 0; jj < nh; ++jj) {   // phase C
                    const double r2 = r2_next;
                    // the entry behind the last neighbour is read too and never used (at worst
                    // nbuf[32]: the first words of this warp's cell table) -- no clamp in the loop
                    r2_next = nbuf[(h0 + jj + 1) * 4 + 3];
                    const double* __restrict__ rt = rbuf + jj * P;
                    if constexpr (PK) {
                        const double e0 = exp_scaled(gam[0], r2, exp_tab);
                        const double e1 = exp_scaled(gam[1], r2, exp_tab);
                        accumulate_packed(a0, e0, e1, g, rt + g * kPackSlots);
                    } else {
                        if (S0 > 0) accumulate<S0>(a0, exp_scaled(gam[0], r2, exp_tab), rt + ioff[0]);
                        if (S1 > 0) accumulate<S1>(a1, exp_scaled(gam[1], r2, exp_tab), rt + ioff[1]);
                        if (S2 > 0) accumulate<S2>(a2, exp_scaled(gam[2], r2, exp_tab), rt + ioff[2]);
                        if (S3 > 0) accumulate<S3>(a3, exp_scaled(gam[3], r2, exp_tab), rt + ioff[3]);
                    }
                }
#pragma unroll
                for (int q = 0; q < kChunk / 4; ++q) {   // l = 0, a quarter of the chunk per lane
                    const int jj = (lane >> 3) + 4 * q;
                    if (jj < nh) acc00 += exp_scaled(gam0, nbuf[(h0 + jj) * 4 + 3], exp_tab);
                }
                __syncwarp();
            }
            }
        };

        if constexpr (LIST) {
        // ---- gather from the rows of soap_list_kernel: 32 neighbours per round trip ----
        const int my_cell = (cc[0] * nc[1] + cc[1]) * nc[2] + cc[2];
        if (cached_cell != my_cell) {   // image shifts of the 27 cell-table entries
            __syncwarp();
            if (lane < 27) {
                const int idx[3] = {lane / 9, (lane / 3) % 3, lane % 3};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int cj = cc[d] + idx[d] - 1;
                    tab_sh[d * kCellTab + lane] = cj < 0 ? -box[d] : (cj >= nc[d] ? box[d] : 0.0);
                }
            }
            __syncwarp();
            cached_cell = my_cell;
        }
        const int* __restrict__ off = a.rowcnt + slot * (size_t)(pl.n_species + 1) + z_sp;
        const int row_first = off[0];
        const int n_nb = off[1] - row_first;
        const int* __restrict__ row = a.rows + slot * (size_t)a.row_cap + row_first;
        const size_t frame_first = (size_t)f * a.n_atoms;
        // coordinates one batch ahead, list codes two batches ahead (two dependent round trips)
        double px = 0.0, py = 0.0, pz = 0.0;
        int pe = 0;
        auto fetch = [&](int code, int h) {
            if (h < n_nb) {
                const size_t q = frame_first + (size_t)(code & ((1 << kListEntryShift) - 1));
                pe = code >> kListEntryShift;
                px = a.sx[q]; py = a.sy[q]; pz = a.sz[q];
            }
        };
        fetch(lane < n_nb ? row[lane] : 0, lane);
        int code_next = 32 + lane < n_nb ? row[32 + lane] : 0;
        for (int h0 = 0; h0 < n_nb; h0 += 32) {
            const int nh = min(32, n_nb - h0);
            if (lane < nh) {
                const double dx = px + tab_sh[pe] - xc, dy = py + tab_sh[kCellTab + pe] - yc;
                const double dz = pz + tab_sh[2 * kCellTab + pe] - zc;
                double* d = nbuf + lane * 4;
                d[0] = dx; d[1] = dy; d[2] = dz; d[3] = dx * dx + dy * dy + dz * dz;
            }
            fetch(code_next, h0 + 32 + lane);
            code_next = h0 + 64 + lane < n_nb ? row[h0 + 64 + lane] : 0;
            __syncwarp();
            process(nh);
        }
        } else {
            // ---- gather ----
            // The (cell range, image shift) entries around the centre's cell are resolved by the
            // lanes in parallel into a per-warp table; consecutive centres of the same cell reuse it.
            int fill = 0;
            const int n_ent_total = ne[0] * ne[1] * ne[2];
            const int my_cell = (cc[0] * nc[1] + cc[1]) * nc[2] + cc[2];
            for (int e0 = 0; e0 < n_ent_total; e0 += kCellTab) {
                if (!(n_ent_total <= kCellTab && cached_cell == my_cell)) {
                    __syncwarp();
                    const int e = e0 + lane;
                    int qb = 0, qe = 0;
                    double sh[3] = {0.0, 0.0, 0.0};
                    if (e < n_ent_total) {
                        int idx[3];
                        idx[2] = e % ne[2];
                        idx[1] = (e / ne[2]) % ne[1];
                        idx[0] = e / (ne[2] * ne[1]);
                        int cj[3];
                        bool ok = true;
    #pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            if (mode[d] == kNarrow) {
                                cj[d] = 0;
                                sh[d] = (double)(idx[d] - nimg[d]) * box[d];
                            } else {
                                cj[d] = cc[d] + idx[d] - 1;
                                if (mode[d] == kWide) {
                                    if (cj[d] < 0) { cj[d] += nc[d]; sh[d] = -box[d]; }
                                    else if (cj[d] >= nc[d]) { cj[d] -= nc[d]; sh[d] = box[d]; }
                                } else if (cj[d] < 0 || cj[d] >= nc[d]) ok = false;
                            }
                        }
                        if (ok) {
                            const int cell = cell_base + (cj[0] * nc[1] + cj[1]) * nc[2] + cj[2];
                            qb = a.cell_start[cell];
                            qe = a.cell_start[cell + 1];
                        }
                    }
                    tab_q[lane] = make_int2(qb, qe);
                    tab_sh[lane] = sh[0];
                    tab_sh[kCellTab + lane] = sh[1];
                    tab_sh[2 * kCellTab + lane] = sh[2];
                    // flat list of 32-candidate batches so that loads can be prefetched across cells
                    const int nbat = qe > qb ? (qe - qb + 31) >> 5 : 0;
                    int incl = nbat;
    #pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int t = __shfl_up_sync(kFullS, incl, off);
                        if (lane >= off) incl += t;
                    }
                    n_batches = __shfl_sync(kFullS, incl, 31);
                    if (n_batches <= kBatchCap) {
                        for (int i = 0; i < nbat; ++i) {
                            bl_q[incl - nbat + i] = make_int2(qb + 32 * i, min(32, qe - qb - 32 * i));
                            bl_e[incl - nbat + i] = lane;
                        }
                    }
                    __syncwarp();
                    cached_cell = n_ent_total <= kCellTab ? my_cell : -1;
                }
                auto consume = [&](double dx, double dy, double dz, double r2, bool hit) {
                    const unsigned m = __ballot_sync(kFullS, hit);
                    if (m == 0u) return;
                    const int cnt = __popc(m), rank = __popc(m & lt_mask);
                    const int room = kListCap - fill;
                    if (hit && rank < room) {
                        double* d = nbuf + (fill + rank) * 4;
                        d[0] = dx; d[1] = dy; d[2] = dz; d[3] = r2;
                    }
                    if (cnt >= room) {
                        __syncwarp();
                        process(kListCap);
                        if (hit && rank >= room) {
                            double* d = nbuf + (rank - room) * 4;
                            d[0] = dx; d[1] = dy; d[2] = dz; d[3] = r2;
                        }
                        fill = cnt - room;
                    } else {
                        fill += cnt;
                    }
                };
                if (n_batches <= kBatchCap) {
                    // two batches in flight ahead of the one being tested
                    double x1 = 0.0, y1 = 0.0, z1 = 0.0, x2 = 0.0, y2 = 0.0, z2 = 0.0;
                    bool v1 = false, v2 = false;
                    auto fetch = [&](int b, double& x, double& y, double& z, bool& valid) {
                        valid = false;
                        if (b < n_batches) {
                            const int2 q = bl_q[b];
                            if (lane < q.y) {
                                x = a.sx[q.x + lane]; y = a.sy[q.x + lane]; z = a.sz[q.x + lane];
                                valid = true;
                            }
                        }
                    };
                    fetch(0, x1, y1, z1, v1);
                    fetch(1, x2, y2, z2, v2);
                    for (int b = 0; b < n_batches; ++b) {
                        const int e = bl_e[b];
                        const double vx = x1, vy = y1, vz = z1;
                        const bool vv = v1;
                        x1 = x2; y1 = y2; z1 = z2; v1 = v2;
                        fetch(b + 2, x2, y2, z2, v2);
                        const double dx = vx + tab_sh[e] - xc, dy = vy + tab_sh[kCellTab + e] - yc;
                        const double dz = vz + tab_sh[2 * kCellTab + e] - zc;
                        const double r2 = dx * dx + dy * dy + dz * dz;
                        consume(dx, dy, dz, r2, vv && r2 <= a.r2max);
                    }
                } else {
                    const int n_ent = min(kCellTab, n_ent_total - e0);
                    for (int t = 0; t < n_ent; ++t) {
                        const int2 qr = tab_q[t];
                        const int qb = qr.x, qe = qr.y;
                        if (qb >= qe) continue;
                        const double shx = tab_sh[t], shy = tab_sh[kCellTab + t];
                        const double shz = tab_sh[2 * kCellTab + t];
                        for (int q0 = qb; q0 < qe; q0 += 32) {
                            const int q = q0 + lane;
                            double dx = 0.0, dy = 0.0, dz = 0.0, r2 = 0.0;
                            bool hit = false;
                            if (q < qe) {
                                dx = a.sx[q] + shx - xc; dy = a.sy[q] + shy - yc;
                                dz = a.sz[q] + shz - zc;
                                r2 = dx * dx + dy * dy + dz * dz;
                                hit = r2 <= a.r2max;
                            }
                            consume(dx, dy, dz, r2, hit);
                        }
                    }
                }
            }
            __syncwarp();
            if (fill > 0) process(fill);
        }

        // ---- registers -> G[z][lm][k] (shared: fused epilogue; global: split mode) ----
        acc00 += __shfl_xor_sync(kFullS, acc00, MM ? 1 : 8);
        acc00 += __shfl_xor_sync(kFullS, acc00, MM ? 2 : 16);
        double* cz = SPLIT ? a.gscr + ((size_t)f * a.n_cent + ci) * c_doubles +
                                 (size_t)z_sp * n_lm * n
                           : rbuf;
        if constexpr (MM) {
#pragma unroll
            for (int i = 0; i < kPeel; ++i) {
                accp[i] += __shfl_xor_sync(kFullS, accp[i], 1);
                accp[i] += __shfl_xor_sync(kFullS, accp[i], 2);
            }
        }
        if constexpr (MM && SPLIT) {
            // C fragments hold G^T: row = column of the tile (lane >> 2), columns = radial 2 ji + {0, 1}
            if (kact && ji == 0) {
                cz[k] = acc00;  // lm = 0: S_0^0 = 1
#pragma unroll
                for (int i = 0; i < kPeel; ++i) {
                    const int l = 4 * (i + 1);
                    if (l <= pl.l_max) cz[(l * l + 2 * l) * n + k] = accp[i];
                }
            }
            const int k0 = kb * kKBlock + 2 * ji;
#pragma unroll
            for (int l = 1; l <= LT; ++l)
#pragma unroll
                for (int t = 0; t < mma_tiles(l); ++t) {
                    const int col = 8 * t + (lane >> 2);
                    if (col < mma_cols(l) && l <= pl.l_max) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (k0 + h < n)
                                cz[(l * l + col) * n + k0 + h] = ct[2 * (mma_tile_base(l) + t) + h];
                    }
                }
        } else if constexpr (MM) {
            // ---- tensor-path epilogue: nothing leaves the registers ----
            // mixing   D_l (n' x cols) = Bm_l (n' x k) * G_l (k x cols): A = the lane's two entries of
            //          the padded Loewdin matrix, B = the C fragments of phase C as they are (K slot
            //          ji <-> radial index 2 ji + h, one product per h)
            // spectrum P_l (n x n') = sum over the columns of w^2 D[n][col] D[n'][col]: the lane's
            //          D element is both its A entry (times w^2) and its B entry (K slot ji <-> column
            //          2 ji + h), so the Gram matrix needs no data movement either
            const int na = lane >> 2;
            double pv[2 * (LT + 1)];
            const double2* __restrict__ bm2 = reinterpret_cast<const double2*>(bmat_s) + na * 4 + ji;
            const double2* __restrict__ wq2 = reinterpret_cast<const double2*>(swlm_s) + ji;
            {
                const double g0a = __shfl_sync(kFullS, acc00, 8 * ji);       // G_0[2 ji]
                const double g0b = __shfl_sync(kFullS, acc00, 8 * ji + 4);   // G_0[2 ji + 1]
                const double2 bm = bm2[0], w = wq2[0];
                double d0 = 0.0, d1 = 0.0, p0 = 0.0, p1 = 0.0;
                dmma_8x8x4(d0, d1, bm.x, na == 0 ? g0a : 0.0);
                dmma_8x8x4(d0, d1, bm.y, na == 0 ? g0b : 0.0);
                dmma_8x8x4(p0, p1, d0 * w.x, d0);
                dmma_8x8x4(p0, p1, d1 * w.y, d1);
                pv[0] = p0;
                pv[1] = p1;
            }
#pragma unroll
            for (int l = 1; l <= LT; ++l) {
                double p0 = 0.0, p1 = 0.0;
                if (l <= pl.l_max) {
                    const double2 bm = bm2[l * 32];
#pragma unroll
                    for (int t = 0; t < mma_tiles(l); ++t) {
                        const double2 w = wq2[(1 + mma_tile_base(l) + t) * 4];
                        double d0 = 0.0, d1 = 0.0;
                        dmma_8x8x4(d0, d1, bm.x, ct[2 * (mma_tile_base(l) + t)]);
                        dmma_8x8x4(d0, d1, bm.y, ct[2 * (mma_tile_base(l) + t) + 1]);
                        dmma_8x8x4(p0, p1, d0 * w.x, d0);
                        dmma_8x8x4(p0, p1, d1 * w.y, d1);
                    }
                    if (mma_peeled(l)) {   // the peeled column as a one-column tile, like l = 0
                        const double ga = __shfl_sync(kFullS, accp[l / 4 - 1], 8 * ji);
                        const double gb = __shfl_sync(kFullS, accp[l / 4 - 1], 8 * ji + 4);
                        const double2 w = wq2[(kTiles + l / 4) * 4];
                        double d0 = 0.0, d1 = 0.0;
                        dmma_8x8x4(d0, d1, bm.x, na == 0 ? ga : 0.0);
                        dmma_8x8x4(d0, d1, bm.y, na == 0 ? gb : 0.0);
                        dmma_8x8x4(p0, p1, d0 * w.x, d0);
                        dmma_8x8x4(p0, p1, d1 * w.y, d1);
                    }
                }
                pv[2 * l] = p0;
                pv[2 * l + 1] = p1;
            }
            // the lane holds P_l[na][2 ji + {0, 1}]; features of one l: rows na, columns nb >= na
            const int nb0 = 2 * ji, blk = n * (n + 1) / 2;
            const int c_first = na * n - (na * (na - 1)) / 2 + nb0 - na;
            const bool ok0 = na <= nb0 && nb0 < n, ok1 = na <= nb0 + 1 && nb0 + 1 < n;
            double* out_row =
                a.out ? a.out + (long long)ci * a.stride_i + (long long)f * a.stride_f : nullptr;
            if constexpr (SEQ) {
                // ring slot f mod delay still holds frame f - delay (if the run got that far)
                double* ring = a.ring + (((size_t)blockIdx.y * a.n_cent + ci) * a.delay +
                                         (size_t)(f % a.delay)) * (size_t)(pl.n_feat + 1);
                // runs overlap by `delay` frames: the spectrum row of a frame (if wanted) is
                // written by the run its index falls into
                if (min(f / a.seq_run, (int)gridDim.y - 1) != (int)blockIdx.y) out_row = nullptr;
                double rv[2 * (LT + 1)];
#pragma unroll
                for (int l = 0; l <= LT; ++l) {   // all the ring loads before the ring stores
                    const int c = l * blk + c_first;
                    rv[2 * l] = (ok0 && l <= pl.l_max) ? ring[c] : 0.0;
                    rv[2 * l + 1] = (ok1 && l <= pl.l_max) ? ring[c + 1] : 0.0;
                }
                double csq = 0.0, dot = 0.0;
#pragma unroll
                for (int l = 0; l <= LT; ++l) {
                    const int c = l * blk + c_first;
                    if (ok0 && l <= pl.l_max) {
                        csq = fma(pv[2 * l], pv[2 * l], csq);
                        dot = fma(pv[2 * l], rv[2 * l], dot);
                        ring[c] = pv[2 * l];
                    }
                    if (ok1 && l <= pl.l_max) {
                        csq = fma(pv[2 * l + 1], pv[2 * l + 1], csq);
                        dot = fma(pv[2 * l + 1], rv[2 * l + 1], dot);
                        ring[c + 1] = pv[2 * l + 1];
                    }
                }
                csq = warp_sum(csq);
                dot = warp_sum(dot);
                if (lane == 0) {
                    double* ring_sq = ring + pl.n_feat;
                    if (fo >= a.delay)
                        a.ts_out[(long long)ci * a.ts_ld + a.ts_col0 + (f - a.delay)] =
                            soap_dist_from_sums(*ring_sq, csq, dot);
                    *ring_sq = csq;
                }
            }
            if (out_row) {
#pragma unroll
                for (int l = 0; l <= LT; ++l) {
                    const int c = l * blk + c_first;
                    if (ok0 && l <= pl.l_max) out_row[c] = pv[2 * l];
                    if (ok1 && l <= pl.l_max) out_row[c + 1] = pv[2 * l + 1];
                }
            }
            __syncwarp();
        } else if (kact) {
            if (g == 0) cz[k] = acc00;  // lm = 0: S_0^0 = 1
            if constexpr (PK) store_packed(a0, g, k, n, cz);
            else store_item<S0>(a0, il[0], k, n, cz);
            store_item<S1>(a1, il[1], k, n, cz);
            store_item<S2>(a2, il[2], k, n, cz);
            store_item<S3>(a3, il[3], k, n, cz);
        }
        __syncwarp();
        if (!SPLIT && !MM) {
            // the harmonics buffer is free now: G sits at its start, the mixed coefficients follow
            SpectrumSink sink;
            sink.out = a.out ? a.out + (long long)ci * a.stride_i + (long long)f * a.stride_f
                             : nullptr;
            sink.ring = nullptr;
            sink.csq = sink.dot = 0.0;
            double* ring_sq = nullptr;
            if (SEQ) {
                // ring slot f mod delay still holds frame f - delay (if the run got that far)
                double* base = a.ring + (((size_t)blockIdx.y * a.n_cent + ci) * a.delay +
                                         (size_t)(f % a.delay)) * (size_t)(pl.n_feat + 1);
                sink.ring = base;
                ring_sq = base + pl.n_feat;
                // runs overlap by `delay` frames: the spectrum row of a frame (if wanted) is
                // written by the run its index falls into
                if (min(f / a.seq_run, (int)gridDim.y - 1) != (int)blockIdx.y) sink.out = nullptr;
            }
            mix_and_spectrum(a, rbuf, rbuf + ((c_doubles + 1) & ~1), bmat, swlm_s, lm_l_s,
                             decode_s, sink, lane);
            if (SEQ) {
                const double csq = warp_sum(sink.csq), dot = warp_sum(sink.dot);
                if (lane == 0) {
                    if (fo >= a.delay)
                        a.ts_out[(long long)ci * a.ts_ld + a.ts_col0 + (f - a.delay)] =
                            soap_dist_from_sums(*ring_sq, csq, dot);
                    *ring_sq = csq;
                }
                __syncwarp();
            }
        }
    }  // slot loop
    }  // pass loop
    }  // frame loop (SEQ)
}

// Split mode, second kernel: one warp per (centre, frame), one angular momentum at a time.  The
// S(2l+1) columns of G for that l are read from the scratch (one column per lane, coalesced),
// mixed into shared memory, and the (species pair, n, n') outputs of that l are computed as 2x2
// register tiles (two 128-bit shared loads per four FMAs).  HBM-bound: 8 S n_lm n bytes read +
// 8 C bytes written per centre; only S(2 l_max+1) n doubles of shared memory per warp.
struct SpecTask {
    int code;   // zi | zj << 4 | na0 << 8 | nb0 << 16
    int base;   // first feature of the species pair
    int blk;    // features per l in that pair
    int pad;
};

template <bool SEQ>
__global__ void __launch_bounds__(128)
soap_spectrum_kernel(const SoapArgs a, const SpecTask* __restrict__ tasks, int n_tasks) {
    extern __shared__ __align__(16) double smem_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* bmat_s = smem_all;
    for (int i = threadIdx.x; i < a.bmat_smem_doubles; i += blockDim.x) bmat_s[i] = a.bmat[i];
    __syncthreads();
    const double* __restrict__ bmat = a.bmat_smem_doubles > 0 ? bmat_s : a.bmat;
    const SoapPlan& pl = a.plan;
    const int n = pl.n_max, n_lm = pl.n_lm, l_max = pl.l_max, S = pl.n_species;
    const int np_stride = (n + 1) & ~1;
    // per warp: mixed coefficients of one l + the outputs of that l (all species pairs, in
    // feature order); block-wide: feature index at l = 0 and per-l stride of every output
    const int n_l = pl.n_feat / (l_max + 1);
    const int mix_doubles = S * (2 * l_max + 1) * np_stride;
    const int warp_doubles = (mix_doubles + n_l + 1) & ~1;
    double* mix = smem_all + a.bmat_smem_doubles + (size_t)warp * warp_doubles;
    double* lst = mix + mix_doubles;
    int* c0_s = reinterpret_cast<int*>(smem_all + a.bmat_smem_doubles +
                                       (size_t)(blockDim.x >> 5) * warp_doubles);
    int* blk_s = c0_s + n_l;
    for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) {
        const SpecTask tk = tasks[t];
        const bool diag = (tk.code & 15) == ((tk.code >> 4) & 15);
        const int na0 = (tk.code >> 8) & 255, nb0 = (tk.code >> 16) & 255;
        const int seg = tk.base / (l_max + 1);   // start of the pair inside one l
        for (int q = 0; q < 4; ++q) {
            const int na = na0 + (q >> 1), nb = nb0 + (q & 1);
            if (na < n && nb < n && (!diag || nb >= na)) {
                const int off = diag ? na * n - (na * (na - 1)) / 2 + nb - na : na * n + nb;
                c0_s[seg + off] = tk.base + off;
                blk_s[seg + off] = tk.blk;
            }
        }
    }
    __syncthreads();
    const int ci = blockIdx.x * (blockDim.x >> 5) + warp;
    if (ci >= a.n_cent) return;   // warps are independent from here: only __syncwarp below
    // SEQ: the warp walks the frames of its run in order and folds every spectrum into the
    // timeSOAP sums against its ring (see soap_kernel); else one (centre, frame) per warp
    const int t0 = SEQ ? blockIdx.y * a.seq_run : 0;
    const int n_fr = SEQ ? min(a.n_frames - a.delay, t0 + a.seq_run) - t0 + a.delay : 1;
    for (int fo = 0; fo < n_fr; ++fo) {
    const int f = SEQ ? t0 + fo : (int)blockIdx.y;
    const double* __restrict__ gsum = a.gscr + ((size_t)f * a.n_cent + ci) * S * n_lm * n;
    SpectrumSink sink;
    sink.out = a.out ? a.out + (long long)ci * a.stride_i + (long long)f * a.stride_f : nullptr;
    sink.ring = nullptr;
    sink.csq = sink.dot = 0.0;
    double* ring_sq = nullptr;
    if (SEQ) {
        double* base = a.ring + (((size_t)blockIdx.y * a.n_cent + ci) * a.delay +
                                 (size_t)(f % a.delay)) * (size_t)(pl.n_feat + 1);
        sink.ring = base;
        ring_sq = base + pl.n_feat;
        if (min(f / a.seq_run, (int)gridDim.y - 1) != (int)blockIdx.y) sink.out = nullptr;
    }

    for (int l = 0; l <= l_max; ++l) {
        const int nm = 2 * l + 1, ncol = S * nm;
        const double* __restrict__ bl = bmat + (size_t)l * n * n;
        for (int col = lane; col < ncol; col += 32) {
            const int z = col / nm, m = col - z * nm;
            const double w = a.swlm[l * l + m];
            const double* __restrict__ v = gsum + ((size_t)z * n_lm + l * l + m) * n;
            double* __restrict__ o = mix + (size_t)col * np_stride;
            if (n == 8) {
                const double2* __restrict__ v2 = reinterpret_cast<const double2*>(v);
                const double2 g0 = v2[0], g1 = v2[1], g2 = v2[2], g3 = v2[3];
                const double2* __restrict__ b2 = reinterpret_cast<const double2*>(bl);
#pragma unroll
                for (int np = 0; np < 8; np += 2) {
                    double r[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double2 b0 = b2[(np + h) * 4], b1 = b2[(np + h) * 4 + 1];
                        const double2 b3 = b2[(np + h) * 4 + 2], b4 = b2[(np + h) * 4 + 3];
                        double sa = b0.x * g0.x, sb = b0.y * g0.y;
                        sa = fma(b1.x, g1.x, sa); sb = fma(b1.y, g1.y, sb);
                        sa = fma(b3.x, g2.x, sa); sb = fma(b3.y, g2.y, sb);
                        sa = fma(b4.x, g3.x, sa); sb = fma(b4.y, g3.y, sb);
                        r[h] = (sa + sb) * w;
                    }
                    reinterpret_cast<double2*>(o)[np >> 1] = make_double2(r[0], r[1]);
                }
            } else {
                for (int np = 0; np < n; ++np) {
                    const double* __restrict__ bm = bl + (size_t)np * n;
                    double s0 = 0.0, s1 = 0.0;
                    int k = 0;
                    for (; k + 1 < n; k += 2) {
                        s0 = fma(bm[k], v[k], s0);
                        s1 = fma(bm[k + 1], v[k + 1], s1);
                    }
                    if (k < n) s0 = fma(bm[k], v[k], s0);
                    o[np] = (s0 + s1) * w;
                }
                if (np_stride > n) o[n] = 0.0;
            }
        }
        __syncwarp();
        for (int t = lane; t < n_tasks; t += 32) {
            const SpecTask tk = tasks[t];
            const int zi = tk.code & 15, zj = (tk.code >> 4) & 15;
            const int na0 = (tk.code >> 8) & 255, nb0 = (tk.code >> 16) & 255;
            const double2* __restrict__ ca =
                reinterpret_cast<const double2*>(mix + (size_t)zi * nm * np_stride + na0);
            const double2* __restrict__ cb =
                reinterpret_cast<const double2*>(mix + (size_t)zj * nm * np_stride + nb0);
            double o00 = 0.0, o01 = 0.0, o10 = 0.0, o11 = 0.0;
            const int step = np_stride >> 1;
            for (int m = 0; m < nm; ++m) {
                const double2 u = ca[m * step], v = cb[m * step];
                o00 = fma(u.x, v.x, o00);
                o01 = fma(u.x, v.y, o01);
                o10 = fma(u.y, v.x, o10);
                o11 = fma(u.y, v.y, o11);
            }
            double* __restrict__ ob = lst + tk.base / (l_max + 1);
            const bool diag = zi == zj;
            const double res[4] = {o00, o01, o10, o11};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int na = na0 + (q >> 1), nb = nb0 + (q & 1);
                if (na < n && nb < n && (!diag || nb >= na)) {
                    const int off = diag ? na * n - (na * (na - 1)) / 2 + nb - na : na * n + nb;
                    ob[off] = res[q];
                }
            }
        }
        __syncwarp();
        // flush the staged outputs of this l in feature order (coalesced); ring loads first
        if (sink.ring) {
            for (int idx = lane; idx < n_l; idx += 32) {
                const double v = lst[idx];
                sink.csq = fma(v, v, sink.csq);
                sink.dot = fma(v, sink.ring[c0_s[idx] + l * blk_s[idx]], sink.dot);
            }
            for (int idx = lane; idx < n_l; idx += 32)
                sink.ring[c0_s[idx] + l * blk_s[idx]] = lst[idx];
        }
        if (sink.out)
            for (int idx = lane; idx < n_l; idx += 32)
                sink.out[c0_s[idx] + l * blk_s[idx]] = lst[idx];
        __syncwarp();
    }
    if (SEQ) {
        const double csq = warp_sum(sink.csq), dot = warp_sum(sink.dot);
        if (lane == 0) {
            if (fo >= a.delay)
                a.ts_out[(long long)ci * a.ts_ld + a.ts_col0 + (f - a.delay)] =
                    soap_dist_from_sums(*ring_sq, csq, dot);
            *ring_sq = csq;
        }
        __syncwarp();
    }
    }  // frame loop (SEQ)
}

// Split mode, second kernel on the fp64 tensor path (n_max <= 8, up to 4 species): the same
// register-to-register products as the fused epilogue of soap_kernel, per angular momentum:
//   mixing    D_z (n' x cols) = Bm_l (n' x k) * G_z,l (k x cols): lane (col = lane >> 2, ji) takes
//             G[z][l^2 + col][2 ji, 2 ji + 1] as its B entries, two products per 8-column tile
//   spectrum  P_zz' (n x n') = sum over the columns of w^2 D_z[n][col] D_z'[n'][col] for z <= z'
// One block of four warps per centre, warp w takes l = w, w + 4, ....  The primitive sums of a
// (centre, frame) -- 8 S (l_max+1)^2 n bytes, 23 KB at cfg4 -- are copied from the scratch to
// shared memory with cp.async; SEQ: the block walks the frames of its run, the copy of frame f + 1
// is in flight while frame f is processed (the version that loaded the fragments straight from
// global memory spent its time waiting for them: 9 % of the HBM bandwidth, tensor pipe 29 % busy),
// and every spectrum is folded into the timeSOAP sums (partial sums of the four warps meet in
// shared memory).  RS (SEQ, delay = 1): the previous spectrum of the centre stays in shared memory,
// one lane-private slot per (l, species pair, element) -- the global ring cost 2 x 8 C bytes of HBM
// traffic per centre-frame (C = 3300 at cfg4: 21 of the 30 GB the kernel moved).
constexpr int kSpecWarps = 4;
__device__ __forceinline__ void cp_async_bytes(double* dst_smem, const double* src, int n_doubles,
                                               int tid, int n_threads) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(dst_smem);
    if ((n_doubles & 1) == 0 && (((size_t)src) & 15) == 0) {
        for (int i = tid; i < n_doubles / 2; i += n_threads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i),
                         "l"(src + 2 * i));
    } else {
        for (int i = tid; i < n_doubles; i += n_threads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * i),
                         "l"(src + i));
    }
    asm volatile("cp.async.commit_group;");
}

// one angular momentum with NT tiles (compile time: exact product counts, no predication)
template <int S, int NT, bool SEQ, bool RS>
__device__ __forceinline__ void spectrum_one_l(const SoapArgs& a, int l, const double* __restrict__ gs,
                                               const double* __restrict__ bm8,
                                               const double* __restrict__ wq, double* ring,
                                               double* ring_s, double* out_row, int lane,
                                               double& csq, double& dot) {
    constexpr int kPairs = S * (S + 1) / 2;
    const int n = a.plan.n_max, n_lm = a.plan.n_lm, l_max = a.plan.l_max;
    const int na = lane >> 2, nb0 = 2 * (lane & 3);
    const int blk_d = n * (n + 1) / 2, blk_o = n * n;
    // offsets of the lane's two elements inside one l block of a diagonal / off-diagonal pair
    const int off_d = na * n - (na * (na - 1)) / 2 + nb0 - na, off_o = na * n + nb0;
    const bool okd0 = na <= nb0 && nb0 < n, okd1 = na <= nb0 + 1 && nb0 + 1 < n;
    const bool oko0 = na < n && nb0 < n, oko1 = na < n && nb0 + 1 < n;
    const double2 bm = *reinterpret_cast<const double2*>(bm8 + l * 64 + na * 8 + nb0);
    // the ring values of this l first: none of their loads waits behind a ring store
    [[maybe_unused]] double rv[kPairs][2];
    if (SEQ) {
        int base = 0, q = 0;
#pragma unroll
        for (int zi = 0; zi < S; ++zi)
#pragma unroll
            for (int zj = zi; zj < S; ++zj, ++q) {
                const bool diag = zi == zj;
                const int c = base + l * (diag ? blk_d : blk_o) + (diag ? off_d : off_o);
                if (RS) {
                    rv[q][0] = ring_s[((l * kPairs + q) * 2) * 32];
                    rv[q][1] = ring_s[((l * kPairs + q) * 2 + 1) * 32];
                } else {
                    rv[q][0] = (diag ? okd0 : oko0) ? ring[c] : 0.0;
                    rv[q][1] = (diag ? okd1 : oko1) ? ring[c + 1] : 0.0;
                }
                base += (diag ? blk_d : blk_o) * (l_max + 1);
            }
    }
    double d[S][NT][2];
#pragma unroll
    for (int z = 0; z < S; ++z)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int col = 8 * t + na;
            double g0 = 0.0, g1 = 0.0;
            if (col < 2 * l + 1) {
                const double* __restrict__ gp = gs + ((size_t)z * n_lm + l * l + col) * n + nb0;
                if (n == 8) {
                    const double2 g = *reinterpret_cast<const double2*>(gp);
                    g0 = g.x;
                    g1 = g.y;
                } else {
                    if (nb0 < n) g0 = gp[0];
                    if (nb0 + 1 < n) g1 = gp[1];
                }
            }
            d[z][t][0] = d[z][t][1] = 0.0;
            dmma_8x8x4(d[z][t][0], d[z][t][1], bm.x, g0);
            dmma_8x8x4(d[z][t][0], d[z][t][1], bm.y, g1);
        }
    int base = 0, q = 0;   // first feature of the species pair, its number
#pragma unroll
    for (int zi = 0; zi < S; ++zi)
#pragma unroll
        for (int zj = zi; zj < S; ++zj, ++q) {
            double p0 = 0.0, p1 = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const double2 w = *reinterpret_cast<const double2*>(wq + (l * 4 + t) * 8 + nb0);
                dmma_8x8x4(p0, p1, d[zi][t][0] * w.x, d[zj][t][0]);
                dmma_8x8x4(p0, p1, d[zi][t][1] * w.y, d[zj][t][1]);
            }
            const bool diag = zi == zj;
            const int c = base + l * (diag ? blk_d : blk_o) + (diag ? off_d : off_o);
            const bool ok0 = diag ? okd0 : oko0, ok1 = diag ? okd1 : oko1;
            if (SEQ) {
                if (ok0) {
                    csq = fma(p0, p0, csq);
                    dot = fma(p0, rv[q][0], dot);
                    if (RS) ring_s[((l * kPairs + q) * 2) * 32] = p0;
                    else ring[c] = p0;
                }
                if (ok1) {
                    csq = fma(p1, p1, csq);
                    dot = fma(p1, rv[q][1], dot);
                    if (RS) ring_s[((l * kPairs + q) * 2 + 1) * 32] = p1;
                    else ring[c + 1] = p1;
                }
            }
            if (out_row) {
                if (ok0) out_row[c] = p0;
                if (ok1) out_row[c + 1] = p1;
            }
            base += (diag ? blk_d : blk_o) * (l_max + 1);
        }
}

template <int S, bool SEQ, bool RS>
__global__ void __launch_bounds__(32 * kSpecWarps)
soap_spectrum_mma_kernel(const SoapArgs a) {
    extern __shared__ __align__(16) double smem_all[];
    const SoapPlan& pl = a.plan;
    const int n = pl.n_max, n_lm = pl.n_lm, l_max = pl.l_max;
    constexpr int kPairs = S * (S + 1) / 2;
    const int g_doubles = S * n_lm * n, g_stride = (g_doubles + 1) & ~1;
    double* bm8 = smem_all;                        // [l][n'][k] padded to 8 x 8
    double* wq = bm8 + (l_max + 1) * 64;           // [l][4 tiles][8] squared m-weights, 0 in the pads
    double* part = wq + (l_max + 1) * 32;          // [2][kSpecWarps][2] partial (csq, dot)
    double* gbuf = part + 4 * kSpecWarps;          // [SEQ ? 2 : 1][g_stride] primitive sums
    [[maybe_unused]] double* ring_s = gbuf + (SEQ ? 2 : 1) * g_stride + (threadIdx.x & 31);
    const int ci = blockIdx.x;
    const int t0 = SEQ ? blockIdx.y * a.seq_run : 0;
    const int n_fr = SEQ ? min(a.n_frames - a.delay, t0 + a.seq_run) - t0 + a.delay : 1;
    const int f_first = SEQ ? t0 : (int)blockIdx.y;
    cp_async_bytes(gbuf, a.gscr + ((size_t)f_first * a.n_cent + ci) * g_doubles, g_doubles,
                   threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < (l_max + 1) * 64; i += blockDim.x) {
        const int l = i >> 6, np = (i >> 3) & 7, kq = i & 7;
        bm8[i] = (np < n && kq < n) ? a.bmat[((size_t)l * n + np) * n + kq] : 0.0;
    }
    for (int i = threadIdx.x; i < (l_max + 1) * 32; i += blockDim.x) {
        const int l = i >> 5, col = i & 31;
        const double w = col < 2 * l + 1 ? a.swlm[l * l + col] : 0.0;
        wq[i] = w * w;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    [[maybe_unused]] double prev_sq = 0.0;
    for (int fo = 0; fo < n_fr; ++fo) {
        const int f = f_first + fo;
        const double* gs = gbuf + (fo & 1) * g_stride;
        // frame f has landed (and the tables, first trip); every warp is done with frame f - 1,
        // whose buffer the copy of frame f + 1 overwrites
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (SEQ && fo + 1 < n_fr)
            cp_async_bytes(gbuf + ((fo + 1) & 1) * g_stride,
                           a.gscr + ((size_t)(f + 1) * a.n_cent + ci) * g_doubles, g_doubles,
                           threadIdx.x, blockDim.x);
        double* out_row =
            a.out ? a.out + (long long)ci * a.stride_i + (long long)f * a.stride_f : nullptr;
        double* ring = nullptr;
        if (SEQ) {
            if (!RS)
                ring = a.ring + (((size_t)blockIdx.y * a.n_cent + ci) * a.delay +
                                 (size_t)(f % a.delay)) * (size_t)(pl.n_feat + 1);
            if (min(f / a.seq_run, (int)gridDim.y - 1) != (int)blockIdx.y) out_row = nullptr;
        }
        double csq = 0.0, dot = 0.0;
        for (int l = warp; l <= l_max; l += kSpecWarps) {
            const int nt = (2 * l + 1 + 7) >> 3;
            if (nt == 1)
                spectrum_one_l<S, 1, SEQ, RS>(a, l, gs, bm8, wq, ring, ring_s, out_row, lane, csq, dot);
            else if (nt == 2)
                spectrum_one_l<S, 2, SEQ, RS>(a, l, gs, bm8, wq, ring, ring_s, out_row, lane, csq, dot);
            else if (nt == 3)
                spectrum_one_l<S, 3, SEQ, RS>(a, l, gs, bm8, wq, ring, ring_s, out_row, lane, csq, dot);
            else
                spectrum_one_l<S, 4, SEQ, RS>(a, l, gs, bm8, wq, ring, ring_s, out_row, lane, csq, dot);
        }
        if (SEQ) {
            csq = warp_sum(csq);
            dot = warp_sum(dot);
            double* pp = part + (fo & 1) * 2 * kSpecWarps;   // double-buffered: the barrier at the top
            if (lane == 0) {                                  // of the next trip separates the uses
                pp[2 * warp] = csq;
                pp[2 * warp + 1] = dot;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double cs = 0.0, dt = 0.0;
#pragma unroll
                for (int w = 0; w < kSpecWarps; ++w) {
                    cs += pp[2 * w];
                    dt += pp[2 * w + 1];
                }
                double* ring_sq = RS ? &prev_sq : ring + pl.n_feat;
                if (fo >= a.delay)
                    a.ts_out[(long long)ci * a.ts_ld + a.ts_col0 + (f - a.delay)] =
                        soap_dist_from_sums(*ring_sq, cs, dt);
                *ring_sq = cs;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct SoapLayout {
    long long total_cells = 0;
    bool periodic = false;
    size_t off_frames = 0, off_wpos = 0, off_cell_id = 0, off_cnt = 0, off_start = 0;
    size_t off_sx = 0, off_sy = 0, off_sz = 0, off_sid = 0, off_cmap = 0;
    size_t off_tile = 0, off_gscr = 0, off_tasks = 0;
    int n_tasks = 0;
    size_t off_gamma = 0, off_bmat = 0, off_wlm = 0, off_lml = 0, off_decode = 0, off_exptab = 0;
    // list mode (soap_list_kernel): fused single-species kernel, periodic, >= 3 cells per axis
    int row_cap = 0;  // 0 = the main kernel searches by itself
    size_t off_rows = 0, off_rowcnt = 0, off_listflag = 0;
    // frame-sequential mode (fused timeSOAP): sorted slot of every atom, per-warp rings
    int seq_run = 0, n_runs = 0;
    size_t off_slot_of = 0, off_ring = 0;
    size_t bytes = 0;
};

// several (species, radial block) passes per centre -> accumulate kernel + spectrum kernel
static bool soap_split(int n_species, int n_max) {
    return n_species * ((n_max + kKBlock - 1) / kKBlock) > 1;
}

static double soap_radius(double r_cut) { return r_cut + std::sqrt(-2.0 * std::log(1e-3)); }

// Shared memory of soap_spectrum_mma_kernel: tables + partial sums + one (fused: two) buffers of
// primitive sums, and -- fused with delay = 1, if a block then still fits twice per SM -- the ring
// of previous spectra: 64 doubles per (l, species pair).  ring_in_smem: no global ring is needed.
constexpr int kSpecWarpsHost = 4;
static bool soap_spectrum_mma_applies(int n_species, int n_max) {
    return DSG_SOAP_MMA && soap_split(n_species, n_max) && n_max <= 8 && n_species <= 4;
}
static size_t soap_spectrum_mma_smem(int n_species, int n_max, int l_max, bool seq, int delay,
                                     bool* ring_in_smem) {
    const size_t n_lm = (size_t)(l_max + 1) * (l_max + 1);
    const size_t g_stride = ((size_t)n_species * n_lm * n_max + 1) & ~(size_t)1;
    const size_t tab_bytes =
        8 * ((size_t)(l_max + 1) * 96 + 4 * kSpecWarpsHost + (seq ? 2 : 1) * g_stride);
    const size_t ring_bytes = 8 * (size_t)(l_max + 1) * (n_species * (n_species + 1) / 2) * 64;
    const bool rs = seq && delay == 1 && tab_bytes + ring_bytes <= 100 * 1024;
    if (ring_in_smem) *ring_in_smem = rs;
    return tab_bytes + (rs ? ring_bytes : 0);
}

// delay > 0: layout of the fused SOAP -> timeSOAP call (dsg_soap_timesoap)
static int soap_layout(int n_frames, int n_atoms, int n_cent, int n_species, int n_max, int l_max,
                       const double* h_box, double r_cut, SoapLayout* L,
                       std::vector<SoapFrame>* frames, int delay = 0) {
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
    // gamma < eta = 1/2, so gamma r^2 < R^2 / 2 must stay below 690 (see exp_scaled)
    if (0.5 * soap_radius(r_cut) * soap_radius(r_cut) > 690.0)
        return set_error(DSG_ERR_UNSUPPORTED, "r_cut=%g: search radius above 37 A is unsupported",
                         r_cut);
    if ((long long)n_frames * n_atoms >= 0x7fffffffLL)
        return set_error(DSG_ERR_OVERFLOW, "n_frames*n_atoms must stay below 2^31 per batch");
    const double r_cell = soap_radius(r_cut) * (1.0 + 1e-9);
    L->periodic = h_box != nullptr;
    long long total = 0;
    bool any_narrow = false;
    double k_max = 0.0;   // expected neighbours inside the search radius at the densest frame
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
                any_narrow = any_narrow || fr.mode[k] == kNarrow;
            }
            const double rs = soap_radius(r_cut);
            const double k_est = (double)n_atoms * 4.18879020478639 * rs * rs * rs /
                                 (h_box[3 * f] * h_box[3 * f + 1] * h_box[3 * f + 2]);
            if (k_est > k_max) k_max = k_est;
        } else {
            cells = (long long)kOpenCellsPerAxis * kOpenCellsPerAxis * kOpenCellsPerAxis;
            for (int k = 0; k < 3; ++k) {
                fr.lo_bits[k] = 0xffffffffu;
                fr.hi_bits[k] = 0u;
            }
        }
        fr.cell_base = (int)total;
        fr.spec_stride = (int)cells;
        total += cells * n_species;
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
    L->off_sid = take(4 * fa);
    L->off_cmap = take(4 * (size_t)n_atoms);
    L->off_tile = take(4 * (((size_t)total + kScanTileItems - 1) / kScanTileItems + 1));
    L->off_gamma = take(8 * (size_t)(l_max + 1) * n_max);
    L->off_bmat = take(8 * ((size_t)(l_max + 1) * n_max * n_max + 2));
    L->off_wlm = take(8 * (size_t)n_lm);
    L->off_lml = take(4 * (size_t)n_lm);
    L->off_decode = take(4 * (size_t)n_feat);
    L->off_exptab = take(8 * (size_t)kExpTab);
    if (soap_split(n_species, n_max)) {  // primitive sums between the two kernels
        L->off_gscr = take(8 * (size_t)n_frames * n_cent * n_species * n_lm * n_max);
        const int tiles = (n_max + 1) / 2;
        L->n_tasks = n_species * (tiles * (tiles + 1) / 2) +
                     n_species * (n_species - 1) / 2 * tiles * tiles;
        L->off_tasks = take(sizeof(SpecTask) * (size_t)L->n_tasks);
    }
    // rows sized for 1.5 x the mean neighbour count + 64 (a Poisson gas of 100 neighbours stays
    // 6 sigma below that); a denser-than-expected region only costs the fallback launch
    // The rows are indexed by sorted slot, i.e. allocated for every atom: with few centres in a large
    // system (fewer than a quarter of the atoms and more than 256 MiB of rows) they would be mostly
    // unused memory, and the in-kernel search only visits the centres anyway.
    const int cap = ((int)(1.5 * k_max) + 64 + 31) & ~31;
    const bool sparse_centres = (long long)n_cent * 4 < (long long)n_atoms &&
                                4 * fa * (size_t)cap > ((size_t)256 << 20);
    if (DSG_SOAP_LIST && h_box && !any_narrow && n_atoms < (1 << kListEntryShift) &&
        k_max < 600.0 && !sparse_centres) {
        L->row_cap = cap;
        L->off_rows = take(4 * fa * (size_t)L->row_cap);
        L->off_rowcnt = take(4 * fa * (size_t)(n_species + 1));
        L->off_listflag = take(4);
    }
    if (delay > 0) {
        if (delay >= n_frames)
            return set_error(DSG_ERR_INVALID_ARGUMENT, "delay=%d must be below n_frames=%d", delay,
                             n_frames);
        // frame pairs per warp: long runs (the `delay` frames a run shares with the next one are
        // computed twice) as long as the grid still has ~4 generations of resident warps
        const long long pairs = n_frames - delay;
        long long run = pairs * n_cent / (148LL * 16 * 4);
        if (run > 64) run = 64;
        if (run < 8LL * delay) run = 8LL * delay;
        if (run > pairs) run = pairs;
        L->seq_run = (int)run;
        L->n_runs = (int)((pairs + run - 1) / run);
        L->off_slot_of = take(4 * fa);
        // (the split-mode tensor-path spectrum kernel keeps a delay-1 ring in shared memory: 6.6 GB of
        // workspace less at cfg4)
        bool ring_in_smem = false;
        if (soap_spectrum_mma_applies(n_species, n_max))
            soap_spectrum_mma_smem(n_species, n_max, l_max, true, delay, &ring_in_smem);
        if (!ring_in_smem)
            L->off_ring = take(8 * (size_t)L->n_runs * n_cent * delay * ((size_t)n_feat + 1));
    }
    L->bytes = o;
    return DSG_OK;
}

// lt = compile-time l cap of the instantiation, s0 = its widest item
static void make_plan(int n_species, int n_max, int l_max, int lt, int s0, SoapPlan* pl) {
    const bool packed = !DSG_SOAP_MMA && lt == 8 && s0 == kPackSlots;
    pl->n_max = n_max;
    pl->l_max = l_max;
    pl->n_species = n_species;
    pl->n_feat = dsg_soap_n_features(n_species, n_max, l_max);
    pl->n_lm = (l_max + 1) * (l_max + 1);
    // harmonics rows (harm_len doubles per neighbour); the stride is = 2 mod 4 so that the
    // phase-B stores of the lanes spread over the banks
    pl->kblocks = (n_max + kKBlock - 1) / kKBlock;
    int p = packed ? kGroups * kPackSlots : harm_len(lt);
    while (p % (lt == 8 ? 4 : 16) != 2) ++p;
    if (DSG_SOAP_MMA) p = mma_row_stride(lt);   // dense rows, odd stride (see soap_kernel)
    pl->row_stride = p;
    int rb = kChunk * p + s0 + 2;  // tail: item rows are read S_t wide past the last row
    // fused epilogue: G and the mixed coefficients both live here (split mode: neither)
    const int need = soap_split(n_species, n_max) ? 0 : 2 * ((pl->n_lm * n_max + 1) & ~1);
    if (rb < need) rb = need;
    rb = (rb + 1) & ~1;
    pl->rbuf_doubles = rb;
    // longest-processing-time packing of l = l_max..1 (size 2l+1) into 4 lane groups
    int load[kGroups] = {0, 0, 0, 0}, cnt[kGroups] = {0, 0, 0, 0};
    for (int g = 0; g < kGroups; ++g)
        for (int t = 0; t < kMaxItems; ++t) {
            pl->item_l[g][t] = -1;
            pl->item_off[g][t] = 0;
        }
    // l = 0 (a single m, weight exp(-gamma_0k r^2) only) is accumulated cooperatively by all lanes
    for (int l = l_max; l >= 1; --l) {
        int best = 0;
        for (int g = 1; g < kGroups; ++g)
            if (load[g] < load[best]) best = g;
        pl->item_off[best][cnt[best]] = packed ? hoff<8, kRowPacked>(l) : harm_off(lt, l);
        pl->item_l[best][cnt[best]++] = l;
        load[best] += 2 * l + 1;
    }
}

template <typename T>
static int soap_impl(const T* d_pos, const int32_t* d_species, const int32_t* d_centres,
                     int n_frames, int n_atoms, int n_cent, int n_species, const double* h_box,
                     double r_cut, int n_max, int l_max, const double* h_alphas,
                     const double* h_betas, void* d_ws, size_t ws_bytes, double* d_out,
                     int64_t stride_i, int64_t stride_f, cudaStream_t st, int delay = 0,
                     double* d_ts = nullptr, int64_t ts_ld = 0, int64_t ts_col0 = 0) {
    SoapLayout L;
    std::vector<SoapFrame> frames;
    int rc = soap_layout(n_frames, n_atoms, n_cent, n_species, n_max, l_max, h_box, r_cut, &L,
                         &frames, delay);
    if (rc) return rc;
    const bool seq = delay > 0;
    if (seq && !d_ts) return set_error(DSG_ERR_INVALID_ARGUMENT, "d_timesoap is NULL");
    if (!d_ws || ws_bytes < L.bytes)
        return set_error(DSG_ERR_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed",
                         ws_bytes, L.bytes);
    if (!d_pos || !d_species || !d_centres || (!d_out && !seq) || !h_alphas || !h_betas)
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
    int* sid = (int*)(ws + L.off_sid);
    int* cmap = (int*)(ws + L.off_cmap);
    int* tile = (int*)(ws + L.off_tile);
    int* slot_of = seq ? (int*)(ws + L.off_slot_of) : nullptr;

    // ---- basis tables (host float64) ----
    const double eta = 0.5;
    const int n_lm = (l_max + 1) * (l_max + 1);
    std::vector<double> gamma((size_t)(l_max + 1) * n_max), bmat((size_t)(l_max + 1) * n_max * n_max),
        swlm(n_lm);
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
            swlm[l * l + mi] = std::sqrt(pi * std::sqrt(8.0 / (2 * l + 1)) * n2);
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
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_gamma, gamma.data(), 8 * gamma.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_bmat, bmat.data(), 8 * bmat.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_wlm, swlm.data(), 8 * swlm.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_lml, lml.data(), 4 * lml.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_decode, decode.data(), 4 * decode.size(),
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemcpyAsync(d_frames, frames.data(), sizeof(SoapFrame) * n_frames,
                                   cudaMemcpyHostToDevice, st));
    double exp_tab[kExpTab];
    for (int j = 0; j < kExpTab; ++j) {   // bit pattern of 2^(j/128) with j << 13 taken off the high word
        const double v = std::exp2((double)j / kExpTab);
        uint64_t bits;
        std::memcpy(&bits, &v, 8);
        bits -= (uint64_t)j << (32 + kExpHiShift);
        std::memcpy(&exp_tab[j], &bits, 8);
    }
    DSG_CUDA_CHECK(cudaMemcpyAsync(ws + L.off_exptab, exp_tab, sizeof(exp_tab),
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
    DSG_CUDA_CHECK(cudaMemsetAsync(cmap, 0xff, 4 * (size_t)n_atoms, st));
    soap_centre_map_kernel<<<(n_cent + 255) / 256, 256, 0, st>>>(d_centres, n_cent, cmap);
    DSG_LAUNCH_CHECK("soap_centre_map_kernel");
    soap_bin_kernel<T><<<agrid, 256, 0, st>>>(d_pos, d_species, n_atoms, d_frames, wpos, cell_id,
                                              cnt);
    DSG_LAUNCH_CHECK("soap_bin_kernel");
    rc = run_scan(cnt, L.total_cells, 1, start, tile, nullptr, st);
    if (rc) return rc;
    soap_scatter_kernel<<<agrid, 256, 0, st>>>(wpos, n_atoms, d_frames, cell_id, cnt, start, sx,
                                               sy, sz, sid, slot_of);
    DSG_LAUNCH_CHECK("soap_scatter_kernel");

    // ---- main kernel ----
    SoapArgs a;
    a.sx = sx;
    a.sy = sy;
    a.sz = sz;
    a.sid = sid;
    a.cell_start = start;
    a.frames = d_frames;
    a.centre_of_atom = cmap;
    a.n_atoms = n_atoms;
    a.n_cent = n_cent;
    a.r2max = radius * radius;
    a.gamma = (const double*)(ws + L.off_gamma);
    a.bmat = (const double*)(ws + L.off_bmat);
    a.exp_tab = (const double*)(ws + L.off_exptab);
    {
        const size_t bd = (size_t)(l_max + 1) * n_max * n_max;
        a.bmat_smem_doubles = bd <= 1024 ? (int)((bd + 1) & ~(size_t)1) : 0;
    }
    a.swlm = (const double*)(ws + L.off_wlm);
    a.lm_l = (const int*)(ws + L.off_lml);
    a.decode = (const int*)(ws + L.off_decode);
    a.out = d_out;
    a.stride_i = stride_i;
    a.stride_f = stride_f;
    a.slot_of_atom = slot_of;
    a.centres = d_centres;
    a.n_frames = n_frames;
    a.delay = delay;
    a.seq_run = L.seq_run;
    a.ring = seq ? (double*)(ws + L.off_ring) : nullptr;
    a.ts_out = d_ts;
    a.ts_ld = ts_ld;
    a.ts_col0 = ts_col0;
    a.rows = nullptr;
    a.rowcnt = nullptr;
    a.row_cap = L.row_cap;
    a.list_flag = nullptr;
    a.list_role = 0;
    if (L.row_cap > 0) {
        a.rows = (int*)(ws + L.off_rows);
        a.rowcnt = (int*)(ws + L.off_rowcnt);
        a.list_flag = (int*)(ws + L.off_listflag);
        a.list_role = 1;
    }

    // template instantiation by l_max: compile-time l cap + slot widths of the packed items
    int lt, s0;
    if (l_max <= 4) { lt = 4; s0 = 9; }
    else if (l_max <= 6) { lt = 6; s0 = 13; }
    else if (l_max < 8 || (l_max == 8 && !DSG_SOAP_PACK)) { lt = 8; s0 = 17; }
    else if (l_max == 8) { lt = 8; s0 = kPackSlots; }  // packed rows: 20 slots per lane
    else if (l_max <= 10) { lt = 10; s0 = 21; }
    else { lt = 12; s0 = 25; }
    make_plan(n_species, n_max, l_max, lt, s0, &a.plan);
    if (a.list_role == 1) {   // neighbour rows for the list-mode kernels (needs the plan)
        DSG_CUDA_CHECK(cudaMemsetAsync(a.list_flag, 0, 4, st));
        dim3 lgrid((n_atoms + 127) / 128, n_frames);
        soap_list_kernel<<<lgrid, 128, 0, st>>>(a);
        DSG_LAUNCH_CHECK("soap_list_kernel");
    }
    const bool split = soap_split(n_species, n_max);
    a.gscr = split ? (double*)(ws + L.off_gscr) : nullptr;
    const size_t per_warp =
        8 * ((((size_t)kListCap * 4 + (size_t)kCellTab * 4 + (size_t)kBatchDoubles +
               (size_t)a.plan.rbuf_doubles) + 1) & ~(size_t)1);
    // two warps per block keep several blocks resident per SM
    // swlm (n_lm doubles) + decode (n_feat u16) + lm_l (n_lm bytes), rounded to 16 bytes
    a.epi_smem_doubles = split ? 0 : (int)((n_lm + (2 * n_feat + n_lm + 7) / 8 + 1) & ~1);
    if (DSG_SOAP_MMA && !split) {   // padded 8 x 8 matrices, squared weights per 8-column tile
        a.bmat_smem_doubles = (lt + 1) * 64;
        a.epi_smem_doubles = 8 * (mma_tile_base(lt + 1) + 1 + lt / 4);
    }
    const size_t shared_tables =
        8 * ((size_t)kExpTabDoubles +
             (split ? 0 : (size_t)a.bmat_smem_doubles + (size_t)a.epi_smem_doubles));
    int warps = DSG_SOAP_MMA ? kListWarps : 2;
    while (warps > 1 && per_warp * warps + shared_tables > 227 * 1024) warps >>= 1;
    const size_t smem = per_warp * warps + shared_tables + DSG_SOAP_EXTRA_SMEM;  // (occupancy probe)
    // list-mode consumer: smaller per-warp part, four warps share the block tables
    const size_t per_warp_l =
        8 * ((((size_t)soap_warp_fixed(true) + (size_t)a.plan.rbuf_doubles) + 1) & ~(size_t)1);
    int warps_l = (DSG_SOAP_MMA || s0 <= kPackSlots) ? kListWarps : 2;
    while (warps_l > 1 && per_warp_l * warps_l + shared_tables > 227 * 1024) warps_l >>= 1;
    const size_t smem_l = per_warp_l * warps_l + shared_tables + DSG_SOAP_EXTRA_SMEM;
    // spectrum kernel: mixed coefficients of one l per warp
    const size_t n_l_feat = (size_t)n_feat / (l_max + 1);   // outputs of one l, staged per warp
    const size_t spec_bytes =
        8 * ((((size_t)n_species * (2 * l_max + 1) * ((n_max + 1) & ~1)) + n_l_feat + 1) &
             ~(size_t)1);
    const size_t spec_block_bytes = 8 * n_l_feat;           // feature index + stride tables
    if (smem > 227 * 1024 ||
        (split && spec_bytes + spec_block_bytes + 8 * (size_t)a.bmat_smem_doubles > 227 * 1024))
        return set_error(DSG_ERR_UNSUPPORTED, "%zu bytes of shared memory per warp needed",
                         split ? spec_bytes : per_warp);
    // about 16 resident-block generations over the 148 SMs; runs of >= 1 slot per warp
    const int target_warps = 148 * 12 * 16;
    int spw = (int)(((long long)n_atoms * n_frames + target_warps - 1) / target_warps);
    if (spw < 1) spw = 1;
    if (spw > 64) spw = 64;
    a.slots_per_warp = spw;
    const int warps_needed = (n_atoms + spw - 1) / spw;
    dim3 grid((warps_needed + warps - 1) / warps, n_frames);
    dim3 grid_l((warps_needed + warps_l - 1) / warps_l, n_frames);
    // frame-sequential grids (fused timeSOAP): one warp per (centre, run of frame pairs)
    dim3 grid_q((n_atoms + warps - 1) / warps, seq ? L.n_runs : 1);
    dim3 grid_ql((n_atoms + warps_l - 1) / warps_l, seq ? L.n_runs : 1);
#define DSG_SOAP_LAUNCH_ONE(KERNEL, GRID, WARPS, SMEM)                                           \
    do {                                                                                         \
        DSG_CUDA_CHECK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            (int)(SMEM)));                                       \
        KERNEL<<<GRID, (WARPS) * 32, SMEM, st>>>(a);                                             \
    } while (0)
#define DSG_SOAP_LAUNCH(LT, A, B, C, D)                                                          \
    do {                                                                                         \
        if (split) {                                                                             \
            if (a.list_role == 1) {                                                              \
                DSG_SOAP_LAUNCH_ONE((soap_kernel<LT, A, B, C, D, true, true>), grid_l, warps_l,  \
                                    smem_l);                                                     \
                a.list_role = 2;                                                                 \
            }                                                                                    \
            DSG_SOAP_LAUNCH_ONE((soap_kernel<LT, A, B, C, D, true, false>), grid, warps, smem);  \
        } else if (seq) {                                                                        \
            if (a.list_role == 1) {                                                              \
                DSG_SOAP_LAUNCH_ONE((soap_kernel<LT, A, B, C, D, false, true, true>), grid_ql,   \
                                    warps_l, smem_l);                                            \
                a.list_role = 2;                                                                 \
            }                                                                                    \
            DSG_SOAP_LAUNCH_ONE((soap_kernel<LT, A, B, C, D, false, false, true>), grid_q,       \
                                warps, smem);                                                    \
        } else {                                                                                 \
            if (a.list_role == 1) {                                                              \
                DSG_SOAP_LAUNCH_ONE((soap_kernel<LT, A, B, C, D, false, true>), grid_l, warps_l, \
                                    smem_l);                                                     \
                a.list_role = 2;   /* the same centres by the in-kernel search iff a row overflowed */ \
            }                                                                                    \
            DSG_SOAP_LAUNCH_ONE((soap_kernel<LT, A, B, C, D, false, false>), grid, warps, smem); \
        }                                                                                        \
    } while (0)
#ifdef DSG_SOAP_ONLY_L8   // development builds: only the l_max = 8 instantiations (compile time)
    if (l_max != 8) return set_error(DSG_ERR_UNSUPPORTED, "development build: l_max = 8 only");
#else
    if (l_max <= 4) DSG_SOAP_LAUNCH(4, 9, 0, 0, 0);
    else if (l_max <= 6) DSG_SOAP_LAUNCH(6, 13, 5, 0, 0);
    else if (l_max < 8 || (l_max == 8 && !DSG_SOAP_PACK)) DSG_SOAP_LAUNCH(8, 17, 9, 0, 0);
    else
#endif
    if (l_max == 8) {
        // the packed kernel hard-codes the packing {8 - g, 1 + g} of lane group g
        for (int g = 0; g < kGroups && !DSG_SOAP_MMA; ++g)
            if (a.plan.item_l[g][0] != 8 - g || a.plan.item_l[g][1] != 1 + g)
                return set_error(DSG_ERR_UNSUPPORTED, "unexpected l packing for l_max = 8");
        DSG_SOAP_LAUNCH(8, kPackSlots, 0, 0, 0);
    }
#ifndef DSG_SOAP_ONLY_L8
    else if (l_max <= 10) DSG_SOAP_LAUNCH(10, 21, 13, 5, 0);
    else DSG_SOAP_LAUNCH(12, 25, 17, 9, 0);
#endif
#undef DSG_SOAP_LAUNCH
#undef DSG_SOAP_LAUNCH_ONE
    DSG_LAUNCH_CHECK("soap_kernel");
    if (soap_spectrum_mma_applies(n_species, n_max)) {
        static_assert(kSpecWarps == kSpecWarpsHost, "spectrum kernel block width");
        bool ring_smem = false;
        const size_t ssm = soap_spectrum_mma_smem(n_species, n_max, l_max, seq, delay, &ring_smem);
        dim3 sgrid(n_cent, seq ? L.n_runs : n_frames);
#define DSG_SPEC_LAUNCH_ONE(S, SEQ, RS)                                                           \
    do {                                                                                          \
        DSG_CUDA_CHECK(cudaFuncSetAttribute((soap_spectrum_mma_kernel<S, SEQ, RS>),               \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                            (int)ssm));                                           \
        soap_spectrum_mma_kernel<S, SEQ, RS><<<sgrid, 32 * kSpecWarps, ssm, st>>>(a);             \
    } while (0)
#define DSG_SPEC_LAUNCH(S)                                                                        \
    do {                                                                                          \
        if (ring_smem) DSG_SPEC_LAUNCH_ONE(S, true, true);                                        \
        else if (seq) DSG_SPEC_LAUNCH_ONE(S, true, false);                                        \
        else DSG_SPEC_LAUNCH_ONE(S, false, false);                                                \
    } while (0)
        if (n_species == 2) DSG_SPEC_LAUNCH(2);
        else if (n_species == 3) DSG_SPEC_LAUNCH(3);
        else DSG_SPEC_LAUNCH(4);
#undef DSG_SPEC_LAUNCH
#undef DSG_SPEC_LAUNCH_ONE
        DSG_LAUNCH_CHECK("soap_spectrum_mma_kernel");
    } else if (split) {
        std::vector<SpecTask> tasks;
        int base = 0;
        const int tiles = (n_max + 1) / 2;
        for (int zi = 0; zi < n_species; ++zi)
            for (int zj = zi; zj < n_species; ++zj) {
                const int blk = zi == zj ? n_max * (n_max + 1) / 2 : n_max * n_max;
                for (int ta = 0; ta < tiles; ++ta)
                    for (int tb = (zi == zj ? ta : 0); tb < tiles; ++tb)
                        tasks.push_back({zi | (zj << 4) | ((2 * ta) << 8) | ((2 * tb) << 16), base,
                                         blk, 0});
                base += blk * (l_max + 1);
            }
        if ((int)tasks.size() != L.n_tasks || base != n_feat)
            return set_error(DSG_ERR_INVALID_ARGUMENT, "spectrum task table mismatch");
        SpecTask* d_tasks = (SpecTask*)(ws + L.off_tasks);
        DSG_CUDA_CHECK(cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(SpecTask) * tasks.size(),
                                       cudaMemcpyHostToDevice, st));
        int sw = 4;  // warps per block, fewer if the coefficients of one l are large
        while (sw > 1 &&
               sw * spec_bytes + spec_block_bytes + 8 * (size_t)a.bmat_smem_doubles > 227 * 1024)
            --sw;
        const size_t ssm = sw * spec_bytes + spec_block_bytes + 8 * (size_t)a.bmat_smem_doubles;
        if (seq) {
            DSG_CUDA_CHECK(cudaFuncSetAttribute(soap_spectrum_kernel<true>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)ssm));
            dim3 sgrid((n_cent + sw - 1) / sw, L.n_runs);
            soap_spectrum_kernel<true><<<sgrid, sw * 32, ssm, st>>>(a, d_tasks, L.n_tasks);
        } else {
            DSG_CUDA_CHECK(cudaFuncSetAttribute(soap_spectrum_kernel<false>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)ssm));
            dim3 sgrid((n_cent + sw - 1) / sw, n_frames);
            soap_spectrum_kernel<false><<<sgrid, sw * 32, ssm, st>>>(a, d_tasks, L.n_tasks);
        }
        DSG_LAUNCH_CHECK("soap_spectrum_kernel");
    }
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

extern "C" int dsg_soap_timesoap_workspace_bytes(int n_frames, int n_atoms, int n_cent,
                                                 int n_species, int n_max, int l_max,
                                                 const double* h_box, double r_cut, int delay,
                                                 size_t* bytes_out) {
    if (!bytes_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "bytes_out is NULL");
    if (delay < 1) return set_error(DSG_ERR_INVALID_ARGUMENT, "delay=%d must be >= 1", delay);
    SoapLayout L;
    int rc = soap_layout(n_frames, n_atoms, n_cent, n_species, n_max, l_max, h_box, r_cut, &L,
                         nullptr, delay);
    if (rc) return rc;
    *bytes_out = L.bytes;
    return DSG_OK;
}

extern "C" int dsg_soap_timesoap(const void* d_pos, int pos_dtype, const int32_t* d_species,
                                 const int32_t* d_centres, int n_frames, int n_atoms, int n_cent,
                                 int n_species, const double* h_box, double r_cut, int n_max,
                                 int l_max, const double* h_alphas, const double* h_betas,
                                 int delay, void* d_workspace, size_t workspace_bytes,
                                 double* d_timesoap, int64_t ts_ld, int64_t ts_col0,
                                 double* d_soap_out, int64_t stride_i, int64_t stride_f,
                                 void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (delay < 1) return set_error(DSG_ERR_INVALID_ARGUMENT, "delay=%d must be >= 1", delay);
    if (pos_dtype == DSG_F32)
        return soap_impl<float>((const float*)d_pos, d_species, d_centres, n_frames, n_atoms,
                                n_cent, n_species, h_box, r_cut, n_max, l_max, h_alphas, h_betas,
                                d_workspace, workspace_bytes, d_soap_out, stride_i, stride_f, st,
                                delay, d_timesoap, ts_ld, ts_col0);
    if (pos_dtype == DSG_F64)
        return soap_impl<double>((const double*)d_pos, d_species, d_centres, n_frames, n_atoms,
                                 n_cent, n_species, h_box, r_cut, n_max, l_max, h_alphas, h_betas,
                                 d_workspace, workspace_bytes, d_soap_out, stride_i, stride_f, st,
                                 delay, d_timesoap, ts_ld, ts_col0);
    return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                     pos_dtype);
}
