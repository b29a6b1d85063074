// neighbors.cu -- batched periodic cell list -> sorted CSR neighbour lists (sm_100a).
//
// Replaces lens/lens.py:29-147 (build_cell_list + neighbor_list_celllist_centers) of the
// reference.  The reference builds a linked-list cell list serially and walks the 27 cells twice
// per centre on CPU threads; here every frame of a batch is processed by the same launches:
//
//   K1 cell_assign   thread / atom : reference cell index (same formula, float64), histogram
//   K2 scan          exclusive scan of the cell histogram of the whole batch -> cell_start
//   K3 cell_scatter  thread / atom : counting-sort scatter into cell order as float4 {x,y,z,id}
//   K4 neigh<COUNT>  warp / centre cell : the 27 neighbour cells are staged through shared
//                    memory once per cell, every centre of the cell tests the staged tile,
//                    warp ballots give per-centre counts
//   K5 scan          counts -> per-frame CSR row pointers (+ int64 frame offsets)
//   K6 neigh<FILL>   same traversal, ballot compaction writes the rows, warp bitonic sort
//
// Bit-exactness: the distance test is decided in float32 only when the float32 value is outside
// a rigorous guard band around (r_cut-1e-6)^2 (frame_finalize computes the band from the
// coordinate magnitudes); inside the band the reference's float64 expression is evaluated.
#include <climits>
#include <cmath>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace dsg {

struct FrameParams {
    double box[3];
    int ncell[3];
    int cell_base;           // first cell of this frame in the batch-wide cell arrays
    unsigned amax_bits[3];   // max |x_k| over env atoms and centres (float bits, atomicMax)
    float r2_lo, r2_hi;      // float32 guard band
    float boxf[3];
    float inv_boxf[3];
    int shift_mode[3];       // rows path: image chosen by the neighbour-cell wrap, not by rint
    float r2_lo_img, r2_hi_img;  // guard band of the plain float32 minimum-image test (pair path)
    int regular;             // every atom sits in [0, box): its cell needed no modulo (pair path)
};

constexpr int kWarpsPerBlock = 8;
constexpr int kStageCap = 128;  // float4 candidates staged per warp (2 KB)
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// scan primitives (int32 items, uniform segments)
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(kFull, v, off);
        if (lane >= off) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across a 256-thread block; returns exclusive prefix,
// *total receives the block total.
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
    __shared__ int warp_sums[kScanThreads / 32];
    __shared__ int block_total;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;
        if (lane == kScanThreads / 32 - 1) block_total = wi;
    }
    __syncthreads();
    int res = incl - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const int* __restrict__ in, long long seg_len, int tiles_per_seg,
                      int* __restrict__ tile_sums) {
    const long long seg = blockIdx.y;
    const long long tile0 = (long long)blockIdx.x * kScanTile;
    const int* src = in + seg * seg_len;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long idx = tile0 + threadIdx.x + (long long)k * kScanThreads;
        if (idx < seg_len) s += src[idx];
    }
    int total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[seg * tiles_per_seg + blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_offsets_kernel(int* __restrict__ tile_sums, int tiles_per_seg,
                         long long* __restrict__ seg_totals) {
    const long long seg = blockIdx.x;
    int* t = tile_sums + seg * tiles_per_seg;
    long long running = 0;
    for (int base = 0; base < tiles_per_seg; base += kScanThreads) {
        int i = base + threadIdx.x;
        int v = i < tiles_per_seg ? t[i] : 0;
        int total;
        int ex = block_excl_scan(v, &total);
        if (i < tiles_per_seg) t[i] = (int)(running + ex);
        running += total;
    }
    if (threadIdx.x == 0 && seg_totals) seg_totals[seg] = running;
}

// out has (seg_len + 1) entries per segment: exclusive prefix + total.
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int* __restrict__ in, long long seg_len, int tiles_per_seg,
                  const int* __restrict__ tile_offsets, int* __restrict__ out) {
    const long long seg = blockIdx.y;
    const long long first = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    const int* src = in + seg * seg_len;
    int* dst = out + seg * (seg_len + 1);
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long idx = first + k;
        v[k] = idx < seg_len ? src[idx] : 0;
        s += v[k];
    }
    int total;
    int ex = block_excl_scan(s, &total) + tile_offsets[seg * tiles_per_seg + blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long idx = first + k;
        if (idx < seg_len) dst[idx] = ex;
        ex += v[k];
        if (idx == seg_len - 1) dst[seg_len] = ex;
    }
}

// frame_off[0..n] = exclusive int64 scan of seg_totals[0..n)
__global__ void frame_offsets_kernel(const long long* __restrict__ seg_totals, int n,
                                     long long* __restrict__ frame_off) {
    // one warp; chunks of 32
    int lane = threadIdx.x;
    long long running = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        long long v = i < n ? seg_totals[i] : 0;
        long long incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            long long t = __shfl_up_sync(kFull, incl, off);
            if (lane >= off) incl += t;
        }
        if (i < n) frame_off[i] = running + incl - v;
        running += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) frame_off[n] = running;
}

int run_scan(const int* in, long long seg_len, int n_seg, int* out, int* tile_tmp,
             long long* seg_totals, cudaStream_t st) {
    int tiles = (int)((seg_len + kScanTile - 1) / kScanTile);
    dim3 grid(tiles, n_seg);
    scan_tile_sums_kernel<<<grid, kScanThreads, 0, st>>>(in, seg_len, tiles, tile_tmp);
    DSG_LAUNCH_CHECK("scan_tile_sums_kernel");
    scan_tile_offsets_kernel<<<n_seg, kScanThreads, 0, st>>>(tile_tmp, tiles, seg_totals);
    DSG_LAUNCH_CHECK("scan_tile_offsets_kernel");
    scan_apply_kernel<<<grid, kScanThreads, 0, st>>>(in, seg_len, tiles, tile_tmp, out);
    DSG_LAUNCH_CHECK("scan_apply_kernel");
    return DSG_OK;
}

// ------------------------------------------------------------------------------------------
// K1 / K3 : cell assignment and counting-sort scatter
// ------------------------------------------------------------------------------------------
// the truncated quotient of cell_of already lies in [0, ncell): no modulo was needed
__device__ __forceinline__ bool in_grid(double x, double box, int ncell) {
    return (long long)(x / box * (double)ncell) < (long long)ncell;
}

template <typename T>
__global__ void __launch_bounds__(256)
cell_assign_kernel(const T* __restrict__ pos, int n_atoms, FrameParams* __restrict__ fp,
                   unsigned* __restrict__ cell_id, int* __restrict__ cell_cnt,
                   unsigned long long* __restrict__ rank_out = nullptr) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const FrameParams* p = fp + f;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (i < n_atoms) {
        const size_t g = ((size_t)f * n_atoms + i) * 3;
        const double x = (double)pos[g], y = (double)pos[g + 1], z = (double)pos[g + 2];
        const int ny = p->ncell[1], nz = p->ncell[2];
        const int cx = cell_of(x, p->box[0], p->ncell[0]);
        const int cy = cell_of(y, p->box[1], ny);
        const int cz = cell_of(z, p->box[2], nz);
        const int c = (cx * ny + cy) * nz + cz;
        // an atom outside [0, box) is binned by truncation + modulo (lens.py:52-54), which is not
        // its geometric cell: the pair path then checks cell adjacency explicitly in this frame
        if (!(x >= 0.0 && y >= 0.0 && z >= 0.0 && in_grid(x, p->box[0], p->ncell[0]) &&
              in_grid(y, p->box[1], ny) && in_grid(z, p->box[2], nz)))
            atomicAnd(&fp[f].regular, 0);
        cell_id[(size_t)f * n_atoms + i] = (unsigned)c;
        // rank_out (pair path): the atom's rank inside its cell, so that the scatter needs no second
        // atomic per atom (it was bound by their latency); else a plain RED, no value comes back
        if (rank_out)
            rank_out[(size_t)f * n_atoms + i] =
                (unsigned long long)atomicAdd(&cell_cnt[p->cell_base + c], 1);
        else
            atomicAdd(&cell_cnt[p->cell_base + c], 1);
        ax = __double2float_ru(fabs(x));
        ay = __double2float_ru(fabs(y));
        az = __double2float_ru(fabs(z));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ax = fmaxf(ax, __shfl_xor_sync(kFull, ax, off));
        ay = fmaxf(ay, __shfl_xor_sync(kFull, ay, off));
        az = fmaxf(az, __shfl_xor_sync(kFull, az, off));
    }
    // one RED.MAX per block and axis (non-negative floats order like their bit patterns); no value
    // is read back, so no thread waits on the atomics
    __shared__ float red[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = ax;
        red[1][warp] = ay;
        red[2][warp] = az;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float m = red[threadIdx.x][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[threadIdx.x][w]);
        atomicMax(&fp[f].amax_bits[threadIdx.x], __float_as_uint(m));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
cell_scatter_kernel(const T* __restrict__ pos, int n_atoms, const FrameParams* __restrict__ fp,
                    const unsigned* __restrict__ cell_id, int* __restrict__ cell_cnt,
                    const int* __restrict__ cell_start, float4* __restrict__ sorted) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const size_t a = (size_t)f * n_atoms + i;
    const int gc = fp[f].cell_base + (int)cell_id[a];
    // cell_cnt is consumed back to zero: slot = start + (remaining - 1)
    const int slot = cell_start[gc] + atomicSub(&cell_cnt[gc], 1) - 1;
    const size_t g = a * 3;
    sorted[slot] = make_float4((float)pos[g], (float)pos[g + 1], (float)pos[g + 2],
                               __int_as_float(i));
}

// Guard band of the float32 distance test (see DESIGN.md "bit-exact membership"):
// |dr2_f32 - dr2_ref| <= tol whenever every |d_k| <= 1.5 r_cut, so
//   dr2_f32 < r2 - tol  => reference accepts,   dr2_f32 > r2 + tol => reference rejects.
__global__ void frame_finalize_kernel(FrameParams* __restrict__ fp, int n_frames, double r_cut,
                                      double r2) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    FrameParams* p = fp + f;
    const double u = 5.9604644775390625e-08;  // 2^-24
    double tol = 8.0 * u * r2, tol_img = 8.0 * u * r2;
    bool exact_only = false;
    for (int k = 0; k < 3; ++k) {
        const double a = (double)__uint_as_float(p->amax_bits[k]);
        const double b = p->box[k] > 0.0 ? p->box[k] : 0.0;
        double delta = 32.0 * u * (a + b);
        {
            // pair path, partner frame: d = fl(x_j - x_i) from the caller's coordinates (rounded
            // to float32 first if they are float64: 2 u a), error u 2a; the image n = rintf(d /
            // b) is removed by ONE fma(-n, b_f32, d): the product is exact, the sum rounds once
            // (u |d'| <= u b / 4 for the pairs that matter) and b_f32 is off by u b, times
            // |n| <= 2a/b + 1.  Total u (6a + 2.25 b) against the exact difference; the
            // reference's own float64 evaluation is within 2^-48 (a + b) of that.
            const double d_img = u * (6.0 * a + 2.25 * b) + 3.5527136788005009e-15 * (a + b);
            tol_img += d_img * (3.0 * r_cut + d_img);
        }
        if (p->shift_mode[k]) {
            // rows path, shift mode: the staged value is the image next to the atom's cell,
            // |x'| <= A = b + edge <= 1.2 b (ncell >= 6), computed in float64 (error
            // <= 2^-52 (a + b)) and rounded once (u A).  The difference is formed as
            // (x'_j + (+-b_f32 - x'_i)) or ((x'_j +- b_f32) - x'_i): roundings u A (x'_j), u A
            // (x'_i), u b (b_f32), u (b + A) (inner sum), u * 0.75 b (outer sum, |d| <= 4.5
            // edges) -- 6.35 u b in total; 8 u b is used.
            delta = 8.0 * u * b + 8.8817841970012523e-16 * (a + b);
        }
        tol += delta * (3.0 * r_cut + delta);
        if (b > 0.0) {
            const double dq = 4.0 * u * (2.0 * a / b + 1.0);  // error of the f32 image quotient
            if (!(dq < 0.25)) exact_only = true;
        }
        p->boxf[k] = (float)p->box[k];
        p->inv_boxf[k] = p->box[k] > 0.0 ? (float)(1.0 / p->box[k]) : 0.f;
    }
    if (!(tol < 0.5 * r2) || exact_only) {
        p->r2_lo = -INFINITY;
        p->r2_hi = INFINITY;
    } else {
        p->r2_lo = __double2float_rd(r2 - tol);
        p->r2_hi = __double2float_ru(r2 + tol);
    }
    if (!(tol_img < 0.5 * r2) || exact_only) {
        p->r2_lo_img = -INFINITY;
        p->r2_hi_img = INFINITY;
    } else {
        p->r2_lo_img = __double2float_rd(r2 - tol_img);
        p->r2_hi_img = __double2float_ru(r2 + tol_img);
    }
}

// ------------------------------------------------------------------------------------------
// K4 / K6 : neighbour traversal
// ------------------------------------------------------------------------------------------
struct NeighArgs {
    const void* pos_env;    // original positions (exact test)
    const void* pos_cent;
    int n_env, n_cent;
    const FrameParams* fp;
    const int* env_start;   // batch-wide cell starts of env atoms
    const int* cent_start;  // same grid, centres
    const float4* env_sorted;
    const float4* cent_sorted;
    int respect_pbc;
    double r2;
    int* counts;                  // COUNT: (F, M)
    const int* indptr;            // FILL: (F, M+1)
    const long long* frame_off;   // FILL: (F+1)
    int* indices;                 // FILL
};

template <typename T>
__device__ __noinline__ bool exact_test(const T* __restrict__ pe, const T* __restrict__ pc,
                                        const double* __restrict__ box, bool pbc, double r2) {
    // lens.py:100-104 in float64, from the caller's original coordinates
    double dr2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double d = (double)pe[k] - (double)pc[k];
        if (pbc && box[k] > 0.0) d -= rint(d / box[k]) * box[k];
        dr2 += d * d;
    }
    return dr2 < r2;
}

__device__ __forceinline__ void warp_sort_row(int* row, int n, int* smem_i, int lane) {
    if (n <= 1) return;
    if (n <= 32) {
        int v = lane < n ? row[lane] : INT_MAX;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int o = __shfl_xor_sync(kFull, v, j);
                const bool asc = (lane & k) == 0;
                const bool lower = (lane & j) == 0;
                v = (lower == asc) ? min(v, o) : max(v, o);
            }
        }
        if (lane < n) row[lane] = v;
        return;
    }
    if (n <= kStageCap * 4) {
        int p2 = 64;
        while (p2 < n) p2 <<= 1;
        for (int t = lane; t < p2; t += 32) smem_i[t] = t < n ? row[t] : INT_MAX;
        __syncwarp();
        for (int k = 2; k <= p2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < (p2 >> 1); t += 32) {
                    const int i0 = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int i1 = i0 | j;
                    const bool asc = (i0 & k) == 0;
                    const int x = smem_i[i0], y = smem_i[i1];
                    if ((x > y) == asc) {
                        smem_i[i0] = y;
                        smem_i[i1] = x;
                    }
                }
                __syncwarp();
            }
        }
        for (int t = lane; t < n; t += 32) row[t] = smem_i[t];
        __syncwarp();
        return;
    }
    // very long rows (rare): in-place odd-even transposition in global memory
    for (int phase = 0; phase < n; ++phase) {
        for (int t = (phase & 1) + 2 * lane; t + 1 < n; t += 64) {
            const int x = row[t], y = row[t + 1];
            if (x > y) {
                row[t] = y;
                row[t + 1] = x;
            }
        }
        __syncwarp();
    }
}

template <typename T, bool FILL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
neigh_kernel(const NeighArgs a) {
    __shared__ float4 stage_all[kWarpsPerBlock][kStageCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const FrameParams* __restrict__ p = a.fp + f;
    const int nx = p->ncell[0], ny = p->ncell[1], nz = p->ncell[2];
    const int c = blockIdx.x * kWarpsPerBlock + warp;  // cell inside the frame
    if (c >= nx * ny * nz) return;
    const int base_cell = p->cell_base;
    const int cbeg = a.cent_start[base_cell + c];
    const int cend = a.cent_start[base_cell + c + 1];
    if (cbeg == cend) return;

    float4* stage = stage_all[warp];
    const int cz = c % nz, cy = (c / nz) % ny, cx = c / (nz * ny);

    // the 27 cells of the reference's traversal (lens.py:92-97); lanes 27..31 idle
    int run_start = 0, run_len = 0;
    if (lane < 27) {
        const int ddx = lane / 9 - 1, ddy = (lane / 3) % 3 - 1, ddz = lane % 3 - 1;
        const int cc = (wrap_cell(cx + ddx, nx) * ny + wrap_cell(cy + ddy, ny)) * nz +
                       wrap_cell(cz + ddz, nz);
        run_start = a.env_start[base_cell + cc];
        run_len = a.env_start[base_cell + cc + 1] - run_start;
    }
    const int run_incl = warp_incl_scan(run_len, lane);
    const int run_excl = run_incl - run_len;
    const int n_cand = __shfl_sync(kFull, run_incl, 31);

    const bool pbc = a.respect_pbc != 0;
    const float bx = p->boxf[0], by = p->boxf[1], bz = p->boxf[2];
    const float ibx = p->inv_boxf[0], iby = p->inv_boxf[1], ibz = p->inv_boxf[2];
    const bool px = pbc && p->box[0] > 0.0, py = pbc && p->box[1] > 0.0, pz = pbc && p->box[2] > 0.0;
    const float r2_lo = p->r2_lo, r2_hi = p->r2_hi;
    const unsigned lt_mask = (1u << lane) - 1u;
    const T* __restrict__ pos_env_f = (const T*)a.pos_env + (size_t)f * a.n_env * 3;
    const T* __restrict__ pos_cent_f = (const T*)a.pos_cent + (size_t)f * a.n_cent * 3;
    const double r2 = a.r2;

    for (int cb = cbeg; cb < cend; cb += 32) {
        const int nb = min(32, cend - cb);
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < nb) me = a.cent_sorted[cb + lane];
        const int my_id = __float_as_int(me.w);
        int my_cnt = 0;
        long long my_row = 0;
        if (FILL && lane < nb)
            my_row = a.frame_off[f] + (long long)a.indptr[(size_t)f * (a.n_cent + 1) + my_id];

        for (int base = 0; base < n_cand; base += kStageCap) {
            const int clen = min(kStageCap, n_cand - base);
            __syncwarp();
            // stage candidates [base, base+clen): flat index -> (run, offset) by a shuffle
            // binary search over the inclusive run prefix
            for (int q0 = 0; q0 < clen; q0 += 32) {
                const int q = base + q0 + lane;
                const bool valid = q0 + lane < clen;
                int r = 0;
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    const int v = __shfl_sync(kFull, run_incl, (r + s - 1) & 31);
                    if (v <= q) r += s;
                }
                r = valid ? r : 0;
                const int src = __shfl_sync(kFull, run_start, r) + (q - __shfl_sync(kFull, run_excl, r));
                if (valid) stage[q0 + lane] = a.env_sorted[src];
            }
            __syncwarp();

            for (int ci = 0; ci < nb; ++ci) {
                const float ccx = __shfl_sync(kFull, me.x, ci);
                const float ccy = __shfl_sync(kFull, me.y, ci);
                const float ccz = __shfl_sync(kFull, me.z, ci);
                const int cid = __shfl_sync(kFull, my_id, ci);
                const long long row = FILL ? __shfl_sync(kFull, my_row, ci) : 0;
                const int cur = __shfl_sync(kFull, my_cnt, ci);
                int add = 0;
                for (int q0 = 0; q0 < clen; q0 += 32) {
                    const bool valid = q0 + lane < clen;
                    const float4 cd = stage[valid ? q0 + lane : 0];
                    const int j = __float_as_int(cd.w);
                    float dx = cd.x - ccx, dy = cd.y - ccy, dz = cd.z - ccz;
                    if (px) dx = fmaf(-rintf(dx * ibx), bx, dx);
                    if (py) dy = fmaf(-rintf(dy * iby), by, dy);
                    if (pz) dz = fmaf(-rintf(dz * ibz), bz, dz);
                    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    bool hit = d2 < r2_lo;
                    if (valid && !hit && !(d2 > r2_hi))
                        hit = exact_test<T>(pos_env_f + (size_t)j * 3, pos_cent_f + (size_t)cid * 3,
                                            p->box, pbc, r2);
                    hit = hit && valid && (j != cid);  // lens.py:104 "j != i"
                    const unsigned m = __ballot_sync(kFull, hit);
                    if (FILL && hit) a.indices[row + cur + add + __popc(m & lt_mask)] = j;
                    add += __popc(m);
                }
                if (lane == ci) my_cnt += add;
            }
        }
        if (!FILL) {
            if (lane < nb) a.counts[(size_t)f * a.n_cent + my_id] = my_cnt;
        } else {
            __syncwarp();
            for (int ci = 0; ci < nb; ++ci) {
                const long long row = __shfl_sync(kFull, my_row, ci);
                const int n = __shfl_sync(kFull, my_cnt, ci);
                warp_sort_row(a.indices + row, n, reinterpret_cast<int*>(stage), lane);
            }
        }
    }
}

// Canonical CSR from the single-traversal rows: one warp per (frame, centre) copies the row to
// its place in the frame's CSR (coalesced) and sorts it there (same sorter as the FILL kernel).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rows_to_csr_kernel(const int* __restrict__ counts, const long long* __restrict__ row_start,
                   const int* __restrict__ rows, const int* __restrict__ indptr,
                   const long long* __restrict__ frame_off, int n_frames, int n_cent,
                   int* __restrict__ indices) {
    __shared__ int sort_tile[kWarpsPerBlock][kStageCap * 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kWarpsPerBlock + warp;
    if (w >= (long long)n_frames * n_cent) return;
    const int f = (int)(w / n_cent), i = (int)(w - (long long)f * n_cent);
    const int n = counts[w];
    const long long src0 = row_start[w];
    if (n <= 0 || src0 < 0) return;
    const int* __restrict__ src = rows + src0;
    int* dst = indices + frame_off[f] + indptr[(size_t)f * (n_cent + 1) + i];
    for (int t = lane; t < n; t += 32) dst[t] = src[t];
    __syncwarp();
    warp_sort_row(dst, n, sort_tile[warp], lane);
}

// Short rows: ONE THREAD per (frame, centre).  The row sits in registers (up to 32 values, loaded in
// chunks of 8 up to the longest row of the warp), is sorted by a bitonic network with compile-time
// indices (8 / 16 / 32 wide, chosen per warp) and written to its CSR place.  A warp that meets a
// row longer than 32 handles its 32 rows with the warp-cooperative copy-and-sort.
template <int N>
__device__ __forceinline__ void sort_regs(int (&v)[32]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int p = i ^ j;
                if (p > i) {
                    const bool asc = (i & k) == 0;
                    const int lo = min(v[i], v[p]), hi = max(v[i], v[p]);
                    v[i] = asc ? lo : hi;
                    v[p] = asc ? hi : lo;
                }
            }
        }
    }
}

constexpr int kRtThreads = 128;
__global__ void __launch_bounds__(kRtThreads)
rows_to_csr_thread_kernel(const int* __restrict__ counts, const long long* __restrict__ row_start,
                          const int* __restrict__ rows, const int* __restrict__ indptr,
                          const long long* __restrict__ frame_off, int n_frames, int n_cent,
                          int* __restrict__ indices) {
    __shared__ int sort_tile[kRtThreads / 32][kStageCap * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long w = (long long)blockIdx.x * kRtThreads + threadIdx.x;
    const bool active = w < (long long)n_frames * n_cent;
    int n = 0;
    long long src0 = -1;
    int* dst = nullptr;
    if (active) {
        const int f = (int)(w / n_cent), i = (int)(w - (long long)f * n_cent);
        n = counts[w];
        src0 = row_start[w];
        if (src0 < 0) n = 0;
        dst = indices + frame_off[f] + indptr[(size_t)f * (n_cent + 1) + i];
    }
    const int wmax = __reduce_max_sync(kFull, n);
    if (wmax == 0) return;
    if (wmax <= 32) {
        int v[32];
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            const bool need = c8 * 8 < wmax;   // warp-uniform
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[c8 * 8 + j] = (need && c8 * 8 + j < n) ? rows[src0 + c8 * 8 + j] : INT_MAX;
        }
        if (wmax <= 8) sort_regs<8>(v);
        else if (wmax <= 16) sort_regs<16>(v);
        else sort_regs<32>(v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < n) dst[j] = v[j];
    } else {
        for (int src = 0; src < 32; ++src) {
            const int n_s = __shfl_sync(kFull, n, src);
            const long long s_s = __shfl_sync(kFull, src0, src);
            const long long d_s = __shfl_sync(kFull, (long long)(dst - indices), src);
            if (n_s <= 0) continue;
            int* d = indices + d_s;
            for (int t = lane; t < n_s; t += 32) d[t] = rows[s_s + t];
            __syncwarp();
            warp_sort_row(d, n_s, sort_tile[warp], lane);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct NeighLayout {
    long long total_cells = 0;
    int max_cells = 0;
    size_t off_fp = 0, off_cell_id_e = 0, off_cnt_e = 0, off_start_e = 0, off_sorted_e = 0;
    size_t off_cell_id_c = 0, off_cnt_c = 0, off_start_c = 0, off_sorted_c = 0;
    size_t off_tile_tmp = 0, off_seg_totals = 0;
    size_t off_slot_cell_e = 0, off_slot_cell_c = 0;  // packed (cx,cy,cz) of a sorted slot
    size_t bytes = 0;
    size_t tile_tmp_ints = 0;
};

static int make_layout(int n_frames, int n_env, int n_cent, bool same, const double* h_box,
                       double r_cut, NeighLayout* L, std::vector<FrameParams>* fps) {
    if (n_frames < 1 || n_frames > 65535)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d must be in [1, 65535] per batch",
                         n_frames);
    if (n_env < 1 || n_cent < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_env=%d and n_cent=%d must be >= 1", n_env,
                         n_cent);
    if (!(r_cut > 0.0)) return set_error(DSG_ERR_INVALID_ARGUMENT, "r_cut=%g must be > 0", r_cut);
    if (!h_box) return set_error(DSG_ERR_INVALID_ARGUMENT, "h_box is NULL");
    if ((long long)n_frames * n_env >= INT_MAX || (long long)n_frames * n_cent >= INT_MAX)
        return set_error(DSG_ERR_OVERFLOW, "n_frames*n_atoms must stay below 2^31 per batch");
    long long total = 0;
    int max_cells = 0;
    if (fps) fps->resize(n_frames);
    for (int f = 0; f < n_frames; ++f) {
        long long nc[3];
        for (int k = 0; k < 3; ++k) {
            const double b = h_box[3 * f + k];
            if (!(b > 0.0) || !std::isfinite(b))
                return set_error(DSG_ERR_INVALID_ARGUMENT,
                                 "box[%d][%d]=%g: the cell list needs a positive box", f, k, b);
            double q = py_floordiv(b, r_cut);  // lens.py:43-45
            long long n = (long long)q;
            if (n < 3) n = 3;
            nc[k] = n;
        }
        const long long cells = nc[0] * nc[1] * nc[2];
        if (cells >= INT_MAX || total + cells >= INT_MAX)
            return set_error(DSG_ERR_OVERFLOW,
                             "cell grid too large (frame %d: %lld cells); use fewer frames per batch",
                             f, cells);
        if (fps) {
            FrameParams& p = (*fps)[f];
            for (int k = 0; k < 3; ++k) {
                p.box[k] = h_box[3 * f + k];
                p.ncell[k] = (int)nc[k];
                p.amax_bits[k] = 0u;
                p.boxf[k] = 0.f;
                p.inv_boxf[k] = 0.f;
                p.shift_mode[k] = 0;
            }
            p.cell_base = (int)total;
            p.r2_lo = p.r2_hi = 0.f;
            p.r2_lo_img = p.r2_hi_img = 0.f;
            p.regular = 1;
        }
        total += cells;
        if (cells > max_cells) max_cells = (int)cells;
    }
    L->total_cells = total;
    L->max_cells = max_cells;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += align_up(bytes);
        return r;
    };
    const size_t fe = (size_t)n_frames * n_env, fc = (size_t)n_frames * n_cent;
    L->off_fp = take(sizeof(FrameParams) * n_frames);
    L->off_cell_id_e = take(4 * fe);
    L->off_cnt_e = take(4 * (size_t)total);
    L->off_start_e = take(4 * ((size_t)total + 1));
    L->off_sorted_e = take(16 * (fe + 4));   // + 4: the pair kernel reads up to 3 slots past a run
    if (!same) {
        L->off_cell_id_c = take(4 * fc);
        L->off_cnt_c = take(4 * (size_t)total);
        L->off_start_c = take(4 * ((size_t)total + 1));
        L->off_sorted_c = take(16 * fc);
    } else {
        L->off_cell_id_c = L->off_cell_id_e;
        L->off_cnt_c = L->off_cnt_e;
        L->off_start_c = L->off_start_e;
        L->off_sorted_c = L->off_sorted_e;
    }
    const size_t tiles_cells = ((size_t)total + kScanTile - 1) / kScanTile;
    const size_t tiles_cnt = (size_t)n_frames * (((size_t)n_cent + kScanTile - 1) / kScanTile);
    L->tile_tmp_ints = tiles_cells > tiles_cnt ? tiles_cells : tiles_cnt;
    L->off_tile_tmp = take(4 * L->tile_tmp_ints);
    L->off_seg_totals = take(8 * (size_t)n_frames);
    L->off_slot_cell_e = take(4 * fe);
    L->off_slot_cell_c = same ? L->off_slot_cell_e : take(4 * fc);
    L->bytes = o;
    return DSG_OK;
}

template <typename T>
static int build_cells(const T* d_pos, int n_frames, int n_atoms, FrameParams* d_fp,
                       unsigned* d_cell_id, int* d_cnt, int* d_start, float4* d_sorted,
                       long long total_cells, int* d_tile_tmp, cudaStream_t st, bool scatter_only) {
    dim3 grid((n_atoms + 255) / 256, n_frames);
    if (!scatter_only) {
        DSG_CUDA_CHECK(cudaMemsetAsync(d_cnt, 0, 4 * (size_t)total_cells, st));
        cell_assign_kernel<T><<<grid, 256, 0, st>>>(d_pos, n_atoms, d_fp, d_cell_id, d_cnt);
        DSG_LAUNCH_CHECK("cell_assign_kernel");
        int rc = run_scan(d_cnt, total_cells, 1, d_start, d_tile_tmp, nullptr, st);
        if (rc) return rc;
    }
    cell_scatter_kernel<T><<<grid, 256, 0, st>>>(d_pos, n_atoms, d_fp, d_cell_id, d_cnt, d_start,
                                                 d_sorted);
    DSG_LAUNCH_CHECK("cell_scatter_kernel");
    return DSG_OK;
}

template <typename T>
static int neigh_impl(bool fill, const void* d_pos_env, const void* d_pos_cent, int n_frames,
                      int n_env, int n_cent, const double* h_box, double r_cut, int respect_pbc,
                      void* d_ws, size_t ws_bytes, int32_t* d_counts, int32_t* d_indptr,
                      int64_t* d_frame_off, int32_t* d_indices, cudaStream_t st) {
    const bool same = d_pos_cent == nullptr;
    if (same && n_cent != n_env)
        return set_error(DSG_ERR_INVALID_ARGUMENT,
                         "d_pos_cent is NULL (centres == env) but n_cent=%d != n_env=%d", n_cent,
                         n_env);
    NeighLayout L;
    std::vector<FrameParams> fps;
    int rc = make_layout(n_frames, n_env, n_cent, same, h_box, r_cut, &L, fill ? nullptr : &fps);
    if (rc) return rc;
    if (!d_ws || ws_bytes < L.bytes)
        return set_error(DSG_ERR_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed",
                         ws_bytes, L.bytes);
    if (!d_pos_env || !d_indptr || !d_frame_off || (!fill && !d_counts) || (fill && !d_indices))
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    char* ws = (char*)d_ws;
    FrameParams* d_fp = (FrameParams*)(ws + L.off_fp);
    unsigned* cid_e = (unsigned*)(ws + L.off_cell_id_e);
    int* cnt_e = (int*)(ws + L.off_cnt_e);
    int* start_e = (int*)(ws + L.off_start_e);
    float4* sorted_e = (float4*)(ws + L.off_sorted_e);
    unsigned* cid_c = (unsigned*)(ws + L.off_cell_id_c);
    int* cnt_c = (int*)(ws + L.off_cnt_c);
    int* start_c = (int*)(ws + L.off_start_c);
    float4* sorted_c = (float4*)(ws + L.off_sorted_c);
    int* tile_tmp = (int*)(ws + L.off_tile_tmp);
    long long* seg_totals = (long long*)(ws + L.off_seg_totals);
    const double rc_eff = r_cut - 1e-6;      // lens.py:78
    const double r2 = rc_eff * rc_eff;

    if (!fill) {
        DSG_CUDA_CHECK(cudaMemcpyAsync(d_fp, fps.data(), sizeof(FrameParams) * n_frames,
                                       cudaMemcpyHostToDevice, st));
        rc = build_cells<T>((const T*)d_pos_env, n_frames, n_env, d_fp, cid_e, cnt_e, start_e,
                            sorted_e, L.total_cells, tile_tmp, st, false);
        if (rc) return rc;
        if (!same) {
            rc = build_cells<T>((const T*)d_pos_cent, n_frames, n_cent, d_fp, cid_c, cnt_c,
                                start_c, sorted_c, L.total_cells, tile_tmp, st, false);
            if (rc) return rc;
        }
        frame_finalize_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(d_fp, n_frames, r_cut, r2);
        DSG_LAUNCH_CHECK("frame_finalize_kernel");
    }

    NeighArgs a;
    a.pos_env = d_pos_env;
    a.pos_cent = same ? d_pos_env : d_pos_cent;
    a.n_env = n_env;
    a.n_cent = n_cent;
    a.fp = d_fp;
    a.env_start = start_e;
    a.cent_start = start_c;
    a.env_sorted = sorted_e;
    a.cent_sorted = sorted_c;
    a.respect_pbc = respect_pbc;
    a.r2 = r2;
    a.counts = d_counts;
    a.indptr = d_indptr;
    a.frame_off = (const long long*)d_frame_off;
    a.indices = d_indices;
    dim3 grid((L.max_cells + kWarpsPerBlock - 1) / kWarpsPerBlock, n_frames);
    if (!fill) {
        neigh_kernel<T, false><<<grid, kWarpsPerBlock * 32, 0, st>>>(a);
        DSG_LAUNCH_CHECK("neigh_kernel<count>");
        rc = run_scan(d_counts, n_cent, n_frames, d_indptr, tile_tmp, seg_totals, st);
        if (rc) return rc;
        frame_offsets_kernel<<<1, 32, 0, st>>>(seg_totals, n_frames, (long long*)d_frame_off);
        DSG_LAUNCH_CHECK("frame_offsets_kernel");
    } else {
        neigh_kernel<T, true><<<grid, kWarpsPerBlock * 32, 0, st>>>(a);
        DSG_LAUNCH_CHECK("neigh_kernel<fill>");
    }
    return DSG_OK;
}


// ------------------------------------------------------------------------------------------
// Rows path: one traversal, unsorted rows at atomically reserved offsets (LENS fast path)
// ------------------------------------------------------------------------------------------
// compute_lens never returns the lists, so for LENS the rows need neither the canonical CSR
// order nor sorting (the warp merge in lens.cu compares all pairs).  Each warp (= one centre cell)
//   1. stages the candidates of its 27 cells once, with the periodic shift of the neighbour cell
//      already applied (axes with >= 6 cells: the image is implied by the cell offset, so the inner
//      test needs no rint; see DESIGN.md "shift mode"),
//   2. tests every centre of the cell against the tile and keeps the hit ballots in registers,
//   3. reserves space for all rows of the cell with ONE atomicAdd on one of 256 cursors,
//   4. writes the rows from the saved ballots (no second distance evaluation).
// Membership is decided exactly as in neigh_kernel (float32 guard band + float64 confirmation).
constexpr int kRowParts = 256;

struct RowsArgs {
    const void* pos_env;
    const void* pos_cent;
    int n_env, n_cent;
    const FrameParams* fp;
    const int* env_start;
    const int* cent_start;
    const float4* env_sorted;   // local representatives on shift-mode axes
    const float4* cent_sorted;
    int respect_pbc;
    double r2;
    int* counts;               // (F, M)
    long long* row_start;      // (F, M) absolute index into rows, -1 if the reservation overflowed
    int* rows;
    long long part_cap;        // capacity of each of the n_parts regions
    int n_parts;               // power of two <= kRowParts
    unsigned long long* cursors;  // [kRowParts]
    long long* status;         // [0] overflow flag
};

// scatter with local representatives: on a shift-mode axis the stored coordinate is the periodic
// image of x closest to the centre of the atom's (reference-formula) cell.
template <typename T>
__global__ void __launch_bounds__(256)
cell_scatter_local_kernel(const T* __restrict__ pos, int n_atoms,
                          const FrameParams* __restrict__ fp, const unsigned* __restrict__ cell_id,
                          int* __restrict__ cell_cnt, const int* __restrict__ cell_start,
                          float4* __restrict__ sorted, unsigned* __restrict__ slot_cell,
                          unsigned long long* __restrict__ rank_io = nullptr) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const FrameParams* p = fp + f;
    const size_t a = (size_t)f * n_atoms + i;
    const int c = (int)cell_id[a];
    const int gc = p->cell_base + c;
    int slot;
    if (rank_io) {   // rank from cell_assign_kernel; the word becomes the pair accumulator (zero)
        slot = cell_start[gc] + (int)rank_io[a];
        rank_io[a] = 0ull;
    } else {
        slot = cell_start[gc] + atomicSub(&cell_cnt[gc], 1) - 1;
    }
    const int ny = p->ncell[1], nz = p->ncell[2];
    const int ck[3] = {c / (ny * nz), (c / nz) % ny, c % nz};
    float out[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double x = (double)pos[a * 3 + k];
        if (p->shift_mode[k]) {
            const double edge = p->box[k] / (double)p->ncell[k];
            x -= rint((x - ((double)ck[k] + 0.5) * edge) / p->box[k]) * p->box[k];
        }
        out[k] = (float)x;
    }
    sorted[slot] = make_float4(out[0], out[1], out[2], __int_as_float(i));
    slot_cell[slot] = ((unsigned)ck[0] << 20) | ((unsigned)ck[1] << 10) | (unsigned)ck[2];
}

// Thread-per-centre variant of the rows path for sparse cells (a few atoms per cell, <= 32
// neighbours per centre) with every axis in shift mode (or pbc off).  A thread walks the 27 cells
// of its centre, stages its hits in shared memory ([hit][thread] layout, conflict-free), the warp
// reserves space for its 32 rows with ONE atomic and copies them out coalesced.  ~4x fewer issued
// instructions per centre than one warp per cell when cells hold 2..5 atoms.
constexpr int kTpcThreads = 128;

// kTpcHitCap = rows that fit the staging tile (32 or 64 hits per centre)
// 6 blocks of 128 threads per SM (80 registers, a few spilled words outside the candidate loop):
// the kernel waits on candidate loads, 24 instead of 20 resident warps is +3.7 % at cfg3; 8 blocks
// (64 registers) spill inside the loop and give the gain back
#ifndef DSG_TPC_MIN_BLOCKS
#define DSG_TPC_MIN_BLOCKS 6
#endif
template <typename T, int kTpcHitCap>
__global__ void __launch_bounds__(kTpcThreads, DSG_TPC_MIN_BLOCKS)
neigh_rows_thread_kernel(const RowsArgs a, const unsigned* __restrict__ cent_slot_cell) {
    __shared__ int stage[kTpcHitCap][kTpcThreads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.y;
    const FrameParams* __restrict__ p = a.fp + f;
    const int nx = p->ncell[0], ny = p->ncell[1], nz = p->ncell[2];
    const int slot_local = blockIdx.x * kTpcThreads + tid;
    const bool active = slot_local < a.n_cent;
    const bool pbc = a.respect_pbc != 0;
    const float bx = pbc ? p->boxf[0] : 0.f, by = pbc ? p->boxf[1] : 0.f, bz = pbc ? p->boxf[2] : 0.f;
    const float r2_lo = p->r2_lo, r2_hi = p->r2_hi;
    const int base_cell = p->cell_base;
    const T* __restrict__ pos_env_f = (const T*)a.pos_env + (size_t)f * a.n_env * 3;
    const T* __restrict__ pos_cent_f = (const T*)a.pos_cent + (size_t)f * a.n_cent * 3;

    int cnt = 0, my_id = 0;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0, cz = 0;
    if (active) {
        const size_t slot = (size_t)f * a.n_cent + slot_local;
        me = a.cent_sorted[slot];
        my_id = __float_as_int(me.w);
        const unsigned pc = cent_slot_cell[slot];
        cx = (int)(pc >> 20), cy = (int)((pc >> 10) & 1023u), cz = (int)(pc & 1023u);
    }
    // The 27 cells of the centre; hits go to the shared staging tile (gdst == nullptr) or, in the
    // rare second pass of a row longer than the tile, straight to the row's place in global memory.
    auto walk = [&](int* __restrict__ gdst) {
        // candidates of a contiguous run of cells, four loads in flight per step
        auto scan_run = [&](int qb, int qe, float ox, float oy, float oz) {
            for (int q0 = qb; q0 < qe; q0 += 4) {
                float4 cd[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    cd[u] = a.env_sorted[min(q0 + u, qe - 1)];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = __float_as_int(cd[u].w);
                    const float dx = cd[u].x + ox, dy = cd[u].y + oy, dz = cd[u].z + oz;
                    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    const bool valid = q0 + u < qe;
                    bool hit = d2 < r2_lo;
                    if (valid && !hit && !(d2 > r2_hi))   // inside the guard band: rare
                        hit = exact_test<T>(pos_env_f + (size_t)j * 3,
                                            pos_cent_f + (size_t)my_id * 3, p->box, pbc, a.r2);
                    hit = hit && valid && j != my_id;  // lens.py:104 "j != i"
                    if (hit) {
                        // branch-free staging: an overflowing row keeps rewriting the last slot
                        if (gdst == nullptr) stage[min(cnt, kTpcHitCap - 1)][tid] = j;
                        else gdst[cnt] = j;
                    }
                    cnt += hit ? 1 : 0;
                }
            }
        };
        const bool z_inside = cz >= 1 && cz <= nz - 2;  // the three z cells are one contiguous run
        for (int ddx = -1; ddx <= 1; ++ddx) {
            const int ax = cx + ddx;
            const int wx = ax < 0 ? ax + nx : (ax >= nx ? ax - nx : ax);
            const float ox = (ax < 0 ? -bx : (ax >= nx ? bx : 0.f)) - me.x;
            for (int ddy = -1; ddy <= 1; ++ddy) {
                const int ay = cy + ddy;
                const int wy = ay < 0 ? ay + ny : (ay >= ny ? ay - ny : ay);
                const float oy = (ay < 0 ? -by : (ay >= ny ? by : 0.f)) - me.y;
                const int rowc = base_cell + (wx * ny + wy) * nz;
                if (z_inside) {
                    scan_run(a.env_start[rowc + cz - 1], a.env_start[rowc + cz + 2], ox, oy, -me.z);
                } else {
                    for (int ddz = -1; ddz <= 1; ++ddz) {
                        const int az = cz + ddz;
                        const int wz = az < 0 ? az + nz : (az >= nz ? az - nz : az);
                        const float oz = (az < 0 ? -bz : (az >= nz ? bz : 0.f)) - me.z;
                        scan_run(a.env_start[rowc + wz], a.env_start[rowc + wz + 1], ox, oy, oz);
                    }
                }
            }
        }
    };
    if (active) walk(nullptr);
    // one reservation per warp for the rows that fit the tile; a longer row (rare) reserves its
    // own place and is written by a second walk of its thread
    const bool long_row = cnt > kTpcHitCap;
    const int cnt_t = long_row ? 0 : cnt;
    const int incl = warp_incl_scan(cnt_t, lane);
    const int total = __shfl_sync(kFull, incl, 31);
    const int part = (blockIdx.x * (kTpcThreads / 32) + warp + f * 131) & (a.n_parts - 1);
    long long base = 0;
    if (lane == 0 && total > 0) {
        const unsigned long long off = atomicAdd(&a.cursors[part], (unsigned long long)total);
        if ((long long)(off + total) > a.part_cap) {
            if (a.status[0] == 0) a.status[0] = 1;
            base = -1;
        } else {
            base = (long long)part * a.part_cap + (long long)off;
        }
    }
    if (blockIdx.x == 0 && f == 0 && tid == 0) a.status[1] = a.n_parts;
    base = __shfl_sync(kFull, base, 0);
    long long my_start = base < 0 ? -1 : base + (long long)(incl - cnt_t);
    if (long_row) {
        const int n_row = cnt;
        const unsigned long long off = atomicAdd(&a.cursors[part], (unsigned long long)n_row);
        if ((long long)(off + n_row) > a.part_cap) {
            if (a.status[0] == 0) a.status[0] = 1;
            my_start = -1;
        } else {
            my_start = (long long)part * a.part_cap + (long long)off;
            cnt = 0;
            walk(a.rows + my_start);
        }
        cnt = n_row;
    }
    if (active) {
        a.counts[(size_t)f * a.n_cent + my_id] = cnt;
        a.row_start[(size_t)f * a.n_cent + my_id] = my_start;
    }
    if (base < 0 || total == 0) return;
    __syncwarp();
    // coalesced copy-out: output position -> (thread, hit) by a shuffle search over the prefix
    int* __restrict__ dst = a.rows + base;
    for (int p0 = 0; p0 < total; p0 += 32) {
        const int pp = p0 + lane;
        int t = 0;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const int v = __shfl_sync(kFull, incl, (t + s - 1) & 31);
            if (v <= pp) t += s;
        }
        t &= 31;
        const int excl_t = __shfl_sync(kFull, incl - cnt_t, t);
        if (pp < total) dst[pp] = stage[pp - excl_t][warp * 32 + t];
    }
}

// RINT = some axis still needs the rint minimum image (fewer than 6 cells along it)
template <typename T, bool RINT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
neigh_rows_kernel(const RowsArgs a) {
    __shared__ float4 stage_all[kWarpsPerBlock][kStageCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const FrameParams* __restrict__ p = a.fp + f;
    const int nx = p->ncell[0], ny = p->ncell[1], nz = p->ncell[2];
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    // the number of reservation regions goes into the status block whatever cell 0 holds (the
    // caller sizes its retry from it)
    if (blockIdx.x == 0 && f == 0 && threadIdx.x == 0) a.status[1] = a.n_parts;
    if (c >= nx * ny * nz) return;
    const int base_cell = p->cell_base;
    const int cbeg = a.cent_start[base_cell + c];
    const int cend = a.cent_start[base_cell + c + 1];
    if (cbeg == cend) return;

    float4* stage = stage_all[warp];
    int cx, cy, cz;
    if (nx * ny * nz < (1 << 20)) {
        // float reciprocals are exact for floor((c + 1/2) / n) below 2^20 cells
        const int xy = (int)(((float)c + 0.5f) * __frcp_rn((float)nz));
        cz = c - xy * nz;
        cx = (int)(((float)xy + 0.5f) * __frcp_rn((float)ny));
        cy = xy - cx * ny;
    } else {
        cz = c % nz; cy = (c / nz) % ny; cx = c / (nz * ny);
    }
    const bool pbc = a.respect_pbc != 0;
    const bool sx_mode = pbc && p->shift_mode[0], sy_mode = pbc && p->shift_mode[1],
               sz_mode = pbc && p->shift_mode[2];
    const float bx = p->boxf[0], by = p->boxf[1], bz = p->boxf[2];

    int run_start = 0, run_len = 0;
    float rsx = 0.f, rsy = 0.f, rsz = 0.f;  // periodic shift of this lane's run
    if (lane < 27) {
        const int ddx = lane / 9 - 1, ddy = (lane / 3) % 3 - 1, ddz = lane % 3 - 1;
        const int ax = cx + ddx, ay = cy + ddy, az = cz + ddz;
        const int wx = ax < 0 ? ax + nx : (ax >= nx ? ax - nx : ax);
        const int wy = ay < 0 ? ay + ny : (ay >= ny ? ay - ny : ay);
        const int wz = az < 0 ? az + nz : (az >= nz ? az - nz : az);
        const int cc = (wx * ny + wy) * nz + wz;
        run_start = a.env_start[base_cell + cc];
        run_len = a.env_start[base_cell + cc + 1] - run_start;
        if (sx_mode) rsx = ax < 0 ? -bx : (ax >= nx ? bx : 0.f);
        if (sy_mode) rsy = ay < 0 ? -by : (ay >= ny ? by : 0.f);
        if (sz_mode) rsz = az < 0 ? -bz : (az >= nz ? bz : 0.f);
    }
    const int run_incl = warp_incl_scan(run_len, lane);
    const int run_excl = run_incl - run_len;
    const int n_cand = __shfl_sync(kFull, run_incl, 31);

    const float ibx = p->inv_boxf[0], iby = p->inv_boxf[1], ibz = p->inv_boxf[2];
    const bool px = RINT && pbc && p->box[0] > 0.0 && !sx_mode,
               py = RINT && pbc && p->box[1] > 0.0 && !sy_mode,
               pz = RINT && pbc && p->box[2] > 0.0 && !sz_mode;
    const float r2_lo = p->r2_lo, r2_hi = p->r2_hi;
    const unsigned lt_mask = (1u << lane) - 1u;
    const T* __restrict__ pos_env_f = (const T*)a.pos_env + (size_t)f * a.n_env * 3;
    const T* __restrict__ pos_cent_f = (const T*)a.pos_cent + (size_t)f * a.n_cent * 3;
    const double r2 = a.r2;
    const int part = (c + f * 131) & (a.n_parts - 1);

    // stage candidates [base, base+clen) with the run's shift applied
    auto stage_chunk = [&](int base, int clen) {
        for (int q0 = 0; q0 < clen; q0 += 32) {
            const int q = base + q0 + lane;
            const bool valid = q0 + lane < clen;
            int r = 0;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const int v = __shfl_sync(kFull, run_incl, (r + s - 1) & 31);
                if (v <= q) r += s;
            }
            r = valid ? r : 0;
            const int src = __shfl_sync(kFull, run_start, r) + (q - __shfl_sync(kFull, run_excl, r));
            const float ox = __shfl_sync(kFull, rsx, r), oy = __shfl_sync(kFull, rsy, r);
            const float oz = __shfl_sync(kFull, rsz, r);
            if (valid) {
                float4 v = a.env_sorted[src];
                v.x += ox; v.y += oy; v.z += oz;
                stage[q0 + lane] = v;
            }
        }
    };
    auto test = [&](const float4 cd, bool valid, float ccx, float ccy, float ccz, int cid) -> bool {
        const int j = __float_as_int(cd.w);
        float dx = cd.x - ccx, dy = cd.y - ccy, dz = cd.z - ccz;
        if (RINT) {
            if (px) dx = fmaf(-rintf(dx * ibx), bx, dx);
            if (py) dy = fmaf(-rintf(dy * iby), by, dy);
            if (pz) dz = fmaf(-rintf(dz * ibz), bz, dz);
        }
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        bool hit = d2 < r2_lo;
        if (valid && !hit && !(d2 > r2_hi))
            hit = exact_test<T>(pos_env_f + (size_t)j * 3, pos_cent_f + (size_t)cid * 3, p->box, pbc,
                                r2);
        return hit && valid && (j != cid);  // lens.py:104 "j != i"
    };
    // reserve `total` row entries for this warp; returns the absolute base or -1 on overflow
    auto reserve = [&](int total) -> long long {
        long long base = 0;
        if (lane == 0) {
            const unsigned long long off = atomicAdd(&a.cursors[part], (unsigned long long)total);
            if ((long long)(off + total) > a.part_cap) {
                a.status[0] = 1;
                base = -1;
            } else {
                base = (long long)part * a.part_cap + (long long)off;
            }
        }
        return __shfl_sync(kFull, base, 0);
    };

    const bool fast = n_cand <= kStageCap;
    if (fast) {
        stage_chunk(0, n_cand);
        __syncwarp();
    }
    for (int cb = cbeg; cb < cend; cb += 32) {
        const int nb = min(32, cend - cb);
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < nb) me = a.cent_sorted[cb + lane];
        const int my_id = __float_as_int(me.w);
        int my_cnt = 0;
        if (fast) {
            unsigned mk0 = 0u, mk1 = 0u, mk2 = 0u, mk3 = 0u;
            for (int ci = 0; ci < nb; ++ci) {
                const float ccx = __shfl_sync(kFull, me.x, ci), ccy = __shfl_sync(kFull, me.y, ci);
                const float ccz = __shfl_sync(kFull, me.z, ci);
                const int cid = __shfl_sync(kFull, my_id, ci);
                unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
                m0 = __ballot_sync(kFull, test(stage[lane], lane < n_cand, ccx, ccy, ccz, cid));
                if (n_cand > 32)
                    m1 = __ballot_sync(kFull, test(stage[32 + lane], 32 + lane < n_cand, ccx, ccy,
                                                   ccz, cid));
                if (n_cand > 64)
                    m2 = __ballot_sync(kFull, test(stage[64 + lane], 64 + lane < n_cand, ccx, ccy,
                                                   ccz, cid));
                if (n_cand > 96)
                    m3 = __ballot_sync(kFull, test(stage[96 + lane], 96 + lane < n_cand, ccx, ccy,
                                                   ccz, cid));
                if (lane == ci) {
                    mk0 = m0; mk1 = m1; mk2 = m2; mk3 = m3;
                    my_cnt = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
                }
            }
            const int incl = warp_incl_scan(my_cnt, lane);
            const int total = __shfl_sync(kFull, incl, 31);
            const long long base = total > 0 ? reserve(total) : 0;
            const long long my_row = base < 0 ? -1 : base + (long long)(incl - my_cnt);
            if (lane < nb) {
                a.counts[(size_t)f * a.n_cent + my_id] = my_cnt;
                a.row_start[(size_t)f * a.n_cent + my_id] = my_row;
            }
            if (base >= 0 && total > 0) {
                for (int ci = 0; ci < nb; ++ci) {
                    const long long row = __shfl_sync(kFull, my_row, ci);
                    const unsigned m0 = __shfl_sync(kFull, mk0, ci), m1 = __shfl_sync(kFull, mk1, ci);
                    const unsigned m2 = __shfl_sync(kFull, mk2, ci), m3 = __shfl_sync(kFull, mk3, ci);
                    int* __restrict__ dst = a.rows + row;
                    int cum = 0;
                    if ((m0 >> lane) & 1u) dst[__popc(m0 & lt_mask)] = __float_as_int(stage[lane].w);
                    cum = __popc(m0);
                    if ((m1 >> lane) & 1u)
                        dst[cum + __popc(m1 & lt_mask)] = __float_as_int(stage[32 + lane].w);
                    cum += __popc(m1);
                    if ((m2 >> lane) & 1u)
                        dst[cum + __popc(m2 & lt_mask)] = __float_as_int(stage[64 + lane].w);
                    cum += __popc(m2);
                    if ((m3 >> lane) & 1u)
                        dst[cum + __popc(m3 & lt_mask)] = __float_as_int(stage[96 + lane].w);
                }
            }
        } else {
            // many candidates: chunked count traversal, one reservation, chunked fill traversal
            long long my_row = 0;
            for (int pass = 0; pass < 2; ++pass) {
                int run_cnt = 0;  // entries of my centre written / counted so far
                for (int base = 0; base < n_cand; base += kStageCap) {
                    const int clen = min(kStageCap, n_cand - base);
                    __syncwarp();
                    stage_chunk(base, clen);
                    __syncwarp();
                    for (int ci = 0; ci < nb; ++ci) {
                        const float ccx = __shfl_sync(kFull, me.x, ci);
                        const float ccy = __shfl_sync(kFull, me.y, ci);
                        const float ccz = __shfl_sync(kFull, me.z, ci);
                        const int cid = __shfl_sync(kFull, my_id, ci);
                        const long long row = __shfl_sync(kFull, my_row, ci);
                        const int cur = __shfl_sync(kFull, run_cnt, ci);
                        int add = 0;
                        for (int q0 = 0; q0 < clen; q0 += 32) {
                            const bool valid = q0 + lane < clen;
                            const float4 cd = stage[valid ? q0 + lane : 0];
                            const bool hit = test(cd, valid, ccx, ccy, ccz, cid);
                            const unsigned m = __ballot_sync(kFull, hit);
                            if (pass == 1 && hit && row >= 0)
                                a.rows[row + cur + add + __popc(m & lt_mask)] = __float_as_int(cd.w);
                            add += __popc(m);
                        }
                        if (lane == ci) run_cnt += add;
                    }
                }
                if (pass == 0) {
                    my_cnt = run_cnt;
                    const int incl = warp_incl_scan(my_cnt, lane);
                    const int total = __shfl_sync(kFull, incl, 31);
                    const long long base = total > 0 ? reserve(total) : 0;
                    my_row = base < 0 ? -1 : base + (long long)(incl - my_cnt);
                    if (lane < nb) {
                        a.counts[(size_t)f * a.n_cent + my_id] = my_cnt;
                        a.row_start[(size_t)f * a.n_cent + my_id] = my_row;
                    }
                    if (total == 0) break;
                }
            }
        }
    }
}

template <typename T>
static int neigh_rows_impl(const void* d_pos_env, const void* d_pos_cent, int n_frames, int n_env,
                           int n_cent, const double* h_box, double r_cut, int respect_pbc,
                           void* d_ws, size_t ws_bytes, int32_t* d_counts, int64_t* d_row_start,
                           int32_t* d_rows, int64_t rows_capacity, int64_t* d_status,
                           int kernel_choice, cudaStream_t st) {
    const bool same = d_pos_cent == nullptr;
    if (same && n_cent != n_env)
        return set_error(DSG_ERR_INVALID_ARGUMENT,
                         "d_pos_cent is NULL (centres == env) but n_cent=%d != n_env=%d", n_cent,
                         n_env);
    NeighLayout L;
    std::vector<FrameParams> fps;
    int rc = make_layout(n_frames, n_env, n_cent, same, h_box, r_cut, &L, &fps);
    if (rc) return rc;
    if (!d_ws || ws_bytes < L.bytes)
        return set_error(DSG_ERR_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed",
                         ws_bytes, L.bytes);
    if (!d_pos_env || !d_counts || !d_row_start || !d_rows || !d_status)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (rows_capacity < kRowParts)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "rows_capacity=%lld is too small",
                         (long long)rows_capacity);
    for (auto& p : fps)
        for (int k = 0; k < 3; ++k) p.shift_mode[k] = (respect_pbc && p.ncell[k] >= 6) ? 1 : 0;
    char* ws = (char*)d_ws;
    FrameParams* d_fp = (FrameParams*)(ws + L.off_fp);
    unsigned* cid_e = (unsigned*)(ws + L.off_cell_id_e);
    int* cnt_e = (int*)(ws + L.off_cnt_e);
    int* start_e = (int*)(ws + L.off_start_e);
    float4* sorted_e = (float4*)(ws + L.off_sorted_e);
    unsigned* cid_c = (unsigned*)(ws + L.off_cell_id_c);
    int* cnt_c = (int*)(ws + L.off_cnt_c);
    int* start_c = (int*)(ws + L.off_start_c);
    float4* sorted_c = (float4*)(ws + L.off_sorted_c);
    int* tile_tmp = (int*)(ws + L.off_tile_tmp);
    // lens.py:78 "dr2 < (r_cut - 1e-6)^2"; the inclusive mode (bit 1 of kernel_choice) is the
    // FastNS rule "dr2 <= r_cut^2" of analysis/spatial_average.py:39-40, kept as a strict test
    // against the next float64 above r_cut^2
    const double rc_eff = r_cut - 1e-6;
    const double r2 = (kernel_choice & 2) ? std::nextafter(r_cut * r_cut, INFINITY)
                                          : rc_eff * rc_eff;

    DSG_CUDA_CHECK(cudaMemcpyAsync(d_fp, fps.data(), sizeof(FrameParams) * n_frames,
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemsetAsync(d_status, 0, 8 * (2 + kRowParts), st));
    unsigned* slot_cell_e = (unsigned*)(ws + L.off_slot_cell_e);
    unsigned* slot_cell_c = (unsigned*)(ws + L.off_slot_cell_c);
    auto build = [&](const T* pos, int n_atoms, unsigned* cid, int* cnt, int* start,
                     float4* sorted, unsigned* slot_cell) -> int {
        dim3 grid((n_atoms + 255) / 256, n_frames);
        DSG_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 4 * (size_t)L.total_cells, st));
        cell_assign_kernel<T><<<grid, 256, 0, st>>>(pos, n_atoms, d_fp, cid, cnt);
        DSG_LAUNCH_CHECK("cell_assign_kernel");
        int r = run_scan(cnt, L.total_cells, 1, start, tile_tmp, nullptr, st);
        if (r) return r;
        cell_scatter_local_kernel<T><<<grid, 256, 0, st>>>(pos, n_atoms, d_fp, cid, cnt, start,
                                                           sorted, slot_cell);
        DSG_LAUNCH_CHECK("cell_scatter_local_kernel");
        return DSG_OK;
    };
    rc = build((const T*)d_pos_env, n_env, cid_e, cnt_e, start_e, sorted_e, slot_cell_e);
    if (rc) return rc;
    if (!same) {
        rc = build((const T*)d_pos_cent, n_cent, cid_c, cnt_c, start_c, sorted_c, slot_cell_c);
        if (rc) return rc;
    }
    frame_finalize_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(d_fp, n_frames, r_cut, r2);
    DSG_LAUNCH_CHECK("frame_finalize_kernel");

    RowsArgs a;
    a.pos_env = d_pos_env;
    a.pos_cent = same ? d_pos_env : d_pos_cent;
    a.n_env = n_env;
    a.n_cent = n_cent;
    a.fp = d_fp;
    a.env_start = start_e;
    a.cent_start = start_c;
    a.env_sorted = sorted_e;
    a.cent_sorted = sorted_c;
    a.respect_pbc = respect_pbc;
    a.r2 = r2;
    a.counts = d_counts;
    a.row_start = (long long*)d_row_start;
    a.rows = d_rows;
    // few, large cells: fewer regions (atomics are no bottleneck then, balance is)
    int n_parts = 1;
    while (n_parts < kRowParts && (long long)n_parts * 512 <= L.total_cells) n_parts <<= 1;
    a.n_parts = n_parts;
    a.part_cap = rows_capacity / n_parts;
    a.status = (long long*)d_status;
    a.cursors = (unsigned long long*)(d_status + 2);
    dim3 grid((L.max_cells + kWarpsPerBlock - 1) / kWarpsPerBlock, n_frames);
    bool any_rint = false, packable = true;
    double cand_sum = 0.0, k_max = 0.0;  // candidates per centre, expected neighbours
    for (const auto& p : fps) {
        for (int k = 0; k < 3; ++k) {
            if (respect_pbc) any_rint = any_rint || !p.shift_mode[k];
            packable = packable && p.ncell[k] <= 1024;
        }
        cand_sum += 27.0 * n_env / ((double)p.ncell[0] * p.ncell[1] * p.ncell[2]);
        const double k_est = n_env / (p.box[0] * p.box[1] * p.box[2]) * 4.18879 * r_cut * r_cut * r_cut;
        if (k_est > k_max) k_max = k_est;
    }
    // sparse cells (few candidates per centre): one thread per centre issues ~4x fewer instructions
    // rows longer than the staging tile are written by a second walk of their thread, so the
    // tile only has to hold the typical row: mean + 3 sigma of a Poisson count
    const double k_hi = k_max + 3.0 * std::sqrt(k_max);
    const bool use_thread_kernel = !(kernel_choice & 1) && !any_rint && packable &&
                                   cand_sum / n_frames <= 200.0 && k_hi <= 64.0;
    if (use_thread_kernel) {
        dim3 tgrid((n_cent + kTpcThreads - 1) / kTpcThreads, n_frames);
        if (k_hi <= 32.0)
            neigh_rows_thread_kernel<T, 32><<<tgrid, kTpcThreads, 0, st>>>(a, slot_cell_c);
        else
            neigh_rows_thread_kernel<T, 64><<<tgrid, kTpcThreads, 0, st>>>(a, slot_cell_c);
    } else if (any_rint) neigh_rows_kernel<T, true><<<grid, kWarpsPerBlock * 32, 0, st>>>(a);
    else neigh_rows_kernel<T, false><<<grid, kWarpsPerBlock * 32, 0, st>>>(a);
    DSG_LAUNCH_CHECK("neigh_rows_kernel");
    return DSG_OK;
}


// ------------------------------------------------------------------------------------------
// Pair path: coordination numbers and LENS straight from the half shell -- no lists at all
// ------------------------------------------------------------------------------------------
// With centres == environment the neighbour relation of the reference is symmetric (the 27-cell
// adjacency is symmetric, and lens.py:17-26 / :100-104 give bit-identical dr2 for (i, j) and
// (j, i): rint is odd), and LENS only needs three integers per centre and frame pair:
//   a = |C_t(i)|,  b = |C_{t+delay}(i)|,  inter = |C_t(i) n C_{t+delay}(i)|   (lens.py:167-188).
// One thread per atom of frame t walks HALF of the 27 cells (own cell behind its own slot + the 13
// cells that follow in (x, y, z) order; the z triples of a column are contiguous, so 5 runs) with
// the same float32 test + float64 confirmation as neigh_rows_thread_kernel.  Every hit (i, j):
//   * counts once for i (register) and once for j (RED.ADD on acc[t][j]);
//   * is looked up again in frame t+delay -- j in C_{t+delay}(i) <=> dr2 < r2 there (plain float32
//     minimum image of the caller's coordinates inside its own guard band, float64 confirmation)
//     and, when that frame has atoms outside [0, box) whose reference cells are not geometric,
//     the two reference cells are adjacent -- and the answer goes to the upper half of the same
//     64-bit word.
// Half the distance tests of the rows path, no rows written, reserved, copied, re-read or merged.
struct PairArgs {
    const void* pos;               // (F, N, 3) caller's positions
    int n_atoms, n_frames, delay;  // delay = 0: coordination numbers only
    const FrameParams* fp;
    const int* cell_start;
    const float4* sorted;          // local representatives on shift-mode axes
    const unsigned* slot_cell;     // packed (cx, cy, cz) of a sorted slot
    const unsigned* cell_id;       // (F, N) reference cell of an atom (by atom index)
    int respect_pbc;
    double r2;
    unsigned long long* acc;       // (F, N): count | intersection << 32
};

#ifndef DSG_PAIR_THREADS
#define DSG_PAIR_THREADS 128
#endif
constexpr int kPairThreads = DSG_PAIR_THREADS;
constexpr int kPairCap = 32;   // staged hits per thread; a longer half row is settled inline
#ifndef DSG_PAIR_MIN_BLOCKS
#define DSG_PAIR_MIN_BLOCKS 6
#endif

template <typename T>
__global__ void __launch_bounds__(kPairThreads, DSG_PAIR_MIN_BLOCKS)
lens_pair_kernel(const PairArgs a) {
    __shared__ int stage[kPairCap][kPairThreads];
    const int tid = threadIdx.x;
    const int f = blockIdx.y;
    const int slot_local = blockIdx.x * kPairThreads + tid;
    if (slot_local >= a.n_atoms) return;   // threads are independent (no warp collectives)
    const FrameParams* __restrict__ p = a.fp + f;
    const int nx = p->ncell[0], ny = p->ncell[1], nz = p->ncell[2];
    const bool pbc = a.respect_pbc != 0;
    const float bx = pbc ? p->boxf[0] : 0.f, by = pbc ? p->boxf[1] : 0.f, bz = pbc ? p->boxf[2] : 0.f;
    const float r2_lo = p->r2_lo, r2_hi = p->r2_hi;
    const int base_cell = p->cell_base;
    const size_t n = (size_t)a.n_atoms;
    const T* __restrict__ pos_f = (const T*)a.pos + (size_t)f * n * 3;
    const size_t slot = (size_t)f * n + slot_local;
    const float4 me = a.sorted[slot];
    const int my_id = __float_as_int(me.w);
    const unsigned pc = a.slot_cell[slot];
    const int cx = (int)(pc >> 20), cy = (int)((pc >> 10) & 1023u), cz = (int)(pc & 1023u);
    unsigned long long* __restrict__ acc_f = a.acc + (size_t)f * n;

    // partner frame t + delay
    const int f2 = f + a.delay;
    const bool paired = a.delay > 0 && f2 < a.n_frames;
    const FrameParams* __restrict__ p2 = a.fp + (paired ? f2 : f);
    const T* __restrict__ pos_2 = (const T*)a.pos + (size_t)(paired ? f2 : f) * n * 3;
    const unsigned* __restrict__ cell_2 = a.cell_id + (size_t)(paired ? f2 : f) * n;
    const float b2x = p2->boxf[0], b2y = p2->boxf[1], b2z = p2->boxf[2];
    const float i2x = pbc ? p2->inv_boxf[0] : 0.f, i2y = pbc ? p2->inv_boxf[1] : 0.f,
                i2z = pbc ? p2->inv_boxf[2] : 0.f;
    const float lo2 = p2->r2_lo_img, hi2 = p2->r2_hi_img;
    const bool regular2 = p2->regular != 0;
    float x2 = 0.f, y2 = 0.f, z2 = 0.f;
    if (paired) {
        x2 = (float)pos_2[(size_t)my_id * 3];
        y2 = (float)pos_2[(size_t)my_id * 3 + 1];
        z2 = (float)pos_2[(size_t)my_id * 3 + 2];
    }
    int own_inter = 0;
    // one hit (my_id, j) of frame t: count it for j and decide membership in frame t + delay
    auto settle = [&](int j, float xj, float yj, float zj) {
        unsigned long long inc = 1ull;
        if (paired) {
            const T* __restrict__ pj = pos_2 + (size_t)j * 3;
            float dx = xj - x2, dy = yj - y2, dz = zj - z2;
            dx = fmaf(-rintf(dx * i2x), b2x, dx);   // i2 = 0 without pbc: the raw difference
            dy = fmaf(-rintf(dy * i2y), b2y, dy);
            dz = fmaf(-rintf(dz * i2z), b2z, dz);
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            bool hit2 = d2 < lo2;
            if (!hit2 && !(d2 > hi2))   // inside the guard band: rare
                hit2 = exact_test<T>(pj, pos_2 + (size_t)my_id * 3, p2->box, pbc, a.r2);
            if (hit2 && !regular2) {
                // atoms outside [0, box) in that frame: the reference only finds j if their
                // (truncation + modulo) cells are adjacent (lens.py:92-97)
                const int m2y = p2->ncell[1], m2z = p2->ncell[2], m2x = p2->ncell[0];
                const int ci = (int)cell_2[my_id], cj = (int)cell_2[j];
                const int ex = abs(ci / (m2y * m2z) - cj / (m2y * m2z));
                const int ey = abs((ci / m2z) % m2y - (cj / m2z) % m2y);
                const int ez = abs(ci % m2z - cj % m2z);
                hit2 = (ex <= 1 || ex == m2x - 1) && (ey <= 1 || ey == m2y - 1) &&
                       (ez <= 1 || ez == m2z - 1);
            }
            own_inter += hit2 ? 1 : 0;
            inc += hit2 ? (1ull << 32) : 0ull;
        }
        atomicAdd(&acc_f[j], inc);
    };

    // The half shell as five columns of cells: c = 0 is my own column (the slots behind mine in my
    // cell, then the cell above), c = 1 the column (cx, cy + 1), c = 2..4 the columns (cx + 1,
    // cy - 1 .. cy + 1), each with the cells cz - 1 .. cz + 1.  The z cells of a column are
    // contiguous in the sorted array: one run per column, or one per cell at the z wrap.  One
    // call site for everything (the kernel stays small enough for the instruction cache); hits
    // number base .. base + kPairCap - 1 of the walk go to the staging tile, so a thread with more
    // hits than the tile holds (rare) walks again with the next base.
    // Hit number h of a walk goes to stage[h - base][tid]; sp points at the slot of the next hit
    // (before or beyond the tile = not staged in this pass), so the hit counter and the store
    // address are one register.
    constexpr int kSlotBytes = kPairThreads * (int)sizeof(int);
    // |d2 - mid| <= half  <=  d2 inside the guard band.  mid, half and d2 - mid are each rounded
    // (an ulp of r2 apiece), so half is widened by 8 ulps of r2_hi: a candidate sent to the
    // confirmation by mistake is still decided exactly, one that slipped past it would not be.
    // A useless band (infinite) confirms everything.
    const bool band_all = !(r2_hi < INFINITY);
    const float r2_mid = band_all ? 0.f : 0.5f * (r2_lo + r2_hi);
    const float r2_half =
        band_all ? INFINITY : 0.5f * (r2_hi - r2_lo) + 8.f * 1.1920929e-7f * r2_hi;
    char* const tile = reinterpret_cast<char*>(&stage[0][tid]);
    int cnt = 0, base = 0;
    do {
        char* sp = tile - base * kSlotBytes;
        // run bounds of column c, loaded one column ahead (they feed the candidate loads: two
        // dependent round trips per column otherwise); only meaningful away from the z wrap
        auto column_of = [&](int c, int& rowc, float& ox, float& oy) {
            const int ddx = c >= 2 ? 1 : 0, ddy = c == 0 ? 0 : (c == 1 ? 1 : c - 3);
            const int ax = cx + ddx, ay = cy + ddy;
            const int wx = ax >= nx ? ax - nx : ax;
            const int wy = ay < 0 ? ay + ny : (ay >= ny ? ay - ny : ay);
            ox = (ax >= nx ? bx : 0.f) - me.x;
            oy = (ay < 0 ? -by : (ay >= ny ? by : 0.f)) - me.y;
            rowc = base_cell + (wx * ny + wy) * nz;
        };
        const bool z_mid = cz >= 1 && cz + 1 < nz;   // every column is one run
        int rowc_n, qb_n = 0, qe_n = 0;
        float ox_n, oy_n;
        column_of(0, rowc_n, ox_n, oy_n);
        if (z_mid) {
            qb_n = a.cell_start[rowc_n + cz];
            qe_n = a.cell_start[rowc_n + cz + 2];
        }
#pragma unroll 1
        for (int c = 0; c < 5; ++c) {
            const int rowc = rowc_n, qb_pre = qb_n, qe_pre = qe_n;
            const float ox = ox_n, oy = oy_n;
            if (c < 4) {
                column_of(c + 1, rowc_n, ox_n, oy_n);
                if (z_mid) {
                    qb_n = a.cell_start[rowc_n + cz - 1];
                    qe_n = a.cell_start[rowc_n + cz + 2];
                }
            }
            const int z_lo = c == 0 ? cz : cz - 1, z_hi = cz + 1;
            const bool one_run = z_lo >= 0 && z_hi < nz;
            const int n_seg = one_run ? 1 : z_hi - z_lo + 1;
#pragma unroll 1
            for (int sg = 0; sg < n_seg; ++sg) {
                int qb, qe;
                float oz = -me.z;
                if (z_mid) {
                    qb = qb_pre;
                    qe = qe_pre;
                } else if (one_run) {
                    qb = a.cell_start[rowc + z_lo];
                    qe = a.cell_start[rowc + z_hi + 1];
                } else {
                    const int az = z_lo + sg;
                    const int wz = az < 0 ? az + nz : (az >= nz ? az - nz : az);
                    oz += az < 0 ? -bz : (az >= nz ? bz : 0.f);
                    qb = a.cell_start[rowc + wz];
                    qe = a.cell_start[rowc + wz + 1];
                }
                if (c == 0 && sg == 0) qb = (int)slot + 1;   // my own cell: only the slots behind me
                // Four candidates per step.  The last step of a run reads up to three slots past
                // it (the next cells; the sorted array is padded at its end) and masks them out.
                // A run whose candidates all fit behind the hits staged so far needs no bounds
                // check on the tile (CHECKED = false: almost always).
                auto scan = [&](auto checked) {
                    constexpr bool CHECKED = decltype(checked)::value;
                    auto put = [&](bool hit, float w) {
                        const bool room =
                            !CHECKED || (unsigned)(sp - tile) < (unsigned)(kPairCap * kSlotBytes);
                        if (hit && room) *reinterpret_cast<int*>(sp) = __float_as_int(w);
                        if (hit) sp += kSlotBytes;
                    };
                    for (int q0 = qb; q0 < qe; q0 += 4) {
                        const float4* __restrict__ cp = a.sorted + q0;
                        float4 cd[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) cd[u] = cp[u];
                        const int n_valid = qe - q0;
                        float d2[4];
                        bool band = false;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float dx = cd[u].x + ox, dy = cd[u].y + oy, dz = cd[u].z + oz;
                            d2[u] = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                            band = band || fabsf(d2[u] - r2_mid) <= r2_half;
                            // j is never my_id (my own slot is in no run): lens.py:104 by construction
                            put(d2[u] < r2_lo && u < n_valid, cd[u].w);
                        }
                        if (band) {   // a candidate of the step sits inside the guard band: rare
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (!(d2[u] < r2_lo) && !(d2[u] > r2_hi) && u < n_valid &&
                                    exact_test<T>(pos_f + (size_t)__float_as_int(cd[u].w) * 3,
                                                  pos_f + (size_t)my_id * 3, p->box, pbc, a.r2))
                                    put(true, cd[u].w);
                            }
                        }
                    }
                };
                const int soff = (int)(sp - tile);
                if (soff >= 0 && soff + (qe - qb) * kSlotBytes <= kPairCap * kSlotBytes)
                    scan(std::false_type{});
                else
                    scan(std::true_type{});
            }
        }
        cnt = (int)(sp - tile) / kSlotBytes + base;
        const int n_stage = min(cnt - base, kPairCap);
        // four hits per step: their partner-frame coordinates are gathered together
#pragma unroll 1
        for (int h0 = 0; h0 < n_stage; h0 += 4) {
            int jj[4];
            float px[4], py[4], pz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                jj[u] = stage[min(h0 + u, n_stage - 1)][tid];
                if (paired) {
                    const T* __restrict__ pj = pos_2 + (size_t)jj[u] * 3;
                    px[u] = (float)pj[0];
                    py[u] = (float)pj[1];
                    pz[u] = (float)pj[2];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (h0 + u < n_stage) settle(jj[u], px[u], py[u], pz[u]);
        }
        base += kPairCap;
    } while (base < cnt);
    atomicAdd(&acc_f[my_id], (unsigned long long)cnt | ((unsigned long long)own_inter << 32));
}

// acc -> counts (F, N) int32 and LENS (N, n_pairs) float64.  One thread owns one atom and a run
// of eight consecutive frames: the 64-bit words of a frame are read coalesced across the atoms of
// a warp, and every thread writes its eight results as one contiguous 64-byte piece of its output
// row (two full sectors) -- no shared-memory transpose, no barrier.
constexpr int kFinRun = 8;
__global__ void __launch_bounds__(256)
pair_finalize_kernel(const unsigned long long* __restrict__ acc, int n_frames, int n_atoms,
                     int delay, int* __restrict__ counts, double* __restrict__ lens,
                     long long ld, long long col0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int t0 = blockIdx.y * kFinRun;
    if (i >= n_atoms) return;
    const bool with_lens = lens != nullptr && delay > 0;
    unsigned long long w[kFinRun];
#pragma unroll
    for (int k = 0; k < kFinRun; ++k)
        w[k] = t0 + k < n_frames ? acc[(size_t)(t0 + k) * n_atoms + i] : 0ull;
    if (counts) {
#pragma unroll
        for (int k = 0; k < kFinRun; ++k)
            if (t0 + k < n_frames) counts[(size_t)(t0 + k) * n_atoms + i] = (int)(unsigned)w[k];
    }
    if (!with_lens) return;
    double* __restrict__ row = lens + (long long)i * ld + col0;
#pragma unroll
    for (int k = 0; k < kFinRun; ++k) {
        const int t = t0 + k;
        if (t + delay < n_frames) {
            // delay 1: the partner word is the next one of my own run (static register index)
            const unsigned long long w2 = (delay == 1 && k + 1 < kFinRun)
                                              ? w[k + 1 < kFinRun ? k + 1 : 0]
                                              : acc[(size_t)(t + delay) * n_atoms + i];
            row[t] = lens_value((int)(unsigned)w[k], (int)(unsigned)w2, (int)(unsigned)(w[k] >> 32));
        }
    }
}

// The pair path applies when one thread per atom is the right granularity (sparse cells, short
// rows), the slot -> cell packing fits and, under pbc, every axis is in shift mode.
static bool pair_path_supported(const std::vector<FrameParams>& fps, int n_atoms, double r_cut,
                                int respect_pbc) {
    double cand_sum = 0.0, k_max = 0.0;
    for (const auto& p : fps) {
        for (int k = 0; k < 3; ++k) {
            if (respect_pbc && p.ncell[k] < 6) return false;
            if (p.ncell[k] > 1024) return false;
        }
        cand_sum += 27.0 * n_atoms / ((double)p.ncell[0] * p.ncell[1] * p.ncell[2]);
        const double k_est =
            n_atoms / (p.box[0] * p.box[1] * p.box[2]) * 4.18879 * r_cut * r_cut * r_cut;
        if (k_est > k_max) k_max = k_est;
    }
    const double half = 0.5 * k_max;
    return cand_sum / (double)fps.size() <= 200.0 && half + 4.0 * std::sqrt(half) <= kPairCap;
}

struct PairLayout {
    NeighLayout L;
    size_t off_acc = 0, bytes = 0;
};

static int pair_layout(int n_frames, int n_atoms, const double* h_box, double r_cut, PairLayout* P,
                       std::vector<FrameParams>* fps) {
    int rc = make_layout(n_frames, n_atoms, n_atoms, true, h_box, r_cut, &P->L, fps);
    if (rc) return rc;
    P->off_acc = align_up(P->L.bytes);
    P->bytes = P->off_acc + align_up(8 * (size_t)n_frames * n_atoms);
    return DSG_OK;
}

template <typename T>
static int lens_direct_impl(const void* d_pos, int n_frames, int n_atoms, const double* h_box,
                            double r_cut, int respect_pbc, int delay, void* d_ws, size_t ws_bytes,
                            int32_t* d_counts, double* d_lens, int64_t lens_stride, int64_t col0,
                            cudaStream_t st) {
    PairLayout P;
    std::vector<FrameParams> fps;
    int rc = pair_layout(n_frames, n_atoms, h_box, r_cut, &P, &fps);
    if (rc) return rc;
    if (!pair_path_supported(fps, n_atoms, r_cut, respect_pbc))
        return set_error(DSG_ERR_UNSUPPORTED,
                         "the pair path needs sparse cells and >= 6 cells per periodic axis; use "
                         "dsg_neigh_rows + dsg_lens_pairs_rows");
    if (!d_ws || ws_bytes < P.bytes)
        return set_error(DSG_ERR_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed",
                         ws_bytes, P.bytes);
    if (!d_pos || (!d_counts && !d_lens))
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (delay < 0 || (d_lens && (delay < 1 || delay >= n_frames)))
        return set_error(DSG_ERR_INVALID_ARGUMENT, "delay=%d needs 1 <= delay < n_frames=%d", delay,
                         n_frames);
    const NeighLayout& L = P.L;
    for (auto& p : fps)
        for (int k = 0; k < 3; ++k) p.shift_mode[k] = respect_pbc ? 1 : 0;
    char* ws = (char*)d_ws;
    FrameParams* d_fp = (FrameParams*)(ws + L.off_fp);
    unsigned* cid = (unsigned*)(ws + L.off_cell_id_e);
    int* cnt = (int*)(ws + L.off_cnt_e);
    int* start = (int*)(ws + L.off_start_e);
    float4* sorted = (float4*)(ws + L.off_sorted_e);
    int* tile_tmp = (int*)(ws + L.off_tile_tmp);
    unsigned* slot_cell = (unsigned*)(ws + L.off_slot_cell_e);
    unsigned long long* acc = (unsigned long long*)(ws + P.off_acc);
    const double rc_eff = r_cut - 1e-6;   // lens.py:78
    const double r2 = rc_eff * rc_eff;

    DSG_CUDA_CHECK(cudaMemcpyAsync(d_fp, fps.data(), sizeof(FrameParams) * n_frames,
                                   cudaMemcpyHostToDevice, st));
    DSG_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 4 * (size_t)L.total_cells, st));
    // the accumulator words first carry the rank of every atom inside its cell (assign -> scatter),
    // the scatter leaves them zero for the pair kernel: no memset, no second atomic per atom
    dim3 agrid((n_atoms + 255) / 256, n_frames);
    cell_assign_kernel<T><<<agrid, 256, 0, st>>>((const T*)d_pos, n_atoms, d_fp, cid, cnt, acc);
    DSG_LAUNCH_CHECK("cell_assign_kernel");
    rc = run_scan(cnt, L.total_cells, 1, start, tile_tmp, nullptr, st);
    if (rc) return rc;
    cell_scatter_local_kernel<T><<<agrid, 256, 0, st>>>((const T*)d_pos, n_atoms, d_fp, cid, cnt,
                                                        start, sorted, slot_cell, acc);
    DSG_LAUNCH_CHECK("cell_scatter_local_kernel");
    frame_finalize_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(d_fp, n_frames, r_cut, r2);
    DSG_LAUNCH_CHECK("frame_finalize_kernel");

    PairArgs a;
    a.pos = d_pos;
    a.n_atoms = n_atoms;
    a.n_frames = n_frames;
    a.delay = d_lens ? delay : 0;
    a.fp = d_fp;
    a.cell_start = start;
    a.sorted = sorted;
    a.slot_cell = slot_cell;
    a.cell_id = cid;
    a.respect_pbc = respect_pbc;
    a.r2 = r2;
    a.acc = acc;
    dim3 pgrid((n_atoms + kPairThreads - 1) / kPairThreads, n_frames);
    lens_pair_kernel<T><<<pgrid, kPairThreads, 0, st>>>(a);
    DSG_LAUNCH_CHECK("lens_pair_kernel");
    dim3 fgrid((n_atoms + 255) / 256, (n_frames + kFinRun - 1) / kFinRun);
    pair_finalize_kernel<<<fgrid, 256, 0, st>>>(acc, n_frames, n_atoms, a.delay, d_counts, d_lens,
                                                lens_stride, col0);
    DSG_LAUNCH_CHECK("pair_finalize_kernel");
    return DSG_OK;
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_neigh_workspace_bytes(int n_frames, int n_env, int n_cent, int same_centres,
                                         const double* h_box, double r_cut, size_t* bytes_out) {
    if (!bytes_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "bytes_out is NULL");
    NeighLayout L;
    int rc = make_layout(n_frames, n_env, n_cent, same_centres != 0, h_box, r_cut, &L, nullptr);
    if (rc) return rc;
    *bytes_out = L.bytes;
    return DSG_OK;
}

extern "C" int dsg_neigh_count(const void* d_pos_env, const void* d_pos_cent, int pos_dtype,
                               int n_frames, int n_env, int n_cent, const double* h_box,
                               double r_cut, int respect_pbc, void* d_workspace,
                               size_t workspace_bytes, int32_t* d_counts, int32_t* d_indptr,
                               int64_t* d_frame_off, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (pos_dtype == DSG_F32)
        return neigh_impl<float>(false, d_pos_env, d_pos_cent, n_frames, n_env, n_cent, h_box,
                                 r_cut, respect_pbc, d_workspace, workspace_bytes, d_counts,
                                 d_indptr, d_frame_off, nullptr, st);
    if (pos_dtype == DSG_F64)
        return neigh_impl<double>(false, d_pos_env, d_pos_cent, n_frames, n_env, n_cent, h_box,
                                  r_cut, respect_pbc, d_workspace, workspace_bytes, d_counts,
                                  d_indptr, d_frame_off, nullptr, st);
    return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                     pos_dtype);
}

extern "C" int dsg_neigh_fill(const void* d_pos_env, const void* d_pos_cent, int pos_dtype,
                              int n_frames, int n_env, int n_cent, const double* h_box,
                              double r_cut, int respect_pbc, void* d_workspace,
                              size_t workspace_bytes, const int32_t* d_indptr,
                              const int64_t* d_frame_off, int32_t* d_indices, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (pos_dtype == DSG_F32)
        return neigh_impl<float>(true, d_pos_env, d_pos_cent, n_frames, n_env, n_cent, h_box,
                                 r_cut, respect_pbc, d_workspace, workspace_bytes, nullptr,
                                 const_cast<int32_t*>(d_indptr),
                                 const_cast<int64_t*>(d_frame_off), d_indices, st);
    if (pos_dtype == DSG_F64)
        return neigh_impl<double>(true, d_pos_env, d_pos_cent, n_frames, n_env, n_cent, h_box,
                                  r_cut, respect_pbc, d_workspace, workspace_bytes, nullptr,
                                  const_cast<int32_t*>(d_indptr),
                                  const_cast<int64_t*>(d_frame_off), d_indices, st);
    return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                     pos_dtype);
}

extern "C" int dsg_neigh_rows(const void* d_pos_env, const void* d_pos_cent, int pos_dtype,
                              int n_frames, int n_env, int n_cent, const double* h_box,
                              double r_cut, int respect_pbc, void* d_workspace,
                              size_t workspace_bytes, int32_t* d_counts, int64_t* d_row_start,
                              int32_t* d_rows, int64_t rows_capacity, int64_t* d_status,
                              void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (pos_dtype == DSG_F32)
        return neigh_rows_impl<float>(d_pos_env, d_pos_cent, n_frames, n_env, n_cent, h_box, r_cut,
                                      respect_pbc & 1, d_workspace, workspace_bytes, d_counts,
                                      d_row_start, d_rows, rows_capacity, d_status,
                                      respect_pbc >> 1, st);
    if (pos_dtype == DSG_F64)
        return neigh_rows_impl<double>(d_pos_env, d_pos_cent, n_frames, n_env, n_cent, h_box,
                                       r_cut, respect_pbc & 1, d_workspace, workspace_bytes,
                                       d_counts, d_row_start, d_rows, rows_capacity, d_status,
                                       respect_pbc >> 1, st);
    return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                     pos_dtype);
}

extern "C" int dsg_rows_csr_workspace_bytes(int n_frames, int n_cent, size_t* bytes_out) {
    if (!bytes_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "bytes_out is NULL");
    if (n_frames < 1 || n_cent < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d, n_cent=%d must be >= 1", n_frames,
                         n_cent);
    const size_t tiles = (size_t)n_frames * (((size_t)n_cent + kScanTile - 1) / kScanTile);
    *bytes_out = align_up(4 * tiles) + align_up(8 * (size_t)n_frames);
    return DSG_OK;
}

extern "C" int dsg_rows_csr_offsets(const int32_t* d_counts, int n_frames, int n_cent,
                                    void* d_workspace, size_t workspace_bytes, int32_t* d_indptr,
                                    int64_t* d_frame_off, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    size_t need = 0;
    int rc = dsg_rows_csr_workspace_bytes(n_frames, n_cent, &need);
    if (rc) return rc;
    if (!d_counts || !d_workspace || !d_indptr || !d_frame_off)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (workspace_bytes < need)
        return set_error(DSG_ERR_WORKSPACE_TOO_SMALL, "workspace has %zu bytes, %zu needed",
                         workspace_bytes, need);
    const size_t tiles = (size_t)n_frames * (((size_t)n_cent + kScanTile - 1) / kScanTile);
    int* tile_tmp = (int*)d_workspace;
    long long* seg_totals = (long long*)((char*)d_workspace + align_up(4 * tiles));
    rc = run_scan(d_counts, n_cent, n_frames, d_indptr, tile_tmp, seg_totals, st);
    if (rc) return rc;
    frame_offsets_kernel<<<1, 32, 0, st>>>(seg_totals, n_frames, (long long*)d_frame_off);
    DSG_LAUNCH_CHECK("frame_offsets_kernel");
    return DSG_OK;
}

extern "C" int dsg_rows_to_csr(const int32_t* d_counts, const int64_t* d_row_start,
                               const int32_t* d_rows, const int32_t* d_indptr,
                               const int64_t* d_frame_off, int n_frames, int n_cent,
                               int32_t* d_indices, int expected_row_len, void* stream) {
    if (!d_counts || !d_row_start || !d_rows || !d_indptr || !d_frame_off || !d_indices)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL device pointer argument");
    if (n_frames < 1 || n_cent < 1)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "n_frames=%d, n_cent=%d must be >= 1", n_frames,
                         n_cent);
    const long long warps = (long long)n_frames * n_cent;
    if (expected_row_len > 0 && expected_row_len <= 24) {   // short rows: one thread per row
        const unsigned blocks = (unsigned)((warps + kRtThreads - 1) / kRtThreads);
        rows_to_csr_thread_kernel<<<blocks, kRtThreads, 0, (cudaStream_t)stream>>>(
            d_counts, (const long long*)d_row_start, d_rows, d_indptr,
            (const long long*)d_frame_off, n_frames, n_cent, d_indices);
        DSG_LAUNCH_CHECK("rows_to_csr_thread_kernel");
        return DSG_OK;
    }
    const unsigned blocks = (unsigned)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
    rows_to_csr_kernel<<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        d_counts, (const long long*)d_row_start, d_rows, d_indptr, (const long long*)d_frame_off,
        n_frames, n_cent, d_indices);
    DSG_LAUNCH_CHECK("rows_to_csr_kernel");
    return DSG_OK;
}

extern "C" int dsg_lens_direct_supported(int n_frames, int n_atoms, const double* h_box,
                                         double r_cut, int respect_pbc) {
    PairLayout P;
    std::vector<FrameParams> fps;
    int rc = pair_layout(n_frames, n_atoms, h_box, r_cut, &P, &fps);
    if (rc) return rc;
    return pair_path_supported(fps, n_atoms, r_cut, respect_pbc) ? 1 : 0;
}

extern "C" int dsg_lens_direct_workspace_bytes(int n_frames, int n_atoms, const double* h_box,
                                               double r_cut, size_t* bytes_out) {
    if (!bytes_out) return set_error(DSG_ERR_INVALID_ARGUMENT, "bytes_out is NULL");
    PairLayout P;
    int rc = pair_layout(n_frames, n_atoms, h_box, r_cut, &P, nullptr);
    if (rc) return rc;
    *bytes_out = P.bytes;
    return DSG_OK;
}

extern "C" int dsg_lens_direct(const void* d_pos, int pos_dtype, int n_frames, int n_atoms,
                               const double* h_box, double r_cut, int respect_pbc, int delay,
                               void* d_workspace, size_t workspace_bytes, int32_t* d_counts,
                               double* d_lens, int64_t lens_stride, int64_t col0, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (pos_dtype == DSG_F32)
        return lens_direct_impl<float>(d_pos, n_frames, n_atoms, h_box, r_cut, respect_pbc != 0,
                                       delay, d_workspace, workspace_bytes, d_counts, d_lens,
                                       lens_stride, col0, st);
    if (pos_dtype == DSG_F64)
        return lens_direct_impl<double>(d_pos, n_frames, n_atoms, h_box, r_cut, respect_pbc != 0,
                                        delay, d_workspace, workspace_bytes, d_counts, d_lens,
                                        lens_stride, col0, st);
    return set_error(DSG_ERR_INVALID_ARGUMENT, "pos_dtype=%d is neither DSG_F32 nor DSG_F64",
                     pos_dtype);
}
