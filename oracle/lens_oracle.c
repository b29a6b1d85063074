/*
 * oracle/lens_oracle.c -- CPU restatement of the reference's numba kernels for the
 * neighbour-list / LENS / coordination-number path.
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded (ctypes) by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product (dynsight_b200) never links,
 * imports or calls anything in oracle/.
 *
 * Each function cites the reference lines it follows
 * (/root/reference/src/dynsight/_internal/lens/lens.py):
 *   dsgo_pbc_diff            <- _pbc_diff                       lens.py:17-26
 *   build_cell_list          <- build_cell_list                 lens.py:29-64
 *   dsgo_neighbors           <- neighbor_list_celllist_centers  lens.py:69-147
 *   dsgo_lens_from_two_csr   <- lens_from_two_csr               lens.py:150-189
 *
 * PARITY PIN: checked bit-exactly against the reference's golden files
 * (tests/trajectory/files/{n_c,lens}.npy, tests/systems/lens_2d_{pbc,fbc}.npy,
 * tests/lens/test_lens/c0-c5, tests/neighbors/test_nn/c0-c4) and against the numba kernels
 * themselves on randomised inputs (tests/golden/make_golden.py, tests/test_oracle_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp; NO -ffast-math so the arithmetic is the plain
 * IEEE evaluation of the reference's expressions).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Python's % on ints: result has the sign of the divisor. */
static inline int py_mod(long long a, int n) {
    long long r = a % n;
    if (r < 0) r += n;
    return (int)r;
}

/* Python/numba float floor division a // b (exact floor of the true quotient, via fmod),
 * used by lens.py:43-45 ``int(box[k] // cell_size)``. */
static inline double py_floordiv(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
    if (div == 0.0) return copysign(0.0, a / b);
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
}

/* lens.py:52-54 -- int(x / box * ncell) % ncell : truncation toward zero, then Python modulo. */
static inline int cell_of(double x, double box, int ncell) {
    double v = x / box * (double)ncell;
    return py_mod((long long)v, ncell);
}

/* lens.py:17-26 */
static inline void dsgo_pbc_diff(double* dx, const double* box) {
    for (int k = 0; k < 3; ++k) {
        if (box[k] > 0.0) dx[k] -= rint(dx[k] / box[k]) * box[k];
    }
}

typedef struct {
    int nx, ny, nz;
    int32_t* head;  /* (ncells) */
    int32_t* next;  /* (n_atoms) */
} cell_list_t;

/* lens.py:29-64 */
static int build_cell_list(const double* pos, int n_atoms, const double* box, double cell_size,
                           cell_list_t* cl) {
    int nx = (int)py_floordiv(box[0], cell_size);
    int ny = (int)py_floordiv(box[1], cell_size);
    int nz = (int)py_floordiv(box[2], cell_size);
    if (nx < 3) nx = 3;
    if (ny < 3) ny = 3;
    if (nz < 3) nz = 3;
    long long n_cells = (long long)nx * ny * nz;
    cl->nx = nx; cl->ny = ny; cl->nz = nz;
    cl->head = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_cells);
    cl->next = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_atoms > 0 ? n_atoms : 1));
    if (!cl->head || !cl->next) return -1;
    for (long long c = 0; c < n_cells; ++c) cl->head[c] = -1;
    for (int i = 0; i < n_atoms; ++i) {
        int cx = cell_of(pos[3 * i + 0], box[0], nx);
        int cy = cell_of(pos[3 * i + 1], box[1], ny);
        int cz = cell_of(pos[3 * i + 2], box[2], nz);
        long long c = (long long)cx * ny * nz + (long long)cy * nz + cz;
        cl->next[i] = cl->head[c];
        cl->head[c] = i;
    }
    return 0;
}

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

/* One traversal of the 27 neighbouring cells of centre i (lens.py:86-106 and :117-139).
 * If out != NULL the accepted env indices are appended to it.  Returns the count. */
static inline int visit_centre(const double* pos_env, const double* pos_cent, int i,
                               const double* box, int respect_pbc, double r_cut2,
                               const cell_list_t* cl, int32_t* out) {
    const int nx = cl->nx, ny = cl->ny, nz = cl->nz;
    int cx = cell_of(pos_cent[3 * i + 0], box[0], nx);
    int cy = cell_of(pos_cent[3 * i + 1], box[1], ny);
    int cz = cell_of(pos_cent[3 * i + 2], box[2], nz);
    int count = 0;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                int nx_ = py_mod(cx + dx, nx);
                int ny_ = py_mod(cy + dy, ny);
                int nz_ = py_mod(cz + dz, nz);
                long long cidx = (long long)nx_ * ny * nz + (long long)ny_ * nz + nz_;
                int j = cl->head[cidx];
                while (j != -1) {
                    double dr[3];
                    dr[0] = pos_env[3 * j + 0] - pos_cent[3 * i + 0];
                    dr[1] = pos_env[3 * j + 1] - pos_cent[3 * i + 1];
                    dr[2] = pos_env[3 * j + 2] - pos_cent[3 * i + 2];
                    if (respect_pbc) dsgo_pbc_diff(dr, box);
                    double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
                    if (j != i && dr2 < r_cut2) {
                        if (out) out[count] = j;
                        ++count;
                    }
                    j = cl->next[j];
                }
            }
    return count;
}

/*
 * lens.py:69-147.  pos_env (n_env,3) f64, pos_cent (n_cent,3) f64, box (3,) f64.
 * indptr_out has n_cent+1 entries.  *indices_out is malloc'ed here (free with dsgo_free).
 * Returns nnz (>= 0) or -1 on allocation failure.
 */
long long dsgo_neighbors(const double* pos_env, int n_env, const double* pos_cent, int n_cent,
                         double r_cut, const double* box, int respect_pbc, int32_t* indptr_out,
                         int32_t** indices_out) {
    const double r_cut2 = (r_cut - 1e-6) * (r_cut - 1e-6);
    cell_list_t cl;
    if (build_cell_list(pos_env, n_env, box, r_cut, &cl) != 0) return -1;

    int32_t* n_neigh = (int32_t*)calloc((size_t)(n_cent > 0 ? n_cent : 1), sizeof(int32_t));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_cent; ++i)
        n_neigh[i] = visit_centre(pos_env, pos_cent, i, box, respect_pbc, r_cut2, &cl, NULL);

    indptr_out[0] = 0;
    for (int i = 0; i < n_cent; ++i) indptr_out[i + 1] = indptr_out[i] + n_neigh[i];
    long long nnz = indptr_out[n_cent];
    int32_t* indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    if (!indices) { free(n_neigh); free(cl.head); free(cl.next); return -1; }

#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_cent; ++i)
        visit_centre(pos_env, pos_cent, i, box, respect_pbc, r_cut2, &cl,
                     indices + indptr_out[i]);

    /* lens.py:142-145 : serial per-row ascending sort */
    for (int i = 0; i < n_cent; ++i) {
        int s = indptr_out[i], e = indptr_out[i + 1];
        if (e > s + 1) qsort(indices + s, (size_t)(e - s), sizeof(int32_t), cmp_i32);
    }
    free(n_neigh); free(cl.head); free(cl.next);
    *indices_out = indices;
    return nnz;
}

void dsgo_free(void* p) { free(p); }

/* lens.py:150-189 */
void dsgo_lens_from_two_csr(const int32_t* indptr1, const int32_t* indices1,
                            const int32_t* indptr2, const int32_t* indices2, int n_centers,
                            double* out) {
    for (int u = 0; u < n_centers; ++u) {
        int s1 = indptr1[u], e1 = indptr1[u + 1];
        int s2 = indptr2[u], e2 = indptr2[u + 1];
        int a = e1 - s1, b = e2 - s2;
        int denom = a + b;
        if (denom <= 0) { out[u] = 0.0; continue; }
        int i = s1, j = s2, inter = 0;
        while (i < e1 && j < e2) {
            int32_t vi = indices1[i], vj = indices2[j];
            if (vi == vj) { ++inter; ++i; ++j; }
            else if (vi < vj) ++i;
            else ++j;
        }
        int numer = (a + b) - 2 * inter;
        out[u] = (double)numer / (double)denom;
    }
}

void dsgo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

int dsgo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
