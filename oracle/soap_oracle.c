/*
 * oracle/soap_oracle.c -- C/OpenMP restatement of the DScribe SOAP-GTO power spectrum, used as
 * the timed CPU baseline for SOAP (DScribe itself is not installed on either box) and as a fast
 * checker at sizes where the numpy restatement (oracle/soap_oracle.py) is too slow.
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py.  Never linked or called by dynsight_b200.
 *
 * Algorithm: see the header of oracle/soap_oracle.py (reference call site
 * src/dynsight/_internal/soapify/saponify.py:98-122; DScribe get_basis_gto + soapGTO.cpp).
 * Differences from the numpy version are organisational only: a cell list (edge >= search radius,
 * like DScribe's CellList) finds the neighbours when the box is >= 3 radii wide, otherwise every
 * image is enumerated; alphas / betas come from the caller (oracle/soap_oracle.py:gto_basis).
 *
 * PARITY PIN: compared with the numpy restatement (itself pinned to the reference's DScribe
 * goldens) in tests/test_oracle_golden.py::test_soap_c_oracle_matches_numpy_oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846
#define LMAX_CAP 16

typedef struct {
    int n;
    int cap;
    double* d; /* 3 doubles per neighbour */
    int* sp;
} nlist_t;

static void nl_push(nlist_t* nl, double dx, double dy, double dz, int sp) {
    if (nl->n == nl->cap) {
        nl->cap = nl->cap ? nl->cap * 2 : 256;
        nl->d = (double*)realloc(nl->d, sizeof(double) * 3 * (size_t)nl->cap);
        nl->sp = (int*)realloc(nl->sp, sizeof(int) * (size_t)nl->cap);
    }
    nl->d[3 * nl->n] = dx; nl->d[3 * nl->n + 1] = dy; nl->d[3 * nl->n + 2] = dz;
    nl->sp[nl->n] = sp;
    nl->n++;
}

/* Orthonormal real spherical harmonics Y_lm(d/|d|), l <= l_max, stored at y[l*l + l + m]. */
static void real_ylm(int l_max, double x, double y_, double z, double* y) {
    double am[LMAX_CAP + 1], bm[LMAX_CAP + 1];
    am[0] = 1.0; bm[0] = 0.0;
    for (int m = 1; m <= l_max; ++m) {
        am[m] = am[m - 1] * x - bm[m - 1] * y_;
        bm[m] = am[m - 1] * y_ + bm[m - 1] * x;
    }
    for (int m = 0; m <= l_max; ++m) {
        double q[LMAX_CAP + 1];
        double dfact = 1.0;
        for (int i = 1; i < 2 * m; i += 2) dfact *= (double)i;
        q[m] = dfact;
        if (m + 1 <= l_max) q[m + 1] = (2 * m + 1) * z * q[m];
        for (int l = m + 2; l <= l_max; ++l)
            q[l] = ((2 * l - 1) * z * q[l - 1] - (l + m - 1) * q[l - 2]) / (double)(l - m);
        for (int l = m; l <= l_max; ++l) {
            double ratio = 1.0; /* (l-m)!/(l+m)! */
            for (int i = l - m + 1; i <= l + m; ++i) ratio /= (double)i;
            double norm = sqrt((2 * l + 1) / (4.0 * PI) * ratio);
            if (m == 0) y[l * l + l] = norm * q[l];
            else {
                y[l * l + l + m] = sqrt(2.0) * norm * q[l] * am[m];
                y[l * l + l - m] = sqrt(2.0) * norm * q[l] * bm[m];
            }
        }
    }
}

static inline double wrap(double x, double b) {
    double w = x - floor(x / b) * b;
    if (w >= b) w -= b;
    if (w < 0) w = 0;
    return w;
}

/*
 * pos (n_atoms,3) f64; species (n_atoms) index in [0,n_species); centres (n_cent) atom indices;
 * cell (3) or NULL (non periodic); alphas (l_max+1, n_max); betas (l_max+1, n_max, n_max);
 * out (n_cent, n_feat).  Returns 0, or -1 on bad arguments.
 */
int dsgo_soap_frame(const double* pos, const int32_t* species, int n_atoms,
                    const int32_t* centres, int n_cent, int n_species, const double* cell,
                    double r_cut, int n_max, int l_max, const double* alphas, const double* betas,
                    double* out) {
    if (l_max > LMAX_CAP || n_max > 64 || n_species > 16) return -1;
    const double eta = 0.5;
    const double radius = r_cut + sqrt(-2.0 * log(1e-3));
    const double r2max = radius * radius;
    const int n_lm = (l_max + 1) * (l_max + 1);
    const int sn = n_species * n_max;
    const int n_feat = sn * (sn + 1) / 2 * (l_max + 1);

    /* optional cell list on wrapped coordinates */
    int use_cells = 0, nc[3] = {1, 1, 1};
    if (cell) {
        use_cells = 1;
        for (int k = 0; k < 3; ++k) {
            nc[k] = (int)floor(cell[k] / (radius * (1.0 + 1e-9)));
            if (nc[k] < 3) use_cells = 0;
        }
    }
    double* wpos = NULL;
    int *head = NULL, *next = NULL;
    if (use_cells) {
        wpos = (double*)malloc(sizeof(double) * 3 * (size_t)n_atoms);
        head = (int*)malloc(sizeof(int) * (size_t)nc[0] * nc[1] * nc[2]);
        next = (int*)malloc(sizeof(int) * (size_t)n_atoms);
        for (int c = 0; c < nc[0] * nc[1] * nc[2]; ++c) head[c] = -1;
        for (int i = 0; i < n_atoms; ++i) {
            int ci[3];
            for (int k = 0; k < 3; ++k) {
                wpos[3 * i + k] = wrap(pos[3 * i + k], cell[k]);
                ci[k] = (int)(wpos[3 * i + k] / cell[k] * nc[k]);
                if (ci[k] >= nc[k]) ci[k] = nc[k] - 1;
            }
            int c = (ci[0] * nc[1] + ci[1]) * nc[2] + ci[2];
            next[i] = head[c];
            head[c] = i;
        }
    }
    /* per-(l,k) constants of the primitive radial functions (DScribe precomputes them too) */
    double* pref = (double*)malloc(sizeof(double) * (size_t)(l_max + 1) * n_max);
    double* gam = (double*)malloc(sizeof(double) * (size_t)(l_max + 1) * n_max);
    for (int l = 0; l <= l_max; ++l)
        for (int k = 0; k < n_max; ++k) {
            const double al = alphas[l * n_max + k];
            pref[l * n_max + k] = pow(PI, 1.5) * pow(eta / (eta + al), l) * pow(eta + al, -1.5);
            gam[l * n_max + k] = al * eta / (al + eta);
        }
    int reps[3] = {0, 0, 0};
    if (cell && !use_cells)
        for (int k = 0; k < 3; ++k) reps[k] = (int)ceil(radius / cell[k]) + 1;

#pragma omp parallel
    {
        nlist_t nl = {0, 0, NULL, NULL};
        double* prim = (double*)malloc(sizeof(double) * (size_t)n_species * n_max * n_lm);
        double* cc = (double*)malloc(sizeof(double) * (size_t)n_species * n_max * n_lm);
        double* ylm = (double*)malloc(sizeof(double) * (size_t)n_lm);
#pragma omp for schedule(dynamic, 8)
        for (int ic = 0; ic < n_cent; ++ic) {
            const int a = centres[ic];
            nl.n = 0;
            if (use_cells) {
                const double* pc = wpos + 3 * a;
                int ci[3];
                for (int k = 0; k < 3; ++k) {
                    ci[k] = (int)(pc[k] / cell[k] * nc[k]);
                    if (ci[k] >= nc[k]) ci[k] = nc[k] - 1;
                }
                for (int ox = -1; ox <= 1; ++ox)
                    for (int oy = -1; oy <= 1; ++oy)
                        for (int oz = -1; oz <= 1; ++oz) {
                            int o[3] = {ox, oy, oz}, cj[3];
                            double sh[3];
                            for (int k = 0; k < 3; ++k) {
                                cj[k] = ci[k] + o[k];
                                sh[k] = 0.0;
                                if (cj[k] < 0) { cj[k] += nc[k]; sh[k] = -cell[k]; }
                                else if (cj[k] >= nc[k]) { cj[k] -= nc[k]; sh[k] = cell[k]; }
                            }
                            int j = head[(cj[0] * nc[1] + cj[1]) * nc[2] + cj[2]];
                            while (j != -1) {
                                double dx = wpos[3 * j] + sh[0] - pc[0];
                                double dy = wpos[3 * j + 1] + sh[1] - pc[1];
                                double dz = wpos[3 * j + 2] + sh[2] - pc[2];
                                if (dx * dx + dy * dy + dz * dz <= r2max)
                                    nl_push(&nl, dx, dy, dz, species[j]);
                                j = next[j];
                            }
                        }
            } else {
                const double* pc = pos + 3 * a;
                for (int sx = -reps[0]; sx <= reps[0]; ++sx)
                    for (int sy = -reps[1]; sy <= reps[1]; ++sy)
                        for (int sz = -reps[2]; sz <= reps[2]; ++sz)
                            for (int j = 0; j < n_atoms; ++j) {
                                double dx = pos[3 * j] - pc[0] + (cell ? sx * cell[0] : 0.0);
                                double dy = pos[3 * j + 1] - pc[1] + (cell ? sy * cell[1] : 0.0);
                                double dz = pos[3 * j + 2] - pc[2] + (cell ? sz * cell[2] : 0.0);
                                if (dx * dx + dy * dy + dz * dz <= r2max)
                                    nl_push(&nl, dx, dy, dz, species[j]);
                            }
            }
            /* primitive sums G^Z_{l k m} */
            memset(prim, 0, sizeof(double) * (size_t)n_species * n_max * n_lm);
            for (int j = 0; j < nl.n; ++j) {
                const double dx = nl.d[3 * j], dy = nl.d[3 * j + 1], dz = nl.d[3 * j + 2];
                const double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2);
                if (r > 0.0) real_ylm(l_max, dx / r, dy / r, dz / r, ylm);
                else real_ylm(l_max, 0.0, 0.0, 1.0, ylm);
                double* pz = prim + (size_t)nl.sp[j] * n_max * n_lm;
                double rl = 1.0;
                for (int l = 0; l <= l_max; ++l) {
                    if (l > 0) rl *= r; /* r^l; for r == 0 only l = 0 survives */
                    for (int k = 0; k < n_max; ++k) {
                        const double radial =
                            pref[l * n_max + k] * rl * exp(-gam[l * n_max + k] * r2);
                        double* dst = pz + (size_t)k * n_lm + l * l;
                        for (int mi = 0; mi <= 2 * l; ++mi) dst[mi] += radial * ylm[l * l + mi];
                    }
                }
            }
            /* c = beta * G */
            for (int z = 0; z < n_species; ++z)
                for (int l = 0; l <= l_max; ++l)
                    for (int n = 0; n < n_max; ++n)
                        for (int mi = 0; mi <= 2 * l; ++mi) {
                            double s = 0.0;
                            for (int k = 0; k < n_max; ++k)
                                s += betas[((size_t)l * n_max + n) * n_max + k] *
                                     prim[((size_t)z * n_max + k) * n_lm + l * l + mi];
                            cc[((size_t)z * n_max + n) * n_lm + l * l + mi] = s;
                        }
            /* power spectrum */
            double* o = out + (size_t)ic * n_feat;
            int col = 0;
            for (int zi = 0; zi < n_species; ++zi)
                for (int zj = zi; zj < n_species; ++zj)
                    for (int l = 0; l <= l_max; ++l) {
                        const double pref = PI * sqrt(8.0 / (2 * l + 1));
                        for (int n = 0; n < n_max; ++n)
                            for (int n2 = (zi == zj ? n : 0); n2 < n_max; ++n2) {
                                const double* ca = cc + ((size_t)zi * n_max + n) * n_lm + l * l;
                                const double* cb = cc + ((size_t)zj * n_max + n2) * n_lm + l * l;
                                double s = 0.0;
                                for (int mi = 0; mi <= 2 * l; ++mi) s += ca[mi] * cb[mi];
                                o[col++] = pref * s;
                            }
                    }
        }
        free(nl.d); free(nl.sp); free(prim); free(cc); free(ylm);
    }
    free(wpos); free(head); free(next); free(pref); free(gam);
    return 0;
}

void dsgo_soap_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}
