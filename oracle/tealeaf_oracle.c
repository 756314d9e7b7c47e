/*
 * tealeaf_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see tealeaf_oracle.h).
 *
 * Restates, in plain C, the arithmetic of the reference TeaLeaf kernels and the
 * call order of its drivers. Every function cites the reference file:line it
 * follows (paths relative to /root/reference/TeaLeaf/).  Build with
 * -ffp-contract=off so that each operation is rounded exactly as written.
 */
#include "tealeaf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define IDX(jj, kk) ((long)(kk) + (long)(jj) * x)
#define INTERIOR(jj, kk) ((kk) >= hd && (kk) < x - hd && (jj) >= hd && (jj) < y - hd)

/* shared.h:59-63 -- the association of the reference macro, verbatim in meaning */
#define SMVP(a, i)                                                          \
    ((1.0 + (kx[(i) + 1] + kx[(i)]) + (ky[(i) + x] + ky[(i)])) * a[(i)]     \
     - (kx[(i) + 1] * a[(i) + 1] + kx[(i)] * a[(i)-1])                      \
     - (ky[(i) + x] * a[(i) + x] + ky[(i)] * a[(i)-x]))

/* ------------------------------------------------------------------------- */
/* Reductions: sycl_shared.hpp:35-79 + the in-kernel tree of e.g. cg.cpp:113-127.
 * 64 consecutive flat indices are summed by an adjacent-pair binary tree
 * (stride 1,2,4,...,32); the per-group results are reduced the same way,
 * repeatedly, until one value remains. Missing tail elements count as 0.0. */
static inline double tree64(double* l)
{
    for (int s = 1; s < 64; s *= 2)
        for (int i = 0; i < 64; i += 2 * s) l[i] = l[i] + l[i + s];
    return l[0];
}

double orc_tree_sum(const double* v, long n)
{
    if (n <= 0) return 0.0;
    if (n == 1) return v[0];
    long ng = (n + 63) / 64;
    double* cur = (double*)malloc(sizeof(double) * (size_t)ng);
    long len = n;
    const double* src = v;
    for (;;) {
        ng = (len + 63) / 64;
        for (long g = 0; g < ng; ++g) {
            double loc[64];
            for (int l = 0; l < 64; ++l) {
                long i = g * 64 + l;
                loc[l] = (i < len) ? src[i] : 0.0;
            }
            cur[g] = tree64(loc); /* in place is safe: group g only reads indices >= g */
        }
        len = ng;
        src = cur;
        if (len == 1) break;
    }
    double out = cur[0];
    free(cur);
    return out;
}

/* Grow-only scratch for the per-work-group partials (the reference's tmpArrayBuff). */
static double* g_partials = NULL;
static long g_partials_cap = 0;
static double* partials(long ng)
{
    if (ng > g_partials_cap) {
        free(g_partials);
        g_partials = (double*)malloc(sizeof(double) * (size_t)ng);
        g_partials_cap = ng;
    }
    return g_partials;
}

/* ------------------------------------------------------------------------- */
/* Optional summation order: replay of the CUDA backend's (deterministic) reduction tree, so that a
 * whole solve can be compared with the GPU bit for bit (iteration counts included).  Mode 0 (default)
 * is the reference's order above; mode 1 replays exploringsycl_b200/csrc/tl_kernels.cu:
 *   thread: sequential over its rows (and its 1 or 2 columns) -> warp xor-butterfly -> 4 warps
 *   (s0+s1)+(s2+s3) = tile partial; 64 tiles per group (lanes of two warps, s0+s1); groups summed
 *   by 128 threads striding, butterfly, (s0+s1)+(s2+s3).                                         */
static int g_sum_mode = 0;
void orc_set_sum_mode(int mode) { g_sum_mode = mode; }

static double butterfly32(double* v) /* lane 0's value after xor-butterfly 1,2,4,8,16 */
{
    for (int o = 1; o < 32; o <<= 1)
        for (int l = 0; l < 32; l += 2 * o) v[l] = v[l] + v[l + o];
    return v[0];
}
static double block128(double* t) /* 128 per-thread values -> CTA value */
{
    double w0 = butterfly32(t), w1 = butterfly32(t + 32), w2 = butterfly32(t + 64), w3 = butterfly32(t + 96);
    return (w0 + w1) + (w2 + w3);
}
enum { SUM_GENERIC = 0, SUM_HOT_W = 1, SUM_HOT_UR = 2, SUM_GENERIC_ALL = 3, SUM_HOT_FS = 4 };
static int gpu_tile_rows(int kind, int nx, int ny)
{
    if (kind != SUM_HOT_W && kind != SUM_HOT_FS) return 8;
    /* tall tiles of the stencil kernels: about 6.75 tiles per SM (tall_tile_rows in tl_kernels.cu) */
    int colb = (nx + 255) / 256;
    int rowblocks = 999 / colb;
    if (rowblocks < 1) rowblocks = 1;
    int rows = (ny + rowblocks - 1) / rowblocks;
    if (rows < 8) rows = 8;
    if (rows > 128) rows = 128;
    return rows;
}
static double gpu_order_sum(const double* val, int x, int y, int hd, int kind)
{
    /* SUM_GENERIC_ALL: the one-pass cg_init kernel runs over ALL cells (tiles start at cell (0,0); the cells outside
     * the interior contribute 0.0); the other kinds tile the interior. */
    const int o = (kind == SUM_GENERIC_ALL) ? 0 : hd;
    const int nx = x - 2 * o, ny = y - 2 * o;
    const int cpt = (kind == SUM_GENERIC || kind == SUM_GENERIC_ALL) ? 1 : 2, tile_cols = 128 * cpt;
    const int rows = gpu_tile_rows(kind, nx, ny);
    const int gx = (nx + tile_cols - 1) / tile_cols, gy = (ny + rows - 1) / rows;
    const long ntiles = (long)gx * gy;
    double* part = (double*)malloc(sizeof(double) * (size_t)ntiles);
#pragma omp parallel for schedule(static)
    for (long tile = 0; tile < ntiles; ++tile) {
        const int bx = (int)(tile % gx), by = (int)(tile / gx);
        const int j0 = o + by * rows, j1 = (j0 + rows < y - o) ? j0 + rows : y - o;
        double t[128];
        for (int tx = 0; tx < 128; ++tx) {
            double acc = 0.0;
            const int kk = o + cpt * (bx * 128 + tx);
            for (int jj = j0; jj < j1; ++jj)
                for (int c = 0; c < cpt; ++c)
                    if (kk + c < x - o) acc += val[(long)(kk + c) + (long)jj * x];
            t[tx] = acc;
        }
        part[tile] = block128(t);
    }
    const long ngroups = (ntiles + 63) / 64;
    double* gp = (double*)malloc(sizeof(double) * (size_t)ngroups);
    for (long g = 0; g < ngroups; ++g) {
        double t[64];
        for (int l = 0; l < 64; ++l) t[l] = (g * 64 + l < ntiles) ? part[g * 64 + l] : 0.0;
        double w0 = butterfly32(t), w1 = butterfly32(t + 32);
        gp[g] = w0 + w1;
    }
    double t[128];
    for (int tid = 0; tid < 128; ++tid) {
        double s = 0.0;
        for (long k = tid; k < ngroups; k += 128) s += gp[k];
        t[tid] = s;
    }
    const double total = block128(t);
    free(part);
    free(gp);
    return total;
}

/* Generic "nd64" reduction kernel shape: range ceil(x*y/64)*64, one work-item per
 * flat index, interior test inside, value 0 elsewhere (e.g. cg.cpp:96-127). */
#define ND64_REDUCE(RESULT, CELL_EXPR) ND64_REDUCE_K(RESULT, CELL_EXPR, SUM_GENERIC)
#define ND64_REDUCE_K(RESULT, CELL_EXPR, KIND)                                      \
    do {                                                                            \
        const long n_ = (long)x * y;                                                \
        if (g_sum_mode == 1) {                                                      \
            double* val_ = (double*)calloc((size_t)n_, sizeof(double));             \
            _Pragma("omp parallel for schedule(static)")                            \
            for (int jj = hd; jj < y - hd; ++jj)                                    \
                for (int kk = hd; kk < x - hd; ++kk) {                              \
                    const long index = (long)kk + (long)jj * x;                     \
                    val_[index] = (CELL_EXPR);                                      \
                }                                                                   \
            (RESULT) = gpu_order_sum(val_, x, y, hd, KIND);                         \
            free(val_);                                                             \
            break;                                                                  \
        }                                                                           \
        const long ng_ = (n_ + 63) / 64;                                            \
        double* part_ = partials(ng_);                                              \
        _Pragma("omp parallel for schedule(static)")                                \
        for (long g_ = 0; g_ < ng_; ++g_) {                                         \
            double loc_[64];                                                        \
            long base_ = g_ * 64;                                                   \
            int jj = (int)(base_ / x), kk = (int)(base_ % x);                       \
            for (int l_ = 0; l_ < 64; ++l_) {                                       \
                const long index = base_ + l_;                                      \
                double val_ = 0.0;                                                  \
                if (index < n_ && INTERIOR(jj, kk)) { val_ = (CELL_EXPR); }         \
                loc_[l_] = val_;                                                    \
                if (++kk == x) { kk = 0; ++jj; }                                    \
            }                                                                       \
            part_[g_] = tree64(loc_);                                               \
        }                                                                           \
        (RESULT) = orc_tree_sum(part_, ng_);                                        \
    } while (0)

/* ------------------------------------------------------------------------- */
/* set_chunk_data.cpp:8-64; kernel_interface.cpp:34-52 computes x_min,y_min. */
void orc_set_chunk_data(int x, int y, int hd, double x_min, double y_min, double dx, double dy,
                        double* vertex_x, double* vertex_y, double* cell_x, double* cell_y,
                        double* volume)
{
    for (int i = 0; i < x + 1; ++i) vertex_x[i] = x_min + dx * ((double)i - (double)hd);
    for (int i = 0; i < y + 1; ++i) vertex_y[i] = y_min + dy * ((double)i - (double)hd);
    for (int i = 0; i < x; ++i) cell_x[i] = 0.5 * (vertex_x[i] + vertex_x[i + 1]);
    for (int i = 0; i < y; ++i) cell_y[i] = 0.5 * (vertex_y[i] + vertex_y[i + 1]);
    const long n = (long)x * y;
    const double v = dx * dy;
    for (long i = 0; i < n; ++i) volume[i] = v;
}

/* set_chunk_state.cpp:8-27 */
void orc_set_chunk_initial_state(int x, int y, double energy, double density,
                                 double* energy0, double* density_f)
{
    const long n = (long)x * y;
    for (long i = 0; i < n; ++i) {
        energy0[i] = energy;
        density_f[i] = density;
    }
}

/* set_chunk_state.cpp:30-92. s_* are the ALREADY SHRUNK extents (parse_config.c:253-260). */
void orc_set_chunk_state(int x, int y, int hd, int geometry, double s_density, double s_energy,
                         double s_xmin, double s_ymin, double s_xmax, double s_ymax, double s_radius,
                         double* energy0, double* density, double* u,
                         const double* cell_x, const double* cell_y,
                         const double* vertex_x, const double* vertex_y)
{
    (void)hd;
    for (int jj = 0; jj < y; ++jj) {
        for (int kk = 0; kk < x; ++kk) {
            const long i = IDX(jj, kk);
            int apply = 0;
            if (geometry == ORC_GEOM_RECT) {
                apply = (vertex_x[kk + 1] >= s_xmin && vertex_x[kk] < s_xmax &&
                         vertex_y[jj + 1] >= s_ymin && vertex_y[jj] < s_ymax);
            } else if (geometry == ORC_GEOM_CIRC) {
                double radius = sqrt((cell_x[kk] - s_xmin) * (cell_x[kk] - s_xmin) +
                                     (cell_y[jj] - s_ymin) * (cell_y[jj] - s_ymin));
                apply = (radius <= s_radius);
            } else if (geometry == ORC_GEOM_POINT) {
                apply = (vertex_x[kk] == s_xmin && vertex_y[jj] == s_ymin);
            }
            if (apply) {
                energy0[i] = s_energy;
                density[i] = s_density;
            }
            if (kk > 0 && kk < x - 1 && jj > 0 && jj < y - 1) u[i] = energy0[i] * density[i];
        }
    }
}

/* store_energy.cpp:6-24 */
void orc_store_energy(int x, int y, const double* energy0, double* energy)
{
    const long n = (long)x * y;
    for (long i = 0; i < n; ++i) energy[i] = energy0[i];
}

/* field_summary.cpp:6-151 (uses energy0; results are ASSIGNED, :147-150) */
void orc_field_summary(int x, int y, int hd, const double* volume, const double* density,
                       const double* energy0, const double* u,
                       double* vol, double* mass, double* ie, double* temp)
{
    double r;
    ND64_REDUCE_K(r, volume[index], SUM_HOT_UR);
    *vol = r;
    ND64_REDUCE_K(r, volume[index] * density[index], SUM_HOT_UR);
    *mass = r;
    ND64_REDUCE_K(r, (volume[index] * density[index]) * energy0[index], SUM_HOT_UR);
    *ie = r;
    ND64_REDUCE_K(r, (volume[index] * density[index]) * u[index], SUM_HOT_UR);
    *temp = r;
}

/* ------------------------------------------------------------------------- */
/* local_halos.cpp:7-102 (index maths derived in SURVEY App. A.3) */
void orc_local_halo(int x, int y, int hd, int depth, int face, double* a)
{
    if (face == ORC_LEFT) {
        for (int jj = 0; jj < y; ++jj)
            for (int d = 0; d < depth; ++d) a[IDX(jj, hd - 1 - d)] = a[IDX(jj, hd + d)];
    } else if (face == ORC_RIGHT) {
        for (int jj = 0; jj < y; ++jj)
            for (int d = 0; d < depth; ++d) a[IDX(jj, x - hd + d)] = a[IDX(jj, x - hd - 1 - d)];
    } else if (face == ORC_TOP) {
        for (int d = 0; d < depth; ++d)
            for (int kk = 0; kk < x; ++kk) a[IDX(y - hd + d, kk)] = a[IDX(y - hd - 1 - d, kk)];
    } else if (face == ORC_BOTTOM) {
        for (int d = 0; d < depth; ++d)
            for (int kk = 0; kk < x; ++kk) a[IDX(hd - 1 - d, kk)] = a[IDX(hd + d, kk)];
    }
}

/* pack_halos.cpp:7-94: the literal flat-index formulas */
void orc_pack(int x, int y, int hd, int depth, int face, const double* field, double* buffer)
{
    if (face == ORC_LEFT || face == ORC_RIGHT) {
        const int base = (face == ORC_LEFT) ? hd : x - hd - depth;
        for (int i = 0; i < y * depth; ++i) {
            const int lines = i / depth;
            const long offset = base + (long)lines * (x - depth);
            buffer[i] = field[offset + i];
        }
    } else {
        const long offset = (face == ORC_TOP) ? (long)x * (y - hd - depth) : (long)x * hd;
        for (int i = 0; i < x * depth; ++i) buffer[i] = field[offset + i];
    }
}

/* pack_halos.cpp:97-184 */
void orc_unpack(int x, int y, int hd, int depth, int face, double* field, const double* buffer)
{
    if (face == ORC_LEFT || face == ORC_RIGHT) {
        const int base = (face == ORC_LEFT) ? hd - depth : x - hd;
        for (int i = 0; i < y * depth; ++i) {
            const int lines = i / depth;
            const long offset = base + (long)lines * (x - depth);
            field[offset + i] = buffer[i];
        }
    } else {
        const long offset = (face == ORC_TOP) ? (long)x * (y - hd) : (long)x * (hd - depth);
        for (int i = 0; i < x * depth; ++i) field[offset + i] = buffer[i];
    }
}

/* ------------------------------------------------------------------------- */
/* cg.cpp:7-134 via kernel_interface.cpp:192-210: cg_init_u, cg_init_k, cg_init_others.
 * *rro is ACCUMULATED (cg.cpp:133). */
void orc_cg_init(int x, int y, int hd, int coefficient, double rx, double ry,
                 const double* density, const double* energy, double* u, double* p, double* r,
                 double* w, double* kx, double* ky, double* rro)
{
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < y; ++jj) {
        for (int kk = 0; kk < x; ++kk) {
            const long i = IDX(jj, kk);
            p[i] = 0.0;
            r[i] = 0.0;
            u[i] = energy[i] * density[i];
            if (jj > 0 && jj < y - 1 && kk > 0 && kk < x - 1)
                w[i] = (coefficient == ORC_CONDUCTIVITY) ? density[i] : 1.0 / density[i];
        }
    }
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - 1; ++jj) {
        for (int kk = hd; kk < x - 1; ++kk) {
            const long i = IDX(jj, kk);
            kx[i] = rx * (w[i - 1] + w[i]) / (2.0 * w[i - 1] * w[i]);
            ky[i] = ry * (w[i - x] + w[i]) / (2.0 * w[i - x] * w[i]);
        }
    }
    /* cg_init_others: w=SMVP(u); r=u-w; p=r; partial = r*p */
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj) {
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            const double smvp = SMVP(u, i);
            w[i] = smvp;
            r[i] = u[i] - w[i];
            p[i] = r[i];
        }
    }
    double s;
    ND64_REDUCE_K(s, r[index] * p[index], SUM_GENERIC_ALL);
    *rro += s;
}

/* cg.cpp:137-195: w = A p ; *pw += sum(w*p) */
void orc_cg_calc_w(int x, int y, int hd, const double* p, const double* kx, const double* ky,
                   double* w, double* pw)
{
    double s;
    ND64_REDUCE_K(s, (w[index] = SMVP(p, index), w[index] * p[index]), SUM_HOT_W);
    *pw += s;
}

/* cg.cpp:198-254: u += alpha p ; r -= alpha w ; *rrn = sum(r*r) (ASSIGNED, :253) */
void orc_cg_calc_ur(int x, int y, int hd, double alpha, const double* p, const double* w,
                    double* u, double* r, double* rrn)
{
    double s;
    ND64_REDUCE_K(s, (u[index] += alpha * p[index], r[index] -= alpha * w[index], r[index] * r[index]), SUM_HOT_UR);
    *rrn = s;
}

/* cg.cpp:257-281: p = beta p + r */
void orc_cg_calc_p(int x, int y, int hd, double beta, const double* r, double* p)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            p[i] = beta * p[i] + r[i];
        }
}

/* ------------------------------------------------------------------------- */
/* cheby.cpp:7-43 */
void orc_cheby_init(int x, int y, int hd, double theta, const double* u, const double* u0,
                    const double* kx, const double* ky, double* p, double* r, double* w)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            const double smvp = SMVP(u, i);
            w[i] = smvp;
            r[i] = u0[i] - w[i];
            p[i] = r[i] / theta;
        }
}

/* kernel_interface.cpp:258-271: cheby_iterate (cheby.cpp:73-110) then cheby_calc_u (:46-70) */
void orc_cheby_iterate(int x, int y, int hd, double alpha, double beta, double* u, const double* u0,
                       const double* kx, const double* ky, double* p, double* r, double* w)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            const double smvp = SMVP(u, i);
            w[i] = smvp;
            r[i] = u0[i] - w[i];
            p[i] = alpha * p[i] + beta * r[i];
        }
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            u[i] += p[i];
        }
}

/* ppcg.cpp:7-31 */
void orc_ppcg_init(int x, int y, int hd, double theta, const double* r, double* sd)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            sd[i] = r[i] / theta;
        }
}

/* kernel_interface.cpp:314-328: ppcg_calc_ur (ppcg.cpp:34-66) then ppcg_calc_sd (:69-94) */
void orc_ppcg_inner_iteration(int x, int y, int hd, double alpha, double beta, double* u, double* r,
                              const double* kx, const double* ky, double* sd)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            const double smvp = SMVP(sd, i);
            r[i] -= smvp;
            u[i] += sd[i];
        }
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            sd[i] = alpha * sd[i] + beta * r[i];
        }
}

/* jacobi.cpp:7-54 */
void orc_jacobi_init(int x, int y, int hd, int coefficient, double rx, double ry,
                     const double* density, const double* energy, double* u0, double* u,
                     double* kx, double* ky)
{
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < y; ++jj)
        for (int kk = 0; kk < x; ++kk) {
            const long i = IDX(jj, kk);
            if (kk > 0 && kk < x - 1 && jj > 0 && jj < y - 1) {
                u0[i] = energy[i] * density[i];
                u[i] = u0[i];
            }
            if (jj >= hd && jj < y - 1 && kk >= hd && kk < x - 1) {
                const int c = (coefficient == ORC_CONDUCTIVITY);
                double dc = c ? density[i] : 1.0 / density[i];
                double dl = c ? density[i - 1] : 1.0 / density[i - 1];
                double dd = c ? density[i - x] : 1.0 / density[i - x];
                kx[i] = rx * (dl + dc) / (2.0 * dl * dc);
                ky[i] = ry * (dd + dc) / (2.0 * dd * dc);
            }
        }
}

/* kernel_interface.cpp:287-300: jacobi_copy_u (jacobi.cpp:120-137) then jacobi_iterate
 * (:57-117); *error is ASSIGNED (:116) */
void orc_jacobi_iterate(int x, int y, int hd, double* u, const double* u0, double* r,
                        const double* kx, const double* ky, double* error)
{
    const long n = (long)x * y;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) r[i] = u[i];
    double s;
    ND64_REDUCE(s, (u[index] = (u0[index]
                                + (kx[index + 1] * r[index + 1] + kx[index] * r[index - 1])
                                + (ky[index + x] * r[index + x] + ky[index] * r[index - x]))
                               / (1.0 + (kx[index] + kx[index + 1]) + (ky[index] + ky[index + x])),
                    fabs(u[index] - r[index])));
    *error = s;
}

/* solver_methods.cpp:7-31 */
void orc_copy_u(int x, int y, int hd, const double* u, double* u0)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) u0[IDX(jj, kk)] = u[IDX(jj, kk)];
}

/* solver_methods.cpp:34-65 */
void orc_calculate_residual(int x, int y, int hd, const double* u, const double* u0,
                            const double* kx, const double* ky, double* r)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            const double smvp = SMVP(u, i);
            r[i] = u0[i] - smvp;
        }
}

/* solver_methods.cpp:68-117; *norm ASSIGNED (:116) */
/* kind: tile geometry of the GPU kernel that computes this sum (only the replay mode looks at it) */
static void calculate_2norm_k(int x, int y, int hd, const double* buffer, double* norm, int kind)
{
    double s;
    if (kind == SUM_HOT_FS) ND64_REDUCE_K(s, buffer[index] * buffer[index], SUM_HOT_FS);
    else ND64_REDUCE_K(s, buffer[index] * buffer[index], SUM_HOT_UR);
    *norm = s;
}
void orc_calculate_2norm(int x, int y, int hd, const double* buffer, double* norm)
{
    calculate_2norm_k(x, y, hd, buffer, norm, SUM_HOT_UR);
}

/* solver_methods.cpp:120-145 */
void orc_finalise(int x, int y, int hd, const double* u, const double* density, double* energy)
{
#pragma omp parallel for schedule(static)
    for (int jj = hd; jj < y - hd; ++jj)
        for (int kk = hd; kk < x - hd; ++kk) {
            const long i = IDX(jj, kk);
            energy[i] = u[i] / density[i];
        }
}

/* ------------------------------------------------------------------------- */
/* initialise.c:34-134 */
int orc_decompose(int grid_x, int grid_y, int num_chunks, int* x_chunks_out, int* y_chunks_out,
                  int* left, int* right, int* bottom, int* top, int* neighbours)
{
    double best_metric = DBL_MAX;
    double x_cells = (double)grid_x, y_cells = (double)grid_y;
    int x_chunks = 0, y_chunks = 0;
    for (int xx = 1; xx <= num_chunks; ++xx) {
        if (num_chunks % xx) continue;
        int yy = num_chunks / xx;
        if (num_chunks % yy) continue;
        double perimeter = ((x_cells / xx) * (x_cells / xx) + (y_cells / yy) * (y_cells / yy)) * 2;
        double area = (x_cells / xx) * (y_cells / yy);
        double current_metric = perimeter / area;
        if (current_metric < best_metric) {
            x_chunks = xx;
            y_chunks = yy;
            best_metric = current_metric;
        }
    }
    if (!x_chunks || !y_chunks) return 1;
    int dx = grid_x / x_chunks, dy = grid_y / y_chunks;
    int mod_x = grid_x % x_chunks, mod_y = grid_y % y_chunks;
    int add_x_prev = 0, add_y_prev = 0;
    for (int yy = 0; yy < y_chunks; ++yy) {
        int add_y = (yy < mod_y);
        for (int xx = 0; xx < x_chunks; ++xx) {
            int add_x = (xx < mod_x);
            int c = xx + yy * x_chunks;
            left[c] = xx * dx + add_x_prev;
            right[c] = left[c] + dx + add_x;
            bottom[c] = yy * dy + add_y_prev;
            top[c] = bottom[c] + dy + add_y;
            neighbours[c * 4 + ORC_LEFT] = (xx == 0) ? ORC_EXTERNAL : c - 1;
            neighbours[c * 4 + ORC_RIGHT] = (xx == x_chunks - 1) ? ORC_EXTERNAL : c + 1;
            neighbours[c * 4 + ORC_BOTTOM] = (yy == 0) ? ORC_EXTERNAL : c - x_chunks;
            neighbours[c * 4 + ORC_TOP] = (yy == y_chunks - 1) ? ORC_EXTERNAL : c + x_chunks;
            add_x_prev += add_x;
        }
        add_x_prev = 0;
        add_y_prev += add_y;
    }
    *x_chunks_out = x_chunks;
    *y_chunks_out = y_chunks;
    return 0;
}

/* eigenvalue_driver.c:71-122 (Numerical-Recipes tqli without eigenvectors) */
static int tqli(double* d, double* e, int n)
{
    int m, l, iter, i;
    double s, r, p, g, f, dd, c, b;
    for (i = 0; i < n - 1; i++) e[i] = e[i + 1];
    e[n - 1] = 0.0;
    for (l = 0; l < n; l++) {
        iter = 0;
        do {
            for (m = l; m < n - 1; m++) {
                dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) + dd == dd) break;
            }
            if (m == l) break;
            if (iter++ == 30) return 2;
            g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            r = sqrt((g * g) + 1.0);
            double sgn = (g) < 0 ? -fabs(r) : fabs(r);
            g = d[m] - d[l] + e[l] / (g + sgn);
            s = c = 1.0;
            p = 0.0;
            for (i = m - 1; i >= l; i--) {
                f = s * e[i];
                b = c * e[i];
                r = sqrt(f * f + g * g);
                e[i + 1] = r;
                if (r == 0.0) {
                    d[i + 1] -= p;
                    e[m] = 0.0;
                    continue;
                }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
            }
            d[l] = d[l] - p;
            e[l] = g;
            e[m] = 0.0;
        } while (m != l);
    }
    return 0;
}

/* eigenvalue_driver.c:11-67 */
int orc_eigenvalues(const double* cg_alphas, const double* cg_betas, int n, double* eigmin, double* eigmax)
{
    double* diag = (double*)calloc((size_t)n + 1, sizeof(double));
    double* offdiag = (double*)calloc((size_t)n + 1, sizeof(double));
    for (int ii = 0; ii < n; ++ii) {
        diag[ii] = 1.0 / cg_alphas[ii];
        if (ii > 0) diag[ii] += cg_betas[ii - 1] / cg_alphas[ii - 1];
        if (ii < n - 1) offdiag[ii + 1] = sqrt(cg_betas[ii]) / cg_alphas[ii];
    }
    int rc = tqli(diag, offdiag, n);
    double mn = DBL_MAX, mx = DBL_MIN;
    for (int ii = 0; ii < n; ++ii) {
        mn = (mn < diag[ii]) ? mn : diag[ii];
        mx = (mx > diag[ii]) ? mx : diag[ii];
    }
    free(diag);
    free(offdiag);
    if (rc) return rc;
    if (mn < 0.0 || mx < 0.0) return 1;
    *eigmin = mn * 0.95;
    *eigmax = mx * 1.05;
    return 0;
}

/* cheby_driver.c:163-183 */
void orc_cheby_coef(double eigmin, double eigmax, int max_iters, double* theta_out,
                    double* alphas, double* betas)
{
    double theta = (eigmax + eigmin) / 2.0;
    double delta = (eigmax - eigmin) / 2.0;
    double sigma = theta / delta;
    double rho_old = 1.0 / sigma;
    for (int ii = 0; ii < max_iters; ++ii) {
        double rho_new = 1.0 / (2.0 * sigma - rho_old);
        double cur_alpha = rho_new * rho_old;
        double cur_beta = 2.0 * rho_new / delta;
        alphas[ii] = cur_alpha;
        betas[ii] = cur_beta;
        rho_old = rho_new;
    }
    *theta_out = theta;
}

/* cheby_driver.c:146-160 (float logf/roundf, as written) */
int orc_cheby_est_iterations(double eigmin, double eigmax, double error, double bb)
{
    double condition_number = eigmax / eigmin;
    double it_alpha = DBL_EPSILON * bb / (4.0 * error);
    double gamm = (sqrt(condition_number) - 1.0) / (sqrt(condition_number) + 1.0);
    return (int)roundf(logf(it_alpha) / (2.0 * logf(gamm)));
}

/* ------------------------------------------------------------------------- */
/* Whole-deck driver: main.c:9-59, initialise.c:11-31, diffuse.c:10-78 and the
 * solver drivers, for N emulated ranks with one chunk each. */
typedef struct {
    int x, y, left, right, bottom, top;
    int nb[4];
    double *density, *energy0, *energy, *u, *u0, *p, *r, *w, *kx, *ky, *sd, *volume;
    double *cell_x, *cell_y, *vertex_x, *vertex_y;
    double *send[4], *recv[4]; /* indexed by face; chunk.c:12-24 */
} chunk_t;

typedef struct {
    const orc_deck* d;
    int nc, hd;
    chunk_t* ch;
    int fields[ORC_NUM_FIELDS];
    double dx, dy;
    double *cg_alphas, *cg_betas, *cheby_alphas, *cheby_betas;
    double theta, eigmin, eigmax;
    long calc_w_calls;
} run_t;

static double* zalloc(long n) { return (double*)calloc((size_t)n, sizeof(double)); }

static double* field_of(chunk_t* c, int f)
{
    switch (f) { /* remote_halo_driver.c:145-168 */
    case ORC_F_DENSITY: return c->density;
    case ORC_F_ENERGY0: return c->energy0;
    case ORC_F_ENERGY1: return c->energy;
    case ORC_F_U: return c->u;
    case ORC_F_P: return c->p;
    default: return c->sd;
    }
}

static void reset_fields(run_t* R) { memset(R->fields, 0, sizeof(R->fields)); }

/* remote_halo_driver.c:132-184 */
static int pack_or_unpack_all(run_t* R, chunk_t* c, int face, int depth, int pack, double* buffer)
{
    const int offset = (face == ORC_LEFT || face == ORC_RIGHT) ? c->y : c->x;
    int len = 0;
    for (int f = 0; f < ORC_NUM_FIELDS; ++f) {
        if (!R->fields[f]) continue;
        double* b = buffer + len;
        len += depth * offset;
        if (pack) orc_pack(c->x, c->y, R->hd, depth, face, field_of(c, f), b);
        else orc_unpack(c->x, c->y, R->hd, depth, face, field_of(c, f), b);
    }
    return len;
}

/* halo_update_driver.c:6-25 + remote_halo_driver.c:11-129 + kernel_interface.cpp:100-127 */
static void halo_update(run_t* R, int depth)
{
    int any = 0;
    for (int f = 0; f < ORC_NUM_FIELDS; ++f) any |= R->fields[f];
    if (!any) return;
    static const int opposite[4] = { ORC_RIGHT, ORC_LEFT, ORC_TOP, ORC_BOTTOM };
    for (int phase = 0; phase < 2; ++phase) { /* L/R completes (incl. unpack) before B/T packs */
        const int f0 = phase ? ORC_BOTTOM : ORC_LEFT;
        int lens[4] = { 0, 0, 0, 0 };
        for (int c = 0; c < R->nc; ++c)
            for (int face = f0; face < f0 + 2; ++face)
                if (R->ch[c].nb[face] != ORC_EXTERNAL)
                    lens[face] = pack_or_unpack_all(R, &R->ch[c], face, depth, 1, R->ch[c].send[face]);
        /* the "messages": my face-send lands in the neighbour's opposite-face recv */
        for (int c = 0; c < R->nc; ++c)
            for (int face = f0; face < f0 + 2; ++face) {
                int n = R->ch[c].nb[face];
                if (n == ORC_EXTERNAL) continue;
                const int offset = (face <= ORC_RIGHT) ? R->ch[c].y : R->ch[c].x;
                int nf = 0;
                for (int f = 0; f < ORC_NUM_FIELDS; ++f) nf += R->fields[f];
                memcpy(R->ch[n].recv[opposite[face]], R->ch[c].send[face],
                       sizeof(double) * (size_t)nf * depth * offset);
            }
        (void)lens;
        for (int c = 0; c < R->nc; ++c)
            for (int face = f0; face < f0 + 2; ++face)
                if (R->ch[c].nb[face] != ORC_EXTERNAL)
                    pack_or_unpack_all(R, &R->ch[c], face, depth, 0, R->ch[c].recv[face]);
    }
    for (int c = 0; c < R->nc; ++c) {
        chunk_t* k = &R->ch[c];
        static const int order[6] = { ORC_F_DENSITY, ORC_F_P, ORC_F_ENERGY0, ORC_F_ENERGY1, ORC_F_U, ORC_F_SD };
        static const int faces[4] = { ORC_LEFT, ORC_RIGHT, ORC_TOP, ORC_BOTTOM };
        for (int o = 0; o < 6; ++o) {
            if (!R->fields[order[o]]) continue;
            for (int q = 0; q < 4; ++q)
                if (k->nb[faces[q]] == ORC_EXTERNAL)
                    orc_local_halo(k->x, k->y, R->hd, depth, faces[q], field_of(k, order[o]));
        }
    }
}

/* sum_over_ranks on per-rank values: added in rank order */
static double sum_ranks(const double* v, int n)
{
    double s = v[0];
    for (int i = 1; i < n; ++i) s += v[i];
    return s;
}

#define FOR_CHUNKS for (int c = 0; c < R->nc; ++c)
#define K (&R->ch[c])

/* cg_driver.c:31-66 */
static void cg_init_driver(run_t* R, double rx, double ry, double* rro)
{
    double loc[64];
    FOR_CHUNKS {
        loc[c] = 0.0;
        orc_cg_init(K->x, K->y, R->hd, R->d->coefficient, rx, ry, K->density, K->energy, K->u, K->p,
                    K->r, K->w, K->kx, K->ky, &loc[c]);
    }
    reset_fields(R);
    R->fields[ORC_F_U] = 1;
    R->fields[ORC_F_P] = 1;
    halo_update(R, 1);
    *rro = sum_ranks(loc, R->nc);
    FOR_CHUNKS orc_copy_u(K->x, K->y, R->hd, K->u, K->u0);
}

/* cg_driver.c:69-124 */
static void cg_main_step(run_t* R, int tt, double* rro, double* error)
{
    double loc[64];
    FOR_CHUNKS {
        loc[c] = 0.0;
        orc_cg_calc_w(K->x, K->y, R->hd, K->p, K->kx, K->ky, K->w, &loc[c]);
    }
    R->calc_w_calls++;
    double pw = sum_ranks(loc, R->nc);
    double alpha = *rro / pw;
    R->cg_alphas[tt] = alpha;
    FOR_CHUNKS {
        loc[c] = 0.0;
        orc_cg_calc_ur(K->x, K->y, R->hd, alpha, K->p, K->w, K->u, K->r, &loc[c]);
    }
    double rrn = sum_ranks(loc, R->nc);
    double beta = rrn / *rro;
    R->cg_betas[tt] = beta;
    FOR_CHUNKS orc_cg_calc_p(K->x, K->y, R->hd, beta, K->r, K->p);
    *error = rrn;
    *rro = rrn;
}

/* cg_driver.c:7-28 */
static int cg_driver(run_t* R, double rx, double ry, double* error)
{
    int tt;
    double rro = 0.0;
    cg_init_driver(R, rx, ry, &rro);
    for (tt = 0; tt < R->d->max_iters; ++tt) {
        cg_main_step(R, tt, &rro, error);
        halo_update(R, 1);
        if (sqrt(fabs(*error)) < R->d->eps) break;
    }
    return tt;
}

/* folded: the GPU backend computes this norm inside the one-pass Chebyshev / PPCG kernel that produces r (tall
 * stencil tiles), not with the calculate_2norm kernel: same values, another summation tree in the replay mode. */
static double norm2_all(run_t* R, int which /*0=r,1=u0*/, int folded)
{
    double loc[64];
    FOR_CHUNKS calculate_2norm_k(K->x, K->y, R->hd, which ? K->u0 : K->r, &loc[c], folded ? SUM_HOT_FS : SUM_HOT_UR);
    return sum_ranks(loc, R->nc);
}

static int switch_rule(const orc_deck* d, int started, int tt, double error)
{
    /* cheby_driver.c:30-32, ppcg_driver.c:27-29; CG_ITERS_FOR_EIGENVALUES=20, ERROR_SWITCH_MAX=1.0 */
    return started || (d->error_switch ? (error < d->eps_lim) && (tt > 20)
                                       : (tt > d->presteps) && (error < 1.0));
}

/* cheby_driver.c:11-143 */
static int cheby_driver(run_t* R, double rx, double ry, double* error, int* n_cheby, int* est, int* rc)
{
    int tt, est_iterations = 0, num_cheby_iters = 0;
    double rro = 0.0;
    cg_init_driver(R, rx, ry, &rro);
    for (tt = 0; tt < R->d->max_iters; ++tt) {
        if (!switch_rule(R->d, num_cheby_iters, tt, *error)) {
            cg_main_step(R, tt, &rro, error);
        } else {
            num_cheby_iters++;
            int calc_2norm;
            double bb = 0.0;
            if (num_cheby_iters == 1) {
                *rc = orc_eigenvalues(R->cg_alphas, R->cg_betas, tt, &R->eigmin, &R->eigmax);
                if (*rc) return tt;
                orc_cheby_coef(R->eigmin, R->eigmax, R->d->max_iters - tt, &R->theta,
                               R->cheby_alphas, R->cheby_betas);
                double loc[64];
                FOR_CHUNKS {
                    orc_calculate_2norm(K->x, K->y, R->hd, K->u0, &loc[c]);
                    orc_cheby_init(K->x, K->y, R->hd, R->theta, K->u, K->u0, K->kx, K->ky, K->p, K->r, K->w);
                }
                reset_fields(R);
                R->fields[ORC_F_U] = 1;
                halo_update(R, 1);
                bb = sum_ranks(loc, R->nc);
                calc_2norm = 1;
            } else {
                calc_2norm = (num_cheby_iters >= est_iterations) && ((tt + 1) % 10 == 0);
            }
            FOR_CHUNKS orc_cheby_iterate(K->x, K->y, R->hd, R->cheby_alphas[num_cheby_iters],
                                         R->cheby_betas[num_cheby_iters], K->u, K->u0, K->kx, K->ky,
                                         K->p, K->r, K->w);
            if (calc_2norm) *error = norm2_all(R, 0, 1);
            if (num_cheby_iters == 1)
                est_iterations = orc_cheby_est_iterations(R->eigmin, R->eigmax, *error, bb);
        }
        halo_update(R, 1);
        if (fabs(*error) < R->d->eps) break;
    }
    *n_cheby = num_cheby_iters;
    *est = est_iterations;
    return tt;
}

/* ppcg_driver.c:10-189 */
static int ppcg_driver(run_t* R, double rx, double ry, double* error, int* n_ppcg, int* rc)
{
    int tt, num_ppcg_iters = 0;
    double rro = 0.0;
    cg_init_driver(R, rx, ry, &rro);
    for (tt = 0; tt < R->d->max_iters; ++tt) {
        if (!switch_rule(R->d, num_ppcg_iters, tt, *error)) {
            cg_main_step(R, tt, &rro, error);
        } else {
            num_ppcg_iters++;
            if (num_ppcg_iters == 1) {
                *rc = orc_eigenvalues(R->cg_alphas, R->cg_betas, tt, &R->eigmin, &R->eigmax);
                if (*rc) return tt;
                orc_cheby_coef(R->eigmin, R->eigmax, R->d->ppcg_inner_steps, &R->theta,
                               R->cheby_alphas, R->cheby_betas);
                /* ppcg_init_driver :66-84. The second sum_over_ranks(rro) of :83 is a
                 * no-op for one rank; the 1-rank semantics are kept for N chunks. */
                FOR_CHUNKS orc_calculate_residual(K->x, K->y, R->hd, K->u, K->u0, K->kx, K->ky, K->r);
                reset_fields(R);
                R->fields[ORC_F_P] = 1;
                halo_update(R, 1);
            }
            /* ppcg_main_step_driver :87-149 */
            double loc[64];
            FOR_CHUNKS {
                loc[c] = 0.0;
                orc_cg_calc_w(K->x, K->y, R->hd, K->p, K->kx, K->ky, K->w, &loc[c]);
            }
            R->calc_w_calls++;
            double pw = sum_ranks(loc, R->nc);
            double alpha = rro / pw;
            FOR_CHUNKS orc_cg_calc_ur(K->x, K->y, R->hd, alpha, K->p, K->w, K->u, K->r, &loc[c]);
            /* ppcg_inner_iterations :152-189 */
            FOR_CHUNKS orc_ppcg_init(K->x, K->y, R->hd, R->theta, K->r, K->sd);
            reset_fields(R);
            R->fields[ORC_F_SD] = 1;
            for (int pp = 0; pp < R->d->ppcg_inner_steps; ++pp) {
                halo_update(R, 1);
                FOR_CHUNKS orc_ppcg_inner_iteration(K->x, K->y, R->hd, R->cheby_alphas[pp],
                                                    R->cheby_betas[pp], K->u, K->r, K->kx, K->ky, K->sd);
            }
            reset_fields(R);
            R->fields[ORC_F_P] = 1;
            double rrn = norm2_all(R, 0, 1);
            double beta = rrn / rro;
            FOR_CHUNKS orc_cg_calc_p(K->x, K->y, R->hd, beta, K->r, K->p);
            *error = rrn;
            rro = rrn;
        }
        halo_update(R, 1);
        if (fabs(*error) < R->d->eps) break;
    }
    *n_ppcg = num_ppcg_iters;
    return tt;
}

/* jacobi_driver.c:7-84 */
static int jacobi_driver(run_t* R, double rx, double ry, double* error)
{
    FOR_CHUNKS {
        orc_jacobi_init(K->x, K->y, R->hd, R->d->coefficient, rx, ry, K->density, K->energy, K->u0,
                        K->u, K->kx, K->ky);
        orc_copy_u(K->x, K->y, R->hd, K->u, K->u0);
    }
    reset_fields(R);
    R->fields[ORC_F_U] = 1;
    int tt;
    for (tt = 0; tt < R->d->max_iters; ++tt) {
        double loc[64];
        FOR_CHUNKS orc_jacobi_iterate(K->x, K->y, R->hd, K->u, K->u0, K->r, K->kx, K->ky, &loc[c]);
        if (tt % 50 == 0) {
            halo_update(R, 1);
            FOR_CHUNKS {
                orc_calculate_residual(K->x, K->y, R->hd, K->u, K->u0, K->kx, K->ky, K->r);
                orc_calculate_2norm(K->x, K->y, R->hd, K->r, &loc[c]);
            }
        }
        *error = sum_ranks(loc, R->nc);
        halo_update(R, 1);
        if (fabs(*error) < R->d->eps) break;
    }
    return tt;
}

void orc_deck_defaults(orc_deck* d)
{
    memset(d, 0, sizeof(*d)); /* settings.h:15-45 */
    d->x_cells = 10; d->y_cells = 10;
    d->xmin = 0.0; d->ymin = 0.0; d->xmax = 100.0; d->ymax = 100.0;
    d->dt_init = 0.1; d->end_step = 2147483647; d->max_iters = 10000; d->eps = 1.0e-15;
    d->solver = ORC_CG; d->coefficient = ORC_CONDUCTIVITY; d->presteps = 30;
    d->ppcg_inner_steps = 10; d->error_switch = 0; d->eps_lim = 1e-5; d->halo_depth = 2;
    d->summary_frequency = 10; d->num_chunks = 1; d->num_states = 0;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int orc_run_deck(const orc_deck* d, orc_result* res, double* u_out, double* energy_out)
{
    run_t Rs, *R = &Rs;
    memset(R, 0, sizeof(*R));
    memset(res, 0, sizeof(*res));
    R->d = d;
    R->nc = d->num_chunks;
    R->hd = d->halo_depth;
    if (R->nc < 1 || R->nc > 64 || d->end_step > 64 || d->num_states < 1) return 10;
    const int hd = R->hd, nc = R->nc;
    R->dx = (d->xmax - d->xmin) / (double)d->x_cells; /* parse_config.c:188-191 */
    R->dy = (d->ymax - d->ymin) / (double)d->y_cells;

    int xc, yc, left[64], right[64], bottom[64], top[64], nb[256];
    if (orc_decompose(d->x_cells, d->y_cells, nc, &xc, &yc, left, right, bottom, top, nb)) return 11;

    R->ch = (chunk_t*)calloc((size_t)nc, sizeof(chunk_t));
    R->cg_alphas = zalloc(d->max_iters); R->cg_betas = zalloc(d->max_iters);
    R->cheby_alphas = zalloc(d->max_iters); R->cheby_betas = zalloc(d->max_iters);
    FOR_CHUNKS {
        K->left = left[c]; K->right = right[c]; K->bottom = bottom[c]; K->top = top[c];
        memcpy(K->nb, nb + 4 * c, sizeof(int) * 4);
        K->x = (right[c] - left[c]) + 2 * hd; /* chunk.c:7-8 */
        K->y = (top[c] - bottom[c]) + 2 * hd;
        const long n = (long)K->x * K->y;
        K->density = zalloc(n); K->energy0 = zalloc(n); K->energy = zalloc(n); K->u = zalloc(n);
        K->u0 = zalloc(n); K->p = zalloc(n); K->r = zalloc(n); K->w = zalloc(n); K->kx = zalloc(n);
        K->ky = zalloc(n); K->sd = zalloc(n); K->volume = zalloc(n);
        K->cell_x = zalloc(K->x); K->cell_y = zalloc(K->y);
        K->vertex_x = zalloc(K->x + 1); K->vertex_y = zalloc(K->y + 1);
        for (int f = 0; f < 4; ++f) {
            long len = (long)((f <= ORC_RIGHT) ? K->y : K->x) * hd * ORC_NUM_FIELDS;
            K->send[f] = zalloc(len);
            K->recv[f] = zalloc(len);
        }
        /* kernel_interface.cpp:34-52 */
        double x_min = d->xmin + R->dx * (double)K->left;
        double y_min = d->ymin + R->dy * (double)K->bottom;
        orc_set_chunk_data(K->x, K->y, hd, x_min, y_min, R->dx, R->dy, K->vertex_x, K->vertex_y,
                           K->cell_x, K->cell_y, K->volume);
        /* kernel_interface.cpp:54-72 */
        orc_set_chunk_initial_state(K->x, K->y, d->states[0].energy, d->states[0].density,
                                    K->energy0, K->density);
        for (int s = 1; s < d->num_states; ++s) {
            const orc_state* st = &d->states[s]; /* parse_config.c:253-260 */
            orc_set_chunk_state(K->x, K->y, hd, st->geometry, st->density, st->energy,
                                st->x_min + R->dx / 100.0, st->y_min + R->dy / 100.0,
                                st->x_max - R->dx / 100.0, st->y_max - R->dy / 100.0, st->radius,
                                K->energy0, K->density, K->u, K->cell_x, K->cell_y, K->vertex_x,
                                K->vertex_y);
        }
    }
    /* initialise.c:23-30 */
    reset_fields(R);
    R->fields[ORC_F_DENSITY] = R->fields[ORC_F_ENERGY0] = R->fields[ORC_F_ENERGY1] = 1;
    halo_update(R, 2);
    FOR_CHUNKS orc_store_energy(K->x, K->y, K->energy0, K->energy);

    int rc = 0;
    for (int step = 0; step < d->end_step && !rc; ++step) { /* diffuse.c:13-16, solve :23-78 */
        double dt = d->dt_init;
        double rx = dt / (R->dx * R->dx);
        double ry = dt / (R->dy * R->dy);
        reset_fields(R);
        R->fields[ORC_F_ENERGY1] = 1;
        R->fields[ORC_F_DENSITY] = 1;
        halo_update(R, 2);
        double error = 1e+10;
        double t0 = now_s();
        int tt = 0, nb_iters = 0, est = 0;
        switch (d->solver) {
        case ORC_JACOBI:
            tt = jacobi_driver(R, rx, ry, &error);
            res->iters_a[step] = tt;
            res->cell_iters += (long)d->x_cells * d->y_cells * (tt < d->max_iters ? tt + 1 : tt);
            break;
        case ORC_CG:
            tt = cg_driver(R, rx, ry, &error);
            res->iters_a[step] = tt; /* cg_driver.c:27 prints tt */
            res->cell_iters += (long)d->x_cells * d->y_cells * (tt < d->max_iters ? tt + 1 : tt);
            break;
        case ORC_CHEBY:
            tt = cheby_driver(R, rx, ry, &error, &nb_iters, &est, &rc);
            res->iters_a[step] = tt - nb_iters + 1; /* cheby_driver.c:73 */
            res->iters_b[step] = nb_iters;
            res->est_iters[step] = est;
            res->cell_iters += (long)d->x_cells * d->y_cells * (tt < d->max_iters ? tt + 1 : tt);
            break;
        case ORC_PPCG:
            tt = ppcg_driver(R, rx, ry, &error, &nb_iters, &rc);
            res->iters_a[step] = tt - nb_iters + 1; /* ppcg_driver.c:59 */
            res->iters_b[step] = nb_iters;
            res->cell_iters += (long)d->x_cells * d->y_cells *
                               ((tt < d->max_iters ? tt + 1 : tt) + (long)nb_iters * d->ppcg_inner_steps);
            break;
        default: rc = 12;
        }
        res->wall_solve_s += now_s() - t0;
        res->error[step] = error;
        res->eigmin[step] = R->eigmin;
        res->eigmax[step] = R->eigmax;
        /* solve_finished_driver.c:7-43 (check_result defaults to 1; the norm is unused) */
        FOR_CHUNKS orc_calculate_residual(K->x, K->y, hd, K->u, K->u0, K->kx, K->ky, K->r);
        (void)norm2_all(R, 0, 0);
        FOR_CHUNKS orc_finalise(K->x, K->y, hd, K->u, K->density, K->energy);
        R->fields[ORC_F_ENERGY1] = 1;
        halo_update(R, 1);
    }
    res->calc_w_calls = R->calc_w_calls;

    /* field_summary_driver.c:8-30 */
    {
        double v[64], m[64], e[64], t[64];
        FOR_CHUNKS orc_field_summary(K->x, K->y, hd, K->volume, K->density, K->energy0, K->u,
                                     &v[c], &m[c], &e[c], &t[c]);
        res->vol = sum_ranks(v, nc); res->mass = sum_ranks(m, nc);
        res->ie = sum_ranks(e, nc); res->temp = sum_ranks(t, nc);
    }
    FOR_CHUNKS {
        for (int jj = hd; jj < K->y - hd; ++jj)
            for (int kk = hd; kk < K->x - hd; ++kk) {
                long g = (long)(K->left + kk - hd) + (long)(K->bottom + jj - hd) * d->x_cells;
                long i = (long)kk + (long)jj * K->x;
                if (u_out) u_out[g] = K->u[i];
                if (energy_out) energy_out[g] = K->energy[i];
            }
    }
    FOR_CHUNKS {
        free(K->density); free(K->energy0); free(K->energy); free(K->u); free(K->u0); free(K->p);
        free(K->r); free(K->w); free(K->kx); free(K->ky); free(K->sd); free(K->volume);
        free(K->cell_x); free(K->cell_y); free(K->vertex_x); free(K->vertex_y);
        for (int f = 0; f < 4; ++f) { free(K->send[f]); free(K->recv[f]); }
    }
    free(R->ch); free(R->cg_alphas); free(R->cg_betas); free(R->cheby_alphas); free(R->cheby_betas);
    return rc;
}
