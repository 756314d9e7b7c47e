// tl_kernels.cu -- sm_100a CUDA kernels of the TeaLeaf backend.
//
// Every kernel cites the reference kernel whose arithmetic it reproduces
// (paths relative to /root/reference/TeaLeaf/c_kernels/sycl/).  Compiled with -fmad=false:
// each fp64 operation is rounded exactly as the reference expression writes it, so all
// element-wise results are bit-identical to the CPU oracle.  Reductions are deterministic
// (fixed per-thread order -> xor-butterfly -> fixed-order sum of per-tile partials by the
// last CTA to finish), single pass: no second launch, no host round trip.
//
// None of this is GEMM-shaped: the path is HBM-bound (0.22 flop/B), so the design rules are
// coalesced 128-bit accesses on 128-byte-aligned rows, enough independent loads in flight,
// and the minimum number of passes over HBM.  Tensor cores / TMEM are not used.
#include <stdlib.h>
#include "tl_internal.h"
#include "tl_device.cuh"

long g_tl_launches = 0;

// ---------------------------------------------------------------------------------------------
// Generic row-tiled kernel: one thread per column, `rows` rows per CTA, optional NR-wide reduction.
// Used for the setup / once-per-timestep kernels.  F: (long i, int jj, int kk, double* acc).
// Fin: (const double* totals, DevScal* S) run by one thread after the grid reduction.
// ---------------------------------------------------------------------------------------------
template <int NR, class F, class Fin>
__global__ void __launch_bounds__(TL_TPB) k_generic(Geo g, int k_lo, int k_hi, int j_lo, int j_hi, int rows,
                                                    F f, Fin fin, RedArgs ra)
{
    const int kk = k_lo + blockIdx.x * TL_TPB + threadIdx.x;
    const int j0 = j_lo + blockIdx.y * rows;
    const int j1 = min(j0 + rows, j_hi);
    double acc[NR > 0 ? NR : 1];
#pragma unroll
    for (int r = 0; r < (NR > 0 ? NR : 1); ++r) acc[r] = 0.0;
    if (kk < k_hi) {
        long i = (long)g.off + (long)j0 * g.pitch + kk;
        for (int jj = j0; jj < j1; ++jj, i += g.pitch) f(i, jj, kk, acc);
    }
    if constexpr (NR > 0) {
        double tot[NR > 0 ? NR : 1];
        const int tile = blockIdx.y * gridDim.x + blockIdx.x;
        if (grid_reduce<(NR > 0 ? NR : 1)>(acc, ra, tile, gridDim.x * gridDim.y, tot) && threadIdx.x == 0)
            fin(tot, ra.S);
    }
}

struct NoFin {
    __device__ void operator()(const double*, DevScal*) const {}
};

template <int NR, class F, class Fin>
static int launch_generic(tl_chunk* c, int k_lo, int k_hi, int j_lo, int j_hi, F f, Fin fin)
{
    if (k_hi <= k_lo || j_hi <= j_lo) return TL_OK;
    const int rows = 8;
    dim3 grid((k_hi - k_lo + TL_TPB - 1) / TL_TPB, (j_hi - j_lo + rows - 1) / rows);
    if (NR > 0 && (long)grid.x * grid.y > c->partial_cap) {
        tl_set_error("partials capacity exceeded");
        return TL_ERR_ARG;
    }
    RedArgs ra{c->partials, c->gpartials, c->gcount, c->partial_cap, c->gpartial_cap, c->scal};
    k_generic<NR, F, Fin><<<grid, TL_TPB, 0, c->stream>>>(c->g, k_lo, k_hi, j_lo, j_hi, rows, f, fin, ra);
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// ---------------------------------------------------------------------------------------------
// Vectorised row-tiled kernel for the streaming / reduction kernels outside the CG loop: two columns per thread
// (double2, 16-byte aligned: the first column of a thread is chosen so that off + kk is even), `rows` rows per CTA,
// and the loads of U rows are issued back to back before any of them is consumed (LD returns a plain struct V of
// loaded values, AP consumes it): the scalar one-row-at-a-time skeleton above keeps 8 bytes per thread in flight and
// reaches 45-75 % of the HBM peak on these kernels, this one 16 x (fields) x U.
// LD: V (long i, int jj, int kk).  AP: (const V&, long i, int jj, int kk, bool v0, bool v1, double* acc); v0 / v1 say
// which of the thread's two columns lie inside [k_lo, k_hi).  Per-thread accumulation order is row-major, column 0
// before column 1 (what the oracle's replay of the reduction tree assumes for the two-column tile).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_mask(double* p, double2 v, bool v0, bool v1)
{
    if (v0 && v1) st2(p, v);
    else if (v0) p[0] = v.x;
    else if (v1) p[1] = v.y;
}
template <int NR, int U, class V, class LD, class AP, class Fin>
__global__ void __launch_bounds__(TL_TPB) k_vec(Geo g, int kbase, int k_lo, int k_hi, int j_lo, int j_hi, int rows,
                                                LD ld, AP ap, Fin fin, RedArgs ra)
{
    const int kk = kbase + 2 * (blockIdx.x * TL_TPB + threadIdx.x);
    const bool v0 = kk >= k_lo && kk < k_hi, v1 = kk + 1 >= k_lo && kk + 1 < k_hi;
    const int j0 = j_lo + blockIdx.y * rows;
    const int j1 = min(j0 + rows, j_hi);
    double acc[NR > 0 ? NR : 1];
#pragma unroll
    for (int r = 0; r < (NR > 0 ? NR : 1); ++r) acc[r] = 0.0;
    if (v0 || v1) {
        long i = (long)g.off + (long)j0 * g.pitch + kk;
        for (int jb = j0; jb < j1; jb += U, i += (long)U * g.pitch) {
            V vals[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (jb + u < j1) vals[u] = ld(i + (long)u * g.pitch, jb + u, kk);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (jb + u < j1) ap(vals[u], i + (long)u * g.pitch, jb + u, kk, v0, v1, acc);
        }
    }
    if constexpr (NR > 0) {
        double tot[NR > 0 ? NR : 1];
        const int tile = blockIdx.y * gridDim.x + blockIdx.x;
        if (grid_reduce<(NR > 0 ? NR : 1)>(acc, ra, tile, gridDim.x * gridDim.y, tot) && threadIdx.x == 0)
            fin(tot, ra.S);
    }
}

template <int NR, int U, class V, class LD, class AP, class Fin>
static int launch_vec(tl_chunk* c, int k_lo, int k_hi, int j_lo, int j_hi, LD ld, AP ap, Fin fin)
{
    if (k_hi <= k_lo || j_hi <= j_lo) return TL_OK;
    const int rows = 8;
    const int kbase = k_lo - ((c->g.off + k_lo) & 1);
    dim3 grid((k_hi - kbase + TL_TILE_COLS - 1) / TL_TILE_COLS, (j_hi - j_lo + rows - 1) / rows);
    if (NR > 0 && (long)grid.x * grid.y > c->partial_cap) {
        tl_set_error("partials capacity exceeded");
        return TL_ERR_ARG;
    }
    RedArgs ra{c->partials, c->gpartials, c->gcount, c->partial_cap, c->gpartial_cap, c->scal};
    k_vec<NR, U, V, LD, AP, Fin><<<grid, TL_TPB, 0, c->stream>>>(c->g, kbase, k_lo, k_hi, j_lo, j_hi, rows, ld, ap, fin, ra);
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}
struct V1 { double2 a; };
struct V2 { double2 a, b; };
struct V3 { double2 a, b, c; };
struct V4 { double2 a, b, c, d; };

#define INTERIOR_RANGE(c) (c)->g.hd, (c)->g.x - (c)->g.hd, (c)->g.hd, (c)->g.y - (c)->g.hd
#define ALL_RANGE(c) 0, (c)->g.x, 0, (c)->g.y

// ---------------------------------------------------------------------------------------------
// Setup kernels
// ---------------------------------------------------------------------------------------------
// set_chunk_data.cpp:8-28
__global__ void k_vertices(int x, int y, int hd, double x_min, double y_min, double dx, double dy,
                           double* vertex_x, double* vertex_y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < x + 1) vertex_x[i] = x_min + dx * ((double)i - (double)hd);
    if (i < y + 1) vertex_y[i] = y_min + dy * ((double)i - (double)hd);
}
// set_chunk_data.cpp:31-64 (cell centres)
__global__ void k_cells(int x, int y, const double* vertex_x, const double* vertex_y, double* cell_x,
                        double* cell_y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < x) cell_x[i] = 0.5 * (vertex_x[i] + vertex_x[i + 1]);
    if (i < y) cell_y[i] = 0.5 * (vertex_y[i] + vertex_y[i + 1]);
}

int tlk_set_chunk_data(tl_chunk* c, double x_min, double y_min, double dx, double dy)
{
    const int n = max(c->g.x, c->g.y) + 1;
    k_vertices<<<(n + 127) / 128, 128, 0, c->stream>>>(c->g.x, c->g.y, c->g.hd, x_min, y_min, dx, dy,
                                                      c->vertex_x, c->vertex_y);
    k_cells<<<(n + 127) / 128, 128, 0, c->stream>>>(c->g.x, c->g.y, c->vertex_x, c->vertex_y, c->cell_x,
                                                   c->cell_y);
    g_tl_launches += 2;
    TL_CUDA(cudaGetLastError());
    double* volume = c->f[TL_FIELD_VOLUME];
    const double v = dx * dy; // set_chunk_data.cpp:54
    return launch_generic<0>(c, ALL_RANGE(c),
                             [=] __device__(long i, int, int, double*) { volume[i] = v; }, NoFin());
}

// set_chunk_state.cpp:8-27
int tlk_set_initial_state(tl_chunk* c, double energy, double density)
{
    double* e0 = c->f[TL_FIELD_ENERGY0];
    double* d = c->f[TL_FIELD_DENSITY];
    return launch_generic<0>(c, ALL_RANGE(c),
                             [=] __device__(long i, int, int, double*) {
                                 e0[i] = energy;
                                 d[i] = density;
                             },
                             NoFin());
}

// set_chunk_state.cpp:30-92
int tlk_set_state(tl_chunk* c, const tl_state* sp)
{
    const tl_state s = *sp;
    double* e0 = c->f[TL_FIELD_ENERGY0];
    double* d = c->f[TL_FIELD_DENSITY];
    double* u = c->f[TL_FIELD_U];
    const double *cx = c->cell_x, *cy = c->cell_y, *vx = c->vertex_x, *vy = c->vertex_y;
    const int x = c->g.x, y = c->g.y;
    return launch_generic<0>(
        c, ALL_RANGE(c),
        [=] __device__(long i, int jj, int kk, double*) {
            bool apply = false;
            if (s.geometry == TL_GEOM_RECTANGULAR) {
                apply = (vx[kk + 1] >= s.x_min && vx[kk] < s.x_max && vy[jj + 1] >= s.y_min &&
                         vy[jj] < s.y_max);
            } else if (s.geometry == TL_GEOM_CIRCULAR) {
                double radius = sqrt((cx[kk] - s.x_min) * (cx[kk] - s.x_min) +
                                     (cy[jj] - s.y_min) * (cy[jj] - s.y_min));
                apply = (radius <= s.radius);
            } else if (s.geometry == TL_GEOM_POINT) {
                apply = (vx[kk] == s.x_min && vy[jj] == s.y_min);
            }
            if (apply) {
                e0[i] = s.energy;
                d[i] = s.density;
            }
            if (kk > 0 && kk < x - 1 && jj > 0 && jj < y - 1) u[i] = e0[i] * d[i];
        },
        NoFin());
}

// store_energy.cpp:6-24 (all cells), solver_methods.cpp:7-31 copy_u (interior), jacobi.cpp:120-137
int tlk_copy_field(tl_chunk* c, int dst, int src, bool interior_only)
{
    double* D = c->f[dst];
    const double* S = c->f[src];
    auto ld = [=] __device__(long i, int, int) { return V1{ld2(S + i)}; };
    auto ap = [=] __device__(const V1& v, long i, int, int, bool v0, bool v1, double*) { st_mask(D + i, v.a, v0, v1); };
    if (interior_only) return launch_vec<0, 4, V1>(c, INTERIOR_RANGE(c), ld, ap, NoFin());
    return launch_vec<0, 4, V1>(c, ALL_RANGE(c), ld, ap, NoFin());
}

// field_summary.cpp:6-151
int tlk_field_summary(tl_chunk* c)
{
    const double *vol = c->f[TL_FIELD_VOLUME], *den = c->f[TL_FIELD_DENSITY], *e0 = c->f[TL_FIELD_ENERGY0],
                 *u = c->f[TL_FIELD_U];
    return launch_vec<4, 2, V4>(
        c, INTERIOR_RANGE(c),
        [=] __device__(long i, int, int) { return V4{ld2(vol + i), ld2(den + i), ld2(e0 + i), ld2(u + i)}; },
        [=] __device__(const V4& v, long, int, int, bool v0, bool v1, double* acc) {
            if (v0) {
                const double cv = v.a.x;
                const double cm = cv * v.b.x;
                acc[0] += cv;
                acc[1] += cm;
                acc[2] += cm * v.c.x;
                acc[3] += cm * v.d.x;
            }
            if (v1) {
                const double cv = v.a.y;
                const double cm = cv * v.b.y;
                acc[0] += cv;
                acc[1] += cm;
                acc[2] += cm * v.c.y;
                acc[3] += cm * v.d.y;
            }
        },
        [=] __device__(const double* t, DevScal* S) {
            S->sums[0] = t[0];
            S->sums[1] = t[1];
            S->sums[2] = t[2];
            S->sums[3] = t[3];
        });
}

// ---------------------------------------------------------------------------------------------
// Halo kernels
// ---------------------------------------------------------------------------------------------
struct FieldList {
    double* f[TL_NUM_EXCHANGE_FIELDS];
    int n;
};

// local_halos.cpp:7-54 (update_left / update_right): every row jj in [0,y), `depth` layers.
__global__ void k_halo_lr(Geo g, FieldList fl, int depth, int do_left, int do_right)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int jj = t / depth, d = t % depth;
    if (jj >= g.y) return;
    double* a = fl.f[blockIdx.y];
    const long row = (long)g.off + (long)jj * g.pitch;
    if (do_left) a[row + g.hd - 1 - d] = a[row + g.hd + d];
    if (do_right) a[row + g.x - g.hd + d] = a[row + g.x - g.hd - 1 - d];
}
// local_halos.cpp:57-102 (update_top / update_bottom): every column kk in [0,x).
__global__ void k_halo_tb(Geo g, FieldList fl, int depth, int do_top, int do_bottom)
{
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= g.x) return;
    double* a = fl.f[blockIdx.y];
    for (int d = 0; d < depth; ++d) {
        if (do_top) a[(long)g.off + (long)(g.y - g.hd + d) * g.pitch + kk] =
            a[(long)g.off + (long)(g.y - g.hd - 1 - d) * g.pitch + kk];
        if (do_bottom) a[(long)g.off + (long)(g.hd - 1 - d) * g.pitch + kk] =
            a[(long)g.off + (long)(g.hd + d) * g.pitch + kk];
    }
}

// kernel_interface.cpp:100-127: per flagged field, faces L, R, T, B where the neighbour is external.
// L/R of all fields in one launch, then T/B in a second (T/B reads the columns L/R just wrote).
int tlk_local_halos(tl_chunk* c, const int fields[6], int depth)
{
    FieldList fl;
    fl.n = 0;
    for (int i = 0; i < TL_NUM_EXCHANGE_FIELDS; ++i)
        if (fields[i]) fl.f[fl.n++] = c->f[i];
    if (!fl.n) return TL_OK;
    const int L = c->nb[TL_FACE_LEFT] == TL_EXTERNAL_FACE, R = c->nb[TL_FACE_RIGHT] == TL_EXTERNAL_FACE;
    const int T = c->nb[TL_FACE_TOP] == TL_EXTERNAL_FACE, B = c->nb[TL_FACE_BOTTOM] == TL_EXTERNAL_FACE;
    if (L || R) {
        dim3 grid((c->g.y * depth + 127) / 128, fl.n);
        k_halo_lr<<<grid, 128, 0, c->stream>>>(c->g, fl, depth, L, R);
        ++g_tl_launches;
    }
    if (T || B) {
        dim3 grid((c->g.x + 127) / 128, fl.n);
        k_halo_tb<<<grid, 128, 0, c->stream>>>(c->g, fl, depth, T, B);
        ++g_tl_launches;
    }
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// pack_halos.cpp:7-184.  Buffer index b: L/R  b = jj*depth + d ; T/B  b = d*x + kk.  Fields are
// concatenated in exchange-index order, depth*y (L/R) or depth*x (T/B) doubles each
// (remote_halo_driver.c:132-184).
__global__ void k_pack(Geo g, FieldList fl, int depth, int face, int pack, double* buf)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool lr = (face == TL_FACE_LEFT || face == TL_FACE_RIGHT);
    const int per_field = depth * (lr ? g.y : g.x);
    if (b >= per_field) return;
    double* a = fl.f[blockIdx.y];
    double* bp = buf + (size_t)blockIdx.y * per_field + b;
    int jj, kk;
    if (lr) {
        jj = b / depth;
        const int d = b % depth;
        if (face == TL_FACE_LEFT) kk = pack ? g.hd + d : g.hd - depth + d;
        else kk = pack ? g.x - g.hd - depth + d : g.x - g.hd + d;
    } else {
        const int d = b / g.x;
        kk = b % g.x;
        if (face == TL_FACE_TOP) jj = pack ? g.y - g.hd - depth + d : g.y - g.hd + d;
        else jj = pack ? g.hd + d : g.hd - depth + d;
    }
    const long i = (long)g.off + (long)jj * g.pitch + kk;
    if (pack) *bp = a[i];
    else a[i] = *bp;
}

int tlk_pack_face(tl_chunk* c, const int fields[6], int depth, int face, bool pack, double* devbuf, int* len)
{
    FieldList fl;
    fl.n = 0;
    for (int i = 0; i < TL_NUM_EXCHANGE_FIELDS; ++i)
        if (fields[i]) fl.f[fl.n++] = c->f[i];
    const bool lr = (face == TL_FACE_LEFT || face == TL_FACE_RIGHT);
    const int per_field = depth * (lr ? c->g.y : c->g.x);
    if (len) *len = per_field * fl.n;
    if (!fl.n) return TL_OK;
    dim3 grid((per_field + 127) / 128, fl.n);
    k_pack<<<grid, 128, 0, c->stream>>>(c->g, fl, depth, face, pack ? 1 : 0, devbuf);
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// One launch per exchange phase and direction: both faces of a phase (L+R or B+T) and all flagged fields.
//   k_pack_send   : gathers into the NEIGHBOUR's receive buffer (peer-mapped, NVLink stores); the last CTA
//                   to finish releases the neighbours' arrival flags (every CTA fences its stores first).
//   k_wait_unpack : thread 0 of each CTA acquires its face's arrival flag, then the CTA scatters; the last CTA
//                   to finish releases the senders' "consumed" (ack) flags.
// Message n of a face (n = 1, 2, ...: a per-face counter, identical on both ends) travels through receive buffer
// n & 1 and raises the arrival flag to n.  The sender of message n first acquires ack >= n - 2: the message that last
// used that buffer has been unpacked.  (With a symmetric exchange the two buffers alone already guarantee this --
// send(n+2) follows my unpack(n+1), which follows the neighbour's send(n+1), which follows its unpack(n) -- the ack
// makes it hold without that argument, e.g. under rank skew with one-directional traffic.)
struct PhaseFaces {
    int face[2];                  // TL_FACE_* or -1
    double* buf[2];               // pack: neighbour's receive buffer (parity n & 1); unpack: mine
    unsigned long long* flag[2];  // pack: neighbour's arrival flag;   unpack: my arrival flag
    unsigned long long* ack[2];   // pack: my ack flag (the neighbour releases it); unpack: the sender's ack flag
    unsigned long long seq[2];    // n
};

__device__ __forceinline__ long halo_cell_index(const Geo& g, int face, int depth, int pack, int b)
{
    const bool lr = (face == TL_FACE_LEFT || face == TL_FACE_RIGHT);
    int jj, kk;
    if (lr) {
        jj = b / depth;
        const int d = b % depth;
        if (face == TL_FACE_LEFT) kk = pack ? g.hd + d : g.hd - depth + d;
        else kk = pack ? g.x - g.hd - depth + d : g.x - g.hd + d;
    } else {
        const int d = b / g.x;
        kk = b % g.x;
        if (face == TL_FACE_TOP) jj = pack ? g.y - g.hd - depth + d : g.y - g.hd + d;
        else jj = pack ? g.hd + d : g.hd - depth + d;
    }
    return (long)g.off + (long)jj * g.pitch + kk;
}

__global__ void k_pack_send(Geo g, FieldList fl, int depth, PhaseFaces pf, DevScal* S)
{
    const int q = blockIdx.z;
    const int face = pf.face[q];
    if (face >= 0) {
        if (pf.seq[q] > 2ull) { // block-uniform
            if (threadIdx.x == 0) spin_flag(pf.ack[q], pf.seq[q] - 2ull, S, 60ull + face);
            __syncthreads();
        }
        const bool lr = (face == TL_FACE_LEFT || face == TL_FACE_RIGHT);
        const int per_field = depth * (lr ? g.y : g.x);
        const int b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b < per_field)
            pf.buf[q][(size_t)blockIdx.y * per_field + b] = fl.f[blockIdx.y][halo_cell_index(g, face, depth, 1, b)];
    }
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        s_last = (atomicAdd(&S->counter[2], 1u) == total - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 2) {
        if (threadIdx.x == 0) S->counter[2] = 0u;
        if (pf.face[threadIdx.x] >= 0) st_release_sys(pf.flag[threadIdx.x], pf.seq[threadIdx.x]);
    }
}

__global__ void k_wait_unpack(Geo g, FieldList fl, int depth, PhaseFaces pf, DevScal* S)
{
    const int q = blockIdx.z;
    const int face = pf.face[q];
    if (face >= 0) { // block-uniform
        if (threadIdx.x == 0) spin_flag(pf.flag[q], pf.seq[q], S, 50ull + face);
        __syncthreads();
        const bool lr = (face == TL_FACE_LEFT || face == TL_FACE_RIGHT);
        const int per_field = depth * (lr ? g.y : g.x);
        const int b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b < per_field)
            fl.f[blockIdx.y][halo_cell_index(g, face, depth, 0, b)] =
                __ldcg(pf.buf[q] + (size_t)blockIdx.y * per_field + b);
    }
    // every CTA's reads of the receive buffer are complete before its ticket; the last one hands the buffers back
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        s_last = (atomicAdd(&S->counter[3], 1u) == total - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 2) {
        if (threadIdx.x == 0) S->counter[3] = 0u;
        if (pf.face[threadIdx.x] >= 0) st_release_sys(pf.ack[threadIdx.x], pf.seq[threadIdx.x]);
    }
}

// faces/bufs/flags/acks/seqs: two entries (one per face of the phase; face -1 = no neighbour on that side)
int tlk_phase_exchange(tl_chunk* c, const int fields[6], int depth, bool send, const int faces[2], double* const bufs[2],
                       unsigned long long* const flags[2], unsigned long long* const acks[2],
                       const unsigned long long seqs[2])
{
    FieldList fl;
    fl.n = 0;
    for (int i = 0; i < TL_NUM_EXCHANGE_FIELDS; ++i)
        if (fields[i]) fl.f[fl.n++] = c->f[i];
    if (!fl.n || (faces[0] < 0 && faces[1] < 0)) return TL_OK;
    PhaseFaces pf;
    int per_field = 0;
    for (int q = 0; q < 2; ++q) {
        pf.face[q] = faces[q];
        pf.buf[q] = bufs[q];
        pf.flag[q] = flags[q];
        pf.ack[q] = acks[q];
        pf.seq[q] = seqs[q];
        if (faces[q] >= 0) {
            const bool lr = (faces[q] == TL_FACE_LEFT || faces[q] == TL_FACE_RIGHT);
            const int n = depth * (lr ? c->g.y : c->g.x);
            per_field = n > per_field ? n : per_field;
        }
    }
    dim3 grid((per_field + 127) / 128, fl.n, 2);
    if (send) k_pack_send<<<grid, 128, 0, c->stream>>>(c->g, fl, depth, pf, c->scal);
    else k_wait_unpack<<<grid, 128, 0, c->stream>>>(c->g, fl, depth, pf, c->scal);
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// ---------------------------------------------------------------------------------------------
// CG
// ---------------------------------------------------------------------------------------------
// cg.cpp:7-134 via kernel_interface.cpp:192-210: cg_init_u, cg_init_k, cg_init_others (+ r.p) in ONE pass:
// 64 B/cell (read energy, density; write u, kx, ky, w, r, p) instead of the 120 B of the three reference kernels.
// u, the conductivity w and kx / ky of the neighbouring cells are recomputed from energy and density (same operations
// on the same inputs: bit-identical to what the owning cell stores); every cell of every field ends up exactly as the
// three-kernel sequence leaves it (p = r = 0 and w = conductivity outside the interior included).
int tlk_cg_init(tl_chunk* c, int coefficient, double rx, double ry)
{
    double *p = c->f[TL_FIELD_P], *r = c->f[TL_FIELD_R], *u = c->f[TL_FIELD_U], *w = c->f[TL_FIELD_W];
    double *kx = c->f[TL_FIELD_KX], *ky = c->f[TL_FIELD_KY];
    const double *den = c->f[TL_FIELD_DENSITY], *en = c->f[TL_FIELD_ENERGY1];
    const int x = c->g.x, y = c->g.y, hd = c->g.hd, pitch = c->g.pitch;
    const bool cond = (coefficient == TL_CONDUCTIVITY);
    return launch_generic<1>(
        c, ALL_RANGE(c),
        [=] __device__(long i, int jj, int kk, double* acc) {
            auto wf = [&](double d) { return cond ? d : 1.0 / d; };       // cg.cpp:34-36 (cg_init_u)
            const bool ring1 = jj > 0 && jj < y - 1 && kk > 0 && kk < x - 1;
            const bool kcell = jj >= hd && jj < y - 1 && kk >= hd && kk < x - 1;
            const bool interior = jj >= hd && jj < y - hd && kk >= hd && kk < x - hd;
            const double dc = den[i];
            const double uc = en[i] * dc;
            u[i] = uc;
            const double wc = wf(dc);
            double kxc = 0.0, kyc = 0.0;
            if (kcell) {                                                   // cg.cpp:61-63 (cg_init_k)
                const double wl = wf(den[i - 1]), wd = wf(den[i - pitch]);
                kxc = rx * (wl + wc) / (2.0 * wl * wc);
                kyc = ry * (wd + wc) / (2.0 * wd * wc);
                kx[i] = kxc;
                ky[i] = kyc;
            }
            if (interior) {                                                // cg.cpp:96-127 (cg_init_others)
                const double wr = wf(den[i + 1]), wu = wf(den[i + pitch]);
                const double kxr = rx * (wc + wr) / (2.0 * wc * wr);       // kx[i + 1]
                const double kyu = ry * (wc + wu) / (2.0 * wc * wu);       // ky[i + x]
                const double sv = smvp(kxc, kxr, kyc, kyu, uc, en[i - 1] * den[i - 1], en[i + 1] * den[i + 1],
                                       en[i - pitch] * den[i - pitch], en[i + pitch] * den[i + pitch]);
                w[i] = sv;
                const double rv = uc - sv;
                r[i] = rv;
                p[i] = rv;
                acc[0] += rv * rv;
            } else {
                p[i] = 0.0;
                r[i] = 0.0;
                if (ring1) w[i] = wc;
            }
        },
        [=] __device__(const double* t, DevScal* S) { S->sums[0] = t[0]; });
}

__global__ void k_seed_rro(DevScal* S) { S->rro = S->sums[0]; }
int tlk_seed_rro(tl_chunk* c)
{
    k_seed_rro<<<1, 1, 0, c->stream>>>(c->scal);
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

__global__ void k_reset_scal(DevScal* S, double eps, int max_iters)
{
    S->rro = S->pw = S->rrn = S->alpha = S->beta = 0.0;
    S->error = 1e+10; // diffuse.c:44
    S->eps = eps;
    S->iters = 0;
    S->conv = 0;
    S->p_pending = 0;
    S->max_iters = max_iters;
    S->conv_mode = 0;
    S->sw_min_iters = -1;
    S->sw_thresh = 0.0;
    S->pad = 0u;
    S->dbg[0] = S->dbg[1] = S->dbg[2] = S->dbg[3] = 0ull;
    S->counter[0] = 0u;
    S->counter[1] = 0u;
    S->counter[2] = 0u;
    S->counter[3] = 0u;
}
int tlk_reset_solve_scalars(tl_chunk* c, double eps, int max_iters)
{
    c->resident_iters = 0;
    k_reset_scal<<<1, 1, 0, c->stream>>>(c->scal, eps, max_iters);
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// Tuning knobs (tl_set_tuning): rows per tile and rows per load batch for each hot kernel family.
// The load batch U is the number of rows whose loads are issued back to back before any of them is
// consumed: it sets the bytes each thread keeps in flight (the kernels are latency-bound, not
// issue-bound; see DESIGN.md "in-flight bytes").
// rows == 0 selects the built-in heuristic (tile_rows).  Defaults from the round-1 sweep on B200
// (profiles/tuning_r01.txt): stencil kernels want tall tiles (the two extra rows of p per tile are
// re-read from L2), the streaming kernels want many small tiles (better tail balance).
static int g_rows[4] = {0, 0, 0, 0};
static int g_batch[4] = {2, 4, 4, 1};
static bool g_batch_user[4] = {false, false, false, false}; // set through tl_set_tuning: no heuristic on top
extern "C" int tl_set_tuning(int kernel, int rows, int batch)
{
    if (kernel < 0 || kernel > 3 || rows < 0 || rows > 256 || (batch != 1 && batch != 2 && batch != 4)) {
        tl_set_error("tl_set_tuning: bad arguments");
        return TL_ERR_ARG;
    }
    g_rows[kernel] = rows;
    g_batch[kernel] = batch;
    g_batch_user[kernel] = true;
    return TL_OK;
}
// TL_TUNE="kernel:rows:batch,..." (e.g. "1:16:4,2:16:4") overrides the defaults at first use: experiments only.
static void tune_from_env()
{
    static bool done = false;
    if (done) return;
    done = true;
    const char* e = getenv("TL_TUNE");
    if (!e) return;
    int k, r, b;
    while (*e) {
        if (sscanf(e, "%d:%d:%d", &k, &r, &b) == 3) tl_set_tuning(k, r, b);
        while (*e && *e != ',') ++e;
        if (*e == ',') ++e;
    }
}

// Stencil kernels: tall tiles (the two extra rows of the operand per tile are re-read from L2), sized so that the whole
// grid is resident at once with some slack: about 6.75 tiles per SM (999 tiles on 148 SMs).  Measured in the resident
// loop at 4000 x 4000 (profiles/pw_rows_r02.txt): 62-66 rows (992-1040 tiles) 0.2446-0.2450 ms per iteration, 55 rows
// (1168 tiles, just under 8 per SM: round 1's choice) 0.2528, 48-52 rows 0.253 (0.29 on two ranks), 80 rows 0.2532.
static int tall_tile_rows(const tl_chunk* c)
{
    const int colb = (c->nx + TL_TILE_COLS - 1) / TL_TILE_COLS;
    int rowblocks = 999 / colb;
    if (rowblocks < 1) rowblocks = 1;
    int rows = (c->ny + rowblocks - 1) / rowblocks;
    if (rows < 8) rows = 8;
    if (rows > 128) rows = 128;
    return rows;
}
int tlk_tile_rows(const tl_chunk* c, int kernel)
{
    tune_from_env();
    if (g_rows[kernel] > 0) return g_rows[kernel];
    if (kernel == TUNE_UR || kernel == TUNE_P) return 8;
    return tall_tile_rows(c);
}
dim3 tlk_hot_grid(const tl_chunk* c, int rows)
{
    return dim3((c->nx + TL_TILE_COLS - 1) / TL_TILE_COLS, (c->ny + rows - 1) / rows);
}
int tlk_hot_check(const tl_chunk* c, dim3 grid)
{
    if ((long)grid.x * grid.y > c->partial_cap) {
        tl_set_error("partials capacity exceeded");
        return TL_ERR_ARG;
    }
    return TL_OK;
}

// cg.cpp:137-195 cg_calc_w:  w = A p (5-point SMVP) fused with the p.w dot product.
// 32 B/cell of HBM traffic: read p, kx, ky; write w.  Rows j-1, j, j+1 of p and rows j, j+1 of ky
// slide through registers, so every element is requested from L2 once per tile.
template <int U, bool MULTI>
__global__ void __launch_bounds__(TL_TPB)
k_cg_calc_w(Geo g, const double* p, const double* __restrict__ kx, const double* __restrict__ ky,
            double* __restrict__ w, double* __restrict__ d_alphas, RedArgs ra, int mode, int rows, int rev,
            const MultiCtx mc)
{
    DevScal* S = ra.S;
    pdl_wait(); // p (and the solver scalars) come from the preceding kernels
    pdl_trigger();
    if (mode == SCAL_DEV) {
        const int conv = sld(&S->conv);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
            S->p_pending = 0;
            if (!conv) stamp(S, sld(&S->iters), 0, 0);
        }
        if (conv) return;
    }
    const HotTile t = hot_tile(g, rows, rev);
    double acc[1] = {0.0};
    if (t.v0) {
        // p is complete when this kernel's dependency wait returns and nobody writes it while the kernel runs (several
        // ranks: the neighbours' halo stores belong to calc_p, whose tail CTA hand-shook before it completed), so the
        // default L1-allocating loads are legal here: the left / right neighbours are scalar L1 hits of lines the
        // adjacent lanes' vector loads bring in.
        long i = t.i;
        const long pitch = g.pitch;
        double2 pm = ld2(p + i - pitch);
        double2 pc = ld2(p + i);
        double pl = p[i - 1], pr = p[i + 2];
        double2 kyc = ld2_ro(ky + i);
        for (int jb = t.j0; jb < t.j1; jb += U, i += U * pitch) {
            double2 pn[U], kyn[U], kxc[U];
            double kxr[U], pln[U], prn[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (jb + u < t.j1) {
                    const long iu = i + u * pitch;
                    pn[u] = ld2(p + iu + pitch);
                    kyn[u] = ld2_ro(ky + iu + pitch);
                    kxc[u] = ld2_ro(kx + iu);
                    kxr[u] = __ldg(kx + iu + 2);
                    pln[u] = p[iu + pitch - 1];
                    prn[u] = p[iu + pitch + 2];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (jb + u < t.j1) {
                    double2 wv;
                    wv.x = smvp(kxc[u].x, kxc[u].y, kyc.x, kyn[u].x, pc.x, pl, pc.y, pm.x, pn[u].x);
                    wv.y = smvp(kxc[u].y, kxr[u], kyc.y, kyn[u].y, pc.y, pc.x, pr, pm.y, pn[u].y);
                    st_pair(w + i + u * pitch, wv, t.v1);
                    acc[0] += wv.x * pc.x;
                    if (t.v1) acc[0] += wv.y * pc.y;
                    pm = pc; pc = pn[u]; kyc = kyn[u]; pl = pln[u]; pr = prn[u];
                }
            }
        }
    }
    double tot[1];
    if (grid_reduce<1>(acc, ra, t.tile, t.ntiles, tot) && threadIdx.x < 32) {
        const int it = (mode == SCAL_DEV) ? sld(&S->iters) : -1;
        if (threadIdx.x == 0) stamp(S, it, 0, 1);
        double pw = tot[0];
        if constexpr (MULTI) pw = mc_allsum_warp(mc, 0, tot[0], S); // sum_over_ranks(pw), cg_driver.c:85, over NVLink
        if (threadIdx.x == 0) {
            S->pw = pw;
            if (mode == SCAL_DEV) {
                const double alpha = S->rro / pw; // cg_driver.c:87
                S->alpha = alpha;
                d_alphas[it] = alpha;             // cg_driver.c:93
                stamp(S, it, 0, 2);
            }
        }
    }
}

const MultiCtx g_single_ctx = {1};

int tlk_cg_calc_w(tl_chunk* c, ScalMode mode, bool rev, const MultiCtx* mc, bool pdl)
{
    const int rows = tlk_tile_rows(c, TUNE_W);
    dim3 grid = tlk_hot_grid(c, rows);
    TL_TRY(tlk_hot_check(c, grid));
    RedArgs ra{c->partials, c->gpartials, c->gcount, c->partial_cap, c->gpartial_cap, c->scal};
#define LAUNCH_W(U)                                                                                          \
    if (mc && mc->num_ranks > 1)                                                                             \
        TL_CUDA(tl_launch(k_cg_calc_w<U, true>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g, c->f[TL_FIELD_P],   \
                          c->f[TL_FIELD_KX], c->f[TL_FIELD_KY], c->f[TL_FIELD_W], c->d_alphas, ra, (int)mode,    \
                          rows, rev ? 1 : 0, *mc));                                                            \
    else                                                                                                     \
        TL_CUDA(tl_launch(k_cg_calc_w<U, false>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g, c->f[TL_FIELD_P],  \
                          c->f[TL_FIELD_KX], c->f[TL_FIELD_KY], c->f[TL_FIELD_W], c->d_alphas, ra, (int)mode,    \
                          rows, rev ? 1 : 0, g_single_ctx))
    switch (g_batch[TUNE_W]) {
    case 1: LAUNCH_W(1); break;
    case 2: LAUNCH_W(2); break;
    default: LAUNCH_W(4); break;
    }
#undef LAUNCH_W
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// cg.cpp:198-254 cg_calc_ur:  u += alpha p ; r -= alpha w ; fused with the r.r reduction.
// 48 B/cell: read u, p, r, w; write u, r.
template <int U, bool MULTI>
__global__ void __launch_bounds__(TL_TPB)
k_cg_calc_ur(Geo g, double* u, double* r, const double* p, const double* w, double* __restrict__ d_betas, RedArgs ra, int mode, double alpha_imm,
             int rows, int rev, double* __restrict__ d_alphas, const MultiCtx mc, int send_r_halo)
{
    DevScal* S = ra.S;
    double alpha = alpha_imm;
    const HotTile t = hot_tile(g, rows, rev, MULTI && send_r_halo);
    const long pitch = g.pitch;
    double acc[1] = {0.0};
    long i = t.i;
    double2 uv[U], rv[U], pv[U], wv[U];
    auto load_ur = [&](int jb) { // u, r: last written by the previous calc_ur (or cg_init), two launches back
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (jb + q < t.j1) {
                uv[q] = ldp2<true>(u + i + q * pitch);
                rv[q] = ldp2<true>(r + i + q * pitch);
            }
        }
    };
    auto load_pw = [&](int jb) { // p, w: written by the kernel just before this one
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (jb + q < t.j1) {
                pv[q] = ldp2<true>(p + i + q * pitch);
                wv[q] = ldp2<true>(w + i + q * pitch);
            }
        }
    };
    auto load_batch = [&](int jb) {
        load_ur(jb);
        load_pw(jb);
    };
    // Programmatic dependent launch: this CTA may be resident while the preceding matvec kernel is still draining;
    // the first rows of u and r are requested before the wait, p and w after it.
    if (t.v0) load_ur(t.j0);
    pdl_wait();
    pdl_trigger();
    if (mode == SCAL_DEV) {
        if (sld(&S->conv)) return;
        alpha = sld(&S->alpha);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) stamp(S, sld(&S->iters), 1, 0);
    }
    if (t.v0) load_pw(t.j0);
    if (t.v0) {
        for (int jb = t.j0; jb < t.j1; jb += U, i += U * pitch) {
            if (jb != t.j0) load_batch(jb);
#pragma unroll
            for (int q = 0; q < U; ++q) {
                if (jb + q < t.j1) {
                    const long iq = i + q * pitch;
                    uv[q].x = uv[q].x + alpha * pv[q].x;
                    uv[q].y = uv[q].y + alpha * pv[q].y;
                    rv[q].x = rv[q].x - alpha * wv[q].x;
                    rv[q].y = rv[q].y - alpha * wv[q].y;
                    st_pair(u + iq, uv[q], t.v1);
                    st_pair(r + iq, rv[q], t.v1);
                    acc[0] += rv[q].x * rv[q].x;
                    if (t.v1) acc[0] += rv[q].y * rv[q].y;
                    if constexpr (MULTI) { // fused loop: the updated r edge cells go to the neighbours' halo of r
                        if (send_r_halo) edge_remote_store<true>(g, mc, mc.nb_f, jb + q, rv[q], t);
                    }
                }
            }
        }
    }
    if constexpr (MULTI) {
        // only the tiles of the top / bottom row made remote halo stores (the columns are parked locally): visible
        // system-wide before this CTA's ticket
        if (send_r_halo && tile_sends_halo(g, mc, t, true)) __threadfence_system();
    }
    double tot[1];
    const ForwardColumns fwd{g, mc, rows, (int)gridDim.x, t.ntiles, MULTI && send_r_halo != 0};
    if (grid_reduce<1, false>(acc, ra, t.tile, t.ntiles, tot, fwd)) {
        const int it = (mode == SCAL_DEV) ? sld(&S->iters) : -1;
        if (threadIdx.x == 0) stamp(S, it, 1, 1);
        double rrn = tot[0];
        if constexpr (MULTI) {
            // warp 0: sum_over_ranks(rrn), cg_driver.c:104, over NVLink; warp 1: fused loop, r's halo handshake (every CTA
            // fenced its halo stores before its ticket)
            if (threadIdx.x < 32) rrn = mc_allsum_warp(mc, 1, tot[0], S);
            else if (send_r_halo && threadIdx.x < 64) { // warp 1: every group has forwarded and fenced its columns
                mc_halo_handshake(mc, mc.nb_f, S, threadIdx.x - 32);
                __syncwarp();
                if (threadIdx.x == 32) stamp(S, it, 1, 3);
            }
        }
        if (threadIdx.x != 0) return;
        S->rrn = rrn;
        if (mode == SCAL_DEV) {
            const double beta = rrn / S->rro; // cg_driver.c:106
            S->beta = beta;
            d_betas[it] = beta;               // cg_driver.c:111
            S->error = rrn;                   // cg_driver.c:122-123
            S->rro = rrn;
            S->iters = it + 1;
            S->p_pending = 1;
            if (conv_test(S, rrn) || it + 1 >= S->max_iters) S->conv = 1; // cg_driver.c:18,24; cheby_driver.c:70
            // cheby_driver.c:30-32 / ppcg_driver.c:27-29: the rule that ends the CG pre-steps, seen from the iteration before
            if (S->sw_min_iters >= 0 && it + 1 > S->sw_min_iters && rrn < S->sw_thresh) S->conv = 1;
            stamp(S, it, 1, 2);
        }
    }
}

int tlk_cg_calc_ur(tl_chunk* c, ScalMode mode, double alpha, bool rev, const MultiCtx* mc, bool send_r_halo, bool pdl)
{
    const int rows = tlk_tile_rows(c, TUNE_UR);
    dim3 grid = tlk_hot_grid(c, rows);
    TL_TRY(tlk_hot_check(c, grid));
    RedArgs ra{c->partials, c->gpartials, c->gcount, c->partial_cap, c->gpartial_cap, c->scal};
    MultiCtx m = g_single_ctx;
    if (mc && mc->num_ranks > 1) {
        m = *mc;
        tlk_set_travelling_field(c, &m, c->f[TL_FIELD_R]); // fused loop: r's edge cells go to the neighbours' halo of r
    }
    mc = &m;
#define LAUNCH_UR(U)                                                                                       \
    if (mc && mc->num_ranks > 1)                                                                           \
        TL_CUDA(tl_launch(k_cg_calc_ur<U, true>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g, c->f[TL_FIELD_U], \
                          c->f[TL_FIELD_R], c->f[TL_FIELD_P], c->f[TL_FIELD_W], c->d_betas, ra, (int)mode,      \
                          alpha, rows, rev ? 1 : 0, c->d_alphas, *mc, send_r_halo ? 1 : 0));                  \
    else                                                                                                   \
        TL_CUDA(tl_launch(k_cg_calc_ur<U, false>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g,                 \
                          c->f[TL_FIELD_U], c->f[TL_FIELD_R], c->f[TL_FIELD_P], c->f[TL_FIELD_W], c->d_betas,  \
                          ra, (int)mode, alpha, rows, rev ? 1 : 0, c->d_alphas, g_single_ctx, 0))
    switch (g_batch[TUNE_UR]) {
    case 1: LAUNCH_UR(1); break;
    case 2: LAUNCH_UR(2); break;
    default: LAUNCH_UR(4); break;
    }
#undef LAUNCH_UR
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// Depth-1 reflective halo of p written by the thread that owns the edge cell (local_halos.cpp, depth 1,
// external faces only; corners end up = double reflection exactly as L/R-then-T/B produces them).
__device__ __forceinline__ void p_halo_store(const Geo& g, double* p, long i, int jj, double2 pv, const HotTile& t,
                                             int halo_mask)
{
    const int last = g.x - g.hd - 1;
    const bool left_edge = (halo_mask & 1) && t.kk == g.hd;
    const bool right0 = (halo_mask & 2) && t.kk == last, right1 = (halo_mask & 2) && t.kk + 1 == last;
    const bool bot = (halo_mask & 4) && jj == g.hd, top = (halo_mask & 8) && jj == g.y - g.hd - 1;
    if (left_edge) p[i - 1] = pv.x;
    if (right0) p[i + 1] = pv.x;
    if (right1) p[i + 2] = pv.y;
    if (bot) {
        st_pair(p + i - g.pitch, pv, t.v1);
        if (left_edge) p[i - g.pitch - 1] = pv.x;
        if (right0) p[i - g.pitch + 1] = pv.x;
        if (right1) p[i - g.pitch + 2] = pv.y;
    }
    if (top) {
        st_pair(p + i + g.pitch, pv, t.v1);
        if (left_edge) p[i + g.pitch - 1] = pv.x;
        if (right0) p[i + g.pitch + 1] = pv.x;
        if (right1) p[i + g.pitch + 2] = pv.y;
    }
}

// cg.cpp:257-281 cg_calc_p:  p = beta p + r.  24 B/cell.
// halo_mask != 0: the CTAs that own chunk-edge cells also write the depth-1 reflective halo of p on
// the external faces in the mask, so the resident loop needs no separate halo launches.
template <int U, bool MULTI>
__global__ void __launch_bounds__(TL_TPB)
k_cg_calc_p(Geo g, double* p, const double* r, DevScal* S, int mode, double beta_imm,
            int rows, int rev, int halo_mask, unsigned int* gcount, const MultiCtx mc)
{
    constexpr bool multi = MULTI;
    double beta = beta_imm;
    const HotTile t = hot_tile(g, rows, rev, MULTI);
    const long pitch = g.pitch;
    long i = t.i;
    int it = -1;
    double2 pv[U], rv[U];
    auto load_p = [&](int jb) { // p: last written by the previous calc_p / calc_pw, at least two launches back
#pragma unroll
        for (int q = 0; q < U; ++q)
            if (jb + q < t.j1) pv[q] = ldp2<true>(p + i + q * pitch);
    };
    auto load_r = [&](int jb) { // r: written by the calc_ur just before this kernel
#pragma unroll
        for (int q = 0; q < U; ++q)
            if (jb + q < t.j1) rv[q] = ldp2<true>(r + i + q * pitch);
    };
    auto load_batch = [&](int jb) {
        load_p(jb);
        load_r(jb);
    };
    if (t.v0) load_p(t.j0); // programmatic dependent launch: requested while calc_ur may still be draining
    pdl_wait();
    pdl_trigger();
    if (mode == SCAL_DEV) {
        if (!sld(&S->p_pending)) return;
        beta = sld(&S->beta);
        it = sld(&S->iters) - 1; // calc_ur of this iteration has already counted it
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) stamp(S, it, 2, 0);
    }
    if (!MULTI && !t.v0) return;
    if (t.v0) load_r(t.j0);
    if (t.v0) {
        const bool edge_tile = (halo_mask || multi) && (t.j0 == g.hd || t.j1 == g.y - g.hd || t.tile % gridDim.x == 0 ||
                                                        t.tile % gridDim.x == gridDim.x - 1);
        for (int jb = t.j0; jb < t.j1; jb += U, i += U * pitch) {
            if (jb != t.j0) load_batch(jb);
#pragma unroll
            for (int q = 0; q < U; ++q) {
                if (jb + q < t.j1) {
                    pv[q].x = beta * pv[q].x + rv[q].x;
                    pv[q].y = beta * pv[q].y + rv[q].y;
                    st_pair(p + i + q * pitch, pv[q], t.v1);
                }
            }
            if (edge_tile) {
#pragma unroll
                for (int q = 0; q < U; ++q)
                    if (jb + q < t.j1) {
                        if (halo_mask) p_halo_store(g, p, i + q * pitch, jb + q, pv[q], t, halo_mask);
                        if (multi) edge_remote_store(g, mc, mc.nb_f, jb + q, pv[q], t);
                    }
            }
        }
    }
    if constexpr (MULTI) {
        // Every CTA makes its remote halo stores visible system-wide before taking its ticket; the last CTA to finish
        // releases the neighbours' per-face flags and acquires its own: when this kernel completes, the halo of p that
        // the next matvec reads is in place (what halo_update_driver.c:22 provides in the reference).
        if (tile_sends_halo(g, mc, t)) __threadfence_system();
        if (grid_last_cta(gcount, &S->counter[1], t.tile, t.ntiles, NoGroupHook()) && threadIdx.x < 32) {
            mc_halo_handshake(mc, mc.nb_f, S, threadIdx.x);
            __syncwarp();
            if (threadIdx.x == 0) stamp(S, it, 2, 3);
        }
    }
}

int tlk_external_mask(const tl_chunk* c)
{
    int mask = 0;
    if (c->nb[TL_FACE_LEFT] == TL_EXTERNAL_FACE) mask |= 1;
    if (c->nb[TL_FACE_RIGHT] == TL_EXTERNAL_FACE) mask |= 2;
    if (c->nb[TL_FACE_BOTTOM] == TL_EXTERNAL_FACE) mask |= 4;
    if (c->nb[TL_FACE_TOP] == TL_EXTERNAL_FACE) mask |= 8;
    return mask;
}

void tlk_set_travelling_field(const tl_chunk* c, MultiCtx* mc, const double* buf)
{
    const size_t slot = (size_t)(buf - c->slab) / c->field_elems;
    for (int f = 0; f < 4; ++f)
        mc->nb_f[f] = c->nb_slab[f] ? c->nb_slab[f] + slot * c->nb_field_elems[f] : nullptr;
}

int tlk_cg_calc_p(tl_chunk* c, ScalMode mode, double beta, bool rev, bool fuse_halo, const MultiCtx* mc, bool pdl)
{
    const int rows = tlk_tile_rows(c, TUNE_P);
    dim3 grid = tlk_hot_grid(c, rows);
    const int mask = fuse_halo ? tlk_external_mask(c) : 0;
    MultiCtx m = g_single_ctx;
    if (mc && mc->num_ranks > 1) {
        m = *mc;
        tlk_set_travelling_field(c, &m, c->f[TL_FIELD_P]); // p's edge cells go to the neighbours' halo of p
    }
#define LAUNCH_P(U)                                                                                         \
    if (m.num_ranks > 1)                                                                                    \
        TL_CUDA(tl_launch(k_cg_calc_p<U, true>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g, c->f[TL_FIELD_P],  \
                          c->f[TL_FIELD_R], c->scal, (int)mode, beta, rows, rev ? 1 : 0, mask, c->gcount, m));  \
    else                                                                                                    \
        TL_CUDA(tl_launch(k_cg_calc_p<U, false>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g, c->f[TL_FIELD_P], \
                          c->f[TL_FIELD_R], c->scal, (int)mode, beta, rows, rev ? 1 : 0, mask, c->gcount, m))
    switch (g_batch[TUNE_P]) {
    case 1: LAUNCH_P(1); break;
    case 2: LAUNCH_P(2); break;
    default: LAUNCH_P(4); break;
    }
#undef LAUNCH_P
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}

// Fused  p = beta p + r  (cg.cpp:257-281, of iteration t-1)  +  w = A p ; p.w  (cg.cpp:137-195, of
// iteration t).  48 B/cell instead of 24 + 32: read p, r, kx, ky; write p, w.
// Each thread updates its own two cells of p; the left/right neighbours' updated values come from the
// adjacent lanes by warp shuffle, only the two edge lanes of a warp recompute them from (p, r) of the
// neighbouring cell -- bit-identical to what the owning thread stores.  Rows j-1 and j+1 of the
// updated p slide through registers (the tile's first/last row recomputes the row outside the tile).
// Reflective boundaries on external faces are applied by index mirroring, so no halo of p is read
// there; on internal faces the halo of p (old) and r must have been exchanged.
// Old p is read by neighbouring tiles while this tile overwrites it, so the update is double
// buffered: reads come from p_in, writes go to p_out (the chunk's P and P2 buffers swap roles).
template <int U, bool MULTI>
__global__ void __launch_bounds__(TL_TPB)
k_cg_calc_pw(Geo g, const double* p_in, double* __restrict__ p_out, const double* r,
             const double* __restrict__ kx, const double* __restrict__ ky, double* __restrict__ w,
             double* __restrict__ d_alphas, double* __restrict__ d_betas, RedArgs ra, int rows, int rev, int ext_mask,
             const MultiCtx mc)
{
    DevScal* S = ra.S;
    pdl_wait(); // r comes from the calc_ur just before this kernel; w may not be overwritten while it still reads it
    pdl_trigger();
    // Same head on one rank and on several: beta, the convergence flag and (multi-rank) the neighbours' halo of r were
    // all settled by the tail CTA of the preceding calc_ur.
    if (sld(&S->conv)) return;
    const double beta = sld(&S->beta);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) stamp(S, sld(&S->iters), 0, 0);
    const HotTile t = hot_tile(g, rows, rev);
    const long pitch = g.pitch;
    const int lane = threadIdx.x & 31;
    const int klo = g.hd, khi = g.x - g.hd - 1, jlo = g.hd, jhi = g.y - g.hd - 1;
    // which neighbours are mirrored (external face at the chunk edge) rather than read
    const bool mir_l = (ext_mask & 1) && t.kk == klo;
    const bool mir_r1 = (ext_mask & 2) && t.kk + 1 == khi; // my cell 1 is the last column
    const bool mir_r0 = (ext_mask & 2) && t.kk == khi;     // my cell 0 is the last column (cell 1 invalid)
    const bool load_l = t.v0 && !mir_l && lane == 0;
    // the right neighbour of cell 1 lives in lane+1 unless I am lane 31 or that lane is past the row end
    const bool load_r = t.v1 && !mir_r1 && (lane == 31 || t.kk + 2 > khi);
    // internal faces (multi-rank): the ring values I compute for the halo cells are also the next
    // iteration's old p there, so they are written to p_out's halo
    const bool halo_l = MULTI && t.v0 && !(ext_mask & 1) && t.kk == klo;
    const bool halo_r = MULTI && !(ext_mask & 2) && ((t.v1 && t.kk + 1 == khi) || (t.v0 && !t.v1 && t.kk == khi));
    // Raw (p, r) pairs: every global load of a row is issued in the row's LOAD phase -- also the two scalar pairs the
    // edge lanes of a warp need for their outer neighbours -- and combined (beta p + r) in the compute phase.  (When the
    // scalar loads sat behind the shuffles the compiler scheduled them after the row's vector loads had returned:
    // two dependent memory latencies per row, +20 % on the multi-rank instantiation.)
    struct Raw2 { double2 p, r; };
    struct Side { double pl, rl, pr, rr; };
    auto load2 = [&](long i) { return Raw2{ldp2<true>(p_in + i), ldp2<true>(r + i)}; };
    auto comb2 = [&](const Raw2& q) { return make_double2(beta * q.p.x + q.r.x, beta * q.p.y + q.r.y); };
    auto load_sides = [&](long i) {
        Side q = {0.0, 0.0, 0.0, 0.0};
        if (load_l) {
            q.pl = ldp1<true>(p_in + i - 1);
            q.rl = ldp1<true>(r + i - 1);
        }
        if (load_r) {
            q.pr = ldp1<true>(p_in + i + 2);
            q.rr = ldp1<true>(r + i + 2);
        }
        return q;
    };
    // left / right neighbours of an updated row held as double2 `c` in every lane (called by all lanes of the warp)
    auto sides = [&](double2 c, const Side& q, double& l, double& rr) {
        const double sl = __shfl_up_sync(0xffffffffu, c.y, 1);
        const double sr = __shfl_down_sync(0xffffffffu, c.x, 1);
        l = mir_l ? c.x : sl;
        if (load_l) l = beta * q.pl + q.rl;
        rr = mir_r1 ? c.y : sr;
        if (load_r) rr = beta * q.pr + q.rr;
    };
    double acc[1] = {0.0};
    long i = t.i;
    double2 pm = make_double2(0.0, 0.0), pc = pm, kyc = pm;
    double pl = 0.0, pr = 0.0;
    {
        const bool mir_b = (ext_mask & 4) && t.j0 == jlo;
        Raw2 rc = {pm, pm}, rm = {pm, pm};
        Side sc = {0.0, 0.0, 0.0, 0.0};
        if (t.v0) {
            rc = load2(i);
            if (!mir_b) rm = load2(i - pitch);
            sc = load_sides(i);
            kyc = ld2_ro(ky + i);
        }
        pc = comb2(rc);
        pm = mir_b ? pc : comb2(rm);
        if (MULTI && t.v0 && !(ext_mask & 4) && t.j0 == jlo) st_pair(p_out + i - pitch, pm, t.v1); // bottom halo row
        sides(pc, sc, pl, pr);
    }
#pragma unroll 1
    for (int jb = t.j0; jb < t.j1; jb += U, i += U * pitch) {
        Raw2 rn[U];
        Side sn[U];
        double2 kyn[U], kxc[U];
        double kxr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            rn[u].p = rn[u].r = make_double2(0.0, 0.0);
            sn[u] = Side{0.0, 0.0, 0.0, 0.0};
            if (t.v0 && jb + u < t.j1) {
                const long iu = i + u * pitch;
                const bool mir_t = (ext_mask & 8) && (jb + u == jhi);
                if (!mir_t) {
                    rn[u] = load2(iu + pitch);
                    sn[u] = load_sides(iu + pitch);
                }
                kyn[u] = ld2_ro(ky + iu + pitch);
                kxc[u] = ld2_ro(kx + iu);
                kxr[u] = __ldg(kx + iu + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (jb + u < t.j1) { // warp-uniform
                const long iu = i + u * pitch;
                const bool mir_t = (ext_mask & 8) && (jb + u == jhi);
                const double2 pn = mir_t ? pc : comb2(rn[u]);
                double pln, prn;
                sides(pn, sn[u], pln, prn); // (the mirrored top row's sides are never used: it is the tile's last row)
                if (t.v0) {
                    double2 wv;
                    const double pr0 = mir_r0 ? pc.x : pc.y; // right neighbour of cell 0
                    wv.x = smvp(kxc[u].x, kxc[u].y, kyc.x, kyn[u].x, pc.x, pl, pr0, pm.x, pn.x);
                    wv.y = smvp(kxc[u].y, kxr[u], kyc.y, kyn[u].y, pc.y, pc.x, pr, pm.y, pn.y);
                    st_pair(w + iu, wv, t.v1);
                    st_pair(p_out + iu, pc, t.v1);
                    if constexpr (MULTI) {
                        if (halo_l) p_out[iu - 1] = pl;
                        if (halo_r) {
                            if (t.v1) p_out[iu + 2] = pr;
                            else p_out[iu + 1] = pc.y;
                        }
                        if (!(ext_mask & 8) && jb + u == jhi) st_pair(p_out + iu + pitch, pn, t.v1); // top halo row
                    }
                    acc[0] += wv.x * pc.x;
                    if (t.v1) acc[0] += wv.y * pc.y;
                    pm = pc; pc = pn; kyc = kyn[u]; pl = pln; pr = prn;
                }
            }
        }
    }
    double tot[1];
    if (grid_reduce<1>(acc, ra, t.tile, t.ntiles, tot) && threadIdx.x < 32) {
        const int it = sld(&S->iters);
        if (threadIdx.x == 0) stamp(S, it, 0, 1);
        double pw = tot[0];
        if constexpr (MULTI) pw = mc_allsum_warp(mc, 0, tot[0], S); // sum_over_ranks(pw), cg_driver.c:85, over NVLink
        if (threadIdx.x == 0) {
            S->pw = pw;
            const double alpha = S->rro / pw; // cg_driver.c:87
            S->alpha = alpha;
            d_alphas[it] = alpha;             // cg_driver.c:93
            S->p_pending = 0; // the pending p update of the previous iteration has now been applied
            stamp(S, it, 0, 2);
        }
    }
}

// p_in = c->f[P], p_out = c->p2; the caller swaps them after the launch.
int tlk_cg_calc_pw(tl_chunk* c, bool rev, const MultiCtx* mc, bool pdl)
{
    const int rows = tlk_tile_rows(c, TUNE_PW);
    dim3 grid = tlk_hot_grid(c, rows);
    TL_TRY(tlk_hot_check(c, grid));
    // optional: TMA bulk-copy row pipeline, persistent CTAs (tl_bulk.cu); bit-identical results; one rank only
    if (tlk_pw_uses_bulk() && !(mc && mc->num_ranks > 1)) {
        TL_TRY(tlk_cg_calc_pw_bulk(c, rev, mc, rows, pdl));
        double* tmp = c->f[TL_FIELD_P];
        c->f[TL_FIELD_P] = c->p2;
        c->p2 = tmp;
        return TL_OK;
    }
    RedArgs ra{c->partials, c->gpartials, c->gcount, c->partial_cap, c->gpartial_cap, c->scal};
    const int mask = tlk_external_mask(c);
#define LAUNCH_PW(U)                                                                                          \
    if (mc && mc->num_ranks > 1)                                                                              \
        TL_CUDA(tl_launch(k_cg_calc_pw<U, true>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g, c->f[TL_FIELD_P],   \
                          c->p2, c->f[TL_FIELD_R], c->f[TL_FIELD_KX], c->f[TL_FIELD_KY], c->f[TL_FIELD_W],        \
                          c->d_alphas, c->d_betas, ra, rows, rev ? 1 : 0, mask, *mc));                          \
    else                                                                                                      \
        TL_CUDA(tl_launch(k_cg_calc_pw<U, false>, grid, dim3(TL_TPB), 0, c->stream, pdl, c->g, c->f[TL_FIELD_P],  \
                          c->p2, c->f[TL_FIELD_R], c->f[TL_FIELD_KX], c->f[TL_FIELD_KY], c->f[TL_FIELD_W],        \
                          c->d_alphas, c->d_betas, ra, rows, rev ? 1 : 0, mask, g_single_ctx))
    // Load batch: one row at a time for tall tiles (4000 x 4000: 65 rows; a batch of 2 costs occupancy and is 2-10 %
    // slower there), two rows for the short tiles of small, L2-resident chunks, where the row-to-row dependent latency
    // is what a tile's time consists of (2000 x 1000: 9 rows, 40.2 -> 37.6 us per iteration; 1000 x 1000: 28.0 -> 25.0;
    // profiles/pw_rows_r02.txt).  Results do not depend on the batch.
    int batch = g_batch[TUNE_PW];
    if (!g_batch_user[TUNE_PW] && rows <= 16) batch = 2;
    switch (batch) {
    case 1: LAUNCH_PW(1); break;
    case 2: LAUNCH_PW(2); break;
    default: LAUNCH_PW(4); break;
    }
#undef LAUNCH_PW
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    double* tmp = c->f[TL_FIELD_P];
    c->f[TL_FIELD_P] = c->p2;
    c->p2 = tmp;
    return TL_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused Chebyshev / PPCG iteration kernels (one HBM pass per iteration instead of two).
//
//   MODE_CHEBY  run_cheby_iterate (kernel_interface.cpp:258-271) = cheby_iterate (cheby.cpp:73-110)
//               then cheby_calc_u (:46-70):   s = A u ; r = u0 - s ; p = alpha p + beta r ; u += p
//               reads u(stencil), u0, p, kx, ky; writes r, p, u'      = 64 B/cell (unfused: 64 + 24)
//   MODE_PPCG   run_ppcg_inner_iteration (kernel_interface.cpp:314-328) = ppcg_calc_ur (ppcg.cpp:34-66)
//               then ppcg_calc_sd (:69-94):   s = A sd ; r -= s ; u += sd ; sd = alpha sd + beta r
//               reads sd(stencil), r, u, kx, ky; writes r, u, sd'     = 64 B/cell (unfused: 56 + 24)
//
// The stencil operand (u resp. sd) is updated in the same pass, so it is double buffered: reads come
// from a_in, the updated operand goes to a_out and the chunk swaps the two buffers.  Reflective
// boundaries of external faces are applied by index mirroring (what halo_update_driver's local update
// would have put in the halo), so no halo kernel runs between iterations on a single chunk; internal
// faces read the exchanged halo.  Same operations in the same order as the two reference kernels:
// results are bit-identical.  `w` (= A u) is never read in the Chebyshev phase and is not stored here.
// Several ranks (MULTI): the thread that owns an edge cell of an internal face also stores the updated operand into the
// neighbour's halo cell of ITS output buffer (both ranks swap buffers in lockstep), and the last CTA of the grid
// hand-shakes with the neighbours -- the exchange of halo_update_driver.c for {u} / {sd}, depth 1, inside the kernel.
// NORM: the kernel also returns sum r.r over the interior (all ranks), i.e. the calculate_2norm(r) that
// cheby_driver.c:128-133 / ppcg_driver.c:136-141 run after it, without another pass over r.
enum { MODE_CHEBY = 0, MODE_PPCG = 1 };
#define TL_FS_CTAS_PER_SM 5 // register cap 102: the U = 2 load batch of five streams needs ~100
template <int MODE, int U, bool MULTI, bool NORM>
__global__ void __launch_bounds__(TL_TPB, TL_FS_CTAS_PER_SM)
k_fused_stencil(Geo g, const double* __restrict__ a_in, double* __restrict__ a_out, double* __restrict__ f1,
                const double* f2_in, double* f2_out, const double* __restrict__ kx, const double* __restrict__ ky,
                double alpha, double beta, int rows, int rev, int ext_mask, RedArgs ra, const MultiCtx mc)
{
    // MODE_CHEBY: f1 = p (in/out), f2_in = u0 (read), f2_out = r (written)
    // MODE_PPCG : f1 = r (in/out), f2_in = f2_out = u (in/out)
    const HotTile t = hot_tile(g, rows, rev);
    double acc[1] = {0.0};
    if (t.v0) {
        const long pitch = g.pitch;
        const int klo = g.hd, khi = g.x - g.hd - 1, jlo = g.hd, jhi = g.y - g.hd - 1;
        const int dl = ((ext_mask & 1) && t.kk == klo) ? 0 : -1;         // left neighbour of cell 0 (mirrored: itself)
        const int dr = ((ext_mask & 2) && t.kk + 1 == khi) ? 1 : 2;      // right neighbour of cell 1
        const bool mir_r0 = (ext_mask & 2) && t.kk == khi;               // cell 0 is the last column
        [[maybe_unused]] const bool sends = MULTI && tile_sends_halo(g, mc, t);
        long i = t.i;
        const long im = ((ext_mask & 4) && t.j0 == jlo) ? i : i - pitch;
        double2 am = ld2_ro(a_in + im);
        double2 ac = ld2_ro(a_in + i);
        double al = __ldg(a_in + i + dl), ar = __ldg(a_in + i + dr);
        double2 kyc = ld2_ro(ky + i);
        for (int jb = t.j0; jb < t.j1; jb += U, i += U * pitch) {
            double2 an[U], kyn[U], kxc[U], x1[U], x2[U];
            double kxr[U], aln[U], arn[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                if (jb + q < t.j1) {
                    const long iq = i + q * pitch;
                    const long in = ((ext_mask & 8) && jb + q == jhi) ? iq : iq + pitch;
                    an[q] = ld2_ro(a_in + in);
                    aln[q] = __ldg(a_in + in + dl);
                    arn[q] = __ldg(a_in + in + dr);
                    kyn[q] = ld2_ro(ky + iq + pitch);
                    kxc[q] = ld2_ro(kx + iq);
                    kxr[q] = __ldg(kx + iq + 2);
                    x1[q] = ld2(f1 + iq);
                    x2[q] = ld2(f2_in + iq);
                }
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                if (jb + q < t.j1) {
                    const long iq = i + q * pitch;
                    double2 sv;
                    sv.x = smvp(kxc[q].x, kxc[q].y, kyc.x, kyn[q].x, ac.x, al, mir_r0 ? ac.x : ac.y, am.x, an[q].x);
                    sv.y = smvp(kxc[q].y, kxr[q], kyc.y, kyn[q].y, ac.y, ac.x, ar, am.y, an[q].y);
                    double2 o1, o2, ao, rn;
                    if (MODE == MODE_CHEBY) {
                        // x1 = p, x2 = u0:  r = u0 - s ; p = alpha p + beta r ; u' = u + p
                        o2.x = x2[q].x - sv.x;
                        o2.y = x2[q].y - sv.y;
                        o1.x = alpha * x1[q].x + beta * o2.x;
                        o1.y = alpha * x1[q].y + beta * o2.y;
                        ao.x = ac.x + o1.x;
                        ao.y = ac.y + o1.y;
                        rn = o2;
                    } else {
                        // x1 = r, x2 = u:  r -= s ; u += sd ; sd' = alpha sd + beta r
                        o1.x = x1[q].x - sv.x;
                        o1.y = x1[q].y - sv.y;
                        o2.x = x2[q].x + ac.x;
                        o2.y = x2[q].y + ac.y;
                        ao.x = alpha * ac.x + beta * o1.x;
                        ao.y = alpha * ac.y + beta * o1.y;
                        rn = o1;
                    }
                    st_pair(f1 + iq, o1, t.v1);
                    st_pair(f2_out + iq, o2, t.v1);
                    st_pair(a_out + iq, ao, t.v1);
                    if constexpr (NORM) {
                        acc[0] += rn.x * rn.x;
                        if (t.v1) acc[0] += rn.y * rn.y;
                    }
                    if constexpr (MULTI) {
                        if (sends) edge_remote_store<true>(g, mc, mc.nb_f, jb + q, ao, t);
                    }
                    am = ac; ac = an[q]; kyc = kyn[q]; al = aln[q]; ar = arn[q];
                }
            }
        }
    }
    if constexpr (MULTI) {
        if (tile_sends_halo(g, mc, t, true)) __threadfence_system(); // remote row stores visible before this CTA's ticket
    }
    const ForwardColumns fwd{g, mc, rows, (int)gridDim.x, t.ntiles, MULTI};
    if constexpr (NORM) {
        double tot[1];
        if (grid_reduce<1, false>(acc, ra, t.tile, t.ntiles, tot, fwd)) {
            double nrm = tot[0];
            if constexpr (MULTI) {
                if (threadIdx.x < 32) nrm = mc_allsum_warp(mc, 1, tot[0], ra.S); // sum_over_ranks, cheby_driver.c:132
                else if (threadIdx.x < 64) mc_halo_handshake(mc, mc.nb_f, ra.S, threadIdx.x - 32);
            }
            if (threadIdx.x == 0) ra.S->sums[0] = nrm;
        }
    } else if constexpr (MULTI) {
        if (grid_last_cta(ra.gcount, &ra.S->counter[1], t.tile, t.ntiles, fwd) && threadIdx.x < 32)
            mc_halo_handshake(mc, mc.nb_f, ra.S, threadIdx.x);
    }
}

// Runs one fused iteration and swaps the operand's two buffers (field = TL_FIELD_U or TL_FIELD_SD).
static int launch_fused_stencil(tl_chunk* c, int mode, double alpha, double beta, const MultiCtx* mc, bool norm)
{
    const int field = (mode == MODE_CHEBY) ? TL_FIELD_U : TL_FIELD_SD;
    tune_from_env();
    const int rows = tlk_tile_rows(c, TUNE_W);
    dim3 grid = tlk_hot_grid(c, rows);
    TL_TRY(tlk_hot_check(c, grid));
    const int mask = tlk_external_mask(c);
    RedArgs ra{c->partials, c->gpartials, c->gcount, c->partial_cap, c->gpartial_cap, c->scal};
    MultiCtx m = g_single_ctx;
    const bool multi = mc && mc->num_ranks > 1;
    if (multi) {
        m = *mc;
        tlk_set_travelling_field(c, &m, c->alt[field]); // the neighbours' copy of the buffer this launch writes
    }
#define LAUNCH_FS(MODE, MULTI, NORM)                                                                            \
    do {                                                                                                        \
        if (MODE == MODE_CHEBY)                                                                                 \
            k_fused_stencil<MODE_CHEBY, 2, MULTI, NORM><<<grid, TL_TPB, 0, c->stream>>>(                        \
                c->g, c->f[TL_FIELD_U], c->alt[TL_FIELD_U], c->f[TL_FIELD_P], c->f[TL_FIELD_U0], c->f[TL_FIELD_R], \
                c->f[TL_FIELD_KX], c->f[TL_FIELD_KY], alpha, beta, rows, 0, mask, ra, m);                       \
        else                                                                                                    \
            k_fused_stencil<MODE_PPCG, 2, MULTI, NORM><<<grid, TL_TPB, 0, c->stream>>>(                         \
                c->g, c->f[TL_FIELD_SD], c->alt[TL_FIELD_SD], c->f[TL_FIELD_R], c->f[TL_FIELD_U], c->f[TL_FIELD_U], \
                c->f[TL_FIELD_KX], c->f[TL_FIELD_KY], alpha, beta, rows, 0, mask, ra, m);                       \
    } while (0)
    if (multi && norm) LAUNCH_FS(mode, true, true);
    else if (multi) LAUNCH_FS(mode, true, false);
    else if (norm) LAUNCH_FS(mode, false, true);
    else LAUNCH_FS(mode, false, false);
#undef LAUNCH_FS
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    double* tmp = c->f[field];
    c->f[field] = c->alt[field];
    c->alt[field] = tmp;
    return TL_OK;
}
int tlk_cheby_fused(tl_chunk* c, double alpha, double beta, const MultiCtx* mc, bool norm)
{
    return launch_fused_stencil(c, MODE_CHEBY, alpha, beta, mc, norm);
}
int tlk_ppcg_fused(tl_chunk* c, double alpha, double beta, const MultiCtx* mc, bool norm)
{
    return launch_fused_stencil(c, MODE_PPCG, alpha, beta, mc, norm);
}

// Puts a double-buffered field back into its slab position (copy if it currently lives in the
// alternate buffer), so that slab-relative peer mappings and later phases see it where they expect.
int tlk_field_home(tl_chunk* c, int field)
{
    double* home = c->slab + (size_t)field * c->field_elems;
    if (c->f[field] == home) return TL_OK;
    TL_CUDA(cudaMemcpyAsync(home, c->f[field], c->field_elems * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if (field == TL_FIELD_P) c->p2 = c->f[field];
    else c->alt[field] = c->f[field];
    c->f[field] = home;
    return TL_OK;
}

// ---------------------------------------------------------------------------------------------
// Chebyshev (cheby.cpp), PPCG (ppcg.cpp), Jacobi (jacobi.cpp), shared solver kernels
// ---------------------------------------------------------------------------------------------
// cheby.cpp:7-43
int tlk_cheby_init(tl_chunk* c, double theta)
{
    double *p = c->f[TL_FIELD_P], *r = c->f[TL_FIELD_R], *w = c->f[TL_FIELD_W];
    const double *u = c->f[TL_FIELD_U], *u0 = c->f[TL_FIELD_U0], *kx = c->f[TL_FIELD_KX], *ky = c->f[TL_FIELD_KY];
    const int pitch = c->g.pitch;
    return launch_generic<0>(c, INTERIOR_RANGE(c),
                             [=] __device__(long i, int, int, double*) {
                                 const double s = smvp(kx[i], kx[i + 1], ky[i], ky[i + pitch], u[i], u[i - 1],
                                                       u[i + 1], u[i - pitch], u[i + pitch]);
                                 w[i] = s;
                                 const double rv = u0[i] - s;
                                 r[i] = rv;
                                 p[i] = rv / theta;
                             },
                             NoFin());
}
// cheby.cpp:73-110
int tlk_cheby_iterate(tl_chunk* c, double alpha, double beta)
{
    double *p = c->f[TL_FIELD_P], *r = c->f[TL_FIELD_R], *w = c->f[TL_FIELD_W];
    const double *u = c->f[TL_FIELD_U], *u0 = c->f[TL_FIELD_U0], *kx = c->f[TL_FIELD_KX], *ky = c->f[TL_FIELD_KY];
    const int pitch = c->g.pitch;
    return launch_generic<0>(c, INTERIOR_RANGE(c),
                             [=] __device__(long i, int, int, double*) {
                                 const double s = smvp(kx[i], kx[i + 1], ky[i], ky[i + pitch], u[i], u[i - 1],
                                                       u[i + 1], u[i - pitch], u[i + pitch]);
                                 w[i] = s;
                                 const double rv = u0[i] - s;
                                 r[i] = rv;
                                 p[i] = alpha * p[i] + beta * rv;
                             },
                             NoFin());
}
// cheby.cpp:46-70
int tlk_cheby_calc_u(tl_chunk* c)
{
    double* u = c->f[TL_FIELD_U];
    const double* p = c->f[TL_FIELD_P];
    return launch_vec<0, 4, V2>(c, INTERIOR_RANGE(c),
                                [=] __device__(long i, int, int) { return V2{ld2(u + i), ld2(p + i)}; },
                                [=] __device__(const V2& v, long i, int, int, bool v0, bool v1, double*) {
                                    st_mask(u + i, make_double2(v.a.x + v.b.x, v.a.y + v.b.y), v0, v1);
                                },
                                NoFin());
}
// ppcg.cpp:7-31
int tlk_ppcg_init(tl_chunk* c, double theta)
{
    double* sd = c->f[TL_FIELD_SD];
    const double* r = c->f[TL_FIELD_R];
    return launch_vec<0, 4, V1>(c, INTERIOR_RANGE(c), [=] __device__(long i, int, int) { return V1{ld2(r + i)}; },
                                [=] __device__(const V1& v, long i, int, int, bool v0, bool v1, double*) {
                                    st_mask(sd + i, make_double2(v.a.x / theta, v.a.y / theta), v0, v1);
                                },
                                NoFin());
}
// ppcg.cpp:34-66
int tlk_ppcg_calc_ur(tl_chunk* c)
{
    double *r = c->f[TL_FIELD_R], *u = c->f[TL_FIELD_U];
    const double *sd = c->f[TL_FIELD_SD], *kx = c->f[TL_FIELD_KX], *ky = c->f[TL_FIELD_KY];
    const int pitch = c->g.pitch;
    return launch_generic<0>(c, INTERIOR_RANGE(c),
                             [=] __device__(long i, int, int, double*) {
                                 const double s = smvp(kx[i], kx[i + 1], ky[i], ky[i + pitch], sd[i], sd[i - 1],
                                                       sd[i + 1], sd[i - pitch], sd[i + pitch]);
                                 r[i] -= s;
                                 u[i] += sd[i];
                             },
                             NoFin());
}
// ppcg.cpp:69-94
int tlk_ppcg_calc_sd(tl_chunk* c, double alpha, double beta)
{
    double* sd = c->f[TL_FIELD_SD];
    const double* r = c->f[TL_FIELD_R];
    return launch_vec<0, 4, V2>(c, INTERIOR_RANGE(c),
                                [=] __device__(long i, int, int) { return V2{ld2(sd + i), ld2(r + i)}; },
                                [=] __device__(const V2& v, long i, int, int, bool v0, bool v1, double*) {
                                    st_mask(sd + i, make_double2(alpha * v.a.x + beta * v.b.x, alpha * v.a.y + beta * v.b.y),
                                            v0, v1);
                                },
                                NoFin());
}
// jacobi.cpp:7-54
int tlk_jacobi_init(tl_chunk* c, int coefficient, double rx, double ry)
{
    double *u = c->f[TL_FIELD_U], *u0 = c->f[TL_FIELD_U0], *kx = c->f[TL_FIELD_KX], *ky = c->f[TL_FIELD_KY];
    const double *den = c->f[TL_FIELD_DENSITY], *en = c->f[TL_FIELD_ENERGY1];
    const int x = c->g.x, y = c->g.y, hd = c->g.hd, pitch = c->g.pitch;
    return launch_generic<0>(
        c, ALL_RANGE(c),
        [=] __device__(long i, int jj, int kk, double*) {
            if (kk > 0 && kk < x - 1 && jj > 0 && jj < y - 1) {
                const double v = en[i] * den[i];
                u0[i] = v;
                u[i] = v;
            }
            if (jj >= hd && jj < y - 1 && kk >= hd && kk < x - 1) {
                const bool cnd = (coefficient == TL_CONDUCTIVITY);
                const double dc = cnd ? den[i] : 1.0 / den[i];
                const double dl = cnd ? den[i - 1] : 1.0 / den[i - 1];
                const double dd = cnd ? den[i - pitch] : 1.0 / den[i - pitch];
                kx[i] = rx * (dl + dc) / (2.0 * dl * dc);
                ky[i] = ry * (dd + dc) / (2.0 * dd * dc);
            }
        },
        NoFin());
}
// kernel_interface.cpp:287-300: jacobi_copy_u (all cells) then jacobi_iterate (jacobi.cpp:57-117)
struct VJac {
    double2 u0, rc, rd, ru, kx, kyc, kyu;
    double rl, rr, kxr;
};
int tlk_jacobi_iterate(tl_chunk* c)
{
    TL_TRY(tlk_copy_field(c, TL_FIELD_R, TL_FIELD_U, false));
    double* u = c->f[TL_FIELD_U];
    const double *u0 = c->f[TL_FIELD_U0], *r = c->f[TL_FIELD_R], *kx = c->f[TL_FIELD_KX], *ky = c->f[TL_FIELD_KY];
    const int pitch = c->g.pitch;
    return launch_vec<1, 2, VJac>(
        c, INTERIOR_RANGE(c),
        [=] __device__(long i, int, int) {
            VJac v;
            v.u0 = ld2(u0 + i);
            v.rc = ld2(r + i);
            v.rd = ld2(r + i - pitch);
            v.ru = ld2(r + i + pitch);
            v.kx = ld2(kx + i);
            v.kyc = ld2(ky + i);
            v.kyu = ld2(ky + i + pitch);
            v.rl = r[i - 1];
            v.rr = r[i + 2];
            v.kxr = kx[i + 2];
            return v;
        },
        [=] __device__(const VJac& v, long i, int, int, bool v0, bool v1, double* acc) {
            double2 o;
            o.x = (v.u0.x + (v.kx.y * v.rc.y + v.kx.x * v.rl) + (v.kyu.x * v.ru.x + v.kyc.x * v.rd.x)) /
                  (1.0 + (v.kx.x + v.kx.y) + (v.kyc.x + v.kyu.x));
            o.y = (v.u0.y + (v.kxr * v.rr + v.kx.y * v.rc.x) + (v.kyu.y * v.ru.y + v.kyc.y * v.rd.y)) /
                  (1.0 + (v.kx.y + v.kxr) + (v.kyc.y + v.kyu.y));
            st_mask(u + i, o, v0, v1);
            if (v0) acc[0] += fabs(o.x - v.rc.x);
            if (v1) acc[0] += fabs(o.y - v.rc.y);
        },
        [=] __device__(const double* t, DevScal* S) { S->sums[0] = t[0]; });
}
// solver_methods.cpp:34-65
int tlk_calculate_residual(tl_chunk* c)
{
    double* r = c->f[TL_FIELD_R];
    const double *u = c->f[TL_FIELD_U], *u0 = c->f[TL_FIELD_U0], *kx = c->f[TL_FIELD_KX], *ky = c->f[TL_FIELD_KY];
    const int pitch = c->g.pitch;
    return launch_generic<0>(c, INTERIOR_RANGE(c),
                             [=] __device__(long i, int, int, double*) {
                                 const double s = smvp(kx[i], kx[i + 1], ky[i], ky[i + pitch], u[i], u[i - 1],
                                                       u[i + 1], u[i - pitch], u[i + pitch]);
                                 r[i] = u0[i] - s;
                             },
                             NoFin());
}
// solver_methods.cpp:68-117
int tlk_calculate_2norm(tl_chunk* c, int field)
{
    const double* b = c->f[field];
    return launch_vec<1, 4, V1>(c, INTERIOR_RANGE(c), [=] __device__(long i, int, int) { return V1{ld2(b + i)}; },
                                [=] __device__(const V1& v, long, int, int, bool v0, bool v1, double* acc) {
                                    if (v0) acc[0] += v.a.x * v.a.x;
                                    if (v1) acc[0] += v.a.y * v.a.y;
                                },
                                [=] __device__(const double* t, DevScal* S) { S->sums[0] = t[0]; });
}
// solver_methods.cpp:120-145
int tlk_finalise(tl_chunk* c)
{
    double* en = c->f[TL_FIELD_ENERGY1];
    const double *u = c->f[TL_FIELD_U], *den = c->f[TL_FIELD_DENSITY];
    return launch_vec<0, 4, V2>(c, INTERIOR_RANGE(c),
                                [=] __device__(long i, int, int) { return V2{ld2(u + i), ld2(den + i)}; },
                                [=] __device__(const V2& v, long i, int, int, bool v0, bool v1, double*) {
                                    st_mask(en + i, make_double2(v.a.x / v.b.x, v.a.y / v.b.y), v0, v1);
                                },
                                NoFin());
}
