// tl_device.cuh -- device-side helpers shared by the kernel translation units (tl_kernels.cu, tl_bulk.cu):
// vector loads/stores, system-scope flag primitives of the NVLink paths, the deterministic single-pass grid
// reduction, the reference SMVP association, the multi-rank slot/flag helpers and the hot-kernel tile geometry.
#pragma once
#include <utility>
#include "tl_internal.h"

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The kernels of the resident CG loop are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the CTAs of kernel N+1 become resident as the CTAs of kernel N
// retire, run their prologue (barrier set-up, index arithmetic, loads of data that kernel N does not write) and block in
// pdl_wait() until kernel N has completed and its writes are visible.  Every such kernel calls pdl_trigger() only AFTER
// its own pdl_wait(), so when a kernel's CTAs run, everything older than its immediate predecessor is complete: data
// written two launches ago may be read before the wait.  Without the launch attribute both calls are no-ops.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute when `pdl` is set.
// (Measured, profiles/pdl_r02.txt: forcing one shared-memory carve-out on all loop kernels so that their CTAs can share
// SMs costs the register-staged kernels 15-30 %: with the L1 shrunk to its minimum they cannot keep enough loads in
// flight.  The kernels therefore keep the driver's default carve-out.)
template <class... KArgs, class... Args>
static inline cudaError_t tl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    bool pdl, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 ld2_ro(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
__device__ __forceinline__ void st_pair(double* p, double2 v, bool v1)
{
    if (v1) st2(p, v);
    else p[0] = v.x;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// system-scope flag primitives of the NVLink paths (halo exchange, resident multi-rank CG loop)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// Spin until *flag >= want.  Bounded (~10 s; later waits bail at once) so that a lost peer cannot hang the GPU.  A timeout
// is FATAL for the solve: the error mark goes to DevScal.pad and to the host-mapped error word, and the convergence stamp is
// set so that every later launch of the resident loop returns at once; the host reports TL_ERR_COMMS at its next
// synchronisation point (tl_check_peer_timeout).
__device__ __forceinline__ void spin_flag(const unsigned long long* flag, unsigned long long want, DevScal* S,
                                          unsigned long long site = 0)
{
    const long long t0 = clock64();
    unsigned long long seen;
    while ((seen = ld_acquire_sys(flag)) < want) {
        if (*(volatile unsigned int*)&S->pad == 0xdeadu) break;
        if (clock64() - t0 > 20000000000LL) { // ~10 s
            if (atomicCAS(&S->pad, 0u, 0xdeadu) == 0u) {
                S->dbg[0] = site;
                S->dbg[1] = want;
                S->dbg[2] = seen;
                S->dbg[3] = (unsigned long long)(blockIdx.y * gridDim.x + blockIdx.x);
                S->conv = 1; // every later launch of the resident loop returns at once: nothing computes on unknown data
                if (S->err_host) *(volatile unsigned int*)S->err_host = 0xdeadu;
                __threadfence_system();
            }
            break;
        }
        __nanosleep(64);
    }
}
// Loads of fields that change inside the solver loop (p, r, u, w).  COH = true: ld.global.cg, served by L2 and never
// allocating in L1.  The streaming kernels use it throughout -- every byte is used once, so L1 brings nothing, and
// the rows they request BEFORE griddepcontrol.wait (programmatic dependent launch: the CTA is resident while the
// predecessor drains) must not leave lines in L1 that a later launch could hit.  After the wait a kernel may use the
// default L1-allocating loads for data nobody writes while it runs (cg_calc_w does, for its neighbour cells): the wait
// makes the predecessors' writes visible, and on several ranks the neighbours' halo stores belong to the predecessor,
// whose tail CTA hand-shook before it completed.  COH = false: the read-only path (ld.global.nc) of the constant
// coefficient fields kx, ky.
template <bool COH>
__device__ __forceinline__ double2 ldp2(const double* p)
{
    if constexpr (COH) return __ldcg(reinterpret_cast<const double2*>(p));
    else return __ldg(reinterpret_cast<const double2*>(p));
}
template <bool COH>
__device__ __forceinline__ double ldp1(const double* p)
{
    if constexpr (COH) return __ldcg(p);
    else return __ldg(p);
}
// Solver scalars written by the tail CTA of the preceding kernel: read through L2 (see ldp2 above)
__device__ __forceinline__ int sld(const int* p) { return __ldcg(p); }
__device__ __forceinline__ double sld(const double* p) { return __ldcg(p); }
struct RedArgs {
    double* partials;   // [NR][cap]   one per tile
    double* gpartials;  // [NR][gcap]  one per group of TL_RED_GROUP tiles
    unsigned int* gcount; // [gcap]    arrival tickets per group (self-resetting)
    int cap, gcap;
    DevScal* S;
};
#define TL_RED_GROUP 64

// Deterministic single-pass grid reduction, two levels.  Every CTA writes its NR tile partials.  Tiles
// are grouped by index (64 per group): the last CTA of a group to arrive adds that group's partials in
// a fixed order -- this happens while the rest of the grid is still streaming -- and the last group to
// finish adds the group sums, again in a fixed order.  The serial tail after the last tile is therefore
// ~ntiles/64 values instead of ntiles.  The result does not depend on arrival order.
// Returns true in EVERY thread of the final CTA with the totals in `tot`.
// CONSUMERS_ONLY: the CTA also holds a producer warp (tl_bulk.cu): only the first TL_TPB threads take part and
// synchronise through named barrier 1 instead of __syncthreads().
template <bool CONSUMERS_ONLY>
__device__ __forceinline__ void red_sync()
{
    if constexpr (CONSUMERS_ONLY) asm volatile("bar.sync 1, %0;" ::"n"(TL_TPB) : "memory");
    else __syncthreads();
}
// `hook(group)` (optional) runs in every thread of the last CTA of each group once all tiles of that group are done
// and their writes are visible -- the resident multi-rank kernels forward the group's parked halo columns there, while
// the rest of the grid is still streaming.
struct NoGroupHook {
    __device__ __forceinline__ void operator()(int) const {}
};
template <int NR, bool CONSUMERS_ONLY = false, class Hook = NoGroupHook>
__device__ __forceinline__ bool grid_reduce(double (&acc)[NR], const RedArgs& ra, int tile, int ntiles,
                                            double (&tot)[NR], Hook hook = Hook())
{
    __shared__ double sm[NR][TL_TPB / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int group = tile / TL_RED_GROUP;
    const int ngroups = (ntiles + TL_RED_GROUP - 1) / TL_RED_GROUP;
    const int gsize = min(TL_RED_GROUP, ntiles - group * TL_RED_GROUP);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double v = warp_sum(acc[r]);
        if (lane == 0) sm[r][wid] = v;
    }
    red_sync<CONSUMERS_ONLY>();
    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double v = (sm[r][0] + sm[r][1]) + (sm[r][2] + sm[r][3]);
            __stcg(&ra.partials[(size_t)r * ra.cap + tile], v);
        }
        __threadfence();
        s_flag = (atomicAdd(&ra.gcount[group], 1u) == (unsigned int)(gsize - 1));
    }
    red_sync<CONSUMERS_ONLY>();
    if (!s_flag) return false;
    // last CTA of this group: group sum (lanes of warps 0,1 hold one tile partial each)
    __threadfence();
    hook(group);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double v = 0.0;
        if (tid < gsize) v = __ldcg(ra.partials + (size_t)r * ra.cap + group * TL_RED_GROUP + tid);
        v = warp_sum(v);
        red_sync<CONSUMERS_ONLY>();
        if (lane == 0) sm[r][wid] = v;
    }
    red_sync<CONSUMERS_ONLY>();
    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) __stcg(&ra.gpartials[(size_t)r * ra.gcap + group], sm[r][0] + sm[r][1]);
        ra.gcount[group] = 0u;
        __threadfence();
        s_flag = (atomicAdd(&ra.S->counter[0], 1u) == (unsigned int)(ngroups - 1));
    }
    red_sync<CONSUMERS_ONLY>();
    if (!s_flag) return false;
    // last group: total
    __threadfence();
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const double* src = ra.gpartials + (size_t)r * ra.gcap;
        double s = 0.0;
#pragma unroll 4
        for (int k = tid; k < ngroups; k += TL_TPB) s += __ldcg(src + k);
        s = warp_sum(s);
        red_sync<CONSUMERS_ONLY>();
        if (lane == 0) sm[r][wid] = s;
    }
    red_sync<CONSUMERS_ONLY>();
#pragma unroll
    for (int r = 0; r < NR; ++r) tot[r] = (sm[r][0] + sm[r][1]) + (sm[r][2] + sm[r][3]);
    if (tid == 0) ra.S->counter[0] = 0u;
    return true;
}

// shared.h:59-63 -- the reference SMVP association, spelled out on registers:
//   (1 + (kx[i+1]+kx[i]) + (ky[i+x]+ky[i]))*a[i] - (kx[i+1]*a[i+1]+kx[i]*a[i-1]) - (ky[i+x]*a[i+x]+ky[i]*a[i-x])
__device__ __forceinline__ double smvp(double kx0, double kx1, double ky0, double ky1, double a,
                                       double al, double ar, double ad, double au)
{
    return (1.0 + (kx1 + kx0) + (ky1 + ky0)) * a - (kx1 * ar + kx0 * al) - (ky1 * au + ky0 * ad);
}


// ---------------------------------------------------------------------------------------------
// Multi-rank helpers of the resident loops (see MultiCtx in tl_internal.h)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Optional time stamps of the resident loop (tl_stamps_enable): one thread, a few stores per kernel.
__device__ __forceinline__ void stamp(DevScal* S, int iter, int kernel, int point)
{
    unsigned long long* st = S->stamps;
    if (st && iter >= 0 && iter < S->stamp_cap)
        st[((size_t)iter * TL_STAMP_KERNELS + kernel) * TL_STAMP_POINTS + point] = global_timer_ns();
}

// "LL" cell: a double travels as two 8-byte words {low half | seq << 32, high half | seq << 32} written by ONE 16-byte
// store.  8-byte words are single-copy atomic, so a reader that finds the expected sequence number in BOTH words holds a
// complete value: no fence, no separate flag, one NVLink hop.
__device__ __forceinline__ void ll_store(unsigned long long* cell, double v, unsigned int seq)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long w0 = ((unsigned long long)seq << 32) | (b & 0xffffffffull);
    const unsigned long long w1 = ((unsigned long long)seq << 32) | (b >> 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const unsigned long long* cell, unsigned int seq, double* v,
                                            unsigned long long* seen)
{
    unsigned long long w0, w1;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(cell) : "memory");
    *seen = w0;
    if ((unsigned int)(w0 >> 32) != seq || (unsigned int)(w1 >> 32) != seq) return false;
    *v = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
    return true;
}
// Spin until the cell carries `seq`; bounded like spin_flag (a lost peer is fatal for the solve, not for the GPU).
__device__ __forceinline__ double ll_wait(const unsigned long long* cell, unsigned int seq, DevScal* S,
                                          unsigned long long site)
{
    const long long t0 = clock64();
    double v = 0.0;
    unsigned long long seen = 0ull;
    while (!ll_try_load(cell, seq, &v, &seen)) {
        if (*(volatile unsigned int*)&S->pad == 0xdeadu) break;
        if (clock64() - t0 > 20000000000LL) { // ~10 s
            if (atomicCAS(&S->pad, 0u, 0xdeadu) == 0u) {
                S->dbg[0] = site;
                S->dbg[1] = seq;
                S->dbg[2] = seen;
                S->dbg[3] = (unsigned long long)(blockIdx.y * gridDim.x + blockIdx.x);
                S->conv = 1;
                if (S->err_host) *(volatile unsigned int*)S->err_host = 0xdeadu;
                __threadfence_system();
            }
            break;
        }
        __nanosleep(32);
    }
    return v;
}
// sum_over_ranks (comms.c:55-61; cg_driver.c:85,104) inside the tail CTA of a reduction kernel.  Called by ALL 32 lanes
// of warp 0: lane r stores this rank's partial into rank r's cell (its own included), then waits for rank r's partial
// in the local copy (the N latencies overlap); the partials are added in rank order, identically on every rank.
__device__ __forceinline__ double mc_allsum_warp(const MultiCtx& mc, int kind, double partial, DevScal* S)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long n = mc.sbase + (unsigned long long)mc.tl + 1ull;
    const unsigned int seq = (unsigned int)n;
    const int par = (int)(n & 1ull);
    double v = 0.0;
    if (lane < mc.num_ranks) {
        ll_store(mc.slots_peer[lane] + 2 * TL_SLOT_IDX(kind, par, mc.rank), partial, seq);
        v = ll_wait(mc.slots_local + 2 * TL_SLOT_IDX(kind, par, lane), seq, S,
                    1000ull + 100ull * kind + 10ull * lane + (unsigned long long)mc.tl * 100000ull);
    }
    double s = __shfl_sync(0xffffffffu, v, 0);
    for (int r = 1; r < mc.num_ranks; ++r) s = s + __shfl_sync(0xffffffffu, v, r);
    return s;
}
// Tail CTA, after every CTA of the grid has fenced its remote halo stores and taken its ticket: threads 0..3 of the
// calling warp release the neighbours' per-face flags (nbf[f] != null: that face has a neighbour) and then acquire
// their own face's flag -- the neighbour's stores into MY halo are then visible to the next kernel.
__device__ __forceinline__ void mc_halo_handshake(const MultiCtx& mc, double* const* nbf, DevScal* S, int lane)
{
    if (lane < 4 && nbf[lane]) {
        const unsigned long long n = mc.hbase + (unsigned long long)mc.tl + 1ull;
        st_release_sys(mc.nb_hflag[lane], n);
        spin_flag(mc.hflags_local + lane, n, S, 20ull + (unsigned long long)lane + (unsigned long long)mc.tl * 100000ull);
    }
}
// "Last CTA of the grid" ticket for kernels without a reduction (calc_p, the one-pass Chebyshev / PPCG kernels), two
// levels like grid_reduce so that thousands of CTAs do not queue on one address: tiles take a ticket of their group of
// TL_RED_GROUP, the last tile of a group takes a ticket of the grid.  Every CTA calls it after its last store (remote
// halo stores already fenced system-wide); returns true in every thread of the last CTA to arrive.
template <class Hook>
__device__ __forceinline__ bool grid_last_cta(unsigned int* gcount, unsigned int* counter, int tile, int ntiles, Hook hook)
{
    __shared__ int s_last_cta;
    const int group = tile / TL_RED_GROUP;
    const int ngroups = (ntiles + TL_RED_GROUP - 1) / TL_RED_GROUP;
    const int gsize = min(TL_RED_GROUP, ntiles - group * TL_RED_GROUP);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last_cta = (atomicAdd(&gcount[group], 1u) == (unsigned int)(gsize - 1));
    }
    __syncthreads();
    if (!s_last_cta) return false;
    // last CTA of its group: every tile of the group is done
    __threadfence();
    hook(group);
    __syncthreads();
    if (threadIdx.x == 0) {
        gcount[group] = 0u;
        __threadfence();
        const int last = (atomicAdd(counter, 1u) == (unsigned int)(ngroups - 1));
        if (last) *counter = 0u;
        s_last_cta = last;
    }
    __syncthreads();
    if (s_last_cta) __threadfence();
    return s_last_cta != 0;
}
__device__ __forceinline__ bool conv_test(const DevScal* S, double rrn)
{
    return S->conv_mode ? (fabs(rrn) < S->eps) : (sqrt(fabs(rrn)) < S->eps);
}

// Hot-kernel tile: TL_TPB threads x 2 columns, `rows` rows.  kk is the first of the thread's two
// columns; (off + kk) is even by construction, so double2 accesses are 16-byte aligned and a warp
// covers 512 contiguous, 128-byte-aligned bytes of a row.
struct HotTile {
    int kk, j0, j1, tile, ntiles;
    bool v0, v1;
    long i;
};
// edges_first (kernels that store halo cells into the neighbours): the top and the bottom row of tiles are scheduled
// first, the others follow in the usual (forward or reversed) order, so the system-scope fence behind the remote stores
// never sits in the kernel's tail.  Only the launch order changes: the tile index (and with it the summation order of
// the grid reduction) is the logical one.
__device__ __forceinline__ HotTile hot_tile(const Geo& g, int rows, int rev, bool edges_first = false)
{
    HotTile t;
    int by = rev ? (gridDim.y - 1 - blockIdx.y) : blockIdx.y;
    if (edges_first && gridDim.y > 2) {
        const int q = blockIdx.y;
        if (q == 0) by = gridDim.y - 1;
        else if (q == 1) by = 0;
        else by = rev ? (gridDim.y - q) : (q - 1); // q = 2 .. gridDim.y-1  ->  gridDim.y-2 .. 1  or  1 .. gridDim.y-2
    }
    const int bx = rev ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    t.kk = g.hd + 2 * (bx * TL_TPB + threadIdx.x);
    t.v0 = t.kk < g.x - g.hd;
    t.v1 = t.kk + 1 < g.x - g.hd;
    t.j0 = g.hd + by * rows;
    t.j1 = min(t.j0 + rows, g.y - g.hd);
    t.i = (long)g.off + (long)t.j0 * g.pitch + t.kk;
    t.tile = by * gridDim.x + bx;
    t.ntiles = gridDim.x * gridDim.y;
    return t;
}

// Multi-rank: the thread that owns an edge cell of an INTERNAL face stores the updated value straight into the
// neighbour's halo cell over NVLink (what pack -> MPI -> unpack does in remote_halo_driver.c for depth 1; the 5-point
// stencil never reads halo corners, so none are sent).
// STAGE_COLS: the left / right COLUMN cells are instead parked in a local column buffer (mc.colbuf) and forwarded by the
// kernel's tail CTA (forward_columns): a column is one 8-byte cell per row, i.e. one strided remote store from one
// thread of every edge tile, and -- what costs -- a system-scope fence in each of those tiles (one tile in 16 at
// 4000 x 4000).  Measured on 2 x 2 ranks: calc_ur 131 us with direct column stores, against 114 us when only a row travels.
template <bool STAGE_COLS = false>
__device__ __forceinline__ void edge_remote_store(const Geo& g, const MultiCtx& mc, double* const* nbf, int jj,
                                                  double2 pv, const HotTile& t)
{
    const int last = g.x - g.hd - 1;
    if (nbf[TL_FACE_LEFT] && t.kk == g.hd) { // my first column -> left neighbour's right halo column
        if constexpr (STAGE_COLS) mc.colbuf[jj] = pv.x;
        else nbf[TL_FACE_LEFT][(long)mc.nb_off[TL_FACE_LEFT] + (long)jj * mc.nb_pitch[TL_FACE_LEFT] +
                               (mc.nb_x[TL_FACE_LEFT] - g.hd)] = pv.x;
    }
    if (nbf[TL_FACE_RIGHT] && (t.kk == last || t.kk + 1 == last)) { // my last column -> right neighbour's left halo column
        const double v = (t.kk == last) ? pv.x : pv.y;
        if constexpr (STAGE_COLS) mc.colbuf[mc.col_cap + jj] = v;
        else nbf[TL_FACE_RIGHT][(long)mc.nb_off[TL_FACE_RIGHT] + (long)jj * mc.nb_pitch[TL_FACE_RIGHT] + (g.hd - 1)] = v;
    }
    if (nbf[TL_FACE_BOTTOM] && jj == g.hd) { // my first row -> bottom neighbour's top halo row
        double* q = nbf[TL_FACE_BOTTOM] + (long)mc.nb_off[TL_FACE_BOTTOM] +
                    (long)(mc.nb_y[TL_FACE_BOTTOM] - g.hd) * mc.nb_pitch[TL_FACE_BOTTOM] + t.kk;
        st_pair(q, pv, t.v1);
    }
    if (nbf[TL_FACE_TOP] && jj == g.y - g.hd - 1) { // my last row -> top neighbour's bottom halo row
        double* q = nbf[TL_FACE_TOP] + (long)mc.nb_off[TL_FACE_TOP] + (long)(g.hd - 1) * mc.nb_pitch[TL_FACE_TOP] +
                    t.kk;
        st_pair(q, pv, t.v1);
    }
}

// Forwarding of the parked halo columns, group by group: run by every thread of the last CTA of tile group `group`
// (TL_RED_GROUP consecutive tiles) once all tiles of the group are done.  The rows covered by the group's left-edge
// (right-edge) tiles are copied from the local column buffer into the left (right) neighbour's halo column, and every
// forwarding thread fences system-wide: the group's ticket of the grid-level count is taken after that, so when the
// kernel's tail CTA releases the per-face flags every column cell has arrived.  Overlaps with the streaming of the rest
// of the grid; the tail itself only hand-shakes.
struct ForwardColumns {
    const Geo& g;
    const MultiCtx& mc;
    int rows, gx, ntiles;
    bool on;
    __device__ __forceinline__ void operator()(int group) const
    {
        if (!on) return;
        bool any = false;
        for (int f = TL_FACE_LEFT; f <= TL_FACE_RIGHT; ++f) {
            if (!mc.nb_f[f]) continue;
            const int bxf = (f == TL_FACE_LEFT) ? 0 : gx - 1;
            const int t_lo = group * TL_RED_GROUP, t_hi = min(t_lo + TL_RED_GROUP, ntiles); // tiles [t_lo, t_hi)
            // tile-rows by with t_lo <= by * gx + bxf < t_hi
            const int by_lo = (t_lo - bxf + gx - 1) / gx > 0 ? (t_lo - bxf + gx - 1) / gx : 0;
            if (t_hi - 1 - bxf < 0) continue;
            const int by_hi = (t_hi - 1 - bxf) / gx;
            const int j_lo = g.hd + by_lo * rows, j_hi = min(g.hd + (by_hi + 1) * rows, g.y - g.hd);
            const double* src = mc.colbuf + (f == TL_FACE_LEFT ? 0 : mc.col_cap);
            double* dst = mc.nb_f[f] + (long)mc.nb_off[f] + (f == TL_FACE_LEFT ? (mc.nb_x[f] - g.hd) : (g.hd - 1));
            const long pitch = mc.nb_pitch[f];
            for (int jj = j_lo + (int)threadIdx.x; jj < j_hi; jj += TL_TPB) {
                dst[(long)jj * pitch] = __ldcg(src + jj);
                any = true;
            }
        }
        if (any) __threadfence_system();
    }
};

// Did this tile make remote halo stores, i.e. does it own cells on a face that has a neighbour?  (Those tiles fence
// their stores system-wide before they take their end-of-kernel ticket.)  staged_cols: the columns went to the local
// column buffer, only the top / bottom rows were stored remotely.
__device__ __forceinline__ bool tile_sends_halo(const Geo& g, const MultiCtx& mc, const HotTile& t, bool staged_cols = false)
{
    const bool rows = (t.j0 == g.hd && mc.nb_f[TL_FACE_BOTTOM]) || (t.j1 == g.y - g.hd && mc.nb_f[TL_FACE_TOP]);
    if (staged_cols) return rows;
    return rows || (t.kk - 2 * (int)threadIdx.x == g.hd && mc.nb_f[TL_FACE_LEFT]) ||
           (t.kk - 2 * (int)threadIdx.x + TL_TILE_COLS >= g.x - g.hd && mc.nb_f[TL_FACE_RIGHT]);
}
