// tl_comms.cu -- the comms layer: replaces the reference's MPI layer (TeaLeaf/comms.c:7-82) for the
// ranks of one NVSwitch-connected node, one process per GPU.
//
//   rendezvous / host scalars : a POSIX shared-memory segment named by the session string
//   halo payloads             : pack kernels store straight into the neighbour's receive buffer
//                               through CUDA-IPC peer mappings (NVLink stores), followed by a
//                               release flag; the receiver's stream spins on the flag (acquire)
//                               and unpacks.  No host staging, no MPI, no NCCL call on this path.
//   host-buffer messages      : send_recv_message semantics of comms.c:31-52 through mailboxes in
//                               the shared segment (the plugin-API path and the CPU tests).
//
// sum_over_ranks adds the per-rank values in rank order on every rank, so the result is
// bit-identical on all ranks and from run to run (MPI_Allreduce's order is unspecified).
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <atomic>
#include "tl_internal.h"

#define TLC_MAX_RANKS TL_MAX_PEERS
#define TLC_SLOT_DOUBLES (6 * 2 * 32776)
#define TLC_MAGIC 0x7ea1eaf5u

struct Mailbox {
    std::atomic<unsigned long long> seq_w, seq_r;
    int len;
    int pad;
    double data[TLC_SLOT_DOUBLES];
};

struct ShmHeader {
    std::atomic<unsigned int> magic;
    std::atomic<int> attached;
    int num_ranks;
    std::atomic<unsigned int> bar_count;
    std::atomic<unsigned int> bar_gen;
    std::atomic<int> aborted;                 // tl_comms_abort(): every blocked or later host wait fails at once
    double red[2][TLC_MAX_RANKS];
    std::atomic<int> nb_list[TLC_MAX_RANKS][4]; // neighbour ranks each rank sends to (registered lazily)
    // GPU arenas
    cudaIpcMemHandle_t arena[TLC_MAX_RANKS];
    cudaIpcMemHandle_t pfield[TLC_MAX_RANKS]; // each rank's field slab (resident CG loop stores halos into p)
    unsigned long long arena_off[TLC_MAX_RANKS];  // byte offset of the arena / p field inside the shared block:
    unsigned long long pfield_off[TLC_MAX_RANKS]; // cudaIpcOpenMemHandle maps the BASE of the allocation block
    unsigned long long arena_face_elems[TLC_MAX_RANKS];
    int arena_device[TLC_MAX_RANKS];
    int geo[TLC_MAX_RANKS][4];                // x, y, pitch, off of each rank's chunk
    unsigned long long field_elems[TLC_MAX_RANKS]; // doubles per field inside each rank's slab
};

struct tl_comms {
    int rank, num_ranks, device, host_only;
    char name[128];
    size_t shm_bytes;
    ShmHeader* hdr;
    Mailbox* boxes; // [num_ranks][4][2]
    unsigned int red_parity;
    unsigned int bar_local_gen;
    // GPU side (after attach)
    double* arena;      // my device arena
    size_t arena_bytes;
    size_t face_elems;  // common (max over ranks) per-face capacity in doubles
    void* peer_arena[TLC_MAX_RANKS];
    void* peer_pfield[TLC_MAX_RANKS];
    void* peer_arena_base[TLC_MAX_RANKS];  // what cudaIpcOpenMemHandle returned (to close)
    void* peer_pfield_base[TLC_MAX_RANKS];
    unsigned long long face_seq[4]; // halo messages exchanged over each of my faces (same count on both ends)
    unsigned long long res_seq;  // resident-loop iterations launched so far (slot / halo flag base)
};

static double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static const double TLC_TIMEOUT_S = 60.0;

#define SPIN_UNTIL(cond, what)                                              \
    do {                                                                    \
        double t0_ = now_s();                                               \
        unsigned long spins_ = 0;                                           \
        while (!(cond)) {                                                   \
            if ((++spins_ & 0x3ff) == 0) {                                  \
                if (k->hdr && k->hdr->aborted.load(std::memory_order_acquire)) { \
                    tl_set_error("comms: a peer rank aborted (while waiting for %s)", what); \
                    return TL_ERR_COMMS;                                    \
                }                                                           \
                sched_yield();                                              \
                if (now_s() - t0_ > TLC_TIMEOUT_S) {                        \
                    tl_set_error("comms timeout waiting for %s", what);     \
                    return TL_ERR_COMMS;                                    \
                }                                                           \
            }                                                               \
        }                                                                   \
    } while (0)

extern "C" int tl_comms_create(tl_comms** out, const char* session, int rank, int num_ranks, int device,
                               int host_only)
{
    TL_CHECK_ARG(out && session && num_ranks >= 1 && num_ranks <= TLC_MAX_RANKS && rank >= 0 && rank < num_ranks,
                 "bad arguments");
    tl_comms* k = new tl_comms();
    memset(k, 0, sizeof(*k));
    k->rank = rank;
    k->num_ranks = num_ranks;
    k->device = device;
    k->host_only = host_only;
    snprintf(k->name, sizeof(k->name), "/tl_b200_%s", session);
    k->shm_bytes = sizeof(ShmHeader) + sizeof(Mailbox) * (size_t)num_ranks * 4 * 2;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(k->name);
        fd = shm_open(k->name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)k->shm_bytes) != 0) {
            tl_set_error("shm_open/ftruncate(%s) failed: %s", k->name, strerror(errno));
            delete k;
            return TL_ERR_COMMS;
        }
    } else {
        double t0 = now_s();
        for (;;) {
            fd = shm_open(k->name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size >= k->shm_bytes) break;
                close(fd);
                fd = -1;
            }
            if (now_s() - t0 > TLC_TIMEOUT_S) {
                tl_set_error("timeout opening shm %s", k->name);
                delete k;
                return TL_ERR_COMMS;
            }
            usleep(1000);
        }
    }
    void* m = mmap(nullptr, k->shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
        tl_set_error("mmap(%s) failed: %s", k->name, strerror(errno));
        delete k;
        return TL_ERR_COMMS;
    }
    k->hdr = (ShmHeader*)m;
    k->boxes = (Mailbox*)((char*)m + sizeof(ShmHeader));
    if (rank == 0) {
        // ftruncate zero-fills; publish
        k->hdr->num_ranks = num_ranks;
        for (int r = 0; r < TLC_MAX_RANKS; ++r)
            for (int i = 0; i < 4; ++i) k->hdr->nb_list[r][i].store(-1);
        k->hdr->magic.store(TLC_MAGIC, std::memory_order_release);
    } else {
        SPIN_UNTIL(k->hdr->magic.load(std::memory_order_acquire) == TLC_MAGIC, "shm header");
    }
    k->hdr->attached.fetch_add(1);
    SPIN_UNTIL(k->hdr->attached.load() >= num_ranks, "all ranks to attach");
    *out = k;
    return tl_comms_barrier(k);
}

extern "C" int tl_comms_destroy(tl_comms* k)
{
    if (!k) return TL_OK;
    if (!k->host_only) {
        for (int r = 0; r < k->num_ranks; ++r)
            if (r != k->rank) {
                if (k->peer_arena_base[r]) cudaIpcCloseMemHandle(k->peer_arena_base[r]);
                if (k->peer_pfield_base[r]) cudaIpcCloseMemHandle(k->peer_pfield_base[r]);
            }
    }
    // Unlink the name BEFORE the last barrier: once any rank has left this function the name no longer
    // resolves to this (still mapped) segment, so an immediate re-create under the same name cannot attach
    // a rank to the dying segment.
    if (k->rank == 0) shm_unlink(k->name);
    tl_comms_barrier(k);
    if (k->arena) cudaFree(k->arena);
    munmap(k->hdr, k->shm_bytes);
    delete k;
    return TL_OK;
}

// abort_comms(), comms.c:79-82 (MPI_Abort): poisons the shared header so that the other ranks' host-side waits
// (barrier, reductions, mailboxes) fail with TL_ERR_COMMS at once instead of running into their time-outs.
extern "C" int tl_comms_abort(tl_comms* k)
{
    if (k && k->hdr) k->hdr->aborted.store(1, std::memory_order_release);
    return TL_OK;
}

extern "C" int tl_comms_rank(const tl_comms* k) { return k ? k->rank : 0; }
extern "C" int tl_comms_size(const tl_comms* k) { return k ? k->num_ranks : 1; }

// barrier(), comms.c:73-76
extern "C" int tl_comms_barrier(tl_comms* k)
{
    if (!k || k->num_ranks == 1) return TL_OK;
    ShmHeader* h = k->hdr;
    const unsigned int gen = h->bar_gen.load(std::memory_order_acquire);
    if (h->bar_count.fetch_add(1, std::memory_order_acq_rel) == (unsigned int)(k->num_ranks - 1)) {
        h->bar_count.store(0, std::memory_order_relaxed);
        h->bar_gen.store(gen + 1, std::memory_order_release);
    } else {
        SPIN_UNTIL(h->bar_gen.load(std::memory_order_acquire) != gen, "barrier");
    }
    return TL_OK;
}

static int reduce_common(tl_comms* k, double* a, bool is_min)
{
    if (!k || k->num_ranks == 1) return TL_OK;
    const unsigned int par = (k->red_parity++) & 1u;
    k->hdr->red[par][k->rank] = *a;
    std::atomic_thread_fence(std::memory_order_release);
    TL_TRY(tl_comms_barrier(k));
    std::atomic_thread_fence(std::memory_order_acquire);
    double v = k->hdr->red[par][0];
    for (int r = 1; r < k->num_ranks; ++r) {
        const double b = k->hdr->red[par][r];
        if (is_min) v = (b < v) ? b : v;
        else v += b; // rank order
    }
    *a = v;
    return TL_OK;
}
// sum_over_ranks, comms.c:55-61
extern "C" int tl_comms_sum(tl_comms* k, double* a) { return reduce_common(k, a, false); }
// min_over_ranks, comms.c:64-70
extern "C" int tl_comms_min(tl_comms* k, double* a) { return reduce_common(k, a, true); }

static int nb_index(tl_comms* k, int src, int dst, bool create)
{
    for (int i = 0; i < 4; ++i) {
        int v = k->hdr->nb_list[src][i].load(std::memory_order_acquire);
        if (v == dst) return i;
        if (v == -1) {
            if (!create) return -1;
            int expect = -1;
            if (k->hdr->nb_list[src][i].compare_exchange_strong(expect, dst)) return i;
            if (expect == dst) return i;
        }
    }
    return -1;
}

// send_recv_message + wait_for_requests, comms.c:31-52, on host buffers.  tl_comms_post is the
// MPI_Isend half (copies the message into this rank's mailbox for `neighbour`), tl_comms_recv the
// MPI_Irecv + MPI_Wait half.  Posting every message of a phase before receiving any keeps the
// exchange free of rank-to-rank chains.
extern "C" int tl_comms_post(tl_comms* k, const double* send_buffer, int buffer_len, int neighbour, int send_tag)
{
    TL_CHECK_ARG(k && send_buffer && buffer_len >= 0 && buffer_len <= TLC_SLOT_DOUBLES && neighbour >= 0 &&
                     neighbour < k->num_ranks && neighbour != k->rank && (send_tag == 0 || send_tag == 1),
                 "bad arguments");
    const int si = nb_index(k, k->rank, neighbour, true);
    TL_CHECK_ARG(si >= 0, "more than 4 neighbours");
    Mailbox* sb = &k->boxes[((size_t)k->rank * 4 + si) * 2 + send_tag];
    SPIN_UNTIL(sb->seq_w.load(std::memory_order_acquire) == sb->seq_r.load(std::memory_order_acquire),
               "mailbox to drain");
    memcpy(sb->data, send_buffer, sizeof(double) * (size_t)buffer_len);
    sb->len = buffer_len;
    sb->seq_w.fetch_add(1, std::memory_order_release);
    return TL_OK;
}

extern "C" int tl_comms_recv(tl_comms* k, double* recv_buffer, int buffer_len, int neighbour, int recv_tag)
{
    TL_CHECK_ARG(k && recv_buffer && buffer_len >= 0 && neighbour >= 0 && neighbour < k->num_ranks &&
                     neighbour != k->rank && (recv_tag == 0 || recv_tag == 1), "bad arguments");
    int ri = -1;
    SPIN_UNTIL((ri = nb_index(k, neighbour, k->rank, false)) >= 0, "neighbour to post");
    Mailbox* rb = &k->boxes[((size_t)neighbour * 4 + ri) * 2 + recv_tag];
    SPIN_UNTIL(rb->seq_w.load(std::memory_order_acquire) > rb->seq_r.load(std::memory_order_acquire),
               "message to arrive");
    if (rb->len != buffer_len) {
        tl_set_error("message length mismatch: expected %d got %d", buffer_len, rb->len);
        return TL_ERR_COMMS;
    }
    memcpy(recv_buffer, rb->data, sizeof(double) * (size_t)buffer_len);
    rb->seq_r.fetch_add(1, std::memory_order_release);
    return TL_OK;
}

extern "C" int tl_comms_send_recv(tl_comms* k, const double* send_buffer, double* recv_buffer, int buffer_len,
                                  int neighbour, int send_tag, int recv_tag)
{
    TL_TRY(tl_comms_post(k, send_buffer, buffer_len, neighbour, send_tag));
    return tl_comms_recv(k, recv_buffer, buffer_len, neighbour, recv_tag);
}

// initialise.c:34-134, for chunk == rank (one chunk per rank)
extern "C" int tl_decompose(int grid_x_cells, int grid_y_cells, int num_chunks, int chunk, int* nx, int* ny,
                            int* left, int* bottom, int neighbours[4], int* x_chunks_out, int* y_chunks_out)
{
    TL_CHECK_ARG(grid_x_cells > 0 && grid_y_cells > 0 && num_chunks >= 1 && chunk >= 0 && chunk < num_chunks,
                 "bad arguments");
    double best = 1.7976931348623157e308; // DBL_MAX
    const double xc = (double)grid_x_cells, yc = (double)grid_y_cells;
    int x_chunks = 0, y_chunks = 0;
    for (int xx = 1; xx <= num_chunks; ++xx) {
        if (num_chunks % xx) continue;
        const int yy = num_chunks / xx;
        if (num_chunks % yy) continue;
        const double perimeter = ((xc / xx) * (xc / xx) + (yc / yy) * (yc / yy)) * 2;
        const double area = (xc / xx) * (yc / yy);
        const double metric = perimeter / area;
        if (metric < best) {
            x_chunks = xx;
            y_chunks = yy;
            best = metric;
        }
    }
    if (!x_chunks || !y_chunks) {
        tl_set_error("Failed to decompose the field with given parameters.");
        return TL_ERR_ARG;
    }
    const int dx = grid_x_cells / x_chunks, dy = grid_y_cells / y_chunks;
    const int mod_x = grid_x_cells % x_chunks, mod_y = grid_y_cells % y_chunks;
    const int xx = chunk % x_chunks, yy = chunk / x_chunks;
    const int add_x = (xx < mod_x), add_y = (yy < mod_y);
    // remainder cells go to the first `mod` chunks of each axis (initialise.c:80-133)
    const int add_x_prev = (xx < mod_x) ? xx : mod_x, add_y_prev = (yy < mod_y) ? yy : mod_y;
    if (nx) *nx = dx + add_x;
    if (ny) *ny = dy + add_y;
    if (left) *left = xx * dx + add_x_prev;
    if (bottom) *bottom = yy * dy + add_y_prev;
    if (neighbours) {
        neighbours[TL_FACE_LEFT] = (xx == 0) ? TL_EXTERNAL_FACE : chunk - 1;
        neighbours[TL_FACE_RIGHT] = (xx == x_chunks - 1) ? TL_EXTERNAL_FACE : chunk + 1;
        neighbours[TL_FACE_BOTTOM] = (yy == 0) ? TL_EXTERNAL_FACE : chunk - x_chunks;
        neighbours[TL_FACE_TOP] = (yy == y_chunks - 1) ? TL_EXTERNAL_FACE : chunk + x_chunks;
    }
    if (x_chunks_out) *x_chunks_out = x_chunks;
    if (y_chunks_out) *y_chunks_out = y_chunks;
    return TL_OK;
}

// ---------------------------------------------------------------------------------------------
// GPU peer path
// ---------------------------------------------------------------------------------------------
// cudaMalloc may sub-allocate from a larger block and CUDA IPC shares the whole block: the opener gets the
// block base.  Offsets are taken with cuMemGetAddressRange (fetched at run time: no link-time libcuda
// dependency, so the library still loads on CPU-only machines).
static int block_offset(const void* ptr, unsigned long long* off)
{
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    static range_fn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !sym) {
            tl_set_error("cudaGetDriverEntryPoint(cuMemGetAddressRange) failed");
            return TL_ERR_CUDA;
        }
        fn = (range_fn)sym;
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (fn(&base, &size, (unsigned long long)ptr) != 0) {
        tl_set_error("cuMemGetAddressRange failed");
        return TL_ERR_CUDA;
    }
    *off = (unsigned long long)ptr - base;
    return TL_OK;
}

// Arena layout (identical on every rank, sized with the max face capacity over ranks), in doubles:
//   recv[face 0..3][parity 0..1][face_elems] | tail:
//   tail+0..3   flags of the generic halo exchange (by receiving face)
//   tail+4..7   p-halo flags of the resident CG loop (by receiving face)
//   tail+8..71  reduction cells  [kind][parity][rank] x 2 words {value half | sequence << 32} (MultiCtx, tl_internal.h)
//   tail+72..75 ack flags of the generic halo exchange (by SENDING face: "message n of this face was unpacked")
static size_t arena_recv_off(const tl_comms* k, int face, int parity)
{
    return ((size_t)face * 2 + parity) * k->face_elems;
}
static size_t arena_flags_off(const tl_comms* k) { return (size_t)8 * k->face_elems; }
static size_t arena_total(const tl_comms* k) { return arena_flags_off(k) + 256; }
#define ARENA_HFLAGS 4
#define ARENA_SLOTS 8
#define ARENA_ACKS 72

extern "C" int tl_comms_attach_chunk(tl_comms* k, tl_chunk* c)
{
    TL_CHECK_ARG(k && c, "null argument");
    c->comms = k;
    if (k->host_only || k->num_ranks == 1) return TL_OK;
    TL_CUDA(cudaSetDevice(c->device));
    ShmHeader* h = k->hdr;
    h->arena_face_elems[k->rank] = c->face_elems;
    h->arena_device[k->rank] = c->device;
    h->geo[k->rank][0] = c->g.x;
    h->geo[k->rank][1] = c->g.y;
    h->geo[k->rank][2] = c->g.pitch;
    h->geo[k->rank][3] = c->g.off;
    h->field_elems[k->rank] = c->field_elems;
    TL_TRY(tl_comms_barrier(k));
    size_t fe = 0;
    for (int r = 0; r < k->num_ranks; ++r) fe = h->arena_face_elems[r] > fe ? h->arena_face_elems[r] : fe;
    k->face_elems = fe;
    k->arena_bytes = (arena_total(k) * sizeof(double) + (2u << 20) - 1) / (2u << 20) * (2u << 20);
    TL_CUDA(cudaMalloc((void**)&k->arena, k->arena_bytes));
    TL_CUDA(cudaMemset(k->arena, 0, k->arena_bytes));
    TL_CUDA(cudaDeviceSynchronize());
    TL_CUDA(cudaIpcGetMemHandle(&h->arena[k->rank], k->arena));
    TL_CUDA(cudaIpcGetMemHandle(&h->pfield[k->rank], c->slab));
    {
        unsigned long long o1 = 0, o2 = 0;
        TL_TRY(block_offset(k->arena, &o1));
        TL_TRY(block_offset(c->slab, &o2));
        h->arena_off[k->rank] = o1;
        h->pfield_off[k->rank] = o2; // offset of the slab inside its allocation block
    }
    std::atomic_thread_fence(std::memory_order_release);
    TL_TRY(tl_comms_barrier(k));
    std::atomic_thread_fence(std::memory_order_acquire);
    k->peer_arena[k->rank] = k->arena;
    // the reduction slots live on every rank: map all arenas; p fields only of the neighbours
    for (int r = 0; r < k->num_ranks; ++r) {
        if (r == k->rank || k->peer_arena[r]) continue;
        cudaError_t e = cudaIpcOpenMemHandle(&k->peer_arena[r], h->arena[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            tl_set_error("cudaIpcOpenMemHandle(arena, rank %d -> %d) failed: %s", k->rank, r, cudaGetErrorString(e));
            return TL_ERR_COMMS;
        }
        k->peer_arena_base[r] = k->peer_arena[r];
        k->peer_arena[r] = (char*)k->peer_arena[r] + h->arena_off[r];
    }
    for (int f = 0; f < 4; ++f) {
        const int n = c->nb[f];
        if (n == TL_EXTERNAL_FACE || k->peer_pfield[n]) continue;
        cudaError_t e = cudaIpcOpenMemHandle(&k->peer_pfield[n], h->pfield[n], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            tl_set_error("cudaIpcOpenMemHandle(p field, rank %d -> %d) failed: %s", k->rank, n, cudaGetErrorString(e));
            return TL_ERR_COMMS;
        }
        k->peer_pfield_base[n] = k->peer_pfield[n];
        k->peer_pfield[n] = (char*)k->peer_pfield[n] + h->pfield_off[n];
    }
    static const int opposite[4] = {TL_FACE_RIGHT, TL_FACE_LEFT, TL_FACE_TOP, TL_FACE_BOTTOM};
    MultiCtx& mc = c->mc;
    memset(&mc, 0, sizeof(mc));
    mc.num_ranks = k->num_ranks;
    mc.rank = k->rank;
    mc.colbuf = c->colbuf;
    mc.col_cap = c->g.y;
    double* tail = k->arena + arena_flags_off(k);
    mc.slots_local = (unsigned long long*)(tail + ARENA_SLOTS);
    mc.hflags_local = (unsigned long long*)(tail + ARENA_HFLAGS);
    for (int r = 0; r < k->num_ranks; ++r) {
        double* ptail = (double*)k->peer_arena[r] + arena_flags_off(k);
        mc.slots_peer[r] = (unsigned long long*)(ptail + ARENA_SLOTS);
    }
    for (int f = 0; f < 4; ++f) {
        const int n = c->nb[f];
        c->nb_recv[f] = nullptr;
        c->nb_slab[f] = nullptr;
        c->nb_flag[f] = nullptr;
        c->nb_ack[f] = nullptr;
        c->my_ack[f] = (unsigned long long*)(tail + ARENA_ACKS) + f;
        if (n == TL_EXTERNAL_FACE) continue;
        double* base = (double*)k->peer_arena[n];
        c->nb_recv[f] = base + arena_recv_off(k, opposite[f], 0);
        c->nb_flag[f] = (unsigned long long*)(base + arena_flags_off(k)) + opposite[f];
        c->nb_ack[f] = (unsigned long long*)(base + arena_flags_off(k) + ARENA_ACKS) + opposite[f];
        c->nb_slab[f] = (double*)k->peer_pfield[n];
        c->nb_field_elems[f] = h->field_elems[n];
        mc.nb_hflag[f] = (unsigned long long*)(base + arena_flags_off(k) + ARENA_HFLAGS) + opposite[f];
        mc.nb_x[f] = h->geo[n][0];
        mc.nb_y[f] = h->geo[n][1];
        mc.nb_pitch[f] = h->geo[n][2];
        mc.nb_off[f] = h->geo[n][3];
    }
    c->has_peers = true;
    return tl_comms_barrier(k);
}

// Sequence bases of the resident loop: every rank launches the same iterations in the same order, so
// a per-endpoint counter stays in lockstep across ranks.
unsigned long long tlc_resident_seq_advance(tl_comms* k, int launched)
{
    const unsigned long long base = k->res_seq;
    k->res_seq += (unsigned long long)launched;
    return base;
}

// remote_halo_driver.c:11-129 over NVLink: L/R fully completes (incl. unpack) before B/T packs.
int tlc_halo_exchange(tl_chunk* c, tl_comms* k, const int fields[6], int depth)
{
    if (!k || k->num_ranks == 1) return TL_OK;
    bool any_nb = false;
    for (int f = 0; f < 4; ++f) any_nb |= (c->nb[f] != TL_EXTERNAL_FACE);
    if (!any_nb) return TL_OK;
    if (k->host_only || !c->has_peers) {
        tl_set_error("halo exchange needs an attached GPU comms endpoint");
        return TL_ERR_COMMS;
    }
    // TL_TEST_SKEW="rank:microseconds": that rank's host sleeps between the send and the unpack launches of every
    // phase (tests only: widens the window in which a neighbour can run ahead by one exchange)
    static int skew_rank = -2, skew_us = 0;
    if (skew_rank == -2) {
        skew_rank = -1;
        const char* e = getenv("TL_TEST_SKEW");
        if (e) sscanf(e, "%d:%d", &skew_rank, &skew_us);
    }
    for (int phase = 0; phase < 2; ++phase) {
        const int f0 = phase ? TL_FACE_BOTTOM : TL_FACE_LEFT;
        int faces[2];
        double* sbuf[2];
        double* rbuf[2];
        unsigned long long* sflag[2];
        unsigned long long* rflag[2];
        unsigned long long* sack[2];
        unsigned long long* rack[2];
        unsigned long long seqs[2];
        for (int q = 0; q < 2; ++q) {
            const int f = f0 + q;
            const bool has = (c->nb[f] != TL_EXTERNAL_FACE);
            // message number on this face: both ends count the same exchanges, and consecutive messages of a face
            // alternate between its two receive buffers
            seqs[q] = has ? ++k->face_seq[f] : 0ull;
            const int par = (int)(seqs[q] & 1ull);
            faces[q] = has ? f : -1;
            sbuf[q] = has ? c->nb_recv[f] + (size_t)par * k->face_elems : nullptr;
            sflag[q] = has ? c->nb_flag[f] : nullptr;
            sack[q] = has ? c->my_ack[f] : nullptr;
            rbuf[q] = has ? k->arena + arena_recv_off(k, f, par) : nullptr;
            rflag[q] = has ? (unsigned long long*)(k->arena + arena_flags_off(k)) + f : nullptr;
            rack[q] = has ? c->nb_ack[f] : nullptr;
        }
        if (faces[0] < 0 && faces[1] < 0) continue;
        // pack both faces of the phase into the neighbours' buffers + release their flags: one launch;
        // acquire my flags + unpack + hand the buffers back: one launch (remote_halo_driver.c:24-126 order: L/R done
        // before B/T packs)
        TL_TRY(tlk_phase_exchange(c, fields, depth, true, faces, sbuf, sflag, sack, seqs));
        if (skew_rank == k->rank && skew_us > 0) usleep(skew_us);
        TL_TRY(tlk_phase_exchange(c, fields, depth, false, faces, rbuf, rflag, rack, seqs));
    }
    return TL_OK;
}

// ---------------------------------------------------------------------------------------------
// Halo-exchange stress (tests/test_multigpu.py): `reps` back-to-back exchanges with NO host synchronisation in
// between.  Before exchange t a kernel fills the interior of the fields with a pattern of (global cell, t, field); after
// it a kernel compares every cell within `depth` of the interior (corners included) with what
// remote_halo_driver.c:24-126 + local_halos.cpp must have produced -- the neighbour's cell, or the reflection at the
// edge of the global mesh -- and counts mismatches on the device.  A receive buffer that is overwritten before it was
// unpacked, or unpacked twice, shows up as a mismatch of exchange t against the data of exchange t +- 1.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double stress_value(int gx, int gy, int rep, int field)
{
    return (double)gx + 10000.0 * (double)gy + 0.25 * (double)rep + 1.0e8 * (double)field;
}
__global__ void k_stress_fill(Geo g, double* f, int left, int bottom, int rep, int field)
{
    const int kk = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y;
    if (kk >= g.x) return;
    const bool interior = kk >= g.hd && kk < g.x - g.hd && jj >= g.hd && jj < g.y - g.hd;
    f[(long)g.off + (long)jj * g.pitch + kk] =
        interior ? stress_value(left + kk - g.hd, bottom + jj - g.hd, rep, field) : -777.0;
}
__global__ void k_stress_check(Geo g, const double* f, int left, int bottom, int gx_cells, int gy_cells, int depth,
                               int rep, int field, unsigned long long* mismatches)
{
    const int kk = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y;
    const int lo = g.hd - depth;
    if (kk < lo || kk >= g.x - lo || jj < lo || jj >= g.y - lo) return;
    int gx = left + kk - g.hd, gy = bottom + jj - g.hd;
    gx = gx < 0 ? -gx - 1 : (gx >= gx_cells ? 2 * gx_cells - gx - 1 : gx);
    gy = gy < 0 ? -gy - 1 : (gy >= gy_cells ? 2 * gy_cells - gy - 1 : gy);
    if (__ldcg(f + (long)g.off + (long)jj * g.pitch + kk) != stress_value(gx, gy, rep, field))
        atomicAdd(mismatches, 1ull);
}

extern "C" int tl_halo_stress(tl_chunk* c, tl_comms* k, int grid_x_cells, int grid_y_cells, int reps, int depth,
                              long* mismatches)
{
    TL_CHECK_ARG(c && k && mismatches && reps > 0 && depth >= 1 && depth <= c->g.hd, "bad arguments");
    TL_CUDA(cudaSetDevice(c->device));
    unsigned long long* d_bad = nullptr;
    TL_CUDA(cudaMalloc((void**)&d_bad, sizeof(unsigned long long)));
    TL_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
    const int fields[2] = {TL_FIELD_U, TL_FIELD_DENSITY};
    int flags[TL_NUM_EXCHANGE_FIELDS] = {0, 0, 0, 0, 0, 0};
    dim3 grid((c->g.x + 127) / 128, c->g.y);
    for (int rep = 0; rep < reps; ++rep) {
        // alternate the field set and the depth: consecutive messages of a face differ in length and content
        const int nf = (rep & 1) ? 2 : 1;
        const int d = (rep % 3 == 2) ? 1 : depth;
        for (int i = 0; i < TL_NUM_EXCHANGE_FIELDS; ++i) flags[i] = 0;
        for (int q = 0; q < nf; ++q) {
            flags[fields[q]] = 1;
            k_stress_fill<<<grid, 128, 0, c->stream>>>(c->g, c->f[fields[q]], c->left, c->bottom, rep, fields[q]);
        }
        TL_TRY(tl_halo_update(c, k, flags, d));
        for (int q = 0; q < nf; ++q)
            k_stress_check<<<grid, 128, 0, c->stream>>>(c->g, c->f[fields[q]], c->left, c->bottom, grid_x_cells,
                                                        grid_y_cells, d, rep, fields[q], d_bad);
        g_tl_launches += 2 * nf;
    }
    TL_CUDA(cudaGetLastError());
    unsigned long long bad = 0;
    TL_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_bad);
    *mismatches = (long)bad;
    return tl_check_peer_timeout(c);
}
