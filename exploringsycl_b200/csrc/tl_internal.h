// tl_internal.h -- internal types of the B200 TeaLeaf backend (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/tealeaf_b200.h"

// ---------------------------------------------------------------------------------------------
// HBM layout of a field (see DESIGN.md "Data layout"):
//   element (jj,kk) lives at  base[off + jj*pitch + kk],  jj in [0,y), kk in [0,x)
//   pitch is a multiple of 16 doubles (128 B) and `off` is chosen so that the first interior
//   column (kk == hd) of every row starts a 128-byte line.  All fields of a chunk share
//   (pitch, off), so one flat index addresses the same cell in every field.
//   The allocation has `off` doubles of lead pad and >= 32 doubles of tail pad: 16-byte vector
//   accesses that straddle a row end stay inside the allocation.
// ---------------------------------------------------------------------------------------------
struct Geo {
    int x, y, hd;   // logical dims incl. halo, halo depth
    int pitch, off; // row pitch and lead offset, in doubles
};

// Slab slots: the TL_NUM_FIELDS public fields, then the second buffers of the double-buffered in-place updates.  They
// live in the slab so that ONE CUDA-IPC mapping lets a neighbour store halo cells into whichever buffer is current.
#define TL_SLAB_P2 (TL_NUM_FIELDS + 0)
#define TL_SLAB_U2 (TL_NUM_FIELDS + 1)
#define TL_SLAB_SD2 (TL_NUM_FIELDS + 2)
#define TL_SLAB_SLOTS (TL_NUM_FIELDS + 3)

#define TL_TPB 128          // threads per CTA of the streaming kernels
#define TL_TILE_COLS 256    // 2 cells per thread
#define TL_MAX_PEERS 8

// Device-resident solver scalars (one per chunk). Written by the tail CTA of the reduction
// kernels, read by the head of the next kernel: alpha/beta never visit the host inside a solve.
struct DevScal {
    double rro, pw, rrn, alpha, beta, error, bb, eps;
    double sums[8];
    int iters;         // CG iterations completed in this solve
    int conv;          // set when sqrt(|rrn|) < eps (cg_driver.c:24) or iters reached max_iters: later launches are no-ops
    int p_pending;     // calc_ur ran: the matching calc_p must still run
    int max_iters;
    int conv_mode;     // 0: sqrt(|rrn|) < eps (cg_driver.c:24); 1: |rrn| < eps (cheby_driver.c:70)
    int sw_min_iters;  // CG pre-steps of the Chebyshev / PPCG drivers: stop once iters > sw_min_iters and rrn < sw_thresh
    double sw_thresh;  // (the switch rule of cheby_driver.c:30-32 evaluated on the device); sw_min_iters < 0: off
    unsigned int counter[8]; // "last CTA done" tickets, one per reduction kernel family
    unsigned int pad;        // 0xdead: a peer wait timed out
    unsigned long long dbg[4]; // first timed-out wait: site, wanted value, seen value, block id
    unsigned int* err_host;  // mapped pinned host word: 0xdead is stored there as well, so the host sees a
                             // timed-out wait at its next synchronisation point without fetching DevScal
    unsigned long long* stamps; // optional %globaltimer stamps of the resident loop kernels (tl_stamps_enable), else null
    int stamp_cap;              // iterations the stamp buffer holds
};
// stamps[(iteration * TL_STAMP_KERNELS + kernel) * TL_STAMP_POINTS + point]
//   kernel: 0 matvec (calc_w / calc_pw), 1 calc_ur, 2 calc_p
//   point : 0 first CTA past its dependency wait, 1 tail CTA holds this rank's partial (all tiles done),
//           2 all ranks' partials gathered (== 1 on one rank), 3 neighbours' halo flags seen / tail done
#define TL_STAMP_KERNELS 3
#define TL_STAMP_POINTS 4

// Multi-rank context of the resident CG loop (passed by value to the hot kernels; num_ranks == 1
// selects the single-GPU behaviour).  All peer pointers are CUDA-IPC mappings of the other ranks'
// HBM: stores to them travel over NVLink / NVSwitch.
//
// All cross-rank waiting happens in the TAIL CTA of the producing kernel (the CTA that completes the
// deterministic grid reduction); the heads of the kernels are exactly the single-GPU heads:
//   slots : [kind 0 = p.w, 1 = r.r][parity][source rank] 16-byte cells, one copy on EVERY rank.  The tail
//           CTA stores its rank's partial into all ranks' copies as two 8-byte words {value half, sequence}
//           (the "LL" scheme of latency-optimised all-reduces: data and flag travel in ONE store, no fence, one
//           NVLink hop), then spins on its own copy until all N cells carry this iteration's sequence number and
//           adds the N partials in rank order (bit-identical on every rank, same order as tl_comms_sum).  It then
//           writes alpha / beta / the convergence flag exactly as the single-GPU tail does.
//   halo  : calc_p (three-kernel loop: p) or calc_ur (fused loop: r) stores its edge cells straight into the
//           neighbour's halo cells; every CTA that made such stores fences them before its ticket, and the
//           tail CTA releases the neighbours' per-face flags and then acquires its own: when the kernel
//           completes, the halo this rank will read next is in place.
#define TL_SLOT_IDX(kind, par, r) (((kind) * 2 + (par)) * TL_MAX_PEERS + (r))
struct MultiCtx {
    int num_ranks, rank;
    int tl;                       // iteration index local to this resident call
    unsigned long long sbase;     // slot sequence number of local iteration tl is sbase + tl + 1
    unsigned long long hbase;     // halo flag value released after local iteration tl is hbase + tl + 1
    unsigned long long* slots_local;  // [TL_SLOT_IDX][2] words
    unsigned long long* hflags_local; // [4] by my face
    unsigned long long* slots_peer[TL_MAX_PEERS];
    double* nb_f[4];              // neighbour's copy of the field whose halo travels in this launch (peer-mapped base),
                                  // by my face; null if the face is external.  Set per launch: p (calc_p), r (fused
                                  // calc_ur), the updated u / sd buffer (one-pass Chebyshev / PPCG kernels)
    unsigned long long* nb_hflag[4]; // neighbour's halo flag of the opposite face
    int nb_pitch[4], nb_x[4], nb_y[4], nb_off[4];
    double* colbuf;               // local: [2][col_cap] parked left / right column cells (edge_remote_store<true>)
    int col_cap;
};

struct tl_chunk {
    int device;
    Geo g;
    int nx, ny;
    int max_iters;
    int nb[4];
    int left, bottom;
    size_t field_elems;           // allocated doubles per 2-D field
    double* slab;                 // one allocation holding all 2-D fields (IPC-shared with neighbours)
    size_t slab_bytes;
    double* f[TL_NUM_FIELDS];     // device, slab + f * field_elems (P may point to p2 in fused mode)
    double *cell_x, *cell_y, *vertex_x, *vertex_y; // device 1-D
    double* p2;                   // second p buffer of the fused p+w kernel (slab slot TL_SLAB_P2; P and P2 swap roles)
    double* alt[TL_NUM_FIELDS];   // second buffers of the one-pass Chebyshev (U) / PPCG (SD) kernels (slab slots
                                  // TL_SLAB_U2 / TL_SLAB_SD2; a field and its alternate swap roles every launch)
    double* colbuf;               // device: 2 * g.y doubles, the resident loops' parked halo columns
    double* nb_slab[4];           // neighbours' slabs (peer-mapped), by my face; null if external or not attached
    size_t nb_field_elems[4];     // doubles per field inside that neighbour's slab
    double* partials;             // device, per-tile partial sums (4 lanes)
    int partial_cap;              // tiles
    double* gpartials;            // device, per-group (64 tiles) sums (4 lanes)
    unsigned int* gcount;         // device, per-group arrival tickets
    int gpartial_cap;
    DevScal* scal;                // device
    DevScal* scal_h;              // pinned host mirror [0] + two polling snapshots [1],[2]
    double* d_alphas;             // device cg_alphas / cg_betas written by the resident loop
    double* d_betas;
    double *cg_alphas, *cg_betas, *cheby_alphas, *cheby_betas; // host (reference host reads/writes)
    double* face_send[4];         // device staging, 6 fields * hd * max(x,y)
    double* face_recv[4];
    size_t face_elems;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    tl_comms* comms;              // non-null once attached
    MultiCtx mc;                  // template filled by tl_comms_attach_chunk (tl / sbase / hbase / nb_f are set per launch)
    double* nb_recv[4];           // neighbours' receive buffers of the generic halo exchange (peer-mapped)
    unsigned long long* nb_flag[4];
    unsigned long long* nb_ack[4];   // neighbour's "buffer consumed" flag for what it SENDS to my face f (I release it)
    unsigned long long* my_ack[4];   // the same kind of flag for what I send through face f (the neighbour releases it)
    unsigned int* err_h;             // pinned + mapped host word behind DevScal.err_host
    bool has_peers;
    int resident_iters;           // CG iterations enqueued so far in the current solve (host bookkeeping)
    unsigned long long* stamps;   // device: optional %globaltimer stamps of the resident loop (tl_stamps_enable)
    int stamp_cap;
};

// error handling ------------------------------------------------------------------------------
void tl_set_error(const char* fmt, ...);
int tl_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define TL_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return tl_cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)
#define TL_CHECK_ARG(cond, msg)                       \
    do {                                              \
        if (!(cond)) {                                \
            tl_set_error("%s: %s", __func__, msg);    \
            return TL_ERR_ARG;                        \
        }                                             \
    } while (0)
#define TL_TRY(call)                 \
    do {                             \
        int rc_ = (call);            \
        if (rc_) return rc_;         \
    } while (0)

extern long g_tl_launches;

// kernel launchers (tl_kernels.cu) --------------------------------------------------------------
// Scalar sources for the solver kernels: either an immediate (host-driven plugin API) or the
// device-resident DevScal (resident loop).
enum ScalMode { SCAL_IMM = 0, SCAL_DEV = 1 };

int tlk_set_chunk_data(tl_chunk* c, double x_min, double y_min, double dx, double dy);
int tlk_set_initial_state(tl_chunk* c, double energy, double density);
int tlk_set_state(tl_chunk* c, const tl_state* s);
int tlk_copy_field(tl_chunk* c, int dst, int src, bool interior_only);
int tlk_field_summary(tl_chunk* c);                       // -> scal->sums[0..3]
int tlk_local_halos(tl_chunk* c, const int fields[6], int depth);
int tlk_pack_face(tl_chunk* c, const int fields[6], int depth, int face, bool pack, double* devbuf, int* len);
int tlk_phase_exchange(tl_chunk* c, const int fields[6], int depth, bool send, const int faces[2], double* const bufs[2],
                       unsigned long long* const flags[2], unsigned long long* const acks[2],
                       const unsigned long long seqs[2]);
int tlk_cg_init(tl_chunk* c, int coefficient, double rx, double ry);  // -> scal->sums[0] (rro part)
int tlk_cg_calc_w(tl_chunk* c, ScalMode mode, bool rev, const MultiCtx* mc = nullptr, bool pdl = false);                // -> scal->pw (& alpha when SCAL_DEV)
int tlk_cg_calc_ur(tl_chunk* c, ScalMode mode, double alpha, bool rev, const MultiCtx* mc = nullptr, bool send_r_halo = false, bool pdl = false); // -> scal->rrn (& beta, conv when SCAL_DEV)
int tlk_cg_calc_p(tl_chunk* c, ScalMode mode, double beta, bool rev, bool fuse_halo, const MultiCtx* mc = nullptr, bool pdl = false);
// neighbours' peer-mapped copy of the buffer `buf` of this chunk (a slab slot), by face, into mc->nb_f
void tlk_set_travelling_field(const tl_chunk* c, MultiCtx* mc, const double* buf);
int tlk_cg_calc_pw(tl_chunk* c, bool rev, const MultiCtx* mc = nullptr, bool pdl = false);                             // fused p-update + matvec (SCAL_DEV only); swaps P/P2
int tlk_cheby_init(tl_chunk* c, double theta);
int tlk_cheby_iterate(tl_chunk* c, double alpha, double beta);
int tlk_cheby_calc_u(tl_chunk* c);
// cheby_iterate + cheby_calc_u in one pass (swaps the U buffers) / ppcg_calc_ur + ppcg_calc_sd in one pass (swaps the SD
// buffers).  mc != null: the updated operand's edge cells are stored into the neighbours' halo and the kernel's last CTA
// hand-shakes with them (no separate halo exchange).  norm: also sum r.r over the interior (all ranks when mc) ->
// scal->sums[0].
int tlk_cheby_fused(tl_chunk* c, double alpha, double beta, const MultiCtx* mc = nullptr, bool norm = false);
int tlk_ppcg_fused(tl_chunk* c, double alpha, double beta, const MultiCtx* mc = nullptr, bool norm = false);
int tlk_field_home(tl_chunk* c, int field);                   // move a double-buffered field back into the slab
int tlk_ppcg_init(tl_chunk* c, double theta);
int tlk_ppcg_calc_ur(tl_chunk* c);
int tlk_ppcg_calc_sd(tl_chunk* c, double alpha, double beta);
int tlk_jacobi_init(tl_chunk* c, int coefficient, double rx, double ry);
int tlk_jacobi_iterate(tl_chunk* c);                      // -> scal->sums[0]
int tlk_calculate_residual(tl_chunk* c);
int tlk_calculate_2norm(tl_chunk* c, int field);          // -> scal->sums[0]
int tlk_finalise(tl_chunk* c);
int tlk_reset_solve_scalars(tl_chunk* c, double eps, int max_iters);
int tlk_seed_rro(tl_chunk* c);                            // scal->rro = scal->sums[0] (after cg_init)

// tile geometry / tuning of the hot kernels (tl_kernels.cu), shared with tl_bulk.cu
enum { TUNE_W = 0, TUNE_UR = 1, TUNE_P = 2, TUNE_PW = 3 };
int tlk_tile_rows(const tl_chunk* c, int kernel);
dim3 tlk_hot_grid(const tl_chunk* c, int rows);
int tlk_hot_check(const tl_chunk* c, dim3 grid);
int tlk_external_mask(const tl_chunk* c);
extern const MultiCtx g_single_ctx;
// tl_bulk.cu: the fused p-update + matvec kernel on a cp.async.bulk shared-memory row pipeline
bool tlk_pw_uses_bulk();
int tlk_cg_calc_pw_bulk(tl_chunk* c, bool rev, const MultiCtx* mc, int rows, bool pdl);

// fetch the DevScal to the pinned mirror and wait
int tl_fetch_scal(tl_chunk* c);
// TL_ERR_COMMS if a device-side wait for a peer rank has timed out since the chunk was created
int tl_check_peer_timeout(tl_chunk* c);
