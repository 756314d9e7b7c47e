/*
 * tealeaf_b200.h -- C-ABI of the B200-native (sm_100a) TeaLeaf solver backend.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.
 * Every `tl_run_*` entry point replaces one function of the reference plugin API
 * TeaLeaf/kernel_interface.h:13-71 (cited per function below); c_kernels/cuda/
 * kernel_interface.cu binds the reference's `run_*(Chunk*, Settings*, ...)`
 * symbols to them so the unmodified reference host (main.c / diffuse.c /
 * drivers/*.c) drives this backend.  `tl_comms_*` replaces TeaLeaf/comms.h:10-20.
 * `tl_solve_*` are the device-resident solver loops (the reference's sanctioned
 * whole-solve hook is DIFFUSE_OVERLOAD, TeaLeaf/main.c:32-36, application.h:9-11).
 *
 * All functions return 0 on success, non-zero on failure; tl_last_error() then
 * describes the failure.  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with TL_ERR_CUDA.
 *
 * Layout contract: host-side field images are dense row-major x*y doubles,
 * i = kk + jj*x with x = nx + 2*halo_depth (TeaLeaf/chunk.c:7-8), exactly the
 * reference layout.  In HBM the backend keeps fields pitched.
 */
#ifndef TEALEAF_B200_H
#define TEALEAF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TL_OK 0
#define TL_ERR_CUDA 1
#define TL_ERR_ARG 2
#define TL_ERR_NUMERIC 3
#define TL_ERR_COMMS 4

/* Faces and exchange-field indices: TeaLeaf/shared.h:33-45 */
#define TL_FACE_LEFT 0
#define TL_FACE_RIGHT 1
#define TL_FACE_BOTTOM 2
#define TL_FACE_TOP 3
#define TL_EXTERNAL_FACE (-1)

/* Field ids. 0..5 are the reference's FIELD_* exchange indices (shared.h:40-45). */
#define TL_FIELD_DENSITY 0
#define TL_FIELD_ENERGY0 1
#define TL_FIELD_ENERGY1 2 /* chunk->energy */
#define TL_FIELD_U 3
#define TL_FIELD_P 4
#define TL_FIELD_SD 5
#define TL_NUM_EXCHANGE_FIELDS 6
#define TL_FIELD_U0 6
#define TL_FIELD_R 7
#define TL_FIELD_W 8
#define TL_FIELD_KX 9
#define TL_FIELD_KY 10
#define TL_FIELD_VOLUME 11
#define TL_NUM_FIELDS 12
/* 1-D arrays (lengths x, y, x+1, y+1) */
#define TL_ARRAY_CELL_X 0
#define TL_ARRAY_CELL_Y 1
#define TL_ARRAY_VERTEX_X 2
#define TL_ARRAY_VERTEX_Y 3

#define TL_CONDUCTIVITY 1       /* shared.h:47 */
#define TL_RECIP_CONDUCTIVITY 2 /* shared.h:48 */

#define TL_SOLVER_JACOBI 0 /* settings.h:48-54 */
#define TL_SOLVER_CG 1
#define TL_SOLVER_CHEBY 2
#define TL_SOLVER_PPCG 3

#define TL_GEOM_RECTANGULAR 0 /* settings.h:126-131 */
#define TL_GEOM_CIRCULAR 1
#define TL_GEOM_POINT 2

typedef struct tl_chunk tl_chunk; /* opaque: one mesh chunk resident on one GPU */
typedef struct tl_comms tl_comms; /* opaque: this rank's endpoint of the comms layer */

/* State as the reference holds it AFTER parsing (settings.h:134-145): extents already
 * shrunk by dx/100 (parse_config.c:253-260). */
typedef struct {
    int geometry;
    double density, energy;
    double x_min, y_min, x_max, y_max, radius;
} tl_state;

const char* tl_last_error(void);
int tl_device_count(void);
const char* tl_version(void);

/* ---- chunk lifecycle: run_kernel_initialise / run_kernel_finalise
 *      (kernel_interface.h:17-20; kernel_initialise.cpp:31-125; chunk.c:4-30) ----
 * nx, ny: interior cells of this chunk; arrays are (nx+2hd) x (ny+2hd).
 * neighbours[4]: chunk ids per face or TL_EXTERNAL_FACE (initialise.c:114-124).
 * left/bottom: global cell offsets of the chunk (initialise.c:107-111).
 * All fields are zero-initialised. */
int tl_chunk_create(tl_chunk** out, int device, int nx, int ny, int halo_depth, int max_iters,
                    const int neighbours[4], int left, int bottom);
int tl_chunk_destroy(tl_chunk* c);
int tl_chunk_dims(const tl_chunk* c, int* x, int* y, int* halo_depth, int* pitch);
int tl_chunk_sync(tl_chunk* c); /* wait for all queued work of this chunk */

/* Dense host image <-> pitched HBM field. host must hold x*y doubles. */
int tl_field_write(tl_chunk* c, int field, const double* host);
int tl_field_read(tl_chunk* c, int field, double* host);
int tl_array_read(tl_chunk* c, int array, double* host);
/* Host coefficient arrays the reference host reads and writes directly
 * (chunk.h:73-76; cg_driver.c:93,111; cheby_driver.c:178-179): max_iters doubles each. */
double* tl_cg_alphas(tl_chunk* c);
double* tl_cg_betas(tl_chunk* c);
double* tl_cheby_alphas(tl_chunk* c);
double* tl_cheby_betas(tl_chunk* c);

/* ---- initialisation kernels ---- */
/* run_set_chunk_data, kernel_interface.h:13-14 (kernel_interface.cpp:34-52) */
int tl_run_set_chunk_data(tl_chunk* c, double grid_x_min, double grid_y_min, double dx, double dy);
/* run_set_chunk_state, kernel_interface.h:15-16 (kernel_interface.cpp:54-72) */
int tl_run_set_chunk_state(tl_chunk* c, int num_states, const tl_state* states);

/* ---- solver-wide kernels ---- */
/* run_local_halos, kernel_interface.h:23-24: reflective update of every field flagged in
 * fields_to_exchange[6] on every external face, `depth` layers. */
int tl_run_local_halos(tl_chunk* c, const int fields_to_exchange[6], int depth);
/* run_pack_or_unpack, kernel_interface.h:25-27: `host_buffer` is HOST memory
 * (chunk->left_send + field offset in the reference), y*depth or x*depth doubles. */
int tl_run_pack_or_unpack(tl_chunk* c, int depth, int face, int pack, int field, double* host_buffer);
/* Same gather/scatter, all flagged fields at once, into/out of the chunk's DEVICE
 * staging buffer for `face` (message layout of remote_halo_driver.c:132-184). Returns the
 * message length in doubles through *len. */
int tl_pack_face_device(tl_chunk* c, const int fields_to_exchange[6], int depth, int face, int pack, int* len);
int tl_face_buffer_read(tl_chunk* c, int face, int send, double* host, int len);
int tl_face_buffer_write(tl_chunk* c, int face, int send, const double* host, int len);
/* run_store_energy, kernel_interface.h:28-29 */
int tl_run_store_energy(tl_chunk* c);
/* run_field_summary, kernel_interface.h:30-32 (values are ASSIGNED, field_summary.cpp:147-150) */
int tl_run_field_summary(tl_chunk* c, double* vol, double* mass, double* ie, double* temp);

/* ---- CG (kernel_interface.h:35-43) ---- */
int tl_run_cg_init(tl_chunk* c, int coefficient, double rx, double ry, double* rro); /* *rro += */
int tl_run_cg_calc_w(tl_chunk* c, double* pw);                                      /* *pw  += */
int tl_run_cg_calc_ur(tl_chunk* c, double alpha, double* rrn);                      /* *rrn  = */
int tl_run_cg_calc_p(tl_chunk* c, double beta);
/* ---- Chebyshev (kernel_interface.h:46-49); theta is chunk->theta ---- */
int tl_run_cheby_init(tl_chunk* c, double theta);
int tl_run_cheby_iterate(tl_chunk* c, double alpha, double beta);
/* ---- Jacobi (kernel_interface.h:52-55) ---- */
int tl_run_jacobi_init(tl_chunk* c, int coefficient, double rx, double ry);
int tl_run_jacobi_iterate(tl_chunk* c, double* error); /* *error = */
/* ---- PPCG (kernel_interface.h:58-61) ---- */
int tl_run_ppcg_init(tl_chunk* c, double theta);
int tl_run_ppcg_inner_iteration(tl_chunk* c, double alpha, double beta);
/* ---- shared solver kernels (kernel_interface.h:64-71) ---- */
int tl_run_copy_u(tl_chunk* c);
int tl_run_calculate_residual(tl_chunk* c);
int tl_run_calculate_2norm(tl_chunk* c, int field, double* norm); /* *norm = */
int tl_run_finalise(tl_chunk* c);

/* ---- comms layer: replaces TeaLeaf/comms.h:10-20 (MPI) for the ranks of ONE node ----
 * One process per GPU. Rendezvous goes through a POSIX shared-memory segment named by
 * `session` (e.g. the torchrun MASTER_PORT); halo payloads and scalars then move GPU to GPU
 * over NVLink through CUDA-IPC mapped peer buffers (or through the shared segment when
 * `host_only` != 0: the no-GPU path used by the CPU tests of the host logic). */
int tl_comms_create(tl_comms** out, const char* session, int rank, int num_ranks, int device, int host_only);
int tl_comms_destroy(tl_comms* k);
int tl_comms_rank(const tl_comms* k);
int tl_comms_size(const tl_comms* k);
int tl_comms_barrier(tl_comms* k);                     /* barrier(),        comms.h:10 */
int tl_comms_abort(tl_comms* k);                       /* abort_comms(),    comms.h:13 (MPI_Abort): the other ranks'
                                                        * host-side waits fail with TL_ERR_COMMS at once */
int tl_comms_sum(tl_comms* k, double* a);              /* sum_over_ranks,   comms.h:14; rank-ordered sum */
int tl_comms_min(tl_comms* k, double* a);              /* min_over_ranks,   comms.h:15 */
/* send_recv_message + wait_for_requests (comms.h:16-20) on HOST buffers: posts my message for
 * `neighbour` under send_tag and receives the neighbour's message posted under recv_tag. */
int tl_comms_send_recv(tl_comms* k, const double* send_buffer, double* recv_buffer, int buffer_len,
                       int neighbour, int send_tag, int recv_tag);
/* The two halves separately (MPI_Isend / MPI_Irecv+MPI_Wait): post every message of a phase, then
 * receive them, as wait_for_requests does after remote_halo_driver.c:32-55 posted them. */
int tl_comms_post(tl_comms* k, const double* send_buffer, int buffer_len, int neighbour, int send_tag);
int tl_comms_recv(tl_comms* k, double* recv_buffer, int buffer_len, int neighbour, int recv_tag);
/* Attach a chunk: exchanges CUDA-IPC handles of its face staging buffers / scalar slots with the
 * neighbouring ranks so that halo_update and the solver loops move data peer to peer. */
int tl_comms_attach_chunk(tl_comms* k, tl_chunk* c);
/* Decomposition of initialise.c:34-134 for chunk == rank: returns nx, ny, left, bottom, neighbours. */
int tl_decompose(int grid_x_cells, int grid_y_cells, int num_chunks, int chunk,
                 int* nx, int* ny, int* left, int* bottom, int neighbours[4],
                 int* x_chunks, int* y_chunks);

/* halo_update_driver (drivers/halo_update_driver.c:6-25): remote exchange (L/R then B/T, each
 * pack -> peer copy -> unpack) followed by the local reflective update. k may be NULL for one rank. */
int tl_halo_update(tl_chunk* c, tl_comms* k, const int fields_to_exchange[6], int depth);

/* Test hook: `reps` back-to-back halo exchanges with no host synchronisation in between, the field contents
 * regenerated and every halo cell verified ON THE DEVICE each time (neighbour's cell or reflection at the edge of
 * the grid_x_cells x grid_y_cells mesh, as remote_halo_driver.c:24-126 + local_halos.cpp produce them).
 * *mismatches = number of wrong halo cells over all exchanges.  Environment TL_TEST_SKEW="rank:usec" delays that
 * rank's host between the send and the unpack launches of every exchange phase. */
int tl_halo_stress(tl_chunk* c, tl_comms* k, int grid_x_cells, int grid_y_cells, int reps, int depth,
                   long* mismatches);

/* ---- device-resident solver loops ---- */
typedef struct {
    int solver;            /* TL_SOLVER_* */
    int coefficient;       /* TL_CONDUCTIVITY */
    int max_iters;         /* settings.h:26 */
    double eps;            /* settings.h:27 */
    int presteps;          /* settings.h:33 */
    int ppcg_inner_steps;  /* settings.h:36 */
    int error_switch;      /* settings.h:32 */
    double eps_lim;        /* settings.h:34 */
    int check_result;      /* settings.h:35: residual + 2norm in solve_finished */
    int fuse_p_into_w;     /* 0: three kernels per CG iteration as in cg_driver.c (p's halo travels between ranks);
                            * non-zero (default 1): p-update fused into the next matvec, two kernels per iteration, 96
                            *    instead of 104 B/cell moved (between ranks r's halo travels).  Non-zero also selects the
                            *    one-pass Chebyshev / PPCG kernels.  Results are bit-identical in every mode. */
    int batch;             /* iterations enqueued between convergence polls (0 = default) */
} tl_solve_opts;

typedef struct {
    int iters_a;      /* printed "CG:" / "Jacobi:" count (cg_driver.c:27, cheby_driver.c:73, ppcg_driver.c:59) */
    int iters_b;      /* printed Cheby / PPCG count */
    int est_iters;    /* cheby_driver.c:74-76 */
    int total_iters;  /* solver-loop iterations executed */
    double error;
    double eigmin, eigmax;
    double gpu_ms;    /* CUDA-event time of the solver loop */
    long kernel_launches;
} tl_solve_info;

void tl_solve_opts_default(tl_solve_opts* o);
/* 1 if the resident CG loop of this chunk runs the fused two-kernel iteration for the given fuse_p_into_w. */
int tl_cg_loop_is_fused(const tl_chunk* c, int fuse_p_into_w);
/* One timestep's linear solve: the body of solve() in diffuse.c:23-64 between the depth-2
 * halo update and solve_finished_driver, i.e. cg_driver / cheby_driver / ppcg_driver /
 * jacobi_driver with the same call order, iteration counts and results. */
int tl_solve(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double rx, double ry, tl_solve_info* info);
/* One full timestep (diffuse.c:23-78 without the printing): depth-2 halo of energy+density,
 * tl_solve, solve_finished_driver (residual/2norm if check_result, finalise, energy halo). */
int tl_timestep(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double dt, double dx, double dy,
                tl_solve_info* info);
/* field_summary_driver.c:8-30: run_field_summary + sum_over_ranks of the four values. */
int tl_field_summary(tl_chunk* c, tl_comms* k, double* vol, double* mass, double* ie, double* temp);
/* End-to-end host-buffer step used by bench.py's e2e leg: copies dense host images of density and
 * energy (pinned or pageable) into HBM, runs tl_timestep, copies energy back and returns the
 * field summary. */
int tl_timestep_host(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double dt, double dx, double dy,
                     const double* density_host, double* energy_host, tl_solve_info* info,
                     double summary[4]);
void* tl_host_alloc_pinned(long bytes);
void tl_host_free_pinned(void* p);

/* Kernel micro-benchmark hook used by bench.py's roofline leg: runs `reps` back-to-back launches of
 * one hot kernel on the chunk's stream and returns the average CUDA-event milliseconds per launch.
 * which: 0 cg_calc_w, 1 cg_calc_ur, 2 cg_calc_p, 3 fused p+w (overwrites p, w and the solver scalars),
 * 4 cheby_iterate, 5 cheby_calc_u, 6 ppcg_calc_ur, 7 ppcg_calc_sd, 8 jacobi_iterate (copy + iterate),
 * 9 calculate_residual, 10 calculate_2norm, 11 field_summary, 12 cg_init (one pass), 13 finalise,
 * 14 copy_u, 15 cheby_init, 16 fused cheby iteration, 17 fused ppcg inner iteration.
 * The timed kernels overwrite solver fields. */
int tl_time_kernel(tl_chunk* c, int which, int reps, double* ms_per_launch);
/* Tuning: rows per tile and rows per load batch (1, 2 or 4) of a hot kernel family
 * (0 cg_calc_w, 1 cg_calc_ur, 2 cg_calc_p, 3 fused p+w). Results do not depend on `batch`;
 * reductions depend on `rows` only through the (deterministic) summation order. */
int tl_set_tuning(int kernel, int rows, int batch);
/* The fused p-update + matvec kernel of the resident CG loop: mode 0 (default) = register-staged kernel, mode 1 =
 * TMA bulk-copy (cp.async.bulk) shared-memory row pipeline with persistent CTAs (measured equal, see tl_bulk.cu).  stages in {3,4,6,8} = rows in
 * flight per CTA; ctas_per_sm 0 = as many as the shared-memory footprint allows (at most 8).  Results are
 * bit-identical in every setting. */
int tl_set_pw_pipeline(int mode, int stages, int ctas_per_sm);
/* Profiling aid: the kernels of the resident CG loop record %globaltimer (ns) at four points -- first CTA past its
 * dependency wait, tail CTA holds this rank's partial sum, all ranks' partials gathered, halo hand-shake done -- for the
 * first `iterations` iterations of every solve (0 disables).  tl_stamps_read copies out
 * [iteration][kernel: 0 matvec, 1 calc_ur, 2 calc_p][point 0..3] and clears the buffer. */
int tl_stamps_enable(tl_chunk* c, int iterations);
int tl_stamps_read(tl_chunk* c, unsigned long long* host, int iterations);
long tl_kernel_launch_count(void); /* kernels launched by this library in this process */
/* CUDA-event timer on the chunk's own stream (torch.cuda.Event only sees torch's stream). */
int tl_timer_start(tl_chunk* c);
int tl_timer_stop(tl_chunk* c, double* elapsed_ms); /* records, synchronises, returns ms since start */

#ifdef __cplusplus
}
#endif
#endif
