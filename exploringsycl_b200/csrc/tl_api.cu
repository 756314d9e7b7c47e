// tl_api.cu -- C-ABI: chunk lifecycle, field I/O and the 22 run_* mirrors of
// TeaLeaf/kernel_interface.h:13-71.  Host-driven ("plugin") path: every reduction is returned
// synchronously through a double*, with the reference's accumulate-vs-assign convention
// (+= for rro, pw; = for rrn, norm, error, summary).
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "tl_internal.h"

static thread_local char g_err[512] = "";

void tl_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int tl_cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    tl_set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return TL_ERR_CUDA;
}

extern "C" const char* tl_last_error(void) { return g_err; }
extern "C" const char* tl_version(void) { return "tealeaf_b200 0.1 (sm_100a)"; }
extern "C" long tl_kernel_launch_count(void) { return g_tl_launches; }
extern "C" int tl_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int dev_zalloc(double** p, size_t elems)
{
    TL_CUDA(cudaMalloc((void**)p, elems * sizeof(double)));
    TL_CUDA(cudaMemset(*p, 0, elems * sizeof(double)));
    return TL_OK;
}

extern "C" int tl_chunk_create(tl_chunk** out, int device, int nx, int ny, int halo_depth, int max_iters,
                               const int neighbours[4], int left, int bottom)
{
    TL_CHECK_ARG(out && nx > 0 && ny > 0 && halo_depth >= 1 && max_iters > 0, "bad arguments");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        tl_set_error("tl_chunk_create: no CUDA device available (%s); this backend has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return TL_ERR_CUDA;
    }
    TL_CHECK_ARG(device >= 0 && device < ndev, "device index out of range");
    TL_CUDA(cudaSetDevice(device));
    tl_chunk* c = new tl_chunk();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->nx = nx;
    c->ny = ny;
    c->max_iters = max_iters;
    c->left = left;
    c->bottom = bottom;
    for (int i = 0; i < 4; ++i) c->nb[i] = neighbours ? neighbours[i] : TL_EXTERNAL_FACE;
    Geo& g = c->g;
    g.hd = halo_depth;
    g.x = nx + 2 * halo_depth; // chunk.c:7-8
    g.y = ny + 2 * halo_depth;
    g.pitch = (g.x + 15) / 16 * 16;
    g.off = (16 - halo_depth % 16) % 16;
    c->field_elems = ((size_t)g.off + (size_t)g.y * g.pitch + 64 + 31) / 32 * 32; // 256-byte multiple
    // One slab for all fields, a 2 MiB multiple: it is its own allocation block, so one CUDA-IPC handle
    // maps every field of this chunk into the neighbouring ranks (tl_comms_attach_chunk).
    c->slab_bytes = (c->field_elems * sizeof(double) * TL_SLAB_SLOTS + (2u << 20) - 1) / (2u << 20) * (2u << 20);
    TL_CUDA(cudaMalloc((void**)&c->slab, c->slab_bytes));
    TL_CUDA(cudaMemset(c->slab, 0, c->slab_bytes));
    for (int f = 0; f < TL_NUM_FIELDS; ++f) c->f[f] = c->slab + (size_t)f * c->field_elems;
    c->p2 = c->slab + (size_t)TL_SLAB_P2 * c->field_elems;
    c->alt[TL_FIELD_U] = c->slab + (size_t)TL_SLAB_U2 * c->field_elems;
    c->alt[TL_FIELD_SD] = c->slab + (size_t)TL_SLAB_SD2 * c->field_elems;
    TL_TRY(dev_zalloc(&c->cell_x, g.x + 2));
    TL_TRY(dev_zalloc(&c->cell_y, g.y + 2));
    TL_TRY(dev_zalloc(&c->vertex_x, g.x + 2));
    TL_TRY(dev_zalloc(&c->vertex_y, g.y + 2));
    c->partial_cap = ((g.x + TL_TPB - 1) / TL_TPB) * ((g.y + 7) / 8) + 64;
    TL_TRY(dev_zalloc(&c->partials, (size_t)c->partial_cap * 4));
    c->gpartial_cap = c->partial_cap / 64 + 2;
    TL_TRY(dev_zalloc(&c->gpartials, (size_t)c->gpartial_cap * 4));
    TL_CUDA(cudaMalloc((void**)&c->gcount, sizeof(unsigned int) * c->gpartial_cap));
    TL_CUDA(cudaMemset(c->gcount, 0, sizeof(unsigned int) * c->gpartial_cap));
    TL_CUDA(cudaMalloc((void**)&c->scal, sizeof(DevScal)));
    TL_CUDA(cudaMemset(c->scal, 0, sizeof(DevScal)));
    TL_CUDA(cudaMallocHost((void**)&c->scal_h, 3 * sizeof(DevScal)));
    memset(c->scal_h, 0, 3 * sizeof(DevScal));
    {   // host-mapped error word: device-side peer waits that time out mark it (spin_flag in tl_kernels.cu)
        TL_CUDA(cudaHostAlloc((void**)&c->err_h, 64, cudaHostAllocMapped));
        *c->err_h = 0u;
        unsigned int* dptr = nullptr;
        TL_CUDA(cudaHostGetDevicePointer((void**)&dptr, c->err_h, 0));
        TL_CUDA(cudaMemcpy((char*)c->scal + offsetof(DevScal, err_host), &dptr, sizeof(dptr), cudaMemcpyHostToDevice));
    }
    TL_TRY(dev_zalloc(&c->colbuf, 2 * (size_t)g.y + 64));
    TL_TRY(dev_zalloc(&c->d_alphas, max_iters + 1));
    TL_TRY(dev_zalloc(&c->d_betas, max_iters + 1));
    // kernel_initialise.cpp:76-79: host coefficient arrays of max_iters doubles, zeroed
    c->cg_alphas = (double*)calloc(max_iters + 1, sizeof(double));
    c->cg_betas = (double*)calloc(max_iters + 1, sizeof(double));
    c->cheby_alphas = (double*)calloc(max_iters + 1, sizeof(double));
    c->cheby_betas = (double*)calloc(max_iters + 1, sizeof(double));
    // face staging: NUM_FIELDS * halo_depth * max(x,y) doubles per face (chunk.c:15-24)
    c->face_elems = (size_t)TL_NUM_EXCHANGE_FIELDS * halo_depth * (size_t)(g.x > g.y ? g.x : g.y);
    for (int fc = 0; fc < 4; ++fc) {
        TL_TRY(dev_zalloc(&c->face_send[fc], c->face_elems));
        TL_TRY(dev_zalloc(&c->face_recv[fc], c->face_elems));
    }
    TL_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    TL_CUDA(cudaEventCreate(&c->ev0));
    TL_CUDA(cudaEventCreate(&c->ev1));
    TL_CUDA(cudaDeviceSynchronize());
    *out = c;
    return TL_OK;
}

extern "C" int tl_chunk_destroy(tl_chunk* c)
{
    if (!c) return TL_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->slab);
    cudaFree(c->cell_x); cudaFree(c->cell_y); cudaFree(c->vertex_x); cudaFree(c->vertex_y);
    cudaFree(c->partials); cudaFree(c->gpartials); cudaFree(c->gcount); cudaFree(c->scal); cudaFreeHost(c->scal_h); cudaFreeHost(c->err_h);
    cudaFree(c->d_alphas); cudaFree(c->d_betas); cudaFree(c->colbuf);
    if (c->stamps) cudaFree(c->stamps);
    free(c->cg_alphas); free(c->cg_betas); free(c->cheby_alphas); free(c->cheby_betas);
    for (int fc = 0; fc < 4; ++fc) { cudaFree(c->face_send[fc]); cudaFree(c->face_recv[fc]); }
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return TL_OK;
}

extern "C" int tl_chunk_dims(const tl_chunk* c, int* x, int* y, int* halo_depth, int* pitch)
{
    TL_CHECK_ARG(c, "null chunk");
    if (x) *x = c->g.x;
    if (y) *y = c->g.y;
    if (halo_depth) *halo_depth = c->g.hd;
    if (pitch) *pitch = c->g.pitch;
    return TL_OK;
}

extern "C" int tl_chunk_sync(tl_chunk* c)
{
    TL_CHECK_ARG(c, "null chunk");
    TL_CUDA(cudaSetDevice(c->device));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return tl_check_peer_timeout(c);
}

extern "C" int tl_field_write(tl_chunk* c, int field, const double* host)
{
    TL_CHECK_ARG(c && host && field >= 0 && field < TL_NUM_FIELDS, "bad arguments");
    TL_CUDA(cudaSetDevice(c->device));
    const Geo& g = c->g;
    TL_CUDA(cudaMemcpy2DAsync(c->f[field] + g.off, (size_t)g.pitch * 8, host, (size_t)g.x * 8,
                              (size_t)g.x * 8, g.y, cudaMemcpyHostToDevice, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return TL_OK;
}

extern "C" int tl_field_read(tl_chunk* c, int field, double* host)
{
    TL_CHECK_ARG(c && host && field >= 0 && field < TL_NUM_FIELDS, "bad arguments");
    TL_CUDA(cudaSetDevice(c->device));
    const Geo& g = c->g;
    TL_CUDA(cudaMemcpy2DAsync(host, (size_t)g.x * 8, c->f[field] + g.off, (size_t)g.pitch * 8,
                              (size_t)g.x * 8, g.y, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return tl_check_peer_timeout(c);
}

extern "C" int tl_array_read(tl_chunk* c, int array, double* host)
{
    TL_CHECK_ARG(c && host && array >= 0 && array <= TL_ARRAY_VERTEX_Y, "bad arguments");
    TL_CUDA(cudaSetDevice(c->device));
    const double* src[4] = {c->cell_x, c->cell_y, c->vertex_x, c->vertex_y};
    const int len[4] = {c->g.x, c->g.y, c->g.x + 1, c->g.y + 1};
    TL_CUDA(cudaMemcpyAsync(host, src[array], (size_t)len[array] * 8, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return TL_OK;
}

// Resident-loop time stamps (DevScal.stamps, tl_internal.h): %globaltimer at four points of every loop kernel.
extern "C" int tl_stamps_enable(tl_chunk* c, int iterations)
{
    TL_CHECK_ARG(c && iterations >= 0, "bad arguments");
    TL_CUDA(cudaSetDevice(c->device));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    if (c->stamps) cudaFree(c->stamps);
    c->stamps = nullptr;
    c->stamp_cap = 0;
    if (iterations > 0) {
        const size_t n = (size_t)iterations * TL_STAMP_KERNELS * TL_STAMP_POINTS;
        TL_CUDA(cudaMalloc((void**)&c->stamps, n * sizeof(unsigned long long)));
        TL_CUDA(cudaMemset(c->stamps, 0, n * sizeof(unsigned long long)));
        c->stamp_cap = iterations;
    }
    TL_CUDA(cudaMemcpy((char*)c->scal + offsetof(DevScal, stamps), &c->stamps, sizeof(c->stamps), cudaMemcpyHostToDevice));
    TL_CUDA(cudaMemcpy((char*)c->scal + offsetof(DevScal, stamp_cap), &c->stamp_cap, sizeof(int), cudaMemcpyHostToDevice));
    return TL_OK;
}
extern "C" int tl_stamps_read(tl_chunk* c, unsigned long long* host, int iterations)
{
    TL_CHECK_ARG(c && host && iterations >= 0 && iterations <= c->stamp_cap, "bad arguments");
    TL_CUDA(cudaSetDevice(c->device));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)iterations * TL_STAMP_KERNELS * TL_STAMP_POINTS;
    TL_CUDA(cudaMemcpy(host, c->stamps, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    TL_CUDA(cudaMemset(c->stamps, 0, n * sizeof(unsigned long long)));
    return TL_OK;
}

extern "C" double* tl_cg_alphas(tl_chunk* c) { return c ? c->cg_alphas : nullptr; }
extern "C" double* tl_cg_betas(tl_chunk* c) { return c ? c->cg_betas : nullptr; }
extern "C" double* tl_cheby_alphas(tl_chunk* c) { return c ? c->cheby_alphas : nullptr; }
extern "C" double* tl_cheby_betas(tl_chunk* c) { return c ? c->cheby_betas : nullptr; }

int tl_check_peer_timeout(tl_chunk* c)
{
    if (c->err_h && *(volatile unsigned int*)c->err_h == 0xdeadu) {
        tl_set_error("a device-side wait for a peer rank timed out (halo exchange or resident solver loop): "
                     "the results of this chunk are not valid");
        return TL_ERR_COMMS;
    }
    return TL_OK;
}

int tl_fetch_scal(tl_chunk* c)
{
    TL_CUDA(cudaMemcpyAsync(c->scal_h, c->scal, sizeof(DevScal), cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return tl_check_peer_timeout(c);
}

#define ENTER(c)                                  \
    TL_CHECK_ARG(c, "null chunk");                \
    TL_CUDA(cudaSetDevice((c)->device))

// kernel_interface.cpp:34-52
extern "C" int tl_run_set_chunk_data(tl_chunk* c, double grid_x_min, double grid_y_min, double dx, double dy)
{
    ENTER(c);
    const double x_min = grid_x_min + dx * (double)c->left;
    const double y_min = grid_y_min + dy * (double)c->bottom;
    return tlk_set_chunk_data(c, x_min, y_min, dx, dy);
}

// kernel_interface.cpp:54-72
extern "C" int tl_run_set_chunk_state(tl_chunk* c, int num_states, const tl_state* states)
{
    ENTER(c);
    TL_CHECK_ARG(num_states >= 1 && states, "need at least one state");
    TL_TRY(tlk_set_initial_state(c, states[0].energy, states[0].density));
    for (int ii = 1; ii < num_states; ++ii) TL_TRY(tlk_set_state(c, &states[ii]));
    return TL_OK;
}

extern "C" int tl_run_local_halos(tl_chunk* c, const int fields_to_exchange[6], int depth)
{
    ENTER(c);
    TL_CHECK_ARG(fields_to_exchange && depth >= 1 && depth <= c->g.hd, "bad depth");
    return tlk_local_halos(c, fields_to_exchange, depth);
}

// kernel_interface.cpp:129-167: device gather/scatter + mirror copy to/from the HOST buffer
extern "C" int tl_run_pack_or_unpack(tl_chunk* c, int depth, int face, int pack, int field, double* host_buffer)
{
    ENTER(c);
    TL_CHECK_ARG(host_buffer && depth >= 1 && depth <= c->g.hd && face >= 0 && face < 4 && field >= 0 &&
                     field < TL_NUM_FIELDS, "bad arguments");
    int fields[TL_NUM_EXCHANGE_FIELDS] = {0, 0, 0, 0, 0, 0};
    // any 2-D field can be packed (the reference passes a FieldBufferType); route through slot 0
    double* saved = c->f[0];
    c->f[0] = c->f[field];
    fields[0] = 1;
    int len = 0, rc;
    const bool lr = (face == TL_FACE_LEFT || face == TL_FACE_RIGHT);
    const size_t bytes = (size_t)depth * (lr ? c->g.y : c->g.x) * sizeof(double);
    if (pack) {
        rc = tlk_pack_face(c, fields, depth, face, true, c->face_send[face], &len);
        c->f[0] = saved;
        TL_TRY(rc);
        TL_CUDA(cudaMemcpyAsync(host_buffer, c->face_send[face], bytes, cudaMemcpyDeviceToHost, c->stream));
        TL_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        cudaError_t e = cudaMemcpyAsync(c->face_recv[face], host_buffer, bytes, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { c->f[0] = saved; return tl_cuda_fail(e, "H2D", __FILE__, __LINE__); }
        rc = tlk_pack_face(c, fields, depth, face, false, c->face_recv[face], &len);
        c->f[0] = saved;
        TL_TRY(rc);
        TL_CUDA(cudaStreamSynchronize(c->stream));
    }
    return TL_OK;
}

extern "C" int tl_pack_face_device(tl_chunk* c, const int fields_to_exchange[6], int depth, int face, int pack,
                                   int* len)
{
    ENTER(c);
    TL_CHECK_ARG(fields_to_exchange && depth >= 1 && depth <= c->g.hd && face >= 0 && face < 4, "bad arguments");
    return tlk_pack_face(c, fields_to_exchange, depth, face, pack != 0,
                         pack ? c->face_send[face] : c->face_recv[face], len);
}

extern "C" int tl_face_buffer_read(tl_chunk* c, int face, int send, double* host, int len)
{
    ENTER(c);
    TL_CHECK_ARG(host && face >= 0 && face < 4 && len >= 0 && (size_t)len <= c->face_elems, "bad arguments");
    TL_CUDA(cudaMemcpyAsync(host, send ? c->face_send[face] : c->face_recv[face], (size_t)len * 8,
                            cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return TL_OK;
}

extern "C" int tl_face_buffer_write(tl_chunk* c, int face, int send, const double* host, int len)
{
    ENTER(c);
    TL_CHECK_ARG(host && face >= 0 && face < 4 && len >= 0 && (size_t)len <= c->face_elems, "bad arguments");
    TL_CUDA(cudaMemcpyAsync(send ? c->face_send[face] : c->face_recv[face], host, (size_t)len * 8,
                            cudaMemcpyHostToDevice, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return TL_OK;
}

extern "C" int tl_run_store_energy(tl_chunk* c)
{
    ENTER(c);
    return tlk_copy_field(c, TL_FIELD_ENERGY1, TL_FIELD_ENERGY0, false);
}

extern "C" int tl_run_field_summary(tl_chunk* c, double* vol, double* mass, double* ie, double* temp)
{
    ENTER(c);
    TL_CHECK_ARG(vol && mass && ie && temp, "null output");
    TL_TRY(tlk_field_summary(c));
    TL_TRY(tl_fetch_scal(c));
    *vol = c->scal_h->sums[0];
    *mass = c->scal_h->sums[1];
    *ie = c->scal_h->sums[2];
    *temp = c->scal_h->sums[3];
    return TL_OK;
}

extern "C" int tl_run_cg_init(tl_chunk* c, int coefficient, double rx, double ry, double* rro)
{
    ENTER(c);
    TL_CHECK_ARG(rro, "null output");
    TL_TRY(tlk_cg_init(c, coefficient, rx, ry));
    TL_TRY(tl_fetch_scal(c));
    *rro += c->scal_h->sums[0]; // cg.cpp:133
    return TL_OK;
}

extern "C" int tl_run_cg_calc_w(tl_chunk* c, double* pw)
{
    ENTER(c);
    TL_CHECK_ARG(pw, "null output");
    TL_TRY(tlk_cg_calc_w(c, SCAL_IMM, false));
    TL_TRY(tl_fetch_scal(c));
    *pw += c->scal_h->pw; // cg.cpp:194
    return TL_OK;
}

extern "C" int tl_run_cg_calc_ur(tl_chunk* c, double alpha, double* rrn)
{
    ENTER(c);
    TL_CHECK_ARG(rrn, "null output");
    TL_TRY(tlk_cg_calc_ur(c, SCAL_IMM, alpha, false));
    TL_TRY(tl_fetch_scal(c));
    *rrn = c->scal_h->rrn; // cg.cpp:253
    return TL_OK;
}

extern "C" int tl_run_cg_calc_p(tl_chunk* c, double beta)
{
    ENTER(c);
    return tlk_cg_calc_p(c, SCAL_IMM, beta, false, false);
}

extern "C" int tl_run_cheby_init(tl_chunk* c, double theta)
{
    ENTER(c);
    return tlk_cheby_init(c, theta);
}

// kernel_interface.cpp:258-271: cheby_iterate then cheby_calc_u
extern "C" int tl_run_cheby_iterate(tl_chunk* c, double alpha, double beta)
{
    ENTER(c);
    TL_TRY(tlk_cheby_iterate(c, alpha, beta));
    return tlk_cheby_calc_u(c);
}

extern "C" int tl_run_jacobi_init(tl_chunk* c, int coefficient, double rx, double ry)
{
    ENTER(c);
    return tlk_jacobi_init(c, coefficient, rx, ry);
}

extern "C" int tl_run_jacobi_iterate(tl_chunk* c, double* error)
{
    ENTER(c);
    TL_CHECK_ARG(error, "null output");
    TL_TRY(tlk_jacobi_iterate(c));
    TL_TRY(tl_fetch_scal(c));
    *error = c->scal_h->sums[0]; // jacobi.cpp:116
    return TL_OK;
}

extern "C" int tl_run_ppcg_init(tl_chunk* c, double theta)
{
    ENTER(c);
    return tlk_ppcg_init(c, theta);
}

// kernel_interface.cpp:314-328: ppcg_calc_ur then ppcg_calc_sd
extern "C" int tl_run_ppcg_inner_iteration(tl_chunk* c, double alpha, double beta)
{
    ENTER(c);
    TL_TRY(tlk_ppcg_calc_ur(c));
    return tlk_ppcg_calc_sd(c, alpha, beta);
}

extern "C" int tl_run_copy_u(tl_chunk* c)
{
    ENTER(c);
    return tlk_copy_field(c, TL_FIELD_U0, TL_FIELD_U, true);
}

extern "C" int tl_run_calculate_residual(tl_chunk* c)
{
    ENTER(c);
    return tlk_calculate_residual(c);
}

extern "C" int tl_run_calculate_2norm(tl_chunk* c, int field, double* norm)
{
    ENTER(c);
    TL_CHECK_ARG(norm && field >= 0 && field < TL_NUM_FIELDS, "bad arguments");
    TL_TRY(tlk_calculate_2norm(c, field));
    TL_TRY(tl_fetch_scal(c));
    *norm = c->scal_h->sums[0]; // solver_methods.cpp:116
    return TL_OK;
}

extern "C" int tl_run_finalise(tl_chunk* c)
{
    ENTER(c);
    return tlk_finalise(c);
}

extern "C" void* tl_host_alloc_pinned(long bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void tl_host_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}
