/*
 * kernel_interface.cpp -- the 22 run_* entry points of TeaLeaf/kernel_interface.h:13-71 for the
 * B200 CUDA backend, with C++ linkage exactly like the reference's own
 * TeaLeaf/c_kernels/sycl/kernel_interface.cpp:34-371.  Each is a thin call into the C-ABI of
 * libtealeaf_b200.so (include/tealeaf_b200.h); failures go through die() (shared.c:98-111), the
 * reference's error convention.
 */
#include <stdlib.h>
#include "kernel_interface.h"
#include "shared.h"
#include "tealeaf_b200.h"

#define TLX(call) do { if ((call) != TL_OK) die(__LINE__, __FILE__, "%s\n", tl_last_error()); } while (0)
#define H(chunk) ((chunk)->ext->handle)

static void flags(Settings* settings, int* f)
{
    for (int ii = 0; ii < NUM_FIELDS; ++ii) f[ii] = settings->fields_to_exchange[ii] ? 1 : 0;
}

// kernel_initialise.cpp:31-80: allocate every field (zeroed) and the host coefficient arrays
void run_kernel_initialise(Chunk* chunk, Settings* settings)
{
    print_and_log(settings, "Performing this solve with the B200 CUDA %s solver\n", settings->solver_name);
    const int hd = settings->halo_depth;
    int ndev = tl_device_count();
    if (ndev < 1) die(__LINE__, __FILE__, "no CUDA device: the B200 backend has no CPU fallback\n");
    TLX(tl_chunk_create(&chunk->ext->handle, settings->rank % ndev, chunk->x - 2*hd, chunk->y - 2*hd, hd,
                        settings->max_iters, chunk->neighbours, chunk->left, chunk->bottom));
    for (int ii = 0; ii < 12; ++ii) chunk->ext->refs[ii].id = ii;
    TlFieldRef* r = chunk->ext->refs;
    chunk->density = &r[TL_FIELD_DENSITY]; chunk->energy0 = &r[TL_FIELD_ENERGY0];
    chunk->energy = &r[TL_FIELD_ENERGY1];  chunk->u = &r[TL_FIELD_U];   chunk->p = &r[TL_FIELD_P];
    chunk->sd = &r[TL_FIELD_SD];           chunk->u0 = &r[TL_FIELD_U0]; chunk->r = &r[TL_FIELD_R];
    chunk->w = &r[TL_FIELD_W];             chunk->kx = &r[TL_FIELD_KX]; chunk->ky = &r[TL_FIELD_KY];
    chunk->volume = &r[TL_FIELD_VOLUME];
    // never touched by any kernel on the path: not allocated
    chunk->density0 = chunk->mi = chunk->x_area = chunk->y_area = NULL;
    chunk->cell_x = chunk->cell_y = chunk->cell_dx = chunk->cell_dy = NULL;
    chunk->vertex_x = chunk->vertex_y = chunk->vertex_dx = chunk->vertex_dy = NULL;
    // the host reads and writes these directly (cg_driver.c:93,111; cheby_driver.c:178-179)
    chunk->cg_alphas = tl_cg_alphas(H(chunk));       chunk->cg_betas = tl_cg_betas(H(chunk));
    chunk->cheby_alphas = tl_cheby_alphas(H(chunk)); chunk->cheby_betas = tl_cheby_betas(H(chunk));
    extern void comms_b200_attach(Chunk* chunk);
    comms_b200_attach(chunk);
}

void run_kernel_finalise(Chunk* chunk, Settings* settings)
{
    (void)settings;
    TLX(tl_chunk_destroy(H(chunk)));
}

void run_set_chunk_data(Chunk* chunk, Settings* settings)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_set_chunk_data(H(chunk), settings->grid_x_min, settings->grid_y_min, settings->dx, settings->dy));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_set_chunk_state(Chunk* chunk, Settings* settings, State* states)
{
    START_PROFILING(settings->kernel_profile);
    tl_state* st = (tl_state*)malloc(sizeof(tl_state) * settings->num_states);
    for (int ii = 0; ii < settings->num_states; ++ii) {
        st[ii].geometry = (int)states[ii].geometry;
        st[ii].density = states[ii].density; st[ii].energy = states[ii].energy;
        st[ii].x_min = states[ii].x_min; st[ii].y_min = states[ii].y_min;
        st[ii].x_max = states[ii].x_max; st[ii].y_max = states[ii].y_max;
        st[ii].radius = states[ii].radius;
    }
    TLX(tl_run_set_chunk_state(H(chunk), settings->num_states, st));
    free(st);
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_local_halos(Chunk* chunk, Settings* settings, int depth)
{
    START_PROFILING(settings->kernel_profile);
    int f[NUM_FIELDS];
    flags(settings, f);
    TLX(tl_run_local_halos(H(chunk), f, depth));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_pack_or_unpack(Chunk* chunk, Settings* settings, int depth, int face, bool pack,
                        FieldBufferType field, double* buffer)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_pack_or_unpack(H(chunk), depth, face, pack ? 1 : 0, field->id, buffer));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_store_energy(Chunk* chunk, Settings* settings)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_store_energy(H(chunk)));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_field_summary(Chunk* chunk, Settings* settings, double* vol, double* mass, double* ie, double* temp)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_field_summary(H(chunk), vol, mass, ie, temp));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_cg_init(Chunk* chunk, Settings* settings, double rx, double ry, double* rro)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_cg_init(H(chunk), settings->coefficient, rx, ry, rro));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_cg_calc_w(Chunk* chunk, Settings* settings, double* pw)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_cg_calc_w(H(chunk), pw));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_cg_calc_ur(Chunk* chunk, Settings* settings, double alpha, double* rrn)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_cg_calc_ur(H(chunk), alpha, rrn));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_cg_calc_p(Chunk* chunk, Settings* settings, double beta)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_cg_calc_p(H(chunk), beta));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_cheby_init(Chunk* chunk, Settings* settings)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_cheby_init(H(chunk), chunk->theta));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_cheby_iterate(Chunk* chunk, Settings* settings, double alpha, double beta)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_cheby_iterate(H(chunk), alpha, beta));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_jacobi_init(Chunk* chunk, Settings* settings, double rx, double ry)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_jacobi_init(H(chunk), settings->coefficient, rx, ry));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_jacobi_iterate(Chunk* chunk, Settings* settings, double* error)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_jacobi_iterate(H(chunk), error));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_ppcg_init(Chunk* chunk, Settings* settings)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_ppcg_init(H(chunk), chunk->theta));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_ppcg_inner_iteration(Chunk* chunk, Settings* settings, double alpha, double beta)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_ppcg_inner_iteration(H(chunk), alpha, beta));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_copy_u(Chunk* chunk, Settings* settings)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_copy_u(H(chunk)));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_calculate_residual(Chunk* chunk, Settings* settings)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_calculate_residual(H(chunk)));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_calculate_2norm(Chunk* chunk, Settings* settings, FieldBufferType buffer, double* norm)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_calculate_2norm(H(chunk), buffer->id, norm));
    STOP_PROFILING(settings->kernel_profile, __func__);
}

void run_finalise(Chunk* chunk, Settings* settings)
{
    START_PROFILING(settings->kernel_profile);
    TLX(tl_run_finalise(H(chunk)));
    STOP_PROFILING(settings->kernel_profile, __func__);
}
