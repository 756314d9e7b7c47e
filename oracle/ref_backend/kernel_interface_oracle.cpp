/*
 * ORACLE backend behind the reference's kernel_interface.h (test infrastructure only).
 *
 * Implements the 22 run_* entry points of TeaLeaf/kernel_interface.h:13-71 (C++ linkage,
 * exactly as the reference's own TeaLeaf/c_kernels/sycl/kernel_interface.cpp:34-371 does)
 * by calling the C restatement in oracle/tealeaf_oracle.c.  Linked with the UNMODIFIED
 * reference host (compiled where it lies under /root/reference) it gives
 * oracle/_ref/tealeaf_ref: the reference's own main()/diffuse()/drivers on the CPU.
 */
#include <stdlib.h>
#include "kernel_interface.h"
#include "shared.h"
#include "../tealeaf_oracle.h"

static double* zeros(long n) { return (double*)calloc((size_t)n, sizeof(double)); }

// kernel_initialise.cpp:31-80 (the reference leaves device buffers uninitialised; we zero)
void run_kernel_initialise(Chunk* chunk, Settings* settings)
{
    const long x = chunk->x, y = chunk->y;
    print_and_log(settings, "Performing this solve with the oracle (CPU restatement) %s solver\n",
                  settings->solver_name);
    chunk->density0 = zeros(x*y); chunk->density = zeros(x*y); chunk->energy0 = zeros(x*y);
    chunk->energy = zeros(x*y); chunk->u = zeros(x*y); chunk->u0 = zeros(x*y);
    chunk->p = zeros(x*y); chunk->r = zeros(x*y); chunk->mi = zeros(x*y); chunk->w = zeros(x*y);
    chunk->kx = zeros(x*y); chunk->ky = zeros(x*y); chunk->sd = zeros(x*y);
    chunk->volume = zeros(x*y); chunk->x_area = zeros((x+1)*y); chunk->y_area = zeros(x*(y+1));
    chunk->cell_x = zeros(x); chunk->cell_y = zeros(y); chunk->cell_dx = zeros(x);
    chunk->cell_dy = zeros(y); chunk->vertex_dx = zeros(x+1); chunk->vertex_dy = zeros(y+1);
    chunk->vertex_x = zeros(x+1); chunk->vertex_y = zeros(y+1);
    chunk->cg_alphas = zeros(settings->max_iters); chunk->cg_betas = zeros(settings->max_iters);
    chunk->cheby_alphas = zeros(settings->max_iters); chunk->cheby_betas = zeros(settings->max_iters);
}

void run_kernel_finalise(Chunk* chunk, Settings* settings)
{
    (void)settings;
    double* all[] = { chunk->density0, chunk->density, chunk->energy0, chunk->energy, chunk->u,
        chunk->u0, chunk->p, chunk->r, chunk->mi, chunk->w, chunk->kx, chunk->ky, chunk->sd,
        chunk->volume, chunk->x_area, chunk->y_area, chunk->cell_x, chunk->cell_y, chunk->cell_dx,
        chunk->cell_dy, chunk->vertex_dx, chunk->vertex_dy, chunk->vertex_x, chunk->vertex_y,
        chunk->cg_alphas, chunk->cg_betas, chunk->cheby_alphas, chunk->cheby_betas };
    for (unsigned i = 0; i < sizeof(all)/sizeof(all[0]); ++i) free(all[i]);
}

void run_set_chunk_data(Chunk* chunk, Settings* settings)
{
    double x_min = settings->grid_x_min + settings->dx*(double)chunk->left;
    double y_min = settings->grid_y_min + settings->dy*(double)chunk->bottom;
    orc_set_chunk_data(chunk->x, chunk->y, settings->halo_depth, x_min, y_min, settings->dx,
        settings->dy, chunk->vertex_x, chunk->vertex_y, chunk->cell_x, chunk->cell_y, chunk->volume);
    for (long i = 0; i < (long)chunk->x*chunk->y; ++i) { chunk->x_area[i] = settings->dy; chunk->y_area[i] = settings->dx; }
}

void run_set_chunk_state(Chunk* chunk, Settings* settings, State* states)
{
    orc_set_chunk_initial_state(chunk->x, chunk->y, states[0].energy, states[0].density,
        chunk->energy0, chunk->density);
    for (int ii = 1; ii < settings->num_states; ++ii)
        orc_set_chunk_state(chunk->x, chunk->y, settings->halo_depth, (int)states[ii].geometry,
            states[ii].density, states[ii].energy, states[ii].x_min, states[ii].y_min,
            states[ii].x_max, states[ii].y_max, states[ii].radius, chunk->energy0, chunk->density,
            chunk->u, chunk->cell_x, chunk->cell_y, chunk->vertex_x, chunk->vertex_y);
}

void run_local_halos(Chunk* chunk, Settings* settings, int depth)
{
    FieldBufferType f[6]; int idx[6] = { FIELD_DENSITY, FIELD_P, FIELD_ENERGY0, FIELD_ENERGY1, FIELD_U, FIELD_SD };
    f[0] = chunk->density; f[1] = chunk->p; f[2] = chunk->energy0; f[3] = chunk->energy; f[4] = chunk->u; f[5] = chunk->sd;
    const int faces[4] = { CHUNK_LEFT, CHUNK_RIGHT, CHUNK_TOP, CHUNK_BOTTOM };
    for (int i = 0; i < 6; ++i) {
        if (!settings->fields_to_exchange[idx[i]]) continue;
        for (int q = 0; q < 4; ++q)
            if (chunk->neighbours[faces[q]] == EXTERNAL_FACE)
                orc_local_halo(chunk->x, chunk->y, settings->halo_depth, depth, faces[q], f[i]);
    }
}

void run_pack_or_unpack(Chunk* chunk, Settings* settings, int depth, int face, bool pack,
                        FieldBufferType field, double* buffer)
{
    if (pack) orc_pack(chunk->x, chunk->y, settings->halo_depth, depth, face, field, buffer);
    else orc_unpack(chunk->x, chunk->y, settings->halo_depth, depth, face, field, buffer);
}

void run_store_energy(Chunk* chunk, Settings* settings)
{
    (void)settings;
    orc_store_energy(chunk->x, chunk->y, chunk->energy0, chunk->energy);
}

void run_field_summary(Chunk* chunk, Settings* settings, double* vol, double* mass, double* ie, double* temp)
{
    orc_field_summary(chunk->x, chunk->y, settings->halo_depth, chunk->volume, chunk->density,
        chunk->energy0, chunk->u, vol, mass, ie, temp);
}

void run_cg_init(Chunk* chunk, Settings* settings, double rx, double ry, double* rro)
{
    orc_cg_init(chunk->x, chunk->y, settings->halo_depth, settings->coefficient, rx, ry,
        chunk->density, chunk->energy, chunk->u, chunk->p, chunk->r, chunk->w, chunk->kx, chunk->ky, rro);
}
void run_cg_calc_w(Chunk* chunk, Settings* settings, double* pw)
{
    orc_cg_calc_w(chunk->x, chunk->y, settings->halo_depth, chunk->p, chunk->kx, chunk->ky, chunk->w, pw);
}
void run_cg_calc_ur(Chunk* chunk, Settings* settings, double alpha, double* rrn)
{
    orc_cg_calc_ur(chunk->x, chunk->y, settings->halo_depth, alpha, chunk->p, chunk->w, chunk->u, chunk->r, rrn);
}
void run_cg_calc_p(Chunk* chunk, Settings* settings, double beta)
{
    orc_cg_calc_p(chunk->x, chunk->y, settings->halo_depth, beta, chunk->r, chunk->p);
}
void run_cheby_init(Chunk* chunk, Settings* settings)
{
    orc_cheby_init(chunk->x, chunk->y, settings->halo_depth, chunk->theta, chunk->u, chunk->u0,
        chunk->kx, chunk->ky, chunk->p, chunk->r, chunk->w);
}
void run_cheby_iterate(Chunk* chunk, Settings* settings, double alpha, double beta)
{
    orc_cheby_iterate(chunk->x, chunk->y, settings->halo_depth, alpha, beta, chunk->u, chunk->u0,
        chunk->kx, chunk->ky, chunk->p, chunk->r, chunk->w);
}
void run_jacobi_init(Chunk* chunk, Settings* settings, double rx, double ry)
{
    orc_jacobi_init(chunk->x, chunk->y, settings->halo_depth, settings->coefficient, rx, ry,
        chunk->density, chunk->energy, chunk->u0, chunk->u, chunk->kx, chunk->ky);
}
void run_jacobi_iterate(Chunk* chunk, Settings* settings, double* error)
{
    orc_jacobi_iterate(chunk->x, chunk->y, settings->halo_depth, chunk->u, chunk->u0, chunk->r,
        chunk->kx, chunk->ky, error);
}
void run_ppcg_init(Chunk* chunk, Settings* settings)
{
    orc_ppcg_init(chunk->x, chunk->y, settings->halo_depth, chunk->theta, chunk->r, chunk->sd);
}
void run_ppcg_inner_iteration(Chunk* chunk, Settings* settings, double alpha, double beta)
{
    orc_ppcg_inner_iteration(chunk->x, chunk->y, settings->halo_depth, alpha, beta, chunk->u,
        chunk->r, chunk->kx, chunk->ky, chunk->sd);
}
void run_copy_u(Chunk* chunk, Settings* settings)
{
    orc_copy_u(chunk->x, chunk->y, settings->halo_depth, chunk->u, chunk->u0);
}
void run_calculate_residual(Chunk* chunk, Settings* settings)
{
    orc_calculate_residual(chunk->x, chunk->y, settings->halo_depth, chunk->u, chunk->u0,
        chunk->kx, chunk->ky, chunk->r);
}
void run_calculate_2norm(Chunk* chunk, Settings* settings, FieldBufferType buffer, double* norm)
{
    orc_calculate_2norm(chunk->x, chunk->y, settings->halo_depth, buffer, norm);
}
void run_finalise(Chunk* chunk, Settings* settings)
{
    orc_finalise(chunk->x, chunk->y, settings->halo_depth, chunk->u, chunk->density, chunk->energy);
}
