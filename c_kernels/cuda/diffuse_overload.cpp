/*
 * diffuse_overload.cpp -- the reference's sanctioned whole-solve hook: with -DDIFFUSE_OVERLOAD
 * main() calls diffuse_overload() instead of diffuse() (TeaLeaf/main.c:32-36, application.h:9-11).
 * Same timestep loop and the same printed lines as TeaLeaf/diffuse.c:10-78, but each timestep runs
 * the device-resident solver loop of libtealeaf_b200.so (alpha/beta never leave the GPU).
 *
 * Reporting.  The reference computes volume, mass, internal energy and temperature in
 * field_summary_driver (drivers/field_summary_driver.c:11-30) and prints only the temperature, and only
 * at the end (:32-52).  Here, in ADDITION to the reference's lines (which stay exactly as they are):
 *   - at every summary step:  "Field summary: \tvol ... mass ... ie ... temp ..."
 *   - at every timestep:      "Solver rate: \t<iterations> iterations, <cell-iterations/s>, <GB/s algorithmic>"
 *   - at the end, on the master rank: a JSON sidecar (default tea.json; TL_REPORT_JSON=<path>) with the deck,
 *     the per-step iteration counts / error / solver time / rates and the four final sums.
 * TL_REPORT=0 switches all of it off (output is then byte-identical in form to the reference's).
 * Algorithmic bytes per cell and iteration: CG 104, Chebyshev 88, PPCG 128 per outer + 80 per
 * inner step, Jacobi 56.
 */
#include <stdlib.h>
#include <string>
#include <vector>
#include "comms.h"
#include "application.h"
#include "drivers/drivers.h"
#include "tealeaf_b200.h"

tl_comms* comms_b200_handle();

namespace {
struct StepRecord {
    int step, iters_a, iters_b, est_iters, total_iters;
    double error, gpu_ms, cell_iters, gbytes;
};

// cell-iterations and algorithmic bytes of one solve
void solve_work(const Settings* s, const tl_solve_info& info, double* cell_iters, double* bytes)
{
    const double cells = (double)s->grid_x_cells * (double)s->grid_y_cells;
    double iters = info.total_iters, bpc;
    switch (s->solver) {
    case JACOBI_SOLVER: bpc = 56.0 * info.total_iters; break;
    case CHEBY_SOLVER: bpc = 104.0 * (info.total_iters - info.iters_b) + 88.0 * info.iters_b; break;
    case PPCG_SOLVER:
        iters += (double)info.iters_b * s->ppcg_inner_steps;
        bpc = 104.0 * (info.total_iters - info.iters_b) + (128.0 + 80.0 * s->ppcg_inner_steps) * info.iters_b;
        break;
    default: bpc = 104.0 * info.total_iters;
    }
    *cell_iters = cells * iters;
    *bytes = cells * bpc;
}

// field_summary_driver.c:11-30: run_field_summary on every chunk of this rank (accumulated), then sum_over_ranks x 4
void four_sums(Chunk* chunks, Settings* s, double sums[4])
{
    sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
    for (int cc = 0; cc < s->num_chunks_per_rank; ++cc) {
        double v[4];
        if (tl_run_field_summary(chunks[cc].ext->handle, &v[0], &v[1], &v[2], &v[3]) != TL_OK)
            die(__LINE__, __FILE__, "%s\n", tl_last_error());
        for (int q = 0; q < 4; ++q) sums[q] += v[q];
    }
    for (int q = 0; q < 4; ++q) sum_over_ranks(s, &sums[q]);
}

void write_sidecar(const Settings* s, const std::vector<StepRecord>& steps, const double sums[4], double wallclock)
{
    const char* path = getenv("TL_REPORT_JSON");
    FILE* f = fopen(path && *path ? path : "tea.json", "w");
    if (!f) return;
    static const char* solver_names[] = {"jacobi", "cg", "chebyshev", "ppcg"};
    double ci = 0.0, gb = 0.0, ms = 0.0;
    long iters = 0;
    for (const StepRecord& r : steps) {
        ci += r.cell_iters;
        gb += r.gbytes;
        ms += r.gpu_ms;
        iters += r.total_iters;
    }
    fprintf(f, "{\n  \"backend\": \"%s\",\n  \"solver\": \"%s\",\n  \"grid\": [%d, %d],\n  \"ranks\": %d,\n"
               "  \"end_step\": %d,\n  \"eps\": %.17g,\n  \"max_iters\": %d,\n",
            tl_version(), solver_names[(int)s->solver], s->grid_x_cells, s->grid_y_cells, s->num_ranks, s->end_step, s->eps,
            s->max_iters);
    fprintf(f, "  \"steps\": [\n");
    for (size_t n = 0; n < steps.size(); ++n) {
        const StepRecord& r = steps[n];
        fprintf(f, "    {\"step\": %d, \"iters_a\": %d, \"iters_b\": %d, \"est_iters\": %d, \"total_iters\": %d, "
                   "\"error\": %.17g, \"solver_ms\": %.6f, \"cell_iters_per_s\": %.6e, \"algorithmic_gb_per_s\": %.3f}%s\n",
                r.step, r.iters_a, r.iters_b, r.est_iters, r.total_iters, r.error, r.gpu_ms,
                r.gpu_ms > 0 ? r.cell_iters / (r.gpu_ms * 1e-3) : 0.0, r.gpu_ms > 0 ? r.gbytes / (r.gpu_ms * 1e-3) : 0.0,
                n + 1 < steps.size() ? "," : "");
    }
    fprintf(f, "  ],\n  \"total_iters\": %ld,\n  \"solver_ms\": %.6f,\n  \"cell_iters_per_s\": %.6e,\n"
               "  \"algorithmic_gb_per_s\": %.3f,\n  \"wallclock_s\": %.6f,\n",
            iters, ms, ms > 0 ? ci / (ms * 1e-3) : 0.0, ms > 0 ? gb / (ms * 1e-3) : 0.0, wallclock);
    fprintf(f, "  \"field_summary\": {\"volume\": %.17g, \"mass\": %.17g, \"internal_energy\": %.17g, "
               "\"temperature\": %.17g}\n}\n", sums[0], sums[1], sums[2], sums[3]);
    fclose(f);
}
} // namespace

void diffuse_overload(Chunk* chunks, Settings* settings)
{
    tl_solve_opts o;
    tl_solve_opts_default(&o);
    o.solver = (int)settings->solver; o.coefficient = settings->coefficient;
    o.max_iters = settings->max_iters; o.eps = settings->eps; o.presteps = settings->presteps;
    o.ppcg_inner_steps = settings->ppcg_inner_steps; o.error_switch = settings->error_switch;
    o.eps_lim = settings->eps_lim; o.check_result = settings->check_result;
    const char* rep = getenv("TL_REPORT");
    const bool report = !(rep && rep[0] == '0');
    std::vector<StepRecord> steps;
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    double wallclock_prev = 0.0, wallclock = 0.0;
    for (int tt = 0; tt < settings->end_step; ++tt) {
        print_and_log(settings, "\nTimestep %d\n", tt+1);
        profiler_start_timer(settings->wallclock_profile);
        tl_solve_info info;
        for (int cc = 0; cc < settings->num_chunks_per_rank; ++cc) {
            if (tl_timestep(chunks[cc].ext->handle, comms_b200_handle(), &o, chunks[cc].dt_init,
                            settings->dx, settings->dy, &info) != TL_OK)
                die(__LINE__, __FILE__, "%s\n", tl_last_error());
        }
        switch (settings->solver) { // cg_driver.c:27, cheby_driver.c:73-76, ppcg_driver.c:59-62, jacobi_driver.c:24
        case JACOBI_SOLVER: print_and_log(settings, "Jacobi: \t\t%d iterations\n", info.iters_a); break;
        case CG_SOLVER: print_and_log(settings, "CG: \t\t\t%d iterations\n", info.iters_a); break;
        case CHEBY_SOLVER:
            print_and_log(settings, "CG: \t\t\t%d iterations\n", info.iters_a);
            print_and_log(settings, "Cheby: \t\t\t%d iterations (%d estimated)\n", info.iters_b, info.est_iters);
            break;
        case PPCG_SOLVER:
            print_and_log(settings, "CG: \t\t\t%d iterations\n", info.iters_a);
            print_and_log(settings, "PPCG: \t\t\t%d iterations (%d inner iterations per)\n", info.iters_b,
                          settings->ppcg_inner_steps);
            break;
        }
        if (tt % settings->summary_frequency == 0) {
            field_summary_driver(chunks, settings, false);
            if (report) { // the four sums the reference computes and drops (field_summary_driver.c:11-30)
                four_sums(chunks, settings, sums);
                print_and_log(settings, "Field summary: \t\tvol %.15e mass %.15e ie %.15e temp %.15e\n", sums[0], sums[1],
                              sums[2], sums[3]);
            }
        }
        profiler_end_timer(settings->wallclock_profile, "Wallclock");
        wallclock = settings->wallclock_profile->profiler_entries[0].time;
        print_and_log(settings, "Wallclock: \t\t%.3lfs\n", wallclock);
        print_and_log(settings, "Avg. time per cell: \t%.6e\n",
                      (wallclock-wallclock_prev) / (settings->grid_x_cells * settings->grid_y_cells));
        print_and_log(settings, "Error: \t\t\t%.6e\n", info.error);
        if (report) {
            StepRecord r = {tt + 1, info.iters_a, info.iters_b, info.est_iters, info.total_iters, info.error, info.gpu_ms, 0.0, 0.0};
            double bytes = 0.0;
            solve_work(settings, info, &r.cell_iters, &bytes);
            r.gbytes = bytes / 1e9;
            steps.push_back(r);
            print_and_log(settings, "Solver rate: \t\t%d iterations in %.3f ms, %.4e cell-iterations/s, %.1f GB/s (algorithmic)\n",
                          info.total_iters, info.gpu_ms, info.gpu_ms > 0 ? r.cell_iters / (info.gpu_ms * 1e-3) : 0.0,
                          info.gpu_ms > 0 ? r.gbytes / (info.gpu_ms * 1e-3) : 0.0);
        }
    }
    field_summary_driver(chunks, settings, true);
    if (report) {
        four_sums(chunks, settings, sums);
        print_and_log(settings, "Field summary: \t\tvol %.15e mass %.15e ie %.15e temp %.15e\n", sums[0], sums[1], sums[2],
                      sums[3]);
        if (settings->rank == MASTER) write_sidecar(settings, steps, sums, wallclock);
    }
}
