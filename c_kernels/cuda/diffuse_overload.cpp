/*
 * diffuse_overload.cpp -- the reference's sanctioned whole-solve hook: with -DDIFFUSE_OVERLOAD
 * main() calls diffuse_overload() instead of diffuse() (TeaLeaf/main.c:32-36, application.h:9-11).
 * Same timestep loop and the same printed lines as TeaLeaf/diffuse.c:10-78, but each timestep runs
 * the device-resident solver loop of libtealeaf_b200.so (alpha/beta never leave the GPU).
 */
#include "comms.h"
#include "application.h"
#include "drivers/drivers.h"
#include "tealeaf_b200.h"

tl_comms* comms_b200_handle();

void diffuse_overload(Chunk* chunks, Settings* settings)
{
    tl_solve_opts o;
    tl_solve_opts_default(&o);
    o.solver = (int)settings->solver; o.coefficient = settings->coefficient;
    o.max_iters = settings->max_iters; o.eps = settings->eps; o.presteps = settings->presteps;
    o.ppcg_inner_steps = settings->ppcg_inner_steps; o.error_switch = settings->error_switch;
    o.eps_lim = settings->eps_lim; o.check_result = settings->check_result;
    double wallclock_prev = 0.0;
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
        if (tt % settings->summary_frequency == 0) field_summary_driver(chunks, settings, false);
        profiler_end_timer(settings->wallclock_profile, "Wallclock");
        double wallclock = settings->wallclock_profile->profiler_entries[0].time;
        print_and_log(settings, "Wallclock: \t\t%.3lfs\n", wallclock);
        print_and_log(settings, "Avg. time per cell: \t%.6e\n",
                      (wallclock-wallclock_prev) / (settings->grid_x_cells * settings->grid_y_cells));
        print_and_log(settings, "Error: \t\t\t%.6e\n", info.error);
    }
    field_summary_driver(chunks, settings, true);
}
