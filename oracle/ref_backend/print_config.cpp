/*
 * print_config.cpp -- test harness (test infrastructure only): runs the REFERENCE's own deck parser
 * (TeaLeaf/parse_config.c:14-74 read_config, compiled in place from /root/reference) on ./tea.in and prints
 * the resulting Settings / State values, so tests/test_host_logic.py can pin the Python mirror
 * exploringsycl_b200.read_config against it.
 */
#include <stdio.h>
#include <stdlib.h>
#include "settings.h"
#include "application.h"

int main()
{
    Settings* settings = (Settings*)malloc(sizeof(Settings));
    set_default_settings(settings);
    settings->rank = 1; // not MASTER: keeps the parser from opening tea.out
    settings->solver_name[0] = 0;
    State* states = NULL;
    read_config(settings, &states);
    printf("grid_x_cells %d\ngrid_y_cells %d\nend_step %d\nmax_iters %d\npresteps %d\nppcg_inner_steps %d\n",
           settings->grid_x_cells, settings->grid_y_cells, settings->end_step, settings->max_iters,
           settings->presteps, settings->ppcg_inner_steps);
    printf("summary_frequency %d\nhalo_depth %d\nnum_states %d\nsolver %d\ncoefficient %d\nerror_switch %d\n"
           "check_result %d\n", settings->summary_frequency, settings->halo_depth, settings->num_states,
           (int)settings->solver, settings->coefficient, (int)settings->error_switch, (int)settings->check_result);
    printf("dt_init %.17g\neps %.17g\neps_lim %.17g\nend_time %.17g\ngrid_x_min %.17g\ngrid_y_min %.17g\n"
           "grid_x_max %.17g\ngrid_y_max %.17g\ndx %.17g\ndy %.17g\n", settings->dt_init, settings->eps,
           settings->eps_lim, settings->end_time, settings->grid_x_min, settings->grid_y_min,
           settings->grid_x_max, settings->grid_y_max, settings->dx, settings->dy);
    for (int ss = 0; ss < settings->num_states; ++ss) {
        if (ss == 0)
            printf("state %d %.17g %.17g\n", ss, states[ss].density, states[ss].energy);
        else
            printf("state %d %.17g %.17g %d %.17g %.17g %.17g %.17g\n", ss, states[ss].density, states[ss].energy,
                   (int)states[ss].geometry, states[ss].x_min, states[ss].y_min, states[ss].x_max, states[ss].y_max);
    }
    return 0;
}
