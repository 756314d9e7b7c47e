/*
 * comms_b200.cpp -- replaces TeaLeaf/comms.c (MPI) for one process per GPU on one NVSwitch node.
 * Same nine functions as TeaLeaf/comms.h:10-20 (MPI-variant signatures), implemented over
 * tl_comms_* of libtealeaf_b200.so.  Ranks are started by any launcher that exports
 * RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT (e.g. `python -m torch.distributed.run --no-python
 * ./tealeaf`); a single process needs no launcher.
 */
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <vector>
#include "comms.h"
#include "chunk.h"
#include "tealeaf_b200.h"

static tl_comms* g_comms = NULL;
static int g_rank = 0, g_size = 1, g_local = 0;

struct PendingMsg { double* send; double* recv; int len, neighbour, send_tag, recv_tag; };
static std::vector<PendingMsg> g_pending;

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

tl_comms* comms_b200_handle() { return g_comms; }

// comms.c:7-10
void initialise_comms(int argc, char** argv)
{
    (void)argc; (void)argv;
    g_rank = env_int("RANK", 0);
    g_size = env_int("WORLD_SIZE", 1);
    g_local = env_int("LOCAL_RANK", g_rank);
    if (g_size > 1) {
        const char* port = getenv("MASTER_PORT");
        char session[96];
        // per-launch name: every rank of one launch has the launcher as parent process
        snprintf(session, sizeof(session), "tealeaf_%s_%d", port ? port : "0", (int)getppid());
        if (tl_comms_create(&g_comms, session, g_rank, g_size, g_local, 0) != TL_OK) {
            fprintf(stderr, "initialise_comms: %s\n", tl_last_error());
            exit(1);
        }
    }
}

// comms.c:13-22
void initialise_ranks(Settings* settings)
{
    settings->rank = g_rank;
    settings->num_ranks = g_size;
    if (settings->rank == MASTER) printf("Successfully initialised %d B200 ranks.\n", settings->num_ranks);
}

void comms_b200_attach(Chunk* chunk)
{
    if (g_comms && tl_comms_attach_chunk(g_comms, chunk->ext->handle) != TL_OK)
        die(__LINE__, __FILE__, "%s\n", tl_last_error());
}

// comms.c:25-28
void finalise_comms()
{
    if (g_comms) tl_comms_destroy(g_comms);
    g_comms = NULL;
}

// comms.c:31-43: post only; completion happens in wait_for_requests like MPI_Isend/Irecv + Waitall
void send_recv_message(Settings* settings, double* send_buffer, double* recv_buffer, int buffer_len,
                       int neighbour, int send_tag, int recv_tag, MPI_Request* send_request,
                       MPI_Request* recv_request)
{
    START_PROFILING(settings->kernel_profile);
    (void)send_request; (void)recv_request;
    PendingMsg m = { send_buffer, recv_buffer, buffer_len, neighbour, send_tag, recv_tag };
    g_pending.push_back(m);
    STOP_PROFILING(settings->kernel_profile, __func__);
}

// comms.c:46-52
void wait_for_requests(Settings* settings, int num_requests, MPI_Request* requests)
{
    START_PROFILING(settings->kernel_profile);
    (void)num_requests; (void)requests;
    for (size_t ii = 0; ii < g_pending.size(); ++ii) { // all sends first, then all receives
        const PendingMsg& m = g_pending[ii];
        if (tl_comms_post(g_comms, m.send, m.len, m.neighbour, m.send_tag) != TL_OK)
            die(__LINE__, __FILE__, "%s\n", tl_last_error());
    }
    for (size_t ii = 0; ii < g_pending.size(); ++ii) {
        const PendingMsg& m = g_pending[ii];
        if (tl_comms_recv(g_comms, m.recv, m.len, m.neighbour, m.recv_tag) != TL_OK)
            die(__LINE__, __FILE__, "%s\n", tl_last_error());
    }
    g_pending.clear();
    STOP_PROFILING(settings->kernel_profile, __func__);
}

// comms.c:55-61
void sum_over_ranks(Settings* settings, double* a)
{
    START_PROFILING(settings->kernel_profile);
    if (g_comms && tl_comms_sum(g_comms, a) != TL_OK) die(__LINE__, __FILE__, "%s\n", tl_last_error());
    STOP_PROFILING(settings->kernel_profile, __func__);
}

// comms.c:64-70
void min_over_ranks(Settings* settings, double* a)
{
    START_PROFILING(settings->kernel_profile);
    if (g_comms && tl_comms_min(g_comms, a) != TL_OK) die(__LINE__, __FILE__, "%s\n", tl_last_error());
    STOP_PROFILING(settings->kernel_profile, __func__);
}

// comms.c:73-76
void barrier()
{
    if (g_comms) tl_comms_barrier(g_comms);
}

// comms.c:79-82
void abort_comms()
{
    if (g_comms) tl_comms_abort(g_comms); // the other ranks' waits fail at once instead of timing out
    exit(1);
}
