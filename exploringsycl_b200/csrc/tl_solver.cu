// tl_solver.cu -- solver drivers: the call order of the reference's drivers/*.c re-expressed for a
// GPU-resident state.
//
//   cg_driver      (drivers/cg_driver.c:7-124)      -> cg_solve_resident(): alpha = rro/pw and
//                  beta = rrn/rro are computed by the tail CTA of the reduction kernels and read
//                  by the next kernel from HBM, so a CG iteration is 3 stream-ordered launches
//                  with no host round trip; the host polls a convergence flag once per batch.
//   cheby_driver   (drivers/cheby_driver.c:11-183)   -> cheby_solve()
//   ppcg_driver    (drivers/ppcg_driver.c:10-189)    -> ppcg_solve()
//   jacobi_driver  (drivers/jacobi_driver.c:7-84)    -> jacobi_solve()
//   eigenvalue_driver (drivers/eigenvalue_driver.c)  -> eigenvalues() / tqli()  (host, fp64)
//
// Iteration counts, switch rules and convergence tests are exactly the reference's.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "tl_internal.h"

int tlc_halo_exchange(tl_chunk* c, tl_comms* k, const int fields[6], int depth);
unsigned long long tlc_resident_seq_advance(tl_comms* k, int launched);

__global__ void k_set_rro(DevScal* S, double rro) { S->rro = rro; }

static void fields_reset(int* f) { memset(f, 0, sizeof(int) * TL_NUM_EXCHANGE_FIELDS); }

// drivers/halo_update_driver.c:6-25
extern "C" int tl_halo_update(tl_chunk* c, tl_comms* k, const int fields_to_exchange[6], int depth)
{
    TL_CHECK_ARG(c && fields_to_exchange, "null argument");
    TL_CHECK_ARG(depth >= 1 && depth <= c->g.hd, "depth must be between 1 and halo_depth");
    TL_CUDA(cudaSetDevice(c->device));
    bool any = false;
    for (int i = 0; i < TL_NUM_EXCHANGE_FIELDS; ++i) any |= (fields_to_exchange[i] != 0);
    if (!any) return TL_OK;
    TL_TRY(tlc_halo_exchange(c, k, fields_to_exchange, depth));
    TL_TRY(tlk_local_halos(c, fields_to_exchange, depth));
    return tl_check_peer_timeout(c); // no sync here: reports a time-out of any EARLIER wait of this chunk
}

extern "C" void tl_solve_opts_default(tl_solve_opts* o)
{
    memset(o, 0, sizeof(*o));
    o->solver = TL_SOLVER_CG;           // settings.h:37
    o->coefficient = TL_CONDUCTIVITY;   // settings.h:31
    o->max_iters = 10000;               // settings.h:26
    o->eps = 1.0e-15;                   // settings.h:27
    o->presteps = 30;                   // settings.h:33
    o->ppcg_inner_steps = 10;           // settings.h:36
    o->error_switch = 0;                // settings.h:32
    o->eps_lim = 1e-5;                  // settings.h:34
    o->check_result = 1;                // settings.h:35
    o->fuse_p_into_w = 1;               // bit-identical to the three-kernel iteration, 96 instead of 104 B/cell
    o->batch = 0;
}

static int sum_ranks(tl_comms* k, double* v) { return (k && tl_comms_size(k) > 1) ? tl_comms_sum(k, v) : TL_OK; }

// ---------------------------------------------------------------------------------------------
// cg_init_driver, cg_driver.c:31-66
// ---------------------------------------------------------------------------------------------
static int cg_init_driver(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double rx, double ry, int* fields,
                          double* rro, bool multi)
{
    TL_TRY(tlk_cg_init(c, o->coefficient, rx, ry));
    fields_reset(fields);
    fields[TL_FIELD_U] = 1;
    fields[TL_FIELD_P] = 1;
    TL_TRY(tl_halo_update(c, k, fields, 1));
    if (multi) {
        TL_TRY(tl_fetch_scal(c));
        *rro = c->scal_h->sums[0];
        TL_TRY(sum_ranks(k, rro));
        k_set_rro<<<1, 1, 0, c->stream>>>(c->scal, *rro); // the resident multi-rank loop starts from it
        ++g_tl_launches;
    } else {
        TL_TRY(tlk_seed_rro(c)); // rro stays in HBM
    }
    return tlk_copy_field(c, TL_FIELD_U0, TL_FIELD_U, true);
}

// cg_main_step_driver, cg_driver.c:69-124, host-driven (used when the scalars must cross ranks
// through the host, and by the Chebyshev / PPCG pre-steps of a multi-rank run).
static int cg_main_step_host(tl_chunk* c, tl_comms* k, int tt, double* rro, double* error, long* launches)
{
    double pw = 0.0;
    TL_TRY(tlk_cg_calc_w(c, SCAL_IMM, false));
    TL_TRY(tl_fetch_scal(c));
    pw += c->scal_h->pw;
    TL_TRY(sum_ranks(k, &pw));
    const double alpha = *rro / pw;
    c->cg_alphas[tt] = alpha;
    TL_TRY(tlk_cg_calc_ur(c, SCAL_IMM, alpha, false));
    TL_TRY(tl_fetch_scal(c));
    double rrn = c->scal_h->rrn;
    TL_TRY(sum_ranks(k, &rrn));
    const double beta = rrn / *rro;
    c->cg_betas[tt] = beta;
    TL_TRY(tlk_cg_calc_p(c, SCAL_IMM, beta, false, false));
    *error = rrn;
    *rro = rrn;
    *launches += 3;
    return TL_OK;
}

// Resident CG iterations: runs until `stop_iters` iterations are done or the convergence test fires.
// abs_test == 0: sqrt(|rrn|) < eps (cg_driver.c:24); abs_test == 1: |rrn| < eps (cheby_driver.c:70,
// ppcg_driver.c:56).  On return the DevScal mirror is current.
//
// One rank (k == null): three (or, fused, two) stream-ordered launches per iteration, alpha and beta never leave HBM.
// N ranks: the SAME kernels and heads; p.w and r.r are combined across ranks inside the tail CTA of the reduction
// kernels (every rank's partial is stored into every rank's slot array over NVLink, rank-ordered sum), and calc_p
// (three kernels: p) or calc_ur (fused: r) stores its edge cells into the neighbours' halo and hand-shakes per face in
// its tail.  No host round trip and no separate halo / all-reduce launches inside the loop; the host only polls the
// convergence flag once per batch.
__global__ void k_set_stop(DevScal* S, int stop_iters, double eps, int abs_test, int sw_min_iters, double sw_thresh)
{
    S->max_iters = stop_iters;
    S->eps = eps;
    S->conv_mode = abs_test;
    S->sw_min_iters = sw_min_iters;
    S->sw_thresh = sw_thresh;
    S->conv = (S->iters >= stop_iters) ? 1 : 0;
}

// Programmatic dependent launch of the resident loops' kernels (tl_device.cuh); TL_PDL=0 disables it (experiments).
static bool use_pdl()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TL_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// sw_min_iters >= 0: additionally stop after the first iteration that leaves iters > sw_min_iters and rrn < sw_thresh
// (the CG pre-steps of the Chebyshev / PPCG drivers end there; decided on the device from the global rrn, so every rank
// stops at the same iteration and the host never has to step the loop one iteration at a time).
static int cg_iterate_resident(tl_chunk* c, tl_comms* k, int stop_iters, double eps, int abs_test, int batch,
                               long* launches, bool fused = false, int sw_min_iters = -1, double sw_thresh = 0.0)
{
    const bool pdl = use_pdl();
    const bool multi = k && tl_comms_size(k) > 1 && c->has_peers;
    k_set_stop<<<1, 1, 0, c->stream>>>(c->scal, stop_iters, eps, abs_test, sw_min_iters, sw_thresh);
    ++g_tl_launches;
    if (batch <= 0) batch = 32;
    // Two pinned status snapshots in flight: batch b+1 is enqueued before batch b's status is read,
    // so the GPU queue never drains; kernels launched after convergence return immediately.
    DevScal* snaps[2] = {c->scal_h + 1, c->scal_h + 2};
    cudaEvent_t ev[2] = {c->ev0, c->ev1};
    const int start = c->resident_iters; // iterations already done in this solve
    if (fused && start != 0) {
        tl_set_error("fused CG iterations must start at iteration 0 of a solve");
        return TL_ERR_ARG;
    }
    int enq = start;
    MultiCtx mc = c->mc;
    const MultiCtx* mcp = nullptr;
    if (multi) {
        mc.sbase = mc.hbase = tlc_resident_seq_advance(k, 0);
        mcp = &mc;
    }
    int nb = 0, n_pw = 0;
    bool done = (enq >= stop_iters);
    const bool send_r = multi; // fused loop on several ranks: calc_ur delivers r's halo to the neighbours
    while (!done) {
        const int todo = (stop_iters - enq) < batch ? (stop_iters - enq) : batch;
        for (int it = 0; it < todo; ++it) {
            mc.tl = enq + it - start;
            if (!fused) {
                // cg_main_step_driver (cg_driver.c:69-124) + the p part of halo_update_driver (:22)
                TL_TRY(tlk_cg_calc_w(c, SCAL_DEV, false, mcp, pdl));
                TL_TRY(tlk_cg_calc_ur(c, SCAL_DEV, 0.0, true, mcp, false, pdl));
                TL_TRY(tlk_cg_calc_p(c, SCAL_DEV, 0.0, false, true, mcp, pdl));
                *launches += 3;
            } else {
                // iteration t >= 1 applies p = beta_{t-1} p + r inside the matvec kernel; across ranks r's halo (not
                // p's) travels and the ring of updated p is recomputed locally
                if (enq + it == 0) {
                    TL_TRY(tlk_cg_calc_w(c, SCAL_DEV, false, mcp, pdl));
                } else {
                    TL_TRY(tlk_cg_calc_pw(c, false, mcp, pdl));
                    ++n_pw;
                }
                TL_TRY(tlk_cg_calc_ur(c, SCAL_DEV, 0.0, true, mcp, send_r, pdl));
                *launches += 2;
            }
        }
        enq += todo;
        TL_CUDA(cudaMemcpyAsync(snaps[nb & 1], c->scal, sizeof(DevScal), cudaMemcpyDeviceToHost, c->stream));
        TL_CUDA(cudaEventRecord(ev[nb & 1], c->stream));
        if (nb > 0) {
            TL_CUDA(cudaEventSynchronize(ev[(nb - 1) & 1]));
            if (snaps[(nb - 1) & 1]->conv) done = true;
        }
        ++nb;
        if (enq >= stop_iters) done = true;
    }
    // every rank launched the same iterations (the convergence decision comes from the same rank-ordered sums, and the
    // poll of batch b happens after batch b + 1 was enqueued on every rank): the sequence bases stay in lockstep
    if (multi) tlc_resident_seq_advance(k, enq - start);
    TL_TRY(tl_fetch_scal(c));
    c->resident_iters = c->scal_h->iters;
    if (fused) {
        // Launches enqueued after convergence returned at once, but the host swapped the P/P2
        // roles for each of them: undo the swaps that did not happen on the device.
        const int executed_pw = c->scal_h->iters > 0 ? c->scal_h->iters - 1 : 0;
        if ((n_pw - executed_pw) & 1) {
            double* tmp = c->f[TL_FIELD_P];
            c->f[TL_FIELD_P] = c->p2;
            c->p2 = tmp;
        }
        // the last iteration's p update is still pending (cg_driver.c:115) -- apply it, with its reflective halo; the
        // caller's halo update (cg_driver.c:22) refreshes the halos shared with neighbouring ranks
        TL_TRY(tlk_cg_calc_p(c, SCAL_DEV, 0.0, false, true));
        *launches += 1;
        // Neighbours address this chunk's p through the slab mapping (three-kernel loop): leave p there.
        if (multi) TL_TRY(tlk_field_home(c, TL_FIELD_P));
    }
    if (c->scal_h->pad == 0xdeadu) {
        tl_set_error("resident CG loop: timed out waiting for a peer rank (site %llu want %llu seen %llu block %llu, "
                     "iters %d)", c->scal_h->dbg[0], c->scal_h->dbg[1], c->scal_h->dbg[2], c->scal_h->dbg[3],
                     c->scal_h->iters);
        return TL_ERR_COMMS;
    }
    return TL_OK;
}

// Which form of the resident CG iteration runs for this chunk and option value (same answer on every rank of a
// decomposition; results are bit-identical either way): fused (calc_pw + calc_ur, 96 B/cell moved) unless
// fuse_p_into_w == 0.  Round 1 fell back to three kernels on decompositions with left/right neighbours; with the
// halo columns parked and forwarded group by group the fused form wins there too (profiles/multi_rank_stamps_r02.txt:
// 2 x 2 ranks 0.262 vs 0.299 ms per iteration, 2 x 4 ranks 0.265 vs 0.301).
extern "C" int tl_cg_loop_is_fused(const tl_chunk* c, int fuse_p_into_w)
{
    return (c && fuse_p_into_w != 0) ? 1 : 0;
}

static bool use_resident_multi(const tl_chunk* c, tl_comms* k)
{
    const char* e = getenv("TL_MULTI_HOST_DRIVEN");
    return k && tl_comms_size(k) > 1 && c->has_peers && !(e && e[0] == '1');
}

static int fetch_cg_coeffs(tl_chunk* c, int n)
{
    if (n <= 0) return TL_OK;
    TL_CUDA(cudaMemcpyAsync(c->cg_alphas, c->d_alphas, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaMemcpyAsync(c->cg_betas, c->d_betas, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return TL_OK;
}

// cg_driver, cg_driver.c:7-28
static int cg_solve(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double rx, double ry, tl_solve_info* info)
{
    const bool multi = k && tl_comms_size(k) > 1;
    int fields[TL_NUM_EXCHANGE_FIELDS];
    double rro = 0.0, error = 1e+10;
    long launches = 0;
    TL_TRY(tlk_reset_solve_scalars(c, o->eps, o->max_iters));
    TL_TRY(cg_init_driver(c, k, o, rx, ry, fields, &rro, multi));
    int tt;
    if (!multi) {
        // p's reflective halo is written by calc_p itself; u's halo is only read after the loop
        // (calculate_residual, solve_finished_driver.c:19), so it is refreshed once at the end.
        TL_TRY(cg_iterate_resident(c, nullptr, o->max_iters, o->eps, 0, o->batch, &launches, o->fuse_p_into_w != 0));
        const DevScal* S = c->scal_h;
        error = S->error;
        const bool converged = sqrt(fabs(error)) < o->eps;
        tt = converged ? S->iters - 1 : o->max_iters; // loop index printed by cg_driver.c:27
        info->total_iters = S->iters;
        TL_TRY(tl_halo_update(c, k, fields, 1));
        TL_TRY(fetch_cg_coeffs(c, S->iters));
    } else if (use_resident_multi(c, k)) {
        const bool fuse_multi = tl_cg_loop_is_fused(c, o->fuse_p_into_w) != 0;
        TL_TRY(cg_iterate_resident(c, k, o->max_iters, o->eps, 0, o->batch, &launches, fuse_multi));
        const DevScal* S = c->scal_h;
        error = S->error;
        const bool converged = sqrt(fabs(error)) < o->eps;
        tt = converged ? S->iters - 1 : o->max_iters;
        info->total_iters = S->iters;
        TL_TRY(tl_halo_update(c, k, fields, 1)); // u's halo (and p's corners) as after cg_driver.c:22
        TL_TRY(fetch_cg_coeffs(c, S->iters));
    } else {
        for (tt = 0; tt < o->max_iters; ++tt) {
            TL_TRY(cg_main_step_host(c, k, tt, &rro, &error, &launches));
            TL_TRY(tl_halo_update(c, k, fields, 1));
            if (sqrt(fabs(error)) < o->eps) break;
        }
        info->total_iters = (tt < o->max_iters) ? tt + 1 : tt;
    }
    info->iters_a = tt;
    info->error = error;
    info->kernel_launches += launches;
    return TL_OK;
}

// ---------------------------------------------------------------------------------------------
// eigenvalue_driver.c:71-122 (tqli) and :11-67
// ---------------------------------------------------------------------------------------------
static int tqli(double* d, double* e, int n)
{
    int m, l, iter, i;
    double s, r, p, g, f, dd, c, b;
    for (i = 0; i < n - 1; i++) e[i] = e[i + 1];
    e[n - 1] = 0.0;
    for (l = 0; l < n; l++) {
        iter = 0;
        do {
            for (m = l; m < n - 1; m++) {
                dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) + dd == dd) break;
            }
            if (m == l) break;
            if (iter++ == 30) {
                tl_set_error("Too many iterations in TQLI routine");
                return TL_ERR_NUMERIC;
            }
            g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            r = sqrt((g * g) + 1.0);
            const double sg = (g < 0) ? -fabs(r) : fabs(r);
            g = d[m] - d[l] + e[l] / (g + sg);
            s = c = 1.0;
            p = 0.0;
            for (i = m - 1; i >= l; i--) {
                f = s * e[i];
                b = c * e[i];
                r = sqrt(f * f + g * g);
                e[i + 1] = r;
                if (r == 0.0) {
                    d[i + 1] -= p;
                    e[m] = 0.0;
                    continue;
                }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
            }
            d[l] = d[l] - p;
            e[l] = g;
            e[m] = 0.0;
        } while (m != l);
    }
    return TL_OK;
}

static int eigenvalues(const tl_chunk* c, int n, double* eigmin, double* eigmax)
{
    std::vector<double> diag(n + 1, 0.0), off(n + 1, 0.0);
    for (int ii = 0; ii < n; ++ii) {
        diag[ii] = 1.0 / c->cg_alphas[ii];
        if (ii > 0) diag[ii] += c->cg_betas[ii - 1] / c->cg_alphas[ii - 1];
        if (ii < n - 1) off[ii + 1] = sqrt(c->cg_betas[ii]) / c->cg_alphas[ii];
    }
    TL_TRY(tqli(diag.data(), off.data(), n));
    double mn = DBL_MAX, mx = DBL_MIN;
    for (int ii = 0; ii < n; ++ii) {
        mn = diag[ii] < mn ? diag[ii] : mn;
        mx = diag[ii] > mx ? diag[ii] : mx;
    }
    if (mn < 0.0 || mx < 0.0) {
        tl_set_error("Calculated negative eigenvalues.");
        return TL_ERR_NUMERIC;
    }
    *eigmin = mn * 0.95;
    *eigmax = mx * 1.05;
    return TL_OK;
}

// cheby_driver.c:163-183
static void cheby_coef(tl_chunk* c, double eigmin, double eigmax, int n, double* theta)
{
    *theta = (eigmax + eigmin) / 2.0;
    const double delta = (eigmax - eigmin) / 2.0;
    const double sigma = *theta / delta;
    double rho_old = 1.0 / sigma;
    for (int ii = 0; ii < n; ++ii) {
        const double rho_new = 1.0 / (2.0 * sigma - rho_old);
        c->cheby_alphas[ii] = rho_new * rho_old;
        c->cheby_betas[ii] = 2.0 * rho_new / delta;
        rho_old = rho_new;
    }
}

static bool switch_rule(const tl_solve_opts* o, int started, int tt, double error)
{
    // cheby_driver.c:30-32 / ppcg_driver.c:27-29; CG_ITERS_FOR_EIGENVALUES 20, ERROR_SWITCH_MAX 1.0
    return started || (o->error_switch ? (error < o->eps_lim) && (tt > 20) : (tt > o->presteps) && (error < 1.0));
}

// The CG pre-steps shared by cheby_driver and ppcg_driver: iterate CG until the switch rule fires
// (or the |error| < eps test of those drivers ends the solve).  Returns tt = index of the first
// non-CG iteration (or the loop index at which the solve ended) and `ended`.
static int cg_presteps(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, int* fields, double* rro, double* error,
                       int* tt_out, bool* ended, long* launches)
{
    const bool multi = k && tl_comms_size(k) > 1;
    int tt = 0;
    *ended = false;
    const bool resident = !multi || use_resident_multi(c, k);
    if (resident) {
        // One resident call: the |error| < eps end test of cheby_driver.c:70 AND the switch rule (tt > presteps &&
        // error < 1, or with errswitch error < eps_lim && tt > 20: cheby_driver.c:30-32) are evaluated on the device after
        // every iteration, so the loop stops exactly where the reference's loop index would leave CG.
        const bool fused = tl_cg_loop_is_fused(c, o->fuse_p_into_w) != 0;
        TL_TRY(cg_iterate_resident(c, multi ? k : nullptr, o->max_iters, o->eps, 1, o->batch, launches, fused,
                                   o->error_switch ? 20 : o->presteps, o->error_switch ? o->eps_lim : 1.0));
        *error = c->scal_h->error;
        if (fabs(*error) < o->eps) {
            *ended = true;
            tt = c->scal_h->iters - 1;
        } else {
            tt = c->scal_h->iters;
        }
        TL_TRY(fetch_cg_coeffs(c, c->scal_h->iters));
        // u and p halos as after halo_update_driver of the last CG iteration (cheby_driver.c:68)
        TL_TRY(tl_halo_update(c, k, fields, 1));
        *rro = c->scal_h->rro;
    } else {
        for (tt = 0; tt < o->max_iters; ++tt) {
            if (switch_rule(o, 0, tt, *error)) break;
            TL_TRY(cg_main_step_host(c, k, tt, rro, error, launches));
            TL_TRY(tl_halo_update(c, k, fields, 1));
            if (fabs(*error) < o->eps) { *ended = true; break; }
        }
    }
    *tt_out = tt;
    return TL_OK;
}

static int norm2(tl_chunk* c, tl_comms* k, int field, double* out)
{
    TL_TRY(tlk_calculate_2norm(c, field));
    TL_TRY(tl_fetch_scal(c));
    *out = c->scal_h->sums[0];
    return sum_ranks(k, out);
}

// cheby_driver.c:11-160
static int cheby_solve(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double rx, double ry, tl_solve_info* info)
{
    const bool multi = k && tl_comms_size(k) > 1;
    int fields[TL_NUM_EXCHANGE_FIELDS];
    double rro = 0.0, error = 1e+10;
    long launches = 0;
    TL_TRY(tlk_reset_solve_scalars(c, o->eps, o->max_iters));
    TL_TRY(cg_init_driver(c, k, o, rx, ry, fields, &rro, multi));
    int tt = 0;
    bool ended = false;
    TL_TRY(cg_presteps(c, k, o, fields, &rro, &error, &tt, &ended, &launches));
    int num_cheby_iters = 0, est_iterations = 0;
    double theta = 0.0;
    const bool fused = o->fuse_p_into_w != 0;
    // One-pass kernel on several ranks: u's edge cells travel inside the kernel (NVLink peer stores + per-face
    // hand-shake in its last CTA) instead of a halo_update_driver round of four launches per iteration.
    const bool inkernel = fused && multi && use_resident_multi(c, k);
    MultiCtx mc = c->mc;
    const MultiCtx* mcp = nullptr;
    int n_seq = 0;
    if (inkernel) {
        mc.sbase = mc.hbase = tlc_resident_seq_advance(k, 0);
        mcp = &mc;
    }
    if (!ended) {
        for (; tt < o->max_iters; ++tt) {
            num_cheby_iters++;
            bool calc_2norm;
            double bb = 0.0;
            if (num_cheby_iters == 1) {
                // cheby_init_driver, cheby_driver.c:80-107
                TL_TRY(eigenvalues(c, tt, &info->eigmin, &info->eigmax));
                cheby_coef(c, info->eigmin, info->eigmax, o->max_iters - tt, &theta);
                TL_TRY(tlk_calculate_2norm(c, TL_FIELD_U0));
                TL_TRY(tl_fetch_scal(c));
                bb = c->scal_h->sums[0];
                TL_TRY(tlk_cheby_init(c, theta));
                fields_reset(fields);
                fields[TL_FIELD_U] = 1;
                TL_TRY(tl_halo_update(c, k, fields, 1));
                TL_TRY(sum_ranks(k, &bb));
                calc_2norm = true;
            } else {
                calc_2norm = (num_cheby_iters >= est_iterations) && ((tt + 1) % 10 == 0);
            }
            // cheby_main_step_driver, cheby_driver.c:110-143
            if (fused) {
                // one pass: cheby_iterate + cheby_calc_u (u double buffered) + on the sampled iterations the 2-norm of r
                mc.tl = n_seq++;
                TL_TRY(tlk_cheby_fused(c, c->cheby_alphas[num_cheby_iters], c->cheby_betas[num_cheby_iters], mcp,
                                       calc_2norm));
                launches += 1;
                if (calc_2norm) {
                    TL_TRY(tl_fetch_scal(c));
                    error = c->scal_h->sums[0];
                    if (!inkernel) TL_TRY(sum_ranks(k, &error)); // in-kernel: already the sum over all ranks
                }
            } else {
                TL_TRY(tlk_cheby_iterate(c, c->cheby_alphas[num_cheby_iters], c->cheby_betas[num_cheby_iters]));
                TL_TRY(tlk_cheby_calc_u(c));
                launches += 2;
                if (calc_2norm) TL_TRY(norm2(c, k, TL_FIELD_R, &error));
            }
            if (num_cheby_iters == 1) {
                // cheby_calc_est_iterations, cheby_driver.c:146-160 (float logf/roundf as written)
                const double cn = info->eigmax / info->eigmin;
                const double it_alpha = DBL_EPSILON * bb / (4.0 * error);
                const double gamm = (sqrt(cn) - 1.0) / (sqrt(cn) + 1.0);
                est_iterations = (int)roundf(logf(it_alpha) / (2.0 * logf(gamm)));
            }
            // The one-pass kernel applies the reflective boundary by mirroring and (several ranks) delivers the
            // neighbours' halo itself: u's halo is only materialised once, after the loop.
            if (!fused || (multi && !inkernel)) TL_TRY(tl_halo_update(c, k, fields, 1));
            if (fabs(error) < o->eps) break;
        }
        if (inkernel) tlc_resident_seq_advance(k, n_seq); // same count on every rank: `error` is a global sum
        if (fused) {
            TL_TRY(tlk_field_home(c, TL_FIELD_U));
            if ((!multi || inkernel) && num_cheby_iters > 0) TL_TRY(tl_halo_update(c, k, fields, 1));
        }
    }
    info->iters_a = tt - num_cheby_iters + 1; // cheby_driver.c:73
    info->iters_b = num_cheby_iters;
    info->est_iters = est_iterations;
    info->total_iters = (tt < o->max_iters) ? tt + 1 : tt;
    info->error = error;
    info->kernel_launches += launches;
    return TL_OK;
}

// ppcg_driver.c:10-189
static int ppcg_solve(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double rx, double ry, tl_solve_info* info)
{
    const bool multi = k && tl_comms_size(k) > 1;
    int fields[TL_NUM_EXCHANGE_FIELDS];
    double rro = 0.0, error = 1e+10;
    long launches = 0;
    TL_TRY(tlk_reset_solve_scalars(c, o->eps, o->max_iters));
    TL_TRY(cg_init_driver(c, k, o, rx, ry, fields, &rro, multi));
    int tt = 0;
    bool ended = false;
    TL_TRY(cg_presteps(c, k, o, fields, &rro, &error, &tt, &ended, &launches));
    int num_ppcg_iters = 0;
    double theta = 0.0;
    const bool fused = o->fuse_p_into_w != 0;
    // see cheby_solve: sd's edge cells travel inside the one-pass kernel
    const bool inkernel = fused && multi && use_resident_multi(c, k);
    MultiCtx mc = c->mc;
    const MultiCtx* mcp = nullptr;
    int n_seq = 0;
    if (inkernel) {
        mc.sbase = mc.hbase = tlc_resident_seq_advance(k, 0);
        mcp = &mc;
    }
    if (!ended) {
        for (; tt < o->max_iters; ++tt) {
            num_ppcg_iters++;
            if (num_ppcg_iters == 1) {
                TL_TRY(eigenvalues(c, tt, &info->eigmin, &info->eigmax));
                cheby_coef(c, info->eigmin, info->eigmax, o->ppcg_inner_steps, &theta);
                // ppcg_init_driver, ppcg_driver.c:66-84.  Its second sum_over_ranks(rro) (:83) would
                // multiply an already global rro by num_ranks; the 1-rank semantics are kept.
                TL_TRY(tlk_calculate_residual(c));
                fields_reset(fields);
                fields[TL_FIELD_P] = 1;
                TL_TRY(tl_halo_update(c, k, fields, 1));
            }
            // ppcg_main_step_driver, ppcg_driver.c:87-149
            double pw = 0.0;
            TL_TRY(tlk_cg_calc_w(c, SCAL_IMM, false));
            TL_TRY(tl_fetch_scal(c));
            pw += c->scal_h->pw;
            TL_TRY(sum_ranks(k, &pw));
            const double alpha = rro / pw;
            TL_TRY(tlk_cg_calc_ur(c, SCAL_IMM, alpha, false)); // its rrn is discarded (:122)
            // ppcg_inner_iterations, ppcg_driver.c:152-189
            TL_TRY(tlk_ppcg_init(c, theta));
            fields_reset(fields);
            fields[TL_FIELD_SD] = 1;
            double rrn = 0.0;
            if (fused && o->ppcg_inner_steps > 0) {
                // one pass per inner step: ppcg_calc_ur + ppcg_calc_sd, sd double buffered, boundary by mirroring; the
                // last step also returns the 2-norm of r (:136-141).  Several ranks: one exchange of sd after
                // ppcg_init, then every step delivers the next step's halo itself.
                if (multi) TL_TRY(tl_halo_update(c, k, fields, 1));
                for (int pp = 0; pp < o->ppcg_inner_steps; ++pp) {
                    const bool last = (pp == o->ppcg_inner_steps - 1);
                    if (multi && !inkernel && pp > 0) TL_TRY(tl_halo_update(c, k, fields, 1));
                    mc.tl = n_seq++;
                    TL_TRY(tlk_ppcg_fused(c, c->cheby_alphas[pp], c->cheby_betas[pp], mcp, last));
                }
                TL_TRY(tl_fetch_scal(c));
                rrn = c->scal_h->sums[0];
                if (!inkernel) TL_TRY(sum_ranks(k, &rrn));
            } else {
                for (int pp = 0; pp < o->ppcg_inner_steps; ++pp) {
                    TL_TRY(tl_halo_update(c, k, fields, 1));
                    TL_TRY(tlk_ppcg_calc_ur(c));
                    TL_TRY(tlk_ppcg_calc_sd(c, c->cheby_alphas[pp], c->cheby_betas[pp]));
                }
                TL_TRY(norm2(c, k, TL_FIELD_R, &rrn));
            }
            launches += 3 + (fused ? 1 : 2) * o->ppcg_inner_steps;
            fields_reset(fields);
            fields[TL_FIELD_P] = 1;
            const double beta = rrn / rro;
            TL_TRY(tlk_cg_calc_p(c, SCAL_IMM, beta, false, false));
            error = rrn;
            rro = rrn;
            TL_TRY(tl_halo_update(c, k, fields, 1));
            if (fabs(error) < o->eps) break;
        }
        if (inkernel) tlc_resident_seq_advance(k, n_seq);
    }
    if (fused) TL_TRY(tlk_field_home(c, TL_FIELD_SD));
    info->iters_a = tt - num_ppcg_iters + 1; // ppcg_driver.c:59
    info->iters_b = num_ppcg_iters;
    info->total_iters = (tt < o->max_iters) ? tt + 1 : tt;
    info->error = error;
    info->kernel_launches += launches;
    return TL_OK;
}

// jacobi_driver.c:7-84
static int jacobi_solve(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double rx, double ry, tl_solve_info* info)
{
    int fields[TL_NUM_EXCHANGE_FIELDS];
    double error = 1e+10;
    TL_TRY(tlk_jacobi_init(c, o->coefficient, rx, ry));
    TL_TRY(tlk_copy_field(c, TL_FIELD_U0, TL_FIELD_U, true));
    fields_reset(fields);
    fields[TL_FIELD_U] = 1;
    int tt;
    for (tt = 0; tt < o->max_iters; ++tt) {
        TL_TRY(tlk_jacobi_iterate(c));
        info->kernel_launches += 2;
        if (tt % 50 == 0) {
            TL_TRY(tl_halo_update(c, k, fields, 1));
            TL_TRY(tlk_calculate_residual(c));
            TL_TRY(tlk_calculate_2norm(c, TL_FIELD_R));
        }
        TL_TRY(tl_fetch_scal(c));
        error = c->scal_h->sums[0];
        TL_TRY(sum_ranks(k, &error));
        TL_TRY(tl_halo_update(c, k, fields, 1));
        if (fabs(error) < o->eps) break;
    }
    info->iters_a = tt;
    info->total_iters = (tt < o->max_iters) ? tt + 1 : tt;
    info->error = error;
    return TL_OK;
}

extern "C" int tl_solve(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double rx, double ry, tl_solve_info* info)
{
    TL_CHECK_ARG(c && o && info, "null argument");
    TL_CHECK_ARG(o->max_iters >= 1 && o->max_iters <= c->max_iters, "max_iters exceeds the chunk's capacity");
    TL_CUDA(cudaSetDevice(c->device));
    memset(info, 0, sizeof(*info));
    const long l0 = g_tl_launches;
    cudaEvent_t e0, e1;
    TL_CUDA(cudaEventCreate(&e0));
    TL_CUDA(cudaEventCreate(&e1));
    TL_CUDA(cudaEventRecord(e0, c->stream));
    int rc;
    switch (o->solver) {
    case TL_SOLVER_CG: rc = cg_solve(c, k, o, rx, ry, info); break;
    case TL_SOLVER_CHEBY: rc = cheby_solve(c, k, o, rx, ry, info); break;
    case TL_SOLVER_PPCG: rc = ppcg_solve(c, k, o, rx, ry, info); break;
    case TL_SOLVER_JACOBI: rc = jacobi_solve(c, k, o, rx, ry, info); break;
    default: tl_set_error("unknown solver %d", o->solver); rc = TL_ERR_ARG;
    }
    if (rc == TL_OK) {
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        info->gpu_ms = ms;
        rc = tl_check_peer_timeout(c); // the stream is idle: every wait of this solve has either passed or marked the word
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    info->kernel_launches = g_tl_launches - l0;
    return rc;
}

// solve(), diffuse.c:23-64 + solve_finished_driver.c:7-43
extern "C" int tl_timestep(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double dt, double dx, double dy,
                           tl_solve_info* info)
{
    TL_CHECK_ARG(c && o && info, "null argument");
    TL_CUDA(cudaSetDevice(c->device));
    double dtmin = dt;
    if (k && tl_comms_size(k) > 1) TL_TRY(tl_comms_min(k, &dtmin)); // diffuse.c:33
    const double rx = dtmin / (dx * dx), ry = dtmin / (dy * dy);
    int fields[TL_NUM_EXCHANGE_FIELDS];
    fields_reset(fields);
    fields[TL_FIELD_ENERGY1] = 1;
    fields[TL_FIELD_DENSITY] = 1;
    TL_TRY(tl_halo_update(c, k, fields, 2));
    TL_TRY(tl_solve(c, k, o, rx, ry, info));
    const long l0 = g_tl_launches;
    // solve_finished_driver.c:7-43
    if (o->check_result) {
        double exact_error = 0.0;
        TL_TRY(tlk_calculate_residual(c));
        TL_TRY(tlk_calculate_2norm(c, TL_FIELD_R));
        if (k && tl_comms_size(k) > 1) {
            TL_TRY(tl_fetch_scal(c));
            exact_error = c->scal_h->sums[0];
            TL_TRY(tl_comms_sum(k, &exact_error));
        }
        (void)exact_error; // never printed by the reference either (solve_finished_driver.c:9-28)
    }
    TL_TRY(tlk_finalise(c));
    // fields_to_exchange keeps the solver's flags and adds ENERGY1 (solve_finished_driver.c:41)
    fields_reset(fields);
    fields[TL_FIELD_ENERGY1] = 1;
    switch (o->solver) {
    case TL_SOLVER_CG: fields[TL_FIELD_U] = fields[TL_FIELD_P] = 1; break;
    case TL_SOLVER_CHEBY: fields[TL_FIELD_U] = 1; if (info->iters_b == 0) fields[TL_FIELD_P] = 1; break;
    case TL_SOLVER_PPCG: fields[TL_FIELD_P] = 1; if (info->iters_b == 0) fields[TL_FIELD_U] = 1; break;
    case TL_SOLVER_JACOBI: fields[TL_FIELD_U] = 1; break;
    }
    TL_TRY(tl_halo_update(c, k, fields, 1));
    info->kernel_launches += g_tl_launches - l0;
    if (k && tl_comms_size(k) > 1) { // the final exchange of the step waits for the neighbours: report a time-out now
        TL_CUDA(cudaStreamSynchronize(c->stream));
        TL_TRY(tl_check_peer_timeout(c));
    }
    return TL_OK;
}

// field_summary_driver.c:8-30
extern "C" int tl_field_summary(tl_chunk* c, tl_comms* k, double* vol, double* mass, double* ie, double* temp)
{
    TL_TRY(tl_run_field_summary(c, vol, mass, ie, temp));
    if (k && tl_comms_size(k) > 1) {
        TL_TRY(tl_comms_sum(k, vol));
        TL_TRY(tl_comms_sum(k, mass));
        TL_TRY(tl_comms_sum(k, ie));
        TL_TRY(tl_comms_sum(k, temp));
    }
    return TL_OK;
}

extern "C" int tl_timestep_host(tl_chunk* c, tl_comms* k, const tl_solve_opts* o, double dt, double dx, double dy,
                                const double* density_host, double* energy_host, tl_solve_info* info,
                                double summary[4])
{
    TL_CHECK_ARG(c && density_host && energy_host && summary, "null argument");
    TL_CUDA(cudaSetDevice(c->device));
    const Geo& g = c->g;
    TL_CUDA(cudaMemcpy2DAsync(c->f[TL_FIELD_DENSITY] + g.off, (size_t)g.pitch * 8, density_host, (size_t)g.x * 8,
                              (size_t)g.x * 8, g.y, cudaMemcpyHostToDevice, c->stream));
    TL_CUDA(cudaMemcpy2DAsync(c->f[TL_FIELD_ENERGY1] + g.off, (size_t)g.pitch * 8, energy_host, (size_t)g.x * 8,
                              (size_t)g.x * 8, g.y, cudaMemcpyHostToDevice, c->stream));
    TL_TRY(tl_timestep(c, k, o, dt, dx, dy, info));
    TL_CUDA(cudaMemcpy2DAsync(energy_host, (size_t)g.x * 8, c->f[TL_FIELD_ENERGY1] + g.off, (size_t)g.pitch * 8,
                              (size_t)g.x * 8, g.y, cudaMemcpyDeviceToHost, c->stream));
    TL_TRY(tl_field_summary(c, k, &summary[0], &summary[1], &summary[2], &summary[3]));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    return TL_OK;
}

// Kernel micro-benchmark hook (bench.py roofline leg).  Uses the current field contents.
extern "C" int tl_time_kernel(tl_chunk* c, int which, int reps, double* ms_per_launch)
{
    TL_CHECK_ARG(c && ms_per_launch && reps > 0 && which >= 0 && which <= 17, "bad arguments");
    TL_CUDA(cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    TL_CUDA(cudaEventCreate(&e0));
    TL_CUDA(cudaEventCreate(&e1));
    if (which == 3) { // the fused kernel reads beta / conv from the device scalars
        TL_TRY(tlk_reset_solve_scalars(c, -1.0, 1 << 30));
    }
    for (int pass = 0; pass < 2; ++pass) { // pass 0 = warm-up
        const int n = pass ? reps : 3;
        if (pass) TL_CUDA(cudaEventRecord(e0, c->stream));
        for (int i = 0; i < n; ++i) {
            if (which == 0) TL_TRY(tlk_cg_calc_w(c, SCAL_IMM, false));
            else if (which == 1) TL_TRY(tlk_cg_calc_ur(c, SCAL_IMM, 1e-9, false));
            else if (which == 2) TL_TRY(tlk_cg_calc_p(c, SCAL_IMM, 0.5, false, false));
            else if (which == 3) TL_TRY(tlk_cg_calc_pw(c, false));
            else if (which == 4) TL_TRY(tlk_cheby_iterate(c, 0.5, 0.25));
            else if (which == 5) TL_TRY(tlk_cheby_calc_u(c));
            else if (which == 6) TL_TRY(tlk_ppcg_calc_ur(c));
            else if (which == 7) TL_TRY(tlk_ppcg_calc_sd(c, 0.5, 0.25));
            else if (which == 8) TL_TRY(tlk_jacobi_iterate(c));
            else if (which == 9) TL_TRY(tlk_calculate_residual(c));
            else if (which == 10) TL_TRY(tlk_calculate_2norm(c, TL_FIELD_R));
            else if (which == 11) TL_TRY(tlk_field_summary(c));
            else if (which == 12) TL_TRY(tlk_cg_init(c, TL_CONDUCTIVITY, 0.5, 0.5));
            else if (which == 13) TL_TRY(tlk_finalise(c));
            else if (which == 14) TL_TRY(tlk_copy_field(c, TL_FIELD_U0, TL_FIELD_U, true));
            else if (which == 15) TL_TRY(tlk_cheby_init(c, 2.0));
            else if (which == 16) TL_TRY(tlk_cheby_fused(c, 0.5, 0.25));
            else TL_TRY(tlk_ppcg_fused(c, 0.5, 0.25));
        }
        if (pass) TL_CUDA(cudaEventRecord(e1, c->stream));
    }
    TL_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    TL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = (double)ms / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return TL_OK;
}

static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
extern "C" int tl_timer_start(tl_chunk* c)
{
    TL_CHECK_ARG(c, "null chunk");
    TL_CUDA(cudaSetDevice(c->device));
    if (!g_t0) {
        TL_CUDA(cudaEventCreate(&g_t0));
        TL_CUDA(cudaEventCreate(&g_t1));
    }
    TL_CUDA(cudaEventRecord(g_t0, c->stream));
    return TL_OK;
}
extern "C" int tl_timer_stop(tl_chunk* c, double* elapsed_ms)
{
    TL_CHECK_ARG(c && elapsed_ms && g_t0, "timer not started");
    TL_CUDA(cudaSetDevice(c->device));
    TL_CUDA(cudaEventRecord(g_t1, c->stream));
    TL_CUDA(cudaEventSynchronize(g_t1));
    float ms = 0.f;
    TL_CUDA(cudaEventElapsedTime(&ms, g_t0, g_t1));
    *elapsed_ms = ms;
    return TL_OK;
}
