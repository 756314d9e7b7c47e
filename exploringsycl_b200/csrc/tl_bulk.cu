// tl_bulk.cu -- the fused p-update + matvec kernel of the resident CG loop on a TMA bulk-copy (cp.async.bulk)
// shared-memory row pipeline.
//
//   p = beta p + r   (cg.cpp:257-281, of iteration t-1)   fused with   w = A p ; p.w   (cg.cpp:137-195, of t)
//
// Same arithmetic, same per-thread accumulation order and same tile geometry (256 columns x `rows` rows, 128 compute
// threads x 2 columns) as k_cg_calc_pw / k_cg_calc_w in tl_kernels.cu: every field value and the p.w sum are
// bit-identical to the register-staged kernel and to the three-kernel iteration.  What changes is how the rows reach
// the SM:
//
//   * persistent CTAs (a multiple of the SM count), each walking the tile list with stride gridDim.x in the launch
//     order of the register kernel, so the L2 ping-pong with calc_ur's reversed traversal is kept;
//   * one producer warp per CTA streams row segments of p, r, kx, ky into an S-stage shared-memory ring with
//     cp.async.bulk (1-D TMA bulk copies, 2 KB each, completion counted in bytes on an mbarrier); it runs ahead of the
//     compute warps by S rows ACROSS tile boundaries, so the bytes in flight per SM no longer depend on registers or
//     on where a tile ends;
//   * the four compute warps read each row once from shared memory (LDS.128), take the left/right neighbours of the
//     updated p from the adjacent lanes by shuffle (the two edge lanes of a warp read them from the staged row instead
//     of issuing scalar global loads), slide rows j-1, j, j+1 through registers and store w and p with 128-bit stores.
//
// One rank only: the multi-rank resident loop always runs the register kernel (tlk_cg_calc_pw); the MULTI parameter
// of this kernel is kept so that its body stays line for line comparable with k_cg_calc_pw, but only MULTI = false is
// instantiated.
//
// STATUS (round 2, measured on B200, profiles/pw_pipeline_r02.txt): bit-identical to the register-staged kernel on
// every mesh tried, and exactly as fast under ncu (126.2 us vs 126.1 us at 4000^2, 749 MB of DRAM traffic and 72 % DRAM
// busy for both: the limiter is the HBM system, not how the rows reach the SM); in the resident loop it is 0.5-1 %
// slower (0.2546 vs 0.2522 ms per iteration) and, because its shared-memory carve-out differs from the streaming
// kernels', it does not overlap with them under programmatic dependent launch (0.2494 vs 0.2469).  The register-staged
// kernel therefore stays the default (tl_set_pw_pipeline(0, ..)); this one is kept selectable (mode 1, TL_PW=1:S:C).
// Staging rows before pdl_wait() is off by default (TL_PW_PRE): with all loop kernels forced onto one carve-out (so that
// CTAs of consecutive kernels really share SMs) solves were no longer reproducible to the last bit.
#include <stdint.h>
#include <stdlib.h>
#include "tl_internal.h"
#include "tl_device.cuh"

namespace {

constexpr int W = TL_TILE_COLS;          // 256 columns per tile
// stage layout, in doubles (every sub-array starts 16-byte aligned; the stage is a multiple of 128 bytes)
constexpr int SP = 0;                    // p   columns k0-2 .. k0+W+1   (W+4)
constexpr int SR = W + 4;                // r   columns k0-2 .. k0+W+1   (W+4)
constexpr int SKX = 2 * (W + 4);         // kx  columns k0   .. k0+W+1   (W+2, padded to W+4)
constexpr int SKY = 3 * (W + 4);         // ky  columns k0   .. k0+W-1   (W)
constexpr int STAGE_DOUBLES = ((3 * (W + 4) + W) + 15) / 16 * 16;
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
constexpr int NTHREADS = TL_TPB + 32;    // four compute warps + one producer warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D TMA bulk copy global -> shared; completion (in bytes) is signalled on `bar`.  dst, src and bytes are multiples of 16.
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

struct TileGeo {
    int tile, k0, j0, j1, nrow;
};
__device__ __forceinline__ TileGeo tile_geo(const Geo& g, int t, int ntiles, int gx, int rows, int rev)
{
    TileGeo q;
    q.tile = rev ? (ntiles - 1 - t) : t; // physical tile index by * gx + bx, as hot_tile() numbers them
    const int bx = q.tile % gx, by = q.tile / gx;
    q.k0 = g.hd + bx * W;
    q.j0 = g.hd + by * rows;
    q.j1 = min(q.j0 + rows, g.y - g.hd);
    q.nrow = q.j1 - q.j0 + 2;            // staged rows j0-1 .. j1
    return q;
}

template <int S, bool MULTI>
__global__ void __launch_bounds__(NTHREADS)
k_cg_calc_pw_bulk(Geo g, const double* p_in, double* __restrict__ p_out, const double* r, const double* kx,
                  const double* ky, double* __restrict__ w, double* __restrict__ d_alphas,
                  double* __restrict__ d_betas, RedArgs ra, int rows, int rev, int ext_mask, int gx, int gy,
                  int pre_mask, const MultiCtx mc)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * STAGE_BYTES);
    uint64_t* empty = full + S;
    DevScal* Sc = ra.S;
    const int ntiles = gx * gy;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);            // the producer's arrive.expect_tx; the copies complete the byte count
            mbar_init(&empty[s], TL_TPB / 32); // one arrival per compute warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Stages one row.  what & 1: arm the barrier with the row's full byte count; & 4: copy p; & 8: ky, kx; & 2: r.
    // p (written two launches back by the previous calc_pw), kx and ky may be requested before pdl_wait(); r is
    // written by the calc_ur just before this kernel.
    auto stage_row = [&](int it, const TileGeo& q, int qq, int what) {
        const int s = it % S;
        const int rem2 = (g.x - q.k0 + 1) & ~1; // columns from k0 to the end of the row, rounded up to 16 bytes
        const uint32_t np8 = 8u * (uint32_t)min(W + 4, rem2 + 2);
        const uint32_t nkx8 = 8u * (uint32_t)min(W + 2, rem2);
        const uint32_t nky8 = 8u * (uint32_t)min(W, rem2);
        const int jr = q.j0 - 1 + qq;
        int js = jr; // source row of p and r: mirrored at external faces (the reflective halo, not stored)
        if (qq == 0 && (ext_mask & 4) && q.j0 == g.hd) js = g.hd;
        if (qq == q.nrow - 1 && (ext_mask & 8) && q.j1 == g.y - g.hd) js = q.j1 - 1;
        const bool want_ky = qq > 0, want_kx = qq > 0 && qq < q.nrow - 1;
        double* st = stages + (size_t)s * STAGE_DOUBLES;
        const long bpr = (long)g.off + (long)js * g.pitch + q.k0 - 2;
        const long bk = (long)g.off + (long)jr * g.pitch + q.k0;
        if (what & 1) {
            mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
            mbar_expect_tx(&full[s], 2u * np8 + (want_ky ? nky8 : 0u) + (want_kx ? nkx8 : 0u));
        }
        if (what & 4) bulk_g2s(st + SP, p_in + bpr, np8, &full[s]);
        if (what & 8) {
            if (want_ky) bulk_g2s(st + SKY, ky + bk, nky8, &full[s]);
            if (want_kx) bulk_g2s(st + SKX, kx + bk, nkx8, &full[s]);
        }
        if (what & 2) bulk_g2s(st + SR, r + bpr, np8, &full[s]);
    };
    // pre_mask: what is requested before the dependency wait (bit 0: p, bit 1: kx and ky); the rest follows it
    const int what_pre = 1 | ((pre_mask & 1) ? 4 : 0) | ((pre_mask & 2) ? 8 : 0);
    const int what_post = 15 & ~what_pre;
    const bool producer = (threadIdx.x == TL_TPB);
    int pre = 0; // rows staged (without r) before the dependency wait
    if (producer) {
        int t = blockIdx.x, qq = 0;
        while (pre_mask >= 0 && pre < S && t < ntiles) {
            const TileGeo q = tile_geo(g, t, ntiles, gx, rows, rev);
            stage_row(pre, q, qq, what_pre);
            ++pre;
            if (++qq == q.nrow) {
                qq = 0;
                t += gridDim.x;
            }
        }
    }
    pdl_wait();
    pdl_trigger();

    // same head on one rank and on several (see k_cg_calc_pw): everything was settled by the preceding calc_ur's tail
    const bool skip = sld(&Sc->conv) != 0;
    const double beta = sld(&Sc->beta);
    if (skip) {
        // converged: this launch is a no-op, but the rows staged before the wait must land before the CTA (and its
        // shared memory) goes away
        if (producer) {
            int t = blockIdx.x, qq = 0;
            for (int it = 0; it < pre; ++it) {
                const TileGeo q = tile_geo(g, t, ntiles, gx, rows, rev);
                stage_row(it, q, qq, what_post);
                mbar_wait(&full[it % S], 0);
                if (++qq == q.nrow) {
                    qq = 0;
                    t += gridDim.x;
                }
            }
        }
        return;
    }

    const int xin = g.x - g.hd; // first column past the interior

    if (threadIdx.x >= TL_TPB) {
        // ------------------------------- producer warp (one elected lane) -------------------------------
        if (!producer) return;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const TileGeo q = tile_geo(g, t, ntiles, gx, rows, rev);
            for (int qq = 0; qq < q.nrow; ++qq, ++it) stage_row(it, q, qq, it < pre ? what_post : 15);
        }
        return;
    }

    // ----------------------------------------- compute warps -----------------------------------------
    const int tid = threadIdx.x, lane = tid & 31;
    const long pitch = g.pitch;
    const int khi = xin - 1, jhi = g.y - g.hd - 1;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const TileGeo q = tile_geo(g, t, ntiles, gx, rows, rev);
        const int kk = q.k0 + 2 * tid;
        const bool v0 = kk < xin, v1 = kk + 1 < xin;
        // which neighbours are mirrored (external face at the chunk edge) rather than read
        const bool mir_l = (ext_mask & 1) && kk == g.hd;
        const bool mir_r1 = (ext_mask & 2) && kk + 1 == khi; // my cell 1 is the last column
        const bool mir_r0 = (ext_mask & 2) && kk == khi;     // my cell 0 is the last column (cell 1 invalid)
        // internal faces (multi-rank): the ring values computed for the halo cells are the next iteration's old p there
        [[maybe_unused]] const bool halo_l = MULTI && v0 && !(ext_mask & 1) && kk == g.hd;
        [[maybe_unused]] const bool halo_r = MULTI && !(ext_mask & 2) && ((v1 && kk + 1 == khi) || (v0 && !v1 && kk == khi));
        double acc[1] = {0.0};
        double2 A = make_double2(0.0, 0.0), B = A, kxB = A, kyB = A;
        double Bl = 0.0, Br = 0.0, kxBr = 0.0;
        for (int qq = 0; qq < q.nrow; ++qq, ++it) {
            const int s = it % S;
            mbar_wait(&full[s], (it / S) & 1);
            const double* st = stages + (size_t)s * STAGE_DOUBLES;
            const double2 pv = lds2(st + SP + 2 * tid + 2), rv = lds2(st + SR + 2 * tid + 2);
            double2 Cc;
            Cc.x = beta * pv.x + rv.x;
            Cc.y = beta * pv.y + rv.y;
            double Cl = __shfl_up_sync(0xffffffffu, Cc.y, 1);
            double Cr = __shfl_down_sync(0xffffffffu, Cc.x, 1);
            if (lane == 0) Cl = beta * st[SP + 2 * tid + 1] + st[SR + 2 * tid + 1];
            if (lane == 31) Cr = beta * st[SP + 2 * tid + 4] + st[SR + 2 * tid + 4];
            double2 kxc = make_double2(0.0, 0.0), kyc = kxc;
            double kxr = 0.0;
            if (qq > 0) {
                kyc = lds2(st + SKY + 2 * tid);
                if (qq < q.nrow - 1) {
                    kxc = lds2(st + SKX + 2 * tid);
                    kxr = __shfl_down_sync(0xffffffffu, kxc.x, 1);
                    if (lane == 31) kxr = st[SKX + 2 * tid + 2];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]); // this warp holds everything it needs from the stage in registers
            if (mir_l) Cl = Cc.x;
            if (mir_r1) Cr = Cc.y;
            const int jr = q.j0 - 1 + qq;
            if (qq == 0) {
                if constexpr (MULTI) { // bottom halo row of p_out on an internal face
                    if (v0 && !(ext_mask & 4) && q.j0 == g.hd)
                        st_pair(p_out + (long)g.off + (long)jr * pitch + kk, Cc, v1);
                }
            } else if (qq >= 2 && v0) {
                const long iu = (long)g.off + (long)(jr - 1) * pitch + kk; // the row being completed: jr - 1
                double2 wv;
                const double pr0 = mir_r0 ? B.x : B.y; // right neighbour of cell 0
                wv.x = smvp(kxB.x, kxB.y, kyB.x, kyc.x, B.x, Bl, pr0, A.x, Cc.x);
                wv.y = smvp(kxB.y, kxBr, kyB.y, kyc.y, B.y, B.x, Br, A.y, Cc.y);
                st_pair(w + iu, wv, v1);
                st_pair(p_out + iu, B, v1);
                if constexpr (MULTI) {
                    if (halo_l) p_out[iu - 1] = Bl;
                    if (halo_r) {
                        if (v1) p_out[iu + 2] = Br;
                        else p_out[iu + 1] = B.y;
                    }
                    if (!(ext_mask & 8) && jr - 1 == jhi) st_pair(p_out + iu + pitch, Cc, v1); // top halo row
                }
                acc[0] += wv.x * B.x;
                if (v1) acc[0] += wv.y * B.y;
            }
            A = B;
            B = Cc;
            Bl = Cl;
            Br = Cr;
            kxB = kxc;
            kxBr = kxr;
            kyB = kyc;
        }
        double tot[1];
        if (grid_reduce<1, true>(acc, ra, q.tile, ntiles, tot) && threadIdx.x < 32) {
            const int itn = sld(&Sc->iters);
            double pw = tot[0];
            if constexpr (MULTI) pw = mc_allsum_warp(mc, 0, tot[0], Sc); // sum_over_ranks(pw), cg_driver.c:85
            if (threadIdx.x == 0) {
                Sc->pw = pw;
                const double alpha = Sc->rro / pw;
                Sc->alpha = alpha;
                d_alphas[itn] = alpha;
                Sc->p_pending = 0; // the pending p update of the previous iteration has now been applied
            }
        }
    }
}

int g_pw_mode = 0;     // 0: register-staged k_cg_calc_pw (default, see STATUS above); 1: bulk-copy pipeline (this file)
int g_pw_stages = 4;
int g_pw_ctas = 0;     // persistent CTAs per SM (0 = what the shared-memory footprint allows, at most 8)
int g_num_sms = 0;
int g_pw_pre = -1;     // staged before the PDL wait: -1 nothing (default); bit 0 p, bit 1 kx/ky (experiments: TL_PW_PRE)

template <int S, bool MULTI>
int launch_pw(tl_chunk* c, bool rev, const MultiCtx* mc, int rows, dim3 tiles, bool pdl)
{
    static int max_ctas = -1; // per instantiation
    const size_t smem = (size_t)S * STAGE_BYTES + 2 * S * sizeof(uint64_t);
    auto kern = k_cg_calc_pw_bulk<S, MULTI>;
    if (max_ctas < 0) {
        TL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const char* cv = getenv("TL_PW_CARVEOUT"); // experiments: -1 leaves the driver's choice
        const int carve = cv ? atoi(cv) : 100;
        if (carve >= 0) TL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        int n = 0;
        TL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, NTHREADS, smem));
        max_ctas = n > 0 ? n : 1;
    }
    if (!g_num_sms) {
        TL_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, c->device));
    }
    int per_sm = max_ctas < 8 ? max_ctas : 8;
    if (g_pw_ctas > 0 && g_pw_ctas < per_sm) per_sm = g_pw_ctas;
    const int ntiles = (int)(tiles.x * tiles.y);
    int grid = g_num_sms * per_sm;
    if (grid > ntiles) grid = ntiles;
    RedArgs ra{c->partials, c->gpartials, c->gcount, c->partial_cap, c->gpartial_cap, c->scal};
    TL_CUDA(tl_launch(kern, dim3(grid), dim3(NTHREADS), smem, c->stream, pdl, c->g, c->f[TL_FIELD_P], c->p2,
                      c->f[TL_FIELD_R], c->f[TL_FIELD_KX], c->f[TL_FIELD_KY], c->f[TL_FIELD_W], c->d_alphas, c->d_betas,
                      ra, rows, rev ? 1 : 0, tlk_external_mask(c), (int)tiles.x, (int)tiles.y, g_pw_pre,
                      MULTI ? *mc : g_single_ctx));
    return TL_OK;
}

} // namespace

// mode 0: register-staged kernel; 1: bulk-copy pipeline.  stages in {3, 4, 6, 8}; ctas_per_sm 0 = automatic.
extern "C" int tl_set_pw_pipeline(int mode, int stages, int ctas_per_sm)
{
    if ((mode != 0 && mode != 1) || (stages != 3 && stages != 4 && stages != 6 && stages != 8) || ctas_per_sm < 0 ||
        ctas_per_sm > 16) {
        tl_set_error("tl_set_pw_pipeline: bad arguments");
        return TL_ERR_ARG;
    }
    g_pw_mode = mode;
    g_pw_stages = stages;
    g_pw_ctas = ctas_per_sm;
    return TL_OK;
}

bool tlk_pw_uses_bulk()
{
    static bool env_done = false;
    if (!env_done) { // TL_PW="mode:stages:ctas" overrides the defaults at first use (experiments only)
        env_done = true;
        const char* e = getenv("TL_PW");
        int m, s, n;
        if (e && sscanf(e, "%d:%d:%d", &m, &s, &n) == 3) tl_set_pw_pipeline(m, s, n);
        const char* pm = getenv("TL_PW_PRE");
        if (pm) g_pw_pre = atoi(pm);
    }
    return g_pw_mode == 1;
}

// p_in = c->f[P], p_out = c->p2 (allocated by the caller); the caller swaps them after the launch.
int tlk_cg_calc_pw_bulk(tl_chunk* c, bool rev, const MultiCtx* mc, int rows, bool pdl)
{
    dim3 tiles = tlk_hot_grid(c, rows);
    TL_TRY(tlk_hot_check(c, tiles));
    if (mc && mc->num_ranks > 1) {
        tl_set_error("the bulk-copy calc_pw pipeline runs on one rank only");
        return TL_ERR_ARG;
    }
#define PW_CASE(S)                                                            \
    case S:                                                                   \
        TL_TRY((launch_pw<S, false>(c, rev, mc, rows, tiles, pdl)));          \
        break
    switch (g_pw_stages) {
        PW_CASE(3);
        PW_CASE(6);
        PW_CASE(8);
    default:
        PW_CASE(4);
    }
#undef PW_CASE
    ++g_tl_launches;
    TL_CUDA(cudaGetLastError());
    return TL_OK;
}
