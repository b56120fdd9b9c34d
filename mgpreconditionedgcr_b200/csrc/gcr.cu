// csrc/gcr.cu -- GCR<num_type> (reference: src/GCR.h:158-302) driven from the host with every
// vector and every scalar (alpha, beta, norms) resident on the device.  Per iteration: 4 fused kernels + the operator
// apply, one 8-byte read-back for the stopping test that overlaps with the direction update.
#include <math.h>

#include <algorithm>

#include "kernels_blas.cuh"
#include "ops.cuh"

int vec_dot_dev(mgcr_ctx* ctx, int64_t n, const c128* a, const c128* b, double* d_out, bool dist, int64_t n_global = 0);
int vec_norm2_dev(mgcr_ctx* ctx, int64_t n, const c128* a, double* d_out, bool dist, int64_t n_global = 0);
int vec_normalise(mgcr_ctx* ctx, int64_t n, c128* a, bool dist, int64_t n_global = 0);

int gcr_solve_lr(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right, const c128* rhs, c128* x, double* hist,
                 int hist_cap, int* iters_out, bool x_zero = false);
int gcr_solve_small(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, const c128* rhs, c128* x, double* hist, int hist_cap,
                    int* iters_out, int storage, int restart, int* handled, bool x_zero);

struct GcrOp : mgcr_op {
    mgcr_op* A = nullptr;
    mgcr_gcr_param prm;
    mgcr_op* right = nullptr;
    mgcr_op* left = nullptr;
    c128* d_rand2 = nullptr;   // cached init_rand(2) start vector (src/GCR.h:63-68)
    ~GcrOp() override { dev_free(ctx, d_rand2); }
    int apply(const c128* x, c128* y) override;
};

// ---- launch tables: exact history length NH (1..16), U elements per thread (dot) / MINB resident CTAs per SM (update)
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

template <int NK, int KS>
static void launch_dot_hist(mgcr_ctx* ctx, const RedGeom& rg, const c128* Ar, const c128* Aps, int64_t stride, const HistList& hl, int nh,
                            int std_conj, double* out, const double* guard, double tol2, const ArPush& push) {
    launch_pdl(ctx, k_gcr_dot_hist<NK, KS>, rg.G, RED_THREADS, 0, rg, Ar, Aps, stride, hl, nh, std_conj, out, ctx->d_partials, ctx->d_ticket, guard, tol2, push);
}
template <int NH>
static int launch_dot_hist_tma(mgcr_ctx* ctx, const RedGeom& rg0, const c128* Ar, const c128* Aps, int64_t stride, const HistList& hl, int std_conj,
                               double* out, const double* guard, double tol2, const ArPush& push) {
    MGCR_TRY(ensure_dyn_smem(ctx, (const void*)k_gcr_dot_hist_tma<NH>, 200 * 1024));
    // tile = 256*ept elements of each of the 1+NH vectors; ring of `stages` tiles in ~150 KB (measured: scripts/kbench_dot5.cu)
    const int ept = NH <= 3 ? 4 : NH <= 7 ? 2 : 1;
    const size_t stage_bytes = (size_t)(1 + NH) * RED_THREADS * ept * sizeof(c128);
    const int stages = (int)std::max<size_t>(2, std::min<size_t>(4, (150 * 1024) / stage_bytes));
    // one CTA per SM (the ring takes the SM's shared memory); each slab's tiles are dealt over the G CTAs
    RedGeom rg = rg0;
    const int64_t tiles = (rg.L + (int64_t)RED_THREADS * ept - 1) / ((int64_t)RED_THREADS * ept);
    rg.G = (int)std::min<int64_t>(148, tiles);
    launch_pdl(ctx, k_gcr_dot_hist_tma<NH>, rg.G, RED_THREADS, stages * stage_bytes, rg, Ar, Aps, stride, hl, std_conj, ept, stages, out,
               ctx->d_partials, ctx->d_ticket, guard, tol2, push);
    return MGCR_OK;
}

// history length -> kernel.  Short histories (and short vectors): register-staged kernel, KS thread groups per CTA with
// <= NK vectors each; nh >= 3 on long vectors: TMA-staged ring (measured on B200, profiles/r01_kbench_dot.txt).
static int dot_hist(mgcr_ctx* ctx, int nh, const RedGeom& rg, int64_t n_global, const c128* Ar, const c128* Aps, int64_t stride, const HistList& hl,
                    int std_conj, double* out, const double* guard, double tol2, const ArPush& push) {
    if (ctx->dot_tma && nh >= 3 && n_global >= ((int64_t)1 << 23) && rg.L >= ((int64_t)1 << 17)) {   // (the GLOBAL length decides: the same kernel at every GPU count)
        switch (nh) {
#define C(NH) case NH: return launch_dot_hist_tma<NH>(ctx, rg, Ar, Aps, stride, hl, std_conj, out, guard, tol2, push);
            C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
#undef C
        }
    }
#define GO(NK, KS) launch_dot_hist<NK, KS>(ctx, rg, Ar, Aps, stride, hl, nh, std_conj, out, guard, tol2, push)
    if (nh <= 3) GO(3, 1);
    else if (nh <= 8) GO(4, 2);
    else GO(4, 4);
#undef GO
    return MGCR_OK;
}

template <int NH, int MINB>
static void launch_update_p(mgcr_ctx* ctx, const RedGeom& rg, const c128* z, const c128* Ar, const c128* r, c128* ps, c128* Aps,
                            int64_t stride, const BetaList& bl, int cur, int first, int last, c128* acc_p, c128* acc_Ap, int std_conj,
                            int bden_off, int keep_Ap, double* scal, double* red_anum, const double* guard, double tol2, const ArWait& aw, const ArPush& push) {
    launch_pdl(ctx, k_gcr_update_p<NH, MINB>, rg.G, RED_THREADS, 0, rg, z, Ar, r, ps, Aps, stride, bl, cur, first, last, acc_p, acc_Ap, std_conj,
               bden_off, keep_Ap, (const double*)scal, red_anum, ctx->d_partials, ctx->d_ticket, guard, tol2, aw, push);
}
static void update_p(mgcr_ctx* ctx, int nh, const RedGeom& rg, const c128* z, const c128* Ar, const c128* r, c128* ps, c128* Aps,
                     int64_t stride, const BetaList& bl, int cur, int first, int last, c128* acc_p, c128* acc_Ap, int std_conj, int bden_off,
                     int keep_Ap, double* scal, double* red_anum, const double* guard, double tol2, const ArWait& aw, const ArPush& push) {
    static const int minb_env = env_int("MGCR_UPD_MINB", 0);   // experiment knob
    const int minb = minb_env ? minb_env : 4;
#define ARGS ctx, rg, z, Ar, r, ps, Aps, stride, bl, cur, first, last, acc_p, acc_Ap, std_conj, bden_off, keep_Ap, scal, red_anum, guard, tol2, aw, push
#define C(NH) case NH: if (minb >= 4) launch_update_p<NH, 4>(ARGS); else if (minb == 3) launch_update_p<NH, 3>(ARGS); else launch_update_p<NH, 2>(ARGS); break;
    switch (nh) {
        C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
    }
#undef C
#undef ARGS
}

// stack of pinned read-back slots / events so that nested solves (preconditioners) never share one
struct SolveSlot { double* h; cudaEvent_t ev; };
static int acquire_slot(mgcr_ctx* ctx, int depth, SolveSlot* s) {
    ARG_CHECK(depth < 12, "GCR: solver nesting deeper than 12");
    while ((int)ctx->depth_events.size() <= depth) {   // owned by the context, destroyed with it
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->depth_events.push_back(e);
    }
    s->ev = ctx->depth_events[(size_t)depth];
    s->h = ctx->h_pinned + 16 + 8 * depth;
    return MGCR_OK;
}

static thread_local int g_depth = 0;
struct DepthGuard { DepthGuard() { g_depth++; } ~DepthGuard() { g_depth--; } };

// x_zero: the caller's start vector is zero and x need not hold it -- the first x += alpha p WRITES alpha p (0 + t = t exactly),
// so neither a memset of x nor the first read of it happens (the multigrid cycle starts every smoother / coarse solve that way:
// a 2.1 GB memset and a 2.1 GB read per level-0 cycle at 512^3)
int gcr_solve(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* right, const c128* rhs, c128* x, double* hist,
              int hist_cap, int* iters_out, bool x_zero = false);
int gcr_solve(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* right, const c128* rhs, c128* x, double* hist,
              int hist_cap, int* iters_out, bool x_zero) {
    return gcr_solve_lr(ctx, A, prm, nullptr, right, rhs, x, hist, hist_cap, iters_out, x_zero);
}

// `left`: the reference's left preconditioner, replicated operation for operation (src/GCR.h:201-204, 245-247): r <- L(r) ONCE,
// after p = rhs and Ap = A p have been formed from the raw right-hand side; then Ar <- L(A r) in every iteration.  The first
// direction is therefore unpreconditioned and what the loop reduces (and prints) is L(rhs) - sum alpha Ap, not a residual of
// the original system -- the reference's behaviour, kept because it is deterministic (SURVEY.md Appendix B, Q4 for the right side).
int gcr_solve_lr(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right, const c128* rhs, c128* x, double* hist,
                 int hist_cap, int* iters_out, bool x_zero) {
    const int64_t n = A->n_local;
    ARG_CHECK(prm->truncation == 0 || prm->restart == 0, "Do not support concurrent restarting and truncation. (src/GCR.h:165)");
    ARG_CHECK(prm->truncation >= 0 && prm->restart >= 0 && prm->max_iter >= 0, "GCR: negative parameter");
    if (prm->truncation == 0 && prm->restart == 0 && prm->verbose) printf("WARNING: Full GCR solve could incur high memory usage!\n");
    int storage = prm->max_iter, restart = prm->max_iter;
    if (prm->truncation != 0) storage = prm->truncation;
    if (prm->restart != 0) { restart = prm->restart; storage = restart; }
    if (storage < 1) storage = 1;   // max_iter = 0 (one smoothing step): the reference writes slot 0 of a 0-length array
    if (restart < 1) restart = 1;
    const bool aliased = (rhs == x);
    const int std_conj = prm->std_conj;
    const bool dist = A->distributed;   // vectors are row slabs: partial inner products are all-reduced (NCCL)
    if (!right && !left) {   // small operators: the whole solve as one persistent cooperative kernel (gcr_small.cu)
        int handled = 0;
        MGCR_TRY(gcr_solve_small(ctx, A, prm, rhs, x, hist, hist_cap, iters_out, storage, restart, &handled, x_zero && rhs != x));
        if (handled) return MGCR_OK;
    }
    DepthGuard dg;
    SolveSlot slot;
    MGCR_TRY(acquire_slot(ctx, g_depth - 1, &slot));

    c128 *r = nullptr, *Ar = nullptr, *z = nullptr, *ps = nullptr, *Aps = nullptr, *acc_p = nullptr, *acc_Ap = nullptr, *Lt = nullptr;
    double* scal = nullptr;
    static const int64_t ring_pad = getenv("MGCR_RING_PAD") ? atoll(getenv("MGCR_RING_PAD")) : 0;   // experiment knob
    const int64_t stride = n + ring_pad;
    const int bden_off = S_BNUM + 2 * storage;
    const int nscal = bden_off + storage;
    // Distributed solves keep this rank's PARTIAL sums in a second block of the same layout (`red`): the kernels reduce into it,
    // the all-reduce reads it and writes the global sums into `scal`.  `scal` therefore only ever holds global values -- the
    // device-side stopping tests of a blind solve decide the same thing on every rank -- and an all-reduce that is repeated after
    // its producer kernel has been skipped (solve converged) reproduces the same sums instead of adding global values up again
    // (round 1 reduced in place: with a loose smoother tolerance ||r||^2 doubled per skipped iteration and the solve resumed).
    int st = MGCR_OK;
    auto cleanup = [&]() {
        dev_free(ctx, r); dev_free(ctx, Ar); dev_free(ctx, z); dev_free(ctx, ps); dev_free(ctx, Aps);
        dev_free(ctx, acc_p); dev_free(ctx, acc_Ap); dev_free(ctx, scal); dev_free(ctx, Lt);
    };
#define GTRY(expr) do { st = (expr); if (st != MGCR_OK) { cleanup(); return st; } } while (0)
#define GCUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { mgcr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); cleanup(); return MGCR_ERR_CUDA; } } while (0)
    GTRY(dev_alloc_t(ctx, (size_t)n, &r));
    GTRY(dev_alloc_t(ctx, (size_t)n, &Ar));
    if (right) GTRY(dev_alloc_t(ctx, (size_t)n, &z));
    if (left) GTRY(dev_alloc_t(ctx, (size_t)n, &Lt));
    // a solve of max_iter iterations stores at most max_iter + 1 directions (slot indices stay below that): short smoother /
    // coarse solves with a long nominal restart do not hold the unused ring slots (15 GB per cycle on the 1024x512x512 lattice).
    // Long rings (full GCR: storage = max_iter, src/GCR.h:171-185, whose `new Field[storage]` fills lazily, :208-209) start with
    // 32 slots and double when the iteration reaches the end: a full GCR that converges in tens of iterations never holds more.
    int slots_alloc = std::min(std::min(storage, prm->max_iter + 1), 32);
    GTRY(dev_alloc_t(ctx, (size_t)stride * slots_alloc, &ps));
    GTRY(dev_alloc_t(ctx, (size_t)stride * slots_alloc, &Aps));
    auto grow_ring = [&](int need_slot) -> int {
        if (need_slot < slots_alloc) return MGCR_OK;
        const int cap = std::min(storage, std::max(2 * slots_alloc, need_slot + 1));
        c128 *nps = nullptr, *nAps = nullptr;
        int s2 = dev_alloc_t(ctx, (size_t)stride * cap, &nps);
        if (s2 == MGCR_OK) s2 = dev_alloc_t(ctx, (size_t)stride * cap, &nAps);
        if (s2 != MGCR_OK) {
            dev_free(ctx, nps);
            mgcr_set_error("GCR: no device memory for %d stored search directions of %lld elements; bound the history with GCR_Param::truncation or "
                           "::restart (src/GCR.h:171-185)", cap, (long long)n);
            return MGCR_ERR_OOM;
        }
        cudaError_t e = cudaMemcpyAsync(nps, ps, sizeof(c128) * (size_t)stride * slots_alloc, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nAps, Aps, sizeof(c128) * (size_t)stride * slots_alloc, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { dev_free(ctx, nps); dev_free(ctx, nAps); mgcr_set_error("GCR ring growth: %s", cudaGetErrorString(e)); return MGCR_ERR_CUDA; }
        dev_free(ctx, ps); dev_free(ctx, Aps);
        ps = nps; Aps = nAps; slots_alloc = cap;
        return MGCR_OK;
    };
    if (storage > GCR_CHUNK) { GTRY(dev_alloc_t(ctx, (size_t)n, &acc_p)); GTRY(dev_alloc_t(ctx, (size_t)n, &acc_Ap)); }
    GTRY(dev_alloc_t(ctx, (size_t)nscal * (dist ? 2 : 1), &scal));
    GCUDA(cudaMemsetAsync(scal, 0, sizeof(double) * nscal * (dist ? 2 : 1), ctx->stream));
    double* const red = dist ? scal + nscal : scal;

    static const int grid_per_sm = getenv("MGCR_GRID_PER_SM") ? atoi(getenv("MGCR_GRID_PER_SM")) : 4;   // experiment knob
    // reduction / streaming shape of every kernel of this solve: independent of the number of GPUs (common.cuh, RedGeom)
    const RedGeom rg = red_geom(ctx, n, dist ? A->n_global : n, grid_per_sm, 2);
    const int grid = rg.G;   // every CTA plays its part in each of the rank's virtual slabs in turn
    // r = rhs ; p = z = R(r) or r ; Ap = A p                                                   (GCR.h:189-192)
    // the operator / preconditioner read rhs directly; r (and p when there is no preconditioner) are written by the init
    // kernel in the pass that forms the first inner products
    if (right) GTRY(right->apply(rhs, ps));
    GTRY(A->apply(right ? ps : rhs, Aps));
    // Short solves nobody watches (the smoothers of the multigrid cycle): no read-back at all, the kernels carry the
    // stopping test themselves (gcr_converged) and the host enqueues max_iter iterations back to back.
    // A right preconditioner does not change that: its applies run whether or not the solve has converged (wasted work in the
    // rare case that a <= 4-iteration solve converges early, but no host round trip per iteration -- the K-cycle's coarse solves
    // were 235 of the 260 host synchronisations of a 512^3 solve in round 1).
    static const int blind_precond = env_int("MGCR_BLIND_PRECOND", 1);   // experiment knob: 0 = round-1 behaviour
    const bool blind = !hist && !iters_out && !prm->verbose && !aliased && !left && prm->max_iter <= 4 && (!right || blind_precond);
    // Distributed blind solves over peer memory: every all-reduce travels inside the kernel that produces the sums (its last CTA
    // posts them to all ranks) and the kernel that needs them (every CTA collects them), common.cuh ArPush / ArWait.  `pushA` /
    // `waitA` carry <r,Ap>, <Ap,Ap> (+ the two norms at the start), `pushB` / `waitB` ||r||^2 with the history inner products.
    const bool fold = blind && dist && p2p_enabled(ctx);
    ArPush pushA, pushB;
    ArWait waitA, waitB;
    memset(&pushA, 0, sizeof pushA); memset(&pushB, 0, sizeof pushB); memset(&waitA, 0, sizeof waitA); memset(&waitB, 0, sizeof waitB);
    if (fold) (void)p2p_allreduce_fold(ctx, 5, red, scal, &pushA, &waitA);   // (refused: the descriptors stay off, stand-alone all-reduce below)
    // (the working copy r = rhs is not made here: the first x / r update reads rhs and writes r.  The aliased solve (rhs IS x, only
    // Arnoldi's inverse iteration) keeps the plain copy, and a left preconditioner forms r itself.)
    // MGCR_BLIND_LEAN=0 (test knob): every vector the reference's loop writes is written -- r = rhs and p0 = rhs copied by the init pass,
    // A p stored by every direction update, the last iteration of a blind solve a full x / r update.  Same bits either way
    // (tests/test_gpu_variants.py).
    static const int lean = env_int("MGCR_BLIND_LEAN", 1);
    const bool r_from_rhs = lean && !aliased && !left;
    // Unpreconditioned blind solve that never comes back to ring slot 0: the first direction p0 = rhs is READ from the right-hand
    // side wherever slot 0 is addressed instead of being copied into the ring (16 bytes per element and solve less)
    const bool p0_is_rhs = lean && blind && !right && !left && !aliased && prm->max_iter <= restart && prm->max_iter <= storage;
    KLAUNCH(ctx, "gcr_init", ((right || p0_is_rhs ? 48. : 64.) - (r_from_rhs ? 16. : 0.)) * n, (launch_pdl(ctx, k_gcr_init, grid, RED_THREADS, 0, rg, rhs, (const c128*)Aps, std_conj, r_from_rhs ? (c128*)nullptr : r, (right || p0_is_rhs) ? (c128*)nullptr : ps, ctx->d_partials, ctx->d_ticket, red, pushA)));
    GCUDA(cudaGetLastError());
    if (left) {
        // r <- L(r) (GCR.h:201-204): the first alpha and the step-0 print use the preconditioned r with the UNpreconditioned Ap
        GTRY(left->apply(rhs, r));
        GTRY(vec_dot_dev(ctx, n, std_conj ? Aps : r, std_conj ? r : Aps, red + S_ANUM, false, dist ? A->n_global : n));   // <r,Ap> (or <Ap,r>), local part
        GTRY(vec_norm2_dev(ctx, n, r, red + S_RR, false, dist ? A->n_global : n));
    }
    if (dist && !pushA.seq) GTRY(dist_allreduce_sum2(ctx, red, scal, 5));
    const double* guard = blind ? scal : nullptr;
    const double tol2 = prm->tol * prm->tol;
    double bb = 1., rr = 1.;
    if (!blind) {
        GCUDA(cudaMemcpyAsync(slot.h, scal + S_BB, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        {
            HostTimer ht(&ctx->host_sync_ms, &ctx->host_sync_calls);
            GCUDA(cudaStreamSynchronize(ctx->stream));
        }
        bb = slot.h[0];
        rr = bb;
    }
    if (left && !blind) {   // step 0 shows ||L(rhs)|| / ||rhs|| (GCR.h:214)
        GCUDA(cudaMemcpyAsync(slot.h, scal + S_RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        GCUDA(cudaStreamSynchronize(ctx->stream));
        rr = slot.h[0];
    }
    if (hist && hist_cap > 0) hist[0] = sqrt(rr) / sqrt(bb);
    if (prm->verbose) printf("Step %d residual norm = %.10e\n", 0, sqrt(rr) / sqrt(bb));

    int iter = 0, g = 0, cur = 0;
    do {
        g++; iter++;
        // alpha, x += alpha p, r -= alpha Ap, ||r||^2                                            (GCR.h:230-233)
        const int xz = (x_zero && !aliased && g == 1) ? 1 : 0;
        const c128* pcur = (p0_is_rhs && cur == 0) ? rhs : ps + (int64_t)cur * stride;
        const c128* Apcur = Aps + (int64_t)cur * stride;
        if (lean && blind && g >= prm->max_iter) {
            // the last pass of a solve nobody watches: x is all that is left to compute (no r, no Ap, no norm)
            const int gx = stream_grid(ctx, n, 4, 4);
            if (waitA.seq) KLAUNCH(ctx, "gcr_update_x", (xz ? 32. : 48.) * n, (launch_pdl(ctx, k_gcr_update_x<true>, gx, RED_THREADS, 0, n, pcur, x, (const double*)scal, xz, guard, tol2, waitA)));
            else KLAUNCH(ctx, "gcr_update_x", (xz ? 32. : 48.) * n, (launch_pdl(ctx, k_gcr_update_x<false>, gx, RED_THREADS, 0, n, pcur, x, (const double*)scal, xz, guard, tol2, waitA)));
        } else {
            const c128* r_in = (g == 1 && r_from_rhs) ? rhs : r;   // first iteration: r = rhs, never copied
            if (waitA.seq) KLAUNCH(ctx, "gcr_update_xr", (xz ? 80. : 96.) * n, (launch_pdl(ctx, k_gcr_update_xr<true>, grid, RED_THREADS, 0, rg, pcur, Apcur, x, r_in, r, scal, red + S_RR, bden_off + cur, xz, ctx->d_partials, ctx->d_ticket, guard, tol2, waitA)));
            else KLAUNCH(ctx, "gcr_update_xr", (xz ? 80. : 96.) * n, (launch_pdl(ctx, k_gcr_update_xr<false>, grid, RED_THREADS, 0, rg, pcur, Apcur, x, r_in, r, scal, red + S_RR, bden_off + cur, xz, ctx->d_partials, ctx->d_ticket, guard, tol2, waitA)));
        }
        GCUDA(cudaGetLastError());
        if (aliased) {   // rhs IS x (src/MG.h:102): the stopping test sees the norm of the updated vector
            KLAUNCH(ctx, "vec_norm2", 16. * n, (launch_pdl(ctx, k_norm2, grid, RED_THREADS, 0, rg, (const c128*)x, ctx->d_partials, ctx->d_ticket, red + S_BB)));
            GCUDA(cudaGetLastError());
            if (dist) GTRY(dist_allreduce_sum2(ctx, red + S_BB, scal + S_BB, 1));
        }
        const bool final_iter = (g >= prm->max_iter);   // nothing after the x update is observable on the last pass
        const c128* zz = r;
        int lim = 0;
        if (!final_iter) {
            if (right) { GTRY(right->apply(r, z)); zz = z; }                                      // flexible form of GCR.h:236-238
            GTRY(A->apply(zz, Ar));                                                               // GCR.h:242
            if (left) { GTRY(left->apply(Ar, Lt)); std::swap(Ar, Lt); }                           // GCR.h:245-247
            lim = std::min(storage, iter);                                                        // GCR.h:251
            memset(&pushB, 0, sizeof pushB); memset(&waitB, 0, sizeof waitB);
            if (fold && lim <= GCR_CHUNK) (void)p2p_allreduce_fold(ctx, 1 + 2 * lim, red + S_RR, scal + S_RR, &pushB, &waitB);
            for (int c0 = 0; c0 < lim; c0 += GCR_CHUNK) {                                         // GCR.h:257-258 numerators
                HistList hl;
                const int cnt = std::min((int)GCR_CHUNK, lim - c0);
                for (int k = 0; k < GCR_CHUNK; k++) hl.slot[k] = k < cnt ? c0 + k : 0;
                ProfScope ps_(ctx, "gcr_dot_hist", 16. * n * (1 + cnt));
                // (distributed: the stopping test sees the GLOBAL ||r||^2 of the previous iteration here -- this iteration's is all-reduced
                // together with the inner products below --, identical on every rank; the output is unused once the solve has converged)
                GTRY(dot_hist(ctx, cnt, rg, dist ? A->n_global : n, Ar, Aps, stride, hl, std_conj, red + S_BNUM + 2 * c0, guard, tol2, pushB));
            }
            GCUDA(cudaGetLastError());
        }
        // (the last pass of a solve nobody watches leaves nothing to reduce: x is final, ||r||^2 is not read)
        if (dist && !(blind && final_iter) && !pushB.seq) GTRY(dist_allreduce_sum2(ctx, red + S_RR, scal + S_RR, 1 + 2 * lim));
        if (!blind) {
            GCUDA(cudaMemcpyAsync(slot.h, scal + S_BB, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            GCUDA(cudaEventRecord(slot.ev, ctx->stream));
        }
        if (!final_iter) {
            // p = z + p_corr, Ap = Ar + Ap_corr into the ring slot, next alpha's inner products          (GCR.h:259-266, 277-287)
            int next_iter = (iter % restart == 0) ? 0 : iter;
            int new_slot = next_iter % storage;
            GTRY(grow_ring(new_slot));
            int nchunks = std::max(1, (lim + GCR_CHUNK - 1) / GCR_CHUNK);
            memset(&pushA, 0, sizeof pushA); memset(&waitA, 0, sizeof waitA);
            if (fold && nchunks == 1) (void)p2p_allreduce_fold(ctx, 3, red + S_ANUM, scal + S_ANUM, &pushA, &waitA);
            for (int c = 0; c < nchunks; c++) {
                BetaList bl;
                const int cnt = std::max(0, std::min((int)GCR_CHUNK, lim - c * GCR_CHUNK));
                for (int k = 0; k < GCR_CHUNK; k++) {
                    bl.slot[k] = k < cnt ? c * GCR_CHUNK + k : 0; bl.num_index[k] = bl.slot[k];
                    bl.poff[k] = (p0_is_rhs && bl.slot[k] == 0) ? (int64_t)(rhs - ps) : (int64_t)bl.slot[k] * stride;
                }
                int first = (c == 0), last = (c == nchunks - 1);
                // the direction formed here feeds the LAST x update of a blind solve: A p is reduced in this pass and never read again
                const int keep_Ap = (lean && blind && g + 1 >= prm->max_iter && last) ? 0 : 1;
                ProfScope ps_(ctx, "gcr_update_p", 16. * n * (2 * cnt + (first ? 0 : 2) + 2 + (last ? 2 + (right ? 1 : 0) : 0) - (keep_Ap ? 0 : 1)));
                update_p(ctx, cnt, rg, zz, Ar, r, ps, Aps, stride, bl, new_slot, first, last, acc_p, acc_Ap, std_conj, bden_off, keep_Ap, scal, red + S_ANUM, guard, tol2, waitB, pushA);
            }
            GCUDA(cudaGetLastError());
            if (dist && !pushA.seq) GTRY(dist_allreduce_sum2(ctx, red + S_ANUM, scal + S_ANUM, 3));
            iter = next_iter;
            cur = new_slot;
        }
        if (blind) continue;   // the `while` below only counts iterations: rr / bb stay at 1
        {
            HostTimer ht(&ctx->host_sync_ms, &ctx->host_sync_calls);
            GCUDA(cudaEventSynchronize(slot.ev));
        }
        if (aliased) bb = slot.h[0];
        rr = slot.h[1];
        if (hist && g < hist_cap) hist[g] = sqrt(rr) / sqrt(bb);
        if (prm->verbose) printf("Step %d residual norm = %.10e\n", g, sqrt(rr) / sqrt(bb));   // GCR.h:271
    } while ((rr / bb) > prm->tol * prm->tol && g < prm->max_iter);                               // GCR.h:288
    if (prm->verbose) {
        if (g == prm->max_iter) printf("GCR did not converge after %d steps! Residual norm = %.10e\n", prm->max_iter, sqrt(rr) / sqrt(bb));
        else printf("GCR converged after %d steps. Residual norm=%.10e\n", g, sqrt(rr) / sqrt(bb));
    }
    if (iters_out) *iters_out = g;
    cleanup();
#undef GTRY
#undef GCUDA
    return MGCR_OK;
}

extern "C" int mgcr_gcr_solve(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right, const mgcr_c128* rhs,
                              mgcr_c128* x, double* hist, int hist_cap, int* iters) {
    ARG_CHECK(ctx && A && prm && rhs && x, "mgcr_gcr_solve: NULL argument");
    ARG_CHECK(!right || right->n_local == A->n_local, "x dimension does not match with Operator! (src/GCR.h:161)");
    ARG_CHECK(!left || left->n_local == A->n_local, "x dimension does not match with Operator! (src/GCR.h:161)");
    ARG_CHECK(!left || rhs != x, "mgcr_gcr_solve: a left-preconditioned solve cannot alias rhs and x");
    return gcr_solve_lr(ctx, A, prm, left, right, (const c128*)rhs, (c128*)x, hist, hist_cap, iters);
}

extern "C" int mgcr_gcr_solve_host(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right,
                                   const mgcr_c128* h_rhs, mgcr_c128* h_x, double* hist, int hist_cap, int* iters) {
    ARG_CHECK(ctx && A && prm && h_rhs && h_x, "mgcr_gcr_solve_host: NULL argument");
    const int64_t n = A->n_local;
    c128 *d_rhs = nullptr, *d_x = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)n, &d_rhs));
    int st = dev_alloc_t(ctx, (size_t)n, &d_x);
    if (st == MGCR_OK) {
        cudaError_t e = cudaMemcpyAsync(d_rhs, h_rhs, sizeof(c128) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, h_x, sizeof(c128) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { mgcr_set_error("solve_host upload: %s", cudaGetErrorString(e)); st = MGCR_ERR_CUDA; }
    }
    if (st == MGCR_OK) st = mgcr_gcr_solve(ctx, A, prm, left, right, (const mgcr_c128*)d_rhs, (mgcr_c128*)d_x, hist, hist_cap, iters);
    if (st == MGCR_OK) {
        cudaError_t e = cudaMemcpyAsync(h_x, d_x, sizeof(c128) * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { mgcr_set_error("solve_host download: %s", cudaGetErrorString(e)); st = MGCR_ERR_CUDA; }
    }
    dev_free(ctx, d_rhs); dev_free(ctx, d_x);
    return st;
}

// ----------------------------------------------------------------------------------------------------------
// GCR as an Operator (src/GCR.h:19, 62-68)
// ----------------------------------------------------------------------------------------------------------
int vec_init_rand_slab(mgcr_ctx* ctx, int seed, int64_t skip, int64_t n, c128* d_out);

int GcrOp::apply(const c128* x, c128* y) {
    ARG_CHECK(x != y, "operator apply: input and output alias");
    if (prm.zero_guess) {
        // (no memset: the solve is told that its start vector is zero)
    } else {
        if (!d_rand2) {
            MGCR_TRY(dev_alloc_t(ctx, (size_t)n_local, &d_rand2));
            int64_t skip = 0;
            if (A->distributed) {   // this rank's slice of the global init_rand(2) stream
                std::vector<int64_t> all;
                MGCR_TRY(dist_allgather_host_i64(ctx, n_local, all));
                for (int rnk = 0; rnk < ctx->rank; rnk++) skip += all[rnk];
            }
            MGCR_TRY(vec_init_rand_slab(ctx, 2, skip, n_local, d_rand2));
        }
        CUDA_TRY(cudaMemcpyAsync(y, d_rand2, sizeof(c128) * n_local, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    mgcr_gcr_param p = prm;
    p.verbose = prm.verbose;
    return gcr_solve_lr(ctx, A, &p, left, right, x, y, nullptr, 0, nullptr, prm.zero_guess != 0);
}

extern "C" int mgcr_gcr_op_create(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right, mgcr_op** out) {
    ARG_CHECK(ctx && A && prm && out, "mgcr_gcr_op_create: NULL argument");
    GcrOp* op = new GcrOp();
    op->kind = OP_GCR; op->ctx = ctx; op->A = A; op->prm = *prm; op->right = right; op->left = left;
    op->n_local = A->n_local; op->n_global = A->n_global; op->distributed = A->distributed;
    *out = op;
    return MGCR_OK;
}

extern "C" int mgcr_gcr_op_retarget(mgcr_op* gcr, mgcr_op* A) {
    ARG_CHECK(gcr && A && gcr->kind == OP_GCR, "mgcr_gcr_op_retarget: not a GCR operator");
    GcrOp* op = static_cast<GcrOp*>(gcr);
    if (op->n_local != A->n_local) { dev_free(op->ctx, op->d_rand2); op->d_rand2 = nullptr; }
    op->A = A; op->n_local = A->n_local; op->n_global = A->n_global; op->distributed = A->distributed;
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// Arnoldi::solve -- near-null vectors by inverse iteration (src/MG.h:90-122; tmp zero-initialised, Q7)
// ----------------------------------------------------------------------------------------------------------
int arnoldi(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* ep, int n_vec, c128* vecs) {
    const int64_t n = A->n_local;
    c128* b = vecs;
    int64_t skip = 0;
    const bool dist = A->distributed;
    if (dist) {
        std::vector<int64_t> all;
        MGCR_TRY(dist_allgather_host_i64(ctx, n, all));
        for (int rnk = 0; rnk < ctx->rank; rnk++) skip += all[rnk];
    }
    MGCR_TRY(vec_init_rand_slab(ctx, 9, skip, n, b));                     // MG.h:96
    mgcr_gcr_param p = *ep;
    p.verbose = 0;
    for (int i = 0; i < 10; i++) {                                        // MG.h:100-104
        MGCR_TRY(gcr_solve(ctx, A, &p, nullptr, b, b, nullptr, 0, nullptr));
        MGCR_TRY(vec_normalise(ctx, n, b, dist, A->n_global));
    }
    const int grid = stream_grid(ctx, n, 8);
    for (int c = 1; c < n_vec; c++) {                                     // MG.h:110-121
        c128* tmp = vecs + (int64_t)c * n;
        CUDA_TRY(cudaMemsetAsync(tmp, 0, sizeof(c128) * n, ctx->stream));
        MGCR_TRY(gcr_solve(ctx, A, &p, nullptr, vecs + (int64_t)(c - 1) * n, tmp, nullptr, 0, nullptr));
        for (int j = 0; j < c; j++) {
            const c128* ej = vecs + (int64_t)j * n;
            MGCR_TRY(vec_dot_dev(ctx, n, ej, tmp, ctx->d_scratch + 16, dist, A->n_global));
            if (n) {
                KLAUNCH(ctx, "vec_axpy", 48. * n, (k_axpy_devscal<<<grid, RED_THREADS, 0, ctx->stream>>>(n, ctx->d_scratch + 16, -1., ej, tmp, tmp)));
                CHECK_LAUNCH();
            }
        }
        MGCR_TRY(vec_normalise(ctx, n, tmp, dist, A->n_global));
    }
    return MGCR_OK;
}

extern "C" int mgcr_arnoldi(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* eigen, int n_vec, mgcr_c128* d_vecs) {
    ARG_CHECK(ctx && A && eigen && d_vecs && n_vec >= 1, "mgcr_arnoldi: bad argument");
    return arnoldi(ctx, A, eigen, n_vec, (c128*)d_vecs);
}
