// csrc/gcr_small.cu -- one whole GCR solve (reference: src/GCR.h:158-302) as ONE persistent
// cooperative kernel, for operators small enough that a host-driven iteration is bound by launch and read-back latency
// rather than by HBM: the coarse levels of the multigrid hierarchy (coarse solves and smoothers of MG.h:405-430).
//
// Every thread owns the elements i = t, t + T, ... (T = total threads) of every vector for the whole solve; the only
// vector read across threads is the operator's input (p at start, r afterwards).  An iteration is three phases
// separated by grid barriers:
//   P1  alpha ; x += alpha p ; r -= alpha Ap ; partial ||r||^2                                 (GCR.h:230-233)
//   P2  Ar = A r (rows owned) ; partial <Ar, Aps[k]> for the lim history vectors                (GCR.h:242, 257-258)
//   P3  beta_k ; p = r + sum -beta_k ps[k] ; Ap = Ar + sum -beta_k Aps[k] -> ring slot ; partial <r,Ap>, <Ap,Ap>
// Inner products: warp shuffle -> one partial per CTA -> after the barrier EVERY CTA sums the partials in the same fixed
// order, so all CTAs hold bit-identical scalars (uniform control flow, run-to-run deterministic) with no second barrier.
// The convergence test runs on the device: the host launches once and never synchronises unless it asked for the
// iteration count or the residual history.
#include <cooperative_groups.h>
#include <math.h>

#include <algorithm>

#include "kernels_blas.cuh"
#include "rows.cuh"

namespace cg = cooperative_groups;

enum { SG_THREADS = 256, SG_MAXH = 16, SG_NV = 2 * SG_MAXH };

struct SmallGcrArgs {
    int64_t n;
    int storage, restart, max_iter, std_conj;
    double tol2;
    const c128* rhs;
    c128* x;
    c128 *r, *Ar, *ps, *Aps;     // ps / Aps: `storage` ring slots of n
    double* partials;            // [2][gridDim][SG_NV]
    double* out;                 // [0] iterations, [1] ||r||^2, [2] ||rhs||^2
    double* hist;                // optional device array of hist_cap doubles
    int hist_cap;
    int x_zero;                  // the start vector is zero and x does not hold it: the first update writes x
};

// CTA partial of NV running sums -> partials[buf][block][k]
template <int NV>
__device__ __forceinline__ void cta_partials(const double (&v)[NV], int nv, double* __restrict__ dst) {
    __shared__ double sm[SG_THREADS / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        if (k < nv) {
            double s = warp_sum(v[k]);
            if (lane == 0) sm[warp][k] = s;
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nv) {
        double s = 0.;
#pragma unroll
        for (int w = 0; w < SG_THREADS / 32; w++) s += sm[w][threadIdx.x];
        dst[(size_t)blockIdx.x * SG_NV + threadIdx.x] = s;
    }
    __syncthreads();
}

// after the grid barrier: every CTA sums all CTA partials in one fixed order -> res[0..nv) in shared memory
__device__ __forceinline__ void sum_partials(const double* __restrict__ src, int nv, double* res /* shared [SG_NV] */) {
    __shared__ double sm2[SG_THREADS / 32][SG_NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < nv; k++) {
        double acc = 0.;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += SG_THREADS) acc += __ldcg(src + (size_t)b * SG_NV + k);
        acc = warp_sum(acc);
        if (lane == 0) sm2[warp][k] = acc;
    }
    __syncthreads();
    if ((int)threadIdx.x < nv) {
        double s = 0.;
#pragma unroll
        for (int w = 0; w < SG_THREADS / 32; w++) s += sm2[w][threadIdx.x];
        res[threadIdx.x] = s;
    }
    __syncthreads();
}

template <class Rows>
__global__ void __launch_bounds__(SG_THREADS) k_gcr_small(Rows M, SmallGcrArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double res[SG_NV];
    __shared__ double bden[SG_MAXH];
    __shared__ c128 beta[SG_MAXH];
    const int64_t T = (int64_t)gridDim.x * SG_THREADS;
    const int64_t t0 = (int64_t)blockIdx.x * SG_THREADS + threadIdx.x;
    const int64_t n = a.n;
    double* const part[2] = {a.partials, a.partials + (size_t)gridDim.x * SG_NV};
    int pb = 0;

    // r = rhs ; p = r                                                                      (GCR.h:189-190)
    c128* p0 = a.ps;
    for (int64_t i = t0; i < n; i += T) { c128 v = a.rhs[i]; a.r[i] = v; p0[i] = v; }
    grid.sync();
    // Ap = A p ; <r,Ap>, <Ap,Ap>, ||rhs||^2                                                 (GCR.h:191, 230)
    {
        double v[4] = {0., 0., 0., 0.};
        for (int64_t i = t0; i < n; i += T) {
            c128 av = M.apply_row(i, p0);
            a.Aps[i] = av;
            c128 rv = a.r[i];
            c128 d = a.std_conj ? cmulc(av, rv) : cmulc(rv, av);
            v[0] += d.x; v[1] += d.y;
            v[2] += av.x * av.x + av.y * av.y;
            v[3] += rv.x * rv.x + rv.y * rv.y;
        }
        cta_partials<4>(v, 4, part[pb]);
    }
    grid.sync();
    sum_partials(part[pb], 4, res);
    pb ^= 1;
    double anum_re = res[0], anum_im = res[1], aden = res[2];
    const double bb = res[3];
    double rr = bb;
    __syncthreads();
    if (t0 == 0 && a.hist && a.hist_cap > 0) a.hist[0] = sqrt(rr) / sqrt(bb);

    int iter = 0, g = 0, cur = 0;
    do {
        g++; iter++;
        // P1: alpha ; x += alpha p ; r -= alpha Ap ; ||r||^2
        const c128 alpha = cdivr(cmake(anum_re, anum_im), aden);
        if (threadIdx.x == 0) bden[cur] = aden;
        {
            const c128* p = a.ps + (int64_t)cur * n;
            const c128* Ap = a.Aps + (int64_t)cur * n;
            double v[1] = {0.};
            for (int64_t i = t0; i < n; i += T) {
                c128 xv = cadd((a.x_zero && g == 1) ? cmake(0., 0.) : a.x[i], cmul(alpha, p[i]));
                c128 rv = csub(a.r[i], cmul(alpha, Ap[i]));
                a.x[i] = xv; a.r[i] = rv;
                v[0] += rv.x * rv.x + rv.y * rv.y;
            }
            cta_partials<1>(v, 1, part[pb]);
        }
        grid.sync();
        sum_partials(part[pb], 1, res);
        pb ^= 1;
        rr = res[0];
        __syncthreads();
        if (t0 == 0 && a.hist && g < a.hist_cap) a.hist[g] = sqrt(rr) / sqrt(bb);
        if (!((rr / bb) > a.tol2 && g < a.max_iter)) break;                                  // GCR.h:288 (uniform: res is identical in all CTAs)
        // P2: Ar = A r ; <Ar, Aps[k]>
        const int lim = min(a.storage, iter);                                               // GCR.h:251
        {
            double v[SG_NV];
#pragma unroll
            for (int k = 0; k < SG_NV; k++) v[k] = 0.;
            for (int64_t i = t0; i < n; i += T) {
                c128 av = M.apply_row(i, a.r);
                a.Ar[i] = av;
#pragma unroll
                for (int k = 0; k < SG_MAXH; k++) {
                    if (k < lim) {
                        c128 h = a.Aps[(int64_t)k * n + i];
                        c128 d = a.std_conj ? cmulc(h, av) : cmulc(av, h);
                        v[2 * k] += d.x; v[2 * k + 1] += d.y;
                    }
                }
            }
            cta_partials<SG_NV>(v, 2 * lim, part[pb]);
        }
        grid.sync();
        sum_partials(part[pb], 2 * lim, res);
        pb ^= 1;
        if ((int)threadIdx.x < lim) beta[threadIdx.x] = cdivr(cmake(res[2 * threadIdx.x], res[2 * threadIdx.x + 1]), bden[threadIdx.x]);
        __syncthreads();
        // P3: p, Ap into the ring slot ; next alpha's inner products                         (GCR.h:259-266, 277-287)
        const int next_iter = (iter % a.restart == 0) ? 0 : iter;
        const int new_slot = next_iter % a.storage;
        {
            c128* pn = a.ps + (int64_t)new_slot * n;
            c128* Apn = a.Aps + (int64_t)new_slot * n;
            double v[3] = {0., 0., 0.};
            for (int64_t i = t0; i < n; i += T) {
                c128 pc = cmake(0., 0.), Apc = cmake(0., 0.);
                for (int k = 0; k < lim; k++) {
                    const c128 bk = beta[k];
                    pc = csub(pc, cmul(bk, a.ps[(int64_t)k * n + i]));
                    Apc = csub(Apc, cmul(bk, a.Aps[(int64_t)k * n + i]));
                }
                const c128 rv = a.r[i];
                pc = cadd(rv, pc);
                Apc = cadd(a.Ar[i], Apc);
                pn[i] = pc; Apn[i] = Apc;
                c128 d = a.std_conj ? cmulc(Apc, rv) : cmulc(rv, Apc);
                v[0] += d.x; v[1] += d.y;
                v[2] += Apc.x * Apc.x + Apc.y * Apc.y;
            }
            __syncthreads();   // res / bden fully consumed before cta_partials reuses shared memory
            cta_partials<3>(v, 3, part[pb]);
        }
        grid.sync();
        sum_partials(part[pb], 3, res);
        pb ^= 1;
        anum_re = res[0]; anum_im = res[1]; aden = res[2];
        __syncthreads();
        iter = next_iter; cur = new_slot;
    } while (true);
    if (t0 == 0) { a.out[0] = (double)g; a.out[1] = rr; a.out[2] = bb; }
}

// ----------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------
template <class Rows>
static int launch_small(mgcr_ctx* ctx, const Rows& rows, SmallGcrArgs& a, int* grid_out) {
    static thread_local int max_blocks_per_sm = -1;
    if (max_blocks_per_sm < 0) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, k_gcr_small<Rows>, SG_THREADS, 0));
        if (max_blocks_per_sm < 1) { mgcr_set_error("small GCR kernel does not fit on an SM"); return MGCR_ERR_CUDA; }
    }
    int64_t want = 0;
    static const int per_sm = getenv("MGCR_SMALL_GRID_PER_SM") ? atoi(getenv("MGCR_SMALL_GRID_PER_SM")) : 2;   // experiment knob
    int64_t cap = (int64_t)ctx->num_sms * std::min(max_blocks_per_sm, per_sm);
    static const int min_rows = getenv("MGCR_SMALL_ROWS_PER_CTA") ? atoi(getenv("MGCR_SMALL_ROWS_PER_CTA")) : SG_THREADS;
    want = (a.n + min_rows - 1) / min_rows;
    int grid = (int)std::max<int64_t>(1, std::min(want, cap));
    *grid_out = grid;
    return MGCR_OK;
}

// Returns MGCR_OK and sets *handled = 1 when the solve was enqueued as one persistent kernel; *handled = 0 when the
// caller must run the host-driven loop (operator has no row access, solve too large, history too long, verbose...).
int gcr_solve_small(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, const c128* rhs, c128* x, double* hist, int hist_cap,
                    int* iters_out, int storage, int restart, int* handled, bool x_zero) {
    *handled = 0;
    const int64_t n = A->n_local;
    if (n == 0 || n > ctx->small_gcr_rows || storage > SG_MAXH || prm->verbose || rhs == x || A->distributed) return MGCR_OK;
    int st = MGCR_OK;
    c128* work = nullptr;
    double* scal = nullptr;
    bool ok = with_rows(A, [&](auto rows) -> int {
        typedef decltype(rows) Rows;
        SmallGcrArgs a;
        a.n = n; a.storage = storage; a.restart = restart; a.max_iter = prm->max_iter; a.std_conj = prm->std_conj;
        a.tol2 = prm->tol * prm->tol;
        a.rhs = rhs; a.x = x; a.x_zero = x_zero ? 1 : 0;
        int grid = 1;
        MGCR_TRY(launch_small(ctx, rows, a, &grid));
        MGCR_TRY(dev_alloc_t(ctx, (size_t)n * (2 + 2 * (size_t)storage), &work));
        const size_t nscal = 2 * (size_t)grid * SG_NV + 4 + (size_t)std::max(hist_cap, 0);
        MGCR_TRY(dev_alloc_t(ctx, nscal, &scal));
        a.r = work; a.Ar = work + n; a.ps = work + 2 * n; a.Aps = work + (2 + (int64_t)storage) * n;
        a.partials = scal; a.out = scal + 2 * (size_t)grid * SG_NV;
        a.hist = (hist && hist_cap > 0) ? a.out + 4 : nullptr;
        a.hist_cap = hist_cap;
        void* params[] = {(void*)&rows, (void*)&a};
        {
            ProfScope ps_(ctx, "gcr_small", 0.);
            CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_gcr_small<Rows>, dim3(grid), dim3(SG_THREADS), params, 0, ctx->stream));
        }
        if (iters_out || a.hist) {
            const int nread = 4 + (a.hist ? hist_cap : 0);
            std::vector<double> h((size_t)nread);
            CUDA_TRY(cudaMemcpyAsync(h.data(), a.out, sizeof(double) * nread, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            const int g = (int)h[0];
            if (iters_out) *iters_out = g;
            if (a.hist) for (int q = 0; q <= g && q < hist_cap; q++) hist[q] = h[4 + q];
        }
        return MGCR_OK;
    }, &st);
    dev_free(ctx, work); dev_free(ctx, scal);
    if (!ok) return MGCR_OK;
    if (st == MGCR_OK) *handled = 1;
    return st;
}
