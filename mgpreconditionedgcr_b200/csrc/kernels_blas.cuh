// csrc/kernels_blas.cuh -- streaming complex128 vector kernels: Field BLAS-1 (src/Fields.h) and the fused kernels of
// the GCR iteration (src/GCR.h:222-288).  All are HBM-bound: one 128-bit access per element and vector, grid-stride
// over a grid sized in multiples of the SM count, reductions through slab_partial / grid_finish (common.cuh: fixed order,
// independent of the number of GPUs).
#pragma once
#include "common.cuh"

#define GRID_STRIDE(i, n) for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)
// the same over virtual slab vs of the vector (RedGeom, common.cuh): this CTA (gridDim.x = G) strides over the elements
// [vs L, (vs + 1) L) with stride G * blockDim.x; nvs = 1, G = gridDim.x, L = n is exactly GRID_STRIDE
#define SLAB_STRIDE(i, rg, vs)                                                                                                              \
    for (int64_t i = (int64_t)(vs) * (rg).L + (int64_t)blockIdx.x * blockDim.x + threadIdx.x, slab_end__ = (int64_t)((vs) + 1) * (rg).L;   \
         i < slab_end__; i += (int64_t)(rg).G * blockDim.x)

static __global__ void __launch_bounds__(RED_THREADS) k_fill(int64_t n, c128 v, c128* __restrict__ out) {
    PDL_ENTRY();
    GRID_STRIDE(i, n) st_stream(out + i, v);
}

// out = a + s*b   (Field::operator+ / operator- / += / -=; the product is formed first, as `s * field[i]`)
static __global__ void __launch_bounds__(RED_THREADS) k_axpy(int64_t n, c128 s, const c128* b, const c128* a, c128* out) {
    PDL_ENTRY();
    GRID_STRIDE(i, n) {
        c128 t = cmul(s, ld_plain(b + i));
        st_stream(out + i, cadd(ld_plain(a + i), t));
    }
}

static __global__ void __launch_bounds__(RED_THREADS) k_scale(int64_t n, c128 s, const c128* a, c128* out) {
    PDL_ENTRY();
    GRID_STRIDE(i, n) st_stream(out + i, cmul(s, ld_plain(a + i)));
}

// a *= 1./sqrt(*nrm2)   (Field::normalise: `field[i] *= 1./norm`, complex *= real-as-complex... the reference
// multiplies by the double 1./norm, i.e. component-wise)
static __global__ void __launch_bounds__(RED_THREADS) k_scale_inv_sqrt(int64_t n, const double* __restrict__ nrm2, c128* a) {
    PDL_ENTRY();
    const double s = 1. / sqrt(*nrm2);
    GRID_STRIDE(i, n) {
        c128 v = ld_plain(a + i);
        st_stream(a + i, cmake(v.x * s, v.y * s));
    }
}

static __global__ void __launch_bounds__(RED_THREADS) k_dot(RedGeom rg, const c128* __restrict__ a, const c128* __restrict__ b,
                                                     double* partials, unsigned int* ticket, double* out) {
    PDL_ENTRY();
    __shared__ SlabSums<2> sums;
    for (int vs = 0; vs < rg.nvs; vs++) {
        double v[2] = {0., 0.};
        SLAB_STRIDE(i, rg, vs) {
            c128 t = cmulc(ld_stream(a + i), ld_stream(b + i));
            v[0] += t.x; v[1] += t.y;
        }
        slab_partial<2>(v, sums, vs);
    }
    grid_finish<2>(sums, partials, ticket, out, rg);
}

static __global__ void __launch_bounds__(RED_THREADS) k_norm2(RedGeom rg, const c128* __restrict__ a, double* partials,
                                                       unsigned int* ticket, double* out) {
    PDL_ENTRY();
    __shared__ SlabSums<1> sums;
    for (int vs = 0; vs < rg.nvs; vs++) {
        double v[1] = {0.};
        SLAB_STRIDE(i, rg, vs) {
            c128 t = ld_stream(a + i);
            v[0] += t.x * t.x + t.y * t.y;
        }
        slab_partial<1>(v, sums, vs);
    }
    grid_finish<1>(sums, partials, ticket, out, rg);
}

// gamma5 permutation along an axis of extent axis_dim with `inner` elements below it (src/Fields.h:310-339):
// out[.., s', ..] = in[.., s, ..] with s' = s^2 for s < 4 (0<->2, 1<->3), identity above.
static __global__ void __launch_bounds__(RED_THREADS) k_gamma5(int64_t n, int64_t inner, int64_t axis_dim, const c128* __restrict__ in,
                                                        c128* __restrict__ out) {
    PDL_ENTRY();
    GRID_STRIDE(j, n) {
        int64_t s = (j / inner) % axis_dim;
        int64_t src = s < 4 ? (s ^ 2) : s;
        st_stream(out + j, ld_plain(in + j + (src - s) * inner));
    }
}

// ----------------------------------------------------------------------------------------------------------
// GCR iteration kernels.  Device-resident scalar block (doubles), one per active solve:
//   S_ANUM (2)  <r,Ap>            (or <Ap,r> with std_conj)           -> alpha numerator   GCR.h:230
//   S_ADEN (1)  <Ap,Ap>                                                -> alpha denominator
//   S_BB   (1)  ||rhs||^2
//   S_RR   (1)  ||r||^2                                                 GCR.h:288
//   S_BNUM (2*storage)  <Ar,Aps[i]>  (or <Aps[i],Ar>)                   GCR.h:258
//   S_BDEN (storage)    ||Aps[i]||^2 cached per ring slot (= the alpha denominator of the iteration that made it)
// [S_ANUM, S_BB) is all-reduced after update_p, [S_RR, S_BNUM + 2*lim) after dot_hist (two all-reduces / iteration).
// ----------------------------------------------------------------------------------------------------------
enum { GCR_CHUNK = 16 };
enum { S_ANUM = 0, S_ADEN = 2, S_BB = 3, S_RR = 4, S_BNUM = 5 };   // S_BDEN = S_BNUM + 2*storage

// Device-side stopping test (GCR.h:288) for solves the host does not watch (short smoother solves inside the multigrid
// cycle): with guard != NULL every kernel of an iteration first checks ||r||^2 <= tol^2 ||rhs||^2 on the scalars the
// previous kernels left and returns at once when the solve has converged, so x and r stop changing exactly where the
// reference's loop would have stopped while the host keeps enqueueing without a read-back.
__device__ __forceinline__ bool gcr_converged(const double* guard, double tol2) {
    return guard != nullptr && guard[S_RR] <= tol2 * guard[S_BB];
}

// init: <r,Ap>, <Ap,Ap>, ||rhs||^2 (r = rhs at start) in one pass -> S_ANUM(2), S_ADEN, S_BB, S_RR (= ||rhs||^2)
// The same pass can also make the solver's working copies r = rhs and p = r (GCR.h:189-190; r_out / p_out != NULL): the host asks for
// them only where somebody needs a COPY -- r is otherwise first written by the first x / r update, which reads rhs (r_in), and an
// unpreconditioned blind solve reads p0 from rhs wherever ring slot 0 is addressed (csrc/gcr.cu).
static __global__ void __launch_bounds__(RED_THREADS) k_gcr_init(RedGeom rg, const c128* __restrict__ r, const c128* __restrict__ Ap,
                                                          int std_conj, c128* __restrict__ r_out, c128* __restrict__ p_out,
                                                          double* partials, unsigned int* ticket, double* out5,
                                                          const __grid_constant__ ArPush push) {
    PDL_ENTRY();
    __shared__ SlabSums<5> sums;
    for (int vs = 0; vs < rg.nvs; vs++) {
        double v[5] = {0., 0., 0., 0., 0.};
        SLAB_STRIDE(i, rg, vs) {
            c128 rv = ld_stream(r + i), av = ld_stream(Ap + i);
            if (r_out) st_stream(r_out + i, rv);
            if (p_out) st_stream(p_out + i, rv);
            c128 t = cmulc(rv, av);
            v[0] += t.x; v[1] += t.y;
            v[2] += av.x * av.x + av.y * av.y;
            v[3] += rv.x * rv.x + rv.y * rv.y;
        }
        if (std_conj) v[1] = -v[1];   // <Ap,r> = conj(<r,Ap>), exactly, term by term
        v[4] = v[3];
        slab_partial<5>(v, sums, vs);
    }
    grid_finish<5>(sums, partials, ticket, out5, rg, 5, &push);
}

// x += alpha p ; r -= alpha Ap ; ||r||^2 -> scal[S_RR]      (GCR.h:230-233)
// FOLD: the variant that receives its scalars through a folded all-reduce (aw.seq != 0), kept apart from the plain one because
// the extra prologue changed the compiler's schedule of the streaming loop (same box: 602 us per launch against 585)
template <bool FOLD>
static __global__ void __launch_bounds__(RED_THREADS, 5) k_gcr_update_xr(RedGeom rg, const c128* __restrict__ p, const c128* __restrict__ Ap,
                                                               c128* x, const c128* r_in, c128* r, double* scal, double* rr_out, int bden_slot, int x_zero,
                                                               double* partials, unsigned int* ticket, const double* guard, double tol2,
                                                               const __grid_constant__ ArWait aw) {
    // r_in: where the residual is read from -- r itself, or the right-hand side in the first iteration (r = rhs then, GCR.h:189: the
    // solver's working copy is first written HERE instead of by a copy in the init kernel, 16 bytes per element less per solve)
    PDL_ENTRY();
    __shared__ SlabSums<1> sums;
    __shared__ double arv[AR_FOLD_MAX];
    // Folded all-reduce (aw.seq != 0, blind distributed solves): <r,Ap>, <Ap,Ap> (and, before the first iteration, ||rhs||^2 twice)
    // arrive through the peer slots.  The stopping test on the scalars earlier kernels left decides FIRST whether anything was sent
    // at all (a producer that saw the solve converged sent nothing); before the first iteration nothing has been left yet and the
    // test runs on the arriving values instead.
    const bool first_batch = FOLD && aw.n == 5;
    bool converged = !first_batch && gcr_converged(guard, tol2);
    if constexpr (FOLD) {
        if (!converged) {
            ar_wait(aw, arv);                                 // -> scal[S_ANUM .. S_ADEN] (first batch: .. S_RR)
            if (first_batch) converged = arv[S_RR] <= tol2 * arv[S_BB];
        }
    }
    if (converged) {
        // x_zero: x has never been written (the caller skipped the memset).  A solve that counts as converged before its first
        // iteration (a zero right-hand side) must still return the zero vector
        if (x_zero)
            for (int vs = 0; vs < rg.nvs; vs++) SLAB_STRIDE(i, rg, vs) st_stream(x + i, cmake(0., 0.));
        return;
    }
    const double aden = FOLD ? arv[S_ADEN] : scal[S_ADEN];
    const c128 alpha = FOLD ? cdivr(cmake(arv[S_ANUM], arv[S_ANUM + 1]), aden) : cdivr(cmake(scal[S_ANUM], scal[S_ANUM + 1]), aden);
    // ||Aps[cur]||^2 never changes while the slot lives: cache it for the beta denominators (GCR.h:258 recomputes it)
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[bden_slot] = aden;
    // two elements per trip: 8 independent 128-bit loads in flight per thread
    const int64_t T = (int64_t)rg.G * blockDim.x;
    for (int vs = 0; vs < rg.nvs; vs++) {
        double v[1] = {0.};
        const int64_t n = (int64_t)(vs + 1) * rg.L;                        // end of this virtual slab
        int64_t i = (int64_t)vs * rg.L + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + T < n; i += 2 * T) {
            const c128 p0 = ld_stream(p + i), a0 = ld_stream(Ap + i), p1 = ld_stream(p + i + T), a1 = ld_stream(Ap + i + T);
            c128 x0 = cmake(0., 0.), x1 = cmake(0., 0.);
            if (!x_zero) { x0 = ld_plain(x + i); x1 = ld_plain(x + i + T); }
            c128 r0 = ld_plain(r_in + i), r1 = ld_plain(r_in + i + T);
            x0 = cadd(x0, cmul(alpha, p0)); r0 = csub(r0, cmul(alpha, a0));
            x1 = cadd(x1, cmul(alpha, p1)); r1 = csub(r1, cmul(alpha, a1));
            st_stream(x + i, x0); st_stream(r + i, r0); st_stream(x + i + T, x1); st_stream(r + i + T, r1);
            v[0] += r0.x * r0.x + r0.y * r0.y;
            v[0] += r1.x * r1.x + r1.y * r1.y;
        }
        for (; i < n; i += T) {
            c128 pv = ld_stream(p + i), av = ld_stream(Ap + i);
            c128 xv = x_zero ? cmake(0., 0.) : ld_plain(x + i), rv = ld_plain(r_in + i);
            xv = cadd(xv, cmul(alpha, pv));
            rv = csub(rv, cmul(alpha, av));
            st_stream(x + i, xv);
            st_stream(r + i, rv);
            v[0] += rv.x * rv.x + rv.y * rv.y;
        }
        slab_partial<1>(v, sums, vs);
    }
    grid_finish<1>(sums, partials, ticket, rr_out, rg);   // scal + S_RR, or this rank's partial block when the solve is distributed
}

// The LAST x update of a solve nobody watches (blind, csrc/gcr.cu): x += alpha p and nothing else -- the residual is not read
// again, its norm is not looked at, so neither Ap nor r is touched (48 instead of 96 bytes per element; 32 when x starts at zero).
template <bool FOLD>
static __global__ void __launch_bounds__(RED_THREADS) k_gcr_update_x(int64_t n, const c128* __restrict__ p, c128* x, const double* scal, int x_zero,
                                                              const double* guard, double tol2, const __grid_constant__ ArWait aw) {
    PDL_ENTRY();
    __shared__ double arv[AR_FOLD_MAX];
    const bool first_batch = FOLD && aw.n == 5;
    bool converged = !first_batch && gcr_converged(guard, tol2);
    if constexpr (FOLD) {
        if (!converged) {
            ar_wait(aw, arv);
            if (first_batch) converged = arv[S_RR] <= tol2 * arv[S_BB];
        }
    }
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    if (converged) {
        if (x_zero)
            for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += T) st_stream(x + i, cmake(0., 0.));
        return;
    }
    const double aden = FOLD ? arv[S_ADEN] : scal[S_ADEN];
    const c128 alpha = FOLD ? cdivr(cmake(arv[S_ANUM], arv[S_ANUM + 1]), aden) : cdivr(cmake(scal[S_ANUM], scal[S_ANUM + 1]), aden);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * T < n; i += 4 * T) {
        c128 pv[4], xv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { pv[u] = ld_stream(p + i + u * T); xv[u] = x_zero ? cmake(0., 0.) : ld_plain(x + i + u * T); }
#pragma unroll
        for (int u = 0; u < 4; u++) st_stream(x + i + u * T, cadd(xv[u], cmul(alpha, pv[u])));
    }
    for (; i < n; i += T) {
        const c128 xv = x_zero ? cmake(0., 0.) : ld_plain(x + i);
        st_stream(x + i, cadd(xv, cmul(alpha, ld_stream(p + i))));
    }
}

// batched <Ar, Aps[slot]> for nh (<= GCR_CHUNK) history vectors in one pass over Ar      (GCR.h:257-258)
// The CTA is split into KS thread groups; group g owns the history vectors g, g+KS, ... (at most NK of them) for the
// element range of the whole CTA, so a thread carries 2*NK accumulators and U*(1+NK) independent 128-bit loads instead
// of 2*nh and 1+nh (long histories otherwise end up with few loads in flight; Ar is read KS times, the repeats hit L2).
struct HistList { int slot[GCR_CHUNK]; };

// one thread group's pass: CNT history vectors (compile time: no predicated loads), U elements per trip
template <int CNT, int U, int NK>
__device__ __forceinline__ void dot_hist_group(int64_t n, int64_t i0, int64_t T, const c128* __restrict__ Ar, const c128* const (&hp)[NK],
                                               double (&v)[2 * NK]) {
    for (; i0 + (U - 1) * T < n; i0 += T * U) {
        c128 a[U], h[U][CNT];
#pragma unroll
        for (int u = 0; u < U; u++) {
            a[u] = ld_stream(Ar + i0 + u * T);
#pragma unroll
            for (int k = 0; k < CNT; k++) h[u][k] = ld_stream(hp[k] + i0 + u * T);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int k = 0; k < CNT; k++) {
                c128 t = cmulc(a[u], h[u][k]);
                v[2 * k] += t.x; v[2 * k + 1] += t.y;
            }
        }
    }
    for (; i0 < n; i0 += T) {                        // tail
        const c128 a = ld_stream(Ar + i0);
#pragma unroll
        for (int k = 0; k < CNT; k++) {
            c128 t = cmulc(a, ld_stream(hp[k] + i0));
            v[2 * k] += t.x; v[2 * k + 1] += t.y;
        }
    }
}

template <int NK, int KS>
static __global__ void __launch_bounds__(RED_THREADS) k_gcr_dot_hist(RedGeom rg, const c128* __restrict__ Ar, const c128* __restrict__ Aps,
                                                              int64_t stride, HistList hl, int nh, int std_conj, double* out /* 2*nh */,
                                                              double* partials, unsigned int* ticket, const double* guard, double tol2,
                                                              const __grid_constant__ ArPush push) {
    PDL_ENTRY();
    if (gcr_converged(guard, tol2)) return;
    constexpr int GT = RED_THREADS / KS;     // threads per group
    constexpr int GW = GT / 32;              // warps per group
    const int g = threadIdx.x / GT, tl = threadIdx.x % GT;
    const int cnt = (nh - g + KS - 1) / KS;  // history vectors of this group: g, g+KS, ...  (warp-uniform)
    const c128* hp[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) hp[k] = Aps + (int64_t)hl.slot[k < cnt ? g + k * KS : 0] * stride;
    const int64_t T = (int64_t)rg.G * GT;
    // warp sums per virtual slab (no block barrier between slabs); value q = 2*kk + c of history vector kk = g + k*KS lives in
    // the warps of group g
    __shared__ double sm[RED_VSLABS][RED_THREADS / 32][2 * NK];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nv = 2 * nh;
    for (int vs = 0; vs < rg.nvs; vs++) {
        double v[2 * NK];
#pragma unroll
        for (int k = 0; k < 2 * NK; k++) v[k] = 0.;
        const int64_t n = (int64_t)(vs + 1) * rg.L;                        // end of this virtual slab
        const int64_t i0 = (int64_t)vs * rg.L + (int64_t)blockIdx.x * GT + tl;
        switch (cnt) {
            case 1: dot_hist_group<1, 4, NK>(n, i0, T, Ar, hp, v); break;
            case 2: if constexpr (NK >= 2) dot_hist_group<2, 2, NK>(n, i0, T, Ar, hp, v); break;
            case 3: if constexpr (NK >= 3) dot_hist_group<3, 2, NK>(n, i0, T, Ar, hp, v); break;
            case 4: if constexpr (NK >= 4) dot_hist_group<4, 1, NK>(n, i0, T, Ar, hp, v); break;
            default: break;
        }
        // <h, a> = conj(<a, h>) term by term and exactly (the products commute, the difference changes sign): the textbook
        // convention only flips the sign of the imaginary sums
        if (std_conj) {
#pragma unroll
            for (int k = 0; k < NK; k++) v[2 * k + 1] = -v[2 * k + 1];
        }
#pragma unroll
        for (int k = 0; k < 2 * NK; k++) {
            double s = warp_sum(v[k]);
            if (lane == 0) sm[vs][warp][k] = s;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < rg.nvs * nv; t += RED_THREADS) {
        const int vs = t / nv, q = t - vs * nv;
        const int kk = q >> 1, c = q & 1;
        const int gg = kk % KS, k = kk / KS;
        double s = 0.;
#pragma unroll
        for (int w = 0; w < GW; w++) s += sm[vs][gg * GW + w][2 * k + c];
        partials[(size_t)(vs * rg.G + blockIdx.x) * MAX_RED_VALUES + q] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA: every (slab, value) pair over all CTAs in a fixed order (combine_partials); sm has been read by everybody
    combine_partials(partials, MAX_RED_VALUES, nv, rg, &sm[0][0][0], out, nv);
    if (push.seq) {   // folded all-reduce of ||r||^2 (left by the x / r update) and these inner products
        __syncthreads();
        ar_push(push);
    }
}

// ----------------------------------------------------------------------------------------------------------
// The same inner products with the operands staged through shared memory by TMA (1-D bulk copies completing on an
// mbarrier): a ring of S stages, each holding a tile of TE elements of Ar and of the NH history vectors.  One elected
// thread keeps S tiles in flight per SM whatever the compiler does with the consumer loop, every thread needs only its
// 2*NH accumulators, and the work is balanced for every NH.  One CTA per SM (the ring takes most of the shared memory).
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

enum { DOT_TMA_MAX_STAGES = 16 };

template <int NH>
static __global__ void __launch_bounds__(RED_THREADS, 1) k_gcr_dot_hist_tma(RedGeom rg, const c128* __restrict__ Ar, const c128* __restrict__ Aps,
                                                                     int64_t stride, HistList hl, int std_conj, int ept, int stages,
                                                                     double* out /* 2*NH */, double* partials, unsigned int* ticket,
                                                                     const double* guard, double tol2, const __grid_constant__ ArPush push) {
    PDL_ENTRY();
    __shared__ SlabSums<2 * NH> sums;
    if (gcr_converged(guard, tol2)) return;
    extern __shared__ __align__(128) unsigned char dot_smem[];
    __shared__ __align__(8) uint64_t full[DOT_TMA_MAX_STAGES];
    const int te = RED_THREADS * ept;                         // elements per tile
    const size_t stage_elems = (size_t)(1 + NH) * te;
    c128* ring = (c128*)dot_smem;
    // The tiles of a virtual slab are dealt round-robin over the G CTAs; this CTA works through its tiles of slab 0, then of
    // slab 1, ... as ONE stream of work items, so the ring stays full across the slab boundaries.
    const int64_t tiles = (rg.L + te - 1) / te;                              // per slab
    const int64_t tpc = blockIdx.x < tiles ? (tiles - blockIdx.x + rg.G - 1) / rg.G : 0;   // tiles of one slab that are this CTA's
    const int64_t items = tpc * rg.nvs;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // (slab, tile-of-slab) of a work item advance by counting: no 64-bit divisions in the loop
    int64_t iss_vs = 0, iss_t = 0;                            // next item to issue (thread 0)
    auto issue = [&](int s) {                                 // thread 0 only
        const int64_t tile = blockIdx.x + iss_t * rg.G;
        const int64_t e0 = iss_vs * rg.L + tile * te;
        const uint32_t cnt = (uint32_t)min((int64_t)te, (iss_vs + 1) * rg.L - e0);
        c128* dst = ring + (size_t)s * stage_elems;
        mbar_expect_tx(&full[s], cnt * 16u * (1 + NH));
        tma_load_1d(dst, Ar + e0, cnt * 16u, &full[s]);
#pragma unroll
        for (int k = 0; k < NH; k++) tma_load_1d(dst + (size_t)(1 + k) * te, Aps + (int64_t)hl.slot[k] * stride + e0, cnt * 16u, &full[s]);
        if (++iss_t == tpc) { iss_t = 0; iss_vs++; }
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < stages; s++)
            if (s < items) issue(s);
    double v[2 * NH];
#pragma unroll
    for (int k = 0; k < 2 * NH; k++) v[k] = 0.;
    int s = 0; uint32_t parity = 0;
    int vs = 0; int64_t t = 0;
    for (int64_t item = 0; item < items; item++) {
        mbar_wait(&full[s], parity);
        const c128* st = ring + (size_t)s * stage_elems;
        const int64_t left = rg.L - (blockIdx.x + t * rg.G) * te;
        for (int q = 0; q < ept; q++) {
            const int e = q * RED_THREADS + threadIdx.x;
            if (e < left) {
                const c128 a = st[e];
#pragma unroll
                for (int k = 0; k < NH; k++) {
                    c128 tt = cmulc(a, st[(size_t)(1 + k) * te + e]);
                    v[2 * k] += tt.x; v[2 * k + 1] += tt.y;
                }
            }
        }
        __syncthreads();                                      // every thread is done with stage s
        if (threadIdx.x == 0 && item + stages < items) issue(s);
        if (++s == stages) { s = 0; parity ^= 1; }
        if (++t == tpc) {                                     // this CTA's last tile of slab vs: its warp sums, fresh accumulators
            if (std_conj) {
#pragma unroll
                for (int k = 0; k < NH; k++) v[2 * k + 1] = -v[2 * k + 1];
            }
            slab_partial<2 * NH>(v, sums, vs);
#pragma unroll
            for (int k = 0; k < 2 * NH; k++) v[k] = 0.;
            t = 0; vs++;
        }
    }
    if (tpc == 0)                                             // (more CTAs than tiles cannot happen: G <= tiles; kept for safety)
        for (int vs = 0; vs < rg.nvs; vs++) slab_partial<2 * NH>(v, sums, vs);
    grid_finish<2 * NH>(sums, partials, ticket, out, rg, 2 * NH, &push);
}

// p_new = z + sum_i(-beta_i ps[i]) ; Ap_new = Ar + sum_i(-beta_i Aps[i]) written into ring slot `cur`, with the next
// alpha's <r,Ap_new>, <Ap_new,Ap_new> reduced in the same pass        (GCR.h:251-266, 286-287, 230)
// History slots are visited in the reference's order i = 0..lim-1.  More than GCR_CHUNK history vectors are
// processed in several passes: all but the last accumulate pc / Apc in two scratch vectors (the ring slot `cur` may
// itself still be unread history), the last one adds z / Ar, writes the slot and reduces.
// NH is the exact number of history vectors of this pass; MINB (resident CTAs per SM) bounds the registers so that
// long histories do not collapse the occupancy.
// poff[k]: element offset of history direction k from `ps` -- slot * stride, or the distance to the right-hand side when slot 0 IS
// the right-hand side (unpreconditioned blind solves: p0 = rhs is never copied into the ring, csrc/gcr.cu)
struct BetaList { int slot[GCR_CHUNK]; int num_index[GCR_CHUNK]; int64_t poff[GCR_CHUNK]; };

template <int NH, int MINB>
static __global__ void __launch_bounds__(RED_THREADS, MINB) k_gcr_update_p(RedGeom rg, const c128* z, const c128* Ar, const c128* r, c128* ps,
                                                                    c128* Aps, int64_t stride, BetaList bl, int cur, int first, int last,
                                                                    c128* acc_p, c128* acc_Ap, int std_conj, int bden_off, int keep_Ap, const double* scal,
                                                                    double* anum_out, double* partials, unsigned int* ticket, const double* guard,
                                                                    double tol2, const __grid_constant__ ArWait aw, const __grid_constant__ ArPush push) {
    PDL_ENTRY();
    __shared__ SlabSums<3> sums;
    __shared__ double arv[AR_FOLD_MAX];
    if (gcr_converged(guard, tol2)) return;   // (folded: on the previous iteration's ||r||^2, as the inner-product kernel did -- it sent nothing)
    constexpr int NHS = NH > 0 ? NH : 1;
    __shared__ c128 beta[NHS];
    if (aw.seq) {
        // folded all-reduce: this iteration's ||r||^2 and the numerators <Ar, Aps[i]> arrive through the peer slots
        // (-> scal[S_RR], scal[S_BNUM ..]); the stopping test of this iteration happens here, identically on every rank
        ar_wait(aw, arv);
        if (arv[0] <= tol2 * guard[S_BB]) return;
    }
    if ((int)threadIdx.x < NH) {
        int q = bl.num_index[threadIdx.x];
        const c128 num = aw.seq ? cmake(arv[S_BNUM - S_RR + 2 * q], arv[S_BNUM - S_RR + 2 * q + 1]) : cmake(scal[S_BNUM + 2 * q], scal[S_BNUM + 2 * q + 1]);
        beta[threadIdx.x] = cdivr(num, scal[bden_off + bl.slot[threadIdx.x]]);
    }
    __syncthreads();
    c128* pout = last ? ps + (int64_t)cur * stride : acc_p;
    c128* Apout = last ? Aps + (int64_t)cur * stride : acc_Ap;
    constexpr int CH = 4;                    // history vectors loaded per batch: 2*CH 128-bit loads in flight per thread
    constexpr int NFULL = (NH / CH) * CH;
    for (int vs = 0; vs < rg.nvs; vs++) {
    double v[3] = {0., 0., 0.};
    SLAB_STRIDE(i, rg, vs) {
        c128 pc = first ? cmake(0., 0.) : ld_plain(acc_p + i);
        c128 Apc = first ? cmake(0., 0.) : ld_plain(acc_Ap + i);
#pragma unroll 1
        for (int k0 = 0; k0 < NFULL; k0 += CH) {
            c128 hp[CH], hA[CH];
#pragma unroll
            for (int k = 0; k < CH; k++) {
                hp[k] = ld_plain(ps + bl.poff[k0 + k] + i);
                hA[k] = ld_plain(Aps + (int64_t)bl.slot[k0 + k] * stride + i);
            }
#pragma unroll
            for (int k = 0; k < CH; k++) {
                pc = csub(pc, cmul(beta[k0 + k], hp[k]));
                Apc = csub(Apc, cmul(beta[k0 + k], hA[k]));
            }
        }
        if (NH > NFULL) {
            constexpr int TL = NH - NFULL > 0 ? NH - NFULL : 1;
            c128 hp[TL], hA[TL];
#pragma unroll
            for (int k = 0; k < NH - NFULL; k++) {
                hp[k] = ld_plain(ps + bl.poff[NFULL + k] + i);
                hA[k] = ld_plain(Aps + (int64_t)bl.slot[NFULL + k] * stride + i);
            }
#pragma unroll
            for (int k = 0; k < NH - NFULL; k++) {
                pc = csub(pc, cmul(beta[NFULL + k], hp[k]));
                Apc = csub(Apc, cmul(beta[NFULL + k], hA[k]));
            }
        }
        if (last) {
            c128 zv = ld_stream(z + i), av = ld_stream(Ar + i);
            pc = cadd(zv, pc);
            Apc = cadd(av, Apc);
            c128 rv = (r == z) ? zv : ld_stream(r + i);
            c128 t = cmulc(rv, Apc);
            v[0] += t.x; v[1] += t.y;
            v[2] += Apc.x * Apc.x + Apc.y * Apc.y;
        }
        st_stream(pout + i, pc);
        if (keep_Ap) st_stream(Apout + i, Apc);   // (0: the next x update is the last of a blind solve and reads p only -- the inner products of Ap are formed here)
    }
    if (std_conj) v[1] = -v[1];   // <Ap,r> = conj(<r,Ap>), exactly, term by term
    if (last) slab_partial<3>(v, sums, vs);
    }
    if (last) grid_finish<3>(sums, partials, ticket, anum_out, rg, 3, &push);   // -> S_ANUM(2), S_ADEN (global block, or this rank's partial block)
}

// out = a + sign * s * b with the complex scalar s in device memory (Gram-Schmidt updates: src/MG.h:116-118, 192-194)
static __global__ void __launch_bounds__(RED_THREADS) k_axpy_devscal(int64_t n, const double* __restrict__ s2, double sign, const c128* b,
                                                              const c128* a, c128* out) {
    PDL_ENTRY();
    const c128 s = cmake(s2[0], s2[1]);
    GRID_STRIDE(i, n) {
        c128 t = cmul(s, ld_plain(b + i));
        c128 av = ld_plain(a + i);
        st_stream(out + i, sign < 0 ? csub(av, t) : cadd(av, t));
    }
}
