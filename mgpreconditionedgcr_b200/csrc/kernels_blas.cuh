// csrc/kernels_blas.cuh -- streaming complex128 vector kernels: Field BLAS-1 (src/Fields.h) and the fused kernels of
// the GCR iteration (src/GCR.h:222-288).  All are HBM-bound: one 128-bit access per element and vector, grid-stride
// over a grid sized in multiples of the SM count, reductions through grid_reduce (common.cuh).
#pragma once
#include "common.cuh"

#define GRID_STRIDE(i, n) for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

static __global__ void __launch_bounds__(RED_THREADS) k_fill(int64_t n, c128 v, c128* __restrict__ out) {
    GRID_STRIDE(i, n) st_stream(out + i, v);
}

// out = a + s*b   (Field::operator+ / operator- / += / -=; the product is formed first, as `s * field[i]`)
static __global__ void __launch_bounds__(RED_THREADS) k_axpy(int64_t n, c128 s, const c128* b, const c128* a, c128* out) {
    GRID_STRIDE(i, n) {
        c128 t = cmul(s, ld_plain(b + i));
        st_stream(out + i, cadd(ld_plain(a + i), t));
    }
}

static __global__ void __launch_bounds__(RED_THREADS) k_scale(int64_t n, c128 s, const c128* a, c128* out) {
    GRID_STRIDE(i, n) st_stream(out + i, cmul(s, ld_plain(a + i)));
}

// a *= 1./sqrt(*nrm2)   (Field::normalise: `field[i] *= 1./norm`, complex *= real-as-complex... the reference
// multiplies by the double 1./norm, i.e. component-wise)
static __global__ void __launch_bounds__(RED_THREADS) k_scale_inv_sqrt(int64_t n, const double* __restrict__ nrm2, c128* a) {
    const double s = 1. / sqrt(*nrm2);
    GRID_STRIDE(i, n) {
        c128 v = ld_plain(a + i);
        st_stream(a + i, cmake(v.x * s, v.y * s));
    }
}

static __global__ void __launch_bounds__(RED_THREADS) k_dot(int64_t n, const c128* __restrict__ a, const c128* __restrict__ b,
                                                     double* partials, unsigned int* ticket, double* out) {
    double v[2] = {0., 0.};
    GRID_STRIDE(i, n) {
        c128 t = cmulc(ld_stream(a + i), ld_stream(b + i));
        v[0] += t.x; v[1] += t.y;
    }
    grid_reduce<2>(v, partials, ticket, out);
}

static __global__ void __launch_bounds__(RED_THREADS) k_norm2(int64_t n, const c128* __restrict__ a, double* partials,
                                                       unsigned int* ticket, double* out) {
    double v[1] = {0.};
    GRID_STRIDE(i, n) {
        c128 t = ld_stream(a + i);
        v[0] += t.x * t.x + t.y * t.y;
    }
    grid_reduce<1>(v, partials, ticket, out);
}

// gamma5 permutation along an axis of extent axis_dim with `inner` elements below it (src/Fields.h:310-339):
// out[.., s', ..] = in[.., s, ..] with s' = s^2 for s < 4 (0<->2, 1<->3), identity above.
static __global__ void __launch_bounds__(RED_THREADS) k_gamma5(int64_t n, int64_t inner, int64_t axis_dim, const c128* __restrict__ in,
                                                        c128* __restrict__ out) {
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

// init: <r,Ap>, <Ap,Ap>, ||rhs||^2 (r = rhs at start) in one pass -> S_ANUM(2), S_ADEN, S_BB
static __global__ void __launch_bounds__(RED_THREADS) k_gcr_init(int64_t n, const c128* __restrict__ r, const c128* __restrict__ Ap,
                                                          int std_conj, double* partials, unsigned int* ticket, double* out4) {
    double v[4] = {0., 0., 0., 0.};
    GRID_STRIDE(i, n) {
        c128 rv = ld_stream(r + i), av = ld_stream(Ap + i);
        c128 t = std_conj ? cmulc(av, rv) : cmulc(rv, av);
        v[0] += t.x; v[1] += t.y;
        v[2] += av.x * av.x + av.y * av.y;
        v[3] += rv.x * rv.x + rv.y * rv.y;
    }
    grid_reduce<4>(v, partials, ticket, out4);
}

// x += alpha p ; r -= alpha Ap ; ||r||^2 -> scal[S_RR]      (GCR.h:230-233)
static __global__ void __launch_bounds__(RED_THREADS) k_gcr_update_xr(int64_t n, const c128* __restrict__ p, const c128* __restrict__ Ap,
                                                               c128* x, c128* r, double* scal, int bden_slot, double* partials,
                                                               unsigned int* ticket) {
    const double aden = scal[S_ADEN];
    const c128 alpha = cdivr(cmake(scal[S_ANUM], scal[S_ANUM + 1]), aden);
    // ||Aps[cur]||^2 never changes while the slot lives: cache it for the beta denominators (GCR.h:258 recomputes it)
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[bden_slot] = aden;
    double v[1] = {0.};
    GRID_STRIDE(i, n) {
        c128 pv = ld_stream(p + i), av = ld_stream(Ap + i);
        c128 xv = ld_plain(x + i), rv = ld_plain(r + i);
        xv = cadd(xv, cmul(alpha, pv));
        rv = csub(rv, cmul(alpha, av));
        st_stream(x + i, xv);
        st_stream(r + i, rv);
        v[0] += rv.x * rv.x + rv.y * rv.y;
    }
    grid_reduce<1>(v, partials, ticket, scal + S_RR);
}

// batched <Ar, Aps[slot]> for NH (<= GCR_CHUNK) history vectors in one pass over Ar      (GCR.h:257-258)
// U elements per thread and trip keep enough 128-bit loads in flight when NH is small.
struct HistList { int slot[GCR_CHUNK]; };

template <int NH, int U>
static __global__ void __launch_bounds__(RED_THREADS) k_gcr_dot_hist(int64_t n, const c128* __restrict__ Ar, const c128* __restrict__ Aps,
                                                              int64_t stride, HistList hl, int std_conj, double* out /* 2*NH */,
                                                              double* partials, unsigned int* ticket) {
    double v[2 * NH];
#pragma unroll
    for (int k = 0; k < 2 * NH; k++) v[k] = 0.;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i0 + (U - 1) * T < n; i0 += T * U) {     // full trips: U*(1+NH) independent 128-bit loads in flight
        c128 a[U], h[U][NH];
#pragma unroll
        for (int u = 0; u < U; u++) {
            a[u] = ld_stream(Ar + i0 + u * T);
#pragma unroll
            for (int k = 0; k < NH; k++) h[u][k] = ld_stream(Aps + (int64_t)hl.slot[k] * stride + i0 + u * T);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int k = 0; k < NH; k++) {
                c128 t = std_conj ? cmulc(h[u][k], a[u]) : cmulc(a[u], h[u][k]);
                v[2 * k] += t.x; v[2 * k + 1] += t.y;
            }
        }
    }
    for (; i0 < n; i0 += T) {                        // tail
        const c128 a = ld_stream(Ar + i0);
#pragma unroll
        for (int k = 0; k < NH; k++) {
            const c128 h = ld_stream(Aps + (int64_t)hl.slot[k] * stride + i0);
            c128 t = std_conj ? cmulc(h, a) : cmulc(a, h);
            v[2 * k] += t.x; v[2 * k + 1] += t.y;
        }
    }
    grid_reduce<2 * NH>(v, partials, ticket, out);
}

// p_new = z + sum_i(-beta_i ps[i]) ; Ap_new = Ar + sum_i(-beta_i Aps[i]) written into ring slot `cur`, with the next
// alpha's <r,Ap_new>, <Ap_new,Ap_new> reduced in the same pass        (GCR.h:251-266, 286-287, 230)
// History slots are visited in the reference's order i = 0..lim-1.  More than GCR_CHUNK history vectors are
// processed in several passes: all but the last accumulate pc / Apc in two scratch vectors (the ring slot `cur` may
// itself still be unread history), the last one adds z / Ar, writes the slot and reduces.
// NH is the exact number of history vectors of this pass; MINB (resident CTAs per SM) bounds the registers so that
// long histories do not collapse the occupancy.
struct BetaList { int slot[GCR_CHUNK]; int num_index[GCR_CHUNK]; };

template <int NH, int MINB>
static __global__ void __launch_bounds__(RED_THREADS, MINB) k_gcr_update_p(int64_t n, const c128* z, const c128* Ar, const c128* r, c128* ps,
                                                                    c128* Aps, int64_t stride, BetaList bl, int cur, int first, int last,
                                                                    c128* acc_p, c128* acc_Ap, int std_conj, int bden_off, double* scal,
                                                                    double* partials, unsigned int* ticket) {
    constexpr int NHS = NH > 0 ? NH : 1;
    __shared__ c128 beta[NHS];
    if ((int)threadIdx.x < NH) {
        int q = bl.num_index[threadIdx.x];
        beta[threadIdx.x] = cdivr(cmake(scal[S_BNUM + 2 * q], scal[S_BNUM + 2 * q + 1]), scal[bden_off + bl.slot[threadIdx.x]]);
    }
    __syncthreads();
    c128* pout = last ? ps + (int64_t)cur * stride : acc_p;
    c128* Apout = last ? Aps + (int64_t)cur * stride : acc_Ap;
    double v[3] = {0., 0., 0.};
    constexpr int CH = 4;                    // history vectors loaded per batch: 2*CH 128-bit loads in flight per thread
    constexpr int NFULL = (NH / CH) * CH;
    GRID_STRIDE(i, n) {
        c128 pc = first ? cmake(0., 0.) : ld_plain(acc_p + i);
        c128 Apc = first ? cmake(0., 0.) : ld_plain(acc_Ap + i);
#pragma unroll 1
        for (int k0 = 0; k0 < NFULL; k0 += CH) {
            c128 hp[CH], hA[CH];
#pragma unroll
            for (int k = 0; k < CH; k++) {
                hp[k] = ld_plain(ps + (int64_t)bl.slot[k0 + k] * stride + i);
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
                hp[k] = ld_plain(ps + (int64_t)bl.slot[NFULL + k] * stride + i);
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
            c128 t = std_conj ? cmulc(Apc, rv) : cmulc(rv, Apc);
            v[0] += t.x; v[1] += t.y;
            v[2] += Apc.x * Apc.x + Apc.y * Apc.y;
        }
        st_stream(pout + i, pc);
        st_stream(Apout + i, Apc);
    }
    if (last) grid_reduce<3>(v, partials, ticket, scal + S_ANUM);   // -> S_ANUM(2), S_ADEN
}

// out = a + sign * s * b with the complex scalar s in device memory (Gram-Schmidt updates: src/MG.h:116-118, 192-194)
static __global__ void __launch_bounds__(RED_THREADS) k_axpy_devscal(int64_t n, const double* __restrict__ s2, double sign, const c128* b,
                                                              const c128* a, c128* out) {
    const c128 s = cmake(s2[0], s2[1]);
    GRID_STRIDE(i, n) {
        c128 t = cmul(s, ld_plain(b + i));
        c128 av = ld_plain(a + i);
        st_stream(out + i, sign < 0 ? csub(av, t) : cadd(av, t));
    }
}
