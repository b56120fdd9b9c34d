// csrc/rows.cuh -- device-side row access to matrix-like operators (sliced-ELL, matrix-free hopping stencil, block-CSR,
// each optionally wrapped as DiracOp = diag - k D).  Two uses:
//   for_each(i, f)   enumerate the (column, value) entries of row i           -> Galerkin coarse-operator assembly (mg.cu)
//   apply_row(i, x)  (A x)_i with exactly the arithmetic of the streaming apply kernels (ops.cu) -> the persistent
//                    small-level GCR kernel (gcr_small.cu), where x changes inside the kernel (plain coherent loads)
#pragma once
#include "ops.cuh"

struct SellRows {
    const int64_t* slice_ptr; const int32_t* col; const c128* val;
    int dirac; c128 k; const double* diag;
    template <class F> __device__ __forceinline__ void for_each(int64_t i, F f) const {
        const int64_t slice = i >> 5; const int lane = (int)(i & 31);
        const int64_t base = slice_ptr[slice];
        const int w = (int)((slice_ptr[slice + 1] - base) >> 5);
        for (int t = 0; t < w; t++) {
            c128 v = val[base + (int64_t)t * 32 + lane];
            if (v.x == 0. && v.y == 0.) continue;
            if (dirac) { v = cmul(k, v); v = cmake(-v.x, -v.y); }     // 1 - k D: src/Operator.h:111-112
            f((int64_t)col[base + (int64_t)t * 32 + lane], v);
        }
        if (dirac) f(i, cmake(diag ? diag[i] : 1., 0.));
    }
    // k_sell_spmv (ops.cu): sum in CSR order, then diag.x_i - k sum
    __device__ __forceinline__ c128 apply_row(int64_t i, const c128* x) const {
        const int64_t slice = i >> 5; const int lane = (int)(i & 31);
        const int64_t base = __ldg(slice_ptr + slice);
        const int w = (int)((__ldg(slice_ptr + slice + 1) - base) >> 5);
        c128 sum = cmake(0., 0.);
        for (int t = 0; t < w; t++) {
            const int64_t e = base + (int64_t)t * 32 + lane;
            sum = cadd(sum, cmul(ld_stream(val + e), x[__ldg(col + e)]));
        }
        if (dirac) {
            c128 xr = x[i];
            if (diag) { double d = __ldg(diag + i); xr = cmake(d * xr.x, d * xr.y); }
            sum = csub(xr, cmul(k, sum));
        }
        return sum;
    }
};

struct HopRows {
    int64_t n2, n1, n0;   // local planes, rows, columns (single GPU: global)
    int dirac; c128 k; const double* diag;
    // slab-partitioned operator: the neighbour planes owned by rank-1 / rank+1 are addressed as ghost columns
    // n_local + p (lower plane) and n_local + ghost_lo + p (upper plane), p = y*n0 + x
    int has_lo, has_hi; int64_t ghost_lo;
    // variable bonds (NULL = unit hopping), layout of HoppingOp::d_face
    const double* fz; const double* fy; const double* fx;
    __device__ __forceinline__ c128 entry(double f) const {
        c128 v = cmake(f, 0.);
        if (dirac) { v = cmul(k, v); v = cmake(-v.x, -v.y); }
        return v;
    }
    template <class F> __device__ __forceinline__ void for_each(int64_t i, F f) const {
        const int64_t x = i % n0, y = (i / n0) % n1, z = i / (n0 * n1);
        const int64_t plane = n0 * n1;
        const bool var = fx != nullptr;
        const c128 v = entry(1.);
        if (z > 0) f(i - plane, var ? entry(fz[i]) : v); else if (has_lo) f(n2 * plane + y * n0 + x, var ? entry(fz[i]) : v);
        if (y > 0) f(i - n0, var ? entry(fy[i - n0]) : v);
        if (x > 0) f(i - 1, var ? entry(fx[i - 1]) : v);
        if (x < n0 - 1) f(i + 1, var ? entry(fx[i]) : v);
        if (y < n1 - 1) f(i + n0, var ? entry(fy[i]) : v);
        if (z < n2 - 1) f(i + plane, var ? entry(fz[i + plane]) : v); else if (has_hi) f(n2 * plane + ghost_lo + y * n0 + x, var ? entry(fz[i + plane]) : v);
        if (dirac) f(i, cmake(diag ? diag[i] : 1., 0.));
    }
    // k_hopping (ops.cu): neighbours in ascending column order z-1, y-1, x-1, x+1, y+1, z+1 starting from the first
    __device__ __forceinline__ c128 apply_row(int64_t i, const c128* x) const {
        const int64_t cx = i % n0, cy = (i / n0) % n1, cz = i / (n0 * n1);
        const int64_t plane = n0 * n1;
        const c128 zero = cmake(0., 0.);
        c128 vzm = cz > 0 ? x[i - plane] : zero, vym = cy > 0 ? x[i - n0] : zero, vxm = cx > 0 ? x[i - 1] : zero;
        c128 vxp = cx < n0 - 1 ? x[i + 1] : zero, vyp = cy < n1 - 1 ? x[i + n0] : zero, vzp = cz < n2 - 1 ? x[i + plane] : zero;
        if (fx) {
            const double czm = cz > 0 ? __ldg(fz + i) : 0., cym = cy > 0 ? __ldg(fy + i - n0) : 0., cxm = cx > 0 ? __ldg(fx + i - 1) : 0.;
            const double cxp = cx < n0 - 1 ? __ldg(fx + i) : 0., cyp = cy < n1 - 1 ? __ldg(fy + i) : 0., czp = cz < n2 - 1 ? __ldg(fz + i + plane) : 0.;
            vzm = cmake(czm * vzm.x, czm * vzm.y); vym = cmake(cym * vym.x, cym * vym.y); vxm = cmake(cxm * vxm.x, cxm * vxm.y);
            vxp = cmake(cxp * vxp.x, cxp * vxp.y); vyp = cmake(cyp * vyp.x, cyp * vyp.y); vzp = cmake(czp * vzp.x, czp * vzp.y);
        }
        c128 s = cadd(vzm, vym);
        s = cadd(s, vxm);
        s = cadd(s, vxp);
        s = cadd(s, vyp);
        s = cadd(s, vzp);
        if (dirac) {
            c128 xr = x[i];
            if (diag) { double d = __ldg(diag + i); xr = cmake(d * xr.x, d * xr.y); }
            s = csub(xr, cmul(k, s));
        }
        return s;
    }
};

struct BlockRows {
    const int32_t* brow; const int32_t* bcol; const c128* bval; int ne;
    template <class F> __device__ __forceinline__ void for_each(int64_t i, F f) const {
        const int64_t R = i / ne; const int r = (int)(i - R * ne);
        for (int l = brow[R]; l < brow[R + 1]; l++) {
            const c128* m = bval + (int64_t)l * ne * ne + r;
            const int64_t c0 = (int64_t)bcol[l] * ne;
            for (int c = 0; c < ne; c++) f(c0 + c, m[(int64_t)c * ne]);
        }
    }
    // k_blockcsr_apply (ops.cu): per block o = sum_c m[r][c] x[c], value += o in block order
    __device__ __forceinline__ c128 apply_row(int64_t i, const c128* x) const {
        const int64_t R = i / ne; const int r = (int)(i - R * ne);
        c128 value = cmake(0., 0.);
        const int lb = __ldg(brow + R), le = __ldg(brow + R + 1);
        for (int l = lb; l < le; l++) {
            const c128* xb = x + (int64_t)__ldg(bcol + l) * ne;
            const c128* m = bval + (int64_t)l * ne * ne + r;
            c128 o = cmake(0., 0.);
            for (int c = 0; c < ne; c++) o = cadd(o, cmul(ld_stream(m + (int64_t)c * ne), xb[c]));
            value = cadd(value, o);
        }
        return value;
    }
};

// Calls f(rows) with the row accessor of a matrix-like operator; returns false when A has none (solvers, callbacks).
// Slab-partitioned operators are accepted only with allow_dist (their rows then reference ghost columns >= n_local,
// lower neighbour's plane first): the multigrid set-up understands those, the persistent small-level solver does not.
template <class F>
static inline bool with_rows(mgcr_op* A, F&& f, int* status, bool allow_dist = false) {
    int dirac = 0; c128 k = cmake(0., 0.); const double* diag = nullptr;
    if (A->kind == OP_DIRAC) {
        DiracOp* d = static_cast<DiracOp*>(A);
        dirac = 1; k = d->k; diag = d->d_diag; A = d->D;
    }
    if (A->kind == OP_SELL) {
        SellOp* s = static_cast<SellOp*>(A);
        if (s->halo) return false;
        *status = f(SellRows{s->d_slice_ptr, s->d_col, s->d_val, dirac, k, diag});
        return true;
    }
    if (A->kind == OP_HOPPING) {
        HoppingOp* h = static_cast<HoppingOp*>(A);
        if (h->distributed && !allow_dist) return false;
        const int lo = h->distributed && h->ctx->rank > 0, hi = h->distributed && h->ctx->rank + 1 < h->ctx->nranks;
        *status = f(HopRows{h->n2_local, h->gdims[1], h->gdims[2], dirac, k, diag, lo, hi, lo ? h->gdims[1] * h->gdims[2] : 0,
                             h->var ? h->d_face[0] : nullptr, h->var ? h->d_face[1] : nullptr, h->var ? h->d_face[2] : nullptr});
        return true;
    }
    if (A->kind == OP_BLOCKCSR && !dirac) {
        BlockCsrOp* bo = static_cast<BlockCsrOp*>(A);
        if (bo->halo && !allow_dist) return false;
        if (!bo->d_bval) return false;   // assembly copy dropped (large operator, streaming image only)
        *status = f(BlockRows{bo->d_brow, bo->d_bcol, bo->d_bval, bo->ne});
        return true;
    }
    return false;
}
