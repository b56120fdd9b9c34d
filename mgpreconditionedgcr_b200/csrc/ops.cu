// csrc/ops.cu -- operator applies: sliced-ELL SpMV (Sparse::operator(), src/Operator.h:330-346), fused DiracOp
// (src/Operator.h:569-574), matrix-free hopping stencil, block-CSR coarse operator (src/HierarchicalSparse.h:101-161).
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include <algorithm>

#include "kernels_blas.cuh"
#include "ops.cuh"

// ----------------------------------------------------------------------------------------------------------
// sliced-ELL SpMV: one thread per row, one warp per slice; lane l streams val/col at base + j*32 + l (coalesced),
// gathers x through L1/L2 and accumulates in CSR order.  DIRAC fuses y = diag.x - k (D x).
// ----------------------------------------------------------------------------------------------------------
template <bool DIRAC>
__global__ void __launch_bounds__(256) k_sell_spmv(int64_t nrow, const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ col,
                                                   const c128* __restrict__ val, const c128* __restrict__ x, const c128* __restrict__ ghost,
                                                   int64_t n_local, c128 k, const double* __restrict__ diag, const c128* __restrict__ bsub,
                                                   c128* __restrict__ y) {
    PDL_ENTRY();
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t slice = row >> 5;
    const int lane = threadIdx.x & 31;
    if (slice * 32 >= nrow) return;
    const int64_t base = __ldg(slice_ptr + slice);
    const int width = (int)((__ldg(slice_ptr + slice + 1) - base) >> 5);
    c128 sum = cmake(0., 0.);
    const int32_t* cp = col + base + lane;
    const c128* vp = val + base + lane;
#pragma unroll 4
    for (int j = 0; j < width; j++) {
        const int32_t c = __ldg(cp + (int64_t)j * 32);
        const c128 v = ld_stream(vp + (int64_t)j * 32);
        const c128 xv = (c < n_local) ? __ldg(x + c) : __ldg(ghost + (c - n_local));
        sum = cadd(sum, cmul(v, xv));
    }
    if (row < nrow) {
        if (DIRAC) {
            c128 xr = __ldg(x + row);
            if (diag) { double d = __ldg(diag + row); xr = cmake(d * xr.x, d * xr.y); }
            sum = csub(xr, cmul(k, sum));
        }
        if (bsub) sum = csub(__ldg(bsub + row), sum);   // residual b - A x
        st_stream(y + row, sum);
    }
}

SellOp::~SellOp() {
    dev_free(ctx, d_slice_ptr); dev_free(ctx, d_col); dev_free(ctx, d_val);
    halo_free(ctx, halo);
}

static int sell_launch(SellOp* op, const c128* x, c128* y, bool dirac, c128 k, const double* diag, const c128* bsub = nullptr) {
    mgcr_ctx* ctx = op->ctx;
    ARG_CHECK(x != y, "operator apply: input and output alias");
    const c128* ghost = nullptr;
    if (op->halo) { MGCR_TRY(halo_exchange(ctx, op->halo, x)); MGCR_TRY(dist_halo_wait(ctx)); ghost = op->halo->ghost_cur; }
    if (op->nrow == 0) return MGCR_OK;
    int grid = (int)((op->nslices * 32 + 255) / 256);
    if (dirac)
        KLAUNCH(ctx, "sell_dirac", op->apply_bytes(), (launch_pdl(ctx, k_sell_spmv<true>, grid, 256, 0, op->nrow, op->d_slice_ptr, op->d_col, op->d_val, x, ghost, op->n_local, k, diag, bsub, y)));
    else
        KLAUNCH(ctx, "sell_spmv", op->apply_bytes(), (launch_pdl(ctx, k_sell_spmv<false>, grid, 256, 0, op->nrow, op->d_slice_ptr, op->d_col, op->d_val, x, ghost, op->n_local, k, diag, bsub, y)));
    CHECK_LAUNCH();
    return MGCR_OK;
}
int SellOp::apply(const c128* x, c128* y) { return sell_launch(this, x, y, false, cmake(0., 0.), nullptr); }
int SellOp::apply_dirac(const c128* x, c128* y, c128 k, const double* diag, const c128* bsub) { return sell_launch(this, x, y, true, k, diag, bsub); }

int sell_build(mgcr_ctx* ctx, int64_t nrow, int64_t ncol_addressable, const int64_t* row, const int64_t* col, const mgcr_c128* val, SellOp* op);   // build.cu
int device_exclusive_scan_i64(mgcr_ctx* ctx, int64_t* d_data, int64_t n, int64_t* d_total);

extern "C" int mgcr_csr_create(mgcr_ctx* ctx, int64_t nrow, int64_t ncol, const int64_t* row, const int64_t* col, const mgcr_c128* val, mgcr_op** out) {
    ARG_CHECK(ctx && out && row && nrow >= 0 && ncol >= 0, "mgcr_csr_create: bad argument");
    ARG_CHECK(ctx->nranks == 1, "mgcr_csr_create: context is distributed, use mgcr_csr_create_dist");
    ARG_CHECK(row[nrow] == 0 || (col && val), "mgcr_csr_create: NULL col/val");
    *out = nullptr;
    SellOp* op = new SellOp();
    op->kind = OP_SELL; op->ctx = ctx; op->ncol = ncol; op->n_local = ncol; op->n_global = ncol;
    int st = sell_build(ctx, nrow, ncol, row, col, val, op);
    if (st != MGCR_OK) { delete op; return st; }
    *out = op;
    return MGCR_OK;
}

// Row-slab-partitioned CSR: this rank holds rows [row_begin, row_end) with GLOBAL column indices.  Columns
// inside the slab address the local vector; every other column becomes a ghost: the sorted list of distinct off-slab
// columns is the ghost buffer's layout (ascending global index = grouped by owner, owners ascending), the lists are
// all-gathered once, and each rank reads off them which of its elements every other rank wants, in that rank's ghost
// order (pack list).  An apply packs, exchanges (one send/recv pair per communicating peer in one NCCL group) and runs the
// same sliced-ELL kernel with columns >= n_local redirected to the ghost buffer.  Collective.
extern "C" int mgcr_csr_create_dist(mgcr_ctx* ctx, int64_t nrow_global, int64_t row_begin, int64_t row_end, const int64_t* row,
                                    const int64_t* col, const mgcr_c128* val, mgcr_op** out) {
    ARG_CHECK(ctx && out && row && row_begin >= 0 && row_end >= row_begin && row_end <= nrow_global, "mgcr_csr_create_dist: bad argument");
    *out = nullptr;
    const int64_t nl = row_end - row_begin, nnz = row[nl];
    ARG_CHECK(nnz == 0 || (col && val), "mgcr_csr_create_dist: NULL col/val");
    if (ctx->nranks == 1) {
        ARG_CHECK(row_begin == 0 && row_end == nrow_global, "mgcr_csr_create_dist: one rank must hold every row");
        return mgcr_csr_create(ctx, nl, nrow_global, row, col, val, out);
    }
    std::vector<int64_t> begins, ends;
    MGCR_TRY(dist_allgather_host_i64(ctx, row_begin, begins));
    MGCR_TRY(dist_allgather_host_i64(ctx, row_end, ends));
    for (int r = 0; r < ctx->nranks; r++) {
        const int64_t expect = r == 0 ? 0 : ends[(size_t)r - 1];
        ARG_CHECK(begins[(size_t)r] == expect && (r + 1 < ctx->nranks || ends[(size_t)r] == nrow_global),
                  "mgcr_csr_create_dist: the row ranges of the ranks are not a contiguous ascending partition of [0, %lld)", (long long)nrow_global);
    }
    // ghost columns
    std::vector<int64_t> ghost;
    int bad = 0;
    for (int64_t l = 0; l < nnz; l++) {
        const int64_t c = col[l];
        if (c < 0 || c >= nrow_global) { bad = 1; continue; }
        if (c < row_begin || c >= row_end) ghost.push_back(c);
    }
    std::sort(ghost.begin(), ghost.end());
    ghost.erase(std::unique(ghost.begin(), ghost.end()), ghost.end());
    std::vector<int64_t> bads;
    MGCR_TRY(dist_allgather_host_i64(ctx, bad, bads));
    for (int64_t b : bads) ARG_CHECK(b == 0, "mgcr_csr_create_dist: column index out of range (src/Operator.h:332 asserts f.field_size() == dim)");
    std::vector<int64_t> lcol((size_t)std::max<int64_t>(nnz, 1));
    for (int64_t l = 0; l < nnz; l++) {
        const int64_t c = col[l];
        lcol[(size_t)l] = (c >= row_begin && c < row_end) ? c - row_begin : nl + (std::lower_bound(ghost.begin(), ghost.end(), c) - ghost.begin());
    }
    // everybody learns everybody's ghost list
    std::vector<int64_t> sizes;
    MGCR_TRY(dist_allgather_host_i64(ctx, (int64_t)ghost.size(), sizes));
    int64_t maxreq = 0;
    for (int64_t v : sizes) maxreq = std::max(maxreq, v);
    std::vector<unsigned char> all;
    {
        std::vector<int64_t> padded((size_t)std::max<int64_t>(maxreq, 1), -1);
        std::copy(ghost.begin(), ghost.end(), padded.begin());
        MGCR_TRY(dist_allgather_host_bytes(ctx, padded.data(), sizeof(int64_t) * (size_t)maxreq, all));
    }
    const int64_t* lists = (const int64_t*)all.data();
    SellOp* op = new SellOp();
    op->kind = OP_SELL; op->ctx = ctx; op->ncol = nl + (int64_t)ghost.size(); op->n_local = nl; op->n_global = nrow_global; op->distributed = true;
    HaloPlan* h = new HaloPlan();
    op->halo = h;
    h->elem = 1; h->n_ghost = (int64_t)ghost.size();
    h->send_off.push_back(0); h->recv_off.push_back(0);
    std::vector<int32_t> send_idx;
    for (int r = 0; r < ctx->nranks; r++) {
        if (r == ctx->rank) continue;
        // what rank r wants from me, in its ghost order
        const int64_t* lr = lists + (size_t)r * (size_t)maxreq;
        const int64_t* lo = std::lower_bound(lr, lr + sizes[(size_t)r], row_begin);
        const int64_t* hi = std::lower_bound(lr, lr + sizes[(size_t)r], row_end);
        // what I want from rank r
        const int64_t g0 = std::lower_bound(ghost.begin(), ghost.end(), begins[(size_t)r]) - ghost.begin();
        const int64_t g1 = std::lower_bound(ghost.begin(), ghost.end(), ends[(size_t)r]) - ghost.begin();
        if (hi == lo && g1 == g0) continue;
        h->peer.push_back(r);
        for (const int64_t* q = lo; q < hi; q++) send_idx.push_back((int32_t)(*q - row_begin));
        h->send_off.push_back((int64_t)send_idx.size());
        h->recv_off.push_back(g1);   // ghosts are sorted by global column: rank r's group ends at g1
        h->send_start.push_back(0);
    }
    h->npeers = (int)h->peer.size();
    int st = sell_build(ctx, nl, op->ncol, row, lcol.data(), val, op);
    if (st == MGCR_OK) st = dev_alloc_t(ctx, std::max<size_t>(send_idx.size(), 1), &h->d_send_idx);
    if (st == MGCR_OK) st = dev_alloc_t(ctx, std::max<size_t>(send_idx.size(), 1), &h->d_send_buf);
    if (st == MGCR_OK) st = dev_alloc_t(ctx, std::max<size_t>(ghost.size(), 1), &h->d_ghost);
    if (st == MGCR_OK && !send_idx.empty()) {
        cudaError_t e = cudaMemcpyAsync(h->d_send_idx, send_idx.data(), sizeof(int32_t) * send_idx.size(), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { mgcr_set_error("mgcr_csr_create_dist: pack-list upload: %s", cudaGetErrorString(e)); st = MGCR_ERR_CUDA; }
    }
    if (st != MGCR_OK) { delete op; return st; }
    *out = op;
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// matrix-free hopping stencil.  A CTA owns a TX x TY tile of the (n1, n0) plane and marches along n2: the plane
// being processed sits in shared memory (with its one-element halo ring) for the x/y neighbours, the z neighbours
// live in registers (prev / cur / next), so every element is read from global memory once per CTA (+ ring).
// Neighbour sum order = ascending column order of the CSR the reference would hold: z-1, y-1, x-1, x+1, y+1, z+1.
// ----------------------------------------------------------------------------------------------------------
enum { HOP_TX = 32, HOP_TY = 16 };

struct HopArgs {
    int64_t n2, n1, n0;      // local planes, rows, columns
    int64_t zc;              // planes per z-chunk
    int64_t z_lo, z_hi;      // planes [z_lo, z_hi) are computed by this launch (interior / boundary split of the halo overlap)
    const c128* x; c128* y;
    const c128* halo_lo; const c128* halo_hi;   // plane below local z=0 / above z=n2-1 (NULL = Dirichlet)
    const uint32_t* flag_lo; const uint32_t* flag_hi; uint32_t flag_seq;   // deferred halo wait (p2p.cu): poll before the first read of halo_lo / halo_hi
    int dirac; c128 k; const double* diag;
    const c128* bsub;        // non-NULL: store b - (A x) (the multigrid residual)
    // variable bond coefficients (NULL = unit hopping): fx[i] / fy[i] = bond between site i and i+1 / i+n0,
    // fz[z*plane + c] = bond between plane z-1 and plane z (n2+1 planes: the first / last is the bond to the slab neighbour)
    const double* fz; const double* fy; const double* fx;
};

__global__ void __launch_bounds__(HOP_TX* HOP_TY) k_hopping(HopArgs a) {
    __shared__ c128 sp[2][HOP_TY + 2][HOP_TX + 2];
    const int tx = threadIdx.x % HOP_TX, ty = threadIdx.x / HOP_TX;
    const int64_t x0 = (int64_t)blockIdx.x * HOP_TX, y0 = (int64_t)blockIdx.y * HOP_TY;
    const int64_t gx = x0 + tx, gy = y0 + ty;
    const int64_t zs = a.z_lo + (int64_t)blockIdx.z * a.zc;
    const int64_t ze = min(zs + a.zc, a.z_hi);
    const int64_t plane = a.n1 * a.n0;
    const bool inb = gx < a.n0 && gy < a.n1;
    const c128 zero = cmake(0., 0.);
    // halo duties of this thread within a plane: left/right column, bottom/top row
    const bool hx_on = (tx == 0 && gx >= 1 && gy < a.n1) || (tx == HOP_TX - 1 && gx + 1 < a.n0 && gy < a.n1);
    const int64_t hx_off = gy * a.n0 + (tx == 0 ? gx - 1 : gx + 1);
    const bool hy_on = (ty == 0 && gy >= 1 && gx < a.n0) || (ty == HOP_TY - 1 && gy + 1 < a.n1 && gx < a.n0);
    const int64_t hy_off = (ty == 0 ? gy - 1 : gy + 1) * a.n0 + gx;
    const int64_t c_off = gy * a.n0 + gx;

    auto plane_ptr = [&](int64_t z) -> const c128* {
        if (z < 0) return a.halo_lo;
        if (z >= a.n2) return a.halo_hi;
        return a.x + z * plane;
    };
    const c128* pp = plane_ptr(zs - 1);
    const c128* pc = plane_ptr(zs);
    c128 prev = (inb && pp) ? ld_stream(pp + c_off) : zero;
    c128 cur = inb ? ld_stream(pc + c_off) : zero;
    c128 hx = hx_on ? ld_stream(pc + hx_off) : zero;
    c128 hy = hy_on ? ld_stream(pc + hy_off) : zero;
    for (int64_t z = zs; z < ze; z++) {
        const c128* pn = plane_ptr(z + 1);
        const c128 next = (inb && pn) ? ld_stream(pn + c_off) : zero;
        c128 nhx = zero, nhy = zero;
        if (z + 1 < ze) {   // ring of the next plane, prefetched one step ahead
            if (hx_on) nhx = ld_stream(pn + hx_off);
            if (hy_on) nhy = ld_stream(pn + hy_off);
        }
        const int b = (int)((z - zs) & 1);
        sp[b][ty + 1][tx + 1] = cur;
        if (tx == 0) sp[b][ty + 1][0] = hx;
        if (tx == HOP_TX - 1) sp[b][ty + 1][HOP_TX + 1] = hx;
        if (ty == 0) sp[b][0][tx + 1] = hy;
        if (ty == HOP_TY - 1) sp[b][HOP_TY + 1][tx + 1] = hy;
        __syncthreads();
        if (inb) {
            c128 s = cadd(prev, sp[b][ty][tx + 1]);
            s = cadd(s, sp[b][ty + 1][tx]);
            s = cadd(s, sp[b][ty + 1][tx + 2]);
            s = cadd(s, sp[b][ty + 2][tx + 1]);
            s = cadd(s, next);
            if (a.dirac) {
                c128 xr = cur;
                if (a.diag) { double d = __ldg(a.diag + z * plane + c_off); xr = cmake(d * xr.x, d * xr.y); }
                s = csub(xr, cmul(a.k, s));
            }
            if (a.bsub) s = csub(__ldg(a.bsub + z * plane + c_off), s);
            st_stream(a.y + z * plane + c_off, s);
        }
        prev = cur; cur = next; hx = nhx; hy = nhy;
    }
}

// Second form of the same stencil, without shared memory or block-wide barriers: a thread owns one (x, y) column of a z-chunk
// and marches along z with the z-neighbours in registers (prev / cur / next, the plane after next already in flight); the
// x and y neighbours are read through L1/L2 -- they were just loaded as own-column elements by the neighbouring lanes /
// warps, so HBM sees every element once (ncu: dram read = 16 B per site, profiles/r01_ncu_full_mg3d_256.md).  Warps never
// wait for each other.  The CTA's threads cover hl_tx consecutive x (whole rows when they fit: long contiguous DRAM bursts)
// times HL_THREADS/hl_tx rows.  It is also the form for 2-D lattices, which are traversed as (ny, 1, nx): 4.4 TB/s against
// 2.4 TB/s for the tile kernel.  Variants measured and rejected on B200 (profiles/r01_stencil_experiments.txt): deeper
// register prefetch of the own column (3.6 TB/s), prefetching the in-plane neighbours one plane ahead (3.8 TB/s), x
// neighbours by warp shuffle (3.4 TB/s), other tile shapes / chunk lengths (all 3.8-4.1 TB/s in 3-D).
enum { HL_THREADS = 512 };

template <bool VAR>
__global__ void __launch_bounds__(HL_THREADS) k_hopping_l1(HopArgs a, int hl_tx) {
    PDL_ENTRY();
    const int hl_ty = HL_THREADS / hl_tx;
    const int64_t gx = (int64_t)blockIdx.x * hl_tx + threadIdx.x % hl_tx;
    const int64_t gy = (int64_t)blockIdx.y * hl_ty + threadIdx.x / hl_tx;
    if (gx >= a.n0 || gy >= a.n1) return;
    const int64_t zs = a.z_lo + (int64_t)blockIdx.z * a.zc;
    const int64_t ze = min(zs + a.zc, a.z_hi);
    const int64_t plane = a.n1 * a.n0;
    const int64_t c_off = gy * a.n0 + gx;
    const c128 zero = cmake(0., 0.);
    const bool xm = gx > 0, xp = gx + 1 < a.n0, ym = gy > 0, yp = gy + 1 < a.n1;
    auto load = [&](int64_t z) -> c128 {   // own column at plane z
        const c128* p = z < 0 ? a.halo_lo : (z >= a.n2 ? a.halo_hi : a.x + z * plane);
        return p ? __ldg(p + c_off) : zero;
    };
    c128 prev = load(zs - 1), cur = load(zs), next = load(zs + 1);
    double fzm = 0., fzp = 0.;   // bonds to the plane below / above the current one
    if (VAR) { fzm = __ldg(a.fz + zs * plane + c_off); fzp = __ldg(a.fz + (zs + 1) * plane + c_off); }
    for (int64_t z = zs; z < ze; z++) {
        const c128 next2 = (z + 2 <= ze) ? load(z + 2) : zero;   // z + 2 == ze is the chunk's upper neighbour plane
        double fzp2 = 0.;
        if (VAR && z + 1 < ze) fzp2 = __ldg(a.fz + (z + 2) * plane + c_off);
        const c128* pc = a.x + z * plane + c_off;
        c128 vym = ym ? __ldg(pc - a.n0) : zero;
        c128 vxm = xm ? __ldg(pc - 1) : zero;
        c128 vxp = xp ? __ldg(pc + 1) : zero;
        c128 vyp = yp ? __ldg(pc + a.n0) : zero;
        c128 vzm = prev, vzp = next;
        if (VAR) {   // real bond x complex neighbour: what the CSR product (f + 0i) * x gives, up to the sign of a zero
            const double* fyc = a.fy + z * plane + c_off;
            const double* fxc = a.fx + z * plane + c_off;
            const double cym = ym ? __ldg(fyc - a.n0) : 0., cxm = xm ? __ldg(fxc - 1) : 0.;
            const double cxp = xp ? __ldg(fxc) : 0., cyp = yp ? __ldg(fyc) : 0.;
            vzm = cmake(fzm * vzm.x, fzm * vzm.y);
            vym = cmake(cym * vym.x, cym * vym.y);
            vxm = cmake(cxm * vxm.x, cxm * vxm.y);
            vxp = cmake(cxp * vxp.x, cxp * vxp.y);
            vyp = cmake(cyp * vyp.x, cyp * vyp.y);
            vzp = cmake(fzp * vzp.x, fzp * vzp.y);
        }
        c128 s = cadd(vzm, vym);
        s = cadd(s, vxm);
        s = cadd(s, vxp);
        s = cadd(s, vyp);
        s = cadd(s, vzp);
        if (a.dirac) {
            c128 xr = cur;
            if (a.diag) { double d = __ldg(a.diag + z * plane + c_off); xr = cmake(d * xr.x, d * xr.y); }
            s = csub(xr, cmul(a.k, s));
        }
        if (a.bsub) s = csub(__ldg(a.bsub + z * plane + c_off), s);
        st_stream(a.y + z * plane + c_off, s);
        prev = cur; cur = next; next = next2;
        fzm = fzp; fzp = fzp2;
    }
}

// Third form: the planes are staged through a shared-memory ring by TMA tensor copies.  A CTA owns a TX x TY tile of the
// (n1, n0) plane and a chunk of planes; ONE producer thread issues, for every plane zs-1 .. ze, a 3-D box copy
// (TX+2) x (TY+2) x 1 of the operand (cp.async.bulk.tensor, completion on an mbarrier) -- elements outside the lattice are
// zero-filled by the copy engine, which is exactly the Dirichlet boundary, so the consumers have no bounds logic on the
// neighbour reads.  16 consumer warps (one site per thread and plane) read the 4 in-plane neighbours and the next plane's
// centre from shared memory (prev / cur stay in registers) and release the stage through a second mbarrier: nobody waits
// on a block-wide barrier, and the bytes in flight per SM are stages x tile, independent of the compiler's load scheduling
// (the register-marching form has one HBM-bound load per thread in flight and sits at 0.63 of the copy peak).
// The operand is described as doubles (2 per element) because the tensor-map element types stop at 8 bytes.
template <int TX, int TY, bool VAR>
struct HopTmaCfg {
    static constexpr int CONSUMERS = TX * TY;
    static constexpr int THREADS = CONSUMERS + 32;
    static constexpr int ROW = TX + 2;
    static constexpr int TILE = ROW * (TY + 2);                       // c128 of the operand per stage
    static constexpr int X_BYTES = ((TILE * 16 + 127) / 128) * 128;
    // variable coefficients: the stage also carries the plane's bonds and diagonal (doubles)
    static constexpr int FZ_OFF = X_BYTES;                            // TX x TY       bond to the plane below
    static constexpr int FY_OFF = FZ_OFF + TX * TY * 8;               // TX x (TY+1)   rows y0-1 .. y0+TY-1
    static constexpr int FXW = TX + 4;                                // columns x0-2 .. x0+TX+1: the box starts on a 16-byte boundary and its rows
                                                                      // are a multiple of 32 bytes (a 66-double box starting at x0-1 traps on B200)
    static constexpr int FX_OFF = FY_OFF + TX * (TY + 1) * 8;         // FXW x TY
    static constexpr int DG_OFF = FX_OFF + FXW * TY * 8;              // TX x TY       diagonal
    static constexpr int STAGE_BYTES = VAR ? DG_OFF + TX * TY * 8 : X_BYTES;
    static_assert(!VAR || (FY_OFF % 128 == 0 && FX_OFF % 128 == 0 && DG_OFF % 128 == 0 && STAGE_BYTES % 128 == 0), "TMA destinations are 128-byte aligned");
};
enum { HOP_TMA_MAX_STAGES = 12, HOP_TMA_SMEM = 112 * 1024 };   // two CTAs per SM: 2 x (112 KB + static + 1 KB reserved) <= 228 KB

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

struct HopMaps { CUtensorMap x, lo, hi, fz, fy, fx, dg, b; };

template <int TX, int TY, bool VAR, bool RES>
__global__ void __launch_bounds__(HopTmaCfg<TX, TY, VAR>::THREADS, 2) k_hopping_tma(const __grid_constant__ HopMaps maps, HopArgs a, int stages) {
    typedef HopTmaCfg<TX, TY, VAR> C;
    // residual form (RES, r = b - A x): the stage of plane z also carries the TX x TY tile of the right-hand side, appended to
    // the operand (and bond) tiles, so that b arrives through the same ring as x instead of one exposed HBM-latency load per
    // thread and plane (0.62 of the copy peak in round 1).  A template parameter: the plain apply compiles to the round-1 code.
    constexpr int B_OFF = C::STAGE_BYTES;
    constexpr int stage_bytes = C::STAGE_BYTES + (RES ? TX * TY * 16 : 0);
    extern __shared__ unsigned char hop_smem_raw[];
    __shared__ __align__(8) uint64_t full[HOP_TMA_MAX_STAGES], empty[HOP_TMA_MAX_STAGES];
    unsigned char* ring = (unsigned char*)(((uintptr_t)hop_smem_raw + 127) & ~(uintptr_t)127);
    const int64_t x0 = (int64_t)blockIdx.x * TX, y0 = (int64_t)blockIdx.y * TY;
    const int64_t zs = a.z_lo + (int64_t)blockIdx.z * a.zc;
    const int64_t ze = min(zs + a.zc, a.z_hi);
    const int nplanes = (int)(ze - zs) + 2;                            // planes zs-1 .. ze
    const bool use_diag = VAR && a.dirac && a.diag;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], C::CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    PDL_ENTRY();
    if (threadIdx.x >= C::CONSUMERS) {
        // ---- producer: one elected thread of the last warp ----
        if (threadIdx.x == C::CONSUMERS) {
            const uint32_t tx_bytes = (uint32_t)(C::TILE * 16) + (VAR ? (uint32_t)(C::DG_OFF - C::FZ_OFF) + (use_diag ? (uint32_t)(TX * TY * 8) : 0u) : 0u);
            for (int p = 0; p < nplanes; p++) {
                const int s = p % stages;
                if (p >= stages) mbar_wait(&empty[s], (uint32_t)((p / stages - 1) & 1));
                const int64_t z = zs - 1 + p;
                const CUtensorMap* m = &maps.x;
                int zc = (int)z;                                       // z = -1 / n2 without a slab neighbour: out of range, zero-filled
                if (z < 0 && a.halo_lo) {
                    // the neighbour's plane may still be on its way: only this CTA's first copy waits for it, the chunks in the
                    // interior of the slab never do.  The copy engine reads through the async proxy: order it after the acquire.
                    if (a.flag_lo) { p2p_flag_wait(a.flag_lo, a.flag_seq); asm volatile("fence.proxy.async;" ::: "memory"); }
                    m = &maps.lo; zc = 0;
                } else if (z >= a.n2 && a.halo_hi) {
                    if (a.flag_hi) { p2p_flag_wait(a.flag_hi, a.flag_seq); asm volatile("fence.proxy.async;" ::: "memory"); }
                    m = &maps.hi; zc = 0;
                }
                unsigned char* st = ring + (size_t)s * stage_bytes;
                const bool own = z >= zs && z < ze;                    // a plane this CTA computes (not a neighbour plane)
                mbar_expect_tx(&full[s], tx_bytes + ((RES && own) ? (uint32_t)(TX * TY * 16) : 0u));
                tma_load_3d(st, m, (int)(2 * (x0 - 1)), (int)(y0 - 1), zc, &full[s]);
                if (VAR) {
                    // bonds below the plane exist for z = 0 .. n2 (n2+1 planes); in-plane bonds and the diagonal only for the
                    // planes that are computed -- neighbour planes get an out-of-range coordinate (zero fill, no traffic)
                    const int zin = (z >= 0 && z < a.n2) ? (int)z : -1;
                    tma_load_3d(st + C::FZ_OFF, &maps.fz, (int)x0, (int)y0, (int)z, &full[s]);
                    tma_load_3d(st + C::FY_OFF, &maps.fy, (int)x0, (int)(y0 - 1), zin, &full[s]);
                    tma_load_3d(st + C::FX_OFF, &maps.fx, (int)(x0 - 2), (int)y0, zin, &full[s]);
                    if (use_diag) tma_load_3d(st + C::DG_OFF, &maps.dg, (int)x0, (int)y0, zin, &full[s]);
                }
                if (RES && own) tma_load_3d(st + B_OFF, &maps.b, (int)(2 * x0), (int)y0, (int)z, &full[s]);
            }
        }
        return;
    }
    // ---- consumers ----
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, lane = threadIdx.x & 31;
    const int64_t gx = x0 + tx, gy = y0 + ty;
    const bool inb = gx < a.n0 && gy < a.n1;
    const bool xp = gx + 1 < a.n0, yp = gy + 1 < a.n1;
    const int64_t plane = a.n1 * a.n0;
    const int64_t c_off = gy * a.n0 + gx;
    const int ctr = (ty + 1) * C::ROW + tx + 1;
    auto stage = [&](int p) -> const unsigned char* { return ring + (size_t)(p % stages) * stage_bytes; };
    mbar_wait(&full[0], 0);
    c128 prev = ((const c128*)stage(0))[ctr];
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[0]);
    mbar_wait(&full[1 % stages], (uint32_t)((1 / stages) & 1));
    c128 cur = ((const c128*)stage(1))[ctr];
    double fzm = 0.;
    if (VAR) fzm = ((const double*)(stage(1) + C::FZ_OFF))[ty * TX + tx];
    for (int p = 1; p + 1 < nplanes; p++) {
        const int64_t z = zs - 1 + p;
        mbar_wait(&full[(p + 1) % stages], (uint32_t)(((p + 1) / stages) & 1));
        const unsigned char* sn = stage(p + 1);
        const unsigned char* sc = stage(p);
        const c128 next = ((const c128*)sn)[ctr];
        const c128* t = (const c128*)sc;
        c128 vzm = prev, vym = t[ctr - C::ROW], vxm = t[ctr - 1], vxp = t[ctr + 1], vyp = t[ctr + C::ROW], vzp = next;
        double fzp = 0., dg = 1.;
        if (VAR) {
            const double* fy = (const double*)(sc + C::FY_OFF);
            const double* fx = (const double*)(sc + C::FX_OFF);
            fzp = ((const double*)(sn + C::FZ_OFF))[ty * TX + tx];
            const double cym = fy[ty * TX + tx];                      // row y-1 (zero-filled below the lattice)
            const double cyp = yp ? fy[(ty + 1) * TX + tx] : 0.;
            const double cxm = fx[ty * C::FXW + tx + 1];              // column x-1 (zero-filled left of the lattice)
            const double cxp = xp ? fx[ty * C::FXW + tx + 2] : 0.;
            if (use_diag) dg = ((const double*)(sc + C::DG_OFF))[ty * TX + tx];
            vzm = cmake(fzm * vzm.x, fzm * vzm.y);
            vym = cmake(cym * vym.x, cym * vym.y);
            vxm = cmake(cxm * vxm.x, cxm * vxm.y);
            vxp = cmake(cxp * vxp.x, cxp * vxp.y);
            vyp = cmake(cyp * vyp.x, cyp * vyp.y);
            vzp = cmake(fzp * vzp.x, fzp * vzp.y);
        }
        c128 bs = cmake(0., 0.);
        if (RES) bs = ((const c128*)(sc + B_OFF))[ty * TX + tx];
        c128 s = cadd(vzm, vym);
        s = cadd(s, vxm);
        s = cadd(s, vxp);
        s = cadd(s, vyp);
        s = cadd(s, vzp);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[p % stages]);               // this warp is done with plane z's stage
        if (inb) {
            if (a.dirac) {
                c128 xr = cur;
                if (VAR) { if (use_diag) xr = cmake(dg * xr.x, dg * xr.y); }
                else if (a.diag) { double d = __ldg(a.diag + z * plane + c_off); xr = cmake(d * xr.x, d * xr.y); }
                s = csub(xr, cmul(a.k, s));
            }
            if (RES) s = csub(bs, s);
            st_stream(a.y + z * plane + c_off, s);
        }
        prev = cur; cur = next; fzm = fzp;
    }
}

typedef CUresult (*tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_fn tensor_map_encoder() {
    static tensor_map_encode_fn fn = []() -> tensor_map_encode_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (tensor_map_encode_fn)p;
    }();
    return fn;
}
// tensor map over `planes` planes of n1 rows of `inner` doubles at `base`, box box0 x box1 x 1, out-of-range elements read as zero
static int hop_tensor_map(CUtensorMap* m, const void* base, int64_t inner, int64_t n1, int64_t planes, int box0, int box1) {
    tensor_map_encode_fn enc = tensor_map_encoder();
    if (!enc) { mgcr_set_error("cuTensorMapEncodeTiled is not available from this driver"); return MGCR_ERR_CUDA; }
    const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)n1, (cuuint64_t)planes};
    const cuuint64_t gstride[2] = {(cuuint64_t)(8 * inner), (cuuint64_t)(8 * inner * n1)};
    const cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { mgcr_set_error("cuTensorMapEncodeTiled failed (%d) for %lld planes of %lld x %lld doubles", (int)r, (long long)planes, (long long)n1, (long long)inner); return MGCR_ERR_CUDA; }
    return MGCR_OK;
}

template <int TX, int TY, bool VAR, bool RES>
static int hop_tma_launch_res(mgcr_ctx* ctx, const HopArgs& a0, int64_t z_lo, int64_t z_hi, const char* name, double bytes) {
    typedef HopTmaCfg<TX, TY, VAR> C;
    HopArgs a = a0;
    static const int stages_env = getenv("MGCR_HOP_STAGES") ? atoi(getenv("MGCR_HOP_STAGES")) : 0;   // experiment knobs
    static const int zc_env = getenv("MGCR_HOP_ZC") ? atoi(getenv("MGCR_HOP_ZC")) : 0;
    // two CTAs per SM (registers); the ring takes what shared memory allows: 8 planes of the operand tile in flight per CTA
    // saturate HBM (profiles/r01_stencil_tma_sweep.txt), 4 with the bond / diagonal tiles riding along
    const int stage_bytes = C::STAGE_BYTES + (RES ? TX * TY * 16 : 0);
    const int stages_max = std::min((int)HOP_TMA_MAX_STAGES, (int)((HOP_TMA_SMEM - 128) / stage_bytes));
    const int stages = std::max(3, std::min(stages_max, stages_env > 0 ? stages_env : 8));
    const size_t smem = (size_t)stages * stage_bytes + 128;
    HopMaps maps;
    MGCR_TRY(hop_tensor_map(&maps.x, a.x, 2 * a.n0, a.n1, a.n2, 2 * (TX + 2), TY + 2));
    maps.lo = maps.x; maps.hi = maps.x; maps.fz = maps.x; maps.fy = maps.x; maps.fx = maps.x; maps.dg = maps.x; maps.b = maps.x;
    if (RES) MGCR_TRY(hop_tensor_map(&maps.b, a.bsub, 2 * a.n0, a.n1, a.n2, 2 * TX, TY));
    if (a.halo_lo) MGCR_TRY(hop_tensor_map(&maps.lo, a.halo_lo, 2 * a.n0, a.n1, 1, 2 * (TX + 2), TY + 2));
    if (a.halo_hi) MGCR_TRY(hop_tensor_map(&maps.hi, a.halo_hi, 2 * a.n0, a.n1, 1, 2 * (TX + 2), TY + 2));
    if (VAR) {
        MGCR_TRY(hop_tensor_map(&maps.fz, a.fz, a.n0, a.n1, a.n2 + 1, TX, TY));
        MGCR_TRY(hop_tensor_map(&maps.fy, a.fy, a.n0, a.n1, a.n2, TX, TY + 1));
        MGCR_TRY(hop_tensor_map(&maps.fx, a.fx, a.n0, a.n1, a.n2, C::FXW, TY));
        if (a.dirac && a.diag) MGCR_TRY(hop_tensor_map(&maps.dg, a.diag, a.n0, a.n1, a.n2, TX, TY));
    }
    MGCR_TRY(ensure_dyn_smem(ctx, (const void*)k_hopping_tma<TX, TY, VAR, RES>, HOP_TMA_SMEM));
    const int64_t nz = z_hi - z_lo;
    dim3 grid((unsigned)((a.n0 + TX - 1) / TX), (unsigned)((a.n1 + TY - 1) / TY), 1);
    // chunks of planes: enough CTAs for ~8 waves of the resident set (the last, partial wave is the tail), but chunks of
    // at least 16 planes (each chunk re-reads its two boundary planes)
    const int64_t tiles = (int64_t)grid.x * grid.y;
    const int64_t resident = (int64_t)ctx->num_sms * 2;
    int64_t nchunks = std::max<int64_t>(1, (8 * resident + tiles - 1) / tiles);
    a.zc = std::max<int64_t>((nz + nchunks - 1) / nchunks, std::min<int64_t>(16, nz));
    if (zc_env > 0) a.zc = std::min<int64_t>(zc_env, nz);
    grid.z = (unsigned)((nz + a.zc - 1) / a.zc);
    ARG_CHECK(grid.y <= 65535 && grid.z <= 65535, "hopping: lattice too large for the launch grid");
    a.z_lo = z_lo; a.z_hi = z_hi;
    KLAUNCH(ctx, name, bytes, (launch_pdl(ctx, k_hopping_tma<TX, TY, VAR, RES>, grid, C::THREADS, smem, maps, a, stages)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

template <int TX, int TY, bool VAR>
static int hop_tma_launch(mgcr_ctx* ctx, const HopArgs& a, int64_t z_lo, int64_t z_hi, const char* name, double bytes) {
    return a.bsub ? hop_tma_launch_res<TX, TY, VAR, true>(ctx, a, z_lo, z_hi, name, bytes) : hop_tma_launch_res<TX, TY, VAR, false>(ctx, a, z_lo, z_hi, name, bytes);
}

HoppingOp::~HoppingOp() {
    for (int d = 0; d < 3; d++) dev_free(ctx, d_face[d]);
    dev_free(ctx, d_halo_lo); dev_free(ctx, d_halo_hi);
    p2p_halo_destroy(ctx, &ph);
}

int HoppingOp::run(const c128* x, c128* y, int dirac, c128 k, const double* diag, const c128* bsub) {
    ARG_CHECK(x != y, "operator apply: input and output alias");
    HopArgs a;
    a.bsub = bsub;
    a.n2 = n2_local; a.n1 = gdims[1]; a.n0 = gdims[2];
    a.x = x; a.y = y; a.dirac = dirac; a.k = k; a.diag = diag;
    a.halo_lo = nullptr; a.halo_hi = nullptr;
    a.flag_lo = nullptr; a.flag_hi = nullptr; a.flag_seq = 0;
    a.fz = d_face[0]; a.fy = d_face[1]; a.fx = d_face[2];
    const int64_t plane = a.n1 * a.n0;
    const int lo = ctx->rank - 1, hi = ctx->rank + 1;
    const bool has_lo = distributed && lo >= 0, has_hi = distributed && hi < ctx->nranks;
    const bool tma_form = ctx->hopping_kernel == 2 && a.n0 >= 64 && a.n1 >= 8 && (!var || d_face[1]) && n_local >= ctx->hopping_tma_rows;
    if (distributed && ph.on) {
        // one plane to each slab neighbour, stored straight into its receive buffer over NVLink (p2p.cu); the TMA-staged kernel
        // waits for the neighbours' planes itself, in the producer threads of the CTAs that touch them
        static const int defer_env = getenv("MGCR_HALO_DEFER") ? atoi(getenv("MGCR_HALO_DEFER")) : 1;
        const bool defer = tma_form && defer_env && n_local > 0;
        MGCR_TRY(p2p_halo_exchange(ctx, &ph, x, x + (n2_local - 1) * plane, &a.halo_lo, &a.halo_hi, defer));
        if (defer) { a.flag_lo = ph.wait_lo; a.flag_hi = ph.wait_hi; a.flag_seq = ph.seq; }
    } else if (distributed) {
        // one plane to each slab neighbour (NCCL send/recv; on the auxiliary stream when the exchange is overlapped)
        cudaStream_t hs;
        MGCR_TRY(dist_halo_begin(ctx, &hs));
        if (has_lo) {
            MGCR_TRY(dist_send(ctx, x, sizeof(c128) * plane, lo, hs));
            MGCR_TRY(dist_recv(ctx, d_halo_lo, sizeof(c128) * plane, lo, hs));
            a.halo_lo = d_halo_lo;
        }
        if (has_hi) {
            MGCR_TRY(dist_send(ctx, x + (n2_local - 1) * plane, sizeof(c128) * plane, hi, hs));
            MGCR_TRY(dist_recv(ctx, d_halo_hi, sizeof(c128) * plane, hi, hs));
            a.halo_hi = d_halo_hi;
        }
        MGCR_TRY(dist_halo_end(ctx));
    }
    if (n_local == 0) return dist_halo_wait(ctx);
    const bool l1_form = ctx->hopping_kernel != 0 || var;   // the tile kernel is unit-hopping only
    int hl_tx = 32;
    static const int hl_tx_max = getenv("MGCR_HL_TX") ? std::min(atoi(getenv("MGCR_HL_TX")), (int)HL_THREADS) : (int)HL_THREADS;   // experiment knob
    while (hl_tx < hl_tx_max && hl_tx * 2 <= a.n0) hl_tx *= 2;
    const int tx = l1_form ? hl_tx : (int)HOP_TX, ty = l1_form ? HL_THREADS / hl_tx : (int)HOP_TY;
    static const int zc_env = getenv("MGCR_HOP_ZC") ? atoi(getenv("MGCR_HOP_ZC")) : 0;   // experiment knob
    const double bytes_per_plane = (apply_bytes() + (diag ? 8. * n_local : 0.) + (bsub ? 16. * n_local : 0.)) / (double)a.n2;
    // TMA-staged form on lattices at least one tile wide with enough sites to fill the machine (small ones are latency-bound
    // and the register-marching form starts faster: 11 us against 13 us at 40 x 33 x 130)
    static const int tma_tile_env = getenv("MGCR_HOP_TILE") ? atoi(getenv("MGCR_HOP_TILE")) : 0;   // experiment knob: 1 = 32 x 16 tile
    auto launch = [&](int64_t z_lo, int64_t z_hi) -> int {   // planes [z_lo, z_hi)
        if (z_hi <= z_lo) return MGCR_OK;
        const int64_t nz = z_hi - z_lo;
        if (tma_form) {
            if (var) return hop_tma_launch<64, 8, true>(ctx, a, z_lo, z_hi, dirac ? "hopping_var_dirac" : "hopping_var", bytes_per_plane * nz);
            const char* nm = dirac ? "hopping_dirac" : "hopping";
            return tma_tile_env == 1 ? hop_tma_launch<32, 16, false>(ctx, a, z_lo, z_hi, nm, bytes_per_plane * nz)
                                     : hop_tma_launch<64, 8, false>(ctx, a, z_lo, z_hi, nm, bytes_per_plane * nz);
        }
        dim3 grid((unsigned)((a.n0 + tx - 1) / tx), (unsigned)((a.n1 + ty - 1) / ty), 1);
        int64_t tiles = (int64_t)grid.x * grid.y;
        int64_t target = (int64_t)ctx->num_sms * 16;
        int64_t nchunks = std::max<int64_t>(1, std::min<int64_t>(nz, (target + tiles - 1) / tiles));
        a.zc = (nz + nchunks - 1) / nchunks;
        if (a.zc < 8 && nz >= 8) a.zc = 8;
        if (zc_env > 0) a.zc = std::min<int64_t>(zc_env, nz);
        grid.z = (unsigned)((nz + a.zc - 1) / a.zc);
        ARG_CHECK(grid.y <= 65535 && grid.z <= 65535, "hopping: lattice too large for the launch grid");
        a.z_lo = z_lo; a.z_hi = z_hi;
        if (var)
            KLAUNCH(ctx, dirac ? "hopping_var_dirac" : "hopping_var", bytes_per_plane * nz, (launch_pdl(ctx, k_hopping_l1<true>, grid, HL_THREADS, 0, a, hl_tx)));
        else if (l1_form)
            KLAUNCH(ctx, dirac ? "hopping_dirac" : "hopping", bytes_per_plane * nz, (launch_pdl(ctx, k_hopping_l1<false>, grid, HL_THREADS, 0, a, hl_tx)));
        else
            KLAUNCH(ctx, dirac ? "hopping_dirac" : "hopping", bytes_per_plane * nz, (k_hopping<<<grid, HOP_TX * HOP_TY, 0, ctx->stream>>>(a)));
        CHECK_LAUNCH();
        return MGCR_OK;
    };
    if (distributed && !ph.on && dist_halo_overlap(ctx) && a.n2 >= 4) {
        // interior planes need no ghost data: they run while the halo planes are in flight
        const int64_t zi0 = has_lo ? 1 : 0, zi1 = has_hi ? a.n2 - 1 : a.n2;
        MGCR_TRY(launch(zi0, zi1));
        MGCR_TRY(dist_halo_wait(ctx));
        MGCR_TRY(launch(0, zi0));
        MGCR_TRY(launch(zi1, a.n2));
    } else {
        MGCR_TRY(dist_halo_wait(ctx));
        MGCR_TRY(launch(0, a.n2));
    }
    return MGCR_OK;
}
int HoppingOp::apply(const c128* x, c128* y) { return run(x, y, 0, cmake(0., 0.), nullptr); }
int HoppingOp::apply_dirac(const c128* x, c128* y, c128 k, const double* diag, const c128* bsub) { return run(x, y, 1, k, diag, bsub); }

// faces: ndim pointers (HOST or DEVICE arrays of n_local doubles, dims order) or NULL for unit hopping
static int hopping_create(mgcr_ctx* ctx, int ndim, const int64_t* dims, const double* const* face, bool face_on_device, mgcr_op** out) {
    ARG_CHECK(ctx && dims && out, "mgcr_hopping_create: NULL argument");
    ARG_CHECK(ndim >= 1 && ndim <= 3, "mgcr_hopping_create: ndim must be 1..3 (got %d)", ndim);
    *out = nullptr;
    HoppingOp* op = new HoppingOp();
    op->kind = OP_HOPPING; op->ctx = ctx; op->ndim = ndim;
    for (int d = 0; d < ndim; d++) {
        if (dims[d] < 1) { delete op; mgcr_set_error("mgcr_hopping_create: dims[%d] < 1", d); return MGCR_ERR_ARG; }
        op->gdims[3 - ndim + d] = dims[d];
    }
    // a 2-D lattice (ny, nx) is traversed as (ny, 1, nx): the kernels stream along their slowest index with the neighbours
    // in that direction held in registers; the operator and the order of the neighbour sum (y-1, x-1, x+1, y+1) are the same
    if (ndim == 2) { op->gdims[0] = dims[0]; op->gdims[1] = 1; op->gdims[2] = dims[1]; }
    int64_t zb = 0, ze = op->gdims[0];
    const int64_t plane = op->gdims[1] * op->gdims[2];
    if (ctx->nranks > 1) {
        if (ndim != 3) { delete op; mgcr_set_error("mgcr_hopping_create: the distributed stencil is 3-D (slabs along dims[0])"); return MGCR_ERR_ARG; }
        int st = mgcr_slab_range(op->gdims[0], ctx->slab_align, ctx->rank, ctx->nranks, &zb, &ze);
        if (st == MGCR_OK && ze <= zb) { mgcr_set_error("mgcr_hopping_create: rank %d owns no plane", ctx->rank); st = MGCR_ERR_ARG; }
        if (st == MGCR_OK) st = dev_alloc_t(ctx, (size_t)plane, &op->d_halo_lo);
        if (st == MGCR_OK) st = dev_alloc_t(ctx, (size_t)plane, &op->d_halo_hi);
        if (st == MGCR_OK) st = p2p_halo_create(ctx, plane, &op->ph);
        if (st != MGCR_OK) { delete op; return st; }
    }
    op->distributed = ctx->nranks > 1;
    op->z_begin = zb; op->n2_local = ze - zb;
    op->n_local = op->n2_local * plane;
    op->n_global = op->gdims[0] * plane;
    if (face) {
        // slot of dims[d] among (z, y, x); the z bonds get one more leading plane: the bond to the lower slab neighbour
        const int slot_of[3][3] = {{2, -1, -1}, {0, 2, -1}, {0, 1, 2}};
        const double* src[3] = {nullptr, nullptr, nullptr};
        for (int d = 0; d < ndim; d++) {
            if (!face[d]) { delete op; mgcr_set_error("mgcr_hopping_create: face[%d] is NULL", d); return MGCR_ERR_ARG; }
            src[slot_of[ndim - 1][d]] = face[d];
        }
        const cudaMemcpyKind kind = face_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        const size_t nl = (size_t)op->n_local;
        int st = dev_alloc_t(ctx, nl + (size_t)plane, &op->d_face[0]);
        if (st == MGCR_OK && src[1]) st = dev_alloc_t(ctx, nl, &op->d_face[1]);
        if (st == MGCR_OK) st = dev_alloc_t(ctx, nl, &op->d_face[2]);
        cudaError_t e = cudaSuccess;
        if (st == MGCR_OK) {
            e = cudaMemsetAsync(op->d_face[0], 0, sizeof(double) * (size_t)plane, ctx->stream);
            if (e == cudaSuccess) e = src[0] ? cudaMemcpyAsync(op->d_face[0] + plane, src[0], sizeof(double) * nl, kind, ctx->stream)
                                             : cudaMemsetAsync(op->d_face[0] + plane, 0, sizeof(double) * nl, ctx->stream);
            if (e == cudaSuccess && src[1]) e = cudaMemcpyAsync(op->d_face[1], src[1], sizeof(double) * nl, kind, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(op->d_face[2], src[2], sizeof(double) * nl, kind, ctx->stream);
        }
        if (st == MGCR_OK && e == cudaSuccess && op->distributed) {
            // the bond between my first plane and the lower neighbour's last plane lives with the lower neighbour
            st = dist_group_begin(ctx);
            if (st == MGCR_OK && ctx->rank + 1 < ctx->nranks) st = dist_send(ctx, op->d_face[0] + (size_t)op->n2_local * plane, sizeof(double) * plane, ctx->rank + 1, ctx->stream);
            if (st == MGCR_OK && ctx->rank > 0) st = dist_recv(ctx, op->d_face[0], sizeof(double) * plane, ctx->rank - 1, ctx->stream);
            if (st == MGCR_OK) st = dist_group_end(ctx);
        }
        // no bond leaves the lattice (Dirichlet): whatever the caller stored for the top plane's upward bonds is dropped
        if (st == MGCR_OK && e == cudaSuccess && ctx->rank + 1 == ctx->nranks)
            e = cudaMemsetAsync(op->d_face[0] + (size_t)op->n2_local * plane, 0, sizeof(double) * (size_t)plane, ctx->stream);
        if (st == MGCR_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (st == MGCR_OK && e != cudaSuccess) { mgcr_set_error("mgcr_hopping_create: bond upload: %s", cudaGetErrorString(e)); st = MGCR_ERR_CUDA; }
        if (st != MGCR_OK) { delete op; return st; }
        op->var = true;
    }
    *out = op;
    return MGCR_OK;
}

extern "C" int mgcr_hopping_create(mgcr_ctx* ctx, int ndim, const int64_t* dims, const double* const* h_face, mgcr_op** out) {
    return hopping_create(ctx, ndim, dims, h_face, false, out);
}
extern "C" int mgcr_hopping_create_dev(mgcr_ctx* ctx, int ndim, const int64_t* dims, const double* const* d_face, mgcr_op** out) {
    ARG_CHECK(d_face, "mgcr_hopping_create_dev: NULL bond arrays");
    return hopping_create(ctx, ndim, dims, d_face, true, out);
}

// ----------------------------------------------------------------------------------------------------------
// DiracOp = diag - k D
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) k_dirac_combine(int64_t n, c128 k, const double* __restrict__ diag, const c128* __restrict__ x,
                                                               c128* y /* in: D x, out: diag.x - k D x */) {
    PDL_ENTRY();
    GRID_STRIDE(i, n) {
        c128 xr = ld_stream(x + i);
        if (diag) { double d = __ldg(diag + i); xr = cmake(d * xr.x, d * xr.y); }
        st_stream(y + i, csub(xr, cmul(k, ld_plain(y + i))));
    }
}

DiracOp::~DiracOp() { dev_free(ctx, d_diag); }

int DiracOp::apply(const c128* x, c128* y) {
    if (D->kind == OP_SELL) return static_cast<SellOp*>(D)->apply_dirac(x, y, k, d_diag);
    if (D->kind == OP_HOPPING) return static_cast<HoppingOp*>(D)->apply_dirac(x, y, k, d_diag);
    MGCR_TRY(D->apply(x, y));
    if (n_local == 0) return MGCR_OK;
    KLAUNCH(ctx, "dirac_combine", 48. * n_local, (launch_pdl(ctx, k_dirac_combine, stream_grid(ctx, n_local, 8), RED_THREADS, 0, n_local, k, (const double*)d_diag, x, y)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

int DiracOp::apply_residual(const c128* x, const c128* b, c128* r) {
    ARG_CHECK(b != r, "residual: right-hand side and output alias");
    if (D->kind == OP_SELL) return static_cast<SellOp*>(D)->apply_dirac(x, r, k, d_diag, b);
    if (D->kind == OP_HOPPING) return static_cast<HoppingOp*>(D)->apply_dirac(x, r, k, d_diag, b);
    return mgcr_op::apply_residual(x, b, r);
}

int vec_axpy(mgcr_ctx* ctx, int64_t n, c128 s, const c128* b, const c128* a, c128* out);
int mgcr_op::apply_residual(const c128* x, const c128* b, c128* r) {
    MGCR_TRY(apply(x, r));
    return vec_axpy(ctx, n_local, cmake(-1., 0.), r, b, r);   // r = b + (-1) r, exactly b - r
}

static int dirac_create(mgcr_ctx* ctx, mgcr_op* D, double k_re, double k_im, const double* h_diag, bool diag_on_device, mgcr_op** out) {
    ARG_CHECK(ctx && D && out, "mgcr_dirac_create: NULL argument");
    *out = nullptr;
    DiracOp* op = new DiracOp();
    op->kind = OP_DIRAC; op->ctx = ctx; op->D = D; op->k = cmake(k_re, k_im);
    op->n_local = D->n_local; op->n_global = D->n_global; op->distributed = D->distributed;
    if (h_diag) {
        int st = dev_alloc_t(ctx, (size_t)op->n_local, &op->d_diag);
        if (st != MGCR_OK) { delete op; return st; }
        cudaError_t e = cudaMemcpyAsync(op->d_diag, h_diag, sizeof(double) * op->n_local, diag_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { delete op; mgcr_set_error("dirac diag upload: %s", cudaGetErrorString(e)); return MGCR_ERR_CUDA; }
    }
    *out = op;
    return MGCR_OK;
}

extern "C" int mgcr_dirac_create(mgcr_ctx* ctx, mgcr_op* D, double k_re, double k_im, const double* h_diag, mgcr_op** out) {
    return dirac_create(ctx, D, k_re, k_im, h_diag, false, out);
}
extern "C" int mgcr_dirac_create_dev(mgcr_ctx* ctx, mgcr_op* D, double k_re, double k_im, const double* d_diag, mgcr_op** out) {
    return dirac_create(ctx, D, k_re, k_im, d_diag, true, out);
}

extern "C" int mgcr_dirac_set_k(mgcr_op* op, double k_re, double k_im) {
    ARG_CHECK(op && op->kind == OP_DIRAC, "mgcr_dirac_set_k: not a DiracOp");
    static_cast<DiracOp*>(op)->k = cmake(k_re, k_im);
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// block-CSR apply: one thread per (block row R, row r inside the block).  Blocks are stored column-major, so for
// each column c the ne threads of a block row read ne consecutive c128 -- coalesced -- and the matching x element is
// a warp-broadcast L1 hit.  Accumulation follows the reference: per block o = sum_c m[r][c] x[c] (sequential), then
// value += o in block order (HierarchicalSparse.h:135-147 with Dense::operator(), Operator.h:159-173).
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_blockcsr_apply(int64_t nb, int ne, const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                                                        const c128* __restrict__ bval, const c128* __restrict__ x, const c128* __restrict__ ghost,
                                                        int64_t nb_local_cols, const c128* __restrict__ bsub, c128* __restrict__ y,
                                                        int64_t row0) {
    const int64_t t = row0 * ne + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // block rows [row0, nb) of this launch
    const int64_t R = t / ne;
    const int r = (int)(t - R * ne);
    if (R >= nb) return;
    c128 value = cmake(0., 0.);
    const int lb = __ldg(brow + R), le = __ldg(brow + R + 1);
    for (int l = lb; l < le; l++) {
        const int64_t bc = __ldg(bcol + l);
        const c128* xb = bc < nb_local_cols ? x + bc * ne : ghost + (bc - nb_local_cols) * ne;
        const c128* m = bval + (int64_t)l * ne * ne + r;
        c128 o = cmake(0., 0.);
#pragma unroll 4
        for (int c = 0; c < ne; c++) o = cadd(o, cmul(ld_stream(m + (int64_t)c * ne), __ldg(xb + c)));
        value = cadd(value, o);
    }
    if (bsub) value = csub(__ldg(bsub + t), value);
    st_stream(y + t, value);
}

// The same for a compile-time block size (the multigrid levels: ne = n_eigen or 2 n_eigen): the column loop is unrolled
// and two blocks are in flight per thread, i.e. 2*NE independent 128-bit matrix loads per thread instead of 4.
template <int NE>
__global__ void __launch_bounds__(256) k_blockcsr_apply_ne(int64_t nb, const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                                                           const c128* __restrict__ bval, const c128* __restrict__ x, const c128* __restrict__ ghost,
                                                           int64_t nb_local_cols, const c128* __restrict__ bsub, c128* __restrict__ y,
                                                           int64_t row0) {
    PDL_ENTRY();
    const int64_t t = row0 * NE + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // block rows [row0, nb) of this launch
    const int64_t R = t / NE;
    const int r = (int)(t - R * NE);
    if (R >= nb) return;
    c128 value = cmake(0., 0.);
    const int lb = __ldg(brow + R), le = __ldg(brow + R + 1);
    int l = lb;
    for (; l + 1 < le; l += 2) {
        const int64_t bc0 = __ldg(bcol + l), bc1 = __ldg(bcol + l + 1);
        const c128* xb0 = bc0 < nb_local_cols ? x + bc0 * NE : ghost + (bc0 - nb_local_cols) * NE;
        const c128* xb1 = bc1 < nb_local_cols ? x + bc1 * NE : ghost + (bc1 - nb_local_cols) * NE;
        const c128* m0 = bval + (int64_t)l * NE * NE + r;
        c128 a0[NE], a1[NE], x0[NE], x1[NE];
#pragma unroll
        for (int c = 0; c < NE; c++) { a0[c] = ld_stream(m0 + c * NE); a1[c] = ld_stream(m0 + NE * NE + c * NE); }
#pragma unroll
        for (int c = 0; c < NE; c++) { x0[c] = __ldg(xb0 + c); x1[c] = __ldg(xb1 + c); }
        c128 o0 = cmake(0., 0.), o1 = cmake(0., 0.);
#pragma unroll
        for (int c = 0; c < NE; c++) { o0 = cadd(o0, cmul(a0[c], x0[c])); o1 = cadd(o1, cmul(a1[c], x1[c])); }
        value = cadd(value, o0);
        value = cadd(value, o1);
    }
    if (l < le) {
        const int64_t bc = __ldg(bcol + l);
        const c128* xb = bc < nb_local_cols ? x + bc * NE : ghost + (bc - nb_local_cols) * NE;
        const c128* m = bval + (int64_t)l * NE * NE + r;
        c128 o = cmake(0., 0.);
#pragma unroll
        for (int c = 0; c < NE; c++) o = cadd(o, cmul(ld_stream(m + c * NE), __ldg(xb + c)));
        value = cadd(value, o);
    }
    if (bsub) value = csub(__ldg(bsub + t), value);
    st_stream(y + t, value);
}

// Sliced streaming image (see BlockCsrOp) applied by a persistent kernel: one CTA per SM, warp i of the CTA owns ring stage i; slices
// are dealt round-robin over all warps of the grid.  A slice (a contiguous blob: per block slot NE*32
// values + 32/NE columns) is fetched into the stage by ONE 1-D bulk copy (cp.async.bulk completing on the stage's
// mbarrier) that the warp issues itself as soon as it has finished with the previous blob.  While the copy is in flight
// the warp gathers the x blocks of that slice through L1/L2 -- the column indices were read one slice ahead -- so the
// DRAM latency of the matrix stream and the L2 latency of the gathers overlap instead of adding up (the first ring
// version, with a producer thread and gathers after the arrival, spent 53 % of its stall samples waiting for refills:
// profiles/r01_ncu_blockcsr_ring.md).  The DRAM stream is as deep as the ring: nst blobs per SM.
// Accumulation order per row is that of k_blockcsr_apply_ne; the padding blocks add exact zeros.
template <int NE> struct RingCfg {
    static constexpr int PRE = NE == 8 ? 2 : 4;                         // block slots whose x gathers are issued before the blob arrives
    static constexpr int MAX_STAGES = NE == 2 ? 26 : 14;                 // = warps per CTA (ne = 2: 7.6 KB blobs, more of them in flight)
    static constexpr int MAX_THREADS = 32 * MAX_STAGES;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int NE>
__global__ void __launch_bounds__(RingCfg<NE>::MAX_THREADS, 1) k_blockcsr_ring(int64_t nb, int64_t nslices, const int64_t* __restrict__ sl_ptr,
                                                                                const unsigned char* __restrict__ blob, const c128* __restrict__ x,
                                                                                const c128* __restrict__ ghost, int64_t nb_local_cols,
                                                                                const c128* __restrict__ bsub, c128* __restrict__ y, int nst, int stage_bytes,
                                                                                const uint32_t* flag_lo, const uint32_t* flag_hi, uint32_t flag_seq) {
    constexpr int S = 32 / NE;
    constexpr int SLOT = NE * 512 + S * 4;                 // bytes per block slot: NE columns x 32 lanes of c128, then S int32 columns
    constexpr int PRE = RingCfg<NE>::PRE;
    extern __shared__ unsigned char ring_raw[];
    __shared__ __align__(8) uint64_t full[RingCfg<NE>::MAX_STAGES];
    unsigned char* ring = (unsigned char*)(((uintptr_t)ring_raw + 127) & ~(uintptr_t)127);
    // slices are dealt round-robin over (CTA, warp): at any time the whole machine streams one window of gridDim.x * nst
    // consecutive slices (148 distant sequential streams, one per CTA, reached only 5.6 TB/s)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t s_end = nslices, stride = (int64_t)gridDim.x * nst;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; i++) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    PDL_ENTRY();
    const int rin = lane / NE;
    // deferred halo wait (p2p.cu): the ghost blocks may still be on their way; a lane polls the neighbours' flags right before
    // its first ghost access, so only the warps whose slices couple to another rank's aggregates ever wait
    bool ghost_ok = (flag_lo == nullptr && flag_hi == nullptr);
    auto ghost_ready = [&]() {
        if (ghost_ok) return;
        if (flag_lo) p2p_flag_wait(flag_lo, flag_seq);
        if (flag_hi) p2p_flag_wait(flag_hi, flag_seq);
        ghost_ok = true;
    };
    unsigned char* stage = ring + (size_t)warp * stage_bytes;
    // Residual form: the 32 right-hand-side elements of the slice ride through the ring as a second bulk copy into the last 512
    // bytes of the stage, completing on the same barrier (read per thread after the products they were one more exposed DRAM round
    // trip per slice: 773 against 695 us per launch on level 1 of the 512^3 solve; no registers are free to prefetch them).
    const c128* bst = (const c128*)(stage + stage_bytes - 512);
    auto fetch = [&](int64_t sl, int64_t base, int w) {    // lane 0: start the copy of a slice into this warp's stage
        uint32_t bb = 0;
        if (bsub) bb = (uint32_t)min((int64_t)32, nb * NE - sl * 32) * 16u;
        mbar_expect_tx(&full[warp], (uint32_t)w * SLOT + bb);
        if (w) tma_load_1d(stage, blob + base * SLOT, (uint32_t)w * SLOT, &full[warp]);
        if (bb) tma_load_1d((void*)bst, bsub + sl * 32, bb, &full[warp]);
    };
    auto load_cols = [&](int64_t base, int w, int32_t (&col)[PRE]) {
#pragma unroll
        for (int i = 0; i < PRE; i++) col[i] = i < w ? __ldg((const int32_t*)(blob + (base + i) * SLOT + NE * 512) + rin) : 0;
    };
    int64_t s = (int64_t)blockIdx.x * nst + warp;
    int64_t base = 0; int w = 0;
    int32_t col[PRE];
    if (s < s_end) {
        base = __ldg(sl_ptr + s);
        w = (int)(__ldg(sl_ptr + s + 1) - base);
        if (lane == 0) fetch(s, base, w);
    }
    load_cols(base, w, col);
    uint32_t parity = 0;
    for (; s < s_end; s += stride) {
        // x blocks of the first PRE slots: in flight together with the blob
        c128 xv[PRE][NE];
#pragma unroll
        for (int i = 0; i < PRE; i++) {
            if (i < w) {
                const int64_t bc = col[i];
                if (bc >= nb_local_cols) ghost_ready();
                const c128* xb = bc < nb_local_cols ? x + bc * NE : ghost + (bc - nb_local_cols) * NE;
#pragma unroll
                for (int c = 0; c < NE; c++) xv[i][c] = __ldg(xb + c);
            }
        }
        // the slice after this one: offsets and column indices (consumed by the next trip)
        const int64_t sn = s + stride;
        int64_t basen = 0; int wn = 0;
        if (sn < s_end) { basen = __ldg(sl_ptr + sn); wn = (int)(__ldg(sl_ptr + sn + 1) - basen); }
        int32_t coln[PRE];
        load_cols(basen, wn, coln);
        mbar_wait(&full[warp], parity);
        parity ^= 1;
        c128 value = cmake(0., 0.);
#pragma unroll
        for (int i = 0; i < PRE; i++) {
            if (i < w) {
                const c128* m = (const c128*)(stage + (size_t)i * SLOT) + lane;
                c128 o = cmake(0., 0.);
#pragma unroll
                for (int c = 0; c < NE; c++) o = cadd(o, cmul(m[c * 32], xv[i][c]));
                value = cadd(value, o);
            }
        }
        for (int l0 = PRE; l0 < w; l0 += PRE) {            // the remaining slots in batches, columns from shared memory
#pragma unroll
            for (int i = 0; i < PRE; i++) {
                if (l0 + i < w) {
                    const int64_t bc = ((const int32_t*)(stage + (size_t)(l0 + i) * SLOT + NE * 512))[rin];
                    if (bc >= nb_local_cols) ghost_ready();
                    const c128* xb = bc < nb_local_cols ? x + bc * NE : ghost + (bc - nb_local_cols) * NE;
#pragma unroll
                    for (int c = 0; c < NE; c++) xv[i][c] = __ldg(xb + c);
                }
            }
#pragma unroll
            for (int i = 0; i < PRE; i++) {
                if (l0 + i < w) {
                    const c128* m = (const c128*)(stage + (size_t)(l0 + i) * SLOT) + lane;
                    c128 o = cmake(0., 0.);
#pragma unroll
                    for (int c = 0; c < NE; c++) o = cadd(o, cmul(m[c * 32], xv[i][c]));
                    value = cadd(value, o);
                }
            }
        }
        const int64_t t = s * 32 + lane;
        if (bsub && t < nb * NE) value = csub(bst[lane], value);
        __syncwarp();                                      // every lane is done with the stage
        if (lane == 0 && sn < s_end) { fence_proxy_async_smem(); fetch(sn, basen, wn); }
        if (t < nb * NE) st_stream(y + t, value);
        base = basen; w = wn;
#pragma unroll
        for (int i = 0; i < PRE; i++) col[i] = coln[i];
    }
    // the exchange is complete when this kernel is: somebody has to have seen both flags even if no row needed a ghost
    if (blockIdx.x == 0 && threadIdx.x == 0) ghost_ready();
}

static __global__ void __launch_bounds__(256) k_slice_width(int64_t nb, int64_t nslices, int S, const int32_t* __restrict__ brow, int64_t* __restrict__ width,
                                                            unsigned long long* __restrict__ wmax) {
    int wm = 0;
    GRID_STRIDE(s, nslices) {
        int w = 0;
        for (int q = 0; q < S; q++) {
            const int64_t R = s * S + q;
            if (R < nb) w = max(w, brow[R + 1] - brow[R]);
        }
        width[s] = w;
        wm = max(wm, w);
    }
    if (wm) atomicMax(wmax, (unsigned long long)wm);
}
// one warp per slice copies its rows from the assembly layout [l][c][r] into the slice's blob; padding = zero blocks
// whose column is the row's own block (always addressable)
static __global__ void __launch_bounds__(256) k_slice_fill(int64_t nb, int64_t nslices, int ne, const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                                                           const c128* __restrict__ bval, const int64_t* __restrict__ sl_ptr, unsigned char* __restrict__ blob) {
    const int S = 32 / ne;
    const int slot = ne * 512 + S * 4;
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= nslices) return;
    const int lane = threadIdx.x & 31, rin = lane / ne, r = lane - rin * ne;
    const int64_t R = s * S + rin;
    const int64_t base = sl_ptr[s];
    const int w = (int)(sl_ptr[s + 1] - base);
    const int first = R < nb ? brow[R] : 0, cnt = R < nb ? brow[R + 1] - brow[R] : 0;
    for (int l = 0; l < w; l++) {
        const bool on = l < cnt;
        unsigned char* p = blob + (base + l) * slot;
        if (r == 0) ((int32_t*)(p + ne * 512))[rin] = on ? bcol[first + l] : (int32_t)min(R, nb - 1);
        for (int c = 0; c < ne; c++)
            ((c128*)p)[c * 32 + lane] = on ? bval[((int64_t)(first + l) * ne + c) * ne + r] : cmake(0., 0.);
    }
}

int BlockCsrOp::build_sliced() {
    sliced = -1;
    static const int enabled = getenv("MGCR_BLOCKCSR_SLICED") ? atoi(getenv("MGCR_BLOCKCSR_SLICED")) : 1;   // experiment knob
    if (!enabled || !(ne == 2 || ne == 4 || ne == 8) || nb * ne < ctx->blockcsr_ring_rows) return MGCR_OK;   // small operators are latency-bound anyway
    const int S = 32 / ne;
    const int slot = ne * 512 + S * 4;
    nslices = (nb + S - 1) / S;
    // slice widths, their maximum and their exclusive scan stay on the device; 16 bytes come back (total slots, widest slice)
    int64_t *d_tot = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)nslices + 1, &d_sl_ptr));
    MGCR_TRY(dev_alloc_t(ctx, 2, &d_tot));
    CUDA_TRY(cudaMemsetAsync(d_tot, 0, 2 * sizeof(int64_t), ctx->stream));
    k_slice_width<<<stream_grid(ctx, nslices, 8), RED_THREADS, 0, ctx->stream>>>(nb, nslices, S, d_brow, d_sl_ptr, (unsigned long long*)(d_tot + 1));
    CHECK_LAUNCH();
    MGCR_TRY(device_exclusive_scan_i64(ctx, d_sl_ptr, nslices, d_tot));
    CUDA_TRY(cudaMemcpyAsync(d_sl_ptr + nslices, d_tot, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    int64_t tot[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(tot, d_tot, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, d_tot);
    sl_slots = tot[0];
    const int64_t wmax = tot[1];
    // ring: one stage (= the widest slice) per consumer warp, as many as fit 200 KB, at most 14 (26 for ne = 2)
    sl_stage_bytes = (int)(((wmax * slot + 127) / 128) * 128) + 512;   // + the slice's 32 right-hand-side elements (residual form)
    static const int nst_env = getenv("MGCR_BLOCKCSR_STAGES") ? atoi(getenv("MGCR_BLOCKCSR_STAGES")) : 0;   // experiment knob
    const int max_stages = ne == 2 ? RingCfg<2>::MAX_STAGES : RingCfg<4>::MAX_STAGES;
    sl_stages = sl_stage_bytes ? (int)std::min<int64_t>(nst_env > 0 ? std::min(nst_env, max_stages) : max_stages, (212 * 1024) / sl_stage_bytes) : 0;
    // very ragged rows (> 25 % padding) or blobs too large for a useful ring: the assembly layout serves better
    if ((double)sl_slots * S > 1.25 * (double)nnzb || sl_stages < 4) { dev_free(ctx, d_sl_ptr); d_sl_ptr = nullptr; return MGCR_OK; }
    MGCR_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(sl_slots, 1) * slot, (void**)&d_sl_blob));
    k_slice_fill<<<(unsigned)((nslices * 32 + 255) / 256), 256, 0, ctx->stream>>>(nb, nslices, ne, d_brow, d_bcol, d_bval, d_sl_ptr, d_sl_blob);
    CHECK_LAUNCH();
    sliced = 1;
    return MGCR_OK;
}

void BlockCsrOp::drop_assembly_values() {
    if (sliced == 1 && d_bval && n_local > ctx->small_gcr_rows) { dev_free(ctx, d_bval); d_bval = nullptr; }
}

template <int NE>
static int blockcsr_ring_launch(BlockCsrOp* op, const c128* x, const c128* ghost, const c128* bsub, c128* y) {
    mgcr_ctx* ctx = op->ctx;
    const size_t smem = (size_t)op->sl_stages * op->sl_stage_bytes + 128;
    MGCR_TRY(ensure_dyn_smem(ctx, (const void*)k_blockcsr_ring<NE>, 213 * 1024));
    const int threads = 32 * op->sl_stages;
    const unsigned grid = (unsigned)std::min<int64_t>(ctx->num_sms, (op->nslices + op->sl_stages - 1) / op->sl_stages);
    const PeerHalo* ph = (op->halo && op->halo_deferred) ? &op->halo->ph : nullptr;
    launch_pdl(ctx, k_blockcsr_ring<NE>, grid, threads, smem, op->nb, op->nslices, (const int64_t*)op->d_sl_ptr, (const unsigned char*)op->d_sl_blob, x, ghost,
               op->n_local / op->ne, bsub, y, op->sl_stages, op->sl_stage_bytes, ph ? ph->wait_lo : (const uint32_t*)nullptr,
               ph ? ph->wait_hi : (const uint32_t*)nullptr, ph ? ph->seq : 0u);
    return MGCR_OK;
}

BlockCsrOp::~BlockCsrOp() {
    dev_free(ctx, d_brow); dev_free(ctx, d_bcol); dev_free(ctx, d_bval);
    dev_free(ctx, d_sl_ptr); dev_free(ctx, d_sl_blob);
    halo_free(ctx, halo);
}

int BlockCsrOp::apply(const c128* x, c128* y) { return run(x, y, nullptr); }
int BlockCsrOp::apply_residual(const c128* x, const c128* b, c128* r) {
    ARG_CHECK(b != r, "residual: right-hand side and output alias");
    return run(x, r, b);
}

int BlockCsrOp::run(const c128* x, c128* y, const c128* bsub) {
    ARG_CHECK(x != y, "operator apply: input and output alias");
    const c128* ghost = nullptr;
    if (sliced == 0 && nb > 0) MGCR_TRY(build_sliced());
    // the ring kernel waits for the neighbours' ghost blocks itself (peer-memory halo only)
    static const int defer_env = getenv("MGCR_HALO_DEFER") ? atoi(getenv("MGCR_HALO_DEFER")) : 1;
    halo_deferred = halo && halo->ph.on && !halo->d_send_idx && sliced == 1 && nb > 0 && defer_env;
    if (halo) { MGCR_TRY(halo_exchange(ctx, halo, x, halo_deferred)); ghost = halo->ghost_cur; }
    if (nb == 0) return dist_halo_wait(ctx);
    const double bytes_per_row = (apply_bytes() + (bsub ? 16. * n_local : 0.)) / (double)nb;
    if (sliced == 1 && !(halo && !halo->ph.on && dist_halo_overlap(ctx))) {
        MGCR_TRY(dist_halo_wait(ctx));
        ProfScope ps_(ctx, "blockcsr_apply", bytes_per_row * nb);
        switch (ne) {
            case 2: MGCR_TRY(blockcsr_ring_launch<2>(this, x, ghost, bsub, y)); break;
            case 4: MGCR_TRY(blockcsr_ring_launch<4>(this, x, ghost, bsub, y)); break;
            default: MGCR_TRY(blockcsr_ring_launch<8>(this, x, ghost, bsub, y)); break;
        }
        CHECK_LAUNCH();
        return MGCR_OK;
    }
    auto launch = [&](int64_t r0, int64_t r1) -> int {   // block rows [r0, r1)
        if (r1 <= r0) return MGCR_OK;
        const int grid = (int)(((r1 - r0) * ne + 255) / 256);
        ProfScope ps_(ctx, "blockcsr_apply", bytes_per_row * (r1 - r0));
        switch (ne) {
            case 2: launch_pdl(ctx, k_blockcsr_apply_ne<2>, grid, 256, 0, r1, (const int32_t*)d_brow, (const int32_t*)d_bcol, (const c128*)d_bval, x, ghost, n_local / ne, bsub, y, r0); break;
            case 4: launch_pdl(ctx, k_blockcsr_apply_ne<4>, grid, 256, 0, r1, (const int32_t*)d_brow, (const int32_t*)d_bcol, (const c128*)d_bval, x, ghost, n_local / ne, bsub, y, r0); break;
            case 8: launch_pdl(ctx, k_blockcsr_apply_ne<8>, grid, 256, 0, r1, (const int32_t*)d_brow, (const int32_t*)d_bcol, (const c128*)d_bval, x, ghost, n_local / ne, bsub, y, r0); break;
            default: k_blockcsr_apply<<<grid, 256, 0, ctx->stream>>>(r1, ne, d_brow, d_bcol, d_bval, x, ghost, n_local / ne, bsub, y, r0);
        }
        return MGCR_OK;
    };
    if (halo && !halo->ph.on && dist_halo_overlap(ctx) && nb > halo_rows_lo + halo_rows_hi) {
        // only the first / last plane of aggregates has ghost columns: everything else runs while the halo is in flight
        MGCR_TRY(launch(halo_rows_lo, nb - halo_rows_hi));
        MGCR_TRY(dist_halo_wait(ctx));
        MGCR_TRY(launch(0, halo_rows_lo));
        MGCR_TRY(launch(nb - halo_rows_hi, nb));
    } else {
        MGCR_TRY(dist_halo_wait(ctx));
        MGCR_TRY(launch(0, nb));
    }
    CHECK_LAUNCH();
    return MGCR_OK;
}

// host block-CSR with ROW-major blocks (the reference's Dense layout) -> device compact column-major
int blockcsr_build(mgcr_ctx* ctx, int64_t nb, int64_t nb_cols_addressable, int ne, const int64_t* brow, const int64_t* bcol,
                   const mgcr_c128* bval, BlockCsrOp* op) {
    ARG_CHECK(nb_cols_addressable * ne < (int64_t)INT32_MAX && brow[nb] < (int64_t)INT32_MAX, "block-CSR upload: index exceeds int32");
    std::vector<int32_t> hrow((size_t)nb + 1, 0), hcol;
    std::vector<c128> hval;
    hcol.reserve((size_t)brow[nb]);
    hval.reserve((size_t)brow[nb] * ne * ne);
    const size_t bsz = (size_t)ne * ne;
    for (int64_t R = 0; R < nb; R++) {
        for (int64_t l = brow[R]; l < brow[R + 1]; l++) {
            const mgcr_c128* m = bval + (size_t)l * bsz;
            bool nz = false;
            for (size_t q = 0; q < bsz; q++) if (m[q].re != 0. || m[q].im != 0.) { nz = true; break; }
            if (!nz) continue;
            ARG_CHECK(bcol[l] >= 0 && bcol[l] < nb_cols_addressable, "block-CSR upload: block column %lld out of range", (long long)bcol[l]);
            hcol.push_back((int32_t)bcol[l]);
            size_t base = hval.size();
            hval.resize(base + bsz);
            for (int r = 0; r < ne; r++) for (int c = 0; c < ne; c++) hval[base + (size_t)c * ne + r] = cmake(m[(size_t)r * ne + c].re, m[(size_t)r * ne + c].im);
        }
        hrow[R + 1] = (int32_t)hcol.size();
    }
    op->nb = nb; op->ne = ne; op->nnzb = (int64_t)hcol.size(); op->nb_cols = nb_cols_addressable;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)nb + 1, &op->d_brow));
    MGCR_TRY(dev_alloc_t(ctx, hcol.size(), &op->d_bcol));
    MGCR_TRY(dev_alloc_t(ctx, hval.size(), &op->d_bval));
    CUDA_TRY(cudaMemcpyAsync(op->d_brow, hrow.data(), sizeof(int32_t) * hrow.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!hcol.empty()) {
        CUDA_TRY(cudaMemcpyAsync(op->d_bcol, hcol.data(), sizeof(int32_t) * hcol.size(), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(op->d_bval, hval.data(), sizeof(c128) * hval.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return MGCR_OK;
}

extern "C" int mgcr_blockcsr_create(mgcr_ctx* ctx, int64_t nb, int ne, const int64_t* brow, const int64_t* bcol, const mgcr_c128* bval, mgcr_op** out) {
    ARG_CHECK(ctx && out && brow && nb >= 0 && ne >= 1, "mgcr_blockcsr_create: bad argument");
    ARG_CHECK(ctx->nranks == 1, "mgcr_blockcsr_create: single-GPU entry point (distributed coarse operators are built by mgcr_mg_create)");
    *out = nullptr;
    BlockCsrOp* op = new BlockCsrOp();
    op->kind = OP_BLOCKCSR; op->ctx = ctx; op->n_local = nb * ne; op->n_global = nb * ne;
    int st = blockcsr_build(ctx, nb, nb, ne, brow, bcol, bval, op);
    if (st != MGCR_OK) { delete op; return st; }
    *out = op;
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// caller-implemented operators
// ----------------------------------------------------------------------------------------------------------
int CallbackOp::apply(const c128* x, c128* y) {
    ARG_CHECK(x != y, "operator apply: input and output alias");
    ctx->launches++;
    int st = fn(user, (const mgcr_c128*)x, (mgcr_c128*)y);
    if (st != MGCR_OK) mgcr_set_error("callback operator returned status %d", st);
    return st;
}

extern "C" int mgcr_callback_op_create(mgcr_ctx* ctx, int64_t n, mgcr_apply_fn fn, void* user, mgcr_op** out) {
    ARG_CHECK(ctx && fn && out && n >= 0, "mgcr_callback_op_create: bad argument");
    CallbackOp* op = new CallbackOp();
    op->kind = OP_CALLBACK; op->ctx = ctx; op->fn = fn; op->user = user;
    op->n_local = n; op->n_global = n;
    *out = op;
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// generic entry points
// ----------------------------------------------------------------------------------------------------------
extern "C" int mgcr_op_apply(mgcr_ctx* ctx, mgcr_op* op, const mgcr_c128* x, mgcr_c128* y) {
    ARG_CHECK(ctx && op && x && y, "mgcr_op_apply: NULL argument");
    return op->apply((const c128*)x, (c128*)y);
}
extern "C" int mgcr_op_residual(mgcr_ctx* ctx, mgcr_op* op, const mgcr_c128* x, const mgcr_c128* b, mgcr_c128* r) {
    ARG_CHECK(ctx && op && x && b && r, "mgcr_op_residual: NULL argument");
    return op->apply_residual((const c128*)x, (const c128*)b, (c128*)r);
}
extern "C" int mgcr_op_dim(mgcr_op* op, int64_t* n_local, int64_t* n_global) {
    ARG_CHECK(op, "mgcr_op_dim: NULL operator");
    if (n_local) *n_local = op->n_local;
    if (n_global) *n_global = op->n_global;
    return MGCR_OK;
}
extern "C" int mgcr_op_apply_bytes(mgcr_op* op, double* bytes) {
    ARG_CHECK(op && bytes, "mgcr_op_apply_bytes: NULL argument");
    *bytes = op->apply_bytes();
    return MGCR_OK;
}
extern "C" int mgcr_op_destroy(mgcr_op* op) {
    if (!op) return MGCR_OK;
    mgcr_ctx* ctx = op->ctx;
    delete op;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// ghost exchange for list-based halos (distributed CSR / block-CSR)
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) k_pack(int64_t n_items, int elem, const int32_t* __restrict__ idx, const c128* __restrict__ x,
                                                      c128* __restrict__ buf) {
    PDL_ENTRY();
    GRID_STRIDE(t, n_items * elem) {
        int64_t it = t / elem; int e = (int)(t - it * elem);
        buf[t] = x[(int64_t)idx[it] * elem + e];
    }
}

int halo_exchange(mgcr_ctx* ctx, HaloPlan* h, const c128* x, bool defer) {
    if (!h || h->npeers == 0) return MGCR_OK;
    h->ghost_cur = h->d_ghost;
    if (h->ph.on && !h->d_send_idx) {
        // slab neighbours, contiguous ranges: stored straight into the neighbours' receive areas over NVLink (p2p.cu)
        const c128 *send_lo = nullptr, *send_hi = nullptr, *recv_lo = nullptr, *recv_hi = nullptr;
        for (int p = 0; p < h->npeers; p++) {
            if (h->peer[p] == ctx->rank - 1) send_lo = x + h->send_start[p] * h->elem;
            if (h->peer[p] == ctx->rank + 1) send_hi = x + h->send_start[p] * h->elem;
        }
        MGCR_TRY(p2p_halo_exchange(ctx, &h->ph, send_lo, send_hi, &recv_lo, &recv_hi, defer));
        h->ghost_cur = recv_lo ? recv_lo : recv_hi;   // [lower plane][upper plane] contiguous; without a lower neighbour the upper one comes first
        return MGCR_OK;
    }
    int64_t n_send = h->send_off[h->npeers];
    if (h->d_send_idx && n_send > 0) {
        KLAUNCH(ctx, "halo_pack", 36. * n_send * h->elem, (k_pack<<<stream_grid(ctx, n_send * h->elem, 4), RED_THREADS, 0, ctx->stream>>>(n_send, h->elem, h->d_send_idx, x, h->d_send_buf)));
        CHECK_LAUNCH();
    }
    cudaStream_t hs;
    MGCR_TRY(dist_halo_begin(ctx, &hs));
    for (int p = 0; p < h->npeers; p++) {
        int64_t ns = h->send_off[p + 1] - h->send_off[p], nr = h->recv_off[p + 1] - h->recv_off[p];
        if (ns > 0) {
            const c128* src = h->d_send_idx ? h->d_send_buf + h->send_off[p] * h->elem : x + h->send_start[p] * h->elem;
            MGCR_TRY(dist_send(ctx, src, sizeof(c128) * ns * h->elem, h->peer[p], hs));
        }
        if (nr > 0) MGCR_TRY(dist_recv(ctx, h->d_ghost + h->recv_off[p] * h->elem, sizeof(c128) * nr * h->elem, h->peer[p], hs));
    }
    MGCR_TRY(dist_halo_end(ctx));
    return MGCR_OK;   // the caller issues dist_halo_wait() before the work that reads the ghosts
}

void halo_free(mgcr_ctx* ctx, HaloPlan* h) {
    if (!h) return;
    p2p_halo_destroy(ctx, &h->ph);
    dev_free(ctx, h->d_send_idx); dev_free(ctx, h->d_send_buf); dev_free(ctx, h->d_ghost);
    delete h;
}
