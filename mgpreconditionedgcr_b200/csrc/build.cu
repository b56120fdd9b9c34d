// csrc/build.cu -- device-side construction of the streaming operator images:
//   * host CSR with int64 indices (what Sparse<long> / read_data hold, src/Operator.h:64, src/Parse.cpp:65-91) -> sliced ELL.
//     The arrays are uploaded as they are and packed on the GPU (slice widths, exclusive scan, one warp per slice fills it);
//     round 1 packed on the host: two more host copies of the operator and a single-threaded width pass.
//   * an exclusive scan of int64 counts that stays on the device (also used for the block-CSR streaming image, ops.cu).
#include <algorithm>

#include "ops.cuh"

// ----------------------------------------------------------------------------------------------------------
// exclusive scan of n int64 values in place; *total (device) receives the sum.  Three passes: every CTA scans a tile of
// SCAN_TILE items and publishes its sum, the sums are scanned the same way (recursively), the offsets are added back.
// ----------------------------------------------------------------------------------------------------------
enum { SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS };

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(int64_t n, int64_t* __restrict__ data, int64_t* __restrict__ tile_sums) {
    __shared__ int64_t warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS], run = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? data[base + k] : 0; run += v[k]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int64_t off = 0;
    for (int w = 0; w < warp; w++) off += warp_tot[w];
    int64_t excl = off + incl - run;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) data[base + k] = excl; excl += v[k]; }
    if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = off + incl;
}
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int64_t n, int64_t* __restrict__ data, const int64_t* __restrict__ tile_offsets) {
    const int64_t off = tile_offsets[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_THREADS) if (base + k < n) data[base + k] += off;
}

int device_exclusive_scan_i64(mgcr_ctx* ctx, int64_t* d_data, int64_t n, int64_t* d_total) {
    if (n <= 0) { CUDA_TRY(cudaMemsetAsync(d_total, 0, sizeof(int64_t), ctx->stream)); return MGCR_OK; }
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* sums = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)tiles + 1, &sums));
    k_scan_tile<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(n, d_data, sums);
    CHECK_LAUNCH();
    if (tiles == 1) {
        CUDA_TRY(cudaMemcpyAsync(d_total, sums, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        MGCR_TRY(device_exclusive_scan_i64(ctx, sums, tiles, d_total));
        k_scan_add<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(n, d_data, sums);
        CHECK_LAUNCH();
    }
    return dev_free(ctx, sums);
}

// ----------------------------------------------------------------------------------------------------------
// CSR -> sliced ELL (slice height 32, column-major inside a slice; see SellOp)
// ----------------------------------------------------------------------------------------------------------
// width[s] = 32 * (longest row of slice s); flags |= 1 when row offsets decrease
static __global__ void __launch_bounds__(256) k_sell_width(int64_t nrow, int64_t nslices, const int64_t* __restrict__ row, int64_t* __restrict__ width,
                                                           int* __restrict__ flags) {
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= nslices) return;
    const int lane = threadIdx.x & 31;
    const int64_t r = s * 32 + lane;
    int64_t w = 0;
    if (r < nrow) {
        w = row[r + 1] - row[r];
        if (w < 0) { atomicOr(flags, 1); w = 0; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) width[s] = 32 * w;
}
// one warp per slice: lane l owns row 32 s + l and writes its entries (CSR order kept) at sp[s] + j*32 + l, zero padding
// beyond the row's end; flags |= 2 when a column index lies outside [0, ncol)
static __global__ void __launch_bounds__(256) k_sell_fill(int64_t nrow, int64_t nslices, int64_t ncol, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                                          const c128* __restrict__ val, const int64_t* __restrict__ sp, int32_t* __restrict__ ocol,
                                                          c128* __restrict__ oval, int* __restrict__ flags) {
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= nslices) return;
    const int lane = threadIdx.x & 31;
    const int64_t r = s * 32 + lane;
    const int64_t base = sp[s];
    const int64_t w = (sp[s + 1] - base) >> 5;
    const int64_t b = r < nrow ? row[r] : 0, e = r < nrow ? row[r + 1] : 0;
    for (int64_t j = 0; j < w; j++) {
        int64_t c = 0;
        c128 v = cmake(0., 0.);
        if (b + j < e) {
            c = col[b + j];
            v = val[b + j];
            if (c < 0 || c >= ncol) { atomicOr(flags, 2); c = 0; }
        }
        ocol[base + j * 32 + lane] = (int32_t)c;
        oval[base + j * 32 + lane] = v;
    }
}

// host CSR (int64) -> device sliced-ELL.  `ncol_addressable`: columns < n_local address x, the rest the ghost buffer.
int sell_build(mgcr_ctx* ctx, int64_t nrow, int64_t ncol_addressable, const int64_t* row, const int64_t* col, const mgcr_c128* val, SellOp* op) {
    ARG_CHECK(ncol_addressable < (int64_t)INT32_MAX, "CSR upload: %lld addressable columns exceed the int32 device index (shard the operator)", (long long)ncol_addressable);
    const int64_t nnz = row[nrow], nslices = (nrow + 31) / 32;
    ARG_CHECK(nnz >= 0, "CSR upload: negative entry count");
    int64_t *d_row = nullptr, *d_col = nullptr, *d_total = nullptr;
    c128* d_val = nullptr;
    int* d_flags = nullptr;
    int st = MGCR_OK;
    auto cleanup = [&]() { dev_free(ctx, d_row); dev_free(ctx, d_col); dev_free(ctx, d_val); dev_free(ctx, d_total); dev_free(ctx, d_flags); };
#define BTRY(expr) do { st = (expr); if (st != MGCR_OK) { cleanup(); return st; } } while (0)
#define BCUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { mgcr_set_error("CSR upload: %s -> %s", #expr, cudaGetErrorString(e__)); cleanup(); return e__ == cudaErrorMemoryAllocation ? MGCR_ERR_OOM : MGCR_ERR_CUDA; } } while (0)
    BTRY(dev_alloc_t(ctx, (size_t)nrow + 1, &d_row));
    BTRY(dev_alloc_t(ctx, (size_t)std::max<int64_t>(nnz, 1), &d_col));
    BTRY(dev_alloc_t(ctx, (size_t)std::max<int64_t>(nnz, 1), &d_val));
    BTRY(dev_alloc_t(ctx, 1, &d_total));
    BTRY(dev_alloc_t(ctx, 2, &d_flags));
    BCUDA(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream));
    BCUDA(cudaMemcpyAsync(d_row, row, sizeof(int64_t) * (size_t)(nrow + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        BCUDA(cudaMemcpyAsync(d_col, col, sizeof(int64_t) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
        BCUDA(cudaMemcpyAsync(d_val, val, sizeof(c128) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    BTRY(dev_alloc_t(ctx, (size_t)nslices + 1, &op->d_slice_ptr));
    const unsigned wgrid = (unsigned)std::max<int64_t>(1, (nslices * 32 + 255) / 256);
    if (nslices) {
        k_sell_width<<<wgrid, 256, 0, ctx->stream>>>(nrow, nslices, d_row, op->d_slice_ptr, d_flags);
        BCUDA(cudaGetLastError());
    }
    BTRY(device_exclusive_scan_i64(ctx, op->d_slice_ptr, nslices, d_total));
    BCUDA(cudaMemcpyAsync(op->d_slice_ptr + nslices, d_total, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    int64_t np = 0;
    BCUDA(cudaMemcpyAsync(&np, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    BCUDA(cudaStreamSynchronize(ctx->stream));
    op->nrow = nrow; op->nnz = nnz; op->nnz_padded = np; op->nslices = nslices;
    BTRY(dev_alloc_t(ctx, (size_t)std::max<int64_t>(np, 1), &op->d_col));
    BTRY(dev_alloc_t(ctx, (size_t)std::max<int64_t>(np, 1), &op->d_val));
    if (nslices) {
        k_sell_fill<<<wgrid, 256, 0, ctx->stream>>>(nrow, nslices, ncol_addressable, d_row, d_col, d_val, op->d_slice_ptr, op->d_col, op->d_val, d_flags);
        BCUDA(cudaGetLastError());
    }
    int flags = 0;
    BCUDA(cudaMemcpyAsync(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BCUDA(cudaStreamSynchronize(ctx->stream));
    cleanup();
#undef BTRY
#undef BCUDA
    ARG_CHECK(!(flags & 1), "CSR upload: row offsets decrease");
    ARG_CHECK(!(flags & 2), "CSR upload: column index out of range (src/Operator.h:332 asserts f.field_size() == dim)");
    return MGCR_OK;
}
