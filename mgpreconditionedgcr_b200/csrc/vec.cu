// csrc/vec.cu -- Field<num_type> storage and BLAS-1 on device (reference: src/Fields.h), Mesh::blocking (src/Mesh.h).
#include <math.h>

#include "common.cuh"
#include "kernels_blas.cuh"

extern "C" int mgcr_vec_alloc(mgcr_ctx* ctx, int64_t n, mgcr_c128** out) {
    ARG_CHECK(ctx && out && n >= 0, "mgcr_vec_alloc: bad argument");
    return dev_alloc(ctx, sizeof(c128) * (size_t)n, (void**)out);
}

extern "C" int mgcr_vec_free(mgcr_ctx* ctx, mgcr_c128* v) {
    ARG_CHECK(ctx, "ctx is NULL");
    return dev_free(ctx, v);
}

extern "C" int mgcr_vec_upload(mgcr_ctx* ctx, mgcr_c128* d_dst, const mgcr_c128* h_src, int64_t n) {
    ARG_CHECK(ctx && (n == 0 || (d_dst && h_src)), "mgcr_vec_upload: NULL buffer");
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, sizeof(c128) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return MGCR_OK;
}

extern "C" int mgcr_vec_download(mgcr_ctx* ctx, mgcr_c128* h_dst, const mgcr_c128* d_src, int64_t n) {
    ARG_CHECK(ctx && (n == 0 || (h_dst && d_src)), "mgcr_vec_download: NULL buffer");
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, sizeof(c128) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return MGCR_OK;
}

extern "C" int mgcr_vec_copy(mgcr_ctx* ctx, int64_t n, const mgcr_c128* src, mgcr_c128* dst) {
    ARG_CHECK(ctx && (n == 0 || (src && dst)), "mgcr_vec_copy: NULL buffer");
    if (n == 0 || src == dst) return MGCR_OK;
    CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(c128) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    return MGCR_OK;
}

extern "C" int mgcr_vec_set_constant(mgcr_ctx* ctx, int64_t n, double re, double im, mgcr_c128* v) {
    ARG_CHECK(ctx && (n == 0 || v), "mgcr_vec_set_constant: NULL buffer");
    if (n == 0) return MGCR_OK;
    if (re == 0. && im == 0.) {
        CUDA_TRY(cudaMemsetAsync(v, 0, sizeof(c128) * (size_t)n, ctx->stream));
        return MGCR_OK;
    }
    KLAUNCH(ctx, "vec_fill", 16. * n, (launch_pdl(ctx, k_fill, stream_grid(ctx, n, 8), RED_THREADS, 0, n, cmake(re, im), (c128*)v)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

int vec_axpy(mgcr_ctx* ctx, int64_t n, c128 s, const c128* b, const c128* a, c128* out) {
    if (n == 0) return MGCR_OK;
    KLAUNCH(ctx, "vec_axpy", 48. * n, (launch_pdl(ctx, k_axpy, stream_grid(ctx, n, 8), RED_THREADS, 0, n, s, b, a, out)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

extern "C" int mgcr_vec_axpy(mgcr_ctx* ctx, int64_t n, double s_re, double s_im, const mgcr_c128* b, const mgcr_c128* a, mgcr_c128* out) {
    ARG_CHECK(ctx && (n == 0 || (a && b && out)), "mgcr_vec_axpy: NULL buffer");
    return vec_axpy(ctx, n, cmake(s_re, s_im), (const c128*)b, (const c128*)a, (c128*)out);
}

int vec_scale(mgcr_ctx* ctx, int64_t n, c128 s, const c128* a, c128* out) {
    if (n == 0) return MGCR_OK;
    KLAUNCH(ctx, "vec_scale", 32. * n, (launch_pdl(ctx, k_scale, stream_grid(ctx, n, 8), RED_THREADS, 0, n, s, a, out)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

extern "C" int mgcr_vec_scale(mgcr_ctx* ctx, int64_t n, double s_re, double s_im, const mgcr_c128* a, mgcr_c128* out) {
    ARG_CHECK(ctx && (n == 0 || (a && out)), "mgcr_vec_scale: NULL buffer");
    return vec_scale(ctx, n, cmake(s_re, s_im), (const c128*)a, (c128*)out);
}

// device-resident result: d_out[0..1] = sum conj(a) b (this rank's part, then all-reduced)
// n_global: the length of the whole (slab-partitioned) vector when known -- the reduction then has the GPU-count-independent
// shape of red_geom; 0 = unknown (plain shape)
int vec_dot_dev(mgcr_ctx* ctx, int64_t n, const c128* a, const c128* b, double* d_out, bool dist, int64_t n_global = 0) {
    const RedGeom rg = red_geom(ctx, n, n_global > 0 ? n_global : (dist ? 0 : n), 4, 2);
    KLAUNCH(ctx, "vec_dot", 32. * n, (launch_pdl(ctx, k_dot, rg.G, RED_THREADS, 0, rg, a, b, ctx->d_partials, ctx->d_ticket, d_out)));
    CHECK_LAUNCH();
    if (dist) MGCR_TRY(dist_allreduce_sum(ctx, d_out, 2));
    return MGCR_OK;
}

int vec_norm2_dev(mgcr_ctx* ctx, int64_t n, const c128* a, double* d_out, bool dist, int64_t n_global = 0) {
    const RedGeom rg = red_geom(ctx, n, n_global > 0 ? n_global : (dist ? 0 : n), 4, 2);
    KLAUNCH(ctx, "vec_norm2", 16. * n, (launch_pdl(ctx, k_norm2, rg.G, RED_THREADS, 0, rg, a, ctx->d_partials, ctx->d_ticket, d_out)));
    CHECK_LAUNCH();
    if (dist) MGCR_TRY(dist_allreduce_sum(ctx, d_out, 1));
    return MGCR_OK;
}

static int read_scalars(mgcr_ctx* ctx, const double* d, int n, double* h) {
    CUDA_TRY(cudaMemcpyAsync(ctx->h_pinned, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) h[i] = ctx->h_pinned[i];
    return MGCR_OK;
}

// length of the whole vector a slab of n elements belongs to, when every rank holds a slab of the same length (0 otherwise):
// lets the Field-level inner products use the GPU-count-independent reduction shape.  One host all-gather per distinct n.
static int64_t global_len(mgcr_ctx* ctx, int64_t n) {
    if (ctx->nranks == 1) return n;
    auto it = ctx->global_len.find(n);
    if (it != ctx->global_len.end()) return it->second;
    std::vector<int64_t> all;
    int64_t g = 0;
    if (dist_allgather_host_i64(ctx, n, all) == MGCR_OK) {
        g = n * ctx->nranks;
        for (int64_t v : all) if (v != n) g = 0;
    }
    ctx->global_len[n] = g;
    return g;
}

extern "C" int mgcr_vec_dot(mgcr_ctx* ctx, int64_t n, const mgcr_c128* a, const mgcr_c128* b, double out[2]) {
    ARG_CHECK(ctx && out && (n == 0 || (a && b)), "mgcr_vec_dot: NULL buffer");
    MGCR_TRY(vec_dot_dev(ctx, n, (const c128*)a, (const c128*)b, ctx->d_scratch, ctx->nranks > 1, global_len(ctx, n)));
    return read_scalars(ctx, ctx->d_scratch, 2, out);
}

// the same on vectors that are NOT row slabs (replicated coarse-level fields, per-rank scratch): no all-reduce, not collective
extern "C" int mgcr_vec_dot_local(mgcr_ctx* ctx, int64_t n, const mgcr_c128* a, const mgcr_c128* b, double out[2]) {
    ARG_CHECK(ctx && out && (n == 0 || (a && b)), "mgcr_vec_dot_local: NULL buffer");
    MGCR_TRY(vec_dot_dev(ctx, n, (const c128*)a, (const c128*)b, ctx->d_scratch, false, n));
    return read_scalars(ctx, ctx->d_scratch, 2, out);
}
extern "C" int mgcr_vec_squarednorm_local(mgcr_ctx* ctx, int64_t n, const mgcr_c128* a, double* out) {
    ARG_CHECK(ctx && out && (n == 0 || a), "mgcr_vec_squarednorm_local: NULL buffer");
    MGCR_TRY(vec_norm2_dev(ctx, n, (const c128*)a, ctx->d_scratch, false, n));
    return read_scalars(ctx, ctx->d_scratch, 1, out);
}

extern "C" int mgcr_vec_squarednorm(mgcr_ctx* ctx, int64_t n, const mgcr_c128* a, double* out) {
    ARG_CHECK(ctx && out && (n == 0 || a), "mgcr_vec_squarednorm: NULL buffer");
    MGCR_TRY(vec_norm2_dev(ctx, n, (const c128*)a, ctx->d_scratch, ctx->nranks > 1, global_len(ctx, n)));
    return read_scalars(ctx, ctx->d_scratch, 1, out);
}

// a *= 1/sqrt(sum |a|^2), scalar never leaves the device (src/Fields.h:237-243)
int vec_normalise(mgcr_ctx* ctx, int64_t n, c128* a, bool dist, int64_t n_global = 0) {
    MGCR_TRY(vec_norm2_dev(ctx, n, a, ctx->d_scratch + 8, dist, n_global));
    KLAUNCH(ctx, "vec_normalise", 32. * n, (launch_pdl(ctx, k_scale_inv_sqrt, stream_grid(ctx, n, 8), RED_THREADS, 0, n, (const double*)(ctx->d_scratch + 8), a)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

extern "C" int mgcr_vec_normalise(mgcr_ctx* ctx, int64_t n, mgcr_c128* a) {
    ARG_CHECK(ctx && (n == 0 || a), "mgcr_vec_normalise: NULL buffer");
    return vec_normalise(ctx, n, (c128*)a, ctx->nranks > 1, global_len(ctx, n));
}

int vec_gamma5(mgcr_ctx* ctx, int64_t n, int64_t inner, int64_t axis_dim, const c128* in, c128* out) {
    KLAUNCH(ctx, "vec_gamma5", 32. * n, (k_gamma5<<<stream_grid(ctx, n, 8), RED_THREADS, 0, ctx->stream>>>(n, inner, axis_dim, in, out)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

extern "C" int mgcr_vec_gamma5(mgcr_ctx* ctx, int ndim, const int64_t* dims, int axis, const mgcr_c128* in, mgcr_c128* out) {
    ARG_CHECK(ctx && dims && in && out && in != out, "mgcr_vec_gamma5: bad buffer");
    ARG_CHECK(axis >= 0 && axis < ndim, "mgcr_vec_gamma5: axis %d out of range", axis);
    // the reference leaves out[...] zero for axis indices >= 4 (Fields.h:316-336 permutes only 0..3)
    int64_t n = 1, inner = 1;
    for (int i = 0; i < ndim; i++) n *= dims[i];
    for (int i = axis + 1; i < ndim; i++) inner *= dims[i];
    return vec_gamma5(ctx, n, inner, dims[axis], (const c128*)in, (c128*)out);
}

// glibc rand() stream on the host, imaginary part drawn first (src/Fields.h:125-135 as compiled by g++)
int vec_init_rand_slab(mgcr_ctx* ctx, int seed, int64_t skip, int64_t n, c128* d_out);
extern "C" int mgcr_vec_init_rand(mgcr_ctx* ctx, int seed, int64_t n, mgcr_c128* d_out) {
    ARG_CHECK(ctx && (n == 0 || d_out), "mgcr_vec_init_rand: NULL buffer");
    return vec_init_rand_slab(ctx, seed, 0, n, (c128*)d_out);
}
extern "C" int mgcr_vec_init_rand_slab(mgcr_ctx* ctx, int seed, int64_t skip, int64_t n, mgcr_c128* d_out) {
    ARG_CHECK(ctx && skip >= 0 && (n == 0 || d_out), "mgcr_vec_init_rand_slab: bad argument");
    return vec_init_rand_slab(ctx, seed, skip, n, (c128*)d_out);
}

// ----------------------------------------------------------------------------------------------------------
// Mesh::blocking (src/Mesh.h:236-298): one thread per site of the 4 masked dims
// ----------------------------------------------------------------------------------------------------------
struct Dims4 { int64_t d[4]; };

__global__ void k_blocking(int64_t nsite, Dims4 sd, Dims4 sub, Dims4 bd, int64_t bs, int64_t* __restrict__ block_map,
                           int32_t* __restrict__ site_block, int32_t* __restrict__ site_off) {
    for (int64_t site = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; site < nsite; site += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = site, idx[4];
#pragma unroll
        for (int c = 3; c >= 0; c--) { idx[c] = rem % sd.d[c]; rem /= sd.d[c]; }
        int64_t b = 0, o = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            b = b * bd.d[c] + idx[c] / sub.d[c];
            o = o * sub.d[c] + idx[c] % sub.d[c];
        }
        if (block_map) block_map[b * bs + o] = site;
        if (site_block) { site_block[site] = (int32_t)b; site_off[site] = (int32_t)o; }
    }
}

int blocking_device(mgcr_ctx* ctx, const int64_t sd[4], const int64_t sub[4], int64_t bd[4], int64_t* d_block_map,
                    int32_t* d_site_block, int32_t* d_site_off) {
    Dims4 a, b, c;
    int64_t nsite = 1, bs = 1;
    for (int i = 0; i < 4; i++) {
        ARG_CHECK(sub[i] > 0 && sd[i] % sub[i] == 0, "blocking: dimension %lld not divisible by block size %lld (src/Mesh.h:245)",
                  (long long)sd[i], (long long)sub[i]);
        bd[i] = sd[i] / sub[i];
        a.d[i] = sd[i]; b.d[i] = sub[i]; c.d[i] = bd[i];
        nsite *= sd[i]; bs *= sub[i];
    }
    KLAUNCH(ctx, "blocking", 8. * nsite, (k_blocking<<<stream_grid(ctx, nsite, 8), RED_THREADS, 0, ctx->stream>>>(nsite, a, b, c, bs, d_block_map, d_site_block, d_site_off)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

extern "C" int mgcr_blocking_build(mgcr_ctx* ctx, int ndim, const int64_t* dims, const int64_t* sub4, const uint8_t* mask,
                                   int64_t* h_block_map, int64_t* h_block_dim4, int64_t* n_blocks_out) {
    ARG_CHECK(ctx && dims && sub4 && mask && h_block_map && h_block_dim4 && n_blocks_out, "mgcr_blocking_build: NULL argument");
    int64_t sd[4], bd[4];
    int c = 0;
    for (int i = 0; i < ndim; i++)
        if (mask[i]) {
            ARG_CHECK(c < 4, "mgcr_blocking_build: more than 4 masked dimensions");
            sd[c++] = dims[i];
        }
    ARG_CHECK(c == 4, "mgcr_blocking_build: exactly 4 dimensions must be masked (src/Mesh.h:61-62)");
    int64_t nsite = sd[0] * sd[1] * sd[2] * sd[3];
    int64_t* d_map = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)nsite, &d_map));
    int st = blocking_device(ctx, sd, sub4, bd, d_map, nullptr, nullptr);
    if (st == MGCR_OK) {
        cudaError_t e = cudaMemcpyAsync(h_block_map, d_map, sizeof(int64_t) * nsite, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { mgcr_set_error("blocking download: %s", cudaGetErrorString(e)); st = MGCR_ERR_CUDA; }
    }
    dev_free(ctx, d_map);
    MGCR_TRY(st);
    for (int i = 0; i < 4; i++) h_block_dim4[i] = bd[i];
    *n_blocks_out = bd[0] * bd[1] * bd[2] * bd[3];
    return MGCR_OK;
}
