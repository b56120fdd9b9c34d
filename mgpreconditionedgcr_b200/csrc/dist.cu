// csrc/dist.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink for exactly two things, the per-level
// halo exchange (send/recv pairs in one group) and the scalar inner-product all-reduce.  Nothing in the reference
// corresponds to this file (it is single-process); the partition follows SURVEY.md 8(e): contiguous row slabs along
// the slowest mesh index.  NCCL is bound at run time (dlopen) so that single-GPU use has no NCCL dependency.
#include <dlfcn.h>

#include "common.cuh"

// minimal NCCL ABI (stable across 2.x): the types below match nccl.h
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclChar_ = 0, ncclInt64_ = 4, ncclFloat64_ = 8 };
enum { ncclSum_ = 0 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;   // optional (NCCL >= 2.18)
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return MGCR_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);   // resolves to an already loaded copy (e.g. torch's) when there is one
        if (lib) break;
    }
    if (!lib) { mgcr_set_error("NCCL not found: %s", dlerror()); return MGCR_ERR_NCCL; }
#define BIND(field, sym)                                                                          \
    *(void**)(&g_nccl.field) = dlsym(lib, sym);                                                   \
    if (!g_nccl.field) { mgcr_set_error("NCCL symbol %s missing", sym); return MGCR_ERR_NCCL; }
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllReduce, "ncclAllReduce");
    BIND(AllGather, "ncclAllGather");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    *(void**)(&g_nccl.CommSplit) = dlsym(lib, "ncclCommSplit");
    g_nccl.lib = lib;
    return MGCR_OK;
}

#define NCCL_TRY(expr)                                                                            \
    do {                                                                                          \
        ncclResult_t r__ = (expr);                                                                \
        if (r__ != ncclSuccess_) {                                                                \
            mgcr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
            return MGCR_ERR_NCCL;                                                                 \
        }                                                                                         \
    } while (0)

extern "C" int mgcr_nccl_unique_id(void* h_id128) {
    ARG_CHECK(h_id128, "mgcr_nccl_unique_id: NULL buffer");
    MGCR_TRY(nccl_load());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(h_id128, &id, 128);
    return MGCR_OK;
}

extern "C" int mgcr_ctx_init_dist(mgcr_ctx* ctx, int rank, int nranks, const void* h_id128) {
    ARG_CHECK(ctx && nranks >= 1 && rank >= 0 && rank < nranks, "mgcr_ctx_init_dist: bad rank %d / %d", rank, nranks);
    ARG_CHECK(ctx->nccl_comm == nullptr, "mgcr_ctx_init_dist: already initialised");
    if (nranks == 1) { ctx->rank = 0; ctx->nranks = 1; return MGCR_OK; }
    ARG_CHECK(h_id128, "mgcr_ctx_init_dist: NULL id");
    MGCR_TRY(nccl_load());
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, h_id128, 128);
    ncclComm_t comm;
    NCCL_TRY(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm; ctx->rank = rank; ctx->nranks = nranks;
    // a second communicator for the halo exchanges: they run on the auxiliary stream while interior rows are computed, and
    // NCCL orders all operations of ONE communicator, whatever stream they are on
    if (g_nccl.CommSplit) {
        ncclComm_t halo = nullptr;
        if (g_nccl.CommSplit(comm, 0, rank, &halo, nullptr) == ncclSuccess_) ctx->nccl_comm_halo = halo;
    }
    // halo exchange and short all-reduces over NVLink peer memory (p2p.cu); NCCL remains the fallback and the bootstrap
    return p2p_init(ctx);
}

extern "C" int mgcr_ctx_set_slab_align(mgcr_ctx* ctx, int64_t align) {
    ARG_CHECK(ctx && align >= 1, "mgcr_ctx_set_slab_align: bad argument");
    ctx->slab_align = align;
    return MGCR_OK;
}

void dist_destroy(mgcr_ctx* ctx) {
    p2p_destroy(ctx);
    if (ctx->nccl_comm_halo && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm_halo);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr; ctx->nccl_comm_halo = nullptr;
}

extern "C" int mgcr_ctx_rank(mgcr_ctx* ctx, int* rank, int* nranks) {
    ARG_CHECK(ctx, "ctx is NULL");
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return MGCR_OK;
}

int dist_allreduce_sum2(mgcr_ctx* ctx, const double* d_in, double* d_out, int n) {
    if (ctx->nranks == 1) {
        if (d_in != d_out) CUDA_TRY(cudaMemcpyAsync(d_out, d_in, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        return MGCR_OK;
    }
    {
        const int st = p2p_allreduce_sum(ctx, d_in, d_out, n);
        if (st != MGCR_ERR_UNSUPPORTED) return st;
    }
    ProfScope ps_(ctx, "nccl_allreduce", 8. * n);
    NCCL_TRY(g_nccl.AllReduce(d_in, d_out, (size_t)n, ncclFloat64_, ncclSum_, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return MGCR_OK;
}
int dist_allreduce_sum(mgcr_ctx* ctx, double* d_buf, int n) { return dist_allreduce_sum2(ctx, d_buf, d_buf, n); }

extern "C" int mgcr_allreduce_sum(mgcr_ctx* ctx, double* d_buf, int n) {
    ARG_CHECK(ctx && d_buf && n >= 0, "mgcr_allreduce_sum: bad argument");
    return dist_allreduce_sum(ctx, d_buf, n);
}

// halo exchanges overlap with interior compute when the second communicator exists and the option is on
bool dist_halo_overlap(mgcr_ctx* ctx) { return ctx->nranks > 1 && ctx->nccl_comm_halo != nullptr && ctx->halo_overlap != 0; }

// Bracket of one halo exchange.  Overlapped: the sends / receives go to the auxiliary stream (after everything enqueued so
// far on the main stream, which produced the data being sent); the caller launches its interior work on the main stream
// and then calls dist_halo_wait() before the work that reads the ghosts.
int dist_halo_begin(mgcr_ctx* ctx, cudaStream_t* stream_out) {
    *stream_out = ctx->stream;
    if (ctx->nranks == 1) return MGCR_OK;
    ctx->launches++;
    if (dist_halo_overlap(ctx)) {
        CUDA_TRY(cudaEventRecord(ctx->ev_a, ctx->stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_a, 0));
        *stream_out = ctx->aux_stream;
    } else if (ctx->profile) {
        prof_begin(ctx, "nccl_halo", 0.);
    }
    NCCL_TRY(g_nccl.GroupStart());
    return MGCR_OK;
}
int dist_halo_end(mgcr_ctx* ctx) {
    if (ctx->nranks == 1) return MGCR_OK;
    NCCL_TRY(g_nccl.GroupEnd());
    if (dist_halo_overlap(ctx)) CUDA_TRY(cudaEventRecord(ctx->ev_b, ctx->aux_stream));
    else if (ctx->profile) prof_end(ctx);
    return MGCR_OK;
}
int dist_halo_wait(mgcr_ctx* ctx) {
    if (dist_halo_overlap(ctx)) CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    return MGCR_OK;
}
int dist_group_begin(mgcr_ctx* ctx) {
    if (ctx->nranks == 1) return MGCR_OK;
    ctx->launches++;
    NCCL_TRY(g_nccl.GroupStart());
    return MGCR_OK;
}
int dist_group_end(mgcr_ctx* ctx) {
    if (ctx->nranks == 1) return MGCR_OK;
    NCCL_TRY(g_nccl.GroupEnd());
    return MGCR_OK;
}
static ncclComm_t comm_for(mgcr_ctx* ctx, cudaStream_t stream) {
    return (ncclComm_t)((stream == ctx->aux_stream && ctx->nccl_comm_halo) ? ctx->nccl_comm_halo : ctx->nccl_comm);
}
int dist_send(mgcr_ctx* ctx, const void* d_send, size_t bytes, int peer, cudaStream_t stream) {
    NCCL_TRY(g_nccl.Send(d_send, bytes, ncclChar_, peer, comm_for(ctx, stream), stream));
    return MGCR_OK;
}
int dist_recv(mgcr_ctx* ctx, void* d_recv, size_t bytes, int peer, cudaStream_t stream) {
    NCCL_TRY(g_nccl.Recv(d_recv, bytes, ncclChar_, peer, comm_for(ctx, stream), stream));
    return MGCR_OK;
}

// equal-size all-gather of raw bytes (coarse-level gather: SURVEY.md 8e item 3)
int dist_allgather(mgcr_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank) {
    if (ctx->nranks == 1) {
        if (d_send != d_recv) CUDA_TRY(cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return MGCR_OK;
    }
    ProfScope ps_(ctx, "nccl_allgather", (double)bytes_per_rank * ctx->nranks);
    NCCL_TRY(g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclChar_, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return MGCR_OK;
}

// every rank contributes one int64, all get the list (setup-time metadata: slab offsets, counts)
int dist_allgather_host_i64(mgcr_ctx* ctx, int64_t mine, std::vector<int64_t>& all) {
    all.assign((size_t)ctx->nranks, mine);
    if (ctx->nranks == 1) return MGCR_OK;
    int64_t* d = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)ctx->nranks + 1, &d));
    CUDA_TRY(cudaMemcpyAsync(d + ctx->nranks, &mine, 8, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(g_nccl.AllGather(d + ctx->nranks, d, 1, ncclInt64_, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(all.data(), d, 8 * (size_t)ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return dev_free(ctx, d);
}

// every rank contributes `bytes` bytes, all get the nranks * bytes concatenation (setup-time metadata)
int dist_allgather_host_bytes(mgcr_ctx* ctx, const void* mine, size_t bytes, std::vector<unsigned char>& all) {
    all.assign(bytes * (size_t)ctx->nranks, 0);
    if (bytes == 0) return MGCR_OK;
    if (ctx->nranks == 1) { memcpy(all.data(), mine, bytes); return MGCR_OK; }
    unsigned char* d = nullptr;
    MGCR_TRY(dev_alloc(ctx, bytes * ((size_t)ctx->nranks + 1), (void**)&d));
    CUDA_TRY(cudaMemcpyAsync(d + bytes * (size_t)ctx->nranks, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(g_nccl.AllGather(d + bytes * (size_t)ctx->nranks, d, bytes, ncclChar_, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(all.data(), d, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return dev_free(ctx, d);
}

// contiguous slabs of [0, n) in units of `align`; the first (n/align) % nranks ranks get one unit more
extern "C" int mgcr_slab_range(int64_t n, int64_t align, int rank, int nranks, int64_t* begin, int64_t* end) {
    ARG_CHECK(begin && end && nranks >= 1 && rank >= 0 && rank < nranks && align >= 1, "mgcr_slab_range: bad argument");
    ARG_CHECK(n % align == 0, "mgcr_slab_range: extent %lld is not a multiple of the aggregate size %lld", (long long)n, (long long)align);
    int64_t units = n / align, q = units / nranks, rem = units % nranks;
    int64_t b = (int64_t)rank * q + (rank < rem ? rank : rem);
    int64_t e = b + q + (rank < rem ? 1 : 0);
    *begin = b * align; *end = e * align;
    return MGCR_OK;
}
