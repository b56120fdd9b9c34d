// csrc/p2p.cu -- the two exchanges of the distributed solve written directly over NVLink peer memory instead of NCCL calls:
//   * slab halo exchange: a kernel STORES this rank's boundary plane into the neighbour's receive buffer (peer pointer from
//     a CUDA IPC handle), fences, raises a sequence flag in the neighbour's memory and waits for the neighbours' flags -- one
//     launch per exchange instead of an ncclSend/ncclRecv group (42 us per exchange at 8 GPUs, profiles/r01_bench_mg3d_512_n8.json);
//   * scalar all-reduce of <= 64 doubles: every rank stores its partial sums into a slot of every peer, each 32-bit half
//     travelling in one 8-byte store together with the sequence number (flag-in-data, no fence), then sums the slots in rank
//     order -- identical bits on every rank, one launch instead of an ncclAllReduce (32 us); in the blind solves not even
//     that: the producer kernel's last CTA posts, the consumer kernel's CTAs collect (p2p_allreduce_fold, common.cuh ArPush / ArWait).
// Each context owns a "heap" (one cudaMalloc) that every peer maps; receive buffers and flags are carved from it by an
// allocator (first fit over a coalescing free list, bump pointer above it) that all ranks call in the same order with the same
// sizes, so an offset means the same thing on every rank.
// Buffers are double-buffered by the parity of the sequence number: a neighbour can be at most one exchange ahead (it cannot
// finish exchange k+1 before this rank has contributed to it, which this rank does only after it has consumed exchange k).
// NCCL stays for communicator bootstrap, set-up traffic, the coarse-level all-gather and reductions longer than 64 doubles.
#include "common.cuh"

int dist_allgather(mgcr_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank);

enum { P2P_ALIGN = 256 };   // (P2P_AR_MAX: common.cuh)

struct PeerState {
    unsigned char* heap = nullptr;
    size_t bytes = 0, used = 0;
    std::map<size_t, size_t> free_list;    // offset -> bytes of the areas handed back (coalesced); first fit before the bump pointer
    std::vector<unsigned char*> peer;      // peer[r] = rank r's heap mapped here (peer[rank] = heap)
    size_t ar_off = 0;                     // all-reduce area: [2][nranks][2 * P2P_AR_MAX] uint64
    uint32_t ar_seq = 0;
    unsigned int* d_ticket = nullptr;      // last-CTA ticket of the halo kernel
};

static PeerState* state(mgcr_ctx* ctx) { return (PeerState*)ctx->p2p; }
bool p2p_enabled(mgcr_ctx* ctx) { return ctx->p2p != nullptr; }

// Every rank calls p2p_alloc / p2p_free in the same order with the same sizes (operators and hierarchies are created and
// destroyed collectively), so first fit over the same free list gives the same offset everywhere.
static size_t p2p_round(size_t bytes) { return (bytes + P2P_ALIGN - 1) / P2P_ALIGN * P2P_ALIGN; }
static bool p2p_fits(PeerState* s, size_t bytes) {
    for (auto& kv : s->free_list) if (kv.second >= bytes) return true;
    return s->used + bytes <= s->bytes;
}
int p2p_alloc(mgcr_ctx* ctx, size_t bytes, size_t* off) {
    PeerState* s = state(ctx);
    bytes = p2p_round(bytes);
    if (!s) return MGCR_ERR_OOM;
    for (auto it = s->free_list.begin(); it != s->free_list.end(); ++it) {
        if (it->second < bytes) continue;
        *off = it->first;
        const size_t rest = it->second - bytes;
        s->free_list.erase(it);
        if (rest) s->free_list[*off + bytes] = rest;
        return MGCR_OK;
    }
    if (s->used + bytes > s->bytes) return MGCR_ERR_OOM;
    *off = s->used;
    s->used += bytes;
    return MGCR_OK;
}
void p2p_free(mgcr_ctx* ctx, size_t off, size_t bytes) {
    PeerState* s = state(ctx);
    if (!s || bytes == 0) return;
    bytes = p2p_round(bytes);
    auto it = s->free_list.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != s->free_list.end() && it->first + it->second == nx->first) { it->second += nx->second; s->free_list.erase(nx); }
    if (it != s->free_list.begin()) {
        auto pv = std::prev(it);
        if (pv->first + pv->second == it->first) { pv->second += it->second; s->free_list.erase(it); it = pv; }
    }
    if (it->first + it->second == s->used) { s->used = it->first; s->free_list.erase(it); }   // the top of the heap: lower the bump pointer
}
unsigned char* p2p_ptr(mgcr_ctx* ctx, int rank, size_t off) { return state(ctx)->peer[(size_t)rank] + off; }

// collective; on any failure the context simply keeps using NCCL for everything
int p2p_init(mgcr_ctx* ctx) {
    static const int enabled = getenv("MGCR_P2P") ? atoi(getenv("MGCR_P2P")) : 1;
    if (!enabled || ctx->nranks == 1) return MGCR_OK;
    static const size_t heap_mb = getenv("MGCR_P2P_HEAP_MB") ? (size_t)atoll(getenv("MGCR_P2P_HEAP_MB")) : 192;
    PeerState* s = new PeerState();
    s->bytes = heap_mb << 20;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&s->heap, s->bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; s->heap = nullptr; }
    if (ok && cudaMemset(s->heap, 0, s->bytes) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, s->heap) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    // handles (+ a validity byte) go round through NCCL
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<unsigned char> all(rec * (size_t)ctx->nranks, 0), one(rec, 0);
    memcpy(one.data(), &mine, sizeof(mine));
    one[sizeof(mine)] = (unsigned char)ok;
    unsigned char* d = nullptr;
    MGCR_TRY(dev_alloc(ctx, rec * ((size_t)ctx->nranks + 1), (void**)&d));
    CUDA_TRY(cudaMemcpyAsync(d + rec * (size_t)ctx->nranks, one.data(), rec, cudaMemcpyHostToDevice, ctx->stream));
    MGCR_TRY(dist_allgather(ctx, d + rec * (size_t)ctx->nranks, d, rec));
    CUDA_TRY(cudaMemcpyAsync(all.data(), d, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, d);
    for (int r = 0; r < ctx->nranks; r++) ok = ok && all[rec * (size_t)r + sizeof(mine)];
    s->peer.assign((size_t)ctx->nranks, nullptr);
    if (ok) {
        for (int r = 0; r < ctx->nranks && ok; r++) {
            if (r == ctx->rank) { s->peer[(size_t)r] = s->heap; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + rec * (size_t)r, sizeof(h));
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            s->peer[(size_t)r] = (unsigned char*)p;
        }
    }
    // every rank must have mapped every peer, or nobody uses the peer path
    double* dflag = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, 1, &dflag));
    double f = ok ? 0. : 1.;
    CUDA_TRY(cudaMemcpyAsync(dflag, &f, 8, cudaMemcpyHostToDevice, ctx->stream));
    MGCR_TRY(dist_allreduce_sum(ctx, dflag, 1));
    CUDA_TRY(cudaMemcpyAsync(&f, dflag, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, dflag);
    if (f != 0.) {
        for (int r = 0; r < ctx->nranks; r++) if (r != ctx->rank && s->peer[(size_t)r]) cudaIpcCloseMemHandle(s->peer[(size_t)r]);
        if (s->heap) cudaFree(s->heap);
        delete s;
        if (getenv("MGCR_VERBOSE")) fprintf(stderr, "mgcr: peer-memory path unavailable on rank %d, using NCCL\n", ctx->rank);
        return MGCR_OK;
    }
    CUDA_TRY(cudaMalloc(&s->d_ticket, 64));
    CUDA_TRY(cudaMemset(s->d_ticket, 0, 64));
    ctx->p2p = s;
    MGCR_TRY(p2p_alloc(ctx, sizeof(uint64_t) * 2 * (size_t)ctx->nranks * 2 * P2P_AR_MAX, &s->ar_off));
    return MGCR_OK;
}

void p2p_destroy(mgcr_ctx* ctx) {
    PeerState* s = state(ctx);
    if (!s) return;
    for (int r = 0; r < ctx->nranks; r++) if (r != ctx->rank && s->peer[(size_t)r]) cudaIpcCloseMemHandle(s->peer[(size_t)r]);
    cudaFree(s->heap); cudaFree(s->d_ticket);
    delete s;
    ctx->p2p = nullptr;
}

// ----------------------------------------------------------------------------------------------------------
// all-reduce
// ----------------------------------------------------------------------------------------------------------
struct PeerPtrs { uint64_t* p[16]; };

// slot layout of one parity: [source rank][2 * P2P_AR_MAX] words, word 2i / 2i+1 = {low / high half of element i, seq}
static __global__ void __launch_bounds__(256) k_p2p_allreduce(PeerPtrs peers, uint64_t* mine, int rank, int nranks, int n, uint32_t seq, const double* in,
                                                              double* out) {
    PDL_ENTRY();
    const int words = 2 * n;
    for (int t = threadIdx.x; t < words * nranks; t += blockDim.x) {
        const int p = t / words, j = t - p * words;
        const uint64_t bits = (uint64_t)__double_as_longlong(in[j >> 1]);
        const uint32_t half = (j & 1) ? (uint32_t)(bits >> 32) : (uint32_t)bits;
        st_volatile_u64(peers.p[p] + (size_t)rank * (2 * P2P_AR_MAX) + j, ((uint64_t)seq << 32) | half);
    }
    __syncthreads();   // `in` has been read by every thread before anybody overwrites it (out may alias in)
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ar_collect(mine, nranks, seq, i);   // fixed tree over the ranks (common.cuh)
}

int p2p_allreduce_fold(mgcr_ctx* ctx, int n, const double* d_src, double* d_dst, ArPush* push, ArWait* wait) {
    static const bool fold_on = !(getenv("MGCR_AR_FOLD") && atoi(getenv("MGCR_AR_FOLD")) == 0);   // experiment knob: 0 = stand-alone kernel
    PeerState* s = state(ctx);
    memset(push, 0, sizeof *push);
    memset(wait, 0, sizeof *wait);
    if (!s || !fold_on || n <= 0 || n > AR_FOLD_MAX || ctx->nranks > 16) return MGCR_ERR_UNSUPPORTED;
    s->ar_seq++;
    if (s->ar_seq == 0) s->ar_seq = 1;
    const size_t par = (size_t)(s->ar_seq & 1) * (size_t)ctx->nranks * 2 * P2P_AR_MAX * sizeof(uint64_t);
    for (int r = 0; r < ctx->nranks; r++) push->peer[r] = (uint64_t*)(s->peer[(size_t)r] + s->ar_off + par);
    push->src = d_src; push->rank = ctx->rank; push->nranks = ctx->nranks; push->n = n; push->seq = s->ar_seq;
    wait->mine = (const uint64_t*)(s->heap + s->ar_off + par); wait->dst = d_dst; wait->nranks = ctx->nranks; wait->n = n; wait->seq = s->ar_seq;
    return MGCR_OK;
}

// returns MGCR_ERR_UNSUPPORTED when the caller has to use NCCL (path off, too many values)
int p2p_allreduce_sum(mgcr_ctx* ctx, const double* d_in, double* d_out, int n) {
    PeerState* s = state(ctx);
    if (!s || n > P2P_AR_MAX || ctx->nranks > 16) return MGCR_ERR_UNSUPPORTED;
    if (n <= 0) return MGCR_OK;
    s->ar_seq++;
    if (s->ar_seq == 0) s->ar_seq = 1;
    const size_t par = (size_t)(s->ar_seq & 1) * (size_t)ctx->nranks * 2 * P2P_AR_MAX * sizeof(uint64_t);
    PeerPtrs pp;
    for (int r = 0; r < ctx->nranks; r++) pp.p[r] = (uint64_t*)(s->peer[(size_t)r] + s->ar_off + par);
    KLAUNCH(ctx, "p2p_allreduce", 8. * n, (launch_pdl(ctx, k_p2p_allreduce, 1, 256, 0, pp, (uint64_t*)(s->heap + s->ar_off + par), ctx->rank, ctx->nranks, n, s->ar_seq, d_in, d_out)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// slab halo exchange
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

static __global__ void __launch_bounds__(256) k_p2p_halo(const c128* __restrict__ src_lo, c128* __restrict__ dst_lo, uint32_t* flag_at_lo,
                                                         const c128* __restrict__ src_hi, c128* __restrict__ dst_hi, uint32_t* flag_at_hi, int64_t n,
                                                         uint32_t seq, unsigned int* ticket, const uint32_t* my_flag_lo, const uint32_t* my_flag_hi) {
    PDL_ENTRY();
    __shared__ bool is_last;
    // four elements per thread and direction in flight (posted 16-byte stores over NVLink), grid sized to the face.  Measured: the
    // 22 us per exchange (every GPU count, mean over all levels) did NOT change against 64 CTAs of one store per thread -- the call is
    // launch + flag latency + the wait for the slower neighbour, not bandwidth (profiles/r02_bench_mg3d_512_n{2,8}_final.json)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        c128 a0, a1, a2, a3, b0, b1, b2, b3;
        if (dst_lo) { a0 = ld_stream(src_lo + i); a1 = ld_stream(src_lo + i + stride); a2 = ld_stream(src_lo + i + 2 * stride); a3 = ld_stream(src_lo + i + 3 * stride); }
        if (dst_hi) { b0 = ld_stream(src_hi + i); b1 = ld_stream(src_hi + i + stride); b2 = ld_stream(src_hi + i + 2 * stride); b3 = ld_stream(src_hi + i + 3 * stride); }
        if (dst_lo) { dst_lo[i] = a0; dst_lo[i + stride] = a1; dst_lo[i + 2 * stride] = a2; dst_lo[i + 3 * stride] = a3; }
        if (dst_hi) { dst_hi[i] = b0; dst_hi[i + stride] = b1; dst_hi[i + 2 * stride] = b2; dst_hi[i + 3 * stride] = b3; }
    }
    for (; i < n; i += stride) {
        if (dst_lo) dst_lo[i] = ld_stream(src_lo + i);
        if (dst_hi) dst_hi[i] = ld_stream(src_hi + i);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    __threadfence_system();
    if (flag_at_lo) st_release_sys_u32(flag_at_lo, seq);
    if (flag_at_hi) st_release_sys_u32(flag_at_hi, seq);
    // the kernel ends when the neighbours' planes of the same exchange have landed here
    SpinGuard guard;
    if (my_flag_lo) while ((int32_t)(ld_acquire_sys_u32(my_flag_lo) - seq) < 0) guard.tick();
    if (my_flag_hi) while ((int32_t)(ld_acquire_sys_u32(my_flag_hi) - seq) < 0) guard.tick();
}

// receive area per parity: [n c128 from the lower neighbour][n c128 from the upper neighbour]; flags: lo at +0, hi at +128
int p2p_halo_create(mgcr_ctx* ctx, int64_t n, PeerHalo* h) {
    h->on = false;
    if (!p2p_enabled(ctx)) return MGCR_OK;
    size_t o0, o1, of;
    // all three or nothing: the decision is the same on every rank (same sizes, same allocation order)
    PeerState* s = state(ctx);
    // one area [buffer 0 | buffer 1 | flags]: all of it or nothing, decided identically on every rank.  When the heap is full
    // the object keeps exchanging through NCCL send/recv; MGCR_VERBOSE reports it (raise MGCR_P2P_HEAP_MB).
    const size_t half = p2p_round(sizeof(c128) * 2 * (size_t)n);
    const size_t need = 2 * half + P2P_ALIGN;
    if (!p2p_fits(s, need)) {
        if (getenv("MGCR_VERBOSE")) fprintf(stderr, "mgcr: peer heap full (%zu of %zu bytes used, %zu wanted): this halo uses NCCL send/recv\n", s->used, s->bytes, need);
        return MGCR_OK;
    }
    MGCR_TRY(p2p_alloc(ctx, need, &o0));
    o1 = o0 + half; of = o1 + half;
    // A recycled area starts with clean sequence flags, and nobody may store into it before every rank has cleaned its own
    // and is done with whatever lived there before (its destroy synchronised its stream): one barrier per creation.
    CUDA_TRY(cudaMemsetAsync(s->heap + of, 0, P2P_ALIGN, ctx->stream));
    double* d_bar = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, 1, &d_bar));
    CUDA_TRY(cudaMemsetAsync(d_bar, 0, sizeof(double), ctx->stream));
    MGCR_TRY(dist_allreduce_sum(ctx, d_bar, 1));   // completes on a rank only when every rank has contributed
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, d_bar);
    h->buf_off[0] = o0; h->buf_off[1] = o1; h->flag_off = of; h->n = n; h->seq = 0; h->on = true;
    return MGCR_OK;
}

// hands the receive area back (collective in the same sense as its creation: every rank destroys the object)
void p2p_halo_destroy(mgcr_ctx* ctx, PeerHalo* h) {
    if (!h->on || !p2p_enabled(ctx)) return;
    cudaStreamSynchronize(ctx->stream);
    p2p_free(ctx, h->buf_off[0], 2 * p2p_round(sizeof(c128) * 2 * (size_t)h->n) + P2P_ALIGN);
    h->on = false;
}

// sends n elements starting at send_lo to the lower neighbour and at send_hi to the upper one; *recv_lo / *recv_hi point at
// what the neighbours sent (valid once the kernel enqueued here has completed, i.e. for everything enqueued after it)
// defer: the kernel only puts and raises the flags; the caller's consumer kernel waits (h->wait_lo / wait_hi / seq), so that the
// wait for the slowest neighbour overlaps with everything of the consumer that needs no ghost data
int p2p_halo_exchange(mgcr_ctx* ctx, PeerHalo* h, const c128* send_lo, const c128* send_hi, const c128** recv_lo, const c128** recv_hi, bool defer) {
    PeerState* s = state(ctx);
    const bool has_lo = ctx->rank > 0, has_hi = ctx->rank + 1 < ctx->nranks;
    h->seq++;
    const size_t buf = h->buf_off[h->seq & 1];
    c128* mine = (c128*)(s->heap + buf);
    *recv_lo = has_lo ? mine : nullptr;
    *recv_hi = has_hi ? mine + h->n : nullptr;
    if (!has_lo && !has_hi) return MGCR_OK;
    c128* dst_lo = has_lo ? (c128*)(s->peer[(size_t)ctx->rank - 1] + buf) + h->n : nullptr;       // I am the lower neighbour's upper neighbour
    c128* dst_hi = has_hi ? (c128*)(s->peer[(size_t)ctx->rank + 1] + buf) : nullptr;
    uint32_t* flag_at_lo = has_lo ? (uint32_t*)(s->peer[(size_t)ctx->rank - 1] + h->flag_off + 128) : nullptr;
    uint32_t* flag_at_hi = has_hi ? (uint32_t*)(s->peer[(size_t)ctx->rank + 1] + h->flag_off) : nullptr;
    const uint32_t* my_lo = has_lo ? (const uint32_t*)(s->heap + h->flag_off) : nullptr;
    const uint32_t* my_hi = has_hi ? (const uint32_t*)(s->heap + h->flag_off + 128) : nullptr;
    h->wait_lo = defer ? my_lo : nullptr; h->wait_hi = defer ? my_hi : nullptr;
    if (defer) { my_lo = nullptr; my_hi = nullptr; }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 2, (h->n + 1023) / 1024));
    KLAUNCH(ctx, "p2p_halo", 32. * h->n * ((has_lo ? 1 : 0) + (has_hi ? 1 : 0)),
            (launch_pdl(ctx, k_p2p_halo, grid, 256, 0, send_lo, dst_lo, flag_at_lo, send_hi, dst_hi, flag_at_hi, h->n, h->seq, s->d_ticket, my_lo, my_hi)));
    CHECK_LAUNCH();
    return MGCR_OK;
}
