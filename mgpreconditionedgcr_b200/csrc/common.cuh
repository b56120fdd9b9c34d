// csrc/common.cuh -- context, error plumbing, complex helpers and the deterministic block/grid reduction shared
// by every kernel file of libmgcr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "../../include/mgcr_b200.h"

// ----------------------------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------------------------
void mgcr_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            mgcr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));           \
            return e__ == cudaErrorMemoryAllocation ? MGCR_ERR_OOM : MGCR_ERR_CUDA;                          \
        }                                                                                                    \
    } while (0)

#define MGCR_TRY(expr)                                                                                       \
    do {                                                                                                     \
        int s__ = (expr);                                                                                    \
        if (s__ != MGCR_OK) return s__;                                                                      \
    } while (0)

#define ARG_CHECK(cond, ...)                                                                                 \
    do {                                                                                                     \
        if (!(cond)) {                                                                                       \
            mgcr_set_error(__VA_ARGS__);                                                                     \
            return MGCR_ERR_ARG;                                                                             \
        }                                                                                                    \
    } while (0)

// ----------------------------------------------------------------------------------------------------------
// complex arithmetic, written out so that every element-wise result is what the reference's std::complex
// arithmetic gives (library built with -fmad=false: these paths are HBM-bound, fused multiply-add buys nothing)
// ----------------------------------------------------------------------------------------------------------
typedef double2 c128;   // .x = re, .y = im ; 16-byte aligned -> one 128-bit load/store

__host__ __device__ __forceinline__ c128 cmake(double re, double im) { c128 r; r.x = re; r.y = im; return r; }
__host__ __device__ __forceinline__ c128 cadd(c128 a, c128 b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ c128 csub(c128 a, c128 b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ c128 cmul(c128 a, c128 b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// conj(a) * b
__host__ __device__ __forceinline__ c128 cmulc(c128 a, c128 b) { return cmake(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
// num / den with a real denominator (all GCR coefficients divide by a squared norm)
__host__ __device__ __forceinline__ c128 cdivr(c128 num, double den) { return cmake(num.x / den, num.y / den); }

// streaming 128-bit accesses: vectors are touched once per kernel, keep them out of L1
__device__ __forceinline__ c128 ld_stream(const c128* p) {
    c128 r;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ c128 ld_plain(const c128* p) { return *p; }
__device__ __forceinline__ void st_stream(c128* p, c128 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ----------------------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------------------
struct ProfEntry { double ms = 0; int64_t calls = 0; double bytes = 0; };
struct ProfPending { const char* name; double bytes; int ev; };

struct mgcr_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;      // halo traffic / boundary work overlapped with interior kernels
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_scal = nullptr;
    // reduction scratch: per-block partial sums, a ticket counter, final scalars on device and pinned host mirror
    double* d_partials = nullptr;           // [MAX_RED_BLOCKS][MAX_RED_VALUES]
    unsigned int* d_ticket = nullptr;
    double* d_scratch = nullptr;            // small device scalar scratch for the BLAS-1 entry points
    double* h_pinned = nullptr;             // pinned host mirror (256 doubles)
    double rand_seconds_ms = 0; int64_t rand_calls = 0;   // host wall-clock (ms) spent drawing + uploading init_rand fields (rand.cu)
    void* h_stage = nullptr;                // pinned staging ring for host-generated fields (rand.cu), allocated on first use
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    int64_t launches = 0;
    // host-side time spent inside the allocator and waiting on the device (diagnostics: reported as the profile classes
    // "host_alloc" / "host_sync" when profiling is on)
    double host_alloc_ms = 0, host_sync_ms = 0;
    int64_t host_alloc_calls = 0, host_sync_calls = 0;
    // device-memory cache behind dev_alloc / dev_free (context.cu)
    std::multimap<size_t, void*> mem_free;     // capacity -> idle buffer
    std::map<void*, size_t> mem_live;          // buffer handed out -> capacity
    size_t mem_free_bytes = 0;
    // distributed
    int rank = 0, nranks = 1;
    int64_t slab_align = 1;                 // slab boundaries along the slowest lattice index are multiples of this
    // tunables (mgcr_ctx_set_option; defaults from the environment variables MGCR_SMALL_GCR_N / MGCR_GATHER_DOFS / MGCR_DOT_TMA)
    int64_t small_gcr_rows = (int64_t)1 << 19;   // operators up to this many rows: whole GCR solve in one persistent kernel
    int64_t gather_dofs = (int64_t)1 << 18;      // distributed coarse systems up to this size are replicated on every rank
    int dot_tma = 1;                              // TMA-staged batched inner products for long vectors
    int hopping_kernel = 2;                       // matrix-free stencil: 2 = TMA-staged ring (lattices of >= hopping_tma_rows sites, else form 1),
                                                  // 1 = register-marching / L1 form, 0 = shared-memory tile form
    int64_t hopping_tma_rows = (int64_t)1 << 19;
    int64_t blockcsr_ring_rows = (int64_t)1 << 18;   // block-CSR operators (ne = 2, 4, 8) of at least this many rows stream through the bulk-copy ring
    void* nccl_comm = nullptr;
    void* p2p = nullptr;                    // peer-memory state (p2p.cu): mapped heaps of all ranks, NULL = NCCL for everything
    void* nccl_comm_halo = nullptr;         // second communicator: halo exchanges on the auxiliary stream
    int red_vslabs = 8;                     // option "red_vslabs" / MGCR_RED_VSLABS: 8 = GPU-count-independent reduction shape (red_geom), 1 = plain
    int pdl = 0;                            // option "pdl" / MGCR_PDL: programmatic dependent launch of the solve's kernels (launch_pdl);
                                            // measured on B200: +4.7 us per launch (512^3 solve 1.234 -> 1.274 s, profiles/r02_knobs_pdl_blind.txt), so off
    int halo_overlap = 0;                   // option: overlap halo exchange with interior rows (measured: no gain at 2 and 8 GPUs,
                                            // the exchanges are latency- and skew-bound; profiles/r01_halo_overlap_n8.txt)
    std::vector<cudaEvent_t> depth_events;  // read-back event of each solver nesting depth (gcr.cu)
    std::map<const void*, int> dyn_smem;    // kernels whose dynamic shared-memory limit has been raised on THIS device
    std::map<const void*, int> resident;    // kernel -> CTAs resident on the whole device (resident_ctas)
    std::map<int64_t, int64_t> global_len;  // slab length -> length of the whole vector when all ranks hold equal slabs (vec.cu)
    // profiling
    bool profile = false;
    std::map<std::string, ProfEntry> prof;
    std::vector<cudaEvent_t> prof_events;   // pairs
    std::vector<ProfPending> prof_pending;
};

enum { MAX_RED_BLOCKS = 8192, MAX_RED_VALUES = 32, RED_THREADS = 256 };   // MAX_RED_BLOCKS: partial slots = CTAs x virtual slabs

// Shape of a reduction over a (possibly slab-partitioned) vector that does NOT depend on the number of GPUs: the GLOBAL vector is
// cut into RED_VSLABS equal virtual slabs; a rank that holds `nvs` of them launches G CTAs, and CTA c plays the virtual CTA (v, c) of
// every slab v in turn: it strides over slab v only, with a stride and a CTA count G that depend on the slab length L alone, and
// leaves one partial per (v, c).  The per-slab sums are combined by a
// balanced binary tree over the slab index -- inside a rank over its own slabs, across ranks (which own aligned subtrees when
// their number is a power of two) by the same tree in the all-reduce (p2p.cu).  One GPU therefore forms exactly the partial sums
// and additions that 2, 4 or 8 GPUs form: residual histories are identical bit for bit at every GPU count.
// nvs = 1, G = gridDim.x, L = n is the plain grid-stride shape (vectors that do not divide, small levels).
enum { RED_VSLABS = 8 };
struct RedGeom { int G; int nvs; int64_t L; };

static inline RedGeom red_geom(const mgcr_ctx* ctx, int64_t n_local, int64_t n_global, int per_sm, int items_per_thread);
// grid for a grid-stride streaming kernel over n items, `per_sm` resident blocks of RED_THREADS per SM
static inline int stream_grid(const mgcr_ctx* ctx, int64_t n, int per_sm, int items_per_thread = 1) {
    int64_t need = (n + (int64_t)RED_THREADS * items_per_thread - 1) / ((int64_t)RED_THREADS * items_per_thread);
    int64_t cap = (int64_t)ctx->num_sms * per_sm;
    if (cap > MAX_RED_BLOCKS) cap = MAX_RED_BLOCKS;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// virtual-slab reduction shape for a vector of n_local elements on this rank out of n_global (= n_local when replicated)
static inline RedGeom red_geom(const mgcr_ctx* ctx, int64_t n_local, int64_t n_global, int per_sm, int items_per_thread) {
    RedGeom rg;
    const int nranks = n_global == n_local ? 1 : ctx->nranks;
    const bool ok = ctx->red_vslabs == RED_VSLABS && nranks >= 1 && RED_VSLABS % nranks == 0 && n_global % RED_VSLABS == 0 &&
                    n_local * nranks == n_global && n_global / RED_VSLABS >= ((int64_t)1 << 16);
    if (!ok) { rg.nvs = 1; rg.L = n_local; rg.G = stream_grid(ctx, n_local, per_sm, items_per_thread); return rg; }
    rg.nvs = RED_VSLABS / nranks;
    rg.L = n_global / RED_VSLABS;
    // G depends on the slab length and on per-launch constants only (148 SMs on every B200): the same on every GPU count
    const int64_t need = (rg.L + (int64_t)RED_THREADS * items_per_thread - 1) / ((int64_t)RED_THREADS * items_per_thread);
    const int64_t cap = std::min<int64_t>((int64_t)148 * per_sm, MAX_RED_BLOCKS / RED_VSLABS);   // 148 SMs on every B200
    rg.G = (int)std::max<int64_t>(1, std::min(need, cap));
    return rg;
}

// profile-aware launch bookkeeping: KLAUNCH(ctx, "name", bytes, kernel<<<...>>>(...)).  With profiling on, every
// launch is bracketed by a pair of pooled events on the launching stream; nothing synchronises until the pool is
// drained (mgcr_ctx_get_profile or pool exhaustion), so the timed region keeps its shape.
void prof_begin(mgcr_ctx* ctx, const char* name, double bytes);
void prof_end(mgcr_ctx* ctx);
struct ProfScope {
    mgcr_ctx* ctx;
    ProfScope(mgcr_ctx* c, const char* n, double bytes = 0.) : ctx(c) {
        ctx->launches++;
        if (ctx->profile) prof_begin(ctx, n, bytes);
    }
    ~ProfScope() { if (ctx->profile) prof_end(ctx); }
};
#define KLAUNCH(ctx, name, bytes, launch_expr)                                                               \
    do {                                                                                                     \
        ProfScope ps__(ctx, name, bytes);                                                                    \
        launch_expr;                                                                                         \
    } while (0)

#define CHECK_LAUNCH() CUDA_TRY(cudaGetLastError())

struct HostTimer {
    double* acc; int64_t* calls;
    std::chrono::steady_clock::time_point t0;
    HostTimer(double* a, int64_t* c) : acc(a), calls(c), t0(std::chrono::steady_clock::now()) {}
    ~HostTimer() { *acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); (*calls)++; }
};

// Device memory for everything the library owns.  Inner solves of the multigrid cycle take and return their Krylov
// workspaces on every call (hundreds of times per outer iteration), so buffers are recycled through a per-context
// cache of exact-size classes: after the first cycle no call reaches the driver.  (The first version used
// cudaMallocAsync; its pool cost 0.5 us per call in some runs and 74 us in others, profiles/r01_allocator_cost.txt.)
// All work of a context is enqueued on one stream by one host thread, so reuse in host order is stream-ordered.
int dev_alloc(mgcr_ctx* ctx, size_t bytes, void** out);
int dev_free(mgcr_ctx* ctx, void* p);
template <typename T> static inline int dev_alloc_t(mgcr_ctx* ctx, size_t count, T** out) {
    return dev_alloc(ctx, count * sizeof(T), (void**)out);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remembered per context, not per process
static inline int ensure_dyn_smem(mgcr_ctx* ctx, const void* kernel, int bytes) {
    auto it = ctx->dyn_smem.find(kernel);
    if (it != ctx->dyn_smem.end() && it->second >= bytes) return MGCR_OK;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ctx->dyn_smem[kernel] = bytes;
    return MGCR_OK;
}

// CTAs of `threads` threads that are resident on the whole device at once for `kernel` (registers, shared memory): the grid
// of a persistent kernel.  A larger grid runs in waves of whole-loop CTAs, and the last, partial wave leaves SMs idle (round 2's
// first persistent restrict: 1184 CTAs of a kernel that fits 5 per SM -> achieved occupancy 51 %, 5.6 instead of 7.1 TB/s).
static inline int resident_ctas(mgcr_ctx* ctx, const void* kernel, int threads, size_t smem = 0) {
    auto it = ctx->resident.find(kernel);
    if (it != ctx->resident.end()) return it->second;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
    const int total = per_sm * ctx->num_sms;
    ctx->resident[kernel] = total;
    return total;
}

// peer-memory exchanges (p2p.cu)
struct PeerHalo {
    bool on = false; size_t buf_off[2] = {0, 0}; size_t flag_off = 0; int64_t n = 0; uint32_t seq = 0;
    // deferred wait (p2p_halo_exchange(..., defer = true)): the put kernel does not wait for the neighbours' planes; the kernel
    // that consumes them polls these flags (NULL = no neighbour on that side) for `seq` right before its first ghost access
    const uint32_t* wait_lo = nullptr; const uint32_t* wait_hi = nullptr;
};
// consumer side of a deferred halo wait: spin until the neighbour's flag has reached seq (ld.acquire.sys)
__device__ __forceinline__ void p2p_flag_wait(const uint32_t* flag, uint32_t seq) {
    uint32_t v;
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - seq) >= 0) break;
        if ((++spins & 0x3fffu) == 0) {   // a peer that never arrives must not hang this GPU for ever: trap after 60 s
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now; else if (now - t0 > 60000000000ull) __trap();
        }
    } while (true);
}
// ---- all-reduce folded into its producer and consumer kernels (p2p.cu owns the slots; gcr.cu uses it for the solves nobody
// watches).  The slot area of a rank holds, per parity of the sequence number, [source rank][2 * P2P_AR_MAX] 64-bit words:
// word 2i / 2i+1 = {low / high half of value i, sequence number} -- data and flag travel in one store.
//   producer: the last CTA of the reducing kernel posts this rank's partial sums into every rank's slots (ar_push);
//   consumer: every CTA of the next kernel polls its own rank's slots, adds the ranks up in the fixed tree order and keeps the
//             sums in shared memory; CTA 0 also stores them where later kernels read them (ar_wait).
// Against a stand-alone all-reduce kernel this removes one launch and two launch gaps per reduction (21 per outer iteration
// of the 512^3 solve on 8 GPUs).
enum { P2P_AR_MAX = 64, AR_FOLD_MAX = 8 };
struct ArPush { uint64_t* peer[16]; const double* src; int rank, nranks, n; uint32_t seq; };   // seq 0: not folded
struct ArWait { const uint64_t* mine; double* dst; int nranks, n; uint32_t seq; };
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// A rank that died (or never reached the matching exchange) must not leave its peers spinning for ever: after 60 s of
// waiting the kernel traps, the stream reports an error and the caller fails loudly.
struct SpinGuard {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ void tick() {
        if ((++spins & 0x3fffu) != 0) return;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 60000000000ull) __trap();
    }
};
// all threads of ONE CTA, after a block barrier that follows the last write to src[0..n)
__device__ __forceinline__ void ar_push(const ArPush& a) {
    const int words = 2 * a.n;
    for (int t = threadIdx.x; t < words * a.nranks; t += blockDim.x) {
        const int p = t / words, j = t - p * words;
        const uint64_t bits = ld_volatile_u64((const uint64_t*)(a.src + (j >> 1)));
        const uint32_t half = (j & 1) ? (uint32_t)(bits >> 32) : (uint32_t)bits;
        st_volatile_u64(a.peer[p] + (size_t)a.rank * (2 * P2P_AR_MAX) + j, ((uint64_t)a.seq << 32) | half);
    }
}
// one value: wait for every rank's contribution and add them as a balanced binary tree over the rank index -- with a power-of-two
// number of ranks this continues the tree each rank summed its own virtual slabs with (RedGeom), so 1, 2, 4 and 8 GPUs perform the
// same additions; identical bits on every rank in any case
static __device__ __noinline__ double ar_collect(const uint64_t* mine, int nranks, uint32_t seq, int i) {
    double part[16];
    for (int r = 0; r < nranks; r++) {
        const uint64_t* w = mine + (size_t)r * (2 * P2P_AR_MAX) + 2 * i;
        uint64_t lo, hi;
        SpinGuard guard;
        while ((uint32_t)((lo = ld_volatile_u64(w)) >> 32) != seq) guard.tick();
        while ((uint32_t)((hi = ld_volatile_u64(w + 1)) >> 32) != seq) guard.tick();
        part[r] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
    }
    for (int w = 1; w < nranks; w <<= 1)
        for (int r = 0; r + w < nranks; r += 2 * w) part[r] += part[r + w];
    return part[0];
}
// every thread of every CTA of the consumer; vals = shared memory, AR_FOLD_MAX doubles
__device__ __forceinline__ void ar_wait(const ArWait& a, double* vals) {
    if ((int)threadIdx.x < a.n) {
        const double s = ar_collect(a.mine, a.nranks, a.seq, threadIdx.x);
        vals[threadIdx.x] = s;
        if (blockIdx.x == 0) a.dst[threadIdx.x] = s;
    }
    __syncthreads();
}
int p2p_init(mgcr_ctx* ctx);
void p2p_destroy(mgcr_ctx* ctx);
// reserves the next all-reduce sequence number for a reduction of n values that travels inside its producer and consumer kernels;
// MGCR_ERR_UNSUPPORTED (and both descriptors off) when the caller has to run a stand-alone all-reduce instead
int p2p_allreduce_fold(mgcr_ctx* ctx, int n, const double* d_src, double* d_dst, ArPush* push, ArWait* wait);
bool p2p_enabled(mgcr_ctx* ctx);
int p2p_allreduce_sum(mgcr_ctx* ctx, const double* d_in, double* d_out, int n);
int p2p_halo_create(mgcr_ctx* ctx, int64_t n, PeerHalo* h);
void p2p_halo_destroy(mgcr_ctx* ctx, PeerHalo* h);
int p2p_halo_exchange(mgcr_ctx* ctx, PeerHalo* h, const c128* send_lo, const c128* send_hi, const c128** recv_lo, const c128** recv_hi, bool defer = false);

// distributed helpers (dist.cu)
void dist_destroy(mgcr_ctx* ctx);
int dist_allreduce_sum(mgcr_ctx* ctx, double* d_buf, int n);
// out[i] = sum over ranks of in[i]; `in` is left untouched (out may alias it), so repeating the call gives the same result
int dist_allreduce_sum2(mgcr_ctx* ctx, const double* d_in, double* d_out, int n);
int dist_sendrecv(mgcr_ctx* ctx, const void* d_send, size_t send_bytes, int send_peer, void* d_recv, size_t recv_bytes,
                  int recv_peer, cudaStream_t stream);
bool dist_halo_overlap(mgcr_ctx* ctx);
int dist_halo_begin(mgcr_ctx* ctx, cudaStream_t* stream_out);
int dist_halo_end(mgcr_ctx* ctx);
int dist_halo_wait(mgcr_ctx* ctx);
int dist_group_begin(mgcr_ctx* ctx);
int dist_group_end(mgcr_ctx* ctx);
int dist_send(mgcr_ctx* ctx, const void* d_send, size_t bytes, int peer, cudaStream_t stream);
int dist_recv(mgcr_ctx* ctx, void* d_recv, size_t bytes, int peer, cudaStream_t stream);
int dist_allgather_host_i64(mgcr_ctx* ctx, int64_t mine, std::vector<int64_t>& all);
int dist_allgather_host_bytes(mgcr_ctx* ctx, const void* mine, size_t bytes, std::vector<unsigned char>& all);
int dist_allgather(mgcr_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank);

// ----------------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A solve is thousands of short kernels in one stream (12 000 per 512^3 solve at 8 GPUs,
// 40-100 us each): between two ordinary launches the GPU drains completely, then pays the launch latency and the CTA
// ramp of the next kernel.  Kernels launched through launch_pdl() carry the programmatic-stream-serialisation attribute:
// their CTAs are scheduled while the previous kernel's last CTAs are still running and block in PDL_ENTRY() (griddepcontrol.wait)
// until that kernel has completed and its writes are visible; PDL_ENTRY() then lets the kernel after this one be scheduled
// in the same way.  Every kernel that is launched this way starts with PDL_ENTRY() before it touches memory; without the
// attribute the two instructions are no-ops, and any other stream operation in between is an ordinary dependency.
// ----------------------------------------------------------------------------------------------------------
#define PDL_ENTRY()                                                       \
    do {                                                                  \
        asm volatile("griddepcontrol.wait;" ::: "memory");                \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   \
    } while (0)

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline void launch_pdl(mgcr_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (ctx->pdl && !ctx->profile) ? 1 : 0;   // the profile's event records sit between the kernels anyway
    (void)cudaLaunchKernelEx(&cfg, kernel, args...);      // a failure is picked up by cudaGetLastError() like a <<<>>> launch
}
#endif

// ----------------------------------------------------------------------------------------------------------
// deterministic reduction: warp shuffle -> shared memory -> one partial per block -> the LAST block to finish
// (ticket counter) sums the partials with a fixed thread assignment and a fixed tree, so the result does not depend
// on block scheduling.  NV values are reduced at once; result[k] is written by the last block only.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// The last CTA's part: value q summed over the CTA partials of every virtual slab (lanes stride over the slab's G CTAs in the order
// b = lane, lane+32, ..., then the shuffle tree), then the slab sums by a balanced binary tree over the slab index.
// All RED_THREADS threads call it.  The (slab, value) pairs are dealt over the eight warps and each lane keeps four loads in
// flight (the additions stay in order): with one warp per VALUE and one load at a time this tail was 148 dependent L2 round trips
// for G = 592 and 8 slabs -- a quarter of the 89 us of a level-1 inner-product kernel at 512^3.
// `scratch`: RED_VSLABS * nv doubles of shared memory that nobody reads any more.
__device__ __forceinline__ void combine_partials(const double* __restrict__ partials, int stride_vals, int nv, const RedGeom& rg, double* scratch,
                                                 double* __restrict__ result, int nwrite) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int task = warp; task < rg.nvs * nv; task += RED_THREADS / 32) {
        const int v = task / nv, q = task - v * nv;
        const double* src = partials + (size_t)v * rg.G * stride_vals + q;
        double acc = 0.;
        int b = lane;
        for (; b + 96 < rg.G; b += 128) {
            const double a0 = __ldcg(src + (size_t)b * stride_vals), a1 = __ldcg(src + (size_t)(b + 32) * stride_vals);
            const double a2 = __ldcg(src + (size_t)(b + 64) * stride_vals), a3 = __ldcg(src + (size_t)(b + 96) * stride_vals);
            acc += a0; acc += a1; acc += a2; acc += a3;
        }
        for (; b < rg.G; b += 32) acc += __ldcg(src + (size_t)b * stride_vals);
        acc = warp_sum(acc);
        if (lane == 0) scratch[v * nv + q] = acc;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nv; q += RED_THREADS) {
        double segs[RED_VSLABS];
#pragma unroll
        for (int v = 0; v < RED_VSLABS; v++) segs[v] = v < rg.nvs ? scratch[v * nv + q] : 0.;
#pragma unroll
        for (int w = 1; w < RED_VSLABS; w <<= 1)
#pragma unroll
            for (int i = 0; i + w < RED_VSLABS; i += 2 * w)
                if (i + w < rg.nvs) segs[i] += segs[i + w];
        if (q < nwrite) result[q] = segs[0];
    }
}

// blockDim.x must be RED_THREADS, gridDim.x = rg.G.  Two steps:
//   slab_partial<NV>(v, vslab)                        after the pass over virtual slab `vslab`: every WARP leaves the sum of its lanes
//                                                     in shared memory -- no block barrier, the warps of a CTA drift from slab to
//                                                     slab independently and the loads keep flowing (a barrier per slab cost 14 % of
//                                                     update_xr at 256^3: 14 elements per thread and slab);
//   grid_finish<NV>(partials, ticket, result, rg)     once: the CTA's partial of every slab = its warp sums added in warp order, then
//                                                     the last CTA to arrive combines all partials (combine_partials).
template <int NV> struct SlabSums { double w[RED_VSLABS][RED_THREADS / 32][NV]; };

template <int NV>
__device__ __forceinline__ void slab_partial(const double (&v)[NV], SlabSums<NV>& sm, int vslab) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = warp_sum(v[k]);
        if (lane == 0) sm.w[vslab][warp][k] = s;
    }
}
template <int NV>
__device__ __forceinline__ bool grid_finish(SlabSums<NV>& sm, double* __restrict__ partials, unsigned int* ticket, double* __restrict__ result,
                                            const RedGeom& rg, int nwrite = NV, const ArPush* push = nullptr) {
    constexpr int NW = RED_THREADS / 32;
    __shared__ bool is_last;
    __syncthreads();
    for (int t = threadIdx.x; t < rg.nvs * NV; t += RED_THREADS) {
        const int vs = t / NV, k = t - vs * NV;
        double s = 0.;
#pragma unroll
        for (int w = 0; w < NW; w++) s += sm.w[vs][w][k];
        partials[(size_t)(vs * rg.G + blockIdx.x) * NV + k] = s;
    }
    __threadfence();                              // this CTA's partials are visible before its ticket is
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicInc(ticket, gridDim.x - 1);   // wraps back to 0 after the last block: self-resetting
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // last block: every (slab, value) pair in a fixed order (combine_partials); the warp sums in shared memory have all been read
    combine_partials(partials, NV, NV, rg, &sm.w[0][0][0], result, nwrite);
    if (push && push->seq) {   // folded all-reduce: this rank's sums leave for every rank's slots right here
        __syncthreads();
        ar_push(*push);
    }
    return true;
}
