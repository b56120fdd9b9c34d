// csrc/context.cu -- context lifetime, error string, stream-ordered memory, profile export.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void mgcr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char* mgcr_last_error(void) { return g_err; }
extern "C" int mgcr_abi_version(void) { return 1; }

// Idle buffers are kept until the device runs out of memory (dev_alloc then trims and retries): a 512^3 solve recycles
// ~100 GB of Krylov workspaces between the levels of one cycle, a fixed cap only makes the cache thrash.

static void mem_trim(mgcr_ctx* ctx) {
    if (ctx->mem_free.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->mem_free) cudaFree(kv.second);
    ctx->mem_free.clear();
    ctx->mem_free_bytes = 0;
}

int dev_alloc(mgcr_ctx* ctx, size_t bytes, void** out) {
    *out = nullptr;
    HostTimer ht(&ctx->host_alloc_ms, &ctx->host_alloc_calls);
    bytes = (bytes + 511) & ~(size_t)511;
    if (bytes == 0) bytes = 512;
    auto it = ctx->mem_free.lower_bound(bytes);
    if (it != ctx->mem_free.end() && it->first <= bytes + bytes / 8) {   // close enough in size: recycle
        *out = it->second;
        ctx->mem_live[it->second] = it->first;
        ctx->mem_free_bytes -= it->first;
        ctx->mem_free.erase(it);
        return MGCR_OK;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation && !ctx->mem_free.empty()) {      // give idle buffers back and retry
        cudaGetLastError();
        mem_trim(ctx);
        e = cudaMalloc(out, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        mgcr_set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? MGCR_ERR_OOM : MGCR_ERR_CUDA;
    }
    ctx->mem_live[*out] = bytes;
    return MGCR_OK;
}

int dev_free(mgcr_ctx* ctx, void* p) {
    if (!p) return MGCR_OK;
    HostTimer ht(&ctx->host_alloc_ms, &ctx->host_alloc_calls);
    auto it = ctx->mem_live.find(p);
    if (it == ctx->mem_live.end()) {
        mgcr_set_error("dev_free: pointer %p was not allocated by this context", p);
        return MGCR_ERR_ARG;
    }
    ctx->mem_free.emplace(it->second, p);
    ctx->mem_free_bytes += it->second;
    ctx->mem_live.erase(it);
    return MGCR_OK;
}

extern "C" int mgcr_ctx_create(int device, mgcr_ctx** out) {
    ARG_CHECK(out != nullptr, "mgcr_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        mgcr_set_error("no CUDA device available (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return MGCR_ERR_CUDA;
    }
    ARG_CHECK(device >= 0 && device < ndev, "mgcr_ctx_create: device %d out of range (have %d)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        mgcr_set_error("device %d is sm_%d%d; libmgcr_b200 carries sm_100a code only", device, prop.major, prop.minor);
        return MGCR_ERR_CUDA;
    }
    mgcr_ctx* c = new mgcr_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_scal, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&c->d_partials, sizeof(double) * MAX_RED_BLOCKS * MAX_RED_VALUES));
    CUDA_TRY(cudaMalloc(&c->d_ticket, sizeof(unsigned int) * 4));
    CUDA_TRY(cudaMemset(c->d_ticket, 0, sizeof(unsigned int) * 4));
    CUDA_TRY(cudaMalloc(&c->d_scratch, sizeof(double) * 256));
    CUDA_TRY(cudaMemset(c->d_scratch, 0, sizeof(double) * 256));
    CUDA_TRY(cudaMallocHost(&c->h_pinned, sizeof(double) * 256));
    if (getenv("MGCR_SMALL_GCR_N")) c->small_gcr_rows = atoll(getenv("MGCR_SMALL_GCR_N"));
    if (getenv("MGCR_GATHER_DOFS")) c->gather_dofs = atoll(getenv("MGCR_GATHER_DOFS"));
    if (getenv("MGCR_DOT_TMA")) c->dot_tma = atoi(getenv("MGCR_DOT_TMA"));
    if (getenv("MGCR_HOPPING_KERNEL")) c->hopping_kernel = atoi(getenv("MGCR_HOPPING_KERNEL"));
    if (getenv("MGCR_HOPPING_TMA_ROWS")) c->hopping_tma_rows = atoll(getenv("MGCR_HOPPING_TMA_ROWS"));
    if (getenv("MGCR_HALO_OVERLAP")) c->halo_overlap = atoi(getenv("MGCR_HALO_OVERLAP"));
    if (getenv("MGCR_PDL")) c->pdl = atoi(getenv("MGCR_PDL"));
    if (getenv("MGCR_RED_VSLABS")) c->red_vslabs = atoi(getenv("MGCR_RED_VSLABS"));
    *out = c;
    return MGCR_OK;
}

extern "C" int mgcr_ctx_set_option(mgcr_ctx* c, const char* key, int64_t value) {
    ARG_CHECK(c && key, "mgcr_ctx_set_option: NULL argument");
    if (!strcmp(key, "small_gcr_rows")) c->small_gcr_rows = value;
    else if (!strcmp(key, "gather_dofs")) c->gather_dofs = value;
    else if (!strcmp(key, "dot_tma")) c->dot_tma = (int)value;
    else if (!strcmp(key, "hopping_kernel")) c->hopping_kernel = (int)value;
    else if (!strcmp(key, "hopping_tma_rows")) c->hopping_tma_rows = value;
    else if (!strcmp(key, "blockcsr_ring_rows")) c->blockcsr_ring_rows = value;
    else if (!strcmp(key, "halo_overlap")) c->halo_overlap = (int)value;
    else if (!strcmp(key, "pdl")) c->pdl = (int)value;
    else if (!strcmp(key, "red_vslabs")) c->red_vslabs = (int)value;
    else { mgcr_set_error("mgcr_ctx_set_option: unknown option '%s'", key); return MGCR_ERR_ARG; }
    return MGCR_OK;
}

extern "C" int mgcr_ctx_destroy(mgcr_ctx* c) {
    if (!c) return MGCR_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->aux_stream);
    dist_destroy(c);
    mem_trim(c);
    for (auto& kv : c->mem_live) cudaFree(kv.first);
    cudaFree(c->d_partials); cudaFree(c->d_ticket); cudaFree(c->d_scratch); cudaFreeHost(c->h_pinned);
    if (c->h_stage) { cudaFreeHost(c->h_stage); cudaEventDestroy(c->ev_stage[0]); cudaEventDestroy(c->ev_stage[1]); }
    cudaEventDestroy(c->ev_a); cudaEventDestroy(c->ev_b); cudaEventDestroy(c->ev_scal);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->depth_events) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream); cudaStreamDestroy(c->aux_stream);
    delete c;
    return MGCR_OK;
}

extern "C" int mgcr_ctx_sync(mgcr_ctx* c) {
    ARG_CHECK(c, "ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return MGCR_OK;
}

extern "C" int mgcr_ctx_stream(mgcr_ctx* c, void** s) {
    ARG_CHECK(c && s, "ctx/stream_out is NULL");
    *s = (void*)c->stream;
    return MGCR_OK;
}

extern "C" int mgcr_ctx_launch_count(mgcr_ctx* c, int64_t* n) {
    ARG_CHECK(c && n, "ctx/count_out is NULL");
    *n = c->launches;
    return MGCR_OK;
}

enum { PROF_POOL_PAIRS = 8192 };

static void prof_drain(mgcr_ctx* c) {
    if (c->prof_pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (const ProfPending& p : c->prof_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->prof_events[2 * p.ev], c->prof_events[2 * p.ev + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
        // MGCR_PROFILE_BY_SIZE=1 (diagnostic): one class per kernel AND problem size ("gcr_init@2.1e+09"), to see which level of a
        // hierarchy a class loses its bandwidth on
        static const bool by_size = getenv("MGCR_PROFILE_BY_SIZE") && atoi(getenv("MGCR_PROFILE_BY_SIZE")) != 0;
        std::string key = p.name;
        if (by_size) { char buf[32]; snprintf(buf, sizeof buf, "@%.1e", p.bytes); key += buf; }
        ProfEntry& e = c->prof[key];
        e.ms += ms; e.calls++; e.bytes += p.bytes;
    }
    c->prof_pending.clear();
}

void prof_begin(mgcr_ctx* c, const char* name, double bytes) {
    if ((int)c->prof_pending.size() >= PROF_POOL_PAIRS) prof_drain(c);
    int idx = (int)c->prof_pending.size();
    while ((int)c->prof_events.size() < 2 * (idx + 1)) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->prof_events.push_back(e);
    }
    c->prof_pending.push_back({name, bytes, idx});
    cudaEventRecord(c->prof_events[2 * idx], c->stream);
}

void prof_end(mgcr_ctx* c) {
    int idx = c->prof_pending.back().ev;
    cudaEventRecord(c->prof_events[2 * idx + 1], c->stream);
}

extern "C" int mgcr_ctx_set_profile(mgcr_ctx* c, int enabled) {
    ARG_CHECK(c, "ctx is NULL");
    prof_drain(c);
    c->profile = enabled != 0;
    c->prof.clear();
    c->host_alloc_ms = c->host_sync_ms = 0; c->host_alloc_calls = c->host_sync_calls = 0;
    return MGCR_OK;
}

extern "C" int mgcr_ctx_get_profile(mgcr_ctx* c, int cap, const char** names, double* ms, int64_t* calls, double* bytes, int* n_out) {
    ARG_CHECK(c && n_out, "ctx/n_out is NULL");
    prof_drain(c);
    if (c->host_alloc_calls) { ProfEntry& e = c->prof["host_alloc"]; e.ms = c->host_alloc_ms; e.calls = c->host_alloc_calls; e.bytes = 0; }
    if (c->host_sync_calls) { ProfEntry& e = c->prof["host_sync"]; e.ms = c->host_sync_ms; e.calls = c->host_sync_calls; e.bytes = 0; }
    int i = 0;
    for (auto& kv : c->prof) {
        if (i < cap) {
            if (names) names[i] = kv.first.c_str();
            if (ms) ms[i] = kv.second.ms;
            if (calls) calls[i] = kv.second.calls;
            if (bytes) bytes[i] = kv.second.bytes;
        }
        i++;
    }
    *n_out = i;
    return MGCR_OK;
}
