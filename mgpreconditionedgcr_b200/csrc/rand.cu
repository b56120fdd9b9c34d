// csrc/rand.cu -- Field::init_rand(seed) (reference: src/Fields.h:125-135): the glibc srand/rand stream, drawn on the host
// and uploaded; the GPU never invents its own random numbers.
//
// glibc's rand() is the additive feedback generator r[i] = r[i-3] + r[i-31] mod 2^32 (TYPE_3: seeded by the Lehmer
// sequence 16807 r mod 2^31-1, first 310 outputs dropped, output r >> 1).  It is a LINEAR recurrence over Z/2^32, so the
// state after n draws is M^n times the state now (M = the 31 x 31 companion matrix): a slab [skip, skip+n) of a
// distributed field starts from a jump of 2*skip draws (about 60 matrix-vector products with the cached powers M^(2^k))
// instead of walking the prefix -- in round 1 rank 7 of 8 generated the whole 268 M-draw stream serially -- and the slab
// itself is cut into chunks that OpenMP threads generate independently, each from its own jumped state, into a pinned
// staging ring that is copied to the device while the next chunk is being drawn.
// The restatement is verified against the C library's rand() on every use; if the first draws ever differ (another
// libc), the C library's own srand/rand is used, serially.
#include <omp.h>

#include <algorithm>

#include "common.cuh"

namespace {

struct Mat31 { uint32_t a[31][31]; };

struct GlibcRand {
    uint32_t ring[31];
    int f, b;
    void seed(unsigned int s) {
        if (s == 0) s = 1;
        int32_t word = (int32_t)s;
        ring[0] = (uint32_t)word;
        for (int i = 1; i < 31; i++) {
            long hi = word / 127773, lo = word % 127773;
            word = (int32_t)(16807 * lo - 2836 * hi);
            if (word < 0) word += 2147483647;
            ring[i] = (uint32_t)word;
        }
        f = 3; b = 0;
        for (int i = 0; i < 310; i++) next();
    }
    inline uint32_t next() {
        ring[f] += ring[b];
        const uint32_t out = ring[f] >> 1;
        if (++f == 31) f = 0;
        if (++b == 31) b = 0;
        return out;
    }
    // state as a vector, oldest value first: v[j] = r[i-31+j]; one draw maps v -> (v[1], ..., v[30], v[0] + v[28])
    void get(uint32_t v[31]) const { for (int j = 0; j < 31; j++) v[j] = ring[(f + j) % 31]; }
    void set(const uint32_t v[31]) { for (int j = 0; j < 31; j++) ring[j] = v[j]; f = 0; b = 28; }
    void jump(uint64_t n);
};

// powers[k] = M^(2^k), built once
const Mat31* jump_powers() {
    static Mat31* powers = []() {
        Mat31* p = new Mat31[64];
        memset(&p[0], 0, sizeof(Mat31));
        for (int j = 0; j < 30; j++) p[0].a[j][j + 1] = 1;
        p[0].a[30][0] = 1; p[0].a[30][28] = 1;
        for (int k = 1; k < 64; k++) {
            const Mat31& m = p[k - 1];
            for (int i = 0; i < 31; i++)
                for (int j = 0; j < 31; j++) {
                    uint32_t s = 0;
                    for (int l = 0; l < 31; l++) s += m.a[i][l] * m.a[l][j];
                    p[k].a[i][j] = s;
                }
        }
        return p;
    }();
    return powers;
}

void GlibcRand::jump(uint64_t n) {
    if (n == 0) return;
    const Mat31* pw = jump_powers();
    uint32_t v[31], w[31];
    get(v);
    for (int k = 0; k < 64 && (n >> k); k++) {
        if (!((n >> k) & 1)) continue;
        for (int i = 0; i < 31; i++) {
            uint32_t s = 0;
            for (int l = 0; l < 31; l++) s += pw[k].a[i][l] * v[l];
            w[i] = s;
        }
        memcpy(v, w, sizeof v);
    }
    set(v);
}

// the restated generator reproduces the C library's rand(), and a jump lands where the walk does
bool glibc_rand_matches(int seed) {
    GlibcRand walk, jumped;
    walk.seed((unsigned int)seed);
    jumped = walk;
    srand(seed);
    for (int i = 0; i < 64; i++) if ((int)walk.next() != rand()) return false;
    jumped.jump(64);
    for (int i = 0; i < 31; i++) if (walk.next() != jumped.next()) return false;
    return true;
}

// elements [skip, skip+n) of Field::init_rand(seed) into a host buffer: element = complex((rand()%2000)/1000.-1, ...), the
// IMAGINARY argument evaluated first (g++), SURVEY.md 8a row a9
void fill_host(int seed, int64_t skip, int64_t n, c128* h, bool restated, int nthreads) {
    if (!restated) {
        srand(seed);
        for (int64_t i = 0; i < 2 * skip; i++) (void)rand();
        for (int64_t i = 0; i < n; i++) {
            double im = (rand() % 2000) / 1000. - 1;
            double re = (rand() % 2000) / 1000. - 1;
            h[i] = cmake(re, im);
        }
        return;
    }
    GlibcRand g0;
    g0.seed((unsigned int)seed);
    const int64_t chunk = 1 << 16;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    (void)jump_powers();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t c = 0; c < nchunks; c++) {
        GlibcRand g = g0;
        const int64_t e0 = c * chunk, e1 = std::min(n, e0 + chunk);
        g.jump(2 * (uint64_t)(skip + e0));
        for (int64_t i = e0; i < e1; i++) {
            double im = (int)(g.next() % 2000) / 1000. - 1;
            double re = (int)(g.next() % 2000) / 1000. - 1;
            h[i] = cmake(re, im);
        }
    }
}

}  // namespace

// host-only entry point (no device needed): the stream itself, for callers that stage their own uploads and for the CPU tests
extern "C" int mgcr_rand_stream(int seed, int64_t skip, int64_t n, mgcr_c128* h_out) {
    ARG_CHECK(skip >= 0 && n >= 0 && (n == 0 || h_out), "mgcr_rand_stream: bad argument");
    fill_host(seed, skip, n, (c128*)h_out, glibc_rand_matches(seed), std::max(1, std::min(omp_get_num_procs(), 16)));
    return MGCR_OK;
}

int vec_init_rand_slab(mgcr_ctx* ctx, int seed, int64_t skip, int64_t n, c128* d_out) {
    if (n == 0) return MGCR_OK;
    HostTimer ht_total(&ctx->rand_seconds_ms, &ctx->rand_calls);
    // persistent pinned staging ring of the context: two halves, one being filled while the other is copied
    const int64_t half = (int64_t)1 << 21;   // 2 Mi elements = 32 MB
    if (!ctx->h_stage) {
        CUDA_TRY(cudaMallocHost(&ctx->h_stage, sizeof(c128) * 2 * (size_t)half));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_stage[0], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_stage[1], cudaEventDisableTiming));
    }
    const bool restated = glibc_rand_matches(seed);
    // one process per GPU on one host: the ranks draw their slabs at the same time and share the cores
    const int nthreads = std::max(1, std::min(omp_get_num_procs() / std::max(1, ctx->nranks), 16));
    c128* stage = (c128*)ctx->h_stage;
    int which = 0;
    for (int64_t e0 = 0; e0 < n; e0 += half, which ^= 1) {
        const int64_t cnt = std::min(half, n - e0);
        if (e0 >= 2 * half) CUDA_TRY(cudaEventSynchronize(ctx->ev_stage[which]));   // the copy that last used this half is done
        if (!restated && e0 > 0) {   // libc fallback: serial stream, continue where the previous chunk stopped
            c128* h = stage + (size_t)which * half;
            for (int64_t i = 0; i < cnt; i++) {
                double im = (rand() % 2000) / 1000. - 1;
                double re = (rand() % 2000) / 1000. - 1;
                h[i] = cmake(re, im);
            }
        } else {
            fill_host(seed, skip + e0, cnt, stage + (size_t)which * half, restated, nthreads);
        }
        CUDA_TRY(cudaMemcpyAsync(d_out + e0, stage + (size_t)which * half, sizeof(c128) * (size_t)cnt, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev_stage[which], ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return MGCR_OK;
}
