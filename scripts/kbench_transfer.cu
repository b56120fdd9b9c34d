// scripts/kbench_transfer.cu -- experiment harness (NOT part of the product; prepared for round 2).
// restrict / prolong of the multigrid cycle sit at 0.65-0.83 of the copy peak because the prolongator layout
// [aggregate][vector][dof in aggregate] makes the fine-lattice side of both kernels 64-byte runs (4 sites of an aggregate
// row).  Candidate: keep the prolongator as ne fine-lattice vectors P_C[e][site] ("the reference's near-null vectors, chopped
// per aggregate, stored once") so that EVERY access of both kernels is a full 512-byte request in lattice order:
//   prolong_c : thread per site, x[site] += sum_e xc[block(site)*ne + e] * P_C[e][site]
//   restrict_c: CTA = 4 x 4 x 32 sites (8 aggregates along x), warp = one row of 32 sites; shuffle over the 4 sites of an
//               aggregate row, shared memory over the 16 rows
// against the product's kernels (same arithmetic as mg.cu: k_prolong, k_restrict_warp<2>) on an n^3 lattice of 4^3 aggregates,
// ne = 4.  Prints time, GB/s of algorithmic bytes 16*V*(1+ne) (+16*V for the += of prolong) and the difference of results.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mgpreconditionedgcr_b200/csrc scripts/kbench_transfer.cu -o /tmp/kbt && /tmp/kbt 256
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
void mgcr_set_error(const char*, ...) {}
void prof_begin(mgcr_ctx*, const char*, double) {}
void prof_end(mgcr_ctx*) {}

enum { NE = 4, SUB = 4, BS = SUB * SUB * SUB };

static __global__ void k_rand(int64_t n, c128* p, unsigned seed) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = cmake((h & 0xffff) / 65536. - .5, (h >> 16) / 65536. - .5);
    }
}
// site of (aggregate b, offset q) and back, n^3 lattice, row-major (src/Mesh.h:146-154, 270-293)
__host__ __device__ inline int64_t site_of(int64_t n, int64_t b, int q) {
    const int64_t nb = n / SUB;
    const int64_t bx = b % nb, by = (b / nb) % nb, bz = b / (nb * nb);
    const int qx = q % SUB, qy = (q / SUB) % SUB, qz = q / (SUB * SUB);
    return ((bz * SUB + qz) * n + by * SUB + qy) * n + bx * SUB + qx;
}
// layout A -> layout C
static __global__ void k_to_c(int64_t n, int64_t nblocks, const c128* PA, c128* PC) {
    const int64_t V = n * n * n;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nblocks * NE * BS; t += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(t % BS); const int e = (int)((t / BS) % NE); const int64_t b = t / (BS * NE);
        PC[(int64_t)e * V + site_of(n, b, q)] = PA[t];
    }
}
// ---- product form (mg.cu) ----
static __global__ void __launch_bounds__(256) k_prolong_a(int64_t n, int64_t total, const c128* __restrict__ P, const c128* __restrict__ xc, c128* __restrict__ xf) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t b = t / BS; const int q = (int)(t - b * BS);
    const c128* pv = P + b * NE * BS + q;
    const c128* a = xc + b * NE;
    c128 acc = cmake(0., 0.);
#pragma unroll
    for (int e = 0; e < NE; e++) acc = cadd(acc, cmul(__ldg(a + e), ld_stream(pv + e * BS)));
    c128* dst = xf + site_of(n, b, q);
    *dst = cadd(*dst, acc);
}
static __global__ void __launch_bounds__(256) k_restrict_a(int64_t n, int64_t nblocks, const c128* __restrict__ P, const c128* __restrict__ xf, c128* __restrict__ xc) {
    const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (b >= nblocks) return;
    const int lane = threadIdx.x & 31;
    c128 xs[2];
#pragma unroll
    for (int j = 0; j < 2; j++) xs[j] = __ldg(xf + site_of(n, b, lane + 32 * j));
#pragma unroll
    for (int e = 0; e < NE; e++) {
        const c128* pv = P + (b * NE + e) * BS;
        double sr = 0., si = 0.;
#pragma unroll
        for (int j = 0; j < 2; j++) { c128 t = cmulc(ld_stream(pv + lane + 32 * j), xs[j]); sr += t.x; si += t.y; }
        sr = warp_sum(sr); si = warp_sum(si);
        if (lane == 0) xc[b * NE + e] = cmake(sr, si);
    }
}

// ---- product layout, a warp works on AW aggregates that are neighbours along x: the fine-lattice side becomes runs of
// 64*AW bytes (AW = 1: the product's kernels), the prolongator side 64-byte pieces of AW contiguous 4 KB chunks ----
template <int AW, bool PERSIST>
static __global__ void __launch_bounds__(256) k_prolong_w(int64_t n, int64_t nblocks, const c128* __restrict__ P, const c128* __restrict__ xc, c128* __restrict__ xf) {
    constexpr int RPT = 8 / AW;                    // aggregate rows (of 4 sites) per aggregate and trip
    const int lane = threadIdx.x & 31;
    const int ox = lane & 3, a = (lane >> 2) % AW, r0 = lane / (4 * AW);
    const int64_t ngroups = nblocks / AW;
    const int64_t nwarps = PERSIST ? ((int64_t)gridDim.x * blockDim.x) >> 5 : ngroups;
    for (int64_t gidx = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; gidx < ngroups; gidx += nwarps) {
        const int64_t b = gidx * AW + a;
        const c128* pb = P + b * NE * BS;
        c128 ce[NE];
#pragma unroll
        for (int e = 0; e < NE; e++) ce[e] = __ldg(xc + b * NE + e);
#pragma unroll
        for (int j = 0; j < 16 / RPT; j += 2) {
            const int qa = (j * RPT + r0) * 4 + ox, qb = ((j + 1) * RPT + r0) * 4 + ox;
            c128* da = xf + site_of(n, b, qa);
            c128* db = xf + site_of(n, b, qb);
            const c128 olda = *da, oldb = *db;
            c128 acca = cmake(0., 0.), accb = cmake(0., 0.);
#pragma unroll
            for (int e = 0; e < NE; e++) {
                acca = cadd(acca, cmul(ce[e], ld_stream(pb + e * BS + qa)));
                accb = cadd(accb, cmul(ce[e], ld_stream(pb + e * BS + qb)));
            }
            *da = cadd(olda, acca);
            *db = cadd(oldb, accb);
        }
        if (!PERSIST) break;
    }
}
template <int AW, bool PERSIST>
static __global__ void __launch_bounds__(256) k_restrict_w(int64_t n, int64_t nblocks, const c128* __restrict__ P, const c128* __restrict__ xf, c128* __restrict__ xc) {
    constexpr int RPT = 8 / AW;
    const int lane = threadIdx.x & 31;
    const int ox = lane & 3, a = (lane >> 2) % AW, r0 = lane / (4 * AW);
    const int64_t ngroups = nblocks / AW;
    const int64_t nwarps = PERSIST ? ((int64_t)gridDim.x * blockDim.x) >> 5 : ngroups;
    for (int64_t gidx = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; gidx < ngroups; gidx += nwarps) {
        const int64_t b = gidx * AW + a;
        const c128* pb = P + b * NE * BS;
        double sr[NE], si[NE];
#pragma unroll
        for (int e = 0; e < NE; e++) { sr[e] = 0.; si[e] = 0.; }
#pragma unroll
        for (int j = 0; j < 16 / RPT; j++) {
            const int q = (j * RPT + r0) * 4 + ox;
            const c128 xv = __ldg(xf + site_of(n, b, q));
#pragma unroll
            for (int e = 0; e < NE; e++) { c128 t = cmulc(ld_stream(pb + e * BS + q), xv); sr[e] += t.x; si[e] += t.y; }
        }
        // lanes of one aggregate: bits 0-1 (ox) and the row bits above the aggregate bits
#pragma unroll
        for (int e = 0; e < NE; e++) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                if (m >= 4 && m < 4 * AW) continue;
                sr[e] += __shfl_xor_sync(0xffffffffu, sr[e], m); si[e] += __shfl_xor_sync(0xffffffffu, si[e], m);
            }
        }
        if (ox == 0 && r0 == 0) {
#pragma unroll
            for (int e = 0; e < NE; e++) xc[b * NE + e] = cmake(sr[e], si[e]);
        }
        if (!PERSIST) break;
    }
}
// ---- candidate form: lattice order everywhere ----
static __global__ void __launch_bounds__(256) k_prolong_c(int64_t n, const c128* __restrict__ PC, const c128* __restrict__ xc, c128* __restrict__ xf) {
    const int64_t V = n * n * n, nb = n / SUB;
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= V) return;
    const int64_t x = s % n, y = (s / n) % n, z = s / (n * n);
    const int64_t b = ((z / SUB) * nb + y / SUB) * nb + x / SUB;
    const c128* a = xc + b * NE;
    c128 acc = cmake(0., 0.);
#pragma unroll
    for (int e = 0; e < NE; e++) acc = cadd(acc, cmul(__ldg(a + e), ld_stream(PC + (int64_t)e * V + s)));
    xf[s] = cadd(xf[s], acc);
}
// CTA = 16 warps = the 16 (z, y) rows of a strip of 8 aggregates along x (32 sites per row)
static __global__ void __launch_bounds__(512) k_restrict_c(int64_t n, const c128* __restrict__ PC, const c128* __restrict__ xf, c128* __restrict__ xc) {
    __shared__ double red[16][8][NE][2];
    const int64_t V = n * n * n, nb = n / SUB;
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;           // row = qz * 4 + qy
    const int64_t strips = n / 32;
    const int64_t strip = blockIdx.x % strips, by = (blockIdx.x / strips) % nb, bz = blockIdx.x / (strips * nb);
    const int64_t s = ((bz * SUB + row / SUB) * n + by * SUB + row % SUB) * n + strip * 32 + lane;
    const c128 xv = __ldg(xf + s);
#pragma unroll
    for (int e = 0; e < NE; e++) {
        c128 t = cmulc(ld_stream(PC + (int64_t)e * V + s), xv);
        double sr = t.x, si = t.y;
        sr += __shfl_xor_sync(0xffffffffu, sr, 1); si += __shfl_xor_sync(0xffffffffu, si, 1);
        sr += __shfl_xor_sync(0xffffffffu, sr, 2); si += __shfl_xor_sync(0xffffffffu, si, 2);
        if ((lane & 3) == 0) { red[row][lane >> 2][e][0] = sr; red[row][lane >> 2][e][1] = si; }
    }
    __syncthreads();
    if (threadIdx.x < 8 * NE * 2) {
        const int c = threadIdx.x & 1, e = (threadIdx.x >> 1) % NE, a = threadIdx.x / (2 * NE);
        double sum = 0.;
#pragma unroll
        for (int r = 0; r < 16; r++) sum += red[r][a][e][c];
        const int64_t b = (bz * nb + by) * nb + strip * 8 + a;
        ((double*)(xc + b * NE + e))[c] = sum;
    }
}

template <class F> static float time_ms(F f, int reps = 10) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) f();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 256;
    if (n % 32) { printf("n must be a multiple of 32\n"); return 1; }
    const int64_t V = n * n * n, nblocks = V / BS;
    c128 *PA, *PC, *xf, *xf2, *xc, *xc2;
    cudaMalloc(&PA, 16 * V * NE); cudaMalloc(&PC, 16 * V * NE); cudaMalloc(&xf, 16 * V); cudaMalloc(&xf2, 16 * V);
    cudaMalloc(&xc, 16 * nblocks * NE); cudaMalloc(&xc2, 16 * nblocks * NE);
    k_rand<<<1184, 256>>>(V * NE, PA, 3); k_rand<<<1184, 256>>>(V, xf, 5); k_rand<<<1184, 256>>>(nblocks * NE, xc, 7);
    k_to_c<<<1184, 256>>>(n, nblocks, PA, PC);
    cudaMemcpy(xf2, xf, 16 * V, cudaMemcpyDeviceToDevice);
    // correctness: one application each
    k_restrict_a<<<(unsigned)((nblocks * 32 + 255) / 256), 256>>>(n, nblocks, PA, xf, xc);
    k_restrict_c<<<(unsigned)(nblocks / 8), 512>>>(n, PC, xf, xc2);
    k_prolong_a<<<(unsigned)((V + 255) / 256), 256>>>(n, V, PA, xc, xf);
    k_prolong_c<<<(unsigned)((V + 255) / 256), 256>>>(n, PC, xc, xf2);
    std::vector<c128> h1((size_t)V), h2((size_t)V), c1((size_t)(nblocks * NE)), c2((size_t)(nblocks * NE));
    cudaMemcpy(h1.data(), xf, 16 * V, cudaMemcpyDeviceToHost); cudaMemcpy(h2.data(), xf2, 16 * V, cudaMemcpyDeviceToHost);
    cudaMemcpy(c1.data(), xc, 16 * nblocks * NE, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data(), xc2, 16 * nblocks * NE, cudaMemcpyDeviceToHost);
    double dp = 0., dr = 0., nr = 0.;
    for (int64_t i = 0; i < V; i++) dp = fmax(dp, fmax(fabs(h1[i].x - h2[i].x), fabs(h1[i].y - h2[i].y)));
    for (int64_t i = 0; i < nblocks * NE; i++) { dr = fmax(dr, fmax(fabs(c1[i].x - c2[i].x), fabs(c1[i].y - c2[i].y))); nr = fmax(nr, fabs(c1[i].x)); }
    printf("n=%lld: prolong max |a - c| = %.3e (same products in the same order: expect 0), restrict max |a - c| / max = %.3e\n", (long long)n, dp, dr / nr);
    const double bytes_r = 16. * V * (1 + NE) + 16. * nblocks * NE, bytes_p = bytes_r + 16. * V;
    float t;
    t = time_ms([&] { k_restrict_a<<<(unsigned)((nblocks * 32 + 255) / 256), 256>>>(n, nblocks, PA, xf, xc); });
    printf("restrict  product layout : %8.1f us %6.0f GB/s\n", t * 1e3, bytes_r / (t * 1e-3) / 1e9);
    t = time_ms([&] { k_restrict_c<<<(unsigned)(nblocks / 8), 512>>>(n, PC, xf, xc2); });
    printf("restrict  lattice layout : %8.1f us %6.0f GB/s\n", t * 1e3, bytes_r / (t * 1e-3) / 1e9);
    t = time_ms([&] { k_prolong_a<<<(unsigned)((V + 255) / 256), 256>>>(n, V, PA, xc, xf); });
    printf("prolong   product layout : %8.1f us %6.0f GB/s\n", t * 1e3, bytes_p / (t * 1e-3) / 1e9);
    t = time_ms([&] { k_prolong_c<<<(unsigned)((V + 255) / 256), 256>>>(n, PC, xc, xf2); });
    printf("prolong   lattice layout : %8.1f us %6.0f GB/s\n", t * 1e3, bytes_p / (t * 1e-3) / 1e9);

    // warp-per-AW-aggregates variants of the product layout
    {
        c128* xc3; cudaMalloc(&xc3, 16 * nblocks * NE);
        c128* xf3; cudaMalloc(&xf3, 16 * V);
        std::vector<c128> c3((size_t)(nblocks * NE)), h3((size_t)V);
        auto check_r = [&](const char* name) {
            cudaMemcpy(c3.data(), xc3, 16 * nblocks * NE, cudaMemcpyDeviceToHost);
            double d = 0.;
            for (int64_t i = 0; i < nblocks * NE; i++) d = fmax(d, fmax(fabs(c1[i].x - c3[i].x), fabs(c1[i].y - c3[i].y)));
            printf("   %s max |diff| / max = %.3e\n", name, d / nr);
        };
#define RUN_R(AW, PERS, GRID) do { \
        cudaMemset(xc3, 0, 16 * nblocks * NE); \
        k_restrict_w<AW, PERS><<<GRID, 256>>>(n, nblocks, PA, xf2, xc3); \
        t = time_ms([&] { k_restrict_w<AW, PERS><<<GRID, 256>>>(n, nblocks, PA, xf2, xc3); }); \
        printf("restrict  warp x %d aggregates %s : %8.1f us %6.0f GB/s\n", AW, PERS ? "persistent" : "one-shot  ", t * 1e3, bytes_r / (t * 1e-3) / 1e9); } while (0)
#define RUN_P(AW, PERS, GRID) do { \
        t = time_ms([&] { k_prolong_w<AW, PERS><<<GRID, 256>>>(n, nblocks, PA, xc, xf3); }); \
        printf("prolong   warp x %d aggregates %s : %8.1f us %6.0f GB/s\n", AW, PERS ? "persistent" : "one-shot  ", t * 1e3, bytes_p / (t * 1e-3) / 1e9); } while (0)
        // (xf2 holds the candidate's prolonged vector; restrict of it by variant vs by the product kernel)
        k_restrict_a<<<(unsigned)((nblocks * 32 + 255) / 256), 256>>>(n, nblocks, PA, xf2, xc);
        cudaMemcpy(c1.data(), xc, 16 * nblocks * NE, cudaMemcpyDeviceToHost);
        nr = 0.; for (int64_t i = 0; i < nblocks * NE; i++) nr = fmax(nr, fabs(c1[i].x));
        const unsigned pg = 148 * 4;
        RUN_R(1, false, (unsigned)((nblocks / 1 * 32 + 255) / 256)); check_r("AW=1");
        RUN_R(2, false, (unsigned)((nblocks / 2 * 32 + 255) / 256)); check_r("AW=2");
        RUN_R(4, false, (unsigned)((nblocks / 4 * 32 + 255) / 256)); check_r("AW=4");
        RUN_R(8, false, (unsigned)((nblocks / 8 * 32 + 255) / 256)); check_r("AW=8");
        RUN_R(1, true, pg); RUN_R(2, true, pg); RUN_R(4, true, pg); RUN_R(8, true, pg); check_r("AW=8 persistent");
        // prolong: bit-exactness of one application against the product kernel
        cudaMemcpy(xf3, xf2, 16 * V, cudaMemcpyDeviceToDevice);
        cudaMemcpy(xf, xf2, 16 * V, cudaMemcpyDeviceToDevice);
        k_prolong_a<<<(unsigned)((V + 255) / 256), 256>>>(n, V, PA, xc, xf);
        k_prolong_w<4, true><<<pg, 256>>>(n, nblocks, PA, xc, xf3);
        cudaMemcpy(h1.data(), xf, 16 * V, cudaMemcpyDeviceToHost); cudaMemcpy(h3.data(), xf3, 16 * V, cudaMemcpyDeviceToHost);
        double d = 0.; for (int64_t i = 0; i < V; i++) d = fmax(d, fmax(fabs(h1[i].x - h3[i].x), fabs(h1[i].y - h3[i].y)));
        printf("   prolong AW=4 persistent vs product: max |diff| = %.3e (expect 0)\n", d);
        RUN_P(1, false, (unsigned)((nblocks / 1 * 32 + 255) / 256));
        RUN_P(2, false, (unsigned)((nblocks / 2 * 32 + 255) / 256));
        RUN_P(4, false, (unsigned)((nblocks / 4 * 32 + 255) / 256));
        RUN_P(8, false, (unsigned)((nblocks / 8 * 32 + 255) / 256));
        RUN_P(1, true, pg); RUN_P(2, true, pg); RUN_P(4, true, pg); RUN_P(8, true, pg);
        RUN_P(4, true, 148 * 3); RUN_P(4, true, 148 * 6); RUN_P(8, true, 148 * 6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
