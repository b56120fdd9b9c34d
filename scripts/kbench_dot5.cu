// scripts/kbench_dot5.cu -- experiment harness (not part of the product): the TMA-staged k_gcr_dot_hist_tma<NH> on n = 2^24,
// run right after a kernel that wrote Ar, for several (elements per thread per tile, stages) choices.
#include "kernels_blas.cuh"
void mgcr_set_error(const char*, ...) {}
void prof_begin(mgcr_ctx*, const char*, double) {}
void prof_end(mgcr_ctx*) {}
static const int64_t N = (int64_t)1 << 24;
static __global__ void k_rand(int64_t n, c128* p, unsigned seed) {
    GRID_STRIDE(i, n) { unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = cmake((h & 0xffff) / 65536. - .5, (h >> 16) / 65536. - .5); }
}
static __global__ void k_ref(int64_t n, const c128* a, const c128* h, double* out) {   // plain check of vector 0
    double x = 0, y = 0;
    GRID_STRIDE(i, n) { c128 t = cmulc(a[i], h[i]); x += t.x; y += t.y; }
    atomicAdd(out, x); atomicAdd(out + 1, y);
}
template <int NH>
static void run(c128* Ar, c128* Aps, c128* src, double* part, unsigned* ticket, double* out) {
    HistList hl; for (int k = 0; k < GCR_CHUNK; k++) hl.slot[k] = k < NH ? k : 0;
    cudaFuncSetAttribute(k_gcr_dot_hist_tma<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int ept : {1, 2, 4}) for (int stages : {2, 3, 4, 6, 8, 12, 16}) {
        size_t smem = (size_t)stages * (1 + NH) * 256 * ept * 16;
        if (smem > 200 * 1024) continue;
        if (smem < 60 * 1024 && stages < 16) continue;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float total = 0; const int reps = 6;
        for (int w = 0; w < reps + 2; w++) {
            k_axpy<<<148 * 8, RED_THREADS>>>(N, cmake(1.0001, 0.), src, src, Ar);
            cudaEventRecord(e0);
            k_gcr_dot_hist_tma<NH><<<148, RED_THREADS, smem>>>(N, Ar, Aps, N, hl, 0, ept, stages, out, part, ticket, nullptr, 0.);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (w >= 2) total += ms;
        }
        float ms = total / reps;
        double h[2], r[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        cudaMemset(out + 40, 0, 16); k_ref<<<592, 256>>>(N, Ar, Aps, out + 40); cudaMemcpy(r, out + 40, 16, cudaMemcpyDeviceToHost);
        printf("NH=%2d ept=%d stages=%2d smem=%3zuK %7.1f us %6.0f GB/s  relerr %.1e %s\n", NH, ept, stages, smem / 1024, ms * 1e3,
               16. * N * (1 + NH) / (ms * 1e-3) / 1e9, fabs(h[0] - r[0]) / fabs(r[0]), cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    c128 *Ar, *Aps, *src; double *part, *out; unsigned* ticket;
    cudaMalloc(&Ar, 16 * N); cudaMalloc(&Aps, 16 * N * 16); cudaMalloc(&src, 16 * N);
    cudaMalloc(&part, 8 * MAX_RED_BLOCKS * MAX_RED_VALUES); cudaMalloc(&out, 8 * 64); cudaMalloc(&ticket, 16); cudaMemset(ticket, 0, 16);
    k_rand<<<1184, 256>>>(N, src, 3); k_rand<<<1184, 256>>>(N * 16, Aps, 7);
    run<1>(Ar, Aps, src, part, ticket, out); run<2>(Ar, Aps, src, part, ticket, out); run<3>(Ar, Aps, src, part, ticket, out);
    run<5>(Ar, Aps, src, part, ticket, out); run<7>(Ar, Aps, src, part, ticket, out); run<10>(Ar, Aps, src, part, ticket, out);
    run<16>(Ar, Aps, src, part, ticket, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
