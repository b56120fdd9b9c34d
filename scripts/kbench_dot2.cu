// scripts/kbench_dot2.cu -- experiment harness (not part of the product): the library's own k_gcr_dot_hist timed (a) alone on
// zero / random data, (b) right after a kernel that wrote Ar (as in the solver, where the operator apply precedes it),
// (c) with cudaMallocAsync storage.   nvcc ... -I mgpreconditionedgcr_b200/csrc scripts/kbench_dot2.cu
#include "kernels_blas.cuh"
void mgcr_set_error(const char*, ...) {}
void prof_begin(mgcr_ctx*, const char*, double) {}
void prof_end(mgcr_ctx*) {}

static const int64_t N = (int64_t)1 << 24;
static __global__ void k_rand(int64_t n, c128* p, unsigned seed) {
    GRID_STRIDE(i, n) { unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = cmake((h & 0xffff) / 65536. - .5, (h >> 16) / 65536. - .5); }
}

template <int NH, int NK, int KS, int U>
static float time_dot(bool writer, c128* Ar, c128* Aps, c128* src, double* part, unsigned* ticket, double* out, int gps) {
    HistList hl; for (int k = 0; k < GCR_CHUNK; k++) hl.slot[k] = k < NH ? k : 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0;
    const int reps = 10;
    for (int w = 0; w < reps + 2; w++) {
        if (writer) k_axpy<<<148 * 8, RED_THREADS>>>(N, cmake(1.0001, 0.), src, src, Ar);
        cudaEventRecord(e0);
        k_gcr_dot_hist<NK, KS><<<148 * gps, RED_THREADS>>>(N, Ar, Aps, N, hl, NH, 0, out, part, ticket, nullptr, 0.);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (w >= 2) total += ms;
    }
    return total / reps;
}

template <int NH, int NK, int KS, int U>
static void report(const char* tag, bool writer, c128* Ar, c128* Aps, c128* src, double* part, unsigned* ticket, double* out) {
    for (int gps : {4, 8}) {
        float ms = time_dot<NH, NK, KS, U>(writer, Ar, Aps, src, part, ticket, out, gps);
        printf("NH=%2d NK=%d KS=%d U=%d %-22s gps=%d %7.1f us %6.0f GB/s\n", NH, NK, KS, U, tag, gps, ms * 1e3, 16. * N * (1 + NH) / (ms * 1e-3) / 1e9);
    }
}

template <int NH, int NK, int KS, int U>
static void all(c128* Ar, c128* Aps, c128* src, c128* ArA, c128* ApsA, double* part, unsigned* ticket, double* out) {
    report<NH, NK, KS, U>("random alone", false, Ar, Aps, src, part, ticket, out);
    report<NH, NK, KS, U>("random after writer", true, Ar, Aps, src, part, ticket, out);
}

int main() {
    c128 *Ar, *Aps, *src, *ArA, *ApsA; double *part, *out; unsigned* ticket;
    cudaMalloc(&Ar, 16 * N); cudaMalloc(&Aps, 16 * N * 16); cudaMalloc(&src, 16 * N);
    cudaMallocAsync(&ArA, 16 * N, 0); ApsA = Aps;
    cudaMalloc(&part, 8 * MAX_RED_BLOCKS * MAX_RED_VALUES); cudaMalloc(&out, 8 * 64); cudaMalloc(&ticket, 16); cudaMemset(ticket, 0, 16);
    k_rand<<<1184, 256>>>(N, src, 3);
    k_rand<<<1184, 256>>>(N, Ar, 1); k_rand<<<1184, 256>>>(N * 16, Aps, 7);
#define ALL(NH, NK, KS, U) all<NH, NK, KS, U>(Ar, Aps, src, ArA, ApsA, part, ticket, out)
    ALL(1, 3, 1, 0); ALL(2, 3, 1, 0); ALL(3, 3, 1, 0); ALL(4, 4, 2, 0); ALL(4, 4, 1, 0); ALL(5, 4, 2, 0); ALL(6, 4, 2, 0); ALL(7, 4, 2, 0); ALL(8, 4, 2, 0);
    ALL(8, 4, 4, 0); ALL(9, 4, 4, 0); ALL(10, 4, 4, 0); ALL(12, 4, 4, 0); ALL(16, 4, 4, 0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
