// scripts/kbench_dot.cu -- experiment harness (not part of the product): variants of the batched history inner product
// <Ar, Aps[k]> k < NH (csrc/kernels_blas.cuh k_gcr_dot_hist) timed alone on n = 2^24 complex128 elements.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false scripts/kbench_dot.cu -o /tmp/kbench_dot && /tmp/kbench_dot
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef double2 c128;
__device__ __forceinline__ c128 ld_nc(const c128* p) {
    c128 r;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ c128 cmulc(c128 a, c128 b) { c128 r; r.x = a.x * b.x + a.y * b.y; r.y = a.x * b.y - a.y * b.x; return r; }
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// KS k-groups per CTA: thread group g = threadIdx.x / (THREADS/KS) handles history vectors k = g, g + KS, ... for the
// element range of the whole CTA (Ar is then read KS times, the repeats hit L1/L2)
template <int NH, int U, int MINB, bool NC, int KS, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB) k_dot(int64_t n, const c128* __restrict__ Ar, const c128* __restrict__ Aps, int64_t stride,
                                                       double* out) {
    constexpr int NK = (NH + KS - 1) / KS;   // history vectors per group
    constexpr int GT = THREADS / KS;         // threads per group
    const int g = threadIdx.x / GT, tl = threadIdx.x % GT;
    double v[2 * NK];
#pragma unroll
    for (int k = 0; k < 2 * NK; k++) v[k] = 0.;
    const int64_t T = (int64_t)gridDim.x * GT;
    int64_t i0 = blockIdx.x * (int64_t)GT + tl;
    for (; i0 + (U - 1) * T < n; i0 += T * U) {
        c128 a[U], h[U][NK];
#pragma unroll
        for (int u = 0; u < U; u++) {
            a[u] = NC ? ld_nc(Ar + i0 + u * T) : Ar[i0 + u * T];
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const int kk = g + k * KS;
                if (kk < NH) h[u][k] = NC ? ld_nc(Aps + (int64_t)kk * stride + i0 + u * T) : Aps[(int64_t)kk * stride + i0 + u * T];
                else h[u][k] = make_double2(0., 0.);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < NK; k++) {
                c128 t = cmulc(a[u], h[u][k]);
                v[2 * k] += t.x; v[2 * k + 1] += t.y;
            }
    }
#pragma unroll
    for (int k = 0; k < 2 * NK; k++) {
        double s = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(out + 2 * (g + (k / 2) * KS) + (k & 1), s);
    }
}

static c128* d_Ar; static c128* d_Aps; static double* d_out;
static const int64_t N = (int64_t)1 << 24;

template <int NH, int U, int MINB, bool NC, int KS, int THREADS>
static void run(const char* tag, int gps) {
    int grid = 148 * gps;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) k_dot<NH, U, MINB, NC, KS, THREADS><<<grid, THREADS>>>(N, d_Ar, d_Aps, N, d_out);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int w = 0; w < reps; w++) k_dot<NH, U, MINB, NC, KS, THREADS><<<grid, THREADS>>>(N, d_Ar, d_Aps, N, d_out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    double gb = 16. * N * (1 + NH) * reps / (ms * 1e-3) / 1e9;
    printf("NH=%2d %-28s gps=%2d  %7.1f us  %6.0f GB/s %s\n", NH, tag, gps, ms * 1e3 / reps, gb, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

template <int NH>
static void sweep() {
    for (int gps : {4, 8, 16}) {
        run<NH, 1, 1, true, 1, 256>("U1 nc", gps);
        if (NH <= 8) run<NH, 2, 1, true, 1, 256>("U2 nc", gps);
        if (NH <= 4) run<NH, 4, 1, true, 1, 256>("U4 nc", gps);
        run<NH, 1, 1, false, 1, 256>("U1 plain", gps);
        run<NH, 1, 3, true, 1, 256>("U1 nc minb3", gps);
        run<NH, 1, 4, true, 1, 256>("U1 nc minb4", gps);
        run<NH, 1, 1, true, 1, 512>("U1 nc t512", gps);
        if (NH >= 4) {
            run<NH, 1, 1, true, 2, 256>("U1 nc ksplit2", gps);
            run<NH, 2, 1, true, 2, 256>("U2 nc ksplit2", gps);
            run<NH, 1, 1, true, 2, 512>("U1 nc ksplit2 t512", gps);
        }
        if (NH >= 8) {
            run<NH, 1, 1, true, 4, 512>("U1 nc ksplit4 t512", gps);
            run<NH, 2, 1, true, 4, 512>("U2 nc ksplit4 t512", gps);
        }
    }
}

int main() {
    cudaMalloc(&d_Ar, sizeof(c128) * N);
    cudaMalloc(&d_Aps, sizeof(c128) * N * 10);
    cudaMalloc(&d_out, sizeof(double) * 64);
    cudaMemset(d_Ar, 0, sizeof(c128) * N);
    cudaMemset(d_Aps, 0, sizeof(c128) * N * 10);
    cudaMemset(d_out, 0, sizeof(double) * 64);
    sweep<1>(); sweep<2>(); sweep<3>(); sweep<5>(); sweep<8>(); sweep<10>();
    return 0;
}
