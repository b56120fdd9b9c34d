// examples/stencil_gcr.cpp -- the matrix-free Stencil operator (include/mgcr/Stencil.h) used exactly like a
// DiracOp(&Sparse, k) of the reference (src/main.cpp:845-858): an anisotropic variable-coefficient 3-D problem
// A = diag - H built both ways, applied and solved with GCR; prints KEY VALUE lines for tests/test_gpu_dropin.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "GCR.h"
#include "Stencil.h"

int main() {
    const long nz = 12, ny = 10, nx = 16, V = nz * ny * nx;
    long dims[3] = {nz, ny, nx};
    const double eps[3] = {1e-4, 1e-2, 1.};
    const long stride[3] = {ny * nx, nx, 1};
    // bonds: a deterministic smooth variation, face[d][i] couples site i and i + stride[d]
    std::vector<std::vector<double>> face(3, std::vector<double>(V, 0.));
    std::vector<double> diag(V, 0.5);
    for (long i = 0; i < V; i++) {
        const long c[3] = {i / (ny * nx), (i / nx) % ny, i % nx};
        for (int d = 0; d < 3; d++)
            if (c[d] + 1 < dims[d]) {
                face[d][i] = eps[d] * (1. + 0.3 * std::sin(0.37 * (double)i + d));
                diag[i] += face[d][i];
                diag[i + stride[d]] += face[d][i];
            }
    }
    const double* fp[3] = {face[0].data(), face[1].data(), face[2].data()};
    Stencil<long> A(dims, 3, 1., fp, diag.data());

    // the same entries as the Sparse H a user of the reference would assemble (ascending columns), wrapped as y = diag.x - Hx
    // through Field arithmetic
    long* ROW = (long*)std::malloc(sizeof(long) * (V + 1));
    long* COL = (long*)std::malloc(sizeof(long) * 6 * V);
    std::complex<double>* VAL = (std::complex<double>*)std::malloc(sizeof(std::complex<double>) * 6 * V);
    long nnz = 0;
    for (long i = 0; i < V; i++) {
        const long c[3] = {i / (ny * nx), (i / nx) % ny, i % nx};
        ROW[i] = nnz;
        for (int d = 0; d < 3; d++) if (c[d] > 0) { COL[nnz] = i - stride[d]; VAL[nnz] = face[d][i - stride[d]]; nnz++; }
        for (int d = 2; d >= 0; d--) if (c[d] + 1 < dims[d]) { COL[nnz] = i + stride[d]; VAL[nnz] = face[d][i]; nnz++; }
    }
    ROW[V] = nnz;
    Sparse<long> H(V, V, ROW, COL, VAL);   // adopts the arrays (src/Operator.h:64)

    Field<long> f(dims, 3);
    f.init_rand(1);
    Field<long> y = A(f);
    Field<long> Hf = H(f);
    double err = 0., ref = 0.;
    for (long i = 0; i < V; i++) {
        const std::complex<double> want = diag[i] * f.val_at(i) - Hf.val_at(i);
        err += std::norm(y.val_at(i) - want);
        ref += std::norm(want);
    }
    printf("APPLY_REL %.3e\n", std::sqrt(err / ref));
    printf("VAL_AT %.17g %.17g %.17g\n", A.val_at(5, 5).real(), A.val_at(5, 6).real(), A.val_at(5, 5 + nx).real());

    GCR_Param<long> p(0, 10, 500, 1e-10, false, nullptr, nullptr);
    GCR<long> gcr(&A, &p);
    Field<long> rhs(dims, 3);
    rhs.init_rand(0);
    Field<long> x(dims, 3);
    x.set_zero();
    gcr.solve(rhs, x);
    Field<long> r = rhs - A(x);
    printf("TRUE_RESIDUAL %.3e\n", r.norm() / rhs.norm());
    printf("DONE 1\n");
    return 0;
}
