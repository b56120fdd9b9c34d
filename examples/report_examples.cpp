// examples/report_examples.cpp -- the usage examples of the reference's report (SemesterProject.pdf sections 3.1 and
// 3.2.3-3.2.5, pp. 6-9) written against the drop-in headers: Field methods, Operator methods, the plain GCR solve and the
// MG-preconditioned solve with exactly the constructor calls the report shows.  Differences forced on any build: `long`
// dims (the report's `int dims[6]` does not convert to `Field<long>`'s `const long*`), field_size() for the report's
// get_size(), the 4^4 lattice of the one matrix the repository ships ("4x4parsed.txt"; the report uses the missing 8^4
// file) and therefore aggregates of 2^4.
//
//   MGCR_DATA_DIR=<dir with 4x4parsed.txt> ./report_examples
#include <complex>
#include <cstdio>

#include "MG.h"
#include "Parse.h"

int main() {
    // ---- 3.2.3 Field methods ----
    int ndim = 6;
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Field<long> fermion1(dims, ndim);
    Field<long> fermion2(dims, ndim);
    fermion1.init_rand(1);
    fermion2.init_rand(2);
    printf("Number of elements in the fermion field = %ld\n", fermion1.field_size());
    std::complex<double> dot_product = fermion1.dot(fermion2);
    printf("Dot product between 2 fermion fields = (%.4e, %.4e)\n", dot_product.real(), dot_product.imag());
    Field<long> addition = fermion1 + fermion2;
    printf("Norm of sum of 2 fermions = %.4e\n", addition.norm());

    // ---- 3.1 / 3.2.4 Operator methods ----
    auto Op = new Sparse(read_data("4x4parsed.txt"));
    Field result = (*Op)(fermion1);
    printf("Norm of D f = %.4e\n", result.norm());

    // ---- 3.2.5 solvers ----
    Field<long> rhs(dims, ndim);
    rhs.init_rand();
    const double k = 0.05 + 8 * ((0.17865 - 0.05) / 10.);
    auto Dirac = new DiracOp<long>(Op, k);
    /* example 1: GCR solver */
    GCR_Param<long> gcr_param(0, 5, 4000, 1e-13, false, nullptr, nullptr);
    GCR gcr(Dirac, &gcr_param);          // gcr is also an Operator, approximately Dirac^-1
    Field resu1 = gcr(rhs);
    /* example 2: MG preconditioned GCR solver */
    GCR_Param<long> eigen(0, 10, 10, 1e-8, false, nullptr, nullptr);
    GCR_Param<long> coarse(0, 10, 50, 1e-2, false, nullptr, nullptr);
    GCR_Param<long> smooth(0, 10, 0, 1e-8, false, nullptr, nullptr);
    eigen.std_conj = coarse.std_conj = smooth.std_conj = true;   // addition, see k_critical_mg_precond.cpp
    auto solver_c = new GCR(&coarse);
    auto solver_s = new GCR(&smooth);
    Mesh mesh(dims, ndim);
    MG_Param<long> mg_param(mesh, 2, 10, &eigen, solver_c, solver_s, 1, nullptr, nullptr);
    auto mg = new MG(&mg_param);
    mg->initialise(Dirac);               // the report leaves this to GCR::solve, where the reference has it commented out
    GCR_Param<long> gcr_param_precond(0, 2, 2000, 1e-13, false, nullptr, mg);
    gcr_param_precond.std_conj = true;
    GCR grc_precond(Dirac, &gcr_param_precond);
    Field<long> resu2(dims, ndim);
    resu2.set_zero();
    grc_precond.solve(rhs, resu2);
    printf("MG-GCR iterations = %d, true residual = %.3e\n", grc_precond.iterations(), (rhs - (*Dirac)(resu2)).norm() / rhs.norm());
    printf("DONE 1\n");
    delete mg; delete solver_c; delete solver_s; delete Dirac; delete Op;
    return 0;
}
