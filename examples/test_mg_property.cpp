// examples/test_mg_property.cpp -- the reference's live diagnostics compiled against the drop-in headers:
//   test_MG_property()  src/main.cpp:877-918  (MG::test_MG + the R T = 1 / T R projector identities)
//   test_hermiticity()  src/main.cpp:541-570  (<v, M w> vs <M v, w> and the value-by-value check)
// The function bodies are the reference's, statement for statement; the only edits are the lattice (the repository ships the
// 4^4 sample, main.cpp names the missing 8^4 file: SURVEY.md fact 10), the aggregate size that goes with it, and KEY VALUE
// lines for tests/test_gpu_dropin.py at the end of each function.  test_gamma5_hermiticity() is an addition: the sample operator
// is not Hermitian but gamma5-Hermitian, which is what the chirality doubling of src/MG.h:316-329 relies on.
//
//   MGCR_DATA_DIR=<dir with 4x4parsed.txt> ./test_mg_property
#include <chrono>
#include <cstdio>

#include "MG.h"
#include "Parse.h"

void test_MG_property() {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Mesh mesh(dims, 6);
    auto D = new Sparse(read_data("4x4parsed.txt"));
    auto Dirac = new DiracOp(D, 0.1);

    GCR_Param<long> eigen(0, 10, 10, 1e-8, false, nullptr, nullptr);
    GCR_Param<long> coarse(0, 10, 1, 1e-8, false, nullptr, nullptr);
    GCR_Param<long> smooth(0, 10, 1, 1e-8, false, nullptr, nullptr);
    auto solver_coarse = new GCR(&coarse);
    auto solver_smooth = new GCR(&smooth);
    MG_Param<long> param(mesh, 2, 2, &eigen, solver_coarse, solver_smooth, 1, nullptr, nullptr);
    auto mg = new MG(Dirac, &param);

    mg->test_MG(Dirac);
    //mg->test_by_value(Dirac);



    Field<long> rhs(dims, 6);
    rhs.init_rand(42);

    // RT Rf = Rf   (RT is identity)
    Field inter1 = mg->restrict(rhs);
    Field inter2 = mg->expand(inter1);
    Field inter3 = mg->restrict(inter2);

    // TR TR f = TR f   (TR is a projection)
    Field inter11 = mg->restrict(inter2);
    Field inter22 = mg->expand(inter11);

    printf("RT - Id identity test difference = %.5e\n", (inter2 - inter22).norm());
    printf("TR TR - TR projector test difference = %.5e\n", (inter3 - inter1).norm());
    //printf("%.5e\t%.5e", inter1.norm(), inter2.norm());

    printf("KV_TEST_MG_GALERKIN %.6e\n", mg->last_test_galerkin);
    printf("KV_TEST_MG_PROJECTOR %.6e\n", mg->last_test_projector);
    printf("KV_RT_IDENTITY %.6e\n", (inter2 - inter22).norm());
    printf("KV_TR_PROJECTOR %.6e\n", (inter3 - inter1).norm());

    delete mg;
    delete solver_coarse;
    delete solver_smooth;
    delete Dirac;
    delete D;
}

void test_hermiticity() {
    auto mat = new Sparse(read_data("4x4parsed.txt"));
    //auto Dirac = new Sparse()

    long dims[1] = {(*mat).get_dim()};
    Field<long> v(dims, 1), w(dims, 1);
    v.init_rand(2); w.init_rand(5);

    double const vmw = v.dot((*mat)(w)).real();
    double const mvw = ((*mat)(v)).dot(w).real();

    if ((vmw-mvw) < 1e-13){
        printf("<v, Mw> = <Mv, w>: Matrix is Hermitian.\n");
    }
    else {
        printf("<v, Mw> != <Mv, w>: Matrix is NOT Hermitian!\n");
    }

    long const dim = (*mat).get_dim();
    for (int i=0; i<dim*dim; i++)
    {
        int const row = i/dim, col = i%dim;
        if(norm((*mat).val_at(row, col) - conj((*mat).val_at(col, row))) > 1e-13) {
            printf("Value-by-value: Matrix is NOT Hermitian!\n");
            break;
        }
    }
    printf("KV_HERMITICITY_DEFECT %.17g\n", vmw - mvw);

    delete mat;
}

// <v, g5 M w> = <g5 M v, w> for the hopping matrix of the sample (gamma5 acts on the spinor axis of the 6-D lattice)
void test_gamma5_hermiticity() {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    auto mat = new Sparse(read_data("4x4parsed.txt"));
    Field<long> v(dims, 6), w(dims, 6);
    v.init_rand(2); w.init_rand(5);
    Field Mw = (*mat)(w), Mv = (*mat)(v);
    std::complex<double> const lhs = v.dot(Mw.gamma5(4));
    std::complex<double> const rhs = (Mv.gamma5(4)).dot(w);
    printf("<v, g5 M w> = (%.12e, %.12e)   <g5 M v, w> = (%.12e, %.12e)\n", lhs.real(), lhs.imag(), rhs.real(), rhs.imag());
    printf("KV_GAMMA5_HERMITICITY_DEFECT %.6e\n", std::abs(lhs - rhs) / std::abs(lhs));
    delete mat;
}

int main() {
    test_MG_property();
    test_hermiticity();
    test_gamma5_hermiticity();
    printf("DONE 1\n");
    return 0;
}
