// examples/dropin_check.cpp -- exercises the drop-in C++ classes (include/mgcr/) the way the reference's own ad-hoc tests do
// (src/main.cpp:343-441 matvec, :687-690 DiracOp identity, :877-918 test_MG_property) and prints KEY VALUE lines that
// tests/test_gpu_dropin.py compares with the golden vectors generated from the unmodified reference.
//
//   MGCR_DATA_DIR=<dir with 4x4parsed.txt> ./dropin_check
#include <chrono>
#include <cstdio>
#include <vector>

#include "MG.h"
#include "Parse.h"

// an Operator written by a user of the library: y = 2 x - k D x, through the public interface only
class Shifted : public Operator<long> {
public:
    Shifted(Sparse<long>* d, double kk) : D(d), k(kk) { this->dim = d->get_dim(); }
    Field<long> operator()(const Field<long>& f) override { return f * 2. - (*D)(f) * k; }
    [[nodiscard]] std::complex<double> val_at(long) const override { return 0; }
    [[nodiscard]] std::complex<double> val_at(long, long) const override { return 0; }
private:
    Sparse<long>* D;
    double k;
};

static void print_field(const char* key, const Field<long>& f, long count) {
    for (long i = 0; i < count; i++) printf("%s %ld %.17g %.17g\n", key, i, f.val_at(i).real(), f.val_at(i).imag());
}

int main() {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Mesh mesh(dims, 6);
    long idx[6] = {1, 2, 3, 0, 2, 1};
    printf("IND_LOC %ld\n", mesh.ind_loc(idx));
    auto D = new Sparse(read_data("4x4parsed.txt"));
    printf("NNZ %ld\n", D->get_nnz());
    const double k = 0.05 + 8 * ((0.17865 - 0.05) / 10.);
    auto Dirac = new DiracOp(D, k);

    Field<long> f(dims, 6);
    f.init_rand(1);
    Field<long> g(dims, 6);
    g.init_rand(0);
    print_field("RAND0", g, 2);
    printf("DOT %.17g %.17g\n", f.dot(g).real(), f.dot(g).imag());
    printf("NORM %.17g\n", f.norm());
    Field Df = (*D)(f);
    Field Af = (*Dirac)(f);
    print_field("SPMV", Df, 4);
    print_field("DIRAC", Af, 4);
    printf("DIRAC_IDENTITY %.3e\n", (Af - (f - Df * k)).norm());                 // src/main.cpp:687-690
    printf("VAL_AT %.17g %.17g\n", Dirac->val_at(0, 12).real(), Dirac->val_at(0, 12).imag());
    Field g5 = f.gamma5(4);
    print_field("GAMMA5", g5, 3);

    // unpreconditioned GCR, the parameters of src/main.cpp:857-858, x0 = 0
    GCR_Param<long> p(0, 5, 4000, 1e-13, true, nullptr, nullptr);
    GCR gcr(Dirac, &p);
    Field<long> x(dims, 6);
    x.set_zero();
    gcr.solve(g, x);
    printf("GCR_ITERS %d\n", gcr.iterations());
    printf("GCR_XNORM %.17g\n", x.norm());
    printf("GCR_TRUE_RES %.6e\n", (g - (*Dirac)(x)).norm() / g.norm());
    // GCR as an operator starts from init_rand(2) (src/GCR.h:63-68)
    p.verbose = false;
    Field xr = gcr(g);
    Field<long> r2(dims, 6);
    r2.init_rand(2);
    printf("GCR_OP_RES %.6e\n", (g - (*Dirac)(xr - r2)).norm() / g.norm());

    // a caller-defined Operator goes through the same solver
    Shifted S(D, k);
    GCR_Param<long> ps(0, 5, 400, 1e-12, false, nullptr, nullptr);
    GCR gs(&S, &ps);
    Field<long> xs(dims, 6);
    xs.set_zero();
    gs.solve(g, xs);
    printf("CALLBACK_ITERS %d\n", gs.iterations());
    printf("CALLBACK_RES %.6e\n", (g - S(xs)).norm() / g.norm());

    // multigrid: the flow of test_MG_property (src/main.cpp:877-918) on the 4^4 sample with sub-blocks of 2
    GCR_Param<long> eigen(0, 10, 10, 1e-8, false, nullptr, nullptr);
    GCR_Param<long> coarse(0, 10, 50, 1e-2, false, nullptr, nullptr);
    GCR_Param<long> smooth(0, 10, 0, 1e-8, false, nullptr, nullptr);
    // the textbook conjugation of the GCR coefficients for everything multigrid: with the reference's convention
    // (src/GCR.h:230,258) preconditioned GCR stagnates on this complex operator
    eigen.std_conj = coarse.std_conj = smooth.std_conj = true;
    auto solver_coarse = new GCR(&coarse);
    auto solver_smooth = new GCR(&smooth);
    MG_Param<long> param(mesh, 2, 2, &eigen, solver_coarse, solver_smooth, 1, nullptr, nullptr);
    auto mg = new MG(Dirac, &param);
    printf("NBLOCKS %ld\n", (long)param.mesh.get_nblocks());
    printf("BLOCK_MAP0");
    for (int i = 0; i < 8; i++) printf(" %ld", param.mesh.get_block_map(0)[i]);
    printf("\nBLOCK_MAP1");
    for (int i = 0; i < 8; i++) printf(" %ld", param.mesh.get_block_map(1)[i]);
    printf("\n");
    std::vector<int64_t> brow, bcol;
    std::vector<std::complex<double>> bval;
    static_cast<MGCoarseOperator<long>*>(mg->coarse_operator())->pattern(brow, bcol, bval);
    printf("COARSE_ROW16 %ld\n", (long)brow[16]);
    printf("COARSE_COLS0");
    for (int i = 0; i < 9; i++) printf(" %ld", (long)bcol[i]);
    printf("\n");
    Field<long> rhs(dims, 6);
    rhs.init_rand(42);
    Field inter1 = mg->restrict(rhs);
    Field inter2 = mg->expand(inter1);
    Field inter3 = mg->restrict(inter2);
    Field inter22 = mg->expand(inter3);
    printf("COARSE_DIM %ld\n", inter1.field_size());
    printf("RT_IDENTITY %.5e\n", (inter3 - inter1).norm());
    printf("TR_PROJECTOR %.5e\n", (inter2 - inter22).norm());
    // Galerkin: m_c R v = R M P R v for v in the range of P
    Field lhs = (*mg->coarse_operator())(inter1);
    Field MPv = (*Dirac)(inter2);
    Field rhs_c = mg->restrict(MPv);
    printf("GALERKIN %.5e\n", (lhs - rhs_c).norm() / rhs_c.norm());
    // MG-preconditioned GCR (src/main.cpp:856, the commented-out line, made to work)
    GCR_Param<long> pmg(0, 2, 2000, 1e-13, false, nullptr, mg);
    pmg.std_conj = true;
    GCR gmg(Dirac, &pmg);
    Field<long> y(dims, 6);
    y.set_zero();
    gmg.solve(g, y);
    printf("MG_GCR_ITERS %d\n", gmg.iterations());
    printf("MG_GCR_TRUE_RES %.6e\n", (g - (*Dirac)(y)).norm() / g.norm());

    delete mg;
    delete solver_coarse;
    delete solver_smooth;
    delete Dirac;
    delete D;
    printf("DONE\n");
    return 0;
}
