// examples/k_critical_mg_precond.cpp -- the reference's live driver flow (src/main.cpp:834-875, `k_critical_mg_precond`)
// written against the drop-in headers in include/mgcr/: the statements below are the reference's own, with the two
// changes its shipped data forces on ANY build of it: the 4^4 lattice / "4x4parsed.txt" (the 8^4 file
// is not shipped) and therefore sub-blocks of 2.  Unlike the reference it also runs the MG-preconditioned solve, which the
// reference leaves commented out because its MG::operator() returns uninitialised memory.
//
//   MGCR_DATA_DIR=<dir with 4x4parsed.txt> ./k_critical_mg_precond
#include <chrono>
#include <cstdio>

#include "MG.h"
#include "Parse.h"

int main() {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Mesh mesh(dims, 6);
    auto D = new Sparse(read_data("4x4parsed.txt"));

    GCR_Param<long> eigen(0, 10, 10, 1e-8, false, nullptr, nullptr);
    GCR_Param<long> coarse(0, 10, 50, 1e-2, false, nullptr, nullptr);
    GCR_Param<long> smooth(0, 10, 0, 1e-8, false, nullptr, nullptr);

    // addition: textbook conjugation of the GCR coefficients inside the multigrid (with the
    // reference's convention preconditioned GCR stagnates on this complex operator)
    eigen.std_conj = coarse.std_conj = smooth.std_conj = true;

    double const st = (0.17865 - 0.05) / 10.;
    for (int exp = 8; exp < 9; exp++) {
        double const k = 0.05 + exp * st;
        printf("k number %d\n", exp);
        auto Dirac = new DiracOp<long>(D, k);

        auto solver_coarse = new GCR(&coarse);
        auto solver_smooth = new GCR(&smooth);
        MG_Param<long> param(mesh, 2, 10, &eigen, solver_coarse, solver_smooth, 1, nullptr, nullptr);
        auto mg = new MG(Dirac, &param);

        GCR_Param<long> gcr_param_mg(0, 2, 2000, 1e-13, true, nullptr, mg);
        gcr_param_mg.std_conj = true;
        GCR_Param<long> gcr_param_new(0, 5, 4000, 1e-13, true, nullptr, nullptr);

        for (int test = 0; test < 1; test++) {
            Field<long> rhs(dims, 6);
            rhs.init_rand(test * 10);
            GCR gcr_plain(Dirac, &gcr_param_new);
            Field x = gcr_plain(rhs);                 // = rand_2 + A^-1 rhs (src/GCR.h:63-68)
            Field<long> x0(dims, 6);
            x0.init_rand(2);
            printf("plain GCR: true residual of (x - x0) = %.6e\n", (rhs - (*Dirac)(x - x0)).norm() / rhs.norm());

            GCR gcr_precond(Dirac, &gcr_param_mg);
            Field<long> y(dims, 6);
            y.set_zero();
            gcr_precond.solve(rhs, y);
            printf("MG-GCR: %d iterations, true residual = %.6e\n", gcr_precond.iterations(), (rhs - (*Dirac)(y)).norm() / rhs.norm());
        }

        delete mg;
        delete Dirac;
        delete solver_coarse;
        delete solver_smooth;
    }
    delete D;
    return 0;
}
