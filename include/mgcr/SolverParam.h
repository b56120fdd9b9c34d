// include/mgcr/SolverParam.h -- drop-in for the reference's src/SolverParam.h: plain parameter records with the same
// member names, defaults and positional constructors (src/SolverParam.h:10-59), holding NON-owning Operator pointers.
// Additions are trailing members with defaults that leave reference code unchanged:
//   GCR_Param::std_conj / zero_guess         -- SURVEY.md Appendix B, Q3 / Q2
//   MG_Param::neg_neighbour_bug              -- Q8 (replicate src/MG.h:263 for bit-parity runs)
#ifndef MGCR_DROPIN_SOLVERPARAM_H
#define MGCR_DROPIN_SOLVERPARAM_H

#include <cstring>

#include "Operator.h"

template <typename num_type>
class SolverParam {
public:
    Operator<num_type>* left_precond = nullptr;
    Operator<num_type>* right_precond = nullptr;
};

template <typename num_type>
class GCR_Param : public SolverParam<num_type> {
public:
    GCR_Param() = default;
    int truncation = 0;   // set to non-zero for truncation
    int restart = 0;      // set to non-zero for restart
    int max_iter = 100;
    double tol = 1e-16;
    bool verbose = true;
    bool std_conj = false;     // false = the reference's alpha = <r,Ap>/<Ap,Ap> (src/GCR.h:230); true = textbook conjugation
    bool zero_guess = false;   // GCR::operator(): false = start from init_rand(2) (src/GCR.h:63-68); true = start from 0

    GCR_Param(const GCR_Param& param) = default;
    GCR_Param(int trunc, int re, int max_it, double tau, bool verb, Operator<num_type>* solver_l, Operator<num_type>* solver_r)
        : truncation(trunc), restart(re), max_iter(max_it), tol(tau), verbose(verb) {
        this->left_precond = solver_l;
        this->right_precond = solver_r;
    }
    GCR_Param& operator=(const GCR_Param& param) = default;

    mgcr_gcr_param c_param() const {   // the C ABI's view of this record
        mgcr_gcr_param p;
        std::memset(&p, 0, sizeof p);   // padding too: callers compare these records bytewise
        p.truncation = truncation; p.restart = restart; p.max_iter = max_iter; p.tol = tol;
        p.verbose = verbose ? 1 : 0; p.std_conj = std_conj ? 1 : 0; p.zero_guess = zero_guess ? 1 : 0;
        return p;
    }
};

template <typename num_type>
class MG_Param : public SolverParam<num_type> {
public:
    Mesh<num_type> mesh;
    num_type subblock_dim = 0;   // dim of subblocks per spacetime direction
    int n_eigen = 0;             // number of eigenvectors to keep per subblock
    GCR_Param<num_type>* eigenvector_precomp_param = nullptr;
    Operator<num_type>* coarse_solver = nullptr;
    Operator<num_type>* smoother_solver = nullptr;
    bool spacetime[6] = {true, true, true, true, false, false};   // spacetime indices mask
    bool spinor[6] = {false, false, false, false, true, false};
    int n_level = 1;             // numbers of coarse grids
    bool neg_neighbour_bug = false;

    MG_Param() = default;
    MG_Param(const MG_Param& param) = default;
    MG_Param(Mesh<num_type> m, num_type subblock, int eigenvecs, GCR_Param<num_type>* eigen_param, Operator<num_type>* solver_coarse,
             Operator<num_type>* solver_smooth, int levels, Operator<num_type>* solver_l, Operator<num_type>* solver_r)
        : mesh(m), subblock_dim(subblock), n_eigen(eigenvecs), eigenvector_precomp_param(eigen_param), coarse_solver(solver_coarse),
          smoother_solver(solver_smooth), n_level(levels) {
        this->left_precond = solver_l;
        this->right_precond = solver_r;
    }
    MG_Param& operator=(const MG_Param& param) = default;
};

#endif  // MGCR_DROPIN_SOLVERPARAM_H
