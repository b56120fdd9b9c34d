// include/mgcr/GCR.h -- drop-in for the reference's src/GCR.h: GCR<num_type>, a restarted / truncated / full GCR solver
// that is itself an Operator (so it can be a preconditioner, smoother or coarse solver), same constructors and members
// (src/GCR.h:18-50).  solve() is one call into the device-resident solver (mgcr_gcr_solve): every vector, inner product
// and coefficient stays in HBM; semantics follow src/GCR.h:158-302 exactly: r = rhs (x0 is never
// used to form the residual), conjugated alpha/beta, restart only drops history, max_iter = 0 runs one iteration.
// Deviations: a right preconditioner is applied in the flexible form z = R(r) (Q4; identical when
// R is null); a left preconditioner is applied exactly as the reference does (src/GCR.h:201-204, 245-247: r <- L(r) once,
// Ar <- L(A r) per iteration); the legacy raw-pointer dense solve (src/GCR.h:79-156, only reachable from a commented-out
// test) is not provided.
#ifndef MGCR_DROPIN_GCR_H
#define MGCR_DROPIN_GCR_H

#include <complex>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "Operator.h"
#include "SolverParam.h"
#include "utils.h"

// for solving Ax = rhs
template <typename num_type>
class GCR : public Operator<num_type> {
public:
    GCR() = default;
    GCR(GCR const& gcr) : param(gcr.param) { if (gcr.A_operator) initialise(gcr.A_operator); }
    explicit GCR(Operator<num_type>* M, GCR_Param<num_type>* gcr_param) : A_operator(M), param(gcr_param) { this->dim = A_operator->get_dim(); }
    GCR(GCR_Param<num_type>* gcr_param) : param(gcr_param) {}   // must be used in conjunction with initialise()
    void initialise(Operator<num_type>* M) override {
        A_operator = M;
        this->dim = A_operator->get_dim();
        this->release_handle();
    }

    // solve for Ax = rhs: x <- x + A^-1 rhs
    void solve(const Field<num_type>& rhs, Field<num_type>& x) {
        assertm(rhs.field_size() == this->dim, "Field dimension does not match with Operator!");
        assertm(x.field_size() == this->dim, "x dimension does not match with Operator!");
        assertm(param->truncation == 0 || param->restart == 0, "Do not support concurrent restarting and truncation.");
        mgcr_gcr_param p = param->c_param();
        std::vector<double> hist;
        if (param->verbose) hist.assign((size_t)param->max_iter + 2, 0.);
        int iters = 0;
        MGCR_CALL(mgcr_gcr_solve(mgcr::context(), A_operator->device_op(), &p, device_of(param->left_precond), device_of(param->right_precond),
                                 mgcr::dev(rhs.device_data()), mgcr::dev(x.device_data()), hist.empty() ? nullptr : hist.data(), (int)hist.size(), &iters));
        last_iterations = iters;
        if (param->verbose) {   // the reference's residual-history file (src/GCR.h:168, 270-274)
            const char* path = std::getenv("MGCR_CONVERGENCE_FILE");
            std::ofstream file(path ? path : "../../data/out_data/convergence.txt");
            for (int g = 1; g <= iters && g < (int)hist.size(); g++) file << g << "\t" << hist[(size_t)g] << "\n";
        }
    }

    // Operator functionality, equivalent to applying M^(-1)
    [[nodiscard]] std::complex<double> val_at(num_type row, num_type col) const override { return A_operator->val_at(row, col); }
    [[nodiscard]] std::complex<double> val_at(num_type location) const override { return A_operator->val_at(location); }
    Field<num_type> operator()(Field<num_type> const& f) override {   // x0 = init_rand(2), then solve (src/GCR.h:62-68)
        Field<num_type> x(f.get_mesh());
        if (param->zero_guess) x.set_zero(); else x.init_rand(2);
        solve(f, x);
        return x;
    }
    mgcr_op* device_op() override {
        mgcr_op* a = A_operator->device_op();
        mgcr_op* r = device_of(param->right_precond);
        mgcr_op* l = device_of(param->left_precond);
        mgcr_gcr_param p = param->c_param();
        if (!this->handle || a != a_seen || r != r_seen || l != l_seen || std::memcmp(&p, &p_seen, sizeof p) != 0) {   // parameters may be edited between calls
            this->release_handle();
            p_seen = p;
            MGCR_CALL(mgcr_gcr_op_create(mgcr::context(), a, &p, l, r, &this->handle));
            a_seen = a; r_seen = r; l_seen = l;
        }
        return this->handle;
    }
    GCR_Param<num_type>* get_param() const { return param; }   // addition (MG reads its solvers' parameters)
    int iterations() const { return last_iterations; }         // addition: iteration count of the last solve()

    ~GCR() override = default;

private:
    static mgcr_op* device_of(Operator<num_type>* op) { return op ? op->device_op() : nullptr; }
    Operator<num_type>* A_operator = nullptr;
    GCR_Param<num_type>* param = nullptr;
    mgcr_op *a_seen = nullptr, *r_seen = nullptr, *l_seen = nullptr;
    mgcr_gcr_param p_seen = {};
    int last_iterations = 0;
};

#endif  // MGCR_DROPIN_GCR_H
