// include/mgcr/MG.h -- drop-in for the reference's src/MG.h: MG<num_type>, the adaptive-aggregation multigrid that is itself
// an Operator (a preconditioner for GCR), and Arnoldi<num_type>, the near-null-vector solver (src/MG.h:20-61, 71-87).
// Set-up and apply run entirely on the device behind mgcr_mg_create / mgcr_mg_cycle:
//   initialise(M)  src/MG.h:131-285  near-null vectors by inverse iteration, chirality doubling, aggregation, per-aggregate
//                                    Gram-Schmidt, Galerkin coarse operator -- aggregation and block-CSR pattern bit-exact
//   restrict / expand  :366-383 / :347-364   one pass over the compact prolongator
//   operator()     the cycle of the report's Algorithm 2 (SemesterProject.pdf p.4) in the structure of :405-430
// Side effects on the caller's objects are the reference's: param->mesh is blocked in place (:155) and the caller's
// smoother / coarse solver objects are re-targeted with initialise() (:136, :283).
// Deviations, all because the reference's behaviour is undefined or not a preconditioner:
//   Q5/Q6  solve() takes x by reference and runs Algorithm 2 (the shipped solve() discards its result);
//   Q7     Arnoldi's work vector starts from zero instead of uninitialised memory;
//   Q8     the negative-neighbour coarse block is the correct Galerkin block unless MG_Param::neg_neighbour_bug is set;
//   Q12    meshes need not be 6-D: up to four aggregated dimensions, any spinor/colour extents, chirality doubling only
//          when the spinor extent is 4; MG_Param::n_level > 1 builds deeper levels (K-cycle) -- the reference stores
//          n_level but never reads it.
// The coarse and smoother solvers must be GCR objects (they are in every reference call site); their parameters are read
// at initialise().  test_MG (src/MG.h:432-512) prints the reference's diagnostics from device-resident operations; test_by_value
// and recursive_solve (declared but unused / undefined in the reference) are not provided.
#ifndef MGCR_DROPIN_MG_H
#define MGCR_DROPIN_MG_H

#include <complex>
#include <vector>

#include "Fields.h"
#include "GCR.h"
#include "HierarchicalSparse.h"
#include "Mesh.h"
#include "Operator.h"
#include "SolverParam.h"
#include "utils.h"

// the Galerkin coarse operator built by MG::initialise, seen as an Operator (what the reference holds in m_coarse)
template <typename num_type>
class MGCoarseOperator : public Operator<num_type> {
public:
    MGCoarseOperator(mgcr_mg* hierarchy, int level) : mg(hierarchy), lvl(level) {
        int64_t nb = 0; int ne_ = 0;
        MGCR_CALL(mgcr_mg_level_info(mg, lvl, nullptr, &nb, &ne_, nullptr));
        n_blocks = nb; ne = ne_;
        this->dim = (num_type)(nb * ne_);
        MGCR_CALL(mgcr_mg_coarse_op(mg, lvl, &borrowed));
    }
    Field<num_type> operator()(const Field<num_type>& f) override {
        Field<num_type> out(f.get_mesh());
        MGCR_CALL(mgcr_op_apply(mgcr::context(), borrowed, mgcr::dev(f.device_data()), mgcr::dev(out.device_data())));
        return out;
    }
    mgcr_op* device_op() override { return borrowed; }   // owned by the hierarchy
    // the reference's pattern: 9 blocks per row, explicit zero blocks, sorted by (row, col) (HierarchicalSparse.h:58-98)
    void pattern(std::vector<int64_t>& row, std::vector<int64_t>& col, std::vector<std::complex<double>>& val) const {
        row.resize((size_t)n_blocks + 1); col.resize((size_t)n_blocks * 9); val.resize((size_t)n_blocks * 9 * ne * ne);
        MGCR_CALL(mgcr_mg_export_coarse(mg, lvl, row.data(), col.data(), mgcr::dev(val.data())));
    }
    [[nodiscard]] std::complex<double> val_at(num_type row, num_type col) const override {
        load();
        const int64_t br = row / ne, bc = col / ne;
        std::complex<double> out(0, 0);
        for (int64_t l = prow[(size_t)br]; l < prow[(size_t)br + 1]; l++)
            if (pcol[(size_t)l] == bc) out += pval[(size_t)(l * ne * ne + (row - br * ne) * ne + (col - bc * ne))];
        return out;
    }
    [[nodiscard]] std::complex<double> val_at(num_type location) const override { load(); return pval[(size_t)location]; }

private:
    void load() const { if (prow.empty()) pattern(prow, pcol, pval); }
    mgcr_mg* mg;
    int lvl;
    mgcr_op* borrowed = nullptr;
    int64_t n_blocks = 0;
    int ne = 0;
    mutable std::vector<int64_t> prow, pcol;
    mutable std::vector<std::complex<double>> pval;
};

template <typename num_type>
class MG : public Operator<num_type> {
public:
    MG() = default;
    MG(MG const& mg) = delete;   // the reference's copy constructor aliases the prolongator and double-frees (:293-314)
    MG(Operator<num_type>* M, MG_Param<num_type>* parameter) : param(parameter) { initialise(M); }
    explicit MG(MG_Param<num_type>* parameter) : param(parameter) {}   // must be used in conjunction with initialise()

    void solve(const Field<num_type>& rhs, Field<num_type>& x) {
        MGCR_CALL(mgcr_mg_cycle(mgcr::context(), hierarchy, 0, mgcr::dev(rhs.device_data()), mgcr::dev(x.device_data())));
    }

    // initialisation in case matrix or/and parameter is/are not known at construction time
    void initialise(Operator<num_type>* M) override {
        m = M;
        this->dim = M->get_dim();
        destroy();
        param->smoother_solver->initialise(m);
        auto* smooth = dynamic_cast<GCR<num_type>*>(param->smoother_solver);
        auto* coarse = dynamic_cast<GCR<num_type>*>(param->coarse_solver);
        if (!smooth || !coarse || !param->eigenvector_precomp_param) {
            std::fprintf(stderr, "mgcr: MG needs GCR objects as smoother_solver / coarse_solver and an eigenvector_precomp_param\n");
            std::abort();
        }
        std::printf("Compute global eigenvectors...\n");
        // level configurations from the mesh and its masks
        Mesh<num_type>& mesh = param->mesh;
        const int nd = mesh.get_ndim();
        std::vector<mgcr_level_cfg> cfg((size_t)std::max(1, param->n_level));
        int64_t site[4] = {1, 1, 1, 1};
        int masked = 0, n_spin = 1, n_col = 1;
        bool layout_ok = true, seen_internal = false;
        for (int i = 0; i < nd; i++) {
            if (i < 6 && param->spacetime[i]) {
                layout_ok = layout_ok && !seen_internal;
                masked++;
            } else {
                seen_internal = true;
            }
        }
        layout_ok = layout_ok && masked >= 1 && masked <= 4;
        for (int i = 0, c = 4 - masked; i < nd && layout_ok; i++) {
            const num_type d = mesh.get_dims()[i];
            if (i < 6 && param->spacetime[i]) site[c++] = d;
            else if (i < 6 && param->spinor[i]) { layout_ok = layout_ok && n_col == 1; n_spin *= (int)d; }
            else n_col *= (int)d;
        }
        if (!layout_ok) {
            std::fprintf(stderr, "mgcr: MG expects the aggregated dimensions first (slowest), then spinor, then colour (SURVEY.md 8a row a11)\n");
            std::abort();
        }
        for (size_t l = 0; l < cfg.size(); l++) {
            mgcr_level_cfg& c = cfg[l];
            for (int i = 0; i < 4; i++) {
                c.site_dims[i] = site[i];
                c.sub[i] = site[i] == 1 ? 1 : (int64_t)param->subblock_dim;
                assertm(site[i] % c.sub[i] == 0, "Dimension not exactly divisible by block size!");
                site[i] /= c.sub[i];
            }
            c.n_spin = n_spin; c.n_col = n_col; c.n_eigen = param->n_eigen;
            // next level: mesh {blocks.., chirality, n_eigen} -- coarse index = block*ne + chirality*n_eigen + i (src/MG.h:321-326, 359)
            const bool doubled = (n_spin == 4);
            n_spin = doubled ? 2 : 1;
            n_col = param->n_eigen;
        }
        mgcr_gcr_param pe = param->eigenvector_precomp_param->c_param();
        mgcr_gcr_param pc = coarse->get_param()->c_param(), ps = smooth->get_param()->c_param();
        const int flags = (param->neg_neighbour_bug ? MGCR_MG_NEG_NEIGHBOUR_BUG : 0) | (pc.std_conj ? MGCR_MG_STD_CONJ : 0);
        MGCR_CALL(mgcr_mg_create(mgcr::context(), M->device_op(), (int)cfg.size(), cfg.data(), &pe, &pc, &ps, flags, nullptr, &hierarchy));
        // domain decomposition in spacetime direction, visible to the caller through param->mesh as in the reference
        mesh.blocking(param->subblock_dim, param->spacetime);
        std::printf("Domain decomposition block size (%ld, %ld, %ld, %ld) with block count (%d, %d, %d, %d)\n", (long)cfg[0].sub[0],
                    (long)cfg[0].sub[1], (long)cfg[0].sub[2], (long)cfg[0].sub[3], mesh.get_block_dim()[0], mesh.get_block_dim()[1],
                    mesh.get_block_dim()[2], mesh.get_block_dim()[3]);
        std::printf("Computing coarse matrix... \n");
        m_coarse = new MGCoarseOperator<num_type>(hierarchy, 0);
        param->coarse_solver->initialise(m_coarse);
        std::printf("Adaptive Multigrid precomputation completed.\n");
    }

    // expand/restrict from/to blocked eigenvectors
    Field<num_type> expand(Field<num_type>& x_coarse) {
        Field<num_type> x_fine(param->mesh);
        MGCR_CALL(mgcr_mg_prolong(mgcr::context(), hierarchy, 0, mgcr::dev(x_coarse.device_data()), mgcr::dev(x_fine.device_data())));
        return x_fine;
    }
    Field<num_type> restrict(Field<num_type>& x_fine) {
        num_type dims[1] = {m_coarse->get_dim()};
        Field<num_type> x_coarse(dims, 1);
        MGCR_CALL(mgcr_mg_restrict(mgcr::context(), hierarchy, 0, mgcr::dev(x_fine.device_data()), mgcr::dev(x_coarse.device_data())));
        return x_coarse;
    }
    // copy of x_fine restricted to the sites of one aggregate, zero elsewhere (set-up helper, src/MG.h:385-403)
    Field<num_type> restrict_block(Field<num_type>& x_fine, num_type block_id, num_type sub_size) {
        const num_type n = x_fine.field_size();
        int64_t n_fine = 0, nb = 0, bl = 0; int ne = 0;
        MGCR_CALL(mgcr_mg_level_info(hierarchy, 0, &n_fine, &nb, &ne, &bl));
        const num_type dof = (num_type)(bl / sub_size);
        std::vector<std::complex<double>> in((size_t)n), out((size_t)n, std::complex<double>(0., 0.));
        x_fine.download(in.data());
        const num_type* sites = param->mesh.get_block_map(block_id);
        for (num_type o = 0; o < sub_size; o++)
            for (num_type d = 0; d < dof; d++) out[(size_t)(sites[o] * dof + d)] = in[(size_t)(sites[o] * dof + d)];
        Field<num_type> res(x_fine.get_mesh());
        res.upload(out.data());
        return res;
    }

    // doubling eigenbasis: v+ = (v + g5 v)/2 -> slot i, v- = (v - g5 v)/2 -> slot i + n_eigen (src/MG.h:316-329)
    void vec_double(Field<num_type>* eigenvecs, Field<num_type>* vecs_doubled) {
        for (int i = 0; i < param->n_eigen; i++) {
            Field<num_type> g5 = eigenvecs[i].gamma5(4);
            vecs_doubled[i] = (eigenvecs[i] + g5) * 0.5;
            vecs_doubled[i + param->n_eigen] = (eigenvecs[i] - g5) * 0.5;
        }
    }

    // operator functionality ~M^(-1)
    [[nodiscard]] std::complex<double> val_at(num_type, num_type) const override { std::printf("Warning: Exact value of MG should not be queried!\n"); return 0; }
    [[nodiscard]] std::complex<double> val_at(num_type) const override { std::printf("Warning: Exact value of MG should not be queried!\n"); return 0; }
    Field<num_type> operator()(Field<num_type> const& f) override {
        Field<num_type> x(f.get_mesh());
        solve(f, x);
        return x;
    }
    mgcr_op* device_op() override {
        if (!this->handle) MGCR_CALL(mgcr_mg_op_create(mgcr::context(), hierarchy, &this->handle));
        return this->handle;
    }

    // The reference's diagnostic (src/MG.h:432-512), same prints in the same order.  The reference sums full-lattice copies of
    // the prolongator columns (`large += prolongator[i][n]`); here sum_b prolongator[b][n] = expand(coarse vector that is 1 at
    // b*ne + n for every aggregate b), everything else is the same chain of restrict / expand / operator applies on the device.
    // The numbers are also kept (last_test_*) so that a caller can assert on them.
    void test_MG(Operator<num_type>* M) {
        const int n = 0;   /* test eigenvector 0 */
        int64_t nb = 0, bl = 0; int ne = 0;
        MGCR_CALL(mgcr_mg_level_info(hierarchy, 0, nullptr, &nb, &ne, &bl));
        num_type cdim[1] = {m_coarse->get_dim()};
        auto column_sum = [&](int col, bool all_blocks) {   // sum over aggregates (or aggregate 0 only) of prolongator[b][col]
            std::vector<std::complex<double>> h((size_t)(nb * ne), std::complex<double>(0., 0.));
            for (int64_t b = 0; b < (all_blocks ? nb : 1); b++) h[(size_t)(b * ne + col)] = 1.;
            Field<num_type> sel(cdim, 1);
            sel.upload(h.data());
            return expand(sel);
        };
        Field<num_type> large = column_sum(n, true);
        large.normalise();
        Field<num_type> p01 = column_sum(1 < ne ? 1 : 0, false);
        std::printf("eigen0.dot(eigen1) = %.5e\n", large.dot(p01).real());
        large = (*M)(large);
        std::printf("M (fine) norm = %f\n", large.norm());
        Field<num_type> small = restrict(large);
        Field<num_type> result = expand(small);
        // m_coarse applied to sum over all eigenvectors and expand
        Field<num_type> eigenbasis = column_sum(n, true);
        eigenbasis.normalise();
        Field<num_type> inter = restrict(eigenbasis);
        Field<num_type> inter1 = (*m_coarse)(inter);
        Field<num_type> result1 = expand(inter1);
        std::printf("m (coarse) norm = %f\n", inter1.norm());
        last_test_galerkin = (result - result1).norm() / result.norm();
        std::printf("Relative Difference between TRM and TmR = %.5e\n", last_test_galerkin);
        std::printf("LHS norm = %.5e\n", result.norm());
        std::printf("RHS norm = %.5e\n", result1.norm());
        std::printf("\n\nTest projector:\n");
        std::printf("norm of eigenvector 0 = %f\n", eigenbasis.norm());
        Field<num_type> r = restrict(eigenbasis);
        std::printf("eigenvector 0 restrict:\n");
        Field<num_type> e = expand(r);
        last_test_projector = (e - eigenbasis).norm();
        std::printf("\neigenvector 0 - expand(restrict(eigenvector 0)) = %.5e\n\n", last_test_projector);
    }
    double last_test_galerkin = -1., last_test_projector = -1.;   // additions: what test_MG printed

    // additions: structure export for parity checks
    mgcr_mg* device_hierarchy() const { return hierarchy; }
    Operator<num_type>* coarse_operator() const { return m_coarse; }

    ~MG() override { destroy(); }

private:
    void destroy() {
        this->release_handle();
        delete m_coarse; m_coarse = nullptr;
        if (hierarchy) mgcr_mg_destroy(hierarchy);
        hierarchy = nullptr;
    }
    MG_Param<num_type>* param = nullptr;
    Operator<num_type>* m = nullptr;
    mgcr_mg* hierarchy = nullptr;                      // prolongator + coarse operators of every level, device resident
    MGCoarseOperator<num_type>* m_coarse = nullptr;
};

// near-null vectors by inverse iteration (src/MG.h:71-122)
template <typename num_type>
class Arnoldi {
public:
    Arnoldi(Arnoldi const& ar) = delete;
    Arnoldi(GCR_Param<num_type>* gcr_param, const int n_eigenvec) : n_vec(n_eigenvec), param(gcr_param) {}

    void solve(Operator<num_type>* m_init, Field<num_type>* eigenvecs, Mesh<num_type> mesh) {
        const int64_t n = (int64_t)mesh.get_size();
        mgcr_c128* buf = nullptr;
        MGCR_CALL(mgcr_vec_alloc(mgcr::context(), n * n_vec, &buf));
        mgcr_gcr_param p = param->c_param();
        MGCR_CALL(mgcr_arnoldi(mgcr::context(), m_init->device_op(), &p, n_vec, buf));
        for (int c = 0; c < n_vec; c++) {
            std::printf("Computing smallest eigenvector %d\n", c);
            Field<num_type> v(mesh);
            MGCR_CALL(mgcr_vec_copy(mgcr::context(), n, buf + (size_t)c * n, mgcr::dev(v.device_data())));
            eigenvecs[c] = v;
        }
        MGCR_CALL(mgcr_vec_free(mgcr::context(), buf));
    }

    ~Arnoldi() = default;

private:
    int n_vec;
    GCR_Param<num_type>* param = nullptr;
};

#endif  // MGCR_DROPIN_MG_H
