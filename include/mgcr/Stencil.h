// include/mgcr/Stencil.h -- ADDITION to the reference's operator family (nothing in src/ corresponds to it): the
// nearest-neighbour operators of BASELINE.json's synthetic configurations as a matrix-free Operator<num_type>,
//     y = diag . x - k (H x),      (H x)_i = sum over the in-range +-1 neighbours j of f_ij x_j,
// i.e. what DiracOp(&Sparse(H), k) computes (src/Operator.h:569-574, same ascending-column summation order as
// Sparse::operator(), :336-343) without storing H: 32 B/row of HBM traffic (56 + 8 with bond and diagonal arrays) instead of
// >= 20 B per stored entry.  It is an Operator like any other: GCR, MG and GCR_Param accept it unchanged.
//   Stencil<long> A(dims, 3, k);                       unit hopping, A = 1 - k H  (configs[1]-[3])
//   Stencil<long> A(dims, 3, 1., faces, diag);         real symmetric bonds faces[d][i] (site i -> its +1 neighbour in
//                                                      dim d) and real diagonal: A = diag - H  (configs[4])
#ifndef MGCR_DROPIN_STENCIL_H
#define MGCR_DROPIN_STENCIL_H

#include <vector>

#include "Operator.h"

template <typename num_type>
class Stencil : public Operator<num_type> {
public:
    Stencil(const num_type* dims, int ndim, std::complex<double> k_factor, const double* const* faces = nullptr, const double* diag = nullptr)
        : k(k_factor), dims_(dims, dims + ndim) {
        this->dim = 1;
        for (num_type d : dims_) this->dim *= d;
        if (faces) for (int d = 0; d < ndim; d++) faces_.emplace_back(faces[d], faces[d] + this->dim);
        if (diag) diag_.assign(diag, diag + this->dim);
    }
    // entries of the operator, as the reference's val_at(row, col) reports them (src/Operator.h:111)
    [[nodiscard]] std::complex<double> val_at(num_type row, num_type col) const override {
        if (row == col) return diag_.empty() ? 1. : diag_[row];
        num_type stride = 1;
        for (int d = (int)dims_.size() - 1; d >= 0; d--) {
            const num_type lo = std::min(row, col), hi = std::max(row, col);
            if (hi - lo == stride && (lo / stride) % dims_[d] + 1 < dims_[d]) return -k * (faces_.empty() ? 1. : faces_[d][lo]);
            stride *= dims_[d];
        }
        return {0., 0.};
    }
    [[nodiscard]] std::complex<double> val_at(num_type /*location*/) const override { return {0., 0.}; }   // nothing is stored
    Field<num_type> operator()(Field<num_type> const& f) override {
        assertm(f.field_size() == this->dim, "Stencil and Field sizes do not match!");
        return this->apply_on_device(f);
    }
    mgcr_op* device_op() override {
        if (!this->handle) {
            std::vector<int64_t> d64(dims_.begin(), dims_.end());
            std::vector<const double*> fp;
            for (auto& f : faces_) fp.push_back(f.data());
            MGCR_CALL(mgcr_hopping_create(mgcr::context(), (int)d64.size(), d64.data(), faces_.empty() ? nullptr : fp.data(), &hop));
            MGCR_CALL(mgcr_dirac_create(mgcr::context(), hop, k.real(), k.imag(), diag_.empty() ? nullptr : diag_.data(), &this->handle));
        }
        return this->handle;
    }
    ~Stencil() override {
        this->release_handle();
        if (hop) mgcr_op_destroy(hop);
    }

private:
    std::complex<double> k;
    std::vector<num_type> dims_;
    std::vector<std::vector<double>> faces_;
    std::vector<double> diag_;
    mgcr_op* hop = nullptr;
};

#endif  // MGCR_DROPIN_STENCIL_H
