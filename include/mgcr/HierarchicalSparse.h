// include/mgcr/HierarchicalSparse.h -- drop-in for the reference's src/HierarchicalSparse.h: block-CSR of dense sub-blocks,
// the storage of the multigrid coarse operator (src/HierarchicalSparse.h:22-48).  The constructor takes the same unordered
// (block, (row, col)) triplets and ADOPTS the block operators (deleted by the destructor, :192-199); the pattern it builds
// is the reference's: triplets sorted row-major by (row, col), duplicates and explicit zero blocks kept, so ROW[r] counts
// every triplet of the rows before r (SURVEY.md 8a row a13, Appendix B Q9/Q10 -- ties are ordered stably here).
// The apply runs on the device: the blocks are uploaded once into the compact column-major compute layout
// (mgcr_blockcsr_create), accumulation order per row = block order, per block = column order, as in :135-147.
#ifndef MGCR_DROPIN_HIERARCHICALSPARSE_H
#define MGCR_DROPIN_HIERARCHICALSPARSE_H

#include <algorithm>
#include <complex>
#include <vector>

#include "Fields.h"
#include "Mesh.h"
#include "Operator.h"

// CRS sparse storage of blocked matrix.  There can be several blocks with the same (row, col) next to each other.
template <typename num_type, typename coarse_num_type>
class HierarchicalSparse : public Operator<num_type> {
public:
    typedef std::pair<Operator<coarse_num_type>*, std::pair<coarse_num_type, coarse_num_type>> Triplet;

    HierarchicalSparse(coarse_num_type block_rows, coarse_num_type block_cols, Triplet* triplets, coarse_num_type triplet_length)
        : nrow(block_rows) {
        sub_dim = triplets[0].first->get_dim();
        this->dim = (num_type)block_cols * sub_dim;
        std::stable_sort(triplets, triplets + triplet_length, [](Triplet const& l, Triplet const& r) { return l.second < r.second; });
        ROW.assign((size_t)block_rows + 1, 0);
        COL.resize((size_t)triplet_length);
        VAL.resize((size_t)triplet_length);
        for (coarse_num_type l = 0; l < triplet_length; l++) {
            VAL[(size_t)l] = triplets[l].first;
            COL[(size_t)l] = triplets[l].second.second;
            ROW[(size_t)triplets[l].second.first + 1]++;
        }
        for (coarse_num_type r = 0; r < block_rows; r++) ROW[(size_t)r + 1] += ROW[(size_t)r];
    }
    HierarchicalSparse(HierarchicalSparse const& m) = delete;   // the reference's copy constructor copies nothing usable (:50-55)

    // Query Sparse matrix information
    [[nodiscard]] num_type get_nrow() const { return (num_type)nrow * sub_dim; }
    [[nodiscard]] num_type get_nnz() const { return (num_type)ROW[(size_t)nrow] * sub_dim * sub_dim; }
    [[nodiscard]] coarse_num_type get_ROW(coarse_num_type r) const { return ROW[(size_t)r]; }   // addition: pattern export
    [[nodiscard]] coarse_num_type get_COL(coarse_num_type l) const { return COL[(size_t)l]; }

    [[nodiscard]] std::complex<double> val_at(num_type row, num_type col) const override {
        const coarse_num_type br = (coarse_num_type)(row / sub_dim), bc = (coarse_num_type)(col / sub_dim);
        std::complex<double> out(0, 0);
        for (coarse_num_type i = ROW[(size_t)br]; i < ROW[(size_t)br + 1]; i++)
            if (COL[(size_t)i] == bc) out += VAL[(size_t)i]->val_at((coarse_num_type)(row - (num_type)br * sub_dim), (coarse_num_type)(col - (num_type)bc * sub_dim));
        return out;
    }
    [[nodiscard]] std::complex<double> val_at(num_type location) const override {
        const num_type bsz = (num_type)sub_dim * sub_dim;
        return VAL[(size_t)(location / bsz)]->val_at((coarse_num_type)(location % bsz));
    }

    Field<num_type> operator()(Field<num_type> const& f) override {
        assertm(f.field_size() == this->dim, "HierarchicalSparse and Field sizes do not match!");
        return this->apply_on_device(f);
    }
    mgcr_op* device_op() override {
        if (!this->handle) {
            const size_t bsz = (size_t)sub_dim * sub_dim;
            std::vector<int64_t> brow(ROW.begin(), ROW.end()), bcol(COL.begin(), COL.end());
            std::vector<std::complex<double>> bval(VAL.size() * bsz);
            for (size_t l = 0; l < VAL.size(); l++)
                for (size_t q = 0; q < bsz; q++) bval[l * bsz + q] = VAL[l]->val_at((coarse_num_type)q);
            MGCR_CALL(mgcr_blockcsr_create(mgcr::context(), (int64_t)nrow, (int)sub_dim, brow.data(), bcol.data(), mgcr::dev(bval.data()), &this->handle));
        }
        return this->handle;
    }

    ~HierarchicalSparse() override {
        for (auto* b : VAL) delete b;
    }

protected:
    std::vector<Operator<coarse_num_type>*> VAL;   // adopted sub-blocks
    std::vector<coarse_num_type> COL;              // column index of each sub-block
    std::vector<coarse_num_type> ROW;              // location where the block row starts
    coarse_num_type nrow = 0;                      // number of block rows
    coarse_num_type sub_dim = 0;
};

#endif  // MGCR_DROPIN_HIERARCHICALSPARSE_H
