// include/mgcr/Mesh.h -- drop-in for the reference's src/Mesh.h: N-d index <-> memory-location arithmetic (row-major, first
// index slowest, src/Mesh.h:146-165, 368-398) and the 4-D aggregation `blocking()` (src/Mesh.h:236-298).  Index arithmetic
// is host-side integer code; the aggregation map is produced by the device kernel behind mgcr_blocking_build (bit-exact
// with the reference, tests/test_gpu_krylov.py::test_blocking_bit_exact) and mirrored on the host for get_block_map().
// Pointer-returning alloc_* members hand out `new[]` arrays the caller deletes, exactly like the reference.
#ifndef MGCR_DROPIN_MESH_H
#define MGCR_DROPIN_MESH_H

#include <cassert>
#include <complex>
#include <vector>

#include "runtime.h"

#ifndef assertm
#define assertm(exp, msg) assert(((void)msg, exp))
#endif

template <typename num_type>
class Mesh {
public:
    Mesh() = default;
    Mesh(Mesh const& m) : dim_(m.dim_), size_(m.size_), sub_dim_(m.sub_dim_), block_size_(m.block_size_), map_(m.map_) {
        for (int i = 0; i < 4; i++) { blocked_ind_[i] = m.blocked_ind_[i]; block_dim_[i] = m.block_dim_[i]; }
    }
    Mesh(num_type const* index_dims, int num_dims) : dim_(index_dims, index_dims + num_dims) {
        size_ = 1;
        for (num_type d : dim_) size_ *= d;
    }
    Mesh<num_type>& operator=(Mesh m) noexcept {
        dim_.swap(m.dim_); map_.swap(m.map_); rows_.clear();
        size_ = m.size_; sub_dim_ = m.sub_dim_; block_size_ = m.block_size_;
        for (int i = 0; i < 4; i++) { blocked_ind_[i] = m.blocked_ind_[i]; block_dim_[i] = m.block_dim_[i]; }
        return *this;
    }

    /* index -> memory location */
    static num_type ind_loc(num_type const* index, num_type const* dims, int ndims) {
        num_type loc = index[0];
        for (int i = 1; i < ndims; i++) loc = loc * dims[i] + index[i];
        return loc;
    }
    num_type ind_loc(num_type const* index) const { return ind_loc(index, dim_.data(), (int)dim_.size()); }

    /* memory location -> index (new[]'d, caller deletes) */
    static num_type* alloc_loc_ind(num_type const loc, num_type const* dims, int const ndims) {
        auto* ind = new num_type[ndims];
        num_type rem = loc;
        for (int i = ndims - 1; i >= 0; i--) {
            if (i == 0) { ind[0] = rem; break; }          // the slowest index absorbs any overflow, as in the reference
            ind[i] = rem % dims[i];
            rem /= dims[i];
        }
        return ind;
    }
    [[nodiscard]] num_type* alloc_loc_ind(num_type const loc) const {
        num_type* ind = alloc_loc_ind(loc, dim_.data(), (int)dim_.size());
        assertm(dim_.empty() || ind[0] < dim_[0], "Error in index computation");
        return ind;
    }

    /* 4-D domain decomposition: block_map[b][o] = spacetime location of offset o of block b (src/Mesh.h:236-298) */
    void blocking(num_type subblock_dim, const bool* blocked_dimensions) {
        sub_dim_ = subblock_dim;
        // Generalisation: fewer than four masked dimensions are padded with leading extent-1
        // dimensions, and an extent-1 dimension is "aggregated" with block size 1; with four masked dimensions of
        // extent > 1 this is exactly the reference's decomposition.
        int masked = 0;
        for (size_t i = 0; i < dim_.size(); i++) masked += blocked_dimensions[i] ? 1 : 0;
        assertm(masked >= 1 && masked <= 4, "blocking() decomposes at most four dimensions");
        const int pad = 4 - masked;
        std::vector<int64_t> dims64((size_t)pad, 1);
        std::vector<uint8_t> mask((size_t)pad, 1);
        int64_t sub4[4];
        int count = 0;
        num_type nsite = 1;
        for (; count < pad; count++) { blocked_ind_[count] = -1; block_dim_[count] = 1; sub4[count] = 1; }
        for (size_t i = 0; i < dim_.size(); i++) {
            dims64.push_back((int64_t)dim_[i]);
            mask.push_back(blocked_dimensions[i] ? 1 : 0);
            if (!blocked_dimensions[i]) continue;
            const num_type sub = dim_[i] == 1 ? (num_type)1 : subblock_dim;
            assertm(dim_[i] % sub == 0, "Dimension not exactly divisible by block size!");
            blocked_ind_[count] = (int)i;
            block_dim_[count] = (int)(dim_[i] / sub);
            sub4[count] = (int64_t)sub;
            nsite *= dim_[i];
            count++;
        }
        block_size_ = (num_type)(sub4[0] * sub4[1] * sub4[2] * sub4[3]);
        std::vector<int64_t> flat((size_t)nsite);
        int64_t bd4[4], nb = 0;
        MGCR_CALL(mgcr_blocking_build(mgcr::context(), (int)dims64.size(), dims64.data(), sub4, mask.data(), flat.data(), bd4, &nb));
        map_.assign(flat.begin(), flat.end());
        rows_.clear();
    }
    num_type block_loc(num_type const block_ind[4], num_type const offset[4]) const {
        num_type bd[4], dims[4];
        for (int i = 0; i < 4; i++) { bd[i] = block_dim_[i]; dims[i] = blocked_ind_[i] < 0 ? (num_type)1 : dim_[blocked_ind_[i]]; }
        return get_block_size_() * ind_loc(block_ind, bd, 4) + ind_loc(offset, dims, 4);
    }
    num_type* alloc_block_spacetime(num_type const block_ind[4], num_type const offset[4]) {
        auto* out = new num_type[4];
        for (int i = 0; i < 4; i++) out[i] = sub_dim_ * block_ind[i] + offset[i];
        return out;
    }
    num_type* alloc_loc_block(num_type loc) {
        auto* out = new num_type[4];
        num_type slice = sub_dim_ * sub_dim_ * sub_dim_;
        for (int i = 0; i < 4; i++) {
            out[i] = loc / slice;
            loc -= out[i] * slice;
            slice /= sub_dim_;
        }
        return out;
    }
    num_type* alloc_loc_block_offset(num_type const loc) {
        auto* out = new num_type[4];
        num_type slice = sub_dim_ * sub_dim_ * sub_dim_;
        for (int i = 0; i < 4; i++) {
            out[i] = loc % slice;
            slice /= sub_dim_;
        }
        return out;
    }
    num_type* alloc_spacetime_dim(const bool* blocked_dimensions) {
        auto* out = new num_type[4];
        int count = 0;
        for (size_t i = 0; i < dim_.size() && count < 4; i++)
            if (blocked_dimensions[i]) out[count++] = dim_[i];
        return out;
    }
    [[nodiscard]] num_type get_nblocks() const { return (num_type)block_dim_[0] * block_dim_[1] * block_dim_[2] * block_dim_[3]; }
    int* get_block_dim() { return block_dim_; }
    num_type get_block_size() { return get_block_size_(); }
    num_type* get_block_map(num_type block_idx) {
        if (rows_.empty() && !map_.empty()) {
            const num_type nb = get_nblocks(), bs = get_block_size_();
            rows_.resize((size_t)nb);
            for (num_type b = 0; b < nb; b++) rows_[(size_t)b] = map_.data() + (size_t)(b * bs);
        }
        return rows_[(size_t)block_idx];
    }
    num_type* alloc_full_index(num_type spacetime_loc, int const spinor, int const colour, const bool* spacetime_dimensions,
                               const bool* spinor_dimension) {
        const int nd = (int)dim_.size();
        auto* out = new num_type[nd];
        num_type* st_dims = alloc_spacetime_dim(spacetime_dimensions);
        num_type* st_index = alloc_loc_ind(spacetime_loc, st_dims, 4);
        int count = 0;
        for (int i = 0; i < nd; i++) out[i] = spacetime_dimensions[i] ? st_index[count++] : (spinor_dimension[i] ? (num_type)spinor : (num_type)colour);
        delete[] st_index;
        delete[] st_dims;
        return out;
    }
    num_type* alloc_full_index(num_type spacetime_loc, num_type* /*dim6*/, int const spinor, int const colour,
                               const bool* spacetime_dimensions, const bool* spinor_dimension) {
        return alloc_full_index(spacetime_loc, spinor, colour, spacetime_dimensions, spinor_dimension);
    }

    [[nodiscard]] num_type get_size() const { return size_; }
    [[nodiscard]] int get_ndim() const { return (int)dim_.size(); }
    num_type* get_dims() { return dim_.data(); }
    const num_type* get_dims() const { return dim_.data(); }
    [[nodiscard]] num_type get_sub_dim() const { return sub_dim_; }   // addition: 0 until blocking() has run

    ~Mesh() = default;

private:
    num_type get_block_size_() const { return block_size_; }
    std::vector<num_type> dim_;
    num_type size_ = 0;
    num_type sub_dim_ = 0;
    num_type block_size_ = 0;
    int blocked_ind_[4] = {0, 0, 0, 0};
    int block_dim_[4] = {0, 0, 0, 0};
    std::vector<num_type> map_;      // [n_blocks][block_size], flat
    std::vector<num_type*> rows_;    // row pointers into map_, built on demand
};

#endif  // MGCR_DROPIN_MESH_H
