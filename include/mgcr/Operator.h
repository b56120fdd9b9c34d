// include/mgcr/Operator.h -- drop-in for the reference's src/Operator.h: the abstract Operator<num_type>
// (src/Operator.h:16-29), Dense (as the block type of the coarse operator, :32-54), Sparse CSR (:56-102) and
// DiracOp = 1 - k D (:105-122).  Host-side members keep the reference's meaning (raw-CSR constructor ADOPTS malloc'd arrays
// and frees them, :64, :547-552; mod_*_at / get_* / val_at work on the host arrays); operator() runs on the B200:
// the first apply uploads the CSR into the device layout (sliced-ELL, int32 columns) and later applies reuse it.
//
// Every Operator can hand the library a device handle (`device_op()`); for subclasses written by the caller the
// default implementation wraps their own operator() in a callback operator, so GCR / MG accept any Operator*, like
// the reference does.
//
// Not provided (not on the solve path, only reachable from commented-out tests): Sparse
// operator+ / operator- / operator*(scalar) / dagger(), Dense operator+ / operator* / dagger().
#ifndef MGCR_DROPIN_OPERATOR_H
#define MGCR_DROPIN_OPERATOR_H

#include <algorithm>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "Fields.h"
#include "Mesh.h"
#include "runtime.h"

// kept for source compatibility with code written against the reference (src/Operator.h:11-12); note that any standard
// header that uses the identifier `zero` must be included BEFORE this file
#ifndef MGCR_NO_COMPAT_MACROS
#define one std::complex<double>(1., 0.)
#define zero std::complex<double>(0., 0.)
#endif

// an object that acts on a field
template <typename num_type>
class Operator {
public:
    virtual Field<num_type> operator()(const Field<num_type>&) = 0;

    [[nodiscard]] num_type get_dim() const { return dim; }
    virtual void initialise(Operator* /*op*/) {}
    [[nodiscard]] virtual std::complex<double> val_at(num_type location) const = 0;
    [[nodiscard]] virtual std::complex<double> val_at(num_type row, num_type col) const = 0;
    virtual ~Operator() { release_handle(); }

    // addition: the device-side operator behind this object (built on first use, owned by this object)
    virtual mgcr_op* device_op() {
        if (!handle) {
            MGCR_CALL(mgcr_callback_op_create(mgcr::context(), (int64_t)dim, &Operator::trampoline, this, &handle));
            handle_dim = dim;
        }
        return handle;
    }

protected:
    num_type dim = 0;
    mgcr_op* handle = nullptr;
    num_type handle_dim = 0;
    void release_handle() {
        if (handle) mgcr_op_destroy(handle);
        handle = nullptr;
    }
    // y = (*this)(x) for device buffers: what every library-backed subclass's operator() boils down to
    Field<num_type> apply_on_device(const Field<num_type>& f) {
        Field<num_type> out(f.get_mesh());
        MGCR_CALL(mgcr_op_apply(mgcr::context(), device_op(), mgcr::dev(f.device_data()), mgcr::dev(out.device_data())));
        return out;
    }

private:
    static int trampoline(void* user, const mgcr_c128* d_x, mgcr_c128* d_y) {
        auto* self = static_cast<Operator*>(user);
        Field<num_type> x = Field<num_type>::device_view(reinterpret_cast<std::complex<double>*>(const_cast<mgcr_c128*>(d_x)), self->dim);
        Field<num_type> y = (*self)(x);
        return mgcr_vec_copy(mgcr::context(), (int64_t)self->dim, mgcr::dev(y.device_data()), d_y);
    }
};

// dense dim x dim row-major matrix; on the solve path only as the block type of HierarchicalSparse
template <typename num_type>
class Dense : public Operator<num_type> {
public:
    Dense() = default;
    Dense(Dense const& d) : mat(d.mat) { this->dim = d.dim; }
    Dense(std::complex<double>* matrix, num_type const dimension) : mat(matrix, matrix + (size_t)dimension * dimension) { this->dim = dimension; }

    [[nodiscard]] std::complex<double> val_at(num_type location) const override { return mat[(size_t)location]; }
    [[nodiscard]] std::complex<double> val_at(num_type row, num_type col) const override { return mat[(size_t)(row * this->dim + col)]; }

    Field<num_type> operator()(const Field<num_type>& f) override {   // row-major matvec, sequential per row (src/Operator.h:159-173)
        assertm(this->dim == f.field_size(), "Dense and Field sizes do not match!");
        return this->apply_on_device(f);
    }
    mgcr_op* device_op() override {
        if (!this->handle) {   // one dense block = a 1 x 1 block-CSR
            const int64_t brow[2] = {0, 1}, bcol[1] = {0};
            MGCR_CALL(mgcr_blockcsr_create(mgcr::context(), 1, (int)this->dim, brow, bcol, mgcr::dev(mat.data()), &this->handle));
        }
        return this->handle;
    }
    const std::complex<double>* data() const { return mat.data(); }

    ~Dense() override = default;

private:
    std::vector<std::complex<double>> mat;
};

template <typename num_type>
class Sparse : public Operator<num_type> {
public:
    Sparse() = default;
    explicit Sparse(num_type rows) {   // empty constructor: row offsets only
        ROW = (num_type*)std::malloc(sizeof(num_type) * (rows + 1));
        nrow = rows; this->dim = rows;
    }
    Sparse(num_type rows, num_type cols, num_type nnz) {
        nrow = rows; this->dim = cols;
        ROW = (num_type*)std::malloc(sizeof(num_type) * (rows + 1));
        ROW[rows] = nnz;
        COL = (num_type*)std::malloc(sizeof(num_type) * nnz);
        VAL = (std::complex<double>*)std::malloc(sizeof(std::complex<double>) * nnz);
    }
    Sparse(Sparse const& matrix) { copy_from(matrix); }
    // raw CSR: ADOPTS the malloc'd arrays (freed by the destructor)
    Sparse(num_type rows, num_type cols, num_type* row, num_type* col, std::complex<double>* val) {
        nrow = rows; this->dim = cols; ROW = row; COL = col; VAL = val;
    }
    // Dense -> Sparse
    Sparse(num_type rows, num_type cols, std::complex<double>* matrix) {
        nrow = rows; this->dim = cols;
        num_type nnz = 0;
        for (num_type i = 0; i < rows * cols; i++) nnz += (matrix[i] != 0.) ? 1 : 0;
        ROW = (num_type*)std::malloc(sizeof(num_type) * (rows + 1));
        COL = (num_type*)std::malloc(sizeof(num_type) * std::max<num_type>(nnz, 1));
        VAL = (std::complex<double>*)std::malloc(sizeof(std::complex<double>) * std::max<num_type>(nnz, 1));
        num_type id = 0;
        for (num_type r = 0; r < rows; r++) {
            ROW[r] = id;
            for (num_type c = 0; c < cols; c++)
                if (matrix[r * cols + c] != 0.) { COL[id] = c; VAL[id] = matrix[r * cols + c]; id++; }
        }
        ROW[rows] = id;
    }
    // unordered Triplet -> Sparse (sorted row-major, duplicates summed; rows without entries are allowed)
    Sparse(num_type rows, num_type cols, std::pair<std::complex<double>, std::pair<num_type, num_type>>* triplets, num_type triplet_length) {
        nrow = rows; this->dim = cols;
        std::sort(triplets, triplets + triplet_length, [](auto const& l, auto const& r) { return l.second < r.second; });
        ROW = (num_type*)std::malloc(sizeof(num_type) * (rows + 1));
        COL = (num_type*)std::malloc(sizeof(num_type) * std::max<num_type>(triplet_length, 1));
        VAL = (std::complex<double>*)std::malloc(sizeof(std::complex<double>) * std::max<num_type>(triplet_length, 1));
        num_type nnz = 0, r = 0;
        ROW[0] = 0;
        for (num_type l = 0; l < triplet_length; l++) {
            const num_type tr = triplets[l].second.first, tc = triplets[l].second.second;
            if (nnz > ROW[r] && tr == r && COL[nnz - 1] == tc) { VAL[nnz - 1] += triplets[l].first; continue; }
            while (r < tr) ROW[++r] = nnz;
            COL[nnz] = tc; VAL[nnz] = triplets[l].first; nnz++;
        }
        while (r < rows) ROW[++r] = nnz;
    }

    // Query Sparse matrix information
    [[nodiscard]] num_type get_nrow() const { return nrow; }
    [[nodiscard]] num_type get_nnz() const { return ROW[nrow]; }
    [[nodiscard]] std::complex<double> val_at(num_type row, num_type col) const override {
        for (num_type l = ROW[row]; l < ROW[row + 1]; l++)
            if (COL[l] == col) return VAL[l];
        return {0., 0.};
    }
    [[nodiscard]] std::complex<double> val_at(num_type location) const override { return VAL[location]; }
    [[nodiscard]] num_type get_COL(num_type location) const { return COL[location]; }
    [[nodiscard]] num_type get_ROW(num_type location) const { return ROW[location]; }

    // for initialisation (they invalidate the device image: the next apply uploads again)
    void mod_COL_at(num_type location, num_type val) const { COL[location] = val; dirty = true; }
    void mod_ROW_at(num_type location, num_type val) const { ROW[location] = val; dirty = true; }
    void mod_VAL_at(num_type location, std::complex<double> val) const { VAL[location] = val; dirty = true; }

    // y = A x (src/Operator.h:330-346: per-row sequential accumulation in CSR order)
    Field<num_type> operator()(Field<num_type> const& f) override {
        assertm(f.field_size() == this->dim, "Sparse and Field sizes do not match!");
        Field<num_type> out = this->apply_on_device(f);
        return out;
    }
    Sparse& operator=(const Sparse& mat) noexcept {   // deep copy
        if (this != &mat) { free_host(); copy_from(mat); dirty = true; }
        return *this;
    }
    mgcr_op* device_op() override {
        if (dirty || !this->handle) {
            this->release_handle();
            std::vector<int64_t> r64, c64;
            const int64_t* rp; const int64_t* cp;
            const num_type nnz = ROW[nrow];
            if (sizeof(num_type) == sizeof(int64_t)) {
                rp = reinterpret_cast<const int64_t*>(ROW); cp = reinterpret_cast<const int64_t*>(COL);
            } else {
                r64.assign(ROW, ROW + nrow + 1); c64.assign(COL, COL + nnz);
                rp = r64.data(); cp = c64.data();
            }
            MGCR_CALL(mgcr_csr_create(mgcr::context(), (int64_t)nrow, (int64_t)this->dim, rp, cp, mgcr::dev(VAL), &this->handle));
            dirty = false;
        }
        return this->handle;
    }

    ~Sparse() override { free_host(); }

protected:
    std::complex<double>* VAL = NULL;
    num_type* COL = NULL;   // column index of each value
    num_type* ROW = NULL;   // location where the row starts
    num_type nrow = 0;
    mutable bool dirty = true;

private:
    void free_host() { std::free(VAL); std::free(COL); std::free(ROW); VAL = NULL; COL = NULL; ROW = NULL; }
    void copy_from(const Sparse& m) {
        nrow = m.nrow; this->dim = m.dim;
        const num_type nnz = m.get_nnz();
        ROW = (num_type*)std::malloc(sizeof(num_type) * (nrow + 1));
        COL = (num_type*)std::malloc(sizeof(num_type) * std::max<num_type>(nnz, 1));
        VAL = (std::complex<double>*)std::malloc(sizeof(std::complex<double>) * std::max<num_type>(nnz, 1));
        std::memcpy(ROW, m.ROW, sizeof(num_type) * (nrow + 1));
        std::memcpy(COL, m.COL, sizeof(num_type) * nnz);
        std::memcpy(VAL, m.VAL, sizeof(std::complex<double>) * nnz);
    }
};

// DiracOp = Id - k * D   (D is borrowed, src/Operator.h:117,121); the apply is ONE fused kernel y = x - k (D x)
template <typename num_type>
class DiracOp : public Operator<num_type> {
public:
    DiracOp(Sparse<num_type>* mat, std::complex<double> k_factor) : k(k_factor), D(mat) { this->dim = mat->get_dim(); }
    DiracOp(DiracOp const& op) : k(op.k), D(op.D) { this->dim = op.dim; }

    [[nodiscard]] std::complex<double> val_at(num_type row, num_type col) const override { return 1. - k * D->val_at(row, col); }
    [[nodiscard]] std::complex<double> val_at(num_type location) const override { return 1. - k * D->val_at(location); }

    Field<num_type> operator()(Field<num_type> const& f) override { return this->apply_on_device(f); }

    void set_k(std::complex<double> new_k) {
        k = new_k;
        if (this->handle) MGCR_CALL(mgcr_dirac_set_k(this->handle, k.real(), k.imag()));
    }
    mgcr_op* device_op() override {
        mgcr_op* d = D->device_op();
        if (!this->handle || d != d_seen) {   // (re)bind when D re-uploaded itself
            this->release_handle();
            MGCR_CALL(mgcr_dirac_create(mgcr::context(), d, k.real(), k.imag(), nullptr, &this->handle));
            d_seen = d;
        }
        return this->handle;
    }
    ~DiracOp() override = default;

private:
    std::complex<double> k = 0.;
    Sparse<num_type>* D;
    mgcr_op* d_seen = nullptr;
};

#endif  // MGCR_DROPIN_OPERATOR_H
