// include/mgcr/Fields.h -- drop-in for the reference's src/Fields.h: Field<num_type> = an owning array of
// std::complex<double> plus its Mesh (by value), with the same constructors, operators and BLAS-1 members
// (src/Fields.h:29-71).  Here the array lives in HBM: every member is one call into libmgcr_b200.so, element access
// (val_at / mod_val_at) moves a single element over PCIe and is meant for tests and set-up code only.
// Semantics kept from the reference: operators return fresh Fields (src/Fields.h:193-214, 246-253); `dot` conjugates the
// LEFT operand (src/Fields.h:217-226); operator= allocates when empty, copies when sizes match, otherwise prints and
// exit(1)s (src/Fields.h:256-286); init_rand(seed) is the glibc rand() stream with the imaginary part drawn first
// (src/Fields.h:125-135 as g++ compiles it).
#ifndef MGCR_DROPIN_FIELDS_H
#define MGCR_DROPIN_FIELDS_H

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>

#include "Mesh.h"
#include "runtime.h"

template <typename num_type>
class Field {
public:
    Field() = default;
    Field(Field const& f) : mesh(f.mesh) {
        allocate();
        if (f.field) MGCR_CALL(mgcr_vec_copy(mgcr::context(), count(), mgcr::dev(f.field), mgcr::dev(field)));
    }
    Field(Field&& f) noexcept : mesh(f.mesh), field(f.field), borrowed(f.borrowed) { f.field = nullptr; }
    Field(Mesh<num_type> m) : mesh(m) { allocate(); }
    Field(const num_type* dimensions, num_type ndim) : mesh(dimensions, (int)ndim) { allocate(); }   // uninitialised field
    Field(const num_type* dimensions, num_type ndim, std::complex<double>* field_init) : mesh(dimensions, (int)ndim) {
        allocate();
        MGCR_CALL(mgcr_vec_upload(mgcr::context(), mgcr::dev(field), mgcr::dev(field_init), count()));
    }
    // addition: a non-owning view of `n` device elements (used to hand device buffers to caller-defined Operators)
    static Field device_view(std::complex<double>* device_ptr, num_type n) {
        Field f;
        num_type d[1] = {n};
        f.mesh = Mesh<num_type>(d, 1);
        f.field = device_ptr;
        f.borrowed = true;
        return f;
    }

    void init_rand(int seed = 1) {
        if (!field) allocate();
        MGCR_CALL(mgcr_vec_init_rand(mgcr::context(), seed, count(), mgcr::dev(field)));
    }
    void set_zero() { MGCR_CALL(mgcr_vec_set_constant(mgcr::context(), count(), 0., 0., mgcr::dev(field))); }
    void set_constant(std::complex<double> c) { MGCR_CALL(mgcr_vec_set_constant(mgcr::context(), count(), c.real(), c.imag(), mgcr::dev(field))); }

    // Query Field information
    [[nodiscard]] num_type* alloc_get_dim() {
        auto* out = (num_type*)std::malloc(sizeof(num_type) * mesh.get_ndim());
        for (int i = 0; i < mesh.get_ndim(); i++) out[i] = mesh.get_dims()[i];
        return out;
    }
    [[nodiscard]] num_type get_ndim() const { return mesh.get_ndim(); }
    [[nodiscard]] num_type field_size() const { return mesh.get_size(); }
    Mesh<num_type> get_mesh() const { return mesh; }
    std::complex<double> val_at(num_type const* index) {
        for (int i = 0; i < mesh.get_ndim(); i++) assertm(index[i] < mesh.get_dims()[i], "Field memory access out of bound!");
        return fetch(mesh.ind_loc(index));
    }
    std::complex<double> val_at(num_type location) const {
        assertm(location < field_size(), "Field memory access out of bound!");
        return fetch(location);
    }
    void mod_val_at(num_type const* index, std::complex<double> new_value) { store(mesh.ind_loc(index), new_value); }
    void mod_val_at(num_type location, std::complex<double> new_value) { store(location, new_value); }

    // Operations
    Field operator+(const Field& f) const { return combine(f, 1.); }
    Field operator-(const Field& f) const { return combine(f, -1.); }
    [[nodiscard]] std::complex<double> dot(const Field& f) const {   // sum_i conj(this_i) * f_i
        assertm(f.field_size() == field_size(), "Field dimensions do not match!");
        double out[2];
        MGCR_CALL(mgcr_vec_dot(mgcr::context(), count(), mgcr::dev(field), mgcr::dev(f.field), out));
        return {out[0], out[1]};
    }
    [[nodiscard]] double squarednorm() const {
        double out = 0.;
        MGCR_CALL(mgcr_vec_squarednorm(mgcr::context(), count(), mgcr::dev(field), &out));
        return out;
    }
    [[nodiscard]] double norm() const { return std::sqrt(squarednorm()); }
    Field operator*(std::complex<double> a) const {
        Field out(mesh);
        MGCR_CALL(mgcr_vec_scale(mgcr::context(), count(), a.real(), a.imag(), mgcr::dev(field), mgcr::dev(out.field)));
        return out;
    }
    Field& operator=(const Field& f) noexcept {
        if (this == &f) return *this;
        if (!field) {
            mesh = f.mesh;
            allocate();
        } else if (f.field_size() != field_size()) {
            std::printf("Field assignment dimension mismatch: %ld vs %ld\n", (long)field_size(), (long)f.field_size());
            std::exit(1);
        }
        MGCR_CALL(mgcr_vec_copy(mgcr::context(), count(), mgcr::dev(f.field), mgcr::dev(field)));
        return *this;
    }
    Field& operator+=(const Field& f) { return accumulate(f, 1.); }
    Field& operator-=(const Field& f) { return accumulate(f, -1.); }
    void normalise() { MGCR_CALL(mgcr_vec_normalise(mgcr::context(), count(), mgcr::dev(field))); }
    Field gamma5(int spinor_index) const {   // permutation 0<->2, 1<->3 along axis `spinor_index`
        Field out(mesh);
        std::vector<int64_t> d(mesh.get_dims(), mesh.get_dims() + mesh.get_ndim());
        MGCR_CALL(mgcr_vec_gamma5(mgcr::context(), mesh.get_ndim(), d.data(), spinor_index, mgcr::dev(field), mgcr::dev(out.field)));
        return out;
    }

    // additions: bulk host <-> device transfer and the raw device pointer (for code that talks to the C ABI directly)
    void download(std::complex<double>* host) const { MGCR_CALL(mgcr_vec_download(mgcr::context(), mgcr::dev(host), mgcr::dev(field), count())); }
    void upload(const std::complex<double>* host) { MGCR_CALL(mgcr_vec_upload(mgcr::context(), mgcr::dev(field), mgcr::dev(host), count())); }
    std::complex<double>* device_data() { return field; }
    const std::complex<double>* device_data() const { return field; }

    ~Field() {
        if (field && !borrowed) mgcr_vec_free(mgcr::context(), mgcr::dev(field));
    }

protected:
    Mesh<num_type> mesh;
    std::complex<double>* field = nullptr;   // DEVICE pointer
    bool borrowed = false;

private:
    int64_t count() const { return (int64_t)mesh.get_size(); }
    void allocate() {
        mgcr_c128* p = nullptr;
        MGCR_CALL(mgcr_vec_alloc(mgcr::context(), count(), &p));
        field = reinterpret_cast<std::complex<double>*>(p);
        borrowed = false;
    }
    std::complex<double> fetch(num_type loc) const {
        std::complex<double> v;
        MGCR_CALL(mgcr_vec_download(mgcr::context(), mgcr::dev(&v), mgcr::dev(field + loc), 1));
        return v;
    }
    void store(num_type loc, std::complex<double> v) { MGCR_CALL(mgcr_vec_upload(mgcr::context(), mgcr::dev(field + loc), mgcr::dev(&v), 1)); }
    Field combine(const Field& f, double sign) const {
        assertm(f.field_size() == field_size(), "Field dimensions do not match!");
        Field out(mesh);
        MGCR_CALL(mgcr_vec_axpy(mgcr::context(), count(), sign, 0., mgcr::dev(f.field), mgcr::dev(field), mgcr::dev(out.field)));
        return out;
    }
    Field& accumulate(const Field& f, double sign) {
        assertm(f.field_size() == field_size(), "Field dimensions do not match!");
        MGCR_CALL(mgcr_vec_axpy(mgcr::context(), count(), sign, 0., mgcr::dev(f.field), mgcr::dev(field), mgcr::dev(field)));
        return *this;
    }
};

#endif  // MGCR_DROPIN_FIELDS_H
