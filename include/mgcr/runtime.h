// include/mgcr/runtime.h -- glue between the drop-in C++ classes (Mesh, Field, Operator, Sparse, DiracOp,
// HierarchicalSparse, GCR, MG, *_Param: same names and signatures as the reference's headers under src/) and the C ABI
// of libmgcr_b200.so (include/mgcr_b200.h).  The reference is single-process and has no notion of a device or context
// (SURVEY.md 2.3), so the classes share one process-wide context: GPU `MGCR_DEVICE` (default 0), created on first use.
//
// Error convention: the reference `assert`s on dimension mismatches with asserts compiled in (src/CMakeLists.txt:6) and
// never returns error codes; the wrappers therefore turn any non-zero status of the C ABI into a message on stderr and
// abort().  There is no CPU path: with no usable B200 the first call aborts with the library's message.
#ifndef MGCR_RUNTIME_H
#define MGCR_RUNTIME_H

#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../mgcr_b200.h"

namespace mgcr {

inline void fail(const char* what, int status) {
    std::fprintf(stderr, "mgcr: %s failed with status %d: %s\n", what, status, mgcr_last_error());
    std::abort();
}

#define MGCR_CALL(expr)                                  \
    do {                                                 \
        int mgcr_st__ = (expr);                          \
        if (mgcr_st__ != MGCR_OK) ::mgcr::fail(#expr, mgcr_st__); \
    } while (0)

struct ContextHolder {
    mgcr_ctx* ctx = nullptr;
    ContextHolder() {
        const char* dev = std::getenv("MGCR_DEVICE");
        MGCR_CALL(mgcr_ctx_create(dev ? std::atoi(dev) : 0, &ctx));
    }
    ~ContextHolder() { /* process exit: device memory is reclaimed by the driver; objects may outlive this holder */ }
};

inline mgcr_ctx* context() {
    static ContextHolder holder;
    return holder.ctx;
}

typedef std::complex<double> cplx;
inline mgcr_c128* dev(cplx* p) { return reinterpret_cast<mgcr_c128*>(p); }
inline const mgcr_c128* dev(const cplx* p) { return reinterpret_cast<const mgcr_c128*>(p); }

}  // namespace mgcr

#endif  // MGCR_RUNTIME_H
