// include/mgcr/utils.h -- placeholder for the reference's src/utils.h.  Its raw-pointer complex loops (vec_add, vec_amult,
// mat_vec, ...) serve only the Dense algebra and the legacy dense GCR (src/Operator.h:127-179, src/GCR.h:90-139), which are
// outside the solve path; the one helper call sites still use, vec_copy, is kept.
#ifndef MGCR_DROPIN_UTILS_H
#define MGCR_DROPIN_UTILS_H
#include <complex>
#include <cstring>

inline void vec_copy(const std::complex<double>* in, std::complex<double>* out, long n) {
    std::memcpy(out, in, sizeof(std::complex<double>) * (size_t)n);
}
#endif
