// placeholder, replaced below
#include "mg.cuh"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { mgcr_set_error(#name ": not built yet"); return MGCR_ERR_UNSUPPORTED; }
STUB(mgcr_mg_create, mgcr_ctx*, mgcr_op*, int, const mgcr_level_cfg*, const mgcr_gcr_param*, const mgcr_gcr_param*, const mgcr_gcr_param*, int, const mgcr_c128*, mgcr_mg**)
STUB(mgcr_mg_destroy, mgcr_mg*)
STUB(mgcr_mg_level_info, mgcr_mg*, int, int64_t*, int64_t*, int*, int64_t*)
STUB(mgcr_mg_export_block_map, mgcr_mg*, int, int64_t*)
STUB(mgcr_mg_export_prolongator, mgcr_mg*, int, mgcr_c128*)
STUB(mgcr_mg_export_coarse, mgcr_mg*, int, int64_t*, int64_t*, mgcr_c128*)
STUB(mgcr_mg_coarse_op, mgcr_mg*, int, mgcr_op**)
STUB(mgcr_mg_restrict, mgcr_ctx*, mgcr_mg*, int, const mgcr_c128*, mgcr_c128*)
STUB(mgcr_mg_prolong, mgcr_ctx*, mgcr_mg*, int, const mgcr_c128*, mgcr_c128*)
STUB(mgcr_mg_cycle, mgcr_ctx*, mgcr_mg*, int, const mgcr_c128*, mgcr_c128*)
STUB(mgcr_mg_op_create, mgcr_ctx*, mgcr_mg*, mgcr_op**)
STUB(mgcr_csr_create_dist, mgcr_ctx*, int64_t, int64_t, int64_t, const int64_t*, const int64_t*, const mgcr_c128*, mgcr_op**)
