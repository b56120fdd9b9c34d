/* include/mgcr_b200.h -- C ABI of the B200-native MG-preconditioned GCR solve path.
 *
 * This is the drop-in boundary: plain pointers, sizes and opaque handles, no C++/torch types.  The C++ host
 * classes in include/mgcr/ (Mesh, Field, Operator, Sparse, DiracOp, HierarchicalSparse, GCR, MG, *_Param --
 * same names and signatures as the reference's headers) are thin wrappers over these entry points, and so are
 * the ctypes bindings the tests and bench.py use.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference repository jing2li/MGPreconditionedGCR).
 *
 * Conventions
 *   - every function returns an int status (MGCR_OK = 0); it never throws, exits or aborts.  The message of the
 *     last failure on the calling thread is available from mgcr_last_error().
 *   - `mgcr_c128` is layout-compatible with std::complex<double> / double _Complex / numpy complex128.
 *   - pointers named d_* are DEVICE pointers (cudaMalloc'd, e.g. from mgcr_vec_alloc or a torch tensor's
 *     data_ptr()); pointers named h_* are HOST pointers.  All pointers are borrowed for the duration of the call
 *     unless the function name says `create` (the returned handle owns device copies of what it was given).
 *   - all work is enqueued on the context's stream; functions that return host-visible numbers synchronise that
 *     stream, the others are asynchronous.  One context = one GPU = one host thread.
 *   - multi-GPU: one process (context) per GPU.  After mgcr_ctx_init_dist every vector is the caller's LOCAL row
 *     slab; inner products are all-reduced and operator applies exchange halos with hand-written kernels over NVLink
 *     peer memory (CUDA IPC; csrc/p2p.cu), NCCL being the bootstrap, the set-up transport and the fallback (MGCR_P2P=0).
 *     Reductions have a GPU-count-independent shape: 1, 2, 4 and 8 GPUs produce identical bits (csrc/common.cuh, RedGeom).
 *   - there is NO CPU fallback: with no usable CUDA device every call fails with MGCR_ERR_CUDA.
 */
#ifndef MGCR_B200_H
#define MGCR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } mgcr_c128;

typedef struct mgcr_ctx mgcr_ctx;   /* device, stream, reduction scratch, NCCL communicator          */
typedef struct mgcr_op mgcr_op;     /* anything that is an Operator<num_type> (src/Operator.h:16-29)  */
typedef struct mgcr_mg mgcr_mg;     /* the multigrid hierarchy built by MG::initialise (src/MG.h:131) */

enum {
    MGCR_OK = 0,
    MGCR_ERR_CUDA = 1,         /* CUDA runtime error / no device                         */
    MGCR_ERR_ARG = 2,          /* invalid argument (the reference's assert sites)        */
    MGCR_ERR_OOM = 3,          /* device or host allocation failed                       */
    MGCR_ERR_NCCL = 4,         /* NCCL missing or failed                                 */
    MGCR_ERR_UNSUPPORTED = 5   /* valid in the reference, not provided by this path      */
};

const char* mgcr_last_error(void);
/* version of the ABI: bumped whenever a signature in this header changes */
int mgcr_abi_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------------------- */
int mgcr_ctx_create(int device, mgcr_ctx** out);
int mgcr_ctx_destroy(mgcr_ctx* ctx);
int mgcr_ctx_sync(mgcr_ctx* ctx);
/* the cudaStream_t all work of this context is enqueued on (for event timing by the caller) */
int mgcr_ctx_stream(mgcr_ctx* ctx, void** stream_out);
/* number of kernels this context has launched so far (bench.py's gpu_launches claim) */
int mgcr_ctx_launch_count(mgcr_ctx* ctx, int64_t* count_out);
/* device-time profile: when enabled every kernel launch is bracketed by pooled CUDA events on the context's stream
 * (no synchronisation until the profile is read), accumulated per kernel class with the algorithmic bytes moved */
int mgcr_ctx_set_profile(mgcr_ctx* ctx, int enabled);
/* copies up to cap entries (names are owned by the context); *n_out = number of kernel classes seen */
int mgcr_ctx_get_profile(mgcr_ctx* ctx, int cap, const char** names, double* ms, int64_t* calls, double* bytes, int* n_out);

/* tunables: "small_gcr_rows" (operators up to this many rows are solved by one persistent kernel, default 2^19; 0 = never),
 * "gather_dofs" (distributed coarse systems up to this size are replicated on every rank, default 2^18),
 * "dot_tma" (1 = TMA-staged batched inner products on long vectors, default 1), "hopping_kernel" (2 = planes staged through
 * shared memory by TMA tensor copies for lattices of at least "hopping_tma_rows" sites (default 2^19), 1 = register-marching
 * stencil kernel, 0 = shared-memory tile kernel), "blockcsr_ring_rows" (block-CSR operators with ne = 2, 4, 8 and at least
 * this many rows are applied from their sliced image streamed through a shared-memory ring of bulk copies, default 2^18), "halo_overlap" (1 = halo exchange on a second stream / communicator while
 * interior rows are computed, default 0: measured no gain) */
int mgcr_ctx_set_option(mgcr_ctx* ctx, const char* key, int64_t value);

/* Multi-GPU (nothing in the reference: it is single-process).  Rank 0 makes an id, the caller broadcasts the
 * 128 bytes by any means (torch.distributed, MPI, a file), every rank then calls init_dist. */
int mgcr_nccl_unique_id(void* h_id128);
int mgcr_ctx_init_dist(mgcr_ctx* ctx, int rank, int nranks, const void* h_id128);
int mgcr_ctx_rank(mgcr_ctx* ctx, int* rank, int* nranks);
/* slab boundaries of every slab-partitioned object created afterwards are multiples of `align` planes of the slowest
 * lattice index (set it to the product of the aggregate sizes of the multigrid levels that stay distributed, so that no
 * aggregate straddles two GPUs; default 1) */
int mgcr_ctx_set_slab_align(mgcr_ctx* ctx, int64_t align);
/* sum-all-reduce of n doubles in device memory (the "scalar dot-product allreduce") */
int mgcr_allreduce_sum(mgcr_ctx* ctx, double* d_buf, int n);
/* contiguous slab of [0, n_slowest) owned by `rank`, aligned to `align` (aggregate size of the slowest dim) */
int mgcr_slab_range(int64_t n_slowest, int64_t align, int rank, int nranks, int64_t* begin, int64_t* end);

/* ---------------------------------------------------------------------------------------------------------
 * Field<num_type> storage and BLAS-1 (src/Fields.h:29-71).  A field is a device array of mgcr_c128.
 * ------------------------------------------------------------------------------------------------------- */
int mgcr_vec_alloc(mgcr_ctx* ctx, int64_t n, mgcr_c128** d_out);                   /* Fields.h:76-104  */
int mgcr_vec_free(mgcr_ctx* ctx, mgcr_c128* d_v);                                  /* Fields.h:341     */
int mgcr_vec_upload(mgcr_ctx* ctx, mgcr_c128* d_dst, const mgcr_c128* h_src, int64_t n);
int mgcr_vec_download(mgcr_ctx* ctx, mgcr_c128* h_dst, const mgcr_c128* d_src, int64_t n);
int mgcr_vec_copy(mgcr_ctx* ctx, int64_t n, const mgcr_c128* d_src, mgcr_c128* d_dst);          /* Fields.h:256-286 operator= */
int mgcr_vec_set_constant(mgcr_ctx* ctx, int64_t n, double re, double im, mgcr_c128* d_v);       /* Fields.h:137-151 set_zero / set_constant */
/* out = a + s*b  (Field operator+ with s=1, operator- with s=-1, += / -= with out == a: Fields.h:192-214,288-308) */
int mgcr_vec_axpy(mgcr_ctx* ctx, int64_t n, double s_re, double s_im, const mgcr_c128* d_b, const mgcr_c128* d_a, mgcr_c128* d_out);
/* out = s * a  (Field operator*(complex): Fields.h:245-253; product formed as s * field[i]) */
int mgcr_vec_scale(mgcr_ctx* ctx, int64_t n, double s_re, double s_im, const mgcr_c128* d_a, mgcr_c128* d_out);
/* sum_i conj(a_i) b_i (Fields.h:216-226).  In a distributed context the vectors are this rank's ROW SLABS: the result is
 * all-reduced over the ranks and the call is COLLECTIVE (every rank must make it). */
int mgcr_vec_dot(mgcr_ctx* ctx, int64_t n, const mgcr_c128* d_a, const mgcr_c128* d_b, double h_out[2]);
/* sum_i |a_i|^2 (Fields.h:228-235); row slabs, all-reduced, collective like mgcr_vec_dot */
int mgcr_vec_squarednorm(mgcr_ctx* ctx, int64_t n, const mgcr_c128* d_a, double* h_out);
/* the same for vectors that are NOT row slabs (replicated coarse-level fields, a rank's own scratch): no all-reduce, not collective */
int mgcr_vec_dot_local(mgcr_ctx* ctx, int64_t n, const mgcr_c128* d_a, const mgcr_c128* d_b, double h_out[2]);
int mgcr_vec_squarednorm_local(mgcr_ctx* ctx, int64_t n, const mgcr_c128* d_a, double* h_out);
/* a *= 1/||a|| (Fields.h:237-243); row slabs, collective like mgcr_vec_dot */
int mgcr_vec_normalise(mgcr_ctx* ctx, int64_t n, mgcr_c128* d_a);
/* permutation 0<->2, 1<->3 along `axis` of a row-major ndim mesh (Fields.h:310-339) */
int mgcr_vec_gamma5(mgcr_ctx* ctx, int ndim, const int64_t* h_dims, int axis, const mgcr_c128* d_in, mgcr_c128* d_out);
/* Field::init_rand(seed) (Fields.h:125-135): the glibc srand/rand stream, drawn on the host (imaginary part first,
 * as g++ evaluates it) and uploaded -- the GPU never re-implements rand(). */
int mgcr_vec_init_rand(mgcr_ctx* ctx, int seed, int64_t n, mgcr_c128* d_out);
/* elements [skip, skip+n) of the same stream: the local slab of a distributed Field::init_rand(seed).  The generator is a
 * linear recurrence, so the slab starts from a jump over its prefix and is drawn by several host threads (csrc/rand.cu) */
int mgcr_vec_init_rand_slab(mgcr_ctx* ctx, int seed, int64_t skip, int64_t n, mgcr_c128* d_out);
/* the same elements into a HOST buffer (no device involved) */
int mgcr_rand_stream(int seed, int64_t skip, int64_t n, mgcr_c128* h_out);

/* Mesh<num_type>::blocking(sub, mask) (src/Mesh.h:236-298), generalised to a per-dimension block size.
 * Exactly 4 dims must be masked.  Writes block_map[b*bs + o] = site (int64, HOST array of prod(masked dims)
 * entries, computed on the device) and block_dim4[4]; *n_blocks_out = number of blocks.  Indivisible dims ->
 * MGCR_ERR_ARG (the assert at Mesh.h:245). */
int mgcr_blocking_build(mgcr_ctx* ctx, int ndim, const int64_t* h_dims, const int64_t* h_sub4, const uint8_t* h_mask,
                        int64_t* h_block_map, int64_t* h_block_dim4, int64_t* n_blocks_out);

/* ---------------------------------------------------------------------------------------------------------
 * Operators (src/Operator.h, src/HierarchicalSparse.h)
 * ------------------------------------------------------------------------------------------------------- */
/* Sparse<num_type>(rows, cols, ROW, COL, VAL) (Operator.h:64): host CSR with int64 indices.  The arrays are
 * copied to the device (into a sliced-ELL layout with int32 columns), the caller keeps ownership of h_*. */
int mgcr_csr_create(mgcr_ctx* ctx, int64_t nrow, int64_t ncol, const int64_t* h_row, const int64_t* h_col,
                    const mgcr_c128* h_val, mgcr_op** out);
/* Distributed variant: this rank passes rows [row_begin, row_end) of a global nrow_global x nrow_global matrix
 * (h_row has row_end-row_begin+1 offsets starting at 0, h_col holds GLOBAL column indices).  Collective. */
int mgcr_csr_create_dist(mgcr_ctx* ctx, int64_t nrow_global, int64_t row_begin, int64_t row_end, const int64_t* h_row,
                         const int64_t* h_col, const mgcr_c128* h_val, mgcr_op** out);
/* Matrix-free nearest-neighbour hopping operator H of an ndim (1..3) Dirichlet lattice, row-major site index
 * (src/Mesh.h:146-154): (H x)_i = sum over the in-range +-1 neighbours j of f_ij x_j -- what a Sparse built through the
 * raw-CSR constructor (src/Operator.h:64) with those entries gives (same ascending-column summation order,
 * src/Operator.h:336-343), without storing the matrix.  h_face == NULL: unit hopping, f = 1.  Otherwise h_face[d]
 * (d = 0..ndim-1, dims order) is a HOST array of n_local doubles: h_face[d][i] = the real, symmetric coefficient of the
 * bond between site i and its +1 neighbour in dim d (entries whose neighbour is outside the lattice are ignored).  In a
 * distributed context dims[0] is split into slabs and the arrays are the local slab; the bond to the upper neighbour's
 * first plane lives with the lower site, the library forwards it once at creation (collective). */
int mgcr_hopping_create(mgcr_ctx* ctx, int ndim, const int64_t* h_dims, const double* const* h_face, mgcr_op** out);
/* the same with the bond arrays already in DEVICE memory (they are copied; the caller keeps ownership) */
int mgcr_hopping_create_dev(mgcr_ctx* ctx, int ndim, const int64_t* h_dims, const double* const* d_face, mgcr_op** out);
/* DiracOp<num_type>(D, k) = 1 - k D (Operator.h:105-122, 556-574).  D is borrowed and must outlive the result.
 * h_diag (optional, HOST array of n doubles) generalises the identity to a real diagonal: y = diag.x - k D x. */
int mgcr_dirac_create(mgcr_ctx* ctx, mgcr_op* D, double k_re, double k_im, const double* h_diag, mgcr_op** out);
/* the same with the diagonal already in DEVICE memory (copied) */
int mgcr_dirac_create_dev(mgcr_ctx* ctx, mgcr_op* D, double k_re, double k_im, const double* d_diag, mgcr_op** out);
int mgcr_dirac_set_k(mgcr_op* dirac, double k_re, double k_im);                      /* Operator.h:118 */
/* HierarchicalSparse<num_type,int>(block_rows, block_cols, triplets, n) (HierarchicalSparse.h:58-98) from an
 * already sorted block-CSR: nb block rows of dense ne x ne row-major blocks (Dense, Operator.h:32-54). */
int mgcr_blockcsr_create(mgcr_ctx* ctx, int64_t nb, int ne, const int64_t* h_brow, const int64_t* h_bcol,
                         const mgcr_c128* h_bval, mgcr_op** out);
/* Any other subclass of Operator<num_type> (Operator.h:16-29 is an open interface): the library calls fn(user, d_x,
 * d_y) with DEVICE pointers of n elements whenever it needs y = A x; fn enqueues its work on the context's stream (or
 * synchronises itself) and returns MGCR_OK.  Lets caller-defined operators be solved / used as preconditioners. */
typedef int (*mgcr_apply_fn)(void* user, const mgcr_c128* d_x, mgcr_c128* d_y);
int mgcr_callback_op_create(mgcr_ctx* ctx, int64_t n, mgcr_apply_fn fn, void* user, mgcr_op** out);
/* Operator::operator()(const Field&) (Operator.h:18): y = A x.  x and y must not alias. */
int mgcr_op_apply(mgcr_ctx* ctx, mgcr_op* op, const mgcr_c128* d_x, mgcr_c128* d_y);
/* r = b - A x in one pass (what `rhs - A(x)` costs three Field temporaries for in the reference; the multigrid cycle's
 * residual).  Operators with their own kernels fold `b -` into the apply's store.  x, b must not alias r. */
int mgcr_op_residual(mgcr_ctx* ctx, mgcr_op* op, const mgcr_c128* d_x, const mgcr_c128* d_b, mgcr_c128* d_r);
int mgcr_op_dim(mgcr_op* op, int64_t* n_local, int64_t* n_global);                   /* Operator.h:20 get_dim  */
/* algorithmic bytes one apply moves (SURVEY 8d formulas for the layout actually traversed) */
int mgcr_op_apply_bytes(mgcr_op* op, double* bytes_out);
int mgcr_op_destroy(mgcr_op* op);

/* ---------------------------------------------------------------------------------------------------------
 * GCR (src/GCR.h, src/SolverParam.h:22-36)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
    int truncation;   /* GCR_Param::truncation (0 = off)                                                   */
    int restart;      /* GCR_Param::restart (0 = off)                                                      */
    int max_iter;     /* GCR_Param::max_iter                                                               */
    double tol;       /* GCR_Param::tol                                                                    */
    int verbose;      /* GCR_Param::verbose: prints "Step %d residual norm = %.10e" per iteration         */
    int std_conj;     /* 0 = the reference's alpha = <r,Ap>/<Ap,Ap> (GCR.h:230); 1 = textbook <Ap,r>/<Ap,Ap> */
    int zero_guess;   /* used by GCR-as-operator: 0 = x0 = init_rand(2) (GCR.h:63-68), 1 = x0 = 0           */
} mgcr_gcr_param;

/* GCR::solve(const Field& rhs, Field& x) (GCR.h:158-302): x += A^-1 rhs (no b - A x0: GCR.h:189).  d_rhs may alias
 * d_x (src/MG.h:102).  `right` (optional) is applied in the flexible form z = R(r).  `left` (optional) is applied exactly as the
 * reference does (GCR.h:201-204, 245-247): r <- L(r) once after p = rhs, Ap = A p were formed from the raw right-hand side, then
 * Ar <- L(A r) in every iteration; what the loop reduces and reports is then L(rhs) - sum alpha Ap.
 * h_hist (optional, hist_cap doubles) receives ||r_g||/||rhs|| for g = 0..iters.  *iters_out = iterations run. */
int mgcr_gcr_solve(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right,
                   const mgcr_c128* d_rhs, mgcr_c128* d_x, double* h_hist, int hist_cap, int* iters_out);
/* The same through HOST buffers: uploads rhs and x0, solves, downloads x (the end-to-end path bench.py times). */
int mgcr_gcr_solve_host(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right,
                        const mgcr_c128* h_rhs, mgcr_c128* h_x, double* h_hist, int hist_cap, int* iters_out);
/* GCR used as an Operator (class GCR : Operator, GCR.h:19; operator() GCR.h:62-68) */
int mgcr_gcr_op_create(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* left, mgcr_op* right, mgcr_op** out);
/* GCR::initialise(Operator*) (GCR.h:31): re-target the solver */
int mgcr_gcr_op_retarget(mgcr_op* gcr, mgcr_op* A);
/* Arnoldi::solve (src/MG.h:90-122): n_vec near-null vectors by inverse iteration, d_vecs = n_vec * n c128 */
int mgcr_arnoldi(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* eigen, int n_vec, mgcr_c128* d_vecs);

/* ---------------------------------------------------------------------------------------------------------
 * MG (src/MG.h, src/SolverParam.h:39-59)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t site_dims[4];  /* the 4 blocked ("spacetime") dims of this level's mesh (Mesh.h:61-62)              */
    int64_t sub[4];        /* aggregate size per dim (MG_Param::subblock_dim, here per dimension)               */
    int n_spin, n_col;     /* dof per site = n_spin*n_col; chirality doubling (MG.h:316-329) iff n_spin == 4    */
    int n_eigen;           /* MG_Param::n_eigen                                                                  */
} mgcr_level_cfg;

enum {
    MGCR_MG_NEG_NEIGHBOUR_BUG = 1,   /* replicate src/MG.h:263 (negative-neighbour block built from prolongator[nb_idx]) */
    MGCR_MG_STD_CONJ = 2             /* inner solvers use the textbook conjugation                                     */
};

/* MG::MG(&param) + MG::initialise(M) (MG.h:23-26, 131-285) for n_level coarse grids.  d_nearnull0 (optional):
 * n_eigen fine-level near-null vectors to use instead of running Arnoldi::solve on level 0. */
int mgcr_mg_create(mgcr_ctx* ctx, mgcr_op* A, int n_level, const mgcr_level_cfg* cfg, const mgcr_gcr_param* eigen,
                   const mgcr_gcr_param* coarse, const mgcr_gcr_param* smooth, int flags, const mgcr_c128* d_nearnull0,
                   mgcr_mg** out);
/* the same with near-null vectors for ANY level: d_nearnull is NULL or an array of n_level device pointers (NULL entries: that
 * level runs Arnoldi::solve); entry l holds cfg[l].n_eigen vectors of the level's (local) length */
int mgcr_mg_create_nn(mgcr_ctx* ctx, mgcr_op* A, int n_level, const mgcr_level_cfg* cfg, const mgcr_gcr_param* eigen,
                      const mgcr_gcr_param* coarse, const mgcr_gcr_param* smooth, int flags, const mgcr_c128* const* d_nearnull,
                      mgcr_mg** out);
int mgcr_mg_destroy(mgcr_mg* mg);
/* wall-clock seconds of the set-up stages (names owned by the hierarchy): "total", "aggregate", "near_null" (Arnoldi::solve,
 * MG.h:142-143; "rand" = its init_rand part), "project_orthonormalise" (MG.h:158-198), "ghost_prolongator", "galerkin"
 * (MG.h:203-281), "coarse_halo_gather", "streaming_image".  *n_out = number of stages. */
int mgcr_mg_setup_profile(mgcr_mg* mg, int cap, const char** names, double* seconds, int* n_out);
/* structure export for parity (all HOST outputs).  level l = 0 .. n_level-1 */
int mgcr_mg_level_info(mgcr_mg* mg, int level, int64_t* n_fine, int64_t* n_blocks, int* ne, int64_t* block_len);
int mgcr_mg_export_block_map(mgcr_mg* mg, int level, int64_t* h_block_map);                 /* Mesh.h:270-293   */
int mgcr_mg_export_prolongator(mgcr_mg* mg, int level, mgcr_c128* h_P);                     /* [nb][ne][bs*dof] */
int mgcr_mg_export_coarse(mgcr_mg* mg, int level, int64_t* h_brow, int64_t* h_bcol, mgcr_c128* h_bval);  /* 9 per row */
int mgcr_mg_coarse_op(mgcr_mg* mg, int level, mgcr_op** borrowed_out);                      /* m_coarse, MG.h:281 */
/* MG::restrict (MG.h:366-383), MG::expand (MG.h:347-364) */
int mgcr_mg_restrict(mgcr_ctx* ctx, mgcr_mg* mg, int level, const mgcr_c128* d_fine, mgcr_c128* d_coarse);
int mgcr_mg_prolong(mgcr_ctx* ctx, mgcr_mg* mg, int level, const mgcr_c128* d_coarse, mgcr_c128* d_fine);
/* one cycle x = MG(b) on `level` (report Algorithm 2; structure of MG.h:405-430) */
int mgcr_mg_cycle(mgcr_ctx* ctx, mgcr_mg* mg, int level, const mgcr_c128* d_b, mgcr_c128* d_x);
/* MG used as an Operator / preconditioner (class MG : Operator, MG.h:21, 124-129) */
int mgcr_mg_op_create(mgcr_ctx* ctx, mgcr_mg* mg, mgcr_op** out);

#ifdef __cplusplus
}
#endif
#endif /* MGCR_B200_H */
