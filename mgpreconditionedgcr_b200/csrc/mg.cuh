// csrc/mg.cuh -- multigrid hierarchy objects behind the opaque mgcr_mg handle (reference: src/MG.h).
#pragma once
#include "ops.cuh"

struct LevelGeom {
    int64_t sd[4];      // site dims of the level (the 4 blocked "spacetime" dims, src/Mesh.h:61-62)
    int64_t sub[4];     // aggregate size per dim
    int64_t bd[4];      // blocks per dim
    int dof;            // dof per site (n_spin * n_col)
    int ne;             // near-null vectors per aggregate (2 n_eigen when chirality-doubled)
    int64_t bs;         // sites per aggregate
    int64_t bl;         // dofs per aggregate = bs * dof
    int64_t nb;         // aggregates
    // slab-partitioned level (all extents above are LOCAL): the slab runs along site dim `pd`, the first dim of extent > 1;
    // planes orthogonal to it are contiguous.  Ghost sites / ghost aggregates = the neighbour ranks' adjacent planes,
    // lower neighbour first.
    int dist, pd, has_lo, has_hi;
    int64_t plane_sites, plane_blocks;   // sites / aggregates per plane
};

struct MgLevel {
    mgcr_level_cfg cfg;
    LevelGeom g;
    mgcr_op* A = nullptr;          // operator of this level (level 0: borrowed from the caller; deeper: previous Ac)
    int64_t nsite = 0, n = 0, nc = 0;
    int K = 0;                     // structural blocks per coarse row
    int64_t* d_block_map = nullptr;   // [nb][bs] -> site   (src/Mesh.h:270-293)
    int32_t* d_site_block = nullptr;  // [nsite]
    int32_t* d_site_off = nullptr;    // [nsite]
    int32_t* d_q_off = nullptr;       // [bl] element offset of in-aggregate position q from the aggregate's first element (restrict / prolong)
    c128* d_P = nullptr;              // compact prolongator [nb][ne][bl]: every fine dof lies in exactly one aggregate
    int8_t* d_bslot = nullptr;        // [nb*K] reference slot (0 self, 2d+1 from the block below in d, 2d+2 from above)
    BlockCsrOp* Ac = nullptr;         // Galerkin coarse operator, owned
    c128 *d_r = nullptr, *d_rc = nullptr, *d_xc = nullptr;   // cycle work vectors
    mgcr_op* deeper = nullptr;        // MG-as-operator of level l+1 (K-cycle preconditioner of the coarse solve)
    // distributed levels
    c128* d_Pg = nullptr;             // prolongator rows of the ghost sites [ghost site][dof][ne]
    bool gather = false;              // the coarse system of this level is replicated on every rank (deeper levels too)
    BlockCsrOp* Ac_full = nullptr;    // replicated coarse operator (gather level only)
    std::vector<int64_t> nc_counts;   // coarse dofs per rank
    int64_t nc_offset = 0, nc_global = 0;
    c128 *d_rc_full = nullptr, *d_xc_full = nullptr, *d_pad = nullptr;
};

struct mgcr_mg {
    mgcr_ctx* ctx = nullptr;
    int n_level = 0;
    std::vector<MgLevel> lv;
    mgcr_gcr_param eigen, coarse, smooth;
    int flags = 0;
    // wall-clock seconds of the set-up stages, summed over the levels (stream synchronised at the stage boundaries)
    std::map<std::string, double> setup_s;
};

// brackets one set-up stage: synchronises the stream on both sides and adds the wall-clock time to mg->setup_s[name]
struct SetupStage {
    mgcr_mg* mg; const char* name; std::chrono::steady_clock::time_point t0;
    SetupStage(mgcr_mg* m, const char* n) : mg(m), name(n) { cudaStreamSynchronize(mg->ctx->stream); t0 = std::chrono::steady_clock::now(); }
    ~SetupStage() {
        cudaStreamSynchronize(mg->ctx->stream);
        mg->setup_s[name] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
};

int gcr_solve(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* prm, mgcr_op* right, const c128* rhs, c128* x, double* hist,
              int hist_cap, int* iters_out, bool x_zero = false);
int arnoldi(mgcr_ctx* ctx, mgcr_op* A, const mgcr_gcr_param* ep, int n_vec, c128* vecs);
int blocking_device(mgcr_ctx* ctx, const int64_t sd[4], const int64_t sub[4], int64_t bd[4], int64_t* d_block_map,
                    int32_t* d_site_block, int32_t* d_site_off);
int vec_gamma5(mgcr_ctx* ctx, int64_t n, int64_t inner, int64_t axis_dim, const c128* in, c128* out);
int vec_axpy(mgcr_ctx* ctx, int64_t n, c128 s, const c128* b, const c128* a, c128* out);
int dist_allgather(mgcr_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank);
