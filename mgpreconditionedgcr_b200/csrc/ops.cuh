// csrc/ops.cuh -- operator objects behind the opaque mgcr_op handle (reference: src/Operator.h:16-29).
#pragma once
#include "common.cuh"

enum OpKind { OP_SELL = 1, OP_HOPPING = 2, OP_DIRAC = 3, OP_BLOCKCSR = 4, OP_GCR = 5, OP_MG = 6, OP_CALLBACK = 7 };

// Ghost-exchange plan of a row-slab-partitioned operator: which local elements each peer needs from us (pack list)
// and where what we receive lands in the ghost buffer that the kernels address as column n_local + g.
struct HaloPlan {
    int npeers = 0;
    std::vector<int> peer;            // rank of each peer
    std::vector<int64_t> send_off;    // [npeers+1] offsets into d_send_idx / d_send_buf (elements)
    std::vector<int64_t> recv_off;    // [npeers+1] offsets into the ghost buffer (elements)
    int32_t* d_send_idx = nullptr;    // local indices to pack (NULL when the send range is contiguous)
    std::vector<int64_t> send_start;  // contiguous case: first local element sent to each peer
    c128* d_send_buf = nullptr;
    c128* d_ghost = nullptr;
    int64_t n_ghost = 0;
    int elem = 1;                     // c128 per exchanged item (ne for block operators)
    PeerHalo ph;                      // slab neighbours over peer memory (p2p.cu) when available: replaces the NCCL send/recv pair
    const c128* ghost_cur = nullptr;  // where the ghosts of the latest exchange are (d_ghost, or the peer-memory receive area)
};

struct mgcr_op {
    OpKind kind;
    mgcr_ctx* ctx = nullptr;
    int64_t n_local = 0;    // rows / vector length held by this rank
    int64_t n_global = 0;
    bool distributed = false;   // rows are this rank's slab of a global operator: inner products over its vectors are all-reduced
    virtual ~mgcr_op() {}
    virtual int apply(const c128* x, c128* y) = 0;
    // r = b - A x (the residual of the multigrid cycle); operators with their own kernels fold `b -` into the apply's store
    virtual int apply_residual(const c128* x, const c128* b, c128* r);
    virtual double apply_bytes() const { return 0.; }
};

// Sliced-ELL (slice height 32) image of a CSR matrix: slice s holds rows 32s..32s+31 padded to the longest of them,
// stored column-major inside the slice so that lane l of a warp reads val[base + j*32 + l] -- every access of the
// matrix is a full 128-bit-per-lane coalesced request; entries keep their CSR order (src/Operator.h:336-343).
struct SellOp : mgcr_op {
    int64_t nrow = 0, ncol = 0, nnz = 0, nnz_padded = 0, nslices = 0;
    int64_t* d_slice_ptr = nullptr;
    int32_t* d_col = nullptr;
    c128* d_val = nullptr;
    HaloPlan* halo = nullptr;
    ~SellOp() override;
    int apply(const c128* x, c128* y) override;
    int apply_dirac(const c128* x, c128* y, c128 k, const double* d_diag, const c128* bsub = nullptr);
    double apply_bytes() const override { return (double)nnz_padded * 20. + (double)(nslices + 1) * 8. + 16. * (double)ncol + 16. * (double)nrow; }
};

// Matrix-free nearest-neighbour hopping operator on an (n2, n1, n0) Dirichlet lattice, n0 fastest.
struct HoppingOp : mgcr_op {
    int ndim = 3;
    int64_t gdims[3] = {1, 1, 1};   // global (n2, n1, n0), missing leading dims are 1
    int64_t n2_local = 1, z_begin = 0;
    // variable bond coefficients (var): d_face[2][i] / d_face[1][i] = bond between site i and i+1 / i+n0; d_face[0] holds
    // n2_local+1 planes, plane z = bonds between plane z-1 and plane z (plane 0: to the lower slab neighbour, zero at the
    // global boundary)
    bool var = false;
    double* d_face[3] = {nullptr, nullptr, nullptr};
    c128* d_halo_lo = nullptr; c128* d_halo_hi = nullptr;   // neighbour planes (distributed, NCCL path)
    PeerHalo ph;                                            // neighbour planes over peer memory (p2p.cu) when available
    ~HoppingOp() override;
    int apply(const c128* x, c128* y) override;
    int apply_dirac(const c128* x, c128* y, c128 k, const double* d_diag, const c128* bsub = nullptr);
    int run(const c128* x, c128* y, int dirac, c128 k, const double* d_diag, const c128* bsub = nullptr);
    double apply_bytes() const override { return (var ? 32. + 8. * ndim : 32.) * (double)n_local; }
};

struct DiracOp : mgcr_op {
    mgcr_op* D = nullptr;
    c128 k = {0., 0.};
    double* d_diag = nullptr;
    ~DiracOp() override;
    int apply(const c128* x, c128* y) override;
    int apply_residual(const c128* x, const c128* b, c128* r) override;
    double apply_bytes() const override { return D->apply_bytes() + (d_diag ? 8. * (double)n_local : 0.); }
};

// Block-CSR of dense ne x ne blocks (src/HierarchicalSparse.h:22-48).  Device compute layout: explicit zero blocks
// dropped, blocks stored COLUMN-major so that the ne threads of a block row read contiguous memory for each column.
struct BlockCsrOp : mgcr_op {
    int64_t nb = 0;       // local block rows
    int64_t nb_cols = 0;  // block columns addressable (local + ghost)
    int ne = 0;
    int64_t nnzb = 0;
    int32_t* d_brow = nullptr;   // [nb+1]
    int32_t* d_bcol = nullptr;   // [nnzb]
    c128* d_bval = nullptr;      // [nnzb][ne(col)][ne(row)]
    HaloPlan* halo = nullptr;
    int64_t halo_rows_lo = 0, halo_rows_hi = 0;   // leading / trailing block rows that reference ghost columns
    // Streaming image for the apply (ne = 2, 4, 8; built on the first apply): slices of 32/ne block rows = one warp, every
    // row of a slice padded with zero blocks to the slice's widest.  A slice is one contiguous blob of block slots, slot l =
    // [ne columns][32 lanes] c128 (lane = (row in slice) * ne + r) followed by the 32/ne column indices of the slot, so that a
    // slice is fetched by ONE bulk copy into shared memory (k_blockcsr_ring); the assembly layout above stays for the
    // Galerkin product of the next level, the persistent small-level solver and the exports.
    int64_t nslices = 0, sl_slots = 0;
    int64_t* d_sl_ptr = nullptr;          // [nslices+1] first block slot of each slice
    unsigned char* d_sl_blob = nullptr;   // [sl_slots] slots of ne*512 + (32/ne)*4 bytes
    int sl_stages = 0, sl_stage_bytes = 0;   // ring geometry: one stage (the widest slice) per consumer warp
    int sliced = 0;                       // 0 = not tried yet, 1 = built, -1 = not applicable
    bool halo_deferred = false;           // the latest halo exchange left its wait to the apply kernel (p2p.cu, deferred wait)
    int build_sliced();
    // after the hierarchy is complete nothing reads the assembly copy of a large operator's values any more: give it back
    void drop_assembly_values();
    ~BlockCsrOp() override;
    int apply(const c128* x, c128* y) override;
    int apply_residual(const c128* x, const c128* b, c128* r) override;
    int run(const c128* x, c128* y, const c128* bsub);
    double apply_bytes() const override { return (double)nnzb * (16. * ne * ne + 4.) + 32. * (double)nb * ne; }
};

// An Operator<num_type> subclass implemented by the caller (src/Operator.h:16-29 is an open interface): apply() calls back
// into host code with DEVICE pointers; the callee enqueues its work on the context's stream (or synchronises itself).
struct CallbackOp : mgcr_op {
    mgcr_apply_fn fn = nullptr;
    void* user = nullptr;
    int apply(const c128* x, c128* y) override;
};

int halo_exchange(mgcr_ctx* ctx, HaloPlan* h, const c128* x, bool defer = false);
void halo_free(mgcr_ctx* ctx, HaloPlan* h);
