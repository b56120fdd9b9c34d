// csrc/mg.cu -- adaptive aggregation multigrid (reference: src/MG.h), device resident.
//   setup  (MG::initialise, MG.h:131-285): blocking -> near-null vectors -> chirality doubling -> block projection ->
//          per-aggregate Gram-Schmidt -> Galerkin coarse blocks, generalised per SURVEY.md Appendix B Q12 to per-dim
//          aggregate sizes, any dof per site and n_level coarse grids.
//   apply  (MG.h:347-383, 405-430): restrict, prolong and the cycle of report Algorithm 2.
// The prolongator is stored compactly ([aggregate][vector][dof in aggregate]): every fine dof lies in exactly one
// aggregate, so the reference's n_blocks*ne full-lattice Fields (MG.h:158-187) collapse to ne numbers per dof.
// The Galerkin product is one pass over the operator's rows: G[B',b] += conj(P_B'[i,:])^T (sum_j M_ij P_b[j,:]) for
// i in B', instead of the reference's n_blocks*9*ne^2 full-lattice matvecs (MG.h:206-278).
#include <math.h>

#include <algorithm>

#include "kernels_blas.cuh"
#include "mg.cuh"
#include "rows.cuh"

// ----------------------------------------------------------------------------------------------------------
// small setup kernels
// ----------------------------------------------------------------------------------------------------------
// v+ = 0.5 (v + g5 v), v- = 0.5 (v - g5 v)                                                      (MG.h:316-329)
static __global__ void __launch_bounds__(RED_THREADS) k_chiral(int64_t n, const c128* __restrict__ v, const c128* __restrict__ g5,
                                                               c128* __restrict__ vp, c128* __restrict__ vm) {
    GRID_STRIDE(i, n) {
        c128 a = v[i], b = g5[i];
        c128 s = cadd(a, b), d = csub(a, b);
        vp[i] = cmake(0.5 * s.x, 0.5 * s.y);
        vm[i] = cmake(0.5 * d.x, 0.5 * d.y);
    }
}

// P[b][e][o*dof+d] = vec[e][block_map[b][o]*dof + d]                                             (MG.h:385-403)
static __global__ void __launch_bounds__(RED_THREADS) k_project(int64_t total, LevelGeom g, int64_t n, const int64_t* __restrict__ block_map,
                                                                const c128* __restrict__ vecs, c128* __restrict__ P) {
    GRID_STRIDE(t, total) {
        int64_t q = t % g.bl;
        int64_t be = t / g.bl;
        int e = (int)(be % g.ne);
        int64_t b = be / g.ne;
        int64_t site = block_map[b * g.bs + q / g.dof];
        P[t] = vecs[(int64_t)e * n + site * g.dof + q % g.dof];
    }
}

// block-level sum of a complex value over the CTA (all threads get the result)
template <int THREADS>
__device__ __forceinline__ c128 cta_sum(c128 v, double* sm /* 2*THREADS/32 + 2 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a = warp_sum(v.x), b = warp_sum(v.y);
    __syncthreads();
    if (lane == 0) { sm[2 * warp] = a; sm[2 * warp + 1] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0., sb = 0.;
        for (int w = 0; w < THREADS / 32; w++) { sa += sm[2 * w]; sb += sm[2 * w + 1]; }
        sm[2 * (THREADS / 32)] = sa; sm[2 * (THREADS / 32) + 1] = sb;
    }
    __syncthreads();
    return cmake(sm[2 * (THREADS / 32)], sm[2 * (THREADS / 32) + 1]);
}

// per-aggregate modified Gram-Schmidt + normalisation, in place                                   (MG.h:189-198)
static __global__ void __launch_bounds__(128) k_block_mgs(LevelGeom g, c128* P) {
    __shared__ double sm[2 * 4 + 2];
    const int64_t b = blockIdx.x;
    c128* Pb = P + b * g.ne * g.bl;
    for (int v = 0; v < g.ne; v++) {
        c128* pv = Pb + (int64_t)v * g.bl;
        for (int j = 0; j < v; j++) {
            const c128* pj = Pb + (int64_t)j * g.bl;
            c128 acc = cmake(0., 0.);
            for (int64_t q = threadIdx.x; q < g.bl; q += 128) { c128 t = cmulc(pj[q], pv[q]); acc.x += t.x; acc.y += t.y; }
            c128 h = cta_sum<128>(acc, sm);
            for (int64_t q = threadIdx.x; q < g.bl; q += 128) pv[q] = csub(pv[q], cmul(h, pj[q]));
            __syncthreads();
        }
        c128 acc = cmake(0., 0.);
        for (int64_t q = threadIdx.x; q < g.bl; q += 128) { c128 t = pv[q]; acc.x += t.x * t.x + t.y * t.y; }
        c128 nn = cta_sum<128>(acc, sm);
        const double s = 1. / sqrt(nn.x);
        for (int64_t q = threadIdx.x; q < g.bl; q += 128) { c128 t = pv[q]; pv[q] = cmake(t.x * s, t.y * s); }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------------------
// Galerkin coarse blocks.  One CTA per coarse block row B.  Reference slots of a row (MG.h:217-276 seen from the row):
//   0        (B, B)
//   2d+1     (B, B-e_d)   the "+d" triplet of column block B-e_d          present iff bd[d] >= 2
//   2d+2     (B, B+e_d)   the "-d" triplet of column block B+e_d          present iff bd[d] >= 3
// (with 2 blocks in a direction both neighbours coincide and the reference keeps the "+" triplet only; neighbours wrap
// periodically as in MG.h:229-231.)  Blocks are written column-major in ascending column order.
// ----------------------------------------------------------------------------------------------------------
struct RowSlots { int K; int slot[9]; int64_t col[9]; };

// On a slab-partitioned level the neighbours across the slab boundary are ghost aggregates (columns nb + ..., lower
// neighbour's plane first); at the global boundary of the partitioned dimension there is no neighbour (the reference's
// periodic wrap-around block, MG.h:229-231, would couple the first and the last GPU; it is structurally zero for the
// Dirichlet operators the distributed configurations use).
// Blocks of a row are ordered by their GLOBAL column: the lower neighbour rank's aggregates come before every local one, the upper
// neighbour's after (slabs are contiguous in the slowest dimension), although their local ghost ids are >= nb.  A slab-partitioned
// coarse operator then adds up the blocks of a row in exactly the order the one-GPU operator does (bit-identical applies).
__host__ __device__ inline void row_slots(const LevelGeom& g, int64_t B, RowSlots* rs) {
    int64_t bi[4], rem = B;
    for (int c = 3; c >= 0; c--) { bi[c] = rem % g.bd[c]; rem /= g.bd[c]; }
    int64_t key[9];   // sort key: global order
    const int64_t LO = -1, HI = (int64_t)1 << 62;
    int K = 0;
    rs->slot[K] = 0; rs->col[K] = B; key[K] = B; K++;
    for (int d = 0; d < 4; d++) {
        int64_t stride = 1;
        for (int c = 3; c > d; c--) stride *= g.bd[c];
        if (g.dist && d == g.pd) {
            const int64_t in_plane = B - bi[d] * stride;   // dims before pd have extent 1
            if (bi[d] > 0) { rs->slot[K] = 2 * d + 1; rs->col[K] = B - stride; key[K] = B - stride; K++; }
            else if (g.has_lo) { rs->slot[K] = 2 * d + 1; rs->col[K] = g.nb + in_plane; key[K] = LO; K++; }
            if (bi[d] + 1 < g.bd[d]) { rs->slot[K] = 2 * d + 2; rs->col[K] = B + stride; key[K] = B + stride; K++; }
            else if (g.has_hi) { rs->slot[K] = 2 * d + 2; rs->col[K] = g.nb + (g.has_lo ? g.plane_blocks : 0) + in_plane; key[K] = HI; K++; }
            continue;
        }
        if (g.bd[d] >= 2) {
            int64_t m = (bi[d] - 1 + g.bd[d]) % g.bd[d];
            rs->slot[K] = 2 * d + 1; rs->col[K] = B + (m - bi[d]) * stride; key[K] = rs->col[K]; K++;
        }
        if (g.bd[d] >= 3) {
            int64_t p = (bi[d] + 1) % g.bd[d];
            rs->slot[K] = 2 * d + 2; rs->col[K] = B + (p - bi[d]) * stride; key[K] = rs->col[K]; K++;
        }
    }
    // ascending global column order, ties (none structurally) by slot: insertion sort
    for (int a = 1; a < K; a++) {
        int s = rs->slot[a]; int64_t c = rs->col[a], k = key[a]; int q = a;
        while (q > 0 && (key[q - 1] > k || (key[q - 1] == k && rs->slot[q - 1] > s))) { rs->slot[q] = rs->slot[q - 1]; rs->col[q] = rs->col[q - 1]; key[q] = key[q - 1]; q--; }
        rs->slot[q] = s; rs->col[q] = c; key[q] = k;
    }
    rs->K = K;
}

// aggregate (as a ghost column id) that owns ghost site gs of a slab-partitioned level
__device__ __forceinline__ int64_t ghost_site_block(const LevelGeom& g, int64_t gs) {
    const int64_t lo_sites = g.has_lo ? g.plane_sites : 0;
    const bool hi = gs >= lo_sites;
    int64_t p = hi ? gs - lo_sites : gs, idx = 0, mul = 1;
    for (int c = 3; c > g.pd; c--) {
        idx += ((p % g.sd[c]) / g.sub[c]) * mul;
        mul *= g.bd[c];
        p /= g.sd[c];
    }
    return g.nb + (hi && g.has_lo ? g.plane_blocks : 0) + idx;
}

template <class Rows>
__global__ void __launch_bounds__(256) k_galerkin(Rows M, LevelGeom g, int KMAX, const int32_t* __restrict__ brow,
                                                  const int64_t* __restrict__ block_map, const int32_t* __restrict__ site_block,
                                                  const int32_t* __restrict__ site_off, const c128* __restrict__ P,
                                                  const c128* __restrict__ Pg, int32_t* __restrict__ bcol_out,
                                                  int8_t* __restrict__ bslot_out, c128* __restrict__ bval_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* G = (c128*)smem_raw;                       // [K][ne*ne] row-major accumulators
    c128* T = G + (int64_t)KMAX * g.ne * g.ne;       // [QC][K][ne]: (M P) rows of the QC fine rows of one pass
    __shared__ RowSlots rs;
    const int64_t B = blockIdx.x;
    const int ne = g.ne;
    const int64_t nsite = g.nb * g.bs;
    const int64_t out0 = brow[B];
    const int QC = max(1, (int)blockDim.x / ne);     // fine rows per pass: thread (qq, c) forms row q0+qq, near-null vector c
    if (threadIdx.x == 0) {
        row_slots(g, B, &rs);
        for (int a = 0; a < rs.K; a++) { bcol_out[out0 + a] = (int32_t)rs.col[a]; bslot_out[out0 + a] = (int8_t)rs.slot[a]; }
    }
    for (int t = threadIdx.x; t < KMAX * ne * ne; t += blockDim.x) G[t] = cmake(0., 0.);
    __syncthreads();
    const int K = rs.K;
    for (int64_t q0 = 0; q0 < g.bl; q0 += QC) {
        const int nq = (int)min((int64_t)QC, g.bl - q0);
        // T[qq][a][c] = sum over the entries (j, v) of fine row i(q0+qq) whose column lies in the block of position a of v * P[j][c]
        {
            const int qq = threadIdx.x / ne, c = threadIdx.x - qq * ne;
            if (qq < nq) {
                const int64_t q = q0 + qq;
                const int64_t i = block_map[B * g.bs + q / g.dof] * g.dof + q % g.dof;
                c128 tacc[9];
#pragma unroll
                for (int a = 0; a < 9; a++) tacc[a] = cmake(0., 0.);
                M.for_each(i, [&](int64_t j, c128 v) {
                    const int64_t js = j / g.dof;
                    int64_t b; c128 pv;
                    if (js < nsite) {
                        b = site_block[js];
                        pv = P[(b * ne + c) * g.bl + (int64_t)site_off[js] * g.dof + (j - js * g.dof)];
                    } else {                                   // ghost site: its prolongator row came from the neighbour rank
                        b = ghost_site_block(g, js - nsite);
                        pv = Pg[(j - nsite * g.dof) * ne + c];
                    }
                    int pos = -1;
#pragma unroll
                    for (int a = 0; a < 9; a++) if (pos < 0 && a < K && rs.col[a] == b) pos = a;
                    if (pos < 0) return;   // not a face neighbour: the reference assembles 9 blocks per row only
#pragma unroll
                    for (int a = 0; a < 9; a++) if (a == pos) tacc[a] = cadd(tacc[a], cmul(v, pv));
                });
#pragma unroll
                for (int a = 0; a < 9; a++) if (a < K) T[((int64_t)qq * K + a) * ne + c] = tacc[a];
            }
        }
        __syncthreads();
        // G[a][r][c] += conj(P_B[q][r]) T[q][a][c], q ascending (the order of the one-row-at-a-time formulation)
        for (int t = threadIdx.x; t < K * ne * ne; t += blockDim.x) {
            const int a = t / (ne * ne), rc = t - a * ne * ne;
            const int r = rc / ne, c = rc - r * ne;
            c128 acc = G[t];
            const c128* pr = P + (B * ne + r) * g.bl + q0;
            for (int qq = 0; qq < nq; qq++) {
                const c128 tv = T[((int64_t)qq * K + a) * ne + c];
                if (tv.x != 0. || tv.y != 0.) acc = cadd(acc, cmulc(pr[qq], tv));
            }
            G[t] = acc;
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < K * ne * ne; t += blockDim.x) {
        const int a = t / (ne * ne), rc = t - a * ne * ne;
        const int r = rc / ne, c = rc - r * ne;
        bval_out[((int64_t)(out0 + a) * ne + c) * ne + r] = G[t];
    }
}

// replicate src/MG.h:263: the (B, B+e_d) block [slot 2d+2, column block b = B+e_d] takes the value of the block the
// reference computed with prolongator[nb_idx], i.e. the "+d" triplet of the same column block: (b+e_d, b) [slot 2d+1]
static __global__ void k_neg_bug(LevelGeom g, const int32_t* __restrict__ brow, const int8_t* __restrict__ bslot, c128* bval) {
    const int64_t B = blockIdx.x;
    const int64_t bsz = (int64_t)g.ne * g.ne;
    for (int l = brow[B]; l < brow[B + 1]; l++) {
        int s = bslot[l];
        if (s == 0 || (s & 1)) continue;
        int d = (s - 2) / 2;
        int64_t bi[4], rem = B;
        for (int c = 3; c >= 0; c--) { bi[c] = rem % g.bd[c]; rem /= g.bd[c]; }
        int64_t stride = 1;
        for (int c = 3; c > d; c--) stride *= g.bd[c];
        int64_t src_row = B + (((bi[d] + 2) % g.bd[d]) - bi[d]) * stride;   // b + e_d = B + 2 e_d
        int sl = -1;
        for (int q = brow[src_row]; q < brow[src_row + 1]; q++) if (bslot[q] == s - 1) sl = q;
        if (sl < 0) continue;
        const c128* src = bval + (int64_t)sl * bsz;
        c128* dst = bval + (int64_t)l * bsz;
        for (int64_t t = threadIdx.x; t < bsz; t += blockDim.x) dst[t] = src[t];
    }
}

// prolongator rows of one plane of sites, packed [site in plane][dof][ne] for the neighbour rank's Galerkin product
static __global__ void __launch_bounds__(RED_THREADS) k_pack_P_plane(LevelGeom g, int64_t first_site, const int32_t* __restrict__ site_block,
                                                                     const int32_t* __restrict__ site_off, const c128* __restrict__ P,
                                                                     c128* __restrict__ buf) {
    const int64_t total = g.plane_sites * g.dof * g.ne;
    GRID_STRIDE(t, total) {
        const int c = (int)(t % g.ne);
        const int64_t sd = t / g.ne;
        const int d = (int)(sd % g.dof);
        const int64_t site = first_site + sd / g.dof;
        buf[t] = P[((int64_t)site_block[site] * g.ne + c) * g.bl + (int64_t)site_off[site] * g.dof + d];
    }
}

// ----------------------------------------------------------------------------------------------------------
// restrict / prolong
// ----------------------------------------------------------------------------------------------------------
// Aggregates are boxes of a structured lattice, so the fine index of (aggregate b, position q inside it) is arithmetic:
// first element of the aggregate (from b's block coordinates, 32-bit: a shard has < 2^31 sites) + a per-level table
// q_off[q] (element offset of in-aggregate position q, bl int32 entries, L1-resident).  Round 1 read the int64 block_map
// (src/Mesh.h:270-293) instead: 8 B per site of extra traffic (+10 % at ne = 4) and a dependent DRAM round trip in front of
// every gather; block_map is now only exported and used by the set-up kernels.
struct AggGeom {
    int bd[4], sub[4], sd[4];
    int dof;
    int linear;   // aggregates are runs of consecutive sites (sub = 1,1,1,s): first element = b * bl
};
__device__ __forceinline__ int64_t agg_first_elem(const AggGeom& a, int64_t bl, int64_t b) {
    if (a.linear) return b * bl;
    unsigned int t = (unsigned int)b;
    const unsigned int b3 = t % (unsigned int)a.bd[3]; t /= (unsigned int)a.bd[3];
    const unsigned int b2 = t % (unsigned int)a.bd[2]; t /= (unsigned int)a.bd[2];
    const unsigned int b1 = t % (unsigned int)a.bd[1]; t /= (unsigned int)a.bd[1];
    const int64_t site = (((int64_t)t * a.sub[0] * a.sd[1] + (int64_t)b1 * a.sub[1]) * a.sd[2] + (int64_t)b2 * a.sub[2]) * a.sd[3] + (int64_t)b3 * a.sub[3];
    return site * a.dof;
}
// q_off[q] for q = o*dof + d: element offset of in-aggregate site o = (o0,o1,o2,o3) row-major over sub
static __global__ void k_agg_offsets(AggGeom a, int bl, int32_t* __restrict__ q_off) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < bl; q += gridDim.x * blockDim.x) {
        int o = q / a.dof;
        const int d = q - o * a.dof;
        const int o3 = o % a.sub[3]; o /= a.sub[3];
        const int o2 = o % a.sub[2]; o /= a.sub[2];
        const int o1 = o % a.sub[1]; o /= a.sub[1];
        const int64_t site = (((int64_t)o * a.sd[1] + o1) * a.sd[2] + o2) * a.sd[3] + o3;
        q_off[q] = (int32_t)(site * a.dof + d);
    }
}

// xc[b*ne+e] = sum_q conj(P[b][e][q]) x[map(b,q)]   (MG.h:366-383).  One CTA per aggregate (grid-stride): the aggregate's
// slice of x is gathered into shared memory once, warp w reduces near-null vectors e = w, w+nw, ... with lanes striding over q.
static __global__ void __launch_bounds__(256) k_restrict(LevelGeom g, AggGeom ag, const int32_t* __restrict__ q_off, const c128* __restrict__ P,
                                                         const c128* __restrict__ xf, c128* __restrict__ xc) {
    PDL_ENTRY();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* xs = (c128*)smem_raw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t b = blockIdx.x; b < g.nb; b += gridDim.x) {
        const c128* xb = xf + agg_first_elem(ag, g.bl, b);
        __syncthreads();                             // the previous aggregate's slice has been consumed
        for (int64_t q = threadIdx.x; q < g.bl; q += blockDim.x) xs[q] = __ldg(xb + __ldg(q_off + q));
        __syncthreads();
        for (int e = warp; e < g.ne; e += nw) {
            const c128* pv = P + (b * g.ne + e) * g.bl;
            double sr = 0., si = 0.;
            for (int64_t q = lane; q < g.bl; q += 32) {
                c128 t = cmulc(ld_stream(pv + q), xs[q]);
                sr += t.x; si += t.y;
            }
            sr = warp_sum(sr); si = warp_sum(si);
            if (lane == 0) xc[b * g.ne + e] = cmake(sr, si);
        }
    }
}

// The same for small aggregates (up to 32*QPL dofs, the synthetic configurations: 4^3 sites x 1..4 dofs): one WARP per
// aggregate, its slice of x held in registers, no shared memory and no block barrier.  Persistent: the warps of a grid
// sized to the machine stride over the aggregates, so that at any time the whole GPU works on one window of consecutive
// aggregates (round 1 launched 262 144 eight-aggregate CTAs per restrict at 512^3).  The per-lane accumulation order is that
// of k_restrict (q = lane, lane+32, ..., then the shuffle tree), so both kernels give bit-identical results.
template <int QPL>
static __global__ void __launch_bounds__(256) k_restrict_warp(LevelGeom g, AggGeom ag, const int32_t* __restrict__ q_off, const c128* __restrict__ P,
                                                              const c128* __restrict__ xf, c128* __restrict__ xc) {
    PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int qo[QPL];
#pragma unroll
    for (int j = 0; j < QPL; j++) qo[j] = lane + 32 * j < g.bl ? __ldg(q_off + lane + 32 * j) : 0;
    for (int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; b < g.nb; b += nwarps) {
        const c128* xb = xf + agg_first_elem(ag, g.bl, b);
        c128 xs[QPL];
#pragma unroll
        for (int j = 0; j < QPL; j++) xs[j] = lane + 32 * j < g.bl ? __ldg(xb + qo[j]) : cmake(0., 0.);
#pragma unroll 4
        for (int e = 0; e < g.ne; e++) {
            const c128* pv = P + (b * g.ne + e) * g.bl;
            double sr = 0., si = 0.;
#pragma unroll
            for (int j = 0; j < QPL; j++) {
                const int64_t q = lane + 32 * j;
                if (q < g.bl) {
                    c128 t = cmulc(ld_stream(pv + q), xs[j]);
                    sr += t.x; si += t.y;
                }
            }
            sr = warp_sum(sr); si = warp_sum(si);
            if (lane == 0) xc[b * g.ne + e] = cmake(sr, si);
        }
    }
}

// The same kernel for ne = 4 with the eight warp reductions (4 vectors x re / im) folded into one halving butterfly: at the
// xor-16 step a lane keeps four of its eight sums and hands the other four to its partner, at xor-8 two of four, at xor-4 one of
// two, then two ordinary steps -- 9 double shuffles instead of 40.  Every value still goes through the tree xor 16, 8, 4, 2, 1 with
// the same pairs (a + b and b + a are the same bits), so the results are those of k_restrict_warp bit for bit; lanes 0, 4, .., 28
// end up holding value (lane >> 2) in the bit order 4, 2, 1 -> lane bits 4, 3, 2.  The reductions were a third of the warp's issue
// slots: in the solve (SM clocks under the power cap) restrict ran 8 % below its standalone rate.
template <int QPL>
static __global__ void __launch_bounds__(256) k_restrict_warp4(LevelGeom g, AggGeom ag, const int32_t* __restrict__ q_off, const c128* __restrict__ P,
                                                               const c128* __restrict__ xf, c128* __restrict__ xc) {
    PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    const int vidx = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
    int qo[QPL];
#pragma unroll
    for (int j = 0; j < QPL; j++) qo[j] = lane + 32 * j < g.bl ? __ldg(q_off + lane + 32 * j) : 0;
    for (int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; b < g.nb; b += nwarps) {
        const c128* xb = xf + agg_first_elem(ag, g.bl, b);
        // dof chunk outermost (per vector the terms are still added in the order j = 0, 1, ...): the x element and its four
        // prolongator elements are one batch of five independent loads, and the batches of different j overlap
        // (long aggregates only: with QPL = 2 everything is in flight at once either way and the vector-outermost form needs 48
        // registers instead of 64 -- 1.63 against 1.88 ms for the finest-level restrict of the 512^3 solve)
        const c128* pv = P + b * 4 * g.bl;
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = 0.;
        if constexpr (QPL <= 2) {
            c128 xs[QPL];
#pragma unroll
            for (int j = 0; j < QPL; j++) xs[j] = lane + 32 * j < g.bl ? __ldg(xb + qo[j]) : cmake(0., 0.);
#pragma unroll
            for (int e = 0; e < 4; e++) {
#pragma unroll
                for (int j = 0; j < QPL; j++) {
                    const int64_t q = lane + 32 * j;
                    if (q < g.bl) {
                        c128 t = cmulc(ld_stream(pv + e * g.bl + q), xs[j]);
                        v[2 * e] += t.x; v[2 * e + 1] += t.y;
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < QPL; j++) {
                const int64_t q = lane + 32 * j;
                if (q < g.bl) {
                    const c128 xv = __ldg(xb + qo[j]);
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        c128 t = cmulc(ld_stream(pv + e * g.bl + q), xv);
                        v[2 * e] += t.x; v[2 * e + 1] += t.y;
                    }
                }
            }
        }
        double w[4], u[2];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double keep = b4 ? v[k + 4] : v[k], send = b4 ? v[k] : v[k + 4];
            w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double keep = b3 ? w[k + 2] : w[k], send = b3 ? w[k] : w[k + 2];
            u[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        double t = (b2 ? u[1] : u[0]) + __shfl_xor_sync(0xffffffffu, b2 ? u[0] : u[1], 4);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        if ((lane & 3) == 0) ((double*)(xc + b * 4))[vidx] = t;
    }
}

// Tiny aggregates (up to G = 8 or 16 dofs, e.g. the 1x1x8 line aggregates of the anisotropic configuration): G lanes per
// aggregate, 32/G aggregates per warp, so that no lane idles.  Lane q of the group holds dof q; the group's shuffle tree
// (xor G/2 .. 1) is the tail of the full warp tree, whose upper steps would only add zeros: bit-identical to k_restrict_warp.
template <int G>
static __global__ void __launch_bounds__(256) k_restrict_sub(LevelGeom g, AggGeom ag, const int32_t* __restrict__ q_off, const c128* __restrict__ P,
                                                             const c128* __restrict__ xf, c128* __restrict__ xc) {
    PDL_ENTRY();
    const int q = (int)(threadIdx.x % G);
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    const int qo = q < g.bl ? __ldg(q_off + q) : 0;
    const int64_t nb_round = (g.nb + (32 / G) - 1) / (32 / G) * (32 / G);   // whole warps run the shuffles together
    for (int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G; b < nb_round; b += T / G) {
        const bool on = b < g.nb && q < g.bl;
        c128 xv = cmake(0., 0.);
        if (on) xv = __ldg(xf + agg_first_elem(ag, g.bl, b) + qo);
        for (int e = 0; e < g.ne; e++) {
            double sr = 0., si = 0.;
            if (on) {
                c128 v = cmulc(ld_stream(P + (b * g.ne + e) * g.bl + q), xv);
                sr = v.x; si = v.y;
            }
#pragma unroll
            for (int o = G / 2; o >= 1; o >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, o);
                si += __shfl_xor_sync(0xffffffffu, si, o);
            }
            if (on && q == 0) xc[b * g.ne + e] = cmake(sr, si);
        }
    }
}

// x[map(b,q)] = sum_e xc[b*ne+e] P[b][e][q]   (MG.h:347-364), e in the reference's order.  One warp per aggregate
// (persistent, like k_restrict_warp): the ne coarse coefficients are warp-uniform, lane l forms the dofs q = l, l+32, ...
static __global__ void __launch_bounds__(256) k_prolong(LevelGeom g, AggGeom ag, const int32_t* __restrict__ q_off, const c128* __restrict__ P,
                                                        const c128* __restrict__ xc, c128* __restrict__ xf, int add) {
    PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; b < g.nb; b += nwarps) {
        c128* xb = xf + agg_first_elem(ag, g.bl, b);
        const c128* pb = P + b * g.ne * g.bl;
        const c128* a = xc + b * g.ne;
        for (int64_t q0 = 0; q0 < g.bl; q0 += 64) {          // two dofs per lane and trip: 2*ne independent loads in flight
            const int64_t qa = q0 + lane, qb = q0 + 32 + lane;
            const bool ona = qa < g.bl, onb = qb < g.bl;
            c128 acca = cmake(0., 0.), accb = cmake(0., 0.);
            c128* da = ona ? xb + __ldg(q_off + qa) : nullptr;
            c128* db = onb ? xb + __ldg(q_off + qb) : nullptr;
            c128 olda = cmake(0., 0.), oldb = cmake(0., 0.);
            if (add) { if (ona) olda = *da; if (onb) oldb = *db; }
#pragma unroll 4
            for (int e = 0; e < g.ne; e++) {
                const c128 ce = __ldg(a + e);
                if (ona) acca = cadd(acca, cmul(ce, ld_stream(pb + (int64_t)e * g.bl + qa)));
                if (onb) accb = cadd(accb, cmul(ce, ld_stream(pb + (int64_t)e * g.bl + qb)));
            }
            if (ona) *da = add ? cadd(olda, acca) : acca;      // add: the cycle's x += P xc
            if (onb) *db = add ? cadd(oldb, accb) : accb;
        }
    }
}

// Short aggregate rows (RL = sub[3]*dof contiguous fine elements; 4 = 64 bytes for the 4^3 scalar aggregates of the finest level):
// a warp works on AW aggregates that are neighbours along x, so that each of its fine-lattice requests is a run of RL*AW elements
// (128 bytes) instead of RL.  Lane l: element l % RL of aggregate (l / RL) % AW, row l / (RL*AW) of the trip's rows.  The sum over e
// is the same per element as in k_prolong: identical bits.  Measured standalone at 512^3 (scripts/kbench_transfer.cu,
// profiles/r02_kbench_transfer_rows.txt): 6.6 TB/s against 6.1 for one aggregate per warp.
template <int RL, int AW, int NE /* compile-time number of near-null vectors, 0 = g.ne */>
static __global__ void __launch_bounds__(256) k_prolong_rows(LevelGeom g, AggGeom ag, const int32_t* __restrict__ q_off, const c128* __restrict__ P,
                                                             const c128* __restrict__ xc, c128* __restrict__ xf, int add) {
    PDL_ENTRY();
    constexpr int RPT = 32 / (RL * AW);               // rows per trip
    const int lane = threadIdx.x & 31;
    const int off = lane % RL, a_in = (lane / RL) % AW, r0 = lane / (RL * AW);
    const int rows = (int)(g.bl / RL);
    const int64_t ngroups = g.nb / AW;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; grp < ngroups; grp += nwarps) {
        const int64_t b = grp * AW + a_in;
        const int ne = NE ? NE : g.ne;
        c128* xb = xf + agg_first_elem(ag, g.bl, b);
        const c128* pb = P + b * ne * g.bl;
        const c128* a = xc + b * ne;
        for (int row = r0; row < rows; row += 2 * RPT) {    // two rows per lane and trip: 2*ne independent loads in flight
            const int qa = row * RL + off, qb = (row + RPT) * RL + off;
            const bool onb = row + RPT < rows;
            c128* da = xb + __ldg(q_off + qa);
            c128* db = onb ? xb + __ldg(q_off + qb) : nullptr;
            c128 olda = cmake(0., 0.), oldb = cmake(0., 0.);
            if (add) { olda = *da; if (onb) oldb = *db; }
            c128 acca = cmake(0., 0.), accb = cmake(0., 0.);
            if (NE) {   // every load of the trip is issued before the first product
                constexpr int NEC = NE ? NE : 1;
                c128 pa[NEC], pq[NEC];
#pragma unroll
                for (int e = 0; e < NEC; e++) {
                    pa[e] = ld_stream(pb + (int64_t)e * g.bl + qa);
                    pq[e] = onb ? ld_stream(pb + (int64_t)e * g.bl + qb) : cmake(0., 0.);
                }
#pragma unroll
                for (int e = 0; e < NEC; e++) {
                    const c128 ce = __ldg(a + e);
                    acca = cadd(acca, cmul(ce, pa[e]));
                    accb = cadd(accb, cmul(ce, pq[e]));
                }
            } else {
#pragma unroll 4
                for (int e = 0; e < ne; e++) {
                    const c128 ce = __ldg(a + e);
                    acca = cadd(acca, cmul(ce, ld_stream(pb + (int64_t)e * g.bl + qa)));
                    if (onb) accb = cadd(accb, cmul(ce, ld_stream(pb + (int64_t)e * g.bl + qb)));
                }
            }
            *da = add ? cadd(olda, acca) : acca;
            if (onb) *db = add ? cadd(oldb, accb) : accb;
        }
    }
}

static AggGeom agg_geom(const LevelGeom& g) {
    AggGeom a;
    for (int i = 0; i < 4; i++) { a.bd[i] = (int)g.bd[i]; a.sub[i] = (int)g.sub[i]; a.sd[i] = (int)g.sd[i]; }
    a.dof = g.dof;
    a.linear = (g.sub[0] == 1 && g.sub[1] == 1 && g.sub[2] == 1) ? 1 : 0;
    return a;
}

// Persistent grid (the CTAs that are resident at once stride over the aggregates) or one warp-task per warp (grid = all the work).
// Measured inside the 512^3 solve (profiles/r02_transfer_knobs.txt): restrict is faster persistent (5.53 against 5.23 TB/s on
// the same box), prolong one-shot (5.74 against 5.60) -- the opposite of what the standalone loop shows for restrict
// (profiles/r02_kbench_transfer_rows.txt), where the SM clock is not under the power cap.  Knobs: MGCR_RESTRICT_ONESHOT,
// MGCR_PROLONG_ONESHOT.
static bool transfer_oneshot() {
    static const int v = getenv("MGCR_RESTRICT_ONESHOT") ? atoi(getenv("MGCR_RESTRICT_ONESHOT")) : 0;
    return v != 0;
}
static bool prolong_oneshot() {
    static const int v = getenv("MGCR_PROLONG_ONESHOT") ? atoi(getenv("MGCR_PROLONG_ONESHOT")) : 1;
    return v != 0;
}
static bool restrict_fold() {   // experiment knob: 0 = one butterfly per value (k_restrict_warp) also for ne = 4
    static const int v = getenv("MGCR_RESTRICT_FOLD") ? atoi(getenv("MGCR_RESTRICT_FOLD")) : 1;
    return v != 0;
}
static int transfer_rows() {
    static const int v = getenv("MGCR_PROLONG_ROWS") ? atoi(getenv("MGCR_PROLONG_ROWS")) : 1;
    return v;
}

static int mg_restrict(mgcr_ctx* ctx, MgLevel& L, const c128* xf, c128* xc) {
    const LevelGeom& g = L.g;
    if (g.nb == 0) return MGCR_OK;
    const double bytes = 16. * L.n * (1 + g.ne) + 16. * L.nc;
    const AggGeom ag = agg_geom(g);
    const int32_t* qo = L.d_q_off;
    const c128* P = L.d_P;
    // persistent grids: exactly the CTAs that are resident at once (or fewer when there is less work)
#define RESTRICT_LAUNCH(KERNEL, lanes_per_agg)                                                                                              \
    do {                                                                                                                                    \
        const int64_t need = (g.nb * (lanes_per_agg) + 255) / 256;                                                                          \
        const int64_t cap = transfer_oneshot() ? need : resident_ctas(ctx, (const void*)KERNEL, 256);                                       \
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cap, need));                                                  \
        launch_pdl(ctx, KERNEL, grid, 256, 0, g, ag, qo, P, xf, xc);                                                                        \
    } while (0)
    if (g.bl <= 256) {
        ProfScope ps_(ctx, "mg_restrict", bytes);
        if (g.bl <= 8) RESTRICT_LAUNCH(k_restrict_sub<8>, 8);
        else if (g.bl <= 16) RESTRICT_LAUNCH(k_restrict_sub<16>, 16);
        else if (g.ne == 4 && restrict_fold()) {
            if (g.bl <= 64) RESTRICT_LAUNCH(k_restrict_warp4<2>, 32);
            else if (g.bl <= 128) RESTRICT_LAUNCH(k_restrict_warp4<4>, 32);
            else RESTRICT_LAUNCH(k_restrict_warp4<8>, 32);
        }
        else if (g.bl <= 64) RESTRICT_LAUNCH(k_restrict_warp<2>, 32);
        else if (g.bl <= 128) RESTRICT_LAUNCH(k_restrict_warp<4>, 32);
        else RESTRICT_LAUNCH(k_restrict_warp<8>, 32);
    } else {
        int threads = 32 * (int)std::min<int64_t>(8, std::max<int64_t>(1, g.ne));
        size_t smem = sizeof(c128) * (size_t)g.bl;
        const unsigned grid = (unsigned)std::min<int64_t>(g.nb, (int64_t)ctx->num_sms * 4);
        KLAUNCH(ctx, "mg_restrict", bytes, (launch_pdl(ctx, k_restrict, grid, threads, smem, g, ag, qo, P, xf, xc)));
    }
#undef RESTRICT_LAUNCH
    CHECK_LAUNCH();
    return MGCR_OK;
}

static int mg_prolong(mgcr_ctx* ctx, MgLevel& L, const c128* xc, c128* xf, bool add = false) {
    const LevelGeom& g = L.g;
    if (L.n == 0) return MGCR_OK;
    const double bytes = 16. * L.n * (1 + g.ne + (add ? 1 : 0)) + 16. * L.nc;
    const AggGeom ag = agg_geom(g);
    const int64_t rl = g.sub[3] * g.dof;
    if (transfer_rows() && !ag.linear && rl == 4 && (g.bl / rl) % 4 == 0 && g.bd[3] % 2 == 0) {   // 64-byte aggregate rows: two aggregates per warp
        const int64_t need = (g.nb / 2 * 32 + 255) / 256;
#define PROLONG_ROWS(KERNEL)                                                                                                                  \
    do {                                                                                                                                      \
        const int64_t cap = prolong_oneshot() ? need : resident_ctas(ctx, (const void*)KERNEL, 256);                                          \
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cap, need));                                                    \
        KLAUNCH(ctx, "mg_prolong", bytes, (launch_pdl(ctx, KERNEL, grid, 256, 0, g, ag, (const int32_t*)L.d_q_off, (const c128*)L.d_P, xc, xf, add ? 1 : 0))); \
    } while (0)
        if (g.ne == 4 && transfer_rows() != 2) PROLONG_ROWS((k_prolong_rows<4, 2, 4>));   // (MGCR_PROLONG_ROWS=2: the runtime-ne form, experiment)
        else PROLONG_ROWS((k_prolong_rows<4, 2, 0>));
#undef PROLONG_ROWS
        CHECK_LAUNCH();
        return MGCR_OK;
    }
    const int64_t need = (g.nb * 32 + 255) / 256;
    const int64_t cap = prolong_oneshot() ? need : resident_ctas(ctx, (const void*)k_prolong, 256);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cap, need));
    KLAUNCH(ctx, "mg_prolong", bytes, (launch_pdl(ctx, k_prolong, grid, 256, 0, g, ag, (const int32_t*)L.d_q_off, (const c128*)L.d_P, xc, xf, add ? 1 : 0)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// MG as an operator + the cycle
// ----------------------------------------------------------------------------------------------------------
static int mg_cycle(mgcr_mg* mg, int l, const c128* b, c128* x);

struct MgOp : mgcr_op {
    mgcr_mg* mg = nullptr;
    int level = 0;
    int apply(const c128* x, c128* y) override { return mg_cycle(mg, level, x, y); }
};

// Report Algorithm 2 (SemesterProject.pdf p.4) in the structure of src/MG.h:405-430 with the defects of SURVEY.md facts
// 6-7 removed:  x = S b ; r = b - A x ; xc = Ac^-1 (P^H r) ; x += P xc ; r = b - A x ; x += S r.  S = one call of the
// smoother GCR (its solve ADDS to x, GCR.h:232).  The coarse solve is GCR on the block-CSR operator, right-
// preconditioned by this same cycle one level down when that level exists (K-cycle), plain on the coarsest level.
static int mg_cycle(mgcr_mg* mg, int l, const c128* b, c128* x) {
    mgcr_ctx* ctx = mg->ctx;
    MgLevel& L = mg->lv[l];
    const int64_t nc = L.nc;
    (void)nc;
    ARG_CHECK(b != x, "MG cycle: input and output alias");
    MGCR_TRY(gcr_solve(ctx, L.A, &mg->smooth, nullptr, b, x, nullptr, 0, nullptr, true));      // x = S b from a zero start: x is written, never read
    MGCR_TRY(L.A->apply_residual(x, b, L.d_r));                            // r = b - A x, one kernel
    MGCR_TRY(mg_restrict(ctx, L, L.d_r, L.d_rc));
    const c128* xc = L.d_xc;
    if (!L.gather) {
        MGCR_TRY(gcr_solve(ctx, L.Ac, &mg->coarse, L.deeper, L.d_rc, L.d_xc, nullptr, 0, nullptr, true));
    } else {
        // coarse-level gather: every rank assembles the whole coarse right-hand side, solves the
        // replicated coarse system (identical arithmetic on every GPU, no further communication) and keeps its slice
        int64_t maxc = 0;
        bool equal = true;
        for (int64_t c : L.nc_counts) { maxc = std::max(maxc, c); equal = equal && c == L.nc_counts[0]; }
        if (equal) {
            MGCR_TRY(dist_allgather(ctx, L.d_rc, L.d_rc_full, sizeof(c128) * (size_t)maxc));
        } else {
            CUDA_TRY(cudaMemcpyAsync(L.d_pad, L.d_rc, sizeof(c128) * nc, cudaMemcpyDeviceToDevice, ctx->stream));
            MGCR_TRY(dist_allgather(ctx, L.d_pad, L.d_pad + maxc, sizeof(c128) * (size_t)maxc));
            int64_t off = 0;
            for (size_t r = 0; r < L.nc_counts.size(); r++) {
                CUDA_TRY(cudaMemcpyAsync(L.d_rc_full + off, L.d_pad + maxc * (int64_t)(r + 1), sizeof(c128) * L.nc_counts[r], cudaMemcpyDeviceToDevice, ctx->stream));
                off += L.nc_counts[r];
            }
        }
        MGCR_TRY(gcr_solve(ctx, L.Ac_full, &mg->coarse, L.deeper, L.d_rc_full, L.d_xc_full, nullptr, 0, nullptr, true));
        xc = L.d_xc_full + L.nc_offset;
    }
    MGCR_TRY(mg_prolong(ctx, L, xc, x, true));                             // x += P xc, one kernel
    MGCR_TRY(L.A->apply_residual(x, b, L.d_r));
    MGCR_TRY(gcr_solve(ctx, L.A, &mg->smooth, nullptr, L.d_r, x, nullptr, 0, nullptr));
    return MGCR_OK;
}

// ----------------------------------------------------------------------------------------------------------
// setup
// ----------------------------------------------------------------------------------------------------------
template <class Rows>
static int galerkin_launch(mgcr_ctx* ctx, MgLevel& L, const Rows& rows, int32_t* bcol, c128* bval) {
    const LevelGeom& g = L.g;
    const int qc = std::max(1, 256 / g.ne);
    size_t smem = sizeof(c128) * ((size_t)L.K * g.ne * g.ne + (size_t)qc * L.K * g.ne);
    ARG_CHECK(smem <= 200 * 1024, "MG setup: %d near-null vectors per aggregate need %zu bytes of shared memory for the coarse blocks", g.ne, smem);
    MGCR_TRY(ensure_dyn_smem(ctx, (const void*)k_galerkin<Rows>, (int)smem));
    KLAUNCH(ctx, "mg_galerkin", 0., (k_galerkin<Rows><<<(unsigned)g.nb, 256, smem, ctx->stream>>>(rows, g, L.K, L.Ac->d_brow, L.d_block_map, L.d_site_block,
                                                                                                L.d_site_off, L.d_P, L.d_Pg, bcol, L.d_bslot, bval)));
    CHECK_LAUNCH();
    return MGCR_OK;
}

static int galerkin(mgcr_ctx* ctx, MgLevel& L, int32_t* bcol, c128* bval) {
    int st = MGCR_OK;
    if (with_rows(L.A, [&](auto rows) { return galerkin_launch(ctx, L, rows, bcol, bval); }, &st, true)) return st;
    mgcr_set_error("MG setup: the operator of this level has no accessible matrix entries (kind %d)", (int)L.A->kind);
    return MGCR_ERR_UNSUPPORTED;
}

// Replicated copy of a slab-partitioned coarse operator: every rank contributes its block rows (ghost columns turned
// into global block columns), all ranks assemble the same block-CSR.  One-off, staged through the host.
static int gather_coarse_operator(mgcr_ctx* ctx, MgLevel& L, int64_t nb_offset, int64_t nb_global, BlockCsrOp** out) {
    const LevelGeom& g = L.g;
    BlockCsrOp* loc = L.Ac;
    const int ne = g.ne;
    const size_t bsz = (size_t)ne * ne;
    std::vector<int64_t> nbs, nnzs;
    MGCR_TRY(dist_allgather_host_i64(ctx, g.nb, nbs));
    MGCR_TRY(dist_allgather_host_i64(ctx, loc->nnzb, nnzs));
    int64_t max_nb = 0, max_nnz = 0, tot_nnz = 0;
    for (size_t r = 0; r < nbs.size(); r++) { max_nb = std::max(max_nb, nbs[r]); max_nnz = std::max(max_nnz, nnzs[r]); tot_nnz += nnzs[r]; }
    ARG_CHECK(tot_nnz < (int64_t)INT32_MAX, "MG setup: gathered coarse operator too large");
    // global block columns on the host
    std::vector<int32_t> hrow((size_t)g.nb + 1), hcol((size_t)loc->nnzb);
    CUDA_TRY(cudaMemcpyAsync(hrow.data(), loc->d_brow, sizeof(int32_t) * hrow.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(hcol.data(), loc->d_bcol, sizeof(int32_t) * hcol.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const int64_t lo_blocks = g.has_lo ? g.plane_blocks : 0;
    for (auto& c : hcol) {
        int64_t cc = c;
        if (cc < g.nb) cc += nb_offset;
        else if (cc - g.nb < lo_blocks) cc = nb_offset - g.plane_blocks + (cc - g.nb);
        else cc = nb_offset + g.nb + (cc - g.nb - lo_blocks);
        c = (int32_t)cc;
    }
    // padded exchange: [row counts | cols | vals]
    int32_t *d_cnt = nullptr, *d_cnt_all = nullptr, *d_col = nullptr, *d_col_all = nullptr;
    c128 *d_val_all = nullptr;
    const int R = ctx->nranks;
    std::vector<int32_t> cnt((size_t)max_nb, 0);
    for (int64_t b = 0; b < g.nb; b++) cnt[b] = hrow[b + 1] - hrow[b];
    std::vector<int32_t> colpad((size_t)max_nnz, 0);
    std::copy(hcol.begin(), hcol.end(), colpad.begin());
    MGCR_TRY(dev_alloc_t(ctx, (size_t)max_nb, &d_cnt));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)max_nb * R, &d_cnt_all));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)max_nnz, &d_col));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)max_nnz * R, &d_col_all));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)max_nnz * bsz * R, &d_val_all));
    CUDA_TRY(cudaMemcpyAsync(d_cnt, cnt.data(), sizeof(int32_t) * max_nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_col, colpad.data(), sizeof(int32_t) * max_nnz, cudaMemcpyHostToDevice, ctx->stream));
    // values: send straight from the local operator (padded region of the receive slots is never read)
    c128* d_val_send = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)max_nnz * bsz, &d_val_send));
    CUDA_TRY(cudaMemcpyAsync(d_val_send, loc->d_bval, sizeof(c128) * (size_t)loc->nnzb * bsz, cudaMemcpyDeviceToDevice, ctx->stream));
    MGCR_TRY(dist_allgather(ctx, d_cnt, d_cnt_all, sizeof(int32_t) * (size_t)max_nb));
    MGCR_TRY(dist_allgather(ctx, d_col, d_col_all, sizeof(int32_t) * (size_t)max_nnz));
    MGCR_TRY(dist_allgather(ctx, d_val_send, d_val_all, sizeof(c128) * (size_t)max_nnz * bsz));
    std::vector<int32_t> cnt_all((size_t)max_nb * R), col_all((size_t)max_nnz * R);
    CUDA_TRY(cudaMemcpyAsync(cnt_all.data(), d_cnt_all, sizeof(int32_t) * cnt_all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(col_all.data(), d_col_all, sizeof(int32_t) * col_all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    BlockCsrOp* full = new BlockCsrOp();
    full->kind = OP_BLOCKCSR; full->ctx = ctx; full->nb = nb_global; full->nb_cols = nb_global; full->ne = ne; full->nnzb = tot_nnz;
    full->n_local = nb_global * ne; full->n_global = full->n_local; full->distributed = false;
    std::vector<int32_t> frow((size_t)nb_global + 1, 0), fcol((size_t)tot_nnz);
    MGCR_TRY(dev_alloc_t(ctx, frow.size(), &full->d_brow));
    MGCR_TRY(dev_alloc_t(ctx, fcol.size(), &full->d_bcol));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)tot_nnz * bsz, &full->d_bval));
    int64_t row = 0, pos = 0;
    for (int r = 0; r < R; r++) {
        for (int64_t b = 0; b < nbs[r]; b++) { frow[row + 1] = frow[row] + cnt_all[(size_t)r * max_nb + b]; row++; }
        std::copy(col_all.begin() + (size_t)r * max_nnz, col_all.begin() + (size_t)r * max_nnz + nnzs[r], fcol.begin() + pos);
        CUDA_TRY(cudaMemcpyAsync(full->d_bval + (size_t)pos * bsz, d_val_all + (size_t)r * max_nnz * bsz, sizeof(c128) * (size_t)nnzs[r] * bsz,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        pos += nnzs[r];
    }
    CUDA_TRY(cudaMemcpyAsync(full->d_brow, frow.data(), sizeof(int32_t) * frow.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(full->d_bcol, fcol.data(), sizeof(int32_t) * fcol.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, d_cnt); dev_free(ctx, d_cnt_all); dev_free(ctx, d_col); dev_free(ctx, d_col_all); dev_free(ctx, d_val_all); dev_free(ctx, d_val_send);
    *out = full;
    return MGCR_OK;
}

static int level_setup(mgcr_mg* mg, int l, const c128* d_nearnull) {
    mgcr_ctx* ctx = mg->ctx;
    MgLevel& L = mg->lv[l];
    const mgcr_level_cfg& c = L.cfg;
    LevelGeom& g = L.g;
    ARG_CHECK(c.n_spin >= 1 && c.n_col >= 1 && c.n_eigen >= 1, "MG level %d: n_spin, n_col, n_eigen must be positive", l);
    g.dof = c.n_spin * c.n_col;
    const bool dist = L.A->distributed;
    g.dist = dist ? 1 : 0; g.pd = 0; g.has_lo = 0; g.has_hi = 0; g.plane_sites = 0; g.plane_blocks = 0;
    int64_t sd_local[4];
    for (int i = 0; i < 4; i++) sd_local[i] = c.site_dims[i];
    if (dist) {
        // cfg carries the GLOBAL lattice; this rank holds a slab of planes of the first dimension of extent > 1
        ARG_CHECK(mg->flags == 0 || !(mg->flags & MGCR_MG_NEG_NEIGHBOUR_BUG), "MG level %d: the negative-neighbour switch is single-GPU only", l);
        int pd = 0;
        while (pd < 3 && c.site_dims[pd] == 1) pd++;
        int64_t plane = 1;
        for (int i = pd + 1; i < 4; i++) plane *= c.site_dims[i];
        ARG_CHECK(L.A->n_local % (plane * g.dof) == 0, "MG level %d: the operator's slab (%lld rows) is not a whole number of lattice planes", l, (long long)L.A->n_local);
        sd_local[pd] = L.A->n_local / (plane * g.dof);
        ARG_CHECK(sd_local[pd] >= 1 && c.sub[pd] >= 1 && sd_local[pd] % c.sub[pd] == 0,
                  "MG level %d: this rank's slab of %lld planes is not a multiple of the aggregate size %lld (mgcr_ctx_set_slab_align)", l,
                  (long long)sd_local[pd], (long long)c.sub[pd]);
        g.pd = pd; g.has_lo = ctx->rank > 0; g.has_hi = ctx->rank + 1 < ctx->nranks;
        g.plane_sites = plane;
    }
    L.nsite = 1; g.bs = 1; g.nb = 1;
    for (int i = 0; i < 4; i++) {
        ARG_CHECK(sd_local[i] >= 1 && c.sub[i] >= 1 && sd_local[i] % c.sub[i] == 0,
                  "MG level %d: dimension %lld is not divisible by the aggregate size %lld (src/Mesh.h:245)", l, (long long)sd_local[i], (long long)c.sub[i]);
        g.sd[i] = sd_local[i]; g.sub[i] = c.sub[i]; g.bd[i] = sd_local[i] / c.sub[i];
        L.nsite *= g.sd[i]; g.bs *= g.sub[i]; g.nb *= g.bd[i];
    }
    if (dist) {
        g.plane_blocks = 1;
        for (int i = g.pd + 1; i < 4; i++) g.plane_blocks *= g.bd[i];
    }
    L.n = L.nsite * g.dof;
    ARG_CHECK(L.n == L.A->n_local, "MG level %d: mesh has %lld dofs, the operator %lld (src/GCR.h:160)", l, (long long)L.n, (long long)L.A->n_local);
    ARG_CHECK(L.nsite < (int64_t)INT32_MAX, "MG level %d: too many sites for one device shard", l);
    const bool doubled = (c.n_spin == 4);
    g.ne = doubled ? 2 * c.n_eigen : c.n_eigen;
    g.bl = g.bs * g.dof;
    L.nc = g.nb * g.ne;
    const int64_t n = L.n;
    const int nv = c.n_eigen, ne = g.ne;
    // aggregation (Mesh::blocking)
    SetupStage* stage = new SetupStage(mg, "aggregate");
    struct StageGuard { SetupStage** s; ~StageGuard() { delete *s; *s = nullptr; } } stage_guard{&stage};
    auto next_stage = [&](const char* name) { delete stage; stage = new SetupStage(mg, name); };
    MGCR_TRY(dev_alloc_t(ctx, (size_t)L.nsite, &L.d_block_map));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)L.nsite, &L.d_site_block));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)L.nsite, &L.d_site_off));
    int64_t bd[4];
    MGCR_TRY(blocking_device(ctx, g.sd, g.sub, bd, L.d_block_map, L.d_site_block, L.d_site_off));
    ARG_CHECK(g.bl < (int64_t)INT32_MAX && L.n < ((int64_t)1 << 40), "MG level %d: aggregate too large", l);
    MGCR_TRY(dev_alloc_t(ctx, (size_t)g.bl, &L.d_q_off));
    k_agg_offsets<<<(unsigned)std::min<int64_t>(64, (g.bl + 255) / 256), 256, 0, ctx->stream>>>(agg_geom(g), (int)g.bl, L.d_q_off);
    CHECK_LAUNCH();
    // near-null vectors (Arnoldi::solve) and chirality doubling
    next_stage("near_null");
    const double rand_ms0 = ctx->rand_seconds_ms;
    c128* ev = nullptr;
    MGCR_TRY(dev_alloc_t(ctx, (size_t)n * ne, &ev));
    c128* base = doubled ? nullptr : ev;
    c128* raw = nullptr;
    if (doubled) { MGCR_TRY(dev_alloc_t(ctx, (size_t)n * nv, &raw)); base = raw; }
    if (d_nearnull) CUDA_TRY(cudaMemcpyAsync(base, d_nearnull, sizeof(c128) * n * nv, cudaMemcpyDeviceToDevice, ctx->stream));
    else MGCR_TRY(arnoldi(ctx, L.A, &mg->eigen, nv, base));
    if (doubled) {
        c128* g5 = nullptr;
        MGCR_TRY(dev_alloc_t(ctx, (size_t)n, &g5));
        for (int i = 0; i < nv; i++) {
            const c128* v = raw + (int64_t)i * n;
            MGCR_TRY(vec_gamma5(ctx, n, c.n_col, c.n_spin, v, g5));
            KLAUNCH(ctx, "mg_chiral", 64. * n, (k_chiral<<<stream_grid(ctx, n, 8), RED_THREADS, 0, ctx->stream>>>(n, v, g5, ev + (int64_t)i * n, ev + (int64_t)(i + nv) * n)));
            CHECK_LAUNCH();
        }
        MGCR_TRY(dev_free(ctx, g5));
        MGCR_TRY(dev_free(ctx, raw));
    }
    // block projection into compact storage + per-aggregate Gram-Schmidt
    mg->setup_s["rand"] += 1e-3 * (ctx->rand_seconds_ms - rand_ms0);   // part of near_null: the init_rand(9) start vector
    next_stage("project_orthonormalise");
    MGCR_TRY(dev_alloc_t(ctx, (size_t)n * ne, &L.d_P));
    KLAUNCH(ctx, "mg_project", 32. * n * ne, (k_project<<<stream_grid(ctx, n * ne, 8), RED_THREADS, 0, ctx->stream>>>(n * ne, g, n, L.d_block_map, ev, L.d_P)));
    CHECK_LAUNCH();
    MGCR_TRY(dev_free(ctx, ev));
    KLAUNCH(ctx, "mg_block_mgs", 0., (k_block_mgs<<<(unsigned)g.nb, 128, 0, ctx->stream>>>(g, L.d_P)));
    CHECK_LAUNCH();
    // slab-partitioned level: the prolongator rows of the neighbour ranks' adjacent planes (ghost sites)
    next_stage("ghost_prolongator");
    const int64_t lo_blocks = g.has_lo ? g.plane_blocks : 0, hi_blocks = g.has_hi ? g.plane_blocks : 0;
    if (dist && (g.has_lo || g.has_hi)) {
        const size_t plane_elems = (size_t)g.plane_sites * g.dof * ne;
        c128 *send_lo = nullptr, *send_hi = nullptr;
        MGCR_TRY(dev_alloc_t(ctx, plane_elems * ((g.has_lo ? 1 : 0) + (g.has_hi ? 1 : 0)), &L.d_Pg));
        MGCR_TRY(dev_alloc_t(ctx, plane_elems, &send_lo));
        MGCR_TRY(dev_alloc_t(ctx, plane_elems, &send_hi));
        const int pgrid = stream_grid(ctx, (int64_t)plane_elems, 4);
        if (g.has_lo) { KLAUNCH(ctx, "mg_pack_P", 32. * plane_elems, (k_pack_P_plane<<<pgrid, RED_THREADS, 0, ctx->stream>>>(g, 0, L.d_site_block, L.d_site_off, L.d_P, send_lo))); CHECK_LAUNCH(); }
        if (g.has_hi) { KLAUNCH(ctx, "mg_pack_P", 32. * plane_elems, (k_pack_P_plane<<<pgrid, RED_THREADS, 0, ctx->stream>>>(g, L.nsite - g.plane_sites, L.d_site_block, L.d_site_off, L.d_P, send_hi))); CHECK_LAUNCH(); }
        // over peer memory when the planes fit the heap (a temporary receive area, handed back right away): the first NCCL
        // send/recv of a process sets up its point-to-point channels, 0.6-1.5 s of the 2-GPU set-up in round 2's first measurement
        PeerHalo tmp;
        MGCR_TRY(p2p_halo_create(ctx, (int64_t)plane_elems, &tmp));
        if (tmp.on) {
            const c128 *recv_lo = nullptr, *recv_hi = nullptr;
            MGCR_TRY(p2p_halo_exchange(ctx, &tmp, send_lo, send_hi, &recv_lo, &recv_hi));
            if (g.has_lo) CUDA_TRY(cudaMemcpyAsync(L.d_Pg, recv_lo, sizeof(c128) * plane_elems, cudaMemcpyDeviceToDevice, ctx->stream));
            if (g.has_hi) CUDA_TRY(cudaMemcpyAsync(L.d_Pg + (g.has_lo ? plane_elems : 0), recv_hi, sizeof(c128) * plane_elems, cudaMemcpyDeviceToDevice, ctx->stream));
            p2p_halo_destroy(ctx, &tmp);
        } else {
            MGCR_TRY(dist_group_begin(ctx));
            if (g.has_lo) {
                MGCR_TRY(dist_send(ctx, send_lo, sizeof(c128) * plane_elems, ctx->rank - 1, ctx->stream));
                MGCR_TRY(dist_recv(ctx, L.d_Pg, sizeof(c128) * plane_elems, ctx->rank - 1, ctx->stream));
            }
            if (g.has_hi) {
                MGCR_TRY(dist_send(ctx, send_hi, sizeof(c128) * plane_elems, ctx->rank + 1, ctx->stream));
                MGCR_TRY(dist_recv(ctx, L.d_Pg + (g.has_lo ? plane_elems : 0), sizeof(c128) * plane_elems, ctx->rank + 1, ctx->stream));
            }
            MGCR_TRY(dist_group_end(ctx));
        }
        MGCR_TRY(dev_free(ctx, send_lo));
        MGCR_TRY(dev_free(ctx, send_hi));
    }
    // Galerkin coarse operator, written straight into the block-CSR compute layout
    next_stage("galerkin");
    std::vector<int32_t> hrow((size_t)g.nb + 1, 0);
    int kmax = 0;
    {
        RowSlots rs;
        for (int64_t B = 0; B < g.nb; B++) {
            row_slots(g, B, &rs);
            hrow[B + 1] = hrow[B] + rs.K;
            kmax = std::max(kmax, rs.K);
        }
    }
    L.K = kmax;
    BlockCsrOp* Ac = new BlockCsrOp();
    Ac->kind = OP_BLOCKCSR; Ac->ctx = ctx; Ac->nb = g.nb; Ac->nb_cols = g.nb + lo_blocks + hi_blocks; Ac->ne = ne; Ac->nnzb = hrow[g.nb];
    Ac->n_local = L.nc; Ac->n_global = L.nc; Ac->distributed = dist;
    L.Ac = Ac;
    ARG_CHECK(Ac->nnzb < (int64_t)INT32_MAX, "MG level %d: coarse operator too large for one device shard", l);
    MGCR_TRY(dev_alloc_t(ctx, (size_t)g.nb + 1, &Ac->d_brow));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)Ac->nnzb, &Ac->d_bcol));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)Ac->nnzb * ne * ne, &Ac->d_bval));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)Ac->nnzb, &L.d_bslot));
    CUDA_TRY(cudaMemcpyAsync(Ac->d_brow, hrow.data(), sizeof(int32_t) * hrow.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    MGCR_TRY(galerkin(ctx, L, Ac->d_bcol, Ac->d_bval));
    if (mg->flags & MGCR_MG_NEG_NEIGHBOUR_BUG) {
        KLAUNCH(ctx, "mg_neg_bug", 0., (k_neg_bug<<<(unsigned)g.nb, 128, 0, ctx->stream>>>(g, Ac->d_brow, L.d_bslot, Ac->d_bval)));
        CHECK_LAUNCH();
    }
    L.nc_global = L.nc; L.nc_offset = 0;
    next_stage("coarse_halo_gather");
    if (dist) {
        // the coarse operator inherits the slab partition: one plane of aggregates to / from each neighbour per apply
        MGCR_TRY(dist_allgather_host_i64(ctx, L.nc, L.nc_counts));
        L.nc_global = 0;
        for (int r = 0; r < ctx->nranks; r++) { if (r < ctx->rank) L.nc_offset += L.nc_counts[r]; L.nc_global += L.nc_counts[r]; }
        Ac->n_global = L.nc_global;
        HaloPlan* h = new HaloPlan();
        h->elem = ne;
        h->send_off.push_back(0); h->recv_off.push_back(0);
        if (g.has_lo) {
            h->peer.push_back(ctx->rank - 1); h->send_start.push_back(0);
            h->send_off.push_back(h->send_off.back() + g.plane_blocks); h->recv_off.push_back(h->recv_off.back() + g.plane_blocks);
        }
        if (g.has_hi) {
            h->peer.push_back(ctx->rank + 1); h->send_start.push_back(g.nb - g.plane_blocks);
            h->send_off.push_back(h->send_off.back() + g.plane_blocks); h->recv_off.push_back(h->recv_off.back() + g.plane_blocks);
        }
        h->npeers = (int)h->peer.size();
        h->n_ghost = lo_blocks + hi_blocks;
        Ac->halo = h;
        Ac->halo_rows_lo = lo_blocks; Ac->halo_rows_hi = hi_blocks;
        MGCR_TRY(dev_alloc_t(ctx, (size_t)std::max<int64_t>(1, h->n_ghost * ne), &h->d_ghost));
        MGCR_TRY(p2p_halo_create(ctx, g.plane_blocks * ne, &h->ph));   // same size and order on every rank
        // gather here?  yes when the next level cannot keep the slab partition (a rank's aggregates no longer divide) or
        // the coarse system is small enough that communication latency dominates (option gather_dofs, default 2^18)
        bool gather = L.nc_global <= ctx->gather_dofs;
        if (!gather && l + 1 < mg->n_level) {
            const int64_t sub_next = mg->lv[l + 1].cfg.sub[g.pd];
            std::vector<int64_t> bds;
            MGCR_TRY(dist_allgather_host_i64(ctx, g.bd[g.pd], bds));
            for (int64_t b : bds) gather = gather || sub_next < 1 || (b % sub_next) != 0;
        }
        L.gather = gather;
        if (gather) {
            int64_t nb_global = L.nc_global / ne, nb_offset = L.nc_offset / ne;
            MGCR_TRY(gather_coarse_operator(ctx, L, nb_offset, nb_global, &L.Ac_full));
            int64_t maxc = 0;
            for (int64_t cnt : L.nc_counts) maxc = std::max(maxc, cnt);
            MGCR_TRY(dev_alloc_t(ctx, (size_t)L.nc_global, &L.d_rc_full));
            MGCR_TRY(dev_alloc_t(ctx, (size_t)L.nc_global, &L.d_xc_full));
            MGCR_TRY(dev_alloc_t(ctx, (size_t)maxc * (ctx->nranks + 1), &L.d_pad));
            CUDA_TRY(cudaMemsetAsync(L.d_pad, 0, sizeof(c128) * (size_t)maxc * (ctx->nranks + 1), ctx->stream));
        }
    }
    // cycle work vectors
    MGCR_TRY(dev_alloc_t(ctx, (size_t)n, &L.d_r));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)L.nc, &L.d_rc));
    MGCR_TRY(dev_alloc_t(ctx, (size_t)L.nc, &L.d_xc));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return MGCR_OK;
}

static void level_free(mgcr_ctx* ctx, MgLevel& L) {
    dev_free(ctx, L.d_q_off);
    dev_free(ctx, L.d_block_map); dev_free(ctx, L.d_site_block); dev_free(ctx, L.d_site_off); dev_free(ctx, L.d_P);
    dev_free(ctx, L.d_bslot); dev_free(ctx, L.d_r); dev_free(ctx, L.d_rc); dev_free(ctx, L.d_xc);
    dev_free(ctx, L.d_Pg); dev_free(ctx, L.d_rc_full); dev_free(ctx, L.d_xc_full); dev_free(ctx, L.d_pad);
    delete L.deeper; L.deeper = nullptr;
    delete L.Ac; L.Ac = nullptr;
    delete L.Ac_full; L.Ac_full = nullptr;
}

extern "C" int mgcr_mg_destroy(mgcr_mg* mg) {
    if (!mg) return MGCR_OK;
    for (auto& L : mg->lv) level_free(mg->ctx, L);
    cudaStreamSynchronize(mg->ctx->stream);
    delete mg;
    return MGCR_OK;
}

extern "C" int mgcr_mg_create(mgcr_ctx* ctx, mgcr_op* A, int n_level, const mgcr_level_cfg* cfg, const mgcr_gcr_param* eigen,
                              const mgcr_gcr_param* coarse, const mgcr_gcr_param* smooth, int flags, const mgcr_c128* d_nearnull0, mgcr_mg** out) {
    std::vector<const mgcr_c128*> nn((size_t)std::max(n_level, 1), nullptr);
    nn[0] = d_nearnull0;
    return mgcr_mg_create_nn(ctx, A, n_level, cfg, eigen, coarse, smooth, flags, nn.data(), out);
}

extern "C" int mgcr_mg_create_nn(mgcr_ctx* ctx, mgcr_op* A, int n_level, const mgcr_level_cfg* cfg, const mgcr_gcr_param* eigen,
                                 const mgcr_gcr_param* coarse, const mgcr_gcr_param* smooth, int flags, const mgcr_c128* const* d_nearnull,
                                 mgcr_mg** out) {
    ARG_CHECK(ctx && A && cfg && eigen && coarse && smooth && out && n_level >= 1, "mgcr_mg_create: bad argument");
    *out = nullptr;
    mgcr_mg* mg = new mgcr_mg();
    mg->ctx = ctx; mg->n_level = n_level; mg->flags = flags;
    mg->eigen = *eigen; mg->coarse = *coarse; mg->smooth = *smooth;
    const int sc = (flags & MGCR_MG_STD_CONJ) ? 1 : 0;
    mg->eigen.std_conj = sc; mg->coarse.std_conj = sc; mg->smooth.std_conj = sc;
    mg->eigen.verbose = 0; mg->coarse.verbose = 0; mg->smooth.verbose = 0;
    mg->lv.resize((size_t)n_level);
    for (int l = 0; l < n_level; l++) mg->lv[l].cfg = cfg[l];
    mgcr_op* cur = A;
    cudaStreamSynchronize(ctx->stream);
    const auto t_total0 = std::chrono::steady_clock::now();
    for (int l = 0; l < n_level; l++) {
        mg->lv[l].A = cur;
        int st = level_setup(mg, l, d_nearnull ? (const c128*)d_nearnull[l] : nullptr);
        if (st != MGCR_OK) { mgcr_mg_destroy(mg); return st; }
        cur = mg->lv[l].gather ? (mgcr_op*)mg->lv[l].Ac_full : (mgcr_op*)mg->lv[l].Ac;
    }
    for (int l = 0; l + 1 < n_level; l++) {
        MgOp* op = new MgOp();
        op->kind = OP_MG; op->ctx = ctx; op->mg = mg; op->level = l + 1;
        op->n_local = mg->lv[l + 1].n; op->n_global = mg->lv[l + 1].A->n_global; op->distributed = mg->lv[l + 1].A->distributed;
        mg->lv[l].deeper = op;
    }
    // the hierarchy is complete: large coarse operators keep only their streaming image (ops.cu, BlockCsrOp::build_sliced)
    cudaStreamSynchronize(ctx->stream);
    const auto t_img0 = std::chrono::steady_clock::now();
    for (int l = 0; l < n_level; l++) {
        BlockCsrOp* used = mg->lv[l].gather ? mg->lv[l].Ac_full : mg->lv[l].Ac;
        if (!used) continue;
        if (used->sliced == 0) { int st = used->build_sliced(); if (st != MGCR_OK) { mgcr_mg_destroy(mg); return st; } }
        used->drop_assembly_values();
    }
    cudaStreamSynchronize(ctx->stream);
    const auto t_end = std::chrono::steady_clock::now();
    mg->setup_s["streaming_image"] = std::chrono::duration<double>(t_end - t_img0).count();
    mg->setup_s["total"] = std::chrono::duration<double>(t_end - t_total0).count();
    *out = mg;
    return MGCR_OK;
}

// wall-clock seconds per set-up stage, summed over the levels: "total", "aggregate", "near_null" (Arnoldi::solve; "rand" is
// its init_rand part), "project_orthonormalise", "ghost_prolongator", "galerkin", "coarse_halo_gather", "streaming_image"
extern "C" int mgcr_mg_setup_profile(mgcr_mg* mg, int cap, const char** names, double* seconds, int* n_out) {
    ARG_CHECK(mg && n_out, "mgcr_mg_setup_profile: NULL argument");
    int i = 0;
    for (auto& kv : mg->setup_s) {
        if (i < cap) { if (names) names[i] = kv.first.c_str(); if (seconds) seconds[i] = kv.second; }
        i++;
    }
    *n_out = i;
    return MGCR_OK;
}

#define LEVEL_CHECK(mg, level) ARG_CHECK((mg) && (level) >= 0 && (level) < (mg)->n_level, "MG: level %d out of range", (level))

extern "C" int mgcr_mg_level_info(mgcr_mg* mg, int level, int64_t* n_fine, int64_t* n_blocks, int* ne, int64_t* block_len) {
    LEVEL_CHECK(mg, level);
    const MgLevel& L = mg->lv[level];
    if (n_fine) *n_fine = L.n;
    if (n_blocks) *n_blocks = L.g.nb;
    if (ne) *ne = L.g.ne;
    if (block_len) *block_len = L.g.bl;
    return MGCR_OK;
}

extern "C" int mgcr_mg_export_block_map(mgcr_mg* mg, int level, int64_t* h) {
    LEVEL_CHECK(mg, level);
    ARG_CHECK(h, "NULL output");
    const MgLevel& L = mg->lv[level];
    CUDA_TRY(cudaMemcpyAsync(h, L.d_block_map, sizeof(int64_t) * L.nsite, cudaMemcpyDeviceToHost, mg->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(mg->ctx->stream));
    return MGCR_OK;
}

extern "C" int mgcr_mg_export_prolongator(mgcr_mg* mg, int level, mgcr_c128* h) {
    LEVEL_CHECK(mg, level);
    ARG_CHECK(h, "NULL output");
    const MgLevel& L = mg->lv[level];
    CUDA_TRY(cudaMemcpyAsync(h, L.d_P, sizeof(c128) * L.n * L.g.ne, cudaMemcpyDeviceToHost, mg->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(mg->ctx->stream));
    return MGCR_OK;
}

// The reference's pattern (HierarchicalSparse.h:58-98 fed by MG.h:217-276): always 9 blocks per row, explicit zero
// blocks where a neighbour coincides with the block itself or with the other neighbour, sorted by (row, col) with
// ties in triplet order; blocks row-major.
extern "C" int mgcr_mg_export_coarse(mgcr_mg* mg, int level, int64_t* h_brow, int64_t* h_bcol, mgcr_c128* h_bval) {
    LEVEL_CHECK(mg, level);
    ARG_CHECK(h_brow && h_bcol && h_bval, "NULL output");
    const MgLevel& L = mg->lv[level];
    const LevelGeom& g = L.g;
    ARG_CHECK(!g.dist, "mgcr_mg_export_coarse: the reference-pattern export is single-GPU (level %d is slab-partitioned)", level);
    ARG_CHECK(L.Ac->d_bval, "mgcr_mg_export_coarse: level %d keeps only its streaming image (operators above small_gcr_rows rows drop the assembly copy)", level);
    const int ne = g.ne;
    const size_t bsz = (size_t)ne * ne;
    std::vector<c128> val((size_t)L.Ac->nnzb * bsz);
    std::vector<int8_t> slot((size_t)L.Ac->nnzb);
    std::vector<int32_t> prow((size_t)g.nb + 1);
    CUDA_TRY(cudaMemcpyAsync(prow.data(), L.Ac->d_brow, sizeof(int32_t) * prow.size(), cudaMemcpyDeviceToHost, mg->ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(val.data(), L.Ac->d_bval, sizeof(c128) * val.size(), cudaMemcpyDeviceToHost, mg->ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(slot.data(), L.d_bslot, slot.size(), cudaMemcpyDeviceToHost, mg->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(mg->ctx->stream));
    for (int64_t R = 0; R < g.nb; R++) {
        h_brow[R] = 9 * R;
        int64_t bi[4], rem = R;
        for (int c = 3; c >= 0; c--) { bi[c] = rem % g.bd[c]; rem /= g.bd[c]; }
        struct Ent { int64_t col; int s; bool zero; } ent[9];
        ent[0] = {R, 0, false};
        for (int d = 0; d < 4; d++) {
            int64_t stride = 1;
            for (int c = 3; c > d; c--) stride *= g.bd[c];
            int64_t m = (bi[d] - 1 + g.bd[d]) % g.bd[d], p = (bi[d] + 1) % g.bd[d];
            ent[2 * d + 1] = {R + (m - bi[d]) * stride, 2 * d + 1, g.bd[d] < 2};
            ent[2 * d + 2] = {R + (p - bi[d]) * stride, 2 * d + 2, g.bd[d] < 3};
        }
        std::stable_sort(ent, ent + 9, [](const Ent& a, const Ent& b) { return a.col < b.col || (a.col == b.col && a.s < b.s); });
        for (int a = 0; a < 9; a++) {
            h_bcol[9 * R + a] = ent[a].col;
            mgcr_c128* dst = h_bval + (size_t)(9 * R + a) * bsz;
            int src = -1;
            if (!ent[a].zero) for (int q = prow[R]; q < prow[R + 1]; q++) if (slot[q] == ent[a].s) src = q;
            for (int r = 0; r < ne; r++) for (int c = 0; c < ne; c++) {
                c128 v = src < 0 ? cmake(0., 0.) : val[(size_t)src * bsz + (size_t)c * ne + r];
                dst[(size_t)r * ne + c].re = v.x; dst[(size_t)r * ne + c].im = v.y;
            }
        }
    }
    h_brow[g.nb] = 9 * g.nb;
    return MGCR_OK;
}

extern "C" int mgcr_mg_coarse_op(mgcr_mg* mg, int level, mgcr_op** out) {
    LEVEL_CHECK(mg, level);
    ARG_CHECK(out, "NULL output");
    *out = mg->lv[level].Ac;
    return MGCR_OK;
}

extern "C" int mgcr_mg_restrict(mgcr_ctx* ctx, mgcr_mg* mg, int level, const mgcr_c128* fine, mgcr_c128* coarse) {
    LEVEL_CHECK(mg, level);
    ARG_CHECK(ctx && fine && coarse, "NULL argument");
    return mg_restrict(ctx, mg->lv[level], (const c128*)fine, (c128*)coarse);
}

extern "C" int mgcr_mg_prolong(mgcr_ctx* ctx, mgcr_mg* mg, int level, const mgcr_c128* coarse, mgcr_c128* fine) {
    LEVEL_CHECK(mg, level);
    ARG_CHECK(ctx && fine && coarse, "NULL argument");
    return mg_prolong(ctx, mg->lv[level], (const c128*)coarse, (c128*)fine);
}

extern "C" int mgcr_mg_cycle(mgcr_ctx* ctx, mgcr_mg* mg, int level, const mgcr_c128* b, mgcr_c128* x) {
    LEVEL_CHECK(mg, level);
    ARG_CHECK(ctx && b && x, "NULL argument");
    return mg_cycle(mg, level, (const c128*)b, (c128*)x);
}

extern "C" int mgcr_mg_op_create(mgcr_ctx* ctx, mgcr_mg* mg, mgcr_op** out) {
    ARG_CHECK(ctx && mg && out, "mgcr_mg_op_create: NULL argument");
    MgOp* op = new MgOp();
    op->kind = OP_MG; op->ctx = ctx; op->mg = mg; op->level = 0;
    op->n_local = mg->lv[0].n; op->n_global = mg->lv[0].A->n_global; op->distributed = mg->lv[0].A->distributed;
    *out = op;
    return MGCR_OK;
}

// the distributed CSR entry point lives here until the slab-partitioned hierarchy lands
