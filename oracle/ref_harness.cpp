// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Drives the UNMODIFIED reference (jing2li/MGPreconditionedGCR) headers, included in place from
// /root/reference/src, to (1) emit golden vectors for tests/golden and (2) time the reference's CPU
// solve for bench.py's `--impl reference` / cpu_baseline legs.  Nothing here is linked into or called
// by the product library.  Built by oracle/Makefile into oracle/_ref/ref_oracle (git-ignored).
//
// Rules followed: every std header before the reference headers (the reference
// defines global macros `one`/`zero`), scalar indices written as (long)0, synthetic operators built with
// the raw-CSR ctor from malloc'd arrays (src/Operator.h:64).  `private`/`protected` are opened with the
// usual test-harness define so the hierarchy built by MG::initialise (src/MG.h:131-285) can be dumped.
// Run with MALLOC_PERTURB_=255 so that the reference's uninitialised buffers (src/MG.h:112,126) are
// all-zero bytes and its output is deterministic.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include <omp.h>

#define private public
#define protected public
#include "Fields.h"
#include "GCR.h"
#include "utils.h"
#include "Parse.h"
#include "Operator.h"
#include "MG.h"
#undef private
#undef protected

typedef std::complex<double> cplx;
static std::string g_out;

static void dump(const std::string& name, const void* p, size_t bytes) {
    std::string path = g_out + "/" + name + ".bin";
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    fwrite(p, 1, bytes, f);
    fclose(f);
}
static void dump_field(const std::string& name, const Field<long>& f) {
    dump(name, f.field, sizeof(cplx) * (size_t)f.field_size());
}
static void dump_f64(const std::string& name, const std::vector<double>& v) { dump(name, v.data(), v.size() * 8); }
static void dump_i64(const std::string& name, const std::vector<long>& v) { dump(name, v.data(), v.size() * 8); }

// Wraps an operator and records ||input|| at full precision on every apply.  GCR::solve applies A to p
// once (src/GCR.h:191) and then to r once per iteration (src/GCR.h:242), so the record after the first
// entry IS the residual history ||r_g||, without touching the reference or parsing its %.10e prints.
struct Recorder : public Operator<long> {
    Operator<long>* inner;
    std::vector<double> norms;
    explicit Recorder(Operator<long>* op) : inner(op) { this->dim = op->get_dim(); }
    Field<long> operator()(const Field<long>& f) override { norms.push_back(f.norm()); return (*inner)(f); }
    std::complex<double> val_at(long l) const override { return inner->val_at(l); }
    std::complex<double> val_at(long r, long c) const override { return inner->val_at(r, c); }
};

// hopping matrix H of an n-d nearest-neighbour lattice (unit off-diagonals, Dirichlet), raw CSR ctor.
static Sparse<long>* make_hopping(const std::vector<long>& dims) {
    long V = 1; for (long d : dims) V *= d;
    int nd = (int)dims.size();
    std::vector<long> stride(nd); long s = 1;
    for (int d = nd - 1; d >= 0; d--) { stride[d] = s; s *= dims[d]; }
    long* ROW = (long*)malloc(sizeof(long) * (V + 1));
    long nnz = 0;
    std::vector<long> idx(nd, 0);
    // count
    for (long i = 0; i < V; i++) {
        long rem = i; ROW[i] = nnz;
        for (int d = 0; d < nd; d++) { idx[d] = rem / stride[d]; rem -= idx[d] * stride[d]; }
        for (int d = 0; d < nd; d++) { if (idx[d] > 0) nnz++; if (idx[d] < dims[d] - 1) nnz++; }
    }
    ROW[V] = nnz;
    long* COL = (long*)malloc(sizeof(long) * nnz);
    cplx* VAL = (cplx*)malloc(sizeof(cplx) * nnz);
    long l = 0;
    for (long i = 0; i < V; i++) {
        long rem = i;
        for (int d = 0; d < nd; d++) { idx[d] = rem / stride[d]; rem -= idx[d] * stride[d]; }
        // ascending column order: -stride[0], -stride[1], ..., +stride[nd-1], ..., +stride[0]
        for (int d = 0; d < nd; d++) if (idx[d] > 0) { COL[l] = i - stride[d]; VAL[l] = 1.; l++; }
        for (int d = nd - 1; d >= 0; d--) if (idx[d] < dims[d] - 1) { COL[l] = i + stride[d]; VAL[l] = 1.; l++; }
    }
    return new Sparse<long>(V, V, ROW, COL, VAL);
}

struct SolveOut { std::vector<double> hist; int iters; double seconds; };

static SolveOut run_gcr(Operator<long>* A, GCR_Param<long>* p, const Field<long>& rhs, Field<long>& x) {
    Recorder rec(A);
    GCR<long> gcr(&rec, p);
    auto t0 = std::chrono::steady_clock::now();
    gcr.solve(rhs, x);
    auto t1 = std::chrono::steady_clock::now();
    SolveOut o;
    o.seconds = std::chrono::duration<double>(t1 - t0).count();
    double rn = rhs.norm();
    // norms[0] = ||p0|| = ||rhs||; norms[g] = ||r_g||
    o.hist.push_back(1.0);
    for (size_t i = 1; i < rec.norms.size(); i++) o.hist.push_back(rec.norms[i] / rn);
    o.iters = (int)rec.norms.size() - 1;
    return o;
}

// ---------------------------------------------------------------------------------------------
// golden: everything the parity tests anchor on, from the reference's own code on its own data.
// ---------------------------------------------------------------------------------------------
static void golden_rand() {
    for (int seed : {0, 1, 2, 9, 42}) {
        long d[1] = {1024};
        Field<long> f(d, 1);
        f.init_rand(seed);
        dump_field("rand_seed" + std::to_string(seed), f);
    }
}

static void golden_c1_matrix(Sparse<long>* D) {
    long nrow = D->get_nrow(), nnz = D->get_nnz();
    std::vector<long> row(nrow + 1), col(nnz);
    std::vector<cplx> val(nnz);
    for (long i = 0; i <= nrow; i++) row[i] = D->get_ROW(i);
    for (long l = 0; l < nnz; l++) { col[l] = D->get_COL(l); val[l] = D->val_at(l); }
    dump_i64("c1_row", row); dump_i64("c1_col", col);
    dump("c1_val", val.data(), val.size() * sizeof(cplx));
}

static void golden_c1_apply(Sparse<long>* D, DiracOp<long>* A) {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Field<long> f(dims, 6);
    f.init_rand(1);
    dump_field("c1_f1", f);
    Field<long> y = (*D)(f);
    dump_field("c1_spmv_f1", y);
    Field<long> z = (*A)(f);
    dump_field("c1_dirac_f1", z);
    Field<long> g(dims, 6);
    g.init_rand(5);
    cplx d = f.dot(g);
    std::vector<double> sc = {d.real(), d.imag(), f.squarednorm(), g.norm()};
    dump_f64("c1_dot_f1_f5__n2_f1__norm_f5", sc);
    Field<long> g5 = f.gamma5(4);
    dump_field("c1_gamma5_f1", g5);
}

static void golden_c1_gcr(DiracOp<long>* A) {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    struct Mode { const char* name; int trunc, restart, max_iter; double tol; };
    Mode modes[] = {{"r5", 0, 5, 4000, 1e-13}, {"r2", 0, 2, 4000, 1e-13}, {"t5", 5, 0, 4000, 1e-13},
                    {"r10", 0, 10, 4000, 1e-10}, {"full100", 0, 0, 100, 1e-10}, {"smooth0", 0, 10, 0, 1e-8}};
    for (auto& m : modes) {
        Field<long> rhs(dims, 6); rhs.init_rand(0);
        Field<long> x(dims, 6); x.set_zero();
        GCR_Param<long> p(m.trunc, m.restart, m.max_iter, m.tol, false, nullptr, nullptr);
        SolveOut o = run_gcr(A, &p, rhs, x);
        dump_f64(std::string("c1_gcr_") + m.name + "_hist", o.hist);
        dump_field(std::string("c1_gcr_") + m.name + "_x", x);
        printf("golden gcr %-8s iters=%d final=%.10e |x|=%.12e  %.3fs\n", m.name, o.iters, o.hist.back(), x.norm(), o.seconds);
    }
    // operator() start: x0 = init_rand(2) (src/GCR.h:63-68)
    {
        Field<long> rhs(dims, 6); rhs.init_rand(0);
        GCR_Param<long> p(0, 5, 4000, 1e-13, false, nullptr, nullptr);
        GCR<long> gcr(A, &p);
        Field<long> x = gcr(rhs);
        dump_field("c1_gcr_call_r5_x", x);
    }
    // aliased solve(b,b) as in Arnoldi (src/MG.h:101-104): one round
    {
        Field<long> b(dims, 6); b.init_rand(9);
        GCR_Param<long> p(0, 10, 10, 1e-8, false, nullptr, nullptr);
        Recorder rec(A);
        GCR<long> gcr(&rec, &p);
        gcr.solve(b, b);
        dump_field("c1_gcr_alias_b9", b);
        std::vector<double> it = {(double)(rec.norms.size() - 1)};
        dump_f64("c1_gcr_alias_iters", it);
    }
}

static void dump_hierarchy(const std::string& tag, MG<long>& mg, MG_Param<long>& param, int n_eigen, long sub) {
    int ne = 2 * n_eigen;
    long nb = param.mesh.get_nblocks();
    long bs = param.mesh.get_block_size();
    std::vector<long> bmap(nb * bs);
    for (long b = 0; b < nb; b++) for (long o = 0; o < bs; o++) bmap[b * bs + o] = param.mesh.get_block_map(b)[o];
    dump_i64(tag + "_block_map", bmap);
    // compact prolongator: [block][e][site-in-block][12 dof]  (dof = spinor*3+colour, contiguous in the field)
    std::vector<cplx> P((size_t)nb * ne * bs * 12);
    for (long b = 0; b < nb; b++)
        for (int e = 0; e < ne; e++)
            for (long o = 0; o < bs; o++)
                for (int s = 0; s < 12; s++)
                    P[((b * ne + e) * bs + o) * 12 + s] = mg.prolongator[b][e].val_at(bmap[b * bs + o] * 12 + s);
    dump(tag + "_prolongator", P.data(), P.size() * sizeof(cplx));
    auto* mc = dynamic_cast<HierarchicalSparse<long, int>*>(mg.m_coarse);
    std::vector<long> row(nb + 1), col(mc->ROW[nb]);
    for (long i = 0; i <= nb; i++) row[i] = mc->ROW[i];
    for (long l = 0; l < mc->ROW[nb]; l++) col[l] = mc->COL[l];
    dump_i64(tag + "_coarse_row", row); dump_i64(tag + "_coarse_col", col);
    std::vector<cplx> val((size_t)mc->ROW[nb] * ne * ne);
    for (long l = 0; l < mc->ROW[nb]; l++)
        for (int i = 0; i < ne * ne; i++) val[l * ne * ne + i] = mc->VAL[l]->val_at((int)i);
    dump(tag + "_coarse_val", val.data(), val.size() * sizeof(cplx));
    // component applies
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Field<long> f(dims, 6); f.init_rand(42);
    Field<long> rc = mg.restrict(f);
    dump_field(tag + "_restrict_f42", rc);
    Field<long> pf = mg.expand(rc);
    dump_field(tag + "_expand_restrict_f42", pf);
    Field<long> mcv = (*mg.m_coarse)(rc);
    dump_field(tag + "_coarse_apply", mcv);
    (void)sub;
}

static void golden_c1_mg(DiracOp<long>* A) {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Mesh<long> mesh(dims, 6);
    struct Cfg { const char* tag; long sub; int n_eigen; };
    Cfg cfgs[] = {{"mg_s2_e2", 2, 2}, {"mg_s1_e1", 1, 1}, {"mg_s2_e3", 2, 3}};
    for (auto& c : cfgs) {
        GCR_Param<long> eigen(0, 10, 10, 1e-8, false, nullptr, nullptr);
        GCR_Param<long> coarse(0, 10, 50, 1e-2, false, nullptr, nullptr);
        GCR_Param<long> smooth(0, 10, 0, 1e-8, false, nullptr, nullptr);
        // near-null vectors straight from the reference's Arnoldi (src/MG.h:90-122)
        {
            std::vector<Field<long>> ev(c.n_eigen);
            Arnoldi<long> ar(&eigen, c.n_eigen);
            ar.solve(A, ev.data(), mesh);
            std::vector<cplx> all;
            for (auto& v : ev) all.insert(all.end(), v.field, v.field + v.field_size());
            dump(std::string(c.tag) + "_nearnull", all.data(), all.size() * sizeof(cplx));
        }
        auto sc = new GCR<long>(&coarse);
        auto ss = new GCR<long>(&smooth);
        MG_Param<long> param(mesh, c.sub, c.n_eigen, &eigen, sc, ss, 1, nullptr, nullptr);
        auto t0 = std::chrono::steady_clock::now();
        MG<long> mg(A, &param);
        auto t1 = std::chrono::steady_clock::now();
        printf("golden %s: MG::initialise %.3f s, blocks=%ld\n", c.tag, std::chrono::duration<double>(t1 - t0).count(),
               (long)param.mesh.get_nblocks());
        dump_hierarchy(c.tag, mg, param, c.n_eigen, c.sub);
        delete sc; delete ss;
    }
}

// Algorithm 2 of the report (SemesterProject.pdf p.4) assembled ONLY from the reference's public pieces
// (MG::restrict / MG::expand / GCR::solve / HierarchicalSparse::operator()), because the shipped
// MG::solve is undefined behaviour.  Outer loop: flexible right-preconditioned GCR
// that is line-for-line src/GCR.h:222-288 when no preconditioner is set.
struct Alg2Cycle {
    Operator<long>* A; MG<long>* mg; GCR<long>* smoother; GCR<long>* coarse;
    Field<long> apply(const Field<long>& b) {
        Field<long> x(b.get_mesh()); x.set_zero();
        smoother->solve(b, x);                        // pre-smoothing from x = 0
        Field<long> r = b - (*A)(x);
        Field<long> rc = mg->restrict(r);
        Field<long> xc(rc.get_mesh()); xc.set_zero();
        coarse->solve(rc, xc);                        // zero initial guess (Q2 switch)
        x += mg->expand(xc);
        Field<long> r2 = b - (*A)(x);
        smoother->solve(r2, x);                       // post-smoothing: x += S (b - A x)
        return x;
    }
};

static void golden_c1_mgsolve(DiracOp<long>* A, bool std_conj, int n_eigen, const std::string& tag) {
    long dims[6] = {4, 4, 4, 4, 4, 3};
    Mesh<long> mesh(dims, 6);
    GCR_Param<long> eigen(0, 10, 10, 1e-8, false, nullptr, nullptr);
    GCR_Param<long> coarse(0, 10, 50, 1e-2, false, nullptr, nullptr);
    GCR_Param<long> smooth(0, 10, 0, 1e-8, false, nullptr, nullptr);
    auto sc = new GCR<long>(&coarse);
    auto ss = new GCR<long>(&smooth);
    MG_Param<long> param(mesh, 2, n_eigen, &eigen, sc, ss, 1, nullptr, nullptr);
    MG<long> mg(A, &param);
    Alg2Cycle cyc{A, &mg, ss, sc};
    {   // one preconditioner application
        Field<long> f(dims, 6); f.init_rand(7);
        Field<long> z = cyc.apply(f);
        dump_field(tag + "_cycle_f7", z);
    }
    Field<long> rhs(dims, 6); rhs.init_rand(0);
    Field<long> x(dims, 6); x.set_zero();
    const int restart = 2, max_iter = 200; const double tol = 1e-13;
    Field<long> r(rhs);
    std::vector<Field<long>> ps(restart), Aps(restart);
    Field<long> p = cyc.apply(r);
    Field<long> Ap = (*A)(p);
    ps[0] = p; Aps[0] = Ap;
    std::vector<double> hist = {1.0};
    int iter = 0, g = 0; double rn = rhs.norm();
    do {
        g++; iter++;
        cplx alpha = (std_conj ? Ap.dot(r) : r.dot(Ap)) / Ap.dot(Ap);
        x = x + p * alpha;
        r = r - Ap * alpha;
        Field<long> z = cyc.apply(r);
        Field<long> Az = (*A)(z);
        int lim = std::min(restart, iter);
        Field<long> pc(dims, 6), Apc(dims, 6); pc.set_zero(); Apc.set_zero();
        for (int i = 0; i < lim; i++) {
            cplx beta = (std_conj ? Aps[i].dot(Az) : Az.dot(Aps[i])) / Aps[i].dot(Aps[i]);
            pc = pc - ps[i] * beta; Apc = Apc - Aps[i] * beta;
        }
        p = z + pc; Ap = Az + Apc;
        hist.push_back(r.norm() / rn);
        if (iter % restart == 0) iter = 0;
        Aps[iter % restart] = Ap; ps[iter % restart] = p;
    } while (r.squarednorm() / rhs.squarednorm() > tol * tol && g < max_iter);
    printf("golden %s: flexible MG-GCR iters=%d final=%.6e\n", tag.c_str(), g, hist.back());
    dump_f64(tag + "_hist", hist);
    dump_field(tag + "_x", x);
    delete sc; delete ss;
}

static void golden_synth() {
    // small synthetic hopping operators through the reference API: A = I - kH (DiracOp), x0 = 0
    struct S { const char* tag; std::vector<long> dims; double m2; int restart; };
    S cases[] = {{"lap2d_48", {48, 48}, 0.01, 10}, {"lap3d_12", {12, 12, 12}, 0.01, 10}};
    for (auto& c : cases) {
        Sparse<long>* H = make_hopping(c.dims);
        double k = 1.0 / (2.0 * c.dims.size() + c.m2);
        DiracOp<long> A(H, k);
        long V = H->get_nrow();
        long d1[1] = {V};
        Field<long> rhs(d1, 1); rhs.init_rand(0);
        Field<long> x(d1, 1); x.set_zero();
        GCR_Param<long> p(0, c.restart, 100000, 1e-10, false, nullptr, nullptr);
        SolveOut o = run_gcr(&A, &p, rhs, x);
        printf("golden synth %s: V=%ld iters=%d final=%.6e\n", c.tag, V, o.iters, o.hist.back());
        dump_f64(std::string(c.tag) + "_hist", o.hist);
        dump_field(std::string(c.tag) + "_x", x);
        delete H;
    }
}

static void golden_blocking() {
    struct B { const char* tag; std::vector<long> dims; long sub; };
    B cases[] = {{"bm_4444_s2", {4, 4, 4, 4, 4, 3}, 2}, {"bm_4444_s1", {4, 4, 4, 4, 4, 3}, 1},
                 {"bm_8484_s4", {8, 4, 8, 4, 4, 3}, 4}, {"bm_6666_s3", {6, 6, 6, 6, 4, 3}, 3}};
    bool mask[6] = {true, true, true, true, false, false};
    for (auto& c : cases) {
        Mesh<long> m(c.dims.data(), 6);
        m.blocking(c.sub, mask);
        long nb = m.get_nblocks(), bs = m.get_block_size();
        std::vector<long> bmap(nb * bs);
        for (long b = 0; b < nb; b++) for (long o = 0; o < bs; o++) bmap[b * bs + o] = m.get_block_map(b)[o];
        dump_i64(c.tag, bmap);
    }
}

// ---------------------------------------------------------------------------------------------
// bench: time the reference's GCR on a synthetic hopping operator (cpu_baseline / --impl reference)
// ---------------------------------------------------------------------------------------------
static int bench(int argc, char** argv) {
    // ref_oracle bench <nd> <d0> .. <m2> <restart> <iters> <nearnull(0)>
    int nd = atoi(argv[2]);
    std::vector<long> dims;
    for (int i = 0; i < nd; i++) dims.push_back(atol(argv[3 + i]));
    double m2 = atof(argv[3 + nd]);
    int restart = atoi(argv[4 + nd]);
    int iters = atoi(argv[5 + nd]);
    auto t0 = std::chrono::steady_clock::now();
    Sparse<long>* H = make_hopping(dims);
    double k = 1.0 / (2.0 * nd + m2);
    DiracOp<long> A(H, k);
    long V = H->get_nrow();
    long d1[1] = {V};
    Field<long> rhs(d1, 1); rhs.init_rand(0);
    Field<long> x(d1, 1); x.set_zero();
    auto t1 = std::chrono::steady_clock::now();
    // one SpMV timing
    auto s0 = std::chrono::steady_clock::now();
    Field<long> y = (*H)(rhs);
    auto s1 = std::chrono::steady_clock::now();
    GCR_Param<long> p(0, restart, iters, 1e-10, false, nullptr, nullptr);
    SolveOut o = run_gcr(&A, &p, rhs, x);
    printf("{\"V\": %ld, \"nnz\": %ld, \"iters\": %d, \"solve_seconds\": %.6f, \"seconds_per_iter\": %.6f, "
           "\"spmv_seconds\": %.6f, \"setup_seconds\": %.3f, \"final_rel_res\": %.10e, \"threads\": 1}\n",
           V, (long)H->get_nnz(), o.iters, o.seconds, o.seconds / std::max(1, o.iters),
           std::chrono::duration<double>(s1 - s0).count(), std::chrono::duration<double>(t1 - t0).count(), o.hist.back());
    delete H;
    return 0;
}

// gcr-file: run the reference GCR on an operator read from raw binary files (used by CPU tests to
// cross-check the C restatement on arbitrary inputs).  dir holds row.bin col.bin val.bin rhs.bin x0.bin
static int gcr_file(int argc, char** argv) {
    std::string dir = argv[2];
    long nrow = atol(argv[3]); long nnz = atol(argv[4]);
    double kre = atof(argv[5]), kim = atof(argv[6]);
    int trunc = atoi(argv[7]), restart = atoi(argv[8]), max_iter = atoi(argv[9]); double tol = atof(argv[10]);
    auto rd = [&](const char* n, void* p, size_t b) {
        FILE* f = fopen((dir + "/" + n).c_str(), "rb"); if (!f) { fprintf(stderr, "missing %s\n", n); exit(2); }
        size_t got = fread(p, 1, b, f); fclose(f); if (got != b) { fprintf(stderr, "short %s\n", n); exit(2); } };
    long* ROW = (long*)malloc(8 * (nrow + 1)); long* COL = (long*)malloc(8 * nnz); cplx* VAL = (cplx*)malloc(16 * nnz);
    rd("row.bin", ROW, 8 * (nrow + 1)); rd("col.bin", COL, 8 * nnz); rd("val.bin", VAL, 16 * nnz);
    Sparse<long> D(nrow, nrow, ROW, COL, VAL);
    long d1[1] = {nrow};
    Field<long> rhs(d1, 1), x(d1, 1);
    rd("rhs.bin", rhs.field, 16 * nrow); rd("x0.bin", x.field, 16 * nrow);
    GCR_Param<long> p(trunc, restart, max_iter, tol, false, nullptr, nullptr);
    g_out = dir;
    Field<long> x0(d1, 1);
    x0 = x;
    DiracOp<long> Ad(&D, cplx(kre, kim));
    Operator<long>* A = (kre == 0. && kim == 0.) ? (Operator<long>*)&D : (Operator<long>*)&Ad;
    SolveOut o = run_gcr(A, &p, rhs, x);
    dump_f64("hist", o.hist);
    dump_field("x", x);
    // the same solve once more WITHOUT the recording wrapper: the reference's own wall-clock for this problem
    Field<long> xt(d1, 1);
    xt = x0;
    GCR<long> plain(A, &p);
    auto t0 = std::chrono::steady_clock::now();
    plain.solve(rhs, xt);
    auto t1 = std::chrono::steady_clock::now();
    printf("{\"V\": %ld, \"nnz\": %ld, \"iters\": %d, \"solve_seconds\": %.6f, \"final_rel_res\": %.10e, \"threads\": 1}\n", nrow, nnz, o.iters,
           std::chrono::duration<double>(t1 - t0).count(), o.hist.back());
    return 0;
}

// gcr-left: the reference GCR with a LEFT preconditioner (src/GCR.h:201-204, 245-247).  L is a caller-defined subclass of the
// reference's open Operator interface: a real diagonal read from diag.bin (Jacobi-like).  Same files as gcr-file.
struct DiagOp : public Operator<long> {
    std::vector<double> d;
    Field<long> operator()(const Field<long>& f) override {
        Field<long> out(f);
        for (long i = 0; i < (long)d.size(); i++) out.field[i] = d[i] * f.field[i];
        return out;
    }
    std::complex<double> val_at(long, long) const override { return 0.; }
    std::complex<double> val_at(long) const override { return 0.; }
    void set_dim(long n) { this->dim = n; }
};
static int gcr_left(int argc, char** argv) {
    std::string dir = argv[2];
    long nrow = atol(argv[3]); long nnz = atol(argv[4]);
    double kre = atof(argv[5]), kim = atof(argv[6]);
    int trunc = atoi(argv[7]), restart = atoi(argv[8]), max_iter = atoi(argv[9]); double tol = atof(argv[10]);
    auto rd = [&](const char* n, void* p, size_t b) {
        FILE* f = fopen((dir + "/" + n).c_str(), "rb"); if (!f) { fprintf(stderr, "missing %s\n", n); exit(2); }
        size_t got = fread(p, 1, b, f); fclose(f); if (got != b) { fprintf(stderr, "short %s\n", n); exit(2); } };
    long* ROW = (long*)malloc(8 * (nrow + 1)); long* COL = (long*)malloc(8 * nnz); cplx* VAL = (cplx*)malloc(16 * nnz);
    rd("row.bin", ROW, 8 * (nrow + 1)); rd("col.bin", COL, 8 * nnz); rd("val.bin", VAL, 16 * nnz);
    Sparse<long> D(nrow, nrow, ROW, COL, VAL);
    DiracOp<long> A(&D, cplx(kre, kim));
    DiagOp L;
    L.d.resize(nrow);
    rd("diag.bin", L.d.data(), 8 * nrow);
    L.set_dim(nrow);
    long d1[1] = {nrow};
    Field<long> rhs(d1, 1), x(d1, 1);
    rd("rhs.bin", rhs.field, 16 * nrow); rd("x0.bin", x.field, 16 * nrow);
    GCR_Param<long> p(trunc, restart, max_iter, tol, false, &L, nullptr);
    g_out = dir;
    // residual history as the reference prints it: sqrt(r.squarednorm()) / rhs.norm() with r the (left-preconditioned) recurrence
    // residual -- recorded by solving with verbose output captured would need stdout parsing; the recurrence is re-derived from x
    // instead: the harness records ||input of A|| like run_gcr does (the input of A in iteration g is r_g)
    SolveOut o = run_gcr(&A, &p, rhs, x);
    o.hist[0] = L(rhs).norm() / rhs.norm();   // the step-0 print of src/GCR.h:214: r = L(rhs) by then
    dump_f64("hist", o.hist);
    dump_field("x", x);
    printf("{\"iters\": %d, \"final_rel_res\": %.10e}\n", o.iters, o.hist.back());
    return 0;
}

// parse: the reference's own I/O pair on a MatrixMarket file: parse_data(<file>) writes ../../data/sample_matrix/parsed.txt
// (src/Parse.cpp:39, relative to the working directory), read_data("parsed.txt") reads it back; the arrays are dumped to <outdir>.
static int parse_mode(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_oracle parse <file.mtx> <outdir>\n"); return 1; }
    parse_data(argv[2]);
    Sparse<long> m = read_data("parsed.txt");
    g_out = argv[3];
    long nrow = m.get_nrow(), nnz = m.get_nnz();
    std::vector<long> row(nrow + 1), col(nnz);
    std::vector<cplx> val(nnz);
    for (long r = 0; r <= nrow; r++) row[r] = m.get_ROW(r);
    for (long l = 0; l < nnz; l++) { col[l] = m.get_COL(l); val[l] = m.val_at(l); }
    dump_i64("row", row);
    dump_i64("col", col);
    FILE* f = fopen((std::string(argv[3]) + "/val.bin").c_str(), "wb"); fwrite(val.data(), 16, val.size(), f); fclose(f);
    printf("PARSED %ld %ld %ld\n", nrow, (long)m.get_dim(), nnz);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_oracle golden <outdir> | bench ... | gcr-file ...\n"); return 1; }
    std::string cmd = argv[1];
    if (cmd == "bench") return bench(argc, argv);
    if (cmd == "gcr-file") return gcr_file(argc, argv);
    if (cmd == "gcr-left") return gcr_left(argc, argv);
    if (cmd == "parse") return parse_mode(argc, argv);
    if (cmd == "golden") {
        g_out = argv[2];
        golden_rand();
        golden_blocking();
        golden_synth();
        auto D = new Sparse<long>(read_data("4x4parsed.txt"));   // needs cwd two levels below data/sample_matrix
        double k = 0.05 + 8 * ((0.17865 - 0.05) / 10.);           // src/main.cpp:845-847
        auto A = new DiracOp<long>(D, k);
        golden_c1_matrix(D);
        golden_c1_apply(D, A);
        golden_c1_gcr(A);
        golden_c1_mg(A);
        golden_c1_mgsolve(A, false, 3, "mgsolve_refconj_e3");
        golden_c1_mgsolve(A, true, 3, "mgsolve_stdconj_e3");
        if (argc > 3) { golden_c1_mgsolve(A, true, 10, "mgsolve_stdconj_e10"); }
        delete A; delete D;
        return 0;
    }
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 1;
}
