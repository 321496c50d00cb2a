// omg_hier.cuh — host-side hierarchy object (levels, vectors, cached CUDA graphs).
#pragma once
#include <map>
#include <tuple>

#include "omg_common.cuh"

struct Globals {
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // halo / copy stream
    // distributed state (omg_dist.cu)
    int rank = 0, nranks = 1;
    void *nccl_comm = nullptr;
};
extern Globals g;

struct Level {
    // geometry
    int n = 0;                 // global rows
    int row0 = 0;              // first owned global row (0 on one GPU)
    int nloc = 0;              // owned rows
    int pad = 0;               // zero pad / halo capacity on each side (multiple of 16)
    bool slab = false;         // rows partitioned across ranks (else replicated / single GPU)
    int halo = 0;              // elements exchanged with each neighbour before an operator application
    int piece_row0 = 0;        // first coarse row this rank's fine slab restricts to
    int piece_n = 0;           // number of coarse rows it produces
    int exc_s0 = 0, exc_s1 = 0;     // exception slots whose row is owned
    int crow_t0 = 0, crow_t1 = 0;   // entries of exc_crows whose coarse row this rank produces
    int ndim = 0;
    int shape[3] = {1, 1, 1};  // level shape, openmg/operators.py:131,136 (C order, as given)
    ColourRule colour{};

    // operator
    int kind = OMG_KIND_CSR;
    BandOp band{};
    int64_t nexc = 0;
    unsigned *exc_mask = nullptr;
    int *exc_wpre = nullptr, *exc_ptr = nullptr, *exc_col = nullptr;
    double *exc_val = nullptr, *exc_diag = nullptr;
    int *exc_rows = nullptr;     // row index per exception slot
    int *exc_crows = nullptr;    // coarse rows whose aggregate contains an exception row (regular R)
    int nexc_crows = 0;
    bool exc_diag_uniform = true; // every exception row has a_ii == band.diag
    int exc_reach = 0;            // max |column - row| over the exception rows
    bool classed = false;         // all exception rows follow the 9 positional stencil classes below
    ClsTab cls{};
    bool classed2 = false;        // 2-D analogue: rows of the first / last grid column carry three correction taps each
    double c2l[3] = {0, 0, 0};    // first column: deltas of the taps -1, -(N+1), +(N-1)
    double c2r[3] = {0, 0, 0};    // last column:  deltas of the taps +1, +(N+1), -(N-1)
    // full CSR (sorted columns, global indices == local on one GPU); may be absent for a band level 0
    int64_t nnz = 0;
    int *ptr = nullptr, *col = nullptr;
    double *val = nullptr, *diag = nullptr;

    // sliced-ELLPACK copy of the CSR (CSR-kind levels: the operator the kernels read)
    int *sell_off = nullptr, *sell_rlen = nullptr, *sell_col = nullptr;
    double *sell_val = nullptr;
    int64_t sell_entries = 0;

    // restriction to level+1
    bool hasR = false;
    bool regular = false;
    RegR reg{};
    int Rk = 0;                // entries per R row (after dedupe)
    int Roffs[8] = {0};
    double Rw = 0.0;
    int nc = 0;                // rows of R (= n of next level)
    int *Rptr = nullptr, *Rcol = nullptr, *RTptr = nullptr, *RTcol = nullptr;   // explicit pattern (non-regular)
    int *Rcc = nullptr;        // explicit: first column per coarse row

    // vectors [pad | nloc | pad]
    double *xa_base = nullptr, *xb_base = nullptr, *b_base = nullptr, *r_base = nullptr;
    double *xa = nullptr, *xb = nullptr, *b = nullptr, *r = nullptr;

    ExcOp exc_op() const { return ExcOp{exc_mask, exc_wpre, exc_ptr, exc_col, exc_val, exc_diag}; }
    CsrOp csr_op() const { return CsrOp{ptr, col, val, diag}; }
    SellOp sell_op() const { return SellOp{sell_off, sell_rlen, sell_col, sell_val, diag}; }
    // matrix bytes one application of a CSR-kind operator moves (12 B per stored slot + row length + a_ii);
    // band levels carry their operator in kernel parameters: 0
    double matrix_bytes() const {
        if (kind != OMG_KIND_CSR) return 0.0;
        return sell_val ? 12.0 * (double)sell_entries + 12.0 * n : 12.0 * (double)nnz + 16.0 * n;
    }
};

struct CycleCfg {
    int pre, post, smoother, with_norm;
    double omega;
    int cur0;   // which level-0 buffer holds the iterate on entry
    bool operator<(const CycleCfg &o) const {
        return std::tie(pre, post, smoother, with_norm, omega, cur0) <
               std::tie(o.pre, o.post, o.smoother, o.with_norm, o.omega, o.cur0);
    }
};

struct CachedGraph {
    cudaGraphExec_t exec = nullptr;
    int cur0_after = 0;
    int64_t launches = 0;
};

// Device-side stop rule of mgSolve (openmg/__init__.py:118-138): the state lives in device memory, every cycle of a
// thresholded solve is wrapped in an IF node of its CUDA graph, so the host enqueues cycles without reading anything
// back in between.
struct StopState {
    double threshold;     // <= 0: no residual test
    double norm;          // ||b - A x||_2 after the last executed cycle
    int max_cycles;       // <= 0: no cycle cap
    int cycle;            // cycles executed
    int done;
    int hist_cap;
};

struct ProfRec {
    const char *name;
    int level;
    double bytes;       // algorithmic bytes of this launch
    cudaEvent_t e0, e1;
};

// Peer-memory halo exchange state (omg_dist.cu): neighbours' level buffers and flag words mapped
// through CUDA IPC, so halos move with copy-engine peer copies ordered by device-side flags.
struct PeerState {
    bool enabled = false;
    unsigned long long *flags = nullptr;        // local: 4 words per slot {ready<-dn, ready<-up, done<-dn, done<-up}
    unsigned long long *epochs = nullptr;       // local: one exchange counter per slot
    unsigned long long *flags_dn = nullptr, *flags_up = nullptr;   // the neighbours' flag arrays (mapped)
    std::vector<double *> base_dn, base_up;     // per slot: the neighbours' buffer base (mapped), slot = 2*level + (xb?1:0)
    std::vector<void *> opened;                 // everything cudaIpcOpenMemHandle returned
    // fused halo pull (omg_stencil.cu k_st3 on slab levels): 8 words per slot
    //   {kernels done, ready<-dn, ready<-up, arrivals, pulled<-dn, pulled<-up, -, -}
    unsigned long long *pull = nullptr;                           // local
    unsigned long long *pull_dn = nullptr, *pull_up = nullptr;    // the neighbours' arrays (mapped)
    int *pull_timeout = nullptr;                // set by a kernel that gave up waiting for a neighbour
    bool pull_ok = false;
};

// what a stencil kernel needs to read the halos of one vector of a slab level straight from the neighbours
struct HaloPull {
    unsigned long long *mine, *peer_dn, *peer_up;   // this slot's 8 words here and on the neighbours (nullptr: no neighbour)
    const double *v_dn, *v_up;                      // the neighbours' copies of the vector (their owned row 0)
    int *timeout;
};

struct omg_hierarchy {
    int flags = 0;
    PeerState peer;
    bool profiling = false;          // per-kernel CUDA-event timing (omg_profile_cycle)
    bool halo_pending = false;       // a halo exchange is in flight on the comm stream
    std::vector<ProfRec> prof;
    int nlev = 0;
    int first_replicated = 0;        // levels [0, first_replicated) are row slabs across ranks
    std::vector<Level> lv;
    std::vector<void *> allocs;      // everything cudaMalloc'ed for this hierarchy
    // coarse solve
    double *Ainv = nullptr;
    int ncoarse = 0;
    // norm reduction scratch
    double *partial = nullptr;
    int npartial = 0;
    double *norm2_dev = nullptr;     // device scalar: sum of squares
    double *norm2_host = nullptr;    // pinned
    // state
    int cur0 = 0;                    // level-0 iterate lives in xa (0) or xb (1)
    // halo exchanges asked for but not yet issued: the consuming stencil kernel may pull them itself (fused), any
    // other consumer makes dist_halo_wait issue them on the comm stream
    std::vector<std::pair<Level *, double *>> halo_req;
    std::map<CycleCfg, CachedGraph> graphs;
    std::map<CycleCfg, CachedGraph> gated;      // the same cycles behind the device-side stop test
    StopState *stop_dev = nullptr, *stop_host = nullptr;     // device / pinned
    double *hist_dev = nullptr;
    int hist_dev_cap = 0;
    int64_t host_syncs = 0;          // stream synchronisations issued by the last omg_solve (diagnostic)
    bool no_gate = false;            // a gated graph could not be instantiated: stop test stays on the host
    int64_t launches = 0;            // kernels launched since last reset
    // timings
    double t_upload_ms = 0, t_galerkin_ms = 0, t_coarse_ms = 0;
    double coarse_defect = 0;          // max |A_L * Ainv - I| of the coarse factor that was kept
};

// allocation helpers (omg_setup.cu)
int h_alloc(omg_hierarchy *h, void **p, size_t bytes, bool zero);
template <class T>
static inline int h_alloc_t(omg_hierarchy *h, T **p, size_t count, bool zero = false) {
    return h_alloc(h, (void **)p, count * sizeof(T), zero);
}
void h_free(omg_hierarchy *h, void *p);

int exclusive_scan_i32(const int *in, int *out, int n, int *total_host, cudaStream_t st);

// omg_setup.cu
int setup_levels(omg_hierarchy *h, int ndim, const int64_t *shape, int coarsestLevel, int minSize);
int build_hierarchy(omg_hierarchy *h);
int materialize_level_csr(omg_hierarchy *h, const Level &L, int **ptr, int **col, double **val, int64_t *nnz);

// omg_cycle.cu
int run_cycle(omg_hierarchy *h, const CycleCfg &cfg);

// omg_dist.cu
int dist_halo_exchange(omg_hierarchy *h, Level &L, double *v);
int dist_halo_wait(omg_hierarchy *h);
int dist_halo_flush(omg_hierarchy *h);
bool dist_pull_params(omg_hierarchy *h, Level &L, const double *v, HaloPull *out);
int dist_pull_timed_out(omg_hierarchy *h, bool *timed_out);
int dist_peer_setup(omg_hierarchy *h);
void dist_peer_teardown(omg_hierarchy *h);
int dist_allgather(omg_hierarchy *h, const double *piece, double *full, size_t count);
int dist_allreduce_sum(omg_hierarchy *h, double *v, size_t count);
void dist_finalize();
extern "C" int omg_partition(int nlevels, const int64_t *level_lead, const int64_t *level_rows,
                             const int32_t *level_regular, int nranks, int rank, int64_t agglomerate_below,
                             int32_t *first_replicated, int64_t *row0, int64_t *nloc);

// RAII event pair around one launch when h->profiling
struct ProfScope {
    omg_hierarchy *h;
    int idx;
    cudaStream_t st;
    ProfScope(omg_hierarchy *h_, const char *name, int level, double bytes, cudaStream_t stream = nullptr)
        : h(h_), idx(-1), st(stream ? stream : g.stream) {
        if (!h->profiling) return;
        ProfRec r{name, level, bytes, nullptr, nullptr};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        h->prof.push_back(r);
        idx = (int)h->prof.size() - 1;
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(h->prof[idx].e1, st);
    }
};
