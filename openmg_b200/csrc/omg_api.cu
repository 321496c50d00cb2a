// omg_api.cu — the extern "C" boundary (include/omg_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <ctype.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "omg_hier.cuh"

Globals g;

static thread_local OmgError tls_err{0, ""};

int omg_set_error(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    tls_err.code = code;
    tls_err.msg = buf;
    return code;
}

// implemented in omg_setup.cu / omg_cycle.cu
int level0_from_csr(omg_hierarchy *h);
int device_restriction_csr(int ndim, const int64_t *shape, int64_t *n_rows, int64_t *nnz, int **dptr, int **dcol,
                           double **dval);
int launch_matvec(omg_hierarchy *h, Level &L, double *x, double *y);
int launch_residual(omg_hierarchy *h, Level &L, double *x, const double *b, double *r);
int launch_resnorm2(omg_hierarchy *h, Level &L, double *x, const double *b, int slot);
double *launch_smooth(omg_hierarchy *h, Level &L, int smoother, double omega, int sweeps, double *cur,
                      const double *b);
int launch_residual_restrict(omg_hierarchy *h, int l, double *x, const double *b, double *rc);
int launch_smooth_residual_restrict(omg_hierarchy *h, int l, int smoother, double omega, int sweeps, double **cur,
                                    const double *b, double *rc);
int launch_prolong_correct(omg_hierarchy *h, int l, const double *e, const double *xi, double *xo);
double *launch_prolong_correct_smooth(omg_hierarchy *h, int l, int smoother, double omega, int sweeps, double *cur,
                                      double *e, const double *b, bool cur_halo_valid);
int launch_coarse_solve(omg_hierarchy *h, const double *b, double *x);
double *cycle_from_level(omg_hierarchy *h, int l, const CycleCfg &cfg, double *cur);

extern "C" {

const char *omg_last_error(void) { return tls_err.msg.c_str(); }

int omg_init(int device) {
    if (g.inited) return OMG_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return omg_set_error(OMG_ENODEV, "no CUDA device available (%s); libomg_b200 has no CPU fallback",
                             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) {
        const char *lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % count : 0;
    }
    if (device >= count) return omg_set_error(OMG_ENODEV, "device %d requested but only %d present", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (p.major != 10)
        return omg_set_error(OMG_ENODEV, "device %d is sm_%d%d; libomg_b200 is built for sm_100a (B200) only",
                             device, p.major, p.minor);
    g.device = device;
    g.sm_count = p.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    {
        int lo_pri = 0, hi_pri = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CUDA_TRY(cudaStreamCreateWithPriority(&g.stream2, cudaStreamNonBlocking, hi_pri));   // comm stream
    }
    g.inited = true;
    return OMG_OK;
}

void omg_finalize(void) {
    if (!g.inited) return;
    dist_finalize();
    cudaStreamDestroy(g.stream);
    cudaStreamDestroy(g.stream2);
    g = Globals();
}

int omg_device_info(char *buf, int buflen, int *sm_count, int64_t *mem_bytes) {
    if (!g.inited) return omg_set_error(OMG_ENODEV, "omg_init() has not succeeded");
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, g.device));
    if (buf && buflen > 0) snprintf(buf, buflen, "%s sm_%d%d %d SMs", p.name, p.major, p.minor, p.multiProcessorCount);
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (mem_bytes) *mem_bytes = (int64_t)p.totalGlobalMem;
    return OMG_OK;
}

// NUMA placement of pinned host buffers.  cudaMallocHost pins the pages where the calling thread's memory policy
// puts them — by default all on the node the process happens to run on, so with one process per GPU the eight
// ranks' host<->device copies all hammer the same memory controllers.  Place the buffer on the NUMA node the GPU
// hangs off (/sys/bus/pci/devices/<bdf>/numa_node) when that is known, else interleave it over all nodes.
// Raw syscalls (no libnuma in the image); any failure leaves the default policy in place.
static int gpu_numa_node() {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, g.device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bdf; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
static int numa_node_count() {
    int n = 0;
    for (; n < 64; ++n) {
        char path[64];
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d", n);
        if (access(path, F_OK) != 0) break;
    }
    return n;
}

int omg_host_alloc(void **ptr, int64_t bytes) {
    if (!g.inited) return omg_set_error(OMG_ENODEV, "omg_init() has not succeeded");
    const int MPOL_DEFAULT_ = 0, MPOL_PREFERRED_ = 1, MPOL_INTERLEAVE_ = 3;
    bool policy_set = false;
    if (!getenv("OMG_NO_NUMA")) {
        int nodes = numa_node_count(), node = gpu_numa_node();
        if (nodes > 1) {
            unsigned long mask = (node >= 0 && node < nodes) ? (1ul << node) : ((1ul << nodes) - 1ul);
            int mode = (node >= 0 && node < nodes) ? MPOL_PREFERRED_ : MPOL_INTERLEAVE_;
            policy_set = syscall(SYS_set_mempolicy, mode, &mask, (unsigned long)(8 * sizeof mask)) == 0;
        }
    }
    cudaError_t e = cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 16));
    if (policy_set) syscall(SYS_set_mempolicy, MPOL_DEFAULT_, nullptr, 0ul);
    if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "cudaMallocHost failed: %s", cudaGetErrorString(e));
    return OMG_OK;
}
int omg_host_free(void *ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return OMG_OK;
}

// ------------------------------------------------------------------ create / destroy

void omg_hierarchy_destroy(omg_hierarchy *h) {
    if (!h) return;
    if (g.inited) cudaStreamSynchronize(g.stream);
    if (g.inited) cudaStreamSynchronize(g.stream2);
    dist_peer_teardown(h);
    for (auto &kv : h->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (auto &kv : h->gated)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (h->stop_host) cudaFreeHost(h->stop_host);
    for (void *p : h->allocs) cudaFree(p);
    if (h->norm2_host) cudaFreeHost(h->norm2_host);
    delete h;
}

int omg_hierarchy_create_csr(omg_hierarchy **out, int ndim, const int64_t *shape, int coarsestLevel, int minSize,
                             int64_t n, const int32_t *indptr, const int32_t *indices, const double *data,
                             int flags) {
    if (!g.inited) return omg_set_error(OMG_ENODEV, "omg_init() has not succeeded");
    if (!out || !shape || !indptr || n <= 0 || n >= (1ll << 31))
        return omg_set_error(OMG_EINVAL, "bad arguments to omg_hierarchy_create_csr");
    omg_hierarchy *h = new omg_hierarchy();
    h->flags = flags;
    h->lv.resize(1);
    Level &L = h->lv[0];
    L.n = (int)n;
    L.nloc = (int)n;
    int64_t nnz = indptr[n];
    L.nnz = nnz;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, g.stream);
    int rc = h_alloc_t(h, &L.ptr, (size_t)n + 1);
    if (rc == OMG_OK) rc = h_alloc_t(h, &L.col, (size_t)std::max<int64_t>(nnz, 1));
    if (rc == OMG_OK) rc = h_alloc_t(h, &L.val, (size_t)std::max<int64_t>(nnz, 1));
    if (rc == OMG_OK) {
        cudaMemcpyAsync(L.ptr, indptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, g.stream);
        cudaMemcpyAsync(L.col, indices, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, g.stream);
        cudaMemcpyAsync(L.val, data, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, g.stream);
        cudaEventRecord(e1, g.stream);
        cudaError_t e = cudaStreamSynchronize(g.stream);
        if (e != cudaSuccess) rc = omg_set_error(OMG_ECUDA, "upload of A failed: %s", cudaGetErrorString(e));
    }
    if (rc == OMG_OK) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        h->t_upload_ms = ms;
        // validate the depth rule BEFORE the expensive work, like mgSolve does (R list first)
        std::vector<Level> keep = h->lv;
        rc = setup_levels(h, ndim, shape, coarsestLevel, minSize);
        if (rc == OMG_OK) {
            // setup_levels resized lv; restore level-0 operator fields it does not own
            Level &Z = h->lv[0];
            Z.ptr = keep[0].ptr;
            Z.col = keep[0].col;
            Z.val = keep[0].val;
            Z.nnz = keep[0].nnz;
            rc = level0_from_csr(h);
        }
        if (rc == OMG_OK) rc = build_hierarchy(h);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != OMG_OK) {
        omg_hierarchy_destroy(h);
        return rc;
    }
    *out = h;
    return OMG_OK;
}

int omg_operator_create_csr(omg_hierarchy **out, int64_t n, const int32_t *indptr, const int32_t *indices,
                            const double *data, int flags) {
    if (!g.inited) return omg_set_error(OMG_ENODEV, "omg_init() has not succeeded");
    if (!out || !indptr || n <= 0 || n >= (1ll << 31))
        return omg_set_error(OMG_EINVAL, "bad arguments to omg_operator_create_csr");
    omg_hierarchy *h = new omg_hierarchy();
    h->flags = flags;
    h->nlev = 1;
    h->lv.resize(1);
    Level &L = h->lv[0];
    L.n = L.nloc = (int)n;
    L.nnz = indptr[n];
    L.ndim = 1;
    L.shape[0] = (int)n;
    L.colour.flat = 1;
    L.colour.alpha = 1;
    L.colour.s1 = 1;
    L.colour.s2 = (int)n;
    int rc = h_alloc_t(h, &L.ptr, (size_t)n + 1);
    if (rc == OMG_OK) rc = h_alloc_t(h, &L.col, (size_t)std::max<int64_t>(L.nnz, 1));
    if (rc == OMG_OK) rc = h_alloc_t(h, &L.val, (size_t)std::max<int64_t>(L.nnz, 1));
    if (rc == OMG_OK) {
        cudaMemcpyAsync(L.ptr, indptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, g.stream);
        cudaMemcpyAsync(L.col, indices, sizeof(int) * (size_t)L.nnz, cudaMemcpyHostToDevice, g.stream);
        cudaMemcpyAsync(L.val, data, sizeof(double) * (size_t)L.nnz, cudaMemcpyHostToDevice, g.stream);
        cudaError_t e = cudaStreamSynchronize(g.stream);
        if (e != cudaSuccess) rc = omg_set_error(OMG_ECUDA, "upload of A failed: %s", cudaGetErrorString(e));
    }
    if (rc == OMG_OK) rc = level0_from_csr(h);
    if (rc == OMG_OK) rc = build_hierarchy(h);
    if (rc != OMG_OK) {
        omg_hierarchy_destroy(h);
        return rc;
    }
    *out = h;
    return OMG_OK;
}

int omg_hierarchy_create_band(omg_hierarchy **out, int ndim, const int64_t *shape, int coarsestLevel, int minSize,
                              int64_t n, double diag, int nband, const int64_t *offsets, const double *coeffs,
                              int flags) {
    if (!g.inited) return omg_set_error(OMG_ENODEV, "omg_init() has not succeeded");
    if (!out || !shape || n <= 0 || n >= (1ll << 31) || nband < 0 || 2 * nband > OMG_MAXBAND || diag == 0.0)
        return omg_set_error(OMG_EINVAL, "bad arguments to omg_hierarchy_create_band");
    omg_hierarchy *h = new omg_hierarchy();
    h->flags = flags;
    h->lv.resize(1);
    h->lv[0].n = (int)n;
    h->lv[0].nloc = (int)n;
    int rc = setup_levels(h, ndim, shape, coarsestLevel, minSize);
    if (rc == OMG_OK) {
        Level &L = h->lv[0];
        std::vector<std::pair<int64_t, double>> taps;
        for (int k = 0; k < nband; ++k) {
            if (offsets[k] <= 0) {
                rc = omg_set_error(OMG_EINVAL, "band offsets must be > 0");
                break;
            }
            if (offsets[k] >= n) continue;   // diagonal lies outside the matrix
            bool merged = false;
            for (auto &t : taps)
                if (t.first == offsets[k]) {
                    t.second += coeffs[k];
                    merged = true;
                }
            if (!merged) taps.push_back({offsets[k], coeffs[k]});
        }
        std::sort(taps.begin(), taps.end());
        L.band.nb = 0;
        L.band.diag = diag;
        for (int k = (int)taps.size() - 1; k >= 0; --k) {
            L.band.off[L.band.nb] = (int)-taps[k].first;
            L.band.coef[L.band.nb++] = taps[k].second;
        }
        for (size_t k = 0; k < taps.size(); ++k) {
            L.band.off[L.band.nb] = (int)taps[k].first;
            L.band.coef[L.band.nb++] = taps[k].second;
        }
        L.kind = OMG_KIND_BAND;
        L.nexc = 0;
        int64_t nnz = n;
        for (auto &t : taps) nnz += 2 * (n - t.first);
        L.nnz = nnz;
        if (flags & OMG_FLAG_FORCE_CSR) {
            int64_t nz = 0;
            rc = materialize_level_csr(h, L, &L.ptr, &L.col, &L.val, &nz);
            if (rc == OMG_OK) {
                h->allocs.push_back(L.ptr);
                h->allocs.push_back(L.col);
                h->allocs.push_back(L.val);
                rc = level0_from_csr(h);
            }
        }
    }
    if (rc == OMG_OK) rc = build_hierarchy(h);
    if (rc != OMG_OK) {
        omg_hierarchy_destroy(h);
        return rc;
    }
    *out = h;
    return OMG_OK;
}

// ------------------------------------------------------------------ introspection / export

#define CHECK_H(h)                                                            \
    do {                                                                      \
        if (!(h)) return omg_set_error(OMG_EINVAL, "null hierarchy handle");  \
    } while (0)
#define CHECK_LEVEL(h, l)                                                                          \
    do {                                                                                           \
        CHECK_H(h);                                                                                \
        if ((l) < 0 || (l) >= (h)->nlev) return omg_set_error(OMG_EINVAL, "level %d out of range", (l)); \
    } while (0)

int omg_level_count(const omg_hierarchy *h, int *nlevels) {
    CHECK_H(h);
    *nlevels = h->nlev;
    return OMG_OK;
}

int omg_level_info(const omg_hierarchy *h, int level, int64_t *n, int64_t *nnzA, int64_t *nnzR, int *kind,
                   int64_t *nexc) {
    CHECK_LEVEL(h, level);
    const Level &L = h->lv[level];
    if (n) *n = L.n;
    if (nnzA) *nnzA = L.nnz;
    if (nnzR) *nnzR = L.hasR ? (int64_t)L.nc * L.Rk : 0;
    if (kind) *kind = L.kind;
    if (nexc) *nexc = L.nexc;
    return OMG_OK;
}

int omg_level_partition(const omg_hierarchy *h, int level, int64_t *row0, int64_t *nloc, int *slab) {
    CHECK_LEVEL(h, level);
    const Level &L = h->lv[level];
    if (row0) *row0 = L.row0;
    if (nloc) *nloc = L.nloc;
    if (slab) *slab = L.slab ? 1 : 0;
    return OMG_OK;
}

int omg_level_band(const omg_hierarchy *h, int level, double *diag, int *nband, int64_t *offsets, double *coeffs) {
    CHECK_LEVEL(h, level);
    const Level &L = h->lv[level];
    if (L.kind == OMG_KIND_CSR) {
        *nband = 0;
        return OMG_OK;
    }
    *diag = L.band.diag;
    *nband = L.band.nb;
    for (int k = 0; k < L.band.nb; ++k) {
        offsets[k] = L.band.off[k];
        coeffs[k] = L.band.coef[k];
    }
    return OMG_OK;
}

int omg_level_export_A(const omg_hierarchy *hc, int level, int32_t *indptr, int32_t *indices, double *data) {
    omg_hierarchy *h = const_cast<omg_hierarchy *>(hc);
    CHECK_LEVEL(h, level);
    const Level &L = h->lv[level];
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    if (L.ptr) {
        CUDA_TRY(cudaMemcpy(indptr, L.ptr, sizeof(int) * ((size_t)L.n + 1), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(indices, L.col, sizeof(int) * (size_t)L.nnz, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(data, L.val, sizeof(double) * (size_t)L.nnz, cudaMemcpyDeviceToHost));
        return OMG_OK;
    }
    int *p = nullptr, *c = nullptr;
    double *v = nullptr;
    int64_t nnz = 0;
    int rc = materialize_level_csr(h, L, &p, &c, &v, &nnz);
    if (rc == OMG_OK) {
        cudaMemcpy(indptr, p, sizeof(int) * ((size_t)L.n + 1), cudaMemcpyDeviceToHost);
        cudaMemcpy(indices, c, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaMemcpy(data, v, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = omg_set_error(OMG_ECUDA, "export failed: %s", cudaGetErrorString(e));
    }
    cudaFree(p);
    cudaFree(c);
    cudaFree(v);
    return rc;
}

int omg_restriction(int ndim, const int64_t *shape, int64_t *n_rows, int64_t *nnz, int32_t *indptr,
                    int32_t *indices, double *data) {
    if (!g.inited) return omg_set_error(OMG_ENODEV, "omg_init() has not succeeded");
    int *p = nullptr, *c = nullptr;
    double *v = nullptr;
    int64_t n = 0, nz = 0;
    bool query = (indptr == nullptr);
    int rc = device_restriction_csr(ndim, shape, &n, &nz, query ? nullptr : &p, query ? nullptr : &c,
                                    query ? nullptr : &v);
    if (rc != OMG_OK) return rc;
    if (n_rows) *n_rows = n;
    if (nnz) *nnz = nz;
    if (!query) {
        cudaMemcpy(indptr, p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost);
        cudaMemcpy(indices, c, sizeof(int) * (size_t)nz, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaMemcpy(data, v, sizeof(double) * (size_t)nz, cudaMemcpyDeviceToHost);
        cudaFree(p);
        cudaFree(c);
        cudaFree(v);
        if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "export failed: %s", cudaGetErrorString(e));
    }
    return OMG_OK;
}

int omg_level_export_R(const omg_hierarchy *h, int level, int32_t *indptr, int32_t *indices, double *data) {
    CHECK_LEVEL(h, level);
    const Level &L = h->lv[level];
    if (!L.hasR) return omg_set_error(OMG_EINVAL, "level %d has no restriction (coarsest)", level);
    int64_t sh[3];
    for (int i = 0; i < L.ndim; ++i) sh[i] = L.shape[i];
    return omg_restriction(L.ndim, sh, nullptr, nullptr, indptr, indices, data);
}

int omg_setup_times(const omg_hierarchy *h, double *upload_ms, double *galerkin_ms, double *coarse_factor_ms) {
    CHECK_H(h);
    if (upload_ms) *upload_ms = h->t_upload_ms;
    if (galerkin_ms) *galerkin_ms = h->t_galerkin_ms;
    if (coarse_factor_ms) *coarse_factor_ms = h->t_coarse_ms;
    return OMG_OK;
}

// ------------------------------------------------------------------ cycles

// host vectors are GLOBAL (n entries); each rank moves the slice of the rows it owns
static int up(const Level &L, double *dst_owned, const double *src_host) {
    CUDA_TRY(cudaMemcpyAsync(dst_owned, src_host + L.row0, sizeof(double) * (size_t)L.nloc, cudaMemcpyHostToDevice,
                             g.stream));
    return OMG_OK;
}
static int down(const Level &L, double *dst_host, const double *src_owned) {
    CUDA_TRY(cudaMemcpyAsync(dst_host + L.row0, src_owned, sizeof(double) * (size_t)L.nloc, cudaMemcpyDeviceToHost,
                             g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    return OMG_OK;
}

static int exec_cycle(omg_hierarchy *h, CycleCfg cfg) {
    cfg.cur0 = h->cur0;
    if (h->flags & OMG_FLAG_NO_GRAPH) return run_cycle(h, cfg);
    auto it = h->graphs.find(cfg);
    if (it == h->graphs.end()) {
        int64_t l0 = h->launches;
        CUDA_TRY(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
        int rc = run_cycle(h, cfg);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(g.stream, &graph);
        if (rc != OMG_OK) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
        CachedGraph cg;
        e = cudaGraphInstantiate(&cg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        cg.cur0_after = h->cur0;
        cg.launches = h->launches - l0;
        h->launches = l0;
        h->cur0 = cfg.cur0;
        it = h->graphs.insert({cfg, cg}).first;
    }
    CUDA_TRY(cudaGraphLaunch(it->second.exec, g.stream));
    h->cur0 = it->second.cur0_after;
    h->launches += it->second.launches;
    return OMG_OK;
}

// ---- device-side stop rule (thresholded solves without a host round trip per cycle)

__global__ void k_stop_gate(cudaGraphConditionalHandle hnd, const StopState *s) {
    cudaGraphSetConditional(hnd, s->done ? 0u : 1u);
}

// after a cycle and its norm: count it, record the norm, apply openmg/__init__.py:121-130
__global__ void k_stop_update(StopState *s, const double *norm2, double *hist) {
    const double nv = sqrt(*norm2);
    const int c = ++s->cycle;
    s->norm = nv;
    if (hist && c <= s->hist_cap) hist[c - 1] = nv;
    const bool cycleStop = s->max_cycles > 0 && c >= s->max_cycles;
    const bool thresholdStop = s->threshold > 0.0 && nv < s->threshold;
    if (cycleStop || thresholdStop) s->done = 1;
}

// graph = [gate kernel] -> IF(!done) { one V-cycle, residual norm, stop update }
static int exec_gated_cycle(omg_hierarchy *h, CycleCfg cfg) {
    cfg.cur0 = h->cur0;
    cfg.with_norm = 1;
    auto it = h->gated.find(cfg);
    if (it == h->gated.end()) {
        int64_t l0 = h->launches;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaGraphCreate(&graph, 0));
        cudaGraphConditionalHandle hnd;
        CUDA_TRY(cudaGraphConditionalHandleCreate(&hnd, graph, 0, 0));
        cudaGraphNode_t gate = nullptr, cond = nullptr;
        {
            cudaGraphNodeParams kp = {cudaGraphNodeTypeKernel};
            const StopState *sp = h->stop_dev;
            void *args[2] = {(void *)&hnd, (void *)&sp};
            kp.kernel.func = (void *)k_stop_gate;
            kp.kernel.gridDim = dim3(1);
            kp.kernel.blockDim = dim3(1);
            kp.kernel.kernelParams = args;
            CUDA_TRY(cudaGraphAddNode(&gate, graph, nullptr, 0, &kp));
        }
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.conditional.handle = hnd;
        cp.conditional.type = cudaGraphCondTypeIf;
        cp.conditional.size = 1;
        CUDA_TRY(cudaGraphAddNode(&cond, graph, &gate, 1, &cp));
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        CUDA_TRY(cudaStreamBeginCaptureToGraph(g.stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        int rc = run_cycle(h, cfg);
        k_stop_update<<<1, 1, 0, g.stream>>>(h->stop_dev, h->norm2_dev, h->hist_dev);
        cudaGraph_t got = nullptr;
        cudaError_t e = cudaStreamEndCapture(g.stream, &got);
        if (rc != OMG_OK || e != cudaSuccess) {
            cudaGraphDestroy(graph);
            if (rc != OMG_OK) return rc;
            return omg_set_error(OMG_ECUDA, "gated graph capture failed: %s", cudaGetErrorString(e));
        }
        CachedGraph cg;
        e = cudaGraphInstantiate(&cg.exec, graph, 0);
        cudaGraphDestroy(graph);
        cg.cur0_after = h->cur0;
        cg.launches = h->launches - l0 + 2;
        h->launches = l0;
        h->cur0 = cfg.cur0;
        if (e != cudaSuccess) {
            // the body holds something a conditional node may not contain (e.g. a captured NCCL collective)
            cudaGetLastError();
            h->no_gate = true;
            return OMG_EUNSUPPORTED;
        }
        it = h->gated.insert({cfg, cg}).first;
    }
    CUDA_TRY(cudaGraphLaunch(it->second.exec, g.stream));
    h->cur0 = it->second.cur0_after;
    h->launches += it->second.launches;
    return OMG_OK;
}

// The cycle loop of mgSolve with the stop test on the device: cycles are enqueued in growing batches (1, 2, 4, ...
// up to 16); once the test fires the remaining graphs of the batch skip their body.  One stream synchronisation per
// batch instead of one per cycle; the result buffer follows from the number of cycles that really ran.
static int solve_gated(omg_hierarchy *h, CycleCfg cfg, int cycles, double threshold, int *cycle_out, double *norm_out,
                       double *norm_hist, int hist_cap) {
    if (!h->stop_dev) {
        OMG_TRY(h_alloc_t(h, &h->stop_dev, 1, true));
        CUDA_TRY(cudaMallocHost((void **)&h->stop_host, sizeof(StopState)));
    }
    const int want_hist = (norm_hist && hist_cap > 0) ? hist_cap : 0;
    if (want_hist > h->hist_dev_cap) {
        if (h->hist_dev) h_free(h, h->hist_dev);
        h->hist_dev = nullptr;
        // the gated graphs captured the old buffer address
        for (auto &kv : h->gated)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        h->gated.clear();
        OMG_TRY(h_alloc_t(h, &h->hist_dev, (size_t)want_hist, true));
        h->hist_dev_cap = want_hist;
    }
    const bool both_disabled = (threshold <= 0.0 && cycles <= 0);
    StopState st{threshold, 0.0, both_disabled ? 1 : cycles, 0, 0, want_hist ? std::min(want_hist, h->hist_dev_cap) : 0};
    *h->stop_host = st;
    CUDA_TRY(cudaMemcpyAsync(h->stop_dev, h->stop_host, sizeof(StopState), cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));     // stop_host is reused as the read-back buffer below
    std::vector<int> cur0_seq(1, h->cur0);          // iterate buffer after k executed cycles
    int issued = 0, batch = 1;
    for (;;) {
        if (cycles > 0) batch = std::min(batch, cycles - issued);
        for (int i = 0; i < batch; ++i) {
            OMG_TRY(exec_gated_cycle(h, cfg));
            cur0_seq.push_back(h->cur0);
        }
        issued += batch;
        CUDA_TRY(cudaMemcpyAsync(h->stop_host, h->stop_dev, sizeof(StopState), cudaMemcpyDeviceToHost, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        h->host_syncs++;
        if (h->stop_host->done || (cycles > 0 && issued >= cycles)) break;
        batch = std::min(batch * 2, 16);
    }
    const int ran = h->stop_host->cycle;
    h->cur0 = cur0_seq[std::min<size_t>((size_t)ran, cur0_seq.size() - 1)];
    *cycle_out = ran;
    *norm_out = h->stop_host->norm;
    if (want_hist && ran > 0) {
        CUDA_TRY(cudaMemcpyAsync(norm_hist, h->hist_dev, sizeof(double) * (size_t)std::min(ran, want_hist),
                                 cudaMemcpyDeviceToHost, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
    }
    return OMG_OK;
}

static int read_norm(omg_hierarchy *h, double *norm) {
    CUDA_TRY(cudaMemcpyAsync(h->norm2_host, h->norm2_dev, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    *norm = sqrt(h->norm2_host[0]);
    return OMG_OK;
}

static int upload_state(omg_hierarchy *h, const double *b_host, const double *x_host, int has_initial) {
    Level &L = h->lv[0];
    size_t bytes = sizeof(double) * (size_t)L.nloc;
    if (b_host) CUDA_TRY(cudaMemcpyAsync(L.b, b_host + L.row0, bytes, cudaMemcpyHostToDevice, g.stream));
    if (has_initial && x_host)
        CUDA_TRY(cudaMemcpyAsync(L.xa, x_host + L.row0, bytes, cudaMemcpyHostToDevice, g.stream));
    else
        CUDA_TRY(cudaMemsetAsync(L.xa, 0, bytes, g.stream));
    h->cur0 = 0;
    return OMG_OK;
}

// a stencil kernel gave up waiting for a neighbour during a fused halo pull: the iterate is invalid
static int check_pull(omg_hierarchy *h) {
    bool bad = false;
    OMG_TRY(dist_pull_timed_out(h, &bad));
    if (bad)
        return omg_set_error(OMG_ECUDA, "a fused halo pull timed out waiting for a neighbouring rank (set "
                                        "OMG_NO_HALO_PULL=1 to exchange halos with separate copies)");
    return OMG_OK;
}

static int download_x(omg_hierarchy *h, double *x_host) {
    if (h->first_replicated > 0) OMG_TRY(check_pull(h));
    Level &L = h->lv[0];
    const double *cur = h->cur0 ? L.xb : L.xa;
    CUDA_TRY(cudaMemcpyAsync(x_host + L.row0, cur, sizeof(double) * (size_t)L.nloc, cudaMemcpyDeviceToHost,
                             g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return OMG_OK;
}

static int check_cfg(const omg_hierarchy *h, int pre, int post, int smoother, double omega) {
    if (pre < 0 || post < 0) return omg_set_error(OMG_EINVAL, "preIterations/postIterations must be >= 0");
    if (smoother < 0 || smoother > 2) return omg_set_error(OMG_EINVAL, "unknown smoother id %d", smoother);
    if (!(omega > 0.0)) return omg_set_error(OMG_EINVAL, "omega must be > 0");
    // the reference's lexicographic sweep is one sequential chain over all rows: it cannot run on row slabs
    if (smoother == OMG_SMOOTH_LEXGS && h->first_replicated > 0)
        return omg_set_error(OMG_EUNSUPPORTED, "smoother 'gs' (lexicographic Gauss-Seidel) is sequential and not "
                             "available on a hierarchy sharded across GPUs; use 'rbgs' or 'jacobi'");
    return OMG_OK;
}

int omg_solve(omg_hierarchy *h, const double *b_host, double *x_host, int has_initial, int pre, int post,
              int smoother, double omega, int cycles, double threshold, int *cycles_done, double *final_norm,
              double *norm_hist, int hist_cap) {
    CHECK_H(h);
    OMG_TRY(check_cfg(h, pre, post, smoother, omega));
    OMG_TRY(upload_state(h, b_host, x_host, has_initial));
    bool every = (threshold > 0.0) || (norm_hist != nullptr && hist_cap > 0);
    CycleCfg cfg{pre, post, smoother, 0, omega, 0};
    int cycle = 0;
    double norm = 0.0;
    // do at least one cycle (openmg/__init__.py:112)
    bool both_disabled = (threshold <= 0.0 && cycles <= 0);
    h->host_syncs = 0;
    // a norm after every cycle (threshold stop, or a residual history was asked for): stop test on the device.  Not on
    // sharded hierarchies: the collectives NCCL captures cannot sit inside a conditional graph node, so there the
    // host reads the (all-reduced) norm back after every cycle as before.
    int grc = OMG_EUNSUPPORTED;
    if (every && !(h->flags & OMG_FLAG_NO_GRAPH) && !h->no_gate && h->first_replicated == 0)
        grc = solve_gated(h, cfg, cycles, threshold, &cycle, &norm, norm_hist, hist_cap);
    if (grc != OMG_OK && grc != OMG_EUNSUPPORTED) return grc;      // unsupported: nothing was launched yet
    if (grc == OMG_OK) {
        if (cycles_done) *cycles_done = cycle;
        if (final_norm) *final_norm = norm;
        if (both_disabled)      // ValueError raised after the first cycle (:118-119)
            return omg_set_error(OMG_EINVAL, "Either parameters['threshold'] or parameters['cycles'] must be > 0.");
        return download_x(h, x_host);
    }
    for (;;) {
        // without a threshold the norm is only needed after the last cycle (:140-141)
        bool last_known = !every && (both_disabled || cycle + 1 >= cycles);
        cfg.with_norm = (every || last_known) ? 1 : 0;
        OMG_TRY(exec_cycle(h, cfg));
        ++cycle;
        if (cfg.with_norm) {
            OMG_TRY(read_norm(h, &norm));
            h->host_syncs++;
            if (norm_hist && cycle <= hist_cap) norm_hist[cycle - 1] = norm;
        }
        if (both_disabled) {   // ValueError raised after the first cycle (:118-119)
            if (cycles_done) *cycles_done = cycle;
            if (final_norm) *final_norm = norm;
            return omg_set_error(OMG_EINVAL, "Either parameters['threshold'] or parameters['cycles'] must be > 0.");
        }
        bool cycleStop = cycles > 0 && cycle >= cycles;                 // :124-125
        bool thresholdStop = threshold > 0.0 && norm < threshold;      // :127-128
        if (cycleStop || thresholdStop) break;
    }
    if (cycles_done) *cycles_done = cycle;
    if (final_norm) *final_norm = norm;
    return download_x(h, x_host);
}

int omg_solve_stats(const omg_hierarchy *h, int64_t *host_syncs, double *coarse_defect) {
    CHECK_H(h);
    if (host_syncs) *host_syncs = h->host_syncs;
    if (coarse_defect) *coarse_defect = h->coarse_defect;
    return OMG_OK;
}

int omg_cycle(omg_hierarchy *h, int level, const double *b_host, double *x_host, int has_initial, int pre,
              int post, int smoother, double omega, double *norm) {
    CHECK_LEVEL(h, level);
    OMG_TRY(check_cfg(h, pre, post, smoother, omega));
    double nv = 0;
    if (level == 0) {
        OMG_TRY(upload_state(h, b_host, x_host, has_initial));
        CycleCfg cfg{pre, post, smoother, 1, omega, 0};
        OMG_TRY(exec_cycle(h, cfg));
        OMG_TRY(read_norm(h, &nv));
        if (norm) *norm = nv;
        return download_x(h, x_host);
    }
    Level &L = h->lv[level];
    OMG_TRY(up(L, L.b, b_host));
    if (has_initial) OMG_TRY(up(L, L.xa, x_host));
    CycleCfg cfg{pre, post, smoother, 0, omega, 0};
    double *cur = cycle_from_level(h, level, cfg, has_initial ? L.xa : nullptr);
    if (level < h->nlev - 1) {
        OMG_TRY(launch_resnorm2(h, L, cur, L.b, 1));
        CUDA_TRY(cudaMemcpyAsync(h->norm2_host + 1, h->norm2_dev + 1, sizeof(double), cudaMemcpyDeviceToHost,
                                 g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        nv = sqrt(h->norm2_host[1]);
    }
    if (norm) *norm = nv;
    return down(L, x_host, cur);
}

int omg_set_rhs(omg_hierarchy *h, const double *b_host) {
    CHECK_H(h);
    OMG_TRY(upload_state(h, b_host, nullptr, 0));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return OMG_OK;
}

int omg_bench_cycles(omg_hierarchy *h, int pre, int post, int smoother, double omega, int ncycles, int with_norm,
                     float *ms, int64_t *launches) {
    CHECK_H(h);
    OMG_TRY(check_cfg(h, pre, post, smoother, omega));
    if (ncycles <= 0) return omg_set_error(OMG_EINVAL, "ncycles must be > 0");
    CycleCfg cfg{pre, post, smoother, with_norm ? 1 : 0, omega, 0};
    // make sure both ping-pong parities are captured outside the timed region
    OMG_TRY(upload_state(h, nullptr, nullptr, 0));
    OMG_TRY(exec_cycle(h, cfg));
    OMG_TRY(exec_cycle(h, cfg));
    OMG_TRY(upload_state(h, nullptr, nullptr, 0));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    int64_t l0 = h->launches;
    CUDA_TRY(cudaEventRecord(e0, g.stream));
    int rc = OMG_OK;
    for (int c = 0; c < ncycles && rc == OMG_OK; ++c) rc = exec_cycle(h, cfg);
    cudaEventRecord(e1, g.stream);
    cudaError_t e = cudaStreamSynchronize(g.stream);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != OMG_OK) return rc;
    if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "bench failed: %s", cudaGetErrorString(e));
    if (h->first_replicated > 0) OMG_TRY(check_pull(h));
    if (ms) *ms = t;
    if (launches) *launches = h->launches - l0;
    return OMG_OK;
}

// Per-kernel CUDA-event timing of `reps` cycles launched directly (no graph): writes a JSON
// array [{"name":..,"level":..,"launches":..,"ms":avg per launch,"bytes":algorithmic per launch},..]
int omg_profile_cycle(omg_hierarchy *h, int pre, int post, int smoother, double omega, int reps, char *json,
                      int cap) {
    CHECK_H(h);
    OMG_TRY(check_cfg(h, pre, post, smoother, omega));
    if (reps <= 0 || !json || cap < 64) return omg_set_error(OMG_EINVAL, "bad arguments to omg_profile_cycle");
    CycleCfg cfg{pre, post, smoother, 0, omega, 0};
    OMG_TRY(upload_state(h, nullptr, nullptr, 0));
    OMG_TRY(run_cycle(h, cfg));   // warm
    OMG_TRY(run_cycle(h, cfg));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    h->profiling = true;
    h->prof.clear();
    int rc = OMG_OK;
    for (int r = 0; r < reps && rc == OMG_OK; ++r) {
        cfg.cur0 = h->cur0;
        rc = run_cycle(h, cfg);
    }
    h->profiling = false;
    cudaError_t e = cudaStreamSynchronize(g.stream);
    struct Agg {
        const char *name;
        int level;
        int n;
        double ms, bytes;
    };
    std::vector<Agg> agg;
    for (auto &r : h->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        bool found = false;
        for (auto &a : agg)
            if (a.level == r.level && strcmp(a.name, r.name) == 0) {
                a.n++;
                a.ms += ms;
                found = true;
                break;
            }
        if (!found) agg.push_back(Agg{r.name, r.level, 1, ms, r.bytes});
    }
    h->prof.clear();
    if (rc != OMG_OK) return rc;
    if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "profile failed: %s", cudaGetErrorString(e));
    std::string out = "[";
    for (size_t i = 0; i < agg.size(); ++i) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s{\"name\":\"%s\",\"level\":%d,\"launches\":%d,\"ms\":%.6f,\"bytes\":%.0f}",
                 i ? "," : "", agg[i].name, agg[i].level, agg[i].n, agg[i].ms / agg[i].n, agg[i].bytes);
        out += buf;
    }
    out += "]";
    if ((int)out.size() + 1 > cap) return omg_set_error(OMG_EINVAL, "profile buffer too small");
    memcpy(json, out.c_str(), out.size() + 1);
    return OMG_OK;
}

int omg_get_solution(omg_hierarchy *h, double *x_host) {
    CHECK_H(h);
    return download_x(h, x_host);
}

int omg_current_norm(omg_hierarchy *h, double *norm) {
    CHECK_H(h);
    Level &L = h->lv[0];
    launch_resnorm2(h, L, h->cur0 ? L.xb : L.xa, L.b, 1);
    CUDA_TRY(cudaMemcpyAsync(h->norm2_host + 1, h->norm2_dev + 1, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    *norm = sqrt(h->norm2_host[1]);
    return OMG_OK;
}

// ------------------------------------------------------------------ unit entry points

int omg_smooth(omg_hierarchy *h, int level, const double *b_host, double *x_host, int sweeps, int smoother,
               double omega) {
    CHECK_LEVEL(h, level);
    OMG_TRY(check_cfg(h, sweeps, 0, smoother, omega));
    Level &L = h->lv[level];
    OMG_TRY(up(L, L.b, b_host));
    OMG_TRY(up(L, L.xa, x_host));
    double *cur = launch_smooth(h, L, smoother, omega, sweeps, L.xa, L.b);
    if (level == 0) h->cur0 = (cur == L.xb);
    return down(L, x_host, cur);
}

int omg_residual_restrict(omg_hierarchy *h, int level, const double *b_host, const double *x_host,
                          double *rc_host) {
    CHECK_LEVEL(h, level);
    if (level >= h->nlev - 1) return omg_set_error(OMG_EINVAL, "level %d has no restriction", level);
    Level &L = h->lv[level];
    Level &C = h->lv[level + 1];
    OMG_TRY(up(L, L.b, b_host));
    OMG_TRY(up(L, L.xa, x_host));
    OMG_TRY(launch_residual_restrict(h, level, L.xa, L.b, C.b));
    return down(C, rc_host, C.b);
}

int omg_smooth_residual_restrict(omg_hierarchy *h, int level, const double *b_host, double *x_host, int sweeps,
                                 int smoother, double omega, double *rc_host) {
    CHECK_LEVEL(h, level);
    if (level >= h->nlev - 1) return omg_set_error(OMG_EINVAL, "level %d has no restriction", level);
    OMG_TRY(check_cfg(h, sweeps, 0, smoother, omega));
    Level &L = h->lv[level];
    Level &C = h->lv[level + 1];
    OMG_TRY(up(L, L.b, b_host));
    OMG_TRY(up(L, L.xa, x_host));
    double *cur = L.xa;
    OMG_TRY(launch_smooth_residual_restrict(h, level, smoother, omega, sweeps, &cur, L.b, C.b));
    if (level == 0) h->cur0 = (cur == L.xb);
    OMG_TRY(down(L, x_host, cur));
    return down(C, rc_host, C.b);
}

int omg_prolong_correct(omg_hierarchy *h, int level, const double *ec_host, double *x_host) {
    CHECK_LEVEL(h, level);
    if (level >= h->nlev - 1) return omg_set_error(OMG_EINVAL, "level %d has no restriction", level);
    Level &L = h->lv[level];
    Level &C = h->lv[level + 1];
    OMG_TRY(up(C, C.xa, ec_host));
    OMG_TRY(up(L, L.xa, x_host));
    OMG_TRY(launch_prolong_correct(h, level, C.xa, L.xa, L.xa));
    return down(L, x_host, L.xa);
}

int omg_prolong_correct_smooth(omg_hierarchy *h, int level, const double *b_host, const double *ec_host,
                               double *x_host, int sweeps, int smoother, double omega) {
    CHECK_LEVEL(h, level);
    if (level >= h->nlev - 1) return omg_set_error(OMG_EINVAL, "level %d has no restriction", level);
    OMG_TRY(check_cfg(h, sweeps, 0, smoother, omega));
    Level &L = h->lv[level];
    Level &C = h->lv[level + 1];
    OMG_TRY(up(C, C.xa, ec_host));
    OMG_TRY(up(L, L.xa, x_host));
    OMG_TRY(up(L, L.b, b_host));
    double *cur = launch_prolong_correct_smooth(h, level, smoother, omega, sweeps, L.xa, C.xa, L.b, false);
    if (level == 0) h->cur0 = (cur == L.xb);
    return down(L, x_host, cur);
}

int omg_smooth_to_threshold(omg_hierarchy *h, int level, const double *b_host, double *x_host, double threshold,
                            int max_sweeps, int smoother, double omega, int *sweeps_done, double *norm) {
    CHECK_LEVEL(h, level);
    OMG_TRY(check_cfg(h, 0, 0, smoother, omega));
    Level &L = h->lv[level];
    OMG_TRY(up(L, L.b, b_host));
    OMG_TRY(up(L, L.xa, x_host));
    double *cur = L.xa;
    int it = 0;
    double nv = 0.0;
    for (;;) {
        OMG_TRY(launch_resnorm2(h, L, cur, L.b, 1));
        CUDA_TRY(cudaMemcpyAsync(h->norm2_host + 1, h->norm2_dev + 1, sizeof(double), cudaMemcpyDeviceToHost,
                                 g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        nv = sqrt(h->norm2_host[1]);
        if (nv < threshold || it >= max_sweeps) break;
        cur = launch_smooth(h, L, smoother, omega, 1, cur, L.b);
        ++it;
    }
    if (level == 0) h->cur0 = (cur == L.xb);
    if (sweeps_done) *sweeps_done = it;
    if (norm) *norm = nv;
    return down(L, x_host, cur);
}

int omg_coarse_solve(omg_hierarchy *h, const double *b_host, double *x_host) {
    CHECK_H(h);
    if (!h->Ainv) return omg_set_error(OMG_EINVAL, "this handle has no direct-solve factor");
    Level &L = h->lv[h->nlev - 1];
    OMG_TRY(up(L, L.b, b_host));
    OMG_TRY(launch_coarse_solve(h, L.b, L.xa));
    return down(L, x_host, L.xa);
}

int omg_residual_norm(omg_hierarchy *h, int level, const double *b_host, const double *x_host, double *norm) {
    CHECK_LEVEL(h, level);
    Level &L = h->lv[level];
    OMG_TRY(up(L, L.b, b_host));
    OMG_TRY(up(L, L.xa, x_host));
    if (level == 0) h->cur0 = 0;
    OMG_TRY(launch_resnorm2(h, L, L.xa, L.b, 1));
    CUDA_TRY(cudaMemcpyAsync(h->norm2_host + 1, h->norm2_dev + 1, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    *norm = sqrt(h->norm2_host[1]);
    return OMG_OK;
}

int omg_residual(omg_hierarchy *h, int level, const double *b_host, const double *x_host, double *r_host) {
    CHECK_LEVEL(h, level);
    Level &L = h->lv[level];
    OMG_TRY(up(L, L.b, b_host));
    OMG_TRY(up(L, L.xa, x_host));
    if (level == 0) h->cur0 = 0;
    OMG_TRY(launch_residual(h, L, L.xa, L.b, L.xb));
    int rc = down(L, r_host, L.xb);
    return rc;
}

int omg_matvec(omg_hierarchy *h, int level, const double *x_host, double *y_host) {
    CHECK_LEVEL(h, level);
    Level &L = h->lv[level];
    OMG_TRY(up(L, L.xa, x_host));
    if (level == 0) h->cur0 = 0;
    OMG_TRY(launch_matvec(h, L, L.xa, L.xb));
    return down(L, y_host, L.xb);
}

}   // extern "C"
