// omg_dist.cu — multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Fine ("slab") levels are partitioned into contiguous row slabs along the leading grid
// dimension, cut so that no restriction aggregate straddles two ranks (R and R^T need no
// communication).  What does cross ranks:
//   * halo planes of a vector before an operator is applied to it      ncclSend/ncclRecv pairs
//   * the coarse right-hand side at the slab -> replicated transition  ncclAllGather
//   * the residual norm                                                 ncclAllReduce (1 double)
// Coarse levels below the agglomeration threshold are replicated on every rank (each GPU runs
// them redundantly: one collective instead of gather + scatter).
//
// NCCL is dlopen'ed (the torch-bundled libnccl.so.2 when the process already loaded it), so a
// single-GPU process never needs it.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include "omg_hier.cuh"

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
static NcclApi nccl;

static int nccl_load() {
    if (nccl.lib) return OMG_OK;
    const char *cands[] = {getenv("OMG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // already in the process (torch)?
    for (int i = 0; !lib && i < 3; ++i)
        if (cands[i]) lib = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return omg_set_error(OMG_ENCCL, "cannot load libnccl.so.2 (%s)", dlerror());
#define SYM(field, name)                                                                     \
    *(void **)(&nccl.field) = dlsym(lib, name);                                              \
    if (!nccl.field) return omg_set_error(OMG_ENCCL, "libnccl.so.2 lacks %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
    nccl.lib = lib;
    return OMG_OK;
}

#define NCCL_TRY(expr)                                                                           \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess)                                                                   \
            return omg_set_error(OMG_ENCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,         \
                                 nccl.GetErrorString ? nccl.GetErrorString(_r) : "nccl error"); \
    } while (0)

extern "C" {

int omg_nccl_unique_id(unsigned char id[128]) {
    OMG_TRY(nccl_load());
    ncclUniqueId u;
    NCCL_TRY(nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return OMG_OK;
}

int omg_dist_init(int rank, int nranks, const unsigned char id[128]) {
    if (!g.inited) return omg_set_error(OMG_ENODEV, "omg_init() has not succeeded");
    if (nranks < 1 || rank < 0 || rank >= nranks) return omg_set_error(OMG_EINVAL, "bad rank %d / %d", rank, nranks);
    if (g.nccl_comm) return omg_set_error(OMG_EINVAL, "omg_dist_init called twice");
    g.rank = rank;
    g.nranks = nranks;
    if (nranks == 1) return OMG_OK;
    OMG_TRY(nccl_load());
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclComm_t comm;
    NCCL_TRY(nccl.CommInitRank(&comm, nranks, u, rank));
    g.nccl_comm = comm;
    return OMG_OK;
}

int omg_dist_rank(int *rank, int *nranks) {
    if (rank) *rank = g.rank;
    if (nranks) *nranks = g.nranks;
    return OMG_OK;
}

// Pure host logic (no device): the slab partition of a hierarchy.  level_lead[l] = leading grid
// extent of level l (shape[0] >> l), level_rows[l] = rows of level l.  Levels [0, *first_replicated)
// are slab levels.  row0/nloc per level for `rank`.
int omg_partition(int nlevels, const int64_t *level_lead, const int64_t *level_rows, const int32_t *level_regular,
                  int nranks, int rank, int64_t agglomerate_below, int32_t *first_replicated, int64_t *row0,
                  int64_t *nloc) {
    if (nlevels < 1 || nranks < 1 || rank < 0 || rank >= nranks) return omg_set_error(OMG_EINVAL, "bad partition request");
    int ld = 0;
    if (nranks > 1) {
        // slab levels: regular restriction below them, enough rows, leading extent divisible so that
        // every rank gets an even, equal number of leading units down to the transition level
        while (ld < nlevels - 1 && level_regular[ld] && level_rows[ld] > agglomerate_below) {
            int64_t lead_next = level_lead[ld + 1];
            if (lead_next % nranks != 0 || level_lead[ld] != 2 * lead_next) break;
            if (level_rows[ld] % level_lead[ld] != 0) break;
            ++ld;
        }
    }
    *first_replicated = ld;
    for (int l = 0; l < nlevels; ++l) {
        if (l < ld) {
            int64_t lead = level_lead[l];
            int64_t per = lead / nranks;            // divisible by construction (lead = 2^(ld-l) * lead_ld)
            int64_t unit = level_rows[l] / lead;
            row0[l] = per * rank * unit;
            nloc[l] = per * unit;
        } else {
            row0[l] = 0;
            nloc[l] = level_rows[l];
        }
    }
    return OMG_OK;
}

}   // extern "C"

// ---------------------------------------------------------------- device-side collectives (library stream)

// All NCCL traffic runs on the library's second (high-priority) stream so that a halo exchange can
// overlap the interior planes of the kernel that consumes it:
//   compute stream:  ... producer | interior z-segments ........ | wait(join) | boundary z-segments
//   comm stream:         wait(fork) | ncclSend/ncclRecv | record(join)
static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

static int comm_fork() {
    if (!ev_fork) {
        CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(ev_fork, g.stream));
    CUDA_TRY(cudaStreamWaitEvent(g.stream2, ev_fork, 0));
    return OMG_OK;
}

// ---- peer-memory path: copy-engine pulls ordered by device-side flags (no SMs, no NCCL kernels)
//
//   signal_ready: "my boundary rows of this buffer are final"   -> epoch written into both neighbours' flags
//   wait_ready  : spin until both neighbours said so
//   2 x cudaMemcpyAsync (peer -> my halos), copy engines over NVLink
//   signal_done : "I have pulled"                                 -> neighbours may overwrite their rows again
//   wait_done   : spin until both neighbours have pulled from me
// Epochs live in device memory, so a captured CUDA graph can be replayed.

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// "my rows are final" (bump the slot's epoch, tell both neighbours), then wait until theirs are
__global__ void k_halo_ready(unsigned long long *epoch, unsigned long long *peer_dn, unsigned long long *peer_up,
                             const unsigned long long *mine_dn, const unsigned long long *mine_up) {
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long e = *epoch + 1ull;
    *epoch = e;
    __threadfence_system();
    if (peer_dn) st_release_sys(peer_dn, e);
    if (peer_up) st_release_sys(peer_up, e);
    if (mine_dn)
        while (ld_acquire_sys(mine_dn) < e) __nanosleep(32);
    if (mine_up)
        while (ld_acquire_sys(mine_up) < e) __nanosleep(32);
    __threadfence_system();
}

// "I have pulled" to both neighbours, then wait until both have pulled from me
__global__ void k_halo_done(const unsigned long long *epoch, unsigned long long *peer_dn, unsigned long long *peer_up,
                            const unsigned long long *mine_dn, const unsigned long long *mine_up) {
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long e = *epoch;
    __threadfence_system();
    if (peer_dn) st_release_sys(peer_dn, e);
    if (peer_up) st_release_sys(peer_up, e);
    if (mine_dn)
        while (ld_acquire_sys(mine_dn) < e) __nanosleep(32);
    if (mine_up)
        while (ld_acquire_sys(mine_up) < e) __nanosleep(32);
}

// small halos: the whole exchange in ONE single-CTA kernel (flags, NVLink loads, flags)
__global__ void __launch_bounds__(1024) k_halo_small(unsigned long long *epoch, unsigned long long *peer_dn_ready,
                                                     unsigned long long *peer_up_ready,
                                                     unsigned long long *peer_dn_done, unsigned long long *peer_up_done,
                                                     const unsigned long long *mine, int has_dn, int has_up,
                                                     const double2 *__restrict__ src_dn, double2 *__restrict__ dst_lo,
                                                     const double2 *__restrict__ src_up, double2 *__restrict__ dst_hi,
                                                     int n2) {
    __shared__ unsigned long long se;
    if (threadIdx.x == 0) {
        unsigned long long e = *epoch + 1ull;
        *epoch = e;
        se = e;
        __threadfence_system();
        if (has_dn) st_release_sys(peer_dn_ready, e);
        if (has_up) st_release_sys(peer_up_ready, e);
        if (has_dn)
            while (ld_acquire_sys(mine + 0) < e) __nanosleep(32);
        if (has_up)
            while (ld_acquire_sys(mine + 1) < e) __nanosleep(32);
        __threadfence_system();
    }
    __syncthreads();
    // NVLink loads (not cached: ld.cv), 8 in flight per thread
    for (int side = 0; side < 2; ++side) {
        const double2 *src = side ? src_up : src_dn;
        double2 *dst = side ? dst_hi : dst_lo;
        if (!(side ? has_up : has_dn)) continue;
        for (int i0 = threadIdx.x; i0 < n2; i0 += 8 * blockDim.x) {
            double2 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                int i = i0 + u * blockDim.x;
                if (i < n2) t[u] = __ldcv(src + i);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                int i = i0 + u * blockDim.x;
                if (i < n2) dst[i] = t[u];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long e = se;
        if (has_dn) st_release_sys(peer_dn_done, e);
        if (has_up) st_release_sys(peer_up_done, e);
        if (has_dn)
            while (ld_acquire_sys(mine + 2) < e) __nanosleep(32);
        if (has_up)
            while (ld_acquire_sys(mine + 3) < e) __nanosleep(32);
    }
}

static int peer_slot(const Level &L, int l, const double *v) {
    if (v == L.xa) return 3 * l + 0;
    if (v == L.xb) return 3 * l + 1;
    if (v == L.b) return 3 * l + 2;
    return -1;
}

static int peer_exchange(omg_hierarchy *h, Level &L, double *v) {
    PeerState &P = h->peer;
    int l = (int)(&L - h->lv.data());
    int slot = peer_slot(L, l, v);
    size_t hw = (size_t)L.halo;
    bool has_dn = g.rank > 0, has_up = g.rank < g.nranks - 1;
    unsigned long long *ep = P.epochs + slot;
    unsigned long long *mine = P.flags + 4 * slot;
    // the flag words I write live in the neighbours' arrays: I am the DN neighbour of `up`, the UP neighbour of `dn`
    unsigned long long *up_ready = has_up ? P.flags_up + 4 * slot + 0 : nullptr;
    unsigned long long *dn_ready = has_dn ? P.flags_dn + 4 * slot + 1 : nullptr;
    unsigned long long *up_done = has_up ? P.flags_up + 4 * slot + 2 : nullptr;
    unsigned long long *dn_done = has_dn ? P.flags_dn + 4 * slot + 3 : nullptr;
    cudaStream_t st = g.stream2;
    size_t base_off = (size_t)L.pad;                                    // owned row 0 inside the allocation
    const double *src_dn = has_dn ? P.base_dn[slot] + base_off + L.nloc - hw : nullptr;   // lower neighbour's top rows
    const double *src_up = has_up ? P.base_up[slot] + base_off : nullptr;                  // upper neighbour's bottom rows
    if (hw <= (1u << 15) + 4096 && (hw & 1) == 0) {
        k_halo_small<<<1, 1024, 0, st>>>(ep, dn_ready, up_ready, dn_done, up_done, mine, has_dn, has_up,
                                         (const double2 *)src_dn, (double2 *)(v - hw), (const double2 *)src_up,
                                         (double2 *)(v + L.nloc), (int)(hw / 2));
        h->launches += 1;
        return OMG_OK;
    }
    k_halo_ready<<<1, 32, 0, st>>>(ep, dn_ready, up_ready, has_dn ? mine + 0 : nullptr, has_up ? mine + 1 : nullptr);
    if (has_dn) CUDA_TRY(cudaMemcpyAsync(v - hw, src_dn, hw * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (has_up) CUDA_TRY(cudaMemcpyAsync(v + L.nloc, src_up, hw * sizeof(double), cudaMemcpyDeviceToDevice, st));
    k_halo_done<<<1, 32, 0, st>>>(ep, dn_done, up_done, has_dn ? mine + 2 : nullptr, has_up ? mine + 3 : nullptr);
    h->launches += 2;
    return OMG_OK;
}

static int allgather_bytes(const void *mine, void *all_host, size_t bytes) {
    unsigned char *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, bytes * (size_t)g.nranks));
    CUDA_TRY(cudaMemcpy(d + bytes * (size_t)g.rank, mine, bytes, cudaMemcpyHostToDevice));
    NCCL_TRY(nccl.AllGather(d + bytes * (size_t)g.rank, d, bytes, ncclChar, (ncclComm_t)g.nccl_comm, g.stream2));
    CUDA_TRY(cudaStreamSynchronize(g.stream2));
    CUDA_TRY(cudaMemcpy(all_host, d, bytes * (size_t)g.nranks, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return OMG_OK;
}

// Map the neighbours' slab-level buffers and flag words (collective: every rank calls it for the same hierarchy)
int dist_peer_setup(omg_hierarchy *h) {
    PeerState &P = h->peer;
    const char *mode = getenv("OMG_HALO");
    if (g.nranks == 1 || h->first_replicated == 0 || (mode && strcmp(mode, "nccl") == 0)) return OMG_OK;
    int nslab = h->first_replicated, nslot = 3 * nslab;
    CUDA_TRY(cudaMalloc((void **)&P.flags, sizeof(unsigned long long) * 4 * nslot));
    CUDA_TRY(cudaMalloc((void **)&P.epochs, sizeof(unsigned long long) * nslot));
    CUDA_TRY(cudaMemset(P.flags, 0, sizeof(unsigned long long) * 4 * nslot));
    CUDA_TRY(cudaMemset(P.epochs, 0, sizeof(unsigned long long) * nslot));
    CUDA_TRY(cudaMalloc((void **)&P.pull, sizeof(unsigned long long) * 8 * nslot));
    CUDA_TRY(cudaMemset(P.pull, 0, sizeof(unsigned long long) * 8 * nslot));
    CUDA_TRY(cudaMalloc((void **)&P.pull_timeout, sizeof(int)));
    CUDA_TRY(cudaMemset(P.pull_timeout, 0, sizeof(int)));
    int nh = 2 + nslot;
    std::vector<cudaIpcMemHandle_t> mine(nh), all((size_t)nh * g.nranks);
    CUDA_TRY(cudaIpcGetMemHandle(&mine[0], P.flags));
    for (int l = 0; l < nslab; ++l) {
        CUDA_TRY(cudaIpcGetMemHandle(&mine[1 + 3 * l], h->lv[l].xa_base));
        CUDA_TRY(cudaIpcGetMemHandle(&mine[2 + 3 * l], h->lv[l].xb_base));
        CUDA_TRY(cudaIpcGetMemHandle(&mine[3 + 3 * l], h->lv[l].b_base));
    }
    CUDA_TRY(cudaIpcGetMemHandle(&mine[1 + nslot], P.pull));
    OMG_TRY(allgather_bytes(mine.data(), all.data(), sizeof(cudaIpcMemHandle_t) * nh));
    P.base_dn.assign(nslot, nullptr);
    P.base_up.assign(nslot, nullptr);
    for (int side = 0; side < 2; ++side) {
        int peer = side == 0 ? g.rank - 1 : g.rank + 1;
        if (peer < 0 || peer >= g.nranks) continue;
        cudaIpcMemHandle_t *ph = &all[(size_t)peer * nh];
        void *p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, ph[0], cudaIpcMemLazyEnablePeerAccess));
        P.opened.push_back(p);
        (side == 0 ? P.flags_dn : P.flags_up) = (unsigned long long *)p;
        for (int s = 0; s < nslot; ++s) {
            CUDA_TRY(cudaIpcOpenMemHandle(&p, ph[1 + s], cudaIpcMemLazyEnablePeerAccess));
            P.opened.push_back(p);
            (side == 0 ? P.base_dn : P.base_up)[s] = (double *)p;
        }
        CUDA_TRY(cudaIpcOpenMemHandle(&p, ph[1 + nslot], cudaIpcMemLazyEnablePeerAccess));
        P.opened.push_back(p);
        (side == 0 ? P.pull_dn : P.pull_up) = (unsigned long long *)p;
    }
    // nobody may start signalling before everyone has mapped and zeroed: a barrier over NCCL
    double *tok = nullptr;
    CUDA_TRY(cudaMalloc(&tok, sizeof(double)));
    CUDA_TRY(cudaMemset(tok, 0, sizeof(double)));
    NCCL_TRY(nccl.AllReduce(tok, tok, 1, ncclFloat64, ncclSum, (ncclComm_t)g.nccl_comm, g.stream2));
    CUDA_TRY(cudaStreamSynchronize(g.stream2));
    cudaFree(tok);
    P.enabled = true;
    P.pull_ok = getenv("OMG_NO_HALO_PULL") == nullptr;
    return OMG_OK;
}

void dist_peer_teardown(omg_hierarchy *h) {
    PeerState &P = h->peer;
    for (void *p : P.opened) cudaIpcCloseMemHandle(p);
    P.opened.clear();
    if (P.flags) cudaFree(P.flags);
    if (P.epochs) cudaFree(P.epochs);
    if (P.pull) cudaFree(P.pull);
    if (P.pull_timeout) cudaFree(P.pull_timeout);
    P.pull = nullptr;
    P.pull_timeout = nullptr;
    P.pull_ok = false;
    P.flags = P.epochs = nullptr;
    P.enabled = false;
}

// Start filling the halos of vector v (owned pointer) of slab level L: hw elements from each
// neighbour.  Asynchronous: dist_halo_wait() makes the compute stream wait for it.
static int halo_exchange_now(omg_hierarchy *h, Level &L, double *v);

// With the fused pull available the exchange is only REQUESTED here: a 3-D stencil kernel that consumes the vector
// reads its boundary planes straight from the neighbours (omg_stencil.cu), any other consumer reaches
// dist_halo_wait, which issues the pending requests on the comm stream first.
int dist_halo_exchange(omg_hierarchy *h, Level &L, double *v) {
    if (g.nranks == 1 || !L.slab) return OMG_OK;
    if (h->peer.pull_ok && peer_slot(L, 0, v) >= 0) {
        for (auto &r : h->halo_req)
            if (r.first == &L && r.second == v) return OMG_OK;
        h->halo_req.push_back({&L, v});
        return OMG_OK;
    }
    return halo_exchange_now(h, L, v);
}

int dist_halo_flush(omg_hierarchy *h) {
    std::vector<std::pair<Level *, double *>> req;
    req.swap(h->halo_req);
    for (auto &r : req) OMG_TRY(halo_exchange_now(h, *r.first, r.second));
    return OMG_OK;
}

// peer pointers for the in-kernel pull of vector v (owned-row-0 pointer) of slab level L
bool dist_pull_params(omg_hierarchy *h, Level &L, const double *v, HaloPull *out) {
    PeerState &P = h->peer;
    if (!P.pull_ok || !L.slab) return false;
    int l = (int)(&L - h->lv.data());
    int slot = peer_slot(L, l, v);
    if (slot < 0) return false;
    bool has_dn = g.rank > 0, has_up = g.rank < g.nranks - 1;
    out->mine = P.pull + 8 * slot;
    out->peer_dn = has_dn ? P.pull_dn + 8 * slot : nullptr;
    out->peer_up = has_up ? P.pull_up + 8 * slot : nullptr;
    out->v_dn = has_dn ? P.base_dn[slot] + L.pad : nullptr;
    out->v_up = has_up ? P.base_up[slot] + L.pad : nullptr;
    out->timeout = P.pull_timeout;
    return true;
}

int dist_pull_timed_out(omg_hierarchy *h, bool *timed_out) {
    *timed_out = false;
    if (!h->peer.pull_timeout) return OMG_OK;
    int f = 0;
    CUDA_TRY(cudaMemcpy(&f, h->peer.pull_timeout, sizeof(int), cudaMemcpyDeviceToHost));
    *timed_out = f != 0;
    return OMG_OK;
}

static int halo_exchange_now(omg_hierarchy *h, Level &L, double *v) {
    OMG_TRY(comm_fork());
    if (h->peer.enabled && peer_slot(L, 0, v) >= 0) {
        ProfScope ps(h, "halo_exchange", (int)(&L - h->lv.data()), 0.0, g.stream2);
        OMG_TRY(peer_exchange(h, L, v));
        CUDA_TRY(cudaEventRecord(ev_join, g.stream2));
        h->halo_pending = true;
        return OMG_OK;
    }
    ProfScope ps(h, "halo_exchange", (int)(&L - h->lv.data()), 0.0, g.stream2);
    ncclComm_t comm = (ncclComm_t)g.nccl_comm;
    size_t hw = (size_t)L.halo;
    NCCL_TRY(nccl.GroupStart());
    if (g.rank > 0) {
        NCCL_TRY(nccl.Send(v, hw, ncclFloat64, g.rank - 1, comm, g.stream2));                  // my bottom rows -> lower
        NCCL_TRY(nccl.Recv(v - hw, hw, ncclFloat64, g.rank - 1, comm, g.stream2));             // lower's top rows
    }
    if (g.rank < g.nranks - 1) {
        NCCL_TRY(nccl.Send(v + L.nloc - hw, hw, ncclFloat64, g.rank + 1, comm, g.stream2));    // my top rows -> upper
        NCCL_TRY(nccl.Recv(v + L.nloc, hw, ncclFloat64, g.rank + 1, comm, g.stream2));         // upper's bottom rows
    }
    NCCL_TRY(nccl.GroupEnd());
    CUDA_TRY(cudaEventRecord(ev_join, g.stream2));
    h->halo_pending = true;
    h->launches++;
    return OMG_OK;
}

// the compute stream waits for every exchange started so far
int dist_halo_wait(omg_hierarchy *h) {
    if (!h->halo_req.empty()) OMG_TRY(dist_halo_flush(h));
    if (!h->halo_pending) return OMG_OK;
    CUDA_TRY(cudaStreamWaitEvent(g.stream, ev_join, 0));
    h->halo_pending = false;
    return OMG_OK;
}

// full[0..n) on every rank from the equal-sized pieces (piece of rank r at full + r*count)
int dist_allgather(omg_hierarchy *h, const double *piece, double *full, size_t count) {
    if (g.nranks == 1) return OMG_OK;
    OMG_TRY(dist_halo_wait(h));
    OMG_TRY(comm_fork());
    {
        ProfScope ps(h, "allgather", -1, 0.0, g.stream2);
        NCCL_TRY(nccl.AllGather(piece, full, count, ncclFloat64, (ncclComm_t)g.nccl_comm, g.stream2));
    }
    CUDA_TRY(cudaEventRecord(ev_join, g.stream2));
    CUDA_TRY(cudaStreamWaitEvent(g.stream, ev_join, 0));
    h->launches++;
    return OMG_OK;
}

int dist_allreduce_sum(omg_hierarchy *h, double *v, size_t count) {
    if (g.nranks == 1) return OMG_OK;
    OMG_TRY(dist_halo_wait(h));
    OMG_TRY(comm_fork());
    NCCL_TRY(nccl.AllReduce(v, v, count, ncclFloat64, ncclSum, (ncclComm_t)g.nccl_comm, g.stream2));
    CUDA_TRY(cudaEventRecord(ev_join, g.stream2));
    CUDA_TRY(cudaStreamWaitEvent(g.stream, ev_join, 0));
    h->launches++;
    return OMG_OK;
}

void dist_finalize() {
    if (g.nccl_comm && nccl.CommDestroy) nccl.CommDestroy((ncclComm_t)g.nccl_comm);
    g.nccl_comm = nullptr;
}
