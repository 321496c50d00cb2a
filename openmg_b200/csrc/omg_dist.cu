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

// Start filling the halos of vector v (owned pointer) of slab level L: hw elements from each
// neighbour.  Asynchronous: dist_halo_wait() makes the compute stream wait for it.
int dist_halo_exchange(omg_hierarchy *h, Level &L, double *v) {
    if (g.nranks == 1 || !L.slab) return OMG_OK;
    OMG_TRY(comm_fork());
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
