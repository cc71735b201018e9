// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference has no distributed code at all; this is new: rows are
// sharded by molecule blocks, columns replicated, so one evaluation needs an all-reduce of the
// small energy/scalar pack and an all-gather of the row-sliced vectors (d or grad).
//
// NCCL is bound at run time with dlopen("libnccl.so.2") - the library PyTorch already loaded in
// the same process - so libupol_b200.so has no link-time NCCL dependency and loads on a
// CPU-only host for the symbol checks.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace upol {

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    bool ok = false;
};

NcclApi &api() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) return a;
#define LOAD(field, sym) *(void **)(&a.field) = dlsym(a.handle, sym)
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(AllGather, "ncclAllGather");
    LOAD(Broadcast, "ncclBroadcast");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
#undef LOAD
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather &&
           a.Broadcast && a.GroupStart && a.GroupEnd;
    return a;
}

struct CommState {
    ncclComm_t comm = nullptr;
    std::vector<int64_t> row_lo, row_hi;   // Drude-row range of every rank
    std::vector<int64_t> site_lo, site_hi; // site range of every rank
};

}  // namespace

int comm_allreduce_sum(upol_system *s, double *buf, int64_t n, cudaStream_t st) {
    if (s->nranks <= 1) return UPOL_OK;
    CommState *c = static_cast<CommState *>(s->comm);
    if (!c || !api().ok) return UPOL_ERR_NCCL;
    return api().AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, c->comm, st) == ncclSuccess ? UPOL_OK : UPOL_ERR_NCCL;
}

// In-place gather of per-rank slices [lo[g], hi[g]) * 3 doubles.  Equal slices -> one ncclAllGather
// (in place: every rank's send buffer is its own slice of the receive buffer); uneven slices (molecule
// blocks differ by one molecule) -> a group of broadcasts, which NCCL fuses into one launch.
static int gather_slices(upol_system *s, CommState *c, const std::vector<int64_t> &lo, const std::vector<int64_t> &hi,
                         double *vec3, cudaStream_t st) {
    NcclApi &a = api();
    bool equal = true;
    const int64_t n0 = hi[0] - lo[0];
    for (int g = 0; g < s->nranks; ++g) equal = equal && (hi[g] - lo[g] == n0) && (lo[g] == lo[0] + g * n0);
    if (equal && n0 > 0) {
        double *base = vec3 + 3 * lo[0];
        return a.AllGather(base + 3 * n0 * s->rank, base, (size_t)(3 * n0), ncclFloat64, c->comm, st) == ncclSuccess
                   ? UPOL_OK : UPOL_ERR_NCCL;
    }
    if (a.GroupStart() != ncclSuccess) return UPOL_ERR_NCCL;
    for (int g = 0; g < s->nranks; ++g) {
        const int64_t n = 3 * (hi[g] - lo[g]);
        if (n <= 0) continue;
        double *p = vec3 + 3 * lo[g];
        if (a.Broadcast(p, p, (size_t)n, ncclFloat64, g, c->comm, st) != ncclSuccess) { a.GroupEnd(); return UPOL_ERR_NCCL; }
    }
    return a.GroupEnd() == ncclSuccess ? UPOL_OK : UPOL_ERR_NCCL;
}

// All-gather of a row-sliced [nd*3] vector, in place: rank g owns rows [row_lo[g], row_hi[g]).
int comm_allgather_rows(upol_system *s, double *vec3, cudaStream_t st) {
    if (s->nranks <= 1) return UPOL_OK;
    CommState *c = static_cast<CommState *>(s->comm);
    if (!c || !api().ok) return UPOL_ERR_NCCL;
    return gather_slices(s, c, c->row_lo, c->row_hi, vec3, st);
}

int comm_allgather_sites(upol_system *s, double *vec3, cudaStream_t st) {
    if (s->nranks <= 1) return UPOL_OK;
    CommState *c = static_cast<CommState *>(s->comm);
    if (!c || !api().ok) return UPOL_ERR_NCCL;
    return gather_slices(s, c, c->site_lo, c->site_hi, vec3, st);
}

void p2p_destroy(upol_system *s) {
    for (void *p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    s->ipc_opened.clear();
    void *ptrs[] = {s->xchg_pack, s->xchg_flags, s->d_peer_x, s->d_peer_pack, s->d_peer_flags};
    for (void *p : ptrs) if (p) cudaFree(p);
    s->xchg_pack = nullptr; s->xchg_flags = nullptr;
    s->d_peer_x = nullptr; s->d_peer_pack = nullptr; s->d_peer_flags = nullptr;
    s->p2p = false;
}

void comm_destroy(upol_system *s) {
    p2p_destroy(s);
    CommState *c = static_cast<CommState *>(s->comm);
    if (!c) return;
    if (c->comm && api().ok) api().CommDestroy(c->comm);
    delete c;
    s->comm = nullptr;
    s->nranks = 1; s->rank = 0;
}

// Balanced molecule blocks: rank g gets molecules [g*nmol/G, (g+1)*nmol/G).
void shard_bounds(int64_t nmol, int nranks, int rank, int64_t *lo, int64_t *hi) {
    *lo = nmol * rank / nranks;
    *hi = nmol * (rank + 1) / nranks;
}

}  // namespace upol

using namespace upol;

extern "C" int upol_comm_unique_id(unsigned char *id_bytes) {
    if (!id_bytes) return UPOL_ERR_ARG;
    if (!api().ok) return UPOL_ERR_NCCL;
    ncclUniqueId id;
    if (api().GetUniqueId(&id) != ncclSuccess) return UPOL_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == UPOL_COMM_ID_BYTES, "id size");
    memcpy(id_bytes, &id, UPOL_COMM_ID_BYTES);
    return UPOL_OK;
}

extern "C" int upol_system_attach_comm(upol_system *s, const unsigned char *id_bytes, int rank, int nranks) try {
    if (!s || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return UPOL_ERR_ARG;
    if (!api().ok) return UPOL_ERR_NCCL;
    comm_destroy(s);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    CommState *c = new CommState();
    ncclUniqueId id;
    memcpy(&id, id_bytes, UPOL_COMM_ID_BYTES);
    int rc = api().CommInitRank(&c->comm, nranks, id, rank);
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != ncclSuccess) { delete c; return UPOL_ERR_NCCL; }
    c->row_lo.resize(nranks); c->row_hi.resize(nranks);
    c->site_lo.resize(nranks); c->site_hi.resize(nranks);
    for (int g = 0; g < nranks; ++g) {
        int64_t ml, mh;
        shard_bounds(s->nmol, nranks, g, &ml, &mh);
        c->row_lo[g] = ml < s->nmol ? s->h_mol_row0[ml] : s->nd;
        c->row_hi[g] = mh < s->nmol ? s->h_mol_row0[mh] : s->nd;
        c->site_lo[g] = ml < s->nmol ? s->h_mol_site0[ml] : s->nsite;
        c->site_hi[g] = mh < s->nmol ? s->h_mol_site0[mh] : s->nsite;
    }
    s->comm = c; s->rank = rank; s->nranks = nranks;
    int64_t ml, mh;
    shard_bounds(s->nmol, nranks, rank, &ml, &mh);
    return upol_system_set_row_range(s, ml, mh);
} catch (...) {
    return UPOL_ERR_ALLOC;
}

// ---- peer-memory exchange (CUDA IPC over NVLink) -------------------------------------------------
extern "C" int upol_system_p2p_export(upol_system *s, unsigned char *blob) {
    if (!s || !blob || s->nranks < 2 || s->nranks > 32) return UPOL_ERR_ARG;   // one warp polls the peers' flags
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    p2p_destroy(s);
    const size_t nflag = 2 * (size_t)s->nranks + 1;
    cudaError_t e = cudaMalloc((void **)&s->xchg_pack, sizeof(double) * (size_t)s->nranks * kMaxPack);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->xchg_flags, sizeof(int) * nflag);
    if (e == cudaSuccess) e = cudaMemset(s->xchg_pack, 0, sizeof(double) * (size_t)s->nranks * kMaxPack);
    if (e == cudaSuccess) e = cudaMemset(s->xchg_flags, 0, sizeof(int) * nflag);
    cudaIpcMemHandle_t h[3];
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h[0], s->dc);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h[1], s->xchg_pack);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h[2], s->xchg_flags);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); return UPOL_ERR_CUDA; }
    static_assert(3 * sizeof(cudaIpcMemHandle_t) == UPOL_P2P_EXPORT_BYTES, "blob size");
    memcpy(blob, h, sizeof(h));
    s->p2p_epoch = 0;
    return UPOL_OK;
}

extern "C" int upol_system_p2p_attach(upol_system *s, const unsigned char *all) try {
    if (!s || !all || s->nranks < 2 || !s->xchg_pack) return UPOL_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    const int nr = s->nranks;
    std::vector<double *> px(nr), pp(nr);
    std::vector<int *> pf(nr);
    cudaError_t e = cudaSuccess;
    for (int r = 0; r < nr && e == cudaSuccess; ++r) {
        if (r == s->rank) { px[r] = s->dc; pp[r] = s->xchg_pack; pf[r] = s->xchg_flags; continue; }
        cudaIpcMemHandle_t h[3];
        memcpy(h, all + (size_t)r * UPOL_P2P_EXPORT_BYTES, sizeof(h));
        void *q[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < 3 && e == cudaSuccess; ++k) {
            e = cudaIpcOpenMemHandle(&q[k], h[k], cudaIpcMemLazyEnablePeerAccess);
            if (e == cudaSuccess) s->ipc_opened.push_back(q[k]);
        }
        px[r] = (double *)q[0]; pp[r] = (double *)q[1]; pf[r] = (int *)q[2];
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_peer_x, sizeof(double *) * nr);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_peer_pack, sizeof(double *) * nr);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_peer_flags, sizeof(int *) * nr);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_peer_x, px.data(), sizeof(double *) * nr, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_peer_pack, pp.data(), sizeof(double *) * nr, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_peer_flags, pf.data(), sizeof(int *) * nr, cudaMemcpyHostToDevice);
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); p2p_destroy(s); return UPOL_ERR_CUDA; }
    s->p2p = true;
    return UPOL_OK;
} catch (...) {
    return UPOL_ERR_ALLOC;
}

extern "C" int upol_system_p2p_disable(upol_system *s) {
    if (!s) return UPOL_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    p2p_destroy(s);
    if (prev >= 0) cudaSetDevice(prev);
    return UPOL_OK;
}
