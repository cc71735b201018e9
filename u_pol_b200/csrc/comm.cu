// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference has no distributed code at all; this is new: rows are
// sharded by molecule blocks, columns replicated, so one evaluation needs an all-reduce of the
// small energy/scalar pack and an all-gather of the row-sliced vectors (d or grad).
//
// NCCL is bound at run time with dlopen("libnccl.so.2") - the library PyTorch already loaded in
// the same process - so libupol_b200.so has no link-time NCCL dependency and loads on a
// CPU-only host for the symbol checks.
#include <dlfcn.h>

#include <algorithm>
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
    void *ptrs[] = {s->xchg_pack, s->xchg_flags, s->d_peer_x, s->d_peer_pack, s->d_peer_flags,
                    s->xg, s->xe, s->xd, s->xflags2, s->d_peer_g, s->d_peer_e, s->d_peer_r, s->d_peer_d, s->d_peer_flags2};
    for (void *p : ptrs) if (p) dfree(p);      // xd may come from the guarded allocator (single-rank sharded host path)
    s->xchg_pack = nullptr; s->xchg_flags = nullptr;
    s->d_peer_x = nullptr; s->d_peer_pack = nullptr; s->d_peer_flags = nullptr;
    s->xg = nullptr; s->xe = nullptr; s->xd = nullptr; s->xflags2 = nullptr;
    s->d_peer_g = nullptr; s->d_peer_e = nullptr; s->d_peer_r = nullptr; s->d_peer_d = nullptr; s->d_peer_flags2 = nullptr;
    s->p2p = false;
}

int comm_site_range(const upol_system *s, int rank, int64_t *site_lo, int64_t *site_hi) {
    if (s->nranks <= 1) { *site_lo = 0; *site_hi = s->nsite; return UPOL_OK; }
    const CommState *c = static_cast<const CommState *>(s->comm);
    if (!c || rank < 0 || rank >= s->nranks) return UPOL_ERR_ARG;
    *site_lo = c->site_lo[rank]; *site_hi = c->site_hi[rank];
    return UPOL_OK;
}

namespace {

constexpr int kNumIpc = 8;   // handles a rank exports: dc, xchg_pack, xchg_flags, xg, xe, xd, r, xflags2

// All-gather by peer stores of the rank's site slice [e0, e1) (in doubles) of r (optional) and of the (nsite,3)
// displacements: every element goes to the same offset of every peer's array; the block that finishes last
// raises the rank's arrival epoch (flag slot 1) on every peer.
__global__ void __launch_bounds__(256)
push_slices_kernel(int64_t e0, int64_t e1, int with_r, const double *__restrict__ r, const double *__restrict__ d,
                   double **peer_r, double **peer_d, int **peer_flags, int *my_flags, int rank, int nranks, int epoch) {
    for (int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += (int64_t)gridDim.x * blockDim.x) {
        const double dv = d[e];
        const double rv = with_r ? r[e] : 0.0;
        for (int g = 0; g < nranks; ++g) {
            if (g == rank) continue;
            peer_d[g][e] = dv;
            if (with_r) peer_r[g][e] = rv;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    int *ticket = my_flags + 3 * nranks;
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *ticket = 0;
        __threadfence_system();
        if ((int)threadIdx.x < nranks && (int)threadIdx.x != rank)
            *((volatile int *)(peer_flags[threadIdx.x] + nranks + rank)) = epoch;
    }
}

__global__ void wait_slices_kernel(const int *my_flags, int rank, int nranks, int epoch, double *poison) {
    bool ok = true;
    if ((int)threadIdx.x < nranks && (int)threadIdx.x != rank) {
        const volatile int *f = my_flags + nranks + threadIdx.x;
        const long long t0 = clock64();
        while (*f < epoch) {
            if (clock64() - t0 > 60000000000LL) { ok = false; break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    // a dead peer must not pass silently: its slice of d becomes NaN, and so does everything computed from it
    if (!ok && poison != nullptr) poison[0] = nan("");
}

}  // namespace

int p2p_allgather_r_d(upol_system *s, bool with_r, cudaStream_t st) {
    if (s->nranks <= 1) return UPOL_OK;
    if (!s->p2p) {
        int rc = with_r ? comm_allgather_sites(s, s->r, st) : UPOL_OK;
        if (rc) return rc;
        return comm_allgather_sites(s, s->xd, st);
    }
    int64_t lo = 0, hi = 0;
    int rc = comm_site_range(s, s->rank, &lo, &hi);
    if (rc) return rc;
    const int epoch = ++s->epoch_push;
    const int64_t n = 3 * (hi - lo);
    const unsigned nb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(148 * 4, (n + 255) / 256));
    push_slices_kernel<<<nb, 256, 0, st>>>(3 * lo, 3 * hi, with_r ? 1 : 0, s->r, s->xd, s->d_peer_r, s->d_peer_d,
                                           s->d_peer_flags2, s->xflags2, s->rank, s->nranks, epoch);
    wait_slices_kernel<<<1, 32, 0, st>>>(s->xflags2, s->rank, s->nranks, epoch, s->xd);
    s->launches += 2;
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
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
    if (guard_enabled()) return UPOL_ERR_ARG;                                  // guarded buffers cannot be IPC-exported (guard.cu)
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    p2p_destroy(s);
    const size_t nflag = 2 * (size_t)s->nranks + 1, nflag2 = 3 * (size_t)s->nranks + 4;
    const size_t ng = 2 * 3 * (size_t)std::max<int64_t>(s->nd, 1), ne = 2 * (size_t)s->nranks * UPOL_E_SIZE;
    const size_t nx = 3 * (size_t)s->nsite;
    cudaError_t e = cudaMalloc((void **)&s->xchg_pack, sizeof(double) * (size_t)s->nranks * kMaxPack);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->xchg_flags, sizeof(int) * nflag);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->xg, sizeof(double) * ng);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->xe, sizeof(double) * ne);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->xd, sizeof(double) * nx);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->xflags2, sizeof(int) * nflag2);
    if (e == cudaSuccess) e = cudaMemset(s->xchg_pack, 0, sizeof(double) * (size_t)s->nranks * kMaxPack);
    if (e == cudaSuccess) e = cudaMemset(s->xchg_flags, 0, sizeof(int) * nflag);
    if (e == cudaSuccess) e = cudaMemset(s->xg, 0, sizeof(double) * ng);
    if (e == cudaSuccess) e = cudaMemset(s->xe, 0, sizeof(double) * ne);
    if (e == cudaSuccess) e = cudaMemset(s->xd, 0, sizeof(double) * nx);
    if (e == cudaSuccess) e = cudaMemset(s->xflags2, 0, sizeof(int) * nflag2);
    cudaIpcMemHandle_t h[kNumIpc];
    void *exported[kNumIpc] = {s->dc, s->xchg_pack, s->xchg_flags, s->xg, s->xe, s->xd, s->r, s->xflags2};
    for (int k = 0; k < kNumIpc && e == cudaSuccess; ++k) e = cudaIpcGetMemHandle(&h[k], exported[k]);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); return UPOL_ERR_CUDA; }
    static_assert(kNumIpc * sizeof(cudaIpcMemHandle_t) == UPOL_P2P_EXPORT_BYTES, "blob size");
    memcpy(blob, h, sizeof(h));
    s->p2p_epoch = 0; s->epoch_eval = 0; s->epoch_push = 0;
    return UPOL_OK;
}

extern "C" int upol_system_p2p_attach(upol_system *s, const unsigned char *all) try {
    if (!s || !all || s->nranks < 2 || !s->xchg_pack) return UPOL_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    const int nr = s->nranks;
    // tab[k][r]: rank r's k-th exported buffer as seen from this device (own buffers for r = rank)
    std::vector<std::vector<void *>> tab(kNumIpc, std::vector<void *>(nr, nullptr));
    void *own[kNumIpc] = {s->dc, s->xchg_pack, s->xchg_flags, s->xg, s->xe, s->xd, s->r, s->xflags2};
    cudaError_t e = cudaSuccess;
    for (int r = 0; r < nr && e == cudaSuccess; ++r) {
        if (r == s->rank) { for (int k = 0; k < kNumIpc; ++k) tab[k][r] = own[k]; continue; }
        cudaIpcMemHandle_t h[kNumIpc];
        memcpy(h, all + (size_t)r * UPOL_P2P_EXPORT_BYTES, sizeof(h));
        for (int k = 0; k < kNumIpc && e == cudaSuccess; ++k) {
            void *q = nullptr;
            e = cudaIpcOpenMemHandle(&q, h[k], cudaIpcMemLazyEnablePeerAccess);
            if (e == cudaSuccess) s->ipc_opened.push_back(q);
            tab[k][r] = q;
        }
    }
    void **dst[kNumIpc] = {(void **)&s->d_peer_x, (void **)&s->d_peer_pack, (void **)&s->d_peer_flags, (void **)&s->d_peer_g,
                           (void **)&s->d_peer_e, (void **)&s->d_peer_d, (void **)&s->d_peer_r, (void **)&s->d_peer_flags2};
    for (int k = 0; k < kNumIpc && e == cudaSuccess; ++k) {
        e = cudaMalloc(dst[k], sizeof(void *) * nr);
        if (e == cudaSuccess) e = cudaMemcpy(*dst[k], tab[k].data(), sizeof(void *) * nr, cudaMemcpyHostToDevice);
    }
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
