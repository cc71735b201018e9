// extern "C" entry points of libupol_b200.so (see include/upol_b200.h for the contract and the
// reference lines each one replaces).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "common.cuh"

using namespace upol;

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

template <typename T>
int dev_alloc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = upol::dmalloc((void **)p, n * sizeof(T));
    if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); return UPOL_ERR_ALLOC; }
    return UPOL_OK;
}

void free_system(upol_system *s) {
    if (!s) return;
    minimize_free(s);
    ewald_free(s);
    comm_destroy(s);
    void *ptrs[] = {s->r, s->qc, s->qs, s->k, s->row_site, s->row_qs, s->row_k, s->exA_lo, s->exA_len,
                    s->exB_lo, s->exB_len, s->rowx, s->rowy, s->rowz, s->srowx, s->srowy, s->srowz,
                    s->cols, s->cols_static, s->th_ptr, s->th_row, s->th_u, s->part, s->dc,
                    s->gc, s->eblock, s->energy, s->e_static, s->phi_static, s->site_row,
                    s->cols_dcore, s->site_x, s->exS_lo, s->cf_tmp,
                    s->cols_x, s->cols_d, s->cols_d2, s->col_qeff, s->ticket, s->site_nslot, s->exP_lo, s->exP_len, s->exN_lo, s->exN_len};
    for (void *p : ptrs) if (p) upol::dfree(p);
    if (s->pin) cudaFreeHost(s->pin);
    if (s->hscratch) upol::dfree(s->hscratch);
    for (cudaEvent_t e : s->tev) cudaEventDestroy(e);
    delete s;
}

// true if `p` is page-locked host memory (then the library copies straight from / to it)
bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

void compute_centre(upol_system *s, const double *r_host) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < s->nsite; ++i)
        for (int c = 0; c < 3; ++c) {
            lo[c] = std::min(lo[c], r_host[3 * i + c]);
            hi[c] = std::max(hi[c], r_host[3 * i + c]);
        }
    set_box(s, lo, hi);
}

}  // namespace

extern "C" {

const char *upol_error_string(int code) {
    switch (code) {
        case UPOL_OK: return "ok";
        case UPOL_ERR_CUDA: return last_error_message();
        case UPOL_ERR_ARG: return "invalid argument";
        case UPOL_ERR_NCCL: return "NCCL error or NCCL library not available";
        case UPOL_ERR_ALLOC: return last_error_message();
        default: return code > 0 ? "numerical flag (see UPOL_FLAG_*)" : "unknown error";
    }
}

int upol_version(void) { return 100; }

double upol_one_4pi_eps0(void) { return kCoulomb; }

int upol_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int upol_system_create_impl(upol_system **out, int device, int64_t nsite, const double *r_core,
                       const double *q_core, const double *q_shell, const double *k,
                       const int32_t *mol_id, int64_t npair, const int32_t *pair_a,
                       const int32_t *pair_b, const double *pair_u) {
    if (!out || nsite <= 0 || !r_core || !q_core || !q_shell || !k || !mol_id || npair < 0) return UPOL_ERR_ARG;
    if (nsite > (int64_t)1 << 30) return UPOL_ERR_ARG;
    *out = nullptr;
    DeviceGuard guard(device);
    if (!guard.ok) { set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__); return UPOL_ERR_CUDA; }

    upol_system *s = new (std::nothrow) upol_system();
    if (!s) return UPOL_ERR_ALLOC;
    s->device = device;
    s->nsite = nsite;
    s->h_mol.assign(mol_id, mol_id + nsite);
    s->h_qc.assign(q_core, q_core + nsite);
    s->h_qs.assign(q_shell, q_shell + nsite);
    s->h_k.assign(k, k + nsite);

    // molecules: contiguous runs of equal mol_id
    s->h_site_row.assign(nsite, -1);
    std::vector<int32_t> site_mol(nsite);
    for (int64_t i = 0; i < nsite; ++i) {
        if (i > 0 && mol_id[i] < mol_id[i - 1]) { delete s; return UPOL_ERR_ARG; }
        if (i == 0 || mol_id[i] != mol_id[i - 1]) {
            s->h_mol_site0.push_back((int32_t)i);
            s->h_mol_nsite.push_back(0);
            s->h_mol_row0.push_back((int32_t)s->h_drude_site.size());
            s->h_mol_nrow.push_back(0);
        }
        const int m = (int)s->h_mol_site0.size() - 1;
        site_mol[i] = m;
        s->h_mol_nsite[m]++;
        if (q_shell[i] != 0.0) {
            s->h_site_row[i] = (int32_t)s->h_drude_site.size();
            s->h_drude_site.push_back((int32_t)i);
            s->h_mol_nrow[m]++;
        }
    }
    s->h_mol = site_mol;   // dense molecule index from here on
    s->nmol = (int64_t)s->h_mol_site0.size();
    s->nd = (int64_t)s->h_drude_site.size();
    s->nd_pad = round_up(std::max<int64_t>(s->nd, 1), 512);
    s->na_pad = round_up(nsite, kTileCols);
    s->nb_pad = round_up(std::max<int64_t>(s->nd, 1), kTileCols);
    s->nn_pad = round_up(std::max<int64_t>(nsite - s->nd, 1), kTileCols);
    s->mol_lo = 0; s->mol_hi = s->nmol; s->row_lo = 0; s->row_hi = s->nd;

    // per-row tables
    std::vector<int32_t> exA_lo(s->nd_pad, 0), exA_len(s->nd_pad, 0), exB_lo(s->nd_pad, 0), exB_len(s->nd_pad, 0);
    std::vector<double> row_qs(std::max<int64_t>(s->nd, 1), 0.0), row_k(std::max<int64_t>(s->nd, 1), 0.0);
    // difference-form mixed kernel: non-polarisable sites get slots of section N in site order
    std::vector<int32_t> site_nslot(nsite, -1), mol_n0(s->nmol, 0), mol_nn(s->nmol, 0);
    {
        int32_t slot = 0;
        for (int64_t i = 0; i < nsite; ++i) {
            const int m = site_mol[i];
            if (i == s->h_mol_site0[m]) mol_n0[m] = slot;
            if (s->h_site_row[i] < 0) { site_nslot[i] = slot++; mol_nn[m]++; }
        }
    }
    std::vector<int32_t> exP_lo(s->nd_pad, 0), exP_len(s->nd_pad, 0), exN_lo(s->nd_pad, 0), exN_len(s->nd_pad, 0);
    for (int64_t i = 0; i < s->nd; ++i) {
        const int site = s->h_drude_site[i];
        const int m = site_mol[site];
        exP_lo[i] = s->h_mol_row0[m];
        exP_len[i] = s->h_mol_nrow[m];
        exN_lo[i] = (int32_t)s->nb_pad + mol_n0[m];              // concatenated column index
        exN_len[i] = mol_nn[m];
        exA_lo[i] = s->h_mol_site0[m];
        exA_len[i] = s->h_mol_nsite[m];
        exB_lo[i] = (int32_t)s->na_pad + s->h_mol_row0[m];   // concatenated column index
        exB_len[i] = s->h_mol_nrow[m];
        row_qs[i] = q_shell[site];
        row_k[i] = k[site];
    }
    // screened pairs -> CSR of ordered pairs over Drude rows
    std::vector<int32_t> th_ptr(s->nd + 1, 0), th_row;
    std::vector<double> th_u;
    {
        std::vector<std::vector<std::pair<int32_t, double>>> adj(s->nd);
        for (int64_t p = 0; p < npair; ++p) {
            const int a = pair_a[p], b = pair_b[p];
            if (a < 0 || b < 0 || a >= nsite || b >= nsite || a == b || site_mol[a] != site_mol[b]) { delete s; return UPOL_ERR_ARG; }
            const int ra = s->h_site_row[a], rb = s->h_site_row[b];
            if (ra < 0 || rb < 0 || pair_u[p] == 0.0) continue;   // S == 0 for a non-polarisable partner / u = 0
            adj[ra].push_back({rb, pair_u[p]});
            adj[rb].push_back({ra, pair_u[p]});
        }
        for (int64_t i = 0; i < s->nd; ++i) {
            std::sort(adj[i].begin(), adj[i].end());
            for (auto &e : adj[i]) { th_row.push_back(e.first); th_u.push_back(e.second); }
            th_ptr[i + 1] = (int32_t)th_row.size();
        }
    }
    s->npair_ordered = (int64_t)th_row.size();

    // Column splits S are chosen per launch so that (row blocks * S) ~ 148*40 (choose_split);
    // with rb row blocks of <= 512 rows, S*stride <= (5920/rb + 1)(512 rb + 511), which the
    // allocation below covers for the full system and for every row shard of it ...
    s->max_split = 256;
    // ... and a launch never has more splits than column tiles, which bounds small systems tightly
    const size_t rows_pad = (size_t)std::max<int64_t>(s->nd_pad, round_up(nsite, 512));
    const size_t tiles_max = (size_t)((s->na_pad + s->nb_pad) / kTileCols);
    const size_t part_elems = (size_t)kPartComps * std::min<size_t>((size_t)148 * 64 * 1024 + rows_pad + 1024,
                                                                    std::min<size_t>(tiles_max, 256) * rows_pad);

    int rc = UPOL_OK;
#define TRY(x) do { rc = (x); if (rc) { free_system(s); return rc; } } while (0)
    TRY(dev_alloc(&s->r, 3 * nsite));
    TRY(dev_alloc(&s->qc, nsite)); TRY(dev_alloc(&s->qs, nsite)); TRY(dev_alloc(&s->k, nsite));
    TRY(dev_alloc(&s->row_site, s->nd)); TRY(dev_alloc(&s->row_qs, s->nd)); TRY(dev_alloc(&s->row_k, s->nd));
    TRY(dev_alloc(&s->site_row, nsite));
    TRY(dev_alloc(&s->exA_lo, s->nd_pad)); TRY(dev_alloc(&s->exA_len, s->nd_pad));
    TRY(dev_alloc(&s->exB_lo, s->nd_pad)); TRY(dev_alloc(&s->exB_len, s->nd_pad));
    TRY(dev_alloc(&s->rowx, s->nd_pad)); TRY(dev_alloc(&s->rowy, s->nd_pad)); TRY(dev_alloc(&s->rowz, s->nd_pad));
    TRY(dev_alloc(&s->srowx, s->nd_pad)); TRY(dev_alloc(&s->srowy, s->nd_pad)); TRY(dev_alloc(&s->srowz, s->nd_pad));
    TRY(dev_alloc(&s->cols, s->na_pad + s->nb_pad));
    TRY(dev_alloc(&s->cols_static, s->na_pad)); TRY(dev_alloc(&s->col_qeff, s->na_pad));
    TRY(dev_alloc(&s->cols_x, s->nb_pad + s->nn_pad)); TRY(dev_alloc(&s->cols_d, s->nb_pad)); TRY(dev_alloc(&s->cols_d2, s->nb_pad));
    TRY(dev_alloc(&s->site_nslot, nsite));
    TRY(dev_alloc(&s->exP_lo, s->nd_pad)); TRY(dev_alloc(&s->exP_len, s->nd_pad));
    TRY(dev_alloc(&s->exN_lo, s->nd_pad)); TRY(dev_alloc(&s->exN_len, s->nd_pad));
    TRY(dev_alloc(&s->th_ptr, s->nd + 1)); TRY(dev_alloc(&s->th_row, th_row.size())); TRY(dev_alloc(&s->th_u, th_u.size()));
    TRY(dev_alloc(&s->part, part_elems));
    s->part_elems = part_elems;
    TRY(dev_alloc(&s->dc, 3 * s->nd)); TRY(dev_alloc(&s->gc, 3 * s->nd));
    s->n_eblock = (int)((std::max<int64_t>(s->nd, nsite) + 7) / 8) + 1;   // finalize blocks hold 32 rows (8 for small systems)
    TRY(dev_alloc(&s->eblock, 5 * (size_t)s->n_eblock));   // kEblockComps per finalize block
    TRY(dev_alloc(&s->ticket, 4));
    TRY(dev_alloc(&s->energy, UPOL_E_SIZE)); TRY(dev_alloc(&s->e_static, 2)); TRY(dev_alloc(&s->phi_static, s->nd_pad));

    {
        cudaError_t e = cudaSuccess;
        auto cp = [&](void *dst, const void *src, size_t bytes) {
            if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
        };
        cp(s->r, r_core, sizeof(double) * 3 * nsite);
        cp(s->qc, q_core, sizeof(double) * nsite);
        cp(s->qs, q_shell, sizeof(double) * nsite);
        cp(s->k, k, sizeof(double) * nsite);
        cp(s->row_site, s->h_drude_site.data(), sizeof(int32_t) * s->nd);
        cp(s->site_row, s->h_site_row.data(), sizeof(int32_t) * nsite);
        cp(s->row_qs, row_qs.data(), sizeof(double) * s->nd);
        cp(s->row_k, row_k.data(), sizeof(double) * s->nd);
        cp(s->exA_lo, exA_lo.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->exA_len, exA_len.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->exB_lo, exB_lo.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->exB_len, exB_len.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->site_nslot, site_nslot.data(), sizeof(int32_t) * nsite);
        cp(s->exP_lo, exP_lo.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->exP_len, exP_len.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->exN_lo, exN_lo.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->exN_len, exN_len.data(), sizeof(int32_t) * s->nd_pad);
        cp(s->th_ptr, th_ptr.data(), sizeof(int32_t) * (s->nd + 1));
        cp(s->th_row, th_row.data(), sizeof(int32_t) * th_row.size());
        cp(s->th_u, th_u.data(), sizeof(double) * th_u.size());
        if (e == cudaSuccess) e = cudaMemset(s->ticket, 0, sizeof(int) * 4);
        if (e == cudaSuccess) e = cudaMemset(s->rowx, 0, sizeof(double) * s->nd_pad);
        if (e == cudaSuccess) e = cudaMemset(s->rowy, 0, sizeof(double) * s->nd_pad);
        if (e == cudaSuccess) e = cudaMemset(s->rowz, 0, sizeof(double) * s->nd_pad);
        if (e == cudaSuccess) e = cudaMemset(s->srowx, 0, sizeof(double) * s->nd_pad);
        if (e == cudaSuccess) e = cudaMemset(s->srowy, 0, sizeof(double) * s->nd_pad);
        if (e == cudaSuccess) e = cudaMemset(s->srowz, 0, sizeof(double) * s->nd_pad);
        if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); free_system(s); return UPOL_ERR_CUDA; }
    }
    compute_centre(s, r_core);
    TRY(sys_upload_positions(s, nullptr, 0));
    if (cudaDeviceSynchronize() != cudaSuccess) { set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__); free_system(s); return UPOL_ERR_CUDA; }
#undef TRY
    *out = s;
    return UPOL_OK;
}

int upol_system_destroy(upol_system *s) {
    if (!s) return UPOL_OK;
    DeviceGuard guard(s->device);
    cudaDeviceSynchronize();
    free_system(s);
    return UPOL_OK;
}

int upol_system_set_positions(upol_system *s, const double *r_core, void *stream) {
    if (!s || !r_core) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    // stage through pinned memory so the copy is a true async DMA; the caller's buffer is free on return
    const size_t n3 = 3 * (size_t)s->nsite;
    int rc = sys_ensure_pinned(s, sizeof(double) * (2 * n3 + UPOL_E_SIZE));
    if (rc) return rc;
    const double *src = r_core;
    if (!is_pinned(r_core)) {                             // page-locked caller buffers are copied from directly
        UPOL_CUDA(cudaStreamSynchronize(st));             // staging buffer may still be in flight
        memcpy(s->pin, r_core, sizeof(double) * n3);
        src = s->pin;
    }
    UPOL_CUDA(cudaMemcpyAsync(s->r, src, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
    // box centre / fixed-point grid from the device copy (one block, 6 doubles back) instead of a host loop over
    // all sites; the synchronisation inside also makes the caller's / the staging buffer reusable on return
    rc = sys_box_from_device(s, st);
    if (rc) return rc;
    return sys_upload_positions(s, nullptr, st);
}

int upol_system_set_positions_dev(upol_system *s, const double *r_core_dev, void *stream) {
    if (!s || !r_core_dev) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    UPOL_CUDA(cudaMemcpyAsync(s->r, r_core_dev, sizeof(double) * 3 * s->nsite, cudaMemcpyDeviceToDevice, st));
    // box centre / fixed-point grid follow the new geometry (6 doubles come back to the host)
    int rc = sys_box_from_device(s, st);
    if (rc) return rc;
    return sys_upload_positions(s, nullptr, st);
}

int64_t upol_system_nsite(const upol_system *s) { return s ? s->nsite : 0; }
int64_t upol_system_ndrude(const upol_system *s) { return s ? s->nd : 0; }

int64_t upol_system_interactions(const upol_system *s, int static_part) {
    if (!s) return 0;
    int64_t total = 0;
    for (int64_t m = s->mol_lo; m < s->mol_hi; ++m) {
        const int64_t rows = s->h_mol_nrow[m];
        const int64_t cols = static_part ? (s->nsite - s->h_mol_nsite[m])
                                         : (s->nsite - s->h_mol_nsite[m]) + (s->nd - s->h_mol_nrow[m]);
        total += rows * cols;
    }
    return total;
}

int upol_system_set_row_range(upol_system *s, int64_t mol_lo, int64_t mol_hi) {
    if (!s || mol_lo < 0 || mol_hi > s->nmol || mol_lo > mol_hi) return UPOL_ERR_ARG;
    s->mol_lo = mol_lo; s->mol_hi = mol_hi;
    s->row_lo = mol_lo < s->nmol ? s->h_mol_row0[mol_lo] : s->nd;
    s->row_hi = mol_hi < s->nmol ? s->h_mol_row0[mol_hi] : s->nd;
    s->static_valid = false;
    return UPOL_OK;
}

int upol_system_periodic_info(upol_system *s, double *box, double *kappa, int32_t *nk) {
    if (!s) return UPOL_ERR_ARG;
    int n = 0;
    int rc = ewald_info(s, box, kappa, &n);
    if (rc == UPOL_OK && nk) *nk = n;
    return rc;
}

int upol_uind_energy_grad_dev(upol_system *s, const double *d, int precision, double *energy,
                              double *grad, void *stream) {
    if (!s || !d || (!energy && !grad)) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    EvalOpts o;
    o.precision = precision;
    o.want_grad = grad != nullptr;
    o.want_energy = energy != nullptr;      // NULL: field pass only (what a minimiser iteration needs)
    o.d_full = d;                           // gathered into s->dc by the prepare kernel
    // Sharded runs with the peer mapping in place: the finalize kernel stores its gradient rows and the rank's
    // energy partials into every peer (fused compute + all-gather / all-reduce over NVLink), no NCCL call
    const bool px = s->p2p && s->nranks > 1;
    double *gbuf = s->gc;
    if (px) { o.gx = make_eval_xchg(s, grad != nullptr); gbuf = s->xg + o.gx.g_off; }
    // one rank evaluating all its rows: the finalize kernel writes the (nsite,3) layout itself, no scatter launch
    const bool direct = grad != nullptr && s->nranks == 1 && s->row_lo == 0 && s->row_hi == s->nd && s->nd > 0;
    if (direct) o.g_full = grad;
    int rc = sys_eval_compact(s, s->dc, o, energy, gbuf, st);
    if (rc) return rc;
    if (grad && !direct) {
        if (s->nranks > 1 && !px) { rc = comm_allgather_rows(s, gbuf, st); if (rc) return rc; }
        rc = sys_scatter_g(s, gbuf, grad, st);
    }
    return rc;
}

// Sharded host path: see the header.
static int upol_uind_energy_grad_host_sharded_impl(upol_system *s, const double *r_core, const double *d, int precision,
                                                   double *energy, double *grad, int64_t *site_lo_out, int64_t *site_hi_out) {
    if (!s || !d || (!energy && !grad)) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    int64_t lo = 0, hi = s->nsite;
    int rc = comm_site_range(s, s->rank, &lo, &hi);
    if (rc) return rc;
    if (site_lo_out) *site_lo_out = lo;
    if (site_hi_out) *site_hi_out = hi;
    const size_t n3 = 3 * (size_t)s->nsite, m3 = 3 * (size_t)(hi - lo), o3 = 3 * (size_t)lo;
    rc = sys_ensure_pinned(s, sizeof(double) * (3 * n3 + UPOL_E_SIZE));      // staging for r, d, grad slices + energy
    if (rc) return rc;
    rc = sys_ensure_hscratch(s, 2 * n3 + UPOL_E_SIZE);
    if (rc) return rc;
    if (!s->xd) UPOL_CUDA(upol::dmalloc((void **)&s->xd, sizeof(double) * n3));      // single rank / NCCL-only runs
    double *gfull = s->hscratch + n3, *e_dev = gfull + n3;
    double *pr = s->pin, *pd = s->pin + m3, *pg = s->pin + 2 * m3, *pe = s->pin + 3 * n3;
    cudaStream_t st = 0;
    // only this rank's site slices of r and d cross PCIe; the peers' slices arrive over NVLink
    if (r_core) {
        const double *src = r_core + o3;
        if (!is_pinned(r_core)) { memcpy(pr, src, sizeof(double) * m3); src = pr; }
        if (m3) UPOL_CUDA(cudaMemcpyAsync(s->r + o3, src, sizeof(double) * m3, cudaMemcpyHostToDevice, st));
    }
    {
        const double *src = d + o3;
        if (!is_pinned(d)) { memcpy(pd, src, sizeof(double) * m3); src = pd; }
        if (m3) UPOL_CUDA(cudaMemcpyAsync(s->xd + o3, src, sizeof(double) * m3, cudaMemcpyHostToDevice, st));
    }
    rc = p2p_allgather_r_d(s, r_core != nullptr, st);
    if (rc) return rc;
    if (r_core) {
        rc = sys_box_from_device(s, st);
        if (rc) return rc;
        rc = sys_upload_positions(s, nullptr, st);
        if (rc) return rc;
    }
    EvalOpts o;
    o.precision = precision;
    o.want_grad = grad != nullptr;
    o.want_energy = energy != nullptr;
    o.d_full = s->xd;
    if (s->p2p && s->nranks > 1) o.gx = make_eval_xchg(s, false);     // energy partials only: every rank keeps its rows
    rc = sys_eval_compact(s, s->dc, o, e_dev, s->gc, st);
    if (rc) return rc;
    const bool e_pin = energy && is_pinned(energy), g_pin = grad && is_pinned(grad);
    if (energy) UPOL_CUDA(cudaMemcpyAsync(e_pin ? energy : pe, e_dev, sizeof(double) * UPOL_E_SIZE, cudaMemcpyDeviceToHost, st));
    if (grad && m3) {
        rc = sys_scatter_g_range(s, s->gc, gfull, lo, hi, st);
        if (rc) return rc;
        UPOL_CUDA(cudaMemcpyAsync(g_pin ? grad + o3 : pg, gfull + o3, sizeof(double) * m3, cudaMemcpyDeviceToHost, st));
    }
    UPOL_CUDA(cudaStreamSynchronize(st));
    if (energy && !e_pin) memcpy(energy, pe, sizeof(double) * UPOL_E_SIZE);
    if (grad && !g_pin && m3) memcpy(grad + o3, pg, sizeof(double) * m3);
    return UPOL_OK;
}

int upol_uind_energy_grad_host(upol_system *s, const double *d, int precision, double *energy, double *grad) {
    if (!s || !d || (!energy && !grad)) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    const size_t n3 = 3 * (size_t)s->nsite;
    int rc = sys_ensure_pinned(s, sizeof(double) * (3 * n3 + UPOL_E_SIZE));      // staging for r, d, grad slices + energy
    if (rc) return rc;
    rc = sys_ensure_hscratch(s, 2 * n3 + UPOL_E_SIZE);
    if (rc) return rc;
    double *dfull = s->hscratch, *gfull = dfull + n3, *e_dev = gfull + n3;
    // page-locked caller buffers are used directly; pageable ones are staged through the library's own
    const bool d_pin = is_pinned(d), g_pin = grad && is_pinned(grad), e_pin = energy && is_pinned(energy);
    double *pd = s->pin, *pg = s->pin + n3, *pe = pg + n3;
    if (!d_pin) memcpy(pd, d, sizeof(double) * n3);
    cudaStream_t st = 0;
    UPOL_CUDA(cudaMemcpyAsync(dfull, d_pin ? d : pd, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
    rc = upol_uind_energy_grad_dev(s, dfull, precision, energy ? e_dev : nullptr, grad ? gfull : nullptr, st);
    if (rc) return rc;
    if (energy) UPOL_CUDA(cudaMemcpyAsync(e_pin ? energy : pe, e_dev, sizeof(double) * UPOL_E_SIZE, cudaMemcpyDeviceToHost, st));
    if (grad) UPOL_CUDA(cudaMemcpyAsync(g_pin ? grad : pg, gfull, sizeof(double) * n3, cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaStreamSynchronize(st));
    if (energy && !e_pin) memcpy(energy, pe, sizeof(double) * UPOL_E_SIZE);
    if (grad && !g_pin) memcpy(grad, pg, sizeof(double) * n3);
    return UPOL_OK;
}

int64_t upol_system_launch_count(const upol_system *s) { return s ? s->launches : 0; }

int upol_system_set_kernel_timing(upol_system *s, int enable) {
    if (!s) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    if (enable && s->tev.empty()) {
        s->tev.resize(512);
        for (auto &e : s->tev) UPOL_CUDA(cudaEventCreate(&e));
    }
    s->timing = enable != 0;
    s->tev_used = 0;
    return UPOL_OK;
}

int upol_system_kernel_time(upol_system *s, double *ms_total, int32_t *launches) {
    if (!s || !ms_total || !launches) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    double tot = 0.0;
    for (int i = 0; i + 1 < s->tev_used; i += 2) {
        UPOL_CUDA(cudaEventSynchronize(s->tev[i + 1]));
        float ms = 0.f;
        UPOL_CUDA(cudaEventElapsedTime(&ms, s->tev[i], s->tev[i + 1]));
        tot += ms;
    }
    *ms_total = tot;
    *launches = s->tev_used / 2;
    s->tev_used = 0;
    return UPOL_OK;
}

static int upol_system_set_periodic_impl(upol_system *s, const double *box, double kappa, int nmax) {
    if (!s) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    cudaDeviceSynchronize();                       // no evaluation may still be reading the old state
    return ewald_setup(s, box, kappa, nmax);
}

static int upol_uind_core_gradient_dev_impl(upol_system *s, const double *d, double *dudr, void *stream) {
    if (!s || !d || !dudr) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = sys_gather_d(s, d, s->dc, st);
    if (rc) return rc;
    return sys_core_forces(s, s->dc, dudr, st);
}

static int upol_coulomb_static_dev_impl(upol_system *s, double *out2, void *stream) {
    if (!s || !out2 || s->ewald) return UPOL_ERR_ARG;           // open boundaries only
    DeviceGuard guard(s->device);
    return sys_coulomb_static(s, out2, (cudaStream_t)stream);
}

int upol_uself_dev(const double *d, const double *k, int64_t n, double *out, void *stream) {
    if (!d || !k || !out || n < 0) return UPOL_ERR_ARG;
    return sys_uself(d, k, n, out, (cudaStream_t)stream);
}

int upol_make_sij_dev(const double *rij, const double *u, int u_is_scalar, int64_t n, double *out, void *stream) {
    if (!rij || !u || !out || n < 0) return UPOL_ERR_ARG;
    return sys_make_sij(rij, u, u_is_scalar, n, out, (cudaStream_t)stream);
}

void upol_min_options_default(upol_min_options *o) {
    if (!o) return;
    o->tol = 1e-4;          // polarization.py:178
    o->maxiter = 500;       // jaxopt default
    o->method = UPOL_LBFGS;
    o->precision = UPOL_FP64;
    o->history = 8;
    o->check_every = 4;
    o->reserved = 0;
}

static int upol_minimize_dev_impl(upol_system *s, double *d, const upol_min_options *opts, upol_min_result *res, void *stream) {
    if (!s || !d || !res) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    upol_min_options o;
    if (opts) o = *opts; else upol_min_options_default(&o);
    if (s->ewald && o.precision != UPOL_FP64) return UPOL_ERR_ARG;   // periodic boxes: fp64 only
    int rc = sys_gather_d(s, d, s->dc, st);
    if (rc) return rc;
    rc = minimize_compact(s, s->dc, &o, res, st);
    if (rc < 0) return rc;
    int rc2 = sys_scatter_d(s, s->dc, d, st);
    if (rc2) return rc2;
    UPOL_CUDA(cudaStreamSynchronize(st));
    return rc;
}

int upol_minimize_host(upol_system *s, double *d, const upol_min_options *opts, upol_min_result *res) {
    if (!s || !d || !res) return UPOL_ERR_ARG;
    DeviceGuard guard(s->device);
    const size_t n3 = 3 * (size_t)s->nsite;
    int rc = sys_ensure_pinned(s, sizeof(double) * (3 * n3 + UPOL_E_SIZE));      // staging for r, d, grad slices + energy
    if (rc) return rc;
    rc = sys_ensure_hscratch(s, 2 * n3 + UPOL_E_SIZE);
    if (rc) return rc;
    double *dfull = s->hscratch;
    memcpy(s->pin, d, sizeof(double) * n3);
    UPOL_CUDA(cudaMemcpyAsync(dfull, s->pin, sizeof(double) * n3, cudaMemcpyHostToDevice, 0));
    rc = upol_minimize_dev(s, dfull, opts, res, 0);
    if (rc < 0) return rc;
    UPOL_CUDA(cudaMemcpyAsync(s->pin, dfull, sizeof(double) * n3, cudaMemcpyDeviceToHost, 0));
    UPOL_CUDA(cudaStreamSynchronize(0));
    memcpy(d, s->pin, sizeof(double) * n3);
    return rc;
}

int upol_uind_energy_grad_host_sharded(upol_system *s, const double *r_core, const double *d, int precision,
                                       double *energy, double *grad, int64_t *site_lo, int64_t *site_hi) {
    try {
        return upol_uind_energy_grad_host_sharded_impl(s, r_core, d, precision, energy, grad, site_lo, site_hi);
    } catch (const std::bad_alloc &) {
        return UPOL_ERR_ALLOC;
    } catch (...) {
        return UPOL_ERR_ARG;
    }
}

// C ABI: no C++ exception may cross the boundary
int upol_system_create(upol_system **out, int device, int64_t nsite, const double *r_core, const double *q_core,
                       const double *q_shell, const double *k, const int32_t *mol_id, int64_t npair,
                       const int32_t *pair_a, const int32_t *pair_b, const double *pair_u) {
    try {
        return upol_system_create_impl(out, device, nsite, r_core, q_core, q_shell, k, mol_id, npair, pair_a, pair_b, pair_u);
    } catch (const std::bad_alloc &) {
        return UPOL_ERR_ALLOC;
    } catch (...) {
        return UPOL_ERR_ARG;
    }
}

// C ABI: no C++ exception may cross the boundary
int upol_minimize_dev(upol_system *s, double *d, const upol_min_options *opts, upol_min_result *res, void *stream) {
    try {
        return upol_minimize_dev_impl(s, d, opts, res, stream);
    } catch (const std::bad_alloc &) {
        return UPOL_ERR_ALLOC;
    } catch (...) {
        return UPOL_ERR_ARG;
    }
}

// C ABI: no C++ exception may cross the boundary
int upol_uind_core_gradient_dev(upol_system *s, const double *d, double *dudr, void *stream) {
    try {
        return upol_uind_core_gradient_dev_impl(s, d, dudr, stream);
    } catch (const std::bad_alloc &) {
        return UPOL_ERR_ALLOC;
    } catch (...) {
        return UPOL_ERR_ARG;
    }
}

// C ABI: no C++ exception may cross the boundary
int upol_coulomb_static_dev(upol_system *s, double *out2, void *stream) {
    try {
        return upol_coulomb_static_dev_impl(s, out2, stream);
    } catch (const std::bad_alloc &) {
        return UPOL_ERR_ALLOC;
    } catch (...) {
        return UPOL_ERR_ARG;
    }
}

int upol_system_set_periodic(upol_system *s, const double *box, double kappa, int32_t nmax) {
    try {
        return upol_system_set_periodic_impl(s, box, kappa, nmax);
    } catch (const std::bad_alloc &) {
        return UPOL_ERR_ALLOC;
    } catch (...) {
        return UPOL_ERR_ARG;
    }
}

}  // extern "C"
