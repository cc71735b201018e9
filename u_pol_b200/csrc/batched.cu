// K6: batched independent small systems - one WARP per system, dense BFGS in shared memory.
//
// Replaces the subprocess-per-dimer loop of U_pol/wrapper.py:265-316 (each iteration starts a new
// interpreter, re-traces JAX and runs drudeOpt on an 18-site dimer) with ONE launch over the
// whole conformer batch (BASELINE config 5: 10 k imidazole dimers, N = 18, N_D = 10, 30 dof).
// All systems share one topology (charges, spring constants, molecule ids, screened pairs) and
// differ in core positions.  Per system: the shells' field from the other molecules' charges
// (same formulas as pair_kernels.cu), Thole intra pairs and springs (as finalize_rows_kernel),
// dense inverse-Hessian BFGS started from the inverse intra-molecular block Hessian (springs +
// Thole dipole tensors; the reference's jaxopt.BFGS uses H0 = I, polarization.py:178), the same derivative-only line search and ||grad||_2 <= tol stop rule as
// minimize.cu.  Everything of a system lives in the shared memory slice of its warp; no global
// traffic besides reading r and writing d, the energy vector and the iteration count.
#include <algorithm>
#include <new>
#include <vector>

#include <mutex>

#include "common.cuh"
#include "thole.cuh"

namespace upol {

namespace {

constexpr int kMaxLs = 12;

struct BatchTopo {          // device pointers, shared by all systems
    int nsite, nd, npairs;  // sites, Drude rows, ordered Thole pairs
    const double *qc;       // [nsite]
    const double *row_qs;   // [nd]
    const double *row_k;    // [nd]
    const int *row_site;    // [nd]
    const int *site_mol;    // [nsite]
    const int *th_ptr;      // [nd+1]
    const int *th_row;      // [npairs]
    const double *th_u;     // [npairs]
};

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-warp working set inside dynamic shared memory.
struct WarpMem {
    double *r;      // [nsite*3] core positions
    double *x;      // [n] current displacements (compact)
    double *g;      // [n]
    double *p;      // [n]
    double *xt;     // [n] trial point
    double *gt;     // [n]
    double *hy;     // [n]
    double *H;      // [n*n] inverse Hessian
};

// Field (and optionally potential) at the shells of one system: lane = Drude row (nd <= 32), each
// lane loops over all columns, so the sum order is fixed (deterministic).
template <bool POT>
__device__ void eval_system(const BatchTopo &t, const WarpMem &m, const double *xcur, double *gout,
                            double *e_inter, double *e_intra, double *e_self, int lane) {
    const int nd = t.nd, nsite = t.nsite;
    // dimer workload: 10 rows x 28 columns, ~140 inter-molecular interactions per evaluation
    double ei = 0.0, ea = 0.0, es = 0.0;
    if (lane < nd) {
        const int sa = t.row_site[lane];
        const int ma = t.site_mol[sa];
        const double dx0 = xcur[3 * lane], dy0 = xcur[3 * lane + 1], dz0 = xcur[3 * lane + 2];
        const double sx = m.r[3 * sa] - dx0, sy = m.r[3 * sa + 1] - dy0, sz = m.r[3 * sa + 2] - dz0;
        double pa = 0.0, pb = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
        for (int j = 0; j < nsite; ++j) {             // core columns
            if (t.site_mol[j] == ma) continue;        // polarization.py:90-91
            const double ddx = m.r[3 * j] - sx, ddy = m.r[3 * j + 1] - sy, ddz = m.r[3 * j + 2] - sz;
            const double r2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (r2 < kZeroR2) continue;               // zero distance: dropped (:33-42; rule in thole.cuh)
            const double inv = rsqrt(r2), tq = t.qc[j] * inv, w = tq * inv * inv;
            if (POT) pa += tq;
            fx += w * ddx; fy += w * ddy; fz += w * ddz;
        }
        for (int b = 0; b < nd; ++b) {                // shell columns
            const int sb = t.row_site[b];
            if (t.site_mol[sb] == ma) continue;
            const double ddx = (m.r[3 * sb] - xcur[3 * b]) - sx, ddy = (m.r[3 * sb + 1] - xcur[3 * b + 1]) - sy,
                         ddz = (m.r[3 * sb + 2] - xcur[3 * b + 2]) - sz;
            const double r2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (r2 < kZeroR2) continue;
            const double inv = rsqrt(r2), tq = t.row_qs[b] * inv, w = tq * inv * inv;
            if (POT) pb += tq;
            fx += w * ddx; fy += w * ddy; fz += w * ddz;
        }
        const double qs = t.row_qs[lane], kk = t.row_k[lane];
        double gx = -kCoulomb * qs * fx + kk * dx0, gy = -kCoulomb * qs * fy + kk * dy0, gz = -kCoulomb * qs * fz + kk * dz0;
        if (POT) { ei = kCoulomb * qs * (pa + 0.5 * pb); es = 0.5 * kk * (dx0 * dx0 + dy0 * dy0 + dz0 * dz0); }
        for (int q = t.th_ptr[lane]; q < t.th_ptr[lane + 1]; ++q) {       // polarization.py:94-105
            const int b = t.th_row[q], sb = t.row_site[b];
            const double u = t.th_u[q];
            const double Rx = m.r[3 * sb] - m.r[3 * sa], Ry = m.r[3 * sb + 1] - m.r[3 * sa + 1], Rz = m.r[3 * sb + 2] - m.r[3 * sa + 2];
            const double bx = xcur[3 * b], by = xcur[3 * b + 1], bz = xcur[3 * b + 2];
            const double Ax = Rx + dx0, Ay = Ry + dy0, Az = Rz + dz0;
            const double Cx = Ax - bx, Cy = Ay - by, Cz = Az - bz;
            double gA, cA, gC, cC;
            thole_g(Ax, Ay, Az, u, gA, cA);
            thole_g(Cx, Cy, Cz, u, gC, cC);
            const double qq = kCoulomb * qs * t.row_qs[b];
            if (POT) {
                double gR, cR, gB, cB;
                thole_g(Rx, Ry, Rz, u, gR, cR);
                thole_g(Rx - bx, Ry - by, Rz - bz, u, gB, cB);
                ea += 0.5 * qq * (gR - gA - gB + gC);
            }
            gx += qq * (cC * Cx - cA * Ax); gy += qq * (cC * Cy - cA * Ay); gz += qq * (cC * Cz - cA * Az);
        }
        gout[3 * lane] = gx; gout[3 * lane + 1] = gy; gout[3 * lane + 2] = gz;
    }
    if (POT) { *e_inter = warp_sum(ei); *e_intra = warp_sum(ea); *e_self = warp_sum(es); }
    __syncwarp();
}

// d-independent part: -K sum_rows qs_i sum_{j not in mol(i)} (qc_j + qs_j/2)/|r_j - r_i|
__device__ double static_part(const BatchTopo &t, const WarpMem &m, const double *site_qs, int lane) {
    double e = 0.0;
    if (lane < t.nd) {
        const int sa = t.row_site[lane], ma = t.site_mol[sa];
        double ph = 0.0;
        for (int j = 0; j < t.nsite; ++j) {
            if (t.site_mol[j] == ma) continue;
            const double ddx = m.r[3 * j] - m.r[3 * sa], ddy = m.r[3 * j + 1] - m.r[3 * sa + 1], ddz = m.r[3 * j + 2] - m.r[3 * sa + 2];
            const double r2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (r2 < kZeroR2) continue;
            ph += (t.qc[j] + 0.5 * site_qs[j]) * rsqrt(r2);
        }
        e = -kCoulomb * t.row_qs[lane] * ph;
    }
    return warp_sum(e);
}

__global__ void batched_bfgs_kernel(BatchTopo t, const double *__restrict__ site_qs, int64_t nbatch,
                                    const double *__restrict__ r_all, double *__restrict__ d_all, double tol,
                                    int maxiter, double *__restrict__ energy, int *__restrict__ iters,
                                    int warps_per_block, int per_warp_doubles) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t sys = (int64_t)blockIdx.x * warps_per_block + warp;
    if (sys >= nbatch) return;
    const int nd = t.nd, nsite = t.nsite, n = 3 * nd;
    double *base = smem + (size_t)warp * per_warp_doubles;
    WarpMem m;
    m.r = base; base += 3 * nsite;
    m.x = base; base += n; m.g = base; base += n; m.p = base; base += n;
    m.xt = base; base += n; m.gt = base; base += n; m.hy = base; base += n;
    m.H = base; base += n * n;

    const double *rg = r_all + sys * 3 * nsite;
    double *dg = d_all + sys * 3 * nsite;
    for (int i = lane; i < 3 * nsite; i += 32) m.r[i] = rg[i];
    for (int i = lane; i < n; i += 32) m.x[i] = dg[3 * t.row_site[i / 3] + i % 3];
    // H0 = inverse of the intra-molecular block Hessian at d = 0 (springs + Thole dipole tensors, as
    // minimize.cu::build_blocks_kernel); inverted in place by the warp (Gauss-Jordan, SPD).
    for (int i = lane; i < n * n; i += 32) m.H[i] = 0.0;
    __syncwarp();
    if (lane < nd) {
        const int a = lane, sa = t.row_site[a];
        const double k = t.row_k[a];
        for (int c = 0; c < 3; ++c) m.H[(3 * a + c) * n + 3 * a + c] = k > 0.0 ? k : 1.0;
        for (int q = t.th_ptr[a]; q < t.th_ptr[a + 1]; ++q) {
            const int b = t.th_row[q], sb = t.row_site[b];
            double G[3][3];
            thole_hessian(m.r[3 * sb] - m.r[3 * sa], m.r[3 * sb + 1] - m.r[3 * sa + 1], m.r[3 * sb + 2] - m.r[3 * sa + 2], t.th_u[q], G);
            const double qq = -kCoulomb * t.row_qs[a] * t.row_qs[b];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) m.H[(3 * a + i) * n + 3 * b + j] = qq * G[i][j];
        }
    }
    __syncwarp();
    bool spd = true;
    for (int c = 0; c < n; ++c) {
        const double piv = m.H[c * n + c];
        if (!(piv > 0.0)) { spd = false; break; }          // warp-uniform (shared memory value)
        const double ip = 1.0 / piv;
        __syncwarp();
        for (int j = lane; j < n; j += 32) m.H[c * n + j] = (j == c ? 1.0 : m.H[c * n + j]) * ip;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            if (i == c) continue;
            const double f = m.H[i * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) m.H[i * n + j] = (j == c ? 0.0 : m.H[i * n + j]) - f * m.H[c * n + j];
        }
        __syncwarp();
    }
    if (!spd) {
        for (int i = lane; i < n * n; i += 32) m.H[i] = 0.0;
        __syncwarp();
        for (int i = lane; i < n; i += 32) { const double k = t.row_k[i / 3]; m.H[i * n + i] = k > 0.0 ? 1.0 / k : 1.0; }
        __syncwarp();
    }

    double e0, e1, e2;
    eval_system<false>(t, m, m.x, m.g, &e0, &e1, &e2, lane);
    int it = 0, flags = 0;
    double gnorm = 0.0;
    for (;;) {
        double gg = 0.0;
        for (int i = lane; i < n; i += 32) gg += m.g[i] * m.g[i];
        gg = warp_sum(gg);
        gnorm = sqrt(gg);
        if (!(gg == gg) || isinf(gg)) { flags |= UPOL_FLAG_NONFINITE; break; }
        if (gnorm <= tol) break;
        if (it >= maxiter) { flags |= UPOL_FLAG_NOT_CONVERGED; break; }
        // p = -H g
        double phi0 = 0.0;
        for (int i = lane; i < n; i += 32) {
            double a = 0.0;
            for (int j = 0; j < n; ++j) a += m.H[i * n + j] * m.g[j];
            m.p[i] = -a;
            phi0 += -a * m.g[i];
        }
        phi0 = warp_sum(phi0);
        __syncwarp();
        if (!(phi0 < 0.0)) {       // lost positive definiteness to round-off: restart from H0
            for (int i = lane; i < n * n; i += 32) m.H[i] = 0.0;
            __syncwarp();
            phi0 = 0.0;
            for (int i = lane; i < n; i += 32) {
                const double k = t.row_k[i / 3], h = k > 0.0 ? 1.0 / k : 1.0;
                m.H[i * n + i] = h; m.p[i] = -h * m.g[i]; phi0 += m.p[i] * m.g[i];
            }
            phi0 = warp_sum(phi0);
            __syncwarp();
        }
        // derivative-only line search (same rules as minimize.cu::control_kernel)
        double alpha = 1.0, alo = 0.0, dlo = phi0, ahi = 0.0, dhi = 0.0, dphi = 0.0;
        for (int ls = 0;; ++ls) {
            for (int i = lane; i < n; i += 32) m.xt[i] = m.x[i] + alpha * m.p[i];
            __syncwarp();
            eval_system<false>(t, m, m.xt, m.gt, &e0, &e1, &e2, lane);
            dphi = 0.0;
            for (int i = lane; i < n; i += 32) dphi += m.gt[i] * m.p[i];
            dphi = warp_sum(dphi);
            if ((dphi >= 0.9 * phi0) && (dphi <= -0.8 * phi0)) break;
            if (ls >= kMaxLs) { flags |= UPOL_FLAG_LINESEARCH; break; }
            if (dphi > 0.0) { ahi = alpha; dhi = dphi; } else { alo = alpha; dlo = dphi; }
            double an;
            if (ahi > 0.0) {
                an = alo - dlo * (ahi - alo) / (dhi - dlo);
                const double w = ahi - alo;
                if (!(an > alo + 0.05 * w) || !(an < ahi - 0.05 * w)) an = 0.5 * (alo + ahi);
            } else {
                const double den = phi0 - dphi;
                an = (den < 0.0) ? alpha * phi0 / den : 4.0 * alpha;
                if (!(an > 1.1 * alpha)) an = 2.0 * alpha;
                if (an > 4.0 * alpha) an = 4.0 * alpha;
            }
            alpha = an;
        }
        // BFGS update of the inverse Hessian with s = alpha p, y = gt - g
        double sy = 0.0;
        for (int i = lane; i < n; i += 32) sy += alpha * m.p[i] * (m.gt[i] - m.g[i]);
        sy = warp_sum(sy);
        if (sy > 1e-300) {
            double yhy = 0.0;
            for (int i = lane; i < n; i += 32) {
                double a = 0.0;
                for (int j = 0; j < n; ++j) a += m.H[i * n + j] * (m.gt[j] - m.g[j]);
                m.hy[i] = a;
                yhy += a * (m.gt[i] - m.g[i]);
            }
            yhy = warp_sum(yhy);
            __syncwarp();
            const double rho = 1.0 / sy, c1 = (1.0 + rho * yhy) * rho;
            for (int e = lane; e < n * n; e += 32) {
                const int i = e / n, j = e % n;
                const double si = alpha * m.p[i], sj = alpha * m.p[j];
                m.H[e] += c1 * si * sj - rho * (m.hy[i] * sj + si * m.hy[j]);
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) { m.x[i] = m.xt[i]; m.g[i] = m.gt[i]; }
        __syncwarp();
        ++it;
    }
    // energy at the minimiser
    double ei, ea, es;
    eval_system<true>(t, m, m.x, m.gt, &ei, &ea, &es, lane);
    const double est = static_part(t, m, site_qs, lane);
    for (int i = lane; i < n; i += 32) dg[3 * t.row_site[i / 3] + i % 3] = m.x[i];
    if (lane == 0) {
        double *e = energy + sys * UPOL_E_SIZE;
        e[UPOL_E_INTER] = ei + est; e[UPOL_E_INTRA] = ea; e[UPOL_E_SELF] = es; e[UPOL_E_STATIC] = est;
        e[UPOL_E_GNORM2] = gnorm * gnorm;
        e[UPOL_E_UIND] = ((ei + est) + ea) + es;
        e[6] = (double)flags; e[7] = 0.0;
        if (iters) iters[sys] = it;
    }
}

}  // namespace

}  // namespace upol

using namespace upol;

namespace {
struct TopoEntry {
    int device = -1;
    std::vector<double> hd;
    std::vector<int> hi;
    double *dblob = nullptr;
    int *iblob = nullptr;
    uint64_t stamp = 0;
};
std::mutex g_topo_mutex;
std::vector<TopoEntry> g_topo_cache;
uint64_t g_topo_stamp = 0;
}  // namespace

static int batched_minimize_impl(int device, int64_t nbatch, int64_t nsite, const double *r_core_dev,
                                         double *d_dev, const double *q_core, const double *q_shell,
                                         const double *k, const int32_t *mol_id, int64_t npair,
                                         const int32_t *pair_a, const int32_t *pair_b, const double *pair_u,
                                         double tol, int32_t maxiter, double *energy_dev,
                                         int32_t *iterations_dev, void *stream) {
    if (nbatch < 0 || nsite <= 0 || nsite > 4096 || !r_core_dev || !d_dev || !q_core || !q_shell || !k || !mol_id ||
        !energy_dev || npair < 0)
        return UPOL_ERR_ARG;
    if (nbatch == 0) return UPOL_OK;
    // device switch undone on EVERY exit path (early CUDA errors included)
    struct Scope {
        int prev = -1, dev = -1;
        ~Scope() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
    } scope;
    scope.dev = device;
    cudaGetDevice(&scope.prev);
    if (scope.prev != device && cudaSetDevice(device) != cudaSuccess) { set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__); return UPOL_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;

    // host topology -> compact tables
    std::vector<int> row_site, site_row(nsite, -1), site_mol(nsite);
    std::vector<double> row_qs, row_k, site_qs(q_shell, q_shell + nsite);
    int mol = -1;
    for (int64_t i = 0; i < nsite; ++i) {
        if (i == 0 || mol_id[i] != mol_id[i - 1]) ++mol;
        site_mol[i] = mol;
        if (q_shell[i] != 0.0) { site_row[i] = (int)row_site.size(); row_site.push_back((int)i); row_qs.push_back(q_shell[i]); row_k.push_back(k[i]); }
    }
    const int nd = (int)row_site.size();
    if (nd > 32) return UPOL_ERR_ARG;       // one lane per Drude row
    std::vector<std::vector<std::pair<int, double>>> adj(std::max(nd, 1));
    for (int64_t p = 0; p < npair; ++p) {
        const int a = pair_a[p], b = pair_b[p];
        if (a < 0 || b < 0 || a >= nsite || b >= nsite || a == b || site_mol[a] != site_mol[b]) return UPOL_ERR_ARG;
        const int ra = site_row[a], rb = site_row[b];
        if (ra < 0 || rb < 0 || pair_u[p] == 0.0) continue;
        adj[ra].push_back({rb, pair_u[p]});
        adj[rb].push_back({ra, pair_u[p]});
    }
    std::vector<int> th_ptr(nd + 1, 0), th_row;
    std::vector<double> th_u;
    for (int i = 0; i < nd; ++i) {
        std::sort(adj[i].begin(), adj[i].end());
        for (auto &e : adj[i]) { th_row.push_back(e.first); th_u.push_back(e.second); }
        th_ptr[i + 1] = (int)th_row.size();
    }
    // one device blob for the topology
    const size_t nd1 = std::max(nd, 1), np1 = std::max<size_t>(th_row.size(), 1);
    const size_t dbl = nsite + nd1 + nd1 + np1 + nsite, itg = nd1 + nsite + (nd + 1) + np1;
    std::vector<double> hd;
    hd.insert(hd.end(), q_core, q_core + nsite);
    hd.insert(hd.end(), row_qs.begin(), row_qs.end()); hd.resize(nsite + nd1, 0.0);
    hd.insert(hd.end(), row_k.begin(), row_k.end()); hd.resize(nsite + 2 * nd1, 0.0);
    hd.insert(hd.end(), th_u.begin(), th_u.end()); hd.resize(nsite + 2 * nd1 + np1, 0.0);
    hd.insert(hd.end(), site_qs.begin(), site_qs.end());
    std::vector<int> hi;
    hi.insert(hi.end(), row_site.begin(), row_site.end()); hi.resize(nd1, 0);
    hi.insert(hi.end(), site_mol.begin(), site_mol.end());
    hi.insert(hi.end(), th_ptr.begin(), th_ptr.end());
    hi.insert(hi.end(), th_row.begin(), th_row.end()); hi.resize(itg, 0);
    // The topology tables live on the device across calls (a scan driver calls this once per batch of conformers
    // with the same force field): a small cache keyed by device and table contents; on a hit nothing is allocated
    // or copied.  Entries are only freed on eviction (cudaFree waits for kernels that may still read them).
    double *dblob = nullptr;
    int *iblob = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_topo_mutex);
        for (auto &c : g_topo_cache)
            if (c.device == device && c.hd == hd && c.hi == hi) { dblob = c.dblob; iblob = c.iblob; c.stamp = ++g_topo_stamp; break; }
        if (!dblob) {
            if (g_topo_cache.size() >= 8) {
                size_t old = 0;
                for (size_t q = 1; q < g_topo_cache.size(); ++q) if (g_topo_cache[q].stamp < g_topo_cache[old].stamp) old = q;
                upol::dfree(g_topo_cache[old].dblob); upol::dfree(g_topo_cache[old].iblob);
                g_topo_cache.erase(g_topo_cache.begin() + old);
            }
            TopoEntry c;
            c.device = device;
            UPOL_CUDA(upol::dmalloc((void **)&c.dblob, dbl * sizeof(double)));
            if (upol::dmalloc((void **)&c.iblob, itg * sizeof(int)) != cudaSuccess) { upol::dfree(c.dblob); set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__); return UPOL_ERR_CUDA; }
            cudaError_t ce = cudaMemcpy(c.dblob, hd.data(), dbl * sizeof(double), cudaMemcpyHostToDevice);
            if (ce == cudaSuccess) ce = cudaMemcpy(c.iblob, hi.data(), itg * sizeof(int), cudaMemcpyHostToDevice);
            if (ce == cudaSuccess) ce = cudaDeviceSynchronize();      // pageable copies on the legacy stream: landed before any stream reads them
            if (ce != cudaSuccess) { upol::dfree(c.dblob); upol::dfree(c.iblob); set_last_cuda_error(ce, __FILE__, __LINE__); return UPOL_ERR_CUDA; }
            c.hd = hd; c.hi = hi; c.stamp = ++g_topo_stamp;
            dblob = c.dblob; iblob = c.iblob;
            g_topo_cache.push_back(std::move(c));
        }
    }
    BatchTopo t;
    t.nsite = (int)nsite; t.nd = nd; t.npairs = (int)th_row.size();
    t.qc = dblob; t.row_qs = dblob + nsite; t.row_k = dblob + nsite + nd1; t.th_u = dblob + nsite + 2 * nd1;
    const double *site_qs_dev = dblob + nsite + 2 * nd1 + np1;
    t.row_site = iblob; t.site_mol = iblob + nd1; t.th_ptr = iblob + nd1 + nsite; t.th_row = iblob + nd1 + nsite + nd + 1;

    const int n = 3 * nd;
    const int per_warp = (int)(3 * nsite + 6 * n + (size_t)n * n);
    const size_t per_warp_bytes = (size_t)per_warp * sizeof(double);
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (per_warp_bytes > (size_t)max_smem) return UPOL_ERR_ARG;
    int wpb = (int)std::min<size_t>(8, std::max<size_t>(1, (size_t)(48 * 1024) / per_warp_bytes));
    const size_t smem = per_warp_bytes * wpb;
    if (smem > 48 * 1024)
        UPOL_CUDA(cudaFuncSetAttribute(batched_bfgs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((nbatch + wpb - 1) / wpb);
    batched_bfgs_kernel<<<blocks, 32 * wpb, smem, st>>>(t, site_qs_dev, nbatch, r_core_dev, d_dev, tol, maxiter,
                                                       energy_dev, iterations_dev, wpb, per_warp);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);     // documented: returns after the batch has finished
    if (e != cudaSuccess) { set_last_cuda_error(e, __FILE__, __LINE__); return UPOL_ERR_CUDA; }
    return UPOL_OK;
}

// C ABI: no C++ exception may cross the boundary
extern "C" int upol_batched_minimize_dev(int device, int64_t nbatch, int64_t nsite, const double *r_core_dev,
                                         double *d_dev, const double *q_core, const double *q_shell,
                                         const double *k, const int32_t *mol_id, int64_t npair,
                                         const int32_t *pair_a, const int32_t *pair_b, const double *pair_u,
                                         double tol, int32_t maxiter, double *energy_dev,
                                         int32_t *iterations_dev, void *stream) {
    try {
        return batched_minimize_impl(device, nbatch, nsite, r_core_dev, d_dev, q_core, q_shell, k, mol_id, npair, pair_a,
                                     pair_b, pair_u, tol, maxiter, energy_dev, iterations_dev, stream);
    } catch (const std::bad_alloc &) {
        return UPOL_ERR_ALLOC;
    } catch (...) {
        return UPOL_ERR_ARG;
    }
}
