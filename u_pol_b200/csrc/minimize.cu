// K5: device-resident minimiser over the Drude displacements.
// Replaces drudeOpt = jaxopt.BFGS(fun=Uind, tol=1e-4).run(Dij0) (U_pol/polarization.py:154-188).
//
// jaxopt's dense BFGS keeps a (3N)^2 inverse Hessian and an energy-based zoom line search; both
// stop working at scale (H is 2.3 TB at 100 k Drudes, and near the
// minimum dE ~ g^2/2k is below the round-off of an N^2-term fp64 sum).  This minimiser is
//   * limited-memory BFGS in the compact (Byrd-Nocedal-Schnabel) form with H0 = gamma*P, where P
//     is the inverse of the MOLECULE-BLOCK Hessian at d = 0: springs on the diagonal and the
//     Thole-damped dipole-dipole tensors -K qs_a qs_b grad grad g(R_ab) between the screened pairs
//     of a molecule (the strongest couplings in the system; their block has condition number ~4
//     for the imidazole force field).  Blocks are built and inverted on the device, one thread
//     per molecule.  Against P = diag(1/k) this takes 12 instead of 20 iterations on the liquid
//     boxes.  U_ind is near-quadratic, so the unit step is almost always accepted (one gradient
//     evaluation per iteration);
//   * a line search driven by directional derivatives only: accept when
//     0.9 phi'(0) <= phi'(a) <= -0.8 phi'(0) (approximate Wolfe), otherwise a safeguarded
//     secant step on phi'; energies are never compared;
//   * same stop rule as the reference: ||grad||_2 <= tol.
// Everything - dot products, the m x m solves, the line-search decisions, the updates - runs on
// the device; scalars never visit the host.  Per gradient evaluation ("slot") the stream holds
//   field pass -> dots -> [all-reduce] -> control (1 thread) -> update -> [all-gather d]
// and every kernel returns immediately once the device-side done flag is set.  The host only
// enqueues slots in chunks and polls the flag (one chunk ahead), so no host round-trip happens
// per iteration.  UPOL_CG selects the memoryless variant (history 1), which is the
// Polak-Ribiere-type preconditioned conjugate gradient under exact line searches.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "thole.cuh"

namespace upol {

namespace {

constexpr int kMaxHist = 16;
constexpr int kDotBlock = 256;
constexpr int kDotVals = 8;          // per history slot: 4 used; last row: 6 used
constexpr int kMaxLineSearch = 12;

// integer control words
enum { I_GATE = 0, I_ITER, I_NEVAL, I_HCOUNT, I_HHEAD, I_LS, I_FLAGS, I_ACCEPT, I_FIRST, I_NEWSLOT, I_PUSHED, I_MAXITER, I_TICKET, I_SIZE = 16 };
// double control words
enum { D_ALPHA = 0, D_ALPHA_PREV, D_PHI0, D_ALO, D_DLO, D_AHI, D_DHI, D_GAMMA, D_GNORM, D_GNORM0, D_TOL, D_SWITCH, D_SWITCH_REL, D_SIZE = 16 };

struct MinWs {
    int m = 0;
    int64_t n = 0;             // 3 * nd
    double *x0 = nullptr, *g0 = nullptr, *p = nullptr, *g = nullptr;
    double *Pg = nullptr, *Py = nullptr;     // [n] preconditioned gradient / gradient change
    double *S = nullptr, *Y = nullptr;       // [m][n] history: s_l and P y_l (y_l itself is never needed)
    double *blk = nullptr;                   // inverse molecule blocks, (3 nrow_m)^2 doubles each
    int64_t *blk_off = nullptr;              // [nmol] offset of a molecule's block
    int *mol_row0 = nullptr, *mol_nrow = nullptr, *row_mol = nullptr;
    int64_t blk_elems = 0;
    int64_t nmol = 0;
    int max_rows = 0;                        // Drude rows of the largest molecule
    double *partial = nullptr;               // [nblocks][(m+1)*kDotVals]
    double *pack = nullptr;                  // [(m+1)*kDotVals]
    double *sc = nullptr;                    // [D_SIZE]
    int *ic = nullptr;                       // [I_SIZE]
    double *coef = nullptr;                  // top[m], u[m] by PHYSICAL slot
    double *RS = nullptr, *YY = nullptr;     // [m*m] chronological
    int nblocks = 0;
    int *h_ic = nullptr;                     // pinned
    double *h_sc = nullptr;                  // pinned
    double *e_final = nullptr;               // [UPOL_E_SIZE]
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // single-rank fast path: a chunk of `check_every` slots captured once into a CUDA graph and replayed
    cudaStream_t gstream = nullptr;          // capture needs a non-legacy stream
    cudaEvent_t gev[2] = {nullptr, nullptr};
    cudaGraphExec_t gexec = nullptr;
    long long gkey[6] = {-1, -1, -1, -1, -1, -1};   // (history | precision, chunk length, row range, x, geometry generation)
    bool graph_failed = false;
    int64_t glaunches = 0;                   // kernels one replay of the graph launches
};

void ws_free(MinWs *w) {
    if (!w) return;
    void *ptrs[] = {w->x0, w->g0, w->p, w->g, w->Pg, w->Py, w->S, w->Y, w->partial, w->pack, w->sc, w->ic,
                    w->coef, w->RS, w->YY, w->e_final, w->blk, w->blk_off, w->mol_row0, w->mol_nrow, w->row_mol};
    for (void *q : ptrs) if (q) upol::dfree(q);
    if (w->h_ic) cudaFreeHost(w->h_ic);
    if (w->h_sc) cudaFreeHost(w->h_sc);
    for (auto &e : w->ev) if (e) cudaEventDestroy(e);
    for (auto &e : w->gev) if (e) cudaEventDestroy(e);
    if (w->gexec) cudaGraphExecDestroy(w->gexec);
    if (w->gstream) cudaStreamDestroy(w->gstream);
    delete w;
}

// One thread per molecule: block Hessian of (springs + Thole intra pairs) at d = 0, inverted in
// place (Gauss-Jordan on an SPD matrix).  H_aa = k_a I,  H_ab = -K qs_a qs_b grad grad g(R_ab)
// (only the g(C) term of polarization.py:94-99 couples d_a and d_b; the others cancel at d = 0).
// A non-positive pivot (would mean an intra-molecular polarisation catastrophe) falls back to
// the diagonal 1/k.
__global__ void build_blocks_kernel(int mol_lo, int mol_hi, const int *__restrict__ mol_row0,
                                    const int *__restrict__ mol_nrow, const int64_t *__restrict__ blk_off,
                                    const int *__restrict__ row_site, const double *__restrict__ r,
                                    const double *__restrict__ row_qs, const double *__restrict__ row_k,
                                    const int *__restrict__ th_ptr, const int *__restrict__ th_row,
                                    const double *__restrict__ th_u, double *__restrict__ blk) {
    const int m = mol_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= mol_hi) return;
    const int r0 = mol_row0[m], nr = mol_nrow[m], nb = 3 * nr;
    if (nr == 0) return;
    double *A = blk + blk_off[m];
    for (int i = 0; i < nb * nb; ++i) A[i] = 0.0;
    bool coupled = false;
    for (int a = 0; a < nr; ++a) {
        const double k = row_k[r0 + a];
        for (int c = 0; c < 3; ++c) A[(3 * a + c) * nb + 3 * a + c] = k > 0.0 ? k : 1.0;
        const int sa = row_site[r0 + a];
        for (int q = th_ptr[r0 + a]; q < th_ptr[r0 + a + 1]; ++q) {
            const int b = th_row[q] - r0, sb = row_site[th_row[q]];
            double G[3][3];
            thole_hessian(r[3 * sb] - r[3 * sa], r[3 * sb + 1] - r[3 * sa + 1], r[3 * sb + 2] - r[3 * sa + 2], th_u[q], G);
            const double qq = -kCoulomb * row_qs[r0 + a] * row_qs[r0 + b];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) A[(3 * a + i) * nb + 3 * b + j] = qq * G[i][j];
            coupled = true;
        }
    }
    bool ok = true;
    if (coupled) {
        for (int c = 0; c < nb && ok; ++c) {          // in-place Gauss-Jordan inverse
            const double piv = A[c * nb + c];
            if (!(piv > 0.0)) { ok = false; break; }
            const double ip = 1.0 / piv;
            A[c * nb + c] = 1.0;
            for (int j = 0; j < nb; ++j) A[c * nb + j] *= ip;
            for (int i = 0; i < nb; ++i) {
                if (i == c) continue;
                const double f = A[i * nb + c];
                if (f == 0.0) continue;
                A[i * nb + c] = 0.0;
                for (int j = 0; j < nb; ++j) A[i * nb + j] -= f * A[c * nb + j];
            }
        }
    }
    if (!coupled || !ok) {
        for (int i = 0; i < nb * nb; ++i) A[i] = 0.0;
        for (int a = 0; a < nr; ++a) {
            const double k = row_k[r0 + a];
            for (int c = 0; c < 3; ++c) A[(3 * a + c) * nb + 3 * a + c] = k > 0.0 ? 1.0 / k : 1.0;
        }
    }
}

// The same, one WARP per molecule with the block in shared memory (3 nrow <= 32): lane j owns column j, so a
// Gauss-Jordan sweep is 3 nrow row operations instead of (3 nrow)^2 element operations of one thread in global
// memory (1.16 ms -> tens of microseconds for 20 000 imidazoles).  Identical arithmetic per element, identical result.
constexpr int kBlkWarps = 4;
__global__ void __launch_bounds__(32 * kBlkWarps)
build_blocks_warp_kernel(int mol_lo, int mol_hi, const int *__restrict__ mol_row0,
                         const int *__restrict__ mol_nrow, const int64_t *__restrict__ blk_off,
                         const int *__restrict__ row_site, const double *__restrict__ r,
                         const double *__restrict__ row_qs, const double *__restrict__ row_k,
                         const int *__restrict__ th_ptr, const int *__restrict__ th_row,
                         const double *__restrict__ th_u, double *__restrict__ blk) {
    __shared__ double sm[kBlkWarps][32][33];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = mol_lo + blockIdx.x * kBlkWarps + wid;
    if (m >= mol_hi) return;
    const int r0 = mol_row0[m], nr = mol_nrow[m], nb = 3 * nr;
    if (nr == 0) return;
    double (*A)[33] = sm[wid];
    for (int i = 0; i < nb; ++i) if (lane < nb) A[i][lane] = 0.0;
    __syncwarp();
    int coupled = 0;
    if (lane < nr) {                                   // lane a fills the three rows of Drude a
        const int a = lane;
        const double k = row_k[r0 + a];
        for (int c = 0; c < 3; ++c) A[3 * a + c][3 * a + c] = k > 0.0 ? k : 1.0;
        const int sa = row_site[r0 + a];
        for (int q = th_ptr[r0 + a]; q < th_ptr[r0 + a + 1]; ++q) {
            const int b = th_row[q] - r0, sb = row_site[th_row[q]];
            double G[3][3];
            thole_hessian(r[3 * sb] - r[3 * sa], r[3 * sb + 1] - r[3 * sa + 1], r[3 * sb + 2] - r[3 * sa + 2], th_u[q], G);
            const double qq = -kCoulomb * row_qs[r0 + a] * row_qs[r0 + b];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) A[3 * a + i][3 * b + j] = qq * G[i][j];
            coupled = 1;
        }
    }
    coupled = __any_sync(0xffffffffu, coupled);
    __syncwarp();
    bool ok = true;
    if (coupled) {
        for (int c = 0; c < nb; ++c) {                 // in-place Gauss-Jordan inverse, lane = column
            const double piv = A[c][c];
            if (!(piv > 0.0)) { ok = false; break; }   // uniform across the warp (shared value)
            const double ip = 1.0 / piv;
            __syncwarp();
            if (lane == c) A[c][c] = 1.0;
            __syncwarp();
            if (lane < nb) A[c][lane] *= ip;
            __syncwarp();
            const double pc = lane < nb ? A[c][lane] : 0.0;
            for (int i = 0; i < nb; ++i) {
                if (i == c) continue;
                const double f = A[i][c];
                __syncwarp();
                if (f == 0.0) continue;
                if (lane == c) A[i][c] = 0.0;
                __syncwarp();
                if (lane < nb) A[i][lane] -= f * pc;
                __syncwarp();
            }
        }
    }
    double *out = blk + blk_off[m];
    if (!coupled || !ok) {
        for (int i = 0; i < nb; ++i)
            if (lane < nb) {
                const double k = row_k[r0 + i / 3];
                out[i * nb + lane] = (i == lane) ? (k > 0.0 ? 1.0 / k : 1.0) : 0.0;
            }
    } else {
        for (int i = 0; i < nb; ++i) if (lane < nb) out[i * nb + lane] = A[i][lane];
    }
}

// Pg = P g and Py = P (g - g0) over the local slice (block mat-vec, blocks are tiny).
__global__ void __launch_bounds__(256)
precond_apply_kernel(int64_t e0, int64_t e1, const int *__restrict__ ic, const int *__restrict__ row_mol,
                     const int *__restrict__ mol_row0, const int *__restrict__ mol_nrow,
                     const int64_t *__restrict__ blk_off, const double *__restrict__ blk,
                     const double *__restrict__ g, const double *__restrict__ g0,
                     double *__restrict__ Pg, double *__restrict__ Py) {
    if (ic[I_GATE] & kGateDone) return;
    const int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e1) return;
    const int m = row_mol[e / 3];
    const int64_t b0 = 3 * (int64_t)mol_row0[m];
    const int nb = 3 * mol_nrow[m];
    const double *row = blk + blk_off[m] + (e - b0) * nb;
    double a = 0.0, b = 0.0;
    for (int j = 0; j < nb; ++j) {
        const double gj = g[b0 + j];
        a += row[j] * gj;
        b += row[j] * (gj - g0[b0 + j]);
    }
    Pg[e] = a; Py[e] = b;
}

// Peer-memory exchange (NVLink P2P, tables built by upol_system_p2p_attach).  epoch = id of the slot.
struct P2P {
    int on, rank, nranks, epoch;
    double **peer_x;        // every rank's displacement vector
    double **peer_pack;     // every rank's mailbox [nranks][kMaxPack]
    int **peer_flags;       // every rank's arrival epochs [0,nranks): packs, [nranks,2nranks): d slices, [2nranks]: ticket
    double *my_pack;        // this rank's mailbox
    int *my_flags;
};

// Spin until flags[i] >= epoch (written by peer i through NVLink); ~2 s time-out instead of a hang.
__device__ __forceinline__ bool wait_epoch(const int *flags, int i, int epoch) {
    const volatile int *f = flags + i;
    const long long t0 = clock64();
    while (*f < epoch) {
        if (clock64() - t0 > 4000000000LL) return false;
        __nanosleep(200);
    }
    __threadfence_system();
    return true;
}

// Before a field pass: all peers' slices of d (sent by their update kernels of the previous slot) are in.
__global__ void wait_x_kernel(int *ic, P2P pp, int epoch_prev) {
    if (ic[I_GATE] & kGateDone) return;
    bool ok = true;
    if ((int)threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank) ok = wait_epoch(pp.my_flags, pp.nranks + threadIdx.x, epoch_prev);
    if (!__all_sync(0xffffffffu, ok) && threadIdx.x == 0) { ic[I_FLAGS] |= UPOL_FLAG_COMM; ic[I_GATE] |= kGateDone; }
}

__global__ void init_state_kernel(int *ic, double *sc, double tol, int phase_bits, int maxiter, double switch_rel) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < I_SIZE; ++i) ic[i] = 0;
    ic[I_MAXITER] = maxiter;
    for (int i = 0; i < D_SIZE; ++i) sc[i] = 0.0;
    ic[I_GATE] = phase_bits;
    ic[I_FIRST] = 1;
    sc[D_ALPHA] = 0.0;      // the first slot evaluates x itself
    sc[D_TOL] = tol;
    sc[D_SWITCH_REL] = switch_rel;
    sc[D_GAMMA] = 1.0;
}

// Dot products of one slot over the local row slice [e0, e1).
// blockIdx.y = l < m : history slot l (physical): s_l.y, y_l.P.y, s_l.g, y_l.P.g      (y = g - g0)
// blockIdx.y = m     : p.y, y.P.y, p.g, y.P.g, g.g, g.P.g
__global__ void __launch_bounds__(kDotBlock)
dots_kernel(int m, int64_t n, int64_t e0, int64_t e1, const int *__restrict__ ic,
            const double *__restrict__ S, const double *__restrict__ Y, const double *__restrict__ g,
            const double *__restrict__ g0, const double *__restrict__ p, const double *__restrict__ Pg,
            const double *__restrict__ Py, double *__restrict__ partial) {
    if (ic[I_GATE] & kGateDone) return;
    __shared__ double sm[kDotVals][kDotBlock / 32];
    const int l = blockIdx.y;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    {   // history slots not yet filled hold zeros, so every slot can be summed unconditionally
        const double *a = l < m ? S + (int64_t)l * n : p;
        const double *b = l < m ? Y + (int64_t)l * n : nullptr;
        for (int64_t e = e0 + (int64_t)blockIdx.x * kDotBlock + threadIdx.x; e < e1; e += (int64_t)gridDim.x * kDotBlock) {
            const double ge = g[e], ye = ge - g0[e];
            if (l < m) {
                const double se = a[e], yl = b[e];          // b holds P y_l (P symmetric)
                acc[0] += se * ye; acc[1] += yl * ye; acc[2] += se * ge; acc[3] += yl * ge;
            } else {
                const double ae = a[e], pye = Py[e];
                acc[0] += ae * ye; acc[1] += ye * pye; acc[2] += ae * ge; acc[3] += pye * ge;
                acc[4] += ge * ge; acc[5] += ge * Pg[e];
            }
        }
    }
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double v = acc[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (ln == 0) sm[c][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int k = 0; k < kDotBlock / 32; ++k) t += sm[threadIdx.x][k];
        partial[((int64_t)blockIdx.x * (m + 1) + l) * kDotVals + threadIdx.x] = t;
    }
}

// One block.  Without peers: pack = sum of the block partials.  With peers: the local sums go into
// every rank's mailbox slot of this rank over NVLink, then the arrival flag of this rank is raised on
// every peer (fence, then flag); the controller below adds the mailboxes in rank order.
__device__ __forceinline__ void reduce_partials(int m, int nblocks, const double *__restrict__ partial,
                                                double *__restrict__ pack, const P2P &pp) {
    const int idx = threadIdx.x, len = (m + 1) * kDotVals;
    double t = 0.0;
    if (idx < len && (idx % kDotVals) < 6)
        for (int b = 0; b < nblocks; ++b) t += partial[(int64_t)b * len + idx];
    if (!pp.on) {
        if (idx < len) pack[idx] = t;
        return;
    }
    if (idx < len)
        for (int r = 0; r < pp.nranks; ++r) pp.peer_pack[r][(int64_t)pp.rank * kMaxPack + idx] = t;
    __threadfence_system();
    __syncthreads();
    if (idx < pp.nranks) *((volatile int *)(pp.peer_flags[idx] + pp.rank)) = pp.epoch;
}

// Stand-alone form, used when an NCCL all-reduce sits between the local sums and the controller.
__global__ void __launch_bounds__(kMaxPack + 24)
dots_reduce_kernel(int m, int nblocks, const int *__restrict__ ic, const double *__restrict__ partial,
                   double *__restrict__ pack, P2P pp) {
    if (ic[I_GATE] & kGateDone) return;
    reduce_partials(m, nblocks, partial, pack, pp);
}

// Single-thread controller: line-search decision, history bookkeeping, m x m solves.
__device__ void control_body(int m, int mixed_enabled, int *__restrict__ ic, double *__restrict__ sc,
                             double *__restrict__ pack, double *__restrict__ RS, double *__restrict__ YY,
                             double *__restrict__ coef) {
    const int maxiter = ic[I_MAXITER];
    const double *last = pack + m * kDotVals;
    const double p_y = last[0], yPy = last[1], p_g = last[2], yPg = last[3], gg = last[4], gPg = last[5];
    ic[I_NEVAL] += 1;
    const double gnorm = sqrt(gg);
    sc[D_GNORM] = gnorm;
    if (!(gg == gg) || isinf(gg)) {            // non-finite gradient
        ic[I_FLAGS] |= UPOL_FLAG_NONFINITE;
        ic[I_GATE] |= kGateDone;
        return;
    }
    if (gnorm <= sc[D_TOL]) {
        if (mixed_enabled && (ic[I_GATE] & kGatePhaseMixed)) {
            // converged on the fp32 field: confirm (and continue) with the fp64 field from here
            ic[I_GATE] = (ic[I_GATE] & ~kGatePhaseMixed) | kGatePhase64;
            ic[I_ACCEPT] = 0;        // re-evaluate the same point: alpha unchanged
            ic[I_PUSHED] = 0;
            return;
        }
        ic[I_GATE] |= kGateDone;
        return;
    }
    bool accept = false, first = ic[I_FIRST] != 0;
    const double alpha = sc[D_ALPHA];
    if (first) {
        sc[D_GNORM0] = gnorm;
        // mixed field error ~1.5e-6 of the gross field (= ||g(d=0)||).  Hand-over swept in round 2
        // (profiles/r02_variants_switch.txt): 1e-4 saves 4 % at 100 k Drudes and 3 % at 20 k against 1e-3; 3e-5 and
        // below start to cost a 14th iteration at 100 k Drudes and leave less than 20x head-room over the noise.
        sc[D_SWITCH] = sc[D_SWITCH_REL] * gnorm;
        accept = true;
    } else {
        const double phi0 = sc[D_PHI0];
        const double dphi = p_g;
        accept = (dphi >= 0.9 * phi0) && (dphi <= -0.8 * phi0);
        if (!accept && ic[I_LS] >= kMaxLineSearch) {
            accept = true;
            ic[I_FLAGS] |= UPOL_FLAG_LINESEARCH;
        }
        if (!accept) {
            // safeguarded secant on phi'(a): bracket [alo (phi' < 0), ahi (phi' > 0)]
            double alo = sc[D_ALO], dlo = sc[D_DLO], ahi = sc[D_AHI], dhi = sc[D_DHI];
            if (dphi > 0.0) { ahi = alpha; dhi = dphi; } else { alo = alpha; dlo = dphi; }
            double an;
            if (ahi > 0.0) {                       // bracketed: secant inside, bisection if degenerate
                an = alo - dlo * (ahi - alo) / (dhi - dlo);
                const double w = ahi - alo;
                if (!(an > alo + 0.05 * w) || !(an < ahi - 0.05 * w)) an = 0.5 * (alo + ahi);
            } else {                               // still descending: extrapolate, at most 4x
                const double den = phi0 - dphi;    // < 0 when curvature is positive
                an = (den < 0.0) ? alpha * phi0 / den : 4.0 * alpha;
                if (!(an > 1.1 * alpha)) an = 2.0 * alpha;
                if (an > 4.0 * alpha) an = 4.0 * alpha;
            }
            sc[D_ALO] = alo; sc[D_DLO] = dlo; sc[D_AHI] = ahi; sc[D_DHI] = dhi;
            sc[D_ALPHA_PREV] = alpha;
            sc[D_ALPHA] = an;
            ic[I_LS] += 1;
            ic[I_ACCEPT] = 0;
            ic[I_PUSHED] = 0;
            return;
        }
    }

    // ---- accepted: close the iteration, update the history, build the next direction -------
    if (!first) ic[I_ITER] += 1;
    if (ic[I_ITER] >= maxiter) {
        ic[I_FLAGS] |= UPOL_FLAG_NOT_CONVERGED;
        ic[I_GATE] |= kGateDone;
        return;
    }
    if (mixed_enabled && (ic[I_GATE] & kGatePhaseMixed) && gnorm <= sc[D_SWITCH])
        ic[I_GATE] = (ic[I_GATE] & ~kGatePhaseMixed) | kGatePhase64;

    int count = ic[I_HCOUNT], head = ic[I_HHEAD];
    int pushed = 0, newslot = 0;
    const double sy = alpha * p_y;
    if (!first && m > 0 && sy > 1e-300 && yPy > 0.0) {
        // chronological matrices: drop the oldest row/column when full
        if (count == m) {
            for (int i = 1; i < m; ++i)
                for (int j = 1; j < m; ++j) { RS[(i - 1) * m + (j - 1)] = RS[i * m + j]; YY[(i - 1) * m + (j - 1)] = YY[i * m + j]; }
            newslot = head;
            head = (head + 1) % m;
            count = m - 1;
        } else {
            newslot = (head + count) % m;
        }
        // new column (logical index count): s_i.y_new, y_i.P.y_new for the kept pairs
        for (int j = 0; j < count; ++j) {
            const int phys = (head + j) % m;
            RS[j * m + count] = pack[phys * kDotVals + 0];
            YY[j * m + count] = pack[phys * kDotVals + 1];
            YY[count * m + j] = pack[phys * kDotVals + 1];
            RS[count * m + j] = 0.0;
        }
        RS[count * m + count] = sy;
        YY[count * m + count] = yPy;
        count += 1;
        pushed = 1;
        sc[D_GAMMA] = sy / yPy;
    }
    ic[I_HCOUNT] = count; ic[I_HHEAD] = head; ic[I_PUSHED] = pushed; ic[I_NEWSLOT] = newslot;

    // p = -H g,  H = gamma P + [S  gamma P Y] [[R^-T (D + gamma Y'PY) R^-1, -R^-T], [-R^-1, 0]] [S' ; gamma Y'P]
    const double gamma = sc[D_GAMMA];
    double v1[kMaxHist], v2[kMaxHist], u[kMaxHist], t[kMaxHist], top[kMaxHist];
    for (int j = 0; j < count; ++j) {
        const int phys = (head + j) % m;
        if (pushed && j == count - 1) { v1[j] = alpha * p_g; v2[j] = gamma * yPg; }
        else { v1[j] = pack[phys * kDotVals + 2]; v2[j] = gamma * pack[phys * kDotVals + 3]; }
    }
    for (int j = count - 1; j >= 0; --j) {            // u = R^-1 v1   (R upper triangular)
        double acc = v1[j];
        for (int l = j + 1; l < count; ++l) acc -= RS[j * m + l] * u[l];
        u[j] = acc / RS[j * m + j];
    }
    for (int j = 0; j < count; ++j) {                 // t = (D + gamma Y'PY) u - v2
        double acc = RS[j * m + j] * u[j] - v2[j];
        for (int l = 0; l < count; ++l) acc += gamma * YY[j * m + l] * u[l];
        t[j] = acc;
    }
    for (int j = 0; j < count; ++j) {                 // top = R^-T t   (R^T lower triangular)
        double acc = t[j];
        for (int l = 0; l < j; ++l) acc -= RS[l * m + j] * top[l];
        top[j] = acc / RS[j * m + j];
    }
    // phi'(0) of the new direction: g.p = -gamma gPg - top.v1 + u.v2
    double phi0n = -gamma * gPg;
    for (int j = 0; j < count; ++j) phi0n += -top[j] * v1[j] + u[j] * v2[j];
    if (!(phi0n < 0.0)) {                             // not a descent direction: restart from H0
        count = 0; head = 0;
        ic[I_HCOUNT] = 0; ic[I_HHEAD] = 0;
        if (pushed) { ic[I_PUSHED] = 0; }
        phi0n = -gamma * gPg;
    }
    for (int l = 0; l < 2 * kMaxHist; ++l) coef[l] = 0.0;
    for (int j = 0; j < count; ++j) {
        const int phys = (head + j) % m;
        coef[phys] = top[j];
        coef[kMaxHist + phys] = u[j];
    }
    sc[D_PHI0] = phi0n;
    sc[D_ALPHA_PREV] = alpha;
    sc[D_ALPHA] = 1.0;
    sc[D_ALO] = 0.0; sc[D_DLO] = phi0n; sc[D_AHI] = 0.0; sc[D_DHI] = 0.0;
    ic[I_LS] = 0;
    ic[I_ACCEPT] = 1;
    ic[I_FIRST] = 0;
}

__global__ void __launch_bounds__(kMaxPack + 24)
control_kernel(int m, int mixed_enabled, int *__restrict__ ic, double *__restrict__ sc,
               double *__restrict__ pack, double *__restrict__ RS, double *__restrict__ YY,
               double *__restrict__ coef, P2P pp, int nblocks_reduce, const double *__restrict__ partial) {
    if (ic[I_GATE] & kGateDone) return;
    if (nblocks_reduce > 0) {
        // single rank or peer-memory exchange: the fold of the block partials happens here (one launch less)
        reduce_partials(m, nblocks_reduce, partial, pack, pp);
        __threadfence();
        __syncthreads();
    }
    if (pp.on) {
        // wait for every rank's partial sums, then add them in rank order: every rank gets the same bits
        __shared__ int bad;
        if (threadIdx.x == 0) bad = 0;
        __syncthreads();
        if ((int)threadIdx.x < pp.nranks && !wait_epoch(pp.my_flags, threadIdx.x, pp.epoch)) bad = 1;
        __syncthreads();
        if (bad) {
            if (threadIdx.x == 0) { ic[I_FLAGS] |= UPOL_FLAG_COMM; ic[I_GATE] |= kGateDone; }
            return;
        }
        const int len = (m + 1) * kDotVals;
        if ((int)threadIdx.x < len) {
            double t = 0.0;
            for (int r = 0; r < pp.nranks; ++r) t += ((volatile double *)pp.my_pack)[(int64_t)r * kMaxPack + threadIdx.x];
            pack[threadIdx.x] = t;
        }
        __threadfence();
        __syncthreads();
    }
    if (threadIdx.x) return;
    control_body(m, mixed_enabled, ic, sc, pack, RS, YY, coef);
}

// Single-rank slot, O(N) part in two launches instead of five.
// (1) dots_fused_kernel: the y = m blocks apply the block preconditioner (Pg = P g, Py = P (g - g0)) to the elements
//     they sum over, every block writes its partial inner products, and the block that finishes LAST folds the
//     partials in fixed order and runs the controller (control_body) - no separate precondition / reduce / control
//     launches.
// (2) update_rows_kernel: one thread per Drude row moves its three components (new pair, new direction, new x) and
//     writes the shell position / shell column / fp32 displacement record the next field pass reads - the job of
//     prepare_shells_kernel, so that launch disappears from the loop too.
__global__ void __launch_bounds__(kDotBlock)
dots_fused_kernel(int m, int64_t n, int64_t e0, int64_t e1, int mixed_enabled, int *__restrict__ ic, double *__restrict__ sc,
                  const double *__restrict__ S, const double *__restrict__ Y, const double *__restrict__ g,
                  const double *__restrict__ g0, const double *__restrict__ p, double *__restrict__ Pg, double *__restrict__ Py,
                  const int *__restrict__ row_mol, const int *__restrict__ mol_row0, const int *__restrict__ mol_nrow,
                  const int64_t *__restrict__ blk_off, const double *__restrict__ blk,
                  double *partial, double *__restrict__ pack, double *__restrict__ RS, double *__restrict__ YY,
                  double *__restrict__ coef) {
    if (ic[I_GATE] & kGateDone) return;
    __shared__ double sm[kDotVals][kDotBlock / 32];
    __shared__ int is_last;
    const int l = blockIdx.y;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t e = e0 + (int64_t)blockIdx.x * kDotBlock + threadIdx.x; e < e1; e += (int64_t)gridDim.x * kDotBlock) {
        const double ge = g[e], ye = ge - g0[e];
        if (l < m) {
            const double se = S[(int64_t)l * n + e], yl = Y[(int64_t)l * n + e];          // Y holds P y_l (P symmetric)
            acc[0] += se * ye; acc[1] += yl * ye; acc[2] += se * ge; acc[3] += yl * ge;
        } else {
            const int mol = row_mol[e / 3];
            const int64_t b0 = 3 * (int64_t)mol_row0[mol];
            const int nb = 3 * mol_nrow[mol];
            const double *row = blk + blk_off[mol] + (e - b0) * nb;
            double pge = 0.0, pye = 0.0;
            for (int j = 0; j < nb; ++j) {
                const double gj = g[b0 + j];
                pge += row[j] * gj;
                pye += row[j] * (gj - g0[b0 + j]);
            }
            Pg[e] = pge; Py[e] = pye;
            const double ae = p[e];
            acc[0] += ae * ye; acc[1] += ye * pye; acc[2] += ae * ge; acc[3] += pye * ge;
            acc[4] += ge * ge; acc[5] += ge * pge;
        }
    }
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double v = acc[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (ln == 0) sm[c][w] = v;
    }
    __syncthreads();
    const int len = (m + 1) * kDotVals;
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int k = 0; k < kDotBlock / 32; ++k) t += sm[threadIdx.x][k];
        partial[((int64_t)blockIdx.x * (m + 1) + l) * kDotVals + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ic + I_TICKET, 1) == (int)(gridDim.x * gridDim.y) - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if ((int)threadIdx.x < len) {
        double t = 0.0;
        if ((threadIdx.x % kDotVals) < 6)
            for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(partial + (int64_t)b * len + threadIdx.x);
        pack[threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x) return;
    ic[I_TICKET] = 0;
    control_body(m, mixed_enabled, ic, sc, pack, RS, YY, coef);
}

__global__ void __launch_bounds__(128)
update_rows_kernel(int m, int64_t n, int row_lo, int row_hi, const int *__restrict__ ic, const double *__restrict__ sc,
                   const double *__restrict__ coef, double *__restrict__ S, double *__restrict__ Y,
                   double *__restrict__ x, double *__restrict__ x0, const double *__restrict__ g,
                   double *__restrict__ g0, double *__restrict__ p, const double *__restrict__ Pg,
                   const double *__restrict__ Py, const int *__restrict__ row_site, const double *__restrict__ r,
                   const double *__restrict__ row_qs, double inv_grid, double *__restrict__ rowx, double *__restrict__ rowy,
                   double *__restrict__ rowz, double4 *__restrict__ colsB, float4 *__restrict__ cols_d, float *__restrict__ cols_d2) {
    if (ic[I_GATE] & kGateDone) return;
    const int i = row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_hi) return;
    const bool accept = ic[I_ACCEPT] != 0, pushed = ic[I_PUSHED] != 0;
    const int ns = ic[I_NEWSLOT];
    const double alpha = sc[D_ALPHA], alpha_prev = sc[D_ALPHA_PREV], gamma = sc[D_GAMMA];
    double xn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int64_t e = 3 * (int64_t)i + c;
        if (accept) {
            const double ge = g[e];
            if (pushed) {
                S[(int64_t)ns * n + e] = alpha_prev * p[e];
                Y[(int64_t)ns * n + e] = Py[e];                 // P y_new
            }
            double pn = -gamma * Pg[e];
            for (int l = 0; l < m; ++l) {
                const double a = coef[l], b = coef[kMaxHist + l];
                if (a != 0.0 || b != 0.0) pn += -a * S[(int64_t)l * n + e] + gamma * b * Y[(int64_t)l * n + e];
            }
            const double xe = x[e];
            x0[e] = xe; g0[e] = ge; p[e] = pn;
            xn[c] = xe + alpha * pn;
        } else {
            xn[c] = x0[e] + alpha * p[e];
        }
        x[e] = xn[c];
    }
    // what prepare_shells_kernel would write for this row (system.cu)
    const int st = row_site[i];
    const double px = r[3 * st] - xn[0], py = r[3 * st + 1] - xn[1], pz = r[3 * st + 2] - xn[2];
    rowx[i] = px; rowy[i] = py; rowz[i] = pz;
    colsB[i] = make_double4(px, py, pz, row_qs[i]);
    const float dx = (float)(xn[0] * inv_grid), dy = (float)(xn[1] * inv_grid), dz = (float)(xn[2] * inv_grid);
    cols_d[i] = make_float4(-2.0f * dx, -2.0f * dy, -2.0f * dz, (float)row_qs[i]);
    cols_d2[i] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// Element-wise update over the local slice: store the new pair, form the new direction, move x.
__global__ void __launch_bounds__(256)
update_kernel(int m, int64_t n, int64_t e0, int64_t e1, const int *__restrict__ ic, const double *__restrict__ sc,
              const double *__restrict__ coef, double *__restrict__ S, double *__restrict__ Y,
              double *__restrict__ x, double *__restrict__ x0, const double *__restrict__ g,
              double *__restrict__ g0, double *__restrict__ p, const double *__restrict__ Pg,
              const double *__restrict__ Py, P2P pp) {
    if (ic[I_GATE] & kGateDone) return;
    const int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e < e1;
    double xnew = 0.0;
    if (live && ic[I_ACCEPT]) {
        const double ge = g[e];
        if (ic[I_PUSHED]) {
            const int ns = ic[I_NEWSLOT];
            S[(int64_t)ns * n + e] = sc[D_ALPHA_PREV] * p[e];
            Y[(int64_t)ns * n + e] = Py[e];                 // P y_new
        }
        const double gamma = sc[D_GAMMA];
        double pn = -gamma * Pg[e];
        for (int l = 0; l < m; ++l) {
            const double a = coef[l], b = coef[kMaxHist + l];
            if (a != 0.0 || b != 0.0) pn += -a * S[(int64_t)l * n + e] + gamma * b * Y[(int64_t)l * n + e];
        }
        const double xe = x[e];
        x0[e] = xe; g0[e] = ge; p[e] = pn;
        xnew = xe + sc[D_ALPHA] * pn;
        x[e] = xnew;
    } else if (live) {
        xnew = x0[e] + sc[D_ALPHA] * p[e];
        x[e] = xnew;
    }
    if (!pp.on) return;
    // fused all-gather: this rank's slice goes straight into every peer's vector (NVLink stores) ...
    if (live)
        for (int r = 0; r < pp.nranks; ++r)
            if (r != pp.rank) pp.peer_x[r][e] = xnew;
    // ... and the block that finishes last raises this rank's arrival flag on every peer
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = (atomicAdd(pp.my_flags + 2 * pp.nranks, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) pp.my_flags[2 * pp.nranks] = 0;
        __threadfence_system();
        if ((int)threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank)
            *((volatile int *)(pp.peer_flags[threadIdx.x] + pp.nranks + pp.rank)) = pp.epoch;
    }
}

}  // namespace

void minimize_free(upol_system *s) {
    if (s->min_ws) { ws_free(static_cast<MinWs *>(s->min_ws)); s->min_ws = nullptr; }
}

static int ws_ensure(upol_system *s, int m, MinWs **out) {
    MinWs *w = static_cast<MinWs *>(s->min_ws);
    const int64_t n = 3 * s->nd;
    if (w && w->m == m && w->n == n) { *out = w; return UPOL_OK; }
    if (w) { ws_free(w); s->min_ws = nullptr; }
    w = new MinWs();
    w->m = m; w->n = n;
    const int64_t nn = std::max<int64_t>(n, 1);
    w->nblocks = (int)std::min<int64_t>(148 * 2, (nn + kDotBlock - 1) / kDotBlock);
    auto A = [&](double **p, size_t cnt) { return upol::dmalloc((void **)p, sizeof(double) * std::max<size_t>(cnt, 1)) == cudaSuccess; };
    bool ok = A(&w->x0, nn) && A(&w->g0, nn) && A(&w->p, nn) && A(&w->g, nn) && A(&w->Pg, nn) && A(&w->Py, nn) &&
              A(&w->S, (size_t)std::max(m, 1) * nn) && A(&w->Y, (size_t)std::max(m, 1) * nn) &&
              A(&w->partial, (size_t)w->nblocks * (m + 1) * kDotVals) && A(&w->pack, (size_t)(m + 1) * kDotVals) &&
              A(&w->sc, D_SIZE) && A(&w->coef, 2 * kMaxHist) && A(&w->RS, (size_t)std::max(m * m, 1)) &&
              A(&w->YY, (size_t)std::max(m * m, 1)) && A(&w->e_final, UPOL_E_SIZE);
    ok = ok && upol::dmalloc((void **)&w->ic, sizeof(int) * I_SIZE) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&w->h_ic, sizeof(int) * I_SIZE * 2) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&w->h_sc, sizeof(double) * D_SIZE * 2) == cudaSuccess;
    for (auto &e : w->ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    // molecule-block tables of the preconditioner
    {
        w->nmol = s->nmol;
        std::vector<int64_t> off(std::max<int64_t>(s->nmol, 1), 0);
        std::vector<int> rmol(std::max<int64_t>(s->nd, 1), 0);
        int64_t tot = 0;
        int max_rows = 0;
        for (int64_t mm = 0; mm < s->nmol; ++mm) {
            off[mm] = tot;
            const int nr = s->h_mol_nrow[mm];
            max_rows = std::max(max_rows, nr);
            tot += (int64_t)9 * nr * nr;
            for (int a = 0; a < nr; ++a) rmol[s->h_mol_row0[mm] + a] = (int)mm;
        }
        w->blk_elems = std::max<int64_t>(tot, 1);
        ok = ok && A(&w->blk, (size_t)w->blk_elems);
        ok = ok && upol::dmalloc((void **)&w->blk_off, sizeof(int64_t) * off.size()) == cudaSuccess;
        ok = ok && upol::dmalloc((void **)&w->mol_row0, sizeof(int) * off.size()) == cudaSuccess;
        ok = ok && upol::dmalloc((void **)&w->mol_nrow, sizeof(int) * off.size()) == cudaSuccess;
        ok = ok && upol::dmalloc((void **)&w->row_mol, sizeof(int) * rmol.size()) == cudaSuccess;
        if (ok && s->nmol > 0) {
            ok = ok && cudaMemcpy(w->blk_off, off.data(), sizeof(int64_t) * s->nmol, cudaMemcpyHostToDevice) == cudaSuccess;
            ok = ok && cudaMemcpy(w->mol_row0, s->h_mol_row0.data(), sizeof(int) * s->nmol, cudaMemcpyHostToDevice) == cudaSuccess;
            ok = ok && cudaMemcpy(w->mol_nrow, s->h_mol_nrow.data(), sizeof(int) * s->nmol, cudaMemcpyHostToDevice) == cudaSuccess;
            ok = ok && cudaMemcpy(w->row_mol, rmol.data(), sizeof(int) * rmol.size(), cudaMemcpyHostToDevice) == cudaSuccess;
            // pageable copies on the legacy stream: make sure they have landed before kernels on the caller's
            // (possibly non-blocking) stream read the tables
            ok = ok && cudaDeviceSynchronize() == cudaSuccess;
        }
        w->max_rows = max_rows;
    }
    if (!ok) { set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__); ws_free(w); return UPOL_ERR_ALLOC; }
    s->min_ws = w;
    *out = w;
    return UPOL_OK;
}

int minimize_compact(upol_system *s, double *x, const upol_min_options *o, upol_min_result *res, cudaStream_t st_caller) {
    memset(res, 0, sizeof(*res));
    int m = o->method == UPOL_CG ? 1 : (o->history > 0 ? o->history : 8);
    m = std::min(m, kMaxHist);
    const int check_every = std::max(1, o->check_every);
    const bool mixed = o->precision == UPOL_MIXED;
    MinWs *w = nullptr;
    int rc = ws_ensure(s, m, &w);
    if (rc) return rc;
    const int64_t n = w->n;
    const int64_t e0 = 3 * s->row_lo, e1 = 3 * s->row_hi;
    const int64_t nloc = std::max<int64_t>(e1 - e0, 0);
    const unsigned ublocks = (unsigned)std::max<int64_t>(1, (nloc + 255) / 256);

    // Single rank, open boundaries: the O(N) part of a slot is two fused launches (dots_fused_kernel,
    // update_rows_kernel), the shells are kept current by the update, and chunks of slots are replayed from a CUDA
    // graph (the whole solve under one jit in the reference, polarization.py:154,178-179).  Capture needs a
    // non-legacy stream, so this path runs on the workspace's own stream, ordered after / before the caller's by events.
    const bool fused = s->nranks == 1 && !s->ewald && s->nd > 0 && s->row_hi > s->row_lo;
    static const bool no_graph = getenv("UPOL_NO_GRAPH") != nullptr;
    cudaStream_t st = st_caller;
    if (fused) {
        if (!w->gstream) {
            UPOL_CUDA(cudaStreamCreateWithFlags(&w->gstream, cudaStreamNonBlocking));
            for (auto &e : w->gev) UPOL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        UPOL_CUDA(cudaEventRecord(w->gev[0], st_caller));
        UPOL_CUDA(cudaStreamWaitEvent(w->gstream, w->gev[0], 0));
        st = w->gstream;
    }

    if (s->mol_hi > s->mol_lo)      // preconditioner blocks of this rank's molecules, current geometry
    {
        if (w->max_rows <= 10)        // 3 nrow <= 32: one warp per molecule, block in shared memory
            build_blocks_warp_kernel<<<(unsigned)((s->mol_hi - s->mol_lo + kBlkWarps - 1) / kBlkWarps), 32 * kBlkWarps, 0, st>>>(
                (int)s->mol_lo, (int)s->mol_hi, w->mol_row0, w->mol_nrow, w->blk_off, s->row_site, s->r, s->row_qs,
                s->row_k, s->th_ptr, s->th_row, s->th_u, w->blk);
        else
            build_blocks_kernel<<<(unsigned)((s->mol_hi - s->mol_lo + 63) / 64), 64, 0, st>>>(
                (int)s->mol_lo, (int)s->mol_hi, w->mol_row0, w->mol_nrow, w->blk_off, s->row_site, s->r, s->row_qs,
                s->row_k, s->th_ptr, s->th_row, s->th_u, w->blk);
    }
    static const double switch_rel = getenv("UPOL_MIXED_SWITCH") ? atof(getenv("UPOL_MIXED_SWITCH")) : 1e-4;
    init_state_kernel<<<1, 1, 0, st>>>(w->ic, w->sc, o->tol, mixed ? kGatePhaseMixed : kGatePhase64, o->maxiter, switch_rel);
    UPOL_CUDA(cudaMemsetAsync(w->S, 0, sizeof(double) * (size_t)std::max(m, 1) * std::max<int64_t>(n, 1), st));
    UPOL_CUDA(cudaMemsetAsync(w->Y, 0, sizeof(double) * (size_t)std::max(m, 1) * std::max<int64_t>(n, 1), st));
    UPOL_CUDA(cudaMemsetAsync(w->p, 0, sizeof(double) * std::max<int64_t>(n, 1), st));
    UPOL_CUDA(cudaMemsetAsync(w->g0, 0, sizeof(double) * std::max<int64_t>(n, 1), st));
    if (n > 0) UPOL_CUDA(cudaMemcpyAsync(w->x0, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    UPOL_CUDA(cudaMemsetAsync(w->partial, 0, sizeof(double) * (size_t)w->nblocks * (m + 1) * kDotVals, st));

    EvalOpts eo;
    eo.precision = mixed ? UPOL_MIXED : UPOL_FP64;
    eo.want_energy = false;
    eo.want_grad = true;
    eo.reduce_ranks = false;
    eo.gate = w->ic + I_GATE;
    eo.phase_switch = mixed;
    eo.shells_ready = fused;          // update_rows_kernel keeps the shell rows / columns current

    P2P pp;
    memset(&pp, 0, sizeof(pp));
    pp.on = (s->p2p && s->nranks > 1) ? 1 : 0;
    pp.rank = s->rank; pp.nranks = s->nranks;
    pp.peer_x = s->d_peer_x; pp.peer_pack = s->d_peer_pack; pp.peer_flags = s->d_peer_flags;
    pp.my_pack = s->xchg_pack; pp.my_flags = s->xchg_flags;
    int slots_in_call = 0;
    if (pp.on) {
        // Nothing else aligns the ranks when they enter this call, and the peer waits below give up after ~2 s:
        // a host-side skew (data loading, a first-call set-up on one rank) must not be mistaken for a dead peer.
        // One tiny NCCL all-reduce on the stream is a device-side barrier without a time-out.
        rc = comm_allreduce_sum(s, w->pack, 1, st);
        if (rc) return rc;
    }
    if (fused) {                       // shells of the starting point; from here on the update kernel writes them
        rc = sys_prepare_shells(s, x, st);
        if (rc) return rc;
    }

    auto enqueue_slot = [&]() -> int {
        if (pp.on) {
            if (slots_in_call > 0) wait_x_kernel<<<1, 32, 0, st>>>(w->ic, pp, s->p2p_epoch);
            pp.epoch = ++s->p2p_epoch;
        }
        ++slots_in_call;
        int r = sys_eval_compact(s, x, eo, s->energy, w->g, st);
        if (r) return r;
        dim3 dg((unsigned)w->nblocks, (unsigned)(m + 1));
        if (fused) {
            dots_fused_kernel<<<dg, kDotBlock, 0, st>>>(m, n, e0, e1, mixed ? 1 : 0, w->ic, w->sc, w->S, w->Y, w->g, w->g0, w->p,
                                                        w->Pg, w->Py, w->row_mol, w->mol_row0, w->mol_nrow, w->blk_off, w->blk,
                                                        w->partial, w->pack, w->RS, w->YY, w->coef);
            update_rows_kernel<<<(unsigned)((s->row_hi - s->row_lo + 127) / 128), 128, 0, st>>>(
                m, n, (int)s->row_lo, (int)s->row_hi, w->ic, w->sc, w->coef, w->S, w->Y, x, w->x0, w->g, w->g0, w->p, w->Pg, w->Py,
                s->row_site, s->r, s->row_qs, s->inv_grid, s->rowx, s->rowy, s->rowz, s->cols + s->na_pad, s->cols_d, s->cols_d2);
            s->launches += 2;
            UPOL_CUDA(cudaGetLastError());
            return UPOL_OK;
        }
        precond_apply_kernel<<<ublocks, 256, 0, st>>>(e0, e1, w->ic, w->row_mol, w->mol_row0, w->mol_nrow, w->blk_off,
                                                      w->blk, w->g, w->g0, w->Pg, w->Py);
        dots_kernel<<<dg, kDotBlock, 0, st>>>(m, n, e0, e1, w->ic, w->S, w->Y, w->g, w->g0, w->p, w->Pg, w->Py, w->partial);
        const bool nccl_between = s->nranks > 1 && !pp.on;
        if (nccl_between) {
            dots_reduce_kernel<<<1, kMaxPack + 24, 0, st>>>(m, w->nblocks, w->ic, w->partial, w->pack, pp);
            r = comm_allreduce_sum(s, w->pack, (int64_t)(m + 1) * kDotVals, st);
            if (r) return r;
            ++s->launches;
        }
        control_kernel<<<1, kMaxPack + 24, 0, st>>>(m, mixed ? 1 : 0, w->ic, w->sc, w->pack, w->RS, w->YY, w->coef, pp,
                                                    nccl_between ? 0 : w->nblocks, w->partial);
        update_kernel<<<ublocks, 256, 0, st>>>(m, n, e0, e1, w->ic, w->sc, w->coef, w->S, w->Y, x, w->x0, w->g, w->g0, w->p, w->Pg, w->Py, pp);
        s->launches += 4;
        if (s->nranks > 1 && !pp.on) { r = comm_allgather_rows(s, x, st); if (r) return r; }
        UPOL_CUDA(cudaGetLastError());
        return UPOL_OK;
    };

    // The graph of one chunk depends only on buffer addresses and launch shapes, never on values (tolerance, iteration
    // limit and all line-search state live in device memory), so it is captured once per (history, precision, chunk
    // length, row range, displacement vector) and replayed by every later minimisation of this system.
    bool use_graph = fused && !no_graph && !w->graph_failed;
    int64_t launches_per_chunk = 0;
    if (use_graph) {
        const long long key[6] = {m * 2 + (mixed ? 1 : 0), check_every, (long long)s->row_lo, (long long)s->row_hi, (long long)(intptr_t)x,
                                  (long long)s->geom_gen};   // box centre and fixed-point grid are kernel arguments
        if (!w->gexec || memcmp(key, w->gkey, sizeof(key)) != 0) {
            if (w->gexec) { cudaGraphExecDestroy(w->gexec); w->gexec = nullptr; }
            cudaGraph_t graph = nullptr;
            const int64_t l0 = s->launches;
            bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                int r = UPOL_OK;
                for (int c = 0; c < check_every && r == UPOL_OK; ++c) r = enqueue_slot();
                ok = (cudaStreamEndCapture(st, &graph) == cudaSuccess) && r == UPOL_OK && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&w->gexec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            w->glaunches = s->launches - l0;
            s->launches = l0;
            slots_in_call = 0;
            if (!ok) {
                cudaGetLastError();
                w->graph_failed = true; use_graph = false;
                if (w->gexec) { cudaGraphExecDestroy(w->gexec); w->gexec = nullptr; }
            } else {
                memcpy(w->gkey, key, sizeof(key));
            }
        }
        launches_per_chunk = w->glaunches;
    }

    static const bool trace = getenv("UPOL_MIN_TRACE") != nullptr;
    // Enqueue chunks of slots, polling the done flag one chunk behind.
    const int max_slots = o->maxiter * 3 + kMaxLineSearch + 8;
    int enq = 0, buf = 0;
    bool done = false;
    bool pending[2] = {false, false};
    while (!done) {
        if (use_graph) {
            UPOL_CUDA(cudaGraphLaunch(w->gexec, st));
            enq += check_every;
            s->launches += launches_per_chunk;
        } else {
            for (int c = 0; c < check_every && enq < max_slots; ++c, ++enq) { rc = enqueue_slot(); if (rc) return rc; }
        }
        UPOL_CUDA(cudaMemcpyAsync(w->h_ic + buf * I_SIZE, w->ic, sizeof(int) * I_SIZE, cudaMemcpyDeviceToHost, st));
        if (trace) {     // debug aid (UPOL_MIN_TRACE=1, use check_every = 1): the controller's state after every chunk
            UPOL_CUDA(cudaMemcpyAsync(w->h_sc, w->sc, sizeof(double) * D_SIZE, cudaMemcpyDeviceToHost, st));
            UPOL_CUDA(cudaStreamSynchronize(st));
            const int *ic = w->h_ic + buf * I_SIZE;
            fprintf(stderr, "[upol minimise] slots %d iter %d evals %d ls %d accept %d gate %d alpha %.4g gnorm %.6e gamma %.4g\n", enq,
                    ic[I_ITER], ic[I_NEVAL], ic[I_LS], ic[I_ACCEPT], ic[I_GATE], w->h_sc[D_ALPHA_PREV], w->h_sc[D_GNORM], w->h_sc[D_GAMMA]);
        }
        UPOL_CUDA(cudaEventRecord(w->ev[buf], st));
        pending[buf] = true;
        const int prev = buf ^ 1;
        if (pending[prev]) {
            UPOL_CUDA(cudaEventSynchronize(w->ev[prev]));
            pending[prev] = false;
            if (w->h_ic[prev * I_SIZE + I_GATE] & kGateDone) done = true;
        }
        if (!done && enq >= max_slots) {
            UPOL_CUDA(cudaEventSynchronize(w->ev[buf]));
            pending[buf] = false;
            done = true;
        }
        buf ^= 1;
    }
    if (pp.on) wait_x_kernel<<<1, 32, 0, st>>>(w->ic, pp, s->p2p_epoch);   // only matters when not converged
    // fp64 energy at the minimiser (x is complete on every rank).  The gradient there is already
    // known from the last slot (its norm is in the control block), so this is a potential-only pass.
    EvalOpts ef;
    ef.precision = UPOL_FP64;
    ef.want_energy = true; ef.want_grad = false; ef.reduce_ranks = true;
    ef.shells_ready = fused;
    if (pp.on) ef.gx = make_eval_xchg(s, false);      // the energy partials meet in the peers' mailboxes, no NCCL
    rc = sys_eval_compact(s, x, ef, w->e_final, w->g, st);
    if (rc) return rc;
    UPOL_CUDA(cudaMemcpyAsync(w->h_ic, w->ic, sizeof(int) * I_SIZE, cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaMemcpyAsync(w->h_sc, w->e_final, sizeof(double) * UPOL_E_SIZE, cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaMemcpyAsync(w->h_sc + D_SIZE, w->sc, sizeof(double) * D_SIZE, cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaStreamSynchronize(st));
    if (fused) {                       // later work on the caller's stream comes after everything enqueued here
        UPOL_CUDA(cudaEventRecord(w->gev[1], st));
        UPOL_CUDA(cudaStreamWaitEvent(st_caller, w->gev[1], 0));
    }
    for (int i = 0; i < UPOL_E_SIZE; ++i) res->energy[i] = w->h_sc[i];
    res->gnorm = w->h_sc[D_SIZE + D_GNORM];
    res->energy[UPOL_E_GNORM2] = res->gnorm * res->gnorm;
    res->iterations = w->h_ic[I_ITER];
    res->evaluations = w->h_ic[I_NEVAL] + 1;
    res->flags = w->h_ic[I_FLAGS];
    if (!(w->h_ic[I_GATE] & kGateDone)) res->flags |= UPOL_FLAG_NOT_CONVERGED;
    if (!(res->gnorm <= o->tol) && !(res->flags & UPOL_FLAG_NONFINITE)) res->flags |= UPOL_FLAG_NOT_CONVERGED;
    else if (res->gnorm <= o->tol) res->flags &= ~UPOL_FLAG_NOT_CONVERGED;
    return res->flags;
}

}  // namespace upol
