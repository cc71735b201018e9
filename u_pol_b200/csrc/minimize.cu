// K5: device-resident minimiser over the Drude displacements.
// Replaces drudeOpt = jaxopt.BFGS(fun=Uind, tol=1e-4).run(Dij0) (U_pol/polarization.py:154-188).
//
// jaxopt's dense BFGS keeps a (3N)^2 inverse Hessian and an energy-based zoom line search; both
// stop working at scale (H is 2.3 TB at 100 k Drudes, and near the
// minimum dE ~ g^2/2k is below the round-off of an N^2-term fp64 sum).  This minimiser is
//   * limited-memory BFGS in the compact (Byrd-Nocedal-Schnabel) form with H0 = gamma*diag(1/k)
//     - U_ind is near-quadratic with a spring-dominated Hessian, so diag(1/k) is a good
//     preconditioner and the unit step is almost always accepted (one gradient per iteration);
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
#include <cstring>

#include "common.cuh"

namespace upol {

namespace {

constexpr int kMaxHist = 16;
constexpr int kDotBlock = 256;
constexpr int kDotVals = 8;          // per history slot: 4 used; last row: 6 used
constexpr int kMaxLineSearch = 12;

// integer control words
enum { I_GATE = 0, I_ITER, I_NEVAL, I_HCOUNT, I_HHEAD, I_LS, I_FLAGS, I_ACCEPT, I_FIRST, I_NEWSLOT, I_PUSHED, I_SIZE = 16 };
// double control words
enum { D_ALPHA = 0, D_ALPHA_PREV, D_PHI0, D_ALO, D_DLO, D_AHI, D_DHI, D_GAMMA, D_GNORM, D_GNORM0, D_TOL, D_SWITCH, D_SIZE = 16 };

struct MinWs {
    int m = 0;
    int64_t n = 0;             // 3 * nd
    double *x0 = nullptr, *g0 = nullptr, *p = nullptr, *g = nullptr, *pinv = nullptr;
    double *S = nullptr, *Y = nullptr;       // [m][n]
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
};

void ws_free(MinWs *w) {
    if (!w) return;
    void *ptrs[] = {w->x0, w->g0, w->p, w->g, w->pinv, w->S, w->Y, w->partial, w->pack, w->sc, w->ic,
                    w->coef, w->RS, w->YY, w->e_final};
    for (void *q : ptrs) if (q) cudaFree(q);
    if (w->h_ic) cudaFreeHost(w->h_ic);
    if (w->h_sc) cudaFreeHost(w->h_sc);
    for (auto &e : w->ev) if (e) cudaEventDestroy(e);
    delete w;
}

__global__ void init_pinv_kernel(int64_t n, const double *__restrict__ row_k, double *__restrict__ pinv) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double k = row_k[e / 3];
    pinv[e] = k > 0.0 ? 1.0 / k : 1.0;
}

__global__ void init_state_kernel(int *ic, double *sc, double tol, int phase_bits) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < I_SIZE; ++i) ic[i] = 0;
    for (int i = 0; i < D_SIZE; ++i) sc[i] = 0.0;
    ic[I_GATE] = phase_bits;
    ic[I_FIRST] = 1;
    sc[D_ALPHA] = 0.0;      // the first slot evaluates x itself
    sc[D_TOL] = tol;
    sc[D_GAMMA] = 1.0;
}

// Dot products of one slot over the local row slice [e0, e1).
// blockIdx.y = l < m : history slot l (physical): s_l.y, y_l.P.y, s_l.g, y_l.P.g      (y = g - g0)
// blockIdx.y = m     : p.y, y.P.y, p.g, y.P.g, g.g, g.P.g
__global__ void __launch_bounds__(kDotBlock)
dots_kernel(int m, int64_t n, int64_t e0, int64_t e1, const int *__restrict__ ic,
            const double *__restrict__ S, const double *__restrict__ Y, const double *__restrict__ g,
            const double *__restrict__ g0, const double *__restrict__ p, const double *__restrict__ pinv,
            double *__restrict__ partial) {
    if (ic[I_GATE] & kGateDone) return;
    __shared__ double sm[kDotVals][kDotBlock / 32];
    const int l = blockIdx.y;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    {   // history slots not yet filled hold zeros, so every slot can be summed unconditionally
        const double *a = l < m ? S + (int64_t)l * n : p;
        const double *b = l < m ? Y + (int64_t)l * n : nullptr;
        for (int64_t e = e0 + (int64_t)blockIdx.x * kDotBlock + threadIdx.x; e < e1; e += (int64_t)gridDim.x * kDotBlock) {
            const double ge = g[e], ye = ge - g0[e], pe = pinv[e];
            if (l < m) {
                const double se = a[e], yl = b[e] * pe;
                acc[0] += se * ye; acc[1] += yl * ye; acc[2] += se * ge; acc[3] += yl * ge;
            } else {
                const double ae = a[e];
                acc[0] += ae * ye; acc[1] += ye * pe * ye; acc[2] += ae * ge; acc[3] += ye * pe * ge;
                acc[4] += ge * ge; acc[5] += ge * pe * ge;
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

__global__ void dots_reduce_kernel(int m, int nblocks, const int *__restrict__ ic,
                                   const double *__restrict__ partial, double *__restrict__ pack) {
    if (ic[I_GATE] & kGateDone) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (m + 1) * kDotVals) return;
    double t = 0.0;
    if ((idx % kDotVals) < 6)
        for (int b = 0; b < nblocks; ++b) t += partial[(int64_t)b * (m + 1) * kDotVals + idx];
    pack[idx] = t;
}

// Single-thread controller: line-search decision, history bookkeeping, m x m solves.
__global__ void control_kernel(int m, int maxiter, int mixed_enabled, int *__restrict__ ic, double *__restrict__ sc,
                               const double *__restrict__ pack, double *__restrict__ RS, double *__restrict__ YY,
                               double *__restrict__ coef) {
    if (threadIdx.x || blockIdx.x) return;
    if (ic[I_GATE] & kGateDone) return;
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
        sc[D_SWITCH] = 1e-3 * gnorm;
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

// Element-wise update over the local slice: store the new pair, form the new direction, move x.
__global__ void __launch_bounds__(256)
update_kernel(int m, int64_t n, int64_t e0, int64_t e1, const int *__restrict__ ic, const double *__restrict__ sc,
              const double *__restrict__ coef, double *__restrict__ S, double *__restrict__ Y,
              double *__restrict__ x, double *__restrict__ x0, const double *__restrict__ g,
              double *__restrict__ g0, double *__restrict__ p, const double *__restrict__ pinv) {
    if (ic[I_GATE] & kGateDone) return;
    const int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e1) return;
    if (ic[I_ACCEPT]) {
        const double ge = g[e], pe = pinv[e];
        if (ic[I_PUSHED]) {
            const int ns = ic[I_NEWSLOT];
            S[(int64_t)ns * n + e] = sc[D_ALPHA_PREV] * p[e];
            Y[(int64_t)ns * n + e] = ge - g0[e];
        }
        const double gamma = sc[D_GAMMA];
        double pn = -gamma * pe * ge;
        for (int l = 0; l < m; ++l) {
            const double a = coef[l], b = coef[kMaxHist + l];
            if (a != 0.0 || b != 0.0) pn += -a * S[(int64_t)l * n + e] + gamma * pe * b * Y[(int64_t)l * n + e];
        }
        const double xe = x[e];
        x0[e] = xe; g0[e] = ge; p[e] = pn;
        x[e] = xe + sc[D_ALPHA] * pn;
    } else {
        x[e] = x0[e] + sc[D_ALPHA] * p[e];
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
    auto A = [&](double **p, size_t cnt) { return cudaMalloc((void **)p, sizeof(double) * std::max<size_t>(cnt, 1)) == cudaSuccess; };
    bool ok = A(&w->x0, nn) && A(&w->g0, nn) && A(&w->p, nn) && A(&w->g, nn) && A(&w->pinv, nn) &&
              A(&w->S, (size_t)std::max(m, 1) * nn) && A(&w->Y, (size_t)std::max(m, 1) * nn) &&
              A(&w->partial, (size_t)w->nblocks * (m + 1) * kDotVals) && A(&w->pack, (size_t)(m + 1) * kDotVals) &&
              A(&w->sc, D_SIZE) && A(&w->coef, 2 * kMaxHist) && A(&w->RS, (size_t)std::max(m * m, 1)) &&
              A(&w->YY, (size_t)std::max(m * m, 1)) && A(&w->e_final, UPOL_E_SIZE);
    ok = ok && cudaMalloc((void **)&w->ic, sizeof(int) * I_SIZE) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&w->h_ic, sizeof(int) * I_SIZE * 2) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&w->h_sc, sizeof(double) * D_SIZE * 2) == cudaSuccess;
    for (auto &e : w->ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__); ws_free(w); return UPOL_ERR_ALLOC; }
    s->min_ws = w;
    *out = w;
    return UPOL_OK;
}

int minimize_compact(upol_system *s, double *x, const upol_min_options *o, upol_min_result *res, cudaStream_t st) {
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

    if (n > 0) init_pinv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, s->row_k, w->pinv);
    init_state_kernel<<<1, 1, 0, st>>>(w->ic, w->sc, o->tol, mixed ? kGatePhaseMixed : kGatePhase64);
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

    auto enqueue_slot = [&]() -> int {
        int r = sys_eval_compact(s, x, eo, s->energy, w->g, st);
        if (r) return r;
        dim3 dg((unsigned)w->nblocks, (unsigned)(m + 1));
        dots_kernel<<<dg, kDotBlock, 0, st>>>(m, n, e0, e1, w->ic, w->S, w->Y, w->g, w->g0, w->p, w->pinv, w->partial);
        dots_reduce_kernel<<<((m + 1) * kDotVals + 127) / 128, 128, 0, st>>>(m, w->nblocks, w->ic, w->partial, w->pack);
        if (s->nranks > 1) { r = comm_allreduce_sum(s, w->pack, (int64_t)(m + 1) * kDotVals, st); if (r) return r; }
        control_kernel<<<1, 1, 0, st>>>(m, o->maxiter, mixed ? 1 : 0, w->ic, w->sc, w->pack, w->RS, w->YY, w->coef);
        update_kernel<<<ublocks, 256, 0, st>>>(m, n, e0, e1, w->ic, w->sc, w->coef, w->S, w->Y, x, w->x0, w->g, w->g0, w->p, w->pinv);
        if (s->nranks > 1) { r = comm_allgather_rows(s, x, st); if (r) return r; }
        UPOL_CUDA(cudaGetLastError());
        return UPOL_OK;
    };

    // Enqueue chunks of slots, polling the done flag one chunk behind.
    const int max_slots = o->maxiter * 3 + kMaxLineSearch + 8;
    int enq = 0, buf = 0;
    bool done = false;
    bool pending[2] = {false, false};
    while (!done) {
        for (int c = 0; c < check_every && enq < max_slots; ++c, ++enq) { rc = enqueue_slot(); if (rc) return rc; }
        UPOL_CUDA(cudaMemcpyAsync(w->h_ic + buf * I_SIZE, w->ic, sizeof(int) * I_SIZE, cudaMemcpyDeviceToHost, st));
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
    // final state + fp64 energy at the minimiser (x is complete on every rank)
    EvalOpts ef;
    ef.precision = UPOL_FP64;
    ef.want_energy = true; ef.want_grad = true; ef.reduce_ranks = true;
    rc = sys_eval_compact(s, x, ef, w->e_final, w->g, st);
    if (rc) return rc;
    UPOL_CUDA(cudaMemcpyAsync(w->h_ic, w->ic, sizeof(int) * I_SIZE, cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaMemcpyAsync(w->h_sc, w->e_final, sizeof(double) * UPOL_E_SIZE, cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < UPOL_E_SIZE; ++i) res->energy[i] = w->h_sc[i];
    res->gnorm = std::sqrt(res->energy[UPOL_E_GNORM2]);
    res->iterations = w->h_ic[I_ITER];
    res->evaluations = w->h_ic[I_NEVAL] + 1;
    res->flags = w->h_ic[I_FLAGS];
    if (!(w->h_ic[I_GATE] & kGateDone)) res->flags |= UPOL_FLAG_NOT_CONVERGED;
    if (!(res->gnorm <= o->tol) && !(res->flags & UPOL_FLAG_NONFINITE)) res->flags |= UPOL_FLAG_NOT_CONVERGED;
    else if (res->gnorm <= o->tol) res->flags &= ~UPOL_FLAG_NOT_CONVERGED;
    return res->flags;
}

}  // namespace upol
