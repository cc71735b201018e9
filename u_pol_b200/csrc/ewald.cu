// Periodic boundaries (SURVEY 8f "next" #3, second half): U_ind and its gradient for an orthorhombic
// periodic box by direct Ewald summation.  The reference is gas phase only
// (benchmarks/OpenMM/water/calc_energy.py:80), so this is an extension of polarization.py:63-151, not a
// restatement; its definition and its checker are oracle/ewald_dense.py.
//
// The evaluation keeps the row structure of the open-boundary path (rows = Drude shells; per row the
// potential of the cores, the potential of the other shells, the field, and the static potential that
// is subtracted row by row) and only replaces the pair function 1/r by the periodic Green's function
// with intra-molecular exclusions:
//
//   real space   erfc(kappa r)/r over the nearest image of every charge of another molecule
//                (kappa >= 13/L_min makes further images < 1e-19), tiled all-pairs like pair_fp64_kernel;
//   reciprocal   (8 pi/V) sum_{k in half space} exp(-k^2/4kappa^2)/k^2 Re[S(k)* e^{ik.x}], with the
//                structure factors of cores, shells and shells-at-their-cores formed once per evaluation;
//   excluded     -erf(kappa r)/r for the charges of the row's own molecule (they are inside S(k)).
//
// The partial sums go into the same part[split][comp][row] slots that finalize_rows_kernel /
// finalize_static_kernel fold, so springs, Thole pairs, K*qs and the energy assembly are shared.
// First correct version: fp64 only, O(N^2) real space + O(N N_k) reciprocal space (no PME); rows shard
// over ranks like the open path (every rank forms the structure factors of all charges itself).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "thole.cuh"

namespace upol {

struct EwaldState {
    double box[3] = {0, 0, 0};
    double kappa = 0.0;
    int nk = 0, nk_pad = 0;
    int4 *kn = nullptr;        // [nk_pad] integer reciprocal vector (padding: 0,0,0)
    double *kA = nullptr;      // [nk_pad] (8 pi/V) exp(-k^2/4 kappa^2)/k^2  (padding: 0)
    double2 *Sc = nullptr;     // [nk_pad] structure factor of the cores (charge qc)
    double2 *Ss0 = nullptr;    // [nk_pad] ... of the shells sitting on their cores (d = 0)
    double2 *Ssd = nullptr;    // [nk_pad] ... of the shells at r - d
    double2 *Sstat = nullptr;  // [nk_pad] Sc + Ss0 / 2 (static columns, core-force pass)
    double2 *Szero = nullptr;  // [nk_pad] zeros
    double4 *cols0 = nullptr;  // [nb_pad] shells at their cores: x, y, z, qs
    double2 *rec = nullptr;    // [3 * max(na_pad, nb_pad)] per-charge records of the structure-factor pass
    double2 *partial = nullptr;  // [kSfacChunks][nk_pad] per-chunk partial structure factors
    bool sfac_valid = false;   // Sc, Ss0, cols0 match the current core positions
    int *mol_site0 = nullptr, *mol_nsite = nullptr;   // [nmol] site interval of a molecule
    int *bad = nullptr;        // [1] set by the geometry check
};

struct EwaldArgs {
    double Lx, Ly, Lz, iLx, iLy, iLz;
    double kappa, c2;          // c2 = 2 kappa / sqrt(pi)
    double cut2;               // r^2 beyond which erfc(kappa r) < 2.2e-17 (below fp64 resolution of the row sums)
};

namespace {

constexpr int kSfacChunks = 16;       // charge chunks of the structure-factor pass (partials per k)
constexpr int kRun = 16;              // reciprocal vectors per run of consecutive n_z (rows kernel)

// nearest image: dx - L * round(dx / L); the round-to-nearest-integer is the add/subtract of 1.5 * 2^52
// (two full-rate DADDs, exact for |dx / L| < 2^51) instead of FRND.F64, which is a conversion-pipe op
__device__ __forceinline__ double wrap(double dx, double L, double iL) {
    const double t = dx * iL;
    const double n = (t + 6755399441055744.0) - 6755399441055744.0;
    return fma(-L, n, dx);
}

// g = erf(kappa r)/r and c with grad g(|x|) = c * x; series below kappa r = 0.05 (also r = 0).
__device__ __forceinline__ void erf_over_r(double r2, double kappa, double c2, double &g, double &c) {
    const double x2 = kappa * kappa * r2;
    if (x2 < 2.5e-3) {
        g = c2 * (1.0 + x2 * (-1.0 / 3 + x2 * (1.0 / 10 + x2 * (-1.0 / 42 + x2 * (1.0 / 216 + x2 * (-1.0 / 1320))))));
        c = c2 * kappa * kappa *
            (-2.0 / 3 + x2 * (2.0 / 5 + x2 * (-1.0 / 7 + x2 * (1.0 / 27 + x2 * (-1.0 / 132 + x2 * (1.0 / 780))))));
    } else {
        const double r = sqrt(r2);
        g = erf(kappa * r) / r;
        c = (c2 * exp(-x2) - g) / r2;
    }
}

// erfcx(x) = exp(x^2) erfc(x) on 0 <= x <= 6.2 as ONE polynomial of degree 18 in t = 4 / (4 + x) (mapped to
// u in [-1, 1]); coefficients fitted in 50-digit arithmetic (Chebyshev series converted to powers of u), Horner in
// fp64 stays within 9e-16 relative of the exact function over the whole range.  erfc(x) = exp(-x^2) erfcx(x) then
// costs one reciprocal, 19 FMAs and one exp instead of the library's erfc + exp (~115 -> ~70 DP instructions per
// screened pair), and every lane of a warp takes the same path (the library erfc branches on its argument).
__device__ __forceinline__ double erfcx_0_to_6(double x) {
    const double t = 4.0 / (4.0 + x);
    const double u = fma(t, 3.2903225806451615, -0.696078431372549 * 3.2903225806451615);
    double p = -3.4857342293044046e-13;
    p = fma(p, u, 2.59344629117771e-11);
    p = fma(p, u, 1.1216234676875397e-11);
    p = fma(p, u, -7.696879080610992e-10);
    p = fma(p, u, -8.849265850355584e-10);
    p = fma(p, u, 1.969172608702299e-08);
    p = fma(p, u, 6.196904809193565e-08);
    p = fma(p, u, -4.128306935973374e-07);
    p = fma(p, u, -3.3569780388182545e-06);
    p = fma(p, u, -1.8669639133269318e-06);
    p = fma(p, u, 0.00010689650697244825);
    p = fma(p, u, 0.0008919441925262346);
    p = fma(p, u, 0.00441708995797159);
    p = fma(p, u, 0.01610750725952156);
    p = fma(p, u, 0.04636062203304619);
    p = fma(p, u, 0.10846332103559755);
    p = fma(p, u, 0.20861363014313378);
    p = fma(p, u, 0.32961043419911595);
    p = fma(p, u, 0.2854341114017986);
    return p;
}

// Real space: rows x columns of other molecules, nearest image.  STATIC: potential only (comp 0).
template <bool STATIC>
__global__ void __launch_bounds__(kBlock) ewald_real_kernel(const PairArgs a, const EwaldArgs w) {
    __shared__ double4 sh[kTileCols];
    __shared__ unsigned char list[kTileCols][kBlock];     // in-range columns of each thread's row (16 KB)
    static_assert(kTileCols <= 256, "column indices are stored as bytes");
    if (a.gate != nullptr && (*a.gate & kGateDone) != 0) return;
    const double4 *__restrict__ cols = static_cast<const double4 *>(a.cols);
    const int tid = threadIdx.x;
    const int lr = blockIdx.x * kBlock + tid;
    const bool live = lr < a.n_rows;
    const int i = a.row0 + (live ? lr : 0);
    const double sx = a.rx[i], sy = a.ry[i], sz = a.rz[i];
    const int loA = a.exA_lo[i], lenA = a.exA_len[i];
    const int loB = STATIC ? 0 : a.exB_lo[i], lenB = STATIC ? 0 : a.exB_len[i];
    double phA = 0.0, phB = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    const int t0 = (int)(((int64_t)a.n_tiles * blockIdx.y) / a.n_split);
    const int t1 = (int)(((int64_t)a.n_tiles * (blockIdx.y + 1)) / a.n_split);
    for (int t = t0; t < t1; ++t) {
        const int base = t * kTileCols;
        __syncthreads();
        sh[tid] = cols[base + tid];
        __syncthreads();
        const bool isB = t >= a.n_tiles_a;
        const int lo = (isB ? loB : loA) - base, len = isB ? lenB : lenA;
        // Two phases per tile.  Only ~40 % of the pairs lie inside the screening range, but in a fused loop
        // nearly every warp iteration has SOME lane in range, so every pair pays for the special functions.
        // Phase 1 only measures distances and lets each thread note the in-range columns of its row (one byte each,
        // list[slot][thread]: conflict-free); phase 2 runs the expensive part over the thread's own list, so a warp
        // iterates max-over-lanes(count) ~ 50 % of the tile instead of 100 %.  Same columns in the same order per
        // row: identical sums.
        int cnt = 0;
#pragma unroll 4
        for (int j = 0; j < kTileCols; ++j) {
            const double4 c = sh[j];
            const double dx = wrap(c.x - sx, w.Lx, w.iLx), dy = wrap(c.y - sy, w.Ly, w.iLy), dz = wrap(c.z - sz, w.Lz, w.iLz);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool ok = !((unsigned)(j - lo) < (unsigned)len) && r2 > kZeroR2 && r2 < w.cut2 && c.w != 0.0;
            if (ok) { list[cnt][tid] = (unsigned char)j; ++cnt; }
        }
        double pt = 0.0;
        for (int q = 0; q < cnt; ++q) {
            const double4 c = sh[list[q][tid]];
            const double dx = wrap(c.x - sx, w.Lx, w.iLx), dy = wrap(c.y - sy, w.Ly, w.iLy), dz = wrap(c.z - sz, w.Lz, w.iLz);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            // 1/r from the MUFU seed + one third-order step (as the open-boundary kernels), erfc(x) = exp(-x^2) erfcx(x)
            const double inv = rsqrt_f64(r2), x = w.kappa * (r2 * inv);
            const double ex = exp(-(w.kappa * w.kappa) * r2);
            const double ec = ex * erfcx_0_to_6(x);
            const double pot = c.w * ec * inv;
            pt += pot;
            if (!STATIC) {
                const double wq = (pot + c.w * w.c2 * ex) * (inv * inv);
                fx = fma(wq, dx, fx); fy = fma(wq, dy, fy); fz = fma(wq, dz, fz);
            }
        }
        if (isB) phB += pt; else phA += pt;
    }
    if (live) {
        double *out = a.part + (int64_t)blockIdx.y * kPartComps * a.part_stride;
        out[lr] = phA;
        if (!STATIC) {
            out[1 * a.part_stride + lr] = phB;
            out[2 * a.part_stride + lr] = fx; out[3 * a.part_stride + lr] = fy; out[4 * a.part_stride + lr] = fz;
        }
    }
}

// Structure factors S(k) = sum_j q_j exp(i k.x_j) of one charge set, in three launches:
//   prep    per charge: fractional coordinates (reduced to [-1/2, 1/2]), q, and the z step exp(2 pi i f_z);
//   partial threads own charges, blocks own (charge chunk, range of runs): per run of kRun consecutive
//           n_z one sincospi per charge, then the phase advances by complex multiplication while
//           2 * kRun register accumulators collect q * phase; the block folds them through shared
//           memory in fixed order and writes one partial per (chunk, k);
//   fold    S(k) = sum over chunks (fixed order: deterministic).
__global__ void ewald_prep_kernel(int ncols, const double4 *__restrict__ cols, const EwaldArgs w, const int *__restrict__ gate,
                                  double2 *__restrict__ rec) {
    if (gate != nullptr && (*gate & kGateDone) != 0) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const double4 c = cols[j];
    double fx = c.x * w.iLx, fy = c.y * w.iLy, fz = c.z * w.iLz;
    fx -= rint(fx); fy -= rint(fy); fz -= rint(fz);
    double wr, wi;
    sincospi(2.0 * fz, &wi, &wr);
    rec[3 * (int64_t)j] = make_double2(fx, fy);
    rec[3 * (int64_t)j + 1] = make_double2(fz, c.w);
    rec[3 * (int64_t)j + 2] = make_double2(wr, wi);
}

constexpr int kSfacPad = 129;      // row stride of the fold buffer (bank-conflict free column reads)

__global__ void __launch_bounds__(128)
ewald_sfac_partial_kernel(int ncols, const double2 *__restrict__ rec, int nk_pad, const int4 *__restrict__ kn,
                          const int *__restrict__ gate, double2 *__restrict__ partial) {
    __shared__ double fold[2 * kRun][kSfacPad];
    __shared__ double quart[4][2 * kRun];
    if (gate != nullptr && (*gate & kGateDone) != 0) return;
    const int tid = threadIdx.x;
    const int per = (ncols + gridDim.x - 1) / gridDim.x;
    const int c0 = blockIdx.x * per, c1 = min(ncols, c0 + per);
    const int nrun = nk_pad / kRun;
    const int r0 = (int)(((int64_t)nrun * blockIdx.y) / gridDim.y), r1 = (int)(((int64_t)nrun * (blockIdx.y + 1)) / gridDim.y);
    for (int run = r0; run < r1; ++run) {
        const int4 n = kn[run * kRun];
        const double nx = n.x, ny = n.y, nz = n.z;
        double are[kRun], aim[kRun];
#pragma unroll
        for (int u = 0; u < kRun; ++u) are[u] = aim[u] = 0.0;
        for (int j = c0 + tid; j < c1; j += 128) {
            const double2 a = rec[3 * (int64_t)j], b = rec[3 * (int64_t)j + 1], st = rec[3 * (int64_t)j + 2];
            double si, co;
            sincospi(2.0 * fma(nx, a.x, fma(ny, a.y, nz * b.x)), &si, &co);
            co *= b.y; si *= b.y;                      // carry q inside the phase
#pragma unroll
            for (int u = 0; u < kRun; ++u) {
                are[u] += co; aim[u] += si;
                const double t2 = fma(co, st.x, -si * st.y);
                si = fma(co, st.y, si * st.x);
                co = t2;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kRun; ++u) { fold[u][tid] = are[u]; fold[kRun + u][tid] = aim[u]; }
        __syncthreads();
        {
            const int v = tid & (2 * kRun - 1), q = tid / (2 * kRun);     // 128 threads = 32 values x 4 quarters
            double sum = 0.0;
            for (int t = 0; t < 32; ++t) sum += fold[v][q * 32 + t];
            quart[q][v] = sum;
        }
        __syncthreads();
        if (tid < kRun) {
            const double re = (quart[0][tid] + quart[1][tid]) + (quart[2][tid] + quart[3][tid]);
            const double im = (quart[0][kRun + tid] + quart[1][kRun + tid]) + (quart[2][kRun + tid] + quart[3][kRun + tid]);
            partial[(int64_t)blockIdx.x * nk_pad + run * kRun + tid] = make_double2(re, im);
        }
    }
}

__global__ void ewald_sfac_fold_kernel(int nchunk, int nk_pad, const double2 *__restrict__ partial, const int *__restrict__ gate,
                                       double2 *__restrict__ out) {
    if (gate != nullptr && (*gate & kGateDone) != 0) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk_pad) return;
    double re = 0.0, im = 0.0;
    for (int c = 0; c < nchunk; ++c) { const double2 p = partial[(int64_t)c * nk_pad + k]; re += p.x; im += p.y; }
    out[k] = make_double2(re, im);
}

// Molecules must be whole (not wrapped atom by atom) and smaller than half the box: the excluded-pair
// terms use the direct intra-molecular distance and the real-space sum the nearest image only.
__global__ void ewald_check_molecules_kernel(int nmol, const int *__restrict__ site0, const int *__restrict__ nsite,
                                             const double *__restrict__ r, double hx, double hy, double hz,
                                             int *__restrict__ bad) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmol) return;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = site0[m]; i < site0[m] + nsite[m]; ++i)
        for (int c = 0; c < 3; ++c) { const double v = r[3 * i + c]; lo[c] = fmin(lo[c], v); hi[c] = fmax(hi[c], v); }
    const bool ok = nsite[m] == 0 || (hi[0] - lo[0] < hx && hi[1] - lo[1] < hy && hi[2] - lo[2] < hz &&
                                      isfinite(hi[0]) && isfinite(hi[1]) && isfinite(hi[2]));
    if (!ok) atomicOr(bad, 1);
}

__global__ void ewald_cols0_kernel(int nd, int nb_pad, const double *__restrict__ x, const double *__restrict__ y,
                                   const double *__restrict__ z, const double *__restrict__ row_qs,
                                   double4 *__restrict__ cols0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb_pad) return;
    cols0[i] = i < nd ? make_double4(x[i], y[i], z[i], row_qs[i]) : make_double4(0.0, 0.0, 0.0, 0.0);
}

__device__ __forceinline__ void kahan_add(double &sum, double &comp, double term) {
    const double y = term - comp;
    const double t = sum + y;
    comp = (t - sum) - y;
    sum = t;
}

// Reciprocal-space potential / field at the rows, k range split over gridDim.y; block row y = 0 also
// removes the row's own molecule (excluded pairs), which the structure factors contain.
template <bool STATIC>
__global__ void __launch_bounds__(128)
ewald_rows_kernel(const PairArgs a, const EwaldArgs w, int nk_pad, const int4 *__restrict__ kn,
                  const double *__restrict__ kA, const double2 *__restrict__ Sc, const double2 *__restrict__ Ss,
                  const double4 *__restrict__ cols, const double4 *__restrict__ cols_static, int na_pad, int slot0,
                  bool with_b) {
    __shared__ double4 s_h[128 / kRun];          // per run: nx, ny, first nz
    __shared__ double s_A[128];
    __shared__ double2 s_c[128], s_s[128];
    if (a.gate != nullptr && (*a.gate & kGateDone) != 0) return;
    const int tid = threadIdx.x;
    const int lr = blockIdx.x * 128 + tid;
    const bool live = lr < a.n_rows;
    const int i = a.row0 + (live ? lr : 0);
    const double sx = a.rx[i], sy = a.ry[i], sz = a.rz[i];
    double f0 = sx * w.iLx, f1 = sy * w.iLy, f2 = sz * w.iLz;
    f0 -= rint(f0); f1 -= rint(f1); f2 -= rint(f2);
    double wr, wi;                                // one step along z: exp(2 pi i f_z)
    sincospi(2.0 * f2, &wi, &wr);
    double pa = 0.0, pb = 0.0, ca = 0.0, cb = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    const int ntile = nk_pad / 128;
    const int k0 = (int)(((int64_t)ntile * blockIdx.y) / gridDim.y), k1 = (int)(((int64_t)ntile * (blockIdx.y + 1)) / gridDim.y);
    for (int t = k0; t < k1; ++t) {
        __syncthreads();
        s_A[tid] = kA[t * 128 + tid];
        s_c[tid] = Sc[t * 128 + tid]; s_s[tid] = Ss[t * 128 + tid];
        if (tid < 128 / kRun) {
            const int4 n = kn[t * 128 + tid * kRun];
            s_h[tid] = make_double4((double)n.x, (double)n.y, (double)n.z, 0.0);
        }
        __syncthreads();
        for (int g = 0; g < 128 / kRun; ++g) {
            // a run of kRun consecutive n_z at fixed (n_x, n_y): one sincospi, then the phase advances by
            // complex multiplication (kRun steps of 1 ulp each; re-seeded at the next run)
            const double4 h = s_h[g];
            double si, co;
            sincospi(2.0 * fma(h.x, f0, fma(h.y, f1, h.z * f2)), &si, &co);
            double lg = 0.0, lgz = 0.0, nz = h.z;
#pragma unroll
            for (int u = 0; u < kRun; ++u) {
                const int j = g * kRun + u;
                const double A = s_A[j];
                const double2 c = s_c[j], sh = s_s[j];
                // Kahan sums: each potential carries a self-like part of ~2 kappa/sqrt(pi) per unit charge
                // that the static pass / the excluded-pair terms cancel again, so over 10^4-10^6 reciprocal
                // vectors a plain running sum would leave sqrt(N_k) * eps of that part in U_ind
                kahan_add(pa, ca, A * fma(c.x, co, c.y * si));
                kahan_add(pb, cb, A * fma(sh.x, co, sh.y * si));
                if (!STATIC) {
                    const double gq = A * fma(c.y + sh.y, co, -(c.x + sh.x) * si);
                    lg += gq; lgz = fma(gq, nz, lgz); nz += 1.0;
                }
                const double t2 = fma(co, wr, -si * wi);
                si = fma(co, wi, si * wr);
                co = t2;
            }
            if (!STATIC) { gx = fma(h.x, lg, gx); gy = fma(h.y, lg, gy); gz += lgz; }
        }
    }
    const double two_pi = 2.0 * kPi;
    gx *= two_pi * w.iLx; gy *= two_pi * w.iLy; gz *= two_pi * w.iLz;
    if (live && blockIdx.y == 0) {
        const int loA = a.exA_lo[i], lenA = a.exA_len[i];
        if (STATIC) {
            // the row's own core sits at distance zero: erf(kappa r)/r -> 2 kappa/sqrt(pi); its own shell
            // is the self term, left out here and in the displaced configuration alike
            for (int b = loA; b < loA + lenA; ++b) {
                const double4 c = cols_static[b];
                const double x = sx - c.x, y = sy - c.y, z = sz - c.z;
                const double r2 = fma(z, z, fma(y, y, x * x));
                double g, cc;
                erf_over_r(r2, w.kappa, w.c2, g, cc);
                pa -= (r2 == 0.0 ? cols[b].w : c.w) * g;
            }
        } else {
            for (int b = loA; b < loA + lenA; ++b) {
                const double4 c = cols[b];
                const double x = sx - c.x, y = sy - c.y, z = sz - c.z;
                double g, cc;
                erf_over_r(fma(z, z, fma(y, y, x * x)), w.kappa, w.c2, g, cc);
                pa -= c.w * g;
                gx -= c.w * cc * x; gy -= c.w * cc * y; gz -= c.w * cc * z;
            }
            const int loB = a.exB_lo[i], lenB = with_b ? a.exB_len[i] : 0, own = na_pad + i;
            for (int b = loB; b < loB + lenB; ++b) {
                if (b == own) continue;
                const double4 c = cols[b];
                const double x = sx - c.x, y = sy - c.y, z = sz - c.z;
                double g, cc;
                erf_over_r(fma(z, z, fma(y, y, x * x)), w.kappa, w.c2, g, cc);
                pb -= c.w * g;
                gx -= c.w * cc * x; gy -= c.w * cc * y; gz -= c.w * cc * z;
            }
        }
    }
    if (live) {
        double *out = a.part + (int64_t)(slot0 + blockIdx.y) * kPartComps * a.part_stride;
        if (STATIC) {
            out[lr] = pa + 0.5 * pb;
        } else {
            out[lr] = pa;
            out[1 * a.part_stride + lr] = pb;
            out[2 * a.part_stride + lr] = gx; out[3 * a.part_stride + lr] = gy; out[4 * a.part_stride + lr] = gz;
        }
    }
}

__global__ void ewald_sstat_kernel(int nk_pad, const double2 *__restrict__ Sc, const double2 *__restrict__ Ss0,
                                   double2 *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nk_pad) out[k] = make_double2(Sc[k].x + 0.5 * Ss0[k].x, Sc[k].y + 0.5 * Ss0[k].y);
}

inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

EwaldArgs make_args(const EwaldState *e) {
    EwaldArgs w;
    w.Lx = e->box[0]; w.Ly = e->box[1]; w.Lz = e->box[2];
    w.iLx = 1.0 / w.Lx; w.iLy = 1.0 / w.Ly; w.iLz = 1.0 / w.Lz;
    w.kappa = e->kappa;
    w.c2 = 2.0 * e->kappa / std::sqrt(kPi);
    w.cut2 = (6.0 / e->kappa) * (6.0 / e->kappa);      // erfc(6) = 2e-17 of the bare 1/r
    return w;
}

// real-space column splits and reciprocal k chunks of one launch, fitted into the part buffer
void choose_slots(const upol_system *s, const EwaldState *e, int n_rows, int n_tiles, int64_t stride, int &s_real, int &s_rec) {
    const int row_blocks = std::max(1, (n_rows + kBlock - 1) / kBlock);
    s_rec = std::max(1, std::min(std::min(16, e->nk_pad / 128), (148 * 2 + row_blocks - 1) / row_blocks));
    s_real = choose_split(s, n_rows, n_tiles, kBlock);
    const int64_t cap = (int64_t)(s->part_elems / ((size_t)kPartComps * (size_t)stride));
    s_rec = (int)std::max<int64_t>(1, std::min<int64_t>(s_rec, cap - 1));
    s_real = (int)std::max<int64_t>(1, std::min<int64_t>(s_real, std::min<int64_t>(cap - s_rec, s->max_split - s_rec)));
}

}  // namespace

void ewald_free(upol_system *s) {
    EwaldState *e = static_cast<EwaldState *>(s->ewald);
    if (!e) return;
    void *ptrs[] = {e->kn, e->kA, e->Sc, e->Ss0, e->Ssd, e->Sstat, e->Szero, e->cols0, e->rec, e->partial, e->mol_site0, e->mol_nsite, e->bad};
    for (void *p : ptrs) if (p) upol::dfree(p);
    delete e;
    s->ewald = nullptr;
}

void ewald_invalidate(upol_system *s) {
    if (s->ewald) static_cast<EwaldState *>(s->ewald)->sfac_valid = false;
}

int ewald_setup(upol_system *s, const double *box, double kappa, int nmax) {
    ewald_free(s);
    s->static_valid = false;
    if (!box) return UPOL_OK;                       // back to open boundaries
    if (!(box[0] > 0.0 && box[1] > 0.0 && box[2] > 0.0) || nmax < 0) return UPOL_ERR_ARG;
    const double lmin = std::min(box[0], std::min(box[1], box[2]));
    if (!(kappa > 0.0)) kappa = 13.0 / lmin;        // erfc(kappa L/2) = erfc(6.5): nearest image suffices
    if (kappa * lmin < 12.0) return UPOL_ERR_ARG;   // the real-space sum only visits the nearest image
    EwaldState *e = new EwaldState();
    s->ewald = e;
    for (int c = 0; c < 3; ++c) e->box[c] = box[c];
    e->kappa = kappa;
    // half space of reciprocal vectors, exp(-k^2/4 kappa^2) >= exp(-41.5) unless a cube |n| <= nmax is asked for
    const double kmax = 2.0 * kappa * std::sqrt(41.5);
    int lim[3];
    for (int c = 0; c < 3; ++c) lim[c] = nmax > 0 ? nmax : (int)std::floor(kmax * box[c] / (2.0 * kPi));
    // refuse set-ups whose reciprocal list would not fit comfortably (a cube needs ~5e4 vectors)
    if ((double)(lim[0] + 1) * (2.0 * lim[1] + 1) * (2.0 * lim[2] + 1) > 3.2e7) { ewald_free(s); return UPOL_ERR_ARG; }
    const double vol = box[0] * box[1] * box[2];
    // Lines of consecutive n_z at fixed (n_x, n_y), each padded to a multiple of kRun with zero-weight
    // entries that continue the line, so that the rows kernel can advance the phase by recurrence.
    std::vector<int4> kn;
    std::vector<double> kA;
    int n_true = 0;
    for (int nx = 0; nx <= lim[0]; ++nx)
        for (int ny = (nx == 0 ? 0 : -lim[1]); ny <= lim[1]; ++ny) {
            int z_lo = 0, z_hi = -1;
            bool any = false;
            for (int nz = ((nx == 0 && ny == 0) ? 1 : -lim[2]); nz <= lim[2]; ++nz) {
                const double kx = 2.0 * kPi * nx / box[0], ky = 2.0 * kPi * ny / box[1], kz = 2.0 * kPi * nz / box[2];
                if (nmax == 0 && kx * kx + ky * ky + kz * kz > kmax * kmax) continue;
                if (!any) { z_lo = nz; any = true; }
                z_hi = nz;
            }
            if (!any) continue;
            const int len = z_hi - z_lo + 1, padded = (int)round_up(len, kRun);
            for (int u = 0; u < padded; ++u) {
                const int nz = z_lo + u;
                const double kx = 2.0 * kPi * nx / box[0], ky = 2.0 * kPi * ny / box[1], kz = 2.0 * kPi * nz / box[2];
                const double k2 = kx * kx + ky * ky + kz * kz;
                kn.push_back(make_int4(nx, ny, nz, 0));
                kA.push_back(u < len ? 8.0 * kPi / vol * std::exp(-k2 / (4.0 * kappa * kappa)) / k2 : 0.0);
            }
            n_true += len;
        }
    e->nk = n_true;
    e->nk_pad = (int)round_up(std::max<int64_t>((int64_t)kn.size(), 1), 128);
    kn.resize(e->nk_pad, make_int4(0, 0, 0, 0));
    kA.resize(e->nk_pad, 0.0);
    auto fail = [&](int rc) { ewald_free(s); return rc; };
    if (upol::dmalloc((void **)&e->kn, sizeof(int4) * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->kA, sizeof(double) * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->Sc, sizeof(double2) * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->Ss0, sizeof(double2) * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->Ssd, sizeof(double2) * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->Sstat, sizeof(double2) * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->Szero, sizeof(double2) * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->cols0, sizeof(double4) * std::max<int64_t>(s->nb_pad, 1)) != cudaSuccess ||
        upol::dmalloc((void **)&e->rec, sizeof(double2) * 3 * std::max<int64_t>(std::max(s->na_pad, s->nb_pad), 1)) != cudaSuccess ||
        upol::dmalloc((void **)&e->partial, sizeof(double2) * (size_t)kSfacChunks * e->nk_pad) != cudaSuccess ||
        upol::dmalloc((void **)&e->mol_site0, sizeof(int) * std::max<int64_t>(s->nmol, 1)) != cudaSuccess ||
        upol::dmalloc((void **)&e->mol_nsite, sizeof(int) * std::max<int64_t>(s->nmol, 1)) != cudaSuccess ||
        upol::dmalloc((void **)&e->bad, sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        return fail(UPOL_ERR_ALLOC);
    }
    if (cudaMemset(e->Szero, 0, sizeof(double2) * e->nk_pad) != cudaSuccess ||
        cudaMemcpy(e->kn, kn.data(), sizeof(int4) * e->nk_pad, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(e->kA, kA.data(), sizeof(double) * e->nk_pad, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(e->mol_site0, s->h_mol_site0.data(), sizeof(int) * s->nmol, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(e->mol_nsite, s->h_mol_nsite.data(), sizeof(int) * s->nmol, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {      // pageable copies on the legacy stream must land before use on any stream
        set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
        return fail(UPOL_ERR_CUDA);
    }
    return UPOL_OK;
}

int ewald_info(const upol_system *s, double *box, double *kappa, int *nk) {
    const EwaldState *e = static_cast<const EwaldState *>(s->ewald);
    if (!e) return UPOL_ERR_ARG;
    if (box) for (int c = 0; c < 3; ++c) box[c] = e->box[c];
    if (kappa) *kappa = e->kappa;
    if (nk) *nk = e->nk;
    return UPOL_OK;
}

static int launch_sfac(EwaldState *e, int ncols, const double4 *cols, const EwaldArgs &w, const int *gate,
                       double2 *out, cudaStream_t st) {
    if (ncols <= 0) {
        UPOL_CUDA(cudaMemsetAsync(out, 0, sizeof(double2) * e->nk_pad, st));
        return UPOL_OK;
    }
    ewald_prep_kernel<<<nblk(ncols, 256), 256, 0, st>>>(ncols, cols, w, gate, e->rec);
    const int chunks = std::max(1, std::min(kSfacChunks, (ncols + 1023) / 1024));
    const int nrun = e->nk_pad / kRun;
    const int ry = std::max(1, std::min(nrun, (148 * 4 + chunks - 1) / chunks));
    dim3 grid((unsigned)chunks, (unsigned)ry);
    ewald_sfac_partial_kernel<<<grid, 128, 0, st>>>(ncols, e->rec, e->nk_pad, e->kn, gate, e->partial);
    ewald_sfac_fold_kernel<<<nblk(e->nk_pad, 256), 256, 0, st>>>(chunks, e->nk_pad, e->partial, gate, out);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

// structure factors that depend on the core positions only
static int ewald_static_sfac(upol_system *s, EwaldState *e, const EwaldArgs &w, cudaStream_t st) {
    if (e->sfac_valid) return UPOL_OK;
    // once per geometry: refuse molecules that are broken across the cell or too large for it
    UPOL_CUDA(cudaMemsetAsync(e->bad, 0, sizeof(int), st));
    ewald_check_molecules_kernel<<<nblk(s->nmol, 128), 128, 0, st>>>((int)s->nmol, e->mol_site0, e->mol_nsite, s->r,
                                                                     0.5 * e->box[0], 0.5 * e->box[1], 0.5 * e->box[2], e->bad);
    int bad = 0;
    UPOL_CUDA(cudaMemcpyAsync(&bad, e->bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaStreamSynchronize(st));
    if (bad) return UPOL_ERR_ARG;
    ewald_cols0_kernel<<<nblk(s->nb_pad, 256), 256, 0, st>>>((int)s->nd, (int)s->nb_pad, s->srowx, s->srowy, s->srowz,
                                                             s->row_qs, e->cols0);
    UPOL_CUDA(cudaGetLastError());
    int rc = launch_sfac(e, (int)s->na_pad, s->cols, w, nullptr, e->Sc, st);
    if (rc) return rc;
    rc = launch_sfac(e, (int)s->nb_pad, e->cols0, w, nullptr, e->Ss0, st);
    if (rc) return rc;
    e->sfac_valid = true;
    return UPOL_OK;
}

// Fills part[0 .. *n_slots) with the static potential partials of the rows (comp 0).
int ewald_static_part(upol_system *s, PairArgs a, int *n_slots, cudaStream_t st) {
    EwaldState *e = static_cast<EwaldState *>(s->ewald);
    const EwaldArgs w = make_args(e);
    int rc = ewald_static_sfac(s, e, w, st);
    if (rc) return rc;
    int s_real, s_rec;
    choose_slots(s, e, a.n_rows, a.n_tiles, a.part_stride, s_real, s_rec);
    a.n_split = s_real;
    a.gate = nullptr;
    dim3 g1(nblk(a.n_rows, kBlock), (unsigned)s_real);
    ewald_real_kernel<true><<<g1, kBlock, 0, st>>>(a, w);
    dim3 g2(nblk(a.n_rows, 128), (unsigned)s_rec);
    ewald_rows_kernel<true><<<g2, 128, 0, st>>>(a, w, e->nk_pad, e->kn, e->kA, e->Sc, e->Ss0, s->cols, s->cols_static,
                                                (int)s->na_pad, s_real, true);
    UPOL_CUDA(cudaGetLastError());
    *n_slots = s_real + s_rec;
    return UPOL_OK;
}

// Fills part[0 .. *n_slots) with potential (comps 0,1) and field (comps 2-4) partials of the rows at
// the current shell positions (a.rx/ry/rz, shell columns already prepared).
int ewald_dynamic_part(upol_system *s, PairArgs a, int *n_slots, cudaStream_t st) {
    EwaldState *e = static_cast<EwaldState *>(s->ewald);
    const EwaldArgs w = make_args(e);
    int rc = ewald_static_sfac(s, e, w, st);
    if (rc) return rc;
    rc = launch_sfac(e, (int)s->nb_pad, s->cols + s->na_pad, w, a.gate, e->Ssd, st);
    if (rc) return rc;
    int s_real, s_rec;
    choose_slots(s, e, a.n_rows, a.n_tiles, a.part_stride, s_real, s_rec);
    a.n_split = s_real;
    dim3 g1(nblk(a.n_rows, kBlock), (unsigned)s_real);
    ewald_real_kernel<false><<<g1, kBlock, 0, st>>>(a, w);
    dim3 g2(nblk(a.n_rows, 128), (unsigned)s_rec);
    ewald_rows_kernel<false><<<g2, 128, 0, st>>>(a, w, e->nk_pad, e->kn, e->kA, e->Sc, e->Ssd, s->cols, s->cols_static,
                                                 (int)s->na_pad, s_real, true);
    UPOL_CUDA(cudaGetLastError());
    *n_slots = s_real + s_rec;
    return UPOL_OK;
}

// Core-force pass: field (comps 2-4 of part[0 .. *n_slots)) at the rows of `a` from ONE column set
// (a.cols, a.n_tiles tiles, excluded interval a.exA_*), whose structure factor is `which`:
// 0 shells at r - d, 1 shells on their cores, 2 static columns (qc + qs/2 on the cores).
int ewald_field_part(upol_system *s, PairArgs a, int which, int *n_slots, cudaStream_t st) {
    EwaldState *e = static_cast<EwaldState *>(s->ewald);
    const EwaldArgs w = make_args(e);
    int rc = ewald_static_sfac(s, e, w, st);
    if (rc) return rc;
    const double2 *S = which == 0 ? e->Ssd : which == 1 ? e->Ss0 : e->Sstat;
    if (which == 2) ewald_sstat_kernel<<<nblk(e->nk_pad, 256), 256, 0, st>>>(e->nk_pad, e->Sc, e->Ss0, e->Sstat);
    a.n_tiles_a = a.n_tiles;                          // a single section
    a.exB_lo = a.exA_lo; a.exB_len = a.exA_len;       // never read: no B tiles, with_b = false
    a.gate = nullptr;
    int s_real, s_rec;
    choose_slots(s, e, a.n_rows, a.n_tiles, a.part_stride, s_real, s_rec);
    a.n_split = s_real;
    dim3 g1(nblk(a.n_rows, kBlock), (unsigned)s_real);
    ewald_real_kernel<false><<<g1, kBlock, 0, st>>>(a, w);
    dim3 g2(nblk(a.n_rows, 128), (unsigned)s_rec);
    ewald_rows_kernel<false><<<g2, 128, 0, st>>>(a, w, e->nk_pad, e->kn, e->kA, S, e->Szero, static_cast<const double4 *>(a.cols),
                                                 s->cols_static, 0, s_real, false);
    UPOL_CUDA(cudaGetLastError());
    *n_slots = s_real + s_rec;
    return UPOL_OK;
}

}  // namespace upol
