// Mixed mode (UPOL_MIXED): fp32 pair arithmetic, fp64 accumulation, for the ENERGY and the field.
//
// Replaces, in one pass over site-paired columns, the inter-molecular part of Ucoul
// (U_pol/polarization.py:82-91), the static subtraction Ucoul_static (:50-61, :149) and the reverse-mode
// gradient jax.grad(Uind, 1) taken inside jaxopt.BFGS (:172-179).
//
// U_ind is the small difference U_coul - U_coul_static of sums that are 1e2..1e6 times larger, so fp32
// pair terms q/|x| cannot deliver it.  The kernel therefore never forms them: with, for row i (Drude
// shell at s_i = r_i - d_i) and site j of another molecule,
//     A = r_j - s_i,   R = r_j - r_i = A - d_i,   C = s_j - s_i = A - d_j,           f(X) = 1/|X|,
// the row's share of (U_coul - U_coul_static) is  K qs_i sum_j { qc_j [f(A)-f(R)] + 1/2 qs_j [f(C)-f(R)] }
// and every bracket is evaluated in the cancellation-free form ("Cancellation")
//     f(X) - f(R) = (|R|^2 - |X|^2) / (|X| |R| (|X| + |R|)),
//     |A|^2 - |R|^2 = 2 A.d_i - |d_i|^2,      |C|^2 - |A|^2 = -2 A.d_j + |d_j|^2,
// i.e. the differences of squares come from dot products with the (small) displacements, never from
// subtracting two norms.  A itself is an EXACT difference of 32-bit fixed-point coordinates (grid
// 2^-k nm about the box centre) converted once to fp32; the displacements enter as fp32 vectors in grid
// units.  |R|^2 = |A|^2 - (|A|^2-|R|^2) and |C|^2 = |A|^2 + (|C|^2-|A|^2) cost one add each, so the three
// norms of a site pair need 3 + 3 + 3 FMAs, not 9 + 6 subtractions.
//
// Field (gradient): q_c A/|A|^3 + q_s C/|C|^3 with C = A - d_j is accumulated as (w_A + w_C) A - w_C d_j;
// the two nearly cancelling weights meet before they touch an accumulator, and C - A = -d_j holds to
// fp32 precision of d_j by construction (the older field kernel formed A and C from two independently
// rounded conversions).  w = q y^3 (1 + 1.5 e), e = 1 - |X|^2 y^2, removes the 3 x 2^-22.9 error of the
// MUFU.RSQ seed y.  The energy needs no such correction: in yA yR / (|A|^2 yA + |R|^2 yR) the seed errors
// cancel to first order (net (eA + eR)/2).
//
// Columns (built by system.cu): section P = polarisable sites in Drude-row order, section N = the other
// sites, each padded to the 128-column tile with zero-charge far-away dummies;
//   cols_x [P|N] int4   fixed-point core position, float bits of q_core
//   cols_d [P]   float4 -2 d_j (grid units), q_shell           (rewritten by every evaluation)
//   cols_d2[P]   float  |d_j|^2
// Rows = Drude shells (fixed-point s_i, fp32 2 d_i and -|d_i|^2 in registers), four per thread, processed
// as two fp32 pairs (FFMA2/FMUL2/FADD2: half the issue slots of scalar code).  Same-molecule columns of a
// row are one index interval per section; tiles overlapping an interval of some row of the warp, and tiles
// in which a distance vanished (the reference drops such terms, polarization.py:33-42; here: below
// 1e-5 nm, as in the fp64 kernels), take a predicated scalar loop - the packed loop is speculative and
// its tile-local sums are discarded when they come out non-finite or absurdly large.
//
// FMA-pipe instructions per (row, polarisable site): 45 with the energy, 26 field only; per (row,
// other site) 23 / 12; MUFU 4 / 2 and 3 / 1.
#include <cstdlib>

#include "common.cuh"
#include "f32x2.cuh"

namespace upol {

namespace {

constexpr float kDropNm = 1.0e-5f;        // distances below this count as zero (term dropped)
constexpr float kBigFieldNm = 1.0e6f;     // e/nm^2: a tile partial beyond this had a (near-)zero distance
constexpr float kBigPotNm = 1.0e4f;       // e/nm
constexpr int kFarFix = (1 << 30) - 1;    // fixed-point coordinate of the zero-charge padding columns (system.cu)

struct DiffConsts {
    float thr2;        // (kDropNm * inv_grid)^2, grid units^2
    float big_f;       // kBigFieldNm in grid units
    float big_p;       // kBigPotNm in grid units
};

// One row x one site, predicated scalar version (excluded columns, zero distances).  Adds into
// ph[0] (cores: -sum qc [f(A)-f(R)]), ph[1] (shells: -sum qs [f(C)-f(R)]) and g[3] (field).
template <bool POT, bool PSEC>
__device__ __forceinline__ void diff_site_careful(const int4 c, const float4 cd, int sx, int sy, int sz, float t2x,
                                                  float t2y, float t2z, float nd2, bool excl, const DiffConsts k,
                                                  float *ph, float *g) {
    const float Ax = (float)(c.x - sx), Ay = (float)(c.y - sy), Az = (float)(c.z - sz);
    const float A2 = fmaf(Az, Az, fmaf(Ay, Ay, Ax * Ax));
    const bool zA = A2 < k.thr2;
    const float yA = zA ? 0.f : rsqrt_f32(A2);
    const float qc = excl ? 0.f : __int_as_float(c.w);
    float yR = 0.f, R2 = 1.f, DA = 0.f;
    bool zR = false;
    if (POT) {
        const float Rx = fmaf(-0.5f, t2x, Ax), Ry = fmaf(-0.5f, t2y, Ay), Rz = fmaf(-0.5f, t2z, Az);
        R2 = fmaf(Rz, Rz, fmaf(Ry, Ry, Rx * Rx));
        zR = R2 < k.thr2;
        yR = zR ? 0.f : rsqrt_f32(R2);
        DA = fmaf(Ax, t2x, fmaf(Ay, t2y, fmaf(Az, t2z, nd2)));
        // f(A) - f(R), a dropped term counting as 0 (then the other one stands alone)
        const float tA = (zA || zR) ? (yA - yR) : -(DA * (yA * yR)) * rcp_f32(fmaf(A2, yA, R2 * yR));
        ph[0] = fmaf(qc, -tA, ph[0]);
    }
    {
        const float y2 = yA * yA, e = fmaf(-A2, y2, 1.0f), z = (qc * yA) * y2;
        const float w = zA ? 0.f : fmaf(z * e, 1.5f, z);
        g[0] = fmaf(w, Ax, g[0]); g[1] = fmaf(w, Ay, g[1]); g[2] = fmaf(w, Az, g[2]);
    }
    if (PSEC) {
        const float qs = excl ? 0.f : cd.w;
        const float Cx = fmaf(0.5f, cd.x, Ax), Cy = fmaf(0.5f, cd.y, Ay), Cz = fmaf(0.5f, cd.z, Az);
        const float C2 = fmaf(Cz, Cz, fmaf(Cy, Cy, Cx * Cx));
        const bool zC = C2 < k.thr2;
        const float yC = zC ? 0.f : rsqrt_f32(C2);
        if (POT) {
            const float dj2 = 0.25f * fmaf(cd.z, cd.z, fmaf(cd.y, cd.y, cd.x * cd.x));
            const float DC = DA + fmaf(Ax, cd.x, fmaf(Ay, cd.y, fmaf(Az, cd.z, dj2)));
            const float tC = (zC || zR) ? (yC - yR) : -(DC * (yC * yR)) * rcp_f32(fmaf(C2, yC, R2 * yR));
            ph[1] = fmaf(qs, -tC, ph[1]);
        }
        const float y2 = yC * yC, e = fmaf(-C2, y2, 1.0f), z = (qs * yC) * y2;
        const float w = zC ? 0.f : fmaf(z * e, 1.5f, z);
        g[0] = fmaf(w, Cx, g[0]); g[1] = fmaf(w, Cy, g[1]); g[2] = fmaf(w, Cz, g[2]);
    }
}

}  // namespace

// Packed tile loops.  MASKED: the tile overlaps the same-molecule interval of some row of the warp; the excluded
// (row, column) pairs are then switched off lane by lane - zero charges, and the column moved to the far corner so
// that every distance in the lane stays finite and benign (a row's own site would otherwise sit at |A| = |d_i|,
// which is exactly 0 at the start of a minimisation) - at the price of ~10 integer-pipe instructions per site; the
// arithmetic is the packed one either way.
template <int NP, bool POT, bool MASKED, int UNR>
__device__ __forceinline__ void diff_tile_P(const int4 *__restrict__ shX, const float4 *__restrict__ shD, const float *__restrict__ shD2,
                                            const float2 *__restrict__ shQ, const int *sx, const int *sy, const int *sz, const int *lo, const int *len,
                                            const f32x2 *T2x, const f32x2 *T2y, const f32x2 *T2z, const f32x2 *ND2,
                                            f32x2 *Fx, f32x2 *Fy, f32x2 *Fz, f32x2 *Hx, f32x2 *Hy, f32x2 *Hz, f32x2 *PA, f32x2 *PC) {
    const f32x2 mone2 = bcast2(-1.0f), m15 = bcast2(-1.5f);
#pragma unroll UNR
    for (int j = 0; j < kTileCols; ++j) {
        const int4 c = shX[j];
        const float4 cd = shD[j];
        const float qc = __int_as_float(c.w), qs = cd.w;
        const f32x2 mx2 = bcast2(cd.x), my2 = bcast2(cd.y), mz2 = bcast2(cd.z), dj2 = bcast2(shD2[j]);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            int cx0 = c.x, cx1 = c.x;
            f32x2 qc2 = bcast2(qc), qs2 = bcast2(qs);
            if (MASKED) {
                const bool e0 = (unsigned)(j - lo[2 * p]) < (unsigned)len[2 * p], e1 = (unsigned)(j - lo[2 * p + 1]) < (unsigned)len[2 * p + 1];
                cx0 = e0 ? kFarFix : c.x; cx1 = e1 ? kFarFix : c.x;
                qc2 = pack2(e0 ? 0.f : qc, e1 ? 0.f : qc); qs2 = pack2(e0 ? 0.f : qs, e1 ? 0.f : qs);
            }
            const f32x2 Ax = pack2((float)(cx0 - sx[2 * p]), (float)(cx1 - sx[2 * p + 1]));
            const f32x2 Ay = pack2((float)(c.y - sy[2 * p]), (float)(c.y - sy[2 * p + 1]));
            const f32x2 Az = pack2((float)(c.z - sz[2 * p]), (float)(c.z - sz[2 * p + 1]));
            const f32x2 A2 = fma2(Az, Az, fma2(Ay, Ay, mul2(Ax, Ax)));
            const f32x2 yA = rsqrt2(A2);
            const f32x2 pa = mul2(qc2, yA);
            const f32x2 DCA = fma2(Ax, mx2, fma2(Ay, my2, fma2(Az, mz2, dj2)));         // |C|^2 - |A|^2
            const f32x2 C2 = add2(A2, DCA);
            const f32x2 yC = rsqrt2(C2);
            const f32x2 pc = mul2(qs2, yC);
            if (POT) {
                const f32x2 DA = fma2(Ax, T2x[p], fma2(Ay, T2y[p], fma2(Az, T2z[p], ND2[p])));  // |A|^2 - |R|^2
                const f32x2 R2 = sub2(A2, DA);
                const f32x2 DC = add2(DA, DCA);                                         // |C|^2 - |R|^2
                const f32x2 yR = rsqrt2(R2);
                const f32x2 nR = mul2(R2, yR);
                // 1/(|A|+|R|) and 1/(|C|+|R|) from ONE reciprocal of their product: the XU pipe (MUFU) is the
                // busiest pipe of this kernel (69 %, stalls on the MIO queue), two packed multiplies are cheaper
                const f32x2 sA = fma2(A2, yA, nR), sC = fma2(C2, yC, nR);
                const f32x2 yRt = mul2(yR, rcp2(mul2(sA, sC)));
                const f32x2 uA = mul2(yRt, sC), uC = mul2(yRt, sA);
                PA[p] = fma2(pa, mul2(DA, uA), PA[p]);        // -qc [f(A) - f(R)]
                PC[p] = fma2(pc, mul2(DC, uC), PC[p]);        // -qs [f(C) - f(R)]
            }
            const f32x2 yA2 = mul2(yA, yA), yC2 = mul2(yC, yC);
            const f32x2 meA = fma2(A2, yA2, mone2), meC = fma2(C2, yC2, mone2);      // -(1 - |X|^2 y^2)
            if (POT) {
                const f32x2 zA = mul2(pa, yA2), zC = mul2(pc, yC2);
                // w_A + w_C with both seed corrections in one go; the -w_C d_j part is |d_j|/|A| times
                // smaller than the gross terms and takes the uncorrected z_C
                const f32x2 wS = fma2(fma2(zC, meC, mul2(zA, meA)), m15, add2(zA, zC));
                Fx[p] = fma2(wS, Ax, Fx[p]); Fy[p] = fma2(wS, Ay, Fy[p]); Fz[p] = fma2(wS, Az, Fz[p]);
                Hx[p] = fma2(zC, mx2, Hx[p]); Hy[p] = fma2(zC, my2, Hy[p]); Hz[p] = fma2(zC, mz2, Hz[p]);
            } else {
                // field only: w = y^3 q (1 + 1.5 e) with q and -1.5 q as broadcast column constants (one packed
                // instruction less per site than the route over q y)
                f32x2 q15A = bcast2(shQ[j].x), q15C = bcast2(shQ[j].y);
                if (MASKED) {
                    const bool e0 = (unsigned)(j - lo[2 * p]) < (unsigned)len[2 * p], e1 = (unsigned)(j - lo[2 * p + 1]) < (unsigned)len[2 * p + 1];
                    q15A = pack2(e0 ? 0.f : shQ[j].x, e1 ? 0.f : shQ[j].x); q15C = pack2(e0 ? 0.f : shQ[j].y, e1 ? 0.f : shQ[j].y);
                }
                const f32x2 cA = fma2(meA, q15A, qc2), cC = fma2(meC, q15C, qs2);
                const f32x2 wA = mul2(mul2(yA2, yA), cA), wC = mul2(mul2(yC2, yC), cC);
                const f32x2 wS = add2(wA, wC);
                Fx[p] = fma2(wS, Ax, Fx[p]); Fy[p] = fma2(wS, Ay, Fy[p]); Fz[p] = fma2(wS, Az, Fz[p]);
                Hx[p] = fma2(wC, mx2, Hx[p]); Hy[p] = fma2(wC, my2, Hy[p]); Hz[p] = fma2(wC, mz2, Hz[p]);
            }
        }
    }
}

template <int NP, bool POT, bool MASKED, int UNR>
__device__ __forceinline__ void diff_tile_N(const int4 *__restrict__ shX, const float2 *__restrict__ shQ, const int *sx, const int *sy, const int *sz,
                                            const int *lo, const int *len, const f32x2 *T2x, const f32x2 *T2y,
                                            const f32x2 *T2z, const f32x2 *ND2, f32x2 *Fx, f32x2 *Fy, f32x2 *Fz, f32x2 *PA) {
    const f32x2 mone2 = bcast2(-1.0f), m15 = bcast2(-1.5f);
#pragma unroll UNR
    for (int j = 0; j < kTileCols; ++j) {
        const int4 c = shX[j];
        const float qc = __int_as_float(c.w);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            int cx0 = c.x, cx1 = c.x;
            f32x2 qc2 = bcast2(qc);
            if (MASKED) {
                const bool e0 = (unsigned)(j - lo[2 * p]) < (unsigned)len[2 * p], e1 = (unsigned)(j - lo[2 * p + 1]) < (unsigned)len[2 * p + 1];
                cx0 = e0 ? kFarFix : c.x; cx1 = e1 ? kFarFix : c.x;
                qc2 = pack2(e0 ? 0.f : qc, e1 ? 0.f : qc);
            }
            const f32x2 Ax = pack2((float)(cx0 - sx[2 * p]), (float)(cx1 - sx[2 * p + 1]));
            const f32x2 Ay = pack2((float)(c.y - sy[2 * p]), (float)(c.y - sy[2 * p + 1]));
            const f32x2 Az = pack2((float)(c.z - sz[2 * p]), (float)(c.z - sz[2 * p + 1]));
            const f32x2 A2 = fma2(Az, Az, fma2(Ay, Ay, mul2(Ax, Ax)));
            const f32x2 yA = rsqrt2(A2);
            const f32x2 pa = mul2(qc2, yA);
            if (POT) {
                const f32x2 DA = fma2(Ax, T2x[p], fma2(Ay, T2y[p], fma2(Az, T2z[p], ND2[p])));
                const f32x2 R2 = sub2(A2, DA);
                const f32x2 yR = rsqrt2(R2);
                const f32x2 uA = mul2(yR, rcp2(fma2(A2, yA, mul2(R2, yR))));
                PA[p] = fma2(pa, mul2(DA, uA), PA[p]);
            }
            const f32x2 yA2 = mul2(yA, yA);
            const f32x2 meA = fma2(A2, yA2, mone2);
            f32x2 wA;
            if (POT) {
                const f32x2 zA = mul2(pa, yA2);
                wA = fma2(mul2(zA, meA), m15, zA);
            } else {
                f32x2 q15A = bcast2(shQ[j].x);
                if (MASKED) {
                    const bool e0 = (unsigned)(j - lo[2 * p]) < (unsigned)len[2 * p], e1 = (unsigned)(j - lo[2 * p + 1]) < (unsigned)len[2 * p + 1];
                    q15A = pack2(e0 ? 0.f : shQ[j].x, e1 ? 0.f : shQ[j].x);
                }
                wA = mul2(mul2(yA2, yA), fma2(meA, q15A, qc2));
            }
            Fx[p] = fma2(wA, Ax, Fx[p]); Fy[p] = fma2(wA, Ay, Fy[p]); Fz[p] = fma2(wA, Az, Fz[p]);
        }
    }
}

template <int ROWS, bool POT, int MINB, int UP, int UN>
__global__ void __launch_bounds__(kBlock, MINB) pair_diff_kernel(const PairArgs a) {
    static_assert(ROWS % 2 == 0, "rows are processed in fp32 pairs");
    constexpr int NP = ROWS / 2;
    __shared__ int4 shX[kTileCols];
    __shared__ float4 shD[kTileCols];
    __shared__ float shD2[kTileCols];
    __shared__ float2 shQ[POT ? 1 : kTileCols];     // field only: -1.5 q_core, -1.5 q_shell of the column
    if (a.gate != nullptr && (*a.gate & a.gate_skip_mask) != 0) return;
    const int4 *__restrict__ cols_x = static_cast<const int4 *>(a.cols);
    const float4 *__restrict__ cols_d = static_cast<const float4 *>(a.cols_d);
    const float *__restrict__ cols_d2 = a.cols_d2;

    DiffConsts kc;
    {
        const float ig = (float)a.inv_grid;
        kc.thr2 = (kDropNm * ig) * (kDropNm * ig);
        kc.big_f = kBigFieldNm / (ig * ig);
        kc.big_p = kBigPotNm / ig;
    }

    const int tid = threadIdx.x;
    const int row_base = blockIdx.x * (kBlock * ROWS) + tid;
    int sx[ROWS], sy[ROWS], sz[ROWS];
    float t2x[ROWS], t2y[ROWS], t2z[ROWS], nd2[ROWS];
    int loA[ROWS], lenA[ROWS], loB[ROWS], lenB[ROWS];
    double fx[ROWS], fy[ROWS], fz[ROWS], pA[POT ? ROWS : 1], pC[POT ? ROWS : 1];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        const int i = a.row0 + (lr < a.n_rows ? lr : 0);
        sx[r] = __double2int_rn((a.rx[i] - a.cx) * a.inv_grid);
        sy[r] = __double2int_rn((a.ry[i] - a.cy) * a.inv_grid);
        sz[r] = __double2int_rn((a.rz[i] - a.cz) * a.inv_grid);
        const float dx = (float)(a.dcomp[3 * i] * a.inv_grid), dy = (float)(a.dcomp[3 * i + 1] * a.inv_grid),
                    dz = (float)(a.dcomp[3 * i + 2] * a.inv_grid);
        t2x[r] = 2.0f * dx; t2y[r] = 2.0f * dy; t2z[r] = 2.0f * dz;
        nd2[r] = -fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        loA[r] = a.exA_lo[i]; lenA[r] = a.exA_len[i];
        loB[r] = a.exB_lo[i]; lenB[r] = a.exB_len[i];
        fx[r] = fy[r] = fz[r] = 0.0;
        if (POT) pA[POT ? r : 0] = pC[POT ? r : 0] = 0.0;
    }

    const int t0 = (int)(((int64_t)a.n_tiles * blockIdx.y) / a.n_split);
    const int t1 = (int)(((int64_t)a.n_tiles * (blockIdx.y + 1)) / a.n_split);

    for (int t = t0; t < t1; ++t) {
        const int base = t * kTileCols;
        const bool isP = t < a.n_tiles_a;
        __syncthreads();
        for (int j = tid; j < kTileCols; j += kBlock) {
            const int4 cx = cols_x[base + j];
            shX[j] = cx;
            float qsj = 0.f;
            if (isP) { const float4 cd = cols_d[base + j]; shD[j] = cd; shD2[j] = cols_d2[base + j]; qsj = cd.w; }
            if (!POT) shQ[POT ? 0 : j] = make_float2(-1.5f * __int_as_float(cx.w), -1.5f * qsj);
        }
        __syncthreads();

        bool need = false;
        int lo[ROWS], len[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            lo[r] = (isP ? loA[r] : loB[r]) - base;
            len[r] = isP ? lenA[r] : lenB[r];
            need |= (lo[r] < kTileCols) && (lo[r] + len[r] > 0);
        }
        // tile-local fp32 sums: field g, potentials ph (cores, shells)
        float gx[ROWS], gy[ROWS], gz[ROWS], phc[ROWS], phs[ROWS];
        bool bad = false;
        {
            const f32x2 zero2 = bcast2(0.f);
            f32x2 Fx[NP], Fy[NP], Fz[NP], Hx[NP], Hy[NP], Hz[NP], PA[NP], PC[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) Fx[p] = Fy[p] = Fz[p] = Hx[p] = Hy[p] = Hz[p] = PA[p] = PC[p] = zero2;
            // the row constants are kept as scalars (the pair-by-pair redo reads them) and paired up per tile
            f32x2 T2x[NP], T2y[NP], T2z[NP], ND2[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                T2x[p] = pack2(t2x[2 * p], t2x[2 * p + 1]); T2y[p] = pack2(t2y[2 * p], t2y[2 * p + 1]);
                T2z[p] = pack2(t2z[2 * p], t2z[2 * p + 1]); ND2[p] = pack2(nd2[2 * p], nd2[2 * p + 1]);
            }
            const bool masked = __any_sync(0xffffffffu, need);
            if (isP) {
                if (masked) diff_tile_P<NP, POT, true, UP>(shX, shD, shD2, shQ, sx, sy, sz, lo, len, T2x, T2y, T2z, ND2, Fx, Fy, Fz, Hx, Hy, Hz, PA, PC);
                else diff_tile_P<NP, POT, false, UP>(shX, shD, shD2, shQ, sx, sy, sz, lo, len, T2x, T2y, T2z, ND2, Fx, Fy, Fz, Hx, Hy, Hz, PA, PC);
            } else {
                if (masked) diff_tile_N<NP, POT, true, UN>(shX, shQ, sx, sy, sz, lo, len, T2x, T2y, T2z, ND2, Fx, Fy, Fz, PA);
                else diff_tile_N<NP, POT, false, UN>(shX, shQ, sx, sy, sz, lo, len, T2x, T2y, T2z, ND2, Fx, Fy, Fz, PA);
            }
            const f32x2 half2 = bcast2(0.5f);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                unpack2(fma2(half2, Hx[p], Fx[p]), gx[2 * p], gx[2 * p + 1]);
                unpack2(fma2(half2, Hy[p], Fy[p]), gy[2 * p], gy[2 * p + 1]);
                unpack2(fma2(half2, Hz[p], Fz[p]), gz[2 * p], gz[2 * p + 1]);
                unpack2(PA[p], phc[2 * p], phc[2 * p + 1]);
                unpack2(PC[p], phs[2 * p], phs[2 * p + 1]);
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                // a vanished distance shows up as inf/NaN (rsqrt(0) = inf, inf * 0 = NaN) or as an absurd sum
                bad |= !(fabsf(gx[r]) + fabsf(gy[r]) + fabsf(gz[r]) < kc.big_f);
                if (POT) bad |= !(fabsf(phc[r]) + fabsf(phs[r]) < kc.big_p);
            }
        }
        if (__any_sync(0xffffffffu, bad)) {
            // some distance in this tile vanished: redo it pair by pair with the reference's rule (term dropped)
#pragma unroll
            for (int r = 0; r < ROWS; ++r) gx[r] = gy[r] = gz[r] = phc[r] = phs[r] = 0.f;
#pragma unroll 1
            for (int j = 0; j < kTileCols; ++j) {
                const int4 c = shX[j];
                const float4 cd = isP ? shD[j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const bool excl = (unsigned)(j - lo[r]) < (unsigned)len[r];
                    float ph[2] = {phc[r], phs[r]}, g[3] = {gx[r], gy[r], gz[r]};
                    if (isP) diff_site_careful<POT, true>(c, cd, sx[r], sy[r], sz[r], t2x[r], t2y[r], t2z[r], nd2[r], excl, kc, ph, g);
                    else diff_site_careful<POT, false>(c, cd, sx[r], sy[r], sz[r], t2x[r], t2y[r], t2z[r], nd2[r], excl, kc, ph, g);
                    phc[r] = ph[0]; phs[r] = ph[1]; gx[r] = g[0]; gy[r] = g[1]; gz[r] = g[2];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            fx[r] += (double)gx[r]; fy[r] += (double)gy[r]; fz[r] += (double)gz[r];
            if (POT) { pA[POT ? r : 0] += (double)phc[r]; pC[POT ? r : 0] += (double)phs[r]; }
        }
    }

    // grid units -> nm: potentials * inv_grid, fields * inv_grid^2.  The potential slots receive the
    // DIFFERENCES sum q [f(X) - f(R)] (static part already removed), so the finalize kernel is told that
    // there is no separate static potential to subtract.
    const double s1 = -a.inv_grid, s2 = a.inv_grid * a.inv_grid;
    double *out = a.part + (int64_t)blockIdx.y * kPartComps * a.part_stride;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        if (lr < a.n_rows) {
            if (POT) {
                out[0 * a.part_stride + lr] = pA[POT ? r : 0] * s1;
                out[1 * a.part_stride + lr] = pC[POT ? r : 0] * s1;
            }
            out[2 * a.part_stride + lr] = fx[r] * s2;
            out[3 * a.part_stride + lr] = fy[r] * s2;
            out[4 * a.part_stride + lr] = fz[r] * s2;
        }
    }
}

// Launch shapes measured on the 100 k-Drude box (profiles/r02_diff_variants.txt, r02_variants_switch.txt): with
// the energy, 2 rows per thread and 4 resident blocks (124-128 registers) 23.0 ms against 24.8 ms for 4 rows (246
// registers, 2 blocks); field only, every shape between 2 rows / 6 blocks (80 registers) and 4 rows / 2 blocks
// (172 registers) lands within 13.8-14.5 ms - occupancy is not the limiter.  ncu (profiles/r02_ncu_full_pairs.txt):
// FMA pipe 65 % busy, XU (MUFU) 44 % / 69 % with the energy, ALU (IADD3, I2FP) 36 %, issue slots 62 %, top
// stalls math-pipe throttle, fixed-latency wait and dispatch stall.  What the loop saturates is the register
// file's read bandwidth (~2 operand words per lane and cycle, scripts/ubench2.cu: an FFMA2 with three 64-bit
// register operands takes 3.1 cycles, with two 2.1): ~114 operand words per (row pair, polarisable site)
// over all pipes, i.e. ~57 cycles at best, against ~68 measured.  Issuing FFMA2s with three register pairs as two
// scalar FFMAs each was measured slower (25.0 ms / 15.0 ms): it adds issue slots without removing operand reads.

// UPOL_DIFF_VARIANT (sweeps only): other launch shapes / unroll factors of the same kernel
static int diff_variant() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("UPOL_DIFF_VARIANT"); v = e ? atoi(e) : 0; }
    return v;
}

int diff_rows_per_block(bool with_pot) {
    const int v = diff_variant();
    if (with_pot) return kBlock * (v == 5 ? 4 : 2);
    return kBlock * ((v == 0 || v == 3) ? 2 : 4);
}

template <int ROWS, bool POT, int MINB, int UP, int UN>
static void launch_diff(const PairArgs &a, cudaStream_t st) {
    dim3 grid((unsigned)((a.n_rows + kBlock * ROWS - 1) / (kBlock * ROWS)), (unsigned)a.n_split);
    pair_diff_kernel<ROWS, POT, MINB, UP, UN><<<grid, kBlock, 0, st>>>(a);
}

int launch_pair_diff(const PairArgs &a, bool with_pot, cudaStream_t st) {
    if (a.n_rows <= 0) return UPOL_OK;
    const int v = diff_variant();
    if (with_pot) {
        switch (v) {
            case 1: launch_diff<2, true, 4, 4, 4>(a, st); break;
            case 2: launch_diff<2, true, 4, 1, 2>(a, st); break;
            case 3: launch_diff<2, true, 4, 2, 4>(a, st); break;
            case 4: launch_diff<2, true, 5, 2, 4>(a, st); break;
            case 5: launch_diff<4, true, 3, 1, 2>(a, st); break;
            default: launch_diff<2, true, 3, 2, 4>(a, st); break;
        }
    } else {
        switch (v) {
            case 1: launch_diff<4, false, 1, 2, 4>(a, st); break;
            case 2: launch_diff<4, false, 3, 2, 4>(a, st); break;
            case 3: launch_diff<2, false, 6, 2, 4>(a, st); break;
            case 4: launch_diff<4, false, 3, 1, 2>(a, st); break;
            default: launch_diff<2, false, 4, 2, 4>(a, st); break;
        }
    }
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

}  // namespace upol
