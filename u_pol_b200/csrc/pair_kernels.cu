// K1/K3 and friends: all-pairs potential + field of point charges at the Drude-shell rows.
//
// Replaces the inter-molecular part of Ucoul and the static subtraction
// (U_pol/polarization.py:63-91, :50-61) and, through the field, the reverse-mode gradient
// jax.grad(Uind, 1) that jaxopt.BFGS takes (polarization.py:172-179).
//
// Formulation: with shells at s_i = r_i - d_i,
//   U_inter(d) = K sum_{i in Drude} qs_i [ sum_{j not in mol(i)} qc_j/|r_j - s_i|
//                                        + 1/2 sum_{j in Drude, not in mol(i)} qs_j/|s_j - s_i| ]
//   dU/dd_i    = -K qs_i sum_j q_j (x_j - s_i)/|x_j - s_i|^3   (+ intra + spring, finalize kernel)
// i.e. rows = Drude shells held in registers, columns = every point charge (cores, then
// shells) streamed through shared memory in tiles of kTileCols.
//
// Kernels in this file
//   pair_fp64_kernel<ROWS, POT, GRAD, MINB, SPLITF>  fp64 potential and/or field (K1); SPLITF keeps the
//                                                     fields of the two column sections apart (core forces)
//                                                     FUSE adds the static potential at the Drude cores (K3)
//                                                     from the same core tiles, 10 DP ops per interaction
// (the mixed-precision kernel lives in pair_diff.cu)
//
// Layout: columns are double4 (x, y, z, q), section A (cores) and section B (shells) each padded to
// a tile multiple with zero-charge dummies, so a tile is never mixed and the 1/2 weight of shell
// columns is applied per tile.  Sites of a molecule are contiguous, so the same-molecule exclusion
// of a row is one index interval per section: only tiles that overlap an interval of some row of
// the warp take the predicated loop; every other tile runs a speculative compare-free loop that
// is redone if it met a zero distance (zero-distance pairs contribute nothing,
// polarization.py:33-42).  Grid: x = row blocks (kBlock threads * ROWS rows), y = column splits;
// every (row block, split) writes its partial sums once - no atomics, deterministic.
#include <cstdlib>

#include "common.cuh"
#include "thole.cuh"

namespace upol {

// Compensated accumulation (Knuth two-sum) of a tile partial into a (hi, lo) pair.  The row
// potentials reach ~2e4 e/nm at 100 k Drudes (all core charges positive, all shells negative) while
// U_ind is their ~1e-6 remainder; adding 128-column tile partials (~10) into such an accumulator
// costs 2e-12 per add in plain fp64, which over 2000 tiles and 1e5 rows is 1e-10 of U_ind.  Seven
// adds per TILE per row make that negligible.
__device__ __forceinline__ void two_sum_acc(double &hi, double &lo, double x) {
    const double s = hi + x;
    const double bp = s - hi;
    lo += (hi - (s - bp)) + (x - bp);
    hi = s;
}

// q / r^3 from r^2 without forming 1/r (field-only variant): seed y (MUFU.RSQ64H, |e| <= 2^-22), e = 1 - x y^2,
// y^3 (1-e)^(-3/2) = y^3 (1 + 3e/2 + O(e^2)); 6 DP ops.  The dropped 15 e^2 / 8 <= 1.1e-13 per term is far below
// the 1e-10 bar of a gradient (the potential, whose row sums cancel to 1e-6 of their size, keeps the third-order
// root below).
__device__ __forceinline__ double q_over_r3_f64(double q, double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double y2 = y * y;
    const double e = fma(-x, y2, 1.0);
    const double c = fma(1.5, e, 1.0);
    return ((q * y) * y2) * c;
}

// SPLITF (core-force pass): the fields of section A and section B are kept apart, because they are
// multiplied by different charges of the row afterwards (output comps 0-2 and 3-5, no potential).
//
// FUSE (with POT): the d-independent static potential  phi_static_i = sum_j (qc_j + qs_j/2) / |r_j - r_i|
// (polarization.py:50-61, :82) is formed in the same pass from the core tiles: R = r_j - r_i = A - d_i, so
// |R|^2 = |A|^2 - 2 A.d_i + |d_i|^2 costs one add and three FMAs on top of the A already in registers.  One
// launch and one column sweep instead of two; the partial goes to slot 5 and finalize_rows_kernel
// subtracts it row by row (and keeps it, so later evaluations of the same geometry skip it).
template <int ROWS, bool POT, bool GRAD, int MINB, bool SPLITF = false, bool FUSE = false>
__global__ void __launch_bounds__(kBlock, MINB) pair_fp64_kernel(const PairArgs a) {
    static_assert(!FUSE || POT, "the static potential rides on the potential pass");
    __shared__ double4 sh[kTileCols];
    __shared__ double shq[FUSE ? kTileCols : 1];
    if (a.gate != nullptr && (*a.gate & a.gate_skip_mask) != 0) return;   // minimiser: converged / other phase
    const double4 *__restrict__ cols = static_cast<const double4 *>(a.cols);

    const int tid = threadIdx.x;
    const int row_base = blockIdx.x * (kBlock * ROWS) + tid;
    double sx[ROWS], sy[ROWS], sz[ROWS];
    int loA[ROWS], lenA[ROWS], loB[ROWS], lenB[ROWS];
    double phA[ROWS], phB[ROWS], plA[ROWS], plB[ROWS], fx[ROWS], fy[ROWS], fz[ROWS];
    double m2x[FUSE ? ROWS : 1], m2y[FUSE ? ROWS : 1], m2z[FUSE ? ROWS : 1], dd[FUSE ? ROWS : 1];   // -2 d_i, |d_i|^2
    double psH[FUSE ? ROWS : 1], psL[FUSE ? ROWS : 1];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        const bool live = lr < a.n_rows;
        const int i = a.row0 + (live ? lr : 0);
        sx[r] = a.rx[i]; sy[r] = a.ry[i]; sz[r] = a.rz[i];
        loA[r] = a.exA_lo[i]; lenA[r] = a.exA_len[i];
        loB[r] = a.exB_lo[i]; lenB[r] = a.exB_len[i];
        phA[r] = phB[r] = plA[r] = plB[r] = fx[r] = fy[r] = fz[r] = 0.0;
        if (FUSE) {
            constexpr int q = FUSE ? 1 : 0;
            const double dx = a.dcomp[3 * i], dy = a.dcomp[3 * i + 1], dz = a.dcomp[3 * i + 2];
            m2x[q * r] = -2.0 * dx; m2y[q * r] = -2.0 * dy; m2z[q * r] = -2.0 * dz;
            dd[q * r] = fma(dz, dz, fma(dy, dy, dx * dx));
            psH[q * r] = psL[q * r] = 0.0;
        }
    }

    const int t0 = (int)(((int64_t)a.n_tiles * blockIdx.y) / a.n_split);
    const int t1 = (int)(((int64_t)a.n_tiles * (blockIdx.y + 1)) / a.n_split);
    const double4 *__restrict__ cols_b = static_cast<const double4 *>(a.cols_b);
    double gx[SPLITF ? ROWS : 1], gy[SPLITF ? ROWS : 1], gz[SPLITF ? ROWS : 1];   // section A field once B starts
    if (SPLITF) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) gx[r] = gy[r] = gz[r] = 0.0;
    }

    for (int t = t0; t < t1; ++t) {
        const int base = t * kTileCols;
        const bool isB = t >= a.n_tiles_a;
        const double4 *src = (cols_b != nullptr && isB) ? cols_b + (int64_t)(t - a.n_tiles_a) * kTileCols : cols + base;
        __syncthreads();
        for (int j = tid; j < kTileCols; j += kBlock) {
            sh[j] = src[j];
            if (FUSE && !isB) shq[j] = a.col_qeff[base + j];
        }
        __syncthreads();
        const bool stat = FUSE && !isB;            // this tile also feeds the static potential
        if (SPLITF && t == a.n_tiles_a) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) { gx[r] = fx[r]; gy[r] = fy[r]; gz[r] = fz[r]; fx[r] = fy[r] = fz[r] = 0.0; }
        }

        bool need = false;
        int lo[ROWS], len[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            lo[r] = (isB ? loB[r] : loA[r]) - base;      // interval relative to the tile
            len[r] = isB ? lenB[r] : lenA[r];
            need |= (lo[r] < kTileCols) && (lo[r] + len[r] > 0);
        }
        // tile-local sums: the compare-free loop below is speculative and may have to be redone
        double ph[ROWS], tx[ROWS], ty[ROWS], tz[ROWS], pst[FUSE ? ROWS : 1];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) { ph[r] = tx[r] = ty[r] = tz[r] = 0.0; if (FUSE) pst[FUSE ? r : 0] = 0.0; }

        bool careful = __any_sync(0xffffffffu, need);
        if (!careful && stat) {
            // core tile of the fused pass: the loop below plus |R|^2 from A and one more reciprocal root.  Unroll factors
            // measured on the 100 k-Drude box (1/2/4/8/16/32 here x 2/4/8 for the shell loop): 8 and 4 are best,
            // 44.55 ms against 45.57 ms for 2 and 4 - same arithmetic and order, different instruction schedule
            int hmin = 0x7fffffff;
#pragma unroll 8
            for (int j = 0; j < kTileCols; ++j) {
                const double4 c = sh[j];
                const double qe = shq[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const double dx = c.x - sx[r], dy = c.y - sy[r], dz = c.z - sz[r];
                    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const double R2 = fma(dx, m2x[FUSE ? r : 0], fma(dy, m2y[FUSE ? r : 0], fma(dz, m2z[FUSE ? r : 0], r2 + dd[FUSE ? r : 0])));
                    hmin = min(hmin, min(__double2hiint(r2), __double2hiint(R2)));
                    pst[FUSE ? r : 0] = fma(qe, rsqrt_f64(R2), pst[FUSE ? r : 0]);
                    const double inv = rsqrt_f64(r2);
                    const double tq = c.w * inv;
                    ph[r] += tq;
                    if (GRAD) {
                        const double w = tq * (inv * inv);
                        tx[r] = fma(w, dx, tx[r]); ty[r] = fma(w, dy, ty[r]); tz[r] = fma(w, dz, tz[r]);
                    }
                }
            }
            careful = __any_sync(0xffffffffu, hmin < kZeroR2Hi);
            if (careful) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) { ph[r] = tx[r] = ty[r] = tz[r] = 0.0; pst[FUSE ? r : 0] = 0.0; }
            }
        } else if (!careful) {
            // no same-molecule column in this tile for any row of the warp.  No per-pair selects either:
            // only an integer min of the high word of r^2 (ALU pipe); a zero-distance pair
            // (polarization.py:33-42: term dropped) is caught after the tile, which is then redone
            // with the predicated loop - the sums are tile-local, so nothing has to be undone.
            int hmin = 0x7fffffff;
#pragma unroll 4
            for (int j = 0; j < kTileCols; ++j) {
                const double4 c = sh[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const double dx = c.x - sx[r], dy = c.y - sy[r], dz = c.z - sz[r];
                    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    hmin = min(hmin, __double2hiint(r2));
                    if (POT) {
                        const double inv = rsqrt_f64(r2);
                        const double tq = c.w * inv;
                        ph[r] += tq;
                        if (GRAD) {
                            const double w = tq * (inv * inv);
                            tx[r] = fma(w, dx, tx[r]); ty[r] = fma(w, dy, ty[r]); tz[r] = fma(w, dz, tz[r]);
                        }
                    } else {
                        const double w = q_over_r3_f64(c.w, r2);
                        tx[r] = fma(w, dx, tx[r]); ty[r] = fma(w, dy, ty[r]); tz[r] = fma(w, dz, tz[r]);
                    }
                }
            }
            careful = __any_sync(0xffffffffu, hmin < kZeroR2Hi);       // r < 1e-5 nm counts as zero distance
            if (careful) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) ph[r] = tx[r] = ty[r] = tz[r] = 0.0;
            }
        }
        if (careful) {
            // predicated tile: same-molecule columns and zero-distance pairs masked out
#pragma unroll 2
            for (int j = 0; j < kTileCols; ++j) {
                const double4 c = sh[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const double dx = c.x - sx[r], dy = c.y - sy[r], dz = c.z - sz[r];
                    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const bool excl = (unsigned)(j - lo[r]) < (unsigned)len[r];
                    if (FUSE && stat) {
                        // f(R) is dropped on its own when ITS distance vanishes (polarization.py:33-42)
                        double R2 = fma(dx, m2x[FUSE ? r : 0], fma(dy, m2y[FUSE ? r : 0], fma(dz, m2z[FUSE ? r : 0], r2 + dd[FUSE ? r : 0])));
                        const bool dropR = excl || (__double2hiint(R2) < kZeroR2Hi);
                        R2 = dropR ? 1.0 : R2;
                        pst[FUSE ? r : 0] = fma(dropR ? 0.0 : shq[FUSE ? j : 0], rsqrt_f64(R2), pst[FUSE ? r : 0]);
                    }
                    const bool drop = excl || (__double2hiint(r2) < kZeroR2Hi);
                    const double q = drop ? 0.0 : c.w;
                    r2 = drop ? 1.0 : r2;
                    if (POT) {
                        const double inv = rsqrt_f64(r2);
                        const double tq = q * inv;
                        ph[r] += tq;
                        if (GRAD) {
                            const double w = tq * (inv * inv);
                            tx[r] = fma(w, dx, tx[r]); ty[r] = fma(w, dy, ty[r]); tz[r] = fma(w, dz, tz[r]);
                        }
                    } else {
                        const double w = q_over_r3_f64(q, r2);
                        tx[r] = fma(w, dx, tx[r]); ty[r] = fma(w, dy, ty[r]); tz[r] = fma(w, dz, tz[r]);
                    }
                }
            }
        }
        if (GRAD) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) { fx[r] += tx[r]; fy[r] += ty[r]; fz[r] += tz[r]; }
        }
        if (POT) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                if (isB) two_sum_acc(phB[r], plB[r], ph[r]); else two_sum_acc(phA[r], plA[r], ph[r]);
                if (FUSE && stat) two_sum_acc(psH[FUSE ? r : 0], psL[FUSE ? r : 0], pst[FUSE ? r : 0]);
            }
        }
    }

    double *out = a.part + (int64_t)blockIdx.y * kPartComps * a.part_stride;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        if (lr < a.n_rows) {
            if (SPLITF) {
                // the split of this block saw B tiles iff t1 > n_tiles_a; before that fx IS the A field
                const bool sawB = t1 > a.n_tiles_a;
                out[0 * a.part_stride + lr] = sawB ? gx[r] : fx[r];
                out[1 * a.part_stride + lr] = sawB ? gy[r] : fy[r];
                out[2 * a.part_stride + lr] = sawB ? gz[r] : fz[r];
                out[3 * a.part_stride + lr] = sawB ? fx[r] : 0.0;
                out[4 * a.part_stride + lr] = sawB ? fy[r] : 0.0;
                out[5 * a.part_stride + lr] = sawB ? fz[r] : 0.0;
                continue;
            }
            if (POT) {
                out[0 * a.part_stride + lr] = phA[r] + plA[r];
                out[1 * a.part_stride + lr] = phB[r] + plB[r];
            }
            if (FUSE) out[5 * a.part_stride + lr] = psH[FUSE ? r : 0] + psL[FUSE ? r : 0];
            if (GRAD) {
                out[2 * a.part_stride + lr] = fx[r];
                out[3 * a.part_stride + lr] = fy[r];
                out[4 * a.part_stride + lr] = fz[r];
            }
        }
    }
}

template <int ROWS, int MINB>
static void launch_fp64_variant(const PairArgs &a, bool with_pot, bool with_grad, bool fuse_static, cudaStream_t st) {
    dim3 grid((unsigned)((a.n_rows + kBlock * ROWS - 1) / (kBlock * ROWS)), (unsigned)a.n_split);
    if (with_pot && fuse_static) {
        if (with_grad) pair_fp64_kernel<ROWS, true, true, ROWS == 1 ? 4 : 3, false, true><<<grid, kBlock, 0, st>>>(a);   // 3 resident blocks: <= 168 registers
        else pair_fp64_kernel<ROWS, true, false, MINB, false, true><<<grid, kBlock, 0, st>>>(a);
    } else if (with_pot && with_grad)
        pair_fp64_kernel<ROWS, true, true, MINB><<<grid, kBlock, 0, st>>>(a);
    else if (with_grad)
        pair_fp64_kernel<ROWS, false, true, MINB><<<grid, kBlock, 0, st>>>(a);
    else
        pair_fp64_kernel<ROWS, true, false, MINB><<<grid, kBlock, 0, st>>>(a);
}

int fp64_rows_per_block() { return kRowTile; }

// Variants tried on the 100 k-Drude box (profiles/r01_variant_sweep.txt): 1/2/3/4 rows per thread,
// min-blocks 1/3/4/6/8 - all within 5 %; occupancy is not the limiter (the FP64 pipe is, ~80 %
// busy), so the plain 2-rows / no-min-blocks build stays.
int launch_pair_fp64(const PairArgs &a, bool with_pot, bool with_grad, bool fuse_static, cudaStream_t st) {
    if (a.n_rows <= 0) return UPOL_OK;
    // Small systems (BASELINE config 2: 3 000 Drude rows): with 2 rows per thread the grid is only ~3 waves of
    // resident blocks and the last, mostly empty wave costs 20 %; one row per thread doubles the block count.
    static const int force_rows = getenv("UPOL_F64_ROWS") ? atoi(getenv("UPOL_F64_ROWS")) : 0;
    const int64_t blocks2 = (int64_t)((a.n_rows + kRowTile - 1) / kRowTile) * a.n_split;
    const bool rows1 = force_rows ? force_rows == 1 : (blocks2 < 148 * 3 * 6 && with_pot && fuse_static);   // measured: only the fused pass gains
    if (rows1) launch_fp64_variant<1, 1>(a, with_pot, with_grad, fuse_static, st);
    else launch_fp64_variant<kRowsPerThread, 1>(a, with_pot, with_grad, fuse_static, st);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

int launch_pair_fp64_two_fields(const PairArgs &a, cudaStream_t st) {
    if (a.n_rows <= 0) return UPOL_OK;
    dim3 grid((unsigned)((a.n_rows + kRowTile - 1) / kRowTile), (unsigned)a.n_split);
    pair_fp64_kernel<kRowsPerThread, false, true, 1, true><<<grid, kBlock, 0, st>>>(a);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

}  // namespace upol
