// K1/K3: all-pairs potential + field of point charges at the Drude-shell rows.
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
// shells) streamed through shared memory in tiles of kTileCols.  The d-independent part
//   E0 = -K sum_i qs_i sum_{j not in mol(i)} (qc_j + qs_j/2)/|r_j - r_i|
// is the same kernel without the field (K3).
//
// Layout: columns are double4/float4 (x, y, z, q), section A (cores) and section B (shells)
// each padded to a tile multiple with zero-charge dummies, so a tile is never mixed and the
// 1/2 weight of shell columns is applied per tile.  Sites of a molecule are contiguous, so
// the same-molecule exclusion of a row is one index interval per section: only tiles that
// overlap an interval of some row of the warp take the predicated path.
// Grid: x = row blocks (kBlock threads * ROWS rows), y = column splits; every (row block,
// split) writes its partial sums once - no atomics, deterministic.
//
// Zero-distance pairs contribute nothing (polarization.py:33-42).
#include "common.cuh"

namespace upol {

__device__ __forceinline__ double rsqrt_f64(double x) {
    // MUFU.RSQ64H seed (rel. error <= 2^-22) + one third-order (Halley) step:
    // 1/sqrt(x) = y (1-e)^(-1/2),  e = 1 - x y^2  ->  y (1 + e/2 + 3e^2/8), error O(e^3) ~ 1e-20
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = y * y;
    double e = fma(-x, t, 1.0);
    double p = fma(0.375, e, 0.5);
    double q = y * e;
    return fma(q, p, y);
}

template <int ROWS, bool POT, bool GRAD>
__global__ void __launch_bounds__(kBlock) pair_fp64_kernel(const PairArgs a) {
    __shared__ double4 sh[kTileCols];
    if (a.gate != nullptr && (*a.gate & a.gate_skip_mask) != 0) return;   // minimiser: converged / other phase
    const double4 *__restrict__ cols = static_cast<const double4 *>(a.cols);

    const int tid = threadIdx.x;
    const int row_base = blockIdx.x * (kBlock * ROWS) + tid;
    double sx[ROWS], sy[ROWS], sz[ROWS];
    int loA[ROWS], lenA[ROWS], loB[ROWS], lenB[ROWS];
    double phA[ROWS], phB[ROWS], fx[ROWS], fy[ROWS], fz[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        const bool live = lr < a.n_rows;
        const int i = a.row0 + (live ? lr : 0);
        sx[r] = a.rx[i]; sy[r] = a.ry[i]; sz[r] = a.rz[i];
        loA[r] = a.exA_lo[i]; lenA[r] = a.exA_len[i];
        loB[r] = a.exB_lo[i]; lenB[r] = a.exB_len[i];
        phA[r] = phB[r] = fx[r] = fy[r] = fz[r] = 0.0;
    }

    const int t0 = (int)(((int64_t)a.n_tiles * blockIdx.y) / a.n_split);
    const int t1 = (int)(((int64_t)a.n_tiles * (blockIdx.y + 1)) / a.n_split);

    for (int t = t0; t < t1; ++t) {
        const int base = t * kTileCols;
        __syncthreads();
        for (int j = tid; j < kTileCols; j += kBlock) sh[j] = cols[base + j];
        __syncthreads();

        const bool isB = t >= a.n_tiles_a;
        bool need = false;
        int lo[ROWS], len[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            lo[r] = (isB ? loB[r] : loA[r]) - base;      // interval relative to the tile
            len[r] = isB ? lenB[r] : lenA[r];
            need |= (lo[r] < kTileCols) && (lo[r] + len[r] > 0);
        }
        double ph[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) ph[r] = 0.0;

        if (__any_sync(0xffffffffu, need)) {
            // predicated tile: same-molecule columns masked out
#pragma unroll 2
            for (int j = 0; j < kTileCols; ++j) {
                const double4 c = sh[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const double dx = c.x - sx[r], dy = c.y - sy[r], dz = c.z - sz[r];
                    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const bool drop = ((unsigned)(j - lo[r]) < (unsigned)len[r]) || (__double2hiint(r2) == 0);
                    const double q = drop ? 0.0 : c.w;
                    r2 = drop ? 1.0 : r2;
                    const double inv = rsqrt_f64(r2);
                    const double tq = q * inv;
                    if (POT) ph[r] += tq;
                    if (GRAD) {
                        const double w = tq * (inv * inv);
                        fx[r] = fma(w, dx, fx[r]); fy[r] = fma(w, dy, fy[r]); fz[r] = fma(w, dz, fz[r]);
                    }
                }
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < kTileCols; ++j) {
                const double4 c = sh[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const double dx = c.x - sx[r], dy = c.y - sy[r], dz = c.z - sz[r];
                    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const bool drop = (__double2hiint(r2) == 0);   // zero distance: term dropped
                    const double q = drop ? 0.0 : c.w;
                    r2 = drop ? 1.0 : r2;
                    const double inv = rsqrt_f64(r2);
                    const double tq = q * inv;
                    if (POT) ph[r] += tq;
                    if (GRAD) {
                        const double w = tq * (inv * inv);
                        fx[r] = fma(w, dx, fx[r]); fy[r] = fma(w, dy, fy[r]); fz[r] = fma(w, dz, fz[r]);
                    }
                }
            }
        }
        if (POT) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                if (isB) phB[r] += ph[r]; else phA[r] += ph[r];
            }
        }
    }

    double *out = a.part + (int64_t)blockIdx.y * kPartComps * a.part_stride;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        if (lr < a.n_rows) {
            if (POT) {
                out[0 * a.part_stride + lr] = phA[r];
                out[1 * a.part_stride + lr] = phB[r];
            }
            if (GRAD) {
                out[2 * a.part_stride + lr] = fx[r];
                out[3 * a.part_stride + lr] = fy[r];
                out[4 * a.part_stride + lr] = fz[r];
            }
        }
    }
}

// Mixed mode (field only): fp32 pair arithmetic on coordinates relative to the box centre,
// per-tile fp32 partial sums promoted to fp64 accumulators after every tile (kTileCols terms).
// The potential is NOT formed in fp32: U_ind is the small difference of two large sums
// (U_coul - U_coul_static), and box-scale fp32 coordinates (ulp ~5e-7 nm) cannot deliver it to
// 1e-5; see DESIGN.md "mixed mode".  Energies always come from the fp64 kernel.
template <int ROWS>
__global__ void __launch_bounds__(kBlock) pair_mixed_kernel(const PairArgs a) {
    __shared__ float4 sh[kTileCols];
    if (a.gate != nullptr && (*a.gate & a.gate_skip_mask) != 0) return;
    const float4 *__restrict__ cols = static_cast<const float4 *>(a.cols);

    const int tid = threadIdx.x;
    const int row_base = blockIdx.x * (kBlock * ROWS) + tid;
    float sx[ROWS], sy[ROWS], sz[ROWS];
    int loA[ROWS], lenA[ROWS], loB[ROWS], lenB[ROWS];
    double fx[ROWS], fy[ROWS], fz[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        const bool live = lr < a.n_rows;
        const int i = a.row0 + (live ? lr : 0);
        sx[r] = (float)(a.rx[i] - a.cx); sy[r] = (float)(a.ry[i] - a.cy); sz[r] = (float)(a.rz[i] - a.cz);
        loA[r] = a.exA_lo[i]; lenA[r] = a.exA_len[i];
        loB[r] = a.exB_lo[i]; lenB[r] = a.exB_len[i];
        fx[r] = fy[r] = fz[r] = 0.0;
    }

    const int t0 = (int)(((int64_t)a.n_tiles * blockIdx.y) / a.n_split);
    const int t1 = (int)(((int64_t)a.n_tiles * (blockIdx.y + 1)) / a.n_split);

    for (int t = t0; t < t1; ++t) {
        const int base = t * kTileCols;
        __syncthreads();
        for (int j = tid; j < kTileCols; j += kBlock) sh[j] = cols[base + j];
        __syncthreads();

        const bool isB = t >= a.n_tiles_a;
        bool need = false;
        int lo[ROWS], len[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            lo[r] = (isB ? loB[r] : loA[r]) - base;
            len[r] = isB ? lenB[r] : lenA[r];
            need |= (lo[r] < kTileCols) && (lo[r] + len[r] > 0);
        }
        float gx[ROWS], gy[ROWS], gz[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) gx[r] = gy[r] = gz[r] = 0.f;

        if (__any_sync(0xffffffffu, need)) {
#pragma unroll 2
            for (int j = 0; j < kTileCols; ++j) {
                const float4 c = sh[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const float dx = c.x - sx[r], dy = c.y - sy[r], dz = c.z - sz[r];
                    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const bool drop = ((unsigned)(j - lo[r]) < (unsigned)len[r]) || (r2 == 0.f);
                    const float q = drop ? 0.f : c.w;
                    r2 = drop ? 1.f : r2;
                    const float inv = rsqrtf(r2);
                    const float w = (q * inv) * (inv * inv);
                    gx[r] = fmaf(w, dx, gx[r]); gy[r] = fmaf(w, dy, gy[r]); gz[r] = fmaf(w, dz, gz[r]);
                }
            }
        } else {
#pragma unroll 8
            for (int j = 0; j < kTileCols; ++j) {
                const float4 c = sh[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const float dx = c.x - sx[r], dy = c.y - sy[r], dz = c.z - sz[r];
                    // different molecules never coincide in a physical system; the clamp only
                    // keeps the arithmetic finite (1e-10 nm is far below fp32 resolution here)
                    const float r2 = fmaxf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)), 1e-20f);
                    const float inv = rsqrtf(r2);
                    const float w = (c.w * inv) * (inv * inv);
                    gx[r] = fmaf(w, dx, gx[r]); gy[r] = fmaf(w, dy, gy[r]); gz[r] = fmaf(w, dz, gz[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) { fx[r] += (double)gx[r]; fy[r] += (double)gy[r]; fz[r] += (double)gz[r]; }
    }

    double *out = a.part + (int64_t)blockIdx.y * kPartComps * a.part_stride;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int lr = row_base + r * kBlock;
        if (lr < a.n_rows) {
            out[2 * a.part_stride + lr] = fx[r];
            out[3 * a.part_stride + lr] = fy[r];
            out[4 * a.part_stride + lr] = fz[r];
        }
    }
}

int launch_pair_fp64(const PairArgs &a, bool with_pot, bool with_grad, cudaStream_t st) {
    if (a.n_rows <= 0) return UPOL_OK;
    dim3 grid((unsigned)((a.n_rows + kRowTile - 1) / kRowTile), (unsigned)a.n_split);
    if (with_pot && with_grad)
        pair_fp64_kernel<kRowsPerThread, true, true><<<grid, kBlock, 0, st>>>(a);
    else if (with_grad)
        pair_fp64_kernel<kRowsPerThread, false, true><<<grid, kBlock, 0, st>>>(a);
    else
        pair_fp64_kernel<kRowsPerThread, true, false><<<grid, kBlock, 0, st>>>(a);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

int launch_pair_mixed(const PairArgs &a, cudaStream_t st) {
    if (a.n_rows <= 0) return UPOL_OK;
    const int rows_per_block = kBlock * kRowsMixed;
    dim3 grid((unsigned)((a.n_rows + rows_per_block - 1) / rows_per_block), (unsigned)a.n_split);
    pair_mixed_kernel<kRowsMixed><<<grid, kBlock, 0, st>>>(a);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

}  // namespace upol
