// System object: HBM-resident geometry/topology, and the evaluation pipeline
//   prepare (K4a: shells s = r - d, shell columns) -> pair kernel (K1, static potential fused) -> finalize
//   (K4b/K2: fold the column splits, scale, spring, Thole intra; the last block sums the energy vector and, in
//   sharded runs, stores the gradient rows and energy partials into the peers).
// Replaces the array construction of U_pol/util.py:45-279 (compact O(N) form) and the body of
// Uind (polarization.py:111-151) + its gradient.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "thole.cuh"

namespace upol {

static thread_local char g_err[256] = "";
void set_last_cuda_error(cudaError_t e, const char *file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
}
const char *last_error_message() { return g_err; }

constexpr double kFar = 1.0e8;   // nm; position of zero-charge padding columns

// ---------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------

constexpr int kFixFar = (1 << 30) - 1;   // fixed-point position of zero-charge padding columns

__device__ __forceinline__ int4 fix_col(double x, double y, double z, double inv_grid, double q) {
    return make_int4(__double2int_rn(x * inv_grid), __double2int_rn(y * inv_grid), __double2int_rn(z * inv_grid),
                     __float_as_int((float)q));
}

// Core columns (section A), static-pass columns and Drude core rows; once per geometry.
__global__ void build_core_columns_kernel(int nsite, int na_pad, const double *__restrict__ r,
                                          const double *__restrict__ qc, const double *__restrict__ qs,
                                          double4 *__restrict__ cols,
                                          double4 *__restrict__ cols_static, double *__restrict__ col_qeff) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na_pad) return;
    if (i < nsite) {
        const double x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
        cols[i] = make_double4(x, y, z, qc[i]);
        cols_static[i] = make_double4(x, y, z, qc[i] + 0.5 * qs[i]);
        col_qeff[i] = qc[i] + 0.5 * qs[i];
    } else {
        cols[i] = make_double4(kFar, kFar, kFar, 0.0);
        cols_static[i] = make_double4(kFar, kFar, kFar, 0.0);
        col_qeff[i] = 0.0;
    }
}

// Mixed mode, difference form: fixed-point core columns, section P (polarisable sites, Drude-row order)
// then section N (the others); once per geometry.
__global__ void build_diff_columns_kernel(int nsite, int nd, int nb_pad, int nn_pad, const double *__restrict__ r,
                                          const double *__restrict__ qc, const int *__restrict__ site_row,
                                          const int *__restrict__ site_nslot, double cx, double cy, double cz,
                                          double inv_grid, int4 *__restrict__ cols_x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_n = nsite - nd;
    if (i >= nd && i < nb_pad) cols_x[i] = make_int4(kFixFar, kFixFar, kFixFar, 0);                      // P padding
    if (i >= n_n && i < nn_pad) cols_x[nb_pad + i] = make_int4(kFixFar, kFixFar, kFixFar, 0);            // N padding
    if (i >= nsite) return;
    const int row = site_row[i];
    const int slot = row >= 0 ? row : nb_pad + site_nslot[i];
    cols_x[slot] = fix_col(r[3 * i] - cx, r[3 * i + 1] - cy, r[3 * i + 2] - cz, inv_grid, qc[i]);
}

__global__ void build_static_rows_kernel(int nd, const int *__restrict__ row_site, const double *__restrict__ r,
                                         double *__restrict__ sx, double *__restrict__ sy, double *__restrict__ sz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    const int s = row_site[i];
    sx[i] = r[3 * s]; sy[i] = r[3 * s + 1]; sz[i] = r[3 * s + 2];
}

// K4a: shell positions and shell columns (section B) from the current displacements.
// With d_full != nullptr the displacements come in the reference's (nsite,3) layout and the compact copy dc
// is written here as well (no separate gather launch).
__global__ void prepare_shells_kernel(int nd, int nb_pad, const int *__restrict__ row_site,
                                      const double *__restrict__ r, const double *__restrict__ d_full, double *__restrict__ dc,
                                      const double *__restrict__ row_qs, double inv_grid,
                                      double *__restrict__ rowx, double *__restrict__ rowy, double *__restrict__ rowz,
                                      double4 *__restrict__ colsB, float4 *__restrict__ cols_d, float *__restrict__ cols_d2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb_pad) return;
    if (i < nd) {
        const int s = row_site[i];
        double ddx, ddy, ddz;
        if (d_full != nullptr) {
            ddx = d_full[3 * s]; ddy = d_full[3 * s + 1]; ddz = d_full[3 * s + 2];
            dc[3 * i] = ddx; dc[3 * i + 1] = ddy; dc[3 * i + 2] = ddz;
        } else {
            ddx = dc[3 * i]; ddy = dc[3 * i + 1]; ddz = dc[3 * i + 2];
        }
        const double x = r[3 * s] - ddx, y = r[3 * s + 1] - ddy, z = r[3 * s + 2] - ddz;
        rowx[i] = x; rowy[i] = y; rowz[i] = z;
        colsB[i] = make_double4(x, y, z, row_qs[i]);
        // difference-form mixed kernel: -2 d_j and |d_j|^2 in grid units, from the SAME fp32 values
        const float dx = (float)(ddx * inv_grid), dy = (float)(ddy * inv_grid), dz = (float)(ddz * inv_grid);
        cols_d[i] = make_float4(-2.0f * dx, -2.0f * dy, -2.0f * dz, (float)row_qs[i]);
        cols_d2[i] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    } else {
        colsB[i] = make_double4(kFar, kFar, kFar, 0.0);
        cols_d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        cols_d2[i] = 0.f;
    }
}

template <int BS>
__device__ __forceinline__ double block_sum(double v, double *sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        for (int k = 0; k < BS / 32; ++k) t += sm[k];
    }
    return t;   // valid in thread 0
}

constexpr int kFinBlock = 128;
constexpr int kEblockComps = 5;   // per finalize block: inter, intra, self, |g|^2, static (= energy[INTER..STATIC])
constexpr int kFinRows = 32;      // rows per finalize block: 32 rows x 4 split lanes
constexpr int kFinLanes = kFinBlock / kFinRows;

// Fold the column-split partials of one row.  A block handles FR rows; thread (lx, ly) sums the
// splits ly, ly+FL, ... (FL = 128 / FR lanes) of row lx for components [c0, c0+NC) with independent loads in
// flight, and the lane sums are combined in fixed order by lane 0 (deterministic).  At small N the split
// count reaches ~100 and a single thread per row would chain ~500 L2 loads; small systems therefore use
// FR = 8 (16 lanes per row, four times as many blocks), large ones FR = 32.
template <int NC, int FR = kFinRows>
__device__ __forceinline__ void fold_splits(const double *__restrict__ part, int64_t stride, int lr, bool live,
                                            int n_split, int c0, double (*sm)[FR][kPartComps], double *out) {
    constexpr int FL = kFinBlock / FR;
    const int lx = threadIdx.x % FR, ly = threadIdx.x / FR;
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    if (live) {
#pragma unroll 4
        for (int s = ly; s < n_split; s += FL) {
            const double *p = part + (int64_t)s * kPartComps * stride + (int64_t)c0 * stride + lr;
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] += p[(int64_t)c * stride];
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NC; ++c) sm[ly][lx][c] = acc[c];
    __syncthreads();
    if (ly == 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double t = sm[0][lx][c];
#pragma unroll
            for (int l = 1; l < FL; ++l) t += sm[l][lx][c];
            out[c] = t;
        }
    }
}

// energy vector from the summed partials (inter, intra, self, |g|^2, static); flags: 1 = report the static part
// (informational, already inside E_INTER), 2 = form the total (single rank), 4 = keep this rank's static sum
constexpr int kEPutStatic = 1, kETotal = 2, kESaveStatic = 4;
__device__ __forceinline__ void write_energy_vector(double *__restrict__ e, const double *res, double *__restrict__ e_static, int flags) {
    if (flags & kESaveStatic) e_static[0] = res[4];
    e[UPOL_E_INTER] = res[0]; e[UPOL_E_INTRA] = res[1]; e[UPOL_E_SELF] = res[2]; e[UPOL_E_GNORM2] = res[3];
    e[UPOL_E_STATIC] = (flags & kEPutStatic) ? e_static[0] : res[4];
    e[6] = 0.0; e[7] = 0.0;
    e[UPOL_E_UIND] = (flags & kETotal) ? (res[0] + res[1]) + res[2] : 0.0;
}

// K4b + K2: per Drude row, fold the column-split partials, apply K*qs, add the spring and the
// intra-molecular Thole pairs, write the gradient row, and emit per-block energy partials.
template <int FR>
__global__ void __launch_bounds__(kFinBlock)
finalize_rows_kernel(int n_rows, int row0, int n_split_pot, int n_split_f64, int n_split_mixed, int fused_static,
                     const int *__restrict__ gate, int phase_switch, const double *__restrict__ part, int64_t stride,
                     const int *__restrict__ row_site, const double *__restrict__ r,
                     const double *__restrict__ dc, const double *__restrict__ row_qs,
                     const double *__restrict__ row_k,
                     const int *__restrict__ th_ptr, const int *__restrict__ th_row,
                     const double *__restrict__ th_u, double *__restrict__ phi_static,
                     bool with_grad, double *__restrict__ gc, double *__restrict__ eblock,
                     int *__restrict__ ticket, double *__restrict__ e_out, double *__restrict__ e_static, int e_flags,
                     const EvalXchg xg, double *__restrict__ g_full, int nd_total, int nsite) {
    __shared__ double sm[kFinBlock / 32];
    __shared__ int is_last;
    __shared__ double fold[kFinBlock / FR][FR][kPartComps];
    const int gw = gate ? *gate : 0;
    if (gw & kGateDone) return;
    // which field kernel ran decides how many column splits hold force partials
    const int n_split_f = (phase_switch && (gw & kGatePhaseMixed)) ? n_split_mixed : n_split_f64;
    const int lr = blockIdx.x * FR + threadIdx.x % FR;
    const bool live = lr < n_rows;
    const bool owner = live && threadIdx.x < FR;
    double pot[2] = {0.0, 0.0}, frc[3] = {0.0, 0.0, 0.0}, stat[1] = {0.0};
    if (n_split_pot > 0) fold_splits<2, FR>(part, stride, lr, live, n_split_pot, 0, fold, pot);
    if (with_grad) fold_splits<3, FR>(part, stride, lr, live, n_split_f, 2, fold, frc);
    // fused pass: the static potential of the row arrives as slot 5 of the same launch; it is kept
    // (phi_static) so that later evaluations of this geometry need not form it again
    if (fused_static) fold_splits<1, FR>(part, stride, lr, live, n_split_pot, 5, fold, stat);
    double e_inter = 0.0, e_intra = 0.0, e_self = 0.0, g2 = 0.0, e_stat = 0.0;
    if (owner) {
        const int i = row0 + lr;
        const double pa = pot[0], pb = pot[1], fx = frc[0], fy = frc[1], fz = frc[2];
        const double qs = row_qs[i], kk = row_k[i];
        const double dx = dc[3 * i], dy = dc[3 * i + 1], dz = dc[3 * i + 2];
        // (U_coul - U_coul_static) share of the row: static potential removed before the row sum
        const double ps = fused_static ? stat[0] : (phi_static ? phi_static[i] : 0.0);
        if (fused_static) { phi_static[i] = ps; e_stat = -kCoulomb * qs * ps; }
        e_inter = kCoulomb * qs * ((pa - ps) + 0.5 * pb);
        e_self = 0.5 * kk * (dx * dx + dy * dy + dz * dz);
        double gx = -kCoulomb * qs * fx + kk * dx;
        double gy = -kCoulomb * qs * fy + kk * dy;
        double gz = -kCoulomb * qs * fz + kk * dz;
        const int sa = row_site[i];
        const double rax = r[3 * sa], ray = r[3 * sa + 1], raz = r[3 * sa + 2];
        for (int p = th_ptr[i]; p < th_ptr[i + 1]; ++p) {
            const int b = th_row[p];
            const double u = th_u[p];
            const int sb = row_site[b];
            const double Rx = r[3 * sb] - rax, Ry = r[3 * sb + 1] - ray, Rz = r[3 * sb + 2] - raz;
            const double bx = dc[3 * b], by = dc[3 * b + 1], bz = dc[3 * b + 2];
            const double Ax = Rx + dx, Ay = Ry + dy, Az = Rz + dz;          // polarization.py:71
            const double Bx = Rx - bx, By = Ry - by, Bz = Rz - bz;          // :72
            const double Cx = Ax - bx, Cy = Ay - by, Cz = Az - bz;          // :73
            double gR, gA, gB, gC, cR, cA, cB, cC;
            thole_g(Rx, Ry, Rz, u, gR, cR);
            thole_g(Ax, Ay, Az, u, gA, cA);
            thole_g(Bx, By, Bz, u, gB, cB);
            thole_g(Cx, Cy, Cz, u, gC, cC);
            const double qq = kCoulomb * qs * row_qs[b];
            e_intra += 0.5 * qq * (gR - gA - gB + gC);                      // :94-99, :107
            gx += qq * (cC * Cx - cA * Ax);
            gy += qq * (cC * Cy - cA * Ay);
            gz += qq * (cC * Cz - cA * Az);
        }
        if (with_grad) {
            gc[3 * i] = gx; gc[3 * i + 1] = gy; gc[3 * i + 2] = gz;
            g2 = gx * gx + gy * gy + gz * gz;
            if (g_full != nullptr) {
                // the reference's (nsite,3) layout written here as well (no separate scatter launch): the row's own
                // site, and exact zeros on the non-polarisable sites up to the next Drude site (and before the first)
                const int s_next = (i + 1 < nd_total) ? row_site[i + 1] : nsite;
                g_full[3 * sa] = gx; g_full[3 * sa + 1] = gy; g_full[3 * sa + 2] = gz;
                for (int q = (i == 0 ? 0 : sa + 1); q < s_next; ++q) {
                    if (q == sa) continue;
                    g_full[3 * q] = 0.0; g_full[3 * q + 1] = 0.0; g_full[3 * q + 2] = 0.0;
                }
            }
            if (xg.on && xg.push_grad) {
                // fused all-gather: the row goes straight into every peer's gradient buffer (NVLink stores)
                for (int rk = 0; rk < xg.nranks; ++rk) {
                    if (rk == xg.rank) continue;
                    double *pg = xg.peer_g[rk] + xg.g_off + 3 * (int64_t)i;
                    pg[0] = gx; pg[1] = gy; pg[2] = gz;
                }
                __threadfence_system();
            }
        }
    }
    const double s0 = block_sum<kFinBlock>(e_inter, sm);
    const double s1 = block_sum<kFinBlock>(e_intra, sm);
    const double s2 = block_sum<kFinBlock>(e_self, sm);
    const double s3 = block_sum<kFinBlock>(g2, sm);
    const double s4 = fused_static ? block_sum<kFinBlock>(e_stat, sm) : 0.0;
    if (threadIdx.x == 0) {
        double *o = eblock + kEblockComps * (int64_t)blockIdx.x;
        o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3; o[4] = s4;
        if (xg.on) __threadfence_system(); else __threadfence();
        is_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    // The block that finishes last sums the per-block partials in FIXED order (deterministic, whichever
    // block that is) and writes the energy vector: no separate reduce / assemble launches, no memset.
    __threadfence();
    double res[kEblockComps];
#pragma unroll
    for (int c = 0; c < kEblockComps; ++c) {
        double v = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += kFinBlock) v += __ldcg(eblock + (int64_t)b * kEblockComps + c);
        res[c] = block_sum<kFinBlock>(v, sm);
    }
    if (threadIdx.x == 0) {
        *ticket = 0;
        if (!xg.on) {
            write_energy_vector(e_out, res, e_static, e_flags);
        } else {
            // the rank's partial sums go into its slot of every rank's mailbox; then, after a system-wide fence
            // (every block fenced its gradient stores before it took its ticket), the arrival epoch is raised
            if (e_flags & kESaveStatic) e_static[0] = res[4];
            if (e_flags & kEPutStatic) res[4] = e_static[0];
            for (int rk = 0; rk < xg.nranks; ++rk) {
                double *m = xg.peer_e[rk] + xg.e_off + xg.rank * UPOL_E_SIZE;
                for (int c = 0; c < kEblockComps; ++c) m[c] = res[c];
            }
            __threadfence_system();
            for (int rk = 0; rk < xg.nranks; ++rk) *((volatile int *)(xg.peer_flags[rk] + xg.rank)) = xg.epoch;
        }
    }
}

// A rank without rows still owes its peers an (all-zero) energy slot and its arrival epoch.
__global__ void eval_xchg_empty_kernel(const EvalXchg xg) {
    if (threadIdx.x || blockIdx.x) return;
    for (int rk = 0; rk < xg.nranks; ++rk) {
        double *m = xg.peer_e[rk] + xg.e_off + xg.rank * UPOL_E_SIZE;
        for (int c = 0; c < kEblockComps; ++c) m[c] = 0.0;
    }
    __threadfence_system();
    for (int rk = 0; rk < xg.nranks; ++rk) *((volatile int *)(xg.peer_flags[rk] + xg.rank)) = xg.epoch;
}

// Spin until flags[i] >= epoch (raised by peer i through NVLink).  The time-out (~30 s) only guards against a
// dead peer: ranks may enter a call seconds apart (host-side skew), which is not an error.
__device__ __forceinline__ bool wait_epoch_long(const int *flags, int i, int epoch) {
    const volatile int *f = flags + i;
    const long long t0 = clock64();
    while (*f < epoch) {
        if (clock64() - t0 > 60000000000LL) return false;
        __nanosleep(100);
    }
    __threadfence_system();
    return true;
}

// Consumer side of the exchange: every peer's rows and energy partials have arrived once all epochs are in;
// the energy vector is the sum of the mailbox slots in RANK ORDER (same bits on every rank).  A peer that never
// delivers turns the energy into NaN instead of hanging the stream.
__global__ void eval_xchg_wait_kernel(const EvalXchg xg, const int *__restrict__ my_flags, const double *__restrict__ my_e,
                                      double *__restrict__ e_out) {
    bool ok = true;
    if ((int)threadIdx.x < xg.nranks) ok = wait_epoch_long(my_flags, threadIdx.x, xg.epoch);
    ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x) return;
    double res[kEblockComps] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int rk = 0; rk < xg.nranks; ++rk)
        for (int c = 0; c < kEblockComps; ++c) res[c] += ((const volatile double *)my_e)[xg.e_off + rk * UPOL_E_SIZE + c];
    if (!ok) { for (int c = 0; c < kEblockComps; ++c) res[c] = nan(""); }
    if (e_out != nullptr) {
        e_out[UPOL_E_INTER] = res[0]; e_out[UPOL_E_INTRA] = res[1]; e_out[UPOL_E_SELF] = res[2]; e_out[UPOL_E_GNORM2] = res[3];
        e_out[UPOL_E_STATIC] = res[4]; e_out[6] = 0.0; e_out[7] = 0.0;
        e_out[UPOL_E_UIND] = (res[0] + res[1]) + res[2];
    }
}

// Static pass epilogue: E0 partial = -K sum_rows qs_i * phi_i.
__global__ void __launch_bounds__(kFinBlock)
finalize_static_kernel(int n_rows, int row0, int n_split, const double *__restrict__ part, int64_t stride,
                       const double *__restrict__ row_qs, double *__restrict__ phi_static,
                       double *__restrict__ eblock) {
    __shared__ double sm[kFinBlock / 32];
    __shared__ double fold[kFinLanes][kFinRows][kPartComps];
    const int lr = blockIdx.x * kFinRows + threadIdx.x % kFinRows;
    const bool live = lr < n_rows;
    double pot[1] = {0.0};
    fold_splits<1>(part, stride, lr, live, n_split, 0, fold, pot);
    double e = 0.0;
    if (live && threadIdx.x < kFinRows) {
        const double pa = pot[0];
        // kept PER ROW: the static potential (~1e4 e/nm at 100 k Drudes) is subtracted from the
        // row's core/shell potential before anything is summed over rows, otherwise U_ind
        // (~1e5 kJ/mol) would be the difference of two ~1e11 kJ/mol totals
        phi_static[row0 + lr] = pa;
        e = -kCoulomb * row_qs[row0 + lr] * pa;
    }
    const double s0 = block_sum<kFinBlock>(e, sm);
    if (threadIdx.x == 0) eblock[blockIdx.x] = s0;
}

// Ordered (deterministic) sum of the per-block partials.
__global__ void reduce_blocks_kernel(int nblock, int ncomp, const double *__restrict__ eblock,
                                     double *__restrict__ out, int out_stride) {
    __shared__ double sm[256 / 32];
    for (int c = 0; c < ncomp; ++c) {
        double v = 0.0;
        for (int b = threadIdx.x; b < nblock; b += 256) v += eblock[(int64_t)b * ncomp + c];
        const double t = block_sum<256>(v, sm);
        if (threadIdx.x == 0) out[c * out_stride] = t;
        __syncthreads();
    }
}

// after the all-reduce over ranks (sharded runs): total from the summed parts
__global__ void assemble_energy_kernel(double *__restrict__ e) {
    if (threadIdx.x == 0 && blockIdx.x == 0) e[UPOL_E_UIND] = (e[UPOL_E_INTER] + e[UPOL_E_INTRA]) + e[UPOL_E_SELF];
}

__global__ void gather3_kernel(int nd, const int *__restrict__ row_site, const double *__restrict__ full,
                               double *__restrict__ compact) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    const int s = row_site[i];
    compact[3 * i] = full[3 * s]; compact[3 * i + 1] = full[3 * s + 1]; compact[3 * i + 2] = full[3 * s + 2];
}

// gradient rows -> the reference's (nsite,3) layout; sites without a Drude (and rows outside [row_lo, row_hi))
// get an exact zero, so no memset precedes it
__global__ void scatter_grad_kernel(int nsite, int row_lo, int row_hi, const int *__restrict__ site_row,
                                    const double *__restrict__ compact, double *__restrict__ full) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsite) return;
    const int row = site_row[i];
    const bool on = row >= row_lo && row < row_hi;
    full[3 * i] = on ? compact[3 * row] : 0.0;
    full[3 * i + 1] = on ? compact[3 * row + 1] : 0.0;
    full[3 * i + 2] = on ? compact[3 * row + 2] : 0.0;
}

__global__ void scatter3_kernel(int nd, const int *__restrict__ row_site, const double *__restrict__ compact,
                                double *__restrict__ full) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    const int s = row_site[i];
    full[3 * s] = compact[3 * i]; full[3 * s + 1] = compact[3 * i + 1]; full[3 * s + 2] = compact[3 * i + 2];
}

// make_Sij on a dense array (polarization.py:27-31)
__global__ void make_sij_kernel(int64_t n, const double *__restrict__ rij, const double *__restrict__ u,
                                int u_is_scalar, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = rij[3 * i], y = rij[3 * i + 1], z = rij[3 * i + 2];
    const double rr = sqrt(x * x + y * y + z * z);
    const double uu = u_is_scalar ? u[0] : u[i];
    out[i] = 1.0 - (1.0 + 0.5 * rr * uu) * exp(-uu * rr);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

int choose_split(const upol_system *s, int64_t n_rows, int n_tiles, int rows_per_block) {
    const int64_t rb = std::max<int64_t>(1, (n_rows + rows_per_block - 1) / rows_per_block);
    static int64_t target = 0;
    if (target == 0) { const char *e = getenv("UPOL_SPLIT_TARGET"); target = e ? atoll(e) : 148 * 40; }   // ~8 waves of resident blocks: short tail
    int64_t sp = (target + rb - 1) / rb;
    sp = std::min<int64_t>(sp, n_tiles);
    sp = std::min<int64_t>(sp, s->max_split);
    // never more splits than the partial-sum buffer holds for this many rows
    sp = std::min<int64_t>(sp, (int64_t)(s->part_elems / ((size_t)kPartComps * (size_t)round_up(std::max<int64_t>(n_rows, 1), 512))));
    return (int)std::max<int64_t>(1, sp);
}

static int64_t part_stride_for(int64_t n_rows) { return round_up(std::max<int64_t>(n_rows, 1), 512); }

static PairArgs base_args(const upol_system *s) {
    PairArgs a;
    memset(&a, 0, sizeof(a));
    a.n_rows = (int)(s->row_hi - s->row_lo);
    a.row0 = (int)s->row_lo;
    a.exA_lo = s->exA_lo; a.exA_len = s->exA_len;
    a.exB_lo = s->exB_lo; a.exB_len = s->exB_len;
    a.n_tiles_a = (int)(s->na_pad / kTileCols);
    a.cx = s->centre[0]; a.cy = s->centre[1]; a.cz = s->centre[2];
    a.inv_grid = s->inv_grid;
    a.part = s->part;
    a.part_stride = part_stride_for(a.n_rows);
    return a;
}

int sys_static_part(upol_system *s, cudaStream_t st) {
    // K3: d-independent remainder of (Ucoul - Ucoul_static): the f(R) terms of
    // polarization.py:82-86 that survive the subtraction of :50-61.
    PairArgs a = base_args(s);
    a.rx = s->srowx; a.ry = s->srowy; a.rz = s->srowz;
    a.cols = s->cols_static;
    a.n_tiles = a.n_tiles_a;
    a.n_split = choose_split(s, a.n_rows, a.n_tiles, fp64_rows_per_block());
    if (a.n_rows > 0) {
        // periodic boxes only (same rows and columns, Ewald pair function, ewald.cu); under open boundaries
        // the static potential is fused into the main fp64 pass (pair_fp64_kernel<FUSE>)
        if (!s->ewald) return UPOL_ERR_ARG;
        int rc = ewald_static_part(s, a, &a.n_split, st);
        if (rc) return rc;
        const int nb = (int)nblk(a.n_rows, kFinRows);
        finalize_static_kernel<<<nb, kFinBlock, 0, st>>>(a.n_rows, a.row0, a.n_split, s->part, a.part_stride,
                                                         s->row_qs, s->phi_static, s->eblock);
        reduce_blocks_kernel<<<1, 256, 0, st>>>(nb, 1, s->eblock, s->e_static, 1);
    } else {
        UPOL_CUDA(cudaMemsetAsync(s->e_static, 0, sizeof(double), st));
    }
    UPOL_CUDA(cudaGetLastError());
    s->static_valid = true;
    return UPOL_OK;
}

int sys_prepare_shells(upol_system *s, const double *d_compact, cudaStream_t st) {
    if (s->nd <= 0) return UPOL_OK;
    prepare_shells_kernel<<<nblk(s->nb_pad, 256), 256, 0, st>>>(
        (int)s->nd, (int)s->nb_pad, s->row_site, s->r, nullptr, const_cast<double *>(d_compact), s->row_qs, s->inv_grid,
        s->rowx, s->rowy, s->rowz, s->cols + s->na_pad, s->cols_d, s->cols_d2);
    ++s->launches;
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

// One evaluation = prepare (shells, shell columns; also gathers d from the (nsite,3) layout when o.d_full is
// set) -> pair kernel(s) -> finalize (fold, K qs, springs, Thole pairs, gradient rows; its last block sums
// the energy vector).  Three launches, no memset.
int sys_eval_compact(upol_system *s, const double *dcmp, const EvalOpts &o, double *energy_dev,
                     double *gcmp, cudaStream_t st) {
    if (o.precision != UPOL_FP64 && o.precision != UPOL_MIXED) return UPOL_ERR_ARG;
    if (s->ewald && (o.precision != UPOL_FP64 || o.phase_switch)) return UPOL_ERR_ARG;
    if (o.d_full != nullptr && dcmp != s->dc) return UPOL_ERR_ARG;
    // Mixed mode: ONE fp32 pass in cancellation-free difference form delivers the row potentials with the
    // static part already removed, and the field (pair_diff.cu); no fp64 pair kernel, no static pass.
    // The minimiser (phase_switch) launches both field kernels and lets the gate word pick one.
    const bool mixed = o.precision == UPOL_MIXED;
    const bool diff_energy = mixed && o.want_energy && !o.phase_switch;
    // fp64 energies need the static potential of the rows: cached per geometry; formed by the Ewald kernels
    // under a box, and inside the main pass (fused) under open boundaries
    const bool fuse_static = o.want_energy && !diff_energy && !s->static_valid && !s->ewald;
    if (o.want_energy && !diff_energy && !s->static_valid && s->ewald) {
        int rc = sys_static_part(s, st);
        if (rc) return rc;
    }
    if (s->nd > 0 && !o.shells_ready) {
        prepare_shells_kernel<<<nblk(s->nb_pad, 256), 256, 0, st>>>(
            (int)s->nd, (int)s->nb_pad, s->row_site, s->r, o.d_full, const_cast<double *>(dcmp), s->row_qs, s->inv_grid,
            s->rowx, s->rowy, s->rowz, s->cols + s->na_pad, s->cols_d, s->cols_d2);
        ++s->launches;
    }
    PairArgs a = base_args(s);
    a.rx = s->rowx; a.ry = s->rowy; a.rz = s->rowz;
    a.n_tiles = (int)((s->na_pad + s->nb_pad) / kTileCols);
    a.gate = o.gate;
    double *e = energy_dev ? energy_dev : s->energy;
    const bool reduce = o.reduce_ranks && s->nranks > 1;
    EvalXchg xg = o.gx;
    if (!reduce || !s->p2p) xg.on = 0;
    if (a.n_rows > 0) {
        const int split64 = choose_split(s, a.n_rows, a.n_tiles, fp64_rows_per_block());
        int split_pot = 0, split_f64 = 0, split_mx = 0;
        const bool f64_field = o.want_grad && (!mixed || o.phase_switch);
        const bool f64_launch = f64_field || (o.want_energy && !diff_energy);
        const bool mx_launch = mixed && (o.want_grad || diff_energy);
        int rc = UPOL_OK;
        if (s->ewald) {
            // periodic box: potential and field partials from the Ewald kernels, same slots
            int slots = 0;
            PairArgs b = a;
            b.cols = s->cols;
            rc = ewald_dynamic_part(s, b, &slots, st);
            if (rc) return rc;
            if (o.want_energy) split_pot = slots;
            split_f64 = slots;
        } else if (f64_launch) {
            PairArgs b = a;
            b.cols = s->cols;
            b.n_split = split64;
            b.dcomp = dcmp; b.col_qeff = s->col_qeff;
            b.gate_skip_mask = kGateDone | (o.phase_switch ? kGatePhaseMixed : 0);
            const bool timed = s->timing && o.gate == nullptr && s->tev_used + 2 <= (int)s->tev.size();
            if (timed) cudaEventRecord(s->tev[s->tev_used], st);
            rc = launch_pair_fp64(b, o.want_energy, f64_field, fuse_static, st);
            if (timed) { cudaEventRecord(s->tev[s->tev_used + 1], st); s->tev_used += 2; }
            if (rc) return rc;
            ++s->launches;
            if (o.want_energy) split_pot = split64;
            if (f64_field) split_f64 = split64;
        }
        if (mx_launch && !s->ewald) {
            PairArgs b = a;
            b.cols = s->cols_x; b.cols_d = s->cols_d; b.cols_d2 = s->cols_d2; b.dcomp = dcmp;
            b.n_tiles_a = (int)(s->nb_pad / kTileCols);
            b.n_tiles = (int)((s->nb_pad + s->nn_pad) / kTileCols);
            b.exA_lo = s->exP_lo; b.exA_len = s->exP_len; b.exB_lo = s->exN_lo; b.exB_len = s->exN_len;
            b.n_split = choose_split(s, a.n_rows, b.n_tiles, diff_rows_per_block(diff_energy));
            b.gate_skip_mask = kGateDone | (o.phase_switch ? kGatePhase64 : 0);
            const bool timed = s->timing && o.gate == nullptr && !f64_launch && s->tev_used + 2 <= (int)s->tev.size();
            if (timed) cudaEventRecord(s->tev[s->tev_used], st);
            rc = launch_pair_diff(b, diff_energy, st);
            if (timed) { cudaEventRecord(s->tev[s->tev_used + 1], st); s->tev_used += 2; }
            if (rc) return rc;
            ++s->launches;
            split_mx = b.n_split;
            if (!o.phase_switch) split_f64 = b.n_split;   // no gate: the mixed partials are THE field
            if (diff_energy) split_pot = b.n_split;
        }
        // small systems: 8 rows per finalize block (16 lanes fold the ~100 column splits of a row), else 32
        const bool fin8 = a.n_rows <= 16384 && std::max(split_pot, std::max(split_f64, split_mx)) >= 16;
        const int nb = (int)nblk(a.n_rows, fin8 ? 8 : kFinRows);
        const int e_flags = ((o.want_energy && !diff_energy && !fuse_static) ? kEPutStatic : 0) | (reduce ? 0 : kETotal) |
                            (fuse_static ? kESaveStatic : 0);
        auto fin = fin8 ? finalize_rows_kernel<8> : finalize_rows_kernel<kFinRows>;
        fin<<<nb, kFinBlock, 0, st>>>(a.n_rows, a.row0, split_pot, split_f64, split_mx, fuse_static ? 1 : 0,
                                      o.gate, o.phase_switch ? 1 : 0, s->part, a.part_stride,
                                      s->row_site, s->r, dcmp, s->row_qs, s->row_k,
                                      s->th_ptr, s->th_row, s->th_u,
                                      (o.want_energy && !diff_energy) ? s->phi_static : nullptr, o.want_grad, gcmp,
                                      s->eblock, s->ticket, e, s->e_static, e_flags, xg,
                                      o.want_grad ? o.g_full : nullptr, (int)s->nd, (int)s->nsite);
        UPOL_CUDA(cudaGetLastError());
        ++s->launches;
        if (fuse_static) s->static_valid = true;
    } else if (xg.on) {
        eval_xchg_empty_kernel<<<1, 32, 0, st>>>(xg);
        ++s->launches;
    } else {
        // a rank without rows contributes zeros (and its cached static sum, which is zero as well)
        UPOL_CUDA(cudaMemsetAsync(e, 0, UPOL_E_SIZE * sizeof(double), st));
    }
    if (xg.on) {
        eval_xchg_wait_kernel<<<1, 32, 0, st>>>(xg, s->xflags2, s->xe, e);
        ++s->launches;
    } else if (reduce) {
        int rc = comm_allreduce_sum(s, e, UPOL_E_SIZE, st);
        if (rc) return rc;
        assemble_energy_kernel<<<1, 32, 0, st>>>(e);
        ++s->launches;
    }
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

EvalXchg make_eval_xchg(upol_system *s, bool push_grad) {
    EvalXchg x = {0, 0, 1, 0, 0, nullptr, nullptr, nullptr, 0, 0};
    if (!s->p2p || s->nranks <= 1) return x;
    x.on = 1; x.rank = s->rank; x.nranks = s->nranks;
    x.epoch = ++s->epoch_eval;
    x.push_grad = push_grad ? 1 : 0;
    x.peer_g = s->d_peer_g; x.peer_e = s->d_peer_e; x.peer_flags = s->d_peer_flags2;
    x.g_off = (int64_t)(x.epoch & 1) * 3 * s->nd;
    x.e_off = (x.epoch & 1) * s->nranks * UPOL_E_SIZE;
    return x;
}

int sys_gather_d(upol_system *s, const double *d_full, double *d_compact, cudaStream_t st) {
    if (s->nd == 0) return UPOL_OK;
    gather3_kernel<<<nblk(s->nd, 256), 256, 0, st>>>((int)s->nd, s->row_site, d_full, d_compact);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

int sys_scatter_g(upol_system *s, const double *g_compact, double *g_full, cudaStream_t st) {
    // after the all-gather of a sharded run every row is present; a single rank shows its row range only
    const int64_t lo = s->nranks > 1 ? 0 : s->row_lo, hi = s->nranks > 1 ? s->nd : s->row_hi;
    scatter_grad_kernel<<<nblk(s->nsite, 256), 256, 0, st>>>((int)s->nsite, (int)lo, (int)hi, s->site_row, g_compact, g_full);
    ++s->launches;
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

// Only the sites [site_lo, site_hi) of the (nsite,3) layout, from this rank's own rows (sharded host path).
__global__ void scatter_grad_range_kernel(int site_lo, int site_hi, int row_lo, int row_hi, const int *__restrict__ site_row,
                                          const double *__restrict__ compact, double *__restrict__ full) {
    const int i = site_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= site_hi) return;
    const int row = site_row[i];
    const bool on = row >= row_lo && row < row_hi;
    full[3 * i] = on ? compact[3 * row] : 0.0;
    full[3 * i + 1] = on ? compact[3 * row + 1] : 0.0;
    full[3 * i + 2] = on ? compact[3 * row + 2] : 0.0;
}

int sys_scatter_g_range(upol_system *s, const double *g_compact, double *g_full, int64_t site_lo, int64_t site_hi, cudaStream_t st) {
    if (site_hi <= site_lo) return UPOL_OK;
    scatter_grad_range_kernel<<<nblk(site_hi - site_lo, 256), 256, 0, st>>>((int)site_lo, (int)site_hi, (int)s->row_lo, (int)s->row_hi,
                                                                            s->site_row, g_compact, g_full);
    ++s->launches;
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

int sys_scatter_d(upol_system *s, const double *d_compact, double *d_full, cudaStream_t st) {
    if (s->nd == 0) return UPOL_OK;
    scatter3_kernel<<<nblk(s->nd, 256), 256, 0, st>>>((int)s->nd, s->row_site, d_compact, d_full);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

int sys_ensure_pinned(upol_system *s, size_t bytes) {
    if (s->pin_bytes >= bytes) return UPOL_OK;
    if (s->pin) cudaFreeHost(s->pin);
    s->pin = nullptr; s->pin_bytes = 0;
    UPOL_CUDA(cudaMallocHost((void **)&s->pin, bytes));
    s->pin_bytes = bytes;
    return UPOL_OK;
}

int sys_ensure_hscratch(upol_system *s, size_t elems) {
    if (s->hscratch_elems >= elems) return UPOL_OK;
    if (s->hscratch) upol::dfree(s->hscratch);
    s->hscratch = nullptr; s->hscratch_elems = 0;
    UPOL_CUDA(upol::dmalloc((void **)&s->hscratch, sizeof(double) * elems));
    s->hscratch_elems = elems;
    return UPOL_OK;
}

// Bounding box of the positions: every block reduces its stripe, the block that finishes last folds the block
// results (a single block took 110 us for 180 000 sites, 2 % of an 8-GPU end-to-end step).
__global__ void __launch_bounds__(256)
bbox_kernel(int nsite, const double *__restrict__ r, double *__restrict__ partial, int *__restrict__ ticket, double *__restrict__ out6) {
    __shared__ double lo[3][8], hi[3][8];
    __shared__ int is_last;
    double l[3] = {1e300, 1e300, 1e300}, h[3] = {-1e300, -1e300, -1e300};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nsite; i += gridDim.x * blockDim.x)
        for (int c = 0; c < 3; ++c) { const double v = r[3 * i + c]; l[c] = fmin(l[c], v); h[c] = fmax(h[c], v); }
    auto block_reduce = [&]() {
        for (int c = 0; c < 3; ++c) {
            for (int o = 16; o > 0; o >>= 1) {
                l[c] = fmin(l[c], __shfl_down_sync(0xffffffffu, l[c], o));
                h[c] = fmax(h[c], __shfl_down_sync(0xffffffffu, h[c], o));
            }
            if ((threadIdx.x & 31) == 0) { lo[c][threadIdx.x >> 5] = l[c]; hi[c][threadIdx.x >> 5] = h[c]; }
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int c = 0; c < 3; ++c) {
                for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo[c][0] = fmin(lo[c][0], lo[c][w]); hi[c][0] = fmax(hi[c][0], hi[c][w]); }
            }
        __syncthreads();
    };
    block_reduce();
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c) { partial[6 * blockIdx.x + c] = lo[c][0]; partial[6 * blockIdx.x + 3 + c] = hi[c][0]; }
        __threadfence();
        is_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int c = 0; c < 3; ++c) { l[c] = 1e300; h[c] = -1e300; }
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
        for (int c = 0; c < 3; ++c) { l[c] = fmin(l[c], __ldcg(partial + 6 * b + c)); h[c] = fmax(h[c], __ldcg(partial + 6 * b + 3 + c)); }
    block_reduce();
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c) { out6[c] = lo[c][0]; out6[3 + c] = hi[c][0]; }
        *ticket = 0;
    }
}

void set_box(upol_system *s, const double *lo, const double *hi) {
    double half = 0.0;
    for (int c = 0; c < 3; ++c) {
        s->centre[c] = s->nsite ? 0.5 * (lo[c] + hi[c]) : 0.0;
        half = std::max(half, s->nsite ? 0.5 * (hi[c] - lo[c]) : 0.0);
    }
    // fixed-point grid of the mixed kernel: 2^-k nm with (half extent + 1 nm of Drude slack)
    // inside +-2^29, so coordinate differences fit a 32-bit int with room to spare
    int e = 0;
    frexp(half + 1.0, &e);                      // half + 1 < 2^e
    s->inv_grid = ldexp(1.0, 29 - e);
}

int sys_box_from_device(upol_system *s, cudaStream_t st) {
    const unsigned nbb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(148, (s->nsite + 2047) / 2048));
    bbox_kernel<<<nbb, 256, 0, st>>>((int)s->nsite, s->r, s->part, s->ticket + 1, s->eblock);
    double h[6];
    UPOL_CUDA(cudaMemcpyAsync(h, s->eblock, sizeof(h), cudaMemcpyDeviceToHost, st));
    UPOL_CUDA(cudaStreamSynchronize(st));
    set_box(s, h, h + 3);
    return UPOL_OK;
}

int sys_upload_positions(upol_system *s, const double *r_dev_or_null, cudaStream_t st) {
    // r is already in s->r (device); (re)build the d-independent columns with the current box
    (void)r_dev_or_null;
    build_core_columns_kernel<<<nblk(s->na_pad, 256), 256, 0, st>>>(
        (int)s->nsite, (int)s->na_pad, s->r, s->qc, s->qs, s->cols, s->cols_static, s->col_qeff);
    build_diff_columns_kernel<<<nblk(std::max<int64_t>(s->nsite, std::max(s->nb_pad, s->nn_pad)), 256), 256, 0, st>>>(
        (int)s->nsite, (int)s->nd, (int)s->nb_pad, (int)s->nn_pad, s->r, s->qc, s->site_row, s->site_nslot,
        s->centre[0], s->centre[1], s->centre[2], s->inv_grid, s->cols_x);
    if (s->nd > 0)
        build_static_rows_kernel<<<nblk(s->nd, 256), 256, 0, st>>>((int)s->nd, s->row_site, s->r,
                                                                    s->srowx, s->srowy, s->srowz);
    UPOL_CUDA(cudaGetLastError());
    s->launches += 2 + (s->nd > 0 ? 1 : 0);
    s->static_valid = false;
    ++s->geom_gen;
    ewald_invalidate(s);
    return UPOL_OK;
}

// Uself = 1/2 sum k |d|^2 (polarization.py:44-48) on plain arrays; one block, fixed summation order
__global__ void uself_kernel(int64_t n, const double *__restrict__ d, const double *__restrict__ k, double *__restrict__ out) {
    __shared__ double sm[256 / 32];
    double v = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256)
        v += 0.5 * k[i] * (d[3 * i] * d[3 * i] + d[3 * i + 1] * d[3 * i + 1] + d[3 * i + 2] * d[3 * i + 2]);
    const double t = block_sum<256>(v, sm);
    if (threadIdx.x == 0) out[0] = t;
}

int sys_uself(const double *d, const double *k, int64_t n, double *out, cudaStream_t st) {
    uself_kernel<<<1, 256, 0, st>>>(n, d, k, out);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

int sys_make_sij(const double *rij, const double *u, int u_is_scalar, int64_t n, double *out, cudaStream_t st) {
    if (n <= 0) return UPOL_OK;
    make_sij_kernel<<<nblk(n, 256), 256, 0, st>>>(n, rij, u, u_is_scalar, out);
    UPOL_CUDA(cudaGetLastError());
    return UPOL_OK;
}

// d-independent sums for Ucoul / Ucoul_static as separate values (polarization.py:50-61,82):
// rows = every site, columns = cores carrying q_row-matching charges.
__global__ void site_rows_kernel(int nsite, const double *__restrict__ r, double *__restrict__ x,
                                 double *__restrict__ y, double *__restrict__ z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsite) return;
    x[i] = r[3 * i]; y[i] = r[3 * i + 1]; z[i] = r[3 * i + 2];
}
__global__ void charge_columns_kernel(int nsite, int na_pad, const double *__restrict__ r,
                                      const double *__restrict__ qc, const double *__restrict__ qs,
                                      int total, double4 *__restrict__ cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na_pad) return;
    if (i < nsite) cols[i] = make_double4(r[3 * i], r[3 * i + 1], r[3 * i + 2], total ? qc[i] + qs[i] : qc[i]);
    else cols[i] = make_double4(kFar, kFar, kFar, 0.0);
}
__global__ void __launch_bounds__(kFinBlock)
finalize_sites_kernel(int nsite, int n_split, const double *__restrict__ part, int64_t stride,
                      const double *__restrict__ qc, const double *__restrict__ qs, int total,
                      double *__restrict__ eblock) {
    __shared__ double sm[kFinBlock / 32];
    __shared__ double fold[kFinLanes][kFinRows][kPartComps];
    const int i = blockIdx.x * kFinRows + threadIdx.x % kFinRows;
    const bool live = i < nsite;
    double pot[1] = {0.0};
    fold_splits<1>(part, stride, i, live, n_split, 0, fold, pot);
    double e = 0.0;
    if (live && threadIdx.x < kFinRows) {
        const double pa = pot[0];
        e = 0.5 * kCoulomb * (total ? qc[i] + qs[i] : qc[i]) * pa;
    }
    const double s0 = block_sum<kFinBlock>(e, sm);
    if (threadIdx.x == 0) eblock[blockIdx.x] = s0;
}

int sys_coulomb_static(upol_system *s, double *out2, cudaStream_t st) {
    // scratch: rows (3*nsite doubles), site exclusion intervals, columns, partials
    const int64_t n = s->nsite;
    // scratch freed on every exit path (a CUDA error on the way must not leak it: Ucoul / Ucoul_static call
    // this once per evaluation)
    struct Scratch {
        double *rows = nullptr, *part = nullptr, *eblock = nullptr;
        int *ex = nullptr;
        double4 *cols = nullptr;
        ~Scratch() { upol::dfree(rows); upol::dfree(part); upol::dfree(eblock); upol::dfree(ex); upol::dfree(cols); }
    } sc;
    const int ntile = (int)(s->na_pad / kTileCols);
    const int64_t rb = (n + kRowTile - 1) / kRowTile;
    int split = (int)std::max<int64_t>(1, std::min<int64_t>(ntile, (148 * 8 + rb - 1) / rb));
    const int64_t stride = part_stride_for(n);
    const int nb = (int)nblk(n, kFinRows);
    UPOL_CUDA(upol::dmalloc((void **)&sc.rows, sizeof(double) * 3 * n));
    UPOL_CUDA(upol::dmalloc((void **)&sc.ex, sizeof(int) * 4 * n));
    UPOL_CUDA(upol::dmalloc((void **)&sc.cols, sizeof(double4) * s->na_pad));
    UPOL_CUDA(upol::dmalloc((void **)&sc.part, sizeof(double) * split * kPartComps * stride));
    UPOL_CUDA(upol::dmalloc((void **)&sc.eblock, sizeof(double) * nb));
    double *rows = sc.rows, *part = sc.part, *eblock = sc.eblock;
    int *ex = sc.ex;
    double4 *cols = sc.cols;
    std::vector<int> hex(4 * n, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int m = s->h_mol[i];
        hex[i] = s->h_mol_site0[m]; hex[n + i] = s->h_mol_nsite[m];
    }
    UPOL_CUDA(cudaMemcpyAsync(ex, hex.data(), sizeof(int) * 4 * n, cudaMemcpyHostToDevice, st));
    site_rows_kernel<<<nblk(n, 256), 256, 0, st>>>((int)n, s->r, rows, rows + n, rows + 2 * n);
    for (int total = 0; total < 2; ++total) {
        charge_columns_kernel<<<nblk(s->na_pad, 256), 256, 0, st>>>((int)n, (int)s->na_pad, s->r, s->qc, s->qs, total, cols);
        PairArgs a;
        memset(&a, 0, sizeof(a));
        a.n_rows = (int)n; a.row0 = 0;
        a.rx = rows; a.ry = rows + n; a.rz = rows + 2 * n;
        a.exA_lo = ex; a.exA_len = ex + n; a.exB_lo = ex + 2 * n; a.exB_len = ex + 3 * n;
        a.cols = cols; a.n_tiles_a = ntile; a.n_tiles = ntile;
        a.n_split = split; a.part = part; a.part_stride = stride;
        int rc = launch_pair_fp64(a, true, false, false, st);
        if (rc) return rc;
        finalize_sites_kernel<<<nb, kFinBlock, 0, st>>>((int)n, split, part, stride, s->qc, s->qs, total, eblock);
        reduce_blocks_kernel<<<1, 256, 0, st>>>(nb, 1, eblock, out2 + total, 1);
    }
    UPOL_CUDA(cudaGetLastError());
    UPOL_CUDA(cudaStreamSynchronize(st));   // hex (host vector) and the scratch die here
    return UPOL_OK;
}

// ---------------------------------------------------------------------------------------
// Core forces: dU_ind/dr_i at fixed displacements (SURVEY 8f "next" #3, first half).  The reference
// has no such function; it is what an MD code needs to move the cores on the U_ind surface.
//   dU/dr_a = [a Drude] K qs_a (F_a - H_a)  +  K qc_a G_a  -  K (qc_a + qs_a/2) H'_a  +  intra
// F_a : field at the shell s_a from all other-molecule charges        (main pass, field only)
// H_a : field at the core r_a from the static charges qc + qs/2        (static columns, field)
// G_a : field at r_a from the other molecules' shells (qs at s)        } one pass over all site rows,
// H'_a: field at r_a from the other molecules' Drude CORES (qs at r)   } two fields kept apart
// intra: all four Thole terms of a screened pair depend on R = r_b - r_a alike.
// ---------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kFinBlock)
fold_fields_kernel(int n_rows, int row0, int64_t dst_stride, int n_split, const double *__restrict__ part,
                   int64_t stride, int c0, double *__restrict__ dst) {
    __shared__ double fold[kFinLanes][kFinRows][kPartComps];
    const int lr = blockIdx.x * kFinRows + threadIdx.x % kFinRows;
    const bool live = lr < n_rows;
    double v[NC];
    fold_splits<NC>(part, stride, lr, live, n_split, c0, fold, v);
    if (live && threadIdx.x < kFinRows) {
#pragma unroll
        for (int c = 0; c < NC; ++c) dst[(int64_t)c * dst_stride + row0 + lr] = v[c];
    }
}

__global__ void build_force_tables_kernel(int nsite, int nd, int nb_pad, const double *__restrict__ r,
                                          const int *__restrict__ row_site, const double *__restrict__ row_qs,
                                          double *__restrict__ sx, double *__restrict__ sy, double *__restrict__ sz,
                                          double4 *__restrict__ cols_dcore) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nsite) { sx[i] = r[3 * i]; sy[i] = r[3 * i + 1]; sz[i] = r[3 * i + 2]; }
    if (i < nb_pad) {
        if (i < nd) { const int s = row_site[i]; cols_dcore[i] = make_double4(r[3 * s], r[3 * s + 1], r[3 * s + 2], row_qs[i]); }
        else cols_dcore[i] = make_double4(kFar, kFar, kFar, 0.0);
    }
}

__global__ void core_force_assemble_kernel(int site_lo, int site_hi, int nsite, int nd, const double *__restrict__ r, const double *__restrict__ dc,
                                           const double *__restrict__ qc, const double *__restrict__ qs,
                                           const int *__restrict__ site_row, const int *__restrict__ row_site,
                                           const double *__restrict__ row_qs, const int *__restrict__ th_ptr,
                                           const int *__restrict__ th_row, const double *__restrict__ th_u,
                                           const double *__restrict__ F, const double *__restrict__ H,
                                           const double *__restrict__ GH, double *__restrict__ out) {
    const int j = site_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= site_hi) return;
    const double qe = qc[j] + 0.5 * qs[j];
    double g[3];
    for (int c = 0; c < 3; ++c) g[c] = kCoulomb * (qc[j] * GH[(int64_t)c * nsite + j] - qe * GH[(int64_t)(3 + c) * nsite + j]);
    const int a = site_row[j];
    if (a >= 0) {
        const double q = row_qs[a];
        for (int c = 0; c < 3; ++c) g[c] += kCoulomb * q * (F[(int64_t)c * nd + a] - H[(int64_t)c * nd + a]);
        const double dx = dc[3 * a], dy = dc[3 * a + 1], dz = dc[3 * a + 2];
        for (int p = th_ptr[a]; p < th_ptr[a + 1]; ++p) {
            const int b = th_row[p], sb = row_site[b];
            const double u = th_u[p];
            const double Rx = r[3 * sb] - r[3 * j], Ry = r[3 * sb + 1] - r[3 * j + 1], Rz = r[3 * sb + 2] - r[3 * j + 2];
            const double bx = dc[3 * b], by = dc[3 * b + 1], bz = dc[3 * b + 2];
            const double Ax = Rx + dx, Ay = Ry + dy, Az = Rz + dz, Bx = Rx - bx, By = Ry - by, Bz = Rz - bz;
            const double Cx = Ax - bx, Cy = Ay - by, Cz = Az - bz;
            double gR, gA, gB, gC, cR, cA, cB, cC;
            thole_g(Rx, Ry, Rz, u, gR, cR); thole_g(Ax, Ay, Az, u, gA, cA);
            thole_g(Bx, By, Bz, u, gB, cB); thole_g(Cx, Cy, Cz, u, gC, cC);
            const double qq = kCoulomb * q * row_qs[b];
            // d/dr_a of qq [g(R) - g(A) - g(B) + g(C)] with R = r_b - r_a: minus the gradient w.r.t. R
            g[0] -= qq * (cR * Rx - cA * Ax - cB * Bx + cC * Cx);
            g[1] -= qq * (cR * Ry - cA * Ay - cB * By + cC * Cy);
            g[2] -= qq * (cR * Rz - cA * Az - cB * Bz + cC * Cz);
        }
    }
    out[3 * j] = g[0]; out[3 * j + 1] = g[1]; out[3 * j + 2] = g[2];
}

int sys_core_forces(upol_system *s, const double *dcmp, double *dudr, cudaStream_t st) {
    const int64_t n = s->nsite, nd = s->nd;
    if (!s->cf_tmp) {      // first use: tables of the site-row pass
        s->ns_pad = round_up(n, 512);
        UPOL_CUDA(upol::dmalloc((void **)&s->cols_dcore, sizeof(double4) * s->nb_pad));
        UPOL_CUDA(upol::dmalloc((void **)&s->site_x, sizeof(double) * 3 * s->ns_pad));
        s->site_y = s->site_x + s->ns_pad; s->site_z = s->site_y + s->ns_pad;
        UPOL_CUDA(upol::dmalloc((void **)&s->exS_lo, sizeof(int) * 3 * s->ns_pad));
        s->exS_len = s->exS_lo + s->ns_pad; s->exS2_lo = s->exS_len + s->ns_pad;
        UPOL_CUDA(upol::dmalloc((void **)&s->cf_tmp, sizeof(double) * (6 * std::max<int64_t>(nd, 1) + 6 * n)));
        std::vector<int> ex(3 * s->ns_pad, 0);
        for (int64_t j = 0; j < n; ++j) {
            const int m = s->h_mol[j];
            ex[j] = s->h_mol_row0[m]; ex[s->ns_pad + j] = s->h_mol_nrow[m]; ex[2 * s->ns_pad + j] = (int)s->nb_pad + s->h_mol_row0[m];
        }
        UPOL_CUDA(cudaMemcpy(s->exS_lo, ex.data(), sizeof(int) * ex.size(), cudaMemcpyHostToDevice));
        UPOL_CUDA(cudaMemset(s->site_x, 0, sizeof(double) * 3 * s->ns_pad));
        // the copies above ran on the legacy stream from pageable memory; `st` may be a non-blocking stream
        // that is not ordered against it, so the tables must have landed before anything is enqueued on `st`
        UPOL_CUDA(cudaDeviceSynchronize());
    }
    UPOL_CUDA(cudaMemsetAsync(dudr, 0, sizeof(double) * 3 * n, st));
    if (nd == 0) return UPOL_OK;
    // this rank's share: Drude rows and sites of its molecule block (all of them on a single rank)
    const int64_t r_lo = s->row_lo, nr = s->row_hi - s->row_lo;
    const int64_t s_lo = s->mol_lo < s->nmol ? s->h_mol_site0[s->mol_lo] : n;
    const int64_t s_hi = s->mol_hi < s->nmol ? s->h_mol_site0[s->mol_hi] : n;
    const int64_t ns = s_hi - s_lo;
    double *F = s->cf_tmp, *H = F + 3 * nd, *GH = H + 3 * nd;
    prepare_shells_kernel<<<nblk(s->nb_pad, 256), 256, 0, st>>>(
        (int)nd, (int)s->nb_pad, s->row_site, s->r, nullptr, const_cast<double *>(dcmp), s->row_qs, s->inv_grid,
        s->rowx, s->rowy, s->rowz, s->cols + s->na_pad, s->cols_d, s->cols_d2);
    build_force_tables_kernel<<<nblk(std::max<int64_t>(n, s->nb_pad), 256), 256, 0, st>>>(
        (int)n, (int)nd, (int)s->nb_pad, s->r, s->row_site, s->row_qs, s->site_x, s->site_y, s->site_z, s->cols_dcore);
    if (s->ewald) {
        // periodic box: the same four fields from the Ewald kernels (ewald.cu), one column set at a time
        int slots = 0, rc = UPOL_OK;
        if (nr > 0) {
            const int nbf = (int)nblk(nr, kFinRows);
            PairArgs a = base_args(s);
            a.rx = s->rowx; a.ry = s->rowy; a.rz = s->rowz;
            a.cols = s->cols; a.n_tiles = (int)((s->na_pad + s->nb_pad) / kTileCols);
            rc = ewald_dynamic_part(s, a, &slots, st);                       // F (also refreshes S of the shells)
            if (rc) return rc;
            fold_fields_kernel<3><<<nbf, kFinBlock, 0, st>>>((int)nr, (int)r_lo, nd, slots, s->part, a.part_stride, 2, F);
            a = base_args(s);
            a.rx = s->srowx; a.ry = s->srowy; a.rz = s->srowz;
            a.cols = s->cols_static; a.n_tiles = a.n_tiles_a;
            rc = ewald_field_part(s, a, 2, &slots, st);                      // H
            if (rc) return rc;
            fold_fields_kernel<3><<<nbf, kFinBlock, 0, st>>>((int)nr, (int)r_lo, nd, slots, s->part, a.part_stride, 2, H);
        }
        if (ns > 0) {
            if (nr == 0) {                                                   // S of the shells is still needed
                PairArgs a0 = base_args(s);
                a0.rx = s->rowx; a0.ry = s->rowy; a0.rz = s->rowz; a0.cols = s->cols;
                a0.n_tiles = (int)((s->na_pad + s->nb_pad) / kTileCols);
                rc = ewald_dynamic_part(s, a0, &slots, st);
                if (rc) return rc;
            }
            PairArgs a = base_args(s);
            a.n_rows = (int)ns; a.row0 = (int)s_lo; a.part_stride = part_stride_for(ns);
            a.rx = s->site_x; a.ry = s->site_y; a.rz = s->site_z;
            a.exA_lo = s->exS_lo; a.exA_len = s->exS_len;
            a.n_tiles = (int)(s->nb_pad / kTileCols);
            a.cols = s->cols + s->na_pad;
            rc = ewald_field_part(s, a, 0, &slots, st);                      // G: shells at r - d
            if (rc) return rc;
            fold_fields_kernel<3><<<nblk(ns, kFinRows), kFinBlock, 0, st>>>((int)ns, (int)s_lo, n, slots, s->part, a.part_stride, 2, GH);
            a.cols = s->cols_dcore;
            rc = ewald_field_part(s, a, 1, &slots, st);                      // H': Drude cores
            if (rc) return rc;
            fold_fields_kernel<3><<<nblk(ns, kFinRows), kFinBlock, 0, st>>>((int)ns, (int)s_lo, n, slots, s->part, a.part_stride, 2, GH + 3 * n);
            core_force_assemble_kernel<<<nblk(ns, 128), 128, 0, st>>>((int)s_lo, (int)s_hi, (int)n, (int)nd, s->r, dcmp, s->qc, s->qs,
                                                                     s->site_row, s->row_site, s->row_qs, s->th_ptr, s->th_row,
                                                                     s->th_u, F, H, GH, dudr);
        }
        UPOL_CUDA(cudaGetLastError());
        if (s->nranks > 1) return comm_allgather_sites(s, dudr, st);
        return UPOL_OK;
    }
    if (nr > 0) {
        const int nbf = (int)nblk(nr, kFinRows);
        {   // F: field at the shells (main pass, field only)
            PairArgs a = base_args(s);
            a.rx = s->rowx; a.ry = s->rowy; a.rz = s->rowz;
            a.cols = s->cols; a.n_tiles = (int)((s->na_pad + s->nb_pad) / kTileCols);
            a.n_split = choose_split(s, nr, a.n_tiles, fp64_rows_per_block());
            int rc = launch_pair_fp64(a, false, true, false, st);
            if (rc) return rc;
            fold_fields_kernel<3><<<nbf, kFinBlock, 0, st>>>((int)nr, (int)r_lo, nd, a.n_split, s->part, a.part_stride, 2, F);
        }
        {   // H: field at the Drude cores from the static charges
            PairArgs a = base_args(s);
            a.rx = s->srowx; a.ry = s->srowy; a.rz = s->srowz;
            a.cols = s->cols_static; a.n_tiles = a.n_tiles_a;
            a.n_split = choose_split(s, nr, a.n_tiles, fp64_rows_per_block());
            int rc = launch_pair_fp64(a, false, true, false, st);
            if (rc) return rc;
            fold_fields_kernel<3><<<nbf, kFinBlock, 0, st>>>((int)nr, (int)r_lo, nd, a.n_split, s->part, a.part_stride, 2, H);
        }
    }
    if (ns > 0) {   // G, H': fields at every core of the block from the other molecules' shells / Drude cores
        PairArgs a = base_args(s);
        a.n_rows = (int)ns; a.row0 = (int)s_lo; a.part_stride = part_stride_for(ns);
        a.rx = s->site_x; a.ry = s->site_y; a.rz = s->site_z;
        a.exA_lo = s->exS_lo; a.exA_len = s->exS_len; a.exB_lo = s->exS2_lo; a.exB_len = s->exS_len;
        a.cols = s->cols + s->na_pad;           // shells (refreshed by prepare_shells above)
        a.cols_b = s->cols_dcore;
        a.n_tiles_a = (int)(s->nb_pad / kTileCols); a.n_tiles = 2 * a.n_tiles_a;
        a.n_split = choose_split(s, ns, a.n_tiles, fp64_rows_per_block());
        int rc = launch_pair_fp64_two_fields(a, st);
        if (rc) return rc;
        fold_fields_kernel<6><<<nblk(ns, kFinRows), kFinBlock, 0, st>>>((int)ns, (int)s_lo, n, a.n_split, s->part, a.part_stride, 0, GH);
        core_force_assemble_kernel<<<nblk(ns, 128), 128, 0, st>>>((int)s_lo, (int)s_hi, (int)n, (int)nd, s->r, dcmp, s->qc, s->qs,
                                                                 s->site_row, s->row_site, s->row_qs, s->th_ptr, s->th_row,
                                                                 s->th_u, F, H, GH, dudr);
    }
    UPOL_CUDA(cudaGetLastError());
    if (s->nranks > 1) return comm_allgather_sites(s, dudr, st);
    return UPOL_OK;
}

}  // namespace upol
