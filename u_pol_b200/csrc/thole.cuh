// Thole-screened pair function of the intra-molecular Drude dipole-dipole term and the zero-distance
// rule, shared by the evaluation pipeline (system.cu), the minimiser's preconditioner (minimize.cu) and
// the batched small-system kernel (batched.cu).
#pragma once

namespace upol {

// Zero-distance rule of the device kernels.  The reference drops a pair term whose distance is EXACTLY zero
// (optax safe_norm / the NaN-safe norm and sum of U_pol/polarization.py:33-42).  For the inter-molecular
// 1/r terms every kernel here (pair_kernels.cu, pair_diff.cu, ewald.cu, batched.cu) drops a term when
// r^2 < kZeroR2 = 1e-10 nm^2 (r < 1e-5 nm, five orders of magnitude below any physical contact): the
// fixed-point / expanded-square forms cannot tell an exact zero from their round-off.  The intra-molecular
// Thole function below keeps the exact test: S(ur)/r is finite at r -> 0, so no threshold is needed there.
constexpr double kZeroR2 = 1.0e-10;
constexpr int kZeroR2Hi = 0x3DDB7CDF;     // high word of kZeroR2 (compare with __double2hiint(r2))

// g(x) = S(u|x|)/|x| with S = 1 - (1 + u r/2) exp(-u r)  (make_Sij, polarization.py:27-31; used at :94-99)
// and the coefficient c of its gradient, grad g(x) = c x.  Zero norm -> 0 (polarization.py:33-37).
__device__ __forceinline__ void thole_g(double x, double y, double z, double u, double &g, double &c) {
    const double r2 = x * x + y * y + z * z;
    if (r2 == 0.0) { g = 0.0; c = 0.0; return; }
    const double rr = sqrt(r2);
    const double ex = exp(-u * rr);
    const double S = 1.0 - (1.0 + 0.5 * rr * u) * ex;
    const double Sp = 0.5 * u * (1.0 + u * rr) * ex;
    const double ir = 1.0 / rr;
    g = S * ir;
    c = (Sp * ir - S * ir * ir) * ir;
}

// Second derivative tensor of g:  grad g = c(r) x,  c = (S' r - S)/r^3;   grad grad g = c I + (c'/r) x x^T,
//   c' = S''/r^2 - 3 (S' r - S)/r^4,  S' = u(1+ur)e^{-ur}/2,  S'' = -u^3 r e^{-ur}/2.
__device__ inline void thole_hessian(double x, double y, double z, double u, double G[3][3]) {
    const double r2 = x * x + y * y + z * z;
    if (r2 == 0.0) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) G[i][j] = 0.0; return; }
    const double r = sqrt(r2), ex = exp(-u * r);
    const double S = 1.0 - (1.0 + 0.5 * r * u) * ex, Sp = 0.5 * u * (1.0 + u * r) * ex, Spp = -0.5 * u * u * u * r * ex;
    const double t = Sp * r - S;
    const double c = t / (r2 * r), cpr = (Spp / r2 - 3.0 * t / (r2 * r2)) / r;
    const double v[3] = {x, y, z};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i][j] = (i == j ? c : 0.0) + cpr * v[i] * v[j];
}

// 1/sqrt(x) in fp64: MUFU.RSQ64H seed (rel. error <= 2^-22) + one third-order (Halley) step:
// 1/sqrt(x) = y (1-e)^(-1/2),  e = 1 - x y^2  ->  y (1 + e (1/2 + 3e/8)), error O(e^3) ~ 1e-20.
// Written as y * fma(e, p, 1) rather than fma(y e, p, y): same five DP instructions, but none of them
// reads three distinct registers (a DFMA with three 64-bit register operands takes 3.1 cycles per warp
// instead of 2.2 - register-operand bandwidth, profiles/r02_ubench2.txt).
__device__ __forceinline__ double rsqrt_f64(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = y * y;
    const double e = fma(-x, t, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double c = fma(e, p, 1.0);
    return y * c;
}

}  // namespace upol
