/*
 * TEST INFRASTRUCTURE - C restatement of the reference's U_ind path, used as the parity
 * checker at sizes the dense torch oracle cannot hold and as the CPU baseline in bench.py.
 * Not part of the product; nothing under u_pol_b200/ links or calls this.
 *
 * Follows U_pol/polarization.py term by term, per ORDERED site pair (i,j), i != j, exactly as
 * the reference's (nmol,nmol,natoms,natoms) tensors enumerate them:
 *   Rij = r_j - r_i                               util.py:78-81
 *   A = Rij + d_i, B = Rij - d_j, C = Rij + d_i - d_j        polarization.py:70-73
 *   |x| = 0 -> +inf (term contributes 0)          polarization.py:33-37
 *   different molecules:  qc_i qc_j/|R| + qs_i qc_j/|A| + qc_i qs_j/|B| + qs_i qs_j/|C|   :82-91
 *                         static: (qc_i+qs_i)(qc_j+qs_j)/|R|                            :50-61
 *   same molecule, a != b: qs_i qs_j [ S(R)/|R| - S(A)/|A| - S(B)/|B| + S(C)/|C| ]        :94-105
 *                         S(x) = 1 - (1 + u|x|/2) exp(-u|x|)                           :27-31
 *   U_ind = K/2 sum(inter+intra) - K/2 sum(static) + 1/2 sum k_i |d_i|^2               :107-109,:149
 * The reference evaluates S (four exp) for every pair, also where u = 0 and S == 0; this port
 * skips the exp when u == 0 (identical result, and it only makes the CPU baseline faster).
 *
 * Gradient: the reference differentiates the expression above by reverse-mode AD
 * (jax.grad inside jaxopt.BFGS, polarization.py:172-179).  Here the same derivative is written
 * out: d_i enters pair (i,j) through A and C and pair (j,i) through B' = -A and C' = -C, so
 *   dU/dd_i = K sum_j [ qs_i qc_j grad f(A) + qs_i qs_j grad f(C) ]           (inter)
 *           + K sum_j qs_i qs_j [ -grad g(A) + grad g(C) ]                    (intra)
 *           + k_i d_i,        f(x) = 1/|x|, g(x) = S(x)/|x|.
 * It is validated against autodiff of the dense oracle in tests/test_oracle_c.py.
 *
 * Rows [row_lo, row_hi) only: the caller can evaluate a bounded sample of a large system
 * (every row still sees all columns), which is how bench.py times the CPU baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const double M_PI_REF = 3.14159265358979323846;
static const double E_CHARGE = 1.602176634e-19;
static const double AVOGADRO = 6.02214076e23;

double upol_oracle_one_4pi_eps0(void) {
    /* polarization.py:19-25 */
    const double eps0 = 1e-6 * 8.8541878128e-12 / (E_CHARGE * E_CHARGE * AVOGADRO);
    return 1 / (4 * M_PI_REF * eps0);
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm asks for all cores back. */
void upol_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int upol_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline double inv_norm(double x, double y, double z, double *r_out) {
    const double r = sqrt(x * x + y * y + z * z);
    *r_out = r;
    return r == 0.0 ? 0.0 : 1.0 / r;          /* 1/inf */
}

/* g(x) = S/|x| and c with grad g = c * x */
static inline void thole_term(double x, double y, double z, double u, double *g, double *c) {
    double r;
    const double ir = inv_norm(x, y, z, &r);
    if (r == 0.0) { *g = 0.0; *c = 0.0; return; }
    const double ex = exp(-u * r);
    const double S = 1.0 - (1.0 + 0.5 * r * u) * ex;
    const double Sp = 0.5 * u * (1.0 + u * r) * ex;
    *g = S * ir;
    *c = (Sp * ir - S * ir * ir) * ir;
}

/*
 * r, d: [nsite*3]; qc, qs, k: [nsite]; sites of molecule m are [m*natoms, (m+1)*natoms);
 * thole_u: [nmol*natoms*natoms] or NULL (reference: u_scale = 0.0, util.py:277).
 * energy[0] = U_ind share of the rows, [1] = K/2 sum(inter+intra), [2] = K/2 sum(static), [3] = U_self.
 * grad: [nsite*3], only rows in range are written (may be NULL).
 */
int upol_oracle_energy_grad(int64_t nsite, int natoms, const double *r, const double *d, const double *qc,
                            const double *qs, const double *k, const double *thole_u, int64_t row_lo,
                            int64_t row_hi, double *energy, double *grad) {
    if (nsite <= 0 || natoms <= 0 || nsite % natoms != 0 || row_lo < 0 || row_hi > nsite) return -2;
    const double K = upol_oracle_one_4pi_eps0();
    double e_coul = 0.0, e_stat = 0.0, e_self = 0.0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : e_coul, e_stat, e_self)
    for (int64_t i = row_lo; i < row_hi; ++i) {
        const int64_t mi = i / natoms;
        const int ai = (int)(i % natoms);
        const double xi = r[3 * i], yi = r[3 * i + 1], zi = r[3 * i + 2];
        const double dxi = d[3 * i], dyi = d[3 * i + 1], dzi = d[3 * i + 2];
        const double qci = qc[i], qsi = qs[i];
        double ec = 0.0, es = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
        for (int64_t j = 0; j < nsite; ++j) {
            if (j == i) continue;                                 /* a == b in the same molecule: masked, :101-105 */
            const double Rx = r[3 * j] - xi, Ry = r[3 * j + 1] - yi, Rz = r[3 * j + 2] - zi;
            const double Ax = Rx + dxi, Ay = Ry + dyi, Az = Rz + dzi;
            const double Bx = Rx - d[3 * j], By = Ry - d[3 * j + 1], Bz = Rz - d[3 * j + 2];
            const double Cx = Ax - d[3 * j], Cy = Ay - d[3 * j + 1], Cz = Az - d[3 * j + 2];
            const double qcj = qc[j], qsj = qs[j];
            if (j / natoms != mi) {
                double rr, ra, rb, rc;
                const double iR = inv_norm(Rx, Ry, Rz, &rr), iA = inv_norm(Ax, Ay, Az, &ra);
                const double iB = inv_norm(Bx, By, Bz, &rb), iC = inv_norm(Cx, Cy, Cz, &rc);
                ec += qci * qcj * iR + qsi * qcj * iA + qci * qsj * iB + qsi * qsj * iC;
                es += (qci + qsi) * (qcj + qsj) * iR;
                const double wa = -qsi * qcj * iA * iA * iA, wc = -qsi * qsj * iC * iC * iC;
                gx += wa * Ax + wc * Cx; gy += wa * Ay + wc * Cy; gz += wa * Az + wc * Cz;
            } else if (thole_u) {
                const double u = thole_u[(mi * natoms + ai) * natoms + (j % natoms)];
                if (u == 0.0) continue;                            /* S == 0 */
                double gR, gA, gB, gC, cR, cA, cB, cC;
                thole_term(Rx, Ry, Rz, u, &gR, &cR);
                thole_term(Ax, Ay, Az, u, &gA, &cA);
                thole_term(Bx, By, Bz, u, &gB, &cB);
                thole_term(Cx, Cy, Cz, u, &gC, &cC);
                const double qq = qsi * qsj;
                ec += qq * (gR - gA - gB + gC);
                gx += qq * (cC * Cx - cA * Ax); gy += qq * (cC * Cy - cA * Ay); gz += qq * (cC * Cz - cA * Az);
            }
        }
        e_coul += ec; e_stat += es;
        e_self += 0.5 * k[i] * (dxi * dxi + dyi * dyi + dzi * dzi);
        if (grad) {
            grad[3 * i] = K * gx + k[i] * dxi;
            grad[3 * i + 1] = K * gy + k[i] * dyi;
            grad[3 * i + 2] = K * gz + k[i] * dzi;
        }
    }
    energy[1] = K * 0.5 * e_coul;
    energy[2] = K * 0.5 * e_stat;
    energy[3] = e_self;
    energy[0] = (energy[1] - energy[2]) + energy[3];
    return 0;
}
