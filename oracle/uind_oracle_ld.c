/*
 * TEST INFRASTRUCTURE - arbiter: U_ind with every operation in 80-bit long double (64-bit mantissa),
 * same per-ordered-pair formulas as oracle/uind_oracle.c (U_pol/polarization.py:50-109,:149).
 *
 * Why it exists: U_ind = U_coul - U_coul_static + U_self is a small remainder of large sums, and at
 * BASELINE sizes the fp64 restatement (uind_oracle.c, like the reference's own fp64 JAX path) is
 * itself only good to ~1e-10 relative (measured against this arbiter: 1.5e-11 at 9 k sites,
 * 4.0e-11 at 36 k, 6.4e-11 at 72 k, growing with N), i.e. at the parity bar.  The CUDA path stays
 * within 2e-12 of the arbiter at those sizes; tests/test_gpu_parity.py checks that.
 * ~4x slower than the fp64 oracle; energy only.
 */
#include <math.h>
#include <stdint.h>
typedef long double R;
static inline R invn(R x, R y, R z, R *r) { *r = sqrtl(x*x + y*y + z*z); return *r == 0 ? 0 : 1 / *r; }
static inline R gterm(R x, R y, R z, R u) { R r; R ir = invn(x, y, z, &r); if (r == 0) return 0; return (1 - (1 + 0.5L*r*u)*expl(-u*r))*ir; }
double upol_oracle_uind_long_double(int64_t n, int na, const double *r, const double *d, const double *qc, const double *qs,
                    const double *k, const double *th) {
    const R E = 1.602176634e-19L, NA = 6.02214076e23L;
    /* same literals as polarization.py:19-25 but evaluated in long double would change K itself: use the double value */
    const double eps0 = 1e-6 * 8.8541878128e-12 / (1.602176634e-19 * 1.602176634e-19 * 6.02214076e23);
    const R K = (R)(1 / (4 * 3.14159265358979323846 * eps0));
    (void)E; (void)NA;
    R ec = 0, es = 0, eself = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ec, es, eself)
    for (int64_t i = 0; i < n; ++i) {
        R xi = r[3*i], yi = r[3*i+1], zi = r[3*i+2], dxi = d[3*i], dyi = d[3*i+1], dzi = d[3*i+2];
        R qci = qc[i], qsi = qs[i], c = 0, s = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            R Rx = r[3*j]-xi, Ry = r[3*j+1]-yi, Rz = r[3*j+2]-zi;
            R Ax = Rx+dxi, Ay = Ry+dyi, Az = Rz+dzi;
            R Bx = Rx-d[3*j], By = Ry-d[3*j+1], Bz = Rz-d[3*j+2];
            R Cx = Ax-d[3*j], Cy = Ay-d[3*j+1], Cz = Az-d[3*j+2];
            R qcj = qc[j], qsj = qs[j], t;
            if (j / na != i / na) {
                R iR = invn(Rx,Ry,Rz,&t), iA = invn(Ax,Ay,Az,&t), iB = invn(Bx,By,Bz,&t), iC = invn(Cx,Cy,Cz,&t);
                c += qci*qcj*iR + qsi*qcj*iA + qci*qsj*iB + qsi*qsj*iC;
                s += (qci+qsi)*(qcj+qsj)*iR;
            } else if (th) {
                R u = th[((i/na)*na + i%na)*na + j%na];
                if (u == 0) continue;
                c += qsi*qsj*(gterm(Rx,Ry,Rz,u) - gterm(Ax,Ay,Az,u) - gterm(Bx,By,Bz,u) + gterm(Cx,Cy,Cz,u));
            }
        }
        ec += c; es += s; eself += 0.5L*k[i]*(dxi*dxi+dyi*dyi+dzi*dzi);
    }
    return (double)((K*0.5L*ec - K*0.5L*es) + eself);
}
