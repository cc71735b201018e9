/*
 * Plain-C use of the C ABI (include/upol_b200.h): a water dimer, energy + gradient + minimisation
 * through host buffers.  Build:
 *   gcc -std=c99 -Iinclude examples/c_api_example.c -Lu_pol_b200/lib -lupol_b200 -Wl,-rpath,$PWD/u_pol_b200/lib -o c_api_example
 * Without a CUDA device it prints the constant and exits (there is no CPU fallback).
 */
#include <stdio.h>
#include <string.h>

#include "upol_b200.h"

int main(void) {
    printf("libupol_b200 version %d, ONE_4PI_EPS0 = %.14f kJ mol^-1 nm e^-2\n", upol_version(), upol_one_4pi_eps0());
    if (upol_device_count() == 0) {
        printf("no CUDA device: compute entry points are unavailable (no CPU fallback)\n");
        return 0;
    }
    /* two SWM4-NDP-style waters (O, H1, H2, M), 0.3 nm apart; U_pol/benchmarks/OpenMM/water/water.xml:26-32 */
    const double K = upol_one_4pi_eps0();
    const double alpha = 0.000978253, qd = -1.71636;
    double r[24] = {0.0, 0.0, 0.0, 0.0757, 0.0, 0.0586, -0.0757, 0.0, 0.0586, 0.0, 0.0, 0.0240,
                    0.30, 0.0, 0.0, 0.3757, 0.0, 0.0586, 0.2243, 0.0, 0.0586, 0.30, 0.0, 0.0240};
    double qc[8] = {1.71636, 0.55733, 0.55733, -1.11466, 1.71636, 0.55733, 0.55733, -1.11466};
    double qs[8] = {qd, 0, 0, 0, qd, 0, 0, 0};
    double k[8] = {0};
    int32_t mol[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    double d[24] = {0}, e[UPOL_E_SIZE], g[24];
    upol_system *sys = NULL;
    upol_min_options opt;
    upol_min_result res;
    int rc;
    k[0] = k[4] = K * qd * qd / alpha;                      /* util.py:267-268 */
    rc = upol_system_create(&sys, 0, 8, r, qc, qs, k, mol, 0, NULL, NULL, NULL);
    if (rc) { printf("create failed: %s\n", upol_error_string(rc)); return 1; }
    rc = upol_uind_energy_grad_host(sys, d, UPOL_FP64, e, g);
    if (rc) { printf("energy failed: %s\n", upol_error_string(rc)); return 1; }
    printf("U_ind(d=0) = %.3e kJ/mol, dU/dd(O1) = (%.4f, %.4f, %.4f) kJ/mol/nm\n", e[UPOL_E_UIND], g[0], g[1], g[2]);
    upol_min_options_default(&opt);
    rc = upol_minimize_host(sys, d, &opt, &res);
    if (rc < 0) { printf("minimise failed: %s\n", upol_error_string(rc)); return 1; }
    printf("minimised U_ind = %.9f kJ/mol in %d iterations, |grad| = %.2e, d(O1) = (%.6f, %.6f, %.6f) nm\n",
           res.energy[UPOL_E_UIND], res.iterations, res.gnorm, d[0], d[1], d[2]);
    upol_system_destroy(sys);
    return 0;
}
