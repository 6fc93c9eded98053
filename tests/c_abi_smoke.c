/* Plain-C host of the drop-in boundary: the default scenario of NTM_MPC_Sim.m (x0 = [0; 2000*pi], N = 3, k_sim = 20)
 * through ntm_mpc_closed_loop with HOST pointers.  Prints uk and the island widths; exit code 0 = ran on the GPU,
 * 3 = no usable CUDA device (expected in the CPU-only build container), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include "ntm_mpc.h"

int main(void) {
    const double pi = 3.141592653589793;
    /* NTM_MPC_Sim.m:5-25,31,37,47-50,59-60 -> the 16-double parameter block of ntm_mpc.h */
    const double j_BS = 73e3, w_dep = 0.024, w_marg = 0.02, w_sat = 0.32, tau_r = 293, rs = 1.55, a = 2.0, eta_CD = 0.9,
                 tau_E0 = 3.7, mu0 = 4e-7 * pi, Lq = 0.87, B_pol = 0.97, m = 2, Cw = 1, tau_A0 = 3e-6, tau_w = 0.188,
                 omega0 = 2 * pi * 420, Ts = 0.1;
    const double kappa = 16 * mu0 * Lq * rs * rs / (0.82 * tau_r * B_pol * pi);
    const double zeta = m * Cw * tau_A0 * tau_A0 * tau_w * a * a * a;
    double prm[NTM_NPARAM] = {
        (4.0 / 3.0) * (kappa * rs / (0.82 * tau_r)) * Ts, Ts / (zeta * a * a * a), 1 - Ts / tau_E0, kappa * Ts * eta_CD / w_dep,
        -4.0 / 3.0 * (kappa * Ts * j_BS * w_sat) / (w_sat * w_sat + w_marg * w_marg), Ts * omega0 / tau_E0,
        w_marg * w_marg, w_dep, 0.0, 2e6, 0.0, 1000 * 2 * pi, 1.0, 0.0, 1.0, 0.0};
    double x0[2] = {0.0, 1000 * 2 * pi}, xk[2 * 21], uk[20], cost;
    int inner[20], qp[20], status, k;
    ntm_handle *h = NULL;
    if (ntm_create(&h, 0) != NTM_OK) { fprintf(stderr, "ntm_create: %s\n", ntm_last_error()); return 3; }
    if (ntm_mpc_closed_loop(h, NTM_LAYOUT_MATLAB, NTM_PROFILE_INNER_FIXED, 1, 3, 20, 10, 1e-14, x0, prm, 1, xk, uk, NULL, &cost,
                            inner, qp, &status) != NTM_OK) {
        fprintf(stderr, "ntm_mpc_closed_loop: %s\n", ntm_last_error());
        return 1;
    }
    printf("status %d cost %.17g\n", status, cost);
    for (k = 0; k < 20; ++k) printf("k %d u %.17g w %.17g inner %d\n", k, uk[k], xk[2 * (k + 1)], inner[k]);
    ntm_destroy(h);
    return status == 0 ? 0 : 2;
}
