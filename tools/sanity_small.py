"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck): all group widths, both Hessian
builds, the materialising entry points and the MEX-facing host path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np
import ntm_mpc
from ntm_mpc import physics

mpc = ntm_mpc.NtmMpc(0)
for cfg, S, N, ks in ((3, 24, 20, 3), (2, 8, 10, 2), (3, 6, 40, 2), (5, 4, 100, 2)):
    P, x0, _ = physics.batch_params(cfg, S=S)
    for prof in (16, 0, ntm_mpc.PROFILE_CONSISTENT | 16, 16 | ntm_mpc.PROFILE_DENSE_G):
        r = mpc.closed_loop(x0, P.T, N=N, k_sim=ks, i_sim=3, profile=prof, want_Uk=True)
        assert r["status"].max() == 0, (cfg, prof, r["status"])
    R = np.abs(np.random.default_rng(N).standard_normal((3, S, N))) * 1e-3 + 1e-3
    for gi in (0, 2):
        Phi, Gam, Lam = mpc.condense(R[0], R[1], R[2], P.T, profile=gi)
    G, F = mpc.hessian_grad(Phi, Gam, Lam, x0, P.T)
    U, it, st = mpc.qp_box(G, F, 0.0, 2e6)
    assert st.max() == 0
    r1, r2, r3 = mpc.rho(x0, P.T); A, B = mpc.lpv_AB(r1, r2, r3, P.T); xn = mpc.plant_step(x0, U[:, 0], P.T)
# EXT instantiations of the fused loop (RK4 plant, state rows refresh / frozen), the Monte-Carlo reduction in both
# layouts' host path, the quadprog shim (general-inequality QP with factor rebuilds)
XB = (0.05, 0.16, 2000.0, 12000.0)
for N, S in ((10, 12), (20, 12), (40, 4), (72, 3)):
    P, x0, _ = physics.batch_params(3, S=S)
    r = mpc.closed_loop(x0, P.T, N=N, k_sim=3, i_sim=2, profile=16 | ntm_mpc.PROFILE_PLANT_RK4)
    for mode in (ntm_mpc.STATE_ROWS_REFRESH, ntm_mpc.STATE_ROWS_FROZEN):
        r = mpc.closed_loop(x0, P.T, N=N, k_sim=3, i_sim=2, profile=16, state_rows=mode, xbounds=XB, want_Uk=True)
        assert set(np.unique(r["status"])) <= {0, 1, 3}, r["status"]
    full = mpc.closed_loop(x0, P.T, N=N, k_sim=3, i_sim=2, profile=16)
    stats = mpc.mc_stats(full["xk"], full["uk"], full["cost"], full["status"], np.ascontiguousarray(P.T))
    assert stats[0] + stats[1] + stats[2] + stats[3] == S
rng = np.random.default_rng(0)
for N in (5, 48, 100):
    M = rng.standard_normal((6, 2 * N, N)); G = 2 * np.einsum("ski,skj->sij", M, M); F = rng.standard_normal((6, N))
    U, it, st = mpc.qp_box(G, F, -0.05, 0.05)
    assert st.max() == 0
print("sanity ok, launches", mpc.launch_count())
