#!/usr/bin/env python
"""State-row QPs inside the loop: warm start (active set of the QP before last) against the cold start (box minimiser +
dual active-set iterations) on the same scenarios -- same unique minimisers, two code paths.
    python tools/check_rows_warm.py [S]        compares; runs itself twice (NTM_ROWS_WARM is read once per process)"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np
BIND = (0.05, 0.16, 2000.0, 12000.0)
SCRIPT = (0.06, 0.15, 200 * np.pi, 10000 * np.pi)


def run(S, out):
    import ntm_mpc
    from ntm_mpc import physics
    mpc = ntm_mpc.NtmMpc(0)
    P, x0, N = physics.batch_params(3, S=S)
    res = {}
    for mode in (1, 2):
        for name, xb in (("bind", BIND), ("script", SCRIPT)):
            g = mpc.closed_loop(x0, P.T, N=N, profile=16, state_rows=mode, xbounds=xb)
            res[f"uk_{mode}_{name}"] = g["uk"]; res[f"st_{mode}_{name}"] = g["status"]; res[f"w_{mode}_{name}"] = g["xk"][:, :, 0]
            res[f"qp_{mode}_{name}"] = g["qp_iters"]
    res["umax"] = P[9]
    np.savez(out, **res)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        run(int(sys.argv[2]), sys.argv[3]); sys.exit(0)
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    tmp = tempfile.mkdtemp()
    outs = {}
    for warm in ("0", "1"):
        out = os.path.join(tmp, f"w{warm}.npz")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--run", str(S), out], env=dict(os.environ, NTM_ROWS_WARM=warm))
        outs[warm] = np.load(out)
    c, w = outs["0"], outs["1"]
    worst = 0
    for mode in (1, 2):
        for name in ("bind", "script"):
            k = f"{mode}_{name}"
            same_st = np.array_equal(c["st_" + k], w["st_" + k])
            nan_eq = np.array_equal(np.isnan(c["uk_" + k]), np.isnan(w["uk_" + k]))
            du = np.nanmax(np.abs(np.nan_to_num(c["uk_" + k]) - np.nan_to_num(w["uk_" + k])), axis=1) / c["umax"]
            nbad = int((du > 1e-6).sum())
            worst = max(worst, nbad)
            print(f"rows mode {mode}, {name} box, S={S}: status identical {same_st}, NaN pattern identical {nan_eq}, infeasible {int((c['st_' + k] == 3).sum())}, "
                  f"scenarios with |du| > 1e-6 umax: {nbad} (max {du.max():.2e}), dual iterations cold {int(c['qp_' + k].sum())} warm {int(w['qp_' + k].sum())}")
            assert same_st and nan_eq
    sys.exit(0 if worst <= max(1, S // 1000) else 1)
