#!/usr/bin/env python
"""Two-phase launch with a longest-first work queue against the single launch in natural order: bit-identical outputs.
    python tools/check_lpt.py [S]       runs itself twice (NTM_LPT is read once per process)"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np


def run(S, out):
    import ntm_mpc
    from ntm_mpc import physics
    mpc = ntm_mpc.NtmMpc(0)
    res = {}
    for cfg, prof in ((3, 16), (3, 0), (4, 16)):
        P, x0, N = physics.batch_params(cfg, S=S)
        g = mpc.closed_loop(x0, P.T, N=N, profile=prof, want_Uk=True)
        for k in ("xk", "uk", "Uk", "cost", "inner_iters", "qp_iters", "status"):
            res[f"{k}_{cfg}_{prof}"] = g[k]
    res["launches"] = np.array([mpc.launch_count()])
    np.savez(out, **res)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        run(int(sys.argv[2]), sys.argv[3]); sys.exit(0)
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    tmp = tempfile.mkdtemp()
    outs = {}
    for lpt in ("0", "1"):
        out = os.path.join(tmp, f"l{lpt}.npz")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--run", str(S), out], env=dict(os.environ, NTM_LPT=lpt))
        outs[lpt] = np.load(out)
    a, b = outs["0"], outs["1"]
    ok = True
    for k in a.files:
        if k == "launches":
            continue
        same = np.array_equal(a[k], b[k], equal_nan=True)
        ok &= same
        if not same:
            print("DIFFERENT:", k)
    print(f"S={S}: kernels launched natural order {int(a['launches'][0])}, two-phase {int(b['launches'][0])}; all outputs bit-identical: {ok}")
    sys.exit(0 if ok and int(b["launches"][0]) > int(a["launches"][0]) else 1)
