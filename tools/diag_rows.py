"""Diagnostic (GPU box): first QP at which the fused loop with state rows and the oracle disagree, replayed through
the per-call kernel ntm_qp_ineq."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
from oracle import ntm_oracle as o
import ntm_mpc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 72
s = int(sys.argv[2]) if len(sys.argv) > 2 else 0
XB = (0.05, 0.16, 2000.0, 12000.0)
h = ntm_mpc.NtmMpc(0)
phys, x0, _ = o.make_batch(3, S=6)
P = np.ascontiguousarray(o.derive_params_batch(phys).T)
g = h.closed_loop(x0, P, N=N, k_sim=4, i_sim=2, profile=o.LITERAL_FIXED.flags(), state_rows=1, xbounds=XB)
print("gpu status", g["status"], "inner", g["inner_iters"][s], "qp", g["qp_iters"][s])
log = []
r = o.closed_loop(o.scenario(phys, s), x0[s], N=N, k_sim=4, i_sim=2, profile=o.LITERAL_FIXED, state_rows=1, xbounds=XB, qp_log=log)
print("oracle status", r["status"], "qp", r["qp_iters"])
p = o.scenario(phys, s)
xmin = np.array([XB[0], XB[2]]); xmax = np.array([XB[1], XB[3]])
for e in log:
    W, L, c = o.getWLc(xmax, xmin, [p["umax"]], [p["umin"]], e["rows"][1], e["rows"][0], e["rows"][2])
    b = c + W @ e["xk"]
    lo, hi, Lg, bg, feas = o.split_rows(L, b)
    Uo, ito, so = o.qp_ineq(e["G"], e["F"], lo, hi, Lg, bg)
    # (a) rows split as the quadprog shim does, (b) every state row as a general row with the plain box (what the loop does)
    Ua, ita, sa = h.qp_ineq(e["G"][None], e["F"][None], lo[None], hi[None], Lg[None], bg[None])
    gen = [i for i in range(L.shape[0]) if not (np.count_nonzero(L[i]) == 1 and abs(abs(L[i]).max() - 1.0) == 0.0 and i % 6 < 2)]
    Lb = L[gen]; bb = b[gen]
    Ub, itb, sb = h.qp_ineq(e["G"][None], e["F"][None], np.full((1, N), p["umin"]), np.full((1, N), p["umax"]), Lb[None], bb[None])
    act = int(np.sum(np.abs(Lg @ Uo - bg) <= 1e-9 * (np.abs(Lg) @ (hi - lo)))) if so == 0 else -1
    nb = int(np.sum((Uo <= lo) | (Uo >= hi))) if so == 0 else -1
    print(f"k={e['k']} it={e['it']} oracle st {so} it {ito} tight rows {act} bound vars {nb} | gpu split st {int(sa[0])} it {int(ita[0])} err {np.max(np.abs(Ua[0]-Uo)) if sa[0]==0 and so==0 else -1:.2e} | gpu all-general st {int(sb[0])} it {int(itb[0])} err {np.max(np.abs(Ub[0]-Uo)) if sb[0]==0 and so==0 else -1:.2e}")
