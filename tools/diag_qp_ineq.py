"""Quick diagnostic of ntm_qp_ineq against the oracle (run on the GPU box)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ntm_oracle as o
import ntm_mpc
from test_gpu_qp_ineq import _random_problem, _mpc_problems
h = ntm_mpc.NtmMpc(0)
rng = np.random.default_rng(0)
for n, M in [(2, 3), (5, 8), (20, 30)]:
    probs = [_random_problem(rng, n, M) for _ in range(64)]
    arr = [np.array([p[i] for p in probs]) for i in range(6)]
    U, it, st = h.qp_ineq(*arr)
    bad = 0
    for s, p in enumerate(probs):
        Uo, ito, so = o.qp_ineq(*p)
        e = np.max(np.abs(U[s] - Uo)) if so == 0 and st[s] == 0 else 0
        if st[s] != so or e > 1e-6:
            bad += 1
            if bad < 4: print("  mismatch", n, M, s, "gpu st", st[s], "it", it[s], "oracle st", so, "it", ito, "err", e)
    print("random", n, M, "bad", bad, "of 64; mean gpu iters", it.mean())
for N in (3, 10, 20, 32, 40):
    S = 32
    probs = _mpc_problems(3, S, N, True, N)
    G = np.array([p[0] for p in probs]); F = np.array([p[1] for p in probs]); L = np.array([p[2] for p in probs]); b = np.array([p[3] for p in probs])
    t0 = time.time(); U, fval, flag = ntm_mpc.quadprog(G, F, L, b); dt = time.time() - t0
    errs = []
    for s in range(S):
        lb, ub, Lg, bg, feas = o.split_rows(L[s], b[s])
        Uo, ito, so = o.qp_ineq(G[s], F[s], lb, ub, Lg, bg)
        errs.append((np.max(np.abs(U[s] - Uo)) / 2e6 if so == 0 else 0.0, int(flag[s]), so))
    print("mpc N", N, "max err %.2e" % max(e[0] for e in errs), "flags", sorted(set((e[1], e[2]) for e in errs)), "%.1f ms" % (dt * 1e3))
