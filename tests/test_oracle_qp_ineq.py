"""CPU tests of the oracle's general-inequality QP (oracle/ntm_oracle.py: qp_ineq, split_rows) -- the checker of
tests/test_gpu_qp_ineq.py.  The reference has no golden vector at the quadprog boundary ("parity
unpinned"); the oracle is pinned here by what defines the answer: an LP solver for feasibility (quadprog exitflag -2),
the KKT conditions with independently computed multipliers, and exhaustive active-set enumeration on small cases."""
import itertools

import numpy as np
from scipy.optimize import linprog

from oracle import ntm_oracle as o


def random_problem(rng, n, M):
    B = rng.standard_normal((n, n))
    G = B @ B.T + 0.05 * np.eye(n)
    F = rng.standard_normal(n) * 3
    lb = -rng.random(n); ub = rng.random(n) + 0.1
    Lg = rng.standard_normal((M, n)); bg = rng.standard_normal(M) * 0.7 + 0.2
    return G, F, lb, ub, Lg, bg


def enumerate_active_sets(G, F, lb, ub, Lg, bg):
    n = len(F)
    A = np.vstack([-np.eye(n), np.eye(n), Lg]); b = np.concatenate([-lb, ub, bg])
    best = None
    for k in range(n + 1):
        for W in itertools.combinations(range(len(b)), k):
            W = list(W)
            if k:
                Nw = A[W]
                if np.linalg.matrix_rank(Nw) < k:
                    continue
                sol = np.linalg.solve(np.block([[G, Nw.T], [Nw, np.zeros((k, k))]]), np.concatenate([-F, b[W]]))
                x = sol[:n]
                if np.any(sol[n:] < -1e-9):
                    continue
            else:
                x = np.linalg.solve(G, -F)
            if np.all(A @ x <= b + 1e-9 * (1 + np.abs(b))):
                f = 0.5 * x @ G @ x + F @ x
                if best is None or f < best[0]:
                    best = (f, x)
    return best


def test_small_problems_against_exhaustive_enumeration():
    rng = np.random.default_rng(1)
    n_inf = 0
    for _ in range(120):
        n = int(rng.integers(1, 5)); M = int(rng.integers(0, 6))
        pr = random_problem(rng, n, M)
        U, it, st = o.qp_ineq(*pr)
        best = enumerate_active_sets(*pr)
        if best is None:
            n_inf += 1
            assert st == o.QP_INFEASIBLE
        else:
            assert st == o.QP_OK and np.max(np.abs(U - best[1])) < 1e-8
    assert 10 < n_inf < 110


def test_feasibility_matches_lp_and_kkt_certificate_holds():
    rng = np.random.default_rng(7)
    n_inf = 0
    for n, M, cnt in [(2, 3, 40), (5, 8, 60), (6, 10, 60), (8, 12, 40), (20, 30, 20), (40, 50, 4)]:
        for _ in range(cnt):
            pr = random_problem(rng, n, M)
            U, it, st = o.qp_ineq(*pr)
            lp = linprog(np.zeros(n), A_ub=pr[4], b_ub=pr[5], bounds=list(zip(pr[2], pr[3])))
            assert (lp.status == 0) == (st == o.QP_OK), (n, M, st, lp.status)
            if st == o.QP_OK:
                stat, viol = o.qp_ineq_kkt_residual(*pr, U)
                assert stat < 1e-8 and viol < 1e-9
                on_lb, on_ub = U <= pr[2], U >= pr[3]
                assert np.all(U[on_lb] == pr[2][on_lb]) and np.all(U[on_ub] == pr[3][on_ub])   # bounds are exact
            else:
                assert st == o.QP_INFEASIBLE
                n_inf += 1
    assert n_inf > 20


def _mpc_rows(s, N, tighten):
    rng = np.random.default_rng(s)
    phys, x0, _ = o.make_batch(3, s + 1)
    p = o.scenario(phys, s); x = x0[s].copy()
    Af, Bf, C = o.model_callables(p)
    xs = [x * (1 + 0.06 * rng.standard_normal(2)) for _ in range(N)]
    R1 = np.array([o.rho1(v, p["w_marg"]) for v in xs]); R2 = np.array([o.rho2(v) for v in xs])
    R3 = np.array([o.rho3(v, p["w_dep"]) for v in xs])
    Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C)
    G, F = o.hessian_grad(Phi, Gam, Lam, x, np.array([0.0, 2000 * np.pi]), np.eye(2))
    xmin = np.array([x[0] * 0.9 if tighten else 0.06, 100 * 2 * np.pi]); xmax = np.array([0.15, 5000 * 2 * np.pi])
    W, L, c = o.getWLc(xmax, xmin, 2e6, 0.0, Gam, Phi, Lam)
    return G, F, L, c + W @ x, Phi, Gam, Lam, x, xmin, xmax


def test_getWLc_rows_split_into_box_feasibility_and_state_rows():
    N = 10
    G, F, L, b, *_ = _mpc_rows(3, N, False)
    lb, ub, Lg, bg, feas = o.split_rows(L, b)
    assert np.all(lb == 0.0) and np.all(ub == 2e6) and feas          # getWLc.m:14-23
    # 4N state rows, of which the omega_1 pair is empty (B = [b; 0], B.m:2), the w_1 and omega_2 pairs touch u_0 only
    assert Lg.shape == (4 * N - 6, N)
    assert L.shape[0] == 6 * N + 4


def test_state_rows_bind_and_the_predicted_states_respect_them():
    N = 10
    changed = 0
    for s in range(1, 12, 2):
        G, F, L, b, Phi, Gam, Lam, x, xmin, xmax = _mpc_rows(s, N, True)
        lb, ub, Lg, bg, feas = o.split_rows(L, b)
        assert feas
        U, it, st = o.qp_ineq(G, F, lb, ub, Lg, bg)
        if st == o.QP_INFEASIBLE:
            continue
        assert st == o.QP_OK
        X = (Phi @ x + Gam @ U + Lam).reshape(N, 2)
        assert np.all(X >= xmin - 1e-7 * np.abs(xmin)) and np.all(X <= xmax + 1e-7 * np.abs(xmax))
        stat, viol = o.qp_ineq_kkt_residual(G, F, lb, ub, Lg, bg, U)
        assert stat < 1e-8 and viol < 1e-9
        Ub = o.qp_box(G, F, lb, ub)[0]
        changed += int(np.max(np.abs(U - Ub)) > 1e-3 * 2e6)
    assert changed >= 2


def test_default_scenario_is_infeasible_through_the_x0_rows():
    """Defect D18: x0(1) = 0 < min_width, getWLc.m:30 constrains x_0 itself."""
    p = o.default_physics(); x = o.default_x0()
    Af, Bf, C = o.model_callables(p)
    R1 = np.full(3, o.rho1(x, p["w_marg"])); R2 = np.full(3, o.rho2(x)); R3 = np.full(3, o.rho3(x, p["w_dep"]))
    Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C)
    W, L, c = o.getWLc([0.15, 5000 * 2 * np.pi], [0.06, 100 * 2 * np.pi], 2e6, 0.0, Gam, Phi, Lam)
    *_, feas = o.split_rows(L, c + W @ x)
    assert not feas


def test_pinned_variables_and_nonfinite_data():
    rng = np.random.default_rng(3)
    G, F, lb, ub, Lg, bg = random_problem(rng, 5, 3)
    ub2 = ub.copy(); ub2[2] = lb[2]
    bg2 = np.abs(bg) + 5.0
    U, it, st = o.qp_ineq(G, F, lb, ub2, Lg, bg2)
    assert st == o.QP_OK and U[2] == lb[2]
    F2 = F.copy(); F2[0] = np.nan
    assert o.qp_ineq(G, F2, lb, ub, Lg, bg)[2] == o.QP_NONFINITE
