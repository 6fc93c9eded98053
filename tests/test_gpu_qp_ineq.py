"""GPU parity tests (-m gpu) of the general-inequality QP -- quadprog(G,F,L,c+W*x) exactly as NTM_MPC_Sim.m:97 calls
it, with the state rows of getWLc.m kept.  The strictly convex QP has a unique minimiser, so the CUDA
dual active-set path (csrc/ntm_device.cuh: qp_ineq_continue) is compared with the oracle's direct-solve restatement
(oracle/ntm_oracle.py: qp_ineq) at TOL_U relative to the box width, and certified independently through the KKT
conditions (multipliers by non-negative least squares)."""
import numpy as np
import pytest

from oracle import ntm_oracle as o

pytestmark = pytest.mark.gpu
TOL_U = 1e-6


@pytest.fixture(scope="module")
def mpc():
    import ntm_mpc
    h = ntm_mpc.NtmMpc(0)
    yield h
    h.close()


def _random_problem(rng, n, M):
    B = rng.standard_normal((n, n))
    G = B @ B.T + 0.05 * np.eye(n)
    F = rng.standard_normal(n) * 3
    lb = -rng.random(n); ub = rng.random(n) + 0.1
    Lg = rng.standard_normal((M, n)); bg = rng.standard_normal(M) * 0.7 + 0.2
    return G, F, lb, ub, Lg, bg


@pytest.mark.parametrize("n,M", [(1, 1), (2, 3), (3, 5), (5, 8), (8, 12), (20, 30), (32, 64), (33, 40), (64, 70)])
def test_random_problems_match_oracle_including_infeasible(mpc, n, M):
    rng = np.random.default_rng(100 * n + M)
    S = 96 if n <= 32 else 24
    probs = [_random_problem(rng, n, M) for _ in range(S)]
    G = np.array([p[0] for p in probs]); F = np.array([p[1] for p in probs])
    lb = np.array([p[2] for p in probs]); ub = np.array([p[3] for p in probs])
    Lg = np.array([p[4] for p in probs]); bg = np.array([p[5] for p in probs])
    U, it, st = mpc.qp_ineq(G, F, lb, ub, Lg, bg)
    n_inf = 0
    for s in range(S):
        Uo, _, so = o.qp_ineq(*probs[s])
        assert st[s] == so, (s, st[s], so)
        if so == o.QP_INFEASIBLE:
            n_inf += 1
            continue
        assert np.max(np.abs(U[s] - Uo) / (ub[s] - lb[s])) < TOL_U, s
        stat, viol = o.qp_ineq_kkt_residual(*probs[s], U[s])
        assert stat < 1e-7 and viol < 1e-8, (s, stat, viol)
    assert 0 < n_inf < S or n <= 2 or M >= 30     # the sample exercises both outcomes


def _mpc_problems(cfg, S, N, tighten, seed):
    """(G, F, L, b) of the first QP of a closed-loop step with non-constant scheduling, getWLc rows from the script's
    state box (NTM_MPC_Sim.m:39-45) -- or a tighter lower width bound so that the state rows bind."""
    rng = np.random.default_rng(seed)
    phys, x0, _ = o.make_batch(cfg, S)
    out = []
    for s in range(S):
        p = o.scenario(phys, s); x = x0[s].copy()
        Af, Bf, C = o.model_callables(p)
        xs = [x * (1 + 0.06 * rng.standard_normal(2)) for _ in range(N)]
        R1 = np.array([o.rho1(v, p["w_marg"]) for v in xs]); R2 = np.array([o.rho2(v) for v in xs])
        R3 = np.array([o.rho3(v, p["w_dep"]) for v in xs])
        Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C)
        G, F = o.hessian_grad(Phi, Gam, Lam, x, np.array([0.0, 2000 * np.pi]), np.eye(2))
        xmin = np.array([0.06, 100 * 2 * np.pi]); xmax = np.array([0.15, 5000 * 2 * np.pi])
        if tighten and s % 2:
            xmin[0] = x[0] * 0.9
        W, L, c = o.getWLc(xmax, xmin, 2e6, 0.0, Gam, Phi, Lam)
        out.append((G, F, L, c + W @ x))
    return out


@pytest.mark.parametrize("N", [3, 10, 20, 32, 40])
def test_quadprog_with_getWLc_rows_matches_oracle(mpc, N):
    import ntm_mpc as m
    S = 48 if N <= 20 else 16
    probs = _mpc_problems(3, S, N, True, N)
    G = np.array([p[0] for p in probs]); F = np.array([p[1] for p in probs])
    L = np.array([p[2] for p in probs]); b = np.array([p[3] for p in probs])
    U, fval, flag = m.quadprog(G, F, L, b)
    n_bind = 0
    for s in range(S):
        lb, ub, Lg, bg, feas = o.split_rows(L[s], b[s])
        if not feas:
            assert flag[s] == -2
            continue
        Uo, _, so = o.qp_ineq(G[s], F[s], lb, ub, Lg, bg)
        assert flag[s] == {0: 1, 1: 0, 2: -3, 3: -2}[so], (s, flag[s], so)
        if so != 0:
            continue
        assert np.max(np.abs(U[s] - Uo)) / 2e6 < TOL_U, (s, np.max(np.abs(U[s] - Uo)) / 2e6)
        Ub = o.qp_box(G[s], F[s], lb, ub)[0]
        n_bind += int(np.max(np.abs(Ub - Uo)) / 2e6 > 1e-3)
        fo = 0.5 * Uo @ G[s] @ Uo + F[s] @ Uo
        assert abs(fval[s] - fo) <= 1e-9 * abs(fo) + 1e-6
    assert n_bind >= (1 if N <= 3 else S // 8)     # the state rows change the answer in a good part of the sample


def test_x0_rows_decide_infeasibility_on_the_host_side():
    """Defect D18: the default scenario starts at w = 0 < min_width, so getWLc's x_0 rows (getWLc.m:30) make the QP
    infeasible for every U -- quadprog's exitflag -2 (NTM_MPC_Sim.m:100-101)."""
    import ntm_mpc as m
    p = o.default_physics(); x = o.default_x0()
    Af, Bf, C = o.model_callables(p)
    N = 3
    R1 = np.full(N, o.rho1(x, p["w_marg"])); R2 = np.full(N, o.rho2(x)); R3 = np.full(N, o.rho3(x, p["w_dep"]))
    Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C)
    G, F = o.hessian_grad(Phi, Gam, Lam, x, np.array([0.0, 2000 * np.pi]), np.eye(2))
    W, L, c = o.getWLc([0.15, 5000 * 2 * np.pi], [0.06, 100 * 2 * np.pi], 2e6, 0.0, Gam, Phi, Lam)
    U, fval, flag = m.quadprog(G, F, L, c + W @ x)
    assert flag == -2


def test_no_general_rows_is_the_box_qp_and_bad_arguments_fail(mpc):
    rng = np.random.default_rng(5)
    G, F, lb, ub, Lg, bg = _random_problem(rng, 6, 4)
    U0, _, st0 = mpc.qp_ineq(G, F, lb, ub, np.zeros((0, 6)), np.zeros(0))
    U1, _, st1 = mpc.qp_box(G, F, lb, ub)
    assert st0[0] == st1[0] == 0 and np.array_equal(U0, U1)
    with pytest.raises(ValueError):
        mpc.qp_ineq(G, F, lb, np.full(6, np.inf), Lg, bg)
    import ntm_mpc as m
    with pytest.raises(NotImplementedError):
        m.quadprog(G, F, Lg, bg, np.ones((1, 6)), np.ones(1), lb, ub)
