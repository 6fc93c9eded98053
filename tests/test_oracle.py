"""Pins the oracle (oracle/ntm_oracle.py and oracle/ntm_oracle.c) -- CPU only.

Order of trust: closed-form known answers (appendix_a.json) -> 50-digit mpmath twin ->
solver-independent KKT certificates + SciPy BVLS -> structural properties -> NumPy-vs-C agreement ->
committed golden closed-loop fixtures.
"""
import json
import math
import os

import mpmath as mp
import numpy as np
import pytest
from scipy.optimize import lsq_linear

from oracle import c_oracle as co
from oracle import ntm_oracle as o

GOLD = os.path.join(os.path.dirname(__file__), "golden")
APPX = json.load(open(os.path.join(GOLD, "appendix_a.json")))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ------------------------------------------------------------------ closed-form known answers
def test_constants_appendix_a():
    p = o.default_physics()
    assert o.kappa_of(p) == pytest.approx(APPX["kappa"], rel=2e-15)
    assert o.zeta_of(p) == pytest.approx(APPX["zeta"], rel=2e-15)
    assert rel(o.C_of(p), APPX["C"]) < 2e-15
    assert o.derive_params(p)[2] == pytest.approx(APPX["a22"], rel=2e-15)


@pytest.mark.parametrize("st", APPX["states"])
def test_rho_A_B_appendix_a(st):
    p = o.default_physics()
    x = np.array(st["x"])
    Af, Bf, _ = o.model_callables(p)
    r1, r2, r3 = o.rho1(x, p["w_marg"]), o.rho2(x), o.rho3(x, p["w_dep"])
    assert r1 == pytest.approx(st["rho1"], rel=1e-14)
    assert o.rho1(x, p["w_marg"], o.RHO1_SQ) == pytest.approx(st["rho1_sq"], rel=1e-14)
    assert r2 == pytest.approx(st["rho2"], rel=1e-14, abs=0)
    assert r3 == pytest.approx(st["rho3"], rel=1e-14)
    A, B = Af(r1, r2), Bf(r3)
    assert A[0, 0] == pytest.approx(st["a11"], rel=1e-15)
    assert A[1, 0] == pytest.approx(st["a21"], rel=1e-13, abs=0)
    assert A[0, 1] == 0.0 and B[1] == 0.0
    assert B[0] == pytest.approx(st["b1"], rel=1e-14)
    # hoisted coefficients (the device parameter block) reproduce A.m / B.m
    prm = o.derive_params(p)
    assert prm[0] * r1 + 1 == A[0, 0]
    assert prm[3] * r3 == B[0]
    assert prm[1] * r2 == pytest.approx(A[1, 0], rel=4e-16, abs=0)


def test_default_condensation_appendix_a():
    p = o.default_physics()
    x0 = o.default_x0()
    Af, Bf, C = o.model_callables(p)
    N = 3
    R1 = np.full(N, o.rho1(x0, p["w_marg"])); R2 = np.full(N, o.rho2(x0)); R3 = np.full(N, o.rho3(x0, p["w_dep"]))
    d = APPX["default_N3"]
    for gi in (o.GAMMA_I_MINUS_J, o.GAMMA_I):            # rho constant -> literal == consistent
        Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C, gi)
        for j in range(N):
            blk = Phi[2 * j:2 * j + 2]
            assert blk[0, 0] == pytest.approx(d["a11_pow"][j], rel=1e-15)
            assert blk[1, 1] == pytest.approx(d["a22_pow"][j], rel=1e-15)
            assert blk[0, 1] == 0 and blk[1, 0] == 0
        b = APPX["states"][0]["b1"]
        expect = np.array([[b, 0, 0], [d["a11_b"], b, 0], [d["a11sq_b"], d["a11_b"], b]])
        assert rel(Gam[0::2], expect) < 1e-14
        assert np.all(Gam[1::2] == 0)
        assert rel(Lam, d["Lambda"]) < 1e-14
        G, F = o.hessian_grad(Phi, Gam, Lam, x0, [p["r1"], p["r2"]], np.eye(2))
        Gu = np.array([[v if v is not None else G[i, j] for j, v in enumerate(row)] for i, row in enumerate(d["G_upper"])])
        assert rel(np.triu(G), np.triu(Gu)) < 1e-13
        assert rel(F, d["F"]) < 1e-12
        U, _, st = o.qp_box(G, F, p["umin"], p["umax"])
        assert st == 0 and np.allclose(U, d["U_first"], rtol=1e-8)


def test_first_qp_saturates_at_x008():
    p = o.default_physics()
    r = o.closed_loop(p, [0.08, 2000 * math.pi], N=3, k_sim=1, i_sim=1)
    assert np.array_equal(r["Uk"][:, 0], APPX["default_N3"]["U_first_x008"])


# ------------------------------------------------------------------ 50-digit twin
def _mp_model(p):
    mp.mp.dps = 50
    P = {k: mp.mpf(repr(v)) for k, v in p.items()}       # exact decimal of the fp64 literal
    P["mu0"] = mp.mpf(p["mu0"]); P["omega0"] = mp.mpf(p["omega0"]); P["r2"] = mp.mpf(p["r2"])
    kappa = 16 * P["mu0"] * P["Lq"] * P["rs"] ** 2 / (mp.mpf("0.82") * P["tau_r"] * P["B_pol"] * mp.mpf(math.pi))
    zeta = P["m"] * P["Cw"] * P["tau_A0"] ** 2 * P["tau_w"] * P["a"] ** 3
    C = [-mp.mpf(4) / 3 * (kappa * P["Ts"] * P["j_BS"] * P["w_sat"]) / (P["w_sat"] ** 2 + P["w_marg"] ** 2),
         P["Ts"] * P["omega0"] / P["tau_E0"]]

    def A(r1, r2):
        return mp.matrix([[mp.mpf(4) / 3 * (kappa * P["rs"] / (mp.mpf("0.82") * P["tau_r"])) * P["Ts"] * r1 + 1, 0],
                          [r2 * P["Ts"] / (zeta * P["a"] ** 3), 1 - P["Ts"] / P["tau_E0"]]])

    def B(r3):
        return mp.matrix([[kappa * P["Ts"] * P["eta_CD"] / P["w_dep"] * r3], [0]])

    return P, kappa, zeta, C, A, B


def test_mpmath_twin_constants_and_condensation():
    p = o.default_physics()
    P, kappa, zeta, C, A, B = _mp_model(p)
    assert abs(o.kappa_of(p) / kappa - 1) < 5e-16
    assert abs(o.zeta_of(p) / zeta - 1) < 5e-16
    Cn = o.C_of(p)
    assert abs(Cn[0] / C[0] - 1) < 1e-15 and abs(Cn[1] / C[1] - 1) < 1e-15
    rng = np.random.default_rng(7)
    Af, Bf, Cf = o.model_callables(p)
    for N in (3, 10):
        w = rng.uniform(0.06, 0.15, N); om = rng.uniform(600, 12000, N)
        R1 = np.array([o.rho1([w[i], om[i]], p["w_marg"]) for i in range(N)])
        R2 = np.array([o.rho2([w[i], om[i]]) for i in range(N)])
        R3 = np.array([o.rho3([w[i], om[i]], p["w_dep"]) for i in range(N)])
        # rho functions against the twin
        for i in range(N):
            ws = mp.mpf(w[i]) / P["w_dep"]
            r3 = (mp.mpf("0.25") + mp.mpf("0.24") * ws) / (1 + mp.mpf("1.5") * ws + mp.mpf("0.43") * ws ** 2 + mp.mpf("0.64") * ws ** 3)
            assert abs(R3[i] / r3 - 1) < 2e-15
            assert abs(R1[i] / (1 / (mp.mpf(w[i]) + P["w_marg"] ** 2)) - 1) < 1e-15
            assert abs(R2[i] / (mp.mpf(w[i]) ** 2 / mp.mpf(om[i])) - 1) < 1e-15
        for gi in (o.GAMMA_I_MINUS_J, o.GAMMA_I):
            Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, Cf, gi)
            mA = [A(mp.mpf(R1[i]), mp.mpf(R2[i])) for i in range(N)]
            mB = [B(mp.mpf(R3[i])) for i in range(N)]
            mPhi = [mA[0]]
            for j in range(1, N):
                mPhi.append(mA[j] * mPhi[-1])
            mLam = [mp.matrix(C)]
            for i in range(1, N):
                mLam.append(mA[i] * mLam[-1] + mp.matrix(C))
            mG = {}
            for i in range(N):
                for j in range(i + 1):
                    if i == j:
                        mG[i, j] = mB[j]
                    else:
                        k = (i - j - 1) if gi == o.GAMMA_I_MINUS_J else i     # 0-based A index
                        mG[i, j] = mA[k] * mG[i - 1, j]
            scale_phi = max(abs(float(mPhi[j][r, c])) for j in range(N) for r in range(2) for c in range(2))
            scale_gam = max(abs(float(v[r])) for v in mG.values() for r in range(2))
            scale_lam = max(abs(float(v[r])) for v in mLam for r in range(2))
            for j in range(N):
                for r in range(2):
                    for c in range(2):
                        assert abs(Phi[2 * j + r, c] - float(mPhi[j][r, c])) <= 1e-13 * scale_phi
                    assert abs(Lam[2 * j + r] - float(mLam[j][r])) <= 1e-13 * scale_lam
            for (i, j), v in mG.items():
                for r in range(2):
                    assert abs(Gam[2 * i + r, j] - float(v[r])) <= 1e-13 * scale_gam


# ------------------------------------------------------------------ structural properties (document D6)
def _random_rho(p, N, rng):
    w = rng.uniform(0.06, 0.15, N); om = rng.uniform(600, 12000, N)
    X = np.stack([w, om], axis=1)
    return (np.array([o.rho1(x, p["w_marg"]) for x in X]), np.array([o.rho2(x) for x in X]),
            np.array([o.rho3(x, p["w_dep"]) for x in X]))


@pytest.mark.parametrize("N", [1, 2, 3, 10, 20])
def test_consistent_condensation_reproduces_rollout_and_literal_does_not(N):
    p = o.default_physics()
    rng = np.random.default_rng(N)
    Af, Bf, C = o.model_callables(p)
    R1, R2, R3 = _random_rho(p, N, rng)
    U = rng.uniform(0, 2e6, N)
    x = np.array([0.09, 5000.0])
    xs, X = x.copy(), []
    for i in range(N):                                            # NTM_MPC_Sim.m:113
        xs = Af(R1[i], R2[i]) @ xs + Bf(R3[i]) * U[i] + C
        X.append(xs)
    X = np.concatenate(X)
    Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C, o.GAMMA_I)
    assert rel(Phi @ x + Gam @ U + Lam, X) < 1e-12
    Phi2, Gam2, Lam2 = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C, o.GAMMA_I_MINUS_J)
    assert np.array_equal(Phi, Phi2) and np.array_equal(Lam, Lam2)
    if N >= 3:
        assert rel(Gam2, Gam) > 1e-3                              # D6: the literal index is a different matrix
    # block lower-triangular, diagonal blocks B_j
    for j in range(N):
        assert np.all(Gam2[:2 * j, j] == 0)
        assert np.array_equal(Gam2[2 * j:2 * j + 2, j], Bf(R3[j]))


# ------------------------------------------------------------------ QP certificates
def _random_qp(N, rng, cond_pow=6):
    M = rng.standard_normal((2 * N, N)) * np.logspace(0, -cond_pow / 2, N)[None, :]
    G = 2 * M.T @ M
    F = rng.standard_normal(N) * np.abs(G).max() * rng.uniform(0.1, 3)
    return M, G, F


@pytest.mark.parametrize("N", [1, 2, 3, 10, 20, 50])
@pytest.mark.parametrize("solver", ["numpy", "c"])
def test_qp_box_kkt_and_bvls(N, solver):
    rng = np.random.default_rng(100 + N)
    for trial in range(6):
        M, G, F = _random_qp(N, rng, cond_pow=rng.integers(0, 8))
        lb = np.full(N, -rng.uniform(0.1, 2)); ub = np.full(N, rng.uniform(0.1, 2))
        U, it, st = (o.qp_box if solver == "numpy" else co.qp_box)(G, F, lb, ub)
        assert st == 0
        assert np.all(U >= lb) and np.all(U <= ub)
        assert o.qp_kkt_residual(G, F, lb, ub, U) < 1e-10
        # second opinion: BVLS on min |M U - d|^2 with 2 M'M = G, -2 M'd = F
        d = np.linalg.lstsq(M.T, -F / 2, rcond=None)[0]
        ref = lsq_linear(M, d, bounds=(lb, ub), method="bvls", tol=1e-14).x
        obj = lambda u: 0.5 * u @ G @ u + F @ u
        assert obj(U) <= obj(ref) + 1e-9 * (abs(obj(ref)) + 1e-300)


def test_qp_box_edge_cases():
    G = np.array([[2.0]]); F = np.array([-2.0])
    assert o.qp_box(G, F, 0, 5)[0][0] == pytest.approx(1.0)
    assert o.qp_box(G, F, 2, 5)[0][0] == 2.0                      # exact lower bound
    assert o.qp_box(G, F, -3, 0.5)[0][0] == 0.5                   # exact upper bound
    U, _, st = o.qp_box(np.array([[np.nan]]), F, 0, 1)
    assert st == 2 and np.isnan(U[0])
    U, _, st = co.qp_box(np.array([[np.nan]]), F, 0, 1)
    assert st == 2 and np.isnan(U[0])


# ------------------------------------------------------------------ C restatement vs NumPy oracle
@pytest.mark.parametrize("N", [1, 3, 10, 20, 100])
@pytest.mark.parametrize("gi", [0, 1])
def test_c_condense_and_hessian_match_numpy(N, gi):
    p = o.default_physics()
    rng = np.random.default_rng(N * 2 + gi)
    Af, Bf, C = o.model_callables(p)
    R1, R2, R3 = _random_rho(p, N, rng)
    Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C, gi)
    flags = o.Profile(gamma_index=gi).flags()
    cPhi, cGam, cLam = co.condense({k: np.array([v]) for k, v in p.items()}, R1, R2, R3, flags)
    assert rel(cPhi, Phi) < 1e-13 and rel(cGam, Gam) < 1e-13 and rel(cLam, Lam) < 1e-13
    x = np.array([0.1, 4000.0]); r = [p["r1"], p["r2"]]; Q = np.array([[1.0, 0.2], [0.2, 3.0]])
    G, F = o.hessian_grad(Phi, Gam, Lam, x, r, Q)
    cG, cF = co.hessian_grad(Phi, Gam, Lam, x, r, Q)
    assert rel(cG, G) < 1e-12 and rel(cF, F) < 1e-11
    assert np.array_equal(cG, cG.T)


@pytest.mark.parametrize("cfg,S", [(1, 1), (2, 8), (3, 8), (4, 8)])
@pytest.mark.parametrize("prof", [o.LITERAL_FIXED, o.CONSISTENT_FIXED, o.LITERAL], ids=["lit_fixed", "con_fixed", "lit_eps"])
def test_c_closed_loop_matches_numpy(cfg, S, prof):
    if cfg == 1 and prof.inner_policy == o.INNER_EPS_BREAK:
        pytest.skip("default scenario + eps_break is bit-chaotic by construction (SURVEY D14); covered by the fixed policy")
    phys, x0, N = o.make_batch(cfg, S=S)
    rc = co.closed_loop_batch(phys, x0, N, flags=prof.flags(), want_Uk=True)
    for s in range(S):
        r = o.closed_loop(o.scenario(phys, s), x0[s], N=N, profile=prof)
        umax = phys["umax"][s]
        assert np.max(np.abs(r["uk"] - rc["uk"][s])) <= 1e-6 * umax
        wref = r["xk"][0]
        assert np.max(np.abs(wref - rc["xk"][s, :, 0])) <= 1e-6 * max(np.max(np.abs(wref)), 1e-3)
        assert np.max(np.abs(r["xk"][1] - rc["xk"][s, :, 1])) <= 1e-6 * np.max(np.abs(r["xk"][1]))
        assert rc["cost"][s] == pytest.approx(r["cost"], rel=1e-6)
        assert rc["status"][s] == r["status"] == 0


def test_the_two_oracles_agree_on_the_heaviest_scenarios_of_config3():
    """Scenarios 62,420 / 60,466 / 60,091 of config 3 are the hardest the batch holds (10 bang-bang switches per QP, or ~10
    free variables in every QP: they end the slowest 8-GPU shard).  The NumPy oracle and its C restatement are two
    independent active-set implementations; on these they must still agree to rounding over the whole closed loop (the GPU
    side of the same statement is tests/test_gpu_parity.py::test_heaviest_scenarios_of_config3_alone)."""
    idx = np.array([62420, 60466, 60091])
    phys, x0, N = o.make_batch(3, S=int(idx.max()) + 1)
    sub = {k: np.asarray(v)[idx].copy() for k, v in phys.items()}
    c = co.closed_loop_batch(sub, x0[idx], N, flags=o.LITERAL_FIXED.flags())
    assert np.all(c["status"] == 0)
    for i in range(len(idx)):
        r = o.closed_loop(o.scenario(sub, i), x0[idx][i], N, profile=o.LITERAL_FIXED)
        du = np.max(np.abs(np.asarray(r["uk"]).reshape(-1) - c["uk"][i])) / sub["umax"][i]
        assert du <= 1e-10, (idx[i], du)


# ------------------------------------------------------------------ committed golden fixtures
@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_golden_fixtures_regress_c_oracle(cfg):
    g = np.load(os.path.join(GOLD, f"closed_loop_config{cfg}.npz"))
    S, N = int(g["S"]), int(g["N"])
    phys = {k[5:]: g[k] for k in g.files if k.startswith("phys_")}
    ph2, x0, N2 = o.make_batch(cfg, S=S)
    assert N2 == N and np.array_equal(x0, g["x0"])
    assert np.array_equal(o.derive_params_batch(ph2), g["params"])
    for name in ("literal_fixed", "literal", "consistent_fixed"):
        if f"{name}_xk" not in g.files:
            continue
        if cfg == 1 and name == "literal":
            continue                                              # bit-chaotic (D14)
        rc = co.closed_loop_batch(phys, x0, N, flags=int(g[f"{name}_flags"]))
        umax = phys["umax"][:, None]
        assert np.max(np.abs(rc["uk"] - g[f"{name}_uk"]) / umax) <= 1e-6
        w = g[f"{name}_xk"][:, :, 0]
        assert np.max(np.abs(rc["xk"][:, :, 0] - w) / np.maximum(np.max(np.abs(w), axis=1, keepdims=True), 1e-3)) <= 1e-6


def test_make_batch_prefix_property_and_ranges():
    for cfg in (2, 3, 4, 5):
        pa, xa, N = o.make_batch(cfg, S=16)
        pb, xb, _ = o.make_batch(cfg, S=64)
        assert np.array_equal(xa, xb[:16])
        for k in pa:
            assert np.array_equal(pa[k], pb[k][:16])
        assert np.all((xb[:, 0] >= 0.06) & (xb[:, 0] <= 0.15))
    p4, x4, _ = o.make_batch(4, S=256)
    assert np.all((p4["umax"] >= 0.2e6) & (p4["umax"] <= 2e6))
    assert np.all((x4[:, 1] >= 200 * math.pi) & (x4[:, 1] <= 4000 * math.pi))


def test_getWLc_structure_and_default_infeasibility():
    """getWLc.m: row blocks of 6 per stage + 4 terminal; D18: x0(1) = 0 < min_width makes the default problem infeasible."""
    p = o.default_physics()
    Af, Bf, C = o.model_callables(p)
    N = 3
    x0 = o.default_x0()
    R1 = np.full(N, o.rho1(x0, p["w_marg"])); R2 = np.full(N, o.rho2(x0)); R3 = np.full(N, o.rho3(x0, p["w_dep"]))
    Phi, Gam, Lam = o.Rho_to_PhiGammaLambda(R1, R2, R3, Af, Bf, C)
    W, L, c = o.getWLc([0.15, 5000 * 2 * math.pi], [0.06, 100 * 2 * math.pi], [2e6], [0.0], Gam, Phi, Lam)
    assert W.shape == (6 * N + 4, 2) and L.shape == (6 * N + 4, N) and c.shape == (6 * N + 4,)
    assert np.array_equal(L[0:6, 0], [-1, 1, 0, 0, 0, 0])                      # Ei on the block diagonal
    assert np.array_equal(W[2:6], [[1, 0], [0, 1], [-1, 0], [0, -1]])          # -Dcal: block 0 constrains x0 itself
    rhs = c + W @ x0
    assert rhs[2] < 0 and np.all(L[2] == 0)                                    # 0 <= -0.06 + 0: infeasible for every U (D18)
