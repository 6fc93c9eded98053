"""CPU oracle for the LPV-MPC step of IsaacSavona/MPC-NTM-Control  (TEST INFRASTRUCTURE ONLY).

This module is the parity checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product path
(``mpc-ntm-control_b200/``) never does and fails loudly when the CUDA library is missing.

What it restates (all citations relative to the upstream reference tree):

* ``NTM_MPC_Sim.m:5-37``   physics literals, ``kappa``, ``zeta``, affine term ``C``
* ``rho1.m:2`` (variant ``rhos.m:18``), ``rho2.m:2``, ``rho3.m:2-3``   scheduling functions
* ``A.m:2``, ``B.m:2``      LPV matrices (B acts as a 2x1 column)
* ``Rho_to_PhiGammaLambda.m:17-52``  horizon condensation
* ``NTM_MPC_Sim.m:67-73,120-121``    G = 2 Gamma' Omega Gamma, F = 2 Gamma' Omega (Phi x + Lambda - R)
* ``NTM_MPC_Sim.m:93-131``           closed loop (QP, rollout with old rho, re-condense, stop rule, plant)

The committed reference does not execute (wrong argument counts, row/column mix-ups, two
closed toolboxes).  This restatement follows the *literal, minimally repaired* (LMR) reading:
repair only what cannot execute, keep every executable expression verbatim.  Each quirk has a
switch (``Profile``) so that the "consistent" reading is one flag away.

PARITY PINNING.  MATLAB/Octave are absent from the build container and the reference ships no
tests, fixtures or golden vectors, so the oracle is pinned by (i) closed-form known answers
derived from the reference formulas (``tests/golden/appendix_a.json``), (ii) a 50-digit mpmath
twin of the same formulas (``tests/test_oracle.py``) and (iii) solver-independent KKT
certificates plus SciPy BVLS for the QP.  The QP arithmetic of the reference lives in the
closed-source MathWorks ``quadprog`` (no version pinned, ``NTM_MPC_Sim.m:88,97``):
**parity unpinned at the QP boundary** -- the box QP is strictly convex, so any exact solver
returns the same unique minimiser up to conditioning.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# Profile switches
# --------------------------------------------------------------------------------------
RHO1_LIN, RHO1_SQ = 0, 1                  # rho1.m:2  vs  rhos.m:18
GAMMA_I_MINUS_J, GAMMA_I = 0, 1           # Rho_to_PhiGammaLambda.m:32 literal index vs intent
F_X0, F_XK = 0, 1                         # NTM_MPC_Sim.m:73,121 uses x0
PLANT_NO_C, PLANT_WITH_C = 0, 1           # NTM_MPC_Sim.m:130 omits +C
INNER_EPS_BREAK, INNER_FIXED = 0, 1       # NTM_MPC_Sim.m:123-126
PLANT_EULER, PLANT_RK4 = 0, 1             # NTM_MPC_Sim.m:130 is the forward-Euler map; RK4 = fidelity option
STATE_ROWS_OFF, STATE_ROWS_REFRESH, STATE_ROWS_FROZEN = 0, 1, 2   # getWLc's state rows in the loop
TAUE_CONST, TAUE_W = 0, 1                 # NTM_MPC_Sim.m:14 "tau_E = tau_E0; currently NOT EXACT FORMULA" vs tau_E(w)


@dataclasses.dataclass(frozen=True)
class Profile:
    rho1_variant: int = RHO1_LIN
    gamma_index: int = GAMMA_I_MINUS_J
    f_state: int = F_X0
    plant_affine: int = PLANT_NO_C
    inner_policy: int = INNER_EPS_BREAK
    plant_integrator: int = PLANT_EULER
    tau_e_model: int = TAUE_CONST

    def flags(self) -> int:
        """Bit-packed form shared with include/ntm_mpc.h (NTM_PROFILE_* bits; bit 5 is the GPU-only DENSE_G switch)."""
        return (self.rho1_variant | (self.gamma_index << 1) | (self.f_state << 2)
                | (self.plant_affine << 3) | (self.inner_policy << 4) | (self.plant_integrator << 6)
                | (self.tau_e_model << 7))


LITERAL = Profile()
CONSISTENT = Profile(RHO1_LIN, GAMMA_I, F_XK, PLANT_WITH_C, INNER_EPS_BREAK)
LITERAL_FIXED = dataclasses.replace(LITERAL, inner_policy=INNER_FIXED)
CONSISTENT_FIXED = dataclasses.replace(CONSISTENT, inner_policy=INNER_FIXED)


# --------------------------------------------------------------------------------------
# Physics constants, NTM_MPC_Sim.m:5-25,31,34,37,47-50,59-60
# --------------------------------------------------------------------------------------
def default_physics() -> Dict[str, float]:
    """The literals of NTM_MPC_Sim.m:5-22 (+ Ts :31, bounds :47-48, cost :59-60)."""
    return dict(
        j_BS=73e3, w_dep=0.024, w_marg=0.02, w_sat=0.32, tau_r=293.0, rs=1.55, a=2.0,
        eta_CD=0.9, tau_E0=3.7, mu0=4e-7 * math.pi, Lq=0.87, B_pol=0.97, m=2.0, Cw=1.0,
        tau_A0=3e-6, tau_w=0.188, omega0=2 * math.pi * 420,
        Ts=0.1, umin=0.0, umax=2e6, r1=0.0, r2=1000 * 2 * math.pi,
        q11=1.0, q12=0.0, q22=1.0,
        c_tauE=0.0,                       # tau_E(w) hook (TAUE_W), see tau_E_of; 0 = the script's constant tau_E
    )


def default_x0() -> np.ndarray:
    """NTM_MPC_Sim.m:34."""
    return np.array([0.0, 1000 * 2 * math.pi])


def kappa_of(p) -> float:
    """NTM_MPC_Sim.m:24."""
    return 16 * p["mu0"] * p["Lq"] * p["rs"] ** 2 / (0.82 * p["tau_r"] * p["B_pol"] * math.pi)


def zeta_of(p) -> float:
    """NTM_MPC_Sim.m:25."""
    return p["m"] * p["Cw"] * p["tau_A0"] ** 2 * p["tau_w"] * p["a"] ** 3


def C_of(p) -> np.ndarray:
    """NTM_MPC_Sim.m:37."""
    kappa = kappa_of(p)
    return np.array([
        -4 / 3 * (kappa * p["Ts"] * p["j_BS"] * p["w_sat"]) / (p["w_sat"] ** 2 + p["w_marg"] ** 2),
        p["Ts"] * p["omega0"] / p["tau_E0"],
    ])


def tau_E_of(p, w, model: int = TAUE_CONST):
    """NTM_MPC_Sim.m:14 sets ``tau_E = tau_E0`` and flags it "currently NOT EXACT FORMULA"; the authors give no other.
    ``TAUE_W`` is the documented hook: ``tau_E(w) = tau_E0 * (1 - c_tauE * w)`` with the coefficient
    ``c_tauE`` [1/m] an ordinary entry of the physics dictionary (slot 15 of the parameter block).  ``c_tauE_belt``
    gives the value of the belt model of confinement degradation by an island [external knowledge: Chang & Callen,
    Nucl. Fusion 30 (1990) 219: d tau_E / tau_E = -4 w r_s^3 / a^4]; ``c_tauE = 0`` is the script as written."""
    if model == TAUE_CONST:
        return p["tau_E0"]
    return p["tau_E0"] * (1 - p.get("c_tauE", 0.0) * w)


def c_tauE_belt(p):
    """4 r_s^3 / a^4 [1/m] from the script's own ``rs`` (:10) and ``a`` (:11): 0.93 per metre on the nominal physics."""
    return 4 * p["rs"] ** 3 / p["a"] ** 4


# --------------------------------------------------------------------------------------
# Scheduling functions
# --------------------------------------------------------------------------------------
def rho1(x, wmarg, variant: int = RHO1_LIN):
    """rho1.m:2  ``1/(x(1)+wmarg^2)``;  variant 'sq' is rhos.m:18 ``1/(x(1)^2+w_marg^2)``."""
    if variant == RHO1_SQ:
        return 1.0 / (x[0] ** 2 + wmarg ** 2)
    return 1.0 / (x[0] + wmarg ** 2)


def rho2(x):
    """rho2.m:2  ``x(1)^2/x(2)``."""
    return x[0] ** 2 / x[1]


def rho3(x, w_dep):
    """rho3.m:2-3."""
    wstar = x[0] / w_dep
    return (0.25 + 0.24 * wstar) / (1 + 1.5 * wstar + 0.43 * wstar ** 2 + 0.64 * wstar ** 3)


# --------------------------------------------------------------------------------------
# LPV matrices
# --------------------------------------------------------------------------------------
def A_mat(r1, r2, kappa, taur, Ts, zeta, rs, a, TE) -> np.ndarray:
    """A.m:2 verbatim (MATLAB evaluates left to right)."""
    return np.array([
        [((4 / 3) * (kappa * rs / (0.82 * taur)) * Ts * r1 + 1), 0.0],
        [((r2 * Ts) / (zeta * a ** 3)), (1 - Ts / TE)],
    ])


def B_mat(r3, wdep, kappa, Ts, etaCD) -> np.ndarray:
    """B.m:2; the reference returns a 1x2 row that must act as the 2x1 column [b;0] (defect D7)."""
    return np.array([(kappa * Ts * etaCD / wdep) * r3, 0.0])


def model_callables(p, tau_E=None) -> Tuple[Callable, Callable, np.ndarray]:
    """The call forms the script uses -- ``A(r1,r2)``, ``B(r3)`` (NTM_MPC_Sim.m:113) -- closed
    over the workspace constants the .m signatures require (repair of D10).  ``tau_E``: the workspace value A.m's
    ``TE`` argument picks up (default: NTM_MPC_Sim.m:14, ``tau_E0``)."""
    kappa, zeta = kappa_of(p), zeta_of(p)
    if tau_E is None:
        tau_E = p["tau_E0"]                                    # NTM_MPC_Sim.m:14

    def Af(r1, r2):
        return A_mat(r1, r2, kappa, p["tau_r"], p["Ts"], zeta, p["rs"], p["a"], tau_E)

    def Bf(r3):
        return B_mat(r3, p["w_dep"], kappa, p["Ts"], p["eta_CD"])

    return Af, Bf, C_of(p)


# --------------------------------------------------------------------------------------
# Horizon condensation, Rho_to_PhiGammaLambda.m
# --------------------------------------------------------------------------------------
def Rho_to_PhiGammaLambda(Rho1, Rho2, Rho3, A, B, C, gamma_index: int = GAMMA_I_MINUS_J):
    """Phi (2N x 2), Gamma (2N x N), Lambda (2N,) from the rho sequences.

    N = numel(Rho1) (repair of D4).  Phi blocks left-multiply like :32 and :51 (repair of D5).
    Gamma off-diagonal blocks use ``A(Rho(i-j))`` literally (:32) or ``A(Rho(i))`` (intent,
    comment :33-34) when ``gamma_index == GAMMA_I``.  MATLAB indices are 1-based; arrays here
    are 0-based, so ``Rho1[i-1]`` is MATLAB's ``Rho1(i)``.
    """
    Rho1 = np.asarray(Rho1, dtype=np.float64).ravel()
    Rho2 = np.asarray(Rho2, dtype=np.float64).ravel()
    Rho3 = np.asarray(Rho3, dtype=np.float64).ravel()
    N = Rho1.size
    nx = 2
    C = np.asarray(C, dtype=np.float64).ravel()

    Phi = np.zeros((nx * N, nx))
    Phi[0:nx, :] = A(Rho1[0], Rho2[0])                                         # :8,18
    for j in range(2, N + 1):                                                  # :20-22
        Phi[(j - 1) * nx:j * nx, :] = A(Rho1[j - 1], Rho2[j - 1]) @ Phi[(j - 2) * nx:(j - 1) * nx, :]

    Gamma = np.zeros((nx * N, N))
    Gamma[0:nx, 0] = B(Rho3[0])                                                # :27
    for i in range(2, N + 1):                                                  # :28
        for j in range(1, i + 1):                                              # :29
            if i != j:
                k = (i - j) if gamma_index == GAMMA_I_MINUS_J else i           # :32
                Gamma[(i - 1) * nx:i * nx, j - 1] = A(Rho1[k - 1], Rho2[k - 1]) @ Gamma[(i - 2) * nx:(i - 1) * nx, j - 1]
            else:
                Gamma[(i - 1) * nx:i * nx, j - 1] = B(Rho3[j - 1])             # :36

    Lambda = np.zeros(nx * N)
    Lambda[0:nx] = C                                                           # :48
    for i in range(2, N + 1):                                                  # :49-52
        Lambda[(i - 1) * nx:i * nx] = A(Rho1[i - 1], Rho2[i - 1]) @ Lambda[(i - 2) * nx:(i - 1) * nx] + C
    return Phi, Gamma, Lambda


def hessian_grad(Phi, Gamma, Lambda, x, r, Q):
    """NTM_MPC_Sim.m:67-73 / :120-121.  Omega = blkdiag(Q,...,Q); R = repmat(r(:),N,1) (repair of D8)."""
    N = Gamma.shape[1]
    Omega = np.kron(np.eye(N), np.asarray(Q, dtype=np.float64))                # :67-70
    R = np.tile(np.asarray(r, dtype=np.float64).ravel(), N)                    # :71 repaired
    G = 2 * Gamma.T @ Omega @ Gamma                                            # :72
    F = 2 * Gamma.T @ Omega @ (Phi @ np.asarray(x, dtype=np.float64) + Lambda - R)   # :73
    return G, F


# --------------------------------------------------------------------------------------
# State + input constraint condensation, getWLc.m  (the "next" row)
# --------------------------------------------------------------------------------------
def _blkdiag(*ms):
    r = sum(m.shape[0] for m in ms); c = sum(m.shape[1] for m in ms)
    out = np.zeros((r, c)); i = j = 0
    for m in ms:
        out[i:i + m.shape[0], j:j + m.shape[1]] = m
        i += m.shape[0]; j += m.shape[1]
    return out


def getWLc(xmax, xmin, umax, umin, Gamma, Phi, Lambda):
    """getWLc.m:1-63 -- stacks input-box and state-box constraints over the horizon into ``L U <= c + W x``.
    Literal except for defect D9 (``Ccal(i) = bi`` assigns a 6-vector into a scalar slot, getWLc.m:51-55): Ccal is
    the stack [bi; ...; bi; bN] the commented-out lines :47-50 describe."""
    xmax = np.asarray(xmax, dtype=np.float64).reshape(-1, 1); xmin = np.asarray(xmin, dtype=np.float64).reshape(-1, 1)
    umax = np.asarray(umax, dtype=np.float64).reshape(-1, 1); umin = np.asarray(umin, dtype=np.float64).reshape(-1, 1)
    nu, nx = umin.shape[0], xmin.shape[0]                                     # :5-6
    N = Phi.shape[0] // nx                                                    # :7
    Mi = np.vstack([np.zeros((nu, nx)), np.zeros((nu, nx)), -np.eye(nx), np.eye(nx)])          # :9-12
    Ei = np.vstack([-np.eye(nu), np.eye(nu), np.zeros((nx, nu)), np.zeros((nx, nu))])          # :14-17
    bi = np.vstack([-umin, umax, -xmin, xmax])                                # :20-23
    MN = np.vstack([-np.eye(nx), np.eye(nx)]); bN = np.vstack([-xmin, xmax])  # :25-26
    Dcal = np.vstack([Mi] + [0 * Mi] * (N - 1) + [0 * MN])                    # :30
    Mcal = MN                                                                 # :33-37
    for _ in range(2, N + 1):
        Mcal = _blkdiag(Mi, Mcal)
    Mcal = np.vstack([np.zeros((Mi.shape[0], Mcal.shape[1])), Mcal])
    Ecal = Ei                                                                 # :40-44
    for _ in range(2, N + 1):
        Ecal = _blkdiag(Ecal, Ei)
    Ecal = np.vstack([Ecal, np.zeros((MN.shape[0], Ecal.shape[1]))])
    Ccal = np.vstack([bi] * N + [bN])                                         # :47-55 repaired (D9)
    L = Mcal @ Gamma + Ecal                                                   # :57
    W = -Dcal - Mcal @ Phi                                                    # :58
    c = Ccal - Mcal @ np.asarray(Lambda, dtype=np.float64).reshape(-1, 1)     # :59
    return W, L, c[:, 0]


# --------------------------------------------------------------------------------------
# Box QP  (stands in for quadprog(G,F,L,c+W*x) with only the Ei/bi input rows of getWLc.m:14-23)
# --------------------------------------------------------------------------------------
def qp_kkt_residual(G, F, lb, ub, U) -> float:
    """Scale-free KKT residual of ``min 1/2 U'GU + F'U, lb<=U<=ub``: the largest violation of
    g_i>=0 at lb, g_i<=0 at ub, g_i=0 inside, relative to |F_i| + sum_j |G_ij||U_j|."""
    G = np.asarray(G); F = np.asarray(F); U = np.asarray(U)
    lb = np.broadcast_to(np.asarray(lb, dtype=np.float64), U.shape)
    ub = np.broadcast_to(np.asarray(ub, dtype=np.float64), U.shape)
    g = G @ U + F
    scale = np.abs(F) + np.abs(G) @ np.abs(U) + 1e-300
    at_lb, at_ub = U <= lb, U >= ub
    viol = np.where(at_lb & at_ub, 0.0, np.where(at_lb, np.maximum(-g, 0), np.where(at_ub, np.maximum(g, 0), np.abs(g))))
    feas = np.maximum(np.maximum(lb - U, U - ub), 0) / (np.abs(ub - lb) + 1e-300)
    return float(max(np.max(viol / scale), np.max(feas)))


def qp_box(G, F, lb, ub, max_iter: Optional[int] = None):
    """Exact primal active-set solve of ``min 1/2 U'GU + F'U  s.t. lb <= U <= ub`` (G SPD).

    Works in scaled variables (U = lb + (ub-lb)*t, 0<=t<=1) because cond(G) reaches 1e9..1e11 in
    raw watts.  Returns ``(U, iterations, status)`` with status 0 = KKT point,
    1 = iteration cap, 2 = non-finite data.  Bound components are *exactly* lb/ub so that the
    reference's bit-sensitive stop rule (NTM_MPC_Sim.m:123) sees saturated iterates as equal.
    """
    G = np.asarray(G, dtype=np.float64)
    F = np.asarray(F, dtype=np.float64).ravel()
    N = F.size
    lb = np.broadcast_to(np.asarray(lb, dtype=np.float64), (N,)).copy()
    ub = np.broadcast_to(np.asarray(ub, dtype=np.float64), (N,)).copy()
    if not (np.all(np.isfinite(G)) and np.all(np.isfinite(F))):
        return np.full(N, np.nan), 0, 2
    rng = ub - lb
    # scaled problem: 1/2 t'Hs t + fs't
    Hs = G * rng[:, None] * rng[None, :]
    fs = (F + G @ lb) * rng
    t = np.zeros(N)
    state = -np.ones(N, dtype=np.int64)          # -1 at lower, +1 at upper, 0 free; start at lb
    tol_g = 1e-13
    if max_iter is None:
        max_iter = 10 * N + 20
    status = 1
    it = 0
    for it in range(1, max_iter + 1):
        free = state == 0
        alpha, block, bstate = 1.0, -1, 0
        if free.any():
            g = Hs @ t + fs
            p = np.zeros(N)
            p[free] = np.linalg.solve(Hs[np.ix_(free, free)], -g[free])
            for i in np.flatnonzero(free):                     # ratio test along p
                if p[i] < 0:
                    a_i = (0.0 - t[i]) / p[i]
                    if a_i < alpha:
                        alpha, block, bstate = a_i, i, -1
                elif p[i] > 0:
                    a_i = (1.0 - t[i]) / p[i]
                    if a_i < alpha:
                        alpha, block, bstate = a_i, i, 1
            t = t + alpha * p
        if block >= 0:                                         # a bound blocks: add it, stay on the arc
            state[block] = bstate
            t[block] = 0.0 if bstate == -1 else 1.0
            continue
        # full step taken: t minimises on the current face -> check the bound multipliers
        g = Hs @ t + fs
        gscale = np.abs(fs) + np.abs(Hs) @ np.abs(t) + 1e-300
        lam = np.where(state == -1, g, np.where(state == 1, -g, 0.0)) / gscale
        worst = int(np.argmin(lam))
        if lam[worst] >= -tol_g:
            status = 0
            break
        state[worst] = 0
    # polish: re-solve the free block from scratch on the final partition, in raw data
    free = state == 0
    U = np.where(state == 1, ub, lb)
    if free.any():
        act = ~free
        rhs = -(F[free] + G[np.ix_(free, act)] @ U[act])
        Hff = G[np.ix_(free, free)] * rng[free][:, None] * rng[free][None, :]
        tf = np.linalg.solve(Hff, rhs * rng[free])
        U[free] = np.minimum(np.maximum(tf * rng[free], lb[free] - 0.0), ub[free])
    return U, it, status


# --------------------------------------------------------------------------------------
# General-inequality QP  (quadprog(G,F,L,c+W*x) as NTM_MPC_Sim.m:97 actually calls it: input box rows
# getWLc.m:14-23 AND the state rows getWLc.m:11-12,25 -- SURVEY 8(f)-1)
# --------------------------------------------------------------------------------------
QP_OK, QP_ITER_CAP, QP_NONFINITE, QP_INFEASIBLE = 0, 1, 2, 3


def split_rows(L, b):
    """Host-side classification of the rows of ``L U <= b`` (what the quadprog gateway does with getWLc's output):
    rows with one non-zero tighten a bound, all-zero rows are feasibility checks (the x_0 block of getWLc.m:30),
    the rest are general rows.  Returns ``(lb, ub, Lg, bg, feasible)``; missing bounds are -inf/+inf."""
    L = np.asarray(L, dtype=np.float64); b = np.asarray(b, dtype=np.float64).ravel()
    M, N = L.shape
    lb = np.full(N, -np.inf); ub = np.full(N, np.inf)
    gen = []
    feasible = True
    for i in range(M):
        nz = np.flatnonzero(L[i])
        if nz.size == 0:
            if b[i] < 0:
                feasible = False
        elif nz.size == 1:
            j = int(nz[0]); a = L[i, j]
            if a > 0:
                ub[j] = min(ub[j], b[i] / a)
            else:
                lb[j] = max(lb[j], b[i] / a)
        else:
            gen.append(i)
    return lb, ub, L[gen], b[gen], feasible


def qp_ineq_kkt_residual(G, F, lb, ub, Lg, bg, U):
    """Certificate for ``min 1/2 U'GU + F'U, lb<=U<=ub, Lg U<=bg``: solves for the multipliers of the rows that
    are tight at U by non-negative least squares and returns (stationarity residual, worst violation), both
    relative.  Independent of how U was obtained."""
    from scipy.optimize import nnls
    G = np.asarray(G); F = np.asarray(F).ravel(); U = np.asarray(U).ravel(); N = U.size
    lb = np.broadcast_to(np.asarray(lb, dtype=np.float64), (N,)); ub = np.broadcast_to(np.asarray(ub, dtype=np.float64), (N,))
    Lg = np.asarray(Lg, dtype=np.float64).reshape(-1, N); bg = np.asarray(bg, dtype=np.float64).ravel()
    rng = ub - lb
    g = (G @ U + F) * rng                                   # gradient in scaled variables
    gs = (np.abs(F) + np.abs(G) @ np.abs(U)) * rng + 1e-300
    rows = Lg * rng[None, :]
    rs = np.abs(rows).sum(axis=1) + 1e-300
    slack = (bg - Lg @ U) / rs
    cols = []
    for j in range(N):
        if U[j] <= lb[j]:
            cols.append(-np.eye(N)[j])
        if U[j] >= ub[j]:
            cols.append(np.eye(N)[j])
    for i in range(Lg.shape[0]):
        if slack[i] <= 1e-9:
            cols.append(rows[i] / rs[i])
    viol = max(float(np.max(-slack, initial=0.0)), float(np.max((lb - U) / rng, initial=0.0)),
               float(np.max((U - ub) / rng, initial=0.0)))
    if not cols:
        return float(np.max(np.abs(g) / gs)), viol
    Nm = np.array(cols).T
    lam, _ = nnls(Nm / gs[:, None], -g / gs)
    return float(np.max(np.abs(Nm @ lam + g) / gs)), viol


def qp_ineq(G, F, lb, ub, Lg, bg, max_iter: Optional[int] = None, tol: float = 1e-9, stuck_tol: float = 1e-6):
    """``min 1/2 U'GU + F'U  s.t. lb <= U <= ub, Lg U <= bg`` (G SPD, finite bounds).

    Two phases.  (1) ``qp_box`` gives the minimiser over the box; if it satisfies the general rows it is the answer
    (the common case).  (2) Otherwise it is a dual-feasible start for a dual active-set continuation
    (Goldfarb & Idnani 1983, restated with direct KKT solves on the free block: the most violated row enters, the
    primal/dual step lengths t2/t1 decide between a full step and dropping the blocking constraint; no step
    possible = infeasible).  Works in scaled variables t = (U-lb)/(ub-lb) with the general rows divided by their
    1-norm.  Returns ``(U, iterations, status)``; status QP_OK / QP_ITER_CAP / QP_NONFINITE / QP_INFEASIBLE
    (quadprog's exitflag 1 / 0 / -- / -2).  Bound components are exactly lb/ub."""
    G = np.asarray(G, dtype=np.float64); F = np.asarray(F, dtype=np.float64).ravel(); N = F.size
    lb = np.broadcast_to(np.asarray(lb, dtype=np.float64), (N,)).copy()
    ub = np.broadcast_to(np.asarray(ub, dtype=np.float64), (N,)).copy()
    Lg = np.asarray(Lg, dtype=np.float64).reshape(-1, N); bg = np.asarray(bg, dtype=np.float64).ravel()
    M = bg.size
    if not (np.all(np.isfinite(G)) and np.all(np.isfinite(F)) and np.all(np.isfinite(Lg)) and np.all(np.isfinite(bg))
            and np.all(np.isfinite(lb)) and np.all(np.isfinite(ub))):
        return np.full(N, np.nan), 0, QP_NONFINITE
    U, it, st = qp_box(G, F, lb, ub)
    if st != 0:
        return U, it, st
    rng = np.where(ub > lb, ub - lb, 1.0)                    # a pinned variable keeps unit scale and 0 <= t <= 0
    hb = np.where(ub > lb, 1.0, 0.0)
    Hs = G * rng[:, None] * rng[None, :]
    fs = (F + G @ lb) * rng
    A = Lg * rng[None, :]
    rs = np.abs(A).sum(axis=1)
    bs = bg - Lg @ lb
    zero = rs == 0
    if np.any(bs[zero] < 0):
        return U, it, QP_INFEASIBLE
    rs = np.where(zero, 1.0, rs)
    A = A / rs[:, None]; bs = np.where(zero, 0.0, bs / rs)
    state = np.where((U >= ub) & (ub > lb), 1, np.where(U <= lb, -1, 0)).astype(np.int64)
    t = np.where(state == 1, 1.0, np.where(state == -1, 0.0, (U - lb) / rng))
    Wg: list = []                                            # active general rows, in order of entry
    g = Hs @ t + fs
    ub_mult = np.where(state == -1, np.maximum(g, 0.0), np.where(state == 1, np.maximum(-g, 0.0), 0.0))
    ug: list = []                                            # multipliers of Wg
    if max_iter is None:
        max_iter = 20 * (N + M) + 50

    Lc = np.linalg.cholesky(Hs)

    def direction(npl):
        """Goldfarb-Idnani step data for the entering normal n+, factorised from scratch every time (Householder
        QR of Lc^{-1} N, N = active normals): z = J2 J2' n+ (primal direction), r = R^{-1} J1' n+ (dual direction,
        returned per bound and per active general row), |d2|^2 = z'n+ and |d|^2 = n+' Hs^{-1} n+."""
        fixed = np.flatnonzero(state != 0)
        q = fixed.size + len(Wg)
        Nact = np.zeros((N, q))
        for k, j in enumerate(fixed):
            Nact[j, k] = 1.0 if state[j] == -1 else -1.0      # t_j >= 0 : +e_j ;  -t_j >= -1 : -e_j
        for k, i in enumerate(Wg):
            Nact[:, fixed.size + k] = -A[i]
        Bm = np.linalg.solve(Lc, Nact)
        Qf, Rf = np.linalg.qr(Bm, mode="complete") if q else (np.eye(N), np.zeros((N, 0)))
        d = Qf.T @ np.linalg.solve(Lc, npl)
        dd2 = float(d[q:] @ d[q:]) if q < N else 0.0
        dd = float(d @ d)
        z = np.linalg.solve(Lc.T, Qf[:, q:] @ d[q:]) if q < N else np.zeros(N)
        r = np.linalg.solve(Rf[:q, :q], d[:q]) if q else np.zeros(0)
        rb = np.zeros(N); rb[fixed] = r[:fixed.size]
        return z, rb, r[fixed.size:], dd2, dd

    status = QP_ITER_CAP
    while it < max_iter:
        it += 1
        v = A @ t - bs if M else np.zeros(0)
        if Wg:
            v[Wg] = -np.inf
        vlo = np.where(state == -1, -np.inf, -t); vhi = np.where(state == 1, -np.inf, t - hb)
        cand = [(float(np.max(v, initial=-np.inf)), 2), (float(np.max(vlo)), 0), (float(np.max(vhi)), 1)]
        worst, kind = max(cand)
        if worst <= tol:
            status = QP_OK
            break
        if kind == 2:
            p = int(np.argmax(v)); npl = -A[p].copy(); bp = -bs[p]
        elif kind == 0:
            p = int(np.argmax(vlo)); npl = np.zeros(N); npl[p] = 1.0; bp = 0.0
        else:
            p = int(np.argmax(vhi)); npl = np.zeros(N); npl[p] = -1.0; bp = -hb[p]
        up = 0.0
        infeasible = False
        stuck = False
        while True:
            z, rb, rg, dd2, dd = direction(npl)
            sp = float(npl @ t) - bp                         # < 0: violated
            zero = not (dd2 > 1e-18 * dd)                    # n+ depends on the active normals: no primal step
            t2 = np.inf if zero else -sp / dd2
            t1, drop = np.inf, None
            for j in range(N):
                if state[j] != 0 and rb[j] > 0 and ub_mult[j] / rb[j] < t1:
                    t1, drop = ub_mult[j] / rb[j], ("b", j)
            for i in range(len(Wg)):
                if rg[i] > 0 and ug[i] / rg[i] < t1:
                    t1, drop = ug[i] / rg[i], ("g", i)
            tau = min(t1, t2)
            if not np.isfinite(tau):
                infeasible = -sp > stuck_tol                 # a rounding-level violation that cannot be removed is not infeasibility
                stuck = not infeasible
                break
            if not zero:
                t = t + tau * z
            ub_mult = np.where(state != 0, np.maximum(ub_mult - tau * rb, 0.0), 0.0)
            ug = [max(ug[i] - tau * rg[i], 0.0) for i in range(len(Wg))]
            up += tau
            if tau == t2:                                    # full step: p joins the active set
                if kind == 2:
                    Wg.append(p); ug.append(up)
                else:
                    state[p] = -1 if kind == 0 else 1
                    t[p] = 0.0 if kind == 0 else hb[p]
                    ub_mult[p] = up
                break
            if drop[0] == "b":                               # partial step: the blocking constraint leaves
                state[drop[1]] = 0; ub_mult[drop[1]] = 0.0
            else:
                Wg.pop(drop[1]); ug.pop(drop[1])
            it += 1
            if it >= max_iter:
                break
        if infeasible:
            status = QP_INFEASIBLE
            break
        if stuck:
            status = QP_OK
            break
    U = np.where(state == 1, ub, np.where(state == -1, lb, lb + rng * t))
    return U, it, status


# --------------------------------------------------------------------------------------
# Per-scenario derived coefficients (the 16-double parameter block of include/ntm_mpc.h)
# --------------------------------------------------------------------------------------
PARAM_NAMES = ("c_a11", "c_a21", "a22", "c_b", "C1", "C2", "wmarg2", "w_dep",
               "umin", "umax", "r1", "r2", "q11", "q12", "q22", "c_tauE")
NPARAM = len(PARAM_NAMES)


def derive_params(p) -> np.ndarray:
    """Hoisted coefficients of A.m:2 / B.m:2 / NTM_MPC_Sim.m:37 in the reference's own
    evaluation order (so c_a11*rho1+1 and c_b*rho3 round exactly like A.m / B.m)."""
    kappa, zeta = kappa_of(p), zeta_of(p)
    C = C_of(p)
    tau_E = p["tau_E0"]
    return _stack([
        (4 / 3) * (kappa * p["rs"] / (0.82 * p["tau_r"])) * p["Ts"],
        p["Ts"] / (zeta * p["a"] ** 3),
        1 - p["Ts"] / tau_E,
        (kappa * p["Ts"] * p["eta_CD"] / p["w_dep"]),
        C[0], C[1],
        p["w_marg"] ** 2, p["w_dep"],
        p["umin"], p["umax"], p["r1"], p["r2"], p["q11"], p["q12"], p["q22"], p.get("c_tauE", 0.0),
    ])


def _stack(vals):
    """np.array for scalars; [NPARAM, S] for a batch dict of arrays."""
    shapes = [np.shape(v) for v in vals]
    if all(sh == () for sh in shapes):
        return np.array(vals, dtype=np.float64)
    S = max(sh[0] for sh in shapes if sh != ())
    return np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), (S,)) for v in vals])


# --------------------------------------------------------------------------------------
# Closed loop, NTM_MPC_Sim.m:63-73 (offline build) and :80-131 (simulation)
# --------------------------------------------------------------------------------------
def plant_step(p, x, u, profile: Profile = LITERAL):
    """NTM_MPC_Sim.m:130: ``x+ = A(rho(x)) x + B(rho(x)) u`` (``+ C`` with PLANT_WITH_C).  That map is one forward-Euler
    step ``x + g(x, u)`` of the GRE model with ``g = Ts * dx/dt``; ``PLANT_RK4`` (not in the reference)
    takes the classical Runge-Kutta step of the same vector field over one sample instead, ``u`` held.  With ``TAUE_W``
    the vector field carries ``tau_E(w)`` of the state it is evaluated at (every RK4 stage its own)."""
    Af, Bf, C = model_callables(p)
    x = np.asarray(x, dtype=np.float64).ravel()

    def euler(z):
        Az = Af if profile.tau_e_model == TAUE_CONST else model_callables(p, tau_E_of(p, z[0], TAUE_W))[0]
        zn = Az(rho1(z, p["w_marg"], profile.rho1_variant), rho2(z)) @ z + Bf(rho3(z, p["w_dep"])) * u
        return zn + C if profile.plant_affine == PLANT_WITH_C else zn

    if profile.plant_integrator == PLANT_EULER:
        return euler(x)
    k1 = euler(x) - x
    y = x + 0.5 * k1; k2 = euler(y) - y
    y = x + 0.5 * k2; k3 = euler(y) - y
    y = x + k3; k4 = euler(y) - y
    return x + ((k1 + 2.0 * k2) + (2.0 * k3 + k4)) / 6.0


def qp_state_rows(G, F, lb, ub, Phi, Gamma, Lambda, xk, xmin, xmax):
    """NTM_MPC_Sim.m:97 as written: ``quadprog(G, F, L, c + W*xk)`` with ``[W, L, c] = getWLc(...)`` (:74) -- the input
    box AND the state rows.  Rows with one non-zero become bounds, empty rows feasibility checks (``split_rows``).
    Returns ``(U, iterations, status)``; an infeasible QP returns NaNs and ``QP_INFEASIBLE`` (exitflag -2)."""
    N = F.size
    W, L, c = getWLc(xmax, xmin, [ub], [lb], Gamma, Phi, Lambda)
    b = c + W @ np.asarray(xk, dtype=np.float64).ravel()
    if not (np.all(np.isfinite(L)) and np.all(np.isfinite(b))):
        return np.full(N, np.nan), 0, QP_NONFINITE
    lo, hi, Lg, bg, feasible = split_rows(L, b)
    if not feasible or np.any(lo > hi):
        return np.full(N, np.nan), 0, QP_INFEASIBLE
    U, it, st = qp_ineq(G, F, lo, hi, Lg, bg)
    if st == QP_INFEASIBLE:
        U = np.full(N, np.nan)
    return U, it, st


def closed_loop(p, x0, N: int = 3, k_sim: int = 20, i_sim: int = 10, eps: float = 1e-14,
                profile: Profile = LITERAL, qp: Callable = qp_box, trace: Optional[list] = None,
                state_rows: int = STATE_ROWS_OFF, xbounds=None, qp_log: Optional[list] = None):
    """One scenario of the repaired script.  Returns a dict with xk (2,k_sim+1), uk (k_sim,),
    Uk (N,k_sim), inner_iters (k_sim,), qp_iters (k_sim,), cost, status.

    ``state_rows``: keep getWLc's state rows in the QP of :97.  ``STATE_ROWS_FROZEN`` is the literal
    script -- ``[W, L, c] = getWLc(...)`` at :74 is outside both loops, so the rows keep the offline condensation --,
    ``STATE_ROWS_REFRESH`` rebuilds them from every re-condensation (:119).  ``xbounds = (xmin1, xmax1, xmin2, xmax2)``
    (:44-45).  An infeasible QP (exitflag -2, :100-101) ends the scenario: status 3, everything not yet produced NaN.

    ``profile.tau_e_model = TAUE_W``: the workspace ``tau_E`` that every ``A(r1, r2)`` call picks up is re-evaluated
    from the MEASURED island width at the top of each time step, ``tau_E = tau_E(xk(1,k))`` (``tau_E_of``), and held
    over the prediction horizon of that step -- the rho's vary along the horizon, ``tau_E`` does not -- and the plant
    step :130 uses the same value (Euler) or the value at each stage's state (RK4)."""
    Af, Bf, C = model_callables(p, tau_E_of(p, np.asarray(x0, dtype=np.float64).ravel()[0], profile.tau_e_model))
    x0 = np.asarray(x0, dtype=np.float64).ravel()
    wmarg, w_dep = p["w_marg"], p["w_dep"]
    Q = np.array([[p["q11"], p["q12"]], [p["q12"], p["q22"]]])
    r = np.array([p["r1"], p["r2"]])
    lb, ub = p["umin"], p["umax"]

    def r1f(x):
        return rho1(x, wmarg, profile.rho1_variant)

    # offline build :63-73
    Rho1 = np.full(N, r1f(x0)); Rho2 = np.full(N, rho2(x0)); Rho3 = np.full(N, rho3(x0, w_dep))
    Phi, Gamma, Lambda = Rho_to_PhiGammaLambda(Rho1, Rho2, Rho3, Af, Bf, C, profile.gamma_index)
    G, F = hessian_grad(Phi, Gamma, Lambda, x0, r, Q)
    if state_rows != STATE_ROWS_OFF:
        xmin = np.array([xbounds[0], xbounds[2]]); xmax = np.array([xbounds[1], xbounds[3]])
        rows_src = (Phi, Gamma, Lambda)                                        # :74

    xk = np.zeros((2, k_sim + 1)); xk[:, 0] = x0                               # :82
    uk = np.zeros(k_sim); Uk = np.zeros((N, k_sim))                            # :83-84
    Uold = np.ones(N)                                                          # :86 (D13)
    inner = np.zeros(k_sim, dtype=np.int32); qpit = np.zeros(k_sim, dtype=np.int32)
    status = 0
    with np.errstate(all="ignore"):
        for k in range(k_sim):                                                 # :93
            if profile.tau_e_model == TAUE_W:
                # a workspace assignment at the top of the time loop: G, F on hand (from the last :119-121 of the previous
                # step) keep the previous value, exactly as the script's G, F keep the previous xk under F_XK
                Af = model_callables(p, tau_E_of(p, xk[0, k], TAUE_W))[0]
            for it in range(1, i_sim + 1):                                     # :94
                if state_rows == STATE_ROWS_OFF:
                    U, nit, st = qp(G, F, lb, ub)                              # :97
                else:
                    if qp_log is not None:
                        qp_log.append(dict(k=k, it=it, G=G.copy(), F=F.copy(), rows=rows_src, xk=xk[:, k].copy()))
                    U, nit, st = qp_state_rows(G, F, lb, ub, *rows_src, xk[:, k], xmin, xmax)
                status = max(status, st)
                qpit[k] += nit
                if st == QP_INFEASIBLE:                                        # :100-101, no U: the script cannot go on
                    inner[k] = it
                    uk[k:] = np.nan; Uk[:, k:] = np.nan; xk[:, k + 1:] = np.nan
                    return dict(xk=xk, uk=uk, Uk=Uk, inner_iters=inner, qp_iters=qpit, cost=float("nan"), status=status)
                Uk[:, k] = U; uk[k] = U[0]                                     # :106-107
                xN = np.zeros((2, N + 1)); xN[:, 0] = xk[:, k]                 # :110
                for i in range(N):                                             # :112-117
                    xN[:, i + 1] = Af(Rho1[i], Rho2[i]) @ xN[:, i] + Bf(Rho3[i]) * U[i] + C
                    Rho1[i] = r1f(xN[:, i]); Rho2[i] = rho2(xN[:, i]); Rho3[i] = rho3(xN[:, i], w_dep)
                Phi, Gamma, Lambda = Rho_to_PhiGammaLambda(Rho1, Rho2, Rho3, Af, Bf, C, profile.gamma_index)  # :119
                xf = x0 if profile.f_state == F_X0 else xk[:, k]
                G, F = hessian_grad(Phi, Gamma, Lambda, xf, r, Q)              # :120-121
                if state_rows == STATE_ROWS_REFRESH:
                    rows_src = (Phi, Gamma, Lambda)
                if trace is not None:
                    trace.append(dict(k=k, it=it, U=U.copy(), Phi=Phi, Gamma=Gamma, Lambda=Lambda, G=G, F=F,
                                      Rho1=Rho1.copy(), Rho2=Rho2.copy(), Rho3=Rho3.copy()))
                inner[k] = it
                if profile.inner_policy == INNER_EPS_BREAK and np.sum(np.abs(Uold - U)) < eps:   # :123
                    break
                Uold = U.copy()                                                # :127
            xk[:, k + 1] = plant_step(p, xk[:, k], uk[k], profile)             # :130
    if not np.all(np.isfinite(xk)):
        status = max(status, 2)
    e = xk[:, 1:] - r[:, None]
    cost = float(np.sum(e * (Q @ e)))
    return dict(xk=xk, uk=uk, Uk=Uk, inner_iters=inner, qp_iters=qpit, cost=cost, status=status)


# --------------------------------------------------------------------------------------
# Synthetic scenario batches, BASELINE.md section 4
# --------------------------------------------------------------------------------------
SAMPLED_KEYS = ("j_BS", "w_dep", "w_marg", "w_sat", "tau_r", "rs", "a", "eta_CD", "tau_E0",
                "Lq", "B_pol", "tau_A0", "tau_w", "omega0")


def make_batch(config: int, S: Optional[int] = None, seed: Optional[int] = None, sample=None):
    """Returns ``(phys, x0[S,2], N)`` for BASELINE config 1..5, ``phys`` a dict name -> array[S].  ``sample`` =
    ``{name: (lo, hi)}``: extra entries (``Cw`` :19 "UNKNOWN!!", ``c_tauE``) drawn from PCG64(seed + 1), sorted names.

    RNG ``numpy.random.Generator(PCG64(seed))``, draws in scenario order (one row of uniforms per
    scenario: the 14 sampled constants, then w0, omega0, umax as the config uses them).  ``S``
    overrides the scenario count; a subsample is a prefix of the full batch."""
    base = default_physics()
    if config == 1:
        return {k: np.array([v]) for k, v in base.items()}, default_x0()[None, :].copy(), 3
    full = {2: (1024, 10), 3: (65536, 20), 4: (1048576, 20), 5: (16384, 100)}[config]
    S = full[0] if S is None else S
    N = full[1]
    rng = np.random.Generator(np.random.PCG64(full[0] if seed is None else seed))
    ndraw = {2: 1, 3: 16, 4: 17, 5: 1}[config]
    u = rng.random((S, ndraw))
    phys = {k: np.full(S, v) for k, v in base.items()}
    col = 0
    if config in (3, 4):
        for key in SAMPLED_KEYS:
            phys[key] = base[key] * (0.8 + (1.2 - 0.8) * u[:, col]); col += 1
    x0 = np.zeros((S, 2))
    x0[:, 0] = 0.06 + (0.15 - 0.06) * u[:, col]; col += 1
    if config in (3, 4):
        x0[:, 1] = 200 * math.pi + (4000 * math.pi - 200 * math.pi) * u[:, col]; col += 1
    else:
        x0[:, 1] = 2000 * math.pi
    if config == 4:
        phys["umax"] = 0.2e6 + (2e6 - 0.2e6) * u[:, col]; col += 1
    if sample:
        u2 = np.random.Generator(np.random.PCG64((full[0] if seed is None else seed) + 1)).random((S, len(sample)))
        for c2, key in enumerate(sorted(sample)):
            phys[key] = sample[key][0] + (sample[key][1] - sample[key][0]) * u2[:, c2]
    return phys, x0, N


def scenario(phys, s: int) -> Dict[str, float]:
    """Scalar physics dict of scenario ``s`` of a batch."""
    return {k: float(v[s]) for k, v in phys.items()}


def derive_params_batch(phys) -> np.ndarray:
    """[NPARAM, S] SoA parameter block (derive_params is elementwise, so it vectorises as is)."""
    out = derive_params({k: np.asarray(v, dtype=np.float64) for k, v in phys.items()})
    return np.ascontiguousarray(out.reshape(NPARAM, -1))


# --------------------------------------------------------------------------------------
# Monte-Carlo statistics (checker of ntm_mc_stats, include/ntm_mpc.h)
# --------------------------------------------------------------------------------------
MC_NBINS = 32
MC_NSTAT = 22 + MC_NBINS


def mc_stats(xk, uk, cost, status, umin, umax, bounds, w_suppressed=0.06, hist_max=0.2) -> np.ndarray:
    """xk [S,k_sim+1,2], uk [S,k_sim], cost [S], status [S], umin/umax scalars or [S],
    bounds = (xmin1, xmax1, xmin2, xmax2) -> the NTM_MC_NSTAT doubles documented in include/ntm_mpc.h."""
    xk = np.asarray(xk, dtype=np.float64); uk = np.asarray(uk, dtype=np.float64)
    S, K = uk.shape
    cost = np.zeros(S) if cost is None else np.asarray(cost, dtype=np.float64)
    status = np.zeros(S, dtype=np.int64) if status is None else np.asarray(status)
    umin = np.broadcast_to(np.asarray(umin, dtype=np.float64), (S,)); umax = np.broadcast_to(np.asarray(umax, dtype=np.float64), (S,))
    out = np.zeros(MC_NSTAT)
    out[6] = out[10] = np.inf; out[7] = out[11] = -np.inf
    out[0] = np.sum(status == 0); out[1] = np.sum(status == 1); out[2] = np.sum(status == 2); out[3] = np.sum(status == 3)
    ok = status < 2                                          # non-finite and infeasible scenarios are only counted
    if not ok.any():
        return out
    x = xk[ok]; u = uk[ok]; c = cost[ok]
    wf = x[:, K, 0]
    out[4] = c.sum(); out[5] = (c * c).sum(); out[6] = c.min(); out[7] = c.max()
    out[8] = wf.sum(); out[9] = (wf * wf).sum(); out[10] = wf.min(); out[11] = wf.max()
    out[12] = np.sum(wf < w_suppressed)
    below = x[:, 1:, 0] < w_suppressed
    reached = below.any(axis=1)
    out[13] = np.sum(np.argmax(below, axis=1)[reached] + 1); out[14] = reached.sum()
    out[15] = np.sum(u <= umin[ok][:, None]); out[16] = np.sum(u >= umax[ok][:, None]); out[17] = u.size; out[18] = u.sum()
    w = x[:, 1:, 0]; om = x[:, 1:, 1]
    out[19] = np.sum((w < bounds[0]) | (w > bounds[1])); out[20] = np.sum((om < bounds[2]) | (om > bounds[3])); out[21] = w.size
    bins = np.where(wf > 0, np.floor(wf / hist_max * MC_NBINS), 0).astype(np.int64)
    bins = np.clip(bins, 0, MC_NBINS - 1)
    out[22:] = np.bincount(bins, minlength=MC_NBINS)
    return out
