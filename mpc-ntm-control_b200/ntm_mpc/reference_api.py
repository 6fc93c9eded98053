"""The reference's own function names and signatures, executed on the GPU.

A MATLAB user of IsaacSavona/MPC-NTM-Control finds here the callables the script resolves by name --
``rho1``, ``rho2``, ``rho3`` (rho1.m, rho2.m, rho3.m), ``A``, ``B`` (A.m, B.m),
``Rho_to_PhiGammaLambda`` (Rho_to_PhiGammaLambda.m), a ``quadprog`` restricted to the input box
(NTM_MPC_Sim.m:97 with the rows of getWLc.m:14-23) and ``NTM_MPC_Sim`` (the script, as a function).
Same argument order and meaning; shapes follow MATLAB (x is [w; omega], B is the 2x1 column the
script needs, defect D7).  Scalars or batches: a leading scenario axis is accepted everywhere.

Like the MEX shims (csrc/mex/ntm_mex.c) the short call forms the script actually uses --
``rho1(x)``, ``A(r1, r2)``, ``B(r3)``, ``Rho_to_PhiGammaLambda(Rho1, Rho2, Rho3)`` -- are accepted
when a workspace has been bound with ``bind_workspace`` (the analogue of the caller-workspace lookup).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import physics
from ._lib import NPARAM, PROFILE_GAMMA_I, PROFILE_INNER_FIXED, PROFILE_RHO1_SQ, NtmError
from .api import NtmMpc

_handle: Optional[NtmMpc] = None
_workspace: Optional[dict] = None


def handle() -> NtmMpc:
    global _handle
    if _handle is None:
        _handle = NtmMpc(0)
    return _handle


def bind_workspace(ws: Optional[dict]) -> None:
    """Bind the script's workspace variables (kappa, tau_r, Ts, zeta, rs, a, tau_E, w_dep, eta_CD, w_marg, C)
    so the short call forms of NTM_MPC_Sim.m:63-66,113-119 resolve their missing arguments."""
    global _workspace
    _workspace = None if ws is None else dict(ws)


def workspace_from_physics(p: dict) -> dict:
    C1, C2 = physics.affine_C(p)
    return dict(kappa=physics.kappa(p), tau_r=p["tau_r"], Ts=p["Ts"], zeta=physics.zeta(p), rs=p["rs"], a=p["a"],
                tau_E=p["tau_E0"], w_dep=p["w_dep"], eta_CD=p["eta_CD"], w_marg=p["w_marg"], C=np.array([C1, C2]))


def _ws(name):
    if _workspace is None or name not in _workspace:
        raise TypeError(f"not enough input arguments: '{name}' missing and no workspace bound (bind_workspace)")
    return _workspace[name]


def _blank_params(**kw) -> np.ndarray:
    p = np.zeros(NPARAM)
    p[7] = 1.0            # w_dep (avoid 0/0 in unused rho3)
    p[12] = p[14] = 1.0
    names = physics.PARAM_NAMES
    for k, v in kw.items():
        p[names.index(k)] = v
    return p


def _x2(x):
    x = np.asarray(x, dtype=np.float64)
    single = x.ndim == 1 or (x.ndim == 2 and x.shape[1] == 1 and x.shape[0] == 2)
    return (x.reshape(1, 2) if single else x.reshape(-1, 2)), single


# ---------------------------------------------------------------------------------------- rho1.m, rho2.m, rho3.m
def rho1(x, wmarg=None, variant: str = "lin"):
    """rho1.m:1-3 ``rho1(x, wmarg)``; ``variant='sq'`` is the draft in rhos.m:17-19."""
    wmarg = _ws("w_marg") if wmarg is None else wmarg
    X, single = _x2(x)
    r1, _, _ = handle().rho(np.column_stack([X[:, 0], np.ones(len(X))]), _blank_params(wmarg2=wmarg ** 2),
                            PROFILE_RHO1_SQ if variant == "sq" else 0)
    return float(r1[0]) if single else r1


def rho2(x):
    """rho2.m:1-3."""
    X, single = _x2(x)
    _, r2, _ = handle().rho(X, _blank_params(wmarg2=1.0))
    return float(r2[0]) if single else r2


def rho3(x, w_dep=None):
    """rho3.m:1-4."""
    w_dep = _ws("w_dep") if w_dep is None else w_dep
    X, single = _x2(x)
    _, _, r3 = handle().rho(np.column_stack([X[:, 0], np.ones(len(X))]), _blank_params(wmarg2=1.0, w_dep=w_dep))
    return float(r3[0]) if single else r3


# ---------------------------------------------------------------------------------------- A.m, B.m
class LpvA:
    """Callable ``A(rho1, rho2)`` closed over the constants of A.m:1 -- what the script passes around."""

    def __init__(self, kappa, taur, Ts, zeta, rs, a, TE):
        self.consts = (kappa, taur, Ts, zeta, rs, a, TE)

    def __call__(self, r1, r2):
        return A(r1, r2, *self.consts)


class LpvB:
    """Callable ``B(rho3)`` closed over the constants of B.m:1."""

    def __init__(self, wdep, kappa, Ts, etaCD):
        self.consts = (wdep, kappa, Ts, etaCD)

    def __call__(self, r3):
        return B(r3, *self.consts)


def A(rho1_, rho2_, kappa=None, taur=None, Ts=None, zeta=None, rs=None, a=None, TE=None):
    """A.m:1-3.  Returns the 2x2 matrix (or [S,2,2])."""
    if kappa is None:
        kappa, taur, Ts, zeta, rs, a, TE = (_ws(k) for k in ("kappa", "tau_r", "Ts", "zeta", "rs", "a", "tau_E"))
    single = np.ndim(rho1_) == 0
    r1 = np.atleast_1d(np.asarray(rho1_, dtype=np.float64)); r2 = np.atleast_1d(np.asarray(rho2_, dtype=np.float64))
    prm = physics.params_from_model_constants(kappa, taur, Ts, zeta, rs, a, TE, wdep=1.0, etaCD=0.0)
    Am, _ = handle().lpv_AB(r1, r2, np.zeros_like(r1), prm)
    return Am[0].copy() if single else Am


def B(rho3_, wdep=None, kappa=None, Ts=None, etaCD=None):
    """B.m:1-3.  Returns the 2x1 column [b; 0] (or [S,2]); the .m file's 1x2 row cannot execute at
    NTM_MPC_Sim.m:113 (defect D7)."""
    if wdep is None:
        wdep, kappa, Ts, etaCD = (_ws(k) for k in ("w_dep", "kappa", "Ts", "eta_CD"))
    single = np.ndim(rho3_) == 0
    r3 = np.atleast_1d(np.asarray(rho3_, dtype=np.float64))
    prm = physics.params_from_model_constants(kappa, 1.0, Ts, 1.0, 1.0, 1.0, 1.0, wdep=wdep, etaCD=etaCD)
    _, Bm = handle().lpv_AB(np.zeros_like(r3), np.zeros_like(r3), r3, prm)
    return Bm[0].copy() if single else Bm


# ---------------------------------------------------------------------------------------- Rho_to_PhiGammaLambda.m
def Rho_to_PhiGammaLambda(Rho1, Rho2, Rho3, A=None, B=None, C=None, gamma_index: str = "i_minus_j"):
    """Rho_to_PhiGammaLambda.m:1.  ``A``/``B`` must be :class:`LpvA` / :class:`LpvB` (the GPU cannot call
    arbitrary host callables -- and there is no CPU fallback); row or column rho vectors both work
    (N = numel, repair of D4); a leading scenario axis batches."""
    if A is None:
        A = LpvA(*(_ws(k) for k in ("kappa", "tau_r", "Ts", "zeta", "rs", "a", "tau_E")))
    if B is None:
        B = LpvB(*(_ws(k) for k in ("w_dep", "kappa", "Ts", "eta_CD")))
    if C is None:
        C = _ws("C")
    if not isinstance(A, LpvA) or not isinstance(B, LpvB):
        raise TypeError("A and B must be ntm_mpc.LpvA / ntm_mpc.LpvB instances")
    R1 = np.asarray(Rho1, dtype=np.float64); single = R1.ndim == 1 or 1 in R1.shape[:2] and R1.ndim == 2
    if single:
        R1 = R1.reshape(1, -1); R2 = np.asarray(Rho2, dtype=np.float64).reshape(1, -1); R3 = np.asarray(Rho3, dtype=np.float64).reshape(1, -1)
    else:
        R2 = np.asarray(Rho2, dtype=np.float64); R3 = np.asarray(Rho3, dtype=np.float64)
    kappa, taur, Ts, zeta, rs, a, TE = A.consts
    wdep, kappa_b, Ts_b, etaCD = B.consts
    if kappa_b != kappa or Ts_b != Ts:
        raise ValueError("A and B were built from different kappa/Ts")
    C = np.asarray(C, dtype=np.float64).ravel()
    prm = physics.params_from_model_constants(kappa, taur, Ts, zeta, rs, a, TE, wdep=wdep, etaCD=etaCD, C=(C[0], C[1]))
    Phi, Gam, Lam = handle().condense(R1, R2, R3, prm, PROFILE_GAMMA_I if gamma_index == "i" else 0)
    if single:
        return Phi[0].copy(), Gam[0].copy(), Lam[0].copy()
    return Phi, Gam, Lam


# ---------------------------------------------------------------------------------------- getWLc.m
def getWLc(xmax, xmin, umax, umin, Gamma, Phi, Lambda):
    """getWLc.m:1 ``[W, L, c] = getWLc(xmax, xmin, umax, umin, Gamma, Phi, Lambda)`` (the 7-argument signature; the
    script's 9-argument call at NTM_MPC_Sim.m:74 cannot execute, defect D9).  A leading scenario axis on Gamma/Phi/
    Lambda batches."""
    single = np.ndim(Gamma) == 2
    W, L, c = handle().getWLc(xmax, xmin, umax, umin, Gamma, Phi, Lambda)
    if single:
        return W[0].copy(), L[0].copy(), c[0].copy()
    return W, L, c


# ---------------------------------------------------------------------------------------- quadprog (box rows)
def split_rows(A, b):
    """Rows of ``A U <= b`` ([S,M,N], [S,M]) by pattern over the batch: one non-zero column -> a bound on that
    variable, none -> a feasibility statement on b, otherwise a general row.  getWLc.m's L has the same pattern for
    every scenario: 2N bound rows (:14-23), 4 empty rows (the x_0 block, :30) and 4N state rows.
    Returns ``(lb [S,N], ub [S,N], Lg [S,Mg,N], bg [S,Mg], feasible [S])``."""
    A = np.asarray(A, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    S, M, N = A.shape
    pattern = np.any(A != 0.0, axis=0)                       # [M, N]
    nnz = pattern.sum(axis=1)
    lb = np.full((S, N), -np.inf); ub = np.full((S, N), np.inf)
    feasible = np.ones(S, dtype=bool)
    for i in np.flatnonzero(nnz == 0):
        feasible &= b[:, i] >= 0.0
    for i in np.flatnonzero(nnz == 1):
        j = int(np.flatnonzero(pattern[i])[0])
        a = A[:, i, j]
        with np.errstate(divide="ignore", invalid="ignore"):
            v = b[:, i] / a
        ub[:, j] = np.where(a > 0, np.minimum(ub[:, j], v), ub[:, j])
        lb[:, j] = np.where(a < 0, np.maximum(lb[:, j], v), lb[:, j])
        feasible &= ~((a == 0) & (b[:, i] < 0))
    gen = np.flatnonzero(nnz > 1)
    return lb, ub, np.ascontiguousarray(A[:, gen, :]), np.ascontiguousarray(b[:, gen]), feasible


def quadprog(H, f, A=None, b=None, Aeq=None, beq=None, lb=None, ub=None, x0=None, options=None):
    """``quadprog(G, F, L, c + W*xk(:,k), [], [], [], [], [], opt)`` -- NTM_MPC_Sim.m:97, MATLAB argument order.

    The rows of ``A`` are split on the host (`split_rows`): the input-box rows of getWLc.m:14-23 become ``lb/ub``
    (merged with any explicit ``lb/ub``), the x_0 rows of getWLc.m:30 decide feasibility outright (defect D18), the
    state rows go to the GPU as general rows (``ntm_qp_ineq``).  Every variable needs finite bounds.  Returns
    ``(U, fval, exitflag)``, exitflag 1 = minimiser, 0 = iteration cap, -2 = infeasible, -3 = non-finite data.
    A leading scenario axis on H (f, A, b) solves a batch."""
    if (Aeq is not None and np.size(Aeq)) or (beq is not None and np.size(beq)):
        raise NotImplementedError("equality constraints: NTM_MPC_Sim.m:97 passes [] for Aeq, beq")
    H = np.asarray(H, dtype=np.float64); single = H.ndim == 2
    Hb = H[None] if single else H
    S, N, _ = Hb.shape
    fb = np.asarray(f, dtype=np.float64).reshape(S, N)
    lo = np.full((S, N), -np.inf) if lb is None or np.size(lb) == 0 else np.broadcast_to(np.asarray(lb, dtype=np.float64), (S, N)).copy()
    hi = np.full((S, N), np.inf) if ub is None or np.size(ub) == 0 else np.broadcast_to(np.asarray(ub, dtype=np.float64), (S, N)).copy()
    feasible = np.ones(S, dtype=bool)
    Lg = np.zeros((S, 0, N)); bg = np.zeros((S, 0))
    if A is not None and np.size(A):
        Ab = np.asarray(A, dtype=np.float64)
        Ab = np.broadcast_to(Ab, (S,) + Ab.shape[-2:])
        bb = np.broadcast_to(np.asarray(b, dtype=np.float64).reshape(-1, Ab.shape[1]), (S, Ab.shape[1]))
        l2, u2, Lg, bg, feasible = split_rows(Ab, bb)
        lo = np.maximum(lo, l2); hi = np.minimum(hi, u2)
    if not (np.all(np.isfinite(lo)) and np.all(np.isfinite(hi))):
        raise ValueError("quadprog shim: every variable needs finite lower and upper bounds (getWLc.m:14-23 provides them)")
    feasible &= np.all(lo <= hi, axis=1)
    hi = np.maximum(hi, lo)
    if Lg.shape[1]:
        U, it, st = handle().qp_ineq(Hb, fb, lo, hi, Lg, bg)
    else:
        U, it, st = handle().qp_box(Hb, fb, lo, hi)
    fval = 0.5 * np.einsum("si,sij,sj->s", U, Hb, U) + np.einsum("si,si->s", fb, U)
    flag = np.where(st == 0, 1, np.where(st == 1, 0, np.where(st == 3, -2, -3)))
    flag = np.where(feasible, flag, -2)
    if single:
        return U[0].copy(), float(fval[0]), int(flag[0])
    return U, fval, flag


# ---------------------------------------------------------------------------------------- the script
def NTM_MPC_Sim(physics_constants: Optional[dict] = None, x0=None, N: int = 3, k_sim: int = 20, i_sim: int = 10,
                epsilon: float = 1e-14, profile: int = 0, inner_policy: str = "eps_break", state_rows: int = 0,
                xmin=None, xmax=None):
    """NTM_MPC_Sim.m as a function: returns ``(xk [2, k_sim+1], uk [1, k_sim], Uk [N, k_sim])`` -- the variables
    the script leaves in the base workspace (:82-84).  A leading scenario axis on ``x0`` (and arrays in
    ``physics_constants``) runs a batch and returns [S, ...] arrays.  ``state_rows`` = 1 / 2 keeps getWLc's state rows
    in the QP of :97 (rebuilt at every :119 / frozen at :74 as the script has it), ``xmin``/``xmax`` default to :44-45;
    an infeasible QP (exitflag -2) leaves NaN from that step on, as the script cannot continue without a U."""
    p = physics.nominal() if physics_constants is None else physics_constants
    x0 = physics.x0_default() if x0 is None else np.asarray(x0, dtype=np.float64)
    single = x0.ndim == 1
    X0 = x0.reshape(-1, 2)
    prm = physics.params_from_physics(p)
    prm = prm if prm.ndim == 1 else np.ascontiguousarray(prm.T)
    if inner_policy == "fixed":
        profile |= PROFILE_INNER_FIXED
    xb = None
    if state_rows:
        lo = (0.06, 100 * 2 * np.pi) if xmin is None else np.asarray(xmin, dtype=np.float64).ravel()
        hi = (0.15, 5000 * 2 * np.pi) if xmax is None else np.asarray(xmax, dtype=np.float64).ravel()
        xb = (lo[0], hi[0], lo[1], hi[1])
    r = handle().closed_loop(X0, prm, N, k_sim, i_sim, epsilon, profile, want_Uk=True, state_rows=state_rows, xbounds=xb)
    xk = r["xk"].transpose(0, 2, 1); uk = r["uk"][:, None, :]; Uk = r["Uk"].transpose(0, 2, 1)
    if single:
        return xk[0], uk[0], Uk[0]
    return xk, uk, Uk
