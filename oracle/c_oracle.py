"""ctypes binding of oracle/libntm_oracle.so (TEST INFRASTRUCTURE ONLY; see ntm_oracle.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libntm_oracle.so")

PHYS_ORDER = ("j_BS", "w_dep", "w_marg", "w_sat", "tau_r", "rs", "a", "eta_CD", "tau_E0", "mu0", "Lq",
              "B_pol", "m", "Cw", "tau_A0", "tau_w", "omega0", "Ts", "umin", "umax", "r1", "r2",
              "q11", "q12", "q22", "c_tauE")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ntm_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libntm_oracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.ntm_oracle_closed_loop_batch.restype = ctypes.c_int
        _lib.ntm_oracle_qp_box.restype = ctypes.c_int
        _lib.ntm_oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def phys_block(phys) -> np.ndarray:
    """[S,26] row-major physics block from a dict name -> array[S] (or scalars); a missing c_tauE is 0."""
    cols = [np.atleast_1d(np.asarray(phys[k] if k in phys else 0.0, dtype=np.float64)) for k in PHYS_ORDER]
    S = max(c.size for c in cols)
    return np.ascontiguousarray(np.stack([np.broadcast_to(c, (S,)) for c in cols], axis=1))


def condense(phys, Rho1, Rho2, Rho3, flags=0):
    pb = phys_block(phys)[0].copy()
    R1, R2, R3 = (np.ascontiguousarray(r, dtype=np.float64).ravel() for r in (Rho1, Rho2, Rho3))
    N = R1.size
    Phi = np.zeros((2, 2 * N)); Gam = np.zeros((N, 2 * N)); Lam = np.zeros(2 * N)     # column-major storage
    lib().ntm_oracle_condense(_p(pb), ctypes.c_int(N), _p(R1), _p(R2), _p(R3), ctypes.c_int(flags), _p(Phi), _p(Gam), _p(Lam))
    return Phi.T.copy(), Gam.T.copy(), Lam


def hessian_grad(Phi, Gamma, Lambda, x, r, Q):
    N = Gamma.shape[1]
    Phi_c = np.ascontiguousarray(np.asarray(Phi, dtype=np.float64).T); Gam_c = np.ascontiguousarray(np.asarray(Gamma, dtype=np.float64).T)
    Lam = np.ascontiguousarray(Lambda, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64); r = np.ascontiguousarray(r, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    Q3 = np.array([Q[0, 0], Q[0, 1], Q[1, 1]])
    G = np.zeros((N, N)); F = np.zeros(N)
    lib().ntm_oracle_hessian_grad(ctypes.c_int(N), _p(Phi_c), _p(Gam_c), _p(Lam), _p(x), _p(r), _p(Q3), _p(G), _p(F))
    return G, F


def qp_box(G, F, lb, ub):
    G = np.ascontiguousarray(G, dtype=np.float64); F = np.ascontiguousarray(F, dtype=np.float64)
    N = F.size
    lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (N,)))
    ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (N,)))
    U = np.zeros(N); it = ctypes.c_int(0); work = np.zeros(2 * N * N + 8 * N + 8)
    st = lib().ntm_oracle_qp_box(ctypes.c_int(N), _p(G), _p(F), _p(lb), _p(ub), _p(U), ctypes.byref(it), _p(work))
    return U, it.value, st


def closed_loop_batch(phys, x0, N, k_sim=20, i_sim=10, eps=1e-14, flags=0, threads=0, want_Uk=False):
    """Runs S scenarios; returns dict of arrays in scenario-slowest layout plus 'threads'."""
    pb = phys_block(phys)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, 2)
    S = x0.shape[0]
    if pb.shape[0] == 1 and S > 1:
        pb = np.ascontiguousarray(np.broadcast_to(pb, (S, pb.shape[1])))
    xk = np.zeros((S, k_sim + 1, 2)); uk = np.zeros((S, k_sim)); cost = np.zeros(S)
    Uk = np.zeros((S, k_sim, N)) if want_Uk else None
    inner = np.zeros((S, k_sim), dtype=np.int32); qpit = np.zeros((S, k_sim), dtype=np.int32)
    status = np.zeros(S, dtype=np.int32)
    used = lib().ntm_oracle_closed_loop_batch(
        ctypes.c_int(S), ctypes.c_int(N), ctypes.c_int(k_sim), ctypes.c_int(i_sim), ctypes.c_double(eps),
        ctypes.c_int(flags), _p(pb), _p(x0), _p(xk), _p(uk), _p(Uk) if want_Uk else None,
        inner.ctypes.data_as(_ip), qpit.ctypes.data_as(_ip), _p(cost), status.ctypes.data_as(_ip), ctypes.c_int(threads))
    if used < 0:
        raise ValueError("N exceeds NTM_ORACLE_MAXN")
    return dict(xk=xk, uk=uk, Uk=Uk, inner_iters=inner, qp_iters=qpit, cost=cost, status=status, threads=used)


def max_threads() -> int:
    return lib().ntm_oracle_max_threads()
