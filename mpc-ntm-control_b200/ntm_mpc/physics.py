"""Host-side parameter handling: the script's physics literals, the derived per-scenario coefficient
block the kernels consume, and the synthetic scenario batches of BASELINE.md section 4.

Reference lines: NTM_MPC_Sim.m:5-25 (literals, kappa, zeta), :31 (Ts), :37 (C), :47-50 (box), :59-60 (Q, r);
A.m:2 / B.m:2 (the coefficient products hoisted here in the reference's own evaluation order).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

from ._lib import NPARAM

PARAM_NAMES = ("c_a11", "c_a21", "a22", "c_b", "C1", "C2", "wmarg2", "w_dep",
               "umin", "umax", "r1", "r2", "q11", "q12", "q22", "c_tauE")

# constants the Monte-Carlo configs resample (config 3/4: nominal x U[0.8, 1.2])
SAMPLED = ("j_BS", "w_dep", "w_marg", "w_sat", "tau_r", "rs", "a", "eta_CD", "tau_E0",
           "Lq", "B_pol", "tau_A0", "tau_w", "omega0")


def nominal() -> Dict[str, float]:
    """NTM_MPC_Sim.m:5-22,31,47-48,59-60."""
    return dict(j_BS=73e3, w_dep=0.024, w_marg=0.02, w_sat=0.32, tau_r=293.0, rs=1.55, a=2.0, eta_CD=0.9,
                tau_E0=3.7, mu0=4e-7 * math.pi, Lq=0.87, B_pol=0.97, m=2.0, Cw=1.0, tau_A0=3e-6, tau_w=0.188,
                omega0=2 * math.pi * 420, Ts=0.1, umin=0.0, umax=2e6, r1=0.0, r2=1000 * 2 * math.pi,
                q11=1.0, q12=0.0, q22=1.0,
                c_tauE=0.0)   # tau_E(w) = tau_E0*(1 - c_tauE*w), read with PROFILE_TAUE_W only (:14 "NOT EXACT FORMULA")


def c_tauE_belt(p):
    """Coefficient of the tau_E(w) hook from the belt model of confinement degradation by an island [external knowledge:
    Chang & Callen, Nucl. Fusion 30 (1990) 219, d tau_E / tau_E = -4 w rs^3 / a^4], with the script's ``rs`` (:10) and
    ``a`` (:11): 0.93 per metre on the nominal physics.  Put it into ``p["c_tauE"]`` and set ``PROFILE_TAUE_W``."""
    return 4 * p["rs"] ** 3 / p["a"] ** 4


def x0_default() -> np.ndarray:
    """NTM_MPC_Sim.m:34."""
    return np.array([0.0, 1000 * 2 * math.pi])


def kappa(p):
    return 16 * p["mu0"] * p["Lq"] * p["rs"] ** 2 / (0.82 * p["tau_r"] * p["B_pol"] * math.pi)      # :24


def zeta(p):
    return p["m"] * p["Cw"] * p["tau_A0"] ** 2 * p["tau_w"] * p["a"] ** 3                            # :25


def affine_C(p):
    k = kappa(p)                                                                                     # :37
    return (-4 / 3 * (k * p["Ts"] * p["j_BS"] * p["w_sat"]) / (p["w_sat"] ** 2 + p["w_marg"] ** 2),
            p["Ts"] * p["omega0"] / p["tau_E0"])


def params_from_physics(p) -> np.ndarray:
    """[NPARAM] for a dict of scalars, [NPARAM, S] (SoA) for a dict of arrays."""
    k, z = kappa(p), zeta(p)
    C1, C2 = affine_C(p)
    tau_E = p["tau_E0"]                                                                              # :14
    vals = [(4 / 3) * (k * p["rs"] / (0.82 * p["tau_r"])) * p["Ts"],        # A.m:2, left to right
            p["Ts"] / (z * p["a"] ** 3),
            1 - p["Ts"] / tau_E,
            (k * p["Ts"] * p["eta_CD"] / p["w_dep"]),                       # B.m:2
            C1, C2, p["w_marg"] ** 2, p["w_dep"], p["umin"], p["umax"], p["r1"], p["r2"],
            p["q11"], p["q12"], p["q22"], p.get("c_tauE", 0.0)]
    shapes = [np.shape(v) for v in vals]
    if all(s == () for s in shapes):
        return np.array(vals, dtype=np.float64)
    S = max(s[0] for s in shapes if s != ())
    return np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), (S,)) for v in vals]))


def params_from_model_constants(kappa_, taur, Ts, zeta_, rs, a, TE, wdep, etaCD, wmarg=0.0, C=(0.0, 0.0),
                                umin=0.0, umax=0.0, r=(0.0, 0.0), Q=((1.0, 0.0), (0.0, 1.0))) -> np.ndarray:
    """Parameter block from the formal arguments of A.m:1 / B.m:1 (what the MATLAB call sites pass)."""
    return np.array([(4 / 3) * (kappa_ * rs / (0.82 * taur)) * Ts, Ts / (zeta_ * a ** 3), 1 - Ts / TE,
                     (kappa_ * Ts * etaCD / wdep), C[0], C[1], wmarg ** 2, wdep, umin, umax, r[0], r[1],
                     Q[0][0], Q[0][1], Q[1][1], 0.0], dtype=np.float64)


CONFIG_SHAPES = {1: (1, 3), 2: (1024, 10), 3: (65536, 20), 4: (1048576, 20), 5: (16384, 100)}


def make_batch(config: int, S: Optional[int] = None, seed: Optional[int] = None,
               sample: Optional[Dict[str, Tuple[float, float]]] = None) -> Tuple[dict, np.ndarray, int]:
    """Synthetic scenario batch of BASELINE.md section 4: ``(physics dict of arrays[S], x0[S,2], N)``.

    RNG numpy Generator(PCG64(seed)), one row of uniforms per scenario in scenario order, so a
    smaller S is a prefix of the full batch.  ``seed`` defaults to the config's scenario count.

    ``sample``: extra physics entries drawn uniformly per scenario, ``{name: (lo, hi)}`` in absolute units -- the hook for
    the constants the authors flag themselves: ``Cw`` (NTM_MPC_Sim.m:19 "UNKNOWN!!", enters zeta :25 and
    through it c_a21 = Ts/(zeta a^3)) and ``c_tauE`` (:14, with ``PROFILE_TAUE_W``).  Drawn from a SECOND generator
    (PCG64(seed + 1), one column per name in sorted order), so the base batch is unchanged by the option."""
    base = nominal()
    Sfull, N = CONFIG_SHAPES[config]
    if config == 1:
        return {k: np.array([v]) for k, v in base.items()}, x0_default()[None, :].copy(), N
    S = Sfull if S is None else S
    rng = np.random.Generator(np.random.PCG64(Sfull if seed is None else seed))
    ndraw = {2: 1, 3: 16, 4: 17, 5: 1}[config]
    u = rng.random((S, ndraw))
    phys = {k: np.full(S, v) for k, v in base.items()}
    col = 0
    if config in (3, 4):
        for key in SAMPLED:
            phys[key] = base[key] * (0.8 + (1.2 - 0.8) * u[:, col]); col += 1
    x0 = np.zeros((S, 2))
    x0[:, 0] = 0.06 + (0.15 - 0.06) * u[:, col]; col += 1                    # :40-41
    if config in (3, 4):
        x0[:, 1] = 200 * math.pi + (4000 * math.pi - 200 * math.pi) * u[:, col]; col += 1
    else:
        x0[:, 1] = 2000 * math.pi
    if config == 4:
        phys["umax"] = 0.2e6 + (2e6 - 0.2e6) * u[:, col]; col += 1
    if sample:
        rng2 = np.random.Generator(np.random.PCG64((Sfull if seed is None else seed) + 1))
        u2 = rng2.random((S, len(sample)))
        for c2, key in enumerate(sorted(sample)):
            if key not in base:
                raise KeyError(f"unknown physics entry {key!r}")
            lo, hi = sample[key]
            phys[key] = lo + (hi - lo) * u2[:, c2]
    return phys, x0, N


def batch_params(config: int, S: Optional[int] = None, seed: Optional[int] = None, sample=None):
    """``(params[NPARAM,S] SoA, x0[S,2], N)`` ready for the C ABI."""
    phys, x0, N = make_batch(config, S, seed, sample)
    P = params_from_physics(phys).reshape(NPARAM, -1)
    return np.ascontiguousarray(P), x0, N
