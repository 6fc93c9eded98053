#!/usr/bin/env python
"""Adjudicate a scenario on which the implementations disagree: the closed loop of the repaired
script in 50-digit arithmetic (mpmath), with every QP solved EXACTLY -- the active set is carried from QP to QP and
re-certified through its KKT conditions in 50 digits, pivoting until they hold -- next to the NumPy oracle, the C oracle
and (when a file with the GPU result is given) the CUDA kernels.  Prints, per implementation, the first MPC step at which
uk leaves the 50-digit trajectory by more than 1e-6 of umax and the largest deviation: "which one is nearest".

    python tools/adjudicate.py <config> <N> <S> <scenario> [k_sim] [i_sim] [flags] [gpu.npz]

CPU only (test/diagnosis infrastructure: imports oracle/).  One N = 33 scenario with k_sim = 8 takes about a minute."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpmath as mp
import numpy as np

from oracle import c_oracle, ntm_oracle as o

mp.mp.dps = 50


def mp_rho(p, x, sq):
    w, om = x
    r1 = 1 / (w * w + p["w_marg"] ** 2) if sq else 1 / (w + p["w_marg"] ** 2)      # rho1.m:2 | rhos.m:18
    r2 = (w * w) / om                                                               # rho2.m:2
    ws = w / p["w_dep"]
    r3 = (mp.mpf("0.25") + mp.mpf("0.24") * ws) / (1 + mp.mpf("1.5") * ws + mp.mpf("0.43") * ws ** 2 + mp.mpf("0.64") * ws ** 3)
    return r1, r2, r3


def mp_model(p):
    kappa = 16 * p["mu0"] * p["Lq"] * p["rs"] ** 2 / (mp.mpf("0.82") * p["tau_r"] * p["B_pol"] * mp.pi)      # NTM_MPC_Sim.m:24
    zeta = p["m"] * p["Cw"] * p["tau_A0"] ** 2 * p["tau_w"] * p["a"] ** 3                                       # :25
    ca11 = (mp.mpf(4) / 3) * (kappa * p["rs"] / (mp.mpf("0.82") * p["tau_r"])) * p["Ts"]                        # A.m:2
    ca21 = p["Ts"] / (zeta * p["a"] ** 3)
    a22 = 1 - p["Ts"] / p["tau_E0"]
    cb = kappa * p["Ts"] * p["eta_CD"] / p["w_dep"]                                                             # B.m:2
    C = (-(mp.mpf(4) / 3) * (kappa * p["Ts"] * p["j_BS"] * p["w_sat"]) / (p["w_sat"] ** 2 + p["w_marg"] ** 2),
         p["Ts"] * p["omega0"] / p["tau_E0"])                                                                   # :37
    return ca11, ca21, a22, cb, C


def mp_condense_GF(rho, model, xF, p, gamma_i):
    """G, F of NTM_MPC_Sim.m:120-121 from the rho sequences (Rho_to_PhiGammaLambda.m:17-52), all in mp."""
    ca11, ca21, a22, cb, C = model
    N = len(rho)
    A = [(ca11 * r[0] + 1, ca21 * r[1]) for r in rho]             # (a11, a21); a12 = 0, a22 shared
    b = [cb * r[2] for r in rho]
    # columns of Gamma: block(i,j) = A_k block(i-1,j), k = i-j (literal) or i
    Gam = [[None] * N for _ in range(N)]
    for jc in range(N):
        g1, g2 = b[jc], mp.mpf(0)
        Gam[jc][jc] = (g1, g2)
        for i in range(jc + 1, N):
            k = i if gamma_i else i - jc - 1
            a11, a21 = A[k]
            g1, g2 = a11 * g1, a21 * g1 + a22 * g2
            Gam[i][jc] = (g1, g2)
    v1, v2 = xF
    E = []
    for i in range(N):
        a11, a21 = A[i]
        v1, v2 = a11 * v1 + C[0], a21 * v1 + a22 * v2 + C[1]
        E.append((v1 - p["r1"], v2 - p["r2"]))
    q11, q12, q22 = p["q11"], p["q12"], p["q22"]
    G = mp.zeros(N, N); F = mp.zeros(N, 1)
    for jc in range(N):
        for l in range(jc + 1):
            s = mp.mpf(0)
            for i in range(jc, N):
                x1, x2 = Gam[i][jc]; y1, y2 = Gam[i][l]
                s += x1 * (q11 * y1 + q12 * y2) + x2 * (q12 * y1 + q22 * y2)
            G[jc, l] = 2 * s; G[l, jc] = 2 * s
        s = mp.mpf(0)
        for i in range(jc, N):
            x1, x2 = Gam[i][jc]; e1, e2 = E[i]
            s += x1 * (q11 * e1 + q12 * e2) + x2 * (q12 * e1 + q22 * e2)
        F[jc] = 2 * s
    return G, F, A, b


def mp_qp(G, F, lb, ub, state):
    """Exact minimiser of 1/2 U'GU + F'U on the box by a primal active-set method in mp, started from partition
    `state` (-1 / 0 / +1 per variable).  Returns U and the final partition."""
    N = G.rows
    U = [lb if s < 0 else (ub if s > 0 else (lb + ub) / 2) for s in state]
    state = list(state)
    for _ in range(20 * N + 50):
        free = [i for i in range(N) if state[i] == 0]
        g = [F[i] + sum(G[i, k] * U[k] for k in range(N)) for i in range(N)]
        if free:
            H = mp.matrix(len(free), len(free)); rhs = mp.matrix(len(free), 1)
            for a, i in enumerate(free):
                rhs[a] = -g[i]
                for c, k in enumerate(free):
                    H[a, c] = G[i, k]
            pf = mp.lu_solve(H, rhs)
            alpha, blk, bst = mp.mpf(1), -1, 0
            for a, i in enumerate(free):
                if pf[a] < 0:
                    t = (lb - U[i]) / pf[a]
                    if t < alpha: alpha, blk, bst = t, i, -1
                elif pf[a] > 0:
                    t = (ub - U[i]) / pf[a]
                    if t < alpha: alpha, blk, bst = t, i, 1
            for a, i in enumerate(free):
                U[i] += alpha * pf[a]
            if blk >= 0:
                state[blk] = bst; U[blk] = lb if bst < 0 else ub
                continue
            g = [F[i] + sum(G[i, k] * U[k] for k in range(N)) for i in range(N)]
        worst, wi = mp.mpf(0), -1
        for i in range(N):
            lam = g[i] if state[i] < 0 else (-g[i] if state[i] > 0 else mp.mpf(0))
            if lam < worst: worst, wi = lam, i
        if wi < 0:
            return U, state
        state[wi] = 0
    raise RuntimeError("mp active set did not terminate")


def mp_closed_loop(pf64, x0, N, k_sim, i_sim, flags, log=None):
    p = {k: mp.mpf(float(v)) for k, v in pf64.items()}            # the SAME fp64 inputs, exactly
    model = mp_model(p)
    ca11, ca21, a22, cb, C = model
    sq, gi, fxk, plc, fixed = bool(flags & 1), bool(flags & 2), bool(flags & 4), bool(flags & 8), bool(flags & 16)
    x = (mp.mpf(float(x0[0])), mp.mpf(float(x0[1])))
    x0m = x
    rho = [mp_rho(p, x, sq)] * N                                   # :63-65
    G, F, A, b = mp_condense_GF(rho, model, x0m, p, gi)
    Uold = [mp.mpf(1)] * N
    state = [-1] * N
    uk, xk, inner = [], [x], []
    lb, ub = p["umin"], p["umax"]
    eps = mp.mpf("1e-14")
    for k in range(k_sim):
        for it in range(1, i_sim + 1):
            U, state = mp_qp(G, F, lb, ub, state)                  # :97
            u0 = U[0]
            xs = [x]
            for i in range(N):                                     # :110-117 (old rho), then re-schedule
                a11, a21 = A[i]
                z = xs[-1]
                xs.append((a11 * z[0] + b[i] * U[i] + C[0], a21 * z[0] + a22 * z[1] + C[1]))
            rho = [mp_rho(p, xs[i], sq) for i in range(N)]
            G, F, A, b = mp_condense_GF(rho, model, x if fxk else x0m, p, gi)      # :119-121
            brk = (not fixed) and sum(abs(Uold[i] - U[i]) for i in range(N)) < eps  # :123
            if brk:
                break
            Uold = U
        inner.append(it)
        r1, r2, r3 = mp_rho(p, x, sq)                              # :130
        x = ((ca11 * r1 + 1) * x[0] + cb * r3 * u0 + (C[0] if plc else 0), ca21 * r2 * x[0] + a22 * x[1] + (C[1] if plc else 0))
        uk.append(u0); xk.append(x)
        if log:
            log(f"   mp step {k}: u = {mp.nstr(u0, 12)}  w = {mp.nstr(x[0], 12)}  inner {it}")
    return np.array([float(u) for u in uk]), np.array([[float(a), float(b_)] for a, b_ in xk]), np.array(inner)


def main():
    cfg, N, S, s = (int(v) for v in sys.argv[1:5])
    k_sim = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    i_sim = int(sys.argv[6]) if len(sys.argv) > 6 else 10
    flags = int(sys.argv[7]) if len(sys.argv) > 7 else 16
    gpu = np.load(sys.argv[8]) if len(sys.argv) > 8 else None
    phys, x0, _ = o.make_batch(cfg, S=S)
    p = o.scenario(phys, s)
    t0 = time.time()
    uk_mp, xk_mp, inner_mp = mp_closed_loop(p, x0[s], N, k_sim, i_sim, flags, log=print if "-v" in sys.argv else None)
    print(f"50-digit closed loop: {time.time() - t0:.1f} s, inner iterations {inner_mp.tolist()}")
    prof = o.Profile(rho1_variant=flags & 1, gamma_index=(flags >> 1) & 1, f_state=(flags >> 2) & 1, plant_affine=(flags >> 3) & 1,
                     inner_policy=(flags >> 4) & 1) if hasattr(o, "Profile") else None
    impls = {}
    try:
        r = o.closed_loop(p, x0[s], N=N, k_sim=k_sim, i_sim=i_sim, profile=prof)
        impls["numpy oracle"] = (r["uk"], r["xk"].T)
    except Exception as exc:                                       # profile constructor differs: report and go on
        print("numpy oracle skipped:", exc)
    sub = {k: np.asarray(v)[s:s + 1] for k, v in phys.items()}
    c = c_oracle.closed_loop_batch(sub, x0[s:s + 1], N, k_sim, i_sim, 1e-14, flags & 31, 1)
    impls["C oracle"] = (c["uk"][0], c["xk"][0])
    if gpu is not None:
        impls["CUDA"] = (gpu["uk"][s], gpu["xk"][s])
    umax = float(p["umax"])
    wref = max(np.max(np.abs(xk_mp[:, 0])), 1e-3)
    print(f"config{cfg} N={N} scenario {s} k_sim={k_sim} i_sim={i_sim} flags={flags}")
    print("  mp uk:", np.array2string(uk_mp, precision=6))
    for name, (uk, xk) in impls.items():
        du = np.abs(uk - uk_mp) / umax
        dw = np.abs(xk[:, 0] - xk_mp[:, 0]) / wref
        first = int(np.argmax(du > 1e-6)) if (du > 1e-6).any() else -1
        print(f"  {name:13s}: max |du|/umax = {du.max():.3e}  max |dw|/w = {dw.max():.3e}  first step off by > 1e-6: {first}")
        print("               uk:", np.array2string(uk, precision=6))


if __name__ == "__main__":
    main()
