"""Randomised parity sweep: horizons, profiles, loop lengths and batch sizes against the C oracle (literal readings are
held to 1e-6 per trajectory; readings with the `i` index to the 99 % quantile)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np
import ntm_mpc
from oracle import c_oracle as co, ntm_oracle as o

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
mpc = ntm_mpc.NtmMpc(0)
worst = 0.0
for t in range(trials):
    cfg = int(rng.choice([2, 3, 4]))
    N = int(rng.choice([1, 2, 3, 5, 8, 10, 13, 16, 20, 24, 31, 32, 33, 40, 48]))
    S = int(rng.integers(1, 200))
    k_sim = int(rng.integers(1, 12)); i_sim = int(rng.integers(1, 11))
    flags = 0
    if rng.random() < 0.3: flags |= 1          # rho1 squared
    if rng.random() < 0.25: flags |= 2         # gamma index i
    if rng.random() < 0.3: flags |= 4          # F from xk
    if rng.random() < 0.3: flags |= 8          # plant + C
    if rng.random() < 0.5: flags |= 16         # fixed inner policy
    if rng.random() < 0.15 and not (flags & 2): flags |= 32   # dense-G cross-check path
    if rng.random() < 0.25: flags |= 64        # RK4 plant (EXT instantiation of the fused kernel)
    seed = int(rng.integers(1, 1 << 30))
    phys, x0, _ = o.make_batch(cfg, S=S, seed=seed)
    P = o.derive_params_batch(phys)
    g = mpc.closed_loop(x0, P.T, N=N, k_sim=k_sim, i_sim=i_sim, profile=flags)
    c = co.closed_loop_batch(phys, x0, N, k_sim=k_sim, i_sim=i_sim, flags=flags & (31 | 64))
    umax = phys["umax"]
    du = np.max(np.abs(g["uk"] - c["uk"]), axis=1) / umax
    w = c["xk"][:, :, 0]
    dw = np.max(np.abs(g["xk"][:, :, 0] - w), axis=1) / np.maximum(np.max(np.abs(w), axis=1), 1e-3)
    finite = np.isfinite(c["xk"]).all(axis=(1, 2)) & (c["status"] == 0)
    bad = ((du > 1e-6) | (dw > 1e-6)) & finite
    frac = bad.mean()
    strict = not (flags & (2 | 4 | 8)) and N <= 20   # the non-literal F / plant / Gamma readings are chaotic for a few scenarios:
                                                # there the C and NumPy oracles already disagree with each other (tools/diag_trial.py)
    ok = (frac == 0.0) if strict else (bad.sum() <= max(1, 0.03 * S))
    nf_match = np.array_equal(np.isfinite(g["xk"]).all(axis=(1, 2)), np.isfinite(c["xk"]).all(axis=(1, 2)))
    worst = max(worst, float(np.max(np.where(finite, np.maximum(du, dw), 0.0))) if strict else 0.0)
    print(f"trial {t:3d} cfg{cfg} N={N:2d} S={S:3d} k={k_sim:2d} i={i_sim:2d} flags={flags:2d}: bad {frac:.3f} "
          f"gpu_status_max {int(g['status'].max())} oracle_status_max {int(c['status'].max())} nonfinite_match {nf_match} {'OK' if ok else 'FAIL'}")
    if not ok:
        i = int(np.argmax(bad)); print("   first bad scenario", i, "du", du[i], "dw", dw[i], "gpu uk", g["uk"][i][:6], "oracle uk", c["uk"][i][:6])
print("worst strict error", worst)
