"""Generates tests/golden/closed_loop_*.npz from the NumPy oracle (oracle/ntm_oracle.py).

The reference is MATLAB and cannot run here (no MATLAB/Octave), so these fixtures are
outputs of the oracle restatement, pinned in turn by appendix_a.json + the mpmath twin in
tests/test_oracle.py.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ntm_oracle as o  # noqa: E402

CASES = [  # (config, S)
    (1, 1), (2, 24), (3, 24), (4, 24), (5, 2),
]
PROFILES = {"literal_fixed": o.LITERAL_FIXED, "literal": o.LITERAL, "consistent_fixed": o.CONSISTENT_FIXED}


def main():
    for cfg, S in CASES:
        phys, x0, N = o.make_batch(cfg, S=S)
        out = dict(config=cfg, S=S, N=N, x0=x0, params=o.derive_params_batch(phys))
        for k, v in phys.items():
            out["phys_" + k] = v
        for name, prof in PROFILES.items():
            if cfg == 5 and name != "literal_fixed":
                continue
            xk = np.zeros((S, 21, 2)); uk = np.zeros((S, 20)); Uk = np.zeros((S, 20, N))
            inner = np.zeros((S, 20), dtype=np.int32); cost = np.zeros(S); status = np.zeros(S, dtype=np.int32)
            for s in range(S):
                r = o.closed_loop(o.scenario(phys, s), x0[s], N=N, profile=prof)
                xk[s] = r["xk"].T; uk[s] = r["uk"]; Uk[s] = r["Uk"].T; inner[s] = r["inner_iters"]
                cost[s] = r["cost"]; status[s] = r["status"]
            out.update({f"{name}_xk": xk, f"{name}_uk": uk, f"{name}_Uk": Uk, f"{name}_inner": inner,
                        f"{name}_cost": cost, f"{name}_status": status, f"{name}_flags": prof.flags()})
        np.savez_compressed(os.path.join(HERE, f"closed_loop_config{cfg}.npz"), **out)
        print("wrote config", cfg)


if __name__ == "__main__":
    main()
