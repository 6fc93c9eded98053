#!/usr/bin/env python
"""GPU box: scenarios on which the CUDA kernels and the C oracle disagree by more than 1e-6 (literal reading), dumped for
adjudication in 50-digit arithmetic (tools/adjudicate.py reads the .npz).  Horizons 20..48, configs 3 / 4, both policies.
    python tools/find_chaotic.py [S] [out.npz]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np
import ntm_mpc
from ntm_mpc import physics
from oracle import c_oracle, ntm_oracle as o

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r2_chaotic.npz")
mpc = ntm_mpc.NtmMpc(0)
threads = len(os.sched_getaffinity(0))
K_SIM = 8
cases = []
tot = {}
for N in (20, 24, 32, 33, 40, 48):
    for cfg in (3, 4):
        for flags in (16, 0):
            phys, x0, _ = o.make_batch(cfg, S=S)
            P = physics.params_from_physics(phys).reshape(16, -1)
            g = mpc.closed_loop(x0, np.ascontiguousarray(P.T), N, K_SIM, 10, 1e-14, flags)
            c = c_oracle.closed_loop_batch(phys, x0, N, K_SIM, 10, 1e-14, flags, threads)
            umax = np.asarray(phys["umax"], dtype=float)
            du = np.max(np.abs(g["uk"] - c["uk"]), axis=1) / umax
            w = c["xk"][:, :, 0]
            dw = np.max(np.abs(g["xk"][:, :, 0] - w), axis=1) / np.maximum(np.max(np.abs(w), axis=1), 1e-3)
            bad = np.flatnonzero((du > 1e-6) | (dw > 1e-6))
            tot[(N, cfg, flags)] = (len(bad), S)
            print(f"N={N:2d} cfg{cfg} flags={flags:2d}: {len(bad)} of {S} disagree  {bad[:8].tolist()}", flush=True)
            for s in bad[:3]:
                cases.append(dict(cfg=cfg, N=N, S=S, s=int(s), flags=flags, k_sim=K_SIM, gpu_uk=g["uk"][s], gpu_xk=g["xk"][s], c_uk=c["uk"][s]))
np.savez(out, cfg=[c["cfg"] for c in cases], N=[c["N"] for c in cases], S=[c["S"] for c in cases], s=[c["s"] for c in cases],
         flags=[c["flags"] for c in cases], k_sim=[c["k_sim"] for c in cases], gpu_uk=np.array([c["gpu_uk"] for c in cases]),
         gpu_xk=np.array([c["gpu_xk"] for c in cases]), c_uk=np.array([c["c_uk"] for c in cases]),
         counts=np.array([[k[0], k[1], k[2], v[0], v[1]] for k, v in tot.items()]))
print("cases", len(cases), "->", out)
