#!/usr/bin/env python
"""Adjudicates every case of gpurun_out/r2_chaotic.npz (tools/find_chaotic.py) in 50-digit arithmetic and prints one row per
case: deviation of the NumPy oracle, the C oracle and the CUDA kernels from the exact closed loop (max |du| / umax)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import adjudicate as A
from oracle import c_oracle, ntm_oracle as o

z = np.load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_chaotic.npz"))
seen = set()
print("cfg  N     s flags | exact-vs: numpy      C-oracle   CUDA      | CUDA-vs-C  | first step off (numpy/C/CUDA) | free u in exact run")
for i in range(len(z["cfg"])):
    cfg, N, S, s, flags, k_sim = (int(z[k][i]) for k in ("cfg", "N", "S", "s", "flags", "k_sim"))
    phys, x0, _ = o.make_batch(cfg, S=S)
    p = o.scenario(phys, s)
    umax = float(p["umax"])
    t0 = time.time()
    uk_mp, xk_mp, inner = A.mp_closed_loop(p, x0[s], N, k_sim, 10, flags)
    prof = o.Profile(rho1_variant=flags & 1, gamma_index=(flags >> 1) & 1, f_state=(flags >> 2) & 1, plant_affine=(flags >> 3) & 1,
                     inner_policy=(flags >> 4) & 1)
    r = o.closed_loop(p, x0[s], N=N, k_sim=k_sim, i_sim=10, profile=prof)
    sub = {k: np.asarray(v)[s:s + 1] for k, v in phys.items()}
    c = c_oracle.closed_loop_batch(sub, x0[s:s + 1], N, k_sim, 10, 1e-14, flags & 31, 1)
    dev = lambda u: np.abs(np.asarray(u) - uk_mp) / umax
    first = lambda u: int(np.argmax(dev(u) > 1e-6)) if (dev(u) > 1e-6).any() else -1
    dn, dc, dg = dev(r["uk"]), dev(c["uk"][0]), dev(z["gpu_uk"][i])
    interior = int(np.sum((uk_mp > 0) & (uk_mp < umax)))
    print(f"{cfg:3d} {N:2d} {s:5d} {flags:4d}  | {dn.max():10.2e} {dc.max():10.2e} {dg.max():10.2e} | {np.max(np.abs(z['gpu_uk'][i] - c['uk'][0])) / umax:10.2e} | "
          f"{first(r['uk']):3d} {first(c['uk'][0]):3d} {first(z['gpu_uk'][i]):3d}   | {interior} of {k_sim}   ({time.time() - t0:.0f} s)", flush=True)
