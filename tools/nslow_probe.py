#!/usr/bin/env python
"""Debug build (-DNTM_DEBUG_NSLOW: inner_iters carries the cumulative number of QPs that took the active-set path):
how well do the first time steps predict a scenario's total?   NTM_MPC_LIB=.../libntm_mpc_dbg.so NTM_LPT=0 python tools/nslow_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np
import ntm_mpc
from ntm_mpc import physics
mpc = ntm_mpc.NtmMpc(0)
for cfg, S in ((3, 8192), (4, 8192)):
    P, x0, N = physics.batch_params(cfg, S=S)
    g = mpc.closed_loop(x0, P.T, N=N, profile=16)
    cum = g["inner_iters"].astype(float)                     # [S, 20] cumulative slow-QP count after each step
    tot = cum[:, -1]
    per = np.diff(np.concatenate([np.zeros((S, 1)), cum], axis=1), axis=1)
    print(f"config {cfg}: slow QPs per scenario: median {np.median(tot):.0f} p90 {np.quantile(tot, .9):.0f} p99 {np.quantile(tot, .99):.0f} max {tot.max():.0f} of 200; mean per step {per.mean(axis=0).round(2).tolist()}")
    for name, key in (("step 0", per[:, 0]), ("step 1", per[:, 1]), ("steps 0-1", cum[:, 1]), ("steps 1-2", cum[:, 2] - cum[:, 0]), ("steps 0-3", cum[:, 3])):
        c = np.corrcoef(key, tot)[0, 1]
        rest = tot - key
        c2 = np.corrcoef(key, rest)[0, 1]
        top = np.argsort(-key, kind="stable")[: S // 20]
        heavy = np.argsort(-tot, kind="stable")[: S // 20]
        print(f"   key = slow QPs in {name:10s}: corr with total {c:.3f}, with the rest {c2:.3f}; top-5% overlap {len(set(top) & set(heavy)) / len(heavy):.2f}")
