#!/usr/bin/env python
"""Dense-Gamma instantiation of the fused loop on long horizons (tensor-core build of G, F for 2 / 4 warps per scenario):
the literal reading through the dense path (NTM_PROFILE_DENSE_G) must equal the Toeplitz path, the consistent reading is
compared with the C oracle; timings at N = 100.  GPU box only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import ntm_mpc
from ntm_mpc import physics
from oracle import c_oracle, ntm_oracle as o
mpc = ntm_mpc.NtmMpc(0)
threads = len(os.sched_getaffinity(0))
for N in (33, 40, 63, 64, 72, 100, 103, 104):
    phys, x0, _ = o.make_batch(5 if N >= 100 else 3, S=64)
    P = physics.params_from_physics(phys).reshape(16, -1); prm = np.ascontiguousarray(P.T)
    a = mpc.closed_loop(x0, prm, N, 6, 10, 1e-14, 16)
    b = mpc.closed_loop(x0, prm, N, 6, 10, 1e-14, 16 | 32)
    umax = np.broadcast_to(np.asarray(phys["umax"], dtype=float), (64,))
    du = np.max(np.abs(a["uk"] - b["uk"]), axis=1) / umax
    c = c_oracle.closed_loop_batch(phys, x0, N, 6, 10, 1e-14, 30, threads)
    g = mpc.closed_loop(x0, prm, N, 6, 10, 1e-14, 30)
    dc = np.max(np.abs(g["uk"] - c["uk"]), axis=1) / umax
    print(f"N={N:3d}: dense-literal vs Toeplitz max du {du.max():.2e} (bad {(du > 1e-6).sum()}/64), status {b['status'].max()}; "
          f"consistent vs C oracle: median du {np.median(dc):.1e}, within 1e-6: {(dc <= 1e-6).mean():.2f}, status {g['status'].max()}", flush=True)
dev = torch.device("cuda:0")
S = 2048
Pp, x0, N = physics.batch_params(5, S=S)
dx = torch.from_numpy(x0).to(dev); dP = torch.from_numpy(np.ascontiguousarray(Pp.T)).to(dev)
xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
for flags, name in ((16, "literal Toeplitz"), (16 | 32, "literal, dense G on tensor cores"), (30, "consistent (dense)")):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc.closed_loop_dev(S, N, 20, 10, 1e-14, flags, 0, dx.data_ptr(), dP.data_ptr(), S, xk.data_ptr(), uk.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
    print(f"config5 S={S} {name}: {e0.elapsed_time(e1):.1f} ms = {S * 20 / e0.elapsed_time(e1) * 1e3:.0f} scenario-steps/s", flush=True)
mpc.reset_stream()
