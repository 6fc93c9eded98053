#!/usr/bin/env python
"""Kernel time of the fused loop against the batch size (config 3 shape, fixed(10)): where does a short launch spend its
time?  20 launches per size after a 200 ms warm-up of the clocks, minimum and median reported."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
Pf, x0f, N = physics.batch_params(3, S=65536)
def run(S, reps):
    dx = torch.from_numpy(x0f[:S]).to(dev); dP = torch.from_numpy(np.ascontiguousarray(Pf.T[:S])).to(dev)
    xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc.closed_loop_dev(S, N, 20, 10, 1e-14, 16, 0, dx.data_ptr(), dP.data_ptr(), S, xk.data_ptr(), uk.data_ptr())
        e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return ts
run(65536, 8)                                   # clocks up
for S in (1, 4, 32, 148, 592, 740, 1480, 2960, 3700, 5920, 8192, 8880, 11840, 16384):
    ts = run(S, 20)
    print(f"S={S:6d}: min {min(ts):7.3f} ms  median {statistics.median(ts):7.3f} ms  -> {S * 20 / min(ts) / 1e3:8.2f} M steps/s", flush=True)
