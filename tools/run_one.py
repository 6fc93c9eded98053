#!/usr/bin/env python
"""ONE scenario of a BASELINE config alone on the GPU (the command ncu captures for the latency of a slow scenario):
    python tools/run_one.py <config> <scenario index>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
cfg, s = int(sys.argv[1]), int(sys.argv[2])
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
P, x0, N = physics.batch_params(cfg, S=s + 1)
dx = torch.from_numpy(x0[s:s + 1].copy()).to(dev); dP = torch.from_numpy(np.ascontiguousarray(P.T[s:s + 1])).to(dev)
xk = torch.empty((1, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((1, 20), dtype=torch.float64, device=dev)
inn = torch.empty((1, 20), dtype=torch.int32, device=dev); qp = torch.empty((1, 20), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    mpc.closed_loop_dev(1, N, 20, 10, 1e-14, 16, 0, dx.data_ptr(), dP.data_ptr(), 1, xk.data_ptr(), uk.data_ptr(), 0, 0, inn.data_ptr(), qp.data_ptr(), 0)
    e1.record(stream); torch.cuda.synchronize()
print(f"config{cfg} scenario {s} N={N}: {e0.elapsed_time(e1):.3f} ms alone, qp iterations {int(qp.sum().item())} over {int(inn.sum().item())} QPs, "
      f"interior inputs {int(((uk > P[8, s]) & (uk < P[9, s])).sum().item())}/20")
mpc.reset_stream()
