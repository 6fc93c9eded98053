#!/usr/bin/env python
"""Kernel time of every 8-GPU shard of config 3 (8,192 scenarios each) on ONE GPU: the strong-scaling step is the max."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
Pf, x0f, N = physics.batch_params(3, S=65536)
PT = np.ascontiguousarray(Pf.T); Ss = 65536 // G
out = []
for g in range(G):
    dP = torch.from_numpy(PT[g * Ss:(g + 1) * Ss].copy()).to(dev); dx = torch.from_numpy(x0f[g * Ss:(g + 1) * Ss].copy()).to(dev)
    xk = torch.empty((Ss, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((Ss, 20), dtype=torch.float64, device=dev)
    ts = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); mpc.closed_loop_dev(Ss, N, 20, 10, 1e-14, 16, 0, dx.data_ptr(), dP.data_ptr(), Ss, xk.data_ptr(), uk.data_ptr())
        e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    out.append(statistics.median(ts[1:]))
print(f"{G} shards of {Ss}: " + " ".join(f"{t:.2f}" for t in out) + f"  max {max(out):.2f} mean {sum(out)/len(out):.2f} ms")
