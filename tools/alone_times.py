#!/usr/bin/env python
"""Alone-time (min of 3 launches) of every scenario of a slice of config 3 -> npy.   python tools/alone_times.py lo hi out.npy"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
lo, hi, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
Pf, x0f, N = physics.batch_params(3, S=hi)
PT = np.ascontiguousarray(Pf.T)
dP = torch.from_numpy(PT[lo:hi].copy()).to(dev); dx = torch.from_numpy(x0f[lo:hi].copy()).to(dev)
S = hi - lo
xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
qp = torch.empty((S, 20), dtype=torch.int32, device=dev); inn = torch.empty((S, 20), dtype=torch.int32, device=dev)
t = np.full(S, 1e9)
for rep in range(3):
    for s in range(S):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc.closed_loop_dev(1, N, 20, 10, 1e-14, 16, 0, dx[s:s + 1].data_ptr(), dP[s:s + 1].data_ptr(), 1, xk[s:s + 1].data_ptr(), uk[s:s + 1].data_ptr(),
                            0, 0, inn[s:s + 1].data_ptr(), qp[s:s + 1].data_ptr(), 0)
        e1.record(stream); torch.cuda.synchronize(); t[s] = min(t[s], e0.elapsed_time(e1))
np.save(out, np.stack([t, qp.sum(dim=1).cpu().numpy().astype(float)]))
print(f"{lo}..{hi}: median {np.median(t):.3f} p99 {np.quantile(t, .99):.3f} max {t.max():.3f} ms at scenario {lo + int(t.argmax())}; sum {t.sum():.0f} ms")
