#!/usr/bin/env python
"""What would a longest-first order of the work queue buy?  Keys = each scenario's ALONE time (measured one by one), the
batch permuted on the host, launch time in natural / longest-first / shortest-first order (config 3 shape, fixed(10))."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
Pf, x0f, N = physics.batch_params(3, S=S)
PT = np.ascontiguousarray(Pf.T)
dP = torch.from_numpy(PT).to(dev); dx = torch.from_numpy(x0f).to(dev)
xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
t = np.zeros(S)
for s in range(S):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    mpc.closed_loop_dev(1, N, 20, 10, 1e-14, 16, 0, dx[s:s + 1].data_ptr(), dP[s:s + 1].data_ptr(), 1, xk[s:s + 1].data_ptr(), uk[s:s + 1].data_ptr())
    e1.record(stream); torch.cuda.synchronize(); t[s] = e0.elapsed_time(e1)
print(f"S={S}: alone-time median {np.median(t):.3f} p90 {np.quantile(t, .9):.3f} p99 {np.quantile(t, .99):.3f} max {t.max():.3f} ms, sum {t.sum():.0f} ms")
def timed(order, name):
    dPo = torch.from_numpy(np.ascontiguousarray(PT[order])).to(dev); dxo = torch.from_numpy(np.ascontiguousarray(x0f[order])).to(dev)
    ts = []
    for rep in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc.closed_loop_dev(S, N, 20, 10, 1e-14, 16, 0, dxo.data_ptr(), dPo.data_ptr(), S, xk.data_ptr(), uk.data_ptr())
        e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"  {name:28s} min {min(ts):.3f} median {statistics.median(ts):.3f} ms")
timed(np.arange(S), "natural order")
timed(np.argsort(-t, kind="stable"), "longest first (alone time)")
timed(np.argsort(t, kind="stable"), "shortest first")
rng = np.random.default_rng(0)
for i in range(3): timed(rng.permutation(S), f"random order {i}")
