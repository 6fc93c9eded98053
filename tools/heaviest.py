#!/usr/bin/env python
"""The slowest scenarios of a slice of config 3, each run ALONE: time, active-set iterations, interior inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
lo, hi = int(sys.argv[1]), int(sys.argv[2])
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
Pf, x0f, N = physics.batch_params(3, S=hi)
PT = np.ascontiguousarray(Pf.T)
dP = torch.from_numpy(PT[lo:hi].copy()).to(dev); dx = torch.from_numpy(x0f[lo:hi].copy()).to(dev)
S = hi - lo
xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
Uk = torch.empty((S, 20, N), dtype=torch.float64, device=dev)
inn = torch.empty((S, 20), dtype=torch.int32, device=dev); qp = torch.empty((S, 20), dtype=torch.int32, device=dev)
t = np.zeros(S)
for s in range(S):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    mpc.closed_loop_dev(1, N, 20, 10, 1e-14, 16, 0, dx[s:s + 1].data_ptr(), dP[s:s + 1].data_ptr(), 1, xk[s:s + 1].data_ptr(), uk[s:s + 1].data_ptr(),
                        Uk[s:s + 1].data_ptr(), 0, inn[s:s + 1].data_ptr(), qp[s:s + 1].data_ptr(), 0)
    e1.record(stream); torch.cuda.synchronize(); t[s] = e0.elapsed_time(e1)
t[0] = np.median(t)
U = Uk.cpu().numpy(); umax = PT[lo:hi, 9][:, None, None]; umin = PT[lo:hi, 8][:, None, None]
nfree = ((U > umin) & (U < umax)).sum(axis=2)          # [S, 20] free components of the last QP of each step
q = qp.cpu().numpy()
print(f"scenarios {lo}..{hi}: alone-time median {np.median(t):.3f} p99 {np.quantile(t, .99):.3f} max {t.max():.3f} ms")
for s in np.argsort(-t)[:8]:
    print(f"  scenario {lo + s}: {t[s]:.3f} ms, active-set iterations {int(q[s].sum())} over 200 QPs (per step {q[s].tolist()}), free components of each step's last QP {nfree[s].tolist()}")
