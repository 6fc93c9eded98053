#!/usr/bin/env python
"""Which scenarios make a short launch long?  Every scenario of a small sample runs ALONE (one warp on an idle GPU):
kernel time against its number of active-set iterations (config 3 and config 2 shapes, fixed(10))."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
for cfg, S in ((3, 48), (2, 48)):
    Pf, x0f, N = physics.batch_params(cfg, S=S)
    dP = torch.from_numpy(np.ascontiguousarray(Pf.T)).to(dev); dx = torch.from_numpy(x0f).to(dev)
    xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
    inn = torch.empty((S, 20), dtype=torch.int32, device=dev); qp = torch.empty((S, 20), dtype=torch.int32, device=dev)
    rows = []
    for s in range(S):
        ts = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mpc.closed_loop_dev(1, N, 20, 10, 1e-14, 16, 0, dx[s:s + 1].data_ptr(), dP[s:s + 1].data_ptr(), 1, xk[s:s + 1].data_ptr(),
                                uk[s:s + 1].data_ptr(), 0, 0, inn[s:s + 1].data_ptr(), qp[s:s + 1].data_ptr(), 0)
            e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        rows.append((s, min(ts), int(qp[s].sum().item()), int((qp[s] > 10).sum().item())))
    t = np.array([r[1] for r in rows]); q = np.array([r[2] for r in rows], dtype=float)
    A = np.stack([np.ones_like(q), q - 200.0], axis=1)
    coef = np.linalg.lstsq(A, t, rcond=None)[0]
    print(f"config {cfg} N={N}: alone-time min {t.min():.3f} median {np.median(t):.3f} max {t.max():.3f} ms; qp iterations per scenario min {q.min():.0f} "
          f"median {np.median(q):.0f} max {q.max():.0f} (200 = one per QP); fit t = {coef[0]:.3f} ms + {coef[1] * 1e3:.2f} us per extra iteration")
    for r in sorted(rows, key=lambda r: -r[1])[:6]:
        print(f"   scenario {r[0]:3d}: {r[1]:.3f} ms, {r[2]} active-set iterations over 200 QPs")
