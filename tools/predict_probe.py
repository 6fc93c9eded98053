#!/usr/bin/env python
"""Is a scenario's cost predictable from its FIRST QP?  Alone-time of each scenario (one warp on an idle GPU) against the
number of interior components / active-set iterations of the first QP (k_sim = 1, i_sim = 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Pf, x0f, N = physics.batch_params(3, S=S)
dP = torch.from_numpy(np.ascontiguousarray(Pf.T)).to(dev); dx = torch.from_numpy(x0f).to(dev)
# first QP of every scenario
xk1 = torch.empty((S, 2, 2), dtype=torch.float64, device=dev); uk1 = torch.empty((S, 1), dtype=torch.float64, device=dev)
Uk1 = torch.empty((S, 1, N), dtype=torch.float64, device=dev); qp1 = torch.empty((S, 1), dtype=torch.int32, device=dev); in1 = torch.empty((S, 1), dtype=torch.int32, device=dev)
mpc.closed_loop_dev(S, N, 1, 1, 1e-14, 16, 0, dx.data_ptr(), dP.data_ptr(), S, xk1.data_ptr(), uk1.data_ptr(), Uk1.data_ptr(), 0, in1.data_ptr(), qp1.data_ptr(), 0)
torch.cuda.synchronize()
U = Uk1[:, 0, :].cpu().numpy(); umax = Pf[9][:, None]; umin = Pf[8][:, None]
nfree = ((U > umin) & (U < umax)).sum(axis=1); it1 = qp1[:, 0].cpu().numpy()
# full run for the iteration counts, alone-times
xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
inn = torch.empty((S, 20), dtype=torch.int32, device=dev); qp = torch.empty((S, 20), dtype=torch.int32, device=dev)
t = np.zeros(S)
for s in range(S):
    best = 1e9
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc.closed_loop_dev(1, N, 20, 10, 1e-14, 16, 0, dx[s:s + 1].data_ptr(), dP[s:s + 1].data_ptr(), 1, xk[s:s + 1].data_ptr(),
                            uk[s:s + 1].data_ptr(), 0, 0, inn[s:s + 1].data_ptr(), qp[s:s + 1].data_ptr(), 0)
        e1.record(stream); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    t[s] = best
ukh = uk.cpu().numpy()
interior_u0 = ((ukh > umin) & (ukh < umax)).mean(axis=1)
print(f"S={S}: alone-time median {np.median(t):.3f} p90 {np.quantile(t, 0.9):.3f} p99 {np.quantile(t, 0.99):.3f} max {t.max():.3f} ms")
for name, key in (("free components of the first QP", nfree), ("iterations of the first QP", it1), ("fraction of interior uk over the run (hindsight)", interior_u0),
                  ("w0", x0f[:, 0]), ("omega0", x0f[:, 1]), ("umax*c_b", Pf[9] * Pf[3])):
    c = np.corrcoef(key.astype(float), t)[0, 1]
    order = np.argsort(-key.astype(float), kind="stable")
    top = order[: max(S // 10, 1)]
    heavy = t > 1.5 * np.median(t)
    print(f"  {name:50s} corr {c:+.2f}; of the {heavy.sum()} heavy scenarios (> 1.5 x median) {heavy[top].sum()} are in the top decile of this key")
print("nfree histogram", np.bincount(nfree, minlength=N + 1).tolist())
print("mean alone-time by nfree:", [(int(k), round(float(t[nfree == k].mean()), 3), int((nfree == k).sum())) for k in np.unique(nfree)])
