#!/usr/bin/env python
"""One resident launch (plus one warm-up) of the fused loop on a BASELINE config: the command ncu captures.
    python tools/run_loop.py <config 2..5> <S> [eps_break|fixed] [state rows 1|2 (binding box of tools/bench_ext.py)]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import ntm_mpc
from ntm_mpc import physics
cfg, S = int(sys.argv[1]), int(sys.argv[2])
flags = 0 if (len(sys.argv) > 3 and sys.argv[3] == 'eps_break') else 16
rows = int(sys.argv[4]) if len(sys.argv) > 4 else 0
BIND = (0.05, 0.16, 2000.0, 12000.0)
mpc = ntm_mpc.NtmMpc(0)
dev = torch.device("cuda:0")
P, x0, N = physics.batch_params(cfg, S=S)
dx = torch.from_numpy(x0).to(dev); dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev)
xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
inn = torch.empty((S, 20), dtype=torch.int32, device=dev); qp = torch.empty((S, 20), dtype=torch.int32, device=dev)
st = torch.empty((S,), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if rows:
        mpc.closed_loop_sc_dev(S, N, 20, 10, 1e-14, flags, 0, dx.data_ptr(), dP.data_ptr(), S, rows, BIND, xk.data_ptr(),
                               uk.data_ptr(), 0, 0, inn.data_ptr(), qp.data_ptr(), st.data_ptr())
    else:
        mpc.closed_loop_dev(S, N, 20, 10, 1e-14, flags, 0, dx.data_ptr(), dP.data_ptr(), S, xk.data_ptr(), uk.data_ptr(), 0, 0,
                            inn.data_ptr(), qp.data_ptr(), st.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(f"config{cfg} S={S} N={N}: {ms:.2f} ms = {S * 20 / ms * 1e3:.0f} scenario-steps/s, qp iters/inner {qp.sum().item() / inn.sum().item():.2f}, "
      f"inner/step {inn.sum().item() / (S * 20):.2f}, status max {st.max().item()}, uk checksum {float(torch.nan_to_num(uk).double().sum().item()):.17g}")
mpc.reset_stream()
