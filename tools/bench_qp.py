"""Micro-benchmark of ntm_qp_box on device-resident data: interior solutions (all free) so every solve pays a full LDL'."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np, torch, ctypes
import ntm_mpc
from ntm_mpc import _lib

mpc = ntm_mpc.NtmMpc(0)
lib = _lib.load()
dev = torch.device("cuda:0")
mpc.set_stream(torch.cuda.current_stream().cuda_stream)
for N, S in ((20, 4736), (64, 1184), (100, 592), (100, 148)):
    rng = np.random.default_rng(N)
    M = rng.standard_normal((S, 2 * N, N))
    G = 2 * np.einsum("ski,skj->sij", M, M)
    F = rng.standard_normal((S, N))
    for box, name in ((1e6, "interior"), (0.05, "mixed")):
        dG = torch.from_numpy(G).to(dev); dF = torch.from_numpy(F).to(dev)
        lb = torch.full((N,), -box, dtype=torch.float64, device=dev); ub = torch.full((N,), box, dtype=torch.float64, device=dev)
        U = torch.empty((S, N), dtype=torch.float64, device=dev); it = torch.empty(S, dtype=torch.int32, device=dev); st = torch.empty(S, dtype=torch.int32, device=dev)
        def run():
            _lib.check(lib.ntm_qp_box_dev(mpc._h, 0, S, N, dG.data_ptr(), dF.data_ptr(), lb.data_ptr(), ub.data_ptr(), 1, U.data_ptr(), it.data_ptr(), st.data_ptr()))
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        iters = it.double().mean().item()
        free = ((U > -box) & (U < box)).double().mean().item()
        print(f"N={N} S={S} {name}: {ms:.3f} ms, mean iters {iters:.1f}, free frac {free:.2f}, status max {int(st.max())}, "
              f"us per QP-iteration per CTA-slot ~ {ms*1e3/ (iters * max(S/ (148*(6 if N<=32 else 1)),1)):.1f}")
