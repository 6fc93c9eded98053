"""Times the fused closed loop under the other profile switches (consistent reading, dense-G cross-check path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
mpc.set_stream(torch.cuda.current_stream().cuda_stream)
for cfg, S in ((3, 65536), (2, 1024), (5, 1024)):
    P, x0, N = physics.batch_params(cfg, S=S)
    dx = torch.from_numpy(x0).to(dev); dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev)
    xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
    inn = torch.empty((S, 20), dtype=torch.int32, device=dev); qp = torch.empty((S, 20), dtype=torch.int32, device=dev); st = torch.empty((S,), dtype=torch.int32, device=dev)
    for name, prof in (("literal_fixed", 16), ("literal_fixed_denseG", 16 | 32), ("consistent_fixed", ntm_mpc.PROFILE_CONSISTENT | 16), ("consistent_eps", ntm_mpc.PROFILE_CONSISTENT)):
        if cfg == 5 and "dense" in name: continue
        def run():
            mpc.closed_loop_dev(S, N, 20, 10, 1e-14, prof, 0, dx.data_ptr(), dP.data_ptr(), S, xk.data_ptr(), uk.data_ptr(), 0, 0, inn.data_ptr(), qp.data_ptr(), st.data_ptr())
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"config{cfg} N={N} S={S} {name}: {ms:.2f} ms  {S*20/ms/1e3:.3f} M scenario-steps/s  inner {inn.double().mean().item():.2f} qp/inner {qp.double().sum().item()/inn.double().sum().item():.2f} status {int(st.max())}")
