"""Times the EXT instantiation of the fused loop: RK4 plant and getWLc's state rows kept in every QP (config 3 shape,
CUDA events, resident inputs), next to the headline kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np, torch
import ntm_mpc
from ntm_mpc import physics
mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
mpc.set_stream(torch.cuda.current_stream().cuda_stream)
WIDE = (-1e3, 1e3, -1e9, 1e9)
SCRIPT = (0.06, 0.15, 200 * np.pi, 10000 * np.pi)           # NTM_MPC_Sim.m:39-45
BIND = (0.05, 0.16, 2000.0, 12000.0)
for cfg, S in ((3, 65536), (2, 1024), (5, 512)):
    P, x0, N = physics.batch_params(cfg, S=S)
    dx = torch.from_numpy(x0).to(dev); dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev)
    xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
    inn = torch.empty((S, 20), dtype=torch.int32, device=dev); qp = torch.empty((S, 20), dtype=torch.int32, device=dev); st = torch.empty((S,), dtype=torch.int32, device=dev)
    cases = [("literal_fixed (headline kernel)", 16, 0, None), ("literal_fixed + RK4 plant", 16 | 64, 0, None),
             ("state rows REFRESH, box never binds", 16, 1, WIDE), ("state rows REFRESH, script box", 16, 1, SCRIPT),
             ("state rows REFRESH, binding box", 16, 1, BIND), ("state rows FROZEN, binding box", 16, 2, BIND)]
    for name, prof, rows, xb in cases:
        if cfg == 5 and rows and N > 100: continue
        def run():
            if rows:
                mpc.closed_loop_sc_dev(S, N, 20, 10, 1e-14, prof, 0, dx.data_ptr(), dP.data_ptr(), S, rows, xb, xk.data_ptr(), uk.data_ptr(), 0, 0, inn.data_ptr(), qp.data_ptr(), st.data_ptr())
            else:
                mpc.closed_loop_dev(S, N, 20, 10, 1e-14, prof, 0, dx.data_ptr(), dP.data_ptr(), S, xk.data_ptr(), uk.data_ptr(), 0, 0, inn.data_ptr(), qp.data_ptr(), st.data_ptr())
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        stv = st.cpu().numpy()
        steps = float((inn > 0).sum().item())
        print(f"config{cfg} N={N} S={S} {name}: {ms:.2f} ms  {S*20/ms/1e3:.3f} M scenario-steps/s nominal, {steps/ms/1e3:.3f} M executed  "
              f"qp/inner {qp.double().sum().item()/max(inn.double().sum().item(),1):.2f}  ok {int((stv==0).sum())} cap {int((stv==1).sum())} nonfinite {int((stv==2).sum())} infeasible {int((stv==3).sum())}", flush=True)
