import sys, time
sys.path.insert(0, 'mpc-ntm-control_b200'); sys.path.insert(0, '.')
import numpy as np, ntm_mpc
from ntm_mpc import physics
mpc = ntm_mpc.NtmMpc(0)
P, x0, N = physics.batch_params(5, S=8)
t=time.time(); r = mpc.closed_loop(x0, P.T, N=N, profile=16, want_Uk=True); print('time', time.time()-t)
np.set_printoptions(linewidth=250)
for s in range(4):
    print('qp', r['qp_iters'][s])
    print('uk', r['uk'][s][:8])
np.savez('gpurun_out/diag5.npz', **{k:v for k,v in r.items() if v is not None})
for S in (148, 296, 592):
    P, x0, N = physics.batch_params(5, S=S)
    t=time.time(); r = mpc.closed_loop(x0, P.T, N=N, k_sim=2, profile=16); print(S, 'k_sim=2 time', time.time()-t, r['qp_iters'].mean())
