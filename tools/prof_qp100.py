import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np, torch
import ntm_mpc
from ntm_mpc import _lib
mpc = ntm_mpc.NtmMpc(0); lib = _lib.load(); dev = torch.device("cuda:0")
mpc.set_stream(torch.cuda.current_stream().cuda_stream)
N, S = 100, 296
rng = np.random.default_rng(N)
M = rng.standard_normal((S, 2 * N, N)); G = 2 * np.einsum("ski,skj->sij", M, M); F = rng.standard_normal((S, N))
dG = torch.from_numpy(G).to(dev); dF = torch.from_numpy(F).to(dev)
lb = torch.full((N,), -1e6, dtype=torch.float64, device=dev); ub = torch.full((N,), 1e6, dtype=torch.float64, device=dev)
U = torch.empty((S, N), dtype=torch.float64, device=dev); it = torch.empty(S, dtype=torch.int32, device=dev); st = torch.empty(S, dtype=torch.int32, device=dev)
for _ in range(3):
    _lib.check(lib.ntm_qp_box_dev(mpc._h, 0, S, N, dG.data_ptr(), dF.data_ptr(), lb.data_ptr(), ub.data_ptr(), 1, U.data_ptr(), it.data_ptr(), st.data_ptr()))
torch.cuda.synchronize()
print("ok", int(st.max()), float(it.double().mean()))
