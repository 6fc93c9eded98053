"""HBM roofline of the small materialising API kernels (device-resident data, MATLAB layout, per-scenario parameters)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np, torch
import ntm_mpc
from ntm_mpc import _lib, physics
mpc = ntm_mpc.NtmMpc(0); lib = _lib.load(); dev = torch.device("cuda:0")
mpc.set_stream(torch.cuda.current_stream().cuda_stream)
try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: peak = 6650.0
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
S = 1 << 22
P, x0, _ = physics.batch_params(4, S=min(S, 1 << 20))
reps = S // P.shape[1]
dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev).repeat(reps, 1).contiguous()
dx = torch.from_numpy(x0).to(dev).repeat(reps, 1).contiguous()
r1 = torch.empty(S, dtype=torch.float64, device=dev); r2 = torch.empty_like(r1); r3 = torch.empty_like(r1)
A = torch.empty((S, 4), dtype=torch.float64, device=dev); B = torch.empty((S, 2), dtype=torch.float64, device=dev)
u = torch.rand(S, dtype=torch.float64, device=dev) * 2e6; xn = torch.empty_like(dx)
for pc, pname in ((S, "per-scenario params"), (1, "shared params")):
    ms = timed(lambda: _lib.check(lib.ntm_rho_dev(mpc._h, 0, 0, S, dx.data_ptr(), dP.data_ptr(), pc, r1.data_ptr(), r2.data_ptr(), r3.data_ptr())))
    by = S * 8 * (2 + 3 + (16 if pc == S else 0)); print(f"ntm_rho      {pname}: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}% of {peak:.0f})")
    ms = timed(lambda: _lib.check(lib.ntm_lpv_AB_dev(mpc._h, 0, S, r1.data_ptr(), r2.data_ptr(), r3.data_ptr(), dP.data_ptr(), pc, A.data_ptr(), B.data_ptr())))
    by = S * 8 * (3 + 6 + (16 if pc == S else 0)); print(f"ntm_lpv_AB   {pname}: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
    ms = timed(lambda: _lib.check(lib.ntm_plant_step_dev(mpc._h, 0, 0, S, dx.data_ptr(), u.data_ptr(), dP.data_ptr(), pc, xn.data_ptr())))
    by = S * 8 * (2 + 1 + 2 + (16 if pc == S else 0)); print(f"ntm_plant    {pname}: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
del dP, dx, r1, r2, r3, A, B, u, xn
S2, N = 65536, 20
Gam = torch.rand((S2, N, 2 * N), dtype=torch.float64, device=dev); Phi = torch.rand((S2, 2, 2 * N), dtype=torch.float64, device=dev); Lam = torch.rand((S2, 2 * N), dtype=torch.float64, device=dev)
R = 6 * N + 4
W = torch.empty((S2, 2, R), dtype=torch.float64, device=dev); L = torch.empty((S2, N, R), dtype=torch.float64, device=dev); c = torch.empty((S2, R), dtype=torch.float64, device=dev)
b = np.array([0.15, 31415.9, 0.06, 628.3, 2e6, 0.0])
ms = timed(lambda: _lib.check(lib.ntm_getWLc_dev(mpc._h, 0, S2, N, b.ctypes.data, Gam.data_ptr(), Phi.data_ptr(), Lam.data_ptr(), W.data_ptr(), L.data_ptr(), c.data_ptr())))
by = S2 * 8 * (2 * N * N + 4 * N + 2 * N + R * (N + 3)); print(f"ntm_getWLc N=20: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
x = torch.rand((S2, 2), dtype=torch.float64, device=dev); prm = torch.from_numpy(physics.params_from_physics(physics.nominal())).to(dev)
G = torch.empty((S2, N, N), dtype=torch.float64, device=dev); F = torch.empty((S2, N), dtype=torch.float64, device=dev)
ms = timed(lambda: _lib.check(lib.ntm_hessian_grad_dev(mpc._h, 0, S2, N, Phi.data_ptr(), Gam.data_ptr(), Lam.data_ptr(), x.data_ptr(), prm.data_ptr(), 1, G.data_ptr(), F.data_ptr())))
by = S2 * 8 * (2 * N * N + 4 * N + 2 * N + 2 + N * N + N); print(f"ntm_hessian_grad N=20: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
del Gam, Phi, Lam, W, L, c, G, F
for N, S3 in ((20, 262144), (64, 32768), (100, 16384)):
    rho = torch.rand((3, S3, N), dtype=torch.float64, device=dev) * 1e-3 + 1e-3
    phi = torch.empty(S3 * 4 * N, dtype=torch.float64, device=dev); gam = torch.empty(S3 * 2 * N * N, dtype=torch.float64, device=dev); lam = torch.empty(S3 * 2 * N, dtype=torch.float64, device=dev)
    for prof, nm in ((0, "literal"), (2, "index i")):
        ms = timed(lambda: mpc.condense_dev(S3, N, prof, 0, rho[0].data_ptr(), rho[1].data_ptr(), rho[2].data_ptr(), prm.data_ptr(), 1, phi.data_ptr(), gam.data_ptr(), lam.data_ptr()))
        by = S3 * 8 * (3 * N + 4 * N + 2 * N * N + 2 * N); print(f"ntm_condense N={N} {nm}: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
    del rho, phi, gam, lam
# Monte-Carlo reduction: reads xk (2(K+1)), uk (K), cost, status (4 B) and umin/umax per scenario
for lay, lname in ((0, "MATLAB layout"), (1, "SoA layout")):
    S3, K = 1 << 20, 20
    xk = torch.rand(S3 * 2 * (K + 1), dtype=torch.float64, device=dev) * 0.2; uk = torch.rand(S3 * K, dtype=torch.float64, device=dev) * 2e6
    cost = torch.rand(S3, dtype=torch.float64, device=dev); st = torch.zeros(S3, dtype=torch.int32, device=dev)
    prm3 = torch.zeros(S3 * 16, dtype=torch.float64, device=dev); out = torch.empty(64, dtype=torch.float64, device=dev)
    bb = np.array([0.06, 0.15, 628.3, 31415.9])
    ms = timed(lambda: _lib.check(lib.ntm_mc_stats_dev(mpc._h, lay, S3, K, xk.data_ptr(), uk.data_ptr(), cost.data_ptr(), st.data_ptr(), prm3.data_ptr(), S3, bb.ctypes.data, 0.06, 0.2, out.data_ptr())))
    by = S3 * (8 * (2 * (K + 1) + K + 1 + 2) + 4); print(f"ntm_mc_stats {lname} S=2^20 k_sim=20: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
    um = torch.zeros(S3, dtype=torch.float64, device=dev); ux = torch.full((S3,), 2e6, dtype=torch.float64, device=dev)
    ms = timed(lambda: _lib.check(lib.ntm_mc_stats_ub_dev(mpc._h, lay, S3, K, xk.data_ptr(), uk.data_ptr(), cost.data_ptr(), st.data_ptr(), um.data_ptr(), ux.data_ptr(), S3, bb.ctypes.data, 0.06, 0.2, out.data_ptr())))
    print(f"ntm_mc_stats_ub {lname} (compact umin/umax arrays): {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
    del xk, uk, cost, st, prm3
# SoA layout (element index slowest, scenario index fastest) of the materialising kernels
for N, S3 in ((20, 262144),):
    rho = torch.rand((3, N, S3), dtype=torch.float64, device=dev) * 1e-3 + 1e-3
    phi = torch.empty(S3 * 4 * N, dtype=torch.float64, device=dev); gam = torch.empty(S3 * 2 * N * N, dtype=torch.float64, device=dev); lam = torch.empty(S3 * 2 * N, dtype=torch.float64, device=dev)
    for prof, nm in ((0, "literal"), (2, "index i")):
        ms = timed(lambda: mpc.condense_dev(S3, N, prof, 1, rho[0].data_ptr(), rho[1].data_ptr(), rho[2].data_ptr(), prm.data_ptr(), 1, phi.data_ptr(), gam.data_ptr(), lam.data_ptr()))
        by = S3 * 8 * (3 * N + 4 * N + 2 * N * N + 2 * N); print(f"ntm_condense SoA N={N} {nm}: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
    del rho, phi, gam, lam
S2, N = 65536, 20
R = 6 * N + 4
Gam = torch.rand((2 * N * N, S2), dtype=torch.float64, device=dev); Phi = torch.rand((4 * N, S2), dtype=torch.float64, device=dev); Lam = torch.rand((2 * N, S2), dtype=torch.float64, device=dev)
W = torch.empty((2 * R, S2), dtype=torch.float64, device=dev); L = torch.empty((N * R, S2), dtype=torch.float64, device=dev); c = torch.empty((R, S2), dtype=torch.float64, device=dev)
b = np.array([0.15, 31415.9, 0.06, 628.3, 2e6, 0.0])
ms = timed(lambda: _lib.check(lib.ntm_getWLc_dev(mpc._h, 1, S2, N, b.ctypes.data, Gam.data_ptr(), Phi.data_ptr(), Lam.data_ptr(), W.data_ptr(), L.data_ptr(), c.data_ptr())))
by = S2 * 8 * (2 * N * N + 4 * N + 2 * N + R * (N + 3)); print(f"ntm_getWLc SoA N=20: {ms:.3f} ms {by/ms/1e6:.0f} GB/s ({by/ms/1e6/peak*100:.0f}%)")
