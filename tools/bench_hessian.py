"""Times ntm_hessian_grad_dev (dense G = 2 Gamma' Omega Gamma) on device-resident data."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np, torch
import ntm_mpc
from ntm_mpc import _lib, physics

mpc = ntm_mpc.NtmMpc(0); lib = _lib.load(); dev = torch.device("cuda:0")
mpc.set_stream(torch.cuda.current_stream().cuda_stream)
peak, _ = mpc.fp64_peak(1 << 14)
dpeak, _ = mpc.dmma_peak(1 << 12)
print(f"DFMA chains {peak:.2f} TFLOP/s, DMMA.8x8x4 chains {dpeak:.2f} TFLOP/s")
prm = torch.from_numpy(physics.params_from_physics(physics.nominal())).to(dev)
for N, S in ((20, 65536), (64, 8192), (100, 16384), (100, 2048)):
    Gam = torch.rand((S, N, 2 * N), dtype=torch.float64, device=dev)
    Phi = torch.rand((S, 2, 2 * N), dtype=torch.float64, device=dev); Lam = torch.rand((S, 2 * N), dtype=torch.float64, device=dev)
    x = torch.rand((S, 2), dtype=torch.float64, device=dev)
    G = torch.empty((S, N, N), dtype=torch.float64, device=dev); F = torch.empty((S, N), dtype=torch.float64, device=dev)
    def run():
        _lib.check(lib.ntm_hessian_grad_dev(mpc._h, 0, S, N, Phi.data_ptr(), Gam.data_ptr(), Lam.data_ptr(), x.data_ptr(), prm.data_ptr(), 1, G.data_ptr(), F.data_ptr()))
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[1]
    flops = S * (4.0 * N ** 3 + 6.0 * N ** 2)          # dense-no-Omega count of SURVEY 8(a9): 2*(2N)*N^2 + Omega + F
    gbytes = S * 8.0 * (2 * N * N + 4 * N + 2 * N + 2 + N * N + N)
    ref = 2 * torch.einsum("sck,sdk->scd", Gam[:4], Gam[:4])
    err = float((G[:4] - ref).abs().max() / ref.abs().max())
    print(f"N={N} S={S}: {ms:.3f} ms  {flops/ms/1e9:.2f} TFLOP/s dense-count ({flops/ms/1e9/peak*100:.1f}% of {peak:.1f}), "
          f"{gbytes/ms/1e6:.0f} GB/s algorithmic, max rel err vs torch {err:.1e}")
