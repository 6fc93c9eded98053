#!/usr/bin/env python
"""Long horizons (N > 32) of the fused loop against the C oracle: config 5 (N = 100) on S scenarios and a few other
horizons, Uk included; then the timing of config 5 at its stated size.  GPU box only.
    python tools/check_long.py [S_parity] [S_timing]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np
import ntm_mpc
from ntm_mpc import physics
from oracle import c_oracle, ntm_oracle as o

S_par = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S_tim = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
mpc = ntm_mpc.NtmMpc(0)
threads = len(os.sched_getaffinity(0))


def compare(cfg, S, N=None, flags=16, k_sim=20):
    phys, x0, Nc = o.make_batch(cfg, S=S)
    N = N or Nc
    P = physics.params_from_physics(phys).reshape(16, -1)
    t0 = time.time()
    g = mpc.closed_loop(x0, np.ascontiguousarray(P.T), N, k_sim, 10, 1e-14, flags, want_Uk=True)
    tg = time.time() - t0
    c = c_oracle.closed_loop_batch(phys, x0, N, k_sim, 10, 1e-14, flags, threads, want_Uk=True)
    umax = np.broadcast_to(np.asarray(phys["umax"], dtype=float), (S,))
    du = np.max(np.abs(g["uk"] - c["uk"]), axis=1) / umax
    dU = np.max(np.abs(g["Uk"] - c["Uk"]).reshape(S, -1), axis=1) / umax
    wref = np.maximum(np.max(np.abs(c["xk"][:, :, 0]), axis=1), 1e-3)
    dw = np.max(np.abs(g["xk"][:, :, 0] - c["xk"][:, :, 0]), axis=1) / wref
    print(f"config{cfg} N={N} S={S} flags={flags}: max du={du.max():.2e} dU={dU.max():.2e} dw={dw.max():.2e} "
          f"bad(>1e-6) u:{(du > 1e-6).sum()} U:{(dU > 1e-6).sum()} w:{(dw > 1e-6).sum()} status gpu max={g['status'].max()} "
          f"oracle max={c['status'].max()} qp_iters/inner gpu={g['qp_iters'].sum() / g['inner_iters'].sum():.2f} "
          f"oracle={c['qp_iters'].sum() / c['inner_iters'].sum():.2f} inner eq={np.mean(g['inner_iters'] == c['inner_iters']):.3f} gpu_s={tg:.2f}", flush=True)


compare(5, S_par)
compare(5, min(S_par, 64), flags=0)
for N in (33, 40, 64, 65, 72, 128):
    compare(3, 64, N=N, k_sim=8)
    compare(4, 64, N=N, k_sim=8)

import torch
dev = torch.device("cuda:0")
P, x0, N = physics.batch_params(5, S=S_tim)
dx = torch.from_numpy(x0).to(dev); dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev)
xk = torch.empty((S_tim, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S_tim, 20), dtype=torch.float64, device=dev)
inn = torch.empty((S_tim, 20), dtype=torch.int32, device=dev); qp = torch.empty((S_tim, 20), dtype=torch.int32, device=dev)
st = torch.empty((S_tim,), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    mpc.closed_loop_dev(S_tim, N, 20, 10, 1e-14, 16, 0, dx.data_ptr(), dP.data_ptr(), S_tim, xk.data_ptr(), uk.data_ptr(), 0, 0,
                        inn.data_ptr(), qp.data_ptr(), st.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"config5 S={S_tim}: {ms:.1f} ms = {S_tim * 20 / ms * 1e3:.0f} scenario-steps/s, qp iters/inner {qp.sum().item() / inn.sum().item():.2f}, "
          f"status max {st.max().item()}", flush=True)
mpc.reset_stream()
