"""Throughput of ntm_qp_ineq on MPC-shaped problems (run on the GPU box): config-3 scenarios, N = 20, first QP of a
closed-loop step with a perturbed scheduling sequence, constraints = getWLc rows with the script's own bounds.
Everything is built through the C ABI (condense -> hessian_grad -> getWLc), rows split as the quadprog shim does."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import ntm_mpc
from ntm_mpc import physics, reference_api
import torch

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
h = ntm_mpc.NtmMpc(0)
prm, x0, _ = physics.batch_params(3, S)
prm = np.ascontiguousarray(prm.T)
rng = np.random.default_rng(1)
R = [np.empty((S, N)) for _ in range(3)]
for i in range(N):
    xs = x0 * (1 + 0.06 * rng.standard_normal((S, 2)))
    r1, r2, r3 = h.rho(xs, prm)
    R[0][:, i], R[1][:, i], R[2][:, i] = r1, r2, r3
Phi, Gam, Lam = h.condense(R[0], R[1], R[2], prm)
G, F = h.hessian_grad(Phi, Gam, Lam, x0, prm)
W, L, c = h.getWLc([0.15, 5000 * 2 * np.pi], [0.06, 100 * 2 * np.pi], 2e6, 0.0, Gam, Phi, Lam)
b = c + np.einsum("smk,sk->sm", W, x0)
lb, ub, Lg, bg, feas = reference_api.split_rows(L, b)
M = Lg.shape[1]
print(f"S={S} N={N} general rows M={M} host-feasible {feas.mean():.3f}")
t0 = time.time(); U, it, st = h.qp_ineq(G, F, lb, ub, Lg, bg); t_host = time.time() - t0
Ub, itb, stb = h.qp_box(G, F, lb, ub)
changed = np.max(np.abs(U - Ub), axis=1) > 1e-3 * 2e6
print(f"status counts {dict(zip(*np.unique(st, return_counts=True)))}; state rows change the answer in {changed.mean():.3f}; "
      f"mean iterations {it.mean():.1f} (box alone {itb.mean():.1f}), max {it.max()}")
# device-resident timing
dev = torch.device("cuda:0")
def T(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dG = T(np.ascontiguousarray(G.transpose(0, 2, 1))); dF = T(F); dlb = T(lb); dub = T(ub)
dL = T(np.ascontiguousarray(Lg.transpose(0, 2, 1))); dbg = T(bg)
dU = torch.empty(S, N, dtype=torch.float64, device=dev); dit = torch.empty(S, dtype=torch.int32, device=dev); dst = torch.empty_like(dit)
lib = h._lib
h.set_stream(None)
def run(ineq):
    if ineq:
        rc = lib.ntm_qp_ineq_dev(h._h, 0, S, N, M, dG.data_ptr(), dF.data_ptr(), dlb.data_ptr(), dub.data_ptr(), S, dL.data_ptr(), dbg.data_ptr(), dU.data_ptr(), dit.data_ptr(), dst.data_ptr())
    else:
        rc = lib.ntm_qp_box_dev(h._h, 0, S, N, dG.data_ptr(), dF.data_ptr(), dlb.data_ptr(), dub.data_ptr(), S, dU.data_ptr(), dit.data_ptr(), dst.data_ptr())
    assert rc == 0
for ineq in (0, 1):
    for _ in range(3): run(ineq)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run(ineq)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(("ntm_qp_ineq" if ineq else "ntm_qp_box "), f"{ms:.3f} ms per batch = {S / ms / 1e3:.2f} M QPs/s")
assert np.array_equal(dU.cpu().numpy(), U)
