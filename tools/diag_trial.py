"""Diagnose one scenario of a stress trial: where do GPU and oracle trajectories part, and does the GPU QP solve the
oracle's QP at that point?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np
import ntm_mpc
from oracle import c_oracle as co, ntm_oracle as o
np.set_printoptions(linewidth=220, precision=6)
seed0, trial, scen = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(seed0)
for t in range(trial + 1):
    cfg = int(rng.choice([2, 3, 4])); N = int(rng.choice([1, 2, 3, 5, 8, 10, 13, 16, 20, 24, 31, 32, 33, 40, 48])); S = int(rng.integers(1, 200))
    k_sim = int(rng.integers(1, 12)); i_sim = int(rng.integers(1, 11)); flags = 0
    if rng.random() < 0.3: flags |= 1
    if rng.random() < 0.25: flags |= 2
    if rng.random() < 0.3: flags |= 4
    if rng.random() < 0.3: flags |= 8
    if rng.random() < 0.5: flags |= 16
    if rng.random() < 0.15 and not (flags & 2): flags |= 32
    seed = int(rng.integers(1, 1 << 30))
print("cfg", cfg, "N", N, "S", S, "k_sim", k_sim, "i_sim", i_sim, "flags", flags, "seed", seed)
phys, x0, _ = o.make_batch(cfg, S=S, seed=seed)
P = o.derive_params_batch(phys)
mpc = ntm_mpc.NtmMpc(0)
g = mpc.closed_loop(x0, P.T, N=N, k_sim=k_sim, i_sim=i_sim, profile=flags, want_Uk=True)
c = co.closed_loop_batch(phys, x0, N, k_sim=k_sim, i_sim=i_sim, flags=flags & 31, want_Uk=True)
s = scen
print("gpu uk", g["uk"][s]); print("orc uk", c["uk"][s])
print("gpu inner", g["inner_iters"][s], "qp", g["qp_iters"][s]); print("orc inner", c["inner_iters"][s], "qp", c["qp_iters"][s])
print("gpu w", g["xk"][s, :, 0]); print("orc w", c["xk"][s, :, 0])
for k in range(k_sim):
    d = np.abs(g["Uk"][s, k] - c["Uk"][s, k]).max() / phys["umax"][s]
    if d > 1e-9:
        print("first Uk difference at step", k, "rel", d)
        print(" gpu Uk", g["Uk"][s, k]); print(" orc Uk", c["Uk"][s, k]); break
# numpy oracle trace for this scenario: QPs of the diverging step, re-solved by the GPU QP kernel
prof = o.Profile(rho1_variant=flags & 1, gamma_index=(flags >> 1) & 1, f_state=(flags >> 2) & 1, plant_affine=(flags >> 3) & 1, inner_policy=(flags >> 4) & 1)
tr = []
r = o.closed_loop(o.scenario(phys, s), x0[s], N=N, k_sim=k_sim, i_sim=i_sim, profile=prof, trace=tr)
print("numpy oracle uk", r["uk"])
Gs = np.stack([t_["G"] for t_ in tr]); Fs = np.stack([t_["F"] for t_ in tr])
U, it, st = mpc.qp_box(Gs, Fs, phys["umin"][s], phys["umax"][s])
worst = 0
for q in range(len(tr)):
    Uo, _, _ = o.qp_box(Gs[q], Fs[q], phys["umin"][s], phys["umax"][s])
    e = np.abs(U[q] - Uo).max() / phys["umax"][s]; kkt = o.qp_kkt_residual(Gs[q], Fs[q], phys["umin"][s], phys["umax"][s], U[q])
    obj = lambda u: 0.5 * u @ Gs[q] @ u + Fs[q] @ u
    if e > 1e-7: print(" QP", q, "(k,it)=", tr[q]["k"], tr[q]["it"], "gpu-vs-oracle", e, "kkt gpu", kkt, "kkt orc", o.qp_kkt_residual(Gs[q], Fs[q], phys["umin"][s], phys["umax"][s], Uo), "obj diff rel", (obj(U[q]) - obj(Uo)) / abs(obj(Uo)), "cond", np.linalg.cond(Gs[q]))
    worst = max(worst, e)
print("worst QP solution difference on the oracle's own (G,F) sequence:", worst, "status", st.max())
