"""CPU tests of the oracle's restatement of the two loop options beyond the box QP: the RK4 plant and
getWLc's state rows kept in the QP of NTM_MPC_Sim.m:97."""
import dataclasses

import numpy as np

from oracle import c_oracle as co
from oracle import ntm_oracle as o

RK4 = dataclasses.replace(o.LITERAL_FIXED, plant_integrator=o.PLANT_RK4)


def test_profile_flag_bits_match_the_header():
    import re, os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "ntm_mpc.h")).read()
    bit = int(re.search(r"NTM_PROFILE_PLANT_RK4\s*=\s*(\d+)", hdr).group(1))
    assert RK4.flags() == o.LITERAL_FIXED.flags() | bit
    assert re.search(r"NTM_STATE_ROWS_OFF = 0, NTM_STATE_ROWS_REFRESH = 1, NTM_STATE_ROWS_FROZEN = 2", hdr)
    assert (o.STATE_ROWS_OFF, o.STATE_ROWS_REFRESH, o.STATE_ROWS_FROZEN) == (0, 1, 2)


def test_rk4_plant_is_fourth_order_on_the_continuous_model():
    """The Euler map of NTM_MPC_Sim.m:130 is x + g(x,u); integrate dx/dt = g/Ts with a tight adaptive solver and check
    that the RK4 option is far closer to it than the Euler map (an unstable, fast scenario makes the gap visible)."""
    from scipy.integrate import solve_ivp
    p = o.default_physics()
    prof_e = dataclasses.replace(o.LITERAL, plant_affine=o.PLANT_WITH_C)
    prof_r = dataclasses.replace(prof_e, plant_integrator=o.PLANT_RK4)
    x = np.array([0.03, 3000.0]); u = 1.5e6
    f = lambda t, z: o.plant_step(p, z, u, prof_e) - z                 # g(x,u), time in units of Ts
    exact = solve_ivp(f, (0.0, 1.0), x, rtol=1e-12, atol=1e-14).y[:, -1]
    e = o.plant_step(p, x, u, prof_e); r = o.plant_step(p, x, u, prof_r)
    err_e = np.max(np.abs(e - exact) / np.abs(exact)); err_r = np.max(np.abs(r - exact) / np.abs(exact))
    assert err_r < 0.05 * err_e and err_e > 1e-6, (err_e, err_r)      # measured: Euler 22 %, RK4 0.16 % (stiff omega equation)


def test_rk4_closed_loop_python_and_c_restatements_agree():
    phys, x0, N = o.make_batch(3, S=6)
    c = co.closed_loop_batch(phys, x0, N, flags=RK4.flags())
    e = co.closed_loop_batch(phys, x0, N, flags=o.LITERAL_FIXED.flags())
    for s in range(3):
        r = o.closed_loop(o.scenario(phys, s), x0[s], N=N, profile=RK4)
        assert np.max(np.abs(r["uk"] - c["uk"][s])) <= 1e-6 * 2e6
        assert np.max(np.abs(r["xk"].T - c["xk"][s]) / np.abs(c["xk"][s]).max(axis=0)) <= 1e-9
    assert np.max(np.abs(c["xk"] - e["xk"])) > 0.0


# ------------------------------------------------------------------ tau_E(w) and Cw hooks
TAUE = dataclasses.replace(o.LITERAL_FIXED, tau_e_model=o.TAUE_W)


def test_taue_flag_bit_and_param_slot_match_the_header():
    import re, os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "ntm_mpc.h")).read()
    bit = int(re.search(r"NTM_PROFILE_TAUE_W\s*=\s*(\d+)", hdr).group(1))
    assert TAUE.flags() == o.LITERAL_FIXED.flags() | bit
    assert re.search(r"\[15\] c_tauE", hdr) and o.PARAM_NAMES[15] == "c_tauE"
    p = o.default_physics()
    assert o.derive_params(p)[15] == 0.0
    p["c_tauE"] = o.c_tauE_belt(p)
    assert abs(p["c_tauE"] - 4 * 1.55 ** 3 / 2.0 ** 4) < 1e-15 and o.derive_params(p)[15] == p["c_tauE"]


def test_taue_hook_is_the_script_when_the_coefficient_is_zero_and_degrades_confinement_otherwise():
    p = o.default_physics()
    x0 = np.array([0.08, 2000 * np.pi])
    a = o.closed_loop(p, x0, N=5, profile=o.LITERAL_FIXED)
    b = o.closed_loop(p, x0, N=5, profile=TAUE)                      # c_tauE = 0: bit-identical to NTM_MPC_Sim.m:14
    assert np.array_equal(a["xk"], b["xk"]) and np.array_equal(a["uk"], b["uk"])
    p["c_tauE"] = o.c_tauE_belt(p)
    assert o.tau_E_of(p, 0.1, o.TAUE_W) == p["tau_E0"] * (1 - p["c_tauE"] * 0.1) < p["tau_E0"]
    assert o.tau_E_of(p, 0.1, o.TAUE_CONST) == p["tau_E0"]
    # plant: the (2,2) entry of A.m:2 with TE = tau_E(w) of the state the map is evaluated at
    x = np.array([0.1, 3000.0])
    e0 = o.plant_step(p, x, 1e6, o.LITERAL); e1 = o.plant_step(p, x, 1e6, dataclasses.replace(o.LITERAL, tau_e_model=o.TAUE_W))
    assert e0[0] == e1[0]                                            # the width equation has no tau_E
    assert np.isclose(e1[1] - e0[1], (p["Ts"] / p["tau_E0"] - p["Ts"] / o.tau_E_of(p, x[0], o.TAUE_W)) * x[1], rtol=1e-9)
    c = o.closed_loop(p, x0, N=5, profile=TAUE)
    assert np.max(np.abs(c["xk"] - a["xk"])) > 0.0


def test_taue_closed_loop_python_and_c_restatements_agree():
    phys, x0, N = o.make_batch(3, S=6, sample={"c_tauE": (0.3, 1.5), "Cw": (0.5, 2.0)})
    assert phys["c_tauE"].min() >= 0.3 and phys["Cw"].max() <= 2.0 and np.ptp(phys["Cw"]) > 0
    base, _, _ = o.make_batch(3, S=6)
    assert np.array_equal(base["tau_r"], phys["tau_r"]) and np.array_equal(base["Cw"], np.ones(6))   # base draws untouched
    for prof in (TAUE, dataclasses.replace(TAUE, plant_integrator=o.PLANT_RK4)):
        c = co.closed_loop_batch(phys, x0, N, flags=prof.flags())
        e = co.closed_loop_batch(phys, x0, N, flags=prof.flags() & ~128)
        for s in range(3):
            r = o.closed_loop(o.scenario(phys, s), x0[s], N=N, profile=prof)
            assert np.max(np.abs(r["uk"] - c["uk"][s])) <= 1e-6 * 2e6
            assert np.max(np.abs(r["xk"].T - c["xk"][s]) / np.abs(c["xk"][s]).max(axis=0)) <= 1e-9
        assert np.max(np.abs(c["xk"] - e["xk"])) > 0.0


def test_host_sampler_hook_matches_the_oracle_twin():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "mpc-ntm-control_b200"))
    from ntm_mpc import physics
    smp = {"Cw": (0.5, 2.0), "c_tauE": (0.0, 1.5)}
    ph, x0h, _ = physics.make_batch(4, S=40, sample=smp)
    po, x0o, _ = o.make_batch(4, S=40, sample=smp)
    assert np.array_equal(x0h, x0o)
    for k in po:
        assert np.array_equal(np.asarray(ph[k]), np.asarray(po[k])), k
    assert np.array_equal(physics.params_from_physics(ph), o.derive_params_batch(po))
    assert physics.c_tauE_belt(physics.nominal()) == o.c_tauE_belt(o.default_physics())


XB = (0.05, 0.16, 2000.0, 12000.0)


def test_state_rows_wide_box_is_the_box_loop():
    phys, x0, N = o.make_batch(3, S=2)
    wide = (-1e3, 1e3, -1e9, 1e9)
    for s in range(2):
        a = o.closed_loop(o.scenario(phys, s), x0[s], N=10, k_sim=4, i_sim=3, profile=o.LITERAL_FIXED,
                          state_rows=o.STATE_ROWS_REFRESH, xbounds=wide)
        b = o.closed_loop(o.scenario(phys, s), x0[s], N=10, k_sim=4, i_sim=3, profile=o.LITERAL_FIXED)
        assert np.allclose(a["uk"], b["uk"], rtol=0, atol=1e-6 * 2e6) and a["status"] == b["status"]


def test_state_rows_outcomes_and_infeasible_fill():
    phys, x0, _ = o.make_batch(3, S=12)
    seen = set()
    for mode in (o.STATE_ROWS_REFRESH, o.STATE_ROWS_FROZEN):
        for s in range(12):
            r = o.closed_loop(o.scenario(phys, s), x0[s], N=10, k_sim=6, i_sim=2, profile=o.LITERAL_FIXED,
                              state_rows=mode, xbounds=XB)
            seen.add(r["status"])
            if r["status"] == o.QP_INFEASIBLE:
                k = int(np.argmax(np.isnan(r["uk"])))
                assert np.all(np.isnan(r["uk"][k:])) and np.all(np.isnan(r["xk"][:, k + 1:])) and np.isnan(r["cost"])
                assert np.all(np.isfinite(r["xk"][:, :k + 1])) and r["inner_iters"][k] >= 1
                assert np.all(r["inner_iters"][k + 1:] == 0)
                if not (XB[0] <= x0[s][0] <= XB[1] and XB[2] <= x0[s][1] <= XB[3]):
                    assert k == 0                                       # the x_0 block of getWLc.m:30
            else:
                # consistent first prediction = refreshed model at x_k: states stay inside the box they were asked to
                assert np.all(np.isfinite(r["xk"]))
    assert o.QP_OK in seen and o.QP_INFEASIBLE in seen


def test_frozen_rows_differ_from_refreshed_rows():
    """:74 is outside the loops: with FROZEN the rows keep rho(x0); once the scheduling moves the two answers differ."""
    phys, x0, _ = o.make_batch(3, S=12)
    diff = 0.0
    for s in range(12):
        a = o.closed_loop(o.scenario(phys, s), x0[s], N=10, k_sim=6, i_sim=2, profile=o.LITERAL_FIXED,
                          state_rows=o.STATE_ROWS_REFRESH, xbounds=XB)
        b = o.closed_loop(o.scenario(phys, s), x0[s], N=10, k_sim=6, i_sim=2, profile=o.LITERAL_FIXED,
                          state_rows=o.STATE_ROWS_FROZEN, xbounds=XB)
        both = ~np.isnan(a["uk"]) & ~np.isnan(b["uk"])
        if both.any():
            diff = max(diff, float(np.max(np.abs(a["uk"][both] - b["uk"][both]))))
    assert diff > 1.0
