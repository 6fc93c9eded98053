"""Monte-Carlo back end: the on-device reduction ntm_mc_stats against the NumPy statistics of the
oracle, the end-to-end ``montecarlo.run`` (SoA layout, device-resident between loop and reduction), and the
trajectory writers.  Counts must match exactly; sums to 1e-12 (different summation order)."""
import os

import numpy as np
import pytest

from oracle import ntm_oracle as o

BOX = (0.06, 0.15, 100 * 2 * np.pi, 5000 * 2 * np.pi)
COUNTS = [0, 1, 2, 3, 12, 13, 14, 15, 16, 17, 19, 20, 21] + list(range(22, 22 + o.MC_NBINS))
SUMS = [4, 5, 6, 7, 8, 9, 10, 11, 18]


def _compare(got, exp):
    assert np.array_equal(got[COUNTS], exp[COUNTS]), (got[COUNTS], exp[COUNTS])
    for i in SUMS:
        assert got[i] == pytest.approx(exp[i], rel=1e-12, abs=1e-300), i


@pytest.fixture(scope="module")
def mpc():
    import ntm_mpc
    h = ntm_mpc.NtmMpc(0)
    yield h
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,S", [(3, 4096), (4, 3000), (2, 1024)])
def test_mc_stats_match_numpy(mpc, cfg, S):
    from ntm_mpc import physics
    prm, x0, N = physics.batch_params(cfg, S)
    prm = np.ascontiguousarray(prm.T)
    x0 = x0.copy()
    x0[5, 1] = 0.0                                               # omega = 0: rho2 divides by zero -> a non-finite scenario
    r = mpc.closed_loop(x0, prm, N, 20, 10, 1e-14, o.LITERAL_FIXED.flags())
    assert r["status"][5] == 2
    got = mpc.mc_stats(r["xk"], r["uk"], r["cost"], r["status"], prm, BOX, 0.06, 0.2)
    exp = o.mc_stats(r["xk"], r["uk"], r["cost"], r["status"], prm[:, 8], prm[:, 9], BOX, 0.06, 0.2)
    _compare(got, exp)
    assert got[2] >= 1 and got[0] + got[1] + got[2] + got[3] == S
    if cfg == 4:
        assert 0.0 < got[16] / got[17] < 1.0                     # config 4: the sampled umax binds on part of the steps


@pytest.mark.gpu
@pytest.mark.parametrize("K,S", [(1, 33), (7, 1000), (20, 31), (64, 257), (400, 70)])
def test_mc_stats_shapes_layouts_and_all_status_codes(mpc, K, S):
    """Synthetic results with every status code, ragged S and odd / long trajectories (k_sim = 400 takes the
    lane-per-sample fallback of the MATLAB-layout kernel), both layouts through the device entry point."""
    import torch
    from ntm_mpc import LAYOUT_MATLAB, LAYOUT_SOA, physics
    rng = np.random.default_rng(K * 1000 + S)
    xk = rng.uniform(0.0, 0.2, (S, K + 1, 2)); xk[:, :, 1] = rng.uniform(0.0, 4e4, (S, K + 1))
    umax = rng.uniform(0.5e6, 2e6, S)
    uk = rng.uniform(0.0, 1.0, (S, K)) * umax[:, None]
    uk[rng.random((S, K)) < 0.3] = 0.0
    hit = rng.random((S, K)) < 0.3
    uk[hit] = np.broadcast_to(umax[:, None], (S, K))[hit]
    cost = rng.uniform(0.0, 10.0, S)
    status = rng.integers(0, 4, S).astype(np.int32)
    bad = status >= 2
    xk[bad, 1:, :] = np.nan; uk[bad] = np.nan; cost[bad] = np.nan
    prm = np.zeros((S, 16)); prm[:, 9] = umax
    exp = o.mc_stats(xk, uk, cost, status, 0.0, umax, BOX, 0.06, 0.2)
    got = mpc.mc_stats(xk, uk, cost, status, prm, BOX, 0.06, 0.2)
    _compare(got, exp)
    dev = torch.device("cuda:0")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = torch.empty(64, dtype=torch.float64, device=dev)
    mpc.set_stream(torch.cuda.current_stream(dev).cuda_stream or None)
    try:
        t = [d(xk.reshape(S, -1).T), d(uk.T), d(cost), d(status), d(prm.T)]
        mpc.mc_stats_dev(S, K, LAYOUT_SOA, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), S,
                         out.data_ptr(), BOX, 0.06, 0.2)
        soa = out.cpu().numpy()[:len(exp)]
    finally:
        mpc.reset_stream()
    _compare(soa, exp)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1], ids=["matlab", "soa"])
def test_mc_stats_with_compact_bound_arrays(mpc, layout):
    """ntm_mc_stats_ub_dev (umin / umax as two compact arrays, or one shared pair) = ntm_mc_stats_dev, value for value."""
    import torch
    rng = np.random.default_rng(5)
    S, K = 3001, 20
    xk = rng.uniform(0.0, 0.2, (S, K + 1, 2)); xk[:, :, 1] = rng.uniform(0.0, 4e4, (S, K + 1))
    umax = rng.uniform(0.5e6, 2e6, S); umin = np.zeros(S)
    uk = rng.uniform(0.0, 1.0, (S, K)) * umax[:, None]; uk[rng.random((S, K)) < 0.3] = 0.0
    cost = rng.uniform(0.0, 10.0, S); status = rng.integers(0, 2, S).astype(np.int32)
    prm = np.zeros((S, 16)); prm[:, 9] = umax
    dev = torch.device("cuda:0")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    if layout == 0:
        t = [d(xk), d(uk), d(cost), d(status), d(prm)]
    else:
        t = [d(xk.reshape(S, -1).T), d(uk.T), d(cost), d(status), d(prm.T)]
    dmin, dmax = d(umin), d(umax)
    o1 = torch.empty(64, dtype=torch.float64, device=dev); o2 = torch.empty(64, dtype=torch.float64, device=dev)
    mpc.set_stream(torch.cuda.current_stream(dev).cuda_stream or None)
    try:
        mpc.mc_stats_dev(S, K, layout, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), S, o1.data_ptr(), BOX, 0.06, 0.2)
        mpc.mc_stats_ub_dev(S, K, layout, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), dmin.data_ptr(), dmax.data_ptr(), S,
                            o2.data_ptr(), BOX, 0.06, 0.2)
        a, b = o1.cpu().numpy()[:54], o2.cpu().numpy()[:54]
    finally:
        mpc.reset_stream()
    _compare(b, a)
    _compare(a, o.mc_stats(xk, uk, cost, status, 0.0, umax, BOX, 0.06, 0.2))


@pytest.mark.gpu
def test_montecarlo_run_is_device_resident_and_matches(mpc, tmp_path):
    from ntm_mpc import montecarlo, physics
    res = montecarlo.run(config=3, S=8192, profile=o.LITERAL_FIXED.flags(), trajectories=True, handle=mpc)
    prm, x0, N = physics.batch_params(3, 8192)
    exp = o.mc_stats(res["xk"], res["uk"], res["cost"], res["status"], prm[8], prm[9], BOX, 0.06, 0.2)
    _compare(res["stats"], exp)
    ref = mpc.closed_loop(x0, np.ascontiguousarray(prm.T), N, 20, 10, 1e-14, o.LITERAL_FIXED.flags())
    assert np.array_equal(ref["xk"], res["xk"]) and np.array_equal(ref["uk"], res["uk"])     # SoA path = MATLAB path
    assert res["scenarios_ok"] + res["scenarios_iter_cap"] == 8192
    assert 0.0 <= res["active_bound_fraction"] <= 1.0 and res["w_final_hist"].sum() == 8192
    montecarlo.save_npz(str(tmp_path / "mc.npz"), res)
    montecarlo.save_mat(str(tmp_path / "mc.mat"), res, scenario=3)
    from scipy.io import loadmat
    m = loadmat(str(tmp_path / "mc.mat"))
    assert m["xk"].shape == (2, 21) and m["uk"].shape == (1, 20)                              # NTM_MPC_Sim.m:82-83
    assert np.array_equal(m["xk"], res["xk"][3].T)
    z = np.load(str(tmp_path / "mc.npz"))
    assert np.array_equal(z["uk"], res["uk"])


def test_numpy_statistics_on_a_hand_made_batch():
    xk = np.zeros((3, 3, 2)); uk = np.zeros((3, 2))
    xk[0, :, 0] = [0.10, 0.08, 0.05]; xk[1, :, 0] = [0.10, 0.12, 0.16]; xk[2, :, 0] = np.nan
    xk[:, :, 1] = 6283.0
    uk[0] = [0.0, 2e6]; uk[1] = [1e6, 5e5]; uk[2] = np.nan
    s = o.mc_stats(xk, uk, np.array([1.0, 3.0, np.nan]), np.array([0, 1, 2]), 0.0, 2e6, BOX, 0.06, 0.2)
    assert list(s[:4]) == [1, 1, 1, 0]
    assert s[4] == 4.0 and s[5] == 10.0 and s[6] == 1.0 and s[7] == 3.0
    assert s[8] == pytest.approx(0.21) and s[10] == 0.05 and s[11] == 0.16
    assert s[12] == 1 and s[13] == 2 and s[14] == 1                 # scenario 0 drops below 0.06 at step 2
    assert s[15] == 1 and s[16] == 1 and s[17] == 4 and s[18] == 3.5e6
    assert s[19] == 2 and s[20] == 0 and s[21] == 4                  # w = 0.05 and w = 0.16 are outside [0.06, 0.15]
    assert s[22 + 8] == 1 and s[22 + 25] == 1 and s[22:].sum() == 2  # 0.05 -> bin 8, 0.16 -> bin 25 of 32 over [0, 0.2)


def _describe_vs_numpy(stats, cost, xk, status):
    from ntm_mpc import montecarlo
    d = montecarlo.describe(stats)
    inc = status < 2                                                 # what the reduction includes: OK and iteration-cap
    assert inc.sum() == d["scenarios_ok"] + d["scenarios_iter_cap"] and inc.sum() > 0
    assert d["cost_mean"] == pytest.approx(np.mean(cost[inc]), rel=1e-12)
    assert d["cost_std"] == pytest.approx(np.std(cost[inc]), rel=1e-7, abs=1e-9)
    assert d["w_final_mean"] == pytest.approx(np.mean(xk[inc, -1, 0]), rel=1e-12)
    assert d["w_final_std"] == pytest.approx(np.std(xk[inc, -1, 0]), rel=1e-7, abs=1e-12)
    assert d["suppressed_fraction"] == pytest.approx(np.mean(xk[inc, -1, 0] < 0.06), rel=1e-12)
    reached = np.any(xk[inc, 1:, 0] < 0.06, axis=1)
    assert d["reached_suppression_fraction"] == pytest.approx(np.mean(reached), rel=1e-12)


def test_describe_means_leave_infeasible_scenarios_out():
    """Moments are normalised by the scenarios that ENTERED the sums (status 0/1): infeasible ones (status 3, routine
    with state rows: the x_0 block rejects any xk outside the box) are counted but excluded, like non-finite ones."""
    rng = np.random.default_rng(7)
    S, K = 200, 6
    xk = rng.uniform(0.03, 0.16, (S, K + 1, 2)); xk[:, :, 1] = 6283.0
    uk = rng.uniform(0.0, 2e6, (S, K)); cost = rng.uniform(1.0, 9.0, S)
    status = rng.choice([0, 0, 0, 1, 2, 3, 3], S).astype(np.int32)
    bad = status >= 2
    xk[bad, 1:, :] = np.nan; uk[bad] = np.nan; cost[bad] = np.nan
    assert (status == 3).sum() > 10
    _describe_vs_numpy(o.mc_stats(xk, uk, cost, status, 0.0, 2e6, BOX, 0.06, 0.2), cost, xk, status)


@pytest.mark.gpu
def test_describe_on_a_gpu_batch_with_infeasible_scenarios(mpc):
    """State rows kept (frozen, the literal reading): part of the sample ends infeasible; describe() of the on-device
    reduction must equal NumPy means over the scenarios that remain."""
    from ntm_mpc import STATE_ROWS_FROZEN, physics
    prm, x0, N = physics.batch_params(3, 2048)
    prm = np.ascontiguousarray(prm.T)
    r = mpc.closed_loop(x0, prm, N, 20, 10, 1e-14, o.LITERAL_FIXED.flags(), state_rows=STATE_ROWS_FROZEN)
    assert (r["status"] == 3).sum() > 50 and (r["status"] < 2).sum() > 50
    got = mpc.mc_stats(r["xk"], r["uk"], r["cost"], r["status"], prm, BOX, 0.06, 0.2)
    _describe_vs_numpy(got, r["cost"], r["xk"], r["status"])


def test_writers_round_trip_without_a_gpu(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpc-ntm-control_b200"))
    from ntm_mpc import montecarlo
    rng = np.random.default_rng(0)
    res = {"xk": rng.random((4, 21, 2)), "uk": rng.random((4, 20)), "cost": rng.random(4), "stats": np.arange(54.0)}
    res.update({k: v for k, v in montecarlo.describe(np.arange(54.0) + 1).items() if np.ndim(v) == 0})
    montecarlo.save_mat(str(tmp_path / "b.mat"), res)
    from scipy.io import loadmat
    m = loadmat(str(tmp_path / "b.mat"))
    assert m["xk"].shape == (2, 21, 4) and m["uk"].shape == (20, 4)
    assert np.array_equal(m["xk"][:, :, 2], res["xk"][2].T)
    montecarlo.save_npz(str(tmp_path / "b.npz"), res)
    assert np.array_equal(np.load(str(tmp_path / "b.npz"))["xk"], res["xk"])
