"""The figure of NTM_MPC_Sim.m:134-161 as an SVG (host-side, no GPU)."""
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from oracle import ntm_oracle as o


def test_default_scenario_figure_has_the_scripts_curves_and_labels(tmp_path):
    from ntm_mpc import plots
    r = o.closed_loop(o.default_physics(), o.default_x0(), N=3, profile=o.LITERAL_FIXED)
    p = tmp_path / "fig.svg"
    svg = plots.trajectory_svg(r["xk"], r["uk"][None, :], str(p))
    root = ET.fromstring(p.read_text(encoding="utf-8"))
    ns = "{http://www.w3.org/2000/svg}"
    lines = [e for e in root.iter(ns + "polyline") if e.get("class") == "stairs"]
    assert len(lines) == 3                                        # w, omega (:141) and P_ECCD (:155)
    npts = [len(e.get("points").split()) for e in lines]
    assert npts == [2 * 21 - 1, 2 * 21 - 1, 2 * 20 - 1]           # stairs: every sample but the first adds a riser
    text = " ".join(t.text or "" for t in root.iter(ns + "text"))
    for label in ("w [m]", "[Hz]", "P_ECCD [W]", "Constrained quasi-LPV MPC State and Input Trajectory", "k", "u"):
        assert label in text
    assert svg == p.read_text(encoding="utf-8")


def test_batched_result_and_nan_tail(tmp_path):
    from ntm_mpc import plots
    xk = np.zeros((4, 6, 2)); uk = np.ones((4, 5))
    xk[2, 3:, :] = np.nan; uk[2, 2:] = np.nan                     # a scenario that turned infeasible at step 2
    svg = plots.plot_result(dict(xk=xk, uk=uk), scenario=2)
    root = ET.fromstring(svg)
    lines = [e for e in root.iter("{http://www.w3.org/2000/svg}polyline")]
    assert [len(e.get("points").split()) for e in lines] == [5, 5, 3]
    with pytest.raises(ValueError):
        plots.trajectory_svg(np.zeros((3, 7)), np.zeros(6))
    with pytest.raises(ValueError):
        plots.trajectory_svg(np.zeros((2, 7)), np.zeros(5))
