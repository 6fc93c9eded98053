"""The figure of NTM_MPC_Sim.m:134-161 ("Plot 4.1.1") without MATLAB: two panels of staircase curves -- the state
trajectory ``stairs(0:k_sim, xk')`` with legend w [m] / omega [Hz] (:141-147) and the input trajectory
``stairs(0:k_sim-1, uk')`` with legend P_ECCD [W] (:155-161) under the title of :159 -- written as a standalone SVG
(no plotting package in the image).  Works on the arrays a closed-loop call returns or on a file written by
``montecarlo.save_npz`` / ``save_mat``.  Host-side post-processing only: nothing here touches the GPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

_COLORS = ("#0072bd", "#d95319", "#edb120", "#7e2f8e")        # MATLAB's default colour order


def _ticks(lo: float, hi: float, n: int = 5) -> List[float]:
    if not (np.isfinite(lo) and np.isfinite(hi)) or hi <= lo:
        return [lo]
    raw = (hi - lo) / n
    mag = 10.0 ** np.floor(np.log10(raw))
    step = min((1.0, 2.0, 5.0, 10.0), key=lambda m: abs(m * mag - raw)) * mag
    first = np.ceil(lo / step) * step
    return [float(v) for v in np.arange(first, hi + 0.5 * step, step)]


def _stairs_points(k0: int, y: np.ndarray) -> List[Tuple[float, float]]:
    """MATLAB ``stairs``: the value at sample k holds until sample k+1; the last sample is a single point."""
    pts: List[Tuple[float, float]] = []
    for i, v in enumerate(y):
        if not np.isfinite(v):
            break
        if pts:
            pts.append((k0 + i, pts[-1][1]))
        pts.append((k0 + i, float(v)))
    return pts


def _panel(x0: float, y0: float, w: float, h: float, series: Sequence[Tuple[str, int, np.ndarray]], xlabel: str,
           ylabel: str, kmax: int) -> List[str]:
    curves = [(name, _stairs_points(k0, np.asarray(y, dtype=np.float64))) for name, k0, y in series]
    ys = [p[1] for _, pts in curves for p in pts] or [0.0]
    lo, hi = min(ys), max(ys)
    if hi <= lo:
        lo, hi = lo - 1.0, hi + 1.0
    pad = 0.05 * (hi - lo)
    lo, hi = lo - pad, hi + pad
    sx = lambda k: x0 + w * k / max(kmax, 1)
    sy = lambda v: y0 + h - h * (v - lo) / (hi - lo)
    out = [f'<rect x="{x0:.1f}" y="{y0:.1f}" width="{w:.1f}" height="{h:.1f}" fill="none" stroke="#222"/>']
    for t in _ticks(lo, hi):
        out.append(f'<line x1="{x0:.1f}" x2="{x0 - 4:.1f}" y1="{sy(t):.1f}" y2="{sy(t):.1f}" stroke="#222"/>')
        out.append(f'<text x="{x0 - 6:.1f}" y="{sy(t) + 3:.1f}" font-size="9" text-anchor="end">{t:.4g}</text>')
    for t in _ticks(0.0, float(kmax)):
        out.append(f'<line x1="{sx(t):.1f}" x2="{sx(t):.1f}" y1="{y0 + h:.1f}" y2="{y0 + h + 4:.1f}" stroke="#222"/>')
        out.append(f'<text x="{sx(t):.1f}" y="{y0 + h + 14:.1f}" font-size="9" text-anchor="middle">{t:.4g}</text>')
    for i, (name, pts) in enumerate(curves):
        col = _COLORS[i % len(_COLORS)]
        if pts:
            path = " ".join(f"{sx(k):.2f},{sy(v):.2f}" for k, v in pts)
            out.append(f'<polyline class="stairs" fill="none" stroke="{col}" stroke-width="1.2" points="{path}"/>')
        out.append(f'<line x1="{x0 + w - 92:.1f}" x2="{x0 + w - 76:.1f}" y1="{y0 + 12 + 12 * i:.1f}" y2="{y0 + 12 + 12 * i:.1f}" stroke="{col}" stroke-width="1.5"/>')
        out.append(f'<text x="{x0 + w - 72:.1f}" y="{y0 + 15 + 12 * i:.1f}" font-size="9">{name}</text>')
    out.append(f'<text x="{x0 + w / 2:.1f}" y="{y0 + h + 30:.1f}" font-size="11" text-anchor="middle" font-style="italic">{xlabel}</text>')
    out.append(f'<text transform="translate({x0 - 44:.1f},{y0 + h / 2:.1f}) rotate(-90)" font-size="10" text-anchor="middle">{ylabel}</text>')
    return out


def trajectory_svg(xk, uk, path: str = None, title: str = "Constrained quasi-LPV MPC State and Input Trajectory") -> str:
    """``xk`` [2, k_sim+1] (or [k_sim+1, 2]), ``uk`` [1, k_sim] (or [k_sim]) of ONE scenario -> SVG text (also written to
    ``path`` when given).  Labels and legends are the script's (:143-147, :157-161)."""
    xk = np.asarray(xk, dtype=np.float64)
    if xk.ndim != 2 or 2 not in xk.shape:
        raise ValueError("xk must be 2 x (k_sim+1) or (k_sim+1) x 2 (one scenario)")
    if xk.shape[0] != 2:
        xk = xk.T
    uk = np.asarray(uk, dtype=np.float64).ravel()
    k_sim = uk.size
    if xk.shape[1] != k_sim + 1:
        raise ValueError("xk needs k_sim+1 samples for k_sim inputs")
    W, H = 860.0, 360.0
    body = [f'<text x="{W / 2:.1f}" y="22" font-size="14" text-anchor="middle">{title}</text>']
    body += _panel(70, 40, 330, 260, [("w [m]", 0, xk[0]), ("ω [Hz]", 0, xk[1])], "k", "x Angle and Frequency Deviation", k_sim)
    body += _panel(500, 40, 330, 260, [("P_ECCD [W]", 0, uk)], "k", "u", max(k_sim - 1, 1))
    svg = (f'<svg xmlns="http://www.w3.org/2000/svg" width="{W:.0f}" height="{H:.0f}" viewBox="0 0 {W:.0f} {H:.0f}" '
           f'font-family="Helvetica, Arial, sans-serif">\n<rect width="100%" height="100%" fill="white"/>\n' + "\n".join(body) + "\n</svg>\n")
    if path is not None:
        with open(path, "w", encoding="utf-8") as f:
            f.write(svg)
    return svg


def plot_result(result: dict, scenario: int = 0, path: str = None) -> str:
    """One scenario of a batched result (``NtmMpc.closed_loop`` / ``montecarlo.run(trajectories=True)``:
    ``xk`` [S, k_sim+1, 2], ``uk`` [S, k_sim])."""
    return trajectory_svg(np.asarray(result["xk"])[scenario].T, np.asarray(result["uk"])[scenario], path)
