"""A plain-C program linked against libntm_mpc.so: the boundary is usable without Python, torch or C++."""
import os
import subprocess

import numpy as np
import pytest

from oracle import ntm_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mpc-ntm-control_b200", "lib")
EXE = os.path.join(LIBDIR, "c_abi_smoke")


def _build():
    import __graft_entry__ as g
    g.build()
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c"),
                           "-o", EXE, "-L", LIBDIR, "-lntm_mpc", "-lm", f"-Wl,-rpath,{LIBDIR}"])


def test_c_host_links_and_fails_loudly_without_gpu():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_host_reproduces_the_default_scenario():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("k ")]
    uk = np.array([float(x[3]) for x in rows]); w = np.array([float(x[5]) for x in rows])
    ref = o.closed_loop(o.default_physics(), o.default_x0(), N=3, profile=o.LITERAL_FIXED)
    assert np.max(np.abs(uk - ref["uk"])) <= 1e-6 * 2e6
    assert np.max(np.abs(w - ref["xk"][0, 1:])) <= 1e-6 * max(np.max(np.abs(ref["xk"][0])), 1e-3)
