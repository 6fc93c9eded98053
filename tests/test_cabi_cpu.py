"""CPU-side checks of the drop-in boundary and the host logic (no GPU, no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ntm_mpc.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from ntm_mpc import _lib
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ntm_[a-z0-9_A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from ntm_mpc import _lib
    names = declared_symbols()
    assert len(names) >= 22
    for n in names:
        assert hasattr(lib, n), f"libntm_mpc.so does not export {n}"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", out), n


def test_version_and_loud_failure_without_gpu(lib):
    import torch
    assert lib.ntm_version() == 100
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.ntm_create(ctypes.byref(h), 0)
    assert rc == 2 and not h.value                                  # NTM_ERR_CUDA, never a silent CPU path
    assert b"no CPU fallback" in lib.ntm_last_error()
    import ntm_mpc
    with pytest.raises(ntm_mpc.NtmError):
        ntm_mpc.NtmMpc(0)
    with pytest.raises(ntm_mpc.NtmError):
        ntm_mpc.rho2([0.1, 1.0])


def test_library_is_sm100a_only():
    from ntm_mpc import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mpc-ntm-control_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "_lib.py", os.path.join(dp, f)


def test_host_physics_matches_oracle_formulas():
    from ntm_mpc import physics
    from oracle import ntm_oracle as o
    assert np.array_equal(physics.params_from_physics(physics.nominal()), o.derive_params(o.default_physics()))
    for cfg in (2, 3, 4, 5):
        P, x0, N = physics.batch_params(cfg, S=300)
        phys, x0o, No = o.make_batch(cfg, S=300)
        assert N == No and np.array_equal(x0, x0o) and np.array_equal(P, o.derive_params_batch(phys))
    p = physics.nominal()
    k, z = physics.kappa(p), physics.zeta(p)
    C = physics.affine_C(p)
    alt = physics.params_from_model_constants(k, p["tau_r"], p["Ts"], z, p["rs"], p["a"], p["tau_E0"], p["w_dep"], p["eta_CD"],
                                              p["w_marg"], C, p["umin"], p["umax"], (p["r1"], p["r2"]))
    assert np.array_equal(alt, physics.params_from_physics(p))


def test_block_layout_helpers_roundtrip():
    from ntm_mpc import api
    a = np.arange(2 * 6 * 3, dtype=np.float64).reshape(2, 6, 3)
    flat = api._blocks_in(a, 2, 6, 3)
    assert flat.shape == (2, 3, 6) and flat[1, 2, 4] == a[1, 4, 2]       # column-major inside the block
    assert np.array_equal(api._blocks_out(flat.ravel(), 2, 6, 3), a)


def test_bench_flop_model_matches_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    for N, f in ((3, 299), (10, 2385), (20, 10890), (100, 778530)):
        assert round(bench.flops_per_inner(N)) == f
