"""The MATLAB MEX gateways (csrc/mex/ntm_mex.c) against a mock mex.h runtime (tests/mock_mex): neither MATLAB nor
Octave exists in the build image.  CPU part: the gateways compile, export mexFunction, and reject bad calls the way
MATLAB would ("not enough input arguments") before touching the GPU.  GPU part (-m gpu): every gateway returns what
the reference function returns, in MATLAB's column-major shapes, including the short call forms of NTM_MPC_Sim.m."""
import ctypes
import math
import os

import numpy as np
import pytest

from oracle import ntm_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEXDIR = os.path.join(ROOT, "mpc-ntm-control_b200", "lib", "mex")
NAMES = ["rho1", "rho2", "rho3", "A", "B", "Rho_to_PhiGammaLambda", "ntm_qp_box", "ntm_mpc_batch", "getWLc", "quadprog"]


class MxArray(ctypes.Structure):
    _fields_ = [("m", ctypes.c_size_t), ("n", ctypes.c_size_t), ("pr", ctypes.POINTER(ctypes.c_double)),
                ("is_double", ctypes.c_int), ("is_complex", ctypes.c_int)]


class Mock:
    def __init__(self):
        import __graft_entry__ as g
        g.build()
        self.rt = ctypes.CDLL(os.path.join(MEXDIR, "libmockmex.so"), mode=ctypes.RTLD_GLOBAL)
        self.rt.mxCreateDoubleMatrix.restype = ctypes.POINTER(MxArray)
        self.rt.mxCreateDoubleMatrix.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
        self.rt.mock_last_error_id.restype = ctypes.c_char_p
        self.rt.mock_last_error_msg.restype = ctypes.c_char_p
        self.rt.mock_set_variable.argtypes = [ctypes.c_char_p, ctypes.POINTER(MxArray)]
        self.rt.mock_call_mex.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        self.gw = {}
        self.keep = []

    def gateway(self, name):
        if name not in self.gw:
            self.gw[name] = ctypes.CDLL(os.path.join(MEXDIR, f"{name}.mexmock.so"))
        return self.gw[name]

    def mx(self, a):
        a = np.atleast_2d(np.asarray(a, dtype=np.float64))
        p = self.rt.mxCreateDoubleMatrix(a.shape[0], a.shape[1], 0)
        flat = np.asfortranarray(a).ravel(order="F")
        ctypes.memmove(p.contents.pr, flat.ctypes.data, flat.nbytes)
        self.keep.append(p)
        return p

    def to_np(self, p):
        m, n = p.contents.m, p.contents.n
        return np.ctypeslib.as_array(p.contents.pr, shape=(m * n,)).copy().reshape((m, n), order="F")

    def workspace(self, ws):
        self.rt.mock_clear_workspace()
        for k, v in ws.items():
            self.rt.mock_set_variable(k.encode(), self.mx(v))

    def call(self, name, args, nlhs=1):
        fn = ctypes.cast(self.gateway(name).mexFunction, ctypes.c_void_p)
        prhs = (ctypes.POINTER(MxArray) * max(len(args), 1))(*[self.mx(a) for a in args])
        plhs = (ctypes.POINTER(MxArray) * max(nlhs, 1))()
        rc = self.rt.mock_call_mex(fn, nlhs, ctypes.cast(plhs, ctypes.c_void_p), len(args), ctypes.cast(prhs, ctypes.c_void_p))
        if rc:
            raise RuntimeError(self.rt.mock_last_error_id().decode() + ": " + self.rt.mock_last_error_msg().decode())
        return [self.to_np(plhs[i]) for i in range(nlhs)]


@pytest.fixture(scope="module")
def mock():
    return Mock()


def script_workspace():
    p = o.default_physics()
    return dict(kappa=o.kappa_of(p), tau_r=p["tau_r"], Ts=p["Ts"], zeta=o.zeta_of(p), rs=p["rs"], a=p["a"], tau_E=p["tau_E0"],
                w_dep=p["w_dep"], eta_CD=p["eta_CD"], w_marg=p["w_marg"], C=o.C_of(p)[:, None])


def test_gateways_build_and_export_mexfunction(mock):
    for n in NAMES:
        assert hasattr(mock.gateway(n), "mexFunction")


def test_bad_calls_raise_like_matlab_before_touching_the_gpu(mock):
    mock.workspace({})
    with pytest.raises(RuntimeError, match="Not enough input arguments"):
        mock.call("rho1", [[0.08, 6000.0]])                          # rho1(x) with no w_marg in the workspace
    with pytest.raises(RuntimeError, match="usage"):
        mock.call("A", [1.0])
    with pytest.raises(RuntimeError, match="usage"):
        mock.call("B", [1.0, 2.0])
    with pytest.raises(RuntimeError, match="2 x S"):
        mock.call("rho2", [[1.0, 2.0, 3.0]])
    with pytest.raises(RuntimeError, match="same length"):
        mock.call("Rho_to_PhiGammaLambda", [[1.0, 2.0], [1.0], [1.0, 2.0]])
    with pytest.raises(RuntimeError, match="N x N"):
        mock.call("ntm_qp_box", [np.eye(3), [1.0, 2.0], 0.0, 1.0])
    with pytest.raises(RuntimeError, match="usage"):
        mock.call("getWLc", [np.zeros((2, 1)), np.zeros((2, 1)), 1.0, 0.0])         # the script's broken call forms never match
    with pytest.raises(RuntimeError, match="16 x 1"):
        mock.call("ntm_mpc_batch", [np.zeros((2, 1)), np.zeros((3, 1)), 3, 20, 10, 1e-14, 0])
    E = np.zeros((0, 0))
    with pytest.raises(RuntimeError, match="finite lower and upper bounds"):
        mock.call("quadprog", [np.eye(2), [[1.0], [2.0]], [[1.0, 1.0]], [[1.0]]])      # a general row but no bounds at all
    with pytest.raises(RuntimeError, match="equality constraints"):
        mock.call("quadprog", [np.eye(2), [[1.0], [2.0]], E, E, [[1.0, 1.0]], [[1.0]], [[0.0], [0.0]], [[1.0], [1.0]]])
    with pytest.raises(RuntimeError, match="M x N"):
        mock.call("quadprog", [np.eye(2), [[1.0], [2.0]], [[1.0, 1.0, 1.0]], [[1.0]]])


@pytest.mark.gpu
def test_gateways_match_the_reference_functions(mock):
    p = o.default_physics()
    ws = script_workspace()
    mock.workspace(ws)
    x = np.array([[0.08], [2000 * math.pi]])
    X = np.array([[0.08, 0.1, 0.0], [2000 * math.pi, 3000.0, 6000.0]])
    # rho1(x), rho1(x, wmarg), batch columns
    assert mock.call("rho1", [x])[0][0, 0] == pytest.approx(o.rho1(x[:, 0], p["w_marg"]), rel=1e-14)
    assert mock.call("rho1", [x, 0.05])[0][0, 0] == pytest.approx(o.rho1(x[:, 0], 0.05), rel=1e-14)
    r2 = mock.call("rho2", [X])[0]
    assert r2.shape == (1, 3) and np.allclose(r2[0], [o.rho2(X[:, i]) for i in range(3)], rtol=1e-14)
    assert mock.call("rho3", [x])[0][0, 0] == pytest.approx(o.rho3(x[:, 0], p["w_dep"]), rel=1e-14)
    # A(r1, r2) short form and the full .m signature
    Af, Bf, C = o.model_callables(p)
    A = mock.call("A", [12.4, 1e-6])[0]
    assert A.shape == (2, 2) and np.allclose(A, Af(12.4, 1e-6), rtol=1e-14)
    A9 = mock.call("A", [12.4, 1e-6, ws["kappa"], ws["tau_r"], ws["Ts"], ws["zeta"], ws["rs"], ws["a"], ws["tau_E"]])[0]
    assert np.array_equal(A, A9)
    B = mock.call("B", [0.03])[0]
    assert B.shape == (2, 1) and np.allclose(B[:, 0], Bf(0.03), rtol=1e-14)                    # a column (D7)
    # Rho_to_PhiGammaLambda with row vectors (the script builds rows, D4)
    R = np.array([[12.0, 13.0, 14.0], [1e-6, 2e-6, 3e-6], [0.03, 0.04, 0.05]])
    Phi, Gam, Lam = mock.call("Rho_to_PhiGammaLambda", [R[0:1], R[1:2], R[2:3]], nlhs=3)
    e = o.Rho_to_PhiGammaLambda(R[0], R[1], R[2], Af, Bf, C)
    assert Phi.shape == (6, 2) and Gam.shape == (6, 3) and Lam.shape == (6, 1)
    for got, exp in ((Phi, e[0]), (Gam, e[1]), (Lam[:, 0], e[2])):
        assert np.max(np.abs(got - exp)) <= 1e-10 * np.max(np.abs(exp))
    # getWLc(xmax, xmin, umax, umin, Gamma, Phi, Lambda)
    W, L, c = mock.call("getWLc", [[[0.15], [31415.0]], [[0.06], [628.0]], 2e6, 0.0, e[1], e[0], e[2][:, None]], nlhs=3)
    eW, eL, ec = o.getWLc([0.15, 31415.0], [0.06, 628.0], [2e6], [0.0], e[1], e[0], e[2])
    assert W.shape == (22, 2) and L.shape == (22, 3) and c.shape == (22, 1)
    assert np.array_equal(W, eW) and np.array_equal(L, eL) and np.array_equal(c[:, 0], ec)
    # quadprog replacement
    G, F = o.hessian_grad(e[0], e[1], e[2], x[:, 0], [p["r1"], p["r2"]], np.eye(2))
    U, flag, it = mock.call("ntm_qp_box", [G, F[:, None], 0.0, 2e6], nlhs=3)
    assert flag[0, 0] == 1 and np.max(np.abs(U[:, 0] - o.qp_box(G, F, 0.0, 2e6)[0])) <= 1e-6 * 2e6
    # quadprog as NTM_MPC_Sim.m:97 calls it: quadprog(G, F, L, c + W*xk(:,k), [], [], [], [], [], opt), state rows kept
    E = np.zeros((0, 0))
    xs = np.array([0.08, 2000 * math.pi])
    Gs, Fs = o.hessian_grad(e[0], e[1], e[2], xs, [p["r1"], p["r2"]], np.eye(2))
    for wmin in (0.06, 0.0795):                                   # the script's bound, and one that binds
        Wq, Lq, cq = o.getWLc([0.15, 31415.0], [wmin, 628.0], [2e6], [0.0], e[1], e[0], e[2])
        bq = cq + Wq @ xs
        Uq, fval, flag = mock.call("quadprog", [Gs, Fs[:, None], Lq, bq[:, None], E, E, E, E, E, E], nlhs=3)
        lbq, ubq, Lgq, bgq, feas = o.split_rows(Lq, bq)
        Uo, _, so = o.qp_ineq(Gs, Fs, lbq, ubq, Lgq, bgq)
        assert feas and flag[0, 0] == {0: 1, 3: -2}[so]
        if so == 0:
            assert np.max(np.abs(Uq[:, 0] - Uo)) <= 1e-6 * 2e6
            assert fval[0, 0] == pytest.approx(0.5 * Uo @ Gs @ Uo + Fs @ Uo, rel=1e-9)
    bq0 = cq + Wq @ np.array([0.0, 2000 * math.pi])               # the default x0: w = 0 < min_width (D18)
    assert mock.call("quadprog", [Gs, Fs[:, None], Lq, bq0[:, None]], nlhs=3)[2][0, 0] == -2
    # batched closed loop = the script's loop
    xk, uk, cost, inner, status = mock.call("ntm_mpc_batch", [o.default_x0()[:, None], o.derive_params(p)[:, None], 3, 20, 10, 1e-14, 16], nlhs=5)
    ref = o.closed_loop(p, o.default_x0(), N=3, profile=o.LITERAL_FIXED)
    assert xk.shape == (42, 1) and uk.shape == (20, 1)
    assert np.max(np.abs(uk[:, 0] - ref["uk"])) <= 1e-6 * 2e6
    assert np.max(np.abs(xk[0::2, 0] - ref["xk"][0])) <= 1e-6 * max(np.max(np.abs(ref["xk"][0])), 1e-3)
    assert status[0, 0] == 0 and np.all(inner == 10)
    # ... with the state rows of :74 kept: the script's own x0 has w = 0 < min_width, exitflag -2 at the first QP (D18)
    xk, uk, cost, inner, status = mock.call("ntm_mpc_batch", [o.default_x0()[:, None], o.derive_params(p)[:, None], 3, 20, 10, 1e-14, 16,
                                                              2, np.array([[0.06], [200 * math.pi]]), np.array([[0.15], [10000 * math.pi]])], nlhs=5)
    assert status[0, 0] == 3 and np.all(np.isnan(uk)) and np.isnan(cost[0, 0]) and inner[0, 0] == 1
    # ... and a feasible start inside a box that binds, against the oracle
    x_in = np.array([0.1, 2000 * math.pi]); xb = (0.05, 0.101, 200 * math.pi, 10000 * math.pi)
    xk, uk, cost, inner, status = mock.call("ntm_mpc_batch", [x_in[:, None], o.derive_params(p)[:, None], 3, 6, 3, 1e-14, 16,
                                                              1, np.array([[xb[0]], [xb[2]]]), np.array([[xb[1]], [xb[3]]])], nlhs=5)
    ref = o.closed_loop(p, x_in, N=3, k_sim=6, i_sim=3, profile=o.LITERAL_FIXED, state_rows=o.STATE_ROWS_REFRESH, xbounds=xb)
    assert status[0, 0] == ref["status"]
    assert np.array_equal(np.isnan(uk[:, 0]), np.isnan(ref["uk"]))
    live = ~np.isnan(ref["uk"])
    assert np.max(np.abs(uk[live, 0] - ref["uk"][live]), initial=0.0) <= 1e-6 * 2e6
    assert mock.rt.mock_lock_count() >= 1                                                       # handle is persistent
