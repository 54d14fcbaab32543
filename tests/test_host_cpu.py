"""CPU tests of the host logic and of the C-ABI library surface (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import golden_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from juqbox_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "juqbox_b200.h")).read()
    declared = set(re.findall(r"\b(jq_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.jq_version()


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly without a device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import juqbox_b200 as jq
    cfg, _ = golden_config("rabi")
    with pytest.raises(Exception) as ei:
        jq.Working_Arrays(cfg.params, 6)
    assert "no CUDA device" in str(ei.value)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "juqbox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "jq_oracle" not in src, f


def test_named_config_shapes_match_survey_table():
    """nsteps and J are computed, not stored; a wrong integer would break golden parity."""
    from juqbox_b200 import configs
    want = {"rabi": (2, 2, 57, 3, 6), "cnot1": (6, 4, 8796, 3, 60), "cnot2": (16, 4, 4472, 4, 80),
            "cnot3": (64, 4, 31325, 3, 180), "risk_neutral": (4, 3, 7937, 5, 48)}
    for name, (n, m, nsteps, J, npar) in want.items():
        c = configs.example(name)
        p = c.params
        assert (p.Ntot, p.N, p.nsteps, p.linear_solver.max_iter, c.nCoeff) == (n, m, nsteps, J, npar), name
    tw = {"rabi": (2, 2, 57, 10, 6), "swap02": (4, 3, 7915, 4, 40), "cnot2": (12, 4, 5985, 5, 80),
          "cnot3": (96, 4, 32386, 6, 270)}
    for name, (n, m, nsteps, J, npar) in tw.items():
        c, _ = golden_config(name)
        p = c.params
        assert (p.Ntot, p.N, p.nsteps, p.linear_solver.max_iter, len(c.pcof0)) == (n, m, nsteps, J, npar), name


def test_setup_helpers():
    import juqbox_b200 as jq
    assert np.allclose(jq.wmatsetup([3], [1]), [0, 0, 0, 1.0])
    assert np.allclose(jq.wmatsetup([4], [2]), [0, 0, 0, 0, 0.1, 1.0])
    w = jq.orig_wmatsetup([2, 2], [1, 2])
    assert w.shape == (12,) and w[0] == 0 and w.max() == pytest.approx(10.0 / 6)   # nForb = 3 + 4 - 1 = 6
    U0 = jq.initial_cond([2, 2], [1, 2])
    assert U0.shape == (12, 4) and np.array_equal(np.nonzero(U0.T)[1], [0, 1, 3, 4])
    o1, o2 = jq.setup_rotmatrices([2, 2], [1, 2], [1.0, 2.0])
    assert np.allclose(o1[:4], 2 * np.pi * np.array([0, 1, 2, 0])) and np.allclose(o2[:4], [0, 0, 0, 4 * np.pi])


def test_noise_shift_matches_reference_model():
    from juqbox_b200.configs import noise_shift
    s = noise_shift(4, [0.5])[0]          # H0[j,j] += 0.01*ep*10^(j-2), j = 2..n (1-based)
    assert np.allclose(s, [0.0, 0.005, 0.05, 0.5])


def test_tikhonov():
    import juqbox_b200 as jq
    cfg, _ = golden_config("swap02")
    pc = cfg.pcof0
    assert jq.tikhonov_pen(pc, cfg.params) == pytest.approx(0.01 * pc @ pc / len(pc))
    assert np.allclose(jq.tikhonov_grad(pc, cfg.params), 2 * 0.01 * pc / len(pc))


def test_multistart_driver_logic_on_a_mock_evaluator():
    """run_optimizer_multistart's host logic (projected L-BFGS in lock step, Armijo backtracking, box handling, Tikhonov term,
    batch independence) on a stand-in for Working_Arrays.evaluate — an anisotropic quadratic bowl per member."""
    import juqbox_b200 as jq
    from juqbox_b200.optimize import IpoptProblemMirror
    cfg, _ = golden_config("rabi")
    p = cfg.params
    p.tik0 = 0.05
    n = 6
    rng = np.random.default_rng(3)
    centre = rng.uniform(-1, 1, n)
    centre[0] = 5.0                                           # outside the box: the optimum sits on the bound
    scale = np.array([1.0, 4.0, 0.5, 2.0, 8.0, 1.0])

    class FakeWA:
        calls = 0

        def evaluate(self, X, shifts=None, weights=None, evaladjoint=True, out=None):
            FakeWA.calls += 1
            X = np.atleast_2d(X)
            d = X - centre
            r = {"infid": 0.5 * (scale * d * d).sum(1, keepdims=True), "leak": np.zeros((len(X), 1))}
            if evaladjoint:
                r["grad"] = (scale * d)[:, None, :]
            return r

    lo, hi = -2.0 * np.ones(n), 2.0 * np.ones(n)
    prob = IpoptProblemMirror(p, FakeWA(), n, lo, hi, maxIter=60, lbfgsMax=6)
    starts = rng.uniform(-2, 2, (5, n))
    X, f, hist = jq.run_optimizer_multistart(prob, starts)
    # exact minimiser of 0.5*sum(scale*(x-c)^2) + tik0*|x|^2/n inside the box
    xs = np.clip(scale * centre / (scale + 2 * p.tik0 / n), lo, hi)
    assert np.all(np.diff(hist, axis=0) <= 1e-15)
    assert np.abs(X - xs).max() < 1e-5 and np.all(X <= hi + 1e-15) and np.all(X >= lo - 1e-15)
    assert np.allclose(X[:, 0], 2.0)                          # the active bound
    X1, f1, _ = jq.run_optimizer_multistart(prob, starts[2:3])
    assert np.array_equal(X1[0], X[2]) and f1[0] == f[2]      # a member's iterates do not depend on the batch


class _QuadraticWA:
    """Stand-in for Working_Arrays.evaluate: infidelity = quadratic bowl, leak = small quadratic, weights honoured."""

    def __init__(self, centre, scale, objFuncType=1):
        self.centre, self.scale, self.calls, self.objFuncType = centre, scale, 0, objFuncType

    def evaluate(self, X, shifts=None, weights=None, evaladjoint=True, out=None):
        self.calls += 1
        X = np.atleast_2d(X)
        d = X - self.centre
        ns = 1 if shifts is None else len(shifts)
        wsum = 1.0 if weights is None else float(np.sum(weights))
        shape = (len(X),) if weights is not None else (len(X), ns)
        infid = np.broadcast_to((0.5 * (self.scale * d * d).sum(1) * wsum).reshape(len(X), *([1] * (len(shape) - 1))), shape).copy()
        leak = 1e-3 * infid
        r = {"infid": infid, "leak": leak, "trace_infid": infid.copy()}
        if evaladjoint:
            g = (self.scale * d) * wsum
            r["infidgrad"] = np.broadcast_to(g.reshape(len(X), *([1] * (len(shape) - 1)), -1), shape + (X.shape[1],)).copy()
            r["leakgrad"] = 1e-3 * r["infidgrad"]
            r["grad"] = r["infidgrad"] + r["leakgrad"]
            if self.objFuncType == 1:                  # the library's convention: infidelgrad aliases totalgrad (evalobjgrad.jl:951)
                r["infidgrad"], r["leakgrad"] = r["grad"], np.zeros_like(r["grad"])
        return r


@pytest.mark.parametrize("objFuncType", [1, 3])
def test_ipopt_callback_layer_and_run_optimizer_on_a_mock_evaluator(objFuncType):
    """eval_f_par / eval_grad_f_par / eval_g_par / eval_jac_g_par (src/ipopt_interface.jl:77-179): last-evaluation cache,
    Tikhonov terms, objFuncType 3 constraint callbacks, convergence history, and the run_optimizer glue around them."""
    import juqbox_b200 as jq
    cfg, _ = golden_config("rabi")
    p = cfg.params
    p.tik0, p.objFuncType, p.leak_ubound = 0.02, objFuncType, 1.0
    n = 6
    rng = np.random.default_rng(8)
    centre, scale = rng.uniform(-0.5, 0.5, n), np.array([1.0, 3.0, 0.5, 2.0, 5.0, 1.0])
    wa = _QuadraticWA(centre, scale, objFuncType)
    x = rng.uniform(-1, 1, n)
    prob = jq.setup_ipopt_problem(p, wa, n, -np.ones(n), np.ones(n), maxIter=200, lbfgsMax=6, ipTol=1e-9)
    f = jq.eval_f_par(x, p, wa)
    quad = 0.5 * (scale * (x - centre) ** 2).sum()
    tik = p.tik0 * (x @ x) / n
    want = quad + tik if objFuncType == 3 else 1.001 * quad + tik        # objFuncType 3: leak is a constraint, not in f (:96-98)
    assert abs(f - want) < 1e-14 and wa.calls == 1
    gbuf = np.zeros(n)
    jq.eval_grad_f_par(x, gbuf, p, wa)
    assert wa.calls == 1                                               # served from the last-evaluation cache
    wantg = scale * (x - centre) * (1.0 if objFuncType == 3 else 1.001) + 2 * p.tik0 * x / n
    assert np.allclose(gbuf, wantg, atol=1e-14)
    if objFuncType == 3:
        gv = np.zeros(1)
        jq.eval_g_par(x, gv, p, wa)
        jac = np.zeros(n)
        jq.eval_jac_g_par(x, np.zeros(0, np.int32), np.zeros(0, np.int32), jac, p, wa)
        assert abs(gv[0] - 1e-3 * quad) < 1e-15 and np.allclose(jac, 1e-3 * scale * (x - centre), atol=1e-15) and wa.calls == 1
    xo = jq.run_optimizer(prob, x)
    xs = scale * centre / (scale * (1.0 if objFuncType == 3 else 1.001) + 2 * p.tik0 / n) * (1.0 if objFuncType == 3 else 1.001)
    assert np.abs(xo - xs).max() < 1e-4, (xo, xs, prob.status)
    assert len(p.objHist) >= 2 and p.objHist[-1] <= p.objHist[0]
