"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against
(1) the reference's golden files with the reference's own acceptance rule (test/evalGrad.jl:43-69) and
(2) the CPU oracle on seeded inputs, at the 1e-10 relative tolerance BASELINE.json's north_star states."""
import numpy as np
import pytest

from helpers import golden_config, ref_pass, with_tikhonov

pytestmark = pytest.mark.gpu

TOL = 1e-10   # north_star: "within 1e-10 relative in objective and gradient"


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _wa(cfg, kernel):
    import juqbox_b200 as jq
    wa = jq.Working_Arrays(cfg.params, len(cfg.pcof0) if cfg.pcof0 is not None else cfg.nCoeff)
    try:
        wa.set_kernel(kernel)
    except Exception as e:
        wa.close()
        pytest.skip(f"kernel {kernel} unavailable for {cfg.name}: {e}")
    return wa


KERNELS = [1, 2, 3, 4, 5]
GOLDEN_CASES = ["rabi", "swap02", "cnot2", "flux", "cnot2-leakieq", "cnot2-jacobi", "cnot3"]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_cuda_matches_reference_golden(case, kernel):
    cfg, g = golden_config(case)
    wa = _wa(cfg, kernel)
    res = wa.evaluate(cfg.pcof0)
    assert wa.last_kernel == kernel
    objv, grad = with_tikhonov(cfg, res)
    ok, dobj, dgrad = ref_pass(objv, grad, g["obj0"], g["grad0"])
    print(case, "kernel", kernel, "objDiff", dobj, "relGradErr", dgrad, "ms", wa.last_kernel_ms)
    wa.close()
    assert ok, (case, kernel, dobj, dgrad)


@pytest.mark.parametrize("kernel", KERNELS)
def test_traceobjgrad_dropin_tuple(kernel):
    """The reference-shaped call returns the reference-shaped tuples (src/evalobjgrad.jl:1032-1035)."""
    import juqbox_b200 as jq
    from oracle import oracle_traceobjgrad
    cfg, _ = golden_config("swap02")
    wa = _wa(cfg, kernel)
    o = oracle_traceobjgrad(cfg.params, cfg.pcof0)
    objfv, totalgrad, primary, secondary, traceInfid, infidelgrad, leakgrad = jq.traceobjgrad(cfg.pcof0, cfg.params, wa, False, True)
    assert abs(objfv - o["objf"][0, 0]) <= TOL * abs(o["objf"][0, 0])
    assert abs(primary - o["infid"][0, 0]) <= TOL and abs(secondary - o["leak"][0, 0]) <= TOL * max(o["leak"][0, 0], 1e-3)
    assert traceInfid == primary and infidelgrad is totalgrad and len(leakgrad) == 0
    assert _rel(totalgrad, o["grad"][0, 0]) < TOL
    f3 = jq.traceobjgrad(cfg.pcof0, cfg.params, wa, False, False)
    assert len(f3) == 3 and abs(f3[0] - objfv) < 1e-14
    with pytest.raises(ValueError):           # reference: error() at src/evalobjgrad.jl:604-606
        jq.traceobjgrad(np.zeros(7), cfg.params, wa)
    wa.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["rabi", "cnot1", "cnot2", "risk_neutral", "cnot3"])
def test_example_configs_batch_vs_oracle(name, kernel):
    """BASELINE configs on seeded synthetic pcof batches (no reference golden exists for these: oracle is the pin)."""
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    cfg = configs.example(name)
    if name == "cnot3" and kernel == 1:
        pytest.skip("generic kernel on the 31325-step cnot3 example: covered by the cnot3 golden")
    nb = 5 if name != "cnot3" else 2
    pc = configs.synthetic_pcof(cfg, nb)
    pc[-1] = np.random.default_rng(7).uniform(-1, 1, cfg.nCoeff) * cfg.maxpar[0]     # full-amplitude stress vector
    shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
    o = oracle_traceobjgrad(cfg.params, pc, shifts, nthreads=8)
    wa = _wa(cfg, kernel)
    r = wa.evaluate(pc, shifts)
    wa.close()
    for k in ("infid", "leak"):
        assert np.all(np.abs(r[k] - o[k]) <= TOL * np.maximum(np.abs(o[k]), 1e-6)), (k, r[k], o[k])
    for b in range(nb):
        for s in range(r["grad"].shape[1]):
            assert _rel(r["grad"][b, s], o["grad"][b, s]) < TOL, (b, s, _rel(r["grad"][b, s], o["grad"][b, s]))


@pytest.mark.parametrize("kernel", KERNELS)
def test_risk_neutral_weighted_sum(kernel):
    """eval_f_g_grad! semantics (src/ipopt_interface.jl:38-65): weighted sums over the quadrature nodes."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    cfg = configs.example("risk_neutral")
    pc = configs.synthetic_pcof(cfg, 2)
    shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes)
    o = oracle_traceobjgrad(cfg.params, pc, shifts, nthreads=8)
    wa = _wa(cfg, kernel)
    r = wa.evaluate(pc, shifts, cfg.weights)
    w = cfg.weights
    assert np.allclose(r["infid"], (o["infid"] * w).sum(1), rtol=TOL, atol=1e-14)
    assert np.allclose(r["leak"], (o["leak"] * w).sum(1), rtol=TOL, atol=1e-14)
    for b in range(2):
        assert _rel(r["grad"][b], (o["grad"][b] * w[:, None]).sum(0)) < TOL
    # Ipopt-callback layer on top (host): cache + Tikhonov
    p = cfg.params
    f = jq.eval_f_par(pc[0], p, wa, cfg.nodes, cfg.weights)
    g = np.zeros(cfg.nCoeff)
    jq.eval_grad_f_par(pc[0], g, p, wa, cfg.nodes, cfg.weights)
    assert abs(f - (r["infid"][0] + r["leak"][0] + jq.tikhonov_pen(pc[0], p))) < 1e-14
    assert _rel(g, r["grad"][0] + jq.tikhonov_grad(pc[0], p)) < 1e-14
    wa.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_ragged_batches_and_determinism(kernel):
    """Batch sizes that do not fill a CTA / a wave give the same per-trajectory results, bit for bit."""
    from juqbox_b200 import configs
    cfg = configs.example("rabi")
    wa = _wa(cfg, kernel)
    pc = configs.synthetic_pcof(cfg, 333)
    full = wa.evaluate(pc)
    again = wa.evaluate(pc)
    assert np.array_equal(full["grad"], again["grad"]) and np.array_equal(full["infid"], again["infid"])
    for nb in (1, 2, 31, 33, 150):
        part = wa.evaluate(pc[:nb])
        assert np.array_equal(part["grad"], full["grad"][:nb]), nb
        assert np.array_equal(part["leak"], full["leak"][:nb]), nb
    wa.close()


def test_full_size_properties_cnot2():
    """BASELINE size (cnot2 example, 4472 steps) through properties that need no oracle:
    (a) the gradient is the derivative of the objective (central differences on the GPU objective),
    (b) zero noise shift == no shift, (c) evaladjoint=False returns the same objective."""
    from juqbox_b200 import configs
    cfg = configs.example("cnot2")
    import juqbox_b200 as jq
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    rng = np.random.default_rng(11)
    p0 = rng.uniform(-1, 1, cfg.nCoeff) * 0.5 * cfg.maxpar[0]
    base = wa.evaluate(p0)
    h = 1e-6
    ks = [0, 13, 41, 79]
    pert = np.stack([p0 + h * np.eye(cfg.nCoeff)[k] for k in ks] + [p0 - h * np.eye(cfg.nCoeff)[k] for k in ks])
    f = wa.evaluate(pert, evaladjoint=False)["objf"][:, 0]
    for a, k in enumerate(ks):
        fd = (f[a] - f[a + len(ks)]) / (2 * h)
        assert abs(fd - base["grad"][0, 0, k]) < 2e-7 * max(1.0, abs(fd)), (k, fd, base["grad"][0, 0, k])
    z = wa.evaluate(p0, np.zeros((1, cfg.params.Ntot)))
    assert np.array_equal(z["grad"], base["grad"])
    assert wa.evaluate(p0, evaladjoint=False)["objf"][0, 0] == base["objf"][0, 0]
    wa.close()


@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("name", ["cnot2", "risk_neutral", "cnot1"])
def test_objfunctype3_second_adjoint_vs_oracle(name, kernel):
    """objFuncType = 3 (leak as inequality constraint): infidelity-only gradient from the adjoint set without forcing
    and leakgrad = totalgrad - infidelgrad (src/evalobjgrad.jl:848-855, :905-918, :940-952)."""
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    cfg = configs.example(name)
    p = cfg.params
    p.objFuncType = 3
    p.T, p.nsteps = p.T / 10.0, max(300, p.nsteps // 10)      # shorter horizon keeps the CPU oracle quick
    pc = configs.synthetic_pcof(cfg, 3) * 30.0
    shifts = configs.noise_shift(p.Ntot, cfg.nodes[:2]) if name == "risk_neutral" else None
    o = oracle_traceobjgrad(p, pc, shifts, nthreads=6)
    wa = _wa(cfg, kernel)
    r = wa.evaluate(pc, shifts)
    assert wa.last_kernel == kernel
    wa.close()
    for key in ("grad", "infidgrad", "leakgrad"):
        for b in range(3):
            for s in range(r[key].shape[1]):
                # leakgrad is a difference of two nearly equal gradients: compare it on the scale of the total gradient
                scale = np.linalg.norm(o["grad"][b, s])
                assert np.linalg.norm(r[key][b, s] - o[key][b, s]) <= TOL * scale, (key, b, s)
    assert np.allclose(r["leakgrad"], r["grad"] - r["infidgrad"], rtol=0, atol=1e-18)


def test_optimizer_glue_reduces_objective():
    """setup_ipopt_problem / run_optimizer mirror: a few L-BFGS-B iterations through the cached callbacks
    (src/ipopt_interface.jl:77-148, 212-240) reduce the risk-neutral objective."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    cfg = configs.example("risk_neutral")
    p = cfg.params
    p.T, p.nsteps = 60.0, 1600
    wa = jq.Working_Arrays(p, cfg.nCoeff)
    pc0 = configs.synthetic_pcof(cfg, 1)[0]
    minC, maxC = jq.assign_thresholds_freq([cfg.maxpar[0]] * p.Nfreq, p.Ncoupled, p.Nfreq, cfg.D1)
    f0 = jq.eval_f_par(pc0, p, wa, cfg.nodes, cfg.weights)
    prob = jq.setup_ipopt_problem(p, wa, cfg.nCoeff, minC, maxC, maxIter=8, lbfgsMax=5, nodes=cfg.nodes, weights=cfg.weights)
    pc = jq.run_optimizer(prob, pc0)
    f1 = jq.eval_f_par(pc, p, wa, cfg.nodes, cfg.weights)
    wa.close()
    assert np.all(pc >= minC - 1e-15) and np.all(pc <= maxC + 1e-15)
    assert f1 < f0 - 1e-3 and len(p.objHist) >= 2 and p.objHist[-1] <= p.objHist[0]


@pytest.mark.parametrize("kernel", KERNELS)
def test_forward_history_matches_oracle(kernel):
    """jq_eval_forward / traceobjgrad(verbose=true, evaladjoint=false): state history Ntot x N x (nsteps/saveEvery + 1)
    (src/evalobjgrad.jl:676-680, 748-752, 2797-2849), from every kernel."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    from oracle import oracle_forward_history
    cfg = configs.example("cnot2")
    p = cfg.params
    p.T, p.nsteps = 5.0, 600
    pc = configs.synthetic_pcof(cfg, 2) * 50
    wa = _wa(cfg, kernel)
    hist, infid, leak = wa.forward_history(pc, save_every=20)
    assert wa.last_kernel == kernel
    assert hist.shape == (2, 1, 31, p.N, p.Ntot)
    for b in range(2):
        want, winf, wleak = oracle_forward_history(p, pc[b], save_every=20)
        assert np.max(np.abs(hist[b, 0] - want)) < 1e-12
        assert abs(infid[b, 0] - winf) < 1e-12 and abs(leak[b, 0] - wleak) <= 1e-10 * max(wleak, 1e-12)
    # the columns stay orthonormal (unitarity of the propagation up to the scheme's error)
    last = hist[0, 0, -1]                      # [N, Ntot]
    assert np.allclose(last.conj() @ last.T, np.eye(p.N), atol=1e-6)
    objfv, uhist, fid = jq.traceobjgrad(pc[0], p, wa, True, False)
    assert uhist.shape == (p.Ntot, p.N, p.nsteps + 1) and np.array_equal(uhist[:, :, 0].real, p.Uinit)
    assert abs(fid - (1 - infid[0, 0])) < 1e-14 and abs(objfv - (infid[0, 0] + leak[0, 0])) < 1e-14
    uend = jq.eval_forward(pc[0], p, wa)
    assert np.array_equal(uend, uhist[:, :, -1])
    with pytest.raises(ValueError):
        wa.forward_history(pc, save_every=7)
    wa.close()


def test_c_abi_error_paths_and_update_target():
    """Error behaviour of the boundary (src/evalobjgrad.jl:604-606, src/bsplines.jl:178-181, :2797-2799) and
    change_target! (src/evalobjgrad.jl:1492-1505) followed by jq_update_target."""
    import ctypes as C
    import juqbox_b200 as jq
    from juqbox_b200 import _lib, configs
    from oracle import oracle_traceobjgrad
    cfg, _ = golden_config("swap02")
    p = cfg.params
    wa = jq.Working_Arrays(p, len(cfg.pcof0))
    lib = _lib.load()
    pc = np.ascontiguousarray(cfg.pcof0)
    null = C.c_void_p()
    # nsamples > 1 without shifts, zero batch, bad kernel id, pcof length not a multiple of 2*Nc*Nfreq
    assert lib.jq_traceobjgrad_batch(wa._handle, 1, pc.ctypes.data_as(C.c_void_p), len(pc), 3, null, null, 1, null, null, null, null, null, null) == -1
    assert "nsamples" in _lib.last_error()
    assert lib.jq_traceobjgrad_batch(wa._handle, 0, pc.ctypes.data_as(C.c_void_p), len(pc), 1, null, null, 1, null, null, null, null, null, null) == -1
    assert lib.jq_set_kernel(wa._handle, 9) == -1
    assert lib.jq_traceobjgrad_batch(wa._handle, 1, pc.ctypes.data_as(C.c_void_p), 38, 1, null, null, 1, null, null, null, null, null, null) == -2
    with pytest.raises(ValueError):
        wa.evaluate(np.zeros(14))               # 14 % (2*Nc) == 0 but not a multiple of 2*Nc*Nfreq
    with pytest.raises(ValueError):
        jq.Working_Arrays(p, 10)
    # a handle survives errors; all-NULL outputs are allowed
    assert lib.jq_traceobjgrad_batch(wa._handle, 1, pc.ctypes.data_as(C.c_void_p), len(pc), 1, null, null, 1, null, null, null, null, null, null) == 0
    before = wa.evaluate(pc)
    # change the target: swap two columns of the target unitary
    newU = (p.Utarget_r + 1j * p.Utarget_i)[:, [1, 0, 2]]
    jq.change_target(p, newU)
    wa.update_target()
    after = wa.evaluate(pc)
    want = oracle_traceobjgrad(p, pc)
    assert abs(after["infid"][0, 0] - want["infid"][0, 0]) < 1e-12 and abs(after["infid"][0, 0] - before["infid"][0, 0]) > 1e-3
    assert _rel(after["grad"][0, 0], want["grad"][0, 0]) < TOL
    assert after["leak"][0, 0] == before["leak"][0, 0]          # the guard-level integral does not depend on the target
    wa.close()
    wa.close()                                   # idempotent


def test_long_pcof_falls_back_when_shared_memory_does_not_fit():
    """Very long coefficient vectors (D1 = 200 knots) overflow the 4-warp CTA's gradient windows; the launch must pick a
    geometry that fits (fewer warps per CTA, else the next kernel) and still match the oracle."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    cfg = configs.example("risk_neutral")
    p = cfg.params
    p.T, p.nsteps = 40.0, 1200
    D1 = 200
    npar = 2 * p.Ncoupled * p.Nfreq * D1
    pc = (np.random.default_rng(3).random((2, npar)) - 0.5) * 0.02
    wa = jq.Working_Arrays(p, npar)
    o = oracle_traceobjgrad(p, pc)
    r = wa.evaluate(pc)
    assert wa.last_kernel == 7        # two candidates: the time-parallel evaluation (its sweeps shrink their CTAs like kernel 3)
    wa.set_kernel(5)
    r5 = wa.evaluate(pc)
    assert wa.last_kernel == 5        # the latency layout (one warp per role), which fits
    wa.set_kernel(3)
    r3 = wa.evaluate(pc)
    assert wa.last_kernel == 3        # the 4-warp fibre kernel shrinks its CTAs (fewer warps) instead of falling back
    wa.close()
    for res in (r, r5, r3):
        for b in range(2):
            assert _rel(res["grad"][b, 0], o["grad"][b, 0]) < TOL and abs(res["infid"][b, 0] - o["infid"][b, 0]) < 1e-12


def test_jacobi_solver_tolerance_exit_vs_oracle():
    """JACOBI_SOLVER (src/linear_solvers.jl:110-152) with a loose tolerance, so the sweep count is decided by the
    residual test, not by max_iter; checked on the generic kernel and, where instantiated, on the fibre kernel (whose
    groups leave the sweep loop independently)."""
    import juqbox_b200 as jq
    from juqbox_b200.params import lsolver_object, JACOBI_SOLVER
    from oracle import oracle_traceobjgrad
    cfg, _ = golden_config("swap02")
    cfg.params.linear_solver = lsolver_object(solver=JACOBI_SOLVER, max_iter=50, tol=1e-9, nrhs=cfg.params.N)
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    res = wa.evaluate(cfg.pcof0)
    assert wa.last_kernel in (1, 3)
    o = oracle_traceobjgrad(cfg.params, cfg.pcof0)
    assert abs(res["infid"][0, 0] - o["infid"][0, 0]) < 1e-10 and abs(res["leak"][0, 0] - o["leak"][0, 0]) < 1e-10
    assert _rel(res["grad"][0, 0], o["grad"][0, 0]) < 1e-8
    cfg.params.linear_solver = lsolver_object(solver=JACOBI_SOLVER, max_iter=50, tol=1e-3, nrhs=cfg.params.N)
    wa2 = jq.Working_Arrays(cfg.params, cfg.nCoeff)          # a different tolerance must change the answer
    for k in (1, 3):                                         # generic and fibre kernels decide the sweep count alike
        try:
            wa2.set_kernel(k)
        except Exception:
            continue
        rk = wa2.evaluate(np.stack([cfg.pcof0, 0.5 * cfg.pcof0, 2.0 * cfg.pcof0]))       # groups of a warp converge at different sweeps
        ok_ = oracle_traceobjgrad(cfg.params, np.stack([cfg.pcof0, 0.5 * cfg.pcof0, 2.0 * cfg.pcof0]))
        assert np.abs(rk["infid"] - ok_["infid"]).max() < 1e-10 and _rel(rk["grad"], ok_["grad"]) < 1e-8, k
    wa2.set_kernel(0)
    res2 = wa2.evaluate(cfg.pcof0)
    o2 = oracle_traceobjgrad(cfg.params, cfg.pcof0)
    assert abs(res2["infid"][0, 0] - o2["infid"][0, 0]) < 1e-10
    assert abs(res2["infid"][0, 0] - res["infid"][0, 0]) > 1e-11
    wa.close(); wa2.close()


@pytest.mark.parametrize("Ne,Ng,kw,kernel", [
    ([2], [1], {}, 3), ([3], [2], {}, 3), ([5], [3], {}, 2), ([2, 2], [0, 0], {}, 3), ([3, 3], [2, 2], {}, 3),
    ([3, 2], [2, 1], {}, 3), ([2, 2, 2], [1, 1, 1], {}, 3), ([2, 2], [2, 2], {"exchange": 0.02}, 3), ([3, 2], [1, 1], {"exchange": 0.05}, 3),
    ([2, 2, 2], [1, 1, 1], {"exchange": 0.02}, 1), ([2, 2], [2, 2], {"Nfreq": 3}, 5), ([2, 2], [2, 2], {}, 5),
    ([2, 2, 1], [2, 2, 3], {"T": 6.0}, 5),
])
def test_other_shapes_auto_kernel_vs_oracle(Ne, Ng, kw, kernel):
    """Shapes beyond the named configs (tools/kernel_coverage.py): the auto-selected kernel is the expected one and
    matches the oracle.  All-4-level qudits go to the tile layout (kernel 4; its latency variant, kernel 5, for a 3-candidate batch;
    the time-parallel evaluation, kernel 7, before either when the problem is long enough).  Exchange couplings a_0' a_q + a_0 a_q' in the drift ride on the fibre kernel's neighbour fetches; a
    coupling between two remote subsystems (three qudits) must fall back to the generic kernel, not be mis-planned."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    cfg = configs.qudit_system(Ne, Ng, **kw)
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    pc = np.random.default_rng(5).uniform(-1, 1, (3, cfg.nCoeff)) * cfg.maxpar[0] * 0.5
    o = oracle_traceobjgrad(cfg.params, pc)
    wa.set_kernel(0)
    r = wa.evaluate(pc)
    auto = wa.last_kernel
    # a 3-candidate launch of a long enough problem with a tile / fibre layout goes to the time-parallel evaluation (kernel 7);
    # JQ_SEG_NTRAJ=0 style selection (no time-parallel path) must give the listed kernel
    assert auto in (kernel, 7), auto
    results = [r]
    if auto == 7:
        import os
        os.environ["JQ_SEG_NTRAJ"] = "0"
        try:
            wb = jq.Working_Arrays(cfg.params, cfg.nCoeff)
            results.append(wb.evaluate(pc))
            assert wb.last_kernel == kernel
            wb.close()
        finally:
            del os.environ["JQ_SEG_NTRAJ"]
    for r in results:
        for b in range(3):
            assert abs(r["infid"][b, 0] - o["infid"][b, 0]) < 1e-12 and abs(r["leak"][b, 0] - o["leak"][b, 0]) < 1e-12
            assert _rel(r["grad"][b, 0], o["grad"][b, 0]) < TOL
    wa.close()


@pytest.mark.parametrize("objFuncType", [1, 3])
def test_weighted_sum_over_many_samples(objFuncType):
    """eval_f_g_grad!'s accumulation (src/ipopt_interface.jl:48-59) with hundreds of quadrature nodes goes through
    the parallel weighted-sum kernel; it must equal the weighted sum of the per-sample outputs of the same library."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    cfg = configs.example("risk_neutral")
    cfg.params.T, cfg.params.nsteps, cfg.params.objFuncType = 30.0, 800, objFuncType
    ns = 333
    eps = np.linspace(-0.06, 0.06, ns)
    w = np.random.default_rng(2).uniform(0.1, 1.0, ns)
    w /= w.sum()
    shifts = configs.noise_shift(cfg.params.Ntot, eps)
    pc = configs.synthetic_pcof(cfg, 2) * 20
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    per = wa.evaluate(pc, shifts)
    tot = wa.evaluate(pc, shifts, w)
    for key in ("infid", "leak", "trace_infid"):
        ref = (per[key] * w[None, :]).sum(axis=1)
        assert np.allclose(tot[key].ravel(), ref, rtol=1e-13, atol=1e-15), key
    for key in ("grad", "infidgrad") + (("leakgrad",) if objFuncType != 1 else ()):
        ref = (per[key] * w[None, :, None]).sum(axis=1)
        assert np.allclose(tot[key].reshape(ref.shape), ref, rtol=1e-12, atol=1e-15), key
    tot2 = wa.evaluate(pc, shifts, w)
    assert np.array_equal(tot["grad"], tot2["grad"])          # deterministic
    wa.close()


@pytest.mark.parametrize("name", ["rabi", "cnot2", "cnot3", "risk_neutral"])
def test_evalctrl_matches_oracle(name):
    """evalctrl (src/plotstatectrl.jl:246): device bcarrier2 on an arbitrary time grid vs the oracle's."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    from oracle import oracle_eval_controls
    cfg = configs.example(name)
    p = cfg.params
    pc = np.random.default_rng(4).uniform(-1, 1, cfg.nCoeff) * cfg.maxpar[0]
    t = np.concatenate([[0.0, p.T], np.random.default_rng(5).uniform(0, p.T, 1000), np.linspace(0, p.T, cfg.D1 - 1)])
    wa = jq.Working_Arrays(p, cfg.nCoeff)
    pv, qv = wa.controls(pc, t)
    po, qo = oracle_eval_controls(p, pc, t)
    scale = np.abs(po).max()
    assert np.abs(pv - po).max() <= 1e-13 * scale and np.abs(qv - qo).max() <= 1e-13 * scale
    pj, qj = jq.evalctrl(p, pc, t, p.Ncoupled, wa)           # 1-based control index, as in the reference
    assert np.array_equal(pj, pv[-1]) and np.array_equal(qj, qv[-1])
    with pytest.raises(ValueError):
        jq.evalctrl(p, pc, t, 0, wa)
    with pytest.raises(ValueError):                          # wrong coefficient count, like bcparams (src/bsplines.jl:178-181)
        wa.controls(pc[:-1], t)
    wa.close()


@pytest.mark.parametrize("risk_neutral", [False, True])
def test_multistart_lockstep_optimizer(risk_neutral):
    """run_optimizer_multistart: B optimisations in lock step, one batched GPU call per objective/gradient request.
    Every member's objective decreases monotonically, stays in the box, and its iterates do not depend on the batch."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    cfg = configs.example("risk_neutral")
    p = cfg.params
    p.T, p.nsteps = 60.0, 1600
    wa = jq.Working_Arrays(p, cfg.nCoeff)
    minC, maxC = jq.assign_thresholds_freq([cfg.maxpar[0]] * p.Nfreq, p.Ncoupled, p.Nfreq, cfg.D1)
    nodes, weights = (cfg.nodes, cfg.weights) if risk_neutral else ((0.0,), (1.0,))
    prob = jq.setup_ipopt_problem(p, wa, cfg.nCoeff, minC, maxC, maxIter=10, lbfgsMax=5, nodes=nodes, weights=weights)
    starts = configs.synthetic_pcof(cfg, 12) * 3.0
    pcs, f, hist = jq.run_optimizer_multistart(prob, starts)
    assert hist.shape[1] == 12 and hist.shape[0] >= 3
    assert np.all(np.diff(hist, axis=0) <= 1e-15)                         # Armijo: never increases
    assert np.all(hist[-1] < hist[0] - 1e-3)                              # every start makes progress
    assert np.all(pcs >= minC - 1e-15) and np.all(pcs <= maxC + 1e-15)
    # the final objective is what the reference-shaped callback reports for that vector
    f3 = jq.eval_f_par(pcs[3], p, wa, nodes, weights)
    assert abs(f3 - f[3]) <= 1e-12
    pcs1, f1, hist1 = jq.run_optimizer_multistart(prob, starts[3:4])
    assert np.allclose(pcs1[0], pcs[3], rtol=0, atol=1e-12) and abs(f1[0] - f[3]) <= 1e-12
    wa.close()


@pytest.mark.parametrize("use_sparse", [False, True])
def test_random_dense_operators_fall_back_to_generic_and_match_oracle(use_sparse):
    """Arbitrary (dense, unstructured) symmetric Hsym / antisymmetric Hanti and a full symmetric drift: nothing the
    register-resident planners recognise, so the generic kernel must take it — and agree with the oracle."""
    import juqbox_b200 as jq
    from juqbox_b200.params import objparams
    from oracle import oracle_traceobjgrad
    rng = np.random.default_rng(12)
    n, m, Nc, Nfreq, D1 = 7, 3, 2, 2, 5
    def sym(a): return (a + a.T) / 2
    H0 = sym(rng.standard_normal((n, n))) * 0.3
    Hs = [sym(rng.standard_normal((n, n))) for _ in range(Nc)]
    Ha = [(lambda a: (a - a.T) / 2)(rng.standard_normal((n, n))) for _ in range(Nc)]
    if use_sparse:                                            # knock out entries so the CSC path has real structure
        for M in Hs + Ha:
            mask = rng.random((n, n)) < 0.5
            mask = mask & mask.T
            M[mask] = 0.0
    U0 = np.eye(n, m)
    Vt = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))[0][:, :m]
    p = objparams([m], [n - m], 3.0, 300, Uinit=U0, Utarget=Vt, Cfreq=rng.standard_normal((Nc, Nfreq)), Rfreq=[1.0, 2.0],
                  Hconst=H0, Hsym_ops=Hs, Hanti_ops=Ha, use_sparse=use_sparse)
    npar = 2 * Nc * Nfreq * D1
    pc = rng.uniform(-0.2, 0.2, (3, npar))
    wa = jq.Working_Arrays(p, npar)
    o = oracle_traceobjgrad(p, pc)
    for want in (0, 1, 6):                # automatic (n < 8: generic), generic, dense tensor-core kernel
        wa.set_kernel(want)
        r = wa.evaluate(pc)
        assert wa.last_kernel == (want or 1)
        for b in range(3):
            assert abs(r["infid"][b, 0] - o["infid"][b, 0]) < 1e-12 and abs(r["leak"][b, 0] - o["leak"][b, 0]) < 1e-12
            assert _rel(r["grad"][b, 0], o["grad"][b, 0]) < TOL
    wa.close()


@pytest.mark.parametrize("name", ["cnot2", "rabi"])
def test_reference_optimised_pulse_on_gpu(name):
    """The reference's own optimised pulse (examples/drives) on the example config: GPU = oracle, and the gate is good."""
    import json
    import os
    import juqbox_b200 as jq
    from helpers import GOLDEN_DIR
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    pc = np.array(json.load(open(os.path.join(GOLDEN_DIR, "drives.json")))[name]["pcof"])
    cfg = configs.example(name)
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    r = wa.evaluate(pc)
    o = oracle_traceobjgrad(cfg.params, pc)
    assert abs(r["infid"][0, 0] - o["infid"][0, 0]) < 1e-12 and abs(r["leak"][0, 0] - o["leak"][0, 0]) < 1e-12
    assert _rel(r["grad"][0, 0], o["grad"][0, 0]) < TOL
    assert abs(r["infid"][0, 0]) < 1e-3
    wa.close()
