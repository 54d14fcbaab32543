"""GPU parity of the time-parallel evaluation (kernel id 7, csrc/jq_seg.cu) through the C ABI: against the reference's goldens with the
reference's own acceptance rule, against the CPU oracle at the north-star tolerance (1e-10 relative), and against the plain kernels at
1e-12 -- the segments are joined through the discrete propagators, so the result is the plain one up to rounding, including the
reference's backward recomputation of the states with the times of its own backward recurrence."""
import numpy as np
import pytest

from helpers import golden_config, ref_pass, with_tikhonov

pytestmark = pytest.mark.gpu
TOL = 1e-10        # north_star: "within 1e-10 relative in objective and gradient"
TIGHT = 1e-12      # against the plain kernels: same arithmetic per step, joins in different order


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _wa7(params, ncoeff, nseg=0):
    import juqbox_b200 as jq
    wa = jq.Working_Arrays(params, ncoeff)
    try:
        wa.set_kernel(7)
    except Exception as e:
        wa.close()
        pytest.skip(f"no time-parallel evaluation for this problem: {e}")
    wa.set_time_segments(nseg)
    return wa


@pytest.mark.parametrize("nseg", [0, 3])
@pytest.mark.parametrize("case", ["swap02", "cnot2", "flux", "cnot3", "rabi", "cnot2-leakieq"])
def test_time_parallel_matches_reference_golden(case, nseg):
    cfg, g = golden_config(case)
    wa = _wa7(cfg.params, len(cfg.pcof0), nseg)
    res = wa.evaluate(cfg.pcof0)
    assert wa.last_kernel == 7 and (nseg == 0 or int(wa.query(7)) == nseg)
    objv, grad = with_tikhonov(cfg, res)
    ok, dobj, dgrad = ref_pass(objv, grad, g["obj0"], g["grad0"])
    print(case, "segments", int(wa.query(7)), "objDiff", dobj, "relGradErr", dgrad, "ms", wa.last_kernel_ms)
    wa.close()
    assert ok, (case, dobj, dgrad)


@pytest.mark.parametrize("name", ["cnot1", "cnot2", "risk_neutral", "cnot3"])
def test_example_configs_vs_oracle_and_plain_kernel(name):
    """BASELINE configs, seeded synthetic pcof incl. a full-amplitude stress vector: oracle at 1e-10, plain kernel at 1e-12."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    cfg = configs.example(name)
    nb = 3 if name != "cnot3" else 2
    pc = configs.synthetic_pcof(cfg, nb)
    pc[-1] = np.random.default_rng(7).uniform(-1, 1, cfg.nCoeff) * cfg.maxpar[0]
    shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes) if name == "risk_neutral" else None
    o = oracle_traceobjgrad(cfg.params, pc, shifts, nthreads=8)
    wa = _wa7(cfg.params, cfg.nCoeff)
    r = wa.evaluate(pc, shifts)
    assert wa.last_kernel == 7 and wa.query(7) > 1
    f = wa.evaluate(pc, shifts, None, False)                  # objective only: no backward sweeps
    wa.set_kernel(3)
    q = wa.evaluate(pc, shifts)
    wa.close()
    for k in ("infid", "leak", "trace_infid"):
        assert np.all(np.abs(r[k] - o[k]) <= TOL * np.maximum(np.abs(o[k]), 1e-6)), (k, r[k], o[k])
        assert np.all(np.abs(r[k] - q[k]) <= TIGHT * np.maximum(np.abs(q[k]), 1e-6)), (k, r[k], q[k])
        assert np.all(np.abs(f[k] - r[k]) <= TIGHT * np.maximum(np.abs(r[k]), 1e-6)), k       # its own number of segments
    for b in range(nb):
        for s in range(r["grad"].shape[1]):
            assert _rel(r["grad"][b, s], o["grad"][b, s]) < TOL, (b, s, _rel(r["grad"][b, s], o["grad"][b, s]))
            assert _rel(r["grad"][b, s], q["grad"][b, s]) < TIGHT, (b, s, _rel(r["grad"][b, s], q["grad"][b, s]))


@pytest.mark.parametrize("nseg", [1, 2, 7, 33, 100])
def test_any_number_of_segments(nseg):
    """Ragged segment lengths, one segment (the joins are the identity), more segments than a wave: same result."""
    cfg, _ = golden_config("swap02")
    cfg.params.T, cfg.params.nsteps = 30.0, 1603
    pc = np.stack([np.asarray(cfg.pcof0) * 3.0, -np.asarray(cfg.pcof0)])
    wa = _wa7(cfg.params, pc.shape[1], nseg)
    r = wa.evaluate(pc)
    assert int(wa.query(7)) == nseg
    again = wa.evaluate(pc)                                   # bit-reproducible: fixed summation orders, no atomics
    assert all(np.array_equal(r[k], again[k]) for k in ("infid", "leak", "grad"))
    wa.set_kernel(3)
    q = wa.evaluate(pc)
    wa.close()
    assert np.allclose(r["infid"], q["infid"], rtol=TIGHT, atol=1e-15) and np.allclose(r["leak"], q["leak"], rtol=TIGHT, atol=1e-18)
    assert _rel(r["grad"], q["grad"]) < TIGHT


@pytest.mark.parametrize("name,T,nsteps,scale", [("rabi", 10.0, 12, 20.0), ("risk_neutral", 30.0, 160, 40.0), ("cnot2", 6.0, 90, 60.0)])
def test_coarse_time_steps_refine_the_backward_boundary_states(name, T, nsteps, scale):
    """With a few large steps the truncated Neumann solves make the backward recomputation of the states visibly irreversible (the
    reference recomputes them like that) and break the symplectic identity the first-order join of the boundary states rests on:
    the refinement passes must bring the time-parallel result back to the plain kernels' (1e-7 without them on the first case)."""
    from juqbox_b200 import configs
    cfg = configs.example(name)
    cfg.params.T, cfg.params.nsteps = T, nsteps
    pc = configs.synthetic_pcof(cfg, 2) * scale
    wa = _wa7(cfg.params, cfg.nCoeff, 3)
    r = wa.evaluate(pc)
    wa.set_kernel(3)
    q = wa.evaluate(pc)
    wa.close()
    assert np.allclose(r["infid"], q["infid"], rtol=TIGHT, atol=1e-15) and np.allclose(r["leak"], q["leak"], rtol=TIGHT, atol=1e-18)
    assert _rel(r["grad"], q["grad"]) < TIGHT, _rel(r["grad"], q["grad"])


@pytest.mark.parametrize("pfid", [1, 3, 4])
def test_pfidtype_and_global_phase(pfid):
    from oracle import oracle_traceobjgrad
    cfg, _ = golden_config("swap02")
    p = cfg.params
    p.T, p.nsteps, p.pFidType, p.globalPhase = 30.0, 1600, pfid, 0.37
    pcs = np.stack([np.asarray(cfg.pcof0) * 3.0, np.asarray(cfg.pcof0) * -1.5])
    if pfid == 3:
        pcs = np.concatenate([pcs, [[0.37], [-1.1]]], axis=1)       # one global phase per candidate, one more gradient entry
    o = oracle_traceobjgrad(p, pcs, None, nthreads=2)
    wa = _wa7(p, pcs.shape[1])
    r = wa.evaluate(pcs)
    wa.close()
    assert r["grad"].shape == o["grad"].shape
    assert np.allclose(r["infid"], o["infid"], rtol=TOL, atol=1e-14)
    for b in range(2):
        assert _rel(r["grad"][b], o["grad"][b]) < TOL


def test_risk_neutral_weighted_sums_fused_entry_and_automatic_choice():
    """eval_f_g_grad! semantics through the time-parallel path; one pcof per call takes it automatically, a large batch does not."""
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    cfg = configs.example("risk_neutral")
    pc = configs.synthetic_pcof(cfg, 1)
    shifts = configs.noise_shift(cfg.params.Ntot, cfg.nodes)
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    a = wa.evaluate(pc, shifts, cfg.weights)
    assert wa.last_kernel == 7, "a single risk-neutral evaluation should take the time-parallel path"
    fa = wa.eval_f_grad(pc[0], shifts, cfg.weights, tik0=0.01)
    assert wa.last_kernel == 7 and fa["evaluated"]
    wa.set_kernel(3)
    b = wa.evaluate(pc, shifts, cfg.weights)
    wa.set_kernel(0)
    big = wa.evaluate(configs.synthetic_pcof(cfg, 2048), shifts, cfg.weights)
    assert wa.last_kernel != 7 and big["infid"].shape == (2048,)
    wa.close()
    assert np.allclose(a["infid"], b["infid"], rtol=TIGHT) and np.allclose(a["leak"], b["leak"], rtol=TIGHT, atol=1e-18)
    assert _rel(a["grad"], b["grad"]) < TIGHT
    tik = 0.01 * float(pc[0] @ pc[0]) / len(pc[0])
    assert abs(fa["f"] - (b["infid"][0] + b["leak"][0] + tik)) < 1e-12
    assert _rel(fa["grad_f"], b["grad"][0] + 2 * 0.01 * pc[0] / len(pc[0])) < TIGHT


@pytest.mark.parametrize("objFuncType", [2, 3])
def test_second_adjoint_set_vs_oracle(objFuncType):
    """objFuncType 2/3 (leakage as a constraint): the infidelity-only gradient comes from a second, unforced adjoint set
    (src/evalobjgrad.jl:848-855) whose boundary values are joined with the adjoint propagators alone."""
    from juqbox_b200 import configs
    from oracle import oracle_traceobjgrad
    cfg = configs.example("cnot2")
    cfg.params.objFuncType = objFuncType
    pc = configs.synthetic_pcof(cfg, 2) * np.array([[1.0], [40.0]])
    o = oracle_traceobjgrad(cfg.params, pc, None, nthreads=2)
    wa = _wa7(cfg.params, cfg.nCoeff)
    r = wa.evaluate(pc)
    assert wa.last_kernel == 7
    wa.set_kernel(3)
    q = wa.evaluate(pc)
    wa.close()
    for k in ("infid", "leak"):
        assert np.all(np.abs(r[k] - o[k]) <= TOL * np.maximum(np.abs(o[k]), 1e-6)), k
    for gk in ("grad", "infidgrad"):
        for b in range(2):
            assert _rel(r[gk][b], o[gk][b]) < TOL, (gk, b, _rel(r[gk][b], o[gk][b]))
            assert _rel(r[gk][b], q[gk][b]) < 1e-11, (gk, b, _rel(r[gk][b], q[gk][b]))
    # leakgrad = totalgrad - infidelgrad (src/evalobjgrad.jl:951), a difference of nearly equal vectors when the leakage is tiny
    # (first candidate: |leakgrad| = 1e-6 |grad|): accurate to the rounding of the two gradients, not of itself
    for b in range(2):
        err = np.linalg.norm(r["leakgrad"][b] - o["leakgrad"][b])
        assert err < TOL * np.linalg.norm(o["leakgrad"][b]) or err < 1e-12 * np.linalg.norm(o["grad"][b]), (b, err)


def test_unsupported_problems_say_so_or_fall_back():
    """The Jacobi solver has no time-parallel path (set_kernel refuses, automatic mode uses the other kernels); state histories too."""
    import juqbox_b200 as jq
    cfg, _ = golden_config("cnot2-jacobi")
    wa = jq.Working_Arrays(cfg.params, len(cfg.pcof0))
    with pytest.raises(Exception):
        wa.set_kernel(7)
    wa.evaluate(cfg.pcof0)
    assert wa.last_kernel != 7
    wa.close()
    cfg, _ = golden_config("cnot2")
    wa = jq.Working_Arrays(cfg.params, len(cfg.pcof0))
    wa.evaluate(cfg.pcof0)
    k_eval = wa.last_kernel
    wa.forward_history(np.atleast_2d(cfg.pcof0), save_every=cfg.params.nsteps // 5 if cfg.params.nsteps % 5 == 0 else cfg.params.nsteps)
    assert wa.last_kernel != 7
    wa.close()
    assert k_eval in (5, 7)
    wa = jq.Working_Arrays(cfg.params, len(cfg.pcof0))
    with pytest.raises(Exception):           # sharing one evaluation over several GPUs needs a communicator first
        wa.comm_set_cooperative(True)
    wa.comm_set_cooperative(False)
    wa.close()
