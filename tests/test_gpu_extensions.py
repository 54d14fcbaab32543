"""GPU parity (through the C ABI) of the SURVEY 8f rank-3 extensions and the fused callback entry against the CPU oracle.
The oracle of these branches is pinned by finite differences and reductions to the golden-pinned core
(tests/test_oracle_extensions.py); the reference holds no golden for them ("parity unpinned by the reference")."""
import numpy as np
import pytest

from helpers import golden_config, ref_pass, with_tikhonov

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _small_swap():
    cfg, _ = golden_config("swap02")
    cfg.params.T, cfg.params.nsteps = 30.0, 1600
    return cfg, np.asarray(cfg.pcof0) * 3.0


def _check(params, pcs, kernels, shifts=None):
    import juqbox_b200 as jq
    from oracle import oracle_traceobjgrad
    o = oracle_traceobjgrad(params, pcs, shifts, nthreads=4)
    wa = jq.Working_Arrays(params, pcs.shape[1])
    seen = []
    for k in kernels:
        try:
            wa.set_kernel(k)
        except Exception:
            continue
        r = wa.evaluate(pcs, shifts)
        seen.append(wa.last_kernel)
        for key in ("infid", "leak", "trace_infid"):
            assert np.all(np.abs(r[key] - o[key]) <= TOL * np.maximum(np.abs(o[key]), 1e-6)), (k, key, r[key], o[key])
        for gk in ("grad", "infidgrad", "leakgrad") if params.objFuncType != 1 else ("grad",):
            assert r[gk].shape == o[gk].shape
            for b in range(pcs.shape[0]):
                assert _rel(r[gk][b], o[gk][b]) < TOL or np.linalg.norm(o[gk][b]) < 1e-13, (k, gk, b, _rel(r[gk][b], o[gk][b]))
    wa.close()
    return seen


@pytest.mark.parametrize("objFuncType", [1, 3])
@pytest.mark.parametrize("pfid", [1, 3, 4])
def test_pfidtype_vs_oracle(pfid, objFuncType):
    cfg, pc = _small_swap()
    p = cfg.params
    p.pFidType, p.globalPhase, p.objFuncType = pfid, 0.37, objFuncType
    pcs = np.stack([pc, 0.5 * pc, -1.5 * pc])
    if pfid == 3:
        pcs = np.concatenate([pcs, [[0.37], [-1.1], [2.5]]], axis=1)       # one global phase per candidate
    seen = _check(p, pcs, (1, 2, 3))
    assert 1 in seen and 3 in seen


def test_pfidtype_on_tile_kernel_vs_oracle():
    from juqbox_b200 import configs
    cfg = configs.qudit_system([2, 2], [2, 2], T=8.0)
    cfg.params.pFidType, cfg.params.globalPhase = 4, -0.6
    pcs = np.random.default_rng(2).uniform(-1, 1, (2, cfg.nCoeff)) * cfg.maxpar[0] * 0.5
    assert 4 in _check(cfg.params, pcs, (4, 3))


def test_dense_forbidden_weights_vs_oracle():
    from juqbox_b200.params import objparams
    cfg, pc = _small_swap()
    p = cfg.params
    n = p.Ntot
    rng = np.random.default_rng(3)
    F = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    F /= np.linalg.norm(F, axis=0)
    for objFuncType in (1, 2):
        pd = objparams(p.Ne, p.Ng, p.T, p.nsteps, Uinit=p.Uinit, Utarget=p.Utarget_r + 1j * p.Utarget_i, Cfreq=p.Cfreq, Rfreq=p.Rfreq,
                       Hconst=p.Hconst, Hsym_ops=p.Hsym_ops, Hanti_ops=p.Hanti_ops, linear_solver=p.linear_solver, objFuncType=objFuncType,
                       use_custom_forbidden=True, forb_states=F, forb_weights=[0.7, 0.2])
        seen = _check(pd, np.stack([pc, 0.3 * pc]), (0,), shifts=np.array([[0.0, 0.001, 0.01, 0.1], [0.0, -0.002, -0.02, -0.2]]))
        assert seen == [1]                    # dense weights: generic kernel


def test_uncoupled_controls_vs_oracle():
    from juqbox_b200 import configs
    from juqbox_b200.params import objparams
    cfg = configs.example("rabi_lab", T=20.0, Pmin=60)
    rng = np.random.default_rng(1)
    pcs = cfg.pcof0[None, :] * 5.0 + 0.3 * cfg.maxpar[0] * rng.standard_normal((3, cfg.nCoeff))
    assert _check(cfg.params, pcs, (0,)) == [1]
    # antisymmetric uncoupled operator next to a coupled control pair, sparse storage, objFuncType 2
    base, pc = _small_swap()
    p = base.params
    a = np.diag(np.sqrt(np.arange(1, p.Ntot)), 1)
    om = np.vstack([p.Cfreq[:1], p.Cfreq[:1] * 0.5])
    pm = objparams(p.Ne, p.Ng, 10.0, 4000, Uinit=p.Uinit, Utarget=p.Utarget_r + 1j * p.Utarget_i, Cfreq=om, Rfreq=[0.0, 0.7],
                   Hconst=p.Hconst, Hsym_ops=[a + a.T], Hanti_ops=[a - a.T], Hunc_ops=[0.3 * (a - a.T)], use_sparse=True, objFuncType=2)
    # NB: the reference asserts Ncoupled == 0 || Nunc == 0 (src/evalobjgrad.jl:176); the kernels take both, Rfreq indexed per
    # uncoupled control as KS! does (Rfreq[q], :2380)
    pm.Rfreq = [0.7, 0.0]
    npar = 2 * 2 * pm.Nfreq * 6
    pcs = rng.uniform(-1, 1, (2, npar)) * 0.02
    assert _check(pm, pcs, (0,)) == [1]


@pytest.mark.parametrize("case", ["rabi", "swap02", "flux", "cnot2", "cnot2-leakieq"])
def test_fused_callback_entry_matches_golden_and_caches(case):
    """jq_eval_f_grad = eval_f_par + eval_grad_f_par (+ eval_g_par / eval_jac_g_par) in one ccall, at the layer the reference's
    goldens are defined on (test/evalGrad.jl:14-26)."""
    import juqbox_b200 as jq
    cfg, g = golden_config(case)
    p = cfg.params
    wa = jq.Working_Arrays(p, len(cfg.pcof0))
    r = wa.eval_f_grad(cfg.pcof0, tik0=p.tik0)
    assert r["evaluated"]
    if p.objFuncType == 1:
        objv, grad = r["f"], r["grad_f"]
    else:
        objv, grad = np.array([r["f"], r["leak"]]), np.concatenate([r["grad_f"], r["leakgrad"]])
    ok, dobj, dgrad = ref_pass(objv, grad, g["obj0"], g["grad0"])
    assert ok, (case, dobj, dgrad)
    r2 = wa.eval_f_grad(cfg.pcof0, tik0=p.tik0)
    assert not r2["evaluated"] and abs(r2["f"] - r["f"]) <= 1e-15 * abs(r["f"]) and np.allclose(r2["grad_f"], r["grad_f"], rtol=1e-14, atol=1e-18)
    r3 = wa.eval_f_grad(cfg.pcof0 * (1 + 1e-9), tik0=p.tik0)
    assert r3["evaluated"]
    # the callback mirrors share that single evaluation
    f = jq.eval_f_par(cfg.pcof0, p, wa)
    gb = np.zeros(len(cfg.pcof0))
    jq.eval_grad_f_par(cfg.pcof0, gb, p, wa)
    assert abs(f - r["f"]) <= 1e-15 * abs(r["f"]) and np.allclose(gb, r["grad_f"], rtol=1e-14, atol=1e-18)
    wa.update_target()
    assert wa.eval_f_grad(cfg.pcof0, tik0=p.tik0)["evaluated"]           # a new target invalidates the cache
    # prior coefficients (usingPriorCoeffs, src/evalobjgrad.jl:2300-2303)
    prior = 0.5 * np.asarray(cfg.pcof0)
    rp = wa.eval_f_grad(cfg.pcof0 * 1.0001, tik0=0.3, prior=prior)
    base = wa.evaluate(cfg.pcof0 * 1.0001)
    d = cfg.pcof0 * 1.0001 - prior
    fbase = base["infid"][0, 0] + (base["leak"][0, 0] if p.objFuncType == 1 else 0.0)
    assert abs(rp["f"] - (fbase + 0.3 * d @ d / len(d))) < 1e-14
    assert _rel(rp["grad_f"], base["infidgrad"][0, 0] + 2 * 0.3 * d / len(d)) < 1e-14
    wa.close()


def test_fused_entry_risk_neutral_and_empty_shard():
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    cfg = configs.example("risk_neutral")
    cfg.params.T, cfg.params.nsteps = 30.0, 800
    pc = configs.synthetic_pcof(cfg, 1)[0] * 20
    sh = configs.noise_shift(cfg.params.Ntot, cfg.nodes)
    wa = jq.Working_Arrays(cfg.params, cfg.nCoeff)
    r = wa.eval_f_grad(pc, sh, cfg.weights, tik0=0.01)
    b = wa.evaluate(pc, sh, cfg.weights)
    assert abs(r["f"] - (b["infid"][0] + b["leak"][0] + 0.01 * pc @ pc / len(pc))) < 1e-14
    assert _rel(r["grad_f"], b["grad"][0] + 2 * 0.01 * pc / len(pc)) < 1e-14
    # an empty sample shard (more ranks than nodes) contributes zeros instead of failing / hanging its peers
    e = wa.evaluate(pc, np.zeros((0, cfg.params.Ntot)), np.zeros(0))
    assert e["infid"][0] == 0.0 and e["leak"][0] == 0.0 and not np.any(e["grad"])
    wa.close()


def _dense_problem(n, m, Nc=2, Nfreq=2, nsteps=240, seed=12, pfid=2):
    from juqbox_b200.params import objparams
    rng = np.random.default_rng(seed)
    sym = lambda a: (a + a.T) / 2
    H0 = sym(rng.standard_normal((n, n))) * 0.3
    Hs = [sym(rng.standard_normal((n, n))) / np.sqrt(n) for _ in range(Nc)]
    Ha = [(lambda a: (a - a.T) / 2)(rng.standard_normal((n, n))) / np.sqrt(n) for _ in range(Nc)]
    Vt = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))[0][:, :m]
    p = objparams([m], [n - m], 3.0, nsteps, Uinit=np.eye(n, m), Utarget=Vt, Cfreq=rng.standard_normal((Nc, Nfreq)), Rfreq=[1.0, 2.0],
                  Hconst=H0, Hsym_ops=Hs, Hanti_ops=Ha)
    p.pFidType, p.globalPhase = pfid, 0.4
    return p, 2 * Nc * Nfreq * 5


@pytest.mark.parametrize("n,m,nsamples,pfid", [(12, 4, 1, 2), (24, 6, 11, 2), (9, 3, 5, 3), (33, 5, 3, 4)])
def test_dense_tensor_core_kernel_vs_oracle_and_generic(n, m, nsamples, pfid):
    """Unstructured dense operators: the FP64-MMA kernel (kernel id 6) batches the noise samples of a candidate as columns of one
    contraction (ragged last tile: 11 = 8 + 3 samples); it must agree with the oracle and with the generic kernel, and the automatic
    selection must pick it for n >= 8."""
    import juqbox_b200 as jq
    from oracle import oracle_traceobjgrad
    p, npar = _dense_problem(n, m, pfid=pfid)
    rng = np.random.default_rng(4)
    pcs = rng.uniform(-0.2, 0.2, (3, npar))
    if pfid == 3:
        pcs = np.concatenate([pcs, [[0.3], [-0.9], [1.7]]], axis=1)
    shifts = None if nsamples == 1 else rng.uniform(-0.05, 0.05, (nsamples, n))
    o = oracle_traceobjgrad(p, pcs, shifts, nthreads=8)
    wa = jq.Working_Arrays(p, pcs.shape[1])
    for want in (0, 6, 1):
        wa.set_kernel(want)
        r = wa.evaluate(pcs, shifts)
        assert wa.last_kernel == (want or 6)
        for key in ("infid", "leak", "trace_infid"):
            assert np.all(np.abs(r[key] - o[key]) <= TOL * np.maximum(np.abs(o[key]), 1e-6)), (want, key)
        for b in range(3):
            for s in range(nsamples):
                assert _rel(r["grad"][b, s], o["grad"][b, s]) < TOL, (want, b, s, _rel(r["grad"][b, s], o["grad"][b, s]))
    if nsamples > 1:                       # weighted sums through the same kernel
        w = rng.uniform(0.1, 1.0, nsamples)
        wa.set_kernel(0)
        rw = wa.evaluate(pcs, shifts, w)
        assert _rel(rw["grad"], (o["grad"] * w[None, :, None]).sum(1)) < TOL
    wa.close()


def test_handle_guards_against_silent_misuse():
    """ADVICE (round 1): params edited after the handle was created must not be silently ignored; device-entry arguments are
    validated before raw pointers reach the library; host-path and device-path calls may be mixed (shared scratch, two streams)."""
    import torch
    import juqbox_b200 as jq
    from juqbox_b200 import configs
    cfg = configs.example("risk_neutral")
    cfg.params.T, cfg.params.nsteps = 30.0, 800
    p = cfg.params
    wa = jq.Working_Arrays(p, cfg.nCoeff)
    pc = configs.synthetic_pcof(cfg, 4) * 20
    sh = configs.noise_shift(p.Ntot, cfg.nodes)
    host = wa.evaluate(pc, sh)
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(dev)
    pcd, shd = torch.from_numpy(pc).to(dev), torch.from_numpy(sh).to(dev)
    with torch.cuda.stream(side):
        d = wa.evaluate_device(pcd, shd, None, True, stream=side)       # caller's stream right after the handle's own stream
    again = wa.evaluate(pc * 1.5, sh)                                     # and back, while `side` may still be running
    side.synchronize()
    assert np.array_equal(d["grad"].cpu().numpy(), host["grad"])
    assert not np.array_equal(again["grad"], host["grad"])
    with pytest.raises(ValueError):
        wa.evaluate_device(pcd.float(), shd)                              # wrong dtype
    with pytest.raises(ValueError):
        wa.evaluate_device(pcd, shd[:, :-1].contiguous())                 # wrong shape
    p.nsteps = 900                                                        # the reference would re-read this on the next call
    with pytest.raises(RuntimeError):
        wa.evaluate(pc, sh)
    wa.close()
