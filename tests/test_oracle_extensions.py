"""Oracle checks of the SURVEY 8f rank-3 extensions: pFidType 1/3/4 + global phase, dense forbidden-state weights
(wmat_real / wmat_imag), uncoupled (lab-frame) controls.  PARITY UNPINNED BY THE REFERENCE (no reference test or golden
exercises these branches): the pins are (1) reductions to the golden-pinned core where the extension degenerates to it,
(2) central finite differences of the restated objective, (3) for the uncoupled branch, the reference's own shipped rabi
pulse reproducing its gate on the lab-frame model of examples/rabi-lab.jl."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR, golden_config
from juqbox_b200 import configs
from oracle import oracle_traceobjgrad


def _fd_check(params, pc, ndir=3, eps=1e-6, tol=2e-6, key="objf", gkey="grad", seed=0):
    rng = np.random.default_rng(seed)
    g = oracle_traceobjgrad(params, pc)[gkey][0, 0]
    for _ in range(ndir):
        d = rng.standard_normal(len(pc))
        d /= np.linalg.norm(d)
        fp = oracle_traceobjgrad(params, pc + eps * d, evaladjoint=False)[key][0, 0]
        fm = oracle_traceobjgrad(params, pc - eps * d, evaladjoint=False)[key][0, 0]
        fd = (fp - fm) / (2 * eps)
        assert abs(fd - g @ d) <= tol * max(abs(fd), np.linalg.norm(g) * 1e-2), (fd, g @ d)
    return g


def _small_swap():
    cfg, _ = golden_config("swap02")
    cfg.params.T, cfg.params.nsteps = 30.0, 1600        # short horizon: FD needs many evaluations
    pc = np.asarray(cfg.pcof0) * 3.0
    return cfg, pc


@pytest.mark.parametrize("pfid", [1, 3, 4])
def test_pfidtype_finite_differences(pfid):
    cfg, pc = _small_swap()
    p = cfg.params
    p.pFidType, p.globalPhase = pfid, 0.37
    if pfid == 3:
        pc = np.concatenate([pc, [0.37]])              # last entry = global phase (src/evalobjgrad.jl:591-596)
    g = _fd_check(p, pc)
    assert len(g) == len(pc)
    o = oracle_traceobjgrad(p, pc)
    # infidelity relations: with s = tr(V'Vtg)/N, type 2 = 1 - |s|^2, type 1 = 1 + |s|^2 - 2 Re(s e^{-i phi}), type 3/4 = 1 - Re(s e^{...})
    p2 = cfg.params
    p2.pFidType = 2
    o2 = oracle_traceobjgrad(p2, pc[:len(cfg.pcof0)])
    assert abs(o["leak"][0, 0] - o2["leak"][0, 0]) < 1e-15          # the guard-level term does not depend on pFidType
    if pfid in (3, 4):
        assert o["infid"][0, 0] >= 0.5 * o2["infid"][0, 0] - 1e-12    # 1 - Re(z) >= (1 - |z|^2) / 2 for |z| <= 1


def test_pfidtype3_phase_is_the_last_entry():
    """pFidType 3 with phase phi in pcof equals pFidType 4 with globalPhase = phi (objective and spline part of the gradient)."""
    cfg, pc = _small_swap()
    p = cfg.params
    p.pFidType, p.globalPhase = 4, -0.81
    o4 = oracle_traceobjgrad(p, pc)
    p.pFidType = 3
    o3 = oracle_traceobjgrad(p, np.concatenate([pc, [-0.81]]))
    assert o3["objf"][0, 0] == o4["objf"][0, 0]
    assert np.array_equal(o3["grad"][0, 0, :-1], o4["grad"][0, 0])


def test_dense_weights_reduce_to_diagonal_and_fd():
    cfg, pc = _small_swap()
    p = cfg.params
    base = oracle_traceobjgrad(p, pc)
    n = p.Ntot
    wd = np.asarray(p.wmat_real).copy()
    # (1) unit-vector forbidden states with the diagonal weights reproduce the Diagonal path (different summation order only)
    from juqbox_b200.params import objparams
    kw = dict(Uinit=p.Uinit, Utarget=p.Utarget_r + 1j * p.Utarget_i, Cfreq=p.Cfreq, Rfreq=p.Rfreq, Hconst=p.Hconst,
              Hsym_ops=p.Hsym_ops, Hanti_ops=p.Hanti_ops, linear_solver=p.linear_solver)
    pd = objparams(p.Ne, p.Ng, p.T, p.nsteps, use_custom_forbidden=True, forb_states=np.eye(n, dtype=complex), forb_weights=wd, **kw)
    o = oracle_traceobjgrad(pd, pc)
    assert abs(o["leak"][0, 0] - base["leak"][0, 0]) <= 1e-13 * max(base["leak"][0, 0], 1e-3)
    assert np.linalg.norm(o["grad"] - base["grad"]) <= 1e-12 * np.linalg.norm(base["grad"])
    # (2) complex superposition states: wmat_imag != 0; finite differences of the restated objective
    rng = np.random.default_rng(3)
    F = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    F /= np.linalg.norm(F, axis=0)
    pd2 = objparams(p.Ne, p.Ng, p.T, p.nsteps, use_custom_forbidden=True, forb_states=F, forb_weights=[0.7, 0.2], **kw)
    assert np.abs(pd2.wmat_imag).max() > 1e-3
    _fd_check(pd2, pc)


def test_uncoupled_control_reduces_to_coupled_when_rfreq_is_zero():
    """ft = 2 (p cos 0 - q sin 0) = 2 p: an uncoupled symmetric control with Rfreq = 0 is a coupled control with Hsym = 2 Hunc and
    Hanti = 0 (whose q-spline then has no effect)."""
    from juqbox_b200.params import objparams
    cfg, pc = _small_swap()
    p = cfg.params
    Z = np.zeros_like(p.Hconst)
    kw = dict(Uinit=p.Uinit, Utarget=p.Utarget_r + 1j * p.Utarget_i, Cfreq=p.Cfreq, Hconst=p.Hconst, linear_solver=p.linear_solver)
    pc_ = objparams(p.Ne, p.Ng, p.T, p.nsteps, Rfreq=[0.0], Hsym_ops=[2 * p.Hsym_ops[0]], Hanti_ops=[Z], **kw)
    pu_ = objparams(p.Ne, p.Ng, p.T, p.nsteps, Rfreq=[0.0], Hunc_ops=[p.Hsym_ops[0]], **kw)
    pc_.wmat_real = pu_.wmat_real = p.wmat_real
    oc, ou = oracle_traceobjgrad(pc_, pc), oracle_traceobjgrad(pu_, pc)
    assert abs(oc["objf"][0, 0] - ou["objf"][0, 0]) < 1e-13
    assert np.linalg.norm(oc["grad"] - ou["grad"]) <= 1e-12 * np.linalg.norm(oc["grad"])


def test_uncoupled_lab_frame_rabi_fd_and_reference_pulse():
    cfg = configs.example("rabi_lab", T=20.0, Pmin=60)
    p = cfg.params
    pc = cfg.pcof0 * 5.0 + 0.3 * cfg.maxpar[0] * np.random.default_rng(1).standard_normal(cfg.nCoeff)
    _fd_check(p, pc, tol=5e-6)
    p.unc_grad_literal = 1           # the reference's own lines for this branch are NOT the gradient of its KS! model
    gl = oracle_traceobjgrad(p, pc)["grad"][0, 0]
    p.unc_grad_literal = 0
    ge = oracle_traceobjgrad(p, pc)["grad"][0, 0]
    assert np.linalg.norm(gl - ge) > 0.5 * np.linalg.norm(ge)
    # the pulse the reference optimised for the rotating-frame rabi model (examples/drives/rabi-pcof-opt-t100.jld2) drives the
    # lab-frame model of examples/rabi-lab.jl to the same gate up to the rotating-wave error
    full = configs.example("rabi_lab")
    drv = json.load(open(os.path.join(GOLDEN_DIR, "drives.json")))["rabi"]
    pcr = np.asarray(drv["pcof"] if isinstance(drv, dict) else drv, dtype=float)
    o = oracle_traceobjgrad(full.params, pcr, evaladjoint=False)
    assert o["infid"][0, 0] < 5e-4, o["infid"]
