"""Named problem configurations, restating the reference's setup scripts.

`test_case(name)` rebuilds the configurations behind the reference's golden files
(/root/reference/test/cases/<name>-setup.jl), `example(name)` the five BASELINE.json
configurations (/root/reference/examples/*.jl run with Integrator = Stormer_Verlet).
Each returns a `Config(params, pcof0, maxpar, ...)`; parameter vectors that the reference
reads from `.dat` files are passed in by the caller (tests read them from tests/golden/).

Synthetic `pcof` batches for throughput runs come from `synthetic_pcof`:
numpy default_rng(2456 + config index) with the amplitude distributions of the examples —
Julia's `rand` stream after Random.seed!(2456) is not reproducible outside Julia.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .params import (objparams, lsolver_object, JACOBI_SOLVER, calculate_timestep, estimate_Neumann, initial_cond,
                     orig_wmatsetup, setup_rotmatrices)

EPS = np.finfo(float).eps


@dataclass
class Config:
    name: str
    params: objparams
    pcof0: Optional[np.ndarray]
    D1: int
    maxpar: List[float]
    nodes: np.ndarray = field(default_factory=lambda: np.array([0.0]))
    weights: np.ndarray = field(default_factory=lambda: np.array([1.0]))

    @property
    def nCoeff(self) -> int:
        p = self.params
        return 2 * (p.Ncoupled + p.Nunc) * p.Nfreq * self.D1


def lowering(nt: int) -> np.ndarray:
    """Standard lowering operator (Bidiagonal(zeros, sqrt.(1:nt-1), :U))."""
    return np.diag(np.sqrt(np.arange(1, nt)), 1)


def kron_ops(Nt):
    """Lowering and number operators of each subsystem embedded in the full space.

    Subsystem 1 varies fastest: a1 = I (x) I (x) a, a2 = I (x) a (x) I, ... (reference
    test/cases/cnot3-setup.jl:80-82).
    """
    ops, nums = [], []
    for k, nt in enumerate(Nt):
        a, num = lowering(nt), np.diag(np.arange(nt, dtype=float))
        for j, ntj in enumerate(Nt):
            if j < k:
                a, num = np.kron(a, np.eye(ntj)), np.kron(num, np.eye(ntj))
            elif j > k:
                a, num = np.kron(np.eye(ntj), a), np.kron(np.eye(ntj), num)
        ops.append(a)
        nums.append(num)
    return ops, nums


def _nsteps_from_eig(T, H0, amats, maxpar, Pmin):
    K1 = H0.astype(complex)
    for a, mp in zip(amats, maxpar):
        K1 = K1 + mp * (a + a.T) + 1j * mp * (a - a.T)
    maxeig = np.max(np.abs(np.linalg.eigvalsh(K1)))
    return int(math.ceil(T * maxeig * Pmin / (2 * np.pi)))


def _rot_target(utarget, Ne, Ng, freqs, T):
    om = setup_rotmatrices(Ne, Ng, freqs)
    if not isinstance(om, tuple):
        om = (om,)
    rot = np.ones(utarget.shape[0], dtype=complex)
    for o in om:
        rot = rot * np.exp(1j * o * T)
    return rot[:, None] * utarget


# ----------------------------------------------------------------------------------------------
# golden (test/cases) configurations
# ----------------------------------------------------------------------------------------------

def _case_rabi(pcof0=None) -> Config:
    # test/cases/rabi-setup.jl
    N, Ng, Ntot = 2, 0, 2
    fa, xa = 0.0, 2 * 0.1099
    T = 2 * np.pi
    theta, aOmega = np.pi / 2, np.pi / T
    ut = np.eye(Ntot, N, dtype=complex)
    ut[0, 0] = math.cos(aOmega * T)
    ut[1, 0] = -(math.sin(theta) + 1j * math.cos(theta)) * math.sin(aOmega * T)
    ut[0, 1] = (math.sin(theta) - 1j * math.cos(theta)) * math.sin(aOmega * T)
    ut[1, 1] = math.cos(aOmega * T)
    vt = _rot_target(ut, [N], [Ng], [fa], T)
    num = np.diag(np.arange(Ntot, dtype=float))
    H0 = -0.5 * (2 * np.pi) * xa * (num @ num - num)
    a = lowering(Ntot)
    maxpar = aOmega
    nsteps = _nsteps_from_eig(T, H0, [a], [maxpar], 80)
    p = objparams([N], [Ng], T, nsteps, Uinit=np.eye(Ntot, N), Utarget=vt, Cfreq=np.zeros((1, 1)),
                  Rfreq=[fa], Hconst=H0, Hsym_ops=[a + a.T], Hanti_ops=[a - a.T])
    D1 = 3
    pc = np.zeros(2 * D1)
    pc[:D1] = aOmega * math.cos(theta)
    pc[D1:] = aOmega * math.sin(theta)
    estimate_Neumann(EPS, p, [maxpar])
    return Config("rabi", p, pc, D1, [maxpar])


def _case_swap02(pcof0) -> Config:
    # test/cases/swap02-setup.jl
    N, Ng = 3, 1
    Ntot = N + Ng
    T = 150.0
    ut = np.zeros((Ntot, N), dtype=complex)
    ut[2, 0] = ut[1, 1] = ut[0, 2] = 1
    vt = _rot_target(ut, [N], [Ng], [4.09947], T)
    xa = 2 * 0.1099
    num = np.diag(np.arange(Ntot, dtype=float))
    H0 = -0.5 * (2 * np.pi) * xa * (num @ num - num)
    a = lowering(Ntot)
    Nfreq = 2
    om = np.zeros((1, Nfreq))
    om[0, 1] = H0[2, 2]
    maxpar = 2 * np.pi * 0.0132 / Nfreq / 2
    nsteps = _nsteps_from_eig(T, H0, [a], [maxpar], 80)
    p = objparams([N], [Ng], T, nsteps, Uinit=np.eye(Ntot, N), Utarget=vt, Cfreq=om, Rfreq=[4.09947],
                  Hconst=H0, Hsym_ops=[a + a.T], Hanti_ops=[a - a.T])
    pc = np.asarray(pcof0, dtype=float)
    estimate_Neumann(EPS, p, [maxpar])
    return Config("swap02", p, pc, len(pc) // (2 * Nfreq), [maxpar])


def _case_cnot2(pcof0, objFuncType=1, name="cnot2", jacobi=False) -> Config:
    # test/cases/cnot2-setup.jl, cnot2-leakieq-setup.jl, cnot2-jacobi-setup.jl (:186 Jacobi solver object; the later
    # estimate_Neumann! call, :259, overwrites its max_iter exactly as it does for the Neumann solver)
    Ne, Ng = [2, 2], [1, 2]
    Nt = [3, 4]
    Ntot, N = 12, 4
    T = 100.0
    fa, fb = 4.10595, 4.81526
    x1, x2, x12 = 2 * 0.1099, 2 * 0.1126, 0.1
    (amat, bmat), (N1, N2) = kron_ops(Nt)
    H0 = -2 * np.pi * (x1 / 2 * (N1 @ N1 - N1) + x2 / 2 * (N2 @ N2 - N2) + x12 * (N1 @ N2))
    maxpar = [0.02, 0.05]
    nsteps = _nsteps_from_eig(T, H0, [amat, bmat], maxpar, 40)
    Nfreq = 2
    om = np.zeros((2, Nfreq))
    om[:, 1] = -2.0 * np.pi * x12
    ut = np.zeros((Ntot, N), dtype=complex)
    ut[0, 0] = ut[1, 1] = ut[3, 3] = ut[4, 2] = 1.0          # Ng1 == 1 branch
    vt = _rot_target(ut, Ne, Ng, [fa, fb], T)
    p = objparams(Ne, Ng, T, nsteps, Uinit=initial_cond(Ne, Ng), Utarget=vt, Cfreq=om, Rfreq=[fa, fb],
                  Hconst=H0, Hsym_ops=[amat + amat.T, bmat + bmat.T], Hanti_ops=[amat - amat.T, bmat - bmat.T],
                  use_sparse=False, objFuncType=objFuncType, leak_ubound=1e-3,
                  linear_solver=lsolver_object(solver=JACOBI_SOLVER, max_iter=100, tol=1e-15, nrhs=4) if jacobi else None)
    p.wmat_real = orig_wmatsetup(Ne, Ng)
    pc = np.asarray(pcof0, dtype=float)
    estimate_Neumann(EPS, p, maxpar)
    return Config(name, p, pc, len(pc) // (2 * 2 * Nfreq), maxpar)


def _cnot3_system(Ng3):
    Ne, Ng = [2, 2, 1], [2, 2, Ng3]
    Nt = [a + b for a, b in zip(Ne, Ng)]
    xa, xb = 2 * 0.1099, 2 * 0.1126
    xs = 0.002494 ** 2 / xa
    xab, xas, xbs = 1.0e-6, math.sqrt(xa * xs), math.sqrt(xb * xs)
    (amat, bmat, cmat), (Na, Nb, Nc) = kron_ops(Nt)
    H0 = -2 * np.pi * (xa / 2 * (Na @ Na - Na) + xb / 2 * (Nb @ Nb - Nb) + xs / 2 * (Nc @ Nc - Nc)
                       + xab * (Na @ Nb) + xas * (Na @ Nc) + xbs * (Nb @ Nc))
    return Ne, Ng, Nt, (xa, xb, xs, xab, xas, xbs), (amat, bmat, cmat), H0


def _cnot3_target(Ne, Ng, Nt, T):
    # test/cases/cnot3-setup.jl:206-235 (Ng[0] == 2 branch); examples/cnot3-setup.jl:163-180 is the same matrix
    G2 = np.zeros((Nt[0] * Nt[1], 4), dtype=complex)
    G2[0, 0] = G2[1, 1] = G2[4, 3] = G2[5, 2] = 1.0
    ut = np.kron(np.eye(Nt[2], Ne[2]), G2)
    return _rot_target(ut, Ne, Ng, [4.10595, 4.81526, 7.8447], T)


def _case_cnot3(pcof0) -> Config:
    # test/cases/cnot3-setup.jl (sparse, Nfreq = 3, 5 guard levels on the resonator)
    Ne, Ng, Nt, (xa, xb, xs, xab, xas, xbs), (amat, bmat, cmat), H0 = _cnot3_system(5)
    T = 550.0
    maxpar = [0.05, 0.1, 0.1]
    nsteps = _nsteps_from_eig(T, H0, [amat, bmat, cmat], maxpar, 40)
    Nfreq = 3
    om = np.zeros((3, Nfreq))
    om[0:2, 1] = -2.0 * np.pi * xa
    om[0:2, 2] = -2.0 * np.pi * xb
    om[2, 1] = -2.0 * np.pi * xas
    om[2, 2] = -2.0 * np.pi * xbs
    p = objparams(Ne, Ng, T, nsteps, Uinit=initial_cond(Ne, Ng), Utarget=_cnot3_target(Ne, Ng, Nt, T), Cfreq=om,
                  Rfreq=[4.10595, 4.81526, 7.8447], Hconst=H0,
                  Hsym_ops=[amat + amat.T, bmat + bmat.T, cmat + cmat.T],
                  Hanti_ops=[amat - amat.T, bmat - bmat.T, cmat - cmat.T], use_sparse=True)
    p.wmat_real = orig_wmatsetup(Ne, Ng)
    pc = np.asarray(pcof0, dtype=float)
    estimate_Neumann(EPS, p, maxpar)
    return Config("cnot3", p, pc, len(pc) // (2 * 3 * Nfreq), maxpar)


def _case_flux(pcof0) -> Config:
    # test/cases/flux-setup.jl (sparse, second "coupled" control has Hsym = a'a and Hanti = 0, tik0 = 0.1)
    N, Ng = 4, 2
    Ntot = N + Ng
    fa, xa, T = 5.0, 0.2, 11.0
    ut = np.eye(Ntot, N, dtype=complex)
    ut[:, 3] = np.eye(Ntot)[:, 2]
    ut[:, 2] = np.eye(Ntot)[:, 3]
    vt = _rot_target(ut, [N], [Ng], [fa], T)
    num = np.diag(np.arange(Ntot, dtype=float))
    H0 = -0.5 * (2 * np.pi) * xa * (num @ num - num)
    a = lowering(Ntot)
    Hsym = [a + a.T, a.T @ a]
    Hanti = [a - a.T, np.zeros((Ntot, Ntot))]
    Nfreq = 2
    om = np.zeros((2, Nfreq))
    om[:, 1] = -2.0 * np.pi * xa
    maxpar, max_flux = 0.08, 2 * np.pi * 5.0
    nsteps = calculate_timestep(T, H0, Hsym, Hanti, [maxpar, max_flux])
    p = objparams([N], [Ng], T, nsteps, Uinit=np.eye(Ntot, N), Utarget=vt, Cfreq=om, Rfreq=[fa, fa],
                  Hconst=H0, Hsym_ops=Hsym, Hanti_ops=Hanti, use_sparse=True)
    p.tik0 = 0.1
    p.traceInfidelityThreshold = 1e-5
    pc = np.asarray(pcof0, dtype=float)
    return Config("flux", p, pc, len(pc) // (2 * 2 * Nfreq), [maxpar, max_flux])


def test_case(name: str, pcof0=None) -> Config:
    """Configuration behind test/reference_solutions/<name>-ref.jld2."""
    if name == "rabi":
        return _case_rabi()
    if name == "swap02":
        return _case_swap02(pcof0)
    if name == "cnot2":
        return _case_cnot2(pcof0)
    if name == "cnot2-leakieq":
        return _case_cnot2(pcof0, objFuncType=3, name="cnot2-leakieq")
    if name == "cnot2-jacobi":
        return _case_cnot2(pcof0, name="cnot2-jacobi", jacobi=True)
    if name == "cnot3":
        return _case_cnot3(pcof0)
    if name == "flux":
        return _case_flux(pcof0)
    raise KeyError(name)


test_case.__test__ = False  # not a pytest test

# ----------------------------------------------------------------------------------------------
# BASELINE.json configurations (examples/, Stormer-Verlet)
# ----------------------------------------------------------------------------------------------
EXAMPLES = ["rabi", "cnot1", "cnot2", "cnot3", "risk_neutral"]


def _ex_rabi() -> Config:
    # examples/rabi-setup.jl
    N, Ng, Ntot = 2, 0, 2
    fa, xa, T = 5.0, 2 * 0.1099, 100.0
    theta, aOmega = np.pi / 4, np.pi / T
    ut = np.eye(Ntot, N, dtype=complex)
    ut[0, 0] = math.cos(aOmega * T)
    ut[1, 0] = -(math.sin(theta) + 1j * math.cos(theta)) * math.sin(aOmega * T)
    ut[0, 1] = (math.sin(theta) - 1j * math.cos(theta)) * math.sin(aOmega * T)
    ut[1, 1] = math.cos(aOmega * T)
    vt = _rot_target(ut, [N], [Ng], [fa], T)
    num = np.diag(np.arange(Ntot, dtype=float))
    H0 = -0.5 * (2 * np.pi) * xa * (num @ num - num)
    a = lowering(Ntot)
    maxpar = aOmega
    nsteps = calculate_timestep(T, H0, [a + a.T], [a - a.T], [maxpar], 80)
    p = objparams([N], [Ng], T, nsteps, Uinit=np.eye(Ntot, N), Utarget=vt, Cfreq=np.zeros((1, 1)), Rfreq=[fa],
                  Hconst=H0, Hsym_ops=[a + a.T], Hanti_ops=[a - a.T])
    D1 = 3
    pc = np.zeros(2 * D1)
    pc[:D1] = aOmega * math.cos(theta)
    pc[D1:] = aOmega * math.sin(theta)
    return Config("rabi", p, pc, D1, [maxpar])


def _ex_rabi_lab(T: float = 100.0, Pmin: int = 100) -> Config:
    # examples/rabi-lab.jl: the single-qubit pi-pulse in the LAB frame, one UNCOUPLED control a + a' with two splines (p, q):
    # f(t) = 2 (p(t) cos(2 pi fa t) - q(t) sin(2 pi fa t))  (KS!, src/evalobjgrad.jl:2372-2387); target not rotated (:79-80)
    N, Ng, Ntot = 2, 0, 2
    fa, xa = 5.0, 2 * 0.1099
    theta, aOmega = np.pi / 4, np.pi / 100.0
    ut = np.eye(Ntot, N, dtype=complex)
    ut[0, 0] = math.cos(aOmega * T)
    ut[1, 0] = -(math.sin(theta) + 1j * math.cos(theta)) * math.sin(aOmega * T)
    ut[0, 1] = (math.sin(theta) - 1j * math.cos(theta)) * math.sin(aOmega * T)
    ut[1, 1] = math.cos(aOmega * T)
    num = np.diag(np.arange(Ntot, dtype=float))
    H0 = 2 * np.pi * (fa * num - 0.5 * xa * (num @ num - num))
    a = lowering(Ntot)
    maxpar = aOmega
    nsteps = calculate_timestep(T, H0, [a + a.T], [np.zeros((Ntot, Ntot))], [maxpar], Pmin)
    p = objparams([N], [Ng], T, nsteps, Uinit=np.eye(Ntot, N), Utarget=ut, Cfreq=np.zeros((1, 1)), Rfreq=[fa],
                  Hconst=H0, Hunc_ops=[a + a.T])
    D1 = 3
    pc = np.zeros(2 * D1)
    pc[:D1] = aOmega * math.cos(theta)
    pc[D1:] = aOmega * math.sin(theta)
    return Config("rabi_lab", p, pc, D1, [maxpar])


def _ex_cnot1() -> Config:
    # examples/cnot1-setup.jl (Integrator_id = 2 there; built here with Stormer-Verlet)
    N, Ng = 4, 2
    Ntot = N + Ng
    T, fa, xa = 100.0, 4.10336, 0.2198
    num = np.diag(np.arange(Ntot, dtype=float))
    H0 = -0.5 * (2 * np.pi) * xa * (num @ num - num)
    a = lowering(Ntot)
    maxctrl = 0.001 * 2 * np.pi * 8.5
    nsteps = calculate_timestep(T, H0, [a + a.T], [a - a.T], [maxctrl])
    Nfreq = 3
    om = np.zeros((1, Nfreq))
    om[0, 1] = -2.0 * np.pi * xa
    om[0, 2] = -2.0 * np.pi * 2.0 * xa
    maxamp = np.zeros(Nfreq)
    maxamp[0] = maxctrl * 0.45
    maxamp[1:] = maxctrl * (1.0 - 0.45) / (Nfreq - 1)
    U0 = initial_cond([N], [Ng])
    gate = np.zeros((N, N), dtype=complex)
    gate[0, 0] = gate[1, 1] = gate[2, 3] = gate[3, 2] = 1.0
    vt = _rot_target(U0 @ gate, [N], [Ng], [fa], T)
    p = objparams([N], [Ng], T, nsteps, Uinit=U0, Utarget=vt, Cfreq=om, Rfreq=[fa], Hconst=H0,
                  Hsym_ops=[a + a.T], Hanti_ops=[a - a.T])
    return Config("cnot1", p, None, 10, [float(maxamp.max())])


def _ex_cnot2(T: float = 50.0) -> Config:
    # examples/cnot2-setup.jl (rotating frame branch, sparse); T is the script's gate duration (50 ns as shipped; the
    # reference also ships pulses optimised for 100 and 200 ns, examples/drives/cnot2-pcof-opt-t{50,100,200}.jld2)
    Ne, Ng = [2, 2], [2, 2]
    Nt = [4, 4]
    fa, fb = 4.10595, 4.81526
    x1, x2, x12 = 2 * 0.1099, 2 * 0.1126, 0.1
    (amat, bmat), (N1, N2) = kron_ops(Nt)
    H0 = 2 * np.pi * ((fa - fa) * N1 + (fb - fb) * N2 - x1 / 2 * (N1 @ N1 - N1) - x2 / 2 * (N2 @ N2 - N2) - x12 * (N1 @ N2))
    Hsym = [amat + amat.T, bmat + bmat.T]
    Hanti = [amat - amat.T, bmat - bmat.T]
    maxpar = [0.040, 0.040]
    nsteps = calculate_timestep(T, H0, Hsym, Hanti, maxpar, 40)
    Nfreq = 2
    om = np.zeros((2, Nfreq))
    om[0, 0] = 2 * np.pi * (fa - fa)
    om[0, 1] = 2 * np.pi * (fa - fa - x12)
    om[1, 0] = 2 * np.pi * (fb - fb)
    om[1, 1] = 2 * np.pi * (fb - fb - x12)
    gate = np.zeros((4, 4), dtype=complex)
    gate[0, 0] = gate[1, 1] = gate[2, 3] = gate[3, 2] = 1.0
    U0 = initial_cond(Ne, Ng)
    vt = _rot_target(U0 @ gate, Ne, Ng, [fa, fb], T)
    p = objparams(Ne, Ng, T, nsteps, Uinit=U0, Utarget=vt, Cfreq=om, Rfreq=[fa, fb], Hconst=H0,
                  Hsym_ops=Hsym, Hanti_ops=Hanti, use_sparse=True)
    estimate_Neumann(1e-12, p, maxpar)
    return Config("cnot2", p, None, 10, maxpar)


def _ex_cnot3(Nfreq: int = 2) -> Config:
    # examples/cnot3-setup.jl (sparse, Nfreq = 2 as shipped, 3 guard levels on the resonator, default J = 3); Nfreq = 3 is the
    # script's alternative branch (:153-158), the one examples/drives/cnot3-pcof-opt.jld2 was optimised with
    Ne, Ng, Nt, (xa, xb, xs, xab, xas, xbs), (amat, bmat, cmat), H0 = _cnot3_system(3)
    T = 550.0
    maxpar = [0.05, 0.1, 0.1]
    Hsym = [amat + amat.T, bmat + bmat.T, cmat + cmat.T]
    Hanti = [amat - amat.T, bmat - bmat.T, cmat - cmat.T]
    nsteps = calculate_timestep(T, H0, Hsym, Hanti, maxpar, 40)
    om = np.zeros((3, Nfreq))
    if Nfreq == 2:
        om[0, 1] = -2.0 * np.pi * xa
        om[1, 1] = -2.0 * np.pi * xb
        om[2, 1] = -2.0 * np.pi * math.sqrt(xas * xbs)
    elif Nfreq == 3:
        om[0:2, 1] = -2.0 * np.pi * xa
        om[0:2, 2] = -2.0 * np.pi * xb
        om[2, 1] = -2.0 * np.pi * xas
        om[2, 2] = -2.0 * np.pi * xbs
    else:
        raise ValueError("cnot3 example: Nfreq must be 2 or 3")
    gate = np.zeros((4, 4), dtype=complex)
    gate[0, 0] = gate[1, 1] = gate[2, 3] = gate[3, 2] = 1.0
    U0 = initial_cond(Ne, Ng)
    vt = _rot_target(U0 @ gate, Ne, Ng, [4.10595, 4.81526, 7.8447], T)
    p = objparams(Ne, Ng, T, nsteps, Uinit=U0, Utarget=vt, Cfreq=om, Rfreq=[4.10595, 4.81526, 7.8447],
                  Hconst=H0, Hsym_ops=Hsym, Hanti_ops=Hanti, use_sparse=True)
    return Config("cnot3", p, None, 15, maxpar)


def _ex_risk_neutral(nquad: int = 9) -> Config:
    # examples/Risk_Neutral/swap-02-risk-neutral.jl with run_all.jl's ep_max = 2pi*2e-2, nquad = 9
    ep_max = 2 * np.pi * 2e-2
    nodes, weights = np.polynomial.legendre.leggauss(nquad)
    nodes = nodes * 0.5 * ep_max
    weights = weights * 0.5
    N, Ng = 3, 1
    Ntot = N + Ng
    T, fa, xa = 300.0, 4.10336, 0.2198
    num = np.diag(np.arange(Ntot, dtype=float))
    H0 = -0.5 * (2 * np.pi) * xa * (num @ num - num)
    ut = np.zeros((Ntot, N), dtype=complex)
    ut[2, 0] = ut[1, 1] = ut[0, 2] = 1
    a = lowering(Ntot)
    Nfreq = 2
    om = np.zeros((1, Nfreq))
    om[0, 1] = -2.0 * np.pi * xa
    maxctrl = 2 * np.pi * 1.2e-2
    maxpar = maxctrl / Nfreq
    nsteps = calculate_timestep(T, H0, [a + a.T], [a - a.T], [maxctrl])
    p = objparams([N], [Ng], T, nsteps, Uinit=initial_cond([N], [Ng]), Utarget=ut, Cfreq=om, Rfreq=[fa],
                  Hconst=H0, Hsym_ops=[a + a.T], Hanti_ops=[a - a.T], wmatScale=1.0)
    estimate_Neumann(EPS, p, [maxpar])
    return Config("risk_neutral", p, None, 12, [maxpar], nodes, weights)


def qudit_system(Ne, Ng, T: float = 20.0, Nfreq: int = 2, D1: int = 6, maxamp: float = 0.03, exchange: float = 0.0,
                 use_sparse: Optional[bool] = None, seed: int = 11) -> Config:
    """Coupled anharmonic qudits in the rotating frame, built the way every examples/*-setup.jl builds its model
    (e.g. examples/cnot2-setup.jl:76-135): H0 = -sum_k x_k/2 (N_k^2 - N_k) - sum_{k<l} x_kl N_k N_l (diagonal),
    one control pair (a_k + a_k', a_k - a_k') per subsystem, carrier frequencies 0 and -x-shift, a random unitary
    permutation of the essential levels as the target.  `exchange` adds J (a_k' a_l + a_k a_l') to H0 (off-diagonal
    drift, Jaynes-Cummings-type coupling), which only the generic kernel handles.  Used to exercise kernel selection
    on shapes beyond the five named configurations; the oracle is the pin.
    """
    Ne, Ng = list(Ne), list(Ng)
    Nt = [a + b for a, b in zip(Ne, Ng)]
    nsub = len(Nt)
    rng = np.random.default_rng(seed)
    amats, nums = kron_ops(Nt)
    xk = [2 * np.pi * (0.2 + 0.02 * k) for k in range(nsub)]
    n = int(np.prod(Nt))
    H0 = np.zeros((n, n))
    for k in range(nsub):
        H0 -= xk[k] / 2 * (nums[k] @ nums[k] - nums[k])
        for l in range(k + 1, nsub):
            H0 -= 2 * np.pi * 0.05 * (nums[k] @ nums[l])
            if exchange:
                H0 += exchange * (amats[k].T @ amats[l] + amats[k] @ amats[l].T)
    Hsym = [a + a.T for a in amats]
    Hanti = [a - a.T for a in amats]
    maxpar = [maxamp] * nsub
    nsteps = calculate_timestep(T, H0, Hsym, Hanti, maxpar, 40)
    om = np.zeros((nsub, Nfreq))
    for f in range(1, Nfreq):
        om[:, f] = [-xk[k] * f for k in range(nsub)]
    U0 = initial_cond(Ne, Ng)
    N = U0.shape[1]
    gate = np.eye(N)[:, rng.permutation(N)].astype(complex)
    if use_sparse is None:
        use_sparse = nsub > 1
    p = objparams(Ne, Ng, T, nsteps, Uinit=U0, Utarget=U0 @ gate, Cfreq=om, Rfreq=[4.0 + k for k in range(nsub)],
                  Hconst=H0, Hsym_ops=Hsym, Hanti_ops=Hanti, use_sparse=use_sparse)
    estimate_Neumann(1e-12, p, maxpar)
    cfg = Config("qudits" + "x".join(str(t) for t in Nt), p, None, D1, maxpar)
    cfg.pcof0 = rng.uniform(-1, 1, cfg.nCoeff) * maxamp * 0.5
    return cfg


def example(name: str, **kw) -> Config:
    """One of the five BASELINE.json configurations (keyword arguments: the knobs the example script itself exposes,
    e.g. example("cnot2", T=100.0))."""
    return {"rabi": _ex_rabi, "cnot1": _ex_cnot1, "cnot2": _ex_cnot2, "cnot3": _ex_cnot3,
            "risk_neutral": _ex_risk_neutral, "rabi_lab": _ex_rabi_lab}[name](**kw)


def synthetic_pcof(cfg: Config, nbatch: int, seed_offset: int = 0) -> np.ndarray:
    """[nbatch, Npar] synthetic parameter vectors with the amplitude distribution of the example script."""
    idx = EXAMPLES.index(cfg.name) if cfg.name in EXAMPLES else 7
    rng = np.random.default_rng(2456 + idx + 1000 * seed_offset)
    n = cfg.nCoeff
    if cfg.name == "rabi":
        base = cfg.pcof0[None, :]
        return base + (rng.random((nbatch, n)) - 0.5) * 0.1 * cfg.maxpar[0]
    if cfg.name == "risk_neutral":
        return (rng.random((nbatch, n)) - 0.5) * 0.1 * cfg.maxpar[0]
    amp = {"cnot1": 0.01 * cfg.maxpar[0], "cnot2": 0.01 * 0.04, "cnot3": 0.01 * 0.05}.get(cfg.name, 0.01 * cfg.maxpar[0])
    return amp * rng.random((nbatch, n))


def noise_shift(n: int, eps) -> np.ndarray:
    """Additive diagonal noise of the risk-neutral sample loop: H0[j,j] += 0.01*eps*10^(j-2), j = 2..n
    (1-based; reference src/ipopt_interface.jl:41-44). Returns [len(eps), n]."""
    eps = np.atleast_1d(np.asarray(eps, dtype=float))
    fac = np.zeros(n)
    for j in range(2, n + 1):
        fac[j - 1] = 0.01 * 10.0 ** (j - 2)
    return eps[:, None] * fac[None, :]
