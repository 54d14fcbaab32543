"""ctypes loader for the CPU oracle (oracle/traceobjgrad_oracle.c).

ORACLE — TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs;
never by the product package.  Takes a juqbox_b200 `objparams`-like object (duck-typed: only plain attributes
are read) so that the oracle and the CUDA path see byte-identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("Nc", C.c_int), ("Nfreq", C.c_int), ("J", C.c_int),
                ("objFuncType", C.c_int), ("sparse", C.c_int), ("nsteps", C.c_int64), ("T", C.c_double),
                ("Uinit", C.c_void_p), ("Vtr", C.c_void_p), ("Vti", C.c_void_p), ("wdiag", C.c_void_p),
                ("Cfreq", C.c_void_p), ("H0", C.c_void_p), ("Hsym", C.c_void_p), ("Hanti", C.c_void_p),
                ("colptr", C.c_void_p), ("rowval", C.c_void_p), ("nzval", C.c_void_p),
                ("solver", C.c_int), ("tol", C.c_double),
                # extensions (parity unpinned by the reference, see the header of traceobjgrad_oracle.c)
                ("pFidType", C.c_int), ("globalPhase", C.c_double), ("wmat_real", C.c_void_p), ("wmat_imag", C.c_void_p),
                ("Nunc", C.c_int), ("Hunc", C.c_void_p), ("isSymm", C.c_void_p), ("Rfreq", C.c_void_p),
                ("unc_grad_literal", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libjq_oracle.so")
    src = os.path.join(_HERE, "traceobjgrad_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.jqo_traceobjgrad_batch.restype = C.c_int
        _LIB.jqo_max_threads.restype = C.c_int
    return _LIB


def max_threads() -> int:
    return int(_lib().jqo_max_threads())


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _csc(a):
    """Dense -> CSC (0-based), dropping exact zeros like Julia's sparse()."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    colptr, rowval, nzval = [0], [], []
    for c in range(n):
        rows = np.nonzero(a[:, c])[0]
        rowval.extend(rows.tolist())
        nzval.extend(a[rows, c].tolist())
        colptr.append(len(rowval))
    return np.array(colptr, dtype=np.int64), np.array(rowval, dtype=np.int64), np.array(nzval, dtype=np.float64)


def _problem(params, keep):
    def ptr(a):
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)
    n, m, Nc = params.Ntot, params.N, params.Ncoupled
    P = _Problem()
    P.n, P.m, P.Nc, P.Nfreq = n, m, Nc, params.Nfreq
    P.J, P.objFuncType, P.sparse = params.linear_solver.max_iter, params.objFuncType, int(bool(params.use_sparse))
    P.nsteps, P.T = params.nsteps, params.T
    P.solver, P.tol = params.linear_solver.solver_id, params.linear_solver.tol
    P.Uinit, P.Vtr, P.Vti = ptr(_f(params.Uinit)), ptr(_f(params.Utarget_r)), ptr(_f(params.Utarget_i))
    Nunc = int(getattr(params, "Nunc", 0))
    wr = np.asarray(params.wmat_real, dtype=np.float64)
    if wr.ndim == 2:                               # custom forbidden states: dense real / imaginary weights
        P.wdiag = ptr(np.ascontiguousarray(np.diag(wr)))
        P.wmat_real = ptr(_f(wr))
        P.wmat_imag = ptr(_f(params.wmat_imag))
    else:
        P.wdiag = ptr(np.ascontiguousarray(wr))
    P.Cfreq = ptr(_f(params.Cfreq[:Nc + Nunc, :]))
    P.pFidType, P.globalPhase = int(getattr(params, "pFidType", 2)), float(getattr(params, "globalPhase", 0.0))
    P.Nunc, P.unc_grad_literal = Nunc, int(getattr(params, "unc_grad_literal", 0))
    hunc = list(getattr(params, "Hunc_ops", []))
    if Nunc:
        P.isSymm = ptr(np.ascontiguousarray(params.isSymm, dtype=np.int32))
        P.Rfreq = ptr(np.ascontiguousarray(np.asarray(params.Rfreq, dtype=np.float64)[:Nunc]))
        if not params.use_sparse:
            P.Hunc = ptr(np.concatenate([_f(h).ravel(order="F") for h in hunc]))
    if params.use_sparse:
        parts = [_csc(h) for h in [params.Hconst] + list(params.Hsym_ops) + list(params.Hanti_ops) + hunc]
        P.colptr = ptr(np.concatenate([p[0] for p in parts]))
        P.rowval = ptr(np.concatenate([p[1] for p in parts] + [np.zeros(1, np.int64)]))
        P.nzval = ptr(np.concatenate([p[2] for p in parts] + [np.zeros(1)]))
    else:
        P.H0 = ptr(_f(params.Hconst))
        P.Hsym = ptr(np.concatenate([_f(h).ravel(order="F") for h in params.Hsym_ops] + [np.zeros(1)]))
        P.Hanti = ptr(np.concatenate([_f(h).ravel(order="F") for h in params.Hanti_ops] + [np.zeros(1)]))
    return P


def oracle_forward_history(params, pcof, shift=None, save_every: int = 1):
    """Forward sweep of ONE trajectory with state history.  Returns (hist [nsave, N, Ntot] complex as Julia's
    Ntot x N x nsave read in C order, infid, leak)."""
    lib = _lib()
    keep = []
    P = _problem(params, keep)
    pcof = np.ascontiguousarray(pcof, dtype=np.float64)
    nsave = params.nsteps // save_every + 1
    hr = np.zeros((nsave, params.N, params.Ntot))
    hi = np.zeros((nsave, params.N, params.Ntot))
    out = np.zeros(2)
    sh = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    lib.jqo_forward_history.restype = C.c_int
    rc = lib.jqo_forward_history(C.byref(P), C.c_int(len(pcof)), pcof.ctypes.data_as(C.c_void_p),
                                 sh.ctypes.data_as(C.c_void_p) if sh is not None else None, C.c_int(save_every),
                                 hr.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("bad pcof length or nsteps not divisible by save_every")
    return hr + 1j * hi, out[0], out[1]


def oracle_eval_controls(params, pcof, times):
    """p_q(t), q_q(t) of every coupled control at `times` (reference evalctrl, plotstatectrl.jl:246).  Returns p, q [Nc, ntimes]."""
    lib = _lib()
    keep = []
    P = _problem(params, keep)
    pcof = np.ascontiguousarray(pcof, dtype=np.float64)
    times = np.ascontiguousarray(times, dtype=np.float64)
    p = np.zeros((params.Ncoupled + int(getattr(params, "Nunc", 0)), len(times)))
    q = np.zeros_like(p)
    lib.jqo_eval_controls.restype = C.c_int
    rc = lib.jqo_eval_controls(C.byref(P), C.c_int(len(pcof)), pcof.ctypes.data_as(C.c_void_p), C.c_int(len(times)),
                               times.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("bad pcof length")
    return p, q


def oracle_traceobjgrad(params, pcof, shifts=None, evaladjoint: bool = True, nthreads: int = 1):
    """Run the oracle on a batch: pcof [nbatch, Npar] (or [Npar]), shifts [nsamples, n] or None.

    Returns dict with objf, infid, leak, trace_infid  [nbatch, nsamples] and grad, infidgrad, leakgrad
    [nbatch, nsamples, Npar] (infidgrad == grad and leakgrad == 0 for objFuncType 1, as in the reference).
    """
    lib = _lib()
    pcof = np.ascontiguousarray(np.atleast_2d(np.asarray(pcof, dtype=np.float64)))
    nbatch, Npar = pcof.shape
    ext = 1 if int(getattr(params, "pFidType", 2)) == 3 else 0      # pFidType 3: last entry = global phase, gradients one longer
    Npar -= ext
    n, m, Nc = params.Ntot, params.N, params.Ncoupled + int(getattr(params, "Nunc", 0))
    keep = []

    def ptr(a):
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    P = _problem(params, keep)
    if shifts is not None:
        shifts = np.ascontiguousarray(np.atleast_2d(np.asarray(shifts, dtype=np.float64)))
        assert shifts.shape[1] == n
        nsamples = shifts.shape[0]
    else:
        nsamples = 1
    ntraj = nbatch * nsamples
    out = np.zeros((ntraj, 4))
    grad = np.zeros((ntraj, Npar + ext))
    infidgrad = np.zeros((ntraj, Npar + ext))
    leakgrad = np.zeros((ntraj, Npar + ext))
    rc = lib.jqo_traceobjgrad_batch(C.byref(P), C.c_int(Npar), C.c_int(nbatch), ptr(pcof), C.c_int(nsamples),
                                    ptr(shifts) if shifts is not None else None, C.c_int(int(evaladjoint)),
                                    C.c_int(nthreads), ptr(out), ptr(grad), ptr(infidgrad), ptr(leakgrad))
    if rc != 0:
        raise ValueError("pcof must have an even number of elements >= %d, not %d" % (3 * 2 * Nc, Npar))
    shp = (nbatch, nsamples)
    return {"objf": out[:, 0].reshape(shp), "infid": out[:, 1].reshape(shp), "leak": out[:, 2].reshape(shp),
            "trace_infid": out[:, 3].reshape(shp), "grad": grad.reshape(shp + (Npar + ext,)),
            "infidgrad": infidgrad.reshape(shp + (Npar + ext,)), "leakgrad": leakgrad.reshape(shp + (Npar + ext,))}
