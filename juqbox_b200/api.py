"""Host-side mirror of the reference's evaluation API, calling the CUDA library through the C ABI.

  Working_Arrays(params, nCoeff)                      src/evalobjgrad.jl:405   -> owns the device handle
  traceobjgrad(pcof0, params, wa, verbose, evaladjoint)  src/evalobjgrad.jl:504
  eval_f_g_grad!(pcof, params, wa, nodes, weights, compute_adjoint)   src/ipopt_interface.jl:24-70
  eval_f_par / eval_grad_f_par / eval_g_par / eval_jac_g_par           src/ipopt_interface.jl:77-179
plus the batched forms the reference only has as serial loops (traceobjgrad_batch, ep_sweep).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .configs import noise_shift
from .params import objparams, tikhonov_grad, tikhonov_pen


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _csc(a):
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    colptr, rowval, nzval = [0], [], []
    for c in range(n):
        rows = np.nonzero(a[:, c])[0]          # sparse() drops exact zeros (src/evalobjgrad.jl:265-288)
        rowval.extend(rows.tolist())
        nzval.extend(a[rows, c].tolist())
        colptr.append(len(rowval))
    return (np.array(colptr, dtype=np.int64), np.array(rowval + [0], dtype=np.int64)[:max(len(rowval), 1)],
            np.array(nzval + [0.0], dtype=np.float64)[:max(len(nzval), 1)], len(nzval))


class Working_Arrays:
    """Device-resident replacement of the reference's Working_Arrays: one handle per (params, nCoeff) and GPU."""

    def __init__(self, params: objparams, nCoeff: int, device: Optional[int] = None):
        lib = _lib.load()
        self._lib = lib
        self.params = params
        self.nCoeff = int(nCoeff)
        nsig = 2 * (params.Ncoupled + params.Nunc)
        nspl = nCoeff - (1 if params.pFidType == 3 else 0)      # pFidType 3: the last entry is the global phase (:591-596)
        if nspl % (nsig * params.Nfreq) != 0 or nspl < 3 * nsig:
            raise ValueError(f"pcof must have an even number of elements >= {3 * nsig}, not {nspl}")
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0")) if "JUQBOX_B200_USE_LOCAL_RANK" in os.environ else 0
        self.device = device
        self._keep = []
        self.comm_size = 1
        self._handle = C.c_void_p()
        pb = self._describe(params)
        _lib.check(lib.jq_create(C.byref(pb), C.c_int(device), C.byref(self._handle)))
        self._snapshot = self._fingerprint(params)

    # The reference re-reads `params` on every traceobjgrad call; the device handle copies it once.  Fields the kernels
    # depend on are fingerprinted so that a later edit (estimate_Neumann, params.nsteps = ..., a new Hconst) is an error
    # instead of being silently ignored.  Utarget is the one supported mutation (update_target).
    @staticmethod
    def _fingerprint(p):
        import hashlib
        hsh = hashlib.sha1()
        for a in [p.Hconst, p.Uinit, p.wmat_real, p.Cfreq] + list(p.Hsym_ops) + list(p.Hanti_ops) + list(p.Hunc_ops) + \
                ([p.wmat_imag] if p.wmat_imag is not None else []):
            hsh.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        return (p.T, p.nsteps, p.linear_solver.max_iter, p.linear_solver.solver_id, p.linear_solver.tol, p.objFuncType, p.pFidType,
                p.globalPhase, tuple(p.Rfreq[:p.Nunc]), hsh.hexdigest())

    def _check_params(self):
        if self._fingerprint(self.params) != self._snapshot:
            raise RuntimeError("params changed after Working_Arrays was created (T, nsteps, linear_solver, objFuncType, Hconst, "
                               "operators, weights or Cfreq): create a new Working_Arrays (only Utarget can be updated in place)")

    # -- problem descriptor -------------------------------------------------------------------
    def _ptr(self, a):
        self._keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    def _op(self, mat, sparse):
        op = _lib.jq_operator()
        if sparse:
            cp, rv, nz, nnz = _csc(mat)
            op.format, op.nnz = _lib.JQ_CSC, nnz
            op.colptr, op.rowval, op.nzval = self._ptr(cp), self._ptr(rv), self._ptr(nz)
        else:
            a = np.asfortranarray(mat, dtype=np.float64)
            op.format, op.nnz = _lib.JQ_DENSE, a.size
            op.nzval = self._ptr(a)
        return op

    def _describe(self, p: objparams):
        pb = _lib.jq_problem()
        pb.n, pb.m, pb.ncoupled, pb.nfreq = p.Ntot, p.N, p.Ncoupled, p.Nfreq
        pb.neumann_terms = p.linear_solver.max_iter
        pb.linear_solver, pb.solver_tol = p.linear_solver.solver_id, p.linear_solver.tol
        pb.obj_func_type, pb.pfid_type = p.objFuncType, p.pFidType
        pb.nsteps, pb.T = p.nsteps, p.T
        pb.uinit = self._ptr(np.asfortranarray(p.Uinit, dtype=np.float64))
        pb.vtarget_r = self._ptr(np.asfortranarray(p.Utarget_r, dtype=np.float64))
        pb.vtarget_i = self._ptr(np.asfortranarray(p.Utarget_i, dtype=np.float64))
        wr = np.asarray(p.wmat_real, dtype=np.float64)
        if wr.ndim == 2:                       # custom forbidden states: dense real / imaginary weights (src/evalobjgrad.jl:214-232)
            pb.wdiag = self._ptr(_f64(np.diag(wr)))
            pb.wmat_real = self._ptr(np.asfortranarray(wr))
            if p.wmat_imag is not None and np.any(p.wmat_imag):
                pb.wmat_imag = self._ptr(np.asfortranarray(p.wmat_imag, dtype=np.float64))
        else:
            pb.wdiag = self._ptr(_f64(wr))
        pb.cfreq = self._ptr(np.asfortranarray(p.Cfreq[:p.Ncoupled + p.Nunc, :], dtype=np.float64))
        pb.global_phase = float(p.globalPhase)
        pb.h0 = self._op(p.Hconst, p.use_sparse)
        OpArr = _lib.jq_operator * max(p.Ncoupled, 1)
        hs = OpArr(*[self._op(h, p.use_sparse) for h in p.Hsym_ops])
        ha = OpArr(*[self._op(h, p.use_sparse) for h in p.Hanti_ops])
        self._keep += [hs, ha]
        pb.hsym, pb.hanti = hs, ha
        pb.nuncoupled = p.Nunc
        if p.Nunc:
            hu = (_lib.jq_operator * p.Nunc)(*[self._op(h, p.use_sparse) for h in p.Hunc_ops])
            self._keep.append(hu)
            pb.hunc = hu
            pb.unc_is_symm = self._ptr(np.ascontiguousarray(p.isSymm, dtype=np.int32))
            pb.unc_rfreq = self._ptr(_f64(np.asarray(p.Rfreq, dtype=np.float64)[:p.Nunc]))
        return pb

    # -- lifecycle ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.jq_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_target(self):
        """Push params.Utarget_r/i to the device after change_target (src/evalobjgrad.jl:1492)."""
        p = self.params
        vr, vi = np.asfortranarray(p.Utarget_r, dtype=np.float64), np.asfortranarray(p.Utarget_i, dtype=np.float64)
        _lib.check(self._lib.jq_update_target(self._handle, vr.ctypes.data_as(C.c_void_p), vi.ctypes.data_as(C.c_void_p)))      # also resets the cache

    def comm_init(self, rank: int, nranks: int, unique_id: bytes = None):
        """Attach an NCCL communicator (jq_comm_init).  With torch.distributed initialised, the 128-byte unique id is
        created on rank 0 and broadcast through it; otherwise pass `unique_id` explicitly."""
        if unique_id is None:
            import torch
            import torch.distributed as dist
            buf = C.create_string_buffer(128)
            if rank == 0:
                _lib.check(self._lib.jq_comm_unique_id(buf))
            t = torch.tensor(list(buf.raw), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda(self.device)
            dist.broadcast(t, src=0)
            unique_id = bytes(t.cpu().tolist())
        assert len(unique_id) == 128
        _lib.check(self._lib.jq_comm_init(self._handle, int(rank), int(nranks), C.c_char_p(unique_id)))
        self.comm_size = int(nranks)        # weighted evaluations now end with the library's own all-reduce

    def comm_set_cooperative(self, on: bool = True):
        """All ranks of the communicator evaluate the SAME arguments together: the time segments of kernel 7 are shared out and the
        propagators all-gathered; every rank returns the single-GPU bits (jq_comm_set_cooperative)."""
        _lib.check(self._lib.jq_comm_set_cooperative(self._handle, int(bool(on))))

    def comm_destroy(self):
        _lib.check(self._lib.jq_comm_destroy(self._handle))
        self.comm_size = 1

    def set_kernel(self, kernel: int):
        """0 = automatic, 1 = generic kernel, 2 = slot layout, 3 = fibre layout, 4 = tile layout, 5 = latency layout (pipelined roles),
        6 = dense-operator kernel on FP64 tensor cores, 7 = time-parallel evaluation (segments of the time axis swept concurrently)."""
        _lib.check(self._lib.jq_set_kernel(self._handle, int(kernel)))

    def set_time_segments(self, nseg: int) -> None:
        """Number of time segments of kernel 7 (0 = automatic)."""
        _lib.check(self._lib.jq_set_time_segments(self._handle, int(nseg)))

    def query(self, what: int) -> float:
        v = C.c_double()
        _lib.check(self._lib.jq_query(self._handle, int(what), C.byref(v)))
        return v.value

    @property
    def last_kernel(self) -> int:
        return int(self.query(0))

    @property
    def last_kernel_ms(self) -> float:
        return self.query(1)

    # -- evaluation ---------------------------------------------------------------------------
    def evaluate(self, pcof, shifts=None, weights=None, evaladjoint=True, out=None):
        """Batched evaluation with host arrays (copies inside): see jq_traceobjgrad_batch.

        `out`: a dict returned by an earlier call with the same shapes (or built by the caller, e.g. on pinned
        memory) whose arrays are reused instead of allocating fresh ones — the Working_Arrays idea applied to outputs.
        With objFuncType == 1 `infidgrad` is the same array as `grad` (the reference's infidelgrad aliases totalgrad,
        src/evalobjgrad.jl:951) and `leakgrad` is a read-only zero view: only one gradient crosses the bus."""
        p = self.params
        self._check_params()
        pcof = _f64(np.atleast_2d(pcof))
        nbatch, npar = pcof.shape
        nsamples = 1
        sp = wp = None
        if shifts is not None:
            shifts = _f64(np.atleast_2d(shifts))
            if shifts.shape[1] != p.Ntot:
                raise ValueError("shifts must be [nsamples, Ntot]")
            nsamples = shifts.shape[0]
            sp = shifts.ctypes.data_as(C.c_void_p)
        if weights is not None:
            weights = _f64(np.atleast_1d(weights))
            if len(weights) != nsamples:
                raise ValueError("weights must have one entry per sample")
            wp = weights.ctypes.data_as(C.c_void_p)
        shape = (nbatch,) if weights is not None else (nbatch, nsamples)
        two_sets = p.objFuncType != 1

        def buf(key, shp):
            a = out.get(key) if out is not None else None
            if a is None or a.shape != shp or a.dtype != np.float64 or not a.flags.c_contiguous or not a.flags.writeable:
                a = np.zeros(shp)
            return a
        res = {k: buf(k, shape) for k in ("infid", "leak", "trace_infid")}
        gptr = [None, None, None]
        if evaladjoint:
            res["grad"] = buf("grad", shape + (npar,))
            gptr[0] = res["grad"].ctypes.data_as(C.c_void_p)
            if two_sets:
                res["infidgrad"] = buf("infidgrad", shape + (npar,))
                res["leakgrad"] = buf("leakgrad", shape + (npar,))
                gptr[1:] = [res[k].ctypes.data_as(C.c_void_p) for k in ("infidgrad", "leakgrad")]
            else:
                res["infidgrad"] = res["grad"]
                res["leakgrad"] = np.broadcast_to(0.0, shape + (npar,))
        out = res
        _lib.check(self._lib.jq_traceobjgrad_batch(
            self._handle, nbatch, pcof.ctypes.data_as(C.c_void_p), npar, nsamples, sp, wp, int(bool(evaladjoint)),
            out["infid"].ctypes.data_as(C.c_void_p), out["leak"].ctypes.data_as(C.c_void_p),
            out["trace_infid"].ctypes.data_as(C.c_void_p), *gptr))
        out["objf"] = out["infid"] + out["leak"]
        return out

    def forward_history(self, pcof, shifts=None, save_every: int = 1):
        """Forward sweep with state history (jq_eval_forward).  Returns (hist, infid, leak): hist complex
        [nbatch, nsamples, nsave, N, Ntot] — for one trajectory hist[b, s].transpose(2, 1, 0) is Julia's Ntot x N x nsave."""
        p = self.params
        self._check_params()
        pcof = _f64(np.atleast_2d(pcof))
        nbatch, npar = pcof.shape
        nsamples, sp = 1, None
        if shifts is not None:
            shifts = _f64(np.atleast_2d(shifts))
            if shifts.shape[1] != p.Ntot:
                raise ValueError("shifts must be [nsamples, Ntot]")
            nsamples, sp = shifts.shape[0], shifts.ctypes.data_as(C.c_void_p)
        if save_every < 1 or p.nsteps % save_every != 0:
            raise ValueError(f"nsteps must be divisible by saveEvery. nsteps={p.nsteps}, saveEvery={save_every}")
        nsave = p.nsteps // save_every + 1
        hr = np.zeros((nbatch, nsamples, nsave, p.N, p.Ntot))
        hi = np.zeros_like(hr)
        infid, leak = np.zeros((nbatch, nsamples)), np.zeros((nbatch, nsamples))
        _lib.check(self._lib.jq_eval_forward(self._handle, nbatch, pcof.ctypes.data_as(C.c_void_p), npar, nsamples, sp, int(save_every),
                                             hr.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                                             infid.ctypes.data_as(C.c_void_p), leak.ctypes.data_as(C.c_void_p)))
        return hr + 1j * hi, infid, leak

    def eval_f_grad(self, pcof, shifts=None, weights=None, tik0: float = 0.0, prior=None):
        """The fused Ipopt-callback entry (jq_eval_f_grad): objective with Tikhonov, its gradient, infidelity, leak and leak
        gradient of ONE pcof in one call, with the last-pcof cache inside the handle.  Returns a dict; `evaluated` is False
        when the call was served from the cache."""
        p = self.params
        self._check_params()
        pcof = _f64(np.ravel(pcof))
        npar = len(pcof)
        nsamples, sp, wp, pp = 1, None, None, None
        if shifts is not None:
            shifts = _f64(np.atleast_2d(shifts))
            if shifts.shape[1] != p.Ntot:
                raise ValueError("shifts must be [nsamples, Ntot]")
            nsamples, sp = shifts.shape[0], shifts.ctypes.data_as(C.c_void_p)
        if weights is not None:
            weights = _f64(np.atleast_1d(weights))
            if len(weights) != nsamples:
                raise ValueError("weights must have one entry per sample")
            wp = weights.ctypes.data_as(C.c_void_p)
        if prior is not None:
            prior = _f64(np.ravel(prior))
            if len(prior) != npar:
                raise ValueError("prior must have the length of pcof")
            pp = prior.ctypes.data_as(C.c_void_p)
        f, infid, leak = C.c_double(), C.c_double(), C.c_double()
        ev = C.c_int32()
        grad, lgrad = np.zeros(npar), np.zeros(npar if p.objFuncType != 1 else 0)
        _lib.check(self._lib.jq_eval_f_grad(self._handle, pcof.ctypes.data_as(C.c_void_p), npar, nsamples, sp, wp, float(tik0), pp,
                                            C.byref(f), grad.ctypes.data_as(C.c_void_p), C.byref(infid), C.byref(leak),
                                            lgrad.ctypes.data_as(C.c_void_p) if len(lgrad) else None, C.byref(ev)))
        return {"f": f.value, "grad_f": grad, "infid": infid.value, "leak": leak.value, "leakgrad": lgrad, "evaluated": bool(ev.value)}

    def controls(self, pcof, times):
        """p_q(t), q_q(t) of every coupled control at `times` (jq_eval_controls).  Returns p, q of shape [Ncoupled, ntimes]."""
        pcof, times = _f64(np.ravel(pcof)), _f64(np.ravel(times))
        p = np.zeros((self.params.Ncoupled + self.params.Nunc, len(times)))
        q = np.zeros_like(p)
        _lib.check(self._lib.jq_eval_controls(self._handle, pcof.ctypes.data_as(C.c_void_p), len(pcof), len(times),
                                              times.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p)))
        return p, q

    def evaluate_device(self, pcof, shifts=None, weights=None, evaladjoint=True, out=None, stream=None):
        """Batched evaluation on torch CUDA tensors (no host copies, asynchronous on `stream` or torch's current
        stream).  pcof [nbatch, npar], shifts [nsamples, n], weights [nsamples]; returns dict of CUDA tensors."""
        import torch
        self._check_params()

        def ok(t, shape, what):
            if not (t.is_cuda and t.device.index == self.device and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == shape):
                raise ValueError(f"{what} must be a contiguous float64 CUDA tensor of shape {shape} on device {self.device}")
        if pcof.dim() != 2:
            raise ValueError("pcof must be [nbatch, npar]")
        ok(pcof, tuple(pcof.shape), "pcof")
        nbatch, npar = pcof.shape
        nsamples = 1 if shifts is None else shifts.shape[0]
        if shifts is not None:
            ok(shifts, (nsamples, self.params.Ntot), "shifts")
        if weights is not None:
            ok(weights, (nsamples,), "weights")
        shape = (nbatch,) if weights is not None else (nbatch, nsamples)
        if out is None:
            out = {k: torch.empty(shape, dtype=torch.float64, device=pcof.device) for k in ("infid", "leak", "trace_infid")}
            if evaladjoint:
                out["grad"] = torch.empty(shape + (npar,), dtype=torch.float64, device=pcof.device)
                if self.params.objFuncType != 1:
                    out["infidgrad"] = torch.empty_like(out["grad"])
                    out["leakgrad"] = torch.empty_like(out["grad"])
        st = stream if stream is not None else torch.cuda.current_stream(pcof.device)

        def dp(t):
            if t is None:
                return None
            return C.c_void_p(t.data_ptr() if t.numel() else pcof.data_ptr())      # empty sample shard: any valid non-null pointer
        _lib.check(self._lib.jq_traceobjgrad_batch_device(
            self._handle, nbatch, dp(pcof), npar, nsamples, dp(shifts), dp(weights), int(bool(evaladjoint)),
            dp(out["infid"]), dp(out["leak"]), dp(out["trace_infid"]), dp(out.get("grad")), dp(out.get("infidgrad")),
            dp(out.get("leakgrad")), C.c_void_p(st.cuda_stream)))
        return out


def traceobjgrad(pcof0, params: objparams, wa: Working_Arrays, verbose: bool = False, evaladjoint: bool = True):
    """Drop-in for the reference method (src/evalobjgrad.jl:504): same return tuples (:1032-1035)."""
    pcof0 = np.asarray(pcof0, dtype=np.float64)
    if verbose:
        if evaladjoint:
            raise NotImplementedError("verbose=true with evaladjoint=true (forward-sensitivity check of one gradient "
                                      "component) stays on the reference's CPU path (SURVEY.md row 14)")
        # verbose && !evaladjoint: (objfv, unitary history Ntot x N x (nsteps+1), fidelity)  (src/evalobjgrad.jl:1029-1031)
        hist, infid, leak = wa.forward_history(pcof0[None, :])
        return infid[0, 0] + leak[0, 0], hist[0, 0].transpose(2, 1, 0), 1.0 - infid[0, 0]
    r = wa.evaluate(pcof0[None, :], evaladjoint=evaladjoint)
    objfv, primary, secondary = r["objf"][0, 0], r["infid"][0, 0], r["leak"][0, 0]
    if not evaladjoint:
        return objfv, primary, secondary
    totalgrad = r["grad"][0, 0]
    if params.objFuncType != 1:
        infidelgrad, leakgrad = r["infidgrad"][0, 0], r["leakgrad"][0, 0]
    else:
        infidelgrad, leakgrad = totalgrad, np.zeros(0)
    return objfv, totalgrad, primary, secondary, r["trace_infid"][0, 0], infidelgrad, leakgrad


def eval_forward(pcof0, params: objparams, wa: Working_Arrays, saveEndOnly: bool = True, saveEvery: int = 1):
    """eval_forward(U0 = params.Uinit, pcof0, params; saveEndOnly, saveEvery) (src/evalobjgrad.jl:2727-2873):
    the propagated state Ntot x N (saveEndOnly) or its history Ntot x N x (nsteps/saveEvery + 1), complex."""
    hist, _, _ = wa.forward_history(np.asarray(pcof0, dtype=np.float64)[None, :], save_every=params.nsteps if saveEndOnly else saveEvery)
    h = hist[0, 0].transpose(2, 1, 0)
    return h[:, :, -1] if saveEndOnly else h


def evalctrl(params: objparams, pcof0, td, jFunc: int, wa: Optional[Working_Arrays] = None):
    """pj, qj = evalctrl(params, pcof0, td, func) (src/plotstatectrl.jl:246-276): control function number `jFunc`
    (1-based, as in the reference) on the time grid `td`, in rad/ns.  Evaluated on the GPU; `wa` avoids building a
    temporary handle."""
    if not 1 <= jFunc <= params.Ncoupled:
        raise ValueError(f"jFunc must be in 1..{params.Ncoupled} (uncoupled controls are not built)")
    pcof0 = np.asarray(pcof0, dtype=np.float64)
    own = wa is None
    if own:
        wa = Working_Arrays(params, len(pcof0))
    try:
        p, q = wa.controls(pcof0, td)
    finally:
        if own:
            wa.close()
    return p[jFunc - 1].copy(), q[jFunc - 1].copy()


def traceobjgrad_batch(pcofs, params: objparams, wa: Working_Arrays, nodes=None, weights=None, evaladjoint=True):
    """All (candidate, noise sample) trajectories in one call; nodes are the epsilons of the risk-neutral model."""
    shifts = None if nodes is None else noise_shift(params.Ntot, nodes)
    return wa.evaluate(pcofs, shifts, weights, evaladjoint)


def eval_f_g_grad(pcof, params: objparams, wa: Working_Arrays, nodes=(0.0,), weights=(1.0,), compute_adjoint=True):
    """eval_f_g_grad! (src/ipopt_interface.jl:24-70): the nquad-sample loop becomes one batched call."""
    pcof = np.asarray(pcof, dtype=np.float64)
    nodes, weights = np.atleast_1d(np.asarray(nodes, float)), np.atleast_1d(np.asarray(weights, float))
    r = wa.evaluate(pcof[None, :], noise_shift(params.Ntot, nodes), weights, compute_adjoint)
    params.last_pcof = pcof.copy()
    params.last_infidelity = float(r["infid"][0])
    params.last_leak = float(r["leak"][0])
    if compute_adjoint:
        params.last_infidelity_grad = r["infidgrad"][0].copy()
        params.last_leak_grad = r["leakgrad"][0].copy() if params.objFuncType != 1 else np.zeros(len(pcof))
    params.lastTraceInfidelity = params.last_infidelity
    params.lastLeakIntegral = params.last_leak


def _stale(pcof, params):
    return len(params.last_pcof) != len(pcof) or np.linalg.norm(pcof - params.last_pcof) > 1.0e-15


def _fused(pcof, params, wa, nodes, weights):
    """One jq_eval_f_grad call: cache test, sample loop, weighted sums and Tikhonov behind the C ABI.
    Keeps the scalar last-evaluation fields of `params` that intermediate_par reads (src/ipopt_interface.jl:67-68,212-228)."""
    nodes, weights = np.atleast_1d(np.asarray(nodes, float)), np.atleast_1d(np.asarray(weights, float))
    r = wa.eval_f_grad(pcof, noise_shift(params.Ntot, nodes), weights, params.tik0, params.priorCoeffs if params.usingPriorCoeffs else None)
    params.last_pcof = np.array(pcof, dtype=np.float64)
    params.last_infidelity, params.last_leak = r["infid"], r["leak"]
    params.lastTraceInfidelity, params.lastLeakIntegral = r["infid"], r["leak"]
    return r


def eval_f_par(pcof, params, wa, nodes=(0.0,), weights=(1.0,)):
    """eval_f_par (src/ipopt_interface.jl:77-100).  With a device handle this is ONE fused C-ABI call (jq_eval_f_grad) shared with
    eval_grad_f_par through the handle's cache; a `wa` without `eval_f_grad` (host mocks) takes the reference's own steps."""
    pcof = np.asarray(pcof, dtype=np.float64)
    if hasattr(wa, "eval_f_grad"):
        return _fused(pcof, params, wa, nodes, weights)["f"]
    if _stale(pcof, params):
        eval_f_g_grad(pcof, params, wa, nodes, weights, True)
    f = params.last_infidelity + params.last_leak if params.objFuncType == 1 else params.last_infidelity
    return f + tikhonov_pen(pcof, params)


def eval_g_par(pcof, g, params, wa, nodes=(0.0,), weights=(1.0,)):
    pcof = np.asarray(pcof, dtype=np.float64)
    if hasattr(wa, "eval_f_grad"):
        g[0] = _fused(pcof, params, wa, nodes, weights)["leak"]
        return g[0]
    if _stale(pcof, params):
        eval_f_g_grad(pcof, params, wa, nodes, weights, True)
    g[0] = params.last_leak
    return g[0]


def eval_grad_f_par(pcof, grad_f, params, wa, nodes=(0.0,), weights=(1.0,)):
    pcof = np.asarray(pcof, dtype=np.float64)
    if hasattr(wa, "eval_f_grad"):
        grad_f[:] = _fused(pcof, params, wa, nodes, weights)["grad_f"]
    else:
        if _stale(pcof, params):
            eval_f_g_grad(pcof, params, wa, nodes, weights, True)
        grad_f[:] = params.last_infidelity_grad + tikhonov_grad(pcof, params)
    if params.save_pcof_hist:
        params.pcof_hist.append(pcof.copy())


def eval_jac_g_par(pcof, rows, cols, jac_g, params, wa, nodes=(0.0,), weights=(1.0,)):
    pcof = np.asarray(pcof, dtype=np.float64)
    if jac_g is None:
        if len(rows) > 0:
            rows[:] = 1
            cols[:] = np.arange(1, len(pcof) + 1)
        return
    if hasattr(wa, "eval_f_grad"):
        r = _fused(pcof, params, wa, nodes, weights)
        if r["evaluated"]:
            return                  # the reference returns without filling jac_g when it had to re-evaluate (:169-173)
        jac_g[:] = r["leakgrad"]
        return
    if _stale(pcof, params):
        eval_f_g_grad(pcof, params, wa, nodes, weights, True)
        return                      # the reference returns without filling jac_g in this branch (:169-173)
    jac_g[:] = params.last_leak_grad
